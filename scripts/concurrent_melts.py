"""Throughput of M independent C2 melts grown concurrently on ONE GPU (one host thread + one CUDA stream per melt):
a 100-chain melt occupies one 16-CTA cluster, i.e. ~1/9 of the SMs."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sarw_b200, cases
cfg = cases.C2
torch.zeros(1, device="cuda")
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
def warm():
    c = sarw_b200.Context(seed=1, **cfg); c.initiate(); c.grow(0, -1); c.terminate(); c.close()
warm(); warm()
_c = sarw_b200.Context(seed=1, **cfg); print('concurrent capacity:', _c.concurrent_capacity()); _c.close()
for M in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4,8,9,12,16").split(",")]:
    streams = [torch.cuda.Stream() for _ in range(M)]
    beads = [0] * M; kern = [0.0] * M; ph = [0.0] * 4
    def work(t):
        for k in range(K):
            ta = time.perf_counter()
            c = sarw_b200.Context(seed=1000 * t + k, **cfg)
            c.set_stream(streams[t].cuda_stream)
            tb = time.perf_counter()
            c.initiate()
            tc = time.perf_counter()
            st = c.grow(0, -1)
            td = time.perf_counter()
            c.terminate()
            beads[t] += st.nBeads; kern[t] += st.growMs
            c.close()
            te = time.perf_counter()
            ph[0] += tb - ta; ph[1] += tc - tb; ph[2] += td - tc; ph[3] += te - td
    work_threads = [threading.Thread(target=work, args=(t,)) for t in range(M)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for th in work_threads: th.start()
    for th in work_threads: th.join()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"M={M:2d} concurrent melts x {K}: {sum(beads)/dt:.3e} beads/s wall, {dt/K*1e3:.2f} ms per wave, kernel {sum(kern)/(M*K):.2f} ms per melt; host ms per melt: create {ph[0]/(M*K)*1e3:.2f} initiate {ph[1]/(M*K)*1e3:.2f} grow {ph[2]/(M*K)*1e3:.2f} close {ph[3]/(M*K)*1e3:.2f}", flush=True)
