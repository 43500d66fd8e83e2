import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sarw_b200, cases
cfg = cases.C2
torch.zeros(1, device="cuda")
stream = torch.cuda.current_stream()
mode = sys.argv[1] if len(sys.argv) > 1 else "pre"
if mode == "pre":
    ctxs = [sarw_b200.Context(seed=i, device=0, **cfg) for i in range(11)]
    for c in ctxs: c.set_stream(stream.cuda_stream); c.initiate(); c.grow(0, -1); c.terminate()
    torch.cuda.synchronize()
    for c in ctxs: c.close()
for it in range(14):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); c = sarw_b200.Context(seed=100 + it, device=0, **cfg); c.set_stream(stream.cuda_stream); torch.cuda.synchronize(); t1 = time.perf_counter()
    c.initiate(); torch.cuda.synchronize(); t2 = time.perf_counter()
    st = c.grow(0, -1); torch.cuda.synchronize(); t3 = time.perf_counter()
    c.terminate(); out = c.download(); torch.cuda.synchronize(); t4 = time.perf_counter()
    c.close(); torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"{mode} it {it}: create {1e3*(t1-t0):.2f} initiate {1e3*(t2-t1):.2f} grow {1e3*(t3-t2):.2f} (kernel {st.growMs:.2f}) download {1e3*(t4-t3):.2f} destroy {1e3*(t5-t4):.2f}", flush=True)
