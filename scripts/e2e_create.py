"""Diagnostics: host time of sarw_create / sarw_set_stream / sync per iteration when contexts run on torch's stream."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sarw_b200, cases
cfg = cases.C2
torch.zeros(1, device="cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
stream = torch.cuda.current_stream() if mode == "default" else torch.cuda.Stream()
print("mode", mode, "stream handle", stream.cuda_stream)
for it in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); c = sarw_b200.Context(seed=100 + it, **cfg)
    t1 = time.perf_counter(); c.set_stream(stream.cuda_stream)
    t2 = time.perf_counter(); torch.cuda.synchronize()
    t3 = time.perf_counter(); c.initiate()
    t4 = time.perf_counter(); st = c.grow(0, -1)
    t5 = time.perf_counter(); c.terminate(); out = c.download()
    t6 = time.perf_counter(); c.close()
    t7 = time.perf_counter(); torch.cuda.synchronize()
    t8 = time.perf_counter()
    print("create %.2f set_stream %.2f sync %.2f initiate %.2f grow %.2f download %.2f close %.2f sync %.2f" % tuple(1e3 * (b - a) for a, b in zip((t0, t1, t2, t3, t4, t5, t6, t7), (t1, t2, t3, t4, t5, t6, t7, t8))), flush=True)
