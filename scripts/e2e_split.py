"""Where does the end-to-end time of one C2 melt through the ABI go? (create / initiate / grow / download / destroy)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sarw_b200, cases
cfg = cases.C2
torch.zeros(1, device="cuda")
acc = {}
def tick(name, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); acc[name] = acc.get(name, 0) + (t - t0); return t
for it in range(12):
    if it == 2: acc = {}
    t = time.perf_counter()
    c = sarw_b200.Context(seed=100 + it, **cfg); t = tick("create", t)
    c.initiate(); t = tick("initiate", t)
    st = c.grow(0, -1); t = tick("grow", t)
    c.terminate(); out = c.download(); t = tick("download", t)
    c.close(); t = tick("destroy", t)
print({k: round(v / 10 * 1e3, 3) for k, v in acc.items()}, "ms per melt; kernel", round(st.growMs, 3))
