"""Where does the end-to-end time of one C2 melt through the ABI go? (create / initiate / grow / download / destroy)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, sarw_b200, cases
cfg = cases.C2
torch.zeros(1, device="cuda")
stream = torch.cuda.current_stream()
cap = 101000
pinned = {"xyz": torch.empty((cap, 3), dtype=torch.float64, pin_memory=True).numpy(), "chainID": torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
          "type": torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(), "bonds": torch.empty((cap, 2), dtype=torch.int64, pin_memory=True).numpy()}
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for setstream in (False, True):
    for pin in (False, True):
        for fl in (False, True):
            acc = {}
            def tick(name, t0):
                torch.cuda.synchronize(); t = time.perf_counter(); acc[name] = acc.get(name, 0) + (t - t0); return t
            for it in range(12):
                if it == 2: acc = {}
                if fl: flush.fill_(float(it)); torch.cuda.synchronize()
                t = time.perf_counter()
                c = sarw_b200.Context(seed=100 + it, **cfg)
                if setstream: c.set_stream(stream.cuda_stream)
                t = tick("create", t)
                c.initiate(); t = tick("initiate", t)
                st = c.grow(0, -1); t = tick("grow", t)
                c.terminate(); out = c.download(pinned if pin else None); t = tick("download", t)
                c.close(); t = tick("destroy", t)
            print(f"set_stream={setstream} pinned={pin} l2flush={fl}:", {k: round(v / 10 * 1e3, 3) for k, v in acc.items()}, "total", round(sum(acc.values()) / 10 * 1e3, 2), "ms; kernel", round(st.growMs, 3), flush=True)
