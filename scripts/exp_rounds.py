"""Experiment: where does the per-round time go for C2? (early rounds vs all rounds vs single-pass trials)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sarw_b200, cases

def run(tag, cfg, maxRounds=-1, **kw):
    for seed in (1, 2):
        s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed,
                           **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")}, **kw)
        s.ctx.initiate(); s.ctx.grow(0, 3)          # warm-up launch
        d0 = s.ctx.stats().as_dict()
        st = s.ctx.grow(0, maxRounds)
        d = s.ctx.stats().as_dict()
        import ctypes, numpy as np
        busy = (ctypes.c_ulonglong * 256)()
        s.ctx.lib.sarw_debug_busy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        s.ctx.lib.sarw_debug_busy(s.ctx.h, busy, 256)
        b = np.array(list(busy)[:cfg["nChains"]], dtype=float) / max(1, d["rounds"])
        print("   busy cycles/round per warp: min %.0f median %.0f max %.0f; first 16:" % (b.min(), np.median(b), b.max()), " ".join("%.0f" % x for x in b[:16]))
        s.close()
        r = d["rounds"] - d0["rounds"]
        print(f"{tag:28s} seed {seed}: rounds {r:5d} growMs {d['growMs']-d0['growMs']:8.3f} -> {(d['growMs']-d0['growMs'])*1e3/max(1,r):6.2f} us/round | "
              f"cyc A {(d['cycPropose']-d0['cycPropose'])/max(1,r):7.0f} V {(d['cycVerify']-d0['cycVerify'])/max(1,r):6.0f} C {(d['cycCleanup']-d0['cycCleanup'])/max(1,r):6.0f} "
              f"trials/round {(d['trials']-d0['trials'])/max(1,r):7.1f} dead {d['deadChains']} | sub " + " ".join(f"{(a-b)/max(1,r):.0f}" for a, b in zip(d['cycSub'], d0['cycSub'])), flush=True)

C2 = cases.C2
run("C2 first 400 rounds", C2, 400)
run("C2 16 chains", dict(C2, nChains=16), 400)
