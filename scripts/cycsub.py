"""Per-phase SM cycles of warp 0 / block 0 of the growth kernel (sarw_stats.cycSub) for a workload; diagnostics."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sarw_b200, cases
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = getattr(cases, name)
for seed in (1, 2, 3):
    s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
    s.run(); d = s.ctx.stats().as_dict()
    import ctypes as C
    pp = (C.c_ulonglong * 8)(); s.ctx.lib.sarw_debug_paths(s.ctx.h, pp); pp = list(pp)
    s.close()
    r = max(1, d["rounds"])
    names = ["V:scan", "V:insert", "V:barrier2", "A:propose", "A:publish", "A:shadow", "A:barrier1", "-"] if cfg["nChains"] <= 256 else ["frame", "stage", "arcs", "classify", "evaluate", "publish", "barrier", "-"]
    sub = d["cycSub"]
    print(name, seed, f"growMs {d['growMs']:.3f} rounds {r} A {d['cycPropose']/r:.0f} V {d['cycVerify']/r:.0f} C {d['cycCleanup']/r:.0f} | " +
          " ".join(f"{n} {v/r:.0f}" for n, v in zip(names, sub)) + " | paths " + " ".join(f"{n}: {pp[4+k]/max(1,d['rounds']):.1f}/round x {pp[k]/max(1,pp[4+k]):.0f} cyc" for k, n in enumerate(["dead", "fast", "fast->hard", "hard"])) + f" | trials/bead {d['trials']/d['nBeads']:.1f} dead {d['deadChains']} arcFallbacks {d['arcFallbacks']} stageFallbacks {d['stageFallbacks']}")
