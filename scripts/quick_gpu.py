"""Quick timing probe on a GPU box (not a bench line): growth ms, rounds, trials for a few configs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sarw_b200, cases

which = sys.argv[1:] or ["PE_IN", "C2", "C3"]
seeds = (1, 2)
mode = 0
if which and which[0].startswith("mode="):
    mode = int(which[0][5:]); which = which[1:]
if which and which[0].startswith("seeds="):
    seeds = tuple(int(x) for x in which[0][6:].split(",")); which = which[1:]
for name in which:
    cfg = getattr(cases, name)
    for seed in seeds:
        t0 = time.time()
        s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed, launchMode=mode, cellEdge=float(os.environ.get("SARW_CELL_EDGE", "0")),
                           **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
        t1 = time.time()
        st = s.run()
        t2 = time.time()
        xyz, cid, typ, bonds = s.ctx.download()
        t3 = time.time()
        d = s.ctx.stats().as_dict()
        s.close()
        print(name, seed, f"create {t1-t0:.3f}s run {t2-t1:.3f}s download {t3-t2:.3f}s | beads {d['nBeads']} rounds {d['rounds']} "
              f"trials {d['trials']} queries {d['queries']} conflicts {d['conflicts']} ovf {d['overflowBeads']} "
              f"seedMs {d['seedMs']:.3f} growMs {d['growMs']:.3f} launches {d['growLaunches']} "
              f"| {d['nBeads']/ (d['growMs']+d['seedMs']) * 1e3:.3e} beads/s, {d['growMs']*1e3/max(1,d['rounds']):.2f} us/round "
              f"| cyc/round A {d['cycPropose']/max(1,d['rounds']):.0f} V {d['cycVerify']/max(1,d['rounds']):.0f} C {d['cycCleanup']/max(1,d['rounds']):.0f} "
              f"cleanRounds {d['cleanupRounds']} dead {d['deadChains']}", flush=True)
