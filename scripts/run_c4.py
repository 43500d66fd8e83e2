"""BASELINE configs[3]: grafted brush, graftFraction 0.5 (5000 of 1e4 chains on z = 0), growthBias 0,0,1, boundary ppf.
Runs it on the GPU, verifies the structure with the brute-force verifier and prints timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.spatial import cKDTree
import sarw_b200, cases

L = 632.2
def brush_grafts(n, seed=2024, sep=1.0):
    """BASELINE.md C4 recipe: x,y ~ U(0,L), z = 0, numpy default_rng(2024), rejected if within 1.0 A (min image) of an earlier point"""
    rng = np.random.default_rng(seed)
    pts = np.zeros((0, 2))
    while len(pts) < n:
        cand = rng.uniform(0, L, (2 * (n - len(pts)) + 16, 2))
        allp = np.concatenate([pts, cand])
        tree = cKDTree(allp, boxsize=(L, L))
        pairs = tree.query_pairs(sep, output_type="ndarray")
        bad = set(pairs[:, 1][pairs[:, 1] >= len(pts)].tolist())          # the later point of each close pair goes
        keep = [i for i in range(len(pts), len(allp)) if i not in bad]
        pts = np.concatenate([pts, allp[keep]])[:n]
    return np.concatenate([pts, np.zeros((n, 1))], axis=1)

bias = tuple(float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,0,1".split(",")))
# NB: BASELINE.md asks for graft points >= 1.0 A apart; with that recipe the UNMODIFIED reference fails its own seeding (an earlier
# chain's second bead lands within minDist of a later graft point, every trial of that chain then collides and
# initiateChain returns false, sarw.cpp:313-314,353; oracle/_ref, the oracle and this GPU path all stop at the same chain:
# 674 atoms for seed 1, 430 for seed 2). Graft points farther apart than bond + minDist = 2.54 A cannot be blocked that way.
sep = float(sys.argv[2]) if len(sys.argv) > 2 else 2.6
grafts = brush_grafts(5000, sep=sep)
cfg = dict(nChains=10000, boxSize=(L, L, L), minDist=1.0, maxTrials=1000, boundary="ppf", nGraftChains=5000, graftLoc="file", growthBias=bias)
for seed in (1, 2):
    s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
    s.readGrafts(grafts)
    t0 = time.time(); st = s.run(); t1 = time.time()
    xyz, cid, typ, bonds = s.ctx.download(); d = s.ctx.stats().as_dict(); s.close()
    print({k: d[k] for k in ('nBeads','target','rounds','initiated','lastProgressed','seedTrials','growMs','seedMs','conflicts')}, flush=True)
    if d['growMs'] <= 0: continue
    v = cases.verify_structure(xyz, cid, typ, bonds, cfg["boxSize"], 1.0, 10000)
    a, b = bonds[:, 0] - 1, bonds[:, 1] - 1
    grown = b >= 2 * 10000
    dz = (xyz[b] - xyz[a])[grown, 2]
    print(f"C4 bias {bias} seed {seed}: beads {d['nBeads']} target {d['target']} rounds {d['rounds']} growMs {d['growMs']:.1f} ({d['nBeads']/d['growMs']*1e3:.3e} beads/s) "
          f"trials/bead {d['trials']/d['nBeads']:.1f} exactEvals {d['queries']:.3e} cyc/round A {d['cycPropose']/d['rounds']:.0f} V {d['cycVerify']/d['rounds']:.0f} C {d['cycCleanup']/d['rounds']:.0f} stageFallbacks {d['stageFallbacks']} arcFallbacks {d['arcFallbacks']} conflicts {d['conflicts']} dead {d['deadChains']} lastProgressed {d['lastProgressed']} | bonds [{v['bond_min']:.6f},{v['bond_max']:.6f}] "
          f"d13 [{v['d13_min']:.6f},{v['d13_max']:.6f}] violations {v['violations']} closest {v['closest_nonbonded']:.5f} z range [{xyz[:,2].min():.3f},{xyz[:,2].max():.3f}] "
          f"min dz of grown bonds {dz.min():.4f} grafted starts on z=0: {np.allclose(xyz[0:10000:2][:5000, 2], 0.0)}", flush=True)
