"""One sarw_rdf call on a finished C2 melt (all pairs to |L/2|, rdf.py's range) — target for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, sarw_b200, cases
cfg = cases.C2
s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=1, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
s.run()
nb = int((np.sqrt(3) * cfg["boxSize"][0] / 2 - 2.0) / 0.5)
gi, ge = s.ctx.rdf(2.0, 0.5, nb)
print("pairs binned", int(gi.sum() + ge.sum()))
s.close()
