"""Timing of the device analysis (sarw_rdf / sarw_chain_shapes) on finished melts; wall clock around the ABI call (gather + sort +
pair kernel + copy back). Not a bench line."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, sarw_b200, cases

for name in sys.argv[1:] or ["C2", "C3"]:
    cfg = getattr(cases, name)
    s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=1, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
    st = s.run()
    N = s.ctx.stats().nBeads
    L = cfg["boxSize"][0]
    jobs = [("12 A cut-off, dr 0.5", 2.0, 0.5, 20), ("20 A cut-off, dr 0.1", 0.0, 0.1, 200)]
    if N <= 200000:
        jobs.append(("all pairs to |L/2| (rdf.py's range)", 2.0, 0.5, int((np.sqrt(3) * L / 2 - 2.0) / 0.5)))
    for label, rMin, dr, nb in jobs:
        s.ctx.rdf(rMin, dr, nb)
        t0 = time.perf_counter(); gi, ge = s.ctx.rdf(rMin, dr, nb); t = time.perf_counter() - t0
        pairs = int(gi.sum() + ge.sum())
        print(f"{name} N={N} rdf {label}: {t*1e3:.1f} ms, {pairs:.3e} pairs binned, {pairs/t:.3e} binned pairs/s", flush=True)
    t0 = time.perf_counter(); ree2, rg2 = s.ctx.chain_shapes(); t = time.perf_counter() - t0
    print(f"{name} chain_shapes: {t*1e3:.2f} ms  <Ree^2> {ree2.mean():.1f} <Rg^2> {rg2.mean():.1f} ratio {ree2.mean()/rg2.mean():.2f}", flush=True)
    if N <= 200000:
        import stats
        xyz, cid, _, _ = s.ctx.download()
        t0 = time.perf_counter(); stats.rdf_intra_inter(xyz, cid, cfg["boxSize"]); t = time.perf_counter() - t0
        print(f"{name} CPU KD-tree RDF (tests/stats.py, 12 A): {t*1e3:.0f} ms")
    s.close()
