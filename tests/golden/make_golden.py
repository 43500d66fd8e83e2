"""Generates tests/golden/* from the UNMODIFIED reference (oracle/_ref), so the fixtures can travel to machines
where /root/reference does not exist. Run from the repo root:  python tests/golden/make_golden.py

  walk_<name>_seed<k>.npz   full walks of the reference SARW class driven by the keyed Philox stand-in RNG
  find_truth.npz            PeriodicLookup::find truth tables on random + adversarial candidates
  stats_c2_reference_mt.npz RDF (intra/inter), <Ree^2>, <Rg^2>, chain-length moments of 12 reference melts (native mt RNG)
  pe_seed1/                 the four output files + summary log the reference's own exporters write for pe.in
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_bindings as ob   # noqa: E402
import cases                   # noqa: E402


def adversarial_candidates(rng, pts, box, thresh, n):
    """Candidates on / around the threshold sphere of stored points, a few ulps either side, plus box-edge cases."""
    box = np.asarray(box, float)
    idx = rng.integers(0, len(pts), n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    scale = np.ones(n)
    kind = rng.integers(0, 6, n)
    for k, ulps in enumerate([0, 1, -1, 4, -4, 64]):
        scale[kind == k] = 1.0 + ulps * 2.0 ** -52
    c = pts[idx] + thresh * scale[:, None] * d
    axis = rng.integers(0, 3, n); axis_mask = rng.random(n) < 0.3          # exactly axis-aligned offsets
    e = np.zeros((n, 3)); e[np.arange(n), axis] = np.sign(rng.normal(size=n))
    c[axis_mask] = pts[idx[axis_mask]] + thresh * scale[axis_mask, None] * e[axis_mask]
    shift = rng.integers(-2, 3, (n, 3)) * box                               # unwrapped images, like chain coordinates
    return c + shift * (rng.random((n, 1)) < 0.5)


def main():
    ob.build_oracles()
    R = ob.RefLib()
    # ---- walks
    walks = {"pe_in": cases.PE_IN, "small_defaults": cases.SMALL_DEFAULTS, "dense_tiny": cases.DENSE_TINY}
    for name, cfg in walks.items():
        for seed in (1, 2, 3):
            r = R.walk(seed=seed, topology=True, **cases.oracle_kwargs(cfg))
            np.savez_compressed(os.path.join(HERE, f"walk_{name}_seed{seed}.npz"), xyz=r.xyz, chainID=r.chainID, type=r.type,
                                bonds=r.bonds, chainLengths=r.chainLengths, countAfterRound=r.countAfterRound,
                                angles=r.angles, dihedrals=r.dihedrals,
                                meta=np.array([r.nAtoms, r.nBonds, r.target, r.rounds, r.initiated, r.lastProgressed, r.trials, r.seedTrials]))
    # grafted / obstacle variants
    box = (24.0, 24.0, 30.0)
    grafts = cases.make_grafts(12, box, seed=5)
    obst = cases.make_obstacles(50, box, seed=9)
    variants = {
        "graft_file": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="file", grafts=grafts),
        "graft_z": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="z"),
        "graft_zcenter": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="zCenter"),
        "graft_order": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="file",
                            growthOrder="grafted", grafts=grafts),
        "obstacles": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, obstacles=obst),
        "pfp": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, boundary="pfp"),
        "fpp": dict(nChains=24, box=box, minDist=1.0, maxTrials=200, boundary="fpp"),
    }
    for name, cfg in variants.items():
        r = R.walk(seed=11, **cfg)
        extra = {k: np.asarray(v) for k, v in cfg.items() if k in ("grafts", "obstacles")}
        np.savez_compressed(os.path.join(HERE, f"walk_{name}.npz"), xyz=r.xyz, chainID=r.chainID, type=r.type, bonds=r.bonds,
                            chainLengths=r.chainLengths, countAfterRound=r.countAfterRound,
                            meta=np.array([r.nAtoms, r.nBonds, r.target, r.rounds, r.initiated, r.lastProgressed, r.trials, r.seedTrials]), **extra)
    # ---- find truth tables
    rng = np.random.default_rng(123)
    tables = {}
    for ti, (box, thresh, npts) in enumerate([((20.0, 20.0, 20.0), 1.0, 3000), ((10.0, 10.0, 40.0), 1.0, 1500),
                                              ((7.3, 9.1, 5.2), 2.0, 60), ((4.0, 4.0, 4.0), 2.0, 8), ((33.3, 21.7, 50.1), 1.25, 4000)]):
        pts = rng.random((npts, 3)) * np.array(box) + rng.integers(-1, 2, (npts, 3)) * np.array(box)
        lk = R.lookup(box, thresh); lk.add(pts)
        q = np.concatenate([rng.random((4000, 3)) * np.array(box) * 3 - np.array(box), adversarial_candidates(rng, pts, box, thresh, 6000)])
        ig = np.where(rng.random(len(q)) < 0.5, -1, rng.integers(0, npts, len(q))).astype(np.int64)
        hit = lk.find(q, ig); lk.close()
        tables[f"box{ti}"] = np.array(box); tables[f"thresh{ti}"] = np.array(thresh); tables[f"pts{ti}"] = pts
        tables[f"q{ti}"] = q; tables[f"ig{ti}"] = ig; tables[f"hit{ti}"] = hit
    np.savez_compressed(os.path.join(HERE, "find_truth.npz"), n=np.array(5), **tables)
    # ---- structure statistics of the reference with its NATIVE random numbers (mt19937_64 re-seeded per trial), 12 seeds of C2
    import stats
    walks = []
    for seed in range(101, 113):
        r = R.walk(seed=seed, rng="mt", **cases.oracle_kwargs(cases.C2))
        walks.append((r.xyz, r.chainID))
    st = stats.summarize(walks, cases.C2["boxSize"], cases.C2["nChains"])
    np.savez_compressed(os.path.join(HERE, "stats_c2_reference_mt.npz"), **st)
    # ---- the reference's own exporters on pe.in (seed 1)
    out = os.path.join(HERE, "pe_seed1")
    os.makedirs(out, exist_ok=True)
    R.walk(seed=1, polymer="POLYETHYLENE", export_dir=out, **cases.oracle_kwargs(cases.PE_IN))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
