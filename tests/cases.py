"""Shared test configurations (keyword sets understood by OracleLib.walk, RefLib.walk and sarw_b200.Context)."""
import numpy as np

PE_IN = dict(nChains=20, boxSize=(10.0, 10.0, 40.0), minDist=1.0, maxTrials=1000, boundary="ppf")   # test/pe.in as the code reads it
C2 = dict(nChains=100, boxSize=(136.2, 136.2, 136.2), minDist=1.0, maxTrials=1000)                  # BASELINE configs[1]
C3 = dict(nChains=10000, boxSize=(632.2, 632.2, 632.2), minDist=1.0, maxTrials=1000)                # BASELINE configs[2]
SMALL_DEFAULTS = dict(nChains=30, boxSize=(30.0, 30.0, 30.0), minDist=2.0, maxTrials=100)            # the code's default minDist/maxTrials
DENSE_TINY = dict(nChains=40, boxSize=(14.0, 14.0, 14.0), minDist=1.0, maxTrials=200, targetMassDensity=4.0)   # crowded: many same-round conflicts, jams
MID = dict(nChains=400, boxSize=(60.0, 60.0, 60.0), minDist=1.0, maxTrials=300)


def oracle_kwargs(cfg):
    d = dict(cfg)
    d["box"] = d.pop("boxSize")
    return d


def make_grafts(n, box, seed=2024, zval=0.0, min_sep=1.0):
    """BASELINE.md C4 recipe: x,y ~ U(0,L), z fixed, rejected if within min_sep (min-image in x,y) of an earlier point."""
    rng = np.random.default_rng(seed)
    pts = []
    while len(pts) < n:
        p = np.array([rng.uniform(0, box[0]), rng.uniform(0, box[1]), zval])
        ok = True
        for q in pts:
            d = p - q
            d[0] -= box[0] * np.round(d[0] / box[0]); d[1] -= box[1] * np.round(d[1] / box[1])
            if np.dot(d, d) <= min_sep * min_sep:
                ok = False; break
        if ok:
            pts.append(p)
    return np.array(pts)


def make_obstacles(n, box, seed=7):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1, (n, 3)) * np.array(box)


def verify_structure(xyz, chainID, typ, bonds, box, minDist, nChains, check_pairs=True):
    """Brute-force verifier (north_star): bonds 1.54 A, 1-3 distance of the 109.5 degree angle, every non-bonded
    pair farther than minDist under the minimum image in all three directions. Returns a dict of extrema."""
    from scipy.spatial import cKDTree
    xyz = np.asarray(xyz); box = np.asarray(box, float)
    out = {}
    a, b = bonds[:, 0] - 1, bonds[:, 1] - 1
    bl = np.linalg.norm(xyz[a] - xyz[b], axis=1)
    out["bond_min"], out["bond_max"] = bl.min(), bl.max()
    assert np.all(chainID[a] == chainID[b])
    # 1-3 distances: for each chain, consecutive beads in acceptance order
    order = np.argsort(chainID, kind="stable")
    cid = chainID[order]; p = xyz[order]
    same2 = cid[2:] == cid[:-2]
    d13 = np.linalg.norm(p[2:] - p[:-2], axis=1)[same2]
    out["d13_min"], out["d13_max"] = (d13.min(), d13.max()) if d13.size else (np.nan, np.nan)
    if check_pairs:
        w = xyz - np.floor(xyz / box) * box
        w = np.minimum(w, np.nextafter(box, 0))
        tree = cKDTree(w, boxsize=box)
        pairs = tree.query_pairs(minDist * 1.0000001, output_type="ndarray")
        # bonded = consecutive beads of one chain
        pos_in_order = np.empty(len(order), np.int64); pos_in_order[order] = np.arange(len(order))
        i, j = pairs[:, 0], pairs[:, 1]
        bonded = (chainID[i] == chainID[j]) & (np.abs(pos_in_order[i] - pos_in_order[j]) == 1)
        nb = pairs[~bonded]
        if len(nb):
            d = w[nb[:, 0]] - w[nb[:, 1]]
            d -= box * np.round(d / box)
            dist = np.linalg.norm(d, axis=1)
            out["closest_nonbonded"] = dist.min()
            out["violations"] = int(np.sum(dist <= minDist))
        else:
            out["closest_nonbonded"] = np.inf
            out["violations"] = 0
    return out
# 1e5 chains, ~100 beads each, in the box of a 1000-bead melt (same chain-tip density as BASELINE configs[2..4]): the size at
# which several same-round conflicts happen every round (scripts/mr_run.py, profiles/r1_multigpu_scaling.md)
BIG = dict(nChains=100000, boxSize=(1362.107287789021,) * 3, minDist=1.0, maxTrials=1000, targetMassDensity=0.092)


# BASELINE configs[3] (the grafted brush) as it can be run: 1e4 chains in the 632.2 A box of configs[2], graftFraction 0.5 = the first
# 5000 chains grafted on z = 0 from a graft file, growthBias 0,0,1, boundary ppf. BASELINE.md's recipe spaces the graft points 1.0 A
# (= minDist) apart; with that spacing the UNMODIFIED reference fails its own seeding (an earlier chain's second bead lands within
# minDist of a later graft point: initiateChain returns false, sarw.cpp:313-314,353 - checked with oracle/_ref). 2.6 A (> bond +
# minDist) cannot be blocked that way, so that is the spacing used everywhere (tests, bench.py, golden hashes).
C4 = dict(nChains=10000, boxSize=(632.2, 632.2, 632.2), minDist=1.0, maxTrials=1000, boundary="ppf", nGraftChains=5000, graftLoc="file",
          growthBias=(0.0, 0.0, 1.0))
C4_GRAFT_SEP = 2.6


def brush_grafts(n=5000, L=632.2, seed=2024, sep=C4_GRAFT_SEP):
    """x,y ~ U(0,L), z = 0, numpy default_rng(seed); a point within `sep` (min image in x,y) of an earlier point is rejected."""
    from scipy.spatial import cKDTree
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


def walk_digest(xyz, bonds, typ):
    """SHA-256 over the bytes of the walk's three output arrays (float64 xyz, int64 bonds, int32 type), as the golden hashes store it."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(xyz, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(bonds, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(typ, dtype=np.int32).tobytes())
    return h.hexdigest()
