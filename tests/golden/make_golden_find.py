"""Adversarial truth tables for PeriodicLookup::find (PeriodicLookup.h:51-75), written by the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden_find.py        ->  tests/golden/find_truth_sparse.npz

1.0e6 candidates, every one of them placed ON the threshold sphere of ONE stored point, 0 / 1 / 4 / 64 ulp inside or outside
(plus exactly axis-aligned offsets and unwrapped periodic images). The point sets are SPARSE (jittered lattices whose spacing keeps
every other point more than 2.5 thresh away from the probe), so the near-threshold pair alone decides each answer and an error in
the `dist <= thresh` comparison (PeriodicLookup.h:67-71) cannot hide behind another hit. The candidates are a pure function of
(table, seed) — `candidates()` below regenerates them in the tests — so only the packed answers and a SHA-256 of the candidate
bytes are committed (an npz of 1e6 candidates would be 24 MB).
"""
import hashlib
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "find_truth_sparse.npz")

# (box, thresh, lattice counts per dimension (spacing = box / counts), candidates)
TABLES = [
    ((90.0, 90.0, 90.0), 1.0, (20, 20, 20), 300_000),           # many cells, cubic
    ((33.75, 22.5, 50.625), 1.25, (6, 4, 9), 250_000),          # anisotropic box
    ((10.0, 10.0, 40.0), 1.0, (2, 2, 8), 250_000),              # test/pe.in's box
    ((4.0, 4.0, 4.0), 2.0, (1, 1, 1), 100_000),                 # S = 2 in every dimension: the point's own images are neighbours
    ((7.3, 9.1, 5.2), 2.0, (1, 1, 1), 100_000),                 # S = (3, 4, 2)
]
ULPS = np.array([0, 1, -1, 4, -4, 64, -64], dtype=np.float64)


def points(ti):
    box, thresh, counts, _ = TABLES[ti]
    rng = np.random.default_rng(1000 + ti)
    box = np.array(box); counts = np.array(counts)
    g = np.stack(np.meshgrid(*[np.arange(c) for c in counts], indexing="ij"), -1).reshape(-1, 3)
    spacing = box / counts
    jitter = (rng.random(g.shape) - 0.5) * np.minimum(thresh, spacing - 3.6 * thresh).clip(0)   # keeps neighbours >= 3.5 thresh apart where the lattice allows
    pts = (g + 0.5) * spacing + jitter
    return pts + rng.integers(-1, 2, pts.shape) * box       # stored unwrapped, like chain coordinates


def candidates(ti):
    """(q, ignore): the candidates of table ti and their ignoreIndex (-1 = none)"""
    box, thresh, _, n = TABLES[ti]
    rng = np.random.default_rng(2000 + ti)
    box = np.array(box)
    pts = points(ti)
    idx = rng.integers(0, len(pts), n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    axis = rng.integers(0, 3, n); aligned = rng.random(n) < 0.25
    e = np.zeros((n, 3)); e[np.arange(n), axis] = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    d[aligned] = e[aligned]
    scale = 1.0 + ULPS[rng.integers(0, len(ULPS), n)] * 2.0 ** -52
    q = pts[idx] + (thresh * scale)[:, None] * d
    q = q + rng.integers(-2, 3, (n, 3)) * box * (rng.random((n, 1)) < 0.5)
    u = rng.random(n)
    ig = np.where(u < 0.12, idx, np.where(u < 0.24, rng.integers(0, len(pts), n), -1)).astype(np.int64)
    return np.ascontiguousarray(q), ig


def digest(q, ig):
    h = hashlib.sha256(); h.update(q.tobytes()); h.update(ig.tobytes())
    return h.hexdigest()


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle_bindings as ob
    ob.build_oracles()
    R = ob.RefLib()
    out = {"n": np.array(len(TABLES))}
    for ti, (box, thresh, _, n) in enumerate(TABLES):
        pts = points(ti)
        q, ig = candidates(ti)
        lk = R.lookup(box, thresh); lk.add(pts)
        hit = lk.find(q, ig); lk.close()
        decided = hit[ig == -1]
        print(f"table {ti}: box {box} thresh {thresh} points {len(pts)} candidates {n} hit fraction {hit.mean():.4f} (no ignoreIndex: {decided.mean():.4f})")
        out[f"hit{ti}"] = np.packbits(hit.astype(np.uint8))
        out[f"sha{ti}"] = np.array(digest(q, ig))
    np.savez_compressed(OUT, **out)
    print("written", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
