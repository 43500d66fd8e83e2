"""Structure statistics for the statistical-parity gate (north_star correctness part 3), restating the math of the
reference's scripts/rdf.py without its hard-coded sizes:

  rdf.py:13-20  rdfPair    minimum-image pair-distance histogram, dr = 0.5, rMin = 2
  rdf.py:39-57  rdfChains  intra-chain / inter-chain split (inter pairs counted twice)
  rdf.py:23-36  rdfUniform normalisation by the pair histogram of a uniform distribution (here: the analytic
                           ideal-gas shell count N*(N-1)/V * 4 pi r^2 dr instead of a sampled grid)

plus end-to-end distance and radius of gyration per chain from the UNWRAPPED coordinates (sarw.cpp:184 never wraps).
Pairs are found with a periodic KD-tree up to rCut (the structure is in the first few shells; rdf.py goes to |L/2|)."""
import numpy as np
from scipy.spatial import cKDTree


def rdf_intra_inter(xyz, chainID, box, rMin=2.0, rCut=12.0, dr=0.5):
    box = np.asarray(box, float)
    w = xyz - np.floor(xyz / box) * box
    w = np.minimum(w, np.nextafter(box, 0))
    tree = cKDTree(w, boxsize=box)
    pairs = tree.query_pairs(rCut, output_type="ndarray")
    d = w[pairs[:, 0]] - w[pairs[:, 1]]
    d -= box * np.floor(0.5 + d / box)                   # rdf.py:15, minimum image
    r = np.sqrt((d * d).sum(1))
    edges = np.arange(rMin, rCut + 1e-9, dr)              # rdf.py:19
    same = chainID[pairs[:, 0]] == chainID[pairs[:, 1]]
    hIntra = np.histogram(r[same], edges)[0] * 2.0       # ordered pairs, like rdf.py's full N x N histogram
    hInter = np.histogram(r[~same], edges)[0] * 2.0
    N = len(xyz); V = box.prod()
    shell = 4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    ideal = N * (N - 1) / V * shell                       # what rdfUniform * N^2 estimates
    return hIntra / ideal, hInter / ideal, 0.5 * (edges[1:] + edges[:-1])


def chain_shapes(xyz, chainID, nChains):
    """end-to-end distance^2 and radius of gyration^2 of every chain (unwrapped coordinates, acceptance order)."""
    order = np.argsort(chainID, kind="stable")
    cid = chainID[order]; p = xyz[order]
    starts = np.searchsorted(cid, np.arange(1, nChains + 1)); ends = np.searchsorted(cid, np.arange(1, nChains + 1), side="right")
    ree2 = np.zeros(nChains); rg2 = np.zeros(nChains); length = ends - starts
    for c in range(nChains):
        q = p[starts[c]:ends[c]]
        if len(q) == 0:
            continue
        ree2[c] = ((q[-1] - q[0]) ** 2).sum()
        rg2[c] = ((q - q.mean(0)) ** 2).sum(1).mean()
    return ree2, rg2, length


def summarize(walks, box, nChains):
    """walks: list of (xyz, chainID). Returns per-seed arrays of the statistics the parity gate compares."""
    out = dict(gIntra=[], gInter=[], ree2=[], rg2=[], cn=[], meanLen=[], stdLen=[])
    for xyz, chainID in walks:
        gi, ge, r = rdf_intra_inter(xyz, chainID, box)
        ree2, rg2, length = chain_shapes(xyz, chainID, nChains)
        out["gIntra"].append(gi); out["gInter"].append(ge); out["r"] = r
        out["ree2"].append(ree2.mean()); out["rg2"].append(rg2.mean())
        out["cn"].append((ree2 / np.maximum(length - 1, 1) / 1.54 ** 2).mean())   # characteristic ratio <R^2>/(n l^2)
        out["meanLen"].append(length.mean()); out["stdLen"].append(length.std())
    return {k: np.array(v) for k, v in out.items()}


# ------------------------------------------------------------------ checks for walks too large to download
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def walk_checksum(xyz, chainID, bonds, countAfterRound, nChains):
    """numpy restatement of sarw_walk_checksum (include/sarw_abi.h): the four words from a walk in acceptance order, as the oracle or
    sarw_download returns it. Slot of an atom: seeds 2*chain + {0,1}; the bead chain c accepted in round r: 2n + r*n + c."""
    with np.errstate(over="ignore"):
        N = len(xyz); n = nChains
        chain = chainID.astype(np.int64) - 1
        nSeed = int(2 * n if N >= 2 * n else N)                        # all chains seeded (or seeding stopped early: only seed atoms exist)
        slot = np.empty(N, np.int64)
        # seed atoms come chain by chain, two each (sarw.cpp:301-310)
        slot[:nSeed] = 2 * chain[:nSeed] + (np.arange(nSeed) & 1)
        if N > nSeed:
            counts = np.asarray(countAfterRound, np.int64)               # atoms after each round
            rnd = np.searchsorted(counts, np.arange(nSeed, N), side="right")
            slot[nSeed:] = 2 * n + rnd * n + chain[nSeed:]
        bits = np.ascontiguousarray(xyz, np.float64).view(np.uint64).reshape(-1, 3)
        h = _mix64(slot.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        for k in range(3):
            h = _mix64(h ^ bits[:, k])
        prev = np.full(N, -1, np.int64)
        if len(bonds):
            prev[bonds[:, 1] - 1] = slot[bonds[:, 0] - 1]                  # Bond(1, pen+1, last+1), sarw.cpp:272,309
        pv = np.where(prev < 0, _M64, prev.astype(np.uint64))
        b = _mix64(h ^ pv)
        return (int(np.add.reduce(h)), int(np.bitwise_xor.reduce(h)) if N else 0, N, int(np.add.reduce(b)))


def verify_box_sample(xyz, ids, prevIds, minDist):
    """Brute-force verifier on a sample returned by sarw_download_box (wrapped coordinates inside one box, no periodic images needed):
    bond lengths of the bonds with both ends in the sample, closest non-bonded pair, pairs at or below minDist."""
    order = np.argsort(ids); xyz, ids, prevIds = xyz[order], ids[order], prevIds[order]
    pos = np.searchsorted(ids, prevIds)
    pos_c = np.minimum(pos, len(ids) - 1)
    have = (prevIds >= 0) & (ids[pos_c] == prevIds)
    bl = np.linalg.norm(xyz[have] - xyz[pos_c[have]], axis=1)
    # a bond whose ends are on opposite sides of the periodic box appears long in wrapped coordinates: only bonds shorter than 2 A count
    bl = bl[bl < 2.0]
    tree = cKDTree(xyz)
    pairs = tree.query_pairs(minDist * 1.0000001, output_type="ndarray")
    i, j = pairs[:, 0], pairs[:, 1]
    bonded = (prevIds[i] == ids[j]) | (prevIds[j] == ids[i])
    nb = pairs[~bonded]
    d = np.linalg.norm(xyz[nb[:, 0]] - xyz[nb[:, 1]], axis=1) if len(nb) else np.array([np.inf])
    return dict(beads=len(ids), bonds=int(have.sum()), bond_min=float(bl.min()) if len(bl) else np.nan, bond_max=float(bl.max()) if len(bl) else np.nan,
                closest_nonbonded=float(d.min()), violations=int((d <= minDist).sum()))
