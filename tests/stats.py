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
