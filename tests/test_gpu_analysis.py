"""GPU parity of the analysis row: sarw_rdf / sarw_chain_shapes through the C ABI against the oracle's
restatement of the reference's scripts/rdf.py (pinned to rdf.py itself by tests/golden/rdf_reference.npz) — integer pair counts
must be identical, chain shapes bit-exact (same summation order)."""
import os
import numpy as np
import pytest

import cases
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def grow(sarw, cfg, seed):
    s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed,
                  **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
    s.run()
    return s


@pytest.mark.parametrize("name,rMin,dr,nBins", [("PE_IN", 2.0, 0.5, 38), ("PE_IN", 0.0, 0.3, 20), ("SMALL_DEFAULTS", 1.0, 0.7, 35),
                                               ("DENSE_TINY", 2.0, 0.5, 20), ("MID", 2.0, 0.5, 24), ("MID", 0.25, 0.11, 400)])
def test_rdf_counts_match_oracle(sarw, oracle, name, rMin, dr, nBins):
    """Same walk (GPU walk downloaded), same bins: counts identical, including the all-pairs range (PE_IN 38 bins reach |L/2|),
    self pairs (rMin = 0), a cut-off well inside the box (MID: tile culling active) and many narrow bins."""
    cfg = getattr(cases, name)
    s = grow(sarw, cfg, seed=3)
    xyz, cid, typ, bonds = s.ctx.download()
    gi, ge = s.ctx.rdf(rMin, dr, nBins)
    s.close()
    oi, oe = oracle.rdf(cfg["boxSize"], xyz, cid, rMin, dr, nBins)
    assert np.array_equal(gi, oi), (np.flatnonzero(gi != oi)[:5], gi[:5], oi[:5])
    assert np.array_equal(ge, oe), (np.flatnonzero(ge != oe)[:5], ge[:5], oe[:5])
    assert gi.sum() > 0 and ge.sum() > 0


def test_rdf_matches_reference_rdf_py_on_the_same_walk(sarw):
    """The pe.in walk of the GPU equals the reference's walk_pe_in fixture decision for decision (test_gpu_walk), so the counts the
    reference's own rdf.py produced for that fixture must come out of sarw_rdf."""
    gold = np.load(os.path.join(GOLDEN, "rdf_reference.npz"))
    assert str(gold["file0"]) == "walk_pe_in_seed1.npz"
    s = grow(sarw, cases.PE_IN, seed=1)
    nBins = len(gold["edges0"]) - 1
    gi, ge = s.ctx.rdf(float(gold["rMin0"]), float(gold["dr0"]), nBins)
    s.close()
    # coordinates differ from the libm-trig reference by < 1e-9 A: counts can only differ where a distance sits within that of an edge
    assert np.abs(gi.astype(np.int64) - gold["intra0"]).sum() <= 2 and np.abs(ge.astype(np.int64) - gold["inter0"]).sum() <= 4
    assert gi.sum() + ge.sum() > 0


@pytest.mark.parametrize("name", ["PE_IN", "SMALL_DEFAULTS", "MID"])
def test_chain_shapes_bit_exact(sarw, oracle, name):
    cfg = getattr(cases, name)
    s = grow(sarw, cfg, seed=2)
    xyz, cid, typ, bonds = s.ctx.download()
    ree2, rg2 = s.ctx.chain_shapes()
    s.close()
    o_ree2, o_rg2 = oracle.chain_shapes(xyz, cid, cfg["nChains"])
    assert np.array_equal(ree2, o_ree2) and np.array_equal(rg2, o_rg2)
    assert (ree2 > 0).all() and (rg2 > 0).all()


def test_rdf_large_melt_cutoff_properties(sarw):
    """C2 melt (1e5 beads), 12 A cut-off: no pair below minDist, the intra-chain first-shell peaks sit at the 1-3 (2.5153 A) and 1-4
    distances, and the inter-chain g(r) tends to 1 (ideal-gas normalisation)."""
    cfg = cases.C2
    s = grow(sarw, cfg, seed=5)
    st = s.ctx.stats()
    gi, ge = s.ctx.rdf(0.0, 0.1, 120)
    s.close()
    N = st.nBeads
    assert gi[0] == N                              # self pairs only in [0, 0.1)
    assert gi[1:10].sum() == 0 and ge[:10].sum() == 0   # nothing at or below minDist = 1.0
    assert gi[15] >= 2 * st.nBonds                 # bonds at 1.54 A, both orders
    assert gi[25] >= 2 * (N - 2 * cfg["nChains"])  # 1-3 pairs at 2.5153 A
    edges = np.arange(121) * 0.1
    shell = 4 / 3 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    V = np.prod(cfg["boxSize"])
    g_inter = ge / (N * (N - 1) / V * shell)
    assert abs(g_inter[100:].mean() - 1) < 0.03
