"""CPU tests: the oracle (oracle/sarw_oracle.c) against Random123 known answers, the committed golden fixtures made
from the unmodified reference (tests/golden/make_golden.py), and — where oracle/_ref is built — the reference live."""
import glob
import math
import os
import numpy as np
import pytest

import cases
from conftest import GOLDEN


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_portable_trig_close_to_libm(oracle):
    xs = np.concatenate([np.linspace(0, 359 * math.pi / 180, 20001), np.random.default_rng(0).random(20000) * 6.2657])
    worst = 0.0
    for x in xs:
        s, c = oracle.sincos_portable(float(x))
        for got, want in ((s, math.sin(x)), (c, math.cos(x))):
            ulp = math.ulp(want) if want != 0 else 2.0 ** -1074
            worst = max(worst, abs(got - want) / max(ulp, 2.0 ** -60))   # absolute floor near zeros of sin/cos
    assert worst <= 2.0, worst


def test_target_count(oracle):
    assert oracle.target_count((10, 10, 40), 0.92) == 158            # pe.in
    assert oracle.target_count((136.2, 136.2, 136.2), 0.92) == 99984
    assert oracle.target_count((632.2, 632.2, 632.2), 0.92) == 9999171


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name,cfg", [("pe_in", cases.PE_IN), ("small_defaults", cases.SMALL_DEFAULTS), ("dense_tiny", cases.DENSE_TINY)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_walk_matches_golden_reference(oracle, name, cfg, seed):
    """libm-trig oracle == unmodified reference, bit for bit (fixtures were written by the reference)."""
    gold = _load(f"walk_{name}_seed{seed}.npz")
    o = oracle.walk(seed=seed, trig="libm", **cases.oracle_kwargs(cfg))
    meta = gold["meta"]
    assert [o.nAtoms, o.nBonds, o.target, o.rounds, o.initiated, o.lastProgressed, o.trials, o.seedTrials] == list(meta)
    assert np.array_equal(o.xyz, gold["xyz"])
    assert np.array_equal(o.chainID, gold["chainID"]) and np.array_equal(o.type, gold["type"])
    assert np.array_equal(o.bonds, gold["bonds"]) and np.array_equal(o.chainLengths, gold["chainLengths"])
    assert np.array_equal(o.countAfterRound, gold["countAfterRound"])
    # portable trig: same decisions, coordinates equal to rounding
    p = oracle.walk(seed=seed, trig="portable", **cases.oracle_kwargs(cfg))
    assert p.nAtoms == o.nAtoms and p.trials == o.trials and np.array_equal(p.bonds, o.bonds)
    assert np.abs(p.xyz - o.xyz).max() < 1e-9


@pytest.mark.parametrize("name", ["graft_file", "graft_z", "graft_zcenter", "graft_order", "obstacles", "pfp", "fpp"])
def test_oracle_variants_match_golden_reference(oracle, name):
    gold = _load(f"walk_{name}.npz")
    box = (24.0, 24.0, 30.0)
    cfg = dict(nChains=24, box=box, minDist=1.0, maxTrials=200)
    if name.startswith("graft"):
        cfg.update(boundary="ppf", nGraftChains=12, graftLoc={"graft_file": "file", "graft_z": "z", "graft_zcenter": "zCenter", "graft_order": "file"}[name])
        if name == "graft_order":
            cfg["growthOrder"] = "grafted"
    if name in ("pfp", "fpp"):
        cfg["boundary"] = name
    if "grafts" in gold.files:
        cfg["grafts"] = gold["grafts"]
    if "obstacles" in gold.files:
        cfg["obstacles"] = gold["obstacles"]
    o = oracle.walk(seed=11, trig="libm", **cfg)
    assert [o.nAtoms, o.nBonds, o.target, o.rounds, o.initiated, o.lastProgressed, o.trials, o.seedTrials] == list(gold["meta"])
    assert np.array_equal(o.xyz, gold["xyz"]) and np.array_equal(o.bonds, gold["bonds"])
    assert np.array_equal(o.chainID, gold["chainID"]) and np.array_equal(o.type, gold["type"])


def test_oracle_find_matches_golden_reference(oracle):
    gold = _load("find_truth.npz")
    for ti in range(int(gold["n"])):
        lk = oracle.lookup(tuple(gold[f"box{ti}"]), float(gold[f"thresh{ti}"]))
        lk.add(gold[f"pts{ti}"])
        got = lk.find(gold[f"q{ti}"], gold[f"ig{ti}"])
        assert np.array_equal(got, gold[f"hit{ti}"]), ti
        # the cell table is a pure accelerator: same answers as all-pairs with the same per-pair arithmetic
        brute = lk.find(gold[f"q{ti}"], gold[f"ig{ti}"], bruteforce=True)
        assert np.array_equal(brute, gold[f"hit{ti}"]), ti
        lk.close()


def _sparse_tables():
    """the 1e6 adversarial candidates of tests/golden/make_golden_find.py, regenerated from their seeds and checked against the
    committed SHA-256 (a numpy whose generator streams differ would fail here, loudly, not silently test something else)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_find", os.path.join(GOLDEN, "make_golden_find.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    gold = np.load(os.path.join(GOLDEN, "find_truth_sparse.npz"))
    for ti in range(int(gold["n"])):
        box, thresh, _, n = m.TABLES[ti]
        pts = m.points(ti); q, ig = m.candidates(ti)
        assert m.digest(q, ig) == str(gold[f"sha{ti}"]), "candidate generator drifted: rerun tests/golden/make_golden_find.py where /root/reference exists"
        yield ti, box, thresh, pts, q, ig, np.unpackbits(gold[f"hit{ti}"])[:n]


def test_oracle_find_matches_reference_on_sparse_adversarial_tables(oracle):
    """1e6 candidates ON the threshold sphere (0/1/4/64 ulp either side) of isolated points: the near-threshold pair decides, so a
    table that is half hits / half misses pins the `dist <= thresh` arithmetic of PeriodicLookup.h:67-71."""
    total = 0
    for ti, box, thresh, pts, q, ig, want in _sparse_tables():
        assert 0.3 < want.mean() < 0.7, (ti, want.mean())
        lk = oracle.lookup(box, thresh); lk.add(pts)
        got = lk.find(q, ig); lk.close()
        assert np.array_equal(got, want), (ti, int(np.sum(got != want)))
        total += len(q)
    assert total >= 1_000_000


def test_oracle_walk_matches_live_reference_c2(oracle, ref):
    """Full BASELINE config 2 (100 x 1000 melt): oracle == unmodified reference, bit for bit."""
    r = ref.walk(seed=5, **cases.oracle_kwargs(cases.C2))
    o = oracle.walk(seed=5, trig="libm", **cases.oracle_kwargs(cases.C2))
    assert (r.nAtoms, r.rounds, r.trials, r.seedTrials) == (o.nAtoms, o.rounds, o.trials, o.seedTrials)
    assert np.array_equal(r.xyz, o.xyz) and np.array_equal(r.bonds, o.bonds) and np.array_equal(r.type, o.type)
    assert np.array_equal(r.countAfterRound, o.countAfterRound)


def test_golden_structure_is_valid():
    """The brute-force verifier on the reference's own pe.in output (documents what 'valid' means)."""
    g = _load("walk_pe_in_seed2.npz")
    v = cases.verify_structure(g["xyz"], g["chainID"], g["type"], g["bonds"], cases.PE_IN["boxSize"], 1.0, 20)
    assert abs(v["bond_min"] - 1.54) < 1e-9 and abs(v["bond_max"] - 1.54) < 1e-9
    assert abs(v["d13_min"] - 2.515256) < 1e-5 and abs(v["d13_max"] - 2.515256) < 1e-5
    assert v["violations"] == 0
    # sarw.cpp:351-352,365: CH3 ends
    assert g["type"][0] == 1 and g["type"][1] == 2
    assert len(g["angles"]) == len(g["xyz"]) - 2 * 20 and len(g["dihedrals"]) == len(g["xyz"]) - 3 * 20


# ------------------------------------------------------------------ analysis row: scripts/rdf.py
def test_rdf_oracle_matches_reference_rdf_py_counts(oracle):
    """tests/golden/rdf_reference.npz was written by the reference's own scripts/rdf.py (rdfPair / rdfChains, imported) on walks of
    the unmodified reference: the oracle's integer pair counts and bin edges must be identical."""
    gold = np.load(os.path.join(GOLDEN, "rdf_reference.npz"))
    for k in range(int(gold["n"])):
        w = np.load(os.path.join(GOLDEN, str(gold[f"file{k}"])))
        edges = gold[f"edges{k}"]; nBins = len(edges) - 1
        rMin, dr = float(gold[f"rMin{k}"]), float(gold[f"dr{k}"])
        assert np.array_equal(oracle.rdf_edges(rMin, dr, nBins), edges)
        intra, inter = oracle.rdf(tuple(gold[f"box{k}"]), w["xyz"], w["chainID"], rMin, dr, nBins)
        assert np.array_equal(intra.astype(np.int64), gold[f"intra{k}"]), k
        assert np.array_equal(inter.astype(np.int64), gold[f"inter{k}"]), k


def test_chain_shapes_oracle_matches_numpy(oracle):
    import stats
    w = np.load(os.path.join(GOLDEN, "walk_small_defaults_seed1.npz"))
    nChains = int(w["chainID"].max())
    ree2, rg2 = oracle.chain_shapes(w["xyz"], w["chainID"], nChains)
    ree2n, rg2n, _ = stats.chain_shapes(w["xyz"], w["chainID"], nChains)
    assert np.allclose(ree2, ree2n, rtol=1e-12, atol=0) and np.allclose(rg2, rg2n, rtol=1e-10, atol=0)


def test_rdf_oracle_counts_every_ordered_pair_once(oracle):
    """Properties of the rdf.py accumulation the device kernel is held to: with rMin = 0 and a range beyond |L/2| every ordered pair
    (self pairs included) lands in exactly one bin, and inter-chain counts are even (each unordered pair is counted twice)."""
    w = np.load(os.path.join(GOLDEN, "walk_dense_tiny_seed1.npz"))
    box = (14.0, 14.0, 14.0)
    n = len(w["xyz"])
    nBins = int(np.ceil(np.sqrt(3) * 7.0 / 0.25)) + 2
    intra, inter = oracle.rdf(box, w["xyz"], w["chainID"], 0.0, 0.25, nBins)
    assert int(intra.sum() + inter.sum()) == n * n
    assert int(intra[0]) >= n                      # the self pairs sit in the first bin
    assert (inter % 2 == 0).all()
    lengths = np.bincount(w["chainID"])[1:]
    assert int(intra.sum()) == int((lengths.astype(np.int64) ** 2).sum())
