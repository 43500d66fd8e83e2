"""GPU parity of the chain-growth path (K1-K4,K7) through the C ABI.

Bar (north_star): bit-exact against the oracle's sequential walk for the same Philox key — same accepted trial of
every chain in every round, same atom order, ids, types, bonds, and coordinates equal bit for bit (the oracle is run
with the same portable sin/cos; its libm-trig twin is pinned bit-exactly to the unmodified reference in
tests/test_oracle.py). Against fixtures written by the unmodified reference itself the decisions must be identical
and coordinates equal to 1e-9 A (libm vs portable trig)."""
import os
import numpy as np
import pytest

import cases
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def run_gpu(sarw, cfg, seed, maxRounds=-1, grafts=None, obstacles=None, **extra):
    s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed,
                  **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")}, **extra)
    if obstacles is not None:
        s.addObstacle(obstacles)
    if grafts is not None:
        s.readGrafts(grafts)
    st = s.run(maxRounds)
    xyz, chainID, typ, bonds = s.ctx.download()
    out = dict(st=st.as_dict(), xyz=xyz, chainID=chainID, type=typ, bonds=bonds, chainLengths=s.ctx.chain_lengths(),
               countAfterRound=s.ctx.round_counts(), stats=s.ctx.stats().as_dict())
    s.close()
    return out


def assert_same_walk(g, o, exact=True):
    st = g["st"]
    assert (st["nBeads"], st["nBonds"], st["target"], st["rounds"], st["initiated"], st["lastProgressed"]) == \
        (o.nAtoms, o.nBonds, o.target, o.rounds, o.initiated, o.lastProgressed)
    assert (st["trials"], st["seedTrials"]) == (o.trials, o.seedTrials)
    assert np.array_equal(g["chainID"], o.chainID)
    assert np.array_equal(g["type"], o.type)
    assert np.array_equal(g["bonds"], o.bonds)
    assert np.array_equal(g["chainLengths"], o.chainLengths)
    assert np.array_equal(g["countAfterRound"], o.countAfterRound)
    if exact:
        assert np.array_equal(g["xyz"], o.xyz), float(np.abs(g["xyz"] - o.xyz).max())
    else:
        assert np.abs(g["xyz"] - o.xyz).max() < 1e-9


@pytest.mark.parametrize("name,cfg", [("pe_in", cases.PE_IN), ("small_defaults", cases.SMALL_DEFAULTS), ("dense_tiny", cases.DENSE_TINY)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_walk_bit_exact_vs_oracle_and_reference_fixture(sarw, oracle, name, cfg, seed):
    g = run_gpu(sarw, cfg, seed)
    o = oracle.walk(seed=seed, trig="portable", **cases.oracle_kwargs(cfg))
    assert_same_walk(g, o, exact=True)
    gold = np.load(os.path.join(GOLDEN, f"walk_{name}_seed{seed}.npz"))   # written by the unmodified reference
    assert list(gold["meta"][:4]) == [g["st"]["nBeads"], g["st"]["nBonds"], g["st"]["target"], g["st"]["rounds"]]
    assert np.array_equal(g["chainID"], gold["chainID"]) and np.array_equal(g["type"], gold["type"])
    assert np.array_equal(g["bonds"], gold["bonds"]) and np.array_equal(g["countAfterRound"], gold["countAfterRound"])
    assert np.abs(g["xyz"] - gold["xyz"]).max() < 1e-9


@pytest.mark.parametrize("name", ["graft_file", "graft_z", "graft_zcenter", "graft_order", "obstacles", "pfp", "fpp"])
def test_walk_variants_vs_oracle_and_reference_fixture(sarw, oracle, name):
    gold = np.load(os.path.join(GOLDEN, f"walk_{name}.npz"))
    cfg = dict(nChains=24, boxSize=(24.0, 24.0, 30.0), minDist=1.0, maxTrials=200)
    if name.startswith("graft"):
        cfg.update(boundary="ppf", nGraftChains=12, graftLoc={"graft_file": "file", "graft_z": "z", "graft_zcenter": "zCenter", "graft_order": "file"}[name])
        if name == "graft_order":
            cfg["growthOrder"] = "grafted"
    if name in ("pfp", "fpp"):
        cfg["boundary"] = name
    grafts = gold["grafts"] if "grafts" in gold.files else None
    obstacles = gold["obstacles"] if "obstacles" in gold.files else None
    g = run_gpu(sarw, cfg, 11, grafts=grafts, obstacles=obstacles)
    o = oracle.walk(seed=11, trig="portable", grafts=grafts, obstacles=obstacles, **cases.oracle_kwargs(cfg))
    assert_same_walk(g, o, exact=True)
    assert np.array_equal(g["bonds"], gold["bonds"]) and np.array_equal(g["type"], gold["type"])
    assert np.abs(g["xyz"] - gold["xyz"]).max() < 1e-9


@pytest.mark.parametrize("seed", [1, 4])
def test_walk_c2_full_size_bit_exact(sarw, oracle, seed):
    """BASELINE configs[1]: 100 chains x 1000 beads, 0.92 g/cc, ppp — every bead bit-exact against the oracle."""
    g = run_gpu(sarw, cases.C2, seed)
    o = oracle.walk(seed=seed, trig="portable", **cases.oracle_kwargs(cases.C2))
    assert_same_walk(g, o, exact=True)
    v = cases.verify_structure(g["xyz"], g["chainID"], g["type"], g["bonds"], cases.C2["boxSize"], 1.0, 100)
    assert abs(v["bond_min"] - 1.54) < 1e-5 and abs(v["bond_max"] - 1.54) < 1e-5
    assert abs(v["d13_min"] - 2.515256) < 1e-5 and abs(v["d13_max"] - 2.515256) < 1e-5
    assert v["violations"] == 0 and v["closest_nonbonded"] > 1.0
    assert g["st"]["target"] <= g["st"]["nBeads"] < g["st"]["target"] + 100


def test_walk_mid_size_many_conflicts_bit_exact(sarw, oracle):
    g = run_gpu(sarw, cases.MID, 9)
    o = oracle.walk(seed=9, trig="portable", **cases.oracle_kwargs(cases.MID))
    assert_same_walk(g, o, exact=True)
    assert g["stats"]["conflicts"] > 0   # the same-round conflict path was exercised


@pytest.mark.parametrize("cap", [1, 2])
def test_walk_with_overflowing_cells_bit_exact(sarw, oracle, cap):
    """cellCapacity forced small: most beads live in the overflow list; results must not change."""
    g = run_gpu(sarw, cases.DENSE_TINY, 5, cellCapacity=cap)
    o = oracle.walk(seed=5, trig="portable", **cases.oracle_kwargs(cases.DENSE_TINY))
    assert_same_walk(g, o, exact=True)
    assert g["stats"]["overflowBeads"] > 0


def test_walk_growth_bias_extension_matches_oracle(sarw, oracle):
    cfg = dict(nChains=24, boxSize=(24.0, 24.0, 60.0), minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12,
               graftLoc="file", growthBias=(0.0, 0.0, 1.0), targetMassDensity=0.3)
    grafts = cases.make_grafts(12, cfg["boxSize"], seed=3)
    g = run_gpu(sarw, cfg, 2, grafts=grafts)
    o = oracle.walk(seed=2, trig="portable", grafts=grafts, **cases.oracle_kwargs(cfg))
    assert_same_walk(g, o, exact=True)
    # every bond of a biased chain points up (bias . step >= 0)
    a, b = g["bonds"][:, 0] - 1, g["bonds"][:, 1] - 1
    grown = b >= 2 * 24
    assert np.all((g["xyz"][b] - g["xyz"][a])[grown, 2] >= 0)


def test_round_by_round_api_equals_whole_loop(sarw, oracle):
    """sarw_propagate_round (one SARW::propagateChain per call) gives the same walk as sarw_grow."""
    cfg = cases.PE_IN
    s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=3, maxTrials=cfg["maxTrials"], boundary=cfg["boundary"])
    assert s.initiateChain()
    rounds = 0
    while s.actualCount() < s.targetCount():          # sarw.cpp:67-69
        if not s.propagateChain():
            break
        rounds += 1
    s.terminateChain()
    xyz, chainID, typ, bonds = s.ctx.download()
    s.close()
    o = oracle.walk(seed=3, trig="portable", **cases.oracle_kwargs(cfg))
    assert rounds == o.rounds
    assert np.array_equal(xyz, o.xyz) and np.array_equal(typ, o.type) and np.array_equal(bonds, o.bonds)


def test_partial_growth_and_resume(sarw, oracle):
    """maxRounds stops early; a second sarw_grow continues the same walk (several kernel launches, same result)."""
    cfg = cases.SMALL_DEFAULTS
    s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=6, maxTrials=cfg["maxTrials"], reserveRounds=4)
    s.ctx.initiate()
    st = s.ctx.grow(0, 5)
    assert st.rounds == 5
    o5 = oracle.walk(seed=6, trig="portable", maxRounds=5, **cases.oracle_kwargs(cfg))
    assert st.nBeads == o5.nAtoms
    st = s.ctx.grow(0, -1)           # storage was reserved for 4 rounds only: exercises the re-allocation path
    s.ctx.terminate()
    xyz, chainID, typ, bonds = s.ctx.download()
    s.close()
    o = oracle.walk(seed=6, trig="portable", **cases.oracle_kwargs(cfg))
    assert st.rounds == o.rounds and np.array_equal(xyz, o.xyz) and np.array_equal(bonds, o.bonds) and np.array_equal(typ, o.type)


def test_determinism_same_seed_twice(sarw):
    a = run_gpu(sarw, cases.MID, 21)
    b = run_gpu(sarw, cases.MID, 21)
    assert np.array_equal(a["xyz"], b["xyz"]) and np.array_equal(a["bonds"], b["bonds"])
    c = run_gpu(sarw, cases.MID, 22)
    assert not np.array_equal(a["xyz"][:100], c["xyz"][:100])


def test_jammed_seeding_matches_oracle(sarw, oracle):
    """Too many chains for the box: initiateChain fails at the same chain, growth is skipped (sarw.cpp:66,313-314)."""
    cfg = dict(nChains=400, boxSize=(8.0, 8.0, 8.0), minDist=1.5, maxTrials=20)
    g = run_gpu(sarw, cfg, 1)
    o = oracle.walk(seed=1, trig="portable", **cases.oracle_kwargs(cfg))
    assert o.initiated == 0
    assert (g["st"]["initiated"], g["st"]["nBeads"], g["st"]["rounds"]) == (0, o.nAtoms, 0)
    assert np.array_equal(g["xyz"], o.xyz) and np.array_equal(g["bonds"], o.bonds) and np.array_equal(g["type"], o.type)


def test_c3_full_size_invariants(sarw):
    """BASELINE configs[2] at full size (1e4 x 1000 = 1e7 beads): no CPU oracle in test time, so size-independent
    properties: bead count window, bond/angle geometry, zero minDist violations, topology consistency."""
    g = run_gpu(sarw, cases.C3, 1)
    st = g["st"]
    assert st["initiated"] == 1 and st["target"] == 9999171
    assert st["target"] <= st["nBeads"] < st["target"] + 10000
    assert st["nBonds"] == st["nBeads"] - 10000
    assert g["chainLengths"].sum() == st["nBeads"] and np.array_equal(np.bincount(g["chainID"], minlength=10001)[1:], g["chainLengths"])
    v = cases.verify_structure(g["xyz"], g["chainID"], g["type"], g["bonds"], cases.C3["boxSize"], 1.0, 10000)
    assert abs(v["bond_min"] - 1.54) < 1e-5 and abs(v["bond_max"] - 1.54) < 1e-5
    assert abs(v["d13_min"] - 2.515256) < 1e-5 and abs(v["d13_max"] - 2.515256) < 1e-5
    assert v["violations"] == 0
    assert np.sum(g["type"] == 1) == 2 * 10000


def test_many_conflicts_large_box_bit_exact(sarw, oracle):
    """4000 short chains at full density in a 100 A box: ~10 same-round conflicts per round, most of them far apart, so the
    parallel conflict clean-up (independent chains resolved at once) and the ordered pass both run; result == oracle."""
    cfg = dict(nChains=4000, boxSize=(100.0, 100.0, 100.0), minDist=1.0, maxTrials=300, targetMassDensity=0.92)
    g = run_gpu(sarw, cfg, 17)
    o = oracle.walk(seed=17, trig="portable", **cases.oracle_kwargs(cfg))
    assert_same_walk(g, o, exact=True)
    assert g["stats"]["conflicts"] > 20


@pytest.mark.parametrize("bias", [(0.0, 0.0, 0.0), (0.0, 0.0, 1.0)])
def test_brush_mid_size_walls_and_bias_bit_exact(sarw, oracle, bias):
    """BASELINE configs[3] in small: half of 400 chains grafted on z = 0, boundary ppf, with and without growthBias. Chains pile
    up at the walls, so the wall / bias arcs, dead-chain detection and crowded neighbourhoods are all exercised; == oracle."""
    cfg = dict(nChains=400, boxSize=(60.0, 60.0, 60.0), minDist=1.0, maxTrials=300, boundary="ppf", nGraftChains=200, graftLoc="file",
               growthBias=bias)
    rng = np.random.default_rng(11)
    pts = []
    while len(pts) < 200:      # graft points farther apart than bond + minDist, so no graft can be blocked by an earlier chain's second bead
        p = np.array([rng.uniform(0, 60), rng.uniform(0, 60), 0.0])
        if all(np.hypot(*(((p - q)[:2] + 30) % 60 - 30)) > 2.6 for q in pts):
            pts.append(p)
    grafts = np.array(pts)
    g = run_gpu(sarw, cfg, 23, grafts=grafts)
    o = oracle.walk(seed=23, trig="portable", grafts=grafts, **cases.oracle_kwargs(cfg))
    assert_same_walk(g, o, exact=True)
    assert g["stats"]["deadChains"] > 0
    assert g["xyz"][:, 2].min() >= 0.0 and g["xyz"][:, 2].max() <= 60.0


EDGE_CASES = {
    "one_chain": dict(nChains=1, boxSize=(20.0, 20.0, 20.0), minDist=1.0, maxTrials=100),
    "max_trials_1": dict(nChains=30, boxSize=(20.0, 20.0, 20.0), minDist=1.0, maxTrials=1),            # seeding can never succeed (sarw.cpp:355)
    "max_trials_2": dict(nChains=30, boxSize=(20.0, 20.0, 20.0), minDist=1.0, maxTrials=2),
    "max_trials_31": dict(nChains=60, boxSize=(18.0, 18.0, 18.0), minDist=1.0, maxTrials=31, targetMassDensity=2.0),
    "max_trials_33": dict(nChains=60, boxSize=(18.0, 18.0, 18.0), minDist=1.0, maxTrials=33, targetMassDensity=2.0),
    "max_trials_129": dict(nChains=60, boxSize=(18.0, 18.0, 18.0), minDist=1.0, maxTrials=129, targetMassDensity=2.5),
    "tiny_box_one_cell": dict(nChains=2, boxSize=(4.1, 4.1, 4.1), minDist=2.0, maxTrials=50, targetMassDensity=3.0),   # S = 1 per dimension here, 2 in the reference
    "flat_box": dict(nChains=12, boxSize=(40.0, 30.0, 4.2), minDist=2.0, maxTrials=100, targetMassDensity=1.5),
    "anisotropic": dict(nChains=50, boxSize=(17.3, 41.9, 9.7), minDist=1.3, maxTrials=150, targetMassDensity=1.2),
    "mindist_above_bond": dict(nChains=20, boxSize=(30.0, 30.0, 30.0), minDist=1.6, maxTrials=200, targetMassDensity=0.5),  # the bonded neighbour is inside minDist: ignoreIndex matters
    "mindist_2_4": dict(nChains=20, boxSize=(40.0, 40.0, 40.0), minDist=2.4, maxTrials=200, targetMassDensity=0.3),
    "small_mindist": dict(nChains=40, boxSize=(16.0, 16.0, 16.0), minDist=0.35, maxTrials=60, targetMassDensity=3.0),
    "nothing_to_grow": dict(nChains=50, boxSize=(12.0, 12.0, 12.0), minDist=1.0, maxTrials=100, targetMassDensity=0.05),     # seeds already exceed the target
    "fractional_box": dict(nChains=25, boxSize=(23.7, 19.2, 31.4), minDist=1.0, maxTrials=120),
    "ppf_zcenter_wall": dict(nChains=30, boxSize=(20.0, 20.0, 24.0), minDist=1.0, maxTrials=150, boundary="ppf", nGraftChains=30, graftLoc="zCenter", targetMassDensity=1.4),
    "grafted_order_z": dict(nChains=30, boxSize=(20.0, 20.0, 30.0), minDist=1.0, maxTrials=150, boundary="ppf", nGraftChains=10, graftLoc="z", growthOrder="grafted"),
}


@pytest.mark.parametrize("name", sorted(EDGE_CASES))
def test_edge_cases_bit_exact(sarw, oracle, name):
    cfg = EDGE_CASES[name]
    for seed in (1, 2):
        g = run_gpu(sarw, cfg, seed)
        o = oracle.walk(seed=seed, trig="portable", **cases.oracle_kwargs(cfg))
        assert_same_walk(g, o, exact=True)


def test_obstacle_ignore_quirk_bit_exact(sarw, oracle):
    """sarw.cpp:198,353: seeding ignores the LAST obstacle (ignoreIndex -1 + nObstacles); growth offsets ignoreIndex by nObstacles."""
    cfg = dict(nChains=16, boxSize=(14.0, 14.0, 14.0), minDist=1.0, maxTrials=100, targetMassDensity=1.5)
    obst = cases.make_obstacles(300, cfg["boxSize"], seed=4)
    g = run_gpu(sarw, cfg, 9, obstacles=obst)
    o = oracle.walk(seed=9, trig="portable", obstacles=obst, **cases.oracle_kwargs(cfg))
    assert_same_walk(g, o, exact=True)


def test_concurrent_melts_equal_sequential_melts(sarw):
    """sarw_concurrent_capacity melts grown at the same time (own stream + own host thread each) give, bit for bit, the walks the
    same seeds give one at a time; three waves so blocks of destroyed contexts are reused while other melts are still running."""
    import threading
    import torch
    cfg = dict(nChains=100, boxSize=(40.0, 40.0, 40.0), minDist=1.0, maxTrials=300)
    probe = sarw.Context(seed=1, **cfg)
    M = probe.concurrent_capacity()
    probe.close()
    assert M >= 2, M

    def melt(seed, stream=None):
        c = sarw.Context(seed=seed, **cfg)
        if stream is not None:
            c.set_stream(stream.cuda_stream)
        c.initiate(); st = c.grow(0, -1); c.terminate()
        out = c.download()
        c.close()
        return st.nBeads, out

    want = {seed: melt(seed) for seed in range(50, 50 + 3 * M)}
    streams = [torch.cuda.Stream() for _ in range(M)]
    got, errs = {}, []

    def worker(m):
        try:
            for wave in range(3):
                seed = 50 + wave * M + m
                got[seed] = melt(seed, streams[m])
        except Exception as e:   # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=worker, args=(m,)) for m in range(M)]
    for t in th: t.start()
    for t in th: t.join()
    assert not errs, errs
    for seed, (n, (xyz, cid, typ, bonds)) in want.items():
        gn, (gx, gc, gt, gb) = got[seed]
        assert gn == n and np.array_equal(gx, xyz) and np.array_equal(gc, cid) and np.array_equal(gt, typ) and np.array_equal(gb, bonds), seed


def test_release_cached_memory_and_reuse(sarw):
    """destroyed contexts park their tables; releasing them and creating again still gives the same walk"""
    cfg = cases.SMALL_DEFAULTS
    def one():
        s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=9, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
        s.run(); out = s.ctx.download(); s.close()
        return out
    a = one()
    assert sarw.release_cached_memory(-1) == 0
    b = one()
    c = one()
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x, z)


def test_row_copy_staging_equals_tensor_copy_staging(sarw, tmp_path):
    """SARW_NO_TMAP=1 forces the per-row bulk copies (the fallback for regions that cross the periodic boundary) everywhere: the walk
    must not change. Checked in a child process (the switch is read at sarw_create) for a cluster-kernel and a grid-kernel config."""
    import hashlib
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import sarw_b200, cases\n"
        "for name in ('MID', 'SMALL_DEFAULTS'):\n"
        "    cfg = getattr(cases, name)\n"
        "    s = sarw_b200.SARW(cfg['nChains'], cfg['boxSize'], cfg['minDist'], seed=4, **{k: v for k, v in cfg.items() if k not in ('nChains', 'boxSize', 'minDist')})\n"
        "    s.run(); xyz, cid, typ, bonds = s.ctx.download(); s.close()\n"
        "    print(name, hashlib.sha256(xyz.tobytes() + bonds.tobytes() + typ.tobytes()).hexdigest())\n"
    ) % (os.path.dirname(sarw.__file__), os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ, SARW_NO_TMAP=flag)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1] and outs[0].count("\n") == 2, outs


# ------------------------------------------------------------------ large walks pinned by hashes of ONE oracle walk each
def _golden_big():
    import json
    return json.load(open(os.path.join(GOLDEN, "big_walk_hashes.json")))


@pytest.mark.parametrize("name", [k for k in ("C3", "C4", "BIG") if k in _golden_big()])
def test_large_walk_equals_oracle_hash(sarw, name):
    """BASELINE configs[2] (1e4 x 1000 melt) and configs[3] (1e4-chain brush, graftFraction 0.5, growthBias 0,0,1, ppf) at FULL size, and the
    1e5-chain melt whose rounds have several same-round conflicts each:
    SHA-256 of xyz / bonds / type equals the hash of the oracle's Philox-keyed walk (tests/golden/make_golden_big.py), i.e. every one
    of the ~1e7 beads is bit-identical to what the serial reference algorithm produces with the same random numbers."""
    gold = _golden_big()[name]
    cfg = getattr(cases, name)
    g = run_gpu(sarw, cfg, gold["seed"], grafts=cases.brush_grafts() if name == "C4" else None)
    st = g["st"]
    assert (st["nBeads"], st["nBonds"], st["rounds"], st["trials"], st["seedTrials"], st["initiated"], st["lastProgressed"]) == \
        (gold["nAtoms"], gold["nBonds"], gold["rounds"], gold["trials"], gold["seedTrials"], gold["initiated"], gold["lastProgressed"])
    assert cases.walk_digest(g["xyz"], g["bonds"], g["type"]) == gold["sha256"]
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(g["chainLengths"], dtype=np.int32).tobytes()).hexdigest() == gold["chainLengths_sha256"]
    if name == "C4":
        assert g["xyz"][:, 2].min() >= 0.0 and g["xyz"][:, 2].max() <= cfg["boxSize"][2]


def test_walk_checksum_matches_numpy_restatement_of_oracle_walk(sarw, oracle):
    """sarw_walk_checksum (the device-side identity of a walk, used where a download is impossible) equals the numpy restatement
    evaluated on the ORACLE's walk, for a melt with conflicts and for the jammed-seeding case."""
    import stats
    for cfg, seed in ((cases.MID, 9), (cases.PE_IN, 2), (dict(nChains=400, boxSize=(8.0, 8.0, 8.0), minDist=1.5, maxTrials=20), 1)):
        s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
        s.run()
        got = s.ctx.walk_checksum()
        s.close()
        o = oracle.walk(seed=seed, trig="portable", **cases.oracle_kwargs(cfg))
        want = stats.walk_checksum(o.xyz, o.chainID, o.bonds, o.countAfterRound, cfg["nChains"])
        assert got == want, (got, want)


def test_download_box_sample_and_verifier(sarw):
    """sarw_download_box returns exactly the beads of the full download whose wrapped position lies in the box; the sample verifier
    finds no minDist violation; after sarw_release_lookup the walk can still be read but not queried."""
    import stats
    cfg = cases.MID
    s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=5, maxTrials=cfg["maxTrials"])
    s.run()
    xyz, cid, typ, bonds = s.ctx.download()
    lo, hi = (10.0, 5.0, 20.0), (40.0, 35.0, 55.0)
    bx, ids, prev, bc = s.ctx.download_box(lo, hi, cap=1000)        # cap too small on purpose: the binding retries with the reported count
    L = np.array(cfg["boxSize"]); w = xyz - np.floor(xyz / L) * L
    inside = np.all((w >= np.array(lo)) & (w < np.array(hi)), axis=1)
    assert len(ids) == inside.sum() and len(ids) > 1000
    assert np.allclose(np.sort(bx[:, 0]), np.sort(w[inside, 0]), rtol=0, atol=0)
    assert sorted(np.bincount(bc, minlength=401)) == sorted(np.bincount(cid[inside], minlength=401))
    v = stats.verify_box_sample(bx, ids, prev, cfg["minDist"])
    assert v["violations"] == 0 and v["closest_nonbonded"] > cfg["minDist"] and abs(v["bond_min"] - 1.54) < 1e-5 and abs(v["bond_max"] - 1.54) < 1e-5
    before = s.ctx.walk_checksum()
    s.ctx.release_lookup()
    assert s.ctx.walk_checksum() == before
    with pytest.raises(sarw.SarwError):
        s.ctx.find(xyz[:4])
    s.close()


def test_async_growth_equals_blocking_growth(sarw, oracle):
    """sarw_grow_async / sarw_wait (seeding + loop enqueued back to back, ONE synchronisation) give the blocking calls' walk, also when
    bead storage runs out mid-walk (reserveRounds) and when seeding fails (the kernel itself skips the loop, sarw.cpp:66)."""
    for cfg, extra in ((cases.SMALL_DEFAULTS, dict(reserveRounds=4)), (cases.MID, {}), (dict(nChains=400, boxSize=(8.0, 8.0, 8.0), minDist=1.5, maxTrials=20), {})):
        kw = {k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")}
        c = sarw.Context(nChains=cfg["nChains"], boxSize=cfg["boxSize"], minDist=cfg["minDist"], seed=6, **kw, **extra)
        c.grow_async(0, -1)
        st = c.wait(); c.terminate()
        xyz, cid, typ, bonds = c.download(); c.close()
        o = oracle.walk(seed=6, trig="portable", **cases.oracle_kwargs(cfg))
        assert (st.nBeads, st.rounds, st.trials, st.initiated, st.lastProgressed) == (o.nAtoms, o.rounds, o.trials, o.initiated, o.lastProgressed)
        assert np.array_equal(xyz, o.xyz) and np.array_equal(bonds, o.bonds) and np.array_equal(typ, o.type)
