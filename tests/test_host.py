"""CPU tests of the host layer that sits either side of the hot path (SURVEY §8f "next" rows): the InputMap-compatible
parser, the O(N) topology builder and the byte-compatible exporters, through the GPU-free sarw_hosttest binary; and
(GPU) the `sarw -i pe.in` drop-in itself against files written by the reference's own exporters."""
import os
import shutil
import struct
import subprocess
import numpy as np
import pytest

import cases
from conftest import GOLDEN, ROOT

BIN = os.path.join(ROOT, "polymer-sarw_b200", "bin")


@pytest.fixture(scope="module")
def hosttest():
    exe = os.path.join(BIN, "sarw_hosttest")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", ROOT, os.path.relpath(exe, ROOT)], check=True)
    return exe


def write_atoms_bin(path, nChains, box, xyz, chainID, typ, bonds, chainLengths):
    with open(path, "wb") as f:
        f.write(struct.pack("<qqq", nChains, len(chainID), len(bonds)))
        f.write(np.asarray(box, "<f8").tobytes())
        f.write(np.ascontiguousarray(xyz, "<f8").tobytes())
        f.write(np.ascontiguousarray(chainID, "<i4").tobytes())
        f.write(np.ascontiguousarray(typ, "<i4").tobytes())
        f.write(np.ascontiguousarray(bonds, "<i8").tobytes())
        f.write(np.ascontiguousarray(chainLengths, "<i4").tobytes())


def test_input_map_semantics(hosttest, tmp_path, monkeypatch):
    """InputMap.cpp:31-43: first two tokens per trimmed line, rest ignored, later keys override, ${VAR} substituted,
    unknown keys kept; numeric values through atof, vectors comma-separated."""
    monkeypatch.setenv("SARW_TEST_BOX", "12.5,7,9")
    p = tmp_path / "x.in"
    p.write_text("  nChains   7   # inline comment\n\nminDist\t1.25 trailing junk\nboxSize ${SARW_TEST_BOX}\nnChains 9\nlonely\nmaxTrials 5e2\nweird key value\n")
    out = subprocess.run([hosttest, "parse", str(p)], capture_output=True, text=True, check=True).stdout.splitlines()
    kv = dict(l.split("=", 1) for l in out if not l.startswith("#"))
    assert kv == {"nChains": "9", "minDist": "1.25", "boxSize": "12.5,7,9", "maxTrials": "5e2", "weird": "key"}
    assert out[-1] == "#boxSize=12.5,7,9 nChains=9 minDist=1.25 maxTrials=500"
    r = subprocess.run([hosttest, "parse", str(tmp_path / "missing.in")], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open system file" in r.stdout          # InputMap.cpp:30


def test_exporters_byte_identical_to_reference(hosttest, tmp_path):
    """Atoms of the reference's own pe.in run -> our topology + exporters -> the very bytes its exporters wrote."""
    g = np.load(os.path.join(GOLDEN, "walk_pe_in_seed1.npz"))
    write_atoms_bin(tmp_path / "a.bin", 20, cases.PE_IN["boxSize"], g["xyz"], g["chainID"], g["type"], g["bonds"], g["chainLengths"])
    for threads in ("1", "3"):
        out = tmp_path / f"out{threads}"; out.mkdir()
        r = subprocess.run([hosttest, "export", str(tmp_path / "a.bin"), str(out), "POLYETHYLENE", threads], capture_output=True, text=True, check=True)
        n = [int(x) for x in r.stdout.split()]
        assert n == [len(g["chainID"]), len(g["bonds"]), len(g["angles"]), len(g["dihedrals"])]
        for name in ("polymer.dat", "polymer.xyz", "polymer.xsf", "chains.dat"):
            assert (out / name).read_bytes() == open(os.path.join(GOLDEN, "pe_seed1", name), "rb").read(), name


@pytest.mark.parametrize("name", ["walk_small_defaults_seed2.npz", "walk_dense_tiny_seed3.npz", "walk_pe_in_seed3.npz"])
def test_topology_matches_reference(hosttest, tmp_path, name):
    """O(N) builder == the reference's O(nChains*N) scan (sarw.cpp:397-422): same angles and dihedrals, same order."""
    g = np.load(os.path.join(GOLDEN, name))
    nChains = len(g["chainLengths"])
    write_atoms_bin(tmp_path / "a.bin", nChains, (30, 30, 30), g["xyz"], g["chainID"], g["type"], g["bonds"], g["chainLengths"])
    subprocess.run([hosttest, "export", str(tmp_path / "a.bin"), str(tmp_path)], check=True, capture_output=True)
    lines = (tmp_path / "polymer.dat").read_text().split("\n")
    ia = lines.index("Angles") + 2
    ang = np.array([[int(x) for x in l.split()] for l in lines[ia:ia + len(g["angles"])]], dtype=np.int64).reshape(-1, 5)
    idh = lines.index("Dihedrals") + 2
    dih = np.array([[int(x) for x in l.split()] for l in lines[idh:idh + len(g["dihedrals"])]], dtype=np.int64).reshape(-1, 6)
    assert np.array_equal(ang[:, 2:], g["angles"]) and np.array_equal(dih[:, 2:], g["dihedrals"])
    assert np.array_equal(ang[:, 0], np.arange(1, len(ang) + 1)) and np.all(ang[:, 1] == 1)


def test_topology_large_random_is_linear_time(hosttest, tmp_path):
    """1e6 synthetic atoms in 1000 chains: finishes in seconds (the reference scan would be ~1e9 steps) and is consistent."""
    nC, rounds = 1000, 998
    chainID = np.concatenate([np.repeat(np.arange(1, nC + 1), 2), np.tile(np.arange(1, nC + 1), rounds)]).astype(np.int32)
    N = len(chainID)
    xyz = np.zeros((N, 3)); typ = np.full(N, 2, np.int32)
    bonds = np.zeros((N - nC, 2), np.int64)
    write_atoms_bin(tmp_path / "a.bin", nC, (100, 100, 100), xyz, chainID, typ, bonds, np.full(nC, rounds + 2, np.int32))
    r = subprocess.run([hosttest, "export", str(tmp_path / "a.bin"), str(tmp_path)], check=True, capture_output=True, text=True, timeout=120)
    n = [int(x) for x in r.stdout.split()]
    assert n == [N, N - nC, N - 2 * nC, N - 3 * nC]


@pytest.mark.gpu
def test_cli_drop_in_pe_in(tmp_path):
    """`sarw -i pe.in` (test/pe.in as shipped) on the GPU: files byte-identical to what the reference's own exporters
    wrote for the same Philox key; log blocks as the reference prints them."""
    exe = os.path.join(BIN, "sarw")
    assert os.path.exists(exe), "build the CLI first (make cli)"
    shutil.copy(os.path.join(GOLDEN, "pe.in"), tmp_path / "pe.in")
    r = subprocess.run([exe, "-i", "pe.in", "--seed", "1"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("polymer.dat", "polymer.xyz", "polymer.xsf", "chains.dat"):
        assert (tmp_path / name).read_bytes() == open(os.path.join(GOLDEN, "pe_seed1", name), "rb").read(), name
    for alias, name in (("polymer.data", "polymer.dat"), ("polymer.XYZ", "polymer.xyz"), ("polymer.XSF", "polymer.xsf")):
        assert (tmp_path / alias).read_bytes() == (tmp_path / name).read_bytes()
    log = r.stdout
    for block in ("\nINPUTS:\nnChains = 20\nboxSize = (10 x 10 x 40)\nminDist = 1\ntargetMassDensity = 0.92\nnGraftChains = 0\ngraftLoc = file\n"
                  "growthOrder = roundRobin\nboundary = ppf\nmaxTrials = 1000\npolymer = POLYETHYLENE\nobstacle = none\n",
                  "\nLOG FLAGS:\nlogProgress = yes\nlogSteps = no\n", "\nPROGRESS: "):
        assert block in log, block
    ref_summary = open(os.path.join(GOLDEN, "pe_seed1", "ref.log")).read()
    assert ref_summary.strip() in log                      # SUMMARY + "Exporting ..." lines, verbatim
    r2 = subprocess.run([exe, "-i", "nope.in"], cwd=tmp_path, capture_output=True, text=True)
    assert r2.returncode == 1 and "Could not open system file 'nope.in' for reading." in r2.stdout


@pytest.mark.gpu
def test_cli_grafted_brush_with_files(tmp_path):
    """grafts.dat + obstacle file through the CLI loaders (sarw.cpp:102-145) == the same walk through the ABI."""
    import sarw_b200
    exe = os.path.join(BIN, "sarw")
    box = (24.0, 24.0, 30.0)
    grafts = cases.make_grafts(14, box, seed=8)          # two extra records: "Removing extra seeds!"
    obst = cases.make_obstacles(20, box, seed=3)
    np.savetxt(tmp_path / "grafts.dat", grafts, fmt="%.17g")
    np.savetxt(tmp_path / "obs.dat", obst, fmt="%.17g")
    (tmp_path / "b.in").write_text("nChains 24\nboxSize 24,24,30\nminDist 1.0\nmaxTrials 200\nboundary ppf\nnGraftChains 12\ngraftLoc file\n"
                                   "obstacle obs.dat\nseed 77\nlogProgress no\n")
    r = subprocess.run([exe, "-i", "b.in", "-o", "b.log", "--timing"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    log = (tmp_path / "b.log").read_text()
    assert "Reading grafts.dat: 14 graft seeds found\nRemoving extra seeds!" in log and "Reading obs.dat: 20 obstacles found" in log
    s = sarw_b200.SARW(24, box, 1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="file", seed=77)
    s.addObstacle(obst); s.readGrafts(grafts[:12]); st = s.run()
    xyz, chainID, typ, bonds = s.ctx.download(); s.close()
    rows = (tmp_path / "polymer.xyz").read_text().split("\n")
    assert int(rows[0]) == st.nBeads
    got = np.array([[float(x) for x in l.split()[1:]] for l in rows[2:2 + st.nBeads]])
    assert np.abs(got - xyz).max() < 1e-6
    assert np.allclose(got[0:24:2], grafts[:12], atol=1e-6)   # grafted chains start on their graft points


@pytest.mark.gpu
def test_cli_rdf_option_writes_rdf_py_counts(tmp_path):
    """`sarw -i pe.in --rdf 2,0.5,38` adds rdf.dat / shapes.dat without touching the reference's output files; the pair counts are
    the ones the reference's own scripts/rdf.py gave for the same walk (tests/golden/rdf_reference.npz, case 0)."""
    exe = os.path.join(BIN, "sarw")
    shutil.copy(os.path.join(GOLDEN, "pe.in"), tmp_path / "pe.in")
    r = subprocess.run([exe, "-i", "pe.in", "--seed", "1", "--rdf", "2,0.5,38"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("polymer.dat", "polymer.xyz", "polymer.xsf", "chains.dat"):
        assert (tmp_path / name).read_bytes() == open(os.path.join(GOLDEN, "pe_seed1", name), "rb").read(), name
    rows = np.loadtxt(tmp_path / "rdf.dat")
    gold = np.load(os.path.join(GOLDEN, "rdf_reference.npz"))
    assert rows.shape == (38, 6)
    assert np.abs(rows[:, 2] - gold["intra0"]).sum() <= 2 and np.abs(rows[:, 3] - gold["inter0"]).sum() <= 4
    shapes = np.loadtxt(tmp_path / "shapes.dat")
    lengths = np.loadtxt(tmp_path / "chains.dat")
    assert shapes.shape == (20, 4) and np.array_equal(shapes[:, 1], lengths) and (shapes[:, 2] > 0).all()


def test_fast_format_equals_printf(hosttest):
    """The exporters' own number formatting (host/fast_format.h) against snprintf("%lf") / ("%lld") on 3e5 x 5 pseudo-random doubles
    (coordinates, negatives, tiny values, dyadic rationals that hit exact ties at six decimals) plus edge cases: zero mismatches."""
    r = subprocess.run([hosttest, "fmtcheck", "300000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    bad, total = (int(x) for x in r.stdout.split())
    assert bad == 0 and total > 1_500_000


def test_exporters_large_multithreaded_equals_single_thread(hosttest, tmp_path):
    """1.1e6 beads: the parallel topology emission (N >= 2^20) and the pipelined multi-batch writer (more chunks than threads) must
    write exactly the bytes the single-threaded path writes."""
    import hashlib
    nC, per = 1100, 1000
    N = nC * per
    rng = np.random.default_rng(3)
    chainID = np.concatenate([np.repeat(np.arange(1, nC + 1), 2), np.tile(np.arange(1, nC + 1), per - 2)]).astype(np.int32)
    xyz = (rng.random((N, 3)) - 0.25) * 300.0                     # negative coordinates too
    xyz[::1000] = np.round(xyz[::1000] * 128) / 128 + 1.0 / 256   # exact ties at the sixth decimal
    typ = np.where(rng.random(N) < 0.01, 1, 2).astype(np.int32)
    bonds = np.zeros((N - nC, 2), np.int64)
    last = 2 * np.arange(1, nC + 1)
    bonds[:nC, 0] = last - 1; bonds[:nC, 1] = last
    b = nC
    for r in range(per - 2):
        ids = 2 * nC + r * nC + np.arange(1, nC + 1)
        bonds[b:b + nC, 0] = last; bonds[b:b + nC, 1] = ids; last = ids; b += nC
    write_atoms_bin(tmp_path / "a.bin", nC, (300, 300, 300), xyz, chainID, typ, bonds, np.full(nC, per, np.int32))
    digests = []
    for threads in ("1", "4"):
        out = tmp_path / f"o{threads}"; out.mkdir()
        r = subprocess.run([hosttest, "export", str(tmp_path / "a.bin"), str(out), "Polyethylene", threads], capture_output=True, text=True, check=True, timeout=300)
        assert [int(x) for x in r.stdout.split()] == [N, N - nC, N - 2 * nC, N - 3 * nC]
        digests.append([hashlib.md5((out / f).read_bytes()).hexdigest() for f in ("polymer.dat", "polymer.xyz", "polymer.xsf", "chains.dat")])
        if threads == "4":   # the text itself against Python's exactly rounded "%f" (same rule as glibc's printf), ties included
            rows = (out / "polymer.xyz").read_text().split("\n")[2:2 + 60000]
            want = ["C\t%f\t%f\t%f" % tuple(p) for p in xyz[:60000]]
            assert rows == want
        for f in ("polymer.dat", "polymer.xyz", "polymer.xsf"):
            (out / f).unlink()
    assert digests[0] == digests[1]


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 3])
def test_cli_gpus_decomposes_the_box_and_writes_the_same_files(tmp_path, gpus):
    """`sarw -i in --gpus N`: one melt over N slab contexts (on a box with fewer devices they share a GPU); the four output files are
    byte-identical to the single-GPU run's, for test/pe.in (walls, checked against the reference's own files) and for a melt with
    same-round conflicts, grafts and obstacles."""
    exe = os.path.join(BIN, "sarw")
    shutil.copy(os.path.join(GOLDEN, "pe.in"), tmp_path / "pe.in")
    r = subprocess.run([exe, "-i", "pe.in", "--seed", "1", "--gpus", str(gpus), "--slab-halo", "4"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "one melt over %d GPUs" % gpus in r.stdout
    for name in ("polymer.dat", "polymer.xyz", "polymer.xsf", "chains.dat"):
        assert (tmp_path / name).read_bytes() == open(os.path.join(GOLDEN, "pe_seed1", name), "rb").read(), name
    box = (60.0, 60.0, 60.0)
    np.savetxt(tmp_path / "grafts.dat", cases.make_grafts(40, box, seed=8), fmt="%.17g")
    np.savetxt(tmp_path / "obs.dat", cases.make_obstacles(50, box, seed=3), fmt="%.17g")
    (tmp_path / "m.in").write_text("nChains 400\nboxSize 60,60,60\nminDist 1.0\nmaxTrials 300\nnGraftChains 40\ngraftLoc file\nobstacle obs.dat\nseed 5\nlogProgress no\n")
    outs = []
    for extra in ([], ["--gpus", str(gpus), "--slab-halo", "4"]):
        d = tmp_path / ("run%d" % len(outs)); d.mkdir()
        for f in ("m.in", "grafts.dat", "obs.dat"): shutil.copy(tmp_path / f, d / f)
        r = subprocess.run([exe, "-i", "m.in"] + extra, cwd=d, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append({n: (d / n).read_bytes() for n in ("polymer.dat", "polymer.xyz", "polymer.xsf", "chains.dat")})
    assert outs[0] == outs[1]
