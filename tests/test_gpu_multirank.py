"""One melt over several ranks (replicated state, partitioned proposals, one all-gather per round).

The protocol and kernels are exercised on ONE GPU by running R contexts ("ranks") side by side in one process and doing the
all-gather with a device concatenation; the result must be bit-identical to the single-context walk for every R. The NCCL
plumbing itself (torch.distributed over NVLink) is exercised by scripts/mr_run.py under torchrun on >= 2 GPUs."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


class CatGather:
    """stands in for torch.distributed when all ranks live in one process: ranks deposit their buffer, the last one concatenates"""
    def __init__(self, n):
        self.n, self.parts = n, {}


def run_ranks(sarw, cfg, seed, R, grafts=None, obstacles=None):
    import torch
    stream = torch.cuda.current_stream()
    ctxs = []
    for r in range(R):
        c = sarw.Context(seed=seed, launchMode=1, **cfg)
        c.set_stream(stream.cuda_stream)
        if obstacles is not None: c.add_obstacles(obstacles)
        if grafts is not None: c.set_grafts(grafts)
        c.initiate()
        ctxs.append(c)
    KB = [c.mr_setup(r, R) for r, c in enumerate(ctxs)]
    K, B = KB[0]
    local = [torch.zeros(K * B, dtype=torch.uint8, device="cuda") for _ in range(R)]
    more = [c.mr_more() for c in ctxs]
    rounds = 0
    while more[0]:
        assert all(m == more[0] for m in more)
        for c, l in zip(ctxs, local):
            c.mr_propose(l.data_ptr())
        allrec = torch.cat(local)                      # what the NCCL all-gather delivers, in rank order
        more = [c.mr_finish(allrec.data_ptr()) for c in ctxs]
        rounds += 1
    outs = []
    for c in ctxs:
        c.mr_commit(); c.terminate()
        xyz, chainID, typ, bonds = c.download()
        outs.append((xyz, chainID, typ, bonds, c.chain_lengths(), c.stats().as_dict()))
        c.close()
    return outs, rounds


def run_single(sarw, cfg, seed, grafts=None, obstacles=None):
    c = sarw.Context(seed=seed, **cfg)
    if obstacles is not None: c.add_obstacles(obstacles)
    if grafts is not None: c.set_grafts(grafts)
    c.initiate(); st = c.grow(0, -1); c.terminate()
    out = c.download() + (c.chain_lengths(), c.stats().as_dict())
    c.close()
    return out


@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("name,cfg", [("mid", cases.MID), ("dense", cases.DENSE_TINY), ("pe_in", cases.PE_IN)])
def test_ranks_give_the_single_gpu_walk_bit_for_bit(sarw, name, cfg, R):
    ref = run_single(sarw, cfg, 13)
    outs, rounds = run_ranks(sarw, cfg, 13, R)
    assert rounds == ref[5]["rounds"]
    for o in outs:                                      # every rank holds the complete, identical result
        for a, b in zip(o[:5], ref[:5]):
            assert np.array_equal(a, b)
        assert (o[5]["nBeads"], o[5]["trials"], o[5]["conflicts"]) == (ref[5]["nBeads"], ref[5]["trials"], ref[5]["conflicts"])


def test_ranks_with_grafts_and_growth_order(sarw):
    cfg = dict(nChains=24, boxSize=(24.0, 24.0, 30.0), minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="file", growthOrder="grafted")
    grafts = cases.make_grafts(12, cfg["boxSize"], seed=5)
    obst = cases.make_obstacles(30, cfg["boxSize"], seed=2)
    ref = run_single(sarw, cfg, 4, grafts=grafts, obstacles=obst)
    outs, _ = run_ranks(sarw, cfg, 4, 2, grafts=grafts, obstacles=obst)
    for o in outs:
        for a, b in zip(o[:5], ref[:5]):
            assert np.array_equal(a, b)


def test_ranks_c2_full_size_matches_oracle(sarw, oracle):
    outs, rounds = run_ranks(sarw, cases.C2, 3, 4)
    o = oracle.walk(seed=3, trig="portable", **cases.oracle_kwargs(cases.C2))
    assert rounds == o.rounds
    xyz, chainID, typ, bonds, lengths, st = outs[2]
    assert np.array_equal(xyz, o.xyz) and np.array_equal(bonds, o.bonds) and np.array_equal(typ, o.type)
    assert st["trials"] == o.trials


# ------------------------------------------------------------------ the peer-memory kernel (sarw_mr_grow_peer)
# The whole loop of every rank runs in ONE persistent launch; records and conflict lists are stored straight into the other
# ranks' buffers and the ranks meet on flags (st.release.sys / ld.acquire.sys). Here the R ranks are R contexts of this process
# on ONE GPU, each on its own stream and host thread, with maxBlocks small enough that all R cooperative grids are resident at
# the same time (sarw_mr_peer_connect_local shares the exchange buffers as plain device pointers instead of CUDA IPC handles).
def run_ranks_peer(sarw, cfg, seed, R, split_verify=True, grafts=None, obstacles=None, maxBlocks=24, slab=False, halo=0, download=True):
    import torch
    streams = [torch.cuda.Stream() for _ in range(R)]
    ctxs = []
    for r in range(R):
        kw = dict(slabRank=r, slabRanks=R, slabHalo=halo) if slab else {}
        c = sarw.Context(seed=seed, launchMode=1 if split_verify else 3, maxBlocks=maxBlocks, **kw, **cfg)
        c.set_stream(streams[r].cuda_stream)
        if obstacles is not None: c.add_obstacles(obstacles)
        if grafts is not None: c.set_grafts(grafts)
        assert c.initiate()
        ctxs.append(c)
    sarw.connect_peers_local(ctxs)
    torch.cuda.synchronize()
    stats = sarw.grow_peer_local(ctxs)
    outs = []
    for c, st in zip(ctxs, stats):
        c.terminate()
        if download:
            xyz, chainID, typ, bonds = c.download()
            outs.append((xyz, chainID, typ, bonds, c.chain_lengths(), c.stats().as_dict()))
        else:
            outs.append((c.walk_checksum(), c.stats().as_dict()))
        c.close()
    return outs


@pytest.mark.parametrize("split_verify", [True, False])
@pytest.mark.parametrize("R", [2, 3, 4])
@pytest.mark.parametrize("name,cfg", [("mid", cases.MID), ("dense", cases.DENSE_TINY)])
def test_peer_kernel_gives_the_single_gpu_walk_bit_for_bit(sarw, name, cfg, R, split_verify):
    ref = run_single(sarw, cfg, 13)
    outs = run_ranks_peer(sarw, cfg, 13, R, split_verify)
    for o in outs:                                      # every rank holds the complete, identical result
        for a, b in zip(o[:5], ref[:5]):
            assert np.array_equal(a, b)
        assert (o[5]["rounds"], o[5]["nBeads"], o[5]["trials"], o[5]["conflicts"]) == (ref[5]["rounds"], ref[5]["nBeads"], ref[5]["trials"], ref[5]["conflicts"])
    if name == "mid":
        assert ref[5]["conflicts"] > 0


@pytest.mark.parametrize("R,split_verify", [(2, True), (4, True), (3, False)])
def test_peer_kernel_c2_full_size_matches_oracle(sarw, oracle, R, split_verify):
    outs = run_ranks_peer(sarw, cases.C2, 3, R, split_verify)
    o = oracle.walk(seed=3, trig="portable", **cases.oracle_kwargs(cases.C2))
    for xyz, chainID, typ, bonds, lengths, st in outs:
        assert st["rounds"] == o.rounds and st["trials"] == o.trials
        assert np.array_equal(xyz, o.xyz) and np.array_equal(bonds, o.bonds) and np.array_equal(typ, o.type)


def test_peer_kernel_many_chains_and_conflicts(sarw, oracle):
    """4000 chains in a 100 A box (about 10 same-round conflicts per round): partitioned verification, conflict-list exchange, the
    parallel and the ordered clean-up all run on every rank; == oracle."""
    cfg = dict(nChains=4000, boxSize=(100.0, 100.0, 100.0), minDist=1.0, maxTrials=300, targetMassDensity=0.92)
    outs = run_ranks_peer(sarw, cfg, 17, 3, True, maxBlocks=64)
    o = oracle.walk(seed=17, trig="portable", **cases.oracle_kwargs(cfg))
    for xyz, chainID, typ, bonds, lengths, st in outs:
        assert st["conflicts"] > 20 and st["trials"] == o.trials
        assert np.array_equal(xyz, o.xyz) and np.array_equal(bonds, o.bonds) and np.array_equal(typ, o.type)


def test_peer_kernel_with_grafts_obstacles_growth_order(sarw):
    cfg = dict(nChains=24, boxSize=(24.0, 24.0, 30.0), minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="file", growthOrder="grafted")
    grafts = cases.make_grafts(12, cfg["boxSize"], seed=5)
    obst = cases.make_obstacles(30, cfg["boxSize"], seed=2)
    ref = run_single(sarw, cfg, 4, grafts=grafts, obstacles=obst)
    for o in run_ranks_peer(sarw, cfg, 4, 2, True, grafts=grafts, obstacles=obst):
        for a, b in zip(o[:5], ref[:5]):
            assert np.array_equal(a, b)


def test_peer_kernel_under_torchrun_two_gpus():
    """the real thing: one process per GPU, CUDA IPC handles, NVLink. Skipped on boxes with fewer than 2 GPUs."""
    import json, os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(root, "scripts", "mr_run.py"), "4000", "100"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["nccl_identical_on_all_ranks"] and line["peer_identical_on_all_ranks"] and line["slab_identical_on_all_ranks"]


# ------------------------------------------------------------------ slab decomposition (sarw_params.slabRanks): the box is split along z
# Rank r holds the cell layers of slab r plus a halo (the rest of the table has no memory behind it), chains belong to the rank whose slab
# holds their tip, same-round conflicts are settled by a fixed-point iteration whose results are exchanged. Same R-contexts-on-one-GPU
# set-up as above; the result must still be the single-GPU walk, bit for bit, on every rank.
@pytest.mark.parametrize("R", [2, 3, 4])
@pytest.mark.parametrize("name,cfg,halo", [("mid", cases.MID, 4), ("mid_wide_halo", cases.MID, 0), ("dense", cases.DENSE_TINY, 4)])
def test_slab_ranks_give_the_single_gpu_walk_bit_for_bit(sarw, name, cfg, halo, R):
    ref = run_single(sarw, cfg, 13)
    outs = run_ranks_peer(sarw, cfg, 13, R, slab=True, halo=halo)
    for o in outs:
        for a, b in zip(o[:5], ref[:5]):
            assert np.array_equal(a, b)
        assert (o[5]["rounds"], o[5]["nBeads"], o[5]["trials"]) == (ref[5]["rounds"], ref[5]["nBeads"], ref[5]["trials"])
    if name == "mid":
        assert ref[5]["conflicts"] > 0 and all(o[5]["conflicts"] == outs[0][5]["conflicts"] for o in outs)


def test_slab_ranks_many_chains_and_conflicts(sarw, oracle):
    """4000 chains in a 100 A box (about 10 same-round conflicts per round, some of them next to each other): the fixed-point clean-up
    and its exchange run on every rank, chains change owner as their tips cross slab boundaries; == oracle."""
    cfg = dict(nChains=4000, boxSize=(100.0, 100.0, 100.0), minDist=1.0, maxTrials=300, targetMassDensity=0.92)
    outs = run_ranks_peer(sarw, cfg, 17, 3, maxBlocks=64, slab=True, halo=4)
    o = oracle.walk(seed=17, trig="portable", **cases.oracle_kwargs(cfg))
    for xyz, chainID, typ, bonds, lengths, st in outs:
        assert st["conflicts"] > 20 and st["trials"] == o.trials and st["rounds"] == o.rounds
        assert np.array_equal(xyz, o.xyz) and np.array_equal(bonds, o.bonds) and np.array_equal(typ, o.type)


def test_slab_ranks_with_grafts_obstacles_growth_order_and_walls(sarw):
    cfg = dict(nChains=24, boxSize=(24.0, 24.0, 30.0), minDist=1.0, maxTrials=200, boundary="ppf", nGraftChains=12, graftLoc="file", growthOrder="grafted")
    grafts = cases.make_grafts(12, cfg["boxSize"], seed=5)
    obst = cases.make_obstacles(30, cfg["boxSize"], seed=2)
    ref = run_single(sarw, cfg, 4, grafts=grafts, obstacles=obst)
    for o in run_ranks_peer(sarw, cfg, 4, 2, grafts=grafts, obstacles=obst, slab=True, halo=4):
        for a, b in zip(o[:5], ref[:5]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("name,R", [("C3", 2), ("BIG", 4)])
def test_slab_ranks_large_walk_equals_oracle_hash(sarw, name, R):
    """BASELINE configs[2] (1e7 beads, 1 GB of cells) over 2 slabs and the 1e5-chain melt (10 GB of cells) over 4: here the windows are real -
    most of the table's address range has no memory behind it, so a stray access would fault - and the walk is still the oracle's."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big_walk_hashes.json")))
    if name not in gold:
        pytest.skip("no golden hash")
    gold = gold[name]
    cfg = getattr(cases, name)
    outs = run_ranks_peer(sarw, cfg, gold["seed"], R, maxBlocks=296 // R, slab=True)
    for xyz, chainID, typ, bonds, lengths, st in outs:
        assert (st["nBeads"], st["rounds"], st["trials"]) == (gold["nAtoms"], gold["rounds"], gold["trials"])
        assert cases.walk_digest(xyz, bonds, typ) == gold["sha256"]


def test_slab_context_refuses_what_it_cannot_do(sarw):
    c = sarw.Context(seed=1, launchMode=1, slabRank=0, slabRanks=2, **cases.C3)
    try:
        with pytest.raises(sarw.SarwError):
            c.find(np.zeros((1, 3)))
        with pytest.raises(sarw.SarwError):
            c.mr_setup(1, 2)          # created as slab 0 of 2
        assert c.initiate()
        with pytest.raises(sarw.SarwError):
            c.grow(0, -1)             # one slab cannot grow alone
    finally:
        c.close()
    with pytest.raises(sarw.SarwError):
        sarw.Context(seed=1, slabRank=2, slabRanks=2, **cases.MID)
