"""world_size-2 gloo test (CPU) of bench.py's N>1 plumbing: ranks shard the work as independent melts with distinct
Philox keys; time is the MAX over ranks, work the SUM; the reference arm prints on rank 0 only."""
import json
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import bench
    dist.init_process_group("gloo")
    r = dist.get_rank()
    seeds = bench.rank_seeds(r, 5)
    t_dev, t_e2e, beads, e2e_beads = bench.aggregate_ranks(dist, "cpu", 1.0 + r, 2.0 - r, 100 * (r + 1), 10 * (r + 1))
    gathered = [None, None]
    dist.all_gather_object(gathered, seeds)
    if r == 0:
        print(json.dumps(dict(t_dev=t_dev, t_e2e=t_e2e, beads=beads, e2e_beads=e2e_beads, seeds=gathered)))
    dist.barrier(); dist.destroy_process_group()
""") % ROOT


def test_two_rank_aggregation_with_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(w)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["t_dev"] == 2.0 and d["t_e2e"] == 2.0            # max over ranks
    assert d["beads"] == 300.0 and d["e2e_beads"] == 30.0     # sum over ranks
    s0, s1 = d["seeds"]
    assert len(set(s0) | set(s1)) == 10                       # no two (rank, step) share a key


def test_reference_arm_prints_only_on_rank0(tmp_path):
    """--impl reference under a multi-rank launch: rank != 0 exits 0 silently (checked without starting the reference)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


WORKER2 = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "polymer-sarw_b200"))
    import torch, torch.distributed as dist
    import sarw_b200
    dist.init_process_group("gloo")
    r, R = dist.get_rank(), dist.get_world_size()
    n = 11                                   # chains; not a multiple of the rank count
    K = (n + R - 1) // R
    # each rank writes the chain index into the 32-byte records of the chains it proposes for (chain i -> rank i %% R, slot i // R)
    local = torch.full((K, 4), -1.0, dtype=torch.float64)
    for i in range(n):
        if i %% R == r:
            local[i // R, 0] = float(i)
    allrec = torch.empty((R * K, 4), dtype=torch.float64)
    dist.all_gather_into_tensor(allrec, local)
    got = [int(allrec[sarw_b200.mr_chain_slot(i, R, K), 0].item()) for i in range(n)]
    if r == 0:
        print(json.dumps(got))
    dist.barrier(); dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_record_layout_after_all_gather_with_gloo(tmp_path):
    """The layout contract of include/sarw_abi.h "multi-rank": after a rank-ordered all-gather chain i's record sits at
    (i % nranks) * K + i // nranks — checked with a real 2-process gloo all_gather_into_tensor."""
    w = tmp_path / "worker2.py"
    w.write_text(WORKER2)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29618", str(w)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert got == list(range(11))


def test_committed_bench_line_has_every_contract_key():
    """profiles/r1_bench_final.json is a line bench.py printed on a B200: the keys the measurement contract names are all there."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, "profiles", "r1_bench_final.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["metric"] == "accepted beads/sec" and line["unit"] == "beads/s" and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert "workload" in line["config"] and "model" not in line["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in line["e2e"], k
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in line["roofline"], k
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in line["cpu_baseline"], k
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in line["clocks"], k
    assert line["gpu_launches"] > 0 and line["e2e"]["value"] != line["value"]
    ref = json.loads(open(os.path.join(root, "profiles", "r1_bench_reference_arm.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "reference"


WORKER3 = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import bench
    dist.init_process_group("gloo")
    r = dist.get_rank()
    out = {}
    # nobody failed: nobody raises
    bench.all_ranks_ok(dist, None, "step 1", device="cpu"); out["clean"] = True
    # rank 1 fails locally (an allocation, a launch): BOTH ranks raise, at the same point, so the next collective still matches
    try:
        bench.all_ranks_ok(dist, MemoryError("no room") if r == 1 else None, f"rank {r}: step 2", device="cpu")
        out["raised"] = False
    except RuntimeError as e:
        out["raised"] = True; out["text"] = str(e)
    t = torch.tensor([r + 1]); dist.all_reduce(t)         # the collective after the failure: must not hang or mismatch
    out["sum"] = int(t.item())
    gathered = [None, None]
    dist.all_gather_object(gathered, out)
    if r == 0: print(json.dumps(gathered))
    dist.barrier(); dist.destroy_process_group()
""") % ROOT


def test_a_local_failure_is_raised_on_every_rank_with_gloo(tmp_path):
    """bench.py's multi-GPU records: a rank that fails on its own (round 2: rank 0 could not allocate its slab of the 1e9-bead melt while
    the other seven went on into the next all-gather and waited ten minutes for it) must take all ranks out of the record together."""
    w = tmp_path / "worker3.py"
    w.write_text(WORKER3)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29619", str(w)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    a, b = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert a["clean"] and b["clean"] and a["raised"] and b["raised"] and a["sum"] == b["sum"] == 3
    assert "failed on another rank" in a["text"] and "no room" in b["text"]


WORKER4 = textwrap.dedent("""
    import os, sys, json, time
    sys.path.insert(0, %r)
    import bench
    r, port = int(sys.argv[1]), int(sys.argv[2])
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    ag = bench.HostAgreement(r, 2, port=port, timeout_s=30)
    out = {}
    ag.agree(None, "step 1"); out["clean"] = not ag.poisoned
    try:
        ag.agree(MemoryError("no room") if r == 1 else None, "step 2")          # a failure that leaves the device usable
        out["soft"] = "not raised"
    except RuntimeError as e:
        out["soft"] = str(e); out["poisoned_after_soft"] = ag.poisoned
    try:
        ag.agree(RuntimeError("illegal memory access") if r == 0 else None, "step 3", hard=(r == 0))   # rank 0 lost its CUDA context
        out["hard"] = "not raised"
    except RuntimeError as e:
        out["hard"] = str(e); out["poisoned_after_hard"] = ag.poisoned
    ag.barrier(20)
    if r == 0:                                                                   # rank 1 is gone: rank 0 must not wait for ever
        t0 = time.time()
        try:
            ag.agree(None, "step 4", timeout_s=3); out["gone"] = "not raised"
        except RuntimeError as e:
            out["gone"] = str(e); out["waited"] = time.time() - t0
    print(json.dumps(out))
""") % ROOT


def test_host_agreement_between_ranks_without_the_device(tmp_path):
    """bench.py's multi-GPU records agree on failures over a TCP store on the host (HostAgreement): soft failures are raised on every rank,
    a rank that lost its CUDA context poisons the run on every rank (no device collective afterwards), a rank that is gone costs a
    bounded wait."""
    import socket
    w = tmp_path / "worker4.py"
    w.write_text(WORKER4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ps = [subprocess.Popen([sys.executable, str(w), str(r), str(port)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in ps:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads([l for l in o.splitlines() if l.startswith("{")][-1]))
    a, b = outs
    for o in outs:
        assert o["clean"] and "no room" in o["soft"] and "rank 1" in o["soft"] and o["poisoned_after_soft"] is False
        assert "illegal memory access" in o["hard"] and "rank 0" in o["hard"] and o["poisoned_after_hard"] is True
    assert "did not answer" in a["gone"] and a["waited"] < 20
