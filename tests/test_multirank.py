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
