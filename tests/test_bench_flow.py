"""bench.py's multi-GPU record (one_melt_record) run end to end on CPU: two gloo ranks, a stand-in for the device side (fake contexts, CUDA
events and streams), the REAL control flow - collectives, the agreement after every step that only one rank can fail at, the record
that comes out. Round 2's 8-GPU run hung for ten minutes because one rank left this flow through an exception; the second scenario
injects exactly that (rank 1 cannot allocate in the third melt) and both ranks must come out together, with the next collective intact."""
import json
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, json, types
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import bench

    # ---- stand-ins for the device: events, streams, "cuda" tensors live on the CPU
    class Event:
        def __init__(self, enable_timing=False): pass
        def record(self, stream=None): pass
        def synchronize(self): pass
        def elapsed_time(self, other): return 12.5
    torch.cuda.Event = Event
    torch.cuda.synchronize = lambda *a, **k: None
    _tensor = torch.tensor
    torch.tensor = lambda *a, **k: _tensor(*a, **{**k, "device": "cpu"}) if k.get("device") == "cuda" else _tensor(*a, **k)
    torch.Tensor.cuda = lambda self, *a, **k: self

    FAIL = os.environ.get("FAKE_FAIL", "")
    created = [0]
    class Stats:
        seedMs = 0.5; nBeads = 1000; trials = 23000; rounds = 10; conflicts = 3; overflowBeads = 7
        cycPropose = 100; cycVerify = 50; cycCleanup = 20; cycSub = [1, 2, 3, 4, 5, 6, 7, 8]
    class Context:
        def __init__(self, **kw):
            created[0] += 1
            self.kw = kw
            if FAIL == "create" and kw.get("slabRank") == 1 and created[0] == 3: raise MemoryError("cuMemCreate failed (2)")
        def set_stream(self, s): pass
        def cell_table_bytes(self): return 123
        def initiate(self): return True
        def mr_setup(self, r, n): return (1, 32)
        def mr_peer_export(self): return bytes(64)
        def mr_peer_connect(self, handles): assert len(handles) == 2 and all(len(h) == 64 for h in handles)
        def mr_grow_peer(self, n):
            if FAIL == "grow" and self.kw.get("slabRank") == 0 and created[0] == 2: raise RuntimeError("multi-rank growth aborted")
            return Stats()
        def grow(self, a, b): return Stats()
        def walk_checksum(self): return (1, 2, 3, 4)
        def download_box(self, lo, hi): return np.zeros((0, 3)), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32)
        def close(self): pass
    fake = types.SimpleNamespace(Context=Context, release_cached_memory=lambda d: 0)

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    env = types.SimpleNamespace(torch=torch, sarw=fake, stream=types.SimpleNamespace(cuda_stream=0), local=0, hbm_peak=6548.2, peak_src="test")
    bench.setup_host_agreement(dist, rank, world, device="cpu")
    out = {"agreement": bench._AGREE is not None}
    w = bench.melt_workload(2000, "test")
    try:
        rec = bench.one_melt_record(env, dist, rank, world, w, 1, 2, single_gpu_check=True, box_samples=0)
        out["rec"] = None if rec is None else {k: rec[k] for k in ("value", "all_ranks_agree", "matches_single_gpu", "steps", "cell_table_bytes_rank0")}
    except Exception as e:
        out["error"] = repr(e)
    t = _tensor([rank + 1]); dist.all_reduce(t); out["sum_after"] = int(t.item())      # the next collective still lines up
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0: print(json.dumps(gathered))
    dist.barrier(); dist.destroy_process_group()
""") % (ROOT, ROOT)


def _run(tmp_path, fail, port):
    w = tmp_path / "flow_worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", FAKE_FAIL=fail)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(w)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])


def test_one_melt_record_flow_two_ranks(tmp_path):
    a, b = _run(tmp_path, "", 29641)
    assert a["agreement"] and b["agreement"]
    assert "error" not in a and "error" not in b, (a, b)
    assert b["rec"] is None                                                   # the record is rank 0's
    rec = a["rec"]
    assert rec["steps"] == 2 and rec["all_ranks_agree"] and rec["matches_single_gpu"] and rec["cell_table_bytes_rank0"] == 123
    assert abs(rec["value"] - 2 * 1000 / (2 * 13.0e-3)) < 1e-6               # beads of the timed steps / (event time + seeding)
    assert a["sum_after"] == b["sum_after"] == 3


def test_a_rank_that_fails_takes_both_ranks_out_of_the_record_together(tmp_path):
    for fail, port, text in (("create", 29643, "cuMemCreate failed"), ("grow", 29645, "growth aborted")):
        a, b = _run(tmp_path, fail, port)
        assert "error" in a and "error" in b, (fail, a, b)                    # BOTH ranks, not just the one that failed
        assert text in a["error"] and text in b["error"]
        assert a["sum_after"] == b["sum_after"] == 3                          # and the collective that follows is not torn
