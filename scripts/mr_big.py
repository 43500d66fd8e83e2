"""Large one-melt run over N GPUs (peer-memory exchange) vs the single-GPU kernel, statistics only (no host download):
   torchrun --nproc-per-node N scripts/mr_big.py [nChains] [beadsPerChain] [nosingle|-] [slab|replicated]
Agreement is checked on the walk checksum (every coordinate and bond) and on the invariants (beads, rounds, trials in reference terms)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200"))
import numpy as np, torch, torch.distributed as dist
import sarw_b200

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
skip_single = len(sys.argv) > 3 and sys.argv[3] == "nosingle"
slab = not (len(sys.argv) > 4 and sys.argv[4] == "replicated")      # default: the box decomposed into z-slabs, one per rank
L = (n * per / 0.03957) ** (1.0 / 3.0)
cfg = dict(nChains=n, boxSize=(L, L, L), minDist=1.0, maxTrials=1000, seed=7, device=local)
stream = torch.cuda.current_stream()

def run(multi):
    t0 = time.time()
    kw = dict(slabRank=rank, slabRanks=world) if multi and slab else {}
    c = sarw_b200.Context(launchMode=1, **kw, **cfg); c.set_stream(stream.cuda_stream)
    cells = c.cell_table_bytes()
    c.initiate()
    if multi: sarw_b200.connect_peers(c, dist, rank, world)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if multi: c.mr_grow_peer(-1)
    else: c.grow(0, -1)
    e1.record(stream); e1.synchronize()
    st = c.stats().as_dict(); st["checksum"] = c.walk_checksum(); st["cellBytes"] = cells; c.close()
    return st, e0.elapsed_time(e1), time.time() - t0

res = {}
if not skip_single:
    st1, ms1, w1 = run(False)
    res.update(single_ms=ms1, single_beads_per_s=st1["nBeads"] / ms1 * 1e3, single=[st1[k] for k in ("nBeads", "rounds", "trials", "conflicts")],
               single_cyc=[round(st1[k] / max(1, st1["rounds"])) for k in ("cycPropose", "cycVerify", "cycCleanup")])
stN, msN, wN = run(True)
t = torch.tensor([msN], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
inv = torch.tensor([float(stN[k]) for k in ("nBeads", "rounds", "trials")] + [float(x & 0xFFFFFFFF) for x in stN["checksum"]] + [float(x >> 32) for x in stN["checksum"]], dtype=torch.float64, device="cuda")
lo, hi = inv.clone(), inv.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    res.update(n_gpus=world, nChains=n, box=L, beads=stN["nBeads"], target=stN["target"], rounds=stN["rounds"], trials_per_bead=stN["trials"] / stN["nBeads"],
               conflicts=stN["conflicts"], overflow=stN["overflowBeads"], dead=stN["deadChains"],
               peer_ms=t[0].item(), peer_beads_per_s=stN["nBeads"] / t[0].item() * 1e3, us_per_round=t[0].item() * 1e3 / max(1, stN["rounds"]),
               peer=[stN[k] for k in ("nBeads", "rounds", "trials", "conflicts")], all_ranks_agree=bool(torch.equal(lo, hi)),
               peer_cyc=[round(stN[k] / max(1, stN["rounds"])) for k in ("cycPropose", "cycVerify", "cycCleanup")],
               peer_sub=[round(x / max(1, stN["rounds"])) for x in stN["cycSub"]], setup_wall_s=wN)
    res.update(mode="slab" if slab else "replicated", cell_bytes_rank0=stN["cellBytes"], checksum=list(stN["checksum"]))
    if not skip_single: res["matches_single_gpu"] = res["single"][:3] == res["peer"][:3] and list(st1["checksum"]) == list(stN["checksum"])
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
