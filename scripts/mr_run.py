"""One melt over N GPUs under torchrun (NCCL all-gather per round): checks bit-identity with the single-GPU walk and times it.
   torchrun --nproc-per-node N scripts/mr_run.py [nChains] [beadsPerChain]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200"))
import numpy as np, torch, torch.distributed as dist
import sarw_b200

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dens = float(sys.argv[3]) if len(sys.argv) > 3 else 0.92      # g/cc; lower = same tip density as a longer-chain melt early in its growth
L = (n * per / (0.03957 * dens / 0.92)) ** (1.0 / 3.0)
cfg = dict(nChains=n, boxSize=(L, L, L), minDist=1.0, maxTrials=1000, seed=7, device=local, targetMassDensity=dens)
stream = torch.cuda.current_stream()

def run(multi):
    slab = dict(slabRank=rank, slabRanks=world) if multi == 3 else {}      # 3: box decomposed into z-slabs, one per rank
    c = sarw_b200.Context(launchMode=1, **slab, **cfg); c.set_stream(stream.cuda_stream)
    cellBytes = c.cell_table_bytes()
    c.initiate()
    if multi >= 2: sarw_b200.connect_peers(c, dist, rank, world)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if multi >= 2: rounds = sarw_b200.grow_multirank_peer(c, dist, rank, world, connected=True)
    elif multi: rounds = sarw_b200.grow_multirank(c, dist, rank, world)
    else: rounds = c.grow(0, -1).rounds
    e1.record(stream); e1.synchronize()
    c.terminate()
    out = c.download(); st = c.stats().as_dict(); c.close()
    st["cellBytes"] = cellBytes
    return out, st, e0.elapsed_time(e1), rounds

single, st1, ms1, r1 = run(0)
multi, stN, msN, rN = run(1)
peer, stP, msP, rP = run(2)
slab, stS, msS, rS = run(3)
same = all(np.array_equal(a, b) for a, b in zip(single, multi)) and r1 == rN
sameP = all(np.array_equal(a, b) for a, b in zip(single, peer)) and r1 == rP
sameS = all(np.array_equal(a, b) for a, b in zip(single, slab)) and r1 == rS
t = torch.tensor([msN, ms1, msP, msS], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = torch.tensor([1.0 if same else 0.0, 1.0 if sameP else 0.0, 1.0 if sameS else 0.0], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
def cyc(st, r):
    return [round(st[k] / max(1, r)) for k in ("cycPropose", "cycVerify", "cycCleanup")]
if rank == 0:
    print("cycles/round [A, V, C]  single", cyc(st1, r1), " nccl", cyc(stN, rN), " peer", cyc(stP, rP), " slab", cyc(stS, rS))
    print("sub-phases/round single", [round(x / max(1, r1)) for x in st1["cycSub"]], " peer", [round(x / max(1, rP)) for x in stP["cycSub"]])
    print(json.dumps(dict(n_gpus=world, nChains=n, box=L, density=dens, conflicts=st1["conflicts"], beads=st1["nBeads"], rounds=r1, nccl_identical_on_all_ranks=bool(ok[0].item() == 1.0), peer_identical_on_all_ranks=bool(ok[1].item() == 1.0), slab_identical_on_all_ranks=bool(ok[2].item() == 1.0),
                          slab_gpu_ms=t[3].item(), slab_beads_per_s=st1["nBeads"] / t[3].item() * 1e3, cell_bytes_single=st1["cellBytes"], cell_bytes_slab_rank0=stS["cellBytes"],
                          peer_gpu_ms=t[2].item(), peer_beads_per_s=st1["nBeads"] / t[2].item() * 1e3, us_per_round_peer=t[2].item() * 1e3 / max(1, rP),
                          single_gpu_ms=t[1].item(), multi_gpu_ms=t[0].item(), single_beads_per_s=st1["nBeads"] / t[1].item() * 1e3,
                          multi_beads_per_s=st1["nBeads"] / t[0].item() * 1e3, us_per_round_multi=t[0].item() * 1e3 / max(1, rN))))
dist.barrier(); dist.destroy_process_group()
