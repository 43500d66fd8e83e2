#!/usr/bin/env python
"""bench.py — accepted beads/s of the chain-growth hot path (BASELINE.json metric) on N B200s of one node.

A step = one complete pass of the hot path — SARW::initiateChain + the growth loop of main() (sarw.cpp:66-78) — over one BATCH of
synthetic melts of BASELINE.json configs[1] (PE melt, 100 chains x 1000 beads, 0.92 g/cc, ppp, minDist 1.0, maxTrials 1000).
A 100-chain melt grows in one thread-block cluster (13 of the 148 SMs), so the batch is M independent melts grown at the same time,
one CUDA stream + one host thread each, M = sarw_concurrent_capacity() (7 on B200) — the same thing the reference arm does with the
host cores (one serial instance per core). `single_melt` in the JSON line is one melt at a time. Every melt uses a fresh Philox
key, so no work is cached between steps.

  value : device-resident throughput — contexts (cell tables) are created before the timed region; a step is
          timed with CUDA events on the stream the kernels run on.
  e2e   : the same metric through the public C ABI with host buffers: sarw_create (allocations + table init),
          H2D of the inputs, seeding, growth, sarw_terminate and sarw_download of every bead/bond into host memory,
          wall clock around the calls.
  N > 1 : one process per GPU (torchrun). The melts of different ranks are independent (no data-path collective yet:
          spatial decomposition with NCCL halo exchange is the next §8e row); "scaling": "weak".
  --impl reference : the UNMODIFIED reference (oracle/_ref, built from /root/reference) timed on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[1] / SURVEY §8d C2
    "C2": dict(name="PE melt 100 chains x 1000 beads, 0.92 g/cc, ppp, minDist 1.0, maxTrials 1000 (BASELINE configs[1])",
               nChains=100, boxSize=(136.2, 136.2, 136.2), minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92),
    # BASELINE.json configs[2] / C3
    "C3": dict(name="PE melt 1e4 chains x 1000 beads, 0.92 g/cc, ppp, minDist 1.0, maxTrials 1000 (BASELINE configs[2])",
               nChains=10000, boxSize=(632.2, 632.2, 632.2), minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92),
}
# ONE melt over all N GPUs (strong scaling): not the headline workload, selected with --workload
ONE_MELT = {
    "ONE_MELT_1E7": dict(name="ONE PE melt of 1e5 chains, 1.0e7 beads (tip density of a 1000-bead melt), grown by all N GPUs together", nChains=100000,
                         boxSize=(1362.107287789021,) * 3, minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.092),
    "ONE_MELT_2E8": dict(name="ONE PE melt of 2e5 chains x 1000 beads (2.0e8 beads, 0.92 g/cc), grown by all N GPUs together", nChains=200000,
                         boxSize=(1716.1476441006014,) * 3, minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92),
    "ONE_MELT_3E8": dict(name="ONE PE melt of 3e5 chains x 1000 beads (3.0e8 beads, 0.92 g/cc; 121 GB of cells per GPU), grown by all N GPUs together", nChains=300000,
                         boxSize=(1964.4986505263048,) * 3, minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture
NCU_TRAFFIC = {"C2": (36.86e6, "profiles/r1_grow_cluster_kernel_C2_raw.csv")}
# algorithmic bytes per overlap query and per accepted bead
B_QUERY = 121.0
B_BEAD = 108.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        if os.environ.get("SARW_BENCH_NOSAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ctx_kwargs(w, seed):
    return dict(nChains=w["nChains"], boxSize=w["boxSize"], minDist=w["minDist"], targetMassDensity=w["targetMassDensity"],
                maxTrials=w["maxTrials"], boundary=w["boundary"], seed=seed)


# ------------------------------------------------------------------ reference on host cores (oracle/_ref)
def _ref_one(args):
    w, seed = args
    import oracle_bindings as ob
    R = ob.RefLib()
    r = R.walk(nChains=w["nChains"], box=w["boxSize"], minDist=w["minDist"], targetMassDensity=w["targetMassDensity"],
               maxTrials=w["maxTrials"], boundary=w["boundary"], rng="mt", seed=seed)
    return r.nAtoms, r.seconds, r.trials, r.rounds


def reference_throughput(w, instances, seed0):
    """`instances` independent melts of the reference, one per host core, run concurrently (the reference itself is
    strictly serial: 1 thread per instance). Returns (beads/s aggregate, wall seconds of the slowest, details)."""
    import multiprocessing as mp
    import oracle_bindings as ob
    if not ob.RefLib.available():
        if os.path.isdir("/root/reference"):
            ob.build_oracles()
    if not ob.RefLib.available():
        return None
    ctx = mp.get_context("fork")
    with ctx.Pool(instances) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_one, [(w, seed0 + i) for i in range(instances)])
        wall = time.perf_counter() - t0
    beads = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    per_core = statistics.median(r[0] / r[1] for r in res)
    return dict(value=beads / slowest, wall=wall, slowest=slowest, per_core=per_core, beads=beads,
                trials_per_bead=sum(r[2] for r in res) / beads, rounds=[r[3] for r in res])


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 32))
    times, values = [], []
    last = None
    for s in range(args.warmup + args.steps):
        r = reference_throughput(w, cores, 1000 + 100 * s)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (needs /root/reference at build time)"}))
            return
        if s >= args.warmup:
            times.append(r["slowest"]); values.append(r["value"]); last = r
    value = sum(values) / len(values)
    sample = (f"{cores} concurrent instances (1 thread each; the reference is serial) of the full workload per step, unmodified reference "
              f"sources (oracle/_ref) with mt19937_64 re-seeded per trial like upstream; seeding + growth loop timed (sarw.cpp:66-78)")
    line = {"impl": "reference", "metric": "accepted beads/sec", "value": value, "unit": "beads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": w["name"], "instances": cores},
            "cpu_baseline": {"value": value, "unit": "beads/s", "cores": cores, "kind": "reference", "sample": sample,
                             "per_core_beads_per_s": last["per_core"], "trials_per_bead": last["trials_per_bead"]},
            "e2e": {"value": value, "unit": "beads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def rank_seeds(rank, total_steps):
    """Distinct Philox keys per (rank, step): ranks shard the work as independent melts."""
    return [10_000 * (rank + 1) + s for s in range(total_steps)]


def aggregate_ranks(dist, device, t_dev, t_e2e, beads, e2e_beads):
    """Contract: time = MAX over ranks, work = SUM over ranks. `dist` is torch.distributed (any backend) or None."""
    import torch
    if dist is None:
        return t_dev, t_e2e, float(beads), float(e2e_beads)
    t = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=device)
    w = torch.tensor([float(beads), float(e2e_beads)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t[0].item(), t[1].item(), w[0].item(), w[1].item()


def run_one_melt(args, w):
    """Strong scaling: every step grows ONE melt with all N GPUs (replicated state, partitioned proposals and verification, records and
    conflict lists exchanged inside the persistent kernel over NVLink peer memory). N = 1 runs the single-GPU kernel."""
    import torch
    import sarw_b200
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    hbm_peak, peak_src = load_peaks()
    times, beads, stats = [], 0, None
    sampler = None
    for s in range(args.warmup + args.steps):
        c = sarw_b200.Context(device=local, launchMode=1, **ctx_kwargs(w, 31_000 + s))
        c.set_stream(stream.cuda_stream)
        c.initiate()
        if world > 1:
            sarw_b200.connect_peers(c, dist, rank, world)
        torch.cuda.synchronize()
        if dist: dist.barrier()
        if s == args.warmup and rank == 0:
            sampler = ClockSampler(local); sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = c.mr_grow_peer(-1) if world > 1 else c.grow(0, -1)
        e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1) + st.seedMs
        if dist:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        if s >= args.warmup:
            times.append(ms); beads += st.nBeads; stats = st
        c.close()
    if rank == 0:
        clocks = sampler.stop() if sampler else None
        t_dev = sum(times) / 1e3
        line = {"metric": "accepted beads/sec", "value": beads / t_dev, "unit": "beads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "multi_gpu": "one melt over all GPUs: replicated state, partitioned proposals + verification, in-kernel peer-memory exchange"
                           if world > 1 else "single GPU", "rounds_per_melt": stats.rounds, "trials_per_bead": stats.trials / max(1, stats.nBeads),
                           "conflicts_per_melt": stats.conflicts, "l2": "tables far larger than L2"},
                "e2e": None, "gpu_launches": int(args.steps * (1 + 4)), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": (stats.trials * B_QUERY + stats.nBeads * B_BEAD) / (times[-1] / 1e3) / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (stats.trials * B_QUERY + stats.nBeads * B_BEAD) / (times[-1] / 1e3) / 1e9 / world / hbm_peak, "traffic": None,
                             "kernel": "sarw::grow_kernel (persistent cooperative launch per GPU)", "peak_source": peak_src,
                             "algorithmic_bytes": f"(trials*{B_QUERY:.0f} + beads*{B_BEAD:.0f}) / N GPUs"},
                "note": "optional strong-scaling workload; the headline is the default workload"}
        print(json.dumps(line))
    if dist:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS) + sorted(ONE_MELT))
    ap.add_argument("--melts", type=int, default=0, help="independent melts grown concurrently per step (0 = sarw_concurrent_capacity of the device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary C3 / find_batch measurements")
    args = ap.parse_args()
    if args.workload in ONE_MELT:
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the one-melt strong-scaling workloads have no CPU arm (hours per melt on one core)"}))
            return
        return run_one_melt(args, ONE_MELT[args.workload])
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    import numpy as np
    import torch
    import sarw_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsarw_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    hbm_peak, peak_src = load_peaks()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2

    def one_melt_device(ctx):
        """initiate + grow + terminate on an existing context (blocking; returns the stats)"""
        ctx.initiate()
        st = ctx.grow(0, -1)
        ctx.terminate()
        return st

    # A melt of <= 256 chains grows in ONE thread-block cluster (13 of 148 SMs for configs[1]); a step is therefore a BATCH of M
    # independent melts grown at the same time, one CUDA stream + one host thread each, M = sarw_concurrent_capacity (7 on B200).
    # The reference arm does the same with the host cores (one serial instance per core). M = 1 for the large workloads.
    probe = sarw_b200.Context(device=local, **ctx_kwargs(w, 1))
    M = args.melts if args.melts > 0 else probe.concurrent_capacity()
    probe.close()
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=M)
    streams = [stream] if M == 1 else [torch.cuda.Stream() for _ in range(M)]

    def make_ctx(seed, m):
        c = sarw_b200.Context(device=local, **ctx_kwargs(w, seed))
        c.set_stream(streams[m].cuda_stream)
        return c

    def batch_device(ctxs):
        """one step: M melts at once; CUDA events on the launching streams: start fans out to every stream, the end event waits for all"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for sm in streams:
            if sm is not stream: sm.wait_event(e0)
        def work(m):
            with torch.cuda.stream(streams[m]):
                st = one_melt_device(ctxs[m])
                ev = torch.cuda.Event(); ev.record(streams[m])
            return st, ev
        res = list(pool.map(work, range(M)))
        for _, ev in res: stream.wait_event(ev)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), [r[0] for r in res]

    total = args.warmup + args.steps
    seeds = rank_seeds(rank, total * M)
    # -------- device-resident arm: the contexts (cell tables) of a step exist before its timed region starts and are destroyed after it,
    # so memory does not grow with --steps
    def step_contexts(i):
        return [make_ctx(seeds[i * M + m], m) for m in range(M)]

    def close_all(row):
        for c in row: c.close()

    # clocks / throttle reasons are sampled by rank 0 only, from before the warm-up (nvidia-smi takes a moment to start) to the end of the
    # timed region: one poller per rank (8 on a full node) holds the driver long enough to stretch the steps of every rank (measured at
    # N = 8: 15.8 ms per step against 9.6 ms end to end)
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    for i in range(args.warmup):
        row = step_contexts(i); flush.fill_(float(i)); batch_device(row); close_all(row)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    step_ms, beads, queries, trials, rounds, launches, grow_ms, seed_ms, conflicts, nMelts = [], 0, 0, 0, 0, 0, 0.0, 0.0, 0, 0
    for i in range(args.warmup, total):
        row = step_contexts(i)
        flush.fill_(float(i))          # L2 flush between timed iterations (outside the event pair)
        torch.cuda.synchronize()
        ms, sts = batch_device(row)
        step_ms.append(ms)
        for st in sts:
            beads += st.nBeads; queries += st.queries; trials += st.trials; rounds += st.rounds; nMelts += 1
            launches += st.growLaunches + st.seedLaunches; grow_ms += st.growMs; seed_ms += st.seedMs; conflicts += st.conflicts
        close_all(row)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sum(step_ms) / 1e3

    # one melt at a time on an otherwise idle GPU (latency of a melt; what `value` was before batching)
    single_ms, single_beads = [], 0
    if M > 1:
        for k in range(args.warmup + args.steps):
            c = sarw_b200.Context(device=local, **ctx_kwargs(w, 900_000 + rank * 1000 + k))
            c.set_stream(stream.cuda_stream)
            flush.fill_(float(k)); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); st = one_melt_device(c); e1.record(stream); e1.synchronize()
            if k >= args.warmup: single_ms.append(e0.elapsed_time(e1)); single_beads += st.nBeads
            c.close()

    # -------- end-to-end arm through the C ABI with host buffers (results land in pinned host memory): the same batch of M melts,
    # every melt created, grown, downloaded and destroyed inside the timed region
    cap = int(sarw_b200.target_count(w["boxSize"], w["targetMassDensity"]) + w["nChains"] + 16)
    pinned = [{"xyz": torch.empty((cap, 3), dtype=torch.float64, pin_memory=True).numpy(), "chainID": torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
               "type": torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(), "bonds": torch.empty((cap, 2), dtype=torch.int64, pin_memory=True).numpy()} for _ in range(M)]
    e2e_times, e2e_beads, h2d, d2h, e2e_launch = [], 0, 0, 0, 0
    import gc

    def e2e_melt(args_):
        m, seed = args_
        c = sarw_b200.Context(device=local, **ctx_kwargs(w, seed))
        c.set_stream(streams[m].cuda_stream)
        c.initiate(); st = c.grow(0, -1); c.terminate()
        xyz, cid, typ, bonds = c.download(pinned[m])
        nbytes = xyz.nbytes + cid.nbytes + typ.nbytes + bonds.nbytes
        c.close()
        return st, nbytes

    for i in range(total):
        gc.collect(); gc.disable()                 # no collector pauses inside the timed call sequence
        flush.fill_(float(i)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = list(pool.map(e2e_melt, [(m, seeds[i * M + m] + 5_000_000) for m in range(M)]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gc.enable()
        if os.environ.get("SARW_BENCH_DEBUG"): print(f"e2e step {i}: {dt*1e3:.3f} ms for {M} melts", file=sys.stderr)
        if i >= args.warmup:
            e2e_times.append(dt)
            for st, nbytes in res:
                e2e_beads += st.nBeads
                h2d += sarw_b200.C.sizeof(sarw_b200.Params) + 64                      # parameters + control block
                d2h += nbytes + 3 * 64 + st.growLaunches * 64 + 8                     # results + control/stat reads
                e2e_launch += st.growLaunches + st.seedLaunches + 5                   # + flag, 2x cub scan, gather
    pool.shutdown()

    # -------- aggregate over ranks (max time, summed work)
    t_e2e = sum(e2e_times)
    t_dev, t_e2e, beads_all, e2e_beads_all = aggregate_ranks(dist, "cuda", t_dev, t_e2e, beads, e2e_beads)

    if rank == 0:
        value = beads_all / t_dev
        # roofline of the dominant kernel (grow_kernel: ~all of the step): algorithmic bytes / its CUDA-event duration
        # bytes per bead = T*B_q + B_a with T = trials per accepted bead in the reference's terms (what the serial
        # loop of sarw.cpp:259 evaluates; trials of provably dead chains are decided without memory traffic here)
        alg_bytes = trials * B_QUERY + beads * B_BEAD
        achieved = alg_bytes / (grow_ms / 1e3) / 1e9       # per launch: summed bytes / summed launch durations
        line = {
            "metric": "accepted beads/sec", "value": value, "unit": "beads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"], "melts_per_step": M,
                       "batch": (f"{M} independent melts of the workload grown at the same time (one CUDA stream + one host thread each; a melt of "
                                 f"{w['nChains']} chains occupies one thread-block cluster, {M} clusters fit the device)") if M > 1 else "one melt per step",
                       "seeds": "distinct Philox key per melt, step and rank", "l2": "256 MB flush write between timed steps",
                       "multi_gpu": "independent melts per rank, no data-path collective" if world > 1 else "single GPU",
                       "rounds_per_melt": rounds / nMelts, "trials_per_bead": trials / max(1, beads), "conflicts_per_melt": conflicts / nMelts},
            "e2e": {"value": e2e_beads_all / t_e2e, "unit": "beads/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "ms_per_step_median": 1e3 * statistics.median(e2e_times), "ms_per_step_max": 1e3 * max(e2e_times),
                    "includes": "per melt: sarw_create (alloc + cell-table init), seeding, growth, terminate, sarw_download to pinned host memory, sarw_destroy"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0], "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1],
                         "kernel": "sarw::grow_cluster_kernel<8> (one thread-block cluster, persistent, one launch per step)" if w["nChains"] <= 128 else "sarw::grow_kernel (persistent cooperative launch)", "peak_source": peak_src,
                         "algorithmic_bytes": f"trials*{B_QUERY:.0f} + beads*{B_BEAD:.0f} (SURVEY §8d, trials in reference terms)",
                         "trials_per_launch": trials / nMelts, "candidates_evaluated_per_launch": queries / nMelts,
                         "avg_launch_ms": grow_ms / nMelts, "concurrent_launches": M,
                         "note": "per launch (one melt); latency-bound: ~1e2 chains per round, ~1e3 dependent rounds. M launches run side by side in a step"},
        }
        if single_ms:
            line["single_melt"] = {"ms_per_melt": statistics.median(single_ms), "beads_per_s": single_beads / (sum(single_ms) / 1e3),
                                   "note": "one melt at a time on an otherwise idle GPU, device time (CUDA events)"}
        if not args.no_extra:
            line["extra"] = extra_measurements(sarw_b200, torch, stream, flush, local)
        if not args.no_cpu_baseline and world == 1:
            cores = max(1, min(os.cpu_count() or 1, 32))
            r = reference_throughput(w, cores, 777)
            if r is not None:
                line["cpu_baseline"] = {"value": r["value"], "unit": "beads/s", "cores": cores, "kind": "reference",
                                        "sample": f"{cores} concurrent single-thread instances of the full workload, unmodified reference (oracle/_ref), "
                                                  f"seeding + growth timed; slowest instance {r['slowest']:.2f} s", "per_core_beads_per_s": r["per_core"],
                                        "trials_per_bead": r["trials_per_bead"]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "beads/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        print(json.dumps(line))
    if dist:
        dist.barrier(); dist.destroy_process_group()


def extra_measurements(sarw_b200, torch, stream, flush, local):
    """Secondary numbers (not the headline): BASELINE configs[2] growth and the overlap-query kernel on its melt."""
    import numpy as np
    out = {}
    w = WORKLOADS["C3"]
    try:
        c = sarw_b200.Context(device=local, **ctx_kwargs(w, 4242))
        c.set_stream(stream.cuda_stream)
        c.initiate(); st = c.grow(0, -1); c.terminate()
        out["C3_growth"] = {"workload": w["name"], "beads": st.nBeads, "rounds": st.rounds, "grow_ms": st.growMs, "seed_ms": st.seedMs,
                            "beads_per_s": st.nBeads / ((st.growMs + st.seedMs) / 1e3), "trials_per_bead": st.trials / st.nBeads,
                            "queries": st.queries, "conflicts": st.conflicts}
        # K5 on the finished melt: 16 Mi random candidates resident in HBM
        n = 1 << 24
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        q = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g) * w["boxSize"][0]
        hit = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            c.find_device(q.data_ptr(), None, n, hit.data_ptr())
        torch.cuda.synchronize()
        ms = []
        for i in range(5):
            flush.fill_(float(i)); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); c.find_device(q.data_ptr(), None, n, hit.data_ptr()); e1.record(stream); e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = statistics.median(ms) / 1e3
        hbm_peak, _ = load_peaks()
        bq = 134.0 + 24 + 1   # SURVEY §8d B_q at full density + candidate read + result write
        out["find_batch"] = {"queries": n, "ms": t * 1e3, "queries_per_s": n / t, "algorithmic_GBps": n * bq / t / 1e9,
                             "frac_of_hbm_peak": n * bq / t / 1e9 / hbm_peak, "hit_fraction": float(hit.float().mean().item()),
                             "bytes_per_query": bq}
        c.close()
    except Exception as e:   # secondary numbers must never take the headline down
        out["error"] = repr(e)
    return out


if __name__ == "__main__":
    main()
