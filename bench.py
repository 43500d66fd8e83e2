#!/usr/bin/env python
"""bench.py — accepted beads/s of the chain-growth hot path (BASELINE.json metric) on N B200s of one node.

A step = one complete pass of the hot path — SARW::initiateChain + the growth loop of main() (sarw.cpp:66-78) — over one BATCH of
synthetic melts of the workload. The headline workload is BASELINE.json configs[1] (PE melt, 100 chains x 1000 beads, 0.92 g/cc,
ppp, minDist 1.0, maxTrials 1000): such a melt grows in ONE thread-block cluster (13 of the 148 SMs), so the batch is M independent
melts grown at the same time, M = sarw_concurrent_capacity() (7 on B200), all driven by ONE host thread (sarw_grow_async on M
streams, then sarw_wait) — the same thing the reference arm does with the host cores (one serial instance per core).
Every melt uses a fresh Philox key, so no work is cached between steps.

  value : device-resident throughput — contexts (cell tables) are created before the timed region; a step is timed with CUDA events
          on the streams the kernels run on.
  e2e   : the same metric through the public C ABI with host buffers: sarw_create (allocations + table init), H2D of the inputs,
          seeding, growth, sarw_terminate and sarw_download of every bead/bond into pinned host memory, sarw_destroy; wall clock.
  workloads : the same two measurements, as full records of their own (repeated steps, roofline, e2e, clocks), for BASELINE
          configs[2] (1e4 x 1000 melt, 1e7 beads) and configs[3] (1e4-chain grafted brush); N = 1 only.
  N > 1 : one process per GPU (torchrun). The headline melts of different ranks are independent ("scaling": "weak", no data-path
          collective). The record `one_melt` is ONE melt of 5e7*N beads grown by all N GPUs together (sarw_mr_grow_peer on slab contexts: the box is
          decomposed into N slabs along z, each rank holds the cell layers of its slab + halo and owns the chains whose tip is inside; bead
          records, conflict lists and re-resolution results are exchanged inside the persistent kernel over NVLink peer memory), with
          every rank's walk checksum compared with the others' and with the single-GPU walk. At N = 8 the record `C5` is
          BASELINE configs[4]: one melt of 1e6 chains x 1000 beads (1e9 beads).
  --impl reference : the UNMODIFIED reference (oracle/_ref, built from /root/reference) timed on the host cores.
"""
import argparse
import gc
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

C4_GRAFT_SEP = 2.6
WORKLOADS = {
    # BASELINE.json configs[1] / SURVEY §8d C2
    "C2": dict(name="PE melt 100 chains x 1000 beads, 0.92 g/cc, ppp, minDist 1.0, maxTrials 1000 (BASELINE configs[1])",
               nChains=100, boxSize=(136.2, 136.2, 136.2), minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92),
    # BASELINE.json configs[2] / C3
    "C3": dict(name="PE melt 1e4 chains x 1000 beads, 0.92 g/cc, ppp, minDist 1.0, maxTrials 1000 (BASELINE configs[2])",
               nChains=10000, boxSize=(632.2, 632.2, 632.2), minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92),
    # BASELINE.json configs[3] / C4. BASELINE.md spaces the graft points 1.0 A (= minDist) apart; with that spacing the UNMODIFIED
    # reference fails its own seeding (a second bead lands within minDist of a later graft point, sarw.cpp:313-314,353), so the graft
    # points here are 2.6 A (> bond + minDist) apart — the same recipe as tests/cases.py C4 and the golden hash.
    "C4": dict(name="grafted brush: 1e4 chains in a 632.2 A box, graftFraction 0.5 (5000 chains grafted on z = 0 from a graft file, points >= 2.6 A apart "
                    "instead of BASELINE.md's 1.0 A, which the reference itself cannot seed), growthBias 0,0,1, boundary ppf, 0.92 g/cc target (BASELINE configs[3])",
               nChains=10000, boxSize=(632.2, 632.2, 632.2), minDist=1.0, maxTrials=1000, boundary="ppf", targetMassDensity=0.92,
               nGraftChains=5000, graftLoc="file", growthBias=(0.0, 0.0, 1.0), grafts="brush"),
}


def melt_workload(nChains, what):
    """a PE melt of nChains x 1000 beads at 0.92 g/cc: the box of BASELINE configs[2] scaled to the chain count"""
    L = 632.2 * (nChains / 10000.0) ** (1.0 / 3.0)
    return dict(name=f"ONE PE melt of {nChains} chains x 1000 beads ({nChains * 1000:.1e} beads, 0.92 g/cc, L = {L:.1f} A), {what}",
                nChains=nChains, boxSize=(L, L, L), minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.92)


# ONE melt over all N GPUs (strong scaling), selected with --workload
ONE_MELT = {
    "ONE_MELT_1E7": dict(name="ONE PE melt of 1e5 chains, 1.0e7 beads (tip density of a 1000-bead melt), grown by all N GPUs together", nChains=100000,
                         boxSize=(1362.107287789021,) * 3, minDist=1.0, maxTrials=1000, boundary="ppp", targetMassDensity=0.092),
    "ONE_MELT_2E8": melt_workload(200000, "grown by all N GPUs together"),
    "ONE_MELT_3E8": melt_workload(300000, "grown by all N GPUs together"),
    "C5": melt_workload(1000000, "BASELINE configs[4], grown by all N GPUs together"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full captures
NCU_TRAFFIC = {"C2": (9.88e6 + 0.02e6, "profiles/r2_grow_cluster_kernel_C2_raw.csv (one melt; round 1 with 128-byte cells: 36.9 MB)"),
               "C3": (5.646e9 + 0.761e9, "profiles/r2_grow_kernel_C3_raw.csv (one melt; round 1 with 128-byte cells: 13.2 GB)")}
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
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

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
    kw = dict(nChains=w["nChains"], boxSize=w["boxSize"], minDist=w["minDist"], targetMassDensity=w["targetMassDensity"],
              maxTrials=w["maxTrials"], boundary=w["boundary"], seed=seed)
    for k in ("nGraftChains", "graftLoc", "growthBias"):
        if k in w: kw[k] = w[k]
    return kw


_grafts_cache = {}


def workload_grafts(w):
    """graft points of a workload (None for melts); the brush uses the recipe of tests/cases.py (the one the golden hash pins)"""
    if w.get("grafts") != "brush":
        return None
    if "brush" not in _grafts_cache:
        import cases
        _grafts_cache["brush"] = cases.brush_grafts(n=w["nGraftChains"], L=w["boxSize"][0], sep=C4_GRAFT_SEP)
    return _grafts_cache["brush"]


# ------------------------------------------------------------------ reference on host cores (oracle/_ref)
def _ref_one(args):
    w, seed, rng, maxRounds = args
    import oracle_bindings as ob
    R = ob.RefLib()
    r = R.walk(nChains=w["nChains"], box=w["boxSize"], minDist=w["minDist"], targetMassDensity=w["targetMassDensity"],
               maxTrials=w["maxTrials"], boundary=w["boundary"], rng=rng, seed=seed, maxRounds=maxRounds)
    return r.nAtoms, r.seconds, r.trials, r.rounds


def reference_throughput(w, instances, seed0, rng="mt", maxRounds=-1):
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
        res = pool.map(_ref_one, [(w, seed0 + i, rng, maxRounds) for i in range(instances)])
        wall = time.perf_counter() - t0
    beads = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    per_core = statistics.median(r[0] / r[1] for r in res)
    return dict(value=beads / slowest, wall=wall, slowest=slowest, per_core=per_core, beads=beads,
                trials_per_bead=sum(r[2] for r in res) / beads, rounds=[r[3] for r in res])


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
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
    keyed = reference_throughput(w, cores, 5000, rng="keyed")
    line = {"impl": "reference", "metric": "accepted beads/sec", "value": value, "unit": "beads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": w["name"], "instances": cores},
            "cpu_baseline": {"value": value, "unit": "beads/s", "cores": cores, "kind": "reference", "sample": sample,
                             "per_core_beads_per_s": last["per_core"], "trials_per_bead": last["trials_per_bead"],
                             "keyed_rng_beads_per_s": keyed["value"] if keyed else None,
                             "keyed_rng_note": "same unmodified reference with the stand-in's Random::seed reduced to a counter set-up (no mt19937_64 "
                                               "re-initialisation per trial): upper bound of the reference's speed whatever JDFTx's Random::seed costs"},
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


class Env:
    """what every measurement needs: the binding, torch, the launching stream, the L2 flush buffer, the device"""
    def __init__(self, sarw_b200, torch, local):
        self.sarw, self.torch, self.local = sarw_b200, torch, local
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2
        self.hbm_peak, self.peak_src = load_peaks()
        self._streams = {}

    def streams(self, M):
        if M == 1:
            return [self.stream]
        if M not in self._streams:
            self._streams[M] = [self.torch.cuda.Stream() for _ in range(M)]
        return self._streams[M]

    def l2_flush(self, i):
        self.flush.fill_(float(i)); self.torch.cuda.synchronize()


def make_ctx(env, w, seed, stream):
    c = env.sarw.Context(device=env.local, **ctx_kwargs(w, seed))
    c.set_stream(stream.cuda_stream)
    g = workload_grafts(w)
    if g is not None: c.set_grafts(g)
    return c


def measure_device(env, w, M, seeds, warmup, steps):
    """device-resident arm: the contexts (cell tables) of a step exist before its timed region starts and are destroyed after it, so
    memory does not grow with the step count. ONE host thread: seeding + the growth loop of every melt are enqueued back to back on the
    melt's stream (sarw_grow_async), then each is waited for (sarw_wait: one synchronisation per melt). CUDA events on the launching
    streams: the start event fans out to every stream, the end event waits for all of them."""
    torch, stream, streams = env.torch, env.stream, env.streams(M)
    step_ms, acc = [], dict(beads=0, queries=0, trials=0, rounds=0, launches=0, grow_ms=0.0, seed_ms=0.0, conflicts=0, melts=0)
    for i in range(warmup + steps):
        ctxs = [make_ctx(env, w, seeds[i * M + m], streams[m]) for m in range(M)]
        env.l2_flush(i)          # L2 flush between timed iterations (outside the event pair)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for sm in streams:
            if sm is not stream: sm.wait_event(e0)
        for c in ctxs: c.grow_async(0, -1)
        for sm in streams:
            if sm is not stream:
                ev = torch.cuda.Event(); ev.record(sm); stream.wait_event(ev)
        e1.record(stream)
        sts = [c.wait() for c in ctxs]
        for c in ctxs: c.terminate()
        e1.synchronize()
        if i >= warmup:
            step_ms.append(e0.elapsed_time(e1))
            for st in sts:
                acc["beads"] += st.nBeads; acc["queries"] += st.queries; acc["trials"] += st.trials; acc["rounds"] += st.rounds; acc["melts"] += 1
                acc["launches"] += st.growLaunches + st.seedLaunches; acc["grow_ms"] += st.growMs; acc["seed_ms"] += st.seedMs; acc["conflicts"] += st.conflicts
        for c in ctxs: c.close()
    torch.cuda.synchronize()
    return step_ms, acc


def measure_e2e(env, w, M, seeds, warmup, steps):
    """end-to-end arm through the C ABI with host buffers (results land in pinned host memory): the same batch of M melts, every melt
    created, seeded, grown, downloaded and destroyed inside the timed region; wall clock, one host thread."""
    torch, streams = env.torch, env.streams(M)
    cap = int(env.sarw.target_count(w["boxSize"], w["targetMassDensity"]) + w["nChains"] + 16)
    pinned = [{"xyz": torch.empty((cap, 3), dtype=torch.float64, pin_memory=True).numpy(), "chainID": torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(),
               "type": torch.empty(cap, dtype=torch.int32, pin_memory=True).numpy(), "bonds": torch.empty((cap, 2), dtype=torch.int64, pin_memory=True).numpy()} for _ in range(M)]
    grafts = workload_grafts(w)
    times, acc = [], dict(beads=0, h2d=0, d2h=0, launches=0)
    for i in range(warmup + steps):
        gc.collect(); gc.disable()                 # no collector pauses inside the timed call sequence
        env.l2_flush(i)
        t0 = time.perf_counter()
        ctxs = []
        for m in range(M):
            c = make_ctx(env, w, seeds[i * M + m] + 5_000_000, streams[m])
            c.grow_async(0, -1)
            ctxs.append(c)
        res = []
        for m, c in enumerate(ctxs):
            st = c.wait(); c.terminate()
            xyz, cid, typ, bonds = c.download(pinned[m])
            res.append((st, xyz.nbytes + cid.nbytes + typ.nbytes + bonds.nbytes))
            c.close()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gc.enable()
        if os.environ.get("SARW_BENCH_DEBUG"): print(f"e2e step {i}: {dt*1e3:.3f} ms for {M} melts", file=sys.stderr)
        if i >= warmup:
            times.append(dt)
            for st, nbytes in res:
                acc["beads"] += st.nBeads
                acc["h2d"] += env.sarw.C.sizeof(env.sarw.Params) + 64 + (0 if grafts is None else grafts.nbytes)   # parameters + control block + graft points
                acc["d2h"] += nbytes + 2 * 64 + st.growLaunches * 64 + 8               # results + control/stat reads
                acc["launches"] += st.growLaunches + st.seedLaunches + 5               # + flag, 2x cub scan, gather
    del pinned
    return times, acc


def grow_kernel_name(w):
    return ("sarw::grow_cluster_kernel<8> (one thread-block cluster, persistent, one launch per melt)" if w["nChains"] <= 128
            else "sarw::grow_kernel<8,false> (persistent cooperative launch, one per melt)")


def roofline_record(env, w, key, acc, M):
    """dominant kernel = the growth kernel (~all of the step). SURVEY §8d: bytes per bead = T*B_q + B_a with T = trials per accepted
    bead in the reference's terms (what the serial loop of sarw.cpp:259 evaluates; trials of provably dead chains are decided without
    memory traffic here). achieved = summed algorithmic bytes / summed launch durations (the kernel's own CUDA-event pair on its stream)."""
    alg = acc["trials"] * B_QUERY + acc["beads"] * B_BEAD
    achieved = alg / (acc["grow_ms"] / 1e3) / 1e9
    traffic, src = NCU_TRAFFIC.get(key, (None, None))
    return {"bound": "hbm", "achieved": achieved, "peak": env.hbm_peak, "unit": "GB/s", "frac": achieved / env.hbm_peak, "traffic": traffic, "traffic_source": src,
            "kernel": grow_kernel_name(w), "peak_source": env.peak_src,
            "algorithmic_bytes": f"trials*{B_QUERY:.0f} + beads*{B_BEAD:.0f} (SURVEY §8d, trials in reference terms)",
            "algorithmic_bytes_per_launch": alg / acc["melts"], "trials_per_launch": acc["trials"] / acc["melts"],
            "candidates_evaluated_per_launch": acc["queries"] / acc["melts"], "avg_launch_ms": acc["grow_ms"] / acc["melts"], "concurrent_launches": M,
            "note": "per launch (one melt); latency-bound (~1e3 dependent rounds per melt), see DESIGN.md section 5. Trials of chains that are provably dead (every "
                    "torsion blocked: the brush has thousands of them) are counted as the reference burns them (maxTrials per round) but cost no memory traffic here, so "
                    "frac can exceed 1 for such workloads"}


def workload_record(env, key, warmup=2, steps=5):
    """a first-class record of a single-GPU workload other than the headline: repeated steps, roofline, e2e, clocks"""
    w = WORKLOADS[key]
    sampler = ClockSampler(env.local).start()
    seeds = [77_000 + s for s in range(warmup + steps)]
    step_ms, acc = measure_device(env, w, 1, seeds, warmup, steps)
    e2e_t, e2e = measure_e2e(env, w, 1, seeds, 1, steps)
    clocks = sampler.stop()
    t = sum(step_ms) / 1e3
    return {"workload": w["name"], "metric": "accepted beads/sec", "unit": "beads/s", "value": acc["beads"] / t, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * t / steps, "ms_per_step_median": statistics.median(step_ms), "ms_per_step_min": min(step_ms), "ms_per_step_max": max(step_ms),
            "beads_per_melt": acc["beads"] / acc["melts"], "rounds_per_melt": acc["rounds"] / acc["melts"], "trials_per_bead": acc["trials"] / max(1, acc["beads"]),
            "conflicts_per_melt": acc["conflicts"] / acc["melts"], "seed_ms_per_melt": acc["seed_ms"] / acc["melts"], "grow_ms_per_melt": acc["grow_ms"] / acc["melts"],
            "l2": "256 MB flush write between timed steps; tables far larger than L2",
            "e2e": {"value": e2e["beads"] / sum(e2e_t), "unit": "beads/s", "h2d_bytes_per_step": e2e["h2d"] // steps, "d2h_bytes_per_step": e2e["d2h"] // steps,
                    "ms_per_step": 1e3 * sum(e2e_t) / steps, "ms_per_step_median": 1e3 * statistics.median(e2e_t), "ms_per_step_max": 1e3 * max(e2e_t),
                    "includes": "sarw_create (alloc + cell-table init), graft upload, seeding, growth, terminate, sarw_download of every bead and bond to pinned host memory, sarw_destroy"},
            "gpu_launches": int(acc["launches"]), "clocks": clocks, "roofline": roofline_record(env, w, key, acc, 1)}


# ------------------------------------------------------------------ ONE melt over all GPUs
class HostAgreement:
    """The ranks agree on "did that step work everywhere?" through a TCP key-value store on the HOST, not through NCCL: a rank whose CUDA
    context died (a faulting kernel) can still say so, and the healthy ranks do not sit in a device collective waiting for it until the
    NCCL watchdog kills them. `poisoned` = some rank lost its context (or stopped answering): no device collective after that."""

    def __init__(self, rank, world, port=None, timeout_s=60):
        import datetime
        import torch.distributed as dist
        self.rank, self.world, self.n, self.poisoned = rank, world, 0, False
        port = port or int(os.environ.get("MASTER_PORT", "29500")) + 29
        self.store = dist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), port, world, rank == 0,
                                   timeout=datetime.timedelta(seconds=timeout_s), wait_for_workers=False)

    def _round(self, mine, timeout_s):
        import datetime
        self.n += 1
        self.store.set(f"k{self.n}/{self.rank}", mine)
        keys = [f"k{self.n}/{r}" for r in range(self.world)]
        try:
            self.store.wait(keys, datetime.timedelta(seconds=timeout_s))
        except Exception as e:   # noqa: BLE001  (a rank that never answers: its process is gone or stuck)
            self.poisoned = True
            raise RuntimeError(f"a rank did not answer within {timeout_s} s ({e!r})")
        return [self.store.get(k).decode(errors="replace") for k in keys]

    def agree(self, err, what, hard=False, timeout_s=200):
        """every rank passes its local exception (or None); raises on ALL ranks if any rank had one"""
        mine = "" if err is None else ("HARD " if hard else "SOFT ") + repr(err)[:400]
        answers = self._round(mine, timeout_s)
        bad = [(r, a) for r, a in enumerate(answers) if a]
        if any(a.startswith("HARD ") for _, a in bad):
            self.poisoned = True
        if bad:
            raise RuntimeError(f"{what}: " + "; ".join(f"rank {r}: {a[5:]}" for r, a in bad))

    def barrier(self, timeout_s=60):
        self._round("", timeout_s)


_AGREE = None   # set by setup_host_agreement when the ranks of a multi-GPU bench run can all reach the store


def cuda_context_alive():
    try:
        import torch
        torch.cuda.synchronize()
        return True
    except Exception:   # noqa: BLE001
        return False


def setup_host_agreement(dist, rank, world, device="cuda"):
    """create the store on every rank and make sure, with one (still safe) collective, that ALL ranks have it; else nobody uses it"""
    global _AGREE
    import torch
    ag = None
    try:
        ag = HostAgreement(rank, world)
    except Exception:   # noqa: BLE001
        ag = None
    t = torch.tensor([1 if ag is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    _AGREE = ag if t.item() == 1 else None
    return _AGREE


def all_ranks_ok(dist, err, what, device="cuda"):
    """Every rank reports whether a step that only IT can fail at (allocation, a launch) went through; if any rank failed, ALL ranks raise
    here, together. Without this a rank that raises leaves the others waiting in the next collective until the NCCL watchdog (minutes).
    Over the host store when there is one (see HostAgreement), else over the process group."""
    if dist is None:
        if err is not None:
            raise err
        return
    if _AGREE is not None:
        _AGREE.agree(err, what, hard=err is not None and not cuda_context_alive())
        return
    import torch
    t = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if t.item():
        raise RuntimeError(f"{what}: {err!r}" if err is not None else f"{what}: failed on another rank")


def connect_peers_checked(c, dist, rank, world):
    """sarw_b200.connect_peers with the local steps (buffer allocation, opening the IPC handles) reported to all ranks before the next
    collective"""
    err, mine = None, None
    try:
        c.mr_setup(rank, world)
        mine = c.mr_peer_export()
    except Exception as e:   # noqa: BLE001
        err = e
    all_ranks_ok(dist, err, f"rank {rank}: exchange buffers")
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    try:
        c.mr_peer_connect(handles)
    except Exception as e:   # noqa: BLE001
        err = e
    all_ranks_ok(dist, err, f"rank {rank}: opening the peers' buffers")


def checksum_words(torch, dist, c):
    """this rank's walk checksum and every rank's (all-gathered as int64 bit patterns)"""
    import numpy as np
    mine = np.array(c.walk_checksum(), dtype=np.uint64)
    if dist is None:
        return mine, [mine]
    t = torch.from_numpy(mine.view(np.int64).copy()).cuda()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return mine, [o.cpu().numpy().view(np.uint64) for o in out]


def one_melt_record(env, dist, rank, world, w, warmup, steps, single_gpu_check=True, box_samples=0, rdf=False, replicated=False):
    """Strong scaling: every step grows ONE melt with all N GPUs: the box is decomposed into N slabs along z, each rank holds the cell layers
    of its slab + halo and proposes / verifies / re-resolves the chains whose tip is inside; bead records, conflict lists and re-resolution
    results are exchanged inside the persistent kernel over NVLink peer memory (SARW_BENCH_REPLICATED=1: round 1's mode with every table
    on every rank). N = 1 runs the single-GPU kernel. Every rank's walk checksum (sarw_walk_checksum: every coordinate and bond) is compared
    with the others'; rank 0 then grows the melt of the last step once more on its own and compares that checksum too. Steps that only
    one rank can fail at are followed by all_ranks_ok, so the ranks leave a failed record together. Returns the record on rank 0."""
    torch, sarw_b200, stream, local = env.torch, env.sarw, env.stream, env.local
    import numpy as np
    times, beads, st, agree, sums = [], 0, None, True, None
    sampler = None
    extra = {}
    for s in range(warmup + steps):
        seed = 31_000 + s
        slab = dict(slabRank=rank, slabRanks=world) if world > 1 and not replicated and not os.environ.get("SARW_BENCH_REPLICATED") else {}
        c, err, cell_bytes = None, None, 0
        try:
            c = sarw_b200.Context(device=local, launchMode=1, **slab, **ctx_kwargs(w, seed))
            c.set_stream(stream.cuda_stream)
            cell_bytes = c.cell_table_bytes()
            c.initiate()
        except Exception as e:   # noqa: BLE001
            err = e
        try:
            all_ranks_ok(dist, err, f"rank {rank}: creating / seeding the context")
        except Exception:
            if c is not None: c.close()
            if sampler: sampler.stop()
            raise
        if world > 1:
            try:
                connect_peers_checked(c, dist, rank, world)
            except Exception:
                c.close()
                raise
        torch.cuda.synchronize()
        if dist: dist.barrier()
        if s == warmup and rank == 0:
            sampler = ClockSampler(local).start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        try:
            st = c.mr_grow_peer(-1) if world > 1 else c.grow(0, -1)
            e1.record(stream); e1.synchronize()
        except Exception as e:   # noqa: BLE001  (a rank whose launch aborts makes the others time out inside their kernels: every rank lands here)
            err = e
        try:
            all_ranks_ok(dist, err, f"rank {rank}: growth")
        except Exception:
            c.close()
            if sampler: sampler.stop()
            raise
        ms = e0.elapsed_time(e1) + st.seedMs
        if dist:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        mine, sums = checksum_words(torch, dist, c)
        agree = agree and all(np.array_equal(x, sums[0]) for x in sums)
        if s >= warmup:
            times.append(ms); beads += st.nBeads
        last = s == warmup + steps - 1
        if last and rank == 0 and (box_samples or rdf):
            try:
                extra = verify_large_walk(c, w, box_samples, rdf)
            except Exception as e:   # noqa: BLE001  (rank 0 only: must not skip the collectives that follow)
                extra = {"verify_error": repr(e)}
        c.close()
    clocks = sampler.stop() if sampler else None
    single = None
    if single_gpu_check and world > 1:
        # the same melt (seed of the last step) on ONE GPU: rank 0 alone, the other ranks wait at the barrier
        if rank == 0:
            c = None
            try:
                sarw_b200.release_cached_memory(local)
                c = sarw_b200.Context(device=local, launchMode=1, **ctx_kwargs(w, 31_000 + warmup + steps - 1))
                c.set_stream(stream.cuda_stream)
                c.initiate()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); st1 = c.grow(0, -1); e1.record(stream); e1.synchronize()
                one = np.array(c.walk_checksum(), dtype=np.uint64)
                single = {"ms": e0.elapsed_time(e1) + st1.seedMs, "beads": st1.nBeads, "checksum_equal": bool(np.array_equal(one, sums[0])),
                          "invariants_equal": (st1.nBeads, st1.rounds, st1.trials) == (st.nBeads, st.rounds, st.trials),
                          "conflict_resolutions": [st1.conflicts, st.conflicts],
                          "conflict_resolutions_note": "re-resolution calls: depends on the clean-up schedule (one ordered pass on one GPU, fixed-point iteration over slabs), not on the walk"}
            except Exception as e:   # noqa: BLE001  (rank 0 only: the other ranks wait at the barrier below)
                single = {"error": repr(e)}
            finally:
                if c is not None: c.close()
                sarw_b200.release_cached_memory(local)   # the single-GPU tables (the whole box) go back to the driver before the next record
        dist.barrier()
    if rank != 0:
        return None
    t_dev = sum(times) / 1e3
    alg = (st.trials * B_QUERY + st.nBeads * B_BEAD) / world
    rounds = max(1, st.rounds)
    rec = {"workload": w["name"], "metric": "accepted beads/sec", "unit": "beads/s", "value": beads / t_dev, "beads": st.nBeads, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": 1e3 * t_dev / steps, "ms_per_step_min": min(times), "ms_per_step_max": max(times), "scaling": "strong",
           "multi_gpu": (("one melt over all GPUs, box decomposed into z-slabs: each rank holds the cell layers of its slab + halo (rest of the table unbacked), proposes / verifies / "
                          "re-resolves the chains whose tip is in its slab; bead records, conflict lists and re-resolution results exchanged inside the persistent kernel over "
                          "NVLink peer memory (st.release.sys / ld.acquire.sys flags); chain states and coordinates replicated") if slab else
                         ("one melt over all GPUs: replicated state, partitioned proposals + verification, records and conflict lists exchanged inside the persistent "
                          "kernel over NVLink peer memory (st.release.sys / ld.acquire.sys flags)")) if world > 1 else "single GPU",
           "cell_table_bytes_rank0": cell_bytes,
           "rounds": st.rounds, "trials_per_bead": st.trials / max(1, st.nBeads), "conflicts": st.conflicts, "overflow_beads": st.overflowBeads,
           "all_ranks_agree": bool(agree), "walk_checksum": [int(x) for x in sums[0]],
           "matches_single_gpu": None if single is None or "error" in single else bool(single["checksum_equal"] and single["invariants_equal"]),
           "single_gpu": single, "speedup_vs_single_gpu": None if single is None or "error" in single else single["ms"] / times[-1],
           "cycles_per_round_rank0": {"propose": st.cycPropose / rounds, "verify": st.cycVerify / rounds, "cleanup": st.cycCleanup / rounds,
                                      "sub": {n: st.cycSub[k] / rounds for k, n in enumerate(("propose_loop", "barrier_after_propose", "peer_handshake", "barrier_after_handshake",
                                                                                                "apply_remote", "barrier_after_apply"))}},
           "clocks": clocks,
           "roofline": {"bound": "hbm", "achieved": alg / (times[-1] / 1e3) / 1e9, "peak": env.hbm_peak, "unit": "GB/s", "frac": alg / (times[-1] / 1e3) / 1e9 / env.hbm_peak,
                        "traffic": None, "kernel": "sarw::grow_kernel<8,true> (persistent cooperative launch per GPU)", "peak_source": env.peak_src,
                        "algorithmic_bytes": f"(trials*{B_QUERY:.0f} + beads*{B_BEAD:.0f}) / N GPUs, per GPU"}}
    rec.update(extra)
    return rec


def verify_large_walk(c, w, box_samples, rdf):
    """checks of a walk too large to download (1e9 beads are 48 GB of state vectors): brute-force verifier on sampled sub-boxes
    (sarw_download_box), and the RDF of the whole melt (sarw_rdf, scripts/rdf.py math) against the reference's C2 melts"""
    import numpy as np
    import stats
    out = {}
    L = w["boxSize"][0]
    if box_samples:
        rng = np.random.default_rng(12345)
        side = min(100.0, L / 2)
        samples = []
        for k in range(box_samples):
            lo = rng.random(3) * (L - side)
            xyz, ids, prev, cid = c.download_box(tuple(lo), tuple(lo + side))
            v = stats.verify_box_sample(xyz, ids, prev, w["minDist"])
            v.update(lo=[float(x) for x in lo], side=side, density_g_cc=len(ids) * 14.0 / 6.022e23 / (side ** 3 * 1e-24))   # sarw.h:4-6 constants
            samples.append(v)
        out["box_samples"] = samples
        out["minDist_violations"] = int(sum(v["violations"] for v in samples))
    if rdf:
        t0 = time.perf_counter()
        c.release_lookup()
        intra, inter = c.rdf(2.0, 0.5, 20)
        st = c.stats()
        edges = 2.0 + 0.5 * np.arange(21)
        shell = 4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
        N = st.nBeads
        ideal = N * (N - 1) / (L ** 3) * shell
        g = (intra.astype(np.float64) + inter.astype(np.float64)) / ideal
        rec = {"seconds": time.perf_counter() - t0, "r": [float(x) for x in 0.5 * (edges[1:] + edges[:-1])], "g_total": [float(x) for x in g]}
        ref_path = os.path.join(ROOT, "tests", "golden", "stats_c2_reference_mt.npz")
        if os.path.exists(ref_path):
            ref = np.load(ref_path)
            gr = ref["gIntra"] + ref["gInter"]                      # 12 melts of BASELINE configs[1] grown by the unmodified reference (native RNG)
            mean, se = gr.mean(0), gr.std(0, ddof=1) / np.sqrt(len(gr))
            tol = 4 * se + 0.01 * np.abs(mean)
            rec.update(reference_mean=[float(x) for x in mean], tolerance=[float(x) for x in tol], max_dev_over_tol=float(np.max(np.abs(g - mean) / tol)),
                       within_tolerance=bool(np.all(np.abs(g - mean) <= tol)),
                       tolerance_def="|g - mean_ref| <= 4 * standard error of the 12 reference melts + 1 % of the reference mean, per 0.5 A bin from 2 to 12 A")
        out["rdf"] = rec
    return out


def run_one_melt(args, w):
    import torch
    import sarw_b200
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    env = Env(sarw_b200, torch, local)
    big = w["nChains"] >= 500000
    rec = one_melt_record(env, dist, rank, world, w, args.warmup, args.steps, single_gpu_check=not args.no_single_check,
                          box_samples=3 if big or args.verify else 0, rdf=args.rdf)
    if rank == 0:
        line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "multi_gpu": rec["multi_gpu"], "l2": "tables far larger than L2"},
                "e2e": None, "gpu_launches": int(args.steps * (1 + 4)), "clocks": rec["clocks"], "roofline": rec["roofline"], "one_melt": rec,
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
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary records (C3 / C4 / find_batch at N = 1, one_melt / C5 at N > 1)")
    ap.add_argument("--no-single-check", action="store_true", help="one-melt workloads: skip the single-GPU re-run of the melt on rank 0")
    ap.add_argument("--verify", action="store_true", help="one-melt workloads: brute-force verifier on sampled sub-boxes")
    ap.add_argument("--rdf", action="store_true", help="one-melt workloads: RDF of the whole melt against the reference's C2 melts")
    args = ap.parse_args()
    if args.workload in ONE_MELT:
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the one-melt strong-scaling workloads have no CPU arm (hours per melt on one core)"}))
            return
        return run_one_melt(args, ONE_MELT[args.workload])
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

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
        import datetime
        # a collective that one rank never joins ends after four minutes instead of the default ten
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    env = Env(sarw_b200, torch, local)
    stream = env.stream

    # A melt of <= 256 chains grows in ONE thread-block cluster (13 of 148 SMs for configs[1]); a step is therefore a BATCH of M
    # independent melts grown at the same time on M streams, M = sarw_concurrent_capacity (7 on B200). M = 1 for the large workloads.
    probe = make_ctx(env, w, 1, stream)
    M = args.melts if args.melts > 0 else probe.concurrent_capacity()
    probe.close()
    total = args.warmup + args.steps
    seeds = rank_seeds(rank, total * M)

    # clocks / throttle reasons are sampled by rank 0 only, from before the warm-up (nvidia-smi takes a moment to start) to the end of the
    # timed region
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    if dist: dist.barrier()
    step_ms, acc = measure_device(env, w, M, seeds, args.warmup, args.steps)
    if dist: dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sum(step_ms) / 1e3

    # one melt at a time on an otherwise idle GPU (latency of a melt)
    single_ms, single_beads = [], 0
    if M > 1:
        s_ms, s_acc = measure_device(env, w, 1, [900_000 + rank * 1000 + k for k in range(total)], args.warmup, args.steps)
        single_ms, single_beads = s_ms, s_acc["beads"]

    e2e_times, e2e = measure_e2e(env, w, M, seeds, args.warmup, args.steps)

    # -------- aggregate over ranks (max time, summed work)
    t_e2e = sum(e2e_times)
    t_dev, t_e2e, beads_all, e2e_beads_all = aggregate_ranks(dist, "cuda", t_dev, t_e2e, acc["beads"], e2e["beads"])

    line = None
    if rank == 0:
        line = {
            "metric": "accepted beads/sec", "value": beads_all / t_dev, "unit": "beads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"], "melts_per_step": M,
                       "batch": (f"{M} independent melts of the workload grown at the same time on {M} CUDA streams driven by one host thread (a melt of "
                                 f"{w['nChains']} chains occupies one thread-block cluster, {M} clusters fit the device)") if M > 1 else "one melt per step",
                       "seeds": "distinct Philox key per melt, step and rank", "l2": "256 MB flush write between timed steps",
                       "multi_gpu": "independent melts per rank, no data-path collective (the record one_melt is ONE melt over all GPUs)" if world > 1 else "single GPU",
                       "rounds_per_melt": acc["rounds"] / acc["melts"], "trials_per_bead": acc["trials"] / max(1, acc["beads"]), "conflicts_per_melt": acc["conflicts"] / acc["melts"]},
            "e2e": {"value": e2e_beads_all / t_e2e, "unit": "beads/s", "h2d_bytes_per_step": e2e["h2d"] // args.steps, "d2h_bytes_per_step": e2e["d2h"] // args.steps,
                    "ms_per_step": 1e3 * t_e2e / args.steps, "ms_per_step_median": 1e3 * statistics.median(e2e_times), "ms_per_step_max": 1e3 * max(e2e_times),
                    "includes": "per melt: sarw_create (alloc + cell-table init), seeding, growth, terminate, sarw_download to pinned host memory, sarw_destroy"},
            "gpu_launches": int(acc["launches"]),
            "clocks": clocks,
            "roofline": roofline_record(env, w, args.workload, acc, M),
        }
        if single_ms:
            line["single_melt"] = {"ms_per_melt": statistics.median(single_ms), "beads_per_s": single_beads / (sum(single_ms) / 1e3),
                                   "note": "one melt at a time on an otherwise idle GPU, device time (CUDA events)"}
    if not args.no_extra:
        if world == 1:
            line["workloads"] = {}
            for key in ("C3", "C4"):
                if key == args.workload: continue
                try:
                    line["workloads"][key] = workload_record(env, key)
                except Exception as e:   # secondary records must never take the headline down
                    line["workloads"][key] = {"error": repr(e)}
            line["extra"] = extra_measurements(env)
        else:
            setup_host_agreement(dist, rank, world)
            # At N = 8 first BASELINE configs[4], ONE melt of 1e9 beads (13.7 GB of cell layers per rank), while no rank carries anything
            # from an earlier record: in round 2's run it came after rank 0's single-GPU comparison melt and rank 0 could not back its slab.
            if world == 8 and not os.environ.get("SARW_BENCH_NO_C5"):
                sarw_b200.release_cached_memory(local)
                try:
                    rec = one_melt_record(env, dist, rank, world, ONE_MELT["C5"], 1, 2, single_gpu_check=False, box_samples=3, rdf=False)
                except Exception as e:
                    rec = {"error": repr(e)}
                if rank == 0: line["C5"] = rec
                sarw_b200.release_cached_memory(local)
            # ONE melt of 5e7*N beads over all N GPUs, checked against the other ranks and against the single-GPU walk
            try:
                if _AGREE is not None and _AGREE.poisoned:
                    raise RuntimeError("skipped: a rank lost its CUDA context in the record before")
                rec = one_melt_record(env, dist, rank, world, melt_workload(50000 * world, "grown by all N GPUs together"), 1, 3)
            except Exception as e:
                rec = {"error": repr(e)}
            if rank == 0: line["one_melt"] = rec
            # the same melt with round 1's scheme for comparison: every table on every rank (so it only fits smaller boxes), chain i proposed
            # by rank i % N, every rank applies every bead; same seeds, so the walk must be the one of the record above
            try:
                if _AGREE is not None and _AGREE.poisoned:
                    raise RuntimeError("skipped: a rank lost its CUDA context in a record before")
                sarw_b200.release_cached_memory(local)
                rec2 = one_melt_record(env, dist, rank, world, melt_workload(50000 * world, "grown by all N GPUs together"), 1, 3, single_gpu_check=False, replicated=True)
                if rank == 0 and isinstance(rec, dict) and "walk_checksum" in rec:
                    rec2["same_walk_as_one_melt"] = rec2["walk_checksum"] == rec["walk_checksum"]
            except Exception as e:
                rec2 = {"error": repr(e)}
            if rank == 0: line["one_melt_replicated_tables"] = rec2
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            cores = host_cores()
            r = reference_throughput(w, cores, 777)
            if r is not None:
                keyed = reference_throughput(w, cores, 888, rng="keyed")
                line["cpu_baseline"] = {"value": r["value"], "unit": "beads/s", "cores": cores, "kind": "reference",
                                        "sample": f"{cores} concurrent single-thread instances of the full workload, unmodified reference (oracle/_ref), "
                                                  f"seeding + growth timed; slowest instance {r['slowest']:.2f} s", "per_core_beads_per_s": r["per_core"],
                                        "trials_per_bead": r["trials_per_bead"], "keyed_rng_beads_per_s": keyed["value"],
                                        "keyed_rng_note": "the same unmodified reference with Random::seed reduced to a counter set-up (no mt19937_64 re-initialisation per "
                                                          "trial): an upper bound of the reference's speed whatever JDFTx's Random::seed costs"}
                for key in ("C3",):
                    if "workloads" in line and key in line["workloads"] and "error" not in line["workloads"][key]:
                        # the reference's cost per bead depends on density, chain length and maxTrials (trials per bead), not on the box: a bounded
                        # sample of the big melt itself would have to be its FIRST rounds, which are ~10x cheaper than the average round (no dead chains yet)
                        line["workloads"][key]["cpu_baseline"] = {
                            "value": r["value"], "unit": "beads/s", "cores": cores, "kind": "reference", "per_core_beads_per_s": r["per_core"],
                            "sample": f"{cores} concurrent single-thread instances of BASELINE configs[1] (same density, chain length, minDist and maxTrials, hence the same "
                                      f"trials per bead: {r['trials_per_bead']:.1f}); a whole configs[2] melt takes one core 253 s (3.95e4 beads/s, SURVEY section 6) - its table no "
                                      f"longer fits the CPU caches"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "beads/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        print(json.dumps(line))
    if dist:
        if _AGREE is not None and _AGREE.poisoned:
            # a rank's CUDA context is gone: no device collective would return; meet on the host and leave without tearing NCCL down
            sys.stdout.flush()
            try:
                _AGREE.barrier(30)
            except Exception:   # noqa: BLE001
                pass
            os._exit(0)
        dist.barrier(); dist.destroy_process_group()


def extra_measurements(env):
    """Secondary number: the overlap-query kernel (K5) on a finished BASELINE configs[2] melt."""
    torch, stream = env.torch, env.stream
    out = {}
    w = WORKLOADS["C3"]
    try:
        c = make_ctx(env, w, 4242, stream)
        c.initiate(); st = c.grow(0, -1); c.terminate()
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
            env.l2_flush(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); c.find_device(q.data_ptr(), None, n, hit.data_ptr()); e1.record(stream); e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = statistics.median(ms) / 1e3
        bq = 134.0 + 24 + 1   # SURVEY §8d B_q at full density + candidate read + result write
        out["find_batch"] = {"queries": n, "table_beads": st.nBeads, "ms": t * 1e3, "queries_per_s": n / t, "algorithmic_GBps": n * bq / t / 1e9,
                             "frac_of_hbm_peak": n * bq / t / 1e9 / env.hbm_peak, "hit_fraction": float(hit.float().mean().item()),
                             "bytes_per_query": bq}
        c.close()
    except Exception as e:   # secondary numbers must never take the headline down
        out["error"] = repr(e)
    return out


if __name__ == "__main__":
    main()
