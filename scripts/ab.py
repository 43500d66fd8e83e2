"""A/B timing probe (not a bench line): C2 one melt, C2 batch of M melts driven by ONE host thread (sarw_grow_async / sarw_wait),
C3 one melt, K5 on the C3 melt. Select the library with SARW_B200_LIB=... ; python scripts/ab.py [c2] [batch] [c3] [find]"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sarw_b200, cases
what = sys.argv[1:] or ["c2", "batch", "c3", "find", "big"]
tag = os.path.basename(os.environ.get("SARW_B200_LIB", "libsarw_b200.so"))
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 18, dtype=torch.float32, device="cuda")

def timed(fn):
    flush.fill_(1.0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); r = fn(); e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1), r

if "c2" in what:
    ms = []
    for k in range(6):
        c = sarw_b200.Context(seed=100 + k, **cases.C2); c.set_stream(stream.cuda_stream)
        t, st = timed(lambda: (c.initiate(), c.grow(0, -1))[1]); ms.append(t)
        d = st.as_dict(); c.close()
    print(f"[{tag}] C2 one melt: median {statistics.median(ms[1:]):.3f} ms (min {min(ms[1:]):.3f}); cyc/round A {d['cycPropose']/d['rounds']:.0f} V {d['cycVerify']/d['rounds']:.0f} C {d['cycCleanup']/d['rounds']:.0f} sub {[round(x/d['rounds']) for x in d['cycSub']]}", flush=True)
if "batch" in what:
    probe = sarw_b200.Context(seed=1, **cases.C2); M = probe.concurrent_capacity(); probe.close()
    for MM in sorted({M, M + 2, max(1, M - 2)}):
        streams = [torch.cuda.Stream() for _ in range(MM)]
        ms, beads = [], 0
        for k in range(6):
            ctxs = [sarw_b200.Context(seed=1000 * k + m, **cases.C2) for m in range(MM)]
            for c, s in zip(ctxs, streams): c.set_stream(s.cuda_stream)
            def go():
                for s in streams: s.wait_stream(stream)
                for c in ctxs: c.grow_async(0, -1)
                sts = [c.wait() for c in ctxs]
                for s in streams: stream.wait_stream(s)
                return sts
            t, sts = timed(go)
            if k: ms.append(t); beads += sum(s.nBeads for s in sts)
            for c in ctxs: c.close()
        print(f"[{tag}] C2 batch of {MM} (capacity {M}) from one host thread: median {statistics.median(ms):.3f} ms per step, {beads / sum(ms) * 1e3:.3e} beads/s", flush=True)
if "c3" in what or "find" in what:
    ms = []
    for k in range(3):
        c = sarw_b200.Context(seed=4242 + k, **cases.C3); c.set_stream(stream.cuda_stream)
        t, st = timed(lambda: (c.initiate(), c.grow(0, -1))[1]); ms.append(t)
        d = st.as_dict()
        if k < 2: c.close()
    print(f"[{tag}] C3 one melt: {['%.2f' % x for x in ms]} ms, {d['nBeads'] / min(ms) * 1e3:.3e} beads/s; ovf {d['overflowBeads']} conflicts {d['conflicts']} cyc/round A {d['cycPropose']/d['rounds']:.0f} V {d['cycVerify']/d['rounds']:.0f} C {d['cycCleanup']/d['rounds']:.0f} sub {[round(x/d['rounds']) for x in d['cycSub']]}", flush=True)
    if "find" in what:
        n = 1 << 24
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        q = torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g) * cases.C3["boxSize"][0]
        hit = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(2): c.find_device(q.data_ptr(), None, n, hit.data_ptr())
        ts = [timed(lambda: c.find_device(q.data_ptr(), None, n, hit.data_ptr()))[0] for _ in range(5)]
        os.environ["SARW_FIND_ORDER"] = "0"
        print(f"[{tag}] K5 16.8M candidates on the C3 melt: median {statistics.median(ts):.3f} ms = {n / statistics.median(ts) / 1e6:.2f} Gq/s, hit fraction {hit.float().mean().item():.4f}", flush=True)
    c.close()
if "big" in what:
    ms = []
    for k in range(3):
        c = sarw_b200.Context(seed=77 + k, **cases.BIG); c.set_stream(stream.cuda_stream)
        t, st = timed(lambda: (c.initiate(), c.grow(0, -1))[1]); ms.append(t)
        d = st.as_dict(); c.close()
    print(f"[{tag}] BIG one melt: {['%.2f' % x for x in ms]} ms, {d['nBeads'] / min(ms) * 1e3:.3e} beads/s; conflicts {d['conflicts']} cyc/round A {d['cycPropose']/d['rounds']:.0f} V {d['cycVerify']/d['rounds']:.0f} C {d['cycCleanup']/d['rounds']:.0f} sub {[round(x/d['rounds']) for x in d['cycSub']]}", flush=True)
