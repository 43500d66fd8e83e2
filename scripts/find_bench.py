"""Stand-alone timing of the overlap-query kernel (K5) on a melt-density random point set; used under ncu."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200"))
import numpy as np, torch, sarw_b200

L = float(sys.argv[1]) if len(sys.argv) > 1 else 632.2
npts = int(0.03957 * L ** 3)
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
gran = int(os.environ.get("L2_FETCH", "0"))
if gran:
    import ctypes
    rt = ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.12"))
    torch.zeros(1, device="cuda")
    print("cudaDeviceSetLimit(MaxL2FetchGranularity,", gran, ") ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran)))
    v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print("limit now", v.value)
rng = np.random.default_rng(1)
ctx = sarw_b200.Context(nChains=1, boxSize=(L, L, L), minDist=1.0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
pts = rng.random((npts, 3)) * L
ctx.add_points(pts)
g = torch.Generator(device="cuda"); g.manual_seed(5)
q = torch.rand((nq, 3), dtype=torch.float64, device="cuda", generator=g) * L
if os.environ.get("PRESORT") == "z":
    q = q[torch.argsort(q[:, 2])].contiguous()
elif os.environ.get("PRESORT") == "zy":
    a = 632.2 / 316
    key = torch.floor(q[:, 2] / a) * 1000 + torch.floor(q[:, 1] / (8 * a))
    q = q[torch.argsort(key)].contiguous()
if os.environ.get("PRESORT") == "cell":
    a = L / int(L / 2.002)
    c = torch.floor(q / a).to(torch.int64)
    q = q[torch.argsort((c[:, 2] * 1000 + c[:, 1]) * 1000 + c[:, 0])].contiguous()
if os.environ.get("PRESORT") == "tile":     # 4x4x4-cell tiles, tiles in x-fastest order
    a = L / int(L / 2.002)
    c = torch.floor(q / (4 * a)).to(torch.int64)
    q = q[torch.argsort((c[:, 2] * 1000 + c[:, 1]) * 1000 + c[:, 0])].contiguous()
hit = torch.empty(nq, dtype=torch.uint8, device="cuda")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
ms = []
for i in range(reps):
    flush.fill_(float(i)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.find_device(q.data_ptr(), None, nq, hit.data_ptr()); e1.record(stream); e1.synchronize()
    ms.append(e0.elapsed_time(e1))
t = sorted(ms)[len(ms) // 2] / 1e3
print(f"L={L} points={npts} queries={nq}: {t*1e3:.3f} ms median, {nq/t:.3e} queries/s, hit fraction {hit.float().mean().item():.4f}, "
      f"algorithmic {nq*159/t/1e9:.1f} GB/s (159 B/query)")
