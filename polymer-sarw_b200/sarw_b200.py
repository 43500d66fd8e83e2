"""ctypes binding of libsarw_b200.so — the same C ABI (include/sarw_abi.h) the C++ `sarw` host links against.

Mirrors the reference's two classes for the hot path so parity tests read like the reference's own calls:

    PeriodicLookup(boxSize, thresh).addPoint / .find        PeriodicLookup.h:26,37,51   (batched here)
    SARW(nChains, boxSize, minDist) .initiateChain() .propagateChain() .terminateChain()   sarw.h:88-118

There is NO fallback: if the shared library is missing or no CUDA device is present, construction raises.
The oracle under oracle/ is never imported from here.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SARW_B200_LIB") or os.path.join(_HERE, "lib", "libsarw_b200.so")   # override: A/B builds of the same library

GRAFT = {"file": 0, "z": 1, "zCenter": 2}
ORDER = {"roundRobin": 0, "grafted": 1}
BOUNDARY = {"ppp": 0, "ppf": 1, "pfp": 2, "fpp": 3}

ABI_SYMBOLS = [
    "sarw_create", "sarw_destroy", "sarw_last_error", "sarw_set_stream", "sarw_add_obstacles", "sarw_set_grafts",
    "sarw_initiate", "sarw_propagate_round", "sarw_grow", "sarw_terminate", "sarw_get_stats", "sarw_target_count",
    "sarw_download", "sarw_chain_lengths", "sarw_round_counts", "sarw_lookup_add_points", "sarw_lookup_find_batch",
    "sarw_lookup_find_batch_device", "sarw_lookup_size", "sarw_rdf", "sarw_chain_shapes", "sarw_concurrent_capacity", "sarw_release_cached_memory",
    "sarw_mr_setup", "sarw_mr_more", "sarw_mr_propose", "sarw_mr_finish", "sarw_mr_commit",
    "sarw_mr_peer_export", "sarw_mr_peer_connect", "sarw_mr_grow_peer", "sarw_mr_peer_local_base", "sarw_mr_peer_connect_local",
    "sarw_grow_async", "sarw_wait", "sarw_walk_checksum", "sarw_download_box", "sarw_release_lookup", "sarw_cell_table_bytes", "sarw_mr_grow_local", "sarw_device_count", "sarw_slab_layers", "sarw_slab_owner",
]


class SarwError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("nChains", C.c_int64), ("boxSize", C.c_double * 3), ("minDist", C.c_double),
                ("targetMassDensity", C.c_double), ("nGraftChains", C.c_int64), ("maxTrials", C.c_int32),
                ("graftLoc", C.c_int32), ("growthOrder", C.c_int32), ("boundary", C.c_int32),
                ("growthBias", C.c_double * 3), ("seed", C.c_uint64), ("device", C.c_int32),
                ("cellCapacity", C.c_int32), ("cellEdge", C.c_double), ("reserveRounds", C.c_int64),
                ("launchMode", C.c_int32), ("maxBlocks", C.c_int32),
                ("slabRank", C.c_int32), ("slabRanks", C.c_int32), ("slabHalo", C.c_int32), ("reserved0", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("nBeads", C.c_int64), ("nBonds", C.c_int64), ("target", C.c_int64), ("rounds", C.c_int64),
                ("initiated", C.c_int32), ("lastProgressed", C.c_int32), ("trials", C.c_uint64),
                ("seedTrials", C.c_uint64), ("queries", C.c_uint64), ("conflicts", C.c_uint64),
                ("overflowBeads", C.c_uint64), ("seedMs", C.c_double), ("growMs", C.c_double),
                ("growLaunches", C.c_int64), ("seedLaunches", C.c_int64), ("cycPropose", C.c_uint64),
                ("cycVerify", C.c_uint64), ("cycCleanup", C.c_uint64), ("cleanupRounds", C.c_int64),
                ("deadChains", C.c_int64), ("stageFallbacks", C.c_uint64), ("arcFallbacks", C.c_uint64),
                ("cycSub", C.c_uint64 * 8)]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "cycSub" else getattr(self, n)) for n, _ in self._fields_}


_lib = None


def load_library(path=LIB_PATH):
    """dlopen libsarw_b200.so and declare the ABI; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise SarwError(f"{path} is missing: build it with `make lib` (python __graft_entry__.py); there is no CPU fallback")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.sarw_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.sarw_destroy.argtypes = [vp]; lib.sarw_destroy.restype = None
    lib.sarw_last_error.argtypes = [vp]; lib.sarw_last_error.restype = C.c_char_p
    lib.sarw_set_stream.argtypes = [vp, vp]
    lib.sarw_add_obstacles.argtypes = [vp, vp, C.c_int64]
    lib.sarw_set_grafts.argtypes = [vp, vp, C.c_int64]
    lib.sarw_initiate.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.sarw_propagate_round.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.sarw_grow.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(Stats)]
    lib.sarw_terminate.argtypes = [vp]
    lib.sarw_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.sarw_target_count.argtypes = [C.POINTER(C.c_double), C.c_double]; lib.sarw_target_count.restype = C.c_int64
    lib.sarw_download.argtypes = [vp, vp, vp, vp, vp]
    lib.sarw_chain_lengths.argtypes = [vp, vp]
    lib.sarw_rdf.argtypes = [vp, C.c_double, C.c_double, C.c_int32, vp, vp]
    lib.sarw_chain_shapes.argtypes = [vp, vp, vp]
    lib.sarw_concurrent_capacity.argtypes = [vp]
    lib.sarw_release_cached_memory.argtypes = [C.c_int32]
    lib.sarw_concurrent_capacity.restype = C.c_int32
    lib.sarw_round_counts.argtypes = [vp, vp, C.c_int64]
    lib.sarw_lookup_add_points.argtypes = [vp, vp, C.c_int64]
    lib.sarw_lookup_find_batch.argtypes = [vp, vp, vp, C.c_int64, vp]
    lib.sarw_lookup_find_batch_device.argtypes = [vp, vp, vp, C.c_int64, vp]
    lib.sarw_lookup_size.argtypes = [vp]; lib.sarw_lookup_size.restype = C.c_int64
    lib.sarw_cell_table_bytes.argtypes = [vp]; lib.sarw_cell_table_bytes.restype = C.c_int64
    lib.sarw_slab_layers.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]; lib.sarw_slab_layers.restype = C.c_int
    lib.sarw_slab_owner.argtypes = [C.c_int32, C.c_int32, C.c_int32]; lib.sarw_slab_owner.restype = C.c_int32
    lib.sarw_mr_setup.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.sarw_mr_more.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.sarw_mr_propose.argtypes = [vp, vp]
    lib.sarw_mr_finish.argtypes = [vp, vp, C.POINTER(C.c_int32)]
    lib.sarw_mr_commit.argtypes = [vp]
    lib.sarw_mr_peer_export.argtypes = [vp, vp]
    lib.sarw_mr_peer_connect.argtypes = [vp, vp]
    lib.sarw_mr_grow_peer.argtypes = [vp, C.c_int64, C.POINTER(Stats)]
    lib.sarw_mr_peer_local_base.argtypes = [vp, C.POINTER(vp)]
    lib.sarw_mr_peer_connect_local.argtypes = [vp, C.POINTER(vp)]
    lib.sarw_grow_async.argtypes = [vp, C.c_int64, C.c_int64]
    lib.sarw_wait.argtypes = [vp, C.POINTER(Stats)]
    lib.sarw_walk_checksum.argtypes = [vp, vp]
    lib.sarw_download_box.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(C.c_int64), vp, vp, vp, vp]
    lib.sarw_release_lookup.argtypes = [vp]
    _lib = lib
    return lib


def target_count(boxSize, targetMassDensity):
    b = (C.c_double * 3)(*boxSize)
    return load_library().sarw_target_count(b, targetMassDensity)


def release_cached_memory(device=-1):
    """hand the parked tables of destroyed contexts back to the driver"""
    return load_library().sarw_release_cached_memory(device)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)


class Context:
    """Owns one sarw_ctx (one GPU). Thin: every method is one ABI call."""

    def __init__(self, nChains=1, boxSize=(10, 10, 10), minDist=2.0, targetMassDensity=0.92, maxTrials=100,
                 boundary="ppp", nGraftChains=0, graftLoc="file", growthOrder="roundRobin", growthBias=(0, 0, 0),
                 seed=1, device=-1, cellCapacity=0, cellEdge=0.0, reserveRounds=0, launchMode=0, maxBlocks=0,
                 slabRank=0, slabRanks=0, slabHalo=0):
        self.lib = load_library()
        p = Params()
        p.nChains = nChains; p.boxSize = (C.c_double * 3)(*boxSize); p.minDist = minDist
        p.targetMassDensity = targetMassDensity; p.nGraftChains = nGraftChains; p.maxTrials = maxTrials
        p.graftLoc = GRAFT.get(graftLoc, 3); p.growthOrder = ORDER.get(growthOrder, 0)
        p.boundary = BOUNDARY.get(boundary, 0)   # only exact ppf/pfp/fpp act, anything else == ppp (sarw.cpp:201-203)
        p.growthBias = (C.c_double * 3)(*growthBias); p.seed = seed; p.device = device
        p.cellCapacity = cellCapacity; p.cellEdge = cellEdge; p.reserveRounds = reserveRounds; p.launchMode = launchMode; p.maxBlocks = maxBlocks
        p.slabRank = slabRank; p.slabRanks = slabRanks; p.slabHalo = slabHalo; p.reserved0 = 0
        self.params = p
        self.nChains = nChains
        h = C.c_void_p()
        if self.lib.sarw_create(C.byref(p), C.byref(h)) != 0:
            raise SarwError(self.lib.sarw_last_error(None).decode())
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise SarwError(self.lib.sarw_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.sarw_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- lookup (PeriodicLookup.h)
    def add_points(self, xyz):
        a = _f64(xyz)
        self._ck(self.lib.sarw_lookup_add_points(self.h, a.ctypes.data, a.shape[0]))

    def add_obstacles(self, xyz):
        a = _f64(xyz)
        self._ck(self.lib.sarw_add_obstacles(self.h, a.ctypes.data, a.shape[0]))

    def find(self, xyz, ignore=None):
        a = _f64(xyz)
        ig = None if ignore is None else np.ascontiguousarray(ignore, dtype=np.int64)
        out = np.zeros(a.shape[0], np.uint8)
        self._ck(self.lib.sarw_lookup_find_batch(self.h, a.ctypes.data, ig.ctypes.data if ig is not None else None,
                                                 a.shape[0], out.ctypes.data))
        return out

    def find_device(self, d_xyz_ptr, d_ignore_ptr, n, d_hit_ptr):
        self._ck(self.lib.sarw_lookup_find_batch_device(self.h, d_xyz_ptr, d_ignore_ptr, n, d_hit_ptr))

    def lookup_size(self):
        return self.lib.sarw_lookup_size(self.h)

    def cell_table_bytes(self):
        return self.lib.sarw_cell_table_bytes(self.h)

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.sarw_set_stream(self.h, cuda_stream_ptr))

    # ---- walk (sarw.cpp)
    def set_grafts(self, xyz):
        a = _f64(xyz)
        self._ck(self.lib.sarw_set_grafts(self.h, a.ctypes.data, a.shape[0]))

    def initiate(self):
        ok = C.c_int32()
        self._ck(self.lib.sarw_initiate(self.h, C.byref(ok)))
        return bool(ok.value)

    def propagate_round(self):
        pr = C.c_int32()
        self._ck(self.lib.sarw_propagate_round(self.h, C.byref(pr)))
        return bool(pr.value)

    def grow(self, target=0, maxRounds=-1):
        s = Stats()
        self._ck(self.lib.sarw_grow(self.h, target, maxRounds, C.byref(s)))
        return s

    def grow_async(self, target=0, maxRounds=-1):
        """seeding (unless initiate() was called) + the growth loop enqueued on the context's stream; returns at once"""
        self._ck(self.lib.sarw_grow_async(self.h, target, maxRounds))

    def wait(self):
        s = Stats()
        self._ck(self.lib.sarw_wait(self.h, C.byref(s)))
        return s

    # ---- one melt over several GPUs (replicated state, partitioned proposals; include/sarw_abi.h "multi-rank")
    def mr_setup(self, rank, nranks):
        k, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.sarw_mr_setup(self.h, rank, nranks, C.byref(k), C.byref(b)))
        return k.value, b.value

    def mr_more(self):
        m = C.c_int32()
        self._ck(self.lib.sarw_mr_more(self.h, C.byref(m)))
        return bool(m.value)

    def mr_propose(self, d_local_ptr):
        self._ck(self.lib.sarw_mr_propose(self.h, d_local_ptr))

    def mr_finish(self, d_all_ptr):
        m = C.c_int32()
        self._ck(self.lib.sarw_mr_finish(self.h, d_all_ptr, C.byref(m)))
        return bool(m.value)

    def mr_commit(self):
        self._ck(self.lib.sarw_mr_commit(self.h))

    def mr_peer_export(self):
        buf = (C.c_ubyte * 64)()
        self._ck(self.lib.sarw_mr_peer_export(self.h, buf))
        return bytes(buf)

    def mr_peer_connect(self, all_handles):
        raw = b"".join(all_handles)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        self._ck(self.lib.sarw_mr_peer_connect(self.h, buf))

    def mr_peer_local_base(self):
        b = C.c_void_p()
        self._ck(self.lib.sarw_mr_peer_local_base(self.h, C.byref(b)))
        return b.value

    def mr_peer_connect_local(self, bases):
        arr = (C.c_void_p * len(bases))(*bases)
        self._ck(self.lib.sarw_mr_peer_connect_local(self.h, arr))

    def mr_grow_peer(self, maxRounds=-1):
        s = Stats()
        self._ck(self.lib.sarw_mr_grow_peer(self.h, maxRounds, C.byref(s)))
        return s

    def terminate(self):
        self._ck(self.lib.sarw_terminate(self.h))

    def stats(self):
        s = Stats()
        self._ck(self.lib.sarw_get_stats(self.h, C.byref(s)))
        return s

    def download(self, out=None):
        """The public state vectors (sarw.h:77-83). `out` may hold preallocated host buffers (e.g. views of pinned torch tensors)
        xyz (>= N x 3 f64), chainID, type (>= N i32), bonds (>= N x 2 i64); views of the right length are returned."""
        s = self.stats()
        if out is None:
            xyz = np.empty((s.nBeads, 3), np.float64); chainID = np.empty(s.nBeads, np.int32)
            typ = np.empty(s.nBeads, np.int32); bonds = np.empty((s.nBonds, 2), np.int64)
        else:
            xyz, chainID, typ, bonds = out["xyz"][:s.nBeads], out["chainID"][:s.nBeads], out["type"][:s.nBeads], out["bonds"][:s.nBonds]
        self._ck(self.lib.sarw_download(self.h, xyz.ctypes.data, chainID.ctypes.data, typ.ctypes.data, bonds.ctypes.data))
        return xyz, chainID, typ, bonds

    def walk_checksum(self):
        """four 64-bit words that identify the walk (every coordinate and bond), computed on the device"""
        out = np.zeros(4, np.uint64)
        self._ck(self.lib.sarw_walk_checksum(self.h, out.ctypes.data))
        return tuple(int(x) for x in out)

    def download_box(self, lo, hi, cap=1 << 22):
        """beads with wrapped position in [lo, hi): (wrapped xyz, slot ids, slot ids of the bonded predecessors, chainID)"""
        lo_ = (C.c_double * 3)(*lo); hi_ = (C.c_double * 3)(*hi)
        xyz = np.empty((cap, 3), np.float64); ids = np.empty(cap, np.int64); prev = np.empty(cap, np.int64); cid = np.empty(cap, np.int32)
        n = C.c_int64()
        self._ck(self.lib.sarw_download_box(self.h, lo_, hi_, cap, C.byref(n), xyz.ctypes.data, ids.ctypes.data, prev.ctypes.data, cid.ctypes.data))
        if n.value > cap:
            return self.download_box(lo, hi, int(n.value))
        m = n.value
        return xyz[:m], ids[:m], prev[:m], cid[:m]

    def release_lookup(self):
        self._ck(self.lib.sarw_release_lookup(self.h))

    def chain_lengths(self):
        out = np.zeros(self.nChains, np.int32)
        self._ck(self.lib.sarw_chain_lengths(self.h, out.ctypes.data))
        return out

    def concurrent_capacity(self):
        return int(self.lib.sarw_concurrent_capacity(self.h))

    def rdf(self, rMin=2.0, dr=0.5, nBins=20):
        """scripts/rdf.py:13-20,39-57 on the device: (intra, inter) integer pair counts per bin, bins rMin + k*dr"""
        intra = np.zeros(nBins, np.uint64); inter = np.zeros(nBins, np.uint64)
        self._ck(self.lib.sarw_rdf(self.h, rMin, dr, nBins, intra.ctypes.data, inter.ctypes.data))
        return intra, inter

    def chain_shapes(self):
        """(end-to-end distance^2, radius of gyration^2) per chain"""
        ree2 = np.zeros(self.nChains); rg2 = np.zeros(self.nChains)
        self._ck(self.lib.sarw_chain_shapes(self.h, ree2.ctypes.data, rg2.ctypes.data))
        return ree2, rg2

    def round_counts(self):
        s = self.stats()
        out = np.zeros(s.rounds, np.int64)
        self._ck(self.lib.sarw_round_counts(self.h, out.ctypes.data, s.rounds))
        return out


class PeriodicLookup:
    """Mirror of the reference class (PeriodicLookup.h:9-76), batched: addPoint/find take (n,3) arrays."""

    def __init__(self, boxSize, thresh, **kw):
        self.ctx = Context(nChains=1, boxSize=boxSize, minDist=thresh, **kw)

    def addPoint(self, pos):
        self.ctx.add_points(pos)

    def find(self, pos, ignoreIndex=None):
        return self.ctx.find(pos, ignoreIndex)

    def close(self):
        self.ctx.close()


class SARW:
    """Mirror of the reference's SARW object for the hot path (sarw.h:59-136, sarw.cpp:220-368)."""

    def __init__(self, nChains, boxSize, minDist, **kw):
        self.ctx = Context(nChains=nChains, boxSize=boxSize, minDist=minDist, **kw)
        self.nChains = nChains
        self.boxSize = tuple(boxSize)

    def readGrafts(self, xyz):
        self.ctx.set_grafts(xyz)

    def addObstacle(self, xyz):
        self.ctx.add_obstacles(xyz)

    def initiateChain(self):
        return self.ctx.initiate()

    def propagateChain(self):
        return self.ctx.propagate_round()

    def terminateChain(self):
        self.ctx.terminate()

    def actualCount(self):
        return self.ctx.stats().nBeads

    def targetCount(self):
        return self.ctx.stats().target

    def run(self, maxRounds=-1):
        """main()'s sequence sarw.cpp:66-79: initiate, grow to the target, terminate. Returns the final stats."""
        self.ctx.initiate()
        st = self.ctx.grow(0, maxRounds)
        self.ctx.terminate()
        return st

    def close(self):
        self.ctx.close()


def grow_multirank(ctx, dist, rank, nranks, device="cuda"):
    """main()'s growth loop (sarw.cpp:67-78) for ONE melt over `nranks` GPUs: every rank holds a context created with the same
    parameters; proposals are partitioned, one all-gather of 32-byte records per round (torch.distributed, NCCL on GPUs).
    `dist` is torch.distributed (initialised) or an object with all_gather_into_tensor(out, inp). Returns rounds run."""
    import torch
    K, B = ctx.mr_setup(rank, nranks)
    local = torch.zeros(K * B, dtype=torch.uint8, device=device)
    allrec = torch.zeros(nranks * K * B, dtype=torch.uint8, device=device)
    rounds = 0
    more = ctx.mr_more()
    while more:
        ctx.mr_propose(local.data_ptr())
        dist.all_gather_into_tensor(allrec, local)
        more = ctx.mr_finish(allrec.data_ptr())
        rounds += 1
    ctx.mr_commit()
    return rounds


def mr_chain_slot(i, nranks, K):
    """Where chain i's record sits after the all-gather (rank-major): (i % nranks) * K + i // nranks."""
    return (i % nranks) * K + i // nranks


def slab_layers(nLayers, ranks, rank, halo):
    """(slab first layer, slab end, window first layer, window layers) of one rank of a decomposed box; needs no GPU"""
    out = (C.c_int32 * 4)()
    if load_library().sarw_slab_layers(nLayers, ranks, rank, halo, out) != 0:
        raise SarwError("sarw_slab_layers: bad arguments")
    return tuple(out)


def slab_owner(nLayers, ranks, layer):
    return load_library().sarw_slab_owner(nLayers, ranks, layer)


def connect_peers(ctx, dist, rank, nranks):
    """One-time set-up of the peer-memory exchange: torch.distributed is used once, to swap the 64-byte CUDA IPC handles."""
    ctx.mr_setup(rank, nranks)
    mine = ctx.mr_peer_export()
    handles = [None] * nranks
    dist.all_gather_object(handles, mine)
    ctx.mr_peer_connect(handles)
    dist.barrier()


def connect_peers_local(ctxs):
    """All ranks' contexts live in this process (ranks side by side on one GPU, each created with maxBlocks so that their cooperative
    grids are resident together): exchange buffers are shared as plain device pointers."""
    R = len(ctxs)
    for r, c in enumerate(ctxs):
        c.mr_setup(r, R)
        c.mr_peer_export()
    bases = [c.mr_peer_local_base() for c in ctxs]
    for c in ctxs:
        c.mr_peer_connect_local(bases)


def grow_peer_local(ctxs):
    """sarw_mr_grow_peer on every rank at the same time, one host thread per rank (the kernels meet inside the launch)."""
    import threading
    out, errs = [None] * len(ctxs), []

    def work(r):
        try:
            out[r] = ctxs[r].mr_grow_peer(-1)
        except Exception as e:   # noqa: BLE001
            errs.append(f"rank {r}: {e!r}")

    th = [threading.Thread(target=work, args=(r,)) for r in range(len(ctxs))]
    for t in th: t.start()
    for t in th: t.join()
    if errs:
        raise SarwError("; ".join(errs))
    return out


def grow_multirank_peer(ctx, dist, rank, nranks, connected=False):
    """Same walk, exchange inside the kernel over NVLink peer memory (no host round trip, no NCCL call per round)."""
    if not connected:
        connect_peers(ctx, dist, rank, nranks)
    return ctx.mr_grow_peer(-1).rounds
