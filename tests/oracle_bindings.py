"""ctypes bindings for the two parity oracles (TEST INFRASTRUCTURE ONLY):

* ``OracleLib``  — oracle/build/liboracle.so, the plain-C restatement (oracle/sarw_oracle.c)
* ``RefLib``     — oracle/_ref/libsarw_ref.so, the UNMODIFIED reference sources behind oracle/ref_harness.cpp

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "build", "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsarw_ref.so")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "sarw_ref")

GRAFT = {"file": 0, "z": 1, "zCenter": 2}
ORDER = {"roundRobin": 0, "grafted": 1}
BOUNDARY = {"ppp": 0, "ppf": 1, "pfp": 2, "fpp": 3}


def build_oracles():
    """make -C oracle (the reference part is skipped by the Makefile when /root/reference is absent)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.DEVNULL)


class OracleParams(C.Structure):
    _fields_ = [("nChains", C.c_int32), ("box", C.c_double * 3), ("minDist", C.c_double),
                ("targetMassDensity", C.c_double), ("nGraftChains", C.c_int32), ("maxTrials", C.c_int32),
                ("graftLoc", C.c_int32), ("growthOrder", C.c_int32), ("boundary", C.c_int32),
                ("growthBias", C.c_double * 3), ("trig", C.c_int32), ("seed", C.c_uint64), ("maxRounds", C.c_int32)]


class RefParams(C.Structure):
    _fields_ = [("nChains", C.c_int32), ("box", C.c_double * 3), ("minDist", C.c_double),
                ("targetMassDensity", C.c_double), ("nGraftChains", C.c_int32), ("maxTrials", C.c_int32),
                ("graftLoc", C.c_char * 16), ("growthOrder", C.c_char * 16), ("boundary", C.c_char * 8),
                ("polymer", C.c_char * 64), ("rngMode", C.c_int32), ("seed", C.c_uint64), ("maxRounds", C.c_int32)]


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _xyz(a):
    if a is None:
        return None, 0
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)
    return a, a.shape[0]


class WalkResult:
    """Plain container: atoms in acceptance order, exactly what polymer.dat would list."""

    def __init__(self):
        self.xyz = self.chainID = self.type = self.bonds = self.chainLengths = self.countAfterRound = None
        self.acceptedTrial = None
        self.nAtoms = self.nBonds = self.target = self.rounds = 0
        self.initiated = self.lastProgressed = 0
        self.seconds = 0.0
        self.trials = self.seedTrials = 0


def _collect(lib, prefix, h, nChains, with_trials):
    r = WalkResult()
    nA, nB, tg = C.c_int64(), C.c_int64(), C.c_int64()
    ini, rounds, lp = C.c_int32(), C.c_int32(), C.c_int32()
    sec = C.c_double()
    tr, st = C.c_uint64(), C.c_uint64()
    getattr(lib, prefix + "_walk_counts")(h, C.byref(nA), C.byref(nB), C.byref(tg), C.byref(ini), C.byref(rounds),
                                          C.byref(lp), C.byref(sec), C.byref(tr), C.byref(st))
    r.nAtoms, r.nBonds, r.target = nA.value, nB.value, tg.value
    r.initiated, r.rounds, r.lastProgressed = ini.value, rounds.value, lp.value
    r.seconds, r.trials, r.seedTrials = sec.value, tr.value, st.value
    r.xyz = np.zeros((r.nAtoms, 3), np.float64)
    r.chainID = np.zeros(r.nAtoms, np.int32)
    r.type = np.zeros(r.nAtoms, np.int32)
    r.bonds = np.zeros((r.nBonds, 2), np.int64)
    r.chainLengths = np.zeros(nChains, np.int32)
    r.countAfterRound = np.zeros(r.rounds, np.int64)
    args = [h, _dptr(r.xyz), r.chainID.ctypes.data_as(C.c_void_p), r.type.ctypes.data_as(C.c_void_p),
            r.bonds.ctypes.data_as(C.c_void_p), r.chainLengths.ctypes.data_as(C.c_void_p),
            r.countAfterRound.ctypes.data_as(C.c_void_p)]
    if with_trials:
        r.acceptedTrial = np.zeros(r.nAtoms, np.int32)
        args.append(r.acceptedTrial.ctypes.data_as(C.c_void_p))
    getattr(lib, prefix + "_walk_get")(*args)
    return r


class OracleLib:
    def __init__(self, path=ORACLE_SO):
        self.lib = lib = C.CDLL(path)
        lib.oracle_lookup_create.restype = C.c_void_p
        lib.oracle_lookup_create.argtypes = [C.c_double] * 4
        lib.oracle_lookup_destroy.argtypes = [C.c_void_p]
        lib.oracle_lookup_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        lib.oracle_lookup_find.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        lib.oracle_lookup_find_bruteforce.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        lib.oracle_target_count.restype = C.c_int64
        lib.oracle_target_count.argtypes = [C.POINTER(C.c_double), C.c_double]
        lib.oracle_walk_run.restype = C.c_void_p
        lib.oracle_walk_run.argtypes = [C.POINTER(OracleParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        lib.oracle_walk_counts.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        lib.oracle_walk_get.argtypes = [C.c_void_p] * 8
        lib.oracle_walk_free.argtypes = [C.c_void_p]
        lib.oracle_sincos_portable.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.oracle_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.oracle_cone_pos.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_int, C.POINTER(C.c_double)]
        lib.oracle_rdf_edges.argtypes = [C.c_double, C.c_double, C.c_int32, C.c_void_p]
        lib.oracle_rdf.argtypes = [C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_chain_shapes.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]

    def rdf_edges(self, rMin, dr, nBins):
        e = np.zeros(nBins + 1)
        self.lib.oracle_rdf_edges(rMin, dr, nBins, e.ctypes.data_as(C.c_void_p))
        return e

    def rdf(self, box, xyz, chain, rMin, dr, nBins):
        """scripts/rdf.py:13-20,39-57 as integer pair counts: (intra, inter)"""
        xyz, _ = _xyz(xyz); chain = np.ascontiguousarray(chain, np.int32)
        intra = np.zeros(nBins, np.uint64); inter = np.zeros(nBins, np.uint64)
        self.lib.oracle_rdf((C.c_double * 3)(*box), xyz.ctypes.data_as(C.c_void_p), chain.ctypes.data_as(C.c_void_p), len(xyz), rMin, dr, nBins,
                            intra.ctypes.data_as(C.c_void_p), inter.ctypes.data_as(C.c_void_p))
        return intra, inter

    def chain_shapes(self, xyz, chain, nChains):
        xyz, _ = _xyz(xyz); chain = np.ascontiguousarray(chain, np.int32)
        ree2 = np.zeros(nChains); rg2 = np.zeros(nChains)
        self.lib.oracle_chain_shapes(xyz.ctypes.data_as(C.c_void_p), chain.ctypes.data_as(C.c_void_p), len(xyz), nChains,
                                     ree2.ctypes.data_as(C.c_void_p), rg2.ctypes.data_as(C.c_void_p))
        return ree2, rg2

    def target_count(self, box, density):
        b = (C.c_double * 3)(*box)
        return self.lib.oracle_target_count(b, density)

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        self.lib.oracle_philox(c, k, o)
        return list(o)

    def sincos_portable(self, x):
        s, c = C.c_double(), C.c_double()
        self.lib.oracle_sincos_portable(x, C.byref(s), C.byref(c))
        return s.value, c.value

    def cone_pos(self, last, pen, u01, trig=1):
        l = (C.c_double * 3)(*last); p = (C.c_double * 3)(*pen); o = (C.c_double * 3)()
        self.lib.oracle_cone_pos(l, p, u01, trig, o)
        return np.array(list(o))

    def lookup(self, box, thresh):
        return _Lookup(self.lib, "oracle", box, thresh)

    def walk(self, nChains, box, minDist=2.0, targetMassDensity=0.92, maxTrials=100, boundary="ppp", nGraftChains=0,
             graftLoc="file", growthOrder="roundRobin", growthBias=(0, 0, 0), trig="portable", seed=1, maxRounds=-1,
             grafts=None, obstacles=None):
        p = OracleParams()
        p.nChains = nChains; p.box = (C.c_double * 3)(*box); p.minDist = minDist; p.targetMassDensity = targetMassDensity
        p.nGraftChains = nGraftChains; p.maxTrials = maxTrials; p.graftLoc = GRAFT.get(graftLoc, 3)
        p.growthOrder = ORDER.get(growthOrder, 0); p.boundary = BOUNDARY.get(boundary, 0)
        p.growthBias = (C.c_double * 3)(*growthBias); p.trig = 1 if trig == "portable" else 0
        p.seed = seed; p.maxRounds = maxRounds
        g, ng = _xyz(grafts); o, no = _xyz(obstacles)
        h = self.lib.oracle_walk_run(C.byref(p), g.ctypes.data if g is not None else None, ng,
                                     o.ctypes.data if o is not None else None, no)
        try:
            return _collect(self.lib, "oracle", h, nChains, True)
        finally:
            self.lib.oracle_walk_free(h)


class _Lookup:
    def __init__(self, lib, prefix, box, thresh):
        self.lib, self.prefix = lib, prefix
        self.h = getattr(lib, prefix + "_lookup_create")(box[0], box[1], box[2], thresh)

    def add(self, xyz):
        a, n = _xyz(xyz)
        getattr(self.lib, self.prefix + "_lookup_add")(self.h, a.ctypes.data, n)

    def find(self, xyz, ignore=None, bruteforce=False):
        a, n = _xyz(xyz)
        ig = None if ignore is None else np.ascontiguousarray(ignore, dtype=np.int64)
        out = np.zeros(n, np.uint8)
        name = self.prefix + ("_lookup_find_bruteforce" if bruteforce else "_lookup_find")
        getattr(self.lib, name)(self.h, a.ctypes.data, ig.ctypes.data if ig is not None else None, n, out.ctypes.data)
        return out

    def close(self):
        if self.h:
            getattr(self.lib, self.prefix + "_lookup_destroy")(self.h)
            self.h = None

    def __del__(self):
        self.close()


class RefLib:
    """The unmodified reference (oracle/_ref). ``available()`` is False when it has not been built."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, path=REF_SO):
        self.lib = lib = C.CDLL(path)
        lib.ref_lookup_create.restype = C.c_void_p
        lib.ref_lookup_create.argtypes = [C.c_double] * 4
        lib.ref_lookup_destroy.argtypes = [C.c_void_p]
        lib.ref_lookup_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        lib.ref_lookup_find.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        lib.ref_walk_run.restype = C.c_void_p
        lib.ref_walk_run.argtypes = [C.POINTER(RefParams), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        lib.ref_walk_counts.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        lib.ref_walk_get.argtypes = [C.c_void_p] * 7
        lib.ref_walk_export.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        lib.ref_walk_export.restype = C.c_int
        lib.ref_walk_topology.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_walk_get_topology.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_walk_free.argtypes = [C.c_void_p]

    def lookup(self, box, thresh):
        return _Lookup(self.lib, "ref", box, thresh)

    def walk(self, nChains, box, minDist=2.0, targetMassDensity=0.92, maxTrials=100, boundary="ppp", nGraftChains=0,
             graftLoc="file", growthOrder="roundRobin", polymer="Polyethylene", rng="keyed", seed=1, maxRounds=-1,
             grafts=None, obstacles=None, export_dir=None, topology=False):
        p = RefParams()
        p.nChains = nChains; p.box = (C.c_double * 3)(*box); p.minDist = minDist; p.targetMassDensity = targetMassDensity
        p.nGraftChains = nGraftChains; p.maxTrials = maxTrials
        p.graftLoc = graftLoc.encode(); p.growthOrder = growthOrder.encode(); p.boundary = boundary.encode()
        p.polymer = polymer.encode(); p.rngMode = 1 if rng == "keyed" else 0; p.seed = seed; p.maxRounds = maxRounds
        g, ng = _xyz(grafts); o, no = _xyz(obstacles)
        h = self.lib.ref_walk_run(C.byref(p), g.ctypes.data if g is not None else None, ng,
                                  o.ctypes.data if o is not None else None, no)
        try:
            r = _collect(self.lib, "ref", h, nChains, False)
            if topology:
                na, nd = C.c_int64(), C.c_int64()
                self.lib.ref_walk_topology(h, C.byref(na), C.byref(nd))
                r.angles = np.zeros((na.value, 3), np.int64); r.dihedrals = np.zeros((nd.value, 4), np.int64)
                self.lib.ref_walk_get_topology(h, r.angles.ctypes.data, r.dihedrals.ctypes.data)
            if export_dir is not None:
                rc = self.lib.ref_walk_export(h, export_dir.encode(), b"ref.log")
                assert rc == 0, rc
            return r
        finally:
            self.lib.ref_walk_free(h)
