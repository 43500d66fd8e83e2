"""Golden hashes of the large walks (BASELINE configs[2], configs[3] and the 1e5-chain case): one Philox-keyed walk of the plain-C
oracle each (portable trig = the bit-exact target of the CUDA path; the oracle itself is pinned to the unmodified reference in
tests/test_oracle.py), reduced to a SHA-256 of xyz / bonds / type plus the walk's invariants. Minutes of CPU once; the GPU tests
then compare a ~60 ms melt against the hash on every run.

    python tests/golden/make_golden_big.py [C3] [C4] [BIG]      ->  tests/golden/big_walk_hashes.json
"""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import cases
import oracle_bindings as ob

OUT = os.path.join(HERE, "big_walk_hashes.json")
which = sys.argv[1:] or ["C3", "C4", "BIG"]
O = ob.OracleLib()
res = json.load(open(OUT)) if os.path.exists(OUT) else {}
for name in which:
    cfg = getattr(cases, name)
    grafts = cases.brush_grafts() if name == "C4" else None
    seed = 1
    t0 = time.time()
    o = O.walk(seed=seed, trig="portable", grafts=grafts, **cases.oracle_kwargs(cfg))
    res[name] = dict(seed=seed, nAtoms=int(o.nAtoms), nBonds=int(o.nBonds), rounds=int(o.rounds), trials=int(o.trials), seedTrials=int(o.seedTrials),
                     initiated=int(o.initiated), lastProgressed=int(o.lastProgressed), sha256=cases.walk_digest(o.xyz, o.bonds, o.type),
                     chainLengths_sha256=__import__("hashlib").sha256(np.ascontiguousarray(o.chainLengths, dtype=np.int32).tobytes()).hexdigest(),
                     oracle_seconds=round(time.time() - t0, 1))
    print(name, res[name], flush=True)
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
