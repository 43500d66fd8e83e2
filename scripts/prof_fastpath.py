"""Fast-path step cycles of propose_chain, averaged over every chain-round (needs the -DSARW_PROF build:
SARW_B200_LIB=polymer-sarw_b200/lib/libsarw_prof.so python scripts/prof_fastpath.py C2)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sarw_b200, cases
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = getattr(cases, name)
lib = sarw_b200.load_library()
out = (C.c_ulonglong * 64)()
for seed in (1, 2):
    s = sarw_b200.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed, **{k: v for k, v in cfg.items() if k not in ("nChains", "boxSize", "minDist")})
    lib.sarw_debug_prof(out, 64, 1)
    s.run(); d = s.ctx.stats().as_dict()
    assert lib.sarw_debug_prof(out, 64, 1) == 0, "not a SARW_PROF build"
    s.close()
    v = list(out); n = max(1, v[15])
    print(f"{name} seed {seed}: growMs {d['growMs']:.3f}, {n} chain-rounds; cycles per chain-round: issue copies {v[0]/n:.0f}, candidates {v[1]/n:.0f}, wait copies {v[2]/n:.0f}, "
          f"compact {v[3]/n:.0f}, ovf-count load {v[5]/n:.0f}, test {v[4]/n:.0f} | band fallbacks {v[11]/n:.4f}, regionOvf {v[12]/n:.4f} | neighbours {v[8]/n:.1f}, cells {v[9]/n:.1f}, rows {v[10]/n:.1f}")
    print(f"   grid-kernel phase A per chain: load state + commit last round {v[48]/n:.0f}, frame + region plan {v[49]/n:.0f}, propose_chain {v[50]/n:.0f}, insert + stores {v[51]/n:.0f}")
    print("   propose time histogram (512-cycle bins, % of chain-rounds):", " ".join(f"{100*x/n:.1f}" for x in v[16:48]))
