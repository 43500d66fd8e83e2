"""Diagnostics of the quad propose path (library built with -DSARW_QUAD_STATS): why chains leave it. python scripts/quad_stats.py [C3|BIG]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sarw_b200, cases
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
c = sarw_b200.Context(seed=4242, **getattr(cases, name))
c.initiate(); st = c.grow(0, -1)
out = (C.c_ulonglong * 8)()
c.lib.sarw_debug_paths.argtypes = [C.c_void_p, C.c_void_p]
c.lib.sarw_debug_paths(c.h, out)
cyc, cnt = list(out[:4]), list(out[4:])
tot = sum(cnt) + cyc[0]
print(f"{name}: rounds {st.rounds} beads {st.nBeads} growMs {st.growMs:.2f}; active chain-rounds {tot}: quad-accepted {cnt[0]} ({cnt[0]/tot:.3f}), hard {cnt[1]} ({cnt[1]/tot:.3f}), "
      f"non-tile {cnt[2]} ({cnt[2]/tot:.3f}), no-fit {cnt[3]-0} (incl. failed8) failed8 {cyc[0]} ({cyc[0]/tot:.3f}); quad steps {cyc[1]}, slow chains {cyc[2]} , steps with >= 3 slow {cyc[3]}")
busy = (C.c_ulonglong * 16)()
c.lib.sarw_debug_busy.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
c.lib.sarw_debug_busy(c.h, busy, 16)
steps = max(1, busy[8])
print(f"  neighbour lists split: count pass {busy[9]/steps:.0f}, scan + write pass {busy[10]/steps:.0f}, overflow-table append + sync {busy[4]/steps:.0f}")
names = ["open+commit", "frame+region", "tma+candidates", "wait copies", "neighbour lists", "test", "insert+stores", "slow path"]
print(f"  warp 0 of block 0: {steps} quad steps, cycles per step: " + ", ".join(f"{n} {busy[k]/steps:.0f}" for k, n in enumerate(names)) + f"; sum {(sum(busy[:8]) + busy[9] + busy[10])/steps:.0f}; cyc/round A {st.cycPropose/st.rounds:.0f} V {st.cycVerify/st.rounds:.0f} C {st.cycCleanup/st.rounds:.0f}")
c.close()
