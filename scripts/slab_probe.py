"""Slab-decomposed melt on ONE GPU (R slab contexts side by side), compared with the single-context walk by checksum (not a bench line).
   python scripts/slab_probe.py nChains boxEdge R maxBlocks [maxRounds] [density] [halo]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "polymer-sarw_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, sarw_b200
n, L, R, mb = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
maxRounds = int(sys.argv[5]) if len(sys.argv) > 5 else -1
dens = float(sys.argv[6]) if len(sys.argv) > 6 else 0.92
halo = int(sys.argv[7]) if len(sys.argv) > 7 else 0
cfg = dict(nChains=n, boxSize=(L, L, L), minDist=1.0, maxTrials=1000, targetMassDensity=dens)
c = sarw_b200.Context(seed=5, launchMode=1, **cfg); c.initiate(); st = c.grow(0, maxRounds); ref = c.walk_checksum(); d1 = st.as_dict(); full = c.cell_table_bytes(); c.close()
print("single", ref, {k: d1[k] for k in ("nBeads", "rounds", "trials", "conflicts", "growMs")}, "cells", full, flush=True)
streams = [torch.cuda.Stream() for _ in range(R)]
ctxs = []
for r in range(R):
    kw = dict(slabRank=r, slabRanks=R, slabHalo=halo) if os.environ.get("SLAB", "1") != "0" else {}
    c = sarw_b200.Context(seed=5, launchMode=1, maxBlocks=mb, **kw, **cfg)
    c.set_stream(streams[r].cuda_stream); assert c.initiate(); ctxs.append(c)
print("slab cell bytes per rank", [c.cell_table_bytes() for c in ctxs], flush=True)
sarw_b200.connect_peers_local(ctxs); torch.cuda.synchronize()
import threading
out = [None] * R
def work(r):
    try: out[r] = ctxs[r].mr_grow_peer(maxRounds)
    except Exception as e: out[r] = e
th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
t0 = time.time()
for t in th: t.start()
for t in th: t.join()
print("wall %.3f s" % (time.time() - t0))
for r, c in enumerate(ctxs):
    if isinstance(out[r], Exception): print("rank", r, "ERROR", out[r]); continue
    d = out[r].as_dict()
    print("rank", r, "same" if c.walk_checksum() == ref else "DIFFERENT", {k: d[k] for k in ("nBeads", "rounds", "trials", "conflicts", "growMs")},
          "cyc/round A V C", [round(d[k] / max(1, d["rounds"])) for k in ("cycPropose", "cycVerify", "cycCleanup")], "sub", [round(x / max(1, d["rounds"])) for x in d["cycSub"]], flush=True)
