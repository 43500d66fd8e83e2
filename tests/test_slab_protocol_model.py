"""A CPU model of the slab-mode round against the serial loop of sarw.cpp:251-281. TEST INFRASTRUCTURE, no product code.

The GPU kernels settle same-round conflicts of a decomposed box by a fixed-point iteration: all open conflicts are re-resolved at once by
their owners, each seeing for every other chain that is being re-resolved its old bead, its new bead or neither (whatever the
interleaving of warps and ranks happens to be), results are exchanged, closed entries of higher chains near a re-resolved chain are re-opened, chains that now collide are added.
The claim: whatever the interleaving, the iteration ends in the state the serial loop produces. Here the round is modelled in numpy
with the oracle's own Philox stream and cone arithmetic, the interleaving is drawn at random, and every round of a crowded melt is
compared with the serial round; the serial model itself is pinned to the oracle's walk.
"""
import numpy as np
import pytest

import cases

BOND = 1.54


class Melt:
    def __init__(self, O, cfg, seed):
        self.O, self.seed = O, seed
        self.n, self.T = cfg["nChains"], cfg["maxTrials"]
        self.box = np.array(cfg["boxSize"], dtype=np.float64)
        self.thresh = float(cfg["minDist"])
        self.Linv = 1.0 / self.box
        kw = cases.oracle_kwargs(cfg)
        w0 = O.walk(seed=seed, trig="portable", maxRounds=0, **kw)             # the oracle's own seeding: two beads per chain
        assert w0.nAtoms == 2 * self.n
        self.kw = kw
        self.beads = [np.array(p) for p in w0.xyz]
        self.last = [2 * i + 1 for i in range(self.n)]
        self.pen = [2 * i for i in range(self.n)]
        self.target = O.target_count(tuple(self.box), cfg.get("targetMassDensity", 0.92))
        edge = max(2.002 * self.thresh, 2.0)
        a = max(L / np.floor(L / edge) for L in self.box)
        self.R2 = (2 * BOND + self.thresh + 4 * a) ** 2                        # independence radius of the clean-up (sarw_kernels.cu)

    def key(self):
        return [self.seed & 0xFFFFFFFF, self.seed >> 32]

    def u_grow(self, chain, rnd, trial):                                       # oracle/philox.h oracle_u_grow
        w = self.O.philox([chain, rnd, (trial - 1) >> 2, 0], self.key())
        return w[(trial - 1) & 3] * (1.0 / 4294967296.0)

    def hits(self, cand, pts):
        """PeriodicLookup.h:53-55,67-71 for one candidate against an (m, 3) array of stored points"""
        if len(pts) == 0:
            return np.zeros(0, bool)
        P = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        v = cand * self.Linv; v = v - np.floor(v)
        vp = P * self.Linv; vp = vp - np.floor(vp)
        dv = v - vp
        dv = dv - np.floor(0.5 + dv)
        c = self.box * dv
        return np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2]) <= self.thresh

    def first_passing(self, chain, rnd, t0, old, extra):
        """lowest trial >= t0 whose candidate is free of the old beads (all but the chain's own last bead) and of `extra`"""
        last, pen = self.beads[self.last[chain]], self.beads[self.pen[chain]]
        for t in range(t0, self.T + 1):
            cand = self.O.cone_pos(last, pen, self.u_grow(chain, rnd, t), trig=1)
            if self.hits(cand, old[self.keep[chain]]).any():
                continue
            if len(extra) and self.hits(cand, extra).any():
                continue
            return t, cand
        return self.T + 1, None

    def start_round(self):
        old = np.array(self.beads)
        self.keep = []
        for i in range(self.n):
            m = np.ones(len(old), bool); m[self.last[i]] = False                # checkCollision ignores the bonded neighbour (sarw.cpp:198)
            self.keep.append(m)
        return old

    def serial_round(self, rnd):
        old = self.start_round()
        out, new = [], []
        for i in range(self.n):                                                 # chain i sees the beads chains < i added this round
            t, p = self.first_passing(i, rnd, 1, old, np.array(new).reshape(-1, 3))
            out.append((t, p))
            if p is not None: new.append(p)
        return out

    def tip_dist2(self, a, b):
        d = self.beads[self.last[a]] - self.beads[self.last[b]]
        d = d - self.box * np.rint(d * self.Linv)
        return float((d * d).sum())

    def slab_round(self, rnd, rng, stats):
        old = self.start_round()
        none = np.zeros((0, 3))
        tent = [self.first_passing(i, rnd, 1, old, none) for i in range(self.n)]    # propose: against the beads of earlier rounds only
        orig = [t for t, _ in tent]
        state = {}                                                                   # chain -> True (open) / False (closed)
        for i in range(1, self.n):                                                   # verify: lower chains' tentative beads
            if tent[i][1] is None: continue
            lower = [tent[j][1] for j in range(i) if tent[j][1] is not None]
            if len(lower) and self.hits(tent[i][1], np.array(lower)).any(): state[i] = True
        stats["conflicts"] += len(state)
        while any(state.values()):
            stats["iterations"] += 1
            before = list(tent)
            results, pushes = {}, set()
            todo = [x for x, o in state.items() if o]
            busy = set(todo)
            rng.shuffle(todo)                                                        # any order: they run at the same time
            for x in todo:
                def seen(j):                                                         # the other chain's bead as this resolver happens to see it:
                    if j not in busy: return before[j]                               # untouched in this pass
                    r = rng.random()                                                 # being re-resolved right now: its old bead, its new one,
                    if r < 0.35: return before[j]                                    # or neither (removed, not yet re-inserted)
                    if r < 0.7 and j in results: return results[j]
                    return (self.T + 1, None)
                lower = [q for q in (seen(j)[1] for j in range(x)) if q is not None]
                t, p = self.first_passing(x, rnd, orig[x], old, np.array(lower).reshape(-1, 3))
                results[x] = (t, p)
                if p is not None:
                    for k in range(x + 1, self.n):
                        q = seen(k)[1]
                        if q is not None and self.hits(q, p.reshape(1, 3)).any(): pushes.add(k)
            for x, r in results.items(): tent[x] = r                                 # exchange + apply
            for x in results: state[x] = False                                       # list update, the same on every rank
            for x in results:
                for y in state:
                    if y > x and not state[y] and self.tip_dist2(x, y) <= self.R2: state[y] = True
            for k in pushes: state[k] = True
            assert stats["iterations"] < 10000
        return tent

    def commit(self, result):
        for i, (t, p) in enumerate(result):
            if p is None: continue
            self.beads.append(p)
            self.pen[i] = self.last[i]; self.last[i] = len(self.beads) - 1


@pytest.mark.parametrize("name,cfg,seed", [
    ("dense_tiny", cases.DENSE_TINY, 3), ("dense_tiny", cases.DENSE_TINY, 11),
    # box smaller than the independence radius: every re-resolution re-opens every closed higher chain, ~30 conflicts per round
    ("packed", dict(nChains=150, boxSize=(12.0, 12.0, 12.0), minDist=1.0, maxTrials=30, targetMassDensity=10.0), 7),
    # box much larger than the radius: ~10 conflicts per round, some of them near each other
    ("roomy", dict(nChains=600, boxSize=(40.0, 40.0, 40.0), minDist=1.0, maxTrials=50, targetMassDensity=1.3), 9)])
def test_fixed_point_clean_up_gives_the_serial_round(oracle, name, cfg, seed):
    m = Melt(oracle, cfg, seed)
    rng = np.random.default_rng(1234 + seed)
    stats = dict(conflicts=0, iterations=0)
    rnd = rounds_with_conflicts = 0
    while len(m.beads) < m.target and rnd < 40:
        before = stats["conflicts"]
        serial = m.serial_round(rnd)
        slab = m.slab_round(rnd, rng, stats)
        rounds_with_conflicts += stats["conflicts"] > before
        for i, ((ts, ps), (tp, pp)) in enumerate(zip(serial, slab)):
            assert ts == tp, (rnd, i, ts, tp)
            assert (ps is None and pp is None) or np.array_equal(ps, pp)
        m.commit(serial)
        if all(t >= m.T for t, _ in serial): break                                  # sarw.cpp:277,282: no chain progressed
        rnd += 1
    assert stats["conflicts"] >= 5, stats                                            # the clean-up was really exercised
    if name in ("packed", "roomy"):
        assert stats["iterations"] > rounds_with_conflicts, stats                    # ... including re-opened / newly hit chains (more than one pass)
    # the serial model is the oracle: same beads in the same (acceptance) order
    w = oracle.walk(seed=seed, trig="portable", maxRounds=rnd + 1, **m.kw)
    assert w.nAtoms == len(m.beads) and np.array_equal(w.xyz, np.array(m.beads))


def test_the_model_notices_a_missing_re_open_rule(oracle):
    """control: without re-opening the closed higher chains near a re-resolved chain, the iteration ends in a state that is NOT the serial
    round's (the packed case) - the model test above can fail"""
    cfg = dict(nChains=150, boxSize=(12.0, 12.0, 12.0), minDist=1.0, maxTrials=30, targetMassDensity=10.0)
    m = Melt(oracle, cfg, 7)
    m.R2 = -1.0                                                                      # nothing is ever "near": the rule never fires
    rng = np.random.default_rng(99)
    differs = False
    for rnd in range(4):
        serial = m.serial_round(rnd)
        slab = m.slab_round(rnd, rng, dict(conflicts=0, iterations=0))
        differs = differs or any(ts != tp for (ts, _), (tp, _) in zip(serial, slab))
        m.commit(serial)
    assert differs
