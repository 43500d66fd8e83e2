"""GPU parity of the overlap query (K5): sarw_lookup_find_batch through the C ABI must return, bit for bit, what
PeriodicLookup::find returns on identical candidates (north_star correctness part 1)."""
import os
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_find_matches_golden_reference_truth_tables(sarw):
    """Truth tables written by the unmodified reference (random + adversarial candidates within a few ulp of the
    threshold sphere, unwrapped images, ignoreIndex cases, S=2 boxes)."""
    gold = np.load(os.path.join(GOLDEN, "find_truth.npz"))
    for ti in range(int(gold["n"])):
        lk = sarw.PeriodicLookup(tuple(gold[f"box{ti}"]), float(gold[f"thresh{ti}"]))
        lk.addPoint(gold[f"pts{ti}"])
        got = lk.find(gold[f"q{ti}"], gold[f"ig{ti}"])
        want = gold[f"hit{ti}"]
        assert np.array_equal(got, want), (ti, int(np.sum(got != want)))
        lk.close()


def test_find_matches_reference_on_sparse_adversarial_tables(sarw):
    """1e6 candidates ON the threshold sphere of isolated points (0/1/4/64 ulp either side, axis-aligned, unwrapped images, S = 2 and
    anisotropic boxes, ignoreIndex): answers of the unmodified reference (tests/golden/make_golden_find.py), about half of them hits."""
    from test_oracle import _sparse_tables
    total = 0
    for ti, box, thresh, pts, q, ig, want in _sparse_tables():
        lk = sarw.PeriodicLookup(box, thresh)
        lk.addPoint(pts)
        got = lk.find(q, ig); lk.close()
        assert np.array_equal(got, want), (ti, int(np.sum(got != want)))
        assert 0.3 < want.mean() < 0.7
        total += len(q)
    assert total >= 1_000_000


@pytest.mark.parametrize("cap", [0, 1, 2])
def test_find_matches_oracle_random_and_overflowing_cells(sarw, oracle, cap):
    """Dense random points (cells overflow into the overflow list when cellCapacity is forced small)."""
    rng = np.random.default_rng(42 + cap)
    box, thresh = (15.0, 17.0, 13.0), 1.0
    pts = rng.random((900, 3)) * np.array(box)
    q = rng.random((20000, 3)) * np.array(box) * 2 - np.array(box) / 2
    ig = np.where(rng.random(len(q)) < 0.3, rng.integers(0, len(pts), len(q)), -1).astype(np.int64)
    olk = oracle.lookup(box, thresh); olk.add(pts)
    want = olk.find(q, ig); olk.close()
    lk = sarw.PeriodicLookup(box, thresh, cellCapacity=cap)
    lk.addPoint(pts[:400]); lk.addPoint(pts[400:])   # two batches: indices continue
    got = lk.find(q, ig)
    assert lk.ctx.lookup_size() == len(pts)
    if cap:
        assert lk.ctx.stats().overflowBeads > 0
    lk.close()
    assert np.array_equal(got, want), int(np.sum(got != want))
    assert 0.05 < want.mean() < 0.95


def test_find_large_melt_like_density(sarw, oracle):
    """BASELINE config-2-sized lookup (1e5 points at melt density in a 136.2 A box), 2e5 candidates."""
    rng = np.random.default_rng(7)
    box, thresh = (136.2, 136.2, 136.2), 1.0
    pts = rng.random((100000, 3)) * np.array(box)
    q = np.concatenate([rng.random((100000, 3)) * np.array(box),
                        pts[rng.integers(0, len(pts), 100000)] + rng.normal(size=(100000, 3)) * 0.6])
    olk = oracle.lookup(box, thresh); olk.add(pts)
    want = olk.find(q); olk.close()
    lk = sarw.PeriodicLookup(box, thresh); lk.addPoint(pts)
    got = lk.find(q); lk.close()
    assert np.array_equal(got, want)


def test_find_layer_ordered_large_batch(sarw, oracle):
    """Batches of >= 2^18 candidates on a table larger than L2 are processed in cell-layer order (counting sort on the device);
    results must not depend on that: 4e5 candidates incl. unwrapped images and ignore indices, against the oracle."""
    rng = np.random.default_rng(11)
    box, thresh = (190.0, 186.0, 200.0), 1.0
    pts = rng.random((270000, 3)) * np.array(box)
    n = 400000
    q = np.concatenate([rng.random((n // 2, 3)) * np.array(box) * 3 - np.array(box),
                        pts[rng.integers(0, len(pts), n // 2)] + rng.normal(size=(n // 2, 3)) * 0.6])
    ig = np.where(rng.random(n) < 0.3, rng.integers(0, len(pts), n), -1).astype(np.int64)
    olk = oracle.lookup(box, thresh); olk.add(pts)
    want = olk.find(q, ig); olk.close()
    lk = sarw.PeriodicLookup(box, thresh); lk.addPoint(pts)
    got = lk.find(q, ig)
    got_small = np.concatenate([lk.find(q[k:k + 100000], ig[k:k + 100000]) for k in range(0, n, 100000)])   # unordered path
    lk.close()
    assert np.array_equal(got, want), int(np.sum(got != want))
    assert np.array_equal(got_small, want)
    assert 0.05 < want.mean() < 0.95


def test_find_empty_and_edge_inputs(sarw):
    lk = sarw.PeriodicLookup((10.0, 10.0, 10.0), 1.0)
    assert lk.find(np.zeros((0, 3))).shape == (0,)
    assert not lk.find(np.array([[1.0, 2.0, 3.0]]))[0]                 # empty table
    lk.addPoint(np.array([[0.0, 0.0, 0.0]]))
    q = np.array([[10.0, 10.0, 10.0], [9.5, 0.0, 0.0], [-0.5, 10.2, 20.1], [1.0, 0.0, 0.0], [1.0000000001, 0.0, 0.0], [5.0, 5.0, 5.0]])
    assert list(lk.find(q)) == [1, 1, 1, 1, 0, 0]                       # periodic images; dist <= thresh collides
    assert list(lk.find(q, np.zeros(len(q), np.int64))) == [0] * 6      # the only point ignored
    lk.close()
