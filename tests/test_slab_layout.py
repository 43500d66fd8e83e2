"""The slab decomposition of one melt over several GPUs as pure functions of (cell layers, ranks, halo): no GPU needed.
sarw_slab_layers / sarw_slab_owner are what sarw_create and the growth kernel compute."""
import itertools

import pytest


@pytest.fixture(scope="module")
def sarw():
    import sarw_b200
    sarw_b200.load_library()
    return sarw_b200


@pytest.mark.parametrize("S,R", [(6, 2), (29, 2), (29, 3), (29, 4), (149, 4), (315, 2), (680, 4), (1467, 8), (8, 8), (1000, 7)])
def test_slabs_tile_the_layers_and_owner_agrees(sarw, S, R):
    edges = [sarw.slab_layers(S, R, r, 0)[:2] for r in range(R)]
    assert edges[0][0] == 0 and edges[-1][1] == S
    for (lo, hi), (lo2, hi2) in zip(edges, edges[1:]):
        assert hi == lo2                                    # contiguous, no gap, no overlap
    sizes = [hi - lo for lo, hi in edges]
    assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1  # balanced to one layer
    owner = [sarw.slab_owner(S, R, z) for z in range(S)]
    for r, (lo, hi) in enumerate(edges):
        assert owner[lo:hi] == [r] * (hi - lo)


@pytest.mark.parametrize("S,R,halo", [(29, 2, 4), (29, 4, 4), (149, 4, 8), (315, 2, 8), (680, 4, 8), (1467, 8, 8), (1467, 8, 16), (6, 2, 4), (40, 3, 20)])
def test_window_is_the_slab_plus_the_halo_on_either_side(sarw, S, R, halo):
    for r in range(R):
        lo, hi, wlo, wlen = sarw.slab_layers(S, R, r, halo)
        window = {(wlo + k) % S for k in range(wlen)}
        need = {(z % S) for z in range(lo - halo, hi + halo)}
        if hi - lo + 2 * halo >= S:
            assert (wlo, wlen) == (0, S)                    # the window would cover the table anyway: the whole table
        else:
            assert wlen == hi - lo + 2 * halo and window == need and 0 <= wlo < S
        # whatever a round can touch from a tip inside the slab (reach <= halo - 2 layers) lies inside the window
        for tip in (lo, hi - 1):
            for d in range(-(halo - 2), halo - 1):
                assert (tip + d) % S in window


def test_c5_memory_per_rank(sarw):
    """BASELINE configs[4] on 8 GPUs: 1467 layers of 1467 x 1467 32-byte cells; a rank holds 183 or 184 layers + 2 x 8"""
    S, cell = 1467, 32
    per_rank = [sarw.slab_layers(S, 8, r, 8)[3] * S * S * cell for r in range(8)]
    assert max(per_rank) < 14.0e9 and S ** 3 * cell > 100e9
    assert sum(sarw.slab_layers(S, 8, r, 0)[1] - sarw.slab_layers(S, 8, r, 0)[0] for r in range(8)) == S


def test_bad_arguments(sarw):
    for args in itertools.product((0, 10), (0, 2), (-1, 0, 2), (-1, 0)):
        S, R, r, h = args
        ok = S >= 1 and R >= 1 and 0 <= r < R and h >= 0
        if ok:
            sarw.slab_layers(S, R, r, h)
        else:
            with pytest.raises(sarw.SarwError):
                sarw.slab_layers(S, R, r, h)
    assert sarw.slab_owner(10, 2, 10) == -1 and sarw.slab_owner(10, 2, -1) == -1
