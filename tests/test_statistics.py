"""Statistical parity (north_star correctness part 3): RDF (scripts/rdf.py math), end-to-end distance, radius of
gyration and chain-length distribution of melts grown with the counter-based Philox stream must match melts grown by the
UNMODIFIED reference with its native random numbers (mt19937_64 re-seeded per trial; fixture
tests/golden/stats_c2_reference_mt.npz: 12 reference melts of BASELINE configs[1]).

Stated tolerance: for every statistic, |mean_ours - mean_ref| <= 4 * sqrt(se_ours^2 + se_ref^2) + 1 % of the reference
mean, where se is the standard error over seeds (seed-to-seed spread of the RDF bins is 3-4 %, of <Ree^2> ~8 %)."""
import os
import numpy as np
import pytest

import cases
import stats
from conftest import GOLDEN


def compare(ours, ref, what):
    for key in ("gInter", "gIntra", "ree2", "rg2", "cn", "meanLen", "stdLen"):
        a, b = np.atleast_2d(ours[key].T).T, np.atleast_2d(ref[key].T).T
        ma, mb = a.mean(0), b.mean(0)
        se = np.sqrt(a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b))
        tol = 4 * se + 0.01 * np.abs(mb)
        assert np.all(np.abs(ma - mb) <= tol), (what, key, ma, mb, tol)


def test_oracle_philox_statistics_match_reference_native_rng(oracle):
    """CPU: the oracle walk driven by Philox (the stream the GPU uses) vs the reference with mt19937_64."""
    ref = np.load(os.path.join(GOLDEN, "stats_c2_reference_mt.npz"))
    walks = []
    for seed in range(1, 7):
        o = oracle.walk(seed=seed, trig="portable", **cases.oracle_kwargs(cases.C2))
        walks.append((o.xyz, o.chainID))
    compare(stats.summarize(walks, cases.C2["boxSize"], 100), ref, "oracle/philox")


@pytest.mark.gpu
def test_gpu_statistics_match_reference_native_rng(sarw):
    """GPU: 12 melts of BASELINE configs[1] through the C ABI vs 12 melts of the unmodified reference."""
    ref = np.load(os.path.join(GOLDEN, "stats_c2_reference_mt.npz"))
    cfg = cases.C2
    walks = []
    for seed in range(201, 213):
        s = sarw.SARW(cfg["nChains"], cfg["boxSize"], cfg["minDist"], seed=seed, maxTrials=cfg["maxTrials"])
        s.run()
        xyz, chainID, typ, bonds = s.ctx.download()
        s.close()
        walks.append((xyz, chainID))
    ours = stats.summarize(walks, cfg["boxSize"], 100)
    compare(ours, ref, "gpu")
