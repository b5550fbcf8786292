"""Pin the C oracle against LIVE calls into the real reference (build container only: skipped where
/root/reference does not exist).  Adversarial inputs the fixtures do not contain."""
import os

import numpy as np
import pytest

from oracle import coracle as co
from oracle.run_reference import reference_available
from tests.helpers import GOLDEN

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")]


@pytest.fixture(scope="module")
def ref():
    from oracle.run_reference import import_reference
    return import_reference()


def _peakachu(ref, G, res, exp, w):
    rc = ref[0]
    core = rc.Peakachu(G, upper=400 * res, res=res)
    core.load_expected({i: v for i, v in enumerate(exp)})
    core.w = w
    return core


def _random_band_matrix(rng, n, density, scale=1.0):
    d = np.abs(np.subtract.outer(np.arange(n), np.arange(n)))
    G = rng.poisson(30.0 / (d + 1.0) ** 0.8 + 0.3).astype(float) * scale * rng.uniform(0.005, 0.02, size=(n, n))
    G[rng.random((n, n)) > density] = 0.0
    return np.triu(G)


@pytest.mark.parametrize("seed,n,density,w,nexp", [(0, 90, 0.9, 6, 501), (1, 120, 0.35, 6, 501), (2, 70, 0.6, 5, 201),
                                                   (3, 260, 0.8, 5, 201), (4, 45, 1.0, 6, 1001)])
def test_candidates_and_windows(ref, seed, n, density, w, nexp):
    rng = np.random.default_rng(seed)
    G = _random_band_matrix(rng, n, density)
    G[5, 30] = np.nan                     # non-finite entries are dropped by Peakachu.__init__
    G[7, 40] = np.inf
    exp = 0.4 / (np.arange(nexp) + 1.0)
    core = _peakachu(ref, G, 10000, exp, w)
    r, c, e = co.candidates(G)
    assert np.array_equal(r, core.ridx) and np.array_equal(c, core.cidx)
    fea, cl = co.getwindow(G, r, c, w, exp)
    rf, rcl = core.getwindow(list(zip(core.ridx, core.cidx)), w)
    if len(rf) == 0:
        assert len(fea) == 0
        return
    assert np.array_equal(cl, rcl)
    assert np.array_equal(fea, rf, equal_nan=True)


def test_constant_and_zero_expected_windows(ref):
    """max == min -> NaN features; a zero in the expected table -> inf/NaN through the gaussian."""
    n = 80
    G = np.triu(np.full((n, n), 0.5))
    G[20:26, 50:58] = 3.0
    exp = np.ones(501)
    exp[37] = 0.0
    core = _peakachu(ref, G, 10000, exp, 6)
    r, c, e = co.candidates(G)
    with np.errstate(all="ignore"):
        rf, rcl = core.getwindow(list(zip(core.ridx, core.cidx)), 6)
    fea, cl = co.getwindow(G, r, c, 6, exp)
    assert np.array_equal(cl, rcl)
    assert np.array_equal(fea, rf, equal_nan=True)
    assert np.isnan(fea).all(axis=1).any() and np.isnan(fea).any(axis=1).sum() > np.isnan(fea).all(axis=1).sum() - 1


def test_forest_with_nan_rows(ref, models):
    rng = np.random.default_rng(5)
    for name, F in (("forest_w6.joblib", 169), ("forest_w5.joblib", 121)):
        model = models(name)
        X = rng.random((3000, F))
        X[rng.random(X.shape) < 0.02] = np.nan
        X[:50] = np.nan
        want = model.predict_proba(X)[:, 1]
        got = co.forest_predict(co.sklearn_forest_arrays(model), X.astype(np.float32))
        assert np.array_equal(want, got)


def test_gap_removal_and_block_means(ref):
    rc, ra, ru, cli = ref
    rng = np.random.default_rng(6)
    n = 150
    M = _random_band_matrix(rng, n, 0.7)
    M[:, [10, 11, 77]] = 0
    M[[10, 11, 77], :] = 0
    bias = rng.uniform(0.05, 0.15, n)
    bias[[10, 11, 77, 120]] = 0
    fu = object.__new__(rc.Fusion)
    fu.fusion_matrix = M
    fu.res = 10000
    fu.hcm = M
    fu.remove_gaps()
    assert np.array_equal(co.remove_gaps_index(M), np.array([fu.index_map[i] for i in range(len(fu.index_map))]))
    valid = np.triu(bias[:, None] * bias > 0)
    for (lo1, hi1, lo2, hi2) in ((0, 60, 0, 60), (0, 60, 60, 150), (60, 150, 60, 150)):
        mask = np.zeros((n, n), dtype=bool)
        mask[lo1:hi1, lo2:hi2] = True
        want = fu._extract_xy(mask & valid, maxdis=2000000, mink=3)
        got = co.block_diag_means(M, bias, (lo1, hi1), (lo2, hi2), 200)
        assert sorted(got) == sorted(want)
        assert all(got[k] == want[k] for k in want)


def test_tiny_matrices(ref):
    rc = ref[0]
    for n in (9, 10, 11, 14):
        G = np.triu(np.random.default_rng(n).random((n, n)))
        if n - 4 <= 5:
            with pytest.raises(ValueError):
                rc.Peakachu(G, upper=4000000, res=10000)
            with pytest.raises(ValueError):
                co.candidates(G)
        else:
            core = rc.Peakachu(G, upper=4000000, res=10000)
            r, c, _ = co.candidates(G)
            assert np.array_equal(r, core.ridx) and np.array_equal(c, core.cidx)
