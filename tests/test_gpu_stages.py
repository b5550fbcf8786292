"""GPU parity, stage by stage, through the C ABI, against the golden fixtures produced by the
real reference and against the C oracle.  Bit-exact everywhere (fp64 / fp32 features / indices)."""
import os

import numpy as np
import pytest

from tests.conftest import CASES
from tests.helpers import GOLDEN, load_case, scale_plan_from_slopes, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from neoloopfinder_b200 import _lib
    assert _lib.device_count() > 0, "no CUDA device visible"
    return _lib.Context.get(0)


@pytest.fixture(scope="module")
def forests(gpu):
    from neoloopfinder_b200.forest import load_forest
    return lambda name: load_forest(os.path.join(GOLDEN, name), gpu)


@pytest.mark.parametrize("name", CASES)
def test_normalisation_stages(name, gpu):
    from neoloopfinder_b200.scoring import DeviceAssembly
    z = load_case(name)
    res = int(z["res"])
    M, bias, bi = z["fusion_matrix"], z["bias"], z["block_index"]
    nb = len(bi)
    asm = DeviceAssembly(M, bias, bi, gpu)
    maxdiag = 2000000 // res
    mean, cnt = asm.diag_means(maxdiag)
    p = 0
    for a in range(nb):
        for b in range(a, nb):
            idx = np.where(~np.isnan(mean[p]))[0]
            assert list(idx) == list(z[f"exps_idx_{a}_{b}"]), (a, b)
            assert np.array_equal(mean[p][idx], z[f"exps_val_{a}_{b}"]), (a, b)
            p += 1
    op, val = scale_plan_from_slopes(z["slopes"])
    asm.set_scales(op, val)
    assert sha(asm.hcm()) == str(z["hcm_sha"])
    assert np.array_equal(asm.kept_index(), z["index_map"])
    assert sha(asm.gap_free()) == str(z["gap_free_sha"])
    asm.close()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("source", ["device", "host"])
def test_scoring_stages(name, source, gpu, forests):
    from neoloopfinder_b200.scoring import DeviceAssembly, ScoreBatch, local_clustering_gpu
    z = load_case(name)
    res = int(z["res"])
    M, bias, bi, exp = z["fusion_matrix"], z["bias"], z["block_index"], z["expected"]
    op, val = scale_plan_from_slopes(z["slopes"])
    asm = DeviceAssembly(M, bias, bi, gpu)
    asm.set_scales(op, val)
    forest = forests(str(z["model"]))
    w = int(z["w"])
    assert forest.w == w
    batch = ScoreBatch(4, 400, gpu)
    if source == "device":
        job = batch.add_assembly(asm)
    else:
        job = batch.add_dense(asm.gap_free())
    # candidates are available before any model is involved (Peakachu.__init__)
    r, c = batch.candidates(job)
    assert np.array_equal(r, z["ridx"]) and np.array_equal(c, z["cidx"])
    batch.keep_features(True)
    thre = float(z["prob"])
    batch.score(forest, exp, thre)
    n_cand, n_scored, n_don = batch.counts()
    assert n_cand[job] == len(z["ridx"])
    assert n_scored[job] == len(z["clist"])
    i, j, p = batch.scored(job)
    assert np.array_equal(np.stack([i, j], 1), z["clist"])
    fea32 = batch.features32(job)
    assert sha(fea32) == str(z["fea32_sha"])
    assert np.array_equal(p, z["probas"])          # bit-exact probabilities
    di, dj, dp, dv = batch.donuts(job)
    o = np.lexsort((di, dj - di))                  # diagonal-major like the reference's clist order
    assert np.array_equal(np.stack([di, dj], 1)[o], z["donut_ij"])
    assert np.array_equal(dv[o], z["donut_val"])
    # float64 getwindow accessor == reference fea
    fea, cl = batch.getwindow(job, z["ridx"], z["cidx"], w, exp)
    assert np.array_equal(cl, z["clist"])
    assert sha(fea) == str(z["fea_sha"])
    # pooling
    chains, im = z["chains"], z["index_map"]
    ij = np.stack([di, dj], 1)
    loops = local_clustering_gpu(ij, dv, res, int(z["min_count"]), 15000, chains[im[ij[:, 0]]], chains[im[ij[:, 1]]], gpu)
    assert sorted(loops) == [tuple(x) for x in z["loop_index"]]
    pm = {(int(a), int(b)): float(q) for a, b, q in zip(di, dj, dp)}
    assert [pm[t] for t in sorted(loops)] == list(z["loop_prob"])
    batch.close()
    asm.close()


def test_pool_cases(gpu):
    from neoloopfinder_b200.scoring import local_clustering_gpu
    z = np.load(os.path.join(GOLDEN, "pool_cases.npz"))
    for k in range(int(z["n_cases"])):
        ij, val, ch = z[f"c{k}_ij"], z[f"c{k}_val"], z[f"c{k}_chain"]
        res, mc, r = (int(v) for v in z[f"c{k}_par"])
        got = local_clustering_gpu(ij, val, res, mc, r, ch[ij[:, 0]], ch[ij[:, 1]], gpu)
        assert sorted(got) == [tuple(x) for x in z[f"c{k}_out"]], f"pool case {k}"


def test_forest_rows_match_sklearn(gpu, forests, models):
    rng = np.random.default_rng(3)
    for name, F in (("forest_w6.joblib", 169), ("forest_w5.joblib", 121)):
        X = rng.random((4000, F)).astype(np.float32)
        X[rng.random(X.shape) < 0.01] = np.nan            # missing values follow missing_go_to_left
        model = models(name)
        ref = model.predict_proba(X)[:, 1]
        got = forests(name).predict_proba1(X)
        assert np.array_equal(ref, got)


def test_batch_of_all_cases_matches_oracle(gpu, forests, models):
    """All 10 kb cases in ONE batch (the production shape) against the C oracle."""
    from oracle import coracle as co
    from neoloopfinder_b200.scoring import ScoreBatch
    names = ["case_transloc_10k", "case_inversion_10k", "case_complex_10k"]
    forest = forests("forest_w6.joblib")
    batch = ScoreBatch(4, 400, gpu)
    mats = []
    for nme in names:
        z = load_case(nme)
        op, val = scale_plan_from_slopes(z["slopes"])
        hcm = co.scale_blocks(z["fusion_matrix"], z["block_index"], op, val)
        kept = co.remove_gaps_index(hcm)
        G = hcm[np.ix_(kept, kept)]
        mats.append((G, z))
        batch.add_dense(G)
    exp = mats[0][1]["expected"]
    batch.score(forest, exp, 0.5)
    arrays = co.sklearn_forest_arrays(models("forest_w6.joblib"))
    for job, (G, z) in enumerate(mats):
        i, j, p = batch.scored(job)
        cl, pr, _ = co.score_matrix(G, exp, arrays, 6)
        assert np.array_equal(np.stack([i, j], 1), cl)
        assert np.array_equal(p, pr)
    batch.close()


def test_constant_windows_give_nan_features_like_sklearn(gpu, forests, models):
    """A flat matrix: every window beyond the diagonal zone is constant, so max == min and the
    min-max normalisation yields NaN features (callers.py:606, util.py:493-497); sklearn >= 1.4
    routes NaN by missing_go_to_left.  Covers the NaN path of the forest kernels."""
    from oracle import coracle as co
    from neoloopfinder_b200.scoring import ScoreBatch
    n = 96
    G = np.triu(np.full((n, n), 0.75))
    G[40:44, 70:75] = 2.5                      # a few non-constant windows too
    exp = np.ones(501)
    forest = forests("forest_w6.joblib")
    model = models("forest_w6.joblib")
    batch = ScoreBatch(4, 400, gpu)
    batch.add_dense(G)
    batch.keep_features(True)
    batch.score(forest, exp, 0.5)
    i, j, p = batch.scored(0)
    fea32 = batch.features32(0)
    assert np.isnan(fea32).all(axis=1).sum() > 1000          # the constant windows
    assert (~np.isnan(fea32).any(axis=1)).sum() > 100          # and ordinary ones
    cl, pr, f32 = co.score_matrix(G, exp, co.sklearn_forest_arrays(model), 6)
    assert np.array_equal(np.stack([i, j], 1), cl)
    assert np.array_equal(fea32.view(np.uint32), f32.view(np.uint32))
    assert np.array_equal(p, pr)
    assert np.array_equal(p, model.predict_proba(fea32)[:, 1])
    batch.close()


@pytest.mark.parametrize("n", [9, 10, 40, 133, 300, 777, 2100])
def test_diagonal_means_and_isotonic_bit_exact(n, gpu):
    """K5: every diagonal mean equals numpy's (pairwise order) and the expected equals sklearn's
    IsotonicRegression(increasing=False, out_of_bounds='clip') prediction, bit for bit."""
    from sklearn.isotonic import IsotonicRegression
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.scoring import ScoreBatch
    rng = np.random.default_rng(n)
    G = np.triu(rng.random((n, n)) * np.exp(-np.abs(np.subtract.outer(np.arange(n), np.arange(n))) / 40.0))
    G[rng.random((n, n)) < 0.3] = 0.0
    batch = ScoreBatch(4, 400, gpu)
    batch.add_dense(G)
    if n - 4 <= 5:
        with pytest.raises(_lib.NLFError):
            batch.counts()
        batch.close()
        return
    means, e = batch.expected(0)
    idx, obs = [], []
    for d in range(4, 401):
        diag = np.diagonal(G, d) if d < n else np.zeros(0)
        if diag.size > 5:
            assert means[d] == diag.mean(), d
            idx.append(d); obs.append(diag.mean())
        else:
            assert np.isnan(means[d])
    ref = IsotonicRegression(increasing=False, out_of_bounds="clip").fit(np.r_[idx], np.r_[obs]).predict(np.arange(4, 401))
    assert np.array_equal(e[4:401], ref)
    batch.close()


def _random_band_matrix(rng, n, density):
    d = np.abs(np.subtract.outer(np.arange(n), np.arange(n)))
    G = rng.poisson(30.0 / (d + 1.0) ** 0.8 + 0.3).astype(float) * rng.uniform(0.005, 0.02, size=(n, n))
    G[rng.random((n, n)) > density] = 0.0
    return np.triu(G)


@pytest.mark.parametrize("seed,n,density,fname,nexp,zero_at", [
    (0, 90, 0.9, "forest_w6.joblib", 501, None), (1, 230, 0.25, "forest_w6.joblib", 501, None),
    (2, 70, 0.6, "forest_w5.joblib", 201, None), (3, 460, 0.8, "forest_w5.joblib", 201, None),
    (4, 150, 0.7, "forest_w6.joblib", 1001, 37), (5, 640, 0.1, "forest_w6.joblib", 501, None)])
def test_adversarial_matrices_match_oracle(seed, n, density, fname, nexp, zero_at, gpu, forests, models):
    """Sparse maps (most windows fail the gates: ragged feature rows), NaN/inf entries, the w=5
    model with a 201-entry expected (windows with d+2w >= 201 stay un-normalised, the boundary falls
    inside a 32-diagonal tile), a zero in the expected table (inf -> NaN features): all bit-exact
    against the C oracle, which tests/test_oracle_vs_reference.py pins to the live reference."""
    from oracle import coracle as co
    from neoloopfinder_b200.scoring import ScoreBatch
    rng = np.random.default_rng(seed)
    G = _random_band_matrix(rng, n, density)
    G[5, 30] = np.nan
    G[7, 40] = np.inf
    exp = 0.4 / (np.arange(nexp) + 1.0)
    if zero_at is not None:
        exp[zero_at] = 0.0
    forest = forests(fname)
    batch = ScoreBatch(4, 400, gpu)
    batch.add_dense(G)
    batch.keep_features(True)
    batch.score(forest, exp, 0.3)
    r, c = batch.candidates(0)
    orr, orc, _ = co.candidates(G)
    assert np.array_equal(r, orr) and np.array_equal(c, orc)
    with np.errstate(all="ignore"):
        cl, pr, f32 = co.score_matrix(G, exp, co.sklearn_forest_arrays(models(fname)), forest.w)
    if len(cl) < 2:
        assert batch.counts()[1][0] == len(cl)
        batch.close()
        return
    i, j, p = batch.scored(0)
    assert np.array_equal(np.stack([i, j], 1), cl)
    g32 = batch.features32(0)
    nan = np.isnan(f32)
    assert np.array_equal(np.isnan(g32), nan)                      # NaN payload bits are not part of the contract
    assert np.array_equal(g32.view(np.uint32)[~nan], f32.view(np.uint32)[~nan])
    assert np.array_equal(p, pr)
    fea64, cl64 = batch.getwindow(0, orr, orc, forest.w, exp)
    with np.errstate(all="ignore"):
        ofea, ocl = co.getwindow(G, orr, orc, forest.w, exp)
    assert np.array_equal(cl64, ocl) and np.array_equal(fea64, ofea, equal_nan=True)
    batch.close()


@pytest.mark.parametrize("n,density,trees,depth", [(330, 0.85, 40, 12), (520, 0.5, 120, None)])
def test_w7_model_matches_oracle(n, density, trees, depth, gpu):
    """A 15 x 15 (w=7) model -- the window Peakachu's constructor sizes its band for -- exercises the
    third instance of the feature kernel (8-row tiles, 225-feature groups) and, with 120 fully grown
    trees, a forest that needs several shared-memory parts.  The model is trained here on windows
    of the same matrix so probabilities spread over [0, 1]."""
    from sklearn.ensemble import RandomForestClassifier
    from oracle import coracle as co
    from neoloopfinder_b200.forest import load_forest
    from neoloopfinder_b200.scoring import ScoreBatch
    rng = np.random.default_rng(n)
    G = _random_band_matrix(rng, n, density)
    exp = 0.4 / (np.arange(501) + 1.0)
    r, c, _ = co.candidates(G)
    fea, cl = co.getwindow(G, r, c, 7, exp)
    assert fea.shape[1] == 225 and len(fea) > 2000
    pick = rng.choice(len(fea), 3000, replace=False)
    y = (fea[pick, 112] + 0.3 * rng.random(3000) > np.median(fea[pick, 112]) + 0.15).astype(int)
    model = RandomForestClassifier(n_estimators=trees, max_depth=depth, random_state=7, n_jobs=1).fit(
        fea[pick].astype(np.float32), y)
    forest = load_forest(model, gpu)
    assert forest.w == 7
    batch = ScoreBatch(4, 400, gpu)
    batch.add_dense(G)
    batch.keep_features(True)
    batch.score(forest, exp, 0.5)
    ocl, opr, of32 = co.score_matrix(G, exp, co.sklearn_forest_arrays(model), 7)
    i, j, p = batch.scored(0)
    assert np.array_equal(np.stack([i, j], 1), ocl)
    assert np.array_equal(batch.features32(0).view(np.uint32), of32.view(np.uint32))
    assert np.array_equal(p, opr)
    assert np.array_equal(p, model.predict_proba(of32)[:, 1])          # and sklearn itself
    assert 0.05 < (p >= 0.5).mean() < 0.95
    fea64, cl64 = batch.getwindow(0, r, c, 7, exp)
    assert np.array_equal(cl64, cl) and np.array_equal(fea64, fea)
    batch.close()


@pytest.mark.parametrize("seed,sizes", [(0, (700, 650)), (1, (300, 40, 520)), (2, (9, 1100))])
def test_pair_diagonal_means_long_blocks(seed, sizes, gpu):
    """K2 with more valid pixels per diagonal than one numpy leaf (128) and ragged blocks, against
    the port's extract_xy (the reference's masked ``M[x, y].mean()`` per diagonal)."""
    from neoloopfinder_b200.scoring import DeviceAssembly
    from oracle import pyport
    rng = np.random.default_rng(seed)
    n = sum(sizes)
    M = np.triu(rng.poisson(4.0, size=(n, n)) * rng.uniform(0.001, 0.02, size=(n, n)))
    M[rng.random((n, n)) < 0.3] = 0.0
    bias = rng.uniform(0.05, 0.15, n)
    bias[rng.random(n) < 0.08] = 0.0
    edges = np.r_[0, np.cumsum(sizes)]
    bi = [(int(edges[k]), int(edges[k + 1])) for k in range(len(sizes))]
    res, maxdiag = 10000, 200
    asm = DeviceAssembly(M, bias, bi, gpu)
    mean, cnt = asm.diag_means(maxdiag)
    valid_bias = np.triu((bias[:, None] * bias) > 0)
    p = 0
    for a, ia in enumerate(bi):
        for b, ib in enumerate(bi):
            if a > b:
                continue
            mask = np.zeros((n, n), dtype=bool)
            mask[ia[0]:ia[1], ib[0]:ib[1]] = True
            want = pyport.extract_xy(M, res, mask & valid_bias, maxdis=maxdiag * res, mink=3)
            got = {int(i): mean[p][i] for i in np.where(~np.isnan(mean[p]))[0]}
            assert got == want, (a, b)
            p += 1
    assert cnt.max() > 128
    asm.close()


def test_gap_columns_with_negative_and_cancelling_entries(gpu):
    """Fusion.remove_gaps tests ``hcm.sum(axis=0) == 0``: columns whose entries cancel are gaps, columns
    holding NaN are not; the parallel sign scan hands exactly those columns to the numpy-order sum."""
    from neoloopfinder_b200.scoring import DeviceAssembly
    rng = np.random.default_rng(5)
    n = 150
    M = np.triu(rng.random((n, n)))
    M[:, 10] = 0.0                                   # plain gap
    M[:, 20] = 0.0
    M[3, 20], M[7, 20] = 1.5, -1.5                   # cancels exactly: gap
    M[:, 30] = 0.0
    M[0, 30], M[1, 30], M[2, 30] = 1e16, 1.0, -1e16  # (1e16 + 1) - 1e16 = 0 in row order: gap
    M[:, 40] = 0.0
    M[0, 40], M[1, 40], M[2, 40] = 1e16, -1e16, 1.0  # same entries, another order: not a gap
    M[5, 50] = -M[5, 50]                             # mixed signs, non-zero sum
    M[:, 60] = 0.0
    M[2, 60] = np.nan                                # NaN sum != 0: kept
    M[:, 70] = 0.0
    M[4, 70] = -0.0
    bias = np.ones(n)
    bi = [(0, 80), (80, n)]
    op = np.array([[1, 2], [0, 1]], np.int32)
    val = np.array([[2.0, 0.5], [1.0, 4.0]])
    asm = DeviceAssembly(M, bias, bi, gpu)
    asm.set_scales(op, val)
    hcm = asm.hcm()
    with np.errstate(invalid="ignore"):
        want = np.where(~(hcm.sum(axis=0) == 0))[0]
    assert np.array_equal(asm.kept_index(), want)
    assert not {10, 20, 30, 70} & set(want.tolist()) and {40, 50, 60} <= set(want.tolist())
    asm.close()


@pytest.mark.parametrize("res", [5000, 10000, 25000])
def test_find_anchors_and_cluster_core_methods(res, gpu):
    """Peakachu.find_anchors / _cluster_core as callable methods against the port, which
    makes the reference's scipy / sklearn calls."""
    from neoloopfinder_b200.callers import Peakachu
    from oracle import pyport
    rng = np.random.default_rng(res)
    n = 120
    G = np.triu(rng.random((n, n)))
    mine = Peakachu(G, upper=100 * res, res=res)
    port = pyport.PeakachuPort(G, upper=100 * res, res=res)
    for trial in range(25):
        centres = rng.integers(20, 400, size=int(rng.integers(1, 8)))
        pos = np.concatenate([rng.integers(c - 6, c + 7, size=int(rng.integers(2, 30))) for c in centres])
        pos = np.concatenate([pos, rng.integers(0, 420, size=int(rng.integers(0, 12)))]).tolist()
        for mc in (2, 3):
            try:
                want = port.find_anchors(pos, min_count=mc, min_dis=15000)
            except Exception as exc:                                  # noqa: BLE001 - scipy refuses degenerate inputs
                pytest.skip("port raised %r" % (exc,))
            assert mine.find_anchors(pos, min_count=mc, min_dis=15000) == want, (trial, mc)
    for trial in range(25):
        m = int(rng.integers(1, 60))
        pts = set()
        while len(pts) < m:
            c = rng.integers(10, 100, size=2)
            pts.add((int(c[0] + rng.integers(-4, 5)), int(c[1] + rng.integers(-4, 5))))
        sort_list = sorted(((float(rng.integers(1, 9)), p) for p in pts), reverse=True)
        r = max(15000 // res, 2)
        v1, f1, v2, f2 = set(), [], set(), []
        port._cluster_core(list(sort_list), r, v2, f2)
        mine._cluster_core(list(sort_list), r, v1, f1)
        assert f1 == f2 and v1 == {(int(a), int(b)) for a, b in v2}, trial
    mine.close()


@pytest.mark.parametrize("w,nexp", [(6, 501), (5, 60), (7, 1001)])
def test_distance_normalize_on_window_stacks(w, nexp, gpu):
    """util.distance_normalize as a standalone call against the port: NaN cells, sparse windows,
    zero lower-left corners, weak centres and windows that reach beyond the expected table."""
    from neoloopfinder_b200.util import distance_normalize
    from oracle import pyport
    rng = np.random.default_rng(100 + w)
    S, n = 2 * w + 1, 400
    pool = rng.poisson(2.0, size=(n, S, S)).astype(float) * rng.uniform(0.01, 0.2, size=(n, S, S))
    pool[rng.random((n, S, S)) < 0.05] = np.nan
    pool[:60][rng.random((60, S, S)) < 0.93] = 0.0             # fail the 10 % non-zero gate
    pool[60:100, :w, :w] = 0.0                                 # lower-left mean 0
    pool[100:140, w, w] *= 1e-3                                # centre below 0.1 of the corner mean
    xi = rng.integers(w, 300, size=n).astype(np.int64)
    yi = xi + rng.integers(w + 2, 90, size=n)
    exp = 1.0 / (np.arange(nexp) + 1.0)
    a, b = pool.copy(), pool.copy()
    want_fea, want_cl = pyport.distance_normalize(a, exp, xi, yi, w)
    got_fea, got_cl = distance_normalize(b, exp, xi, yi, w)
    assert [tuple(map(int, c)) for c in got_cl] == [tuple(map(int, c)) for c in want_cl]
    assert 100 < len(want_cl) < n - 100
    assert all(np.array_equal(x, y) for x, y in zip(got_fea, want_fea))
    assert np.array_equal(a, b, equal_nan=True) and not np.isnan(b).any()      # NaN cleared in the caller's stack
    if nexp == 60:
        assert any((y - x) + 2 * w >= nexp for x, y in want_cl)               # some windows stay un-normalised
