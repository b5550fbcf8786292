"""Size-independent properties at the sizes bench.py measures (BASELINE.json configs[1]: hg38-sized 10 kb map,
3 Mb span, 602-bin assemblies): the public end-to-end call equals the resident batch, re-scoring and job order
do not change anything, thresholds nest, pooled loops are Donuts, and per-matrix counts equal the C oracle."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

N_ASM = 10


@pytest.fixture(scope="module")
def workload():
    import bench
    from neoloopfinder_b200 import _lib
    from neoloopfinder_b200.forest import load_forest
    wl = bench.Workload("cfg2", N_ASM)
    ctx = _lib.Context.get(0)
    clr, lines, expected = wl.build()
    forest = load_forest(os.path.join(GOLDEN, "forest_w6.joblib"), ctx)
    mats, chains = bench.gap_free_matrices(clr, list(lines), lines, expected, ctx, forest, wl.res)
    exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
    assert len(mats) >= 8 and min(m.shape[0] for m in mats) > 500
    return ctx, mats, chains, exp_arr, forest


def _as_sets(loops):
    return [sorted(zip(l[0].tolist(), l[1].tolist(), l[2].tolist())) for l in loops]


def test_end_to_end_call_equals_resident_batch_and_is_idempotent(workload):
    from neoloopfinder_b200.scoring import ScoreBatch, score_and_pool, score_matrices
    ctx, mats, chains, exp_arr, forest = workload
    b = ScoreBatch(4, 400, ctx)
    for G in mats:
        b.add_dense(G)
    loops, counts = score_and_pool(b, forest, exp_arr, 0.9, 10000, 2, chains, False)
    scored = [b.scored(k) for k in range(len(mats))]
    b.reset()                                                        # same matrices, scored again
    loops2, counts2 = score_and_pool(b, forest, exp_arr, 0.9, 10000, 2, chains, False)
    assert _as_sets(loops) == _as_sets(loops2) and all(np.array_equal(x, y) for x, y in zip(counts, counts2))
    for k in (0, len(mats) - 1):
        again = b.scored(k)
        assert all(np.array_equal(x, y) for x, y in zip(scored[k], again))
    b.close()
    for depth in (1, 2, 4):                                          # pipelined sub-batches of the public call
        loops3, counts3 = score_matrices(mats, forest, exp_arr, 0.9, 10000, 2, chains, False, 400, ctx, depth=depth)
        assert _as_sets(loops3) == _as_sets(loops), depth
        assert all(np.array_equal(x, y) for x, y in zip(counts3, counts)), depth
    assert sum(len(l[0]) for l in loops) > 10 and int(counts[1].sum()) > 500000


def test_job_order_does_not_matter(workload):
    from neoloopfinder_b200.scoring import score_matrices
    ctx, mats, chains, exp_arr, forest = workload
    fwd, cf = score_matrices(mats, forest, exp_arr, 0.9, 10000, 2, chains, False, 400, ctx)
    rev, cr = score_matrices(mats[::-1], forest, exp_arr, 0.9, 10000, 2, chains[::-1], False, 400, ctx)
    assert _as_sets(fwd) == _as_sets(rev)[::-1]
    assert all(np.array_equal(x, y[::-1]) for x, y in zip(cf, cr))


def test_thresholds_nest_and_pooled_loops_are_donuts(workload):
    from neoloopfinder_b200.scoring import ScoreBatch
    ctx, mats, _chains, exp_arr, forest = workload
    donuts = {}
    for thre in (0.5, 0.9):
        b = ScoreBatch(4, 400, ctx)
        for G in mats[:4]:
            b.add_dense(G)
        b.score(forest, exp_arr, thre)
        donuts[thre] = [b.donuts(k) for k in range(4)]
        if thre == 0.5:
            probs = []
            for k in range(4):
                si, sj, sp = b.scored(k)
                probs.append(dict(zip(zip(si.tolist(), sj.tolist()), sp.tolist())))
        b.close()
    for k in range(4):
        lo = set(zip(donuts[0.5][k][0].tolist(), donuts[0.5][k][1].tolist()))
        hi = set(zip(donuts[0.9][k][0].tolist(), donuts[0.9][k][1].tolist()))
        assert hi <= lo and len(lo) > len(hi)
        # the Donuts are exactly the scored pixels at or above the threshold, with their contact value
        assert lo == {ij for ij, p in probs[k].items() if p >= 0.5}
        assert hi == {ij for ij, p in probs[k].items() if p >= 0.9}
        G = mats[k]
        i, j, _p, v = donuts[0.5][k]
        assert np.array_equal(v, G[i, j])


def test_no_pool_returns_every_donut_and_pooling_a_subset(workload):
    from neoloopfinder_b200.scoring import score_matrices
    ctx, mats, chains, exp_arr, forest = workload
    pooled, _c = score_matrices(mats[:5], forest, exp_arr, 0.6, 10000, 2, chains[:5], False, 400, ctx)
    plain, counts = score_matrices(mats[:5], forest, exp_arr, 0.6, 10000, 2, chains[:5], True, 400, ctx)
    for k in range(5):
        assert len(plain[k][0]) == counts[2][k]
        assert set(zip(pooled[k][0].tolist(), pooled[k][1].tolist())) <= set(zip(plain[k][0].tolist(), plain[k][1].tolist()))
    assert sum(len(l[0]) for l in pooled) < sum(len(l[0]) for l in plain)


def test_full_size_matrices_match_the_oracle(workload):
    """Two of the 602-bin matrices through the C oracle: pixel list and probabilities bit for bit."""
    import joblib
    from neoloopfinder_b200.scoring import ScoreBatch
    from oracle import coracle as co
    ctx, mats, _chains, exp_arr, forest = workload
    arrays = co.sklearn_forest_arrays(joblib.load(os.path.join(GOLDEN, "forest_w6.joblib")))
    b = ScoreBatch(4, 400, ctx)
    for G in mats:
        b.add_dense(G)
    b.score(forest, exp_arr, 0.9)
    for k in (1, len(mats) - 2):
        cl, pr, _f = co.score_matrix(mats[k], exp_arr, arrays, 6)
        i, j, p = b.scored(k)
        assert np.array_equal(np.stack([i, j], 1), cl) and np.array_equal(p, pr)
    b.close()
