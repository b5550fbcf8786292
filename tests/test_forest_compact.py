"""The compact single-pass forest (internal nodes only, implicit child behind its parent, swap bit, terminal nodes,
table of distinct leaf values; csrc/score.cu build_compact_forest / k_forest_one) checked on the CPU: the host-only
entry nlf_forest_compact_host hands out the encoding, a few lines of Python walk it exactly as the kernel does, and
the result must equal sklearn's predict_proba bit for bit -- and the real reference's probabilities in the fixtures."""
import ctypes as C
import os

import numpy as np
import pytest

from neoloopfinder_b200 import _lib
from neoloopfinder_b200.forest import flatten_forest
from tests.helpers import GOLDEN, load_case


def compact(flat):
    L = _lib.load()
    u32p = C.POINTER(C.c_uint32)
    n_int, n_leaf = C.c_longlong(0), C.c_longlong(0)
    args = (flat["n_trees"], _lib.ptr(flat["tree_off"], C.c_int64), _lib.ptr(flat["left"], C.c_int), _lib.ptr(flat["right"], C.c_int),
            _lib.ptr(flat["feature"], C.c_int), _lib.ptr(flat["threshold"], C.c_double), _lib.ptr(flat["leaf_p1"], C.c_double))
    rc = _lib.check(L.nlf_forest_compact_host(*args, 0, None, None, 0, None, None, C.byref(n_int), C.byref(n_leaf)))
    if rc == 0:
        return None
    thr = np.empty(max(1, n_int.value), np.uint32)
    word = np.empty(max(1, n_int.value), np.uint32)
    leafv = np.empty(n_leaf.value, np.float64)
    root = np.empty(flat["n_trees"], np.uint32)
    rc = _lib.check(L.nlf_forest_compact_host(*args, thr.size, thr.ctypes.data_as(u32p), word.ctypes.data_as(u32p), leafv.size,
                                              _lib.ptr(leafv, C.c_double), root.ctypes.data_as(u32p), C.byref(n_int), C.byref(n_leaf)))
    assert rc == 1
    return thr[:n_int.value].view(np.float32), word[:n_int.value], leafv, root, n_int.value


def walk(X32, thr, word, leafv, root, n_int):
    """k_forest_one on the host: per row, trees in order, leaf values added sequentially in float64."""
    T = len(root)
    out = np.empty(len(X32))
    for r, x in enumerate(X32):
        acc = 0.0
        for t in range(T):
            slot = int(root[t])
            while slot < n_int:
                w = int(word[slot])
                go_left = bool(x[(w >> 22) & 0xFF] <= thr[slot])
                if w & 0x80000000:                                  # CONT: one child is the implicit next node
                    if go_left != bool(w & 0x40000000):             # swap bit: the implicit child is the right one
                        slot += 1
                    else:
                        slot = w & 0x3FFFFF
                else:                                               # terminal: both children are leaves
                    assert w & 0x40000000
                    slot = n_int + ((w & 0x7FF) if go_left else ((w >> 11) & 0x7FF))
            acc = acc + leafv[slot - n_int]
        out[r] = acc / T
    return out


def test_bench_forest_is_half_the_size_and_equals_the_reference_fixture(models):
    z = load_case("case_transloc_10k")
    model = models(str(z["model"]))
    flat = flatten_forest(model)
    thr, word, leafv, root, n_int = compact(flat)
    total = int(flat["tree_off"][-1])
    assert n_int == int((flat["left"] >= 0).sum()) and n_int < total / 2 + 1
    assert len(leafv) == len(np.unique(flat["leaf_p1"][flat["left"] < 0])) <= 2048
    assert 8 * (n_int + len(leafv)) < 132 * 1024                   # whole forest in the shared memory of one CTA
    X32 = z["fea_sample"].astype(np.float32)                         # rows of the reference's own feature matrix
    got = walk(X32, thr, word, leafv, root, n_int)
    assert np.array_equal(got, model.predict_proba(X32)[:, 1])
    assert np.array_equal(got, z["probas"][z["fea_rows"]])           # what the real reference computed for those pixels


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_forests_incl_one_leaf_trees_and_threshold_ties(seed):
    from sklearn.ensemble import RandomForestClassifier
    rng = np.random.default_rng(seed)
    n, F = 70, 12
    X = rng.random((n, F))
    y = np.zeros(n, dtype=int)
    y[rng.choice(n, size=2 + seed, replace=False)] = 1               # rare class: some bootstrap samples see one class only
    rf = RandomForestClassifier(n_estimators=40, max_depth=[3, 8, None][seed], random_state=seed, n_jobs=1).fit(X, y)
    flat = flatten_forest(rf)
    sizes = np.diff(flat["tree_off"])
    enc = compact(flat)
    assert enc is not None
    thr, word, leafv, root, n_int = enc
    if seed == 0:
        assert (sizes == 1).any() and (root >= n_int).any()          # one-leaf trees: the root slot is a leaf
    Xq = rng.random((150, F)).astype(np.float32)
    # rows that sit exactly on thresholds (float32 of the float64 threshold): '<=' must go left
    internal = np.flatnonzero(flat["left"] >= 0)
    for k, node in enumerate(internal[:60]):
        Xq[k, flat["feature"][node]] = np.float32(flat["threshold"][node])
    assert np.array_equal(walk(Xq, thr, word, leafv, root, n_int), rf.predict_proba(Xq)[:, 1])


def test_forests_that_do_not_fit_the_encoding_are_refused():
    # thresholds outside [0, 2): the two flag bits of the staged words would be lost
    flat = dict(n_trees=1, tree_off=np.array([0, 3], np.int64), left=np.array([1, -1, -1], np.int32),
                right=np.array([2, -1, -1], np.int32), feature=np.array([0, -2, -2], np.int32),
                threshold=np.array([2.5, -2.0, -2.0]), leaf_p1=np.array([0.0, 0.25, 0.75]))
    assert compact(flat) is None
    flat["threshold"] = np.array([0.5, -2.0, -2.0])
    thr, word, leafv, root, n_int = compact(flat)
    assert n_int == 1 and sorted(leafv) == [0.25, 0.75] and not (int(word[0]) & 0x80000000)
    # more than 2 048 distinct leaf values among terminal nodes: 11-bit references cannot hold them
    T = 1500
    off = np.arange(0, 3 * T + 1, 3, dtype=np.int64)
    left = np.tile(np.array([1, -1, -1], np.int32), T)
    right = np.tile(np.array([2, -1, -1], np.int32), T)
    feat = np.tile(np.array([0, -2, -2], np.int32), T)
    thr = np.tile(np.array([0.5, -2.0, -2.0]), T)
    p1 = np.zeros(3 * T)
    p1[1::3] = np.arange(T) / (2.0 * T)
    p1[2::3] = 0.5 + np.arange(T) / (2.0 * T)
    big = dict(n_trees=T, tree_off=off, left=left, right=right, feature=feat, threshold=thr, leaf_p1=p1)
    assert compact(big) is None
    p1[1::3] = (np.arange(T) % 500) / 1000.0
    p1[2::3] = 0.5 + (np.arange(T) % 500) / 1000.0
    assert compact(big) is not None
