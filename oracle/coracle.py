"""ctypes wrapper over oracle/_ref/libnlf_oracle.so.  TEST INFRASTRUCTURE ONLY.

High-level helpers mirror the stages of pipeline() (scripts/neoloop-caller:185-218) so the
parity tests read stage by stage.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        L = C.CDLL(path)
        dp, ip, lp, fp, bp = (C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long),
                              C.POINTER(C.c_float), C.POINTER(C.c_ubyte))
        L.nlfo_pairwise_sum.restype = C.c_double
        L.nlfo_pairwise_sum.argtypes = [dp, C.c_long, C.c_long]
        L.nlfo_block_diag_means.restype = None
        L.nlfo_block_diag_means.argtypes = [dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, dp, ip]
        L.nlfo_scale_blocks.restype = None
        L.nlfo_scale_blocks.argtypes = [dp, C.c_int, ip, ip, C.c_int, ip, dp]
        L.nlfo_remove_gaps_index.restype = C.c_int
        L.nlfo_remove_gaps_index.argtypes = [dp, C.c_int, ip]
        L.nlfo_isotonic_decreasing_fit.restype = None
        L.nlfo_isotonic_decreasing_fit.argtypes = [dp, C.c_long, dp]
        L.nlfo_candidates.restype = C.c_long
        L.nlfo_candidates.argtypes = [dp, C.c_int, C.c_int, C.c_int, ip, ip, dp]
        L.nlfo_getwindow.restype = C.c_long
        L.nlfo_getwindow.argtypes = [dp, C.c_int, C.c_int, ip, ip, C.c_long, C.c_int, dp, C.c_int, dp,
                                     C.c_int, dp, ip]
        L.nlfo_forest_predict.restype = None
        L.nlfo_forest_predict.argtypes = [C.c_int, lp, lp, lp, lp, dp, bp, dp, fp, C.c_long, C.c_int, dp]
        L.nlfo_pool_bucket.restype = C.c_int
        L.nlfo_pool_bucket.argtypes = [ip, ip, dp, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, bp]
        L.nlfo_anchor_priorities.restype = C.c_int
        L.nlfo_anchor_priorities.argtypes = [ip, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.nlfo_num_threads.restype = C.c_int
        i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        L.nlfo_assemble_pixels.restype = None
        L.nlfo_assemble_pixels.argtypes = [i64p, i32p, i32p, dp, C.c_int, C.c_int, i64p, i64p, bp, dp, dp]
        L.nlfo_expected_chrom.restype = None
        L.nlfo_expected_chrom.argtypes = [i64p, i32p, i32p, dp, C.c_int, C.c_int64, C.c_int64, C.c_int, dp, lp]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def gaussian_weights(sigma=1.0, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius)[::-1], computed with the
    same numpy operations scipy uses (so it tracks the local numpy build like the reference)."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[::-1]), radius


def pairwise_sum(a):
    a = _f64(a)
    return lib().nlfo_pairwise_sum(_p(a, C.c_double), a.size, 1)


def block_diag_means(M, bias, blk1, blk2, maxdiag, mink=3):
    """-> dict {i: mean} as Fusion._extract_xy returns for maxdis = maxdiag*res."""
    M = _f64(M)
    bias = _f64(bias)
    n = M.shape[0]
    mean = np.empty(maxdiag + 1)
    cnt = np.empty(maxdiag + 1, dtype=np.int32)
    lib().nlfo_block_diag_means(_p(M, C.c_double), n, _p(bias, C.c_double), blk1[0], blk1[1], blk2[0], blk2[1],
                                maxdiag, mink, _p(mean, C.c_double), _p(cnt, C.c_int))
    return {int(i): float(mean[i]) for i in range(maxdiag + 1) if not np.isnan(mean[i])}


def scale_blocks(M, block_index, op, val):
    hcm = _f64(M).copy()
    bi = np.asarray(block_index)
    lo, hi = _i32(bi[:, 0]), _i32(bi[:, 1])
    op = _i32(op)
    val = _f64(val)
    lib().nlfo_scale_blocks(_p(hcm, C.c_double), hcm.shape[0], _p(lo, C.c_int), _p(hi, C.c_int), len(lo),
                            _p(op, C.c_int), _p(val, C.c_double))
    return hcm


def remove_gaps_index(hcm):
    hcm = _f64(hcm)
    n = hcm.shape[0]
    kept = np.empty(n, dtype=np.int32)
    dim = lib().nlfo_remove_gaps_index(_p(hcm, C.c_double), n, _p(kept, C.c_int))
    return kept[:dim].copy()


def isotonic_decreasing(y):
    y = _f64(y)
    out = np.empty_like(y)
    lib().nlfo_isotonic_decreasing_fit(_p(y, C.c_double), y.size, _p(out, C.c_double))
    return out


def candidates(G, lower=4, upper=400):
    G = _f64(G)
    n = G.shape[0]
    e = np.full(upper + 1, np.nan)
    nc = lib().nlfo_candidates(_p(G, C.c_double), n, lower, upper, None, None, _p(e, C.c_double))
    if nc < 0:
        raise ValueError("isotonic fit on an empty sample (matrix has no diagonal longer than 5)")
    r = np.empty(nc, dtype=np.int32)
    c = np.empty(nc, dtype=np.int32)
    lib().nlfo_candidates(_p(G, C.c_double), n, lower, upper, _p(r, C.c_int), _p(c, C.c_int), _p(e, C.c_double))
    return r, c, e


def getwindow(G, xi, yi, w, exp_arr, upper=400, gw=None):
    G = _f64(G)
    xi, yi = _i32(xi), _i32(yi)
    exp_arr = _f64(exp_arr)
    if gw is None:
        gw, radius = gaussian_weights()
    else:
        gw = _f64(gw)
        radius = (gw.size - 1) // 2
    S = 2 * w + 1
    fea = np.empty((max(1, xi.size), S * S), dtype=np.float64)
    clist = np.empty((max(1, xi.size), 2), dtype=np.int32)
    k = lib().nlfo_getwindow(_p(G, C.c_double), G.shape[0], upper, _p(xi, C.c_int), _p(yi, C.c_int), xi.size, w,
                             _p(exp_arr, C.c_double), exp_arr.size, _p(gw, C.c_double), radius,
                             _p(fea, C.c_double), _p(clist, C.c_int))
    return fea[:k], clist[:k]


def sklearn_forest_arrays(model):
    """Concatenate the RAW sklearn tree arrays (no repacking: this is the independent check
    of the product's own flattening)."""
    off = [0]
    left, right, feat, thr, ml, p1 = [], [], [], [], [], []
    cls = list(model.classes_)
    k1 = cls.index(1) if 1 in cls else len(cls) - 1
    for est in model.estimators_:
        t = est.tree_
        left.append(t.children_left.astype(np.int64))
        right.append(t.children_right.astype(np.int64))
        feat.append(np.where(t.feature >= 0, t.feature, 0).astype(np.int64))
        thr.append(t.threshold.astype(np.float64))
        mgl = getattr(t, "missing_go_to_left", None)
        ml.append(np.zeros(t.node_count, np.uint8) if mgl is None else np.asarray(mgl, dtype=np.uint8))
        p1.append(t.value[:, 0, k1].astype(np.float64))
        off.append(off[-1] + t.node_count)
    return (np.array(off, dtype=np.int64), np.concatenate(left), np.concatenate(right), np.concatenate(feat),
            np.concatenate(thr), np.concatenate(ml), np.concatenate(p1))


def forest_predict(arrays, X32):
    off, left, right, feat, thr, ml, p1 = arrays
    X32 = np.ascontiguousarray(X32, dtype=np.float32)
    N, F = X32.shape
    out = np.empty(N, dtype=np.float64)
    lib().nlfo_forest_predict(len(off) - 1, _p(off, C.c_long), _p(left, C.c_long), _p(right, C.c_long),
                              _p(feat, C.c_long), _p(thr, C.c_double), _p(ml, C.c_ubyte), _p(p1, C.c_double),
                              _p(X32, C.c_float), N, F, _p(out, C.c_double))
    return out


def _numpy_tie_order(pos, res, min_count, r):
    """The permutation numpy.argsort (default kind, as scipy's _select_by_peak_distance calls
    it) returns for the peak heights of this coordinate histogram on THIS machine, or None
    when the distance filter cannot remove anything (distance <= 2)."""
    if max(r // res, 2) <= 2 or pos.size == 0:
        return None
    prio = np.empty(pos.size + 1, dtype=np.float64)
    m = lib().nlfo_anchor_priorities(_p(pos, C.c_int), pos.size, res, min_count, r, _p(prio, C.c_double))
    return np.ascontiguousarray(np.argsort(prio[:m]), dtype=np.int32)


def pool_bucket(pi, pj, val, res, min_count=2, r=15000, numpy_ties=True):
    pi, pj, val = _i32(pi), _i32(pj), _f64(val)
    keep = np.zeros(pi.size, dtype=np.uint8)
    xo = _numpy_tie_order(pi, res, min_count, r) if numpy_ties else None
    yo = _numpy_tie_order(pj, res, min_count, r) if numpy_ties else None
    lib().nlfo_pool_bucket(_p(pi, C.c_int), _p(pj, C.c_int), _p(val, C.c_double), pi.size, res, min_count, r,
                           None if xo is None else _p(xo, C.c_int), None if yo is None else _p(yo, C.c_int),
                           _p(keep, C.c_ubyte))
    return keep.astype(bool)


def local_clustering(ij, val, res, min_count=2, r=15000, chain_of_i=None, chain_of_j=None, numpy_ties=True):
    """Peakachu.local_clustering (neoloop/callers.py:668-697): bucket by (chain_i, chain_j)."""
    ij = np.asarray(ij, dtype=np.int64).reshape(-1, 2)
    val = _f64(val)
    if chain_of_i is None:
        key = np.zeros(len(ij), dtype=np.int64)
    else:
        key = np.asarray(chain_of_i, dtype=np.int64) * (1 << 20) + np.asarray(chain_of_j, dtype=np.int64)
    out = set()
    for k in np.unique(key):
        sel = np.where(key == k)[0]
        keep = pool_bucket(ij[sel, 0], ij[sel, 1], val[sel], res, min_count, r, numpy_ties)
        for a, b in ij[sel][keep]:
            out.add((int(a), int(b)))
    return out


def score_matrix(G, exp_arr, model_arrays, w, upper=400, lower=4):
    """gap-free matrix -> (clist, probas, fea32): the scoring path K5-K8 on the CPU."""
    r, c, e = candidates(G, lower, upper)
    fea, clist = getwindow(G, r, c, w, exp_arr, upper)
    fea32 = fea.astype(np.float32)
    if len(fea32) < 2:
        return clist, np.zeros(0), fea32
    return clist, forest_predict(model_arrays, fea32), fea32


def _pixel_table(off, bin2, cnt, w):
    return (np.ascontiguousarray(off, dtype=np.int64), np.ascontiguousarray(bin2, dtype=np.int32),
            np.ascontiguousarray(cnt, dtype=np.int32), _f64(w))


def assemble_pixels(off, bin2, cnt, w, src_lo, src_hi, rev, balance=True):
    """complexSV.reorganize_matrix on cooler's pixel table -> (fusion_matrix, bias)."""
    off, bin2, cnt, w = _pixel_table(off, bin2, cnt, w)
    lo = np.ascontiguousarray(src_lo, dtype=np.int64)
    hi = np.ascontiguousarray(src_hi, dtype=np.int64)
    rv = np.ascontiguousarray(rev, dtype=np.uint8)
    n = int((hi - lo).sum())
    M = np.empty((n, n), dtype=np.float64)
    bias = np.empty(n, dtype=np.float64)
    lib().nlfo_assemble_pixels(_p(off, C.c_int64), _p(bin2, C.c_int32), _p(cnt, C.c_int32), _p(w, C.c_double),
                               1 if balance else 0, len(lo), _p(lo, C.c_int64), _p(hi, C.c_int64), _p(rv, C.c_ubyte),
                               _p(M, C.c_double), _p(bias, C.c_double))
    return M, bias


def expected_chrom(off, bin2, cnt, w, lo, hi, nd, balance=True):
    """util._prepare_core on cooler's pixel table -> (sums[nd], counts[nd])."""
    off, bin2, cnt, w = _pixel_table(off, bin2, cnt, w)
    nd = min(int(nd), int(hi) - int(lo))
    sums = np.empty(nd, dtype=np.float64)
    cnts = np.empty(nd, dtype=np.int64)
    lib().nlfo_expected_chrom(_p(off, C.c_int64), _p(bin2, C.c_int32), _p(cnt, C.c_int32), _p(w, C.c_double),
                              1 if balance else 0, int(lo), int(hi), nd, _p(sums, C.c_double), _p(cnts, C.c_long))
    return sums, cnts
