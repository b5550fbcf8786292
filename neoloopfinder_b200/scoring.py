"""Python face of the CUDA scoring path (batch of gap-free matrices -> Donuts -> pooled loops).

Thin, typed wrappers over include/nlf_b200.h; the reference-shaped classes in ``callers.py`` /
``assembly.py`` are built on these.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, f64, i32, ptr

__all__ = ["gaussian_weights", "ScoreBatch", "DeviceAssembly", "DevicePixels", "pool_buckets", "pool_flat", "local_clustering_gpu",
           "score_and_pool", "score_matrices"]


def gaussian_weights(sigma: float = 1.0, truncate: float = 4.0):
    """Kernel of ``scipy.ndimage.gaussian_filter(sigma=1)`` (scipy/ndimage/_filters.py
    ``_gaussian_kernel1d``), computed on the host with the same numpy operations scipy uses, so it
    tracks the local numpy build exactly as the reference does.  The device never calls exp()."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[::-1], dtype=np.float64), radius


class DevicePixels:
    """A sample's pixel table resident in HBM (cooler's layout: ``bin1_offset`` CSR index, ``bin2``,
    ``count`` for bin1 <= bin2, one balancing weight per bin).  Uploaded once per sample, GPU and
    weight column; assemblies, chromosome bands and the expected prologue are built from it on
    the device (include/nlf_b200.h, nlf_pixels_*)."""

    def __init__(self, bin1_offset, bin2, count, weights, ctx=None):
        self.ctx = ctx or _lib.Context.get()
        off = np.ascontiguousarray(bin1_offset, dtype=np.int64)
        b2 = np.ascontiguousarray(bin2, dtype=np.int32)
        cnt = np.ascontiguousarray(count, dtype=np.int32)
        w = f64(weights)
        self.nbins = int(w.size)
        if off.size != self.nbins + 1 or b2.size != cnt.size or (off.size and int(off[-1]) != b2.size):
            raise ValueError("inconsistent pixel table: bin1_offset / bin2 / count / weights sizes")
        self.nnz = int(b2.size)
        self.h = C.c_void_p()
        check(_lib.load().nlf_pixels_create(self.ctx.h, self.nbins, ptr(off, C.c_int64), ptr(b2, C.c_int),
                                            ptr(cnt, C.c_int), ptr(w, C.c_double), C.byref(self.h)))

    def expected_diagonals(self, lo: int, hi: int, nd: int, balance: bool = True):
        """util._prepare_core for the chromosome of global bins [lo, hi): (sums[nd'], counts[nd'])
        with nd' = min(nd, hi - lo)."""
        nd = min(int(nd), int(hi) - int(lo))
        sums = np.empty(nd, dtype=np.float64)
        cnts = np.empty(nd, dtype=np.int64)
        check(_lib.load().nlf_expected_from_pixels(self.ctx.h, self.h, 1 if balance else 0, int(lo), int(hi), nd,
                                                   ptr(sums, C.c_double), ptr(cnts, C.c_longlong)))
        return sums, cnts

    def close(self):
        if self.h:
            _lib.load().nlf_pixels_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceAssembly:
    """Device-resident fusion matrix of one assembly (K2-K4)."""

    @classmethod
    def from_pixels(cls, pixels: DevicePixels, src_lo, src_hi, reverse, balance: bool = True):
        """complexSV.reorganize_matrix on the device: block b = global bins [src_lo[b], src_hi[b])
        of ``pixels``, traversed in reverse when ``reverse[b]``."""
        self = cls.__new__(cls)
        self.ctx = pixels.ctx
        lo = np.ascontiguousarray(src_lo, dtype=np.int64)
        hi = np.ascontiguousarray(src_hi, dtype=np.int64)
        rev = np.ascontiguousarray(reverse, dtype=np.uint8)
        self.nblk = int(lo.size)
        self.h = C.c_void_p()
        check(_lib.load().nlf_assembly_create_from_pixels(self.ctx.h, pixels.h, 1 if balance else 0, ptr(lo, C.c_int64),
                                                          ptr(hi, C.c_int64), ptr(rev, C.c_ubyte), self.nblk,
                                                          C.byref(self.h)))
        self.n = int((hi - lo).sum())
        self._kept = None
        return self

    def fusion_matrix(self) -> np.ndarray:
        out = np.empty((self.n, self.n), dtype=np.float64)
        check(_lib.load().nlf_assembly_fetch_matrix(self.h, ptr(out, C.c_double), None))
        return out

    def bias(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        check(_lib.load().nlf_assembly_fetch_matrix(self.h, None, ptr(out, C.c_double)))
        return out

    def __init__(self, M, bias, block_index, ctx=None):
        self.ctx = ctx or _lib.Context.get()
        M = f64(M)
        if M.ndim != 2 or M.shape[0] != M.shape[1]:
            raise ValueError("fusion matrix must be square")
        self.n = M.shape[0]
        bias = f64(bias)
        bi = np.asarray(block_index, dtype=np.int64).reshape(-1, 2)
        self.nblk = len(bi)
        lo, hi = i32(bi[:, 0]), i32(bi[:, 1])
        self.h = C.c_void_p()
        check(_lib.load().nlf_assembly_create(self.ctx.h, ptr(M, C.c_double), self.n, ptr(bias, C.c_double),
                                              ptr(lo, C.c_int), ptr(hi, C.c_int), self.nblk, C.byref(self.h)))
        self._kept = None

    def diag_means(self, maxdiag: int, mink: int = 3):
        """-> (mean[npairs, maxdiag+1] with NaN where not kept, count[npairs, maxdiag+1])."""
        npairs = self.nblk * (self.nblk + 1) // 2
        mean = np.empty((npairs, maxdiag + 1), dtype=np.float64)
        cnt = np.empty((npairs, maxdiag + 1), dtype=np.int32)
        check(_lib.load().nlf_assembly_diag_means(self.h, int(maxdiag), int(mink), ptr(mean, C.c_double),
                                                  ptr(cnt, C.c_int)))
        return mean, cnt

    def set_bias(self, bias):
        b = f64(bias)
        if b.size != self.n:
            raise ValueError("bias vector must have one entry per bin")
        check(_lib.load().nlf_assembly_set_bias(self.h, ptr(b, C.c_double)))
        self.ctx.sync()

    def rect_means(self, rects, maxdiag: int, mink: int = 3):
        """Fusion._extract_xy for rectangles [(row_lo, row_hi, col_lo, col_hi), ...] of the matrix:
        -> (mean[nrect, maxdiag+1] with NaN where not kept, count[nrect, maxdiag+1])."""
        r = np.asarray(rects, dtype=np.int64).reshape(-1, 4)
        cols = [i32(r[:, k]) for k in range(4)]
        mean = np.empty((len(r), maxdiag + 1), dtype=np.float64)
        cnt = np.empty((len(r), maxdiag + 1), dtype=np.int32)
        check(_lib.load().nlf_assembly_rect_means(self.h, len(r), ptr(cols[0], C.c_int), ptr(cols[1], C.c_int),
                                                  ptr(cols[2], C.c_int), ptr(cols[3], C.c_int), int(maxdiag), int(mink),
                                                  ptr(mean, C.c_double), ptr(cnt, C.c_int)))
        return mean, cnt

    def set_scales(self, op, val):
        op = i32(op).reshape(self.nblk, self.nblk)
        val = f64(val).reshape(self.nblk, self.nblk)
        check(_lib.load().nlf_assembly_set_scales(self.h, ptr(op, C.c_int), ptr(val, C.c_double)))
        self._kept = None

    def kept_index(self) -> np.ndarray:
        if self._kept is None:
            kept = np.empty(self.n, dtype=np.int32)
            nk = C.c_int(0)
            check(_lib.load().nlf_assembly_remove_gaps(self.h, ptr(kept, C.c_int), C.byref(nk)))
            self._kept = kept[:nk.value].copy()
        return self._kept

    def hcm(self) -> np.ndarray:
        out = np.empty((self.n, self.n), dtype=np.float64)
        check(_lib.load().nlf_assembly_fetch_hcm(self.h, ptr(out, C.c_double)))
        return out

    def gap_free(self) -> np.ndarray:
        nk = len(self.kept_index())
        out = np.empty((nk, nk), dtype=np.float64)
        check(_lib.load().nlf_assembly_fetch_gap_free(self.h, ptr(out, C.c_double)))
        return out

    def close(self):
        if self.h:
            _lib.load().nlf_assembly_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ScoreBatch:
    """A batch of gap-free matrices scored together (K1, K5-K9)."""

    def __init__(self, lower: int = 4, upper: int = 400, ctx=None):
        self.ctx = ctx or _lib.Context.get()
        self.lower, self.upper = int(lower), int(upper)
        self.h = C.c_void_p()
        check(_lib.load().nlf_batch_create(self.ctx.h, self.lower, self.upper, C.byref(self.h)))
        self._keepalive = []
        self.njobs = 0
        self.n = []
        self.w = None
        self._counts = None

    # ---- inputs
    def add_dense(self, G) -> int:
        G = f64(G)
        if G.ndim != 2 or G.shape[0] != G.shape[1]:
            raise ValueError("matrix must be square")
        self._keepalive.append(G)
        j = check(_lib.load().nlf_batch_add_dense(self.h, ptr(G, C.c_double), G.shape[0]))
        self.njobs += 1
        self.n.append(G.shape[0])
        return j

    def add_assembly(self, asm: DeviceAssembly) -> int:
        j = check(_lib.load().nlf_batch_add_assembly(self.h, asm.h))
        self.njobs += 1
        self.n.append(len(asm.kept_index()))
        return j

    def add_pixels(self, pixels: DevicePixels, lo: int, hi: int, balance: bool = True):
        """A whole chromosome (global bins [lo, hi) of a resident pixel table) as a trivial assembly;
        returns (job, kept) with kept[k] = bin (relative to lo) behind gap-free index k."""
        kept = np.empty(int(hi) - int(lo), dtype=np.int32)
        nk = C.c_int(0)
        j = check(_lib.load().nlf_batch_add_pixels(self.h, pixels.h, 1 if balance else 0, int(lo), int(hi),
                                                   ptr(kept, C.c_int), C.byref(nk)))
        self.njobs += 1
        self.n.append(nk.value)
        return j, kept[:nk.value].copy()

    def add_band(self, band) -> int:
        band = f64(band)
        self._keepalive.append(band)
        j = check(_lib.load().nlf_batch_add_band(self.h, ptr(band, C.c_double), band.shape[0], band.shape[1]))
        self.njobs += 1
        self.n.append(band.shape[0])
        return j

    def keep_features(self, on=True):
        check(_lib.load().nlf_batch_keep_features(self.h, 1 if on else 0))

    # ---- compute
    def score(self, forest, exp_arr, thre: float, w: Optional[int] = None, gw=None):
        exp_arr = f64(exp_arr)
        if gw is None:
            gw, radius = gaussian_weights()
        else:
            gw = f64(gw)
            radius = (gw.size - 1) // 2
        self.w = int(forest.w if w is None else w)
        check(_lib.load().nlf_batch_score(self.h, forest.h, ptr(exp_arr, C.c_double), exp_arr.size, self.w,
                                          ptr(gw, C.c_double), radius, float(thre)))
        self._counts = None
        self._keepalive.clear()

    def rethreshold(self, thre: float):
        """Another probability threshold on the probabilities already in HBM (no re-scoring)."""
        check(_lib.load().nlf_batch_rethreshold(self.h, float(thre)))
        self._counts = None

    def reset(self):
        """Forget the results, keep the matrices in HBM (lets the scoring path run again)."""
        check(_lib.load().nlf_batch_reset(self.h))
        self._counts = None

    # ---- outputs
    def counts(self):
        if self._counts is None:
            a = np.zeros(max(1, self.njobs), dtype=np.int64)
            b = np.zeros_like(a)
            c = np.zeros_like(a)
            check(_lib.load().nlf_batch_counts(self.h, ptr(a, C.c_longlong), ptr(b, C.c_longlong), ptr(c, C.c_longlong)))
            self._counts = (a[:self.njobs], b[:self.njobs], c[:self.njobs])
        return self._counts

    def donuts(self, job: int):
        m = int(self.counts()[2][job])
        i = np.empty(m, np.int32); j = np.empty(m, np.int32)
        p = np.empty(m, np.float64); v = np.empty(m, np.float64)
        if m:
            check(_lib.load().nlf_batch_fetch_donuts(self.h, job, ptr(i, C.c_int), ptr(j, C.c_int),
                                                     ptr(p, C.c_double), ptr(v, C.c_double)))
        return i, j, p, v

    def donuts_all(self):
        """(job, i, j, prob, value) arrays of every Donut of the batch: one device-to-host copy."""
        n = C.c_longlong(0)
        L = _lib.load()
        check(L.nlf_batch_fetch_donuts_all(self.h, 0, None, None, None, None, None, C.byref(n)))
        m = n.value
        job = np.empty(m, np.int32); i = np.empty(m, np.int32); j = np.empty(m, np.int32)
        p = np.empty(m, np.float64); v = np.empty(m, np.float64)
        if m:
            check(L.nlf_batch_fetch_donuts_all(self.h, m, ptr(job, C.c_int), ptr(i, C.c_int), ptr(j, C.c_int),
                                               ptr(p, C.c_double), ptr(v, C.c_double), C.byref(n)))
        return job, i, j, p, v

    def scored(self, job: int):
        m = int(self.counts()[1][job])
        i = np.empty(m, np.int32); j = np.empty(m, np.int32); p = np.empty(m, np.float64)
        check(_lib.load().nlf_batch_fetch_scored(self.h, job, m, ptr(i, C.c_int), ptr(j, C.c_int), ptr(p, C.c_double)))
        return i, j, p

    def candidates(self, job: int):
        m = int(self.counts()[0][job])
        r = np.empty(m, np.int32); c = np.empty(m, np.int32)
        check(_lib.load().nlf_batch_fetch_candidates(self.h, job, m, ptr(r, C.c_int), ptr(c, C.c_int)))
        return r, c

    def expected(self, job: int):
        """(diagonal means, isotonic expected) of get_candidates, indexed by the diagonal."""
        means = np.empty(self.upper + 1); e = np.empty(self.upper + 1)
        check(_lib.load().nlf_batch_fetch_expected(self.h, job, ptr(means, C.c_double), ptr(e, C.c_double)))
        return means, e

    def features32(self, job: int):
        m = int(self.counts()[1][job])
        F = (2 * self.w + 1) ** 2
        fea = np.empty((m, F), dtype=np.float32)
        check(_lib.load().nlf_batch_fetch_features32(self.h, job, m, ptr(fea, C.c_float)))
        return fea

    def getwindow(self, job: int, xi, yi, w: int, exp_arr, gw=None):
        xi, yi = i32(xi), i32(yi)
        exp_arr = f64(exp_arr)
        if gw is None:
            gw, radius = gaussian_weights()
        else:
            gw = f64(gw); radius = (gw.size - 1) // 2
        F = (2 * w + 1) ** 2
        fea = np.empty((max(1, xi.size), F), dtype=np.float64)
        clist = np.empty((max(1, xi.size), 2), dtype=np.int32)
        nrows = C.c_longlong(0)
        check(_lib.load().nlf_batch_getwindow(self.h, job, ptr(xi, C.c_int), ptr(yi, C.c_int), xi.size, int(w),
                                              ptr(exp_arr, C.c_double), exp_arr.size, ptr(gw, C.c_double), radius,
                                              ptr(fea, C.c_double), ptr(clist, C.c_int), C.byref(nrows)))
        return fea[:nrows.value], clist[:nrows.value]

    def close(self):
        if self.h:
            _lib.load().nlf_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pool_flat(off, pi, pj, val, res: int, min_count: int = 2, r: int = 15000, ctx=None) -> np.ndarray:
    """Peakachu._local_clustering for buckets given as one concatenated point list
    (points of bucket k = [off[k], off[k+1])) -> keep mask over the points.

    Below 10 kb scipy's peak-distance filter can remove peaks, and among EQUAL peak heights its
    outcome follows numpy.argsort's implementation-defined tie order.  The kernel reports when
    that can matter; only then are the peak heights brought back and numpy.argsort -- the very
    call the reference makes (scipy/signal/_peak_finding.py _select_by_peak_distance) -- asked for
    the permutation, which the pooling kernel then follows.
    """
    ctx = ctx or _lib.Context.get()
    L = _lib.load()
    off = np.ascontiguousarray(off, dtype=np.int64)
    nb = len(off) - 1
    total = int(off[-1]) if nb > 0 else 0
    if nb <= 0 or total == 0:
        return np.zeros(total, dtype=bool)
    pi, pj, val = i32(pi), i32(pj), f64(val)
    xo = yo = None
    xo_off = yo_off = None
    if max(r // res, 2) > 2:
        orders = []
        for pos in (pi, pj):
            prio = np.empty(total, dtype=np.float64)
            poff = np.zeros(nb + 1, dtype=np.int64)
            amb = C.c_int(0)
            check(L.nlf_pool_peak_heights(ctx.h, nb, ptr(off, C.c_int64), ptr(pos, C.c_int), int(res), int(min_count),
                                          int(r), ptr(prio, C.c_double), ptr(poff, C.c_int64), C.byref(amb)))
            orders.append((prio, poff, amb.value))
        if orders[0][2] or orders[1][2]:
            def perm(prio, poff):
                out = np.empty(int(poff[-1]) + 1, dtype=np.int32)
                for k in range(nb):
                    a, b = int(poff[k]), int(poff[k + 1])
                    out[a:b] = np.argsort(prio[a:b])        # default kind, as scipy calls it
                return out
            xo, xo_off = perm(*orders[0][:2]), orders[0][1]
            yo, yo_off = perm(*orders[1][:2]), orders[1][1]
    keep = np.zeros(total, dtype=np.uint8)
    check(L.nlf_pool(ctx.h, nb, ptr(off, C.c_int64), ptr(pi, C.c_int), ptr(pj, C.c_int), ptr(val, C.c_double),
                     int(res), int(min_count), int(r),
                     None if xo is None else ptr(xo, C.c_int), None if xo is None else ptr(xo_off, C.c_int64),
                     None if yo is None else ptr(yo, C.c_int), None if yo is None else ptr(yo_off, C.c_int64),
                     ptr(keep, C.c_ubyte)))
    return keep.astype(bool)


def pool_buckets(buckets: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]], res: int, min_count: int = 2,
                 r: int = 15000, ctx=None) -> List[np.ndarray]:
    """Peakachu._local_clustering for a list of buckets [(i, j, value)] -> list of keep masks."""
    nb = len(buckets)
    if nb == 0:
        return []
    sizes = [len(b[0]) for b in buckets]
    off = np.zeros(nb + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    if int(off[-1]) == 0:
        return [np.zeros(0, dtype=bool) for _ in buckets]
    pi = np.concatenate([np.asarray(b[0]) for b in buckets])
    pj = np.concatenate([np.asarray(b[1]) for b in buckets])
    val = np.concatenate([np.asarray(b[2]) for b in buckets])
    keep = pool_flat(off, pi, pj, val, res, min_count, r, ctx)
    return [keep[off[k]:off[k + 1]] for k in range(nb)]


def local_clustering_gpu(ij, val, res, min_count=2, r=15000, chain_i=None, chain_j=None, ctx=None):
    """Peakachu.local_clustering (callers.py:668-697): bucket by (chain_i, chain_j), pool each
    bucket on the GPU, return the set of kept (i, j)."""
    ij = np.asarray(ij, dtype=np.int64).reshape(-1, 2)
    val = f64(val)
    if len(ij) == 0:
        return set()
    if chain_i is None:
        key = np.zeros(len(ij), dtype=np.int64)
    else:
        key = np.asarray(chain_i, dtype=np.int64) * (1 << 24) + np.asarray(chain_j, dtype=np.int64)
    uniq = np.unique(key)
    sels = [np.where(key == k)[0] for k in uniq]
    masks = pool_buckets([(ij[s, 0], ij[s, 1], val[s]) for s in sels], res, min_count, r, ctx)
    out = set()
    for s, m in zip(sels, masks):
        for a, b in ij[s][m]:
            out.add((int(a), int(b)))
    return out


def _pool_scored(batches: Sequence[ScoreBatch], thre: float, res: int, min_count: int, chains: Optional[Sequence],
                 no_pool: bool, r: int):
    """Threshold results of one or more scored batches -> pooled loops, with ONE pooling pass over
    the Donuts of all of them (jobs are numbered consecutively across the batches)."""
    cand, scored, don, parts = [], [], [], []
    base = 0
    for batch in batches:
        n_cand, n_scored, n_don = batch.counts()
        dj_, di, dj, dp, dv = batch.donuts_all()
        cand.append(n_cand); scored.append(n_scored); don.append(n_don)
        parts.append((dj_.astype(np.int64) + base, di, dj, dp, dv))
        base += batch.njobs
    nj = base
    n_cand, n_scored, n_don = (np.concatenate(x) if len(x) > 1 else x[0] for x in (cand, scored, don))
    if len(parts) > 1:
        dj_, di, dj, dp, dv = (np.concatenate([p[k] for p in parts]) for k in range(5))
    else:
        dj_, di, dj, dp, dv = parts[0]
    empty = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64))
    # callers.py:619-620: fewer than two scored windows -> no call at all for that matrix
    if len(dj_):
        ok = n_scored[dj_] >= 2
        if not ok.all():
            dj_, di, dj, dp, dv = dj_[ok], di[ok], dj[ok], dp[ok], dv[ok]
    if len(dj_) == 0:
        return [empty] * nj, (n_cand, n_scored, n_don)
    if no_pool:
        keep = np.ones(len(dj_), dtype=bool)
        order = np.argsort(dj_, kind="stable")
    else:
        # bucket key = (job, chain of i, chain of j)   (callers.py:671-683)
        dj64 = dj_.astype(np.int64)
        if chains is not None:
            arrs = [np.zeros(0, np.int64) if ch is None else np.asarray(ch, dtype=np.int64) for ch in chains]
            lens = np.fromiter((a.size for a in arrs), np.int64, len(arrs))
            coff = np.zeros(len(arrs) + 1, np.int64)
            np.cumsum(lens, out=coff[1:])
            cat = np.concatenate(arrs + [np.zeros(1, np.int64)])
            has = lens[dj64] > 0
            base = np.where(has, coff[dj64], cat.size - 1)
            ci = cat[base + np.where(has, di, 0)]
            cj = cat[base + np.where(has, dj, 0)]
            key = (dj64 << 40) | (ci << 20) | cj
        else:
            key = dj64 << 40
        order = np.argsort(key, kind="stable")
        ks = key[order]
        starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
        off = np.r_[starts, len(ks)].astype(np.int64)
        keep_sorted = pool_flat(off, di[order], dj[order], dv[order], res, min_count, r, batches[0].ctx)
        keep = np.empty(len(dj_), dtype=bool)
        keep[order] = keep_sorted
    loops = [empty] * nj
    sel = order[keep[order]]
    if len(sel):
        js = dj_[sel]
        oi, oj, op = di[sel], dj[sel], dp[sel]
        cuts = np.flatnonzero(np.r_[True, js[1:] != js[:-1]]).tolist()
        ends = cuts[1:] + [len(sel)]
        for a, e, job in zip(cuts, ends, js[cuts].tolist()):
            loops[job] = (oi[a:e], oj[a:e], op[a:e])
    return loops, (n_cand, n_scored, n_don)


def score_and_pool(batch: ScoreBatch, forest, exp_arr, thre: float, res: int, min_count: int = 2,
                   chains: Optional[Sequence] = None, no_pool: bool = False, r: int = 15000):
    """The scoring path for a whole batch: candidates -> features -> forest -> threshold ->
    pooling (callers.py:614-639 for every matrix of the batch).

    chains[job] = per-gap-free-bin block id (``chains[index_map[i]]`` of the reference) or None.
    Returns (loops, counts): loops[job] = (i, j, prob) arrays of the called loops, counts =
    (n_candidates, n_scored, n_donuts) arrays.
    """
    batch.score(forest, exp_arr, thre, forest.w)
    return _pool_scored([batch], thre, res, min_count, chains, no_pool, r)


PIPELINE_MIN_JOBS = 8       # matrices per sub-batch below which splitting a batch stops paying
PIPELINE_DEPTH = 4          # sub-batches of one score_matrices call
PIPELINE_FIRST = 0.5        # size of the first sub-batch relative to an even split


def score_matrices(mats: Sequence[np.ndarray], forest, exp_arr, thre: float, res: int, min_count: int = 2,
                   chains: Optional[Sequence] = None, no_pool: bool = False, upper: int = 400, ctx=None,
                   depth: Optional[int] = None):
    """Host gap-free matrices in -> called loops out on one GPU (the end-to-end call ``bench.py``
    times: host->device copies, every kernel, device->host results).

    The matrices are cut into up to ``depth`` sub-batches.  Uploads run on the context's copy
    stream, scoring on its main stream, and nothing here waits for the device until every
    sub-batch is enqueued: the upload of sub-batch k+1 overlaps the kernels of sub-batch k.
    Results are independent per matrix, so the split does not change them."""
    mats = list(mats)
    nj = len(mats)
    if depth is None:
        # 4 sub-batches; 8 for long lists (cfg3, 189 matrices: 104.0 -> 101.0 ms per end-to-end step)
        depth = int(os.environ.get("NLF_PIPELINE_DEPTH", PIPELINE_DEPTH if nj < 96 else 2 * PIPELINE_DEPTH))
    depth = max(1, min(depth, nj // PIPELINE_MIN_JOBS))
    # the first upload cannot hide under anything: make the first sub-batch smaller than the others
    first = float(os.environ.get("NLF_PIPELINE_FIRST", PIPELINE_FIRST))
    if depth > 1:
        n0 = max(1, int(round(nj / depth * first)))
        cuts = [0] + [n0 + (nj - n0) * k // (depth - 1) for k in range(depth)]
    else:
        cuts = [0, nj]
    # every sub-batch object first: creating one orders the copy stream after the main stream
    batches = [ScoreBatch(4, upper, ctx) for _ in range(depth)]
    try:
        for b, lo, hi in zip(batches, cuts[:-1], cuts[1:]):
            for G in mats[lo:hi]:
                b.add_dense(G)
            b.score(forest, exp_arr, thre, forest.w)
        return _pool_scored(batches, thre, res, min_count, chains, no_pool, 15000)
    finally:
        for b in batches:
            b.close()
