"""Numeric helpers of the neoloop-caller path (mirror of the four hot functions of
neoloop/util.py; the SV/peak text parsers of that module are out of scope).

``calculate_expected`` is the per-sample prologue (neoloop/util.py:403-468), row f-1 of SURVEY.md
section 8: the masked per-diagonal sums of every chromosome run on the GPU in numpy's summation
order (``nlf_expected_diagonals``); the Python-float accumulation over chromosomes and the
isotonic fit are the reference's own host steps.
"""
from __future__ import annotations

from typing import Dict, Iterable

import numpy as np

__all__ = ["find_chrom_pre", "calculate_expected", "autosomes", "image_normalize", "distance_normalize"]


def find_chrom_pre(chromlabels) -> str:
    """'chr' when the map's chromosome labels carry the prefix (neoloop/util.py:10-17)."""
    return "chr" if chromlabels[0].startswith("chr") else ""


def run_parallel(n_jobs, fn, items):
    """[fn(x) for x in items] on ``n_jobs`` loky worker processes whose numeric libraries are held to ONE thread each.
    The host work farmed out on this path (Huber / isotonic fits on <= 400 points, procedural pixel blocks) is far
    too small for threaded BLAS; left alone, every worker starts cpu_count // n_jobs spinning OpenBLAS / OpenMP
    threads, and with one such pool per GPU rank a box is oversubscribed many times over (8 ranks x 4 workers x 8
    threads on 32 cores made a 2 s phase take minutes)."""
    items = list(items)
    if n_jobs <= 1 or len(items) <= 1:
        return [fn(x) for x in items]
    from joblib import Parallel, delayed, parallel_config
    with parallel_config(backend="loky", inner_max_num_threads=1):
        return Parallel(n_jobs=n_jobs)(delayed(fn)(x) for x in items)


def shutdown_workers():
    """Stop this process's loky worker pool now.  A per-GPU worker process (``cli.call_resolution`` and friends spawn
    one per device) must do this before it returns: multiprocessing joins a child's non-daemonic children at exit, and
    idle loky workers only leave after their 300 s timeout -- the parent would sit in ``shutdown()`` for five minutes."""
    try:
        from joblib.externals.loky import reusable_executor
        ex = getattr(reusable_executor, "_executor", None)        # None: this process never started a pool
        if ex is not None:
            ex.shutdown(wait=True, kill_workers=True)
    except Exception:      # noqa: BLE001 - nothing to stop
        pass


def autosomes(chromnames: Iterable[str]):
    """Chromosome filter of scripts/neoloop-caller:148-152 (drops names with _ M X Y MT EBV)."""
    keep = []
    for c in chromnames:
        if any(tag in c for tag in ("_", "M", "X", "Y", "MT", "EBV")):
            continue
        keep.append(c)
    return keep


def chrom_band(clr, chrom, nd, balance):
    """(band, weights): band[r, i] = balanced contact of bins (r, r+i) of one chromosome, i < nd.
    Sources that can produce a band directly (``fetch_band``) are used as such; any other
    cooler-like object goes through its sparse matrix (O(nnz) scatter, no per-diagonal scans)."""
    col = "weight" if isinstance(balance, bool) else balance
    weights = clr.bins().fetch(chrom)[col].values
    n = weights.size
    nd = min(nd, n)
    if hasattr(clr, "fetch_band"):
        band = clr.fetch_band(chrom, nd, balance)
    else:
        coo = clr.matrix(balance=balance, sparse=True).fetch(chrom).tocoo()
        d = coo.col - coo.row
        ok = (d >= 0) & (d < nd)
        band = np.zeros((n, nd), dtype=np.float64)
        band[coo.row[ok], d[ok]] = coo.data[ok]
    return np.ascontiguousarray(band, dtype=np.float64), weights


def _chrom_diagonal_sums(args):
    """One chromosome: {distance: [sum, count]} over bin pairs whose weights are both finite
    and positive (neoloop/util.py:403-428).  The masked diagonal sums run on the GPU
    (nlf_expected_diagonals) in numpy's summation order."""
    import ctypes as C
    from . import _lib
    clr, chrom, maxdis, balance = args
    if hasattr(clr, "device_pixels"):
        # resident pixel table: band, valid mask and sums stay on the device
        col = "weight" if isinstance(balance, bool) else balance
        px = clr.device_pixels(None, col)
        n = clr._nbins(chrom)
        lo, hi = clr.global_bins(chrom, 0, n)
        sums, cnts = px.expected_diagonals(lo, hi, maxdis + 1, balance=bool(balance))
        return {i: [np.float64(sums[i]), int(cnts[i])] for i in range(min(n - 1, maxdis) + 1) if cnts[i] > 0}
    band, weights = chrom_band(clr, chrom, maxdis + 1, balance)
    n, nd = band.shape
    band = np.where(np.isfinite(band), band, 0.0)         # masked out below; keep the copy finite
    valid = np.ascontiguousarray(np.isfinite(weights) & (weights > 0), dtype=np.uint8)
    sums = np.empty(nd, dtype=np.float64)
    cnts = np.empty(nd, dtype=np.int64)
    ctx = _lib.Context.get()
    _lib.check(_lib.load().nlf_expected_diagonals(ctx.h, _lib.ptr(band, C.c_double), n, nd, _lib.ptr(valid, C.c_ubyte),
                                                  _lib.ptr(sums, C.c_double), _lib.ptr(cnts, C.c_longlong)))
    out = {}
    for i in range(min(n - 1, maxdis) + 1):
        if cnts[i] > 0:
            out[i] = [np.float64(sums[i]), int(cnts[i])]
    return out


def calculate_expected(clr, chroms, maxdis=5000000, balance=True, nproc=1) -> Dict[int, float]:
    """Genome-wide mean contact per genomic distance with a decreasing isotonic fit
    (neoloop/util.py:430-468).  Returns {int distance in bins: float}.  ``nproc`` is accepted for
    interface compatibility; the per-chromosome scans run on the GPU one after the other."""
    from sklearn.isotonic import IsotonicRegression
    maxdis = maxdis // clr.binsize
    per_chrom = [_chrom_diagonal_sums((clr, c, maxdis, balance)) for c in chroms]
    raw = {}
    for i in range(maxdis + 1):
        nume = 0
        denom = 0
        for table in per_chrom:
            if i in table:
                nume += table[i][0]
                denom += table[i][1]
        if denom > 10 and nume > 0:
            raw[i] = nume / denom
    xs = np.r_[sorted(raw)]
    ys = np.r_[[raw[i] for i in xs]]
    fit = IsotonicRegression(increasing=False, out_of_bounds="clip").fit(xs, ys)
    d = np.arange(maxdis + 1)
    return dict(zip(d, fit.predict(d)))


def load_translocation_fil(fil_path, res, minIntra):
    """SV list of ``assemble-complexSVs`` (neoloop/util.py:65-101): lines ``chrA chrB ++|+-|-+|-- posA posB type``;
    a position may be an interval ``s-e`` no wider than one bin (its midpoint is used).  Intra-chromosomal pairs
    are ordered by position, dropped below ``minIntra`` (duplications, '-+', below twice that).  Returns the
    sorted unique records (c1, p1, s1, c2, p2, s2, type) with the ``chr`` prefix stripped."""
    def position(field):
        if '-' not in field:
            return int(field)
        s, e = (int(v) for v in field.split('-'))
        return None if e - s > res else (s + e) // 2

    records = set()
    with open(fil_path) as source:
        for line in source:
            c1, c2, strands, f1, f2, note = line.rstrip().split()
            p1, p2 = position(f1), position(f2)
            if p1 is None or p2 is None:
                continue
            s1, s2 = strands[0], strands[1]
            if c1 == c2:
                if abs(p2 - p1) < minIntra:
                    continue
                if p1 > p2:
                    p1, p2, s1, s2 = p2, p1, s2, s1
                if s1 == '-' and s2 == '+' and abs(p2 - p1) < minIntra * 2:
                    continue
            records.add((c1.lstrip('chr'), p1, s1, c2.lstrip('chr'), p2, s2, note))
    return sorted(records)


def map_coordinates(k_p, res, chrom_1, chrom_2, unit='bp'):
    """(forward, reverse) maps between (chrom, bin start) and the bin index on the rearranged axis of a
    Fusion matrix described by ``k_p`` (neoloop/util.py:335-389): a side given as (high, low) is read backwards."""
    def side(a, b, first):
        lo, hi = (a, b) if a < b else (b, a)
        lo_bin, hi_bin = lo // res, -(-hi // res)
        nbin = hi_bin - lo_bin
        keys = range(lo, hi, res) if unit == 'bp' else range(lo_bin, hi_bin)
        idx = range(first, first + nbin) if a < b else range(first + nbin - 1, first - 1, -1)
        return dict(zip(keys, idx)), nbin

    d1, n1 = side(k_p[0], k_p[1], 0)
    d2, _n2 = side(k_p[2], k_p[3], n1)
    forward = {(chrom_1, k): v for k, v in d1.items()}
    forward.update({(chrom_2, k): v for k, v in d2.items()})
    reverse = {v: k for k, v in forward.items()}
    return forward, reverse


def image_normalize(arr_2d):
    """(a - min) / (max - min) (neoloop/util.py:493-497).  On the scoring path this runs inside
    the CUDA feature kernel; this helper exists for API compatibility on small host arrays."""
    arr_2d = np.asarray(arr_2d)
    return (arr_2d - arr_2d.min()) / (arr_2d.max() - arr_2d.min())


def distance_normalize(arr_pool, exp_bychrom, xi, yi, w):
    """Gates and observed/expected transform of a stack of windows (neoloop/util.py:499-524): NaN -> 0 (in place,
    as the reference does), windows with fewer than 10 % non-zero cells, a non-positive lower-left mean or a
    centre below 0.1 of it are dropped, the others are divided by the expected at each cell's distance unless the
    window reaches beyond the expected table.  Returns (list of windows, list of (x, y)).  On the scoring path this
    is fused into the feature kernel; this entry serves callers that hold their own window stacks."""
    import ctypes as C
    from . import _lib
    xi = np.ascontiguousarray(xi, dtype=np.int32)
    yi = np.ascontiguousarray(yi, dtype=np.int32)
    n = int(xi.size)
    S = 2 * int(w) + 1
    if n == 0:
        return [], []
    if isinstance(arr_pool, np.ndarray) and arr_pool.dtype == np.float64:
        arr_pool[np.isnan(arr_pool)] = 0                      # the reference clears NaN inside the caller's stack
    stack = np.ascontiguousarray(np.asarray(arr_pool, dtype=np.float64)[:n]).reshape(n, S, S)
    exp = np.ascontiguousarray(exp_bychrom, dtype=np.float64)
    out = np.empty_like(stack)
    keep = np.zeros(n, dtype=np.uint8)
    ctx = _lib.Context.get()
    _lib.check(_lib.load().nlf_distance_normalize(ctx.h, _lib.ptr(stack, C.c_double), n, _lib.ptr(xi, C.c_int),
                                                  _lib.ptr(yi, C.c_int), int(w), _lib.ptr(exp, C.c_double), int(exp.size),
                                                  _lib.ptr(out, C.c_double), _lib.ptr(keep, C.c_ubyte)))
    sel = np.flatnonzero(keep)
    return [out[k] for k in sel], [(xi[k], yi[k]) for k in sel]
