"""Genome-wide loop calling: every chromosome as a trivial assembly (BASELINE.json configs[3]).

The reference has no such entry -- ``complexSV.parse_input`` needs at least one SV (neoloop/assembly.py:323-334)
and ``Peakachu.__init__`` takes a dense matrix (neoloop/callers.py:527-535), 19.8 GB for chr1 at 5 kb.  This
module extends the same scoring path to whole chromosomes: a chromosome is what ``Peakachu`` would see for an
assembly made of one block and no junction -- its balanced upper-triangular matrix with the gap columns removed
(``Fusion.remove_gaps``, callers.py:504-522) -- given to the kernels as the band they read anyway
(``upper + 14`` diagonals).  No block rescaling applies (one block, slope 1).  Results equal
``Peakachu(dense gap-free chromosome matrix).predict(...)`` (tests/test_gpu_genome.py).

The sample's pixel table (cooler's layout) is uploaded once per GPU; gap removal and the bands are built on the
device.  Chromosomes are dealt to the GPUs by longest-processing-time on their bin counts, one worker process
per GPU, KB-sized result lists merged on the host: no collective.  A job holds up to 2^(32 - bits(upper)) rows
(8.3 M at upper = 400), so a chromosome is never cut into row shards.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

__all__ = ["call_genome", "call_chromosomes", "genome_lines", "as_pixel_source", "dense_chromosome"]


class _Subset:
    """View of a cooler-like source restricted to some chromosomes (same pixels, same bin table)."""

    def __init__(self, clr, names):
        self._clr = clr
        self.chromnames = list(names)
        self.binsize = clr.binsize
        self.chromsizes = {c: int(clr.chromsizes[c]) for c in names}

    def __getattr__(self, name):
        return getattr(self._clr, name)


def as_pixel_source(clr, chroms: Sequence[str], band_bins: int):
    """``clr`` itself when it is a pixel table that can live in HBM, else a banded, intra-chromosomal pixel table
    of the wanted chromosomes materialised from it (``PixelCooler.from_cooler``)."""
    if hasattr(clr, "device_pixels"):
        return clr
    from .coolshim import PixelCooler
    return PixelCooler.from_cooler(_Subset(clr, chroms), band_bins=band_bins, trans=False)


def dense_chromosome(clr, chrom: str, balance="sweight"):
    """(gap-free dense matrix, kept bins) of one chromosome as the reference's classes would build them: balanced
    matrix, NaN -> 0, np.triu, gap columns removed.  For parity tests on small chromosomes."""
    M = np.array(clr.matrix(balance=balance).fetch(chrom), dtype=np.float64)
    M[np.isnan(M)] = 0
    M = np.triu(M)
    keep = np.where(~(M.sum(axis=0) == 0))[0]
    if 11 > keep.size:
        keep = np.arange(M.shape[0])
    return np.ascontiguousarray(M[np.ix_(keep, keep)]), keep


def call_chromosomes(clr, chroms: Sequence[str], model, expected, prob=0.9, min_count=2, no_pool=False,
                     balance="sweight", ctx=None, stats=None, upper=400):
    """Score the listed chromosomes of a pixel-table source on ONE GPU.
    Returns {chrom: (start1[], start2[], prob[])} with bin starts in bp, ordered by (start1, start2)."""
    from . import _lib
    from .forest import DeviceForest, load_forest
    from .scoring import ScoreBatch, score_and_pool
    ctx = ctx or _lib.Context.get()
    forest = model if isinstance(model, DeviceForest) else load_forest(model, ctx)
    res = clr.binsize
    exp_arr = np.r_[[expected[i] for i in sorted(expected)]]
    col = "weight" if balance in (True, False, None) else balance
    px = clr.device_pixels(ctx, col)
    batch = ScoreBatch(4, upper, ctx)
    kept = []
    try:
        for c in chroms:
            lo, hi = clr.global_bins(c, 0, clr._nbins(c))
            _job, k = batch.add_pixels(px, lo, hi, balance=bool(balance))
            kept.append(k)
        loops, (n_cand, n_scored, n_don) = score_and_pool(batch, forest, exp_arr, prob, res, min_count, None, no_pool)
    finally:
        batch.close()
    if stats is not None:
        stats["candidates"] = stats.get("candidates", 0) + int(n_cand.sum())
        stats["scored"] = stats.get("scored", 0) + int(n_scored.sum())
        stats["donuts"] = stats.get("donuts", 0) + int(n_don.sum())
        stats["bins"] = stats.get("bins", 0) + int(sum(len(k) for k in kept))
    out = {}
    for c, k, (li, lj, lp) in zip(chroms, kept, loops):
        a = k[li].astype(np.int64) * res
        b = k[lj].astype(np.int64) * res
        order = np.lexsort((b, a))
        out[c] = (a[order], b[order], np.asarray(lp)[order])
    return out


def _genome_worker(dev, clr, chroms, model, expected, prob, min_count, no_pool, balance, upper, child=False):
    from . import _lib
    try:
        ctx = _lib.Context.get(dev)
        st = {}
        r = call_chromosomes(clr, chroms, model, expected, prob, min_count, no_pool, balance, ctx, st, upper)
        return r, st
    finally:
        if child:                      # a spawned worker stops its own loky pool before returning (util.shutdown_workers)
            from .util import shutdown_workers
            shutdown_workers()


def call_genome(clr, model, expected, prob=0.9, min_count=2, no_pool=False, chroms: Optional[Iterable[str]] = None,
                balance="sweight", devices: Optional[List[int]] = None, stats=None, workers: Optional[str] = None,
                upper=400):
    """All (or the listed) chromosomes of ``clr`` as trivial assemblies, sharded over the visible GPUs.
    Returns {chrom: (start1[], start2[], prob[])}."""
    from . import _lib
    from .forest import DeviceForest
    from .sharding import lpt_shards
    names = list(chroms) if chroms is not None else list(clr.chromnames)
    src = as_pixel_source(clr, names, upper + 14 + 36)
    if devices is None:
        devices = list(range(_lib.device_count()))
    devices = list(dict.fromkeys(int(d) for d in devices))
    if not devices:
        raise RuntimeError("no CUDA device visible: neoloopfinder_b200 has no CPU fallback")
    weights = [float(src._nbins(c)) for c in names]
    shards = lpt_shards(weights, len(devices))
    jobs = [(d, [names[k] for k in s]) for d, s in zip(devices, shards) if s]
    mode = workers or os.environ.get("NLF_WORKERS", "process")
    if isinstance(model, DeviceForest):
        mode = "thread"
    if len(jobs) <= 1:
        parts = [_genome_worker(d, src, cs, model, expected, prob, min_count, no_pool, balance, upper) for d, cs in jobs]
    elif mode == "thread":
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            parts = list(ex.map(lambda j: _genome_worker(j[0], src, j[1], model, expected, prob, min_count, no_pool,
                                                         balance, upper), jobs))
    else:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=len(jobs), mp_context=mp.get_context("spawn")) as ex:
            futs = [ex.submit(_genome_worker, d, src, cs, model, expected, prob, min_count, no_pool, balance, upper, True)
                    for d, cs in jobs]
            parts = [f.result() for f in futs]
    out = {}
    for r, st in parts:
        out.update(r)
        if stats is not None:
            for k, v in st.items():
                stats[k] = stats.get(k, 0) + v
    return {c: out[c] for c in names}


def genome_lines(calls: Dict[str, tuple], res: int) -> Dict[tuple, list]:
    """Calls of ``call_genome`` in the shape ``merge_results`` produces for assemblies
    ({(c1, s1, e1, c2, s2, e2): [(label, distance, neo flag)]}, scripts/neoloop-caller:161-170): the label is the
    chromosome, the neo-loop flag is 0 (no junction is crossed)."""
    out = {}
    for c, (a, b, _p) in calls.items():
        for s1, s2 in zip(a.tolist(), b.tolist()):
            out[(c, s1, s1 + res, c, s2, s2 + res)] = [(c, s2 - s1, 0)]
    return out
