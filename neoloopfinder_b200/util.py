"""Numeric helpers of the neoloop-caller path (mirror of the four hot functions of
neoloop/util.py; the SV/peak text parsers of that module are out of scope).

``calculate_expected`` is the per-sample prologue (neoloop/util.py:403-468).  It is row f-1 of
SURVEY.md section 8 ("next"): it stays host-side numpy here, restated so that every number is
produced by the same numpy/scipy/sklearn calls as the reference (same ``diagonal(i)[valid].sum()``
pairwise sums, same Python-float accumulation over chromosomes, same isotonic fit).
"""
from __future__ import annotations

from typing import Dict, Iterable

import numpy as np

__all__ = ["find_chrom_pre", "calculate_expected", "autosomes", "image_normalize", "distance_normalize"]


def find_chrom_pre(chromlabels) -> str:
    """'chr' when the map's chromosome labels carry the prefix (neoloop/util.py:10-17)."""
    return "chr" if chromlabels[0].startswith("chr") else ""


def autosomes(chromnames: Iterable[str]):
    """Chromosome filter of scripts/neoloop-caller:148-152 (drops names with _ M X Y MT EBV)."""
    keep = []
    for c in chromnames:
        if any(tag in c for tag in ("_", "M", "X", "Y", "MT", "EBV")):
            continue
        keep.append(c)
    return keep


def _chrom_diagonal_sums(args):
    """One chromosome: {distance: [sum, count]} over bin pairs whose weights are both finite
    and positive (neoloop/util.py:403-428)."""
    clr, chrom, maxdis, balance = args
    hic = clr.matrix(balance=balance, sparse=True).fetch(chrom)
    col = "weight" if isinstance(balance, bool) else balance
    weights = clr.bins().fetch(chrom)[col].values
    ok = np.isfinite(weights) & (weights > 0)
    n = hic.shape[0]
    out = {}
    for i in range(min(n - 1, maxdis) + 1):
        valid = ok if i == 0 else ok[:-i] * ok[i:]
        cur = hic.diagonal(i)[valid]
        if cur.size > 0:
            out[i] = [cur.sum(), cur.size]
    return out


def calculate_expected(clr, chroms, maxdis=5000000, balance=True, nproc=1) -> Dict[int, float]:
    """Genome-wide mean contact per genomic distance with a decreasing isotonic fit
    (neoloop/util.py:430-468).  Returns {int distance in bins: float}."""
    from sklearn.isotonic import IsotonicRegression
    maxdis = maxdis // clr.binsize
    jobs = [(clr, c, maxdis, balance) for c in chroms]
    if nproc == 1:
        per_chrom = [_chrom_diagonal_sums(j) for j in jobs]
    else:
        from multiprocess import Pool
        with Pool(nproc) as pool:
            per_chrom = pool.map(_chrom_diagonal_sums, jobs)
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


def image_normalize(arr_2d):
    """(a - min) / (max - min) (neoloop/util.py:493-497).  On the scoring path this runs inside
    the CUDA feature kernel; this helper exists for API compatibility on small host arrays."""
    arr_2d = np.asarray(arr_2d)
    return (arr_2d - arr_2d.min()) / (arr_2d.max() - arr_2d.min())


def distance_normalize(*_args, **_kwargs):
    raise NotImplementedError(
        "distance_normalize operates on host window stacks in the reference (neoloop/util.py:499-524); "
        "in neoloopfinder_b200 the gates and the O/E transform are fused into the CUDA feature kernel: "
        "call Peakachu.getwindow(coords, w) instead")
