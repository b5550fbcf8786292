"""Synthetic workloads for the neoloop-caller scoring path.

Everything here is deterministic in its integer seeds: SV lists, assembly lines in
the format written by ``assembleSV.output_all`` (neoloop/assembly.py:229-275), the
procedural contact map carrying those SVs, and a synthetic random forest with the
contract the reference expects from a Peakachu model (``feature_importances_.size
== (2w+1)**2`` and ``predict_proba``; neoloop/callers.py:572-577, 622).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .coolshim import HG38_CHROMSIZES, SyntheticCooler

__all__ = [
    "sv_to_assembly_line", "demo_k562_svs", "random_simple_svs", "random_complex_assemblies",
    "make_workload", "train_synthetic_forest", "analytic_expected", "SMALL_CHROMSIZES",
]

SMALL_CHROMSIZES = {"chr1": 12_000_000, "chr2": 9_000_000, "chr3": 7_000_000, "chr4": 6_000_000}


def _strip(c: str) -> str:
    return c[3:] if c.startswith("chr") else c


def sv_to_assembly_line(sv: Sequence, span: int, chromsizes: Dict[str, int], note="translocation") -> str:
    """One ``C#`` line body: ``type,c1,p1,s1,c2,p2,s2<TAB>c1,b1<TAB>c2,b2``.

    Bounds follow ``assembleSV._change_format`` (neoloop/assembly.py:229-256): the far
    end of each retained piece ('+' keeps the left side of a breakpoint, '-' the right).
    """
    c1, p1, s1, c2, p2, s2 = sv[:6]
    b1 = max(0, p1 - span) if s1 == "+" else min(chromsizes[c1], p1 + span)
    b2 = min(chromsizes[c2], p2 + span) if s2 == "-" else max(0, p2 - span)
    tok = ",".join(map(str, (note, _strip(c1), p1, s1, _strip(c2), p2, s2)))
    return "\t".join([tok, f"{_strip(c1)},{b1}", f"{_strip(c2)},{b2}"])


def demo_k562_svs() -> List[Tuple]:
    """The five SVs of demo/K562-test-SVs.txt as (c1, p1, s1, c2, p2, s2, type),
    parsed the way ``load_translocation_fil`` does (neoloop/util.py:65-101)."""
    raw = [
        ("chr22", "chr9", "+-", 23290555, 130731760, "translocation"),
        ("chr9", "chr13", "++", 131280137, 108009063, "translocation"),
        ("chr13", "chr13", "+-", 93371155, 107848624, "deletion"),
        ("chr9", "chr13", "+-", 131280000, 93252000, "translocation"),
        ("chr22", "chr9", "++", 16819349, 131199197, "translocation"),
    ]
    out = []
    for c1, c2, st, p1, p2, note in raw:
        s1, s2 = st[0], st[1]
        if c1 == c2 and p1 > p2:
            p1, p2, s1, s2 = p2, p1, s2, s1
        out.append((c1, p1, s1, c2, p2, s2, note))
    return out


def random_simple_svs(n: int, chromsizes: Dict[str, int], seed: int, margin: int = 4_000_000,
                      min_intra: int = 500_000) -> List[Tuple]:
    """n random translocations / deletions (+-) / inversions (++ or --)."""
    rng = np.random.default_rng([seed, 0x5F])
    names = [c for c in chromsizes if chromsizes[c] > 2 * margin + min_intra and c not in ("chrY",)]
    out = []
    while len(out) < n:
        kind = rng.choice(["translocation", "deletion", "inversion"], p=[0.5, 0.25, 0.25])
        if kind == "translocation":
            c1, c2 = rng.choice(names, size=2, replace=False)
            p1 = int(rng.integers(margin, chromsizes[c1] - margin))
            p2 = int(rng.integers(margin, chromsizes[c2] - margin))
            s1, s2 = rng.choice(["+", "-"]), rng.choice(["+", "-"])
        else:
            c1 = c2 = str(rng.choice(names))
            p1 = int(rng.integers(margin, chromsizes[c1] - margin - min_intra))
            hi = min(chromsizes[c1] - margin, p1 + 20_000_000)
            p2 = int(rng.integers(p1 + min_intra, max(p1 + min_intra + 1, hi)))
            if kind == "deletion":
                s1, s2 = "+", "-"
            else:
                s1 = s2 = str(rng.choice(["+", "-"]))
        out.append((str(c1), p1, str(s1), str(c2), p2, str(s2), str(kind)))
    return out


def random_complex_assemblies(n: int, chromsizes: Dict[str, int], seed: int, span: int,
                              margin: int = 4_000_000) -> Tuple[List[str], List[Tuple]]:
    """n ``A#``-style chains of 2-5 compatible SVs (config 3).

    Chains satisfy the asserts of ``complexSV.reorganize_matrix``
    (neoloop/assembly.py:414-428): consecutive SVs share a chromosome, the traversal
    direction leaving SV k equals the one entering SV k+1, and the middle block has
    positive length in that direction.
    """
    rng = np.random.default_rng([seed, 0xC0])
    names = [c for c in chromsizes if chromsizes[c] > 3 * margin and c not in ("chrY",)]
    lines, all_svs = [], []
    while len(lines) < n:
        k = int(rng.integers(2, 6))
        svs = []
        c_cur = str(rng.choice(names))
        p_cur = int(rng.integers(margin, chromsizes[c_cur] - margin))
        s1 = str(rng.choice(["+", "-"]))
        ok = True
        for t in range(k):
            c2 = str(rng.choice(names))
            p2 = int(rng.integers(margin, chromsizes[c2] - margin))
            s2 = str(rng.choice(["+", "-"]))
            svs.append((c_cur, p_cur, s1, c2, p2, s2, "translocation" if c2 != c_cur else "inversion"))
            if t == k - 1:
                break
            # middle block: leaves p2 in direction '+' (s2 == '-') or '-' (s2 == '+')
            length = int(rng.integers(300_000, 1_500_000))
            if s2 == "-":
                nxt = p2 + length
                s1 = "+"
            else:
                nxt = p2 - length
                s1 = "-"
            if not (margin < nxt < chromsizes[c2] - margin):
                ok = False
                break
            c_cur, p_cur = c2, nxt
        if not ok:
            continue
        toks = [",".join(map(str, (sv[6], _strip(sv[0]), sv[1], sv[2], _strip(sv[3]), sv[4], sv[5]))) for sv in svs]
        f, l = svs[0], svs[-1]
        b1 = max(0, f[1] - span) if f[2] == "+" else min(chromsizes[f[0]], f[1] + span)
        b2 = min(chromsizes[l[3]], l[4] + span) if l[5] == "-" else max(0, l[4] - span)
        lines.append("\t".join(toks + [f"{_strip(f[0])},{b1}", f"{_strip(l[3])},{b2}"]))
        all_svs.extend(svs)
    return lines, all_svs


def make_workload(kind: str, res: int = 10000, seed: int = 0, n_assemblies: int = None,
                  chromsizes: Dict[str, int] = None, span: int = 3_000_000, band_bp: int = 5_000_000):
    """Return (clr, assemblies dict {ID: line body}) for a named synthetic workload.

    kind: 'demo' (cfg1: the five demo SVs), 'simple' (cfg2/5), 'complex' (cfg3).
    """
    cs = dict(chromsizes) if chromsizes is not None else dict(HG38_CHROMSIZES)
    if kind == "demo":
        svs = demo_k562_svs()
        lines = {f"C{i}": sv_to_assembly_line(sv, span, cs, note=sv[6]) for i, sv in enumerate(svs)}
    elif kind == "simple":
        svs = random_simple_svs(n_assemblies or 50, cs, seed, margin=span + 1_000_000)
        lines = {f"C{i}": sv_to_assembly_line(sv, span, cs, note=sv[6]) for i, sv in enumerate(svs)}
    elif kind == "complex":
        body, svs = random_complex_assemblies(n_assemblies or 200, cs, seed, span, margin=span + 1_000_000)
        lines = {f"A{i}": b for i, b in enumerate(body)}
    else:
        raise ValueError(kind)
    amplitude = 255.0 * (res / 10000.0)  # keeps per-pixel depth comparable across resolutions
    clr = SyntheticCooler(binsize=res, chromsizes=cs, seed=seed, amplitude=amplitude, band_bp=band_bp,
                          svs=[sv[:6] for sv in svs])
    return clr, lines


def assembly_blocks(clr, line: str, span: int):
    """[(chrom, lo_bin, hi_bin)] of the blocks ``complexSV(..., flexible=False)`` builds for one assembly
    line (neoloop/assembly.py:408-440), geometry only."""
    from .assembly import complexSV
    fu = complexSV(clr, line, span=span, col="sweight", expected_values={}, flexible=False)
    fu.reorganize_matrix(map_only=True)
    return [clr._bin_extent(tuple(r)) for r in fu.blocks]


def _assembly_pixels(clr, line, span):
    """Stored pixels (global bin1 <= bin2, count) of every block pair of one assembly."""
    blocks = assembly_blocks(clr, line, span)
    offs = {c: clr._chrom_offset(c) for c, _lo, _hi in blocks}
    keys, vals = [], []
    total = np.int64(sum(clr._nbins(c) for c in clr.chromnames))
    for j1, (c1, lo1, hi1) in enumerate(blocks):
        for (c2, lo2, hi2) in blocks[j1:]:
            B = np.asarray(clr._raw_block(c1, lo1, hi1, c2, lo2, hi2))
            rr, cc = np.nonzero(B)
            gi = rr.astype(np.int64) + (lo1 + offs[c1])
            gj = cc.astype(np.int64) + (lo2 + offs[c2])
            keys.append(np.minimum(gi, gj) * total + np.maximum(gi, gj))
            vals.append(B[rr, cc].astype(np.int32))
    return (np.concatenate(keys), np.concatenate(vals)) if keys else (np.zeros(0, np.int64), np.zeros(0, np.int32))


def _assembly_pixels_of(clr, span, line):
    return _assembly_pixels(clr, line, span)


def regional_pixel_cooler(clr, lines: Dict[str, str], span: int, nproc: int = 1):
    """A ``PixelCooler`` (cooler's pixel table, the format that lives in HBM) holding every pixel the
    assemblies in ``lines`` read from ``clr``: all block pairs of every assembly.  The rest of the
    genome stays empty, so an hg38-sized 5 kb table for ~200 assemblies is a few hundred MB instead of
    the whole map; bins, chromosomes and weights are genome-wide as in a real ``.cool``.
    ``complexSV.reorganize_matrix`` reads the same values from it as from ``clr`` itself."""
    from .coolshim import PixelCooler
    from functools import partial
    from .util import run_parallel
    parts = run_parallel(nproc, partial(_assembly_pixels_of, clr, span), list(lines.values()))
    names = list(clr.chromnames)
    nb = [clr._nbins(c) for c in names]
    total = int(sum(nb))
    key = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.int64)
    val = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int32)
    key, first = np.unique(key, return_index=True)              # sorted by (bin1, bin2); regions may overlap
    val = val[first]
    bin1 = key // total
    bin2 = (key - bin1 * total).astype(np.int32)
    indptr = np.zeros(total + 1, dtype=np.int64)
    np.cumsum(np.bincount(bin1, minlength=total), out=indptr[1:])
    weights = {col: np.concatenate([clr._weights(c, 0, n, col) for c, n in zip(names, nb)])
               for col in ("weight", "sweight")}
    sizes = {c: int(clr._chromsizes[c]) for c in names}
    return PixelCooler(clr.binsize, sizes, indptr, bin2, val, weights)


def analytic_expected(clr: SyntheticCooler, maxdis: int = 5_000_000) -> Dict[int, float]:
    """Closed-form stand-in for ``calculate_expected`` on a SyntheticCooler.

    mean balanced contact at distance d = A/(d+1) * E[w]^2 with w ~ U(0.05, 0.15);
    used where the genome-wide prologue (neoloop/util.py:430-468) would take minutes
    on a procedural hg38 map.  Same container type as the reference: {int: float}.
    """
    n = maxdis // clr.binsize
    d = np.arange(n + 1)
    vals = clr.amplitude / (d + 1.0) * (0.1 * 0.1)
    return {int(k): float(v) for k, v in zip(d, vals)}


def train_synthetic_forest(features: np.ndarray, labels: np.ndarray, n_estimators=100, max_depth=20,
                           random_state=42):
    """RandomForestClassifier with upstream Peakachu's shape."""
    from sklearn.ensemble import RandomForestClassifier
    rf = RandomForestClassifier(n_estimators=n_estimators, max_depth=max_depth,
                                random_state=random_state, n_jobs=1)
    rf.fit(np.asarray(features, dtype=np.float64), np.asarray(labels, dtype=np.int64))
    return rf
