"""Cooler-compatible contact-map sources for the neoloop-caller path.

The reference reads Hi-C maps through the third-party ``cooler`` package (HDF5).
Neither ``cooler`` nor ``h5py``/HDF5 exist in this image or on the GPU box
(SURVEY.md section 0, D5/D6), so the path is fed by duck-typed objects that
expose exactly the attributes the reference touches on ``clr``:

* ``clr.binsize``, ``clr.chromnames``, ``clr.chromsizes[c]`` (+ ``.sum()``),
  ``clr.info``                                    (scripts/neoloop-caller:124-128, 248-257)
* ``clr.matrix(balance=B).fetch(r1[, r2])``       (neoloop/assembly.py:399-406)
* ``clr.matrix(balance=B, sparse=True).fetch(c)`` (neoloop/util.py:408)
* ``clr.bins().fetch(r)[col].values`` (writable) and ``['chrom'].size``
                                                  (neoloop/assembly.py:387-397, 440; util.py:411-413)
* ``clr.pixels()[:]['count'].values.sum()``       (scripts/neoloop-caller:251-252)

Three sources are provided:

``SyntheticCooler``  a *procedural* map: every pixel count is a deterministic
    hash of (seed, bin1, bin2), so an hg38-sized "file" costs no storage, any
    region can be fetched in any order, and the reference (oracle) and the
    CUDA path see bit-identical inputs.  The on-disk form is a small JSON spec.
``ArrayCooler``      explicit per-chromosome-pair dense blocks held in memory /
    ``.npz`` (for hand-made test matrices).

``PixelCooler``      cooler's own storage layout -- the genome-wide upper-triangular pixel table
    (``bin1_offset`` index, ``bin2_id``, ``count``) plus the bin table -- held in memory /
    ``.npz``.  This is the on-disk format that replaces ``.cool`` here and the only source that can
    live in HBM: ``device_pixels()`` uploads it once and assemblies, chromosome bands and the
    expected prologue are then built on the GPU.

All are picklable (the reference fans out with loky processes).
"""
from __future__ import annotations

import json
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["SyntheticCooler", "ArrayCooler", "PixelCooler", "load_cooler", "HG38_CHROMSIZES", "parse_region"]

# hg38 primary assembly; autosomes + X sum to 3 031 042 417, the constant at
# scripts/neoloop-caller:257.
HG38_CHROMSIZES: Dict[str, int] = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555,
    "chr5": 181538259, "chr6": 170805979, "chr7": 159345973, "chr8": 145138636,
    "chr9": 138394717, "chr10": 133797422, "chr11": 135086622, "chr12": 133275309,
    "chr13": 114364328, "chr14": 107043718, "chr15": 101991189, "chr16": 90338345,
    "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}


class _Series:
    """Minimal stand-in for the pandas Series the reference indexes.

    ``.values`` returns a fresh *writable* array every time: pandas 3.0 hands
    out read-only views and ``complexSV.get_bias`` writes into the result
    (neoloop/assembly.py:394-395).
    """

    def __init__(self, arr, index=None):
        self._arr = np.asarray(arr)
        self._index = None if index is None else list(index)

    @property
    def values(self):
        return np.array(self._arr, copy=True)

    @property
    def size(self):
        return int(self._arr.size)

    def sum(self):
        return self._arr.sum()

    def __len__(self):
        return int(self._arr.size)

    def __getitem__(self, key):
        if self._index is not None and not isinstance(key, (int, np.integer, slice)):
            return self._arr[self._index.index(key)]
        return self._arr[key]

    def __iter__(self):
        return iter(self._arr)


class _Frame:
    def __init__(self, cols: Dict[str, np.ndarray]):
        self._cols = cols

    def __getitem__(self, col):
        return _Series(self._cols[col])

    def __len__(self):
        return len(next(iter(self._cols.values())))

    @property
    def columns(self):
        return list(self._cols)


def parse_region(region, chromsizes: Dict[str, int]) -> Tuple[str, int, int]:
    """cooler region semantics: 'chr', 'chr:a-b', or (chrom[, start, end])."""
    if isinstance(region, str):
        if ":" in region:
            c, rest = region.split(":")
            a, b = rest.replace(",", "").split("-")
            return c, int(a), int(b)
        return region, 0, int(chromsizes[region])
    region = tuple(region)
    c = region[0]
    if len(region) == 1:
        return c, 0, int(chromsizes[c])
    start = 0 if region[1] is None else int(region[1])
    end = int(chromsizes[c]) if region[2] is None else int(region[2])
    return c, start, end


class _BinsSelector:
    def __init__(self, clr):
        self._clr = clr

    def fetch(self, region):
        c, lo, hi = self._clr._bin_extent(region)
        n = hi - lo
        res = self._clr.binsize
        start = (np.arange(lo, hi, dtype=np.int64)) * res
        end = np.minimum(start + res, self._clr._chromsizes[c])
        cols = {
            "chrom": np.array([c] * n, dtype=object),
            "start": start,
            "end": end,
        }
        for col in self._clr.weight_columns():          # a column the map does not have is a KeyError, as in cooler
            cols[col] = self._clr._weights(c, lo, hi, col)
        return _Frame(cols)

    def __getitem__(self, key):
        # whole bin table (only used with [:])
        cols = {k: [] for k in ("chrom", "start", "end") + tuple(self._clr.weight_columns())}
        for c in self._clr.chromnames:
            fr = self.fetch(c)
            for k in cols:
                cols[k].append(fr._cols[k])
        return _Frame({k: np.concatenate(v) for k, v in cols.items()})


class _MatrixSelector:
    def __init__(self, clr, balance, sparse):
        self._clr = clr
        self._balance = balance
        self._sparse = sparse

    def fetch(self, r1, r2=None):
        if r2 is None:
            r2 = r1
        c1, lo1, hi1 = self._clr._bin_extent(r1)
        c2, lo2, hi2 = self._clr._bin_extent(r2)
        M = self._clr._raw_block(c1, lo1, hi1, c2, lo2, hi2).astype(np.float64)
        bal = self._balance
        if bal is True:
            bal = "weight"
        if bal:
            w1 = self._clr._weights(c1, lo1, hi1, bal)
            w2 = self._clr._weights(c2, lo2, hi2, bal)
            # cooler multiplies count * w_i * w_j (NaN weights give NaN pixels)
            M = M * w1[:, None] * w2[None, :]
        if self._sparse:
            from scipy import sparse as sp
            if bal:
                # cooler's sparse fetch only carries stored (non-zero count) pixels;
                # zero-count pixels stay structural zeros even when a weight is NaN
                raw = self._clr._raw_block(c1, lo1, hi1, c2, lo2, hi2)
                M = np.where(raw != 0, M, 0.0)
            return sp.coo_matrix(M)
        return M


class _PixelSelector:
    def __init__(self, clr):
        self._clr = clr

    def __getitem__(self, key):
        total = []
        names = self._clr.chromnames
        for i, c1 in enumerate(names):
            for c2 in names[i:]:
                n1 = self._clr._nbins(c1)
                n2 = self._clr._nbins(c2)
                B = self._clr._raw_block(c1, 0, n1, c2, 0, n2)
                if c1 == c2:
                    B = np.triu(B)
                total.append(B[B != 0].ravel())
        return _Frame({"count": np.concatenate(total) if total else np.zeros(0)})


class _CoolerBase:
    binsize: int

    def __init__(self, binsize: int, chromsizes: Dict[str, int]):
        self.binsize = int(binsize)
        self._chromsizes = {str(k): int(v) for k, v in chromsizes.items()}
        self.chromnames = list(self._chromsizes)

    # ---- cooler-like surface -------------------------------------------------
    @property
    def chromsizes(self):
        return _Series(np.array([self._chromsizes[c] for c in self.chromnames], dtype=np.int64),
                       index=self.chromnames)

    def matrix(self, balance=True, sparse=False, **_):
        return _MatrixSelector(self, balance, sparse)

    def bins(self):
        return _BinsSelector(self)

    def pixels(self):
        return _PixelSelector(self)

    # ---- helpers -------------------------------------------------------------
    def _nbins(self, c) -> int:
        return int(math.ceil(self._chromsizes[c] / self.binsize))

    def weight_columns(self) -> Tuple[str, ...]:
        return ("weight", "sweight")

    def _bin_extent(self, region) -> Tuple[str, int, int]:
        c, start, end = parse_region(region, self._chromsizes)
        # cooler.util.parse_region: malformed bounds are errors, not empty selections
        if c not in self._chromsizes:
            raise ValueError("Unknown sequence label: {}".format(c))
        if end < start:
            raise ValueError("End cannot be less than start")
        if start < 0 or end > self._chromsizes[c]:
            raise ValueError("Genomic region out of bounds: [{}, {})".format(start, end))
        res = self.binsize
        lo = start // res
        hi = int(math.ceil(end / res))
        hi = min(hi, self._nbins(c))
        lo = max(0, min(lo, hi))
        return c, lo, hi

    def _chrom_offset(self, c) -> int:
        off = 0
        for name in self.chromnames:
            if name == c:
                return off
            off += self._nbins(name)
        raise KeyError(c)


# ----------------------------------------------------------------------------------
# procedural synthetic map
# ----------------------------------------------------------------------------------
_U64 = np.uint64


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wraps modulo 2**64)."""
    with np.errstate(over="ignore"):
        x = (x + _U64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> _U64(27))) * _U64(0x94D049BB133111EB)
        x = x ^ (x >> _U64(31))
    return x


def _uniform01(h: np.ndarray) -> np.ndarray:
    """uint64 hash -> float64 in (0, 1)."""
    return ((h >> _U64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _poisson_from_hash(lam: np.ndarray, h: np.ndarray) -> np.ndarray:
    """Deterministic Poisson-like integer draw with mean ``lam`` keyed by hash ``h``.

    Inverse CDF for lam < 30, rounded normal approximation above. (Synthetic data:
    only determinism and a realistic count distribution matter.)
    """
    lam = np.asarray(lam, dtype=np.float64)
    out = np.zeros(lam.shape, dtype=np.int64)
    u = _uniform01(h)
    small = (lam > 0) & (lam < 30.0)
    if small.any():
        ls = lam[small]
        us = u[small]
        p = np.exp(-ls)
        cdf = p.copy()
        k = np.zeros(ls.shape, dtype=np.int64)
        active = us > cdf
        i = 0
        while active.any() and i < 200:
            i += 1
            p = p * ls / i
            cdf = cdf + p
            k[active] += 1
            active = active & (us > cdf)
        out[small] = k
    big = lam >= 30.0
    if big.any():
        lb = lam[big]
        u1 = u[big]
        u2 = _uniform01(_mix64(h[big] ^ _U64(0xD1B54A32D192ED03)))
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        out[big] = np.maximum(0, np.rint(lb + np.sqrt(lb) * z)).astype(np.int64)
    return out


class SyntheticCooler(_CoolerBase):
    """Procedural Hi-C map (SURVEY.md section 8d "Synthetic inputs").

    count(i, j) ~ Poisson(lambda(i, j)), lambda = A / (|i-j| + 1) inside a band of
    ``band_bp`` on the same chromosome, plus 3x3 loop bumps, plus -- for every SV in
    ``svs`` -- contacts across the junction that follow the same decay in
    *rearranged* coordinates scaled by the allele fraction.  Weights are
    U(0.05, 0.15) with a fraction ``gap_frac`` of NaN bins.

    Parameters
    ----------
    svs : iterable of (c1, p1, s1, c2, p2, s2[, fraction])
        Breakpoints with strands as in the assembly file
        (neoloop/assembly.py:323-334); chromosome names as in ``chromsizes``.
    """

    def __init__(self, binsize=10000, chromsizes=None, seed=0, amplitude=255.0,
                 band_bp=5_000_000, loop_every=40, loop_strength=60.0, gap_frac=0.02,
                 svs: Iterable[Sequence] = (), sv_fraction=0.5, info=None, gap_runs=True):
        super().__init__(binsize, chromsizes if chromsizes is not None else HG38_CHROMSIZES)
        self.seed = int(seed)
        self.amplitude = float(amplitude)
        self.band_bins = int(band_bp // self.binsize)
        self.loop_every = int(loop_every)
        self.loop_strength = float(loop_strength)
        self.gap_frac = float(gap_frac)
        self.gap_runs = bool(gap_runs)
        self.sv_fraction = float(sv_fraction)
        self.svs = [tuple(sv) for sv in svs]
        self._offsets = {}
        off = 0
        for c in self.chromnames:
            self._offsets[c] = off
            off += self._nbins(c)
        self._total_bins = off
        self._loops = {}
        self._info = dict(info) if info else None

    # -- description / persistence ------------------------------------------------
    def spec(self) -> dict:
        return {
            "format": "nlf-synthetic-cooler-v1",
            "binsize": self.binsize, "chromsizes": self._chromsizes, "seed": self.seed,
            "amplitude": self.amplitude, "band_bp": self.band_bins * self.binsize,
            "loop_every": self.loop_every, "loop_strength": self.loop_strength,
            "gap_frac": self.gap_frac, "gap_runs": self.gap_runs,
            "sv_fraction": self.sv_fraction,
            "svs": [list(s) for s in self.svs],
        }

    def save(self, path):
        with open(path, "w") as fh:
            json.dump(self.spec(), fh)

    @classmethod
    def from_spec(cls, spec: dict):
        return cls(binsize=spec["binsize"], chromsizes=spec["chromsizes"], seed=spec["seed"],
                   amplitude=spec["amplitude"], band_bp=spec["band_bp"],
                   loop_every=spec["loop_every"], loop_strength=spec["loop_strength"],
                   gap_frac=spec["gap_frac"], svs=spec.get("svs", ()),
                   sv_fraction=spec.get("sv_fraction", 0.5), gap_runs=spec.get("gap_runs", True))

    @property
    def info(self):
        if self._info is not None:
            return self._info
        # analytic estimate of the intra-chromosomal sum: sum_bins A * H(band)
        h = float(np.sum(1.0 / (np.arange(self.band_bins + 1) + 1.0)))
        return {"sum": int(self._total_bins * self.amplitude * h), "nnz": int(self._total_bins * 50)}

    # -- ingredients -----------------------------------------------------------------
    def _weights(self, c, lo, hi, col) -> np.ndarray:
        g = np.arange(lo, hi, dtype=np.uint64) + _U64(self._offsets[c])
        salt = _U64(0x5851F42D4C957F2D if col == "weight" else 0x14057B7EF767814F)
        h = _mix64(g * _U64(0x2545F4914F6CDD1D) ^ _U64(self.seed) ^ salt)
        w = 0.05 + 0.10 * _uniform01(h)
        # gaps are shared by both weight columns (unmappable bins)
        if self.gap_runs:
            # gaps keyed on groups of 1 bin; a bin is a gap with probability gap_frac
            hg = _mix64(g * _U64(0x9E3779B97F4A7C15) ^ _U64((self.seed * 7919 + 13) & 0xFFFFFFFFFFFFFFFF))
        else:
            hg = _mix64(g ^ _U64((self.seed * 7919 + 13) & 0xFFFFFFFFFFFFFFFF))
        gap = _uniform01(hg) < self.gap_frac
        w = w.astype(np.float64)
        w[gap] = np.nan
        return w

    def _chrom_loops(self, c) -> Tuple[np.ndarray, np.ndarray]:
        """Loop anchor pairs (a, b) in chromosome bin coordinates, sorted by a."""
        if c not in self._loops:
            n = self._nbins(c)
            m = max(0, n // max(1, self.loop_every))
            rng = np.random.default_rng([self.seed, self._offsets[c], 0xA11CE])
            a = np.sort(rng.integers(0, max(1, n - 60), size=m))
            L = rng.integers(8, 51, size=m)
            b = np.minimum(a + L, n - 1)
            self._loops[c] = (a.astype(np.int64), b.astype(np.int64))
        return self._loops[c]

    def _lambda_block(self, c1, lo1, hi1, c2, lo2, hi2) -> np.ndarray:
        n1, n2 = hi1 - lo1, hi2 - lo2
        lam = np.zeros((n1, n2), dtype=np.float64)
        if n1 == 0 or n2 == 0:
            return lam
        i = np.arange(lo1, hi1, dtype=np.int64)[:, None]
        j = np.arange(lo2, hi2, dtype=np.int64)[None, :]
        if c1 == c2:
            d = np.abs(i - j)
            lam += np.where(d <= self.band_bins, self.amplitude / (d + 1.0), 0.0)
            a, b = self._chrom_loops(c1)
            if a.size:
                # loops whose centre may touch the block (either orientation of the symmetric map)
                for (rlo, rhi, clo, chi, transpose) in ((lo1, hi1, lo2, hi2, False), (lo2, hi2, lo1, hi1, True)):
                    sel = (a >= rlo - 1) & (a <= rhi) & (b >= clo - 1) & (b <= chi)
                    for aa, bb in zip(a[sel], b[sel]):
                        for da in (-1, 0, 1):
                            for db in (-1, 0, 1):
                                r, cc = aa + da, bb + db
                                if r > cc:
                                    continue
                                g = self.loop_strength * (1.0 if (da == 0 and db == 0) else 0.5)
                                if not transpose:
                                    if lo1 <= r < hi1 and lo2 <= cc < hi2:
                                        lam[r - lo1, cc - lo2] += g
                                else:
                                    if r == cc:
                                        continue  # diagonal already handled by the first pass
                                    if lo1 <= cc < hi1 and lo2 <= r < hi2:
                                        lam[cc - lo1, r - lo2] += g
        res = self.binsize
        for sv in self.svs:
            sc1, p1, s1, sc2, p2, s2 = sv[:6]
            frac = float(sv[6]) if len(sv) > 6 else self.sv_fraction
            # both orientations of the symmetric map; at either breakpoint '+' keeps the
            # piece left of the position and '-' the piece right of it
            # (neoloop/assembly.py:350-383)
            for (A, B) in (((sc1, p1, s1), (sc2, p2, s2)), ((sc2, p2, s2), (sc1, p1, s1))):
                if A[0] != c1 or B[0] != c2:
                    continue
                bpA, bpB = int(A[1]) // res, int(B[1]) // res
                da = (bpA - i) if A[2] == "+" else (i - bpA)
                db = (bpB - j) if B[2] == "+" else (j - bpB)
                ok = (da >= 0) & (db >= 0)
                dist = da + db + 1
                ok = ok & (dist <= self.band_bins)
                lam += np.where(ok, frac * self.amplitude / np.where(ok, dist + 1.0, 1.0), 0.0)
        return lam

    def fetch_band(self, chrom, ndiag: int, balance="sweight", row_chunk: int = 8192) -> np.ndarray:
        """Upper band of one chromosome without ever forming the dense matrix:
        band[i, d] = balanced contact of bins (i, i+d), d < ndiag (0 beyond the chromosome end,
        0 where a weight is NaN).  Same pixels as ``matrix(...).fetch(chrom)``, including the
        contacts induced by intra-chromosomal SVs."""
        n = self._nbins(chrom)
        off = self._offsets[chrom]
        a, b = self._chrom_loops(chrom)
        band = np.zeros((n, ndiag), dtype=np.float64)
        w = self._weights(chrom, 0, n, balance if balance else "weight") if balance else None
        d = np.arange(ndiag, dtype=np.int64)[None, :]
        for lo in range(0, n, row_chunk):
            hi = min(n, lo + row_chunk)
            i = np.arange(lo, hi, dtype=np.int64)[:, None]
            j = i + d
            inside = j < n
            lam = np.where(inside & (d <= self.band_bins), self.amplitude / (d + 1.0), 0.0)
            sel = (a >= lo - 1) & (a <= hi)
            for aa, bb in zip(a[sel], b[sel]):
                for da in (-1, 0, 1):
                    for db in (-1, 0, 1):
                        r, c = aa + da, bb + db
                        if lo <= r < hi and r <= c < n and c - r < ndiag:
                            lam[r - lo, c - r] += self.loop_strength * (1.0 if (da == 0 and db == 0) else 0.5)
            res = self.binsize
            for sv in self.svs:
                sc1, p1, s1, sc2, p2, s2 = sv[:6]
                if sc1 != chrom or sc2 != chrom:
                    continue
                frac = float(sv[6]) if len(sv) > 6 else self.sv_fraction
                for (A, B) in (((p1, s1), (p2, s2)), ((p2, s2), (p1, s1))):
                    bpA, bpB = int(A[0]) // res, int(B[0]) // res
                    da = (bpA - i) if A[1] == "+" else (i - bpA)
                    db = (bpB - j) if B[1] == "+" else (j - bpB)
                    ok = (da >= 0) & (db >= 0) & inside
                    dist = da + db + 1
                    ok = ok & (dist <= self.band_bins)
                    lam += np.where(ok, frac * self.amplitude / np.where(ok, dist + 1.0, 1.0), 0.0)
            gi = (i + off).astype(np.uint64)
            gj = (np.minimum(j, n - 1) + off).astype(np.uint64)
            with np.errstate(over="ignore"):
                key = gi * _U64(0x100000001B3) + gj * _U64(0xC2B2AE3D27D4EB4F)
            salt = _U64((self.seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
            cnt = _poisson_from_hash(lam, _mix64(key ^ salt)).astype(np.float64)
            if w is not None:
                cnt = cnt * w[lo:hi, None] * w[np.minimum(j, n - 1)]
                cnt[~np.isfinite(cnt)] = 0.0
            cnt[~inside] = 0.0
            band[lo:hi] = cnt
        return band

    def _raw_block(self, c1, lo1, hi1, c2, lo2, hi2) -> np.ndarray:
        lam = self._lambda_block(c1, lo1, hi1, c2, lo2, hi2)
        gi = (np.arange(lo1, hi1, dtype=np.uint64) + _U64(self._offsets[c1]))[:, None]
        gj = (np.arange(lo2, hi2, dtype=np.uint64) + _U64(self._offsets[c2]))[None, :]
        a = np.minimum(gi, gj)
        b = np.maximum(gi, gj)
        with np.errstate(over="ignore"):
            key = a * _U64(0x100000001B3) + b * _U64(0xC2B2AE3D27D4EB4F)
        salt = _U64((self.seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        h = _mix64(key ^ salt)
        return _poisson_from_hash(lam, h)


# ----------------------------------------------------------------------------------
# explicit arrays
# ----------------------------------------------------------------------------------
class ArrayCooler(_CoolerBase):
    """Cooler-like view over explicit dense blocks.

    ``blocks[(c1, c2)]`` (c1 before-or-equal c2 in ``chromnames`` order) holds the raw
    count block (n_bins(c1) x n_bins(c2)); same-chromosome blocks must be symmetric.
    ``weights[col][c]`` are per-bin balancing weights (NaN = masked bin).
    """

    def __init__(self, binsize, chromsizes, blocks, weights, info=None):
        super().__init__(binsize, chromsizes)
        self._blocks = {k: np.asarray(v) for k, v in blocks.items()}
        self._w = {col: {c: np.asarray(a, dtype=np.float64) for c, a in d.items()}
                   for col, d in weights.items()}
        self._info = info

    @property
    def info(self):
        if self._info is not None:
            return self._info
        s = 0
        for (c1, c2), B in self._blocks.items():
            s += np.triu(B).sum() if c1 == c2 else B.sum()
        return {"sum": int(s), "nnz": int(sum((b != 0).sum() for b in self._blocks.values()))}

    def weight_columns(self):
        return tuple(self._w)

    def _weights(self, c, lo, hi, col):
        if col not in self._w:
            raise KeyError(col)                      # cooler: no such balancing column in the bin table
        return np.array(self._w[col][c][lo:hi], dtype=np.float64, copy=True)

    def _raw_block(self, c1, lo1, hi1, c2, lo2, hi2):
        if (c1, c2) in self._blocks:
            return np.array(self._blocks[(c1, c2)][lo1:hi1, lo2:hi2])
        if (c2, c1) in self._blocks:
            return np.array(self._blocks[(c2, c1)][lo2:hi2, lo1:hi1]).T
        return np.zeros((hi1 - lo1, hi2 - lo2))

    def save(self, path):
        payload = {"binsize": np.int64(self.binsize),
                   "chromnames": np.array(self.chromnames),
                   "chromsizes": np.array([self._chromsizes[c] for c in self.chromnames], dtype=np.int64)}
        for (c1, c2), B in self._blocks.items():
            payload[f"block|{c1}|{c2}"] = B
        for col, d in self._w.items():
            for c, a in d.items():
                payload[f"w|{col}|{c}"] = a
        np.savez_compressed(path, **payload)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        names = [str(x) for x in z["chromnames"]]
        sizes = dict(zip(names, (int(x) for x in z["chromsizes"])))
        blocks, weights = {}, {}
        for k in z.files:
            if k.startswith("block|"):
                _, c1, c2 = k.split("|")
                blocks[(c1, c2)] = z[k]
            elif k.startswith("w|"):
                _, col, c = k.split("|")
                weights.setdefault(col, {})[c] = z[k]
        return cls(int(z["binsize"]), sizes, blocks, weights)


# ----------------------------------------------------------------------------------
# cooler's pixel table (host copy + device residency)
# ----------------------------------------------------------------------------------
class PixelCooler(_CoolerBase):
    """Cooler-like view over a genome-wide upper-triangular pixel table in cooler's layout.

    ``bin1_offset[nbins + 1]`` indexes the pixels of every ``bin1`` (cooler's
    ``indexes/bin1_offset``), ``bin2[nnz]`` ascends inside a row with ``bin1 <= bin2``
    (global bin ids, chromosomes concatenated in ``chromnames`` order), ``count[nnz]`` is int32.
    ``weights[col]`` are genome-wide per-bin balancing weights (NaN = masked bin).

    The host methods implement the cooler surface the reference touches (the oracle and the live
    reference read the very same pixels); ``device_pixels`` hands the table to the GPU once.
    """

    FORMAT = "nlf-pixel-cooler-v1"

    def __init__(self, binsize, chromsizes, bin1_offset, bin2, count, weights, info=None):
        super().__init__(binsize, chromsizes)
        self._offsets = {}
        off = 0
        for c in self.chromnames:
            self._offsets[c] = off
            off += self._nbins(c)
        self._total_bins = off
        self.bin1_offset = np.ascontiguousarray(bin1_offset, dtype=np.int64)
        self.bin2 = np.ascontiguousarray(bin2, dtype=np.int32)
        self.count = np.ascontiguousarray(count, dtype=np.int32)
        if self.bin1_offset.size != off + 1 or self.bin2.size != self.count.size or int(self.bin1_offset[-1]) != self.bin2.size:
            raise ValueError("inconsistent pixel table")
        self._w = {col: np.ascontiguousarray(a, dtype=np.float64) for col, a in weights.items()}
        for col, a in self._w.items():
            if a.size != off:
                raise ValueError(f"weight column {col!r} has {a.size} entries for {off} bins")
        self._info = dict(info) if info else None
        self._device = {}
        self._csr = None

    # -- pickling: device handles and the scipy view stay behind ---------------------------------
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_device"] = {}
        st["_csr"] = None
        return st

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def from_cooler(cls, clr, band_bins: Optional[int] = None, trans: bool = True, row_chunk: int = 2048):
        """Materialise the pixel table of any cooler-like source.  ``band_bins`` keeps only
        intra-chromosomal pixels with ``bin2 - bin1 <= band_bins``; ``trans=False`` drops
        inter-chromosomal pixels (enough for the genome-wide call and the expected prologue)."""
        names = list(clr.chromnames)
        sizes = {c: int(clr.chromsizes[c]) for c in names}
        res = int(clr.binsize)
        nb = {c: int(math.ceil(sizes[c] / res)) for c in names}
        offs, off = {}, 0
        for c in names:
            offs[c] = off
            off += nb[c]
        rows_b2, rows_cnt, indptr = [], [], [0]
        nnz = 0
        # sources that can emit a chromosome's band directly (no dense blocks): one vectorised pass per chromosome
        band_source = band_bins is not None and not trans and hasattr(clr, "fetch_band")
        for i1, c1 in enumerate(names):
            n1 = nb[c1]
            if band_source:
                B = np.asarray(clr.fetch_band(c1, band_bins + 1, balance=False))
                if not np.array_equal(B, np.round(B)):
                    raise ValueError("PixelCooler stores integer counts (cooler's int32 `count` column)")
                rr, dd = np.nonzero(B)                        # row-major: rows ascending, distances ascending inside a row
                rows_b2.append((rr + dd + offs[c1]).astype(np.int32))
                rows_cnt.append(B[rr, dd].astype(np.int32))
                per_row = np.bincount(rr, minlength=n1)
                indptr.extend((nnz + np.cumsum(per_row)).tolist())
                nnz += int(per_row.sum())
                continue
            for lo in range(0, n1, row_chunk):
                hi = min(n1, lo + row_chunk)
                per_row_b2 = [[] for _ in range(hi - lo)]
                per_row_cnt = [[] for _ in range(hi - lo)]
                for c2 in names[i1:]:
                    if c2 != c1 and not trans:
                        continue
                    n2 = nb[c2]
                    if c2 == c1:
                        clo = lo
                        chi = n2 if band_bins is None else min(n2, hi + band_bins)
                    else:
                        clo, chi = 0, n2
                    if chi <= clo:
                        continue
                    B = np.asarray(clr._raw_block(c1, lo, hi, c2, clo, chi)) if hasattr(clr, "_raw_block") else \
                        np.nan_to_num(np.asarray(clr.matrix(balance=False).fetch((c1, lo * res, min(hi * res, sizes[c1])),
                                                                                 (c2, clo * res, min(chi * res, sizes[c2])))))
                    if not np.array_equal(B, np.round(B)):
                        raise ValueError("PixelCooler stores integer counts (cooler's int32 `count` column)")
                    rr, cc = np.nonzero(B)
                    if c2 == c1:
                        keep = (cc + clo) >= (rr + lo)
                        if band_bins is not None:
                            keep &= (cc + clo) - (rr + lo) <= band_bins
                        rr, cc = rr[keep], cc[keep]
                    vals = B[rr, cc].astype(np.int32)
                    gcol = (cc + clo + offs[c2]).astype(np.int32)
                    # np.nonzero is row-major: split per row
                    cuts = np.searchsorted(rr, np.arange(hi - lo + 1))
                    for k in range(hi - lo):
                        a, b = cuts[k], cuts[k + 1]
                        if b > a:
                            per_row_b2[k].append(gcol[a:b])
                            per_row_cnt[k].append(vals[a:b])
                for k in range(hi - lo):
                    if per_row_b2[k]:
                        b2 = np.concatenate(per_row_b2[k])
                        rows_b2.append(b2)
                        rows_cnt.append(np.concatenate(per_row_cnt[k]))
                        nnz += b2.size
                    indptr.append(nnz)
        bin2 = np.concatenate(rows_b2) if rows_b2 else np.zeros(0, np.int32)
        count = np.concatenate(rows_cnt) if rows_cnt else np.zeros(0, np.int32)
        weights = {}
        for col in ("weight", "sweight"):
            weights[col] = np.concatenate([np.asarray(clr.bins().fetch(c)[col].values, dtype=np.float64) for c in names])
        return cls(res, sizes, np.asarray(indptr, dtype=np.int64), bin2, count, weights)

    def save(self, path):
        payload = {"format": np.array(self.FORMAT), "binsize": np.int64(self.binsize),
                   "chromnames": np.array(self.chromnames),
                   "chromsizes": np.array([self._chromsizes[c] for c in self.chromnames], dtype=np.int64),
                   "bin1_offset": self.bin1_offset, "bin2": self.bin2, "count": self.count}
        for col, a in self._w.items():
            payload[f"w|{col}"] = a
        np.savez_compressed(path, **payload)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        names = [str(x) for x in z["chromnames"]]
        sizes = dict(zip(names, (int(x) for x in z["chromsizes"])))
        weights = {k.split("|")[1]: z[k] for k in z.files if k.startswith("w|")}
        return cls(int(z["binsize"]), sizes, z["bin1_offset"], z["bin2"], z["count"], weights)

    # -- cooler surface ----------------------------------------------------------------------------
    @property
    def info(self):
        if self._info is not None:
            return self._info
        return {"sum": int(self.count.sum(dtype=np.int64)), "nnz": int(self.count.size)}

    def weight_columns(self):
        return tuple(self._w)

    def _weights(self, c, lo, hi, col):
        if col not in self._w:
            raise KeyError(col)                      # cooler: no such balancing column in the bin table
        o = self._offsets[c]
        return np.array(self._w[col][o + lo:o + hi], dtype=np.float64, copy=True)

    def _upper(self):
        if self._csr is None:
            from scipy import sparse as sp
            n = self._total_bins
            self._csr = sp.csr_matrix((self.count.astype(np.float64), self.bin2, self.bin1_offset), shape=(n, n))
        return self._csr

    def _raw_block(self, c1, lo1, hi1, c2, lo2, hi2):
        """Symmetric completion of the stored upper triangle restricted to a rectangle."""
        U = self._upper()
        a0, a1 = self._offsets[c1] + lo1, self._offsets[c1] + hi1
        b0, b1 = self._offsets[c2] + lo2, self._offsets[c2] + hi2
        B = U[a0:a1, b0:b1].toarray() + U[b0:b1, a0:a1].toarray().T
        # a stored diagonal pixel (g, g) was added twice where the two ranges overlap
        g0, g1 = max(a0, b0), min(a1, b1)
        if g1 > g0:
            g = np.arange(g0, g1)
            B[g - a0, g - b0] -= U.diagonal()[g0:g1]
        return B

    # -- device residency --------------------------------------------------------------------------
    def global_bins(self, chrom, lo_bin, hi_bin):
        o = self._offsets[chrom]
        return o + int(lo_bin), o + int(hi_bin)

    def device_pixels(self, ctx=None, col="sweight"):
        """The table in HBM (one upload per process, GPU and weight column)."""
        import os
        from . import _lib
        from .scoring import DevicePixels
        ctx = ctx or _lib.Context.get()
        if col in (True, False, None):
            col = "weight"                # balance=True / raw counts: the bias column of the reference
        if col not in self._w:
            raise KeyError(col)
        key = (os.getpid(), ctx.device, col)
        if key not in self._device:
            self._device[key] = DevicePixels(self.bin1_offset, self.bin2, self.count, self._w[col], ctx)
        return self._device[key]


def load_cooler(uri: str):
    """Open a contact map given on the ``-H`` command line.

    ``*.json``  SyntheticCooler spec; ``*.npz`` PixelCooler (pixel table) or ArrayCooler (dense blocks).  Real ``.cool``/``.mcool``
    URIs need the ``cooler`` package, which is tried last and reported if absent.
    """
    if uri.endswith(".json"):
        with open(uri) as fh:
            return SyntheticCooler.from_spec(json.load(fh))
    if uri.endswith(".npz"):
        with np.load(uri, allow_pickle=False) as z:
            is_pixels = "format" in z.files and str(z["format"]) == PixelCooler.FORMAT
        return PixelCooler.load(uri) if is_pixels else ArrayCooler.load(uri)
    try:
        import cooler  # type: ignore
    except ImportError as exc:  # pragma: no cover - depends on the image
        raise RuntimeError(
            f"cannot open {uri!r}: the 'cooler' package (HDF5) is not installed; "
            "use a SyntheticCooler .json spec or an ArrayCooler .npz") from exc
    return cooler.Cooler(uri)
