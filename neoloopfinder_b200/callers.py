"""Drop-in for the scoring classes of ``neoloop/callers.py`` on the neoloop-caller path.

Same class names, constructor/method signatures, attribute names and error behaviour as the
reference; every pass over matrix data runs in the CUDA library
(``libnlf_b200.so``) through ``scoring.py``.  Host-only pieces are the ones the survey leaves on
the host: the per-block-pair regression (Spearman + isotonic + Huber on <= 400 points) and
dictionary/coordinate bookkeeping.

``Fusion`` also carries the single-SV analysis of ``assemble-complexSVs``:
``get_matrix``, ``extract_bias``, ``detect_bounds``, ``allele_slope``, ``correct_heterozygosity``.
"""
from __future__ import annotations

import logging
import math
import os

import numpy as np

from . import _lib
from .forest import DeviceForest, load_forest
from .scoring import DeviceAssembly, ScoreBatch, local_clustering_gpu, pool_buckets

log = logging.getLogger(__name__)

__all__ = ["check_increasing", "clean_inputs", "linear_regression_fit", "Fusion", "Peakachu", "combine_annotations"]


def check_increasing(x, y):
    """(warning, increasing) from the sign of Spearman's rho and its Fisher 95 % interval
    (neoloop/callers.py:19-63)."""
    from scipy.stats import spearmanr
    rho, _ = spearmanr(x, y)
    warning = False
    increasing = rho >= 0
    if rho not in (-1.0, 1.0) and len(x) > 3:
        F = 0.5 * math.log((1.0 + rho) / (1.0 - rho))
        se = 1 / math.sqrt(len(x) - 3)
        lo, hi = math.tanh(F - 1.96 * se), math.tanh(F + 1.96 * se)
        if np.sign(lo) != np.sign(hi):
            warning = True
    return warning, increasing


def clean_inputs(Xi, X, Y, increasing_bool):
    """Drop a noisy head of the curve located with an isotonic fit (neoloop/callers.py:65-89)."""
    from sklearn.isotonic import IsotonicRegression
    fit = IsotonicRegression(increasing=increasing_bool).fit(Xi, Y).predict(Xi)
    vi = np.where(np.diff(fit) < 0)[0]
    runs = np.split(vi, np.where(np.diff(vi) != 1)[0] + 1)
    si = 0
    for k in range(len(runs) - 1):
        cur, nxt = runs[k], runs[k + 1]
        if cur.size / (nxt[0] - cur[0]) > 0.5:
            si = cur[0]
            break
    if si / X.size > 0.3:
        si = vi[0]
        if si / X.size > 0.3:
            si = 0
    return X[si:], Y[si:]


def linear_regression_fit(exp1, exp2, min_samples=5):
    """Slope of local vs global expected with a Huber fit (neoloop/callers.py:370-412).
    Module-level so that it can be shipped to worker processes."""
    from sklearn.linear_model import HuberRegressor
    keys = [i for i in sorted(exp1) if i in exp2]
    Xi = np.r_[keys]
    X = np.r_[[exp2[i] for i in keys]]
    Y = np.r_[[exp1[i] for i in keys]]
    if X.size < min_samples:
        return 0, 0, False
    warning, increasing = check_increasing(Xi, Y)
    Xc, Yc = clean_inputs(Xi, X, Y, increasing)
    rscores, slopes = [], []
    for X_, Y_ in ((X, Y), (Xc, Yc)):
        X2 = X_[:, np.newaxis]
        huber = HuberRegressor().fit(X2, Y_)
        inliers = np.logical_not(huber.outliers_)
        if inliers.sum() < min_samples:
            r, s = 0, 0
        else:
            r = huber.score(X2[inliers], Y_[inliers])
            s = huber.coef_[0]
        rscores.append(r)
        slopes.append(s)
    rscores = np.r_[rscores]
    slopes = np.r_[slopes]
    return rscores.max(), slopes[np.argmax(rscores)], warning


_MAXDIS_LADDER = (200000, 400000, 500000, 1000000, 1500000, 2000000)   # callers.py:433, assembly.py:513


def rearranged_matrix(clr, pieces, reversed_flags, balance, ctx=None):
    """np.triu of the contact matrix of ``pieces`` [(chrom, start, end), ...] laid side by side, piece k
    read 3'->5' when ``reversed_flags[k]`` (the construction shared by Fusion.get_matrix, callers.py:137-237,
    and complexSV.reorganize_matrix, assembly.py:468-483).  With a resident pixel table the matrix is built
    on the GPU and only the handle comes back: returns (M or None, DeviceAssembly or None, index)."""
    index, start = [], 0
    for r in pieces:
        nbin = clr.bins().fetch(tuple(r))['chrom'].size
        index.append((start, start + nbin))
        start += nbin
    n = start
    if hasattr(clr, "device_pixels") and n > 0:
        col = 'weight' if balance == False else balance       # noqa: E712 (bias column of the reference)
        px = clr.device_pixels(ctx, col)
        spans = [clr.global_bins(*clr._bin_extent(tuple(r))) for r in pieces]
        asm = DeviceAssembly.from_pixels(px, [a for a, _b in spans], [b for _a, b in spans],
                                         [bool(f) for f in reversed_flags], balance=bool(balance))
        return None, asm, index
    M = np.zeros((n, n))
    for j1, (r1, f1, i1) in enumerate(zip(pieces, reversed_flags, index)):
        for j2, (r2, f2, i2) in enumerate(zip(pieces, reversed_flags, index)):
            if j1 > j2:
                continue
            sub = clr.matrix(balance=balance).fetch(tuple(r1), tuple(r2))
            if f1:
                sub = sub[::-1, :]
            if f2:
                sub = sub[:, ::-1]
            M[i1[0]:i1[1], i2[0]:i2[1]] = sub
    M = np.triu(M)
    if balance:
        M[np.isnan(M)] = 0
    return M, None, index


class Fusion(object):
    """One SV breakpoint pair: the rearranged matrix of its two flanks, the induced-contact bounds and
    the distance-decay slopes against the global expected (neoloop/callers.py:92-522).
    ``assemble-complexSVs`` filters SVs with it (neoloop/assembly.py:12-27); ``complexSV`` inherits the
    regression and gap-removal parts.  Matrix passes (the masked per-diagonal means of ``_extract_xy``)
    run on the GPU; PCA / regression on at most a few hundred points stay on the host."""

    def __init__(self, clr, c1, c2, p1, p2, s1, s2, sv_type, span=5000000,
                 col='sweight', trim=True, protocol='insitu', expected_values=None):
        from .util import find_chrom_pre
        self.clr = clr
        self.p1, self.p2, self.s1, self.s2 = p1, p2, s1, s2
        self.res = clr.binsize
        pre = find_chrom_pre(list(clr.chromnames))
        self.c1 = pre + c1.lstrip('chr')
        self.c2 = pre + c2.lstrip('chr')
        self.chromsize1 = int(clr.chromsizes[self.c1])
        self.chromsize2 = int(clr.chromsizes[self.c2])
        self.balance_type = col
        self.protocol = protocol
        self.name = ','.join(map(str, [sv_type, c1.lstrip('chr'), p1, s1, c2.lstrip('chr'), p2, s2]))
        self._asm = None
        self._fusion_matrix = None
        self._hcm = None
        self.get_matrix(span, col, trim)
        if expected_values is None:
            self.load_expected()
        else:
            self.expected = expected_values

    # ------------------------------------------------------------------ geometry + matrix
    def _flanks(self, span, trim):
        """(k_p, pieces, reversed flags): '+' at breakpoint 1 keeps the side left of it, '-' at breakpoint 2
        the side right of it; the other two cases are read backwards.  Intra-chromosomal inversions and
        '-+' pairs stop halfway between the breakpoints (callers.py:146-214)."""
        res = self.res
        b1, b2, r = self.p1 // res, self.p2 // res, span // res
        intra = self.c1 == self.c2
        half = int((b2 - b1) / 2)
        if self.s1 == '+':
            lo1, hi1 = max(b1 - r, 0) * res, min(b1 * res + res, self.chromsize1)
        else:
            lo1 = b1 * res
            if intra and self.s2 == '-':
                hi1 = min((b1 + r + 1) * res, (b2 - half) * res if trim else b2 * res - res)
            elif intra:
                hi1 = min(b1 + r + 1, b2 - half - 1) * res
            else:
                hi1 = min((b1 + r + 1) * res, self.chromsize1)
        if self.s2 == '-':
            lo2, hi2 = b2 * res, min((b2 + r + 1) * res, self.chromsize2)
        else:
            hi2 = min(b2 * res + res, self.chromsize2)
            if intra and self.s1 == '+':
                lo2 = max(b2 - r, (b1 + half) if trim else (b1 + 1)) * res
            elif intra:
                lo2 = max(b2 - r, b1 + half + 1) * res
            else:
                lo2 = max(b2 - r, 0) * res
        rev = [self.s1 == '-', self.s2 == '+']
        k_p = ([hi1, lo1] if rev[0] else [lo1, hi1]) + ([hi2, lo2] if rev[1] else [lo2, hi2])
        return k_p, [(self.c1, lo1, hi1), (self.c2, lo2, hi2)], rev

    def get_matrix(self, span, col, trim):
        self.k_p, pieces, rev = self._flanks(span, trim)
        self._pieces, self._rev = pieces, rev
        M, asm, index = rearranged_matrix(self.clr, pieces, rev, col, getattr(self, '_ctx', None))
        self._fusion_matrix, self._asm, self._index = M, asm, index
        self._n_bins = index[-1][1]
        self.fusion_point = index[0][1]

    @property
    def fusion_matrix(self):
        if self._fusion_matrix is None and self._asm is not None:
            self._fusion_matrix = self._asm.fusion_matrix()
        return self._fusion_matrix

    @fusion_matrix.setter
    def fusion_matrix(self, M):
        self._fusion_matrix = M

    def _device_assembly(self) -> DeviceAssembly:
        if self._asm is None:
            self._asm = DeviceAssembly(self.fusion_matrix, np.zeros(self._n_bins), self._index, getattr(self, '_ctx', None))
        return self._asm

    def extract_bias(self, col='sweight'):
        """1 / weight per bin of the rearranged axis, 0 where the weight is 0 or NaN (callers.py:240-262)."""
        if col is False or col is None:
            col = 'weight'
        parts = []
        for r, f in zip(self._pieces, self._rev):
            t = np.asarray(self.clr.bins().fetch(tuple(r))[col].values, dtype=np.float64)
            if f:
                t = t[::-1]
            ok = ~((t == 0) | np.isnan(t))
            w = np.zeros_like(t)
            w[ok] = 1 / t[ok]
            parts.append(w)
        self.bias = np.concatenate(parts)

    # ------------------------------------------------------------------ induced-contact bounds
    def detect_bounds(self):
        """First principal component of the smoothed row / column correlation of the junction block
        locates where the induced contacts stop on either side (callers.py:265-312)."""
        from scipy.ndimage import gaussian_filter
        from sklearn.decomposition import PCA
        n, junc = self._n_bins, self.fusion_point
        self.up_i, self.down_i = 0, n - 1
        self.up_var_ratio = self.down_var_ratio = 0
        inter = self.fusion_matrix[:junc, junc:]
        rows, cols = inter.sum(axis=1) != 0, inter.sum(axis=0) != 0
        if rows.sum() < 10 or cols.sum() < 10:
            return
        block = inter[rows][:, cols]
        for upstream in (True, False):
            corr = gaussian_filter(np.corrcoef(block, rowvar=upstream), sigma=1)
            try:
                pca = PCA(n_components=3, whiten=True)
                pc1 = pca.fit_transform(corr)[:, 0]
                ratio = pca.explained_variance_ratio_[0]
                loc = self.locate(pc1, left_most=not upstream)
                if upstream:
                    self.up_i, self.up_var_ratio = np.where(rows)[0][loc], ratio
                else:
                    self.down_i, self.down_var_ratio = np.where(cols)[0][loc] + junc, ratio
            except Exception:     # noqa: BLE001 - the reference falls back to the full flank on any failure
                if upstream:
                    self.up_i, self.up_var_ratio = 0, 0
                else:
                    self.down_i, self.down_var_ratio = n - 1, 0

    def extend_stretches(self, arr, min_seed_len=2, max_gap=1, min_stripe_len=5):
        """Runs of consecutive indices (>= min_seed_len long), joined across gaps of at most max_gap,
        reported as [first, last + 1] when they span min_stripe_len (callers.py:315-333)."""
        arr = np.sort(arr)
        runs = [p for p in np.split(arr, np.where(np.diff(arr) != 1)[0] + 1) if len(p) >= min_seed_len]
        out = []
        cur_lo, cur_hi = runs[0][0], runs[0][-1]          # IndexError when nothing qualifies, as in the reference

        def flush():
            if cur_hi - cur_lo + 1 >= min_stripe_len:
                out.append([cur_lo, cur_hi + 1])
        for p in runs[1:]:
            if p[0] - cur_hi < max_gap + 2:
                cur_hi = p[-1]
            else:
                flush()
                cur_lo, cur_hi = p[0], p[-1]
        flush()
        return out

    def locate(self, pca1, left_most=True):
        loci = self.extend_stretches(np.where(pca1 >= 0)[0]) + self.extend_stretches(np.where(pca1 <= 0)[0])
        loci.sort()
        return loci[0][1] - 1 if left_most else loci[-1][0]

    # ------------------------------------------------------------------ slopes
    def _region_means(self, up_i, down_i):
        """Masked per-diagonal means (Fusion._extract_xy, callers.py:348-367) of the upstream, downstream
        and junction regions, up to 2 Mb: one kernel launch for the three of them."""
        junc = self.fusion_point
        rects = [(up_i, junc, up_i, junc), (junc, down_i, junc, down_i), (up_i, junc, junc, down_i)]
        maxdiag = _MAXDIS_LADDER[-1] // self.res
        mean, _cnt = self._device_assembly().rect_means(rects, maxdiag, mink=3)
        return mean

    def allele_slope(self):
        """Best Huber fit of local vs global expected over six distance cut-offs for the three regions, first
        between the detected bounds, then -- unless all three fits are good -- over the whole flanks
        (callers.py:415-478).  The six cut-offs are prefixes of one per-diagonal table."""
        self.extract_bias(col=self.balance_type)
        asm = self._device_assembly()
        asm.set_bias(self.bias)
        res = self.res
        for up_i, down_i in ((self.up_i, self.down_i), (0, self.bias.size - 1)):
            mean = self._region_means(int(up_i), int(down_i))
            best = []
            for row in mean:
                kept = np.where(~np.isnan(row))[0]
                memo, rs, ss = {}, [], []
                for maxdis in _MAXDIS_LADDER:
                    upto = kept[kept <= maxdis // res]
                    if len(upto) not in memo:
                        memo[len(upto)] = self.linear_regression({int(i): row[i] for i in upto}, self.expected)
                    r, s, _w = memo[len(upto)]
                    rs.append(r)
                    ss.append(s)
                rs = np.r_[rs]
                best.append((rs.max(), ss[rs.argmax()]))
            (self.u_g_r, self.u_g_s), (self.d_g_r, self.d_g_s), (self.m_g_r, self.m_g_s) = best
            if self.u_g_r > 0.6 and self.d_g_r > 0.6 and self.m_g_r > 0.6 and self.m_g_s > 0:
                break

    def correct_heterozygosity(self):
        """Relative slopes of the junction against each flank; ``hcm`` (the rescaled matrix,
        callers.py:482-502) is materialised when the attribute is read."""
        self.u_s = self.d_s = 0
        self._hcm = None
        if self.u_g_r > 0.6 and self.d_g_r > 0.6:
            self.u_s = self.m_g_s / self.u_g_s
            self.d_s = self.m_g_s / self.d_g_s

    @property
    def hcm(self):
        if self._hcm is None:
            junc = self.fusion_point
            h = self.fusion_matrix.copy()
            if self.u_g_r > 0.6 and self.d_g_r > 0.6:
                slope, ms = max(self.u_g_s, self.d_g_s), min(self.u_g_s, self.d_g_s)
                h[:junc, :junc] = h[:junc, :junc] / slope
                h[junc:, junc:] = h[junc:, junc:] / slope
                if self.m_g_r > 0.6 and self.m_g_s / ms > 0.1:
                    box = (slice(self.up_i, junc), slice(junc, self.down_i))
                    h[box] = h[box] * min(1 / self.m_g_s, 5 / slope)
            self._hcm = h
        return self._hcm

    def load_expected(self):
        """Bundled GM12878 in-situ expected tables (neoloop/callers.py:120-134): {distance in bins: mean contact}
        for 5, 10, 20, 25, 40 and 50 kb; any other resolution gets the reference's empty table.  The six pickles
        are the reference's own data files and are not duplicated here: they are looked up in NLF_EXPECTED_DIR,
        in this package's ``data/`` folder and in an installed ``neoloop`` package.  A missing table for a
        supported resolution is an error -- with an empty table every regression would silently return (0, 0, False)
        and every SV / assembly would be rejected."""
        import joblib
        name = {5000: "5k", 10000: "10k", 20000: "20k", 25000: "25k", 40000: "40k", 50000: "50k"}.get(self.res)
        self.expected = {}
        if name is None:
            return
        fname = f"gm.insitu.expected.{name}.pkl"
        folders = [os.environ.get("NLF_EXPECTED_DIR"), os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")]
        try:
            import importlib.util
            spec = importlib.util.find_spec("neoloop")
            if spec is not None and spec.submodule_search_locations:
                folders.append(os.path.join(list(spec.submodule_search_locations)[0], "data"))
        except (ImportError, ValueError):
            pass
        for folder in folders:
            if folder and os.path.exists(os.path.join(folder, fname)):
                self.expected = joblib.load(os.path.join(folder, fname))
                return
        raise FileNotFoundError(
            f"{fname} not found (searched NLF_EXPECTED_DIR, {folders[1]} and an installed neoloop package): pass "
            "expected_values=calculate_expected(clr, ...) or point NLF_EXPECTED_DIR at the reference's neoloop/data")

    def linear_regression(self, exp1, exp2, min_samples=5):
        """Slope of local vs global expected with a Huber fit (neoloop/callers.py:370-412)."""
        return linear_regression_fit(exp1, exp2, min_samples)

    def remove_gaps(self):
        """Gap-free matrix and gap-free -> original index map (neoloop/callers.py:504-522).
        For an assembly (complexSV) the column scan and the compaction run on the GPU and ``gap_free`` is
        materialised on the host only if somebody reads the attribute; a plain Fusion works on its ``hcm``."""
        if getattr(self, "_scale_plan", None) is None and not hasattr(self, "blocks"):
            h = self.hcm
            keep = np.where(~(h.sum(axis=0) == 0))[0]
            if 5 * 2 + 1 > keep.size:
                keep = np.arange(h.shape[0])
            self.index_map = dict(zip(np.arange(len(keep)), keep.astype(np.int64)))
            self._gap_free = h[np.ix_(keep, keep)]
            return
        asm = self._device_assembly()
        kept = asm.kept_index()
        self.index_map = dict(zip(np.arange(len(kept)), kept.astype(np.int64)))
        self._gap_free = None

    @property
    def gap_free(self):
        if getattr(self, "_gap_free", None) is None:
            self._gap_free = _DeviceBackedMatrix(self._device_assembly())
        return self._gap_free


class _DeviceBackedMatrix:
    """Lazy stand-in for the gap-free ndarray: remembers the device assembly it came from, so
    ``Peakachu(fu.gap_free, ...)`` scores straight from HBM; any other consumer that treats it as
    an array (np.asarray, indexing, .sum() ...) triggers one device-to-host copy."""

    def __init__(self, asm):
        self._nlf_assembly = asm
        self._host = None
        n = len(asm.kept_index())
        self.shape = (n, n)
        self.ndim = 2
        self.dtype = np.dtype(np.float64)

    def _materialise(self):
        if self._host is None:
            self._host = self._nlf_assembly.gap_free()
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        return self._materialise()[key]

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._materialise(), name)


class Peakachu():
    """Loop scoring of one gap-free matrix (neoloop/callers.py:525-831) on the GPU."""

    def __init__(self, matrix, upper=4000000, res=10000, protocol='insitu'):
        self.w = 7
        self.r = res
        self.protocol = protocol
        self._upper = upper // res
        self._ctx = _lib.Context.get()
        self._batch = ScoreBatch(4, self._upper, self._ctx)
        asm = getattr(matrix, "_nlf_assembly", None)
        if asm is not None:
            self._job = self._batch.add_assembly(asm)
            self._matrix = matrix
        else:
            self._matrix = np.ascontiguousarray(np.asarray(matrix), dtype=np.float64)
            self._job = self._batch.add_dense(self._matrix)
        self._shape = (self._batch.n[self._job],) * 2
        self._cand = None
        self._M = None
        self.model_path = None
        self._forest = None
        self._model = None
        self.exp_arr = None
        self._scored_thre = None
        # the reference's constructor fails here when no diagonal is longer than 5 bins
        # (IsotonicRegression.fit on an empty sample raises ValueError, callers.py:544-551)
        try:
            self._batch.counts()
        except _lib.NLFError as exc:
            if exc.code == _lib.NLF_ERR_EMPTY_FIT:
                raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required "
                                 "by IsotonicRegression (matrix has no diagonal longer than 5 bins)") from exc
            raise

    # -- attributes of the reference object -------------------------------------------------
    @property
    def M(self):
        """CSR matrix of the finite band (callers.py:531-535); built on the host on demand."""
        if self._M is None:
            from scipy import sparse
            m = np.asarray(self._matrix)
            R, Cc = m.nonzero()
            data = m[R, Cc]
            ok = np.isfinite(data) & (Cc - R > -14) & (Cc - R < self._upper + 14)
            self._M = sparse.csr_matrix((data[ok], (R[ok], Cc[ok])), shape=m.shape)
        return self._M

    def _candidates(self):
        if self._cand is None:
            r, c = self._batch.candidates(self._job)
            self._cand = (r.astype(np.int64), c.astype(np.int64))
        return self._cand

    @property
    def ridx(self):
        return self._candidates()[0]

    @property
    def cidx(self):
        return self._candidates()[1]

    @property
    def model(self):
        if self._model is None and self.model_path is not None:
            import joblib
            self._model = joblib.load(self.model_path)
        return self._model

    # -- reference methods -------------------------------------------------------------------
    def get_candidates(self, lower, upper):
        """Kept for API compatibility: candidates are computed on the GPU for the (4, upper)
        range fixed by the constructor (callers.py:536)."""
        if lower != 4 or upper != self._upper:
            raise ValueError("neoloopfinder_b200 computes candidates for lower=4 and the constructor's upper")
        self._cand = None
        self._candidates()

    def load_models(self, model_fil_path):
        """callers.py:572-577, but the forest is flattened once per file and kept on the GPU."""
        if isinstance(model_fil_path, DeviceForest):
            self._forest = model_fil_path
        else:
            self.model_path = model_fil_path
            self._forest = load_forest(model_fil_path, self._ctx)
        self.w = self._forest.w

    def load_expected(self, expected_values):
        self.exp_arr = np.r_[[expected_values[i] for i in sorted(expected_values)]]

    def getwindow(self, coords, w):
        """Features (float64, rows x (2w+1)^2) and coordinates of the windows that pass the
        reference's masks (callers.py:583-612)."""
        if len(coords) < 2:
            return [], []
        coords = np.r_[coords]
        fea, clist = self._batch.getwindow(self._job, coords[:, 0], coords[:, 1], w, self.exp_arr)
        if len(fea) == 0 and self._n_inside(coords, w) < 2:
            return [], []
        return fea, clist.astype(np.int64)

    def _n_inside(self, coords, w):
        xi, yi = coords[:, 0], coords[:, 1]
        return int(((xi - w >= 0) & (yi + w + 1 <= self._shape[0]) & (yi - xi - 2 >= w)).sum())

    def _score(self, thre):
        if self._forest is None or self.exp_arr is None:
            raise RuntimeError("call load_expected() and load_models() before predict()")
        if self._scored_thre is None:
            self._batch.score(self._forest, self.exp_arr, thre, self.w)
            self._scored_thre = thre
        elif self._scored_thre != thre:
            # the probabilities stay in HBM: only the threshold pass (callers.py:623-628) runs again
            self._batch.rethreshold(thre)
            self._scored_thre = thre

    def predict(self, thre, no_pool=False, min_count=2, index_map=None, chains=None):
        """[(i, j, probability)] of the called loops (callers.py:614-639)."""
        self._score(thre)
        n_cand, n_scored, n_don = self._batch.counts()
        if n_scored[self._job] < 2:
            return []
        di, dj, dp, dv = self._batch.donuts(self._job)
        Donuts = {(int(a), int(b)): v for a, b, v in zip(di, dj, dv)}
        prob = {(int(a), int(b)): float(q) for a, b, q in zip(di, dj, dp)}
        if not no_pool:
            keep = self.local_clustering(Donuts, min_count=min_count, index_map=index_map, chains=chains)
        else:
            keep = Donuts
        return [(i, j, prob[(i, j)]) for i, j in keep]

    def scored_pixels(self):
        """(clist, probas) of the last predict(): every scored window in candidate order."""
        i, j, p = self._batch.scored(self._job)
        return np.stack([i, j], 1).astype(np.int64), p

    def _context(self):
        """The CUDA context of this object; a subclass that never ran ``Peakachu.__init__`` and only set ``self.r``
        (``class Triangle(Peakachu)``, neoloop/visualize/core.py:23, :271, :315) gets the process default."""
        ctx = self.__dict__.get("_ctx")
        if ctx is None:
            ctx = self._ctx = _lib.Context.get()
        return ctx

    def local_clustering(self, Donuts, min_count=2, r=15000, index_map=None, chains=None):
        """Pool thresholded pixels per (chain, chain) region (callers.py:668-697)."""
        if not len(Donuts):
            return set()
        ij = np.array(list(Donuts), dtype=np.int64).reshape(-1, 2)
        val = np.array([Donuts[(i, j)] for i, j in ij], dtype=np.float64)
        if chains is None:
            ci = cj = None
        else:
            if index_map is None:
                xs, ys = ij[:, 0], ij[:, 1]
            else:
                xs = np.array([index_map[i] for i in ij[:, 0]])
                ys = np.array([index_map[j] for j in ij[:, 1]])
            ci = np.array([chains[x] for x in xs])
            cj = np.array([chains[y] for y in ys])
        return local_clustering_gpu(ij, val, self.r, min_count, r, ci, cj, self._context())

    def _local_clustering(self, Donuts, min_count=1, r=15000):
        """One region (callers.py:699-738) -> list of kept (i, j)."""
        if not len(Donuts):
            return []
        ij = np.array(list(Donuts), dtype=np.int64).reshape(-1, 2)
        val = np.array([Donuts[(i, j)] for i, j in ij], dtype=np.float64)
        mask = pool_buckets([(ij[:, 0], ij[:, 1], val)], self.r, min_count, r, self._context())[0]
        return [(int(a), int(b)) for a, b in ij[mask]]

    def find_anchors(self, pos, min_count=3, min_dis=15000, wlen=40000):
        """Anchors {(summit, left, right)} of one coordinate list (callers.py:740-784): histogram peaks at least
        ``min_count`` high and ``min_dis`` apart, widened to the foot of each peak, merged on overlap.  Runs on the
        GPU (the routine the pooling kernel uses); below 10 kb ties between equal peak heights follow
        numpy.argsort, exactly like the pooling path."""
        import ctypes as C
        if self.r <= 10000 and min(wlen // self.r, 20) <= 1:
            raise ValueError("`wlen` must be larger than 1, was %d" % min(wlen // self.r, 20))      # scipy.signal.peak_widths
        pos = np.ascontiguousarray(list(pos), dtype=np.int32)
        if pos.size == 0:
            raise ValueError("min() arg is an empty sequence")              # Counter of nothing, as in the reference
        L = _lib.load()
        order = None
        if max(min_dis // self.r, 2) > 2:
            off = np.array([0, pos.size], dtype=np.int64)
            prio = np.empty(pos.size, dtype=np.float64)
            poff = np.zeros(2, dtype=np.int64)
            amb = C.c_int(0)
            _lib.check(L.nlf_pool_peak_heights(self._context().h, 1, _lib.ptr(off, C.c_int64), _lib.ptr(pos, C.c_int), int(self.r),
                                               int(min_count), int(min_dis), _lib.ptr(prio, C.c_double),
                                               _lib.ptr(poff, C.c_int64), C.byref(amb)))
            if amb.value:
                order = np.zeros(pos.size, dtype=np.int32)
                order[:int(poff[1])] = np.argsort(prio[:int(poff[1])])        # default kind, as scipy calls it
        cap = int(pos.max() - pos.min()) + 1
        out = np.empty(3 * cap, dtype=np.int32)
        n = C.c_int(0)
        _lib.check(L.nlf_pool_find_anchors_wlen(self._context().h, _lib.ptr(pos, C.c_int), int(pos.size), int(self.r), int(min_count),
                                                int(min_dis), int(wlen), None if order is None else _lib.ptr(order, C.c_int),
                                                _lib.ptr(out, C.c_int), cap, C.byref(n)))
        return {(int(a), int(b), int(c)) for a, b, c in out[:3 * n.value].reshape(-1, 3)}

    def _cluster_core(self, sort_list, r, visited, final_list):
        """DBSCAN + greedy pooling of one sorted point list (callers.py:786-831): seeds are appended to
        ``final_list``, every pooled point joins ``visited``.  ``r`` is in bins."""
        import ctypes as C
        pts = [p[1] for p in sort_list]
        m = len(pts)
        if m == 0:
            return
        ij = np.ascontiguousarray(pts, dtype=np.int32).reshape(m, 2)
        pi, pj = np.ascontiguousarray(ij[:, 0]), np.ascontiguousarray(ij[:, 1])
        fin = np.zeros(m, dtype=np.uint8)
        pooled = np.zeros(m, dtype=np.uint8)
        _lib.check(_lib.load().nlf_pool_cluster_core(self._context().h, _lib.ptr(pi, C.c_int), _lib.ptr(pj, C.c_int), m, int(r),
                                                     _lib.ptr(fin, C.c_ubyte), _lib.ptr(pooled, C.c_ubyte)))
        for k in np.flatnonzero(fin):
            final_list.append((int(pi[k]), int(pj[k])))
        visited.update((int(pi[k]), int(pj[k])) for k in np.flatnonzero(pooled))

    def close(self):
        self._batch.close()


def combine_annotations(byres):
    """Merge loop calls of several resolutions, finest first: a coarser loop is dropped when a
    finer call falls inside its two bins, and its labels go to those finer calls
    (neoloop/callers.py:834-868).  Returns sorted rows [c1, s1, e1, c2, s2, e2, label]."""
    levels = sorted(byres)
    pool = byres[levels[0]]
    seen = [levels[0]]
    for res in levels[1:]:
        current = byres[res]
        for loop in current:
            for pr in seen:
                hits = 0
                for i in range(loop[1] // pr, loop[2] // pr + 1):
                    for j in range(loop[4] // pr, loop[5] // pr + 1):
                        key = (loop[0], i * pr, i * pr + pr, loop[3], j * pr, j * pr + pr)
                        if key in pool:
                            hits += 1
                            pool[key].extend(current[loop])
                if hits == 0:
                    pool[loop] = current[loop]
        seen.append(res)
    rows = []
    for key in sorted(pool):
        labels = [",".join([t[0], str(t[1]), str(t[2])]) for t in set(pool[key])]
        rows.append(list(map(str, key)) + [",".join(labels)])
    return rows
