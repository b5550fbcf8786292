"""TEST INFRASTRUCTURE ONLY -- Python restatement of the reference's neoloop-caller path.

A port of what ``pipeline()`` executes (scripts/neoloop-caller:185-218) with the SAME third-party
call sites (numpy reductions, scipy.sparse CSR indexing, scipy.ndimage.gaussian_filter, sklearn
IsotonicRegression / HuberRegressor / RandomForest.predict_proba / dbscan, scipy.signal
find_peaks / peak_widths) and the SAME loop structure (one Python iteration per window, the
nested expected-fill loop, per-assembly model unpickling), so that its run time is the
reference's run time.  It exists because the reference tree does not travel to the GPU box:
there it is (a) the timed CPU baseline (``bench.py`` cpu_baseline / --impl reference, kind
"port") and (b) an end-to-end checker next to the C oracle.

Pinned against the real reference in the build container by tests/test_oracle_vs_reference.py
and, through the golden fixtures, by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from collections import Counter

import numpy as np
from scipy import sparse
from scipy.ndimage import gaussian_filter

# ----------------------------------------------------------------------------- util.py:403-524


def chrom_expected(args):
    clr, c, maxdis, balance = args
    hic = clr.matrix(balance=balance, sparse=True).fetch(c)
    col = 'weight' if type(balance) == bool else balance
    weights = clr.bins().fetch(c)[col].values
    good = np.isfinite(weights) & (weights > 0)
    n = hic.shape[0]
    maxdis = min(n - 1, maxdis)
    table = {}
    for i in range(maxdis + 1):
        valid = good if i == 0 else good[:-i] * good[i:]
        cur = hic.diagonal(i)[valid]
        if cur.size > 0:
            table[i] = [cur.sum(), cur.size]
    return table


def calculate_expected(clr, chroms, maxdis=5000000, balance=True, nproc=1):
    from sklearn.isotonic import IsotonicRegression
    maxdis = maxdis // clr.binsize
    results = [chrom_expected((clr, c, maxdis, balance)) for c in chroms]
    expected = {}
    for i in range(maxdis + 1):
        nume = 0
        denom = 0
        for extract in results:
            if i in extract:
                nume += extract[i][0]
                denom += extract[i][1]
        if (denom > 10) and (nume > 0):
            expected[i] = nume / denom
    IR = IsotonicRegression(increasing=False, out_of_bounds='clip')
    _d = np.r_[sorted(expected)]
    _y = np.r_[[expected[i] for i in _d]]
    IR.fit(_d, _y)
    d = np.arange(maxdis + 1)
    return dict(zip(d, IR.predict(d)))


def oe_window(sub, exp_bychrom, x, y, w):
    """util.distance_normaize_core (util.py:470-491), including its nested fill loop."""
    rows = np.arange(x - w, x + w + 1).reshape((2 * w + 1, 1))
    cols = np.arange(y - w, y + w + 1)
    D = np.abs(cols - rows)
    lo, hi = D.min(), D.max()
    if hi >= exp_bychrom.size:
        return sub
    exp_sub = np.zeros(sub.shape)
    for d in range(lo, hi + 1):
        xi, yi = np.where(D == d)
        for i, j in zip(xi, yi):
            exp_sub[i, j] = exp_bychrom[d]
    return sub / exp_sub


def image_normalize(arr_2d):
    return (arr_2d - arr_2d.min()) / (arr_2d.max() - arr_2d.min())


def distance_normalize(arr_pool, exp_bychrom, xi, yi, w):
    clist, fea = [], []
    for k in range(xi.size):
        x, y = xi[k], yi[k]
        window = arr_pool[k]
        bx, by = np.where(np.isnan(window))
        for i_, j_ in zip(bx, by):
            window[i_, j_] = 0
        if np.count_nonzero(window) < window.size * 0.1:
            continue
        ll_mean = window[:w, :w].mean()
        if ll_mean > 0:
            center = window[w, w]
            if center / ll_mean > 0.1:
                fea.append(oe_window(window, exp_bychrom, x, y, w))
                clist.append((x, y))
    return fea, clist


# ----------------------------------------------------------------------------- callers.py:19-89, 348-412
def check_increasing(x, y):
    from scipy.stats import spearmanr
    rho, _ = spearmanr(x, y)
    warning = False
    increasing_bool = rho >= 0
    if rho not in [-1.0, 1.0] and len(x) > 3:
        F = 0.5 * math.log((1. + rho) / (1. - rho))
        F_se = 1 / math.sqrt(len(x) - 3)
        rho_0 = math.tanh(F - 1.96 * F_se)
        rho_1 = math.tanh(F + 1.96 * F_se)
        if np.sign(rho_0) != np.sign(rho_1):
            warning = True
    return warning, increasing_bool


def clean_inputs(Xi, X, Y, increasing_bool):
    from sklearn.isotonic import IsotonicRegression
    Y1 = IsotonicRegression(increasing=increasing_bool).fit(Xi, Y).predict(Xi)
    vi = np.where(np.diff(Y1) < 0)[0]
    pieces = np.split(vi, np.where(np.diff(vi) != 1)[0] + 1)
    si = 0
    for i in range(len(pieces) - 1):
        p1, p2 = pieces[i], pieces[i + 1]
        if p1.size / (p2[0] - p1[0]) > 0.5:
            si = p1[0]
            break
    if si / X.size > 0.3:
        si = vi[0]
        if si / X.size > 0.3:
            si = 0
    return X[si:], Y[si:]


def extract_xy(M, res, valid_pixel, maxdis=2000000, mink=3):
    x, y = np.where(valid_pixel)
    maxdis = maxdis // res
    dis = y - x
    local_exp = {}
    for i in range(2, M.shape[0] - 1):
        if i > maxdis:
            break
        mask = dis == i
        if mask.sum() >= mink:
            tmp = M[x[mask], y[mask]].mean()
            if tmp > 0:
                local_exp[i] = tmp
    return local_exp


def linear_regression(exp1, exp2, min_samples=5):
    from sklearn.linear_model import HuberRegressor
    X, Y, Xi = [], [], []
    for i in sorted(exp1):
        if i in exp2:
            Xi.append(i)
            X.append(exp2[i])
            Y.append(exp1[i])
    X, Y, Xi = np.r_[X], np.r_[Y], np.r_[Xi]
    if X.size < min_samples:
        return 0, 0, False
    warning, increasing_bool = check_increasing(Xi, Y)
    X_clean, Y_clean = clean_inputs(Xi, X, Y, increasing_bool)
    rscores, slopes = [], []
    for X_, Y_ in ([X, Y], [X_clean, Y_clean]):
        X_ = X_[:, np.newaxis]
        huber = HuberRegressor().fit(X_, Y_)
        inlier_mask = np.logical_not(huber.outliers_)
        if inlier_mask.sum() < min_samples:
            rscore, slope = 0, 0
        else:
            rscore = huber.score(X_[inlier_mask], Y_[inlier_mask])
            slope = huber.coef_[0]
        rscores.append(rscore)
        slopes.append(slope)
    rscores, slopes = np.r_[rscores], np.r_[slopes]
    return rscores.max(), slopes[np.argmax(rscores)], warning


# ----------------------------------------------------------------------------- callers.py:92-502, assembly.py:12-27
class FusionPort:
    """One SV as assemble-complexSVs analyses it (row f-3): matrix of the two flanks, induced-contact bounds,
    slopes of the upstream / downstream / junction regions against the global expected."""

    def __init__(self, clr, c1, c2, p1, p2, s1, s2, sv_type, span=5000000, col='sweight', trim=True, expected_values=None):
        self.clr, self.res = clr, clr.binsize
        self.p1, self.p2, self.s1, self.s2 = p1, p2, s1, s2
        pre = 'chr' if list(clr.chromnames)[0].startswith('chr') else ''
        self.c1, self.c2 = pre + c1.lstrip('chr'), pre + c2.lstrip('chr')
        self.size1, self.size2 = clr.chromsizes[self.c1], clr.chromsizes[self.c2]
        self.col = col
        self.name = ','.join(map(str, [sv_type, c1.lstrip('chr'), p1, s1, c2.lstrip('chr'), p2, s2]))
        self.expected = expected_values
        self.get_matrix(span, col, trim)

    def get_matrix(self, span, col, trim):
        """callers.py:137-237: the four strand combinations, each with its intra-chromosomal special case."""
        res, fetch = self.res, self.clr.matrix(balance=col).fetch
        p1, p2, r = self.p1 // res, self.p2 // res, span // res
        same = self.c1 == self.c2
        mid = int((p2 - p1) / 2)
        if self.s1 == '+' and self.s2 == '+':
            lo2 = max(p2 - r, 0) if not same else (max(p2 - r, p1 + mid) if trim else max(p2 - r, p1 + 1))
            k_p = [max(p1 - r, 0) * res, min(p1 * res + res, self.size1), min(p2 * res + res, self.size2), lo2 * res]
            M1 = fetch((self.c1, k_p[0], k_p[1]))
            M2 = fetch((self.c2, k_p[3], k_p[2]))[::-1, ::-1]
            M3 = fetch((self.c1, k_p[0], k_p[1]), (self.c2, k_p[3], k_p[2]))[:, ::-1]
        elif self.s1 == '+' and self.s2 == '-':
            k_p = [max(p1 - r, 0) * res, min(p1 * res + res, self.size1), p2 * res, min((p2 + r + 1) * res, self.size2)]
            M1 = fetch((self.c1, k_p[0], k_p[1]))
            M2 = fetch((self.c2, k_p[2], k_p[3]))
            M3 = fetch((self.c1, k_p[0], k_p[1]), (self.c2, k_p[2], k_p[3]))
        elif self.s1 == '-' and self.s2 == '-':
            if not same:
                hi1 = min((p1 + r + 1) * res, self.size1)
            elif trim:
                hi1 = min((p1 + r + 1) * res, (p2 - mid) * res)
            else:
                hi1 = min((p1 + r + 1) * res, p2 * res - res)
            k_p = [hi1, p1 * res, p2 * res, min((p2 + r + 1) * res, self.size2)]
            M1 = fetch((self.c1, k_p[1], k_p[0]))[::-1, ::-1]
            M2 = fetch((self.c2, k_p[2], k_p[3]))
            M3 = fetch((self.c1, k_p[1], k_p[0]), (self.c2, k_p[2], k_p[3]))[::-1, :]
        else:
            if not same:
                k_p = [min((p1 + r + 1) * res, self.size1), p1 * res, min(p2 * res + res, self.size2), max(p2 - r, 0) * res]
            else:
                k_p = [min(p1 + r + 1, p2 - mid - 1) * res, p1 * res, min(p2 * res + res, self.size2),
                       max(p2 - r, p1 + mid + 1) * res]
            M1 = fetch((self.c1, k_p[1], k_p[0]))[::-1, ::-1]
            M2 = fetch((self.c2, k_p[3], k_p[2]))[::-1, ::-1]
            M3 = fetch((self.c1, k_p[1], k_p[0]), (self.c2, k_p[3], k_p[2]))[::-1, ::-1]
        n1, n = M1.shape[0], M1.shape[0] + M2.shape[0]
        M = np.zeros((n, n))
        M[:n1, :n1], M[n1:, n1:], M[:n1, n1:] = M1, M2, M3
        M = np.triu(M)
        if col:
            M[np.isnan(M)] = 0
        self.fusion_matrix, self.k_p, self.fusion_point = M, k_p, n1

    def extract_bias(self, col):
        """callers.py:240-262."""
        k_p, parts = self.k_p, []
        for c, a, b in ((self.c1, k_p[0], k_p[1]), (self.c2, k_p[2], k_p[3])):
            t = self.clr.bins().fetch((c, min(a, b), max(a, b)))[col].values
            if a > b:
                t = t[::-1]
            ok = np.logical_not((t == 0) | np.isnan(t))
            wv = np.zeros_like(t)
            wv[ok] = 1 / t[ok]
            parts.append(wv)
        self.bias = np.r_[parts[0], parts[1]]

    @staticmethod
    def _stretches(arr, min_seed_len=2, max_gap=1, min_stripe_len=5):
        """callers.py:315-333."""
        arr = np.sort(arr)
        pieces = [p for p in np.split(arr, np.where(np.diff(arr) != 1)[0] + 1) if len(p) >= min_seed_len]
        stripes, seed = [], pieces[0]
        for p in pieces[1:]:
            if p[0] - seed[-1] < max_gap + 2:
                seed = np.r_[seed, p]
            else:
                if seed[-1] - seed[0] + 1 >= min_stripe_len:
                    stripes.append([seed[0], seed[-1] + 1])
                seed = p
        if seed[-1] - seed[0] + 1 >= min_stripe_len:
            stripes.append([seed[0], seed[-1] + 1])
        return stripes

    def _locate(self, pca1, left_most):
        loci = sorted(self._stretches(np.where(pca1 >= 0)[0]) + self._stretches(np.where(pca1 <= 0)[0]))
        return loci[0][1] - 1 if left_most else loci[-1][0]

    def detect_bounds(self):
        """callers.py:265-312."""
        from scipy.ndimage import gaussian_filter
        from sklearn.decomposition import PCA
        n, junc = self.fusion_matrix.shape[0], self.fusion_point
        self.up_i, self.down_i = 0, n - 1
        inter = self.fusion_matrix[:junc, junc:]
        rowmask, colmask = inter.sum(axis=1) != 0, inter.sum(axis=0) != 0
        if rowmask.sum() >= 10 and colmask.sum() >= 10:
            new = inter[rowmask][:, colmask]
            try:
                pca1 = PCA(n_components=3, whiten=True).fit_transform(gaussian_filter(np.corrcoef(new, rowvar=True), sigma=1))[:, 0]
                self.up_i = np.where(rowmask)[0][self._locate(pca1, False)]
            except Exception:       # noqa: BLE001
                self.up_i = 0
            try:
                pca1 = PCA(n_components=3, whiten=True).fit_transform(gaussian_filter(np.corrcoef(new, rowvar=False), sigma=1))[:, 0]
                self.down_i = np.where(colmask)[0][self._locate(pca1, True)] + junc
            except Exception:       # noqa: BLE001
                self.down_i = n - 1

    def allele_slope(self):
        """callers.py:415-478."""
        self.extract_bias(self.col)
        shape, p1L = self.fusion_matrix.shape, self.fusion_point
        valid_bias = (self.bias[:, np.newaxis] * self.bias) > 0
        for up_i, down_i in ((self.up_i, self.down_i), (0, self.bias.size - 1)):
            best = []
            for rows, cols in (((up_i, p1L), (up_i, p1L)), ((p1L, down_i), (p1L, down_i)), ((up_i, p1L), (p1L, down_i))):
                mask = np.zeros(shape, dtype=bool)
                mask[rows[0]:rows[1], cols[0]:cols[1]] = True
                rs, ss = [], []
                for maxdis in (200000, 400000, 500000, 1000000, 1500000, 2000000):
                    r_, s_, _w = linear_regression(extract_xy(self.fusion_matrix, self.res, mask & valid_bias, maxdis=maxdis, mink=3),
                                                   self.expected)
                    rs.append(r_)
                    ss.append(s_)
                rs = np.r_[rs]
                best.append((rs.max(), ss[rs.argmax()]))
            (self.u_g_r, self.u_g_s), (self.d_g_r, self.d_g_s), (self.m_g_r, self.m_g_s) = best
            if self.u_g_r > 0.6 and self.d_g_r > 0.6 and self.m_g_r > 0.6 and self.m_g_s > 0:
                break

    def correct_heterozygosity(self):
        """callers.py:482-502."""
        junc, hcm = self.fusion_point, self.fusion_matrix.copy()
        self.u_s = self.d_s = 0
        if self.u_g_r > 0.6 and self.d_g_r > 0.6:
            slope, ms = max(self.u_g_s, self.d_g_s), min(self.u_g_s, self.d_g_s)
            hcm[:junc, :junc] = hcm[:junc, :junc] / slope
            hcm[junc:, junc:] = hcm[junc:, junc:] / slope
            self.u_s, self.d_s = self.m_g_s / self.u_g_s, self.m_g_s / self.d_g_s
            if self.m_g_r > 0.6 and self.m_g_s / ms > 0.1:
                hcm[self.up_i:junc, junc:self.down_i] = hcm[self.up_i:junc, junc:self.down_i] * min(1 / self.m_g_s, 5 / slope)
        self.hcm = hcm


# ----------------------------------------------------------------------------- assembly.py:278-596
class AssemblyPort:
    def __init__(self, clr, candidate, span=5000000, col='sweight', flexible=True, slopes={}, expected_values=None):
        self.clr = clr
        self.res = clr.binsize
        self.pre = 'chr' if list(clr.chromnames)[0].startswith('chr') else ''
        self.span = span
        self.flexible = flexible
        self.slopes = slopes
        self.balance_type = col if col in ['weight', 'sweight'] else False
        self.chroms = {c: [0, clr.chromsizes[c]] for c in clr.chromnames}
        parse = candidate.split()
        self.sv_list = []
        for p in parse[:-2]:
            _, c1, p1, s1, c2, p2, s2 = p.split(',')
            self.sv_list.append((c1, c2, int(p1), int(p2), s1, s2))
        b1, b2 = parse[-2].split(','), parse[-1].split(',')
        self.bounds = [(b1[0], int(b1[1])), (b2[0], int(b2[1]))]
        if len(self.sv_list) == 1:
            self.tb, self.to = self.single_block(self.sv_list[0], self.bounds[0], self.bounds[1])
        else:
            self.tb, self.to = [], []
            for i, sv in enumerate(self.sv_list):
                if i == 0:
                    iv, dr = self.single_block(sv, left_bound=self.bounds[0])
                elif i == len(self.sv_list) - 1:
                    iv, dr = self.single_block(sv, right_bound=self.bounds[1])
                else:
                    iv, dr = self.single_block(sv)
                self.tb.extend(iv)
                self.to.extend(dr)
        self.expected = expected_values

    def single_block(self, sv, left_bound=None, right_bound=None):
        c1, c2, p1, p2, s1, s2 = sv
        c1, c2 = self.pre + c1, self.pre + c2
        directions = [s1, '-' if s2 == '+' else '+']
        if s1 == '+':
            if self.flexible and left_bound is not None:
                r1 = [c1, left_bound[1], p1]
            else:
                r1 = [c1, max(self.chroms[c1][0], p1 - self.span), p1]
        else:
            if self.flexible and left_bound is not None:
                r1 = [c1, p1, left_bound[1]]
            else:
                r1 = [c1, p1, min(p1 + self.span, self.chroms[c1][1])]
        if s2 == '-':
            if self.flexible and right_bound is not None:
                r2 = [c2, p2, right_bound[1]]
            else:
                r2 = [c2, p2, min(p2 + self.span, self.chroms[c2][1])]
        else:
            if self.flexible and right_bound is not None:
                r2 = [c2, right_bound[1], p2]
            else:
                r2 = [c2, max(self.chroms[c2][0], p2 - self.span), p2]
        return [r1, r2], directions

    def get_bias(self, r):
        col = 'weight' if self.balance_type == False else self.balance_type  # noqa: E712
        bias = self.clr.bins().fetch(r)[col].values
        bias[np.isnan(bias)] = 0
        return bias

    def get_matrix(self, r1, r2):
        M = self.clr.matrix(balance=self.balance_type).fetch(r1, r2)
        if self.balance_type:
            M[np.isnan(M)] = 0
        return M

    def reorganize_matrix(self):
        tb, to = self.tb, self.to
        blocks, orients = [tb[0]], [to[0]]
        for i in range(1, len(tb) - 1, 2):
            assert to[i] == to[i + 1]
            assert tb[i][0] == tb[i + 1][0]
            if to[i] == '+':
                b1, b2 = tb[i][1], tb[i + 1][2]
                assert b1 < b2
                blocks.append([tb[i][0], b1, b2])
                orients.append('+')
            else:
                b1, b2 = tb[i][2], tb[i + 1][1]
                assert b1 > b2
                blocks.append([tb[i][0], b2, b1])
                orients.append('-')
        blocks.append(tb[-1])
        orients.append(to[-1])
        index, Map = [], []
        start = 0
        for r, o in zip(blocks, orients):
            N = self.clr.bins().fetch(tuple(r))['chrom'].size
            index.append((start, start + N))
            coords = [(r[0], r[1] // self.res * self.res + (i - start) * self.res) for i in range(start, start + N)]
            if o == '+':
                Map.extend(list(zip(coords, range(start, start + N))))
            else:
                Map.extend(list(zip(coords[::-1], range(start, start + N))))
            start = index[-1][1]
        forward = dict(Map)
        M = np.zeros((index[-1][1], index[-1][1]))
        bias_arr = np.r_[[]]
        for r1, o1, i1, j1 in zip(blocks, orients, index, range(len(blocks))):
            tmp = self.get_bias(r1)
            bias_arr = np.r_[bias_arr, tmp if o1 == '+' else tmp[::-1]]
            for r2, o2, i2, j2 in zip(blocks, orients, index, range(len(blocks))):
                if j1 > j2:
                    continue
                tmp = self.get_matrix(r1, r2)
                if (o1 == '+') and (o2 == '-'):
                    tmp = tmp[:, ::-1]
                elif (o1 == '-') and (o2 == '+'):
                    tmp = tmp[::-1, :]
                elif (o1 == '-') and (o2 == '-'):
                    tmp = tmp[::-1, ::-1]
                M[i1[0]:i1[1], i2[0]:i2[1]] = tmp
        self.fusion_matrix = np.triu(M)
        self.Map = forward
        self.reverse = {forward[i]: i for i in forward}
        self.index = index
        self.blocks = blocks
        self.bias = bias_arr
        self.chains = {}
        for i, r in enumerate(index):
            for ri in range(r[0], r[1]):
                self.chains[ri] = i

    def correct_heterozygosity(self):
        Bias_M = self.bias[:, np.newaxis] * self.bias
        valid_bias = np.triu(Bias_M > 0)
        hcm = self.fusion_matrix.copy()
        slopes = {}
        nb = len(self.blocks)
        for i1, j1 in zip(self.index, range(nb)):
            for i2, j2 in zip(self.index, range(nb)):
                if j1 > j2:
                    continue
                inter_mask = np.zeros(hcm.shape, dtype=bool)
                inter_mask[i1[0]:i1[1], i2[0]:i2[1]] = True
                valid_pixel = inter_mask & valid_bias
                _r, _s = [], []
                for maxdis in [200000, 400000, 500000, 1000000, 1500000, 2000000]:
                    local_exp = extract_xy(self.fusion_matrix, self.res, valid_pixel, mink=3, maxdis=maxdis)
                    r, s, w = linear_regression(local_exp, self.expected, min_samples=5)
                    _r.append(r)
                    _s.append(s)
                _r = np.r_[_r]
                slopes[(j1, j2)] = [_r.max(), _s[_r.argmax()]]
        for i1, j1 in zip(self.index, range(nb)):
            for i2, j2 in zip(self.index, range(nb)):
                if j1 > j2:
                    continue
                ms = max(slopes[(j1, j1)][1], slopes[(j2, j2)][1])
                blk = (slice(i1[0], i1[1]), slice(i2[0], i2[1]))
                if (j1, j2) in self.slopes:
                    hcm[blk] = hcm[blk] / self.slopes[(j1, j2)]
                elif (slopes[(j1, j2)][0] > 0.6) and (slopes[(j1, j2)][1] > 0) and (ms > 0):
                    if j1 == j2:
                        hcm[blk] = hcm[blk] / slopes[(j1, j2)][1]
                    else:
                        hcm[blk] = hcm[blk] * min(1 / slopes[(j1, j2)][1], 5 / ms)
        x, y = hcm.nonzero()
        hcm[y, x] = hcm[x, y]
        self.slopes_ = slopes
        self.hcm = np.triu(hcm)

    def remove_gaps(self):
        M = self.hcm
        w = 5
        mask = M.sum(axis=0) == 0
        index = list(np.where(mask)[0])
        convert_index = np.where(np.logical_not(mask))[0]
        if w * 2 + 1 > convert_index.size:
            index = []
            convert_index = np.arange(len(M))
        nM = np.delete(np.delete(M, index, 0), index, 1)
        self.gap_free = nM
        self.index_map = dict(zip(np.arange(len(nM)), convert_index))

    def index_to_coordinates(self, loop_list):
        lines = {}
        for i, j, v in loop_list:
            x, y = self.index_map[i], self.index_map[j]
            gdis = (y - x) * self.res
            l1, l2 = self.reverse[x], self.reverse[y]
            if l1 > l2:
                l1, l2 = l2, l1
            key = (l1[0], l1[1], l1[1] + self.res, l2[0], l2[1], l2[1] + self.res)
            lines[key] = [gdis, 0 if self.chains[x] == self.chains[y] else 1]
        return lines


# ----------------------------------------------------------------------------- callers.py:525-831
class PeakachuPort:
    def __init__(self, matrix, upper=4000000, res=10000):
        self.w = 7
        upper = upper // res
        R, C = matrix.nonzero()
        data = matrix[R, C]
        ok = np.isfinite(data) & (C - R > (-2 * self.w)) & (C - R < (upper + 2 * self.w))
        self.M = sparse.csr_matrix((data[ok], (R[ok], C[ok])), shape=matrix.shape)
        self.get_candidates(4, upper)
        self.r = res

    def get_candidates(self, lower, upper):
        from sklearn.isotonic import IsotonicRegression
        exp_obs, idx_obs = [], []
        for i in range(lower, upper + 1):
            diag = self.M.diagonal(i)
            if diag.size > 5:
                exp_obs.append(diag.mean())
                idx_obs.append(i)
        IR = IsotonicRegression(increasing=False, out_of_bounds='clip')
        IR.fit(np.r_[idx_obs], np.r_[exp_obs])
        d = np.arange(lower, upper + 1)
        expected = dict(zip(d, IR.predict(d)))
        x_arr = np.array([], dtype=int)
        y_arr = np.array([], dtype=int)
        idx = np.arange(self.M.shape[0])
        for i in range(lower, upper + 1):
            diag = self.M.diagonal(i)
            e = expected[i]
            if (diag.size > 1) and (e > 0):
                mask = (diag / e) > 0
                x_arr = np.r_[x_arr, idx[:-i][mask]]
                y_arr = np.r_[y_arr, idx[i:][mask]]
        self.ridx, self.cidx = x_arr, y_arr

    def load_models(self, path):
        import joblib
        self.model = joblib.load(path)
        self.w = int((np.sqrt(self.model.feature_importances_.size) - 1) / 2)

    def load_expected(self, expected_values):
        self.exp_arr = np.r_[[expected_values[i] for i in sorted(expected_values)]]

    def getwindow(self, coords, w):
        if len(coords) < 2:
            return [], []
        coords = np.r_[coords]
        xi, yi = coords[:, 0], coords[:, 1]
        mask = (xi - w >= 0) & (yi + w + 1 <= self.M.shape[0]) & (yi - xi - 2 >= w)
        xi, yi = xi[mask], yi[mask]
        if xi.size < 2:
            return [], []
        seed = np.arange(-w, w + 1)
        delta = np.tile(seed, (seed.size, 1))
        xxx = xi.reshape((xi.size, 1, 1)) + delta.T
        yyy = yi.reshape((yi.size, 1, 1)) + delta
        v = np.array(self.M[xxx.ravel(), yyy.ravel()]).ravel()
        vvv = v.reshape((xi.size, seed.size, seed.size))
        windows, clist = distance_normalize(vvv, self.exp_arr, xi, yi, w)
        fea = []
        for arr in windows:
            fea.append(image_normalize(gaussian_filter(arr, sigma=1, order=0)).ravel())
        return np.r_[fea], np.r_[clist]

    def predict(self, thre, no_pool=False, min_count=2, index_map=None, chains=None, capture=None):
        coords = [(r, c) for r, c in zip(self.ridx, self.cidx)]
        fea, clist = self.getwindow(coords, self.w)
        if capture is not None:
            capture["n_scored"] = len(fea)
        if len(fea) < 2:
            return []
        probas = self.model.predict_proba(fea)[:, 1]
        if capture is not None:
            capture["clist"], capture["probas"] = clist, probas
        pM = sparse.csr_matrix((probas, (clist[:, 0], clist[:, 1])), shape=self.M.shape)
        keep = probas >= thre
        ri, ci = clist[:, 0][keep], clist[:, 1][keep]
        Donuts = {(i, j): self.M[i, j] for i, j in zip(ri, ci)}
        tmp = Donuts if no_pool else self.local_clustering(Donuts, min_count=min_count, index_map=index_map, chains=chains)
        return list({(i, j, pM[i, j]) for i, j in tmp})

    def local_clustering(self, Donuts, min_count=2, r=15000, index_map=None, chains=None):
        byregion = {}
        if chains is None:
            byregion[(0, 0)] = Donuts
        else:
            for i, j in Donuts:
                x, y = (i, j) if index_map is None else (index_map[i], index_map[j])
                byregion.setdefault((chains[x], chains[y]), {})[(i, j)] = Donuts[(i, j)]
        coords = set()
        for k in byregion:
            coords.update(set(self._local_clustering(byregion[k], min_count=min_count, r=r)))
        return coords

    def _local_clustering(self, Donuts, min_count=1, r=15000):
        final_list = []
        x = np.r_[[i[0] for i in Donuts]]
        y = np.r_[[i[1] for i in Donuts]]
        if x.size == 0:
            return final_list
        x_anchors = self.find_anchors(x, min_count=min_count, min_dis=r)
        y_anchors = self.find_anchors(y, min_count=min_count, min_dis=r)
        r = max(r // self.r, 2)
        visited = set()
        lookup = set(zip(x, y))
        for x_a in x_anchors:
            for y_a in y_anchors:
                sort_list = []
                for i in range(x_a[1], x_a[2] + 1):
                    for j in range(y_a[1], y_a[2] + 1):
                        if (i, j) in lookup:
                            sort_list.append((Donuts[(i, j)], (i, j)))
                sort_list.sort(reverse=True)
                self._cluster_core(sort_list, r, visited, final_list)
        sort_list = []
        for i, j in zip(x, y):
            if (i, j) not in visited:
                sort_list.append((Donuts[(i, j)], (i, j)))
        sort_list.sort(reverse=True)
        self._cluster_core(sort_list, r, visited, final_list)
        for i, j in zip(x, y):
            if (i, j) not in visited:
                final_list.append((int(i), int(j)))
        return list(set(final_list))

    def find_anchors(self, pos, min_count=3, min_dis=15000, wlen=40000):
        from scipy.signal import find_peaks, peak_widths
        min_dis = max(min_dis // self.r, 2)
        wlen = min(wlen // self.r, 20) if self.r <= 10000 else 4
        count = Counter(pos)
        refidx = range(min(count), max(count) + 1)
        signal = np.r_[[count[i] for i in refidx]]
        summits = find_peaks(signal, height=min_count, distance=min_dis)[0]
        ordered = sorted(((signal[i], i) for i in summits), reverse=True)
        peaks, records = set(), {}
        for _, i in ordered:
            tmp = peak_widths(signal, [i], rel_height=1, wlen=wlen)[2:4]
            li, ri = int(np.round(tmp[0][0])), int(np.round(tmp[1][0]))
            lb, rb = refidx[li], refidx[ri]
            if not len(peaks):
                peaks.add((refidx[i], lb, rb))
                for b in range(lb, rb + 1):
                    records[b] = (refidx[i], lb, rb)
            else:
                for b in range(lb, rb + 1):
                    if b in records:
                        m_lb, m_rb = min(lb, records[b][1]), max(rb, records[b][2])
                        summit = records[b][0]
                        peaks.remove(records[b])
                        break
                else:
                    m_lb, m_rb, summit = lb, rb, refidx[i]
                peaks.add((summit, m_lb, m_rb))
                for b in range(m_lb, m_rb + 1):
                    records[b] = (summit, m_lb, m_rb)
        return peaks

    def _cluster_core(self, sort_list, r, visited, final_list):
        from scipy.spatial.distance import euclidean
        from sklearn.cluster import dbscan
        pos = np.r_[[i[1] for i in sort_list]]
        if len(pos) >= 2:
            _, labels = dbscan(pos, eps=r, min_samples=2)
            pool = set()
            for i, p in enumerate(sort_list):
                if p[1] in pool:
                    continue
                c = labels[i]
                if c == -1:
                    continue
                sub = pos[labels == c]
                cen, rad = p[1], r
                Local = [p[1]]
                ini = -1
                while len(sub):
                    out = []
                    for q in sub:
                        if tuple(q) in pool:
                            continue
                        if euclidean(q, cen) <= rad:
                            Local.append(tuple(q))
                        else:
                            out.append(tuple(q))
                    if len(out) == ini:
                        break
                    ini = len(out)
                    tmp = np.r_[Local]
                    cen = tuple(tmp.mean(axis=0).round().astype(int))
                    rad = np.int64(np.round(max([euclidean(cen, q) for q in Local]))) + r
                    sub = np.r_[out]
                for q in Local:
                    pool.add(q)
                final_list.append((int(p[1][0]), int(p[1][1])))
            visited.update(pool)
        elif len(pos) == 1:
            final_list.append((int(pos[0][0]), int(pos[0][1])))


# ----------------------------------------------------------------------------- scripts/neoloop-caller:185-218
def pipeline(clr, index, assemblies, rsize, balance, prob, mmp, nopool, models_by_res, expected, capture=None):
    fu = AssemblyPort(clr, assemblies[index], span=rsize, col=balance, expected_values=expected, flexible=False)
    fu.reorganize_matrix()
    if len(fu.Map) != fu.fusion_matrix.shape[0]:
        return None
    fu.correct_heterozygosity()
    fu.remove_gaps()
    core = PeakachuPort(fu.gap_free, upper=400 * fu.res, res=fu.res)
    core.load_expected(expected)
    core.load_models(models_by_res[fu.res])
    loop_index = core.predict(thre=prob, no_pool=nopool, min_count=mmp, index_map=fu.index_map, chains=fu.chains,
                              capture=capture)
    loops = fu.index_to_coordinates(loop_index)
    fu2 = AssemblyPort(clr, assemblies[index], span=rsize, col=balance, expected_values=expected)
    fu2.reorganize_matrix()
    final = {}
    for k in loops:
        if ((k[0], k[1]) in fu2.Map) and ((k[3], k[4]) in fu2.Map):
            final[k] = (index, loops[k][0], loops[k][1])
    return final


def combine_annotations(byres):
    reslist = sorted(byres)
    pool = byres[reslist[0]]
    visited = [reslist[0]]
    for res in reslist[1:]:
        cur = byres[res]
        for l in cur:
            for pr in visited:
                count = 0
                for i in range(l[1] // pr, l[2] // pr + 1):
                    for j in range(l[4] // pr, l[5] // pr + 1):
                        tmp = (l[0], i * pr, i * pr + pr, l[3], j * pr, j * pr + pr)
                        if tmp in pool:
                            count += 1
                            pool[tmp].extend(cur[l])
                if count == 0:
                    pool[l] = cur[l]
        visited.append(res)
    loop_list = []
    for k in sorted(pool):
        labels = [','.join([t[0], str(t[1]), str(t[2])]) for t in set(pool[k])]
        loop_list.append(list(map(str, k)) + [','.join(labels)])
    return loop_list
