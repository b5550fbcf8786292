/*
 * nlf_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C CPU restatement of the arithmetic on the neoloop-caller scoring path,
 * including the third-party arithmetic the reference reaches through numpy, scipy and
 * scikit-learn.  Every function cites the reference location it follows
 * (paths relative to /root/reference, or the third-party module).
 *
 * Build: oracle/build.py  (gcc -O2 -ffp-contract=off: NO fused multiply-add, the
 * reference's numpy/scipy arithmetic rounds after every operation).
 *
 * Parity pinning: checked against tests/golden/ (npz files), which were produced by the real
 * reference in the build container (oracle/gen_golden.py), and -- in that container --
 * against live calls into the reference (tests/test_oracle_vs_reference.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NLFO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * numpy pairwise summation: numpy/_core/src/umath/loops_utils.h.src  @TYPE@_pairwise_sum
 * (what ndarray.sum()/mean() run on a contiguous or strided 1-D float64 buffer).
 * ---------------------------------------------------------------------------------- */
NLFO_API double nlfo_pairwise_sum(const double *a, long n, long stride)
{
    if (n < 8) {
        double res = -0.0;
        for (long i = 0; i < n; i++) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j * stride];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[(i + j) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return nlfo_pairwise_sum(a, n2, stride) + nlfo_pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

/* ------------------------------------------------------------------------------------
 * Fusion._extract_xy (neoloop/callers.py:348-367) on the valid_pixel mask built by
 * complexSV.correct_heterozygosity (neoloop/assembly.py:497-516) for ONE block pair:
 * rows [lo1,hi1) x cols [lo2,hi2) AND triu(bias_i*bias_j > 0).
 * For diagonal i = 2 .. min(n-2, maxdiag): values M[x, x+i] in row order; if count>=mink
 * the numpy mean; kept only if mean > 0.  mean_out[i] = NaN when not kept.
 * The six maxdis values of the reference are nested prefixes of maxdiag = 2 Mb / res.
 * ---------------------------------------------------------------------------------- */
NLFO_API void nlfo_block_diag_means(const double *M, int n, const double *bias, int lo1, int hi1,
                                    int lo2, int hi2, int maxdiag, int mink, double *mean_out,
                                    int *cnt_out)
{
    double *buf = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i <= maxdiag; i++) { mean_out[i] = NAN; cnt_out[i] = 0; }
    for (int i = 2; i < n - 1; i++) {
        if (i > maxdiag) break;
        int cnt = 0;
        for (int x = lo1; x < hi1; x++) {
            int y = x + i;
            if (y < lo2 || y >= hi2 || y >= n) continue;
            if (bias[x] * bias[y] > 0) buf[cnt++] = M[(size_t)x * n + y];
        }
        cnt_out[i] = cnt;
        if (cnt >= mink) {
            double m = nlfo_pairwise_sum(buf, cnt, 1) / (double)cnt;
            if (m > 0) mean_out[i] = m;
        }
    }
    free(buf);
}

/* ------------------------------------------------------------------------------------
 * Per-block-pair rescaling, complexSV.correct_heterozygosity (neoloop/assembly.py:536-549).
 * op[j1*nblk+j2] : 0 = untouched, 1 = divide by val, 2 = multiply by val  (j1 <= j2).
 * ---------------------------------------------------------------------------------- */
NLFO_API void nlfo_scale_blocks(double *hcm, int n, const int *blk_lo, const int *blk_hi, int nblk,
                                const int *op, const double *val)
{
    for (int j1 = 0; j1 < nblk; j1++)
        for (int j2 = j1; j2 < nblk; j2++) {
            int o = op[j1 * nblk + j2];
            double v = val[j1 * nblk + j2];
            if (!o) continue;
            for (int r = blk_lo[j1]; r < blk_hi[j1]; r++)
                for (int c = blk_lo[j2]; c < blk_hi[j2]; c++) {
                    size_t k = (size_t)r * n + c;
                    hcm[k] = (o == 1) ? hcm[k] / v : hcm[k] * v;
                }
        }
}

/* ------------------------------------------------------------------------------------
 * Fusion.remove_gaps (neoloop/callers.py:504-522): mask = M.sum(axis=0) == 0.
 * numpy reduces axis 0 of a C-contiguous matrix row by row: out[j] = (..(M[0,j]+M[1,j])+..).
 * Returns the number of kept columns; kept[] = gap-free -> original index (index_map).
 * ---------------------------------------------------------------------------------- */
NLFO_API int nlfo_remove_gaps_index(const double *hcm, int n, int *kept)
{
    int w = 5, dim = 0;
    for (int j = 0; j < n; j++) {
        double s = hcm[j];
        for (int r = 1; r < n; r++) s += hcm[(size_t)r * n + j];
        if (!(s == 0)) kept[dim++] = j;
    }
    if (w * 2 + 1 > dim) {
        for (int j = 0; j < n; j++) kept[j] = j;
        dim = n;
    }
    return dim;
}

/* ------------------------------------------------------------------------------------
 * scipy.optimize.isotonic_regression -> scipy/optimize/_pava/pava_pybind.cpp "pava"
 * (Busing 2022, algorithm 1, scipy's >= variant), increasing order, in place.
 * Returns the number of blocks.  r must hold n+1 entries.
 * ---------------------------------------------------------------------------------- */
static long pava_increasing(double *x, double *w, long *r, long n)
{
    r[0] = 0; r[1] = 1;
    long b = 0;
    double xb_prev = x[b], wb_prev = w[b];
    for (long i = 1; i < n; ++i) {
        b++;
        double xb = x[i], wb = w[i], sb = 0;
        if (xb_prev >= xb) {
            b--;
            sb = wb_prev * xb_prev + wb * xb;
            wb += wb_prev;
            xb = sb / wb;
            while (i < n - 1 && xb >= x[i + 1]) {
                i++;
                sb += w[i] * x[i];
                wb += w[i];
                xb = sb / wb;
            }
            while (b > 0 && x[b - 1] >= xb) {
                b--;
                sb += w[b] * x[b];
                wb += w[b];
                xb = sb / wb;
            }
        }
        x[b] = xb_prev = xb;
        w[b] = wb_prev = wb;
        r[b + 1] = i + 1;
    }
    long f = n - 1;
    for (long k = b; k >= 0; --k) {
        long t = r[k];
        double xk = x[k];
        for (long i = f; i >= t; --i) x[i] = xk;
        f = t - 1;
    }
    return b + 1;
}

/* sklearn IsotonicRegression(increasing=False, out_of_bounds='clip').fit(X, y).predict(T)
 * for strictly increasing integer X = xs[0..m) with unit weights and integer queries:
 * sklearn/isotonic.py _build_y (decreasing = reverse, PAVA, reverse back; trim_duplicates),
 * predict = clip to [X_min, X_max] then scipy interp1d(kind='linear') == numpy.interp:
 * exact knot value at a knot, and between two knots only flat stretches were trimmed, so
 * the interpolant equals the fitted value.  Hence pred(t) = yhat[clip(t)] exactly. */
NLFO_API void nlfo_isotonic_decreasing_fit(const double *y, long m, double *yhat)
{
    double *x = (double *)malloc(sizeof(double) * (size_t)m);
    double *w = (double *)malloc(sizeof(double) * (size_t)m);
    long *r = (long *)malloc(sizeof(long) * (size_t)(m + 1));
    for (long i = 0; i < m; i++) { x[i] = y[m - 1 - i]; w[i] = 1.0; }
    if (m > 1) pava_increasing(x, w, r, m);
    for (long i = 0; i < m; i++) yhat[i] = x[m - 1 - i];
    free(x); free(w); free(r);
}

/* value of the Peakachu CSR matrix M at (r, c): neoloop/callers.py:531-535 keeps finite
 * non-zeros with -2w < c-r < upper+2w (w = 7 at construction time). */
static inline double band_value(const double *G, int n, int r, int c, int upper)
{
    int d = c - r;
    if (r < 0 || c < 0 || r >= n || c >= n) return 0.0;
    if (!(d > -14 && d < upper + 14)) return 0.0;
    double v = G[(size_t)r * n + c];
    return isfinite(v) ? v : 0.0;
}

/* ------------------------------------------------------------------------------------
 * Peakachu.get_candidates(lower, upper) (neoloop/callers.py:540-570).
 * e_out[d] (d = 0..upper; only lower..upper filled) = isotonic expected.
 * Returns the candidate count (diagonal-major, row-ascending), or -1 when the isotonic fit
 * has no sample (sklearn raises ValueError there).
 * ---------------------------------------------------------------------------------- */
NLFO_API long nlfo_candidates(const double *G, int n, int lower, int upper, int *ridx, int *cidx,
                              double *e_out)
{
    int nd = upper - lower + 1;
    double *means = (double *)malloc(sizeof(double) * (size_t)nd);
    double *fit = (double *)malloc(sizeof(double) * (size_t)nd);
    double *diag = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    long m = 0;   /* diagonals lower..lower+m-1 have size > 5 (sizes decrease with i) */
    for (int i = lower; i <= upper; i++) {
        int len = n - i;
        if (len > 5) {
            for (int k = 0; k < len; k++) diag[k] = band_value(G, n, k, k + i, upper);
            means[m++] = nlfo_pairwise_sum(diag, len, 1) / (double)len;
        }
    }
    if (m == 0) { free(means); free(fit); free(diag); return -1; }
    nlfo_isotonic_decreasing_fit(means, m, fit);
    long nc = 0;
    for (int i = lower; i <= upper; i++) {
        long k = i - lower;
        if (k > m - 1) k = m - 1;            /* out_of_bounds='clip' */
        double e = fit[k];
        e_out[i] = e;
        int len = n - i;
        if (len > 1 && e > 0) {
            for (int r = 0; r < len; r++) {
                double v = band_value(G, n, r, r + i, upper);
                if (v / e > 0) {
                    if (ridx) { ridx[nc] = r; cidx[nc] = r + i; }
                    nc++;
                }
            }
        }
    }
    free(means); free(fit); free(diag);
    return nc;
}

/* scipy.ndimage.correlate1d with a symmetric kernel and mode='reflect'
 * (scipy/ndimage/src/ni_filters.c NI_Correlate1D, symmetric branch): centre tap first,
 * then pairs from the outermost to the innermost, (left + right) * weight, no FMA. */
static inline int reflect_idx(int i, int n)
{
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i;
}

static void gaussian_2d(const double *in, double *out, int S, const double *gw, int radius)
{
    double tmp[32 * 32];
    /* axis 0 */
    for (int a = 0; a < S; a++)
        for (int b = 0; b < S; b++) {
            double v = in[a * S + b] * gw[radius];
            for (int j = -radius; j < 0; j++)
                v += (in[reflect_idx(a + j, S) * S + b] + in[reflect_idx(a - j, S) * S + b]) * gw[radius + j];
            tmp[a * S + b] = v;
        }
    /* axis 1 */
    for (int a = 0; a < S; a++)
        for (int b = 0; b < S; b++) {
            double v = tmp[a * S + b] * gw[radius];
            for (int j = -radius; j < 0; j++)
                v += (tmp[a * S + reflect_idx(b + j, S)] + tmp[a * S + reflect_idx(b - j, S)]) * gw[radius + j];
            out[a * S + b] = v;
        }
}

/* ------------------------------------------------------------------------------------
 * Peakachu.getwindow (neoloop/callers.py:583-612) + util.distance_normalize
 * (neoloop/util.py:499-524) + distance_normaize_core (util.py:470-491) +
 * scipy.ndimage.gaussian_filter(sigma=1) + util.image_normalize (util.py:493-497).
 * Returns the number of feature rows written (fea: rows x S*S fp64, clist: rows x 2).
 * Like the reference, fewer than 2 coordinates (before or after the border mask) give 0.
 * ---------------------------------------------------------------------------------- */
NLFO_API long nlfo_getwindow(const double *G, int n, int upper, const int *xi, const int *yi, long nc,
                             int w, const double *expv, int nexp, const double *gw, int radius,
                             double *fea, int *clist)
{
    int S = 2 * w + 1, F = S * S;
    if (nc < 2) return 0;
    long nmask = 0;
    for (long k = 0; k < nc; k++) {
        int x = xi[k], y = yi[k];
        if (x - w >= 0 && y + w + 1 <= n && y - x - 2 >= w) nmask++;
    }
    if (nmask < 2) return 0;
    /* pass 1: decide which windows survive (sequential so that the output order is the input order) */
    long *sel = (long *)malloc(sizeof(long) * (size_t)nmask);
    long nsel = 0;
    for (long k = 0; k < nc; k++) {
        int x = xi[k], y = yi[k];
        if (!(x - w >= 0 && y + w + 1 <= n && y - x - 2 >= w)) continue;
        double win[32 * 32], blk[16 * 16];
        int nz = 0;
        for (int a = 0; a < S; a++)
            for (int b = 0; b < S; b++) {
                double v = band_value(G, n, x - w + a, y - w + b, upper);
                win[a * S + b] = v;
                nz += (v != 0);
            }
        if ((double)nz < (double)F * 0.1) continue;
        for (int a = 0; a < w; a++)
            for (int b = 0; b < w; b++) blk[a * w + b] = win[a * S + b];
        double ll_mean = nlfo_pairwise_sum(blk, w * w, 1) / (double)(w * w);
        if (ll_mean > 0) {
            double p2ll = win[w * S + w] / ll_mean;
            if (p2ll > 0.1) sel[nsel++] = k;
        }
    }
    /* pass 2: O/E, gaussian, min-max */
#pragma omp parallel for schedule(static)
    for (long s = 0; s < nsel; s++) {
        long k = sel[s];
        int x = xi[k], y = yi[k];
        double win[32 * 32], g[32 * 32];
        int maxdis = (y + w) - (x - w);
        if (maxdis < 0) maxdis = -maxdis;
        int alt = (y - w) - (x + w);
        if (alt < 0) alt = -alt;
        if (alt > maxdis) maxdis = alt;
        for (int a = 0; a < S; a++)
            for (int b = 0; b < S; b++) {
                double v = band_value(G, n, x - w + a, y - w + b, upper);
                if (maxdis < nexp) {
                    int D = (y - w + b) - (x - w + a);
                    if (D < 0) D = -D;
                    v = v / expv[D];
                }
                win[a * S + b] = v;
            }
        gaussian_2d(win, g, S, gw, radius);
        /* numpy min()/max() propagate NaN */
        double mn = g[0], mx = g[0];
        int has_nan = 0;
        for (int t = 0; t < F; t++) {
            if (g[t] != g[t]) has_nan = 1;
            if (g[t] < mn) mn = g[t];
            if (g[t] > mx) mx = g[t];
        }
        if (has_nan) { mn = NAN; mx = NAN; }
        double den = mx - mn;
        for (int t = 0; t < F; t++) fea[(size_t)s * F + t] = (g[t] - mn) / den;
        clist[2 * s] = x;
        clist[2 * s + 1] = y;
    }
    free(sel);
    return nsel;
}

/* ------------------------------------------------------------------------------------
 * RandomForestClassifier.predict_proba(X)[:, 1] (sklearn/ensemble/_forest.py: X cast to
 * float32; all_proba zeros; += tree proba in estimator order; /= n_estimators) with
 * Tree._apply_dense (sklearn/tree/_tree.pyx): float32 feature vs float64 threshold,
 * NaN routed by missing_go_to_left.  Arrays are the raw sklearn tree arrays, concatenated;
 * tree t owns nodes [tree_off[t], tree_off[t+1]); child indices are tree-local.
 * leaf_p1[node] = tree_.value[node, 0, 1].
 * ---------------------------------------------------------------------------------- */
NLFO_API void nlfo_forest_predict(int n_trees, const long *tree_off, const long *left, const long *right,
                                  const long *feature, const double *threshold,
                                  const unsigned char *missing_left, const double *leaf_p1,
                                  const float *X, long N, int F, double *proba)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; i++) {
        const float *x = X + (size_t)i * F;
        double acc = 0.0;
        for (int t = 0; t < n_trees; t++) {
            long base = tree_off[t], node = 0;
            while (left[base + node] != -1) {
                float v = x[feature[base + node]];
                if (v != v)
                    node = missing_left[base + node] ? left[base + node] : right[base + node];
                else if ((double)v <= threshold[base + node])
                    node = left[base + node];
                else
                    node = right[base + node];
            }
            acc += leaf_p1[base + node];
        }
        proba[i] = acc / (double)n_trees;
    }
}

/* ====================================================================================
 * Pooling: Peakachu._local_clustering / find_anchors / _cluster_core
 * (neoloop/callers.py:699-831) for ONE (chain, chain) bucket.
 * ==================================================================================== */

/* scipy.signal._peak_finding_utils._local_maxima_1d: strict rise, plateau midpoint */
static int local_maxima_1d(const double *x, int n, int *mid, int *ledge, int *redge)
{
    int m = 0, i = 1, i_max = n - 1;
    while (i < i_max) {
        if (x[i - 1] < x[i]) {
            int ahead = i + 1;
            while (ahead < i_max && x[ahead] == x[i]) ahead++;
            if (x[ahead] < x[i]) {
                ledge[m] = i;
                redge[m] = ahead - 1;
                mid[m] = (ledge[m] + redge[m]) / 2;
                m++;
                i = ahead;
            }
        }
        i++;
    }
    return m;
}

/* scipy.signal._peak_finding_utils._select_by_peak_distance.  Priority order = ascending
 * argsort of the heights; scipy uses numpy's default (unstable, SIMD-dispatched) argsort, so
 * the order among EQUAL heights is implementation-defined in the reference (it differs
 * between numpy's AVX-512, AVX2 and scalar sorts).  It only matters when two equal summits
 * are closer than `distance`, which is impossible for distance <= 2 (10 kb and coarser).
 * `ext_order` (may be NULL) is the permutation numpy.argsort returned for `prio` on the
 * machine running the comparison; without it the stable order is used. */
static void select_by_distance(const int *peaks, const double *prio, int m, int distance, unsigned char *keep,
                               const int *ext_order)
{
    int *order = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    for (int i = 0; i < m; i++) { order[i] = i; keep[i] = 1; }
    if (ext_order) {
        for (int i = 0; i < m; i++) order[i] = ext_order[i];
    } else {
        /* stable insertion sort, ascending priority */
        for (int i = 1; i < m; i++) {
            int o = order[i], j = i - 1;
            while (j >= 0 && prio[order[j]] > prio[o]) { order[j + 1] = order[j]; j--; }
            order[j + 1] = o;
        }
    }
    for (int i = m - 1; i >= 0; i--) {
        int j = order[i];
        if (!keep[j]) continue;
        int k = j - 1;
        while (k >= 0 && peaks[j] - peaks[k] < distance) { keep[k] = 0; k--; }
        k = j + 1;
        while (k < m && peaks[k] - peaks[j] < distance) { keep[k] = 0; k++; }
    }
    free(order);
}

/* scipy.signal.peak_prominences + peak_widths(rel_height=1.0, wlen) for one peak.
 * (_peak_finding_utils._peak_prominences / _peak_widths) */
static void peak_width_bounds(const double *x, int n, int peak, int wlen_in, double *left_ip, double *right_ip)
{
    int i_min = 0, i_max = n - 1;
    int wlen = wlen_in;
    if (wlen >= 2) {
        /* _arg_wlen_as_expected: ceil + round up to odd;  window = peak +- wlen/2 */
        if (wlen % 2 == 0) wlen += 1;
        int half = wlen / 2;
        i_min = peak - half > 0 ? peak - half : 0;
        i_max = peak + half < n - 1 ? peak + half : n - 1;
    }
    int i = peak, left_base = peak, right_base = peak;
    double left_min = x[peak], right_min = x[peak];
    while (i_min <= i && x[i] <= x[peak]) {
        if (x[i] < left_min) { left_min = x[i]; left_base = i; }
        i--;
    }
    i = peak;
    while (i <= i_max && x[i] <= x[peak]) {
        if (x[i] < right_min) { right_min = x[i]; right_base = i; }
        i++;
    }
    double prominence = x[peak] - (left_min > right_min ? left_min : right_min);
    double height = x[peak] - prominence * 1.0;
    /* _peak_widths */
    i = peak;
    while (left_base < i && height < x[i]) i--;
    double lip = (double)i;
    if (x[i] < height) lip += (height - x[i]) / (x[i + 1] - x[i]);
    i = peak;
    while (i < right_base && height < x[i]) i++;
    double rip = (double)i;
    if (x[i] < height) rip -= (height - x[i]) / (x[i - 1] - x[i]);
    *left_ip = lip;
    *right_ip = rip;
}

typedef struct { int summit, lb, rb; } anchor_t;

/* Peakachu.find_anchors (neoloop/callers.py:740-784). pos: coordinates (with repeats).
 * Returns the number of anchors written. */
static int find_anchors(const int *pos, int npos, int res, int min_count, int r_bp, anchor_t *out,
                        const int *ext_order, double *prio_out, int *n_prio)
{
    int min_dis = r_bp / res; if (min_dis < 2) min_dis = 2;
    int wlen;
    if (res <= 10000) { wlen = 40000 / res; if (wlen > 20) wlen = 20; } else wlen = 4;
    int lo = pos[0], hi = pos[0];
    for (int i = 1; i < npos; i++) { if (pos[i] < lo) lo = pos[i]; if (pos[i] > hi) hi = pos[i]; }
    int n = hi - lo + 1;
    double *sig = (double *)calloc((size_t)n, sizeof(double));
    for (int i = 0; i < npos; i++) sig[pos[i] - lo] += 1.0;
    int *mid = (int *)malloc(sizeof(int) * (size_t)n), *le = (int *)malloc(sizeof(int) * (size_t)n),
        *re = (int *)malloc(sizeof(int) * (size_t)n);
    int m = local_maxima_1d(sig, n, mid, le, re);
    /* height filter */
    int m2 = 0;
    for (int i = 0; i < m; i++) if (sig[mid[i]] >= (double)min_count) mid[m2++] = mid[i];
    m = m2;
    unsigned char *keep = (unsigned char *)malloc((size_t)(m > 0 ? m : 1));
    double *prio = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    for (int i = 0; i < m; i++) prio[i] = sig[mid[i]];
    if (prio_out) {          /* query mode: only report the height-filtered peak priorities */
        for (int i = 0; i < m; i++) prio_out[i] = prio[i];
        *n_prio = m;
        free(sig); free(mid); free(le); free(re); free(keep); free(prio);
        return 0;
    }
    select_by_distance(mid, prio, m, min_dis, keep, ext_order);
    m2 = 0;
    for (int i = 0; i < m; i++) if (keep[i]) mid[m2++] = mid[i];
    m = m2;
    /* sorted_summits.sort(reverse=True) on (count, index) */
    for (int i = 1; i < m; i++) {
        int p = mid[i], j = i - 1;
        while (j >= 0 && (sig[mid[j]] < sig[p] || (sig[mid[j]] == sig[p] && mid[j] < p))) { mid[j + 1] = mid[j]; j--; }
        mid[j + 1] = p;
    }
    /* records[b] = index of the anchor covering bin b (refidx coordinates), -1 if none */
    int *records = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) records[i] = -1;
    int na = 0;            /* anchors; removed ones get lb = INT_MIN marker */
    unsigned char *alive = (unsigned char *)malloc((size_t)(m > 0 ? m : 1));
    for (int s = 0; s < m; s++) {
        int pk = mid[s];
        double lip, rip;
        peak_width_bounds(sig, n, pk, wlen, &lip, &rip);
        int li = (int)nearbyint(lip), ri = (int)nearbyint(rip);   /* np.round: half to even */
        int summit = pk, m_lb = li, m_rb = ri;
        if (na > 0) {
            for (int b = li; b <= ri; b++) {
                if (records[b] >= 0) {
                    int q = records[b];
                    m_lb = li < out[q].lb ? li : out[q].lb;
                    m_rb = ri > out[q].rb ? ri : out[q].rb;
                    summit = out[q].summit;
                    alive[q] = 0;
                    break;
                }
            }
        }
        /* set semantics: an identical (summit, lb, rb) tuple may already be present */
        int idx = -1;
        for (int q = 0; q < na; q++)
            if (alive[q] && out[q].summit == summit && out[q].lb == m_lb && out[q].rb == m_rb) { idx = q; break; }
        if (idx < 0) {
            idx = na++;
            out[idx].summit = summit; out[idx].lb = m_lb; out[idx].rb = m_rb; alive[idx] = 1;
        }
        for (int b = m_lb; b <= m_rb; b++) records[b] = idx;
    }
    int k = 0;
    for (int q = 0; q < na; q++)
        if (alive[q]) { out[k].summit = out[q].summit + lo; out[k].lb = out[q].lb + lo; out[k].rb = out[q].rb + lo; k++; }
    free(sig); free(mid); free(le); free(re); free(keep); free(prio); free(records); free(alive);
    return k;
}

typedef struct { double v; int i, j, id; } spt_t;

static int spt_desc(const void *a, const void *b)
{
    const spt_t *p = (const spt_t *)a, *q = (const spt_t *)b;
    if (p->v != q->v) return p->v > q->v ? -1 : 1;
    if (p->i != q->i) return p->i > q->i ? -1 : 1;
    if (p->j != q->j) return p->j > q->j ? -1 : 1;
    return 0;
}

/* Peakachu._cluster_core (neoloop/callers.py:786-831) on a descending-sorted list.
 * visited[id] / is_final[id] are indexed by the point id in the bucket. */
static void cluster_core(const spt_t *sl, int m, int r, unsigned char *visited, unsigned char *is_final)
{
    if (m == 1) { is_final[sl[0].id] = 1; return; }
    if (m < 2) return;
    /* sklearn.cluster.dbscan(eps=r, min_samples=2): every point with another point within
     * eps is a core point, so labels are the connected components of the eps-graph and
     * isolated points are noise (-1). */
    int *label = (int *)malloc(sizeof(int) * (size_t)m);
    int *stack = (int *)malloc(sizeof(int) * (size_t)m);
    for (int a = 0; a < m; a++) label[a] = -1;
    long r2 = (long)r * r;
    int ncomp = 0;
    for (int a = 0; a < m; a++) {
        if (label[a] != -1) continue;
        int has_nb = 0;
        for (int b = 0; b < m && !has_nb; b++)
            if (b != a) {
                long dx = sl[a].i - sl[b].i, dy = sl[a].j - sl[b].j;
                if (dx * dx + dy * dy <= r2) has_nb = 1;
            }
        if (!has_nb) continue;
        int top = 0;
        stack[top++] = a; label[a] = ncomp;
        while (top) {
            int u = stack[--top];
            for (int b = 0; b < m; b++)
                if (label[b] == -1) {
                    long dx = sl[u].i - sl[b].i, dy = sl[u].j - sl[b].j;
                    if (dx * dx + dy * dy <= r2) { label[b] = ncomp; stack[top++] = b; }
                }
        }
        ncomp++;
    }
    unsigned char *pool = (unsigned char *)calloc((size_t)m, 1);     /* local index */
    int *sub = (int *)malloc(sizeof(int) * (size_t)m), *outl = (int *)malloc(sizeof(int) * (size_t)m);
    int *Li = (int *)malloc(sizeof(int) * (size_t)(m + 1)), *Lj = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    int *Lid = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    for (int p = 0; p < m; p++) {
        if (pool[p]) continue;
        int c = label[p];
        if (c == -1) continue;
        int nsub = 0;
        for (int b = 0; b < m; b++) if (label[b] == c) sub[nsub++] = b;
        long ci = sl[p].i, cj = sl[p].j;
        long rad = r;
        int nl = 0;
        Li[nl] = sl[p].i; Lj[nl] = sl[p].j; Lid[nl] = p; nl++;   /* Local = [p[1]] (seed counted again below) */
        int ini = -1;
        while (nsub) {
            int nout = 0;
            for (int t = 0; t < nsub; t++) {
                int q = sub[t];
                if (pool[q]) continue;
                long dx = sl[q].i - ci, dy = sl[q].j - cj;
                /* euclidean(q, cen) <= rad  <=>  dx^2+dy^2 <= rad^2 for integer inputs */
                if (dx * dx + dy * dy <= rad * rad) { Li[nl] = sl[q].i; Lj[nl] = sl[q].j; Lid[nl] = q; nl++; }
                else outl[nout++] = q;
            }
            if (nout == ini) break;
            ini = nout;
            /* cen = round(mean(Local)) (np.round half-to-even), mean = sequential sum / count */
            double si = 0.0, sj = 0.0;
            for (int t = 0; t < nl; t++) { si += (double)Li[t]; sj += (double)Lj[t]; }
            ci = (long)nearbyint(si / (double)nl);
            cj = (long)nearbyint(sj / (double)nl);
            double mx = 0.0;
            for (int t = 0; t < nl; t++) {
                double dx = (double)(ci - Li[t]), dy = (double)(cj - Lj[t]);
                double d = sqrt(dx * dx + dy * dy);
                if (t == 0 || d > mx) mx = d;
            }
            rad = (long)nearbyint(mx) + r;
            for (int t = 0; t < nout; t++) sub[t] = outl[t];
            nsub = nout;
        }
        for (int t = 0; t < nl; t++) pool[Lid[t]] = 1;
        is_final[sl[p].id] = 1;
    }
    for (int b = 0; b < m; b++) if (pool[b]) visited[sl[b].id] = 1;
    free(label); free(stack); free(pool); free(sub); free(outl); free(Li); free(Lj); free(Lid);
}

/* Peakachu._local_clustering for one bucket.  out_keep[k] = 1 when point k is in the
 * returned list.  Returns the number kept. */
/* priorities (heights) of the height-filtered local maxima of the coordinate histogram, in
 * position order: the array scipy hands to numpy.argsort inside _select_by_peak_distance. */
NLFO_API int nlfo_anchor_priorities(const int *pos, int npos, int res, int min_count, int r_bp, double *prio_out)
{
    int m = 0;
    if (npos == 0) return 0;
    find_anchors(pos, npos, res, min_count, r_bp, NULL, NULL, prio_out, &m);
    return m;
}

NLFO_API int nlfo_pool_bucket(const int *pi, const int *pj, const double *val, int np_, int res,
                              int min_count, int r_bp, const int *x_order, const int *y_order,
                              unsigned char *out_keep)
{
    memset(out_keep, 0, (size_t)np_);
    if (np_ == 0) return 0;
    anchor_t *xa = (anchor_t *)malloc(sizeof(anchor_t) * (size_t)np_ + 16);
    anchor_t *ya = (anchor_t *)malloc(sizeof(anchor_t) * (size_t)np_ + 16);
    int nxa = find_anchors(pi, np_, res, min_count, r_bp, xa, x_order, NULL, NULL);
    int nya = find_anchors(pj, np_, res, min_count, r_bp, ya, y_order, NULL, NULL);
    int r = r_bp / res; if (r < 2) r = 2;
    unsigned char *visited = (unsigned char *)calloc((size_t)np_, 1);
    unsigned char *is_final = (unsigned char *)calloc((size_t)np_, 1);
    spt_t *sl = (spt_t *)malloc(sizeof(spt_t) * (size_t)np_);
    for (int a = 0; a < nxa; a++)
        for (int b = 0; b < nya; b++) {
            int m = 0;
            for (int k = 0; k < np_; k++)
                if (pi[k] >= xa[a].lb && pi[k] <= xa[a].rb && pj[k] >= ya[b].lb && pj[k] <= ya[b].rb) {
                    sl[m].v = val[k]; sl[m].i = pi[k]; sl[m].j = pj[k]; sl[m].id = k; m++;
                }
            qsort(sl, (size_t)m, sizeof(spt_t), spt_desc);
            cluster_core(sl, m, r, visited, is_final);
        }
    int m = 0;
    for (int k = 0; k < np_; k++)
        if (!visited[k]) { sl[m].v = val[k]; sl[m].i = pi[k]; sl[m].j = pj[k]; sl[m].id = k; m++; }
    qsort(sl, (size_t)m, sizeof(spt_t), spt_desc);
    cluster_core(sl, m, r, visited, is_final);
    int kept = 0;
    for (int k = 0; k < np_; k++) {
        if (is_final[k] || !visited[k]) { out_keep[k] = 1; kept++; }
    }
    free(xa); free(ya); free(visited); free(is_final); free(sl);
    return kept;
}

/* ------------------------------------------------------------------------------------
 * Row f-2: the reference's reads of a cooler, restated on cooler's own pixel table
 * (bin1_offset index, bin2 ascending inside a row, bin1 <= bin2, int32 counts, one weight per bin).
 * `clr.matrix(balance=col).fetch(r1, r2)` returns the symmetric completion of the stored upper
 * triangle, balanced as (count * w[row bin]) * w[column bin]; absent pixels are 0.
 * ---------------------------------------------------------------------------------- */
static double px_lookup(const int64_t *off, const int32_t *bin2, const int32_t *cnt, int64_t a, int64_t b)
{
    if (a > b) { int64_t t = a; a = b; b = t; }
    int64_t lo = off[a], hi = off[a + 1];
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (bin2[m] < b) lo = m + 1; else hi = m;
    }
    return (lo < off[a + 1] && bin2[lo] == b) ? (double)cnt[lo] : 0.0;
}

/* complexSV.reorganize_matrix (neoloop/assembly.py:408-492): blocks laid side by side, block k =
 * global bins [src_lo[k], src_hi[k]) read backwards when rev[k]; every block pair j1 <= j2 is
 * fetched, flipped per orientation and pasted (get_matrix clears NaN when balancing,
 * assembly.py:399-406), then np.triu; bias = weights with NaN -> 0 (get_bias, assembly.py:387-397). */
NLFO_API void nlfo_assemble_pixels(const int64_t *off, const int32_t *bin2, const int32_t *cnt, const double *w,
                                   int balance, int nblk, const int64_t *src_lo, const int64_t *src_hi,
                                   const unsigned char *rev, double *M, double *bias)
{
    long n = 0;
    long *start = (long *)malloc(sizeof(long) * (nblk + 1));
    for (int k = 0; k < nblk; k++) { start[k] = n; n += (long)(src_hi[k] - src_lo[k]); }
    start[nblk] = n;
    memset(M, 0, sizeof(double) * (size_t)n * n);
    for (int j1 = 0; j1 < nblk; j1++) {
        const long n1 = (long)(src_hi[j1] - src_lo[j1]);
        for (long i = 0; i < n1; i++) {
            const int64_t a = rev[j1] ? src_hi[j1] - 1 - i : src_lo[j1] + i;
            double v = w[a];
            bias[start[j1] + i] = (v != v) ? 0.0 : v;
        }
        for (int j2 = j1; j2 < nblk; j2++) {
            const long n2 = (long)(src_hi[j2] - src_lo[j2]);
            for (long i = 0; i < n1; i++) {
                const int64_t a = rev[j1] ? src_hi[j1] - 1 - i : src_lo[j1] + i;
                for (long j = 0; j < n2; j++) {
                    const int64_t b = rev[j2] ? src_hi[j2] - 1 - j : src_lo[j2] + j;
                    double v = px_lookup(off, bin2, cnt, a, b);
                    if (balance) {
                        v = (v * w[a]) * w[b];
                        if (v != v) v = 0.0;
                    }
                    const long r = start[j1] + i, c = start[j2] + j;
                    M[(size_t)r * n + c] = (r <= c) ? v : 0.0;          /* np.triu */
                }
            }
        }
    }
    free(start);
}

/* util._prepare_core (neoloop/util.py:403-428) for the chromosome of global bins [lo, hi):
 * sums[i] / cnts[i] = hic.diagonal(i)[valid].sum() / .size for i < nd, valid = both weights finite
 * and positive; numpy's pairwise order. */
NLFO_API void nlfo_expected_chrom(const int64_t *off, const int32_t *bin2, const int32_t *cnt, const double *w,
                                  int balance, int64_t lo, int64_t hi, int nd, double *sums, long *cnts)
{
    const long n = (long)(hi - lo);
    double *buf = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < nd; i++) {
        long m = 0;
        for (long r = 0; r + i < n; r++) {
            const double wa = w[lo + r], wb = w[lo + r + i];
            if (!(isfinite(wa) && wa > 0 && isfinite(wb) && wb > 0)) continue;
            double v = px_lookup(off, bin2, cnt, lo + r, lo + r + i);
            if (balance) v = (v * wa) * wb;
            buf[m++] = v;
        }
        sums[i] = m ? nlfo_pairwise_sum(buf, m, 1) : 0.0;
        cnts[i] = m;
    }
    free(buf);
}

NLFO_API int nlfo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
