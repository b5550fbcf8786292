/*
 * nlf_b200.h -- C ABI of the B200-native neoloop-caller scoring path.
 *
 * One shared library (libnlf_b200.so, hand-written CUDA for sm_100a) behind plain C entry
 * points: pointers and sizes only, caller-owned host buffers, library-owned device memory
 * behind opaque handles, int return codes (0 = ok, <0 = error, message from
 * nlf_last_error()).  Every entry point names the reference interface it replaces
 * (paths relative to the NeoLoopFinder tree).  The reference-side binding is a ctypes stub,
 * shown in INTEGRATION.md and implemented in neoloopfinder_b200/_lib.py.
 *
 * Threading: a context owns one device, its streams and scratch memory; calls on one
 * context must not overlap in time, different contexts are independent (one context per
 * GPU / per worker process, as the reference fans out one process per assembly at
 * scripts/neoloop-caller:155-159).  There is no CPU fallback anywhere in the library.
 */
#ifndef NLF_B200_H
#define NLF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLF_OK 0
#define NLF_ERR_CUDA -1      /* a CUDA runtime call failed                                  */
#define NLF_ERR_ARG -2       /* invalid argument                                            */
#define NLF_ERR_EMPTY_FIT -3 /* no diagonal longer than 5: the reference's isotonic fit raises ValueError */
#define NLF_ERR_LOWER_TRI -4 /* non-zero below the main diagonal: the path expects np.triu input */
#define NLF_ERR_STATE -5     /* calls made in the wrong order                                */

typedef struct nlf_ctx nlf_ctx;
typedef struct nlf_forest nlf_forest;
typedef struct nlf_assembly nlf_assembly;
typedef struct nlf_batch nlf_batch;
typedef struct nlf_pixels nlf_pixels;

/* ---- library / context ----------------------------------------------------------------- */
int nlf_version(void);
const char *nlf_last_error(void);            /* thread-local, valid until the next failing call */
int nlf_device_count(int *count);
int nlf_ctx_create(int device, nlf_ctx **out);
int nlf_ctx_destroy(nlf_ctx *ctx);
int nlf_ctx_sync(nlf_ctx *ctx);
/* CUDA-event timer on the context's compute stream (the stream every kernel is launched on). */
int nlf_timer_start(nlf_ctx *ctx);
int nlf_timer_stop(nlf_ctx *ctx, float *elapsed_ms); /* records, synchronises, returns ms */
/* Kernels launched by this context so far (bench.py's "gpu_launches"). */
int nlf_launch_count(nlf_ctx *ctx, long long *count);
/* Accumulated CUDA-event time per kernel class since the last reset, for the roofline line.
 * which: 0 bandify, 1 diag means+isotonic, 2 feature kernel, 3 forest kernel, 4 threshold. */
int nlf_profile_enable(nlf_ctx *ctx, int on);
int nlf_profile_read(nlf_ctx *ctx, int which, float *total_ms, long long *launches);
int nlf_profile_reset(nlf_ctx *ctx);
/* Measurement helpers: evict L2 between timed iterations; fp64 add/mul issue rate (G op/s,
 * no FMA) measured on this device, the roofline denominator of the feature kernel. */
int nlf_l2_flush(nlf_ctx *ctx);
int nlf_fp64_peak(nlf_ctx *ctx, double *gops);

/* ---- forest: replaces Peakachu.load_models + model.predict_proba
 *      (neoloop/callers.py:572-577, 622; sklearn _forest.py predict_proba, _tree.pyx _apply_dense).
 * The host flattens the sklearn model ONCE per model file into these arrays:
 * tree t owns nodes [tree_off[t], tree_off[t+1]); left/right are tree-local child ids
 * (-1 at leaves); threshold is sklearn's float64 threshold; missing_left its
 * missing_go_to_left flag; leaf_p1[node] = tree_.value[node, 0, class 1]. ----------------- */
int nlf_forest_create(nlf_ctx *ctx, int n_trees, const int64_t *tree_off, const int32_t *left,
                      const int32_t *right, const int32_t *feature, const double *threshold,
                      const uint8_t *missing_left, const double *leaf_p1, int n_features,
                      nlf_forest **out);
/* Host-only view of the compact single-pass forest (internal nodes only; csrc/score.cu k_forest_one) that
 * nlf_forest_create builds next to the general node array: 1 = representable (arrays filled when the capacities
 * suffice; NULL arrays = size query), 0 = not representable (the general kernels are used), < 0 = bad arguments.
 * No CUDA call is made: exists so that the encoding can be checked against sklearn without a GPU. */
int nlf_forest_compact_host(int n_trees, const int64_t *tree_off, const int32_t *left, const int32_t *right,
                            const int32_t *feature, const double *threshold, const double *leaf_p1,
                            long long cap_nodes, uint32_t *node_thr, uint32_t *node_word, long long cap_leaf,
                            double *leafv, uint32_t *root, long long *n_int, long long *n_leaf);
int nlf_forest_destroy(nlf_forest *f);
/* predict_proba(X)[:, 1] for host float32 rows (N x n_features), results to proba[N]. */
int nlf_forest_predict(nlf_forest *f, const float *X, long long N, double *proba);

/* ---- resident pixel table: replaces the cooler reads behind
 *      clr.matrix(balance=col).fetch(r1, r2) and clr.bins().fetch(r)[col]
 *      (neoloop/assembly.py:387-406, neoloop/util.py:408-413).  The sample is uploaded ONCE in
 *      cooler's own pixel layout: bin1_offset[nbins+1] (the `indexes/bin1_offset` CSR index),
 *      bin2[nnz] strictly ascending inside a row with bin1 <= bin2, count[nnz] (int32, cooler's
 *      dtype), and one weight per bin (the chosen balancing column, NaN = masked bin; also the
 *      source of complexSV.bias).  Balanced values are (count * w[row bin]) * w[column bin]. ---- */
int nlf_pixels_create(nlf_ctx *ctx, long long nbins, const int64_t *bin1_offset, const int32_t *bin2,
                      const int32_t *count, const double *weights, nlf_pixels **out);
int nlf_pixels_destroy(nlf_pixels *px);

/* ---- assembly normalisation: replaces complexSV.correct_heterozygosity and
 *      Fusion.remove_gaps (neoloop/assembly.py:495-571, neoloop/callers.py:348-367, 504-522).
 * M is the n x n row-major np.triu fusion matrix (assembly.py:483), bias the per-bin
 * weights (NaN already 0), blk_lo/blk_hi the block index ranges (complexSV.index). -------- */
int nlf_assembly_create(nlf_ctx *ctx, const double *M, int n, const double *bias,
                        const int *blk_lo, const int *blk_hi, int nblk, nlf_assembly **out);
/* complexSV.reorganize_matrix on the device (assembly.py:408-492): block b of the assembly is the
 * global bin range [src_lo[b], src_hi[b]) of the pixel table, traversed in reverse when rev[b]
 * (orientation '-').  Builds np.triu of the rearranged matrix (NaN -> 0 when balance != 0) and
 * the bias vector (get_bias, assembly.py:387-397) without any host matrix.  balance = 0 uses raw
 * counts (balance_type False). */
int nlf_assembly_create_from_pixels(nlf_ctx *ctx, nlf_pixels *px, int balance, const int64_t *src_lo,
                                    const int64_t *src_hi, const uint8_t *rev, int nblk,
                                    nlf_assembly **out);
int nlf_assembly_size(nlf_assembly *a, int *n);
/* Replace the per-bin vector whose pairwise products decide which pixels are valid (a pixel
 * (x, y) counts when bias[x] * bias[y] > 0): Fusion.extract_bias keeps 1 / weight (callers.py:262-283). */
int nlf_assembly_set_bias(nlf_assembly *a, const double *bias);
/* complexSV.fusion_matrix (n x n) and complexSV.bias (n) back on the host; either may be NULL. */
int nlf_assembly_fetch_matrix(nlf_assembly *a, double *M, double *bias);
/* Fusion._extract_xy for every block pair j1<=j2 (pair index p = position in that nested
 * order) up to maxdiag diagonals: mean_out[p*(maxdiag+1)+i] = mean of valid pixels on
 * diagonal i (numpy summation order) when at least mink are valid and the mean is > 0,
 * NaN otherwise; cnt_out = number of valid pixels. */
int nlf_assembly_diag_means(nlf_assembly *a, int maxdiag, int mink, double *mean_out, int *cnt_out);
/* The same for caller-chosen rectangles [row_lo,row_hi) x [col_lo,col_hi) of the matrix (they need
 * not follow the block table): Fusion._extract_xy as Fusion.allele_slope uses it for the upstream,
 * downstream and junction regions between up_i and down_i (callers.py:414-478).  Rectangle p
 * fills mean_out / cnt_out[p*(maxdiag+1) ...]. */
int nlf_assembly_rect_means(nlf_assembly *a, int nrect, const int *row_lo, const int *row_hi,
                            const int *col_lo, const int *col_hi, int maxdiag, int mink,
                            double *mean_out, int *cnt_out);
/* assembly.py:536-549: op[j1*nblk+j2] 0 = keep, 1 = divide by val, 2 = multiply by val. */
int nlf_assembly_set_scales(nlf_assembly *a, const int *op, const double *val);
/* callers.py:504-522: kept_out[k] = original index of gap-free bin k; *n_kept = n'. */
int nlf_assembly_remove_gaps(nlf_assembly *a, int *kept_out, int *n_kept);
int nlf_assembly_fetch_hcm(nlf_assembly *a, double *out);       /* n  x n  (complexSV.hcm)      */
int nlf_assembly_fetch_gap_free(nlf_assembly *a, double *out);  /* n' x n' (Fusion.gap_free)    */
int nlf_assembly_destroy(nlf_assembly *a);

/* ---- per-sample prologue: replaces util._prepare_core (neoloop/util.py:403-428) for one
 *      chromosome given as its band (band[r*nd + i] = balanced M[r, r+i], i < nd = maxdis+1):
 *      sum_out[i] / cnt_out[i] = hic.diagonal(i)[valid].sum() / .size with valid[r] = weight of bin r
 *      finite and > 0 (numpy summation order).  calculate_expected's reduction over chromosomes and
 *      its isotonic fit (util.py:449-466) stay on the host. ------------------------------------ */
int nlf_expected_diagonals(nlf_ctx *ctx, const double *band, int n, int nd, const uint8_t *valid,
                           double *sum_out, long long *cnt_out);

/* The same from the resident pixel table, chromosome = global bins [lo, hi): band, valid mask
 * (weight finite and > 0, util.py:415) and sums never leave the device. */
int nlf_expected_from_pixels(nlf_ctx *ctx, nlf_pixels *px, int balance, long long lo, long long hi,
                             int nd, double *sum_out, long long *cnt_out);

/* ---- scoring: replaces Peakachu.__init__/get_candidates/getwindow/predict up to the
 *      Donuts dictionary (neoloop/callers.py:527-628) and util.distance_normalize /
 *      distance_normaize_core / image_normalize (neoloop/util.py:470-524) for a BATCH of
 *      gap-free matrices (one per assembly).  upper = 400 and lower = 4 in the reference
 *      pipeline (scripts/neoloop-caller:202, callers.py:536). -------------------------------- */
int nlf_batch_create(nlf_ctx *ctx, int lower, int upper, nlf_batch **out);
/* Add one gap-free matrix from host memory (dense n x n row-major, upper-triangular).
 * Returns the job index (>= 0) or an error (< 0).  The copy is asynchronous: G must stay
 * valid until nlf_batch_score has been called. */
int nlf_batch_add_dense(nlf_batch *b, const double *G, int n);
/* Add the gap-free matrix of a normalised assembly without leaving the device. */
int nlf_batch_add_assembly(nlf_batch *b, nlf_assembly *a);
/* Add a matrix given as its band: band[r*pitch + d] = G[r, r+d], d < upper+14 (banded
 * genome-wide input). */
int nlf_batch_add_band(nlf_batch *b, const double *band, int n, int pitch);
/* Add a whole chromosome (global bins [lo, hi) of a resident pixel table) as a trivial assembly:
 * Fusion.remove_gaps (callers.py:504-522; a column is a gap when it holds no non-zero value) and
 * the finite band of the gap-free matrix (callers.py:527-535) are built on the device.
 * kept_out[k] = bin, relative to lo, behind gap-free index k (capacity hi - lo); *n_kept = n'. */
int nlf_batch_add_pixels(nlf_batch *b, nlf_pixels *px, int balance, long long lo, long long hi,
                         int *kept_out, int *n_kept);
/* Keep the float32 feature rows addressable after scoring (needed by
 * nlf_batch_fetch_features32; costs one int per band cell).  Call before nlf_batch_score. */
int nlf_batch_keep_features(nlf_batch *b, int on);
/* Run candidates -> window features -> forest -> threshold for every job (asynchronous).
 * expected[nexp] = Peakachu.exp_arr; w = window half-size (features = (2w+1)^2);
 * gw[2*radius+1] = scipy's gaussian kernel weights for sigma=1 (computed by the host with
 * numpy exactly as scipy does); thre = probability threshold (callers.py:625). */
int nlf_batch_score(nlf_batch *b, nlf_forest *f, const double *expected, int nexp, int w,
                    const double *gw, int radius, double thre);
/* Synchronise and report per-job counts: candidates (len(ridx)), scored windows
 * (len(clist)), pixels with proba >= thre (len(Donuts)).  Arrays of n_jobs entries. */
int nlf_batch_counts(nlf_batch *b, long long *n_cand, long long *n_scored, long long *n_donuts);
/* Donuts of one job: coordinates, probability, and the contact value M[i,j]. */
int nlf_batch_fetch_donuts(nlf_batch *b, int job, int *i, int *j, double *prob, double *value);
/* Donuts of ALL jobs in one call (arrays of capacity cap; any order; job[k] tags the owner).
 * Pass job = NULL to query *n_total only. */
int nlf_batch_fetch_donuts_all(nlf_batch *b, long long cap, int *job, int *i, int *j, double *prob,
                               double *value, long long *n_total);
/* Every scored pixel of one job in candidate order (diagonal-major, row-ascending):
 * clist and probas of Peakachu.predict.  cap = capacity of the arrays. */
int nlf_batch_fetch_scored(nlf_batch *b, int job, long long cap, int *i, int *j, double *prob);
/* Diagonal means and their decreasing isotonic fit (callers.py:540-551) of one job: arrays of
 * upper+1 doubles indexed by the diagonal (entries below `lower` are NaN; means of diagonals with
 * at most 5 entries are NaN). */
int nlf_batch_fetch_expected(nlf_batch *b, int job, double *means, double *e);
/* Peakachu.ridx / cidx (callers.py:553-570) of one job. */
int nlf_batch_fetch_candidates(nlf_batch *b, int job, long long cap, int *ridx, int *cidx);
/* float32 feature rows (what sklearn sees after its cast) of the scored pixels of one job, in
 * the order of nlf_batch_fetch_scored.  fea: n_scored x (2w+1)^2. */
int nlf_batch_fetch_features32(nlf_batch *b, int job, long long cap, float *fea);
/* Peakachu.getwindow(coords, w) for arbitrary coordinates on one job's matrix: float64
 * features, rows only for windows that pass the reference's masks. *n_rows = rows written. */
int nlf_batch_getwindow(nlf_batch *b, int job, const int *xi, const int *yi, long long nc, int w,
                        const double *expected, int nexp, const double *gw, int radius,
                        double *fea, int *clist, long long *n_rows);
/* util.distance_normalize + distance_normaize_core (neoloop/util.py:470-524) on a caller-provided stack of
 * N windows ((2w+1)^2 doubles each, centred at (xi, yi)): NaN -> 0, the count_nonzero and lower-left-mean
 * gates, observed / expected.  out receives every window (normalised when it passes and its largest
 * distance is inside the expected table), pass[k] = 1 when window k is kept. */
int nlf_distance_normalize(nlf_ctx *ctx, const double *windows, long long N, const int *xi, const int *yi,
                           int w, const double *expected, int nexp, double *out, uint8_t *pass);
/* Drop the results but keep the matrices resident, so nlf_batch_score can run again. */
/* Apply another probability threshold to a scored batch (callers.py:623-628 again): the Donut list and the
 * per-job scored / Donut counts are rebuilt from the probabilities kept in HBM; nothing is re-scored. */
int nlf_batch_rethreshold(nlf_batch *b, double thre);
int nlf_batch_reset(nlf_batch *b);
int nlf_batch_njobs(nlf_batch *b);
int nlf_batch_destroy(nlf_batch *b);

/* ---- pooling: replaces Peakachu._local_clustering / find_anchors / _cluster_core
 *      (neoloop/callers.py:699-831) for a set of buckets (one per assembly and chain pair,
 *      callers.py:671-683).  Points of bucket k are [bucket_off[k], bucket_off[k+1]).
 *      x_order / y_order: optional per-bucket permutations (concatenated, offsets in
 *      xo_off / yo_off) that numpy.argsort returned for the peak heights of scipy's
 *      distance filter; NULL = stable order (they only matter below 10 kb).
 *      keep_out[p] = 1 when point p is in the pooled loop list. ------------------------------ */
int nlf_pool(nlf_ctx *ctx, int n_buckets, const int64_t *bucket_off, const int *pi, const int *pj,
             const double *val, int res, int min_count, int r_bp, const int *x_order,
             const int64_t *xo_off, const int *y_order, const int64_t *yo_off, uint8_t *keep_out);
/* Peak heights (position order) that scipy hands to numpy.argsort for the x (axis=0) or
 * y (axis=1) coordinate histogram of every bucket; prio_off[n_buckets+1] receives offsets.
 * *ambiguous = 1 when some bucket has two equal heights closer than the distance filter,
 * i.e. when the reference's result depends on numpy's tie order. */
int nlf_pool_peak_heights(nlf_ctx *ctx, int n_buckets, const int64_t *bucket_off, const int *pos,
                          int res, int min_count, int r_bp, double *prio_out, int64_t *prio_off,
                          int *ambiguous);

/* Peakachu.find_anchors (callers.py:740-784) for one coordinate list: anchors_out[3*k ..] =
 * (summit, left bound, right bound) of anchor k; cap >= max(pos) - min(pos) + 1.  min_dis_bp is the
 * method's min_dis (bp); order: optional numpy.argsort permutation of the peak heights (see
 * nlf_pool_peak_heights), NULL = stable order. */
int nlf_pool_find_anchors(nlf_ctx *ctx, const int *pos, int npos, int res, int min_count, int min_dis_bp,
                          const int *order, int *anchors_out, int cap, int *n_anchors);
/* Same with the caller's peak-width window `wlen_bp` (Peakachu.find_anchors(..., wlen=...), callers.py:740-749:
 * wlen = min(wlen_bp // res, 20) bins at res <= 10 kb, 4 bins above; a window of one bin or less is scipy's
 * "`wlen` must be larger than 1" error). */
int nlf_pool_find_anchors_wlen(nlf_ctx *ctx, const int *pos, int npos, int res, int min_count, int min_dis_bp,
                               int wlen_bp, const int *order, int *anchors_out, int cap, int *n_anchors);
/* Peakachu._cluster_core (callers.py:786-831) on points already sorted by descending
 * (value, (i, j)): is_final_out[k] = 1 when point k is appended to final_list, pooled_out[k] = 1
 * when it joins `visited`.  r_bins is the method's r (bins). */
int nlf_pool_cluster_core(nlf_ctx *ctx, const int *pi, const int *pj, int m, int r_bins,
                          uint8_t *is_final_out, uint8_t *pooled_out);

#ifdef __cplusplus
}
#endif
#endif /* NLF_B200_H */
