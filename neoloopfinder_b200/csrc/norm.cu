// norm.cu -- distance normalisation of one SV-assembly matrix on B200 (sm_100a):
//   K2 masked per-diagonal means per block pair, numpy summation order   (callers.py:348-367,
//                                                                         assembly.py:497-516)
//   K3 per-block-pair rescale                                            (assembly.py:536-557)
//   K4 gap columns (column sums in numpy's row order) + gap-free band    (callers.py:504-522)
// The regression between K2 and K3 (callers.py:370-412: Spearman + isotonic + Huber/L-BFGS on
// <= 400 points per pair) stays on the host; this file moves every pass over the n x n matrix.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

using namespace nlf;

struct nlf_assembly {
    nlf_ctx *ctx = nullptr;
    int n = 0, nblk = 0;
    double *d_M = nullptr, *d_bias = nullptr, *d_val = nullptr;
    int *d_blk = nullptr, *d_op = nullptr, *d_kept = nullptr, *d_blo = nullptr, *d_bhi = nullptr;
    std::vector<int> blo, bhi, kept;
    bool scales_set = false, gaps_done = false;
};

// value of hcm[r][c] = fusion_matrix[r][c] rescaled by the rule of its block pair
__device__ __forceinline__ double scaled_value(const double *__restrict__ M, int n, const int *__restrict__ blk,
                                               const int *__restrict__ op, const double *__restrict__ val, int nblk,
                                               int r, int c)
{
    double v = M[(size_t)r * n + c];
    int p = blk[r] * nblk + blk[c];
    int o = op[p];
    if (o == 1) v = __ddiv_rn(v, val[p]);
    else if (o == 2) v = __dmul_rn(v, val[p]);
    return v;
}

// K2: one WARP per (pair, diagonal).  The lanes walk the diagonal 32 rows at a time, a ballot
// keeps the valid pixels in row order (scratch row of the warp), then lanes 0-7 add them in numpy's
// pairwise order and lane 0 publishes the mean.  (The first version used one thread per diagonal, serial
// over all rows, and a second kernel for the means: 148 + 83 us per assembly on cfg2.)
__global__ void __launch_bounds__(256) k_pair_diag_means(const double *__restrict__ M, int n, const double *__restrict__ bias,
                                  const int *__restrict__ pair_lo1, const int *__restrict__ pair_hi1,
                                  const int *__restrict__ pair_lo2, const int *__restrict__ pair_hi2, int npairs,
                                  int maxdiag, int mink, double *__restrict__ scratch, int maxlen,
                                  int *__restrict__ cnt, double *__restrict__ mean)
{
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nd = maxdiag + 1;
    if (wid >= npairs * nd) return;
    const int lane = threadIdx.x & 31;
    const int p = wid / nd, i = wid % nd;
    double *my = scratch + (size_t)wid * maxlen;
    int k = 0;
    if (i >= 2 && i < n - 1) {          // for i in range(2, M.shape[0]-1)
        const int lo1 = pair_lo1[p], hi1 = pair_hi1[p], lo2 = pair_lo2[p], hi2 = pair_hi2[p];
        const int xs = max(lo1, lo2 - i), xe = min(hi1, min(hi2, n) - i);
        for (int base = xs; base < xe; base += 32) {
            const int x = base + lane;
            bool ok = false;
            if (x < xe) ok = __dmul_rn(bias[x], bias[x + i]) > 0;
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok) my[k + __popc(mask & ((1u << lane) - 1u))] = M[(size_t)x * n + x + i];
            k += __popc(mask);
        }
    }
    __syncwarp();
    double m = __longlong_as_double(0x7ff8000000000000LL);
    if (k >= mink && k > 0 && lane < 8) {
        const double s = np_pairwise_sum_group8([&](int e) { return my[e]; }, k, lane, 0xffu);
        const double mm = __ddiv_rn(s, (double)k);
        if (mm > 0) m = mm;
    }
    if (lane == 0) {
        cnt[wid] = k;
        mean[wid] = m;
    }
}

// K4a fast path: per column, does hcm hold a positive entry (bit 0) / a negative or NaN entry (bit 1)?
// A column without bit 1 sums to zero exactly when it has no positive entry (a sum of non-negative
// doubles with one positive term cannot round to 0), whatever the order; columns with bit 1 take the
// sequential kernel below, which follows numpy's axis-0 order.  Fully parallel, one pass over M.
__global__ void k_gap_flags(const double *__restrict__ M, int n, const int *__restrict__ blk,
                            const int *__restrict__ op, const double *__restrict__ val, int nblk,
                            int *__restrict__ flags)
{
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j >= n) return;
    const int r0 = blockIdx.y * 64, r1 = min(r0 + 64, n);
    int f = 0;
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
        const double v = scaled_value(M, n, blk, op, val, nblk, r, j);
        if (v > 0) f |= 1;
        else if (!(v == 0)) f |= 2;
    }
    if (f) atomicOr(&flags[j], f);
}

// K4a exact path: column sums of hcm in numpy's axis-0 order (row after row), one thread per listed column.
// The rescale rule only changes at block boundaries, so the loop runs block by block with the
// rule in registers and eight independent loads in flight; the additions stay strictly sequential.
__global__ void k_gap_columns(const double *__restrict__ M, int n, const int *__restrict__ blk,
                              const int *__restrict__ blo, const int *__restrict__ bhi,
                              const int *__restrict__ op, const double *__restrict__ val, int nblk,
                              const int *__restrict__ cols, int ncols, unsigned char *__restrict__ is_gap)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncols) return;
    const int j = cols ? cols[t] : t;
    const int bj = blk[j];
    double s = 0.0;
    bool first = true;
    for (int b = 0; b < nblk; b++) {
        const int o = op[b * nblk + bj];
        const double v = val[b * nblk + bj];
        int r = blo[b];
        const int re = bhi[b];
        const double *col = M + j;
        auto sc = [&](double a) { return (o == 1) ? __ddiv_rn(a, v) : (o == 2) ? __dmul_rn(a, v) : a; };
        if (first && r < re) { s = sc(col[(size_t)r * n]); r++; first = false; }     // out[j] starts as M[0][j]
        for (; r + 8 <= re; r += 8) {
            double a0 = col[(size_t)(r + 0) * n], a1 = col[(size_t)(r + 1) * n], a2 = col[(size_t)(r + 2) * n],
                   a3 = col[(size_t)(r + 3) * n], a4 = col[(size_t)(r + 4) * n], a5 = col[(size_t)(r + 5) * n],
                   a6 = col[(size_t)(r + 6) * n], a7 = col[(size_t)(r + 7) * n];
            s = __dadd_rn(s, sc(a0)); s = __dadd_rn(s, sc(a1)); s = __dadd_rn(s, sc(a2)); s = __dadd_rn(s, sc(a3));
            s = __dadd_rn(s, sc(a4)); s = __dadd_rn(s, sc(a5)); s = __dadd_rn(s, sc(a6)); s = __dadd_rn(s, sc(a7));
        }
        for (; r < re; r++) s = __dadd_rn(s, sc(col[(size_t)r * n]));
    }
    is_gap[t] = (s == 0) ? 1 : 0;
}

// K4b: gap-free band straight from the fusion matrix: band[i][d] = hcm[kept[i]][kept[i+d]]
__global__ void k_gap_band(const double *__restrict__ M, int n, const int *__restrict__ blk,
                           const int *__restrict__ op, const double *__restrict__ val, int nblk,
                           const int *__restrict__ kept, int nk, int bw, int pitch, double *__restrict__ band)
{
    int i = blockIdx.y;
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= pitch) return;
    double v = 0.0;
    if (d < bw && i + d < nk) {
        v = scaled_value(M, n, blk, op, val, nblk, kept[i], kept[i + d]);
        if (!isfinite(v)) v = 0.0;
    }
    band[(size_t)i * pitch + d] = v;
}

// dense outputs for the public attributes complexSV.hcm / Fusion.gap_free
__global__ void k_dense_out(const double *__restrict__ M, int n, const int *__restrict__ blk,
                            const int *__restrict__ op, const double *__restrict__ val, int nblk,
                            const int *__restrict__ kept, int nk, double *__restrict__ out)
{
    int i = blockIdx.y;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nk) return;
    int r = kept ? kept[i] : i, c = kept ? kept[j] : j;
    out[(size_t)i * nk + j] = scaled_value(M, n, blk, op, val, nblk, r, c);
}

// a CUDA failure after the assembly object exists hands its device memory back before reporting
#define NLF_CUDA_A(call)                                                                            \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            cudaGetLastError();                                                                     \
            nlf_assembly_destroy(a);                                                                \
            return nlf::fail(NLF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                             cudaGetErrorString(e__));                                              \
        }                                                                                           \
    } while (0)

// block table checks + device allocations shared by the two constructors
static int assembly_alloc(nlf_ctx *ctx, int n, const int *blk_lo, const int *blk_hi, int nblk, nlf_assembly **out)
{
    if (n > 65535) return fail(NLF_ERR_ARG, "assembly of %d bins exceeds the dense limit 65535", n);
    std::vector<int> blk(n, -1);
    for (int b = 0; b < nblk; b++) {
        if (blk_lo[b] < 0 || blk_hi[b] > n || blk_lo[b] > blk_hi[b])
            return fail(NLF_ERR_ARG, "block %d range [%d,%d) outside [0,%d)", b, blk_lo[b], blk_hi[b], n);
        for (int r = blk_lo[b]; r < blk_hi[b]; r++) blk[r] = b;
    }
    for (int r = 0; r < n; r++)
        if (blk[r] < 0) return fail(NLF_ERR_ARG, "bin %d belongs to no block", r);
    for (int b = 0; b < nblk; b++)
        if (blk_lo[b] != (b ? blk_hi[b - 1] : 0))
            return fail(NLF_ERR_ARG, "blocks must tile [0,%d) in ascending order (block %d starts at %d)", n, b, blk_lo[b]);
    NLF_CUDA(cudaSetDevice(ctx->device));
    nlf_assembly *a = new nlf_assembly();
    a->ctx = ctx; a->n = n; a->nblk = nblk;
    a->blo.assign(blk_lo, blk_lo + nblk);
    a->bhi.assign(blk_hi, blk_hi + nblk);
    cudaStream_t s = ctx->stream;
    NLF_CUDA_A(cudaMallocAsync(&a->d_M, sizeof(double) * (size_t)n * n, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_bias, sizeof(double) * n, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_blk, sizeof(int) * n, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_op, sizeof(int) * nblk * nblk, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_val, sizeof(double) * nblk * nblk, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_kept, sizeof(int) * n, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_blo, sizeof(int) * nblk, s));
    NLF_CUDA_A(cudaMallocAsync(&a->d_bhi, sizeof(int) * nblk, s));
    NLF_CUDA_A(cudaMemcpyAsync(a->d_blo, blk_lo, sizeof(int) * nblk, cudaMemcpyHostToDevice, s));
    NLF_CUDA_A(cudaMemcpyAsync(a->d_bhi, blk_hi, sizeof(int) * nblk, cudaMemcpyHostToDevice, s));
    NLF_CUDA_A(cudaMemcpyAsync(a->d_blk, blk.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    NLF_CUDA_A(cudaMemsetAsync(a->d_op, 0, sizeof(int) * nblk * nblk, s));
    NLF_CUDA_A(cudaMemsetAsync(a->d_val, 0, sizeof(double) * nblk * nblk, s));
    // (pageable sources are staged before cudaMemcpyAsync returns: `blk` may go out of scope)
    *out = a;
    return NLF_OK;
}

NLF_EXPORT int nlf_assembly_create(nlf_ctx *ctx, const double *M, int n, const double *bias, const int *blk_lo,
                                   const int *blk_hi, int nblk, nlf_assembly **out)
{
    if (!ctx || !M || !bias || !blk_lo || !blk_hi || !out || n <= 0 || nblk <= 0)
        return fail(NLF_ERR_ARG, "nlf_assembly_create: bad arguments");
    nlf_assembly *a = nullptr;
    int rc = assembly_alloc(ctx, n, blk_lo, blk_hi, nblk, &a);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    NLF_CUDA_A(cudaMemcpyAsync(a->d_M, M, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, s));
    NLF_CUDA_A(cudaMemcpyAsync(a->d_bias, bias, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    *out = a;
    return NLF_OK;
}

// complexSV.reorganize_matrix on the device (assembly.py:408-492): block b = global bins
// [src_lo[b], src_hi[b]) of the resident pixel table, reversed when rev[b]
NLF_EXPORT int nlf_assembly_create_from_pixels(nlf_ctx *ctx, nlf_pixels *px, int balance, const int64_t *src_lo,
                                               const int64_t *src_hi, const uint8_t *rev, int nblk, nlf_assembly **out)
{
    if (!ctx || !px || !src_lo || !src_hi || !rev || !out || nblk <= 0)
        return fail(NLF_ERR_ARG, "nlf_assembly_create_from_pixels: bad arguments");
    if (px->ctx != ctx) return fail(NLF_ERR_ARG, "nlf_assembly_create_from_pixels: pixel table lives on another context");
    std::vector<int> blo(nblk), bhi(nblk);
    std::vector<long long> lo(nblk), hi(nblk);
    long long n = 0;
    for (int b = 0; b < nblk; b++) {
        if (src_lo[b] < 0 || src_hi[b] > px->nbins || src_lo[b] > src_hi[b])
            return fail(NLF_ERR_ARG, "block %d bins [%lld,%lld) outside the pixel table [0,%lld)", b, (long long)src_lo[b],
                        (long long)src_hi[b], px->nbins);
        lo[b] = src_lo[b]; hi[b] = src_hi[b];
        blo[b] = (int)n;
        n += src_hi[b] - src_lo[b];
        if (n > 65535) return fail(NLF_ERR_ARG, "assembly exceeds the dense limit of 65535 bins");
        bhi[b] = (int)n;
    }
    if (n <= 0) return fail(NLF_ERR_ARG, "nlf_assembly_create_from_pixels: empty assembly");
    nlf_assembly *a = nullptr;
    int rc = assembly_alloc(ctx, (int)n, blo.data(), bhi.data(), nblk, &a);
    if (rc) return rc;
    rc = nlf_pixels_fill_assembly(px, balance, lo.data(), hi.data(), rev, blo.data(), nblk, (int)n, a->d_blk, a->d_M, a->d_bias);
    if (rc) { nlf_assembly_destroy(a); return rc; }
    *out = a;
    return NLF_OK;
}

// replace the per-bin validity vector (Fusion.extract_bias keeps 1 / weight, callers.py:262-283)
NLF_EXPORT int nlf_assembly_set_bias(nlf_assembly *a, const double *bias)
{
    if (!a || !bias) return fail(NLF_ERR_ARG, "nlf_assembly_set_bias: NULL argument");
    NLF_CUDA(cudaSetDevice(a->ctx->device));
    NLF_CUDA(cudaMemcpyAsync(a->d_bias, bias, sizeof(double) * a->n, cudaMemcpyHostToDevice, a->ctx->stream));
    return NLF_OK;
}

NLF_EXPORT int nlf_assembly_size(nlf_assembly *a, int *n)
{
    if (!a || !n) return fail(NLF_ERR_ARG, "nlf_assembly_size: NULL argument");
    *n = a->n;
    return NLF_OK;
}

// complexSV.fusion_matrix / complexSV.bias (assembly.py:483-485) back on the host
NLF_EXPORT int nlf_assembly_fetch_matrix(nlf_assembly *a, double *M, double *bias)
{
    if (!a || (!M && !bias)) return fail(NLF_ERR_ARG, "nlf_assembly_fetch_matrix: NULL argument");
    NLF_CUDA(cudaSetDevice(a->ctx->device));
    cudaStream_t s = a->ctx->stream;
    if (M) NLF_CUDA(cudaMemcpyAsync(M, a->d_M, sizeof(double) * (size_t)a->n * a->n, cudaMemcpyDeviceToHost, s));
    if (bias) NLF_CUDA(cudaMemcpyAsync(bias, a->d_bias, sizeof(double) * a->n, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    return NLF_OK;
}

// masked per-diagonal means of arbitrary rectangles [lo1,hi1) x [lo2,hi2) of the fusion matrix
static int rect_means(nlf_assembly *a, int nrect, const int *lo1, const int *hi1, const int *lo2, const int *hi2,
                      int maxdiag, int mink, double *mean_out, int *cnt_out)
{
    nlf_ctx *c = a->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int nd = maxdiag + 1;
    int maxlen = 1;
    for (int p = 0; p < nrect; p++) {
        if (lo1[p] < 0 || hi1[p] > a->n || lo2[p] < 0 || hi2[p] > a->n)
            return fail(NLF_ERR_ARG, "rectangle %d [%d,%d) x [%d,%d) outside the %d-bin matrix", p, lo1[p], hi1[p], lo2[p], hi2[p], a->n);
        maxlen = std::max(maxlen, hi1[p] - lo1[p]);
    }
    const int total = nrect * nd;
    int *d_par, *d_cnt;
    double *d_scratch, *d_mean;
    NLF_CUDA(cudaMallocAsync(&d_par, sizeof(int) * 4 * nrect, s));
    NLF_CUDA(cudaMallocAsync(&d_cnt, sizeof(int) * total, s));
    NLF_CUDA(cudaMallocAsync(&d_mean, sizeof(double) * total, s));
    NLF_CUDA(cudaMallocAsync(&d_scratch, sizeof(double) * (size_t)total * maxlen, s));
    NLF_CUDA(cudaMemcpyAsync(d_par, lo1, sizeof(int) * nrect, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_par + nrect, hi1, sizeof(int) * nrect, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_par + 2 * nrect, lo2, sizeof(int) * nrect, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_par + 3 * nrect, hi2, sizeof(int) * nrect, cudaMemcpyHostToDevice, s));
    {
        KLaunch k(c, P_NORM);
        k_pair_diag_means<<<(total + 7) / 8, 256, 0, s>>>(a->d_M, a->n, a->d_bias, d_par, d_par + nrect, d_par + 2 * nrect,
                                                          d_par + 3 * nrect, nrect, maxdiag, mink, d_scratch, maxlen,
                                                          d_cnt, d_mean);
    }
    NLF_CUDA(cudaGetLastError());
    NLF_CUDA(cudaMemcpyAsync(mean_out, d_mean, sizeof(double) * total, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(cnt_out, d_cnt, sizeof(int) * total, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    NLF_CUDA(cudaFreeAsync(d_par, s));
    NLF_CUDA(cudaFreeAsync(d_cnt, s));
    NLF_CUDA(cudaFreeAsync(d_mean, s));
    NLF_CUDA(cudaFreeAsync(d_scratch, s));
    return NLF_OK;
}

NLF_EXPORT int nlf_assembly_diag_means(nlf_assembly *a, int maxdiag, int mink, double *mean_out, int *cnt_out)
{
    if (!a || maxdiag < 0 || !mean_out || !cnt_out) return fail(NLF_ERR_ARG, "nlf_assembly_diag_means: bad arguments");
    const int nb = a->nblk;
    std::vector<int> lo1, hi1, lo2, hi2;
    for (int j1 = 0; j1 < nb; j1++)
        for (int j2 = j1; j2 < nb; j2++) {
            lo1.push_back(a->blo[j1]); hi1.push_back(a->bhi[j1]);
            lo2.push_back(a->blo[j2]); hi2.push_back(a->bhi[j2]);
        }
    return rect_means(a, (int)lo1.size(), lo1.data(), hi1.data(), lo2.data(), hi2.data(), maxdiag, mink, mean_out, cnt_out);
}

// Fusion._extract_xy for caller-chosen rectangles (Fusion.allele_slope, callers.py:414-478: the
// upstream / downstream / junction regions bounded by up_i and down_i)
NLF_EXPORT int nlf_assembly_rect_means(nlf_assembly *a, int nrect, const int *row_lo, const int *row_hi,
                                       const int *col_lo, const int *col_hi, int maxdiag, int mink,
                                       double *mean_out, int *cnt_out)
{
    if (!a || nrect <= 0 || !row_lo || !row_hi || !col_lo || !col_hi || maxdiag < 0 || !mean_out || !cnt_out)
        return fail(NLF_ERR_ARG, "nlf_assembly_rect_means: bad arguments");
    return rect_means(a, nrect, row_lo, row_hi, col_lo, col_hi, maxdiag, mink, mean_out, cnt_out);
}

NLF_EXPORT int nlf_assembly_set_scales(nlf_assembly *a, const int *op, const double *val)
{
    if (!a || !op || !val) return fail(NLF_ERR_ARG, "nlf_assembly_set_scales: NULL argument");
    nlf_ctx *c = a->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    const int nb2 = a->nblk * a->nblk;
    std::vector<int> o(op, op + nb2);
    for (int j1 = 0; j1 < a->nblk; j1++)
        for (int j2 = 0; j2 < j1; j2++) o[j1 * a->nblk + j2] = 0;   // only j1 <= j2 is ever rescaled
    NLF_CUDA(cudaMemcpyAsync(a->d_op, o.data(), sizeof(int) * nb2, cudaMemcpyHostToDevice, c->stream));
    NLF_CUDA(cudaMemcpyAsync(a->d_val, val, sizeof(double) * nb2, cudaMemcpyHostToDevice, c->stream));
    a->scales_set = true;
    a->gaps_done = false;
    return NLF_OK;
}

static int ensure_gaps(nlf_assembly *a)
{
    if (a->gaps_done) return NLF_OK;
    nlf_ctx *c = a->ctx;
    cudaStream_t s = c->stream;
    const int n = a->n;
    int *d_flags;
    NLF_CUDA(cudaMallocAsync(&d_flags, sizeof(int) * n, s));
    NLF_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * n, s));
    {
        KLaunch k(c, P_NORM);
        dim3 grid((n + 31) / 32, (n + 63) / 64), block(32, 8);
        k_gap_flags<<<grid, block, 0, s>>>(a->d_M, n, a->d_blk, a->d_op, a->d_val, a->nblk, d_flags);
    }
    NLF_CUDA(cudaGetLastError());
    std::vector<int> flags(n);
    NLF_CUDA(cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    NLF_CUDA(cudaFreeAsync(d_flags, s));
    std::vector<unsigned char> gap(n);
    std::vector<int> exact;
    for (int j = 0; j < n; j++) {
        gap[j] = (flags[j] & 1) ? 0 : 1;
        if (flags[j] & 2) exact.push_back(j);
    }
    if (!exact.empty()) {
        // columns with a negative or NaN entry: the sum in numpy's row order decides
        const int ne = (int)exact.size();
        int *d_cols;
        unsigned char *d_gap;
        std::vector<unsigned char> g(ne);
        NLF_CUDA(cudaMallocAsync(&d_cols, sizeof(int) * ne, s));
        NLF_CUDA(cudaMallocAsync(&d_gap, ne, s));
        NLF_CUDA(cudaMemcpyAsync(d_cols, exact.data(), sizeof(int) * ne, cudaMemcpyHostToDevice, s));
        {
            KLaunch k(c, P_NORM);
            k_gap_columns<<<(ne + 31) / 32, 32, 0, s>>>(a->d_M, n, a->d_blk, a->d_blo, a->d_bhi, a->d_op, a->d_val, a->nblk,
                                                         d_cols, ne, d_gap);
        }
        NLF_CUDA(cudaGetLastError());
        NLF_CUDA(cudaMemcpyAsync(g.data(), d_gap, ne, cudaMemcpyDeviceToHost, s));
        NLF_CUDA(cudaStreamSynchronize(s));
        NLF_CUDA(cudaFreeAsync(d_cols, s));
        NLF_CUDA(cudaFreeAsync(d_gap, s));
        for (int t = 0; t < ne; t++) gap[exact[t]] = g[t];
    }
    // index bookkeeping of callers.py:509-516 (w = 5: keep every bin when fewer than 11 survive)
    a->kept.clear();
    for (int j = 0; j < n; j++)
        if (!gap[j]) a->kept.push_back(j);
    if (5 * 2 + 1 > (int)a->kept.size()) {
        a->kept.resize(n);
        for (int j = 0; j < n; j++) a->kept[j] = j;
    }
    NLF_CUDA(cudaMemcpyAsync(a->d_kept, a->kept.data(), sizeof(int) * a->kept.size(), cudaMemcpyHostToDevice, s));
    a->gaps_done = true;
    return NLF_OK;
}

NLF_EXPORT int nlf_assembly_remove_gaps(nlf_assembly *a, int *kept_out, int *n_kept)
{
    if (!a || !n_kept) return fail(NLF_ERR_ARG, "nlf_assembly_remove_gaps: NULL argument");
    NLF_CUDA(cudaSetDevice(a->ctx->device));
    int rc = ensure_gaps(a);
    if (rc) return rc;
    *n_kept = (int)a->kept.size();
    if (kept_out) memcpy(kept_out, a->kept.data(), sizeof(int) * a->kept.size());
    return NLF_OK;
}

static int dense_out(nlf_assembly *a, bool gap_free, double *out)
{
    nlf_ctx *c = a->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    int nk = a->n;
    if (gap_free) {
        int rc = ensure_gaps(a);
        if (rc) return rc;
        nk = (int)a->kept.size();
    }
    double *d_out;
    NLF_CUDA(cudaMallocAsync(&d_out, sizeof(double) * (size_t)nk * nk, s));
    {
        KLaunch k(c, P_NORM);
        dim3 grid((nk + 127) / 128, nk);
        k_dense_out<<<grid, 128, 0, s>>>(a->d_M, a->n, a->d_blk, a->d_op, a->d_val, a->nblk,
                                         gap_free ? a->d_kept : nullptr, nk, d_out);
    }
    NLF_CUDA(cudaGetLastError());
    NLF_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)nk * nk, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    NLF_CUDA(cudaFreeAsync(d_out, s));
    return NLF_OK;
}

NLF_EXPORT int nlf_assembly_fetch_hcm(nlf_assembly *a, double *out)
{
    if (!a || !out) return fail(NLF_ERR_ARG, "nlf_assembly_fetch_hcm: NULL argument");
    return dense_out(a, false, out);
}

NLF_EXPORT int nlf_assembly_fetch_gap_free(nlf_assembly *a, double *out)
{
    if (!a || !out) return fail(NLF_ERR_ARG, "nlf_assembly_fetch_gap_free: NULL argument");
    return dense_out(a, true, out);
}

// used by nlf_batch_add_assembly (score.cu): the gap-free band never leaves the device
int nlf_assembly_export_band(nlf_assembly *a, int bw, int pitch, double **d_band, int *n_out)
{
    nlf_ctx *c = a->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    int rc = ensure_gaps(a);
    if (rc) return rc;
    const int nk = (int)a->kept.size();
    double *band;
    NLF_CUDA(cudaMallocAsync(&band, sizeof(double) * (size_t)nk * pitch, c->stream));
    {
        KLaunch k(c, P_BANDIFY);
        dim3 grid((pitch + 127) / 128, nk);
        k_gap_band<<<grid, 128, 0, c->stream>>>(a->d_M, a->n, a->d_blk, a->d_op, a->d_val, a->nblk, a->d_kept, nk, bw,
                                                pitch, band);
    }
    NLF_CUDA(cudaGetLastError());
    *d_band = band;
    *n_out = nk;
    return NLF_OK;
}

NLF_EXPORT int nlf_assembly_destroy(nlf_assembly *a)
{
    if (!a) return NLF_OK;
    cudaSetDevice(a->ctx->device);
    cudaStream_t s = a->ctx->stream;
    cudaFreeAsync(a->d_M, s); cudaFreeAsync(a->d_bias, s); cudaFreeAsync(a->d_blk, s);
    cudaFreeAsync(a->d_op, s); cudaFreeAsync(a->d_val, s); cudaFreeAsync(a->d_kept, s);
    cudaFreeAsync(a->d_blo, s); cudaFreeAsync(a->d_bhi, s);
    delete a;
    return NLF_OK;
}
