// expected.cu -- per-sample prologue of neoloop-caller on B200 (sm_100a):
//   util._prepare_core (neoloop/util.py:403-428): for one chromosome and every genomic distance i,
//   sum and count of the balanced contacts on diagonal i restricted to bin pairs whose two weights
//   are finite and positive -- `hic.diagonal(i)[valid].sum()` / `.size` in numpy's summation order.
// The chromosome arrives as its band (n x nd, band[r][i] = M[r, r+i]); the cross-chromosome
// accumulation (Python floats) and the isotonic fit of calculate_expected stay on the host.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

using namespace nlf;

// ---- ordered compaction of the valid entries of every diagonal ------------------------------------
// scratch[k * nd + i] = k-th valid entry (row order) of diagonal i: neighbouring diagonals are
// neighbours in memory, so the band reads and the scratch writes of a warp are coalesced.  Rows are
// cut into segments of EXP_SEG rows; a thread owns one (segment, diagonal): count, scan over the
// segments, scatter.  (The first version walked all rows of a diagonal with ONE thread.)
constexpr int EXP_SEG = 128;

__global__ void k_exp_segcount(int n, int nd, const unsigned char *__restrict__ valid, int *__restrict__ segcnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (i >= nd) return;
    const int r0 = seg * EXP_SEG, r1 = min(r0 + EXP_SEG, n - i);      // diagonal i has n-i entries
    int k = 0;
    for (int r = r0; r < r1; r++) k += (valid[r] && valid[r + i]);
    segcnt[(size_t)seg * nd + i] = k;
}

__global__ void k_exp_segscan(int nseg, int nd, int *__restrict__ segcnt, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    int run = 0;
    int seg = 0;
    for (; seg + 8 <= nseg; seg += 8) {             // eight independent loads in flight, then the prefix
        int c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = segcnt[(size_t)(seg + k) * nd + i];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            segcnt[(size_t)(seg + k) * nd + i] = run;   // exclusive prefix: first slot of the segment
            run += c[k];
        }
    }
    for (; seg < nseg; seg++) {
        const int c = segcnt[(size_t)seg * nd + i];
        segcnt[(size_t)seg * nd + i] = run;
        run += c;
    }
    cnt[i] = run;
}

__global__ void k_exp_scatter(const double *__restrict__ band, int n, int nd, const unsigned char *__restrict__ valid,
                              const int *__restrict__ segoff, double *__restrict__ scratch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (i >= nd) return;
    const int r0 = seg * EXP_SEG, r1 = min(r0 + EXP_SEG, n - i);
    int k = segoff[(size_t)seg * nd + i];
    for (int r = r0; r < r1; r++) {
        if (valid[r] && valid[r + i]) {
            scratch[(size_t)k * nd + i] = band[(size_t)r * nd + i];
            k++;
        }
    }
}

// ---- numpy pairwise sum of one diagonal by a whole block -------------------------------------------
// numpy (loops_utils.h.src, DOUBLE_pairwise_sum) splits n > 128 at n2 = n/2 - (n/2)%8 and adds the two
// halves.  The top EXP_DEPTH levels of that recursion are walked explicitly: 8-lane group g follows
// the bits of g from the root and sums its subtree with the sequential group routine (the subtree IS
// numpy's recursive call on that range); the partial sums are then added pairwise up the same tree,
// so every addition numpy makes is made once, with the same operands.  A node that becomes a leaf
// before EXP_DEPTH levels belongs to the group whose remaining path bits are zero.
constexpr int EXP_DEPTH = 5;                    // 32 groups of 8 lanes = 256 threads per diagonal

__device__ __forceinline__ void np_split(int &o, int &m, int right)
{
    int n2 = m / 2;
    n2 -= n2 % 8;
    if (right) { o += n2; m -= n2; } else { m = n2; }
}

__global__ void __launch_bounds__(8 << EXP_DEPTH) k_exp_sum(const double *__restrict__ scratch, int nd,
                                                            const int *__restrict__ cnt, double *__restrict__ sums)
{
    __shared__ double part[1 << EXP_DEPTH];
    const int i = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int j = lane & 7;
    const int g = threadIdx.x >> 3;                               // group id = path from the root, MSB first
    const unsigned gmask = 0xffu << (8 * ((lane >> 3)));
    const int c = cnt[i];
    if (c <= 0) {
        if (threadIdx.x == 0) sums[i] = 0.0;
        return;
    }
    int o = 0, m = c, level = 0;
    bool mine = true;
    for (; level < EXP_DEPTH; level++) {
        if (m <= 128) { mine = ((g & ((1 << (EXP_DEPTH - level)) - 1)) == 0); break; }
        np_split(o, m, (g >> (EXP_DEPTH - 1 - level)) & 1);
    }
    if (mine) {
        const double *base = scratch + i + (size_t)o * nd;
        const size_t p = (size_t)nd;
        const double s = np_pairwise_sum_group8([&](int e) { return base[(size_t)e * p]; }, m, j, gmask);
        if (j == 0) part[g] = s;
    }
    __syncthreads();
    // combine: node (level, q) lives in slot q << (EXP_DEPTH - level); it is internal iff every ancestor
    // and the node itself are longer than 128
    for (int lv = EXP_DEPTH - 1; lv >= 0; lv--) {
        const int q = threadIdx.x;
        if (q < (1 << lv)) {
            int oo = 0, mm = c;
            bool internal = true;
            for (int l = 0; l < lv && internal; l++) {
                if (mm <= 128) internal = false;
                else np_split(oo, mm, (q >> (lv - 1 - l)) & 1);
            }
            if (internal && mm > 128) {
                const int slot = q << (EXP_DEPTH - lv);
                part[slot] = __dadd_rn(part[slot], part[slot + (1 << (EXP_DEPTH - 1 - lv))]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[i] = part[0];
}

// masked per-diagonal sums of a band that is already on the device
static int expected_core(nlf_ctx *c, const double *d_band, int n, int nd, const unsigned char *d_valid, double *sum_out,
                         long long *cnt_out)
{
    cudaStream_t s = c->stream;
    const size_t cells = (size_t)n * nd;
    const int nseg = (n + EXP_SEG - 1) / EXP_SEG;
    double *d_scratch, *d_sums;
    int *d_cnt, *d_seg;
    NLF_CUDA(c->big_alloc((void **)&d_scratch, sizeof(double) * cells));
    NLF_CUDA(cudaMallocAsync(&d_cnt, sizeof(int) * nd, s));
    NLF_CUDA(cudaMallocAsync(&d_sums, sizeof(double) * nd, s));
    NLF_CUDA(cudaMallocAsync(&d_seg, sizeof(int) * (size_t)nseg * nd, s));
    const dim3 grid((nd + 63) / 64, nseg);
    {
        KLaunch k(c, P_NORM);
        k_exp_segcount<<<grid, 64, 0, s>>>(n, nd, d_valid, d_seg);
    }
    {
        KLaunch k(c, P_NORM);
        k_exp_segscan<<<(nd + 63) / 64, 64, 0, s>>>(nseg, nd, d_seg, d_cnt);
    }
    {
        KLaunch k(c, P_NORM);
        k_exp_scatter<<<grid, 64, 0, s>>>(d_band, n, nd, d_valid, d_seg, d_scratch);
    }
    {
        KLaunch k(c, P_NORM);
        k_exp_sum<<<nd, 8 << EXP_DEPTH, 0, s>>>(d_scratch, nd, d_cnt, d_sums);
    }
    NLF_CUDA(cudaGetLastError());
    std::vector<int> cnt(nd);
    NLF_CUDA(cudaMemcpyAsync(sum_out, d_sums, sizeof(double) * nd, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * nd, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < nd; i++) cnt_out[i] = cnt[i];
    c->big_free(d_scratch);
    NLF_CUDA(cudaFreeAsync(d_cnt, s));
    NLF_CUDA(cudaFreeAsync(d_sums, s));
    NLF_CUDA(cudaFreeAsync(d_seg, s));
    return NLF_OK;
}

NLF_EXPORT int nlf_expected_diagonals(nlf_ctx *ctx, const double *band, int n, int nd, const uint8_t *valid,
                                      double *sum_out, long long *cnt_out)
{
    if (!ctx || !band || !valid || !sum_out || !cnt_out || n <= 0 || nd <= 0)
        return fail(NLF_ERR_ARG, "nlf_expected_diagonals: bad arguments");
    if (nd > n) nd = n;
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    double *d_band;
    unsigned char *d_valid;
    const size_t cells = (size_t)n * nd;
    NLF_CUDA(c->big_alloc((void **)&d_band, sizeof(double) * cells));
    NLF_CUDA(cudaMallocAsync(&d_valid, n, s));
    NLF_CUDA(cudaMemcpyAsync(d_band, band, sizeof(double) * cells, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_valid, valid, n, cudaMemcpyHostToDevice, s));
    int rc = expected_core(c, d_band, n, nd, d_valid, sum_out, cnt_out);
    c->big_free(d_band);
    NLF_CUDA(cudaFreeAsync(d_valid, s));
    return rc;
}

// util._prepare_core straight from the resident pixel table: chromosome = global bins [lo, hi)
NLF_EXPORT int nlf_expected_from_pixels(nlf_ctx *ctx, nlf_pixels *px, int balance, long long lo, long long hi, int nd,
                                        double *sum_out, long long *cnt_out)
{
    if (!ctx || !px || !sum_out || !cnt_out || nd <= 0 || lo < 0 || hi <= lo)
        return fail(NLF_ERR_ARG, "nlf_expected_from_pixels: bad arguments");
    if (px->ctx != ctx) return fail(NLF_ERR_ARG, "nlf_expected_from_pixels: pixel table lives on another context");
    if (hi > px->nbins) return fail(NLF_ERR_ARG, "nlf_expected_from_pixels: bins [%lld,%lld) outside [0,%lld)", lo, hi, px->nbins);
    if (hi - lo > 0x7fffffffLL / 2) return fail(NLF_ERR_ARG, "nlf_expected_from_pixels: chromosome too long");
    const int n = (int)(hi - lo);
    if (nd > n) nd = n;
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    double *d_band;
    unsigned char *d_valid;
    NLF_CUDA(c->big_alloc((void **)&d_band, sizeof(double) * (size_t)n * nd));
    NLF_CUDA(cudaMallocAsync(&d_valid, n, s));
    int rc = nlf_pixels_chrom_band(px, balance, lo, hi, nd, d_band, d_valid);
    if (rc == NLF_OK) rc = expected_core(c, d_band, n, nd, d_valid, sum_out, cnt_out);
    c->big_free(d_band);
    NLF_CUDA(cudaFreeAsync(d_valid, s));
    return rc;
}
