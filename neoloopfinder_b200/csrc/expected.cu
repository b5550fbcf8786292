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

// one thread per diagonal: gather the valid entries in row order.  scratch[k * nd + i]: the k-th
// valid entry of diagonal i (neighbouring diagonals are neighbours in memory -> coalesced).
__global__ void k_exp_compact(const double *__restrict__ band, int n, int nd, const unsigned char *__restrict__ valid,
                              double *__restrict__ scratch, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    int k = 0;
    const int rows = n - i;                       // diagonal i has n-i entries
    for (int r = 0; r < rows; r++) {
        if (valid[r] && valid[r + i]) {
            scratch[(size_t)k * nd + i] = band[(size_t)r * nd + i];
            k++;
        }
    }
    cnt[i] = k;
}

// 8 lanes per diagonal: numpy pairwise sum of the gathered entries
__global__ void __launch_bounds__(256) k_exp_sum(const double *__restrict__ scratch, int nd, const int *__restrict__ cnt,
                                                 double *__restrict__ sums)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7;
    const int i = blockIdx.x * 32 + warp * 4 + g;
    if (i >= nd) return;
    const unsigned gmask = 0xffu << (8 * g);
    const int c = cnt[i];
    double s = 0.0;
    if (c > 0) {
        const double *base = scratch + i;
        const size_t p = (size_t)nd;
        s = np_pairwise_sum_group8([&](int e) { return base[(size_t)e * p]; }, c, j, gmask);
    }
    if (j == 0) sums[i] = s;
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
    double *d_band, *d_scratch, *d_sums;
    unsigned char *d_valid;
    int *d_cnt;
    const size_t cells = (size_t)n * nd;
    NLF_CUDA(c->big_alloc((void **)&d_band, sizeof(double) * cells));
    NLF_CUDA(c->big_alloc((void **)&d_scratch, sizeof(double) * cells));
    NLF_CUDA(cudaMallocAsync(&d_valid, n, s));
    NLF_CUDA(cudaMallocAsync(&d_cnt, sizeof(int) * nd, s));
    NLF_CUDA(cudaMallocAsync(&d_sums, sizeof(double) * nd, s));
    NLF_CUDA(cudaMemcpyAsync(d_band, band, sizeof(double) * cells, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_valid, valid, n, cudaMemcpyHostToDevice, s));
    {
        KLaunch k(c, P_NORM);
        k_exp_compact<<<(nd + 63) / 64, 64, 0, s>>>(d_band, n, nd, d_valid, d_scratch, d_cnt);
    }
    {
        KLaunch k(c, P_NORM);
        k_exp_sum<<<(nd + 31) / 32, 256, 0, s>>>(d_scratch, nd, d_cnt, d_sums);
    }
    NLF_CUDA(cudaGetLastError());
    std::vector<int> cnt(nd);
    NLF_CUDA(cudaMemcpyAsync(sum_out, d_sums, sizeof(double) * nd, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * nd, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < nd; i++) cnt_out[i] = cnt[i];
    c->big_free(d_band);
    c->big_free(d_scratch);
    NLF_CUDA(cudaFreeAsync(d_valid, s));
    NLF_CUDA(cudaFreeAsync(d_cnt, s));
    NLF_CUDA(cudaFreeAsync(d_sums, s));
    return NLF_OK;
}
