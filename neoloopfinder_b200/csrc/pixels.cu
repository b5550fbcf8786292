// pixels.cu -- a sample's contact map resident in HBM as cooler's pixel table:
//   bin1_offset[nbins+1] (CSR index), bin2[nnz], count[nnz] for bin1 <= bin2, one weight per bin.
// Everything the reference obtains through `clr.matrix(balance=col).fetch(r1, r2)` and
// `clr.bins().fetch(r)[col]` on the neoloop-caller path is produced from it on the device:
//   * complexSV.reorganize_matrix (neoloop/assembly.py:408-492): the rearranged upper-triangular
//     fusion matrix of an SV assembly and its bias vector, block by block, with orientation flips;
//   * util._prepare_core (neoloop/util.py:403-428): the band of one chromosome for the masked
//     per-diagonal sums of calculate_expected;
//   * the gap-free band of a whole chromosome taken as a trivial assembly (Fusion.remove_gaps,
//     neoloop/callers.py:504-522, then Peakachu.__init__ :527-535) for the genome-wide call.
// Balanced values follow cooler's order: (count * w[row bin]) * w[column bin], NaN -> 0 where the
// reference clears NaN (assembly.py:403-404).  HBM-bound scatter kernels: one warp per source row,
// binary search to the first column of interest, coalesced reads of the row's pixels.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

using namespace nlf;

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long px_lower_bound(const int *__restrict__ bin2, long long a, long long b, long long key)
{
    // first p in [a, b) with bin2[p] >= key
    while (a < b) {
        long long m = (a + b) >> 1;
        if ((long long)bin2[m] < key) a = m + 1; else b = m;
    }
    return a;
}

__device__ __forceinline__ double px_value(int c, const double *__restrict__ w, int balance, long long rowbin, long long colbin)
{
    double v = (double)c;
    if (balance) {
        v = __dmul_rn(__dmul_rn(v, w[rowbin]), w[colbin]);
        if (v != v) v = 0.0;                      // M[np.isnan(M)] = 0 (assembly.py:403-404)
    }
    return v;
}

// rows must be sorted by bin2, bin1 <= bin2 < nbins: the lookups above rely on it
__global__ void k_px_validate(const long long *__restrict__ off, const int *__restrict__ bin2, long long nbins,
                              int *__restrict__ bad)
{
    const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= nbins) return;
    const int lane = threadIdx.x & 31;
    const long long a = off[s], b = off[s + 1];
    for (long long p = a + lane; p < b; p += 32) {
        const long long t = bin2[p];
        bool ok = (t >= s) && (t < nbins);
        if (p > a) ok = ok && ((long long)bin2[p - 1] < t);
        if (!ok) atomicExch(bad, 1);
    }
}

// ---------------------------------------------------------------------------------------------
// reorganize_matrix: warp per (assembly row r, column block bc)
// ---------------------------------------------------------------------------------------------
struct PxBlocks {
    const long long *src_lo, *src_hi;     // global bin range of every block
    const int *out_lo;                    // first assembly index of every block
    const unsigned char *rev;             // block traversed 3'->5'
    int nblk;
};

__device__ __forceinline__ int px_out_index(const PxBlocks &B, int b, long long t)
{
    return B.out_lo[b] + (int)(B.rev[b] ? (B.src_hi[b] - 1 - t) : (t - B.src_lo[b]));
}

__global__ void k_px_assemble(const long long *__restrict__ off, const int *__restrict__ bin2,
                              const int *__restrict__ cnt, const double *__restrict__ w, int balance, PxBlocks B,
                              const int *__restrict__ blk, int n, double *__restrict__ M)
{
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long long)n * B.nblk) return;
    const int lane = threadIdx.x & 31;
    const int r = (int)(wid / B.nblk), bc = (int)(wid % B.nblk);
    const int br = blk[r];
    const long long s = B.rev[br] ? (B.src_hi[br] - 1 - (r - B.out_lo[br])) : (B.src_lo[br] + (r - B.out_lo[br]));
    // stored pixels (s, t) with t >= s and t inside column block bc
    const long long t0 = max(B.src_lo[bc], s), t1 = B.src_hi[bc];
    if (t0 >= t1) return;
    const long long a = off[s], b = off[s + 1];
    long long p = px_lower_bound(bin2, a, b, t0) + lane;
    for (; p < b; p += 32) {
        const long long t = bin2[p];
        if (t >= t1) break;
        const int c = cnt[p];
        const int ct = px_out_index(B, bc, t);
        // the pixel seen as (row block br, column block bc)
        if (br <= bc && (br < bc || r <= ct)) M[(size_t)r * n + ct] = px_value(c, w, balance, s, t);
        // and, mirrored, as (row block bc, column block br): M[t-side][s-side] = (count*w[t])*w[s]
        if (bc <= br && !(bc == br && s == t) && (bc < br || ct <= r))
            M[(size_t)ct * n + r] = px_value(c, w, balance, t, s);
    }
}

__global__ void k_px_bias(const double *__restrict__ w, PxBlocks B, const int *__restrict__ blk, int n,
                          double *__restrict__ bias)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int br = blk[r];
    const long long s = B.rev[br] ? (B.src_hi[br] - 1 - (r - B.out_lo[br])) : (B.src_lo[br] + (r - B.out_lo[br]));
    double v = w[s];
    bias[r] = (v != v) ? 0.0 : v;                 // get_bias (assembly.py:387-397)
}

// ---------------------------------------------------------------------------------------------
// chromosome band for calculate_expected: band[(s-lo)*nd + (t-s)], pre-zeroed
// ---------------------------------------------------------------------------------------------
__global__ void k_px_band(const long long *__restrict__ off, const int *__restrict__ bin2,
                          const int *__restrict__ cnt, const double *__restrict__ w, int balance, long long lo,
                          long long hi, int nd, double *__restrict__ band, unsigned char *__restrict__ valid)
{
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long s = lo + wid;
    if (s >= hi) return;
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        const double ws = w[s];
        valid[wid] = (isfinite(ws) && ws > 0) ? 1 : 0;     // util.py:415
    }
    const long long t1 = min(hi, s + nd);
    const long long b = off[s + 1];
    for (long long p = off[s] + lane; p < b; p += 32) {    // rows start at t = s: no search needed
        const long long t = bin2[p];
        if (t >= t1) break;
        double v = px_value(cnt[p], w, balance, s, t);
        if (!isfinite(v)) v = 0.0;                         // masked by `valid` anyway
        band[(size_t)wid * nd + (t - s)] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// whole chromosome as a trivial assembly: gap columns, then the gap-free band
// ---------------------------------------------------------------------------------------------
__global__ void k_px_colflag(const long long *__restrict__ off, const int *__restrict__ bin2,
                             const int *__restrict__ cnt, const double *__restrict__ w, int balance, long long lo,
                             long long hi, unsigned char *__restrict__ flag)
{
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long s = lo + wid;
    if (s >= hi) return;
    const int lane = threadIdx.x & 31;
    const long long b = off[s + 1];
    for (long long p = off[s] + lane; p < b; p += 32) {
        const long long t = bin2[p];
        if (t >= hi) break;
        if (px_value(cnt[p], w, balance, s, t) != 0.0) flag[t - lo] = 1;   // column t has a non-zero entry
    }
}

__global__ void k_px_gapband(const long long *__restrict__ off, const int *__restrict__ bin2,
                             const int *__restrict__ cnt, const double *__restrict__ w, int balance, long long lo,
                             long long hi, const int *__restrict__ kept, const int *__restrict__ rank, int nk, int bw,
                             int pitch, double *__restrict__ band)
{
    const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= nk) return;
    const int lane = threadIdx.x & 31;
    const long long s = lo + kept[k];
    const long long b = off[s + 1];
    for (long long p = off[s] + lane; p < b; p += 32) {
        const long long t = bin2[p];
        if (t >= hi) break;
        const int rt = rank[t - lo];
        if (rt < 0) continue;
        const int d = rt - (int)k;
        if (d >= bw) break;                        // rank is monotone in t: nothing further fits the band
        double v = px_value(cnt[p], w, balance, s, t);
        if (!isfinite(v)) v = 0.0;                 // Peakachu.__init__ keeps finite values only (callers.py:531)
        band[(size_t)k * pitch + d] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
NLF_EXPORT int nlf_pixels_create(nlf_ctx *ctx, long long nbins, const int64_t *bin1_offset, const int32_t *bin2,
                                 const int32_t *count, const double *weights, nlf_pixels **out)
{
    if (!ctx || !bin1_offset || !weights || !out || nbins <= 0)
        return fail(NLF_ERR_ARG, "nlf_pixels_create: bad arguments");
    if (bin1_offset[0] != 0) return fail(NLF_ERR_ARG, "nlf_pixels_create: bin1_offset[0] must be 0");
    for (long long s = 0; s < nbins; s++)
        if (bin1_offset[s + 1] < bin1_offset[s])
            return fail(NLF_ERR_ARG, "nlf_pixels_create: bin1_offset decreases at bin %lld", s);
    const long long nnz = bin1_offset[nbins];
    if (nnz > 0 && (!bin2 || !count)) return fail(NLF_ERR_ARG, "nlf_pixels_create: NULL pixel columns");
    NLF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    nlf_pixels *px = new nlf_pixels();
    px->ctx = ctx; px->nbins = nbins; px->nnz = nnz;
    const size_t nz = (size_t)std::max(nnz, 1LL);
    int *d_bad = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&px->d_off, sizeof(long long) * (size_t)(nbins + 1)));
    step(cudaMalloc(&px->d_bin2, sizeof(int) * nz));
    step(cudaMalloc(&px->d_cnt, sizeof(int) * nz));
    step(cudaMalloc(&px->d_w, sizeof(double) * (size_t)nbins));
    step(cudaMalloc(&d_bad, sizeof(int)));
    if (e == cudaSuccess) {
        step(cudaMemcpyAsync(px->d_off, bin1_offset, sizeof(long long) * (size_t)(nbins + 1), cudaMemcpyHostToDevice, s));
        if (nnz > 0) {
            step(cudaMemcpyAsync(px->d_bin2, bin2, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, s));
            step(cudaMemcpyAsync(px->d_cnt, count, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, s));
        }
        step(cudaMemcpyAsync(px->d_w, weights, sizeof(double) * (size_t)nbins, cudaMemcpyHostToDevice, s));
        step(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
    }
    int bad = 0;
    if (e == cudaSuccess) {
        {
            KLaunch k(ctx, P_NORM);
            k_px_validate<<<(unsigned)((nbins + 7) / 8), 256, 0, s>>>(px->d_off, px->d_bin2, nbins, d_bad);
        }
        step(cudaGetLastError());
        step(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        step(cudaStreamSynchronize(s));
    }
    if (d_bad) cudaFree(d_bad);
    if (e != cudaSuccess || bad) {
        nlf_pixels_destroy(px);
        if (e != cudaSuccess) return fail(NLF_ERR_CUDA, "nlf_pixels_create: %s", cudaGetErrorString(e));
        return fail(NLF_ERR_ARG, "nlf_pixels_create: rows must hold strictly ascending bin2 with bin1 <= bin2 < nbins");
    }
    *out = px;
    return NLF_OK;
}

NLF_EXPORT int nlf_pixels_destroy(nlf_pixels *px)
{
    if (!px) return NLF_OK;
    cudaSetDevice(px->ctx->device);
    cudaStreamSynchronize(px->ctx->stream);
    cudaFree(px->d_off); cudaFree(px->d_bin2); cudaFree(px->d_cnt); cudaFree(px->d_w);
    delete px;
    return NLF_OK;
}

// fusion matrix + bias of an assembly into caller-provided device buffers (norm.cu owns nlf_assembly)
int nlf_pixels_fill_assembly(nlf_pixels *px, int balance, const long long *src_lo, const long long *src_hi,
                             const unsigned char *rev, const int *out_lo, int nblk, int n, const int *d_blk,
                             double *d_M, double *d_bias)
{
    nlf_ctx *c = px->ctx;
    cudaStream_t s = c->stream;
    // block table: one small allocation [src_lo | src_hi | out_lo | rev]
    const size_t bytes = sizeof(long long) * 2 * nblk + sizeof(int) * nblk + nblk;
    std::vector<unsigned char> host(bytes);
    memcpy(host.data(), src_lo, sizeof(long long) * nblk);
    memcpy(host.data() + sizeof(long long) * nblk, src_hi, sizeof(long long) * nblk);
    memcpy(host.data() + sizeof(long long) * 2 * nblk, out_lo, sizeof(int) * nblk);
    memcpy(host.data() + sizeof(long long) * 2 * nblk + sizeof(int) * nblk, rev, nblk);
    unsigned char *d_tab;
    NLF_CUDA(cudaMallocAsync(&d_tab, bytes, s));
    NLF_CUDA(cudaMemcpyAsync(d_tab, host.data(), bytes, cudaMemcpyHostToDevice, s));
    // (pageable sources are staged before cudaMemcpyAsync returns: `host` may go out of scope)
    PxBlocks B;
    B.src_lo = (const long long *)d_tab;
    B.src_hi = B.src_lo + nblk;
    B.out_lo = (const int *)(d_tab + sizeof(long long) * 2 * nblk);
    B.rev = d_tab + sizeof(long long) * 2 * nblk + sizeof(int) * nblk;
    B.nblk = nblk;
    cudaError_t e = cudaMemsetAsync(d_M, 0, sizeof(double) * (size_t)n * n, s);
    if (e == cudaSuccess) {
        {
            KLaunch k(c, P_NORM);
            const long long warps = (long long)n * nblk;
            k_px_assemble<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(px->d_off, px->d_bin2, px->d_cnt, px->d_w, balance, B,
                                                                       d_blk, n, d_M);
        }
        {
            KLaunch k(c, P_NORM);
            k_px_bias<<<(n + 127) / 128, 128, 0, s>>>(px->d_w, B, d_blk, n, d_bias);
        }
        e = cudaGetLastError();
    }
    cudaFreeAsync(d_tab, s);                       // stream ordered: after the kernels that read it
    NLF_CUDA(e);
    return NLF_OK;
}

// band + valid mask of the chromosome [lo, hi) on the device (expected.cu consumes them)
int nlf_pixels_chrom_band(nlf_pixels *px, int balance, long long lo, long long hi, int nd, double *d_band,
                          unsigned char *d_valid)
{
    nlf_ctx *c = px->ctx;
    cudaStream_t s = c->stream;
    const long long n = hi - lo;
    NLF_CUDA(cudaMemsetAsync(d_band, 0, sizeof(double) * (size_t)n * nd, s));
    {
        KLaunch k(c, P_NORM);
        k_px_band<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(px->d_off, px->d_bin2, px->d_cnt, px->d_w, balance, lo, hi, nd,
                                                           d_band, d_valid);
    }
    NLF_CUDA(cudaGetLastError());
    return NLF_OK;
}

// gap-free band of the chromosome [lo, hi) taken as a trivial assembly; kept = surviving bins
int nlf_pixels_export_band(nlf_pixels *px, int balance, long long lo, long long hi, int bw, int pitch,
                           double **d_band_out, std::vector<int> &kept)
{
    nlf_ctx *c = px->ctx;
    cudaStream_t s = c->stream;
    const int n = (int)(hi - lo);
    unsigned char *d_flag;
    NLF_CUDA(cudaMallocAsync(&d_flag, n, s));
    NLF_CUDA(cudaMemsetAsync(d_flag, 0, n, s));
    {
        KLaunch k(c, P_NORM);
        k_px_colflag<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(px->d_off, px->d_bin2, px->d_cnt, px->d_w, balance, lo, hi, d_flag);
    }
    NLF_CUDA(cudaGetLastError());
    std::vector<unsigned char> flag(n);
    NLF_CUDA(cudaMemcpyAsync(flag.data(), d_flag, n, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    NLF_CUDA(cudaFreeAsync(d_flag, s));
    kept.clear();
    for (int j = 0; j < n; j++)
        if (flag[j]) kept.push_back(j);
    if (5 * 2 + 1 > (int)kept.size()) {            // callers.py:514-516 (w = 5)
        kept.resize(n);
        for (int j = 0; j < n; j++) kept[j] = j;
    }
    const int nk = (int)kept.size();
    std::vector<int> rank(n, -1);
    for (int k = 0; k < nk; k++) rank[kept[k]] = k;
    int *d_kept, *d_rank;
    NLF_CUDA(cudaMallocAsync(&d_kept, sizeof(int) * nk, s));
    NLF_CUDA(cudaMallocAsync(&d_rank, sizeof(int) * n, s));
    NLF_CUDA(cudaMemcpyAsync(d_kept, kept.data(), sizeof(int) * nk, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_rank, rank.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    double *band = nullptr;
    cudaError_t e = c->big_alloc((void **)&band, sizeof(double) * (size_t)nk * pitch);
    if (e == cudaSuccess) e = cudaMemsetAsync(band, 0, sizeof(double) * (size_t)nk * pitch, s);
    if (e == cudaSuccess) {
        KLaunch k(c, P_BANDIFY);
        k_px_gapband<<<(unsigned)((nk + 7) / 8), 256, 0, s>>>(px->d_off, px->d_bin2, px->d_cnt, px->d_w, balance, lo, hi,
                                                               d_kept, d_rank, nk, bw, pitch, band);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);       // rank / kept host vectors are read by the copies above
    cudaFreeAsync(d_kept, s);
    cudaFreeAsync(d_rank, s);
    if (e != cudaSuccess) {
        if (band) c->big_free(band);
        NLF_CUDA(e);
    }
    *d_band_out = band;
    return NLF_OK;
}
