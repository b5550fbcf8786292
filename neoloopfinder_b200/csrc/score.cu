// score.cu -- scoring path of neoloop-caller on B200 (sm_100a):
//   K1 bandify        dense gap-free matrix -> finite band        (callers.py:527-535)
//   K5 diag means + isotonic expected + candidate gate             (callers.py:540-570)
//   K7 window gates, O/E, gaussian, min-max -> float32 features    (callers.py:583-612, util.py:470-524)
//   K8 random-forest traversal, ordered fp64 vote                  (callers.py:622)
//   K9 threshold + compaction of the Donuts                        (callers.py:623-628)
// All floating-point arithmetic that reaches a result is IEEE fp64 with explicit round-to-nearest
// intrinsics (no FMA contraction; the file is also compiled with -fmad=false), in the exact
// operation order of numpy / scipy.ndimage / sklearn as restated in oracle/c/nlf_oracle.c.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <utility>

#include "common.cuh"

namespace nlf {
thread_local char g_err[512] = "";
}
using namespace nlf;

// ---------------------------------------------------------------------------------------------
// data layout in HBM (per job = one gap-free matrix of n bins)
//   band  [n][pitch]  fp64   band[r][d] = G[r, r+d] for 0 <= d < upper+14, finite, else 0
//   prob  [n][ndp]    fp64   prob[x][d-dlo] = forest probability of pixel (x, x+d); NaN = not scored
//   slot  [n][ndp]    int32  feature slot (group*32+lane) of a scored pixel, -1 otherwise (optional)
//   e     [upper+1]   fp64   isotonic expected of Peakachu.get_candidates
// feature groups (shared by all jobs of a chunk)
//   featT [group][F][32] fp32  features of 32 scored pixels (slots group*32 .. +31, densely packed in
//                              the order the feature kernel's tiles claim them), feature-major so that
//                              the forest kernel's per-lane column is bank-conflict free
//   pix   [slot]               {job, (x << dshift) | d} of the pixel in that slot (dshift bits hold any d <= upper)
//   nanmask [group]            bit l set: slot group*32+l has NaN features
// ---------------------------------------------------------------------------------------------
struct JobDev {
    const double *band;
    double *prob;
    int *slot;
    double *means;      // [upper+1] diagonal means (NaN where the diagonal is too short)
    double *e;          // [upper+1]
    int *status;        // [0] error code of the job (0 ok)
    unsigned long long *counts;  // [0] candidates, [1] scored, [2] donuts
    int *don_job, *don_i, *don_j; // Donut outputs: ONE list per batch (capacity don_cap), tagged by job
    double *don_p, *don_v;
    unsigned long long *don_count;
    long long don_cap;
    int job;
    int n;
    int ntx;            // number of x tiles
    int dshift;         // PixRef packing: smallest s with (1 << s) > upper; rows are limited to 2^(32-s) per job
};

// pixel behind feature slot s: .x = job, .y = (matrix row x << JobDev::dshift) | diagonal d
typedef uint2 PixRef;

__constant__ double c_gw[16];   // gaussian weights, radius 4: c_gw[0..8]

// =============================================================================================
// K1 bandify
// =============================================================================================
__global__ void k_bandify(const double *__restrict__ G, int n, int bw, int pitch, double *__restrict__ band,
                          int *__restrict__ status)
{
    int r = blockIdx.x;                                   // rows in grid.x: a job may have more than 65 535 of them
    int d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= pitch) return;
    double v = 0.0;
    if (d < bw && r + d < n) {
        v = G[(size_t)r * n + r + d];
        if (!isfinite(v)) v = 0.0;
    }
    band[(size_t)r * pitch + d] = v;
    // Peakachu keeps 13 sub-diagonals too (C-R > -14); the path feeds np.triu matrices, so any
    // finite non-zero there means the caller broke the contract: report instead of mis-scoring.
    if (d >= 1 && d < 14 && r - d >= 0) {
        double u = G[(size_t)r * n + r - d];
        if (u != 0.0 && isfinite(u)) atomicExch(status, NLF_ERR_LOWER_TRI);
    }
}

// band given by the host with its own pitch -> canonical pitch, finite filter
__global__ void k_band_copy(const double *__restrict__ src, int n, int src_pitch, int bw, int pitch,
                            double *__restrict__ band)
{
    int r = blockIdx.x;
    int d = blockIdx.y * blockDim.x + threadIdx.x;
    if (d >= pitch) return;
    double v = 0.0;
    if (d < bw && d < src_pitch && r + d < n) {
        v = src[(size_t)r * src_pitch + d];
        if (!isfinite(v)) v = 0.0;
    }
    band[(size_t)r * pitch + d] = v;
}

// =============================================================================================
// K5a per-diagonal means in numpy's pairwise order, 8 lanes per diagonal.
// numpy splits n > 128 at n/2 - (n/2)%8, so every leaf starts at a multiple of 8 and only the last
// leaf has a ragged tail: inside a leaf, element e feeds accumulator e%8.  Lane j of an 8-lane group
// therefore owns accumulator j of every leaf (rows off+8i+j), the combine ((r0+r1)+(r2+r3))+((r4+r5)+
// (r6+r7)) is a 3-step xor butterfly (fp add is commutative, the association is numpy's), and the
// recursion over leaves is replayed identically by the 8 lanes.  A warp serves 4 neighbouring
// diagonals: each load instruction touches 8 rows x 32 contiguous bytes.
// =============================================================================================
__global__ void __launch_bounds__(256) k_diag_means(const JobDev *__restrict__ jobs, int lower, int upper, int pitch)
{
    const JobDev J = jobs[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7;
    const int d = lower + blockIdx.x * 32 + warp * 4 + g;
    if (d > upper) return;
    const unsigned gmask = 0xffu << (8 * g);
    const int len = J.n - d;
    double m = __longlong_as_double(0x7ff8000000000000LL);
    if (len > 5) {
        const double *base = J.band + d;
        const size_t p = (size_t)pitch;
        auto ld = [&](int e) { return base[(size_t)e * p]; };
        auto leaf = [&](int off, int cnt) -> double {
            if (cnt < 8) {
                double r = -0.0;
                for (int i = 0; i < cnt; i++) r = __dadd_rn(r, ld(off + i));
                return r;
            }
            const int groups = cnt >> 3;
            double r = ld(off + j);
            int i = 1;
            for (; i + 4 <= groups; i += 4) {
                double a0 = ld(off + 8 * i + j), a1 = ld(off + 8 * (i + 1) + j), a2 = ld(off + 8 * (i + 2) + j),
                       a3 = ld(off + 8 * (i + 3) + j);
                r = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r, a0), a1), a2), a3);
            }
            for (; i < groups; i++) r = __dadd_rn(r, ld(off + 8 * i + j));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
            r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
            for (int e = groups * 8; e < cnt; e++) r = __dadd_rn(r, ld(off + e));
            return r;
        };
        double s;
        if (len <= 128) {
            s = leaf(0, len);
        } else {
            int st_off[48], st_n[48];      // st_n < 0 marks a "combine" entry
            double vs[26];
            int sp = 0, vp = 0;
            st_off[sp] = 0; st_n[sp] = len; sp++;
            while (sp) {
                sp--;
                const int o = st_off[sp], mm = st_n[sp];
                if (mm < 0) {
                    const double r = vs[--vp];
                    const double l = vs[--vp];
                    vs[vp++] = __dadd_rn(l, r);
                } else if (mm <= 128) {
                    vs[vp++] = leaf(o, mm);
                } else {
                    int n2 = mm / 2;
                    n2 -= n2 % 8;
                    st_off[sp] = 0; st_n[sp] = -1; sp++;
                    st_off[sp] = o + n2; st_n[sp] = mm - n2; sp++;
                    st_off[sp] = o; st_n[sp] = n2; sp++;
                }
            }
            s = vs[0];
        }
        m = __ddiv_rn(s, (double)len);
    }
    if (j == 0) J.means[d] = m;
}

// K5b isotonic (decreasing) fit of the means: scipy's PAVA (pava_pybind.cpp) on the reversed
// sequence, one thread per job (<= upper-lower+1 points).  e[d] = fit[min(d, last)].
__global__ void k_isotonic(const JobDev *__restrict__ jobs, int lower, int upper)
{
    // work arrays in shared memory (the fit is a short, strictly sequential scan: latency matters)
    extern __shared__ __align__(16) unsigned char iso_raw[];
    const JobDev J = jobs[blockIdx.x];
    const int last = min(upper, J.n - 6);    // diagonals lower..last have more than 5 entries
    const long long m = (long long)last - lower + 1;
    double *x = reinterpret_cast<double *>(iso_raw);
    double *w = x + (upper + 2);
    int *r = reinterpret_cast<int *>(w + (upper + 2));
    if (m <= 0) {
        if (threadIdx.x == 0) atomicExch(J.status, NLF_ERR_EMPTY_FIT);
        return;
    }
    for (long long i = threadIdx.x; i < m; i += blockDim.x) { x[i] = J.means[last - i]; w[i] = 1.0; }   // reversed
    __syncthreads();
    if (threadIdx.x == 0 && m > 1) {
        r[0] = 0; r[1] = 1;
        int b = 0;
        double xb_prev = x[0], wb_prev = w[0];
        for (int i = 1; i < (int)m; ++i) {
            b++;
            double xb = x[i], wb = w[i], sb = 0;
            if (xb_prev >= xb) {
                b--;
                sb = __dadd_rn(__dmul_rn(wb_prev, xb_prev), __dmul_rn(wb, xb));
                wb = __dadd_rn(wb, wb_prev);
                xb = __ddiv_rn(sb, wb);
                while (i < (int)m - 1 && xb >= x[i + 1]) {
                    i++;
                    sb = __dadd_rn(sb, __dmul_rn(w[i], x[i]));
                    wb = __dadd_rn(wb, w[i]);
                    xb = __ddiv_rn(sb, wb);
                }
                while (b > 0 && x[b - 1] >= xb) {
                    b--;
                    sb = __dadd_rn(sb, __dmul_rn(w[b], x[b]));
                    wb = __dadd_rn(wb, w[b]);
                    xb = __ddiv_rn(sb, wb);
                }
            }
            x[b] = xb_prev = xb;
            w[b] = wb_prev = wb;
            r[b + 1] = i + 1;
        }
        int f = (int)m - 1;
        for (int k = b; k >= 0; --k) {
            const int t = r[k];
            const double xk = x[k];
            for (int i = f; i >= t; --i) x[i] = xk;
            f = t - 1;
        }
    }
    __syncthreads();
    // un-reverse and publish e[d] with out_of_bounds='clip'
    for (int d = threadIdx.x; d <= upper; d += blockDim.x) {
        double v = __longlong_as_double(0x7ff8000000000000LL);
        if (d >= lower) {
            long long k = d - lower;
            if (k > m - 1) k = m - 1;
            v = x[m - 1 - k];
        }
        J.e[d] = v;
    }
}

// shared-memory / mbarrier / TMA helpers (split-phase barriers in k_features, TMA pipeline in the forest kernels)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    const unsigned addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// =============================================================================================
// K7 feature kernel.
// Block = S warps (S = 2W+1); tile = TX matrix rows x 32 consecutive diagonals.  Warp a owns
// window row a, lane l owns diagonal d0+l, so every shared-memory access of a warp is a run of
// consecutive 8-byte words (conflict free) and no lane diverges inside the arithmetic.
//   phase 0  band tile (+halo) -> smem raw R and observed/expected E = R / exp[|col-row|]
//   phase 1  gates per pixel: candidate test, count_nonzero, lower-left mean (numpy order)
//   phase 2  per matrix row x: axis-0 gaussian pass shared by the 32 windows of the row,
//            then per window row the axis-1 pass, min/max, normalise, cast to float32.
// =============================================================================================
template <int W, int TXP>
struct FeatCfg {
    static constexpr int S = 2 * W + 1;
    static constexpr int F = S * S;
    static constexpr int TX = TXP;              // matrix rows per tile
    static constexpr int TR = TX + 2 * W;       // tile rows incl. halo
    static constexpr int TC = 32 + 4 * W;       // tile band columns incl. halo
    static constexpr int VC = 32 + 2 * W;       // columns of the axis-0 result
    static constexpr int NT = 32 * S;
};

__device__ __forceinline__ int reflect_idx(int i, int n)
{
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}

template <int W, int TXP>
__global__ void __launch_bounds__(FeatCfg<W, TXP>::NT, 2)
k_features(const JobDev *__restrict__ jobs, const int *__restrict__ tile_base, int njobs, int tile_offset,
           int lower, int upper, int pitch, int ndp, int ndt, const double *__restrict__ expv, int nexp, PixRef *__restrict__ pix,
           float *__restrict__ featT, unsigned *__restrict__ slot_counter, unsigned max_slots,
           int *__restrict__ overflow, unsigned *__restrict__ nanmask, unsigned slot_base)
{
    using C = FeatCfg<W, TXP>;
    constexpr int S = C::S, F = C::F, TX = C::TX, TR = C::TR, TC = C::TC, VC = C::VC, NT = C::NT;
    static_assert(TC <= 64, "one 64-bit non-zero mask per tile row");
    // raw and observed/expected tile: dynamic shared memory (2 * TR * TC doubles; above 48 KB per CTA in
    // total for the taller tiles, see launch_features)
    extern __shared__ __align__(16) double feat_dyn[];
    double (*Rt)[TC] = reinterpret_cast<double (*)[TC]>(feat_dyn);
    double (*Et)[TC] = reinterpret_cast<double (*)[TC]>(feat_dyn + TR * TC);
    __shared__ unsigned long long nzmask[TR];     // bit cc of row rr: Rt[rr][cc] != 0
    __shared__ double Vx[S][VC];
    __shared__ double2 red_mm[2][S][32];      // {row min, row max}, double buffered
    __shared__ double fin_mn[2][32], fin_den[2][32], fin_rden[2][32];   // window min, max - min, 1 / (max - min)
    __shared__ unsigned passmask[TX];
    __shared__ unsigned s_rowbase[TX];       // first feature slot of every tile row
    __shared__ unsigned s_slots_ok;
    __shared__ unsigned char s_items[2 * TX];
    __shared__ int s_nitems;
    __shared__ __align__(8) unsigned long long s_bar[2], s_fin[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- tile -> (job, x0, d0)
    int t = blockIdx.x + tile_offset;
    int lo = 0, hi = njobs;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_base[mid] <= t) lo = mid; else hi = mid;
    }
    const int job = lo;
    const JobDev J = jobs[job];
    if (*J.status != 0) return;
    const int lt = t - tile_base[job];
    const int dlo = max(lower, W + 2);
    const int x0 = (lt / ndt) * TX;
    const int d0 = dlo + (lt % ndt) * 32;
    const int n = J.n;
    const int bw = upper + 14;

    // ---- phase 0: load tile
    for (int k = tid; k < TR * TC; k += NT) {
        int rr = k / TC, cc = k % TC;
        int r = x0 - W + rr;
        int dd = d0 - 2 * W + cc;
        double v = 0.0;
        if (r >= 0 && r < n && dd >= 0 && dd < bw) v = J.band[(size_t)r * pitch + dd];
        Rt[rr][cc] = v;
        int ad = dd < 0 ? -dd : dd;
        Et[rr][cc] = (ad < nexp) ? __ddiv_rn(v, expv[ad]) : v;
    }
    __syncthreads();
    // one 64-bit mask of the non-zero cells per tile row (two ballots); count_nonzero of any window row is
    // then a shift, a mask and a popcount
    for (int rr = warp; rr < TR; rr += S) {
        const unsigned lo = __ballot_sync(0xffffffffu, Rt[rr][lane] != 0.0);
        const unsigned hi = __ballot_sync(0xffffffffu, (lane + 32 < TC) && Rt[rr][(lane + 32 < TC) ? lane + 32 : 0] != 0.0);
        if (lane == 0) nzmask[rr] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();

    // ---- phase 1: gates; warp handles tile row xi, lane = diagonal
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int xi = warp; xi < TX; xi += S) {
        int x = x0 + xi, d = d0 + lane, y = x + d;
        bool in_band = (x < n) && (d <= upper);
        bool ok = in_band && (x - W >= 0) && (y + W + 1 <= n) && (n - d > 1);
        if (ok) {
            double e = J.e[d];
            double v = Rt[xi + W][lane + 2 * W];
            ok = (e > 0) && (__ddiv_rn(v, e) > 0);
        }
        if (ok) {
            int nz = 0;
#pragma unroll
            for (int a = 0; a < S; a++)
                nz += __popcll((nzmask[xi + a] >> (lane + 2 * W - a)) & ((1ull << S) - 1ull));
            ok = !((double)nz < (double)F * 0.1);
        }
        if (ok) {
            // window[:w,:w].mean(): numpy pairwise order over the C-flattened w x w block
            auto ld = [&](int k) { int a = k / W, b = k % W; return Rt[xi + a][lane + 2 * W + b - a]; };
            double ll = __ddiv_rn(np_leaf_sum(ld, 0, W * W), (double)(W * W));
            ok = (ll > 0);
            if (ok) ok = (__ddiv_rn(Rt[xi + W][lane + 2 * W], ll) > 0.1);
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) passmask[xi] = m;
        if (in_band) {
            size_t o = (size_t)x * ndp + (d - dlo);
            J.prob[o] = qnan;
            if (J.slot) J.slot[o] = -1;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // claim one contiguous range of feature slots for the whole tile (one atomic per tile)
        unsigned cnt = 0;
        for (int xi = 0; xi < TX; xi++) cnt += __popc(passmask[xi]);
        unsigned base = cnt ? atomicAdd(slot_counter, cnt) : 0u;
        unsigned ok = 1u;
        if (cnt && (base + cnt > max_slots || base + cnt < base)) { atomicExch(overflow, 1); ok = 0u; }
        for (int xi = 0; xi < TX; xi++) { s_rowbase[xi] = base; base += __popc(passmask[xi]); }
        s_slots_ok = ok;
    }
    __syncthreads();
    if (!s_slots_ok) return;

    // lanes whose window reaches beyond the expected table stay un-normalised (util.py:480-481)
    unsigned raw_lanes = 0;
    {
        bool raw = (d0 + lane + 2 * W) >= nexp;
        raw_lanes = __ballot_sync(0xffffffffu, raw);
    }

    // ---- phase 2: rows.  Work items = (tile row, O/E or raw source) that have at least one window.
    // Per item: axis-1 pass -> row min/max to shared memory -> mbarrier ARRIVE (s_bar) -> axis-0 pass of
    // the NEXT item (independent work that hides the barrier latency) -> mbarrier WAIT (s_fin) -> normalise,
    // store.  Between the two barriers ONE warp (round robin over the items) folds the 2W+1 row extrema
    // of every window into min, max - min and its reciprocal and publishes them, so the other warps
    // do not repeat that reduction and division.
    if (tid == 0) {
        int k = 0;
        for (int xi = 0; xi < TX; xi++) {
            unsigned pm = passmask[xi];
            if (pm & ~raw_lanes) s_items[k++] = (unsigned char)(xi * 2);
            if (pm & raw_lanes) s_items[k++] = (unsigned char)(xi * 2 + 1);
        }
        s_nitems = k;
        mbar_init(&s_bar[0], NT);
        mbar_init(&s_bar[1], NT);
        mbar_init(&s_fin[0], 32);
        mbar_init(&s_fin[1], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nitems = s_nitems;
    auto stage_a = [&](int item) {
        // axis-0 pass, warp a -> Vx[a][*]  (row a is produced and consumed by warp a only)
        const int xi = item >> 1;
        const double (*T)[TC] = (item & 1) ? Rt : Et;
        const int a = warp;
        int ra[9];
#pragma unroll
        for (int k = 0; k < 9; k++) ra[k] = reflect_idx(a + k - 4, S);
        for (int cc = lane; cc < VC; cc += 32) {
            // src(a', cc) = T[xi + a'][cc - a' + 2W]
            double v = __dmul_rn(T[xi + ra[4]][cc - ra[4] + 2 * W], c_gw[4]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double l = T[xi + ra[j]][cc - ra[j] + 2 * W];
                double r = T[xi + ra[8 - j]][cc - ra[8 - j] + 2 * W];
                v = __dadd_rn(v, __dmul_rn(__dadd_rn(l, r), c_gw[j]));
            }
            Vx[a][cc] = v;
        }
        __syncwarp();
    };
    if (nitems > 0) stage_a(s_items[0]);
    for (int it = 0; it < nitems; it++) {
        const int item = s_items[it];
        const int xi = item >> 1;
        const unsigned pm = passmask[xi];
        const unsigned mm = (item & 1) ? (pm & raw_lanes) : (pm & ~raw_lanes);
        const int x = x0 + xi;
        const unsigned rb = it & 1u;
        // dense packing: the k-th passing pixel of the row takes slot rowbase + k
        const unsigned my_slot = s_rowbase[xi] + __popc(pm & ((1u << lane) - 1u));
        float *fdst = featT + (size_t)(my_slot >> 5) * F * 32 + (my_slot & 31u);
        // stage B: axis-1 pass for window row a = warp, window of lane
        double in[S], out[S];
#pragma unroll
        for (int b = 0; b < S; b++) in[b] = Vx[warp][lane + b];
        __syncwarp();                                  // Vx[warp] may now be overwritten (next stage A)
        double rmin, rmax;
        bool has_nan = false;
#pragma unroll
        for (int b = 0; b < S; b++) {
            double v = __dmul_rn(in[b], c_gw[4]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // pair (b-4+j, b+4-j), outermost first
                const int il = (b - 4 + j) < 0 ? -(b - 4 + j) - 1 : (b - 4 + j);
                const int ir = (b + 4 - j) >= S ? 2 * S - 1 - (b + 4 - j) : (b + 4 - j);
                v = __dadd_rn(v, __dmul_rn(__dadd_rn(in[il], in[ir]), c_gw[j]));
            }
            out[b] = v;
            has_nan |= (v != v);
            if (b == 0) { rmin = v; rmax = v; }
            else { rmin = (v < rmin) ? v : rmin; rmax = (v > rmax) ? v : rmax; }
        }
        if (has_nan) { rmin = qnan; rmax = qnan; }
        red_mm[rb][warp][lane] = make_double2(rmin, rmax);
        mbar_arrive(&s_bar[rb]);                       // release: my row's min/max are published
        const unsigned par = (it >> 1) & 1u;
        if (warp == it % S) {
            mbar_wait(&s_bar[rb], par);                // acquire: all window rows are in
            double2 m0 = red_mm[rb][0][lane];
            double mn = m0.x, mx = m0.y;
            bool nn = (mn != mn);
#pragma unroll
            for (int a = 1; a < S; a++) {
                const double2 m = red_mm[rb][a][lane];
                nn |= (m.x != m.x);
                mn = (m.x < mn) ? m.x : mn;
                mx = (m.y > mx) ? m.y : mx;
            }
            if (nn) { mn = qnan; mx = qnan; }
            const double den = __dsub_rn(mx, mn);
            const bool fast_ok = (den > 1e-290) && (den < 1e300);
            fin_mn[rb][lane] = mn;
            fin_den[rb][lane] = den;
            fin_rden[rb][lane] = fast_ok ? __ddiv_rn(1.0, den) : 0.0;    // 0 marks "always divide"
            mbar_arrive(&s_fin[rb]);
        }
        if (it + 1 < nitems) stage_a(s_items[it + 1]);
        mbar_wait(&s_fin[rb], par);                    // acquire: the window extrema are published
        if ((mm >> lane) & 1u) {
            bool f_nan = false;
            const double mn = fin_mn[rb][lane], den = fin_den[rb][lane], rden = fin_rden[rb][lane];
            // float32(fl64(num/den)) without 13 fp64 divisions: p = fl64(num * fl64(1/den)) is within
            // 2.01 ulp64 of the exact quotient, so it rounds to the same float32 unless it sits within
            // 8 ulp64 of a float32 rounding boundary (low 29 mantissa bits ~ 0x10000000) or below the
            // float32 normal range (p < 2^-125, p != 0 up to doubles that round to 0 anyway); rows with
            // such a value (about 1 cell in 3e7) are redone with the exact division.  In this branch den
            // is finite and positive, so every value is finite, non-negative and no NaN can appear.
            if (rden != 0.0) {
                bool risky = false;
#pragma unroll
                for (int b = 0; b < S; b++) {
                    const double f = __dmul_rn(__dsub_rn(out[b], mn), rden);
                    const unsigned flo = (unsigned)__double2loint(f), fhi = (unsigned)__double2hiint(f);
                    risky |= ((flo & 0x1fffffffu) - 0x0ffffff8u <= 16u) | (fhi - 1u < 0x381fffffu);
                    fdst[(size_t)(warp * S + b) * 32] = __double2float_rn(f);
                }
                if (risky) {
#pragma unroll
                    for (int b = 0; b < S; b++)
                        fdst[(size_t)(warp * S + b) * 32] = __double2float_rn(__ddiv_rn(__dsub_rn(out[b], mn), den));
                }
            } else {
#pragma unroll
                for (int b = 0; b < S; b++) {
                    const double f = __ddiv_rn(__dsub_rn(out[b], mn), den);
                    f_nan |= (f != f);
                    fdst[(size_t)(warp * S + b) * 32] = __double2float_rn(f);
                }
            }
            // windows with NaN features (constant window, or a zero in the expected table) take the
            // general forest kernel, which implements sklearn's missing_go_to_left routing
            if (f_nan) atomicOr(&nanmask[my_slot >> 5], 1u << (my_slot & 31u));
            if (warp == 0) {
                pix[my_slot] = make_uint2((unsigned)job, ((unsigned)x << J.dshift) | (unsigned)(d0 + lane));
                if (J.slot) J.slot[(size_t)x * ndp + (d0 + lane - dlo)] = (int)(slot_base + my_slot);
            }
        }
    }
}

// =============================================================================================
// K8 forest.  One warp per feature group, lane = pixel.  Every lane walks all trees in
// estimator order and adds the class-1 leaf fractions sequentially in fp64 (the order in which
// sklearn accumulates, _forest.py:704-717), then divides by the number of trees.
// Node = 8 bytes {float thr (largest float <= sklearn's float64 threshold), u32 packed}:
//   packed bits 0-7 feature (255 = leaf), bit 8 missing_go_to_left, bits 9-31 right child.
// The left child of node k is k+1 (sklearn grows trees depth-first).
// =============================================================================================
struct ForestDev {
    const uint2 *nodes;       // all trees concatenated
    const double *leaf;       // class-1 fraction per node (only read at leaves)
    const int *tree_off;      // [T+1]
    int n_trees;
    int n_features;
};

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
k_forest(ForestDev fr, const PixRef *__restrict__ pix, const float *__restrict__ featT,
         const unsigned *__restrict__ slot_counter, const JobDev *__restrict__ jobs, int lower, int W, int ndp)
{
    extern __shared__ float fs[];     // [WARPS][F][32]
    const int F = fr.n_features;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *my = fs + (size_t)warp * F * 32;
    const unsigned nslots = *slot_counter;
    const unsigned ngroups = (nslots + 31u) >> 5;
    const int dlo = max(lower, W + 2);
    const int T = fr.n_trees;
    for (unsigned g = blockIdx.x * WARPS + warp; g < ngroups; g += gridDim.x * WARPS) {
        const float4 *src = reinterpret_cast<const float4 *>(featT + (size_t)g * F * 32);
        float4 *dst = reinterpret_cast<float4 *>(my);
        for (int k = lane; k < F * 8; k += 32) dst[k] = src[k];
        __syncwarp();
        const unsigned slot = g * 32u + lane;
        if (slot < nslots) {
            double acc = 0.0;
            int t = 0;
            int base = fr.tree_off[0];
            int node = 0;
            while (t < T) {
                uint2 nd = __ldg(&fr.nodes[base + node]);
                unsigned f = nd.y & 0xffu;
                if (f == 0xffu) {
                    acc = __dadd_rn(acc, __ldg(&fr.leaf[base + node]));
                    t++;
                    node = 0;
                    if (t < T) base = fr.tree_off[t];
                } else {
                    float xv = my[f * 32 + lane];
                    float thr = __uint_as_float(nd.x);
                    bool left = (xv != xv) ? ((nd.y >> 8) & 1u) : (xv <= thr);
                    node = left ? node + 1 : (int)(nd.y >> 9);
                }
            }
            const PixRef px = pix[slot];
            const JobDev J = jobs[px.x];
            J.prob[(size_t)(px.y >> J.dshift) * ndp + ((int)(px.y & ((1u << J.dshift) - 1u)) - dlo)] = __ddiv_rn(acc, (double)T);
        }
        __syncwarp();
    }
}

// tree (0..ntree-1) that owns part-local node k; off[] = part-local tree offsets, off[ntree] = part_nodes
__device__ __forceinline__ int tree_of_node(const int *off, int ntree, int k)
{
    int lo = 0, hi = ntree;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= k) lo = mid; else hi = mid;
    }
    return lo;
}

// K8 (cooperative, dynamic trees) -- the production forest kernel.
// One CTA = 16 warps share ONE feature tile (32 pixels x F floats, TMA bulk copy, double buffered)
// and the part of the forest held in shared memory.  Lane l of every warp works for pixel l:
// it grabs the pixel's next unstarted tree from a shared-memory counter, walks it to the leaf,
// publishes the leaf value in leafv[tree][l] and grabs again, so all 16 lanes-per-pixel stay busy
// until the part's trees are exhausted (no static imbalance, no lock-step between lanes).
// After a block barrier one warp adds the part's leaf values in tree order (fp64, sequential =
// sklearn's accumulation order) while the others already traverse the next tile.
// Pixels whose features contain NaN are skipped here (nanmask) and scored by k_forest_nan.
template <int NW>
__global__ void __launch_bounds__(32 * NW, 1)
k_forest_dyn(const uint2 *__restrict__ nodes2, const int *__restrict__ tree_off, int t0, int t1, int T_total, int F,
             const PixRef *__restrict__ pix, const unsigned *__restrict__ nanmask, const float *__restrict__ featT,
             const unsigned *__restrict__ slot_counter, const JobDev *__restrict__ jobs, int lower, int W, int ndp,
             double *__restrict__ accbuf, int first, int last)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[2];
    const int part_base = tree_off[t0];
    const int part_nodes = tree_off[t1] - part_base;
    const int ntree = t1 - t0;
    const unsigned tile_bytes = (unsigned)F * 128u;
    const unsigned leafv_bytes = (unsigned)ntree * 256u;
    // byte offsets in smem_raw: [tile0][tile1][leafv0][leafv1][cnt0 cnt1][nodes + dummy][roots]
    const unsigned o_leafv = 2 * tile_bytes;
    const unsigned o_cnt = o_leafv + 2 * leafv_bytes;
    const unsigned o_nodes = o_cnt + 256u;
    const unsigned o_root = o_nodes + (unsigned)((part_nodes + 2) & ~1) * 8u;
#define SM_U2(off) (*reinterpret_cast<uint2 *>(smem_raw + (off)))
#define SM_U32(off) (*reinterpret_cast<unsigned *>(smem_raw + (off)))
#define SM_F32(off) (*reinterpret_cast<float *>(smem_raw + (off)))
#define SM_F64(off) (*reinterpret_cast<double *>(smem_raw + (off)))
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k <= ntree; k += blockDim.x)
        SM_U32(o_root + 4u * k) = (unsigned)((k < ntree) ? tree_off[t0 + k] - part_base : part_nodes);
    if (tid < 64) SM_U32(o_cnt + 4u * tid) = 0u;
    __syncthreads();
    for (int k = tid; k < part_nodes; k += blockDim.x) {
        uint2 nd = nodes2[part_base + k];
        if ((int)nd.y < 0) {
            // {flag | feature<<23 | miss<<22 | tree-local right}  ->  {flag | right byte offset | feature*128>>7}
            unsigned r = (nd.y & 0x3fffffu) +
                         SM_U32(o_root + 4u * tree_of_node(reinterpret_cast<const int *>(smem_raw + o_root), ntree, k));
            unsigned feat = (nd.y >> 23) & 0xffu;
            // bits 0-7 feature, bits 8-30 right-child byte offset / 8 is not needed: store the byte
            // offset itself in bits 8-30 (< 2^18 for any part that fits in shared memory)
            nd.y = 0x80000000u | ((o_nodes + (r << 3)) << 8) | feat;
        }
        SM_U2(o_nodes + 8u * k) = nd;
    }
    if (tid == 0) {
        SM_U2(o_nodes + 8u * part_nodes) = make_uint2(0u, 0u);
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int k = tid; k <= ntree; k += blockDim.x) SM_U32(o_root + 4u * k) = o_nodes + 8u * SM_U32(o_root + 4u * k);
    __syncthreads();
    const unsigned nslots = *slot_counter;
    const unsigned ngroups = (nslots + 31u) >> 5;
    const int dlo = max(lower, W + 2);
    unsigned g = blockIdx.x;
    if (g < ngroups && tid == 0) {
        mbar_expect_tx(&mbar[0], tile_bytes);
        tma_load_1d(smem_raw, featT + (size_t)g * F * 32, tile_bytes, &mbar[0]);
    }
    unsigned nanm = (g < ngroups) ? nanmask[g] : 0u;
    for (unsigned it = 0; g < ngroups; it++, g += gridDim.x) {
        const unsigned buf = it & 1u;
        const unsigned gn = g + gridDim.x;
        if (gn < ngroups && tid == 0) {
            // tile[buf^1] was last read during iteration it-1, which ended with a block barrier
            mbar_expect_tx(&mbar[buf ^ 1u], tile_bytes);
            tma_load_1d(smem_raw + (size_t)(buf ^ 1u) * tile_bytes, featT + (size_t)gn * F * 32, tile_bytes, &mbar[buf ^ 1u]);
        }
        const unsigned nanm_n = (gn < ngroups) ? nanmask[gn] : 0u;
        mbar_wait(&mbar[buf], (it >> 1) & 1u);
        const bool active = (g * 32u + lane < nslots) && !((nanm >> lane) & 1u);
        if (active) {
            const unsigned my = buf * tile_bytes + (unsigned)lane * 4u;
            const unsigned lv = o_leafv + buf * leafv_bytes + (unsigned)lane * 8u;
            unsigned *cnt = reinterpret_cast<unsigned *>(smem_raw + o_cnt + buf * 128u + (unsigned)lane * 4u);
            int t = (int)atomicAdd(cnt, 1u);
            while (t < ntree) {
                unsigned na = SM_U32(o_root + 4u * (unsigned)t);
                uint2 nd = SM_U2(na);
                while ((int)nd.y < 0) {
                    const float xv = SM_F32(my + ((nd.y & 0xffu) << 7));
                    na = (xv <= __uint_as_float(nd.x)) ? na + 8u : ((nd.y >> 8) & 0x7fffffu);
                    nd = SM_U2(na);
                }
                SM_U2(lv + 256u * (unsigned)t) = nd;
                t = (int)atomicAdd(cnt, 1u);
            }
        }
        __syncthreads();
        if (warp == (int)(it % NW)) {
            SM_U32(o_cnt + buf * 128u + (unsigned)lane * 4u) = 0u;       // next use is two barriers away
            if (active) {
                const unsigned lv = o_leafv + buf * leafv_bytes + (unsigned)lane * 8u;
                double acc = first ? 0.0 : accbuf[(size_t)g * 32 + lane];
                for (int tt = 0; tt < ntree; tt++) acc = __dadd_rn(acc, SM_F64(lv + 256u * (unsigned)tt));
                if (last) {
                    const PixRef px = pix[g * 32u + lane];
                    const JobDev J = jobs[px.x];
                    J.prob[(size_t)(px.y >> J.dshift) * ndp + ((int)(px.y & ((1u << J.dshift) - 1u)) - dlo)] = __ddiv_rn(acc, (double)T_total);
                } else {
                    accbuf[(size_t)g * 32 + lane] = acc;
                }
            }
        }
        nanm = nanm_n;
    }
#undef SM_U2
#undef SM_U32
#undef SM_F32
#undef SM_F64
}

// shared-window loads of the forest traversal (addresses are absolute shared addresses; node
// addresses carry a +0x40000 bias, see k_forest_pipe)
__device__ __forceinline__ uint2 lds_node(unsigned biased)
{
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2 + -262144];" : "=r"(v.x), "=r"(v.y) : "r"(biased));
    return v;
}
__device__ __forceinline__ unsigned scaled_add4(unsigned a, unsigned b)   // 4*a + b in one instruction
{
    unsigned r;
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ float lds_f32(unsigned addr)
{
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// K8 (cooperative, dynamic trees, barrier-free pipeline).  Same work decomposition as
// k_forest_dyn, but consecutive tiles overlap: there is no block barrier.  Every tile buffer has
// a `full` mbarrier (completed by the TMA copy) and a `done` mbarrier (every thread arrives once
// it finds no unstarted tree left for that tile).  Threads that run out of trees move on to the
// next tile at once; the tile's summing warp (round robin) waits for `done`, adds the leaf values
// in tree order, resets the tree counters and only then issues the TMA copy that refills the
// buffer two tiles later -- which is also what protects leafv/cnt from early reuse.
template <int NW, bool WARP_GRAB, int ILP = 1>
__global__ void __launch_bounds__(32 * NW, 1)
k_forest_pipe(const uint2 *__restrict__ nodes2, const int *__restrict__ tree_off, int t0, int t1, int T_total, int F,
              const PixRef *__restrict__ pix, const unsigned *__restrict__ nanmask, const float *__restrict__ featT,
              const unsigned *__restrict__ slot_counter, const JobDev *__restrict__ jobs, int lower, int W, int ndp,
              double *__restrict__ accbuf, int first, int last)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[2], done[2];
    const int part_base = tree_off[t0];
    const int part_nodes = tree_off[t1] - part_base;
    const int ntree = t1 - t0;
    const unsigned tile_bytes = (unsigned)F * 128u;
    const unsigned leafv_bytes = (unsigned)ntree * 256u;
    const unsigned o_leafv = 2 * tile_bytes;
    const unsigned o_cnt = o_leafv + 2 * leafv_bytes;
    const unsigned o_nodes = o_cnt + 256u;
    const unsigned o_root = o_nodes + (unsigned)((part_nodes + 2) & ~1) * 8u;
#define SM_U2(off) (*reinterpret_cast<uint2 *>(smem_raw + (off)))
#define SM_U32(off) (*reinterpret_cast<unsigned *>(smem_raw + (off)))
#define SM_F32(off) (*reinterpret_cast<float *>(smem_raw + (off)))
#define SM_F64(off) (*reinterpret_cast<double *>(smem_raw + (off)))
    // Staged internal node: y = 1<<31 | (shared address of the right child >> 3) << 16 | feature << 5, so
    // that y >> 13 is that address plus the constant 0x40000 (folded into the load's immediate
    // offset) and (y & 0x1fe0) << 2 is the byte offset of the feature row in the tile.
    const unsigned sbase = smem_u32(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < ntree + 8; k += blockDim.x)
        SM_U32(o_root + 4u * k) = (unsigned)((k < ntree) ? tree_off[t0 + k] - part_base : part_nodes);
    if (tid < 64) SM_U32(o_cnt + 4u * tid) = 0u;
    __syncthreads();
    for (int k = tid; k < part_nodes; k += blockDim.x) {
        uint2 nd = nodes2[part_base + k];
        if ((int)nd.y < 0) {
            unsigned r = (nd.y & 0x3fffffu) +
                         SM_U32(o_root + 4u * tree_of_node(reinterpret_cast<const int *>(smem_raw + o_root), ntree, k));
            unsigned feat = (nd.y >> 23) & 0xffu;
            nd.y = 0x80000000u | (((sbase + o_nodes + (r << 3)) >> 3) << 16) | (feat << 5);
        }
        SM_U2(o_nodes + 8u * k) = nd;
    }
    if (tid == 0) {
        SM_U2(o_nodes + 8u * part_nodes) = make_uint2(0u, 0u);
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&done[0], 32 * NW);
        mbar_init(&done[1], 32 * NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int k = tid; k < ntree + 8; k += blockDim.x) SM_U32(o_root + 4u * k) = 0x40000u + sbase + o_nodes + 8u * SM_U32(o_root + 4u * k);
    __syncthreads();
    const unsigned nslots = *slot_counter;
    const unsigned ngroups = (nslots + 31u) >> 5;
    const int dlo = max(lower, W + 2);
    const unsigned stride = gridDim.x;
    unsigned g = blockIdx.x;
    if (tid == 0) {
        if (g < ngroups) {
            mbar_expect_tx(&full[0], tile_bytes);
            tma_load_1d(smem_raw, featT + (size_t)g * F * 32, tile_bytes, &full[0]);
        }
        if (g + stride < ngroups) {
            mbar_expect_tx(&full[1], tile_bytes);
            tma_load_1d(smem_raw + tile_bytes, featT + (size_t)(g + stride) * F * 32, tile_bytes, &full[1]);
        }
    }
    unsigned nanm = (g < ngroups) ? nanmask[g] : 0u;
    for (unsigned it = 0; g < ngroups; it++, g += stride) {
        const unsigned buf = it & 1u;
        const unsigned par = (it >> 1) & 1u;
        const unsigned gn = g + stride;
        const unsigned nanm_n = (gn < ngroups) ? nanmask[gn] : 0u;
        mbar_wait(&full[buf], par);
        const bool active = (g * 32u + lane < nslots) && !((nanm >> lane) & 1u);
        if (WARP_GRAB) {
            // the warp takes whole trees: its 32 pixels walk the same tree together, so the upper
            // levels are broadcast loads and neighbouring pixels tend to leave at similar depths
            const unsigned my = sbase + buf * tile_bytes + (unsigned)lane * 4u;
            const unsigned lv = o_leafv + buf * leafv_bytes + (unsigned)lane * 8u;
            unsigned *cnt = reinterpret_cast<unsigned *>(smem_raw + o_cnt + buf * 128u);
            int t = 0;
            if (lane == 0) t = (int)atomicAdd(cnt, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            while (t < ntree) {
                if (active) {
                    unsigned na = SM_U32(o_root + 4u * (unsigned)t);
                    uint2 nd = lds_node(na);
                    while ((int)nd.y < 0) {
                        const float xv = lds_f32(scaled_add4(nd.y & 0x1fe0u, my));
                        na = (xv <= __uint_as_float(nd.x)) ? na + 8u : (nd.y >> 13);
                        nd = lds_node(na);
                    }
                    SM_U2(lv + 256u * (unsigned)t) = nd;
                }
                if (lane == 0) t = (int)atomicAdd(cnt, 1u);
                t = __shfl_sync(0xffffffffu, t, 0);
            }
        } else if (ILP > 1) {
            // ILP trees per lane in flight: that many independent shared-memory load chains per warp
            if (active) {
                const unsigned my = sbase + buf * tile_bytes + (unsigned)lane * 4u;
                const unsigned lv = o_leafv + buf * leafv_bytes + (unsigned)lane * 8u;
                unsigned *cnt = reinterpret_cast<unsigned *>(smem_raw + o_cnt + buf * 128u + (unsigned)lane * 4u);
                int t = (int)atomicAdd(cnt, (unsigned)ILP);
                while (t < ntree) {
                    // root table entries [ntree, ntree + 8) point at the all-zero sentinel leaf behind the part
                    unsigned na[ILP];
                    uint2 nd[ILP];
#pragma unroll
                    for (int c = 0; c < ILP; c++) na[c] = SM_U32(o_root + 4u * (unsigned)(t + c));
#pragma unroll
                    for (int c = 0; c < ILP; c++) nd[c] = lds_node(na[c]);
                    while (true) {
                        unsigned any = nd[0].y;
#pragma unroll
                        for (int c = 1; c < ILP; c++) any |= nd[c].y;
                        if ((int)any >= 0) break;
#pragma unroll
                        for (int c = 0; c < ILP; c++) {
                            if ((int)nd[c].y < 0) {
                                const float xv = lds_f32(scaled_add4(nd[c].y & 0x1fe0u, my));
                                na[c] = (xv <= __uint_as_float(nd[c].x)) ? na[c] + 8u : (nd[c].y >> 13);
                                nd[c] = lds_node(na[c]);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < ILP; c++)
                        if (t + c < ntree) SM_U2(lv + 256u * (unsigned)(t + c)) = nd[c];
                    t = (int)atomicAdd(cnt, (unsigned)ILP);
                }
            }
        } else
        if (active) {
            const unsigned my = sbase + buf * tile_bytes + (unsigned)lane * 4u;
            const unsigned lv = o_leafv + buf * leafv_bytes + (unsigned)lane * 8u;
            unsigned *cnt = reinterpret_cast<unsigned *>(smem_raw + o_cnt + buf * 128u + (unsigned)lane * 4u);
            int t = (int)atomicAdd(cnt, 1u);
            while (t < ntree) {
                unsigned na = SM_U32(o_root + 4u * (unsigned)t);
                uint2 nd = lds_node(na);
                while ((int)nd.y < 0) {
                    const float xv = lds_f32(scaled_add4(nd.y & 0x1fe0u, my));
                    na = (xv <= __uint_as_float(nd.x)) ? na + 8u : (nd.y >> 13);
                    nd = lds_node(na);
                }
                SM_U2(lv + 256u * (unsigned)t) = nd;
                t = (int)atomicAdd(cnt, 1u);
            }
        }
        mbar_arrive(&done[buf]);                      // release: this thread's leaf values are published
        if (warp == (int)(it % NW)) {
            mbar_wait(&done[buf], par);               // acquire: every thread is through with this tile
            if (active) {
                const unsigned lv = o_leafv + buf * leafv_bytes + (unsigned)lane * 8u;
                double acc = first ? 0.0 : accbuf[(size_t)g * 32 + lane];
                for (int tt = 0; tt < ntree; tt++) acc = __dadd_rn(acc, SM_F64(lv + 256u * (unsigned)tt));
                if (last) {
                    const PixRef px = pix[g * 32u + lane];
                    const JobDev J = jobs[px.x];
                    J.prob[(size_t)(px.y >> J.dshift) * ndp + ((int)(px.y & ((1u << J.dshift) - 1u)) - dlo)] = __ddiv_rn(acc, (double)T_total);
                } else {
                    accbuf[(size_t)g * 32 + lane] = acc;
                }
            }
            SM_U32(o_cnt + buf * 128u + (unsigned)lane * 4u) = 0u;
            __syncwarp();
            const unsigned g2 = g + 2 * stride;
            if (lane == 0 && g2 < ngroups) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads/writes before the async-proxy write
                mbar_expect_tx(&full[buf], tile_bytes);
                tma_load_1d(smem_raw + (size_t)buf * tile_bytes, featT + (size_t)g2 * F * 32, tile_bytes, &full[buf]);
            }
        }
        nanm = nanm_n;
    }
#undef SM_U2
#undef SM_U32
#undef SM_F32
#undef SM_F64
}

// =============================================================================================
// K8, single pass over a COMPACT forest (the production kernel when the forest fits).
// Only internal nodes are stored (8 bytes: float threshold + packed word); a leaf is a pointer, never a node:
//   * a node with one leaf child keeps its internal child as the IMPLICIT next node (the node that follows
//     it in memory) and points at the leaf explicitly; bit S says that the implicit child is the right one;
//   * a node with two leaf children is TERMINAL: it carries both leaf references;
//   * leaf references index a table of the forest's DISTINCT leaf values (class fractions repeat a lot),
//     stored right behind the nodes, so "explicit child" is one address space.
// For the bundled-shape forest (100 trees, 32 048 sklearn nodes) that is 15 974 nodes + 307 values = 127 KB
// instead of 256 KB: the whole forest sits in the shared memory of ONE CTA, next to a CTA of the feature
// kernel on the same SM -- one forest pass per pixel, no partial sums through HBM, and the fp64 pipe
// (features of the next chunk) and the LSU (traversal of this chunk) work at the same time.
// Global word y: bit 31 CONT (non-terminal), bit 30 S (CONT) / 1 (terminal), bits 29-22 feature,
// bits 21-0: CONT: explicit child slot (node index, or n_int + leaf reference); terminal: left reference |
// right reference << 11.  Staged in shared memory (CONT nodes): 1<<31 | (address of the explicit child >> 3)
// << 16 | feature << 5 | S, as in k_forest_pipe.  Work decomposition, TMA tile pipeline, tree-ordered fp64
// summation: as k_forest_pipe; the traversal stores 16-bit leaf references, the summing warp looks them up.
// =============================================================================================
struct Forest3Dev {
    const uint2 *nodes;       // [n_int]
    const double *leafv;      // [n_leaf] distinct class-1 fractions, all in [0, 2)
    const unsigned *root;     // [n_trees] slot of the root (node index, or n_int + leaf reference for a one-leaf tree)
    int n_int, n_leaf, n_trees;
};

static inline size_t forest3_smem(int F, int T, int n_int, int n_leaf)
{
    return 2 * (size_t)F * 128 + 2 * (size_t)T * 64 + 256 + 8 * (size_t)n_int + 8 * (size_t)n_leaf +
           4 * (size_t)((T + 8 + 3) & ~3) + 16;
}

template <int NW, int ILP>
__global__ void __launch_bounds__(32 * NW, 1)
k_forest_one(Forest3Dev fr, int F, const PixRef *__restrict__ pix, const unsigned *__restrict__ nanmask,
             const float *__restrict__ featT, const unsigned *__restrict__ slot_counter, const JobDev *__restrict__ jobs,
             int lower, int W, int ndp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[2], done[2];
    const int T = fr.n_trees;
    const unsigned tile_bytes = (unsigned)F * 128u;
    const unsigned lid_bytes = (unsigned)T * 64u;
    const unsigned o_lid = 2 * tile_bytes;
    const unsigned o_cnt = o_lid + 2 * lid_bytes;
    const unsigned o_nodes = o_cnt + 256u;
    const unsigned o_leaf = o_nodes + 8u * (unsigned)fr.n_int;
    const unsigned o_root = o_leaf + 8u * (unsigned)fr.n_leaf;
#define SM_U2(off) (*reinterpret_cast<uint2 *>(smem_raw + (off)))
#define SM_U32(off) (*reinterpret_cast<unsigned *>(smem_raw + (off)))
#define SM_U16(off) (*reinterpret_cast<unsigned short *>(smem_raw + (off)))
#define SM_F64(off) (*reinterpret_cast<double *>(smem_raw + (off)))
    const unsigned sbase = smem_u32(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < fr.n_int; k += blockDim.x) {
        uint2 nd = fr.nodes[k];
        if ((int)nd.y < 0)
            nd.y = 0x80000000u | (((sbase + o_nodes + ((nd.y & 0x3fffffu) << 3)) >> 3) << 16) | (((nd.y >> 22) & 0xffu) << 5) |
                   ((nd.y >> 30) & 1u);
        SM_U2(o_nodes + 8u * k) = nd;
    }
    for (int k = tid; k < fr.n_leaf; k += blockDim.x) SM_F64(o_leaf + 8u * k) = fr.leafv[k];
    // root table; entries [T, T + 8) point at leaf slot 0 (a chain that starts there ends at once)
    for (int k = tid; k < T + 8; k += blockDim.x)
        SM_U32(o_root + 4u * k) = 0x40000u + sbase + o_nodes + 8u * (k < T ? fr.root[k] : (unsigned)fr.n_int);
    if (tid < 64) SM_U32(o_cnt + 4u * tid) = 0u;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&done[0], 32 * NW);
        mbar_init(&done[1], 32 * NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned leaf_biased = 0x40000u + sbase + o_leaf;
    const unsigned nslots = *slot_counter;
    const unsigned ngroups = (nslots + 31u) >> 5;
    const int dlo = max(lower, W + 2);
    const unsigned stride = gridDim.x;
    unsigned g = blockIdx.x;
    if (tid == 0) {
        if (g < ngroups) {
            mbar_expect_tx(&full[0], tile_bytes);
            tma_load_1d(smem_raw, featT + (size_t)g * F * 32, tile_bytes, &full[0]);
        }
        if (g + stride < ngroups) {
            mbar_expect_tx(&full[1], tile_bytes);
            tma_load_1d(smem_raw + tile_bytes, featT + (size_t)(g + stride) * F * 32, tile_bytes, &full[1]);
        }
    }
    unsigned nanm = (g < ngroups) ? nanmask[g] : 0u;
    for (unsigned it = 0; g < ngroups; it++, g += stride) {
        const unsigned buf = it & 1u;
        const unsigned par = (it >> 1) & 1u;
        const unsigned gn = g + stride;
        const unsigned nanm_n = (gn < ngroups) ? nanmask[gn] : 0u;
        mbar_wait(&full[buf], par);
        const bool active = (g * 32u + lane < nslots) && !((nanm >> lane) & 1u);
        if (active) {
            const unsigned my = sbase + buf * tile_bytes + (unsigned)lane * 4u;
            const unsigned lid = o_lid + buf * lid_bytes + (unsigned)lane * 2u;
            unsigned *cnt = reinterpret_cast<unsigned *>(smem_raw + o_cnt + buf * 128u + (unsigned)lane * 4u);
            int t = (int)atomicAdd(cnt, (unsigned)ILP);
            while (t < T) {
                unsigned na[ILP];
                uint2 nd[ILP];
#pragma unroll
                for (int c = 0; c < ILP; c++) na[c] = SM_U32(o_root + 4u * (unsigned)(t + c));
#pragma unroll
                for (int c = 0; c < ILP; c++) nd[c] = lds_node(na[c]);
                while (true) {
                    unsigned any = nd[0].y;
#pragma unroll
                    for (int c = 1; c < ILP; c++) any |= nd[c].y;
                    if ((int)any >= 0) break;
#pragma unroll
                    for (int c = 0; c < ILP; c++) {
                        if ((int)nd[c].y < 0) {
                            const float xv = lds_f32(scaled_add4(nd[c].y & 0x1fe0u, my));
                            const bool imp = (xv <= __uint_as_float(nd[c].x)) != (bool)(nd[c].y & 1u);
                            na[c] = imp ? na[c] + 8u : (nd[c].y >> 13);
                            nd[c] = lds_node(na[c]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < ILP; c++) {
                    unsigned ref;
                    if (nd[c].y & 0x40000000u) {
                        // terminal node: one more comparison picks one of its two leaves
                        const float xv = lds_f32(my + ((nd[c].y >> 15) & 0x7f80u));
                        ref = (xv <= __uint_as_float(nd[c].x)) ? (nd[c].y & 0x7ffu) : ((nd[c].y >> 11) & 0x7ffu);
                    } else {
                        ref = (na[c] - leaf_biased) >> 3;       // the address IS the leaf
                    }
                    if (t + c < T) SM_U16(lid + 64u * (unsigned)(t + c)) = (unsigned short)ref;
                }
                t = (int)atomicAdd(cnt, (unsigned)ILP);
            }
        }
        mbar_arrive(&done[buf]);                      // release: this thread's leaf references are published
        if (warp == (int)(it % NW)) {
            mbar_wait(&done[buf], par);               // acquire: every thread is through with this tile
            if (active) {
                const unsigned lid = o_lid + buf * lid_bytes + (unsigned)lane * 2u;
                double acc = 0.0;
                for (int tt = 0; tt < T; tt++)
                    acc = __dadd_rn(acc, SM_F64(o_leaf + 8u * (unsigned)SM_U16(lid + 64u * (unsigned)tt)));
                const PixRef px = pix[g * 32u + lane];
                const JobDev J = jobs[px.x];
                J.prob[(size_t)(px.y >> J.dshift) * ndp + ((int)(px.y & ((1u << J.dshift) - 1u)) - dlo)] = __ddiv_rn(acc, (double)T);
            }
            SM_U32(o_cnt + buf * 128u + (unsigned)lane * 4u) = 0u;
            __syncwarp();
            const unsigned g2 = g + 2 * stride;
            if (lane == 0 && g2 < ngroups) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads/writes before the async-proxy write
                mbar_expect_tx(&full[buf], tile_bytes);
                tma_load_1d(smem_raw + (size_t)buf * tile_bytes, featT + (size_t)g2 * F * 32, tile_bytes, &full[buf]);
            }
        }
        nanm = nanm_n;
    }
#undef SM_U2
#undef SM_U32
#undef SM_U16
#undef SM_F64
}

// pixels with NaN features: general traversal from global memory with missing_go_to_left routing
__global__ void k_forest_nan(ForestDev fr, const PixRef *__restrict__ pix, const unsigned *__restrict__ nanmask,
                             const float *__restrict__ featT, const unsigned *__restrict__ slot_counter,
                             const JobDev *__restrict__ jobs, int lower, int W, int ndp)
{
    const unsigned nslots = *slot_counter;
    const unsigned ngroups = (nslots + 31u) >> 5;
    const int lane = threadIdx.x & 31;
    const int dlo = max(lower, W + 2);
    const int F = fr.n_features;
    for (unsigned g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < ngroups; g += gridDim.x * (blockDim.x >> 5)) {
        const unsigned nm = nanmask[g];
        if (!nm) continue;
        const unsigned slot = g * 32u + lane;
        if (!((nm >> lane) & 1u) || slot >= nslots) continue;
        const float *x = featT + (size_t)g * F * 32 + lane;
        double acc = 0.0;
        for (int t = 0; t < fr.n_trees; t++) {
            int base = fr.tree_off[t], node = 0;
            while (true) {
                uint2 nd = fr.nodes[base + node];
                unsigned f = nd.y & 0xffu;
                if (f == 0xffu) break;
                float xv = x[(size_t)f * 32];
                bool left = (xv != xv) ? ((nd.y >> 8) & 1u) : (xv <= __uint_as_float(nd.x));
                node = left ? node + 1 : (int)(nd.y >> 9);
            }
            acc = __dadd_rn(acc, fr.leaf[base + node]);
        }
        const PixRef px = pix[slot];
        const JobDev J = jobs[px.x];
        J.prob[(size_t)(px.y >> J.dshift) * ndp + ((int)(px.y & ((1u << J.dshift) - 1u)) - dlo)] = __ddiv_rn(acc, (double)fr.n_trees);
    }
}

// forest on plain row-major float32 rows (nlf_forest_predict): thread per row
__global__ void k_forest_rows(ForestDev fr, const float *__restrict__ X, long long N, double *__restrict__ out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *x = X + (size_t)i * fr.n_features;
    double acc = 0.0;
    for (int t = 0; t < fr.n_trees; t++) {
        int base = fr.tree_off[t], node = 0;
        while (true) {
            uint2 nd = __ldg(&fr.nodes[base + node]);
            unsigned f = nd.y & 0xffu;
            if (f == 0xffu) break;
            float xv = x[f];
            bool left = (xv != xv) ? ((nd.y >> 8) & 1u) : (xv <= __uint_as_float(nd.x));
            node = left ? node + 1 : (int)(nd.y >> 9);
        }
        acc = __dadd_rn(acc, fr.leaf[base + node]);
    }
    out[i] = __ddiv_rn(acc, (double)fr.n_trees);
}

// =============================================================================================
// K9 threshold: count scored pixels, append pixels with proba >= thre (callers.py:625-628)
// with their contact value M[i,j].  Also counts the candidates of get_candidates.
// =============================================================================================
__global__ void __launch_bounds__(256) k_threshold(const JobDev *__restrict__ jobs, int lower, int upper, int W, int pitch, int ndp,
                                                   double thre)
{
    // HBM-bound scan of the probability plane: one warp per matrix row, 128-bit loads (a row of `prob` starts on a
    // 256-byte boundary), no index division; the rare pixels above the threshold are appended with one atomic each.
    const JobDev J = jobs[blockIdx.y];
    if (*J.status != 0) return;
    const int dlo = max(lower, W + 2);
    const int nd = upper - dlo + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    unsigned long long scored = 0, donuts = 0;
    for (int x = blockIdx.x * nwarp + warp; x < J.n; x += gridDim.x * nwarp) {
        const double2 *row = reinterpret_cast<const double2 *>(J.prob + (size_t)x * ndp);
        for (int k = lane; 2 * k < nd; k += 32) {
            const double2 pp = row[k];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int dd = 2 * k + h;
                const double p = h ? pp.y : pp.x;
                if (dd < nd && p == p) {
                    scored++;
                    if (p >= thre) {
                        donuts++;
                        unsigned long long pos = atomicAdd(J.don_count, 1ULL);
                        if ((long long)pos < J.don_cap) {
                            const int d = dlo + dd;
                            J.don_job[pos] = J.job;
                            J.don_i[pos] = x;
                            J.don_j[pos] = x + d;
                            J.don_p[pos] = p;
                            J.don_v[pos] = J.band[(size_t)x * pitch + d];
                        }
                    }
                }
            }
        }
    }
    // warp reduce the counters
    for (int o = 16; o; o >>= 1) {
        scored += __shfl_down_sync(0xffffffffu, scored, o);
        donuts += __shfl_down_sync(0xffffffffu, donuts, o);
    }
    if (lane == 0) {
        if (scored) atomicAdd(&J.counts[1], scored);
        if (donuts) atomicAdd(&J.counts[2], donuts);
    }
}

// candidates of Peakachu.get_candidates: all (x, d), lower <= d <= upper, n-d > 1, e[d] > 0,
// band/e > 0.  flag = 1 -> count only; also used for ordered compaction below.
__device__ __forceinline__ bool is_candidate(const JobDev &J, int x, int d, int pitch)
{
    if (x + d >= J.n || J.n - d <= 1) return false;
    double e = J.e[d];
    if (!(e > 0)) return false;
    return __ddiv_rn(J.band[(size_t)x * pitch + d], e) > 0;
}

__global__ void __launch_bounds__(256) k_count_candidates(const JobDev *__restrict__ jobs, int lower, int upper, int pitch)
{
    // HBM-bound scan of the band: one warp per matrix row, 128-bit loads (pitch is a multiple of 4 doubles and the
    // band is 256-byte aligned), the exact division only for cells that are not zero.
    const JobDev J = jobs[blockIdx.y];
    if (*J.status != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int n = J.n;
    unsigned long long c = 0;
    for (int x = blockIdx.x * nwarp + warp; x < n; x += gridDim.x * nwarp) {
        const double2 *row = reinterpret_cast<const double2 *>(J.band + (size_t)x * pitch);
        for (int k = (lower >> 1) + lane; 2 * k <= upper; k += 32) {
            const double2 vv = row[k];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int d = 2 * k + h;
                const double v = h ? vv.y : vv.x;
                if (d >= lower && d <= upper && v != 0.0 && x + d < n && n - d > 1) {
                    const double e = J.e[d];
                    c += (e > 0) && (__ddiv_rn(v, e) > 0);
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if (lane == 0 && c) atomicAdd(&J.counts[0], c);
}

// Ordered (diagonal-major, row-ascending) listing used by the API/test accessors.
// mode 0: candidates; mode 1: scored pixels (prob not NaN).  One block per diagonal:
// pass A counts, pass B (after an exclusive scan of the counts) writes.
__global__ void k_diag_count(JobDev J, int mode, int lower, int upper, int W, int pitch, int ndp, int *cnt)
{
    int d = lower + blockIdx.x;
    const int dlo = max(lower, W + 2);
    int c = 0;
    for (int x = threadIdx.x; x + d < J.n; x += blockDim.x) {
        bool f;
        if (mode == 0) f = is_candidate(J, x, d, pitch);
        else { f = false; if (d >= dlo) { double p = J.prob[(size_t)x * ndp + (d - dlo)]; f = (p == p); } }
        c += f;
    }
    __shared__ int sh[32];
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) s += sh[k];
        cnt[blockIdx.x] = s;
    }
}

__global__ void k_diag_fill(JobDev J, int mode, int lower, int upper, int W, int pitch, int ndp,
                            const long long *__restrict__ off, int *oi, int *oj, double *op, int *oslot)
{
    int d = lower + blockIdx.x;
    const int dlo = max(lower, W + 2);
    __shared__ int wsum[32];
    __shared__ long long s_base;
    if (threadIdx.x == 0) s_base = off[blockIdx.x];
    __syncthreads();
    const int nw = blockDim.x >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int x0 = 0; x0 + d < J.n; x0 += blockDim.x) {
        int x = x0 + threadIdx.x;
        bool f = false;
        double p = 0;
        if (x + d < J.n) {
            if (mode == 0) f = is_candidate(J, x, d, pitch);
            else if (d >= dlo) { p = J.prob[(size_t)x * ndp + (d - dlo)]; f = (p == p); }
        }
        unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int before = 0, tot = 0;
        for (int k = 0; k < nw; k++) { if (k < warp) before += wsum[k]; tot += wsum[k]; }
        if (f) {
            long long pos = s_base + before + __popc(b & ((1u << lane) - 1u));
            oi[pos] = x;
            oj[pos] = x + d;
            if (op) op[pos] = p;
            if (oslot) oslot[pos] = J.slot ? J.slot[(size_t)x * ndp + (d - dlo)] : -1;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
}

__global__ void k_gather_features(const float *__restrict__ featT, const int *__restrict__ slots, long long N, int F,
                                  float *__restrict__ out)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N * F) return;
    long long row = k / F;
    int f = (int)(k % F);
    int s = slots[row];
    out[k] = featT[((size_t)(s >> 5) * F + f) * 32 + (s & 31)];
}

// =============================================================================================
// Peakachu.getwindow for arbitrary coordinates, float64 output (API / test accessor; the hot
// path is k_features).  One warp per coordinate, same arithmetic, order-preserving compaction is
// done by the host wrapper from the per-coordinate pass flags.
// =============================================================================================
template <int W>
__global__ void k_getwindow(JobDev J, int upper, int pitch, const int *__restrict__ xi, const int *__restrict__ yi,
                            long long nc, const double *__restrict__ expv, int nexp, unsigned char *__restrict__ pass,
                            double *__restrict__ fea)
{
    constexpr int S = 2 * W + 1, F = S * S;
    __shared__ double win[4][F], tmp[4][F], red[4][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long k = (long long)blockIdx.x * 4 + warp;
    if (k >= nc) return;
    const int x = xi[k], y = yi[k], n = J.n, bw = upper + 14;
    bool ok = (x - W >= 0) && (y + W + 1 <= n) && (y - x - 2 >= W);
    if (!ok) { if (lane == 0) pass[k] = 0; return; }
    int nz = 0;
    for (int c = lane; c < F; c += 32) {
        int a = c / S, b = c % S;
        int r = x - W + a, cc = y - W + b, d = cc - r;
        double v = 0.0;
        if (d >= 0 && d < bw) v = J.band[(size_t)r * pitch + d];
        win[warp][c] = v;
        nz += (v != 0.0);
    }
    for (int o = 16; o; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
    __syncwarp();
    ok = !((double)nz < (double)F * 0.1);
    if (ok) {
        auto ld = [&](int q) { return win[warp][(q / W) * S + (q % W)]; };
        double ll = __ddiv_rn(np_leaf_sum(ld, 0, W * W), (double)(W * W));
        ok = (ll > 0) && (__ddiv_rn(win[warp][W * S + W], ll) > 0.1);
    }
    if (!ok) { if (lane == 0) pass[k] = 0; return; }
    const bool norm = (y - x + 2 * W) < nexp;
    if (norm)
        for (int c = lane; c < F; c += 32) {
            int a = c / S, b = c % S;
            int d = (y - W + b) - (x - W + a);
            win[warp][c] = __ddiv_rn(win[warp][c], expv[d < 0 ? -d : d]);
        }
    __syncwarp();
    for (int c = lane; c < F; c += 32) {
        int a = c / S, b = c % S;
        double v = __dmul_rn(win[warp][a * S + b], c_gw[4]);
        for (int j = 0; j < 4; j++) {
            double l = win[warp][reflect_idx(a - 4 + j, S) * S + b], r = win[warp][reflect_idx(a + 4 - j, S) * S + b];
            v = __dadd_rn(v, __dmul_rn(__dadd_rn(l, r), c_gw[j]));
        }
        tmp[warp][c] = v;
    }
    __syncwarp();
    double mn = 0, mx = 0;
    bool first = true, nn = false;
    for (int c = lane; c < F; c += 32) {
        int a = c / S, b = c % S;
        double v = __dmul_rn(tmp[warp][a * S + b], c_gw[4]);
        for (int j = 0; j < 4; j++) {
            double l = tmp[warp][a * S + reflect_idx(b - 4 + j, S)], r = tmp[warp][a * S + reflect_idx(b + 4 - j, S)];
            v = __dadd_rn(v, __dmul_rn(__dadd_rn(l, r), c_gw[j]));
        }
        win[warp][c] = v;       // safe: each lane only rewrites cells it alone reads from now on
        nn |= (v != v);
        if (first) { mn = mx = v; first = false; }
        else { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
    }
    red[warp][lane] = mn; red[warp][32 + lane] = mx;
    unsigned anynan = __ballot_sync(0xffffffffu, nn);
    __syncwarp();
    for (int l = 0; l < 32; l++) {
        double u = red[warp][l], v = red[warp][32 + l];
        mn = u < mn ? u : mn; mx = v > mx ? v : mx;
    }
    if (anynan) { mn = mx = __longlong_as_double(0x7ff8000000000000LL); }
    double den = __dsub_rn(mx, mn);
    for (int c = lane; c < F; c += 32) fea[(size_t)k * F + c] = __ddiv_rn(__dsub_rn(win[warp][c], mn), den);
    if (lane == 0) pass[k] = 1;
}

// =============================================================================================
// host side
// =============================================================================================
struct nlf_forest {
    nlf_ctx *ctx;
    ForestDev dev;
    uint2 *d_nodes = nullptr;
    uint2 *d_nodes2 = nullptr;       // shared-memory format (leaf values in place), NULL if unusable
    double *d_leaf = nullptr;
    int *d_off = nullptr;
    long long n_nodes = 0;
    std::vector<int> off;            // host copy of tree offsets
    // compact single-pass format (k_forest_one), NULL when the forest cannot be expressed in it
    uint2 *d_nodes3 = nullptr;
    double *d_leafv3 = nullptr;
    unsigned *d_root3 = nullptr;
    int n_int3 = 0, n_leaf3 = 0;
    double visits_per_tree = 0;      // mean root-to-leaf comparisons (unweighted over leaves; informative)
};

constexpr int NLF_MAX_JOBS = 16384;     // jobs per batch (status / counter pool size)

struct HostJob {
    int n = 0;
    double *d_band = nullptr;
    bool band_big = true;       // from the context's big-buffer cache (false: cudaMallocAsync, assembly export)
    size_t prob_off = 0;         // offset (cells) of this job inside the batch's prob / slot pools
    unsigned long long counts[4] = {0, 0, 0, 0};
    int status = 0;
};

struct nlf_batch {
    nlf_ctx *ctx = nullptr;
    int lower = 4, upper = 400, pitch = 416, bw = 414;
    int W = 0, ndp = 0, ndt = 0, TX = 16;
    bool keep_slots = false;
    std::vector<HostJob> jobs;
    cudaEvent_t ev_inputs = nullptr;         // recorded on the copy stream after the last uploaded input
    bool inputs_pending = false;
    // pools: one allocation per batch instead of a dozen per job
    int *d_status = nullptr;                 // [NLF_MAX_JOBS]
    unsigned long long *d_counts = nullptr;  // [NLF_MAX_JOBS*4 + 1]; the last word counts the Donuts
    double *d_scratch = nullptr;             // per job: means and e (2 * up1 doubles)
    int scratch_jobs = 0;
    double *d_prob = nullptr;
    int *d_slot = nullptr;
    int *d_don_job = nullptr, *d_don_i = nullptr, *d_don_j = nullptr;
    double *d_don_p = nullptr, *d_don_v = nullptr;
    long long don_cap = 0;
    JobDev *d_jobs = nullptr;
    int *d_tile_base = nullptr;
    int jobs_cap = 0;
    PixRef *d_pix = nullptr;
    float *d_featT = nullptr;
    unsigned *d_group_counter = nullptr;
    unsigned *d_nanmask = nullptr;
    int *d_overflow = nullptr;
    double *d_exp = nullptr;
    unsigned max_groups = 0;
    int total_tiles = 0;
    long long total_groups = 0;
    bool prepared = false, scored = false, counted = false;
    // host cache of the Donut list (filled by the first fetch after scoring)
    bool donuts_cached = false;
    unsigned long long n_donuts_total = 0;
    std::vector<int> h_don_job, h_don_i, h_don_j;
    std::vector<double> h_don_p, h_don_v;
    std::vector<std::vector<int>> don_by_job;
};

extern int nlf_assembly_export_band(nlf_assembly *a, int bw, int pitch, double **d_band, int *n_out);

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

NLF_EXPORT int nlf_version(void) { return 100; }
NLF_EXPORT const char *nlf_last_error(void) { return g_err; }

NLF_EXPORT int nlf_device_count(int *count)
{
    NLF_CUDA(cudaGetDeviceCount(count));
    return NLF_OK;
}

NLF_EXPORT int nlf_ctx_create(int device, nlf_ctx **out)
{
    if (!out) return fail(NLF_ERR_ARG, "nlf_ctx_create: out is NULL");
    int ndev = 0;
    NLF_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(NLF_ERR_ARG, "device %d not in [0,%d)", device, ndev);
    NLF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NLF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(NLF_ERR_CUDA, "device %d is sm_%d%d; libnlf_b200 is built for sm_100a only", device, prop.major,
                    prop.minor);
    nlf_ctx *c = new nlf_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    NLF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    NLF_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    NLF_CUDA(cudaEventCreate(&c->t0));
    NLF_CUDA(cudaEventCreate(&c->t1));
    cudaMemPool_t pool;
    NLF_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    NLF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = c;
    return NLF_OK;
}

NLF_EXPORT int nlf_ctx_destroy(nlf_ctx *c)
{
    if (!c) return NLF_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamSynchronize(c->stream);
    c->aux_release();
    c->prof_resolve();
    c->big_release_all();
    c->stage_release_all();
    for (auto e : c->free_events) cudaEventDestroy(e);
    cudaEventDestroy(c->t0);
    cudaEventDestroy(c->t1);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    delete c;
    return NLF_OK;
}

NLF_EXPORT int nlf_ctx_sync(nlf_ctx *c)
{
    NLF_CUDA(cudaSetDevice(c->device));
    NLF_CUDA(cudaStreamSynchronize(c->copy_stream));
    if (c->aux_stream) NLF_CUDA(cudaStreamSynchronize(c->aux_stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    return NLF_OK;
}

NLF_EXPORT int nlf_timer_start(nlf_ctx *c)
{
    NLF_CUDA(cudaSetDevice(c->device));
    NLF_CUDA(cudaEventRecord(c->t0, c->stream));
    NLF_CUDA(cudaStreamWaitEvent(c->copy_stream, c->t0, 0));     // uploads of the timed region start after t0
    return NLF_OK;
}

NLF_EXPORT int nlf_timer_stop(nlf_ctx *c, float *ms)
{
    NLF_CUDA(cudaSetDevice(c->device));
    NLF_CUDA(cudaEventRecord(c->t1, c->stream));
    NLF_CUDA(cudaEventSynchronize(c->t1));
    NLF_CUDA(cudaEventElapsedTime(ms, c->t0, c->t1));
    return NLF_OK;
}

NLF_EXPORT int nlf_launch_count(nlf_ctx *c, long long *count) { *count = c->launches; return NLF_OK; }
NLF_EXPORT int nlf_profile_enable(nlf_ctx *c, int on) { c->profile = on != 0; return NLF_OK; }
NLF_EXPORT int nlf_profile_reset(nlf_ctx *c)
{
    c->prof_resolve();
    for (int i = 0; i < P_NCLASS; i++) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    return NLF_OK;
}
NLF_EXPORT int nlf_profile_read(nlf_ctx *c, int which, float *total_ms, long long *launches)
{
    if (which < 0 || which >= P_NCLASS) return fail(NLF_ERR_ARG, "profile class %d", which);
    NLF_CUDA(cudaSetDevice(c->device));
    c->prof_resolve();
    *total_ms = c->prof_ms[which];
    *launches = c->prof_n[which];
    return NLF_OK;
}

// ---- forest -----------------------------------------------------------------------------------
// Compact forest (k_forest_one): internal nodes only, depth-first with the IMPLICIT child right behind its parent.
// Usable when thresholds and leaf values are non-negative and below 2 (the staged words keep their two top bits for
// flags), there are at most 2048 distinct leaf values among the leaves of terminal nodes (11-bit references) and
// every index fits its field.  Host only (no CUDA call): also exported for the CPU tests (nlf_forest_compact_host).
static bool build_compact_forest(int n_trees, const int64_t *tree_off, const int32_t *left, const int32_t *right,
                                 const int32_t *feature, const double *threshold, const double *leaf_p1,
                                 std::vector<uint2> &nodes3, std::vector<double> &leafv3, std::vector<unsigned> &root3)
{
    const long long total = tree_off[n_trees];
    nodes3.clear();
    leafv3.clear();
    root3.assign((size_t)n_trees, 0u);
    bool v3_ok = true;
    std::map<uint64_t, unsigned> leaf_ref;
    auto ref_of = [&](long long node) -> unsigned {
        const double v = leaf_p1[node];
        uint64_t bits;
        memcpy(&bits, &v, 8);
        auto it = leaf_ref.find(bits);
        if (it != leaf_ref.end()) return it->second;
        const unsigned r = (unsigned)leafv3.size();
        leaf_ref[bits] = r;
        leafv3.push_back(v);
        return r;
    };
    long long n_internal = 0;
    for (long long k = 0; k < total; k++) {
        if (left[k] >= 0) {
            n_internal++;
            if (!(threshold[k] >= 0.0 && threshold[k] < 2.0)) v3_ok = false;
            if (feature[k] < 0 || feature[k] > 255) v3_ok = false;
        } else if (!(leaf_p1[k] >= 0.0 && leaf_p1[k] < 2.0) || std::signbit(leaf_p1[k])) {
            v3_ok = false;
        }
    }
    if (n_internal >= (1 << 21)) v3_ok = false;
    const unsigned LEAF = 0x80000000u;                 // marks "slot = n_internal + reference" until n_internal is final
    struct Item { long long node; long long patch; };  // patch: compact index of the parent whose explicit field waits
    std::vector<std::pair<unsigned, unsigned>> leaf_fix;   // (compact node, leaf reference) of explicit leaf children
    for (int t = 0; t < n_trees && v3_ok; t++) {
        const long long b = tree_off[t];
        if (left[b] < 0) { root3[(size_t)t] = LEAF | ref_of(b); continue; }
        root3[(size_t)t] = (unsigned)nodes3.size();
        std::vector<Item> stack;
        stack.push_back({b, -1});
        while (!stack.empty()) {
            const Item itx = stack.back();
            stack.pop_back();
            const long long k = itx.node;
            const unsigned idx = (unsigned)nodes3.size();
            if (itx.patch >= 0) nodes3[(size_t)itx.patch].y |= idx;
            const long long L = b + left[k], R = b + right[k];
            const bool lleaf = left[L] < 0, rleaf = left[R] < 0;
            uint2 nd;
            float tf = (float)threshold[k];
            if ((double)tf > threshold[k]) tf = nextafterf(tf, -INFINITY);   // largest float <= threshold
            memcpy(&nd.x, &tf, 4);
            nd.y = ((unsigned)feature[k] & 0xffu) << 22;
            if (lleaf && rleaf) {
                const unsigned rl = ref_of(L), rr = ref_of(R);
                if (rl >= 2048u || rr >= 2048u) { v3_ok = false; break; }
                nd.y |= 0x40000000u | rl | (rr << 11);                 // terminal
                nodes3.push_back(nd);
            } else if (!lleaf && !rleaf) {
                nd.y |= 0x80000000u;                                     // implicit left, explicit right (patched later)
                nodes3.push_back(nd);
                stack.push_back({R, (long long)idx});
                stack.push_back({L, -1});                                // popped next: lands at idx + 1
            } else if (rleaf) {
                nd.y |= 0x80000000u;                                     // implicit left, explicit right leaf
                nodes3.push_back(nd);
                leaf_fix.push_back({idx, ref_of(R)});
                stack.push_back({L, -1});
            } else {
                nd.y |= 0x80000000u | 0x40000000u;                       // S: the implicit child is the right one
                nodes3.push_back(nd);
                leaf_fix.push_back({idx, ref_of(L)});
                stack.push_back({R, -1});
            }
        }
    }
    // explicit leaf children and one-leaf trees: slot = number of internal nodes + reference
    const unsigned n_int = (unsigned)nodes3.size();
    if (n_int + leafv3.size() >= (1u << 22) || leafv3.empty()) v3_ok = false;
    if (v3_ok) {
        for (auto &fx : leaf_fix) nodes3[fx.first].y |= n_int + fx.second;
        for (auto &r : root3) if (r & LEAF) r = n_int + (r & ~LEAF);
    }
    return v3_ok;
}

// The compact forest on the host, for checking the encoding without a GPU: returns 1 and fills the arrays when the
// forest is representable and the capacities suffice, 0 when it is not representable, < 0 on bad arguments.
// node_thr[k] = float bits of the threshold, node_word[k] = packed word (see k_forest_one), root[t] = slot of the root.
NLF_EXPORT int nlf_forest_compact_host(int n_trees, const int64_t *tree_off, const int32_t *left, const int32_t *right,
                                       const int32_t *feature, const double *threshold, const double *leaf_p1,
                                       long long cap_nodes, uint32_t *node_thr, uint32_t *node_word, long long cap_leaf,
                                       double *leafv, uint32_t *root, long long *n_int, long long *n_leaf)
{
    if (n_trees <= 0 || !tree_off || !left || !right || !feature || !threshold || !leaf_p1 || !n_int || !n_leaf)
        return fail(NLF_ERR_ARG, "nlf_forest_compact_host: bad arguments");
    std::vector<uint2> nodes3;
    std::vector<double> leafv3;
    std::vector<unsigned> root3;
    if (!build_compact_forest(n_trees, tree_off, left, right, feature, threshold, leaf_p1, nodes3, leafv3, root3)) return 0;
    *n_int = (long long)nodes3.size();
    *n_leaf = (long long)leafv3.size();
    if (!node_thr || !node_word || !leafv || !root) return 1;                 // size query
    if (cap_nodes < *n_int || cap_leaf < *n_leaf) return fail(NLF_ERR_ARG, "nlf_forest_compact_host: capacity too small");
    for (size_t k = 0; k < nodes3.size(); k++) { node_thr[k] = nodes3[k].x; node_word[k] = nodes3[k].y; }
    for (size_t k = 0; k < leafv3.size(); k++) leafv[k] = leafv3[k];
    for (int t = 0; t < n_trees; t++) root[t] = root3[(size_t)t];
    return 1;
}

NLF_EXPORT int nlf_forest_create(nlf_ctx *ctx, int n_trees, const int64_t *tree_off, const int32_t *left,
                                 const int32_t *right, const int32_t *feature, const double *threshold,
                                 const uint8_t *missing_left, const double *leaf_p1, int n_features,
                                 nlf_forest **out)
{
    if (!ctx || !out || n_trees <= 0) return fail(NLF_ERR_ARG, "nlf_forest_create: bad arguments");
    if (n_features <= 0 || n_features >= 255)
        return fail(NLF_ERR_ARG, "n_features=%d not supported (node format holds 1..254)", n_features);
    long long total = tree_off[n_trees];
    std::vector<uint2> nodes((size_t)total);
    std::vector<int> off(n_trees + 1);
    for (int t = 0; t <= n_trees; t++) off[t] = (int)tree_off[t];
    for (int t = 0; t < n_trees; t++) {
        long long b = tree_off[t], cnt = tree_off[t + 1] - b;
        if (cnt >= (1 << 23)) return fail(NLF_ERR_ARG, "tree %d has %lld nodes (limit 2^23)", t, cnt);
        for (long long k = 0; k < cnt; k++) {
            uint2 nd;
            if (left[b + k] < 0) {
                nd.x = 0; nd.y = 0xffu;
            } else {
                if (left[b + k] != k + 1)
                    return fail(NLF_ERR_ARG, "tree %d node %lld: left child %d is not node+1 (depth-first layout expected)",
                                t, k, left[b + k]);
                if (feature[b + k] < 0 || feature[b + k] >= n_features)
                    return fail(NLF_ERR_ARG, "tree %d node %lld: feature %d out of range", t, k, feature[b + k]);
                double th = threshold[b + k];
                float tf = (float)th;
                if ((double)tf > th) tf = nextafterf(tf, -INFINITY);   // largest float <= th
                memcpy(&nd.x, &tf, 4);
                nd.y = (unsigned)feature[b + k] | ((missing_left && missing_left[b + k]) ? 0x100u : 0u) |
                       ((unsigned)right[b + k] << 9);
            }
            nodes[(size_t)(b + k)] = nd;
        }
    }
    // shared-memory format: leaves carry their value in place (needs a clear sign bit); right
    // children stay tree-local here and are rebased when a kernel stages its part.
    std::vector<uint2> nodes2((size_t)total);
    bool v2_ok = true;
    for (int t = 0; t < n_trees && v2_ok; t++) {
        long long b = tree_off[t], cnt = tree_off[t + 1] - b;
        if (cnt >= (1 << 19)) v2_ok = false;
        for (long long k = 0; k < cnt && v2_ok; k++) {
            uint2 nd;
            if (left[b + k] < 0) {
                double v = leaf_p1[b + k];
                if (std::signbit(v)) { v2_ok = false; break; }
                uint64_t bits;
                memcpy(&bits, &v, 8);
                nd.x = (unsigned)(bits & 0xffffffffu);
                nd.y = (unsigned)(bits >> 32);
            } else {
                nd.x = nodes[(size_t)(b + k)].x;
                nd.y = 0x80000000u | ((unsigned)feature[b + k] << 23) |
                       ((missing_left && missing_left[b + k]) ? 0x400000u : 0u) | (unsigned)right[b + k];
            }
            nodes2[(size_t)(b + k)] = nd;
        }
    }
    // compact format (k_forest_one): internal nodes only, see build_compact_forest
    std::vector<uint2> nodes3;
    std::vector<double> leafv3;
    std::vector<unsigned> root3;
    const bool v3_ok = build_compact_forest(n_trees, tree_off, left, right, feature, threshold, leaf_p1, nodes3, leafv3, root3);
    NLF_CUDA(cudaSetDevice(ctx->device));
    nlf_forest *f = new nlf_forest();
    f->ctx = ctx;
    f->n_nodes = total;
    f->off = off;
    if (v2_ok) {
        NLF_CUDA(cudaMalloc(&f->d_nodes2, sizeof(uint2) * (size_t)total));
        NLF_CUDA(cudaMemcpy(f->d_nodes2, nodes2.data(), sizeof(uint2) * (size_t)total, cudaMemcpyHostToDevice));
    }
    NLF_CUDA(cudaMalloc(&f->d_nodes, sizeof(uint2) * (size_t)total));
    NLF_CUDA(cudaMalloc(&f->d_leaf, sizeof(double) * (size_t)total));
    NLF_CUDA(cudaMalloc(&f->d_off, sizeof(int) * (size_t)(n_trees + 1)));
    NLF_CUDA(cudaMemcpy(f->d_nodes, nodes.data(), sizeof(uint2) * (size_t)total, cudaMemcpyHostToDevice));
    NLF_CUDA(cudaMemcpy(f->d_leaf, leaf_p1, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice));
    NLF_CUDA(cudaMemcpy(f->d_off, off.data(), sizeof(int) * (size_t)(n_trees + 1), cudaMemcpyHostToDevice));
    f->dev.nodes = f->d_nodes;
    f->dev.leaf = f->d_leaf;
    f->dev.tree_off = f->d_off;
    f->dev.n_trees = n_trees;
    f->dev.n_features = n_features;
    if (v3_ok) {
        f->n_int3 = (int)nodes3.size();
        f->n_leaf3 = (int)leafv3.size();
        NLF_CUDA(cudaMalloc(&f->d_nodes3, sizeof(uint2) * std::max<size_t>(1, nodes3.size())));
        NLF_CUDA(cudaMalloc(&f->d_leafv3, sizeof(double) * leafv3.size()));
        NLF_CUDA(cudaMalloc(&f->d_root3, sizeof(unsigned) * (size_t)n_trees));
        if (!nodes3.empty())
            NLF_CUDA(cudaMemcpy(f->d_nodes3, nodes3.data(), sizeof(uint2) * nodes3.size(), cudaMemcpyHostToDevice));
        NLF_CUDA(cudaMemcpy(f->d_leafv3, leafv3.data(), sizeof(double) * leafv3.size(), cudaMemcpyHostToDevice));
        NLF_CUDA(cudaMemcpy(f->d_root3, root3.data(), sizeof(unsigned) * (size_t)n_trees, cudaMemcpyHostToDevice));
    }
    *out = f;
    return NLF_OK;
}

NLF_EXPORT int nlf_forest_destroy(nlf_forest *f)
{
    if (!f) return NLF_OK;
    cudaSetDevice(f->ctx->device);
    cudaFree(f->d_nodes);
    if (f->d_nodes2) cudaFree(f->d_nodes2);
    if (f->d_nodes3) cudaFree(f->d_nodes3);
    if (f->d_leafv3) cudaFree(f->d_leafv3);
    if (f->d_root3) cudaFree(f->d_root3);
    cudaFree(f->d_leaf);
    cudaFree(f->d_off);
    delete f;
    return NLF_OK;
}

NLF_EXPORT int nlf_forest_predict(nlf_forest *f, const float *X, long long N, double *proba)
{
    nlf_ctx *c = f->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    if (N <= 0) return NLF_OK;
    float *dX;
    double *dP;
    size_t bytes = sizeof(float) * (size_t)N * f->dev.n_features;
    NLF_CUDA(cudaMallocAsync(&dX, bytes, c->stream));
    NLF_CUDA(cudaMallocAsync(&dP, sizeof(double) * (size_t)N, c->stream));
    NLF_CUDA(cudaMemcpyAsync(dX, X, bytes, cudaMemcpyHostToDevice, c->stream));
    {
        KLaunch k(c, P_FOREST);
        k_forest_rows<<<(unsigned)((N + 127) / 128), 128, 0, c->stream>>>(f->dev, dX, N, dP);
    }
    NLF_CUDA(cudaGetLastError());
    NLF_CUDA(cudaMemcpyAsync(proba, dP, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    NLF_CUDA(cudaFreeAsync(dX, c->stream));
    NLF_CUDA(cudaFreeAsync(dP, c->stream));
    return NLF_OK;
}

// ---- batch ------------------------------------------------------------------------------------
NLF_EXPORT int nlf_batch_create(nlf_ctx *ctx, int lower, int upper, nlf_batch **out)
{
    if (!ctx || !out) return fail(NLF_ERR_ARG, "nlf_batch_create: NULL argument");
    if (lower < 0 || upper < lower || upper > 4000) return fail(NLF_ERR_ARG, "bad band limits lower=%d upper=%d", lower, upper);
    NLF_CUDA(cudaSetDevice(ctx->device));
    nlf_batch *b = new nlf_batch();
    b->ctx = ctx;
    b->lower = lower;
    b->upper = upper;
    b->bw = upper + 14;
    b->pitch = round_up(b->bw, 4);
    cudaStream_t s = ctx->stream;
    NLF_CUDA(cudaMallocAsync(&b->d_status, sizeof(int) * NLF_MAX_JOBS, s));
    NLF_CUDA(cudaMallocAsync(&b->d_counts, sizeof(unsigned long long) * (NLF_MAX_JOBS * 4 + 1), s));
    NLF_CUDA(cudaMemsetAsync(b->d_status, 0, sizeof(int) * NLF_MAX_JOBS, s));
    NLF_CUDA(cudaMemsetAsync(b->d_counts, 0, sizeof(unsigned long long) * (NLF_MAX_JOBS * 4 + 1), s));
    NLF_CUDA(ctx->copy_fence());                    // inputs of this batch come after everything enqueued so far
    *out = b;
    return NLF_OK;
}

// inputs arrive on the copy stream: remember the point the main stream has to wait for
static cudaError_t inputs_uploaded(nlf_batch *b)
{
    cudaError_t e;
    if (!b->ev_inputs && (e = cudaEventCreateWithFlags(&b->ev_inputs, cudaEventDisableTiming)) != cudaSuccess) return e;
    b->inputs_pending = true;
    return cudaEventRecord(b->ev_inputs, b->ctx->copy_stream);
}

static cudaError_t wait_inputs(nlf_batch *b)
{
    if (!b->inputs_pending) return cudaSuccess;
    b->inputs_pending = false;
    return cudaStreamWaitEvent(b->ctx->stream, b->ev_inputs, 0);
}

// PixRef packs (row << dshift) | diagonal in 32 bits
static int dshift_of(const nlf_batch *b)
{
    int s = 1;
    while ((1 << s) <= b->upper) s++;
    return s;
}
static int max_rows(const nlf_batch *b) { return (int)std::min<long long>((1LL << (32 - dshift_of(b))) - 1, 1LL << 24); }

static int check_can_add(nlf_batch *b)
{
    if (b->scored || b->prepared) return fail(NLF_ERR_STATE, "batch already prepared/scored: add matrices first");
    if ((int)b->jobs.size() >= NLF_MAX_JOBS) return fail(NLF_ERR_ARG, "more than %d matrices in one batch", NLF_MAX_JOBS);
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_add_dense(nlf_batch *b, const double *G, int n)
{
    if (!b || !G || n <= 0) return fail(NLF_ERR_ARG, "nlf_batch_add_dense: bad arguments");
    if (n > 65535) return fail(NLF_ERR_ARG, "dense input limited to n <= 65535 (use the band or the pixel-table input), got %d", n);
    int rc = check_can_add(b);
    if (rc) return rc;
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    HostJob j;
    j.n = n;
    double *d_dense;
    cudaStream_t cs = c->copy_stream;
    NLF_CUDA(c->big_alloc((void **)&j.d_band, sizeof(double) * (size_t)n * b->pitch, cs));
    NLF_CUDA(c->stage_acquire(sizeof(double) * (size_t)n * n, (void **)&d_dense));
    if (n < 1024) {
        NLF_CUDA(cudaMemcpyAsync(d_dense, G, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, cs));
    } else {
        // only the cells k_bandify reads (13 sub-diagonals and bw super-diagonals) cross the bus: row
        // blocks of 128 rows, each a 2-D copy of the columns its rows can touch
        const int RB = 128;
        for (int r0 = 0; r0 < n; r0 += RB) {
            const int r1 = std::min(n, r0 + RB);
            const int c0 = std::max(0, r0 - 13), c1 = std::min(n, r1 - 1 + b->bw);
            NLF_CUDA(cudaMemcpy2DAsync(d_dense + (size_t)r0 * n + c0, sizeof(double) * (size_t)n, G + (size_t)r0 * n + c0,
                                       sizeof(double) * (size_t)n, sizeof(double) * (size_t)(c1 - c0), (size_t)(r1 - r0),
                                       cudaMemcpyHostToDevice, cs));
        }
    }
    {
        KLaunch k(c, P_BANDIFY, cs);
        dim3 grid(n, (b->pitch + 127) / 128);
        k_bandify<<<grid, 128, 0, cs>>>(d_dense, n, b->bw, b->pitch, j.d_band, b->d_status + b->jobs.size());
    }
    NLF_CUDA(cudaGetLastError());
    c->stage_release(d_dense);
    NLF_CUDA(inputs_uploaded(b));
    b->jobs.push_back(j);
    return (int)b->jobs.size() - 1;
}

NLF_EXPORT int nlf_batch_add_band(nlf_batch *b, const double *band, int n, int pitch)
{
    if (!b || !band || n <= 0 || pitch <= 0) return fail(NLF_ERR_ARG, "nlf_batch_add_band: bad arguments");
    if (n > max_rows(b)) return fail(NLF_ERR_ARG, "band input limited to %d rows per job at upper=%d, got %d", max_rows(b), b->upper, n);
    int rc = check_can_add(b);
    if (rc) return rc;
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    HostJob j;
    j.n = n;
    double *d_src;
    cudaStream_t cs = c->copy_stream;
    NLF_CUDA(c->big_alloc((void **)&j.d_band, sizeof(double) * (size_t)n * b->pitch, cs));
    NLF_CUDA(c->stage_acquire(sizeof(double) * (size_t)n * pitch, (void **)&d_src));
    NLF_CUDA(cudaMemcpyAsync(d_src, band, sizeof(double) * (size_t)n * pitch, cudaMemcpyHostToDevice, cs));
    {
        KLaunch k(c, P_BANDIFY, cs);
        dim3 grid(n, (b->pitch + 127) / 128);
        k_band_copy<<<grid, 128, 0, cs>>>(d_src, n, pitch, b->bw, b->pitch, j.d_band);
    }
    NLF_CUDA(cudaGetLastError());
    c->stage_release(d_src);
    NLF_CUDA(inputs_uploaded(b));
    b->jobs.push_back(j);
    return (int)b->jobs.size() - 1;
}

NLF_EXPORT int nlf_batch_add_assembly(nlf_batch *b, nlf_assembly *a)
{
    if (!b || !a) return fail(NLF_ERR_ARG, "nlf_batch_add_assembly: NULL argument");
    int rc = check_can_add(b);
    if (rc) return rc;
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    HostJob j;
    rc = nlf_assembly_export_band(a, b->bw, b->pitch, &j.d_band, &j.n);
    if (rc) return rc;
    j.band_big = false;
    b->jobs.push_back(j);
    return (int)b->jobs.size() - 1;
}

// A whole chromosome (global bins [lo, hi) of the resident pixel table) as a trivial assembly:
// gap columns removed (callers.py:504-522), the finite band of the gap-free matrix built on the device
// (callers.py:527-535).  kept_out[k] = bin (relative to lo) behind gap-free index k; capacity hi - lo.
NLF_EXPORT int nlf_batch_add_pixels(nlf_batch *b, nlf_pixels *px, int balance, long long lo, long long hi,
                                    int *kept_out, int *n_kept)
{
    if (!b || !px || !n_kept || lo < 0 || hi <= lo) return fail(NLF_ERR_ARG, "nlf_batch_add_pixels: bad arguments");
    if (px->ctx != b->ctx) return fail(NLF_ERR_ARG, "nlf_batch_add_pixels: pixel table lives on another context");
    if (hi > px->nbins) return fail(NLF_ERR_ARG, "nlf_batch_add_pixels: bins [%lld,%lld) outside [0,%lld)", lo, hi, px->nbins);
    if (hi - lo > max_rows(b)) return fail(NLF_ERR_ARG, "band input limited to %d rows per job at upper=%d, got %lld", max_rows(b), b->upper, hi - lo);
    int rc = check_can_add(b);
    if (rc) return rc;
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    HostJob j;
    std::vector<int> kept;
    rc = nlf_pixels_export_band(px, balance, lo, hi, b->bw, b->pitch, &j.d_band, kept);
    if (rc) return rc;
    j.n = (int)kept.size();
    j.band_big = true;
    *n_kept = j.n;
    if (kept_out) memcpy(kept_out, kept.data(), sizeof(int) * kept.size());
    b->jobs.push_back(j);
    return (int)b->jobs.size() - 1;
}

NLF_EXPORT int nlf_batch_njobs(nlf_batch *b) { return b ? (int)b->jobs.size() : 0; }

// Matrix rows per feature tile.  Taller tiles amortise the per-tile prologue (band tile + halo load, O/E
// division, gates, slot claim: ~29 % of the warp time at 16 rows in the ncu source view) over more rows.
static int feat_tx(int w)
{
    int tx = (w >= 7) ? 16 : 32;
    if (const char *e = getenv("NLF_FEAT_TX")) {
        const int v = atoi(e);
        if ((w >= 7 && (v == 8 || v == 16)) || (w < 7 && (v == 12 || v == 14 || v == 16 || v == 32))) tx = v;
    }
    return tx;
}

// where one chunk of tiles puts its scored pixels: a region of the batch's feature buffers
struct ChunkBuf {
    PixRef *pix;
    float *featT;
    unsigned *nanmask;
    unsigned *counter;
    unsigned max_slots;
    unsigned slot_base;      // first slot of the region inside the batch's buffers (for the kept-slot accessors)
};

template <int W, int TXP>
static const void *feature_kernel() { return reinterpret_cast<const void *>(&k_features<W, TXP>); }

static const void *feature_kernel_ptr(int w, int tx)
{
    if (w == 5)
        return tx == 32 ? feature_kernel<5, 32>() : tx == 16 ? feature_kernel<5, 16>() : tx == 14 ? feature_kernel<5, 14>()
                                                                                                : feature_kernel<5, 12>();
    if (w == 6)
        return tx == 32 ? feature_kernel<6, 32>() : tx == 16 ? feature_kernel<6, 16>() : tx == 14 ? feature_kernel<6, 14>()
                                                                                                : feature_kernel<6, 12>();
    return tx == 16 ? feature_kernel<7, 16>() : feature_kernel<7, 8>();
}

static size_t feature_dyn_smem(int w, int tx) { return sizeof(double) * 2 * (size_t)(tx + 2 * w) * (32 + 4 * w); }

static int launch_features(nlf_batch *b, int w, int tx, int tile_offset, int n_tiles, int nexp, const ChunkBuf &cb)
{
    nlf_ctx *c = b->ctx;
    const void *kern = feature_kernel_ptr(w, tx);
    const size_t dyn = feature_dyn_smem(w, tx);
    static bool configured[64][3][5] = {{{false}}};
    const int ti = tx == 32 ? 0 : tx == 16 ? 1 : tx == 14 ? 2 : tx == 12 ? 3 : 4;
    if (!configured[c->device & 63][w - 5][ti]) {
        NLF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        // same shared-memory carve-out as the forest kernel, so that CTAs of both can share an SM
        NLF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[c->device & 63][w - 5][ti] = true;
    }
    const JobDev *jobs = b->d_jobs;
    const int *tile_base = b->d_tile_base;
    int njobs = (int)b->jobs.size();
    const double *expv = b->d_exp;
    PixRef *pix = cb.pix;
    float *featT = cb.featT;
    unsigned *counter = cb.counter;
    unsigned max_slots = cb.max_slots;
    int *overflow = b->d_overflow;
    unsigned *nanmask = cb.nanmask;
    unsigned slot_base = cb.slot_base;
    void *args[] = {&jobs, &tile_base, &njobs, &tile_offset, &b->lower, &b->upper, &b->pitch, &b->ndp, &b->ndt, &expv, &nexp,
                    &pix, &featT, &counter, &max_slots, &overflow, &nanmask, &slot_base};
    KLaunch k(c, P_FEATURE);
    NLF_CUDA(cudaLaunchKernel(kern, dim3((unsigned)n_tiles), dim3(32u * (2 * w + 1)), args, dyn, c->stream));
    return NLF_OK;
}

// dynamic shared memory of the cooperative forest kernels for a part of `nodes` nodes in `ntree` trees
static inline size_t forest_dyn_smem(size_t nodes, int ntree, int F)
{
    return 2 * (size_t)F * 128 + 2 * (size_t)ntree * 256 + 256 + ((nodes + 2) & ~(size_t)1) * 8 +
           sizeof(int) * (size_t)((ntree + 8 + 3) & ~3) + 64;
}

// How one nlf_batch_score call runs the feature and the forest kernels.
struct ScorePlan {
    bool one = false;        // single-pass compact forest (k_forest_one)
    bool coop = false;       // ... launched on the aux stream, sharing every SM with the feature kernel of the next chunk
    int tx = 32;             // matrix rows per feature tile
    int nw = 32;             // warps of the forest CTA
    size_t forest_smem = 0;
};

template <int NW>
static const void *forest_one_kernel() { return reinterpret_cast<const void *>(&k_forest_one<NW, 2>); }

static const void *forest_one_ptr(int nw)
{
    return nw == 32 ? forest_one_kernel<32>() : nw == 24 ? forest_one_kernel<24>() : nw == 20 ? forest_one_kernel<20>()
                                                                                            : forest_one_kernel<16>();
}

// Decide the plan from the real resource numbers of the compiled kernels: the forest CTA (whole compact forest
// + two feature tiles + leaf references) and ONE feature CTA must fit an SM together (shared memory incl. the
// 1 KB the system reserves per CTA, registers per warp in units of 512); the tile height shrinks until they do.
static int make_plan(nlf_batch *b, nlf_forest *f, int w, int F, ScorePlan *out)
{
    ScorePlan p;
    p.tx = feat_tx(w);
    const size_t smem_block_max = 232448;                      // 227 KB opt-in limit per block on sm_100
    const size_t smem_sm = 233472;                             // 228 KB per SM
    // NLF_SCORE_MODE = parts | one | coop.  Measured on the bench forest (profiles/r02_score_modes.md): the two-part
    // kernel (k_forest_pipe) is the fastest, the single-pass compact kernel pays ~2 extra instructions per node step
    // for its swap bit (+15 %), and sharing the SMs with the feature kernel is limited by the register file of
    // an SM sub-partition (4 feature warps x 72 registers leave room for 20 forest warps, which traverse 34 %
    // slower than 32).  Default: parts while the forest needs at most two of them, else the single pass.
    const char *mode = getenv("NLF_SCORE_MODE");
    const bool small_forest = f->d_nodes2 && (size_t)f->n_nodes * 8 <= 2 * (size_t)150 * 1024;
    const bool allow_one = mode ? strcmp(mode, "parts") != 0 : !small_forest;
    const bool allow_coop = mode && !strcmp(mode, "coop");
    if (f->d_nodes3 && allow_one) {
        p.forest_smem = forest3_smem(F, f->dev.n_trees, f->n_int3, f->n_leaf3);
        cudaFuncAttributes fa;
        NLF_CUDA(cudaFuncGetAttributes(&fa, forest_one_ptr(32)));
        p.one = p.forest_smem + fa.sharedSizeBytes <= smem_block_max;
        if (p.one && allow_coop && !b->keep_slots) {
            const int cands5[] = {32, 16, 14, 12}, cands7[] = {16, 8};
            const int *cands = w >= 7 ? cands7 : cands5;
            const int ncand = w >= 7 ? 2 : 4;
            const char *etx = getenv("NLF_FEAT_TX");
            for (int k = 0; k < ncand && !p.coop; k++) {
                const int tx = cands[k];
                if (etx && tx != p.tx) continue;               // the knob pins the height
                cudaFuncAttributes ka;
                NLF_CUDA(cudaFuncGetAttributes(&ka, feature_kernel_ptr(w, tx)));
                const size_t feat_smem = ka.sharedSizeBytes + feature_dyn_smem(w, tx) + 1024;
                if (feat_smem + p.forest_smem + fa.sharedSizeBytes + 1024 > smem_sm) continue;
                const int feat_regs = ((ka.numRegs * 32 + 511) / 512) * 512 * (2 * w + 1);
                const int nws[] = {32, 24, 20, 16};
                for (int q = 0; q < 4; q++) {
                    cudaFuncAttributes qa;
                    NLF_CUDA(cudaFuncGetAttributes(&qa, forest_one_ptr(nws[q])));
                    if (feat_regs + ((qa.numRegs * 32 + 511) / 512) * 512 * nws[q] <= 65536) {
                        p.coop = true;
                        p.tx = tx;
                        p.nw = nws[q];
                        break;
                    }
                }
            }
        }
        if (const char *e = getenv("NLF_FOREST_NW")) {
            const int v = atoi(e);
            if (v == 32 || v == 24 || v == 20 || v == 16) p.nw = v;
        }
    }
    *out = p;
    return NLF_OK;
}

static int launch_forest_one(nlf_batch *b, nlf_forest *f, int w, int F, const ScorePlan &p, const ChunkBuf &cb, cudaStream_t st)
{
    nlf_ctx *c = b->ctx;
    const void *kern = forest_one_ptr(p.nw);
    static bool configured[64][4] = {{false}};
    const int qi = p.nw == 32 ? 0 : p.nw == 24 ? 1 : p.nw == 20 ? 2 : 3;
    if (!configured[c->device & 63][qi]) {
        NLF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 64));
        NLF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[c->device & 63][qi] = true;
    }
    Forest3Dev fd{f->d_nodes3, f->d_leafv3, f->d_root3, f->n_int3, f->n_leaf3, f->dev.n_trees};
    const PixRef *pix = cb.pix;
    const unsigned *nanmask = cb.nanmask;
    const float *featT = cb.featT;
    const unsigned *counter = cb.counter;
    const JobDev *jobs = b->d_jobs;
    void *args[] = {&fd, &F, &pix, &nanmask, &featT, &counter, &jobs, &b->lower, &w, &b->ndp};
    {
        KLaunch k(c, P_FOREST, st);
        NLF_CUDA(cudaLaunchKernel(kern, dim3((unsigned)c->sm_count), dim3(32u * p.nw), args, p.forest_smem, st));
    }
    {
        // pixels with NaN features (rare: constant windows) through the general kernel
        KLaunch k(c, P_FOREST, st);
        k_forest_nan<<<c->sm_count * 4, 256, 0, st>>>(f->dev, cb.pix, cb.nanmask, cb.featT, cb.counter, b->d_jobs, b->lower, w, b->ndp);
        NLF_CUDA(cudaGetLastError());
    }
    return NLF_OK;
}

// K8 launch (fallback when the compact forest is unusable): shared-memory forest parts when they fit, else the
// global-memory kernel.
static int launch_forest(nlf_batch *b, nlf_forest *f, int w, int F, const ChunkBuf &cb)
{
    nlf_ctx *c = b->ctx;
    const size_t smem_max = 232448;                            // 227 KB opt-in limit per block on sm_100
    const int T = f->dev.n_trees;
    bool use_smem = f->d_nodes2 != nullptr;
    std::vector<int> part_begin;
    if (use_smem) {
        // parts of consecutive trees that fit in shared memory, then an even split of the tree count
        int nparts_needed = 0, t = 0;
        while (t < T && use_smem) {
            int t1 = t;
            while (t1 < T && forest_dyn_smem((size_t)(f->off[t1 + 1] - f->off[t]), t1 + 1 - t, F) <= smem_max - 64) t1++;
            if (t1 == t) use_smem = false;                     // a single tree does not fit
            nparts_needed++;
            t = t1;
        }
        if (use_smem) {
            bool even_ok = true;
            std::vector<int> even;
            for (int p = 0; p <= nparts_needed; p++) even.push_back((int)((long long)T * p / nparts_needed));
            for (int p = 0; p < nparts_needed && even_ok; p++)
                if (forest_dyn_smem((size_t)(f->off[even[p + 1]] - f->off[even[p]]), even[p + 1] - even[p], F) > smem_max - 64)
                    even_ok = false;
            if (even_ok) part_begin = even;
            else {
                t = 0;
                while (t < T) {
                    part_begin.push_back(t);
                    int t1 = t;
                    while (t1 < T && forest_dyn_smem((size_t)(f->off[t1 + 1] - f->off[t]), t1 + 1 - t, F) <= smem_max - 64) t1++;
                    t = t1;
                }
                part_begin.push_back(T);
            }
        }
    }
    if (!use_smem) {
        constexpr int GW = 4;
        size_t smem = sizeof(float) * (size_t)GW * F * 32;
        NLF_CUDA(cudaFuncSetAttribute(k_forest<GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KLaunch k(c, P_FOREST);
        k_forest<GW><<<c->sm_count * 2, 32 * GW, smem, c->stream>>>(f->dev, cb.pix, cb.featT, cb.counter,
                                                                    b->d_jobs, b->lower, w, b->ndp);
        NLF_CUDA(cudaGetLastError());
        return NLF_OK;
    }
    static int cfg = -1;
    if (cfg < 0) {
        const char *e = getenv("NLF_FOREST_CFG");              // tuning knob: default pipe32 with 2 trees per lane in flight;
                                                               // pipe32x1, pipe32x3, pipe16, wpipe32, wpipe16, dyn32, dyn16
        cfg = 0;
        if (e) {
            if (!strcmp(e, "pipe16")) cfg = 1;
            else if (!strcmp(e, "dyn32")) cfg = 2;
            else if (!strcmp(e, "dyn16")) cfg = 3;
            else if (!strcmp(e, "wpipe32")) cfg = 4;
            else if (!strcmp(e, "wpipe16")) cfg = 5;
            else if (!strcmp(e, "pipe32x1")) cfg = 6;
            else if (!strcmp(e, "pipe32x3")) cfg = 7;
        }
    }
    const int nparts = (int)part_begin.size() - 1;
    double *acc = nullptr;
    if (nparts > 1) NLF_CUDA(cudaMallocAsync(&acc, sizeof(double) * (size_t)cb.max_slots, c->stream));
    for (int p = 0; p < nparts; p++) {
        int t0 = part_begin[p], t1 = part_begin[p + 1];
        size_t smem = forest_dyn_smem((size_t)(f->off[t1] - f->off[t0]), t1 - t0, F);
        KLaunch k(c, P_FOREST);
#define NLF_LAUNCH_COOP(KERN, A, ...)                                                                               \
        do {                                                                                                        \
            auto kern = KERN<A, ##__VA_ARGS__>;                                                                     \
            NLF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                        \
                                          (int)(smem_max - 64)));                                                   \
            kern<<<c->sm_count, 32 * A, smem, c->stream>>>(                                                         \
                f->d_nodes2, f->d_off, t0, t1, T, F, cb.pix, cb.nanmask, cb.featT, cb.counter,                      \
                b->d_jobs, b->lower, w, b->ndp, acc, p == 0, p == nparts - 1);                                      \
        } while (0)
        switch (cfg) {
            case 1: NLF_LAUNCH_COOP(k_forest_pipe, 16, false); break;
            case 2: NLF_LAUNCH_COOP(k_forest_dyn, 32); break;
            case 3: NLF_LAUNCH_COOP(k_forest_dyn, 16); break;
            case 4: NLF_LAUNCH_COOP(k_forest_pipe, 32, true); break;
            case 5: NLF_LAUNCH_COOP(k_forest_pipe, 16, true); break;
            case 6: NLF_LAUNCH_COOP(k_forest_pipe, 32, false, 1); break;
            case 7: NLF_LAUNCH_COOP(k_forest_pipe, 32, false, 3); break;
            default: NLF_LAUNCH_COOP(k_forest_pipe, 32, false, 2); break;
        }
#undef NLF_LAUNCH_COOP
        NLF_CUDA(cudaGetLastError());
    }
    {
        // pixels with NaN features (rare: constant windows) through the general kernel
        KLaunch k(c, P_FOREST);
        k_forest_nan<<<c->sm_count * 4, 256, 0, c->stream>>>(f->dev, cb.pix, cb.nanmask, cb.featT,
                                                             cb.counter, b->d_jobs, b->lower, w, b->ndp);
        NLF_CUDA(cudaGetLastError());
    }
    if (acc) NLF_CUDA(cudaFreeAsync(acc, c->stream));
    return NLF_OK;
}

// upload the per-job descriptors (pointers into the batch pools; NULL pools stay NULL)
static int upload_jobs(nlf_batch *b, int w, int tx = 0)
{
    nlf_ctx *c = b->ctx;
    const int nj = (int)b->jobs.size();
    const int TX = b->TX = (tx > 0 ? tx : feat_tx(w));   // FeatCfg<W, TX>::TX
    const int up1 = b->upper + 2;
    std::vector<JobDev> jd(nj);
    std::vector<int> tb(nj + 1);
    int tiles = 0;
    long long groups = 0;
    for (int k = 0; k < nj; k++) {
        HostJob &j = b->jobs[k];
        JobDev &d = jd[k];
        double *sc = b->d_scratch + (size_t)k * 2 * up1;
        d.band = j.d_band;
        d.prob = b->d_prob ? b->d_prob + j.prob_off : nullptr;
        d.slot = b->d_slot ? b->d_slot + j.prob_off : nullptr;
        d.means = sc; d.e = sc + up1;
        d.status = b->d_status + k;
        d.counts = b->d_counts + (size_t)k * 4;
        d.don_job = b->d_don_job; d.don_i = b->d_don_i; d.don_j = b->d_don_j; d.don_p = b->d_don_p; d.don_v = b->d_don_v;
        d.don_count = b->d_counts + (size_t)NLF_MAX_JOBS * 4;
        d.don_cap = b->don_cap;
        d.job = k;
        d.dshift = dshift_of(b);
        d.n = j.n;
        d.ntx = (j.n + TX - 1) / TX;
        tb[k] = tiles;
        tiles += d.ntx * std::max(1, b->ndt);
        groups += (long long)d.ntx * TX * std::max(1, b->ndt);
    }
    tb[nj] = tiles;
    if (b->jobs_cap < nj) {
        if (b->d_jobs) { cudaFreeAsync(b->d_jobs, c->stream); cudaFreeAsync(b->d_tile_base, c->stream); }
        NLF_CUDA(cudaMallocAsync(&b->d_jobs, sizeof(JobDev) * nj, c->stream));
        NLF_CUDA(cudaMallocAsync(&b->d_tile_base, sizeof(int) * (nj + 1), c->stream));
        b->jobs_cap = nj;
    }
    // pageable sources: the copies are staged before these calls return
    NLF_CUDA(cudaMemcpyAsync(b->d_jobs, jd.data(), sizeof(JobDev) * nj, cudaMemcpyHostToDevice, c->stream));
    NLF_CUDA(cudaMemcpyAsync(b->d_tile_base, tb.data(), sizeof(int) * (nj + 1), cudaMemcpyHostToDevice, c->stream));
    b->total_tiles = tiles;
    b->total_groups = groups;
    return NLF_OK;
}

// K5: diagonal means, isotonic expected, candidate count.  Independent of the model.
static int batch_prepare(nlf_batch *b)
{
    NLF_CUDA(wait_inputs(b));
    if (b->prepared) return NLF_OK;
    nlf_ctx *c = b->ctx;
    const int nj = (int)b->jobs.size();
    if (nj == 0) { b->prepared = true; return NLF_OK; }
    if (b->scratch_jobs < nj) {
        if (b->d_scratch) cudaFreeAsync(b->d_scratch, c->stream);
        NLF_CUDA(cudaMallocAsync(&b->d_scratch, sizeof(double) * (size_t)nj * 2 * (b->upper + 2), c->stream));
        b->scratch_jobs = nj;
    }
    int rc = upload_jobs(b, 0);
    if (rc) return rc;
    {
        KLaunch k(c, P_DIAG);
        dim3 grid((b->upper - b->lower + 1 + 31) / 32, nj);
        k_diag_means<<<grid, 256, 0, c->stream>>>(b->d_jobs, b->lower, b->upper, b->pitch);
    }
    {
        KLaunch k(c, P_ISO);
        k_isotonic<<<nj, 128, (size_t)(b->upper + 2) * 20, c->stream>>>(b->d_jobs, b->lower, b->upper);
    }
    {
        KLaunch k(c, P_COUNT);
        dim3 grid(64, nj);
        k_count_candidates<<<grid, 256, 0, c->stream>>>(b->d_jobs, b->lower, b->upper, b->pitch);
    }
    NLF_CUDA(cudaGetLastError());
    b->prepared = true;
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_score(nlf_batch *b, nlf_forest *f, const double *expected, int nexp, int w,
                               const double *gw, int radius, double thre)
{
    if (!b || !f || !expected || !gw) return fail(NLF_ERR_ARG, "nlf_batch_score: NULL argument");
    if (b->scored) return fail(NLF_ERR_STATE, "batch already scored");
    if (radius != 4) return fail(NLF_ERR_ARG, "gaussian radius %d not supported (sigma=1 -> radius 4)", radius);
    if (w != 5 && w != 6 && w != 7) return fail(NLF_ERR_ARG, "window half-size w=%d not supported (5, 6, 7)", w);
    if (f->dev.n_features != (2 * w + 1) * (2 * w + 1))
        return fail(NLF_ERR_ARG, "forest has %d features, window w=%d needs %d", f->dev.n_features, w, (2 * w + 1) * (2 * w + 1));
    if (nexp <= 0) return fail(NLF_ERR_ARG, "empty expected array");
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    const int nj = (int)b->jobs.size();
    if (nj == 0) { b->scored = true; return NLF_OK; }
    int rc = batch_prepare(b);
    if (rc) return rc;
    const int dlo = std::max(b->lower, w + 2);
    b->W = w;
    b->ndt = (b->upper - dlo + 1 + 31) / 32;
    if (b->ndt < 1) b->ndt = 1;
    b->ndp = b->ndt * 32;
    const int F = (2 * w + 1) * (2 * w + 1);

    NLF_CUDA(cudaMemcpyToSymbolAsync(c_gw, gw, sizeof(double) * 9, 0, cudaMemcpyHostToDevice, c->stream));
    NLF_CUDA(cudaMallocAsync(&b->d_exp, sizeof(double) * nexp, c->stream));
    NLF_CUDA(cudaMemcpyAsync(b->d_exp, expected, sizeof(double) * nexp, cudaMemcpyHostToDevice, c->stream));

    size_t cells = 0;
    long long band_cells = 0;
    for (int k = 0; k < nj; k++) {
        HostJob &j = b->jobs[k];
        j.prob_off = cells;
        cells += (size_t)j.n * b->ndp;
        band_cells += (long long)j.n * (b->upper - dlo + 1);
    }
    NLF_CUDA(c->big_alloc((void **)&b->d_prob, sizeof(double) * cells));
    if (b->keep_slots) NLF_CUDA(c->big_alloc((void **)&b->d_slot, sizeof(int) * cells));
    b->don_cap = std::max(1LL, std::min(band_cells, 1LL << 23));
    NLF_CUDA(cudaMallocAsync(&b->d_don_job, sizeof(int) * b->don_cap, c->stream));
    NLF_CUDA(cudaMallocAsync(&b->d_don_i, sizeof(int) * b->don_cap, c->stream));
    NLF_CUDA(cudaMallocAsync(&b->d_don_j, sizeof(int) * b->don_cap, c->stream));
    NLF_CUDA(cudaMallocAsync(&b->d_don_p, sizeof(double) * b->don_cap, c->stream));
    NLF_CUDA(cudaMallocAsync(&b->d_don_v, sizeof(double) * b->don_cap, c->stream));
    ScorePlan plan;
    rc = make_plan(b, f, w, F, &plan);
    if (rc) return rc;
    rc = upload_jobs(b, w, plan.tx);
    if (rc) return rc;
    // Tiles are processed in chunks: every chunk goes through the feature kernel and then the forest kernel
    // (results land in `prob` by pixel coordinates).  With the compact forest the two kernels of DIFFERENT
    // chunks run at the same time: features of chunk k+1 on the main stream, traversal of chunk k on the aux
    // stream, CTAs of both on every SM (fp64 pipe and LSU are different units).  The chunks rotate through
    // NBUF feature buffers whose total worst-case footprint (TX*32 slots of F floats per tile) stays below
    // the budget; a buffer is reused once the forest kernel that read it has finished.
    const int tiles = b->total_tiles;
    const int TXh = b->TX;
    size_t budget_mb = 12288;
    if (const char *e = getenv("NLF_FEATURE_BUDGET_MB")) budget_mb = (size_t)std::max(1, atoi(e));
    const size_t slot_bytes = (size_t)F * 4 + sizeof(PixRef) + 8;
    const size_t tile_bytes = slot_bytes * TXh * 32;
    int nbuf = plan.coop ? 3 : 1;
    long long tiles_per_chunk = (long long)((budget_mb << 20) / (tile_bytes * nbuf));
    if (plan.coop) {
        long long target = std::max<long long>((tiles + 7) / 8, 4LL * c->sm_count);
        if (const char *e = getenv("NLF_CHUNK_TILES")) target = std::max(1, atoi(e));
        tiles_per_chunk = std::min(tiles_per_chunk, target);
    }
    tiles_per_chunk = std::max(1LL, std::min<long long>(tiles_per_chunk, tiles));
    const int n_chunks = (int)((tiles + tiles_per_chunk - 1) / tiles_per_chunk);
    if (n_chunks > 1 && b->keep_slots)
        return fail(NLF_ERR_ARG, "batch needs %d feature chunks: features cannot be kept (raise NLF_FEATURE_BUDGET_MB or split the batch)", n_chunks);
    nbuf = std::min(nbuf, n_chunks);
    const long long groups = tiles_per_chunk * TXh;            // 32 slots per group, TX*32 slots per tile
    b->max_groups = (unsigned)groups;
    NLF_CUDA(c->big_alloc((void **)&b->d_pix, sizeof(PixRef) * (size_t)groups * 32 * nbuf));
    NLF_CUDA(c->big_alloc((void **)&b->d_featT, sizeof(float) * (size_t)groups * F * 32 * nbuf));
    NLF_CUDA(cudaMallocAsync(&b->d_group_counter, sizeof(unsigned) * n_chunks, c->stream));
    NLF_CUDA(c->big_alloc((void **)&b->d_nanmask, sizeof(unsigned) * (size_t)(groups + 1) * nbuf));
    NLF_CUDA(cudaMallocAsync(&b->d_overflow, sizeof(int), c->stream));
    NLF_CUDA(cudaMemsetAsync(b->d_overflow, 0, sizeof(int), c->stream));
    NLF_CUDA(cudaMemsetAsync(b->d_group_counter, 0, sizeof(unsigned) * n_chunks, c->stream));
    const bool two_streams = plan.coop && n_chunks > 1;
    if (two_streams) NLF_CUDA(c->aux_init());
    cudaEvent_t region_a = nullptr;
    if (c->profile) {
        region_a = c->get_event();
        NLF_CUDA(cudaEventRecord(region_a, c->stream));
    }
    for (int ch = 0; ch < n_chunks; ch++) {
        const int t_off = (int)(ch * tiles_per_chunk);
        const int t_cnt = (int)std::min<long long>(tiles_per_chunk, tiles - t_off);
        const int set = ch % nbuf;
        ChunkBuf cb;
        cb.pix = b->d_pix + (size_t)set * groups * 32;
        cb.featT = b->d_featT + (size_t)set * groups * F * 32;
        cb.nanmask = b->d_nanmask + (size_t)set * (groups + 1);
        cb.counter = b->d_group_counter + ch;
        cb.max_slots = (unsigned)(groups * 32);
        cb.slot_base = (unsigned)((size_t)set * groups * 32);
        if (two_streams && ch >= nbuf) NLF_CUDA(cudaStreamWaitEvent(c->stream, c->ev_forest[set], 0));   // buffer `set` is free again
        NLF_CUDA(cudaMemsetAsync(cb.nanmask, 0, sizeof(unsigned) * (size_t)(groups + 1), c->stream));
        rc = launch_features(b, w, TXh, t_off, t_cnt, nexp, cb);
        if (rc) return rc;
        NLF_CUDA(cudaGetLastError());
        if (two_streams) {
            NLF_CUDA(cudaEventRecord(c->ev_feat[set], c->stream));
            NLF_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_feat[set], 0));
            rc = launch_forest_one(b, f, w, F, plan, cb, c->aux_stream);
            if (rc) return rc;
            NLF_CUDA(cudaEventRecord(c->ev_forest[set], c->aux_stream));
        } else if (plan.one) {
            rc = launch_forest_one(b, f, w, F, plan, cb, c->stream);
            if (rc) return rc;
        } else {
            rc = launch_forest(b, f, w, F, cb);
            if (rc) return rc;
        }
    }
    if (two_streams)                                           // the aux stream is in order: its last event covers every chunk
        NLF_CUDA(cudaStreamWaitEvent(c->stream, c->ev_forest[(n_chunks - 1) % nbuf], 0));
    if (region_a) {
        cudaEvent_t region_b = c->get_event();
        NLF_CUDA(cudaEventRecord(region_b, c->stream));
        c->recs.push_back({P_SCORE, region_a, region_b});
    }
    {
        KLaunch k(c, P_THRESH);
        dim3 grid(64, nj);
        k_threshold<<<grid, 256, 0, c->stream>>>(b->d_jobs, b->lower, b->upper, w, b->pitch, b->ndp, thre);
    }
    NLF_CUDA(cudaGetLastError());
    b->scored = true;
    b->counted = false;
    b->donuts_cached = false;
    return NLF_OK;
}

static int batch_sync_counts(nlf_batch *b)
{
    if (!b->prepared) {
        int rc0 = batch_prepare(b);
        if (rc0) return rc0;
    }
    if (b->counted) return NLF_OK;
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    const int nj = (int)b->jobs.size();
    std::vector<unsigned long long> cnt((size_t)nj * 4 + 1, 0ULL);
    std::vector<int> st(std::max(1, nj), 0);
    if (nj) {
        NLF_CUDA(cudaMemcpyAsync(cnt.data(), b->d_counts, sizeof(unsigned long long) * (size_t)nj * 4, cudaMemcpyDeviceToHost, c->stream));
        NLF_CUDA(cudaMemcpyAsync(st.data(), b->d_status, sizeof(int) * nj, cudaMemcpyDeviceToHost, c->stream));
    }
    NLF_CUDA(cudaMemcpyAsync(&cnt[(size_t)nj * 4], b->d_counts + (size_t)NLF_MAX_JOBS * 4, sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, c->stream));
    int overflow = 0;
    if (b->d_overflow)
        NLF_CUDA(cudaMemcpyAsync(&overflow, b->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    if (overflow) return fail(NLF_ERR_CUDA, "feature group buffer overflow (internal sizing error)");
    for (int k = 0; k < nj; k++) {
        memcpy(b->jobs[k].counts, &cnt[(size_t)k * 4], sizeof(unsigned long long) * 4);
        b->jobs[k].status = st[k];
    }
    b->n_donuts_total = cnt[(size_t)nj * 4];
    b->counted = true;
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_counts(nlf_batch *b, long long *n_cand, long long *n_scored, long long *n_donuts)
{
    int rc = batch_sync_counts(b);
    if (rc) return rc;
    int first_err = 0;
    for (size_t k = 0; k < b->jobs.size(); k++) {
        HostJob &j = b->jobs[k];
        if (j.status && !first_err) first_err = j.status;
        if (n_cand) n_cand[k] = j.status ? -1 : (long long)j.counts[0];
        if (n_scored) n_scored[k] = j.status ? -1 : (long long)j.counts[1];
        if (n_donuts) n_donuts[k] = j.status ? -1 : (long long)j.counts[2];
    }
    if (b->scored && (long long)b->n_donuts_total > b->don_cap)
        return fail(NLF_ERR_ARG, "%llu pixels above the threshold exceed the Donut capacity %lld of the batch; raise the threshold or split the batch",
                    b->n_donuts_total, b->don_cap);
    if (first_err == NLF_ERR_EMPTY_FIT)
        return fail(NLF_ERR_EMPTY_FIT, "a matrix has no diagonal longer than 5 bins: the isotonic fit of get_candidates is empty");
    if (first_err == NLF_ERR_LOWER_TRI)
        return fail(NLF_ERR_LOWER_TRI, "a matrix has non-zero entries below the main diagonal (np.triu input expected)");
    if (first_err) return fail(first_err, "job failed with status %d", first_err);
    return NLF_OK;
}

// one device-to-host copy of the whole Donut list, binned by job on the host
static int cache_donuts(nlf_batch *b)
{
    if (b->donuts_cached) return NLF_OK;
    nlf_ctx *c = b->ctx;
    const size_t m = (size_t)std::min<unsigned long long>(b->n_donuts_total, (unsigned long long)b->don_cap);
    b->h_don_job.resize(m); b->h_don_i.resize(m); b->h_don_j.resize(m); b->h_don_p.resize(m); b->h_don_v.resize(m);
    if (m) {
        NLF_CUDA(cudaMemcpyAsync(b->h_don_job.data(), b->d_don_job, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
        NLF_CUDA(cudaMemcpyAsync(b->h_don_i.data(), b->d_don_i, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
        NLF_CUDA(cudaMemcpyAsync(b->h_don_j.data(), b->d_don_j, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
        NLF_CUDA(cudaMemcpyAsync(b->h_don_p.data(), b->d_don_p, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
        NLF_CUDA(cudaMemcpyAsync(b->h_don_v.data(), b->d_don_v, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
        NLF_CUDA(cudaStreamSynchronize(c->stream));
    }
    b->don_by_job.assign(b->jobs.size(), std::vector<int>());
    for (size_t k = 0; k < m; k++) b->don_by_job[(size_t)b->h_don_job[k]].push_back((int)k);
    b->donuts_cached = true;
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_fetch_donuts(nlf_batch *b, int job, int *i, int *j, double *prob, double *value)
{
    if (!b->scored) return fail(NLF_ERR_STATE, "batch not scored yet");
    int rc = batch_sync_counts(b);
    if (rc) return rc;
    if (job < 0 || job >= (int)b->jobs.size()) return fail(NLF_ERR_ARG, "job %d out of range", job);
    HostJob &J = b->jobs[job];
    if (J.status) return fail(J.status, "job %d failed with status %d", job, J.status);
    if ((long long)b->n_donuts_total > b->don_cap) return fail(NLF_ERR_ARG, "Donut capacity exceeded");
    rc = cache_donuts(b);
    if (rc) return rc;
    const std::vector<int> &idx = b->don_by_job[(size_t)job];
    for (size_t k = 0; k < idx.size(); k++) {
        i[k] = b->h_don_i[idx[k]];
        j[k] = b->h_don_j[idx[k]];
        prob[k] = b->h_don_p[idx[k]];
        value[k] = b->h_don_v[idx[k]];
    }
    return NLF_OK;
}

// Per-diagonal means (NaN where the diagonal has <= 5 entries) and the isotonic expected of
// Peakachu.get_candidates for one job: arrays of upper+1 doubles indexed by the diagonal.
NLF_EXPORT int nlf_batch_fetch_expected(nlf_batch *b, int job, double *means, double *e)
{
    if (!b || job < 0 || job >= (int)b->jobs.size()) return fail(NLF_ERR_ARG, "nlf_batch_fetch_expected: bad job");
    int rc = batch_sync_counts(b);
    if (rc) return rc;
    nlf_ctx *c = b->ctx;
    const int up1 = b->upper + 2;
    const double *sc = b->d_scratch + (size_t)job * 2 * up1;
    if (means) NLF_CUDA(cudaMemcpyAsync(means, sc, sizeof(double) * (b->upper + 1), cudaMemcpyDeviceToHost, c->stream));
    if (e) NLF_CUDA(cudaMemcpyAsync(e, sc + up1, sizeof(double) * (b->upper + 1), cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    return NLF_OK;
}

// The whole Donut list of the batch at once (arrays of n_total entries, any order): job, i, j, p, v.
NLF_EXPORT int nlf_batch_fetch_donuts_all(nlf_batch *b, long long cap, int *job, int *i, int *j, double *prob,
                                          double *value, long long *n_total)
{
    if (!b || !n_total) return fail(NLF_ERR_ARG, "nlf_batch_fetch_donuts_all: NULL argument");
    if (!b->scored) return fail(NLF_ERR_STATE, "batch not scored yet");
    int rc = batch_sync_counts(b);
    if (rc) return rc;
    if ((long long)b->n_donuts_total > b->don_cap) return fail(NLF_ERR_ARG, "Donut capacity exceeded");
    *n_total = (long long)b->n_donuts_total;
    if (!job) return NLF_OK;                         // size query
    if (cap < (long long)b->n_donuts_total) return fail(NLF_ERR_ARG, "capacity %lld < %llu Donuts", cap, b->n_donuts_total);
    const size_t m = (size_t)b->n_donuts_total;
    if (m == 0) return NLF_OK;
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaMemcpyAsync(job, b->d_don_job, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaMemcpyAsync(i, b->d_don_i, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaMemcpyAsync(j, b->d_don_j, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaMemcpyAsync(prob, b->d_don_p, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaMemcpyAsync(value, b->d_don_v, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    return NLF_OK;
}

// ordered listing helper (mode 0 candidates, 1 scored)
static int ordered_list(nlf_batch *b, int job, int mode, long long cap, int *oi, int *oj, double *op, float *ofea)
{
    if (mode == 1 && !b->scored) return fail(NLF_ERR_STATE, "batch not scored yet");
    int rc = batch_sync_counts(b);
    if (rc) return rc;
    if (job < 0 || job >= (int)b->jobs.size()) return fail(NLF_ERR_ARG, "job %d out of range", job);
    HostJob &J = b->jobs[job];
    if (J.status) return fail(J.status, "job %d failed with status %d", job, J.status);
    nlf_ctx *c = b->ctx;
    const int nd = b->upper - b->lower + 1;
    JobDev d;
    memset(&d, 0, sizeof(d));
    d.band = J.d_band; d.prob = b->d_prob ? b->d_prob + J.prob_off : nullptr;
    d.slot = b->d_slot ? b->d_slot + J.prob_off : nullptr;
    d.e = b->d_scratch + (size_t)job * 2 * (b->upper + 2) + (b->upper + 2); d.n = J.n; d.status = b->d_status + job;
    int *d_cnt;
    long long *d_off;
    NLF_CUDA(cudaMallocAsync(&d_cnt, sizeof(int) * nd, c->stream));
    NLF_CUDA(cudaMallocAsync(&d_off, sizeof(long long) * nd, c->stream));
    {
        KLaunch k(c, P_THRESH);
        k_diag_count<<<nd, 256, 0, c->stream>>>(d, mode, b->lower, b->upper, b->W, b->pitch, b->ndp, d_cnt);
    }
    std::vector<int> cnt(nd);
    NLF_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * nd, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<long long> off(nd);
    long long tot = 0;
    for (int k = 0; k < nd; k++) { off[k] = tot; tot += cnt[k]; }
    if (tot > cap) return fail(NLF_ERR_ARG, "output capacity %lld too small for %lld rows", cap, tot);
    if (tot > 0) {
        int *d_i, *d_j, *d_slot = nullptr;
        double *d_p = nullptr;
        NLF_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * nd, cudaMemcpyHostToDevice, c->stream));
        NLF_CUDA(cudaMallocAsync(&d_i, sizeof(int) * tot, c->stream));
        NLF_CUDA(cudaMallocAsync(&d_j, sizeof(int) * tot, c->stream));
        if (op) NLF_CUDA(cudaMallocAsync(&d_p, sizeof(double) * tot, c->stream));
        if (ofea) NLF_CUDA(cudaMallocAsync(&d_slot, sizeof(int) * tot, c->stream));
        {
            KLaunch k(c, P_THRESH);
            k_diag_fill<<<nd, 256, 0, c->stream>>>(d, mode, b->lower, b->upper, b->W, b->pitch, b->ndp, d_off, d_i, d_j,
                                                   d_p, d_slot);
        }
        NLF_CUDA(cudaGetLastError());
        if (oi) NLF_CUDA(cudaMemcpyAsync(oi, d_i, sizeof(int) * tot, cudaMemcpyDeviceToHost, c->stream));
        if (oj) NLF_CUDA(cudaMemcpyAsync(oj, d_j, sizeof(int) * tot, cudaMemcpyDeviceToHost, c->stream));
        if (op) NLF_CUDA(cudaMemcpyAsync(op, d_p, sizeof(double) * tot, cudaMemcpyDeviceToHost, c->stream));
        if (ofea) {
            const int F = (2 * b->W + 1) * (2 * b->W + 1);
            float *d_f;
            NLF_CUDA(cudaMallocAsync(&d_f, sizeof(float) * (size_t)tot * F, c->stream));
            {
                KLaunch k(c, P_THRESH);
                long long work = tot * F;
                k_gather_features<<<(unsigned)((work + 255) / 256), 256, 0, c->stream>>>(b->d_featT, d_slot, tot, F, d_f);
            }
            NLF_CUDA(cudaGetLastError());
            NLF_CUDA(cudaMemcpyAsync(ofea, d_f, sizeof(float) * (size_t)tot * F, cudaMemcpyDeviceToHost, c->stream));
            NLF_CUDA(cudaFreeAsync(d_f, c->stream));
            NLF_CUDA(cudaFreeAsync(d_slot, c->stream));
        }
        NLF_CUDA(cudaStreamSynchronize(c->stream));
        NLF_CUDA(cudaFreeAsync(d_i, c->stream));
        NLF_CUDA(cudaFreeAsync(d_j, c->stream));
        if (d_p) NLF_CUDA(cudaFreeAsync(d_p, c->stream));
    }
    NLF_CUDA(cudaFreeAsync(d_cnt, c->stream));
    NLF_CUDA(cudaFreeAsync(d_off, c->stream));
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_fetch_scored(nlf_batch *b, int job, long long cap, int *i, int *j, double *prob)
{
    return ordered_list(b, job, 1, cap, i, j, prob, nullptr);
}

NLF_EXPORT int nlf_batch_fetch_candidates(nlf_batch *b, int job, long long cap, int *ridx, int *cidx)
{
    return ordered_list(b, job, 0, cap, ridx, cidx, nullptr, nullptr);
}

NLF_EXPORT int nlf_batch_fetch_features32(nlf_batch *b, int job, long long cap, float *fea)
{
    if (!b->keep_slots) return fail(NLF_ERR_STATE, "feature slots were not kept: call nlf_batch_keep_features before scoring");
    return ordered_list(b, job, 1, cap, nullptr, nullptr, nullptr, fea);
}

NLF_EXPORT int nlf_batch_keep_features(nlf_batch *b, int on)
{
    if (b->scored) return fail(NLF_ERR_STATE, "batch already scored");
    b->keep_slots = on != 0;
    return NLF_OK;
}

// util.distance_normalize (neoloop/util.py:499-524) + distance_normaize_core (:470-491) on a stack of
// windows handed over by the caller: NaN -> 0, count_nonzero and lower-left-mean gates, O/E division.
// One warp per window; the w x w mean is numpy's (8-accumulator) order, computed by one lane.
__global__ void __launch_bounds__(128) k_distance_normalize(const double *__restrict__ win, long long N, const int *__restrict__ xi,
                                                            const int *__restrict__ yi, int w, const double *__restrict__ expv,
                                                            int nexp, double *__restrict__ out, unsigned char *__restrict__ pass)
{
    __shared__ double tmp[4][15 * 15];
    __shared__ double s_ll[4];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const long long k = (long long)blockIdx.x * 4 + wp;
    if (k >= N) return;
    const int S = 2 * w + 1, F = S * S;
    const double *src = win + (size_t)k * F;
    double *dst = out + (size_t)k * F;
    int nz = 0;
    for (int c = lane; c < F; c += 32) {
        double v = src[c];
        if (v != v) v = 0.0;
        tmp[wp][c] = v;
        nz += (v != 0.0);
    }
    for (int o = 16; o; o >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, o);
    __syncwarp();
    bool ok = !((double)nz < __dmul_rn((double)F, 0.1));
    if (ok) {
        if (lane == 0) {
            const double *t = tmp[wp];
            auto ld = [&](int q) { return t[(q / w) * S + (q % w)]; };
            s_ll[wp] = __ddiv_rn((w * w <= 128) ? np_leaf_sum(ld, 0, w * w) : np_pairwise_sum(ld, w * w), (double)(w * w));
        }
        __syncwarp();
        const double ll = s_ll[wp];
        ok = (ll > 0) && (__ddiv_rn(tmp[wp][w * S + w], ll) > 0.1);
    }
    if (ok) {
        const int x = xi[k], y = yi[k];
        int dmax = abs(y - x) + 2 * w;                          // largest |column - row| inside the window
        const bool normalise = dmax < nexp;
        for (int c = lane; c < F; c += 32) {
            const int a = c / S, b = c % S;
            const int d = abs((y - w + b) - (x - w + a));
            const double v = tmp[wp][c];
            dst[c] = normalise ? __ddiv_rn(v, expv[d]) : v;
        }
    } else {
        for (int c = lane; c < F; c += 32) dst[c] = tmp[wp][c];
    }
    if (lane == 0) pass[k] = ok ? 1 : 0;
}

NLF_EXPORT int nlf_distance_normalize(nlf_ctx *ctx, const double *windows, long long N, const int *xi, const int *yi,
                                      int w, const double *expected, int nexp, double *out, uint8_t *pass)
{
    if (!ctx || N < 0 || w < 1 || w > 7 || nexp <= 0 || !expected || (N > 0 && (!windows || !xi || !yi || !out || !pass)))
        return fail(NLF_ERR_ARG, "nlf_distance_normalize: bad arguments (w must be 1..7)");
    if (N == 0) return NLF_OK;
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t F = (size_t)(2 * w + 1) * (2 * w + 1);
    double *d_win, *d_out, *d_exp;
    int *d_xy;
    unsigned char *d_pass;
    NLF_CUDA(c->big_alloc((void **)&d_win, sizeof(double) * F * N));
    NLF_CUDA(c->big_alloc((void **)&d_out, sizeof(double) * F * N));
    NLF_CUDA(cudaMallocAsync(&d_exp, sizeof(double) * nexp, s));
    NLF_CUDA(cudaMallocAsync(&d_xy, sizeof(int) * 2 * N, s));
    NLF_CUDA(cudaMallocAsync(&d_pass, N, s));
    NLF_CUDA(cudaMemcpyAsync(d_win, windows, sizeof(double) * F * N, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_exp, expected, sizeof(double) * nexp, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_xy, xi, sizeof(int) * N, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_xy + N, yi, sizeof(int) * N, cudaMemcpyHostToDevice, s));
    {
        KLaunch k(c, P_FEATURE);
        k_distance_normalize<<<(unsigned)((N + 3) / 4), 128, 0, s>>>(d_win, N, d_xy, d_xy + N, w, d_exp, nexp, d_out, d_pass);
    }
    NLF_CUDA(cudaGetLastError());
    NLF_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * F * N, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(pass, d_pass, N, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    c->big_free(d_win);
    c->big_free(d_out);
    NLF_CUDA(cudaFreeAsync(d_exp, s));
    NLF_CUDA(cudaFreeAsync(d_xy, s));
    NLF_CUDA(cudaFreeAsync(d_pass, s));
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_getwindow(nlf_batch *b, int job, const int *xi, const int *yi, long long nc, int w,
                                   const double *expected, int nexp, const double *gw, int radius, double *fea,
                                   int *clist, long long *n_rows)
{
    if (!b || job < 0 || job >= (int)b->jobs.size()) return fail(NLF_ERR_ARG, "nlf_batch_getwindow: bad job");
    if (radius != 4) return fail(NLF_ERR_ARG, "gaussian radius %d not supported", radius);
    if (w != 5 && w != 6 && w != 7) return fail(NLF_ERR_ARG, "w=%d not supported", w);
    *n_rows = 0;
    if (nc < 2) return NLF_OK;          // callers.py:587-588
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    HostJob &J = b->jobs[job];
    const int F = (2 * w + 1) * (2 * w + 1);
    // callers.py:592-595: fewer than two coordinates inside the border mask -> nothing
    long long inside = 0;
    for (long long k = 0; k < nc; k++)
        inside += (xi[k] - w >= 0) && (yi[k] + w + 1 <= J.n) && (yi[k] - xi[k] - 2 >= w);
    if (inside < 2) return NLF_OK;
    for (long long k = 0; k < nc; k++)
        if (xi[k] < 0 || yi[k] < 0 || xi[k] >= J.n || yi[k] >= J.n) return fail(NLF_ERR_ARG, "coordinate %lld out of range", k);
    NLF_CUDA(wait_inputs(b));
    int *d_x, *d_y;
    unsigned char *d_pass;
    double *d_fea, *d_exp;
    NLF_CUDA(cudaMallocAsync(&d_x, sizeof(int) * nc, c->stream));
    NLF_CUDA(cudaMallocAsync(&d_y, sizeof(int) * nc, c->stream));
    NLF_CUDA(cudaMallocAsync(&d_pass, nc, c->stream));
    NLF_CUDA(cudaMallocAsync(&d_fea, sizeof(double) * (size_t)nc * F, c->stream));
    NLF_CUDA(cudaMallocAsync(&d_exp, sizeof(double) * nexp, c->stream));
    NLF_CUDA(cudaMemcpyAsync(d_x, xi, sizeof(int) * nc, cudaMemcpyHostToDevice, c->stream));
    NLF_CUDA(cudaMemcpyAsync(d_y, yi, sizeof(int) * nc, cudaMemcpyHostToDevice, c->stream));
    NLF_CUDA(cudaMemcpyAsync(d_exp, expected, sizeof(double) * nexp, cudaMemcpyHostToDevice, c->stream));
    NLF_CUDA(cudaMemcpyToSymbolAsync(c_gw, gw, sizeof(double) * 9, 0, cudaMemcpyHostToDevice, c->stream));
    JobDev d;
    memset(&d, 0, sizeof(d));
    d.band = J.d_band; d.n = J.n;
    {
        KLaunch k(c, P_FEATURE);
        unsigned grid = (unsigned)((nc + 3) / 4);
        if (w == 5) k_getwindow<5><<<grid, 128, 0, c->stream>>>(d, b->upper, b->pitch, d_x, d_y, nc, d_exp, nexp, d_pass, d_fea);
        else if (w == 6) k_getwindow<6><<<grid, 128, 0, c->stream>>>(d, b->upper, b->pitch, d_x, d_y, nc, d_exp, nexp, d_pass, d_fea);
        else k_getwindow<7><<<grid, 128, 0, c->stream>>>(d, b->upper, b->pitch, d_x, d_y, nc, d_exp, nexp, d_pass, d_fea);
    }
    NLF_CUDA(cudaGetLastError());
    std::vector<unsigned char> pass((size_t)nc);
    std::vector<double> all((size_t)nc * F);
    NLF_CUDA(cudaMemcpyAsync(pass.data(), d_pass, nc, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaMemcpyAsync(all.data(), d_fea, sizeof(double) * (size_t)nc * F, cudaMemcpyDeviceToHost, c->stream));
    NLF_CUDA(cudaStreamSynchronize(c->stream));
    long long rows = 0;
    for (long long k = 0; k < nc; k++)
        if (pass[k]) {
            memcpy(fea + (size_t)rows * F, all.data() + (size_t)k * F, sizeof(double) * F);
            clist[2 * rows] = xi[k];
            clist[2 * rows + 1] = yi[k];
            rows++;
        }
    *n_rows = rows;
    NLF_CUDA(cudaFreeAsync(d_x, c->stream));
    NLF_CUDA(cudaFreeAsync(d_y, c->stream));
    NLF_CUDA(cudaFreeAsync(d_pass, c->stream));
    NLF_CUDA(cudaFreeAsync(d_fea, c->stream));
    NLF_CUDA(cudaFreeAsync(d_exp, c->stream));
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_rethreshold(nlf_batch *b, double thre)
{
    if (!b) return fail(NLF_ERR_ARG, "nlf_batch_rethreshold: NULL batch");
    if (!b->scored) return fail(NLF_ERR_STATE, "batch not scored yet");
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    const int nj = (int)b->jobs.size();
    if (nj == 0) return NLF_OK;
    // counts[k] = {candidates, scored, donuts, -}: keep the candidates, rebuild the other two and the Donut list
    NLF_CUDA(cudaMemset2DAsync(b->d_counts + 1, 4 * sizeof(unsigned long long), 0, 2 * sizeof(unsigned long long), (size_t)nj, c->stream));
    NLF_CUDA(cudaMemsetAsync(b->d_counts + (size_t)NLF_MAX_JOBS * 4, 0, sizeof(unsigned long long), c->stream));
    {
        KLaunch k(c, P_THRESH);
        dim3 grid(64, nj);
        k_threshold<<<grid, 256, 0, c->stream>>>(b->d_jobs, b->lower, b->upper, b->W, b->pitch, b->ndp, thre);
    }
    NLF_CUDA(cudaGetLastError());
    b->counted = false;
    b->donuts_cached = false;
    return NLF_OK;
}

static void free_score_outputs(nlf_batch *b)
{
    cudaStream_t s = b->ctx->stream;
    auto fr = [&](auto *&p) { if (p) { cudaFreeAsync(p, s); p = nullptr; } };
    auto frb = [&](auto *&p) { if (p) { b->ctx->big_free(p); p = nullptr; } };
    frb(b->d_prob); frb(b->d_slot); frb(b->d_pix); frb(b->d_featT); frb(b->d_nanmask);
    fr(b->d_don_job); fr(b->d_don_i); fr(b->d_don_j); fr(b->d_don_p); fr(b->d_don_v);
    fr(b->d_group_counter); fr(b->d_overflow); fr(b->d_exp);
}

// Forget the results of a scored batch but keep the matrices resident in HBM, so that the
// whole scoring path (K5-K9) can run again on them.
NLF_EXPORT int nlf_batch_reset(nlf_batch *b)
{
    if (!b) return fail(NLF_ERR_ARG, "nlf_batch_reset: NULL batch");
    nlf_ctx *c = b->ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    free_score_outputs(b);
    const int nj = (int)b->jobs.size();
    if (nj) {
        // keep the statuses the input stage raised (lower-triangular input); the fit status is recomputed
        std::vector<int> st(nj);
        for (int k = 0; k < nj; k++) {
            st[k] = (b->jobs[k].status == NLF_ERR_LOWER_TRI) ? NLF_ERR_LOWER_TRI : 0;
            memset(b->jobs[k].counts, 0, sizeof(b->jobs[k].counts));
        }
        if (b->counted) NLF_CUDA(cudaMemcpyAsync(b->d_status, st.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, c->stream));
        NLF_CUDA(cudaMemsetAsync(b->d_counts, 0, sizeof(unsigned long long) * (size_t)nj * 4, c->stream));
    }
    NLF_CUDA(cudaMemsetAsync(b->d_counts + (size_t)NLF_MAX_JOBS * 4, 0, sizeof(unsigned long long), c->stream));
    b->prepared = b->scored = b->counted = false;
    b->donuts_cached = false;
    NLF_CUDA(c->copy_fence());                      // later inputs come after the resets above
    return NLF_OK;
}

// Evict the L2 (126 MB on B200) by overwriting a 256 MB scratch buffer on the compute stream.
NLF_EXPORT int nlf_l2_flush(nlf_ctx *c)
{
    NLF_CUDA(cudaSetDevice(c->device));
    static thread_local void *scratch[16] = {nullptr};
    const size_t bytes = (size_t)256 << 20;
    if (c->device >= 16) return fail(NLF_ERR_ARG, "device index too large");
    if (!scratch[c->device]) NLF_CUDA(cudaMalloc(&scratch[c->device], bytes));
    NLF_CUDA(cudaMemsetAsync(scratch[c->device], 0, bytes, c->stream));
    return NLF_OK;
}

// fp64 add/mul issue-rate micro-benchmark (no FMA): the roofline denominator of the feature
// kernel, measured in the same run.  16 independent chains per thread, DADD and DMUL alternate.
__global__ void k_fp64_peak(double *out, int iters)
{
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    const double m = 1.0000000001, s = 1e-12;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) { a[k] = __dmul_rn(a[k], m); a[k] = __dadd_rn(a[k], s); }
    }
    double t = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) t += a[k];
    if (t == 12345.678) out[0] = t;
}

NLF_EXPORT int nlf_fp64_peak(nlf_ctx *c, double *gops)
{
    NLF_CUDA(cudaSetDevice(c->device));
    double *d;
    NLF_CUDA(cudaMalloc(&d, 8));
    const int iters = 4096, threads = 256, blocks = c->sm_count * 8;
    cudaEvent_t a, b;
    NLF_CUDA(cudaEventCreate(&a));
    NLF_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        NLF_CUDA(cudaEventRecord(a, c->stream));
        k_fp64_peak<<<blocks, threads, 0, c->stream>>>(d, iters);
        NLF_CUDA(cudaEventRecord(b, c->stream));
        NLF_CUDA(cudaEventSynchronize(b));
        float ms;
        NLF_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 4;
    double ops = (double)blocks * threads * (double)iters * 32.0;
    *gops = ops / (best * 1e-3) / 1e9;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return NLF_OK;
}

NLF_EXPORT int nlf_batch_destroy(nlf_batch *b)
{
    if (!b) return NLF_OK;
    nlf_ctx *c = b->ctx;
    cudaSetDevice(c->device);
    cudaStream_t s = c->stream;
    wait_inputs(b);                                 // uploads in flight finish before their buffers are handed back
    if (b->ev_inputs) cudaEventDestroy(b->ev_inputs);
    free_score_outputs(b);
    for (auto &j : b->jobs)
        if (j.d_band) {
            if (j.band_big) c->big_free(j.d_band); else cudaFreeAsync(j.d_band, s);
        }
    if (b->d_status) cudaFreeAsync(b->d_status, s);
    if (b->d_counts) cudaFreeAsync(b->d_counts, s);
    if (b->d_scratch) cudaFreeAsync(b->d_scratch, s);
    if (b->d_jobs) cudaFreeAsync(b->d_jobs, s);
    if (b->d_tile_base) cudaFreeAsync(b->d_tile_base, s);
    delete b;
    return NLF_OK;
}
