// common.cuh -- shared host/device plumbing of libnlf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only; ranges are no-ops unless a tool (ncu --nvtx, nsys) is attached
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/nlf_b200.h"

#define NLF_EXPORT extern "C" __attribute__((visibility("default")))

namespace nlf {

extern thread_local char g_err[512];

inline int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define NLF_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return nlf::fail(NLF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                             cudaGetErrorString(e__));                                              \
    } while (0)

enum ProfClass { P_BANDIFY = 0, P_DIAG = 1, P_FEATURE = 2, P_FOREST = 3, P_THRESH = 4, P_NORM = 5, P_POOL = 6, P_SCORE = 7, P_ISO = 8, P_COUNT = 9, P_NCLASS = 11 };
// P_SCORE: the whole feature + forest region of one nlf_batch_score call (the two kernels run concurrently on two streams)

// numpy pairwise sum of ld(0..len-1) computed by an aligned group of 8 lanes (lane j = accumulator j
// of every leaf); every lane of the group returns the sum.  numpy splits n > 128 at n/2 - (n/2)%8, so
// leaves start at multiples of 8 and only the last one has a ragged tail; the 8-accumulator combine
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is an xor butterfly (addition commutes, the association is kept).
template <typename Load>
__device__ double np_pairwise_sum_group8(Load ld, int len, int j, unsigned gmask)
{
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
    if (len <= 128) return leaf(0, len);
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
    return vs[0];
}

}  // namespace nlf

struct nlf_ctx {
    // Host-side caches below (big buffers, staging ring, event pool, profiling records, launch counter) are guarded
    // by this mutex, so two host threads that end up on the same context (e.g. a device listed twice) cannot hand
    // one buffer to two users.  Work of one context is still ordered on its single main stream.
    std::recursive_mutex mu;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    long long launches = 0;
    bool profile = false;
    // profiling: (class, start, stop) triples recorded in-stream, resolved lazily
    struct Rec { int cls; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> free_events;
    float prof_ms[nlf::P_NCLASS] = {0};
    long long prof_n[nlf::P_NCLASS] = {0};
    int sm_count = 148;

    // Grow-only cache of large device buffers (>= 1 MB).  Every operation of a context is ordered on
    // its single stream, so a buffer handed back here can be reused by the next request at once;
    // steady-state steps then make no driver allocation at all (cudaMalloc of multi-GB feature
    // buffers costs tens of ms to seconds and showed up as latency spikes in the end-to-end path).
    struct Big { void *p; size_t bytes; bool busy; cudaEvent_t freed; };
    std::vector<Big> bigs;
    // `consumer`: the stream that will touch the buffer first.  Buffers are handed back in
    // main-stream order (big_free records an event there); another stream waits for that event.
    cudaError_t big_alloc(void **out, size_t bytes, cudaStream_t consumer = nullptr)
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        int best = -1;
        for (int i = 0; i < (int)bigs.size(); i++)
            if (!bigs[i].busy && bigs[i].bytes >= bytes && bigs[i].bytes <= bytes + bytes / 4 + (1 << 20) &&
                (best < 0 || bigs[i].bytes < bigs[best].bytes))
                best = i;
        if (best >= 0) {
            bigs[best].busy = true;
            *out = bigs[best].p;
            if (consumer && consumer != stream && bigs[best].freed) return cudaStreamWaitEvent(consumer, bigs[best].freed, 0);
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            // release idle cached buffers and retry once
            cudaGetLastError();
            cudaStreamSynchronize(stream);
            for (auto &b : bigs) if (!b.busy && b.p) { cudaFree(b.p); b.p = nullptr; if (b.freed) cudaEventDestroy(b.freed); }
            std::vector<Big> keep;
            for (auto &b : bigs) if (b.p) keep.push_back(b);
            bigs.swap(keep);
            e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) return e;
        }
        bigs.push_back({p, bytes, true, nullptr});
        *out = p;
        return cudaSuccess;
    }
    void big_free(void *p)
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        if (!p) return;
        for (auto &b : bigs)
            if (b.p == p) {
                b.busy = false;
                if (!b.freed) cudaEventCreateWithFlags(&b.freed, cudaEventDisableTiming);
                if (b.freed) cudaEventRecord(b.freed, stream);
                return;
            }
    }
    void big_release_all()
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        for (auto &b : bigs) { if (b.p) cudaFree(b.p); if (b.freed) cudaEventDestroy(b.freed); }
        bigs.clear();
    }

    // Host -> device uploads of job inputs run on a second stream (copy + the small kernel that
    // turns the upload into the resident band), through a ring of staging buffers, so the inputs
    // of the next job -- or of the next sub-batch -- arrive while kernels of earlier work occupy the
    // main stream.  The batch records an event after its last input; the main stream waits for it
    // before the first kernel that reads the bands.
    cudaStream_t copy_stream = nullptr;
    // forest traversal runs here (highest priority), concurrently with the feature kernel of the next chunk on `stream`
    cudaStream_t aux_stream = nullptr;
    static constexpr int NEVT = 4;
    cudaEvent_t ev_feat[NEVT] = {nullptr}, ev_forest[NEVT] = {nullptr};   // per feature-buffer set (re-recorded every use)
    cudaError_t aux_init()
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        if (aux_stream) return cudaSuccess;
        int lo = 0, hi = 0;
        cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithPriority(&aux_stream, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
        for (int i = 0; i < NEVT; i++) {
            if ((e = cudaEventCreateWithFlags(&ev_feat[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_forest[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    void aux_release()
    {
        if (!aux_stream) return;
        cudaStreamSynchronize(aux_stream);
        for (int i = 0; i < NEVT; i++) { cudaEventDestroy(ev_feat[i]); cudaEventDestroy(ev_forest[i]); }
        cudaStreamDestroy(aux_stream);
        aux_stream = nullptr;
    }
    cudaEvent_t fence = nullptr;
    struct Stage { void *p = nullptr; size_t bytes = 0; cudaEvent_t used = nullptr; bool pending = false; };
    static constexpr int NSTAGE = 3;
    Stage stages[NSTAGE];
    int stage_next = 0;
    // order the copy stream after everything enqueued on the main stream so far
    cudaError_t copy_fence()
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        cudaError_t e;
        if (!fence && (e = cudaEventCreateWithFlags(&fence, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(fence, stream)) != cudaSuccess) return e;
        return cudaStreamWaitEvent(copy_stream, fence, 0);
    }
    // a staging buffer of at least `bytes`, safe to overwrite from the copy stream
    cudaError_t stage_acquire(size_t bytes, void **dev)
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        Stage &s = stages[stage_next];
        stage_next = (stage_next + 1) % NSTAGE;
        cudaError_t e;
        if (!s.used && (e = cudaEventCreateWithFlags(&s.used, cudaEventDisableTiming)) != cudaSuccess) return e;
        if (s.bytes < bytes) {
            if (s.pending) cudaEventSynchronize(s.used);
            s.pending = false;
            if (s.p) cudaFree(s.p);
            s.p = nullptr;
            s.bytes = 0;
            const size_t want = bytes + bytes / 4;
            if (cudaMalloc(&s.p, want) != cudaSuccess) {
                cudaGetLastError();
                if ((e = cudaMalloc(&s.p, bytes)) != cudaSuccess) return e;
                s.bytes = bytes;
            } else {
                s.bytes = want;
            }
        }
        // same stream as the previous user of the buffer: stream order is enough
        *dev = s.p;
        return cudaSuccess;
    }
    void stage_release(void *dev)
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        for (auto &s : stages)
            if (s.p == dev) { cudaEventRecord(s.used, copy_stream); s.pending = true; return; }
    }
    void stage_release_all()
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        for (auto &s : stages) {
            if (s.p) cudaFree(s.p);
            if (s.used) cudaEventDestroy(s.used);
            s = Stage();
        }
        if (fence) cudaEventDestroy(fence);
        fence = nullptr;
    }

    cudaEvent_t get_event()
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void prof_begin(int cls, cudaEvent_t *a, cudaStream_t s)
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        launches++;
        if (!profile) return;
        *a = get_event();
        cudaEventRecord(*a, s);
        (void)cls;
    }
    void prof_end(int cls, cudaEvent_t a, cudaStream_t s)
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        if (!profile) return;
        cudaEvent_t b = get_event();
        cudaEventRecord(b, s);
        recs.push_back({cls, a, b});
    }
    void prof_resolve()
    {
        std::lock_guard<std::recursive_mutex> lock(mu);
        static const bool trace = getenv("NLF_TRACE") != nullptr;     // per-launch timeline on stderr (tuning)
        for (auto &r : recs) {
            float ms = 0;
            cudaEventSynchronize(r.b);
            cudaEventElapsedTime(&ms, r.a, r.b);
            if (trace && !recs.empty()) {
                float t0 = 0;
                cudaEventElapsedTime(&t0, recs.front().a, r.a);
                fprintf(stderr, "nlf-trace class %d start %.3f ms dur %.3f ms\n", r.cls, t0, ms);
            }
            prof_ms[r.cls] += ms;
            prof_n[r.cls] += 1;
            free_events.push_back(r.a);
            free_events.push_back(r.b);
        }
        recs.clear();
    }
};

// A sample's pixel table resident on the device (pixels.cu): cooler's CSR layout, bin1 <= bin2.
struct nlf_pixels {
    nlf_ctx *ctx = nullptr;
    long long nbins = 0, nnz = 0;
    long long *d_off = nullptr;      // [nbins+1]
    int *d_bin2 = nullptr;           // [nnz] ascending inside a row
    int *d_cnt = nullptr;            // [nnz]
    double *d_w = nullptr;           // [nbins] balancing weight per bin (NaN = masked bin)
};
// internal entry points of pixels.cu used by norm.cu / expected.cu / score.cu
int nlf_pixels_fill_assembly(nlf_pixels *px, int balance, const long long *src_lo, const long long *src_hi,
                             const unsigned char *rev, const int *out_lo, int nblk, int n, const int *d_blk,
                             double *d_M, double *d_bias);
int nlf_pixels_chrom_band(nlf_pixels *px, int balance, long long lo, long long hi, int nd, double *d_band,
                          unsigned char *d_valid);
int nlf_pixels_export_band(nlf_pixels *px, int balance, long long lo, long long hi, int bw, int pitch,
                           double **d_band_out, std::vector<int> &kept);

// RAII-free helper: launch bracket used as  { KLaunch k(ctx, cls); kernel<<<...>>>(...); }
struct KLaunch {
    nlf_ctx *c; int cls; cudaEvent_t a = nullptr; cudaStream_t s;
    static const char *range_name(int cls)
    {
        static const char *const names[nlf::P_NCLASS] = {"nlf:bandify", "nlf:diag_means+isotonic", "nlf:features", "nlf:forest",
                                                         "nlf:threshold", "nlf:normalise", "nlf:pool", "nlf:score", "nlf:isotonic", "nlf:count_candidates",
                                                         "nlf:other"};
        return names[(cls >= 0 && cls < nlf::P_NCLASS) ? cls : nlf::P_NCLASS - 1];
    }
    KLaunch(nlf_ctx *ctx, int cls_, cudaStream_t st = nullptr) : c(ctx), cls(cls_), s(st ? st : ctx->stream)
    {
        nvtxRangePushA(range_name(cls));          // NVTX range around every launch: `ncu --nvtx --nvtx-include "nlf:features/"`
        c->prof_begin(cls, &a, s);
    }
    ~KLaunch() { c->prof_end(cls, a, s); nvtxRangePop(); }
};

// ---------------------------------------------------------------------------------------------
// device helpers shared by the kernels: numpy's pairwise summation, exactly
// (numpy/_core/src/umath/loops_utils.h.src DOUBLE_pairwise_sum), no FMA, no reassociation.
// ---------------------------------------------------------------------------------------------
namespace nlf {

template <typename Load>
__device__ __forceinline__ double np_leaf_sum(Load ld, int off, int n)
{
    if (n < 8) {
        double r = -0.0;
        for (int i = 0; i < n; i++) r = __dadd_rn(r, ld(off + i));
        return r;
    }
    double r0 = ld(off + 0), r1 = ld(off + 1), r2 = ld(off + 2), r3 = ld(off + 3);
    double r4 = ld(off + 4), r5 = ld(off + 5), r6 = ld(off + 6), r7 = ld(off + 7);
    int i;
    int lim = n - (n % 8);
    for (i = 8; i < lim; i += 8) {
        double a0 = ld(off + i + 0), a1 = ld(off + i + 1), a2 = ld(off + i + 2), a3 = ld(off + i + 3);
        double a4 = ld(off + i + 4), a5 = ld(off + i + 5), a6 = ld(off + i + 6), a7 = ld(off + i + 7);
        r0 = __dadd_rn(r0, a0); r1 = __dadd_rn(r1, a1); r2 = __dadd_rn(r2, a2); r3 = __dadd_rn(r3, a3);
        r4 = __dadd_rn(r4, a4); r5 = __dadd_rn(r5, a5); r6 = __dadd_rn(r6, a6); r7 = __dadd_rn(r7, a7);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; i++) res = __dadd_rn(res, ld(off + i));
    return res;
}

// sum of ld(0..n-1) in numpy's pairwise order (recursion made explicit; depth <= 24)
template <typename Load>
__device__ double np_pairwise_sum(Load ld, int n)
{
    if (n <= 128) return np_leaf_sum(ld, 0, n);
    int st_off[48], st_n[48];   // st_n < 0 marks a "combine" entry
    double vs[26];
    int sp = 0, vp = 0;
    st_off[sp] = 0; st_n[sp] = n; sp++;
    while (sp) {
        sp--;
        int o = st_off[sp], m = st_n[sp];
        if (m < 0) {
            double r = vs[--vp];
            double l = vs[--vp];
            vs[vp++] = __dadd_rn(l, r);
        } else if (m <= 128) {
            vs[vp++] = np_leaf_sum(ld, o, m);
        } else {
            int n2 = m / 2;
            n2 -= n2 % 8;
            st_off[sp] = 0; st_n[sp] = -1; sp++;
            st_off[sp] = o + n2; st_n[sp] = m - n2; sp++;
            st_off[sp] = o; st_n[sp] = n2; sp++;
        }
    }
    return vs[0];
}

}  // namespace nlf
