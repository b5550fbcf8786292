// pool.cu -- pooling of the thresholded pixels into loop calls on B200 (sm_100a):
//   K10 Peakachu._local_clustering / find_anchors / _cluster_core   (callers.py:699-831)
// restating scipy.signal.find_peaks(height, distance) + peak_widths(rel_height=1, wlen),
// sklearn.cluster.dbscan(eps, min_samples=2) (= connected components of the eps-graph, isolated
// points are noise) and the reference's greedy centroid growth, all on integer coordinates so
// the result is exact.  One bucket (assembly x chain pair) per thread block; the control flow of
// the greedy growth is inherently sequential, so lane 0 drives and the O(m^2) scans are spread
// over the block's threads.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

using namespace nlf;

namespace {

struct Anchor { int summit, lb, rb; };

struct BucketWs {           // per-bucket workspace carved from one device allocation
    double *sig;            // [range]
    double *prio;           // [range]
    int *mid;               // [range]
    int *records;           // [range]
    unsigned char *keep;    // [range]
    unsigned char *alive;   // [range]
    int *order;             // [range]
    int *chain;             // [range]
    char *axis2;            // second copy of the seven arrays above: the y anchors are found concurrently
    size_t axis_bytes;
    Anchor *xa, *ya;        // [range] each
    // per point
    int *sl;                // [m] sorted list (point ids)
    int *label;             // [m]
    int *stack;             // [m]
    int *sub, *outl;        // [m]
    int *Lid;               // [m+1]
    unsigned char *pool;    // [m]
    unsigned char *visited; // [m]
    unsigned char *is_final;// [m]
    int *pr_rect, *pr_pt, *pr_sorted;   // [4m] (anchor rectangle, point) memberships
};

__device__ int d_local_maxima(const double *x, int n, int *mid)
{
    int m = 0, i = 1, i_max = n - 1;
    while (i < i_max) {
        if (x[i - 1] < x[i]) {
            int ahead = i + 1;
            while (ahead < i_max && x[ahead] == x[i]) ahead++;
            if (x[ahead] < x[i]) {
                mid[m++] = (i + ahead - 1) / 2;
                i = ahead;
            }
        }
        i++;
    }
    return m;
}

__device__ void d_peak_width_bounds(const double *x, int n, int peak, int wlen_in, double *left_ip, double *right_ip)
{
    int i_min = 0, i_max = n - 1;
    int wlen = wlen_in;
    if (wlen >= 2) {
        if (wlen % 2 == 0) wlen += 1;
        int half = wlen / 2;
        i_min = max(peak - half, 0);
        i_max = min(peak + half, n - 1);
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
    double prominence = __dsub_rn(x[peak], fmax(left_min, right_min));
    double height = __dsub_rn(x[peak], __dmul_rn(prominence, 1.0));
    i = peak;
    while (left_base < i && height < x[i]) i--;
    double lip = (double)i;
    if (x[i] < height) lip = __dadd_rn(lip, __ddiv_rn(__dsub_rn(height, x[i]), __dsub_rn(x[i + 1], x[i])));
    i = peak;
    while (i < right_base && height < x[i]) i++;
    double rip = (double)i;
    if (x[i] < height) rip = __dsub_rn(rip, __ddiv_rn(__dsub_rn(height, x[i]), __dsub_rn(x[i - 1], x[i])));
    *left_ip = lip;
    *right_ip = rip;
}

// Peakachu.find_anchors (callers.py:740-784); single thread.
// query != 0: only emit the height-filtered peak priorities (prio_out, *n_prio) and report in
// *ambiguous whether two equal heights sit closer than the distance filter.
__device__ int d_find_anchors(const int *pos, int npos, int res, int min_count, int r_bp, BucketWs &ws, Anchor *out,
                              const int *ext_order, int query, double *prio_out, int *n_prio, int *ambiguous,
                              int wlen_bp = 40000)
{
    int min_dis = r_bp / res; if (min_dis < 2) min_dis = 2;
    int wlen;
    if (res <= 10000) { wlen = wlen_bp / res; if (wlen > 20) wlen = 20; } else wlen = 4;
    int lo = pos[0], hi = pos[0];
    for (int i = 1; i < npos; i++) { lo = min(lo, pos[i]); hi = max(hi, pos[i]); }
    int n = hi - lo + 1;
    double *sig = ws.sig;
    for (int i = 0; i < n; i++) sig[i] = 0.0;
    for (int i = 0; i < npos; i++) sig[pos[i] - lo] += 1.0;
    int *mid = ws.mid;
    int m = d_local_maxima(sig, n, mid);
    int m2 = 0;
    for (int i = 0; i < m; i++) if (sig[mid[i]] >= (double)min_count) mid[m2++] = mid[i];
    m = m2;
    double *prio = ws.prio;
    for (int i = 0; i < m; i++) prio[i] = sig[mid[i]];
    if (query) {
        for (int i = 0; i < m; i++) prio_out[i] = prio[i];
        *n_prio = m;
        int amb = 0;
        for (int i = 0; i + 1 < m && !amb; i++)
            for (int k = i + 1; k < m && mid[k] - mid[i] < min_dis; k++)
                if (prio[k] == prio[i]) { amb = 1; break; }
        if (amb) *ambiguous = 1;
        return 0;
    }
    // Peak heights are histogram counts (small non-negative integers), so both orderings below are
    // counting sorts: O(m + max count) for a single thread even on chromosome-sized buckets.
    unsigned char *keep = ws.keep;
    int *order = ws.order;
    int *cnt = ws.records;                      // scratch until the interval painting starts
    int maxc = 0;
    for (int i = 0; i < m; i++) maxc = max(maxc, (int)prio[i]);
    if (min_dis > 2) {
        // _select_by_peak_distance: walk the peaks from the highest priority down; the order among
        // equal priorities is numpy.argsort's (ext_order when the host had to ask), else stable
        for (int i = 0; i < m; i++) keep[i] = 1;
        if (ext_order) {
            for (int i = 0; i < m; i++) order[i] = ext_order[i];
        } else {
            for (int c = 0; c <= maxc + 1; c++) cnt[c] = 0;
            for (int i = 0; i < m; i++) cnt[(int)prio[i] + 1]++;
            for (int c = 0; c < maxc; c++) cnt[c + 1] += cnt[c];
            for (int i = 0; i < m; i++) order[cnt[(int)prio[i]]++] = i;       // stable ascending
        }
        for (int i = m - 1; i >= 0; i--) {
            int j = order[i];
            if (!keep[j]) continue;
            int k = j - 1;
            while (k >= 0 && mid[j] - mid[k] < min_dis) { keep[k] = 0; k--; }
            k = j + 1;
            while (k < m && mid[k] - mid[j] < min_dis) { keep[k] = 0; k++; }
        }
        m2 = 0;
        for (int i = 0; i < m; i++) if (keep[i]) mid[m2++] = mid[i];
        m = m2;
    }   // two local maxima are never adjacent: a distance of 2 bins removes nothing
    // sorted_summits.sort(reverse=True) on (count, index): stable ascending by count, then reversed
    for (int c = 0; c <= maxc + 1; c++) cnt[c] = 0;
    for (int i = 0; i < m; i++) cnt[(int)sig[mid[i]] + 1]++;
    for (int c = 0; c < maxc; c++) cnt[c + 1] += cnt[c];
    for (int i = 0; i < m; i++) order[cnt[(int)sig[mid[i]]]++] = mid[i];
    for (int i = 0; i < m; i++) mid[i] = order[m - 1 - i];
    int *records = ws.records;
    unsigned char *alive = ws.alive;
    // anchors that share a summit are chained (head/next), so "is this exact interval already
    // there" looks at one short chain instead of every anchor
    int *head = ws.order, *nextq = ws.chain;
    for (int i = 0; i < n; i++) { records[i] = -1; head[i] = -1; }
    int na = 0;
    for (int s = 0; s < m; s++) {
        int pk = mid[s];
        double lip, rip;
        d_peak_width_bounds(sig, n, pk, wlen, &lip, &rip);
        int li = (int)rint(lip), ri = (int)rint(rip);     // np.round: half to even
        int summit = pk, m_lb = li, m_rb = ri;
        if (na > 0) {
            for (int b = li; b <= ri; b++) {
                if (records[b] >= 0) {
                    int q = records[b];
                    m_lb = min(li, out[q].lb);
                    m_rb = max(ri, out[q].rb);
                    summit = out[q].summit;
                    alive[q] = 0;
                    break;
                }
            }
        }
        int idx = -1, last = -1;
        for (int q = head[summit]; q >= 0; q = nextq[q]) {
            if (idx < 0 && alive[q] && out[q].lb == m_lb && out[q].rb == m_rb) idx = q;
            last = q;
        }
        if (idx < 0) {
            idx = na++;
            out[idx].summit = summit; out[idx].lb = m_lb; out[idx].rb = m_rb; alive[idx] = 1;
            nextq[idx] = -1;
            if (last >= 0) nextq[last] = idx; else head[summit] = idx;
        }
        for (int b = m_lb; b <= m_rb; b++) records[b] = idx;
    }
    int k = 0;
    for (int q = 0; q < na; q++)
        if (alive[q]) { out[k].summit = out[q].summit + lo; out[k].lb = out[q].lb + lo; out[k].rb = out[q].rb + lo; k++; }
    return k;
}

// (value, (i, j)) descending, as Python's tuple sort(reverse=True)
__device__ __forceinline__ bool pt_before(const int *pi, const int *pj, const double *val, int a, int b)
{
    if (val[a] != val[b]) return val[a] > val[b];
    if (pi[a] != pi[b]) return pi[a] > pi[b];
    return pj[a] > pj[b];
}

// Peakachu._cluster_core (callers.py:786-831) on the sorted id list sl[0..m)
__device__ void d_cluster_core(const int *sl, int m, int r, const int *pi, const int *pj, BucketWs &ws)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_flag, s_cnt, s_a;
    __shared__ unsigned long long s_b;
    __shared__ long long s_si, s_sj;
    if (m == 1) {
        if (tid == 0) ws.is_final[sl[0]] = 1;
        __syncthreads();
        return;
    }
    if (m < 2) return;
    const long long r2 = (long long)r * r;
    int *label = ws.label;
    // dbscan(eps=r, min_samples=2): connected components by min-label propagation
    for (int a = tid; a < m; a += nt) label[a] = a;
    __syncthreads();
    while (true) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        int changed = 0;
        for (int a = tid; a < m; a += nt) {
            int ia = pi[sl[a]], ja = pj[sl[a]], best = label[a];
            for (int b = 0; b < m; b++) {
                long long dx = ia - pi[sl[b]], dy = ja - pj[sl[b]];
                if (dx * dx + dy * dy <= r2) { int lb = label[b]; if (lb < best) best = lb; }
            }
            if (best < label[a]) { label[a] = best; changed = 1; }
        }
        if (changed) s_flag = 1;
        __syncthreads();
        // pointer jumping
        for (int a = tid; a < m; a += nt) { int l = label[a]; while (label[l] < l) l = label[l]; label[a] = l; }
        __syncthreads();
        if (!s_flag) break;
        __syncthreads();
    }
    // noise = component of size 1
    for (int a = tid; a < m; a += nt) ws.stack[a] = 0;
    __syncthreads();
    for (int a = tid; a < m; a += nt) atomicAdd(&ws.stack[label[a]], 1);
    __syncthreads();
    for (int a = tid; a < m; a += nt) { ws.pool[a] = 0; }
    __syncthreads();
    for (int a = tid; a < m; a += nt) if (ws.stack[label[a]] == 1) ws.outl[a] = -1; else ws.outl[a] = label[a];
    __syncthreads();
    for (int a = tid; a < m; a += nt) label[a] = ws.outl[a];
    __syncthreads();

    // greedy growth, seeds in sorted order.  in-set markers: sub[a] = 1 when a is in the
    // current `sub` list, Lid[a] = 1 when a is in Local.
    int *insub = ws.sub;
    int *inloc = ws.Lid;
    for (int p = 0; p < m; p++) {
        if (ws.pool[p]) continue;            // uniform: pool only changes between barriers
        int c = label[p];
        if (c == -1) continue;
        for (int a = tid; a < m; a += nt) { insub[a] = (label[a] == c); inloc[a] = 0; }
        __syncthreads();
        long long cen_i = pi[sl[p]], cen_j = pj[sl[p]], rad = r;
        long long nl = 1, sum_i = cen_i, sum_j = cen_j;      // Local = [p]
        if (tid == 0) inloc[p] = 1;
        int ini = -1;
        int nsub = 1;
        __syncthreads();
        while (true) {
            // size of sub
            if (tid == 0) { s_cnt = 0; s_a = 0; s_si = 0; s_sj = 0; }
            __syncthreads();
            int my_sub = 0;
            for (int a = tid; a < m; a += nt) my_sub += insub[a];
            if (my_sub) atomicAdd(&s_cnt, my_sub);
            __syncthreads();
            nsub = s_cnt;
            __syncthreads();
            if (nsub == 0) break;
            if (tid == 0) { s_cnt = 0; s_a = 0; }
            __syncthreads();
            int add = 0, nout = 0;
            long long asi = 0, asj = 0;
            for (int a = tid; a < m; a += nt) {
                if (!insub[a]) continue;
                if (ws.pool[a]) { insub[a] = 0; continue; }
                long long dx = pi[sl[a]] - cen_i, dy = pj[sl[a]] - cen_j;
                if (dx * dx + dy * dy <= rad * rad) {
                    add++; asi += pi[sl[a]]; asj += pj[sl[a]];
                    inloc[a] = 1; insub[a] = 0;
                } else nout++;
            }
            if (add) { atomicAdd(&s_a, add); atomicAdd((unsigned long long *)&s_si, (unsigned long long)asi);
                       atomicAdd((unsigned long long *)&s_sj, (unsigned long long)asj); }
            if (nout) atomicAdd(&s_cnt, nout);
            __syncthreads();
            int tot_out = s_cnt;
            nl += s_a; sum_i += s_si; sum_j += s_sj;
            __syncthreads();
            if (tot_out == ini) break;
            ini = tot_out;
            // centroid = round(mean(Local)), radius = round(max dist) + r
            cen_i = (long long)rint(__ddiv_rn((double)sum_i, (double)nl));
            cen_j = (long long)rint(__ddiv_rn((double)sum_j, (double)nl));
            if (tid == 0) s_b = 0;
            __syncthreads();
            long long mx2 = 0;
            for (int a = tid; a < m; a += nt)
                if (inloc[a]) {
                    long long dx = pi[sl[a]] - cen_i, dy = pj[sl[a]] - cen_j;
                    long long d2 = dx * dx + dy * dy;
                    if (d2 > mx2) mx2 = d2;
                }
            if (mx2) atomicMax(&s_b, (unsigned long long)mx2);
            __syncthreads();
            rad = (long long)rint(sqrt((double)s_b)) + r;
            __syncthreads();
        }
        for (int a = tid; a < m; a += nt) if (inloc[a]) ws.pool[a] = 1;
        if (tid == 0) ws.is_final[sl[p]] = 1;
        __syncthreads();
    }
    for (int a = tid; a < m; a += nt) if (ws.pool[a]) ws.visited[sl[a]] = 1;
    __syncthreads();
}

struct PoolArgs {
    const long long *bucket_off;
    const int *pi, *pj;
    const double *val;
    const int *x_order, *y_order;
    const long long *xo_off, *yo_off;
    unsigned char *keep_out;
    // workspace
    char *ws_base;
    const long long *ws_off;     // per bucket byte offset
    const int *range;            // per bucket max(range_x, range_y)
    int res, min_count, r_bp;
};

__device__ BucketWs carve(char *base, int range, int m)
{
    BucketWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *p = base + o; o += (bytes + 15) & ~(size_t)15; return p; };
    auto take_axis = [&](BucketWs &v) {
        v.sig = (double *)take(sizeof(double) * range);
        v.prio = (double *)take(sizeof(double) * range);
        v.mid = (int *)take(sizeof(int) * range);
        v.records = (int *)take(sizeof(int) * range);
        v.order = (int *)take(sizeof(int) * range);
        v.chain = (int *)take(sizeof(int) * range);
        v.keep = (unsigned char *)take(range);
        v.alive = (unsigned char *)take(range);
    };
    take_axis(w);
    w.axis2 = base + o;
    {
        BucketWs dummy;
        const size_t before = o;
        take_axis(dummy);
        w.axis_bytes = o - before;
    }
    w.xa = (Anchor *)take(sizeof(Anchor) * range);
    w.ya = (Anchor *)take(sizeof(Anchor) * range);
    w.sl = (int *)take(sizeof(int) * m);
    w.label = (int *)take(sizeof(int) * m);
    w.stack = (int *)take(sizeof(int) * m);
    w.sub = (int *)take(sizeof(int) * m);
    w.outl = (int *)take(sizeof(int) * m);
    w.Lid = (int *)take(sizeof(int) * (m + 1));
    w.pool = (unsigned char *)take(m);
    w.visited = (unsigned char *)take(m);
    w.is_final = (unsigned char *)take(m);
    w.pr_rect = (int *)take(sizeof(int) * 4 * (size_t)m);
    w.pr_pt = (int *)take(sizeof(int) * 4 * (size_t)m);
    w.pr_sorted = (int *)take(sizeof(int) * 4 * (size_t)m);
    return w;
}

__device__ BucketWs axis2_view(const BucketWs &w, int range)
{
    BucketWs v = w;
    size_t o = 0;
    auto take = [&](size_t bytes) { char *p = w.axis2 + o; o += (bytes + 15) & ~(size_t)15; return p; };
    v.sig = (double *)take(sizeof(double) * range);
    v.prio = (double *)take(sizeof(double) * range);
    v.mid = (int *)take(sizeof(int) * range);
    v.records = (int *)take(sizeof(int) * range);
    v.order = (int *)take(sizeof(int) * range);
    v.chain = (int *)take(sizeof(int) * range);
    v.keep = (unsigned char *)take(range);
    v.alive = (unsigned char *)take(range);
    return v;
}

}  // namespace

size_t pool_ws_bytes(int range, int m)
{
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t o = 0;
    o += 2 * (al(sizeof(double) * range) * 2 + al(sizeof(int) * range) * 4 + al(range) * 2);   // two axis scratch sets
    o += al(sizeof(Anchor) * range) * 2;
    o += al(sizeof(int) * m) * 5;
    o += al(sizeof(int) * (m + 1));
    o += al(m) * 3;
    o += al(sizeof(int) * 4 * (size_t)m) * 3;
    return o + 64;
}

// sorted id list of the points selected by `sel` (1 = take), descending (value, (i, j))
__device__ int blk_select_sorted(const unsigned char *sel, int m, const int *pi, const int *pj, const double *val,
                                 int *sl, int *scratch)
{
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // rank among the selected = number of selected points that sort before it
    for (int a = threadIdx.x; a < m; a += blockDim.x) {
        if (!sel[a]) continue;
        int rank = 0;
        for (int b = 0; b < m; b++)
            if (sel[b] && b != a) rank += pt_before(pi, pj, val, b, a);
        sl[rank] = a;
        atomicAdd(&s_n, 1);
    }
    __syncthreads();
    (void)scratch;
    return s_n;
}

__global__ void __launch_bounds__(256) k_pool(PoolArgs A)
{
    const int bk = blockIdx.x;
    const long long p0 = A.bucket_off[bk];
    const int m = (int)(A.bucket_off[bk + 1] - p0);
    if (m == 0) return;
    const int *pi = A.pi + p0, *pj = A.pj + p0;
    const double *val = A.val + p0;
    unsigned char *keep_out = A.keep_out + p0;
    BucketWs ws = carve(A.ws_base + A.ws_off[bk], A.range[bk], m);
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_nxa, s_nya;
    for (int a = tid; a < m; a += nt) { ws.visited[a] = 0; ws.is_final[a] = 0; }
    // the two anchor searches are sequential algorithms: one thread each, in different warps
    if (tid == 0) {
        s_nxa = d_find_anchors(pi, m, A.res, A.min_count, A.r_bp, ws, ws.xa, A.x_order ? A.x_order + A.xo_off[bk] : nullptr,
                               0, nullptr, nullptr, nullptr);
    } else if (tid == 32) {
        BucketWs wy = axis2_view(ws, A.range[bk]);
        s_nya = d_find_anchors(pj, m, A.res, A.min_count, A.r_bp, wy, ws.ya, A.y_order ? A.y_order + A.yo_off[bk] : nullptr,
                               0, nullptr, nullptr, nullptr);
    }
    __syncthreads();
    int r = A.r_bp / A.res; if (r < 2) r = 2;
    unsigned char *sel = (unsigned char *)ws.outl;   // outl is only used inside d_cluster_core, after sel is consumed
    const int nxa = s_nxa, nya = s_nya;
    // The reference walks every (x anchor, y anchor) rectangle and looks its points up
    // (callers.py:711-720).  Rectangles are independent of each other (fresh `pool` per call,
    // append-only results), so only the non-empty ones are visited here: every point emits the
    // rectangles it belongs to, the (rectangle, point) pairs are sorted by rectangle and, inside a
    // rectangle, in the reference's descending (value, (i, j)) order.
    __shared__ int s_npairs;
    const int cap = 4 * m;
    if (tid == 0) s_npairs = 0;
    __syncthreads();
    for (int k = tid; k < m; k += nt) {
        const int x = pi[k], y = pj[k];
        for (int a = 0; a < nxa; a++) {
            if (x < ws.xa[a].lb || x > ws.xa[a].rb) continue;
            for (int b = 0; b < nya; b++) {
                if (y < ws.ya[b].lb || y > ws.ya[b].rb) continue;
                int pos = atomicAdd(&s_npairs, 1);
                if (pos < cap) { ws.pr_rect[pos] = a * nya + b; ws.pr_pt[pos] = k; }
            }
        }
    }
    __syncthreads();
    const int npairs = s_npairs;
    if (npairs <= cap) {
        // rank sort of the pairs (all keys distinct)
        for (int u = tid; u < npairs; u += nt) {
            const int ru = ws.pr_rect[u], pu = ws.pr_pt[u];
            int rank = 0;
            for (int v = 0; v < npairs; v++) {
                const int rv = ws.pr_rect[v];
                rank += (rv < ru) || (rv == ru && v != u && pt_before(pi, pj, val, ws.pr_pt[v], pu));
            }
            ws.pr_sorted[rank] = u;
        }
        __syncthreads();
        int pos = 0;
        while (pos < npairs) {
            const int rect = ws.pr_rect[ws.pr_sorted[pos]];
            int end = pos + 1;
            while (end < npairs && ws.pr_rect[ws.pr_sorted[end]] == rect) end++;
            const int len = end - pos;
            for (int q = tid; q < len; q += nt) ws.sl[q] = ws.pr_pt[ws.pr_sorted[pos + q]];
            __syncthreads();
            d_cluster_core(ws.sl, len, r, pi, pj, ws);
            __syncthreads();
            pos = end;
        }
    } else {
        // heavily overlapping anchors: fall back to the plain double loop
        for (int a = 0; a < nxa; a++)
            for (int b = 0; b < nya; b++) {
                const Anchor X = ws.xa[a], Y = ws.ya[b];
                for (int k = tid; k < m; k += nt)
                    sel[k] = (pi[k] >= X.lb && pi[k] <= X.rb && pj[k] >= Y.lb && pj[k] <= Y.rb);
                __syncthreads();
                int cnt = blk_select_sorted(sel, m, pi, pj, val, ws.sl, nullptr);
                __syncthreads();
                d_cluster_core(ws.sl, cnt, r, pi, pj, ws);
                __syncthreads();
            }
    }
    for (int k = tid; k < m; k += nt) sel[k] = !ws.visited[k];
    __syncthreads();
    int cnt = blk_select_sorted(sel, m, pi, pj, val, ws.sl, nullptr);
    __syncthreads();
    d_cluster_core(ws.sl, cnt, r, pi, pj, ws);
    __syncthreads();
    for (int k = tid; k < m; k += nt) keep_out[k] = (ws.is_final[k] || !ws.visited[k]) ? 1 : 0;
}

// ---- host ---------------------------------------------------------------------------------------
struct PoolPlan {
    std::vector<long long> ws_off;
    std::vector<int> range;
    size_t total = 0;
};

static PoolPlan plan_pool(int nb, const int64_t *off, const int *pi, const int *pj)
{
    PoolPlan P;
    P.ws_off.resize(nb);
    P.range.resize(nb);
    for (int k = 0; k < nb; k++) {
        long long a = off[k], b = off[k + 1];
        int range = 1;
        if (b > a) {
            int lo_i = pi[a], hi_i = pi[a], lo_j = pj[a], hi_j = pj[a];
            for (long long q = a; q < b; q++) {
                lo_i = std::min(lo_i, pi[q]); hi_i = std::max(hi_i, pi[q]);
                lo_j = std::min(lo_j, pj[q]); hi_j = std::max(hi_j, pj[q]);
            }
            range = std::max(hi_i - lo_i, hi_j - lo_j) + 1;
        }
        range = std::max(range, (int)(b - a)) + 2;
        P.range[k] = range;
        P.ws_off[k] = (long long)P.total;
        P.total += pool_ws_bytes(range, (int)(b - a) + 1);
    }
    return P;
}

__global__ void k_pool_query(PoolArgs A, int n_buckets, const int *pos, double *prio_out, long long *n_prio,
                             int *ambiguous)
{
    int bk = blockIdx.x;
    if (bk >= n_buckets || threadIdx.x != 0) return;
    const long long p0 = A.bucket_off[bk];
    const int m = (int)(A.bucket_off[bk + 1] - p0);
    n_prio[bk] = 0;
    if (m == 0) return;
    BucketWs ws = carve(A.ws_base + A.ws_off[bk], A.range[bk], m);
    int np = 0;
    // priorities are written at the bucket's point offset (there are never more peaks than points)
    d_find_anchors(pos + p0, m, A.res, A.min_count, A.r_bp, ws, ws.xa, nullptr, 1, prio_out + p0, &np, ambiguous);
    n_prio[bk] = np;
}

static int upload_common(nlf_ctx *c, int nb, const int64_t *bucket_off, const PoolPlan &P, long long **d_off,
                         long long **d_wsoff, int **d_range, char **d_ws)
{
    cudaStream_t s = c->stream;
    NLF_CUDA(cudaMallocAsync(d_off, sizeof(long long) * (nb + 1), s));
    NLF_CUDA(cudaMallocAsync(d_wsoff, sizeof(long long) * nb, s));
    NLF_CUDA(cudaMallocAsync(d_range, sizeof(int) * nb, s));
    NLF_CUDA(cudaMallocAsync(d_ws, P.total + 256, s));
    NLF_CUDA(cudaMemcpyAsync(*d_off, bucket_off, sizeof(long long) * (nb + 1), cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(*d_wsoff, P.ws_off.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(*d_range, P.range.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s));
    return NLF_OK;
}

NLF_EXPORT int nlf_pool(nlf_ctx *ctx, int n_buckets, const int64_t *bucket_off, const int *pi, const int *pj,
                        const double *val, int res, int min_count, int r_bp, const int *x_order,
                        const int64_t *xo_off, const int *y_order, const int64_t *yo_off, uint8_t *keep_out)
{
    if (!ctx || n_buckets < 0 || !bucket_off) return fail(NLF_ERR_ARG, "nlf_pool: bad arguments");
    if (res <= 0) return fail(NLF_ERR_ARG, "nlf_pool: res must be positive");
    const long long np = bucket_off[n_buckets];
    if (n_buckets == 0 || np == 0) return NLF_OK;
    if (!pi || !pj || !val || !keep_out) return fail(NLF_ERR_ARG, "nlf_pool: NULL point arrays");
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    PoolPlan P = plan_pool(n_buckets, bucket_off, pi, pj);
    long long *d_off, *d_wsoff, *d_xo = nullptr, *d_yo = nullptr;
    int *d_range, *d_pi, *d_pj, *d_xord = nullptr, *d_yord = nullptr;
    double *d_val;
    char *d_ws;
    unsigned char *d_keep;
    int rc = upload_common(c, n_buckets, bucket_off, P, &d_off, &d_wsoff, &d_range, &d_ws);
    if (rc) return rc;
    NLF_CUDA(cudaMallocAsync(&d_pi, sizeof(int) * np, s));
    NLF_CUDA(cudaMallocAsync(&d_pj, sizeof(int) * np, s));
    NLF_CUDA(cudaMallocAsync(&d_val, sizeof(double) * np, s));
    NLF_CUDA(cudaMallocAsync(&d_keep, np, s));
    NLF_CUDA(cudaMemcpyAsync(d_pi, pi, sizeof(int) * np, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_pj, pj, sizeof(int) * np, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_val, val, sizeof(double) * np, cudaMemcpyHostToDevice, s));
    if (x_order && xo_off) {
        long long nx = xo_off[n_buckets];
        NLF_CUDA(cudaMallocAsync(&d_xord, sizeof(int) * std::max(1LL, nx), s));
        NLF_CUDA(cudaMallocAsync(&d_xo, sizeof(long long) * (n_buckets + 1), s));
        NLF_CUDA(cudaMemcpyAsync(d_xord, x_order, sizeof(int) * nx, cudaMemcpyHostToDevice, s));
        NLF_CUDA(cudaMemcpyAsync(d_xo, xo_off, sizeof(long long) * (n_buckets + 1), cudaMemcpyHostToDevice, s));
    }
    if (y_order && yo_off) {
        long long ny = yo_off[n_buckets];
        NLF_CUDA(cudaMallocAsync(&d_yord, sizeof(int) * std::max(1LL, ny), s));
        NLF_CUDA(cudaMallocAsync(&d_yo, sizeof(long long) * (n_buckets + 1), s));
        NLF_CUDA(cudaMemcpyAsync(d_yord, y_order, sizeof(int) * ny, cudaMemcpyHostToDevice, s));
        NLF_CUDA(cudaMemcpyAsync(d_yo, yo_off, sizeof(long long) * (n_buckets + 1), cudaMemcpyHostToDevice, s));
    }
    PoolArgs A;
    A.bucket_off = d_off; A.pi = d_pi; A.pj = d_pj; A.val = d_val;
    A.x_order = d_xord; A.y_order = d_yord; A.xo_off = d_xo; A.yo_off = d_yo;
    A.keep_out = d_keep; A.ws_base = d_ws; A.ws_off = d_wsoff; A.range = d_range;
    A.res = res; A.min_count = min_count; A.r_bp = r_bp;
    {
        KLaunch k(c, P_POOL);
        k_pool<<<n_buckets, 256, 0, s>>>(A);
    }
    NLF_CUDA(cudaGetLastError());
    NLF_CUDA(cudaMemcpyAsync(keep_out, d_keep, np, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(d_off, s); cudaFreeAsync(d_wsoff, s); cudaFreeAsync(d_range, s); cudaFreeAsync(d_ws, s);
    cudaFreeAsync(d_pi, s); cudaFreeAsync(d_pj, s); cudaFreeAsync(d_val, s); cudaFreeAsync(d_keep, s);
    if (d_xord) { cudaFreeAsync(d_xord, s); cudaFreeAsync(d_xo, s); }
    if (d_yord) { cudaFreeAsync(d_yord, s); cudaFreeAsync(d_yo, s); }
    return NLF_OK;
}

NLF_EXPORT int nlf_pool_peak_heights(nlf_ctx *ctx, int n_buckets, const int64_t *bucket_off, const int *pos, int res,
                                     int min_count, int r_bp, double *prio_out, int64_t *prio_off, int *ambiguous)
{
    if (!ctx || n_buckets < 0 || !bucket_off || !prio_off || !ambiguous) return fail(NLF_ERR_ARG, "nlf_pool_peak_heights: bad arguments");
    *ambiguous = 0;
    for (int k = 0; k <= n_buckets; k++) prio_off[k] = 0;
    const long long np = bucket_off[n_buckets];
    if (n_buckets == 0 || np == 0) return NLF_OK;
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    PoolPlan P = plan_pool(n_buckets, bucket_off, pos, pos);
    long long *d_off, *d_wsoff, *d_np;
    int *d_range, *d_pos, *d_amb;
    double *d_prio;
    char *d_ws;
    int rc = upload_common(c, n_buckets, bucket_off, P, &d_off, &d_wsoff, &d_range, &d_ws);
    if (rc) return rc;
    NLF_CUDA(cudaMallocAsync(&d_pos, sizeof(int) * np, s));
    NLF_CUDA(cudaMallocAsync(&d_prio, sizeof(double) * np, s));
    NLF_CUDA(cudaMallocAsync(&d_np, sizeof(long long) * n_buckets, s));
    NLF_CUDA(cudaMallocAsync(&d_amb, sizeof(int), s));
    NLF_CUDA(cudaMemsetAsync(d_amb, 0, sizeof(int), s));
    NLF_CUDA(cudaMemcpyAsync(d_pos, pos, sizeof(int) * np, cudaMemcpyHostToDevice, s));
    PoolArgs A;
    memset(&A, 0, sizeof(A));
    A.bucket_off = d_off; A.ws_base = d_ws; A.ws_off = d_wsoff; A.range = d_range;
    A.res = res; A.min_count = min_count; A.r_bp = r_bp;
    {
        KLaunch k(c, P_POOL);
        k_pool_query<<<n_buckets, 32, 0, s>>>(A, n_buckets, d_pos, d_prio, d_np, d_amb);
    }
    NLF_CUDA(cudaGetLastError());
    std::vector<double> prio((size_t)np);
    std::vector<long long> cnt(n_buckets);
    NLF_CUDA(cudaMemcpyAsync(prio.data(), d_prio, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(cnt.data(), d_np, sizeof(long long) * n_buckets, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(ambiguous, d_amb, sizeof(int), cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    long long o = 0;
    for (int k = 0; k < n_buckets; k++) {
        prio_off[k] = o;
        if (prio_out) memcpy(prio_out + o, prio.data() + bucket_off[k], sizeof(double) * cnt[k]);
        o += cnt[k];
    }
    prio_off[n_buckets] = o;
    cudaFreeAsync(d_off, s); cudaFreeAsync(d_wsoff, s); cudaFreeAsync(d_range, s); cudaFreeAsync(d_ws, s);
    cudaFreeAsync(d_pos, s); cudaFreeAsync(d_prio, s); cudaFreeAsync(d_np, s); cudaFreeAsync(d_amb, s);
    return NLF_OK;
}

// ---------------------------------------------------------------------------------------------
// The two helper methods of Peakachu as callable entry points (the pooling kernel uses the same
// device routines): find_anchors (callers.py:740-784) and _cluster_core (callers.py:786-831).
// ---------------------------------------------------------------------------------------------
__global__ void k_find_anchors_one(const int *pos, int npos, int res, int min_count, int r_bp, int wlen_bp, const int *order,
                                   char *ws_base, int range, int *out, int *n_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    BucketWs ws = carve(ws_base, range, npos);
    const int k = d_find_anchors(pos, npos, res, min_count, r_bp, ws, ws.xa, order, 0, nullptr, nullptr, nullptr, wlen_bp);
    for (int q = 0; q < k; q++) { out[3 * q] = ws.xa[q].summit; out[3 * q + 1] = ws.xa[q].lb; out[3 * q + 2] = ws.xa[q].rb; }
    *n_out = k;
}

__global__ void __launch_bounds__(256) k_cluster_core_one(const int *pi, const int *pj, int m, int r, char *ws_base,
                                                          unsigned char *is_final, unsigned char *pooled)
{
    BucketWs ws = carve(ws_base, 1, m);
    for (int a = threadIdx.x; a < m; a += blockDim.x) { ws.visited[a] = 0; ws.is_final[a] = 0; ws.sl[a] = a; }
    __syncthreads();
    d_cluster_core(ws.sl, m, r, pi, pj, ws);
    __syncthreads();
    for (int a = threadIdx.x; a < m; a += blockDim.x) { is_final[a] = ws.is_final[a]; pooled[a] = ws.visited[a]; }
}

NLF_EXPORT int nlf_pool_find_anchors(nlf_ctx *ctx, const int *pos, int npos, int res, int min_count, int min_dis_bp,
                                     const int *order, int *anchors_out, int cap, int *n_anchors)
{
    return nlf_pool_find_anchors_wlen(ctx, pos, npos, res, min_count, min_dis_bp, 40000, order, anchors_out, cap, n_anchors);
}

NLF_EXPORT int nlf_pool_find_anchors_wlen(nlf_ctx *ctx, const int *pos, int npos, int res, int min_count, int min_dis_bp,
                                          int wlen_bp, const int *order, int *anchors_out, int cap, int *n_anchors)
{
    if (!ctx || !pos || npos <= 0 || res <= 0 || !anchors_out || !n_anchors) return fail(NLF_ERR_ARG, "nlf_pool_find_anchors: bad arguments");
    if (res <= 10000 && std::min(wlen_bp / res, 20) <= 1)
        return fail(NLF_ERR_ARG, "`wlen` must be larger than 1, was %d", std::min(wlen_bp / res, 20));
    int lo = pos[0], hi = pos[0];
    for (int i = 1; i < npos; i++) { lo = std::min(lo, pos[i]); hi = std::max(hi, pos[i]); }
    const int range = std::max(hi - lo + 1, npos) + 2;
    if (cap < hi - lo + 1) return fail(NLF_ERR_ARG, "nlf_pool_find_anchors: anchors_out needs room for %d anchors", hi - lo + 1);
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    int *d_pos, *d_out, *d_n, *d_order = nullptr;
    char *d_ws;
    NLF_CUDA(cudaMallocAsync(&d_pos, sizeof(int) * npos, s));
    NLF_CUDA(cudaMallocAsync(&d_out, sizeof(int) * 3 * (size_t)range, s));
    NLF_CUDA(cudaMallocAsync(&d_n, sizeof(int), s));
    NLF_CUDA(cudaMallocAsync(&d_ws, pool_ws_bytes(range, npos + 1) + 256, s));
    NLF_CUDA(cudaMemcpyAsync(d_pos, pos, sizeof(int) * npos, cudaMemcpyHostToDevice, s));
    if (order) {
        NLF_CUDA(cudaMallocAsync(&d_order, sizeof(int) * range, s));
        NLF_CUDA(cudaMemcpyAsync(d_order, order, sizeof(int) * std::min(range, npos), cudaMemcpyHostToDevice, s));
    }
    {
        KLaunch k(c, P_POOL);
        k_find_anchors_one<<<1, 32, 0, s>>>(d_pos, npos, res, min_count, min_dis_bp, wlen_bp, d_order, d_ws, range, d_out, d_n);
    }
    NLF_CUDA(cudaGetLastError());
    int n = 0;
    NLF_CUDA(cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    if (n > 0) NLF_CUDA(cudaMemcpyAsync(anchors_out, d_out, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    *n_anchors = n;
    cudaFreeAsync(d_pos, s); cudaFreeAsync(d_out, s); cudaFreeAsync(d_n, s); cudaFreeAsync(d_ws, s);
    if (d_order) cudaFreeAsync(d_order, s);
    return NLF_OK;
}

NLF_EXPORT int nlf_pool_cluster_core(nlf_ctx *ctx, const int *pi, const int *pj, int m, int r_bins,
                                     uint8_t *is_final_out, uint8_t *pooled_out)
{
    if (!ctx || m < 0 || (m > 0 && (!pi || !pj || !is_final_out || !pooled_out)) || r_bins < 0)
        return fail(NLF_ERR_ARG, "nlf_pool_cluster_core: bad arguments");
    if (m == 0) return NLF_OK;
    nlf_ctx *c = ctx;
    NLF_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    int *d_pi, *d_pj;
    unsigned char *d_fin, *d_pool;
    char *d_ws;
    NLF_CUDA(cudaMallocAsync(&d_pi, sizeof(int) * m, s));
    NLF_CUDA(cudaMallocAsync(&d_pj, sizeof(int) * m, s));
    NLF_CUDA(cudaMallocAsync(&d_fin, m, s));
    NLF_CUDA(cudaMallocAsync(&d_pool, m, s));
    NLF_CUDA(cudaMallocAsync(&d_ws, pool_ws_bytes(1, m + 1) + 256, s));
    NLF_CUDA(cudaMemcpyAsync(d_pi, pi, sizeof(int) * m, cudaMemcpyHostToDevice, s));
    NLF_CUDA(cudaMemcpyAsync(d_pj, pj, sizeof(int) * m, cudaMemcpyHostToDevice, s));
    {
        KLaunch k(c, P_POOL);
        k_cluster_core_one<<<1, 256, 0, s>>>(d_pi, d_pj, m, r_bins, d_ws, d_fin, d_pool);
    }
    NLF_CUDA(cudaGetLastError());
    NLF_CUDA(cudaMemcpyAsync(is_final_out, d_fin, m, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaMemcpyAsync(pooled_out, d_pool, m, cudaMemcpyDeviceToHost, s));
    NLF_CUDA(cudaStreamSynchronize(s));
    cudaFreeAsync(d_pi, s); cudaFreeAsync(d_pj, s); cudaFreeAsync(d_fin, s); cudaFreeAsync(d_pool, s); cudaFreeAsync(d_ws, s);
    return NLF_OK;
}
