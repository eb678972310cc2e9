// Kernel group (2b): thresholding, ordered compaction of candidate pixels, connected
// components, deblending and flux-weighted centroids.
//
// Replaces (reference, libstacker/src/):
//   getAdjoiningPixels / detectAdjoiningPixel   stardetector.cpp:192-246
//   AdjoiningPixel::deblend / extract / isAdjoining / getGravityCenter  adjoiningpixel.cpp:41-118,172-202
//   AdjoiningPixel::createStar                  adjoiningpixel.cpp:120-150
// The reference's result depends on its traversal orders (raster order of components,
// DFS order of pixels, peel order of the deblender); SURVEY.md Appendix D is the
// behavioural spec.  Here:
//   - candidates (pixels > threshold) are compacted in raster order with per-segment
//     bit masks + an exclusive scan, so "position in the list" == raster rank;
//   - a union-find over the sparse list labels components by their raster-first pixel,
//     which reproduces the reference's component order;
//   - one WARP per component (k_components, <= 256 px) or one CTA per component (k_components_big, incl. the
//     >= 10000-px deblender bypass) replays the reference's DFS and peel loops on shared memory over precomputed
//     neighbour links (no image access, no coordinate searches): traversals whose ORDER is the result run on one
//     thread, everything else (peak search, adjacency stamps, rankings) in parallel.
#pragma once
#include "common.cuh"

namespace oss {

// ---- generic batched exclusive scan of int32 (3 phases) -------------------------------
// n may be device-resident (n_dev[f]) or a host constant (n_dev == nullptr).
#define SCAN_ITEMS 4096 // per block: 1024 threads x 4

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int *total)
{
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = wsum[lane], si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        wsum[lane] = si - s; // exclusive warp offsets
        if (lane == 31) *total = si;
    }
    __syncthreads();
    int r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(1024)
k_scan_phase1(const int *__restrict__ in, long long strideIn, int *__restrict__ out, long long strideOut,
              const int *__restrict__ n_dev, int n_dev_stride, int n_host, int *__restrict__ blocksum, long long strideBS)
{
    // grid.x is a small fixed number of CTAs per frame that stride over the 4096-item blocks actually present (n is usually
    // a device-side count far below the capacity the grid used to be sized for: thousands of empty 1024-thread CTAs at high
    // stream priority were displacing the bulk kernels' CTAs)
    const int f = blockIdx.y;
    const int n = n_dev ? n_dev[(long long)f * n_dev_stride] : n_host;
    const int *src = in + f * strideIn;
    int *dst = out + f * strideOut;
    __shared__ int total;
    for (int blk = blockIdx.x; (long long)blk * SCAN_ITEMS < n; blk += gridDim.x) {
        const long long base = (long long)blk * SCAN_ITEMS;
        int v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            long long i = base + threadIdx.x * 4 + k;
            v[k] = i < n ? src[i] : 0;
            s += v[k];
        }
        int ex = block_exclusive_scan_1024(s, &total);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            long long i = base + threadIdx.x * 4 + k;
            if (i < n) dst[i] = ex;
            ex += v[k];
        }
        if (threadIdx.x == 0) blocksum[f * strideBS + blk] = total;
        __syncthreads(); // `total` is rewritten by the next block
    }
}

// one block per frame: exclusive scan of the block sums, total to out[n] and *total_out
__global__ void __launch_bounds__(1024)
k_scan_phase2(int *__restrict__ blocksum, long long strideBS, const int *__restrict__ n_dev, int n_dev_stride,
              int n_host, int *__restrict__ out, long long strideOut, int write_total_at_n, int *__restrict__ total_out,
              int total_stride)
{
    const int f = blockIdx.x;
    const int n = n_dev ? n_dev[(long long)f * n_dev_stride] : n_host;
    const int nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    int *bs = blocksum + f * strideBS;
    __shared__ int total;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? bs[i] : 0;
        int ex = block_exclusive_scan_1024(v, &total);
        if (i < nb) bs[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (write_total_at_n) out[f * strideOut + n] = carry;
        if (total_out) total_out[(long long)f * total_stride] = carry;
    }
}

__global__ void __launch_bounds__(1024)
k_scan_phase3(int *__restrict__ out, long long strideOut, const int *__restrict__ n_dev, int n_dev_stride, int n_host,
              const int *__restrict__ blocksum, long long strideBS)
{
    const int f = blockIdx.y;
    const int n = n_dev ? n_dev[(long long)f * n_dev_stride] : n_host;
    int *dst = out + f * strideOut;
    for (int blk = blockIdx.x; (long long)blk * SCAN_ITEMS < n; blk += gridDim.x) {
        if (blk == 0) continue;
        const long long base = (long long)blk * SCAN_ITEMS;
        const int add = blocksum[f * strideBS + blk];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            long long i = base + threadIdx.x * 4 + k;
            if (i < n) dst[i] += add;
        }
    }
}

// ---- K4a  per-segment candidate masks ---------------------------------------------------
// A segment is 32 consecutive pixels of one row.  Segments inside a 32 px x 8 rows cell whose maximum (from K3) does not
// exceed the threshold are skipped without touching the image: at t=20 that is > 99 %.
__global__ void __launch_bounds__(256)
k_segment_masks(const float *__restrict__ stars, long long strideStars, const float *__restrict__ segmax,
                long long strideSeg, uint32_t *__restrict__ segmask, int *__restrict__ segcount,
                const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    const long long nsegs = (long long)g.H * g.nseg;
    const long long s = (long long)blockIdx.x * 256 + threadIdx.x;
    if (s >= nsegs) return;
    const float th = ds[f].threshold;
    const int y = (int)(s / g.nseg), sx = (int)(s - (long long)y * g.nseg);
    uint32_t m = 0;
    if (segmax[f * strideSeg + (long long)(y / OSS_SEGROWS) * g.nseg + sx] > th) {
        const float *row = stars + f * strideStars + (long long)y * g.W + sx * OSS_SEG;
        const int n = min(OSS_SEG, g.W - sx * OSS_SEG);
        for (int b = 0; b < n; b++)
            if (row[b] > th) m |= 1u << b;
    }
    segmask[f * strideSeg + s] = m;
    segcount[f * (strideSeg + 1) + s] = __popc(m);
}

// ---- K4c  ordered compaction -----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_compact_candidates(const float *__restrict__ stars, long long strideStars, const uint32_t *__restrict__ segmask,
                     long long strideSeg, const int *__restrict__ segoff, int *__restrict__ cand_idx,
                     float *__restrict__ cand_val, int *__restrict__ owner, DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    const long long nsegs = (long long)g.H * g.nseg;
    const long long s = (long long)blockIdx.x * 256 + threadIdx.x;
    if (s >= nsegs) return;
    if (s == 0) {
        int total = segoff[f * (strideSeg + 1) + nsegs];
        ds[f].ncand = total;
        if (total > g.cap) ds[f].overflow = 1;
    }
    uint32_t m = segmask[f * strideSeg + s];
    if (!m) return;
    int o = segoff[f * (strideSeg + 1) + s];
    const int y = (int)(s / g.nseg), sx = (int)(s - (long long)y * g.nseg);
    const int lin0 = y * g.W + sx * OSS_SEG;
    const float *row = stars + f * strideStars + lin0;
    while (m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        if (o < g.cap) {
            long long q = (long long)f * g.cap + o;
            cand_idx[q] = lin0 + b;
            cand_val[q] = row[b];
            owner[q] = -3; // not yet visited by the component walk
        }
        o++;
    }
}

// position of pixel (x, y) in the candidate list, or -1
__device__ __forceinline__ int cand_lookup(const uint32_t *segmask, const int *segoff, int nseg, int x, int y)
{
    long long s = (long long)y * nseg + (x >> 5);
    uint32_t m = segmask[s];
    int b = x & 31;
    if (!((m >> b) & 1u)) return -1;
    return segoff[s] + __popc(m & ((1u << b) - 1u));
}

// ---- K5a  neighbour links + union-find init -------------------------------------------
// nbr = (up, left, right, down) positions, the reference's push order (stardetector.cpp:231-242)
__global__ void __launch_bounds__(256)
k_links(const int *__restrict__ cand_idx, const uint32_t *__restrict__ segmask, long long strideSeg,
        const int *__restrict__ segoff, int4 *__restrict__ nbr, int *__restrict__ parent,
        int *__restrict__ comp_size, int *__restrict__ comp_peak, const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int n = ds[f].ncand;
    const uint32_t *mk = segmask + f * strideSeg;
    const int *so = segoff + f * (strideSeg + 1);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        long long q = (long long)f * g.cap + i;
        int lin = cand_idx[q];
        int y = lin / g.W, x = lin - y * g.W;
        int4 nb;
        nb.x = y > 0 ? cand_lookup(mk, so, g.nseg, x, y - 1) : -1;
        nb.y = x > 0 ? cand_lookup(mk, so, g.nseg, x - 1, y) : -1;
        nb.z = x + 1 < g.W ? cand_lookup(mk, so, g.nseg, x + 1, y) : -1;
        nb.w = y + 1 < g.H ? cand_lookup(mk, so, g.nseg, x, y + 1) : -1;
        nbr[q] = nb;
        parent[q] = i;
        comp_size[q] = 0;
        comp_peak[q] = 0;
    }
}

__device__ __forceinline__ int uf_find(int *parent, int i)
{
    int p = parent[i];
    while (p != i) { i = p; p = parent[i]; }
    return i;
}

__device__ __forceinline__ void uf_union(int *parent, int a, int b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; } // a > b: hook the larger root under the smaller
        int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// ---- K5b  union with the up and left neighbours (4-connectivity) ------------------------
__global__ void __launch_bounds__(256)
k_ccl_union(const int4 *__restrict__ nbr, int *__restrict__ parent, const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int n = ds[f].ncand;
    int *par = parent + (long long)f * g.cap;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        int4 nb = nbr[(long long)f * g.cap + i];
        if (nb.x >= 0) uf_union(par, i, nb.x);
        if (nb.y >= 0) uf_union(par, i, nb.y);
    }
}

// ---- K5c  flatten labels, component size and peak (label = raster-first pixel) ----------
__global__ void __launch_bounds__(256)
k_ccl_flatten(int *__restrict__ parent, const float *__restrict__ cand_val, int *__restrict__ comp_size,
              int *__restrict__ comp_peak, const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int n = ds[f].ncand;
    int *par = parent + (long long)f * g.cap;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        int r = uf_find(par, i);
        par[i] = r; // benign race: every writer stores a valid ancestor, roots never change here
        atomicAdd(&comp_size[(long long)f * g.cap + r], 1);
        // candidates are > threshold >= 0, so their bit patterns order like ints
        atomicMax(&comp_peak[(long long)f * g.cap + r], __float_as_int(cand_val[(long long)f * g.cap + i]));
    }
}

// flag[i] = 1 for roots whose component passes peak > minPeak (stardetector.cpp:201);
// flag2[i] = component pixel count for those roots (scanned into scratch offsets)
__global__ void __launch_bounds__(256)
k_root_flags(const int *__restrict__ parent, const int *__restrict__ comp_size, const int *__restrict__ comp_peak,
             int *__restrict__ flag, int *__restrict__ flag2, const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int n = ds[f].ncand;
    const float minpeak = ds[f].minpeak;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        long long q = (long long)f * g.cap + i;
        int keep = parent[q] == i && __int_as_float(comp_peak[q]) > minpeak;
        flag[q] = keep;
        flag2[q] = keep ? comp_size[q] : 0;
    }
}

// comp_root[c] = list position of component c's first pixel, comp_off[c] = scratch offset; components above CW_CAP_SMALL pixels
// are also listed by size class for the later k_components / k_components_big instances (scratch region [11 cap, 12 cap):
// nobody else's; counters in DetectState, zeroed by k_finalize_stats)
#define CBIG_SMALL 2048   // first size class of k_components_big
#define CW_CAP_SMALL 64
#define CW_CAP 256
__device__ __forceinline__ int *class_list(int *scratch, long long fb, const Geom &g, int cls) // 0: 65..256 px, 1: ..2048, 2: larger
{
    return scratch + 12 * fb + 11ll * g.cap + (cls == 0 ? 0 : (cls == 1 ? g.cap / 2 : g.cap / 2 + g.cap / 4));
}
__global__ void __launch_bounds__(256)
k_gather_roots(const int *__restrict__ flag, const int *__restrict__ flagscan, const int *__restrict__ sizescan,
               const int *__restrict__ comp_size, int *__restrict__ comp_root, int *__restrict__ comp_off, int *__restrict__ fill,
               int *__restrict__ scratch, DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int n = ds[f].ncand;
    const long long fb = (long long)f * g.cap;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        long long q = fb + i;
        if (flag[q]) {
            const int ci = flagscan[q];
            long long c = fb + ci;
            comp_root[c] = i;
            comp_off[c] = sizescan[q];
            fill[c] = 0; // member counter of k_comp_members (the array is comp_nstar: k_components overwrites it afterwards)
            const int sz = comp_size[q];
            if (sz > CW_CAP_SMALL) {
                const int cls = sz <= CW_CAP ? 0 : (sz <= CBIG_SMALL ? 1 : 2);
                int *cnt = cls == 0 ? &ds[f].nbig : (cls == 1 ? &ds[f].nbig2 : &ds[f].nbig3);
                class_list(scratch, fb, g, cls)[atomicAdd(cnt, 1)] = ci;
            }
        }
    }
}

// ---- K6  per-component replay of the reference's sequential logic -----------------------

// ---- K6  one WARP per component ----------------------------------------------------------
// The component's pixels (k_comp_members: an unordered scatter by label) are staged in shared memory, sorted back into raster
// order, and their neighbour links are translated to local indices; every traversal of the reference then runs on shared
// memory instead of chasing pointers through L2:
//   sequential by lane 0 (their ORDER is the result): the DFS of detectAdjoiningPixel, the deblender's extract();
//   sequential chains spread over lanes 0..3 (one fp chain each, same operation order): gravity centre, createStar's sums;
//   parallel over the warp: first-maximum search (getPeak), isAdjoining stamps, blending count, the (value, extraction
//   order) ranking that replaces createStar's std::sort.
// One thread per component over global memory (the first version) made the largest component of a batch its critical
// path: ~2 us per pixel, 1.3 ms per 17-frame batch; here it is ~0.1 us per pixel.
// Two instances: CAP = 64 (3.2 KB per warp, 8 warps per CTA: ~48 resident warps per SM — the kernel is latency-bound on lane 0's
// traversals, so residency is throughput; it walks over ALL components and skips the larger ones) and CAP = 256 (12.5 KB per
// warp, 4 warps per CTA) over k_gather_roots' list of the 65..256-px components; the two (and k_components_big) are independent
// and run side by side on two streams.
#define CW_WARPS_SMALL 8
#define CW_WARPS 4
template <int CAP> struct CompWarpSmem {
    int gidx[CAP];            // candidate-list positions, ascending = raster order
    float val[CAP];
    int px[CAP], py[CAP];
    short nb[CAP][4];         // local neighbours: up, left, right, down (-1: none)
    short own[CAP];           // -3 unvisited, -1 remaining, -2 dropped, >= 0 accepted region
    short list[CAP];          // the reference's DFS order
    short S[CAP];             // region being peeled, extraction order
    short sorted[CAP];        // positions in S, ascending by (value, extraction order)
    short stamp[CAP];
    int acc_gx[CAP], acc_gy[CAP];
    short stack[4 * CAP + 4];
};
constexpr int kCompSmemBytes = (int)sizeof(CompWarpSmem<CW_CAP>) * CW_WARPS;
constexpr int kCompSmemBytesSmall = (int)sizeof(CompWarpSmem<CW_CAP_SMALL>) * CW_WARPS_SMALL;

// createStar (adjoiningpixel.cpp:120-150) over region S[0..ns) by one warp.  Lanes 0..3 carry one chain each, all with the
// formula acc = float(double(acc) + coord * double(v)): lanes 0 / 1 the flux-weighted x / y sums (coord = x + 0.5, y + 0.5, in
// ascending-value order), lane 2 the weight and lane 3 `value` (coord = 1: double(acc) + double(v) rounded once to fp32 IS the
// fp32 sum — 53 >= 2 * 24 + 2 bits, double rounding is innocuous), lane 3 in extraction order like the reference's first loop.
template <int CAP>
__device__ oss_star create_star_warp(CompWarpSmem<CAP> &sm, int ns, int lane)
{
    float *sv = (float *)sm.stack; // the region's values in extraction order (the DFS stack is idle here)
    for (int q = lane; q < ns; q += 32) sv[q] = sm.val[sm.S[q]];
    __syncwarp();
    float peak = -INFINITY;
    for (int q = lane; q < ns; q += 32) {
        const float v = sv[q];
        peak = fmaxf(peak, v);
        // rank = number of elements ordered before this one: ascending value, ties by extraction order
        int r = 0;
        for (int p2 = 0; p2 < ns; p2++) {
            const float u = sv[p2];
            r += (u < v) || (u == v && p2 < q);
        }
        sm.sorted[r] = (short)q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    __syncwarp();
    float acc = 0.f;
    if (lane < 4) {
        for (int q = 0; q < ns; q++) {
            const int i = sm.S[lane == 3 ? q : sm.sorted[q]];
            const double coord = lane == 0 ? (double)sm.px[i] + 0.5 : (lane == 1 ? (double)sm.py[i] + 0.5 : 1.0);
            acc = (float)__dadd_rn((double)acc, __dmul_rn(coord, (double)sm.val[i]));
        }
    }
    const float xa = __shfl_sync(0xffffffffu, acc, 0), ya = __shfl_sync(0xffffffffu, acc, 1);
    const float wt = __shfl_sync(0xffffffffu, acc, 2), value = __shfl_sync(0xffffffffu, acc, 3);
    oss_star st;
    st.area = ns;
    st.peak = peak;
    st.value = value;
    st.x = (int)__fdiv_rn(xa, wt);
    st.y = (int)__fdiv_rn(ya, wt);
    return st;
}

template <int CAP, int WARPS, bool FROM_LIST>
__global__ void __launch_bounds__(32 * WARPS)
k_components(const int *__restrict__ cand_idx, const float *__restrict__ cand_val, const int4 *__restrict__ nbr,
             int *__restrict__ owner, const int *__restrict__ comp_size, const int *__restrict__ comp_root,
             const int *__restrict__ comp_off, int *__restrict__ comp_nstar, int *__restrict__ scratch,
             oss_star *__restrict__ star_tmp, const DetectState *__restrict__ ds, Geom g)
{
    extern __shared__ __align__(16) unsigned char comp_smem_raw[];
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    CompWarpSmem<CAP> &sm = ((CompWarpSmem<CAP> *)comp_smem_raw)[wid];
    const long long fb = (long long)f * g.cap;
    const int *cidx = cand_idx + fb;
    const float *cval = cand_val + fb;
    const int4 *nb = nbr + fb;
    int *scr = scratch + 12 * fb;
    const int *mid_list = class_list(scratch, fb, g, 0);
    const float step = ds[f].threshold;
    const int nitems = FROM_LIST ? ds[f].nbig : ds[f].ncomp;

    for (int e = blockIdx.x * WARPS + wid; e < nitems; e += gridDim.x * WARPS) {
        const int c = FROM_LIST ? mid_list[e] : e;
        const int root = comp_root[fb + c];
        const int n = comp_size[fb + root];
        const int off = comp_off[fb + c];
        oss_star *out = star_tmp + fb + off;
        if (n > CAP) continue; // another instance's (k_gather_roots' lists)
        // ---- stage: members (any order) -> ascending candidate positions (rank sort; positions are distinct)
        const int *member = scr + off;
        for (int k = lane; k < n; k += 32) sm.px[k] = member[k];
        __syncwarp();
        for (int k = lane; k < n; k += 32) {
            const int v = sm.px[k];
            int r = 0;
            for (int j = 0; j < n; j++) r += sm.px[j] < v;
            sm.gidx[r] = v;
        }
        __syncwarp();
        for (int k = lane; k < n; k += 32) {
            const int i = sm.gidx[k];
            sm.val[k] = cval[i];
            const int lin = cidx[i];
            const int y = lin / g.W;
            sm.py[k] = y;
            sm.px[k] = lin - y * g.W;
            sm.own[k] = -3;
            const int4 q = nb[i];
            const int gl[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int d = 0; d < 4; d++) {
                int loc = -1;
                if (gl[d] >= 0) { // neighbours above the threshold belong to the same component: binary search
                    int lo = 0, hi = n - 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.gidx[mid] < gl[d]) lo = mid + 1; else hi = mid; }
                    loc = lo;
                }
                sm.nb[k][d] = (short)loc;
            }
        }
        __syncwarp();
        // ---- detectAdjoiningPixel: explicit-stack DFS from the raster-first pixel, push order up, left, right, down
        if (lane == 0) {
            int np = 0, sp = 0;
            sm.stack[sp++] = 0;
            while (sp > 0) {
                const int i = sm.stack[--sp];
                if (sm.own[i] != -3) continue;
                sm.own[i] = -1; // REMAINING (for the deblender)
                sm.list[np++] = (short)i;
                // the four links in one load, the four neighbour states in flight together (this loop is a chain of shared-memory
                // latencies on one lane: with a load pair per neighbour behind a branch each it was 10 deep, now 4)
                const short4 q = *reinterpret_cast<const short4 *>(sm.nb[i]);
                const short o0 = sm.own[max((int)q.x, 0)], o1 = sm.own[max((int)q.y, 0)], o2 = sm.own[max((int)q.z, 0)], o3 = sm.own[max((int)q.w, 0)];
                if (q.x >= 0 && o0 == -3) sm.stack[sp++] = q.x;
                if (q.y >= 0 && o1 == -3) sm.stack[sp++] = q.y;
                if (q.z >= 0 && o2 == -3) sm.stack[sp++] = q.z;
                if (q.w >= 0 && o3 == -3) sm.stack[sp++] = q.w;
            }
        }
        __syncwarp();
        // ---- AdjoiningPixel::deblend (adjoiningpixel.cpp:41-102); n < OSS_DEBLEND_MAX here
        int left = n, nacc = 0, iter = 0, nout = 0;
        while (left > 0) {
            iter++;
            // getPeak: first maximum in list order among the remaining pixels
            float mx = 0.f;
            int pos = 0x7fffffff;
            for (int q = lane; q < n; q += 32) {
                const int i = sm.list[q];
                if (sm.own[i] != -1) continue;
                const float v = sm.val[i];
                if (pos == 0x7fffffff || mx < v) { mx = v; pos = q; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float v2 = __shfl_xor_sync(0xffffffffu, mx, o);
                const int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
                // the sequential scan keeps the earlier element unless a later one is strictly larger
                const bool take = p2 != 0x7fffffff && (pos == 0x7fffffff || (p2 < pos ? !(v2 < mx) : mx < v2));
                if (take) { mx = v2; pos = p2; }
            }
            if (pos == 0x7fffffff) break;
            const int pk = sm.list[pos];
            const float T = (float)__dsub_rn((double)mx, __dmul_rn((double)step, sqrt(__dadd_rn((double)__fdiv_rn(mx, step), 1.0)))); // adjoiningpixel.cpp:57
            // extract(): recursive pre-order up, left, right, down == stack with reversed pushes
            int ns = 0;
            if (lane == 0) {
                int sp = 0;
                sm.stack[sp++] = (short)pk;
                while (sp > 0) {
                    const int i = sm.stack[--sp];
                    const short oi = sm.own[i];
                    const float vi = sm.val[i];
                    if (oi != -1 || !(vi > T)) continue;
                    sm.own[i] = -2; // DROPPED unless accepted below
                    sm.S[ns++] = (short)i;
                    const short4 q = *reinterpret_cast<const short4 *>(sm.nb[i]);
                    const short o0 = sm.own[max((int)q.x, 0)], o1 = sm.own[max((int)q.y, 0)], o2 = sm.own[max((int)q.z, 0)], o3 = sm.own[max((int)q.w, 0)];
                    if (q.w >= 0 && o3 == -1) sm.stack[sp++] = q.w;
                    if (q.z >= 0 && o2 == -1) sm.stack[sp++] = q.z;
                    if (q.y >= 0 && o1 == -1) sm.stack[sp++] = q.y;
                    if (q.x >= 0 && o0 == -1) sm.stack[sp++] = q.x;
                }
            }
            ns = __shfl_sync(0xffffffffu, ns, 0);
            if (ns == 0) break; // step <= 0 or NaN: the reference would never terminate
            left -= ns;
            // getGravityCenter (adjoiningpixel.cpp:104-118): fp32, list order; lanes 0..2 carry gx, gy, gw (v * 1.f == v)
            float gacc = 0.f;
            if (lane < 3) {
                for (int q = 0; q < ns; q++) {
                    const int i = sm.S[q];
                    const float coord = lane == 0 ? (float)sm.px[i] : (lane == 1 ? (float)sm.py[i] : 1.f);
                    gacc = __fadd_rn(gacc, __fmul_rn(sm.val[i], coord));
                }
            }
            const float gx = __shfl_sync(0xffffffffu, gacc, 0), gy = __shfl_sync(0xffffffffu, gacc, 1), gw = __shfl_sync(0xffffffffu, gacc, 2);
            const int cx = (int)__fdiv_rn(gx, gw), cy = (int)__fdiv_rn(gy, gw);
            // isAdjoining against every accepted region
            for (int q = lane; q < ns; q += 32) {
                const int i = sm.S[q];
#pragma unroll
                for (int d = 0; d < 4; d++) {
                    const int t = sm.nb[i][d];
                    if (t >= 0) { const int o2 = sm.own[t]; if (o2 >= 0) sm.stamp[o2] = (short)iter; }
                }
            }
            __syncwarp();
            int blending = 0;
            for (int a = lane; a < nacc; a += 32) blending += (sm.stamp[a] == iter) || (sm.acc_gx[a] == cx && sm.acc_gy[a] == cy);
            blending = __any_sync(0xffffffffu, blending != 0);
            if (!blending) {
                for (int q = lane; q < ns; q += 32) sm.own[sm.S[q]] = (short)nacc;
                if (lane == 0) { sm.acc_gx[nacc] = cx; sm.acc_gy[nacc] = cy; sm.stamp[nacc] = 0; }
                nacc++;
                __syncwarp();
                const oss_star st = create_star_warp<CAP>(sm, ns, lane);
                if (lane == 0) out[nout] = st;
                nout++;
            }
            __syncwarp();
        }
        if (lane == 0) comp_nstar[fb + c] = nout;
        __syncwarp();
    }
}

// ---- K6b  one CTA per LARGE component (n > CW_CAP) ----------------------------------------------------------
// Same algorithm and orders as k_components, for the components its per-warp scratch cannot hold (a saturated star with its
// halo, a galaxy core: hundreds to thousands of pixels).  One thread replaying such a component over global memory costs
// ~1 us per pixel and peel iteration (37 ms for one 1400-px blob, tools/big_blob_probe.py); here
//   in shared memory, sequential on thread 0 (their ORDER is the result): the DFS of detectAdjoiningPixel and the
//     deblender's extract() — value, neighbour links (up / down as local indices, left / right are k -+ 1 in raster order),
//     a state byte and the DFS stack (4 n + 4 entries: the reference's recursion pushes duplicates);
//   parallel over the CTA: staging (bitonic sort of the member list back to raster order, links by binary search), getPeak
//     (maximum, ties to the earlier DFS position), isAdjoining stamps, the blending test, createStar's (value, extraction
//     order) ranking;
//   fp chains on lanes 0..3 of warp 0 in the reference's operation order, their operands gathered 32 at a time by the whole
//     warp (coordinates live in global scratch) and handed over by shuffles, the next 32 in flight meanwhile.
// CB = capacity in pixels: two instances, 2048 (43 KB of shared memory: several CTAs per SM) and OSS_DEBLEND_MAX - 1
// (210 KB: one CTA per SM, like k_focas); a CTA skips the list entries of the other instance.
// The deblender bypass (a component of >= 10000 pixels: a satellite trail, a lattice of diffraction spikes) by one CTA over
// global memory: the reference's DFS (its order decides `value`'s summation order and the ties of the sort) on thread 0, then
// createStar with the (value, extraction order) sort as a bitonic network over 64-bit keys in the idle DFS stack and the fp
// chains on lanes 0..3 like everywhere else.  (One thread doing all of it, heap sort included, took ~40 us per pixel.)
__device__ void big_bypass(int c, int root, int n, int off, long long fb, const int *__restrict__ cand_idx, const float *__restrict__ cand_val,
                           const int4 *__restrict__ nbr, int *__restrict__ owner, int *__restrict__ scr, oss_star *__restrict__ star_tmp,
                           int *__restrict__ comp_nstar, float *red, const Geom &g)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = blockDim.x;
    const int *cidx = cand_idx + fb;
    const float *cval = cand_val + fb;
    const int4 *nb = nbr + fb;
    int *own = owner + fb;
    int *list = scr + off;
    const long long stack_at = 6ll * g.cap + 4ll * off + c; // 4 n + 1 entries
    int *stack = scr + stack_at;
    __syncthreads();
    if (tid == 0) { // detectAdjoiningPixel: explicit-stack DFS, push order up, left, right, down
        int np = 0, sp = 0;
        stack[sp++] = root; // the raster-first pixel (the union-find's label)
        while (sp > 0) {
            const int i = stack[--sp];
            if (own[i] != -3) continue;
            own[i] = -1;
            list[np++] = i;
            const int4 q = nb[i];
            if (q.x >= 0 && own[q.x] == -3) stack[sp++] = q.x;
            if (q.y >= 0 && own[q.y] == -3) stack[sp++] = q.y;
            if (q.z >= 0 && own[q.z] == -3) stack[sp++] = q.z;
            if (q.w >= 0 && own[q.w] == -3) stack[sp++] = q.w;
        }
    }
    __syncthreads();
    // keys: value bits (positive floats order like their bit patterns) over the extraction position
    unsigned long long *K = reinterpret_cast<unsigned long long *>(stack + (stack_at & 1));
    int N = 2;
    while (N < n) N <<= 1;
    float peak = -INFINITY;
    for (int q = tid; q < N; q += NT) {
        unsigned long long key = ~0ull;
        if (q < n) {
            const float v = cval[list[q]];
            peak = fmaxf(peak, v);
            key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)q;
        }
        K[q] = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if (lane == 0) red[wid] = peak;
    __syncthreads();
    for (int k2 = 2; k2 <= N; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (N >> 1); t += NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const unsigned long long a = K[i], b = K[l];
                if ((a > b) == ((i & k2) == 0)) { K[i] = b; K[l] = a; }
            }
            __syncthreads();
        }
    if (wid == 0) {
        peak = red[0];
        for (int w2 = 1; w2 < NT / 32; w2++) peak = fmaxf(peak, red[w2]);
        float acc = 0.f;
        auto fetch = [&](int q, float &vs, float &ve, int &x, int &y) {
            vs = ve = 0.f; x = y = 0;
            if (q < n) {
                const int i = list[(int)(K[q] & 0xffffffffu)];
                vs = cval[i];
                const int lin = cidx[i];
                y = lin / g.W;
                x = lin - y * g.W;
                ve = cval[list[q]];
            }
        };
        float vs0, ve0;
        int x0, y0;
        fetch(lane, vs0, ve0, x0, y0);
        for (int q0 = 0; q0 < n; q0 += 32) {
            const float vsc = vs0, vec = ve0;
            const int xc = x0, yc = y0;
            if (q0 + 32 < n) fetch(q0 + 32 + lane, vs0, ve0, x0, y0);
            const int cnt = min(32, n - q0);
            for (int j = 0; j < cnt; j++) {
                const float v_s = __shfl_sync(0xffffffffu, vsc, j), v_e = __shfl_sync(0xffffffffu, vec, j);
                const int x = __shfl_sync(0xffffffffu, xc, j), y = __shfl_sync(0xffffffffu, yc, j);
                const double coord = lane == 0 ? (double)x + 0.5 : (lane == 1 ? (double)y + 0.5 : 1.0);
                acc = (float)__dadd_rn((double)acc, __dmul_rn(coord, (double)(lane == 3 ? v_e : v_s)));
            }
        }
        const float xa = __shfl_sync(0xffffffffu, acc, 0), ya = __shfl_sync(0xffffffffu, acc, 1);
        const float wt = __shfl_sync(0xffffffffu, acc, 2), value = __shfl_sync(0xffffffffu, acc, 3);
        if (lane == 0) {
            oss_star st;
            st.area = n; st.peak = peak; st.value = value;
            st.x = (int)__fdiv_rn(xa, wt);
            st.y = (int)__fdiv_rn(ya, wt);
            star_tmp[fb + off] = st;
            comp_nstar[fb + c] = 1;
        }
    }
    __syncthreads();
}

#define CBIG_NT 256
template <int CB> struct CompBigSmem {
    float val[CB];
    short up[CB], dn[CB];
    short lpos[CB];              // position in the reference's DFS list (getPeak's tie-break)
    short S[CB];                 // region being peeled, extraction order
    short stack[4 * CB + 8];     // DFS stack; staging: the member list (int, padded to a power of two); createStar: values + ranks
    unsigned char st[CB];        // bits 0-1: 0 unvisited, 1 remaining, 2 dropped, 3 accepted; bit 2: left neighbour, bit 3: right
    float red_v[CBIG_NT / 32];
    int red_p[CBIG_NT / 32];
    int bc[4];
};

template <int CB>
__global__ void __launch_bounds__(CBIG_NT)
k_components_big(const int *__restrict__ cand_idx, const float *__restrict__ cand_val, const int4 *__restrict__ nbr,
                 const int *__restrict__ comp_size, const int *__restrict__ comp_root, const int *__restrict__ comp_off,
                 int *__restrict__ comp_nstar, int *__restrict__ scratch, oss_star *__restrict__ star_tmp, int *__restrict__ owner,
                 const DetectState *__restrict__ ds, int nframes, Geom g)
{
    extern __shared__ __align__(16) unsigned char comp_smem_raw[];
    CompBigSmem<CB> &sm = *reinterpret_cast<CompBigSmem<CB> *>(comp_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr bool kLarge = CB > CBIG_SMALL;
    const int G = (int)gridDim.x;
    int base = 0; // (frame, list entry) pairs before this frame: pair k goes to CTA k % G
    for (int f = 0; f < nframes; f++) {
        if (ds[f].overflow) continue;
        const int nbig = kLarge ? ds[f].nbig3 : ds[f].nbig2;
        const long long fb = (long long)f * g.cap;
        int *scr = scratch + 12 * fb;
        const int *big_list = class_list(scratch, fb, g, kLarge ? 2 : 1);
        const float step = ds[f].threshold;
        const int e0 = (((int)blockIdx.x - base) % G + G) % G;
        base += nbig;
        for (int e = e0; e < nbig; e += G) {
            const int c = big_list[e];
            const int n = comp_size[fb + comp_root[fb + c]];
            const int off = comp_off[fb + c];
            if (n >= OSS_DEBLEND_MAX) { // adjoiningpixel.cpp:45-48: no deblending, ONE star from the whole component
                if (CB == OSS_DEBLEND_MAX - 1) big_bypass(c, comp_root[fb + c], n, off, fb, cand_idx, cand_val, nbr, owner, scr, star_tmp, comp_nstar, sm.red_v, g);
                continue;
            }
            const int *member = scr + off;                        // unordered (k_comp_members)
            int *reg = scr + g.cap + off;                         // accepted region of a pixel
            int *px = scr + 2ll * g.cap + off, *py = scr + 6ll * g.cap + 4ll * off + c;
            int *acc_gx = scr + 3ll * g.cap + off, *acc_gy = scr + 4ll * g.cap + off, *stamp = scr + 5ll * g.cap + off;
            oss_star *out = star_tmp + fb + off;
            __syncthreads(); // the previous component's shared memory is dead
            // ---- stage: member list -> ascending (raster order), bitonic sort of ints padded with INT_MAX
            int *A = reinterpret_cast<int *>(sm.stack);
            int N = 2;
            while (N < n) N <<= 1;
            for (int k = tid; k < N; k += CBIG_NT) A[k] = k < n ? member[k] : 0x7fffffff;
            __syncthreads();
            for (int k2 = 2; k2 <= N; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (N >> 1); t += CBIG_NT) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                        const int a = A[i], b = A[l];
                        if ((a > b) == ((i & k2) == 0)) { A[i] = b; A[l] = a; }
                    }
                    __syncthreads();
                }
            for (int k = tid; k < n; k += CBIG_NT) {
                const int i = A[k];
                sm.val[k] = cand_val[fb + i];
                const int lin = cand_idx[fb + i];
                const int y = lin / g.W;
                py[k] = y;
                px[k] = lin - y * g.W;
                const int4 q = nbr[fb + i];
                auto local = [&](int gl) {
                    if (gl < 0) return -1;
                    int lo = 0, hi = n - 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (A[mid] < gl) lo = mid + 1; else hi = mid; }
                    return lo;
                };
                sm.up[k] = (short)local(q.x);
                sm.dn[k] = (short)local(q.w);
                // a left / right neighbour above the threshold is the previous / next entry of the raster-ordered candidate
                // list and of this component's member list
                sm.st[k] = (unsigned char)((q.y >= 0 ? 4 : 0) | (q.z >= 0 ? 8 : 0));
            }
            __syncthreads(); // A (aliases the stack) is dead from here on
            // ---- detectAdjoiningPixel: explicit-stack DFS from the raster-first pixel, push order up, left, right, down
            if (tid == 0) {
                int np = 0, sp = 0;
                sm.stack[sp++] = 0;
                while (sp > 0) {
                    const int i = sm.stack[--sp];
                    const unsigned char s0 = sm.st[i];
                    if (s0 & 3) continue;
                    sm.st[i] = s0 | 1; // REMAINING
                    sm.lpos[i] = (short)np++;
                    const int u = sm.up[i], d = sm.dn[i];
                    const unsigned char su = sm.st[max(u, 0)], sl = sm.st[max(i - 1, 0)], sr = sm.st[min(i + 1, n - 1)], sd = sm.st[max(d, 0)];
                    if (u >= 0 && !(su & 3)) sm.stack[sp++] = (short)u;
                    if ((s0 & 4) && !(sl & 3)) sm.stack[sp++] = (short)(i - 1);
                    if ((s0 & 8) && !(sr & 3)) sm.stack[sp++] = (short)(i + 1);
                    if (d >= 0 && !(sd & 3)) sm.stack[sp++] = (short)d;
                }
            }
            __syncthreads();
            // ---- AdjoiningPixel::deblend (adjoiningpixel.cpp:41-102)
            int left = n, nacc = 0, iter = 0, nout = 0;
            while (left > 0) {
                iter++;
                // getPeak: the maximum among the remaining pixels, ties to the earlier DFS position
                float mx = 0.f;
                int pos = 0x7fffffff, pk = -1;
                for (int k = tid; k < n; k += CBIG_NT) {
                    if ((sm.st[k] & 3) != 1) continue;
                    const float v = sm.val[k];
                    const int lp = sm.lpos[k];
                    if (pos == 0x7fffffff || mx < v || (v == mx && lp < pos)) { mx = v; pos = lp; pk = k; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float v2 = __shfl_xor_sync(0xffffffffu, mx, o);
                    const int p2 = __shfl_xor_sync(0xffffffffu, pos, o), k2 = __shfl_xor_sync(0xffffffffu, pk, o);
                    if (p2 != 0x7fffffff && (pos == 0x7fffffff || mx < v2 || (v2 == mx && p2 < pos))) { mx = v2; pos = p2; pk = k2; }
                }
                if (lane == 0) { sm.red_v[wid] = mx; sm.red_p[wid] = pos == 0x7fffffff ? -1 : (pos | (pk << 16)); }
                __syncthreads();
                mx = 0.f; pos = 0x7fffffff; pk = -1;
#pragma unroll
                for (int w2 = 0; w2 < CBIG_NT / 32; w2++) {
                    const int pp = sm.red_p[w2];
                    if (pp < 0) continue;
                    const float v2 = sm.red_v[w2];
                    const int p2 = pp & 0xffff;
                    if (pos == 0x7fffffff || mx < v2 || (v2 == mx && p2 < pos)) { mx = v2; pos = p2; pk = pp >> 16; }
                }
                if (pk < 0) break;
                const float T = (float)__dsub_rn((double)mx, __dmul_rn((double)step, sqrt(__dadd_rn((double)__fdiv_rn(mx, step), 1.0)))); // adjoiningpixel.cpp:57
                // extract(): recursive pre-order up, left, right, down == stack with reversed pushes
                if (tid == 0) {
                    int ns0 = 0, sp = 0;
                    sm.stack[sp++] = (short)pk;
                    while (sp > 0) {
                        const int i = sm.stack[--sp];
                        const unsigned char s0 = sm.st[i];
                        const float vi = sm.val[i];
                        if ((s0 & 3) != 1 || !(vi > T)) continue;
                        sm.st[i] = (s0 & ~3) | 2; // DROPPED unless accepted below
                        sm.S[ns0++] = (short)i;
                        const int u = sm.up[i], d = sm.dn[i];
                        const unsigned char su = sm.st[max(u, 0)], sl = sm.st[max(i - 1, 0)], sr = sm.st[min(i + 1, n - 1)], sd = sm.st[max(d, 0)];
                        if (d >= 0 && (sd & 3) == 1) sm.stack[sp++] = (short)d;
                        if ((s0 & 8) && (sr & 3) == 1) sm.stack[sp++] = (short)(i + 1);
                        if ((s0 & 4) && (sl & 3) == 1) sm.stack[sp++] = (short)(i - 1);
                        if (u >= 0 && (su & 3) == 1) sm.stack[sp++] = (short)u;
                    }
                    sm.bc[0] = ns0;
                }
                __syncthreads();
                const int ns = sm.bc[0];
                if (ns == 0) break; // step <= 0 or NaN: the reference would never terminate
                left -= ns;
                // getGravityCenter (adjoiningpixel.cpp:104-118): fp32, extraction order; lanes 0..2 of warp 0 carry gx, gy, gw
                if (wid == 0) {
                    float gacc = 0.f;
                    int i0 = lane < ns ? sm.S[lane] : 0;
                    float v0 = sm.val[i0];
                    int x0 = px[i0], y0 = py[i0];
                    for (int q0 = 0; q0 < ns; q0 += 32) {
                        const float vc = v0;
                        const int xc = x0, yc = y0;
                        if (q0 + 32 < ns) { // the next 32 operands are in flight while this chunk's chain runs
                            i0 = q0 + 32 + lane < ns ? sm.S[q0 + 32 + lane] : 0;
                            v0 = sm.val[i0]; x0 = px[i0]; y0 = py[i0];
                        }
                        const int cnt = min(32, ns - q0);
                        for (int j = 0; j < cnt; j++) {
                            const float v = __shfl_sync(0xffffffffu, vc, j);
                            const int x = __shfl_sync(0xffffffffu, xc, j), y = __shfl_sync(0xffffffffu, yc, j);
                            const float coord = lane == 0 ? (float)x : (lane == 1 ? (float)y : 1.f);
                            gacc = __fadd_rn(gacc, __fmul_rn(v, coord));
                        }
                    }
                    const float gx = __shfl_sync(0xffffffffu, gacc, 0), gy = __shfl_sync(0xffffffffu, gacc, 1), gw = __shfl_sync(0xffffffffu, gacc, 2);
                    if (lane == 0) { sm.bc[1] = (int)__fdiv_rn(gx, gw); sm.bc[2] = (int)__fdiv_rn(gy, gw); }
                }
                // isAdjoining against every accepted region
                for (int q = tid; q < ns; q += CBIG_NT) {
                    const int i = sm.S[q];
                    const unsigned char s0 = sm.st[i];
                    const int u = sm.up[i], d = sm.dn[i];
                    if (u >= 0 && (sm.st[u] & 3) == 3) stamp[reg[u]] = iter;
                    if ((s0 & 4) && (sm.st[i - 1] & 3) == 3) stamp[reg[i - 1]] = iter;
                    if ((s0 & 8) && (sm.st[i + 1] & 3) == 3) stamp[reg[i + 1]] = iter;
                    if (d >= 0 && (sm.st[d] & 3) == 3) stamp[reg[d]] = iter;
                }
                __syncthreads();
                const int cx = sm.bc[1], cy = sm.bc[2];
                int blending = 0;
                for (int a = tid; a < nacc; a += CBIG_NT) blending |= (stamp[a] == iter) || (acc_gx[a] == cx && acc_gy[a] == cy);
                blending = __syncthreads_or(blending);
                if (!blending) {
                    for (int q = tid; q < ns; q += CBIG_NT) { const int i = sm.S[q]; sm.st[i] |= 3; reg[i] = nacc; }
                    if (tid == 0) { acc_gx[nacc] = cx; acc_gy[nacc] = cy; stamp[nacc] = 0; }
                    nacc++;
                    // createStar (adjoiningpixel.cpp:120-150): values in extraction order + ranks, in the idle stack
                    float *sv = reinterpret_cast<float *>(sm.stack);
                    short *sorted = reinterpret_cast<short *>(sv + ns);
                    for (int q = tid; q < ns; q += CBIG_NT) sv[q] = sm.val[sm.S[q]];
                    __syncthreads();
                    float peak = -INFINITY;
                    for (int q = tid; q < ns; q += CBIG_NT) {
                        const float v = sv[q];
                        peak = fmaxf(peak, v);
                        int r = 0; // elements ordered before this one: ascending value, ties by extraction order
                        for (int p2 = 0; p2 < ns; p2++) {
                            const float u2 = sv[p2];
                            r += (u2 < v) || (u2 == v && p2 < q);
                        }
                        sorted[r] = (short)q;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
                    if (lane == 0) sm.red_v[wid] = peak;
                    __syncthreads();
                    if (wid == 0) {
                        peak = sm.red_v[0];
#pragma unroll
                        for (int w2 = 1; w2 < CBIG_NT / 32; w2++) peak = fmaxf(peak, sm.red_v[w2]);
                        // lanes 0 / 1: flux-weighted x / y sums in ascending-value order, lane 2 the weight, lane 3 `value` in
                        // extraction order; every step acc = float(double(acc) + coord * double(v)) (see create_star_warp)
                        float acc = 0.f;
                        int is = lane < ns ? sm.S[sorted[lane]] : 0;
                        float vs0 = sm.val[is], ve0 = lane < ns ? sv[lane] : 0.f;
                        int x0 = px[is], y0 = py[is];
                        for (int q0 = 0; q0 < ns; q0 += 32) {
                            const float vsc = vs0, vec = ve0;
                            const int xc = x0, yc = y0;
                            if (q0 + 32 < ns) {
                                const int q = q0 + 32 + lane;
                                is = q < ns ? sm.S[sorted[q]] : 0;
                                vs0 = sm.val[is]; ve0 = q < ns ? sv[q] : 0.f; x0 = px[is]; y0 = py[is];
                            }
                            const int cnt = min(32, ns - q0);
                            for (int j = 0; j < cnt; j++) {
                                const float v_s = __shfl_sync(0xffffffffu, vsc, j), v_e = __shfl_sync(0xffffffffu, vec, j);
                                const int x = __shfl_sync(0xffffffffu, xc, j), y = __shfl_sync(0xffffffffu, yc, j);
                                const double coord = lane == 0 ? (double)x + 0.5 : (lane == 1 ? (double)y + 0.5 : 1.0);
                                acc = (float)__dadd_rn((double)acc, __dmul_rn(coord, (double)(lane == 3 ? v_e : v_s)));
                            }
                        }
                        const float xa = __shfl_sync(0xffffffffu, acc, 0), ya = __shfl_sync(0xffffffffu, acc, 1);
                        const float wt = __shfl_sync(0xffffffffu, acc, 2), value = __shfl_sync(0xffffffffu, acc, 3);
                        if (lane == 0) {
                            oss_star st;
                            st.area = ns; st.peak = peak; st.value = value;
                            st.x = (int)__fdiv_rn(xa, wt);
                            st.y = (int)__fdiv_rn(ya, wt);
                            out[nout] = st;
                        }
                    }
                    nout++;
                }
                __syncthreads();
            }
            if (tid == 0) comp_nstar[fb + c] = nout;
        }
    }
}

// members of every kept component, gathered by label into the component's scratch slice (any order; k_components sorts)
__global__ void __launch_bounds__(256)
k_comp_members(const int *__restrict__ parent, const int *__restrict__ flag, const int *__restrict__ flagscan,
               const int *__restrict__ comp_off, int *__restrict__ fill, int *__restrict__ scratch,
               const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int n = ds[f].ncand;
    const long long fb = (long long)f * g.cap;
    int *member = scratch + 12 * fb;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int r = parent[fb + i];
        if (!flag[fb + r]) continue;
        const int c = flagscan[fb + r];
        const int pos = atomicAdd(&fill[fb + c], 1);
        member[comp_off[fb + c] + pos] = i;
    }
}

// final ordered star list: stars of component c go to star_out[staroff[c] ...]
__global__ void __launch_bounds__(256)
k_gather_stars(const oss_star *__restrict__ star_tmp, const int *__restrict__ comp_off,
               const int *__restrict__ comp_nstar, const int *__restrict__ comp_staroff,
               oss_star *__restrict__ star_out, const DetectState *__restrict__ ds, Geom g)
{
    const int f = blockIdx.y;
    if (ds[f].overflow) return;
    const int ncomp = ds[f].ncomp;
    const long long fb = (long long)f * g.cap;
    for (int c = blockIdx.x * 256 + threadIdx.x; c < ncomp; c += gridDim.x * 256) {
        int n = comp_nstar[fb + c], so = comp_staroff[fb + c], off = comp_off[fb + c];
        for (int s = 0; s < n; s++) star_out[fb + so + s] = star_tmp[fb + off + s];
    }
}

}  // namespace oss
