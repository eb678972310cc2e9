// Kernel group (3): FOCAS-style triangle alignment — triangle lists over the first 40
// stars, similarity voting, match selection and the Lawson-Hanson HFTI least-squares fit
// with the 3-sigma-style clipping loop.  One CTA per light frame, batched.
//
// Replaces (reference, libstacker/src/):
//   generateTriangleList / sidesPos   focas.cpp:5-84
//   findMatches                       focas.cpp:86-144
//   sortTriangles                     focas.cpp:146-210   (quicksort with the (j > 1) guard:
//                                     slot 0 can stay misplaced; emulated literally)
//   binSearchTriangles / checkTolerance  focas.cpp:212-279
//   findTransform                     focas.cpp:281-379
//   hfti_ / h12 / diff                3rdparty/focas/{hfti,h12,diff}.f (prebuilt Fortran)
#pragma once
#include "common.cuh"

namespace oss {

// (i,j,k), i<j<k<40 in lexicographic order, packed i | j<<8 | k<<16.  For n < 40 stars the
// triples with k >= n are skipped, which leaves exactly the lexicographic order for n.
__device__ uint32_t d_tri_ijk[OSS_MAX_TRI];

// Triangle list of the reference frame, built once per stack.
struct RefTriangles {
    float x[OSS_MAX_TRI], y[OSS_MAX_TRI];
    uint32_t s[OSS_MAX_TRI]; // s1 | s2<<8 | s3<<16
    int n;
    int nstars;              // min(40, stars)
    int xy[2 * OSS_NOBJS];
};

// first 40 (x, y) of a detection result -> compact int list for the alignment kernel
__global__ void k_pack_xy(const oss_star *__restrict__ star_out, const DetectState *__restrict__ ds, int cap,
                          int *__restrict__ xy, int *__restrict__ nxy)
{
    const int f = blockIdx.x, i = threadIdx.x;
    int n = min(ds[f].nstars, OSS_NOBJS);
    if (ds[f].overflow) n = 0;
    if (i < n) {
        oss_star s = star_out[(long long)f * cap + i];
        xy[f * 2 * OSS_NOBJS + 2 * i] = s.x;
        xy[f * 2 * OSS_NOBJS + 2 * i + 1] = s.y;
    }
    if (i == 0) nxy[f] = n;
}

__device__ __forceinline__ float star_dist(const int *xy, int a, int b)
{
    float dx = (float)(xy[2 * a] - xy[2 * b]);
    float dy = (float)(xy[2 * a + 1] - xy[2 * b + 1]);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// One triangle (focas.cpp:31-71).  Returns false when it is dropped (b/a >= 0.9 or k >= n).
__device__ __forceinline__ bool make_triangle(const int *xy, int n, uint32_t ijk, float *tx, float *ty, uint32_t *ts)
{
    int i = ijk & 255, j = (ijk >> 8) & 255, k = ijk >> 16;
    if (k >= n) return false;
    float dij = star_dist(xy, i, j), djk = star_dist(xy, j, k), dki = star_dist(xy, k, i);
    float a, b, c;
    int s1, s2, s3;
    if (dki > djk) {
        if (dki > dij) {
            a = dki;
            if (djk > dij) { b = djk; c = dij; s1 = k; s2 = i; s3 = j; }
            else           { b = dij; c = djk; s1 = i; s2 = k; s3 = j; }
        } else { a = dij; b = dki; c = djk; s1 = i; s2 = j; s3 = k; }
    } else if (djk > dij) {
        a = djk;
        if (dij > dki) { b = dij; c = dki; s1 = j; s2 = k; s3 = i; }
        else           { b = dki; c = dij; s1 = k; s2 = j; s3 = i; }
    } else { a = dij; b = djk; c = dki; s1 = j; s2 = i; s3 = k; }
    float r = __fdiv_rn(b, a);
    if (!((double)r < 0.9)) return false;
    *tx = r;
    *ty = __fdiv_rn(c, a);
    *ts = (uint32_t)s1 | ((uint32_t)s2 << 8) | ((uint32_t)s3 << 16);
    return true;
}

// Ordered triangle list for one star list, by the whole CTA (blockDim.x == FOCAS_NT).
// Thread t owns PER consecutive triples; kept ones are written in enumeration order.
#define FOCAS_NT 1024
#define FOCAS_NW (FOCAS_NT / 32)
__device__ int build_triangle_list(const int *xy, int n, float *tx, float *ty, uint32_t *ts, int *scan_tmp)
{
    constexpr int PER = (OSS_MAX_TRI + FOCAS_NT - 1) / FOCAS_NT; // 10
    const int t0 = threadIdx.x * PER;
    unsigned long long keep = 0;
    int cnt = 0;
    for (int q = 0; q < PER; q++) {
        int t = t0 + q;
        float x, y; uint32_t s;
        if (t < OSS_MAX_TRI && make_triangle(xy, n, d_tri_ijk[t], &x, &y, &s)) { keep |= 1ull << q; cnt++; }
    }
    // exclusive scan of cnt over the block
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) scan_tmp[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < FOCAS_NW; i++) { int v = scan_tmp[i]; scan_tmp[i] = run; run += v; }
        scan_tmp[FOCAS_NW] = run;
    }
    __syncthreads();
    int pos = scan_tmp[w] + inc - cnt;
    const int total = scan_tmp[FOCAS_NW];
    for (int q = 0; q < PER; q++) {
        if (!((keep >> q) & 1ull)) continue;
        float x, y; uint32_t s;
        make_triangle(xy, n, d_tri_ijk[t0 + q], &x, &y, &s);
        tx[pos] = x;
        if (ty) { ty[pos] = y; ts[pos] = s; }
        pos++;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(FOCAS_NT)
k_build_ref_triangles(const int *__restrict__ xy, const int *__restrict__ nxy, RefTriangles *__restrict__ ref)
{
    __shared__ int sxy[2 * OSS_NOBJS];
    __shared__ int tmp[FOCAS_NW + 1];
    const int n = nxy[0];
    if (threadIdx.x < 2 * OSS_NOBJS) {
        int v = threadIdx.x < 2 * n ? xy[threadIdx.x] : 0;
        sxy[threadIdx.x] = v;
        ref->xy[threadIdx.x] = v;
    }
    __syncthreads();
    int total = build_triangle_list(sxy, n, ref->x, ref->y, ref->s, tmp);
    if (threadIdx.x == 0) { ref->n = total; ref->nstars = n; }
}

// ---- HFTI (Lawson & Hanson 1973), fp32 storage / fp64 accumulators, no FMA ---------------
// 1-based accessors mirror the Fortran; a is m x 3, b is m x 2, leading dimension 120.
__device__ void dev_h12(int mode, int lpivot, int l1, int m, float *u, int iue, float *up, float *c, int ice, int icv, int ncv)
{
#define U(j) u[((j)-1) * iue]
#define CC(i) c[(i)-1]
    if (0 >= lpivot || lpivot >= l1 || l1 > m) return;
    float cl = fabsf(U(lpivot));
    if (mode != 2) {
        for (int j = l1; j <= m; j++) cl = fmaxf(fabsf(U(j)), cl);
        if (cl <= 0.f) return;
        float clinv = __fdiv_rn(1.f, cl);
        double t = __dmul_rn((double)U(lpivot), (double)clinv);
        double sm = __dmul_rn(t, t);
        for (int j = l1; j <= m; j++) { t = __dmul_rn((double)U(j), (double)clinv); sm = __dadd_rn(sm, __dmul_rn(t, t)); }
        float sm1 = (float)sm;
        cl = __fmul_rn(cl, __fsqrt_rn(sm1));
        if (U(lpivot) > 0.f) cl = -cl;
        *up = __fsub_rn(U(lpivot), cl);
        U(lpivot) = cl;
    } else if (cl <= 0.f) return;
    if (ncv <= 0) return;
    double b = __dmul_rn((double)*up, (double)U(lpivot));
    if (b >= 0.0) return;
    b = __ddiv_rn(1.0, b);
    int i2 = 1 - icv + ice * (lpivot - 1);
    int incr = ice * (l1 - lpivot);
    for (int j = 1; j <= ncv; j++) {
        i2 += icv;
        int i3 = i2 + incr, i4 = i3;
        double sm = __dmul_rn((double)CC(i2), (double)*up);
        for (int i = l1; i <= m; i++) { sm = __dadd_rn(sm, __dmul_rn((double)CC(i3), (double)U(i))); i3 += ice; }
        if (sm != 0.0) {
            sm = __dmul_rn(sm, b);
            CC(i2) = (float)__dadd_rn((double)CC(i2), __dmul_rn(sm, (double)*up));
            for (int i = l1; i <= m; i++) { CC(i4) = (float)__dadd_rn((double)CC(i4), __dmul_rn(sm, (double)U(i))); i4 += ice; }
        }
    }
#undef U
#undef CC
}

__device__ void dev_hfti(float *a, int m, float *b, float tau)
{
    constexpr int MD = OSS_MAX_MATCH, N = 3, NB = 2;
#define A(i, j) a[((j)-1) * MD + ((i)-1)]
#define B(i, j) b[((j)-1) * MD + ((i)-1)]
    float h[N + 1], g[N + 1];
    int ip[N + 1];
    const float factor = 0.001f;
    float hmax = 0.f;
    const int ldiag = m < N ? m : N;
    if (ldiag <= 0) return;
    for (int j = 1; j <= ldiag; j++) {
        int lmax = j;
        bool recompute = true;
        if (j != 1) {
            for (int l = j; l <= N; l++) {
                h[l] = __fsub_rn(h[l], __fmul_rn(A(j - 1, l), A(j - 1, l)));
                if (h[l] > h[lmax]) lmax = l;
            }
            if (__fsub_rn(__fadd_rn(hmax, __fmul_rn(factor, h[lmax])), hmax) > 0.f) recompute = false;
        }
        if (recompute) {
            lmax = j;
            for (int l = j; l <= N; l++) {
                float s = 0.f;
                for (int i = j; i <= m; i++) s = __fadd_rn(s, __fmul_rn(A(i, l), A(i, l)));
                h[l] = s;
                if (h[l] > h[lmax]) lmax = l;
            }
            hmax = h[lmax];
        }
        ip[j] = lmax;
        if (ip[j] != j) {
            for (int i = 1; i <= m; i++) { float t = A(i, j); A(i, j) = A(i, lmax); A(i, lmax) = t; }
            h[lmax] = h[j];
        }
        dev_h12(1, j, j + 1, m, &A(1, j), 1, &h[j], &A(1, j + 1 <= N ? j + 1 : j), 1, MD, N - j);
        dev_h12(2, j, j + 1, m, &A(1, j), 1, &h[j], b, 1, MD, NB);
    }
    int j;
    for (j = 1; j <= ldiag; j++)
        if (fabsf(A(j, j)) <= tau) break;
    const int k = j <= ldiag ? j - 1 : ldiag;
    const int kp1 = k + 1;
    if (k <= 0) {
        for (int jb = 1; jb <= NB; jb++)
            for (int i = 1; i <= N; i++) B(i, jb) = 0.f;
        return;
    }
    if (k != N)
        for (int ii = 1; ii <= k; ii++) {
            int i = kp1 - ii;
            dev_h12(1, i, kp1, N, &A(i, 1), MD, &g[i], a, MD, 1, i - 1);
        }
    for (int jb = 1; jb <= NB; jb++) {
        for (int l = 1; l <= k; l++) {
            double sm = 0.0;
            int i = kp1 - l;
            for (int jj = i + 1; jj <= k; jj++) sm = __dadd_rn(sm, __dmul_rn((double)A(i, jj), (double)B(jj, jb)));
            float sm1 = (float)sm;
            B(i, jb) = __fdiv_rn(__fsub_rn(B(i, jb), sm1), A(i, i));
        }
        if (k != N) {
            for (int jj = kp1; jj <= N; jj++) B(jj, jb) = 0.f;
            for (int i = 1; i <= k; i++) dev_h12(2, i, kp1, N, &A(i, 1), MD, &g[i], &B(1, jb), 1, MD, 1);
        }
        for (int jj = 1; jj <= ldiag; jj++) {
            int jx = ldiag + 1 - jj;
            if (ip[jx] != jx) {
                int l = ip[jx];
                float t = B(l, jb); B(l, jb) = B(jx, jb); B(jx, jb) = t;
            }
        }
    }
#undef A
#undef B
}

// findTransform (focas.cpp:281-379), one thread.  mi/mj are compacted in place.
__device__ int dev_find_transform(int *mi, int *mj, int m, const int *xy1, const int *xy2, float *a, float *b, float *M)
{
    constexpr int MD = OSS_MAX_MATCH;
    for (int i = 0; i < 6; i++) M[i] = 0.f;
    if (m < 6) return -1;
    int j = m < 12 ? m : 12;
    for (int i = 0; i < j; i++) {
        a[i] = (float)xy1[2 * mi[i]];
        a[MD + i] = (float)xy1[2 * mi[i] + 1];
        a[2 * MD + i] = 1.f;
        b[i] = (float)xy2[2 * mj[i]];
        b[MD + i] = (float)xy2[2 * mj[i] + 1];
    }
    dev_hfti(a, j, b, 0.1f);
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) M[r * 3 + c] = b[r * MD + c];
    for (;;) {
        float *res = a, *sorted = a + MD;
        for (int i = 0; i < m; i++) {
            float px = (float)xy1[2 * mi[i]], py = (float)xy1[2 * mi[i] + 1];
            float x = __fadd_rn(__fadd_rn(__fmul_rn(M[0], px), __fmul_rn(M[1], py)), M[2]);
            float y = __fadd_rn(__fadd_rn(__fmul_rn(M[3], px), __fmul_rn(M[4], py)), M[5]);
            x = __fsub_rn(x, (float)xy2[2 * mj[i]]);
            y = __fsub_rn(y, (float)xy2[2 * mj[i] + 1]);
            float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
            int q;
            for (q = i; q > 0 && r2 < sorted[q - 1]; q--) sorted[q] = sorted[q - 1];
            res[i] = r2;
            sorted[q] = r2;
        }
        int cut = (int)(0.6 * (double)m);
        float lim = __fmul_rn(3.f, sorted[cut]);
        if (lim >= sorted[m - 1]) break;
        int kept = 0;
        for (int i = 0; i < m; i++)
            if (res[i] < lim) {
                int i1 = mi[i], i2 = mj[i];
                mi[kept] = i1; mj[kept] = i2;
                a[kept] = (float)xy1[2 * i1];
                a[MD + kept] = (float)xy1[2 * i1 + 1];
                a[2 * MD + kept] = 1.f;
                b[kept] = (float)xy2[2 * i2];
                b[MD + kept] = (float)xy2[2 * i2 + 1];
                kept++;
            }
        m = kept;
        if (m < 6) return -1;
        dev_hfti(a, m, b, 0.1f);
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) M[r * 3 + c] = b[r * MD + c];
    }
    return 0;
}

// ---- K7  one CTA per light -------------------------------------------------------------
// sortTriangles emulation.  The reference's quicksort (focas.cpp:146-210) guards its down-scan
// with (j > 1) instead of (j > l).  Consequences (verified against the reference's object
// code on thousands of inputs, see tests/): partitions with l >= 2 sort correctly; the chain
// of partitions with l == 0 can strand one element at slot 0; every other slot ends sorted.
// So the final array is  [e*] + sorted(all other elements)  where e* is whatever the l == 0
// chain leaves in slot 0.  We therefore
//   (1) replay ONLY the l == 0 chain, each Hoare partition evaluated in parallel from the
//       ordered lists of up-scan stops (x >= pivot) and down-scan stops (x <= pivot, slot 1
//       being a forced stop, slot 0 never visited), keeping just the left part it leaves;
//   (2) sort everything with a bitonic network;
//   (3) address the result through a view that moves e* to slot 0.
#define FOCAS_SORT_MAX 16384
struct FocasSmem {
    float kx[FOCAS_SORT_MAX];          // keys
    unsigned short kid[FOCAS_SORT_MAX]; // payload: generation index of the triangle
    float by[OSS_MAX_TRI];             // generation order   (aliased by the chain's stop lists)
    uint32_t bs[OSS_MAX_TRI];          // generation order
    int table[OSS_NOBJS * OSS_NOBJS];
    int matches[3][OSS_MAX_MATCH + 1];
    float a[3 * OSS_MAX_MATCH], b[2 * OSS_MAX_MATCH];
    int xy1[2 * OSS_NOBJS], xy2[2 * OSS_NOBJS];
    int tmp[FOCAS_NW + 1];
    int scan[FOCAS_NW + 2];
    int s, ifinal, e_id, e_pos;
    int sel[2];
    int rank[OSS_NOBJS];
    unsigned short above[OSS_NOBJS];
    unsigned short afirst[OSS_MAX_TRI], alast[OSS_MAX_TRI]; // vote window per A triangle
};

__device__ __forceinline__ int block_scan_packed(int v, int *wsum, int *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < FOCAS_NW; i++) { int t = wsum[i]; wsum[i] = run; run += t; }
        *total = run;
    }
    __syncthreads();
    int r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

// (1) replay the l == 0 partition chain; returns (in sm.e_id) the triangle stranded at slot 0
__device__ void focas_left_chain(FocasSmem &sm, int nB)
{
    const int tid = threadIdx.x;
    unsigned short *gpos = reinterpret_cast<unsigned short *>(sm.by); // up-scan stops, ascending
    unsigned short *hpos = reinterpret_cast<unsigned short *>(sm.bs); // down-scan stops, ascending (read backwards)
    int r = nB - 1;
    while (r > 0) {
        const float v = sm.kx[r];
        if (r == 1) { // the down-scan cannot move at all: j stays on the pivot
            if (tid == 0 && sm.kx[0] >= v) {
                float tx = sm.kx[0]; sm.kx[0] = sm.kx[1]; sm.kx[1] = tx;
                unsigned short ti = sm.kid[0]; sm.kid[0] = sm.kid[1]; sm.kid[1] = ti;
            }
            break;
        }
        const int per = (r + FOCAS_NT - 1) / FOCAS_NT, p0 = tid * per, p1 = min(r, p0 + per);
        int cg = 0, ch = 0;
        for (int p = p0; p < p1; p++) {
            float x = sm.kx[p];
            cg += x >= v;
            ch += p >= 1 && (x <= v || p == 1);
        }
        int total;
        int ex = block_scan_packed(cg | (ch << 16), sm.scan, &sm.scan[FOCAS_NW + 1]);
        total = sm.scan[FOCAS_NW + 1];
        const int nG = total & 0xffff, nH = total >> 16;
        int g = ex & 0xffff, h = ex >> 16;
        for (int p = p0; p < p1; p++) {
            float x = sm.kx[p];
            if (x >= v) gpos[g++] = (unsigned short)p;
            if (p >= 1 && (x <= v || p == 1)) hpos[h++] = (unsigned short)p;
        }
        __syncthreads();
        if (tid == 0) { // number of swaps: pairs (k-th up stop, k-th down stop) with g_k < h_k, monotone in k
            int lo = 0, hi = min(nG, nH);
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (gpos[mid] < hpos[nH - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            int ifin = lo < nG ? gpos[lo] : r;            // next up-scan stop in the untouched array ...
            if (lo >= 1) ifin = min(ifin, (int)hpos[nH - lo]); // ... or the last swapped-in big element
            sm.s = lo;
            sm.ifinal = ifin;
        }
        __syncthreads();
        const int s = sm.s;
        for (int k = tid; k < s; k += FOCAS_NT) { // what the swaps leave in the LEFT part (sources lie right of it)
            int gp = gpos[k], hp = hpos[nH - 1 - k];
            sm.kx[gp] = sm.kx[hp];
            sm.kid[gp] = sm.kid[hp];
        }
        r = sm.ifinal - 1;
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) sm.e_id = nB > 0 ? sm.kid[0] : -1;
    __syncthreads();
}

// (2) bitonic sort of (kx, kid)[0..N), N a power of two, ascending by key
__device__ void focas_bitonic(FocasSmem &sm, int N)
{
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += FOCAS_NT) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                float a = sm.kx[i], b = sm.kx[l];
                bool asc = (i & k) == 0;
                if ((a > b) == asc && a != b) {
                    sm.kx[i] = b; sm.kx[l] = a;
                    unsigned short t2 = sm.kid[i]; sm.kid[i] = sm.kid[l]; sm.kid[l] = t2;
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(FOCAS_NT)
k_focas(const RefTriangles *__restrict__ ref, const int *__restrict__ xy_tgt, const int *__restrict__ nxy_tgt,
        AlignOut *__restrict__ out, int *__restrict__ valid_total, int *__restrict__ dbg_matches, int *__restrict__ dbg_table,
        long long *__restrict__ dbg_clk)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FocasSmem &sm = *reinterpret_cast<FocasSmem *>(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n2 = nxy_tgt[f];
    long long c0 = clock64();
    if (tid < 2 * OSS_NOBJS) {
        sm.xy1[tid] = ref->xy[tid];
        sm.xy2[tid] = tid < 2 * n2 ? xy_tgt[f * 2 * OSS_NOBJS + tid] : 0;
    }
    // (votes as fire-and-forget global RED.ADD into a per-frame table were measured slower than these shared-memory atomics:
    // 625 k vs 495 k cycles for the vote phase at 8000 triangles)
    for (int i = tid; i < OSS_NOBJS * OSS_NOBJS; i += FOCAS_NT) sm.table[i] = 0;
    for (int i = tid; i < 3 * (OSS_MAX_MATCH + 1); i += FOCAS_NT) (&sm.matches[0][0])[i] = 0;
    __syncthreads();
    // keys only, for the chain replay (by/bs are the chain's scratch)
    const int nB = build_triangle_list(sm.xy2, n2, sm.kx, nullptr, nullptr, sm.tmp);
    for (int i = tid; i < nB; i += FOCAS_NT) sm.kid[i] = (unsigned short)i;
    __syncthreads();
    long long c1 = clock64();
    focas_left_chain(sm, nB);
    long long c2 = clock64();
    // full triangle list again (keys were permuted by the replay), then the real sort
    build_triangle_list(sm.xy2, n2, sm.kx, sm.by, sm.bs, sm.tmp);
    int N = 2;
    while (N < nB) N <<= 1;
    for (int i = tid; i < N; i += FOCAS_NT) {
        sm.kid[i] = (unsigned short)i;
        if (i >= nB) sm.kx[i] = INFINITY;
    }
    __syncthreads();
    if (nB > 1) focas_bitonic(sm, N);
    // (3) view: F[0] = e*, F[i] = S[i-1] for 1 <= i <= e_pos, F[i] = S[i] beyond
    if (tid == 0) sm.e_pos = 0;
    __syncthreads();
    for (int i = tid; i < nB; i += FOCAS_NT)
        if (sm.kid[i] == sm.e_id) sm.e_pos = i;
    __syncthreads();
    const int e_pos = sm.e_pos;
    auto F = [&](int i) { return i == 0 ? e_pos : (i <= e_pos ? i - 1 : i); };
    long long c3 = clock64();

    // ---- vote (binSearchTriangles + checkTolerance, focas.cpp:212-279) ----------------------
    // phase 1, one A triangle per thread: the reference's bisection decides WHETHER a B triangle
    // within tolerance is found (its probes never test slots 0 and n-1, so small runs can be
    // missed) and is replayed literally; the [first,last] window it then widens by linear scans
    // is the in-tolerance run around `mid` plus one element each side (or slot 0 / n-1 when the
    // scan runs off the end), which two more bisections over the sorted part give directly.
    const double tol = 0.002;
    // the reference compares float differences with the doubles 0.002 and 0.002 * 0.002 (checkTolerance, focas.cpp:251-279);
    // for a float d, (double)d < t is d < (the smallest float >= t): the window loop then needs no fp64 conversion per entry
    // (it is issue-bound: one CTA, ~80 B triangles per A triangle).  tests/test_kernel_tables.py checks both constants.
    const float kTolUp = __uint_as_float(0x3b03126fu), kTol2Up = __uint_as_float(0x368637beu);
    const int nA = ref->n;
    auto in_tol = [&](int i, float ax) { return !(fabs((double)__fsub_rn(sm.kx[F(i)], ax)) > tol); };
    for (int t = tid; t < nA; t += FOCAS_NT) {
        const float ax = ref->x[t];
        int lo = 0, hi = nB - 1, mid = 0, guard = 0;
        bool found = false;
        while (!found && hi - lo > 1 && guard++ < 64) {
            mid = (lo + hi) / 2;
            float bx = sm.kx[F(mid)];
            if (fabs((double)__fsub_rn(bx, ax)) < tol) found = true;
            else if ((double)ax < (double)bx - tol) hi = mid;
            else if ((double)ax > (double)bx + tol) lo = mid;
        }
        int first = 2, last = 1;
        if (found) {
            int a = 1, b = mid; // smallest in-tolerance slot p in [1, mid]; first = p - 1
            while (a < b) { int m = (a + b) >> 1; if (in_tol(m, ax)) b = m; else a = m + 1; }
            first = a - 1;
            a = mid; b = nB - 2; // largest in-tolerance slot q in [mid, n-2]; last = q + 1
            if (b < a) b = a;
            while (a < b) { int m = (a + b + 1) >> 1; if (in_tol(m, ax)) a = m; else b = m - 1; }
            last = mid >= nB - 1 ? mid : a + 1;
        }
        sm.afirst[t] = (unsigned short)first;
        sm.alast[t] = (unsigned short)last;
    }
    __syncthreads();
    const long long c3b = clock64();
    // phase 2: a warp takes 32 consecutive A triangles at a time (one coalesced load of their x, y, vertices and window) and walks
    // each one's window with its lanes; votes are integer atomics, so their order does not matter.  Measured at ~8000 triangles
    // per list (clock64, tools/focas_clk.py): 380 k cycles, of which ~100 k the shared-memory atomics (~200 k votes), ~105 k the
    // two tolerance tests and ~175 k the bare window walk (~80 B triangles per A triangle); global RED votes were slower.
    for (int t0 = (tid >> 5) * 32; t0 < nA; t0 += FOCAS_NW * 32) {
        const int lane = tid & 31, tl = t0 + lane;
        float axl = 0.f, ayl = 0.f;
        uint32_t asl = 0;
        int fl = 2, ll = 1;
        if (tl < nA) { axl = ref->x[tl]; ayl = ref->y[tl]; asl = ref->s[tl]; fl = sm.afirst[tl]; ll = sm.alast[tl]; }
        const int cnt = min(32, nA - t0);
        for (int j = 0; j < cnt; j++) {
            const int first = __shfl_sync(0xffffffffu, fl, j), last = __shfl_sync(0xffffffffu, ll, j);
            if (first > last) continue;
            const float ax = __shfl_sync(0xffffffffu, axl, j), ay = __shfl_sync(0xffffffffu, ayl, j);
            const uint32_t as = __shfl_sync(0xffffffffu, asl, j);
            for (int i = first + lane; i <= last; i += 32) {
                const int fi = F(i);
                const int id = sm.kid[fi];
                float dy = __fsub_rn(ay, sm.by[id]);
                if (dy < kTolUp) { // signed test, focas.cpp:261: (double)dy < 0.002  <=>  dy < the smallest float above 0.002
                    float dx = __fsub_rn(ax, sm.kx[fi]);
                    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                    if (d2 < kTol2Up) { // (double)d2 < 4e-6
                        uint32_t bs = sm.bs[id];
                        atomicAdd(&sm.table[(as & 255) * OSS_NOBJS + (bs & 255)], 1);
                        atomicAdd(&sm.table[((as >> 8) & 255) * OSS_NOBJS + ((bs >> 8) & 255)], 1);
                        atomicAdd(&sm.table[(as >> 16) * OSS_NOBJS + (bs >> 16)], 1);
                    }
                }
            }
        }
    }
    __syncthreads();
    long long c4 = clock64();

    // ---- selection (focas.cpp:107-139) --------------------------------------------------------
    // The reference's running insertion keeps, of all voted pairs in row-major scan order, the
    // list sorted by votes (descending, stable), cut after the 40th entry plus its ties, never
    // longer than 120.  Same list, computed by ranking: T = 40th largest vote count; pairs above
    // T are ranked by counting, pairs equal to T by their scan-order prefix.
    {
        constexpr int NT = OSS_NOBJS * OSS_NOBJS, PER = (NT + FOCAS_NT - 1) / FOCAS_NT; // table slots per thread
        const int e0 = tid * PER, e1 = min(NT, e0 + PER);
        int nz = 0, mx = 0;
        for (int e = e0; e < e1; e++) { int n = sm.table[e]; nz += n > 0; mx = max(mx, n); }
        int tot;
        block_scan_packed(nz, sm.scan, &sm.scan[FOCAS_NW + 1]);
        tot = sm.scan[FOCAS_NW + 1];
        if (tid == 0) sm.sel[0] = 0;
        __syncthreads();
        atomicMax(&sm.sel[0], mx);
        __syncthreads();
        const int maxn = sm.sel[0];
        int T = 0; // all voted pairs qualify when there are at most 40 of them
        if (tot > OSS_NOBJS) { // largest v with count(n >= v) >= 40
            int lo = 1, hi = maxn;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1, c = 0;
                for (int e = e0; e < e1; e++) c += sm.table[e] >= mid;
                block_scan_packed(c, sm.scan, &sm.scan[FOCAS_NW + 1]);
                if (sm.scan[FOCAS_NW + 1] >= OSS_NOBJS) lo = mid; else hi = mid - 1;
                __syncthreads();
            }
            T = lo;
        }
        int cAbove = 0, cTie = 0;
        for (int e = e0; e < e1; e++) { int n = sm.table[e]; cAbove += n > T; cTie += (T > 0 && n == T); }
        int ex = block_scan_packed(cAbove | (cTie << 16), sm.scan, &sm.scan[FOCAS_NW + 1]);
        const int c1 = sm.scan[FOCAS_NW + 1] & 0xffff, cT = sm.scan[FOCAS_NW + 1] >> 16;
        const int k = min(c1 + cT, OSS_MAX_MATCH);
        // pairs above T: compact (any order), then rank each by counting over the whole table
        int ia = ex & 0xffff, it = ex >> 16;
        for (int e = e0; e < e1; e++) {
            int n = sm.table[e];
            if (n > T) sm.above[ia++] = (unsigned short)e;
            else if (T > 0 && n == T) {
                int rank = c1 + it++;
                if (rank < k) { sm.matches[0][rank] = e / OSS_NOBJS; sm.matches[1][rank] = e % OSS_NOBJS; sm.matches[2][rank] = n; }
            }
        }
        if (tid < OSS_NOBJS) sm.rank[tid] = 0;
        __syncthreads();
        { // thread -> (pair a, slice sl of the table): 6 slices of 267 slots for up to 40 pairs
            const int a = tid / 6, sl = tid % 6;
            if (a < c1 && tid < 240) {
                const int e = sm.above[a], n = sm.table[e];
                int r = 0;
                for (int q = sl * 267; q < min(NT, sl * 267 + 267); q++) {
                    int m = sm.table[q];
                    r += (m > n) || (m == n && q < e);
                }
                atomicAdd(&sm.rank[a], r);
            }
        }
        __syncthreads();
        if (tid < c1) {
            const int e = sm.above[tid], r = sm.rank[tid];
            sm.matches[0][r] = e / OSS_NOBJS; sm.matches[1][r] = e % OSS_NOBJS; sm.matches[2][r] = sm.table[e];
        }
        if (tid == 0) sm.sel[1] = k;
        __syncthreads();
    }

    if (tid == 0) {
        const int k = sm.sel[1];
        int *mi = sm.matches[0], *mj = sm.matches[1];
        if (dbg_matches)
            for (int r = 0; r < 3; r++)
                for (int i = 0; i < OSS_MAX_MATCH; i++) dbg_matches[(f * 3 + r) * OSS_MAX_MATCH + i] = sm.matches[r][i];
        long long c5 = clock64();
        AlignOut o;
        o.k = k;
        o.err = dev_find_transform(mi, mj, k, sm.xy1, sm.xy2, sm.a, sm.b, o.M);
        out[f] = o;
        if (valid_total && o.err == 0) atomicAdd(valid_total, 1);
        if (dbg_clk) {
            long long c6 = clock64();
            dbg_clk[0] = c1 - c0; dbg_clk[1] = c2 - c1; dbg_clk[2] = c3 - c2; dbg_clk[3] = c4 - c3; dbg_clk[4] = c5 - c4;
            dbg_clk[5] = c6 - c5; dbg_clk[6] = nB; dbg_clk[7] = c3b - c3; // [7]: the bisection phase of the vote
        }
    }
    if (dbg_table)
        for (int i = tid; i < OSS_NOBJS * OSS_NOBJS; i += FOCAS_NT) dbg_table[f * OSS_NOBJS * OSS_NOBJS + i] = sm.table[i];
}

}  // namespace oss
