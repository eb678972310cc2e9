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

// Ordered triangle list for one star list, by the whole CTA (blockDim.x == 256).
// Thread t owns triples [t*39, t*39+39); kept ones are written in enumeration order.
__device__ int build_triangle_list(const int *xy, int n, float *tx, float *ty, uint32_t *ts, int *scan_tmp)
{
    constexpr int PER = (OSS_MAX_TRI + 255) / 256; // 39
    const int t0 = threadIdx.x * PER;
    unsigned long long keep = 0;
    int cnt = 0;
    for (int q = 0; q < PER; q++) {
        int t = t0 + q;
        float x, y; uint32_t s;
        if (t < OSS_MAX_TRI && make_triangle(xy, n, d_tri_ijk[t], &x, &y, &s)) { keep |= 1ull << q; cnt++; }
    }
    // exclusive scan of cnt over 256 threads
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
        for (int i = 0; i < 8; i++) { int v = scan_tmp[i]; scan_tmp[i] = run; run += v; }
        scan_tmp[8] = run;
    }
    __syncthreads();
    int pos = scan_tmp[w] + inc - cnt;
    const int total = scan_tmp[8];
    for (int q = 0; q < PER; q++) {
        if (!((keep >> q) & 1ull)) continue;
        float x, y; uint32_t s;
        make_triangle(xy, n, d_tri_ijk[t0 + q], &x, &y, &s);
        tx[pos] = x; ty[pos] = y; ts[pos] = s;
        pos++;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(256)
k_build_ref_triangles(const int *__restrict__ xy, const int *__restrict__ nxy, RefTriangles *__restrict__ ref)
{
    __shared__ int sxy[2 * OSS_NOBJS];
    __shared__ int tmp[9];
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
struct FocasSmem {
    float bx[OSS_MAX_TRI];      // sort keys (x), permuted by the quicksort
    float by[OSS_MAX_TRI];      // generation order
    uint32_t bs[OSS_MAX_TRI];   // generation order
    unsigned short bid[OSS_MAX_TRI]; // payload riding with bx: index into by/bs
    int table[OSS_NOBJS * OSS_NOBJS];
    int matches[3][OSS_MAX_MATCH + 1];
    float a[3 * OSS_MAX_MATCH], b[2 * OSS_MAX_MATCH];
    int xy1[2 * OSS_NOBJS], xy2[2 * OSS_NOBJS];
    int range[64][2];
    int tmp[9];
    int nB;
};

__global__ void __launch_bounds__(256)
k_focas(const RefTriangles *__restrict__ ref, const int *__restrict__ xy_tgt, const int *__restrict__ nxy_tgt,
        AlignOut *__restrict__ out, int *__restrict__ valid_total, int *__restrict__ dbg_matches, int *__restrict__ dbg_table)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FocasSmem &sm = *reinterpret_cast<FocasSmem *>(smem_raw);
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n2 = nxy_tgt[f];
    if (tid < 2 * OSS_NOBJS) {
        sm.xy1[tid] = ref->xy[tid];
        sm.xy2[tid] = tid < 2 * n2 ? xy_tgt[f * 2 * OSS_NOBJS + tid] : 0;
    }
    for (int i = tid; i < OSS_NOBJS * OSS_NOBJS; i += 256) sm.table[i] = 0;
    for (int i = tid; i < 3 * (OSS_MAX_MATCH + 1); i += 256) (&sm.matches[0][0])[i] = 0;
    __syncthreads();
    const int nB = build_triangle_list(sm.xy2, n2, sm.bx, sm.by, sm.bs, sm.tmp);
    for (int i = tid; i < nB; i += 256) sm.bid[i] = (unsigned short)i;
    __syncthreads();

    // sortTriangles: literal emulation (one thread; sub-ranges are independent, so the
    // order in which they are processed does not change the result)
    if (tid == 0) {
        float *x = sm.bx;
        unsigned short *id = sm.bid;
        int sp = 0;
        sm.range[sp][0] = 0; sm.range[sp][1] = nB - 1; sp++;
        while (sp > 0) {
            sp--;
            int l = sm.range[sp][0], r = sm.range[sp][1];
            while (r > l) {
                float v = x[r];
                int i = l - 1, j = r;
                for (;;) {
                    while (x[++i] < v) {}
                    while (j > 1 && x[--j] > v) {}
                    if (i >= j) break;
                    float tx = x[i]; x[i] = x[j]; x[j] = tx;
                    unsigned short ti = id[i]; id[i] = id[j]; id[j] = ti;
                }
                float tx = x[i]; x[i] = x[r]; x[r] = tx;
                unsigned short ti = id[i]; id[i] = id[r]; id[r] = ti;
                // continue with the smaller part, defer the larger one
                int ll = l, lr = i - 1, rl = i + 1, rr = r;
                if (lr - ll > rr - rl) {
                    if (lr > ll) { sm.range[sp][0] = ll; sm.range[sp][1] = lr; sp++; }
                    l = rl; r = rr;
                } else {
                    if (rr > rl) { sm.range[sp][0] = rl; sm.range[sp][1] = rr; sp++; }
                    l = ll; r = lr;
                }
            }
        }
    }
    __syncthreads();

    // vote: one A triangle per thread step (binSearchTriangles + checkTolerance)
    const double tol = 0.002, tol2 = 0.002 * 0.002;
    const int nA = ref->n;
    for (int t = tid; t < nA; t += 256) {
        const float ax = ref->x[t], ay = ref->y[t];
        int lo = 0, hi = nB - 1, mid = 0, guard = 0;
        bool found = false;
        while (!found && hi - lo > 1 && guard++ < 64) {
            mid = (lo + hi) / 2;
            float bx = sm.bx[mid];
            if (fabs((double)__fsub_rn(bx, ax)) < tol) found = true;
            else if ((double)ax < (double)bx - tol) hi = mid;
            else if ((double)ax > (double)bx + tol) lo = mid;
        }
        if (!found) continue;
        int first, last, i;
        for (i = mid; i > 0; i--)
            if (fabs((double)__fsub_rn(sm.bx[i], ax)) > tol) break;
        first = i;
        for (i = mid; i < nB - 1; i++)
            if (fabs((double)__fsub_rn(sm.bx[i], ax)) > tol) break;
        last = i;
        const uint32_t as = ref->s[t];
        for (i = first; i <= last; i++) {
            int id = sm.bid[i];
            float dy = __fsub_rn(ay, sm.by[id]);
            if ((double)dy < tol) {
                float dx = __fsub_rn(ax, sm.bx[i]);
                float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if ((double)d2 < tol2) {
                    uint32_t bs = sm.bs[id];
                    atomicAdd(&sm.table[(as & 255) * OSS_NOBJS + (bs & 255)], 1);
                    atomicAdd(&sm.table[((as >> 8) & 255) * OSS_NOBJS + ((bs >> 8) & 255)], 1);
                    atomicAdd(&sm.table[(as >> 16) * OSS_NOBJS + (bs >> 16)], 1);
                }
            }
        }
    }
    __syncthreads();

    if (tid == 0) {
        // focas.cpp:107-139: the 40 best-voted pairs plus ties with the 40th (stable, row-major scan)
        const int N = OSS_NOBJS;
        int *mi = sm.matches[0], *mj = sm.matches[1], *mn = sm.matches[2];
        int k = 0;
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                int n = sm.table[i * N + j];
                if (n <= 0) continue;
                bool grow = k < N;
                if (!grow && n < mn[N - 1]) continue;
                int l;
                for (l = k; l > 0 && n > mn[l - 1]; l--) { mi[l] = mi[l - 1]; mj[l] = mj[l - 1]; mn[l] = mn[l - 1]; }
                mi[l] = i; mj[l] = j; mn[l] = n;
                if (grow) k++;
                else {
                    int lim = k < OSS_MAX_MATCH ? k + 1 : k;
                    int nth = mn[N - 1];
                    for (k = N; k < lim && nth == mn[k]; k++) {}
                }
            }
        if (dbg_matches)
            for (int r = 0; r < 3; r++)
                for (int i = 0; i < OSS_MAX_MATCH; i++) dbg_matches[(f * 3 + r) * OSS_MAX_MATCH + i] = sm.matches[r][i];
        AlignOut o;
        o.k = k;
        o.err = dev_find_transform(mi, mj, k, sm.xy1, sm.xy2, sm.a, sm.b, o.M);
        out[f] = o;
        if (valid_total && o.err == 0) atomicAdd(valid_total, 1);
    }
    if (dbg_table)
        for (int i = tid; i < OSS_NOBJS * OSS_NOBJS; i += 256) dbg_table[f * OSS_NOBJS * OSS_NOBJS + i] = sm.table[i];
}

}  // namespace oss
