// Kernel group (2a): sky-background estimation, sky subtraction and the ROI statistics.
//
// Replaces (reference, libstacker/src/stardetector.cpp):
//   generateSkyBackground :158-171  resize(W/8,H/8) -> medianBlur 5 -> resize(W,H) -> GaussianBlur 31x31
//   stars = gray - sky    :120
//   meanStdDev over Rect(W/20, H/20, int(.9W), int(.9H)), threshold / minPeak  :121-129
// Operation orders are the bit-exact models of SURVEY.md Appendix B.3-B.6 (no FMA).
#pragma once
#include "common.cuh"

namespace oss {

// One tap pair of cv::resize INTER_LINEAR along one axis (Appendix B.3).
struct ResizeTap { int i0, i1; float w0, w1; };

// axis_is_x: x zeroes the fraction at the borders, y keeps it but clamps the rows.
__global__ void k_resize_taps(ResizeTap *__restrict__ taps, int src_n, int dst_n, int axis_is_x)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dst_n) return;
    double scale = (double)src_n / (double)dst_n;
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    ResizeTap t;
    if (axis_is_x) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src_n - 1) { f = 0.f; s = src_n - 1; }
        t.i0 = s;
        t.i1 = min(s + 1, src_n - 1);
    } else {
        t.i0 = min(max(s, 0), src_n - 1);
        t.i1 = min(max(s + 1, 0), src_n - 1);
    }
    t.w0 = __fsub_rn(1.f, f);
    t.w1 = f;
    taps[d] = t;
}

__device__ __forceinline__ float lerp2(float a, float w0, float b, float w1)
{
    return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
}

// resize gray (W,H) -> (lw,lh): horizontal lerp of both source rows, then vertical.
__global__ void __launch_bounds__(256)
k_resize_down(const float *__restrict__ gray, long long strideGray, int W, const ResizeTap *__restrict__ tx,
              const ResizeTap *__restrict__ ty, float *__restrict__ lo, long long strideLo, int lw, int lh)
{
    int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (dx >= lw || dy >= lh) return;
    const float *g = gray + blockIdx.z * strideGray;
    ResizeTap a = tx[dx], b = ty[dy];
    const float *r0 = g + (long long)b.i0 * W, *r1 = g + (long long)b.i1 * W;
    float h0 = lerp2(__ldg(r0 + a.i0), a.w0, __ldg(r0 + a.i1), a.w1);
    float h1 = lerp2(__ldg(r1 + a.i0), a.w0, __ldg(r1 + a.i1), a.w1);
    lo[blockIdx.z * strideLo + (long long)dy * lw + dx] = lerp2(h0, b.w0, h1, b.w1);
}

// median-of-25 selection network: 99 compare-exchange pairs (a, b) -> (min, max)
__device__ constexpr unsigned char kMed25[198] = {
    0,1, 3,4, 2,4, 2,3, 6,7, 5,7, 5,6, 9,10, 8,10, 8,9, 12,13, 11,13, 11,12, 15,16, 14,16, 14,15, 18,19, 17,19, 17,18,
    21,22, 20,22, 20,21, 23,24, 2,5, 3,6, 0,6, 0,3, 4,7, 1,7, 1,4, 11,14, 8,14, 8,11, 12,15, 9,15, 9,12, 13,16, 10,16,
    10,13, 20,23, 17,23, 17,20, 21,24, 18,24, 18,21, 19,22, 8,17, 9,18, 0,18, 0,9, 10,19, 1,19, 1,10, 11,20, 2,20, 2,11,
    12,21, 3,21, 3,12, 13,22, 4,22, 4,13, 14,23, 5,23, 5,14, 15,24, 6,24, 6,15, 7,16, 7,19, 13,21, 15,23, 7,13, 7,15,
    1,9, 3,11, 5,17, 11,17, 9,17, 4,10, 6,12, 7,14, 4,6, 4,7, 12,14, 10,14, 6,7, 10,12, 6,10, 6,17, 12,17, 7,17, 7,10,
    12,18, 7,12, 10,18, 12,20, 10,20, 10,12};

// medianBlur ksize 5, BORDER_REPLICATE: exact median of 25 (Appendix B.4)
__global__ void __launch_bounds__(256)
k_median5(const float *__restrict__ lo, float *__restrict__ med, long long strideLo, int lw, int lh)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= lw || y >= lh) return;
    const float *s = lo + blockIdx.z * strideLo;
    float v[25];
#pragma unroll
    for (int dy = -2; dy <= 2; dy++) {
        int yy = min(max(y + dy, 0), lh - 1);
#pragma unroll
        for (int dx = -2; dx <= 2; dx++) {
            int xx = min(max(x + dx, 0), lw - 1);
            v[(dy + 2) * 5 + dx + 2] = __ldg(s + (long long)yy * lw + xx);
        }
    }
    // 99 compare-exchanges that leave the median of 25 in v[12] (a selection network, not a full sort; the
    // pair list is checked exhaustively with the 0/1 principle in tests/test_kernel_tables.py);
    // fully unrolled so v[] stays in registers
#pragma unroll
    for (int i = 0; i < 99; i++) {
        const int a = kMed25[2 * i], b = kMed25[2 * i + 1];
        const float lo_ = fminf(v[a], v[b]), hi_ = fmaxf(v[a], v[b]);
        v[a] = lo_; v[b] = hi_;
    }
    med[blockIdx.z * strideLo + (long long)y * lw + x] = v[12];
}

// cv::getGaussianKernel(31, sigma 5.0, CV_32F) as cv2 4.13 returns it (taps 0..15; symmetric)
__constant__ float c_gauss[31];
static const uint32_t kGaussBits[16] = {
    0x3a68cc99u, 0x3acfe4e8u, 0x3b325fb9u, 0x3b930b60u, 0x3be8ede4u, 0x3c314133u,
    0x3c819931u, 0x3cb61437u, 0x3cf5c7e8u, 0x3d1f615cu, 0x3d4699a0u, 0x3d6dc47au,
    0x3d88bfb5u, 0x3d972180u, 0x3da079edu, 0x3da3b7d6u};

// ---------------------------------------------------------------------------------------
// K3  sky_sub_stats, marching form.  One CTA (SKY_NT = 8 * SKY_H threads) owns a strip of SKY_W = 64 output columns and a
// segment of rows [ya, yb) of one frame and walks down it in steps of SKY_H = 16 rows.  The pieces of step b:
//   hrow     horizontal lerp of the low-res rows that are new for a block ("hrow" ring, keyed by low-res row & 15),
//            read from the segment's low-res patch staged in shared memory once; done one block ahead;
//   up       vertical lerp -> `up2`, the upsampled sky over the block's rows x (64 + 2*15) columns, rows interleaved
//            in pairs (operand layout of the packed row filter); the Gaussian's BORDER_REFLECT_101 is applied to the
//            FULL-RES coordinates of rows and columns; also one block ahead;
//   row      31-tap row filter, taps accumulated k = 0..30 in order; a thread makes 2 rows x 4 columns from one 36-wide
//            register window of row pairs; results go to a ring of SKY_NSLOT blocks of row-filtered rows (the first
//            SKY_LA blocks are mirrored behind the ring so that 34-row windows never wrap);
//   column   column filter (centre + 15 symmetric pairs) of the block whose 15-row lower apron has just been produced
//            (SKY_LA blocks behind): 4 rows x 2 columns per thread from a 34-row window; subtract from the gray tile
//            that cp.async fetched during the row filter; write; fp64 sum / sum of squares over the statistics ROI;
//            per-32-px-segment maximum (lets the thresholding pass skip empty segments).
// Order inside the loop: barrier B | row(b), hrow(b+1) | barrier C | column(b - SKY_LA), up(b+1), row taps of b+2.
// Two barriers per step; every global load of the steady state is issued one phase before its use.
// Marching removes the vertical apron of a tiled design (each row is upsampled and row-filtered
// once per strip instead of 1.23x) and every dependent global load from the steady state.
// HBM traffic: gray read once (4 B/px), stars written once (4 B/px).  About 112 non-fusable
// fp32 operations per pixel make this kernel FP32-pipe bound, not HBM bound.
// Packed fp32x2 arithmetic (Blackwell FMUL2 / FADD2): two IEEE-rounded fp32 operations per issue slot.
// Only products and the symmetric pair sums are packed; every accumulation stays a scalar add.rn —
// ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even with explicit .rn, which would break the
// bit-exact operation order, while it leaves FMUL2 followed by scalar FADD alone (checked in SASS).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t mul2_bcast(float k, f32x2_t b)
{
    f32x2_t r, kk;
    asm("mov.b64 %0, {%1, %1};" : "=l"(kk) : "f"(k));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(kk), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b)
{
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

__device__ __forceinline__ void cp_async16_sky(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}

#ifndef SKY_CTAS
#define SKY_CTAS (512 / SKY_NT)      // 16 warps per SM at 128 registers
#endif
#define SKY_W 64                    // strip width
#ifndef SKY_H
#define SKY_H 16                    // rows per step (16 or 32)
#endif
#define SKY_NT (SKY_H * 8)          // threads per CTA: 8 pixels per thread and step
#define SKY_LA (1 + 30 / SKY_H)     // row-filtered blocks the column pass of a block waits for (its 15-row lower apron)
#ifndef SKY_NSLOT
#define SKY_NSLOT (SKY_LA + 1)      // ring blocks: the LA + 1 a column pass reads; the row pass of the next step overwrites the oldest
#endif                              // one, whose column pass ended before barrier B
#define SKY_R 15                    // Gaussian radius
#define SKY_UW (SKY_W + 2 * SKY_R)  // 94 upsampled columns per strip
#ifndef SKY_ULD
#define SKY_ULD 98                  // float2 per row pair of up2: 34-wide windows starting at 4*gx stay inside; an ODD number
#endif                              // of 16-byte chunks (49), so neighbouring row pairs start 4 banks apart (see the row filter)
#ifndef SKY_HRS
#define SKY_HRS 8                   // hrow ring slots: the low-res rows of ONE block (SKY_H / 8 + 3 <= 7) — hrow(b+1) runs after up(b)
#endif
#define SKY_HLD 96                  // floats per hrow row
#define SKY_RING (SKY_NSLOT * SKY_H) // row-filtered ring rows; the first SKY_LA blocks are mirrored behind the ring
#define SKY_LRC 16                  // low-res patch columns (94/8 + 2 = 14 needed)
#ifndef SKY_SEG_MAX
#define SKY_SEG_MAX 1024            // rows per segment (multiple of SKY_H)
#endif
#define SKY_LRR (SKY_SEG_MAX / 8 + 16) // low-res patch rows of a segment with its aprons
// a row tap with the hrow-ring offsets of its two low-res rows already resolved
struct SkyRowTap { int o0, o1; float w0, w1; };
struct SkySmem {
    float rf[(SKY_RING + SKY_LA * SKY_H) * SKY_W];
    float2 up2[(SKY_H / 2) * SKY_ULD];
    float hr[SKY_HRS * SKY_HLD];             // 6 KB
    float gray[SKY_H * SKY_W];               // cp.async target
    float lr[SKY_LRR * SKY_LRC];             // 9 KB
    ResizeTap tcol[SKY_UW];
    SkyRowTap trow[2][SKY_H];
    int blk_lo[2], blk_hi[2];
    unsigned char mode[SKY_SEG_MAX / SKY_H]; // column-pass variant of every output block of the segment
    double red[2][SKY_NT / 32];
};
constexpr int kSkySmemBytes = (int)sizeof(SkySmem);

// image of the coordinate interval [va, vb] under reflect101(., n): [lo, hi]
__device__ __forceinline__ void reflect_range(int va, int vb, int n, int &lo, int &hi)
{
    if (va < -(n - 1) || vb > 2 * (n - 1)) { lo = 0; hi = n - 1; return; } // several folds: whole axis
    const int ra = reflect101(va, n), rb = reflect101(vb, n);
    lo = (va <= 0 && vb >= 0) ? 0 : min(ra, rb);
    hi = (va <= n - 1 && vb >= n - 1) ? n - 1 : max(ra, rb);
}

// column filter + subtraction + statistics + segment maxima of one 32-row output block.
// MODE 0: block inside the image and inside the statistics ROI; 1: inside the image, outside the ROI;
// 2: anything else (per-pixel tests).
template <int MODE>
__device__ __forceinline__ void sky_column_pass(const float *__restrict__ ring, const float *__restrict__ gtile, int ring_row0, int x0,
                                                int yblk, int tid, float *__restrict__ S, float *__restrict__ SM, int nseg,
                                                const Geom &g, double &sum, double &sq)
{
    const int gy = tid >> 5, cp = tid & 31, lane = cp;
    const int x = x0 + 2 * cp, y = yblk + 4 * gy;
    const f32x2_t *q = (const f32x2_t *)(ring + (ring_row0 + 4 * gy + 1) * SKY_W + 2 * cp);
    // sliding window: tap pair j needs rows 15 - j .. 18 - j and 15 + j .. 18 + j, so the loads are interleaved with the
    // arithmetic (j outer, the four output rows inner) and only ~10 window values are live at a time
    f32x2_t w[34];
#pragma unroll
    for (int j = 15; j < 19; j++) w[j] = q[j * (SKY_W / 2)];
    float ca[4], cb[4];
#pragma unroll
    for (int e = 0; e < 4; e++) unpack2(mul2_bcast(c_gauss[15], w[e + 15]), ca[e], cb[e]);
#pragma unroll
    for (int j = 1; j <= 15; j++) {
        w[18 + j] = q[(18 + j) * (SKY_W / 2)];
        w[15 - j] = q[(15 - j) * (SKY_W / 2)];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float u, v;
            unpack2(mul2_bcast(c_gauss[15 + j], add2(w[e + 15 + j], w[e + 15 - j])), u, v);
            ca[e] = __fadd_rn(ca[e], u);
            cb[e] = __fadd_rn(cb[e], v);
        }
    }
    const float2 *gq = (const float2 *)(gtile + (4 * gy) * SKY_W + 2 * cp);
    float *sp = S + (long long)y * g.W + x;
    float *mp = SM + (long long)y * nseg + (x0 >> 5); // warp-uniform: the two segments of this warp's row
    const bool seg2 = (x0 >> 5) + 1 < nseg;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float a = ca[e], b = cb[e];
        const float2 gv = gq[e * (SKY_W / 2)];
        float va = __fsub_rn(gv.x, a), vb = __fsub_rn(gv.y, b);
        if (MODE < 2) {
            *(float2 *)(sp + e * g.W) = make_float2(va, vb);
            if (MODE == 0) {
                const double da = (double)va, db = (double)vb;
                sum += da; sq = __fma_rn(da, da, sq); // da*da is exact in fp64: same value as sq + da*da
                sum += db; sq = __fma_rn(db, db, sq);
            }
        } else {
            const bool row_ok = y + e < g.H, in_y = y + e >= g.roi_y && y + e < g.roi_y + g.roi_h;
            const bool oka = row_ok && x < g.W, okb = row_ok && x + 1 < g.W;
            if (oka) sp[e * g.W] = va; else va = 0.f;
            if (okb) sp[e * g.W + 1] = vb; else vb = 0.f;
            if (oka && in_y && x >= g.roi_x && x < g.roi_x + g.roi_w) { const double d = (double)va; sum += d; sq = __fma_rn(d, d, sq); }
            if (okb && in_y && x + 1 >= g.roi_x && x + 1 < g.roi_x + g.roi_w) { const double d = (double)vb; sum += d; sq = __fma_rn(d, d, sq); }
        }
        // segment maxima clamped at 0 (thresholds are >= 0 and consumers only test `> threshold`): for non-negative
        // floats the unsigned order of the bit patterns is the float order.  Two full-mask REDUX, one per 16-lane half
        // (a partial-mask reduction compiles to a divergent slow path).
        const unsigned key = __float_as_uint(fmaxf(fmaxf(va, vb), 0.f));
        const unsigned klo = __reduce_max_sync(0xffffffffu, lane < 16 ? key : 0u);
        const unsigned khi = __reduce_max_sync(0xffffffffu, lane < 16 ? 0u : key);
        if (lane == 0 && (MODE < 2 || y + e < g.H)) {
            mp[e * nseg] = __uint_as_float(klo);
            if (MODE < 2 || seg2) mp[e * nseg + 1] = __uint_as_float(khi);
        }
    }
}

__global__ void __launch_bounds__(SKY_NT, SKY_CTAS)
k_sky_sub_stats(const float *__restrict__ gray, long long strideGray, const float *__restrict__ med,
                long long strideLo, int lw, const ResizeTap *__restrict__ ux, const ResizeTap *__restrict__ uy,
                float *__restrict__ stars, long long strideStars, float *__restrict__ segmax, long long strideSeg,
                int nseg, double *__restrict__ partials, long long stridePart, Geom g, int seg_h)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SkySmem &sm = *(SkySmem *)smem_raw;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int x0 = blockIdx.x * SKY_W, ya = blockIdx.y * seg_h, yb = min(ya + seg_h, g.H), f = blockIdx.z;
    const int nk = (yb - ya + SKY_H - 1) / SKY_H; // output blocks 0..nk-1
    const int bmax = nk - 1 + SKY_LA;             // row-filtered blocks 0..bmax; block j = rows ya - 16 + j*SKY_H ...
    const float *L = med + f * strideLo;
    const float *G = gray + f * strideGray;
    float *S = stars + f * strideStars;
    float *SM = segmax + f * strideSeg;

    // row taps of block bb (rows ya - 16 + 32 bb ...), by warp 0: the table lookup is issued a phase early
    auto fetch_row_tap = [&](int bb) { return uy[reflect101(ya - 16 + SKY_H * bb + (lane & (SKY_H - 1)), g.H)]; };
    auto publish_row_taps = [&](int bb, const ResizeTap &t) {
        SkyRowTap r;
        r.o0 = (t.i0 & (SKY_HRS - 1)) * SKY_HLD; r.o1 = (t.i1 & (SKY_HRS - 1)) * SKY_HLD; r.w0 = t.w0; r.w1 = t.w1;
        if (lane < SKY_H) sm.trow[bb & 1][lane] = r;
        const int lo = __reduce_min_sync(0xffffffffu, t.i0), hi = __reduce_max_sync(0xffffffffu, t.i1);
        if (lane == 0) { sm.blk_lo[bb & 1] = lo; sm.blk_hi[bb & 1] = hi; }
    };

    // ---- prologue: column taps, first block's row taps, the segment's low-res patch
    for (int i = tid; i < SKY_UW; i += SKY_NT) sm.tcol[i] = ux[reflect101(x0 - SKY_R + i, g.W)];
    if (wid == 0) publish_row_taps(0, fetch_row_tap(0));
    int xlo, xhi, ylo, yhi;
    reflect_range(x0 - SKY_R, x0 + SKY_W + SKY_R - 1, g.W, xlo, xhi);
    reflect_range(ya - 16, ya - 16 + SKY_H * (bmax + 1) - 1, g.H, ylo, yhi);
    const int pc0 = ux[xlo].i0, pc1 = ux[xhi].i1, pr0 = uy[ylo].i0, pr1 = uy[yhi].i1; // taps are monotone
    const bool patch_ok = pc1 - pc0 < SKY_LRC && pr1 - pr0 < SKY_LRR;
    if (patch_ok) {
        const int nrow = pr1 - pr0 + 1, ncol = pc1 - pc0 + 1;
        for (int i = tid; i < nrow * SKY_LRC; i += SKY_NT) {
            const int r = i >> 4, cc = i & 15;
            if (cc < ncol) sm.lr[i] = __ldg(L + (long long)(pr0 + r) * lw + pc0 + cc);
        }
    }
    __syncthreads();

    int hr_lo = 0, hr_hi = -1; // low-res rows currently valid in the hrow ring (block-uniform)
    // horizontal lerp of the low-res rows block bb needs and the ring does not hold yet: warps 0..2, thread = column
    auto hrow_pass = [&](int bb) {
        if (tid >= SKY_HLD) return; // warps 0..2 only (they all track the ring's valid range)
        const int lo = sm.blk_lo[bb & 1], hi = sm.blk_hi[bb & 1]; // hi - lo <= 32/8 + 2 < SKY_HRS (the scale is 8)
        int first;
        if (lo >= hr_lo && lo <= hr_hi + 1) { first = max(hr_hi + 1, lo); hr_hi = max(hr_hi, hi); }
        else { first = lo; hr_lo = lo; hr_hi = hi; }
        hr_lo = max(hr_lo, hr_hi - (SKY_HRS - 1));
        if (tid < SKY_UW && first <= hi) {
            const ResizeTap a = sm.tcol[tid];
            if (patch_ok) {
                const float *p0 = sm.lr + (first - pr0) * SKY_LRC - pc0;
                for (int r = first; r <= hi; r++, p0 += SKY_LRC)
                    sm.hr[(r & (SKY_HRS - 1)) * SKY_HLD + tid] = lerp2(p0[a.i0], a.w0, p0[a.i1], a.w1);
            } else {
                for (int r = first; r <= hi; r++) {
                    const float *row = L + (long long)r * lw;
                    sm.hr[(r & (SKY_HRS - 1)) * SKY_HLD + tid] = lerp2(__ldg(row + a.i0), a.w0, __ldg(row + a.i1), a.w1);
                }
            }
        }
    };
    // vertical lerp -> up2 for the 16 row pairs x 47 column pairs of block bb: products packed over the column pair,
    // sums scalar (lerp2's order); LDS.64 from the hrow ring, one STS.128 per item
    auto up_pass = [&](int bb) {
        const SkyRowTap *tr = sm.trow[bb & 1];
#pragma unroll
        for (int it = 0; it < ((SKY_H / 2) * (SKY_UW / 2) + SKY_NT - 1) / SKY_NT; it++) {
            const int i = tid + SKY_NT * it;
            if (i < (SKY_H / 2) * (SKY_UW / 2)) {
                const int rp = i / (SKY_UW / 2), c = 2 * (i - rp * (SKY_UW / 2));
                const SkyRowTap t0 = tr[2 * rp], t1 = tr[2 * rp + 1];
                const f32x2_t h0 = *(const f32x2_t *)(sm.hr + t0.o0 + c), h1 = *(const f32x2_t *)(sm.hr + t0.o1 + c);
                const f32x2_t g0 = *(const f32x2_t *)(sm.hr + t1.o0 + c), g1 = *(const f32x2_t *)(sm.hr + t1.o1 + c);
                float a0, a1, b0, b1, c0, c1, d0, d1;
                unpack2(mul2_bcast(t0.w0, h0), a0, a1);
                unpack2(mul2_bcast(t0.w1, h1), b0, b1);
                unpack2(mul2_bcast(t1.w0, g0), c0, c1);
                unpack2(mul2_bcast(t1.w1, g1), d0, d1);
                // (row 2rp, row 2rp+1) of column c, then of column c+1
                *(float4 *)(sm.up2 + rp * SKY_ULD + c) = make_float4(__fadd_rn(a0, b0), __fadd_rn(c0, d0), __fadd_rn(a1, b1), __fadd_rn(c1, d1));
            }
        }
    };
    // gray tile of the next output block -> sm.gray (asynchronously when rows are 16-byte aligned); called once per
    // output block, in order, so the source pointer and the row number just advance
    const bool vec_gray = (g.W & 3) == 0 && x0 + SKY_W <= g.W;
    const int pr = tid >> 4, pq = tid & 15;
    const float *gp = G + (long long)(ya + pr) * g.W + x0 + 4 * pq; // this thread's chunk of row pr (and pr + SKY_H/2)
    int grow = ya + pr;
    const unsigned gdst = (unsigned)__cvta_generic_to_shared(sm.gray + pr * SKY_W + 4 * pq);
    auto gray_prefetch = [&]() {
        if (vec_gray) {
            if (grow < g.H) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gdst), "l"(gp));
            if (grow + SKY_H / 2 < g.H)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gdst + (SKY_H / 2) * SKY_W * 4), "l"(gp + (long long)(SKY_H / 2) * g.W));
        } else {
            const int y0 = grow - pr;
            for (int i = tid; i < SKY_H * SKY_W; i += SKY_NT) {
                const int r = i >> 6, c = i & 63;
                sm.gray[i] = (y0 + r < g.H && x0 + c < g.W) ? __ldg(G + (long long)(y0 + r) * g.W + x0 + c) : 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        gp += (long long)SKY_H * g.W;
        grow += SKY_H;
    };

    // column-pass variant per output block: 0 inside the image and inside the statistics ROI, 1 inside the image and
    // outside the ROI, 2 anything else (per-pixel tests)
    if (tid < nk) {
        const int yblk = ya + SKY_H * tid;
        const bool in_image = x0 + SKY_W <= g.W && yblk + SKY_H <= g.H && (g.W & 1) == 0;
        const bool in_roi = x0 >= g.roi_x && x0 + SKY_W <= g.roi_x + g.roi_w && yblk >= g.roi_y && yblk + SKY_H <= g.roi_y + g.roi_h;
        const bool off_roi = x0 >= g.roi_x + g.roi_w || x0 + SKY_W <= g.roi_x || yblk >= g.roi_y + g.roi_h || yblk + SKY_H <= g.roi_y;
        sm.mode[tid] = (in_image && in_roi) ? 0 : (in_image && off_roi) ? 1 : 2;
    }

    hrow_pass(0);
    __syncthreads();
    up_pass(0);
    if (wid == 0 && bmax >= 1) publish_row_taps(1, fetch_row_tap(1));

    double sum = 0.0, sq = 0.0;
    int slot = 0; // ring block that step b writes = b % SKY_NSLOT

    // Two barriers per step.  Between B and C: row filter of block b + hrow of block b+1.  After C: column pass of
    // output block b - SKY_LA, then (no barrier needed) up2 / row taps for the next step.
    for (int b = 0; b <= bmax; b++) {
        __syncthreads(); // B: up2(b), row taps of b+1 visible; everyone left the previous column pass
        if (b >= SKY_LA) gray_prefetch(); // tile of output block b - SKY_LA; lands behind the row filter
        ResizeTap tnext;
        const bool taps_due = wid == 0 && b + 2 <= bmax;
        if (taps_due) tnext = fetch_row_tap(b + 2); // lands behind the row filter
        {   // rows (2rp, 2rp+1), columns 4gx .. 4gx+3 of this step's row-filtered block
            // a quarter-warp (8 lanes, one LDS.128 wavefront) = 4 neighbouring column groups x 2 neighbouring row pairs:
            // the groups start 32 B apart (banks 0,8,16,24) and the odd chunk stride shifts the second row pair by
            // 4 banks, so the 8 x 16 B of a wavefront cover all 32 banks (a plain (rp, gx) = (tid / 16, tid % 16)
            // mapping is a 2-way conflict on every window load)
            const int rp = 2 * wid + ((lane >> 2) & 1), gx = (lane & 3) | ((lane >> 3) << 2);
            const ulonglong2 *src = (const ulonglong2 *)(sm.up2 + rp * SKY_ULD + 4 * gx);
            // sliding window: tap j needs columns j .. j + 3, so the window loads are interleaved with the arithmetic
            // (taps outer, the four output columns inner: eight independent accumulation chains, few live window values)
            f32x2_t w[36];
            { ulonglong2 v = src[0]; w[0] = v.x; w[1] = v.y; v = src[1]; w[2] = v.x; w[3] = v.y; }
            float o0[4], o1[4];
#pragma unroll
            for (int j = 0; j < 31; j++) {
                if (j & 1) { ulonglong2 v = src[(j + 3) / 2]; w[j + 3] = v.x; w[j + 4] = v.y; }
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    float u, v;
                    unpack2(mul2_bcast(c_gauss[j], w[e + j]), u, v);
                    if (j == 0) { o0[e] = u; o1[e] = v; }
                    else { o0[e] = __fadd_rn(o0[e], u); o1[e] = __fadd_rn(o1[e], v); }
                }
            }
            float *d0 = sm.rf + (slot * SKY_H + 2 * rp) * SKY_W + 4 * gx;
            ((float4 *)d0)[0] = make_float4(o0[0], o0[1], o0[2], o0[3]);
            ((float4 *)(d0 + SKY_W))[0] = make_float4(o1[0], o1[1], o1[2], o1[3]);
            if (slot < SKY_LA) { // mirror of the first blocks behind the ring: 34-row windows never wrap
                ((float4 *)(d0 + SKY_RING * SKY_W))[0] = make_float4(o0[0], o0[1], o0[2], o0[3]);
                ((float4 *)(d0 + (SKY_RING + 1) * SKY_W))[0] = make_float4(o1[0], o1[1], o1[2], o1[3]);
            }
        }
        if (b < bmax) hrow_pass(b + 1);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads(); // C: row-filtered block b, hrow rows of b+1 and the gray tile are visible
        if (b >= SKY_LA) {
            const int yblk = ya + SKY_H * (b - SKY_LA);
            const int ring_row0 = (slot >= SKY_LA ? slot - SKY_LA : slot + SKY_NSLOT - SKY_LA) * SKY_H; // block b - SKY_LA
            const float *gtile = sm.gray;
            const int mode = sm.mode[b - SKY_LA];
            if (mode == 0) sky_column_pass<0>(sm.rf, gtile, ring_row0, x0, yblk, tid, S, SM, nseg, g, sum, sq);
            else if (mode == 1) sky_column_pass<1>(sm.rf, gtile, ring_row0, x0, yblk, tid, S, SM, nseg, g, sum, sq);
            else sky_column_pass<2>(sm.rf, gtile, ring_row0, x0, yblk, tid, S, SM, nseg, g, sum, sq);
        }
        if (b < bmax) {
            up_pass(b + 1);
            if (taps_due) publish_row_taps(b + 2, tnext);
        }
        slot = slot == SKY_NSLOT - 1 ? 0 : slot + 1;
    }
    // block reduction of the fp64 partial sums, fixed order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) { sm.red[0][wid] = sum; sm.red[1][wid] = sq; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < SKY_NT / 32; w++) { a += sm.red[0][w]; b += sm.red[1][w]; }
        const long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        partials[f * stridePart + 2 * blk] = a;
        partials[f * stridePart + 2 * blk + 1] = b;
    }
}

#define SKR_HRS 16                  // hrow ring slots of k_sky_rows (the rows of two consecutive blocks coexist)
// =======================================================================================
// K3 split in two (alternative to k_sky_sub_stats, OSS_B200_SKY_SPLIT=1): the same arithmetic, cut where the data flow is narrowest.
//   k_sky_rows   upsample + 31-tap ROW filter of every image row -> `rf` plane in HBM (4 B/px written).  No vertical
//                dependency at all: a CTA marches down a 64-column strip, one barrier per 16-row step (the upsampled
//                block is double-buffered), FP32-pipe bound (61 + 4.4 non-fusable ops per pixel).
//   k_sky_cols   COLUMN filter + subtraction + statistics + segment maxima: a CTA marches down a strip with a ring of rf
//                rows that cp.async fills from HBM one block ahead (rows beyond the frame are BORDER_REFLECT_101 of rows the
//                first kernel wrote: filtering a mirrored row gives the mirrored row's result), HBM bound: rf read 4 B/px +
//                gray read 4 + stars written 4.
// The fused kernel spends its time in phases (up, row, column, two barriers per step) whose different instruction mixes
// four warps per scheduler cannot overlap well: 59 % FP32-pipe utilisation, while the row filter ALONE reaches 85 %
// (tools/ubench/rowpass.cu).  Split, each kernel has one dense phase; the price is 8 B/px of extra HBM traffic that
// neither kernel's bound notices.
struct SkyRowsSmem {
    float2 up2[2][(SKY_H / 2) * SKY_ULD];    // double-buffered upsampled block (row pairs interleaved)
    float hr[SKR_HRS * SKY_HLD];
    float lr[SKY_LRR * SKY_LRC];
    ResizeTap tcol[SKY_UW];
    SkyRowTap trow[4][SKY_H];
    int blk_lo[4], blk_hi[4];
};
constexpr int kSkyRowsSmemBytes = (int)sizeof(SkyRowsSmem);

__global__ void __launch_bounds__(SKY_NT, 4)
k_sky_rows(const float *__restrict__ med, long long strideLo, int lw, const ResizeTap *__restrict__ ux,
           const ResizeTap *__restrict__ uy, float *__restrict__ rf, long long strideRf, Geom g, int seg_h)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SkyRowsSmem &sm = *(SkyRowsSmem *)smem_raw;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int x0 = blockIdx.x * SKY_W, ya = blockIdx.y * seg_h, yb = min(ya + seg_h, g.H), f = blockIdx.z;
    const int nb = (yb - ya + SKY_H - 1) / SKY_H; // blocks 0..nb-1; block j = rows ya + j*SKY_H ...
    const float *L = med + f * strideLo;
    float *R = rf + f * strideRf;

    auto fetch_row_tap = [&](int bb) { return uy[min(ya + SKY_H * bb + (lane & (SKY_H - 1)), g.H - 1)]; };
    auto publish_row_taps = [&](int bb, const ResizeTap &t) {
        SkyRowTap r;
        r.o0 = (t.i0 & (SKR_HRS - 1)) * SKY_HLD; r.o1 = (t.i1 & (SKR_HRS - 1)) * SKY_HLD; r.w0 = t.w0; r.w1 = t.w1;
        if (lane < SKY_H) sm.trow[bb & 3][lane] = r;
        const int lo = __reduce_min_sync(0xffffffffu, t.i0), hi = __reduce_max_sync(0xffffffffu, t.i1);
        if (lane == 0) { sm.blk_lo[bb & 3] = lo; sm.blk_hi[bb & 3] = hi; }
    };
    // ---- prologue: column taps, row taps of blocks 0..2, the segment's low-res patch
    for (int i = tid; i < SKY_UW; i += SKY_NT) sm.tcol[i] = ux[reflect101(x0 - SKY_R + i, g.W)];
    if (wid == 0) {
        publish_row_taps(0, fetch_row_tap(0));
        if (nb > 1) publish_row_taps(1, fetch_row_tap(1));
        if (nb > 2) publish_row_taps(2, fetch_row_tap(2));
    }
    int xlo, xhi;
    reflect_range(x0 - SKY_R, x0 + SKY_W + SKY_R - 1, g.W, xlo, xhi);
    const int pc0 = ux[xlo].i0, pc1 = ux[xhi].i1, pr0 = uy[ya].i0, pr1 = uy[min(ya + SKY_H * nb, g.H) - 1].i1; // taps are monotone
    const bool patch_ok = pc1 - pc0 < SKY_LRC && pr1 - pr0 < SKY_LRR;
    if (patch_ok) {
        const int nrow = pr1 - pr0 + 1, ncol = pc1 - pc0 + 1;
        for (int i = tid; i < nrow * SKY_LRC; i += SKY_NT) {
            const int r = i >> 4, cc = i & 15;
            if (cc < ncol) sm.lr[i] = __ldg(L + (long long)(pr0 + r) * lw + pc0 + cc);
        }
    }
    __syncthreads();

    int hr_lo = 0, hr_hi = -1; // low-res rows currently valid in the hrow ring (uniform over warps 0..2)
    auto hrow_pass = [&](int bb) {
        if (tid >= SKY_HLD) return;
        const int lo = sm.blk_lo[bb & 3], hi = sm.blk_hi[bb & 3];
        int first;
        if (lo >= hr_lo && lo <= hr_hi + 1) { first = max(hr_hi + 1, lo); hr_hi = max(hr_hi, hi); }
        else { first = lo; hr_lo = lo; hr_hi = hi; }
        hr_lo = max(hr_lo, hr_hi - (SKR_HRS - 1));
        if (tid < SKY_UW && first <= hi) {
            const ResizeTap a = sm.tcol[tid];
            if (patch_ok) {
                const float *p0 = sm.lr + (first - pr0) * SKY_LRC - pc0;
                for (int r = first; r <= hi; r++, p0 += SKY_LRC)
                    sm.hr[(r & (SKR_HRS - 1)) * SKY_HLD + tid] = lerp2(p0[a.i0], a.w0, p0[a.i1], a.w1);
            } else {
                for (int r = first; r <= hi; r++) {
                    const float *row = L + (long long)r * lw;
                    sm.hr[(r & (SKR_HRS - 1)) * SKY_HLD + tid] = lerp2(__ldg(row + a.i0), a.w0, __ldg(row + a.i1), a.w1);
                }
            }
        }
    };
    // this thread's items of the upsample pass (row pair, column pair) are the same in every step: resolved once
    constexpr int UP_ITEMS = (SKY_H / 2) * (SKY_UW / 2), UP_ITERS = (UP_ITEMS + SKY_NT - 1) / SKY_NT;
    int up_rp2[UP_ITERS], up_c[UP_ITERS], up_dst[UP_ITERS];
#pragma unroll
    for (int it = 0; it < UP_ITERS; it++) {
        const int i = tid + SKY_NT * it, rp = i / (SKY_UW / 2), c = 2 * (i - rp * (SKY_UW / 2));
        up_rp2[it] = i < UP_ITEMS ? 2 * rp : -1;
        up_c[it] = c;
        up_dst[it] = rp * SKY_ULD + c;
    }
    auto up_pass = [&](int bb) {
        const SkyRowTap *tr = sm.trow[bb & 3];
        float2 *up2 = sm.up2[bb & 1];
#pragma unroll
        for (int it = 0; it < UP_ITERS; it++) {
            if (up_rp2[it] >= 0) {
                const int c = up_c[it];
                const SkyRowTap t0 = tr[up_rp2[it]], t1 = tr[up_rp2[it] + 1];
                const f32x2_t h0 = *(const f32x2_t *)(sm.hr + t0.o0 + c), h1 = *(const f32x2_t *)(sm.hr + t0.o1 + c);
                const f32x2_t g0 = *(const f32x2_t *)(sm.hr + t1.o0 + c), g1 = *(const f32x2_t *)(sm.hr + t1.o1 + c);
                float a0, a1, b0, b1, c0, c1, d0, d1;
                unpack2(mul2_bcast(t0.w0, h0), a0, a1);
                unpack2(mul2_bcast(t0.w1, h1), b0, b1);
                unpack2(mul2_bcast(t1.w0, g0), c0, c1);
                unpack2(mul2_bcast(t1.w1, g1), d0, d1);
                *(float4 *)(up2 + up_dst[it]) = make_float4(__fadd_rn(a0, b0), __fadd_rn(c0, d0), __fadd_rn(a1, b1), __fadd_rn(c1, d1));
            }
        }
    };
    hrow_pass(0);
    if (nb > 1) hrow_pass(1); // rows of two consecutive blocks coexist in the ring: <= 2 * (16 / 8) + 3 <= SKR_HRS
    __syncthreads();
    up_pass(0);

    const bool vec_out = (g.W & 3) == 0 && x0 + SKY_W <= g.W;
    // this thread's output position (rows 2 rp, 2 rp + 1 of a block, 4 columns), advanced by one block per step
    float *out_run = R + (long long)(ya + 2 * (2 * wid + ((lane >> 2) & 1))) * g.W + x0 + 4 * ((lane & 3) | ((lane >> 3) << 2));
    // One barrier per step.  Phase b: row filter of block b (reads up2[b & 1]) | up of block b+1 (writes up2[(b+1) & 1], reads
    // the hrow rows of block b+1) | hrow of block b+2 (other ring slots) | row taps of block b+3.
    for (int b = 0; b < nb; b++) {
        __syncthreads();
        ResizeTap tnext;
        const bool taps_due = wid == 0 && b + 3 < nb;
        if (taps_due) tnext = fetch_row_tap(b + 3);
        {
            const int rp = 2 * wid + ((lane >> 2) & 1), gx = (lane & 3) | ((lane >> 3) << 2); // conflict-free LDS.128 (see above)
            const ulonglong2 *src = (const ulonglong2 *)(sm.up2[b & 1] + rp * SKY_ULD + 4 * gx);
            f32x2_t w[36];
#pragma unroll
            for (int q = 0; q < 18; q++) { ulonglong2 v = src[q]; w[2 * q] = v.x; w[2 * q + 1] = v.y; }
            float o0[4], o1[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float a0, a1;
                unpack2(mul2_bcast(c_gauss[0], w[e]), a0, a1);
#pragma unroll
                for (int j = 1; j < 31; j++) {
                    float u, v;
                    unpack2(mul2_bcast(c_gauss[j], w[e + j]), u, v);
                    a0 = __fadd_rn(a0, u);
                    a1 = __fadd_rn(a1, v);
                }
                o0[e] = a0; o1[e] = a1;
            }
            const int y = ya + SKY_H * b + 2 * rp, x = x0 + 4 * gx;
            float *d0 = out_run;
            out_run += (long long)SKY_H * g.W;
            if (vec_out) {
                if (y < yb) ((float4 *)d0)[0] = make_float4(o0[0], o0[1], o0[2], o0[3]);
                if (y + 1 < yb) ((float4 *)(d0 + g.W))[0] = make_float4(o1[0], o1[1], o1[2], o1[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (x + e < g.W) {
                        if (y < yb) d0[e] = o0[e];
                        if (y + 1 < yb) d0[g.W + e] = o1[e];
                    }
            }
        }
        if (b + 1 < nb) up_pass(b + 1);
        if (b + 2 < nb) hrow_pass(b + 2);
        if (taps_due) publish_row_taps(b + 3, tnext);
    }
}

#define SKC_D 3                              // prefetch distance in steps: a step is < 1 us, an HBM round trip under load is not
#define SKC_NSLOT (SKY_LA + 1 + SKC_D)       // ring blocks: the SKY_LA + 1 a column pass reads + SKC_D in flight
#define SKC_RING (SKC_NSLOT * SKY_H)
struct SkyColsSmem {
    float rf[(SKC_RING + SKY_LA * SKY_H) * SKY_W]; // ring of row-filtered rows + mirror of the first SKY_LA blocks
    float gray[SKC_D + 1][SKY_H * SKY_W];          // cp.async targets
    unsigned char mode[SKY_SEG_MAX / SKY_H];
    double red[2][SKY_NT / 32];
};
constexpr int kSkyColsSmemBytes = (int)sizeof(SkyColsSmem);

__global__ void __launch_bounds__(SKY_NT, 4)
k_sky_cols(const float *__restrict__ gray, long long strideGray, const float *__restrict__ rf, long long strideRf,
           float *__restrict__ stars, long long strideStars, float *__restrict__ segmax, long long strideSeg, int nseg,
           double *__restrict__ partials, long long stridePart, Geom g, int seg_h)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SkyColsSmem &sm = *(SkyColsSmem *)smem_raw;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int x0 = blockIdx.x * SKY_W, ya = blockIdx.y * seg_h, yb = min(ya + seg_h, g.H), f = blockIdx.z;
    const int nk = (yb - ya + SKY_H - 1) / SKY_H; // output blocks 0..nk-1; rf blocks 0..nk-1+SKY_LA, block j = rows ya - 16 + j*SKY_H ...
    const float *G = gray + f * strideGray;
    const float *R = rf + f * strideRf;
    float *S = stars + f * strideStars;
    float *SM = segmax + f * strideSeg;
    const bool vec = (g.W & 3) == 0 && x0 + SKY_W <= g.W;
    const int pr = tid >> 4, pq = tid & 15; // this thread's 16-byte chunk of rows pr and pr + 8 of a block
    // Everything the step loop needs is kept as running values: shared-memory addresses of this thread's chunk, global
    // pointers that advance by one block (16 rows) per step, ring slots that count modulo SKC_NSLOT.  (Recomputing them per
    // step — reflections, 64-bit row products, j % 6 — cost more instructions than the arithmetic itself, and integer
    // multiply-adds share the FMA pipe with it.)
    const unsigned rf_s = (unsigned)__cvta_generic_to_shared(sm.rf + pr * SKY_W + 4 * pq);
    const unsigned gray_s = (unsigned)__cvta_generic_to_shared(&sm.gray[0][0] + pr * SKY_W + 4 * pq);
    const long long blk_f = (long long)SKY_H * g.W, half_b = (long long)(SKY_H / 2) * g.W * 4; // floats per block, bytes per 8 rows
    const float *rf_g = R + (long long)(ya - 16 + pr) * g.W + x0 + 4 * pq;   // row pr of rf block 0 (only dereferenced when inside the frame)
    const float *gray_g = G + (long long)(ya + pr) * g.W + x0 + 4 * pq;       // row pr of output block 0
    int ld_y0 = ya - 16, ld_slot = 0;   // first row and ring slot of the next rf block to load
    int gr_y0 = ya, gr_buf = 0;         // first row and buffer of the next gray tile to load

    auto cp16 = [](unsigned dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src)); };
    // next rf block -> ring slot ld_slot (+ its mirror copy); rows beyond the frame's ends are BORDER_REFLECT_101 rows
    auto load_next_block = [&]() {
        const unsigned dst = rf_s + ld_slot * (SKY_H * SKY_W * 4);
        if (vec && ld_y0 >= 0 && ld_y0 + SKY_H <= g.H) {          // the usual case: 16 rows inside the frame
            cp16(dst, rf_g);
            cp16(dst + (SKY_H / 2) * SKY_W * 4, (const char *)rf_g + half_b);
            if (ld_slot < SKY_LA) {
                cp16(dst + SKC_RING * SKY_W * 4, rf_g);
                cp16(dst + (SKC_RING + SKY_H / 2) * SKY_W * 4, (const char *)rf_g + half_b);
            }
        } else if (vec) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = pr + 8 * h, y = reflect101(ld_y0 + r, g.H);
                const float *src = R + (long long)y * g.W + x0 + 4 * pq;
                cp16(dst + h * (SKY_H / 2) * SKY_W * 4, src);
                if (ld_slot < SKY_LA) cp16(dst + (SKC_RING + h * (SKY_H / 2)) * SKY_W * 4, src);
            }
        } else {
            for (int i = tid; i < SKY_H * SKY_W; i += SKY_NT) {
                const int r = i >> 6, c = i & 63, y = reflect101(ld_y0 + r, g.H);
                const float v = x0 + c < g.W ? __ldg(R + (long long)y * g.W + x0 + c) : 0.f;
                sm.rf[(ld_slot * SKY_H + r) * SKY_W + c] = v;
                if (ld_slot < SKY_LA) sm.rf[(SKC_RING + ld_slot * SKY_H + r) * SKY_W + c] = v;
            }
        }
        rf_g += blk_f;
        ld_y0 += SKY_H;
        ld_slot = ld_slot == SKC_NSLOT - 1 ? 0 : ld_slot + 1;
    };
    auto load_next_gray = [&]() {
        const unsigned dst = gray_s + gr_buf * (SKY_H * SKY_W * 4);
        if (vec) {
            if (gr_y0 + pr < g.H) cp16(dst, gray_g);
            if (gr_y0 + pr + SKY_H / 2 < g.H) cp16(dst + (SKY_H / 2) * SKY_W * 4, (const char *)gray_g + half_b);
        } else {
            float *dstb = sm.gray[gr_buf];
            for (int i = tid; i < SKY_H * SKY_W; i += SKY_NT) {
                const int r = i >> 6, c = i & 63;
                dstb[i] = (gr_y0 + r < g.H && x0 + c < g.W) ? __ldg(G + (long long)(gr_y0 + r) * g.W + x0 + c) : 0.f;
            }
        }
        gray_g += blk_f;
        gr_y0 += SKY_H;
        gr_buf = gr_buf == SKC_D ? 0 : gr_buf + 1;
    };
    if (tid < nk) {
        const int yblk = ya + SKY_H * tid;
        const bool in_image = x0 + SKY_W <= g.W && yblk + SKY_H <= g.H && (g.W & 1) == 0;
        const bool in_roi = x0 >= g.roi_x && x0 + SKY_W <= g.roi_x + g.roi_w && yblk >= g.roi_y && yblk + SKY_H <= g.roi_y + g.roi_h;
        const bool off_roi = x0 >= g.roi_x + g.roi_w || x0 + SKY_W <= g.roi_x || yblk >= g.roi_y + g.roi_h || yblk + SKY_H <= g.roi_y;
        sm.mode[tid] = (in_image && in_roi) ? 0 : (in_image && off_roi) ? 1 : 2;
    }
    // prologue: SKC_D groups in flight.  Group d holds what output block d still lacks: rf block d + SKY_LA (group 0 also the
    // blocks below it) and its gray tile.  Blocks are loaded strictly in order, so the running state above stays in step.
#pragma unroll
    for (int d = 0; d < SKC_D; d++) {
        if (d == 0) for (int j = 0; j < SKY_LA; j++) load_next_block();
        load_next_block();      // (beyond the segment's last block: rows further down the frame, or reflections — loaded, never read)
        load_next_gray();
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    double sum = 0.0, sq = 0.0;
    int slot = 0, gbuf = 0, yblk = ya;
    // One barrier per step.  Step k: wait for group k (all but the SKC_D - 1 youngest) | barrier | queue group k + SKC_D into the
    // slots output block k - 1 has just released | column pass of output block k (rf blocks k .. k + SKY_LA).
    for (int k = 0; k < nk; k++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(SKC_D - 1) : "memory");
        __syncthreads();
        if (k + SKC_D < nk) { load_next_block(); load_next_gray(); }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int mode = sm.mode[k];
        const float *gt = sm.gray[gbuf];
        if (mode == 0) sky_column_pass<0>(sm.rf, gt, slot * SKY_H, x0, yblk, tid, S, SM, nseg, g, sum, sq);
        else if (mode == 1) sky_column_pass<1>(sm.rf, gt, slot * SKY_H, x0, yblk, tid, S, SM, nseg, g, sum, sq);
        else sky_column_pass<2>(sm.rf, gt, slot * SKY_H, x0, yblk, tid, S, SM, nseg, g, sum, sq);
        yblk += SKY_H;
        slot = slot == SKC_NSLOT - 1 ? 0 : slot + 1;
        gbuf = gbuf == SKC_D ? 0 : gbuf + 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) { sm.red[0][wid] = sum; sm.red[1][wid] = sq; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < SKY_NT / 32; w++) { a += sm.red[0][w]; b += sm.red[1][w]; }
        const long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        partials[f * stridePart + 2 * blk] = a;
        partials[f * stridePart + 2 * blk + 1] = b;
    }
}

// Rows per segment of the marching kernel.  Long segments amortise the warm-up rows (2 * 16 per segment); a launch of
// few frames is cut into shorter ones so that the grid still fills the GPU a few times over (148 SMs x 4 CTAs).
static inline int sky_segment_rows(int W, int H, int nframes)
{
    const int strips = (W + SKY_W - 1) / SKY_W;
    const int want_ctas = 3 * 148 * SKY_CTAS;
    int nseg = (H + SKY_SEG_MAX - 1) / SKY_SEG_MAX;
    const int per_frame = strips * (nframes > 0 ? nframes : 1);
    const int nseg_fill = (want_ctas + per_frame - 1) / per_frame;
    if (nseg_fill > nseg) nseg = nseg_fill;
    const int nseg_cap = (H + 4 * SKY_H - 1) / (4 * SKY_H); // segments of at least 4 steps
    if (nseg > nseg_cap) nseg = nseg_cap;
    if (nseg < 1) nseg = 1;
    const int rows = (H + nseg - 1) / nseg;
    return (rows + SKY_H - 1) / SKY_H * SKY_H;
}
// upper bound of the number of CTAs (= fp64 partial-sum pairs) per frame over all launch sizes
static inline int sky_max_blocks(int W, int H)
{
    const int rows = sky_segment_rows(W, H, 1);
    return ((W + SKY_W - 1) / SKY_W) * ((H + rows - 1) / rows);
}

// sigma, threshold = float(sigma*t*1.5), minPeak = float(sigma*t*2.0)  [stardetector.cpp:125-129]
__global__ void k_finalize_stats(const double *__restrict__ partials, long long stridePart, int nblocks,
                                 DetectState *__restrict__ ds, int threshold_coeff, Geom g)
{
    const int f = blockIdx.x;
    __shared__ double sh[2][256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) {
        a += partials[f * stridePart + 2 * i];
        b += partials[f * stridePart + 2 * i + 1];
    }
    sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) { sh[0][threadIdx.x] += sh[0][threadIdx.x + k]; sh[1][threadIdx.x] += sh[1][threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double scale = 1.0 / ((double)g.roi_w * (double)g.roi_h);
        double m = __dmul_rn(sh[0][0], scale);
        double var = __dsub_rn(__dmul_rn(sh[1][0], scale), __dmul_rn(m, m));
        if (var < 0.0) var = 0.0;
        double sigma = sqrt(var);
        DetectState d;
        d.mean = m; d.sigma = sigma;
        d.threshold = (float)(sigma * (double)threshold_coeff * 1.5);
        d.minpeak = (float)(sigma * (double)threshold_coeff * 2.0);
        d.ncand = 0; d.nroots = 0; d.ncomp = 0; d.nstars = 0; d.overflow = 0; d.pad = 0;
        ds[f] = d;
    }
}

}  // namespace oss
