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
// K3  sky_sub_stats, marching form.  One CTA (SKY_NT = 128 threads) owns a strip of SKY_W = 64 output columns and a
// segment of rows [ya, yb) of one frame and walks down it in steps of SKY_H = 32 rows.  The pieces of step b:
//   hrow     horizontal lerp of the low-res rows that are new for a block ("hrow" ring, keyed by low-res row & 7),
//            read from the segment's low-res patch staged in shared memory once; done one block ahead;
//   up       vertical lerp -> `up2`, the upsampled sky over the block's rows x (64 + 2*15) columns, rows interleaved
//            in pairs (operand layout of the packed row filter); the Gaussian's BORDER_REFLECT_101 is applied to the
//            FULL-RES coordinates of rows and columns; also one block ahead;
//   row      31-tap row filter, taps accumulated k = 0..30 in order; a thread makes 2 rows x 8 columns (16 chains) from a
//            38-wide window of row pairs that SLIDES: the window loads are interleaved with the taps, ~10 values live;
//            results go to a ring of SKY_NSLOT blocks of row-filtered rows (the first SKY_LA blocks are mirrored behind
//            the ring so that 38-row windows never wrap);
//   column   column filter (centre + 15 symmetric pairs) of the block whose 15-row lower apron has just been produced
//            (SKY_LA blocks behind): 8 rows x 2 columns per thread from a sliding 38-row window; subtract from the gray
//            pixels (loaded straight from HBM at the start of the pass); write; fp64 sum / sum of squares over the
//            statistics ROI; maximum per 32 px x 8 rows (lets the thresholding pass skip empty segments).
// Order inside the loop: barrier B | row(b), hrow(b+1) | barrier C | column(b - SKY_LA), up(b+1), row taps of b+2.
// Two barriers per step; every global load of the steady state is issued long before its use.
// Why the tiles are wide: three resources bound this kernel at nearly the same level — the FP32 pipe (113 non-fusable
// operations per pixel), the issue slots and the shared-memory wavefronts.  16 pixels per thread and pass (instead of 8)
// halve the window loads per pixel (row: 19 LDS.128 per 16 px, column: 38 LDS.64 per 16 px) and the per-step overhead.
// Marching removes the vertical apron of a tiled design (each row is upsampled and row-filtered
// once per strip instead of 1.23x) and every dependent global load from the steady state.
// HBM traffic: gray read once (4 B/px), stars written once (4 B/px).
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
// streaming (read-once) 8-byte load
__device__ __forceinline__ float2 ld_stream2(const float *p)
{
    float2 v;
    asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

#ifndef SKY_CTAS
#define SKY_CTAS 4                  // CTAs per SM (shared memory: 4 x 52 KB; registers: <= 128)
#endif
#define SKY_W 64                    // strip width
#define SKY_H 32                    // rows per step
#define SKY_NT 128                  // threads per CTA: 16 pixels per thread in the row pass (2 x 8) and in the column pass (8 x 2)
#define SKY_LA (1 + 30 / SKY_H)     // row-filtered blocks the column pass of a block waits for (its 15-row lower apron)
#define SKY_NSLOT (SKY_LA + 1)      // ring blocks: the LA + 1 a column pass reads; the row pass of the next step overwrites the oldest
                                    // one, whose column pass ended before barrier B
#define SKY_R 15                    // Gaussian radius
#define SKY_UW (SKY_W + 2 * SKY_R)  // 94 upsampled columns per strip
#define SKY_ULD 98                  // float2 per row pair of up2 (>= 94).  98 * 8 B = 16 B mod 128 B: the row filter's LDS.128 of 4
                                    // neighbouring row pairs x 2 neighbouring column groups (64 B apart) cover all 32 banks
#define SKY_HRS 8                   // hrow ring slots: the low-res rows of ONE block (SKY_H / 8 + 3 <= 7) — hrow(b+1) runs after up(b)
#define SKY_HLD 96                  // floats per hrow row
#define SKY_RING (SKY_NSLOT * SKY_H) // row-filtered ring rows; the first SKY_LA blocks are mirrored behind the ring
#define SKY_LRC 16                  // low-res patch columns (94/8 + 2 = 14 needed)
#define SKY_SEG_MAX 1024            // rows per segment (multiple of SKY_H)
#define SKY_LRR (SKY_SEG_MAX / 8 + 16) // low-res patch rows of a segment with its aprons
#ifndef SKY_ROW_ROLLED
#define SKY_ROW_ROLLED 1
#endif
static_assert(SKY_H == 32 && SKY_NT == 128 && SKY_W == 64, "thread mappings of the row and column passes");
// a row tap with the hrow-ring offsets of its two low-res rows already resolved
struct SkyRowTap { int o0, o1; float w0, w1; };
struct SkySmem {
    float rf[(SKY_RING + SKY_LA * SKY_H) * SKY_W]; // 16-byte chunks of a row are XOR-swizzled, see sky_swz
    float2 up2[(SKY_H / 2) * SKY_ULD];
    float hr[SKY_HRS * SKY_HLD];
    float lr[SKY_LRR * SKY_LRC];
    ResizeTap tcol[SKY_UW];
    SkyRowTap trow[2][SKY_H];
    int blk_lo[2], blk_hi[2];
    int patch[4];                            // low-res patch: usable, first row, first column (read by hrow once per step: kept out of
                                             // registers, where they would spill to local memory — an L2 round trip per step)
    unsigned char mode[SKY_SEG_MAX / SKY_H]; // column-pass variant of every output block of the segment
    double red[2][SKY_NT / 32];
};
constexpr int kSkySmemBytes = (int)sizeof(SkySmem);
static_assert(SKY_CTAS * (kSkySmemBytes + 1024) <= 228 * 1024, "shared memory of SKY_CTAS resident CTAs");

// Swizzle of the rf ring: the 16-byte chunk c of a row is stored at chunk c ^ sky_swz(row >> 1).  The row filter stores
// STS.128 from 4 neighbouring row pairs x 2 neighbouring 32-byte column groups per quarter-warp: without the swizzle they would
// share two bank groups (rows are 256 B apart); with it the 8 chunks cover all 32 banks.  The column filter reads LDS.64 of 16
// neighbouring column pairs of ONE row per half-warp: a permutation inside an aligned 128-byte line, conflict-free as before.
__device__ __forceinline__ int sky_swz(int rowpair) { return (rowpair & 1) | ((rowpair & 2) << 1); } // 0, 1, 4, 5

// image of the coordinate interval [va, vb] under reflect101(., n): [lo, hi]
__device__ __forceinline__ void reflect_range(int va, int vb, int n, int &lo, int &hi)
{
    if (va < -(n - 1) || vb > 2 * (n - 1)) { lo = 0; hi = n - 1; return; } // several folds: whole axis
    const int ra = reflect101(va, n), rb = reflect101(vb, n);
    lo = (va <= 0 && vb >= 0) ? 0 : min(ra, rb);
    hi = (va <= n - 1 && vb >= n - 1) ? n - 1 : max(ra, rb);
}

// column filter + subtraction + statistics + segment maxima of one 32-row output block: warp gy makes rows 8 gy .. 8 gy + 7,
// lane cp columns 2 cp, 2 cp + 1.
// MODE 0: block inside the image and inside the statistics ROI; 1: inside the image, outside the ROI;
// 2: anything else (per-pixel tests).
// gray pixels of this thread's 8 rows x 2 columns, issued BEFORE barrier C so that they arrive from HBM while the warp waits and
// filters (inside the column pass ptxas sinks the loads to ~250 instructions before their use: long-scoreboard stalls)
template <int MODE>
__device__ __forceinline__ void sky_load_gray(const float *__restrict__ G, int x0, int yblk, int tid, const Geom &g, float2 (&gv)[8])
{
    const int x = x0 + 2 * (tid & 31), y = yblk + 8 * (tid >> 5);
    const float *p = G + (long long)y * g.W + x;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        if (MODE < 2) gv[e] = ld_stream2(p + (long long)e * g.W);
        else {
            const bool row_ok = y + e < g.H;
            gv[e].x = row_ok && x < g.W ? __ldg(p + (long long)e * g.W) : 0.f;
            gv[e].y = row_ok && x + 1 < g.W ? __ldg(p + (long long)e * g.W + 1) : 0.f;
        }
    }
}

template <int MODE>
__device__ __forceinline__ void sky_column_pass(const float *__restrict__ ring, const float2 (&gv)[8], int ring_row0, int x0,
                                                int yblk, int tid, float *__restrict__ S, float *__restrict__ SM, int nseg,
                                                const Geom &g, double &sum, double &sq)
{
    const int gy = tid >> 5, cp = tid & 31, lane = cp;
    const int x = x0 + 2 * cp, y = yblk + 8 * gy;
    const long long pix = (long long)y * g.W + x;
    // window row j (0..37) = ring row ring_row0 + 8 gy + 1 + j; its swizzle class ((1 + j) >> 1) & 3 is static because
    // ring_row0 + 8 gy is a multiple of 8: four base pointers, immediate row offsets
    const float *rowbase = ring + (ring_row0 + 8 * gy + 1) * SKY_W + 2 * (cp & 1);
    const float *qk[4];
#pragma unroll
    for (int k = 0; k < 4; k++) qk[k] = rowbase + 4 * ((cp >> 1) ^ sky_swz(k));
#define SKY_WROW(j) (*(const f32x2_t *)(qk[(((j) + 1) >> 1) & 3] + (j) * SKY_W))
    // sliding window: tap pair j needs rows 15 - j .. 22 - j and 15 + j .. 22 + j, so the loads are interleaved with the
    // arithmetic (j outer, the eight output rows inner: 16 chains) and only ~18 window values are live at a time
    f32x2_t w[38];
#pragma unroll
    for (int j = 15; j < 23; j++) w[j] = SKY_WROW(j);
    float ca[8], cb[8];
#pragma unroll
    for (int e = 0; e < 8; e++) unpack2(mul2_bcast(c_gauss[15], w[e + 15]), ca[e], cb[e]);
#pragma unroll
    for (int j = 1; j <= 15; j++) {
        w[22 + j] = SKY_WROW(22 + j);
        w[15 - j] = SKY_WROW(15 - j);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            float u, v;
            unpack2(mul2_bcast(c_gauss[15 + j], add2(w[e + 15 + j], w[e + 15 - j])), u, v);
            ca[e] = __fadd_rn(ca[e], u);
            cb[e] = __fadd_rn(cb[e], v);
        }
    }
#undef SKY_WROW
    float *sp = S + pix;
    float vmax = 0.f; // maxima clamped at 0 (thresholds are >= 0 and consumers only test `> threshold`)
#pragma unroll
    for (int e = 0; e < 8; e++) {
        float va = __fsub_rn(gv[e].x, ca[e]), vb = __fsub_rn(gv[e].y, cb[e]);
        if (MODE < 2) {
            *(float2 *)(sp + (long long)e * g.W) = make_float2(va, vb);
            if (MODE == 0) {
                const double da = (double)va, db = (double)vb;
                sum += da; sq = __fma_rn(da, da, sq); // da*da is exact in fp64: same value as sq + da*da
                sum += db; sq = __fma_rn(db, db, sq);
            }
        } else {
            const bool row_ok = y + e < g.H, in_y = y + e >= g.roi_y && y + e < g.roi_y + g.roi_h;
            const bool oka = row_ok && x < g.W, okb = row_ok && x + 1 < g.W;
            if (oka) sp[(long long)e * g.W] = va; else va = 0.f;
            if (okb) sp[(long long)e * g.W + 1] = vb; else vb = 0.f;
            if (oka && in_y && x >= g.roi_x && x < g.roi_x + g.roi_w) { const double d = (double)va; sum += d; sq = __fma_rn(d, d, sq); }
            if (okb && in_y && x + 1 >= g.roi_x && x + 1 < g.roi_x + g.roi_w) { const double d = (double)vb; sum += d; sq = __fma_rn(d, d, sq); }
        }
        vmax = fmaxf(vmax, fmaxf(va, vb));
    }
    // one maximum per 32 px x 8 rows: for non-negative floats the unsigned order of the bit patterns is the float order.
    // Two full-mask REDUX, one per 16-lane half (a partial-mask reduction compiles to a divergent slow path).
    const unsigned key = __float_as_uint(vmax);
    const unsigned klo = __reduce_max_sync(0xffffffffu, lane < 16 ? key : 0u);
    const unsigned khi = __reduce_max_sync(0xffffffffu, lane < 16 ? 0u : key);
    if (lane == 0 && (MODE < 2 || y < g.H)) {
        float *mp = SM + (long long)(y / OSS_SEGROWS) * nseg + (x0 >> 5); // the two cells of this warp's rows
        mp[0] = __uint_as_float(klo);
        if (MODE < 2 || (x0 >> 5) + 1 < nseg) mp[1] = __uint_as_float(khi);
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

    // row taps of block bb (rows ya - 16 + 32 bb ...), by warp 3 (warps 0..2 do the hrow pass): the table lookup is issued a phase early
    auto fetch_row_tap = [&](int bb) { return uy[reflect101(ya - 16 + SKY_H * bb + lane, g.H)]; };
    auto publish_row_taps = [&](int bb, const ResizeTap &t) {
        SkyRowTap r;
        r.o0 = (t.i0 & (SKY_HRS - 1)) * SKY_HLD; r.o1 = (t.i1 & (SKY_HRS - 1)) * SKY_HLD; r.w0 = t.w0; r.w1 = t.w1;
        sm.trow[bb & 1][lane] = r;
        const int lo = __reduce_min_sync(0xffffffffu, t.i0), hi = __reduce_max_sync(0xffffffffu, t.i1);
        if (lane == 0) { sm.blk_lo[bb & 1] = lo; sm.blk_hi[bb & 1] = hi; }
    };

    // ---- prologue: column taps, first block's row taps, the segment's low-res patch
    for (int i = tid; i < SKY_UW; i += SKY_NT) sm.tcol[i] = ux[reflect101(x0 - SKY_R + i, g.W)];
    if (wid == 3) publish_row_taps(0, fetch_row_tap(0));
    int xlo, xhi, ylo, yhi;
    reflect_range(x0 - SKY_R, x0 + SKY_W + SKY_R - 1, g.W, xlo, xhi);
    reflect_range(ya - 16, ya - 16 + SKY_H * (bmax + 1) - 1, g.H, ylo, yhi);
    const int pc0 = ux[xlo].i0, pc1 = ux[xhi].i1, pr0 = uy[ylo].i0, pr1 = uy[yhi].i1; // taps are monotone
    {
        const bool patch_ok = pc1 - pc0 < SKY_LRC && pr1 - pr0 < SKY_LRR;
        if (patch_ok) {
            const int nrow = pr1 - pr0 + 1, ncol = pc1 - pc0 + 1;
            for (int i = tid; i < nrow * SKY_LRC; i += SKY_NT) {
                const int r = i >> 4, cc = i & 15;
                if (cc < ncol) sm.lr[i] = __ldg(L + (long long)(pr0 + r) * lw + pc0 + cc);
            }
        }
        if (tid == 0) { sm.patch[0] = patch_ok; sm.patch[1] = pr0; sm.patch[2] = pc0; }
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
        if (tid >= SKY_UW) { // padding columns 94, 95 (read by the up pass's 4-column items, never by the row filter)
            for (int r = first; r <= hi; r++) sm.hr[(r & (SKY_HRS - 1)) * SKY_HLD + tid] = 0.f;
        } else if (first <= hi) {
            const ResizeTap a = sm.tcol[tid];
            if (sm.patch[0]) {
                const float *p0 = sm.lr + (first - sm.patch[1]) * SKY_LRC - sm.patch[2];
                for (int r = first; r <= hi; r++, p0 += SKY_LRC)
                    sm.hr[(r & (SKY_HRS - 1)) * SKY_HLD + tid] = lerp2(p0[a.i0], a.w0, p0[a.i1], a.w1);
            } else {
                for (int r = first; r <= hi; r++) {
                    const float *row = med + blockIdx.z * strideLo + (long long)r * lw;
                    sm.hr[(r & (SKY_HRS - 1)) * SKY_HLD + tid] = lerp2(__ldg(row + a.i0), a.w0, __ldg(row + a.i1), a.w1);
                }
            }
        }
    };
    // vertical lerp -> up2 for the 16 row pairs x 24 groups of 4 columns of block bb: thread = row pair tid >> 3 (its two row
    // taps are read once) x column groups (tid & 7) + 8 k; products packed over column pairs, sums scalar (lerp2's order);
    // four LDS.128 from the hrow ring and two STS.128 per item.  (Columns 94, 95 are padding: hrow keeps them 0.)
    const int up_rp = tid >> 3, up_c = 4 * (tid & 7);
    auto up_pass = [&](int bb) {
        const SkyRowTap t0 = sm.trow[bb & 1][2 * up_rp], t1 = sm.trow[bb & 1][2 * up_rp + 1];
        const float *h00 = sm.hr + t0.o0 + up_c, *h01 = sm.hr + t0.o1 + up_c, *h10 = sm.hr + t1.o0 + up_c, *h11 = sm.hr + t1.o1 + up_c;
        float2 *dst = sm.up2 + up_rp * SKY_ULD + up_c;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const ulonglong2 a = *(const ulonglong2 *)(h00 + 32 * k), b = *(const ulonglong2 *)(h01 + 32 * k);
            const ulonglong2 c = *(const ulonglong2 *)(h10 + 32 * k), d = *(const ulonglong2 *)(h11 + 32 * k);
            float a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3, d0, d1, d2, d3;
            unpack2(mul2_bcast(t0.w0, a.x), a0, a1); unpack2(mul2_bcast(t0.w0, a.y), a2, a3);
            unpack2(mul2_bcast(t0.w1, b.x), b0, b1); unpack2(mul2_bcast(t0.w1, b.y), b2, b3);
            unpack2(mul2_bcast(t1.w0, c.x), c0, c1); unpack2(mul2_bcast(t1.w0, c.y), c2, c3);
            unpack2(mul2_bcast(t1.w1, d.x), d0, d1); unpack2(mul2_bcast(t1.w1, d.y), d2, d3);
            // (row 2rp, row 2rp+1) of columns c, c+1, then of c+2, c+3
            *(float4 *)(dst + 32 * k) = make_float4(__fadd_rn(a0, b0), __fadd_rn(c0, d0), __fadd_rn(a1, b1), __fadd_rn(c1, d1));
            *(float4 *)(dst + 32 * k + 2) = make_float4(__fadd_rn(a2, b2), __fadd_rn(c2, d2), __fadd_rn(a3, b3), __fadd_rn(c3, d3));
        }
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
    if (wid == 3 && bmax >= 1) publish_row_taps(1, fetch_row_tap(1));

    // row pass: this thread's row pair and column group.  A quarter-warp (8 lanes, one LDS.128 / STS.128 wavefront) = 2 neighbouring
    // column groups x 4 neighbouring row pairs; warp w makes rows 8 w .. 8 w + 7 of the block
    const int rp = 4 * wid + ((lane >> 1) & 3), gx = ((lane >> 3) << 1) | (lane & 1);
    const ulonglong2 *row_src = (const ulonglong2 *)(sm.up2 + rp * SKY_ULD + 8 * gx);
    // store offsets (floats) of the two 16-byte halves of this thread's 8 columns inside a ring row, swizzled
    const int st0 = 4 * ((2 * gx) ^ sky_swz(rp)), st1 = 4 * ((2 * gx + 1) ^ sky_swz(rp));

    double sum = 0.0, sq = 0.0;
    int slot = 0; // ring block that step b writes = b % SKY_NSLOT

    // Two barriers per step.  Between B and C: row filter of block b + hrow of block b+1.  After C: column pass of
    // output block b - SKY_LA, then (no barrier needed) up2 / row taps for the next step.
    for (int b = 0; b <= bmax; b++) {
        __syncthreads(); // B: up2(b), row taps of b+1 visible; everyone left the previous column pass
        ResizeTap tnext;
        const bool taps_due = wid == 3 && b + 2 <= bmax; // warp 3 has no hrow work
        if (taps_due) tnext = fetch_row_tap(b + 2); // lands behind the row filter
        {   // rows (2rp, 2rp+1), columns 8gx .. 8gx+7 of this step's row-filtered block; sliding window: tap j needs columns
            // j .. j + 7, so the window loads are interleaved with the taps (taps outer, the eight columns inner)
#if SKY_ROW_ROLLED
            // taps 0..2 peeled, then two passes of a 14-tap body (code size: the fully unrolled 31 x 8 form is 800 instructions,
            // and the whole step no longer fits the instruction caches).  Window entry i = column 8 gx + i.  The body sees
            // entries base .. base + 21 (base = 2 + 14 t): W[0..7] carried from the previous pass, W[8..21] loaded.
            f32x2_t cw[8];
            float o0[8], o1[8];
            {
                f32x2_t w[10];
#pragma unroll
                for (int q = 0; q < 5; q++) { ulonglong2 v = row_src[q]; w[2 * q] = v.x; w[2 * q + 1] = v.y; }
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        float u, v;
                        unpack2(mul2_bcast(c_gauss[j], w[e + j]), u, v);
                        if (j == 0) { o0[e] = u; o1[e] = v; }
                        else { o0[e] = __fadd_rn(o0[e], u); o1[e] = __fadd_rn(o1[e], v); }
                    }
#pragma unroll
                for (int i = 0; i < 8; i++) cw[i] = w[2 + i];
            }
#pragma unroll 1
            for (int t = 0; t < 2; t++) {
                const ulonglong2 *src = row_src + 5 + 7 * t;
                const float *kt = c_gauss + 3 + 14 * t;
                f32x2_t w[22];
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = cw[i];
#pragma unroll
                for (int jj = 0; jj < 14; jj++) {
                    if (!(jj & 1)) { ulonglong2 v = src[jj / 2]; w[8 + jj] = v.x; w[9 + jj] = v.y; }
                    const float k = kt[jj];
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        float u, v;
                        unpack2(mul2_bcast(k, w[1 + jj + e]), u, v);
                        o0[e] = __fadd_rn(o0[e], u); o1[e] = __fadd_rn(o1[e], v);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; i++) cw[i] = w[14 + i];
            }
#else
            f32x2_t w[38];
#pragma unroll
            for (int q = 0; q < 4; q++) { ulonglong2 v = row_src[q]; w[2 * q] = v.x; w[2 * q + 1] = v.y; }
            float o0[8], o1[8];
#pragma unroll
            for (int j = 0; j < 31; j++) {
                if (j & 1) { ulonglong2 v = row_src[(j + 7) / 2]; w[j + 7] = v.x; w[j + 8] = v.y; }
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    float u, v;
                    unpack2(mul2_bcast(c_gauss[j], w[e + j]), u, v);
                    if (j == 0) { o0[e] = u; o1[e] = v; }
                    else { o0[e] = __fadd_rn(o0[e], u); o1[e] = __fadd_rn(o1[e], v); }
                }
            }
#endif
            float *d0 = sm.rf + (slot * SKY_H + 2 * rp) * SKY_W;
            *(float4 *)(d0 + st0) = make_float4(o0[0], o0[1], o0[2], o0[3]);
            *(float4 *)(d0 + st1) = make_float4(o0[4], o0[5], o0[6], o0[7]);
            *(float4 *)(d0 + SKY_W + st0) = make_float4(o1[0], o1[1], o1[2], o1[3]);
            *(float4 *)(d0 + SKY_W + st1) = make_float4(o1[4], o1[5], o1[6], o1[7]);
            if (slot < SKY_LA) { // mirror of the first blocks behind the ring: 38-row windows never wrap
                float *d1 = d0 + SKY_RING * SKY_W;
                *(float4 *)(d1 + st0) = make_float4(o0[0], o0[1], o0[2], o0[3]);
                *(float4 *)(d1 + st1) = make_float4(o0[4], o0[5], o0[6], o0[7]);
                *(float4 *)(d1 + SKY_W + st0) = make_float4(o1[0], o1[1], o1[2], o1[3]);
                *(float4 *)(d1 + SKY_W + st1) = make_float4(o1[4], o1[5], o1[6], o1[7]);
            }
        }
        if (b < bmax) hrow_pass(b + 1);
        float2 gv[8];
        const int yblk = ya + SKY_H * (b - SKY_LA);
        const int mode = b >= SKY_LA ? sm.mode[b - SKY_LA] : 3;
        if (mode < 2) sky_load_gray<0>(G, x0, yblk, tid, g, gv);
        else if (mode == 2) sky_load_gray<2>(G, x0, yblk, tid, g, gv);
        __syncthreads(); // C: row-filtered block b and the hrow rows of b+1 are visible
        if (b >= SKY_LA) {
            const int ring_row0 = (slot >= SKY_LA ? slot - SKY_LA : slot + SKY_NSLOT - SKY_LA) * SKY_H; // block b - SKY_LA
            if (mode == 0) sky_column_pass<0>(sm.rf, gv, ring_row0, x0, yblk, tid, S, SM, nseg, g, sum, sq);
            else if (mode == 1) sky_column_pass<1>(sm.rf, gv, ring_row0, x0, yblk, tid, S, SM, nseg, g, sum, sq);
            else sky_column_pass<2>(sm.rf, gv, ring_row0, x0, yblk, tid, S, SM, nseg, g, sum, sq);
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
        d.ncand = 0; d.nroots = 0; d.ncomp = 0; d.nstars = 0; d.overflow = 0; d.nbig = 0; d.nbig2 = 0; d.nbig3 = 0;
        ds[f] = d;
    }
}

}  // namespace oss
