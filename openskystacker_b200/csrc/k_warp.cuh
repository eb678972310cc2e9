// Kernel group (4): inverse-map bilinear affine warp fused with the running float32 sum.
//
// Replaces (reference):
//   cv::warpAffine(target, M, INTER_LINEAR + WARP_INVERSE_MAP)   libstacker/src/util.cpp:579
//   cv::add(workingImage, targetAligned)                          libstacker/src/util.cpp:749
//   workingImage /= totalValidImages                              libstacker/src/imagestacker.cpp:361
// Source coordinates are OpenCV's fixed-point ones (10 fractional bits, rounded to 1/32 px,
// cvRound = round-half-even), the blend is fp32 ((a*w00 + b*w01) + c*w10) + d*w11 without
// FMA, taps outside the image read 0 (BORDER_CONSTANT) — SURVEY.md a-14 / Appendix B.7.
#pragma once
#include <cuda.h> // CUtensorMap

#include "common.cuh"

namespace oss {

// ---------------------------------------------------------------------------------------
// K8a  per-frame coordinate tables.  OpenCV's fixed-point source coordinate is separable:
// X = (X0[y] + adelta[x]) >> 5, Y = (Y0[y] + bdelta[x]) >> 5 with adelta/bdelta = cvRound(M*x*1024)
// and X0/Y0 = cvRound((M*y + t)*1024) + 16.  Building the four tables once per frame (W + H
// entries) takes every fp64 operation out of the per-pixel kernel.
__global__ void __launch_bounds__(256)
k_warp_tables(const AlignOut *__restrict__ align, int2 *__restrict__ ab, long long strideAB, int2 *__restrict__ xy,
              long long strideXY, int W, int H)
{
    const int f = blockIdx.y;
    if (align[f].err != 0) return;
    const int i = blockIdx.x * 256 + threadIdx.x;
    double M[6];
#pragma unroll
    for (int q = 0; q < 6; q++) M[q] = (double)align[f].M[q];
    if (i < W) {
        int a = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)i), 1024.0));
        int b = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)i), 1024.0));
        ab[f * strideAB + i] = make_int2(a, b);
    } else if (i < W + H) {
        const int y = i - W;
        int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
        int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
        xy[f * strideXY + y] = make_int2(X0, Y0);
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// K8b  warp + accumulate, tiled.  One CTA owns a 64x32 tile of the accumulator for the whole
// batch: the 24 (RGB) accumulator values of each thread stay in registers across the frames
// (ONE read-modify-write of the sum per batch), and for every frame the tile's source
// footprint (a few rows/columns more than the tile for the small rigid motions of a stack)
// is staged in shared memory with 16-byte cp.async, so the 4 x C bilinear taps per pixel are
// conflict-free LDS instead of 3-4 L1 wavefronts each.  Footprints that do not fit (strong
// rotation / scale) fall back to direct gathers for that frame and tile.
#define WT_W 64
#define WT_H 32
#define WT_ROWS 36
#define WT_COLS 72 // pixels per staged row (multiple of 4)

struct WarpFoot { int cx0, cx1, cy0, cy1; bool empty, fits, interior; };

// Source footprint of the tile [x0..xe] x [y0..ye]: the fixed-point coordinate is a separable sum of two
// monotone tables, so its extremes over the tile sit at the tile's corners.
__device__ __forceinline__ WarpFoot warp_footprint(const int2 *__restrict__ AB, const int2 *__restrict__ XY, int x0, int xe, int y0,
                                                   int ye, int W, int H)
{
    const int2 a0 = AB[x0], a1 = AB[xe], b0 = XY[y0], b1 = XY[ye];
    const int Xlo = min(a0.x, a1.x) + min(b0.x, b1.x), Xhi = max(a0.x, a1.x) + max(b0.x, b1.x);
    const int Ylo = min(a0.y, a1.y) + min(b0.y, b1.y), Yhi = max(a0.y, a1.y) + max(b0.y, b1.y);
    const int sx_lo = Xlo >> 10, sx_hi = (Xhi >> 10) + 1, sy_lo = Ylo >> 10, sy_hi = (Yhi >> 10) + 1;
    WarpFoot t;
    t.cx0 = max(sx_lo, 0); t.cx1 = min(sx_hi, W - 1); t.cy0 = max(sy_lo, 0); t.cy1 = min(sy_hi, H - 1);
    t.empty = t.cx0 > t.cx1 || t.cy0 > t.cy1; // the whole tile maps outside the frame
    // every staged row starts at a pixel index that is a multiple of 4 (16-byte aligned for C = 1 and 3)
    t.fits = !t.empty && (t.cy1 - t.cy0 + 1) <= WT_ROWS && (t.cx1 - t.cx0 + 1 + 3) <= WT_COLS;
    t.interior = sx_lo >= 0 && sx_hi <= W - 1 && sy_lo >= 0 && sy_hi <= H - 1;
    return t;
}

template <int C>
__global__ void __launch_bounds__(256, 3)
k_warp_accumulate_tiled(const float *__restrict__ cal, long long strideCal, const AlignOut *__restrict__ align, int nframes,
                        const int2 *__restrict__ ab, long long strideAB, const int2 *__restrict__ xy, long long strideXY,
                        float *__restrict__ sum, int W, int H, int ytile0, int zero_init)
{
    constexpr int ROWF = WT_COLS * C;           // floats per staged row
    constexpr int CHUNKS = ROWF / 4;            // 16-byte chunks per staged row
    extern __shared__ __align__(16) unsigned char wsm[];
    float *tiles = (float *)wsm;                                  // [2][WT_ROWS * ROWF]
    int2 *sxy = (int2 *)(tiles + 2 * WT_ROWS * ROWF);             // [2][WT_H]
    int *valid = (int *)(sxy + 2 * WT_H);                         // [OSS_MAX_BATCH + 1]
    __shared__ WarpFoot sfoot[OSS_MAX_BATCH];
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31, wid = tid >> 5;
    const int x0 = blockIdx.x * WT_W, y0 = (blockIdx.y + ytile0) * WT_H; // ytile0: first tile row of a banded launch
    const int x = x0 + tx;
    const int xe = min(x0 + WT_W, W) - 1, ye = min(y0 + WT_H, H) - 1; // last pixel of the (possibly clipped) tile
    const bool col_ok = x < W;
    if (tid == 0) { // frames that aligned, in order (util.cpp:746-747 skips the others)
        int n = 0;
        for (int f = 0; f < nframes; f++)
            if (align[f].err == 0) valid[1 + n++] = f;
        valid[0] = n;
    }
    float acc[8][C];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int y = y0 + ty + 4 * k;
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[k][ch] = (col_ok && y < H && !zero_init) ? sum[((long long)y * W + x) * C + ch] : 0.f; // zero_init: the stack's first launch (no memset, no read)
    }
    __syncthreads();
    const int nvalid = valid[0];
    // footprints of all frames up front (one exposed table-load latency per CTA instead of one per frame)
    if (tid < nvalid) {
        const int f = valid[1 + tid];
        sfoot[tid] = warp_footprint(ab + f * strideAB, xy + f * strideXY, x0, xe, y0, ye, W, H);
    }
    __syncthreads();

    // stage valid frame v into buffer `buf` (asynchronously): warp w copies rows w, w+8, ...
    auto stage = [&](int v, int buf) {
        const int f = valid[1 + v];
        if (tid < WT_H) { // this tile's (X0, Y0) rows, 8 bytes each
            unsigned sa = (unsigned)__cvta_generic_to_shared(sxy + buf * WT_H + tid);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(xy + f * strideXY + min(y0 + tid, H - 1)));
        }
        const WarpFoot t = sfoot[v];
        if (t.fits) {
            const float *src = cal + f * strideCal;
            float *dst = tiles + buf * (WT_ROWS * ROWF);
            const int nrows = t.cy1 - t.cy0 + 1;
            for (int r = wid; r < nrows; r += 8) {
                const float *g = src + (((t.cy0 + r) * W + t.cx0) & ~3) * C;
                float *d = dst + r * ROWF;
#pragma unroll
                for (int q = lane; q < CHUNKS; q += 32) cp_async16(d + 4 * q, g + 4 * q);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int2 abx_next = make_int2(0, 0);
    if (nvalid > 0) {
        stage(0, 0);
        if (col_ok) abx_next = ab[valid[1] * strideAB + x];
    }
    for (int v = 0; v < nvalid; v++) {
        const int f = valid[1 + v], buf = v & 1;
        const int2 abx = abx_next;
        cp_async_wait_all();
        __syncthreads(); // buffer `buf` has landed; everyone is done reading buffer `buf ^ 1`
        if (v + 1 < nvalid) {
            stage(v + 1, buf ^ 1);
            if (col_ok) abx_next = ab[valid[2 + v] * strideAB + x]; // consumed one frame later
        }
        if (!col_ok) continue;
        const WarpFoot t = sfoot[v];
        if (t.empty) { // warpAffine writes 0 there
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
                for (int ch = 0; ch < C; ch++) acc[k][ch] = __fadd_rn(acc[k][ch], 0.f);
            continue;
        }
        const float *tile = tiles + buf * (WT_ROWS * ROWF);
        const int2 *rowxy = sxy + buf * WT_H;
        if (t.fits && t.interior && (W & 3) == 0 && y0 + WT_H <= H) {
            // fastest path (the usual one): every tap is inside the frame and inside the staged tile, all 8 rows of the
            // thread exist, and with W % 4 == 0 every staged row starts at the same phase (cx0 & 3) of its 16-byte chunk,
            // so one base pointer serves both tap rows: tap(sy, sx) = tb + sy * ROWF + sx * C
            const float *tb = tile + ((t.cx0 & 3) - t.cx0) * C - t.cy0 * ROWF;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int2 b = rowxy[ty + 4 * k];
                const int X = (b.x + abx.x) >> 5, Y = (b.y + abx.y) >> 5;
                const float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
                const float ix = 1.f - fx, iy = 1.f - fy;
                const float w00 = iy * ix, w01 = iy * fx, w10 = fy * ix, w11 = fy * fx; // exact: multiples of 1/1024
                const float *p0 = tb + (Y >> 5) * ROWF + (X >> 5) * C;
                const float *p1 = p0 + ROWF;
#pragma unroll
                for (int ch = 0; ch < C; ch++) {
                    float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0[ch], w00), __fmul_rn(p0[C + ch], w01)), __fmul_rn(p1[ch], w10)),
                                          __fmul_rn(p1[C + ch], w11));
                    acc[k][ch] = __fadd_rn(acc[k][ch], val);
                }
            }
            continue;
        }
        if (t.fits && t.interior) {
            // fast path: every tap is inside the frame and inside the staged tile
            const int w4 = W & 3, cxa = t.cx0 & 3;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int yy = ty + 4 * k;
                if (y0 + yy >= H) continue;
                const int2 b = rowxy[yy];
                const int X = (b.x + abx.x) >> 5, Y = (b.y + abx.y) >> 5;
                const int sx = X >> 5, sy = Y >> 5;
                const float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
                const float ix = 1.f - fx, iy = 1.f - fy;
                const float w00 = iy * ix, w01 = iy * fx, w10 = fy * ix, w11 = fy * fx; // exact: multiples of 1/1024
                // staged row r starts at linear pixel index ((cy0 + r) * W + cx0) & ~3
                const int r = sy - t.cy0, rw = sy * w4 + cxa; // (sy * W + cx0) mod 4 == (sy * (W mod 4) + cx0 mod 4) mod 4
                const float *p0 = tile + r * ROWF + (sx - t.cx0 + (rw & 3)) * C;
                const float *p1 = tile + (r + 1) * ROWF + (sx - t.cx0 + ((rw + w4) & 3)) * C;
#pragma unroll
                for (int ch = 0; ch < C; ch++) {
                    float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0[ch], w00), __fmul_rn(p0[C + ch], w01)), __fmul_rn(p1[ch], w10)),
                                          __fmul_rn(p1[C + ch], w11));
                    acc[k][ch] = __fadd_rn(acc[k][ch], val);
                }
            }
            continue;
        }
        const float *src = cal + f * strideCal;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int yy = ty + 4 * k;
            if (y0 + yy >= H) continue;
            const int2 b = rowxy[yy];
            const int X = (b.x + abx.x) >> 5, Y = (b.y + abx.y) >> 5;
            const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
            const float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
            const float ix = 1.f - fx, iy = 1.f - fy;
            const float w00 = iy * ix, w01 = iy * fx, w10 = fy * ix, w11 = fy * fx;
            const bool xa = sx >= 0 && sx < W, xb = sx + 1 >= 0 && sx + 1 < W;
            const bool ya = sy >= 0 && sy < H, yb = sy + 1 >= 0 && sy + 1 < H;
            float s00[C], s01[C], s10[C], s11[C];
            if (t.fits) {
                const int r = sy - t.cy0;
                const int o0 = r * ROWF + (sx - t.cx0 + ((sy * W + t.cx0) & 3)) * C;
                const int o1 = (r + 1) * ROWF + (sx - t.cx0 + (((sy + 1) * W + t.cx0) & 3)) * C;
#pragma unroll
                for (int ch = 0; ch < C; ch++) {
                    s00[ch] = (xa && ya) ? tile[o0 + ch] : 0.f;
                    s01[ch] = (xb && ya) ? tile[o0 + C + ch] : 0.f;
                    s10[ch] = (xa && yb) ? tile[o1 + ch] : 0.f;
                    s11[ch] = (xb && yb) ? tile[o1 + C + ch] : 0.f;
                }
            } else {
                const float *p = src + ((long long)sy * W + sx) * C;
#pragma unroll
                for (int ch = 0; ch < C; ch++) {
                    s00[ch] = (xa && ya) ? __ldg(p + ch) : 0.f;
                    s01[ch] = (xb && ya) ? __ldg(p + C + ch) : 0.f;
                    s10[ch] = (xa && yb) ? __ldg(p + (long long)W * C + ch) : 0.f;
                    s11[ch] = (xb && yb) ? __ldg(p + (long long)W * C + C + ch) : 0.f;
                }
            }
#pragma unroll
            for (int ch = 0; ch < C; ch++) {
                float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00[ch], w00), __fmul_rn(s01[ch], w01)), __fmul_rn(s10[ch], w10)),
                                      __fmul_rn(s11[ch], w11));
                acc[k][ch] = __fadd_rn(acc[k][ch], val);
            }
        }
    }
    if (!col_ok) return;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int y = y0 + ty + 4 * k;
        if (y >= H) continue;
#pragma unroll
        for (int ch = 0; ch < C; ch++) sum[((long long)y * W + x) * C + ch] = acc[k][ch];
    }
}

template <int C> constexpr int warp_tiled_smem() { return 2 * WT_ROWS * WT_COLS * C * 4 + 2 * WT_H * 8 + (OSS_MAX_BATCH + 1) * 4; }

// plain copy (keeps the calibrated reference frame, imagestacker.cpp:286)
__global__ void __launch_bounds__(256)
k_copy_f32(const float *__restrict__ src, float *__restrict__ dst, long long count)
{
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 + 4 <= count) ((float4 *)(dst + i0))[0] = ((const float4 *)(src + i0))[0];
    else for (long long i = i0; i < count; i++) dst[i] = src[i];
}

// result = (ref + sum) * r: `sum` holds the lights only, like processConcurrent's partial
// (util.cpp:736,749); the reference frame joins at the end like workingImage += partial
// (imagestacker.cpp:287,349), then Mat /= n  ==  x * float(1.0/n) (imagestacker.cpp:361).
__global__ void __launch_bounds__(256)
k_finalize(const float *__restrict__ sum, const float *__restrict__ ref, float *__restrict__ out, long long count,
           const int *__restrict__ valid_lights)
{
    // totalValidImages is read where it lives (no host round trip between the last warp and this kernel); the reference frame
    // counts as one (imagestacker.cpp:304).  Fewer than 2: the host reports the reference's error, whatever lands in `out`.
    const float r = (float)(1.0 / (double)(*valid_lights + (ref ? 1 : 0)));
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 + 4 <= count) {
        float4 q = ((const float4 *)(sum + i0))[0];
        if (ref) {
            float4 p = ((const float4 *)(ref + i0))[0];
            q.x = __fadd_rn(p.x, q.x); q.y = __fadd_rn(p.y, q.y); q.z = __fadd_rn(p.z, q.z); q.w = __fadd_rn(p.w, q.w);
        }
        q.x = __fmul_rn(q.x, r); q.y = __fmul_rn(q.y, r); q.z = __fmul_rn(q.z, r); q.w = __fmul_rn(q.w, r);
        ((float4 *)(out + i0))[0] = q;
    } else {
        for (long long i = i0; i < count; i++) out[i] = __fmul_rn(ref ? __fadd_rn(ref[i], sum[i]) : sum[i], r);
    }
}


// =======================================================================================
// K8c  warp + accumulate with TMA-staged tiles (the default for geometries whose rows are 16-byte multiples).
//
// Same arithmetic and the same 64x32 accumulator tile per CTA as k_warp_accumulate_tiled, but the source footprint of
// every frame is fetched by ONE `cp.async.bulk.tensor` (3-D tensor map over [slot][row][column*channel] of the calibrated
// planes, box = WT_ROWS x WT_COLS pixels) issued by one thread and completed on an mbarrier:
//   - out-of-bounds parts of the box arrive as zeros, which IS cv::warpAffine's BORDER_CONSTANT 0: border tiles take the
//     same fast path as interior ones (no per-tap bounds tests, no clamped footprint);
//   - every staged row starts at the same pixel (the footprint's first column rounded down to 4 pixels, the 16-byte
//     granularity of TMA coordinates): one base pointer serves all taps;
//   - no `__syncthreads` in the frame loop: consumers wait on the "full" mbarrier of their stage, each warp arrives on the
//     stage's "empty" mbarrier when it has read its taps, and only the producer thread waits for that before it overwrites
//     the stage with the frame after next — warps drift apart by up to a frame instead of meeting at a barrier per frame
//     (the largest stall of the cp.async version).
// The per-row (X0, Y0) coordinate tables of all frames of the launch are staged once in the prologue.
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

struct WarpFootT { int sx0, sy0; bool empty, fits; };

template <int C>
struct WarpTmaSmem {
    float tiles[2][WT_ROWS * WT_COLS * C]; // TMA destinations: 128-byte aligned (WT_ROWS * WT_COLS * C * 4 is a multiple of 128)
    int2 rowxy[OSS_MAX_BATCH][WT_H];
    WarpFootT foot[OSS_MAX_BATCH];
    int valid[OSS_MAX_BATCH + 1];
    unsigned long long full[2], empty[2];
};

template <int C>
__global__ void __launch_bounds__(256, 3)
k_warp_accumulate_tma(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ cal, long long strideCal,
                      const AlignOut *__restrict__ align, int nframes, const int2 *__restrict__ ab, long long strideAB,
                      const int2 *__restrict__ xy, long long strideXY, float *__restrict__ sum, int W, int H, int ytile0, int zero_init)
{
    constexpr int ROWF = WT_COLS * C;
    constexpr unsigned TILE_BYTES = WT_ROWS * ROWF * 4;
    static_assert(TILE_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
    extern __shared__ __align__(16) unsigned char wsm_tma[];
    // TMA destinations need 128-byte alignment: round the dynamic window up (the launch asks for 128 bytes more)
    WarpTmaSmem<C> &sm = *(WarpTmaSmem<C> *)(wsm_tma + ((128u - (smem_u32(wsm_tma) & 127u)) & 127u));
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31;
    const int x0 = blockIdx.x * WT_W, y0 = (blockIdx.y + ytile0) * WT_H;
    const int x = x0 + tx;
    const int xe = min(x0 + WT_W, W) - 1, ye = min(y0 + WT_H, H) - 1;
    const bool col_ok = x < W;
    if (tid == 0) { // frames that aligned, in order (util.cpp:746-747 skips the others)
        int n = 0;
        for (int f = 0; f < nframes; f++)
            if (align[f].err == 0) sm.valid[1 + n++] = f;
        sm.valid[0] = n;
        mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1);
        mbar_init(&sm.empty[0], 8); mbar_init(&sm.empty[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float acc[8][C];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int y = y0 + ty + 4 * k;
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[k][ch] = (col_ok && y < H && !zero_init) ? sum[((long long)y * W + x) * C + ch] : 0.f; // zero_init: the stack's first launch (no memset, no read)
    }
    __syncthreads();
    const int nvalid = sm.valid[0];
    // footprints (unclamped: what falls outside the frame arrives as zeros) and the row tables of every frame
    if (tid < nvalid) {
        const int f = sm.valid[1 + tid];
        const int2 *AB = ab + f * strideAB, *XY = xy + f * strideXY;
        const int2 a0 = AB[x0], a1 = AB[xe], b0 = XY[y0], b1 = XY[ye];
        const int Xlo = min(a0.x, a1.x) + min(b0.x, b1.x), Xhi = max(a0.x, a1.x) + max(b0.x, b1.x);
        const int Ylo = min(a0.y, a1.y) + min(b0.y, b1.y), Yhi = max(a0.y, a1.y) + max(b0.y, b1.y);
        const int sx_lo = Xlo >> 10, sx_hi = (Xhi >> 10) + 1, sy_lo = Ylo >> 10, sy_hi = (Yhi >> 10) + 1;
        WarpFootT t;
        t.sx0 = sx_lo & ~3; // the box must start at a 16-byte multiple of the row: 4 pixels (4 or 12 floats); floors negatives too
        t.sy0 = sy_lo;
        t.empty = sx_hi < 0 || sx_lo > W - 1 || sy_hi < 0 || sy_lo > H - 1; // the whole tile maps outside the frame
        t.fits = (sy_hi - sy_lo + 1) <= WT_ROWS && (sx_hi - t.sx0 + 1) <= WT_COLS && sx_lo > -(1 << 20) && sx_lo < (1 << 20) &&
                 sy_lo > -(1 << 20) && sy_lo < (1 << 20);
        sm.foot[tid] = t;
    }
    for (int i = tid; i < nvalid * WT_H; i += 256) {
        const int v = i / WT_H, r = i - v * WT_H;
        sm.rowxy[v][r] = xy[sm.valid[1 + v] * strideXY + min(y0 + r, H - 1)];
    }
    __syncthreads();

    // producer (thread 0): frame v -> stage v & 1
    auto issue = [&](int v) {
        const WarpFootT t = sm.foot[v];
        const int s = v & 1;
        if (t.fits && !t.empty) {
            mbar_arrive_expect_tx(&sm.full[s], TILE_BYTES);
            tma_load_3d(sm.tiles[s], &tmap, t.sx0 * C, t.sy0, sm.valid[1 + v], &sm.full[s]);
        } else {
            mbar_arrive(&sm.full[s]); // nothing to stage: the phase still completes
        }
    };
    if (tid == 0 && nvalid > 0) issue(0);
    int2 abx_next = make_int2(0, 0);
    if (nvalid > 0 && col_ok) abx_next = ab[sm.valid[1] * strideAB + x];
    for (int v = 0; v < nvalid; v++) {
        const int f = sm.valid[1 + v], s = v & 1;
        const int2 abx = abx_next;
        if (v + 1 < nvalid) {
            if (tid == 0) {
                // stage (v + 1) & 1 was last read for frame v - 1: wait until all 8 warps have arrived for it
                if (v >= 1) mbar_wait(&sm.empty[(v + 1) & 1], ((v - 1) >> 1) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(v + 1);
            }
            if (col_ok) abx_next = ab[sm.valid[2 + v] * strideAB + x]; // consumed one frame later
        }
        mbar_wait(&sm.full[s], (v >> 1) & 1);
        const WarpFootT t = sm.foot[v];
        const int2 *rowxy = sm.rowxy[v];
        if (col_ok) {
            if (t.empty) { // warpAffine writes 0 there
#pragma unroll
                for (int k = 0; k < 8; k++)
#pragma unroll
                    for (int ch = 0; ch < C; ch++) acc[k][ch] = __fadd_rn(acc[k][ch], 0.f);
            } else if (t.fits) {
                // every tap is inside the staged box; taps outside the frame were filled with zeros by the TMA unit.
                // The four products of a channel are two packed multiplies (mul.rn.f32x2: each half an IEEE fp32 product, no
                // contraction with the scalar adds that follow): the kernel is issue-bound, this takes 6 instructions per pixel out
                const float *tb = sm.tiles[s] - t.sx0 * C - t.sy0 * ROWF;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int yy = ty + 4 * k;
                    if (y0 + yy >= H) continue;
                    const int2 b = rowxy[yy];
                    const int X = (b.x + abx.x) >> 5, Y = (b.y + abx.y) >> 5;
                    const float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
                    const float ix = 1.f - fx, iy = 1.f - fy;
                    const unsigned long long wtop = pack_f32x2(iy * ix, iy * fx), wbot = pack_f32x2(fy * ix, fy * fx); // exact: multiples of 1/1024
                    const float *p0 = tb + (Y >> 5) * ROWF + (X >> 5) * C;
                    const float *p1 = p0 + ROWF;
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        float m00, m01, m10, m11;
                        unpack_f32x2(mul_f32x2(pack_f32x2(p0[ch], p0[C + ch]), wtop), m00, m01);
                        unpack_f32x2(mul_f32x2(pack_f32x2(p1[ch], p1[C + ch]), wbot), m10, m11);
                        const float val = __fadd_rn(__fadd_rn(__fadd_rn(m00, m01), m10), m11);
                        acc[k][ch] = __fadd_rn(acc[k][ch], val);
                    }
                }
            } else { // footprint larger than the box (strong rotation / scale): direct gathers with bounds tests
                const float *src = cal + f * strideCal;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int yy = ty + 4 * k;
                    if (y0 + yy >= H) continue;
                    const int2 b = rowxy[yy];
                    const int X = (b.x + abx.x) >> 5, Y = (b.y + abx.y) >> 5;
                    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
                    const float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
                    const float ix = 1.f - fx, iy = 1.f - fy;
                    const float w00 = iy * ix, w01 = iy * fx, w10 = fy * ix, w11 = fy * fx;
                    const bool xa = sx >= 0 && sx < W, xb = sx + 1 >= 0 && sx + 1 < W;
                    const bool ya = sy >= 0 && sy < H, yb = sy + 1 >= 0 && sy + 1 < H;
                    const float *p = src + ((long long)sy * W + sx) * C;
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        const float s00 = (xa && ya) ? __ldg(p + ch) : 0.f, s01 = (xb && ya) ? __ldg(p + C + ch) : 0.f;
                        const float s10 = (xa && yb) ? __ldg(p + (long long)W * C + ch) : 0.f;
                        const float s11 = (xb && yb) ? __ldg(p + (long long)W * C + C + ch) : 0.f;
                        float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, w00), __fmul_rn(s01, w01)), __fmul_rn(s10, w10)), __fmul_rn(s11, w11));
                        acc[k][ch] = __fadd_rn(acc[k][ch], val);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]); // this warp has read its taps of stage s
    }
    if (!col_ok) return;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int y = y0 + ty + 4 * k;
        if (y >= H) continue;
#pragma unroll
        for (int ch = 0; ch < C; ch++) sum[((long long)y * W + x) * C + ch] = acc[k][ch];
    }
}

}  // namespace oss
