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
#include "common.cuh"

namespace oss {

struct WarpCoord { int sx, sy; float w00, w01, w10, w11; };

__device__ __forceinline__ WarpCoord warp_coord(const double *M, int x, int y)
{
    int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
    int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
    int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
    int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
    int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    WarpCoord c;
    c.sx = min(max(X >> 5, -32768), 32767);
    c.sy = min(max(Y >> 5, -32768), 32767);
    float fx = (float)(X & 31) * (1.f / 32), fy = (float)(Y & 31) * (1.f / 32);
    float ix = 1.f - fx, iy = 1.f - fy; // exact: multiples of 1/32
    c.w00 = iy * ix; c.w01 = iy * fx; c.w10 = fy * ix; c.w11 = fy * fx; // exact: multiples of 1/1024
    return c;
}

template <int C>
__device__ __forceinline__ void warp_sample(const float *__restrict__ src, int W, int H, const WarpCoord &c, float *out)
{
    const bool x0 = c.sx >= 0 && c.sx < W, x1 = c.sx + 1 >= 0 && c.sx + 1 < W;
    const bool y0 = c.sy >= 0 && c.sy < H, y1 = c.sy + 1 >= 0 && c.sy + 1 < H;
    const float *p = src + ((long long)c.sy * W + c.sx) * C;
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
        float s00 = (x0 && y0) ? __ldg(p + ch) : 0.f;
        float s01 = (x1 && y0) ? __ldg(p + C + ch) : 0.f;
        float s10 = (x0 && y1) ? __ldg(p + (long long)W * C + ch) : 0.f;
        float s11 = (x1 && y1) ? __ldg(p + (long long)W * C + C + ch) : 0.f;
        out[ch] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s00, c.w00), __fmul_rn(s01, c.w01)), __fmul_rn(s10, c.w10)),
                            __fmul_rn(s11, c.w11));
    }
}

// K8  warp + accumulate for a batch of frames: sum(x,y) += warp_f(x,y) for every aligned
// frame f, in frame order, with ONE read-modify-write of the accumulator per batch.
// blockDim = (32, 8); one pixel per thread.
template <int C>
__global__ void __launch_bounds__(256)
k_warp_accumulate(const float *__restrict__ cal, long long strideCal, const AlignOut *__restrict__ align, int nframes,
                  float *__restrict__ sum, int W, int H)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    float *dst = sum + ((long long)y * W + x) * C;
    float acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) acc[ch] = dst[ch];
    for (int f = 0; f < nframes; f++) {
        if (align[f].err != 0) continue; // util.cpp:746-747
        double M[6];
#pragma unroll
        for (int i = 0; i < 6; i++) M[i] = (double)align[f].M[i];
        WarpCoord c = warp_coord(M, x, y);
        float v[C];
        warp_sample<C>(cal + f * strideCal, W, H, c, v);
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[ch] = __fadd_rn(acc[ch], v[ch]);
    }
#pragma unroll
    for (int ch = 0; ch < C; ch++) dst[ch] = acc[ch];
}

// plain warp (test door for oss_warp_affine)
template <int C>
__global__ void __launch_bounds__(256)
k_warp_only(const float *__restrict__ src, float *__restrict__ dst, const float *__restrict__ Mf, int W, int H)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    double M[6];
#pragma unroll
    for (int i = 0; i < 6; i++) M[i] = (double)Mf[i];
    WarpCoord c = warp_coord(M, x, y);
    float v[C];
    warp_sample<C>(src, W, H, c, v);
#pragma unroll
    for (int ch = 0; ch < C; ch++) dst[((long long)y * W + x) * C + ch] = v[ch];
}

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
k_finalize(const float *__restrict__ sum, const float *__restrict__ ref, float *__restrict__ out, long long count, float r)
{
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

}  // namespace oss
