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
    // 13 rounds of min-extraction, fully unrolled so v[] stays in registers
#pragma unroll
    for (int i = 0; i <= 12; i++) {
#pragma unroll
        for (int j = i + 1; j < 25; j++) {
            float lo_ = fminf(v[i], v[j]), hi_ = fmaxf(v[i], v[j]);
            v[i] = lo_; v[j] = hi_;
        }
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
// K3  sky_sub_stats.  One CTA produces a 64 x 128 tile of stars = gray - sky:
//   phase 1  upsample the median grid over the tile + 15 px apron (horizontal lerp of the few
//            low-res rows the tile touches, then vertical lerp), with the Gaussian's
//            BORDER_REFLECT_101 applied to the FULL-RES coordinates;
//   phase 2  31-tap row filter, taps accumulated k = 0..30 in order; each thread makes 4
//            neighbouring outputs from one 36-float register window (9 LDS.128);
//   phase 3  column filter as centre + 15 symmetric pairs, 4 vertically adjacent outputs per
//            thread from a 34-float window; subtract from gray; write; fp64 sum / sum of
//            squares over the statistics ROI; per-32-px-segment maximum (lets the
//            thresholding pass skip empty segments).
// HBM traffic: gray read once (4 B/px), stars written once (4 B/px).  About 120 non-fusable
// fp32 operations per pixel make this kernel FP32-pipe bound, not HBM bound; the register
// windows keep shared-memory traffic (22 words/px) below the FP32 time.
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

#define SKY_TX 64
#define SKY_TY 128
#define SKY_R 15
#define SKY_AX (SKY_TX + 2 * SKY_R) // 94 columns of apron
#define SKY_AY (SKY_TY + 2 * SKY_R) // 158 rows of apron
#define SKY_UP_LD 100               // 36-float windows starting at 4*gx stay inside the row
#define SKY_HROWS 40                // low-res rows a tile may touch (158/8 + reflections)
constexpr int kSkySmemBytes = (SKY_AY * SKY_UP_LD + SKY_AY * SKY_TX) * 4 + (SKY_AX + SKY_AY) * (int)sizeof(ResizeTap);

__global__ void __launch_bounds__(256, 2)
k_sky_sub_stats(const float *__restrict__ gray, long long strideGray, const float *__restrict__ med,
                long long strideLo, int lw, const ResizeTap *__restrict__ ux, const ResizeTap *__restrict__ uy,
                float *__restrict__ stars, long long strideStars, float *__restrict__ segmax, long long strideSeg,
                int nseg, double *__restrict__ partials, long long stridePart, Geom g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *up = (float *)smem_raw;                  // [SKY_AY][SKY_UP_LD]
    float *rf = up + SKY_AY * SKY_UP_LD;            // [SKY_AY][SKY_TX]
    float *hrow = rf;                               // [SKY_HROWS][SKY_UP_LD], dead before phase 2 writes rf
    ResizeTap *tcol = (ResizeTap *)(rf + SKY_AY * SKY_TX); // [SKY_AX]
    ResizeTap *trow = tcol + SKY_AX;                // [SKY_AY]
    __shared__ int lr_min, lr_max;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * SKY_TX, y0 = blockIdx.y * SKY_TY, f = blockIdx.z;
    const float *L = med + f * strideLo;

    if (tid == 0) { lr_min = 1 << 30; lr_max = -1; }
    __syncthreads();
    for (int i = tid; i < SKY_AX + SKY_AY; i += 256) {
        if (i < SKY_AX) tcol[i] = ux[reflect101(x0 - SKY_R + i, g.W)];
        else {
            ResizeTap t = uy[reflect101(y0 - SKY_R + (i - SKY_AX), g.H)];
            trow[i - SKY_AX] = t;
            atomicMin(&lr_min, t.i0);
            atomicMax(&lr_max, t.i1);
        }
    }
    __syncthreads();
    const int r_lo = lr_min, nlr = lr_max - lr_min + 1;

    // phase 1 (warp w takes rows w, w+8, ...; lanes sweep the 94 columns)
    const int lane = tid & 31, wid = tid >> 5;
    if (nlr <= SKY_HROWS) {
        for (int lr = wid; lr < nlr; lr += 8) {
            const float *row = L + (long long)(r_lo + lr) * lw;
#pragma unroll
            for (int c = lane; c < SKY_AX; c += 32) {
                ResizeTap a = tcol[c];
                hrow[lr * SKY_UP_LD + c] = lerp2(__ldg(row + a.i0), a.w0, __ldg(row + a.i1), a.w1);
            }
        }
        __syncthreads();
        // `up` is stored with rows interleaved in pairs: up2[rp][c] = (up[2rp][c], up[2rp+1][c]), the operand
        // layout of the packed row filter below
        for (int rp = wid; rp < SKY_AY / 2; rp += 8) {
            const ResizeTap b = trow[2 * rp], d = trow[2 * rp + 1];
            const float *h0 = hrow + (b.i0 - r_lo) * SKY_UP_LD, *h1 = hrow + (b.i1 - r_lo) * SKY_UP_LD;
            const float *g0 = hrow + (d.i0 - r_lo) * SKY_UP_LD, *g1 = hrow + (d.i1 - r_lo) * SKY_UP_LD;
            float2 *dst = (float2 *)up + rp * SKY_UP_LD;
#pragma unroll
            for (int c = lane; c < SKY_AX; c += 32)
                dst[c] = make_float2(lerp2(h0[c], b.w0, h1[c], b.w1), lerp2(g0[c], d.w0, g1[c], d.w1));
        }
    } else { // pathological geometry (tiny low-res grid folded many times): direct evaluation
        for (int i = tid; i < SKY_AY * SKY_AX; i += 256) {
            int r = i / SKY_AX, c = i - r * SKY_AX;
            ResizeTap a = tcol[c], b = trow[r];
            const float *r0 = L + (long long)b.i0 * lw, *r1 = L + (long long)b.i1 * lw;
            float h0 = lerp2(__ldg(r0 + a.i0), a.w0, __ldg(r0 + a.i1), a.w1);
            float h1 = lerp2(__ldg(r1 + a.i0), a.w0, __ldg(r1 + a.i1), a.w1);
            up[((r >> 1) * SKY_UP_LD + c) * 2 + (r & 1)] = lerp2(h0, b.w0, h1, b.w1);
        }
    }
    __syncthreads();

    // phase 2: rows (2rp, 2rp+1), columns 4gx .. 4gx+3 of rf from the 36-column window of the row pair
    for (int i = tid; i < (SKY_AY / 2) * (SKY_TX / 4); i += 256) {
        const int rp = i >> 4, gx = i & 15;
        const ulonglong2 *src = (const ulonglong2 *)((const float2 *)up + rp * SKY_UP_LD + 4 * gx);
        f32x2_t w[36];
#pragma unroll
        for (int q = 0; q < 18; q++) { ulonglong2 v = src[q]; w[2 * q] = v.x; w[2 * q + 1] = v.y; }
        float o0[4], o1[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float a, b;
            unpack2(mul2_bcast(c_gauss[0], w[e]), a, b);
#pragma unroll
            for (int j = 1; j < 31; j++) {
                float x, y;
                unpack2(mul2_bcast(c_gauss[j], w[e + j]), x, y);
                a = __fadd_rn(a, x);
                b = __fadd_rn(b, y);
            }
            o0[e] = a; o1[e] = b;
        }
        ((float4 *)(rf + (2 * rp) * SKY_TX + 4 * gx))[0] = make_float4(o0[0], o0[1], o0[2], o0[3]);
        ((float4 *)(rf + (2 * rp + 1) * SKY_TX + 4 * gx))[0] = make_float4(o1[0], o1[1], o1[2], o1[3]);
    }
    __syncthreads();

    // phase 3: out[4gy .. 4gy+3][c] from rf[4gy .. 4gy+33][c]
    double sum = 0.0, sq = 0.0;
    const float *G = gray + f * strideGray;
    float *S = stars + f * strideStars;
    float *SM = segmax + f * strideSeg;
    // tile-uniform classification: interior tiles skip every per-pixel bounds / ROI test
    const bool tile_in_image = x0 + SKY_TX <= g.W && y0 + SKY_TY <= g.H;
    const bool tile_in_roi = x0 >= g.roi_x && x0 + SKY_TX <= g.roi_x + g.roi_w && y0 >= g.roi_y && y0 + SKY_TY <= g.roi_y + g.roi_h;
    const bool tile_off_roi = x0 >= g.roi_x + g.roi_w || x0 + SKY_TX <= g.roi_x || y0 >= g.roi_y + g.roi_h || y0 + SKY_TY <= g.roi_y;
    if (tile_in_image && (tile_in_roi || tile_off_roi) && (g.W & 1) == 0) {
        // interior tiles: thread = rows 4gy..4gy+3 of the column pair (2cp, 2cp+1); one warp spans the 64 columns
        const unsigned half = lane < 16 ? 0x0000ffffu : 0xffff0000u; // a 32-px segment = 16 lanes
        for (int i = tid; i < (SKY_TY / 4) * (SKY_TX / 2); i += 256) {
            const int gy = i >> 5, cp = i & 31;
            const long long p = (long long)(y0 + 4 * gy) * g.W + x0 + 2 * cp;
            float2 gv[4];
#pragma unroll
            for (int e = 0; e < 4; e++) gv[e] = __ldg((const float2 *)(G + p + (long long)e * g.W));
            const f32x2_t *q = (const f32x2_t *)(rf + (4 * gy) * SKY_TX + 2 * cp);
            f32x2_t w[34];
#pragma unroll
            for (int j = 0; j < 34; j++) w[j] = q[j * (SKY_TX / 2)];
            float *mp = SM + (long long)(y0 + 4 * gy) * nseg + ((x0 + 2 * cp) >> 5);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float a, b;
                unpack2(mul2_bcast(c_gauss[15], w[e + 15]), a, b);
#pragma unroll
                for (int j = 1; j <= 15; j++) {
                    float x, y;
                    unpack2(mul2_bcast(c_gauss[15 + j], add2(w[e + 15 + j], w[e + 15 - j])), x, y);
                    a = __fadd_rn(a, x);
                    b = __fadd_rn(b, y);
                }
                const float va = __fsub_rn(gv[e].x, a), vb = __fsub_rn(gv[e].y, b);
                *(float2 *)(S + p + (long long)e * g.W) = make_float2(va, vb);
                if (tile_in_roi) {
                    double da = (double)va, db = (double)vb;
                    sum += da; sq += da * da;
                    sum += db; sq += db * db;
                }
                unsigned key = __float_as_uint(fmaxf(va, vb));
                key ^= (unsigned)((int)key >> 31) | 0x80000000u;
                key = __reduce_max_sync(half, key);
                if ((lane & 15) == 0) {
                    key ^= (key & 0x80000000u) ? 0x80000000u : 0xffffffffu;
                    mp[(long long)e * nseg] = __uint_as_float(key);
                }
            }
        }
    } else
    for (int i = tid; i < (SKY_TY / 4) * SKY_TX; i += 256) {
        const int gy = i >> 6, c = i & 63;
        const int x = x0 + c;
        // gray first: its HBM latency hides behind the window loads and ~190 fp32 operations
        float gv[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int y = y0 + 4 * gy + e;
            gv[e] = (x < g.W && y < g.H) ? __ldg(G + (long long)y * g.W + x) : 0.f;
        }
        const float *q = rf + (4 * gy) * SKY_TX + c;
        float w[34];
#pragma unroll
        for (int j = 0; j < 34; j++) w[j] = q[j * SKY_TX];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float s = __fmul_rn(c_gauss[15], w[e + 15]);
#pragma unroll
            for (int j = 1; j <= 15; j++) s = __fadd_rn(s, __fmul_rn(c_gauss[15 + j], __fadd_rn(w[e + 15 + j], w[e + 15 - j])));
            const int y = y0 + 4 * gy + e;
            float v = -INFINITY;
            if (x < g.W && y < g.H) {
                v = __fsub_rn(gv[e], s);
                S[(long long)y * g.W + x] = v;
                if (x >= g.roi_x && x < g.roi_x + g.roi_w && y >= g.roi_y && y < g.roi_y + g.roi_h) {
                    double d = (double)v;
                    sum += d;
                    sq += d * d;
                }
            }
            // warp maximum through one REDUX on an order-preserving integer image of the float
            unsigned key = __float_as_uint(v);
            key ^= (unsigned)((int)key >> 31) | 0x80000000u;
            key = __reduce_max_sync(0xffffffffu, key);
            if ((tid & 31) == 0 && y < g.H && (x >> 5) < nseg) {
                key ^= (key & 0x80000000u) ? 0x80000000u : 0xffffffffu;
                SM[(long long)y * nseg + (x >> 5)] = __uint_as_float(key);
            }
        }
    }
    // block reduction of the fp64 partial sums, fixed order
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((tid & 31) == 0) { sh[0][tid >> 5] = sum; sh[1][tid >> 5] = sq; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; w++) { a += sh[0][w]; b += sh[1][w]; }
        long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        partials[f * stridePart + 2 * blk] = a;
        partials[f * stridePart + 2 * blk + 1] = b;
    }
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
