// Shared declarations of liboss_b200 (sm_100a only).
//
// Arithmetic contract: every kernel reproduces the reference's CPU results
// bit-for-bit where the reference is deterministic.  The
// whole library is compiled with -fmad=false (no FMA contraction), IEEE division
// and square root (nvcc defaults; never --use_fast_math).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/oss_b200.h"

#define OSS_NOBJS 40      // focas.h:50
#define OSS_MAX_MATCH 120 // focas.h:52
#define OSS_MAX_TRI 9880  // C(40,3)
#define OSS_SEG 32        // pixels per detection segment (one warp-wide row piece)
#define OSS_SEGROWS 8    // rows per segment-maximum cell (k_sky_sub_stats writes one maximum per OSS_SEG px x OSS_SEGROWS rows)
#define OSS_DEBLEND_MAX 10000 // adjoiningpixel.cpp:45

namespace oss {

// Geometry of one stack (all frames alike), plus derived sizes.
struct Geom {
    int W, H, C;
    int sample;       // oss_sample
    int lw, lh;       // low-res sky grid: W/8, H/8 (stardetector.cpp:161)
    int nseg;         // ceil(W / OSS_SEG)
    int roi_x, roi_y, roi_w, roi_h; // stardetector.cpp:121
    long long P;      // W*H
    int cap;          // candidate-pixel capacity per frame
};

// Per-frame detection scalars, produced and consumed on the device (no host round trip
// between the statistics pass and thresholding).
struct DetectState {
    double mean, sigma;
    float threshold, minpeak;
    int ncand;   // pixels > threshold
    int nroots;  // connected components (all)
    int ncomp;   // components with peak > minPeak
    int nstars;
    int overflow;
    int nbig;    // components above CW_CAP_SMALL pixels, listed by the first k_components instance for the second
    int nbig2;   // components of CW_CAP + 1 .. kCompBigSmall pixels, listed by the second k_components instance for k_components_big
    int nbig3;   // ... and the larger ones (second instance of k_components_big, incl. the >= 10000-px bypass)
};

// Result of the alignment kernel for one light.
struct AlignOut {
    int k, err;
    float M[6];
};

// One batch worth of device work space; slot s of an array lives at base + s*stride.
struct Workspace {
    // dense planes
    float *cal;     // [slots][P*C]  calibrated frame
    float *gray;    // [slots][P]    luminance (aliases cal when C == 1)
    float *stars;   // [slots][P]    gray - sky
    float *lo;      // [slots][lw*lh]
    float *med;     // [slots][lw*lh]
    float *segmax;  // [slots][H*nseg]  maxima of 32 px x 8 rows cells, cell (y / 8, x / 32) at (y / 8) * nseg + x / 32
    uint32_t *segmask; // [slots][H*nseg]
    int *segoff;    // [slots][H*nseg + 1]
    double *partials; // [slots][2*stat_blocks]
    DetectState *ds;  // [slots]
    // sparse detection lists (capacity cap per slot)
    int *cand_idx; float *cand_val; int4 *nbr; int *parent; int *owner;
    int *comp_size; int *comp_peak; int *flag; int *flagscan;
    int *comp_root; int *comp_off; int *comp_nstar; int *comp_staroff;
    int *scratch;   // [slots][12*cap]
    oss_star *star_tmp; // [slots][cap]
    oss_star *star_out; // [slots][cap]
    int *blocksum;  // [slots][scan blocks]
    AlignOut *align; // [slots]
    int2 *warp_ab;   // [slots][W]   (adelta, bdelta) of cv::warpAffine's fixed-point coordinates
    int2 *warp_xy;   // [slots][H]   (X0, Y0)
    long long strideP, stridePC, strideLo, strideSeg, stridePart, strideScan;
    int stat_blocks;
};

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// ---- shared-memory addresses and mbarriers (TMA / bulk-copy completion, stage hand-over) --------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bulk asynchronous copy global -> shared of `bytes` (a multiple of 16; both addresses 16-byte aligned), completed on `bar`
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace oss
