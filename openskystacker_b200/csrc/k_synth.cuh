// Synthetic-frame generator (bench / test utility, not on the reference's path): renders the
// seeded star field of SURVEY.md §8(d) straight into HBM so that benchmark inputs do not
// have to be rendered on the host and pushed over PCIe.  Frames produced here are plain
// u16 HWC buffers; parity runs feed the very same bytes to the CPU oracle.
#pragma once
#include "common.cuh"

namespace oss {

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// two uniforms in (0,1] from a counter -> one N(0,1) sample (Box-Muller)
__device__ __forceinline__ float gauss_noise(uint32_t seed, uint32_t idx)
{
    uint32_t a = mix32(idx * 2u + 0x9e3779b9u * seed), b = mix32(idx * 2u + 1u + 0x85ebca6bu * (seed + 1u));
    float u1 = ((float)(a >> 8) + 1.f) * (1.f / 16777216.f), u2 = (float)(b >> 8) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__global__ void k_synth_fill(float *__restrict__ plane, long long P, float v)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) plane[i] = v;
}

// one CTA per star: splat a Gaussian PSF (x, y, amplitude, sigma) onto the plane
__global__ void k_synth_splat(float *__restrict__ plane, int W, int H, const float4 *__restrict__ stars, int n)
{
    const int s = blockIdx.x;
    if (s >= n) return;
    const float4 st = stars[s];
    const int r = (int)ceilf(6.f * st.w);
    const int x0 = (int)st.x - r, y0 = (int)st.y - r, side = 2 * r + 1;
    const float inv = -0.5f / (st.w * st.w);
    for (int i = threadIdx.x; i < side * side; i += blockDim.x) {
        int y = y0 + i / side, x = x0 + i % side;
        if (x < 0 || y < 0 || x >= W || y >= H) continue;
        float dx = (float)x + 0.5f - st.x, dy = (float)y + 0.5f - st.y;
        atomicAdd(&plane[(long long)y * W + x], st.z * __expf((dx * dx + dy * dy) * inv));
    }
}

// sky plane (+ optional vignetting) + noise + pedestal -> u16 HWC with per-channel gains
__global__ void k_synth_quantize(const float *__restrict__ plane, uint16_t *__restrict__ out, int W, int H, int C, float noise,
                                 float pedestal_adu, int vignette, uint32_t seed, float g0, float g1, float g2)
{
    const long long P = (long long)W * H;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float v = plane[p];
        if (vignette) {
            int y = (int)(p / W), x = (int)(p - (long long)y * W);
            float fx = (float)x - 0.5f * W, fy = (float)y - 0.5f * H;
            float r2 = (fx * fx + fy * fy) / (0.25f * ((float)W * W + (float)H * H));
            v *= 1.f - 0.3f * r2;
        }
        float adu = (v + noise * gauss_noise(seed, (uint32_t)p)) * 65535.f + pedestal_adu;
        const float g[3] = {g0, g1, g2};
        for (int c = 0; c < C; c++) out[p * C + c] = (uint16_t)fminf(fmaxf(rintf(adu * g[c]), 0.f), 65535.f);
    }
}

// calibration frames: level + N(0, sigma) ADU, optionally times the vignetting profile (flats)
__global__ void k_synth_calib(uint16_t *__restrict__ out, int W, int H, int C, float level_adu, float sigma_adu, int flat_profile,
                              float pedestal_adu, uint32_t seed)
{
    const long long P = (long long)W * H;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float lvl = level_adu;
        if (flat_profile) {
            int y = (int)(p / W), x = (int)(p - (long long)y * W);
            float fx = (float)x - 0.5f * W, fy = (float)y - 0.5f * H;
            float r2 = (fx * fx + fy * fy) / (0.25f * ((float)W * W + (float)H * H));
            lvl *= 1.f - 0.3f * r2;
        }
        float adu = fmaxf(lvl + pedestal_adu + sigma_adu * gauss_noise(seed, (uint32_t)p), 1.f);
        uint16_t q = (uint16_t)fminf(rintf(adu), 65535.f);
        for (int c = 0; c < C; c++) out[p * C + c] = q;
    }
}

}  // namespace oss
