// Step 10 of the stereo frame through a drawable: StereoRenderTarget.foveated.
// Replaces DepthFirstStereoCopyEncoder.encodeRender (DepthFirstStereoCopyEncoder.swift:28-100) + stereoCopyVertex /
// stereoCopyFragment (DFS.metal:1984-2018): one full-screen triangle per eye, clipped to the eye's viewport, rasterised
// through the drawable's rasterization-rate map, sampling the eye's intermediate image with a linear clamp-to-edge
// sampler. Here: one thread per physical texel of the drawable; the rate map arrives tabulated (gsm.h).
//
// Arithmetic (stated once more in the CPU oracle, which the tests compare against bit for bit): binary32 with
// explicit rounding, texel coordinates snapped to 8 fractional bits, lerp(a, b, f) = f == 0 ? a : fma(f, b - a, a), the
// sample rounded to half (the fragment returns half4), then the attachment conversion: unorm8 = rint(clamp(x) * 255),
// sRGB through a table over the 15361 halfs of [0, 1] built from the IEC 61966-2-1 curve in double.
//
// HBM-bound: per covered texel 4 taps x 8 B (neighbours share them through L1/L2) and one 4 or 8 B store.
#include <cmath>
#include <vector>

#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

namespace {

struct Axis { int a, b; float f; bool in; };

__device__ __forceinline__ Axis axisOf(float s, float o, float extent, bool flip, uint32_t n) {
    Axis r;
    r.in = s >= o && s < __fadd_rn(o, extent);
    float t = __fdiv_rn(__fsub_rn(s, o), extent);
    if (flip) t = __fsub_rn(1.0f, t);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float q = rintf(__fmul_rn(__fmaf_rn(t, (float)n, -0.5f), 256.0f));
    const float x0 = floorf(__fmul_rn(q, 0.00390625f));
    r.f = __fmul_rn(__fsub_rn(q, __fmul_rn(x0, 256.0f)), 0.00390625f);
    const int i = (int)x0, hi = (int)n - 1;
    r.a = min(max(i, 0), hi);
    r.b = min(max(i + 1, 0), hi);
    return r;
}

__device__ __forceinline__ float lerpTap(float a, float b, float f) { return f == 0.0f ? a : __fmaf_rn(f, __fsub_rn(b, a), a); }

__device__ __forceinline__ uint32_t unorm8(__half h) {
    const float f = __half2float(h);
    if (!(f > 0.0f)) return 0u;
    if (f >= 1.0f) return 255u;
    return (uint32_t)__float2int_rn(__fmul_rn(f, 255.0f));
}

__device__ __forceinline__ uint32_t srgb8(__half h, const uint8_t* __restrict__ lut) {
    const float f = __half2float(h);
    if (!(f > 0.0f)) return 0u;
    if (f >= 1.0f) return 255u;
    return __ldg(lut + __half_as_ushort(h));
}

__global__ void __launch_bounds__(256) stereo_copy_kernel(StereoCopyParams p) {
    const uint32_t x = blockIdx.x * 32u + (threadIdx.x & 31u), y = blockIdx.y * 8u + (threadIdx.x >> 5);
    const uint32_t slice = blockIdx.z;
    if (x >= p.textureWidth || y >= p.textureHeight) return;
    const uint32_t layer = p.layerCount ? min(slice, p.layerCount - 1u) : 0u;
    if (p.layerCount && (x >= p.physicalWidth[layer] || y >= p.physicalHeight[layer])) return;
    const float sx = p.layerCount ? __ldg(p.screenX[layer] + x) : __fadd_rn((float)x, 0.5f);
    const float sy = p.layerCount ? __ldg(p.screenY[layer] + y) : __fadd_rn((float)y, 0.5f);
    // left, then right: where two viewports of a shared texture overlap the right eye's draw lands last
    int eye = -1;
    Axis ax, ay;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        if ((p.arrayLength >= 2u ? (uint32_t)e : 0u) != slice) continue;
        const Axis cx = axisOf(sx, p.vp[e][0], p.vp[e][2], false, p.width);
        const Axis cy = axisOf(sy, p.vp[e][1], p.vp[e][3], p.flipY != 0, p.height);
        if (cx.in && cy.in) { eye = e; ax = cx; ay = cy; }
    }
    if (eye < 0) return;
    const __half* src = p.src + (size_t)eye * p.srcEyeStride;  // halfs
    const uint2 t00 = __ldg(reinterpret_cast<const uint2*>(src + (size_t)ay.a * p.srcRowStride + (size_t)ax.a * 4));
    const uint2 t10 = __ldg(reinterpret_cast<const uint2*>(src + (size_t)ay.a * p.srcRowStride + (size_t)ax.b * 4));
    const uint2 t01 = __ldg(reinterpret_cast<const uint2*>(src + (size_t)ay.b * p.srcRowStride + (size_t)ax.a * 4));
    const uint2 t11 = __ldg(reinterpret_cast<const uint2*>(src + (size_t)ay.b * p.srcRowStride + (size_t)ax.b * 4));
    const __half* c00 = reinterpret_cast<const __half*>(&t00);
    const __half* c10 = reinterpret_cast<const __half*>(&t10);
    const __half* c01 = reinterpret_cast<const __half*>(&t01);
    const __half* c11 = reinterpret_cast<const __half*>(&t11);
    __half c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float h0 = lerpTap(__half2float(c00[k]), __half2float(c10[k]), ax.f);
        const float h1 = lerpTap(__half2float(c01[k]), __half2float(c11[k]), ax.f);
        const float v = lerpTap(h0, h1, ay.f);
        c[k] = v != v ? __ushort_as_half((unsigned short)0x7E00u) : __float2half_rn(v);  // one NaN, whatever the payload
    }
    uint8_t* row = reinterpret_cast<uint8_t*>(p.dst) + (size_t)slice * p.sliceBytes + (size_t)y * p.rowBytes;
    switch (p.format) {
    case GSM_PIXEL_RGBA16F: {
        uint2 v;
        v.x = (uint32_t)__half_as_ushort(c[0]) | ((uint32_t)__half_as_ushort(c[1]) << 16);
        v.y = (uint32_t)__half_as_ushort(c[2]) | ((uint32_t)__half_as_ushort(c[3]) << 16);
        *reinterpret_cast<uint2*>(row + (size_t)x * 8) = v;
        break;
    }
    case GSM_PIXEL_BGRA8:
        *reinterpret_cast<uint32_t*>(row + (size_t)x * 4) = unorm8(c[2]) | (unorm8(c[1]) << 8) | (unorm8(c[0]) << 16) | (unorm8(c[3]) << 24);
        break;
    case GSM_PIXEL_BGRA8_SRGB:
        *reinterpret_cast<uint32_t*>(row + (size_t)x * 4) =
            srgb8(c[2], p.srgbLut) | (srgb8(c[1], p.srgbLut) << 8) | (srgb8(c[0], p.srgbLut) << 16) | (unorm8(c[3]) << 24);
        break;
    case GSM_PIXEL_RGBA8:
        *reinterpret_cast<uint32_t*>(row + (size_t)x * 4) = unorm8(c[0]) | (unorm8(c[1]) << 8) | (unorm8(c[2]) << 16) | (unorm8(c[3]) << 24);
        break;
    default:
        *reinterpret_cast<uint32_t*>(row + (size_t)x * 4) =
            srgb8(c[0], p.srgbLut) | (srgb8(c[1], p.srgbLut) << 8) | (srgb8(c[2], p.srgbLut) << 16) | (unorm8(c[3]) << 24);
        break;
    }
}

}  // namespace

void buildSrgbEncodeTable(uint8_t* table) {
    for (uint32_t b = 0; b < kSrgbTableSize; ++b) {
        const uint32_t e = (b >> 10) & 31u, m = b & 1023u;  // b <= 0x3C00: non-negative, finite
        const double c = e ? std::ldexp(1.0 + m / 1024.0, (int)e - 15) : std::ldexp(m / 1024.0, -14);
        const double s = c <= 0.0031308 ? 12.92 * c : 1.055 * std::pow(c, 1.0 / 2.4) - 0.055;
        const double q = std::floor(s * 255.0 + 0.5);
        table[b] = (uint8_t)(q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q));
    }
}

cudaError_t launchStereoCopy(cudaStream_t s, const StereoCopyParams& p) {
    if (p.textureWidth == 0 || p.textureHeight == 0) return cudaSuccess;
    dim3 grid((p.textureWidth + 31u) / 32u, (p.textureHeight + 7u) / 8u, p.arrayLength >= 2u ? 2u : 1u);
    stereo_copy_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gsm
