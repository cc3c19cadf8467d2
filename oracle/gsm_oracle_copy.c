/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the foveated stereo copy. Nothing in the
 * product may import, link or call this file.
 *
 * What it follows: DepthFirstStereoCopyEncoder.swift:28-100 (one render pass, two viewports, vertex amplification, the
 * drawable's MTLRasterizationRateMap attached) and stereoCopyVertex / stereoCopyFragment (DFS.metal:1984-2018): a
 * full-screen triangle whose uv runs (0,0) at NDC(-1,-1) to (1,1) at NDC(1,1), sampled with
 * sampler(address::clamp_to_edge, filter::linear) from slice eyeIndex of the intermediate rgba16f array.
 *
 * PARITY UNPINNED: the reference holds no test or golden vector for this copy, and three things in it are Metal
 * implementation behaviour, not source: the rasterization-rate map's physical->screen mapping (taken here as an INPUT,
 * tabulated by the caller from MTLRasterizationRateMap.mapPhysicalToScreenCoordinates), the bilinear filter's weight
 * precision (restated here as binary32 lerps with 8-bit sub-texel weights), and the float->unorm8 / sRGB attachment conversion (restated as
 * round-to-nearest of the exact IEC 61966-2-1 curve). What IS pinned: at 1:1 (no rate map, viewports (0,0,W,H) and
 * (W,0,W,H), rgba16f) this function reproduces gsmo_stereo_copy, the literal row-flipped copy, bit for bit
 * (tests/test_foveated_oracle.py).
 *
 * Canonical arithmetic (binary32, no contraction except the explicit fmaf):
 *   screen = rate map ? screenX[px], screenY[py] : px + 0.5, py + 0.5
 *   covered iff ox <= sx < ox + vw and oy <= sy < oy + vh (the viewport clips the triangle)
 *   u = (sx - ox) / vw, v = (sy - oy) / vh, flipY: v = 1 - v; both clamped to [0, 1] (DFS.metal:2015)
 *   tx = rint(fmaf(u, W, -0.5) * 256) / 256 (8 bits of sub-texel precision, which is what makes the 1:1 copy exact),
 *   x0 = floor(tx), fx = tx - x0, taps x0 and x0 + 1 clamped to [0, W - 1]; same in y
 *   sample = lerp(h0, h1, fy), h0 = lerp(c00, c10, fx), h1 = lerp(c01, c11, fx), lerp(a, b, f) = f == 0 ? a : fmaf(f, b - a, a),
 *   rounded to half (any NaN becomes 0x7E00)
 *   (the fragment returns half4), then converted to the attachment's format.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "gsm_oracle.h"
#include "gsmo_math.h"

static uint8_t g_srgb[0x3C01];
static int g_srgbReady = 0;

static void buildSrgb(void) {
    if (g_srgbReady) return;
    for (uint32_t b = 0; b <= 0x3C00u; ++b) {
        double c = (double)gsmo_h2f((gsmo_half)b);
        double s = c <= 0.0031308 ? 12.92 * c : 1.055 * pow(c, 1.0 / 2.4) - 0.055;
        double q = floor(s * 255.0 + 0.5);
        g_srgb[b] = (uint8_t)(q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q));
    }
    g_srgbReady = 1;
}

static uint8_t unorm8(gsmo_half h) {
    float f = gsmo_h2f(h);
    if (!(f > 0.0f)) return 0;
    if (f >= 1.0f) return 255;
    return (uint8_t)nearbyintf(f * 255.0f);
}

static uint8_t srgb8(gsmo_half h) {
    float f = gsmo_h2f(h);
    if (!(f > 0.0f)) return 0;
    if (f >= 1.0f) return 255;
    return g_srgb[h];
}

static void storePixel(uint8_t* p, int format, const gsmo_half c[4]) {
    switch (format) {
    case 0: memcpy(p, c, 8); break;
    case 1: p[0] = unorm8(c[2]); p[1] = unorm8(c[1]); p[2] = unorm8(c[0]); p[3] = unorm8(c[3]); break;
    case 2: p[0] = srgb8(c[2]); p[1] = srgb8(c[1]); p[2] = srgb8(c[0]); p[3] = unorm8(c[3]); break;
    case 3: p[0] = unorm8(c[0]); p[1] = unorm8(c[1]); p[2] = unorm8(c[2]); p[3] = unorm8(c[3]); break;
    default: p[0] = srgb8(c[0]); p[1] = srgb8(c[1]); p[2] = srgb8(c[2]); p[3] = unorm8(c[3]); break;
    }
}

/* a zero weight returns the tap itself: keeps -0, and keeps an infinite neighbour from turning a 1:1 copy into NaN */
static float lerp(float a, float b, float f) { return f == 0.0f ? a : fmaf(f, b - a, a); }

static int axis(float s, float o, float extent, int flip, uint32_t n, int* a, int* b, float* frac) {
    if (!(s >= o && s < o + extent)) return 0;
    float t = (s - o) / extent;
    if (flip) t = 1.0f - t;
    t = gsmo_clamp(t, 0.0f, 1.0f);
    /* texel coordinate snapped to 8 fractional bits, as texture units do (D3D11.3 functional spec 7.18.8: "at least 8
     * bits of sub-texel precision"); without it a 1:1 copy would not be exact. Both steps are exact in binary32. */
    float q = nearbyintf(fmaf(t, (float)n, -0.5f) * 256.0f);
    float x0 = floorf(q * 0.00390625f);
    *frac = (q - x0 * 256.0f) * 0.00390625f;
    int i = (int)x0;
    *a = i < 0 ? 0 : (i > (int)n - 1 ? (int)n - 1 : i);
    *b = i + 1 < 0 ? 0 : (i + 1 > (int)n - 1 ? (int)n - 1 : i + 1);
    return 1;
}

void gsmo_stereo_copy_foveated(const gsmo_half* color2, uint32_t width, uint32_t height, int flipY, uint8_t* dst,
                               uint32_t textureWidth, uint32_t textureHeight, uint32_t arrayLength, size_t rowBytes,
                               size_t sliceBytes, int format, uint32_t layerCount, const uint32_t physicalWidth[2],
                               const uint32_t physicalHeight[2], const float* const screenX[2],
                               const float* const screenY[2], const double viewports[8]) {
    buildSrgb();
    const size_t px = format == 0 ? 8 : 4;
    const size_t eyeStride = (size_t)width * height * 4;
    for (int e = 0; e < 2; ++e) { /* left, then right: the right eye wins where shared viewports overlap */
        const uint32_t slice = arrayLength >= 2 ? (uint32_t)e : 0u;
        const uint32_t layer = layerCount ? (slice < layerCount - 1 ? slice : layerCount - 1) : 0u;
        const float ox = (float)viewports[4 * e + 0], oy = (float)viewports[4 * e + 1];
        const float vw = (float)viewports[4 * e + 2], vh = (float)viewports[4 * e + 3];
        uint32_t w = textureWidth, h = textureHeight;
        if (layerCount) {
            if (physicalWidth[layer] < w) w = physicalWidth[layer];
            if (physicalHeight[layer] < h) h = physicalHeight[layer];
        }
        const gsmo_half* src = color2 + (size_t)e * eyeStride;
#pragma omp parallel for schedule(static)
        for (uint32_t y = 0; y < h; ++y) {
            int ya, yb;
            float fy;
            const float sy = layerCount ? screenY[layer][y] : (float)y + 0.5f;
            if (!axis(sy, oy, vh, flipY, height, &ya, &yb, &fy)) continue;
            for (uint32_t x = 0; x < w; ++x) {
                int xa, xb;
                float fx;
                const float sx = layerCount ? screenX[layer][x] : (float)x + 0.5f;
                if (!axis(sx, ox, vw, 0, width, &xa, &xb, &fx)) continue;
                gsmo_half c[4];
                for (int k = 0; k < 4; ++k) {
                    const float c00 = gsmo_h2f(src[((size_t)ya * width + xa) * 4 + k]), c10 = gsmo_h2f(src[((size_t)ya * width + xb) * 4 + k]);
                    const float c01 = gsmo_h2f(src[((size_t)yb * width + xa) * 4 + k]), c11 = gsmo_h2f(src[((size_t)yb * width + xb) * 4 + k]);
                    const float h0 = lerp(c00, c10, fx), h1 = lerp(c01, c11, fx);
                    const float v = lerp(h0, h1, fy);
                    c[k] = v != v ? (gsmo_half)0x7E00u : gsmo_f2h(v); /* one NaN, whatever the payload */
                }
                storePixel(dst + (size_t)slice * sliceBytes + (size_t)y * rowBytes + (size_t)x * px, format, c);
            }
        }
    }
}
