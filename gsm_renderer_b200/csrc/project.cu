// project.cu -- stage 1 (+1.25): project + cull + SH colour + quantise + exact tile count, fused with the
// visibility compaction (ascending gid order, single pass, decoupled look-back).
//
// Replaces depthFirstProjectCullKernel (DFS.metal:46-219), depthFirstStereoProjectCullKernel
// (DFS.metal:341-499) and the 8-pass VisibilityCompactionEncoder (DFS.metal:518-621).
// Arithmetic follows the canonical evaluation order of DESIGN.md section 3 (compiled with -fmad=false).
#include "gsm_common.cuh"
#include "gsm_dmath.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

constexpr uint32_t kProjThreads = 128;  // small CTAs: a slow warp (large splat) pins fewer sibling warp slots

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M3 { V3 c0, c1, c2; };
struct M2 { float m00, m01, m10, m11; };

__device__ __forceinline__ V4 mul44(const float* m, V4 v) {
    V4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}
__device__ __forceinline__ V3 mul33v(const M3& a, V3 v) {
    V3 r;
    r.x = (a.c0.x * v.x + a.c1.x * v.y) + a.c2.x * v.z;
    r.y = (a.c0.y * v.x + a.c1.y * v.y) + a.c2.y * v.z;
    r.z = (a.c0.z * v.x + a.c1.z * v.y) + a.c2.z * v.z;
    return r;
}
__device__ __forceinline__ M3 mul33(const M3& a, const M3& b) {
    M3 r;
    r.c0 = mul33v(a, b.c0);
    r.c1 = mul33v(a, b.c1);
    r.c2 = mul33v(a, b.c2);
    return r;
}

// GaussianShared.h:289-295
__device__ __forceinline__ V4 normalizeQuaternion(V4 q) {
    float d = ((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w;
    float norm = sqrtf(dmax(d, 1e-8f));
    if (norm < 1e-8f) return V4{1.0f, 0.0f, 0.0f, 0.0f};
    return V4{q.x / norm, q.y / norm, q.z / norm, q.w / norm};
}

// GaussianShared.h:297-324 (second normalisation kept, quirk Q1)
__device__ __forceinline__ M3 buildCovariance3D(V3 scale, V4 quat) {
    V4 q = normalizeQuaternion(quat);
    float x = q.x, y = q.y, z = q.z, r = q.w;
    float xx = x * x, yy = y * y, zz = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    V3 row0{1.0f - 2.0f * (yy + zz), 2.0f * (xy - r * z), 2.0f * (xz + r * y)};
    V3 row1{2.0f * (xy + r * z), 1.0f - 2.0f * (xx + zz), 2.0f * (yz - r * x)};
    V3 row2{2.0f * (xz - r * y), 2.0f * (yz + r * x), 1.0f - 2.0f * (xx + yy)};
    // R columns = (row0.k, row1.k, row2.k); RSk = R[k] * scale.k
    V3 RS0{row0.x * scale.x, row1.x * scale.x, row2.x * scale.x};
    V3 RS1{row0.y * scale.y, row1.y * scale.y, row2.y * scale.y};
    V3 RS2{row0.z * scale.z, row1.z * scale.z, row2.z * scale.z};
    M3 c;
    c.c0.x = (RS0.x * RS0.x + RS1.x * RS1.x) + RS2.x * RS2.x;
    c.c0.y = (RS0.x * RS0.y + RS1.x * RS1.y) + RS2.x * RS2.y;
    c.c0.z = (RS0.x * RS0.z + RS1.x * RS1.z) + RS2.x * RS2.z;
    c.c1.x = (RS0.y * RS0.x + RS1.y * RS1.x) + RS2.y * RS2.x;
    c.c1.y = (RS0.y * RS0.y + RS1.y * RS1.y) + RS2.y * RS2.y;
    c.c1.z = (RS0.y * RS0.z + RS1.y * RS1.z) + RS2.y * RS2.z;
    c.c2.x = (RS0.z * RS0.x + RS1.z * RS1.x) + RS2.z * RS2.x;
    c.c2.y = (RS0.z * RS0.y + RS1.z * RS1.y) + RS2.z * RS2.y;
    c.c2.z = (RS0.z * RS0.z + RS1.z * RS1.z) + RS2.z * RS2.z;
    return c;
}

// GaussianShared.h:326-388
__device__ __forceinline__ M2 projectCovariance2D(const M3& cov3d, V3 viewPos, const float* view, const float* proj,
                                                  float width, float height) {
    float absZ = fabsf(viewPos.z);
    float signZ = (viewPos.z >= 0.0f) ? 1.0f : -1.0f;
    float safeAbsZ = dmax(absZ, 1e-4f);
    float invAbsZ = 1.0f / safeAbsZ;
    float invAbsZ2 = invAbsZ * invAbsZ;
    float tanHalfFovX = 1.0f / dmax(fabsf(proj[0]), 1e-4f);
    float tanHalfFovY = 1.0f / dmax(fabsf(proj[5]), 1e-4f);
    float limX = 1.3f * tanHalfFovX;
    float limY = 1.3f * tanHalfFovY;
    float tx = viewPos.x * invAbsZ;
    float ty = viewPos.y * invAbsZ;
    float xClamped = dclamp(tx, -limX, limX) * safeAbsZ;
    float yClamped = dclamp(ty, -limY, limY) * safeAbsZ;
    float focalX = width * fabsf(proj[0]) * 0.5f;
    float focalY = height * fabsf(proj[5]) * 0.5f;
    M3 J;
    J.c0 = V3{focalX * invAbsZ, 0.0f, 0.0f};
    J.c1 = V3{0.0f, focalY * invAbsZ, 0.0f};
    J.c2 = V3{-focalX * xClamped * signZ * invAbsZ2, -focalY * yClamped * signZ * invAbsZ2, 0.0f};
    M3 W;
    W.c0 = V3{view[0], view[1], view[2]};
    W.c1 = V3{view[4], view[5], view[6]};
    W.c2 = V3{view[8], view[9], view[10]};
    M3 T = mul33(J, W);
    M3 Tt;
    Tt.c0 = V3{T.c0.x, T.c1.x, T.c2.x};
    Tt.c1 = V3{T.c0.y, T.c1.y, T.c2.y};
    Tt.c2 = V3{T.c0.z, T.c1.z, T.c2.z};
    M3 covFull = mul33(mul33(T, cov3d), Tt);
    M2 c;
    c.m00 = covFull.c0.x + 0.3f;
    c.m01 = covFull.c0.y;
    c.m10 = covFull.c1.x;
    c.m11 = covFull.c1.y + 0.3f;
    return c;
}

// GaussianShared.h:655-714
__device__ __forceinline__ M2 stabilizeCovariance2D(M2 cov, float width, float height) {
    float maxCond = 256.0f * 256.0f;
    float maxDim = dmax(width, height);
    float maxExtentPx = maxDim * 2.0f;
    float maxEig = maxExtentPx / 3.0f;
    maxEig = maxEig * maxEig;
    float a = cov.m00;
    float b = 0.5f * (cov.m01 + cov.m10);
    float d = cov.m11;
    if (!dfinite(a) || !dfinite(b) || !dfinite(d)) return M2{1.0f, 0.0f, 0.0f, 1.0f};
    a = dmax(a, 1e-4f);
    d = dmax(d, 1e-4f);
    float det = a * d - b * b;
    if (!dfinite(det) || det < 1e-8f) {
        float bump = (1e-8f - det) + 1e-4f;
        a = a + bump;
        d = d + bump;
        det = a * d - b * b;
    }
    float mid = 0.5f * (a + d);
    float disc = dmax(mid * mid - det, 0.0f);
    float sqrtDisc = sqrtf(disc);
    float lambda1 = mid + sqrtDisc;
    float lambda2 = dmax(mid - sqrtDisc, 1e-4f);
    float v1x, v1y;
    if (fabsf(b) > 1e-8f) {
        float vx = b;
        float vy = lambda1 - a;
        float vlen = sqrtf(vx * vx + vy * vy);
        float dn = dmax(vlen, 1e-8f);
        v1x = vx / dn;
        v1y = vy / dn;
    } else if (a >= d) {
        v1x = 1.0f; v1y = 0.0f;
    } else {
        v1x = 0.0f; v1y = 1.0f;
    }
    float v2x = v1y, v2y = -v1x;
    lambda1 = dmin(lambda1, maxEig);
    lambda2 = dmax(lambda2, lambda1 / maxCond);
    M2 o;
    o.m00 = lambda1 * (v1x * v1x) + lambda2 * (v2x * v2x);
    o.m01 = lambda1 * (v1x * v1y) + lambda2 * (v2x * v2y);
    o.m10 = lambda1 * (v1y * v1x) + lambda2 * (v2y * v2x);
    o.m11 = lambda1 * (v1y * v1y) + lambda2 * (v2y * v2y);
    return o;
}

// GaussianShared.h:446-488
__device__ __forceinline__ bool covarianceToThetaSigmas(M2 cov, float& theta, float& sigma1, float& sigma2) {
    float a = cov.m00;
    float b = 0.5f * (cov.m01 + cov.m10);
    float d = cov.m11;
    if (!dfinite(a) || !dfinite(b) || !dfinite(d)) return false;
    a = dmax(a, 1e-8f);
    d = dmax(d, 1e-8f);
    float det = a * d - b * b;
    if (!dfinite(det) || !(det > 0.0f)) return false;
    float mid = 0.5f * (a + d);
    float disc = dmax(mid * mid - det, 0.0f);
    float sqrtDisc = sqrtf(disc);
    float lambda1 = dmax(mid + sqrtDisc, 1e-8f);
    float lambda2 = dmax(mid - sqrtDisc, 1e-8f);
    float v1x, v1y;
    if (fabsf(b) > 1e-8f) {
        float vx = b, vy = lambda1 - a;
        float len = sqrtf(vx * vx + vy * vy);
        v1x = vx / len;
        v1y = vy / len;
    } else if (a >= d) {
        v1x = 1.0f; v1y = 0.0f;
    } else {
        v1x = 0.0f; v1y = 1.0f;
    }
    float t = datan2(v1y, v1x);
    t = dfmod_pi(t);
    if (t < 0.0f) t = t + GSM_PI_F;
    if (t >= GSM_PI_F) t = t - GSM_PI_F;
    theta = t;
    sigma1 = sqrtf(lambda1);
    sigma2 = sqrtf(lambda2);
    return dfinite(t) && dfinite(sigma1) && dfinite(sigma2);
}

// GaussianShared.h:275-278, :739-752
__device__ __forceinline__ bool cullByTotalInk(float opacity, float detCov2d, float depth, float nearPlane,
                                               float farPlane, float totalInkThreshold) {
    if (totalInkThreshold <= 0.0f) return false;
    float totalInk = opacity * 6.283185f * sqrtf(dmax(detCov2d, 1e-12f));
    float adjustedFarPlane = farPlane * 0.02f;
    float s = dclamp((adjustedFarPlane - depth) / (adjustedFarPlane - nearPlane), 0.0f, 1.0f);
    float depthFactor = 1.0f - s * s;
    float adjustedThreshold = depthFactor * totalInkThreshold;
    return totalInk < adjustedThreshold;
}

// GaussianShared.h:402-427
__device__ __forceinline__ void computeOBBExtents(M2 cov, float mult, float& ex, float& ey) {
    float a = cov.m00, b = cov.m01, d = cov.m11;
    float det = a * d - b * b;
    float mid = 0.5f * (a + d);
    float disc = dmax(mid * mid - det, 1e-6f);
    float sqrtDisc = sqrtf(disc);
    float lambda1 = mid + sqrtDisc;
    float lambda2 = dmax(mid - sqrtDisc, 1e-6f);
    float e1 = mult * sqrtf(dmax(lambda1, 1e-6f));
    float e2 = mult * sqrtf(dmax(lambda2, 1e-6f));
    float v1x, v1y;
    if (fabsf(b) > 1e-6f) {
        float vx = b, vy = lambda1 - a;
        float vlen = sqrtf(vx * vx + vy * vy);
        float dn = dmax(vlen, 1e-6f);
        v1x = vx / dn;
        v1y = vy / dn;
    } else if (a >= d) {
        v1x = 1.0f; v1y = 0.0f;
    } else {
        v1x = 0.0f; v1y = 1.0f;
    }
    ex = fabsf(v1x) * e1 + fabsf(v1y) * e2;
    ey = fabsf(v1y) * e1 + fabsf(v1x) * e2;
}

// GaussianShared.h:434-440
__device__ __forceinline__ uint16_t packThetaPi(float theta) {
    theta = dfmod_pi(theta);
    if (theta < 0.0f) theta = theta + GSM_PI_F;
    float u = theta * GSM_THETA_PACK;
    return (uint16_t)dclamp(u + 0.5f, 0.0f, 65535.0f);
}

// GaussianShared.h:791-828
struct TileBounds { int minTX, maxTX, minTY, maxTY; bool valid; };
__device__ __forceinline__ TileBounds computeTileBounds(float sx, float sy, float ex, float ey, float width,
                                                        float height, int tilesX, int tilesY) {
    TileBounds r;
    float xmin = sx - ex, xmax = sx + ex, ymin = sy - ey, ymax = sy + ey;
    float maxW = width - 1.0f, maxH = height - 1.0f;
    xmin = dclamp(xmin, 0.0f, maxW);
    xmax = dclamp(xmax, 0.0f, maxW);
    ymin = dclamp(ymin, 0.0f, maxH);
    ymax = dclamp(ymax, 0.0f, maxH);
    r.minTX = (int)floorf(xmin / (float)kTile);
    r.maxTX = (int)ceilf(xmax / (float)kTile) - 1;
    r.minTY = (int)floorf(ymin / (float)kTile);
    r.maxTY = (int)ceilf(ymax / (float)kTile) - 1;
    r.minTX = max(r.minTX, 0);
    r.minTY = max(r.minTY, 0);
    r.maxTX = min(r.maxTX, tilesX - 1);
    r.maxTY = min(r.maxTY, tilesY - 1);
    r.valid = (r.minTX <= r.maxTX && r.minTY <= r.maxTY);
    return r;
}

// GaussianShared.h:118-121
__device__ __forceinline__ float srgbToLinearChannel(float c) {
    c = dclamp(c, 0.0f, 1.0f);
    return (c <= 0.04045f) ? (c / 12.92f) : dpowr((c + 0.055f) / 1.055f, 2.4f);
}

__device__ __forceinline__ uint8_t quantU8(float v) { return (uint8_t)dclamp(v * 255.0f, 0.0f, 255.0f); }

// DFS.metal:33-37
__device__ __forceinline__ uint32_t float_to_sortable_uint(float v) {
    uint32_t bits = __float_as_uint(v);
    uint32_t mask = (bits & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u;
    return bits ^ mask;
}

// ---------------------------------------------------------------- record / SH loads (128-bit where the layout allows)
struct GaussianIn { V3 pos, scale; V4 rot; float opacity; };

template <bool HALF>
__device__ __forceinline__ GaussianIn loadGaussian(const void* base, uint32_t gid) {
    GaussianIn g;
    if (HALF) {
        const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * (size_t)gid;
        uint4 a = __ldcs(p), b = __ldcs(p + 1);
        g.pos = V3{__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)};
        __half2 h0 = *reinterpret_cast<__half2*>(&a.w);  // opacity, sx
        __half2 h1 = *reinterpret_cast<__half2*>(&b.x);  // sy, sz
        __half2 h2 = *reinterpret_cast<__half2*>(&b.y);  // rx, ry
        __half2 h3 = *reinterpret_cast<__half2*>(&b.z);  // rz, rw
        g.opacity = __low2float(h0);
        g.scale = V3{__high2float(h0), __low2float(h1), __high2float(h1)};
        g.rot = V4{__low2float(h2), __high2float(h2), __low2float(h3), __high2float(h3)};
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(base) + 3 * (size_t)gid;
        uint4 a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2);
        g.pos = V3{__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z)};
        g.opacity = __uint_as_float(a.w);
        g.scale = V3{__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)};
        g.rot = V4{__uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w)};
    }
    return g;
}

// Loads the 3*K coefficients of one Gaussian into registers as float, using the widest aligned vector the
// per-Gaussian stride permits (K=16: 96 B half / 192 B float => 128-bit loads).
template <bool HALF, int K>
__device__ __forceinline__ void loadSH(const void* base, uint32_t gid, float (&out)[3 * K]) {
    constexpr int N = 3 * K;
    constexpr int BYTES = N * (HALF ? 2 : 4);
    constexpr int VEC = (BYTES % 16 == 0) ? 16 : (BYTES % 8 == 0) ? 8 : (BYTES % 4 == 0) ? 4 : 2;
    const char* p = reinterpret_cast<const char*>(base) + (size_t)gid * BYTES;
    constexpr int NV = BYTES / VEC;
    constexpr int WORDS = VEC >= 4 ? VEC / 4 : 1;
    if (VEC >= 4) {
        uint32_t w[NV * WORDS];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (VEC == 16) {
                uint4 v = __ldcs(reinterpret_cast<const uint4*>(p) + i);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            } else if (VEC == 8) {
                uint2 v = __ldcs(reinterpret_cast<const uint2*>(p) + i);
                w[2 * i] = v.x; w[2 * i + 1] = v.y;
            } else {
                w[i] = __ldcs(reinterpret_cast<const uint32_t*>(p) + i);
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (HALF) {
                __half2 h = *reinterpret_cast<__half2*>(&w[i >> 1]);
                out[i] = (i & 1) ? __high2float(h) : __low2float(h);
            } else {
                out[i] = __uint_as_float(w[i]);
            }
        }
    } else {
        const __half* hp = reinterpret_cast<const __half*>(p);
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = __half2float(__ldcs(hp + i));
    }
}

// GaussianShared.h:38-116 with the function constant SH_DEGREE = DEG
template <bool HALF, int DEG>
__device__ __forceinline__ V3 computeSHColor(const void* harmonics, uint32_t gid, V3 pos, V3 cam, uint32_t shComponents) {
    constexpr float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
    if (DEG == 0 || shComponents == 0) {
        float h[3];
        loadSH<HALF, 1>(harmonics, gid, h);
        return V3{h[0] * SH_C0, h[1] * SH_C0, h[2] * SH_C0};
    }
    constexpr int K = DEG == 1 ? 4 : (DEG == 2 ? 9 : 16);
    V3 dv{cam.x - pos.x, cam.y - pos.y, cam.z - pos.z};
    float len = sqrtf((dv.x * dv.x + dv.y * dv.y) + dv.z * dv.z);
    V3 dir{dv.x / len, dv.y / len, dv.z / len};
    float xx = dir.x * dir.x, yy = dir.y * dir.y, zz = dir.z * dir.z;
    float xy = dir.x * dir.y, yz = dir.y * dir.z, xz = dir.x * dir.z;
    float sh[16];
    sh[0] = SH_C0;
    sh[1] = -SH_C1 * dir.y;
    sh[2] = SH_C1 * dir.z;
    sh[3] = -SH_C1 * dir.x;
    if (DEG >= 2) {
        sh[4] = 1.0925484305920792f * xy;
        sh[5] = -1.0925484305920792f * yz;
        sh[6] = 0.31539156525252005f * ((2.0f * zz - xx) - yy);
        sh[7] = -1.0925484305920792f * xz;
        sh[8] = 0.5462742152960396f * (xx - yy);
    }
    if (DEG >= 3) {
        sh[9] = -0.5900435899266435f * dir.y * (3.0f * xx - yy);
        sh[10] = 2.890611442640554f * xy * dir.z;
        sh[11] = -0.4570457994644658f * dir.y * ((4.0f * zz - xx) - yy);
        sh[12] = 0.3731763325901154f * dir.z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
        sh[13] = -0.4570457994644658f * dir.x * ((4.0f * zz - xx) - yy);
        sh[14] = 1.445305721320277f * dir.z * (xx - yy);
        sh[15] = -0.5900435899266435f * dir.x * (xx - 3.0f * yy);
    }
    float h[3 * K];
    loadSH<HALF, K>(harmonics, gid, h);
    V3 color{0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < K; ++i) {
        color.x = color.x + h[i] * sh[i];
        color.y = color.y + h[K + i] * sh[i];
        color.z = color.z + h[2 * K + i] * sh[i];
    }
    return color;
}

__device__ __forceinline__ void writeCulled(const ProjectOut& o, uint32_t gid) {
    o.nTouched[gid] = 0;
    reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(0, -1, 0, -1);
}

// The per-frame zero region (frame state, prefix and status words) is cleared by the frame's first kernel when its grid is
// large enough to do it in a few stores per thread: one operation less on the stream, and nothing between the previous
// frame's blend and this kernel, so the two chain with programmatic dependent launch. Consumers are later kernels.
__device__ __forceinline__ void zeroFrameState(const ProjectOut& o) {
    if (!o.zeroBase) return;
    for (size_t i = (size_t)blockIdx.x * kProjThreads + threadIdx.x; i < o.zeroVecs; i += (size_t)gridDim.x * kProjThreads)
        o.zeroBase[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------- mono kernel
template <bool HALF, int DEG>
__global__ void __launch_bounds__(kProjThreads) project_cull_mono_kernel(const void* __restrict__ gaussians,
                                                                const void* __restrict__ harmonics,
                                                                const __grid_constant__ MonoCam cam, ProjectOut o) {
    pdlLaunchDependents();
    pdlWait();
    zeroFrameState(o);
    const uint32_t tile = blockIdx.x;
    const uint32_t N = cam.gaussianCount;
    const uint32_t numWarpTiles = (N + 31u) / 32u;
    const uint32_t warpTile = tile * (kProjThreads / 32) + (threadIdx.x >> 5);
    if (warpTile >= numWarpTiles) return;  // whole warp past the end
    const uint32_t lid = tile * kProjThreads + threadIdx.x;  // index into the (possibly sharded) input arrays
    const uint32_t gid = o.gidFirst + lid;            // global Gaussian id: what every output is keyed by
    const bool inRange = lid < N;

    __shared__ WarpTileWork s_work[kProjThreads / 32];
    uint32_t touched = 0, key = 0xFFFFFFFFu;
    // state carried from the per-lane projection to the warp-cooperative tile walk
    bool alive = false;
    QuantSplat q = {};
    int bMinTX = 0, bMaxTX = -1, bMinTY = 0, bMaxTY = -1;
    uint32_t nTiles = 0, sColor = 0;
    float keyDepth = 0.0f;
    __half sMeanX = __ushort_as_half((unsigned short)0), sMeanY = sMeanX, sDepth = sMeanX;
    if (inRange) {
        do {
            GaussianIn g = loadGaussian<HALF>(gaussians, lid);
            // (1) DFS.metal:63-69
            float maxScale = dmax(g.scale.x, dmax(g.scale.y, g.scale.z));
            if (maxScale < 0.0005f) break;
            V4 viewPos4 = mul44(cam.view, V4{g.pos.x, g.pos.y, g.pos.z, 1.0f});
            V4 clip = mul44(cam.proj, viewPos4);
            float depth = clip.w;
            if (!(clip.w > cam.nearPlane)) break;  // (2) DFS.metal:76-81
            if (depth > cam.farPlane) break;       // (3) DFS.metal:82-87
            float ndcX = clip.x / clip.w;
            float ndcY = clip.y / clip.w;
            float screenX = (ndcX + 1.0f) * 0.5f * cam.width;
            float screenY = (ndcY + 1.0f) * 0.5f * cam.height;
            if (g.opacity < kAlphaThreshold) break;  // (4) DFS.metal:93-99

            V4 quat = normalizeQuaternion(g.rot);
            M3 cov3d = buildCovariance3D(g.scale, quat);
            M2 cov2d = projectCovariance2D(cov3d, V3{viewPos4.x, viewPos4.y, viewPos4.z}, cam.view, cam.proj,
                                           cam.width, cam.height);
            cov2d = stabilizeCovariance2D(cov2d, cam.width, cam.height);
            float theta, sigma1, sigma2;
            if (!covarianceToThetaSigmas(cov2d, theta, sigma1, sigma2)) break;  // (5) DFS.metal:110-115
            if (3.0f * dmax(sigma1, sigma2) < 0.5f) break;                      // (6) DFS.metal:116-122
            {
                float a = cov2d.m00, b = 0.5f * (cov2d.m01 + cov2d.m10), d = cov2d.m11;
                float detCov = a * d - b * b;
                if (cullByTotalInk(g.opacity, detCov, depth, cam.nearPlane, cam.farPlane, kTotalInkThreshold)) break;  // (7)
            }
            float obbX, obbY;
            computeOBBExtents(cov2d, 3.0f, obbX, obbY);
            if (screenX + obbX < 0.0f || screenX - obbX > cam.width || screenY + obbY < 0.0f ||
                screenY - obbY > cam.height) break;  // (8) DFS.metal:131-137

            V3 color = computeSHColor<HALF, DEG>(harmonics, lid, g.pos, V3{cam.center[0], cam.center[1], cam.center[2]},
                                                 cam.shComponents);
            color.x = dmax(color.x + 0.5f, 0.0f);
            color.y = dmax(color.y + 0.5f, 0.0f);
            color.z = dmax(color.z + 0.5f, 0.0f);
            if (cam.inputIsSRGB > 0.5f) {
                color.x = srgbToLinearChannel(color.x);
                color.y = srgbToLinearChannel(color.y);
                color.z = srgbToLinearChannel(color.z);
            }

            // DFS.metal:143-154
            __half hMeanX = __float2half_rn(screenX), hMeanY = __float2half_rn(screenY);
            uint16_t thetaP = packThetaPi(theta);
            __half hS1 = __float2half_rn(sigma1), hS2 = __float2half_rn(sigma2), hDepth = __float2half_rn(depth);
            uint8_t cR = quantU8(color.x), cG = quantU8(color.y), cB = quantU8(color.z), cO = quantU8(g.opacity);
            uint4 rd;
            rd.x = (uint32_t)__half_as_ushort(hMeanX) | ((uint32_t)__half_as_ushort(hMeanY) << 16);
            rd.y = (uint32_t)thetaP | ((uint32_t)__half_as_ushort(hS1) << 16);
            rd.z = (uint32_t)__half_as_ushort(hS2) | ((uint32_t)__half_as_ushort(hDepth) << 16);
            rd.w = (uint32_t)cR | ((uint32_t)cG << 8) | ((uint32_t)cB << 16) | ((uint32_t)cO << 24);
            reinterpret_cast<uint4*>(o.renderData)[gid] = rd;

            TileBounds tb = computeTileBounds(screenX, screenY, obbX, obbY, cam.width, cam.height, (int)cam.tilesX,
                                              (int)cam.tilesY);

            // DFS.metal:166-205 runs on the quantised record (quirk Q3); the walk itself is warp-cooperative below
            q = makeQuantSplat(hMeanX, hMeanY, thetaP, hS1, hS2, cO);
            bMinTX = tb.minTX; bMaxTX = tb.maxTX; bMinTY = tb.minTY; bMaxTY = tb.maxTY;
            if (q.d2Cutoff >= 0.0f && tb.valid) nTiles = (uint32_t)((tb.maxTX - tb.minTX + 1) * (tb.maxTY - tb.minTY + 1));
            keyDepth = depth;
            sMeanX = hMeanX; sMeanY = hMeanY; sDepth = hDepth;
            sColor = (uint32_t)cR | ((uint32_t)cG << 8) | ((uint32_t)cB << 16) | ((uint32_t)cO << 24);
            alive = true;
        } while (false);
    }
    // exact ellipse-vs-tile count, 32 tiles of the warp's concatenated AABBs per step
    uint2 hitMask;
    const uint32_t cnt = warpCountTiles(s_work[threadIdx.x >> 5], nTiles, q, bMinTX, bMinTY, bMaxTX - bMinTX + 1, hitMask);
    if (inRange) {
        if (alive && cnt > 0) {  // cnt == 0 is exit (9), DFS.metal:207-212 (renderData stays written)
            // the pre-expanded blend record (conicFromThetaSigmas on the same quantised values)
            if (o.blendSplats)
                storeBlendSplat(o.blendSplats + gid, q, sMeanX, sMeanY, (uint8_t)sColor, (uint8_t)(sColor >> 8), (uint8_t)(sColor >> 16),
                                (uint8_t)(sColor >> 24), sDepth);
            reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(bMinTX, bMaxTX, bMinTY, bMaxTY);
            o.nTouched[gid] = cnt;
            o.hitMask[gid] = hitMask;
            touched = cnt;
            key = float_to_sortable_uint(keyDepth);
        } else {
            writeCulled(o, gid);
        }
    }
    if (inRange) o.preDepthKeys[gid] = key;  // compacted by compact_visible_kernel (no inter-warp wait in this kernel)
    recordKeyRange(o.keyRange, key, touched > 0u, warpTile);  // the depth sort's bucket plan needs the frame's key range
}

// ---------------------------------------------------------------- stereo
struct EyeResult {
    float screenX, screenY, theta, sigma1, sigma2, detCov, depth;
    int tb[4];
    bool visible;
};

// DFS.metal:249-339
__device__ __forceinline__ EyeResult projectToEye(V3 scenePos, V3 scale, V4 quat, const float* sceneTransform,
                                                  const float* view, const float* proj, float width, float height,
                                                  float nearPlane, float farPlane, int tilesX, int tilesY) {
    EyeResult r;
    r.visible = false;
    r.theta = 0.0f; r.sigma1 = 0.0f; r.sigma2 = 0.0f; r.detCov = 0.0f; r.screenX = 0.0f; r.screenY = 0.0f;
    r.tb[0] = 0; r.tb[1] = -1; r.tb[2] = 0; r.tb[3] = -1;
    V4 worldPos4 = mul44(sceneTransform, V4{scenePos.x, scenePos.y, scenePos.z, 1.0f});
    V4 viewPos4 = mul44(view, worldPos4);
    V4 clip = mul44(proj, viewPos4);
    float depth = clip.w;
    r.depth = depth;
    if (!(clip.w > nearPlane)) return r;
    if (depth > farPlane) return r;
    float ndcX = clip.x / clip.w, ndcY = clip.y / clip.w;
    r.screenX = (ndcX + 1.0f) * 0.5f * width;
    r.screenY = (ndcY + 1.0f) * 0.5f * height;
    float s0 = sceneTransform[0], s1 = sceneTransform[1], s2 = sceneTransform[2];
    float sceneScale = sqrtf((s0 * s0 + s1 * s1) + s2 * s2);
    M3 cov3d = buildCovariance3D(V3{scale.x * sceneScale, scale.y * sceneScale, scale.z * sceneScale}, quat);
    M2 cov2d = projectCovariance2D(cov3d, V3{viewPos4.x, viewPos4.y, viewPos4.z}, view, proj, width, height);
    cov2d = stabilizeCovariance2D(cov2d, width, height);
    float theta, sigma1, sigma2;
    if (!covarianceToThetaSigmas(cov2d, theta, sigma1, sigma2)) return r;
    r.theta = theta; r.sigma1 = sigma1; r.sigma2 = sigma2;
    float a = cov2d.m00, b = 0.5f * (cov2d.m01 + cov2d.m10), d = cov2d.m11;
    r.detCov = dmax(a * d - b * b, 0.0f);
    if (3.0f * dmax(sigma1, sigma2) < 0.5f) return r;
    float obbX, obbY;
    computeOBBExtents(cov2d, 3.0f, obbX, obbY);
    if (r.screenX + obbX < 0.0f || r.screenX - obbX > width || r.screenY + obbY < 0.0f || r.screenY - obbY > height)
        return r;
    TileBounds tb = computeTileBounds(r.screenX, r.screenY, obbX, obbY, width, height, tilesX, tilesY);
    r.tb[0] = tb.minTX; r.tb[1] = tb.maxTX; r.tb[2] = tb.minTY; r.tb[3] = tb.maxTY;
    r.visible = true;
    return r;
}

// GaussianShared.h:490-510
__device__ __forceinline__ void conicFromThetaSigmasF(float theta, float sigma1, float sigma2, float& A, float& B, float& C) {
    float s, c;
    dsincos(theta, s, c);
    float sig1 = dmax(sigma1, 1e-4f);
    float sig2 = dmax(sigma2, 1e-4f);
    float invVar1 = 1.0f / (sig1 * sig1);
    float invVar2 = 1.0f / (sig2 * sig2);
    float cc = c * c, ss = s * s, cs = c * s;
    A = cc * invVar1 + ss * invVar2;
    B = cs * (invVar1 - invVar2);
    C = ss * invVar1 + cc * invVar2;
}

template <bool HALF, int DEG>
__global__ void __launch_bounds__(kProjThreads) project_cull_stereo_kernel(const void* __restrict__ gaussians,
                                                                  const void* __restrict__ harmonics,
                                                                  const __grid_constant__ StereoCam cam, ProjectOut o) {
    pdlLaunchDependents();
    pdlWait();
    zeroFrameState(o);
    const uint32_t tile = blockIdx.x;
    const uint32_t N = cam.gaussianCount;
    const uint32_t numWarpTiles = (N + 31u) / 32u;
    const uint32_t warpTile = tile * (kProjThreads / 32) + (threadIdx.x >> 5);
    if (warpTile >= numWarpTiles) return;  // whole warp past the end
    const uint32_t lid = tile * kProjThreads + threadIdx.x;  // index into the (possibly sharded) input arrays
    const uint32_t gid = o.gidFirst + lid;            // global Gaussian id: what every output is keyed by
    const bool inRange = lid < N;
    uint32_t touched = 0, key = 0xFFFFFFFFu;
    if (inRange) {
        do {
            GaussianIn g = loadGaussian<HALF>(gaussians, lid);
            float maxScale = dmax(g.scale.x, dmax(g.scale.y, g.scale.z));
            if (maxScale < 0.0005f) break;                 // DFS.metal:359-365
            if (g.opacity < kAlphaThreshold) break;        // DFS.metal:369-375
            V4 quat = normalizeQuaternion(g.rot);
            EyeResult L = projectToEye(g.pos, g.scale, quat, cam.sceneTransform, cam.leftView, cam.leftProj, cam.width,
                                       cam.height, cam.nearPlane, cam.farPlane, (int)cam.tilesX, (int)cam.tilesY);
            EyeResult R = projectToEye(g.pos, g.scale, quat, cam.sceneTransform, cam.rightView, cam.rightProj, cam.width,
                                       cam.height, cam.nearPlane, cam.farPlane, (int)cam.tilesX, (int)cam.tilesY);
            if (!L.visible && !R.visible) break;           // DFS.metal:397-402
            float checkDepth = L.visible ? L.depth : R.depth;
            if (L.visible && R.visible) checkDepth = (L.depth + R.depth) * 0.5f;
            float detCov = L.visible ? L.detCov : R.detCov;
            if (L.visible && R.visible) detCov = dmax(L.detCov, R.detCov);
            if (cullByTotalInk(g.opacity, detCov, checkDepth, cam.nearPlane, cam.farPlane, kTotalInkThreshold)) break;
            V3 mid{(cam.leftCenter[0] + cam.rightCenter[0]) * 0.5f, (cam.leftCenter[1] + cam.rightCenter[1]) * 0.5f,
                   (cam.leftCenter[2] + cam.rightCenter[2]) * 0.5f};
            V3 color = computeSHColor<HALF, DEG>(harmonics, lid, g.pos, mid, cam.shComponents);
            color.x = dmax(color.x + 0.5f, 0.0f);
            color.y = dmax(color.y + 0.5f, 0.0f);
            color.z = dmax(color.z + 0.5f, 0.0f);
            if (cam.inputIsSRGB > 0.5f) {
                color.x = srgbToLinearChannel(color.x);
                color.y = srgbToLinearChannel(color.y);
                color.z = srgbToLinearChannel(color.z);
            }
            int ub[4];
            if (L.visible && R.visible) {
                ub[0] = min(L.tb[0], R.tb[0]); ub[1] = max(L.tb[1], R.tb[1]);
                ub[2] = min(L.tb[2], R.tb[2]); ub[3] = max(L.tb[3], R.tb[3]);
            } else if (L.visible) {
                ub[0] = L.tb[0]; ub[1] = L.tb[1]; ub[2] = L.tb[2]; ub[3] = L.tb[3];
            } else {
                ub[0] = R.tb[0]; ub[1] = R.tb[1]; ub[2] = R.tb[2]; ub[3] = R.tb[3];
            }
            int utx = max(ub[1] - ub[0] + 1, 0), uty = max(ub[3] - ub[2] + 1, 0);
            uint32_t cnt = (uint32_t)(utx * uty);
            if (cnt == 0) break;                           // DFS.metal:444-449

            GSMStereoTiledRenderData rd;
            const unsigned short negInf = 0xFC00u;         // half(-1e10f) = -inf, DFS.metal:461
            if (L.visible) {
                float A, B, C;
                conicFromThetaSigmasF(L.theta, L.sigma1, L.sigma2, A, B, C);
                rd.leftMeanX = __half_as_ushort(__float2half_rn(L.screenX));
                rd.leftMeanY = __half_as_ushort(__float2half_rn(L.screenY));
                rd.leftCxx = __half_as_ushort(__float2half_rn(A));
                rd.leftCyy = __half_as_ushort(__float2half_rn(C));
                rd.leftCxy2 = __half_as_ushort(__float2half_rn(2.0f * B));
                rd.leftDepth = __half_as_ushort(__float2half_rn(L.depth));
            } else {
                rd.leftMeanX = negInf; rd.leftMeanY = negInf; rd.leftCxx = 0; rd.leftCyy = 0; rd.leftCxy2 = 0; rd.leftDepth = 0;
            }
            if (R.visible) {
                float A, B, C;
                conicFromThetaSigmasF(R.theta, R.sigma1, R.sigma2, A, B, C);
                rd.rightMeanX = __half_as_ushort(__float2half_rn(R.screenX));
                rd.rightMeanY = __half_as_ushort(__float2half_rn(R.screenY));
                rd.rightCxx = __half_as_ushort(__float2half_rn(A));
                rd.rightCyy = __half_as_ushort(__float2half_rn(C));
                rd.rightCxy2 = __half_as_ushort(__float2half_rn(2.0f * B));
                rd.rightDepth = __half_as_ushort(__float2half_rn(R.depth));
            } else {
                rd.rightMeanX = negInf; rd.rightMeanY = negInf; rd.rightCxx = 0; rd.rightCyy = 0; rd.rightCxy2 = 0; rd.rightDepth = 0;
            }
            rd.colorR = quantU8(color.x); rd.colorG = quantU8(color.y); rd.colorB = quantU8(color.z);
            rd.opacity = quantU8(g.opacity);
            rd.centerDepth = __half_as_ushort(__float2half_rn(checkDepth));
            rd._pad0 = 0;
            uint4* dst = reinterpret_cast<uint4*>(o.renderData) + 2 * (size_t)gid;
            dst[0] = *reinterpret_cast<uint4*>(&rd);
            dst[1] = *(reinterpret_cast<uint4*>(&rd) + 1);
            reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(ub[0], ub[1], ub[2], ub[3]);
            o.nTouched[gid] = cnt;
            touched = cnt;
            key = float_to_sortable_uint(checkDepth);
        } while (false);
        if (touched == 0) writeCulled(o, gid);
    }
    if (inRange) o.preDepthKeys[gid] = key;  // compacted by compact_visible_kernel (no inter-warp wait in this kernel)
    recordKeyRange(o.keyRange, key, touched > 0u, warpTile);
}

// ---------------------------------------------------------------- stage 1.25: visibility compaction
// Replaces the 8-pass VisibilityCompactionEncoder (DFS.metal:518-621): one pass, 2048 gids per CTA, decoupled
// look-back carrying (visible count, sum of nTouched), so it also yields totalInstances (the atomic of
// DFS.metal:218). It is a separate kernel on purpose: fused into the projection kernel, every warp had to wait
// for all earlier warps' tile walks before it could retire (ncu r1_v3: 36 % of that kernel's instructions were
// look-back spins). Here the work per element is uniform, so the chain never stalls.
// DFS.metal:2184-2203: clamp the raw totals to the buffer capacities, radix-aligned padded counts, overflow flag
__device__ __forceinline__ void writeFrameHeader(GSMDepthFirstHeader* header, uint32_t v, uint32_t i, uint32_t maxGaussians,
                                                 uint32_t maxInstances) {
    uint32_t overflow = 0;
    if (v > maxGaussians) { v = maxGaussians; overflow = 1u; }
    if (i > maxInstances) { i = maxInstances; overflow = 1u; }
    header->visibleCount = v;
    header->totalInstances = i;
    header->paddedVisibleCount = ((v + kRadixAlignment - 1u) / kRadixAlignment) * kRadixAlignment;
    header->paddedInstanceCount = ((i + kRadixAlignment - 1u) / kRadixAlignment) * kRadixAlignment;
    header->overflow = overflow;
    header->padding0 = 0; header->padding1 = 0; header->padding2 = 0;
}

constexpr int kCompactItems = 8;
__global__ void __launch_bounds__(256) compact_visible_kernel(uint32_t N, ProjectOut o) {
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_touched[8];
    __shared__ uint32_t s_tile, s_baseVisible;
    __shared__ uint32_t s_hist[4][256];
    __shared__ uint32_t s_keyMin, s_fineShift;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    pdlLaunchDependents();
    for (int i = tid; i < 4 * 256; i += 256) (&s_hist[0][0])[i] = 0u;
    pdlWait();
    if (o.countPtr) N = min(ldAfterWait(o.countPtr), N);  // routed records: the count exists only on the device
    const uint32_t numTiles = (N + 256u * kCompactItems - 1u) / (256u * kCompactItems);
    if (numTiles == 0u) {  // nothing arrived: no last tile will write the header
        if (blockIdx.x == 0 && tid == 0) {
            o.fs->visibleCountRaw = 0u;
            o.fs->totalInstancesRaw = 0u;
            if (o.header) writeFrameHeader(o.header, 0u, 0u, o.maxOut, o.maxInstances);
        }
        return;
    }
    // the frame's key range (recorded by the projection kernel) fixes the fine bins a sample of the keys is counted into; the
    // depth sort's scatter kernel derives its bucket boundaries from that sample (bucketsort.cu)
    if (o.keyRange && warp == 0) {
        uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, o.keyRange->maxKey[lane]);
        uint32_t lo = __reduce_max_sync(0xFFFFFFFFu, o.keyRange->maxInvKey[lane]);
        if (lane == 0) {
            const bool valid = (hi | lo) != 0u && ~lo <= hi;
            s_keyMin = valid ? ~lo : 0u;
            const uint32_t span = valid ? hi - ~lo : 0xFFFFFFFFu;
            const int bits = 32 - __clz(span);  // 0 for span 0
            s_fineShift = bits > 13 ? (uint32_t)(bits - 13) : 0u;   // (span >> shift) < kDepthFineBins
        }
    }
    __syncthreads();
    const bool planning = o.keyRange != nullptr;
    const uint32_t keyMin = planning ? s_keyMin : 0u, fineShift = planning ? s_fineShift : 0u;
    while (true) {
        if (tid == 0) s_tile = atomicAdd(&o.fs->ticketProject, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t first = tile * 256u * kCompactItems + tid * kCompactItems;  // blocked: 8 consecutive gids per thread
        uint32_t nt[kCompactItems];
        uint32_t cnt = 0, tsum = 0;
#pragma unroll
        for (int i = 0; i < kCompactItems; ++i) {
            const uint32_t l = first + i;
            nt[i] = (l < N) ? (o.recTouched ? o.recTouched[l] : o.nTouched[o.gidFirst + l]) : 0u;
            cnt += nt[i] > 0u ? 1u : 0u;
            tsum += nt[i];
        }
        // the keys of the visible elements do not depend on the prefix: their loads are in flight under the scan and the wait for
        // the predecessors' aggregates instead of behind them
        uint32_t keys[kCompactItems];
#pragma unroll
        for (int i = 0; i < kCompactItems; ++i)
            keys[i] = nt[i] > 0u ? (o.recTouched ? o.recKey[first + i] : o.preDepthKeys[o.gidFirst + first + i]) : 0u;
        uint32_t blockVisible;
        const uint32_t excl = block_exclusive_scan_256(cnt, s_scan, blockVisible);
        for (int off = 16; off > 0; off >>= 1) tsum += __shfl_xor_sync(0xFFFFFFFFu, tsum, off);
        if (lane == 0) s_touched[warp] = tsum;
        __syncthreads();
        if (warp == 0) {
            uint32_t bt = (lane < 8) ? s_touched[lane] : 0u;
            for (int off = 4; off > 0; off >>= 1) bt += __shfl_xor_sync(0xFFFFFFFFu, bt, off);
            bt = __shfl_sync(0xFFFFFFFFu, bt, 0);
            // publish, then sum the predecessors directly: no tile depends on another tile's resolve
            if (lane == 0) prefixPublish2(o.status, o.statusGroups, tile, blockVisible, bt);
            uint32_t ev;
            unsigned long long et;
            prefixResolve2(o.status, o.statusGroups, tile, ev, et);
            if (lane == 0) {
                s_baseVisible = ev;
                if (tile == numTiles - 1) {
                    o.fs->visibleCountRaw = ev + blockVisible;            // DFS.metal:618-620
                    o.fs->totalInstancesRaw = (uint32_t)(et + bt);        // DFS.metal:218 (u32, wraps)
                    // the frame header (prepareIndirectDispatchKernel's role, DFS.metal:2184-2203) rides on the same thread
                    if (o.header) writeFrameHeader(o.header, ev + blockVisible, (uint32_t)(et + bt), o.maxOut, o.maxInstances);
                }
            }
        }
        __syncthreads();
        uint32_t dst = s_baseVisible + excl;
#pragma unroll
        for (int i = 0; i < kCompactItems; ++i) {
            if (nt[i] > 0u) {
                if (dst < o.maxOut) {  // DFS.metal:605
                    // projection: element l is Gaussian gidFirst + l; strip ingest: element l is gathered record l
                    const uint32_t gid = o.recTouched ? o.recGid[first + i] : o.gidFirst + first + i;
                    uint32_t key = keys[i];
                    if (o.depthKey16) {  // DFS.metal:607-612; key is float_to_sortable_uint of a depth > 0
                        uint32_t bits = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
                        key = (uint32_t)(__half_as_ushort(__float2half_rn(__uint_as_float(bits))) ^ 0x8000u);
                    }
                    o.depthKeys[dst] = key;
                    o.primitiveIndices[dst] = (int32_t)gid;
                    for (uint32_t p = 0; p < o.depthPasses; ++p) atomicAdd(&s_hist[p][(key >> (8u * p)) & 0xFFu], 1u);  // the sort's digit histograms
                    if (planning && (dst % kDepthSampleStride) == 0u)   // a sample fixes the bucket boundaries (RED; one per key was 24 us)
                        atomicAdd(&o.fs->fineHist[min((key - keyMin) >> fineShift, kDepthFineBins - 1u)], 1u);
                }
                dst++;
            }
        }
        __syncthreads();
    }
    for (uint32_t i = tid; i < o.depthPasses * 256u; i += 256u) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&o.depthHist[i], v);
    }
}

cudaError_t launchCompactVisible(cudaStream_t s, uint32_t N, const ProjectOut& o, int numSMs) {
    if (N == 0) return cudaSuccess;
    uint32_t tiles = (N + 256u * kCompactItems - 1u) / (256u * kCompactItems);
    uint32_t grid = tiles < (uint32_t)numSMs * 8u ? tiles : (uint32_t)numSMs * 8u;
    launchChained(compact_visible_kernel, grid, 256, s, N, o);
    return cudaGetLastError();
}

const void* kernel_image_probe() { return (const void*)compact_visible_kernel; }

// ---------------------------------------------------------------- GlobalRenderer
// globalProjectCull (GlobalShaders.metal:19-125): the same shared helpers as the DepthFirst projection, but
// ndcToScreenCentered, no far-plane exit, no tile count; 32 x 16-pixel tiles of the renderer's LIMITS.
__device__ __forceinline__ TileBounds computeTileBoundsWH(float sx, float sy, float ex, float ey, float width, float height, int tileW,
                                                          int tileH, int tilesX, int tilesY) {
    TileBounds r;
    float xmin = sx - ex, xmax = sx + ex, ymin = sy - ey, ymax = sy + ey;
    float maxW = width - 1.0f, maxH = height - 1.0f;
    xmin = dclamp(xmin, 0.0f, maxW);
    xmax = dclamp(xmax, 0.0f, maxW);
    ymin = dclamp(ymin, 0.0f, maxH);
    ymax = dclamp(ymax, 0.0f, maxH);
    r.minTX = (int)floorf(xmin / (float)tileW);
    r.maxTX = (int)ceilf(xmax / (float)tileW) - 1;
    r.minTY = (int)floorf(ymin / (float)tileH);
    r.maxTY = (int)ceilf(ymax / (float)tileH) - 1;
    r.minTX = max(r.minTX, 0);
    r.minTY = max(r.minTY, 0);
    r.maxTX = min(r.maxTX, tilesX - 1);
    r.maxTY = min(r.maxTY, tilesY - 1);
    r.valid = (r.minTX <= r.maxTX && r.minTY <= r.maxTY);
    return r;
}

template <bool HALF, int DEG>
__global__ void __launch_bounds__(kProjThreads) global_project_cull_kernel(const void* __restrict__ gaussians, const void* __restrict__ harmonics,
                                                                           const __grid_constant__ MonoCam cam, GlobalFrame f) {
    const uint32_t gid = blockIdx.x * kProjThreads + threadIdx.x;
    if (gid >= cam.gaussianCount) return;
    int4 rect = make_int4(0, -1, 0, -1);
    do {
        GaussianIn g = loadGaussian<HALF>(gaussians, gid);
        float maxScale = dmax(g.scale.x, dmax(g.scale.y, g.scale.z));
        if (maxScale < 0.0005f) break;
        V4 viewPos4 = mul44(cam.view, V4{g.pos.x, g.pos.y, g.pos.z, 1.0f});
        V4 clip = mul44(cam.proj, viewPos4);
        float depth = clip.w;
        if (!(clip.w > cam.nearPlane)) break;
        float ndcX = clip.x / clip.w, ndcY = clip.y / clip.w;
        float screenX = ((ndcX + 1.0f) * cam.width - 1.0f) * 0.5f;    // ndcToScreenCentered, GaussianShared.h:184-189
        float screenY = ((ndcY + 1.0f) * cam.height - 1.0f) * 0.5f;
        if (g.opacity < kAlphaThreshold) break;
        V4 quat = normalizeQuaternion(g.rot);
        M3 cov3d = buildCovariance3D(g.scale, quat);
        M2 cov2d = projectCovariance2D(cov3d, V3{viewPos4.x, viewPos4.y, viewPos4.z}, cam.view, cam.proj, cam.width, cam.height);
        cov2d = stabilizeCovariance2D(cov2d, cam.width, cam.height);
        float theta, sigma1, sigma2;
        if (!covarianceToThetaSigmas(cov2d, theta, sigma1, sigma2)) break;
        if (3.0f * dmax(sigma1, sigma2) < 0.5f) break;
        {
            float a = cov2d.m00, b = 0.5f * (cov2d.m01 + cov2d.m10), d = cov2d.m11;
            float detCov = a * d - b * b;
            if (cullByTotalInk(g.opacity, detCov, depth, cam.nearPlane, cam.farPlane, kTotalInkThreshold)) break;
        }
        float obbX, obbY;
        computeOBBExtents(cov2d, 3.0f, obbX, obbY);
        if (screenX + obbX < 0.0f || screenX - obbX > cam.width || screenY + obbY < 0.0f || screenY - obbY > cam.height) break;
        V3 color = computeSHColor<HALF, DEG>(harmonics, gid, g.pos, V3{cam.center[0], cam.center[1], cam.center[2]}, cam.shComponents);
        color.x = dmax(color.x + 0.5f, 0.0f);
        color.y = dmax(color.y + 0.5f, 0.0f);
        color.z = dmax(color.z + 0.5f, 0.0f);
        if (cam.inputIsSRGB > 0.5f) {
            color.x = srgbToLinearChannel(color.x);
            color.y = srgbToLinearChannel(color.y);
            color.z = srgbToLinearChannel(color.z);
        }
        __half hMeanX = __float2half_rn(screenX), hMeanY = __float2half_rn(screenY);
        uint16_t thetaP = packThetaPi(theta);
        __half hS1 = __float2half_rn(sigma1), hS2 = __float2half_rn(sigma2), hDepth = __float2half_rn(depth);
        uint4 rd;
        rd.x = (uint32_t)__half_as_ushort(hMeanX) | ((uint32_t)__half_as_ushort(hMeanY) << 16);
        rd.y = (uint32_t)thetaP | ((uint32_t)__half_as_ushort(hS1) << 16);
        rd.z = (uint32_t)__half_as_ushort(hS2) | ((uint32_t)__half_as_ushort(hDepth) << 16);
        rd.w = (uint32_t)quantU8(color.x) | ((uint32_t)quantU8(color.y) << 8) | ((uint32_t)quantU8(color.z) << 16) |
               ((uint32_t)quantU8(g.opacity) << 24);
        f.renderData[gid] = rd;
        TileBounds tb = computeTileBoundsWH(screenX, screenY, obbX, obbY, cam.width, cam.height, (int)f.tileW, (int)f.tileH,
                                            (int)f.tilesX, (int)f.tilesY);
        rect = make_int4(tb.minTX, tb.maxTX, tb.minTY, tb.maxTY);
    } while (false);
    f.bounds[gid] = rect;
    f.flags[gid] = (rect.x <= rect.y && rect.z <= rect.w) ? 1u : 0u;   // markVisibilityKernel, GlobalShaders.metal:169-181
}

// GaussianShared.h:595-645 -- gaussianComputePower / gaussianIntersectsTile, opacity passed as the BYTE value (GlobalShaders.metal:589)
__device__ __forceinline__ bool globalSegmentHitsEllipse(float a, float b, float c, float d, float l, float r) {
    float delta = b * b - 4.0f * a * c;
    float t1 = (l - d) * (2.0f * a) + b;
    float t2 = (r - d) * (2.0f * a) + b;
    return delta >= 0.0f && (t1 <= 0.0f || t1 * t1 <= delta) && (t2 >= 0.0f || t2 * t2 <= delta);
}
__device__ __forceinline__ bool globalIntersectsTile(int minX, int minY, int maxX, int maxY, float cx, float cy, float conicX, float conicY,
                                                     float conicZ, float power) {
    if (cx >= (float)minX && cx <= (float)maxX && cy >= (float)minY && cy <= (float)maxY) return true;
    float w = 2.0f * power;
    float dx, dy, a, b, c;
    if (cx * 2.0f < (float)(minX + maxX)) dx = cx - (float)minX; else dx = cx - (float)maxX;
    a = conicZ;
    b = -2.0f * conicY * dx;
    c = conicX * dx * dx - w;
    if (globalSegmentHitsEllipse(a, b, c, cy, (float)minY, (float)maxY)) return true;
    if (cy * 2.0f < (float)(minY + maxY)) dy = cy - (float)minY; else dy = cy - (float)maxY;
    a = conicX;
    b = -2.0f * conicY * dy;
    c = conicZ * dy * dy - w;
    return globalSegmentHitsEllipse(a, b, c, cx, (float)minX, (float)maxX);
}

// tileCountIndirectKernel / tileScatterIndirectKernel + computeSortKeysKernel (GlobalShaders.metal:563-680, :267-295): EMIT false
// counts the tiles of visible Gaussian i, EMIT true stores key [tile:16][half depth ^ 0x8000:16] and the Gaussian's index from
// `writePos` on, bounded by maxAssignments per store. Returns the count.
template <bool EMIT>
__device__ __forceinline__ uint32_t globalWalkTiles(const GlobalFrame& f, uint32_t g, uint32_t writePos, uint32_t* sHist = nullptr) {
    const int4 rect = f.bounds[g];
    if (rect.x > rect.y || rect.z > rect.w) return 0u;
    const uint4 rd = f.renderData[g];
    const float alpha = (float)(rd.w >> 24);
    if (alpha < 1e-4f) return 0u;
    const float cx = __half2float(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)));
    const float cy = __half2float(__ushort_as_half((unsigned short)(rd.x >> 16)));
    const float theta = (float)(rd.y & 0xFFFFu) * GSM_THETA_UNPACK;
    float A, B, C;
    conicFromThetaSigmasF(theta, __half2float(__ushort_as_half((unsigned short)(rd.y >> 16))),
                          __half2float(__ushort_as_half((unsigned short)(rd.z & 0xFFFFu))), A, B, C);
    if (!EMIT) {
        // the render's record (globalRender, GlobalShaders.metal:1036-1187: conic from the quantised theta / sigmas, opacity and
        // colour bytes over 255, all rounded to half): a function of the Gaussian alone, so it is computed here once instead of
        // once per (tile, Gaussian) in the render
        BlendSplat bs;
        bs.mean = *reinterpret_cast<const __half2*>(&rd.x);
        bs.cxx_cyy = __floats2half2_rn(A, C);
        bs.cxy2_op = __halves2half2(__float2half_rn(2.0f * B), u8_over_255h((uint8_t)(rd.w >> 24)));
        bs.rg = __halves2half2(u8_over_255h((uint8_t)(rd.w & 0xFFu)), u8_over_255h((uint8_t)((rd.w >> 8) & 0xFFu)));
        bs.b_depth = __halves2half2(u8_over_255h((uint8_t)((rd.w >> 16) & 0xFFu)), __ushort_as_half((unsigned short)(rd.z >> 16)));
        bs.valid = 1u; bs._pad[0] = 0u; bs._pad[1] = 0u;
        uint4* d = reinterpret_cast<uint4*>(f.blendSplats + g);
        d[0] = *reinterpret_cast<uint4*>(&bs);
        d[1] = *(reinterpret_cast<uint4*>(&bs) + 1);
    }
    const float LN2 = 0.693147180559945f;
    const float power = LN2 * 8.0f + LN2 * (dlog(dmax(alpha, 1e-6f)) * 1.44269504088896341f);
    const uint32_t depthBits = ((rd.z >> 16) ^ 0x8000u) & 0xFFFFu;
    // AABBs of at most 64 tiles (almost all): the count pass leaves a 64-bit hit mask, bit k = tile k of the box in row-major order,
    // and the scatter pass replays it instead of running the ellipse test a second time (as the DepthFirst expansion does)
    const int bw = rect.y - rect.x + 1, bh = rect.w - rect.z + 1;
    const bool masked = bw * bh <= 64;
    uint32_t n = 0u;
    if (EMIT && masked) {
        const uint2 m = f.hitMask[g];
        unsigned long long bits = ((unsigned long long)m.y << 32) | m.x;
        while (bits) {
            const int k = __ffsll((long long)bits) - 1;
            bits &= bits - 1ull;
            if (writePos >= f.maxAssignments) break;   // every later store would be dropped too
            const int ty = rect.z + k / bw, tx = rect.x + k % bw;
            const uint32_t key = ((uint32_t)(ty * (int)f.tilesX + tx) << 16) | depthBits;
            f.sortKeys[writePos] = key;
            f.sortedIndices[writePos] = (int32_t)g;
            if (sHist) {
#pragma unroll
                for (uint32_t p = 0; p < 4u; ++p) atomicAdd(&sHist[p * 256u + ((key >> (8u * p)) & 0xFFu)], 1u);
            }
            writePos++;
            n++;
        }
        return n;
    }
    unsigned long long mask = 0ull;
    for (int ty = rect.z; ty <= rect.w; ++ty)
        for (int tx = rect.x; tx <= rect.y; ++tx) {
            const int px0 = tx * (int)f.tileW, py0 = ty * (int)f.tileH;
            if (globalIntersectsTile(px0, py0, px0 + (int)f.tileW - 1, py0 + (int)f.tileH - 1, cx, cy, A, B, C, power)) {
                if (!EMIT) {
                    n++;
                    if (masked) mask |= 1ull << ((ty - rect.z) * bw + (tx - rect.x));
                } else if (writePos < f.maxAssignments) {
                    const uint32_t key = ((uint32_t)(ty * (int)f.tilesX + tx) << 16) | depthBits;
                    f.sortKeys[writePos] = key;
                    f.sortedIndices[writePos] = (int32_t)g;
                    if (sHist) {   // the sort's four digit histograms, of exactly the keys that are stored
#pragma unroll
                        for (uint32_t p = 0; p < 4u; ++p) atomicAdd(&sHist[p * 256u + ((key >> (8u * p)) & 0xFFu)], 1u);
                    }
                    writePos++;
                    n++;
                }
            }
        }
    if (!EMIT && masked) f.hitMask[g] = make_uint2((uint32_t)mask, (uint32_t)(mask >> 32));
    return n;
}

__global__ void __launch_bounds__(256) global_tile_count_kernel(GlobalFrame f) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < f.capGaussians) f.counts[i] = i < f.header->visibleCount ? globalWalkTiles<false>(f, f.visibleIndices[i], 0u) : 0u;
}
__global__ void __launch_bounds__(256) global_tile_scatter_kernel(GlobalFrame f, uint32_t* __restrict__ sortHist) {
    __shared__ uint32_t s_hist[4 * 256];
    for (int k = threadIdx.x; k < 4 * 256; k += 256) s_hist[k] = 0u;
    __syncthreads();
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < f.header->visibleCount) globalWalkTiles<true>(f, f.visibleIndices[i], f.offsets[i], s_hist);
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * 256; k += 256) {
        const uint32_t v = s_hist[k];
        if (v) atomicAdd(&sortHist[k], v);
    }
}

// ---------------------------------------------------------------- launchers
template <bool HALF>
static cudaError_t launchMono(int deg, dim3 grid, cudaStream_t s, const void* g, const void* h, const MonoCam& cam,
                              const ProjectOut& o) {
    switch (deg) {
        case 0: launchChained(project_cull_mono_kernel<HALF, 0>, grid, kProjThreads, s, g, h, cam, o); break;
        case 1: launchChained(project_cull_mono_kernel<HALF, 1>, grid, kProjThreads, s, g, h, cam, o); break;
        case 2: launchChained(project_cull_mono_kernel<HALF, 2>, grid, kProjThreads, s, g, h, cam, o); break;
        default: launchChained(project_cull_mono_kernel<HALF, 3>, grid, kProjThreads, s, g, h, cam, o); break;
    }
    return cudaGetLastError();
}
template <bool HALF>
static cudaError_t launchStereo(int deg, dim3 grid, cudaStream_t s, const void* g, const void* h, const StereoCam& cam,
                                const ProjectOut& o) {
    switch (deg) {
        case 0: launchChained(project_cull_stereo_kernel<HALF, 0>, grid, kProjThreads, s, g, h, cam, o); break;
        case 1: launchChained(project_cull_stereo_kernel<HALF, 1>, grid, kProjThreads, s, g, h, cam, o); break;
        case 2: launchChained(project_cull_stereo_kernel<HALF, 2>, grid, kProjThreads, s, g, h, cam, o); break;
        default: launchChained(project_cull_stereo_kernel<HALF, 3>, grid, kProjThreads, s, g, h, cam, o); break;
    }
    return cudaGetLastError();
}

// DepthFirstProjectCullEncoder.swift:13-20
int shDegreeFromComponents(uint32_t n) { return n <= 1 ? 0 : (n <= 4 ? 1 : (n <= 9 ? 2 : 3)); }

cudaError_t launchProjectMono(cudaStream_t s, bool halfInput, const void* g, const void* h, const MonoCam& cam,
                              const ProjectOut& o) {
    dim3 grid((cam.gaussianCount + kProjThreads - 1) / kProjThreads);
    int deg = shDegreeFromComponents(cam.shComponents);
    return halfInput ? launchMono<true>(deg, grid, s, g, h, cam, o) : launchMono<false>(deg, grid, s, g, h, cam, o);
}
cudaError_t launchProjectStereo(cudaStream_t s, bool halfInput, const void* g, const void* h, const StereoCam& cam,
                                const ProjectOut& o) {
    dim3 grid((cam.gaussianCount + kProjThreads - 1) / kProjThreads);
    int deg = shDegreeFromComponents(cam.shComponents);
    return halfInput ? launchStereo<true>(deg, grid, s, g, h, cam, o) : launchStereo<false>(deg, grid, s, g, h, cam, o);
}

cudaError_t launchGlobalProject(cudaStream_t s, bool halfInput, const void* g, const void* h, const MonoCam& cam, const GlobalFrame& f) {
    const dim3 grid((cam.gaussianCount + kProjThreads - 1) / kProjThreads);
    const int deg = shDegreeFromComponents(cam.shComponents);
#define GSM_GP(H, D) global_project_cull_kernel<H, D><<<grid, kProjThreads, 0, s>>>(g, h, cam, f)
    if (halfInput) { switch (deg) { case 0: GSM_GP(true, 0); break; case 1: GSM_GP(true, 1); break; case 2: GSM_GP(true, 2); break; default: GSM_GP(true, 3); } }
    else { switch (deg) { case 0: GSM_GP(false, 0); break; case 1: GSM_GP(false, 1); break; case 2: GSM_GP(false, 2); break; default: GSM_GP(false, 3); } }
#undef GSM_GP
    return cudaGetLastError();
}
cudaError_t launchGlobalTileCount(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount) {
    global_tile_count_kernel<<<(gaussianCount + 255u) / 256u, 256, 0, s>>>(f);
    return cudaGetLastError();
}
cudaError_t launchGlobalTileScatter(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount, uint32_t* sortHist) {
    global_tile_scatter_kernel<<<(gaussianCount + 255u) / 256u, 256, 0, s>>>(f, sortHist);
    return cudaGetLastError();
}
}  // namespace gsm
