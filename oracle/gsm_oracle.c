/*
 * gsm_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see gsm_oracle.h).
 *
 * Every function cites the reference lines it restates. Evaluation order is fixed here
 * (left-to-right, one rounding per written operation) because the reference's
 * -ffast-math build does not define one; the CUDA path follows the same order.
 * Multithreaded with OpenMP per stage.
 */
#include "gsm_oracle.h"
#include "gsmo_math.h"

#include <omp.h>
#include <stdlib.h>
#include <time.h>

int gsmo_num_threads(void) { return omp_get_max_threads(); }
void gsmo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------- probes */
void gsmo_probe_sincos(const float* x, float* s, float* c, int n) { for (int i = 0; i < n; ++i) gsmo_sincos(x[i], &s[i], &c[i]); }
void gsmo_probe_log(const float* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = gsmo_log(x[i]); }
void gsmo_probe_atan2(const float* y, const float* x, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = gsmo_atan2(y[i], x[i]); }
void gsmo_probe_powr(const float* x, float yexp, float* r, int n) { for (int i = 0; i < n; ++i) r[i] = gsmo_powr(x[i], yexp); }
void gsmo_probe_hexp(const gsmo_half* x, gsmo_half* y, int n) { for (int i = 0; i < n; ++i) y[i] = gsmo_hexp(x[i]); }
void gsmo_probe_f2h(const float* x, gsmo_half* y, int n) { for (int i = 0; i < n; ++i) y[i] = gsmo_f2h(x[i]); }
void gsmo_probe_minmax(const float* a, const float* b, float* mn, float* mx, int n) { for (int i = 0; i < n; ++i) { mn[i] = gsmo_fmin(a[i], b[i]); mx[i] = gsmo_fmax(a[i], b[i]); } }
void gsmo_probe_hfma(const gsmo_half* a, const gsmo_half* b, const gsmo_half* c, gsmo_half* r, int n) { for (int i = 0; i < n; ++i) r[i] = gsmo_hfma(a[i], b[i], c[i]); }
void gsmo_probe_h2f(const gsmo_half* x, float* y, int n) { for (int i = 0; i < n; ++i) y[i] = gsmo_h2f(x[i]); }

/* ---------------------------------------------------------------- small linear algebra */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;
typedef struct { v3 c0, c1, c2; } m3; /* columns */

/* float4x4 * float4: sum of column*component, left to right (DFS.metal:72-73) */
static v4 mul44(const float* m, v4 v) {
    v4 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
    r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
    return r;
}
/* float3x3 * float3 */
static v3 mul33v(m3 a, v3 v) {
    v3 r;
    r.x = (a.c0.x * v.x + a.c1.x * v.y) + a.c2.x * v.z;
    r.y = (a.c0.y * v.x + a.c1.y * v.y) + a.c2.y * v.z;
    r.z = (a.c0.z * v.x + a.c1.z * v.y) + a.c2.z * v.z;
    return r;
}
static m3 mul33(m3 a, m3 b) {
    m3 r;
    r.c0 = mul33v(a, b.c0);
    r.c1 = mul33v(a, b.c1);
    r.c2 = mul33v(a, b.c2);
    return r;
}
static m3 transpose33(m3 a) {
    m3 r;
    r.c0.x = a.c0.x; r.c0.y = a.c1.x; r.c0.z = a.c2.x;
    r.c1.x = a.c0.y; r.c1.y = a.c1.y; r.c1.z = a.c2.y;
    r.c2.x = a.c0.z; r.c2.y = a.c1.z; r.c2.z = a.c2.z;
    return r;
}

/* 2x2 covariance, cov[col][row] as in MSL float2x2 */
typedef struct { float m00, m01, m10, m11; } m2;

/* ---------------------------------------------------------------- GaussianShared.h helpers */

/* GShared.h:289-295 */
static v4 normalizeQuaternion(v4 q) {
    float d = ((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w;
    float norm = sqrtf(gsmo_fmax(d, 1e-8f));
    if (norm < 1e-8f) { v4 r = {1.0f, 0.0f, 0.0f, 0.0f}; return r; }
    v4 r = {q.x / norm, q.y / norm, q.z / norm, q.w / norm};
    return r;
}

/* GShared.h:297-305 + matrixFromRows :141-147 */
static m3 quaternionToMatrix(v4 q) {
    float x = q.x, y = q.y, z = q.z, r = q.w;
    float xx = x * x, yy = y * y, zz = z * z;
    float xy = x * y, xz = x * z, yz = y * z;
    v3 row0 = {1.0f - 2.0f * (yy + zz), 2.0f * (xy - r * z), 2.0f * (xz + r * y)};
    v3 row1 = {2.0f * (xy + r * z), 1.0f - 2.0f * (xx + zz), 2.0f * (yz - r * x)};
    v3 row2 = {2.0f * (xz - r * y), 2.0f * (yz + r * x), 1.0f - 2.0f * (xx + yy)};
    m3 m;
    m.c0.x = row0.x; m.c0.y = row1.x; m.c0.z = row2.x;
    m.c1.x = row0.y; m.c1.y = row1.y; m.c1.z = row2.y;
    m.c2.x = row0.z; m.c2.y = row1.z; m.c2.z = row2.z;
    return m;
}

/* GShared.h:307-324 (normalises the quaternion a second time, quirk Q1) */
static m3 buildCovariance3D(v3 scale, v4 quat) {
    v4 q = normalizeQuaternion(quat);
    m3 R = quaternionToMatrix(q);
    v3 RS0 = {R.c0.x * scale.x, R.c0.y * scale.x, R.c0.z * scale.x};
    v3 RS1 = {R.c1.x * scale.y, R.c1.y * scale.y, R.c1.z * scale.y};
    v3 RS2 = {R.c2.x * scale.z, R.c2.y * scale.z, R.c2.z * scale.z};
    m3 c;
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

/* GShared.h:326-388 */
static m2 projectCovariance2D(m3 cov3d, v3 viewPos, const float* view, const float* proj,
                              float width, float height) {
    float absZ = fabsf(viewPos.z);
    float signZ = (viewPos.z >= 0.0f) ? 1.0f : -1.0f;
    float safeAbsZ = gsmo_fmax(absZ, 1e-4f);
    float invAbsZ = 1.0f / safeAbsZ;
    float invAbsZ2 = invAbsZ * invAbsZ;

    float tanHalfFovX = 1.0f / gsmo_fmax(fabsf(proj[0]), 1e-4f);
    float tanHalfFovY = 1.0f / gsmo_fmax(fabsf(proj[5]), 1e-4f);
    float limX = 1.3f * tanHalfFovX;
    float limY = 1.3f * tanHalfFovY;

    float tx = viewPos.x * invAbsZ;
    float ty = viewPos.y * invAbsZ;
    float xClamped = gsmo_clamp(tx, -limX, limX) * safeAbsZ;
    float yClamped = gsmo_clamp(ty, -limY, limY) * safeAbsZ;

    float focalX = width * fabsf(proj[0]) * 0.5f;
    float focalY = height * fabsf(proj[5]) * 0.5f;

    m3 J;
    J.c0.x = focalX * invAbsZ; J.c0.y = 0.0f; J.c0.z = 0.0f;
    J.c1.x = 0.0f; J.c1.y = focalY * invAbsZ; J.c1.z = 0.0f;
    J.c2.x = -focalX * xClamped * signZ * invAbsZ2;
    J.c2.y = -focalY * yClamped * signZ * invAbsZ2;
    J.c2.z = 0.0f;

    m3 W;
    W.c0.x = view[0]; W.c0.y = view[1]; W.c0.z = view[2];
    W.c1.x = view[4]; W.c1.y = view[5]; W.c1.z = view[6];
    W.c2.x = view[8]; W.c2.y = view[9]; W.c2.z = view[10];

    m3 T = mul33(J, W);
    m3 covFull = mul33(mul33(T, cov3d), transpose33(T));

    m2 c;
    c.m00 = covFull.c0.x + 0.3f;
    c.m01 = covFull.c0.y;
    c.m10 = covFull.c1.x;
    c.m11 = covFull.c1.y + 0.3f;
    return c;
}

/* GShared.h:655-714 */
static m2 stabilizeCovariance2D(m2 cov, float width, float height) {
    float maxCond = 256.0f * 256.0f;
    float maxDim = gsmo_fmax(width, height);
    float maxExtentPx = maxDim * 2.0f;
    float maxEig = maxExtentPx / 3.0f;
    maxEig = maxEig * maxEig;

    float a = cov.m00;
    float b = 0.5f * (cov.m01 + cov.m10);
    float d = cov.m11;
    if (!gsmo_isfinite(a) || !gsmo_isfinite(b) || !gsmo_isfinite(d)) {
        m2 id = {1.0f, 0.0f, 0.0f, 1.0f};
        return id;
    }
    a = gsmo_fmax(a, 1e-4f);
    d = gsmo_fmax(d, 1e-4f);
    float det = a * d - b * b;
    if (!gsmo_isfinite(det) || det < 1e-8f) {
        float bump = (1e-8f - det) + 1e-4f;
        a = a + bump;
        d = d + bump;
        det = a * d - b * b;
    }
    float mid = 0.5f * (a + d);
    float disc = gsmo_fmax(mid * mid - det, 0.0f);
    float sqrtDisc = sqrtf(disc);
    float lambda1 = mid + sqrtDisc;
    float lambda2 = gsmo_fmax(mid - sqrtDisc, 1e-4f);

    float v1x, v1y;
    if (fabsf(b) > 1e-8f) {
        float vx = b;
        float vy = lambda1 - a;
        float vlen = sqrtf(vx * vx + vy * vy);
        float dn = gsmo_fmax(vlen, 1e-8f);
        v1x = vx / dn;
        v1y = vy / dn;
    } else if (a >= d) {
        v1x = 1.0f; v1y = 0.0f;
    } else {
        v1x = 0.0f; v1y = 1.0f;
    }
    float v2x = v1y, v2y = -v1x;

    lambda1 = gsmo_fmin(lambda1, maxEig);
    lambda2 = gsmo_fmax(lambda2, lambda1 / maxCond);

    m2 o;
    o.m00 = lambda1 * (v1x * v1x) + lambda2 * (v2x * v2x);
    o.m01 = lambda1 * (v1x * v1y) + lambda2 * (v2x * v2y);
    o.m10 = lambda1 * (v1y * v1x) + lambda2 * (v2y * v2x);
    o.m11 = lambda1 * (v1y * v1y) + lambda2 * (v2y * v2y);
    return o;
}

/* GShared.h:446-488 */
static int covarianceToThetaSigmas(m2 cov, float* theta, float* sigma1, float* sigma2) {
    float a = cov.m00;
    float b = 0.5f * (cov.m01 + cov.m10);
    float d = cov.m11;
    if (!gsmo_isfinite(a) || !gsmo_isfinite(b) || !gsmo_isfinite(d)) return 0;
    a = gsmo_fmax(a, 1e-8f);
    d = gsmo_fmax(d, 1e-8f);
    float det = a * d - b * b;
    if (!gsmo_isfinite(det) || !(det > 0.0f)) return 0;
    float mid = 0.5f * (a + d);
    float disc = gsmo_fmax(mid * mid - det, 0.0f);
    float sqrtDisc = sqrtf(disc);
    float lambda1 = gsmo_fmax(mid + sqrtDisc, 1e-8f);
    float lambda2 = gsmo_fmax(mid - sqrtDisc, 1e-8f);
    float v1x, v1y;
    if (fabsf(b) > 1e-8f) {
        float vx = b, vy = lambda1 - a;
        float len = sqrtf(vx * vx + vy * vy);
        v1x = vx / len; /* normalize(): canonical = divide by the length */
        v1y = vy / len;
    } else if (a >= d) {
        v1x = 1.0f; v1y = 0.0f;
    } else {
        v1x = 0.0f; v1y = 1.0f;
    }
    float t = gsmo_atan2(v1y, v1x);
    t = gsmo_fmod_pi(t);
    if (t < 0.0f) t = t + GSMO_PI_F;
    if (t >= GSMO_PI_F) t = t - GSMO_PI_F;
    *theta = t;
    *sigma1 = sqrtf(lambda1);
    *sigma2 = sqrtf(lambda2);
    return gsmo_isfinite(t) && gsmo_isfinite(*sigma1) && gsmo_isfinite(*sigma2);
}

/* GShared.h:275-278: pow(s, 2) is s*s */
static float computeDepthFactor(float depth, float nearPlane, float farPlane) {
    float adjustedFarPlane = farPlane * 0.02f;
    float s = gsmo_clamp((adjustedFarPlane - depth) / (adjustedFarPlane - nearPlane), 0.0f, 1.0f);
    return 1.0f - s * s;
}

/* GShared.h:739-752 */
static int cullByTotalInk(float opacity, float detCov2d, float depth, float nearPlane, float farPlane,
                          float totalInkThreshold) {
    if (totalInkThreshold <= 0.0f) return 0;
    float totalInk = opacity * 6.283185f * sqrtf(gsmo_fmax(detCov2d, 1e-12f));
    float depthFactor = computeDepthFactor(depth, nearPlane, farPlane);
    float adjustedThreshold = depthFactor * totalInkThreshold;
    return totalInk < adjustedThreshold;
}

/* GShared.h:755-768 */
static int cullByTotalInkFromCov(float opacity, m2 cov, float depth, float nearPlane, float farPlane,
                                 float thr) {
    float a = cov.m00;
    float b = 0.5f * (cov.m01 + cov.m10);
    float d = cov.m11;
    float detCov = a * d - b * b;
    return cullByTotalInk(opacity, detCov, depth, nearPlane, farPlane, thr);
}

/* GShared.h:402-427 */
static void computeOBBExtents(m2 cov, float mult, float* ex, float* ey) {
    float a = cov.m00, b = cov.m01, d = cov.m11;
    float det = a * d - b * b;
    float mid = 0.5f * (a + d);
    float disc = gsmo_fmax(mid * mid - det, 1e-6f);
    float sqrtDisc = sqrtf(disc);
    float lambda1 = mid + sqrtDisc;
    float lambda2 = gsmo_fmax(mid - sqrtDisc, 1e-6f);
    float e1 = mult * sqrtf(gsmo_fmax(lambda1, 1e-6f));
    float e2 = mult * sqrtf(gsmo_fmax(lambda2, 1e-6f));
    float v1x, v1y;
    if (fabsf(b) > 1e-6f) {
        float vx = b, vy = lambda1 - a;
        float vlen = sqrtf(vx * vx + vy * vy);
        float dn = gsmo_fmax(vlen, 1e-6f);
        v1x = vx / dn;
        v1y = vy / dn;
    } else if (a >= d) {
        v1x = 1.0f; v1y = 0.0f;
    } else {
        v1x = 0.0f; v1y = 1.0f;
    }
    *ex = fabsf(v1x) * e1 + fabsf(v1y) * e2;
    *ey = fabsf(v1y) * e1 + fabsf(v1x) * e2;
}

/* GShared.h:434-444 */
#define GSMO_THETA_PACK 0x1.45f1c0p+14f   /* binary32(65535.0f / kPiF) */
#define GSMO_THETA_UNPACK 0x1.922148p-15f /* binary32(kPiF / 65535.0f) */
static uint16_t packThetaPi(float theta) {
    theta = gsmo_fmod_pi(theta);
    if (theta < 0.0f) theta = theta + GSMO_PI_F;
    float u = theta * GSMO_THETA_PACK;
    return (uint16_t)gsmo_clamp(u + 0.5f, 0.0f, 65535.0f);
}
static float unpackThetaPi(uint16_t p) { return (float)p * GSMO_THETA_UNPACK; }

/* GShared.h:569-585 */
static void conicFromSigmaTheta(float sigma1, float sigma2, float theta, float* A, float* B, float* C) {
    float s, c;
    gsmo_sincos(theta, &s, &c);
    float invS1sq = 1.0f / gsmo_fmax(sigma1 * sigma1, 1e-12f);
    float invS2sq = 1.0f / gsmo_fmax(sigma2 * sigma2, 1e-12f);
    float cc = c * c, ss = s * s, cs = c * s;
    *A = cc * invS1sq + ss * invS2sq;
    *C = ss * invS1sq + cc * invS2sq;
    *B = cs * (invS1sq - invS2sq);
}

/* GShared.h:490-510 */
static void conicFromThetaSigmas(float theta, float sigma1, float sigma2, float* A, float* B, float* C) {
    float s, c;
    gsmo_sincos(theta, &s, &c);
    float sig1 = gsmo_fmax(sigma1, 1e-4f);
    float sig2 = gsmo_fmax(sigma2, 1e-4f);
    float invVar1 = 1.0f / (sig1 * sig1);
    float invVar2 = 1.0f / (sig2 * sig2);
    float cc = c * c, ss = s * s, cs = c * s;
    *A = cc * invVar1 + ss * invVar2;
    *B = cs * (invVar1 - invVar2);
    *C = ss * invVar1 + cc * invVar2;
}

/* GShared.h:590-593 */
static float computeD2Cutoff(float opacity, float tau) {
    if (opacity < tau) return -1.0f;
    return -2.0f * gsmo_log(tau / opacity);
}

/* GShared.h:518-520 */
static float evalQuad(float x, float y, float a, float b, float c) {
    return (a * x * x + 2.0f * b * x * y) + c * y * y;
}

/* GShared.h:525-564 */
static float minQuadRect(float xmin, float xmax, float ymin, float ymax, float a, float b, float c) {
    if (xmin <= 0.0f && 0.0f <= xmax && ymin <= 0.0f && 0.0f <= ymax) return 0.0f;
    float invA = 1.0f / gsmo_fmax(a, 1e-20f);
    float invC = 1.0f / gsmo_fmax(c, 1e-20f);
    float qmin = INFINITY;
    {
        float x = xmin;
        float y = gsmo_clamp(-(b * invC) * x, ymin, ymax);
        qmin = gsmo_fmin(qmin, evalQuad(x, y, a, b, c));
    }
    {
        float x = xmax;
        float y = gsmo_clamp(-(b * invC) * x, ymin, ymax);
        qmin = gsmo_fmin(qmin, evalQuad(x, y, a, b, c));
    }
    {
        float y = ymin;
        float x = gsmo_clamp(-(b * invA) * y, xmin, xmax);
        qmin = gsmo_fmin(qmin, evalQuad(x, y, a, b, c));
    }
    {
        float y = ymax;
        float x = gsmo_clamp(-(b * invA) * y, xmin, xmax);
        qmin = gsmo_fmin(qmin, evalQuad(x, y, a, b, c));
    }
    return qmin;
}

typedef struct { int minTX, maxTX, minTY, maxTY, valid; } tile_bounds;

/* GShared.h:791-828 */
static tile_bounds computeTileBounds(float sx, float sy, float ex, float ey, float width, float height,
                                     int tileW, int tileH, int tilesX, int tilesY) {
    tile_bounds r;
    float xmin = sx - ex, xmax = sx + ex, ymin = sy - ey, ymax = sy + ey;
    float maxW = width - 1.0f, maxH = height - 1.0f;
    xmin = gsmo_clamp(xmin, 0.0f, maxW);
    xmax = gsmo_clamp(xmax, 0.0f, maxW);
    ymin = gsmo_clamp(ymin, 0.0f, maxH);
    ymax = gsmo_clamp(ymax, 0.0f, maxH);
    r.minTX = (int)floorf(xmin / (float)tileW);
    r.maxTX = (int)ceilf(xmax / (float)tileW) - 1;
    r.minTY = (int)floorf(ymin / (float)tileH);
    r.maxTY = (int)ceilf(ymax / (float)tileH) - 1;
    if (r.minTX < 0) r.minTX = 0;
    if (r.minTY < 0) r.minTY = 0;
    if (r.maxTX > tilesX - 1) r.maxTX = tilesX - 1;
    if (r.maxTY > tilesY - 1) r.maxTY = tilesY - 1;
    r.valid = (r.minTX <= r.maxTX && r.minTY <= r.maxTY);
    return r;
}

/* GShared.h:118-133 */
static float srgbToLinearChannel(float c) {
    c = gsmo_clamp(c, 0.0f, 1.0f);
    return (c <= 0.04045f) ? (c / 12.92f) : gsmo_powr((c + 0.055f) / 1.055f, 2.4f);
}

/* DepthFirstProjectCullEncoder.swift:13-20 */
static int shDegreeFromComponents(uint32_t n) {
    if (n <= 1) return 0;
    if (n <= 4) return 1;
    if (n <= 9) return 2;
    return 3;
}

static inline float loadSH(const void* h, int precision, size_t i) {
    return precision == GSMO_F16 ? gsmo_h2f(((const gsmo_half*)h)[i]) : ((const float*)h)[i];
}

/* GShared.h:38-116 with SH_DEGREE taken from the function constant (always set by the encoder) */
static v3 computeSHColor(const void* harmonics, int precision, uint32_t gid, v3 pos, v3 cam,
                         uint32_t shComponents) {
    static const float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
    static const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
    static const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};
    int degree = shDegreeFromComponents(shComponents);
    v3 color;
    if (degree == 0 || shComponents == 0) {
        size_t base = (size_t)gid * 3u;
        color.x = loadSH(harmonics, precision, base) * SH_C0;
        color.y = loadSH(harmonics, precision, base + 1) * SH_C0;
        color.z = loadSH(harmonics, precision, base + 2) * SH_C0;
        return color;
    }
    v3 dv = {cam.x - pos.x, cam.y - pos.y, cam.z - pos.z};
    float len = sqrtf((dv.x * dv.x + dv.y * dv.y) + dv.z * dv.z);
    v3 dir = {dv.x / len, dv.y / len, dv.z / len}; /* normalize(): canonical = divide by length */
    float xx = dir.x * dir.x, yy = dir.y * dir.y, zz = dir.z * dir.z;
    float xy = dir.x * dir.y, yz = dir.y * dir.z, xz = dir.x * dir.z;
    float sh[16];
    sh[0] = SH_C0;
    sh[1] = -SH_C1 * dir.y;
    sh[2] = SH_C1 * dir.z;
    sh[3] = -SH_C1 * dir.x;
    if (degree >= 2) {
        sh[4] = C2[0] * xy;
        sh[5] = C2[1] * yz;
        sh[6] = C2[2] * ((2.0f * zz - xx) - yy);
        sh[7] = C2[3] * xz;
        sh[8] = C2[4] * (xx - yy);
    }
    if (degree >= 3) {
        sh[9] = C3[0] * dir.y * (3.0f * xx - yy);
        sh[10] = C3[1] * xy * dir.z;
        sh[11] = C3[2] * dir.y * ((4.0f * zz - xx) - yy);
        sh[12] = C3[3] * dir.z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
        sh[13] = C3[4] * dir.x * ((4.0f * zz - xx) - yy);
        sh[14] = C3[5] * dir.z * (xx - yy);
        sh[15] = C3[6] * dir.x * (xx - 3.0f * yy);
    }
    uint32_t coeffs = degree == 1 ? 4u : (degree == 2 ? 9u : 16u);
    size_t base = (size_t)gid * coeffs * 3u;
    color.x = 0.0f; color.y = 0.0f; color.z = 0.0f;
    for (uint32_t i = 0; i < coeffs; ++i) {
        color.x = color.x + loadSH(harmonics, precision, base + i) * sh[i];
        color.y = color.y + loadSH(harmonics, precision, base + coeffs + i) * sh[i];
        color.z = color.z + loadSH(harmonics, precision, base + 2u * coeffs + i) * sh[i];
    }
    return color;
}

/* DFS.metal:33-37 */
static uint32_t float_to_sortable_uint(float v) {
    uint32_t bits = gsmo_f2u(v);
    uint32_t mask = (bits & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u;
    return bits ^ mask;
}
/* DFS.metal:39-43 */
static float sortable_uint_to_float(uint32_t v) {
    uint32_t bits = (v & 0x80000000u) ? (v ^ 0x80000000u) : ~v;
    return gsmo_u2f(bits);
}

static void loadGaussian(const void* g, int precision, uint32_t gid, v3* pos, v3* scale, v4* quat,
                         float* opacity) {
    if (precision == GSMO_F16) {
        const gsmo_packed_f16* p = (const gsmo_packed_f16*)g + gid;
        pos->x = p->px; pos->y = p->py; pos->z = p->pz;
        scale->x = gsmo_h2f(p->sx); scale->y = gsmo_h2f(p->sy); scale->z = gsmo_h2f(p->sz);
        quat->x = gsmo_h2f(p->rx); quat->y = gsmo_h2f(p->ry); quat->z = gsmo_h2f(p->rz); quat->w = gsmo_h2f(p->rw);
        *opacity = gsmo_h2f(p->opacity);
    } else {
        const gsmo_packed_f32* p = (const gsmo_packed_f32*)g + gid;
        pos->x = p->px; pos->y = p->py; pos->z = p->pz;
        scale->x = p->sx; scale->y = p->sy; scale->z = p->sz;
        quat->x = p->rot[0]; quat->y = p->rot[1]; quat->z = p->rot[2]; quat->w = p->rot[3];
        *opacity = p->opacity;
    }
}

static uint8_t quantU8(float v) { return (uint8_t)gsmo_clamp(v * 255.0f, 0.0f, 255.0f); }

static void setCulled(int32_t* bounds, uint32_t* nTouched, uint32_t gid) {
    nTouched[gid] = 0;
    bounds[4 * gid + 0] = 0; bounds[4 * gid + 1] = -1; bounds[4 * gid + 2] = 0; bounds[4 * gid + 3] = -1;
}

/* count tiles whose min quadratic distance is inside the alpha cutoff; optionally emit them.
 * DFS.metal:166-205 (count) and :667-715 (emit) -- both run on the QUANTISED record (quirk Q3). */
static uint32_t walkTiles(const gsmo_render_data* rd, int minTX, int maxTX, int minTY, int maxTY,
                          float alphaThreshold, uint32_t tileW_, uint32_t tileH_, uint32_t tilesX,
                          int emit, int tileId16, void* tileIds, int32_t* instanceIdx,
                          uint32_t writeOffset, uint32_t maxAssignments, int32_t originalIdx) {
    float meanX_q = gsmo_h2f(rd->meanX);
    float meanY_q = gsmo_h2f(rd->meanY);
    float theta_q = unpackThetaPi(rd->theta);
    float sigma1_q = gsmo_h2f(rd->sigma1);
    float sigma2_q = gsmo_h2f(rd->sigma2);
    float opacity_q = (float)rd->opacity * (1.0f / 255.0f);
    float ca, cb, cc;
    conicFromSigmaTheta(sigma1_q, sigma2_q, theta_q, &ca, &cb, &cc);
    float tau = gsmo_fmax(alphaThreshold, 1e-12f);
    float d2Cutoff = computeD2Cutoff(opacity_q, tau);
    uint32_t touched = 0;
    if (d2Cutoff >= 0.0f) {
        float tileW = (float)tileW_, tileH = (float)tileH_;
        for (int ty = minTY; ty <= maxTY; ++ty) {
            float tileMinY = (float)ty * tileH;
            float tileMaxY = tileMinY + tileH;
            float tile_ymin = tileMinY - meanY_q;
            float tile_ymax = tileMaxY - meanY_q;
            for (int tx = minTX; tx <= maxTX; ++tx) {
                float tileMinX = (float)tx * tileW;
                float tileMaxX = tileMinX + tileW;
                float tile_xmin = tileMinX - meanX_q;
                float tile_xmax = tileMaxX - meanX_q;
                float d2min = minQuadRect(tile_xmin, tile_xmax, tile_ymin, tile_ymax, ca, cb, cc);
                if (d2min <= d2Cutoff) {
                    if (!emit) {
                        touched++;
                    } else if (writeOffset < maxAssignments) {
                        uint32_t tileId = (uint32_t)(ty * (int)tilesX + tx);
                        if (tileId16) ((uint16_t*)tileIds)[writeOffset] = (uint16_t)tileId;
                        else ((uint32_t*)tileIds)[writeOffset] = tileId;
                        instanceIdx[writeOffset] = originalIdx;
                        writeOffset++;
                        touched++;
                    }
                }
            }
        }
    }
    return touched;
}

/* ---------------------------------------------------------------- stage 1: project + cull (mono) */
void gsmo_project_cull(const void* gaussians, const void* harmonics, int precision,
                       const gsmo_camera* cam, const gsmo_binning* bin, gsmo_render_data* renderData,
                       int32_t* bounds, uint32_t* preDepthKeys, uint32_t* nTouched,
                       uint32_t* totalInstances) {
    uint64_t total = 0;
    const uint32_t N = cam->gaussianCount;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : total)
    for (uint32_t gid = 0; gid < N; ++gid) {
        v3 position, scale; v4 rot; float opacity;
        loadGaussian(gaussians, precision, gid, &position, &scale, &rot, &opacity);

        /* (1) DFS.metal:63-69, GShared.h:719-722 */
        float maxScale = gsmo_fmax(scale.x, gsmo_fmax(scale.y, scale.z));
        if (maxScale < 0.0005f) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }

        v4 p4 = {position.x, position.y, position.z, 1.0f};
        v4 viewPos4 = mul44(cam->view, p4);
        v4 clip = mul44(cam->proj, viewPos4);
        float depth = clip.w;
        /* (2) DFS.metal:76-81 */
        if (!(clip.w > cam->nearPlane)) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }
        /* (3) DFS.metal:82-87 */
        if (depth > cam->farPlane) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }

        float ndcX = clip.x / clip.w;
        float ndcY = clip.y / clip.w;
        float screenX = (ndcX + 1.0f) * 0.5f * cam->width;  /* GShared.h:150-155 */
        float screenY = (ndcY + 1.0f) * 0.5f * cam->height;

        /* (4) DFS.metal:93-99 */
        if (opacity < bin->alphaThreshold) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }

        v4 quat = normalizeQuaternion(rot);
        m3 cov3d = buildCovariance3D(scale, quat);
        v3 viewPos = {viewPos4.x, viewPos4.y, viewPos4.z};
        m2 cov2d = projectCovariance2D(cov3d, viewPos, cam->view, cam->proj, cam->width, cam->height);
        cov2d = stabilizeCovariance2D(cov2d, cam->width, cam->height);

        float theta, sigma1, sigma2;
        /* (5) DFS.metal:110-115 -- key left stale */
        if (!covarianceToThetaSigmas(cov2d, &theta, &sigma1, &sigma2)) { setCulled(bounds, nTouched, gid); continue; }
        float radius = 3.0f * gsmo_fmax(sigma1, sigma2);
        /* (6) DFS.metal:116-122 -- key left stale */
        if (radius < 0.5f) { setCulled(bounds, nTouched, gid); continue; }
        /* (7) DFS.metal:124-129 */
        if (cullByTotalInkFromCov(opacity, cov2d, depth, cam->nearPlane, cam->farPlane, bin->totalInkThreshold)) {
            setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue;
        }
        float obbX, obbY;
        computeOBBExtents(cov2d, 3.0f, &obbX, &obbY);
        /* (8) DFS.metal:131-137, GShared.h:771-781 -- key left stale */
        if (screenX + obbX < 0.0f || screenX - obbX > cam->width || screenY + obbY < 0.0f ||
            screenY - obbY > cam->height) { setCulled(bounds, nTouched, gid); continue; }

        /* DFS.metal:139-141 */
        v3 color = computeSHColor(harmonics, precision, gid, position, *(const v3*)cam->center, cam->shComponents);
        color.x = gsmo_fmax(color.x + 0.5f, 0.0f);
        color.y = gsmo_fmax(color.y + 0.5f, 0.0f);
        color.z = gsmo_fmax(color.z + 0.5f, 0.0f);
        if (cam->inputIsSRGB > 0.5f) {
            color.x = srgbToLinearChannel(color.x);
            color.y = srgbToLinearChannel(color.y);
            color.z = srgbToLinearChannel(color.z);
        }

        /* DFS.metal:143-154 */
        gsmo_render_data rd;
        rd.meanX = gsmo_f2h(screenX);
        rd.meanY = gsmo_f2h(screenY);
        rd.theta = packThetaPi(theta);
        rd.sigma1 = gsmo_f2h(sigma1);
        rd.sigma2 = gsmo_f2h(sigma2);
        rd.depth = gsmo_f2h(depth);
        rd.colorR = quantU8(color.x);
        rd.colorG = quantU8(color.y);
        rd.colorB = quantU8(color.z);
        rd.opacity = quantU8(opacity);
        renderData[gid] = rd;

        /* DFS.metal:157-164 */
        tile_bounds tb = computeTileBounds(screenX, screenY, obbX, obbY, cam->width, cam->height,
                                           (int)bin->tileWidth, (int)bin->tileHeight, (int)bin->tilesX,
                                           (int)bin->tilesY);
        bounds[4 * gid + 0] = tb.minTX; bounds[4 * gid + 1] = tb.maxTX;
        bounds[4 * gid + 2] = tb.minTY; bounds[4 * gid + 3] = tb.maxTY;

        /* DFS.metal:166-205 */
        uint32_t touched = walkTiles(&rd, tb.minTX, tb.maxTX, tb.minTY, tb.maxTY, bin->alphaThreshold,
                                     bin->tileWidth, bin->tileHeight, bin->tilesX, 0, 0, NULL, NULL, 0, 0, 0);
        /* (9) DFS.metal:207-212 */
        if (touched == 0) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }
        preDepthKeys[gid] = float_to_sortable_uint(depth);
        nTouched[gid] = touched;
        total += touched;
    }
    *totalInstances = (uint32_t)total; /* the device counter is 32-bit (DFS.metal:54) */
}

/* ---------------------------------------------------------------- stage 1s: stereo project */
typedef struct {
    float screenX, screenY, theta, sigma1, sigma2, detCov, obbX, obbY, depth;
    int tb[4];
    uint32_t touchedTiles;
    int visible;
} eye_result;

/* DFS.metal:249-339 */
static eye_result projectToEye(v3 scenePos, v3 scale, v4 quat, const float* sceneTransform,
                               const float* view, const float* proj, float width, float height,
                               float nearPlane, float farPlane, const gsmo_binning* bin) {
    eye_result r;
    memset(&r, 0, sizeof r);
    r.tb[0] = 0; r.tb[1] = -1; r.tb[2] = 0; r.tb[3] = -1;
    v4 sp = {scenePos.x, scenePos.y, scenePos.z, 1.0f};
    v4 worldPos4 = mul44(sceneTransform, sp);
    v4 viewPos4 = mul44(view, worldPos4);
    v4 clip = mul44(proj, viewPos4);
    float depth = clip.w;
    r.depth = depth;
    if (!(clip.w > nearPlane)) return r;
    if (depth > farPlane) return r;
    float ndcX = clip.x / clip.w, ndcY = clip.y / clip.w;
    r.screenX = (ndcX + 1.0f) * 0.5f * width;
    r.screenY = (ndcY + 1.0f) * 0.5f * height;
    float s0 = sceneTransform[0], s1 = sceneTransform[1], s2 = sceneTransform[2];
    float sceneScale = sqrtf((s0 * s0 + s1 * s1) + s2 * s2); /* length(), DFS.metal:293 */
    v3 sc = {scale.x * sceneScale, scale.y * sceneScale, scale.z * sceneScale};
    m3 cov3d = buildCovariance3D(sc, quat);
    v3 viewPos = {viewPos4.x, viewPos4.y, viewPos4.z};
    m2 cov2d = projectCovariance2D(cov3d, viewPos, view, proj, width, height);
    cov2d = stabilizeCovariance2D(cov2d, width, height);
    float theta, sigma1, sigma2;
    if (!covarianceToThetaSigmas(cov2d, &theta, &sigma1, &sigma2)) return r;
    r.theta = theta; r.sigma1 = sigma1; r.sigma2 = sigma2;
    float a = cov2d.m00, b = 0.5f * (cov2d.m01 + cov2d.m10), d = cov2d.m11;
    r.detCov = gsmo_fmax(a * d - b * b, 0.0f);
    float radius = 3.0f * gsmo_fmax(sigma1, sigma2);
    if (radius < 0.5f) return r;
    computeOBBExtents(cov2d, 3.0f, &r.obbX, &r.obbY);
    if (r.screenX + r.obbX < 0.0f || r.screenX - r.obbX > width || r.screenY + r.obbY < 0.0f ||
        r.screenY - r.obbY > height) return r;
    tile_bounds tb = computeTileBounds(r.screenX, r.screenY, r.obbX, r.obbY, width, height,
                                       (int)bin->tileWidth, (int)bin->tileHeight, (int)bin->tilesX,
                                       (int)bin->tilesY);
    r.tb[0] = tb.minTX; r.tb[1] = tb.maxTX; r.tb[2] = tb.minTY; r.tb[3] = tb.maxTY;
    uint32_t cx = tb.valid ? (uint32_t)(tb.maxTX - tb.minTX + 1) : 0;
    uint32_t cy = tb.valid ? (uint32_t)(tb.maxTY - tb.minTY + 1) : 0;
    r.touchedTiles = cx * cy;
    r.visible = 1;
    return r;
}

/* DFS.metal:341-499 */
void gsmo_project_cull_stereo(const void* gaussians, const void* harmonics, int precision,
                              const gsmo_stereo_camera* cam, const gsmo_binning* bin,
                              gsmo_stereo_render_data* renderData, int32_t* bounds,
                              uint32_t* preDepthKeys, uint32_t* nTouched, uint32_t* totalInstances) {
    uint64_t total = 0;
    const uint32_t N = cam->gaussianCount;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : total)
    for (uint32_t gid = 0; gid < N; ++gid) {
        v3 position, scale; v4 rot; float opacity;
        loadGaussian(gaussians, precision, gid, &position, &scale, &rot, &opacity);
        float maxScale = gsmo_fmax(scale.x, gsmo_fmax(scale.y, scale.z));
        if (maxScale < 0.0005f) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }
        if (opacity < bin->alphaThreshold) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }
        v4 quat = normalizeQuaternion(rot);
        eye_result L = projectToEye(position, scale, quat, cam->sceneTransform, cam->leftView, cam->leftProj,
                                    cam->width, cam->height, cam->nearPlane, cam->farPlane, bin);
        eye_result R = projectToEye(position, scale, quat, cam->sceneTransform, cam->rightView, cam->rightProj,
                                    cam->width, cam->height, cam->nearPlane, cam->farPlane, bin);
        if (!L.visible && !R.visible) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }
        float checkDepth = L.visible ? L.depth : R.depth;
        if (L.visible && R.visible) checkDepth = (L.depth + R.depth) * 0.5f;
        float detCov = L.visible ? L.detCov : R.detCov;
        if (L.visible && R.visible) detCov = gsmo_fmax(L.detCov, R.detCov);
        if (cullByTotalInk(opacity, detCov, checkDepth, cam->nearPlane, cam->farPlane, bin->totalInkThreshold)) {
            setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue;
        }
        v3 mid = {(cam->leftCenter[0] + cam->rightCenter[0]) * 0.5f,
                  (cam->leftCenter[1] + cam->rightCenter[1]) * 0.5f,
                  (cam->leftCenter[2] + cam->rightCenter[2]) * 0.5f};
        v3 color = computeSHColor(harmonics, precision, gid, position, mid, cam->shComponents);
        color.x = gsmo_fmax(color.x + 0.5f, 0.0f);
        color.y = gsmo_fmax(color.y + 0.5f, 0.0f);
        color.z = gsmo_fmax(color.z + 0.5f, 0.0f);
        if (cam->inputIsSRGB > 0.5f) {
            color.x = srgbToLinearChannel(color.x);
            color.y = srgbToLinearChannel(color.y);
            color.z = srgbToLinearChannel(color.z);
        }
        int ub[4];
        if (L.visible && R.visible) {
            ub[0] = L.tb[0] < R.tb[0] ? L.tb[0] : R.tb[0];
            ub[1] = L.tb[1] > R.tb[1] ? L.tb[1] : R.tb[1];
            ub[2] = L.tb[2] < R.tb[2] ? L.tb[2] : R.tb[2];
            ub[3] = L.tb[3] > R.tb[3] ? L.tb[3] : R.tb[3];
        } else if (L.visible) {
            memcpy(ub, L.tb, sizeof ub);
        } else {
            memcpy(ub, R.tb, sizeof ub);
        }
        int utx = ub[1] - ub[0] + 1; if (utx < 0) utx = 0;
        int uty = ub[3] - ub[2] + 1; if (uty < 0) uty = 0;
        uint32_t touched = (uint32_t)(utx * uty);
        if (touched == 0) { setCulled(bounds, nTouched, gid); preDepthKeys[gid] = 0xFFFFFFFFu; continue; }

        gsmo_stereo_render_data rd;
        memset(&rd, 0, sizeof rd);
        const gsmo_half negHuge = gsmo_f2h(-1e10f); /* = -inf, DFS.metal:461 */
        if (L.visible) {
            float A, B, C;
            rd.leftMeanX = gsmo_f2h(L.screenX); rd.leftMeanY = gsmo_f2h(L.screenY);
            conicFromThetaSigmas(L.theta, L.sigma1, L.sigma2, &A, &B, &C);
            rd.leftCxx = gsmo_f2h(A); rd.leftCyy = gsmo_f2h(C); rd.leftCxy2 = gsmo_f2h(2.0f * B);
            rd.leftDepth = gsmo_f2h(L.depth);
        } else {
            rd.leftMeanX = negHuge; rd.leftMeanY = negHuge;
        }
        if (R.visible) {
            float A, B, C;
            rd.rightMeanX = gsmo_f2h(R.screenX); rd.rightMeanY = gsmo_f2h(R.screenY);
            conicFromThetaSigmas(R.theta, R.sigma1, R.sigma2, &A, &B, &C);
            rd.rightCxx = gsmo_f2h(A); rd.rightCyy = gsmo_f2h(C); rd.rightCxy2 = gsmo_f2h(2.0f * B);
            rd.rightDepth = gsmo_f2h(R.depth);
        } else {
            rd.rightMeanX = negHuge; rd.rightMeanY = negHuge;
        }
        rd.colorR = quantU8(color.x); rd.colorG = quantU8(color.y); rd.colorB = quantU8(color.z);
        rd.opacity = quantU8(opacity);
        rd.centerDepth = gsmo_f2h(checkDepth);
        rd._pad0 = 0;
        renderData[gid] = rd;
        bounds[4 * gid + 0] = ub[0]; bounds[4 * gid + 1] = ub[1];
        bounds[4 * gid + 2] = ub[2]; bounds[4 * gid + 3] = ub[3];
        nTouched[gid] = touched;
        preDepthKeys[gid] = float_to_sortable_uint(checkDepth);
        total += touched;
    }
    *totalInstances = (uint32_t)total;
}

/* ---------------------------------------------------------------- stage 1.25: compaction */
void gsmo_compact_visible(const uint32_t* nTouched, const uint32_t* preDepthKeys, uint32_t count,
                          uint32_t maxOut, int depthKey16, uint32_t* depthKeys,
                          int32_t* primitiveIndices, uint32_t* visibleCount) {
    /* exclusive scan of (nTouched>0) in gid order (DFS.metal:518-587), then scatter (:589-621) */
    uint32_t out = 0;
    for (uint32_t gid = 0; gid < count; ++gid) {
        if (nTouched[gid] > 0u) {
            if (out < maxOut) {
                uint32_t key = preDepthKeys[gid];
                if (depthKey16) {
                    float depth = sortable_uint_to_float(key);
                    key = (uint32_t)(uint16_t)(gsmo_f2h(depth) ^ 0x8000u);
                }
                depthKeys[out] = key;
                primitiveIndices[out] = (int32_t)gid;
            }
            out++;
        }
    }
    *visibleCount = out;
}

/* DFS.metal:2184-2203 */
void gsmo_prepare_header(uint32_t visibleCount, uint32_t totalInstances, uint32_t maxGaussians,
                         uint32_t maxInstances, gsmo_df_header* h) {
    memset(h, 0, sizeof *h); /* resetDepthFirstStateKernel, DFS.metal:1372-1385 */
    if (visibleCount > maxGaussians) { visibleCount = maxGaussians; h->overflow = 1u; }
    if (totalInstances > maxInstances) { totalInstances = maxInstances; h->overflow = 1u; }
    const uint32_t al = 256u * 4u;
    h->visibleCount = visibleCount;
    h->totalInstances = totalInstances;
    h->paddedVisibleCount = ((visibleCount + al - 1u) / al) * al;
    h->paddedInstanceCount = ((totalInstances + al - 1u) / al) * al;
}

/* ---------------------------------------------------------------- stable LSD radix sort */
#define SORT_IMPL(NAME, KEYT)                                                                         \
    void NAME(KEYT* keys, int32_t* payload, uint32_t count, int numPasses) {                          \
        if (count == 0 || numPasses <= 0) return;                                                     \
        KEYT* k2 = (KEYT*)malloc((size_t)count * sizeof(KEYT));                                       \
        int32_t* p2 = (int32_t*)malloc((size_t)count * sizeof(int32_t));                              \
        KEYT* src = keys; int32_t* srcp = payload; KEYT* dst = k2; int32_t* dstp = p2;                \
        int nt = omp_get_max_threads();                                                               \
        if ((uint32_t)nt > count / 4096u + 1u) nt = (int)(count / 4096u + 1u);                        \
        uint32_t* hist = (uint32_t*)malloc((size_t)nt * 256u * sizeof(uint32_t));                     \
        for (int pass = 0; pass < numPasses; ++pass) {                                                \
            const int shift = 8 * pass; /* value_to_key_at_digit, RadixSortHelpers.h:85-88 */         \
            _Pragma("omp parallel num_threads(nt)")                                                   \
            {                                                                                         \
                int t = omp_get_thread_num();                                                         \
                const int nth = omp_get_num_threads(); /* may be < nt */                             \
                uint32_t lo = (uint32_t)(((uint64_t)count * (uint64_t)t) / (uint64_t)nth);            \
                uint32_t hi = (uint32_t)(((uint64_t)count * (uint64_t)(t + 1)) / (uint64_t)nth);      \
                uint32_t* h = hist + (size_t)t * 256u;                                                \
                memset(h, 0, 256u * sizeof(uint32_t));                                                \
                for (uint32_t i = lo; i < hi; ++i) h[((uint32_t)src[i] >> shift) & 0xFFu]++;          \
                _Pragma("omp barrier")                                                                \
                _Pragma("omp single")                                                                 \
                {                                                                                     \
                    /* bin-major, chunk-minor exclusive scan: lower chunks first inside a bin */      \
                    uint32_t run = 0;                                                                 \
                    for (int b = 0; b < 256; ++b)                                                     \
                        for (int tt = 0; tt < nth; ++tt) {                                            \
                            uint32_t c = hist[(size_t)tt * 256u + b];                                 \
                            hist[(size_t)tt * 256u + b] = run;                                        \
                            run += c;                                                                 \
                        }                                                                             \
                }                                                                                     \
                for (uint32_t i = lo; i < hi; ++i) {                                                  \
                    uint32_t d = h[((uint32_t)src[i] >> shift) & 0xFFu]++;                            \
                    dst[d] = src[i];                                                                  \
                    dstp[d] = srcp[i];                                                                \
                }                                                                                     \
            }                                                                                         \
            KEYT* tk = src; src = dst; dst = tk;                                                      \
            int32_t* tp = srcp; srcp = dstp; dstp = tp;                                               \
        }                                                                                             \
        if (src != keys) { /* odd pass count: copy back (TileSortEncoder.swift:170-177) */            \
            memcpy(keys, src, (size_t)count * sizeof(KEYT));                                          \
            memcpy(payload, srcp, (size_t)count * sizeof(int32_t));                                   \
        }                                                                                             \
        free(hist); free(k2); free(p2);                                                               \
    }

SORT_IMPL(gsmo_sort_pairs_u32, uint32_t)
SORT_IMPL(gsmo_sort_pairs_u16, uint16_t)

/* ---------------------------------------------------------------- stages 3, 4 */
void gsmo_apply_depth_order(const int32_t* sortedIdx, const uint32_t* nTouched, uint32_t visibleCount,
                            uint32_t* ordered) {
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < visibleCount; ++i) {
        int32_t o = sortedIdx[i];
        ordered[i] = (o < 0) ? 0u : nTouched[o];
    }
}

void gsmo_exclusive_scan(const uint32_t* in, uint32_t count, uint32_t* out) {
    uint32_t run = 0;
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t v = in[i];
        out[i] = run;
        run += v;
    }
}

/* ---------------------------------------------------------------- stage 5 */
void gsmo_create_instances(const int32_t* sortedIdx, const uint32_t* instanceOffsets,
                           const int32_t* bounds, const gsmo_render_data* renderData,
                           uint32_t visibleCount, uint32_t tilesX, float alphaThreshold,
                           uint32_t maxAssignments, int tileId16, void* tileIds, int32_t* instanceIdx) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (uint32_t i = 0; i < visibleCount; ++i) {
        int32_t o = sortedIdx[i];
        if (o < 0) continue;
        const int32_t* b = bounds + 4 * (size_t)o;
        if (b[0] > b[1] || b[2] > b[3]) continue;
        walkTiles(&renderData[o], b[0], b[1], b[2], b[3], alphaThreshold, 16u, 16u, tilesX, 1, tileId16,
                  tileIds, instanceIdx, instanceOffsets[i], maxAssignments, o);
    }
}

void gsmo_create_instances_stereo(const int32_t* sortedIdx, const uint32_t* instanceOffsets,
                                  const int32_t* bounds, uint32_t visibleCount, uint32_t tilesX,
                                  uint32_t maxAssignments, int tileId16, void* tileIds,
                                  int32_t* instanceIdx) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (uint32_t i = 0; i < visibleCount; ++i) {
        int32_t o = sortedIdx[i];
        if (o < 0) continue;
        const int32_t* b = bounds + 4 * (size_t)o;
        if (b[0] > b[1] || b[2] > b[3]) continue;
        uint32_t w = instanceOffsets[i];
        for (int ty = b[2]; ty <= b[3]; ++ty)
            for (int tx = b[0]; tx <= b[1]; ++tx)
                if (w < maxAssignments) {
                    uint32_t tileId = (uint32_t)(ty * (int)tilesX + tx);
                    if (tileId16) ((uint16_t*)tileIds)[w] = (uint16_t)tileId;
                    else ((uint32_t*)tileIds)[w] = tileId;
                    instanceIdx[w] = o;
                    w++;
                }
    }
}

int gsmo_tile_sort_passes(uint32_t tileCount) {
    /* bitsNeeded = floor(log2(max(tileCount-1,1))) + 1 ; passes = ceil(bits/8) */
    uint32_t v = tileCount > 0 ? (tileCount - 1 > 1 ? tileCount - 1 : 1) : 1;
    int bits = 0;
    while (v) { bits++; v >>= 1; }
    if (tileCount == 0) bits = 1;
    return (bits + 7) / 8;
}

/* ---------------------------------------------------------------- stage 7 */
void gsmo_extract_ranges(const void* sortedTileIds, int tileId16, uint32_t totalInstances,
                         uint32_t tileCount, gsmo_tile_header* headers, uint32_t* activeTiles,
                         uint32_t* activeTileCount) {
    uint32_t nActive = 0;
    for (uint32_t tile = 0; tile < tileCount; ++tile) {
        if (totalInstances == 0) { headers[tile].offset = 0; headers[tile].count = 0; continue; }
        uint32_t left = 0, right = totalInstances;
        while (left < right) {
            uint32_t mid = (left + right) >> 1;
            uint32_t mt = tileId16 ? (uint32_t)((const uint16_t*)sortedTileIds)[mid]
                                   : ((const uint32_t*)sortedTileIds)[mid];
            if (mt < tile) left = mid + 1; else right = mid;
        }
        uint32_t start = left;
        right = totalInstances;
        while (left < right) {
            uint32_t mid = (left + right) >> 1;
            uint32_t mt = tileId16 ? (uint32_t)((const uint16_t*)sortedTileIds)[mid]
                                   : ((const uint32_t*)sortedTileIds)[mid];
            if (mt <= tile) left = mid + 1; else right = mid;
        }
        uint32_t end = left;
        headers[tile].offset = start;
        headers[tile].count = end > start ? end - start : 0;
        if (headers[tile].count > 0) activeTiles[nActive++] = tile;
    }
    *activeTileCount = nActive;
}

/* ---------------------------------------------------------------- stage 8 */
void gsmo_clear(gsmo_half* color, gsmo_half* depth, uint32_t width, uint32_t height) {
    const size_t P = (size_t)width * height;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < P; ++i) {
        color[4 * i + 0] = 0; color[4 * i + 1] = 0; color[4 * i + 2] = 0; color[4 * i + 3] = 0x3C00u;
        if (depth) depth[i] = 0;
    }
}

#define H_ONE ((gsmo_half)0x3C00u)
#define H_ZERO ((gsmo_half)0x0000u)
#define H_NEG_HALF ((gsmo_half)0xB800u) /* -0.5h */

static inline gsmo_half h_from_uint(uint32_t v) { return gsmo_f2h((float)v); }
static inline int h_lt(gsmo_half a, gsmo_half b) { return gsmo_h2f(a) < gsmo_h2f(b); }
static inline int h_gt(gsmo_half a, gsmo_half b) { return gsmo_h2f(a) > gsmo_h2f(b); }
static inline int h_ge(gsmo_half a, gsmo_half b) { return gsmo_h2f(a) >= gsmo_h2f(b); }
static inline int h_eq0(gsmo_half a) { return gsmo_h2f(a) == 0.0f; }

/* d.x*d.x*cxx + d.y*d.y*cyy + d.x*d.y*cxy2 (DFS.metal:1770) with the contraction a fast-math compiler applies
 * to (m0 + m1) + m2: fma(dx*dy, cxy2, fma(dy*dy, cyy, (dx*dx)*cxx)); every other product is a rounded half mul */
/* Test-only switch (tests/test_blend_conventions.py): 1 = the canonical contracted form (default), 0 = every product
 * and sum rounded separately -- what SURVEY.md H2 first prescribed. It exists so the distance between the two
 * conventions is measured instead of asserted; the device implements the contracted form only. */
static int g_blendContract = 1;
void gsmo_set_blend_contraction(int on) { g_blendContract = on ? 1 : 0; }
int gsmo_get_blend_contraction(void) { return g_blendContract; }
static inline gsmo_half h_mad(gsmo_half a, gsmo_half b, gsmo_half c) {
    return g_blendContract ? gsmo_hfma(a, b, c) : gsmo_hadd(gsmo_hmul(a, b), c);
}

static inline gsmo_half h_power(gsmo_half dx, gsmo_half dy, gsmo_half cxx, gsmo_half cyy, gsmo_half cxy2) {
    gsmo_half t0 = gsmo_hmul(gsmo_hmul(dx, dx), cxx);
    gsmo_half in = h_mad(gsmo_hmul(dy, dy), cyy, t0);
    return h_mad(gsmo_hmul(dx, dy), cxy2, in);
}

/* DFS.metal:1703-1811 */
void gsmo_blend(const gsmo_tile_header* headers, const gsmo_render_data* gaussians,
                const int32_t* sortedGaussianIndices, const uint32_t* activeTiles,
                uint32_t activeTileCount, uint32_t width, uint32_t height, uint32_t tilesX,
                gsmo_half* colorOut, gsmo_half* depthOut) {
    const gsmo_half h255 = gsmo_f2h(255.0f);
    const gsmo_half thr = gsmo_hdiv(H_ONE, h255);  /* half(1.0h/255.0h) */
    const gsmo_half h099 = gsmo_f2h(0.99f);
#pragma omp parallel for schedule(dynamic, 4)
    for (uint32_t tileIdx = 0; tileIdx < activeTileCount; ++tileIdx) {
        uint32_t tile = activeTiles[tileIdx];
        gsmo_tile_header hdr = headers[tile];
        uint32_t tileX = tile % tilesX, tileY = tile / tilesX;
        for (uint32_t ly = 0; ly < 8; ++ly)
            for (uint32_t lx = 0; lx < 8; ++lx) {
                uint32_t baseX = tileX * 16 + lx * 2, baseY = tileY * 16 + ly * 2;
                /* pixel k: 0=(0,0) 1=(1,0) 2=(0,1) 3=(1,1) */
                gsmo_half px[4], py[4], trans[4], col[4][3], dep[4];
                for (int k = 0; k < 4; ++k) {
                    px[k] = h_from_uint(baseX + (uint32_t)(k & 1));
                    py[k] = h_from_uint(baseY + (uint32_t)(k >> 1));
                    trans[k] = H_ONE;
                    col[k][0] = col[k][1] = col[k][2] = H_ZERO;
                    dep[k] = H_ZERO;
                }
                for (uint32_t i = 0; i < hdr.count; ++i) {
                    gsmo_half maxTrans = gsmo_hmax(gsmo_hmax(trans[0], trans[1]), gsmo_hmax(trans[2], trans[3]));
                    if (h_lt(maxTrans, thr)) break;
                    int32_t gi = sortedGaussianIndices[hdr.offset + i];
                    if (gi < 0) continue;
                    gsmo_render_data g = gaussians[gi];
                    float theta = unpackThetaPi(g.theta);
                    float A, B, C;
                    conicFromThetaSigmas(theta, gsmo_h2f(g.sigma1), gsmo_h2f(g.sigma2), &A, &B, &C);
                    gsmo_half cxx = gsmo_f2h(A), cyy = gsmo_f2h(C), cxy2 = gsmo_f2h(2.0f * B);
                    gsmo_half opacity = gsmo_hdiv(gsmo_f2h((float)g.opacity), h255);
                    gsmo_half gc[3] = {gsmo_hdiv(gsmo_f2h((float)g.colorR), h255),
                                       gsmo_hdiv(gsmo_f2h((float)g.colorG), h255),
                                       gsmo_hdiv(gsmo_f2h((float)g.colorB), h255)};
                    gsmo_half a[4];
                    int allZero = 1;
                    for (int k = 0; k < 4; ++k) {
                        gsmo_half dx = gsmo_hsub(px[k], g.meanX), dy = gsmo_hsub(py[k], g.meanY);
                        gsmo_half p = h_power(dx, dy, cxx, cyy, cxy2);
                        a[k] = gsmo_hmin(gsmo_hmul(opacity, gsmo_hexp(gsmo_hmul(H_NEG_HALF, p))), h099);
                        if (!h_eq0(a[k])) allZero = 0;
                    }
                    if (allZero) continue;
                    for (int k = 0; k < 4; ++k) {
                        gsmo_half w = gsmo_hmul(a[k], trans[k]);
                        for (int c = 0; c < 3; ++c) col[k][c] = h_mad(gc[c], w, col[k][c]);  /* color += gColor * (a*T), contracted */
                        dep[k] = h_mad(g.depth, w, dep[k]);
                    }
                    for (int k = 0; k < 4; ++k) trans[k] = gsmo_hmul(trans[k], gsmo_hsub(H_ONE, a[k]));
                }
                for (int k = 0; k < 4; ++k) {
                    uint32_t x = baseX + (uint32_t)(k & 1), y = baseY + (uint32_t)(k >> 1);
                    if (x < width && y < height) {
                        size_t o = (size_t)y * width + x;
                        colorOut[4 * o + 0] = col[k][0];
                        colorOut[4 * o + 1] = col[k][1];
                        colorOut[4 * o + 2] = col[k][2];
                        colorOut[4 * o + 3] = gsmo_hsub(H_ONE, trans[k]);
                        if (depthOut) depthOut[o] = dep[k];
                    }
                }
            }
    }
}

/* DFS.metal:1825-1982 (clear :1813-1823 is done by the caller through gsmo_clear per slice) */
void gsmo_blend_stereo(const gsmo_tile_header* headers, const gsmo_stereo_render_data* gaussians,
                       const int32_t* sortedGaussianIndices, const uint32_t* activeTiles,
                       uint32_t activeTileCount, uint32_t width, uint32_t height, uint32_t tilesX,
                       gsmo_half* color2) {
    const gsmo_half h255 = gsmo_f2h(255.0f);
    const gsmo_half thr = gsmo_hdiv(H_ONE, h255);
    const gsmo_half h099 = gsmo_f2h(0.99f);
    const gsmo_half r2Max = gsmo_f2h(9.0f);
    const gsmo_half hm60000 = gsmo_f2h(-60000.0f);
    const size_t sliceStride = (size_t)width * height * 4;
#pragma omp parallel for schedule(dynamic, 4)
    for (uint32_t tileIdx = 0; tileIdx < activeTileCount; ++tileIdx) {
        uint32_t tile = activeTiles[tileIdx];
        gsmo_tile_header hdr = headers[tile];
        uint32_t tileX = tile % tilesX, tileY = tile / tilesX;
        for (uint32_t ly = 0; ly < 8; ++ly)
            for (uint32_t lx = 0; lx < 8; ++lx) {
                uint32_t baseX = tileX * 16 + lx * 2, baseY = tileY * 16 + ly * 2;
                gsmo_half px[4], py[4], trans[2][4], col[2][4][3];
                for (int k = 0; k < 4; ++k) {
                    px[k] = h_from_uint(baseX + (uint32_t)(k & 1));
                    py[k] = h_from_uint(baseY + (uint32_t)(k >> 1));
                    for (int e = 0; e < 2; ++e) {
                        trans[e][k] = H_ONE;
                        col[e][k][0] = col[e][k][1] = col[e][k][2] = H_ZERO;
                    }
                }
                for (uint32_t i = 0; i < hdr.count; ++i) {
                    gsmo_half mt[2];
                    for (int e = 0; e < 2; ++e)
                        mt[e] = gsmo_hmax(gsmo_hmax(trans[e][0], trans[e][1]), gsmo_hmax(trans[e][2], trans[e][3]));
                    if (h_lt(gsmo_hmax(mt[0], mt[1]), thr)) break;
                    int32_t gi = sortedGaussianIndices[hdr.offset + i];
                    if (gi < 0) continue;
                    gsmo_stereo_render_data g = gaussians[gi];
                    gsmo_half opacity = gsmo_hdiv(gsmo_f2h((float)g.opacity), h255);
                    gsmo_half gc[3] = {gsmo_hdiv(gsmo_f2h((float)g.colorR), h255),
                                       gsmo_hdiv(gsmo_f2h((float)g.colorG), h255),
                                       gsmo_hdiv(gsmo_f2h((float)g.colorB), h255)};
                    for (int e = 0; e < 2; ++e) {
                        if (!h_ge(mt[e], thr)) continue;
                        gsmo_half mx = e ? g.rightMeanX : g.leftMeanX, my = e ? g.rightMeanY : g.leftMeanY;
                        if (!h_ge(mx, hm60000)) continue;
                        gsmo_half cxx = e ? g.rightCxx : g.leftCxx, cyy = e ? g.rightCyy : g.leftCyy;
                        gsmo_half cxy2 = e ? g.rightCxy2 : g.leftCxy2;
                        gsmo_half p[4], a[4];
                        int allOut = 1;
                        for (int k = 0; k < 4; ++k) {
                            p[k] = h_power(gsmo_hsub(px[k], mx), gsmo_hsub(py[k], my), cxx, cyy, cxy2);
                            if (!h_gt(p[k], r2Max)) allOut = 0;
                            a[k] = H_ZERO;
                        }
                        if (!allOut)
                            for (int k = 0; k < 4; ++k)
                                a[k] = h_gt(p[k], r2Max) ? H_ZERO
                                       : gsmo_hmin(gsmo_hmul(opacity, gsmo_hexp(gsmo_hmul(H_NEG_HALF, p[k]))), h099);
                        int allZero = 1;
                        for (int k = 0; k < 4; ++k) if (!h_eq0(a[k])) allZero = 0;
                        if (allZero) continue;
                        for (int k = 0; k < 4; ++k) {
                            gsmo_half w = gsmo_hmul(a[k], trans[e][k]);
                            for (int c = 0; c < 3; ++c) col[e][k][c] = h_mad(gc[c], w, col[e][k][c]);
                        }
                        for (int k = 0; k < 4; ++k) trans[e][k] = gsmo_hmul(trans[e][k], gsmo_hsub(H_ONE, a[k]));
                    }
                }
                for (int k = 0; k < 4; ++k) {
                    uint32_t x = baseX + (uint32_t)(k & 1), y = baseY + (uint32_t)(k >> 1);
                    if (x < width && y < height) {
                        size_t o = ((size_t)y * width + x) * 4;
                        for (int e = 0; e < 2; ++e) {
                            gsmo_half* dst = color2 + (size_t)e * sliceStride + o;
                            dst[0] = col[e][k][0]; dst[1] = col[e][k][1]; dst[2] = col[e][k][2];
                            dst[3] = gsmo_hsub(H_ONE, trans[e][k]);
                        }
                    }
                }
            }
    }
}

void gsmo_stereo_copy(const gsmo_half* color2, uint32_t width, uint32_t height, int flipY,
                      gsmo_half* dst) {
    const size_t sliceStride = (size_t)width * height * 4;
#pragma omp parallel for schedule(static)
    for (uint32_t y = 0; y < height; ++y) {
        uint32_t sy = flipY ? (height - 1 - y) : y;
        for (int e = 0; e < 2; ++e)
            memcpy(dst + ((size_t)y * 2 * width + (size_t)e * width) * 4,
                   color2 + (size_t)e * sliceStride + (size_t)sy * width * 4, (size_t)width * 8);
    }
}

/* ---------------------------------------------------------------- whole frames */
static void binningFor(uint32_t width, uint32_t height, gsmo_binning* b) {
    /* GlobalRenderer.swift:54-69 with DFR.swift:8-9 */
    b->tileWidth = 16; b->tileHeight = 16;
    b->tilesX = (width + 15u) / 16u; b->tilesY = (height + 15u) / 16u;
    b->alphaThreshold = 0.005f; b->totalInkThreshold = 2.0f;
}

static void frame_common(gsmo_frame* f, const gsmo_binning* bin, int stereo, uint32_t N, double* t) {
    /* DFR.swift:291-430 / :649-787 */
    t[1] = now_s();
    gsmo_compact_visible(f->nTouched, f->preDepthKeys, N, f->maxGaussians, f->depthKey16, f->depthKeys,
                         f->primitiveIndices, &f->rawVisibleCount);
    gsmo_prepare_header(f->rawVisibleCount, f->rawTotalInstances, f->maxGaussians, f->maxInstances, &f->header);
    t[2] = now_s();
    uint32_t V = f->header.visibleCount, I = f->header.totalInstances;
    gsmo_sort_pairs_u32(f->depthKeys, f->primitiveIndices, V, f->depthKey16 ? 2 : 4);
    t[3] = now_s();
    gsmo_apply_depth_order(f->primitiveIndices, f->nTouched, V, f->orderedTileCounts);
    gsmo_exclusive_scan(f->orderedTileCounts, V, f->orderedTileCounts);
    t[4] = now_s();
    if (stereo)
        gsmo_create_instances_stereo(f->primitiveIndices, f->orderedTileCounts, f->bounds, V, bin->tilesX,
                                     f->maxInstances, f->tileId16, f->instanceTileIds, f->instanceGaussianIndices);
    else
        gsmo_create_instances(f->primitiveIndices, f->orderedTileCounts, f->bounds,
                              (const gsmo_render_data*)f->renderData, V, bin->tilesX, bin->alphaThreshold,
                              f->maxInstances, f->tileId16, f->instanceTileIds, f->instanceGaussianIndices);
    t[5] = now_s();
    uint32_t T = bin->tilesX * bin->tilesY;
    int passes = gsmo_tile_sort_passes(T);
    if (f->tileId16) gsmo_sort_pairs_u16((uint16_t*)f->instanceTileIds, f->instanceGaussianIndices, I, passes);
    else gsmo_sort_pairs_u32((uint32_t*)f->instanceTileIds, f->instanceGaussianIndices, I, passes);
    t[6] = now_s();
    gsmo_extract_ranges(f->instanceTileIds, f->tileId16, I, T, f->tileHeaders, f->activeTiles, &f->activeTileCount);
    t[7] = now_s();
}

void gsmo_render_mono(gsmo_frame* f, const void* gaussians, const void* harmonics, int precision,
                      const gsmo_camera* cam, uint32_t width, uint32_t height, gsmo_half* color,
                      gsmo_half* depth) {
    /* DFR.swift:249: silent no-op, target untouched */
    if (cam->gaussianCount == 0 || cam->gaussianCount > f->maxGaussians) return;
    gsmo_binning bin;
    binningFor(width, height, &bin);
    double t[10];
    t[0] = now_s();
    gsmo_project_cull(gaussians, harmonics, precision, cam, &bin, (gsmo_render_data*)f->renderData, f->bounds,
                      f->preDepthKeys, f->nTouched, &f->rawTotalInstances);
    frame_common(f, &bin, 0, cam->gaussianCount, t);
    gsmo_clear(color, depth, width, height);
    gsmo_blend(f->tileHeaders, (const gsmo_render_data*)f->renderData, f->instanceGaussianIndices, f->activeTiles,
               f->activeTileCount, width, height, bin.tilesX, color, depth);
    t[8] = now_s();
    for (int i = 0; i < 8; ++i) f->stageSeconds[i] = t[i + 1] - t[i];
    f->stageSeconds[8] = 0.0;
    f->stageSeconds[9] = t[8] - t[0];
}

void gsmo_render_stereo(gsmo_frame* f, const void* gaussians, const void* harmonics, int precision,
                        const gsmo_stereo_camera* cam, uint32_t width, uint32_t height, int flipY,
                        gsmo_half* scratchColor2, gsmo_half* dstSideBySide) {
    if (cam->gaussianCount == 0 || cam->gaussianCount > f->maxGaussians) return; /* DFR.swift:478,607 */
    gsmo_binning bin;
    binningFor(width, height, &bin);
    double t[10];
    t[0] = now_s();
    gsmo_project_cull_stereo(gaussians, harmonics, precision, cam, &bin, (gsmo_stereo_render_data*)f->renderData,
                             f->bounds, f->preDepthKeys, f->nTouched, &f->rawTotalInstances);
    frame_common(f, &bin, 1, cam->gaussianCount, t);
    const size_t slice = (size_t)width * height * 4;
    gsmo_clear(scratchColor2, NULL, width, height);
    gsmo_clear(scratchColor2 + slice, NULL, width, height);
    gsmo_blend_stereo(f->tileHeaders, (const gsmo_stereo_render_data*)f->renderData, f->instanceGaussianIndices,
                      f->activeTiles, f->activeTileCount, width, height, bin.tilesX, scratchColor2);
    t[8] = now_s();
    gsmo_stereo_copy(scratchColor2, width, height, flipY, dstSideBySide);
    t[9] = now_s();
    for (int i = 0; i < 9; ++i) f->stageSeconds[i] = t[i + 1] - t[i];
    f->stageSeconds[9] = t[9] - t[0];
}

/* ================================================================================================================
 * GlobalRenderer: Sources/Renderer/GlobalRenderer/GlobalShaders.metal = "GS.metal",
 * GlobalRenderer.swift = "GR.swift". Same shared helpers as the DepthFirst path (GShared.h); what differs:
 *   - projection: ndcToScreenCentered, no far-plane exit, no tile count (GS.metal:19-125); 32 x 16 tiles;
 *   - visibility = valid bounds (GS.metal:169-208); tiles by gaussianIntersectsTile on the quantised record with
 *     opacity passed as the BYTE value (GS.metal:589-590: `float(g.opacity)` of a uchar, literal) -- two passes
 *     (count, prefix sum, scatter: GS.metal:563-680);
 *   - ONE sort of 32-bit keys [tile:16][half depth ^ 0x8000:16] (GS.metal:267-295; 3 or 4 stable LSD passes,
 *     RadixSortEncoder.swift:52-63 == a stable sort by the whole key); headers by binary search (GS.metal:304-363);
 *   - render: 8 x 8 threads of 4 x 2 pixels per 32 x 16 tile, early exit over the thread's eight pixels
 *     (GS.metal:1036-1187); clear = (0,0,0,1), depth 0 (GS.metal:140-154).
 * Tiles come from the renderer's LIMITS (maxWidth / maxHeight: GR.swift:26-49, RenderParams.width/height too), the
 * camera from the frame's width / height. PARITY UNPINNED by the reference beyond the sort KATs
 * (GlobalUnitTests.swift:23-176): its tests hold no other value and the Metal path cannot run here.
 * ================================================================================================================ */
static float gsmo_log2f_canonical(float x) { return gsmo_log(x) * 1.44269504088896341f; }

/* GShared.h:595-597 */
static float gaussianComputePower(float opacity) {
    const float LN2 = 0.693147180559945f;
    return LN2 * 8.0f + LN2 * gsmo_log2f_canonical(gsmo_fmax(opacity, 1e-6f));
}
/* GShared.h:599-604 */
static int gaussianSegmentIntersectEllipse(float a, float b, float c, float d, float l, float r) {
    float delta = b * b - 4.0f * a * c;
    float t1 = (l - d) * (2.0f * a) + b;
    float t2 = (r - d) * (2.0f * a) + b;
    return delta >= 0.0f && (t1 <= 0.0f || t1 * t1 <= delta) && (t2 >= 0.0f || t2 * t2 <= delta);
}
/* GShared.h:606-645 */
static int gaussianIntersectsTile(int minX, int minY, int maxX, int maxY, float cx, float cy, float conicX, float conicY,
                                  float conicZ, float power) {
    if (cx >= (float)minX && cx <= (float)maxX && cy >= (float)minY && cy <= (float)maxY) return 1;
    float w = 2.0f * power;
    float dx, dy, a, b, c;
    if (cx * 2.0f < (float)(minX + maxX)) dx = cx - (float)minX; else dx = cx - (float)maxX;
    a = conicZ;
    b = -2.0f * conicY * dx;
    c = conicX * dx * dx - w;
    if (gaussianSegmentIntersectEllipse(a, b, c, cy, (float)minY, (float)maxY)) return 1;
    if (cy * 2.0f < (float)(minY + maxY)) dy = cy - (float)minY; else dy = cy - (float)maxY;
    a = conicX;
    b = -2.0f * conicY * dy;
    c = conicZ * dy * dy - w;
    if (gaussianSegmentIntersectEllipse(a, b, c, cx, (float)minX, (float)maxX)) return 1;
    return 0;
}

/* GS.metal:563-680: count (emit == 0) or scatter the tiles of one visible Gaussian */
static uint32_t globalWalkTiles(const gsmo_render_data* g, const int32_t* rect, uint32_t tileW, uint32_t tileH, uint32_t tilesX,
                                int emit, int32_t* tileIds, int32_t* tileIndices, uint32_t writePos, uint32_t maxAssignments,
                                int32_t gaussianIdx) {
    const int minTX = rect[0], maxTX = rect[1], minTY = rect[2], maxTY = rect[3];
    if (minTX > maxTX || minTY > maxTY) return 0;
    const float alpha = (float)g->opacity;   /* the byte value, literal (GS.metal:589) */
    if (alpha < 1e-4f) return 0;
    const float cx = gsmo_h2f(g->meanX), cy = gsmo_h2f(g->meanY);
    const float theta = unpackThetaPi(g->theta);
    float A, B, C;
    conicFromThetaSigmas(theta, gsmo_h2f(g->sigma1), gsmo_h2f(g->sigma2), &A, &B, &C);
    const float power = gaussianComputePower(alpha);
    uint32_t n = 0;
    for (int ty = minTY; ty <= maxTY; ++ty)
        for (int tx = minTX; tx <= maxTX; ++tx) {
            const int px0 = tx * (int)tileW, py0 = ty * (int)tileH;
            if (gaussianIntersectsTile(px0, py0, px0 + (int)tileW - 1, py0 + (int)tileH - 1, cx, cy, A, B, C, power)) {
                if (!emit) n++;
                else if (writePos < maxAssignments) {
                    tileIds[writePos] = ty * (int)tilesX + tx;
                    tileIndices[writePos] = gaussianIdx;
                    writePos++;
                    n++;
                }
            }
        }
    return n;
}

void gsmo_render_global(const void* gaussians, const void* harmonics, int precision, const gsmo_camera* cam,
                        uint32_t maxWidth, uint32_t maxHeight, uint32_t maxGaussians, uint32_t width, uint32_t height,
                        gsmo_half* color, gsmo_half* depthOut, gsmo_render_data* renderData, int32_t* bounds, uint8_t* mask,
                        uint32_t* visibleIndices, uint32_t* sortedKeys, int32_t* sortedIndices, gsmo_tile_header* headers,
                        gsmo_global_info* info) {
    const uint32_t tileW = 32, tileH = 16;
    const uint32_t tilesX = (maxWidth + tileW - 1) / tileW, tilesY = (maxHeight + tileH - 1) / tileH;
    const uint32_t tileCount = tilesX * tilesY > 0 ? tilesX * tilesY : 1;
    const uint32_t maxAssignments = 4u * maxGaussians;   /* GlobalResources.swift:79-81 */
    const float alphaThreshold = 0.005f, totalInkThreshold = 2.0f;   /* GR.swift:45-46 */
    const uint32_t N = cam->gaussianCount;
    memset(info, 0, sizeof *info);

    /* GS.metal:19-125 */
#pragma omp parallel for schedule(dynamic, 4096)
    for (uint32_t gid = 0; gid < N; ++gid) {
        int32_t* rect = bounds + 4 * (size_t)gid;
        mask[gid] = 0; rect[0] = 0; rect[1] = -1; rect[2] = 0; rect[3] = -1;
        v3 position, scale; v4 rot; float opacity;
        loadGaussian(gaussians, precision, gid, &position, &scale, &rot, &opacity);
        float maxScale = gsmo_fmax(scale.x, gsmo_fmax(scale.y, scale.z));
        if (maxScale < 0.0005f) continue;
        v4 p4 = {position.x, position.y, position.z, 1.0f};
        v4 viewPos4 = mul44(cam->view, p4);
        v4 clip = mul44(cam->proj, viewPos4);
        float depth = clip.w;
        if (!(clip.w > cam->nearPlane)) continue;
        float ndcX = clip.x / clip.w, ndcY = clip.y / clip.w;
        float screenX = ((ndcX + 1.0f) * cam->width - 1.0f) * 0.5f;    /* ndcToScreenCentered, GShared.h:184-189 */
        float screenY = ((ndcY + 1.0f) * cam->height - 1.0f) * 0.5f;
        if (opacity < alphaThreshold) continue;
        v4 quat = normalizeQuaternion(rot);
        m3 cov3d = buildCovariance3D(scale, quat);
        v3 viewPos = {viewPos4.x, viewPos4.y, viewPos4.z};
        m2 cov2d = projectCovariance2D(cov3d, viewPos, cam->view, cam->proj, cam->width, cam->height);
        cov2d = stabilizeCovariance2D(cov2d, cam->width, cam->height);
        float theta, sigma1, sigma2;
        if (!covarianceToThetaSigmas(cov2d, &theta, &sigma1, &sigma2)) continue;
        float radius = 3.0f * gsmo_fmax(sigma1, sigma2);
        if (radius < 0.5f) continue;
        if (cullByTotalInkFromCov(opacity, cov2d, depth, cam->nearPlane, cam->farPlane, totalInkThreshold)) continue;
        float obbX, obbY;
        computeOBBExtents(cov2d, 3.0f, &obbX, &obbY);
        if (screenX + obbX < 0.0f || screenX - obbX > cam->width || screenY + obbY < 0.0f || screenY - obbY > cam->height) continue;
        v3 col = computeSHColor(harmonics, precision, gid, position, *(const v3*)cam->center, cam->shComponents);
        col.x = gsmo_fmax(col.x + 0.5f, 0.0f); col.y = gsmo_fmax(col.y + 0.5f, 0.0f); col.z = gsmo_fmax(col.z + 0.5f, 0.0f);
        if (cam->inputIsSRGB > 0.5f) { col.x = srgbToLinearChannel(col.x); col.y = srgbToLinearChannel(col.y); col.z = srgbToLinearChannel(col.z); }
        gsmo_render_data rd;
        rd.meanX = gsmo_f2h(screenX); rd.meanY = gsmo_f2h(screenY);
        rd.theta = packThetaPi(theta);
        rd.sigma1 = gsmo_f2h(sigma1); rd.sigma2 = gsmo_f2h(sigma2);
        rd.depth = gsmo_f2h(depth);
        rd.colorR = quantU8(col.x); rd.colorG = quantU8(col.y); rd.colorB = quantU8(col.z); rd.opacity = quantU8(opacity);
        renderData[gid] = rd;
        tile_bounds tb = computeTileBounds(screenX, screenY, obbX, obbY, cam->width, cam->height, (int)tileW, (int)tileH,
                                           (int)tilesX, (int)tilesY);
        rect[0] = tb.minTX; rect[1] = tb.maxTX; rect[2] = tb.minTY; rect[3] = tb.maxTY;
        mask[gid] = 1;
    }
    /* GS.metal:169-208: visible = valid bounds, compacted in gid order */
    uint32_t V = 0;
    for (uint32_t gid = 0; gid < N; ++gid) {
        const int32_t* r = bounds + 4 * (size_t)gid;
        if (r[0] <= r[1] && r[2] <= r[3]) visibleIndices[V++] = gid;
    }
    info->visibleCount = V;
    /* GS.metal:563-621 + prefix sum: offsets of the UNCLAMPED counts */
    uint32_t* offsets = (uint32_t*)malloc(((size_t)V + 1) * sizeof(uint32_t));
#pragma omp parallel for schedule(dynamic, 1024)
    for (uint32_t i = 0; i < V; ++i) {
        const uint32_t g = visibleIndices[i];
        offsets[i] = globalWalkTiles(&renderData[g], bounds + 4 * (size_t)g, tileW, tileH, tilesX, 0, NULL, NULL, 0, 0, 0);
    }
    uint32_t total = 0;
    for (uint32_t i = 0; i < V; ++i) { const uint32_t c = offsets[i]; offsets[i] = total; total += c; }
    /* GS.metal:623-678: scatter, bounded by maxAssignments per store */
    int32_t* tileIds = (int32_t*)malloc(((size_t)maxAssignments + 1) * sizeof(int32_t));
    int32_t* tileIndices = (int32_t*)malloc(((size_t)maxAssignments + 1) * sizeof(int32_t));
#pragma omp parallel for schedule(dynamic, 1024)
    for (uint32_t i = 0; i < V; ++i) {
        const uint32_t g = visibleIndices[i];
        globalWalkTiles(&renderData[g], bounds + 4 * (size_t)g, tileW, tileH, tilesX, 1, tileIds, tileIndices, offsets[i], maxAssignments,
                        (int32_t)g);
    }
    free(offsets);
    /* GS.metal:685-703: clamp, overflow flag, padded count */
    if (total > maxAssignments) { total = maxAssignments; info->overflow = 1; }
    info->totalAssignments = total;
    info->paddedCount = ((total + 1023u) / 1024u) * 1024u;
    /* GS.metal:267-295 */
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < total; ++i) {
        const int32_t g = tileIndices[i];
        const uint32_t depthBits = (uint32_t)gsmo_f2h(gsmo_h2f(renderData[g].depth)) ^ 0x8000u;
        sortedKeys[i] = ((uint32_t)tileIds[i] << 16) | (depthBits & 0xFFFFu);
        sortedIndices[i] = g;
    }
    free(tileIds); free(tileIndices);
    gsmo_sort_pairs_u32(sortedKeys, sortedIndices, total, 4);
    /* GS.metal:304-363 (the active list's order is the atomic's: here ascending tile) */
    uint32_t* activeTiles = (uint32_t*)malloc((size_t)tileCount * sizeof(uint32_t));
    uint32_t activeCount = 0, cursor = 0;
    for (uint32_t t = 0; t < tileCount; ++t) {
        while (cursor < total && (sortedKeys[cursor] >> 16) < t) cursor++;
        uint32_t end = cursor;
        while (end < total && (sortedKeys[end] >> 16) == t) end++;
        headers[t].offset = total ? cursor : 0; headers[t].count = end - cursor;
        if (end > cursor) activeTiles[activeCount++] = t;
        cursor = end;
    }
    info->activeTileCount = activeCount;
    /* GS.metal:140-154 */
    gsmo_clear(color, depthOut, width, height);
    /* GS.metal:1036-1187 */
    const gsmo_half h255 = gsmo_f2h(255.0f);
    const gsmo_half thr = gsmo_hdiv(H_ONE, h255);
    const gsmo_half h099 = gsmo_f2h(0.99f);
    const uint32_t W = maxWidth, Hh = maxHeight;   /* RenderParams.width / height are the limits (GR.swift:26-28) */
#pragma omp parallel for schedule(dynamic, 2)
    for (uint32_t ti = 0; ti < activeCount; ++ti) {
        const uint32_t tile = activeTiles[ti];
        const gsmo_tile_header hdr = headers[tile];
        const uint32_t tileX = tile % tilesX, tileY = tile / tilesX;
        for (uint32_t ly = 0; ly < 8; ++ly)
            for (uint32_t lx = 0; lx < 8; ++lx) {
                const uint32_t baseX = tileX * 32 + lx * 4, baseY = tileY * 16 + ly * 2;
                /* pixel k = x + 4 * y of the thread's 4 x 2 block */
                gsmo_half px[8], py[8], trans[8], col[8][3], dep[8];
                for (int k = 0; k < 8; ++k) {
                    px[k] = h_from_uint(baseX + (uint32_t)(k & 3));
                    py[k] = h_from_uint(baseY + (uint32_t)(k >> 2));
                    trans[k] = H_ONE; col[k][0] = col[k][1] = col[k][2] = H_ZERO; dep[k] = H_ZERO;
                }
                for (uint32_t i = 0; i < hdr.count; ++i) {
                    gsmo_half m0 = gsmo_hmax(gsmo_hmax(trans[0], trans[1]), gsmo_hmax(trans[2], trans[3]));
                    gsmo_half m1 = gsmo_hmax(gsmo_hmax(trans[4], trans[5]), gsmo_hmax(trans[6], trans[7]));
                    if (h_lt(gsmo_hmax(m0, m1), thr)) break;
                    const int32_t gi = sortedIndices[hdr.offset + i];
                    if (gi < 0) continue;
                    const gsmo_render_data g = renderData[gi];
                    float A, B, C;
                    conicFromThetaSigmas(unpackThetaPi(g.theta), gsmo_h2f(g.sigma1), gsmo_h2f(g.sigma2), &A, &B, &C);
                    const gsmo_half cxx = gsmo_f2h(A), cyy = gsmo_f2h(C), cxy2 = gsmo_f2h(2.0f * B);
                    const gsmo_half opacity = gsmo_hdiv(gsmo_f2h((float)g.opacity), h255);
                    const gsmo_half gc[3] = {gsmo_hdiv(gsmo_f2h((float)g.colorR), h255), gsmo_hdiv(gsmo_f2h((float)g.colorG), h255),
                                             gsmo_hdiv(gsmo_f2h((float)g.colorB), h255)};
                    gsmo_half a[8];
                    int allZero = 1;
                    for (int k = 0; k < 8; ++k) {
                        const gsmo_half dx = gsmo_hsub(px[k], g.meanX), dy = gsmo_hsub(py[k], g.meanY);
                        const gsmo_half p = h_power(dx, dy, cxx, cyy, cxy2);
                        a[k] = gsmo_hmin(gsmo_hmul(opacity, gsmo_hexp(gsmo_hmul(H_NEG_HALF, p))), h099);
                        if (!h_eq0(a[k])) allZero = 0;
                    }
                    if (allZero) continue;
                    for (int k = 0; k < 8; ++k) {
                        const gsmo_half w = gsmo_hmul(a[k], trans[k]);
                        for (int c = 0; c < 3; ++c) col[k][c] = h_mad(gc[c], w, col[k][c]);
                        dep[k] = h_mad(g.depth, w, dep[k]);
                    }
                    for (int k = 0; k < 8; ++k) trans[k] = gsmo_hmul(trans[k], gsmo_hsub(H_ONE, a[k]));
                }
                for (int k = 0; k < 8; ++k) {
                    const uint32_t x = baseX + (uint32_t)(k & 3), y = baseY + (uint32_t)(k >> 2);
                    if (x < W && y < Hh && x < width && y < height) {   /* texture writes outside the target are dropped */
                        const size_t o = (size_t)y * width + x;
                        color[4 * o + 0] = col[k][0]; color[4 * o + 1] = col[k][1]; color[4 * o + 2] = col[k][2];
                        color[4 * o + 3] = gsmo_hsub(H_ONE, trans[k]);
                        if (depthOut) depthOut[o] = dep[k];
                    }
                }
            }
    }
    free(activeTiles);
}
