// gsm_dmath.cuh -- canonical transcendental definitions on the device.
//
// The reference's MSL built-ins (fast::sincos, log, atan2, exp(half), fast::powr, normalize) have no
// specified bits (Metal -ffast-math, compile_shaders.sh:45-53), so tile counts can only be bit-exact
// against a CPU checker if both sides evaluate ONE written-down definition. These are those definitions:
// IEEE binary32 + - * / sqrt in a fixed order (this TU is compiled with -fmad=false, default
// -prec-div/-prec-sqrt/-ftz=false), explicit __fmaf_rn where stated, fixed polynomial coefficients
// (Cephes single-precision sets; degree-5 Chebyshev fit for 2^f). No libdevice transcendental is called.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace gsm {

#define GSM_PI_F 3.14159265358979323846f            // kPiF, GaussianShared.h:432
#define GSM_THETA_PACK 0x1.45f1c0p+14f              // binary32(65535.0f / kPiF), GaussianShared.h:438
#define GSM_THETA_UNPACK 0x1.922148p-15f            // binary32(kPiF / 65535.0f), GaussianShared.h:443

// min/max: a NaN operand loses, -0 orders below +0 -- exactly PTX min.f32/max.f32, one FMNMX each
// (the oracle states the same rule in gsmo_fmin/gsmo_fmax; probe ops 9/10 compare them bit for bit).
__device__ __forceinline__ float dmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float dmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float dclamp(float x, float lo, float hi) { return dmin(dmax(x, lo), hi); }
__device__ __forceinline__ bool dfinite(float x) { return (__float_as_uint(x) & 0x7F800000u) != 0x7F800000u; }

// fast::sincos (GaussianShared.h:495,571)
__device__ __forceinline__ void dsincos(float x, float& sn, float& cs) {
    float ax = fabsf(x);
    float kf = floorf(ax * 0.636619772367581343f + 0.5f);
    int k = (int)kf;
    float r = ax - kf * 1.5703125f;
    r = r - kf * 4.837512969970703125e-4f;
    r = r - kf * 7.54978995489188216e-8f;
    float z = r * r;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
               - 0.5f * z + 1.0f;
    float s, c;
    switch (k & 3) {
        case 0: s = ps; c = pc; break;
        case 1: s = pc; c = -ps; break;
        case 2: s = -ps; c = -pc; break;
        default: s = -pc; c = ps; break;
    }
    if (x < 0.0f) s = -s;
    sn = s;
    cs = c;
}

// log (GaussianShared.h:592)
__device__ __forceinline__ float dlog(float x) {
    uint32_t u = __float_as_uint(x);
    int e = (int)((u >> 23) & 0xFFu) - 126;
    float m = __uint_as_float((u & 0x807FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = m + m - 1.0f;
    } else {
        m = m - 1.0f;
    }
    float z = m * m;
    float y = ((((((((7.0376836292e-2f * m - 1.1514610310e-1f) * m + 1.1676998740e-1f) * m
                    - 1.2420140846e-1f) * m + 1.4249322787e-1f) * m - 1.6668057665e-1f) * m
                 + 2.0000714765e-1f) * m - 2.4999993993e-1f) * m + 3.3333331174e-1f) * m * z;
    float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    z = m + y;
    z = z + 0.693359375f * fe;
    return z;
}

__device__ __forceinline__ float dexp(float x) {
    float n = floorf(1.44269504088896341f * x + 0.5f);
    x = x - n * 0.693359375f;
    x = x - n * -2.12194440e-4f;
    float z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x
           + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    int ni = (int)n;
    return __uint_as_float(__float_as_uint(z) + ((uint32_t)ni << 23));
}

// fast::powr (GaussianShared.h:120)
__device__ __forceinline__ float dpowr(float x, float y) { return dexp(y * dlog(x)); }

__device__ __forceinline__ float datan(float t) {
    float at = fabsf(t);
    float y0, u;
    if (at > 2.414213562373095f) {
        y0 = 1.5707963267948966f;
        u = -(1.0f / at);
    } else if (at > 0.4142135623730950f) {
        y0 = 0.7853981633974483f;
        u = (at - 1.0f) / (at + 1.0f);
    } else {
        y0 = 0.0f;
        u = at;
    }
    float z = u * u;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z
               - 3.33329491539e-1f) * z * u + u;
    float r = y0 + p;
    return (t < 0.0f) ? -r : r;
}

// atan2 (GaussianShared.h:479)
__device__ __forceinline__ float datan2(float y, float x) {
    if (x != x || y != y) return __uint_as_float(0x7FC00000u);
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float a = datan(y / x);
    if (x < 0.0f) {
        if (y < 0.0f) return a - GSM_PI_F;
        return a + GSM_PI_F;
    }
    return a;
}

// fmod(t, kPiF) (GaussianShared.h:436,481)
__device__ __forceinline__ float dfmod_pi(float t) {
    if (!dfinite(t)) return __uint_as_float(0x7FC00000u);
    float a = fabsf(t);
    while (a >= GSM_PI_F) a = a - GSM_PI_F;
    return (t < 0.0f) ? -a : a;
}

// exp(half) -> half for one lane, evaluated in binary32 (DepthFirstShaders.metal:1775).
__device__ __forceinline__ float dhexp_f32(float x) {  // x is an exact half value; returns the binary32 pre-rounding value
    float xc = fminf(fmaxf(x, -17.5f), 11.5f);         // keeps the integer trick in range; selects below fix the ends
    float t = xc * 1.44269504088896341f;
    float zb = t + 12582912.0f;
    float n = zb - 12582912.0f;
    float f = t - n;
    float p = 0x1.5f0890p-10f;
    p = __fmaf_rn(p, f, 0x1.3d1070p-7f);
    p = __fmaf_rn(p, f, 0x1.c6af6cp-5f);
    p = __fmaf_rn(p, f, 0x1.ebf906p-3f);
    p = __fmaf_rn(p, f, 0x1.62e430p-1f);
    p = __fmaf_rn(p, f, 0x1.000002p+0f);
    float r = __uint_as_float(__float_as_uint(p) + (__float_as_uint(zb) << 23));
    r = (x < -17.5f) ? 0.0f : r;
    r = (x > 11.5f) ? __uint_as_float(0x7F800000u) : r;
    return r;
}
__device__ __forceinline__ __half dhexp(__half xh) {
    if (__hisnan(xh)) return __ushort_as_half((unsigned short)0x7FFFu);
    return __float2half_rn(dhexp_f32(__half2float(xh)));
}
__device__ __forceinline__ __half2 dhexp2(__half2 x) {
    float2 xf = __half22float2(x);
    float rx = dhexp_f32(xf.x), ry = dhexp_f32(xf.y);
    __half2 r = __floats2half2_rn(rx, ry);
    // NaN lanes: canonical 0x7FFF
    if (xf.x != xf.x) r = __halves2half2(__ushort_as_half((unsigned short)0x7FFFu), __high2half(r));
    if (xf.y != xf.y) r = __halves2half2(__low2half(r), __ushort_as_half((unsigned short)0x7FFFu));
    return r;
}

// The blend kernel's form of dhexp2: packed binary32 arithmetic (sm_100 FMUL2/FADD2/FFMA2) and no end
// selects. Bit-identical to dhexp2 on every non-NaN input: the clamp is done in half (exact), e^-17.5 < 2^-25
// rounds to +0 and e^11.5 > 65520 rounds to +inf, which are the values dhexp2 selects. A NaN lane comes out
// as +inf instead of NaN; the caller's min(opacity * e, 0.99h) maps both to 0.99h (opacity > 0).
__device__ __forceinline__ __half2 dhexp2_packed(__half2 x) {
    const __half2 xc = __hmax2(__hmin2(x, __float2half2_rn(11.5f)), __float2half2_rn(-17.5f));
    const float2 xf = __half22float2(xc);
    const float2 t = __fmul2_rn(xf, make_float2(1.44269504088896341f, 1.44269504088896341f));
    const float2 zb = __fadd2_rn(t, make_float2(12582912.0f, 12582912.0f));
    const float2 n = __fadd2_rn(zb, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __ffma2_rn(n, make_float2(-1.0f, -1.0f), t);  // t - n (the product is exact)
    float2 p = make_float2(0x1.5f0890p-10f, 0x1.5f0890p-10f);
    p = __ffma2_rn(p, f, make_float2(0x1.3d1070p-7f, 0x1.3d1070p-7f));
    p = __ffma2_rn(p, f, make_float2(0x1.c6af6cp-5f, 0x1.c6af6cp-5f));
    p = __ffma2_rn(p, f, make_float2(0x1.ebf906p-3f, 0x1.ebf906p-3f));
    p = __ffma2_rn(p, f, make_float2(0x1.62e430p-1f, 0x1.62e430p-1f));
    p = __ffma2_rn(p, f, make_float2(0x1.000002p+0f, 0x1.000002p+0f));
    const float rx = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(zb.x) << 23));
    const float ry = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(zb.y) << 23));
    return __floats2half2_rn(rx, ry);
}

// exp(-0.5h * p) for the blend's alpha (DFS.metal:1775), with the -0.5 folded into the binary32 constant: -0.5h * p is exact in
// half except for the smallest subnormals of p (whose exp is 1.0h whichever way the product rounds), scaling a binary32 by 0.5
// is exact, and the clamp of x to [-17.5, 11.5] is the clamp of p to [-23, 35]. Bit-identical to dhexp2_packed(-0.5h * p) on
// all 65536 inputs (tests: gsm_probe_math op 12 against the oracle); two packed FMA-pipe multiplies fewer per splat.
__device__ __forceinline__ __half2 dhexp2_neghalf_packed(__half2 p) {
    const __half2 pc = __hmin2(__hmax2(p, __float2half2_rn(-23.0f)), __float2half2_rn(35.0f));
    const float2 pf = __half22float2(pc);
    const float2 t = __fmul2_rn(pf, make_float2(-0.5f * 1.44269504088896341f, -0.5f * 1.44269504088896341f));
    const float2 zb = __fadd2_rn(t, make_float2(12582912.0f, 12582912.0f));
    const float2 n = __fadd2_rn(zb, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __ffma2_rn(n, make_float2(-1.0f, -1.0f), t);  // t - n (the product is exact)
    float2 q = make_float2(0x1.5f0890p-10f, 0x1.5f0890p-10f);
    q = __ffma2_rn(q, f, make_float2(0x1.3d1070p-7f, 0x1.3d1070p-7f));
    q = __ffma2_rn(q, f, make_float2(0x1.c6af6cp-5f, 0x1.c6af6cp-5f));
    q = __ffma2_rn(q, f, make_float2(0x1.ebf906p-3f, 0x1.ebf906p-3f));
    q = __ffma2_rn(q, f, make_float2(0x1.62e430p-1f, 0x1.62e430p-1f));
    q = __ffma2_rn(q, f, make_float2(0x1.000002p+0f, 0x1.000002p+0f));
    const float rx = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(zb.x) << 23));
    const float ry = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(zb.y) << 23));
    return __floats2half2_rn(rx, ry);
}

// exp(-0.5h * p) with the exponential on the XU pipe (MUFU.EX2) instead of the FMA-pipe polynomial, bit-identical to
// dhexp2_neghalf_packed by construction: t is the same binary32 product the polynomial starts from; e = ex2.approx(t) is within a
// few binary32 ulp of the polynomial's pre-rounding value, so the two can only round to different halfs when e sits that close to a
// half rounding boundary. The guard rounds e(1 - 2^-21) and e(1 + 2^-21) to half (normal, subnormal and zero results alike): equal
// -> that half is the answer; different (about 1 value in 1000) -> the lane pair evaluates the polynomial. gsm_probe_math ops
// 14-16 run this on all 65 536 inputs against the oracle (tests/test_gpu_parity.py::test_math_probes_bit_exact), which is the proof;
// op 15 (no guard) shows the inputs the guard exists for. NaN lanes give NaN here and +inf in the polynomial: the caller's
// min(opacity * e, 0.99h) maps both to 0.99h.
__device__ __forceinline__ float dex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ bool dhexp2_neghalf_mufu_try(__half2 p, __half2& out) {
    const float2 pf = __half22float2(p);
    const float2 t = __fmul2_rn(pf, make_float2(-0.5f * 1.44269504088896341f, -0.5f * 1.44269504088896341f));
    const float2 e = make_float2(dex2_approx(t.x), dex2_approx(t.y));
    const float2 lo = __fmul2_rn(e, make_float2(1.0f - 0x1p-21f, 1.0f - 0x1p-21f));
    const float2 hi = __fmul2_rn(e, make_float2(1.0f + 0x1p-21f, 1.0f + 0x1p-21f));
    const __half2 hl = __floats2half2_rn(lo.x, lo.y), hh = __floats2half2_rn(hi.x, hi.y);
    out = hl;
    return *reinterpret_cast<const uint32_t*>(&hl) == *reinterpret_cast<const uint32_t*>(&hh);
}
__device__ __forceinline__ __half2 dhexp2_neghalf_mufu(__half2 p) {
    __half2 r;
    if (!dhexp2_neghalf_mufu_try(p, r)) r = dhexp2_neghalf_packed(p);
    return r;
}
// the unguarded form (probe / A-B builds only: NOT bit-exact)
__device__ __forceinline__ __half2 dhexp2_neghalf_mufu_raw(__half2 p) {
    const float2 pf = __half22float2(p);
    const float2 t = __fmul2_rn(pf, make_float2(-0.5f * 1.44269504088896341f, -0.5f * 1.44269504088896341f));
    return __floats2half2_rn(dex2_approx(t.x), dex2_approx(t.y));
}

// The form the blend kernels run: exp(-0.5h * p) for two packed pairs with NO guard. t = fma.rm(float(p), c, 3 * 2^-24) -- the
// canonical constant, a round-down fused add of a fixed offset -- then MUFU.EX2 and one rounding to half. A search over the offset and
// the fma's rounding mode on the device (tools/micro/mufu_tune.cu, profiles/r2_mufu_tune.txt) found this form equal to the canonical
// polynomial on every one of the 65 536 half inputs except p = 0x297F (0.04294), which is tested for (one packed compare per pair)
// and sent to the polynomial. The domain is finite and MUFU.EX2 is a fixed function of its input bits, so the exhaustive comparison
// (gsm_probe_math op 17 against the oracle in tests/test_gpu_parity.py, and blendExpSelfTest on every device at renderer creation --
// a device whose MUFU differs runs the polynomial) is the proof, not an error bound.
constexpr float kExpTunedOffset = 3.0f * 0x1p-24f;
constexpr unsigned short kExpTunedException = 0x297Fu;
__device__ __forceinline__ __half2 dhexp2_neghalf_tuned_raw(__half2 p) {
    const float2 t = __ffma2_rd(__half22float2(p), make_float2(-0.5f * 1.44269504088896341f, -0.5f * 1.44269504088896341f),
                                make_float2(kExpTunedOffset, kExpTunedOffset));
    return __floats2half2_rn(dex2_approx(t.x), dex2_approx(t.y));
}
// true if any of the four halves is the exceptional input (or a NaN): two packed compares with two predicate results each (HSETP2), chained through the predicate input
__device__ __forceinline__ bool dhexp2_tuned_exception2(__half2 p0, __half2 p1) {
    const uint32_t xb = (uint32_t)kExpTunedException | ((uint32_t)kExpTunedException << 16);
    uint32_t r;
    asm("{ .reg .pred a, b, c, d;\n\t"
        "setp.ne.f16x2 a|b, %1, %3;\n\t"
        "setp.ne.and.f16x2 c|d, %2, %3, a;\n\t"
        "and.pred c, c, d;\n\t"
        "and.pred c, c, b;\n\t"
        "selp.u32 %0, 0, 1, c; }"
        : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&p0)), "r"(*reinterpret_cast<const uint32_t*>(&p1)), "r"(xb));
    return r != 0u;
}
__device__ __forceinline__ void dhexp2_neghalf_tuned(__half2 p0, __half2 p1, __half2& e0, __half2& e1) {
    e0 = dhexp2_neghalf_tuned_raw(p0);
    e1 = dhexp2_neghalf_tuned_raw(p1);
    if (dhexp2_tuned_exception2(p0, p1)) {
        e0 = dhexp2_neghalf_packed(p0);
        e1 = dhexp2_neghalf_packed(p1);
    }
}

}  // namespace gsm
