/*
 * gsmo_math.h -- deterministic scalar math used by the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or executed by the
 * product path (gsm_renderer_b200/csrc); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Why this file exists: the reference's Metal kernels were compiled with -ffast-math
 * (compile_shaders.sh:45-53) and call MSL built-ins (fast::sincos, log, atan2, exp(half),
 * fast::powr, normalize) whose bits are not specified.  "Bit-exact tile counts" between a
 * CPU oracle and a CUDA kernel therefore needs ONE written-down definition of those
 * built-ins.  This header is that definition (the "canonical semantics"):
 * only IEEE-754 binary32 + - * / sqrt (round-to-nearest-even, no contraction, no FTZ),
 * explicit fmaf where stated, and fixed-coefficient polynomials (Cephes single-precision
 * coefficient sets, Moshier, public domain; the exp2 set is a degree-5 Chebyshev fit).
 * The CUDA side restates the same definitions in gsm_renderer_b200/csrc/gsm_dmath.cuh;
 * tests compare the two implementations exhaustively (all 65536 halfs for exp_h) and on
 * dense sweeps (sincos/log/atan2).
 *
 * Build flags required: -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#ifndef GSMO_MATH_H
#define GSMO_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef uint16_t gsmo_half; /* raw IEEE binary16 bits */

static inline uint32_t gsmo_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float gsmo_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* half <-> float, round-to-nearest-even, overflow to +-inf, subnormals kept (MSL half(x)). */
static inline float gsmo_h2f(gsmo_half h) { _Float16 v; memcpy(&v, &h, 2); return (float)v; }
static inline gsmo_half gsmo_f2h(float f) { _Float16 v = (_Float16)f; gsmo_half h; memcpy(&h, &v, 2); return h; }

/* half arithmetic: each op = exact result rounded once to binary16.  Computing in binary32
 * and rounding again is exact for + - * / (24 >= 2*11+2 bits), so this equals a native
 * half ALU (CUDA __hadd_rn / __hmul_rn). */
static inline gsmo_half gsmo_hadd(gsmo_half a, gsmo_half b) { return gsmo_f2h(gsmo_h2f(a) + gsmo_h2f(b)); }
static inline gsmo_half gsmo_hsub(gsmo_half a, gsmo_half b) { return gsmo_f2h(gsmo_h2f(a) - gsmo_h2f(b)); }
static inline gsmo_half gsmo_hmul(gsmo_half a, gsmo_half b) { return gsmo_f2h(gsmo_h2f(a) * gsmo_h2f(b)); }
static inline gsmo_half gsmo_hdiv(gsmo_half a, gsmo_half b) { return gsmo_f2h(gsmo_h2f(a) / gsmo_h2f(b)); }
/* One correctly rounded (nearest-even) conversion of a finite-or-not double to binary16, done by hand:
 * going through float would round twice (53 -> 24 -> 11 bits), which is not innocuous for an arbitrary double. */
static inline gsmo_half gsmo_d2h(double v) {
    uint64_t u;
    memcpy(&u, &v, 8);
    const gsmo_half sign = (gsmo_half)((u >> 48) & 0x8000u);
    u &= 0x7FFFFFFFFFFFFFFFull;
    if (u > 0x7FF0000000000000ull) return 0x7FFFu;                       /* NaN */
    if (u >= 0x40EFFE0000000000ull) return (gsmo_half)(sign | 0x7C00u);  /* |v| >= 65520: +-inf */
    const int E = (int)(u >> 52) - 1023;
    if (E < -25) return sign;                                            /* |v| < 2^-25: +-0 */
    const uint64_t m = (u & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
    const int shift = (E >= -14) ? 42 : 42 + (-14 - E);                  /* subnormal halfs keep fewer bits */
    uint64_t q = m >> shift;
    const uint64_t rem = m & ((1ull << shift) - 1ull), half = 1ull << (shift - 1);
    if (rem > half || (rem == half && (q & 1ull))) q++;                  /* nearest, ties to even */
    /* normal: biased exponent (E+15) with the implicit bit folded in, so a mantissa carry bumps the exponent */
    const uint32_t h = (E >= -14) ? (uint32_t)(((uint32_t)(E + 14) << 10) + (uint32_t)q) : (uint32_t)q;
    return (gsmo_half)(sign | (gsmo_half)h);
}

/* fused multiply-add in binary16: a*b + c rounded ONCE (device: HFMA2). Used where the Metal compiler contracts
 * a*b + c under -ffast-math. a*b is exact in double (22 significant bits); the sum is exact in double unless the
 * operands are more than 53 bits apart, which an error-free TwoSum detects -- only then the x87 long double
 * (64-bit significand >= the 56-bit span of a half product plus a half addend) path is taken. */
static inline gsmo_half gsmo_hfma(gsmo_half a, gsmo_half b, gsmo_half c) {
    const float fa = gsmo_h2f(a), fb = gsmo_h2f(b), fc = gsmo_h2f(c);
    {
        /* Fast path. a*b is exact in fp32 (22 bits). t = RN32(a*b + c) then RN16(t) equals RN16(a*b + c) unless t
         * sits exactly on a midpoint between two halfs: every such midpoint is an fp32 number and RN32 is monotonic,
         * so t and the exact sum lie on the same side of every midpoint that t is not equal to. For a normal half
         * result (|t| >= 2^-14) the midpoints are the fp32 values whose low 13 mantissa bits are 0x1000. */
        const float t = fa * fb + fc;
        uint32_t u;
        memcpy(&u, &t, 4);
        const uint32_t mag = u & 0x7FFFFFFFu;
        if (mag >= 0x38800000u && mag < 0x7F800000u && (u & 0x1FFFu) != 0x1000u) return gsmo_f2h(t);
    }
    const double ab = (double)fa * (double)fb;
    const double cd = (double)fc;
    const double s = ab + cd;
    if (s - s == 0.0) {                                         /* finite */
        const double bb = s - ab;
        const double err = (ab - (s - bb)) + (cd - bb);
        if (err == 0.0) return gsmo_d2h(s);
        long double r = (long double)ab + (long double)cd;      /* exact */
        _Float16 v = (_Float16)r;
        gsmo_half h;
        memcpy(&h, &v, 2);
        return h;
    }
    return gsmo_f2h((float)s);                                  /* inf / NaN propagate */
}
static inline int gsmo_hisnan(gsmo_half a) { return (a & 0x7FFFu) > 0x7C00u; }
/* min/max: a NaN operand loses (IEEE-754-2008 minNum/maxNum, MSL fmin/fmax); -0 orders below +0
 * (the rule of PTX min/max, so the device uses one FMNMX / HMNMX2 instruction). */
static inline gsmo_half gsmo_hmin(gsmo_half a, gsmo_half b) {
    if (gsmo_hisnan(a)) return b;
    if (gsmo_hisnan(b)) return a;
    float fa = gsmo_h2f(a), fb = gsmo_h2f(b);
    if (fa == fb) return (a & 0x8000u) ? a : b;
    return (fa < fb) ? a : b;
}
static inline gsmo_half gsmo_hmax(gsmo_half a, gsmo_half b) {
    if (gsmo_hisnan(a)) return b;
    if (gsmo_hisnan(b)) return a;
    float fa = gsmo_h2f(a), fb = gsmo_h2f(b);
    if (fa == fb) return (a & 0x8000u) ? b : a;
    return (fa > fb) ? a : b;
}

static inline float gsmo_fmax(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return (gsmo_f2u(a) & 0x80000000u) ? b : a;
    return (a > b) ? a : b;
}
static inline float gsmo_fmin(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return (gsmo_f2u(a) & 0x80000000u) ? a : b;
    return (a < b) ? a : b;
}
static inline float gsmo_clamp(float x, float lo, float hi) { return gsmo_fmin(gsmo_fmax(x, lo), hi); }
static inline int gsmo_isfinite(float x) { return (gsmo_f2u(x) & 0x7F800000u) != 0x7F800000u; }

#define GSMO_PI_F 3.14159265358979323846f /* kPiF, GaussianShared.h:432 (rounds to 0x40490FDB) */

/* sin and cos of one argument (MSL fast::sincos, GaussianShared.h:495,571).
 * k = nearest multiple of pi/2, Cody-Waite 3-constant reduction, Cephes sinf/cosf kernels. */
static inline void gsmo_sincos(float x, float* sn, float* cs) {
    float ax = fabsf(x);
    float kf = floorf(ax * 0.636619772367581343f + 0.5f);
    int k = (int)kf;
    float r = ax - kf * 1.5703125f;               /* pi/2 split: hi (exact product for k < 2^15) */
    r = r - kf * 4.837512969970703125e-4f;        /* mid */
    r = r - kf * 7.54978995489188216e-8f;         /* lo  */
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
    *sn = s;
    *cs = c;
}

/* natural log for x > 0 (MSL log, GaussianShared.h:592). Cephes logf. Subnormals are not
 * pre-scaled (never reached: the argument is tau/opacity in [0.005/1, 1]). */
static inline float gsmo_log(float x) {
    uint32_t u = gsmo_f2u(x);
    int e = (int)((u >> 23) & 0xFFu) - 126;
    float m = gsmo_u2f((u & 0x807FFFFFu) | 0x3F000000u); /* frexp mantissa in [0.5,1) */
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

/* e^x in binary32 (used only by powr below). Cephes expf without the overflow branches:
 * callers keep |x| < 80. */
static inline float gsmo_exp(float x) {
    float n = floorf(1.44269504088896341f * x + 0.5f);
    x = x - n * 0.693359375f;
    x = x - n * -2.12194440e-4f;
    float z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x
           + 4.1665795894e-2f) * x + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    int ni = (int)n;
    return gsmo_u2f(gsmo_f2u(z) + ((uint32_t)ni << 23)); /* ldexp for normal results */
}

/* fast::powr(x, y) for x > 0 (GaussianShared.h:120). */
static inline float gsmo_powr(float x, float y) { return gsmo_exp(y * gsmo_log(x)); }

/* atan for any finite t. Cephes atanf. */
static inline float gsmo_atan(float t) {
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

/* atan2(y, x) (GaussianShared.h:479). Result in [-pi, pi]. */
static inline float gsmo_atan2(float y, float x) {
    if (x != x || y != y) return gsmo_u2f(0x7FC00000u);
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float a = gsmo_atan(y / x);
    if (x < 0.0f) {
        if (y < 0.0f) return a - GSMO_PI_F;
        return a + GSMO_PI_F;
    }
    return a;
}

/* fmod(t, kPiF) for |t| < 4*pi (GaussianShared.h:436,481): exact, sign of t. */
static inline float gsmo_fmod_pi(float t) {
    if (!gsmo_isfinite(t)) return gsmo_u2f(0x7FC00000u);
    float a = fabsf(t);
    while (a >= GSMO_PI_F) a = a - GSMO_PI_F;
    return (t < 0.0f) ? -a : a;
}

/* exp(half) -> half (MSL exp on half, DepthFirstShaders.metal:1775).  binary32 evaluation of
 * 2^(x*log2 e): round-to-nearest integer part by the 1.5*2^23 trick, degree-5 Chebyshev
 * polynomial (max rel err 1.6e-7) evaluated with fmaf, exponent added in the integer domain,
 * single final rounding to binary16 (gives subnormals, 0 and +inf). */
static inline gsmo_half gsmo_hexp(gsmo_half xh) {
    if (gsmo_hisnan(xh)) return 0x7FFFu;
    float x = gsmo_h2f(xh);
    if (x < -17.5f) return 0x0000u; /* e^-17.5 < 2^-25: rounds to +0 */
    if (x > 11.5f) return 0x7C00u;  /* e^11.5 > 65520: rounds to +inf */
    float t = x * 1.44269504088896341f;
    float zb = t + 12582912.0f;
    float n = zb - 12582912.0f;
    float f = t - n;
    float p = 0x1.5f0890p-10f;
    p = fmaf(p, f, 0x1.3d1070p-7f);
    p = fmaf(p, f, 0x1.c6af6cp-5f);
    p = fmaf(p, f, 0x1.ebf906p-3f);
    p = fmaf(p, f, 0x1.62e430p-1f);
    p = fmaf(p, f, 0x1.000002p+0f);
    float r = gsmo_u2f(gsmo_f2u(p) + (gsmo_f2u(zb) << 23));
    return gsmo_f2h(r);
}

#endif /* GSMO_MATH_H */
