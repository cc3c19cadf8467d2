// gsm_tiletest.cuh -- the exact ellipse-vs-tile test on the QUANTISED splat (quirk Q3), shared by the
// tile-count loop of stage 1 (DFS.metal:166-205) and the instance expansion of stage 5 (DFS.metal:667-715).
#pragma once
#include "gsm_common.cuh"
#include "gsm_dmath.cuh"

namespace gsm {

struct QuantSplat {
    float meanX, meanY;     // float(half mean)
    float sn, cs;           // sincos(unpackThetaPi(theta))
    float sigma1, sigma2;   // float(half sigma)
    float ca, cb, cc;       // conicFromSigmaTheta (GaussianShared.h:569-585)
    float d2Cutoff;         // computeD2Cutoff (GaussianShared.h:590-593)
};

__device__ __forceinline__ QuantSplat makeQuantSplat(__half meanX, __half meanY, uint16_t thetaP, __half s1, __half s2,
                                                     uint8_t opacityU8) {
    QuantSplat q;
    q.meanX = __half2float(meanX);
    q.meanY = __half2float(meanY);
    float theta_q = (float)thetaP * GSM_THETA_UNPACK;  // unpackThetaPi, GaussianShared.h:442-444
    q.sigma1 = __half2float(s1);
    q.sigma2 = __half2float(s2);
    float opacity_q = (float)opacityU8 * (1.0f / 255.0f);
    dsincos(theta_q, q.sn, q.cs);
    float invS1sq = 1.0f / dmax(q.sigma1 * q.sigma1, 1e-12f);
    float invS2sq = 1.0f / dmax(q.sigma2 * q.sigma2, 1e-12f);
    float cc = q.cs * q.cs, ss = q.sn * q.sn, cs = q.cs * q.sn;
    q.ca = cc * invS1sq + ss * invS2sq;
    q.cc = ss * invS1sq + cc * invS2sq;
    q.cb = cs * (invS1sq - invS2sq);
    float tau = dmax(kAlphaThreshold, 1e-12f);
    q.d2Cutoff = (opacity_q < tau) ? -1.0f : -2.0f * dlog(tau / opacity_q);
    return q;
}

// conicFromThetaSigmas (GaussianShared.h:490-510) on the same quantised values -- what the blend uses.
__device__ __forceinline__ void conicFromThetaSigmasQ(const QuantSplat& q, float& A, float& B, float& C) {
    float sig1 = dmax(q.sigma1, 1e-4f);
    float sig2 = dmax(q.sigma2, 1e-4f);
    float invVar1 = 1.0f / (sig1 * sig1);
    float invVar2 = 1.0f / (sig2 * sig2);
    float cc = q.cs * q.cs, ss = q.sn * q.sn, cs = q.cs * q.sn;
    A = cc * invVar1 + ss * invVar2;
    B = cs * (invVar1 - invVar2);
    C = ss * invVar1 + cc * invVar2;
}

// GaussianShared.h:518-520
__device__ __forceinline__ float evalQuad(float x, float y, float a, float b, float c) {
    return (a * x * x + 2.0f * b * x * y) + c * y * y;
}

// GaussianShared.h:525-564
__device__ __forceinline__ float minQuadRect(float xmin, float xmax, float ymin, float ymax, float a, float b, float c) {
    if (xmin <= 0.0f && 0.0f <= xmax && ymin <= 0.0f && 0.0f <= ymax) return 0.0f;
    float invA = 1.0f / dmax(a, 1e-20f);
    float invC = 1.0f / dmax(c, 1e-20f);
    float qmin = __uint_as_float(0x7F800000u);
    {
        float x = xmin;
        float y = dclamp(-(b * invC) * x, ymin, ymax);
        qmin = dmin(qmin, evalQuad(x, y, a, b, c));
    }
    {
        float x = xmax;
        float y = dclamp(-(b * invC) * x, ymin, ymax);
        qmin = dmin(qmin, evalQuad(x, y, a, b, c));
    }
    {
        float y = ymin;
        float x = dclamp(-(b * invA) * y, xmin, xmax);
        qmin = dmin(qmin, evalQuad(x, y, a, b, c));
    }
    {
        float y = ymax;
        float x = dclamp(-(b * invA) * y, xmin, xmax);
        qmin = dmin(qmin, evalQuad(x, y, a, b, c));
    }
    return qmin;
}

__device__ __forceinline__ bool tileHit(const QuantSplat& q, int tx, int ty) {
    const float tileW = (float)kTile, tileH = (float)kTile;
    float tileMinY = (float)ty * tileH;
    float tileMaxY = tileMinY + tileH;
    float tile_ymin = tileMinY - q.meanY;
    float tile_ymax = tileMaxY - q.meanY;
    float tileMinX = (float)tx * tileW;
    float tileMaxX = tileMinX + tileW;
    float tile_xmin = tileMinX - q.meanX;
    float tile_xmax = tileMaxX - q.meanX;
    float d2min = minQuadRect(tile_xmin, tile_xmax, tile_ymin, tile_ymax, q.ca, q.cb, q.cc);
    return d2min <= q.d2Cutoff;
}

// half(u8) / 255.0h as one correctly rounded half division (DFS.metal:9-11)
__device__ __forceinline__ __half u8_over_255h(uint8_t v) { return __float2half_rn((float)v / 255.0f); }

// The 32-byte record the blend stage reads: conic and colours rounded to half exactly as
// depthFirstRender does per thread (DFS.metal:1753-1764), computed once per visible Gaussian.
__device__ __forceinline__ void storeBlendSplat(BlendSplat* dst, const QuantSplat& q, __half meanX, __half meanY, uint8_t cR,
                                                uint8_t cG, uint8_t cB, uint8_t cO, __half depth) {
    float A, B, C;
    conicFromThetaSigmasQ(q, A, B, C);
    BlendSplat bs;
    bs.mean = __halves2half2(meanX, meanY);
    bs.cxx_cyy = __floats2half2_rn(A, C);
    bs.cxy2_op = __halves2half2(__float2half_rn(2.0f * B), u8_over_255h(cO));
    bs.rg = __halves2half2(u8_over_255h(cR), u8_over_255h(cG));
    bs.b_depth = __halves2half2(u8_over_255h(cB), depth);
    bs.valid = 1u; bs._pad[0] = 0; bs._pad[1] = 0;
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = *reinterpret_cast<uint4*>(&bs);
    d[1] = *(reinterpret_cast<uint4*>(&bs) + 1);
}

}  // namespace gsm
