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

__device__ __forceinline__ bool tileHitP(float meanX, float meanY, float ca, float cb, float cc, float d2Cutoff, int tx, int ty) {
    const float tileW = (float)kTile, tileH = (float)kTile;
    float tileMinY = (float)ty * tileH;
    float tileMaxY = tileMinY + tileH;
    float tile_ymin = tileMinY - meanY;
    float tile_ymax = tileMaxY - meanY;
    float tileMinX = (float)tx * tileW;
    float tileMaxX = tileMinX + tileW;
    float tile_xmin = tileMinX - meanX;
    float tile_xmax = tileMaxX - meanX;
    float d2min = minQuadRect(tile_xmin, tile_xmax, tile_ymin, tile_ymax, ca, cb, cc);
    return d2min <= d2Cutoff;
}
__device__ __forceinline__ bool tileHit(const QuantSplat& q, int tx, int ty) {
    return tileHitP(q.meanX, q.meanY, q.ca, q.cb, q.cc, q.d2Cutoff, tx, ty);
}

// ---- warp-flattened tile walks -------------------------------------------------------------------
// A thread-per-Gaussian loop over the AABB leaves ~3 of 32 lanes busy (log-normal footprints: the warp runs
// to its largest splat; ncu r1_v1: avg 3 active threads in the loop, 66% barrier stalls). Instead the 32
// lanes of a warp publish their splat parameters to shared memory and walk the CONCATENATION of their tile
// lists 32 tiles at a time; tile j belongs to the lane whose exclusive prefix covers j (5-step search).
// The test itself is unchanged (same function, same operands), so counts and emitted tiles are identical.
struct WarpTileWork {
    float meanX[32], meanY[32], ca[32], cb[32], cc[32], cutoff[32];
    int minTX[32], minTY[32], w[32];
    float rcpW[32];                   // 1 / w, for rowOf()
    uint32_t prefix[32];
    uint32_t counter[32];
    uint32_t maskLo[32], maskHi[32];  // hit bits of the first 64 AABB tiles (row-major), see warpCountTiles
};

__device__ __forceinline__ uint32_t warpExclusiveScan(uint32_t v, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    return inc - v;
}

// k / w for k < 2^23 without the ~20-instruction integer division: the binary32 estimate k * (1 / w) is within k * 2^-23 < 1 / w
// of the quotient, so its floor is exact unless the quotient is an integer the estimate falls just short of -- one fix-up.
__device__ __forceinline__ uint32_t rowOf(uint32_t k, uint32_t w, float rcpW) {
    uint32_t q = (uint32_t)__float2uint_rz(__fmul_rn(__uint2float_rn(k), rcpW));
    if (k - q * w >= w) q++;
    return q;
}

__device__ __forceinline__ void warpPublish(WarpTileWork& s, uint32_t excl, const QuantSplat& q, int minTX, int minTY, int w) {
    const unsigned lane = threadIdx.x & 31u;
    s.rcpW[lane] = 1.0f / (float)(w > 0 ? w : 1);
    s.meanX[lane] = q.meanX; s.meanY[lane] = q.meanY;
    s.ca[lane] = q.ca; s.cb[lane] = q.cb; s.cc[lane] = q.cc; s.cutoff[lane] = q.d2Cutoff;
    s.minTX[lane] = minTX; s.minTY[lane] = minTY; s.w[lane] = w;
    s.prefix[lane] = excl;
    s.counter[lane] = 0u;
    s.maskLo[lane] = 0u;
    s.maskHi[lane] = 0u;
    __syncwarp();
}

__device__ __forceinline__ uint32_t warpOwnerOf(const WarpTileWork& s, uint32_t j) {
    uint32_t lo = 0;  // largest lane with prefix <= j (lanes with no tiles share the prefix of their successor)
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
        if (s.prefix[lo + step] <= j) lo += step;
    return lo;
}

// Every lane passes n = number of AABB tiles to test (0 if none). Returns the lane's own hit count
// (DFS.metal:181-205) and, in `mask`, the hit bits of its first 64 AABB tiles in row-major order: for a splat whose
// AABB has at most kMaskTiles tiles this IS its instance list, and the expansion stage replays it instead of running
// the ellipse test a second time (the reference tests twice, DFS.metal:181-205 and :692-715; same function, same
// operands, same result). Must be called by all 32 lanes.
constexpr uint32_t kMaskTiles = 64;
__device__ __forceinline__ uint32_t warpCountTiles(WarpTileWork& s, uint32_t n, const QuantSplat& q, int minTX, int minTY, int w,
                                                   uint2& mask) {
    const unsigned lane = threadIdx.x & 31u;
    mask = make_uint2(0u, 0u);
    uint32_t total;
    const uint32_t excl = warpExclusiveScan(n, total);
    if (total == 0) return 0u;
    warpPublish(s, excl, q, minTX, minTY, w);
    for (uint32_t j0 = 0; j0 < total; j0 += 32u) {
        const uint32_t j = j0 + lane;
        if (j < total) {
            const uint32_t o = warpOwnerOf(s, j);
            const uint32_t k = j - s.prefix[o];
            const uint32_t ww = (uint32_t)s.w[o];
            const uint32_t row = rowOf(k, ww, s.rcpW[o]);
            const int ty = s.minTY[o] + (int)row, tx = s.minTX[o] + (int)(k - row * ww);
            if (tileHitP(s.meanX[o], s.meanY[o], s.ca[o], s.cb[o], s.cc[o], s.cutoff[o], tx, ty)) {
                if (k < 32u) atomicOr(&s.maskLo[o], 1u << k);
                else if (k < kMaskTiles) atomicOr(&s.maskHi[o], 1u << (k - 32u));
                else atomicAdd(&s.counter[o], 1u);
            }
        }
    }
    __syncwarp();
    mask = make_uint2(s.maskLo[lane], s.maskHi[lane]);
    const uint32_t c = s.counter[lane] + (uint32_t)__popc(mask.x) + (uint32_t)__popc(mask.y);
    __syncwarp();
    return c;
}

// ---- replay of the hit masks (instance expansion of splats with <= kMaskTiles AABB tiles) ----------
struct WarpMaskWork {
    uint32_t prefix[32];
    uint32_t maskLo[32], maskHi[32];
    int minTX[32], minTY[32];
    uint32_t w[32], recip[32];  // AABB width and ceil(2^16 / width): (b * recip) >> 16 == b / width for b < 64, width <= 64
    uint32_t base[32];
    int32_t idx[32];
};

// Every lane passes its splat's hit mask (0 if it has none or takes the tile-test path), AABB origin and width, the
// scan offset and the original index. Instance r of a splat is its r-th set bit; lanes take instances, not splats, so
// the stores are coalesced and the work is balanced. Emission order and the maxAssignments bound are those of
// msdShift of the emit functions: the tile sort runs as one pass on the id's high byte + a local pass (tilesort.cu) and wants the
// histogram of (tileId >> msdShift) & 0xFF instead of the LSD passes' digit histograms.
constexpr uint32_t kNoMsdShift = 0xFFFFFFFFu;

// DFS.metal:692-715. The tile sort's digit histograms are accumulated for the stored instances.
template <typename TileT>
__device__ __forceinline__ void warpEmitMasked(WarpMaskWork& s, uint2 mask, int minTX, int minTY, int w, uint32_t writeBase,
                                               int32_t originalIdx, uint32_t tilesX, uint32_t maxAssignments,
                                               TileT* __restrict__ tileIds, int32_t* __restrict__ instanceIdx, uint32_t* sHist,
                                               uint32_t histPasses, uint32_t msdShift) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n = (uint32_t)__popc(mask.x) + (uint32_t)__popc(mask.y);
    uint32_t total;
    const uint32_t excl = warpExclusiveScan(n, total);
    if (total == 0) return;
    s.prefix[lane] = excl;
    s.maskLo[lane] = mask.x; s.maskHi[lane] = mask.y;
    s.minTX[lane] = minTX; s.minTY[lane] = minTY;
    s.w[lane] = (uint32_t)w;
    s.recip[lane] = w > 0 ? (65536u + (uint32_t)w - 1u) / (uint32_t)w : 0u;
    s.base[lane] = writeBase;
    s.idx[lane] = originalIdx;
    __syncwarp();
    for (uint32_t j0 = 0; j0 < total; j0 += 32u) {
        const uint32_t j = j0 + lane;
        bool stored = false;
        uint32_t tileId = 0xFFFFFFFFu;
        if (j < total) {
            uint32_t o = 0;  // largest lane with prefix <= j
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (s.prefix[o + step] <= j) o += step;
            const uint32_t r = j - s.prefix[o];
            // r-th set bit of the 64-bit mask
            uint32_t word = s.maskLo[o], bit = 0, rr = r;
            const uint32_t cLo = (uint32_t)__popc(word);
            if (rr >= cLo) { rr -= cLo; word = s.maskHi[o]; bit = 32u; }
            uint32_t pos = 0;
#pragma unroll
            for (int wd = 16; wd >= 1; wd >>= 1) {
                const uint32_t c = (uint32_t)__popc((word >> pos) & ((1u << wd) - 1u));
                if (rr >= c) { rr -= c; pos += (uint32_t)wd; }
            }
            bit += pos;
            const uint32_t row = (bit * s.recip[o]) >> 16;
            const int ty = s.minTY[o] + (int)row, tx = s.minTX[o] + (int)(bit - row * s.w[o]);
            tileId = (uint32_t)(ty * (int)tilesX + tx);
            const uint32_t dst = s.base[o] + r;
            if (dst < maxAssignments) {  // DFS.metal:707
                tileIds[dst] = (TileT)tileId;
                instanceIdx[dst] = s.idx[o];
                if (msdShift == kNoMsdShift) atomicAdd(&sHist[tileId & 0xFFu], 1u);
                stored = true;
            }
        }
        if (histPasses > 1 || msdShift != kNoMsdShift) {  // upper digits are shared by most lanes: one shared-memory atomic per distinct value
            const bool msd = msdShift != kNoMsdShift;
            const uint32_t hi = stored ? (tileId >> (msd ? msdShift : 8u)) : 0xFFFFFFFFu;
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, hi);
            if (stored && (peers & ((1u << lane) - 1u)) == 0u) {
                if (msd) atomicAdd(&sHist[hi & 0xFFu], (uint32_t)__popc(peers));   // the bucket histogram of the MSD tile sort
                else for (uint32_t p = 1; p < histPasses; ++p) atomicAdd(&sHist[p * 256u + ((tileId >> (8u * p)) & 0xFFu)], (uint32_t)__popc(peers));
            }
        }
    }
    __syncwarp();
}

// Stereo expansion (createInstancesStereoKernel, DFS.metal:790-864): EVERY tile of the union AABB is an instance, so
// instance r of a splat is simply the r-th tile of its box in row-major order -- no test, no ranking. Lanes take
// instances (coalesced stores, balanced work); same bound and histogram accumulation as warpEmitMasked.
template <typename TileT>
__device__ __forceinline__ void warpEmitBox(WarpMaskWork& s, uint32_t n, int minTX, int minTY, int w, uint32_t writeBase,
                                            int32_t originalIdx, uint32_t tilesX, uint32_t maxAssignments,
                                            TileT* __restrict__ tileIds, int32_t* __restrict__ instanceIdx, uint32_t* sHist,
                                            uint32_t histPasses, uint32_t msdShift) {
    const unsigned lane = threadIdx.x & 31u;
    uint32_t total;
    const uint32_t excl = warpExclusiveScan(n, total);
    if (total == 0) return;
    s.prefix[lane] = excl;
    s.minTX[lane] = minTX; s.minTY[lane] = minTY;
    s.w[lane] = (uint32_t)(w > 0 ? w : 1);
    s.base[lane] = writeBase;
    s.idx[lane] = originalIdx;
    __syncwarp();
    for (uint32_t j0 = 0; j0 < total; j0 += 32u) {
        const uint32_t j = j0 + lane;
        bool stored = false;
        uint32_t tileId = 0xFFFFFFFFu;
        if (j < total) {
            uint32_t o = 0;  // largest lane with prefix <= j
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (s.prefix[o + step] <= j) o += step;
            const uint32_t r = j - s.prefix[o];
            const uint32_t ww = s.w[o], row = r / ww;
            const int ty = s.minTY[o] + (int)row, tx = s.minTX[o] + (int)(r - row * ww);
            tileId = (uint32_t)(ty * (int)tilesX + tx);
            const uint32_t dst = s.base[o] + r;
            if (dst < maxAssignments) {  // DFS.metal:832
                tileIds[dst] = (TileT)tileId;
                instanceIdx[dst] = s.idx[o];
                if (msdShift == kNoMsdShift) atomicAdd(&sHist[tileId & 0xFFu], 1u);
                stored = true;
            }
        }
        if (histPasses > 1 || msdShift != kNoMsdShift) {  // upper digits are shared by most lanes: one shared-memory atomic per distinct value
            const bool msd = msdShift != kNoMsdShift;
            const uint32_t hi = stored ? (tileId >> (msd ? msdShift : 8u)) : 0xFFFFFFFFu;
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, hi);
            if (stored && (peers & ((1u << lane) - 1u)) == 0u) {
                if (msd) atomicAdd(&sHist[hi & 0xFFu], (uint32_t)__popc(peers));   // the bucket histogram of the MSD tile sort
                else for (uint32_t p = 1; p < histPasses; ++p) atomicAdd(&sHist[p * 256u + ((tileId >> (8u * p)) & 0xFFu)], (uint32_t)__popc(peers));
            }
        }
    }
    __syncwarp();
}

// Emits the hit tiles of every lane's splat in row-major order at offsets[lane] + rank (DFS.metal:692-715).
// Items of one owner sit in consecutive lanes, so a match-any group + ballot ranks them in order.
template <typename TileT>
__device__ __forceinline__ void warpEmitTiles(WarpTileWork& s, uint32_t n, const QuantSplat& q, int minTX, int minTY, int w,
                                              uint32_t writeBase, int32_t originalIdx, uint32_t tilesX, uint32_t maxAssignments,
                                              TileT* __restrict__ tileIds, int32_t* __restrict__ instanceIdx, uint32_t* sBase,
                                              int32_t* sIdx, uint32_t* sHist, uint32_t histPasses, uint32_t msdShift) {
    const unsigned lane = threadIdx.x & 31u;
    uint32_t total;
    const uint32_t excl = warpExclusiveScan(n, total);
    if (total == 0) return;
    sBase[lane] = writeBase;
    sIdx[lane] = originalIdx;
    warpPublish(s, excl, q, minTX, minTY, w);
    for (uint32_t j0 = 0; j0 < total; j0 += 32u) {
        const uint32_t j = j0 + lane;
        const bool active = j < total;
        uint32_t o = 0xFFFFFFFFu;
        bool hit = false;
        uint32_t tileId = 0;
        if (active) {
            o = warpOwnerOf(s, j);
            const uint32_t k = j - s.prefix[o];
            const uint32_t ww = (uint32_t)s.w[o];
            const uint32_t row = rowOf(k, ww, s.rcpW[o]);
            const int ty = s.minTY[o] + (int)row, tx = s.minTX[o] + (int)(k - row * ww);
            hit = tileHitP(s.meanX[o], s.meanY[o], s.ca[o], s.cb[o], s.cc[o], s.cutoff[o], tx, ty);
            tileId = (uint32_t)(ty * (int)tilesX + tx);
        }
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, o);
        const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit) & peers;
        uint32_t before = 0;
        if (active) before = s.counter[o];
        __syncwarp();
        if (active && (peers & ((1u << lane) - 1u)) == 0u) s.counter[o] = before + __popc(hits);
        __syncwarp();
        if (histPasses > 1 || msdShift != kNoMsdShift) {
            // higher digits are equal along a tile row, so one shared-memory atomic per lane would serialise; lanes
            // that hit and share the upper bits with their left neighbour form a run and only its head adds the run length
            const bool counted = hit && (sBase[o] + before + __popc(hits & ((1u << lane) - 1u)) < maxAssignments);
            const bool msd = msdShift != kNoMsdShift;
            const uint32_t hi = tileId >> (msd ? msdShift : 8u);
            const uint32_t hiPrev = __shfl_up_sync(0xFFFFFFFFu, hi, 1);
            const unsigned cmask = __ballot_sync(0xFFFFFFFFu, counted);
            const bool head = counted && (lane == 0 || !((cmask >> (lane - 1)) & 1u) || hiPrev != hi);
            const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
            if (head) {
                const unsigned after = heads & ~((2u << lane) - 1u);           // heads to my right
                const unsigned breaks = (~cmask) & ~((2u << lane) - 1u);       // first uncounted lane to my right ends the run too
                const unsigned stop = after | breaks;
                const uint32_t endLane = stop ? (uint32_t)(__ffs(stop) - 1) : 32u;
                const uint32_t runLen = endLane - lane;
                if (msd) atomicAdd(&sHist[hi & 0xFFu], runLen);
                else for (uint32_t p = 1; p < histPasses; ++p) atomicAdd(&sHist[p * 256u + ((tileId >> (8u * p)) & 0xFFu)], runLen);
            }
        }
        if (hit) {
            const uint32_t pos = sBase[o] + before + __popc(hits & ((1u << lane) - 1u));
            if (pos < maxAssignments) {  // DFS.metal:707
                tileIds[pos] = (TileT)tileId;
                instanceIdx[pos] = sIdx[o];
                if (msdShift == kNoMsdShift) atomicAdd(&sHist[tileId & 0xFFu], 1u);  // low digit: neighbouring lanes hold consecutive tiles, no conflict
            }
        }
    }
    __syncwarp();
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
