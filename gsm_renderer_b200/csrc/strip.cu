// strip.cu -- the two extra kernels of the strip-sharded single frame (no reference
// counterpart): pack the compacted splats of a Gaussian shard into 48-byte records for the all-gather, and
// ingest the gathered records on a rank that owns tile rows [rowFirst, rowFirst+rowCount).
//
// Record = {GaussianRenderData 16 B, tile bounds int4 16 B, depth key u32, hit mask 2 x u32, gid u32}.
// Records stay in ascending global gid order (rank-major shards), so the stable depth sort breaks ties
// exactly as the single-GPU frame does and per-tile lists are bit-identical.
#include "gsm_common.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

// SplatRecord: gsm_common.cuh

__global__ void __launch_bounds__(256) pack_records_kernel(const FrameState* __restrict__ fs, const uint32_t* __restrict__ keys,
                                                           const int32_t* __restrict__ gids, const void* __restrict__ renderData,
                                                           const int32_t* __restrict__ bounds, const uint2* __restrict__ hitMask,
                                                           SplatRecord* __restrict__ out, uint32_t cap) {
    const uint32_t count = min(fs->visibleCountRaw, cap);
    for (uint32_t j = blockIdx.x * 256u + threadIdx.x; j < count; j += gridDim.x * 256u) {
        const uint32_t gid = (uint32_t)gids[j];
        SplatRecord r;
        r.renderData = __ldg(reinterpret_cast<const uint4*>(renderData) + gid);
        r.bounds = __ldg(reinterpret_cast<const int4*>(bounds) + gid);
        const uint2 m = __ldg(hitMask + gid);
        r.key = keys[j];
        r.maskLo = m.x; r.maskHi = m.y;
        r.gid = gid;
        uint4* d = reinterpret_cast<uint4*>(out + j);
        const uint4* s = reinterpret_cast<const uint4*>(&r);
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    }
}

// One record per thread. The tile count of a record clipped to this rank's rows needs no ellipse test when the
// projecting rank's hit mask covers the AABB (<= 64 tiles, i.e. almost always): rows are contiguous bit ranges of the
// row-major mask, so clipping is a shift and a mask. Larger AABBs re-run the exact walk on the clipped box. Per-gid
// outputs are written here; the order-preserving compaction, the depth histograms and the frame header come from
// compact_visible_kernel over the per-record (count, key, gid) arrays.
__global__ void __launch_bounds__(256) ingest_records_kernel(const SplatRecord* __restrict__ records, uint32_t recordCount,
                                                             int rowFirst, int rowLast, ProjectOut o, uint32_t* __restrict__ recTouched,
                                                             uint32_t* __restrict__ recKey, uint32_t* __restrict__ recGid) {
    pdlLaunchDependents();
    pdlWait();
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if ((j & ~31u) >= recordCount) return;  // whole warp past the end
    const bool inRange = j < recordCount;
    __shared__ WarpTileWork s_work[8];
    uint32_t gid = 0, nTiles = 0, keyIn = 0xFFFFFFFFu, cnt = 0;
    uint4 rd = make_uint4(0, 0, 0, 0);
    int minTX = 0, maxTX = -1, minTY = 0, maxTY = -1;
    uint2 mask = make_uint2(0u, 0u);
    QuantSplat q = {};
    if (inRange) {
        const uint4* s = reinterpret_cast<const uint4*>(records + j);
        rd = __ldg(s);
        const uint4 bw = __ldg(s + 1);
        const uint4 kw = __ldg(s + 2);
        gid = kw.w;
        keyIn = kw.x;
        minTX = (int)bw.x; maxTX = (int)bw.y; minTY = max((int)bw.z, rowFirst); maxTY = min((int)bw.w, rowLast);
        if (minTX <= maxTX && minTY <= maxTY) {
            const uint32_t w = (uint32_t)(maxTX - minTX + 1);
            const uint32_t fullTiles = w * (uint32_t)((int)bw.w - (int)bw.z + 1);
            if (fullTiles <= kMaskTiles) {
                const uint32_t shift = (uint32_t)(minTY - (int)bw.z) * w, bits = (uint32_t)(maxTY - minTY + 1) * w;
                unsigned long long m = (((unsigned long long)kw.z << 32) | kw.y) >> shift;  // shift < 64: the clipped box is not empty
                if (bits < 64u) m &= (1ull << bits) - 1ull;
                mask = make_uint2((uint32_t)m, (uint32_t)(m >> 32));
                cnt = (uint32_t)__popcll(m);
            } else {
                nTiles = w * (uint32_t)(maxTY - minTY + 1);
            }
        }
        if (cnt > 0 || nTiles > 0)
            q = makeQuantSplat(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)), __ushort_as_half((unsigned short)(rd.x >> 16)),
                               (uint16_t)(rd.y & 0xFFFFu), __ushort_as_half((unsigned short)(rd.y >> 16)),
                               __ushort_as_half((unsigned short)(rd.z & 0xFFFFu)), (uint8_t)(rd.w >> 24));
        if (nTiles > 0 && !(q.d2Cutoff >= 0.0f)) nTiles = 0;
    }
    uint2 walkMask;
    const uint32_t walkCnt = warpCountTiles(s_work[threadIdx.x >> 5], nTiles, q, minTX, minTY, maxTX - minTX + 1, walkMask);
    if (nTiles > 0) { cnt = walkCnt; mask = walkMask; }
    if (inRange) {
        if (cnt > 0) {
            reinterpret_cast<uint4*>(o.renderData)[gid] = rd;
            reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(minTX, maxTX, minTY, maxTY);
            o.nTouched[gid] = cnt;
            o.hitMask[gid] = mask;
            storeBlendSplat(o.blendSplats + gid, q, __ushort_as_half((unsigned short)(rd.x & 0xFFFFu)),
                            __ushort_as_half((unsigned short)(rd.x >> 16)), (uint8_t)rd.w, (uint8_t)(rd.w >> 8), (uint8_t)(rd.w >> 16),
                            (uint8_t)(rd.w >> 24), __ushort_as_half((unsigned short)(rd.z >> 16)));
        }
        recTouched[j] = cnt;
        recKey[j] = keyIn;
        recGid[j] = gid;
    }
}

cudaError_t launchPackRecords(cudaStream_t s, const FrameState* fs, const uint32_t* keys, const int32_t* gids, const void* renderData,
                              const int32_t* bounds, const uint2* hitMask, void* out, uint32_t cap, int numSMs) {
    pack_records_kernel<<<numSMs * 4, 256, 0, s>>>(fs, keys, gids, renderData, bounds, hitMask, (SplatRecord*)out, cap);
    return cudaGetLastError();
}

cudaError_t launchIngestRecords(cudaStream_t s, const void* records, uint32_t recordCount, uint32_t rowFirst, uint32_t rowCount,
                                const ProjectOut& o, uint32_t* recTouched, uint32_t* recKey, uint32_t* recGid) {
    if (recordCount == 0) return cudaSuccess;
    return launchChained(ingest_records_kernel, (recordCount + 255u) / 256u, 256, s, (const SplatRecord*)records, recordCount, (int)rowFirst,
                         (int)(rowFirst + rowCount) - 1, o, recTouched, recKey, recGid);
}

}  // namespace gsm
