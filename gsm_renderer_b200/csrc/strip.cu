// strip.cu -- the two extra kernels of the strip-sharded single frame (no reference
// counterpart): pack the compacted splats of a Gaussian shard into 48-byte records for the all-gather, and
// ingest the gathered records on a rank that owns tile rows [rowFirst, rowFirst+rowCount).
//
// Record = {GaussianRenderData 16 B, tile bounds int4 16 B, depth key u32, nTouched u32, gid u32, pad}.
// Records stay in ascending global gid order (rank-major shards), so the stable depth sort breaks ties
// exactly as the single-GPU frame does and per-tile lists are bit-identical.
#include "gsm_common.cuh"
#include "gsm_compact.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

struct __align__(16) SplatRecord {
    uint4 renderData;
    int4 bounds;
    uint32_t key, nTouched, gid, _pad;
};
static_assert(sizeof(SplatRecord) == GSM_SPLAT_RECORD_BYTES, "record size");

__global__ void __launch_bounds__(256) pack_records_kernel(const FrameState* __restrict__ fs, const uint32_t* __restrict__ keys,
                                                           const int32_t* __restrict__ gids, const void* __restrict__ renderData,
                                                           const int32_t* __restrict__ bounds, const uint32_t* __restrict__ nTouched,
                                                           SplatRecord* __restrict__ out, uint32_t cap) {
    const uint32_t count = min(fs->visibleCountRaw, cap);
    for (uint32_t j = blockIdx.x * 256u + threadIdx.x; j < count; j += gridDim.x * 256u) {
        const uint32_t gid = (uint32_t)gids[j];
        SplatRecord r;
        r.renderData = __ldg(reinterpret_cast<const uint4*>(renderData) + gid);
        r.bounds = __ldg(reinterpret_cast<const int4*>(bounds) + gid);
        r.key = keys[j];
        r.nTouched = nTouched[gid];
        r.gid = gid;
        r._pad = 0;
        uint4* d = reinterpret_cast<uint4*>(out + j);
        const uint4* s = reinterpret_cast<const uint4*>(&r);
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
    }
}

__global__ void __launch_bounds__(256) ingest_records_kernel(const SplatRecord* __restrict__ records, uint32_t recordCount,
                                                             int rowFirst, int rowLast, ProjectOut o) {
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&o.fs->ticketProject, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t numWarpTiles = (recordCount + 31u) / 32u;
    const uint32_t warpTile = tile * 8u + (threadIdx.x >> 5);
    if (warpTile >= numWarpTiles) return;
    const uint32_t j = tile * 256u + threadIdx.x;
    const bool inRange = j < recordCount;
    __shared__ WarpTileWork s_work[8];
    uint32_t touched = 0, key = 0xFFFFFFFFu, gid = 0, nTiles = 0, keyIn = 0;
    uint4 rd = make_uint4(0, 0, 0, 0);
    int minTX = 0, maxTX = -1, minTY = 0, maxTY = -1;
    QuantSplat q = {};
    if (inRange) {
        const uint4* s = reinterpret_cast<const uint4*>(records + j);
        rd = __ldg(s);
        const uint4 bw = __ldg(s + 1);
        const uint4 kw = __ldg(s + 2);
        gid = kw.z;
        keyIn = kw.x;
        minTX = (int)bw.x; maxTX = (int)bw.y; minTY = max((int)bw.z, rowFirst); maxTY = min((int)bw.w, rowLast);
        q = makeQuantSplat(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)), __ushort_as_half((unsigned short)(rd.x >> 16)),
                           (uint16_t)(rd.y & 0xFFFFu), __ushort_as_half((unsigned short)(rd.y >> 16)),
                           __ushort_as_half((unsigned short)(rd.z & 0xFFFFu)), (uint8_t)(rd.w >> 24));
        if (q.d2Cutoff >= 0.0f && minTX <= maxTX && minTY <= maxTY) nTiles = (uint32_t)((maxTX - minTX + 1) * (maxTY - minTY + 1));
    }
    uint2 hitMask;
    const uint32_t cnt = warpCountTiles(s_work[threadIdx.x >> 5], nTiles, q, minTX, minTY, maxTX - minTX + 1, hitMask);
    if (inRange && cnt > 0) {
        reinterpret_cast<uint4*>(o.renderData)[gid] = rd;
        reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(minTX, maxTX, minTY, maxTY);
        o.nTouched[gid] = cnt;
        o.hitMask[gid] = hitMask;
        storeBlendSplat(o.blendSplats + gid, q, __ushort_as_half((unsigned short)(rd.x & 0xFFFFu)),
                        __ushort_as_half((unsigned short)(rd.x >> 16)), (uint8_t)rd.w, (uint8_t)(rd.w >> 8), (uint8_t)(rd.w >> 16),
                        (uint8_t)(rd.w >> 24), __ushort_as_half((unsigned short)(rd.z >> 16)));
        touched = cnt;
        key = keyIn;
    }
    compactAndCount(inRange, gid, touched, key, warpTile, numWarpTiles, o);
}

cudaError_t launchPackRecords(cudaStream_t s, const FrameState* fs, const uint32_t* keys, const int32_t* gids, const void* renderData,
                              const int32_t* bounds, const uint32_t* nTouched, void* out, uint32_t cap, int numSMs) {
    pack_records_kernel<<<numSMs * 4, 256, 0, s>>>(fs, keys, gids, renderData, bounds, nTouched, (SplatRecord*)out, cap);
    return cudaGetLastError();
}

cudaError_t launchIngestRecords(cudaStream_t s, const void* records, uint32_t recordCount, uint32_t rowFirst, uint32_t rowCount,
                                const ProjectOut& o) {
    if (recordCount == 0) return cudaSuccess;
    ingest_records_kernel<<<(recordCount + 255u) / 256u, 256, 0, s>>>((const SplatRecord*)records, recordCount, (int)rowFirst,
                                                                      (int)(rowFirst + rowCount) - 1, o);
    return cudaGetLastError();
}

}  // namespace gsm
