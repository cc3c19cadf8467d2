// gsm_compact.cuh -- the visibility-compaction tail shared by the projection kernels and the strip ingest:
// ranks the visible threads of a 256-thread tile in gid order, resolves the tile's global base by decoupled
// look-back, scatters (depthKey, gid) and accumulates the frame counters.
// Replaces the 8-pass VisibilityCompactionEncoder (DFS.metal:518-621).
#pragma once
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

// Block-wide: ranks the visible threads in gid order, resolves the block's global base by look-back,
// scatters (key, gid) and accumulates the counters.
__device__ __forceinline__ void compactAndCount(bool inRange, uint32_t gid, uint32_t touched, uint32_t key, uint32_t tile,
                                                uint32_t numTiles, const ProjectOut& o) {
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_touchedSum[8];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t flag = (inRange && touched > 0) ? 1u : 0u;
    uint32_t blockVisible;
    uint32_t local = block_exclusive_scan_256(flag, s_scan, blockVisible);
    uint32_t t = touched;
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, off);
    if (lane == 0) s_touchedSum[warp] = t;
    __syncthreads();
    if (warp == 0) {
        uint32_t excl = lookback_exclusive(o.status, tile, blockVisible);
        if (lane == 0) {
            s_base = excl;
            uint32_t sum = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += s_touchedSum[w];
            if (sum) atomicAdd(&o.fs->totalInstancesRaw, sum);
            if (tile == numTiles - 1) o.fs->visibleCountRaw = excl + blockVisible;
        }
    }
    __syncthreads();
    if (flag) {
        uint32_t dst = s_base + local;
        if (dst < o.maxOut) {  // DFS.metal:605
            if (o.depthKey16) {  // DFS.metal:607-612; key is float_to_sortable_uint of a depth > 0
                uint32_t bits = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
                key = (uint32_t)(__half_as_ushort(__float2half_rn(__uint_as_float(bits))) ^ 0x8000u);
            }
            o.depthKeys[dst] = key;
            o.primitiveIndices[dst] = (int32_t)gid;
        }
    }
}

__device__ __forceinline__ void writeCulled(const ProjectOut& o, uint32_t gid) {
    o.nTouched[gid] = 0;
    reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(0, -1, 0, -1);
}


}  // namespace gsm
