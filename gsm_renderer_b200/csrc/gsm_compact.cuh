// gsm_compact.cuh -- the visibility-compaction tail shared by the projection kernels and the strip ingest.
// Replaces the 8-pass VisibilityCompactionEncoder (DFS.metal:518-621) and the atomic of DFS.metal:218.
//
// Granularity is ONE WARP (32 consecutive gids), not the CTA: a warp ranks its visible lanes with a ballot,
// resolves its global base by a decoupled look-back over the warp-tiles before it and scatters -- no
// __syncthreads, so a warp that drew a large splat does not hold its 7 siblings at a barrier (ncu r1_v2:
// 36 % of the project kernel's stall samples were barrier waits). The look-back word carries BOTH running
// sums -- visible count (30 bits) and touched-tile count (32 bits) -- so the last warp-tile publishes
// visibleCount and totalInstances and no global atomics are needed. Ascending gid order is kept: warp-tiles
// are numbered ticket*warpsPerCTA + warp and tickets are handed out in launch order.
#pragma once
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

// status word: [63:62] flag, [61:32] visible count, [31:0] sum of touched tiles (wraps like the reference's u32 atomic)
__device__ __forceinline__ unsigned long long packStatus2(uint32_t flag, uint32_t visible, uint32_t touched) {
    return ((unsigned long long)flag << 62) | ((unsigned long long)(visible & 0x3FFFFFFFu) << 32) | touched;
}

__device__ __forceinline__ void lookback_exclusive2(unsigned long long* status, uint32_t tile, uint32_t aggVisible,
                                                    uint32_t aggTouched, uint32_t& exclVisible, uint32_t& exclTouched) {
    const unsigned lane = threadIdx.x & 31u;
    exclVisible = 0;
    exclTouched = 0;
    if (tile == 0) {
        if (lane == 0) st_u64_relaxed(status, packStatus2(kFlagInclusive, aggVisible, aggTouched));
        return;
    }
    if (lane == 0) st_u64_relaxed(status + tile, packStatus2(kFlagAggregate, aggVisible, aggTouched));
    int look = (int)tile - 1;
    while (true) {
        const int idx = look - (int)lane;
        const unsigned long long s = (idx >= 0) ? ld_status64(&status[idx]) : packStatus2(kFlagInclusive, 0, 0);
        const uint32_t flag = (uint32_t)(s >> 62);
        if (__any_sync(0xFFFFFFFFu, flag == 0)) continue;  // a word in the window is not published yet
        const unsigned incl = __ballot_sync(0xFFFFFFFFu, flag == kFlagInclusive);
        uint32_t v = (uint32_t)(s >> 32) & 0x3FFFFFFFu, t = (uint32_t)s;
        if (incl) {
            const int first = __ffs(incl) - 1;  // nearest predecessor holding an inclusive prefix
            if (lane > (unsigned)first) { v = 0; t = 0; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
        }
        exclVisible += v;
        exclTouched += t;
        if (incl) break;
        look -= 32;
    }
    if (lane == 0) st_u64_relaxed(status + tile, packStatus2(kFlagInclusive, exclVisible + aggVisible, exclTouched + aggTouched));
}

// Must be called by all 32 lanes of the warp. warpTile / numWarpTiles index 32-gid groups in gid order.
__device__ __forceinline__ void compactAndCount(bool inRange, uint32_t gid, uint32_t touched, uint32_t key, uint32_t warpTile,
                                                uint32_t numWarpTiles, const ProjectOut& o) {
    const unsigned lane = threadIdx.x & 31u;
    const bool flag = inRange && touched > 0;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
    const uint32_t warpVisible = __popc(bal);
    uint32_t t = touched;
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, off);
    uint32_t exclVisible, exclTouched;
    lookback_exclusive2(o.status, warpTile, warpVisible, t, exclVisible, exclTouched);
    if (lane == 0 && warpTile == numWarpTiles - 1) {
        o.fs->visibleCountRaw = exclVisible + warpVisible;       // DFS.metal:618-620
        o.fs->totalInstancesRaw = exclTouched + t;               // DFS.metal:218
    }
    if (flag) {
        const uint32_t dst = exclVisible + __popc(bal & ((1u << lane) - 1u));
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
