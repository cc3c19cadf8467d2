// expand.cu -- stages 4+5 in one kernel (stage 3, ordered[i] = nTouched[sortedIdx[i]] -- applyDepthOrderingKernel,
// DFS.metal:623-640 -- is written by the depth sort's last pass): the exclusive scan of ordered[], in place (5-kernel prefix sum, DFS.metal:2036-2139, driver
// InstanceExpansionEncoder.swift:83-176) by a single-pass decoupled look-back over 256-Gaussian tiles, and the
// instance expansion itself: re-run the exact tile test on the quantised record and emit (tileId, originalIdx) at
// the scan offset, row-major ty -> tx, bounded by maxAssignments (createInstancesKernel / ...32 DFS.metal:642-788,
// createInstancesStereoKernel / ...32 DFS.metal:790-864). A tile publishes its aggregate before it starts walking,
// so nobody waits on a neighbour's tile walk. The tile sort's digit histograms are accumulated on the way out.
#include "gsm_common.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

// tools/expand_trace.py builds the library with GSM_EXPAND_TRACE to record a per-tile timeline (globaltimer ns,
// written by thread 0: [0] CTA entry, [1] ticket, [2] counts gathered + scanned, [3] prefix known, [4] bounds/masks
// loaded, [5] warp 0 done, [6] all warps done, [7] SM id); the product build compiles the hooks out.
#ifdef GSM_EXPAND_TRACE
__device__ unsigned long long* g_expandTrace = nullptr;
__device__ __forceinline__ unsigned long long expandNow() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GSM_XTRACE(tile, slot) do { if (threadIdx.x == 0 && g_expandTrace) g_expandTrace[(size_t)(tile) * 8 + (slot)] = expandNow(); } while (0)
#define GSM_XTRACE_SET(tile, slot, v) do { if (threadIdx.x == 0 && g_expandTrace) g_expandTrace[(size_t)(tile) * 8 + (slot)] = (v); } while (0)
#else
#define GSM_XTRACE(tile, slot) do { } while (0)
#define GSM_XTRACE_SET(tile, slot, v) do { } while (0)
#endif

template <typename TileT, bool STEREO>
__global__ void __launch_bounds__(256) create_instances_kernel(const int32_t* __restrict__ sortedIdx,
                                                               const uint32_t* sortedTouched,  // may alias offsets (scanned in place)
                                                               const uint2* __restrict__ hitMask,
                                                               uint32_t* offsets, unsigned long long* scanStatus,
                                                               unsigned long long* scanGroups,
                                                               uint32_t* ticket, const int32_t* __restrict__ bounds,
                                                               const void* __restrict__ renderData,
                                                               TileT* __restrict__ tileIds, int32_t* __restrict__ instanceIdx,
                                                               const GSMDepthFirstHeader* __restrict__ header, uint32_t tilesX,
                                                               uint32_t maxAssignments, uint32_t* __restrict__ tileHist,
                                                               uint32_t tilePasses, uint32_t msdShift) {
    // the mask replay and the tile-test walk run one after the other: one buffer serves both
    union WarpWork { WarpTileWork test; WarpMaskWork replay; };
    __shared__ WarpWork s_warpWork[8];
#ifdef GSM_EXPAND_TRACE
    const unsigned long long traceEntry = expandNow();
#endif
    __shared__ uint32_t s_hist[4][256];  // digit histograms of the emitted tile ids (the tile sort's histogram pass, fused)
    pdlLaunchDependents();
    for (int i = threadIdx.x; i < 4 * 256; i += 256) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    pdlWait();
    __shared__ uint32_t s_base[8][32];
    __shared__ int32_t s_idx[8][32];
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_tile, s_tileBase;
    const uint32_t visibleCount = ldAfterWait(&header->visibleCount);
    const unsigned warp = threadIdx.x >> 5;
    const uint32_t numTiles = (visibleCount + 255u) / 256u;

    // Front half of a tile: counts in depth order, block scan, PUBLISH the tile aggregate. Nothing here waits on another
    // tile, so an aggregate is out about two round trips after its ticket.
    auto front = [&](uint32_t tile, int32_t& idx, uint32_t& excl) {
        const uint32_t i = tile * 256u + threadIdx.x;
        idx = -1;
        if (i < visibleCount) idx = sortedIdx[i];
        // nTouched in depth order (DFS.metal:633-639), written by the depth sort's last pass
        const uint32_t cnt = (i < visibleCount && idx >= 0) ? sortedTouched[i] : 0u;
        uint32_t total;
        excl = block_exclusive_scan_256(cnt, s_scan, total);
        if (threadIdx.x == 0) prefixPublish(scanStatus, scanGroups, tile, total);
        GSM_XTRACE(tile, 2);
    };

    // Software pipeline over tiles A (back half: resolve the prefix, expand) and B (front half): B's aggregate is
    // published BEFORE A is expanded, so the wait for "every predecessor has published" (expansion timeline: ~5 us per
    // tile, as long as the expansion itself) runs under A's expansion instead of in front of it. Whole warps stay
    // together: the tile walk is warp-cooperative.
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    uint32_t tileB = s_tile;
    int32_t idxB = -1;
    uint32_t exclB = 0;
    if (tileB < numTiles) {
#ifdef GSM_EXPAND_TRACE
        { unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); GSM_XTRACE_SET(tileB, 0, traceEntry); GSM_XTRACE_SET(tileB, 7, (unsigned long long)smid); }
#endif
        GSM_XTRACE(tileB, 1);
        front(tileB, idxB, exclB);
    }
    while (tileB < numTiles) {
        const uint32_t tile = tileB;
        const int32_t originalIdx = idxB;
        const uint32_t excl = exclB;
        const uint32_t i = tile * 256u + threadIdx.x;
        __syncthreads();  // everyone has read s_tile
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        // this tile's gathers do not depend on its prefix: in flight under the ticket and B's front half
        int minTX = 0, maxTX = -1, minTY = 0, maxTY = -1;
        uint32_t n = 0;
        QuantSplat q = {};
        uint2 mask = make_uint2(0u, 0u);  // mono, AABB of at most kMaskTiles tiles: replay stage 1's hit bits
        if (i < visibleCount && originalIdx >= 0) {
            const int4 b = __ldg(reinterpret_cast<const int4*>(bounds) + originalIdx);
            if (!STEREO) mask = __ldg(hitMask + originalIdx);
            minTX = b.x; maxTX = b.y; minTY = b.z; maxTY = b.w;
            if (minTX <= maxTX && minTY <= maxTY) {
                n = (uint32_t)((maxTX - minTX + 1) * (maxTY - minTY + 1));
                if (!STEREO) {
                    if (n <= kMaskTiles) {
                        n = 0;
                    } else {
                        mask = make_uint2(0u, 0u);
                        const uint4 rd = __ldg(reinterpret_cast<const uint4*>(renderData) + originalIdx);
                        q = makeQuantSplat(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)),
                                           __ushort_as_half((unsigned short)(rd.x >> 16)), (uint16_t)(rd.y & 0xFFFFu),
                                           __ushort_as_half((unsigned short)(rd.y >> 16)),
                                           __ushort_as_half((unsigned short)(rd.z & 0xFFFFu)), (uint8_t)(rd.w >> 24));
                        if (!(q.d2Cutoff >= 0.0f)) n = 0;
                    }
                }
            } else {
                mask = make_uint2(0u, 0u);
            }
        }
        __syncthreads();
        tileB = s_tile;
        if (tileB < numTiles) {
#ifdef GSM_EXPAND_TRACE
            { unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); GSM_XTRACE_SET(tileB, 0, traceEntry); GSM_XTRACE_SET(tileB, 7, (unsigned long long)smid); }
#endif
            GSM_XTRACE(tileB, 1);
            front(tileB, idxB, exclB);
        }
        if (threadIdx.x < 32) {
            const uint32_t b = prefixResolve(scanStatus, scanGroups, tile);
            if (threadIdx.x == 0) s_tileBase = b;
        }
        __syncthreads();
        const uint32_t writeOffset = s_tileBase + excl;
        GSM_XTRACE(tile, 3);
        if (i < visibleCount) offsets[i] = writeOffset;  // the in-place scan result (debugReadInstanceOffsets)
        GSM_XTRACE(tile, 4);
        if (!STEREO)
            warpEmitMasked<TileT>(s_warpWork[warp].replay, mask, minTX, minTY, maxTX - minTX + 1, writeOffset, originalIdx, tilesX, maxAssignments,
                                  tileIds, instanceIdx, &s_hist[0][0], tilePasses, msdShift);
        if (STEREO) {
            // every tile of the union AABB is an instance, no ellipse test (DFS.metal:816-825)
            warpEmitBox<TileT>(s_warpWork[warp].replay, n, minTX, minTY, maxTX - minTX + 1, writeOffset, originalIdx, tilesX, maxAssignments,
                               tileIds, instanceIdx, &s_hist[0][0], tilePasses, msdShift);
        } else {
            warpEmitTiles<TileT>(s_warpWork[warp].test, n, q, minTX, minTY, maxTX - minTX + 1, writeOffset, originalIdx, tilesX, maxAssignments,
                                 tileIds, instanceIdx, s_base[warp], s_idx[warp], &s_hist[0][0], tilePasses, msdShift);
        }
#ifdef GSM_EXPAND_TRACE
        GSM_XTRACE(tile, 5);
        __syncthreads();
        GSM_XTRACE(tile, 6);
#endif
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (msdShift != kNoMsdShift ? 256u : tilePasses * 256u); i += 256u) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&tileHist[i], v);
    }
}

cudaError_t launchCreateInstances(cudaStream_t s, bool stereo, bool tileId16, const int32_t* sortedIdx, const uint32_t* sortedTouched,
                                  const uint2* hitMask, uint32_t* offsets, unsigned long long* scanStatus, unsigned long long* scanGroups, uint32_t* ticket, const int32_t* bounds, const void* renderData, void* tileIds, int32_t* instanceIdx,
                                  const GSMDepthFirstHeader* header, uint32_t tilesX, uint32_t maxAssignments, uint32_t capVisible,
                                  uint32_t* tileHist, uint32_t tilePasses, int numSMs, uint32_t msdShift) {
    uint32_t grid = (capVisible + 255u) / 256u;
    // persistent, 6 CTAs per SM: 3 left the kernel 25 % slower, 8 changed nothing (profiles/README.md); few CTAs flush the histograms
    if (grid > (uint32_t)numSMs * 6u) grid = (uint32_t)numSMs * 6u;
    if (grid == 0) grid = 1;
#define GSM_LAUNCH(T, ST) launchChained(create_instances_kernel<T, ST>, grid, 256, s, sortedIdx, sortedTouched, hitMask, offsets, scanStatus, scanGroups, ticket, bounds, renderData, (T*)tileIds, instanceIdx, header, tilesX, maxAssignments, tileHist, tilePasses, msdShift)
    if (tileId16) { if (stereo) GSM_LAUNCH(uint16_t, true); else GSM_LAUNCH(uint16_t, false); }
    else { if (stereo) GSM_LAUNCH(uint32_t, true); else GSM_LAUNCH(uint32_t, false); }
#undef GSM_LAUNCH
    return cudaGetLastError();
}

#ifdef GSM_EXPAND_TRACE
extern "C" int gsm_trace_expand_set(void* devBuffer) {
    unsigned long long* p = (unsigned long long*)devBuffer;
    return (int)cudaMemcpyToSymbol(g_expandTrace, &p, sizeof(p));
}
#endif

}  // namespace gsm
