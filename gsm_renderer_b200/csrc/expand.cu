// expand.cu -- stages 3+4+5 in one kernel: ordered[i] = nTouched[sortedIdx[i]] (applyDepthOrderingKernel,
// DFS.metal:623-640), its exclusive scan (5-kernel prefix sum, DFS.metal:2036-2139, driver
// InstanceExpansionEncoder.swift:83-176) by a single-pass decoupled look-back over 256-Gaussian tiles, and the
// instance expansion itself: re-run the exact tile test on the quantised record and emit (tileId, originalIdx) at
// the scan offset, row-major ty -> tx, bounded by maxAssignments (createInstancesKernel / ...32 DFS.metal:642-788,
// createInstancesStereoKernel / ...32 DFS.metal:790-864). A tile publishes its aggregate before it starts walking,
// so nobody waits on a neighbour's tile walk. The tile sort's digit histograms are accumulated on the way out.
#include "gsm_common.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

template <typename TileT, bool STEREO>
__global__ void __launch_bounds__(256) create_instances_kernel(const int32_t* __restrict__ sortedIdx,
                                                               const uint32_t* __restrict__ nTouched,
                                                               const uint2* __restrict__ hitMask,
                                                               uint32_t* __restrict__ offsets, unsigned long long* scanStatus,
                                                               uint32_t* ticket, const int32_t* __restrict__ bounds,
                                                               const void* __restrict__ renderData,
                                                               TileT* __restrict__ tileIds, int32_t* __restrict__ instanceIdx,
                                                               const GSMDepthFirstHeader* __restrict__ header, uint32_t tilesX,
                                                               uint32_t maxAssignments, uint32_t* __restrict__ tileHist,
                                                               uint32_t tilePasses) {
    __shared__ WarpTileWork s_work[8];
    __shared__ WarpMaskWork s_mask[8];
    __shared__ uint32_t s_hist[4][256];  // digit histograms of the emitted tile ids (the tile sort's histogram pass, fused)
    pdlLaunchDependents();
    for (int i = threadIdx.x; i < 4 * 256; i += 256) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    pdlWait();
    __shared__ uint32_t s_base[8][32];
    __shared__ int32_t s_idx[8][32];
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_tile, s_tileBase;
    const uint32_t visibleCount = header->visibleCount;
    const unsigned warp = threadIdx.x >> 5;
    const uint32_t numTiles = (visibleCount + 255u) / 256u;
    // whole warps stay together: the tile walk is warp-cooperative
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t i = tile * 256u + threadIdx.x;
        int32_t originalIdx = -1;
        int minTX = 0, maxTX = -1, minTY = 0, maxTY = -1;
        uint32_t writeOffset = 0, n = 0;
        QuantSplat q = {};
        if (i < visibleCount) originalIdx = sortedIdx[i];
        {   // stages 3+4: gather the tile count in depth order, scan, publish the tile aggregate, look back
            const uint32_t cnt = (originalIdx >= 0) ? __ldg(nTouched + originalIdx) : 0u;  // DFS.metal:633-639
            uint32_t total;
            const uint32_t excl = block_exclusive_scan_256(cnt, s_scan, total);
            if (threadIdx.x < 32) {
                const uint32_t b = lookback_exclusive(scanStatus, tile, total);
                if (threadIdx.x == 0) s_tileBase = b;
            }
            __syncthreads();
            writeOffset = s_tileBase + excl;
            if (i < visibleCount) offsets[i] = writeOffset;  // the in-place scan result (debugReadInstanceOffsets)
        }
        uint2 mask = make_uint2(0u, 0u);  // mono, AABB of at most kMaskTiles tiles: replay stage 1's hit bits
        if (i < visibleCount) {
            if (originalIdx >= 0) {
                const int4 b = __ldg(reinterpret_cast<const int4*>(bounds) + originalIdx);
                minTX = b.x; maxTX = b.y; minTY = b.z; maxTY = b.w;
                if (minTX <= maxTX && minTY <= maxTY) {
                    n = (uint32_t)((maxTX - minTX + 1) * (maxTY - minTY + 1));
                    if (!STEREO) {
                        if (n <= kMaskTiles) {
                            mask = __ldg(hitMask + originalIdx);
                            n = 0;
                        } else {
                            const uint4 rd = __ldg(reinterpret_cast<const uint4*>(renderData) + originalIdx);
                            q = makeQuantSplat(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)),
                                               __ushort_as_half((unsigned short)(rd.x >> 16)), (uint16_t)(rd.y & 0xFFFFu),
                                               __ushort_as_half((unsigned short)(rd.y >> 16)),
                                               __ushort_as_half((unsigned short)(rd.z & 0xFFFFu)), (uint8_t)(rd.w >> 24));
                            if (!(q.d2Cutoff >= 0.0f)) n = 0;
                        }
                    }
                }
            }
        }
        if (!STEREO)
            warpEmitMasked<TileT>(s_mask[warp], mask, minTX, minTY, maxTX - minTX + 1, writeOffset, originalIdx, tilesX, maxAssignments,
                                  tileIds, instanceIdx, &s_hist[0][0], tilePasses);
        if (STEREO) {
            // every tile of the union AABB, no ellipse test (DFS.metal:816-825): a splat whose mean is inside every
            // tile makes tileHitP return true for all of them without changing the walk
            q.meanX = 0.0f; q.meanY = 0.0f; q.ca = 0.0f; q.cb = 0.0f; q.cc = 0.0f;
            q.d2Cutoff = __uint_as_float(0x7F800000u);  // d2min <= +inf always (d2min is never NaN for a = b = c = 0)
        }
        warpEmitTiles<TileT>(s_work[warp], n, q, minTX, minTY, maxTX - minTX + 1, writeOffset, originalIdx, tilesX, maxAssignments,
                             tileIds, instanceIdx, s_base[warp], s_idx[warp], &s_hist[0][0], tilePasses);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tilePasses * 256u; i += 256u) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&tileHist[i], v);
    }
}

cudaError_t launchCreateInstances(cudaStream_t s, bool stereo, bool tileId16, const int32_t* sortedIdx, const uint32_t* nTouched,
                                  const uint2* hitMask, uint32_t* offsets, unsigned long long* scanStatus, uint32_t* ticket, const int32_t* bounds, const void* renderData, void* tileIds, int32_t* instanceIdx,
                                  const GSMDepthFirstHeader* header, uint32_t tilesX, uint32_t maxAssignments, uint32_t capVisible,
                                  uint32_t* tileHist, uint32_t tilePasses, int numSMs) {
    uint32_t grid = (capVisible + 255u) / 256u;
    if (grid > (uint32_t)numSMs * 6u) grid = (uint32_t)numSMs * 6u;  // persistent: few CTAs flush the fused histograms
    if (grid == 0) grid = 1;
#define GSM_LAUNCH(T, ST) launchChained(create_instances_kernel<T, ST>, grid, 256, s, sortedIdx, nTouched, hitMask, offsets, scanStatus, ticket, bounds, renderData, (T*)tileIds, instanceIdx, header, tilesX, maxAssignments, tileHist, tilePasses)
    if (tileId16) { if (stereo) GSM_LAUNCH(uint16_t, true); else GSM_LAUNCH(uint16_t, false); }
    else { if (stereo) GSM_LAUNCH(uint32_t, true); else GSM_LAUNCH(uint32_t, false); }
#undef GSM_LAUNCH
    return cudaGetLastError();
}

}  // namespace gsm
