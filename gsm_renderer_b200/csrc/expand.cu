// expand.cu -- stage 5: per depth-sorted Gaussian, re-run the exact tile test on the quantised record and
// emit (tileId, originalIdx) at the scan offset, row-major ty -> tx, bounded by maxAssignments.
// Replaces createInstancesKernel / ...32 (DFS.metal:642-788) and createInstancesStereoKernel / ...32
// (DFS.metal:790-864).
#include "gsm_common.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

template <typename TileT, bool STEREO>
__global__ void __launch_bounds__(256) create_instances_kernel(const int32_t* __restrict__ sortedIdx,
                                                               const uint32_t* __restrict__ offsets,
                                                               const int32_t* __restrict__ bounds,
                                                               const void* __restrict__ renderData,
                                                               TileT* __restrict__ tileIds, int32_t* __restrict__ instanceIdx,
                                                               const GSMDepthFirstHeader* __restrict__ header, uint32_t tilesX,
                                                               uint32_t maxAssignments) {
    const uint32_t visibleCount = header->visibleCount;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < visibleCount; i += gridDim.x * 256u) {
        const int32_t originalIdx = sortedIdx[i];
        if (originalIdx < 0) continue;
        const int4 b = __ldg(reinterpret_cast<const int4*>(bounds) + originalIdx);
        const int minTX = b.x, maxTX = b.y, minTY = b.z, maxTY = b.w;
        if (minTX > maxTX || minTY > maxTY) continue;
        uint32_t writeOffset = offsets[i];
        if (STEREO) {
            for (int ty = minTY; ty <= maxTY; ++ty)
                for (int tx = minTX; tx <= maxTX; ++tx)
                    if (writeOffset < maxAssignments) {
                        tileIds[writeOffset] = (TileT)(ty * (int)tilesX + tx);
                        instanceIdx[writeOffset] = originalIdx;
                        writeOffset++;
                    }
        } else {
            const uint4 rd = __ldg(reinterpret_cast<const uint4*>(renderData) + originalIdx);
            QuantSplat q = makeQuantSplat(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)),
                                          __ushort_as_half((unsigned short)(rd.x >> 16)), (uint16_t)(rd.y & 0xFFFFu),
                                          __ushort_as_half((unsigned short)(rd.y >> 16)),
                                          __ushort_as_half((unsigned short)(rd.z & 0xFFFFu)), (uint8_t)(rd.w >> 24));
            if (q.d2Cutoff >= 0.0f) {
                for (int ty = minTY; ty <= maxTY; ++ty)
                    for (int tx = minTX; tx <= maxTX; ++tx)
                        if (tileHit(q, tx, ty) && writeOffset < maxAssignments) {
                            tileIds[writeOffset] = (TileT)(ty * (int)tilesX + tx);
                            instanceIdx[writeOffset] = originalIdx;
                            writeOffset++;
                        }
            }
        }
    }
}

cudaError_t launchCreateInstances(cudaStream_t s, bool stereo, bool tileId16, const int32_t* sortedIdx, const uint32_t* offsets,
                                  const int32_t* bounds, const void* renderData, void* tileIds, int32_t* instanceIdx,
                                  const GSMDepthFirstHeader* header, uint32_t tilesX, uint32_t maxAssignments, uint32_t capVisible) {
    uint32_t grid = (capVisible + 255u) / 256u;
    if (grid == 0) grid = 1;
#define GSM_LAUNCH(T, ST) create_instances_kernel<T, ST><<<grid, 256, 0, s>>>(sortedIdx, offsets, bounds, renderData, (T*)tileIds, instanceIdx, header, tilesX, maxAssignments)
    if (tileId16) { if (stereo) GSM_LAUNCH(uint16_t, true); else GSM_LAUNCH(uint16_t, false); }
    else { if (stereo) GSM_LAUNCH(uint32_t, true); else GSM_LAUNCH(uint32_t, false); }
#undef GSM_LAUNCH
    return cudaGetLastError();
}

}  // namespace gsm
