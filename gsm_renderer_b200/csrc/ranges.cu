// ranges.cu -- stage 7: per-tile {offset,count} from the tile-sorted instance ids.
// The reference runs two binary searches per tile (extractTileRangesKernel / ...32, DFS.metal:1258-1370):
// offset = lower_bound(tile), count = upper_bound(tile) - lower_bound(tile). Ids are integers, so
// upper_bound(t) == lower_bound(t+1): one boundary-detection pass over the sorted ids writes lower_bound
// for every tile (empty tiles get the position of the next non-empty tile, exactly what the binary search
// returns), and the headers follow. activeTiles is appended with an atomic, like the reference
// (DFS.metal:1309-1312): its order is nondeterministic there too -- compare as a set.
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

// Boundary detection over the sorted ids. The {offset,count} headers and the active-tile list are written by the
// blend kernel's CTA for each tile (it reads lowerBounds[t], lowerBounds[t+1] anyway), so this stage is one
// fully parallel launch.
template <typename TileT>
__global__ void __launch_bounds__(256) tile_lower_bounds_kernel(const TileT* __restrict__ sortedTileIds,
                                                                const GSMDepthFirstHeader* __restrict__ header,
                                                                uint32_t tileCount, uint32_t* __restrict__ lowerBounds) {
    pdlLaunchDependents();
    pdlWait();
    const uint32_t total = header->totalInstances;
    // boundary i in [0, total]: i == total closes the last run
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i <= total; i += gridDim.x * 256u) {
        int cur = (i < total) ? (int)min((uint32_t)sortedTileIds[i], tileCount) : (int)tileCount;
        int prev = (i > 0) ? (int)min((uint32_t)sortedTileIds[i - 1], tileCount) : -1;
        for (int t = prev + 1; t <= cur; ++t) lowerBounds[t] = i;
    }
}

cudaError_t launchTileRanges(cudaStream_t s, bool tileId16, const void* sortedTileIds, const GSMDepthFirstHeader* header,
                             uint32_t tileCount, uint32_t* lowerBounds, int numSMs) {
    const int grid = numSMs * 8;
    if (tileId16)
        launchChained(tile_lower_bounds_kernel<uint16_t>, grid, 256, s, (const uint16_t*)sortedTileIds, header, tileCount, lowerBounds);
    else
        launchChained(tile_lower_bounds_kernel<uint32_t>, grid, 256, s, (const uint32_t*)sortedTileIds, header, tileCount, lowerBounds);
    return cudaGetLastError();
}

}  // namespace gsm
