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
// fully parallel launch. A thread takes one 128-bit load of consecutive ids (8 x u16 or 4 x u32) plus the id before
// them: with one id per thread and a grid-stride loop the kernel was a chain of ~10 dependent L2 round trips per thread
// (11 us for 5.8 MB, ncu r1_v10); the grid is sized for the capacity and threads beyond totalInstances leave at once.
template <typename TileT>
__global__ void __launch_bounds__(256) tile_lower_bounds_kernel(const TileT* __restrict__ sortedTileIds,
                                                                const GSMDepthFirstHeader* __restrict__ header,
                                                                uint32_t tileCount, uint32_t* __restrict__ lowerBounds,
                                                                uint32_t tileLo, uint32_t tileHi) {
    constexpr uint32_t PER = 16u / sizeof(TileT);
    pdlLaunchDependents();
    pdlWait();
    const uint32_t total = ldAfterWait(&header->totalInstances);
    const uint32_t first = (blockIdx.x * 256u + threadIdx.x) * PER;
    if (first > total) return;  // first == total still closes the last run
    TileT ids[PER];
    if (first + PER <= total) {
        const uint4 w = *reinterpret_cast<const uint4*>(sortedTileIds + first);
        memcpy(ids, &w, 16);
    } else {
#pragma unroll
        for (uint32_t k = 0; k < PER; ++k) ids[k] = (first + k < total) ? sortedTileIds[first + k] : (TileT)0;
    }
    int prev = (first > 0) ? (int)min((uint32_t)sortedTileIds[first - 1], tileCount) : -1;
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
        const uint32_t i = first + k;  // boundary i in [0, total]: i == total closes the last run
        if (i > total) break;
        const int cur = (i < total) ? (int)min((uint32_t)ids[k], tileCount) : (int)tileCount;
        // only entries [tileLo, tileHi] are read (a strip's blend reads its own tiles and the end of the last one): the ids of a
        // strip all lie inside it, so the run that closes everything below the strip would otherwise be one thread storing
        // thousands of words in a row (62 us of a 4K half-frame strip before this clamp)
        for (int t = max(prev + 1, (int)tileLo); t <= min(cur, (int)tileHi); ++t) lowerBounds[t] = i;
        prev = cur;
    }
    // the run that ends exactly at a thread's last id is closed by the next thread (its `prev`); the very last boundary
    // (i == total) is handled above when total falls inside this thread's span, or by the thread whose first == total
}

cudaError_t launchTileRanges(cudaStream_t s, bool tileId16, const void* sortedTileIds, const GSMDepthFirstHeader* header,
                             uint32_t tileCount, uint32_t* lowerBounds, uint32_t capInstances, uint32_t tileLo, uint32_t tileHi) {
    if (tileHi > tileCount) tileHi = tileCount;
    if (tileId16) {
        const uint32_t grid = (capInstances / 8u + 1u + 255u) / 256u;
        return launchChained(tile_lower_bounds_kernel<uint16_t>, grid, 256, s, (const uint16_t*)sortedTileIds, header, tileCount, lowerBounds, tileLo, tileHi);
    }
    const uint32_t grid = (capInstances / 4u + 1u + 255u) / 256u;
    return launchChained(tile_lower_bounds_kernel<uint32_t>, grid, 256, s, (const uint32_t*)sortedTileIds, header, tileCount, lowerBounds, tileLo, tileHi);
}

}  // namespace gsm
