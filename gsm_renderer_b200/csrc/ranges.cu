// ranges.cu -- stage 7: per-tile {offset,count} from the tile-sorted instance ids.
// The reference runs two binary searches per tile (extractTileRangesKernel / ...32, DFS.metal:1258-1370):
// offset = lower_bound(tile), count = upper_bound(tile) - lower_bound(tile). Ids are integers, so
// upper_bound(t) == lower_bound(t+1): one boundary-detection pass over the sorted ids writes lower_bound
// for every tile (empty tiles get the position of the next non-empty tile, exactly what the binary search
// returns), and the headers follow. activeTiles is emitted in ascending tile order (the reference's atomic
// append order is nondeterministic).
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

// Boundary detection over the sorted ids; the last CTA to finish (threadfence + counter) turns the lower bounds
// into headers and the ascending active list, so the stage is one launch.
template <typename TileT>
__global__ void __launch_bounds__(256) tile_ranges_kernel(const TileT* __restrict__ sortedTileIds,
                                                          const GSMDepthFirstHeader* __restrict__ header, uint32_t tileCount,
                                                          uint32_t* __restrict__ lowerBounds,
                                                          GSMGaussianHeader* __restrict__ tileHeaders,
                                                          uint32_t* __restrict__ activeTiles, uint32_t* __restrict__ activeTileCount,
                                                          uint32_t* __restrict__ doneCounter) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_chunkTotal;
    __shared__ bool s_last;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t total = header->totalInstances;
    // boundary i in [0, total]: i == total closes the last run
    for (uint32_t i = blockIdx.x * 256u + tid; i <= total; i += gridDim.x * 256u) {
        int cur = (i < total) ? (int)min((uint32_t)sortedTileIds[i], tileCount) : (int)tileCount;
        int prev = (i > 0) ? (int)min((uint32_t)sortedTileIds[i - 1], tileCount) : -1;
        for (int t = prev + 1; t <= cur; ++t) lowerBounds[t] = i;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(doneCounter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint32_t running = 0;  // same value in every thread
    for (uint32_t base = 0; base < tileCount; base += 256u) {
        const uint32_t t = base + tid;
        uint32_t count = 0;
        if (t < tileCount) {
            GSMGaussianHeader h;
            if (total == 0) {  // DFS.metal:1270-1276
                h.offset = 0; h.count = 0;
            } else {
                const uint32_t lo = __ldcg(lowerBounds + t), hi = __ldcg(lowerBounds + t + 1);
                h.offset = lo;
                h.count = hi > lo ? hi - lo : 0u;
            }
            tileHeaders[t] = h;
            count = h.count;
        }
        const uint32_t flag = count > 0 ? 1u : 0u;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            uint32_t w = (lane < 8) ? s_warp[lane] : 0u, inc = w;
            for (int o = 1; o < 8; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= (unsigned)o) inc += x;
            }
            if (lane < 8) s_warp[lane] = inc - w;
            if (lane == 7) s_chunkTotal = inc;
        }
        __syncthreads();
        if (flag) activeTiles[running + s_warp[warp] + __popc(bal & ((1u << lane) - 1u))] = t;  // DFS.metal:1309-1312
        running += s_chunkTotal;
        __syncthreads();
    }
    if (tid == 0) *activeTileCount = running;
}

cudaError_t launchTileRanges(cudaStream_t s, bool tileId16, const void* sortedTileIds, const GSMDepthFirstHeader* header,
                             uint32_t tileCount, uint32_t* lowerBounds, GSMGaussianHeader* tileHeaders, uint32_t* activeTiles,
                             uint32_t* activeTileCount, uint32_t* doneCounter, int numSMs) {
    const int grid = numSMs * 4;
    if (tileId16)
        tile_ranges_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)sortedTileIds, header, tileCount, lowerBounds, tileHeaders,
                                                           activeTiles, activeTileCount, doneCounter);
    else
        tile_ranges_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)sortedTileIds, header, tileCount, lowerBounds, tileHeaders,
                                                           activeTiles, activeTileCount, doneCounter);
    return cudaGetLastError();
}

}  // namespace gsm
