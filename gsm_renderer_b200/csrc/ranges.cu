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

template <typename TileT>
__global__ void __launch_bounds__(256) tile_lower_bounds_kernel(const TileT* __restrict__ sortedTileIds,
                                                                const GSMDepthFirstHeader* __restrict__ header,
                                                                uint32_t tileCount, uint32_t* __restrict__ lowerBounds) {
    const uint32_t total = header->totalInstances;
    // boundary i in [0, total]: i == total closes the last run
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i <= total; i += gridDim.x * 256u) {
        int cur = (i < total) ? (int)min((uint32_t)sortedTileIds[i], tileCount) : (int)tileCount;
        int prev = (i > 0) ? (int)min((uint32_t)sortedTileIds[i - 1], tileCount) : -1;
        for (int t = prev + 1; t <= cur; ++t) lowerBounds[t] = i;
    }
}

__global__ void __launch_bounds__(1024) tile_headers_kernel(const uint32_t* __restrict__ lowerBounds,
                                                            const GSMDepthFirstHeader* __restrict__ header, uint32_t tileCount,
                                                            GSMGaussianHeader* __restrict__ tileHeaders,
                                                            uint32_t* __restrict__ activeTiles,
                                                            uint32_t* __restrict__ activeTileCount) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_chunkTotal;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t total = header->totalInstances;
    uint32_t running = 0;  // same value in every thread
    for (uint32_t base = 0; base < tileCount; base += 1024u) {
        const uint32_t t = base + tid;
        uint32_t count = 0;
        if (t < tileCount) {
            GSMGaussianHeader h;
            if (total == 0) {  // DFS.metal:1270-1276
                h.offset = 0; h.count = 0;
            } else {
                uint32_t lo = lowerBounds[t], hi = lowerBounds[t + 1];
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
            uint32_t w = s_warp[lane], inc = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= (unsigned)o) inc += x;
            }
            s_warp[lane] = inc - w;
            if (lane == 31) s_chunkTotal = inc;
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
                             uint32_t* activeTileCount, int numSMs) {
    const int grid = numSMs * 8;
    if (tileId16)
        tile_lower_bounds_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)sortedTileIds, header, tileCount, lowerBounds);
    else
        tile_lower_bounds_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)sortedTileIds, header, tileCount, lowerBounds);
    tile_headers_kernel<<<1, 1024, 0, s>>>(lowerBounds, header, tileCount, tileHeaders, activeTiles, activeTileCount);
    return cudaGetLastError();
}

}  // namespace gsm
