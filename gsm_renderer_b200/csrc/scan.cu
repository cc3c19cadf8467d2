// scan.cu -- stages 3+4 fused: ordered[i] = nTouched[sortedIdx[i]] gathered on load, exclusive scan with a
// single-pass decoupled look-back, offsets written once.
// Replaces applyDepthOrderingKernel (DFS.metal:623-640) and the 5-kernel prefix sum (DFS.metal:2036-2139,
// driver InstanceExpansionEncoder.swift:83-176).
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

constexpr int kScanItems = 8;
constexpr int kScanTile = 256 * kScanItems;

__global__ void __launch_bounds__(256) apply_order_scan_kernel(const int32_t* __restrict__ sortedIdx,
                                                               const uint32_t* __restrict__ nTouched,
                                                               uint32_t* __restrict__ offsetsOut,
                                                               const GSMDepthFirstHeader* __restrict__ header,
                                                               unsigned long long* status, uint32_t* ticket, SortReset reset) {
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_tile, s_base;
    const unsigned tid = threadIdx.x;
    const uint32_t count = header->visibleCount;
    {   // reset the tile sort's look-back words for exactly the tiles this frame's totalInstances needs
        const uint32_t words = ((header->totalInstances + reset.tileSize - 1u) / reset.tileSize) * 256u;
        const uint32_t gwords = ((words / 256u + 15u) / 16u) * 256u;
        for (uint32_t p = 0; p < reset.passes; ++p) {
            for (uint32_t i = blockIdx.x * 256u + tid; i < words; i += gridDim.x * 256u) reset.status[(size_t)p * reset.statusStride + i] = 0u;
            for (uint32_t i = blockIdx.x * 256u + tid; i < gwords; i += gridDim.x * 256u) reset.gstatus[(size_t)p * reset.gstatusStride + i] = 0u;
        }
    }
    const uint32_t numTiles = (count + kScanTile - 1) / kScanTile;
    while (true) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t base = tile * kScanTile + tid * kScanItems;  // blocked: a thread owns 8 consecutive elements
        uint32_t v[kScanItems];
        int32_t idx[kScanItems];
        if (base + kScanItems <= count) {
            const int4* p = reinterpret_cast<const int4*>(sortedIdx + base);
            int4 a = p[0], b = p[1];
            idx[0] = a.x; idx[1] = a.y; idx[2] = a.z; idx[3] = a.w;
            idx[4] = b.x; idx[5] = b.y; idx[6] = b.z; idx[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < kScanItems; ++i) idx[i] = (base + i < count) ? sortedIdx[base + i] : -1;
        }
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            v[i] = (idx[i] >= 0) ? __ldg(nTouched + idx[i]) : 0u;  // DFS.metal:633-639
            sum += v[i];
        }
        uint32_t total;
        uint32_t excl = block_exclusive_scan_256(sum, s_scan, total);
        if (tid < 32) {
            uint32_t b = lookback_exclusive(status, tile, total);
            if (tid == 0) s_base = b;
        }
        __syncthreads();
        uint32_t run = s_base + excl;
        uint32_t o[kScanItems];
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) { o[i] = run; run += v[i]; }
        if (base + kScanItems <= count) {
            uint4* q = reinterpret_cast<uint4*>(offsetsOut + base);
            q[0] = make_uint4(o[0], o[1], o[2], o[3]);
            q[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int i = 0; i < kScanItems; ++i) if (base + i < count) offsetsOut[base + i] = o[i];
        }
        __syncthreads();
    }
}

cudaError_t launchApplyOrderScan(cudaStream_t s, const int32_t* sortedIdx, const uint32_t* nTouched, uint32_t* offsetsOut,
                                 const GSMDepthFirstHeader* header, unsigned long long* status, uint32_t* ticket, int numSMs,
                                 const SortReset& reset) {
    apply_order_scan_kernel<<<numSMs * 4, 256, 0, s>>>(sortedIdx, nTouched, offsetsOut, header, status, ticket, reset);
    return cudaGetLastError();
}

}  // namespace gsm
