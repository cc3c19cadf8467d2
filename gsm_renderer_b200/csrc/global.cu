// global.cu -- the GlobalRenderer pipeline's bookkeeping kernels: prefix sums, visibility compaction,
// assignment totals, tile headers. Reference: Sources/Renderer/GlobalRenderer/GlobalShaders.metal ("GS.metal"), orchestration
// GlobalRenderer.swift ("GR.swift"). The projection and the tile walks live next to the shared helpers in project.cu, the
// 32 x 16-tile blend next to the half arithmetic in blend.cu, the sort is sort.cu's Onesweep.
//
//   GS.metal:169-208   markVisibility + prefix sum + scatterCompact  -> exclusive scan of the flags, compact (gid order)
//   GS.metal:386-561   block reduce / scan / apply prefix sum         -> three generic kernels below
//   GS.metal:685-703   clamp to maxAssignments, overflow, paddedCount -> global_assign_totals (tail of the block scan)
//   GS.metal:304-363   buildHeadersFromSorted (binary search)         -> global_headers_kernel
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

namespace {

constexpr uint32_t kScanBlock = 1024;   // values per block of the three-kernel prefix sum

// block sums of `in[0 .. n)` (values beyond *nDev, when given, count as zero)
__global__ void __launch_bounds__(256) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ blockSums) {
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * 4u;
    uint32_t v = 0u;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) if (base + k < n) v += in[base + k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31u) == 0u) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0u;
        for (int w = 0; w < 8; ++w) t += s_warp[w];
        blockSums[blockIdx.x] = t;
    }
}

// in-place exclusive scan of the block sums by ONE CTA; the grand total goes to blockSums[numBlocks]
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t* blockSums, uint32_t numBlocks) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < numBlocks; b0 += 1024u) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < numBlocks ? blockSums[i] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if ((threadIdx.x & 31u) >= (unsigned)o) inc += t;
        }
        if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32u) {
            const uint32_t w = s_warp[threadIdx.x];
            uint32_t winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
                if (threadIdx.x >= (unsigned)o) winc += t;
            }
            s_warp[threadIdx.x] = winc - w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[threadIdx.x >> 5] + inc - v;
        if (i < numBlocks) blockSums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023u) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) blockSums[numBlocks] = s_carry;
}

// out[i] = exclusive prefix of in over the whole array
__global__ void __launch_bounds__(256) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ blockSums,
                                                         uint32_t* __restrict__ out) {
    __shared__ uint32_t s_warp[8];
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * 4u;
    uint32_t v[4], sum = 0u;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) { v[k] = base + k < n ? in[base + k] : 0u; sum += v[k]; }
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if ((threadIdx.x & 31u) >= (unsigned)o) inc += t;
    }
    if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t before = blockSums[blockIdx.x];
    for (unsigned w = 0; w < (threadIdx.x >> 5); ++w) before += s_warp[w];
    uint32_t run = before + inc - sum;
#pragma unroll
    for (uint32_t k = 0; k < 4u; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
}

// scatterCompactKernel (GS.metal:183-208)
__global__ void __launch_bounds__(256) global_compact_kernel(GlobalFrame f, uint32_t gaussianCount) {
    const uint32_t gid = blockIdx.x * 256u + threadIdx.x;
    if (gid >= gaussianCount) return;
    const uint32_t flag = f.flags[gid], pos = f.flagOffsets[gid];
    if (flag) f.visibleIndices[pos] = gid;
    if (gid == gaussianCount - 1u) f.header->visibleCount = pos + flag;
}

// writeTotalCountKernel + prepareAssignmentDispatchKernel (GS.metal:551-561, :685-703)
__global__ void global_assign_totals_kernel(GlobalFrame f, uint32_t numBlocks) {
    uint32_t total = f.blockSums[numBlocks];
    f.header->totalRaw = total;
    uint32_t overflow = 0u;
    if (total > f.maxAssignments) { total = f.maxAssignments; overflow = 1u; }
    f.header->totalAssignments = total;
    f.header->overflow = overflow;
    f.header->paddedCount = ((total + kRadixAlignment - 1u) / kRadixAlignment) * kRadixAlignment;
    f.header->activeTileCount = 0u;
    *f.renderTicket = 0u;
}

// buildHeadersFromSortedKernel (GS.metal:304-363)
__global__ void __launch_bounds__(256) global_headers_kernel(GlobalFrame f, uint32_t tileCount) {
    const uint32_t tile = blockIdx.x * 256u + threadIdx.x;
    if (tile >= tileCount) return;
    const uint32_t total = f.header->totalAssignments;
    GSMGaussianHeader h;
    h.offset = 0u; h.count = 0u;
    if (total > 0u) {
        uint32_t left = 0u, right = total;
        while (left < right) {
            const uint32_t mid = (left + right) >> 1;
            if ((f.sortKeys[mid] >> 16) < tile) left = mid + 1u; else right = mid;
        }
        const uint32_t start = left;
        right = total;
        while (left < right) {
            const uint32_t mid = (left + right) >> 1;
            if ((f.sortKeys[mid] >> 16) <= tile) left = mid + 1u; else right = mid;
        }
        h.offset = start;
        h.count = left > start ? left - start : 0u;
    }
    f.tileHeaders[tile] = h;
    if (h.count > 0u) f.activeTiles[atomicAdd(&f.header->activeTileCount, 1u)] = tile;
}

cudaError_t exclusiveScan(cudaStream_t s, const uint32_t* in, uint32_t n, uint32_t* blockSums, uint32_t* out) {
    const uint32_t blocks = (n + kScanBlock - 1u) / kScanBlock;
    scan_reduce_kernel<<<blocks, 256, 0, s>>>(in, n, blockSums);
    scan_blocks_kernel<<<1, 1024, 0, s>>>(blockSums, blocks);
    scan_apply_kernel<<<blocks, 256, 0, s>>>(in, n, blockSums, out);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launchGlobalCompact(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount) {
    cudaError_t e = exclusiveScan(s, f.flags, gaussianCount, f.blockSums, f.flagOffsets);
    if (e != cudaSuccess) return e;
    global_compact_kernel<<<(gaussianCount + 255u) / 256u, 256, 0, s>>>(f, gaussianCount);
    return cudaGetLastError();
}

cudaError_t launchGlobalAssignOffsets(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount) {
    cudaError_t e = exclusiveScan(s, f.counts, gaussianCount, f.blockSums, f.offsets);   // counts beyond visibleCount are zero
    if (e != cudaSuccess) return e;
    global_assign_totals_kernel<<<1, 1, 0, s>>>(f, (gaussianCount + kScanBlock - 1u) / kScanBlock);
    return cudaGetLastError();
}

cudaError_t launchGlobalHeaders(cudaStream_t s, const GlobalFrame& f) {
    const uint32_t tileCount = f.tilesX * f.tilesY;
    global_headers_kernel<<<(tileCount + 255u) / 256u, 256, 0, s>>>(f, tileCount);
    return cudaGetLastError();
}

}  // namespace gsm
