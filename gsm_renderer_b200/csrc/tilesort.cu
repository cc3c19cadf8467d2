// tilesort.cu -- second half of the tile sort when it runs most-significant-digit first.
//
// The reference sorts the instances' tile ids with two (16-bit ids) stable 8-bit LSD passes of five kernels each
// (tileRadix* DFS.metal:866-1256, TileSortEncoder.swift:51-178) and then extracts the tile ranges (DFS.metal:1258-1370).
// Any stable sort gives the same arrays, and an Onesweep pass costs ~200 instructions per key (ranking with 8 ballots,
// staging, look-back bookkeeping). For frames of up to a few million instances the tile sort is therefore:
//   1. ONE Onesweep pass (sort.cu) on the id's HIGH byte, digit = (id >> L) & 0xFF with L = bits(tileCount - 1) - 8: a stable
//      scatter into <= 256 buckets of 2^L consecutive tiles; the bucket histogram comes fused from the expansion kernel;
//   2. two kernels here finish every bucket with a stable counting sort on the L low bits. Buckets are cut into CHUNKS of
//      4096 ids, one CTA per chunk: tile_chunk_count_kernel counts the ids per tile of its chunk; tile_chunk_place_kernel sums
//      the counts of its bucket's chunks, ranks its ids with L ballots each and stores them at tile offset + ids of the same
//      tile in earlier chunks + rank. The per-tile sums ARE the tile ranges, so the first chunk of a bucket also writes the
//      lower bounds of its tiles and the separate range kernel is not launched.
// Work is balanced by chunk, not by bucket: a bucket under the screen centre holds several times the average. (One persistent
// kernel taking count items, then place items, from a ticket was 3x slower: every item is a chain of dependent round trips.)
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

namespace {

constexpr int kTlThreads = 512;
constexpr int kTlWarps = kTlThreads / 32;
constexpr int kTlItems = 8;                                  // ids per thread and chunk
constexpr uint32_t kTlChunk = kTlThreads * kTlItems;         // 4096 = the Onesweep pass's tile (sortTileSize(16, false))
constexpr int kTlMaxBins = 256;                              // 2^L, L <= 8

template <int LB>
__device__ __forceinline__ uint32_t warpRankLow(uint32_t d, uint32_t* warpRow, unsigned lane) {
    unsigned peers = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < LB; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
        peers &= bit ? bal : ~bal;
    }
    const uint32_t lower = __popc(peers & ((1u << lane) - 1u));
    uint32_t pre = 0;
    if (lower == 0) pre = atomicAdd(warpRow + d, (uint32_t)__popc(peers));
    pre = __shfl_sync(0xFFFFFFFFu, pre, __ffs(peers) - 1);
    return pre + lower;
}

struct TileLocalShared {
    uint32_t rows[kTlWarps][kTlMaxBins];
    uint32_t globalBase[kTlMaxBins];          // output slot of a tile's first id of this chunk, minus its chunk-local offset
    uint32_t stageVals[kTlChunk];             // the chunk in tile order, staged so that the stores are contiguous runs per tile
    unsigned short stageKeys[kTlChunk];
    uint32_t bucketStart[257];    // exclusive prefix of the bucket histogram
    uint32_t chunkBase[257];      // exclusive prefix of the buckets' chunk counts
    uint32_t scan[kTlWarps + 1];
};

// exclusive scan of one value per thread over the CTA (kTlThreads threads)
__device__ __forceinline__ uint32_t blockExclusive512(uint32_t v, uint32_t* smem, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < (unsigned)kTlWarps ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < kTlWarps; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= (unsigned)o) winc += t;
        }
        if (lane < (unsigned)kTlWarps) smem[lane] = winc - w;
        if (lane == kTlWarps - 1) smem[kTlWarps] = winc;
    }
    __syncthreads();
    const uint32_t r = smem[warp] + inc - v;
    total = smem[kTlWarps];
    __syncthreads();
    return r;
}

// PLACE item of chunk c of a bucket: `running[b]` = first output slot of tile b's ids of this chunk
template <int LB>
__device__ __forceinline__ void placeChunk(const unsigned short* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                           unsigned short* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t base,
                                           uint32_t n, uint32_t running, TileLocalShared& sh) {
    constexpr uint32_t BINS = 1u << LB, MASK = BINS - 1u;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // warp-striped: element (warp, item, lane) = warp*ITEMS*32 + item*32 + lane, i.e. index order inside the chunk
    uint32_t key[kTlItems], val[kTlItems], rank[kTlItems];
    const uint32_t wb = warp * (kTlItems * 32u) + lane;
#pragma unroll
    for (int i = 0; i < kTlItems; ++i) {
        const uint32_t j = wb + i * 32u;
        key[i] = j < n ? (uint32_t)keysIn[base + j] : 0xFFFFu;
    }
#pragma unroll
    for (int i = 0; i < kTlItems; ++i) {
        const uint32_t j = wb + i * 32u;
        // padding: last bin, last in index order -- ranked behind every real id, never stored
        rank[i] = warpRankLow<LB>(j < n ? key[i] & MASK : MASK, sh.rows[warp], lane);
    }
#pragma unroll
    for (int i = 0; i < kTlItems; ++i) {
        const uint32_t j = wb + i * 32u;
        val[i] = j < n ? valsIn[base + j] : 0u;
    }
    __syncthreads();
    uint32_t cnt = 0u;
    if (tid < BINS) {   // thread b: exclusive prefix over the warps, the tile's count in this chunk
#pragma unroll
        for (int w = 0; w < kTlWarps; ++w) {
            const uint32_t c = sh.rows[w][tid];
            sh.rows[w][tid] = cnt;
            cnt += c;
        }
    }
    uint32_t chunkTotal;
    const uint32_t binExcl = blockExclusive512(cnt, sh.scan, chunkTotal);   // chunk-local offset of the tile (padding sits last)
    if (tid < BINS) {
#pragma unroll
        for (int w = 0; w < kTlWarps; ++w) sh.rows[w][tid] += binExcl;
        sh.globalBase[tid] = running - binExcl;
    }
    __syncthreads();
    // stage the chunk in tile order (a direct scatter made every 2- and 4-byte store its own 32-byte sector: 43 us), then
    // write it out linearly: ids of one tile are a contiguous run on both sides
#pragma unroll
    for (int i = 0; i < kTlItems; ++i) {
        const uint32_t j = wb + i * 32u;
        const uint32_t p = sh.rows[warp][j < n ? key[i] & MASK : MASK] + rank[i];
        sh.stageKeys[p] = (unsigned short)key[i];
        sh.stageVals[p] = val[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kTlItems; ++i) {
        const uint32_t j = tid + i * kTlThreads;
        if (j < n) {
            const uint32_t k = sh.stageKeys[j];
            const uint32_t dst = sh.globalBase[k & MASK] + j;
            keysOut[dst] = (unsigned short)k;
            valsOut[dst] = sh.stageVals[j];
        }
    }
}

}  // namespace

// chunk numbering shared by the two kernels: bucket offsets and chunk bases from the bucket histogram, computed by ONE warp per
// CTA (eight buckets per lane; with all 16 warps in two block-wide scans this prologue was most of the kernels' instructions)
__device__ __forceinline__ uint32_t numberChunks(const uint32_t* bucketHist, TileLocalShared& sh) {
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    if (tid < 32u) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ldAfterWait(bucketHist + 8u * lane + k);
        uint32_t ids = 0u, chunks = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ids += v[k]; chunks += (v[k] + kTlChunk - 1u) / kTlChunk; }
        uint32_t idInc = ids, chunkInc = chunks;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, idInc, o), c = __shfl_up_sync(0xFFFFFFFFu, chunkInc, o);
            if (lane >= (unsigned)o) { idInc += a; chunkInc += c; }
        }
        uint32_t idRun = idInc - ids, chunkRun = chunkInc - chunks;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sh.bucketStart[8u * lane + k] = idRun; sh.chunkBase[8u * lane + k] = chunkRun;
            idRun += v[k]; chunkRun += (v[k] + kTlChunk - 1u) / kTlChunk;
        }
        if (lane == 31u) { sh.bucketStart[256] = idRun; sh.chunkBase[256] = chunkRun; }
    }
    __syncthreads();
    return sh.chunkBase[256];
}
__device__ __forceinline__ uint32_t bucketOfChunk(uint32_t chunk, const TileLocalShared& sh) {
    uint32_t b = 0;   // largest b with chunkBase[b] <= chunk (the non-empty bucket whose chunks include it)
#pragma unroll
    for (uint32_t step = 128; step > 0; step >>= 1)
        if (b + step <= 256u && sh.chunkBase[b + step] <= chunk) b += step;
    return b;
}

// COUNT: one CTA per chunk -- ids per tile of the chunk (warp-private rows, one shared-memory RED per id)
__global__ void __launch_bounds__(kTlThreads, 4) tile_chunk_count_kernel(const unsigned short* __restrict__ keysIn, const uint32_t* bucketHist,
                                                                         const GSMDepthFirstHeader* header, uint32_t capInstances,
                                                                         uint32_t lowBits, uint32_t* __restrict__ chunkCounts) {
    __shared__ TileLocalShared sh;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t bins = 1u << lowBits;
    pdlLaunchDependents();
    for (uint32_t i = tid; i < kTlWarps * bins; i += kTlThreads) sh.rows[i >> lowBits][i & (bins - 1u)] = 0u;
    pdlWait();
    if (blockIdx.x >= min(ldAfterWait(&header->totalInstances), capInstances) / kTlChunk + 257u) return;   // bound on the chunk count
    const uint32_t totalChunks = numberChunks(bucketHist, sh);
    for (uint32_t chunk = blockIdx.x; chunk < totalChunks; chunk += gridDim.x) {
    const uint32_t b = bucketOfChunk(chunk, sh);
    const uint32_t c = chunk - sh.chunkBase[b];
    const uint32_t bucketBase = sh.bucketStart[b], bucketN = sh.bucketStart[b + 1] - bucketBase;
    const uint32_t base = bucketBase + c * kTlChunk, n = min(kTlChunk, bucketN - c * kTlChunk);
    const uint32_t wb = warp * (kTlItems * 32u) + lane;
    uint32_t key[kTlItems];
#pragma unroll
    for (int i = 0; i < kTlItems; ++i) {
        const uint32_t j = wb + i * 32u;
        key[i] = j < n ? (uint32_t)keysIn[base + j] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int i = 0; i < kTlItems; ++i)
        if (key[i] != 0xFFFFFFFFu) atomicAdd(&sh.rows[warp][key[i] & (bins - 1u)], 1u);
    __syncthreads();
    if (tid < bins) {
        uint32_t cnt = 0u;
#pragma unroll
        for (int w = 0; w < kTlWarps; ++w) cnt += sh.rows[w][tid];
        chunkCounts[(size_t)chunk * bins + tid] = cnt;
    }
    if (chunk + gridDim.x < totalChunks) {   // frames beyond one chunk per CTA
        __syncthreads();
        for (uint32_t i = tid; i < kTlWarps * bins; i += kTlThreads) sh.rows[i >> lowBits][i & (bins - 1u)] = 0u;
        __syncthreads();
    }
    }
}

// PLACE: one CTA per chunk -- per tile of the bucket, ids in all its chunks (-> tile offsets, the tile ranges) and in the chunks
// before this one; then rank and store
__global__ void __launch_bounds__(kTlThreads, 2) tile_chunk_place_kernel(const unsigned short* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                         unsigned short* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                                         const uint32_t* bucketHist, const GSMDepthFirstHeader* header,
                                                                         uint32_t capInstances, uint32_t lowBits, uint32_t tileCount,
                                                                         uint32_t* __restrict__ lowerBounds, const uint32_t* chunkCounts) {
    __shared__ TileLocalShared sh;
    const unsigned tid = threadIdx.x;
    const uint32_t bins = 1u << lowBits;
    pdlLaunchDependents();
    for (uint32_t i = tid; i < kTlWarps * bins; i += kTlThreads) sh.rows[i >> lowBits][i & (bins - 1u)] = 0u;
    pdlWait();
    const uint32_t total = min(ldAfterWait(&header->totalInstances), capInstances);
    if (blockIdx.x >= total / kTlChunk + 257u && blockIdx.x >= 256u) return;   // bound on the chunk count; CTAs 0..255 also own a bucket's empty case
    const uint32_t totalChunks = numberChunks(bucketHist, sh);
    // tiles of EMPTY buckets have no chunk to write their (empty) range: CTA b < 256 does it for bucket b (the grid has >= 256 CTAs)
    if (blockIdx.x < 256u && sh.bucketStart[blockIdx.x + 1] == sh.bucketStart[blockIdx.x]) {
        const uint32_t t = (blockIdx.x << lowBits) + tid;
        if (tid < bins && t < tileCount) lowerBounds[t] = sh.bucketStart[blockIdx.x];
    }
    if (blockIdx.x == 0u && tid == 0) lowerBounds[tileCount] = total;
    for (uint32_t chunk = blockIdx.x; chunk < totalChunks; chunk += gridDim.x) {
    const uint32_t b = bucketOfChunk(chunk, sh);
    const uint32_t c = chunk - sh.chunkBase[b];
    const uint32_t bucketBase = sh.bucketStart[b], bucketN = sh.bucketStart[b + 1] - bucketBase;
    const uint32_t base = bucketBase + c * kTlChunk, n = min(kTlChunk, bucketN - c * kTlChunk);
    uint32_t totalOfBin = 0u, before = 0u;
    if (tid < bins) {
        const uint32_t nChunks = sh.chunkBase[b + 1] - sh.chunkBase[b];
        const uint32_t* row = chunkCounts + (size_t)sh.chunkBase[b] * bins + tid;
        for (uint32_t k0 = 0; k0 < nChunks; k0 += 8u) {
            uint32_t w[8];
#pragma unroll
            for (uint32_t k = 0; k < 8u; ++k) w[k] = k0 + k < nChunks ? __ldcg(row + (size_t)(k0 + k) * bins) : 0u;
#pragma unroll
            for (uint32_t k = 0; k < 8u; ++k) {
                totalOfBin += w[k];
                if (k0 + k < c) before += w[k];
            }
        }
    }
    uint32_t bucketTotal;
    const uint32_t binExcl = blockExclusive512(totalOfBin, sh.scan, bucketTotal);
    const uint32_t tileStart = bucketBase + binExcl;
    if (c == 0u && tid < bins) {   // the tile ranges (DFS.metal:1258-1370): lowerBounds[t] = first instance of tile t
        const uint32_t tile = (b << lowBits) + tid;
        if (tile < tileCount) lowerBounds[tile] = tileStart;
    }
    const uint32_t running = tileStart + before;
    switch (lowBits) {
        case 1: placeChunk<1>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        case 2: placeChunk<2>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        case 3: placeChunk<3>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        case 4: placeChunk<4>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        case 5: placeChunk<5>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        case 6: placeChunk<6>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        case 7: placeChunk<7>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
        default: placeChunk<8>(keysIn, valsIn, keysOut, valsOut, base, n, running, sh); break;
    }
    if (chunk + gridDim.x < totalChunks) {   // frames beyond one chunk per CTA
        __syncthreads();
        for (uint32_t i = tid; i < kTlWarps * bins; i += kTlThreads) sh.rows[i >> lowBits][i & (bins - 1u)] = 0u;
        __syncthreads();
    }
    }
}

// lowBits of the MSD tile sort for a frame of tileCount tiles: 0 = the id has at most 8 bits, the LSD pass alone sorts it
uint32_t tileSortLowBits(uint32_t tileCount) {
    uint32_t v = tileCount > 1u ? tileCount - 1u : 1u, bits = 0;
    while (v) { bits++; v >>= 1; }
    return bits > 8u ? bits - 8u : 0u;
}

cudaError_t launchTileLocalSort(cudaStream_t s, const void* keysIn, const uint32_t* valsIn, void* keysOut, uint32_t* valsOut,
                                const uint32_t* bucketHist, const GSMDepthFirstHeader* header, uint32_t capInstances, uint32_t lowBits,
                                uint32_t tileCount, uint32_t* lowerBounds, uint32_t* chunkCounts, int numSMs) {
    // one CTA per chunk; the chunk count lives on the device, the grid covers its bound and the surplus CTAs return at once
    uint32_t grid = capInstances / kTlChunk + 257u;
    const uint32_t persistent = (uint32_t)numSMs * 8u;   // larger frames: every CTA loops over chunks (>= 256 CTAs: the empty-bucket ranges)
    if (grid > persistent) grid = persistent;
    launchChained(tile_chunk_count_kernel, (int)grid, kTlThreads, s, (const unsigned short*)keysIn, bucketHist, header, capInstances, lowBits,
                  chunkCounts);
    launchChained(tile_chunk_place_kernel, (int)grid, kTlThreads, s, (const unsigned short*)keysIn, valsIn, (unsigned short*)keysOut,
                  valsOut, bucketHist, header, capInstances, lowBits, tileCount, lowerBounds, (const uint32_t*)chunkCounts);
    return cudaGetLastError();
}

}  // namespace gsm
