// bucketsort.cu -- the frame's depth sort as ONE global pass plus one shared-memory pass.
//
// The reference sorts the visible Gaussians' 32-bit depth keys with four 8-bit LSD passes of five kernels each
// (depthSort* DFS.metal:1387-1696, DepthRadixSortEncoder.swift:139-217). Any stable ascending sort yields the same
// arrays bit for bit, so at frame size -- where an LSD pass costs ~14 us of fixed latency however few keys it moves
// (profiles/README.md, sort_trace) -- frames of up to kDepthBucketMaxGaussians Gaussians replace the four passes by:
//   1. bucket_rank_kernel + bucket_scatter_kernel: a stable scatter of (key, gid) into up to 512 BUCKETS of consecutive key
//      ranges holding ~2048 keys each. The bucket boundaries adapt to the depth distribution: the projection kernel recorded the
//      frame's key range, the compaction kernel counted every 8th stored key into an 8192-bin histogram over that range, and
//      every CTA of the rank kernel turns that sample into the same bin -> bucket table (in shared memory, under its key loads).
//      Mechanically this is an Onesweep pass whose "digit" is the bucket id, cut in two at its only global dependency: the rank
//      kernel ranks a tile in index order (9 ballots per key), stores every key's bucket and tile-local position and publishes
//      the tile's bucket counts; the scatter kernel, one kernel boundary later, sums its predecessors' counts AND the bucket
//      totals -- the exact offsets, nothing depends on the sample being representative -- stages the tile in bucket order
//      and writes it out. (As ONE kernel every tile had to wait for every other tile's counts: that needs all tiles' CTAs
//      resident together, which two frames in flight on two streams can deny each other forever.)
//   2. bucket_local_sort_kernel: one CTA per bucket, in shared memory. (key, position) pairs are unique, so ANY sort of the
//      pairs is the stable sort of the keys: the bucket is split into 1024 bins by the top bits of key - bucketMin with one
//      shared-memory atomic per element, then every element is placed by counting the smaller pairs of its own bin (two or
//      three when the keys spread over the bins, as in a thin depth slice). Writes the sorted pairs and, like the LSD sort's
//      last pass, the depth-ordered tile counts (apply-depth-order, DFS.metal:623-640).
//      Buckets that leave that path are still sorted by the same CTA, slower, never wrong: a bin above 32 elements (equal or
//      clustered keys next to an outlier) -> stable ballot-ranked LSD passes in shared memory; more than 4096 keys (thousands
//      of Gaussians inside one fine bin) -> streaming LSD passes through global memory; all keys equal -> a copy.
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

namespace {

constexpr int kBkThreads = 256;
constexpr int kBkWarps = kBkThreads / 32;
constexpr int kBkBins = (int)kDepthMaxBuckets;      // 512: two adjacent bins per thread
constexpr uint32_t kBkGroup = 16;                   // tiles per group of the direct-summation prefix

// Rank of a warp's 32 elements inside their digit, in lane order; the lowest lane of every digit group bumps the warp's
// private counter once (returning atomics on shared memory) and broadcasts the old value.
template <int DB>
__device__ __forceinline__ uint32_t warpRankDigit(uint32_t d, uint32_t* warpRow, unsigned lane) {
    unsigned peers = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < DB; ++b) {
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

// exclusive scan over the CTA of one value per thread (kBkThreads threads); smem: 9 words
__device__ __forceinline__ uint32_t blockExclusive(uint32_t v, uint32_t* smem9, uint32_t& total) {
    return block_exclusive_scan_256(v, smem9, total);
}

__device__ __forceinline__ uint2 ld_status32x2(const uint32_t* p) {   // p 8-byte aligned
    const unsigned long long v = ld_status64(reinterpret_cast<const unsigned long long*>(p));
    return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}

}  // namespace

// ---------------------------------------------------------------- pass 1: stable scatter into buckets
// One tile per CTA, all tiles co-resident (the launcher sizes the grid from the occupancy of this kernel): 8 keys per thread
// while that covers the frame, else 11. Packed scan word: bucket totals (<= 2^20) above tile-local counts (< 2^12).
constexpr int kScatterItemsMax = 11;   // 8 + 3: keeps the kernel's static shared memory under 48 KB
struct ScatterShared {
    uint32_t warpHist[kBkWarps][kBkBins];
    uint32_t binExcl[kBkBins];
    uint32_t globalBase[kBkBins];
    uint32_t scan[9];
    uint32_t keyMin, fineShift, numBuckets;
    // the bin -> bucket table is dead once the keys are ranked; the tile is then staged over it
    union {
        unsigned short bucketOf[kDepthFineBins];
        struct { uint32_t keys[kBkThreads * kScatterItemsMax]; uint32_t vals[kBkThreads * kScatterItemsMax]; } stage;
    } u;
    unsigned short bucket[kBkThreads * kScatterItemsMax];
};

// Every CTA builds the same table: bucket(bin) = exclusive sample prefix / (T / stride), so a bucket expects ~T keys.
// The sample counts (< 65536 per bin: at most 1e6 / 8 samples in all) are loaded coalesced into the table's own storage as
// 16-bit words, then every thread scans its 32 consecutive bins in place.
__device__ __forceinline__ void loadSampleCounts(const uint32_t* __restrict__ fineHist, const KeyRange* keyRange, ScatterShared& sh) {
    constexpr uint32_t PER = kDepthFineBins / kBkThreads;   // 32 bins per thread
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    {
        uint4 v[PER / 4];
#pragma unroll
        for (uint32_t i = 0; i < PER / 4; ++i) v[i] = __ldcg(reinterpret_cast<const uint4*>(fineHist) + i * kBkThreads + tid);
        uint32_t hi = 0u, lo = 0u;
        if (tid < 32u) { hi = __ldcg(&keyRange->maxKey[lane]); lo = __ldcg(&keyRange->maxInvKey[lane]); }
#pragma unroll
        for (uint32_t i = 0; i < PER / 4; ++i)
            *reinterpret_cast<uint2*>(sh.u.bucketOf + 4u * (i * kBkThreads + tid)) = make_uint2(v[i].x | (v[i].y << 16), v[i].z | (v[i].w << 16));
        if (tid < 32u) {   // the frame's key range, as the compaction kernel derived it
            hi = __reduce_max_sync(0xFFFFFFFFu, hi);
            lo = __reduce_max_sync(0xFFFFFFFFu, lo);
            if (lane == 0) {
                const bool valid = (hi | lo) != 0u && ~lo <= hi;
                const uint32_t span = valid ? hi - ~lo : 0xFFFFFFFFu;
                const int bits = 32 - __clz(span);
                sh.keyMin = valid ? ~lo : 0u;
                sh.fineShift = bits > 13 ? (uint32_t)(bits - 13) : 0u;
            }
        }
    }
}
__device__ __forceinline__ void scanSampleCounts(ScatterShared& sh) {   // after loadSampleCounts
    constexpr uint32_t PER = kDepthFineBins / kBkThreads;   // 32 consecutive bins per thread
    constexpr uint32_t TS = kDepthBucketTarget / kDepthSampleStride;
    const unsigned tid = threadIdx.x;
    __syncthreads();
    uint32_t pk[PER / 2];
    uint4* mine = reinterpret_cast<uint4*>(sh.u.bucketOf + tid * PER);
#pragma unroll
    for (uint32_t i = 0; i < PER / 8; ++i) {
        const uint4 v = mine[i];
        pk[4 * i] = v.x; pk[4 * i + 1] = v.y; pk[4 * i + 2] = v.z; pk[4 * i + 3] = v.w;
    }
    uint32_t sum = 0u;
#pragma unroll
    for (uint32_t i = 0; i < PER / 2; ++i) sum += (pk[i] & 0xFFFFu) + (pk[i] >> 16);
    uint32_t total;
    uint32_t excl = blockExclusive(sum, sh.scan, total);
    if (tid == 0) sh.numBuckets = total > 0u ? min((total - 1u) / TS + 1u, kDepthMaxBuckets) : 0u;
#pragma unroll
    for (uint32_t i = 0; i < PER / 2; ++i) {
        const uint32_t c0 = pk[i] & 0xFFFFu, c1 = pk[i] >> 16;
        const uint32_t b0 = min(excl / TS, kDepthMaxBuckets - 1u);
        const uint32_t b1 = min((excl + c0) / TS, kDepthMaxBuckets - 1u);
        pk[i] = b0 | (b1 << 16);
        excl += c0 + c1;
    }
#pragma unroll
    for (uint32_t i = 0; i < PER / 8; ++i) mine[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
    __syncthreads();
}

// Kernel 1a: bucket and final tile-local position of every key of the tile, the tile's bucket counts published.
template <int ITEMS>
__device__ __forceinline__ void bucketRankTile(const uint32_t* __restrict__ keysIn, uint32_t count, const uint32_t* __restrict__ fineHist,
                                               const KeyRange* keyRange, DepthPlan* __restrict__ plan, uint32_t* status, uint32_t* gstatus,
                                               uint32_t* __restrict__ place, ScatterShared& sh) {
    constexpr uint32_t TILE = kBkThreads * ITEMS;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t numTiles = (count + TILE - 1u) / TILE;
    const uint32_t tile = blockIdx.x;
    if (tile >= numTiles) return;
    const uint32_t bin0 = 2u * tid;
    const uint32_t base = tile * TILE;
    const uint32_t tileValid = min(TILE, count - base);

    loadSampleCounts(fineHist, keyRange, sh);
    // warp-striped: element (warp, item, lane) has index base + warp*ITEMS*32 + item*32 + lane
    uint32_t key[ITEMS], br[ITEMS];   // br: bucket in the high half, rank inside the bucket in the low half
    const uint32_t warpBase = warp * ITEMS * 32u + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = warpBase + i * 32u;
        key[i] = (j < tileValid) ? keysIn[base + j] : 0xFFFFFFFFu;
    }
    scanSampleCounts(sh);   // under the loads just issued
    const uint32_t keyMin = sh.keyMin, fineShift = sh.fineShift;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = warpBase + i * 32u;
        // padding sits at the end of the tile, hence at the end of the last bin; it is neither counted nor stored
        br[i] = (j < tileValid) ? (uint32_t)sh.u.bucketOf[min((key[i] - keyMin) >> fineShift, kDepthFineBins - 1u)] : (uint32_t)(kBkBins - 1);
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) br[i] = (br[i] << 16) | warpRankDigit<9>(br[i], sh.warpHist[warp], lane);
    __syncthreads();

    // thread t: bins 2t, 2t+1 -- exclusive prefix over the warps, tile counts
    uint2 binCount = make_uint2(0u, 0u);
#pragma unroll
    for (int w = 0; w < kBkWarps; ++w) {
        uint2* p = reinterpret_cast<uint2*>(&sh.warpHist[w][bin0]);
        const uint2 c = *p;
        *p = binCount;
        binCount.x += c.x; binCount.y += c.y;
    }
    uint2 validCount = binCount;
    if (tid == kBkThreads - 1) validCount.y -= TILE - tileValid;
    // publish the counts: a word per bin and a RED into the group's sums. Nobody waits inside this kernel: the scatter kernel
    // reads them after the kernel boundary (an in-kernel "wait for every tile" needs every tile's CTA resident at once, which two
    // frames on two streams can deny each other forever)
    {
        const uint32_t group = tile / kBkGroup;
        uint32_t* myStatus = status + (size_t)tile * kBkBins + bin0;
        uint32_t* myGroup = gstatus + (size_t)group * kBkBins + bin0;
        *reinterpret_cast<uint2*>(myStatus) = validCount;
        if (validCount.x) atomicAdd(myGroup, validCount.x);   // RED
        if (validCount.y) atomicAdd(myGroup + 1, validCount.y);
    }
    uint32_t scanTotal;
    const uint32_t pairExcl = blockExclusive(binCount.x + binCount.y, sh.scan, scanTotal);
    sh.binExcl[bin0] = pairExcl;
    sh.binExcl[bin0 + 1] = pairExcl + binCount.x;
    if (tile == 0u && tid == 0) { plan->numBuckets = sh.numBuckets; plan->keyMin = keyMin; plan->shift = fineShift; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = warpBase + i * 32u;
        const uint32_t b = br[i] >> 16;
        const uint32_t p = (br[i] & 0xFFFFu) + sh.binExcl[b] + sh.warpHist[warp][b];  // position inside the tile
        if (j < tileValid) place[base + j] = (b << 16) | p;
    }
}

// Kernel 1b: every count is published (kernel boundary): sum the predecessors' counts and the bucket totals, stage the tile in
// bucket order, write it out. No ballots, no waiting.
template <int ITEMS>
__device__ __forceinline__ void bucketScatterTile(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                  uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut, uint32_t count,
                                                  DepthPlan* __restrict__ plan, const uint32_t* status, const uint32_t* gstatus,
                                                  const uint32_t* __restrict__ place, ScatterShared& sh) {
    constexpr uint32_t TILE = kBkThreads * ITEMS;
    const unsigned tid = threadIdx.x;
    const uint32_t numTiles = (count + TILE - 1u) / TILE;
    const uint32_t tile = blockIdx.x;
    if (tile >= numTiles) return;
    const uint32_t numGroups = (numTiles + kBkGroup - 1u) / kBkGroup;
    const uint32_t bin0 = 2u * tid;
    const uint32_t base = tile * TILE;
    const uint32_t tileValid = min(TILE, count - base);
    const uint32_t group = tile / kBkGroup;
    uint32_t key[ITEMS], val[ITEMS], bp[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = tid + i * kBkThreads;
        if (j < tileValid) { key[i] = keysIn[base + j]; val[i] = valsIn[base + j]; bp[i] = place[base + j]; }
    }
    uint2 exclusive = make_uint2(0u, 0u), totals = make_uint2(0u, 0u);
    const uint2 mine = ld_status32x2(status + (size_t)tile * kBkBins + bin0);   // this tile's counts (padding not included)
    {   // issue every load of a batch before using any (15 tile rows + 16 + 12 group rows at most)
        const uint32_t groupStart = group * kBkGroup, nPred = tile - groupStart;
        uint2 pv[kBkGroup];
#pragma unroll
        for (uint32_t k = 0; k < kBkGroup; ++k)
            pv[k] = k < nPred ? ld_status32x2(status + (size_t)(groupStart + k) * kBkBins + bin0) : make_uint2(0u, 0u);
        for (uint32_t b = 0; b < numGroups; b += 16u) {
            uint2 sv[16];
#pragma unroll
            for (uint32_t k = 0; k < 16u; ++k)
                sv[k] = b + k < numGroups ? ld_status32x2(gstatus + (size_t)(b + k) * kBkBins + bin0) : make_uint2(0u, 0u);
#pragma unroll
            for (uint32_t k = 0; k < 16u; ++k) {
                totals.x += sv[k].x; totals.y += sv[k].y;
                if (b + k < group) { exclusive.x += sv[k].x; exclusive.y += sv[k].y; }
            }
        }
#pragma unroll
        for (uint32_t k = 0; k < kBkGroup; ++k) { exclusive.x += pv[k].x; exclusive.y += pv[k].y; }
    }
    // one scan for both prefixes: bucket offsets over the frame (high 20 bits), bin offsets inside the tile (low 12 bits)
    uint32_t scanTotal;
    const uint32_t packedExcl = blockExclusive(((totals.x + totals.y) << 12) | (mine.x + mine.y), sh.scan, scanTotal);
    const uint32_t bucketStart = packedExcl >> 12, pairExcl = packedExcl & 0xFFFu;
    if (tile == 0u) {   // the local pass reads the bucket offsets from the plan
        plan->bucketStart[bin0] = bucketStart;
        plan->bucketStart[bin0 + 1] = bucketStart + totals.x;
        if (tid == kBkThreads - 1) plan->bucketStart[kBkBins] = count;
    }
    sh.globalBase[bin0] = bucketStart + exclusive.x - pairExcl;
    sh.globalBase[bin0 + 1] = bucketStart + totals.x + exclusive.y - (pairExcl + mine.x);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = tid + i * kBkThreads;
        if (j < tileValid) {
            const uint32_t p = bp[i] & 0xFFFFu;
            sh.u.stage.keys[p] = key[i];
            sh.u.stage.vals[p] = val[i];
            sh.bucket[p] = (unsigned short)(bp[i] >> 16);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = tid + i * kBkThreads;
        if (j < tileValid) {
            const uint32_t dst = sh.globalBase[sh.bucket[j]] + j;
            keysOut[dst] = sh.u.stage.keys[j];
            valsOut[dst] = sh.u.stage.vals[j];
        }
    }
}

__global__ void __launch_bounds__(kBkThreads, 3) bucket_rank_kernel(const uint32_t* __restrict__ keysIn, const uint32_t* countPtr, uint32_t countCap,
                                                                    const uint32_t* __restrict__ fineHist, const KeyRange* keyRange,
                                                                    DepthPlan* __restrict__ plan, uint32_t* status, uint32_t* gstatus,
                                                                    uint32_t* __restrict__ place) {
    __shared__ __align__(16) ScatterShared sh;
    const unsigned tid = threadIdx.x;
    pdlLaunchDependents();
    for (int i = tid; i < kBkWarps * kBkBins; i += kBkThreads) (&sh.warpHist[0][0])[i] = 0;
    pdlWait();
    const uint32_t count = min(ldAfterWait(countPtr), countCap);
    if (count == 0u) {
        if (blockIdx.x == 0 && tid == 0) plan->numBuckets = 0u;
        return;
    }
    __syncthreads();
    if (count <= gridDim.x * (uint32_t)kBkThreads * 8u)
        bucketRankTile<8>(keysIn, count, fineHist, keyRange, plan, status, gstatus, place, sh);
    else   // the host routes frames here only while count <= gridDim.x * 2816 (bucketSortCovers)
        bucketRankTile<kScatterItemsMax>(keysIn, count, fineHist, keyRange, plan, status, gstatus, place, sh);
}

__global__ void __launch_bounds__(kBkThreads, 3) bucket_scatter_kernel(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                       uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                                       const uint32_t* countPtr, uint32_t countCap,
                                                                       DepthPlan* __restrict__ plan, const uint32_t* status, const uint32_t* gstatus,
                                                                       const uint32_t* __restrict__ place) {
    __shared__ __align__(16) ScatterShared sh;
    pdlLaunchDependents();
    pdlWait();
    const uint32_t count = min(ldAfterWait(countPtr), countCap);
    if (count == 0u) return;
    if (count <= gridDim.x * (uint32_t)kBkThreads * 8u)
        bucketScatterTile<8>(keysIn, valsIn, keysOut, valsOut, count, plan, status, gstatus, place, sh);
    else
        bucketScatterTile<kScatterItemsMax>(keysIn, valsIn, keysOut, valsOut, count, plan, status, gstatus, place, sh);
}

// ---------------------------------------------------------------- pass 2: one CTA sorts one bucket
constexpr int kLocalThreads = 256;
constexpr int kLocalWarps = kLocalThreads / 32;
constexpr int kLocalFastBins = 1024;                         // fast path: bins by the top 10 bits of key - lo
constexpr uint32_t kLocalMaxBin = 32;                        // fast path: no bin above this many elements
constexpr int kLocalDB = 8;                                  // skewed path: digit bits per ballot-ranked pass
constexpr int kLocalBins = 1 << kLocalDB;
struct LocalShared {
    uint32_t keysA[kDepthBucketCap];   // raw keys in bucket order (both paths)
    union {
        struct {   // fast path
            uint32_t packedB[kDepthBucketCap];   // bin-ordered ((key - lo) & lowMask) << 12 | position
            uint32_t vals[kDepthBucketCap];      // payloads in bucket order
            uint32_t count[kLocalFastBins], start[kLocalFastBins];
        } f;
        struct {   // skewed path
            uint32_t keysB[kDepthBucketCap];
            unsigned short idxA[kDepthBucketCap], idxB[kDepthBucketCap];
            uint32_t rows[kLocalWarps][kLocalBins];
        } s;
    } u;
    unsigned short rank[kDepthBucketCap];
    uint32_t scan[9];
    uint32_t red[2][kLocalWarps];
};
constexpr size_t kLocalSmemBytes = sizeof(LocalShared);

// ranks of one warp's segment (chunks x 32 elements from segBase, lane included) for the digit (key - lo) >> shift
template <int DB>
__device__ __forceinline__ void localRankSegment(const uint32_t* keys, unsigned short* rankOut, uint32_t lo, uint32_t n,
                                                 uint32_t segBase, uint32_t chunks, uint32_t shift, uint32_t* warpRow, unsigned lane) {
    constexpr uint32_t mask = (1u << DB) - 1u;
    uint32_t c = 0;
    for (; c + 4u <= chunks; c += 4u) {
        uint32_t d[4], r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t i = segBase + (c + k) * 32u;
            d[k] = i < n ? ((keys[i] - lo) >> shift) & mask : mask;   // padding: last digit, last in index order
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = warpRankDigit<DB>(d[k], warpRow, lane);
#pragma unroll
        for (int k = 0; k < 4; ++k) rankOut[segBase + (c + k) * 32u] = (unsigned short)r[k];
    }
    for (; c < chunks; ++c) {
        const uint32_t i = segBase + c * 32u;
        const uint32_t d = i < n ? ((keys[i] - lo) >> shift) & mask : mask;
        rankOut[i] = (unsigned short)warpRankDigit<DB>(d, warpRow, lane);
    }
}

// Skewed bucket: stable 8-bit LSD passes ranked with ballots, any distribution. Warp w owns the contiguous segment
// [w*seg, (w+1)*seg), 32 elements per chunk, in index order; positions >= n are padding. Returns whether the result is in A.
__device__ __noinline__ bool skewedBucketSort(LocalShared& sh, uint32_t n, uint32_t lo, int bits) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t chunks = (n + kLocalThreads - 1u) / kLocalThreads;   // per warp
    const uint32_t segBase = warp * chunks * 32u + lane;
    const int passes = (bits + kLocalDB - 1) / kLocalDB;
    uint32_t* src = sh.keysA; uint32_t* dst = sh.u.s.keysB;
    unsigned short* srcIdx = sh.u.s.idxA; unsigned short* dstIdx = sh.u.s.idxB;
    uint32_t* warpRow = sh.u.s.rows[warp];
    __syncthreads();   // the fast path's bin arrays share the rows
    for (int pass = 0; pass < passes; ++pass) {
        const uint32_t shift = (uint32_t)(pass * kLocalDB);
        for (int i = lane; i < kLocalBins; i += 32) warpRow[i] = 0u;
        __syncwarp();
        localRankSegment<kLocalDB>(src, sh.rank, lo, n, segBase, chunks, shift, warpRow, lane);
        __syncthreads();
        {   // thread d: exclusive prefix over the warps, then over the digits
            uint32_t run = 0u;
#pragma unroll
            for (int w = 0; w < kLocalWarps; ++w) {
                const uint32_t c = sh.u.s.rows[w][tid];
                sh.u.s.rows[w][tid] = run;
                run += c;
            }
            uint32_t total;
            const uint32_t excl = blockExclusive(run, sh.scan, total);
#pragma unroll
            for (int w = 0; w < kLocalWarps; ++w) sh.u.s.rows[w][tid] += excl;
        }
        __syncthreads();
        for (uint32_t c = 0; c < chunks; ++c) {   // every warp moves its own segment
            const uint32_t i = segBase + c * 32u;
            const uint32_t k = src[i];
            const uint32_t d = i < n ? ((k - lo) >> shift) & (kLocalBins - 1u) : (uint32_t)(kLocalBins - 1);
            const uint32_t p = warpRow[d] + sh.rank[i];
            dst[p] = k;
            dstIdx[p] = pass == 0 ? (unsigned short)i : srcIdx[i];
        }
        __syncthreads();
        uint32_t* t = src; src = dst; dst = t;
        unsigned short* ti = srcIdx; srcIdx = dstIdx; dstIdx = ti;
    }
    return src == sh.keysA;
}

// A bucket that does not fit shared memory: 8-bit LSD passes over it through global memory, ping-ponging between the
// bucket's range of the two buffer pairs (both are this CTA's alone). Chunks of 2048 keys are ranked in index order and
// scattered with running digit offsets, so every pass is stable. (bufK1, bufV1) holds the input, (bufK0, bufV0) receives
// the output; zero passes (all keys equal) is a copy.
__device__ __noinline__ void streamingBucketSort(uint32_t* bufK0, uint32_t* bufV0, uint32_t* bufK1, uint32_t* bufV1, uint32_t n, uint32_t lo,
                                                 int bits, LocalShared& sh) {
    constexpr uint32_t CH = kLocalThreads * 8u;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t* s_digitBase = sh.keysA;   // 256 words
    const int passes = bits == 0 ? 0 : (bits + 7) / 8;
    uint32_t* srcK = bufK1; uint32_t* srcV = bufV1; uint32_t* dstK = bufK0; uint32_t* dstV = bufV0;
    __syncthreads();
    for (int pass = 0; pass < passes; ++pass) {
        const uint32_t shift = 8u * (uint32_t)pass;
        s_digitBase[tid] = 0u;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += kLocalThreads) atomicAdd(&s_digitBase[((__ldcg(srcK + i) - lo) >> shift) & 0xFFu], 1u);
        __syncthreads();
        uint32_t total;
        const uint32_t digitExcl = blockExclusive(s_digitBase[tid], sh.scan, total);
        s_digitBase[tid] = digitExcl;
        __syncthreads();
        for (uint32_t c0 = 0; c0 < n; c0 += CH) {
            for (int i = lane; i < 256; i += 32) sh.u.s.rows[warp][i] = 0u;
            __syncwarp();
            uint32_t key[8], val[8], rank[8];
            const uint32_t wb = c0 + warp * 256u + lane;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t j = wb + i * 32u;
                key[i] = j < n ? __ldcg(srcK + j) : 0xFFFFFFFFu;
                val[i] = j < n ? __ldcg(srcV + j) : 0u;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t j = wb + i * 32u;
                const uint32_t d = j < n ? ((key[i] - lo) >> shift) & 0xFFu : 0xFFu;   // padding: last digit, last in index order
                rank[i] = warpRankDigit<8>(d, sh.u.s.rows[warp], lane);
            }
            __syncthreads();
            {   // thread d: exclusive prefix over warps on top of the running digit offset, which advances by the chunk's count
                uint32_t run = s_digitBase[tid];
#pragma unroll
                for (int w = 0; w < kLocalWarps; ++w) {
                    const uint32_t c = sh.u.s.rows[w][tid];
                    sh.u.s.rows[w][tid] = run;
                    run += c;
                }
                s_digitBase[tid] = run;   // padding inflates only digit 0xFF of the LAST chunk: never read again
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t j = wb + i * 32u;
                if (j < n) {
                    const uint32_t p = sh.u.s.rows[warp][((key[i] - lo) >> shift) & 0xFFu] + rank[i];
                    dstK[p] = key[i];
                    dstV[p] = val[i];
                }
            }
            __syncthreads();
        }
        uint32_t* t = srcK; srcK = dstK; dstK = t;
        t = srcV; srcV = dstV; dstV = t;
    }
    if (srcK != bufK0) {   // zero or an even number of passes: the result sits in the input pair
        for (uint32_t i = tid; i < n; i += kLocalThreads) { bufK0[i] = __ldcg(srcK + i); bufV0[i] = __ldcg(srcV + i); }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kLocalThreads, 3) bucket_local_sort_kernel(uint32_t* keysIn, uint32_t* valsIn, uint32_t* keysOut, uint32_t* valsOut,
                                                                             const DepthPlan* __restrict__ plan, KeyRange* keyRange,
                                                                             const uint32_t* __restrict__ gatherSrc, uint32_t* __restrict__ gatherDst) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    LocalShared& sh = *reinterpret_cast<LocalShared*>(s_raw);
    static_assert(kLocalFastBins == 4 * kLocalThreads && kLocalBins == kLocalThreads, "bins per thread");

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    pdlLaunchDependents();
    pdlWait();
    // the key range was consumed by the compaction and scatter kernels (complete now); clear it for the next frame's projection
    if (blockIdx.x == gridDim.x - 1 && tid < 64) (&keyRange->maxKey[0])[tid] = 0u;
    const uint32_t numBuckets = ldAfterWait(&plan->numBuckets);
    for (uint32_t bucket = blockIdx.x; bucket < numBuckets; bucket += gridDim.x) {
        const uint32_t start = ldAfterWait(plan->bucketStart + bucket);
        const uint32_t n = ldAfterWait(plan->bucketStart + bucket + 1) - start;
        if (n == 0u) continue;
        // key range of the bucket: the passes sort key - lo, which has `bits` significant bits
        const bool fits = n <= kDepthBucketCap;
        uint32_t lo = 0xFFFFFFFFu, hi = 0u;
#pragma unroll 4
        for (uint32_t i = tid; i < n; i += kLocalThreads) {
            const uint32_t k = keysIn[start + i];
            if (fits) { sh.keysA[i] = k; sh.u.f.vals[i] = valsIn[start + i]; }
            lo = min(lo, k); hi = max(hi, k);
        }
        for (int i = tid; i < kLocalFastBins; i += kLocalThreads) sh.u.f.count[i] = 0u;
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (lane == 0) { sh.red[0][warp] = lo; sh.red[1][warp] = hi; }
        __syncthreads();
        lo = __reduce_min_sync(0xFFFFFFFFu, sh.red[0][lane & 7u]);
        hi = __reduce_max_sync(0xFFFFFFFFu, sh.red[1][lane & 7u]);
        const int bits = 32 - __clz(hi - lo);   // 0: every key equal -- the order is already the stable one

        if (!fits || bits == 0) {
            streamingBucketSort(keysOut + start, valsOut + start, keysIn + start, valsIn + start, n, lo, bits, sh);
            __syncthreads();
            if (gatherDst)
                for (uint32_t i = tid; i < n; i += kLocalThreads) gatherDst[start + i] = __ldg(gatherSrc + __ldcg(valsOut + start + i));
            __syncthreads();
            continue;
        }
        // ---- fast path: bins by the top 10 bits of key - lo, one shared-memory atomic per element (order inside a bin arbitrary)
        const uint32_t binShift = bits > 10 ? (uint32_t)(bits - 10) : 0u;
        const uint32_t lowMask = (1u << binShift) - 1u;
#pragma unroll 4
        for (uint32_t i = tid; i < n; i += kLocalThreads)
            sh.rank[i] = (unsigned short)atomicAdd(&sh.u.f.count[(sh.keysA[i] - lo) >> binShift], 1u);
        __syncthreads();
        bool skewed;
        {
            const uint4 c = *reinterpret_cast<const uint4*>(&sh.u.f.count[4u * tid]);
            uint32_t total;
            const uint32_t excl = blockExclusive(c.x + c.y + c.z + c.w, sh.scan, total);
            *reinterpret_cast<uint4*>(&sh.u.f.start[4u * tid]) = make_uint4(excl, excl + c.x, excl + c.x + c.y, excl + c.x + c.y + c.z);
            // the packed word keeps 20 bits of the key below the bin bits
            skewed = __syncthreads_or(max(max(c.x, c.y), max(c.z, c.w)) > kLocalMaxBin || binShift > 20u) != 0;
        }
        if (!skewed) {
#pragma unroll 4
            for (uint32_t i = tid; i < n; i += kLocalThreads) {
                const uint32_t rel = sh.keysA[i] - lo;
                sh.u.f.packedB[sh.u.f.start[rel >> binShift] + sh.rank[i]] = ((rel & lowMask) << 12) | i;
            }
            __syncthreads();
            // An element's place = its bin's offset + the number of smaller (key, position) words in the bin. The pair goes
            // straight to its final place in global memory, with the tile count fetched through the payload.
#pragma unroll 2
            for (uint32_t p = tid; p < n; p += kLocalThreads) {
                const uint32_t pk = sh.u.f.packedB[p];
                const uint32_t i = pk & 0xFFFu;
                const uint32_t key = sh.keysA[i];
                const uint32_t val = sh.u.f.vals[i];
                const uint32_t bin = (key - lo) >> binShift;
                const uint32_t b = sh.u.f.start[bin], e = b + sh.u.f.count[bin];
                uint32_t touched = 0u;
                if (gatherDst) touched = __ldg(gatherSrc + val);
                uint32_t smaller = 0u;
                for (uint32_t q = b; q < e; ++q) smaller += sh.u.f.packedB[q] < pk ? 1u : 0u;
                const uint32_t dst = start + b + smaller;
                keysOut[dst] = key;
                valsOut[dst] = val;
                if (gatherDst) gatherDst[dst] = touched;
            }
            __syncthreads();
            continue;
        }
        const bool inA = skewedBucketSort(sh, n, lo, bits);
        const uint32_t* sorted = inA ? sh.keysA : sh.u.s.keysB;
        const unsigned short* sortedIdx = inA ? sh.u.s.idxA : sh.u.s.idxB;
        // positions [0, n) hold (key, original position inside the bucket); the payload is fetched through that position and
        // the tile count through the payload -- four elements per trip, every load of a stage issued before any is used
        for (uint32_t j0 = tid; j0 < n; j0 += 4u * kLocalThreads) {
            uint32_t pv[4], pt[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t j = j0 + c * kLocalThreads;
                if (j < n) pv[c] = valsIn[start + sortedIdx[j]];
            }
            if (gatherDst) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t j = j0 + c * kLocalThreads;
                    if (j < n) pt[c] = __ldg(gatherSrc + pv[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t j = j0 + c * kLocalThreads;
                if (j < n) {
                    keysOut[start + j] = sorted[j];
                    valsOut[start + j] = pv[c];
                    if (gatherDst) gatherDst[start + j] = pt[c];
                }
            }
        }
        __syncthreads();
    }
}

uint32_t bucketScatterGrid(int numSMs) {
    static int blocksPerSM = 0;
    if (blocksPerSM == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, bucket_rank_kernel, kBkThreads, 0);
        if (blocksPerSM < 1) blocksPerSM = 1;
        if (blocksPerSM > 3) blocksPerSM = 3;
    }
    return (uint32_t)numSMs * (uint32_t)blocksPerSM;
}

// Frames whose keys may exceed what one wave of scatter tiles covers (or 512 buckets) stay on the LSD passes.
bool bucketSortCovers(uint32_t maxKeys, int numSMs) {
    return maxKeys <= kDepthBucketMaxGaussians && maxKeys <= bucketScatterGrid(numSMs) * (uint32_t)kBkThreads * (uint32_t)kScatterItemsMax &&
           (bucketScatterGrid(numSMs) + kBkGroup - 1u) / kBkGroup <= 32u;
}

// once per device, before the first frame (the local pass needs more than 48 KB of shared memory)
cudaError_t bucketSortPrepareDevice() {
    return cudaFuncSetAttribute(bucket_local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLocalSmemBytes);
}

cudaError_t launchBucketSort(cudaStream_t s, const BucketSortPlan& p) {
    launchChained(bucket_rank_kernel, bucketScatterGrid(p.numSMs), kBkThreads, s, (const uint32_t*)p.k0, p.countPtr, p.countCap, p.fineHist,
                  (const KeyRange*)p.keyRange, p.plan, p.status, p.gstatus, p.place);
    launchChained(bucket_scatter_kernel, bucketScatterGrid(p.numSMs), kBkThreads, s, (const uint32_t*)p.k0, (const uint32_t*)p.v0, p.k1, p.v1,
                  p.countPtr, p.countCap, p.plan, (const uint32_t*)p.status, (const uint32_t*)p.gstatus, (const uint32_t*)p.place);
    launchChainedSmem(bucket_local_sort_kernel, (int)kDepthMaxBuckets, kLocalThreads, s, kLocalSmemBytes, p.k1, p.v1, p.k0, p.v0,
                      (const DepthPlan*)p.plan, p.keyRange, p.gatherSrc, p.gatherDst);
    return cudaGetLastError();
}

}  // namespace gsm
