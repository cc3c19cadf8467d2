// sort.cu -- stable LSD radix sort of (key, payload) pairs, 8-bit digits, Onesweep style:
// one pass over the keys builds every digit histogram, then each digit pass is ONE kernel that ranks a
// tile with warp match-any, resolves its global bin bases by decoupled look-back (one thread per digit)
// and scatters through shared memory. Persistent CTAs take tiles from an atomic ticket so a tile's
// predecessors are always resident (forward progress) and the grid does not depend on the device-side
// count.
//
// Replaces the reference's 5-kernels-per-pass sorts: depthSort* (DFS.metal:1387-1696, driver
// DepthRadixSortEncoder.swift:139-217) and tileRadix* (DFS.metal:866-1256, driver TileSortEncoder.swift:51-178).
// Semantics kept: stable, ascending, first `count` elements, digit = (key >> 8p) & 0xFF
// (RadixSortHelpers.h:85-88). Stability argument: a tile ranks its elements in index order (warp-major,
// then item, then lane == ascending index) and tiles are prefixed in tile order.
#include "gsm_common.cuh"
#include "gsm_kernels.h"

namespace gsm {

// tools/sort_trace.cu builds this file with GSM_SORT_TRACE to record a per-tile timeline (globaltimer ns at each
// phase boundary, written by thread 0); the product build compiles the hooks out.
#ifdef GSM_SORT_TRACE
__device__ unsigned long long* g_sortTrace = nullptr;  // [tile][16]
__device__ __forceinline__ unsigned long long traceNow() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GSM_TRACE(tile, slot) do { if (threadIdx.x == 0 && g_sortTrace) g_sortTrace[(traceBase + (tile)) * 16 + (slot)] = traceNow(); } while (0)
#define GSM_TRACE_SET(tile, slot, v) do { if (threadIdx.x == 0 && g_sortTrace) g_sortTrace[(traceBase + (tile)) * 16 + (slot)] = (v); } while (0)
#else
#define GSM_TRACE(tile, slot) do { } while (0)
#define GSM_TRACE_SET(tile, slot, v) do { } while (0)
#endif

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr uint32_t kStatusValueMask = 0x3FFFFFFFu;
constexpr uint32_t kStatusAggregate = 0x40000000u;
constexpr uint32_t kStatusInclusive = 0x80000000u;
constexpr int kLookBatch = 8;
constexpr uint32_t kLookGroup = 16;
constexpr uint32_t kFlatMaxTiles = 1024;       // direct summation: <= 15 + 64 predecessor words per digit, 64 arrival masks

// Keys per thread: 8 (2048-key tiles) keeps enough tiles in flight when the whole input is a few hundred
// thousand keys and the pass is latency-bound; 16 (4096-key tiles) halves the per-tile overhead (look-back, scans)
// and lengthens the scatter runs once the input is large (profiles/r1_sort_sweep_*: +24 % at 48 M pairs, -35 % at 709 k).
#ifndef GSM_SORT_CTAS_LARGE
#define GSM_SORT_CTAS_LARGE 3   // resident CTAs per SM asked of the 4096-key u32 instance (register cap 85 / 64 / 51)
#endif
#ifndef GSM_SORT_ITEMS_SMALL
#define GSM_SORT_ITEMS_SMALL 8
#endif
#ifndef GSM_SORT_ITEMS_U16
#define GSM_SORT_ITEMS_U16 16
#endif
uint32_t sortTileSize(int keyBits, bool large) {
    return kSortThreads * (keyBits == 16 ? (uint32_t)GSM_SORT_ITEMS_U16 : (large ? 16u : (uint32_t)GSM_SORT_ITEMS_SMALL));
}

// ---- all digit histograms in one read of the keys
template <typename KeyT, int NPASS, int ITEMS>
__global__ void __launch_bounds__(256) radix_histogram_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ countPtr,
                                                              uint32_t countCap, uint32_t* __restrict__ hist,
                                                              uint32_t* __restrict__ status, uint32_t* __restrict__ gstatus,
                                                              uint32_t tilesCap) {
    __shared__ uint32_t s_hist[NPASS][256];
    pdlLaunchDependents();
    for (int i = threadIdx.x; i < NPASS * 256; i += 256) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    pdlWait();
    const uint32_t count = min(ldAfterWait(countPtr), countCap);
    {   // reset the look-back words this frame's passes will use (sized by the device-side count, not the capacity)
        constexpr uint32_t TILE = kSortThreads * ITEMS;
        const uint32_t words = ((count + TILE - 1) / TILE) * 256u;
        const uint32_t gwords = sortGroupRows(words / 256u) * 256u;
        const uint32_t groupsCap = sortGroupRows(tilesCap);
        for (int p = 0; p < NPASS; ++p) {
            for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < words; i += gridDim.x * 256u)
                status[(size_t)p * tilesCap * 256u + i] = 0u;
            for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < gwords; i += gridDim.x * 256u)
                gstatus[(size_t)p * groupsCap * 256u + i] = 0u;
        }
    }
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < count; i += gridDim.x * 256u) {
        uint32_t k = (uint32_t)keys[i];
#pragma unroll
        for (int p = 0; p < NPASS; ++p) atomicAdd(&s_hist[p][(k >> (8 * p)) & 0xFFu], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * 256; i += 256) {
        uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---- one digit pass
// GATHER (the depth sort's last pass): the payload is an index; besides the sorted pairs the pass writes
// gatherDst[sorted position] = gatherSrc[payload], so the consumer reads that array in order instead of
// gathering through the sorted indices at the head of its own dependency chain (expansion timeline: the scan
// waited ~4.5 us per tile for the slowest in-flight predecessor's random gather).
template <typename KeyT, int ITEMS, bool GATHER>
__global__ void __launch_bounds__(kSortThreads, ((sizeof(KeyT) == 4 && ITEMS == 16) ? GSM_SORT_CTAS_LARGE : 3)) onesweep_pass_kernel(const KeyT* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                     KeyT* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                                     const uint32_t* __restrict__ countPtr, uint32_t countCap,
                                                                     const uint32_t* __restrict__ digitHist, uint32_t* status,
                                                                     uint32_t* gstatus, uint32_t* ticket, int shift,
                                                                     const uint32_t* __restrict__ gatherSrc, uint32_t* __restrict__ gatherDst) {
    constexpr int TILE = kSortThreads * ITEMS;
    constexpr KeyT SENTINEL = (KeyT)~(KeyT)0;

    __shared__ uint32_t s_warpHist[kSortWarps][256];
    __shared__ uint32_t s_binExcl[256];     // tile-local exclusive offset of each digit
    __shared__ uint32_t s_globalBase[256];  // global position of the tile's first element of each digit, minus s_binExcl
    __shared__ uint32_t s_histPrefix[256];  // exclusive prefix of the global digit histogram
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_tile;
    __shared__ KeyT s_keys[TILE];
    __shared__ uint32_t s_vals[TILE];

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
#ifdef GSM_SORT_TRACE
    const unsigned long long traceT0 = traceNow();
#endif
    // Prologue: the count and the digit histogram are two independent round trips; the histogram's prefix scan is only
    // needed after the look-back, so it runs under the first tile's key loads.
    // Nothing global is touched before pdlWait(): frames of >= 65 536 Gaussians zero the ticket inside their projection kernel,
    // after that kernel's own wait, so an atomic issued earlier could still see the previous frame's value. The first ticket is
    // taken right after the wait, together with the count and histogram loads (independent round trips, one latency).
    // Tiles must come from the ticket, not from the block index: a CTA that holds a ticketed tile spins on its predecessors,
    // and a predecessor owned by a not-yet-resident CTA (the grid shares the SMs with the early-launched CTAs of the chain's
    // other kernels) would never run -- measured: stereo frames took seconds with block-index first tiles.
    pdlLaunchDependents();
    for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warpHist[0][0])[i] = 0;
    pdlWait();
    uint32_t firstTicket = 0;
    if (tid == 0) firstTicket = atomicAdd(ticket, 1u);
    const uint32_t digitTotal = ldAfterWait(digitHist + tid);
    const uint32_t count = min(ldAfterWait(countPtr), countCap);
    const uint32_t numTiles = (count + TILE - 1) / TILE;
#ifdef GSM_SORT_TRACE
    const size_t traceBase = (size_t)(shift >> 3) * numTiles;
#endif
    if (tid == 0) s_tile = firstTicket;
    __syncthreads();
    bool firstTile = true;

    while (true) {
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t base = tile * TILE;
        const uint32_t tileValid = min((uint32_t)TILE, count - base);
        GSM_TRACE_SET(tile, 0, traceT0);  // kernel entry of the CTA that took this tile
        GSM_TRACE(tile, 1);               // prologue done, ticket known

        // warp-striped load: element (warp, item, lane) has index base + warp*ITEMS*32 + item*32 + lane
        KeyT key[ITEMS];
        uint32_t rank[ITEMS];
        const uint32_t warpBase = warp * ITEMS * 32u + lane;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint32_t j = warpBase + i * 32u;
            key[i] = (j < tileValid) ? keysIn[base + j] : SENTINEL;
        }
        // payload loads are issued now so their latency hides behind the ranking
        uint32_t val[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint32_t j = warpBase + i * 32u;
            val[i] = (j < tileValid) ? valsIn[base + j] : 0u;
        }
        if (firstTile) {  // uniform; under the loads just issued
            uint32_t total;
            s_histPrefix[tid] = block_exclusive_scan_256(digitTotal, s_scan, total);
            firstTile = false;
        }
        // rank inside the warp, in index order. Peers of a lane = lanes holding the same digit, found with 8
        // ballots (one per digit bit; independent across items, so they pipeline) -- MATCH.ANY made this loop
        // latency-bound (ncu r1_v3: 39 % short-scoreboard stalls on its result). The lowest peer bumps the warp's
        // private counter with one shared-memory atomic and broadcasts the old value.
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = ((uint32_t)key[i] >> shift) & 0xFFu;
            unsigned peers = 0xFFFFFFFFu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const bool bit = (d >> b) & 1u;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
                peers &= bit ? bal : ~bal;
            }
            const uint32_t lower = __popc(peers & ((1u << lane) - 1u));
            uint32_t pre = 0;
            if (lower == 0) pre = atomicAdd(&s_warpHist[warp][d], (uint32_t)__popc(peers));
            pre = __shfl_sync(0xFFFFFFFFu, pre, __ffs(peers) - 1);
            rank[i] = pre + lower;
        }
        GSM_TRACE(tile, 2);  // thread 0's warp ranked (keys arrived)
        __syncthreads();
        GSM_TRACE(tile, 3);  // all warps ranked

        // thread d: exclusive prefix over warps, tile count of digit d
        uint32_t binCount = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            uint32_t c = s_warpHist[w][tid];
            s_warpHist[w][tid] = binCount;
            binCount += c;
        }
        // padding lanes carry the sentinel and sit at the END of the tile, hence at the end of their bin;
        // they are not counted globally and never stored.
        const uint32_t sentinelDigit = ((uint32_t)SENTINEL >> shift) & 0xFFu;
        uint32_t validCount = binCount - ((tid == sentinelDigit) ? (TILE - tileValid) : 0u);

        uint32_t* myStatus = status + (size_t)tile * 256u + tid;
        uint32_t exclusive = 0;
        const uint32_t group = tile / kLookGroup;
        uint32_t* myGroup = gstatus + (size_t)group * 256u + tid;
        if (numTiles <= kFlatMaxTiles) {
            // Few tiles (every frame-sized sort): all of them are in flight at once, nothing is "inclusive" yet, and a
            // chained look-back is a string of dependent L2 round trips. Instead a tile publishes its counts once -- a
            // word per digit and a RED into its group's per-digit sums -- and sets its bit in the group's ARRIVAL MASK;
            // one warp polls the masks of the groups up to its own (one or two words per lane), and only when every
            // predecessor has arrived do the 256 digit threads sum the published words, once. Polling the per-digit
            // words themselves costs 256 threads x up to 79 loads per attempt: the pollers then starve the CTAs of their SM
            // that are still ranking (at 5 CTAs per SM the tile sort took 1.1 ms instead of 76 us).
            const uint32_t numGroups = (numTiles + kLookGroup - 1u) / kLookGroup;
            uint32_t* arriveMask = gstatus + (size_t)numGroups * 256u;   // one word per group, after the sum rows
            st_status32(myStatus, validCount);
            if ((group + 1u) * kLookGroup < numTiles) atomicAdd(myGroup, validCount);  // RED: result unused
            __threadfence();   // the counts are visible before the arrival bit
            __syncthreads();
            if (tid == 0) atomicOr(arriveMask + group, 1u << (tile % kLookGroup));
            if (tid < 32u) {
                const uint32_t mine = (1u << (tile % kLookGroup)) - 1u;   // earlier tiles of the own group
                bool ok;
                do {
                    ok = true;
                    for (uint32_t g = lane; g <= group; g += 32u) {
                        const uint32_t m = ld_status32(arriveMask + g);
                        const uint32_t need = g < group ? 0xFFFFu : mine;
                        ok = ok && (m & need) == need;
                    }
                } while (!__all_sync(0xFFFFFFFFu, ok));
                __threadfence();   // the arrival bits were read before the counts are
            }
            __syncthreads();
            const uint32_t groupStart = group * kLookGroup, nPred = tile - groupStart;
            for (uint32_t b = 0; b < group; b += kLookBatch) {
                uint32_t sv[kLookBatch];
#pragma unroll
                for (int k = 0; k < kLookBatch; ++k)
                    sv[k] = (b + k < group) ? ld_status32(gstatus + (size_t)(b + k) * 256u + tid) : 0u;
#pragma unroll
                for (int k = 0; k < kLookBatch; ++k) exclusive += sv[k];
            }
            for (uint32_t b = 0; b < nPred; b += kLookBatch) {
                uint32_t sv[kLookBatch];
#pragma unroll
                for (int k = 0; k < kLookBatch; ++k)
                    sv[k] = (b + k < nPred) ? ld_status32(status + (size_t)(groupStart + b + k) * 256u + tid) : 0u;
#pragma unroll
                for (int k = 0; k < kLookBatch; ++k) exclusive += sv[k];
            }
        } else {
        // Many tiles (streaming): two-level decoupled look-back, one thread per digit. Level 1 walks the tiles of the
        // own group of kLookGroup tiles; level 2 walks per-GROUP words published by each group's last tile. In the
        // steady state the predecessor is already inclusive and the walk is one load.
        const bool groupLeader = (tile % kLookGroup) == kLookGroup - 1;
        if (tile == 0) {
            st_status32(myStatus, kStatusInclusive | validCount);
            if (groupLeader) st_status32(myGroup, kStatusInclusive | validCount);
        } else {
            st_status32(myStatus, kStatusAggregate | validCount);
            bool done = false;
            {   // level 1: predecessors inside the group
                const int groupStart = (int)(group * kLookGroup);
                int look = (int)tile - 1;
                if (look >= groupStart) {  // streaming steady state: the predecessor is usually already inclusive -- one load
                    const uint32_t sw = ld_status32(status + (size_t)look * 256u + tid);
                    if (sw & kStatusInclusive) { exclusive += sw & kStatusValueMask; done = true; }
                    else if (sw & kStatusAggregate) { exclusive += sw & kStatusValueMask; look--; }
                }
                while (!done && look >= groupStart) {
                    uint32_t sv[kLookBatch];
#pragma unroll
                    for (int k = 0; k < kLookBatch; ++k) {
                        const int t = look - k;
                        sv[k] = (t >= groupStart) ? ld_status32(status + (size_t)t * 256u + tid) : 0u;
                    }
                    int consumed = 0;
#pragma unroll
                    for (int k = 0; k < kLookBatch; ++k) {
                        if (!done && consumed == k && look - k >= groupStart) {  // only a contiguous run of published words
                            const uint32_t sw = sv[k];
                            if (sw & kStatusInclusive) { exclusive += sw & kStatusValueMask; done = true; }
                            else if (sw & kStatusAggregate) { exclusive += sw & kStatusValueMask; consumed++; }
                        }
                    }
                    look -= consumed;
                }
            }
            if (!done && group > 0) {
                // the group's last tile now knows the group aggregate: publish it before walking on
                if (groupLeader) st_status32(myGroup, kStatusAggregate | (exclusive + validCount));
                int look = (int)group - 1;  // level 2: earlier groups
                while (!done && look >= 0) {
                    uint32_t sv[kLookBatch];
#pragma unroll
                    for (int k = 0; k < kLookBatch; ++k) {
                        const int g = look - k;
                        sv[k] = (g >= 0) ? ld_status32(gstatus + (size_t)g * 256u + tid) : 0u;
                    }
                    int consumed = 0;
#pragma unroll
                    for (int k = 0; k < kLookBatch; ++k) {
                        if (!done && consumed == k && look - k >= 0) {
                            const uint32_t sw = sv[k];
                            if (sw & kStatusInclusive) { exclusive += sw & kStatusValueMask; done = true; }
                            else if (sw & kStatusAggregate) { exclusive += sw & kStatusValueMask; consumed++; }
                        }
                    }
                    look -= consumed;
                }
            }
            st_status32(myStatus, kStatusInclusive | (exclusive + validCount));
            if (groupLeader) st_status32(myGroup, kStatusInclusive | (exclusive + validCount));
        }
        }
        GSM_TRACE(tile, 4);  // digit 0's prefix known
        uint32_t total;
        uint32_t binExcl = block_exclusive_scan_256(binCount, s_scan, total);
        s_binExcl[tid] = binExcl;
        s_globalBase[tid] = s_histPrefix[tid] + exclusive - binExcl;
        __syncthreads();
        GSM_TRACE(tile, 5);  // every digit's prefix known

        // scatter keys into tile order
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint32_t d = ((uint32_t)key[i] >> shift) & 0xFFu;
            rank[i] += s_binExcl[d] + s_warpHist[warp][d];  // position inside the tile
            s_keys[rank[i]] = key[i];
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) s_vals[rank[i]] = val[i];
        __syncthreads();
        GSM_TRACE(tile, 6);  // tile ordered in shared memory
        // valid elements occupy tile positions [0, tileValid) except that sentinel padding sits at the end of
        // the sentinel digit's bin; bins after it (none: the sentinel digit is 0xFF) would shift.
        uint32_t gdst[GATHER ? ITEMS : 1], gval[GATHER ? ITEMS : 1];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint32_t j = tid + i * kSortThreads;
            if (j < tileValid) {
                KeyT k = s_keys[j];
                uint32_t d = ((uint32_t)k >> shift) & 0xFFu;
                uint32_t dst = s_globalBase[d] + j;
                const uint32_t v = s_vals[j];
                keysOut[dst] = k;
                valsOut[dst] = v;
                if (GATHER) { gdst[i] = dst; gval[i] = __ldg(gatherSrc + v); }
            }
        }
        if (GATHER) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                if (tid + i * kSortThreads < tileValid) gatherDst[gdst[i]] = gval[i];
        }
        __syncthreads();
        GSM_TRACE(tile, 7);  // stores issued
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warpHist[0][0])[i] = 0;
        __syncthreads();
    }
}

// ---- the MSD tile sort's pass (tilesort.cu) on frames of a few million instances: TWO consecutive tiles per CTA, ONE exchange.
// A 16-bit pass over 2.9 M pairs is 710 tiles on 444 resident CTAs: two waves of a latency chain that costs ~17 us however few keys
// it moves (ticket, loads, 128 ballots, publish, wait for every predecessor, scatter). Here a CTA takes a UNIT of two consecutive
// tiles, ranks them one after the other into two shared-memory staging buffers (48 KB), publishes the unit's per-digit counts once,
// waits once, and writes both tiles out: one wave, one wait. Everything else -- ranking in index order, direct summation of the
// predecessors' counts behind arrival masks for few units, two-level look-back for many, tickets after the dependent-launch wait --
// is onesweep_pass_kernel's, with "tile" read as "unit" in the status words.
template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads, 3) onesweep_pair_kernel(const unsigned short* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                        unsigned short* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                                        const uint32_t* __restrict__ countPtr, uint32_t countCap,
                                                                        const uint32_t* __restrict__ digitHist, uint32_t* status,
                                                                        uint32_t* gstatus, uint32_t* ticket, int shift) {
    typedef unsigned short KeyT;
    constexpr int TILE = kSortThreads * ITEMS;
    constexpr uint32_t SENTINEL = 0xFFFFu;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_dyn);                      // [2][TILE]
    KeyT* s_keys = reinterpret_cast<KeyT*>(s_dyn + 2 * TILE * sizeof(uint32_t));  // [2][TILE]
    __shared__ uint32_t s_warpHist[kSortWarps][256];
    __shared__ uint32_t s_binExcl[2][256];     // tile-local exclusive offset of each digit, per tile of the unit
    __shared__ uint32_t s_globalBase[2][256];  // global position of the tile's first element of each digit, minus s_binExcl
    __shared__ uint32_t s_histPrefix[256];
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_unit;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    pdlLaunchDependents();
    for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warpHist[0][0])[i] = 0;
    pdlWait();
    uint32_t firstTicket = 0;
    if (tid == 0) firstTicket = atomicAdd(ticket, 1u);
    const uint32_t digitTotal = ldAfterWait(digitHist + tid);
    const uint32_t count = min(ldAfterWait(countPtr), countCap);
    const uint32_t numTiles = (count + TILE - 1) / TILE;
    // two tiles per unit only where that makes ONE wave out of two: more tiles than resident CTAs, at most twice as many. Fewer
    // tiles are one wave already; with more, a CTA's second unit would queue behind a whole pair (stereo C4, 977 tiles: 91 -> 105 us)
    const uint32_t tpu = (numTiles > gridDim.x && (numTiles + 1u) / 2u <= gridDim.x) ? 2u : 1u;
    const uint32_t numUnits = (numTiles + tpu - 1u) / tpu;
    if (tid == 0) s_unit = firstTicket;
    __syncthreads();
    bool firstTile = true;
    const uint32_t sentinelDigit = (SENTINEL >> shift) & 0xFFu;

    while (true) {
        const uint32_t unit = s_unit;
        if (unit >= numUnits) break;
        uint32_t vc0 = 0u, vc1 = 0u;   // thread d: valid keys of digit d in the unit's first / second tile
#pragma unroll 1
        for (uint32_t k = 0; k < 2u; ++k) {
            const uint32_t tile = unit * tpu + k;
            if (k >= tpu || tile >= numTiles) { s_binExcl[1][tid] = 0u; break; }   // uniform: one tile per unit, or the last unit of an odd tile count
            const uint32_t base = tile * TILE;
            const uint32_t tileValid = min((uint32_t)TILE, count - base);
            uint32_t key[ITEMS], rank[ITEMS], val[ITEMS];
            const uint32_t warpBase = warp * ITEMS * 32u + lane;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t j = warpBase + i * 32u;
                key[i] = (j < tileValid) ? (uint32_t)keysIn[base + j] : SENTINEL;
            }
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t j = warpBase + i * 32u;
                val[i] = (j < tileValid) ? valsIn[base + j] : 0u;
            }
            if (firstTile) {  // uniform; under the loads just issued
                uint32_t total;
                s_histPrefix[tid] = block_exclusive_scan_256(digitTotal, s_scan, total);
                firstTile = false;
            }
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t d = (key[i] >> shift) & 0xFFu;
                unsigned peers = 0xFFFFFFFFu;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const bool bit = (d >> b) & 1u;
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, bit);
                    peers &= bit ? bal : ~bal;
                }
                const uint32_t lower = __popc(peers & ((1u << lane) - 1u));
                uint32_t pre = 0;
                if (lower == 0) pre = atomicAdd(&s_warpHist[warp][d], (uint32_t)__popc(peers));
                pre = __shfl_sync(0xFFFFFFFFu, pre, __ffs(peers) - 1);
                rank[i] = pre + lower;
            }
            __syncthreads();
            uint32_t binCount = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t c = s_warpHist[w][tid];
                s_warpHist[w][tid] = binCount;
                binCount += c;
            }
            const uint32_t validCount = binCount - ((tid == sentinelDigit) ? (TILE - tileValid) : 0u);
            if (k == 0u) vc0 = validCount; else vc1 = validCount;
            uint32_t total;
            const uint32_t binExcl = block_exclusive_scan_256(binCount, s_scan, total);
            s_binExcl[k][tid] = binExcl;
            __syncthreads();
            KeyT* sk = s_keys + k * TILE;
            uint32_t* sv = s_vals + k * TILE;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t d = (key[i] >> shift) & 0xFFu;
                const uint32_t pos = rank[i] + s_binExcl[k][d] + s_warpHist[warp][d];  // position inside the tile
                sk[pos] = (KeyT)key[i];
                sv[pos] = val[i];
            }
            __syncthreads();
            for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&s_warpHist[0][0])[i] = 0;
            __syncthreads();
        }

        const uint32_t validCount = vc0 + vc1;
        uint32_t* myStatus = status + (size_t)unit * 256u + tid;
        uint32_t exclusive = 0;
        const uint32_t group = unit / kLookGroup;
        uint32_t* myGroup = gstatus + (size_t)group * 256u + tid;
        if (numUnits <= kFlatMaxTiles) {
            const uint32_t numGroups = (numUnits + kLookGroup - 1u) / kLookGroup;
            uint32_t* arriveMask = gstatus + (size_t)numGroups * 256u;
            st_status32(myStatus, validCount);
            if ((group + 1u) * kLookGroup < numUnits) atomicAdd(myGroup, validCount);  // RED: result unused
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicOr(arriveMask + group, 1u << (unit % kLookGroup));
            if (tid < 32u) {
                const uint32_t mine = (1u << (unit % kLookGroup)) - 1u;
                bool ok;
                do {
                    ok = true;
                    for (uint32_t g = lane; g <= group; g += 32u) {
                        const uint32_t m = ld_status32(arriveMask + g);
                        const uint32_t need = g < group ? 0xFFFFu : mine;
                        ok = ok && (m & need) == need;
                    }
                } while (!__all_sync(0xFFFFFFFFu, ok));
                __threadfence();
            }
            __syncthreads();
            const uint32_t groupStart = group * kLookGroup, nPred = unit - groupStart;
            for (uint32_t b = 0; b < group; b += kLookBatch) {
                uint32_t sv[kLookBatch];
#pragma unroll
                for (int q = 0; q < kLookBatch; ++q)
                    sv[q] = (b + q < group) ? ld_status32(gstatus + (size_t)(b + q) * 256u + tid) : 0u;
#pragma unroll
                for (int q = 0; q < kLookBatch; ++q) exclusive += sv[q];
            }
            for (uint32_t b = 0; b < nPred; b += kLookBatch) {
                uint32_t sv[kLookBatch];
#pragma unroll
                for (int q = 0; q < kLookBatch; ++q)
                    sv[q] = (b + q < nPred) ? ld_status32(status + (size_t)(groupStart + b + q) * 256u + tid) : 0u;
#pragma unroll
                for (int q = 0; q < kLookBatch; ++q) exclusive += sv[q];
            }
        } else {
            // Many units (streaming sizes): decoupled look-back over the units, one thread per digit, eight status words per round
            // trip. In the steady state the predecessor is already inclusive and the walk is one load. (The two-level form of
            // onesweep_pass_kernel measured slower here: C3 tile sort 339 us against 326 us.)
            if (unit == 0) {
                st_status32(myStatus, kStatusInclusive | validCount);
            } else {
                st_status32(myStatus, kStatusAggregate | validCount);
                bool done = false;
                int look = (int)unit - 1;
                {
                    const uint32_t sw = ld_status32(status + (size_t)look * 256u + tid);
                    if (sw & kStatusInclusive) { exclusive += sw & kStatusValueMask; done = true; }
                    else if (sw & kStatusAggregate) { exclusive += sw & kStatusValueMask; look--; }
                }
                while (!done) {
                    uint32_t sv[kLookBatch];
#pragma unroll
                    for (int q = 0; q < kLookBatch; ++q) {
                        const int t = look - q;
                        sv[q] = (t >= 0) ? ld_status32(status + (size_t)t * 256u + tid) : 0u;
                    }
                    int consumed = 0;
#pragma unroll
                    for (int q = 0; q < kLookBatch; ++q) {
                        if (!done && consumed == q && look - q >= 0) {  // only a contiguous run of published words
                            const uint32_t sw = sv[q];
                            if (sw & kStatusInclusive) { exclusive += sw & kStatusValueMask; done = true; }
                            else if (sw & kStatusAggregate) { exclusive += sw & kStatusValueMask; consumed++; }
                        }
                    }
                    look -= consumed;
                }
                st_status32(myStatus, kStatusInclusive | (exclusive + validCount));
            }
        }
        s_globalBase[0][tid] = s_histPrefix[tid] + exclusive - s_binExcl[0][tid];
        s_globalBase[1][tid] = s_histPrefix[tid] + exclusive + vc0 - s_binExcl[1][tid];
        __syncthreads();
#pragma unroll 1
        for (uint32_t k = 0; k < tpu; ++k) {
            const uint32_t tile = unit * tpu + k;
            if (tile >= numTiles) break;
            const uint32_t tileValid = min((uint32_t)TILE, count - tile * TILE);
            const KeyT* sk = s_keys + k * TILE;
            const uint32_t* sv = s_vals + k * TILE;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t j = tid + i * kSortThreads;
                if (j < tileValid) {
                    const uint32_t kk = sk[j];
                    const uint32_t dst = s_globalBase[k][(kk >> shift) & 0xFFu] + j;
                    keysOut[dst] = (KeyT)kk;
                    valsOut[dst] = sv[j];
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(ticket, 1u);
        __syncthreads();
    }
}

template <typename KeyT, int ITEMS>
static cudaError_t runSort(cudaStream_t s, const SortPlan& p) {
    if (p.numPasses <= 0) return cudaSuccess;
    const int gridHist = p.numSMs * 4;
    KeyT* k0 = (KeyT*)p.k0;
    KeyT* k1 = (KeyT*)p.k1;
    if (!p.histogramReady) switch (p.numPasses) {
        case 1: launchChained(radix_histogram_kernel<KeyT, 1, ITEMS>, gridHist, 256, s, k0, p.countPtr, p.countCap, p.hist, p.status, p.gstatus, p.tilesCap); break;
        case 2: launchChained(radix_histogram_kernel<KeyT, 2, ITEMS>, gridHist, 256, s, k0, p.countPtr, p.countCap, p.hist, p.status, p.gstatus, p.tilesCap); break;
        case 3: launchChained(radix_histogram_kernel<KeyT, 3, ITEMS>, gridHist, 256, s, k0, p.countPtr, p.countCap, p.hist, p.status, p.gstatus, p.tilesCap); break;
        default: launchChained(radix_histogram_kernel<KeyT, 4, ITEMS>, gridHist, 256, s, k0, p.countPtr, p.countCap, p.hist, p.status, p.gstatus, p.tilesCap); break;
    }
    if (sizeof(KeyT) == 2 && p.pairTiles && p.numPasses == 1 && p.gatherSrc == nullptr) {   // the MSD tile sort's pass: two tiles per CTA, one exchange
        constexpr size_t smem = 2u * (size_t)kSortThreads * ITEMS * (sizeof(uint32_t) + sizeof(unsigned short));
        static int pairBlocksPerSM = 0;
        if (pairBlocksPerSM == 0) {
            cudaError_t e = cudaFuncSetAttribute(onesweep_pair_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pairBlocksPerSM, onesweep_pair_kernel<ITEMS>, kSortThreads, smem);
            if (pairBlocksPerSM < 1) pairBlocksPerSM = 1;
        }
        launchChainedSmem(onesweep_pair_kernel<ITEMS>, p.numSMs * pairBlocksPerSM, kSortThreads, s, smem, (const unsigned short*)k0, (const uint32_t*)p.v0,
                          (unsigned short*)k1, p.v1, p.countPtr, p.countCap, (const uint32_t*)p.hist, p.status, p.gstatus, p.tickets, p.shift0);
        return cudaGetLastError();
    }
    static int blocksPerSM = 0;  // per template instance
    if (blocksPerSM == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSM, onesweep_pass_kernel<KeyT, ITEMS, false>, kSortThreads, 0);
        if (blocksPerSM < 1) blocksPerSM = 1;
    }
    const int grid = p.numSMs * blocksPerSM;  // every CTA is resident: the look-back cannot starve
    for (int pass = 0; pass < p.numPasses; ++pass) {
        const bool even = (pass & 1) == 0;
        const bool gather = p.gatherSrc != nullptr && pass == p.numPasses - 1;
        auto kernel = gather ? onesweep_pass_kernel<KeyT, ITEMS, true> : onesweep_pass_kernel<KeyT, ITEMS, false>;
        launchChained(kernel, grid, kSortThreads, s,
                      even ? k0 : k1, even ? p.v0 : p.v1, even ? k1 : k0, even ? p.v1 : p.v0, p.countPtr, p.countCap,
                      p.hist + 256 * pass, p.status + (size_t)pass * p.tilesCap * 256u,
                      p.gstatus + (size_t)pass * sortGroupRows(p.tilesCap) * 256u, p.tickets + pass, p.shift0 + 8 * pass,
                      p.gatherSrc, p.gatherDst);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if ((p.numPasses & 1) && !p.leaveInScratch) {  // odd pass count: result is in the scratch pair, copy back (TileSortEncoder.swift:170-177)
        e = cudaMemcpyAsync(p.k0, p.k1, (size_t)p.countCap * sizeof(KeyT), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(p.v0, p.v1, (size_t)p.countCap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
    }
    return e;
}

cudaError_t launchSort(cudaStream_t s, const SortPlan& p) {
    if (p.keyBits == 16) return runSort<uint16_t, GSM_SORT_ITEMS_U16>(s, p);
    return p.largeTiles ? runSort<uint32_t, 16>(s, p) : runSort<uint32_t, GSM_SORT_ITEMS_SMALL>(s, p);
}

}  // namespace gsm
