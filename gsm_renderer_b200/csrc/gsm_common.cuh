// gsm_common.cuh -- shared device-side declarations of the DepthFirst path (internal; the public
// boundary is include/gsm/gsm.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/gsm/gsm.h"
#include "../../include/gsm/gsm_types.h"

namespace gsm {

// ---- programmatic dependent launch: the frame is a chain of ~12 short kernels, each a few microseconds of fixed
// latency (launch, CTA ramp, first round trips). Every kernel of the chain is launched with the
// programmatic-stream-serialization attribute, lets its successor start early (pdlLaunchDependents) and
// touches global memory only after pdlWait(), which returns once the whole predecessor grid has completed and
// its writes are visible. What runs before the wait (shared-memory clears, ticket atomics on words zeroed by the
// frame's memset) overlaps the predecessor's tail. Every CTA must execute pdlWait() so that "this grid completed"
// implies "its predecessor completed" down the chain.
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdlLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// A scalar the predecessor kernel produced (a device-side count, a header field): read it with this after pdlWait(), never
// through a `const T* __restrict__` pointer -- nvcc treats such loads as invariant and may hoist them ABOVE the inline-asm wait
// (SASS: LDG.E.CONSTANT before ACQBULK; found when the depth sort's local pass read a bucket count of zero). A volatile asm
// load cannot move across the volatile asm wait. tools/check_pdl_hoist.py and tests/test_abi.py scan the built library for it.
__device__ __forceinline__ uint32_t ldAfterWait(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

bool pdlEnabled();  // capi.cu: GSM_PDL=0 in the environment turns the attribute off (A/B measurement)

template <typename... KArgs, typename... Args>
inline cudaError_t launchChainedSmem(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, size_t smemBytes, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdlEnabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launchChained(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdlEnabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kTile = 16;               // DFR.swift:8-9
constexpr float kAlphaThreshold = 0.005f;   // GlobalRenderer.swift:66
constexpr float kTotalInkThreshold = 2.0f;  // GlobalRenderer.swift:67
constexpr uint32_t kRadixAlignment = 1024;  // radixBlockSize*radixGrainSize, DFR.swift:40-41

// Per-frame device state, zeroed by one memset at the start of every frame (replaces
// resetDepthFirstStateKernel DFS.metal:1372 + the ClearCounters blit DFR.swift:267-272).
struct FrameState {
    uint32_t visibleCountRaw;    // written by the last project tile (visibilityScatterCompact, DFS.metal:618-620)
    uint32_t totalInstancesRaw;  // atomic sum of nTouched (DFS.metal:218)
    uint32_t activeTileCount;    // DFS.metal:1310
    uint32_t ticketProject;      // dynamic tile ids => look-back forward progress
    uint32_t ticketScan;
    uint32_t ticketSort[8];      // depth passes 0-3, tile passes 4-7
    uint32_t rangesDone;         // CTAs of the tile-range kernel that finished their boundary scan
    uint32_t ticketRoute;        // group.cu: tiles of the record-routing kernel
    uint32_t routeDone;          // group.cu: CTAs of the routing kernel that have issued all their (remote) stores
    uint32_t hist[8][256];       // global digit histograms: depth passes 0-3, tile passes 4-7
    uint32_t routeTotals[8];     // group.cu: records routed to each destination rank this frame
    uint32_t ticketBlend;        // tiles of the persistent mono blend
    uint32_t ingestDone;         // group.cu: CTAs of the ingest kernel that have read all their records
    uint32_t recordTotal;        // group.cu: records received from all sources this frame (device-side N of the compaction)
    uint32_t compactDone;        // CTAs of the compaction kernel that have issued all their stores (the last one plans the depth sort)
    uint32_t ticketBucket;       // sort.cu: tiles of the depth sort's bucket-scatter pass
    uint32_t ticketLocal;        // sort.cu: buckets of the depth sort's local pass
    uint32_t _pad2[2];
    uint32_t fineHist[8192];     // depth keys per fine bin ((key - keyMin) >> shift), counted by the compaction kernel (kDepthFineBins)
};

// ---- depth sort as ONE global pass + one shared-memory pass (bucketsort.cu: bucket_scatter_kernel, bucket_local_sort_kernel),
// used for frames of up to kDepthBucketMaxGaussians Gaussians (host-side choice; larger frames run the LSD passes of sort.cu).
// The projection kernels record the frame's key range (KeyRange: outside the per-frame zero region because the projection
// kernel itself clears that region; reset by the local-sort kernel). The compaction kernel counts a sample of the keys per fine
// bin of that range (FrameState::fineHist); the scatter kernel turns the sample into bucket boundaries and sums the exact
// bucket offsets (DepthPlan).
constexpr uint32_t kDepthFineBins = 8192;
constexpr uint32_t kDepthMaxBuckets = 512;
constexpr uint32_t kDepthBucketTarget = 2048;   // expected keys per bucket
constexpr uint32_t kDepthBucketCap = 4096;      // what one CTA of the local pass sorts in shared memory (larger buckets stream)
constexpr uint32_t kDepthSampleStride = 8;
constexpr uint32_t kDepthBucketMaxGaussians = 1000000;   // => at most 1e6 / 8 / 256 + 1 = 489 buckets, 326 scatter tiles of 3072 keys
struct KeyRange {
    uint32_t maxKey[32];      // atomicMax of key        (slot = warp tile & 31: same-address REDs serialise)
    uint32_t maxInvKey[32];   // atomicMax of ~key  => min key = ~max; all zero = nothing recorded
};
struct DepthPlan {            // written by the scatter kernel's first tile, read by the local pass
    uint32_t numBuckets;      // 0: nothing to sort
    uint32_t keyMin, shift, _pad;                 // the fine bins' origin and width (diagnostic, gsm_debug_read)
    uint32_t bucketStart[kDepthMaxBuckets + 4];   // exclusive offsets, [kDepthMaxBuckets] = key count
};
__device__ __forceinline__ void recordKeyRange(KeyRange* kr, uint32_t key, bool valid, uint32_t slot) {  // whole warp
    if (!kr) return;
    const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, valid ? key : 0u);
    const uint32_t lo = __reduce_max_sync(0xFFFFFFFFu, valid ? ~key : 0u);
    if ((threadIdx.x & 31u) == 0u && (hi | lo) != 0u) {
        atomicMax(&kr->maxKey[slot & 31u], hi);
        atomicMax(&kr->maxInvKey[slot & 31u], lo);
    }
}

// What the blend stage reads per splat: the quantised record pre-expanded once per visible Gaussian
// (conic from conicFromThetaSigmas GaussianShared.h:490-510 rounded to half exactly as
// DFS.metal:1753-1764 does per thread). 32 bytes = two 128-bit loads.
struct __align__(16) BlendSplat {
    __half2 mean;      // meanX, meanY
    __half2 cxx_cyy;   // half(conic.x), half(conic.z)
    __half2 cxy2_op;   // half(2*conic.y), half(opacity_u8)/255h
    __half2 rg;        // colour R,G as half(u8)/255h
    __half2 b_depth;   // colour B, depth
    uint32_t valid;    // 1
    uint32_t _pad[2];
};
static_assert(sizeof(BlendSplat) == 32, "BlendSplat is 32 bytes");

// What travels between ranks in the strip-sharded frame (strip.cu, group.cu): one compacted, projected splat.
// Records stay in ascending global gid order per source, so the stable depth sort breaks ties as on one GPU.
struct __align__(16) SplatRecord {
    uint4 renderData;
    int4 bounds;
    uint32_t key, maskLo, maskHi, gid;  // mask: stage 1's hit bits of the first 64 AABB tiles (gsm_tiletest.cuh)
};
static_assert(sizeof(SplatRecord) == GSM_SPLAT_RECORD_BYTES, "record size");

// system-scope flag traffic between GPUs (mailboxes in peer memory, group.cu)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Camera constants as the kernels take them (by value, __grid_constant__).
struct MonoCam {
    float view[16];
    float proj[16];
    float center[3];
    float width, height, nearPlane, farPlane;
    uint32_t shComponents, gaussianCount;
    float inputIsSRGB;
    uint32_t tilesX, tilesY;
};
struct StereoCam {
    float leftView[16], leftProj[16], leftCenter[3];
    float rightView[16], rightProj[16], rightCenter[3];
    float sceneTransform[16];
    float width, height, nearPlane, farPlane;
    uint32_t shComponents, gaussianCount;
    float inputIsSRGB;
    uint32_t tilesX, tilesY;
};

// 64-bit look-back word for single-value chained scans: flag in the high half, value in the low half.
constexpr uint32_t kFlagAggregate = 1u, kFlagInclusive = 2u;

__device__ __forceinline__ void st_status64(unsigned long long* p, uint32_t flag, uint32_t value) {
    unsigned long long v = ((unsigned long long)flag << 32) | value;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_u64_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_status64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_status32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Eager variant for software-pipelined consumers: a tile PUBLISHES as soon as its aggregate is known (a word per tile
// and a RED into its group's accumulator, whose top bits count arrivals), goes on with other work, and RESOLVES its
// exclusive prefix later -- by then every predecessor has long published, so the resolve is one round trip and no tile
// ever depends on another tile's resolve (no leader). Group word: [63:40] arrivals, [39:0] sum of the group's aggregates.
constexpr int kEagerShift = 40;
__device__ __forceinline__ void prefixPublish(unsigned long long* tileWords, unsigned long long* groupWords, uint32_t tile,
                                              uint32_t aggregate) {  // one thread
    st_u64_relaxed(tileWords + tile, (1ull << 62) | (unsigned long long)aggregate);
    atomicAdd(groupWords + (tile >> 5), (1ull << kEagerShift) | (unsigned long long)aggregate);
}
__device__ __forceinline__ uint32_t prefixResolve(const unsigned long long* tileWords, const unsigned long long* groupWords,
                                                  uint32_t tile) {  // one full warp; sum modulo 2^32
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t g = tile >> 5, nT = tile & 31u;
    const unsigned long long fullGroup = 32ull << kEagerShift;
    uint32_t sum = 0;
    for (uint32_t base = 0; base == 0 || base < g; base += 128u) {
        while (true) {
            unsigned long long w1 = 1ull << 62, w2[4] = {fullGroup, fullGroup, fullGroup, fullGroup};
            if (base == 0 && lane < nT) w1 = ld_status64(tileWords + (g << 5) + lane);
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                if (base + k * 32u + lane < g) w2[k] = ld_status64(groupWords + base + k * 32u + lane);
            bool ok = (w1 >> 62) != 0ull;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) ok = ok && (w2[k] >> kEagerShift) == 32ull;
            if (!__all_sync(0xFFFFFFFFu, ok)) continue;
            uint32_t v = (uint32_t)w1;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) v += (uint32_t)w2[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            sum += v;
            break;
        }
    }
    return sum;
}

// Dual-sum flavour of the eager prefix (visibility compaction): every word carries a visible count (18 bits) and a
// touched-tile sum (40 bits). Tile word: [63] published. Group word: [63:58] arrivals (32 tiles per group).
// Same-address REDs cost ~0.7 ns each on B200 (tools/micro/atomic_rate.cu): 32 arrivals per word are free.
__device__ __forceinline__ void prefixPublish2(unsigned long long* tileWords, unsigned long long* groupWords, uint32_t tile,
                                               uint32_t visible, uint32_t touched) {  // one thread
    const unsigned long long v = ((unsigned long long)visible << 40) | (unsigned long long)touched;
    st_u64_relaxed(tileWords + tile, (1ull << 63) | v);
    atomicAdd(groupWords + (tile >> 5), (1ull << 58) | v);
}
__device__ __forceinline__ void prefixResolve2(const unsigned long long* tileWords, const unsigned long long* groupWords,
                                               uint32_t tile, uint32_t& exclVisible, unsigned long long& exclTouched) {  // one full warp
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t g = tile >> 5, nT = tile & 31u;
    const unsigned long long fullGroup = 32ull << 58, touchedMask = (1ull << 40) - 1ull;
    uint32_t vis = 0;
    unsigned long long tch = 0;
    for (uint32_t base = 0; base == 0 || base < g; base += 128u) {
        while (true) {
            unsigned long long w1 = 1ull << 63, w2[4] = {fullGroup, fullGroup, fullGroup, fullGroup};
            if (base == 0 && lane < nT) w1 = ld_status64(tileWords + (g << 5) + lane);
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                if (base + k * 32u + lane < g) w2[k] = ld_status64(groupWords + base + k * 32u + lane);
            bool ok = (w1 >> 63) != 0ull;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) ok = ok && (w2[k] >> 58) == 32ull;
            if (!__all_sync(0xFFFFFFFFu, ok)) continue;
            uint32_t v = (uint32_t)(w1 >> 40) & 0x3FFFFu;
            unsigned long long t = w1 & touchedMask;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) { v += (uint32_t)(w2[k] >> 40) & 0x3FFFFu; t += w2[k] & touchedMask; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xFFFFFFFFu, v, o); t += __shfl_xor_sync(0xFFFFFFFFu, t, o); }
            vis += v;
            tch += t;
            break;
        }
    }
    exclVisible = vis;
    exclTouched = tch;
}

// exclusive scan of one value per thread over a 256-thread block; returns the exclusive prefix and
// writes the block total. smem: 9 words.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* smem9, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) smem9[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < 8) ? smem9[lane] : 0u;
        uint32_t winc = w;
        for (int o = 1; o < 8; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= (unsigned)o) winc += t;
        }
        if (lane < 8) smem9[lane] = winc - w;
        if (lane == 7) smem9[8] = winc;
    }
    __syncthreads();
    uint32_t r = smem9[warp] + inc - v;
    total = smem9[8];
    __syncthreads();
    return r;
}

}  // namespace gsm
