// capi.cu -- the C ABI of include/gsm/gsm.h: renderer handle, device arena, stage sequencing on the caller's
// stream, white-box reads. The stage order is the reference's encodeRender / encodeStereoPipeline
// (DepthFirstRenderer.swift:237-465, :595-831) with the fusions listed in DESIGN.md section 5.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cstdlib>
#include "gsm_common.cuh"
#include "gsm_kernels.h"

using namespace gsm;

namespace {

thread_local std::string g_lastError;

gsm_status fail(gsm_status s, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_lastError = buf;
    return s;
}

}  // namespace

namespace gsm {
gsm_status reportFailure(gsm_status s, const char* what, cudaError_t e) { return fail(s, what, e); }  // scene.cu

// Standalone stable pair sort on the current device (gsm_sort_pairs, the Morton pre-sort of scene.cu): scratch is
// allocated per call and the stream is synchronised before it is freed.
struct SortScratchLayout { size_t oHist, oTickets, oStatus, oGStatus, oK1, oV1, total; uint32_t tiles; bool large; };
static SortScratchLayout sortScratchLayout(uint32_t count, int keyBits, int numPasses) {
    SortScratchLayout L;
    L.large = keyBits == 32 && count >= 3000000u;
    const uint32_t tile = sortTileSize(keyBits, L.large);
    L.tiles = (count + tile - 1) / tile;
    const size_t keyBytes = (size_t)count * (keyBits / 8);
    auto up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    // scratch: [count u32][hist 4*256][tickets 4][status passes*tiles*256][k1][v1]
    L.oHist = 256; L.oTickets = L.oHist + 4 * 256 * 4; L.oStatus = up(L.oTickets + 16, 256);
    L.oGStatus = up(L.oStatus + (size_t)numPasses * L.tiles * 256 * 4, 256);
    L.oK1 = up(L.oGStatus + (size_t)numPasses * sortGroupRows(L.tiles) * 256 * 4, 256);
    L.oV1 = up(L.oK1 + keyBytes, 256);
    L.total = up(L.oV1 + (size_t)count * 4, 256);
    return L;
}
size_t sortScratchBytes(uint32_t count, int keyBits, int numPasses) { return sortScratchLayout(count, keyBits, numPasses).total; }

gsm_status sortPairsStandalone(cudaStream_t s, int numSMs, void* keys, void* payload, uint32_t count, int keyBits, int numPasses,
                               void* callerScratch) {
    if (count == 0) return GSM_OK;
    const SortScratchLayout L = sortScratchLayout(count, keyBits, numPasses);
    const bool large = L.large;
    const uint32_t tiles = L.tiles;
    const size_t oHist = L.oHist, oTickets = L.oTickets, oStatus = L.oStatus, oGStatus = L.oGStatus, oK1 = L.oK1, oV1 = L.oV1;
    char* scratch = (char*)callerScratch;
    cudaError_t e = cudaSuccess;
    if (!scratch) {
        e = cudaMalloc((void**)&scratch, L.total);
        if (e != cudaSuccess) return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "sort scratch", e);
    }
    gsm_status st = GSM_OK;
    do {
        if ((e = cudaMemsetAsync(scratch, 0, oK1, s)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(scratch, &count, 4, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
        SortPlan p;
        p.k0 = keys; p.k1 = scratch + oK1; p.v0 = (uint32_t*)payload; p.v1 = (uint32_t*)(scratch + oV1);
        p.countPtr = (const uint32_t*)scratch; p.countCap = count;
        p.hist = (uint32_t*)(scratch + oHist); p.status = (uint32_t*)(scratch + oStatus); p.gstatus = (uint32_t*)(scratch + oGStatus); p.tickets = (uint32_t*)(scratch + oTickets);
        p.tilesCap = tiles; p.keyBits = keyBits; p.numPasses = numPasses; p.numSMs = numSMs; p.histogramReady = false; p.largeTiles = large;
        if ((e = launchSort(s, p)) != cudaSuccess) break;
        if (!callerScratch) e = cudaStreamSynchronize(s);  // the scratch is freed below; a caller's scratch outlives the call
    } while (false);
    if (e != cudaSuccess) st = fail(GSM_ERR_RENDER_FAILED, "sort pairs", e);
    if (!callerScratch) cudaFree(scratch);
    return st;
}
}

namespace {

#define GSM_CUDA(call, what)                                            \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail(GSM_ERR_RENDER_FAILED, what, e__); \
    } while (0)

constexpr uint32_t kMaxSupportedGaussians = 30000000u;  // DFR.swift:7

size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Resources {
    char* arena = nullptr;
    size_t bytes = 0;
    bool stereo = false;
    uint32_t maxGaussians = 0, maxInstances = 0, maxTiles = 0;
    uint32_t depthTilesCap = 0, tileTilesCap = 0;
    uint32_t frameGaussians = 0;  // gaussianCount of the frame being encoded (host-known bound on V; picks the sort tile size)
    // per Gaussian
    void* renderData = nullptr;
    int32_t* bounds = nullptr;
    uint32_t* nTouched = nullptr;
    uint2* hitMask = nullptr;
    BlendSplat* blendSplats = nullptr;
    // per visible (ping-pong pairs)
    uint32_t* depthKeys[2] = {nullptr, nullptr};
    int32_t* primIdx[2] = {nullptr, nullptr};
    uint32_t* offsets = nullptr;
    // per instance
    void* tileIds[2] = {nullptr, nullptr};
    int32_t* instIdx[2] = {nullptr, nullptr};
    // per tile
    uint32_t* lowerBounds = nullptr;
    GSMGaussianHeader* tileHeaders = nullptr;
    uint32_t* activeTiles = nullptr;
    GSMDepthFirstHeader* header = nullptr;
    // depth sort as bucket scatter + local sort (bucketsort.cu): key range recorded by the projection (cleared by its consumer),
    // plan written by the compaction's last CTA. Not part of the per-frame zero region.
    KeyRange* keyRange = nullptr;
    DepthPlan* depthPlan = nullptr;
    // zeroed every frame
    FrameState* fs = nullptr;
    unsigned long long* projStatus = nullptr;
    unsigned long long* scanStatus = nullptr;
    unsigned long long* scanGroups = nullptr;
    unsigned long long* projGroups = nullptr;
    size_t zeroBytes = 0;
    // zeroed by the sort's histogram kernel
    uint32_t* depthSortStatus = nullptr;
    uint32_t* tileSortStatus = nullptr;
    uint32_t* depthSortGStatus = nullptr;
    uint32_t* tileSortGStatus = nullptr;
};

}  // namespace

struct gsm_renderer {
    gsm_config cfg;
    int device = 0;
    int numSMs = 148;
    Resources mono, stereoRes;
    bool profiling = false;
    cudaEvent_t ev[GSM_NUM_STAGES + 1] = {};
    bool evValid = false;
    bool evRecorded = false;
    float stageMs[GSM_NUM_STAGES] = {};
    double lastMs = -1.0;
    // gsm_render_host staging
    cudaStream_t hostStream = nullptr;
    void* stGaussians = nullptr; size_t stGaussiansBytes = 0;
    void* stHarmonics = nullptr; size_t stHarmonicsBytes = 0;
    void* stColor = nullptr; size_t stColorBytes = 0;
    void* stDepth = nullptr; size_t stDepthBytes = 0;
    uint32_t lastTilesX = 0, lastTilesY = 0;
    bool lastStereo = false;
    // StereoRenderTarget.foveated: intermediate image (ensureStereoIntermediateColor, DFR.swift:139-164), rate-map tables, sRGB table
    void* stereoIntermediate = nullptr; size_t stereoIntermediateBytes = 0;
    float* rateTables = nullptr; size_t rateTableFloats = 0;
    uint8_t* srgbLut = nullptr;
    unsigned short* expTable = nullptr;  // exact exp(-0.5h * p) over the non-negative halfs, staged into shared memory by the mono blend
    // GlobalRenderer path (gsm_render_global): its own arena, allocated on the first global frame
    char* globalArena = nullptr;
    gsm::GlobalFrame globalFrame = {};
    char* globalSortScratch = nullptr;   // sort state + ping-pong buffers (sortScratchLayout over 4 * maxGaussians pairs)
    bool lastGlobal = false;
};

namespace gsm {
bool pdlEnabled() {
    static const bool on = [] { const char* e = getenv("GSM_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
}  // namespace gsm

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// DepthFirstViewResources / StereoTiledResources (DepthFirstResources.swift:58-325, :377-615): one arena.
gsm_status ensureResources(gsm_renderer* r, Resources& res, bool stereo) {
    if (res.arena) return GSM_OK;
    const gsm_config& c = r->cfg;
    const uint32_t G = c.maxGaussians < 1 ? 1 : c.maxGaussians;
    const uint32_t I = 4u * G;  // DepthFirstResources.swift:80
    const uint32_t tilesX = (c.maxWidth + kTile - 1) / kTile, tilesY = (c.maxHeight + kTile - 1) / kTile;
    const uint32_t T = tilesX * tilesY < 1 ? 1 : tilesX * tilesY;
    const bool tile16 = c.tileIdPrecision == GSM_KEY_BITS16;
    const size_t tileIdBytes = tile16 ? 2 : 4;
    res.stereo = stereo;
    res.maxGaussians = G; res.maxInstances = I; res.maxTiles = T;
    res.depthTilesCap = (G + sortTileSize(32, false) - 1) / sortTileSize(32, false);  // sized for the smaller tile
    res.tileTilesCap = (I + sortTileSize(tile16 ? 16 : 32, false) - 1) / sortTileSize(tile16 ? 16 : 32, false);

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = alignUp(off + bytes, 256); return o; };
    const size_t oRender = take((size_t)G * (stereo ? 32 : 16));
    const size_t oBounds = take((size_t)G * 16);
    const size_t oTouched = take((size_t)G * 4);
    const size_t oHitMask = stereo ? 0 : take((size_t)G * 8);
    const size_t oBlend = stereo ? 0 : take((size_t)G * sizeof(BlendSplat));
    const size_t oKeys0 = take((size_t)G * 4), oKeys1 = take((size_t)G * 4);
    const size_t oIdx0 = take((size_t)G * 4), oIdx1 = take((size_t)G * 4);
    const size_t oOffsets = take((size_t)G * 4);
    const size_t oTid0 = take((size_t)I * tileIdBytes), oTid1 = take((size_t)I * tileIdBytes);
    const size_t oIi0 = take((size_t)I * 4), oIi1 = take((size_t)I * 4);
    const size_t oLB = take((size_t)(T + 1) * 4);
    const size_t oTH = take((size_t)T * 8);
    const size_t oAT = take((size_t)T * 4);
    const size_t oHeader = take(sizeof(GSMDepthFirstHeader));
    const size_t oKeyRange = take(sizeof(KeyRange));
    const size_t oDepthPlan = take(sizeof(DepthPlan));
    const size_t oZero = off;
    const size_t oFS = take(sizeof(FrameState));
    const size_t oProjStatus = take(((size_t)(G + 31) / 32 + 8) * 8);  // one prefix word per 32-gid warp tile (strip ingest) / 2048-gid tile (compaction)
    const size_t oProjGroups = take(((size_t)(G + 31) / 32 / 32 + 8) * 8);  // one word per group of 32 tiles (prefixTwoLevel)
    const size_t oScanStatus = take(((size_t)(G + 255) / 256 + 8) * 8);  // one prefix word per 256-Gaussian expansion tile
    const size_t oScanGroups = take(((size_t)(G + 255) / 256 / 32 + 8) * 8);
    // the sorts' per-tile status words are part of the per-frame memset too (a few MB: cheaper than a reset kernel's launch)
    const size_t oDepthStatus = take((size_t)4 * res.depthTilesCap * 256 * 4);
    const size_t oTileStatus = take(((size_t)(tile16 ? 2 : 4) * res.tileTilesCap + 258) * 256 * 4);  // + 256 rows: the MSD tile sort numbers its chunks per bucket (tilesort.cu)
    const size_t oDepthGStatus = take((size_t)4 * sortGroupRows(res.depthTilesCap) * 256 * 4);
    const size_t oTileGStatus = take((size_t)(tile16 ? 2 : 4) * sortGroupRows(res.tileTilesCap) * 256 * 4);
    const size_t zeroEnd = off;
    res.bytes = off;

    cudaError_t e = cudaMalloc((void**)&res.arena, res.bytes);
    if (e != cudaSuccess) {
        res.arena = nullptr;
        char msg[128];
        snprintf(msg, sizeof msg, "failed to allocate device arena of %zu bytes", res.bytes);
        return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, msg, e);
    }
    char* a = res.arena;
    res.renderData = a + oRender;
    res.bounds = (int32_t*)(a + oBounds);
    res.nTouched = (uint32_t*)(a + oTouched);
    res.hitMask = stereo ? nullptr : (uint2*)(a + oHitMask);
    res.blendSplats = stereo ? nullptr : (BlendSplat*)(a + oBlend);
    res.depthKeys[0] = (uint32_t*)(a + oKeys0); res.depthKeys[1] = (uint32_t*)(a + oKeys1);
    res.primIdx[0] = (int32_t*)(a + oIdx0); res.primIdx[1] = (int32_t*)(a + oIdx1);
    res.offsets = (uint32_t*)(a + oOffsets);
    res.tileIds[0] = a + oTid0; res.tileIds[1] = a + oTid1;
    res.instIdx[0] = (int32_t*)(a + oIi0); res.instIdx[1] = (int32_t*)(a + oIi1);
    res.lowerBounds = (uint32_t*)(a + oLB);
    res.tileHeaders = (GSMGaussianHeader*)(a + oTH);
    res.activeTiles = (uint32_t*)(a + oAT);
    res.header = (GSMDepthFirstHeader*)(a + oHeader);
    res.keyRange = (KeyRange*)(a + oKeyRange);
    res.depthPlan = (DepthPlan*)(a + oDepthPlan);
    res.fs = (FrameState*)(a + oFS);
    res.projStatus = (unsigned long long*)(a + oProjStatus);
    res.scanStatus = (unsigned long long*)(a + oScanStatus);
    res.scanGroups = (unsigned long long*)(a + oScanGroups);
    res.projGroups = (unsigned long long*)(a + oProjGroups);
    res.zeroBytes = zeroEnd - oZero;
    res.depthSortStatus = (uint32_t*)(a + oDepthStatus);
    res.tileSortStatus = (uint32_t*)(a + oTileStatus);
    res.depthSortGStatus = (uint32_t*)(a + oDepthGStatus);
    res.tileSortGStatus = (uint32_t*)(a + oTileGStatus);
    e = cudaMemset(res.arena, 0, res.bytes);
    if (e != cudaSuccess) return fail(GSM_ERR_RENDER_FAILED, "arena memset", e);
    e = bucketSortPrepareDevice();
    if (e != cudaSuccess) return fail(GSM_ERR_FAILED_TO_CREATE_PIPELINE, "depth sort local pass: shared memory opt-in", e);
    return GSM_OK;
}

void freeResources(Resources& res) {
    if (res.arena) cudaFree(res.arena);
    res = Resources();
}

bool largeSort(uint32_t frameGaussians) { return frameGaussians >= 3000000u; }  // see sort.cu: sortTileSize

// GSM_TILE_MSD=0 in the environment keeps the tile sort on the LSD passes + the range kernel (A/B measurement)
bool tileSortMsdEnabled() {
    static const bool on = [] { const char* e = getenv("GSM_TILE_MSD"); return !(e && e[0] == '0'); }();
    return on;
}
// Frames of at least this many Gaussians keep the two LSD passes (GSM_TILE_MSD_MAX in the environment: A/B measurement). No limit
// by default: most-significant-digit first also wins on the large frames (C5 view, 3 M Gaussians: tile sort + ranges 123 -> 99 us;
// C3, 6 M at 4K, 18.6 M instances: 364 -> 329 us; profiles/r2_tile_msd_large_ab.txt) -- the range kernel disappears there too.
uint32_t tileSortMsdMaxGaussians() {
    static const uint32_t v = [] { const char* e = getenv("GSM_TILE_MSD_MAX"); return e ? (uint32_t)strtoul(e, nullptr, 10) : 0xFFFFFFFFu; }();
    return v;
}
// GSM_TILE_PAIRS=0 in the environment keeps the MSD pass on one tile per CTA (A/B measurement)
bool tilePairsEnabled() {
    static const bool on = [] { const char* e = getenv("GSM_TILE_PAIRS"); return !(e && e[0] == '0'); }();
    return on;
}
// GSM_DEPTH_BUCKETS=0 in the environment keeps the depth sort on the four LSD passes (A/B measurement)
bool depthBucketsEnabled() {
    static const bool on = [] { const char* e = getenv("GSM_DEPTH_BUCKETS"); return !(e && e[0] == '0'); }();
    return on;
}
// The bucket path runs when the projection kernel records the key range of exactly the keys the compaction stores -- 32-bit
// depth keys (16-bit keys are re-encoded inside the compaction) of a frame projected on this GPU -- and the frame is small
// enough for one wave of scatter tiles (bucketsort.cu).
void attachDepthPlan(const gsm_renderer* r, Resources& res, ProjectOut& po, uint32_t gaussianCount) {
    if (depthBucketsEnabled() && r->cfg.depthSortKeyPrecision != GSM_KEY_BITS16 && bucketSortCovers(gaussianCount, r->numSMs)) {
        po.keyRange = res.keyRange;
        po.depthPasses = 0u;   // the bucket sort builds its own counts: no LSD digit histograms from the compaction (0.7 M x 4 shared-memory atomics + the per-CTA flush)
    }
}

int tileSortPasses(uint32_t tileCount) {  // TileSortEncoder.swift:61-62
    uint32_t v = tileCount > 0 ? (tileCount - 1 > 1 ? tileCount - 1 : 1) : 1;
    int bits = 0;
    while (v) { bits++; v >>= 1; }
    if (tileCount == 0) bits = 1;
    return (bits + 7) / 8;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

void recordStage(gsm_renderer* r, cudaStream_t s, int idx) {
    if (r->profiling && r->evValid) cudaEventRecord(r->ev[idx], s);
}

// stages 2-7 shared by the mono and stereo frames (DFR.swift:325-430, :683-787)
gsm_status encodeSortExpandRange(gsm_renderer* r, Resources& res, cudaStream_t s, bool stereo, uint32_t tilesX, uint32_t tilesY,
                                 bool depthHistReady = true, uint32_t tileRowFirst = 0u, uint32_t tileRowCount = 0xFFFFFFFFu,
                                 bool depthPlanned = false) {  // depthHistReady false: a histogram kernel runs first (no producer filled hist[0..3]); depthPlanned: the compaction wrote res.depthPlan
    const gsm_config& c = r->cfg;
    const bool tile16 = c.tileIdPrecision == GSM_KEY_BITS16;
    const bool key16 = c.depthSortKeyPrecision == GSM_KEY_BITS16;
    // stage 2: depth sort (4 passes; 2 with 16-bit depth keys -- the keys stay u32, DepthRadixSortEncoder.swift:45-46)
    SortPlan dp;
    dp.k0 = res.depthKeys[0]; dp.k1 = res.depthKeys[1];
    dp.v0 = (uint32_t*)res.primIdx[0]; dp.v1 = (uint32_t*)res.primIdx[1];
    dp.countPtr = &res.header->visibleCount; dp.countCap = res.maxGaussians;
    dp.gatherSrc = res.nTouched; dp.gatherDst = res.offsets;  // stage 3 rides on the last pass: offsets[i] = nTouched[sortedIdx[i]]
    dp.hist = &res.fs->hist[0][0]; dp.status = res.depthSortStatus; dp.gstatus = res.depthSortGStatus; dp.tickets = &res.fs->ticketSort[0];
    dp.largeTiles = largeSort(res.frameGaussians);
    dp.tilesCap = res.depthTilesCap; dp.keyBits = 32; dp.numPasses = key16 ? 2 : 4; dp.numSMs = r->numSMs;
    dp.histogramReady = depthHistReady;  // compact_visible_kernel filled hist[0..3]; the status words were cleared with the frame state
    if (depthPlanned) {   // one global pass + one shared-memory pass instead of the four LSD passes
        BucketSortPlan bp;
        bp.k0 = res.depthKeys[0]; bp.k1 = res.depthKeys[1]; bp.v0 = dp.v0; bp.v1 = dp.v1;
        bp.countPtr = dp.countPtr; bp.countCap = dp.countCap; bp.plan = res.depthPlan; bp.fineHist = res.fs->fineHist; bp.keyRange = res.keyRange;
        bp.status = res.depthSortStatus; bp.gstatus = res.depthSortGStatus;
        bp.gatherSrc = dp.gatherSrc; bp.gatherDst = dp.gatherDst; bp.numSMs = r->numSMs;
        bp.place = res.offsets;   // free until the local pass writes the depth-ordered tile counts into it
        GSM_CUDA(launchBucketSort(s, bp), "depth sort (buckets)");
    } else {
        GSM_CUDA(launchSort(s, dp), "depth sort");
    }
    recordStage(r, s, 2);
    // stages 3+4
    const int tilePasses = tileSortPasses(tilesX * tilesY);
    recordStage(r, s, 3);
    // 16-bit ids of 9..16 bits on frames small enough for the direct-summation prefix: one pass on the high byte + a local
    // pass per bucket that also writes the tile ranges (tilesort.cu); else the LSD passes + the range kernel
    const uint32_t lowBits = tileSortLowBits(tilesX * tilesY);
    const bool msdTiles = tileSortMsdEnabled() && tile16 && tilePasses == 2 && lowBits > 0u && res.frameGaussians < tileSortMsdMaxGaussians();
    // stage 5
    GSM_CUDA(launchCreateInstances(s, stereo, tile16, res.primIdx[0], res.offsets, res.hitMask, res.offsets, res.scanStatus, res.scanGroups, &res.fs->ticketScan, res.bounds, res.renderData, res.tileIds[0],
                                   res.instIdx[0], res.header, tilesX, res.maxInstances, res.maxGaussians, &res.fs->hist[4][0],
                                   (uint32_t)tilePasses, r->numSMs, msdTiles ? lowBits : 0xFFFFFFFFu), "create instances");
    recordStage(r, s, 4);
    // stage 6
    SortPlan tp;
    tp.k0 = res.tileIds[0]; tp.k1 = res.tileIds[1];
    tp.v0 = (uint32_t*)res.instIdx[0]; tp.v1 = (uint32_t*)res.instIdx[1];
    tp.countPtr = &res.header->totalInstances; tp.countCap = res.maxInstances;
    tp.hist = &res.fs->hist[4][0]; tp.status = res.tileSortStatus; tp.gstatus = res.tileSortGStatus; tp.tickets = &res.fs->ticketSort[4];
    tp.largeTiles = largeSort(res.frameGaussians);
    tp.tilesCap = res.tileTilesCap; tp.keyBits = tile16 ? 16 : 32; tp.numPasses = tilePasses;
    tp.numSMs = r->numSMs; tp.histogramReady = true;  // create_instances_kernel filled hist[4..7]; the status words were cleared with the frame state
    if (msdTiles) { tp.numPasses = 1; tp.shift0 = (int)lowBits; tp.leaveInScratch = true; tp.pairTiles = tilePairsEnabled(); }
    GSM_CUDA(launchSort(s, tp), "tile sort");
    if (msdTiles)
        GSM_CUDA(launchTileLocalSort(s, res.tileIds[1], (const uint32_t*)res.instIdx[1], res.tileIds[0], (uint32_t*)res.instIdx[0],
                                     &res.fs->hist[4][0], res.header, res.maxInstances, lowBits, tilesX * tilesY, res.lowerBounds,
                                     res.tileSortStatus + (size_t)res.tileTilesCap * 256u, r->numSMs),
                 "tile sort (local pass)");
    recordStage(r, s, 5);
    // stage 7
    if (!msdTiles) {
        const uint32_t rowEnd = tileRowCount > tilesY - tileRowFirst ? tilesY : tileRowFirst + tileRowCount;
        GSM_CUDA(launchTileRanges(s, tile16, res.tileIds[0], res.header, tilesX * tilesY, res.lowerBounds, res.maxInstances,
                                  tileRowFirst * tilesX, rowEnd * tilesX), "tile ranges");
    }
    recordStage(r, s, 6);
    return GSM_OK;
}

gsm_status validateFrame(gsm_renderer* r, uint32_t width, uint32_t height, const void* gaussians, const void* harmonics,
                         const void* color) {
    if (!r) return fail(GSM_ERR_INVALID_ARGUMENT, "null renderer");
    if (!gaussians || !harmonics || !color) return fail(GSM_ERR_INVALID_ARGUMENT, "null buffer");
    if (width == 0 || height == 0 || width > r->cfg.maxWidth || height > r->cfg.maxHeight)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "dimensions exceed RendererConfig.maxWidth/maxHeight");
    if (!aligned16(gaussians) || !aligned16(harmonics)) return fail(GSM_ERR_INVALID_ARGUMENT, "input buffers must be 16-byte aligned");
    if (((uintptr_t)color & 7u) != 0) return fail(GSM_ERR_INVALID_ARGUMENT, "colour target must be 8-byte aligned");
    return GSM_OK;
}

void fillMonoCam(MonoCam& mc, const gsm_camera* cam, uint32_t count, uint32_t sh, uint32_t w, uint32_t h, bool srgb) {
    memcpy(mc.view, cam->viewMatrix, 64);
    memcpy(mc.proj, cam->projectionMatrix, 64);
    memcpy(mc.center, cam->position, 12);
    mc.width = (float)w; mc.height = (float)h;
    mc.nearPlane = cam->nearPlane; mc.farPlane = cam->farPlane;
    mc.shComponents = sh; mc.gaussianCount = count;
    mc.inputIsSRGB = srgb ? 1.0f : 0.0f;
    mc.tilesX = (w + kTile - 1) / kTile; mc.tilesY = (h + kTile - 1) / kTile;
}


// ---- strip-sharded frame: what the ingest kernels (strip.cu, group.cu) write, and the stages after them
ProjectOut stripIngestOutputs(gsm_renderer* r, Resources& res) {
    ProjectOut po;
    po.fs = res.fs; po.status = res.projStatus; po.statusGroups = res.projGroups; po.renderData = res.renderData; po.bounds = res.bounds;
    po.nTouched = res.nTouched; po.hitMask = res.hitMask; po.blendSplats = res.blendSplats; po.depthKeys = res.depthKeys[0];
    po.primitiveIndices = res.primIdx[0]; po.maxOut = res.maxGaussians; po.header = res.header; po.maxInstances = res.maxInstances; po.preDepthKeys = res.depthKeys[1];
    po.depthHist = &res.fs->hist[0][0]; po.depthPasses = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 2u : 4u;
    po.depthKey16 = 0; po.gidFirst = 0;
    return po;
}

// Order-preserving compaction of the ingested records (per-record count, key, gid live in three arrays that are free until
// the depth sort runs), then the same stages 2-8 as the single-GPU frame, blending only tile rows [rowFirst, rowFirst+rowCount).
// recordCap: host-side bound on the record count; the exact count is read from po.countPtr on the device when that is set.
gsm_status encodeStripTail(gsm_renderer* r, Resources& res, cudaStream_t s, ProjectOut po, uint32_t recordCap, void* color, void* depth,
                           uint32_t width, uint32_t height, uint32_t tileRowFirst, uint32_t tileRowCount) {
    const uint32_t tilesX = (width + kTile - 1) / kTile, tilesY = (height + kTile - 1) / kTile;
    po.recTouched = res.offsets; po.recKey = res.depthKeys[1]; po.recGid = (uint32_t*)res.primIdx[1];
    GSM_CUDA(launchCompactVisible(s, recordCap, po, r->numSMs), "record compaction");  // also writes the header
    gsm_status st = encodeSortExpandRange(r, res, s, false, tilesX, tilesY, true, tileRowFirst, tileRowCount);  // same stages 2-7 as the single-GPU frame
    if (st != GSM_OK) return st;
    GSM_CUDA(launchBlendMono(s, res.lowerBounds, res.blendSplats, res.instIdx[0], width, height, tilesX, tilesY, tileRowFirst,
                             tileRowCount, (__half*)color, (__half*)depth, TileOut{res.tileHeaders, res.activeTiles, &res.fs->activeTileCount},
                             r->expTable, &res.fs->ticketBlend, r->numSMs), "strip blend");
    return GSM_OK;
}


// Stage 1 + compaction of the gid range [gidFirst, gidFirst+gidCount) of a sharded frame: per-gid outputs keyed by the GLOBAL
// gid, compacted (key, gid) pairs in ascending gid order in depthKeys[0] / primIdx[0], count in fs->visibleCountRaw.
gsm_status encodeShardProject(gsm_renderer* r, Resources& res, cudaStream_t s, const void* gaussians, const void* harmonics,
                              uint32_t gidFirst, uint32_t gidCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width,
                              uint32_t height) {
    res.frameGaussians = gidCount;
    r->lastStereo = false;
    GSM_CUDA(cudaMemsetAsync(res.fs, 0, res.zeroBytes, s), "frame-state memset");
    MonoCam mc;
    fillMonoCam(mc, camera, gidCount, shComponents, width, height, r->cfg.gaussianColorSpace == GSM_COLORSPACE_SRGB);
    ProjectOut po;
    po.fs = res.fs; po.status = res.projStatus; po.statusGroups = res.projGroups; po.renderData = res.renderData; po.bounds = res.bounds;
    po.nTouched = res.nTouched; po.hitMask = res.hitMask; po.blendSplats = nullptr; po.depthKeys = res.depthKeys[0];
    po.primitiveIndices = res.primIdx[0]; po.maxOut = res.maxGaussians; po.header = res.header; po.maxInstances = res.maxInstances; po.preDepthKeys = res.depthKeys[1];
    po.depthHist = &res.fs->hist[0][0]; po.depthPasses = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 2u : 4u;
    po.depthKey16 = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 1u : 0u; po.gidFirst = gidFirst;
    GSM_CUDA(launchProjectMono(s, r->cfg.precision == GSM_PRECISION_FLOAT16, gaussians, harmonics, mc, po), "strip project+cull");
    GSM_CUDA(launchCompactVisible(s, gidCount, po, r->numSMs), "visibility compaction");
    return GSM_OK;
}

}  // namespace

extern "C" {

void gsm_config_default(gsm_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->maxGaussians = 6000000u;  // GRP.swift:211-218
    cfg->maxWidth = 1920;
    cfg->maxHeight = 1080;
    cfg->precision = GSM_PRECISION_FLOAT16;
    cfg->gaussianColorSpace = GSM_COLORSPACE_SRGB;
    cfg->depthSortKeyPrecision = GSM_KEY_BITS32;  // DFR.swift:48-49
    cfg->tileIdPrecision = GSM_KEY_BITS16;
    cfg->device = -1;
    cfg->stereoCopyFlipY = 1;
}

gsm_status gsm_renderer_create(const gsm_config* cfg, gsm_renderer** out) {
    if (!cfg || !out) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (cfg->maxGaussians > kMaxSupportedGaussians)  // DFR.swift:51-56
        return fail(GSM_ERR_INVALID_GAUSSIAN_COUNT, "maxGaussians exceeds 30,000,000");
    if (cfg->precision > 1 || cfg->gaussianColorSpace > 1 ||
        (cfg->depthSortKeyPrecision != 16 && cfg->depthSortKeyPrecision != 32) ||
        (cfg->tileIdPrecision != 16 && cfg->tileIdPrecision != 32))
        return fail(GSM_ERR_INVALID_ARGUMENT, "bad enum value in gsm_config");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(GSM_ERR_DEVICE_NOT_AVAILABLE, "no CUDA device");  // DFR.swift:58-61
    }
    int dev = cfg->device;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= n) return fail(GSM_ERR_DEVICE_NOT_AVAILABLE, "device ordinal out of range");
    // limits: RendererLimits(from:) clamps to >= 1 (GlobalRenderer.swift:17-23)
    gsm_config c = *cfg;
    if (c.maxGaussians < 1) c.maxGaussians = 1;
    if (c.maxWidth < 1) c.maxWidth = 1;
    if (c.maxHeight < 1) c.maxHeight = 1;
    const uint64_t tiles = (uint64_t)((c.maxWidth + kTile - 1) / kTile) * ((c.maxHeight + kTile - 1) / kTile);
    if (c.tileIdPrecision == GSM_KEY_BITS16 && tiles > 65535u)
        return fail(GSM_ERR_INVALID_TILE_COUNT, "more than 65535 tiles need tileIdPrecision 32");
    DeviceGuard guard(dev);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(GSM_ERR_DEVICE_NOT_AVAILABLE, "cudaGetDeviceProperties", e);
    // the kernels are built for sm_100a only: fail loudly anywhere else (no fallback path exists)
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, (const void*)kernel_image_probe());
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(GSM_ERR_FAILED_TO_CREATE_PIPELINE, "sm_100a kernels cannot be loaded on this device", e);
    }
    gsm_renderer* r = new (std::nothrow) gsm_renderer();
    if (!r) return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "host allocation failed");
    r->cfg = c;
    r->cfg.device = dev;
    r->device = dev;
    r->numSMs = prop.multiProcessorCount;
    if (blendUsesExpTable()) {
        e = cudaMalloc((void**)&r->expTable, blendExpTableBytes());
        if (e == cudaSuccess) e = buildBlendExpTable(nullptr, r->expTable);
        if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    }
    if (e == cudaSuccess) e = blendExpSelfTest(dev);  // once per device: which form of exp(-0.5h * p) its blend kernels run
    if (e != cudaSuccess) {
        if (r->expTable) cudaFree(r->expTable);
        delete r;
        return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "blend exp table / self-test", e);
    }
    *out = r;
    return GSM_OK;
}

void gsm_renderer_destroy(gsm_renderer* r) {
    if (!r) return;
    DeviceGuard guard(r->device);
    cudaDeviceSynchronize();
    freeResources(r->mono);
    freeResources(r->stereoRes);
    if (r->evValid) for (auto& e : r->ev) cudaEventDestroy(e);
    if (r->stGaussians) cudaFree(r->stGaussians);
    if (r->stHarmonics) cudaFree(r->stHarmonics);
    if (r->stColor) cudaFree(r->stColor);
    if (r->stDepth) cudaFree(r->stDepth);
    if (r->hostStream) cudaStreamDestroy(r->hostStream);
    if (r->stereoIntermediate) cudaFree(r->stereoIntermediate);
    if (r->rateTables) cudaFree(r->rateTables);
    if (r->srgbLut) cudaFree(r->srgbLut);
    if (r->expTable) cudaFree(r->expTable);
    if (r->globalArena) cudaFree(r->globalArena);
    delete r;
}

gsm_status gsm_render(gsm_renderer* r, void* stream, void* color, void* depth, const void* gaussians, const void* harmonics,
                      uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height) {
    if (!r || !camera) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (gaussianCount == 0 || gaussianCount > r->cfg.maxGaussians) return GSM_OK;  // DFR.swift:249 (quirk Q10)
    gsm_status st = validateFrame(r, width, height, gaussians, harmonics, color);
    if (st != GSM_OK) return st;
    DeviceGuard guard(r->device);
    st = ensureResources(r, r->mono, false);
    if (st != GSM_OK) return st;  // the reference returns silently here (DFR.swift:189); we report
    Resources& res = r->mono;
    res.frameGaussians = gaussianCount;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t tilesX = (width + kTile - 1) / kTile, tilesY = (height + kTile - 1) / kTile;
    r->lastTilesX = tilesX; r->lastTilesY = tilesY; r->lastStereo = false;

    recordStage(r, s, 0);
    // step 0: reset state + counters (DFR.swift:259-272)
    const bool zeroInKernel = gaussianCount >= 65536u;  // enough CTAs to clear the region in a few stores per thread
    if (!zeroInKernel) GSM_CUDA(cudaMemsetAsync(res.fs, 0, res.zeroBytes, s), "frame-state memset");
    // step 1 + 1.25 + 1.5
    MonoCam mc;
    fillMonoCam(mc, camera, gaussianCount, shComponents, width, height, r->cfg.gaussianColorSpace == GSM_COLORSPACE_SRGB);
    ProjectOut po;
    po.fs = res.fs; po.status = res.projStatus; po.statusGroups = res.projGroups; po.renderData = res.renderData; po.bounds = res.bounds;
    po.nTouched = res.nTouched; po.hitMask = res.hitMask; po.blendSplats = res.blendSplats; po.depthKeys = res.depthKeys[0];
    po.primitiveIndices = res.primIdx[0]; po.maxOut = res.maxGaussians; po.header = res.header; po.maxInstances = res.maxInstances; po.preDepthKeys = res.depthKeys[1];
    po.depthHist = &res.fs->hist[0][0]; po.depthPasses = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 2u : 4u;
    po.depthKey16 = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 1u : 0u; po.gidFirst = 0;
    if (zeroInKernel) { po.zeroBase = (uint4*)res.fs; po.zeroVecs = res.zeroBytes / 16; }
    attachDepthPlan(r, res, po, gaussianCount);
    GSM_CUDA(launchProjectMono(s, r->cfg.precision == GSM_PRECISION_FLOAT16, gaussians, harmonics, mc, po), "project+cull");
    GSM_CUDA(launchCompactVisible(s, gaussianCount, po, r->numSMs), "visibility compaction");
    recordStage(r, s, 1);
    st = encodeSortExpandRange(r, res, s, false, tilesX, tilesY, true, 0u, 0xFFFFFFFFu, po.keyRange != nullptr);
    if (st != GSM_OK) return st;
    // step 8: clear + blend (DFR.swift:433-464), fused
    GSM_CUDA(launchBlendMono(s, res.lowerBounds, res.blendSplats, res.instIdx[0], width, height, tilesX, tilesY, 0, tilesY,
                             (__half*)color, (__half*)depth, TileOut{res.tileHeaders, res.activeTiles, &res.fs->activeTileCount},
                             r->expTable, &res.fs->ticketBlend, r->numSMs), "blend");
    recordStage(r, s, 7);
    recordStage(r, s, 8);
    if (r->profiling) r->evRecorded = true;
    return GSM_OK;
}

gsm_status gsm_render_stereo(gsm_renderer* r, void* stream, void* colorSideBySide, const void* gaussians, const void* harmonics,
                             uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* leftEye, const gsm_camera* rightEye,
                             uint32_t width, uint32_t height) {
    return gsm_render_stereo_eyes(r, stream, colorSideBySide, gaussians, harmonics, gaussianCount, shComponents, leftEye, rightEye,
                                  width, height, 3u);
}

// encodeStereoPipeline (DFR.swift:595-831) up to and including the blend. sceneTransform == nullptr: identity (sideBySide).
// The blend writes both eyes side by side into `colorSideBySide` ((2*width) x height rgba16f), rows flipped iff flipY.
static gsm_status encodeStereoFrame(gsm_renderer* r, void* stream, void* colorSideBySide, const void* gaussians,
                                    const void* harmonics, uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* leftEye,
                                    const gsm_camera* rightEye, const float* sceneTransform, uint32_t width, uint32_t height,
                                    uint32_t eyeMask, bool flipY) {
    gsm_status st = validateFrame(r, width, height, gaussians, harmonics, colorSideBySide);
    if (st != GSM_OK) return st;
    DeviceGuard guard(r->device);
    st = ensureResources(r, r->stereoRes, true);
    if (st != GSM_OK) return st;
    Resources& res = r->stereoRes;
    res.frameGaussians = gaussianCount;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t tilesX = (width + kTile - 1) / kTile, tilesY = (height + kTile - 1) / kTile;
    r->lastTilesX = tilesX; r->lastTilesY = tilesY; r->lastStereo = true;

    recordStage(r, s, 0);
    const bool zeroInKernel = gaussianCount >= 65536u;  // enough CTAs to clear the region in a few stores per thread
    if (!zeroInKernel) GSM_CUDA(cudaMemsetAsync(res.fs, 0, res.zeroBytes, s), "frame-state memset");
    // makeStereoCameraUniforms (DFR.swift:554-591): near/far from the left eye; sceneTransform is the configuration's
    // (identity for sideBySide, GRP.swift:111)
    StereoCam sc;
    memcpy(sc.leftView, leftEye->viewMatrix, 64); memcpy(sc.leftProj, leftEye->projectionMatrix, 64);
    memcpy(sc.leftCenter, leftEye->position, 12);
    memcpy(sc.rightView, rightEye->viewMatrix, 64); memcpy(sc.rightProj, rightEye->projectionMatrix, 64);
    memcpy(sc.rightCenter, rightEye->position, 12);
    if (sceneTransform) memcpy(sc.sceneTransform, sceneTransform, 64);
    else {
        memset(sc.sceneTransform, 0, 64);
        sc.sceneTransform[0] = sc.sceneTransform[5] = sc.sceneTransform[10] = sc.sceneTransform[15] = 1.0f;
    }
    sc.width = (float)width; sc.height = (float)height;
    sc.nearPlane = leftEye->nearPlane; sc.farPlane = leftEye->farPlane;
    sc.shComponents = shComponents; sc.gaussianCount = gaussianCount;
    sc.inputIsSRGB = r->cfg.gaussianColorSpace == GSM_COLORSPACE_SRGB ? 1.0f : 0.0f;
    sc.tilesX = tilesX; sc.tilesY = tilesY;
    ProjectOut po;
    po.fs = res.fs; po.status = res.projStatus; po.statusGroups = res.projGroups; po.renderData = res.renderData; po.bounds = res.bounds;
    po.nTouched = res.nTouched; po.hitMask = res.hitMask; po.blendSplats = nullptr; po.depthKeys = res.depthKeys[0];
    po.primitiveIndices = res.primIdx[0]; po.maxOut = res.maxGaussians; po.header = res.header; po.maxInstances = res.maxInstances; po.preDepthKeys = res.depthKeys[1];
    po.depthHist = &res.fs->hist[0][0]; po.depthPasses = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 2u : 4u;
    po.depthKey16 = r->cfg.depthSortKeyPrecision == GSM_KEY_BITS16 ? 1u : 0u; po.gidFirst = 0;
    if (zeroInKernel) { po.zeroBase = (uint4*)res.fs; po.zeroVecs = res.zeroBytes / 16; }
    attachDepthPlan(r, res, po, gaussianCount);
    GSM_CUDA(launchProjectStereo(s, r->cfg.precision == GSM_PRECISION_FLOAT16, gaussians, harmonics, sc, po), "stereo project+cull");
    GSM_CUDA(launchCompactVisible(s, gaussianCount, po, r->numSMs), "visibility compaction");
    recordStage(r, s, 1);
    st = encodeSortExpandRange(r, res, s, true, tilesX, tilesY, true, 0u, 0xFFFFFFFFu, po.keyRange != nullptr);
    if (st != GSM_OK) return st;
    // steps 9+10: clear, blend both eyes, copy into the side-by-side target (DFR.swift:789-830), fused
    GSM_CUDA(launchBlendStereo(s, res.lowerBounds, (const GSMStereoTiledRenderData*)res.renderData, res.instIdx[0], width, height,
                               tilesX, tilesY, (__half*)colorSideBySide, (int)(eyeMask & 3u), flipY ? 1 : 0,
                               TileOut{res.tileHeaders, res.activeTiles, &res.fs->activeTileCount}), "stereo blend");
    recordStage(r, s, 7);
    return GSM_OK;
}

gsm_status gsm_render_stereo_eyes(gsm_renderer* r, void* stream, void* colorSideBySide, const void* gaussians,
                                  const void* harmonics, uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* leftEye,
                                  const gsm_camera* rightEye, uint32_t width, uint32_t height, uint32_t eyeMask) {
    if (!r || !leftEye || !rightEye) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if ((eyeMask & 3u) == 0) return fail(GSM_ERR_INVALID_ARGUMENT, "eyeMask selects no eye");
    if (gaussianCount == 0 || gaussianCount > r->cfg.maxGaussians) return GSM_OK;  // DFR.swift:478,607
    // steps 9+10 fused: at 1:1 the copy (DFR.swift:823-830) is the identity up to the row flip, so the blend writes the target
    gsm_status st = encodeStereoFrame(r, stream, colorSideBySide, gaussians, harmonics, gaussianCount, shComponents, leftEye, rightEye,
                                      nullptr, width, height, eyeMask, r->cfg.stereoCopyFlipY != 0);
    if (st != GSM_OK) return st;
    recordStage(r, (cudaStream_t)stream, 8);
    if (r->profiling) r->evRecorded = true;
    return GSM_OK;
}

// ---- StereoRenderTarget.foveated
static size_t pixelBytes(uint32_t format) { return format == GSM_PIXEL_RGBA16F ? 8u : 4u; }

static gsm_status validateDrawable(const gsm_foveated_drawable* d) {
    if (!d || !d->colorTexture) return fail(GSM_ERR_INVALID_ARGUMENT, "null drawable");
    if (d->colorPixelFormat > GSM_PIXEL_RGBA8_SRGB) return fail(GSM_ERR_INVALID_ARGUMENT, "unsupported drawable pixel format");
    if (d->arrayLength < 1 || d->arrayLength > 2) return fail(GSM_ERR_INVALID_ARGUMENT, "drawable arrayLength must be 1 or 2");
    const size_t px = pixelBytes(d->colorPixelFormat);
    if (d->textureWidth == 0 || d->textureHeight == 0 || d->rowBytes < (size_t)d->textureWidth * px || (d->rowBytes % px) != 0 ||
        ((uintptr_t)d->colorTexture % px) != 0)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "drawable rows must hold textureWidth texels and be texel-aligned");
    if (d->arrayLength == 2 && (d->sliceBytes < d->rowBytes * d->textureHeight || (d->sliceBytes % px) != 0))
        return fail(GSM_ERR_INVALID_DIMENSIONS, "drawable slices overlap");
    if (const gsm_rate_map* m = d->rasterizationRateMap) {
        if (m->layerCount < 1 || m->layerCount > 2) return fail(GSM_ERR_INVALID_ARGUMENT, "rate map layerCount must be 1 or 2");
        for (uint32_t l = 0; l < m->layerCount; ++l)
            if (!m->layers[l].screenX || !m->layers[l].screenY || m->layers[l].physicalWidth == 0 || m->layers[l].physicalHeight == 0)
                return fail(GSM_ERR_INVALID_ARGUMENT, "rate map layer without tables");
    }
    return GSM_OK;
}

// step 10 (DFR.swift:823-830, DepthFirstStereoCopyEncoder.swift:28-100) from a side-by-side intermediate image
static gsm_status encodeStereoCopy(gsm_renderer* r, cudaStream_t s, const void* intermediate, uint32_t width, uint32_t height,
                                   const gsm_foveated_drawable* d, const gsm_viewport* lv, const gsm_viewport* rv) {
    StereoCopyParams p{};
    p.src = (const __half*)intermediate; p.srcEyeStride = (size_t)width * 4; p.srcRowStride = (size_t)width * 8;
    p.width = width; p.height = height;
    p.dst = d->colorTexture; p.textureWidth = d->textureWidth; p.textureHeight = d->textureHeight; p.arrayLength = d->arrayLength;
    p.rowBytes = d->rowBytes; p.sliceBytes = d->arrayLength == 2 ? d->sliceBytes : 0; p.format = d->colorPixelFormat;
    p.flipY = r->cfg.stereoCopyFlipY ? 1 : 0;
    const gsm_viewport* v[2] = {lv, rv};
    for (int e = 0; e < 2; ++e) {
        p.vp[e][0] = (float)v[e]->originX; p.vp[e][1] = (float)v[e]->originY;
        p.vp[e][2] = (float)v[e]->width; p.vp[e][3] = (float)v[e]->height;
    }
    if (d->colorPixelFormat == GSM_PIXEL_BGRA8_SRGB || d->colorPixelFormat == GSM_PIXEL_RGBA8_SRGB) {
        if (!r->srgbLut) {
            std::vector<uint8_t> host(kSrgbTableSize);
            buildSrgbEncodeTable(host.data());
            GSM_CUDA(cudaMalloc(&r->srgbLut, kSrgbTableSize), "sRGB table");
            GSM_CUDA(cudaMemcpy(r->srgbLut, host.data(), kSrgbTableSize, cudaMemcpyHostToDevice), "sRGB table upload");
        }
        p.srgbLut = r->srgbLut;
    }
    if (const gsm_rate_map* m = d->rasterizationRateMap) {
        size_t need = 0;
        for (uint32_t l = 0; l < m->layerCount; ++l) need += (size_t)m->layers[l].physicalWidth + m->layers[l].physicalHeight;
        if (r->rateTableFloats < need) {
            GSM_CUDA(cudaStreamSynchronize(s), "rate-map tables in use");
            if (r->rateTables) cudaFree(r->rateTables);
            r->rateTables = nullptr; r->rateTableFloats = 0;
            GSM_CUDA(cudaMalloc(&r->rateTables, need * sizeof(float)), "rate-map tables");
            r->rateTableFloats = need;
        }
        // pageable host memory: the copy has left the caller's arrays when the call returns
        float* at = r->rateTables;
        p.layerCount = m->layerCount;
        for (uint32_t l = 0; l < m->layerCount; ++l) {
            const gsm_rate_map_layer& L = m->layers[l];
            GSM_CUDA(cudaMemcpyAsync(at, L.screenX, (size_t)L.physicalWidth * 4, cudaMemcpyHostToDevice, s), "rate-map upload");
            p.screenX[l] = at; at += L.physicalWidth;
            GSM_CUDA(cudaMemcpyAsync(at, L.screenY, (size_t)L.physicalHeight * 4, cudaMemcpyHostToDevice, s), "rate-map upload");
            p.screenY[l] = at; at += L.physicalHeight;
            p.physicalWidth[l] = L.physicalWidth; p.physicalHeight[l] = L.physicalHeight;
        }
    }
    GSM_CUDA(launchStereoCopy(s, p), "stereo copy");
    return GSM_OK;
}

gsm_status gsm_stereo_copy(gsm_renderer* r, void* stream, const void* intermediate, uint32_t width, uint32_t height,
                           const gsm_foveated_drawable* drawable, const gsm_viewport* leftViewport, const gsm_viewport* rightViewport) {
    if (!r || !intermediate || !leftViewport || !rightViewport) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (width == 0 || height == 0) return fail(GSM_ERR_INVALID_DIMENSIONS, "empty intermediate image");
    if (((uintptr_t)intermediate & 7u) != 0) return fail(GSM_ERR_INVALID_ARGUMENT, "intermediate image must be 8-byte aligned");
    gsm_status st = validateDrawable(drawable);
    if (st != GSM_OK) return st;
    DeviceGuard guard(r->device);
    return encodeStereoCopy(r, (cudaStream_t)stream, intermediate, width, height, drawable, leftViewport, rightViewport);
}

gsm_status gsm_render_stereo_foveated(gsm_renderer* r, void* stream, const gsm_foveated_drawable* drawable, const void* gaussians,
                                      const void* harmonics, uint32_t gaussianCount, uint32_t shComponents,
                                      const gsm_stereo_configuration* cfg, uint32_t width, uint32_t height) {
    if (!r || !cfg) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (gaussianCount == 0 || gaussianCount > r->cfg.maxGaussians) return GSM_OK;  // DFR.swift:524
    gsm_status st = validateDrawable(drawable);
    if (st != GSM_OK) return st;
    if (width == 0 || height == 0 || width > r->cfg.maxWidth || height > r->cfg.maxHeight)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "dimensions exceed RendererConfig.maxWidth/maxHeight");
    DeviceGuard guard(r->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t need = (size_t)2 * width * height * 8;
    if (r->stereoIntermediateBytes < need) {  // ensureStereoIntermediateColor (DFR.swift:139-164)
        GSM_CUDA(cudaStreamSynchronize(s), "intermediate image in use");
        if (r->stereoIntermediate) cudaFree(r->stereoIntermediate);
        r->stereoIntermediate = nullptr; r->stereoIntermediateBytes = 0;
        GSM_CUDA(cudaMalloc(&r->stereoIntermediate, need), "stereo intermediate image");
        r->stereoIntermediateBytes = need;
    }
    st = encodeStereoFrame(r, stream, r->stereoIntermediate, gaussians, harmonics, gaussianCount, shComponents, &cfg->leftEye.camera,
                           &cfg->rightEye.camera, cfg->sceneTransform, width, height, 3u, false);
    if (st != GSM_OK) return st;
    st = encodeStereoCopy(r, s, r->stereoIntermediate, width, height, drawable, &cfg->leftEye.viewport, &cfg->rightEye.viewport);
    if (st != GSM_OK) return st;
    recordStage(r, s, 8);
    if (r->profiling) r->evRecorded = true;
    return GSM_OK;
}

gsm_status gsm_render_host_async(gsm_renderer* r, const void* hostGaussians, const void* hostHarmonics, uint32_t gaussianCount,
                                 uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height, void* hostColor,
                                 void* hostDepth) {
    if (!r || !camera || !hostGaussians || !hostHarmonics || !hostColor) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (gaussianCount == 0 || gaussianCount > r->cfg.maxGaussians) return GSM_OK;
    if (width == 0 || height == 0 || width > r->cfg.maxWidth || height > r->cfg.maxHeight)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "dimensions exceed RendererConfig.maxWidth/maxHeight");
    DeviceGuard guard(r->device);
    const bool half = r->cfg.precision == GSM_PRECISION_FLOAT16;
    const int deg = shDegreeFromComponents(shComponents);
    const size_t k = deg == 0 ? 1 : (deg == 1 ? 4 : (deg == 2 ? 9 : 16));
    const size_t gBytes = (size_t)gaussianCount * (half ? 32 : 48);
    const size_t hBytes = (size_t)gaussianCount * 3 * k * (half ? 2 : 4);
    const size_t cBytes = (size_t)width * height * 8, dBytes = (size_t)width * height * 2;
    auto ensure = [&](void*& p, size_t& have, size_t need) -> cudaError_t {
        if (have >= need) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; have = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) have = need;
        return e;
    };
    if (!r->hostStream) GSM_CUDA(cudaStreamCreateWithFlags(&r->hostStream, cudaStreamNonBlocking), "stream create");
    GSM_CUDA(ensure(r->stGaussians, r->stGaussiansBytes, gBytes), "staging alloc");
    GSM_CUDA(ensure(r->stHarmonics, r->stHarmonicsBytes, hBytes), "staging alloc");
    GSM_CUDA(ensure(r->stColor, r->stColorBytes, cBytes), "staging alloc");
    if (hostDepth) GSM_CUDA(ensure(r->stDepth, r->stDepthBytes, dBytes), "staging alloc");
    cudaStream_t s = r->hostStream;
    GSM_CUDA(cudaMemcpyAsync(r->stGaussians, hostGaussians, gBytes, cudaMemcpyHostToDevice, s), "H2D gaussians");
    GSM_CUDA(cudaMemcpyAsync(r->stHarmonics, hostHarmonics, hBytes, cudaMemcpyHostToDevice, s), "H2D harmonics");
    gsm_status st = gsm_render(r, s, r->stColor, hostDepth ? r->stDepth : nullptr, r->stGaussians, r->stHarmonics, gaussianCount,
                               shComponents, camera, width, height);
    if (st != GSM_OK) return st;
    GSM_CUDA(cudaMemcpyAsync(hostColor, r->stColor, cBytes, cudaMemcpyDeviceToHost, s), "D2H colour");
    if (hostDepth) GSM_CUDA(cudaMemcpyAsync(hostDepth, r->stDepth, dBytes, cudaMemcpyDeviceToHost, s), "D2H depth");
    return GSM_OK;
}

gsm_status gsm_render_host_wait(gsm_renderer* r) {
    if (!r) return fail(GSM_ERR_INVALID_ARGUMENT, "null renderer");
    if (!r->hostStream) return GSM_OK;
    DeviceGuard guard(r->device);
    GSM_CUDA(cudaStreamSynchronize(r->hostStream), "stream sync");
    return GSM_OK;
}

gsm_status gsm_render_host(gsm_renderer* r, const void* hostGaussians, const void* hostHarmonics, uint32_t gaussianCount,
                           uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height, void* hostColor,
                           void* hostDepth) {
    gsm_status st = gsm_render_host_async(r, hostGaussians, hostHarmonics, gaussianCount, shComponents, camera, width, height,
                                          hostColor, hostDepth);
    if (st != GSM_OK) return st;
    return gsm_render_host_wait(r);
}

gsm_status gsm_strip_project(gsm_renderer* r, void* stream, const void* gaussians, const void* harmonics, uint32_t gidFirst,
                             uint32_t gidCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height,
                             void* recordsOut, uint32_t* hostCount) {
    if (!r || !camera || !recordsOut || !hostCount) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    *hostCount = 0;
    if (gidCount == 0) return GSM_OK;
    if ((uint64_t)gidFirst + gidCount > r->cfg.maxGaussians) return fail(GSM_ERR_INVALID_GAUSSIAN_COUNT, "gid range exceeds maxGaussians");
    gsm_status st = validateFrame(r, width, height, gaussians, harmonics, recordsOut);
    if (st != GSM_OK) return st;
    DeviceGuard guard(r->device);
    st = ensureResources(r, r->mono, false);
    if (st != GSM_OK) return st;
    Resources& res = r->mono;
    cudaStream_t s = (cudaStream_t)stream;
    st = encodeShardProject(r, res, s, gaussians, harmonics, gidFirst, gidCount, shComponents, camera, width, height);
    if (st != GSM_OK) return st;
    GSM_CUDA(launchPackRecords(s, res.fs, res.depthKeys[0], res.primIdx[0], res.renderData, res.bounds, res.hitMask, recordsOut,
                               gidCount, r->numSMs), "pack records");
    GSM_CUDA(cudaMemcpyAsync(hostCount, &res.fs->visibleCountRaw, 4, cudaMemcpyDeviceToHost, s), "count readback");
    GSM_CUDA(cudaStreamSynchronize(s), "strip project sync");
    return GSM_OK;
}

gsm_status gsm_strip_render(gsm_renderer* r, void* stream, void* color, void* depth, const void* records, uint32_t recordCount,
                            uint32_t width, uint32_t height, uint32_t tileRowFirst, uint32_t tileRowCount) {
    if (!r || !color || (!records && recordCount)) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (recordCount > r->cfg.maxGaussians) return fail(GSM_ERR_INVALID_GAUSSIAN_COUNT, "record count exceeds maxGaussians");
    if (width == 0 || height == 0 || width > r->cfg.maxWidth || height > r->cfg.maxHeight)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "dimensions exceed RendererConfig.maxWidth/maxHeight");
    const uint32_t tilesX = (width + kTile - 1) / kTile, tilesY = (height + kTile - 1) / kTile;
    if (tileRowFirst > tilesY || tileRowCount > tilesY - tileRowFirst) return fail(GSM_ERR_INVALID_ARGUMENT, "tile rows out of range");
    DeviceGuard guard(r->device);
    gsm_status st = ensureResources(r, r->mono, false);
    if (st != GSM_OK) return st;
    Resources& res = r->mono;
    res.frameGaussians = recordCount;
    cudaStream_t s = (cudaStream_t)stream;
    r->lastTilesX = tilesX; r->lastTilesY = tilesY; r->lastStereo = false;
    GSM_CUDA(cudaMemsetAsync(res.fs, 0, res.zeroBytes, s), "frame-state memset");
    ProjectOut po = stripIngestOutputs(r, res);
    // no record at all: nothing below writes the frame header (the compaction's last tile does, and it has no tile), and the
    // header lies outside the zeroed region -- clear it so that stages 2-7 and the blend see an empty frame, not the previous one
    if (recordCount == 0) GSM_CUDA(cudaMemsetAsync(res.header, 0, sizeof(GSMDepthFirstHeader), s), "empty-frame header");
    GSM_CUDA(launchIngestRecords(s, records, recordCount, tileRowFirst, tileRowCount, po, res.offsets, res.depthKeys[1],
                                 (uint32_t*)res.primIdx[1]), "ingest records");
    return encodeStripTail(r, res, s, po, recordCount, color, depth, width, height, tileRowFirst, tileRowCount);
}

// ---------------------------------------------------------------- gsm_group: the strip-sharded frame over peer memory (group.cu)
struct gsm_group {
    gsm_renderer* r = nullptr;
    uint32_t rank = 0, world = 1, regionCap = 0, seq = 0;
    char* window = nullptr;           // this rank's exchange window: [GroupMailbox | world receive regions | image colour | image depth]
    size_t windowBytes = 0, oRegions = 0, oImage = 0, imageColorBytes = 0, imageDepthBytes = 0;
    char* base[kGroupMaxRanks] = {};  // every rank's window as mapped here; base[rank] == window
    bool ipcOpened[kGroupMaxRanks] = {};
    uint32_t* routeStatus = nullptr;  // look-back words of the routing kernel
    size_t routeStatusBytes = 0;
    bool connected = false;
};

static gsm_status groupCheck(const gsm_group* g, bool needConnected) {
    if (!g || !g->r) return fail(GSM_ERR_INVALID_ARGUMENT, "null group");
    if (needConnected && !g->connected) return fail(GSM_ERR_INVALID_ARGUMENT, "group is not connected (gsm_group_connect / _connect_local)");
    return GSM_OK;
}

gsm_status gsm_group_create(gsm_renderer* r, uint32_t rank, uint32_t world, uint32_t maxRecordsPerSource, size_t imageColorBytes,
                            size_t imageDepthBytes, gsm_group** out) {
    if (!r || !out) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (world < 1 || world > kGroupMaxRanks || rank >= world) return fail(GSM_ERR_INVALID_ARGUMENT, "group of 1..8 ranks, rank < world");
    if (maxRecordsPerSource == 0 || maxRecordsPerSource > r->cfg.maxGaussians)
        return fail(GSM_ERR_INVALID_GAUSSIAN_COUNT, "maxRecordsPerSource must be in [1, maxGaussians]");
    DeviceGuard guard(r->device);
    gsm_group* g = new (std::nothrow) gsm_group();
    if (!g) return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "host allocation failed");
    g->r = r; g->rank = rank; g->world = world; g->regionCap = maxRecordsPerSource;
    g->oRegions = alignUp(sizeof(GroupMailbox), 256);
    g->oImage = alignUp(g->oRegions + (size_t)world * maxRecordsPerSource * sizeof(SplatRecord), 256);
    g->imageColorBytes = alignUp(imageColorBytes, 256); g->imageDepthBytes = alignUp(imageDepthBytes, 256);
    g->windowBytes = g->oImage + g->imageColorBytes + g->imageDepthBytes;
    g->routeStatusBytes = (size_t)routeStatusWords(maxRecordsPerSource) * 4u;
    cudaError_t e = cudaMalloc((void**)&g->window, g->windowBytes);   // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->routeStatus, g->routeStatusBytes);
    if (e == cudaSuccess) e = cudaMemset(g->window, 0, g->oRegions);   // mailbox: sequence 0 = nothing sent, everything acked
    if (e != cudaSuccess) {
        if (g->window) cudaFree(g->window);
        if (g->routeStatus) cudaFree(g->routeStatus);
        delete g;
        return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "exchange window", e);
    }
    g->base[rank] = g->window;
    if (world == 1) g->connected = true;
    *out = g;
    return GSM_OK;
}

gsm_status gsm_group_export(gsm_group* g, void* handleOut) {
    gsm_status st = groupCheck(g, false);
    if (st != GSM_OK) return st;
    if (!handleOut) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == GSM_GROUP_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(g->r->device);
    cudaIpcMemHandle_t h;
    GSM_CUDA(cudaIpcGetMemHandle(&h, g->window), "cudaIpcGetMemHandle");
    memcpy(handleOut, &h, sizeof h);
    return GSM_OK;
}

gsm_status gsm_group_connect(gsm_group* g, const void* handles) {
    gsm_status st = groupCheck(g, false);
    if (st != GSM_OK) return st;
    if (!handles) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(g->r->device);
    for (uint32_t p = 0; p < g->world; ++p) {
        if (p == g->rank || g->base[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)p * GSM_GROUP_HANDLE_BYTES, sizeof h);
        void* mapped = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(GSM_ERR_DEVICE_NOT_AVAILABLE, "cudaIpcOpenMemHandle (peer window; needs NVLink/PCIe peer access)", e);
        g->base[p] = (char*)mapped;
        g->ipcOpened[p] = true;
    }
    g->connected = true;
    return GSM_OK;
}

gsm_status gsm_group_connect_local(gsm_group* g, gsm_group* const* peers) {
    gsm_status st = groupCheck(g, false);
    if (st != GSM_OK) return st;
    if (!peers) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(g->r->device);
    for (uint32_t p = 0; p < g->world; ++p) {
        if (p == g->rank) continue;
        const gsm_group* q = peers[p];
        if (!q || q->world != g->world || q->rank != p || q->regionCap != g->regionCap)
            return fail(GSM_ERR_INVALID_ARGUMENT, "peer group does not match (world, rank, maxRecordsPerSource)");
        if (q->r->device != g->r->device) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->r->device, q->r->device);
            if (!can) return fail(GSM_ERR_DEVICE_NOT_AVAILABLE, "no peer access between the group's devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(q->r->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(GSM_ERR_DEVICE_NOT_AVAILABLE, "cudaDeviceEnablePeerAccess", e);
            cudaGetLastError();
        }
        g->base[p] = q->window;
    }
    g->connected = true;
    return GSM_OK;
}

void gsm_group_destroy(gsm_group* g) {
    if (!g) return;
    DeviceGuard guard(g->r->device);
    cudaDeviceSynchronize();
    for (uint32_t p = 0; p < kGroupMaxRanks; ++p)
        if (g->ipcOpened[p] && g->base[p]) cudaIpcCloseMemHandle(g->base[p]);
    if (g->window) cudaFree(g->window);
    if (g->routeStatus) cudaFree(g->routeStatus);
    delete g;
}

gsm_status gsm_group_image(gsm_group* g, uint32_t ofRank, void** color, void** depth) {
    gsm_status st = groupCheck(g, true);
    if (st != GSM_OK) return st;
    if (ofRank >= g->world || !g->base[ofRank]) return fail(GSM_ERR_INVALID_ARGUMENT, "rank out of range");
    if (color) *color = g->imageColorBytes ? g->base[ofRank] + g->oImage : nullptr;
    if (depth) *depth = g->imageDepthBytes ? g->base[ofRank] + g->oImage + g->imageColorBytes : nullptr;
    return GSM_OK;
}

static gsm_status checkStrips(const gsm_group* g, const uint32_t* rowStart, uint32_t tilesY) {
    if (!rowStart) return fail(GSM_ERR_INVALID_ARGUMENT, "null strip table");
    if (rowStart[0] != 0 || rowStart[g->world] != tilesY) return fail(GSM_ERR_INVALID_ARGUMENT, "strips must cover tile rows [0, tilesY)");
    for (uint32_t d = 0; d < g->world; ++d)
        if (rowStart[d] > rowStart[d + 1]) return fail(GSM_ERR_INVALID_ARGUMENT, "strip table must be non-decreasing");
    return GSM_OK;
}

gsm_status gsm_group_project_route(gsm_group* g, void* stream, const void* gaussians, const void* harmonics, uint32_t gidFirst,
                                   uint32_t gidCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height,
                                   const uint32_t* stripRowStart) {
    gsm_status st = groupCheck(g, true);
    if (st != GSM_OK) return st;
    gsm_renderer* r = g->r;
    if (!camera) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (gidCount > g->regionCap) return fail(GSM_ERR_INVALID_GAUSSIAN_COUNT, "shard larger than the group's maxRecordsPerSource");
    if ((uint64_t)gidFirst + gidCount > r->cfg.maxGaussians) return fail(GSM_ERR_INVALID_GAUSSIAN_COUNT, "gid range exceeds maxGaussians");
    if (width == 0 || height == 0 || width > r->cfg.maxWidth || height > r->cfg.maxHeight)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "dimensions exceed RendererConfig.maxWidth/maxHeight");
    if (gidCount && (!gaussians || !harmonics || !aligned16(gaussians) || !aligned16(harmonics)))
        return fail(GSM_ERR_INVALID_ARGUMENT, "input buffers must be non-null and 16-byte aligned");
    st = checkStrips(g, stripRowStart, (height + kTile - 1) / kTile);
    if (st != GSM_OK) return st;
    DeviceGuard guard(r->device);
    st = ensureResources(r, r->mono, false);
    if (st != GSM_OK) return st;
    Resources& res = r->mono;
    cudaStream_t s = (cudaStream_t)stream;
    g->seq++;   // every rank calls this once per frame: the sequence numbers agree without being exchanged
    if (gidCount) {
        st = encodeShardProject(r, res, s, gaussians, harmonics, gidFirst, gidCount, shComponents, camera, width, height);
        if (st != GSM_OK) return st;
    } else {
        GSM_CUDA(cudaMemsetAsync(res.fs, 0, res.zeroBytes, s), "frame-state memset");  // visibleCountRaw = 0: the routing kernel publishes zero counts
    }
    GSM_CUDA(cudaMemsetAsync(g->routeStatus, 0, g->routeStatusBytes, s), "route status memset");
    RouteParams P{};
    P.world = g->world; P.rank = g->rank; P.seq = g->seq; P.regionCap = g->regionCap;
    for (uint32_t d = 0; d <= g->world; ++d) P.rowStart[d] = stripRowStart[d];
    for (uint32_t d = 0; d < g->world; ++d) {
        P.region[d] = (SplatRecord*)(g->base[d] + g->oRegions) + (size_t)g->rank * g->regionCap;
        P.mailbox[d] = (GroupMailbox*)g->base[d];
    }
    P.mine = (GroupMailbox*)g->window;
    GSM_CUDA(launchRouteRecords(s, res.fs, res.depthKeys[0], res.primIdx[0], res.renderData, res.bounds, res.hitMask,
                                gidCount ? gidCount : 1u, g->routeStatus, P, r->numSMs), "route records");
    return GSM_OK;
}

gsm_status gsm_group_render_strip(gsm_group* g, void* stream, void* color, void* depth, uint32_t width, uint32_t height,
                                  const uint32_t* stripRowStart) {
    gsm_status st = groupCheck(g, true);
    if (st != GSM_OK) return st;
    gsm_renderer* r = g->r;
    if (!color) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (width == 0 || height == 0 || width > r->cfg.maxWidth || height > r->cfg.maxHeight)
        return fail(GSM_ERR_INVALID_DIMENSIONS, "dimensions exceed RendererConfig.maxWidth/maxHeight");
    const uint32_t tilesX = (width + kTile - 1) / kTile, tilesY = (height + kTile - 1) / kTile;
    st = checkStrips(g, stripRowStart, tilesY);
    if (st != GSM_OK) return st;
    if (g->seq == 0) return fail(GSM_ERR_INVALID_ARGUMENT, "gsm_group_render_strip before gsm_group_project_route");
    DeviceGuard guard(r->device);
    st = ensureResources(r, r->mono, false);
    if (st != GSM_OK) return st;
    Resources& res = r->mono;
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t rowFirst = stripRowStart[g->rank], rowCount = stripRowStart[g->rank + 1] - rowFirst;
    // a strip receives about 1/world of the frame's splats plus those straddling its borders: picks the sorts' tile size only
    res.frameGaussians = g->regionCap;
    r->lastTilesX = tilesX; r->lastTilesY = tilesY; r->lastStereo = false;
    GSM_CUDA(cudaMemsetAsync(res.fs, 0, res.zeroBytes, s), "frame-state memset");
    ProjectOut po = stripIngestOutputs(r, res);
    IngestParams P{};
    P.world = g->world; P.rank = g->rank; P.seq = g->seq; P.regionCap = g->regionCap;
    P.rowFirst = (int)rowFirst; P.rowLast = (int)(rowFirst + rowCount) - 1;
    for (uint32_t sR = 0; sR < g->world; ++sR) {
        P.region[sR] = (const SplatRecord*)(g->window + g->oRegions) + (size_t)sR * g->regionCap;
        P.mailbox[sR] = (GroupMailbox*)g->base[sR];
    }
    P.mine = (GroupMailbox*)g->window;
    GSM_CUDA(launchIngestRouted(s, P, po, res.offsets, res.depthKeys[1], (uint32_t*)res.primIdx[1], r->numSMs), "ingest routed records");
    po.countPtr = &res.fs->recordTotal;
    return encodeStripTail(r, res, s, po, res.maxGaussians, color, depth, width, height, rowFirst, rowCount);
}

gsm_status gsm_group_signal(gsm_group* g, void* stream, uint32_t toRank, uint32_t frameId) {
    gsm_status st = groupCheck(g, true);
    if (st != GSM_OK) return st;
    if (toRank >= g->world) return fail(GSM_ERR_INVALID_ARGUMENT, "rank out of range");
    DeviceGuard guard(g->r->device);
    GSM_CUDA(launchGroupSignal((cudaStream_t)stream, (GroupMailbox*)g->base[toRank], g->rank, frameId), "group signal");
    return GSM_OK;
}

gsm_status gsm_group_wait(gsm_group* g, void* stream, uint32_t fromMask, uint32_t frameId) {
    gsm_status st = groupCheck(g, true);
    if (st != GSM_OK) return st;
    DeviceGuard guard(g->r->device);
    GSM_CUDA(launchGroupWait((cudaStream_t)stream, (const GroupMailbox*)g->window, fromMask & ((1u << g->world) - 1u), frameId), "group wait");
    return GSM_OK;
}

gsm_status gsm_group_record_counts(gsm_group* g, void* stream, uint32_t* countsOut) {
    gsm_status st = groupCheck(g, true);
    if (st != GSM_OK) return st;
    if (!countsOut) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(g->r->device);
    uint32_t host[kGroupMaxRanks];
    GSM_CUDA(cudaMemcpyAsync(host, ((GroupMailbox*)g->window)->recordCount, sizeof host, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "mailbox read");
    GSM_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "mailbox read sync");
    for (uint32_t s2 = 0; s2 < g->world; ++s2) countsOut[s2] = host[s2];
    return GSM_OK;
}

gsm_status gsm_render_strips(gsm_group* g, void* stream, void* color, void* depth, const void* gaussians, const void* harmonics,
                             uint32_t gidFirst, uint32_t gidCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width,
                             uint32_t height, const uint32_t* stripRowStart) {
    gsm_status st = gsm_group_project_route(g, stream, gaussians, harmonics, gidFirst, gidCount, shComponents, camera, width, height,
                                            stripRowStart);
    if (st != GSM_OK) return st;
    return gsm_group_render_strip(g, stream, color, depth, width, height, stripRowStart);
}

double gsm_last_gpu_time_ms(gsm_renderer* r) {
    if (!r || !r->profiling || !r->evRecorded) return -1.0;
    float ms[GSM_NUM_STAGES];
    if (gsm_get_stage_times_ms(r, ms) != GSM_OK) return -1.0;
    return r->lastMs;
}

gsm_status gsm_set_profiling(gsm_renderer* r, int enabled) {
    if (!r) return fail(GSM_ERR_INVALID_ARGUMENT, "null renderer");
    DeviceGuard guard(r->device);
    if (enabled && !r->evValid) {
        for (auto& e : r->ev) GSM_CUDA(cudaEventCreate(&e), "event create");
        r->evValid = true;
    }
    r->profiling = enabled != 0;
    r->evRecorded = false;
    return GSM_OK;
}

gsm_status gsm_get_stage_times_ms(gsm_renderer* r, float* ms) {
    if (!r || !ms) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (!r->profiling || !r->evRecorded) return fail(GSM_ERR_INVALID_ARGUMENT, "profiling is off or no frame recorded");
    DeviceGuard guard(r->device);
    GSM_CUDA(cudaEventSynchronize(r->ev[GSM_NUM_STAGES]), "event sync");
    double total = 0;
    for (int i = 0; i < GSM_NUM_STAGES; ++i) {
        float t = 0;
        GSM_CUDA(cudaEventElapsedTime(&t, r->ev[i], r->ev[i + 1]), "event elapsed");
        ms[i] = r->stageMs[i] = t;
        total += t;
    }
    r->lastMs = total;
    return GSM_OK;
}

const char* gsm_stage_name(int stage) {
    static const char* names[GSM_NUM_STAGES] = {"project", "depthSort", "applyScan", "expand", "tileSort", "ranges", "blend", "copy"};
    return (stage >= 0 && stage < GSM_NUM_STAGES) ? names[stage] : "";
}

size_t gsm_debug_element_size(gsm_renderer* r, int which) {
    if (!r) return 0;
    const bool tile16 = r->cfg.tileIdPrecision == GSM_KEY_BITS16;
    switch (which) {
        case GSM_DBG_HEADER: return sizeof(GSMDepthFirstHeader);
        case GSM_DBG_ACTIVE_TILE_COUNT: return 4;
        case GSM_DBG_SORTED_TILE_IDS: return tile16 ? 2 : 4;
        case GSM_DBG_TILE_BOUNDS: return 16;
        case GSM_DBG_RENDER_DATA: return r->lastStereo ? 32 : 16;
        case GSM_DBG_TILE_HEADERS: return 8;
        case GSM_DBG_DEPTH_SORT_PLAN: return 16;
        case GSM_DBG_SORTED_PRIMITIVE_INDICES: case GSM_DBG_INSTANCE_OFFSETS: case GSM_DBG_N_TOUCHED_TILES:
        case GSM_DBG_INSTANCE_GAUSSIAN_INDICES: case GSM_DBG_DEPTH_KEYS: case GSM_DBG_ACTIVE_TILES:
        case GSM_DBG_SCRATCH_DEPTH_KEYS: case GSM_DBG_SCRATCH_PRIMITIVE_INDICES: return 4;
        default: return 0;
    }
}

gsm_status gsm_debug_read(gsm_renderer* r, void* stream, int which, void* dst, size_t first, size_t count) {
    if (!r || !dst) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    Resources& res = r->lastStereo ? r->stereoRes : r->mono;
    if (!res.arena) return fail(GSM_ERR_INVALID_ARGUMENT, "no frame has been rendered");
    DeviceGuard guard(r->device);
    const char* src = nullptr;
    size_t cap = 0;
    const size_t es = gsm_debug_element_size(r, which);
    switch (which) {
        case GSM_DBG_HEADER: src = (const char*)res.header; cap = 1; break;
        case GSM_DBG_ACTIVE_TILE_COUNT: src = (const char*)&res.fs->activeTileCount; cap = 1; break;
        case GSM_DBG_SORTED_TILE_IDS: src = (const char*)res.tileIds[0]; cap = res.maxInstances; break;
        case GSM_DBG_TILE_BOUNDS: src = (const char*)res.bounds; cap = res.maxGaussians; break;
        case GSM_DBG_SORTED_PRIMITIVE_INDICES: src = (const char*)res.primIdx[0]; cap = res.maxGaussians; break;
        case GSM_DBG_INSTANCE_OFFSETS: src = (const char*)res.offsets; cap = res.maxGaussians; break;
        case GSM_DBG_N_TOUCHED_TILES: src = (const char*)res.nTouched; cap = res.maxGaussians; break;
        case GSM_DBG_INSTANCE_GAUSSIAN_INDICES: src = (const char*)res.instIdx[0]; cap = res.maxInstances; break;
        case GSM_DBG_DEPTH_KEYS: src = (const char*)res.depthKeys[0]; cap = res.maxGaussians; break;
        case GSM_DBG_RENDER_DATA: src = (const char*)res.renderData; cap = res.maxGaussians; break;
        case GSM_DBG_TILE_HEADERS: src = (const char*)res.tileHeaders; cap = res.maxTiles; break;
        case GSM_DBG_ACTIVE_TILES: src = (const char*)res.activeTiles; cap = res.maxTiles; break;
        case GSM_DBG_SCRATCH_DEPTH_KEYS: src = (const char*)res.depthKeys[1]; cap = res.maxGaussians; break;
        case GSM_DBG_SCRATCH_PRIMITIVE_INDICES: src = (const char*)res.primIdx[1]; cap = res.maxGaussians; break;
        case GSM_DBG_DEPTH_SORT_PLAN: src = (const char*)res.depthPlan; cap = 1; break;
        default: return fail(GSM_ERR_INVALID_ARGUMENT, "unknown debug buffer");
    }
    if (first > cap || count > cap - first) return fail(GSM_ERR_INVALID_ARGUMENT, "debug read out of range");
    cudaStream_t s = (cudaStream_t)stream;
    GSM_CUDA(cudaMemcpyAsync(dst, src + first * es, count * es, cudaMemcpyDeviceToHost, s), "debug read");
    GSM_CUDA(cudaStreamSynchronize(s), "debug read sync");
    return GSM_OK;
}

// ---- GlobalRenderer (GlobalRenderer.swift:72-372): the same handle, limits and precision; 32 x 16 tiles of the LIMITS
static gsm_status ensureGlobalResources(gsm_renderer* r) {
    if (r->globalArena) return GSM_OK;
    const gsm_config& c = r->cfg;
    const uint32_t G = c.maxGaussians < 1 ? 1 : c.maxGaussians;
    const uint32_t A = 4u * G;   // GlobalResources.swift:79-81
    const uint32_t tileW = 32, tileH = 16;   // GlobalRenderer.swift:74-75
    const uint32_t tilesX = (c.maxWidth + tileW - 1) / tileW, tilesY = (c.maxHeight + tileH - 1) / tileH;
    const uint32_t T = tilesX * tilesY < 1 ? 1 : tilesX * tilesY;
    if (T > 65536u) return fail(GSM_ERR_INVALID_TILE_COUNT, "GlobalRenderer sort keys hold 16 bits of tile id");
    const SortScratchLayout L = sortScratchLayout(A, 32, 4);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = alignUp(off + bytes, 256); return o; };
    const size_t oRender = take((size_t)G * 16), oBounds = take((size_t)G * 16), oFlags = take((size_t)G * 4), oFlagOff = take((size_t)G * 4);
    const size_t oVisible = take((size_t)G * 4), oCounts = take((size_t)G * 4), oOffsets = take((size_t)G * 4);
    const size_t oBlockSums = take(((size_t)G / 1024 + 8) * 4);
    const size_t oKeys = take((size_t)A * 4), oIdx = take((size_t)A * 4);
    const size_t oHeaders = take((size_t)T * 8), oActive = take((size_t)T * 4), oHeader = take(sizeof(GlobalHeader));
    const size_t oSort = take(L.total);
    const size_t oBlend = take((size_t)G * sizeof(BlendSplat)), oTicket = take(256), oMask = take((size_t)G * 8);
    cudaError_t e = cudaMalloc((void**)&r->globalArena, off);
    if (e != cudaSuccess) { r->globalArena = nullptr; return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "GlobalRenderer arena", e); }
    e = cudaMemset(r->globalArena, 0, off);
    if (e != cudaSuccess) return fail(GSM_ERR_RENDER_FAILED, "GlobalRenderer arena memset", e);
    char* a = r->globalArena;
    GlobalFrame& f = r->globalFrame;
    f.renderData = (uint4*)(a + oRender); f.bounds = (int4*)(a + oBounds); f.flags = (uint32_t*)(a + oFlags);
    f.flagOffsets = (uint32_t*)(a + oFlagOff); f.visibleIndices = (uint32_t*)(a + oVisible); f.counts = (uint32_t*)(a + oCounts);
    f.offsets = (uint32_t*)(a + oOffsets); f.blockSums = (uint32_t*)(a + oBlockSums); f.sortKeys = (uint32_t*)(a + oKeys);
    f.sortedIndices = (int32_t*)(a + oIdx); f.tileHeaders = (GSMGaussianHeader*)(a + oHeaders); f.activeTiles = (uint32_t*)(a + oActive);
    f.header = (GlobalHeader*)(a + oHeader);
    f.blendSplats = (BlendSplat*)(a + oBlend); f.renderTicket = (uint32_t*)(a + oTicket); f.hitMask = (uint2*)(a + oMask);
    f.capGaussians = G; f.maxAssignments = A; f.tileW = tileW; f.tileH = tileH; f.tilesX = tilesX; f.tilesY = tilesY;
    r->globalSortScratch = a + oSort;
    return GSM_OK;
}

gsm_status gsm_render_global(gsm_renderer* r, void* stream, void* color, void* depth, const void* gaussians, const void* harmonics,
                             uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height) {
    if (!r || !camera) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (gaussianCount > r->cfg.maxGaussians) return GSM_OK;   // validateLimits fails: the frame is not encoded (GlobalRenderer.swift:293-297)
    gsm_status st = validateFrame(r, width, height, gaussians, harmonics, color);
    if (st != GSM_OK) return st;
    if (gaussianCount == 0) return GSM_OK;
    DeviceGuard guard(r->device);
    st = ensureGlobalResources(r);
    if (st != GSM_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    const GlobalFrame& f = r->globalFrame;
    r->lastGlobal = true;
    MonoCam mc;
    fillMonoCam(mc, camera, gaussianCount, shComponents, width, height, r->cfg.gaussianColorSpace == GSM_COLORSPACE_SRGB);
    // 1. project + cull (+ visibility flags)  2. compaction in gid order  3. tiles per visible Gaussian  4. offsets + totals
    // 5. scatter keys and indices  6. one stable sort of the 32-bit keys  7. tile headers + active list  8. clear + render
    GSM_CUDA(launchGlobalProject(s, r->cfg.precision == GSM_PRECISION_FLOAT16, gaussians, harmonics, mc, f), "global project+cull");
    GSM_CUDA(launchGlobalCompact(s, f, gaussianCount), "global visibility compaction");
    GSM_CUDA(launchGlobalTileCount(s, f, gaussianCount), "global tile count");
    GSM_CUDA(launchGlobalAssignOffsets(s, f, gaussianCount), "global assignment offsets");
    {
        const SortScratchLayout L = sortScratchLayout(f.maxAssignments, 32, 4);
        char* scratch = r->globalSortScratch;
        GSM_CUDA(cudaMemsetAsync(scratch, 0, L.oK1, s), "global sort state memset");   // histograms, status words, tickets
        // the scatter writes the keys and counts their digits: no histogram kernel in front of the sort
        GSM_CUDA(launchGlobalTileScatter(s, f, gaussianCount, (uint32_t*)(scratch + L.oHist)), "global tile scatter");
        SortPlan p;
        p.k0 = f.sortKeys; p.k1 = scratch + L.oK1; p.v0 = (uint32_t*)f.sortedIndices; p.v1 = (uint32_t*)(scratch + L.oV1);
        p.countPtr = &f.header->totalAssignments; p.countCap = f.maxAssignments;
        p.hist = (uint32_t*)(scratch + L.oHist); p.status = (uint32_t*)(scratch + L.oStatus); p.gstatus = (uint32_t*)(scratch + L.oGStatus);
        p.tickets = (uint32_t*)(scratch + L.oTickets);
        p.tilesCap = L.tiles; p.keyBits = 32; p.numPasses = 4; p.numSMs = r->numSMs; p.histogramReady = true; p.largeTiles = L.large;
        GSM_CUDA(launchSort(s, p), "global sort");   // RadixSortEncoder.swift:52-63 sorts 3 or 4 bytes: the same order
    }
    GSM_CUDA(launchGlobalHeaders(s, f), "global tile headers");
    GSM_CUDA(launchGlobalRender(s, f, width, height, r->cfg.maxWidth, r->cfg.maxHeight, (__half*)color, (__half*)depth, r->numSMs), "global render");
    return GSM_OK;
}

gsm_status gsm_global_debug_read(gsm_renderer* r, void* stream, int which, void* dst, size_t first, size_t count) {
    if (!r || !dst) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (!r->globalArena) return fail(GSM_ERR_INVALID_ARGUMENT, "no global frame has been rendered");
    DeviceGuard guard(r->device);
    const GlobalFrame& f = r->globalFrame;
    const char* src = nullptr;
    size_t es = 0, cap = 0;
    const size_t T = (size_t)f.tilesX * f.tilesY;
    switch (which) {
        case GSM_GDBG_HEADER: src = (const char*)f.header; es = sizeof(GlobalHeader); cap = 1; break;
        case GSM_GDBG_SORTED_KEYS: src = (const char*)f.sortKeys; es = 4; cap = f.maxAssignments; break;
        case GSM_GDBG_SORTED_INDICES: src = (const char*)f.sortedIndices; es = 4; cap = f.maxAssignments; break;
        case GSM_GDBG_TILE_HEADERS: src = (const char*)f.tileHeaders; es = 8; cap = T; break;
        case GSM_GDBG_BOUNDS: src = (const char*)f.bounds; es = 16; cap = f.capGaussians; break;
        case GSM_GDBG_RENDER_DATA: src = (const char*)f.renderData; es = 16; cap = f.capGaussians; break;
        case GSM_GDBG_VISIBLE_INDICES: src = (const char*)f.visibleIndices; es = 4; cap = f.capGaussians; break;
        case GSM_GDBG_ACTIVE_TILES: src = (const char*)f.activeTiles; es = 4; cap = T; break;
        default: return fail(GSM_ERR_INVALID_ARGUMENT, "unknown global debug buffer");
    }
    if (first > cap || count > cap - first) return fail(GSM_ERR_INVALID_ARGUMENT, "debug read out of range");
    cudaStream_t s = (cudaStream_t)stream;
    GSM_CUDA(cudaMemcpyAsync(dst, src + first * es, count * es, cudaMemcpyDeviceToHost, s), "global debug read");
    GSM_CUDA(cudaStreamSynchronize(s), "global debug read sync");
    return GSM_OK;
}

gsm_status gsm_buffer_alloc(int device, size_t bytes, void** out) {
    if (!out) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (device < 0) cudaGetDevice(&device);
    DeviceGuard guard(device);
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(GSM_ERR_FAILED_TO_ALLOCATE_BUFFER, "cudaMalloc", e);
    return GSM_OK;
}
gsm_status gsm_buffer_free(void* p) {
    if (p) GSM_CUDA(cudaFree(p), "cudaFree");
    return GSM_OK;
}
gsm_status gsm_buffer_upload(void* dst, const void* src, size_t bytes, void* stream) {
    GSM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "upload");
    return GSM_OK;
}
gsm_status gsm_buffer_download(void* dst, const void* src, size_t bytes, void* stream) {
    GSM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "download");
    GSM_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "download sync");
    return GSM_OK;
}
gsm_status gsm_stream_create(int device, void** out) {
    if (!out) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (device < 0) cudaGetDevice(&device);
    DeviceGuard guard(device);
    cudaStream_t s;
    GSM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "stream create");
    *out = (void*)s;
    return GSM_OK;
}
gsm_status gsm_stream_synchronize(void* stream) {
    GSM_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "stream sync");
    return GSM_OK;
}
gsm_status gsm_stream_destroy(void* stream) {
    GSM_CUDA(cudaStreamDestroy((cudaStream_t)stream), "stream destroy");
    return GSM_OK;
}

gsm_status gsm_sort_pairs(gsm_renderer* r, void* stream, void* keys, void* payload, uint32_t count, int keyBits, int numPasses) {
    if (!r || !keys || !payload) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if ((keyBits != 16 && keyBits != 32) || numPasses < 1 || numPasses > keyBits / 8) return fail(GSM_ERR_INVALID_ARGUMENT, "bad key width / pass count");
    if (count == 0) return GSM_OK;
    DeviceGuard guard(r->device);
    return gsm::sortPairsStandalone((cudaStream_t)stream, r->numSMs, keys, payload, count, keyBits, numPasses);
}

size_t gsm_sort_pairs_scratch_bytes(uint32_t count, int keyBits, int numPasses) {
    if ((keyBits != 16 && keyBits != 32) || numPasses < 1 || numPasses > keyBits / 8 || count == 0) return 0;
    return gsm::sortScratchBytes(count, keyBits, numPasses);
}

gsm_status gsm_sort_pairs_with_scratch(gsm_renderer* r, void* stream, void* keys, void* payload, uint32_t count, int keyBits,
                                       int numPasses, void* scratch) {
    if (!r || !keys || !payload || !scratch) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if ((keyBits != 16 && keyBits != 32) || numPasses < 1 || numPasses > keyBits / 8) return fail(GSM_ERR_INVALID_ARGUMENT, "bad key width / pass count");
    if (count == 0) return GSM_OK;
    if (((uintptr_t)scratch & 255u) != 0) return fail(GSM_ERR_INVALID_ARGUMENT, "scratch must be 256-byte aligned");
    DeviceGuard guard(r->device);
    return gsm::sortPairsStandalone((cudaStream_t)stream, r->numSMs, keys, payload, count, keyBits, numPasses, scratch);
}

gsm_status gsm_probe_math(int device, int op, const void* a, const void* b, void* out, uint32_t n) {
    if (!a || !out || n == 0) return fail(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (device < 0) cudaGetDevice(&device);
    DeviceGuard guard(device);
    const size_t inEl = (op == 11) ? 6 : ((op == 5 || op == 7 || op == 8 || op == 12 || op == 13 || (op >= 14 && op <= 17)) ? 2 : 4);
    const size_t outEl = (op == 5 || op == 6 || op == 7 || op == 8 || op == 11 || op == 12 || op == 13 || (op >= 14 && op <= 17)) ? 2 : 4;
    void *da = nullptr, *db = nullptr, *dout = nullptr;
    gsm_status st = GSM_OK;
    cudaError_t e;
    do {
        if ((e = cudaMalloc(&da, n * inEl)) != cudaSuccess) break;
        if ((e = cudaMalloc(&dout, n * outEl)) != cudaSuccess) break;
        if ((e = cudaMemcpy(da, a, n * inEl, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (b) {
            if ((e = cudaMalloc(&db, n * inEl)) != cudaSuccess) break;
            if ((e = cudaMemcpy(db, b, n * inEl, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        }
        if (op == 13) {  // the blend's table form of exp(-0.5h * p): table built exactly as gsm_renderer_create builds it
            unsigned short* tab = nullptr;
            if ((e = cudaMalloc((void**)&tab, blendExpTableBytes())) != cudaSuccess) break;
            e = buildBlendExpTable(nullptr, tab);
            if (e == cudaSuccess) e = launchBlendExpProbe(nullptr, tab, (const unsigned short*)da, (unsigned short*)dout, n);
            if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
            cudaFree(tab);
            if (e != cudaSuccess) break;
        } else if ((e = launchProbe(nullptr, op, da, db, dout, n)) != cudaSuccess) break;
        e = cudaMemcpy(out, dout, n * outEl, cudaMemcpyDeviceToHost);
    } while (false);
    if (e != cudaSuccess) st = fail(GSM_ERR_RENDER_FAILED, "gsm_probe_math", e);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return st;
}

int gsm_blend_exp_mode(int device) {
    if (device < 0) cudaGetDevice(&device);
    return blendExpMode(device);
}

const char* gsm_status_string(gsm_status s) {
    switch (s) {
        case GSM_OK: return "ok";
        case GSM_ERR_DEVICE_NOT_AVAILABLE: return "CUDA device not available";
        case GSM_ERR_FAILED_TO_CREATE_PIPELINE: return "failed to create pipeline";
        case GSM_ERR_FAILED_TO_ALLOCATE_BUFFER: return "failed to allocate buffer";
        case GSM_ERR_INVALID_GAUSSIAN_COUNT: return "Gaussian count exceeds maximum";
        case GSM_ERR_INVALID_DIMENSIONS: return "dimensions exceed maximum";
        case GSM_ERR_INVALID_TILE_COUNT: return "tile count exceeds maximum";
        case GSM_ERR_RENDER_FAILED: return "render failed";
        case GSM_ERR_INVALID_ARGUMENT: return "invalid argument";
    }
    switch ((int)s) {  // PLYLoaderError.errorDescription (PLYLoader.swift:218-241), statuses of gsm_scene.h
        case 20: return "Invalid PLY header";
        case 21: return "Unsupported PLY format. Only binary_little_endian is supported.";
        case 22: return "No 'vertex' element found in PLY";
        case 23: return "Missing required properties";
        case 24: return "List properties in vertex element are not supported";
        case 25: return "PLY file has insufficient data for declared vertex count";
        case 26: return "Compressed PLY requires 'chunk' element with bounding box data";
    }
    return "unknown";
}
const char* gsm_last_error_string(void) { return g_lastError.c_str(); }
int gsm_abi_version(void) { return GSM_ABI_VERSION; }

}  // extern "C"
