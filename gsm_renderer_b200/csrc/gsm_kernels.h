// gsm_kernels.h -- host-callable launchers of the stage kernels (internal).
#pragma once
#include "gsm_common.cuh"

namespace gsm {

struct ProjectOut {
    FrameState* fs;
    unsigned long long* status;   // prefix words, one per tile (32-gid warp tiles in the strip ingest, 2048-gid tiles in the compaction)
    unsigned long long* statusGroups;  // one word per group of 32 tiles
    void* renderData;             // GSMGaussianRenderData[] or GSMStereoTiledRenderData[]
    int32_t* bounds;              // int4 per Gaussian
    uint32_t* nTouched;
    uint2* hitMask;               // mono only: hit bits of the first 64 AABB tiles per Gaussian (gsm_tiletest.cuh)
    BlendSplat* blendSplats;      // mono only (may be null)
    const uint32_t* recTouched = nullptr;  // strip ingest: the compaction runs over gathered records (count, key, gid per record)
    const uint32_t* recKey = nullptr;
    const uint32_t* recGid = nullptr;
    const uint32_t* countPtr = nullptr;    // if set, the compaction reads its element count from the device (routed records, group.cu)
    uint32_t* preDepthKeys;       // per gid, before compaction (the reference aliases the sort scratch for it, DFR.swift:282)
    uint32_t* depthKeys;          // compacted, ascending gid
    int32_t* primitiveIndices;
    uint4* zeroBase = nullptr;    // if set, the projection kernel clears the frame's zero region (zeroVecs 16-byte words) itself
    size_t zeroVecs = 0;
    uint32_t maxOut;              // capacity of the compacted arrays (maxGaussians)
    GSMDepthFirstHeader* header;  // if set, the compaction's last tile also writes the frame header (DFS.metal:2184-2203)
    uint32_t maxInstances;
    uint32_t depthKey16;
    uint32_t gidFirst;
    // fused into the compaction kernel: digit histograms of the compacted depth keys
    uint32_t* depthHist;          // [4][256], zeroed with the frame state
    uint32_t depthPasses;
    // depth sort as bucket scatter + local sort (bucketsort.cu): the projection records the key range and the compaction counts
    // a sample of the keys per fine bin of that range. Null: the LSD passes run (large frames, strip ingest, 16-bit depth keys).
    KeyRange* keyRange = nullptr;
};

int shDegreeFromComponents(uint32_t n);
const void* kernel_image_probe();  // a kernel symbol, to test that the sm_100a image loads
cudaError_t launchProjectMono(cudaStream_t s, bool halfInput, const void* g, const void* h, const MonoCam& cam, const ProjectOut& o);
cudaError_t launchProjectStereo(cudaStream_t s, bool halfInput, const void* g, const void* h, const StereoCam& cam, const ProjectOut& o);
cudaError_t launchCompactVisible(cudaStream_t s, uint32_t N, const ProjectOut& o, int numSMs);

// Onesweep radix sort (sort.cu). keys/vals ping-pong between (k0,v0) and (k1,v1); after numPasses the
// result is in (k0,v0) if numPasses is even, else it is copied back. countPtr is read on the device.
// Rows of 256 words of per-group state one sort pass owns for `tiles` tiles: one row of per-digit sums per group of 16 tiles,
// followed by the groups' arrival masks (one word per group, see onesweep_pass_kernel).
__host__ __device__ inline uint32_t sortGroupRows(uint32_t tiles) {
    const uint32_t groups = (tiles + 15u) / 16u;
    return groups + (groups + 255u) / 256u;
}

struct SortPlan {
    void* k0; void* k1; uint32_t* v0; uint32_t* v1;
    const uint32_t* countPtr; uint32_t countCap;
    uint32_t* hist;      // [numPasses][256] zeroed
    uint32_t* status;    // [numPasses][tilesCap][256] per-tile count / look-back words (zeroed with the frame state, or by the histogram kernel)
    uint32_t* gstatus;   // [numPasses][sortGroupRows(tilesCap)][256] per-group words
    uint32_t* tickets;   // [numPasses] zeroed
    uint32_t tilesCap;
    int keyBits;         // 16 or 32
    int numPasses;
    int numSMs;
    bool largeTiles;     // 32-bit keys only: 4096-key tiles (large inputs) instead of 2048
    const uint32_t* gatherSrc = nullptr;  // optional: the last pass also writes gatherDst[i] = gatherSrc[sorted payload i]
    uint32_t* gatherDst = nullptr;
    bool histogramReady; // hist filled and status/gstatus zeroed by earlier kernels of the frame (fused); else a histogram kernel runs
    int shift0 = 0;      // digit of pass p = (key >> (shift0 + 8p)) & 0xFF
    bool leaveInScratch = false;  // odd pass counts: do not copy the result back from (k1, v1) (the MSD tile sort's local pass reads it there)
    bool pairTiles = false;       // one 16-bit pass left in scratch: two consecutive tiles per CTA and one count exchange (onesweep_pair_kernel)
};
uint32_t sortTileSize(int keyBits, bool large);
cudaError_t launchSort(cudaStream_t s, const SortPlan& p);

// The frame's depth sort as bucket scatter + local sort (bucketsort.cu), for frames bucketSortCovers() accepts; the compaction
// kernel must have run with ProjectOut::plan / keyRange set. Result in (k0, v0), like launchSort.
struct BucketSortPlan {
    uint32_t* k0; uint32_t* k1; uint32_t* v0; uint32_t* v1;
    const uint32_t* countPtr; uint32_t countCap;
    DepthPlan* plan;      // bucket offsets, written by the scatter kernel for the local pass
    const uint32_t* fineHist;  // FrameState::fineHist, the compaction's sample of the keys per fine bin
    KeyRange* keyRange;   // cleared for the next frame
    uint32_t* status;     // >= tiles * 512 zeroed words (the LSD passes' status rows: a frame uses one path or the other)
    uint32_t* gstatus;    // >= groups * 512 + groups zeroed words
    uint32_t* place;      // countCap words of scratch: bucket and tile-local position per key, between the rank and the scatter kernel
    const uint32_t* gatherSrc; uint32_t* gatherDst;
    int numSMs;
};
bool bucketSortCovers(uint32_t maxKeys, int numSMs);
cudaError_t bucketSortPrepareDevice();
cudaError_t launchBucketSort(cudaStream_t s, const BucketSortPlan& p);


// instance expansion (expand.cu)
cudaError_t launchCreateInstances(cudaStream_t s, bool stereo, bool tileId16, const int32_t* sortedIdx, const uint32_t* sortedTouched,
                                  const uint2* hitMask, uint32_t* offsets, unsigned long long* scanStatus, unsigned long long* scanGroups, uint32_t* ticket, const int32_t* bounds, const void* renderData, void* tileIds, int32_t* instanceIdx,
                                  const GSMDepthFirstHeader* header, uint32_t tilesX, uint32_t maxAssignments, uint32_t capVisible,
                                  uint32_t* tileHist, uint32_t tilePasses, int numSMs,
                                  uint32_t msdShift = 0xFFFFFFFFu);  // != ~0: histogram of (tileId >> msdShift) & 0xFF instead (tilesort.cu)

// MSD tile sort, second half (tilesort.cu): after ONE onesweep pass with shift0 = lowBits on the ids' high byte, one CTA per
// bucket finishes the stable sort on the low bits into (keysOut, valsOut) and writes the tile ranges (no range kernel).
uint32_t tileSortLowBits(uint32_t tileCount);
cudaError_t launchTileLocalSort(cudaStream_t s, const void* keysIn, const uint32_t* valsIn, void* keysOut, uint32_t* valsOut,
                                const uint32_t* bucketHist, const GSMDepthFirstHeader* header, uint32_t capInstances, uint32_t lowBits,
                                uint32_t tileCount, uint32_t* lowerBounds, uint32_t* chunkCounts, int numSMs);
// chunkCounts: (capInstances / 4096 + 257) rows of 2^lowBits words of scratch

// tile ranges (ranges.cu)
cudaError_t launchTileRanges(cudaStream_t s, bool tileId16, const void* sortedTileIds, const GSMDepthFirstHeader* header,
                             uint32_t tileCount, uint32_t* lowerBounds, uint32_t capInstances, uint32_t tileLo = 0u,
                             uint32_t tileHi = 0xFFFFFFFFu);  // only lowerBounds[tileLo .. tileHi] are written

// blend (blend.cu). Each tile's CTA also writes its GaussianHeader and appends itself to the active list.
struct TileOut { GSMGaussianHeader* tileHeaders; uint32_t* activeTiles; uint32_t* activeTileCount; };
cudaError_t launchBlendMono(cudaStream_t s, const uint32_t* lowerBounds, const BlendSplat* splats, const int32_t* instanceIdx,
                            uint32_t width, uint32_t height, uint32_t tilesX, uint32_t tilesY, uint32_t tileRowFirst,
                            uint32_t tileRowCount, __half* color, __half* depth, TileOut tout, const unsigned short* expTable,
                            uint32_t* ticket, int numSMs);
// exact table of the blend's exp(-0.5h * p) over the non-negative halfs (built once per renderer from the canonical function)
cudaError_t blendExpSelfTest(int device);   // decides, once per device, which form of exp(-0.5h * p) the blend kernels run
int blendExpMode(int device);                // 0 not decided (polynomial), 1 polynomial, 2 tuned XU-pipe (MUFU.EX2) form
size_t blendExpTableBytes();
bool blendUsesExpTable();   // false in the default build: the polynomial is faster (blend.cu, GSM_BLEND_TABLE)
cudaError_t buildBlendExpTable(cudaStream_t s, unsigned short* table);
cudaError_t launchBlendExpProbe(cudaStream_t s, const unsigned short* table, const unsigned short* in, unsigned short* out, uint32_t n);
cudaError_t launchBlendStereo(cudaStream_t s, const uint32_t* lowerBounds, const GSMStereoTiledRenderData* splats,
                              const int32_t* instanceIdx, uint32_t width, uint32_t height, uint32_t tilesX, uint32_t tilesY,
                              __half* dstSideBySide, int eyeMask, int flipY, TileOut tout);

// foveated stereo copy (stereo_copy.cu): DepthFirstStereoCopyEncoder.swift:28-100 as one resampling kernel
constexpr uint32_t kSrgbTableSize = 0x3C01u;  // one byte per half in [0, 1]
struct StereoCopyParams {
    const __half* src;            // intermediate rgba16f; eye e at src + e * srcEyeStride, rows srcRowStride halfs apart
    size_t srcEyeStride, srcRowStride;
    uint32_t width, height;       // per-eye size of the intermediate image
    void* dst;
    uint32_t textureWidth, textureHeight, arrayLength;
    size_t rowBytes, sliceBytes;
    uint32_t format;              // gsm_pixel_format
    int flipY;
    uint32_t layerCount;          // 0 = no rate map
    uint32_t physicalWidth[2], physicalHeight[2];
    const float* screenX[2];      // device
    const float* screenY[2];
    float vp[2][4];               // {originX, originY, width, height} per eye
    const uint8_t* srgbLut;       // device, kSrgbTableSize bytes
};
void buildSrgbEncodeTable(uint8_t* table);
cudaError_t launchStereoCopy(cudaStream_t s, const StereoCopyParams& p);

// error reporting shared with scene.cu (thread-local message behind gsm_last_error_string)
gsm_status reportFailure(gsm_status s, const char* what, cudaError_t e = cudaSuccess);
// callerScratch: optional device buffer of sortScratchBytes(...) bytes (then nothing is allocated and the stream is not synchronised)
size_t sortScratchBytes(uint32_t count, int keyBits, int numPasses);
gsm_status sortPairsStandalone(cudaStream_t s, int numSMs, void* keys, void* payload, uint32_t count, int keyBits, int numPasses,
                               void* callerScratch = nullptr);

// strip-sharded frame (strip.cu)
cudaError_t launchPackRecords(cudaStream_t s, const FrameState* fs, const uint32_t* keys, const int32_t* gids, const void* renderData,
                              const int32_t* bounds, const uint2* hitMask, void* out, uint32_t cap, int numSMs);
cudaError_t launchIngestRecords(cudaStream_t s, const void* records, uint32_t recordCount, uint32_t rowFirst, uint32_t rowCount,
                                const ProjectOut& o, uint32_t* recTouched, uint32_t* recKey, uint32_t* recGid);

// strip-sharded frame over peer memory (group.cu): the exchange is the routing kernel's own stores into the peers' windows
constexpr uint32_t kGroupMaxRanks = 8;
struct GroupMailbox {                      // head of every rank's exchange window; written by peers over NVLink
    uint32_t recordCount[kGroupMaxRanks];  // [s]: records source s routed to this rank ...
    uint32_t recordSeq[kGroupMaxRanks];    // [s]: ... for frame recordSeq[s] (release-stored after the records and the count)
    uint32_t ack[kGroupMaxRanks];          // [d]: destination d has consumed this rank's records of frame ack[d]
    uint32_t frameDone[kGroupMaxRanks];    // [s]: rank s's strip / eye of frame frameDone[s] is in this rank's image
};
struct RouteParams {
    uint32_t world, rank, seq, regionCap;     // regionCap: records per (destination, source) region
    uint32_t rowStart[kGroupMaxRanks + 1];    // strip d = tile rows [rowStart[d], rowStart[d+1])
    SplatRecord* region[kGroupMaxRanks];      // [d]: THIS rank's region inside destination d's window (peer-mapped)
    GroupMailbox* mailbox[kGroupMaxRanks];    // [d]: destination d's mailbox (peer-mapped)
    GroupMailbox* mine;
};
struct IngestParams {
    uint32_t world, rank, seq, regionCap;
    int rowFirst, rowLast;
    const SplatRecord* region[kGroupMaxRanks];  // [s]: source s's region inside THIS rank's window (local memory)
    GroupMailbox* mailbox[kGroupMaxRanks];      // [s]: source s's mailbox (peer-mapped), for the ack
    GroupMailbox* mine;
};
uint32_t routeStatusWords(uint32_t maxRecords);
cudaError_t launchRouteRecords(cudaStream_t s, FrameState* fs, const uint32_t* keys, const int32_t* gids, const void* renderData,
                               const int32_t* bounds, const uint2* hitMask, uint32_t cap, uint32_t* status, const RouteParams& P,
                               int numSMs);
cudaError_t launchIngestRouted(cudaStream_t s, const IngestParams& P, const ProjectOut& o, uint32_t* recTouched, uint32_t* recKey,
                               uint32_t* recGid, int numSMs);
cudaError_t launchGroupSignal(cudaStream_t s, GroupMailbox* to, uint32_t rank, uint32_t seq);
cudaError_t launchGroupWait(cudaStream_t s, const GroupMailbox* mine, uint32_t mask, uint32_t seq);

// ---- GlobalRenderer (global.cu + kernels next to the helpers they share in project.cu / blend.cu): GlobalShaders.metal on the
// DepthFirst path's primitives. One arena per renderer; every count stays on the device.
struct GlobalHeader {               // TileAssignmentHeader (BridgingTypes.h) + the frame's other device-side counts
    uint32_t totalAssignments, paddedCount, overflow, visibleCount, activeTileCount, totalRaw, _pad[2];
};
struct GlobalFrame {
    uint4* renderData;              // GSMGaussianRenderData per Gaussian
    int4* bounds;
    uint32_t* flags;                // 1 = valid bounds (markVisibilityKernel)
    uint32_t* flagOffsets;          // exclusive prefix of flags
    uint32_t* visibleIndices;
    uint32_t* counts;               // tiles per visible Gaussian
    uint32_t* offsets;              // exclusive prefix of counts (unclamped)
    uint32_t* blockSums;            // scan scratch
    uint32_t* sortKeys;             // [tile:16][half depth ^ 0x8000:16]
    int32_t* sortedIndices;
    GSMGaussianHeader* tileHeaders;
    uint32_t* activeTiles;
    GlobalHeader* header;
    BlendSplat* blendSplats;        // the render's 32-byte record per Gaussian (conic and colours as halfs), written by the tile count
    uint2* hitMask;                 // per Gaussian: the tiles of an AABB of <= 64 tiles that pass the ellipse test (count pass -> scatter pass)
    uint32_t* renderTicket;         // tile ticket of the persistent render kernel, zeroed by the totals kernel
    uint32_t capGaussians, maxAssignments, tileW, tileH, tilesX, tilesY;
};
cudaError_t launchGlobalProject(cudaStream_t s, bool halfInput, const void* g, const void* h, const MonoCam& cam, const GlobalFrame& f);
cudaError_t launchGlobalCompact(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount);        // flags -> visibleIndices, visibleCount
cudaError_t launchGlobalTileCount(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount);
cudaError_t launchGlobalAssignOffsets(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount);  // counts -> offsets, header totals
cudaError_t launchGlobalTileScatter(cudaStream_t s, const GlobalFrame& f, uint32_t gaussianCount, uint32_t* sortHist);  // also accumulates the sort's 4 x 256 digit histograms (zeroed by the caller)
cudaError_t launchGlobalHeaders(cudaStream_t s, const GlobalFrame& f);
cudaError_t launchGlobalRender(cudaStream_t s, const GlobalFrame& f, uint32_t width, uint32_t height, uint32_t maxWidth, uint32_t maxHeight,
                               __half* color, __half* depth, int numSMs);

// math probes (probe.cu)
cudaError_t launchProbe(cudaStream_t s, int op, const void* a, const void* b, void* out, uint32_t n);

}  // namespace gsm
