/*
 * gsm.h -- C ABI of the B200-native DepthFirstRenderer path (libgsm_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types. Each entry point
 * cites the reference interface it replaces (paths relative to the reference repo):
 *   GRP.swift = Sources/Renderer/Shared/GaussianRendererProtocol.swift
 *   DFR.swift = Sources/Renderer/DepthFirstRenderer/DepthFirstRenderer.swift
 *   DFUT.swift = Tests/RendererTests/DepthFirstUnitTests.swift
 *
 * Object mapping:  MTLDevice -> CUDA device ordinal;  MTLCommandBuffer -> cudaStream_t (passed as void*;
 * enqueue-only, the caller synchronises, exactly like commit()/waitUntilCompleted());  MTLBuffer -> device
 * pointer;  MTLTexture rgba16Float -> device buffer of W*H*4 halfs, row pitch W*8 B;  r16Float -> W*H halfs.
 * Threading: one host thread per renderer handle (the reference class is @unchecked Sendable with
 * unsynchronised caches, DFR.swift:6,29-31); all work is stream-ordered; gsm_render* never synchronises
 * after the handle's first use and is safe to capture into a CUDA graph from the second call on.
 */
#ifndef GSM_H
#define GSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSM_ABI_VERSION 2

/* RendererError (GRP.swift:274-324), one code per case that can arise on this path. */
typedef enum {
    GSM_OK = 0,
    GSM_ERR_DEVICE_NOT_AVAILABLE = 1,    /* .deviceNotAvailable */
    GSM_ERR_FAILED_TO_CREATE_PIPELINE = 2, /* .failedToCreatePipeline: kernels for sm_100a cannot load */
    GSM_ERR_FAILED_TO_ALLOCATE_BUFFER = 3, /* .failedToAllocateBuffer */
    GSM_ERR_INVALID_GAUSSIAN_COUNT = 4,  /* .invalidGaussianCount: maxGaussians > 30,000,000 (DFR.swift:51-56) */
    GSM_ERR_INVALID_DIMENSIONS = 5,      /* .invalidDimensions */
    GSM_ERR_INVALID_TILE_COUNT = 6,      /* .invalidTileCount: > 65535 tiles with 16-bit tile ids */
    GSM_ERR_RENDER_FAILED = 7,           /* .renderFailed: a CUDA call failed; see gsm_last_error_string */
    GSM_ERR_INVALID_ARGUMENT = 8
} gsm_status;

typedef enum { GSM_PRECISION_FLOAT32 = 0, GSM_PRECISION_FLOAT16 = 1 } gsm_precision;   /* RenderPrecision, GRP.swift:4-7 */
typedef enum { GSM_COLORSPACE_LINEAR = 0, GSM_COLORSPACE_SRGB = 1 } gsm_color_space;   /* GRP.swift:196-201 */
typedef enum { GSM_KEY_BITS16 = 16, GSM_KEY_BITS32 = 32 } gsm_key_precision;           /* RadixSortKeyPrecision */

/* RendererConfig (GRP.swift:195-228) + the two init enums of DepthFirstRenderer.init (DFR.swift:45-50).
 * colorFormat and backToFront are ignored by the DepthFirst path (only HardwareRenderer reads them). */
typedef struct {
    uint32_t maxGaussians;       /* default 6,000,000 */
    uint32_t maxWidth;           /* default 1920 */
    uint32_t maxHeight;          /* default 1080 */
    uint32_t precision;          /* gsm_precision, default FLOAT16 */
    uint32_t gaussianColorSpace; /* gsm_color_space, default SRGB */
    uint32_t depthSortKeyPrecision; /* gsm_key_precision, default BITS32 */
    uint32_t tileIdPrecision;       /* gsm_key_precision, default BITS16 */
    int32_t device;              /* CUDA ordinal; -1 = current device */
    uint32_t stereoCopyFlipY;    /* 1 = literal stereoCopy semantics (rows flipped, quirk Q9); 0 = straight copy */
    uint32_t reserved[7];
} gsm_config;

/* CameraParams (GRP.swift:28-54). Matrices are simd_float4x4: column-major, m[4*col + row]. */
typedef struct {
    float viewMatrix[16];
    float projectionMatrix[16];
    float position[3];
    float focalX, focalY; /* carried, unused on this path */
    float nearPlane, farPlane; /* defaults 0.1, 10.0 */
} gsm_camera;

typedef struct gsm_renderer gsm_renderer;

void gsm_config_default(gsm_config* cfg);

/* DepthFirstRenderer.init (DFR.swift:45-101). Resources are allocated lazily on first render, like
 * ensureMonoResources / ensureStereoResources (DFR.swift:105-137). */
gsm_status gsm_renderer_create(const gsm_config* cfg, gsm_renderer** out);
void gsm_renderer_destroy(gsm_renderer* r);

/* GaussianRenderer.render (GRP.swift:248-256, DFR.swift:166-203). color: rgba16f W*H; depth: r16f W*H or
 * NULL. gaussians: PackedWorldGaussian[count] (FLOAT32) or PackedWorldGaussianHalf[count] (FLOAT16);
 * harmonics: float or half, planar per Gaussian. Returns GSM_OK and enqueues nothing when count == 0 or
 * count > maxGaussians (the reference's silent no-op, DFR.swift:249). */
gsm_status gsm_render(gsm_renderer* r, void* stream, void* color, void* depth, const void* gaussians,
                      const void* harmonics, uint32_t gaussianCount, uint32_t shComponents,
                      const gsm_camera* camera, uint32_t width, uint32_t height);

/* GaussianRenderer.renderStereo with StereoRenderTarget.sideBySide (GRP.swift:264-271, DFR.swift:205-235,
 * :469-512). colorSideBySide: rgba16f (2*width) x height; width/height are per eye. The depth texture of
 * the target is ignored by the reference (DFR.swift:472) and has no parameter here. */
gsm_status gsm_render_stereo(gsm_renderer* r, void* stream, void* colorSideBySide, const void* gaussians,
                             const void* harmonics, uint32_t gaussianCount, uint32_t shComponents,
                             const gsm_camera* leftEye, const gsm_camera* rightEye, uint32_t width,
                             uint32_t height);

/* One-eye-per-GPU split of the joint stereo frame (no reference counterpart): runs the identical
 * joint stages 1-7 and blends only the eyes in eyeMask (bit 0 = left, bit 1 = right) into their half of the
 * side-by-side target; the other half is left untouched. eyeMask 3 == gsm_render_stereo. */
gsm_status gsm_render_stereo_eyes(gsm_renderer* r, void* stream, void* colorSideBySide, const void* gaussians,
                                  const void* harmonics, uint32_t gaussianCount, uint32_t shComponents,
                                  const gsm_camera* leftEye, const gsm_camera* rightEye, uint32_t width,
                                  uint32_t height, uint32_t eyeMask);

/* ---- StereoRenderTarget.foveated (GRP.swift:168-193, :233-239; DFR.swift:516-551, :789-830;
 * DepthFirstStereoCopyEncoder.swift:28-100; stereoCopyVertex/Fragment DFS.metal:1984-2018) ----
 * The reference blends both eyes into an intermediate 2-slice rgba16f array and then draws one full-screen triangle
 * per eye into the drawable, inside that eye's viewport, through the drawable's MTLRasterizationRateMap; the fragment
 * samples the eye's slice with a linear, clamp-to-edge sampler. Here that draw is one resampling copy kernel. The rate
 * map, which Metal keeps opaque, crosses the boundary as what its own mapPhysicalToScreenCoordinates returns: per layer,
 * the screen-space x of the centre of every physical column and the screen-space y of the centre of every physical row
 * (the map is separable: MTLRasterizationRateLayerDescriptor has one horizontal and one vertical rate array). */
typedef enum {
    GSM_PIXEL_RGBA16F = 0,    /* .rgba16Float */
    GSM_PIXEL_BGRA8 = 1,      /* .bgra8Unorm */
    GSM_PIXEL_BGRA8_SRGB = 2, /* .bgra8Unorm_srgb, FoveatedStereoDrawable's default (GRP.swift:184) */
    GSM_PIXEL_RGBA8 = 3,      /* .rgba8Unorm */
    GSM_PIXEL_RGBA8_SRGB = 4  /* .rgba8Unorm_srgb */
} gsm_pixel_format;

/* MTLViewport; znear/zfar do not affect the copy. */
typedef struct { double originX, originY, width, height; } gsm_viewport;

/* EyeView (GRP.swift:68-97). */
typedef struct {
    gsm_viewport viewport;
    gsm_camera camera;
} gsm_eye_view;

/* StereoConfiguration (GRP.swift:100-117). sceneTransform: simd_float4x4, column-major. */
typedef struct {
    gsm_eye_view leftEye, rightEye;
    float sceneTransform[16];
} gsm_stereo_configuration;

/* One layer of a tabulated MTLRasterizationRateMap. screenX[i] / screenY[j]: screen coordinates of the centre of
 * physical column i / row j (HOST pointers, read during the call). */
typedef struct {
    uint32_t physicalWidth, physicalHeight;
    const float* screenX;
    const float* screenY;
} gsm_rate_map_layer;

typedef struct {
    uint32_t layerCount; /* 1 or 2; slice k of the drawable uses layer min(k, layerCount-1) */
    gsm_rate_map_layer layers[2];
} gsm_rate_map;

/* FoveatedStereoDrawable (GRP.swift:168-193). colorTexture: device memory, arrayLength slices of textureHeight rows of
 * rowBytes bytes, sliceBytes apart. arrayLength 2 = layered (left eye -> slice 0, right eye -> slice 1), 1 = shared (both
 * viewports in slice 0) (DepthFirstStereoCopyEncoder.swift:88-94). rasterizationRateMap NULL = no foveation. The depth
 * texture of the drawable is not written on this path (DFR.swift:516-551) and has no field here. */
typedef struct {
    void* colorTexture;
    uint32_t textureWidth, textureHeight, arrayLength;
    size_t rowBytes, sliceBytes;
    uint32_t colorPixelFormat; /* gsm_pixel_format */
    const gsm_rate_map* rasterizationRateMap;
} gsm_foveated_drawable;

/* GaussianRenderer.renderStereo with StereoRenderTarget.foveated (DFR.swift:225-233, :516-551). width/height are the
 * per-eye render size of the intermediate image. Texels of the drawable outside both viewports are left untouched. */
gsm_status gsm_render_stereo_foveated(gsm_renderer* r, void* stream, const gsm_foveated_drawable* drawable,
                                      const void* gaussians, const void* harmonics, uint32_t gaussianCount,
                                      uint32_t shComponents, const gsm_stereo_configuration* configuration,
                                      uint32_t width, uint32_t height);

/* The copy alone (step 10, DFR.swift:823-830): intermediate = rgba16f, (2*width) x height, left eye in columns
 * [0, width), right eye in [width, 2*width). Exposed for tests and for hosts that keep the intermediate image. */
gsm_status gsm_stereo_copy(gsm_renderer* r, void* stream, const void* intermediate, uint32_t width, uint32_t height,
                           const gsm_foveated_drawable* drawable, const gsm_viewport* leftViewport,
                           const gsm_viewport* rightViewport);

/* Same frame as gsm_render but with HOST buffers: copies inputs host->device, renders, copies the images
 * device->host and synchronises. This is what a host with no device allocator of its own calls (and what
 * bench.py times as `e2e`). hostDepth may be NULL. */
gsm_status gsm_render_host(gsm_renderer* r, const void* hostGaussians, const void* hostHarmonics,
                           uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* camera,
                           uint32_t width, uint32_t height, void* hostColor, void* hostDepth);

/* The reference's render() only ENCODES into the caller's MTLCommandBuffer (DFR.swift:196-330); the caller commits
 * and waits, so several frames can be in flight. gsm_render_host_async is that shape for host buffers: it enqueues
 * H2D + frame + D2H on the renderer's own stream and returns; gsm_render_host_wait blocks until the images are in
 * hostColor / hostDepth. The host buffers must stay valid (and should be pinned) until the wait. One frame in
 * flight per renderer: use two renderers to overlap a frame's upload with the previous frame's render + download. */
gsm_status gsm_render_host_async(gsm_renderer* r, const void* hostGaussians, const void* hostHarmonics,
                                 uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* camera,
                                 uint32_t width, uint32_t height, void* hostColor, void* hostDepth);
gsm_status gsm_render_host_wait(gsm_renderer* r);

/* lastGPUTime (GRP.swift:245; declared but never assigned by the reference, DFR.swift:43). Here: device
 * time of the last frame in ms when profiling is on, else a negative number. */
double gsm_last_gpu_time_ms(gsm_renderer* r);

/* Per-stage device timing (cudaEvent pairs around each stage). Costs a host sync per readout, so it is
 * off by default. names: project, depthSort, applyScan, expand, tileSort, ranges, blend, copy. */
#define GSM_NUM_STAGES 8
gsm_status gsm_set_profiling(gsm_renderer* r, int enabled);
gsm_status gsm_get_stage_times_ms(gsm_renderer* r, float* ms /* GSM_NUM_STAGES */);
const char* gsm_stage_name(int stage);

/* White-box reads mirroring the test-side debugRead* helpers (DFUT.swift:911-1252). Each copies
 * `count` elements starting at element `first` into host memory `dst` after synchronising `stream`. */
typedef enum {
    GSM_DBG_HEADER = 0,               /* debugReadHeader: GSMDepthFirstHeader x1 */
    GSM_DBG_ACTIVE_TILE_COUNT = 1,    /* debugReadActiveTileCount: u32 x1 */
    GSM_DBG_SORTED_TILE_IDS = 2,      /* debugReadSortedTileIds: u16 (tileIdPrecision 16) or u32 */
    GSM_DBG_TILE_BOUNDS = 3,          /* debugReadTileBounds / debugReadSingleBounds: int32 x4 per Gaussian */
    GSM_DBG_SORTED_PRIMITIVE_INDICES = 4, /* debugReadSortedPrimitiveIndices(+Range): int32 per visible */
    GSM_DBG_INSTANCE_OFFSETS = 5,     /* debugReadInstanceOffsets == debugReadOrderedTileCounts after the scan */
    GSM_DBG_N_TOUCHED_TILES = 6,      /* debugReadNTouchedTiles: u32 per Gaussian */
    GSM_DBG_INSTANCE_GAUSSIAN_INDICES = 7, /* debugReadInstanceGaussianIndices: int32 per instance, tile-sorted */
    GSM_DBG_DEPTH_KEYS = 8,           /* debugReadDepthKeys: u32 per visible, sorted */
    GSM_DBG_RENDER_DATA = 9,          /* debugReadRenderData: GSMGaussianRenderData (mono) / GSMStereoTiledRenderData */
    GSM_DBG_TILE_HEADERS = 10,        /* debugReadTileHeaders: GSMGaussianHeader per tile */
    GSM_DBG_ACTIVE_TILES = 11,        /* activeTiles list (atomic append order, nondeterministic as in the reference) */
    GSM_DBG_SCRATCH_DEPTH_KEYS = 12,  /* debugReadScratchDepthKeys: ping-pong buffer, content unspecified */
    GSM_DBG_SCRATCH_PRIMITIVE_INDICES = 13, /* debugReadScratchPrimitiveIndices: same */
    GSM_DBG_DEPTH_SORT_PLAN = 14,     /* no reference counterpart: u32 x4 {bucketCount, keyMin, fineShift, 0}: the bucket plan of
                                         the last frame's depth sort (csrc/bucketsort.cu; stale when the LSD passes ran) */
    GSM_DBG_COUNT_
} gsm_debug_buffer;

gsm_status gsm_debug_read(gsm_renderer* r, void* stream, int which, void* dst, size_t first, size_t count);

/* ---- GlobalRenderer (Sources/Renderer/GlobalRenderer/GlobalRenderer.swift:72-372, GlobalShaders.metal): the reference's second
 * renderer behind the same protocol, on the same handle (limits, precision, colour space as RendererConfig; 32 x 16-pixel tiles
 * of the LIMITS, GlobalRenderer.swift:26-49). gsm_render_global replaces GlobalRenderer.render (:206-247): project + cull,
 * visibility compaction, two-pass tile assignment, ONE sort of 32-bit [tile:16][half depth:16] keys, headers by binary search,
 * clear + render of 4 x 2 pixels per thread. Same buffers and errors as gsm_render; gaussianCount > maxGaussians encodes nothing
 * (validateLimits, :293-297). renderStereo has no entry: the reference's is a fatalError (:249-265).
 * gsm_global_debug_read mirrors debugReadTotalAssignments (:200-203) and the buffers the stages leave behind. */
typedef struct {
    uint32_t totalAssignments, paddedCount, overflow;   /* TileAssignmentHeader */
    uint32_t visibleCount, activeTileCount, totalRaw, _pad[2];
} gsm_global_header;
typedef enum {
    GSM_GDBG_HEADER = 0,           /* gsm_global_header x1 */
    GSM_GDBG_SORTED_KEYS = 1,      /* u32 per assignment, sorted */
    GSM_GDBG_SORTED_INDICES = 2,   /* int32 Gaussian index per assignment */
    GSM_GDBG_TILE_HEADERS = 3,     /* GSMGaussianHeader per tile of the limits */
    GSM_GDBG_BOUNDS = 4,           /* int32 x4 per Gaussian ((0,-1,0,-1) = culled) */
    GSM_GDBG_RENDER_DATA = 5,      /* GSMGaussianRenderData per Gaussian (culled entries unspecified) */
    GSM_GDBG_VISIBLE_INDICES = 6,  /* u32 per visible Gaussian, ascending */
    GSM_GDBG_ACTIVE_TILES = 7      /* tile ids, atomic append order */
} gsm_global_debug_buffer;
gsm_status gsm_render_global(gsm_renderer* r, void* stream, void* color, void* depth, const void* gaussians, const void* harmonics,
                             uint32_t gaussianCount, uint32_t shComponents, const gsm_camera* camera, uint32_t width, uint32_t height);
gsm_status gsm_global_debug_read(gsm_renderer* r, void* stream, int which, void* dst, size_t first, size_t count);
size_t gsm_debug_element_size(gsm_renderer* r, int which);

/* Device memory and streams for hosts without a CUDA runtime binding (the Swift facade): the MTLBuffer /
 * MTLCommandQueue replacements. */
gsm_status gsm_buffer_alloc(int device, size_t bytes, void** out);
gsm_status gsm_buffer_free(void* p);
gsm_status gsm_buffer_upload(void* dst, const void* src, size_t bytes, void* stream);
gsm_status gsm_buffer_download(void* dst, const void* src, size_t bytes, void* stream);
gsm_status gsm_stream_create(int device, void** out);
gsm_status gsm_stream_synchronize(void* stream);
gsm_status gsm_stream_destroy(void* stream);

/* Standalone stable radix sort of (key, payload) pairs on the device -- the entry the reference's sort
 * unit tests drive directly (DFUT.swift:120-468 call DepthRadixSortEncoder.encode). keyBits 16 or 32;
 * numPasses 8-bit digits from bit 0. Sorted data ends in keys/payload. */
gsm_status gsm_sort_pairs(gsm_renderer* r, void* stream, void* keys, void* payload, uint32_t count,
                          int keyBits, int numPasses);

/* Strip-sharded frame, library-collective form (kept as the baseline the peer-memory path above is measured against, and for
 * hosts without peer access): a single large frame split into
 * horizontal strips of whole tile rows. Step 1 (per rank): project+cull the gid range
 * [gidFirst, gidFirst+gidCount) and emit compacted 48-byte splat records (count in *hostCount after the
 * call synchronises). Step 2 (per rank, after an all-gather of the records in rank order): sort, expand,
 * tile-sort and blend only tile rows [tileRowFirst, tileRowFirst+tileRowCount) into the full-size targets. */
#define GSM_SPLAT_RECORD_BYTES 48
gsm_status gsm_strip_project(gsm_renderer* r, void* stream, const void* gaussians, const void* harmonics,
                             uint32_t gidFirst, uint32_t gidCount, uint32_t shComponents,
                             const gsm_camera* camera, uint32_t width, uint32_t height, void* recordsOut,
                             uint32_t* hostCount);
gsm_status gsm_strip_render(gsm_renderer* r, void* stream, void* color, void* depth, const void* records,
                            uint32_t recordCount, uint32_t width, uint32_t height, uint32_t tileRowFirst,
                            uint32_t tileRowCount);

/* The same sort without any allocation or host synchronisation (hosts that sort every frame): `scratch` is a device
 * buffer of gsm_sort_pairs_scratch_bytes(count, keyBits, numPasses) bytes, 256-byte aligned, that must stay untouched until
 * the stream has run the sort. */
size_t gsm_sort_pairs_scratch_bytes(uint32_t count, int keyBits, int numPasses);
gsm_status gsm_sort_pairs_with_scratch(gsm_renderer* r, void* stream, void* keys, void* payload, uint32_t count,
                                       int keyBits, int numPasses, void* scratch);

/* ---- gsm_group: one frame split over the GPUs of one node by horizontal strips of whole tile rows (SURVEY.md 8e, config C3;
 * no reference counterpart -- the reference is single-device). One process (or thread) per GPU, one renderer + one group
 * handle per rank. Every rank owns an EXCHANGE WINDOW in its HBM -- a mailbox, one receive region per source rank and an
 * optional frame image -- that every peer maps (cudaIpc between processes: gsm_group_export / gsm_group_connect carry the
 * 64-byte handles over whatever channel the host has; gsm_group_connect_local for ranks living in one process). A frame:
 *   gsm_group_project_route  projects + culls the rank's gid range and STORES each surviving 48-byte splat record straight
 *                            into the window of every rank whose strip it touches (NVLink peer stores issued by the routing
 *                            kernel itself, order-preserving per destination; counts and completion flags travel the same way);
 *   gsm_group_render_strip   waits (on the device) for all sources, ingests only the records of its own strip in global gid
 *                            order, then sorts, expands, tile-sorts and blends tile rows [rowStart[rank], rowStart[rank+1])
 *                            into `color` / `depth` -- full-size images, which may be another rank's window image
 *                            (gsm_group_image), so the frame is assembled by the blend's own stores;
 *   gsm_group_signal / gsm_group_wait  "my part of frame `frameId` is in rank `toRank`'s image" / the image's owner waits, on
 *                            the device, for the ranks in `fromMask`. frameId: any number that grows from frame to frame and
 *                            is the same on all ranks.
 * No call synchronises the host or returns a count to it; all ranks must issue the same sequence of frames (collective
 * semantics). Per-strip lists are bit-identical to the single-GPU frame's (records arrive in global gid order, so the stable
 * depth sort breaks ties the same way); the 4*maxGaussians instance cap applies per strip. */
typedef struct gsm_group gsm_group;
#define GSM_GROUP_HANDLE_BYTES 64
#define GSM_GROUP_MAX_RANKS 8
/* maxRecordsPerSource: the largest gid shard any rank will project (e.g. ceil(N / world)); imageColorBytes / imageDepthBytes:
 * size of the window's frame image (0 = none). The renderer's maxGaussians must cover the WHOLE scene (global gids). */
gsm_status gsm_group_create(gsm_renderer* r, uint32_t rank, uint32_t world, uint32_t maxRecordsPerSource,
                            size_t imageColorBytes, size_t imageDepthBytes, gsm_group** out);
void gsm_group_destroy(gsm_group* g);
gsm_status gsm_group_export(gsm_group* g, void* handleOut /* GSM_GROUP_HANDLE_BYTES */);
gsm_status gsm_group_connect(gsm_group* g, const void* handles /* world * GSM_GROUP_HANDLE_BYTES, rank order */);
gsm_status gsm_group_connect_local(gsm_group* g, gsm_group* const* peers /* world handles of this process, rank order */);
gsm_status gsm_group_image(gsm_group* g, uint32_t ofRank, void** color, void** depth);
/* stripRowStart: world + 1 tile-row boundaries, stripRowStart[0] == 0, stripRowStart[world] == tilesY, non-decreasing. */
gsm_status gsm_group_project_route(gsm_group* g, void* stream, const void* gaussiansShard, const void* harmonicsShard,
                                   uint32_t gidFirst, uint32_t gidCount, uint32_t shComponents, const gsm_camera* camera,
                                   uint32_t width, uint32_t height, const uint32_t* stripRowStart);
gsm_status gsm_group_render_strip(gsm_group* g, void* stream, void* color, void* depth, uint32_t width, uint32_t height,
                                  const uint32_t* stripRowStart);
gsm_status gsm_group_signal(gsm_group* g, void* stream, uint32_t toRank, uint32_t frameId);
gsm_status gsm_group_wait(gsm_group* g, void* stream, uint32_t fromMask, uint32_t frameId);
/* Diagnostic (synchronises `stream`): how many records each source rank routed to this rank in the last frame. */
gsm_status gsm_group_record_counts(gsm_group* g, void* stream, uint32_t* countsOut /* world */);
/* gsm_group_project_route followed by gsm_group_render_strip (ranks in separate processes / on separate streams). */
gsm_status gsm_render_strips(gsm_group* g, void* stream, void* color, void* depth, const void* gaussiansShard,
                             const void* harmonicsShard, uint32_t gidFirst, uint32_t gidCount, uint32_t shComponents,
                             const gsm_camera* camera, uint32_t width, uint32_t height, const uint32_t* stripRowStart);

/* Math probes: run the device restatement of the canonical transcendental definitions on arrays, so the
 * tests can compare them bit-for-bit with the CPU oracle. op: 0 sin, 1 cos, 2 log, 3 atan2(a,b),
 * 4 powr(a, 2.4), 5 half exp (a,out are u16), 6 float->half (out u16). Host pointers. */
gsm_status gsm_probe_math(int device, int op, const void* a, const void* b, void* out, uint32_t n);
/* Which evaluation of exp(-0.5h * p) (DepthFirstShaders.metal:1775-1778, GlobalShaders.metal:1117-1124) the blend kernels of
 * `device` (-1: current) run: 2 = MUFU.EX2 on a tuned argument, 1 = the canonical polynomial, 0 = no renderer created on the
 * device yet. Both give identical bits; gsm_renderer_create proves it on the device over all 65 536 half inputs and falls back to
 * the polynomial if any input differs (environment GSM_BLEND_EXP=poly forces the polynomial). */
int gsm_blend_exp_mode(int device);

const char* gsm_status_string(gsm_status s);
const char* gsm_last_error_string(void);
int gsm_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GSM_H */
