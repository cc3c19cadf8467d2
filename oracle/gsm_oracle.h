/*
 * gsm_oracle.h -- CPU oracle for the DepthFirstRenderer hot path of LuckyIYI/gsm-renderer.
 *
 * TEST INFRASTRUCTURE ONLY (checker / CPU baseline). The product path never links this.
 *
 * A stage-by-stage restatement in plain C of the reference's Metal kernels
 * (Sources/Renderer/DepthFirstRenderer/DepthFirstShaders.metal = "DFS.metal",
 *  Sources/Renderer/Shared/GaussianShared.h = "GShared.h") and of the host-side stage order
 * (Sources/Renderer/DepthFirstRenderer/DepthFirstRenderer.swift = "DFR.swift").
 *
 * Parity pinning status:
 *   - sort stages: PINNED by the reference's own known-answer tests
 *     (DepthFirstUnitTests.swift:120-305 and :308-468, GlobalUnitTests.swift:23-105);
 *   - frame counters: PINNED (weakly) by testDepthFirstPipelineStages
 *     (DepthFirstUnitTests.swift:21-117: overflow==0, 0<V<=1000, I>0);
 *   - tile counts, instance order, tile ranges, pixels: PARITY UNPINNED by the reference --
 *     its tests hold no values for them and the Metal path cannot run here, so this oracle
 *     is the sole definition, as BASELINE.json's north_star prescribes;
 *   - foveated stereo copy (gsm_oracle_copy.c): PARITY UNPINNED -- no reference test, and the rate map,
 *     filter weights and attachment conversion are Metal implementation behaviour; pinned only at 1:1
 *     against the literal copy (tests/test_foveated_oracle.py).
 */
#ifndef GSM_ORACLE_H
#define GSM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t gsmo_half;

/* BridgingTypes.h:58-64 (48 B) */
typedef struct {
    float px, py, pz;
    float opacity;
    float sx, sy, sz;
    float _pad0;
    float rot[4]; /* x, y, z, w(real) */
} gsmo_packed_f32;

/* BridgingTypes.h:67-73 (32 B) */
typedef struct {
    float px, py, pz;
    gsmo_half opacity;
    gsmo_half sx, sy, sz;
    gsmo_half rx, ry, rz, rw;
    gsmo_half _pad0, _pad1;
} gsmo_packed_f16;

/* BridgingTypes.h:76-84 (16 B) */
typedef struct {
    gsmo_half meanX, meanY;
    uint16_t theta;
    gsmo_half sigma1, sigma2;
    gsmo_half depth;
    uint8_t colorR, colorG, colorB, opacity;
} gsmo_render_data;

/* BridgingTypes.h:256-276 (32 B) */
typedef struct {
    gsmo_half leftMeanX, leftMeanY, leftCxx, leftCyy, leftCxy2, leftDepth;
    gsmo_half rightMeanX, rightMeanY, rightCxx, rightCyy, rightCxy2, rightDepth;
    uint8_t colorR, colorG, colorB, opacity;
    gsmo_half centerDepth;
    uint16_t _pad0;
} gsmo_stereo_render_data;

/* BridgingTypes.h:52-55 */
typedef struct { uint32_t offset, count; } gsmo_tile_header;

/* BridgingTypes.h:210-219 */
typedef struct {
    uint32_t visibleCount, totalInstances, paddedVisibleCount, paddedInstanceCount, overflow;
    uint32_t padding0, padding1, padding2;
} gsmo_df_header;

/* The fields of CameraUniforms (BridgingTypes.h:22-39) that the path reads. Matrices are
 * column-major: m[4*c + r]. */
typedef struct {
    float view[16];
    float proj[16];
    float center[3];
    float width, height;
    float nearPlane, farPlane;
    uint32_t shComponents;
    uint32_t gaussianCount;
    float inputIsSRGB;
} gsmo_camera;

/* StereoCameraUniforms (BridgingTypes.h:163-206), the fields the path reads. */
typedef struct {
    float leftView[16], leftProj[16], leftCenter[3];
    float rightView[16], rightProj[16], rightCenter[3];
    float width, height, nearPlane, farPlane;
    uint32_t shComponents, gaussianCount;
    float inputIsSRGB;
    float sceneTransform[16];
} gsmo_stereo_camera;

/* TileBinningParams (BridgingTypes.h:86-97) as built by GlobalRenderer.swift:54-69. */
typedef struct {
    uint32_t tilesX, tilesY, tileWidth, tileHeight;
    float alphaThreshold, totalInkThreshold;
} gsmo_binning;

enum { GSMO_F32 = 0, GSMO_F16 = 1 };

int gsmo_num_threads(void);
void gsmo_set_num_threads(int n);

/* --- math probes (tests compare the CUDA restatement against these) --- */
void gsmo_probe_sincos(const float* x, float* s, float* c, int n);
void gsmo_probe_log(const float* x, float* y, int n);
void gsmo_probe_atan2(const float* y, const float* x, float* r, int n);
void gsmo_probe_powr(const float* x, float yexp, float* r, int n);
void gsmo_probe_hexp(const gsmo_half* x, gsmo_half* y, int n);
void gsmo_probe_f2h(const float* x, gsmo_half* y, int n);
void gsmo_probe_h2f(const gsmo_half* x, float* y, int n);
void gsmo_probe_hfma(const gsmo_half* a, const gsmo_half* b, const gsmo_half* c, gsmo_half* r, int n);
void gsmo_probe_minmax(const float* a, const float* b, float* mn, float* mx, int n);

/* --- stages --- */

/* DFS.metal:46-219. precision selects gsmo_packed_f32+float SH or gsmo_packed_f16+half SH.
 * preDepthKeys entries of the three "stale key" exits (DFS.metal:111-122,:133-137) are left
 * untouched. Returns sum of nTouched in *totalInstances (the atomic of DFS.metal:218). */
void gsmo_project_cull(const void* gaussians, const void* harmonics, int precision,
                       const gsmo_camera* cam, const gsmo_binning* bin,
                       gsmo_render_data* renderData, int32_t* bounds /*4 per gaussian*/,
                       uint32_t* preDepthKeys, uint32_t* nTouched, uint32_t* totalInstances);

/* DFS.metal:341-499 (+ projectToEye :249-339). */
void gsmo_project_cull_stereo(const void* gaussians, const void* harmonics, int precision,
                              const gsmo_stereo_camera* cam, const gsmo_binning* bin,
                              gsmo_stereo_render_data* renderData, int32_t* bounds,
                              uint32_t* preDepthKeys, uint32_t* nTouched, uint32_t* totalInstances);

/* DFS.metal:518-621 + VisibilityCompactionEncoder.swift:44-184: dense (key, gid) in ascending
 * gid order for nTouched>0; *visibleCount is the unclamped count; writes are bounded by maxOut.
 * depthKey16 != 0 applies the function-constant-2 path (DFS.metal:607-612). */
void gsmo_compact_visible(const uint32_t* nTouched, const uint32_t* preDepthKeys, uint32_t count,
                          uint32_t maxOut, int depthKey16, uint32_t* depthKeys,
                          int32_t* primitiveIndices, uint32_t* visibleCount);

/* DFS.metal:2174-2204: clamp, overflow flag, padded counts (multiples of 1024). */
void gsmo_prepare_header(uint32_t visibleCount, uint32_t totalInstances, uint32_t maxGaussians,
                         uint32_t maxInstances, gsmo_df_header* header);

/* Stable ascending LSD radix sort, 8-bit digits, numPasses passes from bit 0
 * (DFS.metal:1387-1696 + DepthRadixSortEncoder.swift:139-217; DFS.metal:866-1256 +
 * TileSortEncoder.swift:51-178). In place on the first count elements. */
void gsmo_sort_pairs_u32(uint32_t* keys, int32_t* payload, uint32_t count, int numPasses);
void gsmo_sort_pairs_u16(uint16_t* keys, int32_t* payload, uint32_t count, int numPasses);

/* DFS.metal:623-640 then DFS.metal:2036-2139 (exclusive scan, in place in the reference). */
void gsmo_apply_depth_order(const int32_t* sortedIdx, const uint32_t* nTouched, uint32_t visibleCount,
                            uint32_t* orderedTileCounts);
void gsmo_exclusive_scan(const uint32_t* in, uint32_t count, uint32_t* out);

/* DFS.metal:642-716 (tileId16 != 0) / :718-788. tileIds is uint16_t* or uint32_t*. */
void gsmo_create_instances(const int32_t* sortedIdx, const uint32_t* instanceOffsets,
                           const int32_t* bounds, const gsmo_render_data* renderData,
                           uint32_t visibleCount, uint32_t tilesX, float alphaThreshold,
                           uint32_t maxAssignments, int tileId16, void* tileIds, int32_t* instanceIdx);
/* DFS.metal:790-864 */
void gsmo_create_instances_stereo(const int32_t* sortedIdx, const uint32_t* instanceOffsets,
                                  const int32_t* bounds, uint32_t visibleCount, uint32_t tilesX,
                                  uint32_t maxAssignments, int tileId16, void* tileIds,
                                  int32_t* instanceIdx);

/* TileSortEncoder.swift:61-62 */
int gsmo_tile_sort_passes(uint32_t tileCount);

/* DFS.metal:1258-1370. activeTiles are emitted in ascending tile order (the reference's
 * atomic append order is nondeterministic; compare as a set). */
void gsmo_extract_ranges(const void* sortedTileIds, int tileId16, uint32_t totalInstances,
                         uint32_t tileCount, gsmo_tile_header* headers, uint32_t* activeTiles,
                         uint32_t* activeTileCount);

/* DFS.metal:2020-2034 then :1703-1811. color = W*H*4 halfs (rgba16f), depth = W*H halfs or NULL. */
void gsmo_clear(gsmo_half* color, gsmo_half* depth, uint32_t width, uint32_t height);
void gsmo_blend(const gsmo_tile_header* headers, const gsmo_render_data* renderData,
                const int32_t* sortedInstanceIdx, const uint32_t* activeTiles,
                uint32_t activeTileCount, uint32_t width, uint32_t height, uint32_t tilesX,
                gsmo_half* color, gsmo_half* depth);

/* DFS.metal:1813-1823 + :1825-1982. color2 = 2 slices of W*H*4 halfs (left then right). */
void gsmo_blend_stereo(const gsmo_tile_header* headers, const gsmo_stereo_render_data* renderData,
                       const int32_t* sortedInstanceIdx, const uint32_t* activeTiles,
                       uint32_t activeTileCount, uint32_t width, uint32_t height, uint32_t tilesX,
                       gsmo_half* color2);
/* DFS.metal:1984-2018 + DepthFirstStereoCopyEncoder.swift:28-100 at 1:1 with viewports
 * (0,0,W,H) and (W,0,W,H): dst is 2W x H rgba16f. flipY != 0 is the literal behaviour
 * (NDC(-1,-1) -> uv(0,0): dst row y = src row H-1-y), flipY == 0 copies rows straight. */
void gsmo_stereo_copy(const gsmo_half* color2, uint32_t width, uint32_t height, int flipY,
                      gsmo_half* dstSideBySide);

/* The same copy through a drawable (FoveatedStereoDrawable, GRP.swift:168-193): any viewports, an optional tabulated
 * rasterization-rate map, five attachment formats (0 rgba16f, 1 bgra8, 2 bgra8_srgb, 3 rgba8, 4 rgba8_srgb). color2 as
 * above (two slices). layerCount 0 = no rate map. viewports = {ox, oy, w, h} left then right. gsm_oracle_copy.c. */
void gsmo_stereo_copy_foveated(const gsmo_half* color2, uint32_t width, uint32_t height, int flipY, uint8_t* dst,
                               uint32_t textureWidth, uint32_t textureHeight, uint32_t arrayLength, size_t rowBytes,
                               size_t sliceBytes, int format, uint32_t layerCount, const uint32_t physicalWidth[2],
                               const uint32_t physicalHeight[2], const float* const screenX[2],
                               const float* const screenY[2], const double viewports[8]);

/* --- whole frames (DFR.swift:237-465 and :595-831); every intermediate is returned --- */
typedef struct {
    /* capacities (caller-set) */
    uint32_t maxGaussians, maxInstances; /* maxInstances = 4*maxGaussians, DepthFirstResources.swift:80 */
    int depthKey16, tileId16;
    /* per-gaussian (caller-allocated, maxGaussians entries) */
    void* renderData;     /* gsmo_render_data or gsmo_stereo_render_data */
    int32_t* bounds;      /* 4 per gaussian */
    uint32_t* nTouched;
    uint32_t* preDepthKeys;
    /* per-visible (maxGaussians entries) */
    uint32_t* depthKeys;
    int32_t* primitiveIndices;
    uint32_t* orderedTileCounts; /* after the scan: instance offsets */
    /* per-instance (maxInstances entries) */
    void* instanceTileIds; /* u16 or u32 */
    int32_t* instanceGaussianIndices;
    /* per-tile */
    gsmo_tile_header* tileHeaders;
    uint32_t* activeTiles;
    /* results */
    gsmo_df_header header;
    uint32_t activeTileCount;
    uint32_t rawVisibleCount, rawTotalInstances;
    double stageSeconds[10]; /* project, compact, depthSort, applyScan, expand, tileSort, ranges, clear+blend, copy, total */
} gsmo_frame;

void gsmo_render_mono(gsmo_frame* f, const void* gaussians, const void* harmonics, int precision,
                      const gsmo_camera* cam, uint32_t width, uint32_t height,
                      gsmo_half* color, gsmo_half* depth);
void gsmo_render_stereo(gsmo_frame* f, const void* gaussians, const void* harmonics, int precision,
                        const gsmo_stereo_camera* cam, uint32_t width, uint32_t height, int flipY,
                        gsmo_half* scratchColor2, gsmo_half* dstSideBySide);

/* test-only: blend contraction convention (1 = canonical fused half FMA, 0 = separately rounded mul/add) */
void gsmo_set_blend_contraction(int on);
int gsmo_get_blend_contraction(void);

/* --- GlobalRenderer (GlobalShaders.metal, GlobalRenderer.swift): one whole frame with its white-box buffers.
 * renderData / bounds / mask: per Gaussian (entries of culled Gaussians: mask 0, bounds (0,-1,0,-1), renderData untouched);
 * visibleIndices: N; sortedKeys / sortedIndices: 4 * maxGaussians; headers: tiles of the LIMITS (32 x 16 px). */
typedef struct { uint32_t totalAssignments, paddedCount, overflow, visibleCount, activeTileCount; } gsmo_global_info;
void gsmo_render_global(const void* gaussians, const void* harmonics, int precision, const gsmo_camera* cam,
                        uint32_t maxWidth, uint32_t maxHeight, uint32_t maxGaussians, uint32_t width, uint32_t height,
                        gsmo_half* color, gsmo_half* depthOut, gsmo_render_data* renderData, int32_t* bounds, uint8_t* mask,
                        uint32_t* visibleIndices, uint32_t* sortedKeys, int32_t* sortedIndices, gsmo_tile_header* headers,
                        gsmo_global_info* info);

#ifdef __cplusplus
}
#endif
#endif
