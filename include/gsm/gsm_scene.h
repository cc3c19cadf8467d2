/*
 * gsm_scene.h -- scene ingest for the DepthFirst path (SURVEY.md 8(f) rank 1 + the Morton pre-sort of rank 2): what the
 * reference's PLYLoader / GaussianSceneBuilder (Sources/Renderer/Utils/PLYLoader.swift, Scene.swift) hand to the renderer.
 * Same C ABI rules as gsm.h: plain pointers and sizes. The file is parsed on the host (header) and DECODED ON THE DEVICE:
 * the caller maps or reads the .ply into host memory and gets PackedWorldGaussian(+Half) and the planar SH buffer in
 * device memory, ready for gsm_render.
 */
#ifndef GSM_SCENE_H
#define GSM_SCENE_H

#include "gsm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Additional gsm_status values returned by the loader (PLYLoaderError, PLYLoader.swift:209-242) */
enum {
    GSM_ERR_PLY_INVALID_HEADER = 20,        /* .invalidHeader / PLYHeader.DecodeError */
    GSM_ERR_PLY_UNSUPPORTED_FORMAT = 21,    /* .unsupportedFormat: only binary_little_endian */
    GSM_ERR_PLY_MISSING_VERTEX = 22,        /* .missingVertexElement */
    GSM_ERR_PLY_MISSING_PROPERTIES = 23,    /* .missingRequiredProperties (x, y, z) */
    GSM_ERR_PLY_LIST_PROPERTY = 24,         /* .listPropertiesNotSupported */
    GSM_ERR_PLY_INSUFFICIENT_DATA = 25,     /* .insufficientData (also: caller's buffers too small) */
    GSM_ERR_PLY_MISSING_CHUNK = 26          /* .missingChunkElement */
};

/* PLYHeader (PLYLoader.swift:6-86), reduced to what sizes the caller's buffers */
typedef struct {
    uint32_t vertexCount;   /* element vertex <count> */
    uint32_t format;        /* 0 ascii, 1 binary_little_endian, 2 binary_big_endian */
    uint32_t compressed;    /* PlayCanvas / splat-transform layout (chunk element + packed_* properties) */
    uint32_t shProperties;  /* SH-like vertex properties (f_dc_*, f_rest_*, sh_*, spherical_harmonics_*); 3 when compressed */
    uint64_t bodyOffset;    /* first byte after end_header */
} gsm_ply_info;

/* GaussianDataset minus the records themselves (Scene.swift:141-157) */
typedef struct {
    uint32_t count;            /* Gaussians kept (placeholder vertices are skipped, PLYLoader.swift:658-660) */
    uint32_t shComponents;     /* GaussianDataset.shComponents */
    uint32_t harmonicsStride;  /* elements per Gaussian in the harmonics buffer */
    uint32_t compressed, scaleIsLogSpace, opacityIsLogit; /* format detection, PLYLoader.swift:620-650 */
    float center[3];           /* what the recentering subtracted (zero if |center| <= 1e-6) */
    float boundsCenter[3];     /* GaussianSceneBuilder.bounds(of:) of the final records (Scene.swift:159-190) */
    float boundsRadius;
} gsm_scene_info;

/* PLYHeader.decodeASCII + the element checks of PLYLoader.load; host only. */
gsm_status gsm_ply_probe(const void* fileBytes, size_t fileSize, gsm_ply_info* info);

/* PLYLoader.load (PLYLoader.swift:246-281) + the packing of PLYBenchmarkTests.swift:139-149. fileBytes is HOST memory
 * (the whole file); gaussiansOut / harmonicsOut are DEVICE buffers of gaussianCapacity PackedWorldGaussian (precision
 * FLOAT32, 48 B) or PackedWorldGaussianHalf (FLOAT16, 32 B) and harmonicsCapacity float / half elements. Synchronises
 * the stream before returning (info is host memory). */
gsm_status gsm_ply_load(int device, void* stream, const void* fileBytes, size_t fileSize, int precision, void* gaussiansOut,
                        void* harmonicsOut, uint32_t gaussianCapacity, size_t harmonicsCapacity, gsm_scene_info* info);

/* GaussianSceneBuilder.sortByMortonCode (Scene.swift:73-138) on the packed device buffers, in place: 21-bit-per-axis Morton
 * codes of the positions normalised to their bounding box, stable ascending sort (ties keep their order), records and
 * harmonics permuted together. */
gsm_status gsm_scene_morton_sort(int device, void* stream, void* gaussians, void* harmonics, uint32_t count,
                                 uint32_t harmonicsStride, int precision);

#ifdef __cplusplus
}
#endif
#endif /* GSM_SCENE_H */
