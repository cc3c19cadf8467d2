/*
 * gsm_types.h -- Linux restatement of the hot-path structs of the reference's RendererTypes C module
 * (Sources/RendererTypes/include/BridgingTypes.h), usable from C, C++, CUDA and as a SwiftPM C target.
 *
 * BridgingTypes.h includes <simd/simd.h> (Apple-only, :5,:13); the 20-line shim below supplies the three
 * simd types it uses with the same size and alignment, so the struct layouts (and the byte offsets the
 * reference documents at :57,:66,:75,:163-206,:256-276) are unchanged. Hardware*-renderer structs are not
 * on the DepthFirst path and are omitted.
 */
#ifndef GSM_TYPES_H
#define GSM_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
#define GSM_STATIC_ASSERT(c, m) static_assert(c, m)
#define GSM_ALIGN(n) alignas(n)
#else
#define GSM_STATIC_ASSERT(c, m) _Static_assert(c, m)
#define GSM_ALIGN(n) _Alignas(n)
#endif

/* ---- simd shim (column-major float4x4; float3 padded to 16 bytes like simd_float3) ---- */
typedef struct { GSM_ALIGN(16) float x; float y, z, w; } gsm_simd_float4;
typedef struct { GSM_ALIGN(16) float x; float y, z, _w; } gsm_simd_float3;
typedef struct { gsm_simd_float4 columns[4]; } gsm_simd_float4x4;

typedef uint16_t GSM_HALF; /* IEEE binary16 bits (MSL half / Swift Float16) */

/* BridgingTypes.h:22-39 */
typedef struct {
    gsm_simd_float4x4 viewMatrix;
    gsm_simd_float4x4 projectionMatrix;
    gsm_simd_float3 cameraCenter;
    float pixelFactor;
    float focalX;
    float focalY;
    float width;
    float height;
    float nearPlane;
    float farPlane;
    uint32_t shComponents;
    uint32_t gaussianCount;
    float inputIsSRGB;
    float _pad1, _pad2, _pad3;
} GSMCameraUniforms;

/* BridgingTypes.h:41-50 */
typedef struct {
    uint32_t width, height, tileWidth, tileHeight, tilesX, tilesY, activeTileCount, gaussianCount;
} GSMRenderParams;

/* BridgingTypes.h:52-55 */
typedef struct { uint32_t offset, count; } GSMGaussianHeader;

/* BridgingTypes.h:58-64, 48 bytes */
typedef struct {
    float px, py, pz;
    float opacity;
    float sx, sy, sz;
    float _pad0;
    gsm_simd_float4 rotation; /* (x, y, z, w) with w the real part */
} GSMPackedWorldGaussian;

/* BridgingTypes.h:67-73, 32 bytes */
typedef struct {
    float px, py, pz;
    GSM_HALF opacity;
    GSM_HALF sx, sy, sz;
    GSM_HALF rx, ry, rz, rw;
    GSM_HALF _pad0, _pad1;
} GSMPackedWorldGaussianHalf;

/* BridgingTypes.h:76-84, 16 bytes */
typedef struct {
    GSM_HALF meanX, meanY;
    uint16_t theta; /* angle in [0, pi) packed to 0..65535 */
    GSM_HALF sigma1, sigma2;
    GSM_HALF depth;
    uint8_t colorR, colorG, colorB;
    uint8_t opacity;
} GSMGaussianRenderData;

/* BridgingTypes.h:86-97 */
typedef struct {
    uint32_t gaussianCount, tilesX, tilesY, tileWidth, tileHeight, surfaceWidth, surfaceHeight, maxCapacity;
    float alphaThreshold;
    float totalInkThreshold;
} GSMTileBinningParams;

/* BridgingTypes.h:116-120 */
typedef struct { uint32_t threadgroupsPerGridX, threadgroupsPerGridY, threadgroupsPerGridZ; } GSMDispatchIndirectArgs;

/* BridgingTypes.h:163-206, 416 bytes */
typedef struct {
    gsm_simd_float4x4 leftViewMatrix;
    gsm_simd_float4x4 leftProjectionMatrix;
    float leftCameraCenterX, leftCameraCenterY, leftCameraCenterZ;
    float leftFocalX, leftFocalY;
    float _padLeft0, _padLeft1, _padLeft2;
    gsm_simd_float4x4 rightViewMatrix;
    gsm_simd_float4x4 rightProjectionMatrix;
    float rightCameraCenterX, rightCameraCenterY, rightCameraCenterZ;
    float rightFocalX, rightFocalY;
    float _padRight0, _padRight1, _padRight2;
    float width, height, nearPlane, farPlane;
    uint32_t shComponents, gaussianCount;
    float inputIsSRGB;
    float _padShared1;
    gsm_simd_float4x4 sceneTransform;
} GSMStereoCameraUniforms;

/* BridgingTypes.h:210-219, 32 bytes */
typedef struct {
    uint32_t visibleCount, totalInstances, paddedVisibleCount, paddedInstanceCount, overflow;
    uint32_t padding0, padding1, padding2;
} GSMDepthFirstHeader;

/* BridgingTypes.h:256-276, 32 bytes */
typedef struct {
    GSM_HALF leftMeanX, leftMeanY, leftCxx, leftCyy, leftCxy2, leftDepth;
    GSM_HALF rightMeanX, rightMeanY, rightCxx, rightCyy, rightCxy2, rightDepth;
    uint8_t colorR, colorG, colorB;
    uint8_t opacity;
    GSM_HALF centerDepth;
    uint16_t _pad0;
} GSMStereoTiledRenderData;

GSM_STATIC_ASSERT(sizeof(GSMPackedWorldGaussian) == 48, "PackedWorldGaussian is 48 bytes");
GSM_STATIC_ASSERT(sizeof(GSMPackedWorldGaussianHalf) == 32, "PackedWorldGaussianHalf is 32 bytes");
GSM_STATIC_ASSERT(sizeof(GSMGaussianRenderData) == 16, "GaussianRenderData is 16 bytes");
GSM_STATIC_ASSERT(sizeof(GSMStereoTiledRenderData) == 32, "StereoTiledRenderData is 32 bytes");
GSM_STATIC_ASSERT(sizeof(GSMDepthFirstHeader) == 32, "DepthFirstHeader is 32 bytes");
GSM_STATIC_ASSERT(sizeof(GSMStereoCameraUniforms) == 416, "StereoCameraUniforms is 416 bytes");
GSM_STATIC_ASSERT(sizeof(GSMCameraUniforms) == 208, "CameraUniforms is 208 bytes");
GSM_STATIC_ASSERT(sizeof(GSMGaussianHeader) == 8, "GaussianHeader is 8 bytes");

#endif /* GSM_TYPES_H */
