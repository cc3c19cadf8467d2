// DepthFirstRenderer.hpp -- header-only C++ facade over the C ABI (gsm.h), mirroring the reference's Swift
// surface name for name so host code reads like the reference's:
//   RendererConfig / GaussianInput / CameraParams / StereoCameraParams / StereoRenderTarget / RendererError
//     <- Sources/Renderer/Shared/GaussianRendererProtocol.swift:9-67,195-239,274-324
//   DepthFirstRenderer(init, render, renderStereo, lastGPUTime)
//     <- Sources/Renderer/DepthFirstRenderer/DepthFirstRenderer.swift:45-101,166-235
//   debugRead*  <- Tests/RendererTests/DepthFirstUnitTests.swift:911-1252
// The Swift toolchain is absent in this image, so this facade (compiled, reference is compiled code) and the
// Python mirror (gsm_renderer_b200/renderer.py) are the executable hosts; swift/ holds the same facade as
// Swift source.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "gsm.h"
#include "gsm_scene.h"
#include "gsm_types.h"

namespace gsm {

enum class RenderPrecision : uint32_t { float32 = GSM_PRECISION_FLOAT32, float16 = GSM_PRECISION_FLOAT16 };
enum class GaussianColorSpace : uint32_t { linear = GSM_COLORSPACE_LINEAR, srgb = GSM_COLORSPACE_SRGB };
enum class RadixSortKeyPrecision : uint32_t { bits16 = GSM_KEY_BITS16, bits32 = GSM_KEY_BITS32 };

// RendererError: one exception type, `status` is the case.
class RendererError : public std::runtime_error {
public:
    RendererError(gsm_status s, const std::string& detail)
        : std::runtime_error(std::string(gsm_status_string(s)) + (detail.empty() ? "" : ": " + detail)), status(s) {}
    gsm_status status;
};

inline void check(gsm_status s) {
    if (s != GSM_OK) throw RendererError(s, gsm_last_error_string());
}

struct RendererConfig {
    int maxGaussians = 6000000;
    int maxWidth = 1920;
    int maxHeight = 1080;
    RenderPrecision precision = RenderPrecision::float16;
    GaussianColorSpace gaussianColorSpace = GaussianColorSpace::srgb;
    bool backToFront = false;  // ignored by DepthFirst (as is colorFormat)
};

// MTLBuffer -> device pointer
struct GaussianInput {
    const void* gaussians;   // PackedWorldGaussian (48 B) or PackedWorldGaussianHalf (32 B)
    const void* harmonics;
    int gaussianCount;
    int shComponents;
};

struct CameraParams {
    std::array<float, 16> viewMatrix;        // column-major (simd_float4x4)
    std::array<float, 16> projectionMatrix;
    std::array<float, 3> position;
    float focalX = 0, focalY = 0;
    float near = 0.1f, far = 10.0f;
    gsm_camera native() const {
        gsm_camera c{};
        for (int i = 0; i < 16; ++i) { c.viewMatrix[i] = viewMatrix[i]; c.projectionMatrix[i] = projectionMatrix[i]; }
        for (int i = 0; i < 3; ++i) c.position[i] = position[i];
        c.focalX = focalX; c.focalY = focalY; c.nearPlane = near; c.farPlane = far;
        return c;
    }
};

struct StereoCameraParams { CameraParams leftEye, rightEye; };

// MTLViewport / EyeView / StereoConfiguration (GaussianRendererProtocol.swift:68-117)
struct Viewport { double originX = 0, originY = 0, width = 0, height = 0; };

struct EyeView {
    Viewport viewport;
    std::array<float, 16> viewMatrix;
    std::array<float, 16> projectionMatrix;
    std::array<float, 3> cameraPosition;
    float focalX = 0, focalY = 0;
    float near = 0.1f, far = 10.0f;
    gsm_eye_view native() const {
        gsm_eye_view e{};
        e.viewport = gsm_viewport{viewport.originX, viewport.originY, viewport.width, viewport.height};
        e.camera = CameraParams{viewMatrix, projectionMatrix, cameraPosition, focalX, focalY, near, far}.native();
        return e;
    }
};

struct StereoConfiguration {
    EyeView leftEye, rightEye;
    std::array<float, 16> sceneTransform{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    gsm_stereo_configuration native() const {
        gsm_stereo_configuration c{};
        c.leftEye = leftEye.native(); c.rightEye = rightEye.native();
        for (int i = 0; i < 16; ++i) c.sceneTransform[i] = sceneTransform[i];
        return c;
    }
};

// FoveatedStereoDrawable (GaussianRendererProtocol.swift:168-193); the rate map is the tabulated form of gsm.h
struct FoveatedStereoDrawable {
    void* colorTexture = nullptr;
    uint32_t textureWidth = 0, textureHeight = 0, arrayLength = 2;
    size_t rowBytes = 0, sliceBytes = 0;               // 0 = tightly packed
    const gsm_rate_map* rasterizationRateMap = nullptr;
    gsm_pixel_format colorPixelFormat = GSM_PIXEL_BGRA8_SRGB;
    gsm_foveated_drawable native() const {
        gsm_foveated_drawable d{};
        const size_t px = colorPixelFormat == GSM_PIXEL_RGBA16F ? 8 : 4;
        d.colorTexture = colorTexture; d.textureWidth = textureWidth; d.textureHeight = textureHeight; d.arrayLength = arrayLength;
        d.rowBytes = rowBytes ? rowBytes : (size_t)textureWidth * px;
        d.sliceBytes = sliceBytes ? sliceBytes : d.rowBytes * textureHeight;
        d.colorPixelFormat = (uint32_t)colorPixelFormat; d.rasterizationRateMap = rasterizationRateMap;
        return d;
    }
};

// StereoRenderTarget (GaussianRendererProtocol.swift:233-239): .sideBySide -- rgba16f (2*width) x height -- or .foveated
struct StereoRenderTarget {
    void* colorTexture = nullptr;
    void* depthTexture = nullptr;  // ignored by the reference too (DepthFirstRenderer.swift:472)
    bool isFoveated = false;
    FoveatedStereoDrawable drawable;
    StereoConfiguration configuration;
    static StereoRenderTarget sideBySide(void* color, void* depth = nullptr) {
        StereoRenderTarget t; t.colorTexture = color; t.depthTexture = depth; return t;
    }
    static StereoRenderTarget foveated(const FoveatedStereoDrawable& drawable, const StereoConfiguration& configuration) {
        StereoRenderTarget t; t.isFoveated = true; t.drawable = drawable; t.configuration = configuration; return t;
    }
};

class DepthFirstRenderer {
public:
    explicit DepthFirstRenderer(int device = -1, RendererConfig config = RendererConfig(),
                                RadixSortKeyPrecision depthSortKeyPrecision = RadixSortKeyPrecision::bits32,
                                RadixSortKeyPrecision tileIdPrecision = RadixSortKeyPrecision::bits16) {
        gsm_config c;
        gsm_config_default(&c);
        c.maxGaussians = (uint32_t)config.maxGaussians;
        c.maxWidth = (uint32_t)config.maxWidth;
        c.maxHeight = (uint32_t)config.maxHeight;
        c.precision = (uint32_t)config.precision;
        c.gaussianColorSpace = (uint32_t)config.gaussianColorSpace;
        c.depthSortKeyPrecision = (uint32_t)depthSortKeyPrecision;
        c.tileIdPrecision = (uint32_t)tileIdPrecision;
        c.device = device;
        check(gsm_renderer_create(&c, &h_));
    }
    ~DepthFirstRenderer() { gsm_renderer_destroy(h_); }
    DepthFirstRenderer(const DepthFirstRenderer&) = delete;
    DepthFirstRenderer& operator=(const DepthFirstRenderer&) = delete;

    // commandBuffer = cudaStream_t; the caller synchronises it (commit + waitUntilCompleted)
    void render(void* commandBuffer, void* colorTexture, void* depthTexture, const GaussianInput& input,
                const CameraParams& camera, int width, int height) {
        gsm_camera c = camera.native();
        check(gsm_render(h_, commandBuffer, colorTexture, depthTexture, input.gaussians, input.harmonics,
                         (uint32_t)input.gaussianCount, (uint32_t)input.shComponents, &c, (uint32_t)width, (uint32_t)height));
    }
    void renderStereo(void* commandBuffer, const StereoRenderTarget& target, const GaussianInput& input,
                      const StereoCameraParams& camera, int width, int height) {
        if (target.isFoveated) {  // DepthFirstRenderer.swift:225-233: the eyes come from the configuration
            gsm_foveated_drawable d = target.drawable.native();
            gsm_stereo_configuration c = target.configuration.native();
            check(gsm_render_stereo_foveated(h_, commandBuffer, &d, input.gaussians, input.harmonics, (uint32_t)input.gaussianCount,
                                             (uint32_t)input.shComponents, &c, (uint32_t)width, (uint32_t)height));
            return;
        }
        gsm_camera l = camera.leftEye.native(), r = camera.rightEye.native();
        check(gsm_render_stereo(h_, commandBuffer, target.colorTexture, input.gaussians, input.harmonics,
                                (uint32_t)input.gaussianCount, (uint32_t)input.shComponents, &l, &r, (uint32_t)width,
                                (uint32_t)height));
    }
    // Host buffers (no device allocator on the caller's side): upload + frame + download. renderHostAsync only encodes,
    // like the reference's render() into a command buffer; waitHost() is the waitUntilCompleted. One frame in flight per
    // renderer -- use two renderers to overlap a frame's upload with the previous frame's render + download.
    void renderHost(const void* hostGaussians, const void* hostHarmonics, int gaussianCount, int shComponents,
                    const CameraParams& camera, int width, int height, void* hostColor, void* hostDepth = nullptr) {
        gsm_camera c = camera.native();
        check(gsm_render_host(h_, hostGaussians, hostHarmonics, (uint32_t)gaussianCount, (uint32_t)shComponents, &c,
                              (uint32_t)width, (uint32_t)height, hostColor, hostDepth));
    }
    void renderHostAsync(const void* hostGaussians, const void* hostHarmonics, int gaussianCount, int shComponents,
                         const CameraParams& camera, int width, int height, void* hostColor, void* hostDepth = nullptr) {
        gsm_camera c = camera.native();
        check(gsm_render_host_async(h_, hostGaussians, hostHarmonics, (uint32_t)gaussianCount, (uint32_t)shComponents, &c,
                                    (uint32_t)width, (uint32_t)height, hostColor, hostDepth));
    }
    void waitHost() { check(gsm_render_host_wait(h_)); }
    double lastGPUTime() const { return gsm_last_gpu_time_ms(h_) * 1e-3; }

    // white-box reads
    GSMDepthFirstHeader debugReadHeader() {
        GSMDepthFirstHeader h{};
        check(gsm_debug_read(h_, nullptr, GSM_DBG_HEADER, &h, 0, 1));
        return h;
    }
    uint32_t debugReadActiveTileCount() {
        uint32_t v = 0;
        check(gsm_debug_read(h_, nullptr, GSM_DBG_ACTIVE_TILE_COUNT, &v, 0, 1));
        return v;
    }
    template <typename T>
    std::vector<T> debugRead(gsm_debug_buffer which, size_t count, size_t first = 0) {
        if (sizeof(T) != gsm_debug_element_size(h_, which)) throw RendererError(GSM_ERR_INVALID_ARGUMENT, "element size mismatch");
        std::vector<T> v(count);
        if (count) check(gsm_debug_read(h_, nullptr, which, v.data(), first, count));
        return v;
    }
    std::vector<int32_t> debugReadSortedPrimitiveIndices(size_t n) { return debugRead<int32_t>(GSM_DBG_SORTED_PRIMITIVE_INDICES, n); }
    std::vector<uint32_t> debugReadDepthKeys(size_t n) { return debugRead<uint32_t>(GSM_DBG_DEPTH_KEYS, n); }
    std::vector<uint32_t> debugReadNTouchedTiles(size_t n) { return debugRead<uint32_t>(GSM_DBG_N_TOUCHED_TILES, n); }
    std::vector<uint32_t> debugReadInstanceOffsets(size_t n) { return debugRead<uint32_t>(GSM_DBG_INSTANCE_OFFSETS, n); }
    std::vector<int32_t> debugReadInstanceGaussianIndices(size_t n) { return debugRead<int32_t>(GSM_DBG_INSTANCE_GAUSSIAN_INDICES, n); }
    std::vector<GSMGaussianHeader> debugReadTileHeaders(size_t n) { return debugRead<GSMGaussianHeader>(GSM_DBG_TILE_HEADERS, n); }

    gsm_renderer* handle() const { return h_; }

private:
    gsm_renderer* h_ = nullptr;
};

// ---- GlobalRenderer (Sources/Renderer/GlobalRenderer/GlobalRenderer.swift:72-372): the reference's second GaussianRenderer,
// same RendererConfig and limits, 32 x 16 tiles, one [tile:16][half depth:16] sort. render() binds to gsm_render_global;
// renderStereo is a fatalError in the reference (:249-255) and throws here.
class GlobalRenderer {
public:
    explicit GlobalRenderer(int device = -1, RendererConfig config = RendererConfig()) {
        gsm_config c;
        gsm_config_default(&c);
        c.maxGaussians = (uint32_t)config.maxGaussians;
        c.maxWidth = (uint32_t)config.maxWidth;
        c.maxHeight = (uint32_t)config.maxHeight;
        c.precision = (uint32_t)config.precision;
        c.gaussianColorSpace = (uint32_t)config.gaussianColorSpace;
        c.device = device;
        check(gsm_renderer_create(&c, &h_));
    }
    ~GlobalRenderer() { gsm_renderer_destroy(h_); }
    GlobalRenderer(const GlobalRenderer&) = delete;
    GlobalRenderer& operator=(const GlobalRenderer&) = delete;

    void render(void* commandBuffer, void* colorTexture, void* depthTexture, const GaussianInput& input,
                const CameraParams& camera, int width, int height) {
        gsm_camera c = camera.native();
        check(gsm_render_global(h_, commandBuffer, colorTexture, depthTexture, input.gaussians, input.harmonics,
                                (uint32_t)input.gaussianCount, (uint32_t)input.shComponents, &c, (uint32_t)width, (uint32_t)height));
    }
    void renderStereo(void*, const StereoRenderTarget&, const GaussianInput&, const StereoCameraParams&, int, int) {
        throw RendererError(GSM_ERR_INVALID_ARGUMENT, "GlobalRenderer does not support stereo rendering. Use DepthFirstRenderer instead.");
    }
    gsm_global_header debugReadHeader() {
        gsm_global_header h{};
        check(gsm_global_debug_read(h_, nullptr, GSM_GDBG_HEADER, &h, 0, 1));
        return h;
    }
    uint32_t debugReadTotalAssignments() { return debugReadHeader().totalAssignments; }   // GlobalRenderer.swift:200-203
    std::vector<uint32_t> debugReadSortedKeys(size_t n) {
        std::vector<uint32_t> v(n);
        if (n) check(gsm_global_debug_read(h_, nullptr, GSM_GDBG_SORTED_KEYS, v.data(), 0, n));
        return v;
    }
    gsm_renderer* handle() const { return h_; }

private:
    gsm_renderer* h_ = nullptr;
};

// ---- scene ingest (gsm_scene.h): PLYLoader.load(url:) + GaussianSceneBuilder (PLYLoader.swift:246-281, Scene.swift:73-190).
// The caller supplies the file in host memory (mmap or read); the records are decoded on the device, straight into the
// renderer's input layout. PLYLoaderError cases arrive as RendererError with the GSM_ERR_PLY_* status.
struct GaussianDataset {
    void* gaussians = nullptr;   // device: PackedWorldGaussian (48 B) or PackedWorldGaussianHalf (32 B) x count
    void* harmonics = nullptr;   // device: float / half, count x harmonicsStride, planar per Gaussian
    gsm_scene_info info{};
    gsm_precision precision = GSM_PRECISION_FLOAT16;
    GaussianInput input() const { return GaussianInput{gaussians, harmonics, (int)info.count, (int)info.shComponents}; }
    void release() { if (gaussians) gsm_buffer_free(gaussians); if (harmonics) gsm_buffer_free(harmonics); gaussians = harmonics = nullptr; }
};

struct PLYLoader {
    static GaussianDataset load(const void* fileBytes, size_t fileSize, int device = -1, void* stream = nullptr,
                                gsm_precision precision = GSM_PRECISION_FLOAT16) {
        gsm_ply_info probe{};
        gsm_status st = gsm_ply_probe(fileBytes, fileSize, &probe);
        if (st != GSM_OK) throw RendererError(st, gsm_last_error_string());
        GaussianDataset d;
        d.precision = precision;
        const size_t n = probe.vertexCount ? probe.vertexCount : 1, sh = probe.shProperties ? probe.shProperties : 1;
        const size_t rec = precision == GSM_PRECISION_FLOAT16 ? 32 : 48, el = precision == GSM_PRECISION_FLOAT16 ? 2 : 4;
        if ((st = gsm_buffer_alloc(device, n * rec, &d.gaussians)) != GSM_OK) throw RendererError(st, gsm_last_error_string());
        if ((st = gsm_buffer_alloc(device, n * sh * el, &d.harmonics)) != GSM_OK) { d.release(); throw RendererError(st, gsm_last_error_string()); }
        st = gsm_ply_load(device, stream, fileBytes, fileSize, (int)precision, d.gaussians, d.harmonics, probe.vertexCount,
                          (size_t)probe.vertexCount * sh, &d.info);
        if (st != GSM_OK) { d.release(); throw RendererError(st, gsm_last_error_string()); }
        return d;
    }
};

struct GaussianSceneBuilder {
    static void sortByMortonCode(GaussianDataset& d, int device = -1, void* stream = nullptr) {
        gsm_status st = gsm_scene_morton_sort(device, stream, d.gaussians, d.harmonics, d.info.count, d.info.harmonicsStride, (int)d.precision);
        if (st != GSM_OK) throw RendererError(st, gsm_last_error_string());
    }
};

}  // namespace gsm
