// facade_driver.cpp -- the reference's testDepthFirstPipelineStages (Tests/RendererTests/DepthFirstUnitTests.swift:21-117)
// restated against the C++ facade: same scene, same camera, same assertions. Also the host-language proof that
// the C ABI is usable without Python or torch (device memory and the stream come from the ABI itself).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gsm/DepthFirstRenderer.hpp"

static std::array<float, 16> makeProjectionMatrix(int width, int height, float near, float far, float fovDegrees) {
    // TestUtils.swift:37-71, OpenCV convention, column-major
    float aspect = float(width) / float(height);
    float fov = fovDegrees * float(M_PI) / 180.0f;
    float f = 1.0f / std::tan(fov / 2.0f);
    std::array<float, 16> m{};
    m[0] = f / aspect;
    m[5] = f;
    m[10] = far / (far - near); m[11] = 1.0f;
    m[14] = -(far * near) / (far - near);
    return m;
}

int main() {
    const int gaussianCount = 1000, width = 640, height = 480;
    std::vector<GSMPackedWorldGaussian> packed(gaussianCount);
    std::vector<float> harmonics;
    for (int i = 0; i < gaussianCount; ++i) {
        int row = i / 32, col = i % 32;
        GSMPackedWorldGaussian g{};
        g.px = float(col) * 0.1f - 1.6f; g.py = float(row) * 0.1f - 1.6f; g.pz = 2.0f + float(i) * 0.001f;
        g.sx = g.sy = g.sz = 0.01f; g.opacity = 0.8f;
        g.rotation.x = 0; g.rotation.y = 0; g.rotation.z = 0; g.rotation.w = 1;
        packed[i] = g;
        harmonics.push_back(float(i % 10) / 10.0f);
        harmonics.push_back(float((i / 10) % 10) / 10.0f);
        harmonics.push_back(float((i / 100) % 10) / 10.0f);
    }
    try {
        gsm::RendererConfig config;
        config.maxGaussians = gaussianCount; config.maxWidth = width; config.maxHeight = height;
        config.precision = gsm::RenderPrecision::float32;
        gsm::DepthFirstRenderer renderer(-1, config);

        void *gaussianBuf, *harmonicsBuf, *colorTexture, *depthTexture, *queue;
        gsm::check(gsm_stream_create(-1, &queue));
        gsm::check(gsm_buffer_alloc(-1, packed.size() * sizeof(GSMPackedWorldGaussian), &gaussianBuf));
        gsm::check(gsm_buffer_alloc(-1, harmonics.size() * sizeof(float), &harmonicsBuf));
        gsm::check(gsm_buffer_alloc(-1, size_t(width) * height * 8, &colorTexture));
        gsm::check(gsm_buffer_alloc(-1, size_t(width) * height * 2, &depthTexture));
        gsm::check(gsm_buffer_upload(gaussianBuf, packed.data(), packed.size() * sizeof(GSMPackedWorldGaussian), queue));
        gsm::check(gsm_buffer_upload(harmonicsBuf, harmonics.data(), harmonics.size() * sizeof(float), queue));

        gsm::CameraParams camera;
        camera.viewMatrix = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        camera.projectionMatrix = makeProjectionMatrix(width, height, 0.1f, 10.0f, 60.0f);
        camera.position = {0, 0, 0};
        camera.focalX = float(width) * 1.5f; camera.focalY = float(height) * 1.5f;
        gsm::GaussianInput input{gaussianBuf, harmonicsBuf, gaussianCount, 1};

        renderer.render(queue, colorTexture, depthTexture, input, camera, width, height);
        gsm::check(gsm_stream_synchronize(queue));  // cb.commit(); cb.waitUntilCompleted()

        GSMDepthFirstHeader header = renderer.debugReadHeader();
        std::printf("visible=%u instances=%u paddedVisible=%u paddedInstances=%u overflow=%u activeTiles=%u\n", header.visibleCount,
                    header.totalInstances, header.paddedVisibleCount, header.paddedInstanceCount, header.overflow,
                    renderer.debugReadActiveTileCount());
        bool ok = header.overflow == 0 && header.visibleCount > 0 && header.visibleCount <= (uint32_t)gaussianCount &&
                  header.totalInstances > 0;
        std::vector<uint16_t> px(size_t(width) * height * 4);
        gsm::check(gsm_buffer_download(px.data(), colorTexture, px.size() * 2, queue));
        size_t nonBlack = 0;
        for (size_t i = 0; i < px.size(); i += 4) nonBlack += (px[i] | px[i + 1] | px[i + 2]) & 0x7FFF ? 1 : 0;
        std::printf("nonBlackPixels=%zu\n", nonBlack);
        // the same scene through the reference's second renderer (GlobalRenderer.swift): 32 x 16 tiles, one sort of [tile | depth] keys
        bool okGlobal = false;
        {
            gsm::GlobalRenderer global(-1, config);
            global.render(queue, colorTexture, depthTexture, input, camera, width, height);
            gsm::check(gsm_stream_synchronize(queue));
            gsm_global_header gh = global.debugReadHeader();
            std::vector<uint32_t> keys = global.debugReadSortedKeys(gh.totalAssignments);
            bool sorted = true;
            for (size_t i = 1; i < keys.size(); ++i) sorted = sorted && keys[i - 1] <= keys[i];
            std::vector<uint16_t> gpx(size_t(width) * height * 4);
            gsm::check(gsm_buffer_download(gpx.data(), colorTexture, gpx.size() * 2, queue));
            size_t nb = 0;
            for (size_t i = 0; i < gpx.size(); i += 4) nb += (gpx[i] | gpx[i + 1] | gpx[i + 2]) & 0x7FFF ? 1 : 0;
            std::printf("global visible=%u assignments=%u overflow=%u activeTiles=%u sorted=%d nonBlackPixels=%zu\n", gh.visibleCount,
                        gh.totalAssignments, gh.overflow, gh.activeTileCount, (int)sorted, nb);
            okGlobal = gh.overflow == 0 && gh.visibleCount > 0 && gh.totalAssignments >= gh.visibleCount && sorted && nb > 0;
        }
        ok = ok && okGlobal;
        gsm_buffer_free(gaussianBuf); gsm_buffer_free(harmonicsBuf); gsm_buffer_free(colorTexture); gsm_buffer_free(depthTexture);
        gsm_stream_destroy(queue);
        std::printf(ok && nonBlack > 0 ? "PASS\n" : "FAIL\n");
        return ok && nonBlack > 0 ? 0 : 1;
    } catch (const gsm::RendererError& e) {
        std::printf("RendererError(%d): %s\n", (int)e.status, e.what());
        return e.status == GSM_ERR_DEVICE_NOT_AVAILABLE ? 3 : 2;
    }
}
