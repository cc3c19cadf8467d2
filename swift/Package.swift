// swift-tools-version: 5.9
// Swift-on-Linux facade over libgsm_b200.so (BASELINE.json north_star). SOURCE ONLY: swiftc is not in the build
// image, so this package has not been compiled here; the same C ABI is exercised by the C++ facade
// (include/gsm/DepthFirstRenderer.hpp, tests/cpp) and the Python mirror (gsm_renderer_b200/renderer.py).
// Build on a Linux box with Swift >= 5.9 and the library built in-tree:
//   swift build -Xlinker -L../gsm_renderer_b200/lib -Xlinker -rpath -Xlinker $PWD/../gsm_renderer_b200/lib
import PackageDescription

let package = Package(
    name: "GSMRendererB200",
    products: [.library(name: "Renderer", targets: ["Renderer"])],
    targets: [
        // the reference's RendererTypes C module (Sources/RendererTypes), Linux restatement: include/gsm/gsm_types.h
        .target(name: "RendererTypes", path: "Sources/RendererTypes", publicHeadersPath: "include"),
        // the C ABI (include/gsm/gsm.h) as a system library
        .systemLibrary(name: "CGSM", path: "Sources/CGSM"),
        .target(name: "Renderer", dependencies: ["RendererTypes", "CGSM"], path: "Sources/Renderer"),
    ]
)
