// DepthFirstRenderer.swift -- Swift facade with the reference's public surface
// (Sources/Renderer/Shared/GaussianRendererProtocol.swift, Sources/Renderer/DepthFirstRenderer/DepthFirstRenderer.swift)
// over the C ABI of include/gsm/gsm.h. MTLCommandBuffer -> CommandBuffer (a CUDA stream), MTLBuffer/MTLTexture ->
// DeviceBuffer (a device pointer + length). Thin on purpose: every call below is one gsm_* call.
import CGSM
import RendererTypes

public enum RenderPrecision: UInt32, Sendable { case float32 = 0, float16 = 1 }
public enum RadixSortKeyPrecision: UInt32, Sendable {
    case bits16 = 16, bits32 = 32
    public var numPasses: Int { self == .bits16 ? 2 : 4 }
}

public enum RendererError: Error, Sendable, CustomStringConvertible {
    case deviceNotAvailable
    case failedToCreatePipeline(String)
    case failedToAllocateBuffer(label: String, size: Int)
    case invalidGaussianCount(provided: Int, maximum: Int)
    case invalidDimensions(width: Int, height: Int, maxWidth: Int, maxHeight: Int)
    case invalidTileCount(provided: Int, maximum: Int)
    case renderFailed(String)

    static func from(_ s: gsm_status, config: RendererConfig? = nil) -> RendererError {
        let detail = String(cString: gsm_last_error_string())
        switch s {
        case GSM_ERR_DEVICE_NOT_AVAILABLE: return .deviceNotAvailable
        case GSM_ERR_FAILED_TO_CREATE_PIPELINE: return .failedToCreatePipeline(detail)
        case GSM_ERR_FAILED_TO_ALLOCATE_BUFFER: return .failedToAllocateBuffer(label: detail, size: 0)
        case GSM_ERR_INVALID_GAUSSIAN_COUNT:
            return .invalidGaussianCount(provided: config?.maxGaussians ?? 0, maximum: 30_000_000)
        case GSM_ERR_INVALID_DIMENSIONS:
            return .invalidDimensions(width: 0, height: 0, maxWidth: config?.maxWidth ?? 0, maxHeight: config?.maxHeight ?? 0)
        case GSM_ERR_INVALID_TILE_COUNT: return .invalidTileCount(provided: 0, maximum: 65535)
        default: return .renderFailed(detail)
        }
    }

    public var description: String {
        switch self {
        case .deviceNotAvailable: "CUDA device not available"
        case let .failedToCreatePipeline(n): "Failed to create pipeline: \(n)"
        case let .failedToAllocateBuffer(l, s): "Failed to allocate buffer '\(l)' with size \(s) bytes"
        case let .invalidGaussianCount(p, m): "Gaussian count \(p) exceeds maximum \(m)"
        case let .invalidDimensions(w, h, mw, mh): "Dimensions \(w)x\(h) exceed maximum \(mw)x\(mh)"
        case let .invalidTileCount(p, m): "Tile count \(p) exceeds maximum \(m)"
        case let .renderFailed(r): "Render failed: \(r)"
        }
    }
}

/// MTLBuffer / MTLTexture replacement: device memory owned by the caller.
public final class DeviceBuffer: @unchecked Sendable {
    public let pointer: UnsafeMutableRawPointer
    public let length: Int
    public init(length: Int, device: Int32 = -1) throws {
        var p: UnsafeMutableRawPointer?
        let s = gsm_buffer_alloc(device, length, &p)
        guard s == GSM_OK, let p else { throw RendererError.failedToAllocateBuffer(label: "DeviceBuffer", size: length) }
        self.pointer = p
        self.length = length
    }
    public convenience init<T>(bytes: [T], queue: CommandBuffer, device: Int32 = -1) throws {
        try self.init(length: bytes.count * MemoryLayout<T>.stride, device: device)
        try bytes.withUnsafeBytes { raw in
            guard gsm_buffer_upload(pointer, raw.baseAddress, raw.count, queue.stream) == GSM_OK else {
                throw RendererError.renderFailed("upload")
            }
        }
    }
    public func download(into dst: UnsafeMutableRawPointer, queue: CommandBuffer) throws {
        guard gsm_buffer_download(dst, pointer, length, queue.stream) == GSM_OK else { throw RendererError.renderFailed("download") }
    }
    deinit { _ = gsm_buffer_free(pointer) }
}

/// MTLCommandBuffer replacement: a CUDA stream. `waitUntilCompleted()` is the caller's synchronisation point.
public final class CommandBuffer: @unchecked Sendable {
    public let stream: UnsafeMutableRawPointer?
    public init(device: Int32 = -1) throws {
        var s: UnsafeMutableRawPointer?
        guard gsm_stream_create(device, &s) == GSM_OK else { throw RendererError.deviceNotAvailable }
        self.stream = s
    }
    public func commit() {}
    public func waitUntilCompleted() { _ = gsm_stream_synchronize(stream) }
    deinit { _ = gsm_stream_destroy(stream) }
}

public struct GaussianInput: Sendable {
    public let gaussians: DeviceBuffer  // PackedWorldGaussian (48 B) or PackedWorldGaussianHalf (32 B)
    public let harmonics: DeviceBuffer
    public let gaussianCount: Int
    public let shComponents: Int
    public init(gaussians: DeviceBuffer, harmonics: DeviceBuffer, gaussianCount: Int, shComponents: Int) {
        self.gaussians = gaussians; self.harmonics = harmonics
        self.gaussianCount = gaussianCount; self.shComponents = shComponents
    }
}

public struct CameraParams: Sendable {
    public let viewMatrix: [Float]        // 16 floats, column-major (simd_float4x4 layout)
    public let projectionMatrix: [Float]
    public let position: SIMD3<Float>
    public let focalX: Float
    public let focalY: Float
    public let near: Float
    public let far: Float
    public init(viewMatrix: [Float], projectionMatrix: [Float], position: SIMD3<Float>, focalX: Float, focalY: Float,
                near: Float = 0.1, far: Float = 10.0) {
        self.viewMatrix = viewMatrix; self.projectionMatrix = projectionMatrix; self.position = position
        self.focalX = focalX; self.focalY = focalY; self.near = near; self.far = far
    }
    func native() -> gsm_camera {
        var c = gsm_camera()
        withUnsafeMutableBytes(of: &c.viewMatrix) { $0.copyBytes(from: viewMatrix.withUnsafeBytes { Array($0) }) }
        withUnsafeMutableBytes(of: &c.projectionMatrix) { $0.copyBytes(from: projectionMatrix.withUnsafeBytes { Array($0) }) }
        c.position = (position.x, position.y, position.z)
        c.focalX = focalX; c.focalY = focalY; c.nearPlane = near; c.farPlane = far
        return c
    }
}

public struct StereoCameraParams: Sendable {
    public let leftEye: CameraParams
    public let rightEye: CameraParams
    public init(leftEye: CameraParams, rightEye: CameraParams) { self.leftEye = leftEye; self.rightEye = rightEye }
}

public struct RendererConfig: Sendable {
    public enum GaussianColorSpace: UInt32, Sendable { case linear = 0, srgb = 1 }
    public let maxGaussians: Int
    public let maxWidth: Int
    public let maxHeight: Int
    public let precision: RenderPrecision
    public let gaussianColorSpace: GaussianColorSpace
    public let backToFront: Bool
    public init(maxGaussians: Int = 6_000_000, maxWidth: Int = 1920, maxHeight: Int = 1080,
                precision: RenderPrecision = .float16, gaussianColorSpace: GaussianColorSpace = .srgb,
                backToFront: Bool = false) {
        self.maxGaussians = maxGaussians; self.maxWidth = maxWidth; self.maxHeight = maxHeight
        self.precision = precision; self.gaussianColorSpace = gaussianColorSpace; self.backToFront = backToFront
    }
}

/// MTLViewport (znear / zfar do not matter to the copy)
public struct Viewport: Sendable {
    public let originX, originY, width, height: Double
    public init(originX: Double, originY: Double, width: Double, height: Double) {
        self.originX = originX; self.originY = originY; self.width = width; self.height = height
    }
}

public struct EyeView: Sendable {
    public let viewport: Viewport
    public let viewMatrix: [Float]
    public let projectionMatrix: [Float]
    public let cameraPosition: SIMD3<Float>
    public let focalX, focalY, near, far: Float
    public init(viewport: Viewport, viewMatrix: [Float], projectionMatrix: [Float], cameraPosition: SIMD3<Float>,
                focalX: Float, focalY: Float, near: Float = 0.1, far: Float = 10.0) {
        self.viewport = viewport; self.viewMatrix = viewMatrix; self.projectionMatrix = projectionMatrix
        self.cameraPosition = cameraPosition; self.focalX = focalX; self.focalY = focalY; self.near = near; self.far = far
    }
    func native() -> gsm_eye_view {
        var e = gsm_eye_view()
        e.viewport = gsm_viewport(originX: viewport.originX, originY: viewport.originY, width: viewport.width, height: viewport.height)
        e.camera = CameraParams(viewMatrix: viewMatrix, projectionMatrix: projectionMatrix, position: cameraPosition,
                                focalX: focalX, focalY: focalY, near: near, far: far).native()
        return e
    }
}

public struct StereoConfiguration: Sendable {
    public let leftEye: EyeView
    public let rightEye: EyeView
    public let sceneTransform: [Float]   // 16 floats, column-major; identity by default
    public init(leftEye: EyeView, rightEye: EyeView,
                sceneTransform: [Float] = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]) {
        self.leftEye = leftEye; self.rightEye = rightEye; self.sceneTransform = sceneTransform
    }
    func native() -> gsm_stereo_configuration {
        var c = gsm_stereo_configuration()
        c.leftEye = leftEye.native(); c.rightEye = rightEye.native()
        withUnsafeMutableBytes(of: &c.sceneTransform) { $0.copyBytes(from: sceneTransform.withUnsafeBytes { Array($0) }) }
        return c
    }
}

/// One layer of MTLRasterizationRateMap, tabulated with its own mapPhysicalToScreenCoordinates: the screen coordinate of the
/// centre of every physical column / row.
public struct RasterizationRateLayer: Sendable {
    public let screenX: [Float]
    public let screenY: [Float]
    public init(screenX: [Float], screenY: [Float]) { self.screenX = screenX; self.screenY = screenY }
}

public struct FoveatedStereoDrawable: Sendable {
    public let colorTexture: DeviceBuffer
    public let textureWidth, textureHeight, arrayLength: Int
    public let rasterizationRateMap: [RasterizationRateLayer]?
    public let colorPixelFormat: gsm_pixel_format
    public init(colorTexture: DeviceBuffer, textureWidth: Int, textureHeight: Int, arrayLength: Int = 2,
                rasterizationRateMap: [RasterizationRateLayer]?, colorPixelFormat: gsm_pixel_format = GSM_PIXEL_BGRA8_SRGB) {
        self.colorTexture = colorTexture; self.textureWidth = textureWidth; self.textureHeight = textureHeight
        self.arrayLength = arrayLength; self.rasterizationRateMap = rasterizationRateMap; self.colorPixelFormat = colorPixelFormat
    }
}

public enum StereoRenderTarget: Sendable {
    /// left eye on the left half, right eye on the right half of an rgba16f (2*width) x height buffer
    case sideBySide(colorTexture: DeviceBuffer, depthTexture: DeviceBuffer?)
    /// a drawable with per-eye viewports and an optional rasterization-rate map
    case foveated(drawable: FoveatedStereoDrawable, configuration: StereoConfiguration)
}

public protocol GaussianRenderer: AnyObject, Sendable {
    var lastGPUTime: Double? { get }
    func render(commandBuffer: CommandBuffer, colorTexture: DeviceBuffer, depthTexture: DeviceBuffer?, input: GaussianInput,
                camera: CameraParams, width: Int, height: Int)
    func renderStereo(commandBuffer: CommandBuffer, target: StereoRenderTarget, input: GaussianInput,
                      camera: StereoCameraParams, width: Int, height: Int)
}

public final class DepthFirstRenderer: GaussianRenderer, @unchecked Sendable {
    private let handle: OpaquePointer
    private let config: RendererConfig
    public var lastGPUTime: Double? {
        let ms = gsm_last_gpu_time_ms(handle)
        return ms < 0 ? nil : ms * 1e-3
    }

    public init(device: Int32? = nil, config: RendererConfig = RendererConfig(),
                depthSortKeyPrecision: RadixSortKeyPrecision = .bits32,
                tileIdPrecision: RadixSortKeyPrecision = .bits16) throws {
        var c = gsm_config()
        gsm_config_default(&c)
        c.maxGaussians = UInt32(config.maxGaussians)
        c.maxWidth = UInt32(config.maxWidth)
        c.maxHeight = UInt32(config.maxHeight)
        c.precision = config.precision.rawValue
        c.gaussianColorSpace = config.gaussianColorSpace.rawValue
        c.depthSortKeyPrecision = depthSortKeyPrecision.rawValue
        c.tileIdPrecision = tileIdPrecision.rawValue
        c.device = device ?? -1
        var h: OpaquePointer?
        let s = gsm_renderer_create(&c, &h)
        guard s == GSM_OK, let h else { throw RendererError.from(s, config: config) }
        self.handle = h
        self.config = config
    }

    deinit { gsm_renderer_destroy(handle) }

    /// Like the reference, failures after init are silent (DepthFirstRenderer.swift:189,249); inspect
    /// gsm_last_error_string() when debugging.
    public func render(commandBuffer: CommandBuffer, colorTexture: DeviceBuffer, depthTexture: DeviceBuffer?,
                       input: GaussianInput, camera: CameraParams, width: Int, height: Int) {
        var cam = camera.native()
        _ = gsm_render(handle, commandBuffer.stream, colorTexture.pointer, depthTexture?.pointer, input.gaussians.pointer,
                       input.harmonics.pointer, UInt32(input.gaussianCount), UInt32(input.shComponents), &cam,
                       UInt32(width), UInt32(height))
    }

    /// Host buffers: upload + frame + download on the renderer's own stream. `renderHostAsync` only encodes (the shape of
    /// render() + commit()); `waitHost()` is waitUntilCompleted. The buffers must stay valid until the wait.
    public func renderHostAsync(gaussians: UnsafeRawPointer, harmonics: UnsafeRawPointer, gaussianCount: Int, shComponents: Int,
                                camera: CameraParams, width: Int, height: Int, color: UnsafeMutableRawPointer,
                                depth: UnsafeMutableRawPointer?) {
        var cam = camera.native()
        _ = gsm_render_host_async(handle, gaussians, harmonics, UInt32(gaussianCount), UInt32(shComponents), &cam,
                                  UInt32(width), UInt32(height), color, depth)
    }
    public func waitHost() { _ = gsm_render_host_wait(handle) }

    public func renderStereo(commandBuffer: CommandBuffer, target: StereoRenderTarget, input: GaussianInput,
                             camera: StereoCameraParams, width: Int, height: Int) {
        if case let .foveated(drawable, configuration) = target {
            renderFoveated(commandBuffer: commandBuffer, drawable: drawable, configuration: configuration, input: input,
                           width: width, height: height)
            return
        }
        guard case let .sideBySide(colorTexture, _) = target else { return }
        var l = camera.leftEye.native(), r = camera.rightEye.native()
        _ = gsm_render_stereo(handle, commandBuffer.stream, colorTexture.pointer, input.gaussians.pointer,
                              input.harmonics.pointer, UInt32(input.gaussianCount), UInt32(input.shComponents), &l, &r,
                              UInt32(width), UInt32(height))
    }

    private func renderFoveated(commandBuffer: CommandBuffer, drawable: FoveatedStereoDrawable, configuration: StereoConfiguration,
                                input: GaussianInput, width: Int, height: Int) {
        let px = drawable.colorPixelFormat == GSM_PIXEL_RGBA16F ? 8 : 4
        var d = gsm_foveated_drawable()
        d.colorTexture = drawable.colorTexture.pointer
        d.textureWidth = UInt32(drawable.textureWidth); d.textureHeight = UInt32(drawable.textureHeight)
        d.arrayLength = UInt32(drawable.arrayLength)
        d.rowBytes = drawable.textureWidth * px; d.sliceBytes = d.rowBytes * drawable.textureHeight
        d.colorPixelFormat = drawable.colorPixelFormat.rawValue
        var cfg = configuration.native()
        func submit(_ map: UnsafePointer<gsm_rate_map>?) {
            d.rasterizationRateMap = map
            _ = gsm_render_stereo_foveated(handle, commandBuffer.stream, &d, input.gaussians.pointer, input.harmonics.pointer,
                                           UInt32(input.gaussianCount), UInt32(input.shComponents), &cfg, UInt32(width), UInt32(height))
        }
        guard let layers = drawable.rasterizationRateMap, let first = layers.first else { submit(nil); return }
        let second = layers.count > 1 ? layers[1] : first
        first.screenX.withUnsafeBufferPointer { x0 in first.screenY.withUnsafeBufferPointer { y0 in
        second.screenX.withUnsafeBufferPointer { x1 in second.screenY.withUnsafeBufferPointer { y1 in
            var m = gsm_rate_map()
            m.layerCount = UInt32(min(layers.count, 2))
            m.layers.0 = gsm_rate_map_layer(physicalWidth: UInt32(x0.count), physicalHeight: UInt32(y0.count),
                                            screenX: x0.baseAddress, screenY: y0.baseAddress)
            m.layers.1 = gsm_rate_map_layer(physicalWidth: UInt32(x1.count), physicalHeight: UInt32(y1.count),
                                            screenX: x1.baseAddress, screenY: y1.baseAddress)
            withUnsafePointer(to: &m) { submit($0) }
        } } } }
    }

    // debugRead* (Tests/RendererTests/DepthFirstUnitTests.swift:911-1252)
    public func debugReadHeader() -> GSMDepthFirstHeader {
        var h = GSMDepthFirstHeader()
        _ = gsm_debug_read(handle, nil, Int32(GSM_DBG_HEADER.rawValue), &h, 0, 1)
        return h
    }
    public func debugReadActiveTileCount() -> UInt32 {
        var v: UInt32 = 0
        _ = gsm_debug_read(handle, nil, Int32(GSM_DBG_ACTIVE_TILE_COUNT.rawValue), &v, 0, 1)
        return v
    }
    public func debugRead<T>(_ which: gsm_debug_buffer, count: Int, first: Int = 0, as _: T.Type) -> [T] {
        guard count > 0 else { return [] }
        return [T](unsafeUninitializedCapacity: count) { buf, n in
            _ = gsm_debug_read(handle, nil, Int32(which.rawValue), buf.baseAddress, first, count)
            n = count
        }
    }
    public func debugReadSortedPrimitiveIndices(count: Int) -> [Int32] { debugRead(GSM_DBG_SORTED_PRIMITIVE_INDICES, count: count, as: Int32.self) }
    public func debugReadDepthKeys(count: Int) -> [UInt32] { debugRead(GSM_DBG_DEPTH_KEYS, count: count, as: UInt32.self) }
    public func debugReadNTouchedTiles(count: Int) -> [UInt32] { debugRead(GSM_DBG_N_TOUCHED_TILES, count: count, as: UInt32.self) }
    public func debugReadInstanceOffsets(count: Int) -> [UInt32] { debugRead(GSM_DBG_INSTANCE_OFFSETS, count: count, as: UInt32.self) }
    public func debugReadInstanceGaussianIndices(count: Int) -> [Int32] { debugRead(GSM_DBG_INSTANCE_GAUSSIAN_INDICES, count: count, as: Int32.self) }
    public func debugReadTileHeaders(count: Int) -> [GSMGaussianHeader] { debugRead(GSM_DBG_TILE_HEADERS, count: count, as: GSMGaussianHeader.self) }
}
