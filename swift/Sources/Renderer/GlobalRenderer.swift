// GlobalRenderer.swift -- Swift facade of the reference's second GaussianRenderer
// (Sources/Renderer/GlobalRenderer/GlobalRenderer.swift:72-372) over gsm_render_global / gsm_global_debug_read of
// include/gsm/gsm.h: 32 x 16-pixel tiles of the renderer's limits, one sort of [tile:16][half depth:16] keys, four assignments
// per Gaussian of capacity. Same RendererConfig, GaussianInput, CameraParams, CommandBuffer and DeviceBuffer as
// DepthFirstRenderer.swift. Thin on purpose: every call below is one gsm_* call.
import CGSM
import RendererTypes

public final class GlobalRenderer: GaussianRenderer, @unchecked Sendable {
    private let handle: OpaquePointer
    private let config: RendererConfig
    public var lastGPUTime: Double? { nil }

    public init(device: Int32? = nil, config: RendererConfig = RendererConfig()) throws {
        var c = gsm_config()
        gsm_config_default(&c)
        c.maxGaussians = UInt32(config.maxGaussians)
        c.maxWidth = UInt32(config.maxWidth)
        c.maxHeight = UInt32(config.maxHeight)
        c.precision = config.precision.rawValue
        c.gaussianColorSpace = config.gaussianColorSpace.rawValue
        c.device = device ?? -1
        var h: OpaquePointer?
        let s = gsm_renderer_create(&c, &h)
        guard s == GSM_OK, let h else { throw RendererError.from(s, config: config) }
        self.handle = h
        self.config = config
    }

    deinit { gsm_renderer_destroy(handle) }

    /// GlobalRenderer.swift:206-247. Frames whose Gaussian count exceeds the limits encode nothing (validateLimits, :293-297).
    public func render(commandBuffer: CommandBuffer, colorTexture: DeviceBuffer, depthTexture: DeviceBuffer?,
                       input: GaussianInput, camera: CameraParams, width: Int, height: Int) {
        var cam = camera.native()
        _ = gsm_render_global(handle, commandBuffer.stream, colorTexture.pointer, depthTexture?.pointer, input.gaussians.pointer,
                              input.harmonics.pointer, UInt32(input.gaussianCount), UInt32(input.shComponents), &cam,
                              UInt32(width), UInt32(height))
    }

    /// GlobalRenderer.swift:249-255: the reference's GlobalRenderer does not render stereo.
    public func renderStereo(commandBuffer: CommandBuffer, target: StereoRenderTarget, input: GaussianInput,
                             camera: StereoCameraParams, width: Int, height: Int) {
        fatalError("GlobalRenderer does not support stereo rendering. Use DepthFirstRenderer instead.")
    }

    /// GlobalRenderer.swift:200-203
    public func debugReadTotalAssignments() -> UInt32 {
        var h = gsm_global_header()
        _ = gsm_global_debug_read(handle, nil, Int32(GSM_GDBG_HEADER.rawValue), &h, 0, 1)
        return h.totalAssignments
    }

    public func debugReadSortedKeys(count: Int) -> [UInt32] {
        var v = [UInt32](repeating: 0, count: count)
        if count > 0 { _ = gsm_global_debug_read(handle, nil, Int32(GSM_GDBG_SORTED_KEYS.rawValue), &v, 0, count) }
        return v
    }
}
