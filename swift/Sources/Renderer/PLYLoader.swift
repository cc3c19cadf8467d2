// PLYLoader / GaussianSceneBuilder over the C ABI of gsm_scene.h (source only: swiftc is not in the build image).
// Mirrors Sources/Renderer/Utils/PLYLoader.swift:246-281 and Scene.swift:73-190 of the reference: the file is mapped on the
// host and decoded ON THE DEVICE, straight into the renderer's input layout (PackedWorldGaussianHalf + planar Float16 SH).
import CGSM
import Foundation

public enum PLYLoaderError: Error, LocalizedError {
    case invalidHeader, unsupportedFormat, missingVertexElement, missingRequiredProperties, listPropertiesNotSupported
    case insufficientData, missingChunkElement
    case renderer(RendererError)

    static func from(_ status: gsm_status) -> PLYLoaderError {
        switch Int(status.rawValue) {
        case 20: return .invalidHeader
        case 21: return .unsupportedFormat
        case 22: return .missingVertexElement
        case 23: return .missingRequiredProperties
        case 24: return .listPropertiesNotSupported
        case 25: return .insufficientData
        case 26: return .missingChunkElement
        default: return .renderer(RendererError.from(status))
        }
    }
    public var errorDescription: String? { String(cString: gsm_last_error_string()) }
}

/// GaussianDataset (Scene.swift:141-157) with the records already packed in device memory.
public struct GaussianDataset {
    public let gaussians: DeviceBuffer
    public let harmonics: DeviceBuffer
    public let count: Int
    public let shComponents: Int
    public let harmonicsStride: Int
    public let boundsCenter: SIMD3<Float>
    public let boundsRadius: Float
    public var input: GaussianInput { GaussianInput(gaussians: gaussians, harmonics: harmonics, gaussianCount: count, shComponents: shComponents) }
}

public enum PLYLoader {
    public static func load(url: URL, device: Int32 = -1, commandBuffer: CommandBuffer) throws -> GaussianDataset {
        let data = try Data(contentsOf: url, options: [.mappedIfSafe])
        return try data.withUnsafeBytes { raw in
            var probe = gsm_ply_info()
            var st = gsm_ply_probe(raw.baseAddress, raw.count, &probe)
            guard st == GSM_OK else { throw PLYLoaderError.from(st) }
            let n = max(Int(probe.vertexCount), 1), sh = max(Int(probe.shProperties), 1)
            let g = try DeviceBuffer(device: device, length: n * 32), h = try DeviceBuffer(device: device, length: n * sh * 2)
            var info = gsm_scene_info()
            st = gsm_ply_load(device, commandBuffer.stream, raw.baseAddress, raw.count, Int32(GSM_PRECISION_FLOAT16.rawValue), g.pointer,
                              h.pointer, probe.vertexCount, Int(probe.vertexCount) * sh, &info)
            guard st == GSM_OK else { throw PLYLoaderError.from(st) }
            return GaussianDataset(gaussians: g, harmonics: h, count: Int(info.count), shComponents: Int(info.shComponents),
                                   harmonicsStride: Int(info.harmonicsStride),
                                   boundsCenter: SIMD3<Float>(info.boundsCenter.0, info.boundsCenter.1, info.boundsCenter.2),
                                   boundsRadius: info.boundsRadius)
        }
    }
}

public enum GaussianSceneBuilder {
    /// GaussianSceneBuilder.sortByMortonCode (Scene.swift:73-138), in place on the device buffers.
    public static func sortByMortonCode(_ d: GaussianDataset, device: Int32 = -1, commandBuffer: CommandBuffer) {
        _ = gsm_scene_morton_sort(device, commandBuffer.stream, d.gaussians.pointer, d.harmonics.pointer, UInt32(d.count),
                                  UInt32(d.harmonicsStride), Int32(GSM_PRECISION_FLOAT16.rawValue))
    }
    public static func bounds(of d: GaussianDataset) -> (center: SIMD3<Float>, radius: Float) { (d.boundsCenter, d.boundsRadius) }
}
