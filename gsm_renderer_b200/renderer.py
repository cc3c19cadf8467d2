"""Host-side mirror of the reference's public surface for the DepthFirst path.

Same names, argument meaning and error behaviour as
Sources/Renderer/Shared/GaussianRendererProtocol.swift (RenderPrecision :4-7, GaussianInput :9-26,
CameraParams :28-54, StereoCameraParams :56-67, RendererConfig :195-228, StereoRenderTarget :233-239,
GaussianRenderer :243-272, RendererError :274-324) and
Sources/Renderer/DepthFirstRenderer/DepthFirstRenderer.swift (init :45-101, render :166-203,
renderStereo :205-235), over the C ABI of include/gsm/gsm.h.

Metal objects map to: MTLCommandBuffer -> a CUDA stream (torch.cuda.Stream, a raw cudaStream_t int, or
None); MTLBuffer / MTLTexture -> anything with .data_ptr() (torch CUDA tensors) or a raw device pointer.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import _native as N


class RenderPrecision(enum.Enum):
    float32 = 0
    float16 = 1


class GaussianColorSpace(enum.IntEnum):
    linear = 0
    srgb = 1


class RadixSortKeyPrecision(enum.IntEnum):
    bits16 = 16
    bits32 = 32

    @property
    def numPasses(self) -> int:
        return 2 if self is RadixSortKeyPrecision.bits16 else 4


class RendererError(Exception):
    """RendererError (GaussianRendererProtocol.swift:274-324)."""

    CASES = {1: "deviceNotAvailable", 2: "failedToCreatePipeline", 3: "failedToAllocateBuffer",
             4: "invalidGaussianCount", 5: "invalidDimensions", 6: "invalidTileCount", 7: "renderFailed",
             8: "invalidArgument"}

    def __init__(self, status: int, detail: str = ""):
        self.status = int(status)
        self.case = self.CASES.get(self.status, "unknown")
        super().__init__(f"{self.case}: {detail}" if detail else self.case)


def _check(status: int) -> None:
    if status != 0:
        raise RendererError(status, (N.lib().gsm_last_error_string() or b"").decode())


@dataclass
class GaussianInput:
    gaussians: Any   # PackedWorldGaussian (48 B) or PackedWorldGaussianHalf (32 B) records on the device
    harmonics: Any   # float32 / float16 SH coefficients, planar per Gaussian
    gaussianCount: int
    shComponents: int


@dataclass
class CameraParams:
    viewMatrix: Any          # 4x4, m[col][row] (simd_float4x4 layout)
    projectionMatrix: Any
    position: Any            # 3 floats
    focalX: float
    focalY: float
    near: float = 0.1
    far: float = 10.0

    def to_native(self) -> N.gsm_camera:
        c = N.gsm_camera()
        c.viewMatrix[:] = np.asarray(self.viewMatrix, np.float32).reshape(16).tolist()
        c.projectionMatrix[:] = np.asarray(self.projectionMatrix, np.float32).reshape(16).tolist()
        c.position[:] = np.asarray(self.position, np.float32).reshape(3).tolist()
        c.focalX, c.focalY = float(self.focalX), float(self.focalY)
        c.nearPlane, c.farPlane = float(self.near), float(self.far)
        return c


@dataclass
class StereoCameraParams:
    leftEye: CameraParams
    rightEye: CameraParams


@dataclass
class RendererConfig:
    maxGaussians: int = 6_000_000
    maxWidth: int = 1920
    maxHeight: int = 1080
    precision: RenderPrecision = RenderPrecision.float16
    colorFormat: str = "bgra8Unorm_srgb"      # ignored by DepthFirst
    gaussianColorSpace: GaussianColorSpace = GaussianColorSpace.srgb
    backToFront: bool = False                 # ignored by DepthFirst


class PixelFormat(enum.IntEnum):
    """MTLPixelFormat values a drawable's colour texture can have here (gsm_pixel_format)."""
    rgba16Float = 0
    bgra8Unorm = 1
    bgra8Unorm_srgb = 2
    rgba8Unorm = 3
    rgba8Unorm_srgb = 4

    @property
    def bytesPerPixel(self) -> int:
        return 8 if self == PixelFormat.rgba16Float else 4


@dataclass
class Viewport:
    """MTLViewport (znear / zfar do not matter to the copy)."""
    originX: float
    originY: float
    width: float
    height: float

    def to_native(self) -> N.gsm_viewport:
        return N.gsm_viewport(float(self.originX), float(self.originY), float(self.width), float(self.height))


@dataclass
class EyeView:
    """EyeView (GRP.swift:68-97)."""
    viewport: Viewport
    viewMatrix: Any
    projectionMatrix: Any
    cameraPosition: Any
    focalX: float
    focalY: float
    near: float = 0.1
    far: float = 10.0

    def to_native(self) -> N.gsm_eye_view:
        cam = CameraParams(self.viewMatrix, self.projectionMatrix, self.cameraPosition, self.focalX, self.focalY, self.near, self.far)
        return N.gsm_eye_view(self.viewport.to_native(), cam.to_native())


@dataclass
class StereoConfiguration:
    """StereoConfiguration (GRP.swift:100-117); sceneTransform is a simd_float4x4 (m[col][row]), identity by default."""
    leftEye: EyeView
    rightEye: EyeView
    sceneTransform: Any = None

    def to_native(self) -> N.gsm_stereo_configuration:
        c = N.gsm_stereo_configuration()
        c.leftEye, c.rightEye = self.leftEye.to_native(), self.rightEye.to_native()
        m = np.eye(4, dtype=np.float32) if self.sceneTransform is None else np.asarray(self.sceneTransform, np.float32)
        c.sceneTransform[:] = m.reshape(16).tolist()
        return c


@dataclass
class RasterizationRateMap:
    """What MTLRasterizationRateMap.mapPhysicalToScreenCoordinates returns, tabulated per layer: (screenX float32[physical
    width], screenY float32[physical height]) = screen coordinates of the physical column / row centres. 1 or 2 layers."""
    layers: Any

    def to_native(self):
        m = N.gsm_rate_map()
        keep = []
        m.layerCount = len(self.layers)
        for i, (sx, sy) in enumerate(self.layers):
            sx, sy = np.ascontiguousarray(sx, np.float32), np.ascontiguousarray(sy, np.float32)
            keep += [sx, sy]
            m.layers[i] = N.gsm_rate_map_layer(sx.size, sy.size, sx.ctypes.data, sy.ctypes.data)
        return m, keep


@dataclass
class FoveatedStereoDrawable:
    """FoveatedStereoDrawable (GRP.swift:168-193). colorTexture: a device tensor of shape (arrayLength, textureHeight, rowBytes)
    bytes or anything laid out that way; arrayLength 2 = layered, 1 = shared. The depth texture is not written on this path."""
    colorTexture: Any
    textureWidth: int
    textureHeight: int
    arrayLength: int = 2
    rasterizationRateMap: Any = None
    colorPixelFormat: PixelFormat = PixelFormat.bgra8Unorm_srgb
    rowBytes: int = 0
    sliceBytes: int = 0
    depthTexture: Any = None

    def to_native(self):
        d = N.gsm_foveated_drawable()
        d.colorTexture = N.ptr(self.colorTexture)
        d.textureWidth, d.textureHeight, d.arrayLength = int(self.textureWidth), int(self.textureHeight), int(self.arrayLength)
        d.rowBytes = int(self.rowBytes) or int(self.textureWidth) * PixelFormat(self.colorPixelFormat).bytesPerPixel
        d.sliceBytes = int(self.sliceBytes) or d.rowBytes * int(self.textureHeight)
        d.colorPixelFormat = int(self.colorPixelFormat)
        keep = None
        if self.rasterizationRateMap is not None:
            m, tables = self.rasterizationRateMap.to_native()
            keep = (m, tables)
            d.rasterizationRateMap = C.pointer(m)
        return d, keep


@dataclass
class StereoRenderTarget:
    """StereoRenderTarget (GRP.swift:233-239): .sideBySide(colorTexture:, depthTexture:) or .foveated(drawable:, configuration:)."""
    colorTexture: Any = None
    depthTexture: Any = None
    kind: str = "sideBySide"
    drawable: Any = None
    configuration: Any = None

    @staticmethod
    def sideBySide(colorTexture, depthTexture=None) -> "StereoRenderTarget":
        return StereoRenderTarget(colorTexture, depthTexture, "sideBySide")

    @staticmethod
    def foveated(drawable: FoveatedStereoDrawable, configuration: StereoConfiguration) -> "StereoRenderTarget":
        return StereoRenderTarget(None, None, "foveated", drawable, configuration)


# debugRead* ids (include/gsm/gsm.h gsm_debug_buffer)
_DBG = dict(header=0, activeTileCount=1, sortedTileIds=2, tileBounds=3, sortedPrimitiveIndices=4,
            instanceOffsets=5, nTouchedTiles=6, instanceGaussianIndices=7, depthKeys=8, renderData=9,
            tileHeaders=10, activeTiles=11, scratchDepthKeys=12, scratchPrimitiveIndices=13, depthSortPlan=14)

RENDER_DATA_DTYPE = np.dtype(
    [("meanX", "<f2"), ("meanY", "<f2"), ("theta", "<u2"), ("sigma1", "<f2"), ("sigma2", "<f2"),
     ("depth", "<f2"), ("colorR", "u1"), ("colorG", "u1"), ("colorB", "u1"), ("opacity", "u1")])
STEREO_RENDER_DATA_DTYPE = np.dtype(
    [("leftMeanX", "<f2"), ("leftMeanY", "<f2"), ("leftCxx", "<f2"), ("leftCyy", "<f2"),
     ("leftCxy2", "<f2"), ("leftDepth", "<f2"),
     ("rightMeanX", "<f2"), ("rightMeanY", "<f2"), ("rightCxx", "<f2"), ("rightCyy", "<f2"),
     ("rightCxy2", "<f2"), ("rightDepth", "<f2"),
     ("colorR", "u1"), ("colorG", "u1"), ("colorB", "u1"), ("opacity", "u1"),
     ("centerDepth", "<f2"), ("_pad0", "<u2")])


class DepthFirstRenderer:
    """DepthFirstRenderer (DepthFirstRenderer.swift:6). One host thread per instance."""

    maxSupportedGaussians = 30_000_000
    tileWidth = 16
    tileHeight = 16

    def __init__(self, device: Optional[int] = None, config: RendererConfig = None,
                 depthSortKeyPrecision: RadixSortKeyPrecision = RadixSortKeyPrecision.bits32,
                 tileIdPrecision: RadixSortKeyPrecision = RadixSortKeyPrecision.bits16,
                 stereoCopyFlipY: bool = True):
        config = config or RendererConfig()
        self.config = config
        self.lastGPUTime: Optional[float] = None
        self._lib = N.lib()
        cfg = N.gsm_config()
        self._lib.gsm_config_default(C.byref(cfg))
        cfg.maxGaussians, cfg.maxWidth, cfg.maxHeight = config.maxGaussians, config.maxWidth, config.maxHeight
        cfg.precision = config.precision.value
        cfg.gaussianColorSpace = int(config.gaussianColorSpace)
        cfg.depthSortKeyPrecision = int(depthSortKeyPrecision)
        cfg.tileIdPrecision = int(tileIdPrecision)
        cfg.device = -1 if device is None else int(device)
        cfg.stereoCopyFlipY = 1 if stereoCopyFlipY else 0
        self._cfg = cfg
        self._tile_dtype = np.uint16 if tileIdPrecision == RadixSortKeyPrecision.bits16 else np.uint32
        h = C.c_void_p()
        _check(self._lib.gsm_renderer_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._stereo_last = False

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.gsm_renderer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- GaussianRenderer protocol
    def render(self, commandBuffer, colorTexture, depthTexture, input: GaussianInput, camera: CameraParams,
               width: int, height: int) -> None:
        cam = camera.to_native()
        _check(self._lib.gsm_render(self._h, N.stream_handle(commandBuffer), N.ptr(colorTexture),
                                    N.ptr(depthTexture), N.ptr(input.gaussians), N.ptr(input.harmonics),
                                    int(input.gaussianCount), int(input.shComponents), C.byref(cam),
                                    int(width), int(height)))
        self._stereo_last = False

    def renderStereo(self, commandBuffer, target: StereoRenderTarget, input: GaussianInput,
                     camera: StereoCameraParams, width: int, height: int, eyeMask: int = 3) -> None:
        """eyeMask (bit 0 left, bit 1 right) is the one-eye-per-GPU extension; 3 is the reference behaviour."""
        if target.kind == "foveated":  # DFR.swift:225-233: the cameras come from the configuration, `camera` is not read
            d, keep = target.drawable.to_native()
            cfg = target.configuration.to_native()
            _check(self._lib.gsm_render_stereo_foveated(self._h, N.stream_handle(commandBuffer), C.byref(d), N.ptr(input.gaussians),
                                                        N.ptr(input.harmonics), int(input.gaussianCount), int(input.shComponents),
                                                        C.byref(cfg), int(width), int(height)))
            del keep
            self._stereo_last = True
            return
        l, r = camera.leftEye.to_native(), camera.rightEye.to_native()
        _check(self._lib.gsm_render_stereo_eyes(self._h, N.stream_handle(commandBuffer), N.ptr(target.colorTexture),
                                                N.ptr(input.gaussians), N.ptr(input.harmonics),
                                                int(input.gaussianCount), int(input.shComponents), C.byref(l),
                                                C.byref(r), int(width), int(height), int(eyeMask)))
        self._stereo_last = True

    def stereoCopy(self, commandBuffer, intermediate, width: int, height: int, drawable: FoveatedStereoDrawable,
                   leftViewport: Viewport, rightViewport: Viewport) -> None:
        """gsm_stereo_copy: step 10 alone (DepthFirstStereoCopyEncoder.encodeRender) from a (2*width) x height rgba16f image."""
        d, keep = drawable.to_native()
        lv, rv = leftViewport.to_native(), rightViewport.to_native()
        _check(self._lib.gsm_stereo_copy(self._h, N.stream_handle(commandBuffer), N.ptr(intermediate), int(width), int(height),
                                         C.byref(d), C.byref(lv), C.byref(rv)))
        del keep

    # -- strip-sharded single frame (multi-GPU helper; gsm_strip_project / gsm_strip_render)
    def stripProject(self, commandBuffer, gaussiansShard, harmonicsShard, gidFirst: int, gidCount: int, shComponents: int,
                     camera: CameraParams, width: int, height: int, recordsOut) -> int:
        cam = camera.to_native()
        n = C.c_uint32(0)
        _check(self._lib.gsm_strip_project(self._h, N.stream_handle(commandBuffer), N.ptr(gaussiansShard), N.ptr(harmonicsShard),
                                           int(gidFirst), int(gidCount), int(shComponents), C.byref(cam), int(width),
                                           int(height), N.ptr(recordsOut), C.byref(n)))
        return int(n.value)

    def stripRender(self, commandBuffer, colorTexture, depthTexture, records, recordCount: int, width: int, height: int,
                    tileRowFirst: int, tileRowCount: int) -> None:
        _check(self._lib.gsm_strip_render(self._h, N.stream_handle(commandBuffer), N.ptr(colorTexture), N.ptr(depthTexture),
                                          N.ptr(records), int(recordCount), int(width), int(height), int(tileRowFirst),
                                          int(tileRowCount)))
        self._stereo_last = False

    def renderHost(self, hostGaussians, hostHarmonics, gaussianCount, shComponents, camera: CameraParams,
                   width, height, hostColor, hostDepth=None) -> None:
        """gsm_render_host: the same frame with host buffers (H2D + render + D2H + sync)."""
        cam = camera.to_native()
        _check(self._lib.gsm_render_host(self._h, N.ptr(hostGaussians), N.ptr(hostHarmonics), int(gaussianCount),
                                         int(shComponents), C.byref(cam), int(width), int(height),
                                         N.ptr(hostColor), N.ptr(hostDepth)))
        self._stereo_last = False

    def renderHostAsync(self, hostGaussians, hostHarmonics, gaussianCount, shComponents, camera: CameraParams,
                        width, height, hostColor, hostDepth=None) -> None:
        """gsm_render_host_async: enqueue H2D + frame + D2H and return (the reference's render() only encodes,
        DFR.swift:196-330). The host buffers must stay alive until waitHost()."""
        cam = camera.to_native()
        _check(self._lib.gsm_render_host_async(self._h, N.ptr(hostGaussians), N.ptr(hostHarmonics), int(gaussianCount),
                                               int(shComponents), C.byref(cam), int(width), int(height),
                                               N.ptr(hostColor), N.ptr(hostDepth)))
        self._stereo_last = False

    def waitHost(self) -> None:
        """gsm_render_host_wait: block until the frame enqueued by renderHostAsync is in the host buffers."""
        _check(self._lib.gsm_render_host_wait(self._h))

    # -- profiling
    def setProfiling(self, enabled: bool) -> None:
        _check(self._lib.gsm_set_profiling(self._h, int(enabled)))

    def stageTimesMs(self) -> dict:
        ms = (C.c_float * N.GSM_NUM_STAGES)()
        _check(self._lib.gsm_get_stage_times_ms(self._h, ms))
        out = {self._lib.gsm_stage_name(i).decode(): float(ms[i]) for i in range(N.GSM_NUM_STAGES)}
        self.lastGPUTime = sum(out.values()) * 1e-3
        return out

    # -- standalone sort (what the reference's sort unit tests drive)
    def sortPairs(self, commandBuffer, keys, payload, count: int, keyBits: int = 32, numPasses: int = 4) -> None:
        _check(self._lib.gsm_sort_pairs(self._h, N.stream_handle(commandBuffer), N.ptr(keys), N.ptr(payload),
                                        int(count), int(keyBits), int(numPasses)))

    def sortPairsScratchBytes(self, count: int, keyBits: int = 32, numPasses: int = 4) -> int:
        return int(self._lib.gsm_sort_pairs_scratch_bytes(int(count), int(keyBits), int(numPasses)))

    def sortPairsWithScratch(self, commandBuffer, keys, payload, count: int, keyBits: int, numPasses: int, scratch) -> None:
        """gsm_sort_pairs_with_scratch: no allocation, no host synchronisation."""
        _check(self._lib.gsm_sort_pairs_with_scratch(self._h, N.stream_handle(commandBuffer), N.ptr(keys), N.ptr(payload),
                                                     int(count), int(keyBits), int(numPasses), N.ptr(scratch)))

    # -- white-box reads (DepthFirstUnitTests.swift:911-1252)
    def _read(self, which: str, dtype, count: int, first: int = 0, stream=None) -> np.ndarray:
        dtype = np.dtype(dtype)
        out = np.empty(count, dtype)
        if count:
            _check(self._lib.gsm_debug_read(self._h, N.stream_handle(stream), _DBG[which], N.ptr(out), first, count))
        return out

    def debugReadHeader(self) -> N.DepthFirstHeader:
        h = N.DepthFirstHeader()
        _check(self._lib.gsm_debug_read(self._h, None, _DBG["header"], C.addressof(h), 0, 1))
        return h

    def debugReadDepthSortPlan(self) -> dict:
        """No reference counterpart: the bucket plan of the last frame's depth sort (bucketCount 0 = nothing was planned)."""
        w = self._read("depthSortPlan", np.dtype((np.uint32, 4)), 1)[0]
        return dict(bucketCount=int(w[0]), keyMin=int(w[1]), fineShift=int(w[2]))

    def debugReadActiveTileCount(self) -> int:
        return int(self._read("activeTileCount", np.uint32, 1)[0])

    def debugReadSortedTileIds(self, count: int) -> np.ndarray:
        return self._read("sortedTileIds", self._tile_dtype, count).astype(np.uint32)

    def debugReadTileBounds(self, count: int) -> np.ndarray:
        return self._read("tileBounds", np.dtype((np.int32, 4)), count)

    def debugReadSingleBounds(self, at: int) -> np.ndarray:
        return self._read("tileBounds", np.dtype((np.int32, 4)), 1, first=at)[0]

    def debugReadSortedPrimitiveIndices(self, count: int) -> np.ndarray:
        return self._read("sortedPrimitiveIndices", np.int32, count)

    def debugReadSortedPrimitiveIndicesRange(self, start: int, count: int) -> np.ndarray:
        return self._read("sortedPrimitiveIndices", np.int32, count, first=start)

    def debugReadInstanceOffsets(self, count: int) -> np.ndarray:
        return self._read("instanceOffsets", np.uint32, count)

    debugReadOrderedTileCounts = debugReadInstanceOffsets  # same buffer after the in-place scan

    def debugReadNTouchedTiles(self, count: int) -> np.ndarray:
        return self._read("nTouchedTiles", np.uint32, count)

    def debugReadInstanceGaussianIndices(self, count: int) -> np.ndarray:
        return self._read("instanceGaussianIndices", np.int32, count)

    def debugReadDepthKeys(self, count: int) -> np.ndarray:
        return self._read("depthKeys", np.uint32, count)

    def debugReadScratchDepthKeys(self, count: int) -> np.ndarray:
        return self._read("scratchDepthKeys", np.uint32, count)

    def debugReadScratchPrimitiveIndices(self, count: int) -> np.ndarray:
        return self._read("scratchPrimitiveIndices", np.int32, count)

    def debugReadRenderData(self, count: int) -> np.ndarray:
        return self._read("renderData", STEREO_RENDER_DATA_DTYPE if self._stereo_last else RENDER_DATA_DTYPE, count)

    def debugReadTileHeaders(self, count: int) -> np.ndarray:
        return self._read("tileHeaders", np.dtype((np.uint32, 2)), count)

    def debugReadActiveTiles(self, count: int) -> np.ndarray:
        return self._read("activeTiles", np.uint32, count)


def probe_math(op: int, a, b=None, device: int = -1) -> np.ndarray:
    """gsm_probe_math: device restatement of the canonical math (0 sin, 1 cos, 2 log, 3 atan2, 4 powr 2.4,
    5 half exp, 6 float->half, 7/8 packed half exp forms, 9/10 min/max, 11 fused half fma on (n,3) triples,
    12 exp(-0.5h * p) with the -0.5 folded in, 13 the same through the blend's shared-memory table, 14 the same on the XU pipe
    (MUFU.EX2 behind the rounding guard), 15 unguarded, 16 the guard's flag, 17 the tuned guard-free form the blend kernels run)."""
    a = np.ascontiguousarray(a)
    n = a.shape[0] if op == 11 else a.size  # op 11 (half fma) takes (n, 3) uint16 triples
    out = np.empty(n if op == 11 else a.shape, np.uint16 if op in (5, 6, 7, 8, 11, 12, 13, 14, 15, 16, 17) else np.float32)
    bp = None if b is None else N.ptr(np.ascontiguousarray(b))
    _check(N.lib().gsm_probe_math(device, op, N.ptr(a), bp, N.ptr(out), n))
    return out


class GlobalRenderer:
    """GlobalRenderer (GlobalRenderer.swift:72): the reference's second renderer behind the same GaussianRenderer protocol --
    32 x 16 tiles of the configured limits, one sort of 32-bit [tile:16][half depth:16] keys, 4 x 2 pixels per thread."""

    maxSupportedGaussians = 30_000_000
    tileWidth = 32
    tileHeight = 16
    _GDBG = dict(header=0, sortedKeys=1, sortedIndices=2, tileHeaders=3, bounds=4, renderData=5, visibleIndices=6, activeTiles=7)
    HEADER_DTYPE = np.dtype([("totalAssignments", "<u4"), ("paddedCount", "<u4"), ("overflow", "<u4"), ("visibleCount", "<u4"),
                             ("activeTileCount", "<u4"), ("totalRaw", "<u4"), ("_pad", "<u4", 2)])

    def __init__(self, device: Optional[int] = None, config: RendererConfig = None):
        self._r = DepthFirstRenderer(device=device, config=config)   # the handle carries limits, precision, colour space
        self.config = self._r.config
        self.lastGPUTime: Optional[float] = None
        self._lib, self._h = self._r._lib, self._r._h

    def close(self) -> None:
        self._r.close()

    def render(self, commandBuffer, colorTexture, depthTexture, input: GaussianInput, camera: CameraParams,
               width: int, height: int) -> None:
        cam = camera.to_native()
        _check(self._lib.gsm_render_global(self._r._h, N.stream_handle(commandBuffer), N.ptr(colorTexture), N.ptr(depthTexture),
                                           N.ptr(input.gaussians), N.ptr(input.harmonics), int(input.gaussianCount),
                                           int(input.shComponents), C.byref(cam), int(width), int(height)))

    def renderStereo(self, commandBuffer, target, input, camera, width: int, height: int) -> None:
        # GlobalRenderer.swift:249-265: fatalError in the reference
        raise NotImplementedError("GlobalRenderer does not support stereo rendering. Use DepthFirstRenderer instead.")

    def _read(self, which: str, dtype, count: int, first: int = 0) -> np.ndarray:
        out = np.empty(count, np.dtype(dtype))
        if count:
            _check(self._lib.gsm_global_debug_read(self._r._h, None, self._GDBG[which], N.ptr(out), first, count))
        return out

    def debugReadHeader(self) -> np.void:
        return self._read("header", self.HEADER_DTYPE, 1)[0]

    def debugReadTotalAssignments(self) -> int:   # GlobalRenderer.swift:200-203
        return int(self.debugReadHeader()["totalAssignments"])

    def debugReadSortedKeys(self, count: int) -> np.ndarray:
        return self._read("sortedKeys", np.uint32, count)

    def debugReadSortedIndices(self, count: int) -> np.ndarray:
        return self._read("sortedIndices", np.int32, count)

    def debugReadTileHeaders(self, count: int) -> np.ndarray:
        return self._read("tileHeaders", np.dtype((np.uint32, 2)), count)

    def debugReadBounds(self, count: int) -> np.ndarray:
        return self._read("bounds", np.dtype((np.int32, 4)), count)

    def debugReadRenderData(self, count: int) -> np.ndarray:
        return self._read("renderData", RENDER_DATA_DTYPE, count)

    def debugReadVisibleIndices(self, count: int) -> np.ndarray:
        return self._read("visibleIndices", np.uint32, count)

    def debugReadActiveTiles(self, count: int) -> np.ndarray:
        return self._read("activeTiles", np.uint32, count)
