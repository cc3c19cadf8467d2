"""Host-side mirror of the reference's scene ingest for the DepthFirst path (SURVEY.md 8(f) rank 1 + Morton pre-sort):
`PLYLoader.load(url:)` (Sources/Renderer/Utils/PLYLoader.swift:246-281) and `GaussianSceneBuilder` (Scene.swift:73-190),
over the C ABI of include/gsm/gsm_scene.h. The file is mapped on the host and DECODED ON THE DEVICE; what comes back is a
`GaussianDataset` whose buffers are device tensors already packed as PackedWorldGaussian(+Half) + planar SH, i.e. a
`GaussianInput` for `DepthFirstRenderer.render`. No CPU decode path exists here.
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
from dataclasses import dataclass
from typing import Any, Tuple

import numpy as np

from . import _native as N
from .renderer import GaussianInput, RenderPrecision, RendererError, _check


class PLYLoaderError(RendererError):
    """PLYLoaderError / PLYHeader.DecodeError (PLYLoader.swift:88-110, :209-242)."""

    CASES = dict(RendererError.CASES)
    CASES.update({20: "invalidHeader", 21: "unsupportedFormat", 22: "missingVertexElement", 23: "missingRequiredProperties",
                  24: "listPropertiesNotSupported", 25: "insufficientData", 26: "missingChunkElement"})


def _check_ply(status: int) -> None:
    if status != 0:
        raise PLYLoaderError(status, (N.lib().gsm_last_error_string() or b"").decode())


@dataclass
class PLYHeaderInfo:
    vertexCount: int
    format: str
    compressed: bool
    shProperties: int
    bodyOffset: int


@dataclass
class GaussianDataset:
    """GaussianDataset (Scene.swift:141-157) with the records already packed for the renderer (device memory)."""
    gaussians: Any          # torch uint8 tensor: count x 48 B (float32) or 32 B (float16) records
    harmonics: Any          # torch float32 / float16 tensor, count x harmonicsStride, planar [R.., G.., B..] per Gaussian
    count: int
    shComponents: int
    harmonicsStride: int
    precision: RenderPrecision
    compressed: bool
    scaleIsLogSpace: bool
    opacityIsLogit: bool
    center: Tuple[float, float, float]        # subtracted by the recentering
    boundsCenter: Tuple[float, float, float]  # GaussianSceneBuilder.bounds(of: records)
    boundsRadius: float

    def input(self) -> GaussianInput:
        return GaussianInput(self.gaussians, self.harmonics, self.count, self.shComponents)


def _file_bytes(path_or_bytes):
    """(keepalive, address, size) of the whole file in host memory; paths are memory-mapped like Data(.mappedIfSafe)."""
    if isinstance(path_or_bytes, np.ndarray):  # the file already in host memory (e.g. a pinned buffer)
        a = np.ascontiguousarray(path_or_bytes).view(np.uint8).reshape(-1)
        return a, a.ctypes.data, a.size
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        a = np.frombuffer(path_or_bytes, dtype=np.uint8)
        return a, a.ctypes.data, a.size
    f = open(os.fspath(path_or_bytes), "rb")
    size = os.fstat(f.fileno()).st_size
    if size == 0:
        f.close()
        a = np.zeros(1, np.uint8)
        return a, a.ctypes.data, 0
    m = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    a = np.frombuffer(m, dtype=np.uint8)
    return (f, m, a), a.ctypes.data, size


_FORMATS = {0: "ascii", 1: "binary_little_endian", 2: "binary_big_endian"}


class PLYLoader:
    @staticmethod
    def probe(path_or_bytes) -> PLYHeaderInfo:
        keep, addr, size = _file_bytes(path_or_bytes)
        info = N.gsm_ply_info()
        _check_ply(N.lib().gsm_ply_probe(addr, size, C.byref(info)))
        del keep
        return PLYHeaderInfo(info.vertexCount, _FORMATS.get(info.format, "?"), bool(info.compressed), info.shProperties,
                             info.bodyOffset)

    @staticmethod
    def load(path_or_bytes, device: int = 0, precision: RenderPrecision = RenderPrecision.float16, stream=None) -> GaussianDataset:
        """PLYLoader.load(url:) + the packing of PLYBenchmarkTests.swift:139-149, decoded on `device`."""
        import torch
        keep, addr, size = _file_bytes(path_or_bytes)
        info = N.gsm_ply_info()
        _check_ply(N.lib().gsm_ply_probe(addr, size, C.byref(info)))
        n = int(info.vertexCount)
        half = precision == RenderPrecision.float16
        dev = torch.device("cuda", device)
        rec = 32 if half else 48
        stride = int(info.shProperties)
        g = torch.empty((max(n, 1), rec), dtype=torch.uint8, device=dev)
        h = torch.empty((max(n, 1), max(stride, 1)), dtype=torch.float16 if half else torch.float32, device=dev)
        out = N.gsm_scene_info()
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        _check_ply(N.lib().gsm_ply_load(device, N.stream_handle(s), addr, size, precision.value, g.data_ptr(), h.data_ptr(),
                                        n, n * max(stride, 1), C.byref(out)))
        del keep
        m, hs = int(out.count), int(out.harmonicsStride)
        return GaussianDataset(g[:m], h.reshape(-1)[:m * hs].reshape(m, hs) if hs else h[:0, :0], m, int(out.shComponents), hs,
                               precision, bool(out.compressed), bool(out.scaleIsLogSpace), bool(out.opacityIsLogit),
                               tuple(out.center), tuple(out.boundsCenter), float(out.boundsRadius))


class GaussianSceneBuilder:
    @staticmethod
    def sortByMortonCode(dataset: GaussianDataset, stream=None) -> None:
        """GaussianSceneBuilder.sortByMortonCode (Scene.swift:73-138), in place on the device buffers."""
        import torch
        dev = dataset.gaussians.device
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        _check(N.lib().gsm_scene_morton_sort(dev.index or 0, N.stream_handle(s), dataset.gaussians.data_ptr(),
                                             dataset.harmonics.data_ptr() if dataset.harmonicsStride else None, dataset.count,
                                             dataset.harmonicsStride, dataset.precision.value))

    @staticmethod
    def bounds(dataset: GaussianDataset):
        """GaussianSceneBuilder.bounds(of:) (Scene.swift:159-190), computed by the loader on the float records."""
        return dataset.boundsCenter, dataset.boundsRadius
