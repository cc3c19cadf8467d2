"""ctypes binding of the CPU oracle (oracle/libgsm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (gsm_renderer_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgsm_oracle.so")

F32, F16 = 0, 1

RENDER_DATA_DTYPE = np.dtype(
    [("meanX", "<f2"), ("meanY", "<f2"), ("theta", "<u2"), ("sigma1", "<f2"), ("sigma2", "<f2"),
     ("depth", "<f2"), ("colorR", "u1"), ("colorG", "u1"), ("colorB", "u1"), ("opacity", "u1")]
)
STEREO_RENDER_DATA_DTYPE = np.dtype(
    [("leftMeanX", "<f2"), ("leftMeanY", "<f2"), ("leftCxx", "<f2"), ("leftCyy", "<f2"),
     ("leftCxy2", "<f2"), ("leftDepth", "<f2"),
     ("rightMeanX", "<f2"), ("rightMeanY", "<f2"), ("rightCxx", "<f2"), ("rightCyy", "<f2"),
     ("rightCxy2", "<f2"), ("rightDepth", "<f2"),
     ("colorR", "u1"), ("colorG", "u1"), ("colorB", "u1"), ("opacity", "u1"),
     ("centerDepth", "<f2"), ("_pad0", "<u2")]
)
assert RENDER_DATA_DTYPE.itemsize == 16 and STEREO_RENDER_DATA_DTYPE.itemsize == 32


class Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("center", C.c_float * 3),
                ("width", C.c_float), ("height", C.c_float), ("nearPlane", C.c_float),
                ("farPlane", C.c_float), ("shComponents", C.c_uint32), ("gaussianCount", C.c_uint32),
                ("inputIsSRGB", C.c_float)]


class StereoCamera(C.Structure):
    _fields_ = [("leftView", C.c_float * 16), ("leftProj", C.c_float * 16), ("leftCenter", C.c_float * 3),
                ("rightView", C.c_float * 16), ("rightProj", C.c_float * 16), ("rightCenter", C.c_float * 3),
                ("width", C.c_float), ("height", C.c_float), ("nearPlane", C.c_float),
                ("farPlane", C.c_float), ("shComponents", C.c_uint32), ("gaussianCount", C.c_uint32),
                ("inputIsSRGB", C.c_float), ("sceneTransform", C.c_float * 16)]


class Binning(C.Structure):
    _fields_ = [("tilesX", C.c_uint32), ("tilesY", C.c_uint32), ("tileWidth", C.c_uint32),
                ("tileHeight", C.c_uint32), ("alphaThreshold", C.c_float),
                ("totalInkThreshold", C.c_float)]


class DFHeader(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("visibleCount", "totalInstances", "paddedVisibleCount",
                                          "paddedInstanceCount", "overflow", "padding0", "padding1",
                                          "padding2")]


class Frame(C.Structure):
    _fields_ = [("maxGaussians", C.c_uint32), ("maxInstances", C.c_uint32),
                ("depthKey16", C.c_int), ("tileId16", C.c_int),
                ("renderData", C.c_void_p), ("bounds", C.c_void_p), ("nTouched", C.c_void_p),
                ("preDepthKeys", C.c_void_p), ("depthKeys", C.c_void_p),
                ("primitiveIndices", C.c_void_p), ("orderedTileCounts", C.c_void_p),
                ("instanceTileIds", C.c_void_p), ("instanceGaussianIndices", C.c_void_p),
                ("tileHeaders", C.c_void_p), ("activeTiles", C.c_void_p),
                ("header", DFHeader), ("activeTileCount", C.c_uint32),
                ("rawVisibleCount", C.c_uint32), ("rawTotalInstances", C.c_uint32),
                ("stageSeconds", C.c_double * 10)]


STAGE_NAMES = ("project", "compact", "depthSort", "applyScan", "expand", "tileSort", "ranges",
               "clearBlend", "copy", "total")


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc). Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("gsm_oracle.c", "gsm_oracle_ply.c", "gsm_oracle_copy.c", "gsm_oracle.h", "gsmo_math.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.gsmo_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def binning_for(width: int, height: int) -> Binning:
    return Binning((width + 15) // 16, (height + 15) // 16, 16, 16, 0.005, 2.0)


def make_camera(view, proj, center, width, height, near, far, sh_components, count, srgb) -> Camera:
    cam = Camera()
    cam.view[:] = np.asarray(view, np.float32).reshape(16).tolist()
    cam.proj[:] = np.asarray(proj, np.float32).reshape(16).tolist()
    cam.center[:] = np.asarray(center, np.float32).reshape(3).tolist()
    cam.width, cam.height = float(width), float(height)
    cam.nearPlane, cam.farPlane = float(near), float(far)
    cam.shComponents, cam.gaussianCount = int(sh_components), int(count)
    cam.inputIsSRGB = 1.0 if srgb else 0.0
    return cam


def make_stereo_camera(lview, lproj, lcenter, rview, rproj, rcenter, width, height, near, far,
                       sh_components, count, srgb, scene=None) -> StereoCamera:
    cam = StereoCamera()
    cam.leftView[:] = np.asarray(lview, np.float32).reshape(16).tolist()
    cam.leftProj[:] = np.asarray(lproj, np.float32).reshape(16).tolist()
    cam.leftCenter[:] = np.asarray(lcenter, np.float32).reshape(3).tolist()
    cam.rightView[:] = np.asarray(rview, np.float32).reshape(16).tolist()
    cam.rightProj[:] = np.asarray(rproj, np.float32).reshape(16).tolist()
    cam.rightCenter[:] = np.asarray(rcenter, np.float32).reshape(3).tolist()
    cam.width, cam.height = float(width), float(height)
    cam.nearPlane, cam.farPlane = float(near), float(far)
    cam.shComponents, cam.gaussianCount = int(sh_components), int(count)
    cam.inputIsSRGB = 1.0 if srgb else 0.0
    scene = np.eye(4, dtype=np.float32) if scene is None else np.asarray(scene, np.float32)
    cam.sceneTransform[:] = scene.reshape(16).tolist()
    return cam


# ------------------------------------------------------------------ math probes
def probe_sincos(x):
    x = np.ascontiguousarray(x, np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().gsmo_probe_sincos(_p(x), _p(s), _p(c), C.c_int(x.size))
    return s, c


def probe_log(x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().gsmo_probe_log(_p(x), _p(y), C.c_int(x.size))
    return y


def probe_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    r = np.empty_like(x)
    lib().gsmo_probe_atan2(_p(y), _p(x), _p(r), C.c_int(x.size))
    return r


def probe_powr(x, yexp):
    x = np.ascontiguousarray(x, np.float32)
    r = np.empty_like(x)
    lib().gsmo_probe_powr(_p(x), C.c_float(yexp), _p(r), C.c_int(x.size))
    return r


def probe_hexp(xbits):
    x = np.ascontiguousarray(xbits, np.uint16)
    y = np.empty_like(x)
    lib().gsmo_probe_hexp(_p(x), _p(y), C.c_int(x.size))
    return y


def probe_hfma(a, b, c):
    a, b, c = (np.ascontiguousarray(x, np.uint16) for x in (a, b, c))
    r = np.empty_like(a)
    lib().gsmo_probe_hfma(_p(a), _p(b), _p(c), _p(r), C.c_int(a.size))
    return r


def probe_minmax(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    mn, mx = np.empty_like(a), np.empty_like(a)
    lib().gsmo_probe_minmax(_p(a), _p(b), _p(mn), _p(mx), C.c_int(a.size))
    return mn, mx


def probe_f2h(x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty(x.shape, np.uint16)
    lib().gsmo_probe_f2h(_p(x), _p(y), C.c_int(x.size))
    return y


# ------------------------------------------------------------------ stages
def sort_pairs_u32(keys, payload, passes=4):
    keys = np.ascontiguousarray(keys, np.uint32).copy()
    payload = np.ascontiguousarray(payload, np.int32).copy()
    lib().gsmo_sort_pairs_u32(_p(keys), _p(payload), C.c_uint32(keys.size), C.c_int(passes))
    return keys, payload


def sort_pairs_u16(keys, payload, passes=2):
    keys = np.ascontiguousarray(keys, np.uint16).copy()
    payload = np.ascontiguousarray(payload, np.int32).copy()
    lib().gsmo_sort_pairs_u16(_p(keys), _p(payload), C.c_uint32(keys.size), C.c_int(passes))
    return keys, payload


def exclusive_scan(x):
    x = np.ascontiguousarray(x, np.uint32)
    y = np.empty_like(x)
    lib().gsmo_exclusive_scan(_p(x), C.c_uint32(x.size), _p(y))
    return y


def tile_sort_passes(tile_count: int) -> int:
    return int(lib().gsmo_tile_sort_passes(C.c_uint32(tile_count)))


class OracleFrame:
    """Runs a whole frame through the oracle and keeps every intermediate as numpy arrays."""

    def __init__(self, max_gaussians: int, max_width: int, max_height: int, stereo: bool = False,
                 depth_key16: bool = False, tile_id16: bool = True):
        G = int(max_gaussians)
        self.G, self.I = G, 4 * G
        self.stereo = stereo
        T = ((max_width + 15) // 16) * ((max_height + 15) // 16)
        self.renderData = np.zeros(G, STEREO_RENDER_DATA_DTYPE if stereo else RENDER_DATA_DTYPE)
        self.bounds = np.zeros((G, 4), np.int32)
        self.nTouched = np.zeros(G, np.uint32)
        self.preDepthKeys = np.zeros(G, np.uint32)
        self.depthKeys = np.zeros(G, np.uint32)
        self.primitiveIndices = np.zeros(G, np.int32)
        self.orderedTileCounts = np.zeros(G, np.uint32)
        self.instanceTileIds = np.zeros(self.I, np.uint16 if tile_id16 else np.uint32)
        self.instanceGaussianIndices = np.zeros(self.I, np.int32)
        self.tileHeaders = np.zeros((max(T, 1), 2), np.uint32)
        self.activeTiles = np.zeros(max(T, 1), np.uint32)
        f = Frame()
        f.maxGaussians, f.maxInstances = G, 4 * G
        f.depthKey16, f.tileId16 = int(depth_key16), int(tile_id16)
        for name in ("renderData", "bounds", "nTouched", "preDepthKeys", "depthKeys", "primitiveIndices",
                     "orderedTileCounts", "instanceTileIds", "instanceGaussianIndices", "tileHeaders",
                     "activeTiles"):
            setattr(f, name, getattr(self, name).ctypes.data)
        self.f = f

    @property
    def header(self):
        return self.f.header

    @property
    def stage_seconds(self):
        return dict(zip(STAGE_NAMES, list(self.f.stageSeconds)))

    def render_mono(self, gaussians, harmonics, precision, cam: Camera, width, height, want_depth=True):
        color = np.full((height, width, 4), 0x7E00, np.uint16)  # NaN pattern: "untouched"
        depth = np.full((height, width), 0x7E00, np.uint16) if want_depth else None
        lib().gsmo_render_mono(C.byref(self.f), _p(gaussians), _p(harmonics), C.c_int(precision),
                               C.byref(cam), C.c_uint32(width), C.c_uint32(height), _p(color), _p(depth))
        return color, depth

    def render_stereo(self, gaussians, harmonics, precision, cam: StereoCamera, width, height, flip_y=True):
        scratch = np.zeros((2, height, width, 4), np.uint16)
        dst = np.full((height, 2 * width, 4), 0x7E00, np.uint16)
        lib().gsmo_render_stereo(C.byref(self.f), _p(gaussians), _p(harmonics), C.c_int(precision),
                                 C.byref(cam), C.c_uint32(width), C.c_uint32(height), C.c_int(int(flip_y)),
                                 _p(scratch), _p(dst))
        return dst, scratch


# ---------------------------------------------------------------- foveated stereo copy (gsm_oracle_copy.c)
PIXEL_BYTES = {0: 8, 1: 4, 2: 4, 3: 4, 4: 4}


def stereo_copy_foveated(color2, flip_y, tex_w, tex_h, array_length, fmt, viewports, rate_layers=None, dst=None, row_bytes=None):
    """color2: (2, H, W, 4) uint16 halfs. viewports: ((ox, oy, w, h), (ox, oy, w, h)). rate_layers: None or a list of 1-2
    (screenX float32[physW], screenY float32[physH]). Returns the drawable bytes (array_length, tex_h, row_bytes) uint8;
    texels no viewport covers keep what `dst` held (0xAB if not given)."""
    color2 = np.ascontiguousarray(color2, np.uint16)
    _, H, W, _ = color2.shape
    px = PIXEL_BYTES[fmt]
    row_bytes = row_bytes or tex_w * px
    if dst is None:
        dst = np.full((array_length, tex_h, row_bytes), 0xAB, np.uint8)
    layers = rate_layers or []
    sx = [np.ascontiguousarray(l[0], np.float32) for l in layers]
    sy = [np.ascontiguousarray(l[1], np.float32) for l in layers]
    pw = (C.c_uint32 * 2)(*([a.size for a in sx] + [0, 0])[:2])
    ph = (C.c_uint32 * 2)(*([a.size for a in sy] + [0, 0])[:2])
    fx = (C.c_void_p * 2)(*([a.ctypes.data for a in sx] + [None, None])[:2])
    fy = (C.c_void_p * 2)(*([a.ctypes.data for a in sy] + [None, None])[:2])
    vp = (C.c_double * 8)(*[float(v) for e in viewports for v in e])
    f = lib().gsmo_stereo_copy_foveated
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t,
                  C.c_size_t, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(color2), W, H, int(flip_y), _p(dst), tex_w, tex_h, array_length, row_bytes, tex_h * row_bytes, fmt, len(layers),
      pw, ph, fx, fy, vp)
    return dst


# ---------------------------------------------------------------- scene ingest (gsm_oracle_ply.c)
class PlyResult(C.Structure):
    _fields_ = [("count", C.c_uint32), ("shComponents", C.c_uint32), ("harmonicsStride", C.c_uint32),
                ("compressed", C.c_uint32), ("scaleIsLogSpace", C.c_uint32), ("opacityIsLogit", C.c_uint32),
                ("center", C.c_float * 3), ("boundsCenter", C.c_float * 3), ("boundsRadius", C.c_float)]


PLY_ERRORS = {1: "invalidHeader", 2: "unsupportedFormat", 3: "missingVertexElement", 4: "missingRequiredProperties",
              5: "listPropertiesNotSupported", 6: "insufficientData", 7: "missingChunkElement", 8: "invalidHeader"}


def ply_load(data: bytes):
    """PLYLoader.load restated (oracle). Returns dict(pos, scale, rot, opacity, harmonics, result) or raises ValueError(case)."""
    l = lib()
    buf = np.frombuffer(data, dtype=np.uint8)
    n = C.c_uint32(0)
    shp = C.c_uint32(0)
    l.gsmo_ply_probe.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    rc = l.gsmo_ply_probe(buf.ctypes.data, buf.size, C.byref(n), C.byref(shp))
    if rc:
        raise ValueError(PLY_ERRORS.get(rc, str(rc)))
    cnt = max(int(n.value), 1)
    pos = np.zeros((cnt, 3), np.float32); scale = np.zeros((cnt, 3), np.float32)
    rot = np.zeros((cnt, 4), np.float32); op = np.zeros(cnt, np.float32)
    har = np.zeros(cnt * max(int(shp.value), 3), np.float32)
    res = PlyResult()
    l.gsmo_ply_load.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 5 + [C.c_size_t, C.POINTER(PlyResult)]
    rc = l.gsmo_ply_load(buf.ctypes.data, buf.size, _p(pos), _p(scale), _p(rot), _p(op), _p(har), har.size, C.byref(res))
    if rc:
        raise ValueError(PLY_ERRORS.get(rc, str(rc)))
    m, hs = int(res.count), int(res.harmonicsStride)
    return {"pos": pos[:m], "scale": scale[:m], "rot": rot[:m], "opacity": op[:m],
            "harmonics": har[:m * hs].reshape(m, hs) if hs else np.zeros((m, 0), np.float32), "result": res}


def pack_gaussians(rec, half: bool, order=None) -> np.ndarray:
    l = lib()
    m = len(rec["pos"])
    out = np.zeros((m, 32 if half else 48), np.uint8)
    l.gsmo_pack_gaussians.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
    o = None if order is None else np.ascontiguousarray(order, np.uint32)
    l.gsmo_pack_gaussians(_p(np.ascontiguousarray(rec["pos"])), _p(np.ascontiguousarray(rec["scale"])),
                          _p(np.ascontiguousarray(rec["rot"])), _p(np.ascontiguousarray(rec["opacity"])), m, _p(o), int(half), _p(out))
    return out


def pack_harmonics(har: np.ndarray, half: bool, order=None) -> np.ndarray:
    l = lib()
    m, hs = har.shape
    out = np.zeros((m, hs), np.float16 if half else np.float32)
    if m * hs == 0:
        return out
    l.gsmo_pack_harmonics.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
    o = None if order is None else np.ascontiguousarray(order, np.uint32)
    l.gsmo_pack_harmonics(_p(np.ascontiguousarray(har)), m, hs, _p(o), int(half), _p(out))
    return out


def morton_order(pos: np.ndarray):
    l = lib()
    n = len(pos)
    codes = np.zeros(n, np.uint64)
    order = np.zeros(n, np.uint32)
    l.gsmo_morton_order.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    l.gsmo_morton_order(_p(np.ascontiguousarray(pos, np.float32)), n, _p(codes), _p(order))
    return codes, order


# ---------------------------------------------------------------- GlobalRenderer frame (gsmo_render_global)
class GlobalInfo(C.Structure):
    _fields_ = [("totalAssignments", C.c_uint32), ("paddedCount", C.c_uint32), ("overflow", C.c_uint32),
                ("visibleCount", C.c_uint32), ("activeTileCount", C.c_uint32)]


class OracleGlobalFrame:
    """One GlobalRenderer frame through the oracle, every white-box buffer kept as numpy arrays."""

    def __init__(self, max_gaussians: int, max_width: int, max_height: int):
        self.G, self.maxW, self.maxH = int(max_gaussians), int(max_width), int(max_height)
        self.tilesX, self.tilesY = (self.maxW + 31) // 32, (self.maxH + 15) // 16
        T = max(1, self.tilesX * self.tilesY)
        G, A = self.G, 4 * self.G
        self.renderData = np.zeros(G, RENDER_DATA_DTYPE)
        self.bounds = np.zeros((G, 4), np.int32)
        self.mask = np.zeros(G, np.uint8)
        self.visibleIndices = np.zeros(G, np.uint32)
        self.sortedKeys = np.zeros(A, np.uint32)
        self.sortedIndices = np.zeros(A, np.int32)
        self.tileHeaders = np.zeros((T, 2), np.uint32)
        self.info = GlobalInfo()

    def render(self, gaussians, harmonics, precision, cam: Camera, width, height, want_depth=True):
        color = np.full((height, width, 4), 0x7E00, np.uint16)
        depth = np.full((height, width), 0x7E00, np.uint16) if want_depth else None
        f = lib().gsmo_render_global
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Camera)] + [C.c_uint32] * 5 + [C.c_void_p] * 9 + [C.POINTER(GlobalInfo)]
        f(_p(gaussians), _p(harmonics), int(precision), C.byref(cam), self.maxW, self.maxH, self.G, int(width), int(height),
          _p(color), _p(depth), _p(self.renderData), _p(self.bounds), _p(self.mask), _p(self.visibleIndices), _p(self.sortedKeys),
          _p(self.sortedIndices), _p(self.tileHeaders), C.byref(self.info))
        return color, depth
