"""ctypes binding of the C ABI in include/gsm/gsm.h (gsm_renderer_b200/lib/libgsm_b200.so).

There is no fallback: if the CUDA library is missing or does not load, importing a renderer fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSM_B200_LIB points at an instrumented build of the same library (tools/: kernel timelines); never at a fallback
LIB_PATH = os.environ.get("GSM_B200_LIB") or os.path.join(_HERE, "lib", "libgsm_b200.so")

GSM_NUM_STAGES = 8


class gsm_config(C.Structure):
    _fields_ = [("maxGaussians", C.c_uint32), ("maxWidth", C.c_uint32), ("maxHeight", C.c_uint32),
                ("precision", C.c_uint32), ("gaussianColorSpace", C.c_uint32),
                ("depthSortKeyPrecision", C.c_uint32), ("tileIdPrecision", C.c_uint32),
                ("device", C.c_int32), ("stereoCopyFlipY", C.c_uint32), ("reserved", C.c_uint32 * 7)]


class gsm_camera(C.Structure):
    _fields_ = [("viewMatrix", C.c_float * 16), ("projectionMatrix", C.c_float * 16),
                ("position", C.c_float * 3), ("focalX", C.c_float), ("focalY", C.c_float),
                ("nearPlane", C.c_float), ("farPlane", C.c_float)]


class gsm_viewport(C.Structure):
    _fields_ = [("originX", C.c_double), ("originY", C.c_double), ("width", C.c_double), ("height", C.c_double)]


class gsm_eye_view(C.Structure):
    _fields_ = [("viewport", gsm_viewport), ("camera", gsm_camera)]


class gsm_stereo_configuration(C.Structure):
    _fields_ = [("leftEye", gsm_eye_view), ("rightEye", gsm_eye_view), ("sceneTransform", C.c_float * 16)]


class gsm_rate_map_layer(C.Structure):
    _fields_ = [("physicalWidth", C.c_uint32), ("physicalHeight", C.c_uint32), ("screenX", C.c_void_p), ("screenY", C.c_void_p)]


class gsm_rate_map(C.Structure):
    _fields_ = [("layerCount", C.c_uint32), ("layers", gsm_rate_map_layer * 2)]


class gsm_foveated_drawable(C.Structure):
    _fields_ = [("colorTexture", C.c_void_p), ("textureWidth", C.c_uint32), ("textureHeight", C.c_uint32),
                ("arrayLength", C.c_uint32), ("rowBytes", C.c_size_t), ("sliceBytes", C.c_size_t),
                ("colorPixelFormat", C.c_uint32), ("rasterizationRateMap", C.POINTER(gsm_rate_map))]


class gsm_ply_info(C.Structure):  # include/gsm/gsm_scene.h
    _fields_ = [("vertexCount", C.c_uint32), ("format", C.c_uint32), ("compressed", C.c_uint32),
                ("shProperties", C.c_uint32), ("bodyOffset", C.c_uint64)]


class gsm_scene_info(C.Structure):  # include/gsm/gsm_scene.h
    _fields_ = [("count", C.c_uint32), ("shComponents", C.c_uint32), ("harmonicsStride", C.c_uint32),
                ("compressed", C.c_uint32), ("scaleIsLogSpace", C.c_uint32), ("opacityIsLogit", C.c_uint32),
                ("center", C.c_float * 3), ("boundsCenter", C.c_float * 3), ("boundsRadius", C.c_float)]


class DepthFirstHeader(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("visibleCount", "totalInstances", "paddedVisibleCount",
                                          "paddedInstanceCount", "overflow", "padding0", "padding1",
                                          "padding2")]


# every symbol include/gsm/gsm.h declares (tests check the library exports each one)
EXPORTS = {
    "gsm_config_default": (None, [C.POINTER(gsm_config)]),
    "gsm_renderer_create": (C.c_int, [C.POINTER(gsm_config), C.POINTER(C.c_void_p)]),
    "gsm_renderer_destroy": (None, [C.c_void_p]),
    "gsm_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_uint32, C.c_uint32, C.POINTER(gsm_camera), C.c_uint32, C.c_uint32]),
    "gsm_render_stereo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.c_uint32, C.POINTER(gsm_camera), C.POINTER(gsm_camera),
                                    C.c_uint32, C.c_uint32]),
    "gsm_render_stereo_foveated": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(gsm_foveated_drawable), C.c_void_p, C.c_void_p,
                                             C.c_uint32, C.c_uint32, C.POINTER(gsm_stereo_configuration), C.c_uint32, C.c_uint32]),
    "gsm_stereo_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(gsm_foveated_drawable),
                                  C.POINTER(gsm_viewport), C.POINTER(gsm_viewport)]),
    "gsm_render_stereo_eyes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_uint32, C.POINTER(gsm_camera), C.POINTER(gsm_camera),
                                         C.c_uint32, C.c_uint32, C.c_uint32]),
    "gsm_render_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                  C.POINTER(gsm_camera), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gsm_render_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                        C.POINTER(gsm_camera), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gsm_render_host_wait": (C.c_int, [C.c_void_p]),
    "gsm_ply_probe": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(gsm_ply_info)]),
    "gsm_ply_load": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32,
                               C.c_size_t, C.POINTER(gsm_scene_info)]),
    "gsm_scene_morton_sort": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]),
    "gsm_last_gpu_time_ms": (C.c_double, [C.c_void_p]),
    "gsm_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "gsm_get_stage_times_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "gsm_stage_name": (C.c_char_p, [C.c_int]),
    "gsm_debug_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "gsm_render_global": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.c_uint32, C.POINTER(gsm_camera), C.c_uint32, C.c_uint32]),
    "gsm_global_debug_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "gsm_debug_element_size": (C.c_size_t, [C.c_void_p, C.c_int]),
    "gsm_buffer_alloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "gsm_buffer_free": (C.c_int, [C.c_void_p]),
    "gsm_buffer_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gsm_buffer_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gsm_stream_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "gsm_stream_synchronize": (C.c_int, [C.c_void_p]),
    "gsm_stream_destroy": (C.c_int, [C.c_void_p]),
    "gsm_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int]),
    "gsm_sort_pairs_scratch_bytes": (C.c_size_t, [C.c_uint32, C.c_int, C.c_int]),
    "gsm_sort_pairs_with_scratch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]),
    "gsm_strip_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.POINTER(gsm_camera), C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.POINTER(C.c_uint32)]),
    "gsm_strip_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "gsm_group_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p)]),
    "gsm_group_destroy": (None, [C.c_void_p]),
    "gsm_group_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gsm_group_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gsm_group_connect_local": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gsm_group_image": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "gsm_group_project_route": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.POINTER(gsm_camera), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "gsm_group_render_strip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "gsm_group_signal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "gsm_group_wait": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]),
    "gsm_group_record_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]),
    "gsm_render_strips": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.POINTER(gsm_camera), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    "gsm_probe_math": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "gsm_blend_exp_mode": (C.c_int, [C.c_int]),
    "gsm_status_string": (C.c_char_p, [C.c_int]),
    "gsm_last_error_string": (C.c_char_p, []),
    "gsm_abi_version": (C.c_int, []),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libgsm_b200.so. Raises if it is absent -- there is no CPU or PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "gsm_renderer_b200 has no fallback path.")
        try:
            l = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)  # AttributeError here means the header and the library diverged
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def ptr(obj) -> int | None:
    """Device or host pointer of a torch tensor / numpy array / int / None."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return obj
    if hasattr(obj, "data_ptr"):
        return int(obj.data_ptr())
    if hasattr(obj, "ctypes"):
        return int(obj.ctypes.data)
    raise TypeError(f"cannot take the address of {type(obj)!r}")


def stream_handle(stream) -> int | None:
    """cudaStream_t of a torch.cuda.Stream / int / None (None = the legacy default stream)."""
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream or None
    if hasattr(stream, "cuda_stream"):
        return int(stream.cuda_stream) or None
    raise TypeError(f"not a stream: {type(stream)!r}")
