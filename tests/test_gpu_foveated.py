"""GPU-vs-oracle parity of StereoRenderTarget.foveated through the C ABI: the resampling copy alone
on random images, and whole foveated frames (scene transform, viewports, rate map, attachment format). Bit for bit."""
import numpy as np
import pytest

from gsm_renderer_b200 import synthetic as syn
from tests import foveation_util as fv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    import tests.parity_util as p
    return p


def _renderer(W, H, n=1000, flip=True, precision="float16"):
    from gsm_renderer_b200.renderer import DepthFirstRenderer, GaussianColorSpace, RendererConfig, RenderPrecision
    return DepthFirstRenderer(device=0, config=RendererConfig(
        maxGaussians=n, maxWidth=W, maxHeight=H,
        precision=RenderPrecision.float16 if precision == "float16" else RenderPrecision.float32,
        gaussianColorSpace=GaussianColorSpace.linear), stereoCopyFlipY=flip)


def _sbs(c2):
    """(2, H, W, 4) slices -> (H, 2W, 4) side by side, the layout of the device's intermediate image."""
    return np.ascontiguousarray(np.concatenate([c2[0], c2[1]], axis=1))


CASES = [
    # fmt, arrayLength, flip, texture (w, h), viewports, rate layers (None | "one" | "two"), row padding
    (0, 1, True, None, "sbs", None, 0),
    (0, 2, False, None, "full", None, 0),
    (2, 2, True, None, "full", "two", 0),
    (2, 1, True, None, "sbs", "one", 64),
    (1, 2, False, (300, 200), "scaled", None, 12),
    (3, 2, True, (257, 131), "offset", "one", 0),
    (4, 1, False, (500, 190), "overlap", None, 4),
    (0, 2, True, (97, 61), "scaled", "two", 8),
]


@pytest.mark.parametrize("fmt,array_length,flip,tex,vpkind,rate,pad", CASES)
def test_stereo_copy_matches_oracle_gpu(oracle, pu, fmt, array_length, flip, tex, vpkind, rate, pad):
    import torch
    from gsm_renderer_b200.renderer import FoveatedStereoDrawable, PixelFormat, RasterizationRateMap, Viewport
    rng = np.random.default_rng(fmt * 100 + array_length * 10 + (rate is not None))
    W, H = 211, 157
    if fmt == 0:
        c2 = (rng.standard_normal((2, H, W, 4)) * np.exp(rng.uniform(-6, 6, (2, H, W, 4)))).astype(np.float16)
        c2[0, 5, 7] = [np.inf, -np.inf, np.nan, -0.0]
    else:
        c2 = rng.uniform(-0.1, 1.1, (2, H, W, 4)).astype(np.float16)
        c2[1, 9, 3] = [np.nan, 2.0, -3.0, 0.5]
    c2 = c2.view(np.uint16)
    screen = {"sbs": (2 * W, H), "full": (W, H), "scaled": (320, 240), "offset": (400, 260), "overlap": (500, 190)}[vpkind]
    vps = {"sbs": ((0, 0, W, H), (W, 0, W, H)), "full": ((0, 0, W, H), (0, 0, W, H)),
           "scaled": ((10.25, 7.5, 290.5, 220.75), (0, 0, 320, 240)), "offset": ((30, 20, W, H), (100.5, 60.25, W * 1.25, H * 0.8)),
           "overlap": ((0, 0, 300, 190), (200, 0, 300, 190))}[vpkind]
    layers = None
    if rate == "one":
        layers = [fv.layer(screen[0], screen[1], fv.FOVEATED_H, fv.FOVEATED_V)]
    elif rate == "two":
        layers = [fv.layer(screen[0], screen[1], fv.FOVEATED_H, fv.FOVEATED_V), fv.layer(screen[0], screen[1], (1.0, 0.4), (0.5, 1.0, 0.5))]
    if tex is None:
        tex = (max(l[0].size for l in layers), max(l[1].size for l in layers)) if layers else screen
    tw, th = tex
    px = 8 if fmt == 0 else 4
    row_bytes = tw * px + pad * px
    ref = oracle.stereo_copy_foveated(c2, flip, tw, th, array_length, fmt, vps, rate_layers=layers, row_bytes=row_bytes)
    r = _renderer(W, H, flip=flip)
    dev = torch.device("cuda:0")
    src = torch.from_numpy(_sbs(c2).view(np.int16)).to(dev)
    dst = torch.full((array_length, th, row_bytes), 0xAB, dtype=torch.uint8, device=dev)
    d = FoveatedStereoDrawable(dst, tw, th, array_length, RasterizationRateMap(layers) if layers else None, PixelFormat(fmt), row_bytes)
    r.stereoCopy(torch.cuda.current_stream(), src, W, H, d, Viewport(*vps[0]), Viewport(*vps[1]))
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    assert (ref != 0xAB).any()
    if not np.array_equal(got, ref):
        bad = np.argwhere(got != ref)
        raise AssertionError(f"{len(bad)} bytes differ, first at {bad[0].tolist()}: gpu {got[tuple(bad[0])]} oracle {ref[tuple(bad[0])]}")
    r.close()


def _eye_views(W, H, vps):
    from gsm_renderer_b200.renderer import EyeView, Viewport
    proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
    fx, fy = syn.focal_lengths(W, H)
    lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    lv[3, 0], rv[3, 0] = 0.032, -0.032
    L = EyeView(Viewport(*vps[0]), lv, proj, (-0.032, 0, 0), fx, fy, 0.1, 100.0)
    R = EyeView(Viewport(*vps[1]), rv, proj, (0.032, 0, 0), fx, fy, 0.1, 100.0)
    return L, R


def _scene_transform():
    a = 0.2
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)  # m[col][row]: rotation about y
    m[:3, :3] *= np.float32(1.1)
    m[3, :3] = [0.05, -0.02, 0.3]
    return m


@pytest.mark.parametrize("precision,fmt,array_length,rate", [("float16", 2, 2, True), ("float32", 0, 1, False), ("float16", 3, 1, True)])
def test_foveated_frame_gpu(oracle, pu, precision, fmt, array_length, rate):
    import torch
    from gsm_renderer_b200.renderer import (FoveatedStereoDrawable, GaussianInput, PixelFormat, RasterizationRateMap,
                                            StereoConfiguration, StereoRenderTarget)
    cl = syn.synthetic_cloud(20_000, 2, seed=21, scale_median=0.02)
    g, h = pu.make_scene_inputs(cl, precision)
    W, H = 480, 272
    screen = (W, H) if array_length == 2 else (2 * W, H)
    vps = ((0, 0, W, H), (0, 0, W, H)) if array_length == 2 else ((0, 0, W, H), (W, 0, W, H))
    layers = [fv.layer(screen[0], screen[1], fv.FOVEATED_H, fv.FOVEATED_V)] if rate else None
    tw, th = (layers[0][0].size, layers[0][1].size) if rate else screen
    L, R = _eye_views(W, H, vps)
    scene = _scene_transform()
    ocam = oracle.make_stereo_camera(L.viewMatrix, L.projectionMatrix, L.cameraPosition, R.viewMatrix, R.projectionMatrix,
                                     R.cameraPosition, W, H, 0.1, 100.0, cl.sh_components, cl.count, False, scene=scene)
    fr = oracle.OracleFrame(cl.count, W, H, stereo=True)
    _, slices = fr.render_stereo(g, h, oracle.F16 if precision == "float16" else oracle.F32, ocam, W, H, flip_y=True)
    ref = oracle.stereo_copy_foveated(slices, True, tw, th, array_length, fmt, vps, rate_layers=layers)
    r = _renderer(W, H, cl.count, True, precision)
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    tht = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    px = 8 if fmt == 0 else 4
    dst = torch.full((array_length, th, tw * px), 0xAB, dtype=torch.uint8, device=dev)
    d = FoveatedStereoDrawable(dst, tw, th, array_length, RasterizationRateMap(layers) if layers else None, PixelFormat(fmt))
    target = StereoRenderTarget.foveated(d, StereoConfiguration(L, R, scene))
    for _ in range(2):  # the second frame reuses the intermediate image and the uploaded tables
        r.renderStereo(torch.cuda.current_stream(), target, GaussianInput(tg, tht, cl.count, cl.sh_components), None, W, H)
    torch.cuda.synchronize()
    pu.compare_white_box(r, fr, W, H, cl.count, stereo=True)
    got = dst.cpu().numpy()
    assert np.array_equal(got, ref), f"{np.count_nonzero(got != ref)} bytes differ"
    assert np.count_nonzero(got.reshape(-1, px) != 0xAB) > got.size // 4  # something was drawn
    r.close()


def test_foveated_one_to_one_equals_side_by_side_gpu(pu):
    """A shared rgba16f drawable with viewports (0,0,W,H), (W,0,W,H) and no rate map is the sideBySide target."""
    import torch
    from gsm_renderer_b200.renderer import (CameraParams, FoveatedStereoDrawable, GaussianInput, PixelFormat, StereoCameraParams,
                                            StereoConfiguration, StereoRenderTarget)
    cl = syn.synthetic_cloud(50_000, 3, seed=5, scale_median=0.02)
    g, h = pu.make_scene_inputs(cl, "float16")
    W, H = 640, 360
    L, R = _eye_views(W, H, ((0, 0, W, H), (W, 0, W, H)))
    r = _renderer(W, H, cl.count, True)
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    inp = GaussianInput(tg, th, cl.count, cl.sh_components)
    sbs = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    cams = StereoCameraParams(CameraParams(L.viewMatrix, L.projectionMatrix, L.cameraPosition, L.focalX, L.focalY, 0.1, 100.0),
                              CameraParams(R.viewMatrix, R.projectionMatrix, R.cameraPosition, R.focalX, R.focalY, 0.1, 100.0))
    s = torch.cuda.current_stream()
    r.renderStereo(s, StereoRenderTarget.sideBySide(sbs), inp, cams, W, H)
    dst = torch.zeros((1, H, 2 * W * 8), dtype=torch.uint8, device=dev)
    d = FoveatedStereoDrawable(dst, 2 * W, H, 1, None, PixelFormat.rgba16Float)
    r.renderStereo(s, StereoRenderTarget.foveated(d, StereoConfiguration(L, R)), inp, None, W, H)
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy().reshape(-1), sbs.cpu().numpy().view(np.uint8).reshape(-1))
    r.close()


def test_foveated_argument_errors_gpu(pu):
    import torch
    from gsm_renderer_b200.renderer import FoveatedStereoDrawable, PixelFormat, RendererError, Viewport
    r = _renderer(64, 64)
    dev = torch.device("cuda:0")
    src = torch.zeros((64, 128, 4), dtype=torch.int16, device=dev)
    dst = torch.zeros((2, 64, 64 * 4), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream()
    vp = Viewport(0, 0, 64, 64)
    with pytest.raises(RendererError):
        r.stereoCopy(s, src, 64, 64, FoveatedStereoDrawable(dst, 64, 64, 3, None, PixelFormat.bgra8Unorm), vp, vp)
    with pytest.raises(RendererError):
        r.stereoCopy(s, src, 64, 64, FoveatedStereoDrawable(dst, 64, 64, 2, None, PixelFormat.bgra8Unorm, rowBytes=100), vp, vp)
    with pytest.raises(RendererError):
        r.stereoCopy(s, src, 0, 64, FoveatedStereoDrawable(dst, 64, 64, 2, None, PixelFormat.bgra8Unorm), vp, vp)
    r.close()
