"""BASELINE.json's configurations at FULL size, GPU vs oracle bit for bit through the C ABI (the oracle needs a few seconds per
frame at these sizes, so each configuration is one frame): C2 = the bench workload, C3 = 6 M Gaussians at 3840x2160,
C4 = stereo 2 x 1080p at 1 M, C5 = one 720p view of the 3 M cloud. C1 is tests/test_gpu_parity.py::test_config1_50k_sh1_f32_gpu."""
import numpy as np
import pytest

import bench
from gsm_renderer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    import tests.parity_util as p
    return p


def _workload(name):
    N, deg, prec, W, H, sm, _ = bench.WORKLOADS[name]
    return syn.synthetic_cloud(N, deg, seed=42, scale_median=sm), prec, W, H


def test_config2_bench_workload_gpu(oracle, pu):
    cl, prec, W, H = _workload("C2")
    res = pu.run_mono_case(oracle, cl, prec, W, H, near=bench.NEAR, far=bench.FAR)
    assert res["V"] == 709_202 and res["I"] == 2_909_567  # the counts bench.py reports for this seed


def test_config5_one_view_gpu(oracle, pu):
    cl, prec, W, H = _workload("C5v")
    res = pu.run_mono_case(oracle, cl, prec, W, H, near=bench.NEAR, far=bench.FAR)
    assert res["V"] > 1_500_000 and res["I"] > res["V"]


def test_config3_6m_at_4k_gpu(oracle, pu):
    """tools/mgpu_bench.py's C3 cloud on one GPU (the strip-sharded form is checked against this path in test_gpu_multi.py)."""
    cl = syn.synthetic_cloud(6_000_000, 3, seed=42, scale_median=0.008)
    res = pu.run_mono_case(oracle, cl, "float16", 3840, 2160, near=bench.NEAR, far=bench.FAR)
    assert res["V"] > 3_000_000 and res["I"] > 10_000_000


def _stereo_case(oracle, pu, cl, prec, W, H, max_gaussians):
    import torch
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision, StereoCameraParams, StereoRenderTarget)
    g, h = pu.make_scene_inputs(cl, prec)
    proj = syn.make_projection_matrix(W, H, bench.NEAR, bench.FAR)
    fx, fy = syn.focal_lengths(W, H)
    lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    lv[3, 0], rv[3, 0] = 0.032, -0.032
    cams = StereoCameraParams(CameraParams(lv, proj, (-0.032, 0, 0), fx, fy, bench.NEAR, bench.FAR),
                              CameraParams(rv, proj, (0.032, 0, 0), fx, fy, bench.NEAR, bench.FAR))
    ocam = oracle.make_stereo_camera(lv, proj, (-0.032, 0, 0), rv, proj, (0.032, 0, 0), W, H, bench.NEAR, bench.FAR,
                                     cl.sh_components, cl.count, False)
    fr = oracle.OracleFrame(max_gaussians, W, H, stereo=True)
    ref, _ = fr.render_stereo(g, h, oracle.F16, ocam, W, H, flip_y=True)
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=max_gaussians, maxWidth=W, maxHeight=H,
                                                           precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear), stereoCopyFlipY=True)
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    sbs = torch.full((H, 2 * W, 4), 0x7E00, dtype=torch.int16, device=dev)
    r.renderStereo(torch.cuda.current_stream(), StereoRenderTarget.sideBySide(sbs), GaussianInput(tg, th, cl.count, cl.sh_components),
                   cams, W, H)
    torch.cuda.synchronize()
    pu.compare_white_box(r, fr, W, H, cl.count, stereo=True)
    pu.compare_pixels(sbs.cpu().numpy().view(np.uint16), ref, True, True, "stereo colour")
    hd = r.debugReadHeader()
    out = (hd.visibleCount, hd.totalInstances, hd.overflow)
    r.close()
    return out


def test_config4_stereo_1m_gpu(oracle, pu):
    """C4 at the reference's default capacity (RendererConfig.maxGaussians = 6 000 000, GRP.swift:211-218 => 24 M instances):
    the union boxes of the 1 M cloud fit, so the frame is complete (overflow == 0), not the nearest 4 M instances."""
    cl, prec, W, H = _workload("C2")
    V, I, overflow = _stereo_case(oracle, pu, cl, prec, W, H, 6_000_000)
    assert overflow == 0 and I > 4_000_000, (V, I, overflow)


def test_stereo_instance_overflow_gpu(oracle, pu):
    """The truncating case stays covered, small: capacity == count, union boxes exceed 4 * maxGaussians (DFS.metal:707)."""
    cl = syn.synthetic_cloud(60_000, 3, seed=42, scale_median=0.03)
    V, I, overflow = _stereo_case(oracle, pu, cl, "float16", 1920, 1080, cl.count)
    assert overflow == 1 and I == 4 * cl.count, (V, I, overflow)
