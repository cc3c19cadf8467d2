"""GlobalRenderer (GlobalRenderer.swift, GlobalShaders.metal) through the C ABI (gsm_render_global)
against the oracle's restatement (oracle/gsm_oracle.c, gsmo_render_global): bounds, visibility, assignment totals, the sorted
[tile:16][half depth:16] keys and indices, tile headers, active set and every pixel, bit for bit. The reference pins only the sort
(GlobalUnitTests.swift:23-176, tests/test_oracle_kats.py); everything else is parity-unpinned, as for the DepthFirst path."""
import numpy as np
import pytest

from gsm_renderer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    import tests.parity_util as p
    return p


def _run(oracle, pu, cloud, precision, W, H, max_wh=None, max_gaussians=None, srgb=False, near=0.1, far=100.0, view=None,
         position=(0.0, 0.0, 0.0)):
    import torch
    from gsm_renderer_b200.renderer import (GaussianColorSpace, GaussianInput, GlobalRenderer, RendererConfig, RenderPrecision)
    maxW, maxH = max_wh or (W, H)
    G = max_gaussians or cloud.count
    g, h = pu.make_scene_inputs(cloud, precision)
    cam = pu.default_camera(W, H, near, far, view, position)
    ocam = oracle.make_camera(cam.viewMatrix, cam.projectionMatrix, cam.position, W, H, cam.near, cam.far, cloud.sh_components, g.shape[0], srgb)
    fr = oracle.OracleGlobalFrame(G, maxW, maxH)
    oc, od = fr.render(g, h, oracle.F16 if precision == "float16" else oracle.F32, ocam, W, H)
    dev = torch.device("cuda:0")
    r = GlobalRenderer(device=0, config=RendererConfig(maxGaussians=G, maxWidth=maxW, maxHeight=maxH,
                                                       precision=RenderPrecision.float16 if precision == "float16" else RenderPrecision.float32,
                                                       gaussianColorSpace=GaussianColorSpace.srgb if srgb else GaussianColorSpace.linear))
    try:
        tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
        th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
        color = torch.full((H, W, 4), 0x7E00, dtype=torch.int16, device=dev)
        depth = torch.full((H, W), 0x7E00, dtype=torch.int16, device=dev)
        r.render(torch.cuda.current_stream(), color, depth, GaussianInput(tg, th, g.shape[0], cloud.sh_components), cam, W, H)
        torch.cuda.synchronize()
        gc, gd = color.cpu().numpy().view(np.uint16), depth.cpu().numpy().view(np.uint16)
        hd = r.debugReadHeader()
        info = fr.info
        N = g.shape[0]
        assert np.array_equal(r.debugReadBounds(N), fr.bounds[:N]), "bounds"
        vis = fr.mask[:N] > 0
        assert r.debugReadRenderData(N)[vis].tobytes() == fr.renderData[:N][vis].tobytes(), "renderData of unculled Gaussians"
        for k in ("visibleCount", "totalAssignments", "paddedCount", "overflow", "activeTileCount"):
            assert int(hd[k]) == int(getattr(info, k)), f"header.{k}: gpu {int(hd[k])} oracle {int(getattr(info, k))}"
        V, A = int(info.visibleCount), int(info.totalAssignments)
        assert np.array_equal(r.debugReadVisibleIndices(V), fr.visibleIndices[:V]), "visible indices"
        assert np.array_equal(r.debugReadSortedKeys(A), fr.sortedKeys[:A]), "sorted keys"
        assert np.array_equal(r.debugReadSortedIndices(A), fr.sortedIndices[:A]), "sorted indices"
        T = fr.tileHeaders.shape[0]
        assert np.array_equal(r.debugReadTileHeaders(T), fr.tileHeaders), "tile headers"
        act = np.sort(r.debugReadActiveTiles(int(info.activeTileCount)))
        assert np.array_equal(act, np.nonzero(fr.tileHeaders[:, 1] > 0)[0].astype(np.uint32)), "active tiles (as a set)"
        pu.compare_pixels(gc, oc, True, precision == "float16", "colour")
        pu.compare_pixels(gd, od, True, precision == "float16", "depth")
        return dict(V=V, A=A, active=int(info.activeTileCount), overflow=int(info.overflow))
    finally:
        r.close()


@pytest.mark.parametrize("precision,deg,srgb", [("float16", 3, False), ("float32", 1, True), ("float16", 0, False), ("float32", 2, False)])
def test_global_frames_gpu(oracle, pu, precision, deg, srgb):
    cl = syn.synthetic_cloud(50_000, deg, seed=40 + deg, scale_median=0.015)
    res = _run(oracle, pu, cl, precision, 1280, 720, srgb=srgb)
    assert res["V"] > 20_000 and res["A"] > res["V"] and res["overflow"] == 0


def test_global_reference_fixture_gpu(oracle, pu):
    # the reference's own small scenes (TestUtils.swift:144-231) and a moved camera
    res = _run(oracle, pu, syn.pipeline_stages_scene(), "float32", 640, 480, near=0.1, far=10.0, srgb=True)
    assert 0 < res["V"] <= 1000
    cl = syn.synthetic_cloud(30_000, 2, seed=9, scale_median=0.02)
    view, pos = syn.orbit_cameras(3, seed=3)[2]
    _run(oracle, pu, cl, "float16", 1280, 720, view=view, position=pos)


def test_global_limits_larger_than_frame_gpu(oracle, pu):
    # tiles come from the LIMITS (GlobalRenderer.swift:26-49), the camera from the frame: a 1280x720 frame on a 1920x1080 renderer
    cl = syn.synthetic_cloud(30_000, 1, seed=5, scale_median=0.02)
    _run(oracle, pu, cl, "float16", 1280, 720, max_wh=(1920, 1080))
    _run(oracle, pu, cl, "float16", 1919, 1079, max_wh=(1920, 1080))


def test_global_overflow_and_edges_gpu(oracle, pu):
    # 4 * maxGaussians assignments exceeded: clamped, flagged, later stores dropped (GlobalShaders.metal:656-662, :695-700)
    big = syn.generate_visible_gaussians(600, seed=42)
    res = _run(oracle, pu, big, "float32", 640, 480, near=0.1, far=10.0)
    assert res["overflow"] == 1
    # nothing visible; one Gaussian
    base = syn.synthetic_cloud(64, 0, seed=1)
    behind = syn.Cloud(base.positions * [1, 1, -1], base.scales, base.rotations, base.opacities, base.harmonics, 1)
    assert _run(oracle, pu, behind, "float32", 320, 200)["V"] == 0
    one = syn.Cloud(base.positions[:1] * 0 + [[0, 0, 5]], base.scales[:1] * 0 + 0.05, base.rotations[:1], base.opacities[:1] * 0 + 0.9,
                    base.harmonics[:1], 1)
    assert _run(oracle, pu, one, "float32", 320, 200)["V"] == 1


def test_global_config2_size_gpu(oracle, pu):
    # the bench workload's cloud through the second renderer (1 M Gaussians, SH3, float16, 1080p)
    cl = syn.synthetic_cloud(1_000_000, 3, seed=42, scale_median=0.015)
    res = _run(oracle, pu, cl, "float16", 1920, 1080)
    assert res["V"] > 600_000 and res["overflow"] == 0
