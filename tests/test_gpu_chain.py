"""Back-to-back frames, chained with programmatic dependent launch and never synchronised in between (parity
item 3): sizes at which the whole projection grid is co-resident (65 536 .. 150 000 Gaussians, where the
frame state is cleared INSIDE the projection kernel), the camera alternating between a near-empty and a full view so that a
stale ticket / status word of the previous frame would name a valid tile of this one. Every frame is kept and compared.
Also: a frame captured into a CUDA graph (gsm.h: "safe to capture into a CUDA graph from the second call on") replays to
the same bytes as the stream-launched frame, mono and stereo."""
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


def _cams(W, H):
    from gsm_renderer_b200.renderer import CameraParams
    proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
    fx, fy = syn.focal_lengths(W, H)
    full = CameraParams(np.eye(4, dtype=np.float32), proj, (0, 0, 0), fx, fy, 0.1, 100.0)
    # looking away from the cloud along +x: a sliver of it stays in view (few visible Gaussians, few tiles)
    v = syn.look_at_opencv((0.0, 0.0, 0.0), (1.0, 0.0, 0.25))
    side = CameraParams(v, proj, (0, 0, 0), fx, fy, 0.1, 100.0)
    return full, side


@pytest.mark.parametrize("n", [65_536, 100_000, 150_000])
def test_alternating_frames_back_to_back_gpu(pu, n):
    import torch
    from gsm_renderer_b200.renderer import DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig, RenderPrecision
    W, H = 640, 368
    cl = syn.synthetic_cloud(n, 1, seed=5 + n, scale_median=0.02)
    g, h = pu.make_scene_inputs(cl, "float16")
    full, side = _cams(W, H)
    dev = torch.device("cuda:0")
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=n, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear))
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    inp = GaussianInput(tg, th, n, 4)
    s = torch.cuda.current_stream()
    refs, counts = [], []
    for cam in (full, side):
        ref = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
        r.render(s, ref, None, inp, cam, W, H)
        torch.cuda.synchronize()
        hd = r.debugReadHeader()
        refs.append(ref)
        counts.append((hd.visibleCount, hd.totalInstances))
    assert counts[0][0] > 4 * max(counts[1][0], 1), f"the side view is not much emptier than the full one: {counts}"
    frames = 200
    out = torch.zeros((frames, H, W, 4), dtype=torch.int16, device=dev)
    for i in range(frames):   # no synchronisation: 200 frames chained on one stream
        r.render(s, out[i], None, inp, full if i % 2 == 0 else side, W, H)
    torch.cuda.synchronize()
    bad = [i for i in range(frames) if not torch.equal(out[i], refs[i % 2])]
    assert not bad, f"{len(bad)} of {frames} chained frames differ from the synchronised ones, first {bad[:8]}"
    r.close()


def _capture(fn, stream):
    import torch
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        fn(torch.cuda.current_stream())
    return g


def test_cuda_graph_capture_mono_and_stereo_gpu(pu):
    import torch
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision, StereoCameraParams, StereoRenderTarget)
    W, H, n = 960, 540, 120_000
    cl = syn.synthetic_cloud(n, 3, seed=11, scale_median=0.015)
    g, h = pu.make_scene_inputs(cl, "float16")
    full, side = _cams(W, H)
    dev = torch.device("cuda:0")
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=n, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear))
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    inp = GaussianInput(tg, th, n, 16)
    s = torch.cuda.current_stream()
    ref = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    refd = torch.zeros((H, W), dtype=torch.int16, device=dev)
    r.render(s, ref, refd, inp, full, W, H)          # first call: allocates the arena (not capturable), also the reference frame
    torch.cuda.synchronize()
    out = torch.zeros_like(ref)
    outd = torch.zeros_like(refd)
    cs = torch.cuda.Stream(device=dev)
    graph = _capture(lambda st: r.render(st, out, outd, inp, full, W, H), cs)
    for _ in range(3):
        out.fill_(0x7E00)
        outd.fill_(0x7E00)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref) and torch.equal(outd, refd), "graph-replayed mono frame differs from the stream-launched one"
    # a different camera between replays must not leak state into the captured frame
    r.render(s, torch.zeros_like(ref), None, inp, side, W, H)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)

    lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    lv[3, 0], rv[3, 0] = 0.032, -0.032
    proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
    fx, fy = syn.focal_lengths(W, H)
    cams = StereoCameraParams(CameraParams(lv, proj, (-0.032, 0, 0), fx, fy, 0.1, 100.0), CameraParams(rv, proj, (0.032, 0, 0), fx, fy, 0.1, 100.0))
    sref = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    r.renderStereo(s, StereoRenderTarget.sideBySide(sref), inp, cams, W, H)
    torch.cuda.synchronize()
    sout = torch.zeros_like(sref)
    sgraph = _capture(lambda st: r.renderStereo(st, StereoRenderTarget.sideBySide(sout), inp, cams, W, H), cs)
    for _ in range(2):
        sout.fill_(0x7E00)
        sgraph.replay()
        torch.cuda.synchronize()
        assert torch.equal(sout, sref), "graph-replayed stereo frame differs from the stream-launched one"
    r.close()


def test_stereo_chained_frames_do_not_stall_gpu(pu):
    """Round 2 regression: with block-index first tiles in the sort passes, a stereo frame's dependent-launch chain starved
    itself (CTAs holding ticketed tiles spun on tiles of CTAs that were not resident yet) and a frame took SECONDS while still
    producing the right image. Frames chained on one stream must take milliseconds, and equal the synchronised frame."""
    import torch
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision, StereoCameraParams, StereoRenderTarget)
    W, H, n = 1920, 1080, 400_000
    cl = syn.synthetic_cloud(n, 3, seed=42, scale_median=0.015)
    g, h = pu.make_scene_inputs(cl, "float16")
    dev = torch.device("cuda:0")
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=2_000_000, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear))
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    inp = GaussianInput(tg, th, n, 16)
    proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
    fx, fy = syn.focal_lengths(W, H)
    lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    lv[3, 0], rv[3, 0] = 0.032, -0.032
    cams = StereoCameraParams(CameraParams(lv, proj, (-0.032, 0, 0), fx, fy, 0.1, 100.0), CameraParams(rv, proj, (0.032, 0, 0), fx, fy, 0.1, 100.0))
    s = torch.cuda.current_stream()
    ref = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    r.renderStereo(s, StereoRenderTarget.sideBySide(ref), inp, cams, W, H)
    torch.cuda.synchronize()
    out = torch.zeros_like(ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r.renderStereo(s, StereoRenderTarget.sideBySide(out), inp, cams, W, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    assert torch.equal(out, ref)
    assert ms < 20.0, f"{ms:.1f} ms per chained stereo frame: the dependent-launch chain is starving itself"
    r.close()
