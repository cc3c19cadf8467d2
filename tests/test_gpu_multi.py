"""Multi-GPU paths against the single-GPU frame, bit for bit. On a one-GPU box the ranks
are emulated in sequence on cuda:0 (same kernels, same records); with >= 2 GPUs a real NCCL run is added."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gsm_renderer_b200 import multigpu as mg
from gsm_renderer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mk(cl, precision, W, H, maxG):
    import torch
    from gsm_renderer_b200.renderer import (DepthFirstRenderer, GaussianColorSpace, RendererConfig, RenderPrecision)
    cfg = RendererConfig(maxGaussians=maxG, maxWidth=W, maxHeight=H,
                         precision=RenderPrecision.float16 if precision == "float16" else RenderPrecision.float32,
                         gaussianColorSpace=GaussianColorSpace.linear)
    return DepthFirstRenderer(device=0, config=cfg)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_strip_sharded_frame_equals_single_gpu(world):
    import torch
    import tests.parity_util as pu
    from gsm_renderer_b200.renderer import GaussianInput
    cl = syn.synthetic_cloud(80_000, 3, seed=13, scale_median=0.015)  # no 4N overflow: truncation is per strip
    precision, W, H = "float16", 1920, 1080
    g, h = pu.make_scene_inputs(cl, precision)
    cam = pu.default_camera(W, H)
    dev = torch.device("cuda:0")
    r = _mk(cl, precision, W, H, cl.count)
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    ref_c = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    ref_d = torch.zeros((H, W), dtype=torch.int16, device=dev)
    s = torch.cuda.current_stream()
    r.render(s, ref_c, ref_d, GaussianInput(tg, th, cl.count, 16), cam, W, H)
    torch.cuda.synchronize()
    T = 120 * 68
    ref_headers = r.debugReadTileHeaders(T)
    assert r.debugReadHeader().overflow == 0
    I = r.debugReadHeader().totalInstances
    ref_inst = r.debugReadInstanceGaussianIndices(I)
    # ---- emulated ranks: project shards, concatenate records in rank order, render strips
    shards = mg.partition_range(cl.count, world)
    recs = []
    for (a, c) in shards:
        scratch = torch.zeros(max(c, 1) * mg.RECORD_BYTES, dtype=torch.uint8, device=dev)
        n = r.stripProject(s, tg[a * 32:(a + c) * 32], th[a * 96:(a + c) * 96], a, c, 16, cam, W, H, scratch)
        recs.append(scratch[: n * mg.RECORD_BYTES].clone())
    allrec = torch.cat(recs)
    total = allrec.numel() // mg.RECORD_BYTES
    out_c = torch.full((H, W, 4), 0x7E00, dtype=torch.int16, device=dev)
    out_d = torch.full((H, W), 0x7E00, dtype=torch.int16, device=dev)
    strips = mg.partition_tile_rows(68, world)
    for (row0, rows) in strips:
        r.stripRender(s, out_c, out_d, allrec, total, W, H, row0, rows)
        torch.cuda.synchronize()
        # per-tile lists of the strip == the single-GPU lists (offsets are strip-local)
        hd = r.debugReadTileHeaders(T)
        inst = r.debugReadInstanceGaussianIndices(r.debugReadHeader().totalInstances)
        for t in (row0 * 120, row0 * 120 + 61, (row0 + rows) * 120 - 1):
            a0, c0 = ref_headers[t]
            a1, c1 = hd[t]
            assert c0 == c1 and np.array_equal(ref_inst[a0:a0 + c0], inst[a1:a1 + c1])
        assert np.array_equal(hd[row0 * 120:(row0 + rows) * 120, 1], ref_headers[row0 * 120:(row0 + rows) * 120, 1])
    torch.cuda.synchronize()
    assert torch.equal(out_c, ref_c) and torch.equal(out_d, ref_d)
    r.close()


def test_empty_strip_frame_after_a_full_one_shows_the_clear_colour():
    """gsm_strip_render with recordCount == 0 launches nothing that writes the frame header; without clearing
    it the sorts and the blend would run on the previous frame's counts. A strip that sees nothing (the camera turned away) must
    show the clear colour -- alpha 1, everything else 0 -- and report zero visible Gaussians and instances."""
    import torch
    import tests.parity_util as pu
    cl = syn.synthetic_cloud(30_000, 1, seed=21, scale_median=0.02)
    precision, W, H = "float16", 1280, 720
    g, h = pu.make_scene_inputs(cl, precision)
    cam = pu.default_camera(W, H)
    dev = torch.device("cuda:0")
    r = _mk(cl, precision, W, H, cl.count)
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    s = torch.cuda.current_stream()
    scratch = torch.zeros(cl.count * mg.RECORD_BYTES, dtype=torch.uint8, device=dev)
    n = r.stripProject(s, tg, th, 0, cl.count, cl.sh_components, cam, W, H, scratch)
    assert n > 1000
    out_c = torch.full((H, W, 4), 0x7E00, dtype=torch.int16, device=dev)
    out_d = torch.full((H, W), 0x7E00, dtype=torch.int16, device=dev)
    rows = (H + 15) // 16
    r.stripRender(s, out_c, out_d, scratch, n, W, H, 0, rows)          # a full frame first
    torch.cuda.synchronize()
    assert r.debugReadHeader().visibleCount > 1000 and int((out_c[..., 0] != 0).sum()) > 0
    out_c.fill_(0x7E00)
    out_d.fill_(0x7E00)
    r.stripRender(s, out_c, out_d, scratch, 0, W, H, 0, rows)          # then a frame without a single record
    torch.cuda.synchronize()
    hd = r.debugReadHeader()
    assert hd.visibleCount == 0 and hd.totalInstances == 0
    c = out_c.cpu().numpy().view(np.uint16)
    assert np.all(c[..., :3] == 0) and np.all(c[..., 3] == 0x3C00)     # colour (0, 0, 0), alpha 1.0h on every pixel
    assert np.all(out_d.cpu().numpy().view(np.uint16) == 0)
    r.close()


def test_stereo_eye_split_equals_joint():
    import torch
    import tests.parity_util as pu
    from gsm_renderer_b200.renderer import GaussianInput, StereoRenderTarget
    from tests.test_gpu_parity import _stereo_inputs
    cl = syn.synthetic_cloud(60_000, 3, seed=3, scale_median=0.02)
    W, H = 960, 540
    g, h = pu.make_scene_inputs(cl, "float16")
    dev = torch.device("cuda:0")
    r = _mk(cl, "float16", W, H, cl.count)
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    cams = _stereo_inputs(W, H)
    inp = GaussianInput(tg, th, cl.count, 16)
    s = torch.cuda.current_stream()
    joint = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    r.renderStereo(s, StereoRenderTarget.sideBySide(joint), inp, cams, W, H)
    split = torch.full((H, 2 * W, 4), 0x7E00, dtype=torch.int16, device=dev)
    r.renderStereo(s, StereoRenderTarget.sideBySide(split), inp, cams, W, H, eyeMask=1)
    torch.cuda.synchronize()
    assert torch.equal(split[:, :W], joint[:, :W]) and bool((split[:, W:] == 0x7E00).all())
    r.renderStereo(s, StereoRenderTarget.sideBySide(split), inp, cams, W, H, eyeMask=2)
    torch.cuda.synchronize()
    assert torch.equal(split, joint)
    r.close()


def test_nccl_two_ranks_strips_and_views():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0 and "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
