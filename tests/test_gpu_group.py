"""gsm_group: the strip-sharded frame whose exchange step is the routing kernel's own peer stores (csrc/group.cu), against the
single-GPU frame bit for bit. Ranks are emulated on cuda:0 (one renderer + one group per rank, windows connected with
gsm_group_connect_local: the same kernels and the same mailbox protocol as between GPUs, the "peer" stores just land in local
HBM); tests/mgpu_worker.py runs the same comparison between processes over cudaIpc when the box has >= 2 GPUs."""
import numpy as np
import pytest

from gsm_renderer_b200 import multigpu as mg
from gsm_renderer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _renderer(N, W, H):
    from gsm_renderer_b200.renderer import DepthFirstRenderer, GaussianColorSpace, RendererConfig, RenderPrecision
    return DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                              gaussianColorSpace=GaussianColorSpace.linear))


def _cams(W, H):
    from gsm_renderer_b200.renderer import CameraParams
    proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
    fx, fy = syn.focal_lengths(W, H)
    full = CameraParams(np.eye(4, dtype=np.float32), proj, (0, 0, 0), fx, fy, 0.1, 100.0)
    side = CameraParams(syn.look_at_opencv((0.0, 0.0, 0.0), (1.0, 0.0, 0.25)), proj, (0, 0, 0), fx, fy, 0.1, 100.0)
    away = CameraParams(syn.look_at_opencv((0.0, 0.0, 0.0), (0.0, 0.0, -1.0)), proj, (0, 0, 0), fx, fy, 0.1, 100.0)
    return full, side, away


@pytest.mark.parametrize("world,weighted", [(2, False), (4, False), (8, False), (3, True)])
def test_group_strips_equal_single_gpu(world, weighted):
    import torch
    import tests.parity_util as pu
    from gsm_renderer_b200.renderer import GaussianInput
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    W, H, N = 1920, 1080, 80_000
    tilesX, tilesY = 120, 68
    cl = syn.synthetic_cloud(N, 3, seed=13, scale_median=0.015)
    g, h = pu.make_scene_inputs(cl, "float16")
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    s = torch.cuda.current_stream()
    single = _renderer(N, W, H)
    shards = mg.partition_range(N, world)
    cap = max(c for _, c in shards)
    rs = [_renderer(N, W, H) for _ in range(world)]
    groups = [mg.RendererGroup(rs[k], k, world, cap, W * H * 8, W * H * 2) for k in range(world)]
    for gr in groups:
        gr.connect_local(groups)
    if weighted:  # uneven strips, one of them empty
        rows = [0, 30, 30, tilesY]
    else:
        rows = mg.strip_row_starts(mg.partition_tile_rows(tilesY, world))
    img_c, img_d = groups[0].image_tensors(0, W, H, dev)
    frame_id = 0
    for cam in _cams(W, H) + _cams(W, H)[:1]:   # full, a sliver, nothing visible at all, full again
        ref_c = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
        ref_d = torch.zeros((H, W), dtype=torch.int16, device=dev)
        single.render(s, ref_c, ref_d, GaussianInput(tg, th, N, 16), cam, W, H)
        torch.cuda.synchronize()
        ref_headers = single.debugReadTileHeaders(tilesX * tilesY)
        ref_inst = single.debugReadInstanceGaussianIndices(single.debugReadHeader().totalInstances)
        assert single.debugReadHeader().overflow == 0
        frame_id += 1
        img_c.fill_(0x7E00)
        img_d.fill_(0x7E00)
        for k, (a, c) in enumerate(shards):   # phase 1 on every rank, then phase 2: the ranks share one stream here
            groups[k].projectRoute(s, tg[a * 32:(a + c) * 32], th[a * 96:(a + c) * 96], a, c, 16, cam, W, H, rows)
        for k in range(world):
            groups[k].renderStrip(s, img_c, img_d, W, H, rows)
            groups[k].signal(s, 0, frame_id)
        groups[0].wait(s, (1 << world) - 1, frame_id)
        torch.cuda.synchronize()
        assert torch.equal(img_c, ref_c), f"assembled colour differs from the single-GPU frame (world {world}, frame {frame_id})"
        assert torch.equal(img_d, ref_d), f"assembled depth differs from the single-GPU frame (world {world}, frame {frame_id})"
        total_records = 0
        for k in range(world):  # per-tile lists of every strip == the single-GPU lists (offsets are strip-local)
            row0, row1 = rows[k], rows[k + 1]
            hdk = rs[k].debugReadHeader()
            total_records += hdk.visibleCount
            if row1 == row0:
                assert hdk.visibleCount == 0 and hdk.totalInstances == 0
                continue
            hd = rs[k].debugReadTileHeaders(tilesX * tilesY)
            inst = rs[k].debugReadInstanceGaussianIndices(hdk.totalInstances)
            assert np.array_equal(hd[row0 * tilesX:row1 * tilesX, 1], ref_headers[row0 * tilesX:row1 * tilesX, 1])
            for t in (row0 * tilesX, row0 * tilesX + 61, row1 * tilesX - 1):
                a0, c0 = ref_headers[t]
                a1, c1 = hd[t]
                assert c0 == c1 and np.array_equal(ref_inst[a0:a0 + c0], inst[a1:a1 + c1])
        V = single.debugReadHeader().visibleCount
        assert V <= total_records <= 2 * V + world, "a rank ingests only the records of its strip (plus those straddling a border)"
    for gr in groups:
        gr.close()
    for r in rs + [single]:
        r.close()


def test_group_argument_errors():
    import torch
    from gsm_renderer_b200.renderer import RendererError
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    r = _renderer(1000, 64, 64)
    with pytest.raises(RendererError):
        mg.RendererGroup(r, 2, 2, 500)          # rank >= world
    with pytest.raises(RendererError):
        mg.RendererGroup(r, 0, 9, 500)          # more than 8 ranks
    with pytest.raises(RendererError):
        mg.RendererGroup(r, 0, 2, 5000)         # shard larger than maxGaussians
    gr = mg.RendererGroup(r, 0, 2, 500)
    full = _cams(64, 64)[0]
    with pytest.raises(RendererError):          # not connected
        gr.projectRoute(None, 0, 0, 0, 0, 1, full, 64, 64, [0, 2, 4])
    gr.close()
    r.close()
