"""torchrun worker: real NCCL run of the strip-sharded frame (all-gather of splat records) and the stereo eye
split on N GPUs, each compared with the single-GPU frame rendered locally on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsm_renderer_b200 import multigpu as mg  # noqa: E402
from gsm_renderer_b200 import synthetic as syn  # noqa: E402
from gsm_renderer_b200.renderer import (DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,  # noqa: E402
                                        RenderPrecision, StereoRenderTarget)
import tests.parity_util as pu  # noqa: E402
from tests.test_gpu_parity import _stereo_inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cl = syn.synthetic_cloud(200_000, 3, seed=17, scale_median=0.015)
    W, H = 1920, 1080
    g, h = pu.make_scene_inputs(cl, "float16")
    cam = pu.default_camera(W, H)
    cfg = RendererConfig(maxGaussians=cl.count, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                         gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=local, config=cfg)
    s = torch.cuda.current_stream()
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    ref = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    r.render(s, ref, None, GaussianInput(tg, th, cl.count, 16), cam, W, H)
    # strip-sharded frame: this rank only holds its shard of the Gaussians
    a, c = mg.partition_range(cl.count, world)[rank]
    shard_g, shard_h = tg[a * 32:(a + c) * 32].clone(), th[a * 96:(a + c) * 96].clone()
    scratch = torch.zeros(max(c, 1) * mg.RECORD_BYTES, dtype=torch.uint8, device=dev)
    out = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    _, counts, strip = mg.render_strips(r, dist, rank, world, s, shard_g, shard_h, (a, c), 16, cam, W, H, out, None, scratch)
    strips = mg.partition_tile_rows(68, world)
    mg.gather_strips(dist, rank, world, out, strips, W, H, root=0)
    torch.cuda.synchronize()
    y0, y1 = strip[0] * 16, min(H, (strip[0] + strip[1]) * 16)
    assert torch.equal(out[y0:y1], ref[y0:y1]), f"rank {rank}: strip differs"
    if rank == 0:
        assert torch.equal(out, ref), "assembled frame differs from the single-GPU frame"
    # ---- the same frame over peer memory (gsm_group): routing kernel stores records into the peers' windows, every rank blends
    # its strip straight into rank 0's window image; several frames with moving cameras exercise the ack / sequence protocol
    cap = max(cc for _, cc in mg.partition_range(cl.count, world))
    grp = mg.RendererGroup(r, rank, world, cap, W * H * 8, W * H * 2)
    grp.connect_distributed(dist)
    rows = mg.strip_row_starts(strips)
    pc, pd = grp.image_ptrs(0)                        # rank 0's image as mapped on this rank (raw device pointers)
    if rank == 0:
        img_c, img_d = grp.image_tensors(0, W, H, dev)
    from tests.test_gpu_group import _cams
    cams3 = _cams(W, H)
    refd = torch.zeros((H, W), dtype=torch.int16, device=dev)
    for f, camf in enumerate([cam, cams3[1], cams3[2], cam, cams3[1]], start=1):
        r.render(s, ref, refd, GaussianInput(tg, th, cl.count, 16), camf, W, H)   # single-GPU frame (arena is re-used below)
        torch.cuda.synchronize()
        dist.barrier()                       # rank 0's previous comparison is over before anyone overwrites its image
        grp.renderStrips(s, pc, pd, shard_g, shard_h, a, c, 16, camf, W, H, rows)
        grp.signal(s, 0, f)
        if rank == 0:
            grp.wait(s, (1 << world) - 1, f)
            torch.cuda.synchronize()
            assert torch.equal(img_c, ref), f"peer-memory frame {f}: assembled colour differs from the single-GPU frame"
            assert torch.equal(img_d, refd), f"peer-memory frame {f}: assembled depth differs from the single-GPU frame"
        torch.cuda.synchronize()
    # back to back without host synchronisation in between (flow control on the device only)
    for f in range(6, 26):
        grp.renderStrips(s, pc, pd, shard_g, shard_h, a, c, 16, cam if f % 2 else cams3[1], W, H, rows)
        grp.signal(s, 0, f)
        if rank == 0:
            grp.wait(s, (1 << world) - 1, f)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        r.render(s, ref, refd, GaussianInput(tg, th, cl.count, 16), cam, W, H)
        torch.cuda.synchronize()
        assert torch.equal(img_c, ref) and torch.equal(img_d, refd), "peer-memory frames back to back: last frame differs"
    dist.barrier()
    # stereo: one eye per GPU
    if world >= 2:
        cams = _stereo_inputs(960, 540)
        small = RendererConfig(maxGaussians=cl.count, maxWidth=960, maxHeight=540, precision=RenderPrecision.float16,
                               gaussianColorSpace=GaussianColorSpace.linear)
        rs = DepthFirstRenderer(device=local, config=small)
        joint = torch.zeros((540, 1920, 4), dtype=torch.int16, device=dev)
        inp = GaussianInput(tg, th, cl.count, 16)
        rs.renderStereo(s, StereoRenderTarget.sideBySide(joint), inp, cams, 960, 540)
        tgt = torch.zeros((540, 1920, 4), dtype=torch.int16, device=dev)
        if rank < 2:  # ranks 0/1 take one eye each; point-to-point only, no collective
            mg.render_stereo_split(rs, dist, rank, 2, s, tgt, inp, cams, 960, 540)
        torch.cuda.synchronize()
        if rank == 0:
            assert torch.equal(tgt, joint), "eye-split stereo differs from the joint frame"
        # the same split with the right eye blended straight into rank 0's window image over NVLink (no copy, no collective)
        sgrp = mg.RendererGroup(rs, rank, world, 1, 540 * 1920 * 8, 0)
        sgrp.connect_distributed(dist)
        sp, _ = sgrp.image_ptrs(0)
        if rank == 0:
            simg, _ = sgrp.image_tensors(0, 1920, 540, dev, depth=False)
            simg.fill_(0x7E00)
        torch.cuda.synchronize()
        dist.barrier()
        if rank < 2:
            rs.renderStereo(s, StereoRenderTarget.sideBySide(sp), inp, cams, 960, 540, eyeMask=1 << rank)
            sgrp.signal(s, 0, 1)
        if rank == 0:
            sgrp.wait(s, 0b11, 1)
            torch.cuda.synchronize()
            assert torch.equal(simg, joint), "peer-memory eye split differs from the joint frame"
        torch.cuda.synchronize()
        dist.barrier()
        sgrp.close()
        rs.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK", counts)
    grp.close()
    r.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
