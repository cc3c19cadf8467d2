"""Multi-GPU measurements of the shard modes other than "views" (SURVEY.md 8(e)), run under torchrun:
  C3  one large frame (6 M Gaussians SH3 f16, 3840x2160): Gaussians sharded by gid range, ONE NCCL all-gather of the
      48-byte splat records, each rank sorts + blends its strip of tile rows, strips gathered on rank 0;
  C4  stereo 2 x (1920x1080) at 1 M Gaussians: identical joint stages on ranks 0/1, one eye blended per GPU, one
      point-to-point copy of the right half.
Prints one JSON line (rank 0). Device time = CUDA events on each rank, max over ranks; inputs resident in HBM.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mgpu_bench.py [--small]"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsm_renderer_b200 import multigpu as mg, synthetic as syn
from gsm_renderer_b200.renderer import (DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig, RenderPrecision,
                                        StereoRenderTarget, CameraParams, StereoCameraParams)

NEAR, FAR = 0.1, 100.0


def camera(W, H, tx=0.0):
    proj = syn.make_projection_matrix(W, H, NEAR, FAR)
    fx, fy = syn.focal_lengths(W, H)
    v = np.eye(4, dtype=np.float32); v[3, 0] = tx
    return CameraParams(v, proj, (-tx, 0, 0), fx, fy, NEAR, FAR)


def timed(fn, steps, warm, dev, world):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    small = "--small" in sys.argv
    out = {"n_gpus": world}
    s = torch.cuda.current_stream()
    # ---- C3
    N, W, H = (600_000, 1920, 1080) if small else (6_000_000, 3840, 2160)
    cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=0.008 if not small else 0.015)
    g, h = cl.pack("float16")
    cam = camera(W, H)
    cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=local, config=cfg)
    a, c = mg.partition_range(N, world)[rank]
    tg = torch.from_numpy(np.ascontiguousarray(g).view(np.uint8).reshape(-1)[a * 32:(a + c) * 32]).to(dev)
    th = torch.from_numpy(np.ascontiguousarray(h).view(np.uint8).reshape(-1)[a * 96:(a + c) * 96]).to(dev)
    scratch = torch.zeros(max(c, 1) * mg.RECORD_BYTES, dtype=torch.uint8, device=dev)
    img = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    tilesY = (H + 15) // 16
    strips = mg.partition_tile_rows(tilesY, world)
    info = {}

    def frame_c3():
        _, counts, strip = mg.render_strips(r, dist, rank, world, s, tg, th, (a, c), 16, cam, W, H, img, None, scratch)
        mg.gather_strips(dist, rank, world, img, strips, W, H, root=0)
        info["records"] = int(sum(counts))
    ms = timed(frame_c3, 5 if not small else 3, 2, dev, world)
    hd = r.debugReadHeader()
    ov = torch.tensor([hd.overflow, hd.totalInstances], dtype=torch.int64, device=dev)
    dist.all_reduce(ov, op=dist.ReduceOp.MAX)
    out["C3"] = {"workload": f"{N} Gaussians SH3 f16, {W}x{H}, strips over {world} GPU(s), all-gather of {info['records']} records x 48 B",
                 "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "allgather_bytes": info["records"] * 48,
                 "max_strip_instances": int(ov[1].item()), "any_strip_overflow": int(ov[0].item())}
    del r, tg, th, scratch, img
    torch.cuda.empty_cache()
    # ---- C4
    if world >= 2:
        N2, W2, H2 = (200_000, 960, 540) if small else (1_000_000, 1920, 1080)
        cl2 = syn.synthetic_cloud(N2, 3, seed=42, scale_median=0.015)
        g2, h2 = cl2.pack("float16")
        cfg2 = RendererConfig(maxGaussians=N2, maxWidth=W2, maxHeight=H2, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
        rs = DepthFirstRenderer(device=local, config=cfg2)
        tg2 = torch.from_numpy(np.ascontiguousarray(g2).view(np.uint8).reshape(-1)).to(dev)
        th2 = torch.from_numpy(np.ascontiguousarray(h2).view(np.uint8).reshape(-1)).to(dev)
        cams = StereoCameraParams(camera(W2, H2, 0.032), camera(W2, H2, -0.032))
        tgt = torch.zeros((H2, 2 * W2, 4), dtype=torch.int16, device=dev)
        inp = GaussianInput(tg2, th2, N2, 16)

        def frame_c4_split():
            mg.render_stereo_split(rs, dist, rank, world, s, tgt, inp, cams, W2, H2)

        def frame_c4_joint():
            rs.renderStereo(s, StereoRenderTarget.sideBySide(tgt), inp, cams, W2, H2)
        ms_split = timed(frame_c4_split, 10, 3, dev, world)
        ms_joint = timed(frame_c4_joint, 10, 3, dev, world)
        out["C4"] = {"workload": f"stereo 2x({W2}x{H2}), {N2} Gaussians SH3 f16", "ms_one_eye_per_gpu": ms_split,
                     "ms_joint_on_one_gpu": ms_joint, "frames_per_s_split": 1e3 / ms_split, "frames_per_s_joint": 1e3 / ms_joint}
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
