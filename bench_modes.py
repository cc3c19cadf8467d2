"""bench_modes.py -- the three multi-GPU shard modes of SURVEY.md 8(e), measured inside bench.py's own run so that the
driver's per-N lines carry them:

  c3_strips  BASELINE config 3: ONE 3840x2160 frame of a 6 M-Gaussian SH3 cloud; Gaussians shard by gid range, the routing
             kernel stores each projected splat into the window of the rank(s) whose strip it touches (gsm_group, NVLink peer
             stores, no library collective, no host in the loop), every rank sorts + blends its strip of tile rows straight into
             rank 0's image. Checked in-run: the assembled image equals rank 0's own single-GPU frame byte for byte. The round-1
             form (padded NCCL all-gather of all records + host-side counts) is timed beside it as the library baseline.
  c5_views   BASELINE config 5: 256 seeded orbit poses of a 3 M-Gaussian SH3 cloud at 1280x720, views split round-robin over
             the ranks, scene replicated, NO collective.
  c4_eyes    BASELINE config 4: stereo 2 x (1920x1080) at 1 M Gaussians, identical joint stages 1-7 on ranks 0 and 1, one eye
             blended per GPU, the right eye written straight into rank 0's side-by-side target over NVLink.

Timing: CUDA events on each rank's stream around K back-to-back frames after warm-up, max over ranks (all_reduce MAX);
inputs resident in HBM. Nothing here runs at import; bench.py calls run_modes()."""
from __future__ import annotations

import math
import time

import numpy as np

NEAR, FAR = 0.1, 100.0


def _camera(CameraParams, syn, W, H, view=None, pos=(0.0, 0.0, 0.0)):
    proj = syn.make_projection_matrix(W, H, NEAR, FAR)
    fx, fy = syn.focal_lengths(W, H)
    return CameraParams(np.eye(4, dtype=np.float32) if view is None else view, proj, pos, fx, fy, NEAR, FAR)


def orbit_poses(syn, n_views: int, seed: int = 42):
    """n_views camera poses on a seeded orbit around the cloud's centre (0, 0, 11), radius 11 (pose 0 = the origin camera the
    cloud was built for), with a seeded vertical wobble; every pose looks at the centre."""
    rng = np.random.default_rng(seed)
    phase, amp = rng.uniform(0, 2 * math.pi), rng.uniform(0.5, 1.5)
    centre = np.array([0.0, 0.0, 11.0])
    poses = []
    for k in range(n_views):
        phi = 2 * math.pi * k / n_views
        eye = centre + np.array([11.0 * math.sin(phi), amp * math.sin(2 * phi + phase), -11.0 * math.cos(phi)])
        poses.append((syn.look_at_opencv(eye, centre), eye.astype(np.float32)))
    return poses


def _timed(torch, dist, dev, world, fn, steps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _upload(torch, arr, dev):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(dev)


def mode_c3_strips(torch, dist, rank, world, local, small=False):
    from gsm_renderer_b200 import multigpu as mg, synthetic as syn
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision)
    dev = torch.device("cuda", local)
    N, W, H, sm = (600_000, 1920, 1080, 0.015) if small else (6_000_000, 3840, 2160, 0.008)
    t0 = time.perf_counter()
    cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=sm)
    g, h = cl.pack("float16")
    gen_s = time.perf_counter() - t0
    cam = _camera(CameraParams, syn, W, H)
    cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=local, config=cfg)
    s = torch.cuda.current_stream()
    shards = mg.partition_range(N, world)
    a, c = shards[rank]
    gb, hb = np.ascontiguousarray(g).view(np.uint8).reshape(-1), np.ascontiguousarray(h).view(np.uint8).reshape(-1)
    out = {"workload": f"C3: {N} Gaussians SH3 float16, {W}x{H}, one frame split into {world} strip(s) of tile rows", "n_gpus": world}
    tilesY = (H + 15) // 16
    strips = mg.partition_tile_rows(tilesY, world)
    rows = mg.strip_row_starts(strips)
    # ---- rank 0: the single-GPU frame (reference image and the 1-GPU time the efficiency is relative to)
    ms_single, ref_c, ref_d = None, None, None
    if rank == 0:
        tg, th = _upload(torch, gb, dev), _upload(torch, hb, dev)
        ref_c = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
        ref_d = torch.zeros((H, W), dtype=torch.int16, device=dev)
        inp = GaussianInput(tg, th, N, 16)
        for _ in range(2):
            r.render(s, ref_c, ref_d, inp, cam, W, H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r.render(s, ref_c, ref_d, inp, cam, W, H)
        e1.record()
        torch.cuda.synchronize()
        ms_single = e0.elapsed_time(e1) / 5
        hd = r.debugReadHeader()
        out.update({"ms_single_gpu": ms_single, "V": hd.visibleCount, "I": hd.totalInstances, "overflow_single_gpu": hd.overflow})
        shard_g, shard_h = tg[a * 32:(a + c) * 32], th[a * 96:(a + c) * 96]
    else:
        shard_g, shard_h = _upload(torch, gb[a * 32:(a + c) * 32], dev), _upload(torch, hb[a * 96:(a + c) * 96], dev)
    if world == 1:
        out.update({"ms_per_frame": ms_single, "frames_per_s": 1e3 / ms_single, "equals_single_gpu": True, "exchange": "none (one GPU)"})
        r.close()
        return out
    # ---- peer-memory path
    cap = max(cc for _, cc in shards)
    grp = mg.RendererGroup(r, rank, world, cap, W * H * 8, W * H * 2)
    grp.connect_distributed(dist)
    pc, pd = grp.image_ptrs(0)
    frame = [0]

    def frame_group():
        frame[0] += 1
        grp.renderStrips(s, pc, pd, shard_g, shard_h, a, c, 16, cam, W, H, rows)
        grp.signal(s, 0, frame[0])
        if rank == 0:
            grp.wait(s, (1 << world) - 1, frame[0])

    ms = _timed(torch, dist, dev, world, frame_group, 10 if not small else 4, 3)
    # phase split on this rank: [project + compaction + routing (the exchange is its stores)] | [wait + ingest + sort + blend]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    dist.barrier()
    frame[0] += 1
    ev[0].record()
    grp.projectRoute(s, shard_g, shard_h, a, c, 16, cam, W, H, rows)
    ev[1].record()
    grp.renderStrip(s, pc, pd, W, H, rows)
    ev[2].record()
    grp.signal(s, 0, frame[0])
    if rank == 0:
        grp.wait(s, (1 << world) - 1, frame[0])
    torch.cuda.synchronize()
    counts = grp.recordCounts(s)
    hd = r.debugReadHeader()
    stats = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), float(sum(counts)), float(sum(counts) - counts[rank]),
                          float(hd.totalInstances), float(hd.overflow)], dtype=torch.float64, device=dev)
    allstats = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)
    equal = None
    if rank == 0:
        img_c, img_d = grp.image_tensors(0, W, H, dev)
        equal = bool(torch.equal(img_c, ref_c)) and bool(torch.equal(img_d, ref_d))
    # ---- library baseline: round 1's padded NCCL all-gather of every record to every rank + host-side counts
    scratch = torch.zeros(max(c, 1) * mg.RECORD_BYTES, dtype=torch.uint8, device=dev)
    img = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)

    def frame_allgather():
        mg.render_strips(r, dist, rank, world, s, shard_g, shard_h, (a, c), 16, cam, W, H, img, None, scratch, strips=strips)
        mg.gather_strips(dist, rank, world, img, strips, W, H, root=0)

    ms_ag = _timed(torch, dist, dev, world, frame_allgather, 5 if not small else 2, 2)
    if rank == 0:
        per = [[float(v) for v in t.tolist()] for t in allstats]
        recv = [int(p[2]) for p in per]
        out.update({
            "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "efficiency_vs_single_gpu": ms_single / (world * ms), "speedup_vs_single_gpu": ms_single / ms,
            "equals_single_gpu": equal,
            "exchange": "routing kernel's own peer stores (NVLink), order-preserving per destination; counts + flags via mailboxes; no host sync",
            "records_received_per_rank": recv, "records_received_remote_per_rank": [int(p[3]) for p in per],
            "exchange_bytes_total": int(sum(p[3] for p in per)) * mg.RECORD_BYTES,
            "phase_ms_per_rank": {"project_compact_route": [p[0] for p in per], "wait_ingest_sort_blend": [p[1] for p in per]},
            "strip_instances_per_rank": [int(p[4]) for p in per], "strip_instances_max": int(max(p[4] for p in per)),
            "any_strip_overflow": int(max(p[5] for p in per)),
            "ms_per_frame_nccl_allgather_baseline": ms_ag,
            "limiter": _c3_limiter(per, ms, ms_single, world),
            "cloud_generation_s": gen_s,
        })
    grp.close()
    r.close()
    return out


def _c3_limiter(per, ms, ms_single, world):
    route = max(p[0] for p in per)
    tail = max(p[1] for p in per)
    inst = [p[4] for p in per]
    imb = max(inst) / (sum(inst) / len(inst)) if sum(inst) else 1.0
    return (f"slowest rank: {route:.3f} ms project+route (projection shards 1/{world}, routing is latency + NVLink stores) + {tail:.3f} ms "
            f"strip tail; instance imbalance max/mean {imb:.2f} (strips are balanced by rows, not by work)")


def mode_c5_views(torch, dist, rank, world, local, n_views=256, small=False):
    from gsm_renderer_b200 import multigpu as mg, synthetic as syn
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision)
    dev = torch.device("cuda", local)
    N, W, H, sm = (300_000, 1280, 720, 0.02) if small else (3_000_000, 1280, 720, 0.012)
    cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=sm)
    g, h = cl.pack("float16")
    tg, th = _upload(torch, g, dev), _upload(torch, h, dev)
    cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=local, config=cfg)
    s = torch.cuda.current_stream()
    inp = GaussianInput(tg, th, N, 16)
    poses = orbit_poses(syn, n_views)
    mine = mg.partition_views(n_views, world, rank)
    cams = [_camera(CameraParams, syn, W, H, poses[v][0], poses[v][1]) for v in mine]
    color = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    depth = torch.zeros((H, W), dtype=torch.int16, device=dev)
    for cam in cams[:3]:
        r.render(s, color, depth, inp, cam, W, H)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for cam in cams:
        r.render(s, color, depth, inp, cam, W, H)
    e1.record()
    torch.cuda.synchronize()
    ms_rank = e0.elapsed_time(e1)
    # per-view counts of a few poses (overflow means the nearest 4*N instances were kept, as the reference would)
    over, maxI = 0, 0
    for cam in cams[:: max(1, len(cams) // 8)]:
        r.render(s, color, depth, inp, cam, W, H)
        torch.cuda.synchronize()
        hd = r.debugReadHeader()
        over += int(hd.overflow)
        maxI = max(maxI, int(hd.totalInstances))
    t = torch.tensor([ms_rank, float(over), float(maxI)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_all = float(t[0].item())
    r.close()
    return {"workload": f"C5: {n_views} orbit poses of {N} Gaussians SH3 float16 at {W}x{H}, {len(mine)} views on rank 0, scene replicated, no collective",
            "n_gpus": world, "views": n_views, "ms_total_max_over_ranks": ms_all, "frames_per_s": n_views / (ms_all * 1e-3),
            "ms_per_view_per_gpu": ms_all / max(1, len(mg.partition_views(n_views, world, 0))),
            "sampled_views_overflowing_any_rank": int(t[1].item()), "sampled_max_instances": int(t[2].item())}


def mode_c4_eyes(torch, dist, rank, world, local, small=False):
    from gsm_renderer_b200 import multigpu as mg, synthetic as syn
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision, StereoCameraParams, StereoRenderTarget)
    dev = torch.device("cuda", local)
    N, W, H = (200_000, 960, 540) if small else (1_000_000, 1920, 1080)
    maxG = 6_000_000  # the reference's default capacity (GRP.swift:211-218): the union boxes of this cloud need > 4 * N instances
    cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=0.015)
    g, h = cl.pack("float16")
    tg, th = _upload(torch, g, dev), _upload(torch, h, dev)
    cfg = RendererConfig(maxGaussians=maxG, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=local, config=cfg)
    s = torch.cuda.current_stream()
    inp = GaussianInput(tg, th, N, 16)
    lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    lv[3, 0], rv[3, 0] = 0.032, -0.032
    cams = StereoCameraParams(_camera(CameraParams, syn, W, H, lv, (-0.032, 0, 0)), _camera(CameraParams, syn, W, H, rv, (0.032, 0, 0)))
    joint = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)

    def frame_joint():
        r.renderStereo(s, StereoRenderTarget.sideBySide(joint), inp, cams, W, H)

    ms_joint = _timed(torch, dist, dev, world, frame_joint, 10, 3)
    hd = r.debugReadHeader()
    out = {"workload": f"C4: stereo 2x({W}x{H}), {N} Gaussians SH3 float16, renderStereo side-by-side, maxGaussians {maxG}", "n_gpus": world,
           "ms_joint_on_one_gpu": ms_joint, "frames_per_s_joint": 1e3 / ms_joint, "V": hd.visibleCount, "I": hd.totalInstances, "overflow": hd.overflow}
    if world >= 2:
        grp = mg.RendererGroup(r, rank, world, 1, H * 2 * W * 8, 0)
        grp.connect_distributed(dist)
        pc, _ = grp.image_ptrs(0)
        frame = [0]

        def frame_split():
            frame[0] += 1
            if rank < 2:
                r.renderStereo(s, StereoRenderTarget.sideBySide(pc), inp, cams, W, H, eyeMask=1 << rank)
                grp.signal(s, 0, frame[0])
            if rank == 0:
                grp.wait(s, 0b11, frame[0])

        ms_split = _timed(torch, dist, dev, world, frame_split, 10, 3)
        equal = None
        if rank == 0:
            img, _ = grp.image_tensors(0, 2 * W, H, dev, depth=False)
            equal = bool(torch.equal(img, joint))
        out.update({"ms_one_eye_per_gpu": ms_split, "frames_per_s_split": 1e3 / ms_split, "equals_joint_frame": equal,
                    "transfer": "rank 1 blends the right eye straight into rank 0's side-by-side target (NVLink peer stores), one flag"})
        dist.barrier()
        grp.close()
    r.close()
    return out


def run_modes(torch, dist, rank, world, local, small=False):
    """All three modes; returns the dict for the bench line's `modes` key (complete on rank 0)."""
    modes = {}
    for name, fn in (("c3_strips", mode_c3_strips), ("c5_views", mode_c5_views), ("c4_eyes", mode_c4_eyes)):
        t0 = time.perf_counter()
        try:
            modes[name] = fn(torch, dist, rank, world, local, small=small)
        except Exception as e:  # a failing mode must not take the headline line down with it; it is reported, not hidden
            import traceback
            modes[name] = {"error": f"{type(e).__name__}: {e}", "traceback": traceback.format_exc()[-1500:]}
        modes[name]["wall_s"] = time.perf_counter() - t0
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        if world > 1:
            dist.barrier()
    return modes


if __name__ == "__main__":   # diagnostic entry: torchrun ... bench_modes.py [--small] [mode ...]
    import json
    import os
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3_strips", "c5_views", "c4_eyes"]
    fns = {"c3_strips": mode_c3_strips, "c5_views": mode_c5_views, "c4_eyes": mode_c4_eyes}
    for n in names:
        res = fns[n](torch, dist if world > 1 else None, rank, world, local, small="--small" in sys.argv)
        if rank == 0:
            print(json.dumps({n: res}))
    if world > 1:
        dist.destroy_process_group()
