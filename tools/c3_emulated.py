"""The C3 strip-sharded frame with WORLD emulated ranks on ONE GPU (group windows connected locally): run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel times of one rank's phases (peer stores land in local HBM here).
  python tools/c3_emulated.py [world] [frames]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200 import multigpu as mg, synthetic as syn
from gsm_renderer_b200.renderer import CameraParams, DepthFirstRenderer, GaussianColorSpace, RendererConfig, RenderPrecision
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N, W, H = 6_000_000, 3840, 2160
cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=0.008)
g, h = cl.pack("float16")
dev = torch.device("cuda:0")
tg = torch.from_numpy(np.ascontiguousarray(g).view(np.uint8).reshape(-1)).to(dev)
th = torch.from_numpy(np.ascontiguousarray(h).view(np.uint8).reshape(-1)).to(dev)
proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
fx, fy = syn.focal_lengths(W, H)
cam = CameraParams(np.eye(4, dtype=np.float32), proj, (0, 0, 0), fx, fy, 0.1, 100.0)
cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
shards = mg.partition_range(N, world)
cap = max(c for _, c in shards)
rs = [DepthFirstRenderer(device=0, config=cfg) for _ in range(world)]
groups = [mg.RendererGroup(rs[k], k, world, cap, W * H * 8 if k == 0 else 0, W * H * 2 if k == 0 else 0) for k in range(world)]
for gr in groups:
    gr.connect_local(groups)
rows = mg.strip_row_starts(mg.partition_tile_rows((H + 15) // 16, world))
pc, pd = groups[0].image_ptrs(0)
s = torch.cuda.current_stream()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(world)]
for f in range(1, frames + 1):
    for k, (a, c) in enumerate(shards):
        ev[k][0].record()
        groups[k].projectRoute(s, tg[a * 32:(a + c) * 32], th[a * 96:(a + c) * 96], a, c, 16, cam, W, H, rows)
        ev[k][1].record()
    for k in range(world):
        e = torch.cuda.Event(enable_timing=True); e.record()
        groups[k].renderStrip(s, pc, pd, W, H, rows)
        ev[k][2].record()
        ev[k][1] = (ev[k][1], e)
        groups[k].signal(s, 0, f)
    groups[0].wait(s, (1 << world) - 1, f)
    torch.cuda.synchronize()
    print(json.dumps({"frame": f, "project_route_ms": [ev[k][0].elapsed_time(ev[k][1][0]) for k in range(world)],
                      "strip_ms": [ev[k][1][1].elapsed_time(ev[k][2]) for k in range(world)],
                      "records": [sum(groups[k].recordCounts(s)) for k in range(world)]}))
    for k in range(world):
        ev[k] = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
