"""C2 frame time as a function of the renderer's CAPACITY (maxGaussians): the per-frame state a frame clears and the grids it
launches are sized from the capacity, the scene is the same 1 M Gaussians. Usage: python tools/capacity_probe.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200 import synthetic as syn  # noqa: E402
from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,  # noqa: E402
                                        RenderPrecision)

NEAR, FAR = 0.1, 100.0
N, W, H = 1_000_000, 1920, 1080
dev = torch.device("cuda", 0)
cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=0.015)
g, h = cl.pack("float16")
tg = torch.from_numpy(np.ascontiguousarray(g).view(np.uint8).reshape(-1)).to(dev)
th = torch.from_numpy(np.ascontiguousarray(h).view(np.uint8).reshape(-1)).to(dev)
proj = syn.make_projection_matrix(W, H, NEAR, FAR)
fx, fy = syn.focal_lengths(W, H)
cam = CameraParams(np.eye(4, dtype=np.float32), proj, (0, 0, 0), fx, fy, NEAR, FAR)
inp = GaussianInput(tg, th, N, 16)
s = torch.cuda.current_stream(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
color = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
depth = torch.zeros((H, W), dtype=torch.int16, device=dev)
out = {}
for cap in (1_000_000, 2_000_000, 6_000_000, 30_000_000):
    cfg = RendererConfig(maxGaussians=cap, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=0, config=cfg)
    ms = []
    for i in range(25):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        r.render(s, color, depth, inp, cam, W, H)
        b.record(s)
        b.synchronize()
        if i >= 5:
            ms.append(a.elapsed_time(b))
    r.setProfiling(True)
    r.render(s, color, depth, inp, cam, W, H)
    torch.cuda.synchronize()
    out[str(cap)] = {"ms_median": float(np.median(ms)), "stages_us": {k: round(v * 1e3, 1) for k, v in r.stageTimesMs().items()}}
    r.close()
print(json.dumps(out, indent=1))
