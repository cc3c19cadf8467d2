"""GlobalRenderer on the bench workload: C2 cloud (1 M Gaussians, SH3, float16) at 1920x1080 on one GPU,
frame time from CUDA events with an L2 flush between frames, beside the DepthFirst frame of the same scene.
Usage: python tools/global_bench.py > profiles/r2_global_bench.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200 import synthetic as syn  # noqa: E402
from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, GlobalRenderer,  # noqa: E402
                                        RendererConfig, RenderPrecision)

NEAR, FAR = 0.1, 100.0


def main():
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
    out = {"workload": f"C2 cloud through both renderers: {N} Gaussians SH3 f16, {W}x{H}, one GPU, L2 flushed between frames"}
    for name, cls in (("global", GlobalRenderer), ("depthFirst", DepthFirstRenderer)):
        cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
        r = cls(device=0, config=cfg)
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
        out[name] = {"ms_median": float(np.median(ms)), "ms_min": float(min(ms)), "frames_per_s": 1e3 / float(np.median(ms))}
        hd = r.debugReadHeader()
        if name == "global":
            out[name].update(visible=int(hd["visibleCount"]), assignments=int(hd["totalAssignments"]), activeTiles=int(hd["activeTileCount"]),
                             tiles="32x16 px")
        r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
