"""How much of the mono blend's inner loop is spent on splats that no lane of a warp can see (diagnostic).
Build the instrumented library, then run on the GPU box:
  nvcc <flags of __graft_entry__.NVCC_FLAGS> -DGSM_BLEND_STATS -o tools/bin/libgsm_stats.so gsm_renderer_b200/csrc/*.cu
  GSM_B200_LIB=tools/bin/libgsm_stats.so python tools/blend_stats.py"""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gsm_renderer_b200 import synthetic as syn
from gsm_renderer_b200 import _native as N
from gsm_renderer_b200.renderer import (DepthFirstRenderer, RendererConfig, GaussianInput, CameraParams, RenderPrecision,
                                        GaussianColorSpace)

cloud, g, h, spec = bench.build_workload("C2")
Ng, deg, prec, W, H, _ = spec
K = syn.SH_COEFFS[deg]
r = DepthFirstRenderer(device=0, config=RendererConfig(
    maxGaussians=Ng, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16 if prec == "float16" else RenderPrecision.float32,
    gaussianColorSpace=GaussianColorSpace.linear))
tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).cuda(); th = torch.from_numpy(h.view(np.uint8).reshape(-1)).cuda()
color = torch.zeros((H, W, 4), dtype=torch.float16, device="cuda"); depth = torch.zeros((H, W), dtype=torch.float16, device="cuda")
proj = syn.make_projection_matrix(W, H, bench.NEAR, bench.FAR); fx, fy = syn.focal_lengths(W, H)
cam = CameraParams(np.eye(4, dtype=np.float32), proj, np.zeros(3, np.float32), fx, fy, bench.NEAR, bench.FAR)
inp = GaussianInput(tg, th, Ng, K)
s = torch.cuda.current_stream()
out = (ctypes.c_ulonglong * 4)()
lib = ctypes.CDLL(N.LIB_PATH if not os.environ.get("GSM_B200_LIB") else os.environ["GSM_B200_LIB"])
r.render(s, color, depth, inp, cam, W, H)
torch.cuda.synchronize()
assert lib.gsm_blend_stats_read(out, 1) == 0
r.render(s, color, depth, inp, cam, W, H)
torch.cuda.synchronize()
assert lib.gsm_blend_stats_read(out, 1) == 0
wi, wfar, li, lfull = (int(x) for x in out)
print(json.dumps({"warp_evaluations": wi, "no_lane_uses_splat": wfar, "no_lane_uses_frac": wfar / wi, "lane_evaluations": li,
                  "lanes_using_splat": lfull, "lane_use_frac": lfull / li,
                  "instances_in_lists": None}))
