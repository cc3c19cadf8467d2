"""Per-tile timeline of create_instances_kernel on the bench workload (diagnostic).
Build the instrumented library, then run on the GPU box:
  nvcc <flags of __graft_entry__.NVCC_FLAGS> -DGSM_EXPAND_TRACE -o tools/bin/libgsm_trace.so gsm_renderer_b200/csrc/*.cu
  GSM_B200_LIB=tools/bin/libgsm_trace.so python tools/expand_trace.py > gpurun_out/expand_trace.csv"""
import ctypes, os, sys
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
tiles = (Ng + 255) // 256
trace = torch.zeros(tiles * 8, dtype=torch.int64, device="cuda")
for _ in range(5):
    r.render(s, color, depth, inp, cam, W, H)
torch.cuda.synchronize()
assert N.lib().gsm_trace_expand_set(ctypes.c_void_p(trace.data_ptr())) == 0
r.render(s, color, depth, inp, cam, W, H)
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(-1, 8)
t = t[t[:, 1] != 0]
t0 = t[:, 0].min()
for i, row in enumerate(t):
    print(i, *(int(x - t0) for x in row[:7]), int(row[7]), sep=",")
