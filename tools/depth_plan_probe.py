"""How the depth sort of a bench workload ran: plan (mode, key range, buckets) and the stage times (diagnostic)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gsm_renderer_b200 import synthetic as syn
from gsm_renderer_b200.renderer import CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig, RenderPrecision

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cloud, g, h, spec = bench.build_workload(name)
N, deg, prec, W, H, desc = spec
dev = torch.device("cuda:0")
r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H,
                       precision=RenderPrecision.float16 if prec == "float16" else RenderPrecision.float32,
                       gaussianColorSpace=GaussianColorSpace.linear))
proj = syn.make_projection_matrix(W, H, bench.NEAR, bench.FAR)
fx, fy = syn.focal_lengths(W, H)
cam = CameraParams(np.eye(4, dtype=np.float32), proj, np.zeros(3, np.float32), fx, fy, bench.NEAR, bench.FAR)
tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
color = torch.zeros((H, W, 4), dtype=torch.float16, device=dev)
depth = torch.zeros((H, W), dtype=torch.float16, device=dev)
inp = GaussianInput(tg, th, N, syn.SH_COEFFS[deg])
s = torch.cuda.current_stream()
for _ in range(5):
    r.render(s, color, depth, inp, cam, W, H)
torch.cuda.synchronize()
hd = r.debugReadHeader()
plan = r.debugReadDepthSortPlan()
keys = r.debugReadDepthKeys(hd.visibleCount)
r.setProfiling(True)
r.render(s, color, depth, inp, cam, W, H)
torch.cuda.synchronize()
print(json.dumps({"workload": name, "V": hd.visibleCount, "I": hd.totalInstances, "plan": plan,
                  "sorted": bool(np.all(keys[1:] >= keys[:-1])), "key_span_bits": int(keys[-1] - keys[0]).bit_length() if len(keys) else 0,
                  "stage_ms": r.stageTimesMs()}))
