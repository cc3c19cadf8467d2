"""Back-to-back (unsynchronised) stereo frames at several renderer capacities; run with GSM_PDL=0/1 (diagnostic)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200 import synthetic as syn
from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                        RenderPrecision, StereoCameraParams, StereoRenderTarget)
N, W, H = 1_000_000, 1920, 1080
cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=0.015)
g, h = cl.pack("float16")
dev = torch.device("cuda:0")
tg = torch.from_numpy(np.ascontiguousarray(g).view(np.uint8).reshape(-1)).to(dev)
th = torch.from_numpy(np.ascontiguousarray(h).view(np.uint8).reshape(-1)).to(dev)
proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
fx, fy = syn.focal_lengths(W, H)
lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
lv[3, 0], rv[3, 0] = 0.032, -0.032
cams = StereoCameraParams(CameraParams(lv, proj, (-0.032, 0, 0), fx, fy, 0.1, 100.0), CameraParams(rv, proj, (0.032, 0, 0), fx, fy, 0.1, 100.0))
mono = CameraParams(np.eye(4, dtype=np.float32), proj, (0, 0, 0), fx, fy, 0.1, 100.0)
s = torch.cuda.current_stream()
for maxG in [int(x) for x in sys.argv[1:]] or [1_000_000, 1_400_000, 6_000_000]:
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=maxG, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear))
    tgt = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    col = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    inp = GaussianInput(tg, th, N, 16)
    res = {"maxG": maxG, "pdl": os.environ.get("GSM_PDL", "1")}
    for name, fn in (("stereo", lambda: r.renderStereo(s, StereoRenderTarget.sideBySide(tgt), inp, cams, W, H)),
                     ("mono", lambda: r.render(s, col, None, inp, mono, W, H))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_b2b_ms"] = e0.elapsed_time(e1) / 6
        e0.record()
        for _ in range(6):
            fn()
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_synced_ms"] = e0.elapsed_time(e1) / 6
    print(json.dumps(res))
    r.close()
