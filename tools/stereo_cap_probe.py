"""Stage times of the C4 stereo frame at several renderer capacities (diagnostic)."""
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
s = torch.cuda.current_stream()
for maxG in [int(x) for x in sys.argv[1:]] or [1_000_000, 1_400_000, 2_000_000, 6_000_000]:
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=maxG, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear))
    tgt = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    inp = GaussianInput(tg, th, N, 16)
    for _ in range(2):
        r.renderStereo(s, StereoRenderTarget.sideBySide(tgt), inp, cams, W, H)
    torch.cuda.synchronize()
    r.setProfiling(True)
    r.renderStereo(s, StereoRenderTarget.sideBySide(tgt), inp, cams, W, H)
    torch.cuda.synchronize()
    st = r.stageTimesMs()
    hd = r.debugReadHeader()
    T = 120 * 68
    mx = int(r.debugReadTileHeaders(T)[:, 1].max())
    print(json.dumps({"maxG": maxG, "V": hd.visibleCount, "I": hd.totalInstances, "overflow": hd.overflow, "maxPerTile": mx, "stage_ms": st}))
    r.close()
