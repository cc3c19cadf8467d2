"""C4 on one GPU: joint stereo frame time and stage times (BASELINE.json configs[3]). Usage: python tools/stereo_bench.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200 import synthetic as syn  # noqa: E402
from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, EyeView, FoveatedStereoDrawable,  # noqa: E402
                                        GaussianColorSpace, GaussianInput, PixelFormat, RasterizationRateMap, RendererConfig,
                                        RenderPrecision, StereoCameraParams, StereoConfiguration, StereoRenderTarget, Viewport)
from tests import foveation_util as fv  # noqa: E402

NEAR, FAR = 0.1, 100.0


def camera(W, H, tx=0.0):
    proj = syn.make_projection_matrix(W, H, NEAR, FAR)
    fx, fy = syn.focal_lengths(W, H)
    v = np.eye(4, dtype=np.float32)
    v[3, 0] = tx
    return CameraParams(v, proj, (-tx, 0, 0), fx, fy, NEAR, FAR)


def main():
    N, W, H = 1_000_000, 1920, 1080
    dev = torch.device("cuda", 0)
    cl = syn.synthetic_cloud(N, 3, seed=42, scale_median=0.015)
    g, h = cl.pack("float16")
    cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=0, config=cfg)
    tg = torch.from_numpy(np.ascontiguousarray(g).view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(np.ascontiguousarray(h).view(np.uint8).reshape(-1)).to(dev)
    cams = StereoCameraParams(camera(W, H, 0.032), camera(W, H, -0.032))
    tgt = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    inp = GaussianInput(tg, th, N, 16)
    s = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for i in range(13):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        r.renderStereo(s, StereoRenderTarget.sideBySide(tgt), inp, cams, W, H)
        b.record(s)
        b.synchronize()
        if i >= 3:
            ms.append(a.elapsed_time(b))
    out = {"workload": f"stereo 2x({W}x{H}), {N} Gaussians SH3 f16, one GPU", "ms_median": float(np.median(ms)), "ms_min": min(ms)}
    try:
        r.setProfiling(True)
        r.renderStereo(s, StereoRenderTarget.sideBySide(tgt), inp, cams, W, H)
        torch.cuda.synchronize()
        out["stages_ms"] = r.stageTimesMs()
    except Exception as e:  # noqa: BLE001
        out["stages_ms"] = str(e)
    hd = r.debugReadHeader()
    out["instances"] = int(hd.totalInstances)
    # StereoRenderTarget.foveated: layered bgra8Unorm_srgb drawable behind a rate map
    sx, sy = fv.layer(W, H, fv.FOVEATED_H, fv.FOVEATED_V)
    tw, thh = sx.size, sy.size
    dst = torch.zeros((2, thh, tw * 4), dtype=torch.uint8, device=dev)
    d = FoveatedStereoDrawable(dst, tw, thh, 2, RasterizationRateMap([(sx, sy)]), PixelFormat.bgra8Unorm_srgb)
    L, R = cams.leftEye, cams.rightEye
    vp = Viewport(0, 0, W, H)
    cfgs = StereoConfiguration(EyeView(vp, L.viewMatrix, L.projectionMatrix, L.position, L.focalX, L.focalY, L.near, L.far),
                               EyeView(vp, R.viewMatrix, R.projectionMatrix, R.position, R.focalX, R.focalY, R.near, R.far))
    target = StereoRenderTarget.foveated(d, cfgs)
    r.setProfiling(False)
    ms = []
    for i in range(13):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        r.renderStereo(s, target, inp, None, W, H)
        b.record(s)
        b.synchronize()
        if i >= 3:
            ms.append(a.elapsed_time(b))
    inter = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
    cp = []
    for i in range(13):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        r.stereoCopy(s, inter, W, H, d, vp, vp)
        b.record(s)
        b.synchronize()
        if i >= 3:
            cp.append(a.elapsed_time(b))
    copy_bytes = 2 * W * H * 8 + 2 * tw * thh * 4   # every intermediate texel read once + every drawable texel written once
    out["foveated"] = {"drawable": f"2 x {tw}x{thh} bgra8Unorm_srgb behind a {len(fv.FOVEATED_H)}x{len(fv.FOVEATED_V)}-cell rate map",
                       "ms_median": float(np.median(ms)), "copy_ms_median": float(np.median(cp)),
                       "copy_includes": "upload of the rate-map tables + the resampling kernel",
                       "copy_algorithmic_bytes": copy_bytes, "copy_GBps": copy_bytes / (float(np.median(cp)) * 1e-3) / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
