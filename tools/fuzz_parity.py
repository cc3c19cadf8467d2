"""Randomised parity sweep on the GPU box: random clouds (sizes, scales, SH degree, precision, colour space), random camera
poses (inside / outside / behind the cloud, near-plane crossings), odd resolutions and the precision knobs, every white-box
buffer and every pixel compared bit for bit with the CPU oracle. `python tools/fuzz_parity.py [cases] [seed]`"""
import math, os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob
from gsm_renderer_b200 import synthetic as syn
import tests.parity_util as pu

ob.build()


def run_stereo_case(cl, prec, W, H, near, far, view, pos, flip, half_ipd):
    """The joint stereo frame (both eyes from one sort) against the oracle: white-box buffers + the side-by-side image."""
    import torch
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision, StereoCameraParams, StereoRenderTarget)
    g, h = pu.make_scene_inputs(cl, prec)
    proj = syn.make_projection_matrix(W, H, near, far)
    fx, fy = syn.focal_lengths(W, H)
    base = np.eye(4, dtype=np.float32) if view is None else np.asarray(view, np.float32)
    lv, rv = base.copy(), base.copy()
    lv[3, 0] += half_ipd; rv[3, 0] -= half_ipd      # eyes offset along the camera's x axis (numpy row 3 = matrix column 3)
    p = np.asarray(pos, np.float32)
    cams = StereoCameraParams(CameraParams(lv, proj, tuple(p), fx, fy, near, far), CameraParams(rv, proj, tuple(p), fx, fy, near, far))
    ocam = ob.make_stereo_camera(lv, proj, tuple(p), rv, proj, tuple(p), W, H, near, far, cl.sh_components, cl.count, False)
    fr = ob.OracleFrame(cl.count, W, H, stereo=True)
    ref, _ = fr.render_stereo(g, h, ob.F16 if prec == "float16" else ob.F32, ocam, W, H, flip_y=flip)
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=cl.count, maxWidth=W, maxHeight=H,
                           precision=RenderPrecision.float16 if prec == "float16" else RenderPrecision.float32,
                           gaussianColorSpace=GaussianColorSpace.linear), stereoCopyFlipY=flip)
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev); th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    sbs = torch.full((H, 2 * W, 4), 0x7E00, dtype=torch.int16, device=dev)
    r.renderStereo(torch.cuda.current_stream(), StereoRenderTarget.sideBySide(sbs), GaussianInput(tg, th, cl.count, cl.sh_components), cams, W, H)
    torch.cuda.synchronize()
    try:
        V, I = pu.compare_white_box(r, fr, W, H, cl.count, stereo=True)
        pu.compare_pixels(sbs.cpu().numpy().view(np.uint16), ref, True, prec == "float16", "stereo colour")
    finally:
        r.close()
    return dict(V=V, I=I, activeTiles=int(fr.f.activeTileCount))


cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
fails = 0
t0 = time.time()
for c in range(cases):
    n = int(rng.choice([1, 37, 1000, 5000, 20000, 70000, 150000]))
    deg = int(rng.integers(0, 4))
    prec = str(rng.choice(["float16", "float32"]))
    srgb = bool(rng.integers(0, 2))
    W, H = [(1920, 1080), (1280, 720), (640, 360), (333, 77), (1919, 1079), (64, 64), (2048, 16)][int(rng.integers(0, 7))]
    scale = float(np.exp(rng.uniform(math.log(0.003), math.log(0.3))))
    near = float(rng.choice([0.01, 0.1, 1.0, 5.0])); far = float(rng.choice([20.0, 100.0, 1000.0]))
    cl = syn.synthetic_cloud(n, deg, seed=int(rng.integers(0, 1 << 30)), scale_median=scale, scale_sigma=float(rng.uniform(0.2, 1.2)),
                             z_range=(float(rng.uniform(0.05, 3.0)), float(rng.uniform(4.0, 60.0))))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        view, pos = None, (0.0, 0.0, 0.0)
    else:
        eye = {1: rng.normal(0, 4, 3) + np.array([0, 0, 11.0]),          # inside the cloud
               2: np.array([0, 0, 11.0]) + 25.0 * rng.normal(0, 1, 3) / 1.7,  # far outside
               3: np.array([rng.normal(0, 1), rng.normal(0, 1), 30.0])}[kind]      # behind, looking back
        target = np.array([0, 0, 11.0]) + rng.normal(0, 2, 3)
        view, pos = syn.look_at_opencv(eye, target), tuple(np.asarray(eye, np.float32))
    dk16 = bool(rng.integers(0, 4) == 0); t16 = bool(rng.integers(0, 4) != 0)
    desc = dict(n=n, deg=deg, prec=prec, srgb=srgb, W=W, H=H, scale=round(scale, 4), near=near, far=far, cam=kind, depthKey16=dk16, tileId16=t16)
    stereo = bool(rng.integers(0, 5) == 0) and W * 2 <= 4096
    desc["stereo"] = stereo
    try:
        if stereo:
            res = run_stereo_case(cl, prec, W, H, near, far, view, pos, bool(rng.integers(0, 2)), float(rng.choice([0.0, 0.032, 0.2])))
        else:
            res = pu.run_mono_case(ob, cl, prec, W, H, near, far, srgb=srgb, depth_key16=dk16, tile_id16=t16, view=view, position=pos)
        print(c, "ok", desc, {k: res[k] for k in ("V", "I", "activeTiles")}, flush=True)
    except Exception as e:  # noqa: BLE001
        fails += 1
        print(c, "FAIL", desc, repr(e)[:300], flush=True)
        traceback.print_exc(limit=2)
print(f"{cases} cases, {fails} failures, {time.time() - t0:.1f} s")
sys.exit(1 if fails else 0)
