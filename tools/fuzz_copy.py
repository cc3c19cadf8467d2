"""Randomised GPU-vs-oracle sweep of the foveated stereo copy (gsm_stereo_copy): image size, drawable size and layout, pixel
format, row padding, viewports (fractional, overlapping, partly or wholly outside the drawable, magnifying and minifying), rate
maps (0-2 layers, random cell rates), flip. Bit for bit, untouched texels included. Usage: python tools/fuzz_copy.py [cases] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob
from gsm_renderer_b200.renderer import (DepthFirstRenderer, FoveatedStereoDrawable, GaussianColorSpace, PixelFormat,
                                        RasterizationRateMap, RendererConfig, RenderPrecision, Viewport)
from tests import foveation_util as fv

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
dev = torch.device("cuda:0")
fails = 0
t0 = time.time()
for c in range(cases):
    W, H = int(rng.integers(1, 300)), int(rng.integers(1, 200))
    fmt = int(rng.integers(0, 5))
    flip = bool(rng.integers(0, 2))
    array_length = int(rng.integers(1, 3))
    screen = (float(rng.uniform(8, 600)), float(rng.uniform(8, 400)))
    nl = int(rng.integers(0, 3))
    layers = None
    if nl:
        layers = [fv.layer(screen[0], screen[1], rng.uniform(0.1, 1.0, int(rng.integers(1, 7))), rng.uniform(0.1, 1.0, int(rng.integers(1, 6))))
                  for _ in range(nl)]
        tw, th = max(l[0].size for l in layers) + int(rng.integers(0, 3)), max(l[1].size for l in layers) + int(rng.integers(0, 3))
    else:
        tw, th = int(np.ceil(screen[0])), int(np.ceil(screen[1]))
    def vp():
        w, h = float(rng.uniform(1, screen[0] * 1.2)), float(rng.uniform(1, screen[1] * 1.2))
        ox, oy = float(rng.uniform(-0.3 * screen[0], screen[0])), float(rng.uniform(-0.3 * screen[1], screen[1]))
        if rng.integers(0, 3) == 0:
            ox, oy, w, h = float(int(ox)), float(int(oy)), float(max(1, int(w))), float(max(1, int(h)))
        return (ox, oy, w, h)
    vps = (vp(), vp())
    pad = int(rng.integers(0, 9))
    px = 8 if fmt == 0 else 4
    row_bytes = (tw + pad) * px
    if fmt == 0:
        c2 = (rng.standard_normal((2, H, W, 4)) * np.exp(rng.uniform(-6, 6, (2, H, W, 4)))).astype(np.float16)
    else:
        c2 = rng.uniform(-0.2, 1.2, (2, H, W, 4)).astype(np.float16)
    if rng.integers(0, 2):
        c2.reshape(-1)[rng.integers(0, c2.size, 4)] = [np.inf, -np.inf, np.nan, -0.0]
    c2 = c2.view(np.uint16)
    ref = ob.stereo_copy_foveated(c2, flip, tw, th, array_length, fmt, vps, rate_layers=layers, row_bytes=row_bytes)
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=16, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear), stereoCopyFlipY=flip)
    src = torch.from_numpy(np.ascontiguousarray(np.concatenate([c2[0], c2[1]], axis=1)).view(np.int16)).to(dev)
    dst = torch.full((array_length, th, row_bytes), 0xAB, dtype=torch.uint8, device=dev)
    d = FoveatedStereoDrawable(dst, tw, th, array_length, RasterizationRateMap(layers) if layers else None, PixelFormat(fmt), row_bytes)
    r.stereoCopy(torch.cuda.current_stream(), src, W, H, d, Viewport(*vps[0]), Viewport(*vps[1]))
    torch.cuda.synchronize()
    got = dst.cpu().numpy()
    r.close()
    ok = np.array_equal(got, ref)
    if not ok:
        fails += 1
        bad = np.argwhere(got != ref)
        print(c, "FAIL", dict(W=W, H=H, fmt=fmt, flip=flip, array_length=array_length, tex=(tw, th), vps=vps, layers=nl), len(bad), bad[0].tolist())
print(f"{cases} cases, {fails} failures, {time.time() - t0:.1f} s")
sys.exit(1 if fails else 0)
