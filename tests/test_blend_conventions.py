"""How far apart are the blend's arithmetic conventions? (parity item 1.)

The canonical blend contracts `a*b + c` into one fused half FMA because the Metal library is built
-ffast-math; SURVEY.md H2 first prescribed separately rounded half mul/add. Nothing in the reference pins either, so this
file QUANTIFIES the choice instead of asserting it:
  * contracted vs uncontracted oracle frames on the same tile lists: max-abs and PSNR per channel;
  * both against an independent float64 "textbook" blend (numpy, written from DFS.metal:1703-1811 without looking at the
    half arithmetic): the half conventions must sit within the rounding noise of a binary16 accumulator from it, and the
    contracted form must not be further from the float64 frame than the uncontracted one by more than noise.
CPU only: oracle + numpy.
"""
import math

import numpy as np
import pytest

from gsm_renderer_b200 import synthetic as syn
from oracle import binding as ob

W, H, N = 640, 368, 60_000
NEAR, FAR = 0.1, 100.0


def _frames():
    ob.build()
    cloud = syn.synthetic_cloud(N, 3, seed=7, scale_median=0.03)
    g, h = cloud.pack("float16")
    proj = syn.make_projection_matrix(W, H, NEAR, FAR)
    cam = ob.make_camera(np.eye(4), proj, (0, 0, 0), W, H, NEAR, FAR, 16, N, False)
    fr = ob.OracleFrame(N, W, H)
    out = {}
    try:
        for name, on in (("contracted", 1), ("uncontracted", 0)):
            ob.lib().gsmo_set_blend_contraction(on)
            c, d = fr.render_mono(g, h, ob.F16, cam, W, H)
            out[name] = (c.view(np.float16).astype(np.float64), d.view(np.float16).astype(np.float64))
    finally:
        ob.lib().gsmo_set_blend_contraction(1)
    return fr, out


def _textbook_tile(fr, tile, tiles_x):
    """float64 blend of one tile's list, quad early exit included (DFS.metal:1746-1747): returns (16,16,4) colour, (16,16) depth."""
    off, cnt = (int(v) for v in fr.tileHeaders[tile])
    tx, ty = tile % tiles_x, tile // tiles_x
    xs = (tx * 16 + np.arange(16)).astype(np.float64)
    ys = (ty * 16 + np.arange(16)).astype(np.float64)
    X, Y = np.meshgrid(xs, ys)
    T = np.ones((16, 16))
    Cc = np.zeros((16, 16, 3))
    D = np.zeros((16, 16))
    thr = 1.0 / 255.0
    for i in range(cnt):
        # a quad (2x2) stops when its max transmittance is below 1/255
        quadMax = T.reshape(8, 2, 8, 2).max(axis=(1, 3))
        openPx = np.repeat(np.repeat(quadMax >= thr, 2, axis=0), 2, axis=1)
        if not openPx.any():
            break
        gi = int(fr.instanceGaussianIndices[off + i])
        if gi < 0:
            continue
        rd = fr.renderData[gi]
        theta = float(rd["theta"]) * math.pi / 65535.0
        s1, s2 = max(float(rd["sigma1"]), 1e-4), max(float(rd["sigma2"]), 1e-4)
        c, s = math.cos(theta), math.sin(theta)
        iv1, iv2 = 1.0 / (s1 * s1), 1.0 / (s2 * s2)
        A, B, Cq = c * c * iv1 + s * s * iv2, c * s * (iv1 - iv2), s * s * iv1 + c * c * iv2
        dx, dy = X - float(rd["meanX"]), Y - float(rd["meanY"])
        p = dx * dx * A + dy * dy * Cq + dx * dy * (2.0 * B)
        a = np.minimum(float(rd["opacity"]) / 255.0 * np.exp(-0.5 * p), 0.99)
        a = np.where(openPx, a, 0.0)
        w = a * T
        col = np.array([float(rd["colorR"]), float(rd["colorG"]), float(rd["colorB"])]) / 255.0
        Cc += w[..., None] * col
        D += w * float(rd["depth"])
        T = T * (1.0 - a)
    return np.concatenate([Cc, (1.0 - T)[..., None]], axis=2), D


def _psnr(a, b):
    mse = float(np.mean((a - b) ** 2))
    return 99.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def test_contraction_conventions_are_quantified():
    fr, out = _frames()
    cC, dC = out["contracted"]
    cU, dU = out["uncontracted"]
    assert fr.header.overflow == 0 and fr.header.totalInstances > 100_000
    diff = np.abs(cC - cU)
    max_abs = float(diff.max())
    psnr = _psnr(cC, cU)
    frac_equal = float((cC == cU).mean())
    ddiff = float(np.max(np.abs(dC - dU) / np.maximum(np.abs(dU), 1.0)))
    print(f"contracted vs uncontracted: colour max-abs {max_abs:.3e}, PSNR {psnr:.1f} dB, identical halfs {100 * frac_equal:.1f} %, "
          f"depth max rel {ddiff:.3e}")
    # both are roundings of the same real-number expression: they may differ by accumulated half ulps only
    # measured on this scene: max-abs 1.7e-2 (a handful of pixels where a quad's early exit flips one splat earlier or later),
    # 99.9th percentile 1.7e-3, mean 1.0e-4, PSNR 72.1 dB, 71 % of the halfs identical, depth max rel 2.0e-2
    assert max_abs <= 2.5e-2, max_abs
    assert float(np.quantile(diff, 0.999)) <= 4e-3
    assert psnr >= 65.0, psnr
    assert ddiff <= 3e-2, ddiff

    # independent float64 evaluation over the oracle's own tile lists, on a sample of the busiest and of random tiles
    tiles_x = (W + 15) // 16
    counts = fr.tileHeaders[:, 1].astype(np.int64)
    rng = np.random.default_rng(3)
    active = np.nonzero(counts)[0]
    sample = np.unique(np.concatenate([active[np.argsort(counts[active])[-12:]], rng.choice(active, 36, replace=False)]))
    errC, errU, errDC = [], [], []
    for t in sample:
        ref, dref = _textbook_tile(fr, int(t), tiles_x)
        tx, ty = int(t) % tiles_x, int(t) // tiles_x
        ys, xs = slice(ty * 16, min(ty * 16 + 16, H)), slice(tx * 16, min(tx * 16 + 16, W))
        hh, ww = ys.stop - ys.start, xs.stop - xs.start
        errC.append(np.abs(cC[ys, xs] - ref[:hh, :ww]).max())
        errU.append(np.abs(cU[ys, xs] - ref[:hh, :ww]).max())
        errDC.append(np.max(np.abs(dC[ys, xs] - dref[:hh, :ww]) / np.maximum(np.abs(dref[:hh, :ww]), 1.0)))
    eC, eU, eD = float(np.max(errC)), float(np.max(errU)), float(np.max(errDC))
    print(f"vs float64 textbook blend on {len(sample)} tiles (max list {int(counts[sample].max())}): contracted {eC:.3e}, "
          f"uncontracted {eU:.3e}, contracted depth rel {eD:.3e}")
    # a binary16 accumulator over lists of hundreds of splats cannot hold north_star's 2^-10 against float64 (the reference's
    # own half arithmetic does not either); what is asserted is that the half frames are rounding noise away from it
    # and that contracting does not make it worse
    # measured: contracted 1.54e-2 (median tile 2.8e-3), uncontracted 1.79e-2 (median 3.0e-3), depth 1.4e-2 relative --
    # the contracted form is the closer of the two to the float64 frame
    assert eC <= 2.5e-2 and eU <= 2.5e-2, (eC, eU)
    assert eC <= eU + 4e-3, (eC, eU)
    assert eD <= 2.5e-2, eD
