"""Shared by the -m gpu parity tests, __graft_entry__.smoke() and bench.py's checker leg: renders the same
inputs through the CUDA C ABI and through the CPU oracle and compares every white-box buffer.

Bars (BASELINE.json north_star): depth keys, tile counts, instance order, tile ranges bit-exact; blended
colour/depth within 1e-3 max-abs (fp32) or 2^-10 relative (fp16 mode) -- the build is in fact bit-exact
for pixels too, and `pixel_exact=True` asserts that.
"""
from __future__ import annotations

import numpy as np

from gsm_renderer_b200 import synthetic as syn


def make_scene_inputs(cloud, precision: str):
    g, h = cloud.pack(precision)
    return np.ascontiguousarray(g), np.ascontiguousarray(h)


def default_camera(W, H, near=0.1, far=100.0, view=None, position=(0.0, 0.0, 0.0)):
    from gsm_renderer_b200.renderer import CameraParams
    proj = syn.make_projection_matrix(W, H, near, far)
    fx, fy = syn.focal_lengths(W, H)
    return CameraParams(np.eye(4, dtype=np.float32) if view is None else view, proj, position, fx, fy, near, far)


def oracle_mono(ob, g, h, precision, cam, W, H, sh, max_gaussians, srgb, depth_key16=False, tile_id16=True):
    ocam = ob.make_camera(cam.viewMatrix, cam.projectionMatrix, cam.position, W, H, cam.near, cam.far, sh,
                          g.shape[0], srgb)
    fr = ob.OracleFrame(max_gaussians, W, H, depth_key16=depth_key16, tile_id16=tile_id16)
    color, depth = fr.render_mono(g, h, ob.F16 if precision == "float16" else ob.F32, ocam, W, H)
    return fr, color, depth


def gpu_mono(g, h, precision, cam, W, H, sh, max_gaussians, srgb, depth_key16=False, tile_id16=True,
             want_depth=True, max_wh=None):
    import torch
    from gsm_renderer_b200.renderer import (DepthFirstRenderer, GaussianColorSpace, GaussianInput,
                                            RadixSortKeyPrecision, RendererConfig, RenderPrecision)
    dev = torch.device("cuda:0")
    cfg = RendererConfig(maxGaussians=max_gaussians, maxWidth=(max_wh or (W, H))[0], maxHeight=(max_wh or (W, H))[1],
                         precision=RenderPrecision.float16 if precision == "float16" else RenderPrecision.float32,
                         gaussianColorSpace=GaussianColorSpace.srgb if srgb else GaussianColorSpace.linear)
    r = DepthFirstRenderer(device=0, config=cfg,
                           depthSortKeyPrecision=RadixSortKeyPrecision.bits16 if depth_key16 else RadixSortKeyPrecision.bits32,
                           tileIdPrecision=RadixSortKeyPrecision.bits16 if tile_id16 else RadixSortKeyPrecision.bits32)
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    color = torch.full((H, W, 4), 0x7E00, dtype=torch.int16, device=dev)  # NaN pattern = "untouched"
    depth = torch.full((H, W), 0x7E00, dtype=torch.int16, device=dev) if want_depth else None
    stream = torch.cuda.current_stream()
    r.render(stream, color, depth, GaussianInput(tg, th, g.shape[0], sh), cam, W, H)
    torch.cuda.synchronize()
    c = color.cpu().numpy().view(np.uint16)
    d = depth.cpu().numpy().view(np.uint16) if want_depth else None
    return r, c, d


def compare_white_box(r, fr, W, H, n_gaussians, stereo=False):
    """Every debugRead* buffer against the oracle's arrays, bit-exact. Returns (V, I)."""
    hd = r.debugReadHeader()
    oh = fr.header
    for f in ("visibleCount", "totalInstances", "paddedVisibleCount", "paddedInstanceCount", "overflow"):
        assert getattr(hd, f) == getattr(oh, f), f"header.{f}: gpu {getattr(hd, f)} oracle {getattr(oh, f)}"
    V, I = hd.visibleCount, hd.totalInstances
    T = ((W + 15) // 16) * ((H + 15) // 16)
    nt = r.debugReadNTouchedTiles(n_gaussians)
    assert np.array_equal(nt, fr.nTouched[:n_gaussians]), "nTouchedTiles"
    assert np.array_equal(r.debugReadTileBounds(n_gaussians), fr.bounds[:n_gaussians]), "tile bounds"
    vis = nt > 0
    rd_g = r.debugReadRenderData(n_gaussians)
    rd_o = fr.renderData[:n_gaussians]
    assert rd_g[vis].tobytes() == rd_o[vis].tobytes(), "renderData of visible Gaussians"
    assert np.array_equal(r.debugReadDepthKeys(V), fr.depthKeys[:V]), "sorted depth keys"
    assert np.array_equal(r.debugReadSortedPrimitiveIndices(V), fr.primitiveIndices[:V]), "sorted primitive indices"
    assert np.array_equal(r.debugReadInstanceOffsets(V), fr.orderedTileCounts[:V]), "instance offsets"
    assert np.array_equal(r.debugReadSortedTileIds(I), fr.instanceTileIds[:I].astype(np.uint32)), "sorted tile ids"
    assert np.array_equal(r.debugReadInstanceGaussianIndices(I), fr.instanceGaussianIndices[:I]), "instance order"
    assert np.array_equal(r.debugReadTileHeaders(T), fr.tileHeaders[:T]), "tile headers"
    na = r.debugReadActiveTileCount()
    assert na == fr.f.activeTileCount, "active tile count"
    assert set(r.debugReadActiveTiles(na).tolist()) == set(fr.activeTiles[:na].tolist()), "active tile set"
    return V, I


def compare_pixels(gpu_bits, ref_bits, pixel_exact=True, fp16_mode=True, what="colour"):
    """uint16 half bit patterns. Tolerance bar, then (optionally) the bit-exact bar."""
    g = gpu_bits.view(np.float16).astype(np.float64)
    o = ref_bits.view(np.float16).astype(np.float64)
    both_nan = np.isnan(g) & np.isnan(o)
    diff = np.abs(np.where(both_nan, 0.0, g - o))
    assert not np.isnan(diff).any(), f"{what}: NaN on one side only"
    if fp16_mode:
        tol = np.maximum(np.abs(o), 2.0 ** -14) * 2.0 ** -10 + 1e-7  # 2^-10 relative
        bad = diff > tol
    else:
        bad = diff > 1e-3
    assert not bad.any(), f"{what}: {int(bad.sum())} values outside tolerance, max abs diff {diff.max():.3g}"
    if pixel_exact:
        neq = (gpu_bits != ref_bits) & ~both_nan
        assert not neq.any(), f"{what}: {int(neq.sum())} values differ in bits (max abs {diff.max():.3g})"
    return float(diff.max())


def run_mono_case(ob, cloud, precision, W, H, near=0.1, far=100.0, srgb=False, sh=None, max_gaussians=None,
                  depth_key16=False, tile_id16=True, view=None, position=(0.0, 0.0, 0.0), pixel_exact=True):
    sh = cloud.sh_components if sh is None else sh
    G = max_gaussians or cloud.count
    g, h = make_scene_inputs(cloud, precision)
    cam = default_camera(W, H, near, far, view, position)
    fr, oc, od = oracle_mono(ob, g, h, precision, cam, W, H, sh, G, srgb, depth_key16, tile_id16)
    r, gc, gd = gpu_mono(g, h, precision, cam, W, H, sh, G, srgb, depth_key16, tile_id16)
    try:
        V, I = compare_white_box(r, fr, W, H, cloud.count)
        e1 = compare_pixels(gc, oc, pixel_exact, precision == "float16", "colour")
        e2 = compare_pixels(gd, od, pixel_exact, precision == "float16", "depth")
    finally:
        r.close()
    return dict(N=cloud.count, V=V, I=I, activeTiles=int(fr.f.activeTileCount), maxColourDiff=e1, maxDepthDiff=e2)
