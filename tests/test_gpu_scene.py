"""Scene ingest on the device vs the CPU oracle (bit-exact): PLY decode (standard + compressed), SH re-layout, recentering,
placeholder compaction, bounds, packing, Morton pre-sort, and a frame rendered from a loaded file."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import ply_util as pu  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    from oracle import binding
    binding.build()
    return binding


def _load_both(ob, data, half):
    from gsm_renderer_b200.scene import PLYLoader
    from gsm_renderer_b200.renderer import RenderPrecision
    ds = PLYLoader.load(data, device=0, precision=RenderPrecision.float16 if half else RenderPrecision.float32)
    ref = ob.ply_load(data)
    return ds, ref


def _assert_same(ob, ds, ref, half, order=None):
    res = ref["result"]
    assert ds.count == res.count and ds.shComponents == res.shComponents and ds.harmonicsStride == res.harmonicsStride
    assert ds.compressed == bool(res.compressed) and ds.scaleIsLogSpace == bool(res.scaleIsLogSpace)
    assert ds.opacityIsLogit == bool(res.opacityIsLogit)
    assert np.array_equal(np.array(ds.center, np.float32), np.array(res.center[:], np.float32))
    assert np.array_equal(np.array(ds.boundsCenter, np.float32), np.array(res.boundsCenter[:], np.float32))
    assert np.float32(ds.boundsRadius) == np.float32(res.boundsRadius)
    g = ds.gaussians.cpu().numpy()
    assert np.array_equal(g, ob.pack_gaussians(ref, half, order)), "packed records differ"
    if ds.harmonicsStride:
        h = ds.harmonics.cpu().numpy()
        exp = ob.pack_harmonics(ref["harmonics"], half, order)
        assert np.array_equal(h.view(np.uint16 if half else np.uint32), exp.view(np.uint16 if half else np.uint32)), "harmonics differ"


@pytest.mark.parametrize("half", [True, False])
@pytest.mark.parametrize("deg,placeholders,log_scale,logit,n", [(3, 0, True, True, 20000), (3, 97, True, True, 33333),
                                                                  (1, 0, False, False, 5000), (0, 1, True, False, 257), (2, 0, False, True, 1)])
def test_standard_ply_matches_oracle(ob, half, deg, placeholders, log_scale, logit, n):
    props, cols = pu.standard_scene(n, deg, seed=11 + deg, log_scale=log_scale, logit_opacity=logit, placeholders=min(placeholders, n - 1) if n > 1 else 0)
    data = pu.write_ply(props, cols)
    ds, ref = _load_both(ob, data, half)
    _assert_same(ob, ds, ref, half)


def test_unaligned_body_mixed_types_and_aliases(ob):
    rng = np.random.default_rng(5)
    n = 4097
    cols = {"px": rng.normal(0, 1, n), "py": rng.normal(0, 1, n), "pz": rng.normal(0, 1, n), "sx": rng.uniform(0.01, 0.3, n),
            "sy": rng.uniform(0.01, 0.3, n), "sz": rng.uniform(0.01, 0.3, n), "qw": rng.normal(0, 1, n), "qx": rng.normal(0, 1, n),
            "qy": rng.normal(0, 1, n), "qz": rng.normal(0, 1, n), "alpha": rng.integers(0, 256, n), "sh_2": rng.normal(0, 1, n),
            "sh_0": rng.integers(-100, 100, n), "sh_1": rng.normal(0, 1, n), "junk": rng.integers(0, 60000, n), "tag": rng.integers(-100, 100, n)}
    props = [("char", "tag"), ("double", "px"), ("float", "py"), ("double", "pz"), ("ushort", "junk"), ("float", "sx"), ("float", "sy"),
             ("float", "sz"), ("float", "qw"), ("float", "qx"), ("float", "qy"), ("float", "qz"), ("uchar", "alpha"), ("float", "sh_2"),
             ("short", "sh_0"), ("double", "sh_1")]
    for comment in ("comment a", "comment ab", "comment abc", "comment abcd"):  # every body alignment modulo 4
        data = pu.write_ply(props, cols, eol="\r\n", extra_header=(comment,))
        ds, ref = _load_both(ob, data, True)
        _assert_same(ob, ds, ref, True)


def test_no_sh_and_missing_optional_properties(ob):
    rng = np.random.default_rng(2)
    n = 1000
    cols = {"x": rng.normal(0, 1, n), "y": rng.normal(0, 1, n), "z": rng.normal(0, 1, n)}
    data = pu.write_ply([("float", "x"), ("float", "y"), ("float", "z")], cols)
    ds, ref = _load_both(ob, data, False)
    assert ds.shComponents == 0 and ds.harmonicsStride == 0 and ds.count == n
    # scale = exp(0) = 1, opacity = sigmoid(0) = 0.5; the absent rotation normalises 0 * (1 / 0) = NaN on both sides (the
    # NaN's payload is the one thing IEEE leaves to the implementation, so compare as floats)
    g = ds.gaussians.cpu().numpy().view(np.float32)
    exp = ob.pack_gaussians(ref, False).view(np.float32)
    assert np.array_equal(g, exp, equal_nan=True) and np.isnan(g[:, 8:12]).all()
    assert np.array_equal(g[:, 3], np.full(n, 0.5, np.float32)) and np.array_equal(g[:, 4:7], np.ones((n, 3), np.float32))


@pytest.mark.parametrize("half", [True, False])
def test_compressed_ply_matches_oracle(ob, half):
    for n, sh in ((700, True), (256, False), (100000, False)):
        data = pu.compressed_scene(n, seed=n, with_sh_element=sh)
        ds, ref = _load_both(ob, data, half)
        assert ds.compressed
        _assert_same(ob, ds, ref, half)


def test_loader_errors_are_the_reference_cases(ob):
    from gsm_renderer_b200.scene import PLYLoader, PLYLoaderError
    props, cols = pu.standard_scene(10, 0)
    for data, case in ((pu.write_ply(props, cols, fmt="ascii"), "unsupportedFormat"), (pu.write_ply(props, cols, element="face"), "missingVertexElement"),
                       (pu.write_ply(props, cols)[:-8], "insufficientData"), (pu.write_ply(props[1:], cols), "missingRequiredProperties")):
        with pytest.raises(PLYLoaderError) as e:
            PLYLoader.load(data)
        assert e.value.case == case
        with pytest.raises(ValueError, match=case):
            ob.ply_load(data)


@pytest.mark.parametrize("half", [True, False])
def test_morton_sort_matches_oracle(ob, half):
    from gsm_renderer_b200.scene import GaussianSceneBuilder
    props, cols = pu.standard_scene(50000, 1, seed=21)
    for k in "xyz":                      # duplicates: ties must keep their order
        cols[k][1000:1200] = cols[k][1000]
    data = pu.write_ply(props, cols)
    ds, ref = _load_both(ob, data, half)
    codes, order = ob.morton_order(ref["pos"])
    GaussianSceneBuilder.sortByMortonCode(ds)
    _assert_same(ob, ds, ref, half, order)
    assert np.all(np.diff(codes[order].astype(np.float64)) >= 0)


def test_render_loaded_scene_matches_oracle_frame(ob):
    """A frame rendered from a loaded + Morton-sorted file equals the oracle's frame of the oracle-loaded data."""
    from gsm_renderer_b200.scene import PLYLoader, GaussianSceneBuilder
    from gsm_renderer_b200.renderer import RenderPrecision
    from tests import parity_util as pq
    from gsm_renderer_b200 import synthetic as syn
    cl = syn.synthetic_cloud(20000, 3, seed=4, scale_median=0.02)
    n = cl.count
    # write the synthetic cloud as a standard PLY: log scales, logit opacities, PLY SH order
    pos = np.asarray(cl.positions, np.float32); sc = np.asarray(cl.scales, np.float32)
    rot = np.asarray(cl.rotations, np.float32); op = np.clip(np.asarray(cl.opacities, np.float32), 1e-4, 1 - 1e-4)
    sh = np.asarray(cl.harmonics, np.float32).reshape(n, 3, 16)  # planar [R.., G.., B..]
    cols = {"x": pos[:, 0], "y": pos[:, 1], "z": pos[:, 2], "opacity": np.log(op / (1 - op))}
    props = [("float", "x"), ("float", "y"), ("float", "z")]
    for c in range(3):
        cols[f"f_dc_{c}"] = sh[:, c, 0]; props.append(("float", f"f_dc_{c}"))
    for c in range(45):
        cols[f"f_rest_{c}"] = sh[:, c // 15, 1 + c % 15]; props.append(("float", f"f_rest_{c}"))
    props.append(("float", "opacity"))
    for c in range(3):
        cols[f"scale_{c}"] = np.log(sc[:, c]); props.append(("float", f"scale_{c}"))
    for c, src in enumerate((3, 0, 1, 2)):  # rot_0 = w
        cols[f"rot_{c}"] = rot[:, src]; props.append(("float", f"rot_{c}"))
    data = pu.write_ply(props, cols)
    ds = PLYLoader.load(data, precision=RenderPrecision.float16)
    GaussianSceneBuilder.sortByMortonCode(ds)
    ref = ob.ply_load(data)
    _, order = ob.morton_order(ref["pos"])
    g = ob.pack_gaussians(ref, True, order); h = ob.pack_harmonics(ref["harmonics"], True, order)
    assert np.array_equal(ds.gaussians.cpu().numpy(), g)
    assert np.array_equal(ds.harmonics.cpu().numpy().view(np.uint16), h.view(np.uint16))
    W, H = 640, 360
    # the loader recentred the cloud around the origin: put the camera 11 units back (numpy row 3 = matrix column 3)
    view = np.eye(4, dtype=np.float32); view[3, 2] = 11.0
    cam = pq.default_camera(W, H, view=view)
    gg, hh = g.reshape(-1).view(np.uint8).reshape(ds.count, 32), h
    r, c, d = pq.gpu_mono(gg, hh, "float16", cam, W, H, 16, ds.count, False)
    fr, oc, od = pq.oracle_mono(ob, gg, hh, "float16", cam, W, H, 16, ds.count, False)
    V, I = pq.compare_white_box(r, fr, W, H, ds.count)
    r.close()
    assert V > 1000
    assert np.array_equal(c, np.asarray(oc).view(np.uint16).reshape(c.shape)) and np.array_equal(d, np.asarray(od).view(np.uint16).reshape(d.shape))


def test_empty_file_and_very_wide_records(ob):
    from gsm_renderer_b200.scene import PLYLoader
    # no vertices at all: count 0, bounds (zero, 1.0) like GaussianSceneBuilder.bounds(of: [])
    data = pu.write_ply([("float", "x"), ("float", "y"), ("float", "z")], {"x": np.zeros(0), "y": np.zeros(0), "z": np.zeros(0)}, count=0)
    ds = PLYLoader.load(data)
    ref = ob.ply_load(data)
    assert ds.count == 0 == ref["result"].count and ds.boundsRadius == 1.0 == ref["result"].boundsRadius
    # records too wide to stage 128 of them in shared memory (> 1090 B): the direct-read path
    props, cols = pu.standard_scene(3000, 3, seed=3, placeholders=4)
    rng = np.random.default_rng(1)
    for k in range(140):
        cols[f"junk_{k}"] = rng.normal(0, 1, 3000)
        props.insert(3 + k, ("double", f"junk_{k}"))
    data = pu.write_ply(props, cols)
    for half in (True, False):
        ds, ref = _load_both(ob, data, half)
        _assert_same(ob, ds, ref, half)
