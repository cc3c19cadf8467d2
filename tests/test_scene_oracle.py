"""Scene ingest (SURVEY.md 8(f) rank 1 + Morton pre-sort), CPU side: the oracle's restatement of PLYLoader.swift /
Scene.swift is pinned against an independent numpy reading of the same rules, and the product's host-side header parser
(gsm_ply_probe, and the error paths of gsm_ply_load that fail before any CUDA call) is checked against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import ply_util as pu  # noqa: E402


@pytest.fixture(scope="module")
def ob():
    from oracle import binding
    binding.build()
    return binding


def _f32(a):
    return np.asarray(a).astype(np.float32)


def _numpy_standard(props, cols, sh_names):
    """Independent restatement (numpy, float32) of PLYLoader.swift:517-741 for all-float files without aliases."""
    n = len(cols["x"])
    s = np.stack([_f32(cols[f"scale_{k}"]) for k in range(3)], 1)
    op = _f32(cols["opacity"])
    keep = ~((s[:, 0] == 2) & (s[:, 1] == 2) & (s[:, 2] == 2) & (np.abs(op - np.float32(4.8402)) < np.float32(0.001)))
    samp = slice(0, min(100, n))
    s0 = s[samp, 0]
    scale_log = True
    if not (s0 < 0).any() and not (s0 > 1).any() and 0 < s0.astype(np.float32).sum(dtype=np.float32) / np.float32(len(s0)) < 0.5:
        scale_log = False
    op_logit = bool(op[samp].min() < 0 or op[samp].max() > 1)
    pos = np.stack([_f32(cols[k]) for k in "xyz"], 1)[keep]
    c = (pos.min(0) + pos.max(0)) * np.float32(0.5)
    if np.sqrt((c * c).sum(dtype=np.float32)) > 1e-6:
        pos = pos - c
    else:
        c = np.zeros(3, np.float32)
    sh = np.stack([_f32(cols[nm]) for nm in sh_names], 1)[keep] if sh_names else np.zeros((keep.sum(), 0), np.float32)
    K = len(sh_names) // 3
    hoc = K - 1
    planar = np.zeros_like(sh)
    if K:
        planar[:, 0] = sh[:, 0]; planar[:, 1:K] = sh[:, 3:3 + hoc]
        planar[:, K] = sh[:, 1]; planar[:, K + 1:2 * K] = sh[:, 3 + hoc:3 + 2 * hoc]
        planar[:, 2 * K] = sh[:, 2]; planar[:, 2 * K + 1:3 * K] = sh[:, 3 + 2 * hoc:3 + 3 * hoc]
    return dict(keep=keep, pos=pos, center=c, scale_log=scale_log, op_logit=op_logit, sh=planar, K=K, raw_scale=s[keep], raw_op=op[keep])


@pytest.mark.parametrize("deg,placeholders,log_scale,logit", [(3, 0, True, True), (3, 7, True, True), (1, 0, False, False),
                                                                (0, 3, True, False), (2, 0, False, True)])
def test_oracle_standard_layout_matches_numpy(ob, deg, placeholders, log_scale, logit):
    props, cols = pu.standard_scene(1500, deg, seed=deg + placeholders, log_scale=log_scale, logit_opacity=logit,
                                    placeholders=placeholders)
    r = ob.ply_load(pu.write_ply(props, cols))
    k = (deg + 1) ** 2
    names = [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(3 * (k - 1))]
    ref = _numpy_standard(props, cols, names)
    res = r["result"]
    assert res.count == ref["keep"].sum() == 1500 - placeholders
    assert res.shComponents == k and res.harmonicsStride == 3 * k
    assert bool(res.scaleIsLogSpace) == ref["scale_log"] == log_scale and bool(res.opacityIsLogit) == ref["op_logit"] == logit
    assert np.array_equal(r["pos"], ref["pos"]) and np.array_equal(np.array(res.center[:], np.float32), ref["center"])
    assert np.array_equal(r["harmonics"], ref["sh"])
    if log_scale:
        assert np.allclose(r["scale"], np.exp(ref["raw_scale"].astype(np.float64)), rtol=3e-7)
    else:
        assert np.array_equal(r["scale"], ref["raw_scale"])
    if logit:
        assert np.allclose(r["opacity"], 1 / (1 + np.exp(-ref["raw_op"].astype(np.float64))), rtol=5e-7)
    else:
        assert np.array_equal(r["opacity"], ref["raw_op"])
    q = np.stack([_f32(cols[f"rot_{i}"]) for i in (1, 2, 3, 0)], 1)[ref["keep"]].astype(np.float64)
    assert np.allclose(r["rot"], q / np.linalg.norm(q, axis=1, keepdims=True), atol=3e-7)  # (x, y, z, w) = rot_1..3, rot_0


def test_oracle_aliases_types_and_crlf(ob):
    rng = np.random.default_rng(5)
    n = 300
    cols = {"px": rng.normal(0, 1, n), "py": rng.normal(0, 1, n), "pz": rng.normal(0, 1, n), "sx": rng.uniform(0.01, 0.3, n),
            "sy": rng.uniform(0.01, 0.3, n), "sz": rng.uniform(0.01, 0.3, n), "qw": rng.normal(0, 1, n), "qx": rng.normal(0, 1, n),
            "qy": rng.normal(0, 1, n), "qz": rng.normal(0, 1, n), "alpha": rng.integers(0, 256, n), "sh_2": rng.normal(0, 1, n),
            "sh_0": rng.integers(-100, 100, n), "sh_1": rng.normal(0, 1, n), "junk": rng.integers(0, 60000, n)}
    props = [("double", "px"), ("float", "py"), ("double", "pz"), ("ushort", "junk"), ("float", "sx"), ("float", "sy"), ("float", "sz"),
             ("float", "qw"), ("float", "qx"), ("float", "qy"), ("float", "qz"), ("uchar", "alpha"), ("float", "sh_2"),
             ("short", "sh_0"), ("double", "sh_1")]
    data = pu.write_ply(props, cols, eol="\r\n", extra_header=("comment made by a test", "obj_info whatever"))
    r = ob.ply_load(data)
    res = r["result"]
    assert res.count == n and res.shComponents == 1 and res.harmonicsStride == 3
    assert not res.scaleIsLogSpace and not res.opacityIsLogit      # small positive scales; uchar alpha / 255 in [0, 1]
    assert np.array_equal(r["opacity"], (cols["alpha"].astype(np.uint8).astype(np.float32) / np.float32(255.0)))
    assert np.array_equal(r["scale"][:, 0], _f32(cols["sx"]))
    # SH sorted by key sh_0, sh_1, sh_2 whatever the property order; planar with K = 1 is [R0, G0, B0]
    exp = np.stack([cols["sh_0"].astype(np.int16).astype(np.float32), cols["sh_1"].astype(np.float64).astype(np.float32), _f32(cols["sh_2"])], 1)
    assert np.array_equal(r["harmonics"], exp)
    x = cols["px"].astype(np.float64).astype(np.float32)
    assert np.array_equal(r["pos"][:, 0], x - (x.min() + x.max()) * np.float32(0.5))


def test_oracle_header_errors(ob):
    props, cols = pu.standard_scene(10, 0)
    with pytest.raises(ValueError, match="unsupportedFormat"):
        ob.ply_load(pu.write_ply(props, cols, fmt="ascii"))
    with pytest.raises(ValueError, match="missingVertexElement"):
        ob.ply_load(pu.write_ply(props, cols, element="face"))
    with pytest.raises(ValueError, match="insufficientData"):
        ob.ply_load(pu.write_ply(props, cols)[:-8])
    with pytest.raises(ValueError, match="missingRequiredProperties"):
        ob.ply_load(pu.write_ply(props[1:], cols))
    with pytest.raises(ValueError, match="invalidHeader"):
        ob.ply_load(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\n")
    with pytest.raises(ValueError, match="invalidHeader"):
        ob.ply_load(pu.write_ply(props, cols).replace(b"comment", b"remark") if b"comment" in pu.write_ply(props, cols)
                    else pu.write_ply(props, cols, extra_header=("remark unknown keyword",)))
    with pytest.raises(ValueError, match="listPropertiesNotSupported"):
        ob.ply_load(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nproperty float x\nproperty float y\nproperty float z\n"
                    b"property list uchar int idx\nend_header\n")


def test_oracle_compressed_layout_matches_numpy(ob):
    n = 700
    data = pu.compressed_scene(n, seed=3, with_sh_element=True)
    r = ob.ply_load(data)
    res = r["result"]
    assert res.compressed and res.count == n and res.shComponents == 1 and res.harmonicsStride == 3
    body = data[data.index(b"end_header\n") + 11:]
    nch = (n + 255) // 256
    ch = np.frombuffer(body[:nch * 72], "<f4").reshape(nch, 18)
    vx = np.frombuffer(body[nch * 72:nch * 72 + n * 16], "<u4").reshape(n, 4)
    f = np.float32

    def unorm(v, bits):
        m = (1 << bits) - 1
        return (v & m).astype(f) / f(m)

    def lerp(a, b, t):
        return a * (f(1) - t) + b * t
    c = ch[np.arange(n) // 256]
    pos = np.stack([lerp(c[:, 0], c[:, 3], unorm(vx[:, 0] >> 21, 11)), lerp(c[:, 1], c[:, 4], unorm(vx[:, 0] >> 11, 10)),
                    lerp(c[:, 2], c[:, 5], unorm(vx[:, 0], 11))], 1)
    cen = (pos.min(0) + pos.max(0)) * f(0.5)
    assert np.array_equal(r["pos"], pos - cen)
    assert np.array_equal(r["opacity"], unorm(vx[:, 3], 8))
    shc0 = f(0.28209479177387814)
    assert np.array_equal(r["harmonics"][:, 0], (lerp(c[:, 12], c[:, 15], unorm(vx[:, 3] >> 24, 8)) - f(0.5)) / shc0)
    norm = f(1) / (np.sqrt(f(2)) * f(0.5))
    a, b, cc = [(unorm(vx[:, 1] >> s, 10) - f(0.5)) * norm for s in (20, 10, 0)]
    m = np.sqrt(np.maximum(f(0), f(1) - ((a * a + b * b) + cc * cc)))
    sel = vx[:, 1] >> 30
    exp = np.where((sel == 0)[:, None], np.stack([a, b, cc, m], 1), np.where((sel == 1)[:, None], np.stack([m, b, cc, a], 1),
                   np.where((sel == 2)[:, None], np.stack([b, m, cc, a], 1), np.stack([b, cc, m, a], 1))))
    assert np.array_equal(r["rot"], exp)
    ls = lerp(c[:, 6], c[:, 9], unorm(vx[:, 2] >> 21, 11))
    assert np.allclose(r["scale"][:, 0], np.exp(ls.astype(np.float64)), rtol=3e-7)


def test_oracle_morton_matches_python(ob):
    rng = np.random.default_rng(9)
    pos = rng.normal(0, 1, (500, 3)).astype(np.float32)
    pos[100:110] = pos[100]  # ties keep their order
    codes, order = ob.morton_order(pos)
    mn, mx = pos.min(0), pos.max(0)
    inv = np.where(mx - mn > 1e-6, np.float32(1) / (mx - mn), np.float32(0)).astype(np.float32)
    q = np.clip(((pos - mn) * inv) * np.float32(2097151.0), 0, np.float32(2097151.0)).astype(np.uint64)

    def expand(v):
        out = 0
        for b in range(21):
            out |= ((int(v) >> b) & 1) << (3 * b)
        return out
    exp = np.array([expand(x) | (expand(y) << 1) | (expand(z) << 2) for x, y, z in q], np.uint64)
    assert np.array_equal(codes, exp)
    assert np.array_equal(order, np.argsort(exp, kind="stable").astype(np.uint32))


def test_product_header_parser_matches_oracle(ob):
    """gsm_ply_probe and the pre-CUDA error paths of gsm_ply_load (no GPU needed)."""
    from gsm_renderer_b200 import _native as N
    lib = N.lib()

    def probe(data):
        a = np.frombuffer(data, np.uint8)
        info = N.gsm_ply_info()
        return lib.gsm_ply_probe(a.ctypes.data, a.size, C.byref(info)), info

    def load_status(data):
        a = np.frombuffer(data, np.uint8)
        out = N.gsm_scene_info()
        return lib.gsm_ply_load(0, None, a.ctypes.data, a.size, 1, 1, 1, 0, 0, C.byref(out))  # dummy non-null pointers, capacity 0
    props, cols = pu.standard_scene(64, 2, placeholders=2)
    data = pu.write_ply(props, cols, eol="\r\n")
    st, info = probe(data)
    assert st == 0 and info.vertexCount == 64 and info.format == 1 and not info.compressed and info.shProperties == 27
    assert info.bodyOffset == data.index(b"end_header\r\n") + 12
    st, info = probe(pu.compressed_scene(300))
    assert st == 0 and info.compressed and info.vertexCount == 300 and info.shProperties == 3
    assert load_status(pu.write_ply(props, cols, fmt="ascii")) == 21
    assert load_status(pu.write_ply(props, cols, element="face")) == 22
    assert load_status(pu.write_ply(props[1:], cols)) == 23
    assert load_status(pu.write_ply(props, cols)[:-8]) == 25
    assert load_status(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\n") == 20
    assert load_status(pu.write_ply(props, cols, extra_header=("remark unknown keyword",))) == 20
    assert load_status(data) == 25   # valid file, but the caller's buffers (capacity 0) are too small: refused before any CUDA call
    assert b"too small" in lib.gsm_last_error_string()
