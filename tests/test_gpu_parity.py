"""GPU-vs-oracle parity through the C ABI (run on the B200 box: pytest -m gpu).

Every stage's white-box buffer is compared bit-for-bit with the CPU oracle on the same seeded inputs;
pixels are held to the north_star tolerance AND to bit-exactness."""
import numpy as np
import pytest

from gsm_renderer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("pytest -m gpu needs a CUDA device")
    import tests.parity_util as p
    return p


# ---------------------------------------------------------------- math: device restatement == oracle, bit for bit
def test_math_probes_bit_exact(oracle):
    from gsm_renderer_b200.renderer import probe_math
    rng = np.random.default_rng(1)
    th = (np.arange(65536, dtype=np.float32) * np.float32(np.float32(np.pi) / np.float32(65535.0)))
    x = np.concatenate([th, rng.uniform(-30, 30, 200_000).astype(np.float32)])
    s, c = oracle.probe_sincos(x)
    assert np.array_equal(probe_math(0, x).view(np.uint32), s.view(np.uint32))
    assert np.array_equal(probe_math(1, x).view(np.uint32), c.view(np.uint32))
    xl = np.exp(rng.uniform(-14, 14, 300_000)).astype(np.float32)
    assert np.array_equal(probe_math(2, xl).view(np.uint32), oracle.probe_log(xl).view(np.uint32))
    ay, ax = rng.normal(0, 1, 300_000).astype(np.float32), rng.normal(0, 1, 300_000).astype(np.float32)
    ay[:1000] = 0.0
    ax[500:1500] = 0.0
    assert np.array_equal(probe_math(3, ay, ax).view(np.uint32), oracle.probe_atan2(ay, ax).view(np.uint32))
    xp = rng.uniform(0.03, 1.0, 300_000).astype(np.float32)
    assert np.array_equal(probe_math(4, xp).view(np.uint32), oracle.probe_powr(xp, 2.4).view(np.uint32))
    # min/max: NaN operand loses, -0 < +0 (one FMNMX on the device)
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.nan, np.inf, -np.inf, 1e-45, -1e-45, 3.5], np.float32)
    ma, mb = [x.ravel() for x in np.meshgrid(sp, sp)]
    ma = np.concatenate([ma, rng.normal(0, 1, 10000).astype(np.float32)])
    mb = np.concatenate([mb, rng.normal(0, 1, 10000).astype(np.float32)])
    omn, omx = oracle.probe_minmax(ma, mb)
    gmn, gmx = probe_math(9, ma, mb), probe_math(10, ma, mb)
    nn = ~(np.isnan(ma) & np.isnan(mb))  # both NaN: payload unspecified
    assert np.array_equal(gmn.view(np.uint32)[nn], omn.view(np.uint32)[nn])
    assert np.array_equal(gmx.view(np.uint32)[nn], omx.view(np.uint32)[nn])
    # fused half fma (HFMA2) == the oracle's exact long-double fma, incl. subnormals, overflow, cancellation
    tri = rng.integers(0, 65536, (400_000, 3)).astype(np.uint16)
    small = (rng.normal(0, 1, (200_000, 3)) * np.array([300.0, 0.01, 3.0])).astype(np.float16).view(np.uint16)
    tri = np.concatenate([tri, small])
    got = probe_math(11, tri)
    want = oracle.probe_hfma(tri[:, 0].copy(), tri[:, 1].copy(), tri[:, 2].copy())
    isn = np.isnan(want.view(np.float16))
    assert np.array_equal(np.isnan(got.view(np.float16)), isn) and np.array_equal(got[~isn], want[~isn])
    bits = np.arange(65536, dtype=np.uint16)
    ref = oracle.probe_hexp(bits)
    assert np.array_equal(probe_math(5, bits), ref)          # scalar half exp, all 65536 inputs
    assert np.array_equal(probe_math(7, bits), ref)          # packed half2 variant
    notnan = ~np.isnan(bits.view(np.float16))
    assert np.array_equal(probe_math(8, bits)[notnan], ref[notnan])  # FFMA2 form used by the blend kernel
    # exp(-0.5h * p) with the -0.5 folded into the binary32 constant (what the blend kernels call), all 65536 p
    with np.errstate(invalid="ignore"):
        x = (np.float16(-0.5) * bits.view(np.float16)).astype(np.float16)
    assert np.array_equal(probe_math(12, bits)[notnan], oracle.probe_hexp(x.view(np.uint16))[notnan])
    # the mono blend's shared-memory table form: identical to the polynomial on ALL 65 536 inputs (NaNs included: the table is
    # built from that very function), hence to the oracle wherever the input is not a NaN
    tab = probe_math(13, bits)
    assert np.array_equal(tab, probe_math(12, bits))
    assert np.array_equal(tab[notnan], oracle.probe_hexp(x.view(np.uint16))[notnan])
    shuffled = rng.permutation(bits)                          # pairs of unrelated values share a half2 in the probe
    assert np.array_equal(probe_math(13, shuffled), probe_math(12, shuffled))
    # the blend kernels' XU-pipe form (MUFU.EX2 + rounding guard, gsm_dmath.cuh): identical to the polynomial, hence to the oracle,
    # on every non-NaN input, whatever value shares the half2; the guard catches every input the bare MUFU.EX2 gets wrong
    mufu, raw, flag = probe_math(14, bits), probe_math(15, bits), probe_math(16, bits)
    assert np.array_equal(mufu[notnan], probe_math(12, bits)[notnan])
    assert np.array_equal(probe_math(14, shuffled)[~np.isnan(shuffled.view(np.float16))],
                          probe_math(12, shuffled)[~np.isnan(shuffled.view(np.float16))])
    wrong = notnan & (raw != probe_math(12, bits))
    assert not np.any(wrong & (flag == 0)), "an input the unguarded form gets wrong passes the guard"
    assert int(flag[notnan].sum()) < 400                     # the guard is the rare path (about 1 input in 1000 of [0, 35])
    # the form the blend kernels RUN: MUFU.EX2 on a tuned argument, no guard, one exceptional input (gsm_dmath.cuh) -- equal to the
    # oracle on every non-NaN input, whichever value shares the half2
    tuned = probe_math(17, bits)
    assert np.array_equal(tuned[notnan], oracle.probe_hexp(x.view(np.uint16))[notnan])
    assert np.array_equal(probe_math(17, shuffled)[~np.isnan(shuffled.view(np.float16))],
                          probe_math(12, shuffled)[~np.isnan(shuffled.view(np.float16))])
    xf = np.concatenate([rng.normal(0, 300, 200_000), [65504, 65520, 1e10, -1e10, 6e-8, 2.9e-8, 0.0]]).astype(np.float32)
    assert np.array_equal(probe_math(6, xf), oracle.probe_f2h(xf))


def test_blend_runs_the_xu_pipe_exp_and_equals_the_polynomial_build_gpu(pu):
    """gsm_renderer_create's self-test must select the MUFU.EX2 form on a B200 (mode 2; a silent fall-back to the polynomial would
    hide a regression), and a process forced to the polynomial (GSM_BLEND_EXP=poly) must produce the same mono, stereo and Global
    frames byte for byte."""
    import hashlib, json, os, subprocess, sys
    code = r'''
import hashlib, json, sys
import numpy as np, torch
sys.path.insert(0, %r)
from gsm_renderer_b200 import synthetic as syn
from gsm_renderer_b200 import _native as N
from gsm_renderer_b200.renderer import (DepthFirstRenderer, GlobalRenderer, GaussianInput, RendererConfig, RenderPrecision, StereoCameraParams, StereoRenderTarget)
import tests.parity_util as pu
cl = syn.synthetic_cloud(60_000, 2, seed=11, scale_median=0.02)
g, h = pu.make_scene_inputs(cl, "float16")
W, H = 1280, 720
cam = pu.default_camera(W, H)
dev = torch.device("cuda:0")
tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev); th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
inp = GaussianInput(tg, th, g.shape[0], cl.sh_components)
out = {}
for name, cls in (("mono", DepthFirstRenderer), ("global", GlobalRenderer)):
    r = cls(device=0, config=RendererConfig(maxGaussians=cl.count, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16))
    color = torch.zeros((H, W, 4), dtype=torch.int16, device=dev); depth = torch.zeros((H, W), dtype=torch.int16, device=dev)
    r.render(torch.cuda.current_stream(), color, depth, inp, cam, W, H)
    torch.cuda.synchronize()
    out[name] = hashlib.sha256(color.cpu().numpy().tobytes() + depth.cpu().numpy().tobytes()).hexdigest()
    if name == "mono":
        camR = pu.default_camera(W, H, position=(0.065, 0.0, 0.0))
        sbs = torch.zeros((H, 2 * W, 4), dtype=torch.int16, device=dev)
        r.renderStereo(torch.cuda.current_stream(), StereoRenderTarget.sideBySide(sbs), inp, StereoCameraParams(cam, camR), W, H)
        torch.cuda.synchronize()
        out["stereo"] = hashlib.sha256(sbs.cpu().numpy().tobytes()).hexdigest()
    r.close()
out["mode"] = int(N.lib().gsm_blend_exp_mode(0))
print(json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for label, env in (("tuned", {}), ("poly", {"GSM_BLEND_EXP": "poly"})):
        e = dict(os.environ); e.pop("GSM_BLEND_EXP", None); e.update(env)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res[label] = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["tuned"]["mode"] == 2, "the self-test rejected the MUFU.EX2 form on this device"
    assert res["poly"]["mode"] == 1
    for k in ("mono", "global", "stereo"):
        assert res["tuned"][k] == res["poly"][k], k


# ---------------------------------------------------------------- sorts: the reference's own KATs on the device
def _gpu_sort(keys, payload, key_bits, passes):
    import torch
    from gsm_renderer_b200.renderer import DepthFirstRenderer, RendererConfig
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=1024, maxWidth=64, maxHeight=64))
    k = torch.from_numpy(keys.copy()).cuda()
    p = torch.from_numpy(payload.copy()).cuda()
    r.sortPairs(torch.cuda.current_stream(), k, p, keys.size, key_bits, passes)
    torch.cuda.synchronize()
    out = k.cpu().numpy(), p.cpu().numpy()
    r.close()
    return out


def test_depth_sort_simple_kat_gpu():
    k, p = _gpu_sort(np.arange(10, 0, -1).astype(np.int32).view(np.uint32).astype(np.uint32).view(np.int32),
                     (np.arange(10) * 100).astype(np.int32), 32, 4)
    assert k.tolist() == list(range(1, 11))
    assert p.tolist() == [900, 800, 700, 600, 500, 400, 300, 200, 100, 0]  # DepthFirstUnitTests.swift:145,304


def test_depth_sort_at_scale_kat_gpu(oracle):
    n = 1_000_000
    i = np.arange(n, dtype=np.int64)
    keys = ((i * 37 + 12345) & 0xFFFF).astype(np.uint32)  # DepthFirstUnitTests.swift:313-315
    k, p = _gpu_sort(keys.view(np.int32), i.astype(np.int32), 32, 4)
    ok, op = oracle.sort_pairs_u32(keys, i.astype(np.int32), 4)
    assert np.array_equal(k.view(np.uint32), ok) and np.array_equal(p, op)


@pytest.mark.parametrize("n", [1, 31, 2047, 2048, 2049, 4096, 4097, 123_457])
def test_sort_edge_sizes_gpu(oracle, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys[rng.integers(0, n, max(1, n // 3))] = keys[0]  # ties: stability is visible in the payload
    pay = np.arange(n, dtype=np.int32)
    k, p = _gpu_sort(keys.view(np.int32), pay, 32, 4)
    ok, op = oracle.sort_pairs_u32(keys, pay, 4)
    assert np.array_equal(k.view(np.uint32), ok) and np.array_equal(p, op)
    k16 = (keys & 0x1FFF).astype(np.uint16)
    k, p = _gpu_sort(k16.view(np.int16), pay, 16, 2)
    ok, op = oracle.sort_pairs_u16(k16, pay, 2)
    assert np.array_equal(k.view(np.uint16), ok) and np.array_equal(p, op)


# ---------------------------------------------------------------- whole frames, every stage compared
def test_pipeline_stages_scene_gpu(oracle, pu):
    # DepthFirstUnitTests.swift:21-117 (float32 records, shComponents 1, default sRGB colour space)
    cl = syn.pipeline_stages_scene()
    res = pu.run_mono_case(oracle, cl, "float32", 640, 480, 0.1, 10.0, srgb=True, sh=1)
    assert 0 < res["V"] <= 1000 and res["I"] > 0


@pytest.mark.parametrize("precision,deg,srgb", [("float32", 1, False), ("float16", 3, False), ("float32", 0, True),
                                                ("float16", 2, True), ("float32", 3, False), ("float16", 1, False)])
def test_synthetic_frames_gpu(oracle, pu, precision, deg, srgb):
    cl = syn.synthetic_cloud(60_000, deg, seed=21 + deg, scale_median=0.015)
    res = pu.run_mono_case(oracle, cl, precision, 1920, 1080, srgb=srgb)
    assert res["V"] > 20_000 and res["I"] > res["V"]


def test_config1_50k_sh1_f32_gpu(oracle, pu):
    # BASELINE.json configs[0]: 50k Gaussians, SH degree 1, 1920x1080, float32
    cl = syn.synthetic_cloud(50_000, 1, seed=42, scale_median=0.015)
    res = pu.run_mono_case(oracle, cl, "float32", 1920, 1080)
    assert res["V"] > 20_000


def test_odd_sizes_and_partial_tiles_gpu(oracle, pu):
    cl = syn.synthetic_cloud(20_000, 1, seed=5, scale_median=0.03)
    for (W, H) in [(1919, 1079), (333, 77), (17, 33), (16, 16)]:
        pu.run_mono_case(oracle, cl, "float16", W, H, max_gaussians=cl.count)


def test_reference_fixture_overflow_gpu(oracle, pu):
    # generateVisibleGaussians: ~100 tiles per splat => I > 4N: clamp + overflow flag
    cl = syn.generate_visible_gaussians(3000, 42)
    res = pu.run_mono_case(oracle, cl, "float32", 640, 480, 0.1, 10.0, srgb=False, sh=0)
    assert res["I"] == 4 * 3000
    cl = syn.generate_grid_gaussians(2500, 42)
    pu.run_mono_case(oracle, cl, "float16", 800, 600, 0.1, 10.0, srgb=True, sh=0)


def test_edge_cases_gpu(oracle, pu):
    base = syn.synthetic_cloud(4096, 0, seed=9)
    # all culled (behind the camera): target cleared to (0,0,0,1)
    cl = syn.synthetic_cloud(4096, 0, seed=9)
    cl.positions[:, 2] = -3.0
    res = pu.run_mono_case(oracle, cl, "float32", 200, 120)
    assert res["V"] == 0 and res["I"] == 0
    # one Gaussian
    one = syn.Cloud(base.positions[:1] * 0 + [[0, 0, 5]], base.scales[:1] * 0 + 0.05, base.rotations[:1],
                    base.opacities[:1] * 0 + 0.9, base.harmonics[:1], 1)
    res = pu.run_mono_case(oracle, one, "float32", 320, 200)
    assert res["V"] == 1
    # a Gaussian covering the whole screen + dx,dy > 256 px (half overflow paths of the blend)
    big = syn.Cloud(np.array([[0, 0, 3], [0.5, 0.2, 4]], np.float32), np.array([[3, 3, 3], [2, 0.01, 2]], np.float32),
                    base.rotations[:2], np.array([0.9, 0.7], np.float32), base.harmonics[:2], 1)
    pu.run_mono_case(oracle, big, "float32", 1280, 720, max_gaussians=4096)
    # ties in depth (identical z) + counts straddling multiples of 256/1024/2048
    for n in (255, 256, 257, 1023, 1025, 2047, 2049):
        t = syn.synthetic_cloud(n, 1, seed=n, scale_median=0.03)
        t.positions[:, 2] = np.float32(6.0)
        pu.run_mono_case(oracle, t, "float16", 640, 360)


def test_silent_noop_gpu(oracle, pu):
    # DepthFirstRenderer.swift:249: count > maxGaussians -> nothing enqueued, target untouched
    cl = syn.synthetic_cloud(300, 0, seed=2)
    g, h = pu.make_scene_inputs(cl, "float32")
    cam = pu.default_camera(64, 64)
    r, c, d = pu.gpu_mono(g, h, "float32", cam, 64, 64, 0, max_gaussians=100, srgb=False)
    r.close()
    assert np.all(c == 0x7E00) and np.all(d == 0x7E00)


def test_precision_knobs_gpu(oracle, pu):
    cl = syn.synthetic_cloud(30_000, 1, seed=77, scale_median=0.02)
    pu.run_mono_case(oracle, cl, "float16", 1280, 720, depth_key16=True)    # .bits16 depth keys (DFS.metal:607-612)
    pu.run_mono_case(oracle, cl, "float16", 1280, 720, tile_id16=False)     # .bits32 tile ids (DFS.metal:718)


def test_moved_camera_gpu(oracle, pu):
    cl = syn.synthetic_cloud(40_000, 2, seed=31, scale_median=0.02)
    view, pos = syn.orbit_cameras(3, seed=3)[2]
    pu.run_mono_case(oracle, cl, "float16", 1280, 720, view=view, position=pos)


def test_render_host_matches_device_path(oracle, pu):
    import torch
    cl = syn.synthetic_cloud(20_000, 3, seed=4, scale_median=0.02)
    g, h = pu.make_scene_inputs(cl, "float16")
    W, H = 640, 360
    cam = pu.default_camera(W, H)
    r, c, d = pu.gpu_mono(g, h, "float16", cam, W, H, 16, cl.count, False)
    hc = np.zeros((H, W, 4), np.uint16)
    hd = np.zeros((H, W), np.uint16)
    r.renderHost(g, h, cl.count, 16, cam, W, H, hc, hd)
    assert np.array_equal(hc, c) and np.array_equal(hd, d)
    # encode now, wait later (two frames queued back to back on the renderer's stream)
    hc2 = np.zeros((H, W, 4), np.uint16)
    hd2 = np.zeros((H, W), np.uint16)
    r.renderHostAsync(g, h, cl.count, 16, cam, W, H, hc2, hd2)
    r.waitHost()
    r.close()
    assert np.array_equal(hc2, c) and np.array_equal(hd2, d)


# ---------------------------------------------------------------- stereo
def _stereo_inputs(W, H):
    from gsm_renderer_b200.renderer import CameraParams, StereoCameraParams
    proj = syn.make_projection_matrix(W, H, 0.1, 100.0)
    fx, fy = syn.focal_lengths(W, H)
    lv, rv = np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    lv[3, 0], rv[3, 0] = 0.032, -0.032
    L = CameraParams(lv, proj, (-0.032, 0, 0), fx, fy, 0.1, 100.0)
    R = CameraParams(rv, proj, (0.032, 0, 0), fx, fy, 0.1, 100.0)
    return StereoCameraParams(L, R)


@pytest.mark.parametrize("precision,deg,flip", [("float16", 3, True), ("float32", 1, False)])
def test_stereo_frame_gpu(oracle, pu, precision, deg, flip):
    import torch
    from gsm_renderer_b200.renderer import (DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig,
                                            RenderPrecision, StereoRenderTarget)
    cl = syn.synthetic_cloud(30_000, deg, seed=8, scale_median=0.02)
    g, h = pu.make_scene_inputs(cl, precision)
    W, H = 640, 360
    cams = _stereo_inputs(W, H)
    ocam = oracle.make_stereo_camera(cams.leftEye.viewMatrix, cams.leftEye.projectionMatrix, cams.leftEye.position,
                                     cams.rightEye.viewMatrix, cams.rightEye.projectionMatrix, cams.rightEye.position,
                                     W, H, 0.1, 100.0, cl.sh_components, cl.count, False)
    fr = oracle.OracleFrame(cl.count, W, H, stereo=True)
    ref, _ = fr.render_stereo(g, h, oracle.F16 if precision == "float16" else oracle.F32, ocam, W, H, flip_y=flip)
    r = DepthFirstRenderer(device=0, config=RendererConfig(
        maxGaussians=cl.count, maxWidth=W, maxHeight=H,
        precision=RenderPrecision.float16 if precision == "float16" else RenderPrecision.float32,
        gaussianColorSpace=GaussianColorSpace.linear), stereoCopyFlipY=flip)
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    sbs = torch.full((H, 2 * W, 4), 0x7E00, dtype=torch.int16, device=dev)
    r.renderStereo(torch.cuda.current_stream(), StereoRenderTarget.sideBySide(sbs), GaussianInput(tg, th, cl.count, cl.sh_components),
                   cams, W, H)
    torch.cuda.synchronize()
    out = sbs.cpu().numpy().view(np.uint16)
    pu.compare_white_box(r, fr, W, H, cl.count, stereo=True)
    pu.compare_pixels(out, ref, True, precision == "float16", "stereo colour")
    r.close()


def test_repeated_frames_are_identical_gpu(pu):
    """Every frame of the same input equals the first, with back-to-back frames chained on one stream: a race in the
    publish / resolve prefix schemes, the arrival masks of the sort passes or the in-kernel clearing of the frame state
    would show up as a differing image, key list or instance list."""
    import torch
    from gsm_renderer_b200.renderer import DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig, RenderPrecision
    cl = syn.synthetic_cloud(300_000, 3, seed=23, scale_median=0.015)
    g, h = pu.make_scene_inputs(cl, "float16")
    W, H = 1920, 1080
    cam = pu.default_camera(W, H)
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=cl.count, maxWidth=W, maxHeight=H,
                                                           precision=RenderPrecision.float16, gaussianColorSpace=GaussianColorSpace.linear))
    dev = torch.device("cuda:0")
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    inp = GaussianInput(tg, th, cl.count, 16)
    s = torch.cuda.current_stream()
    ref = torch.zeros((H, W, 4), dtype=torch.int16, device=dev)
    r.render(s, ref, None, inp, cam, W, H)
    torch.cuda.synchronize()
    hd = r.debugReadHeader()
    keys = r.debugReadDepthKeys(hd.visibleCount).copy()
    inst = r.debugReadInstanceGaussianIndices(hd.totalInstances).copy()
    out = torch.zeros_like(ref)
    for i in range(60):
        r.render(s, out, None, inp, cam, W, H)
        if i % 10 == 9:
            torch.cuda.synchronize()
            assert torch.equal(out, ref), f"frame {i} differs from the first"
            h2 = r.debugReadHeader()
            assert (h2.visibleCount, h2.totalInstances) == (hd.visibleCount, hd.totalInstances)
            assert np.array_equal(r.debugReadDepthKeys(hd.visibleCount), keys)
            assert np.array_equal(r.debugReadInstanceGaussianIndices(hd.totalInstances), inst)
    r.close()


def test_sorts_on_the_streaming_path_gpu(oracle):
    """More than 1024 tiles per pass: the chained two-level look-back instead of the direct summation."""
    rng = np.random.default_rng(77)
    n = 6_000_000                                   # 1465 tiles of 4096 keys
    keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    keys[rng.integers(0, n, n // 5)] = keys[7]      # a heavy tie class: stability shows in the payload
    pay = np.arange(n, dtype=np.int32)
    k, p = _gpu_sort(keys.view(np.int32), pay, 32, 4)
    ok, op = oracle.sort_pairs_u32(keys, pay, 4)
    assert np.array_equal(k.view(np.uint32), ok) and np.array_equal(p, op)
    n = 5_000_000                                   # 1221 tiles of 4096 16-bit keys
    k16 = rng.integers(0, 8160, n).astype(np.uint16)
    pay = np.arange(n, dtype=np.int32)
    k, p = _gpu_sort(k16.view(np.int16), pay, 16, 2)
    ok, op = oracle.sort_pairs_u16(k16, pay, 2)
    assert np.array_equal(k.view(np.uint16), ok) and np.array_equal(p, op)


def test_three_million_gaussians_frame_gpu(oracle, pu):
    """N >= 3 M switches the frame's 32-bit sorts to 4096-key tiles (and the last depth pass still gathers nTouched)."""
    cl = syn.synthetic_cloud(3_100_000, 3, seed=9, scale_median=0.006)
    res = pu.run_mono_case(oracle, cl, "float16", 1280, 720)
    assert res["V"] > 1_000_000 and res["I"] > res["V"]


def test_sorts_equal_cub_library_bar_gpu():
    """cub::DeviceRadixSort::SortPairs (stable) output byte-equal to the hand-written onesweep sorts, on the
    frame's own key distributions and on adversarial ones (all equal, two values, already sorted, reversed)."""
    import ctypes as C
    import os
    import torch
    from gsm_renderer_b200.renderer import DepthFirstRenderer, RendererConfig
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bin", "libcub_bar.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    cub = C.CDLL(path)
    cub.cub_sort_pairs.restype = C.c_int
    cub.cub_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, C.c_void_p]
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=1024, maxWidth=64, maxHeight=64))
    rng = np.random.default_rng(5)
    s = torch.cuda.current_stream()

    def check(keys_np, bits, passes):
        n = keys_np.size
        keys = torch.from_numpy(keys_np.view(np.int32 if bits == 32 else np.int16)).cuda()
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        assert cub.cub_sort_pairs(keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), n, bits, 0, 8 * passes,
                                  s.cuda_stream, 0, None) == 0
        k, v = keys.clone(), vals.clone()
        r.sortPairs(s, k, v, n, bits, passes)
        torch.cuda.synchronize()
        assert torch.equal(k, ko) and torch.equal(v, vo), f"n={n} bits={bits} passes={passes}"

    for n in (1, 255, 2049, 70_000, 709_202):
        depth = rng.uniform(2.0, 20.0, n).astype(np.float32)
        check(depth.view(np.uint32) | np.uint32(0x80000000), 32, 4)
        check(rng.integers(0, 8160, n, dtype=np.int64).astype(np.uint16), 16, 2)
    n = 300_000
    check(np.full(n, 0x80001234, np.uint32), 32, 4)
    check(rng.integers(0, 2, n, dtype=np.int64).astype(np.uint32) * np.uint32(0x01000000), 32, 4)
    check(np.arange(n, dtype=np.uint32), 32, 4)
    check(np.arange(n, dtype=np.uint32)[::-1].copy(), 32, 4)
    check((np.arange(n, dtype=np.uint32) % 8160).astype(np.uint16), 16, 2)
    check(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32), 32, 2)   # partial sort: 2 passes == bits [0, 16)
    r.close()
