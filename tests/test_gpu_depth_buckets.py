"""The round-2 sorts on the inputs that leave their fast paths. Depth sort's bucket path (csrc/bucketsort.cu): every white-box buffer
and every pixel against the oracle, whose depth sort is the reference's four 8-bit LSD passes (DFS.metal:1387-1696).

 * thousands of Gaussians inside ONE fine bin of the key histogram (a thin slab, with two outliers stretching the key range):
   the bucket exceeds one CTA's shared memory and is sorted by the streaming passes;
 * thousands of EQUAL depths (span 0: the local pass is a copy, the stable order is the gid order);
 * a bucket plan built from a sample that misses most of a cluster (keys arranged so that every 8th stored key is an outlier);
 * frames just below / above the size at which the host switches to the LSD passes;
 * the same frame with GSM_DEPTH_BUCKETS=0 (LSD passes) in a subprocess produces identical bytes."""
import os
import subprocess
import sys

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


def _slab(n, z0, dz, seed, outliers=((0.0, 0.0, 0.6), (0.0, 0.0, 80.0))):
    cl = syn.synthetic_cloud(n, 1, seed=seed, scale_median=0.01)
    rng = np.random.default_rng(seed)
    z = (z0 + rng.uniform(0.0, dz, n)).astype(np.float32)
    lat = cl.positions[:, :2] / cl.positions[:, 2:3]
    cl.positions[:, 2] = z
    cl.positions[:, :2] = lat * z[:, None]
    for i, o in enumerate(outliers):
        cl.positions[i] = o
        cl.scales[i] = 0.02
        cl.opacities[i] = 0.9
    return cl


def test_dense_fine_bin_streams_gpu(oracle, pu):
    # key range 0.6 .. 80 (7 octaves => fine bins of 8192 ulps); 40 000 distinct depths inside 5 .. 5.002 share one or two bins
    cl = _slab(40_000, 5.0, 0.002, seed=5)
    res = pu.run_mono_case(oracle, cl, "float16", 1280, 720)
    assert res["V"] > 15_000


def test_equal_depths_gpu(oracle, pu):
    cl = _slab(30_000, 6.0, 0.0, seed=6)
    res = pu.run_mono_case(oracle, cl, "float16", 1280, 720)
    assert res["V"] > 15_000
    # without outliers the key range is a single value (span 0, shift 0, one fine bin)
    cl = _slab(9_000, 6.0, 0.0, seed=7, outliers=())
    res = pu.run_mono_case(oracle, cl, "float32", 640, 360)
    assert res["V"] > 3_000


def test_sample_misses_cluster_gpu(oracle, pu):
    # every 8th gid far away, the other seven in a tight cluster: whichever residue the sample hits, the plan's boundaries
    # are far from balanced and the exact offsets must still come out right
    n = 48_000
    cl = _slab(n, 4.0, 0.01, seed=8, outliers=())
    far = np.arange(n) % 8 == 0
    rng = np.random.default_rng(9)
    z = rng.uniform(20.0, 60.0, int(far.sum())).astype(np.float32)
    lat = cl.positions[far, :2] / cl.positions[far, 2:3]
    cl.positions[far, 2] = z
    cl.positions[far, :2] = lat * z[:, None]
    res = pu.run_mono_case(oracle, cl, "float16", 1280, 720)
    assert res["V"] > 15_000


@pytest.mark.parametrize("n", [999_000, 1_001_000])
def test_switch_to_lsd_passes_gpu(oracle, pu, n):
    # kDepthBucketMaxGaussians = 1 000 000: one frame on each side of the host-side switch (small splats keep the oracle fast)
    cl = syn.synthetic_cloud(n, 0, seed=3, scale_median=0.004)
    res = pu.run_mono_case(oracle, cl, "float16", 640, 360)
    assert res["V"] > 100_000


def test_bucket_path_equals_lsd_path_gpu(tmp_path):
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "import tests.parity_util as pu\n"
        "from gsm_renderer_b200 import synthetic as syn\n"
        "cl = syn.synthetic_cloud(300_000, 2, seed=12, scale_median=0.012)\n"
        "g, h = pu.make_scene_inputs(cl, 'float16')\n"
        "cam = pu.default_camera(1920, 1080)\n"
        "r, c, d = pu.gpu_mono(g, h, 'float16', cam, 1920, 1080, cl.sh_components, cl.count, False)\n"
        "hd = r.debugReadHeader()\n"
        "np.savez(sys.argv[1], c=c, d=d, keys=r.debugReadDepthKeys(hd.visibleCount), idx=r.debugReadSortedPrimitiveIndices(hd.visibleCount),\n"
        "         off=r.debugReadInstanceOffsets(hd.visibleCount), inst=r.debugReadInstanceGaussianIndices(hd.totalInstances),\n"
        "         plan=np.array([r.debugReadDepthSortPlan()['bucketCount']]))\n"
        "r.close()\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = {}
    for mode in ("1", "0"):
        f = str(tmp_path / f"frame{mode}.npz")
        env = dict(os.environ, GSM_DEPTH_BUCKETS=mode)
        subprocess.run([sys.executable, "-c", code, f], check=True, env=env, timeout=300)
        out[mode] = np.load(f)
    assert int(out["1"]["plan"][0]) > 50 and int(out["0"]["plan"][0]) == 0   # the two runs really took different paths
    for k in ("c", "d", "keys", "idx", "off", "inst"):
        assert np.array_equal(out["1"][k], out["0"][k]), k


# ---- tile sort, most-significant digit first (csrc/tilesort.cu): every low-bit width L = bits(tileCount - 1) - 8
@pytest.mark.parametrize("W,H", [(320, 400),      # 20 x 25 = 500 tiles, L = 1
                                 (512, 512),      # 1024 tiles, L = 2 (ids use exactly 10 bits)
                                 (720, 720),      # 2025 tiles, L = 3
                                 (1280, 720),     # 3600 tiles, L = 4
                                 (2560, 1440),    # 14400 tiles, L = 6
                                 (3840, 2160),    # 32400 tiles, L = 7
                                 (4096, 4080),    # 65280 tiles, L = 8 (the largest count 16-bit ids allow is 65535)
                                 (257 * 16, 16),  # 257 tiles in one row, L = 1: a bucket of one tile at the end
                                 (4110, 17)])     # 257 x 2 tiles through the odd-size path
def test_tile_sort_low_bit_widths_gpu(oracle, pu, W, H):
    cl = syn.synthetic_cloud(30_000, 1, seed=W + H, scale_median=0.02)
    res = pu.run_mono_case(oracle, cl, "float16", W, H)
    assert res["I"] >= res["V"] > 0


def test_tile_msd_equals_lsd_path_gpu(tmp_path):
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "import tests.parity_util as pu\n"
        "from gsm_renderer_b200 import synthetic as syn\n"
        "cl = syn.synthetic_cloud(400_000, 2, seed=13, scale_median=0.02)\n"
        "g, h = pu.make_scene_inputs(cl, 'float16')\n"
        "cam = pu.default_camera(1920, 1080)\n"
        "r, c, d = pu.gpu_mono(g, h, 'float16', cam, 1920, 1080, cl.sh_components, cl.count, False)\n"
        "hd = r.debugReadHeader()\n"
        "np.savez(sys.argv[1], c=c, d=d, ids=r.debugReadSortedTileIds(hd.totalInstances), inst=r.debugReadInstanceGaussianIndices(hd.totalInstances),\n"
        "         th=r.debugReadTileHeaders(120 * 68))\n"
        "r.close()\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = {}
    for mode in ("1", "0"):
        f = str(tmp_path / f"frame{mode}.npz")
        subprocess.run([sys.executable, "-c", code, f], check=True, env=dict(os.environ, GSM_TILE_MSD=mode), timeout=300)
        out[mode] = np.load(f)
    for k in ("c", "d", "ids", "inst", "th"):
        assert np.array_equal(out["1"][k], out["0"][k]), k
