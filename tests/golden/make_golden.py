"""Regenerates tests/golden/*.json.

Two kinds of fixtures:
  * reference_kats.json -- the known-answer vectors the reference's own tests hold for this path, restated
    (DepthFirstUnitTests.swift:125-145,304 ; :309-317 ; GlobalUnitTests.swift:31-39 ; DepthFirstUnitTests.swift:21-117).
  * foveated_copy_digests.json -- SHA-256 of the drawable bytes the oracle's foveated stereo copy writes for three seeded cases
    (oracle self-regression digests, like the next one).
  * oracle_digests.json -- SHA-256 of every white-box buffer and of the pixels of a few small frames rendered by
    the CPU oracle. These are ORACLE self-regression digests (the reference pins no such values); they freeze the
    canonical semantics between rounds and let the CUDA path be checked without re-running the oracle.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gsm_renderer_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

SCENES = {
    # name: (cloud factory, precision, W, H, near, far, srgb, sh)
    "pipeline_stages_f32_640x480": (lambda: syn.pipeline_stages_scene(), "float32", 640, 480, 0.1, 10.0, True, 1),
    "synthetic_8k_sh3_f16_1280x720": (lambda: syn.synthetic_cloud(8000, 3, seed=42, scale_median=0.02), "float16", 1280, 720, 0.1, 100.0, False, 16),
    "synthetic_5k_sh1_f32_333x77": (lambda: syn.synthetic_cloud(5000, 1, seed=7, scale_median=0.03), "float32", 333, 77, 0.1, 100.0, True, 4),
    "grid_fixture_1500_f16_800x600": (lambda: syn.generate_grid_gaussians(1500, 42), "float16", 800, 600, 0.1, 10.0, True, 0),
}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_digests(header, nTouched, bounds, depthKeys, primIdx, offsets, tileIds, instIdx, tileHeaders, activeTiles, color, depth):
    V, I = header["visibleCount"], header["totalInstances"]
    return {
        "header": header,
        "nTouched": sha(nTouched), "bounds": sha(bounds), "depthKeys": sha(depthKeys[:V]), "primitiveIndices": sha(primIdx[:V]),
        "instanceOffsets": sha(offsets[:V]), "sortedTileIds": sha(tileIds[:I].astype(np.uint32)),
        "instanceGaussianIndices": sha(instIdx[:I]), "tileHeaders": sha(tileHeaders),
        "activeTilesSorted": sha(np.sort(activeTiles)), "color": sha(color), "depth": sha(depth),
    }


def oracle_digests():
    from oracle import binding as ob
    ob.build()
    out = {}
    for name, (mk, prec, W, H, near, far, srgb, sh) in SCENES.items():
        cl = mk()
        g, h = cl.pack(prec)
        proj = syn.make_projection_matrix(W, H, near, far)
        cam = ob.make_camera(np.eye(4), proj, (0, 0, 0), W, H, near, far, sh, cl.count, srgb)
        fr = ob.OracleFrame(cl.count, W, H)
        color, depth = fr.render_mono(g, h, ob.F16 if prec == "float16" else ob.F32, cam, W, H)
        T = ((W + 15) // 16) * ((H + 15) // 16)
        hd = fr.header
        header = {k: int(getattr(hd, k)) for k in ("visibleCount", "totalInstances", "paddedVisibleCount", "paddedInstanceCount", "overflow")}
        out[name] = frame_digests(header, fr.nTouched[:cl.count], fr.bounds[:cl.count], fr.depthKeys, fr.primitiveIndices,
                                  fr.orderedTileCounts, fr.instanceTileIds, fr.instanceGaussianIndices, fr.tileHeaders[:T],
                                  fr.activeTiles[:fr.f.activeTileCount], color, depth)
    return out


COPY_CASES = {
    # name: (W, H, format, arrayLength, flip, texture (w, h) or None = from the rate map, viewports, rate-map cell rates or None)
    "sbs_rgba16f_1to1": (96, 54, 0, 1, True, (192, 54), ((0, 0, 96, 54), (96, 0, 96, 54)), None),
    "layered_bgra8_srgb_ratemap": (120, 68, 2, 2, True, None, ((0, 0, 120, 68), (0, 0, 120, 68)), ((0.25, 0.5, 1.0, 1.0, 0.5, 0.25), (0.3, 0.75, 1.0, 0.6))),
    "shared_rgba8_scaled_overlap": (77, 41, 3, 1, False, (150, 60), ((3.25, 2.5, 90.5, 50.75), (60, 5, 80, 44)), None),
}


def copy_case_inputs(name):
    """Seeded intermediate image + drawable description of one COPY_CASES entry (shared by the CPU and GPU tests)."""
    sys.path.insert(0, os.path.join(ROOT, "tests")) if os.path.join(ROOT, "tests") not in sys.path else None
    import foveation_util as fv
    W, H, fmt, array_length, flip, tex, vps, rates = COPY_CASES[name]
    rng = np.random.default_rng(sum(name.encode()))
    c2 = (rng.uniform(-0.1, 1.1, (2, H, W, 4)) if fmt else rng.standard_normal((2, H, W, 4)) * 3).astype(np.float16).view(np.uint16)
    layers = None
    if rates:
        sw, sh_ = max(v[0] + v[2] for v in vps), max(v[1] + v[3] for v in vps)
        layers = [fv.layer(sw, sh_, rates[0], rates[1])]
        tex = (layers[0][0].size, layers[0][1].size)
    return c2, W, H, fmt, array_length, flip, tex, vps, layers


def copy_digests():
    from oracle import binding as ob
    ob.build()
    out = {}
    for name in COPY_CASES:
        c2, W, H, fmt, array_length, flip, tex, vps, layers = copy_case_inputs(name)
        out[name] = sha(ob.stereo_copy_foveated(c2, flip, tex[0], tex[1], array_length, fmt, vps, rate_layers=layers))
    return out


def reference_kats():
    r = syn.Drand48(42)
    keys = []
    for _ in range(1024):
        tile = int(r() * 10)
        depth = np.float32(r() * 100.0)
        keys.append((tile << 16) | ((int(np.float16(depth).view(np.uint16)) ^ 0x8000) & 0xFFFF))
    keys = np.array(keys, np.uint32)
    order = np.argsort(keys, kind="stable")
    i = np.arange(1_000_000, dtype=np.int64)
    big = ((i * 37 + 12345) & 0xFFFF).astype(np.uint32)
    bo = np.argsort(big, kind="stable")
    return {
        "depthSortSimple": {"keys": list(range(10, 0, -1)), "payload": [i * 100 for i in range(10)],
                            "sortedPayload": [900, 800, 700, 600, 500, 400, 300, 200, 100, 0]},
        "depthSortAtScale": {"n": 1_000_000, "recipe": "(i*37+12345)&0xFFFF", "sortedKeysSha": sha(big[bo]),
                             "stablePayloadSha": sha(bo.astype(np.int32))},
        "globalRadixRecipe": {"seed": 42, "n": 1024, "first8Keys": [int(k) for k in keys[:8]], "sortedKeysSha": sha(keys[order]),
                              "stablePayloadSha": sha(order.astype(np.int32))},
        "pipelineStages": {"assert": "overflow==0, 0<visibleCount<=1000, totalInstances>0"},
    }


# ---- GlobalRenderer frames: digests of the oracle's GlobalRenderer restatement on the same scenes
GLOBAL_SCENES = ("pipeline_stages_f32_640x480", "synthetic_8k_sh3_f16_1280x720", "synthetic_5k_sh1_f32_333x77")


def global_frame_digests(header, bounds, visibleIndices, sortedKeys, sortedIndices, tileHeaders, activeTiles, color, depth):
    V, A = header["visibleCount"], header["totalAssignments"]
    return {"header": header, "bounds": sha(bounds), "visibleIndices": sha(visibleIndices[:V]), "sortedKeys": sha(sortedKeys[:A]),
            "sortedIndices": sha(sortedIndices[:A]), "tileHeaders": sha(tileHeaders),
            "activeTilesSorted": sha(np.sort(np.asarray(activeTiles).astype(np.uint32))), "color": sha(color), "depth": sha(depth)}


def global_digests():
    from oracle import binding as ob
    ob.build()
    out = {}
    for name in GLOBAL_SCENES:
        mk, prec, W, H, near, far, srgb, sh = SCENES[name]
        cl = mk()
        g, h = cl.pack(prec)
        proj = syn.make_projection_matrix(W, H, near, far)
        cam = ob.make_camera(np.eye(4), proj, (0, 0, 0), W, H, near, far, sh, cl.count, srgb)
        fr = ob.OracleGlobalFrame(cl.count, W, H)
        color, depth = fr.render(g, h, ob.F16 if prec == "float16" else ob.F32, cam, W, H)
        info = fr.info
        header = {k: int(getattr(info, k)) for k in ("visibleCount", "totalAssignments", "paddedCount", "overflow", "activeTileCount")}
        active = np.nonzero(fr.tileHeaders[:, 1] > 0)[0]
        out[name] = global_frame_digests(header, fr.bounds[:cl.count], fr.visibleIndices, fr.sortedKeys, fr.sortedIndices,
                                         fr.tileHeaders, active, color, depth)
    return out


if __name__ == "__main__":
    json.dump(reference_kats(), open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1, sort_keys=True)
    json.dump(oracle_digests(), open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1, sort_keys=True)
    json.dump(copy_digests(), open(os.path.join(HERE, "foveated_copy_digests.json"), "w"), indent=1, sort_keys=True)
    json.dump(global_digests(), open(os.path.join(HERE, "global_digests.json"), "w"), indent=1, sort_keys=True)
    print("wrote", HERE)
