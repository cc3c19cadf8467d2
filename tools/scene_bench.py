"""Scene-ingest measurement: decodes a synthetic standard 3DGS .ply (SH3, 248 B per vertex) on
the device and on the CPU oracle, and times the Morton pre-sort. Prints one JSON line. Run on the GPU box:
    python tools/scene_bench.py [N]
`decode_gbps_device` counts file bytes / time of the two decode kernels alone (CUDA events around gsm_ply_load minus the
H2D copy is not separable from outside, so the whole call is reported as `load_ms`, and the kernels are listed by ncu in
profiles/)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import ply_util as pu
from gsm_renderer_b200.scene import PLYLoader, GaussianSceneBuilder
from gsm_renderer_b200.renderer import RenderPrecision

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
props, cols = pu.standard_scene(n, 3, seed=1)
data = pu.write_ply(props, cols)
pinned = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
host = pinned.numpy()
for _ in range(2):
    ds = PLYLoader.load(host, precision=RenderPrecision.float16)
torch.cuda.synchronize()
t = []
for _ in range(5):
    t0 = time.perf_counter(); ds = PLYLoader.load(host, precision=RenderPrecision.float16); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
load_ms = 1e3 * min(t)
tm = []
for _ in range(3):
    ds = PLYLoader.load(host, precision=RenderPrecision.float16)
    torch.cuda.synchronize(); t0 = time.perf_counter(); GaussianSceneBuilder.sortByMortonCode(ds); torch.cuda.synchronize(); tm.append(time.perf_counter() - t0)
from oracle import binding as ob
ob.build()
t0 = time.perf_counter(); ref = ob.ply_load(data); cpu_load = time.perf_counter() - t0
t0 = time.perf_counter(); codes, order = ob.morton_order(ref["pos"]); g = ob.pack_gaussians(ref, True, order); cpu_morton = time.perf_counter() - t0
same = bool(np.array_equal(ds.gaussians.cpu().numpy(), g))
print(json.dumps({"workload": f"standard 3DGS PLY, {n} vertices, SH3 float32 properties, {len(data) / 1e6:.1f} MB", "load_ms": load_ms,
                  "load_GBps_incl_h2d": len(data) / (load_ms * 1e-3) / 1e9, "morton_sort_ms": 1e3 * min(tm),
                  "cpu_oracle_load_ms": 1e3 * cpu_load, "cpu_oracle_morton_pack_ms": 1e3 * cpu_morton, "cpu_cores": 1,
                  "matches_oracle": same, "count": ds.count, "boundsRadius": ds.boundsRadius}))
