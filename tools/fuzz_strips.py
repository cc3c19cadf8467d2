"""Randomised check of the strip-sharded frame (emulated ranks on one GPU): random scenes, cameras, shard counts and UNEVEN
strip partitions; the assembled image and depth must equal the single-GPU frame bit for bit.
`python tools/fuzz_strips.py [cases] [seed]`"""
import math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200 import multigpu as mg, synthetic as syn
from gsm_renderer_b200.renderer import DepthFirstRenderer, GaussianColorSpace, GaussianInput, RendererConfig, RenderPrecision
import tests.parity_util as pu

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 99)
dev = torch.device("cuda:0")
fails = 0
for c in range(cases):
    n = int(rng.choice([300, 5000, 40000, 120000]))
    W, H = [(1920, 1080), (1280, 720), (333, 77), (640, 360)][int(rng.integers(0, 4))]
    scale = float(np.exp(rng.uniform(math.log(0.003), math.log(0.05))))
    cl = syn.synthetic_cloud(n, 3, seed=int(rng.integers(0, 1 << 30)), scale_median=scale, scale_sigma=float(rng.uniform(0.3, 1.3)))
    if rng.integers(0, 2):
        eye = rng.normal(0, 4, 3) + np.array([0, 0, 11.0])
        view, pos = syn.look_at_opencv(eye, np.array([0, 0, 11.0]) + rng.normal(0, 2, 3)), tuple(np.asarray(eye, np.float32))
    else:
        view, pos = None, (0.0, 0.0, 0.0)
    cam = pu.default_camera(W, H, 0.1, 100.0, view, pos)
    g, h = pu.make_scene_inputs(cl, "float16")
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=n, maxWidth=W, maxHeight=H, precision=RenderPrecision.float16,
                                                           gaussianColorSpace=GaussianColorSpace.linear))
    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev); th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    s = torch.cuda.current_stream()
    ref_c = torch.zeros((H, W, 4), dtype=torch.int16, device=dev); ref_d = torch.zeros((H, W), dtype=torch.int16, device=dev)
    r.render(s, ref_c, ref_d, GaussianInput(tg, th, n, 16), cam, W, H)
    torch.cuda.synchronize()
    if r.debugReadHeader().overflow:
        r.close(); print(c, "skip (4N overflow: truncation is per strip)"); continue
    world = int(rng.integers(1, 7))
    cuts = sorted(set(int(x) for x in rng.integers(1, n, world - 1))) if n > 1 and world > 1 else []
    bounds = [0] + cuts + [n]
    recs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        cnt = b - a
        scratch = torch.zeros(max(cnt, 1) * mg.RECORD_BYTES, dtype=torch.uint8, device=dev)
        k = r.stripProject(s, tg[a * 32:b * 32], th[a * 96:b * 96], a, cnt, 16, cam, W, H, scratch)
        recs.append(scratch[: k * mg.RECORD_BYTES].clone())
    allrec = torch.cat(recs) if recs else torch.zeros(0, dtype=torch.uint8, device=dev)
    total = allrec.numel() // mg.RECORD_BYTES
    tilesY = (H + 15) // 16
    nstrips = int(rng.integers(1, min(7, tilesY) + 1))
    rcuts = sorted(set(int(x) for x in rng.integers(1, tilesY, nstrips - 1))) if tilesY > 1 and nstrips > 1 else []
    rb = [0] + rcuts + [tilesY]
    out_c = torch.full((H, W, 4), 0x7E00, dtype=torch.int16, device=dev); out_d = torch.full((H, W), 0x7E00, dtype=torch.int16, device=dev)
    for a, b in zip(rb[:-1], rb[1:]):
        r.stripRender(s, out_c, out_d, allrec, total, W, H, a, b - a)
    torch.cuda.synchronize()
    ok = bool(torch.equal(out_c, ref_c) and torch.equal(out_d, ref_d))
    fails += 0 if ok else 1
    print(c, "ok" if ok else "FAIL", dict(n=n, W=W, H=H, scale=round(scale, 4), shards=len(bounds) - 1, strips=len(rb) - 1, records=total), flush=True)
    r.close()
print(f"{cases} cases, {fails} failures")
sys.exit(1 if fails else 0)
