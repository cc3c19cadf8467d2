"""Launches gsm_sort_pairs at several sizes (run under `ncu --metrics gpu__time_duration.sum` to read the
per-kernel durations: fixed latency vs throughput of the onesweep passes)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsm_renderer_b200.renderer import DepthFirstRenderer, RendererConfig
r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=1024, maxWidth=64, maxHeight=64))
s = torch.cuda.current_stream()
rng = np.random.default_rng(0)
SIZES = [int(x) for x in sys.argv[1:]] or [50_000, 200_000, 709_000, 2_900_000, 12_000_000, 48_000_000]
for n in SIZES:
    k32 = torch.from_numpy(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.int32)).cuda()
    p = torch.arange(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        kk = k32.clone(); pp = p.clone()
        r.sortPairs(s, kk, pp, n, 32, 4)
    k16 = torch.from_numpy(rng.integers(0, 8160, n, dtype=np.int64).astype(np.int16)).cuda()
    for _ in range(2):
        kk = k16.clone(); pp = p.clone()
        r.sortPairs(s, kk, pp, n, 16, 2)
    torch.cuda.synchronize()
    print("done", n)
