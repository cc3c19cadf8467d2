"""Sort sweep with the library bar: gsm_sort_pairs_with_scratch (histogram kernel + onesweep passes, no allocation) against
cub::DeviceRadixSort::SortPairs (tools/cub_bar.cu) on the same arrays, both timed with CUDA events around the sort alone
(the pristine input is restored before every repetition, outside the events), and compared byte for byte.
Key distributions are the frame's: 32-bit depth keys = sortable bits of depths uniform in [2, 20]; 16-bit tile ids uniform
in [0, 8160) (1080p). Prints one CSV row per (size, key type).
  python tools/sort_sweep.py [sizes...]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsm_renderer_b200.renderer import DepthFirstRenderer, RendererConfig  # noqa: E402

cub = C.CDLL(os.path.join(ROOT, "tools", "bin", "libcub_bar.so"))
cub.cub_sort_pairs.restype = C.c_int
cub.cub_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_int, C.POINTER(C.c_float)]


def cub_sort(keys, vals, key_bits, end_bit, reps):
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    ms = C.c_float(0)
    rc = cub.cub_sort_pairs(keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), keys.numel(), key_bits, 0, end_bit,
                            torch.cuda.current_stream().cuda_stream, reps, C.byref(ms))
    assert rc == 0, rc
    return ko, vo, ms.value * 1e3


def ours_sort(r, keys, vals, key_bits, passes, reps):
    n = keys.numel()
    scratch = torch.empty(r.sortPairsScratchBytes(n, key_bits, passes) + 256, dtype=torch.uint8, device="cuda")
    off = (-scratch.data_ptr()) % 256
    sp = scratch.data_ptr() + off
    s = torch.cuda.current_stream()
    k, v = keys.clone(), vals.clone()
    r.sortPairsWithScratch(s, k, v, n, key_bits, passes, sp)   # warm-up + the result that is compared
    torch.cuda.synchronize()
    kk, vv = keys.clone(), vals.clone()
    tot = 0.0
    for _ in range(reps):
        kk.copy_(keys); vv.copy_(vals)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.sortPairsWithScratch(s, kk, vv, n, key_bits, passes, sp)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return k, v, tot / reps * 1e3


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [50_000, 200_000, 709_202, 2_909_567, 12_000_000, 48_000_000]
    r = DepthFirstRenderer(device=0, config=RendererConfig(maxGaussians=1024, maxWidth=64, maxHeight=64))
    rng = np.random.default_rng(0)
    print("n,keys,ours_us,cub_us,ours_over_cub,equal,ours_GBps,cub_GBps")
    for n in sizes:
        depth = rng.uniform(2.0, 20.0, n).astype(np.float32)
        k32 = torch.from_numpy((depth.view(np.uint32) | np.uint32(0x80000000)).view(np.int32)).cuda()   # float_to_sortable_uint of a positive float
        k16 = torch.from_numpy(rng.integers(0, 8160, n, dtype=np.int64).astype(np.int16)).cuda()
        p = torch.arange(n, dtype=torch.int32, device="cuda")
        reps = 20 if n <= 3_000_000 else 5
        for name, keys, bits, passes, bytes_per in (("u32", k32, 32, 4, 68), ("u16", k16, 16, 2, 26)):
            ok, ov, ours_us = ours_sort(r, keys, p, bits, passes, reps)
            ck, cv, cub_us = cub_sort(keys, p, bits, bits, reps)
            eq = bool(torch.equal(ok, ck)) and bool(torch.equal(ov, cv))
            print(f"{n},{name},{ours_us:.1f},{cub_us:.1f},{ours_us / cub_us:.3f},{eq},{bytes_per * n / ours_us / 1e3:.0f},{bytes_per * n / cub_us / 1e3:.0f}", flush=True)
    r.close()


if __name__ == "__main__":
    main()
