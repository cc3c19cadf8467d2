"""Host-side logic of the multi-GPU partitioning on CPU: pure partition functions, and the
all-gather-v of splat records over a world_size-2 gloo group (the N>1 path without GPUs)."""
import os
import socket

import numpy as np
import pytest

from gsm_renderer_b200 import multigpu as mg


def test_partition_range_and_views():
    for n, p in [(10, 3), (6_000_000, 8), (5, 8), (0, 2), (135, 4)]:
        parts = mg.partition_range(n, p)
        assert sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(p - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    v = [mg.partition_views(256, 8, r) for r in range(8)]
    assert all(len(x) == 32 for x in v) and sorted(sum(v, [])) == list(range(256))


def test_partition_tile_rows_weighted():
    strips = mg.partition_tile_rows(135, 4)
    assert strips == [(0, 34), (34, 34), (68, 34), (102, 33)]
    w = np.zeros(68)
    w[30:40] = 1.0  # all the work in the middle rows
    strips = mg.partition_tile_rows(68, 4, w)
    assert sum(c for _, c in strips) == 68 and strips[0][0] == 0
    assert all(strips[i][0] + strips[i][1] == strips[i + 1][0] for i in range(3))
    loads = [w[a:a + c].sum() for a, c in strips]
    assert max(loads) <= 4.0  # balanced by weight, not by row count
    assert mg.partition_tile_rows(68, 4, np.zeros(68)) == mg.partition_range(68, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = counts[rank]
        local = torch.from_numpy(rng.integers(0, 256, (n + 7) * mg.RECORD_BYTES, dtype=np.uint8))  # trailing garbage
        rec, got = mg.all_gather_records(dist, local, n, world, torch.device("cpu"))
        # strips gather: each rank owns rows of an image
        H, W = 40, 8
        strips = mg.partition_tile_rows((H + 15) // 16, world)
        img = torch.full((H, W, 4), float(rank + 1))
        out = mg.gather_strips(dist, rank, world, img, strips, W, H, root=0)
        q.put((rank, got, rec.numpy().tobytes(), local[: n * mg.RECORD_BYTES].numpy().tobytes(),
               out.numpy().copy() if rank == 0 else None, strips))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[5, 3], [0, 4], [7, 0]])
def test_all_gather_records_gloo_world2(counts):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = res[0][3] + res[1][3]  # rank-major concatenation == global gid order
    for r in (0, 1):
        assert res[r][1] == counts
        assert res[r][2] == expect
    img, strips = res[0][4], res[0][5]
    for r, (a, c) in enumerate(strips):
        y0, y1 = a * 16, min(40, (a + c) * 16)
        assert np.all(img[y0:y1] == r + 1)
