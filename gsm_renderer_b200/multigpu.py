"""Multi-GPU partitioning of the DepthFirst path. The reference is single-device; these are the
only places where the path shards, one process per GPU over torch.distributed:

  * multi-view batches (C5): views are independent -> round-robin over ranks, NO collective;
  * stereo (C4): both ranks run the identical joint stages 1-7 and blend one eye each (eyeMask), then the
    right-eye rank sends its half of the side-by-side target (one peer copy / send), NO collective;
  * one large frame (C3): Gaussians shard by contiguous gid range for projection, ONE all-gather of the compacted
    48-byte splat records over NCCL/NVLink (rank-major == global gid order, so stable-sort ties break exactly as on
    one GPU), then every rank sorts/expands/blends its own horizontal strip of whole tile rows.

Host logic here is backend-agnostic (gloo on CPU tensors in the tests, nccl on the box).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

RECORD_BYTES = 48


def partition_range(n: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous (first, count) shards of [0, n): sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(n, parts)
    out, first = [], 0
    for r in range(parts):
        c = base + (1 if r < rem else 0)
        out.append((first, c))
        first += c
    return out


def partition_views(n_views: int, world: int, rank: int) -> List[int]:
    """Round-robin view ids of `rank` (C5: 256 views over 8 GPUs = 32 each)."""
    return list(range(rank, n_views, world))


def partition_tile_rows(tiles_y: int, world: int, row_weights: Sequence[float] | None = None) -> List[Tuple[int, int]]:
    """Horizontal strips of whole tile rows, (rowFirst, rowCount) per rank. Balanced by row count, or by the
    given per-row weights (e.g. sum of nTouched per row from the previous frame)."""
    if row_weights is None:
        return partition_range(tiles_y, world)
    w = np.asarray(row_weights, np.float64)
    assert w.shape[0] == tiles_y
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    if total <= 0:
        return partition_range(tiles_y, world)
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), tiles_y))
    cuts.append(tiles_y)
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


def _bytes(t):
    """Contiguous uint8 copy of a tensor (NCCL has no int16; images travel as raw bytes)."""
    import torch
    return t.contiguous().view(-1).view(torch.uint8)


def all_gather_records(dist, local_records, local_count: int, world: int, device):
    """All-gather-v of compacted splat records in rank order. local_records: uint8 tensor [>= local_count*48].
    Returns (records uint8 [total*48], counts list). One small all-gather of counts + one padded all-gather."""
    import torch
    counts_t = torch.tensor([local_count], dtype=torch.int64, device=device)
    gathered = [torch.zeros_like(counts_t) for _ in range(world)]
    dist.all_gather(gathered, counts_t)
    counts = [int(c.item()) for c in gathered]
    mx = max(max(counts), 1)
    pad = torch.zeros(mx * RECORD_BYTES, dtype=torch.uint8, device=device)
    pad[: local_count * RECORD_BYTES] = local_records[: local_count * RECORD_BYTES]
    out = torch.empty(world * mx * RECORD_BYTES, dtype=torch.uint8, device=device)
    if hasattr(dist, "all_gather_into_tensor") and device.type == "cuda":
        dist.all_gather_into_tensor(out, pad)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out = torch.cat(parts)
    chunks = [out[r * mx * RECORD_BYTES: r * mx * RECORD_BYTES + counts[r] * RECORD_BYTES] for r in range(world)]
    return torch.cat(chunks) if world > 1 else chunks[0].clone(), counts


def render_strips(renderer, dist, rank: int, world: int, stream, gaussians_shard, harmonics_shard, shard: Tuple[int, int],
                  sh_components: int, camera, width: int, height: int, color, depth, records_scratch,
                  strips: List[Tuple[int, int]] | None = None):
    """One strip-sharded frame on this rank. gaussians_shard/harmonics_shard hold only gids [shard[0], +shard[1]).
    color/depth are full-size targets; this rank fills rows of its strip. Returns (records, counts, strip)."""
    import torch
    tiles_y = (height + 15) // 16
    strips = strips or partition_tile_rows(tiles_y, world)
    n_local = renderer.stripProject(stream, gaussians_shard, harmonics_shard, shard[0], shard[1], sh_components, camera,
                                    width, height, records_scratch)
    if world > 1:
        records, counts = all_gather_records(dist, records_scratch, n_local, world, records_scratch.device)
    else:
        records, counts = records_scratch[: n_local * RECORD_BYTES], [n_local]
    total = sum(counts)
    renderer.stripRender(stream, color, depth, records, total, width, height, strips[rank][0], strips[rank][1])
    return records, counts, strips[rank]


def gather_strips(dist, rank: int, world: int, color, strips: List[Tuple[int, int]], width: int, height: int, root: int = 0):
    """Assemble the strip-disjoint image on `root` (send/recv of each rank's rows; peer copies over NVLink under nccl)."""
    if world == 1:
        return color
    for r in range(world):
        if r == root:
            continue
        y0 = strips[r][0] * 16
        y1 = min(height, (strips[r][0] + strips[r][1]) * 16)
        if y1 <= y0:
            continue
        if rank == r:
            dist.send(_bytes(color[y0:y1]), dst=root)
        elif rank == root:
            buf = _bytes(color[y0:y1])
            dist.recv(buf, src=r)
            color[y0:y1] = buf.view(color.dtype).view(color[y0:y1].shape)
    return color


def render_stereo_split(renderer, dist, rank: int, world: int, stream, target, gaussian_input, cameras, width: int, height: int):
    """C4: one eye per GPU. Ranks 0/1 render eye 0/1 of the joint stereo frame; rank 1 sends its half to rank 0."""
    from .renderer import StereoRenderTarget
    if world == 1:
        renderer.renderStereo(stream, StereoRenderTarget.sideBySide(target), gaussian_input, cameras, width, height)
        return target
    eye = rank % 2
    renderer.renderStereo(stream, StereoRenderTarget.sideBySide(target), gaussian_input, cameras, width, height, eyeMask=1 << eye)
    # The right half travels as one contiguous byte buffer. With two ranks it goes as a broadcast from rank 1: a collective runs
    # on all of NCCL's channels, while a point-to-point send of the same 16.6 MB was measured at 2.2-2.8 ms per frame.
    right = target[:, width:]
    stage = getattr(renderer, "_stereo_half_stage", None)
    if stage is None or stage.numel() != right.numel() * right.element_size() or stage.device != target.device:
        import torch
        stage = torch.empty(right.numel() * right.element_size(), dtype=torch.uint8, device=target.device)
        renderer._stereo_half_stage = stage
    if rank == 1:
        stage.view(target.dtype).view(right.shape).copy_(right)
    if world == 2:
        dist.broadcast(stage, src=1)
    elif rank == 1:
        dist.send(stage, dst=0)
    elif rank == 0:
        dist.recv(stage, src=1)
    if rank == 0:
        right.copy_(stage.view(target.dtype).view(right.shape))
    return target
