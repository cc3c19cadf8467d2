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


def _require_current_stream(stream) -> None:
    """The torch / NCCL work of the library-collective paths below runs on torch's CURRENT stream, the renderer's kernels on
    `stream`: they are only ordered if those are the same stream."""
    import torch
    if not torch.cuda.is_available() or stream is None:
        return
    cur = torch.cuda.current_stream()
    h = int(stream.cuda_stream) if hasattr(stream, "cuda_stream") else int(stream)
    if h != int(cur.cuda_stream):
        raise ValueError("pass torch.cuda.current_stream(): the collectives of this path run on torch's current stream")


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
    _require_current_stream(stream)
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
    _require_current_stream(stream)
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


# ---------------------------------------------------------------------------------------------------------------
# gsm_group: the strip-sharded frame over NVLink peer memory (include/gsm/gsm.h, csrc/group.cu). The exchange is done
# by the routing kernel's own stores into the peers' windows; torch.distributed only carries the 64-byte IPC handles.
class _DevArray:
    """A device pointer as something torch.as_tensor() can wrap without copying (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def strip_row_starts(strips: Sequence[Tuple[int, int]]) -> List[int]:
    """(rowFirst, rowCount) per rank -> the world+1 boundaries gsm_group_* take."""
    out = [s[0] for s in strips] + [strips[-1][0] + strips[-1][1]]
    assert out[0] == 0 and all(a <= b for a, b in zip(out, out[1:]))
    return out


class RendererGroup:
    """One rank of a strip-sharded frame (gsm_group). The renderer's maxGaussians must cover the whole scene."""

    HANDLE_BYTES = 64

    def __init__(self, renderer, rank: int, world: int, max_records_per_source: int, image_color_bytes: int = 0,
                 image_depth_bytes: int = 0):
        import ctypes as C
        from . import _native as N
        from .renderer import _check
        self._C, self._N, self._check = C, N, _check
        self._lib = N.lib()
        self.renderer, self.rank, self.world = renderer, int(rank), int(world)
        self.image_color_bytes, self.image_depth_bytes = int(image_color_bytes), int(image_depth_bytes)
        h = C.c_void_p()
        _check(self._lib.gsm_group_create(renderer._h, self.rank, self.world, int(max_records_per_source), self.image_color_bytes,
                                          self.image_depth_bytes, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gsm_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self) -> bytes:
        buf = self._C.create_string_buffer(self.HANDLE_BYTES)
        self._check(self._lib.gsm_group_export(self._h, buf))
        return bytes(buf.raw)

    def connect(self, handles: Sequence[bytes]) -> None:
        assert len(handles) == self.world and all(len(h) == self.HANDLE_BYTES for h in handles)
        self._check(self._lib.gsm_group_connect(self._h, self._C.create_string_buffer(b"".join(handles), self.world * self.HANDLE_BYTES)))

    def connect_local(self, peers: Sequence["RendererGroup"]) -> None:
        arr = (self._C.c_void_p * self.world)(*[p._h for p in peers])
        self._check(self._lib.gsm_group_connect_local(self._h, arr))

    def connect_distributed(self, dist) -> None:
        """Exchange the IPC handles over torch.distributed (any backend) and map every peer's window."""
        if self.world == 1:
            return
        handles = [None] * self.world
        dist.all_gather_object(handles, self.export())
        self.connect(handles)

    def image_ptrs(self, of_rank: int):
        c, d = self._C.c_void_p(), self._C.c_void_p()
        self._check(self._lib.gsm_group_image(self._h, int(of_rank), self._C.byref(c), self._C.byref(d)))
        return c.value, d.value

    def image_tensors(self, of_rank: int, width: int, height: int, device, depth: bool = True):
        """(colour int16 [H, W, 4], depth int16 [H, W] or None) views of rank `of_rank`'s window image, as mapped here."""
        import torch
        c, d = self.image_ptrs(of_rank)
        assert c and width * height * 8 <= self.image_color_bytes
        color = torch.as_tensor(_DevArray(c, (height, width, 4), "<i2"), device=device)
        dep = None
        if depth and d and width * height * 2 <= self.image_depth_bytes:
            dep = torch.as_tensor(_DevArray(d, (height, width), "<i2"), device=device)
        return color, dep

    def _rows(self, row_starts):
        assert len(row_starts) == self.world + 1
        return (self._C.c_uint32 * (self.world + 1))(*[int(v) for v in row_starts])

    def projectRoute(self, stream, gaussians_shard, harmonics_shard, gid_first: int, gid_count: int, sh_components: int, camera,
                     width: int, height: int, row_starts) -> None:
        N = self._N
        cam = camera.to_native()
        self._check(self._lib.gsm_group_project_route(self._h, N.stream_handle(stream), N.ptr(gaussians_shard), N.ptr(harmonics_shard),
                                                      int(gid_first), int(gid_count), int(sh_components), self._C.byref(cam),
                                                      int(width), int(height), self._rows(row_starts)))

    def renderStrip(self, stream, color, depth, width: int, height: int, row_starts) -> None:
        N = self._N
        self._check(self._lib.gsm_group_render_strip(self._h, N.stream_handle(stream), N.ptr(color), N.ptr(depth), int(width),
                                                     int(height), self._rows(row_starts)))
        self.renderer._stereo_last = False

    def renderStrips(self, stream, color, depth, gaussians_shard, harmonics_shard, gid_first: int, gid_count: int,
                     sh_components: int, camera, width: int, height: int, row_starts) -> None:
        N = self._N
        cam = camera.to_native()
        self._check(self._lib.gsm_render_strips(self._h, N.stream_handle(stream), N.ptr(color), N.ptr(depth), N.ptr(gaussians_shard),
                                                N.ptr(harmonics_shard), int(gid_first), int(gid_count), int(sh_components),
                                                self._C.byref(cam), int(width), int(height), self._rows(row_starts)))
        self.renderer._stereo_last = False

    def recordCounts(self, stream=None) -> List[int]:
        """Records each source rank routed to this rank in the last frame (diagnostic; synchronises the stream)."""
        out = (self._C.c_uint32 * self.world)()
        self._check(self._lib.gsm_group_record_counts(self._h, self._N.stream_handle(stream), out))
        return [int(v) for v in out]

    def signal(self, stream, to_rank: int, frame_id: int) -> None:
        self._check(self._lib.gsm_group_signal(self._h, self._N.stream_handle(stream), int(to_rank), int(frame_id) & 0xFFFFFFFF))

    def wait(self, stream, from_mask: int, frame_id: int) -> None:
        self._check(self._lib.gsm_group_wait(self._h, self._N.stream_handle(stream), int(from_mask), int(frame_id) & 0xFFFFFFFF))
