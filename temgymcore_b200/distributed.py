"""Multi-GPU sharding of the hot path: one process per GPU (``torch.distributed``).

The path shards in the two ways it does naturally:

* rays / scan positions: contiguous chunks per rank, model descriptor replicated,
  NO communication (``shard_range`` + ``run_to_end_abcd`` on the local chunk);
* field sum: detector ROWS are sharded (tile-aligned blocks), every rank needs every
  beamlet -> the (nb, 12) coefficient table is broadcast from the source rank, each rank
  sums its rows, and the row blocks are all-gathered into the full image.  Sharding the
  beamlets instead would need a reduce-scatter and change the summation order.

Collectives go through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the
CPU tests, where the per-rank compute is injected).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

ROW_ALIGN = 32  # the field kernel's tile height: aligned shards keep tile origins identical


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ``[begin, end)`` of ``n`` independent units for ``rank``; sizes differ by
    at most one; every unit is owned by exactly one rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def row_shards(H: int, world: int, align: int = ROW_ALIGN):
    """Tile-aligned detector row blocks ``[(row0, nrows), ...]`` for every rank (some may be
    empty when H is small).  Blocks are whole multiples of ``align`` rows except the last."""
    nblocks = (H + align - 1) // align
    out = []
    for r in range(world):
        b0, b1 = shard_range(nblocks, r, world)
        r0, r1 = min(H, b0 * align), min(H, b1 * align)
        out.append((r0, r1 - r0))
    return out


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_table(poly, nb: int, src: int = 0, group=None):
    """Broadcast the beamlet coefficient table (96 B per beamlet) from ``src``."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(poly, src=src, group=group)
    return poly


def gather_rows(local_rows, H: int, W: int, group=None):
    """All-gather the per-rank row blocks into the full ``(H, W)`` image on every rank."""
    import torch
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    shards = row_shards(H, world)
    max_rows = max(n for _, n in shards)
    pad = torch.zeros((max_rows, W), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    # complex tensors travel as their real view (gloo / older NCCL builds)
    buf = torch.view_as_real(pad).contiguous() if pad.is_complex() else pad
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = []
    for (r0, n), t in zip(shards, out):
        t = torch.view_as_complex(t) if local_rows.is_complex() else t
        parts.append(t[:n])
    return torch.cat(parts, dim=0)


def make_gaussian_image_sharded(gaussian_rays, model, *, cull_bits=None, out_dtype=None,
                                method="auto", group=None, src: int = 0,
                                rows_fn: Optional[Callable] = None,
                                table_fn: Optional[Callable] = None):
    """Row-sharded ``make_gaussian_image`` over the ranks of ``group``.

    ``src`` traces the central rays and builds the coefficient table; the table is
    broadcast; every rank sums its row block; the blocks are all-gathered.  ``rows_fn`` /
    ``table_fn`` exist so the plumbing can be exercised on CPU with gloo.
    """
    import torch
    dist = _dist()
    from .gaussian import _field_sum_grid, beamlet_polynomials
    grid = model[-1]
    H, W = int(grid.shape[0]), int(grid.shape[1])
    inited = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if inited else 1
    rank = dist.get_rank(group) if inited else 0
    table_fn = table_fn or (lambda: beamlet_polynomials(gaussian_rays, model))
    poly, nb, dev = table_fn()          # every rank holds the inputs; src's table wins
    poly = broadcast_table(poly, nb, src=src, group=group)
    r0, nr = row_shards(H, world)[rank]
    rows_fn = rows_fn or (lambda p, n, row0, nrows: _field_sum_grid(
        p, n, grid, dev, row0=row0, nrows=nrows, out_dtype=out_dtype, cull_bits=cull_bits,
        method=method))
    local = rows_fn(poly, nb, r0, nr)
    return gather_rows(local, H, W, group=group)
