"""Multi-GPU sharding of the two hot paths: one process per GPU, cells split across ranks, no data-path collective.

Every grid cell (lat x lon x member) is independent in both paths - the reference parallelises over spatial Dask
blocks only (hdp/threshold.py:161, hdp/metric.py:444) - so rank ``r`` of ``world`` owns the contiguous range
:func:`cell_range` of the flattened cell index (latitude bands of a gridded field, whole members of an ensemble),
runs thresholds -> metrics on it with the thresholds staying on its GPU, and only the finished outputs cross
NVLink in :func:`gather_cells` (one ``all_gather`` over NCCL; ``gloo`` on CPU tensors in the tests).
The small index tables (window rows, day-of-year map, season ranges, definitions) are replicated.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

CELL_ALIGN = 32          # shard boundaries fall on multiples of 32 cells: one 128-byte line of a [T, C] float32 row


def cell_range(n_cells: int, rank: int, world: int, align: int = CELL_ALIGN) -> Tuple[int, int]:
    """``[c0, c1)`` owned by ``rank``: ``ceil(n_cells / align)`` blocks of ``align`` cells dealt out as evenly as
    possible, lower ranks first.  Ranges are contiguous, disjoint, ordered by rank and cover ``[0, n_cells)``."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    if n_cells < 0 or align <= 0:
        raise ValueError("bad n_cells / align")
    blocks = (n_cells + align - 1) // align
    per, extra = divmod(blocks, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(b0 * align, n_cells), min(b1 * align, n_cells)


def all_ranges(n_cells: int, world: int, align: int = CELL_ALIGN) -> List[Tuple[int, int]]:
    return [cell_range(n_cells, r, world, align) for r in range(world)]


def shard_cells(x, rank: int, world: int, dim: int = -1, align: int = CELL_ALIGN):
    """View of this rank's cells of a tensor / ndarray whose ``dim`` is the flattened cell axis."""
    c0, c1 = cell_range(x.shape[dim], rank, world, align)
    index = [slice(None)] * x.ndim
    index[dim] = slice(c0, c1)
    return x[tuple(index)]


def gather_cells(local, n_cells: int, dim: int = -1, group=None, align: int = CELL_ALIGN):
    """All ranks' shards concatenated along the cell axis ``dim`` (every rank gets the full array).

    ``local`` is this rank's shard (a torch tensor on the backend's device: CUDA for NCCL, CPU for gloo).  Shards may
    differ in size by one block, so they travel padded to the largest and are trimmed on arrival."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if local.shape[dim] != n_cells:
            raise ValueError("single-process gather: the local shard must be the whole array")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = all_ranges(n_cells, world, align)
    c0, c1 = ranges[rank]
    if local.shape[dim] != c1 - c0:
        raise ValueError(f"rank {rank} holds {local.shape[dim]} cells, expected {c1 - c0}")
    widest = max(b - a for a, b in ranges)
    even = all(b - a == widest for a, b in ranges)
    dim = dim % local.dim()
    if dim == 0 and even and local.is_contiguous():
        # cell-major shards of equal size (thresholds [C, n_doy, P]): the gathered buffer IS the result, no repacking
        as_bytes = local.view(torch.uint8) if local.dtype == torch.uint16 else local
        bucket = torch.empty((world * as_bytes.shape[0],) + tuple(as_bytes.shape[1:]), dtype=as_bytes.dtype, device=as_bytes.device)
        dist.all_gather_into_tensor(bucket, as_bytes, group=group)
        return bucket.view(torch.uint16) if local.dtype == torch.uint16 else bucket
    if dim == local.dim() - 1 and local.is_contiguous():
        # cell-minor shards (metrics [4, P, D, Y, C]): every rank's block travels as it lies in memory ([rows, cells_r]); the
        # blocks are then laid side by side along the cell axis with one strided copy each - no transposition, no padding pass
        rows = 1
        for n in local.shape[:-1]:
            rows *= int(n)
        flat = local.reshape(rows * (c1 - c0))
        as_bytes = flat.view(torch.uint8) if flat.dtype == torch.uint16 else flat
        per = rows * (2 if local.dtype == torch.uint16 else 1)       # bucket elements per cell (every rank agrees, empty shards too)
        bucket = torch.empty((world, widest * per), dtype=as_bytes.dtype, device=as_bytes.device)
        if even:
            dist.all_gather_into_tensor(bucket.view(-1), as_bytes, group=group)
        else:
            mine = bucket[rank]
            mine[: as_bytes.numel()].copy_(as_bytes)
            dist.all_gather_into_tensor(bucket.view(-1), mine.clone(), group=group)
        out = torch.empty(tuple(local.shape[:-1]) + (n_cells,), dtype=local.dtype, device=local.device)
        out2 = out.view(rows, n_cells)
        for r, (a, b) in enumerate(ranges):
            blk = bucket[r, : (b - a) * per]
            blk = blk.view(torch.uint16) if local.dtype == torch.uint16 else blk
            out2[:, a:b].copy_(blk.view(rows, b - a))
        return out
    moved = local.movedim(dim, 0).contiguous()
    if moved.shape[0] < widest:
        pad = torch.zeros((widest - moved.shape[0],) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
        moved = torch.cat([moved, pad], dim=0)
    as_bytes = moved.view(torch.uint8) if moved.dtype == torch.uint16 else moved     # NCCL/gloo have no uint16 type
    bucket = torch.empty((world * as_bytes.shape[0],) + tuple(as_bytes.shape[1:]), dtype=as_bytes.dtype, device=as_bytes.device)
    dist.all_gather_into_tensor(bucket, as_bytes, group=group)
    if moved.dtype == torch.uint16:
        bucket = bucket.view(torch.uint16)
    bucket = bucket.view((world, widest) + tuple(moved.shape[1:]))
    parts = [bucket[r, : b - a] for r, (a, b) in enumerate(ranges)]
    return torch.cat(parts, dim=0).movedim(0, dim)


def gather_shards(local, n_cells: int, dim: int = -1, group=None, align: int = CELL_ALIGN):
    """The NVLink part of the gather alone: every rank's shard as it lies in memory, side by side in ONE buffer
    ``[world, widest shard (flattened)]`` plus the cell ranges - no repacking pass on the device.  This is all a consumer
    needs that assembles the result elsewhere (each shard is a column block ``[..., c0:c1]`` of the full array: a strided
    device-to-host copy or a view per shard); :func:`gather_cells` adds the repacking into one dense array."""
    import torch
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    ranges = all_ranges(n_cells, world, align)
    dim = dim % local.dim()
    if local.shape[dim] != ranges[rank][1] - ranges[rank][0]:
        raise ValueError(f"rank {rank} holds {local.shape[dim]} cells, expected {ranges[rank][1] - ranges[rank][0]}")
    flat = local.contiguous().reshape(-1)
    as_bytes = flat.view(torch.uint8) if flat.dtype == torch.uint16 else flat          # NCCL/gloo have no uint16 type
    per_cell = as_bytes.numel() // max(local.shape[dim], 1) if local.shape[dim] else 0
    if not local.shape[dim]:                                                            # empty shard: derive the size from the shape
        per_cell = (2 if local.dtype == torch.uint16 else 1)
        for i, n in enumerate(local.shape):
            if i != dim:
                per_cell *= int(n)
    widest = max(b - a for a, b in ranges)
    bucket = torch.empty((world, widest * per_cell), dtype=as_bytes.dtype, device=as_bytes.device)
    if not multi:
        bucket[0].copy_(as_bytes)
        return bucket, ranges
    if as_bytes.numel() == widest * per_cell:
        dist.all_gather_into_tensor(bucket.view(-1), as_bytes, group=group)
    else:
        mine = torch.zeros(widest * per_cell, dtype=as_bytes.dtype, device=as_bytes.device)
        mine[: as_bytes.numel()].copy_(as_bytes)
        dist.all_gather_into_tensor(bucket.view(-1), mine, group=group)
    return bucket, ranges


def member_pieces(c0: int, c1: int, grid_cells: int) -> List[Tuple[int, int, int]]:
    """The flattened (member, grid cell) range ``[c0, c1)`` as ``(member, g0, g1)`` pieces, one per member it touches
    (an ensemble sharded by flattened cell index: every piece meets the member-free thresholds of grid cells g0..g1)."""
    pieces = []
    c = c0
    while c < c1:
        m, g0 = divmod(c, grid_cells)
        g1 = min(grid_cells, g0 + (c1 - c))
        pieces.append((m, g0, g1))
        c += g1 - g0
    return pieces


def run_sharded(base_tc, run_tc, window_tables, percentiles, doy_map, defs, season_north, season_south, is_south=None,
                group=None, gather: bool = True, kernels=None, local_of: Optional[int] = None, out=None):
    """Both paths on this rank's cells, then (optionally) the output gather.

    ``base_tc`` / ``run_tc`` are the FULL ``[T, C]`` float32 arrays as seen by this rank (views are taken, nothing is
    copied).  With ``local_of = C`` they are instead this rank's shard of a ``C``-cell problem that no rank holds whole
    (``is_south`` local as well); ``out = (thresholds, metrics)`` are optional preallocated local outputs.  Returns ``(thresholds [C, n_doy, P] float64,
    metrics [4, P, D, Y, C] uint16)``.  ``kernels`` defaults to the CUDA entry points of :mod:`hdp_b200._core`; the
    CPU tests inject a stand-in so that the partition/gather logic is exercised without a GPU."""
    import torch.distributed as dist

    if kernels is None:
        from . import _core
        kernels = (_core.thresholds_array, _core.metrics_array)
    thresholds_fn, metrics_fn = kernels
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    if run_tc.shape[1] != base_tc.shape[1]:
        raise ValueError("baseline and measure must cover the same cells")
    if local_of is None:
        C = base_tc.shape[1]
        c0, c1 = cell_range(C, rank, world)
        south = None if is_south is None else is_south[c0:c1]
        base_tc, run_tc = base_tc[:, c0:c1], run_tc[:, c0:c1]
    else:
        C = int(local_of)
        c0, c1 = cell_range(C, rank, world)
        if base_tc.shape[1] != c1 - c0:
            raise ValueError(f"rank {rank} holds {base_tc.shape[1]} cells, expected {c1 - c0}")
        south = is_south
    kw_t = {} if out is None else {"out": out[0]}
    kw_m = {} if out is None else {"out": out[1]}
    thr = thresholds_fn(base_tc, window_tables, percentiles, **kw_t)
    met = metrics_fn(run_tc, thr, doy_map, defs, season_north, season_south, south, **kw_m)
    if gather and multi:
        thr = gather_cells(thr, C, dim=0, group=group)
        met = gather_cells(met, C, dim=-1, group=group)
    return thr, met
