"""Event sharding over the GPUs of one node.

Every processor of a dspeed chain works on one waveform at a time (each gufunc core
dimension lies inside a row, reference processors/__init__.py:47-59), so a table of N events
shards into contiguous row ranges with NO exchange on the hot path: one process per GPU
(``torch.distributed``, NCCL on GPUs / gloo in the CPU tests), constants replicated, and a
single gather of the output tables at the end.  The reference has no
counterpart: it scales by the user launching one process per file or channel.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import tables


def shard_rows(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous, balanced row range [begin, end) of `rank` (the first n_rows % world ranks get
    one extra row); ranges of all ranks tile [0, n_rows) in rank order"""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} of {world}")
    base, extra = divmod(int(n_rows), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _column_tensor(col) -> torch.Tensor:
    nda = col.nda if hasattr(col, "nda") else col
    return nda if isinstance(nda, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(nda))


def gather_table(tb_local, n_rows: int, dst: int = 0, group=None):
    """The one collective of the path: gather the per-rank output tables (row ranges of
    :func:`shard_rows`, in rank order) on rank `dst`.  Returns the full table there, None elsewhere.
    Columns travel as they are (device tensors over NCCL, host arrays over gloo)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return tb_local
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    full = {}
    for name in tb_local.keys():
        col = tb_local[name]
        if tables.kind_of(col) not in ("array", "aoesa"):
            raise TypeError(f"gather of column {name}: only fixed-shape columns are gathered")
        t = _column_tensor(col)
        # unsigned 16/32/64-bit columns (e.g. the uint32 peak counters of get_multi_local_extrema) travel as the
        # signed type of the same width: the collectives' dtype support for them varies between backends
        wire = {torch.uint16: torch.int16, torch.uint32: torch.int32, torch.uint64: torch.int64}.get(t.dtype)
        dtype = t.dtype
        if wire is not None:
            t = t.view(wire)
        # equal-sized pieces (all_gather), padded to the longest shard
        piece = t.new_zeros((longest, *t.shape[1:]))
        piece[: t.shape[0]] = t
        pieces = [torch.empty_like(piece) for _ in range(world)]
        dist.all_gather(pieces, piece, group=group)
        if rank == dst:
            cat = torch.cat([p[: e - b] for p, (b, e) in zip(pieces, sizes)], dim=0)
            if wire is not None:
                cat = cat.view(dtype)
            full[name] = type(col)(cat if cat.is_cuda else cat.numpy(), attrs=dict(col.attrs))
    return tables.Table(full, size=n_rows) if rank == dst else None


def build_dsp_sharded(raw_table, dsp_config, database=None, outputs=None, block_width=None, device=None,
                      gather_to: int | None = 0, group=None):
    """``build_dsp`` on this rank's shard of `raw_table` (every rank holds, or can address, the
    whole raw table -- e.g. the same LH5 file); the output tables are gathered on rank
    `gather_to` (None: every rank keeps its shard)."""
    from .build_dsp import build_dsp

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(raw_table)
    begin, end = shard_rows(n, rank, world)
    # the reference's own row-range arguments (build_dsp.py:36-39) select the shard
    out = build_dsp(raw_table, dsp_config=dsp_config, database=database, outputs=outputs, i_start=begin,
                    n_entries=end - begin, block_width=block_width, device=device)
    if gather_to is None or world == 1:
        return out
    return gather_table(out, n, dst=gather_to, group=group)
