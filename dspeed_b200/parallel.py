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


def _as_t(nda) -> torch.Tensor:
    return nda if isinstance(nda, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(nda))


def _parts(name, col):
    """fixed-shape per-row pieces of a column: [(key, tensor [rows, ...])] and its ragged payload (or None)"""
    kind = tables.kind_of(col)
    if kind in ("array", "aoesa"):
        return kind, [("", _as_t(col.nda))], None
    if kind == "wftable":
        vals = tables.wf_values(col)
        if tables.kind_of(vals) != "aoesa":
            raise TypeError(f"gather of column {name}: waveform values must have a fixed length")
        return kind, [("values", _as_t(vals.nda)), ("t0", _as_t(col.t0.nda)), ("dt", _as_t(col.dt.nda))], None
    if kind == "vov":
        cl = _as_t(col.cumulative_length.nda).to(torch.int64)
        lens = torch.diff(cl, prepend=cl.new_zeros(1)).to(torch.int32)
        total = int(cl[-1]) if len(cl) else 0
        return kind, [("len", lens)], _as_t(col.flattened_data.nda)[:total]
    raise TypeError(f"gather of column {name}: unsupported column type {type(col).__name__}")


def _bytes2d(t: torch.Tensor, rows: int) -> torch.Tensor:
    """[rows, ...] tensor -> [rows, bytes per row] uint8 view (copy only when it is not contiguous)"""
    per_row = int(np.prod(t.shape[1:])) if t.ndim > 1 else 1
    if rows == 0:
        return torch.empty((0, per_row * t.element_size()), dtype=torch.uint8, device=t.device)
    return t.reshape(rows, per_row).contiguous().view(torch.uint8).reshape(rows, -1)


def gather_table(tb_local, n_rows: int, dst: int = 0, group=None):
    """The one collective of the path: gather the per-rank output tables (row ranges of :func:`shard_rows`, in rank
    order) on rank `dst`.  Returns the full table there, None elsewhere.

    All fixed-shape columns (scalars, equal-sized arrays, waveform tables, the per-row lengths of variable-length
    columns) are packed into ONE [rows, bytes per row] buffer per rank and travel in a single ``dist.gather`` to `dst`
    -- only `dst` receives anything.  Variable-length (VectorOfVectors) payloads need their sizes first: one more tiny
    exchange and one more gather, only when such columns exist.  Over NCCL the packed buffer is staged on the rank's
    device (host columns are copied there once); over gloo it stays on the host."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return tb_local
    nccl = dist.get_backend(group) == "nccl"
    wire_dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    rows = sizes[rank][1] - sizes[rank][0]
    longest = max(e - b for b, e in sizes)
    names = list(tb_local.keys())
    layout, pieces, ragged = [], [], []
    for name in names:
        kind, parts, payload = _parts(name, tb_local[name])
        for key, t in parts:
            if t.shape[0] != rows:
                raise ValueError(f"column {name}: {t.shape[0]} rows on rank {rank}, the shard has {rows}")
            b = _bytes2d(t.to(wire_dev), rows)
            layout.append((name, kind, key, t.dtype, tuple(t.shape[1:]), b.shape[1]))
            pieces.append(b)
        if payload is not None:
            ragged.append((name, payload))
    width = sum(l[5] for l in layout)
    packed = torch.zeros((longest, width), dtype=torch.uint8, device=wire_dev)
    if rows and pieces:
        packed[:rows] = torch.cat(pieces, dim=1)
    recv = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, recv, dst=dst, group=group)
    # ---- variable-length payloads: sizes, then one padded gather of all of them ---------------------------------
    rag = {}
    if ragged:
        tot = torch.tensor([p.numel() * p.element_size() for _, p in ragged], dtype=torch.int64, device=wire_dev)
        tots = [torch.empty_like(tot) for _ in range(world)]
        dist.all_gather(tots, tot, group=group)
        cap = max(int(t.sum()) for t in tots)
        buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=wire_dev)
        off = 0
        for _, p in ragged:
            nb = p.numel() * p.element_size()
            if nb:
                buf[off:off + nb] = p.to(wire_dev).contiguous().view(torch.uint8).reshape(-1)
            off += nb
        rrecv = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, rrecv, dst=dst, group=group)
        if rank == dst:
            for j, (name, p) in enumerate(ragged):
                chunks = []
                for r in range(world):
                    o = int(tots[r][:j].sum())
                    chunks.append(rrecv[r][o:o + int(tots[r][j])])
                rag[name] = torch.cat(chunks).contiguous().view(p.dtype)
    if rank != dst:
        return None
    body = torch.cat([recv[r][: e - b] for r, (b, e) in enumerate(sizes)], dim=0)
    got, off = {}, 0
    for name, kind, key, dtype, shape, nb in layout:
        t = body[:, off:off + nb].reshape(-1).clone().view(dtype).reshape(n_rows, *shape)    # (fresh, aligned storage)
        got.setdefault(name, {})[key] = t
        off += nb

    def host(t):
        return t if t.is_cuda else t.numpy()

    full = {}
    for name in names:
        col, parts = tb_local[name], got[name]
        kind = tables.kind_of(col)
        if kind in ("array", "aoesa"):
            full[name] = type(col)(host(parts[""]), attrs=dict(col.attrs))
        elif kind == "wftable":
            full[name] = tables.WaveformTable(size=n_rows, t0=tables.Array(host(parts["t0"]), attrs=dict(col.t0.attrs)),
                                              dt=tables.Array(host(parts["dt"]), attrs=dict(col.dt.attrs)),
                                              values=host(parts["values"]), attrs=dict(col.attrs))
            full[name].t0_units, full[name].dt_units = getattr(col, "t0_units", None), getattr(col, "dt_units", None)
        else:
            cl = torch.cumsum(parts["len"].to(torch.int64), 0).to(torch.uint32)
            full[name] = tables.VectorOfVectors(flattened_data=tables.Array(host(rag[name])),
                                                cumulative_length=tables.Array(host(cl)), attrs=dict(col.attrs))
    return tables.Table(full, size=n_rows)


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
    # the reference's own row-range arguments (build_dsp.py:36-39) select the shard; a rank without rows (fewer
    # events than ranks) still builds the chain on one row, so that it knows the columns it contributes nothing to
    empty = end == begin
    out = build_dsp(raw_table, dsp_config=dsp_config, database=database, outputs=outputs,
                    i_start=min(begin, max(n - 1, 0)) if empty else begin, n_entries=1 if empty else end - begin,
                    block_width=block_width, device=device)
    if empty and out is not None:
        out.resize(0)
    if gather_to is None or world == 1:
        return out
    return gather_table(out, n, dst=gather_to, group=group)
