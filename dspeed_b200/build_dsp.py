"""``build_dsp`` -- the entry point of dspeed, same signature and behaviour
(reference: src/dspeed/build_dsp.py:27-452), driving the B200 processing chain.

In-memory inputs (a ``Table`` of columns: our :mod:`dspeed_b200.tables` classes, real
``lgdo`` objects, or an iterator of such tables) are processed directly.  LH5 *files* are
read and written through ``lh5`` (legend-lh5io) when that package is importable on the
box; it is an un-vendored dependency of the reference and is not part of this
repository, so without it a file path raises a clear error.

Differences that are deliberate: ``buffer_len`` (rows per chunk read from a file) and
``block_width`` (rows per device step) default to large values -- a B200 wants tens of
thousands of waveforms in flight -- and results do not depend on either.
"""

from __future__ import annotations

import logging
import os
import re
import time
from collections.abc import Collection, Mapping
from copy import deepcopy
from fnmatch import fnmatch

from . import tables
from .errors import DSPFatal, ProcessingChainError
from .processing_chain import build_processing_chain

log = logging.getLogger("dspeed")

try:  # optional: LH5 file I/O
    import lh5  # type: ignore
except ImportError:  # pragma: no cover - depends on the box
    lh5 = None


def _load_yaml_or_json(path):
    from yaml import safe_load

    with open(os.path.expandvars(os.path.expanduser(path))) as f:
        return safe_load(f)


def _is_table(obj) -> bool:
    return isinstance(obj, Mapping) and hasattr(obj, "__len__")


def _slice_table(tb, start, stop):
    """row range [start, stop) of an in-memory table (views, no copies)"""
    if start == 0 and stop >= len(tb):
        return tb
    kind = tables.kind_of(tb)
    if kind == "wftable":
        vals = tables.wf_values(tb)
        return tables.WaveformTable(
            size=stop - start, t0=tables.Array(tb.t0.nda[start:stop], attrs=tb.t0.attrs),
            dt=tables.Array(tb.dt.nda[start:stop], attrs=tb.dt.attrs),
            values=tables.ArrayOfEqualSizedArrays(vals.nda[start:stop], attrs=getattr(vals, "attrs", {})),
            attrs=getattr(tb, "attrs", {}))
    if kind == "table":
        return tables.Table({k: _slice_table(v, start, stop) for k, v in tb.items()}, size=stop - start,
                            attrs=getattr(tb, "attrs", {}))
    if kind in ("array", "aoesa"):
        cls = tables.ArrayOfEqualSizedArrays if kind == "aoesa" else tables.Array
        return cls(tb.nda[start:stop], attrs=getattr(tb, "attrs", {}))
    if kind in ("numpy", "tensor"):
        return tb[start:stop]
    raise ProcessingChainError(f"cannot slice a column of kind {kind}")


def build_dsp(
    raw_in,
    dsp_out: str | None = None,
    dsp_config=None,
    lh5_tables=None,
    base_group: str = None,
    database=None,
    outputs: Collection[str] = None,
    write_mode: str = None,
    entry_list=None,
    entry_mask=None,
    i_start: int = 0,
    n_entries: int | None = None,
    buffer_len: int = None,
    block_width: int = None,
    chan_config=None,
    device=None,
):
    """Convert raw-tier data into dsp-tier data by running the processing chain defined
    by ``dsp_config`` on the GPU.  Arguments as in the reference (build_dsp.py:27-126).
    Returns the output ``Table`` / ``Struct`` when ``dsp_out`` is ``None``."""
    db_parser = re.compile(r"(?![^\w_.])db\.[\w_.]+")

    if isinstance(lh5_tables, str):
        lh5_tables = [lh5_tables]

    in_memory = _is_table(raw_in) or (hasattr(raw_in, "__iter__") and not isinstance(raw_in, str))
    if in_memory:
        if base_group is None:
            base_group = ""
        if lh5_tables is None:
            lh5_tables = [""]
        elif len(lh5_tables) > 1:
            raise RuntimeError("Cannot have more than one value in lh5_tables for input of type Table or iterator")
    elif isinstance(raw_in, str):
        if lh5 is None:
            raise RuntimeError(
                "reading LH5 files needs the `lh5` (legend-lh5io) package, which is not installed here; "
                "pass an in-memory Table instead")
        if base_group is None:
            base_group = "raw" if lh5.ls(raw_in, "raw") else ""
        if lh5_tables is None:
            lh5_tables = lh5.ls(raw_in, f"{base_group}/*")
        else:
            lh5_tables = [tab for wc in lh5_tables for tab in lh5.ls(raw_in, f"{base_group}/{wc}")]
        found = []
        for tb in lh5_tables:
            if lh5.ls(raw_in, f"{tb}/*") == [f"{tb}/raw"]:
                found.append(f"{tb}/raw")
            elif lh5.ls(raw_in, tb):
                found.append(tb)
        lh5_tables = found
        if len(lh5_tables) == 0:
            raise RuntimeError(f"could not find any valid LH5 table in {raw_in}")
    else:
        raise RuntimeError(f"raw_in was not a file name, Table, or iterator: {raw_in}")

    if isinstance(dsp_config, str):
        dsp_config = _load_yaml_or_json(dsp_config)
    if isinstance(chan_config, str):
        chan_config = _load_yaml_or_json(chan_config)
    elif chan_config is None:
        chan_config = {}
    chan_config = dict(chan_config)
    for chan, config in chan_config.items():
        if isinstance(config, str):
            chan_config[chan] = _load_yaml_or_json(config)
    if isinstance(database, str):
        database = _load_yaml_or_json(database)
    if database and not isinstance(database, Mapping):
        raise ValueError("input database is not a valid JSON or YAML file or dict")

    if dsp_out is None:
        dsp_st = tables.Struct()
        store = None
    else:
        if lh5 is None:
            raise RuntimeError("writing LH5 files needs the `lh5` (legend-lh5io) package")
        if write_mode is None and os.path.isfile(dsp_out):
            raise FileExistsError(f"output file {dsp_out} exists. Set the 'write_mode' keyword")
        if write_mode == "r" and os.path.isfile(dsp_out):
            os.remove(dsp_out)
        dsp_st = None
        store = lh5.LH5Store(keep_open=True)

    for tb in lh5_tables:
        this_config = dsp_config
        for pat, config in chan_config.items():
            if fnmatch(tb, pat):
                this_config = config
                break
        if this_config is None:
            raise ProcessingChainError(f"no dsp_config for table '{tb}'")

        if tb not in ("", "raw"):
            chan_name = next(k for k in tb.split("/") if k not in ("", "raw"))
            db_dict = database.get(chan_name) if database else None
        else:
            db_dict = database

        if isinstance(raw_in, str):
            lh5_in = lh5.LH5Iterator(raw_in, tb, entry_list=entry_list, entry_mask=entry_mask, i_start=i_start,
                                     n_entries=n_entries, buffer_len=buffer_len or 65536)
        else:
            lh5_in = raw_in

        # auxiliary ("friend") inputs declared by the config
        config_inputs = this_config.get("inputs", [])
        if isinstance(config_inputs, Mapping):
            config_inputs = [config_inputs]
        for ci in config_inputs:
            file, group = ci["file"], ci["group"]
            prefix, suffix = ci.get("prefix", ""), ci.get("suffix", "")
            for what in ("file", "group"):
                val = file if what == "file" else group
                if isinstance(val, str) and db_parser.fullmatch(val):
                    try:
                        node = db_dict
                        for key in val.split(".")[1:]:
                            node = node[key]
                    except (KeyError, TypeError):
                        raise ProcessingChainError(f"did not find {val} in database.")
                    if what == "file":
                        file = node
                    else:
                        group = node
            if _is_table(file):  # an in-memory friend table
                for k, v in file.items():
                    lh5_in[prefix + k + suffix] = v
            elif lh5 is not None and isinstance(lh5_in, lh5.LH5Iterator):
                lh5_in.add_friend(
                    lh5.LH5Iterator(file, group, entry_list=entry_list, entry_mask=entry_mask, i_start=i_start,
                                    n_entries=n_entries, buffer_len=buffer_len or 65536),
                    prefix=prefix, suffix=suffix)
            elif lh5 is not None:
                lh5_in.join(lh5.LH5Store().read(group, file, n_rows=len(lh5_in)), prefix=prefix, suffix=suffix)
            else:
                raise RuntimeError("auxiliary file inputs need the `lh5` package")

        processors = this_config["processors"]
        _outputs = this_config["outputs"] if outputs is None else outputs

        is_iter = lh5 is not None and isinstance(lh5_in, lh5.LH5Iterator)
        if is_iter:
            tot_n_rows = len(lh5_in) if n_entries is None else min(n_entries, len(lh5_in))
            lh5_it = lh5_in
            lh5_it.n_entries = tot_n_rows
            tb_in = next(iter(lh5_in))
        elif _is_table(lh5_in):
            tot_n_rows = len(lh5_in) - i_start if n_entries is None else min(n_entries, len(lh5_in) - i_start)
            tb_in = _slice_table(lh5_in, i_start, i_start + tot_n_rows)
            lh5_it = [tb_in]
        else:  # a plain python iterator of tables
            lh5_it = iter(lh5_in)
            tb_in = next(lh5_it)
            lh5_it = _chain_first(tb_in, lh5_it)
            tot_n_rows = None

        log.info(f"Processing table {tb} with {tot_n_rows} rows")
        start = time.time()
        proc_chain, field_mask, tb_out = build_processing_chain(
            processors, tb_in, db_dict=db_dict, outputs=_outputs, block_width=block_width, device=device)
        if is_iter:
            lh5_it.reset_field_mask(field_mask)
        loading_time = time.time() - start
        processing_time = write_time = 0.0

        dsp_name = tb.replace("raw", "dsp")
        tb_fill = None
        if store is None:
            tb_fill = []
        i_entry = 0
        curr = time.time()
        for tb_in in lh5_it:
            loading_time += time.time() - curr
            t0 = time.time()
            if is_iter:
                i_entry = lh5_it.current_i_entry
            try:
                if len(tb_out) != len(tb_in):
                    tb_out.resize(len(tb_in))
                proc_chain(tb_in, tb_out)
            except DSPFatal as e:
                e.wf_range = f"{i_entry}-{i_entry + len(tb_in)}"
                raise e
            processing_time += time.time() - t0
            t0 = time.time()
            if store is not None:
                store.write(obj=_to_lgdo(tb_out), name=dsp_name, lh5_file=dsp_out,
                            wo_mode="o" if write_mode == "u" else "a", write_start=i_start + i_entry,
                            n_rows=len(tb_in))
            else:
                tb_fill.append(deepcopy(tb_out) if not _single_chunk(lh5_it) else tb_out)
            write_time += time.time() - t0
            if not is_iter:
                i_entry += len(tb_in)
            curr = time.time()

        log.info(f"Table {tb} processed in {time.time() - start:.2f} seconds")
        log.debug(f"Table {tb} loading time: {loading_time:.2f} s, processing {processing_time:.2f} s, "
                  f"write {write_time:.2f} s")
        if log.isEnabledFor(logging.DEBUG):
            for proc, t in sorted(proc_chain.get_timing().items(), key=lambda kv: kv[1], reverse=True):
                log.debug(f"{proc}: {t:.3f} s")

        if store is None:
            result = tb_fill[0] if len(tb_fill) == 1 else _concat_tables(tb_fill)
            result.proc_chain = proc_chain
            if dsp_name != "":
                groups = dsp_name.split("/")
                tb_name = groups.pop(-1)
                node = dsp_st
                for gr in groups:
                    node = node.setdefault(gr, tables.Struct())
                node[tb_name] = result
            else:
                dsp_st = result

    if store is None:
        return dsp_st
    return None


def _single_chunk(it) -> bool:
    return isinstance(it, list) and len(it) == 1


def _chain_first(first, rest):
    yield first
    yield from rest


def _concat_data(arrays):
    """row-wise concatenation of numpy arrays or torch tensors (a mixed list ends up on the host)"""
    import numpy as np
    import torch

    if all(isinstance(a, torch.Tensor) for a in arrays):
        return torch.cat(list(arrays))
    return np.concatenate([a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in arrays])


def _concat_columns(cols):
    """one output column from the same column of consecutive chunks, whatever its kind (the reference appends chunks
    generically through lh5 / lgdo; here: arrays, arrays of equal-sized arrays, waveform tables, vectors of vectors,
    nested tables; data on the host or on the device)"""
    import numpy as np

    first = cols[0]
    kind = tables.kind_of(first)
    if kind == "vov":
        ends, base = [], 0
        for c in cols:
            cl = c.cumulative_length.nda
            cl = cl.cpu().numpy() if hasattr(cl, "cpu") else np.asarray(cl)
            ends.append(cl.astype(np.int64) + base)
            base = int(ends[-1][-1]) if len(cl) else base
        flat = [c.flattened_data.nda[: int(c.cumulative_length.nda[-1]) if len(c) else 0] for c in cols]
        return tables.VectorOfVectors(flattened_data=tables.Array(_concat_data(flat), attrs=dict(first.flattened_data.attrs)),
                                      cumulative_length=tables.Array(np.concatenate(ends).astype(np.uint32)),
                                      attrs=dict(first.attrs))
    if kind == "wftable":
        t0, dt, values = (_concat_columns([getattr(c, f) if f != "values" else tables.wf_values(c) for c in cols])
                          for f in ("t0", "dt", "values"))
        return tables.WaveformTable(size=len(t0), t0=t0, dt=dt, values=values, attrs=dict(first.attrs))
    if kind == "table":
        out = tables.Table(size=sum(len(c) for c in cols), attrs=dict(first.attrs))
        for k in first:
            out.add_field(k, _concat_columns([c[k] for c in cols]))
        return out
    if kind in ("array", "aoesa"):
        return type(first)(_concat_data([c.nda for c in cols]), attrs=dict(first.attrs))
    raise TypeError(f"cannot concatenate output columns of kind {kind}")


def _concat_tables(parts):
    out = tables.Table(size=sum(len(p) for p in parts), attrs=parts[0].attrs)
    for k in parts[0]:
        out.add_field(k, _concat_columns([p[k] for p in parts]))
    return out


def _to_lgdo(tb):
    """convert our light-weight tables into real lgdo objects for LH5Store.write"""
    import lgdo  # type: ignore
    import numpy as np

    out = lgdo.Table(size=len(tb))
    for k, v in tb.items():
        if tables.kind_of(v) == "wftable":
            out.add_field(k, lgdo.WaveformTable(t0=np.asarray(v.t0.nda), t0_units=v.t0_units, dt=np.asarray(v.dt.nda),
                                                dt_units=v.dt_units, values=np.asarray(v.values.nda), attrs=v.attrs))
        elif tables.kind_of(v) == "aoesa":
            out.add_field(k, lgdo.ArrayOfEqualSizedArrays(nda=np.asarray(v.nda), attrs=v.attrs))
        else:
            out.add_field(k, lgdo.Array(np.asarray(v.nda), attrs=v.attrs))
    return out
