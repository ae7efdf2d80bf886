"""Light-weight stand-ins for the ``lgdo`` containers the chain reads and writes.

The reference moves data in ``lgdo.Table`` / ``Array`` / ``ArrayOfEqualSizedArrays`` /
``VectorOfVectors`` / ``WaveformTable`` objects (processing_chain.py:1984-2360,
build_dsp.py:346-432).  ``lgdo`` is an un-vendored dependency that is not present in the
build image, so the chain is written against the small duck-typed protocol below
(``.nda``, ``.attrs``, ``.resize``; ``.values/.t0/.dt`` for waveform tables;
``.flattened_data/.cumulative_length`` for vectors of vectors).  Real ``lgdo`` objects
satisfy the same protocol and are accepted as they are (see :func:`kind_of`).

Storage (``nda``) is a numpy array (host; pinned when allocated through
:func:`pinned_empty`) or a torch tensor (device-resident columns are processed without
any host round trip).
"""

from __future__ import annotations

from collections.abc import Mapping

import numpy as np
import torch


def pinned_empty(shape, dtype) -> np.ndarray:
    """Host array backed by page-locked memory when CUDA is available (so H2D/D2H copies
    can be asynchronous and overlap with kernels); plain numpy otherwise."""
    dtype = np.dtype(dtype)
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__iter__") else (shape,)))
    if torch.cuda.is_available():
        try:
            t = torch.empty(shape, dtype=_np_to_torch(dtype), pin_memory=True)
            a = t.numpy()
            _PINNED_OWNERS[id(a)] = t
            return a
        except (RuntimeError, KeyError):
            pass
    return np.empty(shape, dtype)


_PINNED_OWNERS: dict[int, torch.Tensor] = {}

_NP2T = {
    "float32": torch.float32, "float64": torch.float64, "uint16": torch.uint16, "int16": torch.int16,
    "int32": torch.int32, "uint32": torch.uint32, "int64": torch.int64, "uint64": torch.uint64,
    "uint8": torch.uint8, "int8": torch.int8, "bool": torch.bool,
}


def _np_to_torch(dt) -> torch.dtype:
    return _NP2T[np.dtype(dt).name]


def np_dtype_of(x) -> np.dtype:
    """numpy dtype of a numpy array or torch tensor"""
    if isinstance(x, torch.Tensor):
        return np.dtype(str(x.dtype).replace("torch.", ""))
    return np.asarray(x).dtype


class LGDO:
    def __init__(self, attrs=None):
        self.attrs = dict(attrs or {})


class Scalar(LGDO):
    def __init__(self, value, attrs=None):
        super().__init__(attrs)
        self.value = value


class Array(LGDO):
    """1-D (or N-D) column: ``nda[row, ...]``"""

    def __init__(self, nda=None, shape=None, dtype=None, fill_val=None, attrs=None):
        super().__init__(attrs)
        if nda is None:
            shape = (shape,) if np.isscalar(shape) else tuple(shape)
            nda = pinned_empty(shape, dtype or np.float64)
            if fill_val is not None:
                nda[...] = fill_val
            else:
                nda[...] = 0
        self.nda = nda

    @property
    def dtype(self):
        return np_dtype_of(self.nda)

    @property
    def shape(self):
        return tuple(self.nda.shape)

    def __len__(self):
        return int(self.nda.shape[0])

    def __getitem__(self, key):
        return self.nda[key]

    def __setitem__(self, key, val):
        self.nda[key] = val

    def resize(self, new_size: int):
        new_size = int(new_size)
        if new_size == len(self):
            return
        if isinstance(self.nda, torch.Tensor):
            new = torch.zeros((new_size,) + tuple(self.nda.shape[1:]), dtype=self.nda.dtype, device=self.nda.device)
        else:
            new = pinned_empty((new_size,) + tuple(self.nda.shape[1:]), self.nda.dtype)
            new[...] = 0
        k = min(new_size, len(self))
        new[:k] = self.nda[:k]
        self.nda = new

    def form_datatype(self):
        return f"array<{self.nda.ndim}>{{{self.dtype}}}"


class ArrayOfEqualSizedArrays(Array):
    """2-D column: one fixed-length vector per row"""

    def form_datatype(self):
        return f"array_of_equalsized_arrays<1,{self.nda.ndim - 1}>{{{self.dtype}}}"


class VectorOfVectors(LGDO):
    """Ragged column: ``flattened_data`` + ``cumulative_length`` (end offsets)."""

    def __init__(self, flattened_data=None, cumulative_length=None, shape_guess=(0, 0), dtype=None, attrs=None):
        super().__init__(attrs)
        if flattened_data is None:
            n, m = shape_guess
            flattened_data = Array(shape=(max(int(n * m), 0),), dtype=dtype or np.float64)
            cumulative_length = Array(shape=(int(n),), dtype=np.uint32)
        if not isinstance(flattened_data, Array):
            flattened_data = Array(np.asarray(flattened_data))
        if not isinstance(cumulative_length, Array):
            cumulative_length = Array(np.asarray(cumulative_length, dtype=np.uint32))
        self.flattened_data = flattened_data
        self.cumulative_length = cumulative_length

    @property
    def dtype(self):
        return self.flattened_data.dtype

    def __len__(self):
        return len(self.cumulative_length)

    def resize(self, new_size: int):
        old = len(self)
        self.cumulative_length.resize(new_size)
        if new_size > old and old > 0:
            self.cumulative_length.nda[old:] = self.cumulative_length.nda[old - 1]

    def __getitem__(self, i):
        cl = self.cumulative_length.nda
        lo = int(cl[i - 1]) if i > 0 else 0
        return self.flattened_data.nda[lo : int(cl[i])]

    def _set_vector_unsafe(self, start: int, vectors, lens):
        """write rows ``start..`` from a padded 2-D array and per-row lengths"""
        cl = self.cumulative_length.nda
        off = int(cl[start - 1]) if start > 0 else 0
        lens = np.asarray(lens).astype(np.int64)
        ends = off + np.cumsum(lens)
        need = int(ends[-1]) if len(ends) else off
        if need > len(self.flattened_data):
            self.flattened_data.resize(need)
        vectors = np.asarray(vectors)
        mask = np.arange(vectors.shape[1])[None, :] < lens[:, None]
        self.flattened_data.nda[off:need] = vectors[mask]
        cl[start : start + len(lens)] = ends


class Table(LGDO, dict):
    """Column dictionary with a common length."""

    def __init__(self, col_dict=None, size=None, attrs=None):
        LGDO.__init__(self, attrs)
        dict.__init__(self)
        self.size = int(size) if size is not None else None
        if col_dict:
            for k, v in col_dict.items():
                self.add_field(k, v)
        if self.size is None:
            self.size = 0

    def add_field(self, name, obj):
        self[name] = obj
        if self.size is None:
            self.size = len(obj)

    def __len__(self):
        return int(self.size or 0)

    def resize(self, new_size):
        self.size = int(new_size)
        for v in self.values():
            if hasattr(v, "resize"):
                v.resize(new_size)

    def keys_list(self):
        return list(dict.keys(self))


class Struct(LGDO, dict):
    def __init__(self, obj_dict=None, attrs=None):
        LGDO.__init__(self, attrs)
        dict.__init__(self, obj_dict or {})


class WaveformTable(Table):
    """``t0``, ``dt`` (per-row scalars with ``units`` attrs) and ``values``."""

    def __init__(self, size=None, t0=0, t0_units=None, dt=1, dt_units=None, values=None, wf_len=None, dtype=None,
                 attrs=None):
        if values is not None and not isinstance(values, (Array, VectorOfVectors)) and not hasattr(values, "nda"):
            values = ArrayOfEqualSizedArrays(values)
        if size is None:
            size = len(values) if values is not None else 0
        if values is None:
            values = ArrayOfEqualSizedArrays(shape=(size, wf_len or 0), dtype=dtype or np.float64)
        if not isinstance(t0, Array) and not hasattr(t0, "nda"):
            t0v = t0
            t0 = Array(shape=(size,), dtype=np.float64)
            t0.nda[...] = t0v
        if not isinstance(dt, Array) and not hasattr(dt, "nda"):
            dtv = dt
            dt = Array(shape=(size,), dtype=np.float64)
            dt.nda[...] = dtv
        if t0_units is not None:
            t0.attrs["units"] = str(t0_units)
        if dt_units is not None:
            dt.attrs["units"] = str(dt_units)
        Table.__init__(self, {"t0": t0, "dt": dt, "values": values}, size=size, attrs=attrs)

    t0 = property(lambda self: self["t0"])
    dt = property(lambda self: self["dt"])
    values_ = property(lambda self: self["values"])

    @property
    def values(self):  # shadows dict.values on purpose, like lgdo.WaveformTable
        return self["values"]

    @property
    def t0_units(self):
        return self["t0"].attrs.get("units", None)

    @t0_units.setter
    def t0_units(self, u):
        self["t0"].attrs["units"] = str(u)

    @property
    def dt_units(self):
        return self["dt"].attrs.get("units", None)

    @dt_units.setter
    def dt_units(self, u):
        self["dt"].attrs["units"] = str(u)

    @property
    def wf_len(self):
        v = self["values"]
        return v.nda.shape[1] if hasattr(v, "nda") else None

    def resize(self, new_size):
        self.size = int(new_size)
        for k in ("t0", "dt", "values"):
            self[k].resize(new_size)


# ---- duck typing (our classes and real lgdo objects alike) ---------------------------
def kind_of(obj) -> str:
    """'numpy' | 'tensor' | 'wftable' | 'vov' | 'aoesa' | 'array' | 'table' | 'unknown'"""
    if isinstance(obj, np.ndarray):
        return "numpy"
    if isinstance(obj, torch.Tensor):
        return "tensor"
    if hasattr(obj, "flattened_data") and hasattr(obj, "cumulative_length"):
        return "vov"
    if all(hasattr(obj, a) for a in ("t0", "dt")) and _has_values(obj):
        return "wftable"
    if hasattr(obj, "nda"):
        return "aoesa" if getattr(obj.nda, "ndim", 1) > 1 else "array"
    if isinstance(obj, Mapping):
        return "table"
    return "unknown"


def _has_values(obj) -> bool:
    try:
        v = obj["values"] if isinstance(obj, Mapping) and "values" in obj else obj.values
        return hasattr(v, "nda") or hasattr(v, "flattened_data")
    except Exception:
        return False


def wf_values(obj):
    return obj["values"] if isinstance(obj, Mapping) and "values" in obj else obj.values
