"""Device-resident ``ProcessingChain``: the variable/buffer manager, the expression and
recipe compiler, and the block executor of dspeed, rebuilt for B200.

Interface and semantics follow the reference (src/dspeed/processing_chain.py; cited
per class): the same JSON/YAML recipe schema, ``get_variable`` expression grammar, type /
shape / unit / coordinate-grid deduction, unit-to-sample conversion and constant
folding.  What changes is *where things live and run*:

* every variable is a ``torch`` CUDA tensor ``[block_width, *shape]`` carrying the
  reference's metadata (dtype, shape, grid, unit, is_coord) -- reference
  ``ProcChainVar`` :147-377;
* processors are the CUDA processors of :mod:`dspeed_b200.processors` (numpy ufuncs used
  as glue map to device element-wise ops, :mod:`dspeed_b200.numpy_bridge`); a processor
  without a device implementation is a set-up error -- there is no CPU fallback;
* ``execute`` streams blocks: pinned host -> device copy of the input columns, all
  processor launches of the block on one CUDA stream, device -> pinned host copy of the
  output columns; data-dependent ``DSPFatal`` conditions are collected in a device-side
  record and raised once per block (reference :1144-1163);
* ``block_width`` keeps its meaning (rows per execute step) but defaults to a large
  value: results do not depend on it.

When :mod:`dspeed_b200.fusion` recognises the chain (or a prefix of it) it replaces the
per-processor launches by ONE waveform-resident kernel per block; see DESIGN.md.
"""

from __future__ import annotations

import ast
import importlib
import logging
import os
import time
from collections.abc import Collection, MutableMapping
from dataclasses import dataclass
from numbers import Real
from typing import Any

import numpy as np
import torch

from . import _lib, binding, grammar, numpy_bridge, recipe, tables
from . import processors as device_processors
from .processors import _i32, _i64, _vp
from .errors import DSPFatal, ProcessingChainError
from .tables import kind_of
from .units import Quantity, Unit, as_unit, from_foreign, is_in_registry, ureg

log = logging.getLogger("dspeed")

auto = "auto"

MAX_FATAL_SLOTS = 1024

#: rows per execute step when the caller does not choose (reference default is 16 rows of
#: numpy work; a B200 wants tens of thousands of waveforms in flight)
DEFAULT_BLOCK_WIDTH = 16384


class EndExecute(Exception):
    """raised by an input manager when the chunk is exhausted (reference :41-42)"""


def fatal_message(code: int) -> str:
    extra = {numpy_bridge.FATAL_CONVERT_INT: "Cannot convert to integer. Use round or astype",
             numpy_bridge.FATAL_GET_RANGE: "i is out of range"}
    return extra.get(code) or device_processors._lib.fatal_message(code)


def _np2t(dt) -> torch.dtype:
    return tables._np_to_torch(dt)


# ======================================================================================
# coordinate grids and variables
# ======================================================================================
@dataclass
class CoordinateGrid:
    """Period and offset of the sample grid a time coordinate is measured on
    (reference :67-144).  ``offset`` is a Quantity or a per-event variable (the
    waveform's ``t0``)."""

    period: Any
    offset: Any = 0

    def __post_init__(self) -> None:
        if isinstance(self.period, CoordinateGrid):
            self.offset = self.period.offset
            self.period = self.period.period
        elif isinstance(self.period, ProcChainVar):
            if self.period.grid in (None, auto):
                raise ProcessingChainError(f"{self.period} does not have an assigned coordinate grid")
            self.offset = self.period.offset
            self.period = self.period.period
        elif isinstance(self.period, Collection) and not isinstance(self.period, str):
            self.period, self.offset = self.period
        self.period = from_foreign(self.period)
        self.offset = from_foreign(self.offset)
        if isinstance(self.period, str):
            self.period = Quantity(1.0, self.period)
        elif isinstance(self.period, Unit):
            self.period = Quantity(1.0, self.period)
        if isinstance(self.offset, Real):
            self.offset = self.offset * self.period
        if not isinstance(self.period, Quantity) or not isinstance(self.offset, (Quantity, ProcChainVar)):
            raise ProcessingChainError(f"invalid coordinate grid ({self.period}, {self.offset})")

    def __eq__(self, other) -> bool:
        if not isinstance(other, CoordinateGrid):
            return False
        return self.period == other.period and (
            self.offset is other.offset if isinstance(self.offset, ProcChainVar) else self.offset == other.offset
        )

    def unit_str(self) -> str:
        return format(self.period.u, "~") or str(self.period.u)

    def get_period(self, unit) -> float:
        if isinstance(unit, str):
            unit = ureg.Quantity(unit)
        return float(self.period / unit)

    def get_offset(self, unit=None):
        """offset in ``unit`` (default: in periods): a float, or the device buffer of the
        offset variable converted to that unit"""
        if unit is None:
            unit = self.period
        elif isinstance(unit, str):
            unit = ureg.Quantity(unit)
        if isinstance(self.offset, ProcChainVar):
            return self.offset.get_buffer(CoordinateGrid(unit))
        return float(self.offset / unit)

    def __str__(self) -> str:
        off = self.offset.name if isinstance(self.offset, ProcChainVar) else str(self.offset)
        return f"({self.period},{off})"


class ProcChainVar:
    """A named chain variable and its device block buffer(s) (reference :147-377).

    Coordinate variables (times) keep one buffer per unit system they are requested in;
    the extra buffers are filled by conversion processors appended to the chain."""

    def __init__(self, proc_chain, name, shape=auto, dtype=auto, grid=auto, unit=auto, is_coord=auto,
                 vector_len=None, is_const=False):
        assert isinstance(proc_chain, ProcessingChain) and isinstance(name, str)
        self.proc_chain = proc_chain
        self.name = name
        self._native, self._native_system, self._twins = None, None, []
        self.shape = shape
        self.dtype = dtype
        self.grid = grid
        self.unit = unit
        self.is_coord = is_coord
        self.vector_len = vector_len
        self.is_const = is_const
        log.debug(f"added variable: {self.description()}")

    def __setattr__(self, name: str, value: Any) -> None:
        if value is auto:
            pass
        elif name == "shape":
            value = tuple(int(d) for d in value) if hasattr(value, "__iter__") else (int(value),)
        elif name == "dtype" and not isinstance(value, np.dtype):
            value = np.dtype(value)
        elif name == "grid" and not isinstance(value, CoordinateGrid) and value is not None:
            if isinstance(value, str):
                value = CoordinateGrid(value, 0)
            elif isinstance(value, Collection):
                value = CoordinateGrid(*value)
            else:
                value = CoordinateGrid(value, 0)
        elif name == "unit" and value is not None:
            value = from_foreign(value)
        elif name == "is_coord":
            value = bool(value)
        elif name == "vector_len" and value is not None:
            if not isinstance(value, ProcChainVar):
                value = self.proc_chain.get_variable(value)
            value.update_auto(shape=(), grid=None, unit=None, is_coord=False)
        super().__setattr__(name, value)

    # -- device storage -----------------------------------------------------------------------------------------------
    # A variable owns ONE block buffer in its native unit system (`_native`; coordinates: samples of the producing
    # waveform's grid).  When somebody asks for it in another unit system (an output column in ns, another waveform's
    # grid after slicing) a converted twin is created once and a conversion step is appended to the chain behind the
    # producer; `_twins` remembers them per unit system.  `_native_system` is set by the first request that names one.
    def _allocate(self) -> torch.Tensor:
        if self._native is None:
            for what in ("shape", "dtype"):
                if getattr(self, what) is auto:
                    raise ProcessingChainError(f"cannot deduce {what} of {self.name}")
            rows = 1 if self.is_const else self.proc_chain._block_width
            self._native = torch.zeros((rows,) + self.shape, dtype=_np2t(self.dtype), device=self.proc_chain.device)
        return self._native

    def _unit_system(self, unit):
        """the unit system a request means: the variable's own by default; a registry unit stands for a grid of that
        period without offset; anything else (None, 'ADC', ...) is no system at all"""
        if unit is None:
            unit = self.grid if self.is_coord else self.unit
        if not isinstance(unit, CoordinateGrid) and is_in_registry(unit):
            unit = CoordinateGrid(unit)
        return unit

    def get_buffer(self, unit=None) -> torch.Tensor:
        native = self._allocate()
        system = self._unit_system(unit)
        if self._native_system is None and not self._twins:
            if self.is_coord is True and not isinstance(self.grid, CoordinateGrid) and system is not None:
                self.grid = CoordinateGrid(system)     # a coordinate without a grid adopts the first system asked for
            if not isinstance(system, CoordinateGrid):
                return native
            self._native_system = system                # the first named system is the one the native buffer is in
        if not isinstance(system, CoordinateGrid) or system == self._native_system:
            return native
        for twin, twin_system in self._twins:
            if twin_system == system:
                return twin
        step = UnitConversionManager(self, system)
        self._twins.append((step.out_buffer, system))
        self.proc_chain._proc_managers.append(step)
        log.debug(f"added conversion: {step}")
        return step.out_buffer

    def all_buffers(self) -> list:
        """[(tensor, unit system or None)] of everything this variable has allocated (chain compilers map storages back
        to variables with it)"""
        if self._native is None:
            return []
        return [(self._native, self._native_system)] + list(self._twins)

    # legacy view of the storage: a tensor, or [(tensor, system), ...] once unit systems are involved
    @property
    def _buffer(self):
        if self._native is None:
            return None
        if self._native_system is None and not self._twins:
            return self._native
        return self.all_buffers()

    @_buffer.setter
    def _buffer(self, value):
        if value is None:
            self._native, self._native_system, self._twins = None, None, []
        elif isinstance(value, list):
            (self._native, self._native_system), self._twins = value[0], list(value[1:])
        else:
            self._native, self._native_system, self._twins = value, None, []

    @property
    def buffer(self):
        return self.get_buffer()

    @property
    def period(self):
        return self.grid.period if self.grid else None

    @property
    def offset(self):
        return self.grid.offset if self.grid else None

    def description(self) -> str:
        return (f"{self.name}(shape: {self.shape}, dtype: {self.dtype}, grid: {self.grid}, "
                f"unit: {self.unit}, is_coord: {self.is_coord})")

    def update_auto(self, shape=auto, dtype=auto, grid=auto, unit=auto, is_coord=auto, period=None, offset=0,
                    vector_len=None) -> None:
        """fill in attributes that are still ``auto`` (reference :334-374)"""
        if grid is auto and period is not None:
            if isinstance(offset, str):
                offset = self.proc_chain.get_variable(offset, expr_only=True)
            grid = CoordinateGrid(period, offset)
        if self.shape is auto and shape is not auto:
            self.shape = shape
        if self.dtype is auto and dtype is not auto:
            self.dtype = dtype
        if self.grid is auto and grid is not auto:
            self.grid = grid
        if self.unit is auto and unit is not auto:
            self.unit = unit
        if self.is_coord is auto and is_coord is not auto:
            self.is_coord = is_coord
        if self.vector_len is None and vector_len is not None:
            self.vector_len = vector_len

    def __str__(self) -> str:
        return self.name


# ======================================================================================
# the chain
# ======================================================================================
class ProcessingChain:
    """Variables, processors and I/O links of one DSP chain, executed block by block on
    one CUDA device (reference :380-1482)."""

    def __init__(self, block_width: int = None, buffer_len: int = None, device=None) -> None:
        self._vars_dict: dict[str, ProcChainVar] = {}
        self._proc_managers: list = []
        self._input_managers: dict = {}
        self._output_managers: dict = {}
        self._block_width = int(block_width) if block_width else DEFAULT_BLOCK_WIDTH
        self._buffer_len = buffer_len
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("dspeed_b200.ProcessingChain needs a CUDA device (there is no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type not in ("cuda", "meta"):
            # "meta" = plan-only build (no data, nothing executes): used to inspect / test the
            # compiled chain without a GPU.  There is deliberately no "cpu" execution device.
            raise RuntimeError(f"unsupported device {self.device}: dspeed_b200 executes on CUDA only")
        #: device-side DSPFatal records, one int32[4] row per bound processor
        self.fatal = torch.zeros((MAX_FATAL_SLOTS, 4), dtype=torch.int32, device=self.device)
        self._fatal_owner: list = []
        self._fused = None  # set by fusion.try_fuse()
        #: launches = hand-written CUDA kernels launched; glue_ops = per-event scalar torch ops
        self.stats = {"launches": 0, "glue_ops": 0, "blocks": 0, "h2d_bytes": 0, "d2h_bytes": 0}
        self._event_timing = False

    # -- variables -----------------------------------------------------------------
    def add_variable(self, name, dtype=auto, shape=auto, grid=auto, unit=auto, is_coord=auto, period=None, offset=0,
                     vector_len=None) -> ProcChainVar:
        self._validate_name(name, raise_exception=True)
        if name in self._vars_dict:
            raise ProcessingChainError(name + " is already in variable list")
        if grid is auto and period is not None:
            if isinstance(offset, str):
                offset = self.get_variable(offset, expr_only=True)
            grid = CoordinateGrid(period, offset)
        var = ProcChainVar(self, name, shape=shape, dtype=dtype, grid=grid, unit=unit, is_coord=is_coord,
                           vector_len=vector_len)
        self._vars_dict[name] = var
        return var

    def set_constant(self, varname, val, dtype=None, unit=None) -> ProcChainVar:
        """make a variable a constant with the given value (reference :487-524)"""
        param = self.get_variable(varname)
        if not param.is_const and param._buffer is not None:
            raise ProcessingChainError(f"{param} is already defined, cannot set_constant")
        param.is_const = True
        val = from_foreign(val)
        if isinstance(val, Quantity):
            unit = val.u
            val = val.m
        if isinstance(val, torch.Tensor):
            val = val.detach().cpu().numpy()
        val = np.array(val, dtype=dtype)
        param.update_auto(shape=val.shape, dtype=val.dtype, unit=unit, is_coord=False)
        param.host_value = np.ascontiguousarray(val).astype(param.dtype, casting="unsafe")  # for the chain compilers
        buf = param.get_buffer()
        buf.copy_(torch.from_numpy(np.ascontiguousarray(val).astype(param.dtype, casting="unsafe")).reshape(buf.shape))
        log.debug(f"set constant: {param.description()} = {val}")
        return param

    # -- I/O links -----------------------------------------------------------------
    def link_io_buffer(self, varname, buff=None, output=False):
        """link a variable to a host/device column (reference :526-621)"""
        self._validate_name(varname, raise_exception=True)
        var = self.get_variable(varname, expr_only=True)
        if var is None:
            var = self.add_variable(varname)
        io_managers = self._output_managers if output else self._input_managers
        if not isinstance(var, ProcChainVar):
            raise ProcessingChainError("Must link an input buffer to a processing chain variable")

        if buff is None:
            dtype = var.dtype
            n = self._buffer_len
            if isinstance(var.grid, CoordinateGrid) and not var.is_coord:
                if var.vector_len is None:
                    buff = tables.WaveformTable(size=n, wf_len=var.shape[0], dtype=dtype)
                else:
                    buff = tables.WaveformTable(size=n, values=tables.VectorOfVectors(shape_guess=(n, 0), dtype=dtype))
            elif len(var.shape) == 0:
                buff = tables.Array(shape=(n,), dtype=dtype)
            elif var.vector_len is not None:
                buff = tables.VectorOfVectors(shape_guess=(n, 0), dtype=dtype)
            else:
                buff = tables.ArrayOfEqualSizedArrays(shape=(n, *var.shape), dtype=dtype)

        if varname in io_managers:
            io_managers[varname].set_buffer(buff)
            return buff

        kind = kind_of(buff)
        if kind in ("numpy", "tensor"):
            man = NumpyIOManager(buff, var)
        elif kind == "wftable":
            man = WaveformIOManager(buff, var)
        elif kind == "vov":
            man = VectorOfVectorsIOManager(buff, var)
        elif kind in ("array", "aoesa"):
            man = ArrayIOManager(buff, var)
        else:
            raise ProcessingChainError("Could not link input buffer of unknown type", str(buff))
        log.debug(f"added {'output' if output else 'input'} buffer: {man}")
        io_managers[varname] = man
        return buff

    def link_input_buffer(self, varname, buff=None):
        return self.link_io_buffer(varname, buff, output=False)

    def link_output_buffer(self, varname, buff=None):
        return self.link_io_buffer(varname, buff, output=True)

    # -- processors ----------------------------------------------------------------
    def add_processor(self, func, *args, signature=None, types=None, coord_grid=None) -> None:
        """bind a processor to variables / constants (reference :635-663)"""
        params = []
        kw_params = {}
        for param in args:
            if isinstance(param, str):
                param = self.get_variable(param)
            if isinstance(param, MutableMapping):
                kw_params.update(param)
            else:
                params.append(param)
        if coord_grid is not None:
            coord_grid = CoordinateGrid(coord_grid)
        proc_man = ProcessorManager(self, func, params, kw_params, signature, types, coord_grid)
        self._proc_managers.append(proc_man)
        log.debug(f"added processor: {proc_man}")

    # -- execution -----------------------------------------------------------------
    def execute(self, start: int = 0, stop: int = None) -> None:
        """run the chain over rows [start, stop) of the linked buffers, block by block"""
        if stop is None:
            stop = self._buffer_len
        if self._fused is not None and self._fused.can_run(self):
            self._fused.execute(self, start, stop)
            return
        with torch.cuda.device(self.device):
            for i in range(start, stop, self._block_width):
                try:
                    self._execute_procs(i, min(i + self._block_width, stop))
                except EndExecute:
                    break
            torch.cuda.current_stream(self.device).synchronize()

    def __call__(self, tb_in, out=None):
        """process a whole table (reference :675-716)"""
        self._buffer_len = len(tb_in)
        for varname in self._input_managers:
            if varname not in tb_in:
                raise ProcessingChainError(f"Require column {varname} in tb_in")
            self.link_input_buffer(varname, tb_in[varname])
        if out is None:
            out = tables.Table({v: self.link_output_buffer(v) for v in self._output_managers}, size=len(tb_in))
        else:
            for varname in self._output_managers:
                if varname not in out:
                    raise ProcessingChainError(f"Require column {varname} in out")
                self.link_output_buffer(varname, out[varname])
        self.execute()
        return out

    def _execute_procs(self, begin: int, end: int) -> None:
        """one block: copy in, launch every processor, copy out, raise recorded DSPFatal"""
        for in_man in self._input_managers.values():
            in_man.read(begin, end)
        current = None
        try:
            for proc_man in self._proc_managers:
                current = proc_man
                proc_man.execute()
        except DSPFatal as e:
            e.processor = str(current)
            e.wf_range = (begin, end)
            raise e
        self.stats["blocks"] += 1
        self._raise_recorded_fatal(begin, end)
        for out_man in self._output_managers.values():
            out_man.write(begin, end)

    def _new_fatal_slot(self, owner) -> torch.Tensor:
        k = len(self._fatal_owner)
        if k >= MAX_FATAL_SLOTS:
            k = MAX_FATAL_SLOTS - 1
        else:
            self._fatal_owner.append(owner)
        return self.fatal[k]

    def _raise_recorded_fatal(self, begin, end, block_width=None) -> None:
        n = max(1, len(self._fatal_owner))
        rec = self.fatal[:n].cpu()  # one small synchronising read
        hit = torch.nonzero(rec[:, 0])
        if hit.numel():
            k = int(hit[0])
            code = int(rec[k, 0])
            row = (int(rec[k, 2]) << 31) | int(rec[k, 1])
            self.fatal.zero_()
            e = DSPFatal(fatal_message(code))
            e.code = code
            if block_width:   # whole-call check: report the block that holds the recorded row
                b0 = begin + (max(row - begin, 0) // block_width) * block_width
                begin, end = b0, min(b0 + block_width, end)
            e.wf_range = (begin, end)
            if k < len(self._fatal_owner):
                e.processor = str(self._fatal_owner[k])
            raise e

    def get_timing(self) -> dict[str, float]:
        """cumulative time per processor in seconds (reference :1188-1190): device time
        measured with CUDA events after :meth:`enable_event_timing`, else host launch time"""
        if self._event_timing:
            self._resolve_events()
            return {str(proc): getattr(proc, "device_time", 0.0) for proc in self._proc_managers}
        return {str(proc): proc.time_total for proc in self._proc_managers}

    def enable_event_timing(self, on: bool = True) -> None:
        """bracket every processor launch with CUDA events on the launch stream"""
        self._event_timing = on
        self._pending_events = []
        if on:
            for proc in self._proc_managers:
                proc.device_time = 0.0
                proc.device_calls = 0

    def _resolve_events(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
        for proc, e0, e1 in self._pending_events:
            proc.device_time = getattr(proc, "device_time", 0.0) + e0.elapsed_time(e1) * 1e-3
            proc.device_calls = getattr(proc, "device_calls", 0) + 1
        self._pending_events = []

    def __str__(self) -> str:
        return ("Input variables:\n  " + "\n  ".join(str(m) for m in self._input_managers.values())
                + "\nProcessors:\n  " + "\n  ".join(str(m) for m in self._proc_managers)
                + "\nOutput variables:\n  " + "\n  ".join(str(m) for m in self._output_managers.values()))

    # -- expression grammar (reference :718-1130) ------------------------------------
    def get_variable(self, expr: str, get_names_only: bool = False, expr_only: bool = False) -> Any:
        """Evaluate ``expr`` in the chain's expression language (:mod:`dspeed_b200.grammar`): a variable name (created
        on first use), ``name(shape, dtype, ...)`` declarations, arithmetic / comparison / ternary expressions (each
        emits a processor), ``wf[a:b:c]`` views, ``round/floor/ceil/trunc/astype/where/isnan/isfinite/len`` calls, unit
        names, ``np.pi``-style constants and ``kw=expr``.  ``get_names_only``: only list the variable names the text
        mentions (nothing is created)."""
        try:
            stmt = ast.parse(expr).body[0]
            if get_names_only:
                return grammar.names_read(stmt.value, expr, type(self))
            var = grammar.Evaluator(self, expr)(stmt.value)
        except ProcessingChainError:
            raise
        except Exception as e:
            raise ProcessingChainError("Could not parse expression:\n  " + expr) from e
        if isinstance(stmt, ast.Expr):
            return var
        if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1:
            if expr_only:
                raise ProcessingChainError("kwarg assignment is not allowed in this context\n  " + expr)
            return {stmt.targets[0].id: var}
        raise ProcessingChainError("Could not parse expression:\n  " + expr)

    def _emit(self, func, params) -> None:
        """append one glue processor (an element-wise ufunc of the expression language) to the chain"""
        proc_man = ProcessorManager(self, func, params)
        self._proc_managers.append(proc_man)
        log.debug(f"added processor: {proc_man}")

    def _validate_name(self, name: str, raise_exception: bool = False) -> bool:
        ok = grammar.is_variable_name(name, type(self))
        if raise_exception and not ok:
            raise ProcessingChainError(f"{name} is not a valid variable name")
        return ok

    # -- helper functions callable inside expressions: dspeed_b200.grammar.HELPERS -------
    def _add_conversion(self, var, grid, out, mode=None, out_dtype=None) -> None:
        """append the unit conversion that expresses coordinate `var` on `grid` and writes into `out`"""
        conversion = UnitConversionManager(var, grid, mode=mode, out_dtype=out_dtype)
        out._buffer = conversion.out_buffer
        self._proc_managers.append(conversion)
        log.debug(f"added conversion: {conversion}")

    def _where(self, condition, a, b, dtype=auto):
        return grammar.h_where(self, condition, a, b, dtype)

    func_list = grammar.HELPERS
    module_list = {"np": np, "numpy": np}
    #: the classes the expression evaluator instantiates
    Variable = ProcChainVar
    Grid = CoordinateGrid


# ======================================================================================
# processor binding
# ======================================================================================
class ProcessorManager:
    """Freezes one processor call: picks the type loop, solves the gufunc dimensions,
    deduces ``auto`` variables, converts unit-carrying scalars to samples and binds the
    device buffers -- the launch descriptor of a block step (reference :1485-1803)."""

    def __init__(self, proc_chain, func, params, kw_params=None, signature=None, types=None, grid=None) -> None:
        assert isinstance(proc_chain, ProcessingChain) and callable(func) and isinstance(params, Collection)
        kw_params = kw_params or {}
        self.proc_chain = proc_chain
        self.params = [from_foreign(p) for p in params]
        self.kw_params = {k: from_foreign(v) for k, v in kw_params.items()}
        self.args: list = []
        self.kwargs: dict = {}
        self.time_total = 0.0
        self.n_calls = 0

        # the device implementation behind this callable (raises if there is none)
        self.host_func = func
        self.processor = numpy_bridge.device_equivalent(func, signature)

        # ---- the call's launch descriptor (binding.py): type loop, shape solution, lowered operands ------------------
        self.signature = signature if signature is not None else getattr(self.processor, "signature", None)
        if self.signature is None:
            self.signature = binding.scalar_layout(self.processor.nin, self.processor.nout)
        layouts = binding.core_dims(self.signature)
        operands = self.params + list(self.kw_params.values())
        if len(layouts) != len(operands):
            raise ProcessingChainError(
                f"expected {len(layouts)} arguments from signature {self.signature}; found "
                f"{len(operands)}: ({', '.join(str(p) for p in operands)})")
        arrays = [p if isinstance(p, (ProcChainVar, np.ndarray)) else None for p in operands]
        if types is None:
            types = list(getattr(self.processor, "types", None) or [])
        solution, call_grid = binding.solve_shapes(layouts, arrays, proc_chain._block_width, func.__name__)
        self.types = binding.pick_type_loop(types, arrays, self if types else func.__name__)
        # sampling grid of the call: given, else the first gridded operand's, else that of a coordinate operand
        self.grid = grid or call_grid or next((p.grid for p in operands if isinstance(p, ProcChainVar) and p.is_coord is True), None)
        keys = [None] * len(self.params) + list(self.kw_params.keys())
        for key, opnd, names, dtype in zip(keys, operands, layouts, self.types):
            missing = [n for n in names if n not in solution.core]
            if missing:
                raise ProcessingChainError(f"could not deduce dimension {missing[0]} for {opnd}")
            axes = solution.axes_of(names)
            bound = self._lower(opnd, dtype, tuple(ax.extent for ax in axes), axes[-1].grid if axes else None)
            if key is None:
                self.args.append(bound)
            else:
                self.kwargs[key] = bound

        self._sync_timing = bool(int(os.environ.get("DSPEED_B200_TIMING", "0")))
        self.fatal = proc_chain._new_fatal_slot(self)

    def _lower(self, opnd, dtype: np.dtype, shape: tuple, axis_grid):
        """one operand -> what the kernels take: a broadcastable view of the variable's block buffer (deducing the
        variable's open properties from the solution first), a scalar in samples, byte codes, or a device constant"""
        if isinstance(opnd, ProcChainVar):
            unit, is_coord, grid_override = binding.coordinate_role(opnd, self.grid)
            opnd.update_auto(shape=shape, dtype=np.dtype(dtype), grid=grid_override if grid_override is not None else axis_grid,
                             unit=unit, is_coord=is_coord)
            return binding.block_view(opnd.get_buffer(self.grid if opnd.is_coord else None), shape)
        if isinstance(opnd, str):
            return binding.text_operand(opnd, dtype, shape)
        if isinstance(opnd, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(opnd.astype(dtype))).to(self.proc_chain.device)
        if opnd is None:
            return None
        return binding.scalar_in_samples(opnd, dtype, self.grid, CoordinateGrid)

    def execute(self) -> None:
        start = time.perf_counter()
        if self.proc_chain._event_timing:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self.processor(*self.args, fatal=self.fatal, **self.kwargs)
            e1.record()
            self.proc_chain._pending_events.append((self, e0, e1))
        else:
            self.processor(*self.args, fatal=self.fatal, **self.kwargs)
        if self._sync_timing:
            torch.cuda.current_stream(self.proc_chain.device).synchronize()
        self.time_total += time.perf_counter() - start
        self.n_calls += 1
        if getattr(self.processor, "native_kernel", False):
            self.proc_chain.stats["launches"] += getattr(self.processor, "launches_per_call", 1)
        else:
            self.proc_chain.stats["glue_ops"] += 1

    def __str__(self) -> str:
        return (self.host_func.__name__ + "("
                + ", ".join([str(p) for p in self.params] + [f"{k}={v}" for k, v in self.kw_params.items()]) + ")")


class UnitConversionManager(ProcessorManager):
    """Converts a coordinate variable between unit systems / grids:
    ``(buf + offset_in) * ratio - offset_out`` in float64, optionally rounded
    (reference :1806-1908 and processors/unit_conversion.py:16-78)."""

    def __init__(self, var: ProcChainVar, unit, mode=None, out_dtype=None) -> None:
        if mode not in (None, "round", "floor", "ceil", "trunc"):
            raise ProcessingChainError("Mode must be round, floor, ceil or trunc")
        self.proc_chain = var.proc_chain
        self.mode = mode
        self.host_func = self.processor = numpy_bridge.convert
        stored = var.all_buffers()[0] if var._native_system is not None or var._twins else (var._buffer, None)
        source, source_system = stored
        if source_system is None:                    # a plain value in its own unit (no grid involved so far)
            source_system = var.unit
            if isinstance(source_system, str) and source_system in ureg:
                source_system = ureg.Quantity(source_system)
        self.params = [var]
        self.kw_params = {"from": source_system, "to": unit.period if isinstance(unit, CoordinateGrid) else unit}
        ratio, shift_in, shift_out = self._terms(source_system, unit)

        def per_row(x):     # per-event offsets broadcast along the variable's own axes
            return x.reshape(x.shape[0], *[1] * (source.ndim - x.ndim)) if isinstance(x, torch.Tensor) else x

        self.in_is_int = not np.issubdtype(var.dtype, np.floating)
        self.out_buffer = torch.zeros_like(source, dtype=_np2t(np.dtype(out_dtype) if out_dtype is not None else var.dtype))
        self.args = [source, per_row(shift_in), per_row(shift_out), ratio, self.out_buffer]
        self.kwargs = {}
        self.time_total, self.n_calls, self._sync_timing = 0.0, 0, False
        self.fatal = self.proc_chain._new_fatal_slot(self)

    @staticmethod
    def _terms(source, target):
        """(ratio, offset_in, offset_out) of ``(x + offset_in) * ratio - offset_out``: `x` counted in `source` (a grid:
        samples from its offset; a unit / quantity: multiples of it) expressed in `target`"""
        shift_out = 0
        if isinstance(target, CoordinateGrid):
            shift_out, target = target.get_offset(), target.period
        if isinstance(source, CoordinateGrid):
            return source.get_period(target), source.get_offset(), shift_out
        if isinstance(source, (Unit, Quantity)):
            one = lambda u: u if isinstance(u, Quantity) else Quantity(1.0, u)   # noqa: E731
            return float(one(source) / one(ureg.Quantity(target) if isinstance(target, str) else target)), 0, shift_out
        # a unitless count: 1 / (size of the target unit)
        return (1.0 / target.m if isinstance(target, Quantity) else 1 / float(target)), 0, shift_out

    def execute(self) -> None:
        start = time.perf_counter()
        numpy_bridge.convert(*self.args, mode=self.mode, int_check=self.in_is_int and self.mode is None,
                             fatal=self.fatal)
        self.time_total += time.perf_counter() - start
        self.n_calls += 1
        self.proc_chain.stats["glue_ops"] += 1

    def __str__(self) -> str:
        return f"convert({self.params[0]}, from={self.kw_params['from']}, to={self.kw_params['to']})"


# ======================================================================================
# I/O managers: chunk column <-> device block buffer
# ======================================================================================
def _as_tensor(x) -> torch.Tensor:
    """zero-copy torch view of a host numpy array / pass-through for tensors"""
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(x)


class IOManager:
    def _count(self, nbytes, out):
        self.var.proc_chain.stats["d2h_bytes" if out else "h2d_bytes"] += int(nbytes)


class NumpyIOManager(IOManager):
    """plain numpy array or torch tensor column (reference :1942-1981)"""

    def __init__(self, io_buf, var: ProcChainVar) -> None:
        var.update_auto(dtype=tables.np_dtype_of(io_buf), shape=tuple(io_buf.shape[1:]))
        self.var = var
        self.raw_var = var.buffer
        self.set_buffer(io_buf)

    def set_buffer(self, io_buf) -> None:
        if not isinstance(io_buf, (np.ndarray, torch.Tensor)):
            raise ProcessingChainError(f"{self.var} must be set using a numpy array or a torch tensor")
        if self.var.shape != tuple(io_buf.shape[1:]) or self.var.dtype != tables.np_dtype_of(io_buf):
            raise ProcessingChainError(f"array<{tuple(io_buf.shape)}>{{{tables.np_dtype_of(io_buf)}}} "
                                       f"is not compatible with variable {self.var}")
        self.io_buf = io_buf
        self.io_t = _as_tensor(io_buf)

    def read(self, start: int, end: int) -> None:
        if start >= self.io_t.shape[0]:
            raise EndExecute
        end = min(end, self.io_t.shape[0])
        src = self.io_t[start:end]
        self.raw_var[0 : end - start].copy_(src, non_blocking=True)
        if not src.is_cuda:
            self._count(src.numel() * src.element_size(), False)

    def write(self, start: int, end: int) -> None:
        dst = self.io_t[start:end]
        dst.copy_(self.raw_var[0 : end - start] if not self.var.is_const else self.raw_var.expand(end - start, *self.var.shape),
                  non_blocking=True)
        if not dst.is_cuda:
            self._count(dst.numel() * dst.element_size(), True)

    def __str__(self) -> str:
        return f"{self.var} linked to array(shape={tuple(self.io_buf.shape)}, dtype={tables.np_dtype_of(self.io_buf)})"


def _resolve_io_unit(var: ProcChainVar, unit):
    """unit system in which a column exchanges a variable (reference :1990-2011)"""
    if isinstance(var.unit, (CoordinateGrid, Quantity, Unit)):
        if isinstance(var.unit, CoordinateGrid):
            var_u = var.unit.period.u
        elif isinstance(var.unit, Quantity):
            var_u = var.unit.u
        else:
            var_u = var.unit
        if unit is None:
            unit = var_u
        elif ureg.is_compatible_with(var_u, unit):
            unit = as_unit(unit)
        else:
            raise ProcessingChainError(f"array and variable {var} have incompatible units ({var_u} and {unit})")
    elif isinstance(var.unit, str) and unit is None:
        unit = var.unit
    return unit


class ArrayIOManager(IOManager):
    """``Array`` / ``ArrayOfEqualSizedArrays`` column (reference :1984-2124)"""

    def __init__(self, io_array, var: ProcChainVar) -> None:
        unit = io_array.attrs.get("units", None)
        var.update_auto(dtype=tables.np_dtype_of(io_array.nda), shape=tuple(io_array.nda.shape[1:]), unit=unit)
        unit = _resolve_io_unit(var, unit)
        self.var = var
        self.raw_var = var.get_buffer(unit)
        self.set_buffer(io_array)

    def set_buffer(self, io_array) -> None:
        if not hasattr(io_array, "nda"):
            raise ProcessingChainError(f"{self.var} must be set using an Array")
        if "units" not in io_array.attrs and self.var.unit is not None:
            u = self.var.unit
            io_array.attrs["units"] = str(u.u) if isinstance(u, Quantity) else str(u)
        if (self.var.shape != tuple(io_array.nda.shape[1:])
                or np.dtype(str(self.raw_var.dtype).replace("torch.", "")) != tables.np_dtype_of(io_array.nda)):
            raise ProcessingChainError(f"LGDO object {io_array.form_datatype()} is incompatible with {self.var}")
        self.io_array = io_array

    def read(self, start: int, end: int) -> None:
        n = len(self.io_array)
        if start >= n:
            raise EndExecute
        end = min(end, n)
        src = _as_tensor(self.io_array.nda)[start:end]
        self.raw_var[0 : end - start].copy_(src, non_blocking=True)
        if not src.is_cuda:
            self._count(src.numel() * src.element_size(), False)

    def write(self, start: int, end: int) -> None:
        if len(self.io_array) < end:
            self.io_array.resize(end)
        dst = _as_tensor(self.io_array.nda)[start:end]
        src = self.raw_var.expand(end - start, *self.raw_var.shape[1:]) if self.var.is_const \
            else self.raw_var[0 : end - start]
        dst.copy_(src, non_blocking=True)
        if not dst.is_cuda:
            self._count(dst.numel() * dst.element_size(), True)

    def __str__(self) -> str:
        return (f"{self.var} linked to Array(shape={tuple(self.io_array.nda.shape)}, "
                f"dtype={tables.np_dtype_of(self.io_array.nda)}, attrs={self.io_array.attrs})")


class VectorOfVectorsIOManager(IOManager):
    """ragged column <-> NaN/zero padded block + length variable (reference :2127-2260).
    The ragged <-> padded conversion is done on the host side of the copy."""

    def __init__(self, io_vov, var: ProcChainVar) -> None:
        if var.vector_len is None:
            var.vector_len = ProcChainVar(var.proc_chain, f"len({var.name})", shape=(), dtype="uint32", grid=None,
                                          unit=None)
        if not np.issubdtype(var.vector_len.dtype, np.integer):
            raise ProcessingChainError(f"{var.vector_len} must be an integer to act as a vector len")
        self.unit = io_vov.attrs.get("units", None)
        var.update_auto(dtype=io_vov.dtype, unit=self.unit)
        self.unit = _resolve_io_unit(var, self.unit)
        self.var = var
        self.raw_var = None
        self.len_var = var.vector_len.get_buffer()
        self.set_buffer(io_vov)
        if var.shape is not auto:
            # the buffer in the column's unit system now: a coordinate variable (e.g. peak positions in ns) gets its
            # conversion manager appended at build time, not after the first block has already run
            self.raw_var = self.var.get_buffer(self.unit)

    def set_buffer(self, io_vov) -> None:
        if "units" not in io_vov.attrs and self.var.unit is not None:
            u = self.var.unit
            io_vov.attrs["units"] = str(u.u) if isinstance(u, Quantity) else str(u)
        if self.var.dtype != io_vov.dtype:
            raise ProcessingChainError(f"VectorOfVectors of {io_vov.dtype} is incompatible with {self.var}")
        self.io_vov = io_vov

    def _ensure_raw(self, start, end):
        if self.raw_var is None:
            if self.var.shape is auto:
                cl = np.asarray(self.io_vov.cumulative_length.nda)
                lens = np.diff(np.concatenate([[0 if start == 0 else cl[start - 1]], cl[start:end]]))
                self.var.update_auto(shape=2 * int(lens.max() if len(lens) else 1))
                log.warning(f"No maximum length provided for VectorOfVectors {self.var}; using {self.var.shape}")
            self.raw_var = self.var.get_buffer(self.unit)

    def read(self, start: int, end: int) -> None:
        n = len(self.io_vov)
        if start >= n:
            raise EndExecute
        end = min(end, n)
        self._ensure_raw(start, end)
        cl = np.asarray(self.io_vov.cumulative_length.nda).astype(np.int64)
        lo = np.concatenate([[0 if start == 0 else cl[start - 1]], cl[start : end - 1]])
        lens = cl[start:end] - lo
        width = self.raw_var.shape[1]
        if len(lens) and lens.max() > width:
            raise DSPFatal("VectorOfVectors entry has length larger than array variable length")
        flat = np.asarray(self.io_vov.flattened_data.nda)
        fill = 0 if np.issubdtype(self.var.dtype, np.integer) else np.nan
        pad = np.full((end - start, width), fill, dtype=self.var.dtype)
        mask = np.arange(width)[None, :] < lens[:, None]
        pad[mask] = flat[int(lo[0]) if len(lo) else 0 : int(cl[end - 1]) if end > start else 0]
        self.raw_var[: end - start].copy_(torch.from_numpy(pad))
        self.len_var[: end - start].copy_(torch.from_numpy(lens.astype(np.uint32).astype(np.int64)).to(self.len_var.dtype))
        self._count(pad.nbytes, False)

    def write(self, start: int, end: int) -> None:
        """The padded [block, width] buffer is compacted ON THE DEVICE (csrc/sipm.cu: scan of the lengths -> end offsets
        and `cumulative_length`, one warp per row scatters its entries); only the ragged data and the offsets cross to
        the host column -- or nothing at all when the column's arrays are device tensors."""
        self._ensure_raw(start, end)
        nrow = end - start
        if nrow <= 0:
            return
        blk = self.raw_var[:nrow]
        dev = blk.device
        lens = self.len_var[:nrow]
        if lens.dtype != torch.uint32:
            lens = lens.to(torch.int64).clamp_(min=0).to(torch.uint32)
        lens = lens.contiguous()
        if len(self.io_vov) < end:
            self.io_vov.resize(end)
        cl = self.io_vov.cumulative_length.nda
        base = int(cl[start - 1]) if start > 0 else 0
        L = _lib.lib()
        stream = _vp(torch.cuda.current_stream(dev).cuda_stream)
        ends = torch.empty(nrow, dtype=torch.int64, device=dev)
        cum = torch.empty(nrow, dtype=torch.uint32, device=dev)
        width = int(blk.shape[1]) if blk.ndim == 2 else 1
        rc = L.dspb_vov_offsets(_vp(lens.data_ptr()), _i64(nrow), _i64(width), _i64(base), _vp(ends.data_ptr()),
                                _vp(cum.data_ptr()), stream)
        if rc:
            raise RuntimeError(f"dspb_vov_offsets: error {rc}")
        total = int(ends[-1].item()) - base          # (one 8-byte read back: the size of the ragged payload)
        flat_d = torch.empty(max(total, 1), dtype=blk.dtype, device=dev)
        rc = L.dspb_vov_compact(_vp(blk.data_ptr()), _i64(blk.stride(0)), _i32(blk.element_size()), _vp(lens.data_ptr()),
                                _i64(width), _vp(ends.data_ptr()), _i64(nrow), _i64(base), _vp(flat_d.data_ptr()), stream)
        if rc:
            raise RuntimeError(f"dspb_vov_compact: error {rc}")
        need = base + total
        if need > len(self.io_vov.flattened_data):
            self.io_vov.flattened_data.resize(need)
        fd = _as_tensor(self.io_vov.flattened_data.nda)
        fd[base:need].copy_(flat_d[:total])
        _as_tensor(cl)[start:end].copy_(cum)
        if not fd.is_cuda:
            self._count(total * blk.element_size() + 4 * nrow, True)

    def __str__(self) -> str:
        return f"{self.var} linked to VectorOfVectors(vector_len={self.var.vector_len}, attrs={self.io_vov.attrs})"


class WaveformIOManager(IOManager):
    """``WaveformTable``: ``values`` block + per-event ``t0`` offset variable + ``dt``
    period (reference :2263-2360)"""

    def __init__(self, wf_table, variable: ProcChainVar) -> None:
        dt_units = getattr(wf_table, "dt_units", None)
        t0_units = getattr(wf_table, "t0_units", None)
        if dt_units is None:
            dt_units = t0_units
        elif t0_units is None:
            t0_units = dt_units
        self.wf_var = variable
        self.var = variable
        if (self.wf_var.grid is auto and isinstance(dt_units, str) and dt_units in ureg
                and isinstance(t0_units, str) and t0_units in ureg):
            if not len(wf_table.dt):
                # (the sampling period is the first entry of the per-event `dt` column, reference :2286)
                raise ProcessingChainError(f"waveform table of {variable.name} is empty: it carries no sampling period")
            dt0 = float(np.asarray(_host_view(wf_table.dt.nda)[0:1])[0])
            self.wf_var.update_auto(
                grid=CoordinateGrid(
                    ureg.Quantity(dt0, dt_units),
                    ProcChainVar(self.wf_var.proc_chain, self.wf_var.name + "_dt", shape=(),
                                 dtype=tables.np_dtype_of(wf_table.t0.nda), grid=None, unit=dt_units, is_coord=True)),
                is_coord=False)
        else:
            self.wf_var.update_auto(grid=None, is_coord=False)

        values = tables.wf_values(wf_table)
        if kind_of(values) == "vov":
            self.val_ioman = VectorOfVectorsIOManager(values, self.wf_var)
        else:
            self.val_ioman = ArrayIOManager(values, self.wf_var)
        if dt_units is None:
            dt_units = self.wf_var.grid.unit_str()
            t0_units = self.wf_var.grid.unit_str()
        self.t0_var = self.wf_var.grid.get_offset(t0_units)
        self.variable_t0 = isinstance(self.t0_var, torch.Tensor)
        self.set_buffer(wf_table)

    def set_buffer(self, wf_table) -> None:
        if "units" not in wf_table.attrs and self.wf_var.unit is not None:
            u = self.wf_var.unit
            wf_table.attrs["units"] = str(u.u) if isinstance(u, Quantity) else str(u)
        self.io_wf = wf_table
        self.val_ioman.set_buffer(tables.wf_values(wf_table))
        self._is_output_target = False

    def prepare_output(self) -> None:
        """stamp the grid onto an output table (period, constant offset, units)"""
        if not self.variable_t0:
            self.io_wf.t0.nda[...] = float(self.wf_var.offset / Quantity(1.0, self.wf_var.period.u))
        dt_units = self.wf_var.period.u
        self.io_wf.dt.nda[...] = self.wf_var.grid.get_period(Quantity(1.0, dt_units))
        self.io_wf.dt_units = str(dt_units)
        self.io_wf.t0_units = str(dt_units)

    def read(self, start: int, end: int) -> None:
        n = len(self.io_wf)
        if start >= n:
            raise EndExecute
        end = min(end, n)
        self.val_ioman.read(start, end)
        if self.variable_t0:
            src = _as_tensor(self.io_wf.t0.nda)[start:end]
            self.t0_var[0 : end - start].copy_(src, non_blocking=True)

    def write(self, start: int, end: int, t0_src=None) -> None:
        """`t0_src`: the per-event offsets of rows [start, end) when the caller did not fill the block's offset
        variable (the fused tiers never run the input managers' read(): they hand over the input column's rows)"""
        if len(self.io_wf) < end:
            self.io_wf.resize(end)
        if not self._is_output_target:
            self.prepare_output()
            self._is_output_target = True
        self.val_ioman.write(start, end)
        if self.variable_t0:
            src = self.t0_var[0 : end - start] if t0_src is None else t0_src
            _as_tensor(self.io_wf.t0.nda)[start:end].copy_(src, non_blocking=True)

    def __str__(self) -> str:
        return f"{self.wf_var} linked to WaveformTable(values({self.val_ioman}))"


def _host_view(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


# ======================================================================================
# recipe compiler
# ======================================================================================
_MODULE_ALIASES = {
    "dspeed.processors": "dspeed_b200.processors",
    "pygama.dsp.processors": "dspeed_b200.processors",
    "dspeed_b200.processors": "dspeed_b200.processors",
}


def _resolve_function(module_name: str, func_name: str):
    """processor callable named by a recipe.  ``dspeed.processors.X`` resolves to the
    CUDA processor of the same name; numpy / scipy functions are resolved as such and
    mapped to a device implementation when bound (or rejected: no CPU fallback)."""
    # (the reference's configs also name the processor's own sub-module, e.g. "dspeed.processors.histogram")
    if module_name in _MODULE_ALIASES or any(module_name.startswith(m + ".") for m in _MODULE_ALIASES):
        try:
            return getattr(device_processors, func_name)
        except AttributeError as e:
            raise ProcessingChainError(str(e)) from e
    module = importlib.import_module(module_name)
    return getattr(module, func_name)


def build_processing_chain(processors, tb_in=None, db_dict=None, outputs=None, block_width=None, device=None):
    """Compile a JSON/YAML/dict recipe into a :class:`ProcessingChain` (same schema and semantics as the reference's
    build_processing_chain, :2363-2872).  Parsing and scheduling live in :mod:`dspeed_b200.recipe`; this function
    instantiates the scheduled steps on a chain and links the table columns.

    Returns ``(proc_chain, field_mask, tb_out)``."""
    n_rows = len(tb_in) if tb_in is not None else 1
    chain = ProcessingChain(block_width, n_rows, device=device)
    steps, config_outputs, db = recipe.parse(processors, db_dict, lambda text: chain.get_variable(text, True),
                                             ProcessingChain.module_list, ProcessingChain.func_list)
    if outputs is None:
        if config_outputs is None:
            raise ValueError("outputs not provided")
        outputs = config_outputs
    ordered, columns, copies, produced = recipe.schedule(steps, outputs)

    for name in columns:
        if tb_in is None or name not in tb_in:
            log.warning(f"'{name}' not found in input files or dsp config.")
        try:
            chain.link_input_buffer(name, tb_in[name])
        except Exception as e:
            raise ProcessingChainError(f"Exception raised while linking input buffer '{name}'.") from e

    for step in ordered:
        try:
            _instantiate(chain, step, db)
        except Exception as e:
            raise ProcessingChainError("Exception raised while attempting to add processor:\n" + recipe.describe(step)) from e

    tb_out = tables.Table(size=n_rows)
    for name in copies:          # requested names that no step produces: copied through from the input table
        if tb_in is None or name not in tb_in:
            log.warning(f"'{name}' not found in input files or dsp . Building output without it!")
            continue
        try:
            chain.link_input_buffer(name, tb_in[name])
            col = chain.link_output_buffer(name)
            col.attrs.update(getattr(tb_in[name], "attrs", {}))
            col.resize(len(tb_out))
            tb_out.add_field(name, col)
        except Exception as e:
            raise ProcessingChainError(f"Exception raised while linking copy buffer '{name}'.") from e
    for name in produced:
        try:
            col = chain.link_output_buffer(name)
            entry = steps[name].entry
            col.attrs.update(entry.get("lh5_attrs", {}))
            if text := entry.get("description"):
                col.attrs["description"] = text
            col.resize(len(tb_out))
            tb_out.add_field(name, col)
        except Exception as e:
            raise ProcessingChainError(f"Exception raised while linking output buffer {name}.") from e

    chain.recipe_info = {"proc_par_list": [st.key for st in ordered], "outputs": list(outputs)}
    if chain.device.type == "cuda" and os.environ.get("DSPEED_B200_FUSE", "1") != "0":
        from . import fusion

        fusion.try_fuse(chain)
    return (chain, columns + copies, tb_out)


def _instantiate(chain: ProcessingChain, step, db) -> None:
    """one scheduled recipe step -> variables + (a launch descriptor | folded constants | an alias)"""
    entry = step.entry
    if step.module is None:
        # a bare expression: the grammar builds whatever processors it needs; the name becomes an alias of the result
        value = chain.get_variable(step.args[0])
        if isinstance(value, ProcChainVar):
            var = chain.add_variable(name=step.key, dtype=value.dtype, shape=value.shape, grid=value.grid, unit=value.unit,
                                     is_coord=value.is_coord)
            var._buffer = value._buffer
            var.alias_of = value
        else:
            chain.set_constant(varname=step.key, val=value)
        return

    func = _resolve_function(step.module, step.function)
    if "unit" in entry:
        for i, name in enumerate(step.outputs):
            unit = entry["unit"]
            chain.add_variable(name, unit=unit[i] if isinstance(unit, list) else unit)
    options = dict(entry.get("kwargs", {}))
    options.update({k: entry[k] for k in ("signature", "types", "coord_grid") if k in entry})

    if "init_args" in entry:        # the named callable is a factory (iir_filter, ...): call it with these first
        pos, named = [], {}
        for arg in entry["init_args"]:
            if isinstance(arg, str):
                arg = db.substitute(arg, entry)
                if isinstance(arg, str):
                    arg = chain.get_variable(arg)
            if isinstance(arg, MutableMapping):
                named.update(arg)
            else:
                pos.append(arg)
        func = func(*pos, **named)

    operands, keyword_operands, results = [], {}, []
    all_const = True
    for arg in step.args:
        value = chain.get_variable(arg) if isinstance(arg, str) else arg
        if isinstance(value, MutableMapping):      # name=value in the argument list
            keyword_operands.update(value)
            value = list(value.values())[0]
        else:
            operands.append(value)                  # variables, numbers, quantities, string literals ('s', 'n', ...)
        if isinstance(value, ProcChainVar):
            if value.name in step.outputs:
                results.append(value)
            elif not value.is_const:
                all_const = False

    if not all_const:
        grid = options.get("coord_grid")
        man = ProcessorManager(chain, func, operands, keyword_operands, options.get("signature"), options.get("types"),
                               CoordinateGrid(grid) if grid is not None else None)
        chain._proc_managers.append(man)
        log.debug(f"added processor: {man}")
    elif results:
        # constant folding: every input is a constant, so the processor runs once now and its outputs are constants
        # (cusp / zac / t0 / gaussian kernels); the provenance lets the chain compilers recognise structured kernels
        for var in results:
            var.is_const = True
        man = ProcessorManager(chain, func, operands, keyword_operands, options.get("signature"), options.get("types"))
        if chain.device.type != "meta":
            man.execute()
            chain._raise_recorded_fatal(0, 0)
        for var in results:
            var.const_origin = (man.processor.__name__, list(man.args))
    else:
        values = func(*operands, **keyword_operands)
        for name, val in zip(step.outputs, [values] if len(step.outputs) == 1 else values):
            chain.set_constant(name, val)
