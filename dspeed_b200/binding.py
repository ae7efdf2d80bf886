"""Launch-descriptor compiler: binds the operands of ONE processor call of a recipe to device buffers.

A dspeed processor is a generalised ufunc: its layout string (``"(n),(),(m)->(n)"``) names the *core* dimensions of
every operand, everything in front of them is looped over.  In a processing chain the outermost loop dimension is
always the block of events; a variable's own shape is its per-event shape.  Compiling a call therefore means

1. choosing the type loop -- the first entry of the processor's type table every already-typed operand can be cast to
   (what numpy's ufunc machinery would pick);
2. solving the shapes -- every shaped operand contributes *bindings* (core dimension name -> extent) and a *loop shape*
   (its leading per-event axes, broadcast against the others numpy-style); still-untyped / unshaped (``auto``) output
   variables then take their shape, dtype, sampling grid and coordinate flag from the solution;
3. lowering the operands -- variables become views of their block buffers with singleton axes where they broadcast,
   unit-carrying scalars are expressed in samples of the call's sampling grid, strings become byte codes, host arrays
   are uploaded once.

The result (``args`` / ``kwargs``) is what every execution tier consumes: the per-processor kernels take it as is, the
chain compilers (``fusion``, ``codegen``, ``warpchain``) lower it further into one kernel.  Semantics follow the
reference's ``ProcessorManager`` (src/dspeed/processing_chain.py:1485-1803) -- same type choice, same unit -> sample
conversion, same deduction of ``auto`` variables -- so that existing recipes compile to the same plan
(tests/test_chain_plan.py); the formulation (constraint collection + broadcast, instead of the reference's
dimension-list rewriting) is this repository's.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch

from .errors import ProcessingChainError
from .units import Quantity, Unit, is_in_registry, to_period_units, ureg

_OPERAND = re.compile(r"\((.*?)\)")


def core_dims(layout: str) -> list[list[str]]:
    """``"(n),(),(m, k)->(n)"`` -> ``[["n"], [], ["m", "k"], ["n"]]``"""
    return [[d.strip() for d in grp.split(",") if d.strip()] for grp in _OPERAND.findall(layout)]


def scalar_layout(nin: int, nout: int) -> str:
    return ",".join(["()"] * nin) + "->" + ",".join(["()"] * nout)


def pick_type_loop(type_table, operands, describe) -> list[np.dtype]:
    """first type-table entry (``"fii->f"``) that every typed operand casts to safely"""
    if isinstance(type_table, str) or not hasattr(type_table, "__iter__"):
        type_table = [type_table]
    table = [t.replace("->", "") for t in type_table]
    if not table:
        raise ProcessingChainError(f"could not find a type signature list for {describe}. Please supply a valid list of types.")
    for pos, opnd in enumerate(operands):
        dt = getattr(opnd, "dtype", None)
        if dt is None or not _is_concrete(dt):
            continue
        have = np.dtype(dt).char
        table = [t for t in table if np.can_cast(have, t[pos])]
    if not table:
        raise ProcessingChainError(f"could not find a type signature matching the types of the variables given for "
                                   f"{describe} (types: {list(type_table)})")
    return [np.dtype(c) for c in table[0]]


def _is_concrete(x) -> bool:
    return not (isinstance(x, str) and x == "auto") and x is not None and type(x).__name__ != "_Auto"


@dataclass
class Axis:
    """one solved dimension: its extent and the sampling grid of the data laid out along it (if any)"""
    extent: int
    grid: Any = None


@dataclass
class ShapeSolution:
    core: dict = field(default_factory=dict)       # core dimension name -> Axis
    loop: list = field(default_factory=list)       # per-event loop axes (outermost first), broadcast over all operands

    def axes_of(self, names) -> list[Axis]:
        return list(self.loop) + [self.core[n] for n in names]


def solve_shapes(layouts, operands, block_width, describe) -> tuple[ShapeSolution, Any]:
    """Collect the core-dimension bindings and the common loop shape from every operand whose shape is known.
    Returns the solution and the sampling grid of the call (the grid of the first gridded, non-coordinate operand)."""
    sol = ShapeSolution()
    call_grid = None
    for names, opnd in zip(layouts, operands):
        shape = getattr(opnd, "shape", None)
        if shape is None or not _is_concrete(shape) or not hasattr(opnd, "dtype"):
            continue
        shape = tuple(int(s) for s in shape)
        grid = getattr(opnd, "grid", None)
        if not _is_concrete(grid) or getattr(opnd, "is_coord", False):
            grid = None      # (coordinates -- and variables not yet known not to be one -- lend no grid)
        if call_grid is None and grid:
            call_grid = grid
        k = len(names)
        if len(shape) < k:
            # fewer axes than core dimensions: the operand's missing leading axis is the block itself (a per-event
            # scalar handed to a waveform argument); its remaining axes bind the trailing names
            lead, shape_core = names[: k - len(shape)], shape
            if len(lead) > 1:
                raise ProcessingChainError(f"failed to broadcast array dimensions for {describe}: {opnd} has too few axes")
            _bind(sol, lead[0], block_width, None, describe)
            names_core = names[k - len(shape):]
            loop_part = ()
        else:
            names_core, shape_core, loop_part = names, shape[len(shape) - k:], shape[: len(shape) - k]
        for j, (nm, ext) in enumerate(zip(names_core, shape_core)):
            # only the innermost axis of an operand is sampled on its grid
            _bind(sol, nm, ext, grid if j == len(names_core) - 1 else None, describe)
        _broadcast_loop(sol, loop_part, grid if not names_core else None, opnd, describe)
    return sol, call_grid


def _bind(sol: ShapeSolution, name: str, extent: int, grid, describe):
    ax = sol.core.get(name)
    if ax is None:
        sol.core[name] = Axis(extent, grid)
    elif ax.extent != extent:
        raise ProcessingChainError(f"failed to broadcast array dimensions for {describe}. Could not find consistent value "
                                   f"for dimension {name}")
    elif not ax.grid:
        ax.grid = grid


def _broadcast_loop(sol: ShapeSolution, part, grid, opnd, describe):
    """numpy broadcasting of an operand's leading per-event axes against the loop shape found so far"""
    part = list(part)
    while len(sol.loop) < len(part):
        sol.loop.insert(0, Axis(1))
    offset = len(sol.loop) - len(part)
    for j, ext in enumerate(part):
        ax = sol.loop[offset + j]
        if ax.extent == 1:
            ax.extent = ext
        elif ext not in (1, ax.extent):
            raise ProcessingChainError(f"failed to broadcast array dimensions for {describe}. Input arrays do not have "
                                       f"consistent outer dimensions; found {tuple(getattr(opnd, 'shape', ()))} for {opnd}")
        if j == len(part) - 1 and grid and not ax.grid:
            ax.grid = grid


def block_view(buf: torch.Tensor, target_shape) -> torch.Tensor:
    """view of a block buffer ``[block, ...]`` with singleton axes inserted where it broadcasts against
    ``[block, *target_shape]``"""
    have = list(buf.shape)
    want = list(target_shape)
    out = list(have)
    for back in range(1, len(want) + 1):
        if len(out) < back or out[-back] != want[-back]:
            out.insert(len(out) - back + 1, 1)
    return buf if out == have else buf.reshape(out)


def scalar_in_samples(value, dtype: np.dtype, grid, grid_type):
    """python / unit-carrying scalar -> numpy scalar of the loop's type; quantities with a time-like unit are expressed
    in samples of `grid` (``10*us`` on a 16 ns grid -> 625); integer slots round to nearest"""
    if isinstance(value, Unit):
        value = Quantity(1.0, value)
    if isinstance(value, Quantity):
        if value.u.dimensionless:
            value = float(value)
        elif not isinstance(grid, grid_type):
            raise ProcessingChainError(f"could not find valid conversion for {value}; CoordinateGrid is {grid}")
        else:
            try:
                value = to_period_units(value, grid.period)
            except ValueError as e:
                raise ProcessingChainError(str(e)) from e
    return dtype.type(np.round(value)) if np.issubdtype(dtype, np.integer) else dtype.type(value)


def text_operand(text: str, dtype: np.dtype, shape):
    """a string bound to an integer slot is its ASCII codes (``'n'`` -> 110); anything else stays text"""
    if not np.issubdtype(dtype, np.integer):
        return text
    try:
        codes = np.frombuffer(text.encode("ascii"), dtype).reshape(shape)
    except ValueError:
        raise ProcessingChainError(f"could not convert string '{text}' into byte-array of type {dtype} and shape {shape}")
    return int(codes.reshape(-1)[0]) if codes.size == 1 else codes


def coordinate_role(var, call_grid):
    """(unit, is_coord, grid override) a variable takes from the call's grid: a variable that already is a coordinate
    is read in the grid's period units; a still-undecided variable whose unit is a time (compatible with the period)
    becomes a coordinate on that grid"""
    if var.is_coord is True and call_grid is not None:
        return str(call_grid.period.u), False, call_grid
    if is_in_registry(var.unit) and call_grid is not None and ureg.is_compatible_with(call_grid.period, var.unit):
        return None, True, call_grid
    return None, False, None
