"""The expression language of a dspeed configuration (argument texts such as ``"wf_blsub[0:750]"``,
``"tp_max - 2*us"``, ``"where(a > b, a, 1*ns)"``, ``"vt_max(20, 'f', vector_len=n_max)"``, ``"kw=expr"``).

What the language means is the reference's (src/dspeed/processing_chain.py:718-1130: names create variables on first
use, calls to unknown names declare variables, arithmetic on variables emits element-wise processors, subscripts are
views that share the parent's buffer, units are names of the unit registry); how it is processed here is different.
The text is handled by TWO independent passes over its Python AST instead of one evaluator with a dry-run flag:

``names_read``  a static walk that lists the variable names a text mentions, in evaluation order -- used by the recipe
                compiler to find dependencies before anything exists (no chain state is touched, nothing is created);
``Evaluator``   a dispatch-table interpreter (one handler per AST node type) that builds variables, views and glue
                processors on a chain.

Both take their vocabulary -- callable helper names, module aliases, unit names -- from the chain class, so the two passes
cannot disagree about what is a name.
"""
from __future__ import annotations

import ast
import re
from numbers import Real

import numpy as np

from . import numpy_bridge
from .errors import ProcessingChainError
from .units import Quantity, Unit, as_unit, from_foreign, is_in_registry, ureg

#: AST operator -> (element-wise ufunc the processor manager binds, text template of the result's name)
OPERATORS = {
    ast.Add: (np.add, "{}+{}"),
    ast.Sub: (np.subtract, "{}-{}"),
    ast.Mult: (np.multiply, "{}*{}"),
    ast.Div: (np.divide, "{}/{}"),
    ast.FloorDiv: (np.floor_divide, "{}//{}"),
    ast.USub: (np.negative, "-{}"),
    ast.Eq: (np.equal, "{}=={}"),
    ast.NotEq: (np.not_equal, "{}!={}"),
    ast.Lt: (np.less, "{}<{}"),
    ast.LtE: (np.less_equal, "{}<={}"),
    ast.Gt: (np.greater, "{}>{}"),
    ast.GtE: (np.greater_equal, "{}>={}"),
}
_SCALING = (ast.Mult, ast.Div, ast.FloorDiv)     # operators whose result unit is the combination of both units

_PLAIN = {np.add: lambda a, b: a + b, np.subtract: lambda a, b: a - b, np.multiply: lambda a, b: a * b,
          np.divide: lambda a, b: a / b, np.floor_divide: lambda a, b: a // b,
          np.equal: lambda a, b: a == b, np.not_equal: lambda a, b: not (a == b),
          np.less: lambda a, b: a < b, np.less_equal: lambda a, b: a <= b,
          np.greater: lambda a, b: a > b, np.greater_equal: lambda a, b: a >= b}


def fold(op, lhs, rhs):
    """an operator applied to two build-time constants (numbers, units, quantities)"""
    lhs, rhs = from_foreign(lhs), from_foreign(rhs)
    if not isinstance(lhs, (Quantity, Unit)) and not isinstance(rhs, (Quantity, Unit)):
        return op(lhs, rhs)
    ret = _PLAIN[op](lhs, rhs)
    if isinstance(ret, Unit):
        ret = Quantity(1.0, ret)
    if isinstance(ret, Quantity) and ret.u.dimensionless:
        ret = float(ret)
    return ret


def is_variable_name(text: str, vocabulary) -> bool:
    return bool(re.match(r"\A\w+$", text)) and text not in vocabulary.func_list and text not in ureg \
        and text not in vocabulary.module_list


def _source(node, text: str) -> str:
    return text[node.col_offset: node.end_col_offset]


# ----------------------------------------------------------------------------------------------------------------------
# pass 1: names
# ----------------------------------------------------------------------------------------------------------------------
def names_read(tree, text: str, vocabulary) -> list[str]:
    """variable names mentioned by the expression `tree` (parsed from `text`), in the order an evaluation meets them"""
    found: list[str] = []

    def walk(node):
        if node is None or isinstance(node, (ast.Constant, ast.List)):
            return
        if isinstance(node, ast.Name):
            if node.id not in ureg:
                found.append(node.id)
        elif isinstance(node, ast.BinOp):
            walk(node.left)
            walk(node.right)
        elif isinstance(node, ast.UnaryOp):
            walk(node.operand)
        elif isinstance(node, ast.Compare):
            if len(node.comparators) != 1:
                raise ProcessingChainError("Compound comparisons are not supported.")
            walk(node.left)
            walk(node.comparators[0])
        elif isinstance(node, ast.IfExp):
            walk(node.test)
            walk(node.body)
            walk(node.orelse)
        elif isinstance(node, ast.Subscript):
            walk(node.value)     # (indices are constants or already-scheduled variables: the reference's name pass
            #                       does not descend into them either, processing_chain.py:948-950)
        elif isinstance(node, ast.Attribute):
            if _source(node.value, text) not in vocabulary.module_list:
                walk(node.value)
        elif isinstance(node, ast.Call):
            for a in node.args:
                walk(a)
            for kw in node.keywords:
                walk(kw.value)
            callee = getattr(node.func, "id", None)
            if callee in vocabulary.func_list:
                return
            if callee is None or not is_variable_name(callee, vocabulary):
                raise ProcessingChainError(f"do not recognize call to {callee}")
            found.append(callee)     # name(shape, dtype, ...) declares / refines a variable
        else:
            raise ProcessingChainError(f"cannot parse AST nodes of type {type(node).__name__}")

    walk(tree)
    return found


# ----------------------------------------------------------------------------------------------------------------------
# pass 2: evaluation
# ----------------------------------------------------------------------------------------------------------------------
class Evaluator:
    """evaluates one expression text on `chain`: returns a variable, a view, a constant, a unit / quantity, or (for
    ``kw=expr`` statements, handled by the chain) the value of the right-hand side"""

    def __init__(self, chain, text: str):
        self.chain = chain
        self.text = text
        self.Var = type(chain).Variable       # the chain's variable class (ProcChainVar)
        self.Grid = type(chain).Grid

    def __call__(self, node):
        if node is None:
            return None
        handler = getattr(self, "_on_" + type(node).__name__, None)
        if handler is None:
            raise ProcessingChainError(f"cannot parse AST nodes of type {type(node).__name__}")
        return handler(node)

    def _is_var(self, x) -> bool:
        return isinstance(x, self.Var)

    # -- leaves --------------------------------------------------------------------------------------------------------
    def _on_List(self, node):
        return np.array(ast.literal_eval(_source(node, self.text)))

    def _on_Constant(self, node):
        return node.value

    def _on_Name(self, node):
        if node.id in ureg:
            return ureg(node.id)
        known = self.chain._vars_dict.get(node.id)
        return known if known is not None else self.chain.add_variable(node.id)

    def _on_Attribute(self, node):
        prefix = _source(node.value, self.text)
        modules = self.chain.module_list
        if prefix in modules:
            value = getattr(modules[prefix], node.attr)
            if not isinstance(value, Real):
                raise ProcessingChainError(f"Attribute {node.attr} from {prefix} is not an int or float...")
            return value
        return getattr(self(node.value), node.attr)

    # -- operators -----------------------------------------------------------------------------------------------------
    def _result_unit(self, node_op, op, template, lhs, rhs):
        """unit of ``lhs <op> rhs`` for two variables: registry units combine algebraically, free-text units textually"""
        lu, ru = lhs.unit, rhs.unit
        if is_in_registry(lu) and is_in_registry(ru):
            unit = op(Quantity(1.0, as_unit(lu)), Quantity(1.0, as_unit(ru))).u
            return None if unit.dimensionless else unit
        if lu is not None and ru is not None:
            return template.format(str(lu), str(ru)) if type(node_op) in _SCALING else str(lu)
        return lu if lu is not None else ru

    def _on_BinOp(self, node):
        lhs, rhs = self(node.left), self(node.right)
        op, template = OPERATORS[type(node.op)]
        lvar, rvar = self._is_var(lhs), self._is_var(rhs)
        if not lvar and not rvar:
            return fold(op, lhs, rhs)
        name = "(" + template.format(str(lhs), str(rhs)) + ")"
        if lvar and rvar:
            both_coords = lhs.is_coord is True and rhs.is_coord is True     # a difference of two times is a duration
            out = self.Var(self.chain, name, unit=self._result_unit(node.op, op, template, lhs, rhs),
                           grid=None if lhs.is_coord and rhs.is_coord else "auto",
                           is_coord=False if both_coords else "auto")
        else:
            v = lhs if lvar else rhs
            out = self.Var(self.chain, name, unit=v.unit, is_coord=v.is_coord)
        self.chain._emit(op, [lhs, rhs, out])
        return out

    def _on_UnaryOp(self, node):
        operand = self(node.operand)
        op, template = OPERATORS[type(node.op)]
        if not self._is_var(operand):
            return op(operand)
        out = self.Var(self.chain, "(" + template.format(str(operand)) + ")", operand.shape, operand.dtype, operand.grid,
                       operand.unit, operand.is_coord)
        self.chain._emit(op, [operand, out])
        return out

    def _on_Compare(self, node):
        if len(node.comparators) != 1:
            raise ProcessingChainError("Compound comparisons are not supported.")
        lhs, rhs = self(node.left), self(node.comparators[0])
        op, template = OPERATORS[type(node.ops[0])]
        if not self._is_var(lhs) and not self._is_var(rhs):
            return fold(op, lhs, rhs)
        out = self.Var(self.chain, "(" + template.format(str(lhs), str(rhs)) + ")")
        self.chain._emit(op, [lhs, rhs, out])
        return out

    def _on_IfExp(self, node):
        return self.chain._where(self(node.test), self(node.body), self(node.orelse))

    # -- calls: helper functions and variable declarations -------------------------------------------------------------
    def _on_Call(self, node):
        callee = getattr(node.func, "id", None)
        args = [self(a) for a in node.args]
        kwargs = {kw.arg: self(kw.value) for kw in node.keywords}
        helper = self.chain.func_list.get(callee)
        if helper is not None:
            return helper(self.chain, *args, **kwargs)
        if callee is None or not is_variable_name(callee, self.chain):
            raise ProcessingChainError(f"do not recognize call to {callee}")
        known = self.chain._vars_dict.get(callee)
        if known is not None:
            known.update_auto(*args, **kwargs)
            return known
        # positional declaration arguments follow the variable's own order: shape, dtype, grid, unit, is_coord
        declared = dict(zip(("shape", "dtype", "grid", "unit", "is_coord"), args))
        return self.chain.add_variable(callee, **{**declared, **kwargs})

    # -- subscripts: views ---------------------------------------------------------------------------------------------
    def _index(self, node, parent, wrap_len=None):
        """one index / slice bound: a constant number of samples (quantities are divided by the parent's period,
        non-integers rounded with a warning), or a per-event variable; negative constants wrap around `wrap_len`"""
        value = self(node)
        if value is None or self._is_var(value):
            return value
        if isinstance(value, Quantity):
            value = float(value / parent.period)
        if isinstance(value, Real):
            nearest = int(round(value))
            if abs(value - nearest) > 0.0001:
                import logging

                logging.getLogger("dspeed").warning(f"slice value is non-integer. Rounding to {nearest}")
            value = nearest
        if value < 0 and wrap_len is not None:
            value = self.chain.get_variable(f"{wrap_len}{value}")
        return value

    def _on_Subscript(self, node):
        parent = self(node.value)
        if not self._is_var(parent) or not len(parent.shape) > 0:
            raise ProcessingChainError("Cannot apply subscript to", node.value)
        if isinstance(node.slice, ast.Tuple):
            raise ProcessingChainError("Tuple still isn't implemented...")
        if isinstance(node.slice, ast.Slice):
            return self._slice_view(node, parent)
        index = self._index(node.slice, parent, parent.vector_len)
        if isinstance(index, int):
            return self._make_view(node, parent, parent.buffer[..., index], f"{parent}[{index}]",
                                   parent.grid if parent.is_coord else None)
        # a per-event index: one element read per event (get.py:10-91), NaN / integer-max beyond the end
        out = self.Var(self.chain, name=f"{parent}[{index}]", shape=(), dtype=parent.dtype,
                       grid=parent.grid if parent.is_coord else None, unit=parent.unit, is_coord=parent.is_coord)
        beyond = np.nan if np.issubdtype(parent.dtype, np.floating) else np.iinfo(parent.dtype).max
        self.chain._emit(numpy_bridge.get_default, [parent, index, beyond, out])
        return out

    def _slice_view(self, node, parent):
        lo, hi, step = (self._index(part, parent) for part in (node.slice.lower, node.slice.upper, node.slice.step))
        if any(self._is_var(x) for x in (lo, hi, step)):
            raise ProcessingChainError("Slice values must be constants")

        def txt(x, lead=""):
            return "" if x is None else lead + str(x)

        name = f"{parent}[{txt(lo)}:{txt(hi)}{txt(step, ':')}]"
        grid = None
        if parent.grid is not None:
            # the view is sampled `step` times coarser and starts `lo` samples later
            period = parent.period if step is None else parent.period * step
            offset = parent.offset
            if lo is not None and lo > 0:
                shift = lo * parent.period
                if self._is_var(offset):
                    moved = self.Var(self.chain, name=f"({offset}+{shift})", is_coord=True)
                    self.chain._emit(np.add, [offset, shift, moved])
                    offset = moved
                else:
                    offset = offset + shift
            grid = self.Grid(period, offset)
        return self._make_view(node, parent, parent.buffer[..., slice(lo, hi, step)], name, grid)

    def _make_view(self, node, parent, storage, name, grid):
        view = self.Var(self.chain, name, shape=tuple(storage.shape[1:]), dtype=parent.dtype, grid=grid,
                        unit=parent.unit, is_coord=parent.is_coord)
        view._buffer = [(storage, parent._buffer[0][1])] if view.is_coord else storage
        view.view_of = (parent, node.slice)
        return view


# ----------------------------------------------------------------------------------------------------------------------
# helper functions callable inside expressions: len, isnan, isfinite, astype, round / floor / ceil / trunc, where,
# loadlh5 (semantics: src/dspeed/processing_chain.py:1177-1482).  Each takes the chain first; HELPERS is the table the
# chain class publishes as `func_list`.
# ----------------------------------------------------------------------------------------------------------------------
def _like(chain, name, var, dtype=None, grid="same"):
    """a new variable with the layout and metadata of `var` (optionally another dtype / grid)"""
    return type(chain).Variable(chain, name, var.shape, var.dtype if dtype is None else dtype,
                                var.grid if isinstance(grid, str) and grid == "same" else grid, var.unit, var.is_coord)


def h_len(chain, var):
    if var is None:
        return None
    if not isinstance(var, type(chain).Variable):
        raise ProcessingChainError(f"cannot call len() on {var}")
    if var.vector_len is not None:        # variable-length vectors: their per-event length variable
        return var.vector_len
    if len(var.shape) != 1:
        raise ProcessingChainError(f"{var} has wrong number of dims")
    return var.shape[0]


def _predicate(ufunc, label):
    def helper(chain, var):
        if var is None:
            return None
        if not isinstance(var, type(chain).Variable):
            return ufunc(var)
        out = _like(chain, f"{label}({var})", var, dtype="bool")
        chain._emit(ufunc, [var, out])
        return out

    return helper


def h_astype(chain, var, dtype):
    dtype = np.dtype(dtype)
    if var is None:
        return None
    if not isinstance(var, type(chain).Variable):
        raise ProcessingChainError(f"cannot call astype() on {var}")
    out = _like(chain, f"{var}.astype(`{dtype.char}`)", var, dtype=dtype)
    chain._emit(numpy_bridge.make_astype(var.dtype, dtype), [var, out])
    return out


def _rounder(mode):
    """round / floor / ceil / trunc of a constant (host arithmetic), of a value variable (element-wise kernel) or of a
    coordinate (a unit conversion onto the coarser grid)"""
    if mode not in numpy_bridge.HOST_ROUNDERS:
        raise ProcessingChainError("Mode must be round, floor, ceil or trunc")

    def helper(chain, var, to_nearest=1, dtype=None):
        if var is None:
            return None
        step = from_foreign(to_nearest)
        Var, Grid = type(chain).Variable, type(chain).Grid
        if not isinstance(var, Var):
            host = numpy_bridge.HOST_ROUNDERS[mode]
            if isinstance(var, Quantity) and isinstance(step, Quantity):
                one = Quantity(1.0, step.u)
                return host(float(var / one), step.m) * one
            return host(float(var) if isinstance(var, Quantity) else var, step)
        name = f"{mode}({var}, {step})"
        dtype = var.dtype if dtype is None else np.dtype(dtype)
        if not var.is_coord:
            out = _like(chain, name, var, dtype=dtype)
            chain.add_processor(getattr(numpy_bridge, f"{mode}_to_nearest"), var, step, out)
            return out
        if isinstance(step, Real):
            grid = Grid(var.grid.period * step, var.grid.offset)
        elif isinstance(step, (Unit, Quantity)):
            grid = Grid(step, var.grid.offset)
        else:
            grid = step                    # an explicit grid
        out = _like(chain, name, var, dtype=dtype, grid=grid)
        chain._add_conversion(var, grid, out, mode=mode, out_dtype=dtype)
        return out

    return helper


def _registry_unit(u):
    return as_unit(u) if is_in_registry(u) else u


def h_where(chain, condition, a, b, dtype="auto"):
    """``where(cond, a, b)`` / ``a if cond else b``: the result takes grid, coordinate flag and unit from whichever
    operands are variables; constants with units are expressed in the unit (or grid period) of the variable side"""
    if condition is None:
        return None
    Var, Grid = type(chain).Variable, type(chain).Grid
    if not (isinstance(condition, Var) and condition.dtype == "?"):
        raise ProcessingChainError(f"{condition} must be a boolean variable")
    a, b = from_foreign(a), from_foreign(b)
    name = f"where({condition}, {a}, {b})"
    variables = [x for x in (a, b) if isinstance(x, Var)]
    clash = ProcessingChainError(f"{a} and {b} do not have compatible units")
    if len(variables) == 2:
        if a.period != b.period:
            raise ProcessingChainError(f"Cannot select between {a} and {b} with different periods")
        if a.is_coord != b.is_coord:
            raise ProcessingChainError(f"Cannot select between {a} and {b} with different is_coord")
        is_coord = a.is_coord
        # different offsets: the offset itself is selected per event
        grid = a.grid if a.offset == b.offset else Grid(a.period, h_where(chain, condition, a.offset, b.offset))
        ua, ub = _registry_unit(a.unit), _registry_unit(b.unit)
        if ua != ub and ua and ub:
            raise clash
        unit = ua if (ua == ub or not ub) else ub
    elif len(variables) == 1:
        var = variables[0]
        const = b if var is a else a
        grid, is_coord, unit = var.grid, var.is_coord, var.unit or None
        if unit is not None and isinstance(const, Quantity):
            if not is_in_registry(var.unit):
                raise clash
            unit = var.period if is_coord else Quantity(1, as_unit(var.unit))
            const = float(const / (1 * unit))
            a, b = (a, const) if var is a else (const, b)
    else:
        grid, is_coord, unit = None, False, None
        if isinstance(a, Quantity):
            unit = a.u
            if isinstance(b, Quantity):
                b = float(b / Quantity(1.0, unit))
            a = a.m
        elif isinstance(b, Quantity):
            unit, b = b.u, b.m
    out = Var(chain, name, "auto", dtype, grid, unit, is_coord)
    chain._emit(numpy_bridge.where, [condition, a, b, out])
    return out


def h_loadlh5(chain, path_to_file, path_in_file):
    try:
        import lh5
    except ImportError as e:
        raise ProcessingChainError("loadlh5() needs the legend-lh5io package") from e
    try:
        obj = lh5.read(path_in_file, path_to_file)
    except (ValueError, OSError):
        raise ProcessingChainError(f"LH5 file not found: {path_to_file}")
    return obj.value if hasattr(obj, "value") else obj.nda


HELPERS = {
    "len": h_len,
    "isfinite": _predicate(np.isfinite, "isfinite"),
    "isnan": _predicate(np.isnan, "isnan"),
    "round": _rounder("round"),
    "floor": _rounder("floor"),
    "ceil": _rounder("ceil"),
    "trunc": _rounder("trunc"),
    "astype": h_astype,
    "where": h_where,
    "loadlh5": h_loadlh5,
}
