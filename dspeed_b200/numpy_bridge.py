"""Device stand-ins for the *numpy-level glue* a processing chain uses between the
waveform processors: the numpy ufuncs the expression parser emits
(processing_chain.py:46-59 of the reference: add, subtract, multiply, divide,
floor_divide, negative, comparisons, isnan, isfinite), ``numpy.amax`` configured as a
processor (icpc-dsp-config.json:123-129), and the small helper processors of the
reference (``where.py:12-54``, ``get.py:10-91``, ``round_to_nearest.py:11-200``,
``unit_conversion.py:16-78``, ``astype`` processing_chain.py:1269-1300).

These act on per-event scalars (``[block]`` or ``[block, k]`` tensors); they are
evaluated with torch element-wise ops in the loop dtype numpy's type resolution would
pick (IEEE basic operations: results are bit-identical to numpy's).  Inside the fused
chain kernel the same operations are scalar epilogues and launch nothing.
"""

from __future__ import annotations

import numpy as np
import torch

from . import processors as P
from .errors import DSPFatal, ProcessingChainError
from .tables import _np_to_torch

FATAL_CONVERT_INT = 31
FATAL_GET_RANGE = 32


def _t(x, like: torch.Tensor, dtype=None):
    """python / numpy scalar -> python scalar; tensors pass (optionally cast)"""
    if isinstance(x, torch.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).to(like.device, dtype or _np_to_torch(x.dtype))
    if isinstance(x, np.generic):
        return x.item()
    return x


def _record_fatal(fatal, bad_mask: torch.Tensor, code: int):
    if fatal is None:
        if bool(bad_mask.any()):
            e = DSPFatal(P._lib.fatal_message(code))
            e.code = code
            raise e
        return
    flag = bad_mask.any().to(torch.int32) * code
    rec = fatal.reshape(-1)
    rec[0:1].copy_(torch.where(rec[0:1] == 0, flag.reshape(1), rec[0:1]))


class ElementwiseOp:
    """A numpy ufunc evaluated on the device (scalar signature, numpy's type table)."""

    device_processor = True
    launches_per_call = 1

    def __init__(self, ufunc, torch_fn, kind):
        self.ufunc = ufunc
        self.__name__ = ufunc.__name__
        self.torch_fn = torch_fn
        self.kind = kind  # 'arith' | 'compare' | 'predicate'
        self.signature = None
        self.nin, self.nout = ufunc.nin, ufunc.nout
        ok = set("?bBhHiIlLqQfd")
        self.types = [t for t in ufunc.types if set(t.replace("->", "")) <= ok]

    def __call__(self, *args, fatal=None, **kwargs):
        out = args[-1]
        ins = args[:-1]
        if not isinstance(out, torch.Tensor):
            raise ProcessingChainError(f"{self.__name__}: output must be a device tensor")
        if self.kind == "arith":
            cast = out.dtype
        else:
            cast = None
        ins = [_t(a, out, cast) for a in ins]
        if self.kind != "arith":
            # comparison / predicate loops run in the common input type
            tens = [a for a in ins if isinstance(a, torch.Tensor)]
            if len(tens) == 2 and tens[0].dtype != tens[1].dtype:
                common = torch.promote_types(tens[0].dtype, tens[1].dtype)
                ins = [a.to(common) if isinstance(a, torch.Tensor) else a for a in ins]
        if not any(isinstance(a, torch.Tensor) for a in ins):
            ins[0] = torch.as_tensor(ins[0], device=out.device)
        res = self.torch_fn(*ins)
        out.copy_(res if res.shape == out.shape else res.expand_as(out))


def _true_div(a, b):
    return torch.true_divide(a, b)


def _neg(a):
    return torch.neg(a)


_UFUNCS = {
    np.add: ElementwiseOp(np.add, torch.add, "arith"),
    np.subtract: ElementwiseOp(np.subtract, torch.sub, "arith"),
    np.multiply: ElementwiseOp(np.multiply, torch.mul, "arith"),
    np.divide: ElementwiseOp(np.divide, _true_div, "arith"),
    np.floor_divide: ElementwiseOp(np.floor_divide, torch.floor_divide, "arith"),
    np.negative: ElementwiseOp(np.negative, _neg, "arith"),
    np.equal: ElementwiseOp(np.equal, torch.eq, "compare"),
    np.not_equal: ElementwiseOp(np.not_equal, torch.ne, "compare"),
    np.less: ElementwiseOp(np.less, torch.lt, "compare"),
    np.less_equal: ElementwiseOp(np.less_equal, torch.le, "compare"),
    np.greater: ElementwiseOp(np.greater, torch.gt, "compare"),
    np.greater_equal: ElementwiseOp(np.greater_equal, torch.ge, "compare"),
    np.isnan: ElementwiseOp(np.isnan, torch.isnan, "predicate"),
    np.isfinite: ElementwiseOp(np.isfinite, torch.isfinite, "predicate"),
    np.absolute: ElementwiseOp(np.absolute, torch.abs, "arith"),
    np.sqrt: ElementwiseOp(np.sqrt, torch.sqrt, "arith"),
    np.maximum: ElementwiseOp(np.maximum, torch.maximum, "arith"),
    np.minimum: ElementwiseOp(np.minimum, torch.minimum, "arith"),
}


class _Helper:
    """small scalar helper processor with a numba-style type table"""

    device_processor = True
    launches_per_call = 1

    def __init__(self, name, signature, types, fn, nin, nout):
        self.__name__ = name
        self.signature = signature
        self.types = types
        self.fn = fn
        self.nin, self.nout = nin, nout

    def __call__(self, *args, fatal=None, **kwargs):
        return self.fn(*args, fatal=fatal, **kwargs)


_ALL = ["B", "H", "I", "L", "b", "h", "i", "l", "f", "d"]


def _where_impl(cond, a, b, out, fatal=None):
    a = _t(a, out, out.dtype)
    b = _t(b, out, out.dtype)
    if not isinstance(a, torch.Tensor):
        a = torch.as_tensor(a, dtype=out.dtype, device=out.device)
    if not isinstance(b, torch.Tensor):
        b = torch.as_tensor(b, dtype=out.dtype, device=out.device)
    out.copy_(torch.where(cond, a, b))


where = _Helper("where", None, [f"?{t}{t}->{t}" for t in _ALL], _where_impl, 3, 1)


def _gather(a_in, i, out):
    n = a_in.shape[-1]
    if isinstance(i, torch.Tensor):
        idx = i.to(torch.int64)
    else:
        idx = torch.full(out.shape, int(_t(i, out)), dtype=torch.int64, device=out.device)
    idx = idx.reshape(out.shape)
    valid = (idx >= -n) & (idx < n)
    safe = torch.where(idx < 0, idx + n, idx).clamp(0, n - 1)
    a2 = a_in if a_in.shape[0] == out.shape[0] else a_in.expand(out.shape[0], *a_in.shape[1:])
    val = torch.gather(a2, -1, safe.unsqueeze(-1)).squeeze(-1)
    return val, valid


def _get_default_impl(a_in, i, default, out, fatal=None):
    val, valid = _gather(a_in, i, out)
    ok = valid & ~torch.isnan(val) if val.is_floating_point() else valid
    d = _t(default, out, out.dtype)
    if not isinstance(d, torch.Tensor):
        d = torch.as_tensor(d, dtype=out.dtype, device=out.device)
    out.copy_(torch.where(ok, val.to(out.dtype), d))


def _get_impl(a_in, i, out, fatal=None):
    val, valid = _gather(a_in, i, out)
    _record_fatal(fatal, ~valid, FATAL_GET_RANGE)
    out.copy_(val.to(out.dtype))


get_default = _Helper("get_default", "(n),(),()->()", [f"{t}l{t}->{t}" for t in _ALL], _get_default_impl, 3, 1)
get = _Helper("get", "(n),()->()", [f"{t}l->{t}" for t in ["b", "h", "i", "l", "B", "H", "I", "L", "f", "d"]],
              _get_impl, 2, 1)


def _make_rounder(name, tfn):
    def impl(val, to_nearest, out, fatal=None):
        tn = _t(to_nearest, out)
        v = val.to(torch.float64) if not val.is_floating_point() else val
        if isinstance(tn, torch.Tensor) and not tn.is_floating_point():
            tn = tn.to(torch.float64)
        res = tn * tfn(v / tn)
        if val.is_floating_point():
            res = torch.where(torch.isnan(val), val, res)
        out.copy_(res)

    return _Helper(name, None, [f"{t}{t}->{t}" for t in _ALL], impl, 2, 1)


round_to_nearest = _make_rounder("round_to_nearest", torch.round)
floor_to_nearest = _make_rounder("floor_to_nearest", torch.floor)
ceil_to_nearest = _make_rounder("ceil_to_nearest", torch.ceil)
trunc_to_nearest = _make_rounder("trunc_to_nearest", torch.trunc)

HOST_ROUNDERS = {
    "round": lambda v, t=1: t * np.rint(v / t),
    "floor": lambda v, t=1: t * np.floor(v / t),
    "ceil": lambda v, t=1: t * np.ceil(v / t),
    "trunc": lambda v, t=1: t * np.trunc(v / t),
}


def _convert_impl(buf_in, offset_in, offset_out, ratio, out, mode=None, int_check=False, fatal=None):
    """(buf + offset_in) * ratio - offset_out in float64, cast to the output dtype
    (unit_conversion.py:16-78)"""
    oi = offset_in.to(torch.float64) if isinstance(offset_in, torch.Tensor) else float(offset_in)
    oo = offset_out.to(torch.float64) if isinstance(offset_out, torch.Tensor) else float(offset_out)
    tmp = (buf_in.to(torch.float64) + oi) * float(ratio) - oo
    if mode == "round":
        tmp = torch.round(tmp)
    elif mode == "floor":
        tmp = torch.floor(tmp)
    elif mode == "ceil":
        tmp = torch.ceil(tmp)
    elif mode == "trunc":
        tmp = torch.trunc(tmp)
    elif int_check:
        r = torch.round(tmp)
        _record_fatal(fatal, ~((tmp - r).abs() < 1.0e-5), FATAL_CONVERT_INT)
        tmp = r
    out.copy_(tmp)


convert = _Helper("convert", None, ["fddd->f", "dddd->d"], _convert_impl, 4, 1)


def make_astype(in_dtype, out_dtype):
    """``astype`` processor (unsafe cast, like ``np.copyto(casting='unsafe')``)"""

    def impl(a_in, a_out, fatal=None):
        a_out.copy_(a_in)

    return _Helper("astype", "()->()", [f"{np.dtype(in_dtype).char}->{np.dtype(out_dtype).char}"], impl, 1, 1)


def device_equivalent(func, signature=None):
    """The device implementation behind a callable named in a recipe, or a set-up error:
    dspeed_b200 never executes a processor on the host."""
    if getattr(func, "device_processor", False):
        return func
    if isinstance(func, np.ufunc) and func in _UFUNCS:
        return _UFUNCS[func]
    name = getattr(func, "__name__", str(func))
    if func in (np.amax, np.max):
        return P.amax
    if name in P._REGISTRY:  # e.g. the reference's own numba function object was passed in
        return P._REGISTRY[name]
    helpers = {
        "where": where, "get": get, "get_default": get_default, "round_to_nearest": round_to_nearest,
        "floor_to_nearest": floor_to_nearest, "ceil_to_nearest": ceil_to_nearest,
        "trunc_to_nearest": trunc_to_nearest,
    }
    if name in helpers:
        return helpers[name]
    raise ProcessingChainError(
        f"processor '{name}' has no B200 implementation (not on the hot path, see DESIGN.md); "
        "dspeed_b200 has no CPU fallback")
