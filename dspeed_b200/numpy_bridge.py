"""Device stand-ins for the *numpy-level glue* a processing chain uses between the
waveform processors: the numpy ufuncs the expression parser emits
(processing_chain.py:46-59 of the reference: add, subtract, multiply, divide,
floor_divide, negative, comparisons, isnan, isfinite), ``numpy.amax`` configured as a
processor (icpc-dsp-config.json:123-129), and the small helper processors of the
reference (``where.py:12-54``, ``get.py:10-91``, ``round_to_nearest.py:11-200``,
``unit_conversion.py:16-78``, ``astype`` processing_chain.py:1269-1300).

These act on per-event scalars and waveforms (``[block]`` or ``[block, k]`` tensors, constants, immediates); they are
evaluated by ONE hand-written element-wise kernel family (``csrc/glue.cu``: ``dspb_glue`` / ``dspb_glue_get``) in the
loop type numpy's type resolution would pick (IEEE basic operations: results are bit-identical to numpy's).  Inside
the fused chain kernels the same operations are scalar epilogues and launch nothing.
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


# ---- the hand-written glue kernels (csrc/glue.cu) -------------------------------------------------------------------
_GT = {torch.float32: 0, torch.float64: 1, torch.uint16: 2, torch.int16: 3, torch.int32: 4, torch.uint32: 5, torch.int64: 6,
       torch.bool: 7, torch.int8: 8, torch.uint8: 9, torch.uint64: 10}
_GOP = {"add": 0, "subtract": 1, "multiply": 2, "divide": 3, "true_divide": 3, "floor_divide": 4, "negative": 5, "absolute": 6,
        "sqrt": 7, "maximum": 8, "minimum": 9, "equal": 16, "not_equal": 17, "less": 18, "less_equal": 19, "greater": 20,
        "greater_equal": 21, "isnan": 22, "isfinite": 23, "where": 32, "copy": 33, "round": 40, "floor": 41, "ceil": 42,
        "trunc": 43, "convert": 48}
_NP_OF = {torch.float32: np.float32, torch.float64: np.float64, torch.uint16: np.uint16, torch.int16: np.int16,
          torch.int32: np.int32, torch.uint32: np.uint32, torch.int64: np.int64, torch.bool: np.bool_, torch.int8: np.int8,
          torch.uint8: np.uint8, torch.uint64: np.uint64}
_vp, _i64, _i32, _f64 = P._vp, P._i64, P._i32, P._f64


def _loop_of(dtype) -> int:
    """loop type code of the glue kernel: 0 float32, 1 float64, 2 int64 (every integer loop)"""
    return 0 if dtype == torch.float32 else (1 if dtype == torch.float64 else 2)


def _common_loop(ins) -> int:
    """loop of a comparison / predicate: numpy's result type of the array operands (python scalars are weak; a python
    float against integer arrays makes the loop float64)"""
    dts = [_NP_OF[a.dtype] for a in ins if isinstance(a, torch.Tensor)]
    res = np.result_type(*dts) if dts else np.dtype(np.float64)
    if res.kind not in "f" and any(isinstance(a, float) for a in ins):
        res = np.dtype(np.float64)
    return 0 if res == np.float32 else (1 if res == np.float64 else 2)


def _opnd(x, out: torch.Tensor, keep: list):
    """(ptr, row stride, column stride, dtype code, immediate) of one operand against `out` ([rows] or [rows, inner])"""
    if x is None:
        return [_vp(None), _i64(0), _i64(0), _i32(1), _f64(0.0)]
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x)).to(out.device)
    if isinstance(x, np.generic):
        x = x.item()
    if not isinstance(x, torch.Tensor):
        return [_vp(None), _i64(0), _i64(0), _i32(1), _f64(float(x))]
    t = x
    if t.device != out.device:
        t = t.to(out.device)
    if t.dtype not in _GT:
        raise ProcessingChainError(f"glue operand dtype {t.dtype} is not supported on the device")
    rows = out.shape[0]
    inner = out.shape[1] if out.ndim > 1 else 1
    if t.ndim == 0:
        t = t.reshape(1)
    if out.ndim > 1 and t.ndim == 1:
        t = t.unsqueeze(1)          # a per-event scalar against a waveform
    if t.ndim > 2:
        t = t.reshape(t.shape[0], -1)
    if t.shape[0] not in (1, rows) or (t.ndim > 1 and t.shape[1] not in (1, inner)):
        raise ProcessingChainError(f"glue operand of shape {tuple(x.shape)} does not broadcast to {tuple(out.shape)}")
    rs = t.stride(0) if (t.shape[0] == rows and rows > 1) else 0
    cs = t.stride(1) if (t.ndim > 1 and t.shape[1] == inner and inner > 1) else 0
    keep.append(t)
    return [_vp(t.data_ptr()), _i64(rs), _i64(cs), _i32(_GT[t.dtype]), _f64(0.0)]


def _glue(op: str, loop: int, out: torch.Tensor, a, b=None, c=None, ratio=1.0, mode=0):
    """one launch of the element-wise glue kernel (csrc/glue.cu: dspb_glue)"""
    if not (isinstance(out, torch.Tensor) and out.is_cuda):
        raise ProcessingChainError(f"{op}: output must be a CUDA tensor (no CPU fallback)")
    o = out if out.ndim <= 2 else out.reshape(out.shape[0], -1)
    if o.ndim == 2 and o.shape[1] > 1 and o.stride(1) != 1:
        raise ProcessingChainError(f"{op}: output must be contiguous along its last axis")
    if out.dtype not in _GT:
        raise ProcessingChainError(f"{op}: output dtype {out.dtype} is not supported on the device")
    keep = []
    rows = o.shape[0]
    inner = o.shape[1] if o.ndim > 1 else 1
    with torch.cuda.device(out.device):
        rc = P._lib.lib().dspb_glue(_i32(_GOP[op]), _i32(loop), _i64(rows), _i64(inner), _vp(o.data_ptr()),
                                    _i64(o.stride(0) if rows > 1 else inner), _i32(_GT[out.dtype]),
                                    *_opnd(a, o, keep), *_opnd(b, o, keep), *_opnd(c, o, keep), _f64(float(ratio)), _i32(mode),
                                    P._stream_ptr(out.device))
    if rc:
        raise RuntimeError(f"dspb_glue({op}) failed with {rc}")


class ElementwiseOp:
    """A numpy ufunc evaluated on the device by the hand-written glue kernel (scalar signature, numpy's type table).
    Arithmetic runs in the output's type (the loop numpy's type resolution picked for the bound variables),
    comparisons / predicates in the common type of their inputs."""

    device_processor = True
    launches_per_call = 1

    def __init__(self, ufunc, kind):
        self.ufunc = ufunc
        self.__name__ = ufunc.__name__
        self.kind = kind  # 'arith' | 'compare' | 'predicate'
        self.signature = None
        self.nin, self.nout = ufunc.nin, ufunc.nout
        ok = set("?bBhHiIlLqQfd")
        self.types = [t for t in ufunc.types if set(t.replace("->", "")) <= ok]

    def __call__(self, *args, fatal=None, **kwargs):
        out = args[-1]
        ins = [a.item() if isinstance(a, np.generic) else a for a in args[:-1]]
        if not isinstance(out, torch.Tensor):
            raise ProcessingChainError(f"{self.__name__}: output must be a device tensor")
        loop = _loop_of(out.dtype) if self.kind == "arith" else _common_loop(ins)
        _glue(self.__name__, loop, out, ins[0], ins[1] if len(ins) > 1 else None)


_UFUNCS = {
    np.add: ElementwiseOp(np.add, "arith"),
    np.subtract: ElementwiseOp(np.subtract, "arith"),
    np.multiply: ElementwiseOp(np.multiply, "arith"),
    np.divide: ElementwiseOp(np.divide, "arith"),
    np.floor_divide: ElementwiseOp(np.floor_divide, "arith"),
    np.negative: ElementwiseOp(np.negative, "arith"),
    np.equal: ElementwiseOp(np.equal, "compare"),
    np.not_equal: ElementwiseOp(np.not_equal, "compare"),
    np.less: ElementwiseOp(np.less, "compare"),
    np.less_equal: ElementwiseOp(np.less_equal, "compare"),
    np.greater: ElementwiseOp(np.greater, "compare"),
    np.greater_equal: ElementwiseOp(np.greater_equal, "compare"),
    np.isnan: ElementwiseOp(np.isnan, "predicate"),
    np.isfinite: ElementwiseOp(np.isfinite, "predicate"),
    np.absolute: ElementwiseOp(np.absolute, "arith"),
    np.sqrt: ElementwiseOp(np.sqrt, "arith"),
    np.maximum: ElementwiseOp(np.maximum, "arith"),
    np.minimum: ElementwiseOp(np.minimum, "arith"),
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
    """where.py:12-54"""
    _glue("where", _loop_of(out.dtype), out, cond, a, b)


where = _Helper("where", None, [f"?{t}{t}->{t}" for t in _ALL], _where_impl, 3, 1)


def _get_call(a_in, i, default, out, use_default, fatal):
    """get.py:10-91 through dspb_glue_get: out[r] = a[r, i[r]], negative indices wrap"""
    if not (isinstance(a_in, torch.Tensor) and a_in.is_cuda and out.is_cuda):
        raise ProcessingChainError("get: operands must be CUDA tensors (no CPU fallback)")
    a2 = a_in if a_in.ndim == 2 else a_in.reshape(a_in.shape[0], -1)
    if a2.stride(1) != 1:
        a2 = a2.contiguous()
    rows = out.reshape(-1).shape[0]
    keep = [a2]
    o1 = out.reshape(rows, 1) if out.ndim == 1 else out
    own = fatal is None
    if own:
        fatal = torch.zeros(4, dtype=torch.int32, device=out.device)
    iop = _opnd(i, o1, keep)
    dop = _opnd(default, o1, keep)
    with torch.cuda.device(out.device):
        rc = P._lib.lib().dspb_glue_get(_i32(_loop_of(out.dtype)), _i64(rows), _vp(a2.data_ptr()),
                                        _i64(a2.stride(0) if a2.shape[0] == rows and rows > 1 else 0), _i32(_GT[a2.dtype]),
                                        _i64(a2.shape[1]), iop[0], iop[1], iop[3], iop[4], dop[0], dop[1], dop[3], dop[4],
                                        _i32(1 if use_default else 0), _vp(out.data_ptr()), _i32(_GT[out.dtype]),
                                        _vp(fatal.data_ptr()), P._stream_ptr(out.device))
    if rc:
        raise RuntimeError(f"dspb_glue_get failed with {rc}")
    if own and not use_default:
        P.raise_if_fatal(fatal, "get")


def _get_default_impl(a_in, i, default, out, fatal=None):
    _get_call(a_in, i, default, out, True, fatal)


def _get_impl(a_in, i, out, fatal=None):
    _get_call(a_in, i, None, out, False, fatal)


get_default = _Helper("get_default", "(n),(),()->()", [f"{t}l{t}->{t}" for t in _ALL], _get_default_impl, 3, 1)
get = _Helper("get", "(n),()->()", [f"{t}l->{t}" for t in ["b", "h", "i", "l", "B", "H", "I", "L", "f", "d"]],
              _get_impl, 2, 1)


def _make_rounder(name, op):
    def impl(val, to_nearest, out, fatal=None):
        """round_to_nearest.py:11-200: to_nearest * f(val / to_nearest), NaN stays NaN; integer loops go through
        float64 like numpy's true division"""
        loop = _loop_of(out.dtype)
        _glue(op, 1 if loop == 2 else loop, out, val, to_nearest)

    return _Helper(name, None, [f"{t}{t}->{t}" for t in _ALL], impl, 2, 1)


round_to_nearest = _make_rounder("round_to_nearest", "round")
floor_to_nearest = _make_rounder("floor_to_nearest", "floor")
ceil_to_nearest = _make_rounder("ceil_to_nearest", "ceil")
trunc_to_nearest = _make_rounder("trunc_to_nearest", "trunc")

HOST_ROUNDERS = {
    "round": lambda v, t=1: t * np.rint(v / t),
    "floor": lambda v, t=1: t * np.floor(v / t),
    "ceil": lambda v, t=1: t * np.ceil(v / t),
    "trunc": lambda v, t=1: t * np.trunc(v / t),
}


def _convert_impl(buf_in, offset_in, offset_out, ratio, out, mode=None, int_check=False, fatal=None):
    """(buf + offset_in) * ratio - offset_out in float64, cast to the output dtype (unit_conversion.py:16-78)"""
    if int_check and mode is None:
        # integer coordinate that must convert to whole samples: round, then check on the device result
        _glue("convert", 1, out, buf_in, offset_in, offset_out, ratio=ratio, mode=1)
        exact = torch.empty(out.shape, dtype=torch.float64, device=out.device)
        _glue("convert", 1, exact, buf_in, offset_in, offset_out, ratio=ratio, mode=0)
        _record_fatal(fatal, ~((exact - torch.round(exact)).abs() < 1.0e-5), FATAL_CONVERT_INT)
        return
    _glue("convert", 1, out, buf_in, offset_in, offset_out, ratio=ratio,
          mode={None: 0, "round": 1, "floor": 2, "ceil": 3, "trunc": 4}[mode])


convert = _Helper("convert", None, ["fddd->f", "dddd->d"], _convert_impl, 4, 1)


def make_astype(in_dtype, out_dtype):
    """``astype`` processor (unsafe cast, like ``np.copyto(casting='unsafe')``)"""

    def impl(a_in, a_out, fatal=None):
        _glue("copy", _loop_of(a_in.dtype) if isinstance(a_in, torch.Tensor) else 1, a_out, a_in)

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
