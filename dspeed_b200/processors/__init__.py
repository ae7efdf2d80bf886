"""``dspeed_b200.processors`` -- the hot-path processors of ``dspeed.processors`` as
hand-written sm_100a CUDA kernels behind the reference's gufunc protocol.

Every processor keeps the reference's name, argument order and meaning
(src/dspeed/processors/__init__.py:66-159) and exposes what the chain's
``ProcessorManager`` needs (processing_chain.py:1528-1543): ``__name__``,
``.signature``, ``.types``, and a void call ``proc(*inputs, *outputs)`` that writes
caller-owned output arrays in place.

Arrays are ``torch`` CUDA tensors (the chain's block buffers).  NumPy arrays are also
accepted, exactly like the reference's gufuncs: they are staged to the device, the same
kernels run, and the outputs are copied back into the caller's arrays -- there is NO
CPU implementation behind these functions; without the CUDA library they raise.

Data-dependent ``DSPFatal`` conditions are recorded on the device.  Called stand-alone
a processor synchronises and raises immediately (reference behaviour); the processing
chain instead passes ``fatal=<int32[4] device tensor>`` and checks once per block.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _lib
from ..errors import DSPFatal

__all__: list[str] = []

_DT_CODE = {
    torch.float32: 0, torch.float64: 1, torch.uint16: 2, torch.int16: 3, torch.int32: 4, torch.uint32: 5,
}
_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f64 = C.c_double


class _Call:
    """Marshals one processor call into the C ABI of include/dspeed_b200.h."""

    def __init__(self, T: torch.dtype, device: torch.device):
        self.T = T
        self.device = device
        self.n_rows = 1
        self.keep = []  # keep temporaries alive until the launch is enqueued

    # -- row count -----------------------------------------------------------------
    def rows_from(self, *arrs):
        n = 1
        for a in arrs:
            if isinstance(a, torch.Tensor) and a.ndim >= 1:
                lead = a.shape[0]
                if lead != 1:
                    if n != 1 and lead != n:
                        raise ValueError(f"cannot broadcast leading dimensions {n} and {lead}")
                    n = lead
        self.n_rows = n
        return n

    # -- operand marshalling ---------------------------------------------------------
    def wave_in(self, w: torch.Tensor):
        """-> (ptr, row_stride, dtype code), n"""
        if w.ndim == 1:
            w = w.unsqueeze(0)
        if w.ndim != 2:
            w = w.reshape(-1, w.shape[-1])
        if w.dtype not in _DT_CODE:
            w = w.to(self.T)
        if w.shape[-1] > 1 and w.stride(-1) != 1:
            w = w.contiguous()
        self.keep.append(w)
        rs = 0 if w.shape[0] == 1 and self.n_rows > 1 else w.stride(0)
        return [_vp(w.data_ptr()), _i64(rs), _i32(_DT_CODE[w.dtype])], w.shape[-1]

    def wave_out(self, w: torch.Tensor, n: int | None = None):
        if w.ndim == 1:
            w = w.unsqueeze(0)
        if w.dtype != self.T:
            raise TypeError(f"output dtype {w.dtype} does not match the {self.T} type loop")
        if w.shape[-1] > 1 and w.stride(-1) != 1:
            raise ValueError("output waveforms must be contiguous along the sample axis")
        if n is not None and w.shape[-1] != n:
            raise DSPFatal(f"Output waveform has length {w.shape[-1]}; expect {n}")
        return [_vp(w.data_ptr()), _i64(w.stride(0))], w.shape[-1]

    def scalar_in(self, x):
        """per-row scalar: python/numpy number -> immediate; tensor -> device array"""
        if isinstance(x, torch.Tensor):
            t = x.reshape(-1)
            if t.dtype != self.T:
                t = t.to(self.T)
            if t.numel() == 1:
                stride = 0
            else:
                if t.numel() != self.n_rows:
                    raise ValueError(f"scalar argument has {t.numel()} entries for {self.n_rows} rows")
                stride = t.stride(0)
            self.keep.append(t)
            return [_vp(t.data_ptr()), _i64(stride), _f64(0.0)]
        return [_vp(None), _i64(0), _f64(float(x))]

    def scalar_out(self, x: torch.Tensor):
        t = x.reshape(-1) if x.ndim != 1 else x
        if t.dtype != self.T:
            raise TypeError(f"output dtype {t.dtype} does not match the {self.T} type loop")
        if t.numel() != self.n_rows or (t.numel() > 1 and t.stride(0) != 1):
            raise ValueError("scalar outputs must be contiguous [n_rows] arrays")
        return _vp(t.data_ptr())


def _as_int(x) -> int:
    if isinstance(x, torch.Tensor):
        x = x.reshape(-1)[0].item()
    if isinstance(x, np.ndarray):
        x = x.reshape(-1)[0]
    if isinstance(x, (str, bytes)):
        return ord(x)
    return int(x)


def _as_float(x) -> float:
    if isinstance(x, torch.Tensor):
        x = x.reshape(-1)[0].item()
    if isinstance(x, np.ndarray):
        x = x.reshape(-1)[0]
    return float(x)


def _stream_ptr(device) -> _vp:
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def raise_if_fatal(fatal: torch.Tensor, processor: str | None = None, wf_range=None) -> None:
    """Synchronise on the device fatal record and raise the reference's DSPFatal."""
    rec = fatal.cpu()
    code = int(rec[0])
    if code:
        e = DSPFatal(_lib.fatal_message(code))
        e.code = code
        e.row = int(rec[1]) | (int(rec[2]) << 31)
        e.processor = processor
        e.wf_range = wf_range
        raise e


class DeviceProcessor:
    """A CUDA processor obeying the gufunc protocol of the reference."""

    def __init__(self, name, signature, types, impl, nout, doc="", out_args=None):
        self.__name__ = name
        self.out_args = out_args      # positions of the output arguments when they are not the last `nout` ones
        self.__doc__ = doc
        self.signature = signature
        self.types = list(types)
        self.impl = impl
        self.nout = nout
        self.nin = signature.count("(") - nout if signature else None
        self.device_processor = True
        self.native_kernel = True  # a hand-written CUDA kernel of this repository
        self.launches_per_call = 1

    def __repr__(self):
        return f"<dspeed_b200 processor {self.__name__} {self.signature}>"

    def __call__(self, *args, fatal: torch.Tensor | None = None, **kwargs):
        host = [isinstance(a, np.ndarray) and a.ndim > 0 for a in args]
        any_dev = any(isinstance(a, torch.Tensor) for a in args)
        if any(host) or not any_dev:
            return self._call_host(args, kwargs)
        dev = next(a.device for a in args if isinstance(a, torch.Tensor))
        if dev.type != "cuda":
            raise RuntimeError(f"{self.__name__}: tensors must live on a CUDA device (no CPU fallback)")
        own_fatal = fatal is None
        if own_fatal:
            fatal = torch.zeros(4, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = self.impl(*args, fatal=fatal, **kwargs)
        if rc:
            if rc < 0:
                raise RuntimeError(f"{self.__name__}: CUDA error {-rc}")
            e = DSPFatal(_lib.fatal_message(rc))
            e.code = rc
            raise e
        if own_fatal:
            raise_if_fatal(fatal, self.__name__)

    def _call_host(self, args, kwargs):
        """NumPy front door: stage to the device, run the same kernels, copy back."""
        if not torch.cuda.is_available():
            raise RuntimeError(
                f"{self.__name__}: no CUDA device available and dspeed_b200 has no CPU fallback"
            )
        dev = torch.device("cuda", torch.cuda.current_device())
        nargs = len(args)
        dargs = []
        outs = []
        for i, a in enumerate(args):
            is_out = (i in self.out_args) if self.out_args is not None else i >= nargs - self.nout
            if isinstance(a, np.ndarray) and a.ndim > 0:
                if is_out:
                    if not a.flags.writeable:
                        raise ValueError("output array is read-only")
                    t = torch.empty(a.shape, dtype=_torch_dtype(a.dtype), device=dev)
                    outs.append((a, t))
                else:
                    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                dargs.append(t)
            elif isinstance(a, np.ndarray):
                dargs.append(a.item())
            else:
                dargs.append(a)
        self.__call__(*dargs, **kwargs)
        for a, t in outs:
            np.copyto(a, t.cpu().numpy())


def _torch_dtype(dt) -> torch.dtype:
    return {
        np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
        np.dtype("uint16"): torch.uint16, np.dtype("int16"): torch.int16,
        np.dtype("int32"): torch.int32, np.dtype("uint32"): torch.uint32,
        np.dtype("int64"): torch.int64, np.dtype("uint8"): torch.uint8, np.dtype("int8"): torch.int8,
        np.dtype("bool"): torch.bool,
    }[np.dtype(dt)]


def _sfx(T):
    return "_f32" if T == torch.float32 else "_f64"


def _out_T(*outs):
    for o in outs:
        if isinstance(o, torch.Tensor) and o.dtype in (torch.float32, torch.float64):
            return o.dtype
    raise TypeError("outputs must be float32 or float64 device tensors")


def _fn(name, T):
    return getattr(_lib.lib(), name + _sfx(T))


def _tail(fatal, dev):
    return [_vp(fatal.data_ptr() if fatal is not None else None), _stream_ptr(dev)]


_REGISTRY: dict[str, DeviceProcessor] = {}


def _register(name, signature, types, nout, out_args=None):
    def deco(f):
        p = DeviceProcessor(name, signature, types, f, nout, f.__doc__ or "", out_args)
        _REGISTRY[name] = p
        globals()[name] = p
        __all__.append(name)
        return p

    return deco


_FD = ["f->f", "d->d"]  # placeholder, replaced per processor


# ---------------------------------------------------------------------------------------
# processors
# ---------------------------------------------------------------------------------------
@_register("bl_subtract", "(n),()->(n)", ["ff->f", "dd->d"], 1)
def _bl_subtract(w_in, a_baseline, w_out, fatal=None):
    """bl_subtract.py:11-46"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, a_baseline, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_bl_subtract", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(a_baseline), *wo, *_tail(fatal, w_out.device))


@_register("min_max", "(n)->(),(),(),()", ["f->ffff", "d->dddd"], 4)
def _min_max(w_in, t_min, t_max, a_min, a_max, fatal=None):
    """min_max.py:11-82"""
    T = _out_T(a_min)
    c = _Call(T, a_min.device)
    c.rows_from(w_in, a_min)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_min_max", T)(*wi, _i64(c.n_rows), _i64(n), c.scalar_out(t_min), c.scalar_out(t_max),
                                  c.scalar_out(a_min), c.scalar_out(a_max), *_tail(fatal, a_min.device))


@_register("amax", "(n),()->()", ["fi->f", "di->d"], 1)
def _amax(w_in, axis, a_max, fatal=None):
    """numpy.amax(w, axis, out) as configured in icpc-dsp-config.json:123-129"""
    T = _out_T(a_max)
    c = _Call(T, a_max.device)
    c.rows_from(w_in, a_max)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_amax", T)(*wi, _i64(c.n_rows), _i64(n), c.scalar_out(a_max), *_tail(fatal, a_max.device))


@_register("min_max_norm", "(n),(),()->(n)", ["fff->f", "ddd->d"], 1)
def _min_max_norm(w_in, a_min, a_max, w_out, fatal=None):
    """min_max.py:85-140"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, a_min, a_max, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_min_max_norm", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(a_min), *c.scalar_in(a_max), *wo,
                                       *_tail(fatal, w_out.device))


@_register("linear_slope_fit", "(n)->(),(),(),()", ["f->ffff", "d->dddd"], 4)
def _linear_slope_fit(w_in, mean, stdev, slope, intercept, fatal=None):
    """linear_slope_fit.py:11-90 (the reference's mean/stdev baseline processor)"""
    T = _out_T(mean)
    c = _Call(T, mean.device)
    c.rows_from(w_in, mean)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_linear_slope_fit", T)(*wi, _i64(c.n_rows), _i64(n), c.scalar_out(mean), c.scalar_out(stdev),
                                           c.scalar_out(slope), c.scalar_out(intercept), *_tail(fatal, mean.device))


@_register("mean_stdev", "(n)->(),()", ["f->ff", "d->dd"], 2)
def _mean_stdev(w_in, mean, stdev, fatal=None):
    """Two-output alias of linear_slope_fit (BASELINE.json names the baseline processor
    ``mean_stdev``; the reference snapshot only has linear_slope_fit.py:11-90)."""
    slope = torch.empty_like(mean)
    icpt = torch.empty_like(mean)
    return _linear_slope_fit.impl(w_in, mean, stdev, slope, icpt, fatal=fatal)


@_register("linear_slope_diff", "(n),(),()->(),()", ["fff->ff", "ddd->dd"], 2)
def _linear_slope_diff(w_in, slope, intercept, mean, rms, fatal=None):
    """linear_slope_fit.py:93-158"""
    T = _out_T(mean)
    c = _Call(T, mean.device)
    c.rows_from(w_in, slope, intercept, mean)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_linear_slope_diff", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(slope), *c.scalar_in(intercept),
                                            c.scalar_out(mean), c.scalar_out(rms), *_tail(fatal, mean.device))


@_register("mean_below_threshold", "(n),()->()", ["ff->f", "dd->d"], 1)
def _mean_below_threshold(w_in, threshold, result, fatal=None):
    """arithmetic.py:9-62"""
    T = _out_T(result)
    c = _Call(T, result.device)
    c.rows_from(w_in, threshold, result)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_mean_below_threshold", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(threshold),
                                               c.scalar_out(result), *_tail(fatal, result.device))


@_register("pole_zero", "(n),()->(n)", ["ff->f", "dd->d"], 1)
def _pole_zero(w_in, t_tau, w_out, fatal=None):
    """pole_zero.py:24-77"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, t_tau, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_pole_zero", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(t_tau), *wo, *_tail(fatal, w_out.device))


@_register("double_pole_zero", "(n),(),(),()->(n)", ["ffff->f", "dddd->d"], 1)
def _double_pole_zero(w_in, t_tau1, t_tau2, frac, w_out, fatal=None):
    """pole_zero.py:82-198"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, t_tau1, t_tau2, frac, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_double_pole_zero", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(t_tau1), *c.scalar_in(t_tau2),
                                           *c.scalar_in(frac), *wo, *_tail(fatal, w_out.device))


def _trap_common(w_in, rise, flat, w_out, norm, fatal):
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_trap_filter", T)(*wi, _i64(c.n_rows), _i64(n), _i32(_as_int(rise)), _i32(_as_int(flat)), _i32(norm),
                                      *wo, *_tail(fatal, w_out.device))


@_register("trap_filter", "(n),(),()->(n)", ["fii->f", "dii->d"], 1)
def _trap_filter(w_in, rise, flat, w_out, fatal=None):
    """trap_filters.py:12-76"""
    return _trap_common(w_in, rise, flat, w_out, 0, fatal)


@_register("trap_norm", "(n),(),()->(n)", ["fii->f", "dii->d"], 1)
def _trap_norm(w_in, rise, flat, w_out, fatal=None):
    """trap_filters.py:79-149"""
    return _trap_common(w_in, rise, flat, w_out, 1, fatal)


@_register("asym_trap_filter", "(n),(),(),()->(n)", ["fiii->f", "diii->d"], 1)
def _asym_trap_filter(w_in, rise, flat, fall, w_out, fatal=None):
    """trap_filters.py:152-227"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_asym_trap_filter", T)(*wi, _i64(c.n_rows), _i64(n), _i32(_as_int(rise)), _i32(_as_int(flat)),
                                           _i32(_as_int(fall)), *wo, *_tail(fatal, w_out.device))


@_register("trap_pickoff", "(n),(),(),()->()", ["fiif->f", "diid->d"], 1)
def _trap_pickoff(w_in, rise, flat, t_pickoff, a_out, fatal=None):
    """trap_filters.py:230-301"""
    T = _out_T(a_out)
    c = _Call(T, a_out.device)
    c.rows_from(w_in, t_pickoff, a_out)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_trap_pickoff", T)(*wi, _i64(c.n_rows), _i64(n), _i32(_as_int(rise)), _i32(_as_int(flat)),
                                       *c.scalar_in(t_pickoff), c.scalar_out(a_out), *_tail(fatal, a_out.device))


def _mw_common(w_in, length, w_out, kind, fatal):
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_moving_window", T)(*wi, _i64(c.n_rows), _i64(n), _f64(_as_float(length)), _i32(kind), *wo,
                                        *_tail(fatal, w_out.device))


@_register("moving_window_left", "(n),()->(n)", ["ff->f", "dd->d"], 1)
def _moving_window_left(w_in, length, w_out, fatal=None):
    """moving_windows.py:12-61"""
    return _mw_common(w_in, length, w_out, 0, fatal)


@_register("moving_window_right", "(n),()->(n)", ["ff->f", "dd->d"], 1)
def _moving_window_right(w_in, length, w_out, fatal=None):
    """moving_windows.py:64-114"""
    return _mw_common(w_in, length, w_out, 1, fatal)


@_register("moving_window_multi", "(n),(),(),()->(n)", ["fffi->f", "dddi->d"], 1)
def _moving_window_multi(w_in, length, num_mw, mw_type, w_out, fatal=None):
    """moving_windows.py:117-203"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_moving_window_multi", T)(*wi, _i64(c.n_rows), _i64(n), _f64(_as_float(length)),
                                              _f64(_as_float(num_mw)), _i32(_as_int(mw_type)), *wo,
                                              *_tail(fatal, w_out.device))


@_register("avg_current", "(n),(),(m)", ["fff->", "ddd->"], 1)
def _avg_current(w_in, length, w_out, fatal=None):
    """moving_windows.py:206-249"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, m = c.wave_out(w_out)
    return _fn("dspb_avg_current", T)(*wi, _i64(c.n_rows), _i64(n), _f64(_as_float(length)), *wo, _i64(m),
                                      *_tail(fatal, w_out.device))


@_register("time_point_thresh", "(n),(),(),()->()", ["ffff->f", "dddd->d"], 1)
def _time_point_thresh(w_in, a_threshold, t_start, walk_forward, t_out, fatal=None):
    """time_point_thresh.py:12-92"""
    T = _out_T(t_out)
    c = _Call(T, t_out.device)
    c.rows_from(w_in, a_threshold, t_start, walk_forward, t_out)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_time_point_thresh", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(a_threshold), *c.scalar_in(t_start),
                                            *c.scalar_in(walk_forward), c.scalar_out(t_out), *_tail(fatal, t_out.device))


@_register("interpolated_time_point_thresh", "(n),(),(),(),()->()", ["ffflb->f", "dddlb->d"], 1)
def _interpolated_time_point_thresh(w_in, a_threshold, t_start, walk_forward, mode_in, t_out, fatal=None):
    """time_point_thresh.py:95-222"""
    T = _out_T(t_out)
    c = _Call(T, t_out.device)
    c.rows_from(w_in, a_threshold, t_start, t_out)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_interpolated_time_point_thresh", T)(
        *wi, _i64(c.n_rows), _i64(n), *c.scalar_in(a_threshold), *c.scalar_in(t_start), _i64(_as_int(walk_forward)),
        _i32(_as_int(mode_in)), c.scalar_out(t_out), *_tail(fatal, t_out.device))


@_register("multi_time_point_thresh", "(n),(m),(),(),()->(m)", ["ffffb->f", "ddddb->d"], 1)
def _multi_time_point_thresh(w_in, a_threshold, t_start, polarity, mode_in, t_out, fatal=None):
    """time_point_thresh.py:225-401"""
    T = _out_T(t_out)
    c = _Call(T, t_out.device)
    c.rows_from(w_in, t_start, t_out)
    wi, n = c.wave_in(w_in)
    thr = a_threshold.to(T)
    if thr.ndim == 1:
        thr = thr.unsqueeze(0)
    thr = thr.contiguous()
    c.keep.append(thr)
    m = thr.shape[-1]
    rs = m if thr.shape[0] == c.n_rows and c.n_rows > 1 else (m if thr.shape[0] == c.n_rows else 0)
    if thr.shape[0] == 1:
        rs = 0
    to = t_out if t_out.ndim == 2 else t_out.unsqueeze(0)
    assert to.is_contiguous() and to.shape[-1] == m
    return _fn("dspb_multi_time_point_thresh", T)(
        *wi, _i64(c.n_rows), _i64(n), _vp(thr.data_ptr()), _i64(m), _i64(rs), *c.scalar_in(t_start),
        _f64(_as_float(polarity)), _i32(_as_int(mode_in)), _vp(to.data_ptr()), *_tail(fatal, t_out.device))


@_register("fixed_time_pickoff", "(n),(),()->()", ["ffb->f", "ddb->d"], 1)
def _fixed_time_pickoff(w_in, t_in, mode_in, a_out, fatal=None):
    """fixed_time_pickoff.py:12-125"""
    T = _out_T(a_out)
    c = _Call(T, a_out.device)
    c.rows_from(w_in, t_in, a_out)
    wi, n = c.wave_in(w_in)
    return _fn("dspb_fixed_time_pickoff", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(t_in), _i32(_as_int(mode_in)),
                                             c.scalar_out(a_out), *_tail(fatal, a_out.device))


@_register("windower", "(n),(),(m)", ["fff->", "ddd->"], 1)
def _windower(w_in, t0_in, w_out, fatal=None):
    """windower.py:12-54"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, t0_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, m = c.wave_out(w_out)
    return _fn("dspb_windower", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(t0_in), *wo, _i64(m),
                                   *_tail(fatal, w_out.device))


@_register("upsampler", "(n),(),(m)", ["fff->", "ddd->"], 1)
def _upsampler(w_in, upsample, w_out, fatal=None):
    """upsampler.py:14-49"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, m = c.wave_out(w_out)
    return _fn("dspb_upsampler", T)(*wi, _i64(c.n_rows), _i64(n), _f64(_as_float(upsample)), *wo, _i64(m),
                                    *_tail(fatal, w_out.device))


#: float32 convolutions of whole blocks with a generic kernel of at least this many taps run on the
#: tensor cores (csrc/conv_tc.cu: banded Toeplitz GEMM, 3xTF32, tcgen05 + TMEM + TMA); 0 disables the path
TC_CONV_MIN_TAPS = int(os.environ.get("DSPEED_B200_TC_CONV_MIN_TAPS", "128"))
_tc_workspace: dict = {}


def _convolve_tc(w_in, k, mode, w_out) -> bool:
    """tensor-core path of `convolve_wf`; False when the call does not fit it"""
    if (TC_CONV_MIN_TAPS <= 0 or k.numel() < TC_CONV_MIN_TAPS or w_in.dtype != torch.float32 or w_out.dtype != torch.float32
            or w_in.ndim != 2 or w_out.ndim != 2 or w_in.stride(1) != 1 or w_out.stride(1) != 1
            or w_in.stride(0) % 4 or w_in.data_ptr() % 16 or w_in.shape[0] != w_out.shape[0]
            or mode not in "fvs" or k.numel() > w_in.shape[1]):   # (never on the row count: results must not depend on block_width)
        return False
    L = _lib.lib()
    L.dspb_convolve_tc_workspace.restype = C.c_int64
    need = int(L.dspb_convolve_tc_workspace(_i64(k.numel())))
    key = (w_in.device, need, torch.cuda.current_stream(w_in.device).cuda_stream)   # the tiles are rebuilt per call, on the call's stream
    ws = _tc_workspace.get(key)
    if ws is None:
        ws = _tc_workspace[key] = torch.empty(need, dtype=torch.float32, device=w_in.device)
    rc = L.dspb_convolve_tc_f32(_vp(w_in.data_ptr()), _i64(w_in.stride(0)), _i64(w_in.shape[0]), _i64(w_in.shape[1]),
                                _vp(k.data_ptr()), _i64(k.numel()), _i32(ord(mode)), _vp(w_out.data_ptr()),
                                _i64(w_out.stride(0)), _i64(w_out.shape[1]), _vp(ws.data_ptr()), _i64(need),
                                _vp(torch.cuda.current_stream(w_in.device).cuda_stream))
    if rc < 0:
        raise RuntimeError(f"dspb_convolve_tc_f32: CUDA error {-rc}")
    return rc == 0   # (NaN rows -> NaN outputs, convolutions.py:44-46, are handled by the launcher's second kernel)


def _convolve_common(w_in, kernel, mode_in, w_out, fatal):
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, p = c.wave_out(w_out)
    k = kernel.reshape(-1).to(T).contiguous()
    c.keep.append(k)
    if isinstance(w_in, torch.Tensor) and _convolve_tc(w_in, k, chr(_as_int(mode_in)), w_out):
        return 0
    return _fn("dspb_convolve_wf", T)(*wi, _i64(c.n_rows), _i64(n), _vp(k.data_ptr()), _i64(k.numel()),
                                      _i32(_as_int(mode_in)), *wo, _i64(p), *_tail(fatal, w_out.device))


@_register("convolve_wf", "(n),(m),(),(p)", ["ffbf->", "ddbd->"], 1)
def _convolve_wf(w_in, kernel, mode_in, w_out, fatal=None):
    """convolutions.py:14-72"""
    return _convolve_common(w_in, kernel, mode_in, w_out, fatal)


@_register("fft_convolve_wf", "(n),(m),(),(p)", ["ffbf", "ddbd"], 1)
def _fft_convolve_wf(w_in, kernel, mode_in, w_out, fatal=None):
    """convolutions.py:75-119 -- same sums as convolve_wf, evaluated directly on the
    device (the reference's in-place zeroing of NaN input rows is not reproduced; NaN
    rows give NaN outputs)."""
    return _convolve_common(w_in, kernel, mode_in, w_out, fatal)


@_register("get_multi_local_extrema", "(n),(),(),(),(),(),(m),(m),(),()", ["ffffffffII->", "ddddddddII->"], 4)
def _get_multi_local_extrema(w_in, a_delta_max_in, a_delta_min_in, search_direction, a_abs_max_in, a_abs_min_in,
                             vt_max_out, vt_min_out, n_max_out, n_min_out, fatal=None):
    """get_multi_local_extrema.py:12-306"""
    T = _out_T(vt_max_out)
    c = _Call(T, vt_max_out.device)
    c.rows_from(w_in, vt_max_out)
    wi, n = c.wave_in(w_in)
    vmx = vt_max_out if vt_max_out.ndim == 2 else vt_max_out.unsqueeze(0)
    vmn = vt_min_out if vt_min_out.ndim == 2 else vt_min_out.unsqueeze(0)
    assert vmx.is_contiguous() and vmn.is_contiguous() and vmx.shape == vmn.shape
    assert n_max_out.dtype == torch.uint32 and n_min_out.dtype == torch.uint32
    m = vmx.shape[-1]
    return _fn("dspb_get_multi_local_extrema", T)(
        *wi, _i64(c.n_rows), _i64(n), _f64(_as_float(a_delta_max_in)), _f64(_as_float(a_delta_min_in)),
        _f64(_as_float(search_direction)), *c.scalar_in(a_abs_max_in), *c.scalar_in(a_abs_min_in),
        _vp(vmx.data_ptr()), _vp(vmn.data_ptr()), _i64(m), _vp(n_max_out.data_ptr()), _vp(n_min_out.data_ptr()),
        *_tail(fatal, vt_max_out.device))


@_register("recursive_filter", "(n),(p),(q),(),()->(n)", ["fddff->f", "ddddd->d"], 1)
def _recursive_filter(w_in, a, b, init_in, init_out, w_out, fatal=None):
    """recursive_filter.py:12-93 (len(b) <= 3)"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, init_in, init_out, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    av = np.ascontiguousarray(np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a, np.float64).reshape(-1))
    bv = np.ascontiguousarray(np.asarray(b.cpu() if isinstance(b, torch.Tensor) else b, np.float64).reshape(-1))
    if np.isnan(av).any() or np.isnan(bv).any():
        w_out.fill_(float("nan"))
        return 0
    if bv.size > 3 or av.size > 8:      # beyond the order-2 affine scan: the sequential float64 recursion (csrc/sipm.cu)
        return _fn("dspb_recursive_filter_general", T)(
            *wi, _i64(c.n_rows), _i64(n), av.ctypes.data_as(_vp), _i64(av.size), bv.ctypes.data_as(_vp), _i64(bv.size),
            *c.scalar_in(init_in), *c.scalar_in(init_out), *wo, *_tail(fatal, w_out.device))
    return _fn("dspb_recursive_filter", T)(
        *wi, _i64(c.n_rows), _i64(n), av.ctypes.data_as(_vp), _i64(av.size), bv.ctypes.data_as(_vp), _i64(bv.size),
        *c.scalar_in(init_in), *c.scalar_in(init_out), *wo, *_tail(fatal, w_out.device))


# ---- the IIR family built on recursive_filter (pole_zero.py:201-342, rc_cr2.py, iir_filter.py) ------------------------
def _rc_exp(tau) -> float:
    """pole_zero.py:13-20"""
    tau = _as_float(tau)
    return float(np.exp(-1.0 / tau)) if tau != 0 else 0.0


def _first_sample(w_in, T):
    """w_in[..., 0] as a per-row scalar of the loop's type"""
    w = w_in if w_in.ndim == 2 else w_in.unsqueeze(0)
    return w[:, 0].to(T).contiguous()


@_register("convolve_exp", "(n),()->(n)", ["ff->f", "dd->d"], 1)
def _convolve_exp(w_in, tau, w_out, fatal=None):
    """pole_zero.py:201-232: recursive_filter with a = [1], b = [1, -exp(-1 / tau)], both histories = w_in[0]"""
    T = _out_T(w_out)
    x0 = _first_sample(w_in, T)
    tau_t = np.float32(_as_float(tau)) if T == torch.float32 else _as_float(tau)
    return _recursive_filter.impl(w_in, np.ones(1), np.array([1.0, -_rc_exp(tau_t)]), x0, x0, w_out, fatal=fatal)


@_register("convolve_damped_oscillator", "(n),(),(),()->(n)", ["fddd->f", "dddd->d"], 1)
def _convolve_damped_oscillator(w_in, tau, omega, phase, w_out, fatal=None):
    """pole_zero.py:235-281"""
    T = _out_T(w_out)
    x0 = _first_sample(w_in, T)
    rc, om, ph = _rc_exp(tau), _as_float(omega), _as_float(phase)
    a = np.array([np.cos(ph), -rc * np.cos(om - ph)])
    b = np.array([1.0, -2 * rc * np.cos(om), rc * rc])
    return _recursive_filter.impl(w_in, a, b, x0, x0, w_out, fatal=fatal)


@_register("inject_damped_oscillation", "(n),(),(),(),()->(n)", ["fdddd->f", "ddddd->d"], 1)
def _inject_damped_oscillation(w_in, tau, omega, phase, frac, w_out, fatal=None):
    """pole_zero.py:284-342"""
    T = _out_T(w_out)
    fr = _as_float(frac)
    if not 0 <= fr <= 1:
        return 37      # DSPB_FATAL_INJ_FRAC ("frac must be between zero and one.")
    x0 = _first_sample(w_in, T)
    rc, om, ph = _rc_exp(tau), _as_float(omega), _as_float(phase)
    cw, cp, cwp = np.cos(om), np.cos(ph), np.cos(om - ph)
    a = np.array([1 + fr * cp, -(2 * rc * cw + fr * cp + fr * rc * cwp), rc * (rc + fr * cwp)])
    b = np.array([1.0, -2 * rc * cw, rc * rc])
    return _recursive_filter.impl(w_in, a, b, x0, 0.0, w_out, fatal=fatal)


@_register("rc_cr2", "(n),()->(n)", ["ff->f", "dd->d"], 1)
def _rc_cr2(w_in, t_tau, w_out, fatal=None):
    """rc_cr2.py:11-93"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, t_tau, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    return _fn("dspb_rc_cr2", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(t_tau), *wo, *_tail(fatal, w_out.device))


def _iir_processor(a, b, name, init_out):
    """iir_filter.py:103-113, 161-170, 219-228: a (n)->(n) processor running recursive_filter with designed coefficients;
    input history w_in[0], output history `init_out` * w_in[0]"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)

    def impl(w_in, w_out, fatal=None):
        T = _out_T(w_out)
        x0 = _first_sample(w_in, T)
        io = 0.0 if init_out == 0 else (x0 if init_out == 1 else (x0.to(torch.float64) * init_out).to(T))
        return _recursive_filter.impl(w_in, a, b, x0, io, w_out, fatal=fatal)

    return DeviceProcessor(name, "(n)->(n)", ["ff->f", "dd->d"], impl, 1, "recursive_filter with scipy.signal-designed coefficients")


def _f_samp(f_samp):
    """a waveform variable stands for its sampling frequency (iir_filter.py:71-72)"""
    if hasattr(f_samp, "period") and hasattr(f_samp, "proc_chain"):
        return 1 / f_samp.period
    return f_samp


def _norm_freq(freq, f_samp, two):
    def one(f):
        return float(2 * f / f_samp) if f_samp is not None else float(f)

    if two:
        if not (hasattr(freq, "__len__") and len(freq) == 2):
            raise DSPFatal("bandpass / bandstop filter requires two freq values")
        fc = [one(f) for f in freq]
        ok = all(0 <= f <= 1 for f in fc)
    else:
        fc = one(freq)
        ok = 0 <= fc <= 1
    if not ok:
        raise DSPFatal("Critical frequency must be positive and < nyquist frequency")
    return fc


def iir_filter(freq, order, rp=None, rs=None, f_samp=None, ftype="butter", btype="lowpass"):
    """iir_filter.py:18-113 -- factory (``init_args`` of a recipe): designs the filter with scipy.signal.iirfilter on the
    host at set-up time (the reference does the same) and returns the device processor that applies it."""
    import scipy.signal as sg

    f_samp = _f_samp(f_samp)
    if btype in ("lowpass", "highpass"):
        fc = _norm_freq(freq, f_samp, False)
    elif btype in ("bandpass", "bandstop"):
        fc = _norm_freq(freq, f_samp, True)
    else:
        raise DSPFatal("Invalid type of filter")
    a, b = sg.iirfilter(order, fc, rp=rp, rs=rs, btype=btype, ftype=ftype)
    return _iir_processor(a, b, f"{ftype}({freq}, {order}, {btype})", float(np.sum(a) / np.sum(b)))


def notch_filter(freq, bandwidth, f_samp=None):
    """iir_filter.py:115-170"""
    import scipy.signal as sg

    a, b = sg.iirnotch(_norm_freq(freq, _f_samp(f_samp), False), float(freq / bandwidth))
    return _iir_processor(a, b, f"notch({freq}, {bandwidth})", 1)


def peak_filter(freq, bandwidth, f_samp=None):
    """iir_filter.py:173-228"""
    import scipy.signal as sg

    a, b = sg.iirpeak(_norm_freq(freq, _f_samp(f_samp), False), float(freq / bandwidth))
    return _iir_processor(a, b, f"peak({freq}, {bandwidth})", 0)


__all__ += ["iir_filter", "notch_filter", "peak_filter"]


# ---- set-up time kernel generators (const-folded by the chain compiler) ------------------
def _kernel_out(kernel):
    T = _out_T(kernel)
    k = kernel.reshape(-1)
    assert k.is_contiguous()
    return T, k


@_register("cusp_filter", "(),(),(),(n)", ["ffff->", "dddd->"], 1)
def _cusp_filter(sigma, flat, decay, kernel, fatal=None):
    """energy_kernels.py:12-73"""
    T, k = _kernel_out(kernel)
    r = (lambda v: float(np.float32(_as_float(v)))) if T == torch.float32 else _as_float
    return _fn("dspb_cusp_filter", T)(_f64(r(sigma)), _f64(r(flat)), _f64(r(decay)), _vp(k.data_ptr()), _i64(k.numel()),
                                      _stream_ptr(kernel.device))


@_register("zac_filter", "(),(),(),(n)", ["ffff->", "dddd->"], 1)
def _zac_filter(sigma, flat, decay, kernel, fatal=None):
    """energy_kernels.py:76-157"""
    T, k = _kernel_out(kernel)
    r = (lambda v: float(np.float32(_as_float(v)))) if T == torch.float32 else _as_float
    return _fn("dspb_zac_filter", T)(_f64(r(sigma)), _f64(r(flat)), _f64(r(decay)), _vp(k.data_ptr()), _i64(k.numel()),
                                     _stream_ptr(kernel.device))


@_register("t0_filter", "(),(),(n)", ["fff->", "ddd->"], 1)
def _t0_filter(rise, fall, kernel, fatal=None):
    """kernels.py:12-61"""
    T, k = _kernel_out(kernel)
    r = (lambda v: float(np.float32(_as_float(v)))) if T == torch.float32 else _as_float
    return _fn("dspb_t0_filter", T)(_f64(r(rise)), _f64(r(fall)), _vp(k.data_ptr()), _i64(k.numel()),
                                    _stream_ptr(kernel.device))


# ---- the rest of the SiPM / LAr chain (tests/configs/sipm-dsp-config.json; csrc/sipm.cu) -----------------------------
def _list2d(c, a, T, what):
    """[n_rows, m] list-valued operand (histogram weights / borders, index lists) -> (tensor, ptr, row stride, m)"""
    t = a if a.ndim == 2 else a.unsqueeze(0)
    if t.dtype != T:
        raise TypeError(f"{what} dtype {t.dtype} does not match the {T} type loop")
    if t.shape[-1] > 1 and t.stride(-1) != 1:
        t = t.contiguous()
    c.keep.append(t)
    rs = 0 if t.shape[0] == 1 and c.n_rows > 1 else t.stride(0)
    return t, _vp(t.data_ptr()), _i64(rs), t.shape[-1]


def _rows_of_wave(c, w_in, *more):
    """row count of a call whose list-valued operands have their own core dimensions: a 1-D waveform is ONE row"""
    if isinstance(w_in, torch.Tensor) and w_in.ndim == 1:
        c.n_rows = 1
        return 1
    return c.rows_from(w_in, *[m for m in more if isinstance(m, torch.Tensor) and m.ndim >= 2])


@_register("histogram", "(n),(m),(p)", ["fff->", "ddd->"], 2)
def _histogram(w_in, weights_out, borders_out, fatal=None):
    """histogram.py:14-89"""
    T = _out_T(weights_out)
    c = _Call(T, weights_out.device)
    _rows_of_wave(c, w_in, weights_out)
    wi, n = c.wave_in(w_in)
    _, wp, wrs, m = _list2d(c, weights_out, T, "weights_out")
    _, bp, brs, p = _list2d(c, borders_out, T, "borders_out")
    return _fn("dspb_histogram", T)(*wi, _i64(c.n_rows), _i64(n), wp, wrs, _i64(m), bp, brs, _i64(p),
                                    *_tail(fatal, weights_out.device))


@_register("histogram_around_mode", "(n),(),(),(m),(p)", ["fffff->", "ddddd->"], 2)
def _histogram_around_mode(w_in, center, bin_width, weights_out, borders_out, fatal=None):
    """histogram.py:92-204"""
    T = _out_T(weights_out)
    c = _Call(T, weights_out.device)
    _rows_of_wave(c, w_in, weights_out)
    wi, n = c.wave_in(w_in)
    _, wp, wrs, m = _list2d(c, weights_out, T, "weights_out")
    _, bp, brs, p = _list2d(c, borders_out, T, "borders_out")
    return _fn("dspb_histogram_around_mode", T)(*wi, _i64(c.n_rows), _i64(n), *c.scalar_in(center), *c.scalar_in(bin_width),
                                                wp, wrs, _i64(m), bp, brs, _i64(p), *_tail(fatal, weights_out.device))


@_register("histogram_stats", "(n),(m),(),(),(),()", ["ffffff->", "dddddd->"], 3, out_args=(2, 3, 4))
def _histogram_stats(weights_in, edges_in, mode_out, max_out, fwhm_out, max_in, fatal=None):
    """histogram_stats.py:146-261 (outputs are arguments 3-5, `max_in` is the last INPUT, like the reference)"""
    T = _out_T(max_out)
    c = _Call(T, max_out.device)
    _rows_of_wave(c, weights_in)
    _, wp, wrs, m = _list2d(c, weights_in.to(T) if weights_in.dtype != T else weights_in, T, "weights_in")
    _, ep, ers, p = _list2d(c, edges_in.to(T) if edges_in.dtype != T else edges_in, T, "edges_in")
    return _fn("dspb_histogram_stats", T)(wp, wrs, _i64(m), ep, ers, _i64(p), _i64(c.n_rows), c.scalar_out(mode_out),
                                          c.scalar_out(max_out), c.scalar_out(fwhm_out), *c.scalar_in(max_in),
                                          *_tail(fatal, max_out.device))


@_register("histogram_peakstats", "(n),(m),(),(),()->(),()", ["fffii->ff", "dddii->dd"], 2)
def _histogram_peakstats(weights_in, edges_in, max_in, skip_zeroes, width_type, mode_out, width_out, fatal=None):
    """histogram_stats.py:12-143"""
    T = _out_T(mode_out)
    c = _Call(T, mode_out.device)
    _rows_of_wave(c, weights_in)
    _, wp, wrs, m = _list2d(c, weights_in.to(T) if weights_in.dtype != T else weights_in, T, "weights_in")
    _, ep, ers, p = _list2d(c, edges_in.to(T) if edges_in.dtype != T else edges_in, T, "edges_in")
    return _fn("dspb_histogram_peakstats", T)(wp, wrs, _i64(m), ep, ers, _i64(p), _i64(c.n_rows), *c.scalar_in(max_in),
                                              _i32(_as_int(skip_zeroes)), _i32(_as_int(width_type)), c.scalar_out(mode_out),
                                              c.scalar_out(width_out), *_tail(fatal, mode_out.device))


@_register("peak_snr_threshold", "(n),(m),(),()->(m),()", ["ffff->fI", "dddd->dI"], 2)
def _peak_snr_threshold(w_in, idx_in, ratio_in, width_in, idx_out, n_idx_out, fatal=None):
    """peak_snr_threshold.py:11-71"""
    T = _out_T(idx_out)
    c = _Call(T, idx_out.device)
    _rows_of_wave(c, w_in, idx_out)
    wi, n = c.wave_in(w_in)
    _, ip, irs, m = _list2d(c, idx_in, T, "idx_in")
    _, op, ors, mo = _list2d(c, idx_out, T, "idx_out")
    assert mo == m and n_idx_out.dtype == torch.uint32
    return _fn("dspb_peak_snr_threshold", T)(*wi, _i64(c.n_rows), _i64(n), ip, irs, _i64(m), *c.scalar_in(ratio_in),
                                             *c.scalar_in(width_in), op, ors, _vp(n_idx_out.data_ptr()),
                                             *_tail(fatal, idx_out.device))


@_register("multi_a_filter", "(n),(m)->(m)", ["ff->f", "dd->d"], 1)
def _multi_a_filter(w_in, vt_maxs_in, va_max_out, fatal=None):
    """multi_a_filter.py:11-57"""
    T = _out_T(va_max_out)
    c = _Call(T, va_max_out.device)
    _rows_of_wave(c, w_in, va_max_out)
    wi, n = c.wave_in(w_in)
    _, ip, irs, m = _list2d(c, vt_maxs_in, T, "vt_maxs_in")
    _, op, ors, mo = _list2d(c, va_max_out, T, "va_max_out")
    assert mo == m
    return _fn("dspb_multi_a_filter", T)(*wi, _i64(c.n_rows), _i64(n), ip, irs, _i64(m), op, ors,
                                         *_tail(fatal, va_max_out.device))


@_register("reflected_convolve_wf", "(n),(m),(p)", ["fff->", "ddd->"], 1)
def _reflected_convolve_wf(w_in, kernel, w_out, fatal=None):
    """convolutions.py:122-182"""
    T = _out_T(w_out)
    c = _Call(T, w_out.device)
    c.rows_from(w_in, w_out)
    wi, n = c.wave_in(w_in)
    wo, _ = c.wave_out(w_out, n)
    k = kernel.reshape(-1).to(T).contiguous()
    c.keep.append(k)
    return _fn("dspb_reflected_convolve_wf", T)(*wi, _i64(c.n_rows), _i64(n), _vp(k.data_ptr()), _i64(k.numel()), *wo,
                                                *_tail(fatal, w_out.device))


@_register("gaussian_filter1d", "(),(),(n)", ["fff->", "ddd->"], 1)
def _gaussian_filter1d(sigma, truncate, weights, fatal=None):
    """gaussian_filter1d.py:46-82"""
    T, k = _kernel_out(weights)
    return _fn("dspb_gaussian_filter1d", T)(_f64(_as_float(sigma)), _f64(_as_float(truncate)), _vp(k.data_ptr()),
                                            _i64(k.numel()), _stream_ptr(weights.device))


@_register("moving_slope", "(n)", ["f->", "d->"], 1)
def _moving_slope(kernel, fatal=None):
    """kernels.py:64-103"""
    T, k = _kernel_out(kernel)
    return _fn("dspb_moving_slope", T)(_vp(k.data_ptr()), _i64(k.numel()), _stream_ptr(kernel.device))


@_register("step", "(),(n)", ["ff->", "dd->"], 1)
def _step(weight_pos, kernel, fatal=None):
    """kernels.py:106-142"""
    T, k = _kernel_out(kernel)
    return _fn("dspb_step", T)(_f64(_as_float(weight_pos)), _vp(k.data_ptr()), _i64(k.numel()), _stream_ptr(kernel.device))


@_register("dplms", "(n,n),(m),(),(),(),()->(n)", ["ffffff->f", "dddddd->d"], 1)
def _dplms(noise_mat, reference, a1, a2, a3, ff, kernel, fatal=None):
    """energy_kernels.py:160-272 -- the DPLMS optimum filter (set-up time, const-folded by the chain compiler).  The
    penalised normal equations (a1 N + a2 R + a3 1 1^T) x = r are assembled and solved in float64 ON THE DEVICE
    (torch.linalg.solve: cuSOLVER LU, a library call at set-up time like the reference's numpy.linalg.solve); the
    kernel is the flipped solution, normalised by the maximum of its 'valid' convolution with the reference pulse."""
    T, k = _kernel_out(kernel)
    dev = kernel.device
    length = k.numel()

    def dev64(x):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        return t.to(device=dev, dtype=T).to(torch.float64)     # values as the loop's type holds them

    nm = dev64(noise_mat)
    if nm.ndim == 3:
        nm = nm[0]
    ref = dev64(reference).reshape(-1)
    a1, a2, a3, ff = (float(np.dtype("f4" if T == torch.float32 else "f8").type(_as_float(v))) for v in (a1, a2, a3, ff))

    def bad(msg):
        e = DSPFatal(msg)
        e.code = 0
        raise e

    if nm.shape[0] != length:
        bad("The length of the filter is not consistent with the noise matrix")
    if ref.numel() <= 0:
        bad("The length of the reference signal must be positive")
    if a1 <= 0:
        bad("The penalized coefficient for the noise must be positive")
    if a2 <= 0:
        bad("The penalized coefficient for the reference must be positive")
    if a3 <= 0:
        bad("The penalized coefficient for the zero area must be positive")
    if ff <= 0:
        bad("The penalized coefficient for the ref matrix must be positive")
    if ff != 1:
        bad("The penalized coefficient for the ref matrix must be 0 or 1")
    ssize = ref.numel()
    flo, fhi = int(ssize / 2 - length / 2), int(ssize / 2 + length / 2)
    ref_mat = torch.zeros((length, length), dtype=torch.float64, device=dev)
    ref_sig = torch.zeros(length, dtype=torch.float64, device=dev)
    shifts = (-1, 0, 1)
    for i in shifts:
        seg = ref[flo + i: fhi + i]
        ref_mat += torch.outer(seg, seg)
        ref_sig += seg
    ref_mat /= len(shifts)
    mat = a1 * nm + a2 * ref_mat + a3
    sol = torch.linalg.solve(mat, ref_sig)
    k.copy_(torch.flip(sol, (0,)).to(T))                 # kernel[:] = np.flip(solve(...)) (stored in the loop's type)
    # y = np.convolve(reference, kernel, 'valid'); kernel /= max(y)
    kk = k.to(torch.float64)
    y = torch.nn.functional.conv1d(ref.view(1, 1, -1), torch.flip(kk, (0,)).view(1, 1, -1)).view(-1)
    k.copy_((k.to(torch.float64) / y.max()).to(T) if T == torch.float64 else (k / y.max().to(T)))
    return 0


def __getattr__(name):
    raise AttributeError(
        f"module {__name__} has no attribute {name}: not on the B200 hot path "
        "(see DESIGN.md 'out of scope'); there is no CPU fallback"
    )
