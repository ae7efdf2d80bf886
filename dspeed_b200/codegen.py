"""Chain code generator: turns a compiled :class:`ProcessingChain` into ONE specialised,
straight-line CUDA kernel (sm_100a), compiles it with nvcc and launches it per block.

The reference runs ~60 gufunc calls per 16-row block and writes every intermediate
waveform to memory (processing_chain.py:1144-1163).  Here the frozen launch descriptors
(``ProcessorManager.args``, reference :1493-1792) are lowered to a kernel in which

* every thread owns a 16-sample chunk of the waveform and keeps it in registers from
  one processor to the next (load -> min_max -> bl_subtract -> linear_slope_fit ->
  pole_zero ... never leave the register file);
* recursive filters (pole-zero, trapezoids, moving windows, run-structured convolution
  kernels) are chunk-local running sums plus one block scan of the chunk totals;
* block collectives are split into a *put* before and a *get* after a barrier, and all
  independent collectives of consecutive processors share ONE barrier ("round");
* per-event scalars are registers, scalar glue (``np.multiply``, unit conversion,
  ``round``) is inline arithmetic;
* parameters (tap positions, window lengths, slices) are literals, so all index
  arithmetic folds at compile time.

The device routines are hand-written (``csrc/chain_rt.cuh``, ``csrc/row_ops.cuh``); this
module only sequences them.  Chains containing a processor without a specialised
emitter stay on the interpreted program of :mod:`dspeed_b200.fusion`.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import logging
import math
import os
import re
import subprocess
import threading

import numpy as np
import torch

from . import _lib, fusion, numpy_bridge
from . import processors as P
from .fusion import _DT, FusedChain, NotFusable, _storage, _storage_id

log = logging.getLogger("dspeed")

CHK = 16
NT = 512
# block scans of filter running sums entirely in float32 (chain_rt.cuh: put_scan_ff / get_excl_ff); "0": float64 across warps
_FF = "ff" if os.environ.get("DSPEED_B200_F32_SCANS", "1") != "0" else "f"
MAX_SMEM = 227 * 1024
CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_chains")
_build_lock = threading.Lock()
_loaded: dict[str, C.CDLL] = {}

_CTYPE = {torch.float32: "float", torch.float64: "double", torch.uint16: "uint16_t", torch.int16: "int16_t",
          torch.int32: "int32_t", torch.uint32: "uint32_t", torch.int64: "long long"}


class NotSpecializable(NotFusable):
    pass


def _lit(v) -> str:
    """C++ literal of a double (hex float: exact)"""
    v = float(v)
    if math.isnan(v):
        return "CUDART_NAN"
    if math.isinf(v):
        return "CUDART_INF" if v > 0 else "(-CUDART_INF)"
    return v.hex()


def _flit(v) -> str:
    """C++ literal of a float32 value"""
    v = float(np.float32(v))
    if math.isnan(v):
        return "CUDART_NAN_F"
    if math.isinf(v):
        return "CUDART_INF_F" if v > 0 else "(-CUDART_INF_F)"
    return v.hex() + "f"


class Wave:
    def __init__(self, wid, n):
        self.id = wid
        self.n = n               # root length
        self.slot = None
        self.reg = None          # name of the float[16] register chunk (own chunk), when live
        self.nan = "0"           # uniform int expression: != 0 -> the whole wave is NaN
        self.uses = []           # node indices reading it
        self.producer = None
        self.is_input = False
        self.needs_slot = True
        self.scatter = False     # written at other threads' positions (needs a barrier before any read)
        self.name = f"W{wid}"


def host_kernel(origin, length):
    """float32 kernel array of a const-folded generator (planning on the meta device only;
    on a CUDA device the array computed by the device generator is used)"""
    name, args = origin
    f = np.float32
    if name == "t0_filter":
        rise, fall = (float(f(P._as_float(x))) for x in args[:2])
        k = np.zeros(length, np.float64)
        for i in range(int(rise)):
            k[i] = 2 * (int(rise) - i) / (rise * (rise + 1))
        k[int(rise):] = -1 / fall
        return k.astype(f)
    return None


class SpecChain(FusedChain):
    kernel_name = "k_chain_spec (specialised waveform-resident chain kernel)"

    # ------------------------------------------------------------------------------------
    # analysis
    # ------------------------------------------------------------------------------------
    def _compile(self, chain, slot_limit=None):
        entry_state = dict(self.__dict__)        # a second planning pass starts from exactly this state
        self._slot_limit = slot_limit
        managers = list(chain._proc_managers)
        self.n_managers = len(managers)
        self.meta = chain.device.type == "meta"
        fusion._ALIAS.clear()
        self.cse_skipped = 0
        self.ptrs = []
        self.ptr_index = {}
        self.waves: dict[int, Wave] = {}
        self.svar: dict[int, str] = {}          # scalar storage -> C variable
        self.stype: dict[str, str] = {}         # C variable -> "float" | "double"
        self.never_nan: set[str] = set()        # scalars of integer input columns
        self.const_storage = {}
        self.input_wave = {}
        self.input_scalar = {}
        self.nodes = []
        self.conv_lowering = []
        self.prolog = []                         # input scalar loads
        self._sources = None

        all_vars = list(chain._vars_dict.values())
        for pm in managers:
            for prm in list(pm.params) + list(pm.kw_params.values()):
                if hasattr(prm, "proc_chain") and prm not in all_vars:
                    all_vars.append(prm)
        self.var_of_storage = {}
        for v in all_vars:
            bufs = v.all_buffers()
            for b, _ in bufs:
                if isinstance(b, torch.Tensor):
                    self.var_of_storage.setdefault(_storage(b), v)
                    if v.is_const:
                        self.const_storage[_storage(b)] = b
        for name, man in chain._input_managers.items():
            self._register_input(man)

        # identical wave producers share one result (wf_etrap == wf_trap with the default db)
        seen = {}
        self.skip = set()
        for i, pm in enumerate(managers):
            if not getattr(pm.processor, "native_kernel", False) or getattr(pm.processor, "nout", 1) != 1:
                continue
            out = pm.args[-1]
            if not (isinstance(out, torch.Tensor) and out.ndim == 2 and out.storage_offset() == 0):
                continue
            key = [getattr(pm.processor, "__name__", "")]
            for x in pm.args[:-1]:
                if isinstance(x, torch.Tensor):
                    key.append(("t", _storage(x), x.storage_offset(), tuple(x.shape), tuple(x.stride())))
                else:
                    key.append(("c", repr(x)))
            key = tuple(key)
            if key in seen and tuple(seen[key].shape) == tuple(out.shape):
                fusion._ALIAS[_storage_id(out)] = _storage(seen[key])
                self.skip.add(i)
            else:
                seen[key] = out

        self.out_wave_storages = {}
        self.out_scalars = []
        for name, man in chain._output_managers.items():
            rv = self._out_raw(man)
            if rv.ndim >= 2:
                self.out_wave_storages[_storage(rv)] = rv
        # ---- nodes -------------------------------------------------------------------------
        for i, pm in enumerate(managers):
            self._cur = i
            self._fatal_idx = pm.fatal.storage_offset() // 4
            if i in self.skip:
                self.cse_skipped += 1
                continue
            self._lower(pm)
        # outputs
        for name, man in chain._output_managers.items():
            rv = self._out_raw(man)
            st = _storage(rv)
            if rv.ndim == 1:
                if st in self.const_storage:
                    continue
                if st in self.input_scalar and st not in self.svar:
                    self._sc(rv)
                if st not in self.svar:
                    raise NotSpecializable(f"output {name} is not produced by a specialisable processor")
                if rv.dtype not in (torch.float32, torch.float64, torch.int32, torch.uint32):
                    raise NotSpecializable(f"output dtype {rv.dtype}")
                self.out_scalars.append((self.svar[st], self._ptr(("buf", rv)), _CTYPE[rv.dtype]))
            else:
                if st not in self.waves or self.waves[st].is_input:
                    raise NotSpecializable("waveform pass-through outputs stay on the copy path")
                self._node("store_wave", ins=[(self.waves[st], rv.storage_offset() % max(1, rv.stride(0)), rv.shape[1])],
                           ptr=self._ptr(("buf", rv)))
        if len(self.ptrs) > 64:
            raise NotSpecializable("too many distinct device pointers")
        self.used_scalars = {name for (name, _, _) in self.out_scalars}
        for nd in self.nodes:
            for key in ("x", "y", "thr", "start", "walk", "t", "t0", "b", "tau", "oi", "oo"):
                if key in nd and isinstance(nd[key], str):
                    self.used_scalars.add(nd[key])
        self.out_of = {}
        for k, (name, pi, ct) in enumerate(self.out_scalars):
            self.out_of.setdefault(name, []).append((pi, ct, k))
        self._schedule()
        self._emit_kernel()
        # Small chains need few shared-memory slots: a second pass with the pool cut to what the first one used makes
        # room for TWO resident CTAs per SM (two waveforms in flight hide each other's barrier and shuffle latencies)
        # (pool = 2 x the logical slots: the two row frames are then disjoint, the strongest form of the rotation rule)
        need = max(2, 2 * self.logical_slots_used)
        if (slot_limit is None and need < self.total_slots
                and os.environ.get("DSPEED_B200_OCCUPANCY", "2") != "1"
                and self.fixed_bytes + need * self.slot_words * 4 <= (MAX_SMEM - 2048) // 2):
            self.__dict__.clear()
            self.__dict__.update(entry_state)
            return self._compile(chain, slot_limit=need)
        self._build()

    # -- operands ----------------------------------------------------------------------------
    def _node(self, kind, **kw):
        nd = dict(kind=kind, idx=len(self.nodes), fatal=self._fatal_idx if hasattr(self, "_fatal_idx") else 0, **kw)
        for w, _, _ in nd.get("ins", []):
            w.uses.append(nd["idx"])
        for w in nd.get("wouts", []):
            w.producer = nd["idx"]
        self.nodes.append(nd)
        return nd

    def _const_value(self, t: torch.Tensor):
        var = self.var_of_storage.get(_storage(t))
        hv = getattr(var, "host_value", None)
        if hv is not None:
            return np.asarray(hv).reshape(-1)
        if self.meta:
            origin = getattr(var, "const_origin", None)
            if origin is not None:
                k = host_kernel(origin, int(t.numel()))
                if k is not None:
                    return k
            raise NotSpecializable("constant without a host value on the meta device")
        return self.const_storage[_storage(t)].reshape(-1).detach().cpu().numpy()

    def _sc(self, x):
        """C expression (double) of a per-event scalar operand, and the names it depends on"""
        if isinstance(x, torch.Tensor):
            st = _storage(x)
            if st in self.const_storage:
                v = self._const_value(x)
                if v.size != 1:
                    raise NotSpecializable("non-scalar constant used as a scalar")
                return _lit(v[0])
            if x.numel() != x.shape[0]:
                raise NotSpecializable("vector-valued per-event variable")
            if st not in self.svar:
                if st not in self.input_scalar:
                    raise NotSpecializable("scalar operand read before it is produced")
                man, what = self.input_scalar[st]
                src = man.t0_var if what == "t0" else man.raw_var
                if src.dtype not in _CTYPE:
                    raise NotSpecializable(f"scalar input dtype {src.dtype}")
                name = self._new_svar(st, src.dtype)
                if src.dtype not in (torch.float32, torch.float64):
                    self.never_nan.add(name)
                pi = self._ptr(("in", man, what))
                self.prolog.append((name, _CTYPE[src.dtype], pi))
            return self.svar[st]
        if x is None:
            raise NotSpecializable("None argument")
        return _lit(float(x))

    def _new_svar(self, st, dtype=None):
        """per-event scalars are registers of the variable's own type: float32 values stay float
        (no float64 round trips on the serial scalar path); everything else is carried as double"""
        name = f"s{len(self.svar)}"
        self.svar[st] = name
        self.stype[name] = "float" if dtype == torch.float32 else "double"
        return name

    def _asg(self, name, expr):
        return f"{name} = ({self.stype[name]})({expr});"

    def _sout(self, t: torch.Tensor) -> str:
        if not isinstance(t, torch.Tensor) or t.numel() != t.shape[0]:
            raise NotSpecializable("scalar output must be a [block] tensor")
        st = _storage(t)
        if st in self.svar:
            raise NotSpecializable("scalar variable written twice")
        return self._new_svar(st, t.dtype)

    def _win(self, t: torch.Tensor, need_zero_offset=False):
        if t.ndim != 2 or (t.shape[-1] > 1 and t.stride(-1) != 1):
            raise NotSpecializable("unsupported waveform view (stride)")
        st = _storage(t)
        w = self.waves.get(st)
        if w is None:
            if st not in self.input_wave:
                raise NotSpecializable("waveform operand read before it is produced")
            man, what = self.input_wave[st]
            rv = man.raw_var
            if rv.dtype not in _CTYPE:
                raise NotSpecializable(f"input waveform dtype {rv.dtype}")
            w = Wave(len(self.waves), int(rv.shape[1]))
            w.is_input = True
            w.int_valued = rv.dtype in (torch.uint16, torch.int16)
            self.waves[st] = w
            self._node("load", wouts=[w], ptr=self._ptr(("in", man, what)), dtype=rv.dtype, n=int(rv.shape[1]))
        off = t.storage_offset() % max(1, w.n) if t.storage_offset() else 0
        if need_zero_offset and off != 0:
            raise NotSpecializable("processor needs an unsliced waveform")
        if w.n > CHK * NT:
            raise NotSpecializable("waveform longer than 8192 samples")
        return w, int(off), int(t.shape[1])

    def _wout(self, t: torch.Tensor) -> Wave:
        if t.ndim != 2 or t.storage_offset() != 0:
            raise NotSpecializable("waveform outputs must be whole buffers")
        st = _storage(t)
        if st in self.waves:
            raise NotSpecializable("waveform written twice")
        n = int(t.shape[1])
        if n > CHK * NT:
            raise NotSpecializable("waveform longer than 8192 samples")
        w = Wave(len(self.waves), n)
        self.waves[st] = w
        return w

    # -- lowering: ProcessorManager -> node ----------------------------------------------------
    def _lower(self, pm):
        from . import processing_chain as pc

        if isinstance(pm, pc.UnitConversionManager):
            buf, off_in, off_out, ratio, out = pm.args
            if buf.numel() != buf.shape[0] or out.dtype not in (torch.float32, torch.float64):
                raise NotSpecializable("unit conversion of a non-scalar / integer variable")
            if pm.in_is_int and pm.mode is None:
                raise NotSpecializable("integer conversion check")
            self._node("sc_convert", x=self._sc(buf), oi=self._sc(off_in), oo=self._sc(off_out), ratio=float(ratio),
                       mode=pm.mode, f32=out.dtype == torch.float32, out=self._sout(out))
            return
        proc = pm.processor
        name = proc.__name__
        a = pm.args
        f32 = not any(t.char == "d" for t in pm.types)
        if isinstance(proc, numpy_bridge.ElementwiseOp):
            if any(isinstance(x, torch.Tensor) and x.numel() != x.shape[0] and _storage(x) not in self.const_storage
                   for x in a):
                raise NotSpecializable(f"element-wise {name} on waveforms")
            out = a[-1]
            if out.dtype not in (torch.float32, torch.float64):
                raise NotSpecializable(f"{name} with {out.dtype} output")
            if name in ("add", "subtract", "multiply", "divide", "floor_divide"):
                self._node("sc_bin", op=name, x=self._sc(a[0]), y=self._sc(a[1]), f32=out.dtype == torch.float32,
                           out=self._sout(out))
                return
            if name == "negative":
                self._node("sc_neg", x=self._sc(a[0]), out=self._sout(out))
                return
            raise NotSpecializable(f"element-wise {name}")
        if not getattr(proc, "native_kernel", False):
            raise NotSpecializable(f"helper processor {name}")
        if not f32:
            raise NotSpecializable(f"{name}: float64 type loop")

        if name == "bl_subtract":
            w, off, n = self._win(a[0], True)
            if n != w.n:
                raise NotSpecializable("bl_subtract on a slice")
            self._node("bl_sub", ins=[(w, off, n)], b=self._sc(a[1]), wouts=[self._wout(a[2])])
        elif name in ("min_max", "amax"):
            w, off, n = self._win(a[0])
            outs = [self._sout(x) for x in a[1:5]] if name == "min_max" else [None, None, None, self._sout(a[2])]
            self._node("min_max", ins=[(w, off, n)], outs=outs)
        elif name in ("linear_slope_fit", "mean_stdev"):
            w, off, n = self._win(a[0])
            outs = [self._sout(x) for x in a[1:]]
            outs += [None] * (4 - len(outs))
            self._node("lsf", ins=[(w, off, n)], outs=outs)
        elif name == "pole_zero":
            w, off, n = self._win(a[0], True)
            if n != w.n:
                raise NotSpecializable("pole_zero on a slice")
            self._node("pole_zero", ins=[(w, off, n)], tau=self._sc(a[1]), wouts=[self._wout(a[2])])
        elif name == "double_pole_zero":
            w, off, n = self._win(a[0], True)
            t1, t2, fr = (self._sc(x) for x in a[1:4])
            if n != w.n or n <= 3 or not all(x.startswith(("0x", "-0x")) for x in (t1, t2, fr)):
                raise NotSpecializable("double_pole_zero: slice / non-constant time constants (interpreted tier)")
            w.force_slot = True       # the two samples in front of a chunk are read from the slot
            self._node("dpz", ins=[(w, 0, n)], tau1=float(np.float32(float.fromhex(t1))),
                       tau2=float(np.float32(float.fromhex(t2))), frac=float(np.float32(float.fromhex(fr))),
                       wouts=[self._wout(a[4])])
        elif name in ("trap_filter", "trap_norm"):
            w, off, n = self._win(a[0], True)
            rise, flat = int(a[1]), int(a[2])
            if n != w.n or rise < 0 or flat < 0 or 2 * rise + flat > n or (rise == 0 and name == "trap_norm"):
                raise NotSpecializable("trapezoid arguments (the interpreted path raises the DSPFatal)")
            self._node("fir", ins=[(w, 0, n)], taps=[(0, 1.0), (rise, -1.0), (rise + flat, -1.0), (2 * rise + flat, 1.0)],
                       scale=(1.0 / rise) if name == "trap_norm" else 1.0, p=n, extra=None, wouts=[self._wout(a[3])])
        elif name == "asym_trap_filter":
            w, off, n = self._win(a[0], True)
            rise, flat, fall = int(a[1]), int(a[2]), int(a[3])
            if n != w.n or min(rise, flat, fall) < 0 or rise + flat + fall > n or rise == 0 or fall == 0:
                raise NotSpecializable("trapezoid arguments")
            self._node("fir", ins=[(w, 0, n)], taps=[(0, 1.0 / rise), (rise, -1.0 / rise), (rise + flat, -1.0 / fall),
                                                     (rise + flat + fall, 1.0 / fall)],
                       scale=1.0, p=n, extra=None, wouts=[self._wout(a[4])])
        elif name == "trap_pickoff":
            w, off, n = self._win(a[0], True)
            rise, flat = int(a[1]), int(a[2])
            if n != w.n or rise <= 0 or flat < 0 or 2 * rise + flat > n:
                raise NotSpecializable("trapezoid arguments")
            self._node("trap_pickoff", ins=[(w, 0, n)], rise=rise, flat=flat, t=self._sc(a[3]), out=self._sout(a[4]))
        elif name in ("moving_window_left", "moving_window_right", "moving_window_multi"):
            w, off, n = self._win(a[0], True)
            length = float(np.float32(a[1]))
            if n != w.n or length != np.floor(length) or not (1 <= int(length) < n):
                raise NotSpecializable("moving-window arguments")
            if name == "moving_window_multi":
                num, typ = float(np.float32(a[2])), int(a[3])
                if num != np.floor(num) or num < 1 or typ not in (0, 1, 2):
                    raise NotSpecializable("moving-window arguments")
                dirs = [("r" if ((k & 1) and typ == 0) or typ == 2 else "l") for k in range(int(num))]
                out = self._wout(a[4])
            else:
                dirs = ["l" if name.endswith("left") else "r"]
                out = self._wout(a[2])
            self._node("mw", ins=[(w, 0, n)], L=int(length), dirs=dirs, wouts=[out])
        elif name == "avg_current":
            w, off, n = self._win(a[0], True)
            length = float(np.float32(a[1]))
            out = self._wout(a[2])
            if n != w.n or length != np.floor(length) or not (1 <= int(length) < n) or out.n != n - int(length):
                raise NotSpecializable("avg_current arguments")
            self._node("avg_current", ins=[(w, 0, n)], L=int(length), wouts=[out])
        elif name == "time_point_thresh":
            w, off, n = self._win(a[0], True)
            if n != w.n:
                raise NotSpecializable("time_point_thresh on a slice")
            self._node("tpt", ins=[(w, 0, n)], thr=self._sc(a[1]), start=self._sc(a[2]), walk=self._sc(a[3]),
                       out=self._sout(a[4]))
        elif name == "fixed_time_pickoff":
            w, off, n = self._win(a[0], True)
            mode = P._as_int(a[2])
            if n != w.n or chr(mode) == "s":
                raise NotSpecializable("fixed_time_pickoff spline mode / slice")
            self._node("ftp", ins=[(w, 0, n)], t=self._sc(a[1]), mode=mode, out=self._sout(a[3]))
        elif name == "windower":
            w, off, n = self._win(a[0], True)
            out = self._wout(a[2])
            if n != w.n or out.n >= n:
                raise NotSpecializable("windower arguments")
            self._node("windower", ins=[(w, 0, n)], t0=self._sc(a[1]), wouts=[out])
        elif name == "upsampler":
            w, off, n = self._win(a[0], True)
            up = float(np.float32(a[1]))
            if n != w.n or up != np.floor(up) or not up >= 1:
                raise NotSpecializable("non-integer upsample factor")
            self._node("upsampler", ins=[(w, 0, n)], up=int(up), wouts=[self._wout(a[2])])
        elif name in ("convolve_wf", "fft_convolve_wf"):
            self._lower_conv(pm)
        else:
            raise NotSpecializable(f"no specialised emitter for {name}")

    def _lower_conv(self, pm):
        w_in, kernel, mode, w_out = pm.args
        if not isinstance(kernel, torch.Tensor) or _storage(kernel) not in self.const_storage:
            raise NotSpecializable("convolution kernel is not a constant")
        w, off, n = self._win(w_in)
        if off != 0:
            raise NotSpecializable("convolution of an offset slice")
        m = int(kernel.numel())
        mode = chr(P._as_int(mode))
        if m > n or mode not in "fvs":
            raise NotSpecializable("convolution arguments")
        p = {"f": n + m - 1, "v": n - m + 1, "s": n}[mode]
        coff = {"f": 0, "v": m - 1, "s": (m - 1) // 2}[mode]
        var = self.var_of_storage.get(_storage(kernel))
        origin = getattr(var, "const_origin", None)
        # (1) cusp / zac kernels: weighted prefix sums (verified against the array on a real device)
        if origin and origin[0] in ("cusp_filter", "zac_filter") and mode == "v":
            if self.meta:   # plan only: no kernel array to verify the analytic model against
                res = fusion.seg_params(origin, m)
                seg = res[0] if res else None
            else:
                seg = self._seg_model(kernel, self._const_value(kernel).astype(np.float32))
            if seg is not None and 13 * p * 8 <= 4 * (CHK * NT):
                out = self._wout(w_out)
                if out.n != p:
                    raise NotSpecializable("convolution output length mismatch")
                self._node("conv_seg", ins=[(w, 0, n)], seg=seg, p=p, wouts=[out])
                self.conv_lowering.append(("seg", m, "zac" if seg[9] else "cusp"))
                return
        k = self._const_value(kernel).astype(np.float32)
        if np.isnan(k).any():
            raise NotSpecializable("NaN in convolution kernel")
        # (2) run-structured kernels: sparse first difference -> sparse FIR + running sum
        dk = np.diff(np.concatenate([[0.0], k.astype(np.float64), [0.0]]))
        taps = np.flatnonzero(dk)
        if len(taps) <= 24 and coff <= 1024 and p <= CHK * NT:
            out = self._wout(w_out)
            if out.n != p:
                raise NotSpecializable("convolution output length mismatch")
            extra = [float(x) for x in k[:coff]] if coff > 0 else None   # full[coff-1] = sum_j k[j] y[coff-1-j]
            self._node("fir", ins=[(w, 0, n)], taps=[(int(t) - coff, float(dk[t])) for t in taps], scale=1.0, p=p,
                       extra=extra, wouts=[out])
            self.conv_lowering.append(("runs", m, len(taps)))
            return
        raise NotSpecializable("generic convolution kernel (direct lowering lives in the interpreted path)")

    # ------------------------------------------------------------------------------------
    # scheduling: hoist chunk-local reductions behind their producer, group FIRs
    # ------------------------------------------------------------------------------------
    def _schedule(self):
        nodes = self.nodes
        # A windowed filter (trapezoid ...) whose output is only picked off at single positions is
        # never materialised: the scalar warp evaluates it there from its input wave (window sums).
        for nd in nodes:
            if nd["kind"] != "fir" or nd["extra"] is not None:
                continue
            out = nd["wouts"][0]
            users = [nodes[u] for u in out.uses]
            taps = sorted(nd["taps"])
            if (users and all(u["kind"] == "ftp" and chr(u["mode"]) in "infcl" for u in users)
                    and abs(sum(c for _, c in taps)) < 1e-12 and taps[0][0] >= 0 and taps[-1][0] <= 2048):
                nd["kind"] = "fir_lazy"
                for u in users:
                    u["lazy"] = nd
                    u["ins"] = [nd["ins"][0]]
                    nd["ins"][0][0].uses.append(u["idx"])
                out.uses = []
        order = []
        placed = set()
        by_wave_reductions = {}
        for nd in nodes:
            if nd["kind"] in ("min_max", "lsf") or (nd["kind"] == "ftp" and nd["t"].startswith(("0x", "-0x"))):
                by_wave_reductions.setdefault(nd["ins"][0][0].id, []).append(nd)

        segs_of_wave = {}
        for nd in nodes:
            if nd["kind"] == "conv_seg":
                segs_of_wave.setdefault(nd["ins"][0][0].id, []).append(nd)

        def place_seg(nd):
            """cusp / zac convolutions with the same input and shared parameters are evaluated together"""
            if nd["idx"] in placed:
                return
            key = (nd["ins"][0][0].id, nd["ins"][0][2], tuple(nd["seg"][:6]))
            group = [x for x in nodes if x["kind"] == "conv_seg" and x["idx"] not in placed
                     and (x["ins"][0][0].id, x["ins"][0][2], tuple(x["seg"][:6])) == key][:2]
            if nd["ins"][0][2] >= CHK * NT:
                placed.add(nd["idx"])
                order.append(nd)
                return
            for g in group:
                placed.add(g["idx"])
            order.append(dict(kind="conv_seg_group", idx=nd["idx"], members=group, ins=[nd["ins"][0]],
                              wouts=[g["wouts"][0] for g in group], fatal=nd["fatal"]))
            for g in group:
                for r in by_wave_reductions.get(g["wouts"][0].id, []):
                    place(r)

        def place(nd):
            if nd["idx"] in placed:
                return
            placed.add(nd["idx"])
            order.append(nd)
            for w in nd.get("wouts", []):
                for r in by_wave_reductions.get(w.id, []):
                    place(r)
            # hoist the structure-aware convolutions of a fresh wave right behind it (its chunk is
            # still in registers, and its slot can be recycled early)
            for w in nd.get("wouts", []):
                for sg in segs_of_wave.get(w.id, []):
                    place_seg(sg)

        for nd in nodes:
            if nd["idx"] in placed:
                continue
            if nd["kind"] == "conv_seg":
                place_seg(nd)
                continue
            if nd["kind"] == "fir":
                # all FIR-running-sum filters of the same input wave are evaluated together
                w = nd["ins"][0][0]
                group = [x for x in nodes if x["kind"] == "fir" and x["ins"][0][0] is w and x["idx"] not in placed]
                for g in group:
                    placed.add(g["idx"])
                order.append(dict(kind="fir_group", idx=nd["idx"], members=group, ins=[nd["ins"][0]],
                                  wouts=[g["wouts"][0] for g in group], fatal=nd["fatal"]))
                for g in group:
                    for r in by_wave_reductions.get(g["wouts"][0].id, []):
                        place(r)
                continue
            place(nd)
        # Fill the block stream's wait for a scalar of the scalar stream.  A windower needs its start
        # sample from a threshold search (scalar warp, latency bound): the block warps would idle at that
        # point.  A fit / extremum reduction whose results only the scalar stream (outputs) consumes, and
        # whose input wave is still in its slot at that point anyway, is evaluated there instead of right
        # behind its producer (the chunk is re-read from the slot: 4 shared-memory loads per thread).
        if os.environ.get("DSPEED_B200_FILL_WAIT", "1") != "0":
            keys = ("x", "y", "thr", "start", "walk", "t", "t0", "b", "tau", "oi", "oo")
            wpos = next((k for k, nd in enumerate(order) if nd["kind"] == "windower" and not str(nd["t0"]).startswith(("0x", "-0x"))),
                        None)
            if wpos is not None:
                wave_last = {}
                for k, nd in enumerate(order):
                    for m in (nd["members"] if nd["kind"] in ("fir_group", "conv_seg_group", "tpt_chain") else [nd]):
                        for (w, _, _) in m.get("ins", []):
                            wave_last[w.id] = k
                best = None
                for k, nd in enumerate(order[:wpos]):
                    if nd["kind"] != "lsf":
                        continue
                    w, off, n = nd["ins"][0]
                    outs = {o for o in nd["outs"] if o}
                    used_between = any(isinstance(x.get(key), str) and x[key] in outs for x in order[k + 1:wpos + 1] for key in keys)
                    if used_between or wave_last.get(w.id, -1) < wpos or w.is_input:
                        continue
                    if best is None or n > best[1]:
                        best = (k, n)
                if best is not None:
                    nd = order.pop(best[0])
                    order.insert(wpos - 1, nd)     # wpos shifted by one after the pop
                    w = nd["ins"][0][0]
                    w.force_slot = True
        order = self._chain_searches(order)
        self.order = order
        # last use positions (in emission order) for slot liveness
        pos = {}
        for k, nd in enumerate(order):
            members = nd["members"] if nd["kind"] in ("fir_group", "conv_seg_group", "tpt_chain") else [nd]
            for m in members:
                pos[m["idx"]] = k
        for w in self.waves.values():
            w.last = max([pos[u] for u in w.uses if u in pos] + [-1])
            local = True
            for u in w.uses:
                nd = nodes[u]
                if nd["kind"] == "store_wave":
                    local &= nd["ins"][0][1] == 0
                else:
                    local &= nd["kind"] in ("min_max", "lsf", "bl_sub", "pole_zero")
            first = pos.get(w.producer, 0) if w.producer is not None else 0
            w.needs_slot = not (local and w.last - first <= 6) or getattr(w, "force_slot", False)

    def _chain_searches(self, order):
        """Backward threshold searches on one waveform in which each search starts at the result of the previous one
        (tp_95 <- tp_99, tp_90 <- tp_95 ...) become ONE node: the scalar warp resolves them in a single walk
        (chain_rt.cuh: tpt_chain_bwd).  Nodes between the members (threshold glue, unrelated searches) move in front
        of the chain when their inputs are ready there, else behind it; a member whose threshold depends on the
        chain itself ends it."""
        if os.environ.get("DSPEED_B200_TPT_CHAINS", "1") == "0":
            return order
        lit = ("0x", "-0x", "CUDART")

        def sins(nd):
            vals = []
            for key in self._SKEYS:
                v = nd.get(key)
                if isinstance(v, str) and not v.startswith(lit):
                    vals.append(v)
            return vals

        def souts(nd):
            return [o for o in ([nd.get("out")] + list(nd.get("outs", []))) if o]

        def is_bwd(nd):
            return nd["kind"] == "tpt" and nd["walk"].startswith(lit[:2]) and float.fromhex(nd["walk"]) == 0.0

        k = 0
        while k < len(order):
            head = order[k]
            if not is_bwd(head):
                k += 1
                continue
            members, pos_m = [head], [k]
            q = k + 1
            while q < len(order):
                nd = order[q]
                if (is_bwd(nd) and nd["ins"][0][0] is head["ins"][0][0] and nd["start"] == members[-1]["out"]
                        and nd["fatal"] is not None):
                    members.append(nd)
                    pos_m.append(q)
                q += 1
            if len(members) < 2 or len(members) > 8:
                k += 1
                continue
            # nodes between the first and the last member
            chain_outs = {m["out"] for m in members}
            between = [order[q] for q in range(pos_m[0] + 1, pos_m[-1]) if q not in pos_m]
            front, back, tainted = [], [], set(chain_outs)
            ok = True
            for nd in between:
                if nd["kind"] in ("fir_group", "conv_seg_group") or nd.get("wouts") or nd["kind"] in ("load", "store_wave"):
                    ok = False      # block-stream work inside the span: leave the program alone
                    break
                if any(i in tainted for i in sins(nd)):
                    back.append(nd)
                    tainted.update(souts(nd))
                else:
                    front.append(nd)
            # a member must not depend on anything that moved behind the chain
            moved_back = {o for nd in back for o in souts(nd)}
            if not ok or any(i in moved_back for m in members for i in sins(m)):
                k += 1
                continue
            node = dict(kind="tpt_chain", idx=head["idx"], members=members, ins=[head["ins"][0]], outs=[m["out"] for m in members],
                        thrs=[m["thr"] for m in members], start=head["start"], fatal=head["fatal"])
            order = order[:pos_m[0]] + front + [node] + back + order[pos_m[-1] + 1:]
            k = pos_m[0] + len(front) + 1
        return order

    # ------------------------------------------------------------------------------------
    # emission
    # ------------------------------------------------------------------------------------
    def _emit_kernel(self):
        # Two instruction streams per row: the BLOCK stream (warps 0-15, everything that touches
        # waveforms) and the SCALAR stream (warp 16: finishing of reductions, threshold searches,
        # pick-offs, scalar glue, output stores).  They meet only through named-barrier events
        # (B -> S: "partials / slots are ready", S -> B: "this scalar is ready") and at the row end,
        # so the serial per-event scalar chain runs concurrently with the block work.
        self.LB, self.LS, self.LS_tag = [], [], []
        self.pending = set()
        self.dirty = set()
        self.xread = set()
        self.posts = []
        self.sdom = {}              # scalar variable -> "s" (scalar warp only); default: both streams
        self.stored = set()
        self.post_dirty = []
        self.nd_used = 0
        self.ni_used = 0
        self.n_slots = 0
        self.tmp = 0
        self.live_regs = []
        self.n_mbd = 0              # mailbox entries (16 doubles / 16 ints each), unique per row
        self.n_mbi = 0
        self.n_bc = 0
        self.s_dirty = False        # the block stream produced something the scalar warp will read
        self.b2s_count = 0          # B -> S events since the last point where S provably caught up
        self.s2b = {}               # scalar -> (event id, bc index) published by the scalar warp
        self.s2b_ids = [0]          # barrier id for a scalar published by the scalar warp
        self.s_seq = 0              # number of scalar-stream nodes emitted
        self.s_done = 0             # ... of which the block stream knows they are finished
        self.flag_s = {}            # block-stream NaN flag -> its copy in the scalar stream
        self.s_deferred = []
        self.s_deferred_urgent = []
        self._cur_def = []
        slot_len = max([w.n for w in self.waves.values()] + [CHK])
        self.nchunks = (slot_len + CHK - 1) // CHK
        self.psp = 4 * self.nchunks + 8
        self.slot_words = 4 * self.psp
        # waves the scalar warp reads (threshold searches, pick-offs); "late" ones are still read
        # after the scalar warp has published the last scalar the block stream waits for
        self.s_read = {nd["ins"][0][0].id for nd in self.nodes if nd["kind"] in ("tpt", "ftp", "trap_pickoff")}
        pos_of = {nd["idx"]: k for k, nd in enumerate(self.order)}
        for k, nd in enumerate(self.order):
            for m in nd.get("members", []):
                pos_of[m["idx"]] = k
        b_need = set()
        for nd in self.nodes:
            for key in ("b", "tau", "t0"):
                if key in nd and isinstance(nd[key], str):
                    b_need.add(nd[key])
        last_pub = max([pos_of.get(nd["idx"], -1) for nd in self.nodes
                        if (nd.get("out") in b_need or any(o in b_need for o in nd.get("outs", []) if o))] + [-1])
        self.s_late = {nd["ins"][0][0].id for nd in self.nodes
                       if nd["kind"] in ("tpt", "ftp", "trap_pickoff") and pos_of.get(nd["idx"], 0) > last_pub + 1}
        # waves searched by the scalar warp carry a min/max summary (built when they are stored)
        self.summ = {}          # wave id -> (first summary cell, one cell per row parity?)
        self.n_summ = 0
        for nd in self.nodes:
            if nd["kind"] == "tpt":
                w = nd["ins"][0][0]
                if w.id not in self.summ:
                    par2 = w.id in self.s_late   # still searched while the block stream is in the next row
                    self.summ[w.id] = (self.n_summ, par2)
                    self.n_summ += 2 if par2 else 1
        # fixed pool of physical slots: everything the 227 KB allow
        self.MB_BUDGET = 2048       # mailbox bytes per row parity
        fixed_est = 2048 + 8192 + 2048 + 2 * self.MB_BUDGET + self.n_summ * 4224
        total_slots = (MAX_SMEM - fixed_est) // (self.slot_words * 4)
        if total_slots < 2:
            raise NotSpecializable("waveforms too long for the shared-memory resident layout")
        total_slots -= total_slots % 2       # two frames of the pool alternate with the row parity
        if self._slot_limit:
            total_slots = min(total_slots, self._slot_limit)
        self.free_cols = [(k, 0, self.nchunks) for k in range(total_slots)]
        self.total_slots = total_slots
        self.written_slots = set()
        self.top_slots = set()
        # scalars the block stream needs from the scalar warp
        self.b_needed = set()
        for nd in self.nodes:
            for key in ("b", "tau", "t0"):
                if key in nd and isinstance(nd[key], str) and not nd[key].startswith(("0x", "-0x", "CUDART")):
                    self.b_needed.add(nd[key])
        self.urgent_names = set()
        for nd in self.nodes:
            if nd.get("out") in self.b_needed:
                for key in ("thr", "start", "walk", "t", "x", "y", "oi", "oo"):
                    if isinstance(nd.get(key), str) and not nd[key].startswith(("0x", "-0x", "CUDART")):
                        self.urgent_names.add(nd[key])
        regions = self._plan_regions()
        self._plan_scalar_warps()
        self.region_nw = None
        for k, nd in enumerate(self.order):
            self.pos = k
            self._es_later_end()
            if k in regions:
                # a run of nodes that only touch short waveforms: warps that own no chunk of them skip the arithmetic
                # and only mirror the synchronisation (see _expand_regions)
                self._close_round()
                self.region_nw = regions[k][1]
                self.LB.append(f"//@REGION_BEGIN {self.region_nw}")
            self.LB.append(f"// ---- [{k}] {self._describe_node(nd)}")
            self._ls_add([f"// ---- [{k}] {self._describe_node(nd)}"], "all")
            getattr(self, "_e_" + nd["kind"])(nd)
            self.LB.append("PROF_MARK(%d);" % k)
            self._ls_add(["PROF_MARK_S(%d);" % k], "all")
            self._release(k)
            if self.region_nw is not None and any(k == e for (e, _) in regions.values()):
                self._close_round()
                self.LB.append("//@REGION_END")
                self.region_nw = None
                for w in list(self.live_regs):    # register chunks defined inside the region are out of scope
                    w.reg = None
                self.live_regs = []
        self._close_round()
        self._flush_s()
        # scalar outputs not stored at their definition (pass-through input scalars)
        for k, (name, pi, ct) in enumerate(self.out_scalars):
            if name not in self.stored:
                self._es(f"if (lane == {k % 32}) (({ct}*)A.p[{pi}])[row] = ({ct}){name};", tag="leaf")
        self.LB.append("PROF_MARK(%d);" % len(self.order))
        self._ls_add(["PROF_MARK_S(%d);" % len(self.order)], "all")
        self.mb_bytes = (self.n_mbd * 16 * 8 + self.n_mbi * 16 * 4 + 15) & ~15
        if self.mb_bytes > self.MB_BUDGET:
            raise NotSpecializable("too many reductions for the mailbox")
        # Scratch + CScr/bc + prof stamps + mailbox (two row parities) + summaries
        self.fixed_bytes = 2048 + 8192 + 2048 + 2 * self.MB_BUDGET + self.n_summ * 4224
        self.logical_slots_used = self.n_slots
        self.n_slots = self.total_slots
        self.smem_bytes = self.fixed_bytes + self.n_slots * self.slot_words * 4
        self._pipeline_rows()
        if self.smem_bytes > MAX_SMEM:
            raise NotSpecializable("not enough shared memory for the live waveforms")
        self.program_text = "\n".join(f"{k:3d} {self._describe_node(nd)}" for k, nd in enumerate(self.order))
        self.code = self.order  # (for len(code) users)

    _SKEYS = ("x", "y", "thr", "start", "walk", "t", "t0", "b", "tau", "oi", "oo")

    def _plan_scalar_warps(self):
        """Split the per-event scalar work over TWO scalar warps when it is the longer of the two pipelines.

        Warp 1 (warp 16 of the CTA) owns everything a scalar the block stream waits for depends on (the path to
        tp_0_est in the ICPC chain), publishes it, and then carries on with its share of the rest; warp 2 (warp 17)
        takes whole dependency components of the remaining searches / pick-offs / glue.  Reductions that are only
        finished from the mailbox (min_max, linear_slope_fit, convolution sinks) are cheap and have no scalar inputs:
        they are evaluated by every warp that needs one of their results and stored by one.  The only scalars that
        cross are results of warp 1's latency-critical path, handed over once per row (`xfer`, one named-barrier
        event); warp 2 never feeds warp 1 or the block stream.
        self.swarp: program position -> set of scalar warps; self.store_owner: position -> warp that stores the outputs
        """
        order = self.order
        lit = ("0x", "-0x", "CUDART")
        info = {}
        for k, nd in enumerate(order):
            kind = nd["kind"]
            if kind in ("fir_group", "conv_seg_group"):
                continue
            outs = [o for o in ([nd.get("out")] + list(nd.get("outs", []))) if o]
            ins = [nd[key] for key in self._SKEYS if isinstance(nd.get(key), str) and not nd[key].startswith(lit)]
            ins += [t for t in nd.get("thrs", []) if not t.startswith(lit)]
            if not outs:
                continue
            fin = kind in ("min_max", "lsf") or (kind == "ftp" and not ins and nd.get("lazy") is None and any(
                w.producer is not None and self.nodes[w.producer]["kind"] == "conv_seg" for (w, _, _) in nd["ins"]))
            cost = {"tpt": 10, "tpt_chain": 10 + 2 * len(nd.get("members", [])), "ftp": 8 if nd.get("lazy") is not None else 4,
                    "trap_pickoff": 8, "min_max": 2, "lsf": 3}.get(kind, 1)
            info[k] = dict(outs=outs, ins=ins, fin=fin, cost=cost)
        producer = {o: k for k, d in info.items() for o in d["outs"]}
        # glue that only depends on inputs and mailbox results (unit offsets, thresholds ...) is as cheap to repeat
        for k in sorted(info):
            d = info[k]
            if not d["fin"] and order[k]["kind"].startswith("sc_") and all(
                    i not in producer or info[producer[i]]["fin"] for i in d["ins"]):
                d["fin"] = True
        # urgent: ancestors of the scalars the block stream needs
        urgent, stack = set(), [producer[n] for n in self.b_needed if n in producer]
        while stack:
            k = stack.pop()
            if k in urgent:
                continue
            urgent.add(k)
            stack.extend(producer[i] for i in info[k]["ins"] if i in producer)
        self.swarp = {k: {1} for k in info}
        self.store_owner = {k: 1 for k in info}
        self.xfer = []
        self.n_swarps = 1
        regular = [k for k, d in info.items() if not d["fin"] and k not in urgent]
        if os.environ.get("DSPEED_B200_SCALAR_WARPS", "2") == "1" or not regular:
            return
        # dependency components of the remaining regular nodes
        parent = {k: k for k in regular}

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        for k in regular:
            for i in info[k]["ins"]:
                pk = producer.get(i)
                if pk in parent:
                    parent[find(k)] = find(pk)
        comps = {}
        for k in regular:
            comps.setdefault(find(k), []).append(k)
        load = {1: sum(info[k]["cost"] for k in urgent), 2: 0}
        assign = {}
        for root, members in sorted(comps.items(), key=lambda kv: -sum(info[k]["cost"] for k in kv[1])):
            wsel = 2 if load[2] <= load[1] else 1
            load[wsel] += sum(info[k]["cost"] for k in members)
            for k in members:
                assign[k] = wsel
        if load[2] < 20:
            return      # too little to hand over: one scalar warp as before
        self.n_swarps = 2
        for k, wsel in assign.items():
            self.swarp[k] = {wsel}
            self.store_owner[k] = wsel
        # mailbox finishers: wherever a consumer lives; chain outputs are stored by warp 2 (warp 1 keeps the critical path)
        for k in sorted(info, reverse=True):
            d = info[k]
            if not d["fin"]:
                continue
            need = set()
            for k2, d2 in info.items():
                if k2 > k and any(o in d2["ins"] for o in d["outs"]):
                    need |= self.swarp[k2]
            if k in urgent:
                need.add(1)
            has_out = any(o in self.out_of for o in d["outs"])
            if has_out or not need:
                need.add(2)
            self.swarp[k] = need
            self.store_owner[k] = 2 if 2 in need else 1
        # scalars of warp 1's regular nodes that warp 2 reads
        for k, d in info.items():
            if 2 in self.swarp[k] and not d["fin"]:
                for i in d["ins"]:
                    pk = producer.get(i)
                    if pk is not None and not info[pk]["fin"] and self.swarp[pk] == {1} and i not in self.xfer:
                        if pk not in urgent:
                            raise NotSpecializable("internal: scalar warps are not independent")
                        self.xfer.append(i)

    def _split_scalar(self):
        """tagged scalar-stream lines -> one list per scalar warp (see _plan_scalar_warps)"""
        ns = self.n_swarps
        streams = {wn: [] for wn in range(1, ns + 1)}
        tags = {wn: [] for wn in range(1, ns + 1)}
        for tag, ln in zip(self.LS_tag, self.LS):
            if tag in ("all", "sync"):
                targets = range(1, ns + 1)
            elif tag == "leaf":
                targets = [ns]
            else:
                targets = sorted(w_ for w_ in self.swarp.get(tag, {1}) if w_ <= ns)
            for wn in targets:
                if "/*store*/" in ln and not isinstance(tag, str) and self.store_owner.get(tag, 1) != wn and len(targets) > 1:
                    continue
                if wn > 1 and ln.startswith("PROF_MARK_S("):
                    ln = ln.replace("PROF_MARK_S(", "PROF_MARK_S2(")
                streams[wn].append(ln)
                tags[wn].append(tag)
        if ns == 2 and self.xfer:
            if self.n_bc + len(self.xfer) > 16:
                raise NotSpecializable("too many scalars cross between the streams")
            cell = {n: 15 - i for i, n in enumerate(self.xfer)}
            owner = {n: next(k for k, nd in enumerate(self.order)
                             if n == nd.get("out") or n in [o for o in nd.get("outs", []) if o]) for n in self.xfer}
            # warp 1: each scalar goes to its cell behind the node that defines it, the event behind the last one
            last_line = {}
            for idx, tag in enumerate(tags[1]):
                for n, k in owner.items():
                    if tag == k:
                        last_line[n] = idx
            ins_at = {}
            for n, idx in last_line.items():
                ins_at.setdefault(idx, []).append(f"if (lane == 0) bc[{cell[n]}] = (double){n};")
            ev = max(last_line.values())
            out1 = []
            for idx, ln in enumerate(streams[1]):
                out1.append(ln)
                out1.extend(ins_at.get(idx, []))
                if idx == ev:
                    out1.append("EV_ARRIVE(EVX);   // hand-over to the second scalar warp")
            streams[1] = out1
            # warp 2: fetched in front of the first node that reads one of them
            users = set()
            for k, nd in enumerate(self.order):
                if 2 in self.swarp.get(k, set()) and any(nd.get(key) in cell for key in self._SKEYS):
                    users.add(k)
            first = next(idx for idx, tag in enumerate(tags[2]) if tag in users)
            fetch = ["EV_WAIT(EVX);"] + [self._asg(n, f"bc[{cell[n]}]") for n in self.xfer]
            streams[2] = streams[2][:first] + fetch + streams[2][first:]
        return streams

    def _extent(self, nd):
        """number of samples the block-stream code of a node spans (0: scalar-stream only)"""
        kind = nd["kind"]
        if kind in ("tpt", "tpt_chain", "ftp", "trap_pickoff", "fir_lazy", "sc_bin", "sc_neg", "sc_convert"):
            return 0
        if kind == "windower":
            return nd["wouts"][0].n + CHK
        if kind == "min_max" and (nd.get("from_conv") or any(
                w.producer is not None and self.nodes[w.producer]["kind"] == "conv_seg" for (w, _, _) in nd["ins"])):
            return 0    # (the convolution hands its maxima to the scalar warp itself; at worst the region is missed)
        if kind in ("min_max", "lsf", "avg_current", "upsampler", "mw"):
            return max([w.n for (w, _, _) in nd.get("ins", [])] + [w.n for w in nd.get("wouts", [])])
        return CHK * NT

    def _plan_regions(self):
        """{first node: (last node, warps)} for runs of consecutive nodes whose waveforms fit into at most 12 warps"""
        if os.environ.get("DSPEED_B200_REGIONS", "1") == "0":
            return {}
        lim = 12 * 32 * CHK
        ext = [self._extent(nd) for nd in self.order]
        regions, k = {}, 0
        while k < len(ext):
            if not (0 < ext[k] <= lim):
                k += 1
                continue
            e = k
            last = k
            while e + 1 < len(ext) and ext[e + 1] <= lim:
                e += 1
                if ext[e] > 0:
                    last = e
            if sum(1 for q in range(k, last + 1) if ext[q] > 0) >= 2:
                regions[k] = (last, -(-max(ext[k:last + 1]) // (32 * CHK)))
            k = last + 1
        return regions

    def _nwarg(self):
        return f", {self.region_nw}" if getattr(self, "region_nw", None) else ""

    def _expand_regions(self, lines):
        """//@REGION markers -> `if (warp < nw) { ... } else { <synchronisation only> }`.  The idle warps execute the
        same barriers, events and scalar fetches in the same order (the hazard waits that _pipeline_rows inserted
        included); flags that outlive the region are declared in front of it."""
        out, k = [], 0
        keep = re.compile(r"^\s*(BSYNC\(\)|EV_ARRIVE|EV_WAIT|if \(it > 0\) EV_WAIT|par \^= 1|s\d+ = )")
        while k < len(lines):
            ln = lines[k]
            if not ln.startswith("//@REGION_BEGIN"):
                out.append(ln)
                k += 1
                continue
            nw = int(ln.split()[1])
            e = next(i for i in range(k + 1, len(lines)) if lines[i].startswith("//@REGION_END"))
            body, idle, decls = [], [], []
            for b in lines[k + 1:e]:
                m = re.match(r"^(\s*)const int (nan\w+) = (get_imax\(.*)$", b)
                if m:
                    decls.append(f"int {m.group(2)} = 0;")
                    b = f"{m.group(1)}{m.group(2)} = {m.group(3)}"
                    idle.append(b)
                elif keep.match(b):
                    idle.append(b)
                body.append(b)
            out.extend(decls)
            out.append(f"if (warp < {nw}) {{   // ---- region: waveforms of at most {nw * 32 * CHK} samples")
            out.extend(body)
            out.append("} else {   // warps without a chunk of these waveforms: synchronisation only")
            out.extend(idle)
            out.append("}")
            k = e + 1
        return out

    def _pipeline_rows(self):
        """Software pipelining across rows.  The block stream starts row r+1 while the scalar warps are
        still finishing row r; mailbox / broadcast cells and the B -> S event barriers alternate with
        the row parity, and two waits protect the shared-memory slots: before its first write to
        the slot of a wave a scalar warp reads late, the block stream waits for "scalar warps done
        with the previous row" (barrier 15); for the writes before that point (the head: raw-data
        front end, scratch tables) it waits for a progress event (barrier 14) that every scalar warp
        posts right after its last read of any slot the head overwrites."""
        def overlap(a, b):
            return a[0] == b[0] and a[1] < b[2] and b[1] < a[2]

        self.SW = self._split_scalar()          # scalar warp -> lines
        ns = self.n_swarps
        reads = {}          # warp -> [(ordinal, line index, region)] of slot reads, in program order
        for wn, lines in self.SW.items():
            reads[wn] = []
            for idx, ln in enumerate(lines):
                if ln.startswith("//@R "):
                    sl, c0, c1, seq = (int(x) for x in ln.split()[1:5])
                    reads[wn].append((seq, idx, (sl, c0, c1)))     # (seq: one number per reading node, increasing)
        writes = []         # (LB index, region as seen from the previous row's frame)
        for k, ln in enumerate(self.LB):
            if ln.startswith("//@W "):
                sl, c0, c1 = (int(x) for x in ln.split()[1:4])
                writes.append((k, (self._rot(sl) if sl < 1000 else sl, c0, c1)))
        # group the block stream's writes by node; need[w][k] = latest read of scalar warp w (of the previous
        # row) that node k's writes collide with
        node_of, cur = {}, -1
        starts = {}
        for k, ln in enumerate(self.LB):
            if ln.startswith("// ---- ["):
                cur += 1
                starts[cur] = k
            node_of[k] = cur
        need = {wn: {} for wn in reads}
        for k, reg in writes:
            for wn in reads:
                q = max([o for o, _, r in reads[wn] if overlap(reg, r)] + [0])
                need[wn][node_of[k]] = max(need[wn].get(node_of[k], 0), q)
        self.late_idx, self.progress_idx, self.progress_seq = None, None, 0
        hot = sorted({nk for wn in need for nk, q in need[wn].items() if q > 0})
        prog = {wn: 0 for wn in reads}      # ordinal of the read behind which warp wn posts its progress event
        if hot:
            first = hot[0]
            prog = {wn: need[wn].get(first, 0) for wn in reads}   # early group: covered by the progress event
            self.progress_idx = starts[first]
            self.progress_seq = 1
            later = [nk for nk in hot if any(need[wn].get(nk, 0) > prog[wn] for wn in reads)]
            if later:
                self.late_idx = starts[later[0]]     # everything beyond waits for "scalar warps done"
            if all(prog[wn] >= max(o for o, _, _ in reads[wn]) for wn in reads if reads[wn]):
                # the first collision is already with the scalar warps' last reads
                self.late_idx, self.progress_idx, self.progress_seq = starts[first], None, 0
        # A scalar warp must not post the progress event of row r + 1 before every block warp has consumed the one of
        # row r (counted arrivals: a second arrival in the same phase completes it early and the late block warp hangs in
        # the next one).  It therefore posts behind its wait for the first block -> scalar event that the block stream
        # signals AFTER its own progress wait: passing that wait proves that all block warps are beyond theirs.
        safe_ev = None
        if self.progress_seq:
            m = next((re.search(r"EV_ARRIVE\(EVB\((\d+)\)\)", ln) for ln in self.LB[self.progress_idx:]
                      if "EV_ARRIVE(EVB(" in ln), None)
            if m is None:
                self.late_idx, self.progress_idx, self.progress_seq = starts[first], None, 0
            else:
                safe_ev = f"EV_WAIT(EVB({m.group(1)}));"
        # scalar streams: progress event behind read `prog[w]` (at the next node header), done event at the end
        for wn in list(self.SW):
            LS = list(self.SW[wn])
            if self.progress_seq:
                if prog[wn] > 0:
                    k = max(idx for o, idx, _ in reads[wn] if o == prog[wn])
                    nxt = next((i for i in range(k + 1, len(LS)) if LS[i].startswith("// ---- [")), len(LS))
                else:
                    nxt = 0     # nothing of this warp collides: it may be overwritten right away
                nxt = max(nxt, LS.index(safe_ev) + 1)
                LS.insert(nxt, "EV_ARRIVE(14);")
            LS.append("EV_ARRIVE(15);")
            self.SW[wn] = LS
        LB = list(self.LB)
        # The "done" wait must precede the row's LAST block -> scalar event: the scalar warps cannot finish row r + 1
        # (and arrive on barrier 15 again) before they have consumed that event, so every block warp has passed its wait for
        # row r by then.  With the wait behind the last event (chains without slot hazards used to get it at the very end)
        # a scalar warp with little work per row could arrive twice in one phase of barrier 15 while a block warp was
        # still on its way to the wait: the phase completed early and the late warp deadlocked in the next one.
        last_ev = max([k for k, ln in enumerate(LB) if "EV_ARRIVE(EVB(" in ln] + [-1])
        pos = self.late_idx if self.late_idx is not None else len(LB)
        if 0 <= last_ev < pos:
            pos = last_ev
        # (never inside a short-waveform region: its idle branch is generated from these lines, see _expand_regions,
        # which copies the waits, but keep the wait on the common path where the hazard analysis put it)
        LB.insert(pos, "if (it > 0) EV_WAIT(15);   // the scalar warps are done with the previous row")
        if self.progress_seq:   # (inserted second: progress_idx < late_idx)
            LB.insert(self.progress_idx, "if (it > 0) EV_WAIT(14);   // the scalar warps are past their reads of these slots")
        self.LB = LB

    def _describe_node(self, nd):
        if nd["kind"] == "conv_seg_group":
            return "conv_seg_group " + " | ".join(f"{m['wouts'][0].name}<-{m['ins'][0][0].name}[:{m['ins'][0][2]}] "
                                                  f"{'zac' if m['seg'][9] else 'cusp'} L={int(m['seg'][3])}" for m in nd["members"])
        if nd["kind"] == "fir_group":
            return "fir_group " + " | ".join(f"{m['wouts'][0].name}<-{m['ins'][0][0].name} taps={len(m['taps'])}" for m in nd["members"])
        ins = ",".join(f"{w.name}[{o}:{o + n}]" for w, o, n in nd.get("ins", []))
        outs = ",".join(w.name for w in nd.get("wouts", []))
        so = nd.get("out") or ",".join(str(x) for x in nd.get("outs", []) if x)
        return f"{nd['kind']} {ins} -> {outs}{so}"

    # -- stream / round framework -------------------------------------------------------------
    def _e(self, *lines):
        """block stream"""
        self.LB.extend(lines)

    def _ls_add(self, lines, tag=None):
        """scalar-stream lines, tagged with the program position of the node they belong to ("all": every scalar
        warp, "sync": block -> scalar events, "leaf": the last scalar warp); see _split_scalar"""
        tag = self.pos if tag is None else tag
        for ln in lines:
            self.LS.append(ln)
            self.LS_tag.append(tag)

    def _es(self, *lines, urgent=False, tag=None):
        """scalar stream"""
        self._flush_s(urgent)
        self._ls_add(lines, tag)

    def _es_later(self, *lines):
        """scalar-stream code that only consumes mailbox partials: consecutive reductions are
        collected and released behind ONE block -> scalar event.  A node's block that feeds a scalar
        the block stream waits for is released first (and alone, when the consumer is on that path)."""
        self._cur_def.extend((self.pos, ln) for ln in lines)

    def _es_later_end(self):
        """end of a node: file its deferred block as urgent or not"""
        if self._cur_def:
            urgent = any(re.match(rf"\s*{n} = ", ln) for _, ln in self._cur_def for n in self.urgent_names)
            (self.s_deferred_urgent if urgent else self.s_deferred).extend(self._cur_def)
            self._cur_def = []

    def _flush_s(self, urgent=False):
        self._es_later_end()
        if self.s_deferred_urgent or (self.s_deferred and not urgent):
            self._sync_s()
        if self.s_deferred_urgent:
            d, self.s_deferred_urgent = self.s_deferred_urgent, []
            for tag, ln in d:
                self._ls_add([ln], tag)
        if self.s_deferred and not urgent:
            d, self.s_deferred = self.s_deferred, []
            for tag, ln in d:
                self._ls_add([ln], tag)

    def _round_open(self):
        return bool(self.posts or self.nd_used or self.ni_used or self.pending)

    def _close_round(self):
        if self._round_open():
            self._e("BSYNC();")
            self._e(*self.posts)
            self._e("par ^= 1;")
        self.posts = []
        self.post_mail = False
        self.pending.clear()
        self.dirty.clear()
        self.xread.clear()
        self.dirty.update(self.post_dirty)   # chunks stored by the post-barrier code
        self.post_dirty = []
        self.nd_used = self.ni_used = 0

    def _barrier(self):
        if self._round_open():
            self._close_round()
        else:
            self._e("BSYNC();")
            self.dirty.clear()
            self.xread.clear()

    def _need(self, *exprs):
        """block-stream names (waves, flags) that must be defined now"""
        for e in exprs:
            if e is not None and str(e) in self.pending:
                self._close_round()
                return

    # ---- events -----------------------------------------------------------------------------
    def _sync_s(self):
        """the scalar warp is about to read something the block stream produced: one event
        (bar.arrive by the 512 block threads, bar.sync by the scalar warp)"""
        if not self.s_dirty:
            return
        if getattr(self, "post_mail", False):
            self._close_round()      # mailbox entries written by post-barrier code precede the event
        self.s_dirty = False
        if self.b2s_count >= 5:
            raise NotSpecializable("more block -> scalar events per row than named barriers")
        eid = self.b2s_count
        self.b2s_count += 1
        self._e(f"EV_ARRIVE(EVB({eid}));")
        self._ls_add([f"EV_WAIT(EVB({eid}));"], "sync")

    def _is_s(self, e):
        return e is not None and self.sdom.get(str(e)) == "s"

    def _def_s(self, name):
        """`name` was just defined in the scalar stream; publish it when the block stream needs it"""
        self.sdom[name] = "s"
        if name in self.b_needed:
            if not self.s2b_ids or self.n_bc >= 16:
                raise NotSpecializable("too many scalars flow from the scalar warp to the block warps")
            eid, k = self.s2b_ids.pop(0), self.n_bc
            self.n_bc += 1
            self._es(f"if (lane == 0) bc[{k}] = (double){name};", f"EV_ARRIVE({eid});")
            self.s2b[name] = (eid, k, self.s_seq)

    def _def_s_later(self, name):
        self.sdom[name] = "s"
        if name in self.b_needed:
            if not self.s2b_ids or self.n_bc >= 16:
                raise NotSpecializable("too many scalars flow from the scalar warp to the block warps")
            eid, k = self.s2b_ids.pop(0), self.n_bc
            self.n_bc += 1
            self._es_later(f"if (lane == 0) bc[{k}] = (double){name};", f"EV_ARRIVE({eid});")
            self.s2b[name] = (eid, k, self.s_seq)

    def _need_all(self, *exprs):
        """scalars the block stream must hold: wait for the scalar warp's event and fetch them (or close the round that
        defines a scalar the block warps finish themselves)"""
        self._need(*exprs)
        for e in exprs:
            if self._is_s(e):
                name = str(e)
                if name not in self.s2b:
                    raise NotSpecializable("internal: scalar not published to the block stream")
                eid, k, seq = self.s2b.pop(name)
                self._flush_s()
                self._close_round()
                self._e(f"EV_WAIT({eid});", self._asg(name, f"bc[{k}]"))
                self.s2b_ids.append(eid)
                self.s_done = max(self.s_done, seq)
                self.sdom[name] = "both"

    def _flag_s(self, flag):
        """copy of a block-stream NaN flag in the scalar stream"""
        if flag == "0":
            return "0"
        if flag not in self.flag_s:
            self._need(flag)
            k = self.n_mbi
            self.n_mbi += 1
            self._e(f"if (tid == 0) MBI({k})[0] = {flag};")
            self.s_dirty = True
            self._sync_s()
            nm = flag + "_s"
            self._es(f"const int {nm} = MBI({k})[0];", tag="all")
            self.flag_s[flag] = nm
        return self.flag_s[flag]

    def _s_wave(self, w: Wave):
        """the scalar warp is about to read the slot of `w`"""
        self._need(w.name)
        if w.slot is None:
            raise NotSpecializable("internal: scalar-warp access to a register-only wave")
        self._sync_s()
        self.s_seq += 1
        w.s_last = self.s_seq if self.swarp.get(self.pos, {1}) == {1} else 10 ** 9   # (no event proves the 2nd warp's progress)
        self._es(f"//@R {w.slot[0]} {w.slot[1]} {w.slot[1] + w.slot[2]} {self.s_seq}")

    @staticmethod
    def _wmark(sl):
        return f"//@W {sl[0]} {sl[1]} {sl[1] + sl[2]}"

    def _summ(self, w: Wave) -> str:
        k, par2 = self.summ[w.id]
        return f"SUMM({k} + rp)" if par2 else f"SUMM({k})"

    def _summ_mark(self, w: Wave) -> str:
        """summaries that exist once are part of the cross-row hazard analysis (pseudo slot 1000 + k)"""
        k, par2 = self.summ[w.id]
        return "//" if par2 else f"//@W {1000 + k} 0 1"

    def _stores(self, name):
        """statements (scalar warp) that write a just-defined scalar to its output columns: results
        leave the register file as soon as they are final"""
        self.stored.add(name)
        return [f"if (lane == {k % 32}) (({ct}*)A.p[{pi}])[row] = ({ct}){name};   /*store*/" for (pi, ct, k) in self.out_of.get(name, [])]

    def _mbd(self, k=1):
        b = self.n_mbd
        self.n_mbd += k
        return b

    def _mbi(self, k=1):
        b = self.n_mbi
        self.n_mbi += k
        return b

    def _alloc_d(self, k):
        if self.nd_used + k > 16:
            self._close_round()
        b = self.nd_used
        self.nd_used += k
        return b

    def _alloc_i(self, k):
        if self.ni_used + k > 8:
            self._close_round()
        b = self.ni_used
        self.ni_used += k
        return b

    def _t(self, prefix="t"):
        self.tmp += 1
        return f"{prefix}{self.tmp}"

    # -- slots and register chunks ---------------------------------------------------------
    # A physical slot has `nchunks` chunk columns; a wave of n samples needs ceil(n / 16) of them,
    # so several short waves (windowed leading edge, cusp / zac outputs ...) share one slot.
    # Rows alternate between two frames of the slot pool: logical slot k is physical slot k in even
    # rows and (k + T/2) % T in odd rows (T = pool size, even).  A full-size wave the scalar warp
    # reads late is placed in a logical slot whose image in the other frame has not been written by
    # the head of a row: then the head of row r+1 never touches what the scalar warp still reads
    # of row r, and (by symmetry) neither does its own late-read wave.
    def _rot(self, slot):
        return (slot + self.total_slots // 2) % self.total_slots

    def _slot_alloc(self, ncols=None, top=False, late=False):
        ncols = self.nchunks if ncols is None else min(self.nchunks, ncols)
        fits = [iv for iv in self.free_cols if iv[2] - iv[1] >= ncols]
        if not fits:
            raise NotSpecializable("not enough shared memory for the live waveforms")
        if top:
            # 1st choice: the image in the other frame is untouched so far; 2nd: it at least does not
            # hold a wave the scalar warp still reads late in the row
            safe = ([iv for iv in fits if self._rot(iv[0]) not in self.written_slots]
                    or [iv for iv in fits if self._rot(iv[0]) not in self.top_slots])
            slot, c0, c1 = min(safe or fits, key=lambda iv: iv[0])
            if late:
                self.top_slots.add(slot)
        else:
            slot, c0, c1 = min(fits, key=lambda iv: (iv[2] - iv[1] if ncols < self.nchunks else 0, iv[0]))
        self.free_cols.remove((slot, c0, c1))
        if c1 - c0 > ncols:
            self.free_cols.append((slot, c0 + ncols, c1))
        self.n_slots = max(self.n_slots, slot + 1)
        self.written_slots.add(slot)
        return (slot, c0, ncols)

    def _slot_alloc_adjacent(self, k):
        """k whole physical slots that are adjacent in shared memory in BOTH row frames (large scratch
        tables): inside one half of the pool"""
        full = sorted(sl for (sl, c0, c1) in self.free_cols if c0 == 0 and c1 == self.nchunks)
        half = self.total_slots // 2
        start = None
        for a in full:
            if all((a + d) in full for d in range(k)) and a // half == (a + k - 1) // half:
                start = a
                break
        if start is None:
            raise NotSpecializable("not enough shared memory for the scratch table")
        res = []
        for d in range(k):
            self.free_cols.remove((start + d, 0, self.nchunks))
            res.append((start + d, 0, self.nchunks))
            self.written_slots.add(start + d)
        self.n_slots = max(self.n_slots, start + k)
        return res

    def _slot_free(self, sl):
        slot, c0, nc = sl
        self.free_cols.append((slot, c0, c0 + nc))
        # merge adjacent free ranges of the same slot
        self.free_cols.sort()
        merged = []
        for iv in self.free_cols:
            if merged and merged[-1][0] == iv[0] and merged[-1][2] == iv[1]:
                merged[-1] = (iv[0], merged[-1][1], iv[2])
            else:
                merged.append(iv)
        self.free_cols = merged

    def _release(self, k):
        for w in self.waves.values():
            if getattr(w, "s_last", 0) > self.s_done:
                continue   # the scalar warp may still be reading it (its slot is kept to the row end)
            if w.slot is not None and w.last <= k and not getattr(w, "released", False):
                self._slot_free(w.slot)
                w.released = True

    def _slot(self, w: Wave) -> str:
        return self._slot_expr(w.slot)

    def _zc(self, w: Wave) -> int:
        """index (relative to the wave's slot pointer) of the always-zero pad column"""
        return self.nchunks - w.slot[1]

    @staticmethod
    def _slot_expr(sl) -> str:
        return f"SLOT({sl[0]})" if sl[1] == 0 else f"(SLOT({sl[0]}) + {4 * sl[1]})"

    def _give_slot(self, w: Wave, post=False):
        if w.slot is None:
            w.slot = self._slot_alloc((w.n + CHK - 1) // CHK, top=w.id in self.s_read, late=w.id in self.s_late)
            # a pre-barrier store must not overtake other threads still reading the old tenant
            if not post and w.slot[0] in self.xread:
                self._barrier()

    def _chunk(self, w: Wave) -> str:
        """register chunk (own 16 samples) of a wave"""
        self._need(w.name)
        if w.reg is not None:
            return w.reg
        if w.slot is None:
            raise NotSpecializable("internal: wave has neither registers nor a slot")
        if w.scatter and w.slot[0] in self.dirty:
            self._barrier()
        r = self._t("r")
        self._e(f"float {r}[16]; ld_chunk_n({self._slot(w)}, tid, {w.n}, {r});")
        self._set_reg(w, r)
        return r

    def _set_reg(self, w: Wave, r: str):
        w.reg = r
        if w in self.live_regs:
            self.live_regs.remove(w)
        self.live_regs.append(w)
        droppable = [x for x in self.live_regs if x.needs_slot]
        while len(droppable) > 2:
            x = droppable.pop(0)
            x.reg = None
            self.live_regs.remove(x)

    def _store(self, w: Wave, r: str):
        """own chunk -> slot (when other threads / later phases need the wave)"""
        if w.needs_slot:
            self._give_slot(w)
            self._e(self._wmark(w.slot), f"st_chunk_n({self._slot(w)}, tid, {w.n}, {r});")
            if w.id in self.summ:
                self._e(self._summ_mark(w), f"put_summary({self._summ(w)}, {r}, {w.n}, tid, lane, warp);")
            self.dirty.add(w.slot[0])
            self.s_dirty = True
        self._set_reg(w, r)

    def _post_store(self, w: Wave, r: str):
        """like _store, but the chunk is produced by post-barrier code of the open round"""
        if w.needs_slot:
            self._give_slot(w, post=True)
            self.posts.append(self._wmark(w.slot))
            self.posts.append(f"st_chunk_n({self._slot(w)}, tid, {w.n}, {r});")
            if w.id in self.summ:
                self.posts.append(self._summ_mark(w))
                self.posts.append(f"put_summary({self._summ(w)}, {r}, {w.n}, tid, lane, warp);")
            self.post_dirty.append(w.slot[0])
            self.s_dirty = True
        self._set_reg(w, r)

    def _visible(self, w: Wave):
        """make the slot of `w` readable at other threads' positions"""
        self._need(w.name)
        if w.slot is None:
            raise NotSpecializable("internal: cross-thread access to a register-only wave")
        if w.slot[0] in self.dirty:
            self._barrier()
        self.xread.add(w.slot[0])

    # -- node emitters -----------------------------------------------------------------------
    def _e_load(self, nd):
        w = nd["wouts"][0]
        r = self._t("r")
        pi, n, dt = nd["ptr"], nd["n"], nd["dtype"]
        ct = _CTYPE[dt]
        self._e(f"float {r}[16];")
        if dt == torch.uint16 and n % 16 == 0 and not hasattr(self, "prefetch"):
            # software prefetch: the raw chunk of the NEXT row is requested from HBM while this row is
            # processed (two 128-bit loads per thread in flight for a whole row time)
            self.prefetch = (pi, n)
            self._e(f"if (16 * tid < {n}) {{ unpack_u16(pf0, {r}); unpack_u16(pf1, {r} + 8); }}",
                    f"else {{ _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {r}[j] = 0.f; }}",
                    f"if (16 * tid < {n} && row + gridDim.x < A.n_rows) {{",
                    f"  const uint16_t* g_ = (const uint16_t*)A.p[{pi}] + (row + gridDim.x) * A.s[{pi}] + 16 * tid;",
                    "  pf0 = ldg_nc(g_); pf1 = ldg_nc(g_ + 8);",
                    "}")
            self.aligned_ptrs = getattr(self, "aligned_ptrs", []) + [pi]
        elif dt == torch.uint16 and n % 16 == 0:
            self._e(f"ldg_chunk_u16((const uint16_t*)A.p[{pi}] + row * A.s[{pi}], tid, {n}, {r});")
            self.aligned_ptrs = getattr(self, "aligned_ptrs", []) + [pi]
        else:
            f = self._t("f")
            self._e(f"int {f} = ldg_chunk_any<{ct}>((const {ct}*)A.p[{pi}] + row * A.s[{pi}], tid, {n}, {r});")
            if dt in (torch.float32, torch.float64):
                si = self._alloc_i(1)
                nf = f"nan{w.name}"
                self._e(f"put_imax(CSI({si}), {f}, lane, warp);")
                self.posts.append(f"const int {nf} = get_imax(CSI({si}), lane);")
                self.pending.add(nf)
                w.nan = nf
        self._store(w, r)

    def _nan_guard(self, flags):
        fl = [f for f in flags if f != "0"]
        return " || ".join(f"({f})" for f in fl) if fl else None

    def _wrange(self, lo, hi):
        """block warps that own a chunk of the samples [lo, hi)"""
        return lo // (32 * CHK), min(NT // 32, -(-hi // (32 * CHK)))

    def _guard(self, w0, w1):
        """(opening, closing, number of warps) of a warp-uniform guard for the warps [w0, w1)"""
        if w0 <= 0 and w1 >= NT // 32:
            return "{", "}", NT // 32
        cond = f"warp < {w1}" if w0 <= 0 else (f"warp >= {w0}" if w1 >= NT // 32 else f"warp >= {w0} && warp < {w1}")
        return f"if ({cond}) {{", "}", w1 - w0

    def _e_min_max(self, nd):
        # min_max.py:11-82 / numpy.amax: value via FMNMX; outputs nobody reads (and that are no chain outputs) are not
        # computed at all.  Only the warps that own part of the range take part; chunks inside the range run
        # unmasked.  The block warps deposit their partial results in the mailbox; the scalar warp combines them.
        # First-occurrence index: on the latency-critical path (a scalar the block stream waits for depends on it)
        # every thread looks up the position of its own extremum before the block-wide result is known; otherwise
        # the look-up happens after the round's barrier, in the one warp that owns the block extremum.
        w, off, n = nd["ins"][0]
        outs = [o if (o and o in self.used_scalars) else None for o in nd["outs"]]
        if not any(outs):
            return
        if nd.get("from_conv"):
            # the convolution already deposited the per-warp maxima in the mailbox
            g = self._nan_guard([self._flag_s(w.nan)])
            self.s_seq += 1
            self._es_later(self._asg(outs[3], f"{'(' + g + ') ? CUDART_NAN_F : ' if g else ''}get_fmax(MBI({nd['from_conv'][1]}), lane)"),
                           *self._stores(outs[3]))
            self._def_s_later(outs[3])
            return
        self._need(w.nan)
        r = self._chunk(w)
        lo, hi = off, min(off + n, w.n)
        t0, t1 = -(-lo // CHK), hi // CHK                  # chunks entirely inside the range
        partial = (lo % CHK != 0) + (hi % CHK != 0)
        w0, w1 = self._wrange(lo, hi)
        g_open, g_close, g_n = self._guard(w0, w1)
        rng = f", {w0}, {w1}" if g_n < NT // 32 else ""
        todo = []
        for (it, iv, fn, put_a, get_a, put_v, get_v, ident) in (
                (0, 2, "min", "put_argmin", "get_argmin", "put_fmin", "get_fmin", "CUDART_INF_F"),
                (1, 3, "max", "put_argmax", "get_argmax", "put_fmax", "get_fmax", "(-CUDART_INF_F)")):
            if outs[it] is None and outs[iv] is None:
                continue
            m = self._t("m")

            def local(f, extra=""):
                full = f"{f}_local<true>({r}, {extra}16 * tid, {lo}, {hi})"
                part = f"{f}_local<false>({r}, {extra}16 * tid, {lo}, {hi})"
                if not partial and t0 <= w0 * 32 and t1 >= min(w1 * 32, NT):     # every thread of the guarded warps is interior
                    return [f"  {{res}} = {full};   //@X {g_n}"]
                return [f"  if (tid >= {t0} && tid < {t1}) {{res}} = {full};   //@X {-(-t1 // 32) - t0 // 32}",
                        f"  else if (16 * tid < {hi} && 16 * tid + 16 > {lo}) {{res}} = {part};   //@X {max(1, partial)}"]

            urgent_idx = outs[it] is not None and (outs[it] in self.urgent_names or outs[it] in self.b_needed
                                                   or os.environ.get("DSPEED_B200_LATE_ARG", "1") == "0")
            whole = not partial and t0 == 0 and t1 * CHK >= w.n and lo == 0
            if (outs[it] is not None and not urgent_idx and w.is_input and whole and w.n == CHK * NT
                    and getattr(w, "int_valued", False)):
                # raw ADC words: value and first position in one pass (no equality look-up)
                mb = self._mbi(2)
                jx = self._t("jx")
                self._e(f"int {jx}; const float {m} = arg{fn}_packed({r}, {jx});",
                        f"{put_a}(MBI({mb}), {m}, 16 * tid + {jx}, lane, warp);")
                todo.append((it, iv, get_a, mb, "arg"))
                continue
            self._e(f"float {m} = {ident};", g_open, *[ln.format(res=m) for ln in local(fn)])
            if outs[it] is not None and (urgent_idx or w.slot is None or not w.needs_slot):
                mb = self._mbi(2)
                ix = self._t("ix")
                self._e(f"  int {ix} = 0x7fffffff;", *[ln.format(res=ix) for ln in local("first_eq", f"{m}, ")],
                        f"  {put_a}(MBI({mb}), {m}, {ix}, lane, warp);   //@X {g_n}", g_close)
                todo.append((it, iv, get_a, mb, "arg"))
            elif outs[it] is not None:
                # the position is looked up after the barrier by the warp that holds the block extremum, which
                # re-reads its chunk from the slot (no register chunk stays live across the barrier)
                mb = self._mbi(2)
                ix, q = self._t("ix"), self._t("q")
                self._e(f"  {put_v}(MBI({mb}), {m}, lane, warp);   //@X {g_n}", g_close,
                        f"if (tid == 0) MBI({mb + 1})[0] = 0x7fffffff;")
                self.posts.append(f"{g_open} const float g_ = {get_v}(MBI({mb}), lane{rng});   //@X {g_n}")
                self.posts.append(f"  if ({m} == g_) {{ int {ix} = 0x7fffffff; float {q}[16]; ld_chunk({self._slot(w)}, tid, {q});   //@X 1")
                self.posts.extend("  " + ln.format(res=ix).replace(r + ",", q + ",").replace("//@X", "//@X 1 //")
                                  for ln in local("first_eq", "g_, "))
                self.posts.append(f"    if ({ix} != 0x7fffffff) atomicMin(&MBI({mb + 1})[0], {ix}); }} {g_close}   //@X 1")
                self.post_mail = True
                todo.append((it, iv, get_v, mb, "late"))
            else:
                mb = self._mbi(1)
                self._e(f"  {put_v}(MBI({mb}), {m}, lane, warp);   //@X {g_n}", g_close)
                todo.append((it, iv, get_v, mb, "val"))
        g = self._nan_guard([self._flag_s(w.nan)])
        gq = f"({g}) ? CUDART_NAN_F : " if g else ""
        self.s_dirty = True
        self.s_seq += 1
        for (it, iv, get, mb, how) in todo:
            if how == "arg":
                v, i = self._t("v"), self._t("i")
                self._es_later(f"float {v}; int {i}; {get}(MBI({mb}), lane, {v}, {i}{rng});", self._asg(outs[it], f"{gq}(float){i}"))
                if outs[iv]:
                    self._es_later(self._asg(outs[iv], f"{gq}{v}"))
            elif how == "late":
                self._es_later(self._asg(outs[it], f"{gq}(float)MBI({mb + 1})[0]"))
                if outs[iv]:
                    self._es_later(self._asg(outs[iv], f"{gq}{get}(MBI({mb}), lane{rng})"))
            else:
                self._es_later(self._asg(outs[iv], f"{gq}{get}(MBI({mb}), lane{rng})"))
        for o in outs:
            if o:
                self._es_later(*self._stores(o))
                self._def_s_later(o)

    def _e_lsf(self, nd):
        # linear_slope_fit.py:11-90.  Threads whose chunk lies inside [off, off + n) run the unmasked float32-local
        # sums, the (at most two) chunks that straddle an end the masked ones, and only the warps that hold part of
        # the range take part in the block sum.
        w, off, n = nd["ins"][0]
        outs = nd["outs"]
        self._need(w.nan)
        r = self._chunk(w)
        mb = self._mbd(3)
        a, b, c = self._t("sy"), self._t("sxy"), self._t("syy")
        lo, hi = off, off + n
        t0, t1 = -(-lo // CHK), hi // CHK           # chunks entirely inside the range
        w0, w1 = lo // (32 * CHK), min(NT // 32, -(-hi // (32 * CHK)))
        self._e(f"double {a} = 0.0, {b} = 0.0, {c} = 0.0;",
                f"if (tid >= {t0} && tid < {t1}) lsf_local_f<true>({r}, 16 * tid, {lo}, {hi}, {a}, {b}, {c});   //@X {-(-t1 // 32) - t0 // 32}",
                f"else if (16 * tid < {hi} && 16 * tid + 16 > {lo}) lsf_local_f<false>({r}, 16 * tid, {lo}, {hi}, {a}, {b}, {c});   //@X {(lo % CHK != 0) + (hi % CHK != 0)}",
                f"if (warp >= {w0} && warp < {w1}) {{ put_sum(MBD({mb}), {a}, lane, warp); put_sum(MBD({mb + 1}), {b}, lane, warp); "
                f"put_sum(MBD({mb + 2}), {c}, lane, warp); }}   //@X {w1 - w0}")
        # A mean the BLOCK stream itself needs next (baseline -> bl_subtract): every block warp finishes that one sum
        # after the round's barrier (16 loads + a warp sum) instead of waiting for the scalar warp's round trip --
        # mailbox event, float64 finishing of all four results, broadcast cell, event back -- with the SM idle.
        if outs[0] and outs[0] in self.b_needed and os.environ.get("DSPEED_B200_BLOCK_MEAN", "1") != "0":
            sd = self._alloc_d(1)
            gb = self._nan_guard([w.nan])
            self._e(f"if (warp >= {w0} && warp < {w1}) put_sum(CSD({sd}), {a}, lane, warp);   //@X {w1 - w0}")
            self.posts.append(self._asg(outs[0], f"{'(' + gb + ') ? CUDART_NAN_F : ' if gb else ''}"
                                                 f"(float)div_by(get_sum_r(CSD({sd}), lane, {w0}, {w1}), (double){n}, 1.0 / (double){n})"))
            self.pending.add(outs[0])
            self.b_needed.discard(outs[0])        # (no publication by the scalar warp)
            self.block_finished = getattr(self, "block_finished", set()) | {outs[0]}
        g = self._nan_guard([self._flag_s(w.nan)])
        self.s_dirty = True
        self.s_seq += 1
        f = [self._t("f") for _ in range(4)]
        post = (f"float {f[0]}, {f[1]}, {f[2]}, {f[3]}; lsf_finish({n}, get_sum_r(MBD({mb}), lane, {w0}, {w1}), "
                f"get_sum_r(MBD({mb + 1}), lane, {w0}, {w1}), get_sum_r(MBD({mb + 2}), lane, {w0}, {w1}), "
                f"{f[0]}, {f[1]}, {f[2]}, {f[3]});")
        if g:
            post += f" if ({g}) {{ {f[0]} = {f[1]} = {f[2]} = {f[3]} = CUDART_NAN_F; }}"
        self._es_later(post)
        for k in range(4):
            if outs[k]:
                self._es_later(self._asg(outs[k], f[k]), *self._stores(outs[k]))
                self._def_s_later(outs[k])
                if outs[k] in getattr(self, "block_finished", ()):
                    self.sdom[outs[k]] = "both"       # the block stream computed its own copy

    def _e_bl_sub(self, nd):
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        self._need_all(nd["b"])
        self._need(w.nan)
        r = self._chunk(w)
        o = self._t("r")
        b = self._t("b")
        self._e(f"const float {b} = (float)({nd['b']});",
                f"float {o}[16];",
                f"_Pragma(\"unroll\") for (int j = 0; j < 16; j++) {o}[j] = {r}[j] - {b};")
        flags = [w.nan] + ([f"({b} != {b})"] if not (nd["b"].startswith(("0x", "-0x")) or nd["b"] in self.never_nan) else [])
        g = self._nan_guard(flags)
        if g:
            nf = f"nan{out.name}"
            self._e(f"const int {nf} = {g};")
            out.nan = nf
        self._store(out, o)

    def _runtime_const(self, literal: str) -> str:
        """A database constant that changes per channel / run without changing the program (the pole-zero time
        constant): passed through the launch arguments (Args.c[]) instead of being baked into the source, so that
        chains that differ only in such values share ONE compiled kernel."""
        self.rt_consts = getattr(self, "rt_consts", [])
        self.rt_consts.append(float.fromhex(literal))
        return f"A.c[{len(self.rt_consts) - 1}]"

    def _e_pole_zero(self, nd):
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        omc_arg = None
        if nd["tau"].startswith(("0x", "-0x")) and os.environ.get("DSPEED_B200_BAKE_CONSTANTS", "0") != "1":
            # 1 - exp(-1 / tau) of the float32 time constant, evaluated once on the host in float64
            tau32 = float(np.float32(float.fromhex(nd["tau"])))
            if tau32 == tau32 and tau32 != 0.0:
                omc_arg = self._runtime_const((-math.expm1(-1.0 / tau32)).hex())
        const_tau_arg = omc_arg is not None
        self._need_all(nd["tau"])
        self._need(w.nan)
        r = self._chunk(w)
        sd = self._alloc_d(1)
        tot, incl, omc = self._t("tot"), self._t("incl"), self._t("omc")
        self._e(f"const double {omc} = {omc_arg};" if omc_arg is not None else
                f"const double {omc} = -expm1(-1.0 / (double)(float)({nd['tau']}));",
                f"const double {tot} = (double)chunk_sum_f({r});",
                f"const double {incl} = put_scan(CSD({sd}), {tot}, lane, warp);")
        o, dum = self._t("r"), self._t("tt")
        self.posts.append(f"float {o}[16]; double {dum}; "
                          f"pz_chunk_f({r}, get_excl(CSD({sd}), {incl}, {tot}, lane, warp, {dum}), {omc}, {o});")
        const_tau = nd["tau"].startswith(("0x", "-0x")) or const_tau_arg
        flags = [w.nan] + ([] if const_tau else [f"((float)({nd['tau']}) != (float)({nd['tau']}))"])
        g = self._nan_guard(flags)
        if g:
            nf = f"nan{out.name}"
            self.posts.append(f"const int {nf} = {g};")
            out.nan = nf
            self.pending.add(nf)
        self.pending.add(out.name)
        self._post_store(out, o)   # the output chunk exists after the barrier

    def _e_dpz(self, nd):
        # pole_zero.py:82-198 as a geometric scan followed by a plain scan (chain_rt.cuh: dpz_local), two rounds
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        self._need(w.nan)
        x = self._chunk(w)
        self._visible(w)
        a_, b_, f = math.exp(-1.0 / nd["tau1"]), math.exp(-1.0 / nd["tau2"]), nd["frac"]
        if not all(math.isfinite(v) for v in (a_, b_, f)):
            raise NotSpecializable("double_pole_zero: NaN constants (the interpreted tier writes the NaN outputs)")
        r = -1.0 * (f * b_ - f * a_ - b_)           # transfer_denom_2; transfer_denom_1 = -(1 + r)
        n1, n2 = -1.0 * (a_ + b_), a_ * b_
        R = r ** CHK
        rp, rw, rl, g, cl, vt, vin, incl, tot, incl2 = (self._t(p) for p in ("rp", "rw", "rl", "g", "cl", "vt", "vin", "incl", "tot", "incl"))
        self.static_arrays = getattr(self, "static_arrays", [])
        self.static_arrays.append(f"__device__ const double {rl}[32] = {{{', '.join(_lit(R ** k) for k in range(32))}}};")
        gs, acc = [], 0.0
        for j in range(CHK):
            acc += r ** (j + 1)
            gs.append(acc)
        sd = self._alloc_d(1)
        sl = self._slot(w)
        self._e(f"const double {rp}[5] = {{{', '.join(_lit(R ** (1 << k)) for k in range(5))}}};",
                f"const double {rw}[5] = {{{', '.join(_lit((R ** 32) ** (1 << k)) for k in range(5))}}};",
                f"const double {g}[16] = {{{', '.join(_lit(v) for v in gs)}}};",
                f"double {cl}[16], {vt}; dpz_local({x}, 16 * tid >= 1 ? at({sl}, 16 * tid - 1) : 0.f, "
                f"16 * tid >= 2 ? at({sl}, 16 * tid - 2) : 0.f, 16 * tid, {n}, {_lit(r)}, {_lit(n1)}, {_lit(n2)}, {cl}, {vt});",
                f"const double {incl} = put_scan_geo(CSD({sd}), {vt}, {rp}, lane, warp);")
        self.posts.append(f"const double {vin} = get_excl_geo(CSD({sd}), {incl}, {rw}, {rl}[lane], lane, warp);")
        self.posts.append(f"const double {tot} = fma({vin}, {g}[15], {cl}[15]);")
        self.pending.add(vin)
        self._close_round()
        sd2 = self._alloc_d(1)
        self._e(f"const double {incl2} = put_scan(CSD({sd2}), {tot}, lane, warp);")
        o, dum, win = self._t("r"), self._t("tt"), self._t("win")
        self.posts.append(f"double {dum}; const double {win} = get_excl(CSD({sd2}), {incl2}, {tot}, lane, warp, {dum});")
        self.posts.append(f"float {o}[16]; _Pragma(\"unroll\") for (int j = 0; j < 16; j++) "
                          f"{o}[j] = (float)({win} + fma({vin}, {g}[j], {cl}[j]));")
        out.nan = w.nan
        self.pending.add(out.name)
        self._post_store(out, o)

    def _e_fir_group(self, nd):
        members = nd["members"]
        w = nd["ins"][0][0]
        n = nd["ins"][0][2]
        self._need(w.nan)
        own = self._chunk(w)          # own chunk in registers (tap offset 0)
        if w.slot is None:
            raise NotSpecializable("internal: FIR input without a slot")
        self._visible(w)
        # batches of at most 3 filters per round (register pressure)
        self._e("PROF_SUB_RESET();" if getattr(self, "_psub", False) else "PROF_SUB_BEGIN();")
        self._psub = True
        for b0 in range(0, len(members), 3):
            batch = members[b0:b0 + 3]
            recs = []
            for mi, m in enumerate(batch):
                out = m["wouts"][0]
                d, tot, incl = self._t("d"), self._t("tot"), self._t("incl")
                self._e(f"float {d}[16];")
                fresh = [True]      # the first term ASSIGNS the accumulators (no 0 + x additions, no zero fill)

                def acc(c, term, ind=""):
                    loop = f"{ind}_Pragma(\"unroll\") for (int j = 0; j < 16; j++) "
                    if fresh[0]:
                        fresh[0] = False
                        rhs = term if c == 1.0 else f"-{term}" if c == -1.0 else f"{_flit(c)} * {term}"
                        return f"{loop}{d}[j] = {rhs};"
                    if c == 1.0:
                        return f"{loop}{d}[j] += {term};"
                    if c == -1.0:
                        return f"{loop}{d}[j] -= {term};"
                    return f"{loop}{d}[j] = fmaf({_flit(c)}, {term}, {d}[j]);"
                # Taps whose offsets lie within 16 samples of each other share ONE span load
                # (16 + width samples, 128-bit loads): the filters are shared-memory-bandwidth
                # bound, so a cluster costs (16 + width) / 4 loads instead of 5 per tap.  Inside a
                # cluster, consecutive taps whose coefficients agree to float32 rounding of the
                # kernel (a linear ramp's first difference) become one sliding-window sum.
                taps = sorted(m["taps"])
                zc = self._zc(w)
                tol = 4e-7 * max(abs(x[1]) for x in taps)
                k = 0
                while k < len(taps):
                    ts0 = taps[k][0]
                    e = k
                    while e + 1 < len(taps) and taps[e + 1][0] - ts0 <= 16:
                        e += 1
                    cluster = taps[k:e + 1]
                    k = e + 1
                    if len(cluster) == 1 and ts0 == 0:
                        self._e(acc(cluster[0][1], f"{own}[j]"))
                        continue
                    width = cluster[-1][0] - ts0
                    cnt = 16 + width
                    v = self._t("v")
                    # v[q] = x[16 t + q - (ts0 + width)]  ->  tap ts reads v[j + (ts0 + width - ts)]
                    self._e(f"{{ float {v}[{cnt}]; ld_span<{-(ts0 + width)}, {cnt}>({self._slot(w)}, tid, {n}, {zc}, {v});")
                    q = 0
                    while q < len(cluster):
                        ts, c = cluster[q]
                        run = 1
                        while (q + run < len(cluster) and cluster[q + run][0] == ts + run
                               and abs(cluster[q + run][1] - c) <= tol):
                            run += 1
                        if run >= 3:
                            cm = sum(x[1] for x in cluster[q:q + run]) / run
                            lo = ts0 + width - (ts + run - 1)      # v index of the oldest sample of the window at j = 0
                            ws = self._t("ws")
                            upd = f"{d}[j] = {_flit(cm)} * {ws};" if fresh[0] else f"{d}[j] = fmaf({_flit(cm)}, {ws}, {d}[j]);"
                            fresh[0] = False
                            self._e(f"  float {ws} = {v}[{lo}]; _Pragma(\"unroll\") for (int u = 1; u < {run}; u++) {ws} += {v}[{lo} + u];",
                                    f"  _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {{ {upd} "
                                    f"if (j < 15) {ws} += {v}[{lo} + j + {run}] - {v}[{lo} + j]; }}")
                            q += run
                            continue
                        off = ts0 + width - ts
                        self._e(acc(c, f"{v}[j + {off}]", "  "))
                        q += 1
                    self._e("}")
                sd = self._alloc_d(1)
                self._e(f"const float {tot} = cumsum_local({d});",
                        f"const float {incl} = put_scan_{_FF}(CSD({sd}), {tot}, lane, warp);")
                ex = None
                if m["extra"]:
                    kx = self._t("kx")
                    ne = len(m["extra"])
                    self.static_arrays = getattr(self, "static_arrays", [])
                    self.static_arrays.append(f"__device__ const float {kx}[{ne}] = {{{', '.join(_flit(v) for v in m['extra'])}}};")
                    sx = self._alloc_d(1)
                    term = self._t("ex")
                    self._e(f"double {term} = 0.0;",
                            f"for (int q = tid; q < {ne}; q += 512) {term} += (double)({kx}[q] * at({self._slot(w)}, {ne - 1} - q));",
                            f"put_sum(CSD({sx}), {term}, lane, warp);")
                    ex = sx
                self._e(f"PROF_SUB({8 + b0 + mi});")
                recs.append((m, d, tot, incl, sd, ex))
            for (m, d, tot, incl, sd, ex) in recs:
                out = m["wouts"][0]
                o, off, dum = self._t("r"), self._t("off"), self._t("tt")
                sc = _flit(m["scale"])
                extra = f" + get_sum(CSD({ex}), lane)" if ex is not None else ""
                self.posts.append(
                    f"const float {off} = (float)((get_excl_{_FF}(CSD({sd}), {incl}, {tot}, lane, warp){extra}) * (double){sc});"
                    if ex is not None or _FF == "f" else
                    f"const float {off} = get_excl_ff(CSD({sd}), {incl}, {tot}, lane, warp) * {sc};")
                self.posts.append(f"float {o}[16]; _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {o}[j] = fmaf({d}[j], {sc}, {off});")
                out.nan = w.nan
                self.pending.add(out.name)
                self._post_store(out, o)
            self._close_round()

    def _e_tpt(self, nd):
        w, off, n = nd["ins"][0]
        self._s_wave(w)
        f = self._t("f")
        g = self._nan_guard([self._flag_s(w.nan)])
        if getattr(w, "scatter", False):
            raise NotSpecializable("threshold search on a scattered waveform")
        if not self.summ[w.id][1]:
            self._es(f"//@R {1000 + self.summ[w.id][0]} 0 1 {self.s_seq}")
        call = (f"tpt_w({self._slot(w)}, {self._summ(w)}, {n}, (float)({nd['thr']}), (float)({nd['start']}), "
                f"(float)({nd['walk']}), {f}, lane)")
        self._es(f"int {f} = 0;",
                 self._asg(nd['out'], f"{'(' + g + ') ? CUDART_NAN_F : ' if g else ''}{call}"),
                 f"if ({f} && lane == 0) raise_fatal(A.fatal ? A.fatal + 4 * {nd['fatal']} : nullptr, {f}, A.row0 + row);",
                 *self._stores(nd["out"]), urgent=nd["out"] in self.b_needed)
        self._def_s(nd["out"])

    def _e_tpt_chain(self, nd):
        w, off, n = nd["ins"][0]
        self._s_wave(w)
        f, th, res = self._t("f"), self._t("th"), self._t("tc")
        g = self._nan_guard([self._flag_s(w.nan)])
        if getattr(w, "scatter", False):
            raise NotSpecializable("threshold search on a scattered waveform")
        if not self.summ[w.id][1]:
            self._es(f"//@R {1000 + self.summ[w.id][0]} 0 1 {self.s_seq}")
        K = len(nd["members"])
        start = f"(float)({nd['start']})" if not g else f"(({g}) ? CUDART_NAN_F : (float)({nd['start']}))"
        lines = [f"int {f} = 0; float {res}[{K}];",
                 f"{{ const float {th}[{K}] = {{{', '.join('(float)(' + t + ')' for t in nd['thrs'])}}};",
                 f"  tpt_chain_bwd<{K}>({self._slot(w)}, {self._summ(w)}, {n}, {th}, {start}, {res}, {f}, lane); }}",
                 f"if ({f} && lane == 0) raise_fatal(A.fatal ? A.fatal + 4 * {nd['fatal']} : nullptr, {f}, A.row0 + row);"]
        for j, m in enumerate(nd["members"]):
            lines.append(self._asg(m["out"], f"{res}[{j}]"))
            lines.extend(self._stores(m["out"]))
        self._es(*lines, urgent=any(m["out"] in self.b_needed for m in nd["members"]))
        for m in nd["members"]:
            self._def_s(m["out"])

    def _e_trap_pickoff(self, nd):
        # trap_filters.py:230-301 : the normalised trapezoid at ONE index from two window sums,
        # evaluated by the scalar warp (NaN when the windows do not fit, fatal on a fractional index)
        w, off, n = nd["ins"][0]
        self._s_wave(w)
        g = self._nan_guard([self._flag_s(w.nan)])
        r_, fl = nd["rise"], nd["flat"]
        t, st, res, f = (self._t(x) for x in ("t", "st", "res", "f"))
        sl = self._slot(w)
        self._es(f"int {f} = 0; float {res} = CUDART_NAN_F;",
                 f"{{ const float {t} = (float)({nd['t']});",
                 f"  if ({t} == {t}{' && !(' + g + ')' if g else ''}) {{",
                 f"    if (floorf({t}) != {t}) {f} = DSPB_FATAL_PICKOFF_NONINT;",
                 f"    else {{ const long long {st} = (long long)({t} + 1.0f);",
                 f"      if ({st} <= {n} && {st} >= {2 * r_ + fl})",
                 f"        {res} = (wrange_sum({sl}, {n}, (int){st} - {r_}, (int){st}, lane) - "
                 f"wrange_sum({sl}, {n}, (int){st} - {2 * r_ + fl}, (int){st} - {r_ + fl}, lane)) / {float(r_)}f; }}",
                 "  } }",
                 self._asg(nd['out'], res),
                 f"if ({f} && lane == 0) raise_fatal(A.fatal ? A.fatal + 4 * {nd['fatal']} : nullptr, {f}, A.row0 + row);",
                 *self._stores(nd["out"]))
        self._def_s(nd["out"])

    def _e_fir_lazy(self, nd):
        pass   # evaluated at its pick-off positions (see _e_ftp)

    def _e_ftp_lazy(self, nd):
        fir = nd["lazy"]
        w, off, n = nd["ins"][0]
        self._s_wave(w)
        g = self._nan_guard([self._flag_s(w.nan)])
        taps = sorted(fir["taps"])
        sc = _flit(fir["scale"])
        t, i, v0, v1, res, f = (self._t(x) for x in ("t", "i", "v", "v", "res", "f"))
        sl = self._slot(w)
        # out[i] = scale * sum_k C_k * sum_{m in (i - t_{k+1}, i - t_k]} y[m],  C_k = c_0 + ... + c_k
        terms, cum = [], 0.0
        for k, (ts, c) in enumerate(taps):
            cum += c
            if abs(cum) > 1e-12 and k + 1 < len(taps):
                terms.append(f"{_flit(cum)} * wrange_sum({sl}, {n}, {i} - {taps[k + 1][0]} + 1, {i} - {ts} + 1, lane)")
        step = " + ".join(f"{_flit(c)} * at0({sl}, {n}, {i} + 1 - {ts})" for ts, c in taps)
        mode = chr(nd["mode"])
        if mode == "i":
            frac = f"{f} = DSPB_FATAL_FTP_INT;"
        elif mode == "f":
            frac = f"{res} = {v0};"
        else:
            frac = f"const float {v1} = {v0} + {sc} * ({step}); "
            if mode == "c":
                frac += f"{res} = {v1};"
            elif mode == "n":
                frac += f"{res} = ({t} - (float){i} < 0.5f) ? {v0} : {v1};"
            else:
                frac += (f"const double t0_ = (double){t} - (double){i}; "
                         f"{res} = (float)((1.0 - t0_) * (double){v0} + t0_ * (double){v1});")
        self._es(f"int {f} = 0; float {res} = CUDART_NAN_F;",
                 f"{{ const float {t} = (float)({nd['t']});",
                 f"  if ({t} == {t} && {t} >= 0.f && {t} <= {float(n - 1)}f{' && !(' + g + ')' if g else ''}) {{",
                 f"    const int {i} = (int){t};",
                 f"    const float {v0} = {sc} * ({' + '.join(terms) if terms else '0.f'});",
                 f"    if ((float){i} == {t}) {res} = {v0}; else {{ {frac} }}",
                 "  } }",
                 self._asg(nd['out'], res),
                 f"if ({f} && lane == 0) raise_fatal(A.fatal ? A.fatal + 4 * {nd['fatal']} : nullptr, {f}, A.row0 + row);",
                 *self._stores(nd["out"]))
        self._def_s(nd["out"])

    def _e_ftp(self, nd):
        if nd.get("lazy") is not None:
            return self._e_ftp_lazy(nd)
        if nd.get("from_conv"):
            w = nd["ins"][0][0]
            g = self._nan_guard([self._flag_s(w.nan)])
            self.s_seq += 1
            self._es_later(self._asg(nd["out"], f"{'(' + g + ') ? CUDART_NAN_F : ' if g else ''}(float)MBD({nd['from_conv'][1]})[0]"),
                           *self._stores(nd["out"]))
            self._def_s_later(nd["out"])
            return
        w, off, n = nd["ins"][0]
        self._s_wave(w)
        f = self._t("f")
        g = self._nan_guard([self._flag_s(w.nan)])
        call = f"op_fixed_time_pickoff<float>({self._slot(w)}, {n}, (float)({nd['t']}), {nd['mode']}, {f})"
        self._es(f"int {f} = 0;",
                 self._asg(nd['out'], f"{'(' + g + ') ? CUDART_NAN_F : ' if g else ''}{call}"),
                 f"if ({f} && lane == 0) raise_fatal(A.fatal ? A.fatal + 4 * {nd['fatal']} : nullptr, {f}, A.row0 + row);",
                 *self._stores(nd["out"]))
        self._def_s(nd["out"])

    def _e_windower(self, nd):
        # windower.py:12-54 : out[k] = in[t0 + k], NaN where the window leaves the waveform; one
        # sample per thread (the window start is data dependent)
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        m = out.n
        mc = (m + CHK - 1) // CHK * CHK
        self._need_all(nd["t0"])
        self._need(w.nan)
        self._visible(w)
        out.needs_slot = True
        self._give_slot(out)
        pad, beg, tf = self._t("pad"), self._t("beg"), self._t("tf")
        si = self._alloc_i(1)
        g = self._nan_guard([w.nan, f"({tf} != {tf})"])
        so = self._slot(out)
        self._e(self._wmark(out.slot), f"int {pad} = 0;",
                f"const float {tf} = (float)({nd['t0']});",
                f"int {beg} = ({tf} == {tf}) ? (int)fminf(fmaxf({tf}, -1.0e9f), 1.0e9f) : 0; if ({beg} > {n}) {beg} = {n};",
                f"for (int k = tid; k < {mc}; k += {32 * (self.region_nw or NT // 32)}) {{ const int q = {beg} + k; float v = 0.f; "
                f"if (k < {m}) {{ if (!({g}) && q >= 0 && q < {n}) v = at({self._slot(w)}, q); else {{ {pad} = 1; v = CUDART_NAN_F; }} }} "
                f"{so}[sidx(k)] = v; }}",
                f"put_imax(CSI({si}), {pad}, lane, warp);")
        nf = f"nan{out.name}"
        self.posts.append(f"const int {nf} = get_imax(CSI({si}), lane{self._nwarg()});")
        self.pending.add(nf)
        out.nan = nf
        out.nan_elementwise = True
        out.scatter = True
        out.reg = None
        self.dirty.add(out.slot[0])
        self.s_dirty = True

    def _e_avg_current(self, nd):
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        L = nd["L"]
        self._need(w.nan)
        self._visible(w)
        a = self._chunk(w)
        b, o = self._t("r"), self._t("r")
        self._e(f"float {b}[16]; ld_shift<{L}>({self._slot(w)}, tid, {n}, {self._zc(w)}, {b});",
                f"float {o}[16]; _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {o}[j] = ({b}[j] - {a}[j]) / {_flit(L)};")
        out.nan = w.nan
        self._store(out, o)

    def _e_upsampler(self, nd):
        # upsampler.py:14-49 with an integer factor: output t takes input floor((t + half) / up);
        # outputs beyond the last input's span stay NaN
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        up, m = nd["up"], out.n
        half = up // 2
        self._need(w.nan)
        self._visible(w)
        o = self._t("r")
        holes = m > n * up - half
        g = self._nan_guard([w.nan])
        ok = f"{' && !(' + g + ')' if g else ''}"
        if up >= CHK and up % CHK == 0:
            # a chunk of 16 outputs sees at most two inputs: q0 = floor((16 t + half) / up) and q0 + 1
            q0, v0, v1, sw = self._t("q"), self._t("v"), self._t("v"), self._t("sw")
            self._e(f"float {o}[16];",
                    f"const int {q0} = (16 * tid + {half}) / {up};",
                    f"const int {sw} = ({q0} + 1) * {up} - {half} - 16 * tid;   // first j that belongs to input q0 + 1",
                    f"const float {v0} = ({q0} < {n}{ok}) ? at({self._slot(w)}, {q0}) : CUDART_NAN_F;",
                    f"const float {v1} = ({q0} + 1 < {n}{ok}) ? at({self._slot(w)}, {q0} + 1) : CUDART_NAN_F;",
                    f"_Pragma(\"unroll\") for (int j = 0; j < 16; j++) {o}[j] = 16 * tid + j >= {m} ? 0.f : (j < {sw} ? {v0} : {v1});")
        else:
            self._e(f"float {o}[16];",
                    f"_Pragma(\"unroll\") for (int j = 0; j < 16; j++) {{ const int t = 16 * tid + j; const int q = (t + {half}) / {up}; "
                    f"{o}[j] = t >= {m} ? 0.f : ((q < {n}{ok}) ? at({self._slot(w)}, q) : CUDART_NAN_F); }}")
        if holes:
            nf = f"nan{out.name}"
            self._e(f"const int {nf} = 1;")
            out.nan = nf
        else:
            out.nan = w.nan
        out.nan_elementwise = True
        self._store(out, o)

    def _e_mw(self, nd):
        # moving_windows.py:12-203 : successive boxcar means with edge-value padding, each one a
        # chunk-local running sum of (x[i] - x[i -/+ L]) / L plus one block scan.  Only the warps that own a
        # chunk of the wave take part, and chunks away from the ends run the unmasked difference.
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        L = nd["L"]
        il = _flit(1.0 / float(np.float32(L)))
        dirs = nd["dirs"]
        temps = []
        src = w
        nwarg = self._nwarg()
        g_n = self.region_nw or NT // 32
        for k, dr in enumerate(dirs):
            last = k == len(dirs) - 1
            if last:
                dst = out
            else:
                if len(temps) < 2:
                    tw = Wave(10000 + 10 * self.pos + k, n)
                    tw.needs_slot = True
                    tw.last = self.pos
                    temps.append(tw)
                dst = temps[k % 2]
            self._need(src.nan)
            self._visible(src)
            x = self._chunk(src)
            sh, d, tot, incl, e0, tt = (self._t(p) for p in ("r", "d", "tot", "incl", "e", "tt"))
            sd = self._alloc_d(1)
            sl = self._slot(src)
            self._e(f"float {sh}[16], {d}[16];")
            if dr == "l":
                # out[0] = x[0]; out[i] = out[i-1] + (x[i] - x[max(i-L,0)]) / L
                ta, tb = -(-L // CHK), n // CHK     # chunks with every i in [L, n)
                self._e(f"  ld_shift<{-L}>({sl}, tid, {n}, {self._zc(src)}, {sh}); const float {e0} = at({sl}, 0);   //@X {g_n}",
                        f"  if (tid >= {ta} && tid < {tb}) {{ _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {d}[j] = ({x}[j] - {sh}[j]) * {il}; }}   //@X {g_n}",
                        f"  else {{ _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {{ const int i = 16 * tid + j; "
                        f"{d}[j] = i >= {n} ? 0.f : (i == 0 ? {e0} : ({x}[j] - (i >= {L} ? {sh}[j] : {e0})) * {il}); }} }}   //@X 2",
                        f"const float {tot} = cumsum_local({d});   //@X {g_n}",
                        f"const float {incl} = put_scan_{_FF}(CSD({sd}), {tot}, lane, warp);   //@X {g_n}")
                get = f"get_excl_{_FF}(CSD({sd}), {incl}, {tot}, lane, warp{nwarg})"
            else:
                # mirror image: out[n-1] = x[n-1]; out[i] = out[i+1] + (x[i] - x[min(i+L,n-1)]) / L
                ta, tb = 0, max(0, (n - L) // CHK)  # chunks with every i + L <= n - 1
                self._e(f"  ld_shift<{L}>({sl}, tid, {n}, {self._zc(src)}, {sh}); const float {e0} = at({sl}, {n - 1});   //@X {g_n}",
                        f"  if (tid >= {ta} && tid < {tb}) {{ _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {d}[j] = ({x}[j] - {sh}[j]) * {il}; }}   //@X {g_n}",
                        f"  else {{ _Pragma(\"unroll\") for (int j = 0; j < 16; j++) {{ const int i = 16 * tid + j; "
                        f"{d}[j] = i >= {n} ? 0.f : (i == {n - 1} ? {e0} : ({x}[j] - (i + {L} <= {n - 1} ? {sh}[j] : {e0})) * {il}); }} }}   //@X 2",
                        f"const float {tot} = cumsum_local_rev({d});   //@X {g_n}",
                        f"const float {incl} = put_scan_rev_{_FF}(CSD({sd}), {tot}, lane, warp);   //@X {g_n}")
                get = f"get_excl_rev_{_FF}(CSD({sd}), {incl}, {tot}, lane, warp{nwarg})"
            o, offv = self._t("r"), self._t("off")
            self.posts.append(f"float {o}[16]; const float {offv} = (float){get}; _Pragma(\"unroll\") for (int j = 0; j < 16; j++) "
                              f"{o}[j] = {d}[j] + {offv};   //@X {g_n}")
            dst.nan = src.nan
            dst.reg = None
            self.pending.add(dst.name)
            self._post_store(dst, o)
            self._close_round()
            src = dst
        for tw in temps:
            if tw.slot is not None:
                self._slot_free(tw.slot)
            if tw in self.live_regs:
                self.live_regs.remove(tw)

    def _e_conv_seg(self, nd):
        raise NotSpecializable("cusp/zac convolution over a full-length waveform (interpreted path)")
        w, off, n = nd["ins"][0]
        out = nd["wouts"][0]
        self._need(w.nan)
        self._visible(w)
        self._close_round()
        scratch = self._slot_alloc()
        if scratch[0] in self.xread or scratch[0] in self.dirty:
            self._barrier()
        self._give_slot(out)
        seg = nd["seg"]
        prm = self._t("prm")
        self.static_arrays = getattr(self, "static_arrays", [])
        self.static_arrays.append(f"__device__ const double {prm}[10] = {{{', '.join(_lit(v) for v in seg)}}};")
        self._e(f"op_conv_seg<float>({self._slot(w)}, {n}, {self._slot(out)}, reinterpret_cast<double*>(SLOT({scratch[0]})), {prm}, sc);")
        self._slot_free(scratch)
        out.nan = w.nan
        out.reg = None
        self.dirty.discard(out.slot[0])

    def _conv_sink_plan(self, m, p):
        """(amax nodes, pick-off nodes) when the convolution output needs no slot, else None"""
        out = m["wouts"][0]
        amax, picks = [], []
        for u in out.uses:
            un = self.nodes[u]
            if un["kind"] == "min_max":
                used = [o if (o and o in self.used_scalars) else None for o in un["outs"]]
                if any(used[:3]) or un["ins"][0][1] != 0 or un["ins"][0][2] != out.n:
                    return None
                amax.append(un)
            elif un["kind"] == "ftp" and un["t"].startswith(("0x", "-0x")) and un.get("lazy") is None:
                t = float.fromhex(un["t"])
                if t != int(t) or not (0 <= t < p) or picks:
                    return None
                picks.append(un)
            else:
                return None
        if p > NT:
            return None
        return amax, picks

    def _e_conv_seg_group(self, nd):
        members = nd["members"]
        w, off, n = nd["ins"][0]
        self._need(w.nan)
        x = self._chunk(w)
        self._visible(w)
        self._close_round()
        two = len(members) == 2
        seg = members[0]["seg"]
        sigma, lt, fl, L, c, inv2S = seg[:6]
        p = n - int(L) + 1
        poly = any(m["seg"][7] != 0.0 for m in members)
        nq = 5 if poly else 3
        cw = (((p + CHK - 1) // CHK) + 1) | 1
        need = nq * (4 * CHK * cw + NT + 32) * 8
        nsl = -(-need // (self.slot_words * 4))
        scratch = self._slot_alloc_adjacent(nsl)
        outs, sinks = [], []
        for m in members:
            out = m["wouts"][0]
            plan = self._conv_sink_plan(m, p)
            if plan is None:
                self._give_slot(out, post=True)
                outs.append(out)
                sinks.append(f"SegSink{{{self._slot(out)}, nullptr, nullptr, 0}}")
                continue
            amax, picks = plan
            mb = self._mbi(1) if amax else None
            pk = self._mbd(1) if picks else None
            sinks.append(f"SegSink{{nullptr, {'MBI(%d)' % mb if amax else 'nullptr'}, "
                         f"{'MBD(%d)' % pk if picks else 'nullptr'}, {int(float.fromhex(picks[0]['t'])) if picks else 0}}}")
            for un in amax:
                un["from_conv"] = ("max", mb)
            for un in picks:
                un["from_conv"] = ("pick", pk)
            out.needs_slot = False
            out.virtual = True
        if not two:
            sinks.append(sinks[0])
        busy = self.xread | self.dirty
        if any(sl[0] in busy for sl in scratch) or any(o.slot[0] in busy for o in outs):
            self._barrier()
        so = [f"SegOut{{{_lit(m['seg'][6])}, {_lit(m['seg'][7])}, {_lit(m['seg'][8])}}}" for m in members]
        if not two:
            so.append(so[0])
        self._e(*[self._wmark(sl) for sl in scratch], *[self._wmark(o.slot) for o in outs])
        pw = self._t("pw")
        self.static_arrays = getattr(self, "static_arrays", [])
        vals = [math.exp(o / sigma) for o in range(p)] + [math.exp(-o / sigma) for o in range(p)]
        self.static_arrays.append(f"__device__ const double {pw}[{2 * p}] = {{{', '.join(_lit(v) for v in vals)}}};")
        # pass-1 variant: when the threads beyond the end of the input can take all (band, chunk) pairs, the chunk
        # sums are float32 inside a chunk / float64 across chunks and the band deposits move to those helper threads
        nch = -(-n // CHK)
        hb = -(-nch // 32) * 32
        bases = [int(L), int(L - lt), int(L - 1 - lt - fl), 0]
        pairs = sum(((bb + p - 1) >> 4) - (bb >> 4) + 1 for bb in bases)
        helpers = os.environ.get("DSPEED_B200_CONV_HELPERS", "1") != "0" and pairs <= NT - hb and 16.0 * nch / sigma < 600.0
        tmpl = f"{'true' if poly else 'false'}, {'true' if two else 'false'}"
        # c = exp(-1 / decay) is the only number of the model that comes from the (per-channel) decay constant: a launch
        # argument, so that channels share the compiled kernel
        c_arg = _lit(c) if os.environ.get("DSPEED_B200_BAKE_CONSTANTS", "0") == "1" else self._runtime_const(float(c).hex())
        args_a = (f"{self._slot(w)}, {x}, {n}, {_lit(sigma)}, {int(lt)}, {int(fl)}, {int(L)}, {c_arg}, {_lit(inv2S)}, "
                  f"{_lit(math.exp(-1.0 / sigma))}, {_lit(math.exp(1.0 / sigma))}, {_lit(math.exp((L - 1) / sigma))}, {pw}")
        args_b = (f"{so[0]}, {so[1]}, {sinks[0]}, {sinks[1]}, reinterpret_cast<double*>(SLOT({scratch[0][0]})), "
                  f"tid, lane, warp);")
        if helpers:
            pwc, wm, wp = self._t("pwc"), self._t("wm"), self._t("wp")
            tab = [math.exp(-16.0 * t / sigma) if t <= nch else 0.0 for t in range(NT)] + \
                  [math.exp(16.0 * t / sigma) if t <= nch else 0.0 for t in range(NT)]
            self.static_arrays.append(f"__device__ const double {pwc}[{2 * NT}] = {{{', '.join(_lit(v) for v in tab)}}};")
            self._e(f"const float {wm}[16] = {{{', '.join(_flit(math.exp(-k / sigma)) for k in range(16))}}};",
                    f"const float {wp}[16] = {{{', '.join(_flit(math.exp(k / sigma)) for k in range(16))}}};",
                    f"conv_seg_chunked_h<{tmpl}, {hb}, {NT}>({args_a}, {pwc}, {wm}, {wp}, {args_b}")
        else:
            self._e(f"conv_seg_chunked<{tmpl}>({args_a}, {args_b}")
        # the band table overwrote the (always-zero) pad columns of its slots
        self._e(f"zero_pads(SLOT({scratch[0][0]}), {self.slot_words}, {self.nchunks}, 0, {len(scratch)}, tid);")
        for sl in scratch:
            self.dirty.add(sl[0])
            self._slot_free(sl)
        self.s_dirty = True
        for m in members:
            m["wouts"][0].nan = w.nan
            m["wouts"][0].reg = None

    def _sc_emit(self, nd, expr, *operands):
        """scalar glue runs in the scalar stream; when none of its operands lives there (inputs,
        constants) it is simply evaluated by both streams"""
        out = nd["out"]
        if any(self._is_s(o) for o in operands):
            self.s_seq += 1
            self._es(self._asg(out, expr), *self._stores(out))
            self._def_s(out)
        else:
            self._e(self._asg(out, expr))
            self._es(self._asg(out, expr), *self._stores(out))

    def _e_sc_bin(self, nd):
        op = {"add": "+", "subtract": "-", "multiply": "*", "divide": "/"}.get(nd["op"])
        if nd["f32"]:
            ex = f"(float)({nd['x']}) {op} (float)({nd['y']})" if op else f"floorf((float)({nd['x']}) / (float)({nd['y']}))"
        else:
            ex = (f"(double)({nd['x']}) {op} (double)({nd['y']})" if op
                  else f"floor((double)({nd['x']}) / (double)({nd['y']}))")
        self._sc_emit(nd, ex, nd["x"], nd["y"])

    def _e_sc_neg(self, nd):
        self._sc_emit(nd, f"-({nd['x']})", nd["x"])

    def _e_sc_convert(self, nd):
        ex = f"((double)({nd['x']}) + (double)({nd['oi']})) * {_lit(nd['ratio'])} - (double)({nd['oo']})"
        fn = {None: "", "round": "rint", "floor": "floor", "ceil": "ceil", "trunc": "trunc"}[nd["mode"]]
        ex = f"{fn}({ex})"
        if nd["f32"]:
            ex = f"(double)(float)({ex})"
        self._sc_emit(nd, ex, nd["x"], nd["oi"], nd["oo"])

    def _e_store_wave(self, nd):
        w, off, n = nd["ins"][0]
        pi = nd["ptr"]
        self._need(w.nan)
        g = self._nan_guard([w.nan])
        dst = f"(float*)A.p[{pi}] + row * A.s[{pi}]"
        if off == 0:
            r = self._chunk(w)
            body = f"stg_chunk({dst}, tid, {n}, {r});"
        else:
            self._visible(w)
            body = f"for (int q = tid; q < {n}; q += 512) ({dst})[q] = at({self._slot(w)}, {off} + q);"
        if g and not getattr(w, "nan_elementwise", False):
            self._e(f"if ({g}) store_row_nan<float>({dst}, {n}); else {{ {body} }}")
        else:
            self._e(body)

    # ------------------------------------------------------------------------------------
    # source, build, launch
    # ------------------------------------------------------------------------------------
    def source(self) -> str:
        np_ = max(1, len(self.ptrs))
        ind = "\n        "
        body_b = ind.join(self._expand_regions(self.LB))
        ns = self.n_swarps
        nthr = NT + 32 * ns
        occ = getattr(self, "occ", 1)
        body_s = ind.join(self.SW[1])
        body_s2 = ind.join(self.SW[2]) if ns == 2 else ""
        # per-event input scalars (baseline, t0 ...) are requested one row ahead, like the raw chunk: the block stream
        # needs them a few hundred instructions into the row, an HBM latency it would otherwise wait out every row
        prolog, scalar_prefetch = [], []
        for k, (name, ct, pi) in enumerate(self.prolog):
            ld = f"((const {ct}*)A.p[{pi}])[%s * A.s[{pi}]]"
            if not any(nd.get(key) == name for nd in self.nodes for key in ("b", "tau", "t0")):
                # only the scalar stream reads it (late in the row: the load has long completed): no extra registers
                prolog.append(self._asg(name, ld % "row"))
                continue
            scalar_prefetch.append(f"{ct} pin{k} = 0; if (blockIdx.x < A.n_rows) pin{k} = {ld % '(long long)blockIdx.x'};")
            prolog.append(self._asg(name, f"pin{k}"))
            prolog.append(f"if (row + gridDim.x < A.n_rows) pin{k} = {ld % '(row + gridDim.x)'};")
        prolog = "\n      ".join(prolog)
        scalar_prefetch = "\n  ".join(scalar_prefetch)
        names = sorted(set(self.svar.values()), key=lambda x: int(x[1:]))
        decl = " ".join(f"{ty} " + ", ".join(n for n in names if self.stype[n] == ty) + ";"
                        for ty in ("float", "double") if any(self.stype[n] == ty for n in names))
        arrays = "\n".join(getattr(self, "static_arrays", []))
        aligned = getattr(self, "aligned_ptrs", [])
        align_check = "".join(f"  if (((uintptr_t)ptrs[{i}] & 15) || (strides[{i}] & 7)) return DSPB_ERR_UNSUPPORTED;\n" for i in aligned)
        prefetch_init = ""
        if hasattr(self, "prefetch"):
            pi, n = self.prefetch
            prefetch_init = (f"uint4 pf0 = make_uint4(0, 0, 0, 0), pf1 = pf0;   // raw chunk of the next row (prefetched)\n"
                             f"  if (!scalar_warp && 16 * tid < {n} && blockIdx.x < A.n_rows) {{\n"
                             f"    const uint16_t* g_ = (const uint16_t*)A.p[{pi}] + (long long)blockIdx.x * A.s[{pi}] + 16 * tid;\n"
                             f"    pf0 = ldg_nc(g_); pf1 = ldg_nc(g_ + 8);\n  }}")
        mbd_off = 2048 + 8192 + 2048
        mbi_off = mbd_off + self.n_mbd * 16 * 8
        summ_off = mbd_off + 2 * self.MB_BUDGET
        return f"""// generated by dspeed_b200/codegen.py -- do not edit
#define DSPB_PSP {self.psp}
#define DSPB_PROF_OFF (2048 + 8192)
// the 16 block warps synchronise on named barrier 1; the scalar warps (warp 16 ...) never join it
#define BSYNC() asm volatile("bar.sync 1, 512;" ::: "memory")
// events between the block stream and the scalar warps: counted arrivals on named barriers.  Barrier 0 (scalar warp 1
// -> block stream) involves the 512 block threads and ONE scalar warp, the block -> scalar events and the progress /
// done events (14 / 15) all {ns} scalar warp(s), the hand-over between the scalar warps (EVX) only those two.
#define EV_COUNT(id) ((id) == 0 ? 544 : ((id) == 12 || (id) == 13) ? 64 : {nthr})
#define EV_ARRIVE(id) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(EV_COUNT(id)) : "memory")
#define EV_WAIT(id) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(EV_COUNT(id)) : "memory")
#include "chain_rt.cuh"
using namespace dspb;
using namespace crt;
namespace {{
constexpr int NP = {np_};
constexpr int N_NODES = {len(self.order) + 1};
constexpr int NC = {max(1, len(getattr(self, "rt_consts", [])))};   // run-time scalar constants (database values)
struct Args {{
  const void* p[NP];
  long long s[NP];
  long long n_rows, row0;
  int* fatal;
  long long* prof;
  double c[NC];
}};
{arrays}
// logical slot k of this row: physical slot k (even rows) or (k + {self.total_slots // 2}) % {self.total_slots} (odd rows)
#define SLOT(k) (slots + (rp ? (((k) + {self.total_slots // 2}) % {self.total_slots}) * {self.slot_words} : (k) * {self.slot_words}))
#define CSD(k) (cs->d[par][k])
#define CSI(k) (cs->i[par][k])
#define MBD(k) (mbd + 16 * (k))
#define MBI(k) (mbi + 16 * (k))
#define SUMM(k) (summ + (k))
#ifdef DSPB_PROFILE   // tracing build (SpecChain.profile): per-node SM-cycle stamps of CTA 0
// per-node SM cycles of both streams of CTA 0, accumulated in shared memory (time between consecutive
// marks of a stream, waits included), flushed when the CTA exits; rows overlap as in production
#define PROF_MARK(k) if (A.prof && tid == 0 && (k) < 128) {{ const long long t_ = clock64(); prof_ts[k] += t_ - prof_prev; prof_prev = t_; }}
#define PROF_MARK_SX(k) if (A.prof && lane == 0 && (k) < 128) {{ const long long t_ = clock64(); prof_ts[128 + (k)] += t_ - prof_prev; prof_prev = t_; }}
#ifdef DSPB_PROF_S2     // the scalar-stream stamps come from the SECOND scalar warp
#define PROF_MARK_S(k)
#define PROF_MARK_S2(k) PROF_MARK_SX(k)
#else
#define PROF_MARK_S(k) PROF_MARK_SX(k)
#define PROF_MARK_S2(k)
#endif
#else
#define PROF_MARK(k)
#define PROF_MARK_S(k)
#define PROF_MARK_S2(k)
#endif

__global__ void __launch_bounds__({nthr}, {occ}) k_chain_spec(const __grid_constant__ Args A) {{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CScr* cs = reinterpret_cast<CScr*>(smem_raw + 2048);
  long long* prof_ts = reinterpret_cast<long long*>(smem_raw + 2048 + 8192);
  WaveSummary* summ = reinterpret_cast<WaveSummary*>(smem_raw + {summ_off});
  float* slots = reinterpret_cast<float*>(smem_raw + {self.fixed_bytes});
  (void)prof_ts; (void)cs; (void)summ;
  // (the shuffle makes the warp index provably warp-uniform: branches on it are uniform branches and the
  // collectives inside them need no re-convergence code)
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool scalar_warp = warp >= 16;
  int par = 0;
  if (!scalar_warp) {{
    zero_pads(slots, {self.slot_words}, {self.nchunks}, 0, {self.n_slots}, tid);
    BSYNC();
  }}
  int it = 0;
  {prefetch_init}
  {scalar_prefetch}
#ifdef DSPB_PROFILE
  for (int k = tid; k < 256; k += {nthr}) prof_ts[k] = 0;
  __syncthreads();
  long long prof_prev = clock64();
#endif
  for (long long row = blockIdx.x; row < A.n_rows; row += gridDim.x, it++) {{
    // cells that carry data between the two streams alternate with the row parity, so the block
    // stream can run one row ahead of the scalar warp
    const int rp = it & 1;
    double* bc = reinterpret_cast<double*>(smem_raw + 2048 + 5120) + 16 * rp;        // scalar warp -> block warps
    double* mbd = reinterpret_cast<double*>(smem_raw + {mbd_off} + {self.MB_BUDGET} * rp);  // block warps -> scalar warp
    int* mbi = reinterpret_cast<int*>(smem_raw + {mbi_off} + {self.MB_BUDGET} * rp);
    (void)bc; (void)mbd; (void)mbi;
#define EVB(k) (2 + 5 * rp + (k))
#define EVX (12 + rp)
    {{
      {decl}
      {prolog}
      if (!scalar_warp) {{
        // =============================== block stream (warps 0-15) ===============================
        {body_b}
      }} else if (warp == 16) {{
        // =============================== scalar stream (warp 16) =================================
        {body_s}
      }} else {{
        // =============================== second scalar stream (warp 17) ===========================
        {body_s2}
      }}
    }}
#undef EVB
#undef EVX
  }}
  // consume the scalar warp's last "done" / "progress" events
  if (!scalar_warp && it > 0) {{
    EV_WAIT(15);
    {"EV_WAIT(14);" if self.progress_seq else ""}
  }}
#ifdef DSPB_PROFILE
  __syncthreads();
  if (A.prof && blockIdx.x == 0)
    for (int k = tid; k < N_NODES && k < 128; k += {nthr}) {{
      A.prof[k] += prof_ts[k];
      A.prof[N_NODES + k] += prof_ts[128 + k];
      if (k < 32) A.prof[2 * N_NODES + k] += prof_ts[96 + k];   // PROF_SUB stamps inside routines
    }}
#endif
}}
}}  // namespace

extern "C" int chain_smem_bytes() {{ return {self.smem_bytes}; }}
extern "C" int chain_n_nodes() {{ return N_NODES; }}
extern "C" int chain_launch(const void* const* ptrs, long long n_ptrs, long long n_rows, int* fatal, long long* prof,
                            int num_sms, void* stream) {{
  if (n_ptrs != NP && !(n_ptrs == 0 && NP == 1)) return DSPB_ERR_UNSUPPORTED;
  if (n_rows <= 0) return 0;
  const long long* strides = reinterpret_cast<const long long*>(ptrs + n_ptrs);
{align_check}  Args a;
  for (int i = 0; i < n_ptrs; i++) {{ a.p[i] = ptrs[i]; a.s[i] = strides[i]; }}
  a.n_rows = n_rows;
  a.row0 = strides[n_ptrs];
  const double* consts = reinterpret_cast<const double*>(strides + n_ptrs + 1);   // NC doubles behind the row offset
  for (int i = 0; i < NC; i++) a.c[i] = consts[i];
  a.fatal = fatal;
  a.prof = prof;
  cudaError_t e = cudaFuncSetAttribute(k_chain_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, {self.smem_bytes});
  if (e != cudaSuccess) return -(int)e;
  const long long ctas = (long long)num_sms * {occ};      // persistent CTAs: {occ} resident per SM
  const int grid = (int)(n_rows < ctas ? n_rows : ctas);
  k_chain_spec<<<grid, {nthr}, {self.smem_bytes}, (cudaStream_t)stream>>>(a);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}}
"""

    def _build(self):
        # two resident CTAs per SM when shared memory allows it and the register budget (65536 / (2 x threads)) costs
        # no spills; else one CTA with the full register file
        self.occ = 1
        if self.smem_bytes <= (MAX_SMEM - 2048) // 2 and os.environ.get("DSPEED_B200_OCCUPANCY", "2") != "1":
            self.occ = 2
            if spill_bytes(build_source(self.source())[0]) > 0:
                self.occ = 1
        src = self.source()
        self.lib_path, self.src_path = build_source(src)
        self.handle = C.c_void_p(1)  # marks "runnable" for can_run()
        if self.meta:
            return
        self.lib = load_chain_lib(self.lib_path)
        self.num_sms = torch.cuda.get_device_properties(self.chain.device).multi_processor_count
        self.d_prof = None

    def _launch(self, arr, n, n_rows, fatal_ptr, stream):
        prof = C.c_void_p(self.d_prof.data_ptr()) if self.d_prof is not None else C.c_void_p(0)
        # the launch table: n pointers, n strides, the row offset, then the run-time constants (doubles)
        consts = list(getattr(self, "rt_consts", [])) or [0.0]
        ext = (C.c_int64 * (2 * n + 1 + len(consts)))()
        C.memmove(ext, arr, 8 * (2 * n + 1))
        C.memmove(C.byref(ext, 8 * (2 * n + 1)), (C.c_double * len(consts))(*consts), 8 * len(consts))
        return self.lib.chain_launch(C.cast(ext, C.c_void_p), C.c_int64(n), C.c_int64(n_rows), C.c_void_p(fatal_ptr), prof,
                                     C.c_int(self.num_sms), C.c_void_p(stream))

    def profile(self, run, repeats=1):
        """per-node SM cycles of CTA 0 (see fusion.profile_fused); runs a tracing build of the kernel"""
        n = 2 * (len(self.order) + 1)
        so, _ = build_source(self.source(), flags=("-DDSPB_PROFILE",))
        prod_lib, self.lib = self.lib, load_chain_lib(so)
        self.d_prof = torch.zeros(n + 32, dtype=torch.int64, device=self.chain.device)
        try:
            run()   # warm-up of the tracing build
            self.d_prof.zero_()
            for _ in range(repeats):
                run()
            torch.cuda.synchronize(self.chain.device)
            cyc = self.d_prof.cpu().numpy().astype(np.float64)
        finally:
            self.d_prof = None
            self.lib = prod_lib
        tot = cyc[:n].sum() or 1.0
        text = self.program_text.split("\n") + ["store scalars"]
        text = ["B " + t for t in text] + ["S " + t for t in text]   # block stream, then scalar stream
        text += [f"B sub-stamp {k}" for k in range(32)]              # PROF_SUB stamps (inside routines)
        return [(cyc[i], cyc[i] / tot, text[i]) for i in range(n + 32) if i < n or cyc[i] > 0]

    def __del__(self):
        pass


# ----------------------------------------------------------------------------------------
# compile cache (in-tree: the built libraries travel with the repository snapshot)
# ----------------------------------------------------------------------------------------
def _headers_digest() -> str:
    h = hashlib.sha1()
    for f in ("common.cuh", "row_ops.cuh", "conv_ops.cuh", "chain_rt.cuh", "warp_rt.cuh"):
        h.update(open(os.path.join(_lib.CSRC, f), "rb").read())
    h.update(open(os.path.join(_lib.INCLUDE, "dspeed_b200.h"), "rb").read())
    return h.hexdigest()


def cache_dirs():
    """where compiled chain kernels live: the in-tree cache (prebuilt by build(); travels with the repository) first,
    then a per-user cache for installations whose package directory is read-only ($DSPEED_B200_CACHE overrides)"""
    user = os.environ.get("DSPEED_B200_CACHE") or os.path.join(
        os.environ.get("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")), "dspeed_b200", "chains")
    return [CACHE_DIR, user]


def build_source(src: str, flags=()):
    """compile one generated kernel for sm_100a (cached by content hash).  Safe across processes (torchrun ranks with a
    cold cache): an exclusive file lock per kernel, the source written through a private temporary and renamed, the
    library likewise.  A cache that cannot be written raises NotSpecializable: the chain falls back to the interpreted
    program instead of crashing."""
    import fcntl

    extra = list(flags) + os.environ.get("DSPEED_B200_NVCC_EXTRA", "").split()   # experiments: -D switches
    tag = hashlib.sha1((src + _headers_digest() + " ".join(extra)).encode()).hexdigest()[:16]
    for d in cache_dirs():                          # a kernel that exists anywhere is used as is
        so = os.path.join(d, f"chain_{tag}.so")
        if os.path.exists(so):
            return so, os.path.join(d, f"chain_{tag}.cu")
    cache = None
    for d in cache_dirs():
        try:
            os.makedirs(d, exist_ok=True)
            if os.access(d, os.W_OK):
                cache = d
                break
        except OSError:
            continue
    if cache is None:
        raise NotSpecializable("no writable cache directory for the generated kernel (set DSPEED_B200_CACHE)")
    so = os.path.join(cache, f"chain_{tag}.so")
    cu = os.path.join(cache, f"chain_{tag}.cu")
    try:
        lock_file = open(os.path.join(cache, f".chain_{tag}.lock"), "w")
    except OSError as e:
        raise NotSpecializable(f"cannot lock the kernel cache: {e}")
    with _build_lock, lock_file:
        fcntl.flock(lock_file, fcntl.LOCK_EX)
        if not os.path.exists(so):
            try:
                tmp_cu = cu + f".tmp{os.getpid()}"
                with open(tmp_cu, "w") as f:
                    f.write(src)
                os.replace(tmp_cu, cu)
            except OSError as e:
                raise NotSpecializable(f"cannot write the generated kernel: {e}")
            tmp = so + f".tmp{os.getpid()}"
            cmd = [_lib._nvcc(), *_lib.NVCC_FLAGS, "-Xptxas", "-v", *extra, "-I", _lib.INCLUDE, "-I", _lib.CSRC,
                   "-o", tmp, cu]
            env = dict(os.environ)
            env.pop("CC", None)
            env.pop("CXX", None)
            log.info("compiling specialised chain kernel: " + " ".join(cmd))
            try:
                done = subprocess.run(cmd, env=env, check=True, capture_output=True, text=True)
            except FileNotFoundError as e:
                raise NotSpecializable(f"nvcc not available: {e}")
            except subprocess.CalledProcessError as e:
                raise RuntimeError(f"nvcc failed on {cu}:\n{e.stderr[-4000:]}")
            spilled = sum(int(m) for m in re.findall(r"(\d+) bytes spill stores", done.stderr))
            with open(so + ".spills", "w") as f:     # ptxas' verdict travels with the library (see spill_bytes)
                f.write(str(spilled))
            os.replace(tmp, so)
    return so, cu


def spill_bytes(so: str) -> int:
    """bytes of register spill stores ptxas reported when `so` was built (0 if unknown)"""
    try:
        with open(so + ".spills") as f:
            return int(f.read())
    except (OSError, ValueError):
        return 0


def load_chain_lib(path: str) -> C.CDLL:
    if path not in _loaded:
        _loaded[path] = C.CDLL(path)
    return _loaded[path]


def prebuild(config, wf_len=8192, with_baseline=True, dt_ns=16):
    """Plan `config` on the meta device (no GPU needed) and compile its specialised kernel into the
    in-tree cache, so that a GPU process that builds the same chain finds it ready
    (``__graft_entry__.build`` does this for the shipped configurations)."""
    from . import tables
    from .processing_chain import build_processing_chain

    n = 4
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=dt_ns, dt_units="ns",
                              values=np.zeros((n, wf_len), np.uint16))
    cols = {"waveform": wf}
    if with_baseline:
        cols["baseline"] = tables.Array(np.zeros(n, np.uint16))
    chain, _, _ = build_processing_chain(config, tables.Table(cols, size=n), block_width=16, device="meta")
    try:
        sc = SpecChain(chain)
    except NotFusable as e:
        return None, str(e)
    return sc.lib_path, sc.program_text
