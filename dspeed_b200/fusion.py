"""Chain compiler: lowers a compiled :class:`ProcessingChain` (its frozen list of
``ProcessorManager`` launch descriptors) into ONE program for the waveform-resident CUDA
kernel (``csrc/fused.cu``).

What the reference does per block -- ~60 gufunc calls, every intermediate waveform
written to and re-read from memory (processing_chain.py:1144-1163) -- becomes one kernel
launch per block: the raw row is read from HBM once, intermediates live in shared-memory
slots (allocated here by liveness), per-event scalars in a shared scalar file, scalar
glue (``np.multiply``, unit conversions, ``round``) folds into scalar instructions, and
only requested outputs are written back.

Every instruction carries the index of the processor it came from, so a data-dependent
``DSPFatal`` is still attributed to the right processor.  If any processor of the chain
has no lowering the whole chain stays on the per-processor path (still on the GPU).

Convolution lowering is decided from the constant kernel array (see ``_lower_conv``):
run-structured kernels -> sparse FIR + float64 cumulative sum (exact); cusp/zac kernels
whose analytic model matches the array -> weighted prefix sums; everything else ->
register-tiled direct convolution.
"""

from __future__ import annotations

import ctypes as C
import logging
import os

import numpy as np
import torch

from . import _lib, numpy_bridge
from . import processors as P

log = logging.getLogger("dspeed")

IARGS = 15
MAX_SREG = 128
MAX_PTRS = 64
MAX_SMEM = 227 * 1024
FIXED_SMEM = 2048 + MAX_SREG * 8 + 16 * 4 + 32 * 48 + 384 * 8 + 160 * 64
FUSED_THREADS = 512
R_OUT, CH = 8, 64

(OP_END, OP_LOAD_WAVE, OP_LOAD_SCALAR, OP_STORE_SCALAR, OP_STORE_WAVE, OP_BL_SUB, OP_MIN_MAX, OP_LSF, OP_POLE_ZERO,
 OP_DPZ, OP_TRAP, OP_ASYM, OP_TRAP_PICKOFF, OP_MW, OP_AVG_CURRENT, OP_TPT, OP_ITPT, OP_FTP, OP_WINDOWER, OP_UPSAMPLER,
 OP_CONV_DIRECT, OP_CONV_RUNS, OP_CONV_SEG, OP_SC_BIN, OP_SC_CONVERT, OP_SC_UNARY, OP_MIN_MAX_NORM, OP_LSD,
 OP_MBT, OP_PREFIX, OP_PFIR) = range(31)

_DT = {torch.float32: 0, torch.float64: 1, torch.uint16: 2, torch.int16: 3, torch.int32: 4, torch.uint32: 5,
       torch.int64: 6}


#: rows per launch when all inputs and outputs are device-resident (nothing to stage)
RESIDENT_ROWS_PER_LAUNCH = int(os.environ.get("DSPEED_B200_RESIDENT_ROWS", 1 << 22))


class NotFusable(Exception):
    pass


def seg_params(origin, L):
    """(CONV_SEG parameters, float64 model of the kernel array, is_zac) of a cusp_filter /
    zac_filter kernel of length L built from `origin` = (generator name, args); None when the
    arguments are outside what the weighted-prefix-sum evaluation supports"""
    f = np.float32
    sigma, flat, decay = (float(f(P._as_float(x))) for x in origin[1][:3])
    if sigma <= 0 or flat < 0 or decay <= 0 or flat != np.floor(flat):
        return None
    lt = int((L - flat) / 2)
    fl = int(flat)
    if lt < 2 or lt + fl + 1 >= L or L / sigma > 40.0:
        return None
    zac = origin[0] == "zac_filter"
    i = np.arange(L, dtype=np.float64)
    S = np.sinh(lt / sigma)
    shape = np.zeros(L)
    shape[:lt] = np.sinh(i[:lt] / sigma) / S
    shape[lt : lt + fl + 1] = 1.0
    shape[lt + fl + 1 :] = np.sinh((L - i[lt + fl + 1 :]) / sigma) / S
    beta, h = 0.0, lt / 2.0
    if zac:
        par = np.zeros(L)
        par[:lt] = (i[:lt] - h) ** 2 - h**2
        par[lt + fl + 1 :] = (L - i[lt + fl + 1 :] - h) ** 2 - h**2
        beta = -shape.sum() / par.sum()
        shape = shape + beta * par
    c = float(np.exp(-1.0 / decay))
    model = shape.copy()
    model[1:] -= c * shape[:-1]
    prm = [sigma, float(lt), float(fl), float(L), c, 1.0 / (2.0 * S), float(shape[L - 1]), beta, h,
           1.0 if zac else 0.0]
    return prm, model, zac


def _slot_words(n: int) -> int:
    return (n + (n >> 5) + 1 + 3) & ~3


_ALIAS: dict[int, int] = {}


def _storage_id(t: torch.Tensor) -> int:
    """identity of the storage a tensor views (the StorageImpl address: also defined on the
    "meta" device, where data pointers are all null, so programs can be planned without a GPU)"""
    return t.untyped_storage()._cdata


def _storage(t: torch.Tensor) -> int:
    st = _storage_id(t)
    return _ALIAS.get(st, st)


class _Wave:
    """a root waveform buffer living in a shared-memory slot"""

    def __init__(self, length):
        self.length = length
        self.slot = None
        self.last_use = -1


class FusedChain:
    kernel_name = "k_chain (fused waveform-resident chain kernel)"

    def __init__(self, chain):
        self.chain = chain
        self.handle = C.c_void_p()
        self.device_time = 0.0
        self.device_calls = 0
        self.rows_per_launch = 0
        self._events = []
        self._compile(chain)

    # ------------------------------------------------------------------------------------
    # compilation
    # ------------------------------------------------------------------------------------
    def _compile(self, chain):
        managers = list(chain._proc_managers)
        self.n_managers = len(managers)
        _ALIAS.clear()
        self.prefix: dict[int, int] = {}      # wave storage -> first slot of its float64 prefix array
        self.prefix_last: dict[int, int] = {}  # wave storage -> last manager index that wants the prefix
        self.cse_skipped = 0
        self.code: list[list[int]] = []
        self.consts: list[float] = []
        self.ptrs: list = []          # ("in", manager, what) | ("buf", tensor) | ("const", tensor)
        self.ptr_index: dict = {}
        self.sreg: dict[int, int] = {}   # storage ptr -> scalar register
        self.waves: dict[int, _Wave] = {}
        self.const_storage: dict[int, torch.Tensor] = {}
        self.input_wave: dict[int, tuple] = {}
        self.input_scalar: dict[int, tuple] = {}
        self.n_out_rows_buffers = []

        # ---- classify storages ------------------------------------------------------------
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

        # ---- pass 0: identical wave producers (same processor, same bound inputs) share one
        # result, e.g. wf_etrap == wf_trap with the default database ------------------------------
        seen = {}
        self.skip = set()
        for i, pm in enumerate(managers):
            name = getattr(pm.processor, "__name__", "")
            if not getattr(pm.processor, "native_kernel", False) or getattr(pm.processor, "nout", 1) != 1:
                continue
            out = pm.args[-1]
            if not (isinstance(out, torch.Tensor) and out.ndim == 2 and out.storage_offset() == 0):
                continue
            key = [name]
            for x in pm.args[:-1]:
                if isinstance(x, torch.Tensor):
                    key.append(("t", _storage(x), x.storage_offset(), tuple(x.shape), tuple(x.stride())))
                else:
                    key.append(("c", repr(x)))
            key = tuple(key)
            if key in seen and tuple(seen[key].shape) == tuple(out.shape):
                _ALIAS[_storage_id(out)] = _storage(seen[key])
                self.skip.add(i)
            else:
                seen[key] = out

        # ---- pass 1: last use of every wave storage ------------------------------------------
        for i, pm in enumerate(managers):
            if i not in self.skip and getattr(pm.processor, "__name__", "") in (
                    "trap_filter", "trap_norm", "asym_trap_filter", "convolve_wf", "fft_convolve_wf"):
                x = pm.args[0]
                if isinstance(x, torch.Tensor) and x.ndim == 2:
                    self.prefix_last[_storage(x)] = i
            for a in pm.args:
                if isinstance(a, torch.Tensor) and a.ndim >= 2 and _storage(a) not in self.const_storage:
                    w = self.waves.setdefault(_storage(a), _Wave(self._root_len(a)))
                    w.last_use = i
        out_wave_storages = {}
        for name, man in chain._output_managers.items():
            rv = self._out_raw(man)
            if rv.ndim >= 2:
                out_wave_storages[_storage(rv)] = rv
        self.slot_len = max([w.length for w in self.waves.values()] + [1])
        slot_bytes = _slot_words(self.slot_len) * 4
        self.max_slots = min(8, (MAX_SMEM - FIXED_SMEM) // slot_bytes)
        if self.max_slots < 2:
            raise NotFusable("waveforms too long for the shared-memory resident layout")
        self.free_slots = list(range(self.max_slots))
        self.slots_used = 0

        # ---- pass 2: lower every manager ---------------------------------------------------------
        for i, pm in enumerate(managers):
            self._cur = i
            self._fatal_idx = pm.fatal.storage_offset() // 4
            if i in self.skip:
                self.cse_skipped += 1
            else:
                self._lower(pm)
            # outputs that are chain outputs are stored as soon as they are produced
            for a in self._outputs_of(pm):
                if isinstance(a, torch.Tensor) and a.ndim >= 2 and _storage(a) in out_wave_storages \
                        and _storage(a) in self.waves and self.waves[_storage(a)].slot is not None:
                    rv = out_wave_storages[_storage(a)]
                    self._emit(OP_STORE_WAVE, self.waves[_storage(a)].slot, rv.storage_offset() % max(1, rv.stride(0)),
                               rv.shape[1], self._ptr(("buf", rv)))
            self._release_dead(i)
        # pass-through outputs (an input column copied to the output) and scalar outputs
        for name, man in chain._output_managers.items():
            rv = self._out_raw(man)
            if rv.ndim == 1:
                st = _storage(rv)
                if st in self.const_storage:
                    continue  # constants are written by the output manager itself
                if st in self.input_scalar and st not in self.sreg:
                    self._scalar_operand(rv)
                if st not in self.sreg:
                    raise NotFusable(f"output {name} is not produced by a fusable processor")
                if rv.dtype not in (torch.float32, torch.float64, torch.int32, torch.uint32):
                    raise NotFusable(f"output dtype {rv.dtype}")
                self._emit(OP_STORE_SCALAR, self.sreg[st], self._ptr(("buf", rv)), _DT[rv.dtype])
            else:
                st = _storage(rv)
                if st in self.input_wave and (st not in self.waves or self.waves[st].slot is None):
                    raise NotFusable("waveform pass-through outputs stay on the copy path")
        if len(self.ptrs) > MAX_PTRS:
            raise NotFusable("too many distinct device pointers")
        self.n_slots = self.slots_used
        flat = [self.n_slots, self.slot_len, len(self.code)]
        for ins in self.code:
            flat.extend(ins)
        code = np.asarray(flat, dtype=np.int32)
        consts = np.asarray(self.consts if self.consts else [0.0], dtype=np.float64)
        lib = _lib.lib()
        with torch.cuda.device(chain.device):
            rc = lib.dspb_chain_create(code.ctypes.data_as(C.c_void_p), C.c_int64(code.size),
                                       consts.ctypes.data_as(C.c_void_p), C.c_int64(len(self.consts)),
                                       C.byref(self.handle))
        if rc:
            raise NotFusable(f"dspb_chain_create failed with {rc}")
        lib.dspb_chain_smem_bytes.restype = C.c_int64
        self.smem_bytes = int(lib.dspb_chain_smem_bytes(self.handle))
        self.program_text = self._describe()

    # -- inputs ----------------------------------------------------------------------------------
    def _register_input(self, man):
        from . import processing_chain as pc

        if isinstance(man, pc.WaveformIOManager):
            vm = man.val_ioman
            if not isinstance(vm, pc.ArrayIOManager):
                raise NotFusable("vector-of-vector waveform input")
            self.input_wave[_storage(vm.raw_var)] = (vm, "nda")
            if man.variable_t0:
                self.input_scalar[_storage(man.t0_var)] = (man, "t0")
        elif isinstance(man, (pc.ArrayIOManager, pc.NumpyIOManager)):
            rv = man.raw_var
            if rv.ndim >= 2:
                self.input_wave[_storage(rv)] = (man, "nda")
            else:
                self.input_scalar[_storage(rv)] = (man, "nda")
        else:
            raise NotFusable(f"input manager {type(man).__name__}")

    @staticmethod
    def _out_raw(man):
        from . import processing_chain as pc

        if isinstance(man, pc.WaveformIOManager):
            return man.val_ioman.raw_var
        if isinstance(man, pc.VectorOfVectorsIOManager):
            raise NotFusable("vector-of-vector output")
        return man.raw_var

    # -- helpers -----------------------------------------------------------------------------------
    def _emit(self, op, *args, fatal=True):
        a = [int(x) for x in args] + [0] * (IARGS - len(args))
        if len(a) > IARGS:
            raise NotFusable("instruction too long")
        a[14] = int(self._fatal_idx) if fatal else 0
        self.code.append([op] + a)

    def _const(self, *vals) -> int:
        idx = len(self.consts)
        self.consts.extend(float(v) for v in vals)
        return idx

    def _ptr(self, key) -> int:
        k = (key[0], _storage(key[1]) if isinstance(key[1], torch.Tensor) else id(key[1]),
             key[1].storage_offset() if isinstance(key[1], torch.Tensor) else 0) + tuple(key[2:])
        if k not in self.ptr_index:
            self.ptr_index[k] = len(self.ptrs)
            self.ptrs.append(key)
        return self.ptr_index[k]

    @staticmethod
    def _root_len(t: torch.Tensor) -> int:
        """row length of the root buffer a (possibly sliced) wave tensor belongs to"""
        return int(t.stride(0)) if t.shape[0] > 1 or t.stride(0) >= t.shape[-1] else int(t.shape[-1])

    def _alloc_slot(self) -> int:
        if not self.free_slots:
            self._evict_prefixes()
        if not self.free_slots:
            raise NotFusable("not enough shared-memory slots for the live waveforms")
        s = self.free_slots.pop(0)
        self.slots_used = max(self.slots_used, s + 1)
        return s

    def _free(self, *slots):
        self.free_slots.extend(slots)
        self.free_slots.sort()

    def _evict_prefixes(self, keep=None):
        """prefix arrays are caches: give their slot pairs back under pressure (they are
        recomputed by a PREFIX instruction when the next consumer needs them)"""
        for st in list(self.prefix):
            if st != keep:
                s0 = self.prefix.pop(st)
                self._free(s0, s0 + 1)

    def _prefix_of(self, t: torch.Tensor, slot_in: int, n: int):
        """first slot of the float64 inclusive-prefix array of a whole waveform (two adjacent
        slots), emitting the PREFIX instruction when it is not resident; None if there is no room"""
        st = _storage(t)
        if st in self.prefix:
            return self.prefix[st]
        if 8 * (n + (n >> 5) + 1) > 2 * _slot_words(self.slot_len) * 4:
            return None

        def find_pair():
            for a in self.free_slots:
                if a + 1 in self.free_slots:
                    return a
            return None

        s0 = find_pair()
        if s0 is None:
            self._evict_prefixes()
            s0 = find_pair()
        if s0 is None or len(self.free_slots) < 3:  # keep one slot for the consumer's output
            return None
        self.free_slots.remove(s0)
        self.free_slots.remove(s0 + 1)
        self.slots_used = max(self.slots_used, s0 + 2)
        self.prefix[st] = s0
        self._emit(OP_PREFIX, slot_in, 0, n, s0)
        return s0

    def _pfir(self, t_in, s, n, so, p, taps, shift) -> bool:
        """out[i] = sum_s c_s P[i + shift - t_s]; False when the prefix array cannot be resident"""
        if len(taps) > 32:
            return False
        # the output slot must not be one of the prefix slots: allocate the prefix first
        s0 = self._prefix_of(t_in, s, n)
        if s0 is None or so in (s0, s0 + 1):
            return False
        pairs = []
        for t, c in taps:
            pairs += [float(t), float(c)]
        self._emit(OP_PFIR, s, 0, n, so, p, self._const(*pairs), len(taps), shift, s0)
        return True

    def _release_dead(self, i):
        for st, w in self.waves.items():
            if w.slot is not None and w.last_use <= i:
                self._free(w.slot)
                w.slot = None
                w.dead = True
        for st in list(self.prefix):
            if self.prefix_last.get(st, -1) <= i:
                s0 = self.prefix.pop(st)
                self._free(s0, s0 + 1)

    def _wave_in(self, t: torch.Tensor, need_zero_offset=False):
        """(slot, offset, n) of an input waveform operand; loads chain inputs on first use"""
        if t.ndim != 2 or (t.shape[-1] > 1 and t.stride(-1) != 1):
            raise NotFusable("unsupported waveform view (stride)")
        st = _storage(t)
        w = self.waves.get(st)
        if w is None:
            raise NotFusable("unknown waveform operand")
        if w.slot is None:
            if getattr(w, "dead", False):
                raise NotFusable("waveform used after its slot was released")
            if st not in self.input_wave:
                raise NotFusable("waveform operand read before it is produced")
            w.slot = self._alloc_slot()
            man, what = self.input_wave[st]
            rv = man.raw_var
            self._emit(OP_LOAD_WAVE, w.slot, self._ptr(("in", man, what)), rv.shape[1], _DT[rv.dtype])
        off = t.storage_offset() % max(1, w.length) if t.storage_offset() else 0
        if need_zero_offset and off != 0:
            raise NotFusable("processor needs an unsliced waveform")
        return w.slot, int(off), int(t.shape[1])

    def _wave_out(self, t: torch.Tensor):
        if t.ndim != 2 or t.storage_offset() != 0:
            raise NotFusable("waveform outputs must be whole buffers")
        st = _storage(t)
        w = self.waves.setdefault(st, _Wave(int(t.shape[1])))
        if w.slot is None:
            w.slot = self._alloc_slot()
            w.dead = False
        return w.slot, int(t.shape[1])

    def _scalar_operand(self, x):
        """(kind, index) of a per-event scalar operand: 0 = scalar register, 1 = constant"""
        if isinstance(x, torch.Tensor):
            st = _storage(x)
            if st in self.const_storage:
                v = self.const_storage[st].reshape(-1)
                if v.numel() != 1:
                    raise NotFusable("non-scalar constant used as a scalar")
                return 1, self._const(float(v.cpu()[0]))
            if x.numel() != x.shape[0]:
                raise NotFusable("vector-valued per-event variable")
            if st not in self.sreg:
                if st not in self.input_scalar:
                    raise NotFusable("scalar operand read before it is produced")
                man, what = self.input_scalar[st]
                src = man.t0_var if what == "t0" else man.raw_var
                if src.dtype not in _DT:
                    raise NotFusable(f"scalar input dtype {src.dtype}")
                self.sreg[st] = self._new_reg()
                self._emit(OP_LOAD_SCALAR, self.sreg[st], self._ptr(("in", man, what)), _DT[src.dtype], 0)
            return 0, self.sreg[st]
        if x is None:
            raise NotFusable("None argument")
        return 1, self._const(float(x))

    def _new_reg(self) -> int:
        n = len(self.sreg) + getattr(self, "_extra_regs", 0)
        if n >= MAX_SREG:
            raise NotFusable("too many per-event scalars")
        return n

    def _scalar_out(self, t: torch.Tensor) -> int:
        if not isinstance(t, torch.Tensor) or t.numel() != t.shape[0]:
            raise NotFusable("scalar output must be a [block] tensor")
        st = _storage(t)
        if st not in self.sreg:
            self.sreg[st] = self._new_reg()
        return self.sreg[st]

    @staticmethod
    def _outputs_of(pm):
        n = getattr(pm.processor, "nout", 1) or 1
        return pm.args[-n:]

    # -- lowering ------------------------------------------------------------------------------------
    def _lower(self, pm):
        from . import processing_chain as pc

        if isinstance(pm, pc.UnitConversionManager):
            buf, off_in, off_out, ratio, out = pm.args
            if buf.numel() != buf.shape[0] or out.dtype not in (torch.float32, torch.float64):
                raise NotFusable("unit conversion of a non-scalar / integer variable")
            ki, ii = self._scalar_operand(buf)
            if ki != 0:
                raise NotFusable("conversion of a constant")
            oi = self._scalar_operand(off_in)
            oo = self._scalar_operand(off_out)
            mode = {None: 0, "round": 1, "floor": 2, "ceil": 3, "trunc": 4}[pm.mode]
            if pm.in_is_int and pm.mode is None:
                raise NotFusable("integer conversion check")
            self._emit(OP_SC_CONVERT, ii, oi[0], oi[1], oo[0], oo[1], self._const(ratio), self._scalar_out(out), mode,
                       1 if out.dtype == torch.float32 else 0)
            return

        proc = pm.processor
        name = proc.__name__
        a = pm.args
        f32 = not any(t.char == "d" for t in pm.types)

        if isinstance(proc, numpy_bridge.ElementwiseOp):
            if any(isinstance(x, torch.Tensor) and x.numel() != x.shape[0] and _storage(x) not in self.const_storage
                   for x in a):
                raise NotFusable(f"element-wise {name} on waveforms")
            out = a[-1]
            if out.dtype not in (torch.float32, torch.float64):
                raise NotFusable(f"{name} with {out.dtype} output")
            dt = 0 if out.dtype == torch.float32 else 1
            if name in ("add", "subtract", "multiply", "divide", "floor_divide"):
                x, y = self._scalar_operand(a[0]), self._scalar_operand(a[1])
                code = ["add", "subtract", "multiply", "divide", "floor_divide"].index(name)
                self._emit(OP_SC_BIN, code, x[0], x[1], y[0], y[1], self._scalar_out(out), dt)
                return
            if name == "negative":
                x = self._scalar_operand(a[0])
                self._emit(OP_SC_UNARY, 0, x[0], x[1], 0, 0, self._scalar_out(out))
                return
            raise NotFusable(f"element-wise {name}")

        if not getattr(proc, "native_kernel", False):
            raise NotFusable(f"helper processor {name}")
        if not f32:
            raise NotFusable(f"{name}: float64 type loop")

        if name == "bl_subtract":
            s, off, n = self._wave_in(a[0])
            k = self._scalar_operand(a[1])
            so, _ = self._wave_out(a[2])
            self._emit(OP_BL_SUB, s, off, n, so, k[0], k[1])
        elif name in ("min_max", "amax"):
            s, off, n = self._wave_in(a[0])
            if name == "min_max":
                regs = [self._scalar_out(x) for x in a[1:5]]
            else:
                regs = [-1, -1, -1, self._scalar_out(a[2])]
            self._emit(OP_MIN_MAX, s, off, n, *regs)
        elif name in ("linear_slope_fit", "mean_stdev"):
            s, off, n = self._wave_in(a[0])
            regs = [self._scalar_out(x) for x in a[1:]]
            regs += [-1] * (4 - len(regs))
            self._emit(OP_LSF, s, off, n, *regs)
        elif name == "linear_slope_diff":
            s, off, n = self._wave_in(a[0])
            sl, ic = self._scalar_operand(a[1]), self._scalar_operand(a[2])
            self._emit(OP_LSD, s, off, n, sl[0], sl[1], ic[0], ic[1], self._scalar_out(a[3]), self._scalar_out(a[4]))
        elif name == "mean_below_threshold":
            s, off, n = self._wave_in(a[0])
            th = self._scalar_operand(a[1])
            self._emit(OP_MBT, s, off, n, th[0], th[1], self._scalar_out(a[2]))
        elif name == "pole_zero":
            s, off, n = self._wave_in(a[0], True)
            k = self._scalar_operand(a[1])
            so, _ = self._wave_out(a[2])
            self._emit(OP_POLE_ZERO, s, 0, n, so, k[0], k[1])
        elif name == "double_pole_zero":
            s, off, n = self._wave_in(a[0], True)
            if n <= 3:
                raise NotFusable("double_pole_zero on a too short waveform")
            k1, k2, k3 = (self._scalar_operand(x) for x in a[1:4])
            so, _ = self._wave_out(a[4])
            self._emit(OP_DPZ, s, 0, n, so, k1[0], k1[1], k2[0], k2[1], k3[0], k3[1])
        elif name in ("trap_filter", "trap_norm"):
            s, off, n = self._wave_in(a[0], True)
            rise, flat = int(a[1]), int(a[2])
            if rise < 0 or flat < 0 or 2 * rise + flat > n:
                raise NotFusable("invalid trapezoid arguments (the per-processor path raises the DSPFatal)")
            norm = name == "trap_norm"
            if rise == 0 and norm:
                raise NotFusable("trap_norm with rise == 0")
            c = 1.0 / rise if norm else 1.0
            s0 = self._prefix_of(a[0], s, n)   # before the output slot, so the pair stays adjacent
            so, _ = self._wave_out(a[3])
            if s0 is None or not self._pfir(a[0], s, n, so, n, [(0, c), (rise, -c), (rise + flat, -c),
                                                              (2 * rise + flat, c)], 0):
                self._emit(OP_TRAP, s, 0, n, so, rise, flat, 1 if norm else 0)
        elif name == "asym_trap_filter":
            s, off, n = self._wave_in(a[0], True)
            rise, flat, fall = int(a[1]), int(a[2]), int(a[3])
            if min(rise, flat, fall) < 0 or rise + flat + fall > n:
                raise NotFusable("invalid trapezoid arguments")
            if rise == 0 or fall == 0:
                raise NotFusable("asym_trap_filter with a zero-length side")
            s0 = self._prefix_of(a[0], s, n)
            so, _ = self._wave_out(a[4])
            if s0 is None or not self._pfir(a[0], s, n, so, n, [(0, 1.0 / rise), (rise, -1.0 / rise),
                                                              (rise + flat, -1.0 / fall),
                                                              (rise + flat + fall, 1.0 / fall)], 0):
                self._emit(OP_ASYM, s, 0, n, so, rise, flat, fall)
        elif name == "trap_pickoff":
            s, off, n = self._wave_in(a[0], True)
            rise, flat = int(a[1]), int(a[2])
            if rise < 0 or flat < 0 or 2 * rise + flat > n:
                raise NotFusable("invalid trapezoid arguments")
            t = self._scalar_operand(a[3])
            self._emit(OP_TRAP_PICKOFF, s, 0, n, rise, flat, t[0], t[1], self._scalar_out(a[4]))
        elif name in ("moving_window_left", "moving_window_right", "moving_window_multi"):
            s, off, n = self._wave_in(a[0], True)
            length = float(np.float32(a[1]))
            if name == "moving_window_multi":
                num, typ = float(np.float32(a[2])), int(a[3])
                if length != np.floor(length) or num != np.floor(num) or not (0 <= int(length) < n) or num < 0:
                    raise NotFusable("invalid moving-window arguments")
                so, _ = self._wave_out(a[4])
                tmp = self._alloc_slot()
                self._emit(OP_MW, s, 0, n, so, tmp, self._const(length), 2, int(num), typ)
                self._free(tmp)
            else:
                if not (0 <= length < n):
                    raise NotFusable("invalid moving-window arguments")
                so, _ = self._wave_out(a[2])
                self._emit(OP_MW, s, 0, n, so, 0, self._const(length), 0 if name.endswith("left") else 1, 0, 0)
        elif name == "avg_current":
            s, off, n = self._wave_in(a[0], True)
            length = float(np.float32(a[1]))
            so, n_out = self._wave_out(a[2])
            if not (0 <= length < n) or n_out != n - int(length):
                raise NotFusable("invalid avg_current arguments")
            self._emit(OP_AVG_CURRENT, s, 0, n, so, n_out, self._const(length))
        elif name == "time_point_thresh":
            s, off, n = self._wave_in(a[0], True)
            th, ts, wk = (self._scalar_operand(x) for x in a[1:4])
            self._emit(OP_TPT, s, 0, n, th[0], th[1], ts[0], ts[1], wk[0], wk[1], self._scalar_out(a[4]))
        elif name == "interpolated_time_point_thresh":
            s, off, n = self._wave_in(a[0], True)
            th, ts = self._scalar_operand(a[1]), self._scalar_operand(a[2])
            self._emit(OP_ITPT, s, 0, n, th[0], th[1], ts[0], ts[1], int(a[3]), P._as_int(a[4]), self._scalar_out(a[5]))
        elif name == "fixed_time_pickoff":
            s, off, n = self._wave_in(a[0], True)
            t = self._scalar_operand(a[1])
            self._emit(OP_FTP, s, 0, n, t[0], t[1], P._as_int(a[2]), self._scalar_out(a[3]))
        elif name == "windower":
            s, off, n = self._wave_in(a[0], True)
            t0 = self._scalar_operand(a[1])
            so, m = self._wave_out(a[2])
            if m >= n:
                raise NotFusable("invalid windower length")
            self._emit(OP_WINDOWER, s, 0, n, so, m, t0[0], t0[1])
        elif name == "upsampler":
            s, off, n = self._wave_in(a[0], True)
            up = float(np.float32(a[1]))
            if not up > 0:
                raise NotFusable("invalid upsample factor")
            so, m = self._wave_out(a[2])
            self._emit(OP_UPSAMPLER, s, 0, n, so, m, self._const(up))
        elif name == "min_max_norm":
            s, off, n = self._wave_in(a[0], True)
            mn, mx = self._scalar_operand(a[1]), self._scalar_operand(a[2])
            so, _ = self._wave_out(a[3])
            self._emit(OP_MIN_MAX_NORM, s, 0, n, so, mn[0], mn[1], mx[0], mx[1])
        elif name in ("convolve_wf", "fft_convolve_wf"):
            self._lower_conv(pm)
        else:
            raise NotFusable(f"no fused lowering for {name}")

    # -- convolutions ----------------------------------------------------------------------------------
    def _lower_conv(self, pm):
        w_in, kernel, mode, w_out = pm.args
        if not isinstance(kernel, torch.Tensor) or _storage(kernel) not in self.const_storage:
            raise NotFusable("convolution kernel is not a constant")
        k = kernel.reshape(-1).detach().cpu().numpy().astype(np.float32)
        s, off, n = self._wave_in(w_in)
        m = int(k.size)
        mode = chr(P._as_int(mode))
        if m > n or mode not in "fvs":
            raise NotFusable("invalid convolution arguments")
        p = {"f": n + m - 1, "v": n - m + 1, "s": n}[mode]
        coff = {"f": 0, "v": m - 1, "s": (m - 1) // 2}[mode]
        if np.isnan(k).any():
            raise NotFusable("NaN in convolution kernel")
        choice = os.environ.get("DSPEED_B200_CONV", "auto")
        dk0 = np.diff(np.concatenate([[0.0], k.astype(np.float64), [0.0]]))
        runs_ok = choice in ("auto", "runs") and len(np.flatnonzero(dk0)) <= 24
        s0 = self._prefix_of(w_in, s, n) if (runs_ok and off == 0) else None
        so, p_out = self._wave_out(w_out)
        if p_out != p:
            raise NotFusable("convolution output length mismatch")

        # (1) run-structured kernels: sparse first difference -> FIR + cumulative sum (exact)
        dk = np.diff(np.concatenate([[0.0], k.astype(np.float64), [0.0]]))
        taps = np.flatnonzero(dk)
        if runs_ok and s0 is not None and self._pfir(w_in, s, n, so, p, [(int(t), float(dk[t])) for t in taps], coff):
            self.conv_lowering = getattr(self, "conv_lowering", []) + [("runs", m, len(taps))]
            return
        if runs_ok:
            pairs = []
            for t in taps:
                pairs += [float(t), float(dk[t])]
            self._emit(OP_CONV_RUNS, s, off, n, so, p, self._const(*pairs), len(taps), coff)
            self.conv_lowering = getattr(self, "conv_lowering", []) + [("runs", m, len(taps))]
            return

        # (2) cusp / zac kernels whose analytic model reproduces the array
        seg = self._seg_model(kernel, k) if (choice in ("auto", "seg") and mode == "v" and off == 0) else None
        if seg is not None and 13 * p * 8 <= _slot_words(self.slot_len) * 4:
            scratch = self._alloc_slot()
            self._emit(OP_CONV_SEG, s, 0, n, so, p, self._const(*seg), 0, 0, scratch)
            self._free(scratch)
            self.conv_lowering = getattr(self, "conv_lowering", []) + [("seg", m, "zac" if seg[9] else "cusp")]
            return

        # (3) direct
        G = (p + R_OUT - 1) // R_OUT
        S, scratch = 1, 0
        if G < FUSED_THREADS:
            S = max(1, min(FUSED_THREADS // G, (m + CH - 1) // CH))
            while S > 1 and S * G * R_OUT * 8 > _slot_words(self.slot_len) * 4:
                S -= 1
        if S > 1:
            scratch = self._alloc_slot()
        self._emit(OP_CONV_DIRECT, s, off, n, so, p, self._ptr(("const", kernel)), m, coff, scratch, S)
        if S > 1:
            self._free(scratch)
        self.conv_lowering = getattr(self, "conv_lowering", []) + [("direct", m, S)]

    def _seg_model(self, kernel_t, k):
        """parameters of the CONV_SEG lowering if the kernel came from cusp_filter / zac_filter
        (energy_kernels.py:12-157) AND the float64 analytic model reproduces the actual array
        to float32 rounding; None otherwise."""
        var = self.var_of_storage.get(_storage(kernel_t))
        origin = getattr(var, "const_origin", None)
        if not origin or origin[0] not in ("cusp_filter", "zac_filter"):
            return None
        res = seg_params(origin, int(k.size))
        if res is None:
            return None
        prm, model, zac = res
        # the reference rounds the shape to float32 before the differencing (cusp) or only the
        # final kernel (zac): allow exactly that much
        tol = 2.5e-7 if not zac else 4e-10 * max(1.0, float(np.abs(model).max()) / 1.7e-3)
        err = float(np.abs(model - k.astype(np.float64)).max())
        if not err <= tol:
            log.debug(f"cusp/zac model rejected: residual {err:.3e} > {tol:.1e}")
            return None
        return prm

    def _describe(self) -> str:
        names = ["END", "LOAD_WAVE", "LOAD_SCALAR", "STORE_SCALAR", "STORE_WAVE", "BL_SUB", "MIN_MAX", "LSF", "POLE_ZERO",
                 "DPZ", "TRAP", "ASYM", "TRAP_PICKOFF", "MW", "AVG_CURRENT", "TPT", "ITPT", "FTP", "WINDOWER",
                 "UPSAMPLER", "CONV_DIRECT", "CONV_RUNS", "CONV_SEG", "SC_BIN", "SC_CONVERT", "SC_UNARY",
                 "MIN_MAX_NORM", "LSD", "MBT", "PREFIX", "PFIR"]
        return "\n".join(f"{i:3d} {names[ins[0]]:13s} {ins[1:]}" for i, ins in enumerate(self.code))

    # ------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------
    def can_run(self, chain) -> bool:
        return self.handle.value is not None and len(chain._proc_managers) == self.n_managers

    # -- input staging: host columns are copied into device staging buffers on a separate
    # copy stream, double-buffered, so the H2D transfer of block k+1 overlaps the kernel of
    # block k; device-resident columns are read in place (no copy at all) -------------------
    def _input_sources(self):
        """[(ptr-table index, source column, staging buffers[2] or None)]"""
        if getattr(self, "_sources", None) is not None:
            return self._sources
        srcs = []
        for idx, key in enumerate(self.ptrs):
            if key[0] != "in":
                continue
            man, what = key[1], key[2]
            buf = man.t0_var if what == "t0" else man.raw_var
            srcs.append([idx, man, what, buf, [buf, None]])
        self._sources = srcs
        return srcs

    @staticmethod
    def _column(man, what):
        if what == "t0":
            return man.io_wf.t0.nda
        return man.io_array.nda if hasattr(man, "io_array") else man.io_buf

    @staticmethod
    def _t0_source(chain, out_man):
        """input `t0` column that a waveform output's per-event offset variable mirrors (same block buffer), or
        NotFusable when the offset is derived (another unit system / a conversion): those chains take the
        per-processor path, where the conversion managers fill it"""
        from . import processing_chain as pc

        for m in chain._input_managers.values():
            if (isinstance(m, pc.WaveformIOManager) and m.variable_t0 and isinstance(out_man.t0_var, torch.Tensor)
                    and _storage(m.t0_var) == _storage(out_man.t0_var)
                    and m.t0_var.storage_offset() == out_man.t0_var.storage_offset()):
                return m.io_wf.t0.nda
        raise NotFusable("waveform output whose per-event t0 is not an input column")

    def _check_outputs(self, chain):
        from . import processing_chain as pc

        for man in chain._output_managers.values():
            if isinstance(man, pc.WaveformIOManager) and man.variable_t0:
                self._t0_source(chain, man)

    def _resident(self, src) -> bool:
        """is this input column consumed in place (already on the device, same dtype, unit stride)?"""
        idx, man, what, buf, staging = src
        col = self._column(man, what)
        return bool(isinstance(col, torch.Tensor) and col.is_cuda and col.dtype == buf.dtype
                    and (col.ndim == 1 or col.stride(-1) == 1))

    def _stage_block(self, begin, end, stage, copy_stream):
        """enqueue the H2D copies of rows [begin, end) into staging set `stage`; returns
        {ptr index: (device pointer, row stride)}"""
        res = {}
        for src in self._input_sources():
            idx, man, what, buf, staging = src
            col = self._column(man, what)
            if self._resident(src):
                view = col[begin:end]
                res[idx] = (view.data_ptr(), view.stride(0))
                continue
            if staging[stage] is None:
                staging[stage] = torch.empty_like(buf)
            dst = staging[stage]
            t = col if isinstance(col, torch.Tensor) else torch.from_numpy(col)
            if not t.is_cuda and not getattr(self, "_warned_pageable", False) and not t.is_pinned():
                # a copy from pageable memory is staged by the driver and synchronous: no overlap with the kernels
                self._warned_pageable = True
                log.warning("input column %r lives in pageable host memory: host-to-device copies will not overlap the "
                            "kernels (allocate it with dspeed_b200.tables.pinned_empty, or torch .pin_memory())",
                            f"{getattr(getattr(man, 'var', None), 'name', '?')}.{what}")
            n = end - begin
            with torch.cuda.stream(copy_stream):
                dst[:n].copy_(t[begin:end], non_blocking=True)
            if not t.is_cuda:
                self.chain.stats["h2d_bytes"] += n * dst[0].numel() * dst.element_size()
            res[idx] = (dst.data_ptr(), dst.stride(0))
        return res

    def execute(self, chain, start, stop):
        from . import processing_chain as pc

        lib = _lib.lib()
        n_in = min((len(m.io_wf) if isinstance(m, pc.WaveformIOManager) else
                    (len(m.io_array) if hasattr(m, "io_array") else m.io_t.shape[0]))
                   for m in chain._input_managers.values()) if chain._input_managers else stop
        stop = min(stop, n_in)
        bw = chain._block_width
        if stop <= start:
            return
        static = {}
        for idx, key in enumerate(self.ptrs):
            if key[0] != "in":
                t = key[1]
                static[idx] = (t.data_ptr(), t.stride(0) if t.ndim >= 1 else 0)
        # Output columns that already live on this device are written by the kernel in place (row
        # `begin` of the column is row 0 of the launch); only host columns go through the block
        # buffer and a D2H copy.
        direct, copied = {}, []
        for man in chain._output_managers.values():
            try:
                rv = self._out_raw(man)
            except NotFusable:
                copied.append(man)
                continue
            idx = self.ptr_index.get(("buf", _storage(rv), rv.storage_offset()))
            col = None
            if isinstance(man, pc.ArrayIOManager):
                col = man.io_array.nda
            elif isinstance(man, pc.NumpyIOManager):
                col = man.io_buf
            if (idx is not None and isinstance(col, torch.Tensor) and col.is_cuda and col.device == rv.device
                    and col.dtype == rv.dtype and col.shape[0] >= stop and tuple(col.shape[1:]) == tuple(rv.shape[1:])
                    and (col.ndim == 1 or col.stride(-1) == 1) and not getattr(man.var, "is_const", False)):
                direct[idx] = col
            else:
                copied.append(man)
        # Blocks exist for the staging buffers (host columns in, block buffers out).  When every column the
        # kernel touches already lives on the device, nothing is staged and the whole range is ONE launch
        # (persistent CTAs stride over the rows): no per-block launch gaps or tails.
        if not copied and all(self._resident(src) for src in self._input_sources()):
            launch_rows = max(bw, RESIDENT_ROWS_PER_LAUNCH)
        else:
            launch_rows = bw
        blocks = [(b, min(b + launch_rows, stop)) for b in range(start, stop, launch_rows)]
        with torch.cuda.device(chain.device):
            compute = torch.cuda.current_stream(chain.device)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=chain.device)
                self._ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_free = [torch.cuda.Event(), torch.cuda.Event()]
            cs = self._copy_stream
            cs.wait_stream(compute)  # staging buffers may still be read by earlier work
            staged = {}

            def issue(k):
                st = k % 2
                if k >= 2:
                    cs.wait_event(self._ev_free[st])  # the kernel that read this staging set is done
                staged[k] = self._stage_block(blocks[k][0], blocks[k][1], st, cs)
                self._ev_copied[st].record(cs)

            issue(0)
            for k, (begin, end) in enumerate(blocks):
                if k + 1 < len(blocks):
                    issue(k + 1)  # overlaps with the kernel of block k
                compute.wait_event(self._ev_copied[k % 2])
                table = dict(static)
                table.update(staged.pop(k))
                for idx, col in direct.items():
                    view = col[begin:end]
                    table[idx] = (view.data_ptr(), view.stride(0))
                n = len(self.ptrs)
                ptrs = [table[i][0] for i in range(n)]
                strides = [table[i][1] for i in range(n)]
                arr = (C.c_int64 * (2 * n + 1))(*ptrs, *strides, begin)
                if chain._event_timing:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                rc = self._launch(arr, n, end - begin, chain.fatal.data_ptr(), compute.cuda_stream)
                if chain._event_timing:
                    e1.record()
                    self._events.append((e0, e1))
                self._ev_free[k % 2].record(compute)
                if rc:
                    raise RuntimeError(f"fused chain launch failed with {rc}")
                self.rows_per_launch = max(self.rows_per_launch, end - begin)
                chain.stats["launches"] += 1
                chain.stats["blocks"] += 1
                for out_man in copied:
                    if isinstance(out_man, pc.WaveformIOManager) and out_man.variable_t0:
                        # per-event t0 of a waveform output: the rows of the input column the offset variable mirrors
                        # (the block's offset buffer is only a staging area here, and not even that for device-resident
                        # inputs or chains whose kernel never reads t0)
                        col = self._t0_source(chain, out_man)
                        t = col if isinstance(col, torch.Tensor) else torch.from_numpy(np.asarray(col))
                        out_man.write(begin, end, t0_src=t[begin:end])
                    else:
                        out_man.write(begin, end)
            compute.synchronize()
            # data-dependent DSPFatal conditions are recorded on the device with their row: one
            # check per call (the reference raises after the offending block, :1156-1159)
            chain._raise_recorded_fatal(start, stop, block_width=bw)
        if self._events:
            for e0, e1 in self._events:
                self.device_time += e0.elapsed_time(e1) * 1e-3
                self.device_calls += 1
            self._events = []

    def _launch(self, arr, n, n_rows, fatal_ptr, stream):
        return _lib.lib().dspb_chain_launch(self.handle, C.cast(arr, C.c_void_p), C.c_int64(n), C.c_int64(n_rows),
                                            C.c_void_p(fatal_ptr), C.c_void_p(stream))

    def __del__(self):
        try:
            if self.handle.value:
                _lib.lib().dspb_chain_destroy(self.handle)
        except Exception:
            pass


def try_fuse(chain) -> bool:
    """attach a fused program to the chain if every processor has a lowering: the specialised
    (generated, straight-line) kernel when every processor has an emitter, else the interpreted
    program"""
    if os.environ.get("DSPEED_B200_SPECIALIZE", "1") != "0":
        from . import codegen, warpchain

        # short waveforms (<= 2048 samples): one warp per waveform, vector outputs supported
        try:
            chain._fused = warpchain.WarpChain(chain)
            chain._fused._check_outputs(chain)
            log.debug(f"warp-per-waveform chain kernel:\n{chain._fused.program_text}")
            return True
        except NotFusable as e:
            chain._not_warp_reason = str(e)
        try:
            chain._fused = codegen.SpecChain(chain)
            chain._fused._check_outputs(chain)
            log.debug(f"specialised chain kernel:\n{chain._fused.program_text}")
            return True
        except NotFusable as e:
            log.info(f"chain not specialised ({e}); trying the interpreted program")
            chain._not_specialised_reason = str(e)
    try:
        chain._fused = FusedChain(chain)
        chain._fused._check_outputs(chain)
        log.debug(f"fused chain program:\n{chain._fused.program_text}")
        return True
    except NotFusable as e:
        log.info(f"chain not fused ({e}); running one kernel per processor")
        chain._fused = None
        chain._not_fused_reason = str(e)
        return False


def profile_fused(chain, run, repeats: int = 1):
    """Per-instruction SM-cycle profile of the fused program (CTA 0): calls ``run()``
    `repeats` times with the counters enabled and returns [(cycles, share, text)] in
    program order -- the device-side counterpart of the reference's per-processor timers
    (processing_chain.py:1778-1781)."""
    fc = chain._fused
    if fc is None:
        raise RuntimeError("chain is not fused")
    if hasattr(fc, "profile"):
        return fc.profile(run, repeats)
    lib = _lib.lib()
    n = len(fc.code)
    lib.dspb_chain_profile(fc.handle, 1, None)
    for _ in range(repeats):
        run()
    out = (C.c_int64 * n)()
    lib.dspb_chain_profile(fc.handle, 0, out)
    cyc = np.array(out[:], dtype=np.float64)
    tot = cyc.sum() or 1.0
    text = fc.program_text.split("\n")
    return [(cyc[i], cyc[i] / tot, text[i]) for i in range(n)]
