"""Warp-per-waveform chain kernels: the specialised tier for SHORT waveforms and vector outputs.

The block-per-waveform kernels of ``codegen.py`` own one 8192-sample waveform per CTA; for the
SiPM / LAr-instrumentation chains (BASELINE.json config 4: ~2000-sample waveforms, peak lists as
``[block, 20]`` outputs; reference config ``tests/configs/sipm-dsp-config.json``) a CTA is far too
wide, so this generator emits ONE straight-line kernel in which every *warp* owns a waveform
(``csrc/warp_rt.cuh``): lane ``l`` keeps samples ``[CH l, CH l + CH)`` in registers from one
processor to the next, the raw ``uint16`` row of the warp's next waveform is staged into shared
memory with ``cp.async`` while the current one is processed, window filters exchange halos by
shuffles, and ``get_multi_local_extrema`` (the reference's sequential peak-detection state machine,
``processors/get_multi_local_extrema.py:12-306``) is walked event-by-event by the whole warp.  There
is no block-wide barrier and no HBM traffic besides the raw row and the outputs.

Processors with an emitter here: ``bl_subtract``, ``moving_window_left/right/multi``, ``avg_current``,
``min_max`` / ``amax``, ``get_multi_local_extrema`` (search directions 0, 1, 3), per-event scalar
arithmetic, and the unit conversions of scalars and of ``[block, m]`` vector variables
(``processing_chain.py:1806-1908``).  Anything else raises ``NotSpecializable`` and the chain falls back
to the other GPU tiers (``fusion.try_fuse``).
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import numpy_bridge
from .codegen import spill_bytes, _CTYPE, NotSpecializable, _flit, _lit, build_source, load_chain_lib
from .fusion import FusedChain, NotFusable, _storage

MAX_LEN = 2048
WARPS_PER_CTA = 4


class _Wave:
    def __init__(self, n, reg, nan="false"):
        self.n = n
        self.reg = reg
        self.nan = nan      # warp-uniform bool expression: the whole wave is NaN


class WarpChain(FusedChain):
    kernel_name = "k_chain_warp (warp-per-waveform chain kernel)"

    # ------------------------------------------------------------------------------------
    def _compile(self, chain):
        from . import processing_chain as pc

        managers = list(chain._proc_managers)
        self.n_managers = len(managers)
        self.meta = chain.device.type == "meta"
        self.ptrs, self.ptr_index = [], {}
        self.waves: dict[int, _Wave] = {}
        self.svar: dict[int, str] = {}       # per-event scalar storage -> C variable
        self.stype: dict[str, str] = {}
        self.vvar: dict[int, tuple] = {}     # [block, m] vector storage -> (C variable, m)
        self.const_storage, self.var_of_storage = {}, {}
        self.input_wave, self.input_scalar = {}, {}
        self.lines: list[str] = []
        self.text: list[str] = []
        self.aligned_ptrs: list[int] = []
        self._sources = None
        self._tmp = 0
        self.uses_slot = False

        all_vars = list(chain._vars_dict.values())
        for pm in managers:
            for prm in list(pm.params) + list(pm.kw_params.values()):
                if hasattr(prm, "proc_chain") and prm not in all_vars:
                    all_vars.append(prm)
        for v in all_vars:
            bufs = v.all_buffers()
            for b, _ in bufs:
                if isinstance(b, torch.Tensor):
                    self.var_of_storage.setdefault(_storage(b), v)
                    if v.is_const:
                        self.const_storage[_storage(b)] = b
        for man in chain._input_managers.values():
            self._register_input(man)
        lens = [int(m.raw_var.shape[1]) for (m, _) in self.input_wave.values()]
        if not lens:
            raise NotSpecializable("no waveform input")
        if max(lens) > MAX_LEN:
            raise NotSpecializable(f"waveforms longer than {MAX_LEN} samples run on the block-per-waveform kernels")
        self.CH = 8
        while 32 * self.CH < max(lens):
            self.CH *= 2

        for i, pm in enumerate(managers):
            self._fatal_idx = pm.fatal.storage_offset() // 4
            self._lower(pm)
        # ---- outputs ------------------------------------------------------------------------------
        n_out = 0
        for name, man in chain._output_managers.items():
            rv = self._out_raw(man)
            st = _storage(rv)
            if st in self.const_storage:
                continue
            pi = self._ptr(("buf", rv))
            if rv.dtype not in _CTYPE:
                raise NotSpecializable(f"output dtype {rv.dtype}")
            ct = _CTYPE[rv.dtype]
            if rv.ndim == 1:
                if st in self.input_scalar and st not in self.svar:
                    self._sc(rv)
                if st not in self.svar:
                    raise NotSpecializable(f"output {name} is not produced by a processor of this tier")
                self._e(f"if (lane == 0) (({ct}*)A.p[{pi}])[row * A.s[{pi}]] = ({ct})({self.svar[st]});")
            elif st in self.vvar:
                var, m = self.vvar[st]
                if tuple(rv.shape[1:]) != (m,):
                    raise NotSpecializable("vector output shape")
                self._e(f"if (lane < {m}) (({ct}*)A.p[{pi}])[row * A.s[{pi}] + lane] = ({ct})({var});")
            elif st in self.waves and rv.ndim == 2 and rv.dtype == torch.float32 and rv.storage_offset() == 0:
                w = self.waves[st]
                if getattr(w, "is_input", False) or rv.shape[1] != w.n:
                    raise NotSpecializable("waveform pass-through outputs stay on the copy path")
                self._e(f"stg_chunk<CH>((float*)A.p[{pi}] + row * A.s[{pi}], lane, {w.n}, {w.reg}, {w.nan});")
            else:
                raise NotSpecializable(f"output {name}: unsupported shape / producer")
            n_out += 1
        if n_out == 0:
            raise NotSpecializable("no outputs")
        if len(self.ptrs) > 64:
            raise NotSpecializable("too many distinct device pointers")
        self.program_text = "\n".join(f"{i:3d} {t}" for i, t in enumerate(self.text))
        self._build()

    # -- helpers -----------------------------------------------------------------------------------
    def _e(self, *lines):
        self.lines.extend(lines)

    def _t(self, p="t"):
        self._tmp += 1
        return f"{p}{self._tmp}"

    def _const_value(self, t):
        var = self.var_of_storage.get(_storage(t))
        hv = getattr(var, "host_value", None)
        if hv is not None:
            return np.asarray(hv).reshape(-1)
        if self.meta:
            raise NotSpecializable("constant without a host value on the meta device")
        return self.const_storage[_storage(t)].reshape(-1).detach().cpu().numpy()

    def _new_svar(self, st, dtype):
        name = f"s{len(self.svar)}"
        self.svar[st] = name
        self.stype[name] = ("float" if dtype == torch.float32 else
                            "unsigned" if dtype == torch.uint32 else "int" if dtype == torch.int32 else "double")
        return name

    def _sc(self, x):
        """C expression of a per-event scalar operand (constants become literals)"""
        if isinstance(x, torch.Tensor):
            st = _storage(x)
            if st in self.const_storage:
                v = self._const_value(x)
                if v.size != 1:
                    raise NotSpecializable("non-scalar constant used as a scalar")
                return _lit(v[0])
            if x.numel() != x.shape[0]:
                raise NotSpecializable("vector-valued variable used as a scalar")
            if st not in self.svar:
                if st not in self.input_scalar:
                    raise NotSpecializable("scalar operand read before it is produced")
                man, what = self.input_scalar[st]
                src = man.t0_var if what == "t0" else man.raw_var
                if src.dtype not in _CTYPE:
                    raise NotSpecializable(f"scalar input dtype {src.dtype}")
                name = self._new_svar(st, src.dtype)
                pi = self._ptr(("in", man, what))
                self._e(f"const {self.stype[name]} {name} = ({self.stype[name]})((const {_CTYPE[src.dtype]}*)A.p[{pi}])"
                        f"[row * A.s[{pi}]];")
            return self.svar[st]
        if x is None:
            raise NotSpecializable("None argument")
        return _lit(float(x))

    def _const_arg(self, x, what):
        if isinstance(x, torch.Tensor):
            if _storage(x) not in self.const_storage:
                raise NotSpecializable(f"{what} must be a constant in this tier")
            v = self._const_value(x)
            if v.size != 1:
                raise NotSpecializable(f"{what} must be a scalar")
            return float(v[0])
        return float(x)

    def _sout(self, t):
        if not isinstance(t, torch.Tensor) or t.numel() != t.shape[0]:
            raise NotSpecializable("scalar output must be a [block] tensor")
        st = _storage(t)
        if st in self.svar:
            raise NotSpecializable("scalar variable written twice")
        if t.dtype not in (torch.float32, torch.float64, torch.uint32, torch.int32):
            raise NotSpecializable(f"scalar dtype {t.dtype}")
        return self._new_svar(st, t.dtype)

    def _win(self, t, whole=True):
        if not isinstance(t, torch.Tensor) or t.ndim != 2 or (t.shape[-1] > 1 and t.stride(-1) != 1):
            raise NotSpecializable("unsupported waveform view")
        st = _storage(t)
        w = self.waves.get(st)
        if w is None:
            if st not in self.input_wave:
                raise NotSpecializable("waveform operand read before it is produced")
            man, what = self.input_wave[st]
            rv = man.raw_var
            if rv.dtype not in (torch.uint16, torch.int16):
                raise NotSpecializable(f"input waveform dtype {rv.dtype} (16-bit ADC samples only)")
            n = int(rv.shape[1])
            if n % 8 or n < 32:
                raise NotSpecializable("input waveform length must be a multiple of 8")
            if hasattr(self, "in_ptr"):
                raise NotSpecializable("more than one waveform input")
            self.in_ptr = self._ptr(("in", man, what))
            self.in_n, self.in_signed = n, rv.dtype == torch.int16
            self.aligned_ptrs.append(self.in_ptr)
            w = _Wave(n, "r_in")
            w.is_input = True
            self.waves[st] = w
            self.text.append(f"load {rv.dtype} [{n}] -> r_in")
        off = t.storage_offset() % max(1, w.n) if t.storage_offset() else 0
        n = int(t.shape[1])
        if whole and (off != 0 or n != w.n):
            raise NotSpecializable("processor on a waveform slice")
        return w, int(off), n

    def _wout(self, t, n_expected=None):
        if t.ndim != 2 or t.storage_offset() != 0 or t.dtype != torch.float32:
            raise NotSpecializable("waveform outputs must be whole float32 buffers")
        st = _storage(t)
        if st in self.waves:
            raise NotSpecializable("waveform written twice")
        n = int(t.shape[1])
        if n > 32 * self.CH or (n_expected is not None and n != n_expected):
            raise NotSpecializable("waveform output length")
        w = _Wave(n, self._t("r"))
        self.waves[st] = w
        return w

    # -- lowering ------------------------------------------------------------------------------------
    def _lower(self, pm):
        from . import processing_chain as pc

        if isinstance(pm, pc.UnitConversionManager):
            return self._lower_convert(pm)
        proc = pm.processor
        name = proc.__name__
        a = pm.args
        if isinstance(proc, numpy_bridge.ElementwiseOp):
            op = {"add": "+", "subtract": "-", "multiply": "*", "divide": "/"}.get(name)
            out = a[-1]
            if op is None or out.dtype not in (torch.float32, torch.float64) or out.numel() != out.shape[0]:
                raise NotSpecializable(f"element-wise {name}")
            ty = "float" if out.dtype == torch.float32 else "double"
            x, y = self._sc(a[0]), self._sc(a[1])
            o = self._sout(out)
            self._e(f"const {self.stype[o]} {o} = ({ty})({x}) {op} ({ty})({y});")
            self.text.append(f"{name} -> {o}")
            return
        if not getattr(proc, "native_kernel", False):
            raise NotSpecializable(f"helper processor {name}")
        if any(t.char == "d" for t in pm.types):
            raise NotSpecializable(f"{name}: float64 type loop")
        CH = self.CH
        if name == "bl_subtract":
            w, _, n = self._win(a[0])
            b = self._sc(a[1])
            out = self._wout(a[2], n)
            b_is_float = isinstance(a[1], torch.Tensor) and a[1].dtype in (torch.float32, torch.float64) and \
                not b.startswith(("0x", "-0x"))
            bb = self._t("b")
            self._e(f"const float {bb} = (float)({b});",
                    f"float {out.reg}[CH]; bl_sub<CH>({w.reg}, {bb}, {out.reg});")
            out.nan = f"({w.nan} || {bb} != {bb})" if b_is_float or w.nan != "false" else "false"
            self.text.append(f"bl_sub {w.reg} -> {out.reg}")
        elif name in ("moving_window_left", "moving_window_right", "moving_window_multi"):
            w, _, n = self._win(a[0])
            length = float(np.float32(self._const_arg(a[1], "window length")))
            if length != np.floor(length) or not (1 <= int(length) < n):
                raise NotSpecializable("moving-window arguments (another tier raises the DSPFatal)")
            L = int(length)
            if L > CH:
                raise NotSpecializable("moving window longer than a lane chunk")
            if name == "moving_window_multi":
                num, typ = float(np.float32(self._const_arg(a[2], "num_mw"))), int(self._const_arg(a[3], "mw_type"))
                if num != np.floor(num) or num < 1 or typ not in (0, 1, 2):
                    raise NotSpecializable("moving-window arguments")
                dirs = ["r" if ((k & 1) and typ == 0) or typ == 2 else "l" for k in range(int(num))]
                out = self._wout(a[4], n)
            else:
                dirs = ["l" if name.endswith("left") else "r"]
                out = self._wout(a[2], n)
            il = _flit(1.0 / float(np.float32(L)))
            src = w.reg
            for k, d in enumerate(dirs):
                dst = out.reg if k == len(dirs) - 1 else self._t("r")
                if d == "l":
                    self._e(f"float {dst}[CH]; mw_left<CH, {L}, {n}>({src}, {dst}, {il}, lane);")
                else:
                    self._e(f"float {dst}[CH]; mw_right<CH, {L}, {n}>({src}, {dst}, {il}, lane);")
                src = dst
            out.nan = w.nan
            self.text.append(f"mw L={L} {''.join(dirs)} {w.reg} -> {out.reg}")
        elif name == "avg_current":
            w, _, n = self._win(a[0])
            length = float(np.float32(self._const_arg(a[1], "window length")))
            if length != np.floor(length) or not (1 <= int(length) < n) or int(length) > CH:
                raise NotSpecializable("avg_current arguments")
            L = int(length)
            out = self._wout(a[2], n - L)
            self._e(f"float {out.reg}[CH]; avg_current<CH, {L}>({w.reg}, {out.reg}, {_flit(1.0 / float(np.float32(L)))}, lane); "
                    f"edge_extend<CH, {n - L}>({out.reg}, lane);")
            out.nan = w.nan
            self.text.append(f"avg_current L={L} {w.reg} -> {out.reg}")
        elif name in ("min_max", "amax"):
            w, off, n = self._win(a[0], whole=False)
            outs = list(a[1:5]) if name == "min_max" else [None, None, None, a[2]]
            tm, tM, am, aM = (self._t("m") for _ in range(4))
            self._e(f"float {tm}, {tM}, {am}, {aM}; min_max<CH>({w.reg}, {off}, {off + n}, lane, {tm}, {tM}, {am}, {aM});")
            for o, v in zip(outs, (tm, tM, am, aM)):
                if o is not None:
                    s = self._sout(o)
                    self._e(f"const {self.stype[s]} {s} = ({self.stype[s]})({w.nan} ? CUDART_NAN_F : {v});")
            self.text.append(f"min_max {w.reg}[{off}:{off + n}]")
        elif name == "get_multi_local_extrema":
            self._lower_extrema(pm)
        else:
            raise NotSpecializable(f"no warp-tier emitter for {name}")

    def _lower_extrema(self, pm):
        a = pm.args
        w, _, n = self._win(a[0])
        d_max, d_min, sdir, ab_max, ab_min = (np.float32(self._const_arg(x, "extrema parameter")) for x in a[1:6])
        vmax_t, vmin_t, nmax_t, nmin_t = a[6:10]
        if any(not isinstance(t, torch.Tensor) for t in (vmax_t, vmin_t, nmax_t, nmin_t)):
            raise NotSpecializable("extrema outputs")
        if vmax_t.ndim != 2 or vmin_t.ndim != 2:
            raise NotSpecializable("extrema lists must be [block, m] variables")
        m = int(vmax_t.shape[1])
        if tuple(vmin_t.shape) != tuple(vmax_t.shape) or m > 32 or not m < n:
            raise NotSpecializable("extrema list shape (another tier raises the DSPFatal)")
        if not (d_max >= 0 and d_min >= 0) or float(sdir) not in (0.0, 1.0, 3.0) or np.isnan(ab_max) or np.isnan(ab_min):
            raise NotSpecializable("extrema arguments (another tier handles them)")
        if vmax_t.dtype != torch.float32 or vmin_t.dtype != torch.float32:
            raise NotSpecializable("extrema list dtype")
        for t in (vmax_t, vmin_t):
            if _storage(t) in self.vvar or t.storage_offset() != 0:
                raise NotSpecializable("vector variable written twice")
        sdir = int(sdir)
        self.uses_slot = True
        cs, bx, bn, c0, c1 = (self._t(p) for p in ("cs", "bx", "bn", "c", "c"))
        vx, vn = self._t("v"), self._t("v")
        nx, nn = self._sout(nmax_t), self._sout(nmin_t)
        prm = f"{_flit(d_max)}, {_flit(d_min)}, {_flit(ab_max)}, {_flit(ab_min)}, {m}"
        fetch = []
        if self._alias_wanted() and not getattr(self, "_row_fetched", False):
            # the next row is requested now (the register chunk dies with the summary) and staged at the end of the row
            self._row_fetched = True
            pi = self.in_ptr
            fetch = [f"RowPieces<{self.in_n}> nxt_row; fetch_row_16<{self.in_n}>(nxt_row, (const uint16_t*)A.p[{pi}] + "
                     f"(row + wstride) * A.s[{pi}], lane, row + wstride < A.n_rows);"]
        self._e(f"st_chunk<CH>(S, lane, {w.reg});",
                f"const ChunkSumm<CH> {cs} = chunk_summary<CH>({w.reg});",
                "__syncwarp();", *fetch,
                f"unsigned long long {bx} = 0ull, {bn} = 0ull; int {c0} = 0, {c1} = 0; (void){c0}; (void){c1};",
                f"if (!({w.nan})) {{")
        if sdir in (0, 3):
            self._e(f"  peak_walk<CH, false>(S, {n}, {prm}, {cs}, {bx}, {bn}, {c0}, {c1});")
        if sdir in (1, 3):
            self._e(f"  peak_walk<CH, true>(S, {n}, {prm}, {cs}, {bx}, {bn}, {c0}, {c1});")
        desc = "true" if sdir == 1 else "false"
        self._e("}",
                f"float {vx}, {vn};",
                f"const {self.stype[nx]} {nx} = emit_list<CH>({bx}, {m}, {desc}, lbuf, lane, {vx});",
                f"const {self.stype[nn]} {nn} = emit_list<CH>({bn}, {m}, {desc}, lbuf, lane, {vn});")
        self.vvar[_storage(vmax_t)] = (vx, m)
        self.vvar[_storage(vmin_t)] = (vn, m)
        self.text.append(f"get_multi_local_extrema dir={sdir} m={m} {w.reg} -> {vx}, {vn}, {nx}, {nn}")

    def _lower_convert(self, pm):
        buf, off_in, off_out, ratio, out = pm.args
        if pm.in_is_int and pm.mode is None:
            raise NotSpecializable("integer conversion check")
        if out.dtype not in (torch.float32, torch.float64):
            raise NotSpecializable("conversion to an integer variable")
        fn = {None: "", "round": "rint", "floor": "floor", "ceil": "ceil", "trunc": "trunc"}[pm.mode]
        oi, oo = self._sc(off_in), self._sc(off_out)
        st = _storage(buf)
        if st in self.vvar:
            var, m = self.vvar[st]
            if tuple(out.shape[1:]) != (m,) or out.storage_offset() != 0 or _storage(out) in self.vvar:
                raise NotSpecializable("vector conversion shape")
            o = self._t("v")
            ty = "float" if out.dtype == torch.float32 else "double"
            self._e(f"const {ty} {o} = ({ty}){fn}(((double)({var}) + (double)({oi})) * {_lit(ratio)} - (double)({oo}));")
            self.vvar[_storage(out)] = (o, m)
            self.text.append(f"convert {var} -> {o}")
            return
        if buf.numel() != buf.shape[0]:
            raise NotSpecializable("unit conversion of a waveform")
        x = self._sc(buf)
        o = self._sout(out)
        self._e(f"const {self.stype[o]} {o} = ({self.stype[o]}){fn}(((double)({x}) + (double)({oi})) * {_lit(ratio)} - (double)({oo}));")
        self.text.append(f"convert {x} -> {o}")

    # ------------------------------------------------------------------------------------
    # source
    # ------------------------------------------------------------------------------------
    def _geometry(self):
        CH = self.CH
        raw = 32 * (2 * CH + 16)
        slot = 32 * (CH + 4) * 4
        # Shared memory per warp: the raw staging buffer of the next row and -- only when a peak finder walks a wave --
        # one float copy of that wave.  The copy is dead from the end of the walk to the next row's store, the staging
        # buffer from the register read of a row to the arrival of the next one: they are ALIASED (max instead of sum),
        # which lets 50 % more warps be resident; the kernel is bound by the latency of dependent shuffle / ballot
        # chains, so resident warps are what counts (C4: 16 -> 24 warps per SM, + 14 %).  The next row then travels
        # through registers: requested before the walk, written to the staging buffer after it (_lower_extrema).
        self.alias = self.uses_slot and self._alias_wanted()
        if not self.uses_slot:
            self.warp_smem, self.s_off, self.lbuf_off = raw + 128, 0, raw
        elif self.alias:
            self.warp_smem, self.s_off, self.lbuf_off = max(raw, slot) + 128, 0, max(raw, slot)
        else:
            self.warp_smem, self.s_off, self.lbuf_off = raw + slot + 128, raw, raw + slot
        # resident CTAs of 4 warps: the register file allows 8 CTAs at 64 registers (CH <= 32: the chunk is <= 32
        # registers) and 6 at 80 (CH = 64); shared memory may allow fewer
        self.ctas_per_sm = int(os.environ.get("DSPEED_B200_WARP_CTAS", "0")) or (6 if CH >= 32 else 8)
        if getattr(self, "max_ctas", None):
            self.ctas_per_sm = min(self.ctas_per_sm, self.max_ctas)
        self.raw_bytes, self.slot_bytes = raw, slot
        self.smem_bytes = WARPS_PER_CTA * self.warp_smem
        while self.ctas_per_sm > 1 and self.ctas_per_sm * (self.smem_bytes + 1024) > 227 * 1024:
            self.ctas_per_sm -= 1

    @staticmethod
    def _alias_wanted() -> bool:
        return os.environ.get("DSPEED_B200_WARP_ALIAS", "1") != "0"

    def source(self) -> str:
        if not hasattr(self, "in_ptr"):
            raise NotSpecializable("chain without a waveform load")
        self._geometry()
        np_ = max(1, len(self.ptrs))
        body = "\n    ".join(self.lines)
        pi, n = self.in_ptr, self.in_n
        sg = "true" if self.in_signed else "false"
        align_check = "".join(f"  if (((uintptr_t)ptrs[{i}] & 15) || (strides[{i}] & 7)) return DSPB_ERR_UNSUPPORTED;\n"
                              for i in self.aligned_ptrs)
        prefetch = (f"if (row + wstride < A.n_rows) stage_row_16<CH>(raw, (const uint16_t*)A.p[{pi}] + (row + wstride) * "
                    f"A.s[{pi}], {n}, lane);")
        if self.alias:
            prefetch_early = "// (the staging buffer is aliased with the wave copy: the next row travels through registers)"
            prefetch_late = f"stage_pieces_16<CH, {n}>(raw, nxt_row, lane);"
        else:
            prefetch_early = "// the raw row of this warp's next waveform travels while this one is processed\n    " + prefetch
            prefetch_late = ""
        return f"""// generated by dspeed_b200/warpchain.py -- do not edit
#include "warp_rt.cuh"
using namespace wrt;
namespace {{
constexpr int NP = {np_};
constexpr int CH = {self.CH};
constexpr int WPC = {WARPS_PER_CTA};
struct Args {{
  const void* p[NP];
  long long s[NP];
  long long n_rows, row0;
  int* fatal;
}};

__global__ void __launch_bounds__(32 * WPC, {self.ctas_per_sm}) k_chain_warp(const __grid_constant__ Args A) {{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* raw = smem_raw + warp * {self.warp_smem};
  float* S = reinterpret_cast<float*>(raw + {self.s_off});
  float* lbuf = reinterpret_cast<float*>(raw + {self.lbuf_off});
  (void)S; (void)lbuf;
  const long long wstride = (long long)gridDim.x * WPC;
  long long row = (long long)blockIdx.x * WPC + warp;
  if (row < A.n_rows) stage_row_16<CH>(raw, (const uint16_t*)A.p[{pi}] + row * A.s[{pi}], {n}, lane);
  for (; row < A.n_rows; row += wstride) {{
    cp_async_wait_all();
    __syncwarp();
    float r_in[CH];
    read_chunk_16<CH, {sg}>(raw, lane, r_in);
    edge_extend<CH, {n}>(r_in, lane);
    __syncwarp();
    {prefetch_early}
    {body}
    __syncwarp();
    {prefetch_late}
  }}
}}
}}  // namespace

extern "C" int chain_smem_bytes() {{ return {self.smem_bytes}; }}
extern "C" int chain_n_nodes() {{ return {len(self.text)}; }}
extern "C" int chain_launch(const void* const* ptrs, long long n_ptrs, long long n_rows, int* fatal, long long* prof,
                            int num_sms, void* stream) {{
  (void)prof;
  if (n_ptrs != NP) return DSPB_ERR_UNSUPPORTED;
  if (n_rows <= 0) return 0;
  const long long* strides = reinterpret_cast<const long long*>(ptrs + n_ptrs);
{align_check}  Args a;
  for (int i = 0; i < n_ptrs; i++) {{ a.p[i] = ptrs[i]; a.s[i] = strides[i]; }}
  a.n_rows = n_rows;
  a.row0 = strides[n_ptrs];
  a.fatal = fatal;
  cudaError_t e = cudaFuncSetAttribute(k_chain_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, {self.smem_bytes});
  if (e != cudaSuccess) return -(int)e;
  long long grid = (n_rows + WPC - 1) / WPC;
  const long long cap = (long long)num_sms * {self.ctas_per_sm};
  if (grid > cap) grid = cap;
  k_chain_warp<<<(int)grid, 32 * WPC, {self.smem_bytes}, (cudaStream_t)stream>>>(a);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}}
"""

    def _build(self):
        # most resident CTAs first; a register budget that makes ptxas spill more than a few words inside the row loop
        # costs more than the extra warps bring: step down (6 CTAs = 80 registers, 5 = 96, 4 = 128)
        self.max_ctas = None
        while True:
            self.lib_path, self.src_path = build_source(self.source())
            if spill_bytes(self.lib_path) <= 64 or self.ctas_per_sm <= 4:
                break
            self.max_ctas = self.ctas_per_sm - 1
        self.handle = C.c_void_p(1)
        if self.meta:
            return
        self.lib = load_chain_lib(self.lib_path)
        self.num_sms = torch.cuda.get_device_properties(self.chain.device).multi_processor_count

    def _launch(self, arr, n, n_rows, fatal_ptr, stream):
        return self.lib.chain_launch(C.cast(arr, C.c_void_p), C.c_int64(n), C.c_int64(n_rows), C.c_void_p(fatal_ptr),
                                     C.c_void_p(0), C.c_int(self.num_sms), C.c_void_p(stream))

    def __del__(self):
        pass


def prebuild(config, wf_len=2000, with_baseline=True, dt_ns=16):
    """plan `config` on the meta device and compile its warp-tier kernel into the in-tree cache"""
    from . import tables
    from .processing_chain import build_processing_chain

    n = 4
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=dt_ns, dt_units="ns", values=np.zeros((n, wf_len), np.uint16))
    cols = {"waveform": wf}
    if with_baseline:
        cols["baseline"] = tables.Array(np.zeros(n, np.uint16))
    chain, _, _ = build_processing_chain(config, tables.Table(cols, size=n), block_width=16, device="meta")
    try:
        wc = WarpChain(chain)
    except NotFusable as e:
        return None, str(e)
    return wc.lib_path, wc.program_text
