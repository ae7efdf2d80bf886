"""Host-side checks of the chain compiler (no GPU): the recipe is compiled on torch's
"meta" device (plan only: buffers without data, nothing executes) and the resulting
launch plan is inspected -- processor order, type loops, unit -> sample conversion,
coordinate bookkeeping, output columns."""

import json
import os

import numpy as np
import pytest
import torch
import yaml

from dspeed_b200 import tables
from dspeed_b200.errors import ProcessingChainError
from dspeed_b200.processing_chain import build_processing_chain

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ICPC = os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")
REF_ICPC = "/root/reference/tests/configs/icpc-dsp-config.json"


def raw_table(n=8, wf_len=8192, dt=16):
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=dt, dt_units="ns", values=np.zeros((n, wf_len), np.uint16))
    return tables.Table({"waveform": wf, "baseline": tables.Array(np.zeros(n, np.uint16))}, size=n)


def plan(cfg, db=None, **kw):
    chain, mask, tb_out = build_processing_chain(cfg, raw_table(**kw), db_dict=db, block_width=16, device="meta")

    def fmt(a):
        if isinstance(a, torch.Tensor):
            return f"T{tuple(a.shape)}:{str(a.dtype)[6:]}"
        return repr(a)

    procs = [(str(pm), [fmt(a) for a in pm.args]) for pm in chain._proc_managers]
    def col_dtype(v):
        return str(v.values.dtype) if tables.kind_of(v) == "wftable" else str(v.dtype)

    cols = {k: (dict(v.attrs), col_dtype(v)) for k, v in tb_out.items()}
    return chain, procs, cols, mask


def test_icpc_plan_in_sample_units():
    chain, procs, cols, mask = plan(yaml.safe_load(open(ICPC)))
    assert mask == ["waveform", "baseline"]
    by_name = {}
    for name, args in procs:
        by_name.setdefault(name.split("(")[0], []).append((name, args))
    # unit -> sample conversion on the 16 ns grid (reference processing_chain.py:1747-1770)
    trap = [a for n, a in by_name["trap_norm"]]
    assert ["np.int32(625)", "np.int32(188)"] in [a[1:3] for a in trap]   # 10 us, 3.008 us
    assert ["np.int32(250)", "np.int32(6)"] in [a[1:3] for a in trap]     # 4 us, 96 ns
    assert by_name["asym_trap_filter"][0][1][1:4] == ["np.int32(8)", "np.int32(4)", "np.int32(125)"]
    assert by_name["pole_zero"][0][1][1] == "np.float32(27460.5)"
    # slices are views with the right lengths; kernels are constants of the right length
    assert by_name["linear_slope_fit"][0][1][0] == "T(16, 750):float32"
    assert by_name["linear_slope_fit"][1][1][0] == "T(16, 6692):float32"
    conv = by_name["fft_convolve_wf"][0][1]
    assert conv[0] == "T(16, 6092):float32" and conv[1] == "T(1, 5792):float32" and conv[3] == "T(16, 301):float32"
    assert by_name["convolve_wf"][0][1][1] == "T(1, 133):float32"
    # scalar glue picks numpy's float32 loop and rounds the literal to float32
    assert ("multiply(0.99, trapTmax, (0.99*trapTmax))", ["np.float32(0.99)", "T(16,):float32", "T(16,):float32"]) in procs
    assert by_name["add"][-1][1] in (["T(16,):float32", "np.float32(150.0)", "T(16,):float32"],
                                    ["T(16,):float32", "T(16,):float32", "T(16,):float32"])
    # time outputs are converted from samples of the waveform grid to ns: (i + t0/dt) * 16
    conv_ns = [a for n, a in procs if n.startswith("convert(tp_0_est")]
    assert conv_ns and conv_ns[0][3] == "16.0"
    assert cols["tp_0_est"][0]["units"] == "ns" and cols["trapEmax"][0]["units"] == "ADC"
    assert cols["A_max"][0]["units"] == "ADC/sample"
    assert len(cols) == 34 and all(dt == "float32" for _, dt in cols.values())
    # only what the outputs need is scheduled: every processor output is used
    assert len(procs) == 69


@pytest.mark.skipif(not os.path.exists(REF_ICPC), reason="reference tree not present (GPU box)")
def test_reference_config_compiles_to_the_same_plan():
    """an existing dspeed config runs unchanged: the reference's own JSON and our YAML
    restatement of it compile into the identical launch plan"""
    a = plan(json.load(open(REF_ICPC)))
    b = plan(yaml.safe_load(open(ICPC)))
    assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3]


REF_SIPM = "/root/reference/tests/configs/sipm-dsp-config.json"
SIPM_LAR = os.path.join(os.path.dirname(ICPC), "sipm_lar.yaml")


def sipm_plan(cfg, db=None):
    from dspeed_b200 import synth

    n = 8
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns",
                              values=synth.sipm_waveforms(n, seed=1)["values"].numpy())
    chain, _, tb_out = build_processing_chain(cfg, tables.Table({"waveform": wf}, size=n), db_dict=db, block_width=16, device="meta")
    return chain, [str(pm) for pm in chain._proc_managers], {k: (type(v).__name__, dict(v.attrs)) for k, v in tb_out.items()}


def test_sipm_lar_plan():
    """the reference's SiPM chain: float64 loops forced by the 'd' gaussian kernel, a per-event threshold expression,
    variable-length (VectorOfVectors) outputs whose lengths are other outputs of the chain"""
    chain, procs, cols = sipm_plan(yaml.safe_load(open(SIPM_LAR)))
    assert procs == [
        "reflected_convolve_wf(waveform, gaus_kernel, wf_gaus)", "avg_current(wf_gaus, 5, curr)",
        "histogram(curr, hist_weights, hist_borders)", "histogram_stats(hist_weights, hist_borders, idx_out_c, max_out, fwhm, nan)",
        "multiply(3, fwhm, (3*fwhm))",
        "get_multi_local_extrema(curr, 5, 0.1, 1, (3*fwhm), 0, vt_max_candidate_out, vt_min_out, n_max_out, n_min_out)",
        "peak_snr_threshold(curr, vt_max_candidate_out, 0.8, 10, trigger_pos, no_out)", "multi_a_filter(curr, trigger_pos, energies)"]
    assert cols == {"energies": ("VectorOfVectors", {"units": "ADC"}), "trigger_pos": ("VectorOfVectors", {"units": "ns"})}
    v = chain._vars_dict
    assert str(v["wf_gaus"].dtype) == "float64" and str(v["curr"].dtype) == "float64" and v["curr"].shape == (1995,)
    assert v["gaus_kernel"].is_const and v["gaus_kernel"].shape == (9,)          # const-folded at build time
    assert v["trigger_pos"].vector_len is v["no_out"] and v["energies"].shape == (20,)
    # database override of the kernel width changes the (const-folded) kernel length
    chain2, _, _ = sipm_plan(yaml.safe_load(open(SIPM_LAR)), db={"gauss": {"width": 2, "trunc": 3}})
    assert chain2._vars_dict["gaus_kernel"].shape == (13,)


@pytest.mark.skipif(not os.path.exists(REF_SIPM), reason="reference tree not present (GPU box)")
def test_reference_sipm_config_compiles_to_the_same_plan():
    a = sipm_plan(json.load(open(REF_SIPM)))
    b = sipm_plan(yaml.safe_load(open(SIPM_LAR)))
    assert a[1] == b[1] and a[2] == b[2]


def test_database_overrides_and_errors():
    cfg = yaml.safe_load(open(ICPC))
    _, procs, _, _ = plan(cfg, db={"ttrap": {"rise": "8*us", "flat": "2*us"}, "pz": {"tau": "400*us"}})
    trap = [a for n, a in procs if n.startswith("trap_norm(wf_pz, 8")]
    assert trap and trap[0][1:3] == ["np.int32(500)", "np.int32(125)"]
    assert [a for n, a in procs if n.startswith("pole_zero")][0][1] == "np.float32(25000.0)"
    with pytest.raises(ProcessingChainError):
        plan({"outputs": ["x"], "processors": {"x": {"function": "trap_norm", "module": "dspeed.processors",
                                                      "args": ["waveform", "db.missing", "1*us", "x"]}}})
    with pytest.raises(ProcessingChainError):  # no device implementation -> set-up error, never a CPU fallback
        plan({"outputs": ["x"], "processors": {"x": {"function": "psd", "module": "dspeed.processors",
                                                      "args": ["waveform", "x"]}}})
    with pytest.raises(ProcessingChainError):  # circular reference
        plan({"outputs": ["a"], "processors": {"a": "b + 1", "b": "a + 1"}})


def test_expression_grammar():
    cfg = {
        "outputs": ["e", "late", "n", "sel", "wf_half", "pi2"],
        "processors": {
            "wf_bl": "dspeed.processors.bl_subtract(waveform, baseline, wf_bl(unit='ADC'))",
            "mx": {"function": "numpy.amax(wf_bl[100:2000:2], 1, mx)", "kwargs": {"signature": "(n),()->()", "types": ["fi->f"]}},
            "e": "-mx * 2 + 1",
            "late": "mx > 100",
            "n": "len(wf_bl)",
            "sel": "mx if late else 0",
            "wf_half": "wf_bl[:4096]",
            "pi2": "np.pi * 2",
        },
    }
    chain, procs, cols, _ = plan(cfg)
    names = [n for n, _ in procs]
    assert "amax(wf_bl[100:2000:2], 1, mx)" in names
    assert [a for n, a in procs if n.startswith("amax")][0][0] == "T(16, 950):float32"
    assert any(n.startswith("negative(mx") for n in names) and any(n.startswith("greater(mx") for n in names)
    assert any(n.startswith("where(late, mx, 0") for n in names)
    assert cols["late"][1] == "bool" and cols["wf_half"][1] == "float32"
    assert chain._vars_dict["n"].is_const and chain._vars_dict["pi2"].is_const
    # a strided slice scales the grid period, a start shifts the offset (reference :1032-1054)
    v = chain.get_variable("wf_bl[100:2000:2]")
    assert float(v.grid.period / chain._vars_dict["wf_bl"].grid.period) == 2.0


def test_units_algebra():
    from dspeed_b200.units import Quantity, to_period_units, ureg

    p = Quantity(16.0, "ns")
    assert to_period_units(10 * ureg("us"), p) == 625.0
    assert to_period_units(3.008 * ureg("us"), p) == 188.0
    assert abs(to_period_units(Quantity(2, "MHz"), p) - 0.032) < 1e-15
    assert float((128 * ureg("ns") + 2 * ureg("us")) / p) == 133.0
    assert "ns" in ureg and "ADC" not in ureg
    assert Quantity(1, "us") == Quantity(1000, "ns")


def test_tensor_core_convolution_workspace_query():
    """host-side entry point of csrc/conv_tc.cu (no GPU work): Toeplitz tiles hi + lo of the (K + lead + 127)-wide band
    in 32-column tiles of 128 rows, plus the kernel-sum scalar"""
    import ctypes as C

    from dspeed_b200 import _lib

    L = _lib.lib()
    L.dspb_convolve_tc_workspace.restype = C.c_int64
    for K in (1, 128, 1024, 4096, 5792):
        nk = (K + 3 + 127 + 31) // 32
        assert L.dspb_convolve_tc_workspace(C.c_int64(K)) == 2 * nk * 128 * 32 + 32
