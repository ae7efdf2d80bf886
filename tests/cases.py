"""Shared case table: the per-processor cases of tests/golden/processors.npz
(generated from the reference by oracle/gen_golden.py), expressed once so the CPU
oracle (-m "not gpu") and the CUDA processors (-m gpu) run exactly the same inputs.

A case is (golden_key_prefix, processor_name, inputs, out_specs) where ``inputs``
are the gufunc input arguments in the reference's order and ``out_specs`` is a list
of (golden key, core shape, dtype-or-None) for the outputs."""

from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@dataclass
class Case:
    name: str
    proc: str
    inputs: list
    outs: list  # [(golden key, core shape tuple, dtype or None)]
    exact: bool = True
    rtol: float = 0.0  # relative to max|golden| when not exact
    extra: dict = field(default_factory=dict)


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def processor_cases(tags=("f", "d")):
    g = load("processors.npz")
    w32 = g["base/w"]
    cases = []
    for tag in tags:
        arr = w32 if tag == "f" else w32.astype(np.float64)
        dt = arr.dtype.type
        n = arr.shape[1]
        add = lambda *a, **k: cases.append(Case(*a, **k))  # noqa: E731
        add(f"bl_subtract_{tag}", "bl_subtract", [arr, dt(3.25)], [(f"bl_subtract_{tag}/out", (n,), None)])
        add(f"linear_slope_fit_{tag}", "linear_slope_fit", [arr], [(f"linear_slope_fit_{tag}/out{i}", (), None) for i in range(4)])
        add(f"linear_slope_diff_{tag}", "linear_slope_diff", [arr, g[f"linear_slope_fit_{tag}/out2"], g[f"linear_slope_fit_{tag}/out3"]],
            [(f"linear_slope_diff_{tag}/out{i}", (), None) for i in range(2)])
        add(f"mean_below_threshold_{tag}", "mean_below_threshold", [arr, dt(0.5)], [(f"mean_below_threshold_{tag}/out", (), None)])
        add(f"pole_zero_{tag}", "pole_zero", [arr, dt(45.5)], [(f"pole_zero_{tag}/out", (n,), None)])
        add(f"double_pole_zero_{tag}", "double_pole_zero", [arr, dt(45.5), dt(7.25), dt(0.03)], [(f"double_pole_zero_{tag}/out", (n,), None)])
        for r, f in ((10, 4), (1, 0), (50, 100), (0, 5), (3, 0)):
            add(f"trap_filter_{tag}_{r}_{f}", "trap_filter", [arr, r, f], [(f"trap_filter_{tag}_{r}_{f}/out", (n,), None)])
            if r > 0:
                add(f"trap_norm_{tag}_{r}_{f}", "trap_norm", [arr, r, f], [(f"trap_norm_{tag}_{r}_{f}/out", (n,), None)])
        for r, f, fl in ((8, 4, 125), (10, 0, 10), (1, 1, 1), (100, 50, 150)):
            add(f"asym_trap_filter_{tag}_{r}_{f}_{fl}", "asym_trap_filter", [arr, r, f, fl], [(f"asym_trap_filter_{tag}_{r}_{f}_{fl}/out", (n,), None)])
        for r, f, t in ((10, 4, 100.0), (10, 4, 22.0), (10, 4, 23.0), (10, 4, 299.0), (10, 4, 300.0), (50, 100, 250.0)):
            add(f"trap_pickoff_{tag}_{r}_{f}_{int(t)}", "trap_pickoff", [arr, r, f, dt(t)], [(f"trap_pickoff_{tag}_{r}_{f}_{int(t)}/out", (), None)])
        for L in (1, 2, 7, 48, 299, 10.5):
            add(f"moving_window_left_{tag}_{L}", "moving_window_left", [arr, dt(L)], [(f"moving_window_left_{tag}_{L}/out", (n,), None)])
            add(f"moving_window_right_{tag}_{L}", "moving_window_right", [arr, dt(L)], [(f"moving_window_right_{tag}_{L}/out", (n,), None)])
        for L, num, typ in ((48, 3, 0), (5, 1, 0), (5, 2, 1), (5, 4, 2), (16, 0, 0), (1, 3, 0)):
            add(f"moving_window_multi_{tag}_{L}_{num}_{typ}", "moving_window_multi", [arr, dt(L), dt(num), typ],
                [(f"moving_window_multi_{tag}_{L}_{num}_{typ}/out", (n,), None)])
        for L in (1, 3, 100):
            add(f"avg_current_{tag}_{L}", "avg_current", [arr, dt(L)], [(f"avg_current_{tag}_{L}/out", (n - L,), None)])
        thr, ts = g[f"tpt_{tag}/thr"], g[f"tpt_{tag}/ts"]
        for wf in (0, 1):
            add(f"time_point_thresh_{tag}_{wf}", "time_point_thresh", [arr, thr, ts, wf], [(f"time_point_thresh_{tag}_{wf}/out", (), None)])
            for mode in "iafbcrnl":
                add(f"interpolated_time_point_thresh_{tag}_{wf}_{mode}", "interpolated_time_point_thresh", [arr, thr, ts, wf, mode],
                    [(f"interpolated_time_point_thresh_{tag}_{wf}_{mode}/out", (), None)])
        mthr = g[f"mtpt_{tag}/thr"]
        for pol in (1, -1):
            for mode in "iafbcrnl":
                add(f"multi_time_point_thresh_{tag}_{pol}_{mode}", "multi_time_point_thresh", [arr, mthr, ts, pol, mode],
                    [(f"multi_time_point_thresh_{tag}_{pol}_{mode}/out", (4,), None)])
        tpk = g[f"ftp_{tag}/t"]
        for mode in "nfclhs":
            add(f"fixed_time_pickoff_{tag}_{mode}", "fixed_time_pickoff", [arr, tpk, mode], [(f"fixed_time_pickoff_{tag}_{mode}/out", (), None)])
        add(f"fixed_time_pickoff_{tag}_i", "fixed_time_pickoff", [arr, np.array([3, 0, 299, 300, -1], dt), "i"], [(f"fixed_time_pickoff_{tag}_i/out", (), None)])
        add(f"min_max_{tag}", "min_max", [arr], [(f"min_max_{tag}/out{i}", (), None) for i in range(4)])
        add(f"min_max_norm_{tag}", "min_max_norm", [arr, g[f"min_max_{tag}/out2"], g[f"min_max_{tag}/out3"]], [(f"min_max_norm_{tag}/out", (n,), None)])
        add(f"windower_{tag}", "windower", [arr, g[f"windower_{tag}/t0"]], [(f"windower_{tag}/out", (101,), None)])
        for up, m in ((16, 4784), (4, 1000), (3, 950), (2.5, 700)):
            add(f"upsampler_{tag}_{up}_{m}", "upsampler", [arr, dt(up)], [(f"upsampler_{tag}_{up}_{m}/out", (m,), None)])
        kern = g[f"conv_{tag}/kernel"]
        tol = 2e-6 if tag == "f" else 1e-13
        for mode, m in (("f", n + 32), ("v", n - 32), ("s", n)):
            add(f"convolve_wf_{tag}_{mode}", "convolve_wf", [arr, kern, mode], [(f"convolve_wf_{tag}_{mode}/out", (m,), None)], exact=False, rtol=tol)
            add(f"fft_convolve_wf_{tag}_{mode}", "fft_convolve_wf", [arr, kern, mode], [(f"fft_convolve_wf_{tag}_{mode}/out", (m,), None)], exact=False, rtol=tol)
        add(f"convolve_wf_{tag}_s_even", "convolve_wf", [arr, g[f"conv_{tag}/kernel_even"], "s"], [(f"convolve_wf_{tag}_s_even/out", (n,), None)], exact=False, rtol=tol)
        add(f"recursive_filter_{tag}", "recursive_filter", [arr, np.array([1.0, -0.5]), np.array([1.0, -0.9, 0.1]), dt(0.5), dt(-0.25)],
            [(f"recursive_filter_{tag}/out", (n,), None)])
        for sd in (0, 1, 2, 3):
            for dmax, dmin, amax, amin in ((8.0, 8.0, -1e9, 1e9), (3.0, 1.0, 0.0, 20.0), (20.0, 5.0, 10.0, 0.0)):
                k = f"get_multi_local_extrema_{tag}_{sd}_{dmax}_{dmin}_{amax}_{amin}"
                add(k, "get_multi_local_extrema", [arr, dmax, dmin, sd, amax, amin],
                    [(k + "/vmax", (6,), None), (k + "/vmin", (6,), None), (k + "/nmax", (), np.uint32), (k + "/nmin", (), np.uint32)])
    return g, cases


def run_oracle(case: Case):
    """Evaluate a case with the CPU oracle; returns a tuple of outputs."""
    from oracle import oracle as O

    fn = getattr(O, case.proc)
    args = list(case.inputs)
    # oracle functions that need the output length
    if case.proc in ("windower", "upsampler", "get_multi_local_extrema"):
        args.append(case.outs[0][1][0])
    res = fn(*args)
    return res if isinstance(res, tuple) else (res,)


def compare(case: Case, golden, results, extra_rtol: float = 0.0, force_tol: bool = False):
    """Assert results against the golden arrays of a case.  ``extra_rtol`` > 0 relaxes
    an exact case to a tolerance relative to max|golden| (used by the CUDA tests for
    float outputs, where the reference's own sequential-float32 rounding is not
    reproducible)."""
    for (key, _, _), got in zip(case.outs, results):
        ref = golden[key]
        got = np.asarray(got)
        assert got.shape == ref.shape, (key, got.shape, ref.shape)
        assert np.array_equal(np.isnan(ref), np.isnan(got)), f"{key}: NaN pattern differs"
        if case.exact and not force_tol:
            assert np.array_equal(ref, got, equal_nan=True), f"{key}: not bit-exact"
        else:
            tol = max(case.rtol, extra_rtol)
            fin = np.isfinite(ref)
            if fin.any():
                scale = np.abs(ref[fin]).max()
                err = np.abs(ref[fin].astype(np.float64) - got[fin].astype(np.float64)).max()
                assert err <= tol * max(scale, 1e-30), f"{key}: err {err:.3e} > {tol:.1e} * {scale:.3e}"
            assert np.array_equal(np.isinf(ref), np.isinf(got)), f"{key}: inf pattern differs"
