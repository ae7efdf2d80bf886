"""Parity of the CUDA processors (through the C ABI) with the reference, on the golden
cases recorded from the reference's own processors (tests/golden/processors.npz).

Bit-exact: index / integer / min-max / copy-type outputs.  Float outputs of processors
whose reference implementation is a sequential float32 recursion are compared to
1e-5 of the golden array's maximum (the reference's own rounding drift is 2-4e-6,
SURVEY.md fact 5; the CUDA kernels accumulate in float64 and round once)."""

import numpy as np
import pytest

from tests import cases as C

pytestmark = pytest.mark.gpu

GOLD, CASES = C.processor_cases()

# processors whose CUDA result must equal the reference bit for bit
EXACT = {
    "bl_subtract", "min_max", "min_max_norm", "avg_current", "time_point_thresh",
    "interpolated_time_point_thresh", "multi_time_point_thresh", "windower", "upsampler",
    "get_multi_local_extrema",
}
EXACT_FTP_MODES = {"n", "f", "c", "i"}
FLOAT_RTOL = 1e-5
F64_RTOL = 1e-11


def run_cuda(case):
    import dspeed_b200.processors as P

    proc = getattr(P, case.proc)
    nr = case.inputs[0].shape[0]
    dt = case.inputs[0].dtype
    outs = []
    for key, core, odt in case.outs:
        o = np.full((nr,) + tuple(core), 7, dtype=odt or dt)
        outs.append(o)
    proc(*case.inputs, *outs)
    return outs


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_cuda_processor_matches_reference(case):
    is_f64 = case.inputs[0].dtype == np.float64
    exact = case.proc in EXACT or (case.proc == "fixed_time_pickoff" and case.inputs[2] in EXACT_FTP_MODES)
    # degenerate arguments for which the reference itself produces garbage (division by a
    # zero-length window); only the NaN pattern is not even stable there
    if case.proc in ("trap_filter", "trap_norm") and case.inputs[1] == 0:
        pytest.skip("rise == 0: reference reads w_out[-1] (uninitialised NaN)")
    res = run_cuda(case)
    if exact:
        C.compare(case, GOLD, res)
    else:
        C.compare(case, GOLD, res, extra_rtol=F64_RTOL if is_f64 else FLOAT_RTOL, force_tol=True)


def test_fatal_conditions():
    import dspeed_b200.processors as P
    from dspeed_b200.errors import DSPFatal

    w = np.array([-1, 0, 1, 2, 3, 4, -1, 0, 1, 2, 3, 4], np.float32)[None, :]
    out = np.zeros(1, np.float32)
    # reference tests/processors/test_time_point_thresh.py:66-78
    with pytest.raises(DSPFatal):
        P.time_point_thresh(w, 1.0, 10.5, 0, out)
    with pytest.raises(DSPFatal):
        P.time_point_thresh(w, 1.0, 11, 0.5, out)
    with pytest.raises(DSPFatal):
        P.time_point_thresh(w, 1.0, 12, 0, out)
    # test_fixed_time_pickoff.py:34-47
    with pytest.raises(DSPFatal):
        P.fixed_time_pickoff(w, 1.5, "i", out)
    with pytest.raises(DSPFatal):
        P.fixed_time_pickoff(w, 1.5, " ", out)
    wo = np.zeros_like(w)
    with pytest.raises(DSPFatal):
        P.trap_norm(w, 5, 3, wo)
    with pytest.raises(DSPFatal):
        P.double_pole_zero(w[:, :2], 10.0, 20.0, 0.1, wo[:, :2])


def test_reference_known_answers():
    """Known-answer vectors of the reference's own unit tests."""
    import dspeed_b200.processors as P

    f = np.float32
    w = np.array([-1, 0, 1, 2, 3, 4, -1, 0, 1, 2, 3, 4], f)[None, :]
    out = np.zeros(1, f)
    P.time_point_thresh(w, 1.0, 11, 0, out); assert out[0] == 8.0
    P.time_point_thresh(w, 3.0, 0, 1, out); assert out[0] == 4.0
    P.time_point_thresh(np.array([[5, 4, 3, 2, 1, 0, -1]], f), 2.5, 0, 1, out); assert out[0] == 2.0
    P.time_point_thresh(np.array([[0, 1, 2, 3, 4, 5]], f), 2.5, 0, 1, out); assert out[0] == 2.0
    for args, exp in (((1, 11, 0, "i"), 7), ((3, 0, 1, "i"), 4), ((1, 11, 0, "f"), 8), ((3, 0, 1, "f"), 5),
                      ((1, 11, 0, "c"), 7), ((3, 0, 1, "c"), 4), ((1, 11, 0, "n"), 7.5), ((3, 0, 1, "n"), 4.5),
                      ((1.5, 11, 0, "l"), 8.5), ((3.5, 0, 1, "l"), 4.5)):
        P.interpolated_time_point_thresh(w, *args, out)
        assert out[0] == exp, (args, out[0])
    P.interpolated_time_point_thresh(w, 1, 12, 0, "i", out); assert np.isnan(out[0])
    # fixed_time_pickoff (test_fixed_time_pickoff.py:49-90)
    r = np.arange(20, dtype=f)[None, :]
    for t, m, exp in ((3, "i", 3), (3.5, "n", 4), (3.5, "f", 3), (3.5, "c", 4), (3.5, "l", 3.5), (3.5, "h", 3.5), (3.5, "s", 3.5)):
        P.fixed_time_pickoff(r, t, m, out); assert out[0] == exp, (t, m, out[0])
    s = np.sin(np.arange(20)).astype(np.float64)[None, :]
    o64 = np.zeros(1)
    for m, exp in (("n", 0.1411200080598672), ("f", 0.1411200080598672), ("c", -0.7568024953079282),
                   ("l", -0.08336061778208165), ("h", -0.09054574599004982), ("s", -0.10707938709427486)):
        P.fixed_time_pickoff(s, 3.25, m, o64); assert abs(o64[0] - exp) < 1e-12, (m, o64[0])
    P.fixed_time_pickoff(s, 0.2, "h", o64); assert abs(o64[0] - 0.1806725096462211) < 1e-12
    P.fixed_time_pickoff(s, 18.2, "h", o64); assert abs(o64[0] + 0.6150034250096629) < 1e-12
    for t in (-1, 20, np.nan):
        P.fixed_time_pickoff(r, t, "l", out); assert np.isnan(out[0])
    # pole_zero (test_pole_zero.py:14-48)
    tt = np.arange(8192, dtype=np.float64)
    wexp = np.concatenate([np.zeros(20), 17500 * np.exp(-tt / 30000)])[None, :]
    for dt, rtol in ((np.float32, 1e-6), (np.float64, 1e-7)):
        wo = np.zeros_like(wexp, dtype=dt)
        P.pole_zero(wexp.astype(dt), 30000, wo)
        exp = np.concatenate([np.zeros(20), np.full(8192, 17500.0)])
        assert np.allclose(wo[0], exp, rtol=rtol, atol=17500 * rtol)
    # double_pole_zero (test_pole_zero.py:51-96)
    wexp = np.concatenate([np.zeros(20), 17500 * (0.02 * np.exp(-tt / 1000) + 0.98 * np.exp(-tt / 30000))])[None, :8192]
    for dt, rtol in ((np.float32, 1e-6), (np.float64, 1e-7)):
        wo = np.zeros_like(wexp, dtype=dt)
        P.double_pole_zero(wexp.astype(dt), 1000, 30000, 0.98, wo)
        exp = np.concatenate([np.zeros(20), np.full(8192, 17500.0)])[:8192]
        assert np.allclose(wo[0], exp, rtol=rtol, atol=17500 * rtol * 3)
    # min_max ties -> first index (min_max.py:73-77)
    o = [np.zeros(1, f) for _ in range(4)]
    P.min_max(np.array([[5, 9, 9, 1, 1, 7]], np.uint16), *o)
    assert [x[0] for x in o] == [3, 1, 1, 9]
    # get_multi_local_extrema (test_get_multi_local_extrema.py:22-301)
    tri = np.array([0, 0, 1, 2, 3, 4, 5, 4, 3, 4, 5, 6, 7, 8, 9, 10, 9, 8, 7, 6, 5, 4, 3, 4, 5, 4, 3, 2, 1, 0, 0], f)[None, :]
    nan = np.nan
    for args, emax, emin, nmx, nmn in (
        ((3, 3, 0, 0, 20), [15, nan, nan], [nan, nan, nan], 1, 0),
        ((3, 1, 0, 0, 20), [15, 24, nan], [22, nan, nan], 2, 1),
        ((3, 1, 1, 0, 20), [15, 6, nan], [8, nan, nan], 2, 1),
        ((3, 1, 2, 0, 20), [15, nan, nan], [nan, nan, nan], 1, 0),
        ((3, 1, 3, 0, 20), [6, 15, 24], [8, 22, nan], 3, 2),
        ((3, 1, 3, 8, 20), [15, nan, nan], [8, 22, nan], 1, 2),
    ):
        vmax, vmin = np.zeros((1, 3), f), np.zeros((1, 3), f)
        a, b = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        P.get_multi_local_extrema(tri, *args, vmax, vmin, a, b)
        assert np.array_equal(vmax[0], np.array(emax, f), equal_nan=True), (args, vmax)
        assert np.array_equal(vmin[0], np.array(emin, f), equal_nan=True), (args, vmin)
        assert (a[0], b[0]) == (nmx, nmn), (args, a, b)


def test_kernel_generators():
    import dspeed_b200.processors as P

    k = C.load("kernels.npz")
    prm = {5792: (1250.0, 188.0, 28125.0), 301: (100.5, 10.0, 500.0), 64: (20.0, 0.0, 100.0)}
    for key in k.files:
        parts = key.split("_")
        dt = np.float32 if parts[1] == "f" else np.float64
        ref = k[key]
        got = np.zeros_like(ref)
        if parts[0] == "cusp":
            P.cusp_filter(*[dt(x) for x in prm[int(parts[2])]], got)
        elif parts[0] == "zac":
            P.zac_filter(*[dt(x) for x in prm[int(parts[2])]], got)
        elif parts[0] == "t0":
            P.t0_filter(dt(int(parts[2])), dt(int(parts[3])), got)
        else:
            continue
        # the kernels are differences of O(1) shape values: device float64 sinh/exp may
        # differ from libm by an ulp or two of 1.0 (2e-16); the float32 kernels are rounded
        # from the float64 shape *before* the differencing, so they agree essentially
        # bit for bit (<= 4e-7 of the largest tap = 3e-10 absolute for the ICPC cusp)
        err = np.abs(got - ref).max()
        if dt == np.float32:
            assert err <= 4e-7 * np.abs(ref).max(), (key, err, np.abs(ref).max())
        else:
            assert err <= 2e-15, (key, err)
