"""The IIR family built on recursive_filter (pole_zero.py:201-342 convolve_exp / convolve_damped_oscillator /
inject_damped_oscillation, rc_cr2.py, the iir_filter.py factories) against vectors recorded from the reference
(tests/golden/iir_family.npz): the CPU oracle's recursion on CPU, the device processors on the GPU."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "iir_family.npz"))
DT = {"f": np.float32, "d": np.float64}


def rc_exp(tau):
    return np.exp(-1.0 / tau) if tau != 0 else 0.0


def designs():
    import scipy.signal as sg

    return {"butter4_lp": (*sg.iirfilter(4, 0.1, btype="lowpass", ftype="butter"), None),
            "cheby1_3_hp": (*sg.iirfilter(3, 0.2, rp=1.0, btype="highpass", ftype="cheby1"), None),
            "butter2_bp": (*sg.iirfilter(2, [0.05, 0.2], btype="bandpass", ftype="butter"), None),
            "notch": (*sg.iirnotch(0.24, 10.0), 1.0), "peak": (*sg.iirpeak(0.24, 10.0), 0.0)}


def close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.abs(a - b).max() <= rtol * np.abs(b).max()


@pytest.mark.parametrize("t", ["f", "d"])
def test_oracle_recursion_reproduces_the_reference(t):
    from oracle import oracle as O

    dt = DT[t]
    w = G["values"].astype(dt) - G["baseline"].astype(dt)[:, None]
    rt = 1e-6 if t == "f" else 1e-13
    for tau in (50.0, 400.5):
        close(O.recursive_filter(w, [1.0], [1.0, -rc_exp(float(dt(tau)))], w[:, 0], w[:, 0]), G[f"cexp_{t}_{tau}"], rt)
    rc = rc_exp(120.0)
    close(O.recursive_filter(w, [np.cos(0.7), -rc * np.cos(0.3 - 0.7)], [1, -2 * rc * np.cos(0.3), rc * rc], w[:, 0], w[:, 0]), G[f"cdo_{t}"], rt)
    for tag, (a, b, gain) in designs().items():
        gain = sum(a) / sum(b) if gain is None else gain
        close(O.recursive_filter(w, a, b, w[:, 0], (gain * w[:, 0]).astype(dt)), G[f"iir_{tag}_{t}"], rt)


@pytest.mark.gpu
@pytest.mark.parametrize("t", ["f", "d"])
def test_device_iir_family(t):
    import dspeed_b200.processors as P
    from dspeed_b200.errors import DSPFatal

    dt = DT[t]
    w = G["values"].astype(dt) - G["baseline"].astype(dt)[:, None]
    rt = 2e-6 if t == "f" else 1e-12
    for tau in (50.0, 400.5):
        o = np.zeros_like(w)
        P.convolve_exp(w, dt(tau), o)
        close(o, G[f"cexp_{t}_{tau}"], rt)
        o = np.zeros_like(w)
        P.rc_cr2(w, dt(tau), o)
        close(o, G[f"rccr2_{t}_{tau}"], rt)
    o = np.zeros_like(w)
    P.convolve_damped_oscillator(w, 120.0, 0.3, 0.7, o)
    close(o, G[f"cdo_{t}"], rt)
    o = np.zeros_like(w)
    P.inject_damped_oscillation(w, 120.0, 0.3, 0.7, 0.05, o)
    close(o, G[f"ido_{t}"], rt)
    with pytest.raises(DSPFatal):
        P.inject_damped_oscillation(w, 120.0, 0.3, 0.7, 1.5, o)
    procs = {"butter4_lp": P.iir_filter(0.1, 4), "cheby1_3_hp": P.iir_filter(0.2, 3, rp=1.0, ftype="cheby1", btype="highpass"),
             "butter2_bp": P.iir_filter([0.05, 0.2], 2, btype="bandpass"), "notch": P.notch_filter(0.24, 0.024),
             "peak": P.peak_filter(0.24, 0.024)}
    for tag, proc in procs.items():
        o = np.zeros_like(w)
        proc(w, o)
        close(o, G[f"iir_{tag}_{t}"], rt)
    with pytest.raises(DSPFatal):
        P.iir_filter(1.5, 2)
    # order-4 recursion straight through recursive_filter (beyond the order-2 affine scan)
    a, b, _ = designs()["butter4_lp"]
    o = np.zeros_like(w)
    P.recursive_filter(w, a, b, w[:, 0].copy(), ((sum(a) / sum(b)) * w[:, 0]).astype(dt), o)
    close(o, G[f"iir_butter4_lp_{t}"], rt)
