"""The processors of the reference's SiPM chain (csrc/sipm.cu through the C ABI) against the vectors recorded from
the reference's numba / numpy processors (tests/golden/sipm_processors.npz): histogram weights, borders, indices,
counts and pass-through amplitudes bit-for-bit; convolutions / kernels to float rounding."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sipm_processors.npz"))
DT = {"f": np.float32, "d": np.float64}


def P():
    import dspeed_b200.processors as p

    return p


def close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(b)
    if ok.any():
        assert np.abs(a[ok] - b[ok]).max() <= rtol * max(np.abs(b[ok]).max(), 1e-30)


@pytest.mark.parametrize("t", ["f", "d"])
def test_gaussian_kernel_and_reflected_convolution(t):
    dt = DT[t]
    for sig, trunc in ((1.0, 4.0), (2.5, 3.0)):
        ref = G[f"gaus_{t}_{sig}_{trunc}"]
        k = np.zeros(len(ref), dt)
        P().gaussian_filter1d(dt(sig), dt(trunc), k)
        close(k, ref, 2e-7 if t == "f" else 1e-15)
    w = G["values"].astype(dt)
    for tag, kname in (("wf_gaus", "1.0_4.0"), ("wf_gaus2", "2.5_3.0")):
        out = np.zeros_like(w)
        P().reflected_convolve_wf(w, G[f"gaus_{t}_{kname}"], out)
        close(out, G[f"{tag}_{t}"], 1e-6 if t == "f" else 1e-14)
    # uint16 input rows are converted on load like numpy casts them into the float loop
    out = np.zeros_like(w)
    P().reflected_convolve_wf(G["values"], G[f"gaus_{t}_1.0_4.0"], out)
    close(out, G[f"wf_gaus_{t}"], 1e-6 if t == "f" else 1e-14)
    # NaN in the waveform -> NaN row
    wn = w[:2].copy()
    wn[1, 7] = np.nan
    out = np.zeros_like(wn)
    P().reflected_convolve_wf(wn, G[f"gaus_{t}_1.0_4.0"], out)
    assert np.isnan(out[1]).all() and not np.isnan(out[0]).any()


@pytest.mark.parametrize("t", ["f", "d"])
def test_histogram_and_stats(t):
    dt = DT[t]
    curr = G[f"curr_{t}"]
    n = len(curr)
    hw, hb = np.zeros((n, 100), dt), np.zeros((n, 101), dt)
    P().histogram(curr, hw, hb)
    assert np.array_equal(hw, G[f"hist_w_{t}"]) and np.array_equal(hb, G[f"hist_b_{t}"])
    rw, rb = np.zeros((n, 40), dt), np.zeros((n, 41), dt)
    P().histogram(G["values"].astype(dt), rw, rb)
    assert np.array_equal(rw, G[f"hist_raw_w_{t}"]) and np.array_equal(rb, G[f"hist_raw_b_{t}"])
    for tag, mx in (("hs", np.nan), ("hs2", 0.5)):
        idx, m, fw = np.zeros(n, dt), np.zeros(n, dt), np.zeros(n, dt)
        P().histogram_stats(hw, hb, idx, m, fw, dt(mx))
        assert np.array_equal(idx, G[f"{tag}_idx_{t}"], equal_nan=True)
        assert np.array_equal(m, G[f"{tag}_max_{t}"], equal_nan=True)
        assert np.array_equal(fw, G[f"{tag}_fwhm_{t}"], equal_nan=True)
    # NaN sample: zero weights, NaN borders (histogram.py:69-72); constant row: all-zero weights (delta == 0)
    x = curr[:2].copy()
    x[0, 3] = np.nan
    x[1, :] = 2.5
    hw2, hb2 = np.ones((2, 10), dt), np.ones((2, 11), dt)
    P().histogram(x, hw2, hb2)
    assert (hw2 == 0).all() and np.isnan(hb2[0]).all() and (hb2[1] == 2.5).all()


@pytest.mark.parametrize("t", ["f", "d"])
def test_histogram_around_mode_and_peakstats(t):
    from dspeed_b200.errors import DSPFatal

    dt = DT[t]
    curr = G[f"curr_{t}"]
    n = len(curr)
    for tag, center, bw, nb in (("a", np.nan, 1.0, 101), ("b", np.nan, 0.5, 64), ("c", 3.0, 2.0, 31)):
        aw, ab = np.zeros((n, nb), dt), np.zeros((n, nb + 1), dt)
        P().histogram_around_mode(curr, dt(center), dt(bw), aw, ab)
        assert np.array_equal(aw, G[f"ham_w_{tag}_{t}"]), tag
        assert np.array_equal(ab, G[f"ham_b_{tag}_{t}"]), tag
        for skip in (0, 1):
            for wt in range(5):
                mo, wo = np.zeros(n, dt), np.zeros(n, dt)
                P().histogram_peakstats(aw, ab, dt(np.nan), np.int32(skip), np.int32(wt), mo, wo)
                assert np.array_equal(mo, G[f"hps_mode_{tag}_{skip}_{wt}_{t}"], equal_nan=True)
                assert np.array_equal(wo, G[f"hps_width_{tag}_{skip}_{wt}_{t}"], equal_nan=True), (tag, skip, wt)
        mo, wo = np.zeros(n, dt), np.zeros(n, dt)
        P().histogram_peakstats(aw, ab, dt(1.25), np.int32(0), np.int32(0), mo, wo)
        assert np.array_equal(mo, G[f"hps_mode_{tag}_given_{t}"], equal_nan=True)
        assert np.array_equal(wo, G[f"hps_width_{tag}_given_{t}"], equal_nan=True)
    with pytest.raises(DSPFatal):    # length mismatch (histogram.py:152-153)
        P().histogram_around_mode(curr, dt(np.nan), dt(1), np.zeros((n, 10), dt), np.zeros((n, 10), dt))
    with pytest.raises(DSPFatal):    # unknown width type (histogram_stats.py:142-143)
        P().histogram_peakstats(aw, ab, dt(np.nan), np.int32(0), np.int32(7), np.zeros(n, dt), np.zeros(n, dt))
    bad = curr[:1].copy()
    bad[0, 0] = np.nan
    with pytest.raises(DSPFatal):    # histogram.py:157-158
        P().histogram_around_mode(bad, dt(np.nan), dt(1), np.zeros((1, 11), dt), np.zeros((1, 12), dt))


def test_reference_histogram_known_answers():
    """tests/processors/test_histogram.py:9-60 of the reference"""
    from dspeed_b200.errors import DSPFatal

    vals = np.arange(100) * 2 / 3
    with pytest.raises(DSPFatal):
        P().histogram(vals, np.zeros(10), np.zeros(10))
    hw, he = np.zeros(66), np.zeros(67)
    P().histogram(vals, hw, he)
    assert all(he == np.arange(67)) and all(hw[0::2] == 2) and all(hw[1::2] == 1)
    for tag in "abc":
        aw, ab = np.zeros(11, np.float32), np.zeros(12, np.float32)
        P().histogram_around_mode(G[f"kat_ham_in_{tag}"], np.float32(np.nan), np.float32(1.0), aw, ab)
        assert np.array_equal(aw, G[f"kat_ham_w_{tag}"]) and np.array_equal(ab, G[f"kat_ham_b_{tag}"])


@pytest.mark.parametrize("t", ["f", "d"])
def test_peak_selection_and_amplitudes(t):
    dt = DT[t]
    curr, vmax = G[f"curr_{t}"], G[f"vt_max_{t}"]
    n = len(curr)
    for tag, ratio, width in (("a", 0.8, 10), ("b", 0.3, 4)):
        trig, no = np.zeros((n, 20), dt), np.zeros(n, np.uint32)
        P().peak_snr_threshold(curr, vmax, dt(ratio), dt(width), trig, no)
        assert np.array_equal(trig, G[f"trig_{tag}_{t}"], equal_nan=True)
        assert np.array_equal(no, G[f"n_trig_{tag}_{t}"])
        en = np.zeros((n, 20), dt)
        P().multi_a_filter(curr, trig, en)
        assert np.array_equal(en, G[f"energies_{tag}_{t}"], equal_nan=True)


@pytest.mark.parametrize("t", ["f", "d"])
def test_dplms_and_small_kernels(t):
    """tests/processors/test_dplms.py:10-35 of the reference: its 50 x 50 noise matrix, fatals, delta / pulse references"""
    from dspeed_b200.errors import DSPFatal

    dt = DT[t]
    nmat, ref = G["dplms_nmat"].astype(dt), G["dplms_ref"].astype(dt)
    k = np.zeros(50, dt)
    for args in (([], 1, 1, 1, 1), (ref, -1, 1, 1, 1), (ref, 1, -1, 1, 1), (ref, 1, 1, -1, 1), (ref, 1, 1, 1, -1), (ref, 1, 1, 1, 2)):
        with pytest.raises(DSPFatal):
            P().dplms(nmat, np.asarray(args[0], dt), *[dt(x) for x in args[1:]], k)
    P().dplms(nmat, ref, dt(1), dt(1), dt(1), dt(1), k)
    close(k, G[f"dplms_{t}_1111"], 3e-5 if t == "f" else 1e-9)
    P().dplms(nmat, G["dplms_ref_pulse"].astype(dt), dt(50), dt(0.1), dt(1), dt(1), k)
    close(k, G[f"dplms_{t}_pulse"], 3e-5 if t == "f" else 1e-9)
    K = np.load(os.path.join(os.path.dirname(__file__), "golden", "kernels.npz"))
    for n in (5, 32):
        k = np.zeros(n, dt)
        P().moving_slope(k)
        close(k, K[f"moving_slope_{t}_{n}"], 2e-7 if t == "f" else 1e-15)
        k = np.zeros(n, dt)
        P().step(dt(1), k)
        assert np.array_equal(k, K[f"step_{t}_{n}"])
