"""The checker's restatement of the SiPM-chain processors (oracle/sipm_oracle.py) against the vectors recorded from
the reference's own numba / numpy processors (tests/golden/sipm_processors.npz, oracle/gen_golden.py): histogram
weights, indices and counts bit-for-bit, floats to rounding."""
import os

import numpy as np
import pytest

from oracle import sipm_oracle as S

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sipm_processors.npz"))
DT = {"f": np.float32, "d": np.float64}


def close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = np.isfinite(b)
    scale = max(np.abs(b[ok]).max(), 1e-30) if ok.any() else 1.0
    assert np.abs(a[ok] - b[ok]).max() <= rtol * scale if ok.any() else True


@pytest.mark.parametrize("t", ["f", "d"])
def test_gaussian_and_reflected_convolution(t):
    for sig, trunc in ((1.0, 4.0), (2.5, 3.0)):
        k = S.gaussian_filter1d(sig, trunc, DT[t])
        close(k, G[f"gaus_{t}_{sig}_{trunc}"], 1e-7 if t == "f" else 1e-15)
    w = G["values"].astype(DT[t])
    close(S.reflected_convolve_wf(w, G[f"gaus_{t}_1.0_4.0"]), G[f"wf_gaus_{t}"], 1e-6 if t == "f" else 1e-14)
    close(S.reflected_convolve_wf(w, G[f"gaus_{t}_2.5_3.0"]), G[f"wf_gaus2_{t}"], 1e-6 if t == "f" else 1e-14)


@pytest.mark.parametrize("t", ["f", "d"])
def test_histogram_and_stats(t):
    curr = G[f"curr_{t}"]
    hw, hb = S.histogram(curr, 100)
    assert np.array_equal(hw, G[f"hist_w_{t}"])
    assert np.array_equal(hb, G[f"hist_b_{t}"])
    rw, rb = S.histogram(G["values"].astype(DT[t]), 40)
    assert np.array_equal(rw, G[f"hist_raw_w_{t}"]) and np.array_equal(rb, G[f"hist_raw_b_{t}"])
    for tag, mx in (("hs", np.nan), ("hs2", 0.5)):
        idx, m, fw = S.histogram_stats(G[f"hist_w_{t}"], G[f"hist_b_{t}"], mx)
        assert np.array_equal(idx, G[f"{tag}_idx_{t}"], equal_nan=True)
        assert np.array_equal(m, G[f"{tag}_max_{t}"], equal_nan=True)
        assert np.array_equal(fw, G[f"{tag}_fwhm_{t}"], equal_nan=True)


@pytest.mark.parametrize("t", ["f", "d"])
def test_histogram_around_mode_and_peakstats(t):
    curr = G[f"curr_{t}"]
    for tag, center, bw, nb in (("a", np.nan, 1.0, 101), ("b", np.nan, 0.5, 64), ("c", 3.0, 2.0, 31)):
        aw, ab = S.histogram_around_mode(curr, center, bw, nb)
        assert np.array_equal(aw, G[f"ham_w_{tag}_{t}"]), tag
        assert np.array_equal(ab, G[f"ham_b_{tag}_{t}"]), tag
        for skip in (0, 1):
            for wt in range(5):
                mo, wo = S.histogram_peakstats(aw, ab, np.nan, skip, wt)
                assert np.array_equal(mo, G[f"hps_mode_{tag}_{skip}_{wt}_{t}"], equal_nan=True)
                assert np.array_equal(wo, G[f"hps_width_{tag}_{skip}_{wt}_{t}"], equal_nan=True), (tag, skip, wt)
        mo, wo = S.histogram_peakstats(aw, ab, 1.25, 0, 0)
        assert np.array_equal(mo, G[f"hps_mode_{tag}_given_{t}"], equal_nan=True)
        assert np.array_equal(wo, G[f"hps_width_{tag}_given_{t}"], equal_nan=True)


def test_reference_histogram_known_answers():
    """tests/processors/test_histogram.py:9-19 and :22-60 of the reference"""
    hw, hb = S.histogram(G["kat_hist_in"], 66)
    assert np.array_equal(hw[0], G["kat_hist_w"]) and np.array_equal(hb[0], G["kat_hist_b"])
    assert all(hb[0] == np.arange(67)) and all(hw[0][0::2] == 2) and all(hw[0][1::2] == 1)
    for tag in "abc":
        aw, ab = S.histogram_around_mode(G[f"kat_ham_in_{tag}"], np.nan, 1.0, 11)
        assert np.array_equal(aw[0], G[f"kat_ham_w_{tag}"]) and np.array_equal(ab[0], G[f"kat_ham_b_{tag}"])
    assert G["kat_ham_w_a"].max() == 3 and G["kat_ham_w_a"].sum() == 7 and G["kat_ham_w_b"].sum() < 8


@pytest.mark.parametrize("t", ["f", "d"])
def test_peak_selection_and_amplitudes(t):
    curr, vmax = G[f"curr_{t}"], G[f"vt_max_{t}"]
    for tag, ratio, width in (("a", 0.8, 10), ("b", 0.3, 4)):
        trig, no = S.peak_snr_threshold(curr, vmax, ratio, width)
        assert np.array_equal(trig, G[f"trig_{tag}_{t}"], equal_nan=True)
        assert np.array_equal(no, G[f"n_trig_{tag}_{t}"])
        en = S.multi_a_filter(curr, trig)
        assert np.array_equal(en, G[f"energies_{tag}_{t}"], equal_nan=True)


@pytest.mark.parametrize("t", ["f", "d"])
def test_dplms(t):
    """tests/processors/test_dplms.py:10-35: the reference's 50 x 50 noise matrix, delta and pulse references"""
    rt = 2e-5 if t == "f" else 1e-10
    k = S.dplms(G["dplms_nmat"].astype(DT[t]), G["dplms_ref"].astype(DT[t]), 1, 1, 1, 1, 50, DT[t])
    close(k, G[f"dplms_{t}_1111"], rt)
    k = S.dplms(G["dplms_nmat"].astype(DT[t]), G["dplms_ref_pulse"].astype(DT[t]), 50, 0.1, 1, 1, 50, DT[t])
    close(k, G[f"dplms_{t}_pulse"], rt)
