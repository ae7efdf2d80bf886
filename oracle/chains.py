"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Oracle chain runners: the benchmark chains of BASELINE.json sequenced by hand in
*sample units* (sampling period 16 ns), every step evaluated by the CPU oracle
(``oracle.oracle``).  This restates what the reference's ``ProcessingChain`` would
execute for ``tests/configs/icpc-dsp-config.json`` (order from its dependency
resolution, ``processing_chain.py:2601-2651``; unit -> sample conversion
``processing_chain.py:1747-1770``), independently of the
product's chain compiler, so it also checks the product's unit/grid bookkeeping.
"""

from __future__ import annotations

import numpy as np

from . import oracle as O

f32 = np.float32


def icpc_constants():
    return {
        "t0_kernel": O.t0_filter(8, 125, 133),
        "cusp_kernel": O.cusp_filter(1250, 188, 28125, 5792),
        "zac_kernel": O.zac_filter(1250, 188, 28125, 5792),
    }


def _conv_library(w, kernel, mode, threads):
    """The reference's own library calls for the convolutions -- numpy.convolve per row
    (convolutions.py:72) for short kernels, scipy.signal.fftconvolve per block
    (convolutions.py:118) for the long cusp/zac kernels -- spread over row chunks with a
    thread pool (both release the GIL).  Used for the timed CPU baseline only: it is what
    the reference executes and faster than the oracle's float64 direct sum."""
    from concurrent.futures import ThreadPoolExecutor

    from scipy.signal import fftconvolve

    full = {"f": "full", "v": "valid", "s": "same"}[mode]
    n_rows = w.shape[0]
    threads = max(1, min(threads, n_rows))
    bounds = np.linspace(0, n_rows, threads + 1).astype(int)

    def work(i):
        blk = w[bounds[i] : bounds[i + 1]]
        if len(kernel) > 512:
            return fftconvolve(blk, kernel.reshape(1, -1), mode=full, axes=-1).astype(np.float32)
        return np.stack([np.convolve(r, kernel, mode=full) for r in blk]).astype(np.float32) if len(blk) else \
            np.zeros((0, 0), np.float32)

    with ThreadPoolExecutor(threads) as ex:
        parts = [p for p in ex.map(work, range(threads)) if p.shape[0]]
    return np.concatenate(parts)


def icpc_chain(values: np.ndarray, baseline: np.ndarray, tau: float = 27460.5, consts=None, keep_waveforms: bool = True,
               conv: str = "direct", threads: int = 1) -> dict:
    """ICPC HPGe chain (icpc-dsp-config.json) on uint16 ``values [n, 8192]``.
    Returns sample-domain results (time points are sample indices; multiply by
    16 ns and add t0 for the chain's ``ns`` outputs).  ``conv="library"`` evaluates the
    three convolutions with the reference's numpy/scipy calls instead of the oracle's
    float64 direct sum (CPU-baseline timing)."""
    c = consts or icpc_constants()
    if conv == "library":
        convolve = lambda w, k, m: _conv_library(np.ascontiguousarray(w), k, m, threads)  # noqa: E731
    else:
        convolve = O.convolve_wf
    o = {}
    wf = values.astype(np.float32)
    o["tp_min"], o["tp_max"], o["wf_min"], o["wf_max"] = O.min_max(wf)
    blsub = O.bl_subtract(wf, baseline.astype(np.float32))
    o["bl_mean"], o["bl_std"], o["bl_slope"], o["bl_intercept"] = O.linear_slope_fit(blsub[:, 0:750])
    pz = O.pole_zero(blsub, tau)
    o["pz_mean"], o["pz_std"], o["pz_slope"], o["pz_intercept"] = O.linear_slope_fit(pz[:, 1500:])
    t0f = convolve(pz, c["t0_kernel"], "s")
    atrap = O.asym_trap_filter(pz, 8, 4, 125)
    o["conv_tmin"], o["tp_start"], o["conv_min"], o["conv_max"] = O.min_max(t0f)
    o["tp_0_atrap"] = O.time_point_thresh(atrap, o["bl_std"], o["tp_start"], 0)
    o["tp_0_est"] = O.time_point_thresh(t0f, o["bl_std"], o["tp_start"], 0)
    trap = O.trap_norm(pz, 625, 188)
    o["trapTmax"] = np.amax(trap, 1)
    o["trapEmax"] = o["trapTmax"]
    o["trapEftp_t"] = np.rint((f32(o["tp_0_est"] + f32(625)) + f32(150)).astype(np.float64)).astype(np.float32)
    o["trapEftp"] = O.fixed_time_pickoff(trap, o["trapEftp_t"], "l")
    cusp = convolve(np.ascontiguousarray(blsub[:, :6092]), c["cusp_kernel"], "v")
    zac = convolve(np.ascontiguousarray(blsub[:, :6092]), c["zac_kernel"], "v")
    o["cuspEmax"], o["zacEmax"] = np.amax(cusp, 1), np.amax(zac, 1)
    o["cuspEftp"] = O.fixed_time_pickoff(cusp, 50, "i")
    o["zacEftp"] = O.fixed_time_pickoff(zac, 50, "i")
    o["tp_100"] = O.time_point_thresh(pz, o["trapTmax"], o["tp_0_est"], 1)
    o["tp_99"] = O.time_point_thresh(pz, f32(0.99) * o["trapTmax"], o["tp_0_est"], 1)
    prev = "tp_99"
    for name, frac in (("tp_95", 0.95), ("tp_90", 0.9), ("tp_80", 0.8), ("tp_50", 0.5), ("tp_20", 0.2), ("tp_10", 0.1), ("tp_01", 0.01)):
        o[name] = O.time_point_thresh(pz, o["trapTmax"] * f32(frac), o[prev], 0)
        prev = name
    trap2 = O.trap_norm(pz, 250, 6)
    o["trapQftp"] = O.fixed_time_pickoff(trap2, f32(o["tp_0_est"] + f32(506)), "l")
    o["QDrift"] = f32(o["trapQftp"] * f32(16))
    with np.errstate(all="ignore"):
        o["dt_eff"] = f32(o["QDrift"] / o["trapTmax"])
    le = O.windower(pz, o["tp_0_est"], 301)
    curr = O.avg_current(le, 1)
    curr_up = O.upsampler(curr, 16, 4784)
    curr_av = O.moving_window_multi(curr_up, 48, 3, 0)
    o["aoe_t_min"], o["tp_aoe_max"], o["A_min"], o["A_max"] = O.min_max(curr_av)
    o["tp_aoe_samp"] = f32(o["tp_0_est"] + f32(o["tp_aoe_max"] / f32(16)))
    if keep_waveforms:
        o.update(wf_blsub=blsub, wf_pz=pz, wf_t0_filter=t0f, wf_atrap=atrap, wf_trap=trap, wf_trap2=trap2,
                 wf_cusp=cusp, wf_zac=zac, wf_le=le, curr=curr, curr_up=curr_up, curr_av=curr_av)
    return o


#: the 34 outputs of icpc-dsp-config.json and whether each is a time coordinate
#: (converted to ns on output: (samples + t0/dt) * dt)
ICPC_OUTPUTS = [
    "tp_min", "tp_max", "wf_min", "wf_max", "bl_mean", "bl_std", "bl_slope", "bl_intercept",
    "pz_slope", "pz_std", "pz_mean", "trapTmax", "tp_0_est", "tp_0_atrap", "tp_10", "tp_20",
    "tp_50", "tp_80", "tp_90", "tp_99", "tp_100", "tp_01", "tp_95", "A_max", "QDrift", "dt_eff",
    "tp_aoe_max", "tp_aoe_samp", "trapEmax", "trapEftp", "cuspEmax", "zacEmax", "zacEftp", "cuspEftp",
]
ICPC_TIME_OUTPUTS = {
    "tp_min", "tp_max", "tp_0_est", "tp_0_atrap", "tp_10", "tp_20", "tp_50", "tp_80", "tp_90",
    "tp_99", "tp_100", "tp_01", "tp_95", "tp_aoe_samp",
}


def minimal_chain(values: np.ndarray, tau: float = 27460.5) -> dict:
    """BASELINE.json config 0: mean/stdev baseline (linear_slope_fit on the first
    750 samples) + pole_zero + trap_norm + amax + trap_pickoff."""
    o = {}
    wf = values.astype(np.float32)
    o["bl_mean"], o["bl_std"], o["bl_slope"], o["bl_intercept"] = O.linear_slope_fit(wf[:, 0:750])
    blsub = O.bl_subtract(wf, o["bl_mean"])
    pz = O.pole_zero(blsub, tau)
    trap = O.trap_norm(pz, 625, 188)
    o["trapEmax"] = np.amax(trap, 1)
    o["tp_max"] = np.argmax(trap, 1).astype(np.float32)
    o["trapEpick"] = O.trap_pickoff(pz, 625, 188, o["tp_max"])
    return o


def sipm_chain(values: np.ndarray, baseline: np.ndarray, keep_waveforms: bool = True) -> dict:
    """BASELINE.json config 3 (SiPM): bl_subtract + moving_window_multi + avg_current
    + get_multi_local_extrema (20-slot outputs)."""
    o = {}
    blsub = O.bl_subtract(values.astype(np.float32), baseline.astype(np.float32))
    mw = O.moving_window_multi(blsub, 8, 2, 0)
    curr = O.avg_current(mw, 4)
    o["vt_max"], o["vt_min"], o["n_max"], o["n_min"] = O.get_multi_local_extrema(mw, 12.0, 6.0, 3, 15.0, 1000.0, 20)
    if keep_waveforms:
        o.update(wf_blsub=blsub, wf_mw=mw, curr=curr)
    return o


def sipm_lar_chain(values: np.ndarray, gauss_width: float = 1.0, gauss_trunc: float = 4.0, dt_ns: float = 16.0) -> dict:
    """The reference's SiPM / LAr chain (tests/configs/sipm-dsp-config.json), sequenced like processing_chain.py does:
    the float64 gaussian kernel forces the float64 type loops downstream.  Returns the padded per-row lists, their
    lengths and the ragged (VectorOfVectors) form of the two outputs.  `trigger_pos` stays a SAMPLE index although the
    config labels it "ns": wf_gaus / curr are declared with their own length expressions ("(n),(m),(p)" signature),
    so they inherit no coordinate grid from `waveform` (processing_chain.py:1654-1715) and nothing downstream is a
    coordinate -- the same quirk SURVEY Appendix C notes for tp_aoe_max of the ICPC chain."""
    from oracle import sipm_oracle as S

    o = {}
    k = S.gaussian_filter1d(gauss_width, gauss_trunc, np.float64)
    wf_gaus = S.reflected_convolve_wf(values.astype(np.float64), k)
    curr = O.avg_current(wf_gaus, 5)
    hw, hb = S.histogram(curr, 100)
    _, _, fwhm = S.histogram_stats(hw, hb, np.nan)
    n_rows = len(values)
    vmax = np.empty((n_rows, 20))
    nmax = np.empty(n_rows, np.uint32)
    for r in range(n_rows):      # per-row absolute threshold 3 * fwhm (the C oracle takes one value per call)
        vm, _, nm, _ = O.get_multi_local_extrema(curr[r], 5.0, 0.1, 1, 3.0 * fwhm[r], 0.0, 20)
        vmax[r], nmax[r] = vm, nm
    trig, no = S.peak_snr_threshold(curr, vmax, 0.8, 10)
    en = S.multi_a_filter(curr, trig)
    o.update(curr=curr, fwhm=fwhm, vt_max_candidate=vmax, n_max=nmax, trigger_pos_samples=trig, n_trig=no, energies_padded=en)
    o["cumulative_length"] = np.cumsum(no.astype(np.int64)).astype(np.uint32)
    mask = np.arange(20)[None, :] < no[:, None]
    o["trigger_pos_flat"] = trig[mask]
    o["energies_flat"] = en[mask]
    return o
