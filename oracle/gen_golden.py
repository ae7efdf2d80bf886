"""TEST INFRASTRUCTURE -- generates the committed golden vectors in tests/golden/.

Runs ONLY in the build container, where the reference lives at /root/reference:
it imports the reference's own numba/numpy/scipy processors (unmodified, through
a stub parent package that skips ``dspeed/__init__.py`` -- that file needs
lgdo/lh5 which are not installed) and records their outputs on seeded inputs.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

Nothing on the GPU box reads /root/reference; the parity tests read the .npz
files this script wrote.
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/dspeed"
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")


def import_reference():
    m = types.ModuleType("dspeed")
    m.__path__ = [REF]
    sys.modules["dspeed"] = m
    import dspeed.processors as P  # noqa: E402

    return P


def load_synth():
    import importlib.util

    spec = importlib.util.spec_from_file_location("synth", os.path.join(REPO, "dspeed_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def f32(x):
    return np.asarray(x, np.float32)


def hpge_inputs(synth):
    """10 rows: 6 ordinary pulses, a pile-up, a saturated pulse, a flat (pulse-free)
    row and a very small pulse."""
    d = synth.hpge_waveforms(6, seed=20261018)
    vals = d["values"].numpy().copy()
    bl = d["baseline"].numpy().copy()
    rng = np.random.default_rng(7)
    t = np.arange(8192, dtype=np.float64)
    tau = synth.HPGE_TAU_SAMPLES

    def pulse(a, t0, tr):
        x = np.clip(t - t0, 0, None)
        return a * (t >= t0) * (1 - np.exp(-x / tr)) * np.exp(-x / tau)

    extra = []
    extra_bl = []
    # pile-up
    b = 12000.0
    extra.append(b + pulse(9000, 3900, 12) + pulse(5000, 5200, 20) + rng.normal(0, 4, 8192))
    extra_bl.append(b)
    # saturated
    b = 14000.0
    extra.append(b + pulse(70000, 4000, 10) + rng.normal(0, 4, 8192))
    extra_bl.append(b)
    # flat
    b = 11000.0
    extra.append(b + rng.normal(0, 4, 8192))
    extra_bl.append(b)
    # tiny pulse
    b = 10500.0
    extra.append(b + pulse(60, 4100, 15) + rng.normal(0, 4, 8192))
    extra_bl.append(b)
    extra = np.clip(np.round(np.array(extra)), 0, 65535).astype(np.uint16)
    vals = np.concatenate([vals, extra])
    bl = np.concatenate([bl, np.round(extra_bl).astype(np.uint16)])
    return vals, bl


def run_icpc_chain(P, values, baseline):
    """The reference ICPC chain (tests/configs/icpc-dsp-config.json) in sample units
    (dt = 16 ns), every step executed by the reference's own processors.  See
    SURVEY.md appendix C for the unit -> sample derivation."""
    n_rows, L = values.shape
    o = {}
    z = lambda *s: np.zeros(s, np.float32)  # noqa: E731
    sc = lambda: np.zeros(n_rows, np.float32)  # noqa: E731

    o["tp_min"], o["tp_max"], o["wf_min"], o["wf_max"] = sc(), sc(), sc(), sc()
    P.min_max(values, o["tp_min"], o["tp_max"], o["wf_min"], o["wf_max"])
    o["wf_blsub"] = z(n_rows, L)
    P.bl_subtract(values, baseline, o["wf_blsub"])
    for k in ("bl_mean", "bl_std", "bl_slope", "bl_intercept", "pz_mean", "pz_std", "pz_slope", "pz_intercept"):
        o[k] = sc()
    P.linear_slope_fit(o["wf_blsub"][:, 0:750], o["bl_mean"], o["bl_std"], o["bl_slope"], o["bl_intercept"])
    o["wf_pz"] = z(n_rows, L)
    P.pole_zero(o["wf_blsub"], np.float32(27460.5), o["wf_pz"])
    P.linear_slope_fit(o["wf_pz"][:, 1500:], o["pz_mean"], o["pz_std"], o["pz_slope"], o["pz_intercept"])
    o["t0_kernel"] = z(133)
    P.t0_filter(np.float32(8), np.float32(125), o["t0_kernel"])
    o["wf_t0_filter"] = z(n_rows, L)
    P.convolve_wf(o["wf_pz"], o["t0_kernel"], np.int8(ord("s")), o["wf_t0_filter"])
    o["wf_atrap"] = z(n_rows, L)
    P.asym_trap_filter(o["wf_pz"], 8, 4, 125, o["wf_atrap"])
    o["conv_tmin"], o["tp_start"], o["conv_min"], o["conv_max"] = sc(), sc(), sc(), sc()
    P.min_max(o["wf_t0_filter"], o["conv_tmin"], o["tp_start"], o["conv_min"], o["conv_max"])
    o["tp_0_atrap"], o["tp_0_est"] = sc(), sc()
    P.time_point_thresh(o["wf_atrap"], o["bl_std"], o["tp_start"], 0, o["tp_0_atrap"])
    P.time_point_thresh(o["wf_t0_filter"], o["bl_std"], o["tp_start"], 0, o["tp_0_est"])
    o["wf_trap"] = z(n_rows, L)
    P.trap_norm(o["wf_pz"], 625, 188, o["wf_trap"])
    o["trapTmax"] = np.amax(o["wf_trap"], 1)
    o["wf_etrap"] = o["wf_trap"]
    o["trapEmax"] = o["trapTmax"]
    o["trapEftp_t"] = np.rint(
        (f32(o["tp_0_est"] + np.float32(625)) + np.float32(150)).astype(np.float64)
    ).astype(np.float32)
    o["trapEftp"] = sc()
    P.fixed_time_pickoff(o["wf_etrap"], o["trapEftp_t"], np.int8(ord("l")), o["trapEftp"])
    o["cusp_kernel"], o["zac_kernel"] = z(5792), z(5792)
    P.cusp_filter(np.float32(1250), np.float32(188), np.float32(28125), o["cusp_kernel"])
    P.zac_filter(np.float32(1250), np.float32(188), np.float32(28125), o["zac_kernel"])
    o["wf_cusp"], o["wf_zac"] = z(n_rows, 301), z(n_rows, 301)
    P.fft_convolve_wf(o["wf_blsub"][:, :6092].copy(), o["cusp_kernel"], np.int8(ord("v")), o["wf_cusp"])
    P.fft_convolve_wf(o["wf_blsub"][:, :6092].copy(), o["zac_kernel"], np.int8(ord("v")), o["wf_zac"])
    # direct (np.convolve) evaluation of the same convolutions, for error budgeting
    o["wf_cusp_direct"], o["wf_zac_direct"] = z(n_rows, 301), z(n_rows, 301)
    P.convolve_wf(o["wf_blsub"][:, :6092].copy(), o["cusp_kernel"], np.int8(ord("v")), o["wf_cusp_direct"])
    P.convolve_wf(o["wf_blsub"][:, :6092].copy(), o["zac_kernel"], np.int8(ord("v")), o["wf_zac_direct"])
    o["cuspEmax"] = np.amax(o["wf_cusp"], 1)
    o["zacEmax"] = np.amax(o["wf_zac"], 1)
    o["cuspEftp"], o["zacEftp"] = sc(), sc()
    P.fixed_time_pickoff(o["wf_cusp"], np.float32(50), np.int8(ord("i")), o["cuspEftp"])
    P.fixed_time_pickoff(o["wf_zac"], np.float32(50), np.int8(ord("i")), o["zacEftp"])
    o["tp_100"], o["tp_99"] = sc(), sc()
    P.time_point_thresh(o["wf_pz"], o["trapTmax"], o["tp_0_est"], 1, o["tp_100"])
    P.time_point_thresh(o["wf_pz"], np.float32(0.99) * o["trapTmax"], o["tp_0_est"], 1, o["tp_99"])
    prev = "tp_99"
    for name, frac in (("tp_95", 0.95), ("tp_90", 0.9), ("tp_80", 0.8), ("tp_50", 0.5), ("tp_20", 0.2), ("tp_10", 0.1), ("tp_01", 0.01)):
        o[name] = sc()
        P.time_point_thresh(o["wf_pz"], o["trapTmax"] * np.float32(frac), o[prev], 0, o[name])
        prev = name
    o["wf_trap2"] = z(n_rows, L)
    P.trap_norm(o["wf_pz"], 250, 6, o["wf_trap2"])
    o["trapQftp"] = sc()
    P.fixed_time_pickoff(o["wf_trap2"], f32(o["tp_0_est"] + np.float32(506)), np.int8(ord("l")), o["trapQftp"])
    o["QDrift"] = f32(o["trapQftp"] * np.float32(16))
    o["dt_eff"] = f32(o["QDrift"] / o["trapTmax"])
    o["wf_le"] = z(n_rows, 301)
    P.windower(o["wf_pz"], o["tp_0_est"], o["wf_le"])
    o["curr"] = z(n_rows, 300)
    P.avg_current(o["wf_le"], 1, o["curr"])
    o["curr_up"] = z(n_rows, 4784)
    P.upsampler(o["curr"], 16, o["curr_up"])
    o["curr_av"] = z(n_rows, 4784)
    P.moving_window_multi(o["curr_up"], 48, 3, 0, o["curr_av"])
    o["aoe_t_min"], o["tp_aoe_max"], o["A_min"], o["A_max"] = sc(), sc(), sc(), sc()
    P.min_max(o["curr_av"], o["aoe_t_min"], o["tp_aoe_max"], o["A_min"], o["A_max"])
    o["tp_aoe_samp"] = f32(o["tp_0_est"] + f32(o["tp_aoe_max"] / np.float32(16)))
    return o


def processor_cases(P):
    """Short-vector cases over assorted parameters for every hot-path processor.
    Returns {key: array}; keys '<case>/in_k' and '<case>/out_k'."""
    rng = np.random.default_rng(99)
    g = {}

    def smooth(n_rows, n, scale=100.0):
        w = np.cumsum(rng.normal(0, 1, (n_rows, n)), axis=1) * scale / np.sqrt(n)
        return w.astype(np.float32)

    w = smooth(5, 300)
    w[4, 17] = np.nan  # NaN row -> NaN outputs
    wd = w.astype(np.float64)
    g["base/w"] = w

    for tag, arr in (("f", w), ("d", wd)):
        dt = arr.dtype
        z = lambda *s: np.zeros(s, dt)  # noqa: E731
        nr, n = arr.shape
        o = z(nr, n)
        P.bl_subtract(arr, dt.type(3.25), o)
        g[f"bl_subtract_{tag}/out"] = o
        outs = [z(nr) for _ in range(4)]
        P.linear_slope_fit(arr, *outs)
        for i, x in enumerate(outs):
            g[f"linear_slope_fit_{tag}/out{i}"] = x
        outs2 = [z(nr) for _ in range(2)]
        P.linear_slope_diff(arr, outs[2], outs[3], *outs2)
        for i, x in enumerate(outs2):
            g[f"linear_slope_diff_{tag}/out{i}"] = x
        o = z(nr)
        P.mean_below_threshold(arr, dt.type(0.5), o)
        g[f"mean_below_threshold_{tag}/out"] = o
        o = z(nr, n)
        P.pole_zero(arr, dt.type(45.5), o)
        g[f"pole_zero_{tag}/out"] = o
        o = z(nr, n)
        P.double_pole_zero(arr, dt.type(45.5), dt.type(7.25), dt.type(0.03), o)
        g[f"double_pole_zero_{tag}/out"] = o
        for (r, f) in ((10, 4), (1, 0), (50, 100), (0, 5), (3, 0)):
            o = z(nr, n)
            P.trap_filter(arr, r, f, o)
            g[f"trap_filter_{tag}_{r}_{f}/out"] = o
            if r > 0:
                o = z(nr, n)
                P.trap_norm(arr, r, f, o)
                g[f"trap_norm_{tag}_{r}_{f}/out"] = o
        for (r, f, fl) in ((8, 4, 125), (10, 0, 10), (1, 1, 1), (100, 50, 150)):
            o = z(nr, n)
            P.asym_trap_filter(arr, r, f, fl, o)
            g[f"asym_trap_filter_{tag}_{r}_{f}_{fl}/out"] = o
        for (r, f, t) in ((10, 4, 100.0), (10, 4, 22.0), (10, 4, 23.0), (10, 4, 299.0), (10, 4, 300.0), (50, 100, 250.0)):
            o = z(nr)
            P.trap_pickoff(arr, r, f, dt.type(t), o)
            g[f"trap_pickoff_{tag}_{r}_{f}_{int(t)}/out"] = o
        for L in (1, 2, 7, 48, 299, 10.5):
            o = z(nr, n)
            P.moving_window_left(arr, dt.type(L), o)
            g[f"moving_window_left_{tag}_{L}/out"] = o
            o = z(nr, n)
            P.moving_window_right(arr, dt.type(L), o)
            g[f"moving_window_right_{tag}_{L}/out"] = o
        for (L, num, typ) in ((48, 3, 0), (5, 1, 0), (5, 2, 1), (5, 4, 2), (16, 0, 0), (1, 3, 0)):
            o = z(nr, n)
            P.moving_window_multi(arr, dt.type(L), dt.type(num), typ, o)
            g[f"moving_window_multi_{tag}_{L}_{num}_{typ}/out"] = o
        for L in (1, 3, 100):
            o = z(nr, n - L)
            P.avg_current(arr, dt.type(L), o)
            g[f"avg_current_{tag}_{L}/out"] = o
        # threshold searches: per-row thresholds and starts
        thr = np.array([5.0, -3.0, 0.0, 20.0, 1.0], dt)
        ts = np.array([150, 0, 299, 10, 3], dt)
        g[f"tpt_{tag}/thr"], g[f"tpt_{tag}/ts"] = thr, ts
        for wf in (0, 1):
            o = z(nr)
            P.time_point_thresh(arr, thr, ts, wf, o)
            g[f"time_point_thresh_{tag}_{wf}/out"] = o
            for mode in "iafbcrnl":
                o = z(nr)
                P.interpolated_time_point_thresh(arr, thr, ts, wf, np.int8(ord(mode)), o)
                g[f"interpolated_time_point_thresh_{tag}_{wf}_{mode}/out"] = o
        mthr = np.array([[-10.0, 0.0, 4.0, 12.0], [3.0, -3.0, 1.0, 0.5], [0.0, 1.0, 2.0, 3.0], [50.0, -50.0, 5.0, 7.0], [1, 2, 3, 4]], dt)
        g[f"mtpt_{tag}/thr"] = mthr
        for pol in (1, -1):
            for mode in "iafbcrnl":
                o = z(nr, 4)
                P.multi_time_point_thresh(arr, mthr, ts, pol, np.int8(ord(mode)), o)
                g[f"multi_time_point_thresh_{tag}_{pol}_{mode}/out"] = o
        tpk = np.array([3.0, 3.25, 0.2, 297.5, 298.9], dt)
        g[f"ftp_{tag}/t"] = tpk
        for mode in "nfclhs":
            o = z(nr)
            P.fixed_time_pickoff(arr, tpk, np.int8(ord(mode)), o)
            g[f"fixed_time_pickoff_{tag}_{mode}/out"] = o
        o = z(nr)
        P.fixed_time_pickoff(arr, np.array([3, 0, 299, 300, -1], dt), np.int8(ord("i")), o)
        g[f"fixed_time_pickoff_{tag}_i/out"] = o
        outs = [z(nr) for _ in range(4)]
        P.min_max(arr, *outs)
        for i, x in enumerate(outs):
            g[f"min_max_{tag}/out{i}"] = x
        o = z(nr, n)
        P.min_max_norm(arr, outs[2], outs[3], o)
        g[f"min_max_norm_{tag}/out"] = o
        t0s = np.array([10.0, -20.0, 250.0, 310.0, 5.0], dt)
        g[f"windower_{tag}/t0"] = t0s
        o = z(nr, 101)
        P.windower(arr, t0s, o)
        g[f"windower_{tag}/out"] = o
        for (up, m) in ((16, 4784), (4, 1000), (3, 950), (2.5, 700)):
            o = z(nr, m)
            P.upsampler(arr, dt.type(up), o)
            g[f"upsampler_{tag}_{up}_{m}/out"] = o
        kern = rng.normal(0, 1, 33).astype(dt)
        g[f"conv_{tag}/kernel"] = kern
        for mode, m in (("f", n + 32), ("v", n - 32), ("s", n)):
            o = z(nr, m)
            P.convolve_wf(arr, kern, np.int8(ord(mode)), o)
            g[f"convolve_wf_{tag}_{mode}/out"] = o
            o = z(nr, m)
            P.fft_convolve_wf(arr.copy(), kern, np.int8(ord(mode)), o)
            g[f"fft_convolve_wf_{tag}_{mode}/out"] = o
        kern2 = rng.normal(0, 1, 34).astype(dt)
        g[f"conv_{tag}/kernel_even"] = kern2
        o = z(nr, n)
        P.convolve_wf(arr, kern2, np.int8(ord("s")), o)
        g[f"convolve_wf_{tag}_s_even/out"] = o
        a = np.array([1.0, -0.5], np.float64)
        b = np.array([1.0, -0.9, 0.1], np.float64)
        o = z(nr, n)
        P.recursive_filter(arr, a, b, dt.type(0.5), dt.type(-0.25), o)
        g[f"recursive_filter_{tag}/out"] = o
        for sd in (0, 1, 2, 3):
            for (dmax, dmin, amax, amin) in ((8.0, 8.0, -1e9, 1e9), (3.0, 1.0, 0.0, 20.0), (20.0, 5.0, 10.0, 0.0)):
                vmax, vmin = z(nr, 6), z(nr, 6)
                nmax, nmin = np.zeros(nr, np.uint32), np.zeros(nr, np.uint32)
                P.get_multi_local_extrema(arr, dmax, dmin, sd, amax, amin, vmax, vmin, nmax, nmin)
                k = f"get_multi_local_extrema_{tag}_{sd}_{dmax}_{dmin}_{amax}_{amin}"
                g[k + "/vmax"], g[k + "/vmin"], g[k + "/nmax"], g[k + "/nmin"] = vmax, vmin, nmax, nmin
    return g


def kernel_cases(P):
    g = {}
    for dt in (np.float32, np.float64):
        tag = "f" if dt == np.float32 else "d"
        for (sigma, flat, decay, n) in ((1250.0, 188.0, 28125.0, 5792), (100.5, 10.0, 500.0, 301), (20.0, 0.0, 100.0, 64)):
            k = np.zeros(n, dt)
            P.cusp_filter(dt(sigma), dt(flat), dt(decay), k)
            g[f"cusp_{tag}_{n}"] = k
            k = np.zeros(n, dt)
            P.zac_filter(dt(sigma), dt(flat), dt(decay), k)
            g[f"zac_{tag}_{n}"] = k
        for (rise, fall) in ((8, 125), (1, 3), (20, 20)):
            k = np.zeros(rise + fall, dt)
            P.t0_filter(dt(rise), dt(fall), k)
            g[f"t0_{tag}_{rise}_{fall}"] = k
        for n in (5, 32):
            k = np.zeros(n, dt)
            P.moving_slope(k)
            g[f"moving_slope_{tag}_{n}"] = k
            k = np.zeros(n, dt)
            P.step(dt(1), k)
            g[f"step_{tag}_{n}"] = k
    return g


def sipm_chain(P, synth):
    d = synth.sipm_waveforms(12, seed=555)
    vals = d["values"].numpy()
    bl = d["baseline"].numpy()
    n_rows, L = vals.shape
    o = {"values": vals, "baseline": bl}
    z = lambda *s: np.zeros(s, np.float32)  # noqa: E731
    o["wf_blsub"] = z(n_rows, L)
    P.bl_subtract(vals, bl, o["wf_blsub"])
    o["wf_mw"] = z(n_rows, L)
    P.moving_window_multi(o["wf_blsub"], 8, 2, 0, o["wf_mw"])
    o["curr"] = z(n_rows, L - 4)
    P.avg_current(o["wf_mw"], 4, o["curr"])
    for sd in (0, 1, 2, 3):
        vmax, vmin = z(n_rows, 20), z(n_rows, 20)
        nmax, nmin = np.zeros(n_rows, np.uint32), np.zeros(n_rows, np.uint32)
        P.get_multi_local_extrema(o["wf_mw"], 12.0, 6.0, sd, 15.0, 1000.0, vmax, vmin, nmax, nmin)
        o[f"vt_max_{sd}"], o[f"vt_min_{sd}"], o[f"n_max_{sd}"], o[f"n_min_{sd}"] = vmax, vmin, nmax, nmin
    return o


def sipm_processor_cases(P, synth):
    """the processors of the reference's SiPM chain (tests/configs/sipm-dsp-config.json) and the remaining kernel
    generators, float32 and float64 type loops, on 12 synthetic SiPM rows + the reference tests' own vectors"""
    import dspeed.processors.gaussian_filter1d as G
    from dspeed.processors.convolutions import reflected_convolve_wf
    from dspeed.processors.histogram import histogram, histogram_around_mode
    from dspeed.processors.histogram_stats import histogram_peakstats, histogram_stats
    from dspeed.processors.multi_a_filter import multi_a_filter
    from dspeed.processors.peak_snr_threshold import peak_snr_threshold

    g = {}
    d = synth.sipm_waveforms(12, seed=777)
    vals = d["values"].numpy()
    n_rows, L = vals.shape
    g["values"] = vals
    for dt in (np.float32, np.float64):
        t = "f" if dt == np.float32 else "d"
        # --- gaussian kernels (sipm-dsp-config.json:4-16: width 1, trunc 4 -> 9 taps)
        for (sig, trunc) in ((1.0, 4.0), (2.5, 3.0)):
            k = np.zeros(int(trunc * sig + 0.5) * 2 + 1, dt)
            G.gaussian_filter1d(dt(sig), dt(trunc), k)
            g[f"gaus_{t}_{sig}_{trunc}"] = k
        k = g[f"gaus_{t}_1.0_4.0"]
        w = vals.astype(dt)
        wg = np.zeros((n_rows, L), dt)
        reflected_convolve_wf(w, k, wg)
        g[f"wf_gaus_{t}"] = wg
        k2 = g[f"gaus_{t}_2.5_3.0"]
        wg2 = np.zeros((n_rows, L), dt)
        reflected_convolve_wf(w, k2, wg2)
        g[f"wf_gaus2_{t}"] = wg2
        curr = np.zeros((n_rows, L - 5), dt)
        P.avg_current(wg, 5, curr)
        g[f"curr_{t}"] = curr
        hw, hb = np.zeros((n_rows, 100), dt), np.zeros((n_rows, 101), dt)
        histogram(curr, hw, hb)
        g[f"hist_w_{t}"], g[f"hist_b_{t}"] = hw, hb
        idx, mx, fw = np.zeros(n_rows, dt), np.zeros(n_rows, dt), np.zeros(n_rows, dt)
        histogram_stats(hw, hb, idx, mx, fw, dt(np.nan))
        g[f"hs_idx_{t}"], g[f"hs_max_{t}"], g[f"hs_fwhm_{t}"] = idx, mx, fw
        idx2, mx2, fw2 = np.zeros(n_rows, dt), np.zeros(n_rows, dt), np.zeros(n_rows, dt)
        histogram_stats(hw, hb, idx2, mx2, fw2, dt(0.5))
        g[f"hs2_idx_{t}"], g[f"hs2_max_{t}"], g[f"hs2_fwhm_{t}"] = idx2, mx2, fw2
        # histogram around the mode (mode search / given centre), then the peak statistics in every mode
        for tag, center, bw, nb in (("a", np.nan, 1.0, 101), ("b", np.nan, 0.5, 64), ("c", 3.0, 2.0, 31)):
            aw, ab = np.zeros((n_rows, nb), dt), np.zeros((n_rows, nb + 1), dt)
            histogram_around_mode(curr, dt(center), dt(bw), aw, ab)
            g[f"ham_w_{tag}_{t}"], g[f"ham_b_{tag}_{t}"] = aw, ab
            for skip in (0, 1):
                for wt in range(5):
                    mo, wo = np.zeros(n_rows, dt), np.zeros(n_rows, dt)
                    histogram_peakstats(aw, ab, dt(np.nan), np.int32(skip), np.int32(wt), mo, wo)
                    g[f"hps_mode_{tag}_{skip}_{wt}_{t}"], g[f"hps_width_{tag}_{skip}_{wt}_{t}"] = mo, wo
            mo, wo = np.zeros(n_rows, dt), np.zeros(n_rows, dt)
            histogram_peakstats(aw, ab, dt(1.25), np.int32(0), np.int32(0), mo, wo)
            g[f"hps_mode_{tag}_given_{t}"], g[f"hps_width_{tag}_given_{t}"] = mo, wo
        # raw-waveform histogram of integer ADC values (aliasing case of the reference's doc-string)
        rw, rb = np.zeros((n_rows, 40), dt), np.zeros((n_rows, 41), dt)
        histogram(w, rw, rb)
        g[f"hist_raw_w_{t}"], g[f"hist_raw_b_{t}"] = rw, rb
        # peak finding with a per-row absolute threshold (3 * fwhm), SNR cut, amplitudes
        vmax, vmin = np.zeros((n_rows, 20), dt), np.zeros((n_rows, 20), dt)
        nmax, nmin = np.zeros(n_rows, np.uint32), np.zeros(n_rows, np.uint32)
        for r in range(n_rows):
            P.get_multi_local_extrema(curr[r], dt(5), dt(0.1), dt(1), dt(3) * fw[r], dt(0), vmax[r], vmin[r], nmax[r:r + 1], nmin[r:r + 1])
        g[f"vt_max_{t}"], g[f"vt_min_{t}"], g[f"n_max_{t}"], g[f"n_min_{t}"] = vmax, vmin, nmax, nmin
        for tag, ratio, width in (("a", 0.8, 10), ("b", 0.3, 4)):
            trig, no = np.zeros((n_rows, 20), dt), np.zeros(n_rows, np.uint32)
            peak_snr_threshold(curr, vmax, dt(ratio), dt(width), trig, no)
            g[f"trig_{tag}_{t}"], g[f"n_trig_{tag}_{t}"] = trig, no
            en = np.zeros((n_rows, 20), dt)
            multi_a_filter(curr, trig, en)
            g[f"energies_{tag}_{t}"] = en
    # reference KATs (tests/processors/test_histogram.py:9-19)
    v = np.arange(100) * 2 / 3
    hw, hb = np.zeros(66), np.zeros(67)
    histogram(v, hw, hb)
    g["kat_hist_in"], g["kat_hist_w"], g["kat_hist_b"] = v, hw, hb
    for tag, wi in (("a", [1, 2, 2, 2, 3, 4, 5]), ("b", [1, 2, 2, 2, 3, 4, 5, 100]), ("c", [1, 2, 2, 2, 3, 4, 5, -100])):
        wi = np.array(wi, np.float32)
        aw, ab = np.zeros(11, np.float32), np.zeros(12, np.float32)
        histogram_around_mode(wi, np.float32(np.nan), np.float32(1.0), aw, ab)
        g[f"kat_ham_in_{tag}"], g[f"kat_ham_w_{tag}"], g[f"kat_ham_b_{tag}"] = wi, aw, ab
    # dplms: the reference test's 50 x 50 noise matrix (tests/processors/dplms_noise_mat.dat) and delta reference
    nmat = np.array([[float(x) for x in ln.split(" ")] for ln in open("/root/reference/tests/processors/dplms_noise_mat.dat")])
    ref = np.zeros(100)
    ref[49:50] = 1
    g["dplms_nmat"], g["dplms_ref"] = nmat, ref
    for dt in (np.float32, np.float64):
        t = "f" if dt == np.float32 else "d"
        k = np.zeros(50, dt)
        P.dplms(nmat.astype(dt), ref.astype(dt), dt(1), dt(1), dt(1), dt(1), k)
        g[f"dplms_{t}_1111"] = k
        k = np.zeros(50, dt)
        rr = (1 - np.exp(-np.arange(100) / 5.0)) * np.exp(-np.arange(100) / 400.0)
        P.dplms(nmat.astype(dt), rr.astype(dt), dt(50), dt(0.1), dt(1), dt(1), k)
        g[f"dplms_{t}_pulse"] = k
        g["dplms_ref_pulse"] = rr
    return g


def iir_cases(P, synth):
    """the IIR family (pole_zero.py:201-342, rc_cr2.py, iir_filter.py) on 3 HPGe rows (1024 samples around the pulse); the iir_filter factories need
    pint, so their scipy.signal designs are repeated here and applied with the reference's own recursive_filter"""
    import scipy.signal as sg
    from dspeed.processors.pole_zero import convolve_damped_oscillator, convolve_exp, inject_damped_oscillation
    from dspeed.processors.rc_cr2 import rc_cr2

    d = synth.hpge_waveforms(3, seed=4242)
    vals = d["values"].numpy()[:, 3600:4624].copy()
    bl = d["baseline"].numpy()
    g = {"values": vals, "baseline": bl}
    for dt in (np.float32, np.float64):
        t = "f" if dt == np.float32 else "d"
        w = (vals.astype(dt) - bl.astype(dt)[:, None])
        for tau in (50.0, 400.5):
            o = np.zeros_like(w)
            convolve_exp(w, dt(tau), o)
            g[f"cexp_{t}_{tau}"] = o
            o = np.zeros_like(w)
            rc_cr2(w, dt(tau), o)
            g[f"rccr2_{t}_{tau}"] = o
        o = np.zeros_like(w)
        convolve_damped_oscillator(w, 120.0, 0.3, 0.7, o)
        g[f"cdo_{t}"] = o
        o = np.zeros_like(w)
        inject_damped_oscillation(w, 120.0, 0.3, 0.7, 0.05, o)
        g[f"ido_{t}"] = o
        for tag, (a, b, gain) in {
            "butter4_lp": (*sg.iirfilter(4, 0.1, btype="lowpass", ftype="butter"), None),
            "cheby1_3_hp": (*sg.iirfilter(3, 0.2, rp=1.0, btype="highpass", ftype="cheby1"), None),
            "butter2_bp": (*sg.iirfilter(2, [0.05, 0.2], btype="bandpass", ftype="butter"), None),
            "notch": (*sg.iirnotch(0.24, 10.0), 1.0),
            "peak": (*sg.iirpeak(0.24, 10.0), 0.0),
        }.items():
            if gain is None:
                gain = sum(a) / sum(b)
            o = np.zeros_like(w)
            P.recursive_filter(w, a, b, w[..., 0], (gain * w[..., 0]).astype(dt), o)
            g[f"iir_{tag}_{t}"] = o
    return g


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "iir":
        P = import_reference()
        out_dir = os.path.join(REPO, "tests", "golden")
        np.savez_compressed(os.path.join(out_dir, "iir_family.npz"), **iir_cases(P, load_synth()))
        print("iir_family.npz", os.path.getsize(os.path.join(out_dir, "iir_family.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sipm":
        P = import_reference()
        out_dir = os.path.join(REPO, "tests", "golden")
        np.savez_compressed(os.path.join(out_dir, "sipm_processors.npz"), **sipm_processor_cases(P, load_synth()))
        print("sipm_processors.npz", os.path.getsize(os.path.join(out_dir, "sipm_processors.npz")))
        return
    P = import_reference()
    synth = load_synth()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    vals, bl = hpge_inputs(synth)
    chain = run_icpc_chain(P, vals, bl)
    chain["values"], chain["baseline"] = vals, bl
    # keep the file small: full-length intermediates only for 4 rows, scalars for all
    keep_rows = [0, 1, 6, 7]
    slim = {}
    for k, v in chain.items():
        if v.ndim == 2 and v.shape[0] == vals.shape[0] and v.shape[1] > 400 and k != "values":
            slim[k + "__rows"] = v[keep_rows]
        else:
            slim[k] = v
    slim["keep_rows"] = np.array(keep_rows)
    np.savez_compressed(os.path.join(out_dir, "hpge_chain.npz"), **slim)

    np.savez_compressed(os.path.join(out_dir, "processors.npz"), **processor_cases(P))
    np.savez_compressed(os.path.join(out_dir, "kernels.npz"), **kernel_cases(P))
    np.savez_compressed(os.path.join(out_dir, "sipm_chain.npz"), **sipm_chain(P, synth))
    np.savez_compressed(os.path.join(out_dir, "sipm_processors.npz"), **sipm_processor_cases(P, synth))
    np.savez_compressed(os.path.join(out_dir, "iir_family.npz"), **iir_cases(P, synth))
    for f in sorted(os.listdir(out_dir)):
        print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
