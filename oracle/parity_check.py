"""TEST INFRASTRUCTURE (checker): parity report of a B200 ICPC-chain result against the CPU oracle chain.

Used by ``bench.py`` (the benchmarked launch verifies itself against the rows the ``cpu_baseline`` leg computes anyway)
and by the GPU tests.  Never imported by the product path (``dspeed_b200/``).

Rules (north_star; the same as tests/parity.py, but reporting instead of asserting):

* index / extremum outputs (``tp_min, tp_max, wf_min, wf_max``): bit-exact;
* float32 outputs: ``|got - ref| <= 1e-5 * max|ref|`` over the batch (3e-5 for ``dt_eff``, a quotient of two
  such quantities);
* threshold-crossing times: bit-exact, except rows whose crossing is *marginal* -- the oracle's waveform lies
  within ``1e-5 * max|waveform|`` of the threshold next to the oracle's or the GPU's crossing (the threshold itself
  carries float32 drift).  Every differing row must be marginal (else a violation) and marginal rows of the search
  that seeds the others (``tp_0_est``) must stay below 0.1 % of the batch; quantities derived from ``tp_0_est`` are
  compared on the rows where it agrees.
"""
from __future__ import annotations

import numpy as np

EXACT = ["tp_min", "tp_max", "wf_min", "wf_max"]
TP_CHAIN = ["tp_0_est", "tp_0_atrap", "tp_100", "tp_99", "tp_95", "tp_90", "tp_80", "tp_50", "tp_20", "tp_10", "tp_01"]
FLOATS = ["bl_mean", "bl_std", "bl_slope", "bl_intercept", "pz_slope", "pz_std", "pz_mean", "trapTmax", "trapEmax",
          "cuspEmax", "zacEmax", "zacEftp", "cuspEftp"]
DEPEND_ON_T0 = ["A_max", "QDrift", "dt_eff", "tp_aoe_max", "tp_aoe_samp", "trapEftp"]
FLOAT_RTOL = 1e-5
TP_FRACS = (("tp_95", 0.95), ("tp_90", 0.9), ("tp_80", 0.8), ("tp_50", 0.5), ("tp_20", 0.2), ("tp_10", 0.1), ("tp_01", 0.01))


def _marginal(w, thr, idx, tol):
    if not np.isfinite(idx):
        return bool((np.abs(w - thr) <= tol).any())
    i = int(idx)
    lo, hi = max(i - 1, 0), min(i + 2, len(w))
    return bool((np.abs(w[lo:hi] - thr) <= tol).any())


def thresholds(o):
    """time point -> (oracle waveform name, per-row threshold) as the ICPC config defines them"""
    f32 = np.float32
    thr = {"tp_0_est": ("wf_t0_filter", o["bl_std"]), "tp_0_atrap": ("wf_atrap", o["bl_std"]),
           "tp_100": ("wf_pz", o["trapTmax"]), "tp_99": ("wf_pz", f32(0.99) * o["trapTmax"])}
    for name, frac in TP_FRACS:
        thr[name] = ("wf_pz", o["trapTmax"] * f32(frac))
    return thr


def icpc_report(got: dict, o: dict, waves_of=None, dt_ns: float = 16.0) -> dict:
    """`got`: output columns of the B200 chain (time points in ns), `o`: oracle results in samples.  `waves_of(rows)`
    returns the oracle's intermediate waveforms (wf_pz, wf_t0_filter, wf_atrap) of the given rows -- called only for
    rows whose time points differ (when `o` does not carry the waveforms itself)."""
    f32 = np.float32
    n = len(o["trapEmax"])
    viol = []
    rep = {"rows": int(n), "exact_ok": True, "max_rel_err": 0.0, "worst_float": None, "marginal_tp_rows": 0,
           "marginal_tp0_rows": 0, "t0_masked_frac": 0.0}
    samples = {k: (np.asarray(got[k], np.float64) / dt_ns).astype(f32) for k in got if k.startswith("tp_") and k != "tp_aoe_max"}
    for k in EXACT:
        ref = (o[k] * (dt_ns if k.startswith("tp_") else 1.0)).astype(f32)
        if not np.array_equal(np.asarray(got[k]), ref, equal_nan=True):
            rep["exact_ok"] = False
            viol.append(f"{k}: not bit-exact ({int((np.asarray(got[k]) != ref).sum())} rows)")

    def close(name, g, r, mask=None, rtol=FLOAT_RTOL):
        g, r = np.asarray(g, np.float64), np.asarray(r, np.float64)
        mask = np.ones(r.shape, bool) if mask is None else mask
        if not np.array_equal(np.isnan(g[mask]), np.isnan(r[mask])):
            viol.append(f"{name}: NaN pattern differs")
            return
        ok = mask & np.isfinite(r)
        if not ok.any():
            return
        scale = max(np.abs(r[ok]).max(), 1e-30)
        err = np.abs(g[ok] - r[ok]).max() / scale
        if err / (rtol / FLOAT_RTOL) > rep["max_rel_err"]:
            rep["max_rel_err"], rep["worst_float"] = float(err / (rtol / FLOAT_RTOL)), name
        if err > rtol:
            viol.append(f"{name}: max err {err:.3e} of the scale > {rtol:.0e}")

    for k in FLOATS:
        close(k, got[k], o[k])
    thr = thresholds(o)
    agree, differing = {}, set()
    for k in TP_CHAIN:
        g, r = samples[k], np.asarray(o[k])
        agree[k] = (g == r) | (np.isnan(g) & np.isnan(r))
        differing.update(np.flatnonzero(~agree[k]).tolist())
    if differing:
        rows = np.array(sorted(differing))
        waves = waves_of(rows) if waves_of is not None and "wf_pz" not in o else {k: np.asarray(o[k])[rows] for k in ("wf_pz", "wf_t0_filter", "wf_atrap")}
        where = {int(r): j for j, r in enumerate(rows)}
        marg = set()
        for k in TP_CHAIN:
            wname, th = thr[k]
            for r in np.flatnonzero(~agree[k]):
                w = waves[wname][where[int(r)]]
                tol = FLOAT_RTOL * np.abs(w).max()
                if _marginal(w, th[r], o[k][r], tol) or _marginal(w, th[r], samples[k][r], tol):
                    marg.add(int(r))
                elif k != "tp_0_est" and not agree["tp_0_est"][r]:
                    marg.add(int(r))      # starts from a shifted (marginal) tp_0_est: follows that row's verdict
                else:
                    viol.append(f"{k}: row {int(r)} differs ({samples[k][r]} vs {o[k][r]}) and the crossing is not marginal")
        rep["marginal_tp_rows"] = len(marg)
    t0_ok = agree["tp_0_est"]
    rep["marginal_tp0_rows"] = int((~t0_ok).sum())
    rep["t0_masked_frac"] = float((~t0_ok).mean())
    if rep["t0_masked_frac"] > 1e-3:
        viol.append(f"tp_0_est differs on {rep['marginal_tp0_rows']} of {n} rows (> 0.1 %)")
    aoe_ok = t0_ok & ((np.asarray(got["tp_aoe_max"]) == o["tp_aoe_max"]) | (np.isnan(got["tp_aoe_max"]) & np.isnan(o["tp_aoe_max"])))
    for k in DEPEND_ON_T0:
        g = samples[k] if k == "tp_aoe_samp" else np.asarray(got[k])
        if k == "tp_aoe_samp":
            close(k, g, o[k], mask=aoe_ok)
        elif k == "tp_aoe_max":
            if not np.array_equal(g[t0_ok], np.asarray(o[k])[t0_ok], equal_nan=True):
                # an arg-max over a smooth, triple-boxcar-filtered current: ties within float32 rounding move it
                # (its maximum, A_max, is compared above; the position may move between samples that are equal within
                # the float tolerance -- neighbouring samples of a flat top or two equally high peaks)
                bad = np.flatnonzero(t0_ok & ~((g == o[k]) | (np.isnan(g) & np.isnan(o[k]))))
                rep["tp_aoe_max_moved_rows"] = int(len(bad))
                cav = np.asarray(o["curr_av"])[bad] if "curr_av" in o else (waves_of(bad)["curr_av"] if waves_of is not None else None)
                for j, r in enumerate(bad):
                    if cav is None or not (np.isfinite(g[r]) and np.isfinite(o[k][r])):
                        viol.append(f"tp_aoe_max: row {int(r)} differs ({g[r]} vs {o[k][r]})")
                        continue
                    w = cav[j]
                    if not (0 <= g[r] < len(w)):
                        viol.append(f"tp_aoe_max: row {int(r)}: index {g[r]} outside the waveform")
                        continue
                    if abs(float(w[int(g[r])]) - float(w[int(o[k][r])])) > FLOAT_RTOL * np.abs(w).max():
                        viol.append(f"tp_aoe_max: row {int(r)} moved from {o[k][r]} to {g[r]} and the samples differ")
                if len(bad) > max(1, n // 1000):
                    viol.append(f"tp_aoe_max: {len(bad)} rows differ (> 0.1 %)")
        else:
            close(k, g, o[k], mask=t0_ok, rtol=3e-5 if k == "dt_eff" else FLOAT_RTOL)
    rep["violations"] = viol[:8]
    rep["ok"] = not viol
    return rep
