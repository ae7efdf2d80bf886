"""Throughput of the other BASELINE.json configurations on one B200 (device-resident inputs,
CUDA-event timing, >= 3 warm-up passes): C1 minimal energy chain, C4 SiPM chain, C5 long-kernel
convolution sweep (cusp / zac / dplms kernels of length 256 ... 4096, direct SMEM kernel vs 3xTF32 tensor-core
Toeplitz GEMM), each with a parity check of a sample of the measured outputs against the CPU oracle.

    python scripts/bench_configs.py            one JSON line per measurement (kept under profiles/)
    bench.py                                   calls run_all() and puts the result under "configs" of its JSON line
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
from dspeed_b200 import processors as P, synth, tables  # noqa: E402
from dspeed_b200.processing_chain import build_processing_chain  # noqa: E402


def _peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1613.1))
    except Exception:
        return 6650.0, 1613.1


def timed(fn, reps=3, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


C1_CFG = {
    "outputs": ["bl_mean", "bl_std", "trapEmax", "tp_max", "trapEpick"],
    "processors": {
        "bl_mean, bl_std, bl_slope, bl_intercept": {
            "function": "linear_slope_fit", "module": "dspeed.processors",
            "args": ["waveform[0:750]", "bl_mean", "bl_std", "bl_slope", "bl_intercept"], "unit": ["ADC"] * 4},
        "wf_blsub": "dspeed.processors.bl_subtract(waveform, bl_mean, wf_blsub(unit='ADC'))",
        "wf_pz": {"function": "dspeed.processors.pole_zero(wf_blsub, db.pz.tau, wf_pz)", "unit": "ADC",
                  "defaults": {"db.pz.tau": "27460.5"}},
        "wf_trap": {"function": "dspeed.processors.trap_norm(wf_pz, 10*us, 3.008*us, wf_trap)", "unit": "ADC"},
        "tmn, tp_max, emn, trapEmax": {"function": "dspeed.processors.min_max(wf_trap, tmn, tp_max, emn, trapEmax)",
                                       "unit": ["ns", "ns", "ADC", "ADC"]},
        "trapEpick": {"function": "dspeed.processors.trap_pickoff(wf_pz, 10*us, 3.008*us, tp_max, trapEpick)", "unit": "ADC"},
    },
}


def _wf_table(values, n, dev):
    return tables.WaveformTable(size=n, t0=tables.Array(torch.zeros(n, dtype=torch.float64, device=dev), attrs={"units": "ns"}),
                                dt=tables.Array(torch.full((n,), 16.0, dtype=torch.float64, device=dev), attrs={"units": "ns"}),
                                values=values)


def run_c1(dev, n=262144, check_rows=4096):
    """C1: minimal energy chain (linear_slope_fit + bl_subtract + pole_zero + trap_norm + min_max + trap_pickoff)"""
    hbm, _ = _peak()
    d = synth.hpge_waveforms(n, seed=5, device=dev)
    tb = tables.Table({"waveform": _wf_table(d["values"], n, dev)}, size=n)
    chain, _, tb_out = build_processing_chain(C1_CFG, tb, device=dev)
    out = tables.Table({k: tables.Array(torch.empty(n, dtype=torch.float32, device=dev)) for k in tb_out}, size=n)
    t = timed(lambda: chain(tb, out))
    bytes_per_wf = 8192 * 2 + 4 * len(C1_CFG["outputs"])
    res = {"config": "C1 minimal energy chain (lsf + bl_subtract + pole_zero + trap_norm + min_max + trap_pickoff), L=8192",
           "rows": n, "waveforms_per_s": n / t, "ms": t * 1e3,
           "tier": type(chain._fused).__name__ if chain._fused is not None else "per-processor kernels",
           "algorithmic_bytes_per_wf": bytes_per_wf, "frac_of_hbm_roofline": n / t * bytes_per_wf / 1e9 / hbm}
    # parity of the measured outputs: the oracle's processors in the chain's order on the first rows
    try:
        from oracle import oracle as O

        m = min(check_rows, n)
        v = d["values"][:m].cpu().numpy().astype(np.float32)
        mean, std, _, _ = O.linear_slope_fit(v[:, :750])
        pz = O.pole_zero(O.bl_subtract(v, mean), np.float32(27460.5))
        trap = O.trap_norm(pz, 625, 188)
        _, tmax, _, emax = O.min_max(trap)
        pick = O.trap_pickoff(pz, 625, 188, tmax)
        ref = {"bl_mean": mean, "bl_std": std, "trapEmax": emax, "tp_max": tmax * 16.0, "trapEpick": pick}
        scale = float(np.abs(pz).max())
        worst, exact = 0.0, True
        for k, r in ref.items():
            g = out[k].nda[:m].cpu().numpy().astype(np.float64)
            if k == "tp_max":
                # arg-max of a trapezoid flat top: equal within float32 rounding of the top counts as a tie
                bad = np.flatnonzero(g != r)
                tie = [abs(float(trap[i, int(g[i] / 16)]) - float(emax[i])) <= 1e-5 * scale for i in bad]
                exact &= all(tie)
                res["tp_max_tie_rows"] = int(len(bad))
            else:
                sc = scale if k != "bl_mean" else float(np.abs(v).max())
                worst = max(worst, float(np.nanmax(np.abs(g - r.astype(np.float64))) / sc))
        res["parity"] = {"rows": m, "max_rel_err": worst, "index_ok": bool(exact), "ok": bool(exact and worst <= 1e-5)}
    except Exception as e:      # the checker is optional here (tests carry the parity proper)
        res["parity"] = {"error": f"{type(e).__name__}: {e}"}
    return res


def run_c4(dev, n=None, check_rows=8192):
    """C4: SiPM chain (bl_subtract + moving_window_multi + get_multi_local_extrema), 4 M short waveforms"""
    import yaml

    hbm, _ = _peak()
    n = n or int(os.environ.get("DSPB_C4_ROWS", 1 << 22))
    d = synth.sipm_waveforms(n, seed=9, device=dev)
    cfg4 = yaml.safe_load(open(os.path.join(REPO, "dspeed_b200", "configs", "sipm_peaks.yaml")))
    tb = tables.Table({"waveform": _wf_table(d["values"], n, dev), "baseline": tables.Array(d["baseline"])}, size=n)
    chain, _, tb_out = build_processing_chain(cfg4, tb, device=dev, block_width=int(os.environ.get("DSPB_C4_BLOCK", 0)) or None)
    # device-resident output columns (written in place by the kernel), like the ICPC bench
    out4 = tables.Table({k: type(v)(torch.empty(tuple(v.nda.shape), dtype=getattr(torch, str(v.nda.dtype)), device=dev),
                                    attrs=dict(v.attrs)) for k, v in tb_out.items()}, size=n)
    t = timed(lambda: chain(tb, out4), reps=3, warm=3)
    L = d["values"].shape[1]
    b4 = L * 2 + 2 + 4 * 40 + 8
    res = {"config": f"C4 SiPM chain (bl_subtract + moving_window_multi + get_multi_local_extrema), L={L}, device-resident in/out",
           "rows": n, "block_width": chain._block_width, "waveforms_per_s": n / t, "ms": t * 1e3,
           "tier": type(chain._fused).__name__ if chain._fused is not None else "per-processor kernels",
           "algorithmic_bytes_per_wf": b4, "achieved_GBps": n / t * b4 / 1e9, "frac_of_hbm_roofline": n / t * b4 / 1e9 / hbm}
    try:
        from oracle import oracle as O

        m = min(check_rows, n)
        sv, sb = d["values"][:m].cpu().numpy(), d["baseline"][:m].cpu().numpy()
        mw = O.moving_window_multi(O.bl_subtract(sv.astype(np.float32), sb.astype(np.float32)), 8, 2, 0)
        vmax, vmin, nmax, nmin = O.get_multi_local_extrema(mw, 12.0, 6.0, 3, 15.0, 1000.0, 20)
        ok = (np.array_equal(out4["vt_max"].nda[:m].cpu().numpy() / 16.0, vmax, equal_nan=True)
              and np.array_equal(out4["vt_min"].nda[:m].cpu().numpy() / 16.0, vmin, equal_nan=True)
              and np.array_equal(out4["n_max"].nda[:m].cpu().numpy(), nmax) and np.array_equal(out4["n_min"].nda[:m].cpu().numpy(), nmin))
        res["parity"] = {"rows": m, "bit_exact_index_lists_and_counts": bool(ok), "ok": bool(ok)}
    except Exception as e:
        res["parity"] = {"error": f"{type(e).__name__}: {e}"}
    return res


def conv_kernels(K, dev):
    """the three long-kernel families of BASELINE.json config 5 at length K (float32, on the device)"""
    ks = {}
    k = torch.empty(K, dtype=torch.float32, device=dev)
    # cusp / zac with the ICPC chain's shaping parameters scaled to the kernel length (sigma = 20 us, flat = 3 us at 5792 taps)
    sc = K / 5792.0
    P.cusp_filter(np.float32(1250.0 * sc), np.float32(max(2.0, round(188 * sc))), np.float32(28125.0), k)
    ks["cusp"] = k.clone()
    P.zac_filter(np.float32(1250.0 * sc), np.float32(max(2.0, round(188 * sc))), np.float32(28125.0), k)
    ks["zac"] = k.clone()
    if hasattr(P, "dplms") or "dplms" in P.__all__:
        # DPLMS optimum filter from a synthetic noise matrix (white + 1/f-like correlated noise) and a pulse reference
        rng = np.random.default_rng(K)
        lag = np.abs(np.arange(K)[:, None] - np.arange(K)[None, :])
        nmat = (np.eye(K) * 16.0 + 4.0 * np.exp(-lag / 64.0)).astype(np.float32)
        t = np.arange(2 * K) - K // 2            # reference pulse of twice the kernel length (energy_kernels.py:232-246)
        ref = np.where(t >= 0, (1.0 - np.exp(-np.clip(t, 0, None) / max(4.0, 0.02 * K))) * np.exp(-np.clip(t, 0, None) / 27460.5), 0.0)
        ref = ref.astype(np.float32)
        kd = torch.empty(K, dtype=torch.float32, device=dev)
        P.dplms(torch.from_numpy(nmat).to(dev), torch.from_numpy(ref).to(dev), np.float32(1.0), np.float32(1.0), np.float32(1.0),
                np.float32(1.0), kd)
        ks["dplms"] = kd
        del rng
    return ks


def run_c5(dev, n=65536, Ks=(256, 512, 1024, 2048, 4096), check_rows=32):
    """C5: 'valid' convolution of 8192-sample waveforms with cusp / zac / dplms kernels of length K:
    direct SMEM-tiled kernel vs 3xTF32 tensor-core Toeplitz GEMM, parity of both against the float64 sums"""
    _, tf = _peak()
    d = synth.hpge_waveforms(n, seed=11, device=dev)
    x = (d["values"].to(torch.float32) - d["baseline"].to(torch.float32)[:, None]).contiguous()
    xs = x[:check_rows].cpu().numpy().astype(np.float64)
    rows = []
    for K in Ks:
        for fam, k in conv_kernels(K, dev).items():
            out = torch.empty((n, 8192 - K + 1), dtype=torch.float32, device=dev)
            flops = 2.0 * K * (8192 - K + 1)
            kh = k.cpu().numpy().astype(np.float64)
            ref = np.stack([np.convolve(r, kh, "valid") for r in xs])       # float64 truth of the reference's sums
            scale = max(np.abs(ref).max(), 1e-30)
            for tc in (False, True):
                P.TC_CONV_MIN_TAPS = 1 if tc else 0
                t = timed(lambda: P.convolve_wf(x, k, np.int8(ord("v")), out), reps=2, warm=3)
                err = float(np.abs(out[:check_rows].cpu().numpy().astype(np.float64) - ref).max() / scale)
                issued = 3 * 2.0 * 128 * (-(-(K + 127) // 32) * 32) * (-(-(8192 - K + 1) // 128)) if tc else None
                rows.append({"kernel": fam, "K": K, "path": "tensor cores (3xTF32 Toeplitz GEMM, tcgen05)" if tc else "direct SMEM-tiled fp32",
                             "waveforms_per_s": n / t, "ms": t * 1e3, "useful_TFLOP_per_s": flops * n / t / 1e12,
                             "issued_tf32_TFLOP_per_s": issued * n / t / 1e12 if tc else None,
                             "tensor_util_vs_bf16_peak": (issued * n / t / 1e12 / tf) if tc else None,
                             "max_err_of_scale": err, "within_1e-5": bool(err <= 1e-5)})
            del out
    P.TC_CONV_MIN_TAPS = int(os.environ.get("DSPEED_B200_TC_CONV_MIN_TAPS", "128"))
    return {"config": "C5 convolve_wf 'valid', L=8192, kernels cusp / zac / dplms of length K", "rows": n, "sweep": rows,
            "ok": all(r["within_1e-5"] for r in rows)}


def run_all(dev, sel=None):
    sel = sel or os.environ.get("DSPB_CONFIGS", "C1,C4,C5").split(",")
    res = {}
    for name, fn in (("C1", run_c1), ("C4", run_c4), ("C5", run_c5)):
        if name in sel:
            try:
                res[name] = fn(dev)
            except Exception as e:       # a configuration that fails is reported, the headline line still prints
                res[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    out = run_all(torch.device("cuda", 0))
    for k, v in out.items():
        if k == "C5" and "sweep" in v:
            for r in v["sweep"]:
                print(json.dumps({"config": f"C5 {r['kernel']} K={r['K']} ({r['path']})", **r}), flush=True)
        else:
            print(json.dumps(v), flush=True)
