"""Throughput of the other BASELINE.json configurations on one B200 (device-resident inputs,
CUDA-event timing, >= 3 warm-up passes): C1 minimal energy chain, C4 SiPM chain, C5 long-kernel
convolution sweep.  Prints one JSON line per measurement (kept under profiles/)."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dspeed_b200 import processors as P, synth, tables  # noqa: E402
from dspeed_b200.processing_chain import build_processing_chain  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0


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


def chain_bench(name, cfg, tb, n, bytes_per_wf, out_cols):
    chain, _, tb_out = build_processing_chain(cfg, tb, device=dev)
    out = tables.Table({k: (tables.Array(torch.empty(n, dtype=torch.float32, device=dev)) if k in out_cols else tb_out[k])
                        for k in tb_out}, size=n)
    t = timed(lambda: chain(tb, out))
    tier = type(chain._fused).__name__ if chain._fused is not None else "per-processor kernels"
    print(json.dumps({"config": name, "rows": n, "waveforms_per_s": n / t, "ms": t * 1e3, "tier": tier,
                      "algorithmic_bytes_per_wf": bytes_per_wf, "frac_of_hbm_roofline": n / t * bytes_per_wf / 1e9 / PEAK}), flush=True)


SEL = os.environ.get("DSPB_CONFIGS", "C1,C4,C5").split(",")   # e.g. DSPB_CONFIGS=C4

if "C1" in SEL:
    # ---- C1: minimal energy chain ---------------------------------------------------------------------
    n = 262144
    d = synth.hpge_waveforms(n, seed=5, device=dev)
    wf = tables.WaveformTable(size=n, t0=tables.Array(d["t0"], attrs={"units": "ns"}), dt=tables.Array(d["dt"], attrs={"units": "ns"}),
                              values=d["values"])
    cfg1 = {
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
    chain_bench("C1 minimal energy chain (lsf + bl_subtract + pole_zero + trap_norm + min_max + trap_pickoff), L=8192",
                cfg1, tables.Table({"waveform": wf}, size=n), n, 8192 * 2 + 2 + 4 * 5, cfg1["outputs"])
    del d, wf
    torch.cuda.empty_cache()

if "C4" in SEL:
    # ---- C4: SiPM chain ---------------------------------------------------------------------------------
    n = int(os.environ.get("DSPB_C4_ROWS", 1 << 22))   # BASELINE.json config 4: 4 M short waveforms
    d = synth.sipm_waveforms(n, seed=9, device=dev)
    wf = tables.WaveformTable(size=n, t0=tables.Array(torch.zeros(n, dtype=torch.float64, device=dev), attrs={"units": "ns"}),
                              dt=tables.Array(torch.full((n,), 16.0, dtype=torch.float64, device=dev), attrs={"units": "ns"}),
                              values=d["values"])
    import yaml

    cfg4 = yaml.safe_load(open(os.path.join(REPO, "dspeed_b200", "configs", "sipm_peaks.yaml")))
    tb = tables.Table({"waveform": wf, "baseline": tables.Array(d["baseline"])}, size=n)
    chain, _, tb_out = build_processing_chain(cfg4, tb, device=dev, block_width=int(os.environ.get("DSPB_C4_BLOCK", 0)) or None)
    # device-resident output columns (written in place by the kernel), like the ICPC bench
    out4 = tables.Table({k: type(v)(torch.empty(tuple(v.nda.shape), dtype=getattr(torch, str(v.nda.dtype)), device=dev),
                                    attrs=dict(v.attrs)) for k, v in tb_out.items()}, size=n)
    t = timed(lambda: chain(tb, out4), reps=3, warm=3)
    L = d["values"].shape[1]
    b4 = L * 2 + 2 + 4 * 40 + 8
    print(json.dumps({"config": f"C4 SiPM chain (bl_subtract + moving_window_multi + get_multi_local_extrema), L={L}, device-resident in/out",
                      "rows": n, "block_width": chain._block_width, "waveforms_per_s": n / t, "ms": t * 1e3,
                      "tier": type(chain._fused).__name__ if chain._fused is not None else "per-processor kernels",
                      "algorithmic_bytes_per_wf": b4, "achieved_GBps": n / t * b4 / 1e9,
                      "frac_of_hbm_roofline": n / t * b4 / 1e9 / PEAK}), flush=True)
    del out4
    del d, wf, tb, chain
    torch.cuda.empty_cache()

if "C5" in SEL:
    # ---- C5: long-kernel 'valid' convolution sweep (direct SMEM-tiled kernel) ------------------------------------
    n = 65536
    d = synth.hpge_waveforms(n, seed=11, device=dev)
    x = (d["values"].to(torch.float32) - d["baseline"].to(torch.float32)[:, None]).contiguous()
    rng = np.random.default_rng(0)
    for K in (256, 512, 1024, 2048, 4096):
        k = torch.from_numpy(rng.standard_normal(K).astype(np.float32)).to(dev)
        out = torch.empty((n, 8192 - K + 1), dtype=torch.float32, device=dev)
        flops = 2.0 * K * (8192 - K + 1)
        for tc in (False, True):
            P.TC_CONV_MIN_TAPS = 1 if tc else 0
            t = timed(lambda: P.convolve_wf(x, k, np.int8(ord("v")), out), reps=2, warm=3)
            # tensor-pipe work actually issued: 3 TF32 products over the (K + 127)-wide band of every 128-output tile
            issued = 3 * 2.0 * 128 * (-(-(K + 127) // 32) * 32) * (-(-(8192 - K + 1) // 128)) if tc else None
            how = ("tensor cores: banded Toeplitz GEMM, 3xTF32, tcgen05 + TMEM + TMA" if tc
                   else "direct, SMEM-tiled, fp32 FMA + fp64 chunk sums")
            print(json.dumps({"config": f"C5 convolve_wf 'valid', generic kernel K={K}, L=8192 ({how})",
                              "rows": n, "waveforms_per_s": n / t, "ms": t * 1e3, "useful_TFLOP_per_s": flops * n / t / 1e12,
                              "issued_tf32_TFLOP_per_s": issued * n / t / 1e12 if tc else None}), flush=True)
        P.TC_CONV_MIN_TAPS = int(os.environ.get("DSPEED_B200_TC_CONV_MIN_TAPS", "128"))

