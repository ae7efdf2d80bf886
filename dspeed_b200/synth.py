"""Deterministic synthetic raw data shared by the parity tests, the CPU baseline and
``bench.py``: HPGe-like ``uint16`` waveforms with the DAQ
``baseline`` column, and SiPM-like short traces.  Pure torch so the same code fills a
host tensor (parity tests: identical inputs for the oracle and the CUDA path) or a
device tensor (full-size benchmark blocks).
"""

from __future__ import annotations

import math

import torch

HPGE_TAU_SAMPLES = 27460.5  # matches the reference's test database (tests/test_build_dsp.py:23)


def hpge_waveforms(
    n_rows: int,
    wf_len: int = 8192,
    seed: int = 1234,
    device: str | torch.device = "cpu",
    stress: bool = False,
    chunk: int = 16384,
) -> dict[str, torch.Tensor]:
    """HPGe charge-sensitive-preamp pulses.

    Per row: baseline ~ U(10000, 15000); gaussian noise sigma 4 ADC; one pulse at
    t0 ~ U{3800..4200} with amplitude ~ U(500, 20000), saturating-exponential rise
    with time constant ~ U(5, 40) samples, exponential decay tau = 27460.5 samples;
    rounded and clipped to [0, 65535].  ``stress=True`` additionally makes ~1 % of the
    rows pile-up (second pulse), ~0.1 % saturated and ~0.5 % pulse-free (NaN paths).

    Returns ``{"values": uint16 [n, L], "baseline": uint16 [n], "t0": f64 [n] (=0),
    "dt": f64 [n] (=16 ns)}``.
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    values = torch.empty((n_rows, wf_len), dtype=torch.uint16, device=dev)
    baseline = torch.empty((n_rows,), dtype=torch.uint16, device=dev)
    t = torch.arange(wf_len, device=dev, dtype=torch.float32)[None, :]
    scale = wf_len / 8192.0
    for lo in range(0, n_rows, chunk):
        hi = min(lo + chunk, n_rows)
        m = hi - lo
        u = torch.rand((m, 8), generator=g, device=dev)
        bl = 10000.0 + 5000.0 * u[:, 0:1]
        amp = 500.0 + 19500.0 * u[:, 1:2]
        t0 = torch.floor((3800.0 + 401.0 * u[:, 2:3]) * scale)
        rise = 5.0 + 35.0 * u[:, 3:4]

        def pulse(a, t_start, tr):
            x = (t - t_start).clamp_min(0.0)
            on = (t >= t_start).to(torch.float32)
            return a * on * (1.0 - torch.exp(-x / tr)) * torch.exp(-x / HPGE_TAU_SAMPLES)

        wf = bl + pulse(amp, t0, rise)
        if stress:
            pile = (u[:, 4:5] < 0.01).to(torch.float32)
            t1 = t0 + torch.floor((200.0 + 2500.0 * u[:, 5:6]) * scale)
            wf = wf + pile * pulse(0.6 * amp, t1, rise)
            sat = (u[:, 6:7] < 0.001).to(torch.float32)
            wf = wf + sat * pulse(60000.0 * torch.ones_like(amp), t0, rise)
            flat = (u[:, 7:8] < 0.005).to(torch.float32)
            wf = flat * bl + (1.0 - flat) * wf
        wf = wf + 4.0 * torch.randn((m, wf_len), generator=g, device=dev)
        values[lo:hi] = torch.round(wf).clamp_(0, 65535).to(torch.int32).to(torch.uint16)
        baseline[lo:hi] = torch.round(bl[:, 0]).to(torch.int32).to(torch.uint16)
    return {
        "values": values,
        "baseline": baseline,
        "t0": torch.zeros(n_rows, dtype=torch.float64, device=dev),
        "dt": torch.full((n_rows,), 16.0, dtype=torch.float64, device=dev),
    }


def sipm_waveforms(
    n_rows: int,
    wf_len: int = 2000,
    seed: int = 4321,
    device: str | torch.device = "cpu",
    chunk: int = 65536,
    max_pulses: int = 8,
) -> dict[str, torch.Tensor]:
    """SiPM-like traces: baseline ~ U(2000, 3000), noise sigma 3 ADC, Poisson(2)
    single-photo-electron pulses (amplitude ~ N(40, 8) ADC, 5-sample rise,
    100-sample decay) at uniform times."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    values = torch.empty((n_rows, wf_len), dtype=torch.uint16, device=dev)
    baseline = torch.empty((n_rows,), dtype=torch.uint16, device=dev)
    t = torch.arange(wf_len, device=dev, dtype=torch.float32)[None, :]
    # Poisson(2) CDF thresholds for "pulse k present" (k = 0..max_pulses-1)
    pmf = [math.exp(-2.0) * 2.0**k / math.factorial(k) for k in range(max_pulses + 1)]
    cdf = torch.tensor([sum(pmf[: k + 1]) for k in range(max_pulses)], device=dev)
    for lo in range(0, n_rows, chunk):
        hi = min(lo + chunk, n_rows)
        m = hi - lo
        bl = 2000.0 + 1000.0 * torch.rand((m, 1), generator=g, device=dev)
        wf = bl.expand(m, wf_len).clone()
        npe_u = torch.rand((m, 1), generator=g, device=dev)
        present = (npe_u > cdf[None, :]).to(torch.float32)  # [m, max_pulses]
        tpos = torch.floor(torch.rand((m, max_pulses), generator=g, device=dev) * (wf_len - 50))
        amp = 40.0 + 8.0 * torch.randn((m, max_pulses), generator=g, device=dev)
        for k in range(max_pulses):
            x = (t - tpos[:, k : k + 1]).clamp_min(0.0)
            on = (t >= tpos[:, k : k + 1]).to(torch.float32)
            wf = wf + present[:, k : k + 1] * amp[:, k : k + 1] * on * (
                1.0 - torch.exp(-x / 5.0)
            ) * torch.exp(-x / 100.0) * 1.3
        wf = wf + 3.0 * torch.randn((m, wf_len), generator=g, device=dev)
        values[lo:hi] = torch.round(wf).clamp_(0, 65535).to(torch.int32).to(torch.uint16)
        baseline[lo:hi] = torch.round(bl[:, 0]).to(torch.int32).to(torch.uint16)
    return {
        "values": values,
        "baseline": baseline,
        "t0": torch.zeros(n_rows, dtype=torch.float64, device=dev),
        "dt": torch.full((n_rows,), 16.0, dtype=torch.float64, device=dev),
    }
