"""Parity of the warp-per-waveform chain kernels (dspeed_b200/warpchain.py, BASELINE.json config 4) against the
CPU oracle.  bl_subtract of integer samples and power-of-two boxcars are exact in float32 in any summation order, so
everything here -- smoothed waveforms, extrema index lists, counts -- must be BIT-EXACT
(a16 `get_multi_local_extrema` "bit-exact lists & counts")."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stress_rows(rng, n_rows, wf_len):
    """rows the SiPM generator never makes: dense oscillations (lists overflow), plateaus and exact ties, flat
    rows, steps at the row ends, saw-teeth whose period divides / straddles the lane chunks"""
    t = np.arange(wf_len)
    rows = []
    for k in range(n_rows):
        kind = k % 8
        base = 2000 + rng.integers(0, 1000)
        if kind == 0:    # dense oscillation: far more than 20 extrema
            w = base + 40 * np.sin(t * (0.05 + 0.4 * rng.random())) + rng.normal(0, 3, wf_len)
        elif kind == 1:  # plateaus: ties between equal maxima, first occurrence must win
            w = base + 30 * (np.floor(t / (16 + rng.integers(0, 40))) % 2)
        elif kind == 2:  # flat
            w = np.full(wf_len, base, float)
        elif kind == 3:  # pulses right at the edges
            w = base + rng.normal(0, 2, wf_len)
            w[: 3 + rng.integers(0, 5)] += 60
            w[-(3 + rng.integers(0, 5)):] += 60
        elif kind == 4:  # saw-tooth aligned with 64-sample lane chunks
            w = base + (t % 64) * 1.5
        elif kind == 5:  # saw-tooth straddling chunks, falling
            w = base + 80 - (t % 97) * 0.8 + rng.normal(0, 1, wf_len)
        elif kind == 6:  # random walk
            w = base + np.cumsum(rng.normal(0, 4, wf_len))
        else:            # sparse big pulses + ringing
            w = base + rng.normal(0, 3, wf_len)
            for p in rng.integers(0, wf_len - 120, 6):
                w[p:p + 120] += 80 * np.exp(-np.arange(120) / 30.0) * np.cos(np.arange(120) / 6.0)
        rows.append(np.clip(np.round(w), 0, 65535))
    return np.asarray(rows, np.uint16)


def _run(vals, bl, cfg, block_width=None):
    from dspeed_b200 import tables
    from dspeed_b200.processing_chain import build_processing_chain
    from dspeed_b200.warpchain import WarpChain

    n = len(vals)
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=vals)
    tb = tables.Table({"waveform": wf, "baseline": tables.Array(bl)}, size=n)
    chain, _, tb_out = build_processing_chain(cfg, tb, block_width=block_width, device="cuda")
    assert isinstance(chain._fused, WarpChain), (getattr(chain, "_not_warp_reason", None), type(chain._fused))
    chain(tb, tb_out)
    assert chain.stats["launches"] >= 1
    return {k: np.asarray(v.nda if hasattr(v, "nda") else v.values.nda) for k, v in tb_out.items()}


def _cfg(sdir, m, d_max=12.0, d_min=6.0, a_max=15.0, a_min=1000.0, L=8, num=2, typ=0, extra_outputs=()):
    return {
        "outputs": ["vt_max", "vt_min", "n_max", "n_min", *extra_outputs],
        "processors": {
            "wf_blsub": "dspeed.processors.bl_subtract(waveform, baseline, wf_blsub(unit='ADC'))",
            "wf_mw": {"function": f"dspeed.processors.moving_window_multi(wf_blsub, {L}, {num}, {typ}, wf_mw)", "unit": "ADC"},
            "vt_max, vt_min, n_max, n_min": {
                "function": "get_multi_local_extrema", "module": "dspeed.processors",
                "args": ["wf_mw", d_max, d_min, sdir, a_max, a_min, f"vt_max({m}, 'f')", f"vt_min({m}, 'f')", "n_max", "n_min"],
                "unit": ["ns", "ns", "none", "none"]},
            "t_mn, t_mx, a_mn, a_mx": {"function": "dspeed.processors.min_max(wf_mw, t_mn, t_mx, a_mn, a_mx)",
                                       "unit": ["ns", "ns", "ADC", "ADC"]},
        },
    }


def _oracle(vals, bl, sdir, m, d_max=12.0, d_min=6.0, a_max=15.0, a_min=1000.0, L=8, num=2, typ=0):
    from oracle import oracle as O

    blsub = O.bl_subtract(vals.astype(np.float32), bl.astype(np.float32))
    mw = O.moving_window_multi(blsub, L, num, typ)
    vmax, vmin, nmax, nmin = O.get_multi_local_extrema(mw, d_max, d_min, sdir, a_max, a_min, m)
    return mw, vmax, vmin, nmax, nmin


def _check_lists(out, vmax, vmin, nmax, nmin):
    for k, ref in (("vt_max", vmax), ("vt_min", vmin)):
        got = out[k] / 16.0   # ns -> samples (dt = 16 ns, t0 = 0)
        same = (got == ref) | (np.isnan(got) & np.isnan(ref))
        bad = np.flatnonzero(~same.all(axis=1))
        assert bad.size == 0, f"{k}: {bad.size} rows differ, first {bad[:5]}: got {got[bad[0]]} ref {ref[bad[0]]}"
    assert np.array_equal(out["n_max"], nmax)
    assert np.array_equal(out["n_min"], nmin)


@pytest.mark.parametrize("sdir", [3, 0, 1])
def test_sipm_chain_bit_exact(sdir):
    from dspeed_b200 import synth

    d = synth.sipm_waveforms(6000, seed=31)
    rng = np.random.default_rng(5)
    stress = _stress_rows(rng, 2048, 2000)
    vals = np.concatenate([d["values"].numpy(), stress])
    bl = np.concatenate([d["baseline"].numpy(), np.full(len(stress), 2400, np.uint16)])
    out = _run(vals, bl, _cfg(sdir, 20, extra_outputs=("wf_mw", "a_mx", "t_mx", "a_mn", "t_mn")), block_width=3000)
    mw, vmax, vmin, nmax, nmin = _oracle(vals, bl, sdir, 20)
    assert np.array_equal(out["wf_mw"], mw)
    _check_lists(out, vmax, vmin, nmax, nmin)
    assert nmax.max() == 20 and nmax.min() == 0          # lists overflow and stay empty somewhere
    assert np.array_equal(out["a_mx"], mw.max(axis=1)) and np.array_equal(out["a_mn"], mw.min(axis=1))
    assert np.array_equal(out["t_mx"] / 16.0, mw.argmax(axis=1).astype(np.float32))
    assert np.array_equal(out["t_mn"] / 16.0, mw.argmin(axis=1).astype(np.float32))


@pytest.mark.parametrize("wf_len,m,sdir", [(1000, 5, 3), (256, 3, 3), (2048, 32, 3), (512, 8, 1), (1024, 7, 0)])
def test_other_lengths_and_list_sizes(wf_len, m, sdir):
    rng = np.random.default_rng(wf_len + m)
    vals = _stress_rows(rng, 1024, wf_len)
    bl = rng.integers(2000, 2600, len(vals)).astype(np.uint16)
    kw = dict(d_max=9.0, d_min=4.0, a_max=-50.0, a_min=30.0, L=4, num=3, typ=0)
    out = _run(vals, bl, _cfg(sdir, m, **kw))
    mw, vmax, vmin, nmax, nmin = _oracle(vals, bl, sdir, m, **kw)
    _check_lists(out, vmax, vmin, nmax, nmin)


def test_zero_delta_and_thresholds_that_never_pass():
    rng = np.random.default_rng(77)
    vals = _stress_rows(rng, 512, 2000)
    bl = np.full(len(vals), 2300, np.uint16)
    for kw in (dict(d_max=0.0, d_min=0.0, a_max=-1e9, a_min=1e9), dict(d_max=5.0, d_min=5.0, a_max=1e9, a_min=-1e9)):
        out = _run(vals, bl, _cfg(3, 20, **kw))
        _, vmax, vmin, nmax, nmin = _oracle(vals, bl, 3, 20, **kw)
        _check_lists(out, vmax, vmin, nmax, nmin)


def test_device_resident_columns_and_determinism():
    import torch

    from dspeed_b200 import synth, tables
    from dspeed_b200.processing_chain import build_processing_chain

    n = 50000
    d = synth.sipm_waveforms(n, seed=2, device="cuda")
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=d["values"])
    tb = tables.Table({"waveform": wf, "baseline": tables.Array(d["baseline"])}, size=n)
    chain, _, tb_out = build_processing_chain(_cfg(3, 20), tb, device="cuda")
    outs = []
    for _ in range(2):
        chain(tb, tb_out)
        outs.append({k: np.array(v.nda.cpu() if isinstance(v.nda, torch.Tensor) else v.nda) for k, v in tb_out.items()})
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k
    vals, bl = d["values"].cpu().numpy(), d["baseline"].cpu().numpy()
    _, vmax, vmin, nmax, nmin = _oracle(vals[:8000], bl[:8000], 3, 20)
    _check_lists({k: v[:8000] for k, v in outs[0].items()}, vmax, vmin, nmax, nmin)
