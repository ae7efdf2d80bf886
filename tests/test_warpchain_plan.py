"""Host-side checks of the warp-per-waveform chain generator (no GPU): BASELINE.json config 4 (SiPM chain) is planned
on the "meta" device, the generated kernel is inspected and compiled for sm_100a."""

import os
import shutil

import numpy as np
import pytest

from dspeed_b200 import tables, warpchain
from dspeed_b200.fusion import NotFusable
from dspeed_b200.processing_chain import build_processing_chain

SIPM = {
    "outputs": ["vt_max", "vt_min", "n_max", "n_min"],
    "processors": {
        "wf_blsub": "dspeed.processors.bl_subtract(waveform, baseline, wf_blsub(unit='ADC'))",
        "wf_mw": {"function": "dspeed.processors.moving_window_multi(wf_blsub, 8, 2, 0, wf_mw)", "unit": "ADC"},
        "vt_max, vt_min, n_max, n_min": {
            "function": "get_multi_local_extrema", "module": "dspeed.processors",
            "args": ["wf_mw", 12.0, 6.0, 3, 15.0, 1000.0, "vt_max(20, 'f')", "vt_min(20, 'f')", "n_max", "n_min"],
            "unit": ["ns", "ns", "none", "none"]},
    },
}


def _plan(cfg, wf_len):
    n = 4
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=np.zeros((n, wf_len), np.uint16))
    tb = tables.Table({"waveform": wf, "baseline": tables.Array(np.zeros(n, np.uint16))}, size=n)
    chain, _, _ = build_processing_chain(cfg, tb, block_width=16, device="meta")
    build = warpchain.WarpChain._build
    warpchain.WarpChain._build = lambda self: None     # plan only
    try:
        return warpchain.WarpChain(chain)
    finally:
        warpchain.WarpChain._build = build


def test_sipm_program():
    wc = _plan(SIPM, 2000)
    text = wc.program_text
    assert "bl_sub" in text and "mw L=8 lr" in text and "get_multi_local_extrema dir=3 m=20" in text
    assert wc.CH == 64
    src = wc.source()
    # one warp per waveform: no block-wide barrier, raw row staged asynchronously, both walks present
    assert "__syncthreads" not in src and "stage_row_16<CH>" in src
    assert "peak_walk<CH, false>" in src and "peak_walk<CH, true>" in src
    # the staging buffer is aliased with the wave copy of the peak finder (the next row travels through registers:
    # requested before the walk, staged after it), so 24 warps per SM fit the shared memory
    assert wc.alias and "fetch_row_16<2000>" in src and "stage_pieces_16<CH, 2000>" in src
    assert src.index("fetch_row_16<2000>") < src.index("peak_walk<CH, false>") < src.index("stage_pieces_16<CH, 2000>")
    assert wc.ctas_per_sm * warpchain.WARPS_PER_CTA == 24
    assert wc.ctas_per_sm * (wc.smem_bytes + 1024) <= 227 * 1024
    # vector outputs are converted to ns per element and stored by lanes < m
    assert src.count("if (lane < 20)") == 2


def test_lane_chunk_follows_the_waveform_length():
    assert _plan(SIPM, 1000).CH == 32
    assert _plan(SIPM, 256).CH == 8


def test_chains_outside_the_tier_are_refused():
    with pytest.raises(NotFusable):
        _plan(SIPM, 8192)       # long waveforms: block-per-waveform kernels
    cfg = {"outputs": ["wf_pz"], "processors": {
        "wf_blsub": "dspeed.processors.bl_subtract(waveform, baseline, wf_blsub(unit='ADC'))",
        "wf_pz": {"function": "dspeed.processors.pole_zero(wf_blsub, 1000*ns, wf_pz)", "unit": "ADC"}}}
    with pytest.raises(NotFusable):
        _plan(cfg, 2000)        # no warp-tier emitter
    bad = {**SIPM, "processors": {**SIPM["processors"]}}
    bad["processors"]["vt_max, vt_min, n_max, n_min"] = {
        **SIPM["processors"]["vt_max, vt_min, n_max, n_min"],
        "args": ["wf_mw", 12.0, 6.0, 2, 15.0, 1000.0, "vt_max(20, 'f')", "vt_min(20, 'f')", "n_max", "n_min"]}
    with pytest.raises(NotFusable):
        _plan(bad, 2000)        # search direction 2 stays on the per-processor kernel


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_generated_kernel_compiles_for_sm100a():
    path, text = warpchain.prebuild(SIPM)
    assert path and os.path.exists(path) and os.path.getsize(path) > 10000
