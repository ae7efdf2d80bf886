"""The reference's SiPM / LAr chain (sipm-dsp-config.json, restated in dspeed_b200/configs/sipm_lar.yaml) end to end
on the device through build_dsp: float64 type loops, a per-event threshold expression (3 * fwhm) feeding the peak
finder, and variable-length VectorOfVectors outputs compacted on the device -- against the sequenced CPU oracle."""
import os

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(REPO, "dspeed_b200", "configs", "sipm_lar.yaml")


def run(vals, block_width=None, device=None):
    import torch

    from dspeed_b200 import tables
    from dspeed_b200.build_dsp import build_dsp

    n = len(vals)
    v = torch.from_numpy(vals).to(device) if device else vals
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=v)
    return build_dsp(tables.Table({"waveform": wf}, size=n), dsp_config=yaml.safe_load(open(CFG)), block_width=block_width)


@pytest.mark.parametrize("block_width", [None, 100])
def test_sipm_lar_chain_matches_oracle(block_width):
    from dspeed_b200 import synth, tables
    from oracle import chains

    d = synth.sipm_waveforms(700, seed=31)
    vals = d["values"].numpy()
    o = chains.sipm_lar_chain(vals)
    out = run(vals, block_width=block_width)
    for name in ("trigger_pos", "energies"):
        assert isinstance(out[name], tables.VectorOfVectors)
        cl = np.asarray(out[name].cumulative_length.nda)
        assert np.array_equal(cl, o["cumulative_length"]), name
    assert o["cumulative_length"][-1] > 500, "the synthetic SiPM rows carry photo-electron pulses"
    tp = np.asarray(out["trigger_pos"].flattened_data.nda)[: o["cumulative_length"][-1]]
    en = np.asarray(out["energies"].flattened_data.nda)[: o["cumulative_length"][-1]]
    assert tp.dtype == np.float64 and out["trigger_pos"].attrs["units"] == "ns"
    assert np.array_equal(tp, o["trigger_pos_flat"])                       # indices: bit-exact
    assert np.abs(en - o["energies_flat"]).max() <= 1e-12 * np.abs(o["energies_flat"]).max()
    # row access of the ragged column
    r = int(np.argmax(o["n_trig"]))
    assert np.array_equal(out["trigger_pos"][r], o["trigger_pos_samples"][r, : o["n_trig"][r]])


def test_vov_compaction_kernels():
    """dspb_vov_offsets / dspb_vov_compact: lengths -> end offsets + cumulative_length, padded block -> ragged data"""
    import ctypes as C

    import torch

    from dspeed_b200 import _lib

    rng = np.random.default_rng(5)
    for n_rows, width, dtype in ((5000, 20, torch.float64), (1, 7, torch.float32), (3333, 33, torch.float32)):
        lens = rng.integers(0, width + 3, n_rows).astype(np.uint32)      # some exceed the width: clamped
        blk = torch.arange(n_rows * width, dtype=dtype, device="cuda").reshape(n_rows, width)
        lens_d = torch.from_numpy(lens.astype(np.int64)).to("cuda").to(torch.uint32)
        ends = torch.empty(n_rows, dtype=torch.int64, device="cuda")
        cum = torch.empty(n_rows, dtype=torch.uint32, device="cuda")
        base = 17
        L = _lib.lib()
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        assert L.dspb_vov_offsets(vp(lens_d.data_ptr()), i64(n_rows), i64(width), i64(base), vp(ends.data_ptr()),
                                  vp(cum.data_ptr()), vp(0)) == 0
        cl = base + np.cumsum(np.minimum(lens, width).astype(np.int64))
        assert np.array_equal(ends.cpu().numpy(), cl) and np.array_equal(cum.cpu().numpy(), cl.astype(np.uint32))
        flat = torch.full((int(cl[-1]) - base + 1,), -1, dtype=dtype, device="cuda")
        assert L.dspb_vov_compact(vp(blk.data_ptr()), i64(blk.stride(0)), i32(blk.element_size()), vp(lens_d.data_ptr()),
                                  i64(width), vp(ends.data_ptr()), i64(n_rows), i64(base), vp(flat.data_ptr()), vp(0)) == 0
        b = blk.cpu().numpy()
        mask = np.arange(width)[None, :] < np.minimum(lens, width)[:, None]
        assert np.array_equal(flat.cpu().numpy()[:-1], b[mask]) and flat[-1].item() == -1
