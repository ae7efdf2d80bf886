"""Run the ICPC chain through build_dsp on the GPU and print the deviation of every
output from the CPU oracle chain (diagnostic; the assertions live in tests/)."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from dspeed_b200 import synth, tables  # noqa: E402
from dspeed_b200.build_dsp import build_dsp  # noqa: E402
from oracle import chains  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
import yaml  # noqa: E402

cfg = yaml.safe_load(open(os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")))
d = synth.hpge_waveforms(n, seed=11, stress=True)
vals, bl = d["values"].numpy(), d["baseline"].numpy()
wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=vals)
tb = tables.Table({"waveform": wf, "baseline": tables.Array(bl)}, size=n)
t = time.time()
out = build_dsp(tb, dsp_config=cfg, database={"pz": {"tau": "27460.5*16*ns"}}, block_width=int(os.environ.get("BW", 256)))
print("build_dsp", time.time() - t, "s")
o = chains.icpc_chain(vals, bl)
for k in chains.ICPC_OUTPUTS:
    got = np.asarray(out[k].nda)
    ref = o[k].astype(np.float64)
    if k in chains.ICPC_TIME_OUTPUTS:
        ref = ref * 16.0
    ref = ref.astype(np.float32)
    nanm = (np.isnan(ref) != np.isnan(got)).sum()
    ok = ~np.isnan(ref) & ~np.isnan(got)
    diff = np.abs(ref[ok].astype(np.float64) - got[ok])
    nbad = (diff > 0).sum()
    sc = np.abs(ref[ok]).max() if ok.any() else 1
    print(f"{k:14s} nan_mismatch={nanm:3d} n_diff={nbad:4d}/{ok.sum()} max_abs={diff.max() if diff.size else 0:.3e} rel_to_max={diff.max()/sc if diff.size else 0:.2e}")
