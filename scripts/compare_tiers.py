import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from tests.test_chain_gpu import run_icpc
from dspeed_b200 import synth
d = synth.hpge_waveforms(20000, seed=321, stress=True)
vals, bl = d["values"].numpy(), d["baseline"].numpy()
a = run_icpc(vals, bl, block_width=20000, device="cuda")
os.environ["DSPEED_B200_SPECIALIZE"] = "0"
b = run_icpc(vals, bl, block_width=20000, device="cuda")
for k in a:
    same = (a[k] == b[k]) | (np.isnan(a[k]) & np.isnan(b[k]))
    rel = np.nanmax(np.abs(a[k].astype(np.float64) - b[k]) / (np.abs(b[k]) + 1e-30)) if not same.all() else 0.0
    print(f"{k:14s} identical {same.mean():.5f}  max rel diff {rel:.2e}")
