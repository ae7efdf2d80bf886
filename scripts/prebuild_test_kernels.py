"""Pre-compile (here, no GPU) the warp-tier kernels the GPU tests will ask for, so the GPU box finds them in the
in-tree cache (dspeed_b200/_chains/) instead of spending box time in nvcc."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dspeed_b200 import warpchain  # noqa: E402
from tests.test_warpchain_gpu import _cfg  # noqa: E402

jobs = [(2000, _cfg(s, 20, extra_outputs=("wf_mw", "a_mx", "t_mx", "a_mn", "t_mn"))) for s in (3, 0, 1)]
kw = dict(d_max=9.0, d_min=4.0, a_max=-50.0, a_min=30.0, L=4, num=3, typ=0)
jobs += [(n, _cfg(s, m, **kw)) for (n, m, s) in [(1000, 5, 3), (256, 3, 3), (2048, 32, 3), (512, 8, 1), (1024, 7, 0)]]
jobs += [(2000, _cfg(3, 20, **k)) for k in (dict(d_max=0.0, d_min=0.0, a_max=-1e9, a_min=1e9),
                                            dict(d_max=5.0, d_min=5.0, a_max=1e9, a_min=-1e9))]
jobs += [(2000, _cfg(3, 20))]
for n, cfg in jobs:
    print(n, warpchain.prebuild(cfg, wf_len=n)[0])

# alternative lowerings of the ICPC chain exercised by tests/test_chain_gpu.py::test_code_generation_variants_agree
import yaml  # noqa: E402

from dspeed_b200 import codegen  # noqa: E402

icpc = yaml.safe_load(open(os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")))
for var in ("DSPEED_B200_CONV_HELPERS", "DSPEED_B200_FILL_WAIT"):
    os.environ[var] = "0"
    try:
        print(var, codegen.prebuild(icpc)[0])
    finally:
        os.environ.pop(var, None)

# small specialised chains of tests/test_chain_gpu.py
dpz = {
    "outputs": ["wf_dpz", "dpz_max", "t_max", "trapEmax"],
    "processors": {
        "wf_blsub": "dspeed.processors.bl_subtract(waveform, baseline, wf_blsub(unit='ADC'))",
        "wf_dpz": {"function": "dspeed.processors.double_pole_zero(wf_blsub, 27460.5, 1200.25, 0.025, wf_dpz)", "unit": "ADC"},
        "t_min, t_max, dpz_min, dpz_max": {"function": "dspeed.processors.min_max(wf_dpz, t_min, t_max, dpz_min, dpz_max)",
                                           "unit": ["ns", "ns", "ADC", "ADC"]},
        "wf_trap": {"function": "dspeed.processors.trap_norm(wf_dpz, 10*us, 3.008*us, wf_trap)", "unit": "ADC"},
        "trapEmax": {"function": "numpy.amax(wf_trap, 1, trapEmax)", "unit": "ADC"},
    },
}
tiny = {
    "outputs": ["wf_pz", "pz_max"],
    "processors": {
        "wf_blsub": "dspeed.processors.bl_subtract(waveform, baseline, wf_blsub(unit='ADC'))",
        "wf_pz": {"function": "dspeed.processors.pole_zero(wf_blsub, 27460.5, wf_pz)", "unit": "ADC"},
        "pz_max": {"function": "numpy.amax(wf_pz, 1, pz_max)", "unit": "ADC"},
    },
}
for cfg in (dpz, tiny):
    print(codegen.prebuild(cfg)[0])

# the minimal energy chain of BASELINE.json config 1 (scripts/bench_configs.py measures it)
from scripts.bench_configs import C1_CFG  # noqa: E402

print(codegen.prebuild(C1_CFG, with_baseline=False)[0])
