"""Plan a chain on the meta device (no GPU): print the specialised program, write and
compile the generated kernel.  usage: plan_chain.py [config.yaml] [--no-build]"""
import os
import sys

import numpy as np
import yaml

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dspeed_b200 import codegen, tables  # noqa: E402
from dspeed_b200.processing_chain import build_processing_chain  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cfgp = args[0] if args else os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")
n = 8
wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=np.zeros((n, 8192), np.uint16))
tb = tables.Table({"waveform": wf, "baseline": tables.Array(np.zeros(n, np.uint16))}, size=n)
chain, mask, tb_out = build_processing_chain(yaml.safe_load(open(cfgp)), tb, block_width=16, device="meta")
if "--no-build" in sys.argv:
    codegen.SpecChain._build = lambda self: None
sc = codegen.SpecChain(chain)
print(sc.program_text)
print("slots", sc.n_slots, "smem", sc.smem_bytes, "conv", sc.conv_lowering, "cse", sc.cse_skipped)
open("/tmp/chain_spec.cu", "w").write(sc.source())
print("source: /tmp/chain_spec.cu", getattr(sc, "lib_path", ""))
