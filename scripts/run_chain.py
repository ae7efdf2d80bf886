"""Run the ICPC chain on synthetic device-resident waveforms (profiling driver).
usage: run_chain.py <rows> <repeats> [block_width]"""
import os
import sys
import time

import torch
import yaml

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dspeed_b200 import synth, tables  # noqa: E402
from dspeed_b200.processing_chain import build_processing_chain  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bw = int(sys.argv[3]) if len(sys.argv) > 3 else None
dev = torch.device("cuda", 0)
d = synth.hpge_waveforms(n, seed=1, device=dev, stress=True)
wf = tables.WaveformTable(size=n, t0=tables.Array(d["t0"], attrs={"units": "ns"}), dt=tables.Array(d["dt"], attrs={"units": "ns"}),
                          values=d["values"])
tb = tables.Table({"waveform": wf, "baseline": tables.Array(d["baseline"])}, size=n)
cfg = yaml.safe_load(open(os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")))
chain, _, tb_out = build_processing_chain(cfg, tb, block_width=bw, device=dev)
print("fused:", chain._fused is not None, getattr(chain, "_not_fused_reason", ""))
if chain._fused is not None and os.environ.get("SHOW_PROGRAM"):
    print(chain._fused.program_text)
out = tables.Table({k: tables.Array(torch.empty(n, dtype=torch.float32, device=dev)) for k in tb_out}, size=n)
for i in range(reps):
    torch.cuda.synchronize()
    t = time.perf_counter()
    chain(tb, out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"pass {i}: {dt * 1e3:.2f} ms  {n / dt / 1e6:.3f} M wf/s")
if chain._fused is not None and os.environ.get("PROFILE_PROGRAM"):
    from dspeed_b200 import fusion
    prof = fusion.profile_fused(chain, lambda: chain(tb, out))
    tot = sum(p[0] for p in prof)
    rows_cta0 = -(-n // 148) if (bw or 16384) >= n else None
    print(f"per-instruction cycles of CTA 0 (total {tot:.0f} cycles; smem {chain._fused.smem_bytes} B, slots {chain._fused.n_slots})")
    for c, share, text in prof:
        print(f"{c:12.0f} {100 * share:5.1f}%  {text}")
if chain._fused is not None and os.environ.get("SAVE_KERNEL") and hasattr(chain._fused, "lib_path"):
    import shutil
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    shutil.copy(chain._fused.lib_path, os.path.join(REPO, "gpurun_out", "chain_spec.so"))
    shutil.copy(chain._fused.src_path, os.path.join(REPO, "gpurun_out", "chain_spec.cu"))
