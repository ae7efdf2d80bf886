"""The reference arm of bench.py runs without a GPU (CPU oracle chain on the host cores): check the JSON line it prints
against the contract the driver parses (keys, units, the zero-copy e2e block, the cpu_baseline description)."""

import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-rows", "96"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "waveforms/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("waveforms/s for HPGe DSP chain") and line["value"] > 0
    assert line["steps"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["e2e"] == {"value": line["value"], "unit": "waveforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "96" in cb["sample"]
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
