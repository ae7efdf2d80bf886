"""Summarise an ncu --set full capture of the fused kernel per source region (op / routine):
joins the per-SASS-instruction counters of `ncu --page source --csv` with the line table of
the cubin (nvdisasm -g).  usage: ncu_regions.py <report.ncu-rep> [kernel-substring]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "k_chain"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "dspeed_b200", "libdspeed_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
sass = None
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in out and "File" in out:
            sass = out
            break
infn, cur, seq = False, None, []
for line in sass.split("\n"):
    if line.startswith(".text.") or re.match(r"^\s*\.section\s+\.text", line):
        infn = kname in line
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        seq.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))[2:]
src = {}
for f in os.listdir(os.path.join(REPO, "dspeed_b200", "csrc")):
    src[f] = open(os.path.join(REPO, "dspeed_b200", "csrc", f)).read().split("\n")


def region(f, ln):
    if f not in src:
        return f
    for k in range(ln - 1, -1, -1):
        t = src[f][k]
        m = re.search(r"case (OP_\w+):", t)
        if m and f == "fused.cu":
            return "fused:" + m.group(1)
        m = re.search(r"__device__.*?\b(\w+)\s*\(", t)
        if m and not t.strip().startswith("//"):
            return f.split(".")[0] + ":" + m.group(1)
        if "__global__" in t:
            return f.split(".")[0] + ":kernel-main"
    return f


agg, samp = collections.Counter(), collections.Counter()
for i, k in enumerate(seq[: len(rows)]):
    try:
        c, s = int(rows[i][5]), int(rows[i][4])
    except (ValueError, IndexError):
        continue
    r = region(*k) if k else "?"
    agg[r] += c
    samp[r] += s
tot, ts = sum(agg.values()), sum(samp.values())
print(f"warp-instructions executed: {tot}; samples: {ts}")
for k, c in agg.most_common(30):
    print(f"{100 * c / tot:5.1f}% inst  {100 * samp[k] / max(ts, 1):5.1f}% stall-samples  {k}")
