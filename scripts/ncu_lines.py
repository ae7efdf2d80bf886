"""Per-routine breakdown of an `ncu --set full --import-source on` capture of a specialised chain
kernel: joins the per-SASS-instruction counters of `ncu --page source --csv` with the line table
(nvdisasm -g) of the same kernel built locally.
usage: ncu_lines.py <report.ncu-rep> <chain .so> [top]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
seq, cur = [], None
for line in dis.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        seq.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]
rows = [r for r in rows[2:] if len(r) > 5]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
cop = hdr.index("Source")
if len(rows) != len(seq):
    print(f"warning: {len(rows)} profiled instructions vs {len(seq)} disassembled", file=sys.stderr)
srcs = {}


def func_of(f, ln):
    if f not in srcs:
        try:
            srcs[f] = open(f).read().split("\n")
        except OSError:
            alt = os.path.join(os.path.dirname(os.path.abspath(so)), "chain_spec.cu")
            srcs[f] = open(alt).read().split("\n") if os.path.basename(f).startswith("chain_") and os.path.exists(alt) else []
    lines = srcs[f]
    base = os.path.basename(f)
    if base.startswith("chain_") and base.endswith(".cu"):
        # generated file: attribute to the node comment above the line
        for k in range(min(ln, len(lines)) - 1, -1, -1):
            m = re.search(r"// ---- \[(\d+)\] (\S+)", lines[k])
            if m:
                return f"gen:[{m.group(1)}]{m.group(2)}"
        return "gen:prologue"
    for k in range(min(ln, len(lines)) - 1, -1, -1):
        t = lines[k]
        m = re.search(r"__device__.*?\b(\w+)\s*\(", t)
        if m and not t.strip().startswith("//"):
            return base.split(".")[0] + ":" + m.group(1)
    return base


inst, samp, ops = collections.Counter(), collections.Counter(), collections.Counter()
for r, k in zip(rows, seq):
    try:
        c, s = int(r[ci]), int(r[cs])
    except ValueError:
        continue
    key = func_of(*k) if k else "?"
    inst[key] += c
    samp[key] += s
    ops[r[cop].split()[0].split(".")[0] if r[cop].split()[0][0] != "@" else r[cop].split()[1].split(".")[0]] += c
ti, ts = sum(inst.values()), sum(samp.values())
print(f"warp instructions: {ti}   stall samples: {ts}")
for k, c in inst.most_common(top):
    print(f"{100 * c / ti:5.1f}% inst {100 * samp[k] / max(ts, 1):5.1f}% samples  {k}")
print("-- opcode mix")
for k, c in ops.most_common(25):
    print(f"{100 * c / ti:5.1f}%  {k}")

# optional: per-source-line counts of one file (4th argument = file basename, e.g. warp_rt.cuh)
if len(sys.argv) > 4:
    want = sys.argv[4]
    per_line = collections.Counter()
    per_line_s = collections.Counter()
    for r, k in zip(rows, seq):
        if not k or os.path.basename(k[0]) != want:
            continue
        try:
            per_line[k[1]] += int(r[ci])
            per_line_s[k[1]] += int(r[cs])
        except ValueError:
            pass
    lines = srcs.get([f for f in srcs if os.path.basename(f) == want][0], [])
    print(f"-- per line of {want} (% of all warp instructions, % of samples)")
    for ln in sorted(per_line):
        if per_line[ln] * 1000 > ti or per_line_s[ln] * 300 > ts:
            print(f"{100 * per_line[ln] / ti:5.2f}% {100 * per_line_s[ln] / max(ts, 1):5.2f}%  {ln:5d}: {lines[ln - 1].strip()[:110] if ln <= len(lines) else ''}")
