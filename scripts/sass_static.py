"""Static instruction budget of a generated chain kernel (no GPU needed).

Compiles the generated ``.cu`` to a cubin (sm_100a, -lineinfo), disassembles it with
``nvdisasm -gi`` and attributes every SASS instruction of the kernel to (a) the node of the
generated program that the outermost source line belongs to and the stream it runs in (block
warps x16, scalar warp x1) and (b) the innermost run-time routine.  The generated kernels are
straight-line code, so the static count weighted by the number of warps that execute a stream is an
upper bound of the warp-instructions per waveform that ncu reports (divergent early-outs and
helper-thread branches are counted in full).

usage: sass_static.py <chain.cu> [--nodes] [--routines] [--ops] [--node K]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble(cu, extra=()):
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, "k.cubin")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin",
           "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "dspeed_b200", "csrc"), *extra, "-o", cubin, cu]
    r = subprocess.run(cmd + ["-Xptxas", "-v"], env=env, capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr[-3000:])
    regs = re.findall(r"Used (\d+) registers", r.stderr)
    spills = re.findall(r"(\d+) bytes spill stores", r.stderr)
    return subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout, regs, spills


def analyse(cu, extra=()):
    sass, regs, spills = disassemble(cu, extra)
    src = open(cu).read().split("\n")
    base = os.path.basename(cu)
    # generated line -> (stream, node label)
    where = {}
    stream, node = "P", "prologue"
    region = None      # default warp weight inside a short-waveform region
    for i, ln in enumerate(src, 1):
        mr = re.search(r"if \(warp < (\d+)\) \{   // ---- region", ln)
        if mr:
            region = ("body", int(mr.group(1)))
        elif region and "// warps without a chunk of these waveforms" in ln:
            region = ("idle", 16 - region[1])
        elif region and region[0] == "idle" and ln.strip() == "}":
            region = None
        if "block stream (warps" in ln:
            stream = "B"
        elif "scalar stream (warp" in ln:
            stream = "S"
        elif ln.startswith("  // consume the scalar warp"):
            stream, node = "P", "epilogue"
        m = re.match(r"\s*// ---- \[(\d+)\] (.*)", ln)
        if m:
            node = f"[{int(m.group(1)):2d}] {m.group(2)[:70]}"
        mx = re.search(r"//@X (\d+)", ln)
        wt = int(mx.group(1)) if mx else (region[1] if region else None)
        if mx and region:
            wt = min(wt, region[1])
        where[i] = (stream, node, wt)
    hdr = {}
    for f in os.listdir(os.path.join(REPO, "dspeed_b200", "csrc")):
        hdr[f] = open(os.path.join(REPO, "dspeed_b200", "csrc", f)).read().split("\n")

    def routine(f, ln):
        f = os.path.basename(f)
        if f not in hdr:
            return f
        for k in range(min(ln, len(hdr[f])) - 1, -1, -1):
            t = hdr[f][k]
            m = re.search(r"__device__.*?\b(\w+)\s*\(", t)
            if m and not t.strip().startswith("//"):
                return m.group(1)
        return f

    rows = []   # (section, stream, node, routine, opcode)
    section, cur, fresh = None, None, True
    for line in sass.split("\n"):
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            section, cur = m.group(1), None
            continue
        if "//## File" in line:
            parts = re.findall(r'File "([^"]+)", line (\d+)', line)
            if fresh:       # first marker after an instruction: innermost location first
                cur, fresh = [], False
            cur = (cur or []) + [(p, int(n)) for p, n in parts]
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and section:
            fresh = True
            op = m.group(1).split(".")[0]
            if cur is None:
                rows.append((section, "?", "?", "?", op, None))
                continue
            outer = next(((p, n) for p, n in reversed(cur) if os.path.basename(p) == base), None)
            inner = cur[0]
            st, nd, wx = where.get(outer[1], ("?", "?", None)) if outer else ("F", section[:40], None)
            rt = routine(*inner) if os.path.basename(inner[0]) != base else "gen"
            rows.append((section, st, nd, rt, op, wx))
    return rows, regs, spills


def main():
    cu = sys.argv[1]
    extra = [a for a in sys.argv[2:] if a.startswith("-D")]
    rows, regs, spills = analyse(cu, extra)
    w0 = {"B": 16, "S": 1, "P": 17, "?": 1, "F": 1}
    # lines tagged //@X n by the generator run in n warps only
    rows = [(sec, st, nd, rt, op, (wx if (wx is not None and st == "B") else w0[st])) for sec, st, nd, rt, op, wx in rows]
    tot = collections.Counter()
    for sec, st, nd, rt, op, wt in rows:
        tot[st] += 1
    print(f"registers {regs}  spill-store bytes {spills}")
    print("static SASS instructions per stream:", dict(tot))
    est = sum(wt for _, st, _, _, _, wt in rows if st in "BS")
    print(f"upper bound warp-instructions / waveform (16 x block + 1 x scalar): {est}")
    by_node = collections.Counter()
    for sec, st, nd, rt, op, wt in rows:
        if st in "BS":
            by_node[(st, nd)] += wt
    print("-- per node (weighted)")
    for (st, nd), c in sorted(by_node.items(), key=lambda kv: -kv[1])[:40]:
        print(f"{c:7d} {100 * c / est:5.1f}%  {st} {nd}")
    by_rt = collections.Counter()
    for sec, st, nd, rt, op, wt in rows:
        if st in "BS":
            by_rt[rt] += wt
    print("-- per routine (weighted)")
    for rt, c in by_rt.most_common(30):
        print(f"{c:7d} {100 * c / est:5.1f}%  {rt}")
    by_op = collections.Counter()
    for sec, st, nd, rt, op, wt in rows:
        if st in "BS":
            by_op[op] += wt
    print("-- opcode mix (weighted)")
    print("  ".join(f"{op} {100 * c / est:.1f}" for op, c in by_op.most_common(28)))
    for a in sys.argv[2:]:
        if a.startswith("--node="):
            key = a.split("=", 1)[1]
            sel = collections.Counter()
            for sec, st, nd, rt, op, wt in rows:
                if st == "B" and nd.startswith(key):
                    sel[(rt, op)] += 1
            print(f"-- node {key}: routine/opcode (per warp)")
            for (rt, op), c in sel.most_common(40):
                print(f"{c:6d}  {rt:24s} {op}")


if __name__ == "__main__":
    main()
