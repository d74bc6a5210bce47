#!/usr/bin/env python
"""Attribute the executed warp-instructions of a kernel (ncu SASS page of an .ncu-rep) to CUDA source lines,
using `nvdisasm -gi` line info of the same build (the .so must be the one that was profiled).
   python profiles/linemap.py X.ncu-rep <kernel-substring-of-mangled-name> [units] [--inner]
Default: attribute to the OUTERMOST frame (the line of the kernel body); --inner: innermost frame in csrc/."""
import csv, io, os, re, subprocess, sys, tempfile, glob
from collections import Counter, defaultdict
rep, kname = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
inner = "--inner" in sys.argv
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("DRONECU_LIB", os.path.join(root, "drone_rl_b200", "libdronecu.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
insts = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout
    cur, chain, fresh = None, [], True
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            cur = ln[6:].rstrip(":")
            if kname in cur and insts is None: insts = []; active = True
            else: active = False
            chain = []
            continue
        if cur is None or not (kname in cur): continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if fresh: chain = []; fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            if m.group(3): chain.append((os.path.basename(m.group(3)), int(m.group(4))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and active:
            insts.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
            fresh = True
    if insts: break
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr, rows = None, []
for r in src:
    if r and r[0] == "Address":
        if hdr is not None: break
        hdr = r
    elif hdr is not None and r and r[0].startswith("0x"): rows.append(r)
ie, ins = hdr.index("Instructions Executed"), hdr.index("# Samples")
print(f"# sass rows ncu={len(rows)} nvdisasm={len(insts)}")
n = min(len(rows), len(insts))
cnt, smp = Counter(), Counter()
ours = lambda f: f.endswith((".cuh", ".cu", ".h")) and "crt" not in f
for k in range(n):
    ch = insts[k][2]
    if not ch: key = ("?", 0)
    elif inner:
        o = [c for c in ch if c[0] in os.listdir(os.path.join(root, "drone_rl_b200", "csrc"))]
        key = o[0] if o else ch[-1]
    else: key = ch[-1]
    cnt[key] += int(rows[k][ie]); smp[key] += int(rows[k][ins])
tot = sum(cnt.values())
print(f"total executed {tot}" + (f" = {tot/units:.1f} per unit" if units else ""))
cache = {}
def text(f, l):
    p = os.path.join(root, "drone_rl_b200", "csrc", f)
    if not os.path.isfile(p): return ""
    if f not in cache: cache[f] = open(p).read().splitlines()
    return cache[f][l-1].strip()[:80] if 0 < l <= len(cache[f]) else ""
for (f, l), c in sorted(cnt.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if c * 500 < tot: continue
    print(f"{f:20s}:{l:4d} {100*c/tot:5.1f}% " + (f"{c/units:7.1f} " if units else f"{c:10d} ") + f"smp {smp[(f,l)]:6d}  {text(f,l)}")
