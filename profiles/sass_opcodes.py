#!/usr/bin/env python
"""Per-kernel SASS opcode evidence for libdronecu.so: counts of the tcgen05 / TMEM / TMA / bulk-copy mnemonics
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,
UBLKCP = cp.async.bulk, SYNCS = mbarrier ops) from `cuobjdump -sass`.

    python profiles/sass_opcodes.py [path/to/lib.so] > profiles/sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "drone_rl_b200", "libdronecu.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.TANH",
         "MUFU.EX2", "MUFU.RCP", "HMMA", "FFMA", "FFMA2", "FMUL2", "FADD2", "STG.E.128", "LDG.E.128", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("dronecu::", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):      # FFMA does not count FFMA2
                counts[kern][w] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
print(f"# {'kernel':<58} {'instrs':>7}  " + " ".join(f"{w}" for w in WATCH))
for k, c in counts.items():
    row = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{k:<60} {total[k]:>7}  {row}")
