#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell paths (tcgen05 MMA, TMA, TMEM) in libtmr_b200.so:
sass_summary.py [lib] > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tmrnet_b200", "libtmr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "FFMA2", "SHFL", "MUFU", "LDG", "STG", "F2FP"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = cur.replace("tmr::umma::", "").replace("tmr::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# SASS opcode summary of `{os.path.basename(lib)}` (cuobjdump -sass, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `UTMALDG`/`UTMASTG` = TMA tensor load / store, `LDTM` = tcgen05.ld (TMEM -> registers), "
      "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `FFMA2` = packed fp32x2 FMA.\n")
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
for k, c in counts.items():
    if total[k] == 0:
        continue
    print(f"| `{k}` | {total[k]} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")
