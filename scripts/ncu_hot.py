#!/usr/bin/env python
"""Hottest SASS instructions of an `ncu --set full --import-source on` capture (source page, SASS view):
ncu_hot.py <report.ncu-rep> [top N]  -> address, samples, share, dominant stall reasons, instruction."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r["# Samples"] or 0) for r in rows)
rows_s = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))
print(f"total samples {tot}, instructions {len(rows)}")
for r in rows_s[:top]:
    n = int(r["# Samples"] or 0)
    st = sorted(((int(r[k] or 0), k[6:]) for k in stalls), reverse=True)[:3]
    print(f"{r['Address'][-5:]} {n:6d} {100*n/tot:5.1f}%  {' '.join(f'{k}:{v}' for v, k in st if v):40s} {r['Source'][:90]}")
agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
print("stall totals:", " ".join(f"{k[6:]}:{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
