#!/usr/bin/env python
"""Turn an `ncu --set full` capture of the LSTM recurrence kernel into profiles/lstm_traffic.json, the file
bench.py reads `roofline.traffic` from (so the number in the bench line is the committed measurement of the
CURRENT kernel, not a literal in bench.py).

    python scripts/ncu_traffic.py gpurun_out/prof_lstm.ncu-rep <clips per launch> <recurrent steps per launch> \
        <algorithmic bytes per clip and step> "<kernel description>" [kernel-name regex]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, clips, steps, alg, desc = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
pat = re.compile(sys.argv[6] if len(sys.argv) > 6 else "lstm")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6,
                "nsecond": 1e-9, "msecond": 1e-3, "second": 1}.get(u, 1)


launches = []
for r in rows[2:]:
    if len(r) < len(hdr) or not pat.search(r[col["Kernel Name"]]):
        continue
    launches.append({"kernel": r[col["Kernel Name"]], "dram_read": val(r, "dram__bytes_read.sum"),
                     "dram_write": val(r, "dram__bytes_write.sum"), "duration_s": val(r, "gpu__time_duration.sum")})
tot = sum(l["dram_read"] + l["dram_write"] for l in launches) / len(launches)
res = {"kernel": desc, "source": f"ncu --set full capture {os.path.basename(rep)} ({len(launches)} launches, {clips} clips x {steps} steps each)",
       "dram_bytes_per_launch": tot, "clips_per_launch": clips, "steps_per_launch": steps,
       "dram_bytes_per_clip_step": tot / clips / steps, "algorithmic_bytes_per_clip_step": alg,
       "duration_us_under_ncu": [round(l["duration_s"] * 1e6, 1) for l in launches], "launches": launches}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "lstm_traffic.json")
json.dump(res, open(path, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "launches"}, indent=1))
