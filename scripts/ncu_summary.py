#!/usr/bin/env python
"""Key metrics of an `ncu --set full` capture as a markdown table (for profiles/):
ncu_summary.py <report.ncu-rep> "<title>" > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio"]
print(f"# {title}\n\n`{rep.split('/')[-1]}` (ncu --set full --clock-control none; one launch per row)\n")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"**{name.split('(')[0]}**\n\n| metric | value | unit |\n|---|---:|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| `{w}` | {r[i]} | {units[i]} |")
    st = [(float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("\nwarp stalls per issued instruction: " + ", ".join(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for v, h in sorted(st, reverse=True)[:7]) + "\n")
