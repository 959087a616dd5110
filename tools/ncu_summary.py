"""CPU: raw-page extract of ncu reports into a markdown table (one column per captured launch).
usage: python tools/ncu_summary.py out.md "title" report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
        "sm__icc_request_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
lines = [f"# {title}\n\n"]
for rep in reps:
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        lines.append(f"## {rep}: no launches captured\n\n")
        continue
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    lines.append(f"## `{rep.split('/')[-1]}`\n\n| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |\n|---|---|" + "---:|" * (len(rows) - 2) + "\n")
    lines.append("| kernel | | " + " | ".join(v[ki].split("(")[0][-46:] for v in rows[2:]) + " |\n")
    for i, h in enumerate(hdr):
        if h in WANT:
            lines.append(f"| {h} | {units[i]} | " + " | ".join(v[i] for v in rows[2:]) + " |\n")
    lines.append("\n")
open(out, "w").write("".join(lines))
print("wrote", out)
