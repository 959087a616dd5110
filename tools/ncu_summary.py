"""CPU: raw-page extract of ncu reports into a markdown table (one column per captured launch).
usage: python tools/ncu_summary.py out.md "title" report1.ncu-rep [report2.ncu-rep ...]
       python tools/ncu_summary.py --by-kernel out.md "title" report.ncu-rep      (one ROW per kernel name: its last launch)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
        "sm__icc_request_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
by_kernel = "--by-kernel" in sys.argv
argv = [a for a in sys.argv[1:] if a != "--by-kernel"]
out, title, reps = argv[0], argv[1], argv[2:]
if by_kernel:
    r = subprocess.run(["ncu", "-i", reps[0], "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    last, count = {}, {}
    for v in rows[2:]:
        name = v[col["Kernel Name"]].split("(")[0].replace("void ", "")
        last[name] = v
        count[name] = count.get(name, 0) + 1
    f = lambda v, k: float(v[col[k]].replace(",", "")) if k in col and v[col[k]] not in ("", "n/a") else float("nan")
    unit = {h: rows[1][i] for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    L = [f"# {title}\n\n", f"`{reps[0].split('/')[-1]}`: {len(rows) - 2} launches captured with `ncu --set full --clock-control none`; one row per "
         "kernel (its LAST launch; `n` = launches captured). DRAM bytes = `dram__bytes_read.sum + dram__bytes_write.sum`; GB/s = "
         "those bytes / `gpu__time_duration.sum` (cold-cache, serialised replays: shares, not absolutes).\n\n",
         "| kernel | n | us | DRAM MB | GB/s | DRAM % of peak | SM throughput % | achieved occupancy % | regs | grid x block |\n",
         "|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n"]
    for name, v in last.items():
        us = f(v, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit["gpu__time_duration.sum"], 1.0)
        by = (f(v, "dram__bytes_read.sum") * scale.get(unit["dram__bytes_read.sum"], 1.0) +
              f(v, "dram__bytes_write.sum") * scale.get(unit["dram__bytes_write.sum"], 1.0))
        L.append(f"| `{name[:60]}` | {count[name]} | {us:.1f} | {by / 1e6:.1f} | {by / us / 1e3:.0f} | "
                 f"{f(v, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{f(v, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{f(v, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {v[col['launch__registers_per_thread']]} | "
                 f"{v[col['launch__grid_size']]} x {v[col['launch__block_size']]} |\n")
    open(out, "w").write("".join(L))
    print("wrote", out)
    sys.exit(0)
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
