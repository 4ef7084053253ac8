"""Summarise one `ncu --set full` capture of k_fb2 for profiles/: python scripts/ncu_summary.py rep.ncu-rep TAG CELLS READS
Writes profiles/TAG_k_fb2_ncu_metrics.txt (selected raw metrics) and rewrites profiles/ncu_traffic.json (what bench.py
reads for `roofline.traffic` and `thread_inst_per_cell`)."""
import csv, json, os, subprocess, sys
rep, tag, cells, reads = sys.argv[1], sys.argv[2], int(float(sys.argv[3])), int(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, vals = rows[0], rows[1], rows[2]
m = {n: (units[i], vals[i]) for i, n in enumerate(h)}
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.max.pct_of_peak_sustained_active", "smsp__issue_active.min.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
want += sorted(n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"))
def f(name):
    u, v = m[name]
    v = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    return v * scale
kname = vals[h.index("Kernel Name")] if "Kernel Name" in h else "k_fb2"
lines = ["# ncu --set full --clock-control none --import-source on -k regex:k_fb2 -s 1 -c 1 python scripts/tune.py %d \"\" (shipped libphmm_sm100.so)" % reads,
         "# kernel: %s; %d reads = one region per resident block; cells %.4e" % (kname, reads, cells)]
for n in want:
    if n in m:
        lines.append("%-90s %-16s %s" % (n, m[n][0], m[n][1]))
wi, tpw = f("smsp__inst_executed.sum"), f("smsp__thread_inst_executed_per_inst_executed.ratio")
wf = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in m else float("nan")
lines.append("thread-instructions per cell: %.0f x %.2f / %d = %.1f" % (wi, tpw, cells, wi * tpw / cells))
lines.append("shared-memory wavefronts per cell: %.2f  (one per cycle per SM: ceiling %.1f Gcell/s at 1965 MHz)" % (wf / cells, 148 * 1.965 / (wf / cells)))
open(os.path.join(root, "profiles", tag + "_k_fb2_ncu_metrics.txt"), "w").write("\n".join(lines) + "\n")
rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
json.dump({"kernel": "%s (%s)" % (kname, tag), "dram_bytes_read": rd, "dram_bytes_write": wr, "cells": cells,
           "bytes_per_cell": (rd + wr) / cells, "gpu_time_duration": "%s %s" % (m["gpu__time_duration.sum"][1], m["gpu__time_duration.sum"][0]),
           "note": "one region per resident block (740 = 148 SMs x 5); bytes and instructions scale with cells",
           "capture": lines[0][2:] + ", profiles/%s_k_fb2_ncu_metrics.txt" % tag, "reads": reads, "warp_inst_executed": wi,
           "thread_inst_per_warp_inst": tpw, "thread_inst_per_cell": wi * tpw / cells, "shared_wavefronts_per_cell": wf / cells},
          open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("\n".join(lines[-2:]))
