#!/usr/bin/env python3
"""Turns the ncu artefacts of a gpurun call into the tracked summaries under profiles/.
usage: ncu_summarize.py <step_full.ncu-rep> <launches.csv> <tag>      (run where ncu is installed)"""
import csv, io, json, subprocess, sys
from collections import OrderedDict

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
out = []
for r in rows[2:]:
    out.append("\n## " + r[h.index("Kernel Name")] + "\n\n| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        if m in h:
            out.append(f"| `{m}` | {r[h.index(m)]} | {u[h.index(m)]} |")
open(f"profiles/{tag}_ncu_kernels.md", "w").write(
    f"# {tag}: one resident step (8192 pairs x 10 kbp, 5 %, CIGAR), every kernel, `ncu --set full --clock-control none`\n" + "\n".join(out) + "\n")

lr = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
lh = lr[0]
ki, vi = lh.index("Kernel Name"), lh.index("Metric Value")
agg = OrderedDict()
for r in lr[1:]:
    k = r[ki].split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
lines = ["# Launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` under ncu "
         "(gpu__time_duration.sum, --clock-control none; resident steps of 8192 pairs + e2e chunks of 4096)\n",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.2f} % |")
open(f"profiles/{tag}_launches_bench.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
