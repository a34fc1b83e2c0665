"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export: the metrics B200_PROFILING.md names, one kernel per block.
  python tools/ncu_summary.py <raw.csv> [out.txt]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed_pipe_uniform.sum",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "gpc__cycles_elapsed.max", "sm__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    out.append(f"== {d.get('Kernel Name', '?')[:110]}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
    for k in KEYS:
        if k in d and d[k] != "":
            out.append(f"  {k:88s} {d[k]:>22s} {units[hdr.index(k)]}")
    extra = [k for k in hdr if ("tma" in k.lower() or "pipe_fp64" in k) and k not in KEYS and d.get(k, "") not in ("", "0")]
    for k in extra[:12]:
        out.append(f"  {k:88s} {d[k]:>22s} {units[hdr.index(k)]}")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
