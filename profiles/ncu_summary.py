#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> {kernel: {selected metrics}} as JSON (first instance of every kernel name)."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
ix = {h: i for i, h in enumerate(hdr)}
units = rows[1]
out = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    if name in out:
        continue
    out[name] = {m: (r[ix[m]] + " " + units[ix[m]]).strip() for m in want if m in ix}
json.dump(out, sys.stdout, indent=1)
