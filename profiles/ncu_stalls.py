#!/usr/bin/env python
"""Warp-state (stall reason) shares, issue utilisation and instruction-cache hit rate of an ncu capture: ncu_stalls.py <rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]
for v in rows[2:]:
    d = dict(zip(h, v))
    st = {k.split("stalled_")[1]: float(d[k]) for k in h if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k}
    tot = sum(st.values()) or 1
    print(d.get("Kernel Name", "?"), "duration_ms", float(d["gpu__time_duration.sum"]) if "gpu__time_duration.sum" in d else "?")
    print("  issue_active %% of active: %s   icc hit rate %%: %s   inst/cycle/SM: %s" % (d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), d.get("sm__icc_request_hit_rate.pct"), d.get("sm__inst_executed.sum.per_cycle_active")))
    print("  " + "  ".join(f"{k} {100 * x / tot:.1f}" for k, x in sorted(st.items(), key=lambda kv: -kv[1]) if x / tot > 0.02))
