#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  profiles/r1_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  profiles/r1_ncu_full.md [profiles/traffic.json]
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki].split("(")[0], []).append(v * scale)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("source: `%s`\n\n| kernel | launches | mean us | share of listed time |\n|---|---|---|---|\n" % src)
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print(open(dst).read())


def full(src, dst, traffic_json=None):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    traffic = {}
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (one launch per row; per-launch values)\n\nsource: `%s`\n\n" % src)
        for r in data:
            name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name)
            vals = {}
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    vals[m] = (r[i], units[i])
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            f.write("\n")
            try:
                def to_bytes(m):
                    v, u = vals[m]
                    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                tb = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                traffic.setdefault(name, []).append(tb)           # one entry per launch of the step
            except Exception:
                pass
    if traffic_json:
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
