#!/usr/bin/env python3
"""Turns ncu captures under gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1a.csv  > profiles/r1a_launches.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof_trace_r1d.ncu-rep > profiles/r1d_k_trace.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: per-kernel device time (cold-cache, serialised under ncu: compare shares)")
    for k, v in tot.most_common():
        print(f"{k:50s} launches={cnt[k]:5d} total_ms={v / 1e6:10.3f} share={v / total * 100:5.1f}% avg_us={v / cnt[k] / 1e3:10.1f}")


def kernel(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        print(f"# {path}: {d.get('Kernel Name', '?')[:100]}")
        for k in KEYS:
            if k in d:
                print(f"{k:85s} {units[hdr.index(k)]:12s} {d[k]}")
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        elif hdr and r[0].isdigit() and len(r) > 2 and r[2] == "-":
            try:
                agg.append((int(r[ci]), int(r[ti]), int(r[si]), cur, int(r[0]), r[1].strip()[:90]))
            except ValueError:
                pass
    if agg:
        tot, ts = sum(a[0] for a in agg), max(sum(a[2] for a in agg), 1)
        byfile = collections.Counter()
        for a in agg:
            byfile[a[3]] += a[0]
        print("# warp instructions by source file:", {k: f"{v / tot * 100:.1f}%" for k, v in byfile.most_common()})
        print("# hottest source lines: %instructions  %stall-samples  active-threads/instruction  file:line  source")
        for a in sorted(agg, reverse=True)[:top]:
            print(f"{a[0] / tot * 100:5.1f}% {a[2] / ts * 100:5.1f}% {a[1] / max(a[0], 1):5.1f}  {a[3]}:{a[4]}  {a[5]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
