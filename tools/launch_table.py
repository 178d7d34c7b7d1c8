#!/usr/bin/env python3
"""Per-kernel table (time share, issue utilisation, threads/instruction, instruction-fetch stalls, registers) from an
ncu --csv launch list taken with the metrics named in profiles/README.md."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0].replace('void ', '').replace('psyd::', '')
    v = float(row['Metric Value'].replace(',', ''))
    if row['Metric Name'] == 'gpu__time_duration.sum':
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1}[row['Metric Unit']]
    agg[name][row['Metric Name']].append(v)
tot = sum(sum(m['gpu__time_duration.sum']) for m in agg.values())
print(f"# {sys.argv[1]}: serialised, cold-cache launches under ncu: compare shares, not absolute times")
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1]['gpu__time_duration.sum'])):
    t = m['gpu__time_duration.sum']
    w = [x / sum(t) for x in t]
    avg = lambda key: sum(a * b for a, b in zip(m[key], w)) if key in m else float('nan')
    print(f"{k[:40]:40s} n={len(t):4d} ms={sum(t):8.2f} share={sum(t) / tot * 100:5.1f}% "
          f"issue={avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}% "
          f"thr/inst={avg('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} "
          f"no_inst={avg('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'):5.2f} "
          f"regs={m['launch__registers_per_thread'][0] if 'launch__registers_per_thread' in m else 0:.0f}")
