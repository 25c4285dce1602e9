#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_langevin_r1.ncu-rep profiles/r01_langevin_full.csv
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__waves_per_multiprocessor',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
    'sm__cycles_active.min', 'sm__cycles_active.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
]


def launches(src, dst):
  rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
  hdr = rows[0]
  k, v, m = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
  tot = collections.OrderedDict()
  for r in rows[1:]:
    if len(r) <= v or r[m] != 'gpu__time_duration.sum':
      continue
    name = r[k].split('(')[0]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += float(r[v].replace(',', ''))
  unit = rows[1][hdr.index('Metric Unit')]
  total = sum(t[1] for t in tot.values())
  with open(dst, 'w') as f:
    f.write(f'# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), source {src}\n\n')
    f.write('Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
    f.write(f'| kernel | launches | total ({unit}) | share |\n|---|---:|---:|---:|\n')
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
      f.write(f'| `{name}` | {n} | {t:.0f} | {100 * t / total:.2f}% |\n')
  print(open(dst).read())


def full(src, dst):
  out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  with open(dst, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + [d[hdr.index('Kernel Name')].split('(')[0] + f' #{i}' for i, d in enumerate(data)])
    for mname in METRICS:
      if mname in hdr:
        j = hdr.index(mname)
        w.writerow([mname, units[j]] + [d[j] for d in data])
  print(open(dst).read())


if __name__ == '__main__':
  {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
