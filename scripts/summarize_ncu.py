#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_langevin_r1.ncu-rep profiles/r01_langevin_full.csv
  python scripts/summarize_ncu.py opcodes gpurun_out/prof.ncu-rep profiles/r02_x_opcodes.md <units per launch> <unit name>
      per-opcode EXECUTED thread-instructions per unit of work (trajectory-step / spin-update), by pipe and by role,
      from the source page of a capture taken with --import-source on
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


ALU_OPS = {'SHF', 'LOP3', 'IADD3', 'VIADD', 'ISETP', 'FSETP', 'LEA', 'I2FP', 'FSEL', 'FMNMX', 'PRMT', 'PLOP3', 'SEL', 'VIMNMX',
           'POPC', 'FLO', 'BREV', 'IABS', 'FCHK', 'I2F', 'F2I', 'F2F', 'FRND', 'MOV', 'SHL', 'SHR', 'P2R', 'R2P', 'BMSK', 'SGXT',
           'IMNMX', 'VABSDIFF', 'LOP', 'VOTE', 'REDUX'}
FMA_OPS = {'FFMA', 'FMUL', 'FADD', 'IMAD', 'HFMA2', 'DFMA', 'DADD', 'DMUL', 'FFMA2', 'FMUL2', 'FADD2'}   # *2: sm_100a packed f32x2


def _role(op):
  base = op.split('.')[0]
  if base in ('SHF', 'LOP3', 'IADD3', 'VIADD', 'PRMT', 'POPC', 'LEA') or op.startswith(('IMAD.IADD', 'IMAD.MOV', 'IMAD.SHL', 'IMAD.U32', 'IMAD.WIDE', 'IMAD.X', 'IMAD.HI')) or base == 'IMAD':
    return 'integer (threefry hash, bit-sliced lattice, addresses)'
  if base == 'MUFU':
    return 'transcendental (MUFU)'
  if base in ('FFMA', 'FMUL', 'FADD', 'FFMA2', 'FMUL2', 'FADD2', 'FSETP', 'FSEL', 'FMNMX', 'I2FP', 'F2F', 'FCHK', 'HFMA2', 'FRND', 'DFMA', 'DADD', 'DMUL', 'F2I', 'I2F'):
    return 'floating point'
  if base in ('LDS', 'STS', 'LDG', 'STG', 'LDC', 'LDCU', 'LDL', 'STL', 'ATOMS', 'ATOMG', 'RED', 'ATOM', 'LDSM', 'MEMBAR', 'ERRBAR', 'CGAERRBAR', 'CCTL'):
    return 'memory'
  return 'control / loop / uniform'


def opcodes(src, dst, units, unit_name='unit'):
  import re
  units = float(units)
  out = subprocess.run(['ncu', '-i', src, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  kernel = rows[0][1] if rows and len(rows[0]) > 1 else '?'
  hdr = rows[1]
  ix = {h: i for i, h in enumerate(hdr)}
  tot, roles = collections.Counter(), collections.Counter()
  for r in rows[2:]:
    if len(r) < len(hdr):
      continue
    m = re.match(r'(@!?U?P\d\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
    if not m:
      continue
    op = m.group(2)
    n = int(r[ix['Instructions Executed']] or 0) * 32 / units
    key = op.split('.')[0] if not op.startswith(('IMAD', 'LDS', 'MUFU', 'BAR')) else '.'.join(op.split('.')[:2])
    tot[key] += n
    roles[_role(op)] += n
  total = sum(tot.values())
  alu = sum(v for k, v in tot.items() if k.split('.')[0] in ALU_OPS)
  with open(dst, 'w') as f:
    f.write(f'# executed thread-instructions per {unit_name}, by SASS opcode\n\n`{kernel}`\n\nsource: `{src}` '
            f'(ncu --set full --import-source on, source page, "Instructions Executed" x 32 / {units:.6g} {unit_name}s per launch)\n\n')
    f.write(f'**total {total:.1f}**, of which on the half-rate ALU pipe **{alu:.1f}**\n\n| role | instr / {unit_name} |\n|---|---:|\n')
    for k, v in roles.most_common():
      f.write(f'| {k} | {v:.1f} |\n')
    f.write(f'\n| opcode | instr / {unit_name} | pipe |\n|---|---:|---|\n')
    for k, v in tot.most_common():
      if v < 0.005:
        continue
      base = k.split('.')[0]
      pipe = 'ALU' if base in ALU_OPS else ('FMA' if base in FMA_OPS else ('XU' if base == 'MUFU' else ''))
      f.write(f'| {k} | {v:.2f} | {pipe} |\n')
  print(open(dst).read())
  return total, alu, kernel


if __name__ == '__main__':
  {'launches': launches, 'full': full, 'opcodes': opcodes}[sys.argv[1]](*sys.argv[2:])
