#!/usr/bin/env python
"""Dev helper: static opcode count of the innermost step loop of a Langevin kernel, from `cuobjdump -sass`.

  python scripts/hot_loop_static.py <lib.so> [kernel-name substring]   (default: the STRICT persistent kernel, D = 13)

The loop is found as the backward branch whose body holds the most FFMA; the rarely taken patch blocks that sit inside
the address range are counted too (a fixed ~60 instructions), so compare builds with each other, not with ncu.
"""
import collections, re, subprocess, sys
lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else 'langevin_persistent_kernelILi2ELb0ELi13E'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ALU = ('LOP3', 'SHF', 'IADD3', 'VIADD', 'ISETP', 'LEA', 'I2FP', 'FSETP', 'PRMT', 'POPC', 'MOV', 'SEL', 'PLOP3', 'F2F', 'F2I', 'I2F', 'FMNMX', 'VIMNMX')
def report(name, lo, hi, b):
  h = collections.Counter(o if o.startswith(('IMAD', 'MUFU', 'LDS')) else o.split('.')[0] for _, o in b)
  alu = sum(v for k, v in h.items() if k in ALU)
  print(f'{name}\n  loop {lo:#x}..{hi:#x}: {len(b)} instructions, {alu} on the ALU pipe')
  print('  ' + ', '.join(f'{k} {v}' for k, v in h.most_common()))


cur, body = None, []
kernels = {}
for line in out.splitlines():
  m = re.match(r'\s*Function : (\S+)', line)
  if m:
    cur = m.group(1); kernels[cur] = []; continue
  m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);', line)
  if m and cur: kernels[cur].append((int(m.group(1), 16), m.group(3), m.group(4)))
for name, ins in kernels.items():
  if want not in name: continue
  loops = []
  for a, op, rest in ins:
    if op.startswith('BRA'):
      t = re.search(r'0x([0-9a-f]+)', rest)
      if t and int(t.group(1), 16) < a:
        lo = int(t.group(1), 16)
        b = [(x, o) for x, o, _ in ins if lo <= x <= a]
        if sum(o.startswith('FFMA') for _, o in b) >= 20: loops.append((len(b), lo, a, b))
  loops.sort()
  loops = [l for l in loops if l[0] < 2 * loops[0][0]]   # the step loops (free / periodic shift), not their parents
  for n_, lo, hi, b in loops:
    report(name, lo, hi, b)
