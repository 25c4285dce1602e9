#!/usr/bin/env python
"""Histogram SASS opcodes between two addresses of a cuobjdump -sass dump (dev helper)."""
import re, sys, collections
path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
skip = [tuple(int(v, 16) for v in s.split('-')) for s in sys.argv[4:]]
h = collections.Counter()
for line in open(path):
  m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
  if not m: continue
  a = int(m.group(1), 16)
  if a < lo or a > hi or any(s <= a <= e for s, e in skip): continue
  op = m.group(3)
  h[op.split('.')[0] if not op.startswith(('IMAD', 'LDS', 'MUFU')) else op] += 1
tot = sum(h.values())
for k, v in h.most_common(): print(f'{k:16s} {v}')
print('total', tot)
