"""Dev: how the hardware distributes the warps of a single-wave launch over SMs and SM sub-partitions."""
import os, sys, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import _devlib, _lib
lib = _lib.lib()
for blocks, threads in ((1563, 64), (1776, 64), (782, 128), (888, 128)):
  out = torch.zeros(blocks, dtype=torch.int64, device='cuda')
  sink = torch.zeros(4, dtype=torch.int32, device='cuda')
  _lib.check(_devlib.lib().nqdev_microbench(12, blocks, threads, 200, _lib.ptr(out), _lib.ptr(sink), None))
  torch.cuda.synchronize()
  v = out.cpu().numpy()
  per_sm = collections.Counter(); per_smsp = collections.Counter()
  for x in v:
    sm = int(x & 0xffff)
    per_sm[sm] += threads // 32
    for w in range(threads // 32):
      per_smsp[(sm, (int(x >> (16 + 8 * w)) & 0xff) % 4)] += 1
  sm_counts = collections.Counter(per_sm.values()); smsp_counts = collections.Counter(per_smsp.values())
  print(f'{blocks} CTAs x {threads} threads: SMs used {len(per_sm)}, warps/SM histogram {dict(sorted(sm_counts.items()))}, '
        f'warps/SMSP histogram {dict(sorted(smsp_counts.items()))} (SMSPs used {len(per_smsp)})')
