"""Dev: kernel-only timing of the Ising sweep at configs 3 and 4 (spin-updates/s)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import ising, random as nqr
from oracle import ising_ref as I, jaxlike as J
dev = torch.device('cuda')
CONFIGS = ((64, 4096, 1001), (1024, 256, 101), (128, 2048, 201), (256, 1024, 101), (2048, 148, 21), (4096, 148, 6))
if len(sys.argv) > 1:   # e.g. ising_bench.py 64,4736,1001 64,2368,1001
  CONFIGS = tuple(tuple(int(v) for v in a.split(',')) for a in sys.argv[1:])
for (L, R, T) in CONFIGS:
  times = np.linspace(0, 1, T, dtype=np.float32)
  lt = ising.log_temp_baseline(0.69, 10, 1)(times); h = ising.field_baseline(-1, 1)(times)
  s0 = -np.ones((L, L), np.float32) if L == 64 else ising.random_spins([L, L], .5, nqr.PRNGKey(1))
  seeds = nqr.split(nqr.PRNGKey(0), R)
  _, _, spins, L_, batched, seeds_t, broadcast, bits = ising._prepare(ising.IsingParameters(lt, h), s0, seeds)
  for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run = ising._run_kernel(lt, h, bits, broadcast, seeds_t, L, want_partials=True); e1.record()
    torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  print(f'L={L} R={R} T={T}: {ms:.2f} ms  {R * (T - 1) * L * L / ms / 1e-3:.4g} spin-updates/s')
