"""Dev: one launch of each headline Ising kernel for ncu (config 3 with and without warp-team segments; config 4
through the CTA-team segment scheduler).  usage: ising_profile.py [c3|c3_noseg|c4]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
if which == 'c3_noseg':
  os.environ['NQ_ISING_NO_SEGMENTS'] = '1'
from noneq_opt_b200 import ising, parameterization as p10n, random as nqr

def schedule(w):
  return ising.IsingSchedule(
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w[0]), 0., 0.), ising.log_temp_baseline(0.69, 10, 1)),
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w[1]), 0., 0.), ising.field_baseline(-1, 1)))

if which.startswith('c3'):
  L, R, T = 64, 4096, 1001
  w = (0.1 * np.random.RandomState(0).normal(size=(2, 16))).astype(np.float32)
  params = schedule(w)(np.linspace(0, 1, T, dtype=np.float32))
  spins0 = -torch.ones((L, L), dtype=torch.float32, device='cuda')
else:
  L, R, T = 1024, 256, 201
  params = schedule(np.zeros((2, 16), np.float32))(np.linspace(0, 1, 10001, dtype=np.float32)[:T])
  spins0 = ising.random_spins([L, L], 0.5, nqr.PRNGKey(1))
seeds = nqr.split(nqr.PRNGKey(0), R)
log_temp, field, spins, L_, batched, seeds_t, broadcast, bits = ising._prepare(params, spins0, seeds)
for _ in range(2):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  ising._run_kernel(log_temp, field, bits, broadcast, seeds_t, L, want_partials=which.startswith('c3'))
  e1.record(); torch.cuda.synchronize()
print(which, 'ms', e0.elapsed_time(e1), 'updates/s %.4g' % (R * (T - 1) * L * L / (e0.elapsed_time(e1) * 1e-3)))
