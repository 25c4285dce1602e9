"""Dev: where does the host-side time of the Ising gradient call go (config 3)?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import ising, parameterization as p10n, random as nqr
L, R, T = 64, 4096, 1001
w = (0.1 * np.random.RandomState(0).normal(size=(2, 16))).astype(np.float32)
sched = ising.IsingSchedule(
    p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w[0]), 0., 0.), ising.log_temp_baseline(0.69, 10, 1)),
    p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w[1]), 0., 0.), ising.field_baseline(-1, 1)))
times = np.linspace(0, 1, T, dtype=np.float32)
spins0 = -torch.ones((L, L), dtype=torch.float32, device='cuda')
seeds = nqr.split(nqr.PRNGKey(0), R)
est = ising.estimate_gradient_rev(ising.total_entropy_production)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
  t0 = sync(); est(sched, times, spins0, seeds); t1 = sync()
print('total ms', 1e3 * (t1 - t0))
t0 = sync(); tabs = ising._schedule_tables(sched, times); t1 = sync(); print('schedule tables', 1e3 * (t1 - t0))
_, lt, h, Jl, Jh = tabs
t0 = sync(); thr, lut = ising.class_tables(lt, h); t1 = sync(); print('class tables', 1e3 * (t1 - t0))
params = ising.IsingParameters(lt, h)
t0 = sync(); prep = ising._prepare(params, spins0, seeds); t1 = sync(); print('prepare', 1e3 * (t1 - t0))
_, _, spins, L_, batched, seeds_t, broadcast, bits = prep
t0 = sync(); run = ising._run_kernel(lt, h, bits, broadcast, seeds_t, L, want_partials=True); t1 = sync(); print('run kernel (incl tables)', 1e3 * (t1 - t0))
t0 = sync(); out = ising._loss_and_sensitivities('entropy', run, lt, h, L); t1 = sync(); print('loss+sens', 1e3 * (t1 - t0))
