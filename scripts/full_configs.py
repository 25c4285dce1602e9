"""Dev: every BASELINE.json config at its FULL size on one GPU, with a size-independent check each.
Writes gpurun_out/full_configs.json."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, ising, parameterization as p10n, random as nqr, space

res = {}
_, shift = space.free()
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)


def timed(fn, reps=3):
  out = None
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1), out


# C1: 1k x 1k spring, Chebyshev protocol + mean-work gradient; linear protocol => <W> -> 25/e
c1 = bc.fit_linear_to_cheby(1.0, 1e-3, 5., 10., 12)
seeds = nqr.split(nqr.PRNGKey(0), 1000)
ms, g = timed(lambda: bc.gradient_terms(bc.V_spring(1.0), c1, np.full(1000, 5., np.float32), seeds, 5., 10., 100, shift, 1000,
                                        1e-3, 1.0, 1.0, 1.0))
res['c1'] = dict(ms=ms, traj_steps_per_s=1000 * 1100 / ms * 1e3, mean_work=float(g['moments'][1] / g['moments'][0]),
                 expect_mean_work_about=25 / np.e)

# C2 / C5: double well, forward + gradient; mean work must agree between ensemble sizes (same physics)
pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
c2 = np.array([20., 19.95996] + [0.] * 11, np.float32)
for name, B in (('c2', 100_000), ('c5_one_gpu', 10_000_000), ('c5_shard_of_8', 1_250_000)):
  seeds = nqr.split(nqr.PRNGKey(0), B)
  x0 = torch.zeros(B, device='cuda')
  ms, g = timed(lambda: bc.gradient_terms(pot, c2, x0, seeds, 0., 40., 0, shift, 10_000, 1e-8, KT, 1.0, KT / 1e7), reps=4 if B == 100_000 else 3)
  m = g['moments'].cpu().numpy()
  res[name] = dict(ms=ms, traj_steps_per_s=B * 1e4 / ms * 1e3, mean_work=m[1] / m[0],
                   work_std_err=float(np.sqrt((m[2] / m[0] - (m[1] / m[0]) ** 2) / m[0])),
                   grad0=float((g['grad_sums'][0, 0] + g['grad_sums'][1, 0]) / B))
  del seeds, x0, g

# C3: 64^2 x 4096 replicas x 1000 sweeps with the REINFORCE gradient; heat = T * entropy production at fixed parameters
w = (0.1 * np.random.RandomState(0).normal(size=(2, 16))).astype(np.float32)
sched = ising.IsingSchedule(
    p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w[0]), 0., 0.), ising.log_temp_baseline(0.69, 10, 1)),
    p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(w[1]), 0., 0.), ising.field_baseline(-1, 1)))
times = np.linspace(0, 1, 1001, dtype=np.float32)
seeds = nqr.split(nqr.PRNGKey(0), 4096)
est = ising.estimate_gradient_rev(ising.total_entropy_production)
ms, (grads, summary) = timed(lambda: est(sched, times, -np.ones((64, 64), np.float32), seeds))
res['c3'] = dict(ms=ms, spin_updates_per_s=4096 * 1000 * 64 * 64 / ms * 1e3,
                 mean_total_entropy_production=float(summary.entropy_production.double().sum(1).mean()),
                 grad_field_w0=float(grads.field.wrapped.wrapped.weights.double().mean(0)[0]),
                 final_magnetization=float(summary.magnetization[:, -1].double().mean()))

# C4: 1024^2 x 256 replicas x 10 000 sweeps, forward + summary; running energy == energy measured from the final lattice
T = 10_001
times = np.linspace(0, 1, T, dtype=np.float32)
lt, h = ising.log_temp_baseline(0.69, 10, 1)(times), ising.field_baseline(-1, 1)(times)
s0 = ising.random_spins([1024, 1024], .5, nqr.PRNGKey(1))
seeds = nqr.split(nqr.PRNGKey(0), 256)
t0 = time.perf_counter()
ms, (final, summary) = timed(lambda: ising.simulate_ising(ising.IsingParameters(lt, h), s0, seeds), reps=1)
spins = final.spins[:8].to(torch.float64)
nbr = sum(torch.roll(spins, sh, ax) for ax in (1, 2) for sh in (-1, 1))
e_meas = -(spins * (nbr / 2 + float(h[-1]))).sum((1, 2))
res['c4'] = dict(ms=ms, spin_updates_per_s=256 * (T - 1) * 1024 * 1024 / ms * 1e3,
                 energy_rel_err_after_1e4_sweeps=float(((summary.energy[:8, -1].double() - e_meas).abs() / e_meas.abs()).max()),
                 final_magnetization=float(summary.magnetization[:, -1].double().mean()))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/full_configs.json', 'w'), indent=1)
print(json.dumps(res, indent=1))
