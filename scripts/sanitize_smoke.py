"""Dev: one small invocation of every hot kernel, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
the pair kernel (producer / consumer warps, named barriers), the static and the persistent (FIFO ring, __threadfence)
Langevin launches, RECORD + checkpoint + replay, the trainer kernels, warp-team / CTA-team / segment-scheduled /
global-lattice / generic Ising kernels and the gradient assembly.

    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py      (one tool per gpurun call)

The schedulers' waits are unbounded here (NQ_WAIT_POLLS=0): under a sanitizer a legitimate wait can take seconds."""
import os, sys
os.environ.setdefault('NQ_WAIT_POLLS', '0')
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, ising, parameterization as p10n, random as nqr, space, training

_, shift = space.free()
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
coeffs = np.array([20., 19.95996] + [0.] * 11, np.float32)
key = nqr.PRNGKey(0)
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  est = bc.estimate_gradient_boltzinit(200, pot, 0., 40., 3, shift, 300, 1e-8, KT, 1.0, KT / 1e7)
  g, _ = est(coeffs, np.zeros((200, 1, 1), np.float32), key)
  bc.run_brownian_opt(pot, coeffs, np.zeros((70, 1, 1), np.float32), 0., 40., 0, shift, nqr.split(key, 70), 300, 1e-8, KT, 1.0, KT / 1e7)
  ck = bc.estimate_gradient_boltzinit_checkpointed(130, pot, 0., 40., 0, shift, 256, 1e-8, KT, 1.0, KT / 1e7, checkpoint_every=64)
  ck(coeffs, np.zeros((130, 1, 1), np.float32), key)
# persistent scheduler (FIFO ring in global memory): forced on from 148 * 64 * 2 trajectories
os.environ['NQ_PERSISTENT'] = '1'
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  n_p = 148 * 64 * 2 + 37
  bc.gradient_terms(pot, coeffs, np.zeros(n_p, np.float32), nqr.split(key, n_p), 0., 40., 0, shift, 520, 1e-8, KT, 1.0, KT / 1e7)
del os.environ['NQ_PERSISTENT']
# periodic shift (the other specialisation of the step loop) and the spring potential
bc.set_math_mode('strict')
_, pshift = space.periodic(50.0)
bc.gradient_terms(pot, coeffs, np.zeros(20000, np.float32), nqr.split(key, 20000), 0., 40., 2, pshift, 130, 1e-8, KT, 1.0, KT / 1e7)
bc.gradient_terms(bc.V_spring(1.0), coeffs, np.full(300, 5., np.float32), nqr.split(key, 300), 5., 10., 0, shift, 100, 1e-3, 1.0, 1.0, 1.0)
tr = training.LangevinTrainer(96, bc.V_spring(1.0), bc.fit_linear_to_cheby(1.0, 5e-3, 5., 10., 12), 5., 10., 2, shift, 200,
                              5e-3, 1.0, 1.0, 1.0, key, eqm=dict(sim_length=20, dt=5e-3, kT=1.0, gamma=1.0), use_graph=False)
tr.run(2)
for (shape, R, T) in (((64, 64), 9, 6), ((128, 128), 5, 4), ((256, 256), 3, 4), ((10, 10), 3, 5), ((12,), 2, 4), ((8, 6, 4), 2, 4)):
  times = np.linspace(0, 1, T, dtype=np.float32)
  params = ising.IsingParameters(ising.log_temp_baseline(0.69, 10, 1)(times), ising.field_baseline(-1, 1)(times))
  s0 = ising.random_spins(list(shape), .5, nqr.PRNGKey(1))
  ising.simulate_ising(params, s0, nqr.split(key, R), return_states=True)
# segment-scheduled CTA team (R > #SMs, ticket counter + lattice hand-over through the workspace) and the
# global-lattice variant (L >= 2048: colour planes in HBM / L2)
for (Lb, Rb, Tb) in ((512, 150, 7), (2048, 2, 3)):
  times = np.linspace(0, 1, Tb, dtype=np.float32)
  params = ising.IsingParameters(ising.log_temp_baseline(0.69, 10, 1)(times), ising.field_baseline(-1, 1)(times))
  ising.simulate_ising(params, ising.random_spins([Lb, Lb], .5, nqr.PRNGKey(1)), nqr.split(key, Rb))
sched = ising.IsingSchedule(p10n.Chebyshev(np.array([0.3, -0.2, 0.1], np.float32)), p10n.Chebyshev(np.array([-0.4, 0.6, 0.2], np.float32)))
ising.estimate_gradient_rev(ising.total_entropy_production)(sched, np.linspace(0, 1, 5, dtype=np.float32),
                                                            -np.ones((64, 64), np.float32), nqr.split(key, 6))
torch.cuda.synchronize()
print('sanitize smoke done')
