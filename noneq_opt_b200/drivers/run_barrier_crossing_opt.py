"""Optimise the trap protocol of the barrier-crossing model with Adam on REINFORCE gradients.

usage: run_barrier_crossing_opt end_time dt degree kappa batch_size opt_steps
(the argument list, constants and output files of the reference script, lines 66-159).

`device_loop=True` (or NQ_DEVICE_LOOP=1) runs the same loop through `training.LangevinTrainer`:
every step is one CUDA-graph replay with the key chain, the protocol table, the gradient sums and
the Adam state resident on the device; the pickles have the same structure.
"""
import os
import sys

import numpy as np

from .. import barrier_crossing, optimizers, random as nqrandom, space, training
from . import dump, to_numpy, wall_clock_key

INIT_COEFFS = np.array([2.00000000e+01, 1.99599600e+01, -2.63245988e-15, 8.41484236e-16, 0.00000000e+00,
                        -7.47209997e-15, -1.69086360e-14, 1.27892999e-14, 3.02606960e-15, -6.05805446e-15,
                        -6.36600761e-16, -5.53701086e-15, -1.14938537e-14], np.float32)


def main(argv, savedir='output/', quiet=False, device_loop=None, key=None):
  end_time, dt = float(argv[0]), float(argv[1])
  degree, kappa = int(argv[2]), float(argv[3])
  batch_size, opt_steps = int(argv[4]), int(argv[5])
  N = dim = 1
  simulation_steps = int(end_time / dt)
  D = 1e7           # nm^2/s
  kT = 4.183        # pN nm at 303 K
  beta = 1.0 / kT
  mgamma = kT / D
  r0_init, r0_final = 0, 40.
  init_coeffs = INIT_COEFFS
  Neq = 0
  x_m, delta_E = 14., 0
  kappa_l = kappa_r = kappa / (beta * x_m ** 2)
  k_trap = 1.5
  init_position = r0_init * np.ones((N, dim), np.float32)
  _, shift_fn = space.free()
  dt_eq, eqm_sim_length = 1e-8, 1000

  energy_fn = barrier_crossing.V_biomolecule_shifted(kappa_l, kappa_r, x_m, delta_E, k_trap, beta)
  eqm_fn = barrier_crossing.mapped_brownian_eqm(batch_size, energy_fn, init_position, r0_init, shift_fn,
                                                eqm_sim_length, dt_eq, kT, mass=1.0, gamma=mgamma)
  optimizer = optimizers.adam(optimizers.exponential_decay(0.1, opt_steps, 0.001))
  opt_state = optimizer.init_fn(init_coeffs)
  grad_fxn = barrier_crossing.estimate_gradient_boltzinit(batch_size, energy_fn, r0_init, r0_final, Neq, shift_fn,
                                                          simulation_steps, dt, kT, 1.0, mgamma, record_positions=False)
  coeffs_ = [(0,) + (optimizer.params_fn(opt_state),)]
  all_works = []
  key = wall_clock_key() if key is None else key
  if device_loop is None:
    device_loop = os.environ.get('NQ_DEVICE_LOOP', '0') == '1'
  if device_loop:
    # mapped_brownian_eqm ignores the dt/kT/mass/gamma it is given (barrier_crossing.py:250)
    trainer = training.LangevinTrainer(batch_size, energy_fn, init_coeffs, r0_init, r0_final, Neq, shift_fn,
                                       simulation_steps, dt, kT, 1.0, mgamma, key, init_position=r0_init,
                                       eqm=dict(sim_length=eqm_sim_length, dt=1e-6, kT=1e-5, gamma=1.0),
                                       step_size=0.1, decay_steps=opt_steps, decay_rate=0.001)
    out = trainer.run(opt_steps, record_works=True)
    hist, works = to_numpy(out['coeffs']), to_numpy(out['works'])
    all_works = [works[j] for j in range(opt_steps)]
    coeffs_ += [((j + 1),) + (hist[j + 1],) for j in range(0, opt_steps, 100)]
    if not quiet:
      for j in range(opt_steps):
        print(f'step {j}: mean work {float(all_works[j].mean()):.4f}')
    opt_steps_run = 0
  else:
    opt_steps_run = opt_steps
  for j in range(opt_steps_run):
    key, splitoon, splitoonoon = nqrandom.split_host(key, 3)
    init_pos_batch = eqm_fn(splitoon)
    grad, (_, summary) = grad_fxn(optimizer.params_fn(opt_state), init_pos_batch, splitoonoon)
    opt_state = optimizer.update_fn(j, grad, opt_state)
    all_works.append(to_numpy(summary[2]))
    if j % 100 == 0:
      coeffs_.append(((j + 1),) + (optimizer.params_fn(opt_state),))
    if not quiet:
      print(f'step {j}: mean work {float(all_works[-1].mean()):.4f}')
  coeffs_.append((opt_steps,) + ((hist[opt_steps] if device_loop else optimizer.params_fn(opt_state)),))

  prefix = 'end_time_%.10f_dt_%.10f_chebdeg_%i_kappa_%.4f_batch_%i_trainsteps_%i_' % (
      end_time, dt, degree, kappa, batch_size, opt_steps)
  dump(all_works, savedir + prefix + 'all_works.pkl')
  dump(coeffs_, savedir + prefix + 'coeffs.pkl')
  return coeffs_, all_works


if __name__ == '__main__':
  main(sys.argv[1:])
