"""Forward evaluation of the linear-in-Chebyshev trap protocol of the barrier-crossing model.

usage: run_barrier_crossing_fwd end_time dt degree kappa batch_size savestring
(reference script lines 62-150: all positions, log-probs and works are pickled).
"""
import sys

import numpy as np

from .. import barrier_crossing, random as nqrandom, space
from . import dump, to_numpy, wall_clock_key


def main(argv, savedir='output/'):
  end_time, dt = float(argv[0]), float(argv[1])
  kappa, batch_size, savestring = float(argv[3]), int(argv[4]), str(argv[5])
  N = dim = 1
  simulation_steps = int(end_time / dt)
  D = 2e6
  kT = 4.183
  beta = 1.0 / kT
  mgamma = kT / D
  r0_init, r0_final = 0, 40.
  degree = 12
  init_coeffs = barrier_crossing.fit_linear_to_cheby(end_time, dt, r0_init, r0_final, degree)
  Neq = 0
  x_m, delta_E = 14., 0
  kappa_l = kappa_r = kappa / (beta * x_m ** 2)
  k_trap = 0.4
  init_position = r0_init * np.ones((N, dim), np.float32)
  key = wall_clock_key()
  key, split1 = nqrandom.split_host(key, 2)
  _, shift_fn = space.free()
  dt_eq, eqm_sim_length = 1e-6, 50
  energy_fn = barrier_crossing.V_biomolecule_shifted(kappa_l, kappa_r, x_m, delta_E, k_trap, beta)
  eqm_fn = barrier_crossing.mapped_brownian_eqm(batch_size, energy_fn, init_position, r0_init, shift_fn,
                                                eqm_sim_length, dt_eq, kT, mass=1.0, gamma=mgamma)
  split1, splitoon, splitoonoon = nqrandom.split_host(split1, 3)
  init_pos_batch = eqm_fn(splitoon)
  seeds = nqrandom.split(splitoonoon, batch_size)
  all_pos, all_log_probs, all_works = barrier_crossing.run_brownian_opt(
      energy_fn, init_coeffs, init_pos_batch, r0_init, r0_final, Neq, shift_fn, seeds, simulation_steps, dt, kT,
      mass=1.0, gamma=mgamma)
  print('average work: ', float(all_works.sum(1).mean()), tuple(all_works.shape))
  prefix = 'end_time_%i_dt_%i_kappa_%.3f_batch_%i_' % (end_time, dt, kappa_l, batch_size)
  out = dict(all_pos=all_pos, all_works=all_works, all_log_probs=all_log_probs)
  for name, arr in out.items():
    dump(to_numpy(arr), savedir + prefix + savestring + name + '.pkl')
  return out


if __name__ == '__main__':
  main(sys.argv[1:])
