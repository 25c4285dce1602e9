"""Optimise the (log-temperature, field) protocol of the 2-D Ising model for minimum entropy production.

usage: run_ising_opt size seed time_steps degree batch_size training_steps save_every
(the argument list, schedule construction and output files of run_ising_opt/run_ising_opt.py).
"""
import sys

import numpy as np

from .. import ising, optimizers, parameterization as p10n
from . import dump, to_numpy


def main(argv, savedir='output/', quiet=False):
  size, seed, time_steps = int(argv[0]), int(argv[1]), int(argv[2])
  field_degree = log_temp_degree = int(argv[3])
  batch_size, training_steps, save_every = int(argv[4]), int(argv[5]), int(argv[6])
  optimizer = optimizers.adam(optimizers.exponential_decay(0.1, training_steps, .01))
  max_grad_ = 1
  Tbaseline = ising.log_temp_baseline(min_temp=0.65)
  Fbaseline = ising.field_baseline()
  initial_log_temp_schedule = p10n.AddBaseline(
      p10n.ConstrainEndpoints(p10n.Chebyshev(np.zeros(log_temp_degree, np.float32)), y0=0., y1=0.), baseline=Tbaseline)
  initial_field_schedule = p10n.AddBaseline(
      p10n.ConstrainEndpoints(p10n.Chebyshev(np.zeros(field_degree, np.float32)), y0=0., y1=0.), baseline=Fbaseline)
  assert initial_field_schedule.domain == (0., 1.)
  assert initial_log_temp_schedule.domain == (0., 1.)
  initial_schedule = ising.IsingSchedule(initial_log_temp_schedule, initial_field_schedule)
  stream = ising.seed_stream(0)
  state = optimizer.init_fn(initial_schedule)
  initial_spins = -np.ones([size, size], np.float32)
  train_step = ising.get_train_step(optimizer, initial_spins, batch_size, time_steps, ising.total_entropy_production,
                                    mode='fwd', max_grad=max_grad_)
  summaries, entropies, log_temp_weights, fields_weights = [], [], [], []
  times = np.linspace(0, 1, time_steps, dtype=np.float32)
  params0 = initial_schedule(times)
  initial_spin_state = ising.IsingState(initial_spins, ising.map_slice(params0, 0))
  log_p_init = float(-ising.energy(initial_spin_state)) / float(np.exp(params0.log_temp[0]))
  final_kT = float(np.exp(params0.log_temp[-1]))
  for j in range(training_steps):
    state, summary = train_step(state, j, next(stream))
    summary = to_numpy(summary)
    log_p_final = -summary.energy[:, -1] / final_kT          # endpoint correction, lines 103-108
    final_entropies = np.array(summary.entropy_production)
    final_entropies[:, -1] = final_entropies[:, -1] + (log_p_init - log_p_final)
    entropies.append(final_entropies)
    if (j % save_every == 0) or (j == training_steps - 1):
      if not quiet:
        print('saving iteration ', j, ' mean entropy production ', float(final_entropies.sum(1).mean()))
      summaries.append(summary)
      sched = optimizer.params_fn(state)
      log_temp_weights.append(np.array(sched.log_temp.variables['wrapped'].variables['wrapped'].weights))
      fields_weights.append(np.array(sched.field.variables['wrapped'].variables['wrapped'].weights))
  prefix = '%ix%ilattice_timesteps_%i_chebdeg_%i_batch_%i_trainsteps_%i_' % (
      size, size, time_steps, field_degree, batch_size, training_steps)
  dump(summaries, savedir + prefix + 'summaries.pkl')
  dump(entropies, savedir + prefix + 'entropies.pkl')
  dump(log_temp_weights, savedir + prefix + 'log_temp_weights.pkl')
  dump(fields_weights, savedir + prefix + 'field_weights.pkl')
  return dict(entropies=entropies, log_temp_weights=log_temp_weights, field_weights=fields_weights)


if __name__ == '__main__':
  main(sys.argv[1:])
