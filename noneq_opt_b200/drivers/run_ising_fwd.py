"""Forward evaluation of a stored Ising protocol (weights pickled by run_ising_opt).

usage: run_ising_fwd size seed time_steps samples prot_size prot_seed prot_time_steps prot_degree \
                     prot_batch_size prot_training_steps
(experiments/ising/run_ising_fwd.py; the hard-coded cluster paths of the reference are replaced by
`loaddir` / `savedir`, the near-equilibrium comparison protocol by an optional CSV `theory_csv`).
"""
import csv
import pickle
import sys

import numpy as np

from .. import ising, parameterization as p10n, random as nqrandom
from . import dump, to_numpy, wall_clock_key


def entropy_with_endpoint_correction(parameters, initial_spins, summary):
  state0 = ising.IsingState(initial_spins, ising.map_slice(parameters, 0))
  log_p_init = float(-ising.energy(state0)) / float(np.exp(parameters.log_temp[0]))
  final_kT = float(np.exp(parameters.log_temp[-1]))
  log_p_final = -summary.energy[:, -1] / final_kT
  ent = np.array(summary.entropy_production)
  ent[:, -1] = ent[:, -1] + (log_p_init - log_p_final)
  return ent


def main(argv, loaddir='output/', savedir='output/', theory_csv=None):
  size, time_steps, samples = int(argv[0]), int(argv[2]), int(argv[3])
  prot_size, prot_time_steps, prot_degree = int(argv[4]), int(argv[6]), int(argv[7])
  prot_batch_size, prot_training_steps = int(argv[8]), int(argv[9])
  prefix = '%ix%ilattice_timesteps_%i_chebdeg_%i_batch_%i_trainsteps_%i_' % (
      prot_size, prot_size, prot_time_steps, prot_degree, prot_batch_size, prot_training_steps)
  with open(loaddir + prefix + 'log_temp_weights.pkl', 'rb') as f:
    temp_weights = pickle.load(f)[-1]
  with open(loaddir + prefix + 'field_weights.pkl', 'rb') as f:
    field_weights = pickle.load(f)[-1]
  Tbaseline = ising.log_temp_baseline(min_temp=0.69, degree=2. if prot_time_steps == 11 else 1.)
  Fbaseline = ising.field_baseline()
  schedules = {'AD': ising.IsingSchedule(
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(temp_weights), y0=0., y1=0.), baseline=Tbaseline),
      p10n.AddBaseline(p10n.ConstrainEndpoints(p10n.Chebyshev(field_weights), y0=0., y1=0.), baseline=Fbaseline))}
  if theory_csv:
    rows = [(float(r[0]), float(r[1])) for r in csv.reader(open(theory_csv))]
    temps, fields = np.array([r[0] for r in rows]), np.array([r[1] for r in rows])
    schedules['th_2015'] = ising.IsingSchedule(
        p10n.PiecewiseLinear(np.log(temps[1:-2]), y0=np.log(temps[0]), y1=np.log(temps[-1])),
        p10n.PiecewiseLinear(fields[1:-2], y0=fields[0], y1=fields[-1]))
  times = np.linspace(0, 1, time_steps, dtype=np.float32)
  initial_spins = -np.ones([size, size], np.float32)
  out = {}
  for name, schedule in schedules.items():
    parameters = schedule(times)
    seeds = nqrandom.split(wall_clock_key(), samples)
    final_state, summary = ising.simulate_ising(parameters, initial_spins, seeds)
    summary = to_numpy(summary)
    ent = entropy_with_endpoint_correction(parameters, initial_spins, summary)
    out[name] = dict(summary=summary, entropies=ent, mean_entropy=float(ent.sum(1).mean()),
                     sem_entropy=float(ent.sum(1).std() / np.sqrt(samples)))
    print(name, 'total entropy production: %.4f +- %.4f' % (out[name]['mean_entropy'], out[name]['sem_entropy']))
  dump(out, savedir + prefix + 'fwd_%ix%i_timesteps_%i_samples_%i.pkl' % (size, size, time_steps, samples))
  return out


if __name__ == '__main__':
  main(sys.argv[1:])
