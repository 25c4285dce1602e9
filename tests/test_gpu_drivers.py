"""GPU: the command-line drivers run end to end with the reference's positional
arguments and write the reference's pickle files."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_barrier_crossing_drivers(cuda, tmp_path, monkeypatch):
  from noneq_opt_b200.drivers import run_barrier_crossing_fwd, run_barrier_crossing_opt
  monkeypatch.setenv('NQ_SEED', '7')
  out = str(tmp_path) + '/'
  coeffs, works = run_barrier_crossing_opt.main(['2e-6', '1e-8', '12', '21.3863', '512', '4'], savedir=out, quiet=True)
  assert len(works) == 4 and works[0].shape == (512,) and np.isfinite(works[-1]).all()
  assert coeffs[0][0] == 0 and coeffs[-1][0] == 4 and coeffs[-1][1].shape == (13,)
  assert not np.array_equal(coeffs[0][1], coeffs[-1][1])           # Adam moved the protocol
  prefix = 'end_time_%.10f_dt_%.10f_chebdeg_%i_kappa_%.4f_batch_%i_trainsteps_%i_' % (2e-6, 1e-8, 12, 21.3863, 512, 4)
  assert len(pickle.load(open(out + prefix + 'all_works.pkl', 'rb'))) == 4
  assert pickle.load(open(out + prefix + 'coeffs.pkl', 'rb'))[-1][0] == 4
  # the same loop as CUDA-graph replays with device-resident state: same key, same trajectory of coefficients
  coeffs_d, works_d = run_barrier_crossing_opt.main(['2e-6', '1e-8', '12', '21.3863', '512', '4'], savedir=out,
                                                    quiet=True, device_loop=True)
  assert [c[0] for c in coeffs_d] == [c[0] for c in coeffs]
  np.testing.assert_allclose(coeffs_d[-1][1], coeffs[-1][1], rtol=0, atol=2e-6 * np.abs(coeffs[-1][1]).max())
  np.testing.assert_allclose(works_d[-1], works[-1], rtol=1e-5, atol=1e-5)
  res = run_barrier_crossing_fwd.main(['2e-6', '1e-8', '12', '21.3863', '64', 'test_'], savedir=out)
  assert res['all_pos'].shape == (64, 201, 1, 1) and res['all_works'].shape == (64, 201)
  assert sorted(f for f in os.listdir(out) if 'test_' in f)[0].endswith('all_log_probs.pkl')


def test_ising_drivers(cuda, tmp_path, monkeypatch):
  from noneq_opt_b200.drivers import run_ising_fwd, run_ising_opt
  monkeypatch.setenv('NQ_SEED', '3')
  out = str(tmp_path) + '/'
  res = run_ising_opt.main(['16', '0', '11', '8', '64', '3', '1'], savedir=out, quiet=True)
  assert len(res['entropies']) == 3 and res['entropies'][0].shape == (64, 10)
  assert res['field_weights'][-1].shape == (8,) and np.isfinite(res['field_weights'][-1]).all()
  assert np.abs(res['field_weights'][-1]).max() > 0               # the optimiser stepped
  prefix = '16x16lattice_timesteps_11_chebdeg_8_batch_64_trainsteps_3_'
  for name in ('summaries', 'entropies', 'log_temp_weights', 'field_weights'):
    assert os.path.exists(out + prefix + name + '.pkl')
  fwd = run_ising_fwd.main(['64', '0', '21', '32', '16', '0', '11', '8', '64', '3'], loaddir=out, savedir=out)
  assert fwd['AD']['entropies'].shape == (32, 20) and np.isfinite(fwd['AD']['mean_entropy'])
