"""Generate tests/golden/reference_*.npz by executing the REFERENCE's own source
(/root/reference/noneq_opt/*.py, imported verbatim, never copied) on top of oracle/jax_shim -- a
torch-CPU float32 stand-in for jax / jax_md / tensorflow_probability / distrax whose random streams
come from oracle/jaxlike.py (KAT-pinned threefry).  Run in the build container only:

    python tests/golden/make_reference_golden.py

The fixtures pin the oracle (tests/test_oracle_reference.py) and the CUDA path (tests/test_gpu_reference.py)
to the reference's source: step order, key plumbing, work / log-prob bookkeeping, estimator objectives.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get('NQ_REFERENCE_ROOT', '/root/reference')
sys.path[:0] = [os.path.join(ROOT, 'oracle', 'jax_shim'), REFERENCE, ROOT]

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from jax_md import space  # noqa: E402
from noneq_opt import barrier_crossing as ref_bc  # noqa: E402  (the reference, verbatim)
from noneq_opt import ising as ref_ising  # noqa: E402
from noneq_opt import parameterization as ref_p10n  # noqa: E402
from noneq_opt import potentials as ref_potentials  # noqa: E402


def to_np(x):
  return jax.tree_map(lambda l: l.detach().numpy() if hasattr(l, 'detach') else np.asarray(l), x)


def langevin_cases():
  out = {}
  _, shift = space.free()
  kT = 4.183
  beta, x_m = 1.0 / kT, 14.0
  kappa = 21.3863 / (beta * x_m ** 2)
  cases = {
      'spring': dict(energy=ref_potentials.V_spring(1.0), coeffs=np.asarray(ref_bc.fit_linear_to_cheby(1.0, 1e-3, 5., 10., 12)),
                     r0=(5., 10.), x0=5.0, Neq=20, S=300, dt=1e-3, kT=1.0, mass=1.0, gamma=1.0),
      'shifted': dict(energy=ref_bc.V_biomolecule_shifted(kappa, kappa, x_m, 0., 1.5, beta),
                      coeffs=np.array([20., 19.95996, 0.3, -0.2, 0.1] + [0.] * 8, np.float32),
                      r0=(0., 40.), x0=0.0, Neq=0, S=400, dt=1e-8, kT=kT, mass=1.0, gamma=kT / 1e7),
  }
  # a longer horizon across the barrier (2500 steps): rounding differences get time to act
  cases['shifted_long'] = dict(cases['shifted'], S=2500, dt=4e-8)
  # V_biomolecule (wells at -x_m, +x_m) in a periodic box, with equilibration steps
  cases['biomolecule_periodic'] = dict(energy=ref_bc.V_biomolecule(kappa, kappa, x_m / 2, 0.5, 1.5, beta),
                                       coeffs=np.array([10., 9.5, 0.4, -0.3] + [0.] * 3, np.float32),
                                       r0=(-7., 7.), x0=-7.0, Neq=10, S=300, dt=2e-8, kT=kT, mass=1.0, gamma=kT / 1e7,
                                       box=45.0)
  for name, c in cases.items():
    if 'box' in c:
      _, shift = space.periodic(c['box'])
    else:
      _, shift = space.free()
    B = 6
    coeffs = jnp.array(np.asarray(c['coeffs'], np.float32))
    init_pos = jnp.array(np.full((B, 1, 1), c['x0'], np.float32))
    seed = jax.random.PRNGKey(3)
    # forward trajectories, one call per key, exactly as the estimators vmap it (barrier_crossing.py:221)
    seeds = jax.random.split(seed, B)
    pos, lps, works = [], [], []
    for b in range(B):
      p, l, w = ref_bc.run_brownian_opt(c['energy'], coeffs, init_pos[b], c['r0'][0], c['r0'][1], c['Neq'], shift,
                                        seeds[b], c['S'], c['dt'], c['kT'], c['mass'], c['gamma'])
      pos.append(to_np(p)); lps.append(to_np(l)); works.append(to_np(w))
    est = ref_bc.estimate_gradient_boltzinit(B, c['energy'], c['r0'][0], c['r0'][1], c['Neq'], shift, c['S'], c['dt'],
                                             c['kT'], c['mass'], c['gamma'])
    grad, (estimator, (positions, tot_lp, tot_w)) = est(coeffs, init_pos, seed)
    # the log-prob-only and reparameterisation-only estimators (:262-311) and the Boltzmann pre-pass (:227-256)
    g_logp, _ = ref_bc.estimate_gradient_boltzinit_logP(B, c['energy'], c['r0'][0], c['r0'][1], c['Neq'], shift, c['S'],
                                                        c['dt'], c['kT'], c['mass'], c['gamma'])(coeffs, init_pos, seed)
    g_rep, _ = ref_bc.estimate_gradient_boltzinit_reparam(B, c['energy'], c['r0'][0], c['r0'][1], c['Neq'], shift, c['S'],
                                                          c['dt'], c['kT'], c['mass'], c['gamma'])(coeffs, init_pos, seed)
    eqm = ref_bc.mapped_brownian_eqm(B, c['energy'], init_pos[0], c['r0'][0], shift, 40, c['dt'], c['kT'], c['mass'],
                                     c['gamma'])(jax.random.PRNGKey(5))
    out[name] = dict(grad_logP=to_np(g_logp), grad_reparam=to_np(g_rep), eqm=to_np(eqm), box=np.float32(c.get('box', 0.)),
                     coeffs=np.asarray(c['coeffs'], np.float32), r0=np.array(c['r0'], np.float32), x0=np.float32(c['x0']),
                     Neq=c['Neq'], S=c['S'], dt=c['dt'], kT=c['kT'], mass=c['mass'], gamma=c['gamma'], seed=np.asarray(seed),
                     positions=np.stack(pos), log_probs=np.stack(lps), works=np.stack(works),
                     grad=to_np(grad), estimator=to_np(estimator), tot_log_prob=to_np(tot_lp), total_work=to_np(tot_w),
                     est_positions=to_np(positions))
  return out


def ising_cases():
  out = {}
  for name, shape, T, R in (('small', (16, 16), 12, 3), ('fastpath', (64, 64), 5, 2), ('chain', (48,), 7, 2),
                            ('cube', (4, 6, 8), 6, 2)):
    L = shape[0]
    w_lt = np.array([0.3, -0.2, 0.1], np.float32)
    w_h = np.array([-0.4, 0.6, 0.2], np.float32)
    schedule = ref_ising.IsingSchedule(ref_p10n.Chebyshev(jnp.array(w_lt)), ref_p10n.Chebyshev(jnp.array(w_h)))
    times = jnp.linspace(0, 1, T)
    params = schedule(times)
    spins0 = ref_ising.random_spins(shape, 0.5, jax.random.PRNGKey(11))
    seeds = jax.random.split(jax.random.PRNGKey(7), R)
    finals, summaries, grads = [], [], {'total_entropy_production': [], 'total_work': []}
    states = []
    for r in range(R):
      final_state, (summary, traj) = ref_ising.simulate_ising(params, spins0, seeds[r], return_states=True)
      finals.append(to_np(final_state.spins))
      states.append(to_np(traj.spins))
      summaries.append(np.stack([to_np(getattr(summary, f)) for f in summary._fields]))
      for loss in grads:
        g, _ = ref_ising.estimate_gradient_rev(getattr(ref_ising, loss))(schedule, times, spins0, seeds[r])
        grads[loss].append(np.stack([to_np(g.log_temp.weights), to_np(g.field.weights)]))
    out[name] = dict(L=L, T=T, w_lt=w_lt, w_h=w_h, times=to_np(times), log_temp=to_np(params.log_temp),
                     field=to_np(params.field), spins0=to_np(spins0), seeds=np.asarray(seeds),
                     final_spins=np.stack(finals), states=np.stack(states), summary=np.stack(summaries),
                     grad_entropy=np.stack(grads['total_entropy_production']), grad_work=np.stack(grads['total_work']),
                     summary_fields=np.array(list(ref_ising.IsingSummary._fields)))
  return out


def train_step_case():
  """ising.get_train_step (:263-290) driven for three steps: batch split, vmapped reverse-mode estimator,
  batch mean, clip_grads, Adam.  (Adam itself comes from the shim's restatement of
  jax.example_libraries.optimizers; the step glue is the reference's.)"""
  from jax.example_libraries import optimizers as jopt
  w_lt, w_h = np.array([0.3, -0.2, 0.1], np.float32), np.array([-0.4, 0.6, 0.2], np.float32)
  schedule = ref_ising.IsingSchedule(ref_p10n.Chebyshev(jnp.array(w_lt)), ref_p10n.Chebyshev(jnp.array(w_h)))
  spins0 = -jnp.ones([16, 16])
  B, T, lr = 5, 7, 1e-2
  opt = jopt.adam(lr)
  state = opt.init_fn(schedule)
  step_fn = ref_ising.get_train_step(opt, spins0, B, T, ref_ising.total_entropy_production, 'rev')
  key = jax.random.PRNGKey(21)
  hist, ep = [], []
  for j in range(3):
    key, split = jax.random.split(key)
    state, summary = step_fn(state, j, split)
    p = opt.params_fn(state)
    hist.append(np.stack([to_np(p.log_temp.weights), to_np(p.field.weights)]))
    ep.append(to_np(summary.entropy_production))
  return dict(train=dict(w_lt=w_lt, w_h=w_h, B=B, T=T, lr=lr, seed=np.asarray(jax.random.PRNGKey(21)),
                         weights=np.stack(hist), entropy_production=np.stack(ep)))


def schedule_cases():
  """The protocol algebra of parameterization.py and the baselines of ising.py:45-55, evaluated by the reference."""
  x = jnp.linspace(0, 1, 33)
  rs = np.random.RandomState(4)
  w = (0.3 * rs.normal(size=7)).astype(np.float32)
  knots = (np.linspace(-1, 2, 8)[1:-1] + 0.2 * rs.normal(size=6)).astype(np.float32)
  cheb = ref_p10n.Chebyshev(jnp.array(w))
  pwl = ref_p10n.PiecewiseLinear(jnp.array(knots), jnp.array(0.), jnp.array(1.), jnp.array(-1.), jnp.array(2.))
  lt_base, h_base = ref_ising.log_temp_baseline(0.69, 10., 1), ref_ising.field_baseline(-1., 1.)
  composed = ref_p10n.AddBaseline(ref_p10n.ConstrainEndpoints(cheb, 0., 0.), lt_base)
  moved = ref_p10n.ChangeDomain(cheb, jnp.array(2.), jnp.array(5.))
  x25 = jnp.linspace(2, 5, 33)
  return dict(all=dict(x=to_np(x), w=w, knots=knots, chebyshev=to_np(cheb(x)), piecewise=to_np(pwl(x)),
                       log_temp_baseline=to_np(lt_base(x)), field_baseline=to_np(h_base(x)), composed=to_np(composed(x)),
                       x25=to_np(x25), change_domain=to_np(moved(x25)),
                       constant=to_np(ref_p10n.Constant(jnp.array(0.7))(x))))


def multi_particle_cases():
  """simulate.brownian (simulate.py:33-99) on rank-generic states [N, dim] with the reference's own energy closures
  (which read position[0][0]: potentials.py:27-52): four apply_fn steps, positions / keys / log_prob after each."""
  from noneq_opt import simulate as ref_simulate
  kT = 4.183
  beta, x_m = 1.0 / kT, 14.0
  kappa = 21.3863 / (beta * x_m ** 2)
  out = {}
  for name, energy, shape, r0, dt, kT_, gamma, box in (
      ('spring_7x3', ref_potentials.V_spring(2.0), (7, 3), 0.5, 1e-3, 1.0, 0.1, None),
      ('shifted_4x1', ref_potentials.V_biomolecule_shifted(kappa, kappa, x_m, 0., 1.5, beta), (4, 1), 3.0, 1e-8, kT, kT / 1e7, None),
      ('spring_periodic_5x2', ref_potentials.V_spring(1.0), (5, 2), 1.0, 1e-2, 1.0, 0.5, 3.0)):
    _, shift = space.periodic(box) if box else space.free()
    init_fn, apply_fn = ref_simulate.brownian(energy, shift, dt, kT_, gamma)
    R0 = jnp.array((np.arange(np.prod(shape), dtype=np.float32).reshape(shape) * 0.25 + 0.5).astype(np.float32))
    state = init_fn(jax.random.PRNGKey(9), R0)
    pos, keys, lps = [], [], []
    for k in range(4):
      state = apply_fn(state, jnp.array(np.float32(k * dt)), r0=r0)
      pos.append(to_np(state.position)); keys.append(np.asarray(to_np(state.rng))); lps.append(np.float32(to_np(state.log_prob)))
    out[name] = dict(R0=to_np(R0), r0=np.float32(r0), dt=dt, kT=kT_, gamma=gamma, box=np.float32(box or 0.),
                     positions=np.stack(pos), keys=np.stack(keys), log_probs=np.array(lps, np.float32))
  return out


def ising_long_case():
  """simulate_ising + estimate_gradient_rev over 200 sweeps (T = 201) of a 16 x 16 lattice under the drivers' schedule
  composition: pins the per-sweep key derivation split(seed, T-1), the float32 summary bookkeeping and the REINFORCE
  gradient at a config-like horizon to the reference's own code."""
  L, T, R = 16, 201, 2
  w = (0.1 * np.random.RandomState(0).normal(size=(2, 8))).astype(np.float32)
  schedule = ref_ising.IsingSchedule(
      ref_p10n.AddBaseline(ref_p10n.ConstrainEndpoints(ref_p10n.Chebyshev(jnp.array(w[0])), 0., 0.), ref_ising.log_temp_baseline()),
      ref_p10n.AddBaseline(ref_p10n.ConstrainEndpoints(ref_p10n.Chebyshev(jnp.array(w[1])), 0., 0.), ref_ising.field_baseline()))
  times = jnp.linspace(0, 1, T)
  params = schedule(times)
  spins0 = -jnp.ones([L, L])
  seeds = jax.random.split(jax.random.PRNGKey(0), R)
  finals, states, summaries, grads = [], [], [], {'total_entropy_production': [], 'total_work': []}
  for r in range(R):
    final_state, (summary, traj) = ref_ising.simulate_ising(params, spins0, seeds[r], return_states=True)
    finals.append(to_np(final_state.spins))
    states.append(to_np(traj.spins)[[0, 49, 99, 149, 199]])
    summaries.append(np.stack([to_np(getattr(summary, f)) for f in summary._fields]))
    for loss in grads:
      g, _ = ref_ising.estimate_gradient_rev(getattr(ref_ising, loss))(schedule, times, spins0, seeds[r])
      grads[loss].append(np.stack([to_np(g.log_temp.wrapped.wrapped.weights), to_np(g.field.wrapped.wrapped.weights)]))
  return dict(long=dict(L=L, T=T, w=w, log_temp=to_np(params.log_temp), field=to_np(params.field), seeds=np.asarray(seeds),
                        final_spins=np.stack(finals), states=np.stack(states), state_sweeps=np.array([1, 50, 100, 150, 200]),
                        summary=np.stack(summaries), grad_entropy=np.stack(grads['total_entropy_production']),
                        grad_work=np.stack(grads['total_work'])))


if __name__ == '__main__':
  only = sys.argv[1:]      # e.g. `make_reference_golden.py reference_extra` regenerates one file
  jobs = (('reference_langevin', langevin_cases), ('reference_ising', ising_cases),
          ('reference_schedules', schedule_cases), ('reference_train_step', train_step_case),
          ('reference_extra', lambda: dict(multi_particle_cases(), **ising_long_case())))
  for prefix, make in jobs:
    if only and prefix not in only:
      continue
    cases = make()
    flat = {f'{k}/{f}': v for k, d in cases.items() for f, v in d.items()}
    np.savez_compressed(os.path.join(HERE, prefix + '.npz'), **flat)
    print(prefix, {k: {f: np.shape(v) for f, v in d.items()} for k, d in cases.items()})
