"""CPU oracle for the noneq_opt ensemble hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference
(mc2engel/noneq_opt, mounted read-only at /root/reference in the build
container) for the path named by BASELINE.json's `north_star`:

  * `oracle.jaxlike`            -- the un-vendored third-party semantics the path
                                   depends on (jax.random threefry streams,
                                   uniform/normal sampling, XLA's f32 erf_inv,
                                   tfp Normal, distrax Bernoulli, jax_md glue)
  * `oracle.parameterization_ref` -- noneq_opt/parameterization.py
  * `oracle.langevin_ref`       -- noneq_opt/barrier_crossing.py, simulate.py,
                                   potentials.py
  * `oracle.ising_ref`          -- noneq_opt/ising.py
  * `oracle/c/noneq_oracle.c`   -- the same algorithms in plain C (OpenMP over
                                   trajectories / replicas) for the timed CPU baseline

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import, link or execute anything in here, and only
as the checker / the timed CPU arm.  The product package `noneq_opt_b200` never
imports it and has no CPU fallback.

PARITY PINNING STATUS
---------------------
The reference is pure Python on top of jax / tensorflow_probability / distrax /
jax_md, none of which is installed in the build image (no network), so it cannot
be imported as is and no outputs of a live XLA could be recorded.  What pins this
oracle:

  * THE REFERENCE'S OWN SOURCE, executed here.  `oracle/jax_shim/` provides the
    small part of those packages' public API the path uses -- on torch-CPU float32
    tensors, with torch autograd standing in for `jax.grad` and the threefry
    streams of `oracle.jaxlike` -- so that `/root/reference/noneq_opt/*.py` is
    imported VERBATIM (never copied) and run by
    `tests/golden/make_reference_golden.py`; its outputs are committed as
    `tests/golden/reference_{langevin,ising}.npz`.  `tests/test_oracle_reference.py`
    holds the oracle to them: `run_brownian_opt` positions / log-probs / works
    (bit-identical for the harmonic trap, <= 4e-7 for the double well),
    `estimate_gradient_boltzinit` objective and gradient, `simulate_ising` spin
    histories (bit-exact) and summaries, `estimate_gradient_rev` for both losses.
    This pins the oracle's reading of the reference -- step order, key plumbing,
    work / log-prob bookkeeping, what autodiff makes of the estimator objectives.
    It does NOT pin the last bits of XLA's exp / log / erf_inv / sigmoid or its
    reduction order: torch's float32 kernels stand in for those.
  * RNG stream: Random123 threefry2x32-20 known-answer vectors and the values
    published in the jax documentation / test-suite (`split(PRNGKey(0))`,
    `random_bits(PRNGKey(1701))`, `uniform(PRNGKey(0))`, `normal(PRNGKey(0),(1,))`)
    -- tests/test_oracle_rng.py.
  * Ising: the reference's own exact tests (tests/ising_test.py:15-66 --
    sum_neighbors, flip_logits, even_odd_masks) and its property tests
    (detailed balance, heat == T * entropy production, fwd == rev gradient).
  * Langevin: the reference holds NO valid result-pinning test of its own for this
    path (tests/barrier_crossing_test.py is stale, tests/simulate_test.py is smoke
    only), and a live JAX run is not possible here; in that strict sense Langevin
    float parity against XLA remains "parity unpinned".  It is pinned through the
    reference-source fixtures above, the RNG KATs, the physics anchors of BASELINE.md
    (25/e mean work for the linear protocol) and finite-difference checks of the
    closed-form gradients.

Floating point convention of the oracle (both numpy and C forms): arithmetic
(+,-,*,/,sqrt) is IEEE binary32, one rounding per operation, no FMA contraction,
in the operation order of the reference source; transcendental functions
(exp, log, log1p) are evaluated in binary64 and rounded once to binary32
("correctly rounded f32" up to double-rounding), so that the oracle gives the
same bits on every host.
"""
