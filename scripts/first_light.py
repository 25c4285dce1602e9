"""Dev script: first GPU contact -- RNG parity, Langevin parity vs the oracle, issue-rate
microbenchmark and a first timing of config 2.  Writes gpurun_out/first_light.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import _devlib, _lib, barrier_crossing as bc, random as nqr, space  # noqa: E402
from oracle import jaxlike as J, langevin_ref as L  # noqa: E402

res = {}
dev = torch.device('cuda')
torch.cuda.init()
print(torch.cuda.get_device_name(0))

# ---- RNG
key = J.PRNGKey(1701)
res['rng_split'] = bool((nqr.keys_to_numpy(nqr.split(key, 1001)) == J.split(key, 1001)).all())
res['rng_bits'] = bool((nqr.bits(key, (777,)).cpu().numpy().view(np.uint32) == J.random_bits(key, (777,))).all())
res['rng_uniform'] = bool((nqr.uniform(key, (1000,)).cpu().numpy() == J.uniform(key, (1000,))).all())
gn, on = nqr.normal(key, (100001,)).cpu().numpy(), J.normal(key, (100001,))
res['rng_normal_maxrel'] = float(np.max(np.abs(gn - on) / np.maximum(np.abs(on), 1e-30)))
res['rng_normal_exact_frac'] = float((gn == on).mean())
print(res)


def parity(name, pot_o, pot_g, coeffs, x0v, r0i, r0f, Neq, S, B, dt, kT, mass, gamma, mode):
  bc.set_math_mode(mode)
  _, shift = space.free()
  seed = J.PRNGKey(0)
  x0 = np.full(B, x0v, np.float32)
  t = time.time()
  o = L.estimate_gradient(pot_o, coeffs, x0, seed, r0i, r0f, Neq, S, dt, kT, mass, gamma, record=True)
  o64 = L.estimate_gradient(pot_o, coeffs, x0, seed, r0i, r0f, Neq, S, dt, kT, mass, gamma, dtype=np.float64, record=True)
  t_or = time.time() - t
  seeds = nqr.split(seed, B)
  g = bc.gradient_terms(pot_g, coeffs, x0.reshape(B, 1, 1), seeds, r0i, r0f, Neq, shift, S, dt, kT, mass, gamma)
  pos, lps, works = bc.run_brownian_opt(pot_g, coeffs, x0.reshape(B, 1, 1), r0i, r0f, Neq, shift, seeds, S, dt, kT, mass, gamma)
  torch.cuda.synchronize()
  pos = pos.cpu().numpy()[:, :, 0, 0]
  W = g['work'].cpu().numpy()
  LP = g['logp'].cpu().numpy()
  scale = np.abs(o['traj']['positions']).max()
  out = dict(
      oracle_s=t_or,
      pos_maxrel=float(np.abs(pos - o['traj']['positions']).max() / scale),
      pos64_maxrel_gpu=float(np.abs(pos - o64['traj']['positions']).max() / scale),
      pos64_maxrel_oracle=float(np.abs(o['traj']['positions'] - o64['traj']['positions']).max() / scale),
      work_maxrel=float((np.abs(W - o['total_work']) / np.abs(o['total_work']).max()).max()),
      work64_maxrel_gpu=float((np.abs(W - o64['total_work']) / np.abs(o64['total_work']).max()).max()),
      work64_maxrel_oracle=float((np.abs(o['total_work'] - o64['total_work']) / np.abs(o64['total_work']).max()).max()),
      logp_maxrel=float((np.abs(LP - o['tot_log_prob']) / np.abs(o['tot_log_prob']).max()).max()),
      stepwork_maxabs=float(np.abs(works.cpu().numpy() - o['traj']['works']).max()),
      steplogp_maxabs=float(np.abs(lps.cpu().numpy() - o['traj']['log_probs']).max()),
      mean_work=float(W.mean()), mean_work_oracle=float(o['total_work'].mean()),
  )
  gs = g['grad_sums'].cpu().numpy() / B
  for i, nm in enumerate(['grad_logP', 'grad_reparam']):
    ref = o[nm]
    out[nm + '_maxrel'] = float(np.abs(gs[i] - ref).max() / np.abs(ref).max())
    out[nm + '_vs64'] = float(np.abs(gs[i] - o64[nm]).max() / np.abs(o64[nm]).max())
    out[nm + '_oracle_vs64'] = float(np.abs(ref - o64[nm]).max() / np.abs(o64[nm]).max())
  res[f'{name}_{mode}'] = out
  print(name, mode, json.dumps(out, indent=1))


kT = 4.183
beta = 1 / kT
xm = 14.
kap = 21.3863 / (beta * xm ** 2)
c2 = np.array([2.00000000e+01, 1.99599600e+01, -2.63245988e-15, 8.41484236e-16, 0.0, -7.47209997e-15,
               -1.69086360e-14, 1.27892999e-14, 3.02606960e-15, -6.05805446e-15, -6.36600761e-16,
               -5.53701086e-15, -1.14938537e-14], np.float32)
for mode in ('strict', 'fast'):
  parity('c1', L.V_spring(1.0), bc.V_spring(1.0), L.fit_linear_to_cheby(1.0, 1e-3, 5, 10, 12), 5.0, 5., 10., 100,
         1000, 1000, 1e-3, 1.0, 1.0, 1.0, mode)
  parity('c2s', L.V_biomolecule_shifted(kap, kap, xm, 0., 1.5, beta), bc.V_biomolecule_shifted(kap, kap, xm, 0., 1.5, beta),
         c2, 0.0, 0., 40., 0, 2000, 512, 1e-8, kT, 1.0, kT / 1e7, mode)

# ---- microbench
lib = _lib.lib()
names = ['LOP3', 'SHF', 'IADD', 'IMAD', 'FFMA', 'IMAD.WIDE+LOP', 'tf_round', 'MUFU.EX2', 'tf_round_wide',
         'LOP3+FFMA', 'LOP3+IMAD', 'threefry_block']
ops_per_body = [1, 1, 1, 1, 1, 2, 3, 1, 3, 2, 2, 73]
mb = {}
for threads in (256, 1024):
  for w, nm in enumerate(names):
    blocks, iters = 148 * (2 if threads == 1024 else 8), 2000 if w != 11 else 100
    cyc = torch.zeros(blocks, dtype=torch.int64, device=dev)
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    for rep in range(2):
      _lib.check(_devlib.lib().nqdev_microbench(w, blocks, threads, iters, _lib.ptr(cyc), _lib.ptr(sink), None))
      torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_devlib.lib().nqdev_microbench(w, blocks, threads, iters, _lib.ptr(cyc), _lib.ptr(sink), None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_instr = blocks * threads * iters * 8 * 8 * ops_per_body[w]
    per_sm_resident = blocks // 148
    mean_cyc = float(cyc.double().mean())
    mb[f'{nm}@{threads}'] = dict(ms=ms, thread_instr_per_s=n_instr / (ms * 1e-3),
                                 instr_per_clk_per_sm=threads * per_sm_resident * iters * 64 * ops_per_body[w] / mean_cyc)
    print(nm, threads, mb[f'{nm}@{threads}'])
res['microbench'] = mb

# ---- timing config 2
B, S = 100_000, 10_000
_, shift = space.free()
pot = bc.V_biomolecule_shifted(kap, kap, xm, 0., 1.5, beta)
seeds = nqr.split(J.PRNGKey(0), B)
x0 = torch.zeros(B, 1, 1, device=dev)
tim = {}
for mode in ('strict', 'fast'):
  bc.set_math_mode(mode)
  for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g = bc.gradient_terms(pot, c2, x0, seeds, 0., 40., 0, shift, S, 1e-8, kT, 1.0, kT / 1e7)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
  tim[mode] = dict(ms=ms, traj_steps_per_s=B * S / (ms * 1e-3), mean_work=float(g['moments'][1] / B),
                   grad=(g['grad_sums'].sum(0) / B).tolist())
  print(mode, tim[mode])
res['c2_timing'] = tim
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/first_light.json', 'w'), indent=1)
