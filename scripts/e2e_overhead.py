"""Dev: where the host time of one end-to-end gradient call goes (config 2)."""
import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, distributed as nqd, random as nqr, space
KT = 4.183; BETA = 1 / KT; XM = 14.; KAPPA = 21.3863 / (BETA * XM ** 2)
pot = bc.V_biomolecule_shifted(KAPPA, KAPPA, XM, 0., 1.5, BETA)
_, shift = space.free()
B, S = 100_000, 10_000
coeffs = torch.tensor([20., 19.95996] + [0.] * 11).pin_memory()
x0 = torch.zeros(B).pin_memory()
seed = nqr.PRNGKey(0)
fn = nqd.langevin_gradient_sharded(B, pot, 0., 40., 0, shift, S, 1e-8, KT, 1.0, KT / 1e7)
def call():
  c = coeffs.to('cuda', non_blocking=True); x = x0.to('cuda', non_blocking=True)
  return fn(c, x, seed)
for _ in range(3):
  call(); torch.cuda.synchronize()
ts = []
for _ in range(10):
  torch.cuda.synchronize(); t0 = time.perf_counter(); call(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
  ts.append((t1 - t0, t2 - t0))
print('host time to return (ms):', np.median([a for a, _ in ts]) * 1e3, ' total (ms):', np.median([b for _, b in ts]) * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
  call()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)

out_host = [torch.empty(B).pin_memory() for _ in range(3)] + [torch.empty(13).pin_memory()]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def call_full(with_flush):
  if with_flush:
    flush.fill_(1)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  grad, aux = call()
  out_host[3].copy_(grad, non_blocking=True)
  for h, k in zip(out_host, ('work', 'logp', 'estimator')):
    h.copy_(aux['local'][k], non_blocking=True)
  torch.cuda.synchronize()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1)
for wf in (False, True):
  ts = [call_full(wf) for _ in range(8)]
  print('with D2H, flush =', wf, ': event ms', np.median(ts))
def resident(with_flush):
  if with_flush:
    flush.fill_(1)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); call(); e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1)
for wf in (False, True):
  print('no D2H, flush =', wf, ': event ms', np.median([resident(wf) for _ in range(8)]))
