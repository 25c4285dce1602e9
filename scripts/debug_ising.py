import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import ising
from oracle import ising_ref as I, jaxlike as J
np.set_printoptions(linewidth=200, precision=6, suppress=True)
for L in (64, 10, 256):
  T = 4
  lt = np.log(np.array([2.0, 1.5, 2.5, 1.0], np.float32)); h = np.array([-0.5, 0.2, 0.4, -0.3], np.float32)
  s0 = I.random_spins((L, L), .5, J.PRNGKey(3)); seed = J.PRNGKey(9)
  trace = []
  fin, summ = I.simulate_ising(lt, h, s0, seed, trace=trace)
  _, _, spins, _, _, seeds, broadcast, bits = ising._prepare(ising.IsingParameters(lt, h), s0, seed)
  run = ising._run_kernel(lt, h, bits, broadcast, seeds, L, want_partials=True, want_counts=True)
  S = run.summary.cpu().numpy()[:, 0]; P = run.partials.cpu().numpy()[:, 0]; C = run.counts.cpu().numpy()[0]
  print('L', L, 'MB0', run.MB0.cpu().numpy(), 'oracle M0,B0', s0.sum(), (s0 * I.sum_neighbors(s0)).sum() // 2)
  for f, name in enumerate(I.IsingSummary._fields):
    print(f'{name:20s} gpu {S[f]}  oracle {getattr(summ, name)}')
  thr, lut = ising.class_tables(lt, h)
  cf = I.closed_form_gradient('total_work', lt, h, np.eye(T), np.eye(T), s0, seed)
  print('dfwd_dh gpu', P[0], 'oracle', cf['per_step']['dfwd_dh'][1:])
  print('dfwd_dl gpu', P[1], 'oracle', cf['per_step']['dfwd_dl'][1:])
  # host recomputation from counts
  M = s0.sum(); B = (s0 * I.sum_neighbors(s0)).sum() // 2
  for t in range(1, T):
    dM = dB = 0; sF = sAs = lF = lAs = 0.0
    for half in range(2):
      for up in (0, 1):
        for n in range(5):
          slot = 8 * up + n; s = 1 if up else -1; nsum = 2 * n - 4
          A, F = C[t - 1, half, 0, slot], C[t - 1, half, 1, slot]
          dM += -2 * s * F; dB += -2 * s * nsum * F
          sAs += s * A * lut[t - 1, 3, slot]; lAs += lut[t - 1, 0, slot] * A * lut[t - 1, 3, slot]; lF += F * lut[t - 1, 0, slot]
    M += dM; B += dB
    einv = np.exp(-np.float64(lt[t]))
    print(' host-from-counts t', t, 'E_fin', -(B + h[t] * M), 'dfwd_dh', -2 * einv * (-0.5 * dM - sAs), 'dfwd_dl', -(lF - lAs))
