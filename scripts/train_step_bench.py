"""Dev script: training steps per second of config 1 (10^3 trajectories x 10^3 steps, spring) through
(a) the step-by-step host loop, (b) LangevinTrainer without a graph, (c) LangevinTrainer's CUDA graph."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200 import barrier_crossing as bc, optimizers, random as nqr, space, training  # noqa: E402

_, shift = space.free()
B, S, Neq, dt = 1000, 1000, 100, 1e-3
energy = bc.V_spring(1.0)
coeffs0 = bc.fit_linear_to_cheby(1.0, dt, 5., 10., 12)
key = np.array([0, 7], np.uint32)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200


def host_loop(n):
  opt = optimizers.adam(optimizers.exponential_decay(0.1, N, 0.001))
  state = opt.init_fn(coeffs0)
  grad_fn = bc.estimate_gradient_boltzinit(B, energy, 5., 10., Neq, shift, S, dt, 1.0, 1.0, 1.0, record_positions=False)
  k = key
  for j in range(n):
    k, k1, k2 = nqr.split_host(k, 3)
    grad, _ = grad_fn(opt.params_fn(state), np.full((B, 1, 1), 5., np.float32), k2)
    state = opt.update_fn(j, grad, state)
  torch.cuda.synchronize()
  return opt.params_fn(state)


host_loop(5)
t = time.perf_counter(); host_loop(N); t_host = time.perf_counter() - t
print(f'host loop      : {N / t_host:9.1f} steps/s ({1e3 * t_host / N:.3f} ms/step)')
for graph in (False, True):
  tr = training.LangevinTrainer(B, energy, coeffs0, 5., 10., Neq, shift, S, dt, 1.0, 1.0, 1.0, key,
                                step_size=0.1, decay_steps=N, decay_rate=0.001, use_graph=graph)
  tr.run(5); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t = time.perf_counter(); e0.record()
  for _ in range(N):
    tr.step()
  e1.record(); torch.cuda.synchronize(); dt_w = time.perf_counter() - t
  print(f'trainer graph={graph!s:5}: {N / dt_w:9.1f} steps/s ({1e3 * dt_w / N:.3f} ms/step wall, {e0.elapsed_time(e1) / N:.3f} ms/step device)')
