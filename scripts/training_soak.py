"""Dev: a full optimisation of the barrier-crossing protocol through the device-resident loop
(run_barrier_crossing_opt's loop at the reference's own settings: end_time 1e-4 s, dt 1e-8 s, degree 12,
kappa 21.3863, batch 10^4 .. 10^5): mean work must go down, nothing may hang or turn non-finite."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noneq_opt_b200.drivers import run_barrier_crossing_opt as drv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
os.makedirs('gpurun_out/soak', exist_ok=True)
t0 = time.perf_counter()
coeffs, works = drv.main(['1e-4', '1e-8', '12', '21.3863', str(B), str(steps)], savedir='gpurun_out/soak/', quiet=True,
                         device_loop=True, key=np.array([0, 2024], np.uint32))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mw = np.array([w.astype(np.float64).mean() for w in works])
print(f'{steps} training steps of {B} x 10^4 trajectory-steps in {dt:.1f} s ({B * 1e4 * steps / dt:.3g} traj-steps/s incl. equilibration pre-pass)')
print('mean work first 5   :', np.round(mw[:5], 3))
print('mean work last 5    :', np.round(mw[-5:], 3))
print('all finite          :', bool(np.isfinite(mw).all() and np.isfinite(coeffs[-1][1]).all()))
print('final coefficients  :', np.round(coeffs[-1][1], 3))
for f in os.listdir('gpurun_out/soak'):
  os.remove(os.path.join('gpurun_out/soak', f))
