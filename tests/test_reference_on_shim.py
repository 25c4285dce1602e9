"""The reference's OWN tests, run over oracle/jax_shim (only where /root/reference is mounted: the build
container).  This is what makes the shim a credible stand-in when it generates tests/golden/reference_*.npz:
the reference's exact-value tests of sum_neighbors / flip_logits / even_odd_masks, its detailed-balance and
heat = T x entropy-production property tests on a small lattice, its forward-mode == reverse-mode gradient
check and its control_flow tests pass on it.  Nothing here touches the GPU or the product package."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get('NQ_REFERENCE_ROOT', '/root/reference')

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'tests')),
                                reason='the reference tree is only mounted in the build container')


def run_reference_tests(files, select):
  env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1',
             PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'oracle', 'jax_shim'), REFERENCE, ROOT]))
  cmd = [sys.executable, '-m', 'pytest', '-q', '-p', 'no:cacheprovider', '-k', select] + \
        [os.path.join(REFERENCE, 'tests', f) for f in files]
  return subprocess.run(cmd, cwd='/tmp', env=env, capture_output=True, text=True, timeout=900)


def test_reference_ising_library_and_gradient_tests_pass_on_the_shim():
  r = run_reference_tests(['ising_test.py'], 'TestLibraryFunctions or test_gradient_estimates_equivalent')
  tail = r.stdout.strip().splitlines()[-1]
  assert r.returncode == 0 and ' passed' in tail and 'failed' not in tail, r.stdout[-2000:]


def test_reference_control_flow_tests_pass_on_the_shim():
  r = run_reference_tests(['control_flow_test.py'], 'test')
  tail = r.stdout.strip().splitlines()[-1]
  assert r.returncode == 0 and ' passed' in tail and 'failed' not in tail, r.stdout[-2000:]
