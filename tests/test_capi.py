"""The C-ABI library builds, loads and exports every symbol include/noneq_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from noneq_opt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'noneq_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(nq_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(built_lib):
  handle = ctypes.CDLL(built_lib)
  declared = _declared_symbols()
  assert len(declared) >= 20
  missing = [s for s in declared if not hasattr(handle, s)]
  assert not missing, f'declared in the header but not exported: {missing}'


def test_binding_covers_header(built_lib):
  assert sorted(_lib.SIGNATURES) == _declared_symbols()
  assert _lib.lib().nq_version() == 100


def test_struct_layouts_match_header(built_lib):
  # 4 int32 + 11 double + 2 int64 ; 2 int32 + int64 + 2 int32 + (1 + 3) int32
  assert ctypes.sizeof(_lib.NqLangevinDesc) == 16 + 11 * 8 + 16
  assert ctypes.sizeof(_lib.NqIsingDesc) == 40


def test_argument_errors_do_not_need_a_gpu(built_lib):
  lib = _lib.lib()
  desc = _lib.NqLangevinDesc(potential=7, steps=10, dt=1e-3, kT=1.0, gamma=1.0, mass=1.0)
  rc = lib.nq_langevin_forward(ctypes.byref(desc), None, None, None, 1, None, None, None, None, 1, None, None,
                               None, None, 0, None)
  assert rc == -1 and b'potential' in lib.nq_last_error()
  idesc = _lib.NqIsingDesc(L=7, T=3, R=1)
  rc = lib.nq_ising_run(ctypes.byref(idesc), None, None, None, None, None, None, None, None, None, None, None,
                        None, 0, None)
  assert rc == -1 and b'must be even' in lib.nq_last_error()
  assert lib.nq_ising_words_per_lattice(64) == 128
  assert lib.nq_langevin_workspace_bytes(100000, 13) > 0


def test_product_never_imports_the_oracle():
  pkg = os.path.join(ROOT, 'noneq_opt_b200')
  for name in os.listdir(pkg):
    if name.endswith('.py'):
      src = open(os.path.join(pkg, name)).read()
      assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), name
