"""The STRICT kernels' main-path math (csrc/nq_common.cuh: logf_main, log1pf_main, div_main and the two-instruction
uniform of normal_from_bits) must return the SAME BITS as the CUDA library functions / the literal transcription
of jax.random.normal they replace -- they are instruction-count optimisations, not approximations.  Checked on the
device through the developer library (libnoneq_b200_dev.so), exhaustively over the operand windows the kernels
guarantee."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(which, first, count, c=0.0):
  from noneq_opt_b200 import _devlib
  out = torch.zeros(2, dtype=torch.int64, device='cuda')
  rc = _devlib.lib().nqdev_math_check(which, first, count, c, ctypes.c_void_p(out.data_ptr()), None)
  assert rc == 0
  torch.cuda.synchronize()
  bad, where = (int(v) & 0xFFFFFFFFFFFFFFFF for v in out.tolist())
  return bad, where


def _bits(exp2):   # bit pattern of 2**exp2
  return (127 + exp2) << 23


def test_logf_main_is_logf_on_every_float_of_its_window(cuda):
  # Landscape::eval guards 2^-62 <= A + B < 2^62: all 124 * 2^23 floats of that window
  bad, where = _check(0, _bits(-62), _bits(62) - _bits(-62))
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'


def test_log1pf_main_is_log1pf_on_every_square_below_one(cuda):
  # erf_inv passes -u*u with |u| <= 1 - 3 * 2^-24: every float in [2^-60, 1) (and well below what u*u reaches)
  bad, where = _check(1, _bits(-60), _bits(0) - _bits(-60))
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'


@pytest.mark.parametrize('c', [-4.183, -1.0, -1e-5, -0.0259, -300.0, 2.5e9, -3e-12])
def test_div_main_is_ieee_division_on_its_window(cuda, c):
  bad, where = _check(2, _bits(-62), _bits(62) - _bits(-62), c)
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'


def test_normal_draw_is_the_literal_jax_formula_for_every_mantissa(cuda):
  # jax.random.normal uses the top 23 bits of the word: all 2^23 patterns
  bad, where = _check(3, 0, 1 << 23)
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'


def test_expf_main_is_expf_on_every_float(cuda):
  # the landscape's exponents are non-positive, but the sequence is libdevice's whole expf: all 2^32 patterns
  # (NaN payloads included -- both return the same quiet NaN)
  bad, where = _check(4, 0, 1 << 32)
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'


def test_packed_expf_pair_is_expf_in_both_halves(cuda):
  # expf_main2 (FFMA2 / FADD2 / FMUL2 on a register pair): {exp(x), exp(-x)} for every float x
  bad, where = _check(5, 0, 1 << 32)
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'


@pytest.mark.parametrize('bde', [0.0, 0.7, -1.3])
def test_packed_landscape_is_the_scalar_reference_order(cuda, bde):
  # Landscape::eval evaluates both wells as f32x2 pairs; ptxas contracts packed multiply + packed add into FFMA2 even
  # with .rn on both, so the right well's "+ beta dE" is a scalar add -- A, B and A + B must carry the bits of the
  # one-rounding-per-operation scalar form for every float x (folded into [-64, 64))
  bad, where = _check(6, 0, 1 << 32, bde)
  assert bad == 0, f'{bad} mismatches, first at pattern {where:#x}'
