// Batched overdamped-Langevin integrator with per-step work / log-prob accumulation, the fused
// REINFORCE / log-prob / reparameterisation gradient moments, and the checkpoint + replay
// ("adjoint") pair (sm_100a).  Kernel templates; instantiated per potential in langevin_pot*.cu.
//
// Replaces the XLA computation traced from the reference's
//   brownian / apply_fn              noneq_opt/barrier_crossing.py:28-94
//   run_brownian_opt (lax.scan)      noneq_opt/barrier_crossing.py:97-143
//   estimate_gradient_boltzinit*     noneq_opt/barrier_crossing.py:204-224, 262-333
//   run_brownian_stationary_trap     noneq_opt/barrier_crossing.py:227-241
//
// Design: one thread per trajectory; position, threefry key chain, work / log-prob sums and the
// 2*D gradient accumulators live in registers for the whole trajectory.  The trap table r[k] and
// the protocol Jacobian phi[k][:] are staged tile by tile into shared memory and read as
// warp-uniform broadcasts.  The key chain does not depend on the positions, so it is software
// pipelined one step ahead: each iteration hashes the pending sub-key (the normal draw of this
// step) and splits the next key at the same time -- three independent threefry2x32-20 blocks in
// flight per thread.  HBM traffic is ~24 B per *trajectory*; the kernel is bound by instruction
// issue.
#pragma once
#include <math.h>

#ifndef NQ_PIN_STATIC
#define NQ_PIN_STATIC 1
#endif
#ifndef NQ_PIN_CONSTANTS
#define NQ_PIN_CONSTANTS 1
#endif
#ifndef NQ_FAST_EXP   // precision-study toggles of the FAST path (scripts/precision_study.py)
#define NQ_FAST_EXP 0
#endif
#ifndef NQ_FAST_RCP
#define NQ_FAST_RCP 0
#endif
#ifndef NQ_FAST_NORMAL   // 1: FAST draws are the exact jax.random.normal values (the hot loops read them from the table,
#define NQ_FAST_NORMAL 1  // the other kernels compute them), 0: the MUFU approximation of round 1 outside the hot loops
#endif

#include "nq_common.cuh"

namespace nq {

// STRICT work increment: 0 = the reference's literal difference of two total energies (barrier_crossing.py:123-124),
// 1 = the same quantity in closed form, (k_s/2)[(x-r_k)^2 - (x-r_{k-1})^2] = a_k x + b_k (the landscape term cancels
// exactly), one rounding per step
#ifndef NQ_STRICT_ALG_WORK
#define NQ_STRICT_ALG_WORK 1
#endif
#ifndef NQ_LANGEVIN_UNROLL
#define NQ_LANGEVIN_UNROLL 1
#endif
constexpr int kUnroll = NQ_LANGEVIN_UNROLL;
#ifndef NQ_LANGEVIN_TILE
#define NQ_LANGEVIN_TILE 128
#endif
constexpr int kTile = NQ_LANGEVIN_TILE;    // steps staged per shared-memory tile
#ifndef NQ_LANGEVIN_THREADS
#define NQ_LANGEVIN_THREADS 64
#endif
constexpr int kThreads = NQ_LANGEVIN_THREADS;  // trajectories per CTA (small CTAs balance 1e5 trajectories over 148 SMs)
constexpr int kCtaScale = 64 / kThreads;       // CTAs per 64 trajectories
static_assert(kThreads == 32 || kThreads == 64, "32 or 64 trajectories per CTA");
constexpr bool kAlgWork = NQ_STRICT_ALG_WORK != 0;
#ifndef NQ_PACKED_GRAD   // STRICT gradient moments as FFMA2 pairs (bit-identical to the scalar FFMAs; fewer issue slots).
// Not in FAST mode: measured 9.4e10 against 1.06e11 scalar -- ptxas shuffles the pairs through 25 extra moves per step there
#define NQ_PACKED_GRAD 1
#endif
constexpr bool kPackedGrad = NQ_PACKED_GRAD != 0;
#ifndef NQ_NORMAL_TABLE   // STRICT hot loops: the normal draw as one load from the device-wide 2^23-entry table
#define NQ_NORMAL_TABLE 1
#endif
constexpr bool kNormalTable = NQ_NORMAL_TABLE != 0;
#ifndef NQ_NORMAL_TABLE_FAST   // FAST hot loops too: the table holds the exact draw, so FAST keeps only its landscape approximations
#define NQ_NORMAL_TABLE_FAST 1
#endif
constexpr bool kNormalTableFast = kNormalTable && NQ_NORMAL_TABLE_FAST != 0;
#ifndef NQ_PACKED_LAND   // STRICT landscape: the two wells' squares, exponentials and force terms as f32x2 pairs
#define NQ_PACKED_LAND 1
#endif
constexpr bool kPackedLand = NQ_PACKED_LAND != 0;
#ifndef NQ_TAB_IN_PHI   // D = 13: (r_k, a_k, b_k) ride in the three padding slots of the 16-float phi row (one LDS.128 fewer per step)
#define NQ_TAB_IN_PHI 1
#endif
constexpr bool tab_in_phi(int D) { return NQ_TAB_IN_PHI != 0 && D > 0 && (D + 3) / 4 * 4 - D == 3; }
// resident CTAs per SM the register budget is held to, so that config 2 (1563 CTAs of 64) is one wave
// Resident CTAs per SM the kernels are compiled for (the register budget).  Measured at the config-5 shard size
// (1.25e6 trajectories, static launch; gpurun_out/variants_r02m.log, 1e10 traj-steps/s STRICT / FAST):
// 11 CTAs (88 registers) 8.61 / 10.1, 10: 8.34 / 10.7, 9: 8.36 / 10.7, 8 (128 registers): 8.24 / 10.8 --
// STRICT wants the warps, FAST the registers.
#ifndef NQ_STATIC_MINB_STRICT
#define NQ_STATIC_MINB_STRICT 11
#endif
#ifndef NQ_STATIC_MINB_FAST
#define NQ_STATIC_MINB_FAST 8
#endif
constexpr int min_blocks(int D, bool fast) {
  return kCtaScale * (D <= 8 ? 16 : (D <= 16 ? (fast ? NQ_STATIC_MINB_FAST : NQ_STATIC_MINB_STRICT) : (D <= 24 ? 8 : 6)));
}
// The persistent kernels run 8 (STRICT) / 6 (FAST) CTAs per SM at config 2 (measured best), so they may use up to 128
// registers: STRICT uses 108, +3 % over the 96-register build of round 1 (gpurun_out/variants_r02j.log: 8.39 against
// 8.14e10).
#ifndef NQ_PERSIST_MINB
#define NQ_PERSIST_MINB 8
#endif
constexpr int persistent_min_blocks(int D, bool fast) { return D > 16 ? min_blocks(D, fast) : NQ_PERSIST_MINB * kCtaScale; }

struct LangevinConsts {
  int potential, shift_kind;
  float hk, k_s, inv_beta, cl, cr, bde, xm, two_xm;
  float nu, dt, sigma, inv_sigma, log_norm, box;
  float ca;     // nu*dt*k_s/sigma : d logp_k / d r_k = z_k * ca
  float nudt;   // nu*dt (fast mode only)
  // FAST-math constants (see fast_mean): log2(e)-scaled exponents, (nu dt / beta)-scaled force terms
  float f_cl2, f_ncr2, f_nbde2, f_clm, f_crm, f_cu;
  // STRICT: width of the operand window [2^-62, 2^62) in which the division / logarithm of A + B take their
  // main paths (div_main, logf_main); 0 sends every step through the library functions (extreme 1 / beta)
  uint32_t guard_span;
  f32x2 well_c2;                    // {cl, -cr}: the packed landscape's constant pair (kPackedLand)
  float exp_k1, exp_k252, l1p_c9;   // StrictRegs: register-held multiplicands of expf_main / log1pf_main
  float guard_lo;   // 2^-62 when A + B is bounded above by construction (in_window_low), else +inf
  uint32_t one;   // = 1: multiplicand of the IMAD-form additions of the key chain (threefry2x32_inj)
  long long S, Neq;
};

int make_langevin_consts(const NqLangevinDesc* d, LangevinConsts* c);

// ------------------------------------------------------------------------------ device math
// STRICT (FAST=false): the reference's operation order with one binary32 rounding per operation
// (__f*_rn intrinsics are never contracted), IEEE division/sqrt, accurate expf/logf/log1pf.
// FAST: FMA contraction, MUFU exp2/log2/rcp, work increment in its algebraic form.
template <bool FAST> __device__ __forceinline__ float fmul(float a, float b) { return FAST ? a * b : __fmul_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fadd(float a, float b) { return FAST ? a + b : __fadd_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fsub(float a, float b) { return FAST ? a - b : __fsub_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fexp(float a) { return FAST ? __expf(a) : expf(a); }
template <bool FAST> __device__ __forceinline__ float flog(float a) { return FAST ? __logf(a) : logf(a); }
template <bool FAST> __device__ __forceinline__ float fdiv(float a, float b) { return FAST ? __fdividef(a, b) : __fdiv_rn(a, b); }

// division and logarithm of A + B outside the main-path window (never taken for physical parameters:
// A + B < 2^-62 needs a particle > 9 well-widths from both wells); kept out of line
static __device__ __noinline__ float2 div_log_library(float c, float s) {
  return make_float2(__fdiv_rn(c, s), logf(s));
}

// energy pieces shared between increment_work (:123-124) and the force (:60,68)
template <int POT, bool FAST>
struct Landscape {
  float ul, ur, A, B, sAB;
  float g, lg;   // STRICT: -(1/beta) / (A + B) and log(A + B)
  // CHECKED = false: the caller tests in_window() itself, after its straight-line code (hot loops)
  template <bool CHECKED = true>
  __device__ __forceinline__ void eval(const LangevinConsts& c, float x) {
    if (POT == NQ_POT_SPRING) return;
    if (POT == NQ_POT_BIOMOLECULE) {
      ul = fadd<FAST>(x, c.xm);
      ur = fsub<FAST>(x, c.xm);
    } else {
      ul = x;
      ur = fsub<FAST>(x, c.two_xm);
    }
    if (!FAST && kPackedLand) {
      // both wells side by side in FFMA2 / FMUL2 / FADD2 pairs {left, right}: every half is the scalar operation with
      // the scalar rounding (the right well carries -cr and -bde: negation commutes with every rounding, so the pair
      // holds a and -v bit for bit)
      // (the right well's "+ beta dE" stays a scalar add: ptxas contracts a packed multiply and a packed add into
      // one FFMA2 even when both carry .rn, which would skip the rounding of cr ur^2 the reference performs)
      float a0, ncv;
      f2_unpack(f2_mul(c.well_c2, f2_sq(f2_pack(ul, ur))), a0, ncv);
      const f32x2 arg2 = f2_pack(a0, __fsub_rn(ncv, c.bde));
      const f32x2 AB2 = expf_main2(arg2, c.exp_k1, c.exp_k252);
      f2_unpack(AB2, A, B);
      sAB = __fadd_rn(A, B);
      g = div_main(-c.inv_beta, sAB);
      lg = logf_main(sAB);
      if (CHECKED && !in_window(c)) fix_range(c);
      return;
    }
    const float a = fmul<FAST>(c.cl, fmul<FAST>(ul, ul));
    const float v = fadd<FAST>(fmul<FAST>(c.cr, fmul<FAST>(ur, ur)), c.bde);
    A = FAST ? __expf(a) : expf_main(a, c.exp_k1, c.exp_k252);
    B = FAST ? __expf(-v) : expf_main(-v, c.exp_k1, c.exp_k252);
    sAB = fadd<FAST>(A, B);
    if (!FAST) {
      // main paths, no branch (whichever of the two the caller does not use is dead code); callers patch
      // the result with fix_range() when !in_window()
      g = div_main(-c.inv_beta, sAB);
      lg = logf_main(sAB);
      if (CHECKED && !in_window(c)) fix_range(c);
    }
  }
  // STRICT: A + B inside the operand window of div_main / logf_main?  (always true for SPRING)
  __device__ __forceinline__ bool in_window(const LangevinConsts& c) const {
    return POT == NQ_POT_SPRING || (__float_as_uint(sAB) - 0x20800000u) < c.guard_span;
  }
  // the same test as one float compare, for parameters whose A + B cannot exceed the window (guard_lo)
  __device__ __forceinline__ bool in_window_low(const LangevinConsts& c) const {
    return POT == NQ_POT_SPRING || sAB >= c.guard_lo;
  }
  __device__ __forceinline__ void fix_range(const LangevinConsts& c) {
    const float2 t = div_log_library(-c.inv_beta, sAB);
    g = t.x;
    lg = t.y;
  }
  __device__ __forceinline__ float Em(const LangevinConsts& c) const {
    return POT == NQ_POT_SPRING ? 0.f : (FAST ? -c.inv_beta * __logf(sAB) : __fmul_rn(-c.inv_beta, lg));
  }
  // HALF of dEm/dx, in the order of jax's VJP rules (log: g/x, exp: g*ans, x**2: g*(2x)).  The factor 2 of
  // d(x^2)/dx is a power of two: rn(t * (2 u)) = 2 rn(t u) and rn(2 a + 2 b) = 2 rn(a + b) exactly (no
  // intermediate of this path comes near the subnormal range), so the reference's 2ul, 2ur and 2(x - r)
  // are carried as ONE factor that the step applies in its final fused add -- same bits, three multiplies
  // fewer per step.
  __device__ __forceinline__ float half_dEm(const LangevinConsts& c) const {
    if (POT == NQ_POT_SPRING) return 0.f;
    if (!FAST && kPackedLand) {   // {ga, gb} -> {(ga cl) ul, ((-gb) cr) ur}, pairwise, then the one scalar add
      const f32x2 t2 = f2_mul(f2_mul(f2_mul(f2_pack(g, g), f2_pack(A, B)), c.well_c2), f2_pack(ul, ur));
      float ta, tb;
      f2_unpack(t2, ta, tb);
      return __fadd_rn(ta, tb);
    }
    const float gq = FAST ? __fdividef(-c.inv_beta, sAB) : g;
    const float ga = fmul<FAST>(gq, A), gb = fmul<FAST>(gq, B);
    const float ta = fmul<FAST>(fmul<FAST>(ga, c.cl), ul);
    const float tb = fmul<FAST>(fmul<FAST>(-gb, c.cr), ur);
    return fadd<FAST>(ta, tb);
  }
};

template <bool FAST>
__device__ __forceinline__ float spring_energy(const LangevinConsts& c, float x, float r) {
  const float u = fsub<FAST>(x, r);
  return fmul<FAST>(c.hk, fmul<FAST>(u, u));
}

// dW = E(x; r_cur) - E(x; r_prev), position BEFORE the kick (:123-124, 128)
template <int POT, bool FAST>
__device__ __forceinline__ float work_increment(const LangevinConsts& c, const Landscape<POT, FAST>& land, float x,
                                                float r_prev, float r_cur) {
  if (FAST) {  // (k_s/2) [(x-r1)^2 - (x-r0)^2] = (k_s/2) (r0 - r1) (2x - r0 - r1); the landscape term cancels
    return c.hk * (r_prev - r_cur) * ((x - r_prev) + (x - r_cur));
  }
  if (POT == NQ_POT_SPRING) return __fsub_rn(spring_energy<false>(c, x, r_cur), spring_energy<false>(c, x, r_prev));
  const float em = land.Em(c);
  return __fsub_rn(__fadd_rn(em, spring_energy<false>(c, x, r_cur)), __fadd_rn(em, spring_energy<false>(c, x, r_prev)));
}

// Shared-memory reads through an explicit 32-bit shared-window address: keeps the per-iteration
// generic->shared address arithmetic (S2UR/ULEA/UMOV, ~9 issue slots per step) out of the hot loop.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ float py_mod(float x, float box) {
  float r = fmodf(x, box);
  if (r != 0.f && ((r < 0.f) != (box < 0.f))) r += box;
  return r;
}

// The software-pipelined key chain of one trajectory.  Invariant at the top of step k:
// (k0, k1) is the chain key AFTER the split of step k, (s0, s1) the sub-key step k samples with.
// INJ: 0 plain threefry; 1 key-injection adds as IMADs (threefry2x32_inj); 2 every add as IMAD.  Measured at
// config 2 (gpurun_out/variants_r02b.log, 1e10 traj-steps/s, STRICT / FAST): 0: 8.15 / 9.54, 1: 7.96 / 9.89,
// 2: 7.67 / 9.49 -- the FAST loop (fewer FMA-pipe instructions) gains from moving the injections off the ALU
// pipe, the STRICT loop loses, and true IMADs (half rate, one pipe) for every add lose in both.
#ifndef NQ_RNG_INJ_STRICT
#define NQ_RNG_INJ_STRICT 0
#endif
#ifndef NQ_RNG_INJ_FAST
#define NQ_RNG_INJ_FAST 1
#endif
template <int INJ>
struct RngPipeT {
  uint32_t k0, k1, s0, s1;
  uint32_t one = 1u;   // set from LangevinConsts::one by the hot loops (a value the compiler cannot fold)
  // state.rng -> pipeline (performs the split of the first step)
  __device__ __forceinline__ void prime(uint32_t r0, uint32_t r1) { split2(r0, r1, k0, k1, s0, s1); }
  // bits for this step's normal draw; advances the chain by one step
  __device__ __forceinline__ uint32_t next() {
    uint32_t bits, unused, n0, n1, t0, t1;
    if (INJ) {
      const KeyInject ks = key_inject(s0, s1, one);
      threefry2x32_inj<INJ == 2>(ks, ks.ks0, ks.ks1, one, bits, unused);   // normal(split, [1,1,1]) :88
      split2_inj<INJ == 2>(k0, k1, one, n0, n1, t0, t1);       // key, split = split(state.rng) of the NEXT step :79
    } else if (NQ_TF_WIDE_MASK != 0u) {   // experiment: see threefry2x32_w
      threefry2x32_w(key_schedule(s0, s1), 0u, 0u, one, bits, unused);
      const KeySchedule kk = key_schedule(k0, k1);
      threefry2x32_w(kk, 0u, 2u, one, n0, t0);
      threefry2x32_w(kk, 1u, 3u, one, n1, t1);
    } else {
      threefry2x32(key_schedule(s0, s1), 0u, 0u, bits, unused);  // normal(split, [1,1,1]) :88
      split2(k0, k1, n0, n1, t0, t1);                            // key, split = split(state.rng) of the NEXT step :79
    }
    k0 = n0; k1 = n1; s0 = t0; s1 = t1;
    return bits;
  }
};
template <bool FAST> using RngPipeFor = RngPipeT<FAST ? NQ_RNG_INJ_FAST : NQ_RNG_INJ_STRICT>;
using RngPipe = RngPipeT<0>;   // kernels off the hot path

// One apply_fn (barrier_crossing.py:77-92) given the landscape already evaluated at x and this step's
// standard-normal draw.  Returns dR; updates x; writes z (standardised noise) and, if WANT_LP, lp.
// HALF of the mean displacement: mean = F * nu * dt (:67-69) = 2 m2
template <int POT, bool FAST>
__device__ __forceinline__ float half_mean(const LangevinConsts& c, const Landscape<POT, FAST>& land, float x, float r) {
  // -F / 2 = (dEm/dx)/2 + (k_s/2)(x - r); the reference forms dEs = (k_s/2) * (2 (x - r)), see half_dEm
  const float es = fmul<FAST>(c.hk, fsub<FAST>(x, r));
  const float q = POT == NQ_POT_SPRING ? es : fadd<FAST>(land.half_dEm(c), es);
  return FAST ? -q * c.nudt : __fmul_rn(__fmul_rn(-q, c.nu), c.dt);
}
// dR = sampled * scale + loc (:88) in one rounding of xi sigma + 2 m2
template <bool FAST>
__device__ __forceinline__ float kick(const LangevinConsts& c, float m2, float xi) {
  return FAST ? fmaf(2.0f, m2, xi * c.sigma) : __fmaf_rn(2.0f, m2, __fmul_rn(xi, c.sigma));
}

template <int POT, bool FAST, bool PERIODIC, bool WANT_LP>
__device__ __forceinline__ float brownian_apply_xi(const LangevinConsts& c, const Landscape<POT, FAST>& land, float& x,
                                                   float xi, float r, float& z, float& lp) {
  const float dR = kick<FAST>(c, half_mean<POT, FAST>(c, land, x, r), xi);
  // Normal.log_prob's z = x/scale - loc/scale equals xi up to one rounding of dR: using xi changes
  // each log-prob by < 1e-6 absolute (tolerance 1e-5 relative of ~9) and is 7 instructions cheaper
  // than two correctly rounded quotients; the position and work paths stay in reference order.
  z = xi;
  if (WANT_LP) lp = fsub<FAST>(fmul<FAST>(-0.5f, fmul<FAST>(z, z)), c.log_norm);
  x = fadd<FAST>(x, dR);
  if (PERIODIC) x = py_mod(x, c.box);
  return dR;
}

template <int POT, bool FAST, bool PERIODIC, bool WANT_LP>
__device__ __forceinline__ float brownian_apply(const LangevinConsts& c, const Landscape<POT, FAST>& land, float& x,
                                                uint32_t bits, float r, float& z, float& lp) {
  return brownian_apply_xi<POT, FAST, PERIODIC, WANT_LP>(c, land, x, normal_from_bits<FAST>(bits), r, z, lp);
}

// run-time shift kind -> the specialised step (equilibration, stationary trap, replay: not the hot loop)
template <int POT, bool FAST>
__device__ __forceinline__ float brownian_apply_any(const LangevinConsts& c, const Landscape<POT, FAST>& land, float& x,
                                                    float xi, float r, float& z, float& lp) {
  return c.shift_kind == NQ_SHIFT_PERIODIC ? brownian_apply_xi<POT, FAST, true, true>(c, land, x, xi, r, z, lp)
                                           : brownian_apply_xi<POT, FAST, false, true>(c, land, x, xi, r, z, lp);
}

// ------------------------------------------------------------------------------ FAST-math step
// The same step with the algebra refactored for instruction count: constants
// pre-multiplied on the host (log2 e into the exponents, nu*dt/beta into the force), MUFU
// ex2/lg2/rcp/rsq without denormal fix-ups, sqrt(2) folded into the erfinv coefficients, FMA
// everywhere.  Agrees with the STRICT step to ~1e-7 per step; not bit-faithful to the reference's
// operation order.
__device__ __forceinline__ float ex2_approx(float a) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float lg2_approx(float a) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float rsq_approx(float a) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

// sqrt(2) * erfinv(u) with Giles' polynomials (coefficients pre-scaled by sqrt(2))
__device__ __forceinline__ float normal_from_bits_fast(uint32_t bits) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u);        // [1, 2)
  // 2(f-1) + nextafter(-1,0) in one exact FMA plus one rounding: bit-identical to jax's uniform(-1,1)
  // (2f - 3 is exact; adding 2^-24 rounds the same real number jax rounds), and never below -1 + 2^-24
  const float u = fmaf(2.0f, f, -3.0f) + 5.9604644775390625e-8f;
  const float lg = lg2_approx(fmaf(-u, u, 1.0f));                     // log2(1 - u^2)
  constexpr float S2 = 1.41421356237309504880f;
  float p;
  if (lg > -7.2134752044448170f) {                                    // w = -ln(1-u^2) < 5
    const float w = fmaf(lg, -0.69314718055994531f, -2.5f);
    p = S2 * 2.81022636e-08f;
    p = fmaf(p, w, S2 * 3.43273939e-07f);
    p = fmaf(p, w, S2 * -3.5233877e-06f);
    p = fmaf(p, w, S2 * -4.39150654e-06f);
    p = fmaf(p, w, S2 * 0.00021858087f);
    p = fmaf(p, w, S2 * -0.00125372503f);
    p = fmaf(p, w, S2 * -0.00417768164f);
    p = fmaf(p, w, S2 * 0.246640727f);
    p = fmaf(p, w, S2 * 1.50140941f);
  } else {
    const float w0 = lg * -0.69314718055994531f;
    const float w = fmaf(w0, rsq_approx(w0), -3.0f);                  // sqrt(w) - 3
    p = S2 * -0.000200214257f;
    p = fmaf(p, w, S2 * 0.000100950558f);
    p = fmaf(p, w, S2 * 0.00134934322f);
    p = fmaf(p, w, S2 * -0.00367342844f);
    p = fmaf(p, w, S2 * 0.00573950773f);
    p = fmaf(p, w, S2 * -0.0076224613f);
    p = fmaf(p, w, S2 * 0.00943887047f);
    p = fmaf(p, w, S2 * 1.00167406f);
    p = fmaf(p, w, S2 * 2.83297682f);
  }
  return p * u;
}

// mean displacement F*nu*dt at x for trap centre r
template <int POT>
__device__ __forceinline__ float fast_mean(const LangevinConsts& c, float x, float r) {
  const float u = x - r;
  if (POT == NQ_POT_SPRING) return -c.f_cu * u;
  const float ul = POT == NQ_POT_BIOMOLECULE ? x + c.xm : x;
  const float ur = POT == NQ_POT_BIOMOLECULE ? x - c.xm : x - c.two_xm;
#if NQ_FAST_EXP == 1
  const float A = expf(c.cl * (ul * ul));
  const float B = expf(-fmaf(c.cr, ur * ur, c.bde));
#else
  const float A = ex2_approx(c.f_cl2 * (ul * ul));                    // exp(cl ul^2)
  const float B = ex2_approx(fmaf(c.f_ncr2, ur * ur, c.f_nbde2));     // exp(-(cr ur^2 + beta dE))
#endif
  // -dEm/dx * nu dt = (nu dt / beta) (2 cl A ul - 2 cr B ur) / (A + B)
  const float num = fmaf(c.f_clm, A * ul, -c.f_crm * (B * ur));
#if NQ_FAST_RCP == 1
  return fmaf(-c.f_cu, u, __fdiv_rn(num, A + B));
#else
  return fmaf(-c.f_cu, u, num * rcp_approx(A + B));
#endif
}

__device__ __forceinline__ float fast_normal(uint32_t bits) {
#if NQ_FAST_NORMAL == 1
  return normal_from_bits<false>(bits);
#else
  return normal_from_bits_fast(bits);
#endif
}

// PERIODIC: 1 / 0 = the shift kind known at compile time (the hot loops), -1 = read it from the constants
template <int POT, int PERIODIC = -1>
__device__ __forceinline__ float fast_step_z(const LangevinConsts& c, float& x, float z, float r) {
  const float mean = fast_mean<POT>(c, x, r);
  const float dR = fmaf(z, c.sigma, mean);
  x += dR;
  if (PERIODIC < 0 ? c.shift_kind == NQ_SHIFT_PERIODIC : PERIODIC == 1) x = py_mod(x, c.box);
  return dR;
}

template <int POT, int PERIODIC = -1>
__device__ __forceinline__ float fast_step(const LangevinConsts& c, float& x, uint32_t bits, float r, float& z) {
  const float mean = fast_mean<POT>(c, x, r);   // (this statement order gives the better ptxas schedule)
  z = fast_normal(bits);
  const float dR = fmaf(z, c.sigma, mean);
  x += dR;
  if (PERIODIC < 0 ? c.shift_kind == NQ_SHIFT_PERIODIC : PERIODIC == 1) x = py_mod(x, c.box);
  return dR;
}

// ------------------------------------------------------------------------------ main kernel
struct LangevinIO {
  const float* r_table;  // [S+1]
  const float* ztable;   // [2^23] normal draw of every 23-bit pattern (library-owned, see normal_table()); STRICT only
  const float* phi;      // [S-1][D] or null
  const float* x0;       // [n]
  const uint32_t* seeds; // [n][2]
  float* work;           // [n]
  float* logp;           // [n]
  float* estimator;      // [n] or null
  float* x_final;        // [n] or null
  float* positions;      // time-major [n_rec][n] or null
  long long pos_stride;
  float* step_works;     // [S+1][n] or null
  float* step_logps;     // [S][n] or null
  float* ckpt_x;         // [n_ckpt][n] or null
  uint32_t* ckpt_rng;    // [n_ckpt][4][n] or null
  long long ckpt_every;
  double* partials;      // [gridDim.x][NV] block partial sums
  long long n;
  int D;
};

constexpr int pad4(int D) { return (D + 3) / 4 * 4; }

// NV layout of a partial row: [0..7] moments, [8 .. 8+DPS) logP grads, [8+DPS .. 8+2DPS) reparam grads
// Steps [k_begin, k_end] of the 64 trajectories of group `block`.  first: initialise from x0 / seeds
// (+ equilibration); otherwise restore the carry from `state`.  last: write the outputs and the
// group's partial sums; otherwise save the carry.  state layout: [word][npad] (coalesced).
constexpr int state_words(int D) { return 9 + 2 * (D > 0 ? D : 1) + (kDebugChecks ? 1 : 0); }
// debug builds: the carry ends with a checksum of (position, key chain, chunk index) -- a chunk that started before
// its predecessor's carry was complete and visible (an ordering bug of the FIFO hand-over) trips the assert
__device__ __forceinline__ uint32_t carry_checksum(float x, uint32_t k0, uint32_t k1, uint32_t s0, uint32_t s1, long long next_k) {
  return (__float_as_uint(x) * 0x9E3779B1u) ^ (k0 + 0x85EBCA6Bu * k1) ^ (s0 * 0xC2B2AE35u + s1) ^ (uint32_t)next_k;
}

// The steps of one staged tile (r_s / phi_s / tab_s hold steps t0 .. t0 + len - 1).  Per step: work increment at
// the position before the kick, the kick, log-prob and gradient moments.  The log-prob sum is carried as
// sum(z^2) (one FMA per step) and turned into sum(-z^2/2 - log_norm) per tile in binary64 -- at least as
// accurate as summing the reference's rounded per-step terms, which RECORD emits on the side.
template <int POT, bool FAST, int D>
struct StepCtx {
  const LangevinConsts& c;
  const LangevinIO& io;
  float& x;
  RngPipeFor<FAST>& rng;
  Landscape<POT, FAST>& land;
  float& z;
  float& lp;
  float* gA;
  float* gB;
  float& Wt;
  float& LPt;
  long long gid;
  bool live;
};

template <int POT, bool FAST, int D, bool RECORD, bool PERIODIC>
__device__ __forceinline__ void tile_steps(StepCtx<POT, FAST, D>& s, long long t0, int len, uint32_t r_addr,
                                           uint32_t phi_addr, uint32_t tab_addr) {
  constexpr bool GRAD = D > 0;
  constexpr int DPS = GRAD ? pad4(D) : 4;
  constexpr bool TIP = tab_in_phi(D) && (FAST || kAlgWork);
  const LangevinConsts& c = s.c;
  const LangevinIO& io = s.io;
  float x = s.x, z = s.z, lp = s.lp, Wt = s.Wt, LPt = s.LPt;
  // packed accumulator pairs (j, j + 1); the phi rows are zero-padded, so an odd D's spare half stays zero
  constexpr int NP = GRAD ? (D + 1) / 2 : 1;
  f32x2 gA2[NP], gB2[NP];
  if (GRAD && kPackedGrad && !FAST) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      gA2[j] = f2_pack(s.gA[2 * j], 2 * j + 1 < D ? s.gA[2 * j + 1 < D ? 2 * j + 1 : 0] : 0.f);
      gB2[j] = f2_pack(s.gB[2 * j], 2 * j + 1 < D ? s.gB[2 * j + 1 < D ? 2 * j + 1 : 0] : 0.f);
    }
  }
#pragma unroll kUnroll
  for (int i = 0; i < len; ++i, r_addr += 4, phi_addr += 4 * DPS, tab_addr += 16) {
    if (RECORD && io.ckpt_x) {  // checkpoint BEFORE step k = t0+i when (k-1) % ckpt_every == 0
      const long long k = t0 + i;
      if ((k - 1) % io.ckpt_every == 0 && s.live) {
        const long long q = (k - 1) / io.ckpt_every;
        io.ckpt_x[q * io.n + s.gid] = x;
        uint32_t* dst = io.ckpt_rng + q * 4 * io.n + s.gid;
        dst[0] = s.rng.k0; dst[io.n] = s.rng.k1; dst[2 * io.n] = s.rng.s0; dst[3 * io.n] = s.rng.s1;
      }
    }
    float dW, dR;
    const float x_before = x;
    float4 lastq = make_float4(0.f, 0.f, 0.f, 0.f);   // TIP: the row's last quad = (phi_{D-1}, r_k, a_k, b_k)
    if (TIP) lastq = lds_f32x4(phi_addr + 4 * (DPS - 4));
    if (FAST) {
      const float4 tab = TIP ? make_float4(lastq.y, lastq.z, lastq.w, 0.f) : lds_f32x4(tab_addr);
      dW = fmaf(tab.y, x, tab.z);
      if (kNormalTableFast) {   // the exact jax draw from the table instead of the MUFU approximation of it
        z = __ldg(io.ztable + (s.rng.next() >> 9));
        dR = fast_step_z<POT, PERIODIC ? 1 : 0>(c, x, z, tab.x);
      } else {
        dR = fast_step<POT, PERIODIC ? 1 : 0>(c, x, s.rng.next(), tab.x, z);
      }
      if (RECORD) lp = fmaf(-0.5f * z, z, -c.log_norm);
    } else {
      // One straight-line block: key chain, central-range normal draw, landscape on the main paths, work
      // increment, mean -- no branch, so ptxas interleaves the three threefry blocks with the dependent
      // floating-point chains.  The two rare cases -- a draw in erf_inv's tail range (0.3 % of the lanes) and
      // A + B outside the main-path window (never, for physical parameters) -- are patched afterwards.
      float r_prev, r_cur, wa, wb;
      if (kAlgWork) {
        const float4 tab = TIP ? make_float4(lastq.y, lastq.z, lastq.w, 0.f) : lds_f32x4(tab_addr);
        r_cur = tab.x; wa = tab.y; wb = tab.z;
      } else {
        r_prev = lds_f32(r_addr);
        r_cur = lds_f32(r_addr + 4);
      }
      float u = 0.f, w = 0.f;
      if (kNormalTable) {
        // jax.random.normal reads only the top 23 bits of the word (Appendix A.5), so the whole draw -- uniform,
        // log1p, both erf_inv ranges, sqrt(2) -- is a function of 2^23 patterns: one load from the 32 MB table the
        // library builds once per device with normal_from_bits<false> itself (L2-resident on B200).
        // Same bits as computing it; ~36 issue slots and the tail-range branch fewer per step.
        z = __ldg(io.ztable + (s.rng.next() >> 9));
      } else {
        u = uniform_pm1_from_bits(s.rng.next());
        w = erfinv_w<false>(u, c.l1p_c9);
        z = __fmul_rn(1.41421356237309504880f, __fmul_rn(erfinv_poly_central<false>(w), u));
        asm volatile("" : "+f"(z));   // keep the central polynomial in the straight-line block (the compiler would sink it into the w < 5 arm)
      }
      s.land.template eval<false>(c, x);
      dW = kAlgWork ? __fmaf_rn(wa, x, wb) : work_increment<POT, FAST>(c, s.land, x, r_prev, r_cur);
      float m2 = half_mean<POT, FAST>(c, s.land, x, r_cur);
      // ONE predicate for the rare cases: A + B <= 2 always (two exponentials of non-positive arguments), so only
      // the lower edge of the main-path window can be missed and a float compare tests it (NaN -> patched)
      // (a warp vote in front of this branch does not remove its reconvergence barrier pair: tried, two instructions more)
      if (!((kNormalTable || w < 5.0f) && s.land.in_window_low(c))) {
        if (!kNormalTable && !(w < 5.0f)) z = __fmul_rn(1.41421356237309504880f, __fmul_rn(erfinv_poly_tail<false>(w), u));
        // with the table the predicate is the window test itself: no second test (where guard_lo is +inf every
        // step comes here, and the library division / logarithm are right everywhere)
        if (kNormalTable || !s.land.in_window(c)) {
          s.land.fix_range(c);
          if (!kAlgWork) dW = work_increment<POT, FAST>(c, s.land, x, r_prev, r_cur);
          m2 = half_mean<POT, FAST>(c, s.land, x, r_cur);
        }
      }
      dR = kick<FAST>(c, m2, z);
      if (RECORD) lp = __fsub_rn(__fmul_rn(-0.5f, __fmul_rn(z, z)), c.log_norm);
      x = __fadd_rn(x, dR);
      if (PERIODIC) x = py_mod(x, c.box);
    }
    Wt += dW;
    LPt = fmaf(z, z, LPt);  // sum of z^2; turned into the log-prob sum at the end of the tile
    if (GRAD) {
      // d logp_k / d r_k = z * ca and d W / d r_k = k_s * dR: the constant factors are applied
      // once, after the trajectory
      // d W / d r_k = -k_s (x_{k-1} - r_k) + k_s (x_k - r_k) = k_s (x_k - x_{k-1}); that is k_s dR unless a
      // periodic shift wrapped the position
      const float a = z, b = PERIODIC ? x - x_before : dR;
      if (kPackedGrad && !FAST) {   // two accumulators per FFMA2 (same roundings as 2 D scalar FFMAs)
        const f32x2 a2 = f2_pack(a, a), b2 = f2_pack(b, b);
#pragma unroll
        for (int q = 0; q < DPS / 4; ++q) {
          const float4 p = (TIP && q == DPS / 4 - 1) ? lastq : lds_f32x4(phi_addr + 16 * q);
          if (4 * q < D) {
            const f32x2 p01 = f2_pack(p.x, p.y);
            gA2[2 * q] = f2_fma(a2, p01, gA2[2 * q]);
            gB2[2 * q] = f2_fma(b2, p01, gB2[2 * q]);
          }
          if (4 * q + 2 < D) {
            const f32x2 p23 = f2_pack(p.z, p.w);
            gA2[2 * q + 1] = f2_fma(a2, p23, gA2[2 * q + 1]);
            gB2[2 * q + 1] = f2_fma(b2, p23, gB2[2 * q + 1]);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < DPS / 4; ++q) {
          const float4 p = (TIP && q == DPS / 4 - 1) ? lastq : lds_f32x4(phi_addr + 16 * q);
          const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (4 * q + e < D) {
              s.gA[4 * q + e] = fmaf(a, pv[e], s.gA[4 * q + e]);
              s.gB[4 * q + e] = fmaf(b, pv[e], s.gB[4 * q + e]);
            }
          }
        }
      }
    }
    if (RECORD && s.live) {
      const long long k = t0 + i;
      if (io.step_works) io.step_works[k * io.n + s.gid] = dW;
      if (io.step_logps) io.step_logps[(k - 1) * io.n + s.gid] = lp;
      if (io.positions && (k % io.pos_stride) == 0) io.positions[(k / io.pos_stride) * io.n + s.gid] = x;
    }
  }
  if (GRAD && kPackedGrad && !FAST) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float lo, hi;
      f2_unpack(gA2[j], lo, hi);
      s.gA[2 * j] = lo;
      if (2 * j + 1 < D) s.gA[2 * j + 1 < D ? 2 * j + 1 : 0] = hi;
      f2_unpack(gB2[j], lo, hi);
      s.gB[2 * j] = lo;
      if (2 * j + 1 < D) s.gB[2 * j + 1 < D ? 2 * j + 1 : 0] = hi;
    }
  }
  s.x = x; s.z = z; s.lp = lp; s.Wt = Wt; s.LPt = LPt;
}

// PIN: hold the per-step constants in registers.  ptxas otherwise re-reads them from the constant bank every
// iteration (13 LDCU per step in the STRICT loop, 3 % of its issue slots); passing them through an opaque
// instruction makes their values unknown to it, so they stay where they are.  Only where the register budget has room (persistent STRICT).
__device__ __forceinline__ void pin_constants(LangevinConsts& c) {
  // (an empty inline asm vanishes before ptxas and a shuffle of a warp-uniform value is folded away; a round
  // trip through shared memory with a volatile load is opaque -- 11 loads per chunk of >= 128 steps)
  constexpr int N = 14;
  __shared__ float pin_s[N];
  float* f[N] = {&c.hk, &c.inv_beta, &c.cl, &c.cr, &c.bde, &c.xm, &c.two_xm, &c.nu, &c.dt, &c.sigma, &c.log_norm,
                 &c.exp_k1, &c.exp_k252, &c.l1p_c9};
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) pin_s[i] = *f[i];
  }
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(pin_s);
#pragma unroll
  for (int i = 0; i < N; ++i) *f[i] = lds_f32(base + 4 * i);
  // the constant pair, re-read as a 64-bit word so that it arrives in (and keeps) an aligned register pair
  __shared__ __align__(8) f32x2 pin2_s[1];
  if (threadIdx.x == 0) pin2_s[0] = c.well_c2;
  __syncthreads();
  const uint32_t base2 = (uint32_t)__cvta_generic_to_shared(pin2_s);
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(c.well_c2) : "r"(base2));
}

template <int POT, bool FAST, int D, bool RECORD, bool PIN = false>
__device__ __forceinline__ void langevin_chunk(const LangevinConsts& c_in, const LangevinIO& io, long long block,
                                               long long k_begin, long long k_end, bool first, bool last,
                                               uint32_t* state, long long npad) {
  LangevinConsts c = c_in;
  if (PIN) pin_constants(c);
  constexpr bool GRAD = D > 0;
  constexpr int DA = GRAD ? D : 1;            // accumulators per term
  constexpr int DPS = GRAD ? pad4(D) : 4;     // padded row length in shared memory
  __shared__ float r_s[kTile + 1];
  __shared__ __align__(16) float4 tab_s[(FAST || kAlgWork) ? kTile : 1];  // FAST: (r_k, a_k, b_k, -) with dW_k = a_k x + b_k
  __shared__ __align__(16) float phi_s[GRAD ? kTile * DPS : 4];
  __shared__ double red_s[kThreads / 32][8 + 2 * DPS];

  const long long gid = block * kThreads + threadIdx.x;
  const bool live = gid < io.n;
  const long long idx = live ? gid : io.n - 1;

  float x;
  RngPipeFor<FAST> rng;
  rng.one = c.one;
  double W = 0.0, LP = 0.0;
  float gA[DA], gB[DA];
  Landscape<POT, FAST> land;
  float z, lp;
  if (first) {
    x = io.x0[idx];
    {
      // key, split = random.split(key); init(split, ...)     barrier_crossing.py:133,138
      uint32_t n0, n1, r0, r1;
      split2(io.seeds[2 * idx], io.seeds[2 * idx + 1], n0, n1, r0, r1);
      rng.prime(r0, r1);
    }
    const float r_eq = io.r_table[0];
    for (long long e = 0; e < c.Neq; ++e) {  // equilibrate :115-121
      if (FAST) {
        fast_step<POT>(c, x, rng.next(), r_eq, z);
      } else {
        land.eval(c, x);
        brownian_apply_any<POT, FAST>(c, land, x, normal_from_bits<FAST>(rng.next()), r_eq, z, lp);
      }
    }
    if (RECORD && io.positions && live) io.positions[gid] = x;      // positions[0] = x_eq  :142
    if (RECORD && io.step_works && live) io.step_works[gid] = 0.f;  // works[0] = 0        :141
#pragma unroll
    for (int j = 0; j < DA; ++j) gA[j] = gB[j] = 0.f;
  } else {
    const uint32_t* sp = state + idx;
    x = __uint_as_float(sp[0]);
    rng.k0 = sp[npad]; rng.k1 = sp[2 * npad]; rng.s0 = sp[3 * npad]; rng.s1 = sp[4 * npad];
    W = __hiloint2double((int)sp[6 * npad], (int)sp[5 * npad]);
    LP = __hiloint2double((int)sp[8 * npad], (int)sp[7 * npad]);
#pragma unroll
    for (int j = 0; j < DA; ++j) {
      gA[j] = __uint_as_float(sp[(9 + j) * npad]);
      gB[j] = __uint_as_float(sp[(9 + DA + j) * npad]);
    }
    if (kDebugChecks && live)
      NQ_DBG_ASSERT(sp[(9 + 2 * DA) * npad] == carry_checksum(x, rng.k0, rng.k1, rng.s0, rng.s1, k_begin));
  }

  for (long long t0 = k_begin; t0 <= k_end; t0 += kTile) {
    const int len = (int)min((long long)kTile, k_end - t0 + 1);
    __syncthreads();
    NQ_DBG_ASSERT(t0 >= 1 && t0 - 1 + len <= c.S && len >= 1 && len <= kTile);
    for (int i = threadIdx.x; i <= len; i += kThreads) r_s[i] = io.r_table[t0 - 1 + i];
    if (GRAD) {
      // phi row for step k is k-1 (k = 1..S-1); step S (fixed end point) contributes nothing
      for (int i = threadIdx.x; i < len * DPS; i += kThreads) {
        const int row = i / DPS, col = i - row * DPS;
        const long long k = t0 + row;
        phi_s[i] = (col < io.D && k <= c.S - 1) ? io.phi[(k - 1) * io.D + col] : 0.f;  // io.D <= D (template)
      }
    }
    if (FAST || kAlgWork) {
      // dW_k = (k_s/2)[(x-r_k)^2 - (x-r_{k-1})^2] = a_k x + b_k,  a_k = k_s (r_{k-1} - r_k),  b_k = -a_k (r_{k-1} + r_k)/2
      if (tab_in_phi(D)) __syncthreads();   // the padding slots of the rows written above
      for (int i = threadIdx.x; i < len; i += kThreads) {
        const float r0 = io.r_table[t0 - 1 + i], r1 = io.r_table[t0 + i];
        const float a = 2.0f * c.hk * (r0 - r1), b = -0.5f * a * (r0 + r1);
        if (tab_in_phi(D)) {
          phi_s[i * DPS + DPS - 3] = r1; phi_s[i * DPS + DPS - 2] = a; phi_s[i * DPS + DPS - 1] = b;
        } else {
          tab_s[i] = make_float4(r1, a, b, 0.f);
        }
      }
    }
    __syncthreads();
    float Wt = 0.f, LPt = 0.f;
    StepCtx<POT, FAST, D> ctx{c, io, x, rng, land, z, lp, gA, gB, Wt, LPt, gid, live};
    const uint32_t r_addr = (uint32_t)__cvta_generic_to_shared(r_s);
    const uint32_t phi_addr = (uint32_t)__cvta_generic_to_shared(phi_s);
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab_s);
    // the shift kind is uniform over the launch: one branch per tile selects the loop specialised for it
    if (c.shift_kind == NQ_SHIFT_PERIODIC) tile_steps<POT, FAST, D, RECORD, true>(ctx, t0, len, r_addr, phi_addr, tab_addr);
    else tile_steps<POT, FAST, D, RECORD, false>(ctx, t0, len, r_addr, phi_addr, tab_addr);
    W += (double)Wt;   // two-level accumulation: f32 within a tile, binary64 across tiles
    LP += -0.5 * (double)LPt - (double)len * (double)c.log_norm;
  }
  if (!last) {
    if (live) {
      uint32_t* sp = state + gid;
      sp[0] = __float_as_uint(x);
      sp[npad] = rng.k0; sp[2 * npad] = rng.k1; sp[3 * npad] = rng.s0; sp[4 * npad] = rng.s1;
      sp[5 * npad] = (uint32_t)__double2loint(W); sp[6 * npad] = (uint32_t)__double2hiint(W);
      sp[7 * npad] = (uint32_t)__double2loint(LP); sp[8 * npad] = (uint32_t)__double2hiint(LP);
#pragma unroll
      for (int j = 0; j < DA; ++j) {
        sp[(9 + j) * npad] = __float_as_uint(gA[j]);
        sp[(9 + DA + j) * npad] = __float_as_uint(gB[j]);
      }
      if (kDebugChecks) sp[(9 + 2 * DA) * npad] = carry_checksum(x, rng.k0, rng.k1, rng.s0, rng.s1, k_end + 1);
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < DA; ++j) {
    gA[j] *= c.ca;
    gB[j] *= c.k_s;
  }

  const float Wf = (float)W, LPf = (float)LP;
  const float est = __fadd_rn(__fmul_rn(LPf, Wf), Wf);  // tot_log_prob * sg(total_work) + total_work  :212
  if (live) {
    io.work[gid] = Wf;
    io.logp[gid] = LPf;
    if (io.estimator) io.estimator[gid] = est;
    if (io.x_final) io.x_final[gid] = x;
  }

  // ---- block reduction of moments and gradient sums (binary64, fixed order => deterministic)
  const double m = live ? 1.0 : 0.0;
  const double Wd = (double)Wf, Ld = (double)LPf;
  constexpr int NV = 8 + 2 * DPS;
  double v[NV];
  v[NQ_MOM_N] = m;
  v[NQ_MOM_W] = m * Wd;
  v[NQ_MOM_W2] = m * Wd * Wd;
  v[NQ_MOM_LOGP] = m * Ld;
  v[NQ_MOM_WLOGP] = m * Wd * Ld;
  v[NQ_MOM_EST] = m * (double)est;
  v[6] = 0.0;
  v[7] = 0.0;
#pragma unroll
  for (int j = 0; j < DPS; ++j) {
    v[8 + j] = (GRAD && j < D) ? m * Wd * (double)gA[j < DA ? j : 0] : 0.0;
    v[8 + DPS + j] = (GRAD && j < D) ? m * (double)gB[j < DA ? j : 0] : 0.0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const double s = warp_sum(v[q]);
    if (lane == 0) red_s[warp][q] = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < NV; q += kThreads) {  // NV = 8 + 2*DPS reaches 72 > kThreads for D > 28
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red_s[w][q];
    io.partials[block * NV + q] = s;
  }
}

template <int POT, bool FAST, int D, bool RECORD>
__global__ void __launch_bounds__(kThreads, min_blocks(D, FAST)) langevin_kernel(const LangevinConsts c, const LangevinIO io) {
  langevin_chunk<POT, FAST, D, RECORD, NQ_PIN_STATIC && !FAST && !RECORD>(c, io, blockIdx.x, 1, c.S, true, true, nullptr, 0);
}

// ------------------------------------------------------------------------------ small ensembles
// Below ~10^4 trajectories a warp is alone on its SM sub-partition and the main kernel is
// latency-bound: one step costs ~1050 cycles (STRICT) because the three threefry blocks + normal draw
// and the force / work chain are issued back to back from ONE in-order instruction stream, although
// the random stream does not depend on the positions.  Here every group of 32 trajectories gets
// THREE warps in a pipeline:
//   key warp    the serial key chain  key, sub = split(key)  (:79) -- two independent threefry blocks per step,
//               the only step-to-step dependence of the random stream (~390 cycles per step: the floor);
//   draw warp   bits = threefry(sub, (0, 0)); xi = sqrt(2) erf_inv(...) (:88) -- no dependence between steps, so it
//               works on four steps at a time (instruction-level parallelism instead of a second draw warp);
//   consumer    integration, work, log-prob and gradient moments.
// They meet in two shared-memory rings of two 32-step halves (sub-keys; draws), each guarded by four named
// barriers (FULL / EMPTY per half).  Per-trajectory results are bit-identical to langevin_kernel (same
// functions, same order).  Round 1 shipped the two-warp form (key chain + draws in one producer warp):
// 670 cycles per step STRICT at config 1, set by the producer (scripts/latency_scan.py).
constexpr int kPairThreads = 96;  // warp 0: consumer, warp 1: draw, warp 2: key chain
constexpr int kBarCount = 64;     // every named barrier joins two warps
constexpr int kRingHalf = 32;     // steps per ring half

// STRICT normal draws of four independent words at once: the central erf_inv polynomials in straight-line code (so
// that the four threefry blocks and the four polynomial chains interleave), tail-range lanes patched afterwards
__device__ __forceinline__ void normal4_strict(const uint32_t (&bits)[4], float (&z)[4]) {
  float u[4], w[4];
  bool tail = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    u[j] = uniform_pm1_from_bits(bits[j]);
    w[j] = erfinv_w<false>(u[j]);
    z[j] = __fmul_rn(1.41421356237309504880f, __fmul_rn(erfinv_poly_central<false>(w[j]), u[j]));
    asm volatile("" : "+f"(z[j]));
    tail |= !(w[j] < 5.0f);
  }
  if (tail) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (!(w[j] < 5.0f)) z[j] = __fmul_rn(1.41421356237309504880f, __fmul_rn(erfinv_poly_tail<false>(w[j]), u[j]));
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int POT, bool FAST, int D>
__global__ void __launch_bounds__(kPairThreads) langevin_pair_kernel(const LangevinConsts c, const LangevinIO io) {
  constexpr bool GRAD = D > 0;
  constexpr int DA = GRAD ? D : 1;
  constexpr int DPS = GRAD ? pad4(D) : 4;
  constexpr int kFull = 1, kEmpty = 3;   // draw ring: barrier ids kFull + h, kEmpty + h (0 is __syncthreads)
  constexpr int kFullK = 5, kEmptyK = 7; // sub-key ring
  __shared__ float ring_s[2][kRingHalf][32];
  __shared__ uint32_t key_ring_s[2][kRingHalf][2][32];
  __shared__ long long ring_tag_s[kDebugChecks ? 2 * kRingHalf : 1];   // debug: global step index held by each ring slot
  __shared__ float r_s[kTile + 1];
  __shared__ __align__(16) float4 tab_s[(FAST || kAlgWork) ? kTile : 1];
  __shared__ __align__(16) float phi_s[GRAD ? kTile * DPS : 4];

  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;     // 0 consumer, 1 draw, 2 key chain
  const long long gid = (long long)blockIdx.x * 32 + lane;
  const bool live = gid < io.n;
  const long long idx = live ? gid : io.n - 1;
  const long long total = c.Neq + c.S;
  const long long n_chunks = (total + kRingHalf - 1) / kRingHalf;

  if (role == 2) {   // ---- key chain
    uint32_t k0, k1, s0, s1;
    {
      uint32_t n0, n1, r0, r1;
      split2(io.seeds[2 * idx], io.seeds[2 * idx + 1], n0, n1, r0, r1);   // barrier_crossing.py:133,138
      split2(r0, r1, k0, k1, s0, s1);                                     // the split of the first step (:79)
    }
    for (long long ch = 0; ch < n_chunks; ++ch) {
      const int h = (int)(ch & 1);
      if (ch >= 2) named_bar_sync(kEmptyK + h, kBarCount);   // the draw warp has read this half
      const int len = (int)min((long long)kRingHalf, total - ch * kRingHalf);
      for (int i = 0; i < len; ++i) {
        key_ring_s[h][i][0][lane] = s0;
        key_ring_s[h][i][1][lane] = s1;
        uint32_t n0, n1;
        split2(k0, k1, n0, n1, s0, s1);   // key, split = split(state.rng) of the NEXT step
        k0 = n0; k1 = n1;
      }
      __threadfence_block();
      named_bar_arrive(kFullK + h, kBarCount);
    }
    return;
  }
  if (role == 1) {   // ---- draws
    for (long long ch = 0; ch < n_chunks; ++ch) {
      const int h = (int)(ch & 1);
      named_bar_sync(kFullK + h, kBarCount);                 // sub-keys of this half are there
      if (ch >= 2) named_bar_sync(kEmpty + h, kBarCount);    // the consumer is done with this half of the draws
      const int len = (int)min((long long)kRingHalf, total - ch * kRingHalf);
      int i = 0;
      for (; i + 4 <= len; i += 4) {
        uint32_t bits[4], unused;
        float z[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          threefry2x32(key_schedule(key_ring_s[h][i + j][0][lane], key_ring_s[h][i + j][1][lane]), 0u, 0u, bits[j], unused);
        if (FAST && NQ_FAST_NORMAL != 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) z[j] = fast_normal(bits[j]);
        } else {
          normal4_strict(bits, z);   // both math modes draw the exact jax.random.normal values
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ring_s[h][i + j][lane] = z[j];
      }
      for (; i < len; ++i) {
        uint32_t bits, unused;
        threefry2x32(key_schedule(key_ring_s[h][i][0][lane], key_ring_s[h][i][1][lane]), 0u, 0u, bits, unused);
        ring_s[h][i][lane] = FAST ? fast_normal(bits) : normal_from_bits<false>(bits);
      }
      if (kDebugChecks && lane == 0)
        for (int q = 0; q < len; ++q) ring_tag_s[h * kRingHalf + q] = ch * kRingHalf + q;
      __threadfence_block();
      if (ch + 2 < n_chunks) named_bar_arrive(kEmptyK + h, kBarCount);   // the key warp may refill this half
      named_bar_arrive(kFull + h, kBarCount);
    }
    return;
  }

  // ---- consumer
  long long chunk = 0;
  int pos = 0;
  auto next_z = [&]() -> float {
    const int h = (int)(chunk & 1);
    if (pos == 0) {
      // release the half finished last (its values have all been used) once a later chunk needs it
      if (chunk >= 1 && chunk + 1 < n_chunks) named_bar_arrive(kEmpty + (h ^ 1), kBarCount);
      named_bar_sync(kFull + h, kBarCount);
    }
    const float z = ring_s[h][pos][lane];
    // debug: the slot must hold exactly the draw of this step (not yet overwritten, already produced)
    if (kDebugChecks) NQ_DBG_ASSERT(((volatile long long*)ring_tag_s)[h * kRingHalf + pos] == chunk * kRingHalf + pos);
    if (++pos == kRingHalf) {
      pos = 0;
      ++chunk;
    }
    return z;
  };

  float x = io.x0[idx];
  double W = 0.0, LP = 0.0;
  float gA[DA], gB[DA];
#pragma unroll
  for (int j = 0; j < DA; ++j) gA[j] = gB[j] = 0.f;
  Landscape<POT, FAST> land;
  float z, lp;
  {
    const float r_eq = io.r_table[0];
    for (long long e = 0; e < c.Neq; ++e) {  // equilibrate :115-121
      if (FAST) {
        fast_step_z<POT>(c, x, next_z(), r_eq);
      } else {
        land.eval(c, x);
        brownian_apply_any<POT, FAST>(c, land, x, next_z(), r_eq, z, lp);
      }
    }
  }
  // optional trajectory record (time-major [S / stride + 1][n], row 0 = the equilibrated start :142): the
  // consumer is not the slower stage, so the stores are free here
  float* pos_out = io.positions ? io.positions + gid : nullptr;
  if (pos_out && live) pos_out[0] = x;
  for (long long t0 = 1; t0 <= c.S; t0 += kTile) {
    const int len = (int)min((long long)kTile, c.S - t0 + 1);
    __syncwarp();
    for (int i = lane; i <= len; i += 32) r_s[i] = io.r_table[t0 - 1 + i];
    if (FAST || kAlgWork) {
      for (int i = lane; i < len; i += 32) {
        const float r0 = io.r_table[t0 - 1 + i], r1 = io.r_table[t0 + i];
        const float a = 2.0f * c.hk * (r0 - r1);
        tab_s[i] = make_float4(r1, a, -0.5f * a * (r0 + r1), 0.f);
      }
    }
    if (GRAD) {
      for (int i = lane; i < len * DPS; i += 32) {
        const int row = i / DPS, col = i - row * DPS;
        const long long k = t0 + row;
        phi_s[i] = (col < io.D && k <= c.S - 1) ? io.phi[(k - 1) * io.D + col] : 0.f;
      }
    }
    __syncwarp();
    float Wt = 0.f, LPt = 0.f;
    uint32_t r_addr = (uint32_t)__cvta_generic_to_shared(r_s);
    uint32_t phi_addr = (uint32_t)__cvta_generic_to_shared(phi_s);
    uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab_s);
    for (int i = 0; i < len; ++i, r_addr += 4, phi_addr += 4 * DPS, tab_addr += 16) {
      float dW, dR;
      const float x_before = x;
      if (FAST) {
        const float4 tab = lds_f32x4(tab_addr);
        dW = fmaf(tab.y, x, tab.z);
        z = next_z();
        dR = fast_step_z<POT>(c, x, z, tab.x);
      } else {
        float r_cur;
        land.eval(c, x);
        if (kAlgWork) {
          const float4 tab = lds_f32x4(tab_addr);
          r_cur = tab.x;
          dW = __fmaf_rn(tab.y, x, tab.z);
        } else {
          r_cur = lds_f32(r_addr + 4);
          dW = work_increment<POT, FAST>(c, land, x, lds_f32(r_addr), r_cur);
        }
        dR = brownian_apply_any<POT, FAST>(c, land, x, next_z(), r_cur, z, lp);
      }
      Wt += dW;
      LPt = fmaf(z, z, LPt);   // as in tile_steps
      if (pos_out && live) {
        const long long k = t0 + i;
        if (k % io.pos_stride == 0) pos_out[(k / io.pos_stride) * io.n] = x;
      }
      if (GRAD) {
        const float a = z, b = c.shift_kind == NQ_SHIFT_PERIODIC ? x - x_before : dR;
#pragma unroll
        for (int q = 0; q < DPS / 4; ++q) {
          const float4 p = lds_f32x4(phi_addr + 16 * q);
          const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (4 * q + e < D) {
              gA[4 * q + e] = fmaf(a, pv[e], gA[4 * q + e]);
              gB[4 * q + e] = fmaf(b, pv[e], gB[4 * q + e]);
            }
          }
        }
      }
    }
    W += (double)Wt;
    LP += -0.5 * (double)LPt - (double)len * (double)c.log_norm;
  }
#pragma unroll
  for (int j = 0; j < DA; ++j) {
    gA[j] *= c.ca;
    gB[j] *= c.k_s;
  }
  const float Wf = (float)W, LPf = (float)LP;
  const float est = __fadd_rn(__fmul_rn(LPf, Wf), Wf);
  if (live) {
    io.work[gid] = Wf;
    io.logp[gid] = LPf;
    if (io.estimator) io.estimator[gid] = est;
    if (io.x_final) io.x_final[gid] = x;
  }
  const double m = live ? 1.0 : 0.0;
  const double Wd = (double)Wf, Ld = (double)LPf;
  constexpr int NV = 8 + 2 * DPS;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double val = 0.0;
    if (v == NQ_MOM_N) val = m;
    else if (v == NQ_MOM_W) val = m * Wd;
    else if (v == NQ_MOM_W2) val = m * Wd * Wd;
    else if (v == NQ_MOM_LOGP) val = m * Ld;
    else if (v == NQ_MOM_WLOGP) val = m * Wd * Ld;
    else if (v == NQ_MOM_EST) val = m * (double)est;
    else if (v >= 8 && GRAD) {
      const int j = (v - 8) % DPS, which = (v - 8) / DPS;
      if (j < D) val = which == 0 ? m * Wd * (double)gA[j < DA ? j : 0] : m * (double)gB[j < DA ? j : 0];
    }
    const double s = warp_sum(val);
    if (lane == 0) io.partials[(long long)blockIdx.x * NV + v] = s;
  }
}

// Persistent variant with a FIFO ready-queue of trajectory groups ("chains").  With 1e5
// trajectories there are 3125 warps for 592 SM sub-partitions: a static launch leaves some with 6
// warps and some with 5 for the whole run (makespan 6 units for an average of 5.28).  Here fewer
// CTAs than chains are resident (the same number on every SM); each CTA pops a chain, advances it
// by one chunk of steps, saves the carry and pushes the chain back, so chains rotate over all SMs
// and every sub-partition carries the same load all the time.
//   ring[t]  : chain id served by pop ticket t (-1 until published); tickets 0..groups-1 are
//              pre-filled, every completed non-final chunk publishes one more entry;
//   n_items  = groups * n_chunks tickets in total, so a pop on an unpublished entry always has a
//              running chunk that will publish it: waits terminate.
struct PersistentCtl {
  unsigned int* head;        // pop tickets handed out
  unsigned int* tail;        // push tickets handed out (starts at groups)
  int* ring;                 // [groups * n_chunks]
  int* progress;             // [groups] chunks completed
  unsigned int* status;      // 0 = ok; set to 1 when a bounded wait ran out (the launch is then abandoned)
  unsigned int max_polls;    // bound of a wait on the ring, in polls of 100 ns; 0 = unbounded
  uint32_t* state;           // [state_words][npad]
  long long groups, npad, chunk, n_chunks;
};

static __global__ void persistent_init_kernel(PersistentCtl ctl) {
  const long long n_items = ctl.groups * ctl.n_chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (long long)gridDim.x * blockDim.x) {
    ctl.ring[i] = i < ctl.groups ? (int)i : -1;
    if (i < ctl.groups) ctl.progress[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *ctl.head = 0u;
    *ctl.tail = (unsigned int)ctl.groups;
    *ctl.status = 0u;
  }
}

template <int POT, bool FAST, int D>
__global__ void __launch_bounds__(kThreads, persistent_min_blocks(D, FAST)) langevin_persistent_kernel(const LangevinConsts c,
                                                                                      const LangevinIO io,
                                                                                      const PersistentCtl ctl) {
  __shared__ int chain_s, chunk_s;
  const unsigned long long n_items = (unsigned long long)ctl.groups * ctl.n_chunks;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int t = atomicAdd(ctl.head, 1u);
      int g = -2;
      if (t < n_items) {
        // A popped ticket whose entry is not published yet always has a running chunk that will publish it
        // (n_items tickets in total), provided the resident CTAs make progress together -- true for one
        // kernel owning the GPU, not guaranteed under a debugger, MPS time-slicing or pre-emption.  The wait
        // is therefore bounded (~3 s by default, NQ_WAIT_POLLS; 0 = unbounded): when it runs out the launch
        // is abandoned CLEANLY -- status word set, this CTA leaves, the others follow as their own waits run
        // out or the queue drains -- and the finalize kernel turns the status into NaN sums.  No __trap():
        // a trap would poison the whole CUDA context of the caller.
        volatile int* slot = ctl.ring + t;
        volatile unsigned int* dead = ctl.status;
        unsigned int spins = 0;
        while ((g = *slot) < 0) {
          __nanosleep(100);
          if ((ctl.max_polls && ++spins > ctl.max_polls) || *dead) {
            atomicExch(ctl.status, 1u);
            g = -2;
            break;
          }
        }
        if (g >= 0) {
          __threadfence();
          chunk_s = ctl.progress[g];
        }
      }
      chain_s = g;
    }
    __syncthreads();
    const long long g = chain_s;
    if (g < 0) return;
    const long long ch = chunk_s;
    const long long k_begin = ch * ctl.chunk + 1;
    const long long k_end = min(c.S, (ch + 1) * ctl.chunk);
    const bool last = ch == ctl.n_chunks - 1;
    langevin_chunk<POT, FAST, D, false, NQ_PIN_CONSTANTS && !FAST>(c, io, g, k_begin, k_end, ch == 0, last, ctl.state, ctl.npad);
    if (!last) {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        ctl.progress[g] = (int)ch + 1;
        __threadfence();
        const unsigned int s = atomicAdd(ctl.tail, 1u);
        atomicExch(ctl.ring + s, (int)g);
      }
    }
  }
}

// ------------------------------------------------------------------------------ replay ("adjoint")
// One thread per (trajectory, segment): regenerate the segment from its checkpoint and accumulate
//   gbar[k] += lbar_n * dlogp_{n,k}/dr_k + wbar_n * dW_n/dr_k     for every step k of the segment.
// d logp_k/d r_k = z_k * ca ; d W/d r_k = k_s (x_k - x_{k-1}) for 1 <= k <= S-1,
// d W/d r_S = -dE/dr(x_{S-1}; r_S) = k_s (x_{S-1} - r_S), d W/d r_0 = -k_s (x_0 - r_0)... the two end
// points are fixed by the schedule (barrier_crossing.py:149-150) but reported for completeness.
struct ReplayIO {
  const float* r_table;
  const float* ckpt_x;
  const uint32_t* ckpt_rng;
  const float* lbar;
  const float* wbar;
  double* gbar;  // [S+1]
  long long n, ckpt_every, n_ckpt;
};

template <int POT, bool FAST>
__global__ void __launch_bounds__(128) replay_kernel(const LangevinConsts c, const ReplayIO io) {
  extern __shared__ double acc_s[];  // [ckpt_every + 1]
  const long long q = blockIdx.y;    // segment
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gid < io.n;
  const long long idx = live ? gid : io.n - 1;
  const long long k_first = q * io.ckpt_every + 1;
  const int len = (int)min(io.ckpt_every, c.S - k_first + 1);
  for (int i = threadIdx.x; i <= len; i += blockDim.x) acc_s[i] = 0.0;
  __syncthreads();
  float x = io.ckpt_x[q * io.n + idx];
  RngPipe rng;
  {
    const uint32_t* src = io.ckpt_rng + q * 4 * io.n + idx;
    rng.k0 = src[0]; rng.k1 = src[io.n]; rng.s0 = src[2 * io.n]; rng.s1 = src[3 * io.n];
  }
  const float lb = live ? io.lbar[idx] : 0.f, wb = live ? io.wbar[idx] : 0.f;
  Landscape<POT, FAST> land;
  float z, lp;
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < len; ++i) {
    const long long k = k_first + i;
    const float r_prev = io.r_table[k - 1], r_cur = io.r_table[k];
    const float x_before = x;
    if (FAST) {
      fast_step<POT>(c, x, rng.next(), r_cur, z);
    } else {
      land.eval(c, x);
      brownian_apply_any<POT, FAST>(c, land, x, normal_from_bits<FAST>(rng.next()), r_cur, z, lp);
    }
    // dW_k = E(x_{k-1}; r_k) - E(x_{k-1}; r_{k-1}):  d/dr_k = -k_s (x_{k-1} - r_k), d/dr_{k-1} = +k_s (x_{k-1} - r_{k-1})
    double g_cur = (double)lb * (double)(z * c.ca) - (double)wb * (double)(c.k_s * (x_before - r_cur));
    double g_prev = (double)wb * (double)(c.k_s * (x_before - r_prev));
    g_cur = warp_sum(g_cur);
    g_prev = warp_sum(g_prev);
    if (lane == 0) {
      atomicAdd(&acc_s[i + 1], g_cur);
      atomicAdd(&acc_s[i], g_prev);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= len; i += blockDim.x) {
    const double v = acc_s[i];
    if (v != 0.0) atomicAdd(&io.gbar[k_first - 1 + i], v);
  }
}

// stationary trap: same integrator, constant r0, no work bookkeeping
template <int POT, bool FAST>
__global__ void __launch_bounds__(kThreads) stationary_kernel(const LangevinConsts c, float r0, const float* x0,
                                                              const uint32_t* seeds, long long n, float* x_final,
                                                              float* positions) {
  const long long gid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gid >= n) return;
  float x = x0[gid];
  RngPipe rng;
  {
    uint32_t n0, n1, r0k, r1k;
    split2(seeds[2 * gid], seeds[2 * gid + 1], n0, n1, r0k, r1k);  // :228
    rng.prime(r0k, r1k);
  }
  Landscape<POT, FAST> land;
  float z, lp;
  for (long long s = 0; s < c.S; ++s) {
    if (FAST) {
      fast_step<POT>(c, x, rng.next(), r0, z);
    } else {
      land.eval(c, x);
      brownian_apply_any<POT, FAST>(c, land, x, normal_from_bits<FAST>(rng.next()), r0, z, lp);
    }
    if (positions) positions[s * n + gid] = x;
  }
  if (x_final) x_final[gid] = x;
}

// a single apply_fn step on caller-held state (the state's rng is split in place)
template <int POT, bool FAST>
__global__ void __launch_bounds__(kThreads) apply_kernel(const LangevinConsts c, float r0, float* x_io, uint32_t* rng_io,
                                                         float* log_prob, long long n) {
  const long long gid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gid >= n) return;
  float x = x_io[gid];
  uint32_t n0, n1, s0, s1, bits, unused;
  split2(rng_io[2 * gid], rng_io[2 * gid + 1], n0, n1, s0, s1);
  threefry2x32(key_schedule(s0, s1), 0u, 0u, bits, unused);
  Landscape<POT, FAST> land;
  float z, lp;
  if (FAST) {
    fast_step<POT>(c, x, bits, r0, z);
    lp = fmaf(-0.5f * z, z, -c.log_norm);
  } else {
    land.eval(c, x);
    brownian_apply_any<POT, FAST>(c, land, x, normal_from_bits<FAST>(bits), r0, z, lp);
  }
  x_io[gid] = x;
  rng_io[2 * gid] = n0;
  rng_io[2 * gid + 1] = n1;
  if (log_prob) log_prob[gid] = lp;
}

// apply_fn on a rank-generic state: position [N][dim] (E = N * dim elements), ONE key per state (simulate.py:66-97,
// barrier_crossing.py:66-92).  dist.sample(seed=split) draws jax.random.normal(split, [1, N, dim]): element e
// consumes word e of threefry_2x32(split, iota(E)) -- counter pairing (e, e + ceil(E/2)), odd sizes padded -- so a
// thread owns a pair.  The reference's energy closures read position[0][0] only (potentials.py:27-52): element 0
// feels landscape + trap, every other element has zero force and diffuses freely.  log_prob sums
// Normal.log_prob = -((dR/sigma) - (mean/sigma))^2 / 2 - log_norm over all elements (tfp's form, exact here;
// binary64 accumulation, rounded once).  One CTA per state.
template <int POT>
__global__ void __launch_bounds__(128) apply_nd_kernel(const LangevinConsts c, float r0, float* pos, uint32_t* rng_io,
                                                       float* log_prob, long long n_states, int E) {
  __shared__ uint32_t key_s[4];
  __shared__ double red_s[4];
  const long long b = blockIdx.x;
  if (b >= n_states) return;
  float* x = pos + b * E;
  if (threadIdx.x == 0) {
    uint32_t n0, n1, s0, s1;
    split2(rng_io[2 * b], rng_io[2 * b + 1], n0, n1, s0, s1);   // key, split = random.split(state.rng)  :79
    key_s[0] = n0; key_s[1] = n1; key_s[2] = s0; key_s[3] = s1;
  }
  __syncthreads();
  const KeySchedule ks = key_schedule(key_s[2], key_s[3]);
  const int half = (E + 1) / 2;
  double acc = 0.0;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const int j = i + half;
    uint32_t y0, y1;
    threefry2x32(ks, (uint32_t)i, j < E ? (uint32_t)j : 0u, y0, y1);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int e = w ? j : i;
      if (e >= E) continue;
      const float xi = normal_from_bits<false>(w ? y1 : y0);
      const float xe = x[e];
      float mean = 0.f;
      if (e == 0) {   // force on position[0][0] in the reference's order (:60, 67-69)
        Landscape<POT, false> land;
        land.eval(c, xe);
        mean = __fmul_rn(2.0f, half_mean<POT, false>(c, land, xe, r0));
      }
      const float dR = __fadd_rn(__fmul_rn(xi, c.sigma), mean);
      const float z = __fsub_rn(__fdiv_rn(dR, c.sigma), __fdiv_rn(mean, c.sigma));
      acc += (double)__fsub_rn(__fmul_rn(-0.5f, __fmul_rn(z, z)), c.log_norm);
      float xn = __fadd_rn(xe, dR);
      if (c.shift_kind == NQ_SHIFT_PERIODIC) xn = py_mod(xn, c.box);
      x[e] = xn;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (log_prob) log_prob[b] = (float)((red_s[0] + red_s[1]) + (red_s[2] + red_s[3]));
    rng_io[2 * b] = key_s[0];
    rng_io[2 * b + 1] = key_s[1];
  }
}

// per-potential launchers (defined in langevin_pot*.cu through NQ_DEFINE_POT_LAUNCHERS)
using MainLauncher = int (*)(const LangevinConsts&, const LangevinIO&, int blocks, bool fast, bool record, cudaStream_t);
using PairLauncher = int (*)(const LangevinConsts&, const LangevinIO&, int blocks, bool fast, cudaStream_t);
using PersistentLauncher = int (*)(const LangevinConsts&, const LangevinIO&, const PersistentCtl&, int blocks, bool fast,
                                   cudaStream_t, int* max_blocks_per_sm);
using ReplayLauncher = int (*)(const LangevinConsts&, const ReplayIO&, dim3 grid, size_t smem, bool fast, cudaStream_t);
using StationaryLauncher = int (*)(const LangevinConsts&, float r0, const float*, const uint32_t*, long long, float*,
                                   float*, int blocks, bool fast, cudaStream_t);
using ApplyLauncher = int (*)(const LangevinConsts&, float r0, float*, uint32_t*, float*, long long, int blocks,
                              bool fast, cudaStream_t);
using ApplyNdLauncher = int (*)(const LangevinConsts&, float r0, float*, uint32_t*, float*, long long, int E, cudaStream_t);

#define NQ_MAIN_CASE(POT, DD)                                                                          \
  case DD:                                                                                             \
    if (fast) langevin_kernel<POT, true, DD, false><<<blocks, kThreads, 0, st>>>(c, io);               \
    else langevin_kernel<POT, false, DD, false><<<blocks, kThreads, 0, st>>>(c, io);                   \
    break;

#define NQ_PAIR_CASE(POT, DD)                                                                          \
  case DD:                                                                                             \
    if (fast) langevin_pair_kernel<POT, true, DD><<<blocks, kPairThreads, 0, st>>>(c, io);             \
    else langevin_pair_kernel<POT, false, DD><<<blocks, kPairThreads, 0, st>>>(c, io);                 \
    break;

// occ != nullptr: only report resident CTAs per SM (no launch)
#define NQ_PERSIST_CASE(POT, DD)                                                                       \
  case DD:                                                                                             \
    if (occ) {                                                                                         \
      if (fast) cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, langevin_persistent_kernel<POT, true, DD>, kThreads, 0);  \
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, langevin_persistent_kernel<POT, false, DD>, kThreads, 0);      \
    } else if (fast) langevin_persistent_kernel<POT, true, DD><<<blocks, kThreads, 0, st>>>(c, io, ctl); \
    else langevin_persistent_kernel<POT, false, DD><<<blocks, kThreads, 0, st>>>(c, io, ctl);          \
    break;

#define NQ_DEFINE_POT_LAUNCHERS(POT, NAME)                                                             \
  int NAME##_main(const LangevinConsts& c, const LangevinIO& io, int blocks, bool fast, bool record,   \
                  cudaStream_t st) {                                                                   \
    if (io.D == 0) {                                                                                   \
      if (record) {                                                                                    \
        if (fast) langevin_kernel<POT, true, 0, true><<<blocks, kThreads, 0, st>>>(c, io);             \
        else langevin_kernel<POT, false, 0, true><<<blocks, kThreads, 0, st>>>(c, io);                 \
      } else {                                                                                         \
        if (fast) langevin_kernel<POT, true, 0, false><<<blocks, kThreads, 0, st>>>(c, io);            \
        else langevin_kernel<POT, false, 0, false><<<blocks, kThreads, 0, st>>>(c, io);                \
      }                                                                                                \
      return 0;                                                                                        \
    }                                                                                                  \
    switch (io.D <= 13 && io.D > 12 ? 13 : pad4(io.D)) {                                               \
      NQ_MAIN_CASE(POT, 4) NQ_MAIN_CASE(POT, 8) NQ_MAIN_CASE(POT, 12) NQ_MAIN_CASE(POT, 13)            \
      NQ_MAIN_CASE(POT, 16) NQ_MAIN_CASE(POT, 20) NQ_MAIN_CASE(POT, 24) NQ_MAIN_CASE(POT, 28)          \
      NQ_MAIN_CASE(POT, 32)                                                                            \
      default: return -1;                                                                              \
    }                                                                                                  \
    return 0;                                                                                          \
  }                                                                                                    \
  int NAME##_pair(const LangevinConsts& c, const LangevinIO& io, int blocks, bool fast, cudaStream_t st) { \
    switch (io.D == 0 ? 0 : (io.D <= 13 && io.D > 12 ? 13 : pad4(io.D))) {                             \
      NQ_PAIR_CASE(POT, 0) NQ_PAIR_CASE(POT, 4) NQ_PAIR_CASE(POT, 8) NQ_PAIR_CASE(POT, 12)             \
      NQ_PAIR_CASE(POT, 13) NQ_PAIR_CASE(POT, 16) NQ_PAIR_CASE(POT, 20) NQ_PAIR_CASE(POT, 24)          \
      NQ_PAIR_CASE(POT, 28) NQ_PAIR_CASE(POT, 32)                                                      \
      default: return -1;                                                                              \
    }                                                                                                  \
    return 0;                                                                                          \
  }                                                                                                    \
  int NAME##_persistent(const LangevinConsts& c, const LangevinIO& io, const PersistentCtl& ctl,       \
                        int blocks, bool fast, cudaStream_t st, int* occ) {                            \
    switch (io.D == 0 ? 0 : (io.D <= 13 && io.D > 12 ? 13 : pad4(io.D))) {                             \
      NQ_PERSIST_CASE(POT, 0) NQ_PERSIST_CASE(POT, 4) NQ_PERSIST_CASE(POT, 8) NQ_PERSIST_CASE(POT, 12) \
      NQ_PERSIST_CASE(POT, 13) NQ_PERSIST_CASE(POT, 16) NQ_PERSIST_CASE(POT, 20)                       \
      NQ_PERSIST_CASE(POT, 24) NQ_PERSIST_CASE(POT, 28) NQ_PERSIST_CASE(POT, 32)                       \
      default: return -1;                                                                              \
    }                                                                                                  \
    return 0;                                                                                          \
  }                                                                                                    \
  int NAME##_replay(const LangevinConsts& c, const ReplayIO& io, dim3 grid, size_t smem, bool fast,    \
                    cudaStream_t st) {                                                                 \
    if (fast) replay_kernel<POT, true><<<grid, 128, smem, st>>>(c, io);                                \
    else replay_kernel<POT, false><<<grid, 128, smem, st>>>(c, io);                                    \
    return 0;                                                                                          \
  }                                                                                                    \
  int NAME##_stationary(const LangevinConsts& c, float r0, const float* x0, const uint32_t* seeds,     \
                        long long n, float* xf, float* pos, int blocks, bool fast, cudaStream_t st) {  \
    if (fast) stationary_kernel<POT, true><<<blocks, kThreads, 0, st>>>(c, r0, x0, seeds, n, xf, pos); \
    else stationary_kernel<POT, false><<<blocks, kThreads, 0, st>>>(c, r0, x0, seeds, n, xf, pos);     \
    return 0;                                                                                          \
  }                                                                                                    \
  int NAME##_apply(const LangevinConsts& c, float r0, float* x, uint32_t* rng, float* lp, long long n, \
                   int blocks, bool fast, cudaStream_t st) {                                           \
    if (fast) apply_kernel<POT, true><<<blocks, kThreads, 0, st>>>(c, r0, x, rng, lp, n);              \
    else apply_kernel<POT, false><<<blocks, kThreads, 0, st>>>(c, r0, x, rng, lp, n);                  \
    return 0;                                                                                          \
  }                                                                                                    \
  int NAME##_apply_nd(const LangevinConsts& c, float r0, float* x, uint32_t* rng, float* lp,           \
                      long long n, int E, cudaStream_t st) {                                           \
    apply_nd_kernel<POT><<<(unsigned)n, 128, 0, st>>>(c, r0, x, rng, lp, n, E);                        \
    return 0;                                                                                          \
  }

#define NQ_DECLARE_POT_LAUNCHERS(NAME)                                                                 \
  int NAME##_main(const LangevinConsts&, const LangevinIO&, int, bool, bool, cudaStream_t);            \
  int NAME##_pair(const LangevinConsts&, const LangevinIO&, int, bool, cudaStream_t);                  \
  int NAME##_persistent(const LangevinConsts&, const LangevinIO&, const PersistentCtl&, int, bool,    \
                        cudaStream_t, int*);                                                          \
  int NAME##_replay(const LangevinConsts&, const ReplayIO&, dim3, size_t, bool, cudaStream_t);         \
  int NAME##_stationary(const LangevinConsts&, float, const float*, const uint32_t*, long long, float*, \
                        float*, int, bool, cudaStream_t);                                              \
  int NAME##_apply(const LangevinConsts&, float, float*, uint32_t*, float*, long long, int, bool, cudaStream_t); \
  int NAME##_apply_nd(const LangevinConsts&, float, float*, uint32_t*, float*, long long, int, cudaStream_t);

NQ_DECLARE_POT_LAUNCHERS(pot_spring)
NQ_DECLARE_POT_LAUNCHERS(pot_biomolecule)
NQ_DECLARE_POT_LAUNCHERS(pot_shifted)

}  // namespace nq
