// Batched overdamped-Langevin integrator with per-step work / log-prob accumulation and the
// fused REINFORCE / log-prob / reparameterisation gradient moments (sm_100a).
//
// Replaces the XLA computation traced from the reference's
//   brownian / apply_fn              noneq_opt/barrier_crossing.py:28-94
//   run_brownian_opt (lax.scan)      noneq_opt/barrier_crossing.py:97-143
//   estimate_gradient_boltzinit*     noneq_opt/barrier_crossing.py:204-224, 262-333
//   run_brownian_stationary_trap     noneq_opt/barrier_crossing.py:227-241
//
// Design: one thread per trajectory; position, threefry key chain, work / log-prob sums and the
// 2*D gradient accumulators live in registers for the whole trajectory.  The trap table r[k]
// and the protocol Jacobian phi[k][:] are staged tile by tile (TILE steps) into shared memory
// and read as warp-uniform broadcasts.  HBM traffic is ~24 B per *trajectory*; the kernel is
// bound by instruction issue (3 threefry2x32-20 blocks per step), see DESIGN.md.
#include <math.h>

#include "nq_common.cuh"

namespace nq {

constexpr int kTile = 128;     // steps staged per shared-memory tile
constexpr int kThreads = 64;   // trajectories per CTA (small CTAs balance 1e5 trajectories over 148 SMs)
// resident CTAs per SM the register budget is held to, so that config 2 (1563 CTAs) is one wave
constexpr int min_blocks(int DP) { return DP <= 8 ? 16 : (DP <= 16 ? 11 : (DP <= 24 ? 8 : 6)); }

struct LangevinConsts {
  int potential, shift_kind;
  float hk, k_s, inv_beta, cl, cr, bde, xm, two_xm;
  float nu, dt, sigma, inv_sigma, log_norm, box;
  float ca;  // nu*dt*k_s/sigma : d logp_k / d r_k = z_k * ca
  long long S, Neq;
};

static int make_consts(const NqLangevinDesc* d, LangevinConsts* c) {
  NQ_CHECK_ARG(d != nullptr, "desc is null");
  NQ_CHECK_ARG(d->potential >= 0 && d->potential <= 2, "unknown potential kind %d", d->potential);
  NQ_CHECK_ARG(d->shift_kind == NQ_SHIFT_FREE || d->shift_kind == NQ_SHIFT_PERIODIC,
               "unknown shift kind %d", d->shift_kind);
  NQ_CHECK_ARG(d->steps >= 1, "steps must be >= 1 (got %lld)", (long long)d->steps);
  NQ_CHECK_ARG(d->eq_steps >= 0, "eq_steps must be >= 0");
  NQ_CHECK_ARG(d->dt > 0 && d->kT > 0 && d->gamma > 0 && d->mass > 0, "dt, kT, gamma, mass must be > 0");
  if (d->potential != NQ_POT_SPRING) NQ_CHECK_ARG(d->beta > 0, "beta must be > 0");
  if (d->shift_kind == NQ_SHIFT_PERIODIC) NQ_CHECK_ARG(d->box > 0, "periodic box must be > 0");
  c->potential = d->potential;
  c->shift_kind = d->shift_kind;
  // python-float constants are folded in binary64 and cast once (jax weak typing)
  c->hk = (float)(d->k_s / 2);
  c->k_s = (float)d->k_s;
  c->inv_beta = d->potential == NQ_POT_SPRING ? 0.f : (float)(1. / d->beta);
  c->cl = (float)(-0.5 * d->beta * d->kappa_l);
  c->cr = (float)(0.5 * d->beta * d->kappa_r);
  c->bde = (float)(d->beta * d->delta_E);
  c->xm = (float)d->x_m;
  c->two_xm = (float)(2. * d->x_m);
  // brownian._dist (barrier_crossing.py:66-71), float32 one rounding per op
  volatile float mg = (float)d->mass * (float)d->gamma;
  volatile float nu = 1.0f / mg;
  volatile float v1 = 2.0f * (float)d->kT;
  volatile float v2 = v1 * (float)d->dt;
  volatile float var = v2 * nu;
  c->nu = nu;
  c->dt = (float)d->dt;
  c->sigma = sqrtf(var);
  c->inv_sigma = (float)(1.0 / (double)c->sigma);
  volatile float ls = (float)log((double)c->sigma);
  c->log_norm = (float)0.91893853320467274178 + ls;
  c->box = (float)d->box;
  c->ca = (float)((double)nu * (double)c->dt * (double)c->k_s / (double)c->sigma);
  c->S = d->steps;
  c->Neq = d->eq_steps;
  return NQ_OK;
}

// ------------------------------------------------------------------------------ device math
template <bool FAST> __device__ __forceinline__ float fmul(float a, float b) { return FAST ? a * b : __fmul_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fadd(float a, float b) { return FAST ? a + b : __fadd_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fsub(float a, float b) { return FAST ? a - b : __fsub_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float fexp(float a) { return FAST ? __expf(a) : expf(a); }
template <bool FAST> __device__ __forceinline__ float flog(float a) { return FAST ? __logf(a) : logf(a); }
template <bool FAST> __device__ __forceinline__ float fdiv(float a, float b) { return FAST ? __fdividef(a, b) : __fdiv_rn(a, b); }

// energy pieces shared between increment_work (:123-124) and the force (:60,68)
template <int POT, bool FAST>
struct Landscape {
  float ul, ur, A, B, sAB, Em;
  __device__ __forceinline__ void eval(const LangevinConsts& c, float x) {
    if (POT == NQ_POT_SPRING) {
      Em = 0.f;
      return;
    }
    if (POT == NQ_POT_BIOMOLECULE) {
      ul = fadd<FAST>(x, c.xm);
      ur = fsub<FAST>(x, c.xm);
    } else {
      ul = x;
      ur = fsub<FAST>(x, c.two_xm);
    }
    float a = fmul<FAST>(c.cl, fmul<FAST>(ul, ul));
    float v = fadd<FAST>(fmul<FAST>(c.cr, fmul<FAST>(ur, ur)), c.bde);
    A = fexp<FAST>(a);
    B = fexp<FAST>(-v);
    sAB = fadd<FAST>(A, B);
    Em = fmul<FAST>(-c.inv_beta, flog<FAST>(sAB));
  }
  // dEm/dx in the order of jax's VJP rules (log: g/x, exp: g*ans, x**2: g*(2x))
  __device__ __forceinline__ float dEm(const LangevinConsts& c) const {
    if (POT == NQ_POT_SPRING) return 0.f;
    float g = fdiv<FAST>(-c.inv_beta, sAB);
    float ga = fmul<FAST>(g, A), gb = fmul<FAST>(g, B);
    float dxa = fmul<FAST>(fmul<FAST>(ga, c.cl), fmul<FAST>(2.0f, ul));
    float dxb = fmul<FAST>(fmul<FAST>(-gb, c.cr), fmul<FAST>(2.0f, ur));
    return fadd<FAST>(dxa, dxb);
  }
};

template <bool FAST>
__device__ __forceinline__ float spring_energy(const LangevinConsts& c, float x, float r) {
  float u = fsub<FAST>(x, r);
  return fmul<FAST>(c.hk, fmul<FAST>(u, u));
}

__device__ __forceinline__ float py_mod(float x, float box) {
  float r = fmodf(x, box);
  if (r != 0.f && ((r < 0.f) != (box < 0.f))) r += box;
  return r;
}

// One apply_fn (barrier_crossing.py:77-92) given the landscape already evaluated at x.
// Returns dR; updates x and the key chain; writes z (standardised noise incl. rounding) and lp.
template <int POT, bool FAST>
__device__ __forceinline__ float brownian_apply(const LangevinConsts& c, const Landscape<POT, FAST>& land,
                                                float& x, uint32_t& k0, uint32_t& k1, float r,
                                                float& z, float& lp) {
  float u = fsub<FAST>(x, r);
  float dEs = fmul<FAST>(c.hk, fmul<FAST>(2.0f, u));
  float F = POT == NQ_POT_SPRING ? -dEs : -fadd<FAST>(land.dEm(c), dEs);
  float mean = fmul<FAST>(fmul<FAST>(F, c.nu), c.dt);
  uint32_t n0, n1, s0, s1;
  split2(k0, k1, n0, n1, s0, s1);  // key, split = jax.random.split(state.rng)   :79
  k0 = n0;
  k1 = n1;
  uint32_t bits, unused;
  threefry2x32(key_schedule(s0, s1), 0u, 0u, bits, unused);  // normal(split, [1,1,1])
  float xi = normal_from_bits<FAST>(bits);
  float dR = fadd<FAST>(fmul<FAST>(xi, c.sigma), mean);  // sampled * scale + loc
  if (FAST) {
    z = (dR - mean) * c.inv_sigma;
  } else {
    z = __fsub_rn(__fdiv_rn(dR, c.sigma), __fdiv_rn(mean, c.sigma));
  }
  lp = fsub<FAST>(fmul<FAST>(-0.5f, fmul<FAST>(z, z)), c.log_norm);
  x = fadd<FAST>(x, dR);
  if (c.shift_kind == NQ_SHIFT_PERIODIC) x = py_mod(x, c.box);
  return dR;
}

// ------------------------------------------------------------------------------ main kernel
struct LangevinIO {
  const float* r_table;  // [S+1]
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
  double* partials;      // [gridDim.x][NV] block partial sums
  long long n;
  int D;
};

// NV layout of a partial row: [0..7] moments, [8 .. 8+DP) logP grads, [8+DP .. 8+2DP) reparam grads
template <int POT, bool FAST, int DP, bool RECORD>
__global__ void __launch_bounds__(kThreads, min_blocks(DP)) langevin_kernel(const LangevinConsts c, const LangevinIO io) {
  constexpr bool GRAD = DP > 0;
  constexpr int DPS = GRAD ? DP : 4;
  __shared__ float r_s[kTile + 1];
  __shared__ __align__(16) float phi_s[GRAD ? kTile * DPS : 4];
  __shared__ double red_s[kThreads / 32][8 + 2 * DPS];

  const long long gid = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool live = gid < io.n;
  const long long idx = live ? gid : io.n - 1;

  float x = io.x0[idx];
  uint32_t k0, k1;
  {
    // key, split = random.split(key); init(split, ...)     barrier_crossing.py:133,138
    uint32_t n0, n1;
    split2(io.seeds[2 * idx], io.seeds[2 * idx + 1], n0, n1, k0, k1);
  }
  Landscape<POT, FAST> land;
  float z, lp;
  {
    const float r_eq = io.r_table[0];
    for (long long e = 0; e < c.Neq; ++e) {  // equilibrate :115-121
      land.eval(c, x);
      brownian_apply<POT, FAST>(c, land, x, k0, k1, r_eq, z, lp);
    }
  }
  if (RECORD && io.positions && live) io.positions[gid] = x;   // positions[0] = x_eq  :142
  if (RECORD && io.step_works && live) io.step_works[gid] = 0.f;  // works[0] = 0      :141

  double W = 0.0, LP = 0.0;
  float gA[DPS], gB[DPS];
#pragma unroll
  for (int j = 0; j < DPS; ++j) gA[j] = gB[j] = 0.f;

  for (long long t0 = 1; t0 <= c.S; t0 += kTile) {
    const int len = (int)min((long long)kTile, c.S - t0 + 1);
    __syncthreads();
    for (int i = threadIdx.x; i <= len; i += kThreads) r_s[i] = io.r_table[t0 - 1 + i];
    if (GRAD) {
      // phi row for step k is k-1 (k = 1..S-1); step S (fixed end point) contributes nothing
      for (int i = threadIdx.x; i < len * DPS; i += kThreads) {
        const int row = i / DPS, col = i - row * DPS;
        const long long k = t0 + row;
        phi_s[i] = (col < io.D && k <= c.S - 1) ? io.phi[(k - 1) * io.D + col] : 0.f;
      }
    }
    __syncthreads();
    float Wt = 0.f, LPt = 0.f;
#pragma unroll 1
    for (int i = 0; i < len; ++i) {
      const float r_prev = r_s[i], r_cur = r_s[i + 1];
      land.eval(c, x);
      // increment_work :123-124 -- position BEFORE the kick
      const float e1 = fadd<FAST>(land.Em, spring_energy<FAST>(c, x, r_cur));
      const float e0 = fadd<FAST>(land.Em, spring_energy<FAST>(c, x, r_prev));
      const float dW = POT == NQ_POT_SPRING
                           ? fsub<FAST>(spring_energy<FAST>(c, x, r_cur), spring_energy<FAST>(c, x, r_prev))
                           : fsub<FAST>(e1, e0);
      const float dR = brownian_apply<POT, FAST>(c, land, x, k0, k1, r_cur, z, lp);
      Wt += dW;
      LPt += lp;
      if (GRAD) {
        const float a = z * c.ca;     // d logp_k / d r_k
        const float b = c.k_s * dR;   // d W / d r_k
        const float4* row = reinterpret_cast<const float4*>(phi_s + i * DPS);
#pragma unroll
        for (int q = 0; q < DPS / 4; ++q) {
          const float4 p = row[q];
          gA[4 * q + 0] = fmaf(a, p.x, gA[4 * q + 0]); gB[4 * q + 0] = fmaf(b, p.x, gB[4 * q + 0]);
          gA[4 * q + 1] = fmaf(a, p.y, gA[4 * q + 1]); gB[4 * q + 1] = fmaf(b, p.y, gB[4 * q + 1]);
          gA[4 * q + 2] = fmaf(a, p.z, gA[4 * q + 2]); gB[4 * q + 2] = fmaf(b, p.z, gB[4 * q + 2]);
          gA[4 * q + 3] = fmaf(a, p.w, gA[4 * q + 3]); gB[4 * q + 3] = fmaf(b, p.w, gB[4 * q + 3]);
        }
      }
      if (RECORD && live) {
        const long long k = t0 + i;
        if (io.step_works) io.step_works[k * io.n + gid] = dW;
        if (io.step_logps) io.step_logps[(k - 1) * io.n + gid] = lp;
        if (io.positions && (k % io.pos_stride) == 0) io.positions[(k / io.pos_stride) * io.n + gid] = x;
      }
    }
    W += (double)Wt;   // two-level accumulation: f32 within a tile, binary64 across tiles
    LP += (double)LPt;
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
  double v[8 + 2 * DPS];
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
    v[8 + j] = GRAD ? m * Wd * (double)gA[j] : 0.0;
    v[8 + DPS + j] = GRAD ? m * (double)gB[j] : 0.0;
  }
  constexpr int NV = 8 + 2 * DPS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const double s = warp_sum(v[q]);
    if (lane == 0) red_s[warp][q] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red_s[w][threadIdx.x];
    io.partials[(long long)blockIdx.x * NV + threadIdx.x] = s;
  }
}

// sums block partials in block order; out_mom[8], out_grad[2][D]
__global__ void langevin_finalize_kernel(const double* partials, int nblocks, int NV, int DP, int D,
                                         double* out_mom, double* out_grad) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= NV) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partials[(long long)b * NV + q];
  if (q < 8) {
    if (out_mom) out_mom[q] = s;
  } else if (out_grad) {
    const int j = (q - 8) % DP, which = (q - 8) / DP;
    if (j < D) out_grad[which * D + j] = s;
  }
}

template <int POT, bool FAST, int DP, bool RECORD>
static void launch_one(const LangevinConsts& c, const LangevinIO& io, int blocks, cudaStream_t st) {
  langevin_kernel<POT, FAST, DP, RECORD><<<blocks, kThreads, 0, st>>>(c, io);
}

template <int DP, bool RECORD>
static int launch_pot(const LangevinConsts& c, const LangevinIO& io, int blocks, bool fast, cudaStream_t st) {
  switch (c.potential) {
    case NQ_POT_SPRING:
      fast ? launch_one<NQ_POT_SPRING, true, DP, RECORD>(c, io, blocks, st)
           : launch_one<NQ_POT_SPRING, false, DP, RECORD>(c, io, blocks, st);
      break;
    case NQ_POT_BIOMOLECULE:
      fast ? launch_one<NQ_POT_BIOMOLECULE, true, DP, RECORD>(c, io, blocks, st)
           : launch_one<NQ_POT_BIOMOLECULE, false, DP, RECORD>(c, io, blocks, st);
      break;
    default:
      fast ? launch_one<NQ_POT_BIOMOLECULE_SHIFTED, true, DP, RECORD>(c, io, blocks, st)
           : launch_one<NQ_POT_BIOMOLECULE_SHIFTED, false, DP, RECORD>(c, io, blocks, st);
      break;
  }
  NQ_LAUNCHED();
  return NQ_OK;
}

static int pad_D(int D) { return (D + 3) / 4 * 4; }

static int run(const NqLangevinDesc* desc, const LangevinIO& io_in, double* moments, double* grad_sums,
               void* workspace, size_t workspace_bytes, cudaStream_t st) {
  LangevinConsts c;
  int rc = make_consts(desc, &c);
  if (rc) return rc;
  LangevinIO io = io_in;
  NQ_CHECK_ARG(io.n >= 1, "n must be >= 1 (got %lld)", io.n);
  NQ_CHECK_ARG(io.r_table && io.x0 && io.seeds && io.work && io.logp, "null required pointer");
  NQ_CHECK_ARG(io.D >= 0 && io.D <= 32, "D must be in [0, 32] (got %d)", io.D);
  NQ_CHECK_ARG(io.D == 0 || io.phi, "phi is null");
  NQ_CHECK_ARG(desc->math_mode == NQ_MATH_STRICT || desc->math_mode == NQ_MATH_FAST, "bad math_mode");
  const bool fast = desc->math_mode == NQ_MATH_FAST;
  const int DP = pad_D(io.D);
  const int blocks = (int)((io.n + kThreads - 1) / kThreads);
  const int NV = 8 + 2 * (DP > 0 ? DP : 4);
  const size_t need = (size_t)blocks * NV * sizeof(double);
  if (workspace_bytes < need || workspace == nullptr) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return NQ_ERR_WORKSPACE;
  }
  io.partials = (double*)workspace;
  const bool record = io.positions || io.step_works || io.step_logps;
  if (record) NQ_CHECK_ARG(io.D == 0, "per-step recording is only available on the forward entry point");
  if (io.positions) NQ_CHECK_ARG(io.pos_stride >= 1, "pos_stride must be >= 1");
  switch (DP) {
    case 0:
      rc = record ? launch_pot<0, true>(c, io, blocks, fast, st) : launch_pot<0, false>(c, io, blocks, fast, st);
      break;
    case 4: rc = launch_pot<4, false>(c, io, blocks, fast, st); break;
    case 8: rc = launch_pot<8, false>(c, io, blocks, fast, st); break;
    case 12: rc = launch_pot<12, false>(c, io, blocks, fast, st); break;
    case 16: rc = launch_pot<16, false>(c, io, blocks, fast, st); break;
    case 20: rc = launch_pot<20, false>(c, io, blocks, fast, st); break;
    case 24: rc = launch_pot<24, false>(c, io, blocks, fast, st); break;
    case 28: rc = launch_pot<28, false>(c, io, blocks, fast, st); break;
    default: rc = launch_pot<32, false>(c, io, blocks, fast, st); break;
  }
  if (rc) return rc;
  if (moments || grad_sums) {
    langevin_finalize_kernel<<<1, 128, 0, st>>>(io.partials, blocks, NV, DP > 0 ? DP : 4, io.D, moments, grad_sums);
    NQ_LAUNCHED();
  }
  return NQ_OK;
}

// stationary trap: same integrator, constant r0, no work bookkeeping
template <int POT, bool FAST>
__global__ void __launch_bounds__(kThreads) stationary_kernel(const LangevinConsts c, float r0, const float* x0,
                                                              const uint32_t* seeds, long long n, float* x_final,
                                                              float* positions) {
  const long long gid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gid >= n) return;
  float x = x0[gid];
  uint32_t k0, k1, n0, n1;
  split2(seeds[2 * gid], seeds[2 * gid + 1], n0, n1, k0, k1);  // :228
  Landscape<POT, FAST> land;
  float z, lp;
  for (long long s = 0; s < c.S; ++s) {
    land.eval(c, x);
    brownian_apply<POT, FAST>(c, land, x, k0, k1, r0, z, lp);
    if (positions) positions[s * n + gid] = x;
  }
  if (x_final) x_final[gid] = x;
}

// a single apply_fn step on caller-held state (no initial split: the state's rng is used as is)
template <int POT, bool FAST>
__global__ void __launch_bounds__(kThreads) apply_kernel(const LangevinConsts c, float r0, float* x_io,
                                                         uint32_t* rng, float* log_prob, long long n) {
  const long long gid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (gid >= n) return;
  float x = x_io[gid];
  uint32_t k0 = rng[2 * gid], k1 = rng[2 * gid + 1];
  Landscape<POT, FAST> land;
  float z, lp;
  land.eval(c, x);
  brownian_apply<POT, FAST>(c, land, x, k0, k1, r0, z, lp);
  x_io[gid] = x;
  rng[2 * gid] = k0;
  rng[2 * gid + 1] = k1;
  if (log_prob) log_prob[gid] = lp;
}

}  // namespace nq

using namespace nq;

extern "C" size_t nq_langevin_workspace_bytes(int64_t n, int32_t D) {
  if (n < 1) n = 1;
  const int DP = D > 0 ? pad_D(D) : 4;
  const size_t blocks = (size_t)((n + kThreads - 1) / kThreads);
  return blocks * (8 + 2 * DP) * sizeof(double);
}

extern "C" int nq_langevin_forward(const NqLangevinDesc* desc, const float* r_table, const float* x0,
                                   const uint32_t* seeds, int64_t n, float* work, float* logp, float* x_final,
                                   float* positions, int64_t pos_stride, float* step_works, float* step_logps,
                                   double* moments, void* workspace, size_t workspace_bytes, void* stream) {
  LangevinIO io{};
  io.r_table = r_table; io.phi = nullptr; io.x0 = x0; io.seeds = seeds;
  io.work = work; io.logp = logp; io.estimator = nullptr; io.x_final = x_final;
  io.positions = positions; io.pos_stride = pos_stride; io.step_works = step_works; io.step_logps = step_logps;
  io.n = n; io.D = 0;
  return run(desc, io, moments, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int nq_langevin_grad(const NqLangevinDesc* desc, const float* r_table, const float* phi, int32_t D,
                                const float* x0, const uint32_t* seeds, int64_t n, float* work, float* logp,
                                float* estimator, float* x_final, double* grad_sums, double* moments,
                                void* workspace, size_t workspace_bytes, void* stream) {
  NQ_CHECK_ARG(D >= 1, "D must be >= 1 for the gradient entry point");
  NQ_CHECK_ARG(grad_sums != nullptr, "grad_sums is null");
  LangevinIO io{};
  io.r_table = r_table; io.phi = phi; io.x0 = x0; io.seeds = seeds;
  io.work = work; io.logp = logp; io.estimator = estimator; io.x_final = x_final;
  io.n = n; io.D = D;
  return run(desc, io, moments, grad_sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int nq_langevin_stationary(const NqLangevinDesc* desc, float r0, const float* x0, const uint32_t* seeds,
                                      int64_t n, float* x_final, float* positions, void* stream) {
  LangevinConsts c;
  int rc = make_consts(desc, &c);
  if (rc) return rc;
  NQ_CHECK_ARG(n >= 1 && x0 && seeds, "bad arguments");
  const int blocks = (int)((n + kThreads - 1) / kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = desc->math_mode == NQ_MATH_FAST;
#define NQ_ST(P)                                                                                        \
  fast ? stationary_kernel<P, true><<<blocks, kThreads, 0, st>>>(c, r0, x0, seeds, n, x_final, positions) \
       : stationary_kernel<P, false><<<blocks, kThreads, 0, st>>>(c, r0, x0, seeds, n, x_final, positions)
  switch (c.potential) {
    case NQ_POT_SPRING: NQ_ST(NQ_POT_SPRING); break;
    case NQ_POT_BIOMOLECULE: NQ_ST(NQ_POT_BIOMOLECULE); break;
    default: NQ_ST(NQ_POT_BIOMOLECULE_SHIFTED); break;
  }
#undef NQ_ST
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_langevin_apply(const NqLangevinDesc* desc, float r0, float* x, uint32_t* rng, float* log_prob,
                                 int64_t n, void* stream) {
  NqLangevinDesc d1 = *desc;
  if (d1.steps < 1) d1.steps = 1;
  LangevinConsts c;
  int rc = make_consts(&d1, &c);
  if (rc) return rc;
  NQ_CHECK_ARG(n >= 1 && x && rng, "bad arguments");
  const int blocks = (int)((n + kThreads - 1) / kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = desc->math_mode == NQ_MATH_FAST;
#define NQ_AP(P)                                                                              \
  fast ? apply_kernel<P, true><<<blocks, kThreads, 0, st>>>(c, r0, x, rng, log_prob, n)       \
       : apply_kernel<P, false><<<blocks, kThreads, 0, st>>>(c, r0, x, rng, log_prob, n)
  switch (c.potential) {
    case NQ_POT_SPRING: NQ_AP(NQ_POT_SPRING); break;
    case NQ_POT_BIOMOLECULE: NQ_AP(NQ_POT_BIOMOLECULE); break;
    default: NQ_AP(NQ_POT_BIOMOLECULE_SHIFTED); break;
  }
#undef NQ_AP
  NQ_LAUNCHED();
  return NQ_OK;
}
