// Host side of the Langevin entry points of include/noneq_b200.h: descriptor -> kernel constants
// (float32, reference operation order), argument checks, kernel dispatch, deterministic final
// reduction.  Kernels: langevin_impl.cuh.
#include <cstring>
#include <mutex>
#include <stdlib.h>

#include "langevin_impl.cuh"

namespace nq {

int make_langevin_consts(const NqLangevinDesc* d, LangevinConsts* c) {
  NQ_CHECK_ARG(d != nullptr, "desc is null");
  NQ_CHECK_ARG(d->potential >= 0 && d->potential <= 2, "unknown potential kind %d", d->potential);
  NQ_CHECK_ARG(d->shift_kind == NQ_SHIFT_FREE || d->shift_kind == NQ_SHIFT_PERIODIC, "unknown shift kind %d",
               d->shift_kind);
  NQ_CHECK_ARG(d->math_mode == NQ_MATH_STRICT || d->math_mode == NQ_MATH_FAST, "unknown math_mode %d", d->math_mode);
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
  // brownian._dist (barrier_crossing.py:66-71), float32, one rounding per operation
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
  c->nudt = (float)((double)nu * (double)c->dt);
  {
    const double log2e = 1.4426950408889634, nudt = (double)nu * (double)c->dt;
    const double ib = d->potential == NQ_POT_SPRING ? 0.0 : 1.0 / d->beta;
    c->f_cl2 = (float)((double)c->cl * log2e);
    c->f_ncr2 = (float)(-(double)c->cr * log2e);
    c->f_nbde2 = (float)(-(double)c->bde * log2e);
    c->f_clm = (float)(nudt * ib * 2.0 * (double)c->cl);
    c->f_crm = (float)(nudt * ib * 2.0 * (double)c->cr);
    c->f_cu = (float)(2.0 * (double)c->hk * nudt);
  }
  // STRICT main-path window of the division / logarithm of A + B (div_main, logf_main): the numerator
  // 1 / beta must sit inside [2^-62, 2^62) too, otherwise every step takes the library functions
  c->guard_span = (c->inv_beta >= 0x1p-62f && c->inv_beta < 0x1p62f) ? 0x3e000000u : 0u;
  // Lower edge of that window as a float compare (the hot loop's single rare-case predicate): valid only where A + B
  // cannot leave the window at the top, i.e. both exponents are bounded above (A <= 1 for cl <= 0, B <= exp(-bde)
  // for cr >= 0); otherwise +inf sends every step through the patch block, which tests the full window
  {
    const bool bounded = c->cl <= 0.f && c->cr >= 0.f && c->bde > -40.f;   // A + B < 1 + e^40 < 2^62
    c->guard_lo = (c->guard_span != 0u && bounded) ? 0x1p-62f : __builtin_inff();
  }
  {
    const StrictRegs sr;
    c->exp_k1 = sr.exp_k1; c->exp_k252 = sr.exp_k252; c->l1p_c9 = sr.l1p_c9;
    const float pc[2] = {c->cl, -c->cr};   // f32x2: low half first
    memcpy(&c->well_c2, pc, 8);
  }
  c->one = 1u;
  c->S = d->steps;
  c->Neq = d->eq_steps;
  return NQ_OK;
}

// Sums the block partials; out_mom[8], out_grad[2][D].  One warp per column: lane l adds blocks
// l, l+32, ... in order and the lanes are combined by a fixed shuffle tree, so the result does not
// depend on scheduling (bit-deterministic gradients).
__global__ void __launch_bounds__(256) langevin_finalize_kernel(const double* partials, int nblocks, int NV, int DPS,
                                                                 int D, double* out_mom, double* out_grad,
                                                                 const unsigned int* status) {
  const int q = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= NV) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(long long)b * NV + q];
  s = warp_sum(s);
  if (status && *status) s = __longlong_as_double(0x7ff8000000000000ll);   // abandoned persistent launch: NaN, loudly
  if (lane != 0) return;
  if (q < 8) {
    if (out_mom) out_mom[q] = s;
  } else if (out_grad) {
    const int j = (q - 8) % DPS, which = (q - 8) / DPS;
    if (j < D) out_grad[which * D + j] = s;
  }
}

struct PotLaunchers {
  MainLauncher main;
  PairLauncher pair;
  PersistentLauncher persistent;
  ReplayLauncher replay;
  StationaryLauncher stationary;
  ApplyLauncher apply;
  ApplyNdLauncher apply_nd;
};

static PotLaunchers launchers(int potential) {
  switch (potential) {
    case NQ_POT_SPRING: return {pot_spring_main, pot_spring_pair, pot_spring_persistent, pot_spring_replay, pot_spring_stationary, pot_spring_apply, pot_spring_apply_nd};
    case NQ_POT_BIOMOLECULE:
      return {pot_biomolecule_main, pot_biomolecule_pair, pot_biomolecule_persistent, pot_biomolecule_replay, pot_biomolecule_stationary, pot_biomolecule_apply, pot_biomolecule_apply_nd};
    default: return {pot_shifted_main, pot_shifted_pair, pot_shifted_persistent, pot_shifted_replay, pot_shifted_stationary, pot_shifted_apply, pot_shifted_apply_nd};
  }
}

// template D actually launched for a requested D (see NQ_DEFINE_POT_LAUNCHERS)
static int kernel_D(int D) { return D == 0 ? 0 : (D == 13 ? 13 : pad4(D)); }
static int partial_row(int D) { return 8 + 2 * (D > 0 ? pad4(kernel_D(D)) : 4); }

// Dynamic (persistent) scheduling: a fixed number of CTAs per SM pop (trajectory group, chunk of steps)
// items from a FIFO ring, so that work flows to whichever SM is free (langevin_persistent_kernel).
// Measured on B200: with 8 CTAs per SM (4 warps per sub-partition) and 128-step chunks
// config 2 runs at 7.5e10 traj-steps/s against 6.75e10 for the one-wave static launch (STRICT); FAST
// prefers 6 CTAs per SM and 256-step chunks (1.00e11 against 8.4e10).  The static launch stays better
// for small ensembles (hand-off overhead in the latency-bound regime) and, in STRICT mode, for
// ensembles of many waves (it keeps 12 CTAs per SM resident).  NQ_PERSISTENT = 0 / 1 forces the choice;
// NQ_PERSISTENT_CHUNK and NQ_PERSISTENT_PER_SM override the two parameters.
// The thresholds below are per SM (measured on B200, 148 SMs; they scale with the SM count of the device the
// call runs on).  The NQ_* environment overrides exist for the A/B tests and scans under tests/ and scripts/;
// every form gives bit-identical results, so they are tuning knobs, not semantics.
static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      sms = n;
    else
      return 148;   // no device (host-only argument checks)
  }
  return sms;
}
static long long chunk_steps(bool fast) {
  const char* env = getenv("NQ_PERSISTENT_CHUNK");
  const long long v = env ? atoll(env) : (fast ? 256 : 128);
  return v < 32 ? 32 : v;
}
static long long persistent_ctas_per_sm(bool fast, long long n) {
  const char* env = getenv("NQ_PERSISTENT_PER_SM");
  // 8 CTAs of 64 per SM (FAST: always; r02w: 1.03e11 against 1.00e11 at 6); STRICT beyond ~880 trajectories per SM: 10
  return env ? atoll(env) : kCtaScale * ((fast || n <= 880ll * sm_count()) ? 8 : 10);
}
static bool use_persistent(long long n, long long S, bool record, bool fast) {
  if (record || S < 2 * chunk_steps(fast)) return false;
  const char* env = getenv("NQ_PERSISTENT");
  if (env != nullptr) return env[0] == '1' && n >= (long long)sm_count() * kThreads * 2;
  // from ~270 trajectories per SM (below: hand-off overhead in the latency regime) to ~2030 per SM in STRICT
  // mode (above: the static launch keeps 11 CTAs per SM resident over many waves and is faster)
  return n > 270ll * sm_count() && (fast || n <= 2030ll * sm_count());
}
static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// Small ensembles run a producer + consumer warp per 32 trajectories (langevin_pair_kernel): up to
// this many trajectories the GPU is latency-bound and overlapping the random stream with the
// integration wins; beyond it the one-warp-per-32-trajectories kernel has the higher throughput.
// NQ_PAIR_MAX overrides the threshold (0 = never).
constexpr long long kPairLimit = 65536;   // nq_langevin_workspace_bytes sizes the partial rows for 32-trajectory CTAs up to here
static long long pair_max() {
  const char* env = getenv("NQ_PAIR_MAX");
  // ~83 trajectories per SM (12 288 on B200): measured (scripts/latency_scan.py, r02), the three-warp pipeline wins
  // up to here; beyond it the one-warp kernel is as fast (STRICT) or faster (FAST)
  const long long v = env ? atoll(env) : 83ll * sm_count() + 4;
  return v > kPairLimit ? kPairLimit : v;
}
static bool use_pair(long long n, bool record) { return !record && n <= pair_max(); }
static size_t persistent_bytes(long long n, int D) {
  const long long groups = (n + kThreads - 1) / kThreads;
  const long long n_chunks_max = 1 << 16;  // callers size the workspace without knowing S: bound the ring by n
  (void)n_chunks_max;
  return align256((size_t)state_words(kernel_D(D)) * groups * kThreads * 4) + align256((size_t)(groups + 4) * 4);
}

static int normal_table(cudaStream_t st, const float** out);

static int run(const NqLangevinDesc* desc, const LangevinIO& io_in, double* moments, double* grad_sums,
               void* workspace, size_t workspace_bytes, cudaStream_t st) {
  LangevinConsts c;
  int rc = make_langevin_consts(desc, &c);
  if (rc) return rc;
  LangevinIO io = io_in;
  NQ_CHECK_ARG(io.n >= 1, "n must be >= 1 (got %lld)", io.n);
  NQ_CHECK_ARG(io.r_table && io.x0 && io.seeds && io.work && io.logp, "null required pointer");
  NQ_CHECK_ARG(io.D >= 0 && io.D <= 32, "D must be in [0, 32] (got %d)", io.D);
  NQ_CHECK_ARG(io.D == 0 || io.phi, "phi is null");
  const bool fast = desc->math_mode == NQ_MATH_FAST;
  const int NV = partial_row(io.D);
  // the pair kernel can record positions on the side (also with D > 0); everything else that records runs the
  // one-warp RECORD kernel, forward only
  const bool record_other = io.step_works || io.step_logps || io.ckpt_x;
  const bool pair = use_pair(io.n, record_other);
  const bool record = record_other || (io.positions && !pair);
  const int blocks = (int)(pair ? (io.n + 31) / 32 : (io.n + kThreads - 1) / kThreads);
  const long long kChunkSteps = chunk_steps(fast);
  const bool persistent = !pair && use_persistent(io.n, c.S, record, fast);
  const size_t part_bytes = align256((size_t)blocks * NV * sizeof(double));
  const size_t ring_bytes = persistent ? align256((size_t)blocks * ((c.S + kChunkSteps - 1) / kChunkSteps) * 4) : 0;
  size_t need = part_bytes + (persistent ? persistent_bytes(io.n, io.D) : 0);
  bool persistent_ok = persistent;
  int* ring = nullptr;
  if (persistent) {
    // the ring (4 B per work item) does not fit the S-independent workspace bound when S is huge: fall back
    if (workspace_bytes >= need + ring_bytes) {
      ring = (int*)((char*)workspace + need);
      need += ring_bytes;
    } else {
      persistent_ok = false;
      need = part_bytes;
    }
  }
  if (workspace_bytes < need || workspace == nullptr) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return NQ_ERR_WORKSPACE;
  }
  io.partials = (double*)workspace;
  if (fast ? kNormalTableFast : kNormalTable) {
    rc = normal_table(st, &io.ztable);
    if (rc != NQ_OK) return rc;
  }
  const unsigned int* status = nullptr;
  if (record) NQ_CHECK_ARG(io.D == 0, "per-step recording is only available on the forward entry points");
  if (io.positions) NQ_CHECK_ARG(io.pos_stride >= 1, "pos_stride must be >= 1");
  if (io.ckpt_x) NQ_CHECK_ARG(io.ckpt_rng && io.ckpt_every >= 1, "checkpointing needs ckpt_rng and ckpt_every >= 1");
  if (persistent_ok) {
    PersistentCtl ctl;
    char* base = (char*)workspace + part_bytes;
    ctl.groups = blocks;
    ctl.npad = (long long)blocks * kThreads;
    ctl.chunk = kChunkSteps;
    ctl.n_chunks = (c.S + kChunkSteps - 1) / kChunkSteps;
    ctl.state = (uint32_t*)base;
    const size_t state_bytes = align256((size_t)state_words(kernel_D(io.D)) * ctl.npad * 4);
    ctl.head = (unsigned int*)(base + state_bytes);
    ctl.tail = ctl.head + 1;
    ctl.status = ctl.head + 2;
    ctl.progress = (int*)(ctl.head + 4);
    {
      const char* env = getenv("NQ_WAIT_POLLS");   // bound of a scheduler wait in 100 ns polls; 0 = unbounded (sanitizers)
      ctl.max_polls = env ? (unsigned int)strtoul(env, nullptr, 10) : (1u << 25);
    }
    status = ctl.status;
    ctl.ring = ring;
    persistent_init_kernel<<<64, 256, 0, st>>>(ctl);
    NQ_LAUNCHED();
    int occ = 0, sms = 0, dev = 0;
    NQ_CUDA(cudaGetDevice(&dev));
    NQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    rc = launchers(c.potential).persistent(c, io, ctl, 0, fast, st, &occ);
    NQ_CHECK_ARG(rc == 0 && occ > 0, "no persistent kernel for D=%d", io.D);
    // fewer resident CTAs than chains, the same number on every SM
    long long per_sm = persistent_ctas_per_sm(fast, io.n);
    if (per_sm > occ) per_sm = occ;
    if (per_sm < 1) per_sm = 1;
    rc = launchers(c.potential).persistent(c, io, ctl, (int)(per_sm * sms), fast, st, nullptr);
    NQ_CHECK_ARG(rc == 0, "no persistent kernel for D=%d", io.D);
  } else if (pair) {
    rc = launchers(c.potential).pair(c, io, blocks, fast, st);
    NQ_CHECK_ARG(rc == 0, "no kernel for D=%d", io.D);
  } else {
    rc = launchers(c.potential).main(c, io, blocks, fast, record, st);
    NQ_CHECK_ARG(rc == 0, "no kernel for D=%d", io.D);
  }
  NQ_LAUNCHED();
  if (moments || grad_sums) {
    langevin_finalize_kernel<<<(NV + 7) / 8, 256, 0, st>>>(io.partials, blocks, NV, (NV - 8) / 2, io.D, moments, grad_sums,
                                                        status);
    NQ_LAUNCHED();
  }
  return NQ_OK;
}

// ---- the normal-draw table of the STRICT hot loops
// z[p] = normal_from_bits<false>(p << 9) for the 2^23 patterns jax.random.normal distinguishes: 32 MB, built once per
// device by the same device function the kernels used to call per step, so a lookup returns the same bits.  The table
// lives for the life of the process (like a cuBLAS workspace); it is the one allocation the library makes.
constexpr long long kNormalTableSize = 1ll << 23;
__global__ void __launch_bounds__(256) normal_table_kernel(float* table) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < kNormalTableSize) table[p] = normal_from_bits<false>((uint32_t)p << 9);
}
static int normal_table(cudaStream_t st, const float** out) {
  static std::mutex mu;
  static float* tables[64] = {};
  int dev = 0;
  NQ_CUDA(cudaGetDevice(&dev));
  NQ_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (tables[dev] == nullptr) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NQ_CUDA(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone) {
      set_error("the normal-draw table of this device does not exist yet and cannot be built during stream capture: "
                "call nq_langevin_prepare() (or run one STRICT Langevin call) before capturing");
      return NQ_ERR_UNSUPPORTED;
    }
    float* t = nullptr;
    NQ_CUDA(cudaMalloc(&t, kNormalTableSize * sizeof(float)));
    normal_table_kernel<<<(unsigned)(kNormalTableSize / 256), 256, 0, st>>>(t);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // complete before any other stream can use it
    if (e != cudaSuccess) {
      cudaFree(t);
      return cuda_fail(e, "normal table build");
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    tables[dev] = t;
  }
  *out = tables[dev];
  return NQ_OK;
}

__global__ void zero_doubles_kernel(double* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

}  // namespace nq

using namespace nq;

extern "C" int nq_langevin_prepare(void* stream) {
  NQ_RANGE();
  const float* t = nullptr;
  return kNormalTable ? normal_table((cudaStream_t)stream, &t) : NQ_OK;
}

extern "C" size_t nq_langevin_workspace_bytes(int64_t n, int32_t D) {
  if (n < 1) n = 1;
  if (D < 0) D = 0;
  if (D > 32) D = 32;
  const size_t blocks = (size_t)((n + kThreads - 1) / kThreads);
  // + ring of work items for up to 256 chunks of kChunkSteps steps (S <= 131072); longer runs use the static kernel
  // block partial rows: small ensembles may run the pair kernel (32 trajectories per CTA)
  const size_t pblocks = n <= kPairLimit ? (size_t)((n + 31) / 32) : blocks;
  return align256(pblocks * partial_row(D) * sizeof(double)) + persistent_bytes(n, D) + align256(blocks * 256 * 4);
}

extern "C" int nq_langevin_forward(const NqLangevinDesc* desc, const float* r_table, const float* x0,
                                   const uint32_t* seeds, int64_t n, float* work, float* logp, float* x_final,
                                   float* positions, int64_t pos_stride, float* step_works, float* step_logps,
                                   double* moments, void* workspace, size_t workspace_bytes, void* stream) {
  NQ_RANGE();
  LangevinIO io{};
  io.r_table = r_table; io.x0 = x0; io.seeds = seeds;
  io.work = work; io.logp = logp; io.x_final = x_final;
  io.positions = positions; io.pos_stride = pos_stride; io.step_works = step_works; io.step_logps = step_logps;
  io.n = n; io.D = 0;
  return run(desc, io, moments, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int nq_langevin_grad(const NqLangevinDesc* desc, const float* r_table, const float* phi, int32_t D,
                                const float* x0, const uint32_t* seeds, int64_t n, float* work, float* logp,
                                float* estimator, float* x_final, double* grad_sums, double* moments,
                                void* workspace, size_t workspace_bytes, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(D >= 1, "D must be >= 1 for the gradient entry point");
  NQ_CHECK_ARG(grad_sums != nullptr, "grad_sums is null");
  LangevinIO io{};
  io.r_table = r_table; io.phi = phi; io.x0 = x0; io.seeds = seeds;
  io.work = work; io.logp = logp; io.estimator = estimator; io.x_final = x_final;
  io.n = n; io.D = D;
  return run(desc, io, moments, grad_sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int nq_langevin_grad_record(const NqLangevinDesc* desc, const float* r_table, const float* phi, int32_t D,
                                       const float* x0, const uint32_t* seeds, int64_t n, float* work, float* logp,
                                       float* estimator, float* x_final, float* positions, int64_t pos_stride,
                                       double* grad_sums, double* moments, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(D >= 1, "D must be >= 1 for the gradient entry point");
  NQ_CHECK_ARG(grad_sums != nullptr && positions != nullptr && pos_stride >= 1, "grad_sums / positions / pos_stride");
  if (!use_pair(n, false)) {
    set_error("the fused gradient pass records positions only for small ensembles (n <= %lld); run nq_langevin_forward "
              "for the record", pair_max());
    return NQ_ERR_UNSUPPORTED;
  }
  LangevinIO io{};
  io.r_table = r_table; io.phi = phi; io.x0 = x0; io.seeds = seeds;
  io.work = work; io.logp = logp; io.estimator = estimator; io.x_final = x_final;
  io.positions = positions; io.pos_stride = pos_stride;
  io.n = n; io.D = D;
  return run(desc, io, moments, grad_sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t nq_langevin_ckpt_count(int64_t steps, int64_t ckpt_every) {
  return ckpt_every > 0 && steps > 0 ? (size_t)((steps + ckpt_every - 1) / ckpt_every) : 0;
}

extern "C" int nq_langevin_forward_ckpt(const NqLangevinDesc* desc, const float* r_table, const float* x0,
                                        const uint32_t* seeds, int64_t n, int64_t ckpt_every, float* work,
                                        float* logp, float* x_final, float* ckpt_x, uint32_t* ckpt_rng,
                                        double* moments, void* workspace, size_t workspace_bytes, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(ckpt_x && ckpt_rng && ckpt_every >= 1, "checkpoint buffers / interval missing");
  LangevinIO io{};
  io.r_table = r_table; io.x0 = x0; io.seeds = seeds;
  io.work = work; io.logp = logp; io.x_final = x_final;
  io.ckpt_x = ckpt_x; io.ckpt_rng = ckpt_rng; io.ckpt_every = ckpt_every; io.pos_stride = 1;
  io.n = n; io.D = 0;
  return run(desc, io, moments, nullptr, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int nq_langevin_replay(const NqLangevinDesc* desc, const float* r_table, int64_t n, int64_t ckpt_every,
                                  const float* ckpt_x, const uint32_t* ckpt_rng, const float* lbar,
                                  const float* wbar, double* gbar, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  NQ_RANGE();
  (void)workspace;
  (void)workspace_bytes;
  LangevinConsts c;
  int rc = make_langevin_consts(desc, &c);
  if (rc) return rc;
  NQ_CHECK_ARG(n >= 1 && ckpt_every >= 1 && r_table && ckpt_x && ckpt_rng && lbar && wbar && gbar, "bad arguments");
  NQ_CHECK_ARG(ckpt_every <= 4096, "ckpt_every must be <= 4096 (shared-memory accumulator)");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_ckpt = (c.S + ckpt_every - 1) / ckpt_every;
  NQ_CHECK_ARG(n_ckpt <= 65535, "too many segments (%lld); raise ckpt_every", n_ckpt);
  zero_doubles_kernel<<<(unsigned)((c.S + 1 + 255) / 256), 256, 0, st>>>(gbar, c.S + 1);
  NQ_LAUNCHED();
  ReplayIO io{r_table, ckpt_x, ckpt_rng, lbar, wbar, gbar, n, ckpt_every, n_ckpt};
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)n_ckpt);
  const size_t smem = (size_t)(ckpt_every + 1) * sizeof(double);
  launchers(c.potential).replay(c, io, grid, smem, desc->math_mode == NQ_MATH_FAST, st);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_langevin_stationary(const NqLangevinDesc* desc, float r0, const float* x0, const uint32_t* seeds,
                                      int64_t n, float* x_final, float* positions, void* stream) {
  NQ_RANGE();
  LangevinConsts c;
  int rc = make_langevin_consts(desc, &c);
  if (rc) return rc;
  NQ_CHECK_ARG(n >= 1 && x0 && seeds, "bad arguments");
  const int blocks = (int)((n + kThreads - 1) / kThreads);
  launchers(c.potential).stationary(c, r0, x0, seeds, n, x_final, positions, blocks, desc->math_mode == NQ_MATH_FAST,
                                    (cudaStream_t)stream);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_langevin_apply(const NqLangevinDesc* desc, float r0, float* x, uint32_t* rng, float* log_prob,
                                 int64_t n, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(desc != nullptr, "desc is null");
  NqLangevinDesc d1 = *desc;
  if (d1.steps < 1) d1.steps = 1;
  LangevinConsts c;
  int rc = make_langevin_consts(&d1, &c);
  if (rc) return rc;
  NQ_CHECK_ARG(n >= 1 && x && rng, "bad arguments");
  const int blocks = (int)((n + kThreads - 1) / kThreads);
  launchers(c.potential).apply(c, r0, x, rng, log_prob, n, blocks, desc->math_mode == NQ_MATH_FAST, (cudaStream_t)stream);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_langevin_apply_nd(const NqLangevinDesc* desc, float r0, float* position, uint32_t* rng, float* log_prob,
                                    int64_t n_states, int32_t n_particles, int32_t dim, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(desc != nullptr, "desc is null");
  NqLangevinDesc d1 = *desc;
  if (d1.steps < 1) d1.steps = 1;
  LangevinConsts c;
  int rc = make_langevin_consts(&d1, &c);
  if (rc) return rc;
  NQ_CHECK_ARG(n_states >= 1 && n_states <= 0x7fffffffll && position && rng, "bad arguments");
  NQ_CHECK_ARG(n_particles >= 1 && dim >= 1 && (long long)n_particles * dim <= 0x3fffffffll, "bad state shape %d x %d",
               n_particles, dim);
  launchers(c.potential).apply_nd(c, r0, position, rng, log_prob, n_states, n_particles * dim, (cudaStream_t)stream);
  NQ_LAUNCHED();
  return NQ_OK;
}
