// Device-resident pieces of the optimisation loop around the hot path: the
// reference's drivers call jax.random.split, the protocol evaluation, the estimator and the
// jax.example_libraries Adam update once per training step from Python
// (run_barrier_crossing_opt.py:113-137, ising.py:263-290).  With these three kernels every operand
// of a step -- chain key, per-trajectory seeds, trap table, gradient sums, Adam moments, step
// counter -- lives in HBM, so a whole step is a fixed launch sequence that the host captures once
// in a CUDA graph and replays without synchronising.
#include <math.h>

#include "nq_common.cuh"

namespace nq {

// out[i] = element (first+i) of threefry_2x32(key, iota(size)) with the key read from device memory
__global__ void random_bits_devkey_kernel(const uint32_t* __restrict__ key, unsigned long long size,
                                          unsigned long long first, unsigned long long count, uint32_t* out) {
  const KeySchedule ks = key_schedule(key[0], key[1]);
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x)
    out[i] = random_bits_at(ks, first + i, size);
}

// table[0] = first, table[1 + i] = offset[i] + sum_j variables[j] * basis[i][j], table[K + 1] = last.
// The sum runs j = 0..D-1 with one binary32 rounding per product and per addition: the order of
// Chebyshev.__call__ (parameterization.py:175-179 as restated by noneq_opt_b200.parameterization).
__global__ void schedule_affine_kernel(const float* __restrict__ basis, const float* __restrict__ variables,
                                       const float* __restrict__ offset, float first, float last, long long K, int D,
                                       float* __restrict__ table) {
  extern __shared__ float w_s[];
  for (int j = threadIdx.x; j < D; j += blockDim.x) w_s[j] = variables[j];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < K + 2; i += (long long)gridDim.x * blockDim.x) {
    float v;
    if (i == 0) {
      v = first;
    } else if (i == K + 1) {
      v = last;
    } else {
      const float* row = basis + (i - 1) * D;
      float acc = 0.f;
      for (int j = 0; j < D; ++j) acc = __fadd_rn(acc, __fmul_rn(w_s[j], row[j]));
      v = offset ? __fadd_rn(acc, offset[i - 1]) : acc;
    }
    table[i] = v;
  }
}

// table[i] = ((scale[i] * sum_j variables[j] * basis[i][j]) + offset1[i]) + offset2[i], one binary32 rounding per
// operation: ConstrainEndpoints (envelope * wrapped + linear, parameterization.py:295-322) around Chebyshev
// (:175-179) inside AddBaseline (wrapped + baseline, :325-344) -- the schedule composition of the Ising drivers
// (run_ising_opt.py:45-66).  scale / offset1 / offset2 are nullable.
__global__ void schedule_compose_kernel(const float* __restrict__ basis, const float* __restrict__ variables,
                                        const float* __restrict__ scale, const float* __restrict__ offset1,
                                        const float* __restrict__ offset2, long long K, int D, float* __restrict__ table) {
  extern __shared__ float w_s[];
  for (int j = threadIdx.x; j < D; j += blockDim.x) w_s[j] = variables[j];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < K; i += (long long)gridDim.x * blockDim.x) {
    const float* row = basis + i * D;
    float v = 0.f;
    for (int j = 0; j < D; ++j) v = __fadd_rn(v, __fmul_rn(w_s[j], row[j]));
    if (scale) v = __fmul_rn(scale[i], v);
    if (offset1) v = __fadd_rn(v, offset1[i]);
    if (offset2) v = __fadd_rn(v, offset2[i]);
    table[i] = v;
  }
}

// Per-sweep class tables of the Glauber sweep on the device -- the arithmetic of noneq_opt_b200.ising.class_tables
// (itself flip_logits + distrax.Bernoulli, ising.py:93-96, 114-117): float32 operations in the reference's order,
// transcendentals evaluated in binary64 and rounded once to float32.  One thread per (sweep, class).
//   thr[t][slot] = ceil(sigmoid32(logit) * 2^23);  lut[t][0..3][slot] = logit, log_sigmoid32(-logit),
//   log_sigmoid32(logit), sigmoid(logit) in binary64;  slot = 8 * (spin up) + #up-neighbours
__device__ __forceinline__ float exp32(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_sigmoid32(float x) {
  const float a = -x;
  const float l1p = (float)log1p((double)exp32(-fabsf(a)));
  return -__fadd_rn(fmaxf(a, 0.f), l1p);
}
__global__ void ising_class_tables_kernel(const float* __restrict__ log_temp, const float* __restrict__ field, int T,
                                          int ndim, uint32_t* __restrict__ thr, double* __restrict__ lut) {
  const int ncls = 2 * ndim + 1;
  const long long n = (long long)(T - 1) * 16;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(q / 16), slot = (int)(q % 16), up = slot >> 3, n_up = slot & 7;
    uint32_t th = 0u;
    double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
    if (n_up < ncls) {
      const float lt = log_temp[t + 1], h = field[t + 1];
      const float einv = exp32(-lt);
      const float sp = up ? 1.0f : -1.0f;
      const float nsum = (float)(2 * n_up - 2 * ndim);
      const float logit = __fmul_rn(__fmul_rn(__fmul_rn(-2.0f, sp), __fadd_rn(nsum, h)), einv);
      const float prob = __fdiv_rn(1.0f, __fadd_rn(1.0f, exp32(-logit)));
      th = (uint32_t)ceil((double)prob * 8388608.0);
      l0 = (double)logit;
      l1 = (double)log_sigmoid32(-logit);
      l2 = (double)log_sigmoid32(logit);
      l3 = 1.0 / (1.0 + exp(-(double)logit));
    }
    thr[(size_t)t * 16 + slot] = th;
    double* L = lut + (size_t)t * 4 * 16;
    L[slot] = l0; L[16 + slot] = l1; L[32 + slot] = l2; L[48 + slot] = l3;
  }
}

// mean over the batch of the per-replica gradient (w_logp * logP + w_rep * reparam), replica terms rounded to
// float32 first (what vmap(grad) returns) and averaged in binary64 (fixed order) -> float32 [Dl + Dh]
__global__ void __launch_bounds__(256) ising_grad_mean_kernel(const double* __restrict__ logP_lt, const double* __restrict__ logP_h,
                                                              const double* __restrict__ rep_lt, const double* __restrict__ rep_h,
                                                              long long R, int Dl, int Dh, double w_logp, double w_rep,
                                                              float* __restrict__ out) {
  __shared__ double red_s[8];
  const int j = blockIdx.x;                       // one CTA per variable
  const bool is_h = j >= Dl;
  const int c = is_h ? j - Dl : j, D = is_h ? Dh : Dl;
  const double* a = is_h ? logP_h : logP_lt;
  const double* b = is_h ? rep_h : rep_lt;
  double s = 0.0;
  for (long long r = threadIdx.x; r < R; r += blockDim.x)
    s += (double)(float)(w_logp * a[r * D + c] + w_rep * b[r * D + c]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red_s[w];
    out[j] = (float)(tot / (double)R);
  }
}

struct AdamArgs {
  double step_size, decay_steps, decay_rate, b1, b2, eps, max_norm, w0, w1, denom;
  int P;
  const double* gs0;
  const double* gs1;
  const float* g32;
  float* x;
  float* m;
  float* v;
  long long* step;
  float* grad_out;
};

// One CTA.  g = float((w0 gs0 + w1 gs1) / denom)  [or g32]; clip_grads (global l2 norm, float64);
// jax.example_libraries.optimizers.adam in binary32 with one rounding per operation; ++*step.
__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a) {
  __shared__ double red_s[256];
  __shared__ float scale_s;
  const long long i = *a.step;
  double sq = 0.0;
  for (int j = threadIdx.x; j < a.P; j += blockDim.x) {
    const float g = a.g32 ? a.g32[j] : (float)((a.w0 * a.gs0[j] + (a.gs1 ? a.w1 * a.gs1[j] : 0.0)) / a.denom);
    sq += (double)g * (double)g;
  }
  red_s[threadIdx.x] = sq;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red_s[threadIdx.x] += red_s[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double norm = sqrt(red_s[0]);
    scale_s = (a.max_norm > 0.0 && !(norm < a.max_norm)) ? (float)(a.max_norm / norm) : 1.0f;
  }
  __syncthreads();
  const bool clip = scale_s != 1.0f;
  const double lr64 = a.decay_steps > 0.0 ? a.step_size * pow(a.decay_rate, (double)i / a.decay_steps) : a.step_size;
  const float lr = (float)lr64;
  const float c1 = (float)(1.0 - a.b1), c2 = (float)(1.0 - a.b2), b1 = (float)a.b1, b2 = (float)a.b2;
  const float d1 = (float)(1.0 - pow(a.b1, (double)(i + 1))), d2 = (float)(1.0 - pow(a.b2, (double)(i + 1)));
  const float eps = (float)a.eps;
  for (int j = threadIdx.x; j < a.P; j += blockDim.x) {
    float g = a.g32 ? a.g32[j] : (float)((a.w0 * a.gs0[j] + (a.gs1 ? a.w1 * a.gs1[j] : 0.0)) / a.denom);
    if (clip) g = __fmul_rn(g, scale_s);
    const float m = __fadd_rn(__fmul_rn(c1, g), __fmul_rn(b1, a.m[j]));
    const float v = __fadd_rn(__fmul_rn(c2, __fmul_rn(g, g)), __fmul_rn(b2, a.v[j]));
    const float mhat = __fdiv_rn(m, d1), vhat = __fdiv_rn(v, d2);
    a.x[j] = __fsub_rn(a.x[j], __fdiv_rn(__fmul_rn(lr, mhat), __fadd_rn(__fsqrt_rn(vhat), eps)));
    a.m[j] = m;
    a.v[j] = v;
    if (a.grad_out) a.grad_out[j] = g;
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.step = i + 1;
}

}  // namespace nq

using namespace nq;

extern "C" int nq_random_split_dev(const uint32_t* key_dev, int64_t num, int64_t first, int64_t count,
                                   uint32_t* out, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(key_dev && out, "null pointer");
  NQ_CHECK_ARG(num >= 0 && first >= 0 && count >= 0 && first + count <= num, "bad slice [%lld,+%lld) of %lld",
               (long long)first, (long long)count, (long long)num);
  NQ_CHECK_ARG(2 * num <= 0xFFFFFFFFll, "size exceeds the 32-bit counter range of the original threefry layout");
  NQ_CHECK_ARG(out + 2 * count <= key_dev || key_dev + 2 <= out, "out must not alias the key");
  if (count == 0) return NQ_OK;
  const long long words = 2 * count, blocks = (words + 255) / 256;
  random_bits_devkey_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, (cudaStream_t)stream>>>(
      key_dev, 2ull * num, 2ull * first, (unsigned long long)words, out);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_schedule_affine(const float* basis, const float* variables, const float* offset, float first,
                                  float last, int64_t K, int32_t D, float* table, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(basis && variables && table, "null pointer");
  NQ_CHECK_ARG(K >= 0 && D >= 1 && D <= 4096, "need K >= 0 and 1 <= D <= 4096 (got K=%lld D=%d)", (long long)K, D);
  const long long blocks = (K + 2 + 255) / 256;
  schedule_affine_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, D * sizeof(float),
                           (cudaStream_t)stream>>>(basis, variables, offset, first, last, K, D, table);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_adam_step(const NqAdamDesc* d, int32_t P, const double* grad_sums0, const double* grad_sums1,
                            const float* grad_f32, float* x, float* m, float* v, int64_t* step_dev,
                            float* grad_out, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(d && x && m && v && step_dev && P >= 1, "null pointer or P < 1");
  NQ_CHECK_ARG((grad_sums0 != nullptr) != (grad_f32 != nullptr), "pass either grad_sums0 (f64 sums) or grad_f32");
  NQ_CHECK_ARG(grad_f32 || d->denom != 0.0, "denom must be non-zero");
  NQ_CHECK_ARG(d->b1 >= 0 && d->b1 < 1 && d->b2 >= 0 && d->b2 < 1, "b1, b2 must lie in [0, 1)");
  AdamArgs a{d->step_size, d->decay_steps, d->decay_rate, d->b1, d->b2, d->eps, d->max_norm, d->w0, d->w1, d->denom,
             P, grad_sums0, grad_sums1, grad_f32, x, m, v, (long long*)step_dev, grad_out};
  adam_step_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_schedule_compose(const float* basis, const float* variables, const float* scale, const float* offset1,
                                   const float* offset2, int64_t K, int32_t D, float* table, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(basis && variables && table, "null pointer");
  NQ_CHECK_ARG(K >= 1 && D >= 1 && D <= 4096, "need K >= 1 and 1 <= D <= 4096 (got K=%lld D=%d)", (long long)K, D);
  const long long blocks = (K + 255) / 256;
  schedule_compose_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, D * sizeof(float), (cudaStream_t)stream>>>(
      basis, variables, scale, offset1, offset2, K, D, table);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_class_tables(const float* log_temp, const float* field, int32_t T, int32_t ndim, uint32_t* thr,
                                     double* lut, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(log_temp && field && thr && lut, "null pointer");
  NQ_CHECK_ARG(T >= 2 && ndim >= 1 && ndim <= 3, "need T >= 2 and 1 <= ndim <= 3");
  const long long n = (long long)(T - 1) * 16, blocks = (n + 127) / 128;
  ising_class_tables_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 128, 0, (cudaStream_t)stream>>>(
      log_temp, field, T, ndim, thr, lut);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_grad_mean(const double* logP_lt, const double* logP_h, const double* rep_lt, const double* rep_h,
                                  int64_t R, int32_t Dl, int32_t Dh, double w_logp, double w_rep, float* mean_out,
                                  void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(logP_lt && logP_h && rep_lt && rep_h && mean_out, "null pointer");
  NQ_CHECK_ARG(R >= 1 && Dl >= 1 && Dh >= 1, "need R, Dl, Dh >= 1");
  ising_grad_mean_kernel<<<(unsigned)(Dl + Dh), 256, 0, (cudaStream_t)stream>>>(logP_lt, logP_h, rep_lt, rep_h, R, Dl, Dh,
                                                                              w_logp, w_rep, mean_out);
  NQ_LAUNCHED();
  return NQ_OK;
}
