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
