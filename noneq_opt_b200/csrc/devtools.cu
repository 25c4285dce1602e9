// Developer tools, built into libnoneq_b200_dev.so (NOT part of the product ABI, include/noneq_b200_dev.h):
//   nqdev_microbench  -- issue-rate calibration for the instruction mixes the hot path is built from
//                        ("first kernel PR must include a microbenchmark");
//   nqdev_math_check  -- bit-identity of the STRICT main-path math (nq_common.cuh: logf_main, log1pf_main,
//                        div_main, normal_from_bits) against the CUDA library functions / the literal
//                        transcription of jax.random.normal, over ranges of input bit patterns.
//
// Each thread runs 8 independent dependency chains; one loop iteration issues kOps instructions
// of the mix per chain-slot.  The host divides (threads * iters * ops_per_iter) by the elapsed SM
// cycles to get thread-instructions / clk / SM for that mix.
#include "langevin_impl.cuh"   // (brings nq_common.cuh; for the landscape check)

namespace nq {
// self-contained: the dev library does not link core.cu
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return NQ_ERR_CUDA; }
std::atomic<uint64_t> g_launches{0};

constexpr int kChains = 8;

template <int WHICH>
__device__ __forceinline__ void body(uint32_t (&a)[kChains], uint32_t (&b)[kChains], float (&f)[kChains]) {
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    if (WHICH == 0) {  // LOP3 (alu)
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
    } else if (WHICH == 1) {  // SHF (alu)
      asm volatile("shf.l.wrap.b32 %0, %0, %0, 13;" : "+r"(a[c]));
    } else if (WHICH == 2) {  // integer add (ptxas picks IADD3 or IMAD.IADD)
      asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
    } else if (WHICH == 3) {  // IMAD (fma pipe)
      asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(b[c]));
    } else if (WHICH == 4) {  // FFMA
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(__uint_as_float(b[c])));
    } else if (WHICH == 5) {  // IMAD.WIDE
      uint64_t w;
      asm volatile("mul.wide.u32 %0, %1, 8192;" : "=l"(w) : "r"(a[c]));
      a[c] = (uint32_t)w | (uint32_t)(w >> 32);
    } else if (WHICH == 6) {  // one threefry round: add, rotate, xor
      a[c] += b[c];
      b[c] = rotl32(b[c], 13);
      b[c] ^= a[c];
    } else if (WHICH == 7) {  // MUFU.EX2
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
    } else if (WHICH == 8) {  // threefry round with the rotate done as IMAD.WIDE + LOP3
      a[c] += b[c];
      uint64_t w;
      asm volatile("mul.wide.u32 %0, %1, 8192;" : "=l"(w) : "r"(b[c]));
      b[c] = ((uint32_t)w | (uint32_t)(w >> 32)) ^ a[c];
    } else if (WHICH == 9) {  // LOP3 + FFMA pair (dual-pipe issue)
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(__uint_as_float(b[c])));
    } else if (WHICH == 10) {  // LOP3 + IMAD pair
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
      asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[c]) : "r"(a[c]));
    } else if (WHICH == 11) {  // full threefry2x32-20 block per chain (73 ops nominal)
      uint32_t y0, y1;
      threefry2x32(key_schedule(a[c], b[c]), a[c], b[c], y0, y1);
      a[c] = y0;
      b[c] = y1;
    }
  }
}

template <int WHICH>
__global__ void microbench_kernel(int iters, unsigned long long* out_cycles, uint32_t* sink) {
  uint32_t a[kChains], b[kChains];
  float f[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    a[c] = threadIdx.x * 2654435761u + c;
    b[c] = blockIdx.x * 40503u + c * 7u + 1u;
    f[c] = 1.0f + 1e-3f * (float)((threadIdx.x + c) & 7);
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) body<WHICH>(a, b, f);
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc ^= a[c] ^ b[c] ^ __float_as_uint(f[c]);
  if (acc == 0x12345u) sink[0] = acc;
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

}  // namespace nq

using namespace nq;

// which == 12: where the hardware places warps.  out[block] = smid | warp slot of warp 0 << 16 | warp slot of
// warp 1 << 24 (warp slot % 4 = SM sub-partition); the kernel spins `iters` x 1000 cycles so that the whole grid is
// resident at once, with the register footprint of the Langevin kernel (80 registers, 64-thread CTAs).
__global__ void __launch_bounds__(128) placement_kernel(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ unsigned slots[32];
  unsigned smid, warpid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
  if ((threadIdx.x & 31) == 0) slots[threadIdx.x >> 5] = warpid;
  __syncthreads();
  const long long t0 = clock64();
  while (clock64() - t0 < 1000ll * iters) { }
  if (threadIdx.x == 0) {
    unsigned long long v = smid;
    for (unsigned w = 0; w < blockDim.x / 32 && w < 6; ++w) v |= (unsigned long long)(slots[w] & 0xff) << (16 + 8 * w);
    out[blockIdx.x] = v;
  }
  if (iters < 0) sink[0] = smid;
}

extern "C" int nqdev_microbench(int32_t which, int32_t blocks, int32_t threads, int32_t iters, uint64_t* out_cycles,
                             uint32_t* sink, void* stream) {
  NQ_CHECK_ARG(which >= 0 && which <= 12, "unknown microbenchmark %d", which);
  NQ_CHECK_ARG(blocks >= 1 && threads >= 32 && threads <= 1024 && iters >= 1 && out_cycles && sink, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* oc = (unsigned long long*)out_cycles;
#define NQ_MB(W) case W: microbench_kernel<W><<<blocks, threads, 0, st>>>(iters, oc, sink); break;
  if (which == 12) {
    placement_kernel<<<blocks, threads, 0, st>>>(iters, oc, sink);
    NQ_LAUNCHED();
    return NQ_OK;
  }
  switch (which) {
    NQ_MB(0) NQ_MB(1) NQ_MB(2) NQ_MB(3) NQ_MB(4) NQ_MB(5) NQ_MB(6) NQ_MB(7) NQ_MB(8) NQ_MB(9) NQ_MB(10) NQ_MB(11)
  }
#undef NQ_MB
  NQ_LAUNCHED();
  return NQ_OK;
}

// ------------------------------------------------------------------------------ math check
namespace nq {

// jax.random.normal from a raw word exactly as jax writes it, on the CUDA library
// functions: the form the STRICT kernels used before the main-path rewrite
__device__ float normal_literal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));
  float w = -log1pf(-__fmul_rn(u, u));
  float p;
#define NQ_H(c) p = __fadd_rn(c, __fmul_rn(p, w));
  if (w < 5.0f) {
    w = __fadd_rn(w, -2.5f);
    p = 2.81022636e-08f;
    NQ_H(3.43273939e-07f) NQ_H(-3.5233877e-06f) NQ_H(-4.39150654e-06f) NQ_H(0.00021858087f)
    NQ_H(-0.00125372503f) NQ_H(-0.00417768164f) NQ_H(0.246640727f) NQ_H(1.50140941f)
  } else {
    w = __fadd_rn(__fsqrt_rn(w), -3.0f);
    p = -0.000200214257f;
    NQ_H(0.000100950558f) NQ_H(0.00134934322f) NQ_H(-0.00367342844f) NQ_H(0.00573950773f)
    NQ_H(-0.0076224613f) NQ_H(0.00943887047f) NQ_H(1.00167406f) NQ_H(2.83297682f)
  }
#undef NQ_H
  const float r = __fmul_rn(p, u);
  return __fmul_rn(1.41421356237309504880f, fabsf(u) == 1.0f ? u * __int_as_float(0x7f800000) : r);
}

// which: 0 logf_main(x) vs logf(x); 1 log1pf_main(-x) vs log1pf(-x); 2 div_main(c, x) vs c / x (IEEE);
//        3 normal_from_bits<false>(p << 9) vs normal_literal(p << 9), p = the bit pattern itself;
//        4 expf_main(x) vs expf(x); 5 expf_main2({x, -x}) vs {expf(x), expf(-x)};
//        6 the packed landscape's A, B, A + B vs the scalar reference order (c = beta dE).
// x = the float with bit pattern p, p in [first, first + count).  out[0] = mismatches, out[1] = first mismatching p.
__global__ void math_check_kernel(int which, unsigned long long first, unsigned long long count, float c,
                                  unsigned long long* out) {
  unsigned long long bad = 0, where = ~0ull;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t p = (uint32_t)(first + i);
    const float x = __uint_as_float(p);
    float a, b;
    if (which == 0) { a = logf_main(x); b = logf(x); }
    else if (which == 1) { a = log1pf_main(-x); b = log1pf(-x); }
    else if (which == 2) { a = div_main(c, x); b = __fdiv_rn(c, x); }
    else if (which == 3) { a = normal_from_bits<false>(p << 9); b = normal_literal(p << 9); }
    else if (which == 4) { const StrictRegs sr; a = expf_main(x, sr.exp_k1, sr.exp_k252); b = expf(x); }
    else if (which == 6) {
      // the packed landscape (Landscape::eval: both wells as f32x2 pairs) against the scalar reference order
      // A = exp(cl ul^2), B = exp(-(cr ur^2 + bde)) with the library expf, every operation rounded separately; x is
      // the pattern's float folded into [-64, 64), c = beta dE (a non-zero value catches a contracted multiply-add)
      LangevinConsts k{};
      const StrictRegs sr;
      k.cl = -0.0543f; k.cr = 0.0571f; k.bde = c; k.two_xm = 28.0f; k.inv_beta = 4.183f;
      k.exp_k1 = sr.exp_k1; k.exp_k252 = sr.exp_k252;
      k.well_c2 = f2_pack(k.cl, -k.cr);
      k.guard_span = 0x3e000000u; k.guard_lo = 0x1p-62f;
      const float xx = fmodf(x, 64.0f);
      Landscape<NQ_POT_BIOMOLECULE_SHIFTED, false> land;
      land.template eval<false>(k, xx);
      const float ur = __fsub_rn(xx, k.two_xm);
      const float A = expf(__fmul_rn(k.cl, __fmul_rn(xx, xx)));
      const float B = expf(-__fadd_rn(__fmul_rn(k.cr, __fmul_rn(ur, ur)), k.bde));
      a = land.A; b = A;
      if (__float_as_uint(land.B) != __float_as_uint(B) || __float_as_uint(land.sAB) != __float_as_uint(__fadd_rn(A, B)))
        a = __uint_as_float(~__float_as_uint(b));
      if (!(fabsf(x) < 1e30f)) a = b;   // NaN / inf patterns: fmodf is not the subject
    }
    else {   // both halves of the packed form: lo = exp(x), hi = exp(-x)
      const StrictRegs sr;
      float hi;
      f2_unpack(expf_main2(f2_pack(x, -x), sr.exp_k1, sr.exp_k252), a, hi);
      b = expf(x);
      if (__float_as_uint(hi) != __float_as_uint(expf(-x))) a = __uint_as_float(~__float_as_uint(b));
    }
    if (__float_as_uint(a) != __float_as_uint(b)) {
      ++bad;
      if (first + i < where) where = first + i;
    }
  }
  if (bad) {
    atomicAdd(out, bad);
    atomicMin(out + 1, where);
  }
}

}  // namespace nq

extern "C" int nqdev_math_check(int32_t which, uint64_t first, uint64_t count, float c, uint64_t* out /*device [2]*/,
                                void* stream) {
  if (which < 0 || which > 6 || !out) return NQ_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long init[2] = {0ull, ~0ull};
  if (cudaMemcpyAsync(out, init, sizeof(init), cudaMemcpyHostToDevice, st) != cudaSuccess) return NQ_ERR_CUDA;
  math_check_kernel<<<148 * 8, 256, 0, st>>>(which, first, count, c, (unsigned long long*)out);
  return cudaGetLastError() == cudaSuccess ? NQ_OK : NQ_ERR_CUDA;
}
