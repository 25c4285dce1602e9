// Library plumbing (errors, version, launch counter) and the jax.random-compatible
// key / bits / uniform / normal generators.
#include <string.h>

#include <dlfcn.h>

#include "nq_common.cuh"

namespace nq {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return NQ_ERR_CUDA;
}

// out[i] = element (first+i) of threefry_2x32(key, iota(size)), optionally transformed
enum { kRawBits = 0, kUniform = 1, kNormal = 2 };

template <int MODE>
__global__ void random_bits_kernel(uint32_t k0, uint32_t k1, unsigned long long size, unsigned long long first,
                                   unsigned long long count, float lo, float hi, void* out) {
  const KeySchedule ks = key_schedule(k0, k1);
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t bits = random_bits_at(ks, first + i, size);
    if (MODE == kRawBits) {
      ((uint32_t*)out)[i] = bits;
    } else if (MODE == kUniform) {
      // max(lo, floats * (hi - lo) + lo)     jax/_src/random.py _uniform
      const float f = bits_to_unit(bits);
      ((float*)out)[i] = fmaxf(lo, __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo));
    } else {
      ((float*)out)[i] = normal_from_bits<false>(bits);
    }
  }
}

template <int MODE>
static int launch_bits(const uint32_t key[2], int64_t size, int64_t first, int64_t count, float lo, float hi,
                       void* out, void* stream) {
  NQ_CHECK_ARG(key && out, "null pointer");
  NQ_CHECK_ARG(size >= 0 && first >= 0 && count >= 0 && first + count <= size, "bad slice [%lld,+%lld) of %lld",
               (long long)first, (long long)count, (long long)size);
  NQ_CHECK_ARG(size <= 0xFFFFFFFFll, "size exceeds the 32-bit counter range of the original threefry layout");
  if (count == 0) return NQ_OK;
  const int threads = 256;
  const long long blocks = (count + threads - 1) / threads;
  random_bits_kernel<MODE><<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), threads, 0, (cudaStream_t)stream>>>(
      key[0], key[1], (unsigned long long)size, (unsigned long long)first, (unsigned long long)count, lo, hi, out);
  NQ_LAUNCHED();
  return NQ_OK;
}

}  // namespace nq

using namespace nq;

extern "C" int nq_version(void) { return NQ_VERSION; }
extern "C" const char* nq_last_error(void) { return g_err; }
extern "C" uint64_t nq_launch_count(void) { return g_launches.load(); }

// ncclAllReduce(sendbuf, recvbuf, count, ncclFloat64 = 8, ncclSum = 0, comm, stream), looked up once
typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static NcclAllReduceFn find_nccl_allreduce() {
  static std::atomic<NcclAllReduceFn> cached{nullptr};
  NcclAllReduceFn fn = cached.load();
  if (fn) return fn;
  void* h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_NOLOAD);  // the copy torch (or the caller) already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_LAZY);
  if (!h) h = dlopen("libnccl.so", RTLD_LAZY);
  if (h) fn = (NcclAllReduceFn)dlsym(h, "ncclAllReduce");
  if (fn) cached.store(fn);
  return fn;
}

extern "C" int nq_allreduce_f64(void* nccl_comm, double* buf, int64_t n, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(nccl_comm != nullptr && buf != nullptr && n >= 1, "need a communicator, a buffer and n >= 1");
  NcclAllReduceFn fn = find_nccl_allreduce();
  if (!fn) {
    const char* why = dlerror();   // NULL once any dlopen succeeded: then the library loaded but lacks the symbol
    set_error("ncclAllReduce is not available in this process: %s", why ? why : "libnccl loaded, symbol ncclAllReduce missing");
    return NQ_ERR_CUDA;
  }
  const int rc = fn(buf, buf, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) {
    set_error("ncclAllReduce failed with ncclResult_t %d", rc);
    return NQ_ERR_CUDA;
  }
  return NQ_OK;
}

extern "C" int nq_random_split(const uint32_t key[2], int64_t num, int64_t first, int64_t count, uint32_t* out,
                               void* stream) {
  NQ_RANGE();
  // split(key, num) = threefry_2x32(key, iota(2*num)).reshape(num, 2)
  return launch_bits<kRawBits>(key, 2 * num, 2 * first, 2 * count, 0.f, 1.f, out, stream);
}

extern "C" int nq_langevin_keys(const uint32_t seed[2], int64_t B_total, int64_t n0, int64_t n, uint32_t* keys,
                                void* stream) {
  NQ_RANGE();
  return nq_random_split(seed, B_total, n0, n, keys, stream);
}

extern "C" int nq_random_bits(const uint32_t key[2], int64_t size, int64_t first, int64_t count, uint32_t* out,
                              void* stream) {
  NQ_RANGE();
  return launch_bits<kRawBits>(key, size, first, count, 0.f, 1.f, out, stream);
}

extern "C" int nq_random_uniform(const uint32_t key[2], int64_t size, int64_t first, int64_t count, float lo,
                                 float hi, float* out, void* stream) {
  NQ_RANGE();
  return launch_bits<kUniform>(key, size, first, count, lo, hi, out, stream);
}

extern "C" int nq_random_normal(const uint32_t key[2], int64_t size, int64_t first, int64_t count, float* out,
                                void* stream) {
  NQ_RANGE();
  return launch_bits<kNormal>(key, size, first, count, 0.f, 1.f, out, stream);
}
