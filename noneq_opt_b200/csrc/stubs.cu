// TEMPORARY: entry points declared in include/noneq_b200.h that are not implemented yet.
#include "nq_common.cuh"
using namespace nq;
#define NQ_STUB(name) { set_error(name " is not implemented yet"); return NQ_ERR_UNSUPPORTED; }
extern "C" size_t nq_langevin_ckpt_count(int64_t steps, int64_t every) { return every > 0 ? (size_t)((steps + every - 1) / every) : 0; }
extern "C" int nq_langevin_forward_ckpt(const NqLangevinDesc*, const float*, const float*, const uint32_t*, int64_t, int64_t, float*, float*, float*, float*, uint32_t*, double*, void*, size_t, void*) NQ_STUB("nq_langevin_forward_ckpt")
extern "C" int nq_langevin_replay(const NqLangevinDesc*, const float*, int64_t, int64_t, const float*, const uint32_t*, const float*, const float*, double*, void*, size_t, void*) NQ_STUB("nq_langevin_replay")
extern "C" size_t nq_ising_words_per_lattice(int32_t L) { return (size_t)L * L / 32; }
extern "C" size_t nq_ising_workspace_bytes(const NqIsingDesc*) { return 0; }
extern "C" int nq_ising_run(const NqIsingDesc*, const uint32_t*, const double*, const float*, const float*, const uint32_t*, const uint32_t*, uint32_t*, float*, float*, int32_t*, uint32_t*, void*, size_t, void*) NQ_STUB("nq_ising_run")
extern "C" int nq_ising_random_spins(const uint32_t*, int32_t, float, uint32_t*, void*) NQ_STUB("nq_ising_random_spins")
extern "C" int nq_ising_pack(const void*, int32_t, int64_t, uint32_t*, void*) NQ_STUB("nq_ising_pack")
extern "C" int nq_ising_unpack(const uint32_t*, int64_t, int32_t, void*, void*) NQ_STUB("nq_ising_unpack")
extern "C" int nq_ising_measure(const uint32_t*, int32_t, int64_t, int64_t*, void*) NQ_STUB("nq_ising_measure")
