// TEMPORARY: entry points declared in include/noneq_b200.h that are not implemented yet.
#include "nq_common.cuh"
using namespace nq;
#define NQ_STUB(name) { set_error(name " is not implemented yet"); return NQ_ERR_UNSUPPORTED; }
extern "C" size_t nq_langevin_ckpt_count(int64_t steps, int64_t every) { return every > 0 ? (size_t)((steps + every - 1) / every) : 0; }
extern "C" int nq_langevin_forward_ckpt(const NqLangevinDesc*, const float*, const float*, const uint32_t*, int64_t, int64_t, float*, float*, float*, float*, uint32_t*, double*, void*, size_t, void*) NQ_STUB("nq_langevin_forward_ckpt")
extern "C" int nq_langevin_replay(const NqLangevinDesc*, const float*, int64_t, int64_t, const float*, const uint32_t*, const float*, const float*, double*, void*, size_t, void*) NQ_STUB("nq_langevin_replay")
