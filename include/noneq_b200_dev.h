/* Developer tools of the noneq_opt B200 build (libnoneq_b200_dev.so).  NOT part of the product ABI
 * (include/noneq_b200.h): nothing under noneq_opt_b200/*.py calls these; scripts/ and tests/ do. */
#ifndef NONEQ_B200_DEV_H_
#define NONEQ_B200_DEV_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* issue-rate microbenchmark: runs `iters` dependent-chain iterations of
 * instruction mix `which` (0..11; 12 = warp placement probe) on `blocks` x `threads`; elapsed SM
 * cycles per block -> out_cycles.                                                             */
int nqdev_microbench(int32_t which, int32_t blocks, int32_t threads, int32_t iters,
                     uint64_t* out_cycles /*device [blocks]*/, uint32_t* sink /*device*/, void* stream);

/* bit-identity of the STRICT main-path math against the CUDA library functions over the input bit
 * patterns [first, first + count):  which = 0 logf_main vs logf, 1 log1pf_main(-x) vs log1pf(-x),
 * 2 div_main(c, x) vs IEEE c / x, 3 normal_from_bits(p << 9) vs the literal jax.random.normal
 * transcription, 4 expf_main vs expf, 5 expf_main2({x, -x}) vs {expf(x), expf(-x)}, 6 the packed landscape's
 * A, B, A + B at x vs the scalar reference order (c = beta dE).  out[0] = number of mismatches, out[1] = first mismatching pattern.            */
int nqdev_math_check(int32_t which, uint64_t first, uint64_t count, float c, uint64_t* out /*device [2]*/,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NONEQ_B200_DEV_H_ */
