// Instantiations of the Langevin kernels for potential NQ_POT_SPRING (one translation unit per potential so
// that the template instantiations compile in parallel).
#include "langevin_impl.cuh"

namespace nq {
NQ_DEFINE_POT_LAUNCHERS(NQ_POT_SPRING, pot_spring)
}  // namespace nq
