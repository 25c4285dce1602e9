/* A plain C client of libnoneq_b200.so: no Python, no torch -- cudaMalloc'd buffers and the C ABI of
 * include/noneq_b200.h only.  It runs config 1 of BASELINE.json (1000 trajectories x (100 + 1000) steps in
 * a moving harmonic trap, linear protocol) forward + gradient and prints the ensemble moments and the
 * gradient of <W> with respect to the first protocol coefficients.
 *
 *   nvcc examples/c_client.c -Iinclude -Lnoneq_opt_b200 -lnoneq_b200 -Xlinker -rpath=$PWD/noneq_opt_b200 -o c_client
 *
 * The trap table and its Jacobian are supplied by the caller: here a straight line r(k) and the two basis
 * columns {1, t} of that line (any protocol works: the library contracts d/dr_k with whatever phi it is given).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "noneq_b200.h"

#define CK(call)                                                                 \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define NQ(call)                                                                 \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != NQ_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, nq_last_error()); return 3; } \
  } while (0)

int main(void) {
  enum { B = 1000, S = 1000, NEQ = 100, D = 2 };
  const float r0 = 5.f, r1 = 10.f;
  NqLangevinDesc desc = {0};
  desc.potential = NQ_POT_SPRING;
  desc.shift_kind = NQ_SHIFT_FREE;
  desc.math_mode = NQ_MATH_STRICT;
  desc.k_s = 1.0; desc.beta = 1.0;
  desc.dt = 1e-3; desc.kT = 1.0; desc.gamma = 1.0; desc.mass = 1.0;
  desc.steps = S; desc.eq_steps = NEQ;

  /* host side of the protocol: r[0] = r0, r[k] = a + b t_k (k = 1..S-1, t_k = (k-1)/(S-2)), r[S] = r1 */
  float* r_h = (float*)malloc((S + 1) * sizeof(float));
  float* phi_h = (float*)malloc((size_t)(S - 1) * D * sizeof(float));
  const float a = r0 + (r1 - r0) / S, b = (r1 - r0) * (S - 2) / S;
  r_h[0] = r0; r_h[S] = r1;
  for (int k = 1; k < S; ++k) {
    const float t = (float)(k - 1) / (float)(S - 2);
    r_h[k] = a + b * t;
    phi_h[(k - 1) * D + 0] = 1.f;
    phi_h[(k - 1) * D + 1] = t;
  }
  float* x0_h = (float*)malloc(B * sizeof(float));
  for (int i = 0; i < B; ++i) x0_h[i] = r0;

  float *r_d, *phi_d, *x0_d, *work_d, *logp_d, *est_d, *xf_d;
  uint32_t* seeds_d;
  double *grad_d, *mom_d;
  void* ws_d;
  const size_t ws_bytes = nq_langevin_workspace_bytes(B, D);
  CK(cudaMalloc((void**)&r_d, (S + 1) * sizeof(float)));
  CK(cudaMalloc((void**)&phi_d, (size_t)(S - 1) * D * sizeof(float)));
  CK(cudaMalloc((void**)&x0_d, B * sizeof(float)));
  CK(cudaMalloc((void**)&work_d, B * sizeof(float)));
  CK(cudaMalloc((void**)&logp_d, B * sizeof(float)));
  CK(cudaMalloc((void**)&est_d, B * sizeof(float)));
  CK(cudaMalloc((void**)&xf_d, B * sizeof(float)));
  CK(cudaMalloc((void**)&seeds_d, (size_t)B * 2 * sizeof(uint32_t)));
  CK(cudaMalloc((void**)&grad_d, 2 * D * sizeof(double)));
  CK(cudaMalloc((void**)&mom_d, NQ_MOM_COUNT * sizeof(double)));
  CK(cudaMalloc(&ws_d, ws_bytes));
  CK(cudaMemcpy(r_d, r_h, (S + 1) * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(phi_d, phi_h, (size_t)(S - 1) * D * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(x0_d, x0_h, B * sizeof(float), cudaMemcpyHostToDevice));

  const uint32_t key[2] = {0u, 0u}; /* jax.random.PRNGKey(0) */
  NQ(nq_langevin_keys(key, B, 0, B, seeds_d, NULL));
  NQ(nq_langevin_grad(&desc, r_d, phi_d, D, x0_d, seeds_d, B, work_d, logp_d, est_d, xf_d, grad_d, mom_d, ws_d, ws_bytes,
                      NULL));
  CK(cudaDeviceSynchronize());
  double mom[NQ_MOM_COUNT], grad[2 * D];
  CK(cudaMemcpy(mom, mom_d, sizeof(mom), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(grad, grad_d, sizeof(grad), cudaMemcpyDeviceToHost));
  printf("n %.0f mean_work %.9f mean_logp %.6f\n", mom[NQ_MOM_N], mom[NQ_MOM_W] / mom[NQ_MOM_N], mom[NQ_MOM_LOGP] / mom[NQ_MOM_N]);
  for (int j = 0; j < D; ++j)
    printf("grad[%d] logP-term %.9e reparam-term %.9e\n", j, grad[j] / B, grad[D + j] / B);
  printf("kernels launched %llu, library version %d\n", (unsigned long long)nq_launch_count(), nq_version());
  return 0;
}
