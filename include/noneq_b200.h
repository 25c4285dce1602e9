/* noneq_b200.h -- C ABI of libnoneq_b200.so (sm_100a), the drop-in boundary of the
 * noneq_opt ensemble hot path.
 *
 * The reference (mc2engel/noneq_opt) has no FFI: its hot path is reached through plain
 * Python functions which trace into one XLA computation.  Each entry point below replaces
 * the body of one of those functions (cited as file:line relative to the reference root);
 * the Python package `noneq_opt_b200` keeps the reference signatures and calls these
 * through ctypes (see INTEGRATION.md for the binding a maintainer of the reference adds).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked "host";
 *     the library never allocates or frees caller-visible memory (workspaces are sized
 *     by nq_*_workspace_bytes and passed in);
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising;
 *   - return 0 on success, a negative NQ_ERR_* otherwise; nq_last_error() returns a
 *     thread-local message for the last failure; no exceptions cross the ABI;
 *   - re-entrant: no global mutable state apart from the thread-local error string;
 *   - keys are raw uint32[2] threefry keys exactly as jax.random.PRNGKey / split
 *     produce them (original, non-partitionable counter layout).
 */
#ifndef NONEQ_B200_H_
#define NONEQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQ_VERSION 100

#define NQ_OK 0
#define NQ_ERR_INVALID_ARGUMENT (-1)
#define NQ_ERR_UNSUPPORTED (-2)
#define NQ_ERR_CUDA (-3)
#define NQ_ERR_WORKSPACE (-4)

int nq_version(void);
const char* nq_last_error(void);
/* number of kernels this library has launched since load (all threads); bench.py's gpu_launches */
uint64_t nq_launch_count(void);

/* ------------------------------------------------------------------ random (jax.random) */
/* jax.random.split(key, num)[first : first+count]  ->  out[count][2].
 * Replaces jax.random.split at barrier_crossing.py:221,253; ising.py:283. */
int nq_random_split(const uint32_t key[2] /*host*/, int64_t num, int64_t first, int64_t count,
                    uint32_t* out, void* stream);
/* The per-trajectory keys of a rank's shard: rows [n0, n0+n) of jax.random.split(seed, B_total)
 * (barrier_crossing.py:221; SURVEY 8b/8e: the counter pairing depends on the GLOBAL batch, so results
 * do not depend on the number of ranks).  Same as nq_random_split(seed, B_total, n0, n, ...).    */
int nq_langevin_keys(const uint32_t seed[2] /*host*/, int64_t B_total, int64_t n0, int64_t n,
                     uint32_t* keys /*[n][2]*/, void* stream);
/* jax.random.bits(key, (size,), uint32)[first : first+count] */
int nq_random_bits(const uint32_t key[2] /*host*/, int64_t size, int64_t first, int64_t count,
                   uint32_t* out, void* stream);
/* jax.random.uniform(key,(size,),f32,lo,hi) / jax.random.normal(key,(size,),f32) slices */
int nq_random_uniform(const uint32_t key[2] /*host*/, int64_t size, int64_t first, int64_t count,
                      float lo, float hi, float* out, void* stream);
int nq_random_normal(const uint32_t key[2] /*host*/, int64_t size, int64_t first, int64_t count,
                     float* out, void* stream);

/* ------------------------------------------------------------------ Langevin */
#define NQ_POT_SPRING 0              /* potentials.py:47-52   V_spring(k_s)               */
#define NQ_POT_BIOMOLECULE 1         /* potentials.py:27-35   V_biomolecule(...)          */
#define NQ_POT_BIOMOLECULE_SHIFTED 2 /* potentials.py:37-45   V_biomolecule_shifted(...)  */

#define NQ_SHIFT_FREE 0     /* jax_md.space.free():      R + dR            */
#define NQ_SHIFT_PERIODIC 1 /* jax_md.space.periodic(b): mod(R + dR, box)  */

#define NQ_MATH_STRICT 0 /* reference operation order, IEEE div/sqrt, accurate exp/log, no FMA */
#define NQ_MATH_FAST 1   /* MUFU exp2/log2/rcp, FMA contraction allowed                       */

typedef struct NqLangevinDesc {
  int32_t potential;  /* NQ_POT_*   */
  int32_t shift_kind; /* NQ_SHIFT_* */
  int32_t math_mode;  /* NQ_MATH_*  */
  int32_t reserved;
  /* potential parameters, python floats of the reference's factory call */
  double k_s, kappa_l, kappa_r, x_m, delta_E, beta;
  /* brownian(energy, shift, dt, kT, gamma) + init_fn(mass)  barrier_crossing.py:28-75 */
  double dt, kT, gamma, mass, box;
  int64_t steps;     /* simulation_steps S            barrier_crossing.py:97 */
  int64_t eq_steps;  /* Neq, at r0 = r_table[0]       barrier_crossing.py:115-121 */
} NqLangevinDesc;

/* moments[] layout written by the Langevin entry points (sums over the n trajectories of the call) */
#define NQ_MOM_N 0
#define NQ_MOM_W 1
#define NQ_MOM_W2 2
#define NQ_MOM_LOGP 3
#define NQ_MOM_WLOGP 4
#define NQ_MOM_EST 5
#define NQ_MOM_COUNT 8

size_t nq_langevin_workspace_bytes(int64_t n, int32_t D);

/* Builds the per-device normal-draw table of the STRICT kernels (jax.random.normal, jax/_src/random.py _normal_real:
 * sqrt(2) erf_inv(uniform(-1, 1)), reads 23 bits of each random word; the table holds its value for all 2^23
 * patterns, 32 MB, computed on the device by the same code the other kernels evaluate per draw).  Every STRICT
 * Langevin entry point builds it on first use; call this explicitly before CAPTURING such calls in a CUDA graph
 * (the build allocates and synchronises `stream`, which capture forbids).  Idempotent. */
int nq_langevin_prepare(void* stream);

/* run_brownian_opt (barrier_crossing.py:97-143) for n trajectories, given the trap table
 * r_table[S+1] (MakeSchedule_chebyshev :145-151, built by the host) and the per-trajectory
 * seeds (rows of jax.random.split(seed, B), :221).  The kernel performs the
 * `key, split = random.split(key)` of :133 itself.
 *   work[n]   = works.sum()      logp[n] = log_probs.sum()     x_final[n] = positions[-1]
 *   positions : nullable, TIME-MAJOR [n_rec][n], n_rec = S/pos_stride + 1, row i = step i*pos_stride
 *               (row 0 is the equilibrated start, :142)
 *   step_works: nullable, time-major [S+1][n] (row 0 = 0, :141); step_logps: nullable [S][n]
 *   moments   : nullable double[NQ_MOM_COUNT]                                               */
int nq_langevin_forward(const NqLangevinDesc* desc /*host*/, const float* r_table,
                        const float* x0, const uint32_t* seeds, int64_t n,
                        float* work, float* logp, float* x_final,
                        float* positions, int64_t pos_stride,
                        float* step_works, float* step_logps,
                        double* moments, void* workspace, size_t workspace_bytes, void* stream);

/* estimate_gradient_boltzinit / _logP / _reparam / _REINFORCE_recordall
 * (barrier_crossing.py:204-224, 262-333): one fused pass that also accumulates, per
 * trajectory, A_j = sum_k dlogp_k/dr_k * phi[k-1][j] and B_j = sum_k dW/dr_k * phi[k-1][j]
 * (k = 1..S-1; phi[S-1][D] = d r_table[1..S-1] / d variables, row-major) and reduces
 *   grad_sums[0][j] = sum_n W_n A_nj   (log-prob / REINFORCE score term, :270)
 *   grad_sums[1][j] = sum_n B_nj       (reparameterisation term, :271)
 * over the n trajectories of this call (NOT divided by the batch size: the caller divides
 * after summing over shards/ranks).  estimator[n] = logp*W + W (:212).                     */
int nq_langevin_grad(const NqLangevinDesc* desc /*host*/, const float* r_table,
                     const float* phi, int32_t D,
                     const float* x0, const uint32_t* seeds, int64_t n,
                     float* work, float* logp, float* estimator, float* x_final,
                     double* grad_sums /*[2][D]*/, double* moments,
                     void* workspace, size_t workspace_bytes, void* stream);

/* The same fused pass that ALSO records the trajectories -- the `positions` the reference's estimators
 * return next to the gradient (barrier_crossing.py:213, 223) -- time-major [S / pos_stride + 1][n], row 0 the
 * equilibrated start.  Available where the library runs its producer/consumer launch (small ensembles, n <=
 * 9472 by default); NQ_ERR_UNSUPPORTED otherwise: the caller then records with nq_langevin_forward.   */
int nq_langevin_grad_record(const NqLangevinDesc* desc /*host*/, const float* r_table,
                            const float* phi, int32_t D,
                            const float* x0, const uint32_t* seeds, int64_t n,
                            float* work, float* logp, float* estimator, float* x_final,
                            float* positions, int64_t pos_stride,
                            double* grad_sums /*[2][D]*/, double* moments,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Checkpointed forward + replay ("adjoint") pair: the forward pass stores (x, rng) every
 * `ckpt_every` steps; the replay pass regenerates each segment from its checkpoint and
 * accumulates the gradient wrt EVERY trap position,
 *   gbar[k] = sum_n ( lbar[n] * dlogp_{n,k}/dr_k + wbar[n] * dW_n/dr_k ),  k = 0..S
 * for arbitrary per-trajectory cotangents (REINFORCE: lbar = W, wbar = 1).  This is the
 * reverse pass of jax.grad through lax.scan / control_flow.checkpoint_scan
 * (control_flow.py:16-31) for this integrator, where positions carry no derivative
 * (stop_gradient, barrier_crossing.py:88).                                                 */
size_t nq_langevin_ckpt_count(int64_t steps, int64_t ckpt_every);
int nq_langevin_forward_ckpt(const NqLangevinDesc* desc /*host*/, const float* r_table,
                             const float* x0, const uint32_t* seeds, int64_t n,
                             int64_t ckpt_every,
                             float* work, float* logp, float* x_final,
                             float* ckpt_x /*[n_ckpt][n]*/, uint32_t* ckpt_rng /*[n_ckpt][4][n]: chain key + pending sub-key*/,
                             double* moments, void* workspace, size_t workspace_bytes, void* stream);
int nq_langevin_replay(const NqLangevinDesc* desc /*host*/, const float* r_table,
                       int64_t n, int64_t ckpt_every,
                       const float* ckpt_x, const uint32_t* ckpt_rng,
                       const float* lbar /*[n]*/, const float* wbar /*[n]*/,
                       double* gbar /*[S+1]*/, void* workspace, size_t workspace_bytes, void* stream);

/* run_brownian_stationary_trap (barrier_crossing.py:227-241 / simulate.py:101-115):
 * S steps at constant r0; x_final[n] = traj[-1] (single_brownian_eqm :243-247);
 * positions nullable time-major [S][n] (single_brownian_stiffness simulate.py:117-121).   */
int nq_langevin_stationary(const NqLangevinDesc* desc /*host*/, float r0,
                           const float* x0, const uint32_t* seeds, int64_t n,
                           float* x_final, float* positions, void* stream);

/* One apply_fn step (barrier_crossing.py:77-92 / simulate.py:82-97) for n independent states:
 * rng[n][2] is split in place (`key, split = jax.random.split(state.rng)`), x[n] is advanced by
 * one Euler-Maruyama step at trap position r0, log_prob[n] receives Normal.log_prob(dR).     */
int nq_langevin_apply(const NqLangevinDesc* desc /*host*/, float r0, float* x, uint32_t* rng,
                      float* log_prob, int64_t n, void* stream);

/* apply_fn on rank-generic states (simulate.py:66-97): position [n_states][n_particles][dim], ONE key per
 * state; the draw is jax.random.normal(split, [1, n_particles, dim]) (one threefry stream over all elements),
 * log_prob [n_states] the sum of Normal.log_prob over every element.  The reference's energy closures read
 * position[0][0] only (potentials.py:27-52), so element 0 feels landscape + trap and the others diffuse freely. */
int nq_langevin_apply_nd(const NqLangevinDesc* desc, float r0, float* position, uint32_t* rng /*[n_states][2]*/,
                         float* log_prob /*nullable [n_states]*/, int64_t n_states, int32_t n_particles, int32_t dim,
                         void* stream);

/* ------------------------------------------------------------------ Ising */
#define NQ_ISING_NCLASS 10 /* (spin in {-1,+1}) x (neighbour sum in {-4,-2,0,2,4}) */
/* class table slot = 8*(s>0) + (nsum+4)/2 ; slots 5-7 and 13-15 unused */

/* per-(replica, sweep) float outputs, plane-major: out[f][r][t], t = 0..T-2 */
#define NQ_IS_WORK 0
#define NQ_IS_HEAT 1
#define NQ_IS_FWD 2
#define NQ_IS_REV 3
#define NQ_IS_EP 4
#define NQ_IS_MAG 5
#define NQ_IS_ENERGY 6
#define NQ_IS_NSUMMARY 7
/* partial derivatives used by the gradient estimators (Appendix B.2 of SURVEY.md) */
#define NQ_IS_DFWD_DH 0
#define NQ_IS_DFWD_DL 1
#define NQ_IS_DREV_DH 2
#define NQ_IS_DREV_DL 3
#define NQ_IS_NPARTIAL 4

typedef struct NqIsingDesc {
  int32_t L;          /* lattice side of the 2-D square lattice (used when ndim == 0) ising.py:99-106 */
  int32_t T;          /* time_steps; T-1 sweeps                      ising.py:153-173 */
  int64_t R;          /* replicas (rows of split(seed, batch_size),  ising.py:283)   */
  int32_t spins0_broadcast; /* 1: spins0 holds ONE lattice shared by all replicas    */
  int32_t sweep_key_mode;   /* 0: sweep t uses split(seed_r, T-1)[t-1] (simulate_ising :161);
                               1: T must be 2 and the single sweep uses seed_r itself (update :132) */
  int32_t ndim;             /* 0: 2-D square L x L; 1..3: general periodic lattice dims[0..ndim-1], all even
                               (sum_neighbors :75-83 and even_odd_masks :99-106 are rank-generic)        */
  int32_t dims[3];
} NqIsingDesc;

size_t nq_ising_words_per_lattice(int32_t L); /* = L*L/32 (bit b of word w = site 32*w+b, row-major; 1 = spin +1) */
/* Scratch nq_ising_run wants for this shape on the current device.  Required (NQ_ERR_WORKSPACE
 * otherwise) when the lattice exceeds shared memory (L >= 2048, or a large general lattice).
 * Optional for 512 <= L <= 1024 with R above and not a multiple of the SM count: there it holds the
 * replicas between time segments so that the last, partly filled round of CTAs is 1/n as long;
 * without it the run is one CTA per replica and gives bit-identical results. */
size_t nq_ising_workspace_bytes(const NqIsingDesc* desc);

/* simulate_ising (ising.py:153-173) for R replicas: T-1 x update (:129-150), each two
 * masked_update half-sweeps (:109-126, odd-parity sites first :99-106), with
 * seeds_t = split(seed_r, T-1)[t], (seed_even, seed_odd) = split(seed_t, 2) and the
 * full-lattice uniform draw of distrax.Bernoulli reproduced counter-for-counter.
 *   thr     [T-1][16] uint32 : acceptance thresholds ceil(sigmoid(logit_c) * 2^23) per class
 *                              (entries >= NQ_ISING_NCLASS unused); flip iff (bits>>9) < thr
 *   lut     [T-1][4][16] double : per class  logit, log_sigmoid(-logit), log_sigmoid(logit) (the
 *                              float32 values of the reference's arithmetic) and sigmoid(logit) (binary64)
 *   field, log_temp [T] float : the schedule tables (IsingParameters :14-16)
 *   spins0  bit-packed [R or 1][L*L/32];  seeds [R][2]
 *   spins_out bit-packed [R][L*L/32]  (final_state.spins)
 *   summary [NQ_IS_NSUMMARY][R][T-1] float   (IsingSummary :33-40)
 *   partials nullable [NQ_IS_NPARTIAL][R][T-1] float
 *   counts  nullable [R][T-1][2][2][16] int32 : per half-sweep, [0]=active sites, [1]=flipped sites per class
 *   states  nullable bit-packed [R][T-1][L*L/32]  (return_states=True :166-167)           */
int nq_ising_run(const NqIsingDesc* desc /*host*/, const uint32_t* thr, const double* lut,
                 const float* field, const float* log_temp,
                 const uint32_t* spins0, const uint32_t* seeds,
                 uint32_t* spins_out, float* summary, float* partials,
                 int32_t* counts, uint32_t* states,
                 void* workspace, size_t workspace_bytes, void* stream);

/* random_spins(shape, p, seed) (ising.py:71-72) bit-packed, n_sites = prod(shape): bit i = uniform(seed)[i] < p */
int nq_ising_random_spins(const uint32_t key[2] /*host*/, int64_t n_sites, float p, uint32_t* out_bits,
                          void* stream);
/* pack / unpack between +-1 spins (int32 or float32 elements) and the bit-packed layout */
int nq_ising_pack(const void* spins, int32_t is_float, int64_t n_sites, uint32_t* bits, void* stream);
int nq_ising_unpack(const uint32_t* bits, int64_t n_sites, int32_t is_float, void* spins, void* stream);
/* sum_neighbors / energy / magnetisation of bit-packed lattices: out[r] = {M, bond sum B} as int64[2] */
int nq_ising_measure(const uint32_t* bits, int32_t ndim, const int32_t dims[3] /*host*/, int64_t R,
                     int64_t* out_MB, void* stream);

/* Gradient assembly of estimate_gradient_{rev,fwd}* (ising.py:179-236, 294-599): per replica, the loss
 * and the contractions of the per-sweep sensitivities with the schedule Jacobians (what jax.grad /
 * jvp through the schedule pytree produce):
 *   logP_x[r][j] = loss_r * sum_t d fwd_lp / d x_t * J_x[t][j]    (gradient of sum(fwd_lp) * stop_gradient(loss))
 *   rep_x[r][j]  = sum_t d loss / d x_t * J_x[t][j]               (gradient of the loss itself)
 * for x in {log_temp (J_lt [T][Dl]), field (J_h [T][Dh])}, float64, row-major.
 *   kind 0: loss = total_work (:252-255); kind 1: loss = total_entropy_production (:239-249), both in
 *   closed form from `summary`, `partials` (of nq_ising_run) and MB0 (nq_ising_measure of the initial
 *   lattice(s): [R or 1][2]); kind 2: an arbitrary LossFn whose value loss_ext[R] and sensitivities
 *   dL_dh_ext / dL_dl_ext [R][T] the caller obtained elsewhere (the Python host uses torch autograd).
 * log_temp_ends = {log_temp[0], log_temp[T-1]}, field0 = field[0] (host floats).                    */
int nq_ising_grad_assemble(int32_t kind, int32_t T, int64_t R, int64_t n_sites, int32_t broadcast0,
                           const float* log_temp_ends /*host [2]*/, float field0,
                           const float* summary, const float* partials, const int64_t* MB0,
                           const double* J_lt, int32_t Dl, const double* J_h, int32_t Dh,
                           const double* dL_dh_ext, const double* dL_dl_ext, const float* loss_ext,
                           float* loss /*[R]*/, double* logP_lt /*[R][Dl]*/, double* logP_h /*[R][Dh]*/,
                           double* rep_lt, double* rep_h, void* stream);

/* The same assembly (kinds 0 and 1) with the schedule end points read from the DEVICE tables log_temp[T],
 * field[T] the sweep ran on -- no host copy of the schedule (device-resident training step, ising.py:263-290). */
int nq_ising_grad_assemble_dev(int32_t kind, int32_t T, int64_t R, int64_t n_sites, int32_t broadcast0,
                               const float* log_temp /*device [T]*/, const float* field /*device [T]*/,
                               const float* summary, const float* partials, const int64_t* MB0,
                               const double* J_lt, int32_t Dl, const double* J_h, int32_t Dh,
                               float* loss, double* logP_lt, double* logP_h, double* rep_lt, double* rep_h,
                               void* stream);

/* ------------------------------------------------------------------ device-resident training step */
/* The pieces that let one optimisation step of the reference's drivers
 * (run_barrier_crossing_opt.py:113-137; ising.py:263-290) run as a fixed launch sequence with
 * every operand in device memory, so that the host captures it once in a CUDA graph.          */

/* jax.random.split(key, num)[first : first+count] with the key READ FROM DEVICE MEMORY
 * (`key, a, b = split(key, 3)` of run_barrier_crossing_opt.py:125 without a host round trip). */
int nq_random_split_dev(const uint32_t* key_dev /*device [2]*/, int64_t num, int64_t first, int64_t count,
                        uint32_t* out, void* stream);

/* Protocol table from its variables: table[0] = first, table[K+1] = last,
 *   table[1+i] = offset[i] + sum_j variables[j] * basis[i][j]      (offset nullable)
 * i.e. MakeSchedule_chebyshev (barrier_crossing.py:145-151) for basis = the monic Chebyshev basis
 * on the step grid, and any other Parameterization of parameterization.py:67-344 (all are affine in
 * their variables) for basis = its Jacobian, offset = its value at variables = 0.  basis is the
 * same [K][D] array nq_langevin_grad takes as phi.                                              */
int nq_schedule_affine(const float* basis, const float* variables, const float* offset, float first, float last,
                       int64_t K, int32_t D, float* table /*[K+2]*/, void* stream);

/* Schedule table of the Ising drivers' composition (run_ising_opt.py:45-66)
 *   AddBaseline(ConstrainEndpoints(Chebyshev(variables), y0, y1), baseline)   (parameterization.py:144-179, 295-344):
 *   table[i] = ((scale[i] * sum_j variables[j] * basis[i][j]) + offset1[i]) + offset2[i]
 * with one binary32 rounding per operation in that order (scale = the end-point envelope, offset1 = its linear
 * part, offset2 = the baseline; each nullable), all operands in device memory.                                  */
int nq_schedule_compose(const float* basis /*[K][D]*/, const float* variables /*[D]*/, const float* scale /*[K]*/,
                        const float* offset1 /*[K]*/, const float* offset2 /*[K]*/, int64_t K, int32_t D,
                        float* table /*[K]*/, void* stream);

/* Per-sweep class tables nq_ising_run consumes, computed ON THE DEVICE from device schedule tables: the
 * arithmetic of flip_logits (ising.py:93-96) and distrax.Bernoulli's sigmoid / log_sigmoid in float32 with
 * binary64-evaluated transcendentals rounded once (bit-equal to the host tables of noneq_opt_b200.ising).
 * thr [T-1][16] uint32, lut [T-1][4][16] float64; ndim = lattice rank.                                         */
int nq_ising_class_tables(const float* log_temp /*device [T]*/, const float* field /*device [T]*/, int32_t T,
                          int32_t ndim, uint32_t* thr, double* lut, void* stream);

/* Batch mean of the per-replica schedule gradients nq_ising_grad_assemble leaves on the device
 * (tree_map(mean) of ising.py:286): mean_out[Dl + Dh] = mean_r float32(w_logp * logP[r] + w_rep * rep[r]),
 * log-temperature variables first.  Feed mean_out to nq_adam_step(grad_f32 = ...).                           */
int nq_ising_grad_mean(const double* logP_lt, const double* logP_h, const double* rep_lt, const double* rep_h,
                       int64_t R, int32_t Dl, int32_t Dh, double w_logp, double w_rep, float* mean_out /*[Dl+Dh]*/,
                       void* stream);

typedef struct NqAdamDesc {
  double step_size, decay_steps, decay_rate; /* lr(i) = step_size * decay_rate^(i/decay_steps);
                                                decay_steps <= 0: constant  (optimizers.exponential_decay) */
  double b1, b2, eps;                        /* optimizers.adam defaults 0.9, 0.999, 1e-8 */
  double max_norm;                           /* optimizers.clip_grads (ising.py:287); <= 0: no clipping */
  double w0, w1, denom;                      /* g = (w0*grad_sums0 + w1*grad_sums1) / denom  (barrier_crossing.py:222) */
} NqAdamDesc;

/* clip_grads + jax.example_libraries.optimizers.adam update of x[P] (binary32, one rounding per
 * operation, as the reference computes it), with the gradient taken either from the float64 sums
 * nq_langevin_grad leaves on the device (grad_sums1 nullable) or from grad_f32[P]; increments the
 * device step counter; grad_out (nullable) receives the clipped gradient.                       */
int nq_adam_step(const NqAdamDesc* desc /*host*/, int32_t P, const double* grad_sums0, const double* grad_sums1,
                 const float* grad_f32, float* x, float* m, float* v, int64_t* step_dev, float* grad_out,
                 void* stream);

/* ------------------------------------------------------------------ multi-GPU combine */
/* The one collective of the path: in-place sum of `n` f64 values (gradient sums and
 * moments) over the ranks of `nccl_comm`, an ncclComm_t the CALLER created.  The reference has no
 * multi-device code; each rank runs its contiguous slice of split(seed, B) and calls this once.
 * ncclAllReduce is resolved at run time from the libnccl the process already holds (else the
 * system one), so the library itself carries no NCCL link dependency.  nq_langevin_grad /
 * nq_ising_run leave their sums on the device; enqueue this on the same stream.               */
int nq_allreduce_f64(void* nccl_comm, double* buf /*device [n]*/, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NONEQ_B200_H_ */
