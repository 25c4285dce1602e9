/* noneq_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY), plain C restatement of the reference's
 * hot path, used (a) to cross-check the numpy oracle at sizes numpy is too slow for and (b) as the
 * timed CPU arm of bench.py (`cpu_baseline`, `--impl reference`, kind "port"): OpenMP over
 * trajectories / replicas on all host cores.  The real reference (JAX on CPU) cannot be installed
 * in this image; see oracle/__init__.py for the pinning status.
 *
 * Follows, operation by operation in float32 (transcendentals in binary64 rounded once):
 *   noneq_opt/barrier_crossing.py:28-143, 176-194, 204-224   (Langevin: brownian, run_brownian_opt,
 *                                                             potentials, REINFORCE estimator)
 *   noneq_opt/ising.py:75-173                                 (Ising: sum_neighbors .. simulate_ising)
 * with the jax / tfp / distrax semantics of oracle/jaxlike.py (threefry2x32-20 original layout,
 * uniform, XLA ErfInv32, Normal.log_prob, Bernoulli sample/log_prob).
 *
 * Like the reference, the Ising update draws the full L*L uniforms of every half-sweep and
 * evaluates logits / log-probs on the whole lattice before masking (ising.py:114-117).
 *
 * Build: oracle/c/Makefile  ->  oracle/c/_build/libnoneq_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ threefry2x32-20 */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* y0, uint32_t* y1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *y0 = x0;
  *y1 = x1;
}

/* element i of threefry_2x32(key, iota(size)) */
static uint32_t random_bits_at(uint32_t k0, uint32_t k1, uint64_t i, uint64_t size) {
  uint64_t half = (size + 1) / 2;
  uint32_t y0, y1;
  if (i < half) {
    uint64_t j = i + half;
    threefry2x32(k0, k1, (uint32_t)i, j < size ? (uint32_t)j : 0u, &y0, &y1);
    return y0;
  }
  threefry2x32(k0, k1, (uint32_t)(i - half), (uint32_t)i, &y0, &y1);
  return y1;
}

static void split2(uint32_t k0, uint32_t k1, uint32_t* n0, uint32_t* n1, uint32_t* s0, uint32_t* s1) {
  threefry2x32(k0, k1, 0u, 2u, n0, s0);
  threefry2x32(k0, k1, 1u, 3u, n1, s1);
}

static inline float bits_to_unit(uint32_t bits) {
  union { uint32_t u; float f; } v;
  v.u = (bits >> 9) | 0x3F800000u;
  return v.f - 1.0f;
}

static inline float exp32(float x) { return (float)exp((double)x); }
static inline float log32(float x) { return (float)log((double)x); }
static inline float log1p32(float x) { return (float)log1p((double)x); }

static float erf_inv32(float x) {
  static const float A[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                             -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float B[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                             -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  volatile float xx = x * x;
  float w = -log1p32(-xx);
  const float* c = w < 5.0f ? A : B;
  volatile float w2 = w < 5.0f ? w - 2.5f : sqrtf(w) - 3.0f;
  volatile float p = c[0];
  for (int i = 1; i < 9; ++i) {
    volatile float t = p * w2;
    p = c[i] + t;
  }
  if (fabsf(x) == 1.0f) return x * INFINITY;
  volatile float r = p * x;
  return r;
}

static float normal_from_bits(uint32_t bits) {
  const float lo = -0.99999994f;
  volatile float f = bits_to_unit(bits);
  volatile float t = f * 2.0f;
  volatile float u = t + lo;
  float uu = u > lo ? u : lo;
  volatile float r = 1.41421356237309504880f * erf_inv32(uu);
  return r;
}

/* ------------------------------------------------------------------ Langevin */
typedef struct {
  int potential, periodic;
  float hk, k_s, inv_beta, cl, cr, bde, xm, two_xm;
  float nu, dt, sigma, log_norm, box;
} LConsts;

static void make_consts(LConsts* c, int potential, int periodic, double k_s, double kappa_l, double kappa_r,
                        double x_m, double delta_E, double beta, double dt, double kT, double gamma, double mass,
                        double box) {
  c->potential = potential;
  c->periodic = periodic;
  c->hk = (float)(k_s / 2);
  c->k_s = (float)k_s;
  c->inv_beta = potential == 0 ? 0.f : (float)(1. / beta);
  c->cl = (float)(-0.5 * beta * kappa_l);
  c->cr = (float)(0.5 * beta * kappa_r);
  c->bde = (float)(beta * delta_E);
  c->xm = (float)x_m;
  c->two_xm = (float)(2. * x_m);
  volatile float mg = (float)mass * (float)gamma;
  volatile float nu = 1.0f / mg;
  volatile float v1 = 2.0f * (float)kT;
  volatile float v2 = v1 * (float)dt;
  volatile float var = v2 * nu;
  c->nu = nu;
  c->dt = (float)dt;
  c->sigma = sqrtf(var);
  volatile float ls = log32(c->sigma);
  c->log_norm = (float)0.91893853320467274178 + ls;
  c->box = (float)box;
}

/* every binary op through a volatile float: one binary32 rounding per operation, no contraction */
#define F(expr) ({ volatile float v__ = (expr); v__; })

typedef struct { float ul, ur, A, B, sAB, Em; } Land;

static void land_eval(const LConsts* c, float x, Land* l) {
  if (c->potential == 0) { l->Em = 0.f; return; }
  if (c->potential == 1) { l->ul = F(x + c->xm); l->ur = F(x - c->xm); }
  else { l->ul = x; l->ur = F(x - c->two_xm); }
  float a = F(c->cl * F(l->ul * l->ul));
  float v = F(F(c->cr * F(l->ur * l->ur)) + c->bde);
  l->A = exp32(a);
  l->B = exp32(-v);
  l->sAB = F(l->A + l->B);
  l->Em = F(-c->inv_beta * log32(l->sAB));
}

static float spring_energy(const LConsts* c, float x, float r) {
  float u = F(x - r);
  return F(c->hk * F(u * u));
}

static float force_at(const LConsts* c, const Land* l, float x, float r) {
  float u = F(x - r);
  float dEs = F(c->hk * F(2.0f * u));
  if (c->potential == 0) return -dEs;
  float g = F(-c->inv_beta / l->sAB);
  float ga = F(g * l->A), gb = F(g * l->B);
  float dxa = F(F(ga * c->cl) * F(2.0f * l->ul));
  float dxb = F(F(-gb * c->cr) * F(2.0f * l->ur));
  return -F(F(dxa + dxb) + dEs);
}

static float py_modf(float x, float box) {
  float r = fmodf(x, box);
  if (r != 0.f && ((r < 0.f) != (box < 0.f))) r += box;
  return r;
}

/* one apply_fn: returns dR, updates x / key, outputs z and lp */
static float brownian_apply(const LConsts* c, const Land* l, float* x, uint32_t* k0, uint32_t* k1, float r,
                            float* z, float* lp, float* mean_out) {
  float Fr = force_at(c, l, *x, r);
  float mean = F(F(Fr * c->nu) * c->dt);
  uint32_t n0, n1, s0, s1, bits, unused;
  split2(*k0, *k1, &n0, &n1, &s0, &s1);
  *k0 = n0; *k1 = n1;
  threefry2x32(s0, s1, 0u, 0u, &bits, &unused);
  float xi = normal_from_bits(bits);
  float dR = F(F(xi * c->sigma) + mean);
  *z = F(F(dR / c->sigma) - F(mean / c->sigma));
  *lp = F(F(-0.5f * F(*z * *z)) - c->log_norm);
  *x = F(*x + dR);
  if (c->periodic) *x = py_modf(*x, c->box);
  *mean_out = mean;
  return dR;
}

/* estimate_gradient_boltzinit for trajectories [0, n): per-trajectory work / logp (+ optional positions
 * [n][S+1]) and the un-normalised gradient sums grad[2][D] (logP term, reparam term) in binary64. */
void oracle_langevin(int potential, int periodic, double k_s, double kappa_l, double kappa_r, double x_m,
                     double delta_E, double beta, double dt, double kT, double gamma, double mass, double box,
                     long S, long Neq, const float* r, const float* phi, int D, const float* x0,
                     const uint32_t* seeds, long n, float* work, float* logp, double* grad, float* positions) {
  LConsts c;
  make_consts(&c, potential, periodic, k_s, kappa_l, kappa_r, x_m, delta_E, beta, dt, kT, gamma, mass, box);
  const int Dn = D > 0 ? D : 1;
  if (grad) memset(grad, 0, sizeof(double) * 2 * Dn);
#pragma omp parallel
  {
    double* gl = (double*)calloc(2 * Dn, sizeof(double));
    double* ga = (double*)calloc(2 * Dn, sizeof(double));
#pragma omp for schedule(dynamic, 16)
    for (long t = 0; t < n; ++t) {
      float x = x0[t];
      uint32_t k0, k1, d0, d1;
      split2(seeds[2 * t], seeds[2 * t + 1], &d0, &d1, &k0, &k1);
      Land l;
      float z, lp, mean;
      for (long e = 0; e < Neq; ++e) {
        land_eval(&c, x, &l);
        brownian_apply(&c, &l, &x, &k0, &k1, r[0], &z, &lp, &mean);
      }
      if (positions) positions[t * (S + 1)] = x;
      double W = 0.0, LP = 0.0;
      for (int j = 0; j < 2 * Dn; ++j) ga[j] = 0.0;
      for (long k = 1; k <= S; ++k) {
        land_eval(&c, x, &l);
        float dW;
        if (c.potential == 0) dW = F(spring_energy(&c, x, r[k]) - spring_energy(&c, x, r[k - 1]));
        else dW = F(F(l.Em + spring_energy(&c, x, r[k])) - F(l.Em + spring_energy(&c, x, r[k - 1])));
        float dR = brownian_apply(&c, &l, &x, &k0, &k1, r[k], &z, &lp, &mean);
        W += (double)dW;       /* exact sum of the f32 terms, rounded once below */
        LP += (double)lp;
        if (positions) positions[t * (S + 1) + k] = x;
        if (D > 0 && k <= S - 1) {
          float a = F(F(F(F(z / c.sigma) * c.nu) * c.dt) * c.k_s);
          float b = F(c.k_s * dR);
          const float* row = phi + (k - 1) * D;
          for (int j = 0; j < D; ++j) {
            ga[j] += (double)a * (double)row[j];
            ga[Dn + j] += (double)b * (double)row[j];
          }
        }
      }
      work[t] = (float)W;
      logp[t] = (float)LP;
      for (int j = 0; j < D; ++j) {
        gl[j] += (double)work[t] * ga[j];
        gl[Dn + j] += ga[Dn + j];
      }
    }
    if (grad) {
#pragma omp critical
      for (int j = 0; j < 2 * Dn; ++j) grad[j] += gl[j];
    }
    free(gl);
    free(ga);
  }
}

/* ------------------------------------------------------------------ Ising */
static inline int wrap(int i, int L) { return i < 0 ? i + L : (i >= L ? i - L : i); }

static float ising_energy(const int8_t* s, int L, float field) {
  double acc = 0.0;
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int nb = s[wrap(i - 1, L) * L + j] + s[wrap(i + 1, L) * L + j] + s[i * L + wrap(j - 1, L)] + s[i * L + wrap(j + 1, L)];
      float t = F((float)(-s[i * L + j]) * F((float)nb / 2.0f + field));
      acc += (double)t;
    }
  return (float)acc;
}

static inline float log_sigmoid32(float x) {
  float a = -x;
  float m = a > 0.f ? a : 0.f;
  return -F(m + log1p32(exp32(-fabsf(a))));
}

/* masked_update (ising.py:109-126): full-lattice logits / uniforms / log-probs, then mask */
static void masked_update(int8_t* s, int L, float log_temp, float field, int parity, uint32_t k0, uint32_t k1,
                          float* fwd, float* rev, float* logits, int8_t* flips) {
  const int N = L * L;
  const float einv = exp32(-log_temp);
  double f = 0.0, r = 0.0;
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int p = i * L + j;
      int nb = s[wrap(i - 1, L) * L + j] + s[wrap(i + 1, L) * L + j] + s[i * L + wrap(j - 1, L)] + s[i * L + wrap(j + 1, L)];
      float lg = F(F((float)(-2 * s[p]) * F((float)nb + field)) * einv);
      logits[p] = lg;
      float u = bits_to_unit(random_bits_at(k0, k1, (uint64_t)p, (uint64_t)N));
      float prob = F(1.0f / F(1.0f + exp32(-lg)));
      int mask = ((i + j) & 1) == parity;
      flips[p] = (int8_t)((u < prob) && mask);
      if (mask) f += (double)(flips[p] ? log_sigmoid32(lg) : log_sigmoid32(-lg));
    }
  for (int p = 0; p < N; ++p)
    if (flips[p]) s[p] = (int8_t)(-s[p]);
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int p = i * L + j;
      if ((((i + j) & 1) == parity)) {
        int nb = s[wrap(i - 1, L) * L + j] + s[wrap(i + 1, L) * L + j] + s[i * L + wrap(j - 1, L)] + s[i * L + wrap(j + 1, L)];
        float lg = F(F((float)(-2 * s[p]) * F((float)nb + field)) * einv);
        r += (double)(flips[p] ? log_sigmoid32(lg) : log_sigmoid32(-lg));
      }
    }
  *fwd = (float)f;
  *rev = (float)r;
}

/* simulate_ising for R replicas; summary [R][7][T-1]; spins int8 +-1 */
// snaps (nullable): [R][n_snap][L*L] lattices after sweeps snap_t[0..n_snap) (1-based sweep index t, ascending)
void oracle_ising(int L, int T, long R, const float* log_temp, const float* field, const int8_t* spins0,
                  int broadcast0, const uint32_t* seeds, int8_t* final_spins, float* summary, int n_snap,
                  const int* snap_t, int8_t* snaps) {
  const int N = L * L;
#pragma omp parallel
  {
    int8_t* s = (int8_t*)malloc(N);
    int8_t* flips = (int8_t*)malloc(N);
    float* logits = (float*)malloc(sizeof(float) * N);
#pragma omp for schedule(dynamic, 1)
    for (long rep = 0; rep < R; ++rep) {
      memcpy(s, spins0 + (broadcast0 ? 0 : rep * N), N);
      const uint32_t r0 = seeds[2 * rep], r1 = seeds[2 * rep + 1];
      float* out = summary + rep * 7 * (T - 1);
      for (int t = 1; t < T; ++t) {
        uint32_t s0 = random_bits_at(r0, r1, 2ull * (t - 1), 2ull * (T - 1));
        uint32_t s1 = random_bits_at(r0, r1, 2ull * (t - 1) + 1, 2ull * (T - 1));
        uint32_t e0, e1, o0, o1;
        split2(s0, s1, &e0, &e1, &o0, &o1);
        float e_init = ising_energy(s, L, field[t - 1]);
        float e_mid = ising_energy(s, L, field[t]);
        float ef, er, of, orv;
        masked_update(s, L, log_temp[t], field[t], 1, e0, e1, &ef, &er, logits, flips);
        masked_update(s, L, log_temp[t], field[t], 0, o0, o1, &of, &orv, logits, flips);
        float e_fin = ising_energy(s, L, field[t]);
        double m = 0.0;
        for (int p = 0; p < N; ++p) m += s[p];
        float fwd = F(ef + of), rev = F(er + orv);
        out[0 * (T - 1) + t - 1] = F(e_mid - e_init);
        out[1 * (T - 1) + t - 1] = F(e_mid - e_fin);
        out[2 * (T - 1) + t - 1] = fwd;
        out[3 * (T - 1) + t - 1] = rev;
        out[4 * (T - 1) + t - 1] = F(fwd - rev);
        out[5 * (T - 1) + t - 1] = (float)(m / N);
        out[6 * (T - 1) + t - 1] = e_fin;
        if (snaps)
          for (int q = 0; q < n_snap; ++q)
            if (snap_t[q] == t) memcpy(snaps + ((size_t)rep * n_snap + q) * N, s, N);
      }
      if (final_spins) memcpy(final_spins + rep * N, s, N);
    }
    free(s);
    free(flips);
    free(logits);
  }
}

// n <= 0: every core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers; the timed
// CPU arms of bench.py call this so that "all host cores" means the same thing at every --gpus N.
int oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n <= 0) n = omp_get_num_procs();
  omp_set_dynamic(0);
  omp_set_num_threads(n);
  return n;
#else
  (void)n;
  return 1;
#endif
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
