// Checkerboard Glauber sweeps of the 2-D periodic Ising model under a time-varying field and
// temperature, with trajectory log-probabilities and their parameter derivatives (sm_100a).
//
// Replaces the XLA computation traced from the reference's
//   masked_update / update / simulate_ising   noneq_opt/ising.py:109-173
//   (sum_neighbors :75-83, energy :85-86, flip_logits :93-96, even_odd_masks :99-106)
// and provides the per-sweep sufficient statistics from which IsingSummary (:33-40) and the
// gradient estimators (:179-236, 294-599) are assembled.
//
// Design
//   * the lattice is bit-packed (1 = spin +1) and lives in SHARED MEMORY for the whole
//     trajectory, stored checkerboard-compressed: plane c holds the sites updated in
//     half-sweep c (c = 0: (i+j) odd -- the reference's "even" mask; c = 1: (i+j) even), so
//     a 32-bit word is 32 simultaneously-updatable sites and all four neighbours of a word
//     are whole words of the other plane (one of them funnel-shifted by one site);
//   * uniforms are jax's: site (i,j) of a half-sweep consumes counter i*L+j of
//     threefry_2x32(half_key, iota(L*L)), whose counter pairing (p, p+L*L/2) makes ONE
//     threefry block serve site (i,j) and site (i+L/2,j) -- same colour when L % 4 == 0 --
//     so the half of the draw the reference discards is never generated;
//   * acceptance is an integer compare (bits>>9) < thr[class] against a per-sweep table of
//     ceil(sigmoid(logit)*2^23) for the 2x5 (spin, #up-neighbours) classes, so spin histories
//     do not depend on device transcendentals;
//   * neighbour counts are bit-sliced (full adders on 32 sites at a time); per-class counts of
//     active and flipped sites come from popc and are reduced with redux.sync; every float
//     summary is a fixed linear form of those integer counts, evaluated in binary64.
#include <math.h>

#include "nq_common.cuh"

namespace nq {

#ifndef NQ_ISING_TEAMS     // warp teams (replicas) per CTA of the warp-team kernel
#define NQ_ISING_TEAMS 4
#endif
#ifndef NQ_ISING_MINB
#define NQ_ISING_MINB (NQ_ISING_TEAMS == 4 ? 4 : 1)
#endif
#ifndef NQ_ISING_BLOCKS_PER_ITER
#define NQ_ISING_BLOCKS_PER_ITER 8
#endif
constexpr int kIsingClasses = 16;  // class = 8*(spin up) + (#up neighbours 0..4)

// ------------------------------------------------------------------------------ bit helpers
__host__ __device__ __forceinline__ uint32_t compact_even_bits(uint32_t x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0f0f0f0fu;
  x = (x | (x >> 4)) & 0x00ff00ffu;
  x = (x | (x >> 8)) & 0x0000ffffu;
  return x;
}
__host__ __device__ __forceinline__ uint32_t spread_to_even_bits(uint32_t x) {
  x &= 0x0000ffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
// 8 bits -> 8 nibbles (bit k -> bit 4k)
__host__ __device__ __forceinline__ uint32_t spread8_to_nibbles(uint32_t v) {
  v &= 0xffu;
  v = (v | (v << 12)) & 0x000f000fu;
  v = (v | (v << 6)) & 0x03030303u;
  v = (v | (v << 3)) & 0x11111111u;
  return v;
}

// row-major external word pair (sites 64w..64w+63 of row i) <-> the two colour planes
__device__ __forceinline__ void split_colours(int row, uint32_t lo, uint32_t hi, uint32_t& c0, uint32_t& c1) {
  const uint32_t even = compact_even_bits(lo) | (compact_even_bits(hi) << 16);
  const uint32_t odd = compact_even_bits(lo >> 1) | (compact_even_bits(hi >> 1) << 16);
  // colour 0 = (i+j) odd: odd j on even rows, even j on odd rows
  c0 = (row & 1) ? even : odd;
  c1 = (row & 1) ? odd : even;
}
__device__ __forceinline__ void merge_colours(int row, uint32_t c0, uint32_t c1, uint32_t& lo, uint32_t& hi) {
  const uint32_t even = (row & 1) ? c0 : c1, odd = (row & 1) ? c1 : c0;
  lo = spread_to_even_bits(even) | (spread_to_even_bits(odd) << 1);
  hi = spread_to_even_bits(even >> 16) | (spread_to_even_bits(odd >> 16) << 1);
}

struct Planes {
  uint32_t S, C0, C1, C2;
};

// number of up neighbours of 32 sites at once: 4 words -> 3 bit planes
__device__ __forceinline__ void count_planes(uint32_t a, uint32_t b, uint32_t c, uint32_t d, Planes& p) {
  const uint32_t s1 = a ^ b, k1 = a & b, s2 = c ^ d, k2 = c & d;
  p.C0 = s1 ^ s2;
  const uint32_t t = s1 & s2;
  p.C1 = k1 ^ k2 ^ t;
  p.C2 = (k1 & k2) | (t & (k1 ^ k2));
}

struct IsingShape {
  int L, NW, T;     // NW = L/64 words per row per colour (2-D fast path)
  long long R;
  int broadcast0;
  int direct_keys;
  int ndim, d0, d1, d2;  // general lattice d0 x d1 x d2 (unused trailing dims = 1), row-major
  int N;                 // sites
  signed char* scratch;  // generic path: global spin scratch [R][N] when N does not fit shared memory
  uint32_t* lat_global;  // fast path, lattices beyond shared memory: colour planes [R][2][L][NW] in HBM/L2
  uint32_t one;          // = 1 at run time (see threefry2x32_allmad)
  // segment scheduling (CTA-team kernel, R not a multiple of the resident CTAs): the T-1 sweeps of a
  // replica are cut into n_seg segments handed out through a ticket counter, so the last wave is
  // 1/n_seg as long.  Between segments a replica lives in seg_lat / seg_mb.
  int n_seg;
  unsigned* seg_ctl;     // [0] ticket counter, [1 + r] segments of replica r completed, [1 + R] status (zeroed per launch)
  long long* seg_mb;     // [R][2] spin sum, bond sum carried between segments
  uint32_t* seg_lat;     // [R][2][L][NW] colour planes carried between segments
  unsigned seg_max_polls;  // bound of a wait for the previous segment, in polls of 256 ns; 0 = unbounded
  unsigned seg_units;      // tickets per segment index: R (CTA team) or R / teams-per-CTA (warp teams)
};

struct IsingIO {
  const uint32_t* thr;    // [T-1][16]
  const double* lut;      // [T-1][4][16]: logit, log_sigmoid(-logit), log_sigmoid(logit), sigmoid(logit)
  const float* field;     // [T]
  const float* log_temp;  // [T]
  const uint32_t* spins0;
  const uint32_t* seeds;  // [R][2]
  uint32_t* spins_out;
  float* summary;         // [7][R][T-1]
  float* partials;        // [4][R][T-1] or null
  int32_t* counts;        // [R][T-1][2][2][16] or null
  uint32_t* states;       // [R][T-1][L*L/32] or null
};

// neighbour planes of word w of row r of colour c (reads the OTHER colour plane)
__device__ __forceinline__ void load_planes(const uint32_t* lat, const IsingShape& sh, int c, int r, int w, Planes& p) {
  const int L = sh.L, NW = sh.NW;
  const uint32_t* own = lat + (size_t)c * L * NW;
  const uint32_t* oth = lat + (size_t)(1 - c) * L * NW;
  const int ru = r == 0 ? L - 1 : r - 1, rd = r == L - 1 ? 0 : r + 1;
  const uint32_t up = oth[ru * NW + w], dn = oth[rd * NW + w], X = oth[r * NW + w];
  uint32_t Xs;
  if (((r & 1) == 0) == (c == 0)) {  // second in-row neighbour is site k+1
    const uint32_t Xn = oth[r * NW + (w + 1 == NW ? 0 : w + 1)];
    Xs = __funnelshift_r(X, Xn, 1);
  } else {  // site k-1
    const uint32_t Xp = oth[r * NW + (w == 0 ? NW - 1 : w - 1)];
    Xs = __funnelshift_l(Xp, X, 1);
  }
  p.S = own[r * NW + w];
  count_planes(up, dn, X, Xs, p);
}

// 8 sites -> nibble word of class indices (8*s + cnt), via the pre-shifted spread tables
__device__ __forceinline__ uint32_t class_nibbles(const uint32_t* spread, const Planes& p, int q) {
  const uint32_t sel = 0x4440u | (uint32_t)q;  // byte q, zero-extended
  return spread[3 * 256 + __byte_perm(p.S, 0, sel)] | spread[2 * 256 + __byte_perm(p.C2, 0, sel)] |
         spread[1 * 256 + __byte_perm(p.C1, 0, sel)] | spread[__byte_perm(p.C0, 0, sel)];
}

// thresholds of the two sites whose class nibbles form byte `byte` of N: one 64-bit lookup in the
// per-sweep pair table thr2[256] = {thr[b & 15], thr[b >> 4]}
__device__ __forceinline__ uint2 thr_pair(const uint2* thr2, uint32_t N, int byte) {
  return thr2[__byte_perm(N, 0, 0x4440 | byte)];
}

__device__ __forceinline__ void build_pair_table(uint2* dst, const uint32_t* thr16, int tid, int tsize) {
  for (int b = tid; b < 256; b += tsize) dst[b] = make_uint2(thr16[b & 15], thr16[b >> 4]);
}

// per-class popcounts of active and flipped sites of one word
__device__ __forceinline__ void histogram(const Planes& p, uint32_t F, int (&hA)[10], int (&hF)[10]) {
  const uint32_t E[5] = {~(p.C2 | p.C1 | p.C0), ~p.C2 & ~p.C1 & p.C0, p.C1 & ~p.C0, p.C1 & p.C0, p.C2};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t up = E[k] & p.S, down = E[k] & ~p.S;
    hA[k] += __popc(down);
    hA[5 + k] += __popc(up);
    hF[k] += __popc(down & F);
    hF[5 + k] += __popc(up & F);
  }
}

// One unit = word w of rows i and i+L/2 of colour c: 32 threefry blocks, 64 site updates.
// ALLMAD: every threefry addition as IMAD (FMA pipe).  Measured on B200: +5 % for the 1024-thread
// CTA-per-replica kernel (ALU pipe 81 % busy, FMA pipe 20 %), -5 % for the warp-per-replica kernel
// (which is closer to its issue limit), so it is chosen per kernel.
// MODE 0: plain threefry; 1 (ALLMAD): every addition as IMAD; 2 (INJ): only the key injections as IMADs, with the
// injection words ks_b + i formed once per half-sweep (threefry2x32_inj) -- the warp-team kernel's IADD3s
// (3.5 per spin-update on the binding ALU pipe, profiles/r02e_ising_c3_noseg_opcodes.md) move to the FMA pipe.
#ifndef NQ_ISING_WARP_MODE
#define NQ_ISING_WARP_MODE 2
#endif
#ifndef NQ_ISING_CTA_MODE
#define NQ_ISING_CTA_MODE 1
#endif
// 1: the 23 uniform bits jax keeps of a word (bits >> 9) as the high half of bits * 2^23 (IMAD.HI, FMA pipe, fused
// with the subtraction of the threshold) instead of a shift on the binding ALU pipe; the 2^23 comes from the
// run-time 1 so that ptxas cannot turn the multiply back into a shift
// 1: blocks enter the hash after the first round's addition (threefry2x32_pre): the counter words of a unit advance
// by constants, so x1 and x0 + x1 are per-unit bases plus constants -- one addition fewer per block
#ifndef NQ_ISING_PRESUM
#define NQ_ISING_PRESUM 1
#endif
constexpr bool kPreSumOn = NQ_ISING_PRESUM != 0;   // warp teams only (measured, r02z: 64^2 5.31e11 -> 5.52e11, 128^2 4.96e11 ->
// 5.17e11; the CTA team, whose adds are all IMADs, loses 3 %: 1024^2 5.97e11 -> 5.80e11 -- ptxas hoists the products and
// adds the constants on the ALU pipe)
#ifndef NQ_ISING_MULHI
#define NQ_ISING_MULHI 0
#endif
template <int MODE>
__device__ __forceinline__ void update_unit(uint32_t* lat, const IsingShape& sh, const uint32_t* spread,
                                            const uint2* thr2, const KeySchedule& ks, int c, int i, int w,
                                            int (&hA)[10], int (&hF)[10]) {
  constexpr bool ALLMAD = MODE == 1;
  constexpr bool kPreSum = kPreSumOn && !ALLMAD;
  KeyInject kinj;
  if (MODE == 2 || kPreSum) kinj = key_inject(ks.ks0, ks.ks1, sh.one);
  const int L = sh.L, NW = sh.NW, i2 = i + L / 2;
  Planes pa, pb;
  load_planes(lat, sh, c, i, w, pa);
  load_planes(lat, sh, c, i2, w, pb);
  const int off = c == 0 ? 1 - (i & 1) : (i & 1);
  const uint32_t cbase = (uint32_t)i * L + 2u * (32u * w) + off;
  const uint32_t half = (uint32_t)L * L / 2;
  const uint32_t b1 = cbase + half + ks.ks1, b01 = 2u * cbase + half + ks.ks0 + ks.ks1;   // kPreSum bases
  // Sites are visited from 31 down to 0 so that the acceptance bits can be shifted in from the
  // right: accept <=> (bits >> 9) - thr < 0, and one funnel shift moves that sign bit into the mask
  // (the subtraction can issue on the FMA pipe; the ALU pipe, which bounds this kernel, sees one
  // shift instead of compare + select + or).
  uint32_t Fa = 0, Fb = 0, Na = 0, Nb = 0;
  const uint32_t two23 = sh.one << 23, sixteen = sh.one << 4;
#if NQ_ISING_BLOCKS_PER_ITER == 8
#pragma unroll 1
  for (int q = 3; q >= 0; --q) {
    Na = class_nibbles(spread, pa, q);
    Nb = class_nibbles(spread, pb, q);
#pragma unroll
    for (int j = 3; j >= 0; --j) {
#else
  // 4 threefry blocks (8 site updates) per loop iteration: a ~5 KB body that stays in the
  // instruction cache when several warps of a sub-partition are in different phases
#pragma unroll 1
  for (int qq = 7; qq >= 0; --qq) {
    const int q = qq >> 1;
    if (qq & 1) {
      Na = class_nibbles(spread, pa, q);
      Nb = class_nibbles(spread, pb, q);
    }
#pragma unroll
    for (int jj = 1; jj >= 0; --jj) {
      const int j = 2 * (qq & 1) + jj;
#endif
      const uint2 ta = thr_pair(thr2, Na, j), tb = thr_pair(thr2, Nb, j);
#pragma unroll
      for (int e = 1; e >= 0; --e) {
        const uint32_t ctr = cbase + 2u * (8u * q + 2u * j + e);
        uint32_t y0, y1;
        if (kPreSum) {
          // x1 = ctr + half + ks1 and x0 + x1 = 2 ctr + half + ks0 + ks1: both are (a per-unit base) + (16 q resp. 32 q)
          // + a compile-time constant, so a block enters the hash after its first addition with two adds, not three
          const uint32_t o = 2u * (2u * j + e);
          const uint32_t bq1 = ALLMAD ? mad1((uint32_t)q, sixteen, b1) : b1 + 16u * q;      // once per q
          const uint32_t bq01 = ALLMAD ? mad1((uint32_t)q, 2u * sixteen, b01) : b01 + 32u * q;
          const uint32_t x1 = ALLMAD ? mad1(bq1, sh.one, o) : bq1 + o;
          const uint32_t s01 = ALLMAD ? mad1(bq01, sh.one, 2u * o) : bq01 + 2u * o;
          threefry2x32_pre<MODE>(kinj, s01, x1, sh.one, y0, y1);
        } else if (ALLMAD) threefry2x32_allmad(ks, ctr, ctr + half, y0, y1, sh.one);
        else if (MODE == 2) threefry2x32_inj<false>(kinj, ctr + ks.ks0, ctr + half + ks.ks1, sh.one, y0, y1);
        else threefry2x32(ks, ctr, ctr + half, y0, y1);
        const uint32_t da = (NQ_ISING_MULHI ? __umulhi(y0, two23) : (y0 >> 9)) - (e ? ta.y : ta.x);
        const uint32_t db = (NQ_ISING_MULHI ? __umulhi(y1, two23) : (y1 >> 9)) - (e ? tb.y : tb.x);
        Fa = __funnelshift_l(da, Fa, 1);
        Fb = __funnelshift_l(db, Fb, 1);
      }
    }
  }
  uint32_t* own = lat + (size_t)c * L * NW;
  own[i * NW + w] = pa.S ^ Fa;
  own[i2 * NW + w] = pb.S ^ Fb;
  histogram(pa, Fa, hA, hF);
  histogram(pb, Fb, hA, hF);
}

// ------------------------------------------------------------------------------ summaries
// A site has 2*ndim neighbours, so ncls = 2*ndim+1 up-neighbour counts per spin value.  Class
// c in [0, 2*ncls): spin = c >= ncls ? +1 : -1, #up-neighbours = c % ncls; table slot = 8*(spin up) + n_up
__device__ __forceinline__ int slot_of(int c, int ncls) { return c >= ncls ? 8 + (c - ncls) : c; }

// Integer sufficient statistics of one sweep -> the float outputs.  Executed by one full warp;
// lane c < 10 owns class c and holds A[h], F[h] (active / flipped counts of its class in
// half-sweep h).  M, B: spin sum and bond sum BEFORE the sweep; updated to AFTER.
__device__ __forceinline__ void emit_summary(const int (&A)[2], const int (&F)[2], const IsingIO& io,
                                             const IsingShape& sh, long long rep, int t, long long& M,
                                             long long& B, int lane) {
  const int ncls = 2 * sh.ndim + 1;
  const bool on = lane < 2 * ncls;
  const int c10 = on ? lane : 0;
  const int slot = slot_of(c10, ncls);
  const double* lut = io.lut + (size_t)(t - 1) * 4 * kIsingClasses;
  const double ell = lut[slot], lsn = lut[kIsingClasses + slot], lsp = lut[2 * kIsingClasses + slot],
               sig = lut[3 * kIsingClasses + slot];
  const int s = c10 >= ncls ? 1 : -1;
  const int nsum = 2 * (c10 % ncls) - 2 * sh.ndim;
  double fwd[2], rev[2], lF = 0.0, sAs = 0.0, lAs = 0.0;
  int dM = 0, dB = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const double a = on ? (double)A[h] : 0.0, f = on ? (double)F[h] : 0.0;
    fwd[h] = warp_sum((a - f) * lsn + f * lsp);  // sum_active (1-flip) log_sigmoid(-l) + flip log_sigmoid(l)
    rev[h] = warp_sum(a * lsn);                  // reverse logits: -l on flipped sites, l elsewhere
    lF += f * ell;
    sAs += s * a * sig;
    lAs += ell * a * sig;
    if (on) {
      dM += -2 * s * F[h];
      dB += -2 * s * nsum * F[h];
    }
  }
  lF = warp_sum(lF);
  sAs = warp_sum(sAs);
  lAs = warp_sum(lAs);
  dM = __reduce_add_sync(0xffffffffu, dM);
  dB = __reduce_add_sync(0xffffffffu, dB);

  const double h_old = (double)io.field[t - 1], h_new = (double)io.field[t];
  const double Md = (double)M, Bd = (double)B;
  const long long M2 = M + dM, B2 = B + dB;
  // energy(state) = -(B + h M)   (ising.py:85-86); f32 scalars are subtracted in f32 as in update() :143-144
  const float e_init = (float)(-(Bd + h_old * Md));
  const float e_mid = (float)(-(Bd + h_new * Md));
  const float e_fin = (float)(-((double)B2 + h_new * (double)M2));
  const float fwd32 = __fadd_rn((float)fwd[0], (float)fwd[1]);
  const float rev32 = __fadd_rn((float)rev[0], (float)rev[1]);
  const size_t RT = (size_t)sh.R * (sh.T - 1);
  const size_t o = (size_t)rep * (sh.T - 1) + (t - 1);
  float v = 0.f;
  switch (lane) {
    case NQ_IS_WORK: v = __fsub_rn(e_mid, e_init); break;
    case NQ_IS_HEAT: v = __fsub_rn(e_mid, e_fin); break;
    case NQ_IS_FWD: v = fwd32; break;
    case NQ_IS_REV: v = rev32; break;
    case NQ_IS_EP: v = __fsub_rn(fwd32, rev32); break;
    case NQ_IS_MAG: v = (float)((double)M2 / (double)sh.N); break;
    case NQ_IS_ENERGY: v = e_fin; break;
    default: break;
  }
  if (lane < NQ_IS_NSUMMARY) io.summary[lane * RT + o] = v;
  if (io.partials && lane >= 8 && lane < 8 + NQ_IS_NPARTIAL) {
    const double einv = exp(-(double)io.log_temp[t]);
    const double sF = -0.5 * (double)dM;  // sum over flipped sites of s
    double pv;
    switch (lane - 8) {
      case NQ_IS_DFWD_DH: pv = -2.0 * einv * (sF - sAs); break;
      case NQ_IS_DFWD_DL: pv = -(lF - lAs); break;
      case NQ_IS_DREV_DH: pv = 2.0 * einv * sAs; break;
      default: pv = lAs; break;
    }
    io.partials[(lane - 8) * RT + o] = (float)pv;
  }
  if (io.counts && on) {
    int32_t* cnt = io.counts + o * (2 * 2 * kIsingClasses);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      cnt[(h * 2 + 0) * kIsingClasses + slot] = A[h];
      cnt[(h * 2 + 1) * kIsingClasses + slot] = F[h];
    }
  }
  M = M2;
  B = B2;
}

// per-sweep half keys: seed_t = split(seed_r, T-1)[t-1]; (seed_even, seed_odd) = split(seed_t, 2)
__device__ __forceinline__ void sweep_keys(const KeySchedule& ks_rep, int t, int T, bool direct,
                                           KeySchedule& k_first, KeySchedule& k_second) {
  const uint64_t size = 2ull * (T - 1);
  uint32_t s0 = ks_rep.ks0, s1 = ks_rep.ks1;
  if (!direct) {
    s0 = random_bits_at(ks_rep, 2ull * (t - 1), size);
    s1 = random_bits_at(ks_rep, 2ull * (t - 1) + 1, size);
  }
  uint32_t e0, e1, o0, o1;
  split2(s0, s1, e0, e1, o0, o1);
  k_first = key_schedule(e0, e1);   // "even" mask = odd-parity sites, updated first  (ising.py:132-136)
  k_second = key_schedule(o0, o1);
}

// ------------------------------------------------------------------------------ fast kernel
// Shared memory: spread[4][256] | per team: lattice[2][L][NW], thr[2][16] | (CTA team) hist[2][2][2][16]
// GLOBAL_LAT: the colour planes of a replica live in global memory (L2-resident) instead of shared
// memory -- same code, for lattices whose 2 x L x L/64 words exceed 227 KB (L >= 2048).
// Warp teams derive the per-sweep half keys (two words of split(seed, T-1) and their split: four threefry blocks,
// ~280 warp-instructions, 8 % of a 64^2 sweep when every lane repeats them) for kKeyBatch sweeps at a time, lane j
// taking sweep t + j, and read them back from shared memory: ~9 instructions per sweep instead of 280.
#ifndef NQ_ISING_KEY_BATCH
#define NQ_ISING_KEY_BATCH 32
#endif
constexpr int kKeyBatch = NQ_ISING_KEY_BATCH;   // 0: every sweep derives its own keys (the CTA-team form)
static_assert(kKeyBatch == 0 || kKeyBatch == 32, "one lane per batched sweep");

// CTA-per-replica kernels: the half-sweep keys of the next kKeyBatch sweeps, derived by warp 0 (lane j takes sweep
// t + j) into a CTA-shared cache kc[kKeyBatch][4]; every thread then reads its sweep's four words (the warp-team
// kernel does the same per warp, see ising_fast_kernel).  Called by ALL threads of the CTA at every sweep.
__device__ __forceinline__ void cta_sweep_keys(uint32_t* kc, const KeySchedule& ks_rep, int t, int T, bool direct,
                                               KeySchedule& k_first, KeySchedule& k_second, int t_first = 1, int t_end = -1) {
  if (kKeyBatch == 0) {
    sweep_keys(ks_rep, t, T, direct, k_first, k_second);
    return;
  }
  if (t_end < 0) t_end = T;
  const int j = (t - t_first) % kKeyBatch;   // (t_first, t_end): the sweeps of this pass (a segment, or the whole run)
  if (j == 0) {
    __syncthreads();   // every thread has read the previous batch
    if (threadIdx.x < 32 && t + (int)threadIdx.x < t_end) {
      KeySchedule a, b;
      sweep_keys(ks_rep, t + (int)threadIdx.x, T, direct, a, b);
      kc[4 * threadIdx.x + 0] = a.ks0; kc[4 * threadIdx.x + 1] = a.ks1;
      kc[4 * threadIdx.x + 2] = b.ks0; kc[4 * threadIdx.x + 3] = b.ks1;
    }
    __syncthreads();
  }
  k_first = key_schedule(kc[4 * j + 0], kc[4 * j + 1]);
  k_second = key_schedule(kc[4 * j + 2], kc[4 * j + 3]);
}

template <bool WARP_TEAM, bool GLOBAL_LAT = false, bool SEG = false>
__global__ void __launch_bounds__(WARP_TEAM ? 32 * NQ_ISING_TEAMS : 1024, WARP_TEAM ? NQ_ISING_MINB : 1) ising_fast_kernel(const IsingShape sh, const IsingIO io) {
  static_assert(!SEG || !GLOBAL_LAT, "segments are scheduled per CTA with the lattice in shared memory");
  extern __shared__ uint32_t smem[];
  uint32_t* spread = smem;
  const int L = sh.L, NW = sh.NW, T = sh.T;
  const int lat_words = 2 * L * NW;
  constexpr int kThrWords = 2 * 256 * 2;  // two parity buffers of the 256-entry pair table
  const int smem_lat_words = GLOBAL_LAT ? 0 : lat_words;
  // warp teams also keep the half-sweep keys of the next kKeyBatch sweeps (kKeyBatch x 4 words): see the sweep loop
  const int team_words = smem_lat_words + kThrWords + (WARP_TEAM ? 2 * 2 * 2 * kIsingClasses + 4 * kKeyBatch : 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int team = WARP_TEAM ? warp : 0;
  const int tid = WARP_TEAM ? lane : threadIdx.x;          // index within the team
  const int tsize = WARP_TEAM ? 32 : blockDim.x;
  const int nteams = WARP_TEAM ? blockDim.x / 32 : 1;
  uint32_t* team_base = smem + 4 * 256 + team * team_words;
  const long long rep_early = (long long)blockIdx.x * nteams + team;
  uint32_t* lat = GLOBAL_LAT ? sh.lat_global + (size_t)(rep_early < sh.R ? rep_early : 0) * lat_words : team_base;
  // Pair-threshold tables (two sweep-parity buffers).  They depend on the sweep only, so the warp teams of a CTA
  // -- which start every sweep together -- share ONE copy (team 0's area) and build it cooperatively.
  uint2* thr_s = reinterpret_cast<uint2*>((WARP_TEAM ? smem + 4 * 256 : team_base) + smem_lat_words);
  // (segment-scheduled warp teams: R is a multiple of the teams per CTA, every team is live)
  const long long live = SEG ? nteams : sh.R - (long long)blockIdx.x * nteams;
  const int btid = WARP_TEAM ? team * 32 + lane : threadIdx.x;                                   // index among the builders
  const int bsize = WARP_TEAM ? 32 * (int)(live < nteams ? live : nteams) : blockDim.x;
  // per-sweep class histograms [parity][half][active|flipped][16]: per team (warp teams) or per CTA
  int* hist_s = WARP_TEAM ? reinterpret_cast<int*>(team_base + smem_lat_words + kThrWords)
                          : reinterpret_cast<int*>(smem + 4 * 256 + nteams * team_words);

  for (int v = threadIdx.x; v < 4 * 256; v += blockDim.x) spread[v] = spread8_to_nibbles(v & 255) << (v >> 8);
  if (!WARP_TEAM)
    for (int v = threadIdx.x; v < 2 * 2 * 2 * kIsingClasses; v += blockDim.x) hist_s[v] = 0;
  __syncthreads();
  auto team_sync = [&]() {
    if (WARP_TEAM) __syncwarp(); else __syncthreads();
  };
  const size_t ext_words = (size_t)L * L / 32;
  __shared__ unsigned ticket_s;
  for (;;) {  // one pass per (replica, segment); a single pass when !SEG
  long long rep = (long long)blockIdx.x * nteams + team;
  int seg = 0, t_first = 1, t_end = T;
  if (SEG) {
    // ticket it -> segment it / U of work unit it % U; a unit is a replica (CTA team) or the nteams replicas a CTA of
    // warp teams runs in lock step
    const unsigned U = sh.seg_units;
    const unsigned total = U * (unsigned)sh.n_seg;
    if (threadIdx.x == 0) {
      const unsigned it = atomicAdd(&sh.seg_ctl[0], 1u);
      if (it < total && it >= U) {
        // the previous segment of this unit was handed out U tickets ago: it is running or done
        const volatile unsigned* done = sh.seg_ctl + 1 + it % U;
        const unsigned want = it / U;
        // bounded wait (NQ_WAIT_POLLS; 0 = unbounded): when it runs out the launch is abandoned cleanly --
        // status word set, every CTA leaves at its next ticket, ising_status_kernel turns the summaries into
        // NaN -- instead of a __trap() that would poison the caller's CUDA context
        volatile unsigned* dead = sh.seg_ctl + 1 + U;
        unsigned polls = 0;
        bool lost = false;
        while (*done < want) {
          __nanosleep(256);
          if ((sh.seg_max_polls && ++polls > sh.seg_max_polls) || *dead) { lost = true; break; }
        }
        __threadfence();
        if (lost) atomicExch(sh.seg_ctl + 1 + U, 1u);
      }
      ticket_s = (it < total && *(volatile unsigned*)(sh.seg_ctl + 1 + U) == 0u) ? it : 0xffffffffu;
    }
    __syncthreads();
    const unsigned it = ticket_s;
    __syncthreads();
    if (it >= total) return;
    rep = (long long)(it % U) * nteams + team;
    seg = it / U;
    t_first = 1 + (int)((long long)seg * (T - 1) / sh.n_seg);
    t_end = 1 + (int)((long long)(seg + 1) * (T - 1) / sh.n_seg);
  } else if (rep >= sh.R) {
    return;  // whole team leaves together (warp-uniform / CTA-uniform)
  }

  // ---- load the lattice (row-major bits -> colour planes)
  if (SEG && seg > 0) {
    const uint32_t* src = sh.seg_lat + (size_t)rep * lat_words;
    for (int u = tid; u < lat_words; u += tsize) lat[u] = __ldcg(src + u);  // written by another SM: skip L1
  } else {
    const uint32_t* ext0 = io.spins0 + (sh.broadcast0 ? 0 : rep * ext_words);
    for (int u = tid; u < L * NW; u += tsize) {
      const int r = u / NW, w = u - r * NW;
      uint32_t c0, c1;
      split_colours(r, ext0[r * (L / 32) + 2 * w], ext0[r * (L / 32) + 2 * w + 1], c0, c1);
      lat[u] = c0;
      lat[L * NW + u] = c1;
    }
  }
  if (t_first < t_end)  // first sweep of the pass -> the buffer of its parity
    build_pair_table(thr_s + (t_first & 1) * 256, io.thr + (size_t)(t_first - 1) * kIsingClasses, btid, bsize);
  if (WARP_TEAM) __syncthreads(); else team_sync();

  // ---- initial spin sum M and bond sum B (each bond joins a colour-0 and a colour-1 site)
  long long M, B;
  if (SEG && seg > 0) {
    M = __ldcg(sh.seg_mb + 2 * rep);
    B = __ldcg(sh.seg_mb + 2 * rep + 1);
    if (kDebugChecks) {   // the lattice handed over by the previous segment must be the one (M, B) describe
      int up = 0;
      for (int u = tid; u < 2 * L * NW; u += tsize) up += __popc(lat[u]);
      up = __reduce_add_sync(0xffffffffu, up);
      if (WARP_TEAM) {
        NQ_DBG_ASSERT(M == 2ll * up - (long long)L * L);
      } else {
        __shared__ int dbg_up_s;
        if (threadIdx.x == 0) dbg_up_s = 0;
        __syncthreads();
        if (lane == 0) atomicAdd(&dbg_up_s, up);
        __syncthreads();
        NQ_DBG_ASSERT(M == 2ll * dbg_up_s - (long long)L * L);
        __syncthreads();
      }
    }
  } else {
    int hA[10], hF[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) hA[k] = hF[k] = 0;
    int up = 0;
    for (int u = tid; u < L * NW; u += tsize) {
      const int r = u / NW, w = u - r * NW;
      Planes p;
      load_planes(lat, sh, 0, r, w, p);
      histogram(p, 0u, hA, hF);
      up += __popc(lat[u]) + __popc(lat[L * NW + u]);
    }
    int b = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) b += (k >= 5 ? 1 : -1) * (2 * (k % 5) - 4) * hA[k];
    up = __reduce_add_sync(0xffffffffu, up);
    b = __reduce_add_sync(0xffffffffu, b);
    if (!WARP_TEAM) {
      int* tmp = hist_s + 2 * 2 * 2 * kIsingClasses;  // 2 scratch ints
      if (threadIdx.x < 2) tmp[threadIdx.x] = 0;
      __syncthreads();
      if (lane == 0) { atomicAdd(&tmp[0], up); atomicAdd(&tmp[1], b); }
      __syncthreads();
      up = tmp[0];
      b = tmp[1];
    }
    M = 2ll * up - (long long)L * L;
    B = b;
  }

  const KeySchedule ks_rep = key_schedule(io.seeds[2 * rep], io.seeds[2 * rep + 1]);
  const int units = (L / 2) * NW;
  for (int t = t_first; t < t_end; ++t) {
    // the teams of a CTA start every sweep together: they then run the same code (instruction cache) and the
    // shared threshold table of this sweep is complete / the one of the previous sweep is free
    if (WARP_TEAM) __syncthreads();
    KeySchedule kh[2];
    if (WARP_TEAM && kKeyBatch > 0) {
      uint32_t* kc = reinterpret_cast<uint32_t*>(hist_s + 2 * 2 * 2 * kIsingClasses);   // [kKeyBatch][4]
      const int j = (t - t_first) % kKeyBatch;
      if (j == 0) {
        __syncwarp();   // the previous batch has been read by every lane
        if (t + lane < t_end) {
          KeySchedule a, b;
          sweep_keys(ks_rep, t + lane, T, sh.direct_keys != 0, a, b);
          kc[4 * lane + 0] = a.ks0; kc[4 * lane + 1] = a.ks1; kc[4 * lane + 2] = b.ks0; kc[4 * lane + 3] = b.ks1;
        }
        __syncwarp();
      }
      kh[0] = key_schedule(kc[4 * j + 0], kc[4 * j + 1]);
      kh[1] = key_schedule(kc[4 * j + 2], kc[4 * j + 3]);
    } else if (!WARP_TEAM) {
      cta_sweep_keys(reinterpret_cast<uint32_t*>(hist_s + 2 * 2 * 2 * kIsingClasses + 2), ks_rep, t, T, sh.direct_keys != 0,
                     kh[0], kh[1], t_first, t_end);
    } else {
      sweep_keys(ks_rep, t, T, sh.direct_keys != 0, kh[0], kh[1]);
    }
    const uint2* thr_cur = thr_s + (t & 1) * 256;
    if (t + 1 < t_end) build_pair_table(thr_s + ((t + 1) & 1) * 256, io.thr + (size_t)t * kIsingClasses, btid, bsize);
    // warp teams share ONE copy of the half-sweep code between the colours (their phases differ, so the
    // smaller footprint matters to the instruction cache: +1.6 %); the CTA team runs in lock step and is
    // faster with the two colours unrolled (+2.5 %)
    constexpr int kColourUnroll = WARP_TEAM ? 1 : 2;
#pragma unroll kColourUnroll
    for (int c = 0; c < 2; ++c) {
      int At[10], Ft[10];
      int hA[10], hF[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) hA[k] = hF[k] = 0;
      for (int u = tid; u < units; u += tsize) {
        const int i = u / NW, w = u - i * NW;
        update_unit<WARP_TEAM ? NQ_ISING_WARP_MODE : NQ_ISING_CTA_MODE>(lat, sh, spread, thr_cur, kh[c], c, i, w, hA, hF);
      }
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        At[k] = __reduce_add_sync(0xffffffffu, hA[k]);
        Ft[k] = __reduce_add_sync(0xffffffffu, hF[k]);
      }
      if (lane == 0) {
        int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          if (WARP_TEAM) {
            hs[k] = At[k];
            hs[kIsingClasses + k] = Ft[k];
          } else {
            if (At[k]) atomicAdd(&hs[k], At[k]);
            if (Ft[k]) atomicAdd(&hs[kIsingClasses + k], Ft[k]);
          }
        }
      }
      team_sync();
    }
    if (io.states) {  // return_states=True: the lattice after this sweep, row-major bits
      uint32_t* dst = io.states + ((size_t)rep * (T - 1) + (t - 1)) * ext_words;
      for (int u = tid; u < L * NW; u += tsize) {
        const int r = u / NW, w = u - r * NW;
        uint32_t lo, hi;
        merge_colours(r, lat[u], lat[L * NW + u], lo, hi);
        dst[r * (L / 32) + 2 * w] = lo;
        dst[r * (L / 32) + 2 * w + 1] = hi;
      }
      if (!WARP_TEAM) __syncthreads();  // the next sweep rewrites the lattice
    }
    if (WARP_TEAM || warp == 0) {
      int a2[2], f2[2];
      if (WARP_TEAM) __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
        a2[c] = lane < 10 ? hs[lane] : 0;
        f2[c] = lane < 10 ? hs[kIsingClasses + lane] : 0;
      }
      __syncwarp();
      if (!WARP_TEAM) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
          if (lane < 10) { hs[lane] = 0; hs[kIsingClasses + lane] = 0; }
        }
      }
      emit_summary(a2, f2, io, sh, rep, t, M, B, lane);
    }
  }
  team_sync();
  if (SEG && seg + 1 < sh.n_seg) {
    uint32_t* dst = sh.seg_lat + (size_t)rep * lat_words;
    for (int u = tid; u < lat_words; u += tsize) dst[u] = lat[u];
    if (tid == 0) {  // the warp that ran emit_summary carries M, B (warp 0 of a CTA team; every warp team's own)
      sh.seg_mb[2 * rep] = M;
      sh.seg_mb[2 * rep + 1] = B;
    }
  } else if (io.spins_out) {
    uint32_t* dst = io.spins_out + (size_t)rep * ext_words;
    for (int u = tid; u < L * NW; u += tsize) {
      const int r = u / NW, w = u - r * NW;
      uint32_t lo, hi;
      merge_colours(r, lat[u], lat[L * NW + u], lo, hi);
      dst[r * (L / 32) + 2 * w] = lo;
      dst[r * (L / 32) + 2 * w + 1] = hi;
    }
  }
  if (!SEG) break;
  __syncthreads();  // every thread's stores are ordered before thread 0's fence + publish
  if (threadIdx.x == 0) {
    __threadfence();
    atomicExch(&sh.seg_ctl[1 + rep / nteams], (unsigned)(seg + 1));
  }
  }
}

// an abandoned segment-scheduled launch (bounded wait ran out) must not look like a result: NaN summaries
__global__ void ising_status_kernel(const unsigned* status, float* summary, size_t n) {
  if (*status == 0u) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    summary[i] = __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------ bit-packed 3-D kernel
// The fast path for cubic / cuboid lattices d0 x d1 x d2 (the reference's own 3-D test shape is 64^3,
// tests/ising_test.py:71; sum_neighbors / even_odd_masks are rank-generic, ising.py:75-106).  Same design as the 2-D
// CTA-team kernel -- lattice bit-packed in shared memory as two colour planes, a 32-bit word = 32 simultaneously
// updatable sites of one row, jax's counter pairing (p, p + N/2) = rows (i, j) and (i + d0/2, j) served by ONE threefry
// block, integer thresholds, bit-sliced neighbour counts, class histograms by popc -- with "rows" = the d0 * d1
// lines along the last axis, six neighbour words per word (j -+ 1, i -+ 1 whole words; k -+ 1 in-row, one of them
// funnel-shifted) and 2 x 7 site classes (spin, #up-neighbours 0..6).
// Requires d2 % 64 == 0 (whole words per colour per row), d1 even, d0 % 4 == 0 (the counter partner has the same
// colour) and 2 d0 d1 d2 / 64 words of shared memory.
struct Shape3 {
  int d0, d1, NW;   // NW = d2 / 64 words per row per colour
};

__device__ __forceinline__ void count_planes6(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, Planes& p) {
  const uint32_t s1 = a ^ b ^ c, k1 = (a & b) | (c & (a ^ b));   // a + b + c = s1 + 2 k1
  const uint32_t s2 = d ^ e ^ f, k2 = (d & e) | (f & (d ^ e));
  p.C0 = s1 ^ s2;
  const uint32_t t = s1 & s2;
  p.C1 = k1 ^ k2 ^ t;
  p.C2 = (k1 & k2) | (t & (k1 ^ k2));
}

__device__ __forceinline__ void load_planes3(const uint32_t* lat, const Shape3& g, int c, int i, int j, int w, Planes& p) {
  const int NW = g.NW, rows = g.d0 * g.d1, r = i * g.d1 + j;
  const uint32_t* own = lat + (size_t)c * rows * NW;
  const uint32_t* oth = lat + (size_t)(1 - c) * rows * NW;
  const int ju = j == 0 ? g.d1 - 1 : j - 1, jd = j == g.d1 - 1 ? 0 : j + 1;
  const int iu = i == 0 ? g.d0 - 1 : i - 1, id = i == g.d0 - 1 ? 0 : i + 1;
  const uint32_t X = oth[r * NW + w];
  uint32_t Xs;
  if ((((i + j) & 1) == 0) == (c == 0)) {  // second in-row neighbour is site k+1
    const uint32_t Xn = oth[r * NW + (w + 1 == NW ? 0 : w + 1)];
    Xs = __funnelshift_r(X, Xn, 1);
  } else {  // site k-1
    const uint32_t Xp = oth[r * NW + (w == 0 ? NW - 1 : w - 1)];
    Xs = __funnelshift_l(Xp, X, 1);
  }
  p.S = own[r * NW + w];
  count_planes6(oth[(i * g.d1 + ju) * NW + w], oth[(i * g.d1 + jd) * NW + w], oth[(iu * g.d1 + j) * NW + w],
                oth[(id * g.d1 + j) * NW + w], X, Xs, p);
}

// per-class popcounts, 7 neighbour counts per spin value: h[k] spin down, h[7 + k] spin up
__device__ __forceinline__ void histogram7(const Planes& p, uint32_t F, int (&hA)[14], int (&hF)[14]) {
  const uint32_t n2 = ~p.C2, n1 = ~p.C1, n0 = ~p.C0;
  const uint32_t E[7] = {n2 & n1 & n0, n2 & n1 & p.C0, n2 & p.C1 & n0, n2 & p.C1 & p.C0,
                         p.C2 & n1 & n0, p.C2 & n1 & p.C0, p.C2 & p.C1 & n0};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const uint32_t up = E[k] & p.S, down = E[k] & ~p.S;
    hA[k] += __popc(down);
    hA[7 + k] += __popc(up);
    hF[k] += __popc(down & F);
    hF[7 + k] += __popc(up & F);
  }
}

// word w of rows (i, j) and (i + d0/2, j) of colour c: 32 threefry blocks, 64 site updates
__device__ __forceinline__ void update_unit3(uint32_t* lat, const Shape3& g, const uint32_t* spread, const uint2* thr2,
                                             const KeySchedule& ks, uint32_t one, int c, int i, int j, int w,
                                             int (&hA)[14], int (&hF)[14]) {
  const int NW = g.NW, rows = g.d0 * g.d1, i2 = i + g.d0 / 2, d2 = 64 * NW;
  Planes pa, pb;
  load_planes3(lat, g, c, i, j, w, pa);
  load_planes3(lat, g, c, i2, j, w, pb);
  const int rp = (i + j) & 1;                      // (i2 + j) has the same parity: d0 / 2 is even
  const int off = c == 0 ? 1 - rp : rp;
  const uint32_t cbase = (uint32_t)(i * g.d1 + j) * d2 + 2u * (32u * w) + off;
  const uint32_t half = (uint32_t)rows * d2 / 2;
  uint32_t Fa = 0, Fb = 0;
#pragma unroll 1
  for (int q = 3; q >= 0; --q) {
    const uint32_t Na = class_nibbles(spread, pa, q), Nb = class_nibbles(spread, pb, q);
#pragma unroll
    for (int jj = 3; jj >= 0; --jj) {
      const uint2 ta = thr_pair(thr2, Na, jj), tb = thr_pair(thr2, Nb, jj);
#pragma unroll
      for (int e = 1; e >= 0; --e) {
        const uint32_t ctr = cbase + 2u * (8u * q + 2u * jj + e);
        uint32_t y0, y1;
        threefry2x32_allmad(ks, ctr, ctr + half, y0, y1, one);
        const uint32_t da = (y0 >> 9) - (e ? ta.y : ta.x);
        const uint32_t db = (y1 >> 9) - (e ? tb.y : tb.x);
        Fa = __funnelshift_l(da, Fa, 1);
        Fb = __funnelshift_l(db, Fb, 1);
      }
    }
  }
  uint32_t* own = lat + (size_t)c * rows * NW;
  own[(i * g.d1 + j) * NW + w] = pa.S ^ Fa;
  own[(i2 * g.d1 + j) * NW + w] = pb.S ^ Fb;
  histogram7(pa, Fa, hA, hF);
  histogram7(pb, Fb, hA, hF);
}

// Shared memory: spread[4][256] | lattice[2][d0 d1][NW] | pair thresholds[2][256] (uint2) | hist[2][2][2][16] | 2 ints
__global__ void __launch_bounds__(1024, 1) ising_fast3d_kernel(const IsingShape sh, const IsingIO io) {
  extern __shared__ uint32_t smem[];
  const Shape3 g{sh.d0, sh.d1, sh.d2 / 64};
  const int NW = g.NW, rows = g.d0 * g.d1, d2 = sh.d2, T = sh.T;
  const int plane_words = rows * NW;
  uint32_t* spread = smem;
  uint32_t* lat = smem + 4 * 256;
  uint2* thr_s = reinterpret_cast<uint2*>(lat + 2 * plane_words);
  int* hist_s = reinterpret_cast<int*>(lat + 2 * plane_words + 2 * 256 * 2);
  uint32_t* kc = reinterpret_cast<uint32_t*>(hist_s + 2 * 2 * 2 * kIsingClasses + 2);   // [kKeyBatch][4] sweep keys
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x, tsize = blockDim.x;
  const long long rep = blockIdx.x;
  const size_t ext_words = (size_t)sh.N / 32;
  const int row_ext = d2 / 32;

  for (int v = tid; v < 4 * 256; v += tsize) spread[v] = spread8_to_nibbles(v & 255) << (v >> 8);
  for (int v = tid; v < 2 * 2 * 2 * kIsingClasses + 2; v += tsize) hist_s[v] = 0;
  // ---- load the lattice (row-major bits -> colour planes); colour 0 = odd index sum (the reference's "even" mask)
  const uint32_t* ext0 = io.spins0 + (sh.broadcast0 ? 0 : rep * ext_words);
  for (int u = tid; u < plane_words; u += tsize) {
    const int r = u / NW, w = u - r * NW;
    uint32_t c0, c1;
    split_colours((r / g.d1 + r % g.d1) & 1, ext0[r * row_ext + 2 * w], ext0[r * row_ext + 2 * w + 1], c0, c1);
    lat[u] = c0;
    lat[plane_words + u] = c1;
  }
  if (T > 1) build_pair_table(thr_s + 256, io.thr, tid, tsize);   // sweep t = 1 -> parity buffer 1
  __syncthreads();

  // ---- initial spin sum M and bond sum B (each bond joins a colour-0 and a colour-1 site)
  long long M, B;
  {
    int hA[14], hF[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) hA[k] = hF[k] = 0;
    int up = 0;
    for (int u = tid; u < plane_words; u += tsize) {
      const int r = u / NW, w = u - r * NW;
      Planes p;
      load_planes3(lat, g, 0, r / g.d1, r % g.d1, w, p);
      histogram7(p, 0u, hA, hF);
      up += __popc(lat[u]) + __popc(lat[plane_words + u]);
    }
    int b = 0;
#pragma unroll
    for (int k = 0; k < 14; ++k) b += (k >= 7 ? 1 : -1) * (2 * (k % 7) - 6) * hA[k];
    up = __reduce_add_sync(0xffffffffu, up);
    b = __reduce_add_sync(0xffffffffu, b);
    int* tmp = hist_s + 2 * 2 * 2 * kIsingClasses;
    if (lane == 0) { atomicAdd(&tmp[0], up); atomicAdd(&tmp[1], b); }
    __syncthreads();
    M = 2ll * tmp[0] - (long long)sh.N;
    B = tmp[1];
  }

  const KeySchedule ks_rep = key_schedule(io.seeds[2 * rep], io.seeds[2 * rep + 1]);
  const int units = (g.d0 / 2) * g.d1 * NW;
  for (int t = 1; t < T; ++t) {
    KeySchedule kh[2];
    cta_sweep_keys(kc, ks_rep, t, T, sh.direct_keys != 0, kh[0], kh[1]);
    const uint2* thr_cur = thr_s + (t & 1) * 256;
    if (t + 1 < T) build_pair_table(thr_s + ((t + 1) & 1) * 256, io.thr + (size_t)t * kIsingClasses, tid, tsize);
    for (int c = 0; c < 2; ++c) {
      int hA[14], hF[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) hA[k] = hF[k] = 0;
      for (int u = tid; u < units; u += tsize) {
        const int r = u / NW, w = u - r * NW;
        update_unit3(lat, g, spread, thr_cur, kh[c], sh.one, c, r / g.d1, r % g.d1, w, hA, hF);
      }
      int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        const int a = __reduce_add_sync(0xffffffffu, hA[k]), f = __reduce_add_sync(0xffffffffu, hF[k]);
        if (lane == 0) {
          if (a) atomicAdd(&hs[k], a);
          if (f) atomicAdd(&hs[kIsingClasses + k], f);
        }
      }
      __syncthreads();
    }
    if (io.states) {  // return_states=True: the lattice after this sweep, row-major bits
      uint32_t* dst = io.states + ((size_t)rep * (T - 1) + (t - 1)) * ext_words;
      for (int u = tid; u < plane_words; u += tsize) {
        const int r = u / NW, w = u - r * NW;
        uint32_t lo, hi;
        merge_colours((r / g.d1 + r % g.d1) & 1, lat[u], lat[plane_words + u], lo, hi);
        dst[r * row_ext + 2 * w] = lo;
        dst[r * row_ext + 2 * w + 1] = hi;
      }
    }
    if (warp == 0) {
      int a2[2], f2[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
        a2[c] = lane < 14 ? hs[lane] : 0;
        f2[c] = lane < 14 ? hs[kIsingClasses + lane] : 0;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
        if (lane < 14) { hs[lane] = 0; hs[kIsingClasses + lane] = 0; }
      }
      emit_summary(a2, f2, io, sh, rep, t, M, B, lane);
    }
    __syncthreads();   // (states written, histograms of this parity cleared before they are used again)
  }
  if (io.spins_out) {
    uint32_t* dst = io.spins_out + (size_t)rep * ext_words;
    for (int u = tid; u < plane_words; u += tsize) {
      const int r = u / NW, w = u - r * NW;
      uint32_t lo, hi;
      merge_colours((r / g.d1 + r % g.d1) & 1, lat[u], lat[plane_words + u], lo, hi);
      dst[r * row_ext + 2 * w] = lo;
      dst[r * row_ext + 2 * w + 1] = hi;
    }
  }
}

static size_t fast3d_smem_bytes(const int* dims) {
  const size_t plane_words = (size_t)dims[0] * dims[1] * (dims[2] / 64);
  return (4 * 256 + 2 * plane_words + 2 * 256 * 2 + 2 * 2 * 2 * kIsingClasses + 2 + 4 * kKeyBatch) * 4;
}
static bool fast3d_ok(int ndim, const int* dims) {
  return ndim == 3 && dims[2] % 64 == 0 && dims[1] % 2 == 0 && dims[0] % 4 == 0 && fast3d_smem_bytes(dims) <= 227 * 1024 &&
         getenv("NQ_ISING_NO_FAST3D") == nullptr;
}

// ------------------------------------------------------------------------------ bit-packed 1-D kernel
// Periodic chains of N sites with N % 4 == 0 (the reference's own 1-D test shape is 10 000, tests/ising_test.py:71;
// sum_neighbors / even_odd_masks are rank-generic, ising.py:75-106).  Same design as the 2-D / 3-D kernels: two colour
// planes in shared memory (colour 0 = odd sites, the reference's "even" mask, updated first), bit-sliced neighbour
// counts, integer thresholds through the nibble tables, class histograms by popc, 2 x 3 site classes.  jax's counter
// pairing (p, p + N/2) joins two sites of the same colour because N/2 is even; to make the partner a whole word away
// whatever N is, each colour plane is stored as two HALVES of Ph = N/4 sites, each padded to Wh = ceil(Ph / 32) words
// (padding bits stay zero): site j = h Ph + k of a colour is bit k % 32 of word h Wh + k / 32, and ONE threefry block
// (p, p + N/2) serves bit k of both halves -- half a block per spin-update here too.  Only the two half ends need
// care: the neighbour of a half's last site is bit 0 of the other half (the chain closes through them).
struct Shape1 {
  int Ph, Wh, tail;   // sites per half, words per half, Ph % 32 (0 = the last word is full)
};

// colour plane c: words [c][h][w]; neighbour planes of word w of half h
__device__ __forceinline__ void load_planes1(const uint32_t* lat, const Shape1& g, int c, int h, int w, Planes& p) {
  const uint32_t* own = lat + (size_t)c * 2 * g.Wh;
  const uint32_t* oth = lat + (size_t)(1 - c) * 2 * g.Wh;
  const uint32_t X = oth[h * g.Wh + w];
  const bool last = w == g.Wh - 1;
  uint32_t Xs;
  if (c == 0) {   // odd site p = 2 j + 1: neighbours are the even sites j and j + 1
    const uint32_t NX = last ? oth[(1 - h) * g.Wh] : oth[h * g.Wh + w + 1];
    Xs = (last && g.tail) ? ((X >> 1) | ((NX & 1u) << (g.tail - 1))) : __funnelshift_r(X, NX, 1);
  } else {        // even site p = 2 j: neighbours are the odd sites j and j - 1
    const uint32_t prev = w > 0 ? oth[h * g.Wh + w - 1] >> 31
                                : (oth[(1 - h) * g.Wh + g.Wh - 1] >> ((g.tail ? g.tail : 32) - 1)) & 1u;
    Xs = (X << 1) | prev;
  }
  p.S = own[h * g.Wh + w];
  p.C0 = X ^ Xs;
  p.C1 = X & Xs;
  p.C2 = 0u;
}

// per-class popcounts over the valid bits, 3 neighbour counts per spin value: h[k] spin down, h[3 + k] spin up
__device__ __forceinline__ void histogram3(const Planes& p, uint32_t F, uint32_t valid, int (&hA)[6], int (&hF)[6]) {
  const uint32_t E[3] = {~p.C1 & ~p.C0 & valid, p.C0 & valid, p.C1 & valid};   // C1 and C0 are never both set
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint32_t up = E[k] & p.S, down = E[k] & ~p.S;
    hA[k] += __popc(down);
    hA[3 + k] += __popc(up);
    hF[k] += __popc(down & F);
    hF[3 + k] += __popc(up & F);
  }
}

// word w of both halves of colour c: 32 threefry blocks, up to 64 site updates
__device__ __forceinline__ void update_unit1(uint32_t* lat, const Shape1& g, const uint32_t* spread, const uint2* thr2,
                                             const KeySchedule& ks, uint32_t one, int c, int w, int (&hA)[6],
                                             int (&hF)[6]) {
  Planes pa, pb;
  load_planes1(lat, g, c, 0, w, pa);
  load_planes1(lat, g, c, 1, w, pb);
  const uint32_t valid = (w == g.Wh - 1 && g.tail) ? ((1u << g.tail) - 1u) : 0xffffffffu;
  const uint32_t cbase = 2u * (32u * w) + (c == 0 ? 1u : 0u);   // site of bit 0 of the half-0 word
  const uint32_t half = 4u * (uint32_t)g.Ph / 2u;              // N / 2
  uint32_t Fa = 0, Fb = 0;
#pragma unroll 1
  for (int q = 3; q >= 0; --q) {
    const uint32_t Na = class_nibbles(spread, pa, q), Nb = class_nibbles(spread, pb, q);
#pragma unroll
    for (int jj = 3; jj >= 0; --jj) {
      const uint2 ta = thr_pair(thr2, Na, jj), tb = thr_pair(thr2, Nb, jj);
#pragma unroll
      for (int e = 1; e >= 0; --e) {
        const uint32_t ctr = cbase + 2u * (8u * q + 2u * jj + e);
        uint32_t y0, y1;
        threefry2x32_allmad(ks, ctr, ctr + half, y0, y1, one);
        const uint32_t da = (y0 >> 9) - (e ? ta.y : ta.x);
        const uint32_t db = (y1 >> 9) - (e ? tb.y : tb.x);
        Fa = __funnelshift_l(da, Fa, 1);
        Fb = __funnelshift_l(db, Fb, 1);
      }
    }
  }
  Fa &= valid;   // the padding bits of a half's last word are not sites
  Fb &= valid;
  uint32_t* own = lat + (size_t)c * 2 * g.Wh;
  own[w] = pa.S ^ Fa;
  own[g.Wh + w] = pb.S ^ Fb;
  histogram3(pa, Fa, valid, hA, hF);
  histogram3(pb, Fb, valid, hA, hF);
}

// (plane word, bit) of site p
__device__ __forceinline__ void locate1(const Shape1& g, int p, int& word, int& bit) {
  const int c = (p & 1) ? 0 : 1, j = p >> 1, h = j >= g.Ph ? 1 : 0, k = j - h * g.Ph;
  word = (c * 2 + h) * g.Wh + (k >> 5);
  bit = k & 31;
}
// colour planes -> 32 consecutive sites starting at 32 W (row-major external layout)
__device__ __forceinline__ uint32_t gather_word1(const uint32_t* lat, const Shape1& g, int W, int N) {
  uint32_t out = 0;
  for (int b = 0; b < 32 && 32 * W + b < N; ++b) {
    int word, bit;
    locate1(g, 32 * W + b, word, bit);
    out |= ((lat[word] >> bit) & 1u) << b;
  }
  return out;
}

// Shared memory: spread[4][256] | lattice[2][2][Wh] | pair thresholds[2][256] (uint2) | hist[2][2][2][16] | 2 ints
__global__ void __launch_bounds__(1024, 1) ising_fast1d_kernel(const IsingShape sh, const IsingIO io) {
  extern __shared__ uint32_t smem[];
  const int N = sh.N, T = sh.T;
  const Shape1 g{N / 4, (N / 4 + 31) / 32, (N / 4) % 32};
  const int plane_words = 2 * g.Wh;
  uint32_t* spread = smem;
  uint32_t* lat = smem + 4 * 256;
  uint2* thr_s = reinterpret_cast<uint2*>(lat + 2 * plane_words);
  int* hist_s = reinterpret_cast<int*>(lat + 2 * plane_words + 2 * 256 * 2);
  uint32_t* kc = reinterpret_cast<uint32_t*>(hist_s + 2 * 2 * 2 * kIsingClasses + 2);   // [kKeyBatch][4] sweep keys
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x, tsize = blockDim.x;
  const long long rep = blockIdx.x;
  const int ext_words = (N + 31) / 32;

  for (int v = tid; v < 4 * 256; v += tsize) spread[v] = spread8_to_nibbles(v & 255) << (v >> 8);
  for (int v = tid; v < 2 * 2 * 2 * kIsingClasses + 2; v += tsize) hist_s[v] = 0;
  // ---- load the chain (row-major bits -> halved colour planes, zero padding)
  const uint32_t* ext0 = io.spins0 + (sh.broadcast0 ? 0 : rep * (size_t)ext_words);
  for (int u = tid; u < 2 * plane_words; u += tsize) {
    const int c = u / plane_words, hw = u - c * plane_words, h = hw / g.Wh, w = hw - h * g.Wh;
    uint32_t word = 0;
    for (int b = 0; b < 32; ++b) {
      const int k = 32 * w + b;
      if (k >= g.Ph) break;
      const int p = 2 * (h * g.Ph + k) + (c == 0 ? 1 : 0);
      word |= ((ext0[p >> 5] >> (p & 31)) & 1u) << b;
    }
    lat[u] = word;
  }
  if (T > 1) build_pair_table(thr_s + 256, io.thr, tid, tsize);   // sweep t = 1 -> parity buffer 1
  __syncthreads();

  // ---- initial spin sum M and bond sum B (each bond joins an odd and an even site)
  long long M, B;
  {
    int hA[6], hF[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) hA[k] = hF[k] = 0;
    int up = 0;
    for (int u = tid; u < plane_words; u += tsize) {
      const int h = u / g.Wh, w = u - h * g.Wh;
      const uint32_t valid = (w == g.Wh - 1 && g.tail) ? ((1u << g.tail) - 1u) : 0xffffffffu;
      Planes p;
      load_planes1(lat, g, 0, h, w, p);
      histogram3(p, 0u, valid, hA, hF);
      up += __popc(lat[u]) + __popc(lat[plane_words + u]);
    }
    int b = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) b += (k >= 3 ? 1 : -1) * (2 * (k % 3) - 2) * hA[k];
    up = __reduce_add_sync(0xffffffffu, up);
    b = __reduce_add_sync(0xffffffffu, b);
    int* tmp = hist_s + 2 * 2 * 2 * kIsingClasses;
    if (lane == 0) { atomicAdd(&tmp[0], up); atomicAdd(&tmp[1], b); }
    __syncthreads();
    M = 2ll * tmp[0] - (long long)N;
    B = tmp[1];
  }

  const KeySchedule ks_rep = key_schedule(io.seeds[2 * rep], io.seeds[2 * rep + 1]);
  for (int t = 1; t < T; ++t) {
    KeySchedule kh[2];
    cta_sweep_keys(kc, ks_rep, t, T, sh.direct_keys != 0, kh[0], kh[1]);
    const uint2* thr_cur = thr_s + (t & 1) * 256;
    if (t + 1 < T) build_pair_table(thr_s + ((t + 1) & 1) * 256, io.thr + (size_t)t * kIsingClasses, tid, tsize);
    for (int c = 0; c < 2; ++c) {
      int hA[6], hF[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) hA[k] = hF[k] = 0;
      for (int w = tid; w < g.Wh; w += tsize) update_unit1(lat, g, spread, thr_cur, kh[c], sh.one, c, w, hA, hF);
      int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int a = __reduce_add_sync(0xffffffffu, hA[k]), f = __reduce_add_sync(0xffffffffu, hF[k]);
        if (lane == 0) {
          if (a) atomicAdd(&hs[k], a);
          if (f) atomicAdd(&hs[kIsingClasses + k], f);
        }
      }
      __syncthreads();
    }
    if (io.states) {  // return_states=True: the chain after this sweep, row-major bits
      uint32_t* dst = io.states + ((size_t)rep * (T - 1) + (t - 1)) * ext_words;
      for (int W = tid; W < ext_words; W += tsize) dst[W] = gather_word1(lat, g, W, N);
    }
    if (warp == 0) {
      int a2[2], f2[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
        a2[c] = lane < 6 ? hs[lane] : 0;
        f2[c] = lane < 6 ? hs[kIsingClasses + lane] : 0;
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        int* hs = hist_s + ((t & 1) * 2 + c) * 2 * kIsingClasses;
        if (lane < 6) { hs[lane] = 0; hs[kIsingClasses + lane] = 0; }
      }
      emit_summary(a2, f2, io, sh, rep, t, M, B, lane);
    }
    __syncthreads();   // (states written, histograms of this parity cleared before they are used again)
  }
  if (io.spins_out) {
    uint32_t* dst = io.spins_out + (size_t)rep * ext_words;
    for (int W = tid; W < ext_words; W += tsize) dst[W] = gather_word1(lat, g, W, N);
  }
}

static size_t fast1d_smem_bytes(long long N) {
  const size_t Wh = (size_t)((N / 4 + 31) / 32);
  return (4 * 256 + 4 * Wh + 2 * 256 * 2 + 2 * 2 * 2 * kIsingClasses + 2 + 4 * kKeyBatch) * 4;
}
static bool fast1d_ok(int ndim, const int* dims) {
  return ndim == 1 && dims[0] % 4 == 0 && dims[0] >= 8 && fast1d_smem_bytes(dims[0]) <= 227 * 1024 &&
         getenv("NQ_ISING_NO_FAST1D") == nullptr;
}

// ------------------------------------------------------------------------------ generic kernel
// Any even-sided 1-D / 2-D / 3-D periodic lattice (the reference's tests use 10000, 256x256 and
// 64^3): one CTA per replica, one spin per byte in shared memory (or in a global scratch when the
// lattice does not fit), one threefry block per updated site -- the (p, p+N/2) counter partner may
// have the other colour here.  Correctness path; the tuned path is ising_fast_kernel.
__global__ void __launch_bounds__(256) ising_generic_kernel(const IsingShape sh, const IsingIO io) {
  extern __shared__ uint32_t smem[];
  int* hist_s = reinterpret_cast<int*>(smem);               // [2][2][16]
  int* scratch = hist_s + 2 * 2 * kIsingClasses;            // [2]
  uint32_t* thr_s = reinterpret_cast<uint32_t*>(scratch + 2);  // [16]
  uint32_t* kc = thr_s + kIsingClasses;                        // [kKeyBatch][4] sweep keys (cta_sweep_keys)
  const int T = sh.T, N = sh.N, d0 = sh.d0, d1 = sh.d1, d2 = sh.d2, ndim = sh.ndim;
  const long long rep = blockIdx.x;
  signed char* sp = sh.scratch ? sh.scratch + (size_t)rep * N : reinterpret_cast<signed char*>(kc + 4 * kKeyBatch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ncls = 2 * ndim + 1;
  const size_t ext_words = ((size_t)N + 31) / 32;
  const uint32_t* ext0 = io.spins0 + (sh.broadcast0 ? 0 : rep * ext_words);
  for (int p = threadIdx.x; p < N; p += blockDim.x) sp[p] = ((ext0[p >> 5] >> (p & 31)) & 1u) ? 1 : -1;
  if (threadIdx.x < 2) scratch[threadIdx.x] = 0;
  __syncthreads();
  // (parity of the index sum, sum of the 2*ndim neighbours) of site p
  auto site = [&](int p, int& parity) {
    const int k = p % d2, ij = p / d2, j = ij % d1, i = ij / d1;
    parity = (i + j + k) & 1;
    int n = 0;
    if (ndim >= 1) n += sp[((i == 0 ? d0 - 1 : i - 1) * d1 + j) * d2 + k] + sp[((i == d0 - 1 ? 0 : i + 1) * d1 + j) * d2 + k];
    if (ndim >= 2) n += sp[(i * d1 + (j == 0 ? d1 - 1 : j - 1)) * d2 + k] + sp[(i * d1 + (j == d1 - 1 ? 0 : j + 1)) * d2 + k];
    if (ndim >= 3) n += sp[(i * d1 + j) * d2 + (k == 0 ? d2 - 1 : k - 1)] + sp[(i * d1 + j) * d2 + (k == d2 - 1 ? 0 : k + 1)];
    return n;
  };
  {
    int m = 0, b = 0;
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
      int par;
      const int n = site(p, par);
      m += sp[p];
      if (par) b += sp[p] * n;   // every bond joins an odd-parity and an even-parity site
    }
    m = __reduce_add_sync(0xffffffffu, m);
    b = __reduce_add_sync(0xffffffffu, b);
    if (lane == 0) { atomicAdd(&scratch[0], m); atomicAdd(&scratch[1], b); }
  }
  __syncthreads();
  long long M = scratch[0], B = scratch[1];
  const KeySchedule ks_rep = key_schedule(io.seeds[2 * rep], io.seeds[2 * rep + 1]);
  for (int t = 1; t < T; ++t) {
    KeySchedule kh[2];
    cta_sweep_keys(kc, ks_rep, t, T, sh.direct_keys != 0, kh[0], kh[1]);
    if (threadIdx.x < kIsingClasses) thr_s[threadIdx.x] = io.thr[(size_t)(t - 1) * kIsingClasses + threadIdx.x];
    for (int v = threadIdx.x; v < 2 * 2 * kIsingClasses; v += blockDim.x) hist_s[v] = 0;
    __syncthreads();
    for (int c = 0; c < 2; ++c) {
      const int want = c == 0 ? 1 : 0;  // parity of the index sum updated in this half-sweep
      for (int p = threadIdx.x; p < N; p += blockDim.x) {
        int par;
        const int n = site(p, par);
        if (par != want) continue;
        const int s = sp[p];
        const int slot = (s > 0 ? 8 : 0) + (n + 2 * ndim) / 2;
        const uint32_t bits = random_bits_at(kh[c], (uint64_t)p, (uint64_t)N);
        const bool flip = (bits >> 9) < thr_s[slot];
        atomicAdd(&hist_s[(c * 2 + 0) * kIsingClasses + slot], 1);
        if (flip) {
          atomicAdd(&hist_s[(c * 2 + 1) * kIsingClasses + slot], 1);
          sp[p] = (signed char)(-s);
        }
      }
      __threadfence_block();
      __syncthreads();
    }
    if (io.states) {
      uint32_t* dst = io.states + ((size_t)rep * (T - 1) + (t - 1)) * ext_words;
      for (int w = threadIdx.x; w < (int)ext_words; w += blockDim.x) {
        uint32_t word = 0;
        for (int b = 0; b < 32 && 32 * w + b < N; ++b) word |= (sp[32 * w + b] > 0 ? 1u : 0u) << b;
        dst[w] = word;
      }
    }
    if (warp == 0) {
      int a2[2], f2[2];
      for (int c = 0; c < 2; ++c) {
        const int slot = slot_of(lane < 2 * ncls ? lane : 0, ncls);
        a2[c] = lane < 2 * ncls ? hist_s[(c * 2 + 0) * kIsingClasses + slot] : 0;
        f2[c] = lane < 2 * ncls ? hist_s[(c * 2 + 1) * kIsingClasses + slot] : 0;
      }
      emit_summary(a2, f2, io, sh, rep, t, M, B, lane);
    }
    __syncthreads();
  }
  if (io.spins_out) {
    uint32_t* dst = io.spins_out + (size_t)rep * ext_words;
    for (int w = threadIdx.x; w < (int)ext_words; w += blockDim.x) {
      uint32_t word = 0;
      for (int b = 0; b < 32 && 32 * w + b < N; ++b) word |= (sp[32 * w + b] > 0 ? 1u : 0u) << b;
      dst[w] = word;
    }
  }
}

// ------------------------------------------------------------------------------ utilities
template <typename T>
__global__ void pack_kernel(const T* spins, long long n_sites, uint32_t* bits) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 32 >= n_sites) return;
  uint32_t word = 0;
  for (int b = 0; b < 32 && w * 32 + b < n_sites; ++b) word |= (spins[w * 32 + b] > (T)0 ? 1u : 0u) << b;
  bits[w] = word;
}

template <typename T>
__global__ void unpack_kernel(const uint32_t* bits, long long n_sites, T* spins) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sites) return;
  spins[p] = ((bits[p >> 5] >> (p & 31)) & 1u) ? (T)1 : (T)-1;
}

// random_spins: 2 * (uniform(seed, (L*L,)) < p) - 1    (ising.py:71-72)
__global__ void random_spins_kernel(uint32_t k0, uint32_t k1, long long n_sites, uint32_t thr, uint32_t* bits) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 32 >= n_sites) return;
  const KeySchedule ks = key_schedule(k0, k1);
  uint32_t word = 0;
  for (int b = 0; b < 32 && w * 32 + b < n_sites; ++b)
    word |= ((random_bits_at(ks, (uint64_t)(w * 32 + b), (uint64_t)n_sites) >> 9) < thr ? 1u : 0u) << b;
  bits[w] = word;
}

// out[r] += {sum of spins, sum over bonds of s_i s_j} of a d0 x d1 x d2 periodic lattice (out zeroed by
// the caller; blockIdx.y = replica, blockIdx.x strides over the lattice).  One neighbour per axis
// counts every bond once.
__device__ __forceinline__ void measure_flush(long long m, long long b, long long* out, long long rep) {
  m = __reduce_add_sync(0xffffffffu, (int)m);   // per-thread partial sums fit 32 bits (see the callers' grids)
  b = __reduce_add_sync(0xffffffffu, (int)b);
  __shared__ long long acc[2];
  if (threadIdx.x < 2) acc[threadIdx.x] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicAdd((unsigned long long*)&acc[0], (unsigned long long)m);
    atomicAdd((unsigned long long*)&acc[1], (unsigned long long)b);
  }
  __syncthreads();
  if (threadIdx.x < 2) atomicAdd((unsigned long long*)&out[2 * rep + threadIdx.x], (unsigned long long)acc[threadIdx.x]);
}

// Contiguous axis a multiple of 32: 32 sites per step.  A bond joins equal spins (+1) or different
// spins (-1): per word and axis, 32 - 2 popc(word ^ neighbour word).
__global__ void __launch_bounds__(256) measure_words_kernel(const uint32_t* bits, int ndim, int nA, int nB, int wpl,
                                                            long long* out) {
  const long long rep = blockIdx.y;
  const long long words = (long long)nA * nB * wpl;
  const uint32_t* x = bits + rep * words;
  int m = 0, b = 0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
    const int wi = (int)(w % wpl);
    const long long line = w / wpl;
    const int lb = (int)(line % nB), la = (int)(line / nB);
    const uint32_t v = x[w];
    m += 2 * __popc(v) - 32;
    const uint32_t nxt = x[line * wpl + (wi + 1 == wpl ? 0 : wi + 1)];
    b += 32 - 2 * __popc(v ^ ((v >> 1) | (nxt << 31)));
    if (ndim >= 2) b += 32 - 2 * __popc(v ^ x[((long long)la * nB + (lb + 1 == nB ? 0 : lb + 1)) * wpl + wi]);
    if (ndim >= 3) b += 32 - 2 * __popc(v ^ x[((long long)(la + 1 == nA ? 0 : la + 1) * nB + lb) * wpl + wi]);
  }
  measure_flush(m, b, out, rep);
}

__global__ void __launch_bounds__(256) measure_sites_kernel(const uint32_t* bits, int ndim, int d0, int d1, int d2,
                                                            long long* out) {
  const long long rep = blockIdx.y;
  const int N = d0 * d1 * d2;
  const size_t ext_words = ((size_t)N + 31) / 32;
  const uint32_t* x = bits + rep * ext_words;
  auto spin = [&](int p) { return ((x[p >> 5] >> (p & 31)) & 1u) ? 1 : -1; };
  int m = 0, b = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(p % d2), ij = (int)(p / d2), j = ij % d1, i = ij / d1;
    const int s = spin((int)p);
    m += s;
    int fwd = 0;
    if (ndim >= 1) fwd += spin(((i == d0 - 1 ? 0 : i + 1) * d1 + j) * d2 + k);
    if (ndim >= 2) fwd += spin((i * d1 + (j == d1 - 1 ? 0 : j + 1)) * d2 + k);
    if (ndim >= 3) fwd += spin((i * d1 + j) * d2 + (k == d2 - 1 ? 0 : k + 1));
    b += s * fwd;
  }
  measure_flush(m, b, out, rep);
}

// ------------------------------------------------------------------------------ gradient assembly
// Per replica r: the loss and the four contractions of the per-sweep sensitivities with the schedule
// Jacobians,
//   logP_x[r][j] = loss_r * sum_t dF/dx_t J_x[t][j],   rep_x[r][j] = sum_t dLoss/dx_t J_x[t][j],   x in {log_temp, field}
// with dF/dx_0 = 0, dF/dx_t = partials[DFWD_Dx][r][t-1].  kind 0: loss = total_work, 1: loss =
// total_entropy_production (both closed form from summary / partials / the initial spin and bond
// sums), 2: loss and d loss / d x_t supplied by the caller (torch autograd of an arbitrary LossFn).
// One CTA per replica; the sweeps are tiled through shared memory and every sum over t is a fixed
// tree (deterministic).
struct IsingGradArgs {
  int kind, T, Dl, Dh, broadcast0;
  long long R;
  double n_sites, lt0, ltf, h0;      // schedule end points as float64 of the float32 values
  float exp_lt0, exp_ltf;            // float32(exp(log_temp)) at both ends, for the loss in reference arithmetic
  const float* lt_tab;               // device schedule tables [T] (nullable): when given, the four end-point values
  const float* h_tab;                //   above are read from them instead (device-resident training step)
  const float* summary;              // [7][R][T-1]
  const float* partials;             // [4][R][T-1]
  const long long* MB0;              // [R or 1][2]
  const double* J_lt;                // [T][Dl]
  const double* J_h;                 // [T][Dh]
  const double* dL_dh_ext;           // kind 2: [R][T]
  const double* dL_dl_ext;
  const float* loss_ext;             // kind 2: [R]
  float* loss;                       // [R]
  double* logP_lt;                   // [R][Dl]
  double* logP_h;                    // [R][Dh]
  double* rep_lt;
  double* rep_h;
};

constexpr int kGradTile = 512;

__device__ __forceinline__ double block_sum_128(double v, double* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  return (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]);
}

__global__ void __launch_bounds__(128) ising_grad_kernel(const IsingGradArgs a_in) {
  IsingGradArgs a = a_in;
  if (a.lt_tab) {
    a.lt0 = (double)a.lt_tab[0]; a.ltf = (double)a.lt_tab[a.T - 1]; a.h0 = (double)a.h_tab[0];
    a.exp_lt0 = (float)exp(a.lt0); a.exp_ltf = (float)exp(a.ltf);
  }
  __shared__ double sens_s[4][kGradTile];   // dF/dl, dF/dh, dL/dl, dL/dh of the current tile of sweeps
  __shared__ double red_s[4];
  __shared__ double acc_s[4][32];
  const long long r = blockIdx.x;
  const int T = a.T, Tm1 = T - 1;
  const long long RT = a.R * Tm1;
  const float* S = a.summary + r * Tm1;
  const float* P = a.partials ? a.partials + r * Tm1 : nullptr;
  const double M0 = (double)a.MB0[a.broadcast0 ? 0 : 2 * r], B0 = (double)a.MB0[(a.broadcast0 ? 0 : 2 * r) + 1];
  auto M_after = [&](int i) { return rint((double)S[NQ_IS_MAG * RT + i] * a.n_sites); };  // spin sum after sweep i+1
  auto M_before = [&](int i) { return i == 0 ? M0 : M_after(i - 1); };                       // before sweep i+1

  // ---- loss
  float loss;
  if (a.kind == 2) {
    loss = a.loss_ext[r];
  } else {
    const int plane = a.kind == 0 ? NQ_IS_WORK : NQ_IS_EP;
    double s = 0.0;
    for (int t = threadIdx.x; t < Tm1; t += blockDim.x) s += (double)S[plane * RT + t];
    const float traj = (float)block_sum_128(s, red_s);           // exact sum of the float32 terms, rounded once
    if (a.kind == 0) {
      loss = traj;
    } else {
      const float e_init = (float)(-(B0 + a.h0 * M0));
      const float lp_init = __fdiv_rn(-e_init, a.exp_lt0);
      const float lp_fin = __fdiv_rn(-S[NQ_IS_ENERGY * RT + Tm1 - 1], a.exp_ltf);
      loss = __fadd_rn(__fsub_rn(lp_init, lp_fin), traj);
    }
  }
  if (threadIdx.x == 0) a.loss[r] = loss;
  for (int j = threadIdx.x; j < 4 * 32; j += blockDim.x) acc_s[j / 32][j % 32] = 0.0;
  const double e0 = exp(-a.lt0), ef = exp(-a.ltf);

  for (int t0 = 0; t0 < T; t0 += kGradTile) {
    const int len = min(kGradTile, T - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const int t = t0 + i;
      double fl = 0.0, fh = 0.0, ll = 0.0, lh = 0.0;
      if (t >= 1 && P) {
        fh = (double)P[NQ_IS_DFWD_DH * RT + t - 1];
        fl = (double)P[NQ_IS_DFWD_DL * RT + t - 1];
      }
      if (a.kind == 2) {
        lh = a.dL_dh_ext[r * T + t];
        ll = a.dL_dl_ext[r * T + t];
      } else if (a.kind == 0) {                       // work_t = -(h_t - h_{t-1}) M_{t-1}
        if (t >= 1) lh -= M_before(t - 1);
        if (t <= T - 2) lh += M_before(t);
      } else {
        if (t >= 1) {
          lh = fh - (double)P[NQ_IS_DREV_DH * RT + t - 1];
          ll = fl - (double)P[NQ_IS_DREV_DL * RT + t - 1];
        }
        // end-point term log_prob(initial; params[0]) - log_prob(final; params[T-1]), log_prob = (B + h M) e^{-lt}
        if (t == 0) {
          lh += M0 * e0;
          ll += -(B0 + a.h0 * M0) * e0;
        }
        if (t == T - 1) {
          lh += -M_after(Tm1 - 1) * ef;
          ll += -(double)S[NQ_IS_ENERGY * RT + Tm1 - 1] * ef;
        }
      }
      sens_s[0][i] = fl; sens_s[1][i] = fh; sens_s[2][i] = ll; sens_s[3][i] = lh;
    }
    __syncthreads();
    // The four [len] x [len, D] products of the tile (log-prob / loss term x log-temperature / field Jacobian): one warp
    // per product, lane j owns column j and walks the tile's rows -- every load of a Jacobian row is one coalesced
    // line (the first version gave each warp single columns, 128-byte-strided loads and a shuffle tree per column:
    // 1.1 ms per call at config 3, 3.5 % of the step).  D <= 32 (acc_s).
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
      const bool is_rep = warp >= 2, is_h = (warp & 1) != 0;
      const int D = is_h ? a.Dh : a.Dl;
      const double* Jm = (is_h ? a.J_h : a.J_lt) + (long long)t0 * D + lane;
      const double* v = sens_s[(is_rep ? 2 : 0) + (is_h ? 1 : 0)];
      // D <= 16: the two half-warps take alternate rows (still one coalesced line per row pair)
      const int span = D <= 16 ? 16 : 32, groups = 32 / span, col = lane % span, g = lane / span;
      double s0 = 0.0, s1 = 0.0;   // two chains per lane
      if (col < D) {
        const double* Jc = Jm - lane + col;
        int i = g;
        for (; i + groups < len; i += 2 * groups) {
          s0 += v[i] * Jc[(long long)i * D];
          s1 += v[i + groups] * Jc[(long long)(i + groups) * D];
        }
        if (i < len) s0 += v[i] * Jc[(long long)i * D];
      }
      double sum = s0 + s1;
      if (groups == 2) sum += __shfl_down_sync(0xffffffffu, sum, 16);
      if (lane < D) acc_s[(is_rep ? 2 : 0) + (is_h ? 1 : 0)][lane] += sum;
    }
  }
  __syncthreads();
  const double lw = (double)loss;
  for (int j = threadIdx.x; j < a.Dl; j += blockDim.x) {
    a.logP_lt[r * a.Dl + j] = lw * acc_s[0][j];
    a.rep_lt[r * a.Dl + j] = acc_s[2][j];
  }
  for (int j = threadIdx.x; j < a.Dh; j += blockDim.x) {
    a.logP_h[r * a.Dh + j] = lw * acc_s[1][j];
    a.rep_h[r * a.Dh + j] = acc_s[3][j];
  }
}

static bool fast_path(int L) { return L % 64 == 0; }
static size_t fast_smem_bytes(int L, bool warp_team, int nteams, bool global_lat = false) {
  const size_t lat_words = global_lat ? 0 : 2 * (size_t)L * (L / 64);
  size_t words = 4 * 256 + nteams * (lat_words + 2 * 256 * 2 + (warp_team ? 2 * 2 * 2 * kIsingClasses + 4 * kKeyBatch : 0));
  if (!warp_team) words += 2 * 2 * 2 * kIsingClasses + 2 + 4 * kKeyBatch;   // histograms, 2 scratch ints, sweep-key cache
  return words * 4;
}

// Segment scheduling plan for the CTA-team kernel (see IsingShape::n_seg).  R replicas on `slots`
// resident CTAs take ceil(R / slots) rounds; cutting every replica into n segments makes that
// ceil(R n / slots) / n.  n = 1 (off) unless some n <= 64 saves at least 3 %.
struct SegPlan {
  int n_seg = 1;
  int grid = 0;
  unsigned units = 0;
  size_t ctl_bytes = 0, mb_bytes = 0, lat_bytes = 0;
  size_t bytes() const { return ctl_bytes + mb_bytes + lat_bytes; }
};
template <bool W, bool G, bool S>
__global__ void ising_fast_kernel(const IsingShape, const IsingIO);
static SegPlan plan_segments(int L, long long R, int T) {
  SegPlan p;
  if (R * 64 >= 0x7fffffffll || T - 1 < 2) return p;
  const bool warp_team = L <= 128;
  const int nteams = warp_team ? NQ_ISING_TEAMS : 1;
  // Warp-team segments are OFF unless NQ_ISING_WARP_SEGMENTS=1: measured at config 3 (profiles/r02e_ising_c3_*_full.csv)
  // 33.8 ms segmented against 32.8 ms for the plain 1.73-wave launch.  The SMs of the second, partly filled wave hold
  // three CTAs instead of four and -- the kernel being bound by the ALU pipe, not by latency -- finish them
  // proportionally sooner, so the wave tail the segments were meant to remove costs ~1 % (6.92 CTA-runs per SM
  // against a maximum of 7), while the segmented form adds a CTA-wide ticket barrier per segment.
  if (warp_team && (getenv("NQ_ISING_WARP_SEGMENTS") == nullptr || R % nteams)) return p;   // (lock-stepped teams: whole CTAs only)
  const size_t smem = fast_smem_bytes(L, warp_team, nteams);
  if (smem > 227 * 1024) return p;
  const int units = (L / 2) * (L / 64);
  const int threads = warp_team ? 32 * nteams : (units >= 1024 ? 1024 : ((units + 31) / 32) * 32);
  int dev = 0, sms = 0, occ = 0;
  bool ok = cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
  if (ok && warp_team)
    ok = cudaFuncSetAttribute(ising_fast_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ising_fast_kernel<true, false, true>, threads, smem) == cudaSuccess;
  else if (ok)
    ok = cudaFuncSetAttribute(ising_fast_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ising_fast_kernel<false, false, true>, threads, smem) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();  // no device (host-only symbol checks): no plan
    return p;
  }
  // CTA team with two or more CTAs per SM: a lone CTA in the last round already runs faster -- measured no gain.
  // Warp teams (several CTAs per SM by design): the last, partly filled wave runs with fewer warps per SM and a
  // lower ALU-pipe utilisation (profiles/r01h_ising_warp_full.csv: SMs active 55.6M .. 64.4M cycles), segments pay.
  if (occ < 1 || (!warp_team && occ != 1)) return p;
  const long long slots = (long long)sms * occ;      // resident CTAs
  const long long U = R / nteams;                     // work units (CTAs' worth of replicas)
  if (slots <= 0 || U <= slots || U % slots == 0) return p;
  double best = (double)((U + slots - 1) / slots);
  const int n_max = T - 1 < 64 ? T - 1 : 64;
  for (int n = 2; n <= n_max; ++n) {
    const double rounds = (double)((U * n + slots - 1) / slots) / n;
    if (rounds < 0.97 * best) {
      best = rounds;
      p.n_seg = n;
    }
  }
  if (const char* env = getenv("NQ_ISING_SEGMENTS")) {   // override for experiments (>= 2)
    const int n = atoi(env);
    if (n >= 2 && n <= T - 1) p.n_seg = n;
  }
  if (p.n_seg == 1) return p;
  // keep a segment under ~512 sweeps: a CTA that has to wait for the previous segment of its replica
  // waits at most one segment, which must stay far below the kernel's bounded-wait limit.  Doubling
  // n never lengthens the schedule (ceil(2Rn/s)/(2n) <= ceil(Rn/s)/n).
  while ((T - 1) / p.n_seg > 512 && 2 * p.n_seg <= T - 1 && U * 2 * p.n_seg < 0x7fffffffll) p.n_seg *= 2;
  p.grid = (int)slots;
  p.units = (unsigned)U;
  p.ctl_bytes = (((size_t)U + 2) * 4 + 15) / 16 * 16;
  p.mb_bytes = (size_t)R * 2 * 8;
  p.lat_bytes = (size_t)R * 2 * L * (L / 64) * 4;
  return p;
}
static const size_t kGenericSmemHeader = (2 * 2 * kIsingClasses + 2 + kIsingClasses + 4 * kKeyBatch) * 4;
static size_t generic_smem_bytes(long long N) {
  return kGenericSmemHeader + (((size_t)N + 3) / 4) * 4;
}

// lattice geometry of a descriptor: ndim == 0 means the 2-D square L x L
static int lattice_dims(const NqIsingDesc* d, int dims[3]) {
  int ndim = d->ndim;
  if (ndim == 0) {
    dims[0] = dims[1] = d->L;
    dims[2] = 1;
    return 2;
  }
  dims[0] = d->dims[0];
  dims[1] = ndim >= 2 ? d->dims[1] : 1;
  dims[2] = ndim >= 3 ? d->dims[2] : 1;
  return ndim;
}

}  // namespace nq

using namespace nq;

extern "C" size_t nq_ising_words_per_lattice(int32_t L) { return ((size_t)L * L + 31) / 32; }

// the generic path keeps one byte per spin in a global scratch when the lattice exceeds shared memory
extern "C" size_t nq_ising_workspace_bytes(const NqIsingDesc* desc) {
  if (!desc) return 0;
  int dims[3];
  const int ndim = lattice_dims(desc, dims);
  const long long N = (long long)dims[0] * dims[1] * dims[2];
  if (ndim == 2 && dims[0] == dims[1] && fast_path(dims[0])) {
    if (fast_smem_bytes(dims[0], dims[0] <= 128, dims[0] <= 128 ? NQ_ISING_TEAMS : 1) <= 227 * 1024)
      return plan_segments(dims[0], desc->R, desc->T).bytes();  // optional: without it the run is unsegmented
    return (size_t)desc->R * 2 * dims[0] * (dims[0] / 64) * 4;   // colour planes in global memory
  }
  if (fast3d_ok(ndim, dims)) return 0;   // bit-packed 3-D kernel: lattice in shared memory
  if (fast1d_ok(ndim, dims)) return 0;   // bit-packed 1-D kernel: chain in shared memory
  if (generic_smem_bytes(N) <= 200 * 1024) return 0;
  return (size_t)desc->R * N;
}

extern "C" int nq_ising_run(const NqIsingDesc* desc, const uint32_t* thr, const double* lut, const float* field,
                            const float* log_temp, const uint32_t* spins0, const uint32_t* seeds,
                            uint32_t* spins_out, float* summary, float* partials, int32_t* counts, uint32_t* states,
                            void* workspace, size_t workspace_bytes, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(desc != nullptr, "desc is null");
  int dims[3];
  const int ndim = lattice_dims(desc, dims);
  const int T = desc->T;
  NQ_CHECK_ARG(ndim >= 1 && ndim <= 3, "lattices must be 1-, 2- or 3-dimensional (got ndim=%d)", ndim);
  for (int a = 0; a < ndim; ++a) {
    if (dims[a] <= 0 || (dims[a] & 1)) {  // ising.py:101-103
      if (ndim == 1) set_error("All entries in `shape` must be even; got (%d,).", dims[0]);
      else if (ndim == 2) set_error("All entries in `shape` must be even; got (%d, %d).", dims[0], dims[1]);
      else set_error("All entries in `shape` must be even; got (%d, %d, %d).", dims[0], dims[1], dims[2]);
      return NQ_ERR_INVALID_ARGUMENT;
    }
  }
  const long long N = (long long)dims[0] * dims[1] * dims[2];
  NQ_CHECK_ARG(T >= 1 && desc->R >= 1, "need T >= 1 and R >= 1 (got T=%d R=%lld)", T, (long long)desc->R);
  NQ_CHECK_ARG(thr && lut && field && log_temp && spins0 && seeds && summary, "null required pointer");
  NQ_CHECK_ARG(N <= 0x3fffffffll, "lattice too large");
  NQ_CHECK_ARG(desc->sweep_key_mode == 0 || (desc->sweep_key_mode == 1 && T == 2), "sweep_key_mode 1 needs T == 2");
  const int L = dims[0];
  IsingShape sh{L, L / 64, T, desc->R, desc->spins0_broadcast, desc->sweep_key_mode, ndim, dims[0], dims[1], dims[2],
                (int)N, nullptr, nullptr, 1u, 1, nullptr, nullptr, nullptr, 0u, 0u};
  IsingIO io{thr, lut, field, log_temp, spins0, seeds, spins_out, summary, partials, counts, states};
  cudaStream_t st = (cudaStream_t)stream;
  const bool square2d = ndim == 2 && dims[0] == dims[1];
  const bool warp_team = L <= 128;
  const int nteams = warp_team ? NQ_ISING_TEAMS : 1;
  if (square2d && fast_path(L)) {
    const int units = (L / 2) * (L / 64);
    const int threads = warp_team ? 32 * nteams : (units >= 1024 ? 1024 : ((units + 31) / 32) * 32);
    const bool global_lat = fast_smem_bytes(L, warp_team, nteams) > 227 * 1024;
    const size_t smem = fast_smem_bytes(L, warp_team, nteams, global_lat);
    const unsigned blocks = (unsigned)((desc->R + nteams - 1) / nteams);
    if (global_lat) {
      const size_t need = (size_t)desc->R * 2 * L * (L / 64) * 4;
      if (workspace == nullptr || workspace_bytes < need) {
        set_error("L=%d keeps its lattice in global memory and needs a %zu-byte workspace (nq_ising_workspace_bytes)", L, need);
        return NQ_ERR_WORKSPACE;
      }
      sh.lat_global = (uint32_t*)workspace;
      NQ_CUDA(cudaFuncSetAttribute(ising_fast_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ising_fast_kernel<false, true><<<blocks, threads, smem, st>>>(sh, io);
    } else {
      const SegPlan plan = plan_segments(L, desc->R, T);
      if (plan.n_seg > 1 && workspace != nullptr && workspace_bytes >= plan.bytes() && getenv("NQ_ISING_NO_SEGMENTS") == nullptr) {
        char* ws = (char*)workspace;
        sh.n_seg = plan.n_seg;
        sh.seg_units = plan.units;
        sh.seg_ctl = (unsigned*)ws;
        sh.seg_mb = (long long*)(ws + plan.ctl_bytes);
        sh.seg_lat = (uint32_t*)(ws + plan.ctl_bytes + plan.mb_bytes);
        {
          const char* env = getenv("NQ_WAIT_POLLS");
          sh.seg_max_polls = env ? (unsigned)strtoul(env, nullptr, 10) : (1u << 24);
        }
        NQ_CUDA(cudaMemsetAsync(sh.seg_ctl, 0, plan.ctl_bytes, st));
        if (warp_team) ising_fast_kernel<true, false, true><<<plan.grid, threads, smem, st>>>(sh, io);
        else ising_fast_kernel<false, false, true><<<plan.grid, threads, smem, st>>>(sh, io);
        ising_status_kernel<<<64, 256, 0, st>>>(sh.seg_ctl + 1 + plan.units, io.summary,
                                                (size_t)NQ_IS_NSUMMARY * desc->R * (T - 1));
      } else if (warp_team) {
        NQ_CUDA(cudaFuncSetAttribute(ising_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ising_fast_kernel<true><<<blocks, threads, smem, st>>>(sh, io);
      } else {
        NQ_CUDA(cudaFuncSetAttribute(ising_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ising_fast_kernel<false><<<blocks, threads, smem, st>>>(sh, io);
      }
    }
  } else if (fast3d_ok(ndim, dims)) {
    const size_t smem = fast3d_smem_bytes(dims);
    const int units = (dims[0] / 2) * dims[1] * (dims[2] / 64);
    const int threads = units >= 1024 ? 1024 : ((units + 31) / 32) * 32;
    NQ_CUDA(cudaFuncSetAttribute(ising_fast3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ising_fast3d_kernel<<<(unsigned)desc->R, threads, smem, st>>>(sh, io);
  } else if (fast1d_ok(ndim, dims)) {
    const size_t smem = fast1d_smem_bytes(N);
    const int units = (int)((N / 4 + 31) / 32);
    const int threads = units >= 1024 ? 1024 : ((units + 31) / 32) * 32;
    NQ_CUDA(cudaFuncSetAttribute(ising_fast1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ising_fast1d_kernel<<<(unsigned)desc->R, threads, smem, st>>>(sh, io);
  } else {
    size_t smem = generic_smem_bytes(N);
    if (smem > 200 * 1024) {   // spins live in the caller's workspace (L2-resident), tables stay in shared memory
      const size_t need = (size_t)desc->R * N;
      if (workspace == nullptr || workspace_bytes < need) {
        set_error("lattice of %lld sites needs a %zu-byte workspace (nq_ising_workspace_bytes)", N, need);
        return NQ_ERR_WORKSPACE;
      }
      sh.scratch = (signed char*)workspace;
      smem = kGenericSmemHeader;
    }
    NQ_CUDA(cudaFuncSetAttribute(ising_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ising_generic_kernel<<<(unsigned)desc->R, 256, smem, st>>>(sh, io);
  }
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_random_spins(const uint32_t key[2], int64_t n_sites, float p, uint32_t* out_bits,
                                     void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(key && out_bits && n_sites > 0 && n_sites <= 0x7fffffffll, "bad arguments");
  const long long n = n_sites;
  // u < p  with u = m * 2^-23  <=>  m < ceil(p * 2^23)
  double scaled = ceil((double)p * 8388608.0);
  if (scaled < 0) scaled = 0;
  if (scaled > 8388608.0) scaled = 8388608.0;
  const long long words = (n + 31) / 32;
  random_spins_kernel<<<(unsigned)((words + 127) / 128), 128, 0, (cudaStream_t)stream>>>(key[0], key[1], n, (uint32_t)scaled,
                                                                                       out_bits);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_pack(const void* spins, int32_t is_float, int64_t n_sites, uint32_t* bits, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(spins && bits && n_sites > 0, "bad arguments");
  const long long words = (n_sites + 31) / 32;
  const unsigned blocks = (unsigned)((words + 127) / 128);
  if (is_float) pack_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)spins, n_sites, bits);
  else pack_kernel<int32_t><<<blocks, 128, 0, (cudaStream_t)stream>>>((const int32_t*)spins, n_sites, bits);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_unpack(const uint32_t* bits, int64_t n_sites, int32_t is_float, void* spins, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(spins && bits && n_sites > 0, "bad arguments");
  const unsigned blocks = (unsigned)((n_sites + 255) / 256);
  if (is_float) unpack_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(bits, n_sites, (float*)spins);
  else unpack_kernel<int32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(bits, n_sites, (int32_t*)spins);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_measure(const uint32_t* bits, int32_t ndim, const int32_t dims[3], int64_t R,
                                int64_t* out_MB, void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(bits && out_MB && dims && ndim >= 1 && ndim <= 3 && R > 0, "bad arguments");
  const int d0 = dims[0], d1 = ndim >= 2 ? dims[1] : 1, d2 = ndim >= 3 ? dims[2] : 1;
  NQ_CHECK_ARG(d0 > 0 && d1 > 0 && d2 > 0, "bad lattice shape");
  const long long N = (long long)d0 * d1 * d2;
  NQ_CHECK_ARG(N <= 0x3fffffffll, "lattice too large (N=%lld)", N);
  cudaStream_t st = (cudaStream_t)stream;
  NQ_CUDA(cudaMemsetAsync(out_MB, 0, (size_t)R * 2 * sizeof(int64_t), st));
  const int last = ndim == 1 ? d0 : (ndim == 2 ? d1 : d2);
  const size_t ext_words = ((size_t)N + 31) / 32;
  for (int64_t r0 = 0; r0 < R; r0 += 65535) {   // gridDim.y limit
    const unsigned nr = (unsigned)(R - r0 < 65535 ? R - r0 : 65535);
    const uint32_t* src = bits + (size_t)r0 * ext_words;
    long long* dst = (long long*)out_MB + 2 * r0;
    if (last % 32 == 0) {
      // axes padded at the front: (nA, nB, last); a thread sums at most ~2^7 words of 32 sites x 3 axes
      const int nA = ndim == 3 ? d0 : 1, nB = ndim == 3 ? d1 : (ndim == 2 ? d0 : 1);
      long long bx = (N / 32 + 255) / 256;
      if (bx > 1024) bx = 1024;
      measure_words_kernel<<<dim3((unsigned)bx, nr), 256, 0, st>>>(src, ndim, nA, nB, last / 32, dst);
    } else {
      long long bx = (N + 255) / 256;
      if (bx > 1024) bx = 1024;
      measure_sites_kernel<<<dim3((unsigned)bx, nr), 256, 0, st>>>(src, ndim, d0, d1, d2, dst);
    }
  }
  NQ_LAUNCHED();
  return NQ_OK;
}

static int grad_assemble(int32_t kind, int32_t T, int64_t R, int64_t n_sites, int32_t broadcast0,
                         const float* log_temp_ends /*host [2]*/, float field0, const float* lt_tab, const float* h_tab,
                         const float* summary, const float* partials, const int64_t* MB0,
                         const double* J_lt, int32_t Dl, const double* J_h, int32_t Dh,
                         const double* dL_dh_ext, const double* dL_dl_ext, const float* loss_ext,
                         float* loss, double* logP_lt, double* logP_h, double* rep_lt, double* rep_h, void* stream) {
  NQ_CHECK_ARG(kind >= 0 && kind <= 2, "kind must be 0 (work), 1 (entropy production) or 2 (external)");
  NQ_CHECK_ARG(T >= 2 && R >= 1 && n_sites >= 1, "need T >= 2, R >= 1, n_sites >= 1");
  // the spin sum of a sweep is recovered as rint(float32(M / N) * N): exact while N <= 2^23 or N is a power of two
  NQ_CHECK_ARG(n_sites <= (1ll << 23) || (n_sites & (n_sites - 1)) == 0,
               "lattices of more than 2^23 sites must have a power-of-two site count (got %lld)", (long long)n_sites);
  NQ_CHECK_ARG(Dl >= 1 && Dl <= 32 && Dh >= 1 && Dh <= 32, "1 <= variables per schedule component <= 32 (got %d, %d)", Dl, Dh);
  NQ_CHECK_ARG((log_temp_ends || (lt_tab && h_tab)) && summary && MB0 && J_lt && J_h && loss && logP_lt && logP_h && rep_lt && rep_h,
               "null pointer");
  NQ_CHECK_ARG(kind == 2 || partials || kind == 0, "partials are required for the entropy-production loss");
  NQ_CHECK_ARG(kind != 1 || partials, "partials are required for the entropy-production loss");
  NQ_CHECK_ARG(kind != 2 || (dL_dh_ext && dL_dl_ext && loss_ext), "kind 2 needs loss_ext, dL_dh_ext, dL_dl_ext");
  IsingGradArgs a{};
  a.kind = kind; a.T = T; a.Dl = Dl; a.Dh = Dh; a.broadcast0 = broadcast0; a.R = R;
  a.n_sites = (double)n_sites;
  if (lt_tab) {
    a.lt_tab = lt_tab; a.h_tab = h_tab;
  } else {
    a.lt0 = (double)log_temp_ends[0]; a.ltf = (double)log_temp_ends[1]; a.h0 = (double)field0;
    a.exp_lt0 = (float)exp((double)log_temp_ends[0]); a.exp_ltf = (float)exp((double)log_temp_ends[1]);
  }
  a.summary = summary; a.partials = partials; a.MB0 = (const long long*)MB0;
  a.J_lt = J_lt; a.J_h = J_h; a.dL_dh_ext = dL_dh_ext; a.dL_dl_ext = dL_dl_ext; a.loss_ext = loss_ext;
  a.loss = loss; a.logP_lt = logP_lt; a.logP_h = logP_h; a.rep_lt = rep_lt; a.rep_h = rep_h;
  ising_grad_kernel<<<(unsigned)R, 128, 0, (cudaStream_t)stream>>>(a);
  NQ_LAUNCHED();
  return NQ_OK;
}

extern "C" int nq_ising_grad_assemble(int32_t kind, int32_t T, int64_t R, int64_t n_sites, int32_t broadcast0,
                                      const float* log_temp_ends /*host [2]*/, float field0,
                                      const float* summary, const float* partials, const int64_t* MB0,
                                      const double* J_lt, int32_t Dl, const double* J_h, int32_t Dh,
                                      const double* dL_dh_ext, const double* dL_dl_ext, const float* loss_ext,
                                      float* loss, double* logP_lt, double* logP_h, double* rep_lt, double* rep_h,
                                      void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(log_temp_ends != nullptr, "log_temp_ends is null");
  return grad_assemble(kind, T, R, n_sites, broadcast0, log_temp_ends, field0, nullptr, nullptr, summary, partials, MB0,
                       J_lt, Dl, J_h, Dh, dL_dh_ext, dL_dl_ext, loss_ext, loss, logP_lt, logP_h, rep_lt, rep_h, stream);
}

extern "C" int nq_ising_grad_assemble_dev(int32_t kind, int32_t T, int64_t R, int64_t n_sites, int32_t broadcast0,
                                          const float* log_temp /*device [T]*/, const float* field /*device [T]*/,
                                          const float* summary, const float* partials, const int64_t* MB0,
                                          const double* J_lt, int32_t Dl, const double* J_h, int32_t Dh,
                                          float* loss, double* logP_lt, double* logP_h, double* rep_lt, double* rep_h,
                                          void* stream) {
  NQ_RANGE();
  NQ_CHECK_ARG(kind == 0 || kind == 1, "kind must be 0 (work) or 1 (entropy production)");
  NQ_CHECK_ARG(log_temp && field, "schedule tables are null");
  return grad_assemble(kind, T, R, n_sites, broadcast0, nullptr, 0.f, log_temp, field, summary, partials, MB0, J_lt, Dl,
                       J_h, Dh, nullptr, nullptr, nullptr, loss, logP_lt, logP_h, rep_lt, rep_h, stream);
}
