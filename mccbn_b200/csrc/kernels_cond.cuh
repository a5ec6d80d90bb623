// kernels_cond.cuh -- conditional ("truncated-exponential") proposal E-step for sm_100a: the backward
// (Hamming-ball), bernoulli and add-remove schemes of the reference, and the OT-CBN proposal (X == Y).
//
// Replaces, per (observation i, sample l):
//   neighbors / replicate          src/backward_sampling.cpp:37-54, src/mcem_hcbn.cpp:497-513  X = Y ^ flip[l mod n]
//   coin_tossing                   src/backward_sampling.cpp:64-77                              X = Y ^ Bernoulli(eps) mask
//   add-remove proposal            src/mcem_hcbn.cpp:389-459, src/add_remove.cpp:17-141          X = closure(Y +/- one event)
//   generate_mutation_times        src/add_remove.cpp:156-203   T_s ~ Exp(lambda_s); per node in topological order
//                                     m = max_pa T; X_j ? z ~ TExp(lambda_j, 0, T_s - m), T_j = m + z
//                                                       : z ~ Exp(lambda_j),            T_j = max(m, T_s) + z;  Z_j = T_j - m
//   cbn_density_log + weights      src/add_remove.cpp:206-216, src/mcem_hcbn.cpp:524-532, :559-562
//
// The importance weight is evaluated through the algebraic identity (checked against the oracle in
// tests/test_oracle_exact.py::test_k3_density_identity)
//     log P(Z) - log q_Z = sum_{X_j=1} log F_j(T_s - m_j)  -  sum_{X_j=0} lambda_j max(0, T_s - m_j),
// i.e. the lambda*z terms cancel, so no log/exp of the density itself is needed.  A sample is incompatible
// (weight 0, reference: any Z_j < 0) iff some mutated event has an unmutated parent -- pure bit work.
// Weights are continuous here, so each lane keeps a running log2 reference (online log-sum-exp) and the warp
// aligns the lanes before the fp64 reduction.
#pragma once
#include "device_types.cuh"
#include "kernels_forward.cuh"
#include "philox.cuh"

namespace mccbn {

enum CondMode { kCondBackward = 0, kCondBernoulli = 1, kCondOtcbn = 2, kCondAddRemove = 3 };

struct CondArgs {
  const uint64_t* obs_bits;
  const double* times;
  const DevModel* model;
  const uint64_t* ball;  // backward: flip masks of the Hamming ball (event-id bit order), n_ball entries
  double* part;
  int N, L;              // L = samples per observation actually drawn (backward: reps * n_ball)
  int n_ball, reps;      // backward only
  int chunks, chunk_len; // backward: chunk over ball genotypes; bernoulli/otcbn: chunk over samples
  long long n_items;
  long long obs_offset;
  uint32_t key0, key1;
  uint32_t cw;
  int PS;
  unsigned long long* next_item;  // work counter (zeroed before the launch): items cost between nothing and reps x
                                  // |ball| samples (incompatible genotypes are skipped), so warps claim them dynamically
};

template <typename real>
struct CondMath;
template <>
struct CondMath<float> {
  // scale constants per node: lam2 = lambda * log2(e), inv = ln2 / lambda
  static __device__ __forceinline__ float u01(uint32_t b) { return __uint_as_float((b >> 9) | 0x3f800000u) - 1.0f; }  // [0,1)
  // F(a) = 1 - exp(-lambda a); clamped away from 0 so that a compatible sample never loses its (tiny) weight to rounding
  static __device__ __forceinline__ float F_of(float a, float lam2) { return fmaxf(1.0f - fast_ex2(-a * lam2), 1.17549435e-38f); }
  static __device__ __forceinline__ float z_of(float u, float F, float inv) { return -fast_lg2(1.0f - u * F) * inv; }
  static __device__ __forceinline__ float lg2(float x) { return fast_lg2(x); }
  static __device__ __forceinline__ float ex2(float x) { return fast_ex2(x); }
};
template <>
struct CondMath<double> {
  static __device__ __forceinline__ double F_of(double a, double lam2) { return -expm1(-a * lam2 * 0.693147180559945309417); }
  static __device__ __forceinline__ double z_of(double u, double F, double inv) {
    return -log1p(-u * F) * inv * 1.44269504088896340736;
  }
  static __device__ __forceinline__ double lg2(double x) { return log2(x); }
  static __device__ __forceinline__ double ex2(double x) { return exp2(x); }
};

template <typename real>
__device__ __forceinline__ real cond_uniform(const uint4& r, int w) {
  if (sizeof(real) == 4) {
    const uint32_t bits = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
    return (real)CondMath<float>::u01(bits);
  } else {
    const uint32_t hi = w == 0 ? r.x : r.z, lo = w == 0 ? r.y : r.w;
    const uint64_t m = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);
    return (real)((double)m * (1.0 / 9007199254740992.0));  // [0, 1)
  }
}

// One conditional draw.  Xt = hidden genotype (topological bit order).  Returns log2 of P(Z)/q_Z.
// Random stream (fp32): the same packed 24-bit fields as the forward sampler -- field 0 = sampling time, field k + 1 =
// node k -- so p = 25 costs 5 Philox blocks instead of 7.
template <typename real, int PMAX, bool TIMES_GIVEN>
__device__ __forceinline__ real sample_conditional(const PosetProg& pg, const uint32_t tcol_s,
                                                   const real* __restrict__ s_lam2, const real* __restrict__ s_inv,
                                                   const uint32_t l, const uint32_t gi, const uint32_t k0,
                                                   const uint32_t k1, const uint32_t cw, const real ts_given,
                                                   const uint32_t Xt, real (&Z)[PMAX], real& Ts_out) {
  constexpr int UPB = RealTraits<real>::kUniPerBlock;
  constexpr int NB = fwd_blocks_for_fields(PMAX + 1);
  uint32_t W[4 * NB];             // fp32 path
  uint4 r = make_uint4(0, 0, 0, 0);  // fp64 path
  const uint32_t one = pg.one_bits;
  real Ts = ts_given;
  if constexpr (sizeof(real) == 4) {
    const uint4 r0 = philox4x32_10(l, gi, 0u, cw, k0, k1);
    W[0] = r0.x, W[1] = r0.y, W[2] = r0.z, W[3] = r0.w;
    if (!TIMES_GIVEN) Ts = CondMath<real>::z_of((real)field_u01_open1(W, 0, one), (real)1, s_inv[PMAX]);  // Exp(lambda_s)
  } else if (!TIMES_GIVEN) {
    r = philox4x32_10(l, gi, 0u, cw, k0, k1);
    Ts = CondMath<real>::z_of(cond_uniform<real>(r, 0), (real)1, s_inv[PMAX]);
  }
  real lw2 = (real)0;
  const unsigned multi_mask = pg.multi_mask;
  // all PMAX slots run unconditionally; slots k >= p have lam2 = inv = 0 and X_k = 0: they contribute nothing
#pragma unroll
  for (int k = 0; k < PMAX; ++k) {
    real u;
    if constexpr (sizeof(real) == 4) {
      const int b = fwd_field_block(k + 1);
      if (b != fwd_field_block(k)) {  // compile-time after unrolling: block b is first touched by node k
        if (k < pg.p) {               // warp-uniform
          const uint4 rb = philox4x32_10(l, gi, (uint32_t)b, cw, k0, k1);
          W[4 * b + 0] = rb.x, W[4 * b + 1] = rb.y, W[4 * b + 2] = rb.z, W[4 * b + 3] = rb.w;
        } else {
          W[4 * b + 0] = W[4 * b + 1] = W[4 * b + 2] = W[4 * b + 3] = 0u;
        }
      }
      u = (real)field_u01_open1(W, k + 1, one);
    } else {
      const int ui = k + (TIMES_GIVEN ? 0 : 1);
      if (ui % UPB == 0) r = philox4x32_10(l, gi, (uint32_t)(ui / UPB), cw, k0, k1);
      u = cond_uniform<real>(r, ui % UPB);
    }
    const real m = parents_max<real>(pg, multi_mask, k, tcol_s);
    const bool xk = (Xt >> k) & 1u;
    const real a = Ts - m;
    const real Fx = CondMath<real>::F_of(a, s_lam2[k]);
    const real F = xk ? Fx : (real)1;
    const real z = CondMath<real>::z_of(u, F, s_inv[k]);
    const real base = xk ? m : max(m, Ts);
    const real t = base + z;
    sts_at<real>(tcol_s + k * (uint32_t)(kFwdBlock * sizeof(real)), t);
    Z[k] = t - m;
    lw2 += xk ? CondMath<real>::lg2(Fx) : -s_lam2[k] * max(a, (real)0);
  }
  Ts_out = Ts;
  return lw2;
}

// compatibility of a genotype in topological bit order: every mutated event has all parents mutated
template <int PMAX, typename Prog>
__device__ __forceinline__ bool compatible_topo(const Prog& pg, const typename Prog::word_t Xt) {
  typename Prog::word_t bad = 0;
#pragma unroll
  for (int k = 0; k < PMAX; ++k)
    if (k < pg.p) bad |= ((Xt >> k) & 1u) ? (pg.parent_mask[k] & ~Xt) : 0;
  return bad == 0;
}

#ifndef MCCBN_COND_MINB
#define MCCBN_COND_MINB 4  // resident blocks per SM (ncu r1e: 2 blocks left the schedulers at IPC 0.9 of 2.5)
#endif

// The generic sampler interprets the poset program at run time (event times staged in shared memory); a
// poset-specialised sampler generated by jit.cuh has the same call signature with the poset unrolled into straight-line
// code (event times in registers, no parent loads / stores).
template <typename real, int PMAX, bool TIMES_GIVEN>
struct GenericCondSampler {
  static constexpr bool kStagesT = true;
  __device__ __forceinline__ real operator()(const PosetProg& pg, const uint32_t tcol_s, const real* __restrict__ s_lam2,
                                             const real* __restrict__ s_inv, const uint32_t l, const uint32_t gi,
                                             const uint32_t k0, const uint32_t k1, const uint32_t cw, const real ts_given,
                                             const uint32_t Xt, real (&Z)[PMAX], real& Ts_out) const {
    return sample_conditional<real, PMAX, TIMES_GIVEN>(pg, tcol_s, s_lam2, s_inv, l, gi, k0, k1, cw, ts_given, Xt, Z,
                                                       Ts_out);
  }
};

template <typename real, int PMAX, bool TIMES_GIVEN, int MODE, typename Sampler, typename Prog>
__device__ __forceinline__ void estep_conditional_body(const CondArgs& a, const Prog& pg, const Sampler sampler) {
  typedef typename Prog::word_t word_t;  // uint32_t (p <= 32) or uint64_t (specialised kernels only, p <= 64)
  static_assert(MODE != kCondAddRemove || sizeof(word_t) == 4, "add-remove maps proposals to lanes: p <= 32");
  __shared__ real s_T[Sampler::kStagesT ? (PMAX + 1) * kFwdBlock : 1];
  __shared__ real s_lam2[PMAX + 1];
  __shared__ real s_inv[PMAX + 1];
  __shared__ real s_lpyx2[68];  // log2 P(Y|X) by Hamming distance
  // add-remove scheme: transitive closures and "expected time to reach" scales (scale_path_to_mutation,
  // src/add_remove.cpp:17-37) by topological position; per warp the categorical over the <= p + 1 proposals of the
  // current observation (cumulative probabilities, hidden genotype and constant log2-weight part of every proposal)
  constexpr int kAR = MODE == kCondAddRemove ? PMAX : 1;
  __shared__ uint32_t s_anc[kAR], s_desc[kAR];
  __shared__ real s_sc[kAR];
  __shared__ real s_cdf[kFwdBlock / 32][kAR + 1];
  __shared__ real s_lc[kFwdBlock / 32][kAR + 1];
  __shared__ uint32_t s_Xc[kFwdBlock / 32][kAR + 1];

  const DevModel* __restrict__ mdl = a.model;
  const int p = pg.p;
  const int lane = threadIdx.x & 31;
  constexpr int kWarps = kFwdBlock / 32;
  if (Sampler::kStagesT) s_T[PMAX * kFwdBlock + threadIdx.x] = (real)0;  // dummy parent row
  const uint32_t tcol_s = (uint32_t)__cvta_generic_to_shared(s_T + (Sampler::kStagesT ? threadIdx.x : 0));
  const double kLog2e = 1.44269504088896340736;

  for (int t = threadIdx.x; t <= PMAX; t += kFwdBlock) {
    const double lam = t < p ? mdl->lambda[pg.ev[t]] : mdl->lambda_s;
    const bool live = t < p || t == PMAX;
    s_lam2[t] = live ? (real)(lam * kLog2e) : (real)0;
    s_inv[t] = live ? (real)(1.0 / (lam * kLog2e)) : (real)0;
  }
  for (int d = threadIdx.x; d < 68; d += kFwdBlock) {
    double lp = 0.0;
    if ((MODE == kCondBackward || MODE == kCondAddRemove) && d <= p) {
      // log_bernoulli_process (src/mcem_hcbn.cpp:84-100), incl. the eps == 0 special case
      const double eps = mdl->eps;
      if (eps == 0.0)
        lp = d ? log(eps + 2.2204460492503131e-16 /* DBL_EPSILON */) * d + log1p(-eps - 2.2204460492503131e-16 /* DBL_EPSILON */) * (p - d) : 0.0;
      else
        lp = log(eps) * d + log1p(-eps) * (p - d);
    }
    s_lpyx2[d] = (real)(lp * kLog2e);
  }
  if constexpr (MODE == kCondAddRemove) if (threadIdx.x == 0) {
    double sc[PMAX];
    for (int k = 0; k < p; ++k) {  // topological order: parents before children
      uint32_t an = 1u << k;
      double mx = -1.0;
      for (uint32_t m = pg.parent_mask[k]; m; m &= m - 1) {
        const int j = __ffs((int)m) - 1;
        an |= s_anc[j];
        mx = fmax(mx, sc[j]);
      }
      s_anc[k] = an;
      sc[k] = 1.0 / mdl->lambda[pg.ev[k]] + (mx > -1.0 ? mx : 0.0);
      s_sc[k] = (real)sc[k];
      s_desc[k] = 1u << k;
    }
    for (int k = p - 1; k >= 0; --k)
      for (uint32_t m = pg.parent_mask[k]; m; m &= m - 1) s_desc[__ffs((int)m) - 1] |= s_desc[k];
  }
  __syncthreads();

  const long long gw = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * kWarps;
  const uint32_t cw = a.cw;
  const uint32_t key1 = iter_key(a.key1, mdl);
  const uint32_t eps_thr = (MODE == kCondBernoulli) ? (uint32_t)fmin(mdl->eps * 4294967296.0, 4294967295.0) : 0u;

  (void)gw;
  (void)nw;
  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(a.next_item, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= a.n_items) break;
    const int i = (int)(item / a.chunks);
    const int ch = (int)(item - (long long)i * a.chunks);
    const uint64_t yw = a.obs_bits[i];
    const word_t yt = to_topo_bits<PMAX>(pg, yw);
    const real ts_given = TIMES_GIVEN ? (real)a.times[i] : (real)0;
    const uint32_t gi = (uint32_t)(a.obs_offset + i);

    real acc[PMAX];
#pragma unroll
    for (int k = 0; k < PMAX; ++k) acc[k] = (real)0;
    real aw = (real)0, aw2 = (real)0, awd = (real)0, npos = (real)0;
    real mref = (real)-1e30;  // this lane's log2 reference

    auto consume = [&](const uint32_t l, const word_t Xt, const int d, const bool valid, const real lw2_extra) {
      real Z[PMAX];
      real Ts;
      const real lw2s = sampler(pg, tcol_s, s_lam2, s_inv, l, gi, a.key0, key1, cw, ts_given, Xt, Z, Ts);
      const real lw2 = lw2s + lw2_extra;
      const bool ok = valid && (lw2 > (real)-1e30);  // F <= 0 (incompatible) gives NaN / -inf -> not ok
      if (ok && lw2 > mref) {  // rare after the first few samples: move the reference up (with head-room)
        const real nref = lw2 + (real)4;
        const real s = CondMath<real>::ex2(mref - nref);
        aw *= s;
        aw2 *= s * s;
        awd *= s;
#pragma unroll
        for (int k = 0; k < PMAX; ++k) acc[k] *= s;
        mref = nref;
      }
      const real wr = ok ? CondMath<real>::ex2(lw2 - mref) : (real)0;
      aw += wr;
      aw2 = fma(wr, wr, aw2);
      awd = fma(wr, (real)d, awd);
      npos += (wr > (real)0) ? (real)1 : (real)0;
#pragma unroll
      for (int k = 0; k < PMAX; ++k) acc[k] = fma(wr, Z[k], acc[k]);
    };

    double lscale_extra = 0.0;
    if (MODE == kCondBackward) {
      // number of compatible genotypes in the whole ball (:525: nrows_sum - #incompatible)
      int n_compat = 0;
      for (int g = lane; g < a.n_ball; g += 32)
        n_compat += compatible_topo<PMAX>(pg, to_topo_bits<PMAX>(pg, yw ^ a.ball[g]));
      n_compat = __reduce_add_sync(0xffffffffu, n_compat);
      lscale_extra = log((double)n_compat);
      const int g_begin = ch * a.chunk_len, g_end = min(a.n_ball, g_begin + a.chunk_len);
      for (int g = g_begin; g < g_end; ++g) {
        const uint64_t flip = a.ball[g];
        const word_t Xt = to_topo_bits<PMAX>(pg, yw ^ flip);
        if (!compatible_topo<PMAX>(pg, Xt)) continue;  // warp-uniform
        const int d = __popcll(flip);
        for (int r0 = 0; r0 < a.reps; r0 += 32) {
          const int rep = r0 + lane;
          // sample index as in the reference's replicate(reps, 1) layout: row l = g + n_ball * rep
          consume((uint32_t)(g + a.n_ball * rep), Xt, d, rep < a.reps, s_lpyx2[d]);
        }
      }
    } else if constexpr (MODE == kCondAddRemove) {
      // The reference draws (move, idx_add, idx_remove) independently and uses one of the two indices
      // (src/mcem_hcbn.cpp:401-437); the pair (move, used index) is ONE categorical variable over <= p + 1 proposals:
      //   add event k (y_k = 0):     prob P(add) a_k / sum a,  a_k = 1 / scale_k;   X = closure_up(Y + k)
      //   remove event k (y_k = 1):  prob P(rem) r_k / sum r,  r_k = scale_k;       X = closure_down(Y - k)
      //   stand still (Y compatible): X = Y
      // with closure_up / closure_down of the whole genotype when Y is incompatible (add_all / remove_all,
      // src/add_remove.cpp:39-56, :74-91) and of the touched event only when it is compatible (:58-72, :93-107).
      const int w = threadIdx.x >> 5;
      const bool live = lane < p;
      const bool yk = live && ((yt >> lane) & 1u);
      const bool compat = compatible_topo<PMAX>(pg, yt);
      const int mutations = __popc(yt);
      const uint32_t anc = live ? s_anc[lane] : 0u, desc = live ? s_desc[lane] : 0u;
      const uint32_t up = __reduce_or_sync(0xffffffffu, yk ? anc : 0u);                 // Y plus all its ancestors
      const uint32_t down = __ballot_sync(0xffffffffu, yk && (anc & ~yt) == 0u);        // largest compatible subset
      const real sck = live ? s_sc[lane] : (real)1;
      real wa = (live && !yk) ? (real)1 / sck : (real)0, wr = yk ? sck : (real)0;
      real sa = wa, sr = wr;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
      }
      real p_add, p_rem, p_stand;
      if (!compat) p_add = (real)0.5, p_rem = (real)0.5, p_stand = (real)0;
      else if (mutations == 0) p_add = (real)0.5, p_rem = (real)0, p_stand = (real)0.5;
      else if (mutations == p) p_add = (real)0, p_rem = (real)0.5, p_stand = (real)0.5;
      else p_add = p_rem = p_stand = (real)(1.0 / 3.0);
      const real pmf = !live ? (real)0 : (yk ? p_rem * wr / sr : p_add * wa / sa);
      real cdf = pmf;  // inclusive scan over lanes = proposals in topological-position order
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const real t = __shfl_up_sync(0xffffffffu, cdf, o);
        if (lane >= o) cdf += t;
      }
      const uint32_t Xk = yk ? ((compat ? yt : down) & ~desc) : ((compat ? yt : up) | anc);
      const uint32_t has = __ballot_sync(0xffffffffu, pmf > (real)0);
      const int cmax = p_stand > (real)0 ? p : 31 - __clz((int)has);  // last proposal that can be drawn
      __syncwarp();
      if (live) {
        s_cdf[w][lane] = cdf;
        s_Xc[w][lane] = Xk;
        s_lc[w][lane] = s_lpyx2[__popc(Xk ^ yt)] - CondMath<real>::lg2(fmax(pmf, (real)1e-37));
      }
      if (lane == 0) {
        s_cdf[w][p] = (real)2;
        s_Xc[w][p] = yt;
        s_lc[w][p] = s_lpyx2[0] - CondMath<real>::lg2(fmax(p_stand, (real)1e-37));
      }
      __syncwarp();
      const int l_begin = ch * a.chunk_len, l_end = min(a.L, l_begin + a.chunk_len);
      for (int l0 = l_begin; l0 < l_end; l0 += 32) {
        const int l = l0 + lane;
        const uint4 rc = philox4x32_10((uint32_t)l, gi, 64u, cw, a.key0, key1);
        const real u = cond_uniform<real>(rc, 0);
        int cidx = 0;
        for (int k = 0; k < p; ++k) cidx += (s_cdf[w][k] <= u) ? 1 : 0;  // broadcast reads
        cidx = min(cidx, cmax);
        const uint32_t Xt = s_Xc[w][cidx];
        consume((uint32_t)l, Xt, __popc(Xt ^ yt), l < l_end, s_lc[w][cidx]);
      }
      __syncwarp();
    } else {
      const int l_begin = ch * a.chunk_len, l_end = min(a.L, l_begin + a.chunk_len);
      for (int l0 = l_begin; l0 < l_end; l0 += 32) {
        const int l = l0 + lane;
        word_t Xt = yt;
        if constexpr (MODE == kCondBernoulli) {
          // coin_tossing: flip each bit independently with probability eps (32 random bits per coin)
          uint4 rc = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int k = 0; k < PMAX; ++k) {
            if (k < p) {
              if ((k & 3) == 0) rc = philox4x32_10((uint32_t)l, gi, 64u + (uint32_t)(k >> 2), cw, a.key0, key1);
              const uint32_t bits = (k & 3) == 0 ? rc.x : ((k & 3) == 1 ? rc.y : ((k & 3) == 2 ? rc.z : rc.w));
              Xt ^= (bits < eps_thr) ? ((word_t)1 << k) : (word_t)0;
            }
          }
        }
        const bool comp = compatible_topo<PMAX>(pg, Xt);
        consume((uint32_t)l, Xt, popc_word((word_t)(Xt ^ yt)), (l < l_end) && comp, s_lpyx2[popc_word((word_t)(Xt ^ yt))]);
      }
    }

    // align the lanes to the warp-wide reference, then reduce in fp64
    real mmax = mref;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
    const real s = (mref > (real)-1e29) ? CondMath<real>::ex2(mref - mmax) : (real)0;
    double* out = a.part + (size_t)item * a.PS;
    double v0 = (double)(aw * s), v1 = (double)(aw2 * s * s), v2 = (double)(awd * s), v3 = (double)npos;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      v3 += __shfl_xor_sync(0xffffffffu, v3, o);
    }
    if (lane == 0) {
      out[0] = v0;
      out[1] = v1;
      out[2] = v2;
      out[3] = v3;
      // natural-log scale of this record: true weights = exp(lscale) * relative weights
      out[4] = (mmax > (real)-1e29) ? (double)mmax * 0.693147180559945309417 + lscale_extra : -__longlong_as_double(0x7ff0000000000000LL) /* -inf */;
    }
    double mine, mine_hi = 0.0;  // lane k owns slots k and (p > 32) k + 32
    {
      double v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = k < PMAX ? (double)(acc[k] * s) : 0.0;
      mine = warp_transpose_sum(v, lane);
      if constexpr (PMAX > 32) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = k + 32 < PMAX ? (double)(acc[k + 32] * s) : 0.0;
        mine_hi = warp_transpose_sum(v, lane);
      }
    }
    if (lane < p) out[kStatHdr + lane] = mine;
    if (PMAX > 32 && lane + 32 < p) out[kStatHdr + 32 + lane] = mine_hi;
  }
}

template <typename real, int PMAX, bool TIMES_GIVEN, int MODE>
__global__ void __launch_bounds__(kFwdBlock, sizeof(real) == 4 ? MCCBN_COND_MINB : 1)
    estep_conditional_kernel(const CondArgs a, const __grid_constant__ PosetProg pg) {
  estep_conditional_body<real, PMAX, TIMES_GIVEN, MODE>(a, pg, GenericCondSampler<real, PMAX, TIMES_GIVEN>{});
}

}  // namespace mccbn
