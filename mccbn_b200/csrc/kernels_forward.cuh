// kernels_forward.cuh -- forward-scheme E-step (stages 2-4 of the hot path) for sm_100a.
//
// Replaces, per (observation i, sample l):
//   sample_genotypes     src/mcem_hcbn.cpp:138-177   Z_j ~ Exp(lambda_j), T_s ~ Exp(lambda_s),
//                                                    T_j = Z_j + max(0, max_{u in pa(j)} T_u), X_j = [T_j <= T_s]
//   hamming_dist_mat     src/mcem_hcbn.cpp:326-330   d = popc(X ^ Y_i)
//   forward weights      src/mcem_hcbn.cpp:386-388   w = eps^d (1-eps)^(p-d)
//   per-observation sums src/mcem_hcbn.cpp:683-694   sum w, sum w d, sum w Z_j   (+ sum w^2 for :994)
//
// Mapping: one WARP per (observation, chunk of samples); lane = sample.  The node loop is fully unrolled over
// topological positions, so the p waiting times Z_k of the current sample and the p+3 running sums live in
// registers with static indices.  Event times T_k that have children are staged in a per-thread shared-memory
// column (a runtime poset cannot index registers); the poset itself is a __grid_constant__ parameter, i.e. the
// two parent offsets of node k are constant-bank operands and every node costs exactly two LDS + one FMNMX
// (absent parents read a row that holds 0); nodes with more than two parents take a warp-uniform rare path.
// Weights are tracked relative to the best (smallest-distance) sample seen so far, r^(d - dref) in fp32 with
// r = eps/(1-eps) <= 1 -- a running log-sum-exp -- and the per-lane fp32 partial sums over <= L/32 samples are
// reduced across the warp in fp64.
// Random bits (fp32 path): the p + 1 uniforms of a sample are 24-bit fields packed back to back in the Philox
// output stream (field 0 = sampling time, field k + 1 = node k; 23 bits of each are used as the mantissa), so a
// sample of p = 30 events costs 6 Philox4x32-10 blocks instead of 8.  What is staged and accumulated per node is
// e_k = log2(u_k) <= 0; the scale c_k = -ln2/lambda_k enters the event time through one FFMA
// (T_k = e_k c_k + max parents) and the sufficient statistic once per record (sum_l w_l Z_lk = c_k sum_l w_l e_lk).
#pragma once
#include "device_types.cuh"
#include "philox.cuh"

namespace mccbn {

// lambda-derived scales as immediate constant-bank operands: used ONLY by the NVRTC-specialised kernels (jit.cuh), whose
// module -- and therefore whose copy of this symbol -- belongs to one context.  The kernels compiled into the library
// read the scales of their candidate from FwdArgs::scale (device memory -> shared memory): contexts that share a device
// share the library's module, so a module-wide symbol would be raced by their streams.
__constant__ FwdScale c_scale;

struct FwdArgs {
  const uint64_t* obs_bits;  // N
  const double* times;       // N (TIMES_GIVEN only)
  const DevModel* model;     // this candidate's parameters (weight tables; fp64 mode: lambda)
  const FwdScale* scale;     // this candidate's exponential scales (fp32 path)
  double* part;              // n_items x PS
  int N;
  int L;
  int chunks;                // chunks per observation
  int chunk_len;             // samples per chunk (multiple of 32)
  long long n_items;         // N * chunks
  long long obs_offset;      // global index of local observation 0
  uint32_t key0, key1;       // Philox key (seed, iteration tag); the iteration field of key1 is taken from
                             // model->iter on the device (iter_key), so that a captured CUDA graph of one EM iteration
                             // can be replayed
  uint32_t cw;               // Philox counter word 3 (evaluation / candidate id)
  int PS;                    // record stride = kStatHdr + p
  unsigned long long* next_item;  // optional work counter (zeroed before the launch): warps claim items dynamically
};

// key word 1 = (EM iteration << 4) | purpose (make_key, philox.cuh); the iteration is read from the model on the device
__device__ __forceinline__ uint32_t iter_key(const uint32_t key1, const DevModel* __restrict__ mdl) {
  return ((uint32_t)mdl->iter << 4) | (key1 & 0xFu);
}

template <typename real>
struct RealTraits;
template <>
struct RealTraits<float> {
  static constexpr int kUniPerBlock = 4;  // conditional kernels: one 32-bit word per uniform
};
template <>
struct RealTraits<double> {
  static constexpr int kUniPerBlock = 2;
  // e = ln(u) <= 0;  Z = e * c with c = -1/lambda
  static __device__ __forceinline__ double log_u(const uint4& r, int w) {
    return log(w == 0 ? u01_open0_d(r.x, r.y) : u01_open0_d(r.z, r.w));
  }
};

// ---- fp32 path: 24-bit uniform fields packed in the Philox output stream ----------------------------------------
// Field f occupies stream bits [24 f, 24 f + 24); word w of block b is stream word 4 b + w.  23 of the 24 bits become
// the mantissa of a float in [1, 2); u = 2 - that lies in (0, 1].  `f` is a compile-time constant wherever this is
// called (unrolled loops / generated code), so every case below folds to one or two instructions.
__host__ __device__ constexpr int fwd_field_last_word(int f) { return (24 * f + 23) / 32; }
__host__ __device__ constexpr int fwd_field_block(int f) { return fwd_field_last_word(f) / 4; }
// Philox blocks needed for fields 0 .. nf-1
__host__ __device__ constexpr int fwd_blocks_for_fields(int nf) { return nf <= 0 ? 0 : fwd_field_block(nf - 1) + 1; }

// `one` = 0x3f800000 held in a register (PosetProg::one_bits): with an immediate the compiler needs two LOP3 for
// (x & mask) | one, with a register operand it is a single three-input LOP3.
// field -> float in [1, 2) with the 23 field bits as mantissa
template <int NW>
__device__ __forceinline__ float field_f12(const uint32_t (&W)[NW], const int f, const uint32_t one) {
  const int s = 24 * f, wi = s >> 5, o = s & 31;
  uint32_t m;
  if (o == 0) m = (W[wi] & 0x7fffffu) | one;                                          // stream bits [s, s+23)
  else if (o == 8) m = (W[wi] >> 9) | 0x3f800000u;                                    // [s+1, s+24): one LEA.HI
  else m = (__funnelshift_r(W[wi], W[wi + 1 < NW ? wi + 1 : wi], o) & 0x7fffffu) | one;  // o = 16, 24: [s, s+23)
  return __uint_as_float(m);
}
// u in (0, 1] (forward draws: log(u)); u in [0, 1) (conditional draws: inverse CDF of the truncated exponential)
template <int NW>
__device__ __forceinline__ float field_u01(const uint32_t (&W)[NW], const int f, const uint32_t one) {
  return 2.0f - field_f12(W, f, one);
}
template <int NW>
__device__ __forceinline__ float field_u01_open1(const uint32_t (&W)[NW], const int f, const uint32_t one) {
  return field_f12(W, f, one) - 1.0f;
}

// Z staging: each thread owns a row of PMAX values (e_k) stored as 16-byte vectors; an odd number of vectors per
// row makes the 128-bit accesses of a warp bank-conflict free.
template <typename real, int PMAX>
struct ZStage {
  static constexpr int kVec = 16 / (int)sizeof(real);
  static constexpr int kStrideVec = ((PMAX + kVec - 1) / kVec) | 1;
  static constexpr int kStride = kStrideVec * kVec;  // elements per thread
  static constexpr uint32_t kRowBytes = (uint32_t)(kStride * sizeof(real));
};

// per-node exponential scale, by topological position; element PMAX = the sampling time's scale.  A shared-memory array
// filled at kernel start (fp32: from FwdArgs::scale; fp64 validation mode: -1/lambda from the model).
template <typename real, int PMAX>
struct ScaleSrc {
  const real* s;  // [PMAX + 1]
  __device__ __forceinline__ real operator()(int k) const { return s[k < 0 ? PMAX : k]; }
};

// shared-memory load through a 32-bit shared-space address (keeps LDS even where the compiler loses track of the
// address space)
template <typename real>
__device__ __forceinline__ real lds_at(uint32_t saddr);
template <>
__device__ __forceinline__ float lds_at<float>(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
template <>
__device__ __forceinline__ double lds_at<double>(uint32_t saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_at_f(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v)); }
__device__ __forceinline__ void sts_at_d(uint32_t saddr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(saddr), "d"(v)); }
template <typename real>
__device__ __forceinline__ void sts_at(uint32_t saddr, real v) {
  if (sizeof(real) == 4) sts_at_f(saddr, (float)v);
  else sts_at_d(saddr, (double)v);
}
__device__ __forceinline__ void sts128_f(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ float4 lds128_f(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// Warp reduction of a vector of per-lane partial sums: v[j] of every lane is summed over the 32 lanes, and lane l returns
// the total of component l.  A butterfly transpose: at distance o = 16, 8, .. 1 a lane keeps the half of its components
// that its bit o selects, sends the other half to lane ^ o and adds what it receives -- 31 exchanges for 32 components,
// where 32 independent xor-reductions take 160 (an fp64 exchange is two SHFL: 62 instead of 320 per record; the record
// epilogue is per (observation, chunk) item, which for short chunks -- L = 1000 gives 11 sample iterations per item --
// was a quarter of the kernel).  Fixed exchange order: the result is deterministic and the same on every lane layout.
__device__ __forceinline__ double warp_transpose_sum(double (&v)[32], const int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const double lo = v[i], hi = v[i + o];
      const double send = up ? lo : hi;
      const double keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// max over the parents of node k (0 without parents): two unconditional loads (absent parents read the dummy
// row, which holds 0); nodes with in-degree > 2 (warp-uniform test) read four more rows unconditionally, and only
// in-degree > 6 falls back to a loop.
__device__ __forceinline__ int popc_word(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int popc_word(uint64_t x) { return __popcll(x); }

// LEAN = 3: the poset has no node with more than three parents (the host checks; true for most candidate posets of an
// annealing run): three unconditional loads (absent parents read the dummy row), one FMNMX3, and none of the per-node
// in-degree tests.  LEAN = 0 is the general path (any in-degree).  (Six unconditional loads, LEAN = 6, were tried for
// in-degree <= 6: 227 LDS per sample at PMAX = 32 and register spills -- slower than the general path, kept out.)
template <typename real, typename Prog, int LEAN = 0>
__device__ __forceinline__ real parents_max(const Prog& pg, const typename Prog::word_t multi_mask, const int k,
                                            const uint32_t tcol_s) {
  const uint2 o = pg.off[k];
  if constexpr (LEAN == 3)
    return max(max(lds_at<real>(tcol_s + o.x), lds_at<real>(tcol_s + o.y)), lds_at<real>(tcol_s + pg.xo[k].x));
  real m = max(lds_at<real>(tcol_s + o.x), lds_at<real>(tcol_s + o.y));
  if ((multi_mask >> k) & 1u) {
    const uint4 x = pg.xo[k];
    m = max(m, max(max(lds_at<real>(tcol_s + x.x), lds_at<real>(tcol_s + x.y)),
                   max(lds_at<real>(tcol_s + x.z), lds_at<real>(tcol_s + x.w))));
    if ((pg.huge_mask >> k) & 1u) {
      const int e1 = pg.xstart[k + 1];
#pragma unroll 1
      for (int e = pg.xstart[k]; e < e1; ++e)
        m = max(m, lds_at<real>(tcol_s + (uint32_t)pg.xoff[e] * (sizeof(real) / 4)));
    }
  }
  return m;
}

// One forward draw for sample l of (global) observation gi.  e_k = log(u_k) (log2 in fp32, ln in fp64) is staged in this
// thread's row at zrow_s (ZStage layout).  All PMAX slots run unconditionally (no per-node branch): slots k >= p have
// scale 0 and no parents; the final shift / valid_mask removes them from X.
//
// fp32 path -- event times RELATIVE to the sampling time: T'_k = T_k - T_s = e_k c_k + max(-T_s, max_parents T'), staged
// in this thread's shared-memory column (element k at tcol_s + k * kFwdBlock * 4; the dummy row PMAX, which absent
// parents read, holds -T_s of the current sample).  The hidden genotype bit X_k = [T_k <= T_s] is then the SIGN BIT of
// T'_k, shifted into X with one funnel shift per node (no compare, no predicated add): the node at topological
// position k ends up at bit p - 1 - k of X (to_topo_bits_rev gives the observed genotype in the same order).
// T'_k == 0 exactly (an fp32 tie, probability ~1e-8 per node) counts as "not yet mutated"; the specialised sampler of
// jit.cuh performs the identical arithmetic, so both kernels produce the same X bit for bit.
// fp64 path (validation mode) -- absolute times and the reference's comparison T_k <= T_s, X_k at bit p - 1 - k too.
// Random stream of a sample (fp32): field 0 = sampling time, field k + 1 = node k (field_u01); Philox block b is
// generated right before the first field that touches it, and only if that field belongs to a real node.
template <typename real, int PMAX, bool TIMES_GIVEN, typename Prog, int LEAN = 0>
__device__ __forceinline__ void sample_forward(const Prog& pg, const uint32_t tcol_s, const uint32_t zrow_s,
                                               const uint32_t l, const uint32_t gi, const uint32_t k0,
                                               const uint32_t k1, const uint32_t cw, const real ts_given,
                                               const ScaleSrc<real, PMAX> scale, typename Prog::word_t& X,
                                               real& Ts_out) {
  typedef typename Prog::word_t word_t;
  constexpr uint32_t kRow = kFwdBlock * sizeof(real);
  const word_t multi_mask = pg.multi_mask;
  real Ts = ts_given;
  word_t x = 0;
  if constexpr (sizeof(real) == 4) {
    constexpr int NB = fwd_blocks_for_fields(PMAX + 1);
    uint32_t W[4 * NB];
    float zq[4];
    const uint32_t one = pg.one_bits;
#pragma unroll
    for (int k = 0; k < PMAX; ++k) {
      const int b = fwd_field_block(k + 1);
      if (k == 0 || b != fwd_field_block(k)) {  // compile-time after unrolling
        // block b is first touched by node k; for k == 0 this is block 0, which also holds field 0
        if (k == 0 || k < pg.p) {               // warp-uniform
          const uint4 r = philox4x32_10(l, gi, (uint32_t)b, cw, k0, k1);
          W[4 * b + 0] = r.x, W[4 * b + 1] = r.y, W[4 * b + 2] = r.z, W[4 * b + 3] = r.w;
        } else {
          W[4 * b + 0] = W[4 * b + 1] = W[4 * b + 2] = W[4 * b + 3] = 0u;
        }
        if (k == 0) {
          if (!TIMES_GIVEN) Ts = fast_lg2(field_u01(W, 0, one)) * scale(-1);
          sts_at_f(tcol_s + PMAX * kRow, -Ts);  // dummy parent row: max(-T_s, ...) is the reference's T_max = 0
        }
      }
      const float e = fast_lg2(field_u01(W, k + 1, one));  // log2(u) <= 0;  Z = e * c,  c = -ln2/lambda
      const float t = fmaf(e, scale(k), parents_max<real, Prog, LEAN>(pg, multi_mask, k, tcol_s));  // T_k - T_s
      sts_at_f(tcol_s + k * kRow, t);
      zq[k & 3] = e;
      if ((k & 3) == 3) sts128_f(zrow_s + (uint32_t)(k - 3) * 4u, zq[0], zq[1], zq[2], zq[3]);
      if constexpr (sizeof(word_t) == 4) x = __funnelshift_l(__float_as_uint(t), x, 1);
      else x = (x << 1) | (word_t)(__float_as_uint(t) >> 31);
    }
  } else {
    constexpr int UPB = RealTraits<double>::kUniPerBlock;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (!TIMES_GIVEN) {
      r = philox4x32_10(l, gi, 255u, cw, k0, k1);
      Ts = RealTraits<double>::log_u(r, 0) * scale(-1);
    }
#pragma unroll
    for (int k = 0; k < PMAX; ++k) {
      if (k % UPB == 0) r = philox4x32_10(l, gi, (uint32_t)(k / UPB), cw, k0, k1);
      const double e = RealTraits<double>::log_u(r, k % UPB);
      const double t = fma(e, scale(k), parents_max<real>(pg, multi_mask, k, tcol_s));
      sts_at_d(tcol_s + k * kRow, t);
      sts_at_d(zrow_s + (uint32_t)k * 8u, e);
      x = (x << 1) | (word_t)(t <= Ts ? 1u : 0u);
    }
  }
  X = (x >> (PMAX - pg.p)) & pg.valid_mask;  // drop the padding slots: node k -> bit p - 1 - k
  Ts_out = Ts;
}

// observed genotype word (event-id bit order) -> topological-position bit order
template <int PMAX, typename Prog>
__device__ __forceinline__ typename Prog::word_t to_topo_bits(const Prog& pg, const uint64_t yw) {
  typedef typename Prog::word_t word_t;
  word_t yt = 0;
  const word_t y = (word_t)yw;  // p <= 32: 32-bit shifts
#pragma unroll
  for (int k = 0; k < PMAX; ++k)
    if (k < pg.p) yt |= ((y >> pg.ev[k]) & (word_t)1) << k;
  return yt;
}

// the same in the bit order of the forward samplers' X: node k (topological position) -> bit p - 1 - k
template <int PMAX, typename Prog>
__device__ __forceinline__ typename Prog::word_t to_topo_bits_rev(const Prog& pg, const uint64_t yw) {
  typedef typename Prog::word_t word_t;
  word_t yt = 0;
  const word_t y = (word_t)yw;  // p <= 32: events 0..31 live in the low word -- 32-bit shifts
#pragma unroll
  for (int k = 0; k < PMAX; ++k)
    if (k < pg.p) yt |= ((y >> pg.ev[k]) & (word_t)1) << (pg.p - 1 - k);
  return yt;
}

// fill the block's scale array (ScaleSrc): fp32 from the candidate's FwdScale, fp64 (validation mode) from the model
template <typename real, int PMAX, typename Prog>
__device__ __forceinline__ void load_scales(const Prog& pg, const DevModel* __restrict__ mdl,
                                            const FwdScale* __restrict__ sc, real* s_scale) {
  for (int t = threadIdx.x; t <= PMAX; t += kFwdBlock) {
    if (sizeof(real) == 8)
      s_scale[t] = t < pg.p ? (real)(-1.0 / mdl->lambda[pg.ev[t]]) : (t == PMAX ? (real)(-1.0 / mdl->lambda_s) : (real)0);
    else
      s_scale[t] = t < pg.p ? (real)sc->c[t] : (t == PMAX ? (real)sc->cs : (real)0);
  }
}

#ifndef MCCBN_FWD_BIGP
#define MCCBN_FWD_BIGP 28       // PMAX >= this: 5 blocks/SM (102 registers), below: 6 blocks/SM (85 registers)
#endif
#ifndef MCCBN_FWD_MINB_BIG
#define MCCBN_FWD_MINB_BIG 5
#endif
template <typename real, int PMAX>
__host__ __device__ constexpr size_t fwd_smem_bytes(bool stages_t = true) {
  return (stages_t ? sizeof(real) * (size_t)(PMAX + 1) * kFwdBlock : 0) +
         (size_t)ZStage<real, PMAX>::kRowBytes * kFwdBlock;
}

// The generic sampler interprets the poset program at run time (event times staged in shared memory).  A
// poset-specialised sampler generated by jit.cuh has the same call signature with the poset unrolled into
// straight-line code (event times in registers).
template <typename real, int PMAX, bool TIMES_GIVEN, int LEAN = 0>
struct GenericSampler {
  static constexpr bool kStagesT = true;
  static constexpr bool kStagesE = true;  // e_k goes through the thread's shared-memory row; E is left untouched
  template <typename Prog>
  __device__ __forceinline__ void operator()(const Prog& pg, const uint32_t tcol_s, const uint32_t zrow_s,
                                             const uint32_t l, const uint32_t gi, const uint32_t k0,
                                             const uint32_t k1, const uint32_t cw, const real ts_given,
                                             const ScaleSrc<real, PMAX> scale, typename Prog::word_t& X,
                                             real& Ts_out, real (&E)[PMAX]) const {
    (void)E;
    sample_forward<real, PMAX, TIMES_GIVEN, Prog, LEAN>(pg, tcol_s, zrow_s, l, gi, k0, k1, cw, ts_given, scale, X, Ts_out);
  }
};

template <typename real, int PMAX, bool TIMES_GIVEN, typename Sampler, typename Prog>
__device__ __forceinline__ void estep_forward_body(const FwdArgs& a, const Prog& pg, const Sampler sampler) {
  typedef typename Prog::word_t word_t;
  // dynamic shared memory (fwd_smem_bytes): [rows 0..PMAX-1 staged T_k, row PMAX the dummy row,] then -- samplers that
  // stage e_k = log(u_k) -- one row per thread (ZStage; frees PMAX registers per thread)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using ZS = ZStage<real, PMAX>;
  real* s_T = reinterpret_cast<real*>(smem_raw);
  real* s_Z = s_T + (Sampler::kStagesT ? (PMAX + 1) * kFwdBlock : 0);
  __shared__ real s_rtab[kRtabN];
  __shared__ real s_scale[PMAX + 1];

  const DevModel* __restrict__ mdl = a.model;
  const int p = pg.p;
  const int lane = threadIdx.x & 31;
  constexpr int kWarps = kFwdBlock / 32;
  if (Sampler::kStagesT) s_T[PMAX * kFwdBlock + threadIdx.x] = (real)0;
  if (Sampler::kStagesE)
    for (int k = 0; k < ZS::kStride; ++k) s_Z[threadIdx.x * ZS::kStride + k] = (real)0;  // unused node slots stay 0
  const uint32_t tcol_s = (uint32_t)__cvta_generic_to_shared(s_T + threadIdx.x);
  const uint32_t zrow_s = (uint32_t)__cvta_generic_to_shared(s_Z + threadIdx.x * ZS::kStride);

  for (int t = threadIdx.x; t < kRtabN; t += kFwdBlock)
    s_rtab[t] = sizeof(real) == 4 ? (real)mdl->rtab[t] : (real)mdl->rtab_d[t];
  load_scales<real, PMAX>(pg, mdl, a.scale, s_scale);
  __syncthreads();

  const ScaleSrc<real, PMAX> scale{reinterpret_cast<const real*>(s_scale)};

  const long long gw = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * kWarps;
  const int flip = mdl->flip;
  const uint32_t key1 = iter_key(a.key1, mdl);

  for (long long item = gw;; item += nw) {
    if (a.next_item) {  // warp-uniform
      if (lane == 0) item = (long long)atomicAdd(a.next_item, 1ull);
      item = __shfl_sync(0xffffffffu, item, 0);
    }
    if (item >= a.n_items) break;
    const int i = (int)(item / a.chunks);
    const int ch = (int)(item - (long long)i * a.chunks);
    // observed genotype in the samplers' bit order.  Weights decrease in de = flip ? p - d : d (flip: eps > 1/2), and
    // p - d = popc(X ^ ~Y): complementing Y once per item makes popc return de directly.
    const word_t yt = to_topo_bits_rev<PMAX>(pg, a.obs_bits[i]);
    const word_t ye = flip ? (word_t)(~yt & pg.valid_mask) : yt;
    const real ts_given = TIMES_GIVEN ? (real)a.times[i] : (real)0;
    const uint32_t gi = (uint32_t)(a.obs_offset + i);

    real acc[PMAX];
#pragma unroll
    for (int k = 0; k < PMAX; ++k) acc[k] = (real)0;
    real aw = (real)0, aw2 = (real)0, awd = (real)0;  // awd = sum w' de
    int dref = 1 << 20;

    const int l_begin = ch * a.chunk_len;
    const int l_end = min(a.L, l_begin + a.chunk_len);
    for (int l0 = l_begin; l0 < l_end; l0 += 32) {
      const int l = l0 + lane;
      word_t X;
      real Ts;
      real E[PMAX];
      sampler(pg, tcol_s, zrow_s, (uint32_t)l, gi, a.key0, key1, a.cw, ts_given, scale, X, Ts, E);
      // lanes beyond the chunk get a distance that maps to the zero tail of the weight table
      const int de = l < l_end ? popc_word((word_t)(X ^ ye)) : (1 << 19);
      const int dmin = __reduce_min_sync(0xffffffffu, de);
      if (dmin < dref) {  // warp-uniform: a better sample arrived, rescale the running sums
        const real s = s_rtab[min(dref - dmin, kRtabN - 1)];
        aw *= s;
        aw2 *= s * s;
        awd *= s;
#pragma unroll
        for (int k = 0; k < PMAX; ++k) acc[k] *= s;
        dref = dmin;
      }
      const real wr = s_rtab[min(de - dref, kRtabN - 1)];
      aw += wr;
      aw2 = fma(wr, wr, aw2);
      awd = fma(wr, (real)de, awd);  // lanes beyond the chunk: wr == 0
      if constexpr (!Sampler::kStagesE) {
#pragma unroll
        for (int k = 0; k < PMAX; ++k) acc[k] = fma(wr, E[k], acc[k]);
      } else if constexpr (sizeof(real) == 4) {
#pragma unroll
        for (int q = 0; q < PMAX / 4; ++q) {
          const float4 e = lds128_f(zrow_s + (uint32_t)q * 16u);
          acc[4 * q + 0] = fmaf(wr, e.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(wr, e.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(wr, e.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(wr, e.w, acc[4 * q + 3]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < PMAX; ++k) acc[k] = fma(wr, lds_at<real>(zrow_s + (uint32_t)k * 8u), acc[k]);
      }
    }

    // warp reduction in fp64 (all lanes share dref), lane k writes slot kStatHdr + k
    double* out = a.part + (size_t)item * a.PS;
    double v0 = (double)aw, v1 = (double)aw2, v2 = (double)awd;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if (lane == 0) {
      out[0] = v0;
      out[1] = v1;
      out[2] = flip ? (double)p * v0 - v2 : v2;  // sum w' d from sum w' de
      out[3] = (double)(l_end - l_begin);  // every forward sample has a positive weight (unless eps == 0)
      // natural-log scale of the record: true w = exp(scale) * w', scale = log(base) + dref * log(r)
      out[4] = mdl->log_base + (dref > 0 ? (double)dref * mdl->log_r : 0.0);
    }
    double mine, mine_hi = 0.0;  // lane k owns slots k and (p > 32) k + 32
    {
      double v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = k < PMAX ? (double)acc[k] : 0.0;
      mine = warp_transpose_sum(v, lane);
      if constexpr (PMAX > 32) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = k + 32 < PMAX ? (double)acc[k + 32] : 0.0;
        mine_hi = warp_transpose_sum(v, lane);
      }
    }
    // acc holds sum w' e_k: the waiting time is Z_k = e_k c_k (c_k = -ln2/lambda_k in fp32, -1/lambda_k in fp64)
    if (lane < p) out[kStatHdr + lane] = mine * (double)scale(lane);
    if (PMAX > 32 && lane + 32 < p) out[kStatHdr + 32 + lane] = mine_hi * (double)scale(lane + 32);
  }
}

template <typename real, int PMAX, bool TIMES_GIVEN, int LEAN = 0>
__global__ void __launch_bounds__(kFwdBlock, sizeof(real) == 4 ? (PMAX >= MCCBN_FWD_BIGP ? MCCBN_FWD_MINB_BIG : 6) : 1)
    estep_forward_kernel(const FwdArgs a, const __grid_constant__ PosetProg pg) {
  estep_forward_body<real, PMAX, TIMES_GIVEN>(a, pg, GenericSampler<real, PMAX, TIMES_GIVEN, LEAN>{});
}

// 32 < p <= 64: the same body on 64-bit genotype words (PosetProgWide); two resident blocks per SM
constexpr int kFwdWideMinB = 2;
template <int PMAX, bool TIMES_GIVEN>
__global__ void __launch_bounds__(kFwdBlock, kFwdWideMinB)
    estep_forward_wide_kernel(const FwdArgs a, const __grid_constant__ PosetProgWide pg) {
  estep_forward_body<float, PMAX, TIMES_GIVEN>(a, pg, GenericSampler<float, PMAX, TIMES_GIVEN>{});
}

// Per-sample dump for ONE observation (the `_importance_weight_genotype` seam, src/mcem_hcbn.cpp:866-912):
// same Philox stream as the production kernel, writes raw Z (L x p, event-id columns), dist, X, T_s.
struct FwdDumpArgs {
  uint64_t obs_word;
  double time;
  const DevModel* model;
  const FwdScale* scale;
  double* Z_out;    // L x p column-major by EVENT id
  double* T_out;    // L x p event times (sample_times, src/mcem_hcbn.cpp:184-213), or nullptr
  int* dist_out;    // L
  uint32_t* X_out;  // L, event-id bit order (p <= 32)
  double* Ts_out;   // L
  int L;
  uint32_t gi;
  uint32_t key0, key1, cw;
};

template <typename real, int PMAX, bool TIMES_GIVEN>
__global__ void __launch_bounds__(kFwdBlock) forward_dump_kernel(const FwdDumpArgs a,
                                                                 const __grid_constant__ PosetProg pg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* s_T = reinterpret_cast<real*>(smem_raw);
  using ZS = ZStage<real, PMAX>;
  real* s_Z = s_T + (PMAX + 1) * kFwdBlock;
  __shared__ real s_scale[PMAX + 1];
  const int p = pg.p;
  s_T[PMAX * kFwdBlock + threadIdx.x] = (real)0;
  const uint32_t tcol_s = (uint32_t)__cvta_generic_to_shared(s_T + threadIdx.x);
  const uint32_t zrow_s = (uint32_t)__cvta_generic_to_shared(s_Z + threadIdx.x * ZS::kStride);
  load_scales<real, PMAX>(pg, a.model, a.scale, s_scale);
  __syncthreads();
  const ScaleSrc<real, PMAX> scale{reinterpret_cast<const real*>(s_scale)};
  const uint32_t yt = to_topo_bits_rev<PMAX>(pg, a.obs_word);
  for (int l = blockIdx.x * kFwdBlock + threadIdx.x; l < a.L; l += gridDim.x * kFwdBlock) {
    uint32_t X;
    real Ts;
    sample_forward<real, PMAX, TIMES_GIVEN>(pg, tcol_s, zrow_s, (uint32_t)l, a.gi, a.key0, a.key1, a.cw, (real)a.time,
                                            scale, X, Ts);
    a.dist_out[l] = __popc(X ^ yt);
    uint32_t xe = 0u;
#pragma unroll
    for (int k = 0; k < PMAX; ++k)
      if (k < p) {
        a.Z_out[(size_t)pg.ev[k] * a.L + l] = (double)s_Z[threadIdx.x * ZS::kStride + k] * (double)scale(k);
        // the fp32 sampler stages T_k - T_s
        if (a.T_out)
          a.T_out[(size_t)pg.ev[k] * a.L + l] =
              (double)s_T[k * kFwdBlock + threadIdx.x] + (sizeof(real) == 4 ? (double)Ts : 0.0);
        xe |= ((X >> (p - 1 - k)) & 1u) << pg.ev[k];
      }
    a.X_out[l] = xe;
    a.Ts_out[l] = (double)Ts;
  }
}

#ifndef __CUDACC_RTC__  // not part of the NVRTC-specialised modules
// generate_genotypes (src/mcem_hcbn.cpp:215-236): X_kj = [T_sum(k, j) <= T_s(k)]; T_s ~ Exp(lambda_s) unless given.
// One thread per row; T_sum is N x p column-major by event id, the output an N x p int matrix.
__global__ void generate_genotypes_kernel(const double* __restrict__ T_sum, double* __restrict__ T_sampling,
                                          int* __restrict__ samples, int N, int p, double lambda_s, int times_given,
                                          uint32_t key0, uint32_t key1) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    double ts;
    if (times_given) {
      ts = T_sampling[i];
    } else {
      const uint4 r = philox4x32_10((uint32_t)i, 0u, 0u, 0u, key0, key1);
      ts = -log(u01_open0_d(r.x, r.y)) / lambda_s;
      T_sampling[i] = ts;
    }
    for (int j = 0; j < p; ++j) samples[(size_t)j * N + i] = T_sum[(size_t)j * N + i] <= ts ? 1 : 0;
  }
}

#endif  // __CUDACC_RTC__

}  // namespace mccbn
