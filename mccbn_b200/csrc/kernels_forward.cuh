// kernels_forward.cuh -- forward-scheme E-step (stages 2-4 of the hot path) for sm_100a.
//
// Replaces, per (observation i, sample l):
//   sample_genotypes     src/mcem_hcbn.cpp:138-177   Z_j ~ Exp(lambda_j), T_s ~ Exp(lambda_s),
//                                                    T_j = Z_j + max(0, max_{u in pa(j)} T_u), X_j = [T_j <= T_s]
//   hamming_dist_mat     src/mcem_hcbn.cpp:326-330   d = popc(X ^ Y_i)
//   forward weights      src/mcem_hcbn.cpp:386-388   w = eps^d (1-eps)^(p-d)
//   per-observation sums src/mcem_hcbn.cpp:683-694   sum w, sum w d, sum w Z_j   (+ sum w^2 for :994)
//
// Mapping: one WARP per (observation, chunk of samples); lane = sample; the p waiting times Z_k of the current
// sample live in registers (static indices: the node loop is unrolled over topological positions), the event
// times T_k that have children are staged in a per-thread shared-memory column (runtime posets cannot index
// registers), parent lists come from __constant__ memory (warp-uniform -> ULDC).  Weights are tracked relative to
// the best (smallest-distance) sample seen so far, r^(d - dref) in fp32 with r = eps/(1-eps) <= 1, a running
// log-sum-exp: per-lane fp32 partial sums over <= L/32 samples are then reduced across the warp in fp64.
#pragma once
#include "device_types.cuh"
#include "philox.cuh"

namespace mccbn {

__constant__ FwdConst c_fwd[kMaxCand];

struct FwdArgs {
  const uint64_t* obs_bits;  // N
  const double* times;       // N (TIMES_GIVEN only)
  const DevModel* models;    // n_cand
  double* part;              // n_cand x n_items x PS
  int N;
  int L;
  int chunks;                // chunks per observation
  int chunk_len;             // samples per chunk (multiple of 32 except the last)
  long long n_items;         // N * chunks
  long long obs_offset;      // global index of local observation 0
  uint32_t key0, key1;       // Philox key (seed, iteration tag)
  uint32_t cand_ctr[kMaxCand];  // Philox counter word 3 per candidate (evaluation id)
  int PS;                    // stats stride = kStatHdr + p
};

template <typename real>
struct RealTraits;
template <>
struct RealTraits<float> {
  static constexpr int kUniPerBlock = 4;
  static __device__ __forceinline__ float expo(const uint4& r, int w, float c) {
    const uint32_t bits = w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
    return fast_lg2(u01_open0(bits)) * c;  // c = -ln2/lambda  =>  -ln(u)/lambda
  }
};
template <>
struct RealTraits<double> {
  static constexpr int kUniPerBlock = 2;
  static __device__ __forceinline__ double expo(const uint4& r, int w, double c) {
    const double u = w == 0 ? u01_open0_d(r.x, r.y) : u01_open0_d(r.z, r.w);
    return log(u) * c;  // c = -1/lambda
  }
};

// One forward draw for sample l of (global) observation gi.  Z[k] = waiting time of the event at topological
// position k, X = hidden genotype in topological-position space.  `Tcol` is this thread's staging column
// (element k at Tcol[k * kFwdBlock]); cexp[k] is the per-node exponential scale (register/constant operand).
template <typename real, int PMAX, bool TIMES_GIVEN, typename ScaleFn>
__device__ __forceinline__ void sample_forward(const FwdConst& fc, const int p, real* __restrict__ Tcol,
                                               const uint32_t l, const uint32_t gi, const uint32_t k0,
                                               const uint32_t k1, const uint32_t cw, const real ts_given,
                                               ScaleFn scale, real (&Z)[PMAX], uint32_t& X, real& Ts_out) {
  constexpr int UPB = RealTraits<real>::kUniPerBlock;
  uint4 r = make_uint4(0, 0, 0, 0);
  real Ts = ts_given;
  if (!TIMES_GIVEN) {
    r = philox4x32_10(l, gi, 0u, cw, k0, k1);
    Ts = RealTraits<real>::expo(r, 0, scale(-1));
  }
  X = 0u;
#pragma unroll
  for (int k = 0; k < PMAX; ++k) {
    if (k < p) {
      const int ui = k + (TIMES_GIVEN ? 0 : 1);  // index of this node's uniform in the sample's stream
      if (ui % UPB == 0) r = philox4x32_10(l, gi, (uint32_t)(ui / UPB), cw, k0, k1);
      const real z = RealTraits<real>::expo(r, ui % UPB, scale(k));
      real tmax = (real)0;
      const int e1 = fc.pstart[k + 1];
      for (int e = fc.pstart[k]; e < e1; ++e)
        tmax = max(tmax, *reinterpret_cast<const real*>(reinterpret_cast<const char*>(Tcol) +
                                                        (size_t)fc.poff[e] * (sizeof(real) / 4)));
      const real t = z + tmax;
      if ((fc.store_mask >> k) & 1u) Tcol[k * kFwdBlock] = t;
      Z[k] = z;
      X |= (t <= Ts) ? (1u << k) : 0u;
    }
  }
  Ts_out = Ts;
}

// observed genotype word (event-id bit order) -> topological-position bit order
template <int PMAX>
__device__ __forceinline__ uint32_t to_topo_bits(const FwdConst& fc, const int p, const uint64_t yw) {
  uint32_t yt = 0u;
#pragma unroll
  for (int k = 0; k < PMAX; ++k)
    if (k < p) yt |= (uint32_t)((yw >> fc.ev[k]) & 1ull) << k;
  return yt;
}

template <typename real, int PMAX, bool TIMES_GIVEN>
__global__ void __launch_bounds__(kFwdBlock) estep_forward_kernel(const FwdArgs a) {
  __shared__ real s_T[PMAX * kFwdBlock];
  __shared__ real s_rtab[kRtabN];
  __shared__ real s_scale[PMAX + 1];  // fp64 mode only: -1/lambda per position, [PMAX] = -1/lambda_s

  const int cand = blockIdx.y;
  const FwdConst& fc = c_fwd[cand];
  const DevModel* __restrict__ mdl = a.models + cand;
  const int p = fc.p;
  const int lane = threadIdx.x & 31;
  constexpr int kWarps = kFwdBlock / 32;
  real* Tcol = s_T + threadIdx.x;

  for (int t = threadIdx.x; t < kRtabN; t += kFwdBlock)
    s_rtab[t] = sizeof(real) == 4 ? (real)mdl->rtab[t] : (real)mdl->rtab_d[t];
  if (sizeof(real) == 8) {
    for (int t = threadIdx.x; t <= PMAX; t += kFwdBlock) {
      double lam = t < p ? mdl->lambda[fc.ev[t]] : mdl->lambda_s;
      s_scale[t] = (real)(-1.0 / lam);
    }
  }
  __syncthreads();

  auto scale = [&](int k) -> real {
    if (sizeof(real) == 4) return (real)(k < 0 ? fc.cs : fc.c[k]);
    return s_scale[k < 0 ? PMAX : k];
  };

  const long long gw = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * kWarps;
  const uint32_t cw = a.cand_ctr[cand];
  const int flip = fc.flip;

  for (long long item = gw; item < a.n_items; item += nw) {
    const int i = (int)(item / a.chunks);
    const int ch = (int)(item - (long long)i * a.chunks);
    const uint32_t yt = to_topo_bits<PMAX>(fc, p, a.obs_bits[i]);
    const real ts_given = TIMES_GIVEN ? (real)a.times[i] : (real)0;
    const uint32_t gi = (uint32_t)(a.obs_offset + i);

    real acc[PMAX];
#pragma unroll
    for (int k = 0; k < PMAX; ++k) acc[k] = (real)0;
    real aw = (real)0, aw2 = (real)0, awd = (real)0;
    int dref = 1 << 20;

    const int l_begin = ch * a.chunk_len;
    const int l_end = min(a.L, l_begin + a.chunk_len);
    for (int l0 = l_begin; l0 < l_end; l0 += 32) {
      const int l = l0 + lane;
      const bool valid = l < l_end;
      real Z[PMAX];
      uint32_t X;
      real Ts;
      sample_forward<real, PMAX, TIMES_GIVEN>(fc, p, Tcol, (uint32_t)l, gi, a.key0, a.key1, cw, ts_given, scale, Z,
                                              X, Ts);
      const int d = __popc(X ^ yt);
      const int de = flip ? p - d : d;  // weights decrease in `de`
      const int dmin = __reduce_min_sync(0xffffffffu, valid ? de : (1 << 19));
      if (dmin < dref) {  // warp-uniform: a better sample arrived, rescale the running sums
        const real s = s_rtab[min(dref - dmin, kRtabN - 1)];
        aw *= s;
        aw2 *= s * s;
        awd *= s;
#pragma unroll
        for (int k = 0; k < PMAX; ++k) acc[k] *= s;
        dref = dmin;
      }
      const real wr = valid ? s_rtab[min(de - dref, kRtabN - 1)] : (real)0;
      aw += wr;
      aw2 = fma(wr, wr, aw2);
      awd = fma(wr, (real)d, awd);
#pragma unroll
      for (int k = 0; k < PMAX; ++k)
        if (k < p) acc[k] = fma(wr, Z[k], acc[k]);
    }

    // warp reduction in fp64 (all lanes share dref), lane s writes slot s
    double* out = a.part + ((size_t)cand * a.n_items + item) * a.PS;
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
      out[2] = v2;
      out[3] = (double)(l_end - l_begin);  // every forward sample has a positive weight (unless eps == 0)
      // natural-log scale of the record: true w = exp(scale) * w', scale = log(base) + dref * log(r)
      out[4] = mdl->log_base + (dref > 0 ? (double)dref * mdl->log_r : 0.0);
    }
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < PMAX; ++k) {
      if (k < p) {
        double v = (double)acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == (k & 31)) mine = v;
      }
    }
    if (lane < p) out[kStatHdr + lane] = mine;
  }
}

// Per-sample dump for ONE observation (the `_importance_weight_genotype` seam, src/mcem_hcbn.cpp:866-912):
// same Philox stream as the production kernel, writes raw Z (L x p, event-id columns), dist, weight pieces.
struct FwdDumpArgs {
  uint64_t obs_word;
  double time;
  const DevModel* model;
  double* Z_out;    // L x p column-major by EVENT id
  int* dist_out;    // L
  uint32_t* X_out;  // L, event-id bit order (p <= 32)
  double* Ts_out;   // L
  int L;
  uint32_t gi;
  uint32_t key0, key1, cw;
};

template <typename real, int PMAX, bool TIMES_GIVEN>
__global__ void __launch_bounds__(kFwdBlock) forward_dump_kernel(const FwdDumpArgs a) {
  __shared__ real s_T[PMAX * kFwdBlock];
  __shared__ real s_scale[PMAX + 1];
  const FwdConst& fc = c_fwd[0];
  const int p = fc.p;
  if (sizeof(real) == 8) {
    for (int t = threadIdx.x; t <= PMAX; t += kFwdBlock) {
      double lam = t < p ? a.model->lambda[fc.ev[t]] : a.model->lambda_s;
      s_scale[t] = (real)(-1.0 / lam);
    }
  }
  __syncthreads();
  auto scale = [&](int k) -> real {
    if (sizeof(real) == 4) return (real)(k < 0 ? fc.cs : fc.c[k]);
    return s_scale[k < 0 ? PMAX : k];
  };
  const uint32_t yt = to_topo_bits<PMAX>(fc, p, a.obs_word);
  for (int l = blockIdx.x * kFwdBlock + threadIdx.x; l < a.L; l += gridDim.x * kFwdBlock) {
    real Z[PMAX];
    uint32_t X;
    real Ts;
    sample_forward<real, PMAX, TIMES_GIVEN>(fc, p, s_T + threadIdx.x, (uint32_t)l, a.gi, a.key0, a.key1, a.cw,
                                            (real)a.time, scale, Z, X, Ts);
    a.dist_out[l] = __popc(X ^ yt);
    uint32_t xe = 0u;
#pragma unroll
    for (int k = 0; k < PMAX; ++k)
      if (k < p) {
        a.Z_out[(size_t)fc.ev[k] * a.L + l] = (double)Z[k];
        xe |= ((X >> k) & 1u) << fc.ev[k];
      }
    a.X_out[l] = xe;
    a.Ts_out[l] = (double)Ts;
  }
}

}  // namespace mccbn
