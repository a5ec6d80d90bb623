// kernels_pool.cuh -- the reference's "pool" scheme (src/mcem_hcbn.cpp:460-491, :650-654, :668-677) for sm_100a.
//
// Reference: once per EM iteration a pool of K = p*L prior draws (Z_k, T_k) is generated (sample_times :184-213);
// for every observation i the pool genotypes X_k = [T_k <= T_s,ik] are formed with K fresh sampling times (or the
// given t_i) (generate_genotypes :215-236), q_k = eps^d (1-eps)^(p-d) with d = d(X_k, Y_i), and L indices are
// resampled with replacement from Discrete(q) (rdiscrete_std :44-54); all L samples get the weight sum(q)/K.
// Hence  P^(Y_i) = sum_k q_k / K  deterministically given the pool, and the resampled means of d and Z are
// unbiased estimates of  sum_k q_k g_k / sum_k q_k.  The device path evaluates those conditional expectations
// directly (Rao-Blackwellisation of the resampling step: same estimand, less Monte-Carlo noise; documented in
// DESIGN.md) -- the O(N K p) comparison work is the reference's.
//
// Mapping: pool entries are staged tile by tile in shared memory; one THREAD per observation sweeps the tile with its
// p+2 running sums in registers.  The genotype of entry k at sampling time t is a PREFIX of its events sorted by event
// time, so the pool stores per entry the sorted times and the p + 1 prefix genotypes: X_k(t) costs a 5-step binary
// search in the entry's sorted row (per-lane shared-memory addresses inside one 128-byte row: conflict-free) and one
// mask load instead of p compares.
#pragma once
#include "device_types.cuh"
#include "kernels_forward.cuh"
#include "philox.cuh"

namespace mccbn {

#ifndef MCCBN_POOL_UNROLL
#define MCCBN_POOL_UNROLL 4  // 1 / 2 / 4 / 8 -> 35.0 / 32.4 / 26.3 / 26.5 ms on cfg 5 (independent rank searches in flight)
#endif
#define MCCBN_PRAGMA_(x) _Pragma(#x)
#define MCCBN_PRAGMA(x) MCCBN_PRAGMA_(x)
#define MCCBN_POOL_UNROLL_PRAGMA MCCBN_PRAGMA(unroll MCCBN_POOL_UNROLL)

constexpr int kPoolBlock = 128;  // observations per block
constexpr int kPoolTile = 64;    // pool entries per shared-memory tile

// Pool generation: entry k draws Z_j ~ Exp(lambda_j) and propagates T_j (sample_times, src/mcem_hcbn.cpp:184-213).
// Output per entry: Tp[k][PMAX] the event times in ASCENDING order (padded with +inf), Mp[k][PMAX + 1] the prefix
// genotypes in topological-position bit order (Mp[k][r] = events with the r smallest times; Mp[k][0] = 0),
// Zp[k][PMAX] the waiting times by topological position (padded with 0).
template <int PMAX>
__global__ void __launch_bounds__(kFwdBlock) pool_generate_kernel(float* __restrict__ Tp, uint32_t* __restrict__ Mp,
                                                                  float* __restrict__ Zp, int K, uint32_t key0,
                                                                  uint32_t key1_purpose, uint32_t cw,
                                                                  const DevModel* __restrict__ mdl,
                                                                  const FwdScale* __restrict__ sc,
                                                                  const __grid_constant__ PosetProg pg) {
  const uint32_t key1 = iter_key(key1_purpose, mdl);
  __shared__ float s_scale[PMAX + 1];
  load_scales<float, PMAX>(pg, mdl, sc, s_scale);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_T = reinterpret_cast<float*>(smem_raw);
  using ZS = ZStage<float, PMAX>;
  float* s_Z = s_T + (PMAX + 1) * kFwdBlock;
  s_T[PMAX * kFwdBlock + threadIdx.x] = 0.f;
  const uint32_t tcol_s = (uint32_t)__cvta_generic_to_shared(s_T + threadIdx.x);
  const uint32_t zrow_s = (uint32_t)__cvta_generic_to_shared(s_Z + threadIdx.x * ZS::kStride);
  __syncthreads();
  const ScaleSrc<float, PMAX> scale{s_scale};
  const int p = pg.p;
  for (int k = blockIdx.x * kFwdBlock + threadIdx.x; k < K; k += gridDim.x * kFwdBlock) {
    uint32_t X;
    float Ts;
    // TIMES_GIVEN = true: no sampling time is drawn here (it is drawn per (observation, entry) later)
    sample_forward<float, PMAX, true>(pg, tcol_s, zrow_s, (uint32_t)k, 0xffffffffu, key0, key1, cw, 0.f, scale, X, Ts);
    // insertion sort of (time, position) in this thread's T column (in place; the column is rewritten by the next draw)
    unsigned char ord[PMAX];
    for (int j = 0; j < p; ++j) {
      const float t = s_T[j * kFwdBlock + threadIdx.x];
      int r = j;
      while (r > 0 && s_T[(r - 1) * kFwdBlock + threadIdx.x] > t) {
        s_T[r * kFwdBlock + threadIdx.x] = s_T[(r - 1) * kFwdBlock + threadIdx.x];
        ord[r] = ord[r - 1];
        --r;
      }
      s_T[r * kFwdBlock + threadIdx.x] = t;
      ord[r] = (unsigned char)j;
    }
    uint32_t m = 0u;
    Mp[(size_t)k * (PMAX + 1)] = 0u;
    for (int j = 0; j < PMAX; ++j) {
      if (j < p) m |= 1u << ord[j];
      Tp[(size_t)k * PMAX + j] = j < p ? s_T[j * kFwdBlock + threadIdx.x] : __int_as_float(0x7f800000);
      Mp[(size_t)k * (PMAX + 1) + j + 1] = m;
      Zp[(size_t)k * PMAX + j] = j < p ? s_Z[threadIdx.x * ZS::kStride + j] * scale(j) : 0.f;  // Z = e c
    }
  }
}

struct PoolArgs {
  const uint64_t* obs_bits;
  const double* times;
  const DevModel* model;
  const FwdScale* scale;
  const float* Tp;     // K x PMAX sorted event times
  const uint32_t* Mp;  // K x (PMAX + 1) prefix genotypes
  const float* Zp;     // K x PMAX
  double* part;     // N x chunks x PS
  int N, K, chunks, chunk_len;  // chunk over pool entries (multiple of kPoolTile)
  long long obs_offset;
  uint32_t key0, key1, cw;
  int PS;
};

template <int PMAX, bool TIMES_GIVEN>
__global__ void __launch_bounds__(kPoolBlock) estep_pool_kernel(const PoolArgs a, const __grid_constant__ PosetProg pg) {
  __shared__ __align__(16) float s_Tt[kPoolTile * PMAX];
  __shared__ __align__(16) float s_Zt[kPoolTile * PMAX];
  __shared__ uint32_t s_Mt[kPoolTile * (PMAX + 1)];
  static_assert((PMAX & (PMAX - 1)) == 0, "the rank search needs a power-of-two row length");
  __shared__ float s_rtab[kRtabN];
  const DevModel* __restrict__ mdl = a.model;
  const int p = pg.p;
  for (int t = threadIdx.x; t < kRtabN; t += kPoolBlock) s_rtab[t] = mdl->rtab[t];
  const int flip = mdl->flip;
  const uint32_t key1 = iter_key(a.key1, mdl);
  const int ch = blockIdx.y;
  const int i = blockIdx.x * kPoolBlock + threadIdx.x;
  const bool live = i < a.N;
  const uint32_t yt = live ? to_topo_bits<PMAX>(pg, a.obs_bits[i]) : 0u;
  const float ts_given = (TIMES_GIVEN && live) ? (float)a.times[i] : 0.f;
  const uint32_t gi = (uint32_t)(a.obs_offset + i);
  const float cs = a.scale->cs;

  float acc[PMAX];
#pragma unroll
  for (int j = 0; j < PMAX; ++j) acc[j] = 0.f;
  float aw = 0.f, aw2 = 0.f, awd = 0.f;
  int dref = 1 << 20;

  const int k_begin = ch * a.chunk_len, k_end = min(a.K, k_begin + a.chunk_len);
  for (int k0 = k_begin; k0 < k_end; k0 += kPoolTile) {
    __syncthreads();
    const int nk = min(kPoolTile, k_end - k0);
    for (int t = threadIdx.x; t < nk * PMAX / 4; t += kPoolBlock) {  // coalesced float4 copy of the tile
      reinterpret_cast<float4*>(s_Tt)[t] = reinterpret_cast<const float4*>(a.Tp + (size_t)k0 * PMAX)[t];
      reinterpret_cast<float4*>(s_Zt)[t] = reinterpret_cast<const float4*>(a.Zp + (size_t)k0 * PMAX)[t];
    }
    for (int t = threadIdx.x; t < nk * (PMAX + 1); t += kPoolBlock) s_Mt[t] = a.Mp[(size_t)k0 * (PMAX + 1) + t];
    __syncthreads();
    uint4 r = make_uint4(0, 0, 0, 0);
    MCCBN_POOL_UNROLL_PRAGMA
    for (int kk = 0; kk < nk; ++kk) {
      float Ts = ts_given;
      if (!TIMES_GIVEN) {
        const int k = k0 + kk;
        if ((kk & 3) == 0) r = philox4x32_10((uint32_t)(k >> 2), gi, 0u, a.cw, a.key0, key1);
        const uint32_t bits = (kk & 3) == 0 ? r.x : ((kk & 3) == 1 ? r.y : ((kk & 3) == 2 ? r.z : r.w));
        Ts = fast_lg2(u01_open0(bits)) * cs;
      }
      // rank = #{j : T_kj <= Ts} by binary search in the entry's sorted row, X = prefix genotype of that rank
      const float* Tk = s_Tt + kk * PMAX;
      int rank = 0;
#pragma unroll
      for (int step = PMAX / 2; step >= 1; step >>= 1)
        if (Tk[rank + step - 1] <= Ts) rank += step;
      if (Tk[rank] <= Ts) rank += 1;  // rank in [0, PMAX]: the loop alone reaches PMAX - 1
      const uint32_t X = s_Mt[kk * (PMAX + 1) + rank];
      const int d = __popc(X ^ yt);
      const int de = flip ? p - d : d;
      if (de < dref) {  // per-thread running reference (lanes are different observations); rare
        const float s = s_rtab[min(dref - de, kRtabN - 1)];
        aw *= s;
        aw2 *= s * s;
        awd *= s;
#pragma unroll
        for (int j = 0; j < PMAX; ++j) acc[j] *= s;
        dref = de;
      }
      const float wr = s_rtab[min(de - dref, kRtabN - 1)];
      aw += wr;
      aw2 = fmaf(wr, wr, aw2);
      awd = fmaf(wr, (float)d, awd);
      const float4* Zk = reinterpret_cast<const float4*>(s_Zt + kk * PMAX);
#pragma unroll
      for (int q = 0; q < PMAX / 4; ++q) {
        const float4 z4 = Zk[q];
        acc[4 * q] = fmaf(wr, z4.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(wr, z4.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(wr, z4.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(wr, z4.w, acc[4 * q + 3]);
      }
    }
  }
  if (live) {
    double* out = a.part + ((size_t)i * a.chunks + ch) * a.PS;
    out[0] = (double)aw;
    out[1] = (double)aw2;
    out[2] = (double)awd;
    out[3] = (double)(k_end - k_begin);
    out[4] = mdl->log_base + (dref > 0 ? (double)dref * mdl->log_r : 0.0);
#pragma unroll
    for (int j = 0; j < PMAX; ++j)
      if (j < p) out[kStatHdr + j] = (double)acc[j];
  }
}

}  // namespace mccbn
