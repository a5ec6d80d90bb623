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
constexpr int kPoolFlush = 512;  // pool entries between two flushes of a thread's fp32 partial sums into its fp64 record

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
  // pool_resample = 1 (the reference's L-index resampling, src/mcem_hcbn.cpp:477-484); chunks == 1 in this mode
  int* dref;        // N: reference distance of the observation's relative weights (written by estep_pool_kernel)
  double* qtot;     // N: sum_k w'_k in fp64 with that fixed reference (estep_pool_resample_kernel, pass 0)
  int L;            // indices drawn per observation
  uint32_t rkey1;   // Philox key word 1 of the resampling stream
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

  // The per-thread sums are fp32 only over kPoolFlush pool entries: then they are added to the observation's fp64
  // record (L2-resident; rescaled first if the reference distance has moved since the last flush) -- the same
  // "short fp32 partial sums, fp64 beyond" rule as in the forward kernel (<= L / 32 samples per lane there).
  double* out = a.part + ((size_t)(live ? i : 0) * a.chunks + ch) * a.PS;
  int fdref = 1 << 20;  // reference distance of what the record holds; 1 << 20: nothing flushed yet
  auto flush = [&]() {
    if (!live || dref == (1 << 20)) return;
    if (fdref == (1 << 20)) {
      out[0] = (double)aw, out[1] = (double)aw2, out[2] = (double)awd;
#pragma unroll
      for (int j = 0; j < PMAX; ++j)
        if (j < p) out[kStatHdr + j] = (double)acc[j];
    } else {
      const double sc = fdref != dref ? mdl->rtab_d[min(fdref - dref, kRtabN - 1)] : 1.0;  // dref only decreases
      out[0] = out[0] * sc + (double)aw;
      out[1] = out[1] * sc * sc + (double)aw2;
      out[2] = out[2] * sc + (double)awd;
#pragma unroll
      for (int j = 0; j < PMAX; ++j)
        if (j < p) out[kStatHdr + j] = out[kStatHdr + j] * sc + (double)acc[j];
    }
    fdref = dref;
    aw = aw2 = awd = 0.f;
#pragma unroll
    for (int j = 0; j < PMAX; ++j) acc[j] = 0.f;
  };

  const int k_begin = ch * a.chunk_len, k_end = min(a.K, k_begin + a.chunk_len);
  for (int k0 = k_begin; k0 < k_end; k0 += kPoolTile) {
    __syncthreads();
    if (k0 > k_begin && ((k0 - k_begin) % kPoolFlush) == 0) flush();
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
  flush();
  if (live) {
    out[3] = (double)(k_end - k_begin);
    out[4] = mdl->log_base + (dref > 0 ? (double)dref * mdl->log_r : 0.0);
    if (a.dref) a.dref[i] = dref;
  }
}

// ---- tensor-core variant of estep_pool_kernel -------------------------------------------------------------------
// The weighted sums  S_ij = sum_k q_ik Z_kj  ARE a contraction: (observations x pool entries) weights against (pool
// entries x events) waiting times.  In estep_pool_kernel they cost PMAX FFMA + PMAX/4 broadcast LDS.128 per
// (observation, entry) pair -- 40 of the kernel's 112 instructions per pair.  Here a warp (32 observations, lane =
// observation for everything that is NOT a contraction: sampling time, rank search, distance, weight, the three header
// sums) hands the 32 x 8 weights of eight pool entries to `mma.sync.m16n8k8` (TF32 inputs, fp32 accumulation) against
// the entries' 8 x PMAX waiting times.  fp32 accuracy is kept by splitting both operands into a TF32 head and the
// fp32 remainder (q = qh + ql, Z = Zh + Zl: three products qh Zh + ql Zh + qh Zl; the dropped ql Zl term is 2^-22
// relative).  The accumulator of observation row r lives in the lanes of its fragment row, so when a lane lowers its
// running reference distance the scale factor travels to those lanes by shuffle (rare: at most p times per row).
__device__ __forceinline__ uint32_t pool_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void pool_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#ifndef MCCBN_POOL_MMA_TERMS
#define MCCBN_POOL_MMA_TERMS 3  // 3: qh Zh + ql Zh + qh Zl (fp32 accuracy); fewer only in profiling builds
#endif
constexpr int kPoolZStride = kPoolTile + 4;  // Z tile stored [event][entry], stride 68: B-fragment loads hit 32 banks
constexpr int kPoolQStride = 12;             // weights staged [observation][8 entries], stride 12: A-fragment loads too

template <int PMAX, bool TIMES_GIVEN>
__global__ void __launch_bounds__(kPoolBlock, 4) estep_pool_mma_kernel(const PoolArgs a,
                                                                       const __grid_constant__ PosetProg pg) {
  static_assert(PMAX % 8 == 0 && (PMAX & (PMAX - 1)) == 0, "PMAX: power of two, multiple of the mma N = 8");
  constexpr int NCB = PMAX / 8;  // column blocks of the accumulator
  __shared__ __align__(16) float s_Tt[kPoolTile * PMAX];
  __shared__ uint32_t s_Mt[kPoolTile * (PMAX + 1)];
  __shared__ uint32_t s_Zh[PMAX * kPoolZStride], s_Zl[PMAX * kPoolZStride];
  __shared__ __align__(16) uint32_t s_Q[kPoolBlock / 32][2][32 * kPoolQStride];
  __shared__ float s_rtab[kRtabN];
  const DevModel* __restrict__ mdl = a.model;
  const int p = pg.p;
  for (int t = threadIdx.x; t < kRtabN; t += kPoolBlock) s_rtab[t] = mdl->rtab[t];
  const int flip = mdl->flip;
  const uint32_t key1 = iter_key(a.key1, mdl);
  const int ch = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;  // fragment coordinates: groupID, thread-in-group
  const int i = blockIdx.x * kPoolBlock + threadIdx.x;
  const int i_warp = blockIdx.x * kPoolBlock + w * 32;  // observation of fragment row 0 of this warp
  const bool live = i < a.N;
  const uint32_t yt = live ? to_topo_bits<PMAX>(pg, a.obs_bits[i]) : 0u;
  const float ts_given = (TIMES_GIVEN && live) ? (float)a.times[i] : 0.f;
  const uint32_t gi = (uint32_t)(a.obs_offset + i);
  const float cs = a.scale->cs;

  float acc[2][NCB][4];  // accumulator fragments: [row block][column block][c0..c3]
#pragma unroll
  for (int rb = 0; rb < 2; ++rb)
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[rb][cb][e] = 0.f;
  float aw = 0.f, aw2 = 0.f, awd = 0.f;
  int dref = 1 << 20;

  double* out = a.part + ((size_t)(live ? i : 0) * a.chunks + ch) * a.PS;
  int fdref = 1 << 20;  // reference distance of what this lane's record holds; 1 << 20: nothing flushed yet
  // fp32 partial sums -> the observations' fp64 records (same rule as estep_pool_kernel: at most kPoolFlush entries in fp32).
  // The header sums belong to the lane, the S_ij of fragment row r to the lanes that hold that row.
  auto flush = [&]() {
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = 16 * rb + g + 8 * h;
        const int dref_r = __shfl_sync(0xffffffffu, dref, row), fdref_r = __shfl_sync(0xffffffffu, fdref, row);
        const int i_r = i_warp + row;
        if (i_r < a.N && dref_r != (1 << 20)) {
          double* o = a.part + ((size_t)i_r * a.chunks + ch) * a.PS + kStatHdr;
          const bool first = fdref_r == (1 << 20);
          const double sc = (!first && fdref_r != dref_r) ? mdl->rtab_d[min(fdref_r - dref_r, kRtabN - 1)] : 1.0;
#pragma unroll
          for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = 8 * cb + 2 * t4 + e;
              if (col < p) o[col] = first ? (double)acc[rb][cb][2 * h + e] : o[col] * sc + (double)acc[rb][cb][2 * h + e];
            }
        }
      }
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[rb][cb][e] = 0.f;
    if (live && dref != (1 << 20)) {
      if (fdref == (1 << 20)) {
        out[0] = (double)aw, out[1] = (double)aw2, out[2] = (double)awd;
      } else {
        const double sc = fdref != dref ? mdl->rtab_d[min(fdref - dref, kRtabN - 1)] : 1.0;  // dref only decreases
        out[0] = out[0] * sc + (double)aw;
        out[1] = out[1] * sc * sc + (double)aw2;
        out[2] = out[2] * sc + (double)awd;
      }
      fdref = dref;
    }
    aw = aw2 = awd = 0.f;
  };

  uint32_t* Qh = s_Q[w][0];
  uint32_t* Ql = s_Q[w][1];
  const int k_begin = ch * a.chunk_len, k_end = min(a.K, k_begin + a.chunk_len);
  for (int k0 = k_begin; k0 < k_end; k0 += kPoolTile) {
    __syncthreads();
    if (k0 > k_begin && ((k0 - k_begin) % kPoolFlush) == 0) flush();
    const int nk = min(kPoolTile, k_end - k0);
    for (int t = threadIdx.x; t < nk * PMAX / 4; t += kPoolBlock)  // coalesced float4 copy of the sorted times
      reinterpret_cast<float4*>(s_Tt)[t] = reinterpret_cast<const float4*>(a.Tp + (size_t)k0 * PMAX)[t];
    for (int t = threadIdx.x; t < kPoolTile * PMAX / 4; t += kPoolBlock) {  // Z tile: transposed, split into TF32 head + remainder
      const int e = t / (PMAX / 4), cg = t % (PMAX / 4);
      const float4 z = e < nk ? reinterpret_cast<const float4*>(a.Zp + (size_t)k0 * PMAX)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t hi = pool_tf32(zz[c4]);
        s_Zh[(4 * cg + c4) * kPoolZStride + e] = hi;
        s_Zl[(4 * cg + c4) * kPoolZStride + e] = __float_as_uint(zz[c4] - __uint_as_float(hi));
      }
    }
    for (int t = threadIdx.x; t < nk * (PMAX + 1); t += kPoolBlock) s_Mt[t] = a.Mp[(size_t)k0 * (PMAX + 1) + t];
    __syncthreads();
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int kk = 0; kk < nk; kk += 8) {
      // --- lane = observation: distances of the eight entries kk .. kk + 7
      int dj[8], dej[8];
      int dmin = dref;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float Ts = ts_given;
        if (!TIMES_GIVEN) {
          if ((j & 3) == 0) r = philox4x32_10((uint32_t)((k0 + kk + j) >> 2), gi, 0u, a.cw, a.key0, key1);
          const uint32_t bits = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
          Ts = fast_lg2(u01_open0(bits)) * cs;
        }
        const float* Tk = s_Tt + (kk + j) * PMAX;
        int rank = 0;
#pragma unroll
        for (int step = PMAX / 2; step >= 1; step >>= 1)
          if (Tk[rank + step - 1] <= Ts) rank += step;
        if (Tk[rank] <= Ts) rank += 1;
        const uint32_t X = s_Mt[(kk + j) * (PMAX + 1) + rank];
        const int d = __popc(X ^ yt);
        dj[j] = d;
        dej[j] = (kk + j < nk) ? (flip ? p - d : d) : (1 << 19);  // beyond the tile: the zero tail of the weight table
        dmin = min(dmin, dej[j]);
      }
      // --- running reference distance of this observation; the scale of what has been accumulated so far
      float myscale = 1.f;
      if (dmin < dref) {
        myscale = s_rtab[min(dref - dmin, kRtabN - 1)];
        aw *= myscale;
        aw2 *= myscale * myscale;
        awd *= myscale;
        dref = dmin;
      }
      uint32_t qh[8], ql[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float q = s_rtab[min(dej[j] - dref, kRtabN - 1)];
        aw += q;
        aw2 = fmaf(q, q, aw2);
        awd = fmaf(q, (float)dj[j], awd);
        qh[j] = pool_tf32(q);
        ql[j] = __float_as_uint(q - __uint_as_float(qh[j]));
      }
      if (__ballot_sync(0xffffffffu, myscale != 1.f)) {  // some row moved its reference: rescale its accumulator row
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          const float s0 = __shfl_sync(0xffffffffu, myscale, 16 * rb + g), s1 = __shfl_sync(0xffffffffu, myscale, 16 * rb + g + 8);
#pragma unroll
          for (int cb = 0; cb < NCB; ++cb) {
            acc[rb][cb][0] *= s0, acc[rb][cb][1] *= s0;
            acc[rb][cb][2] *= s1, acc[rb][cb][3] *= s1;
          }
        }
      }
      // --- the 32 x 8 weights of this group -> A fragments (through the warp's staging rows)
      __syncwarp();
      *reinterpret_cast<uint4*>(Qh + lane * kPoolQStride) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
      *reinterpret_cast<uint4*>(Qh + lane * kPoolQStride + 4) = make_uint4(qh[4], qh[5], qh[6], qh[7]);
      *reinterpret_cast<uint4*>(Ql + lane * kPoolQStride) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
      *reinterpret_cast<uint4*>(Ql + lane * kPoolQStride + 4) = make_uint4(ql[4], ql[5], ql[6], ql[7]);
      __syncwarp();
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        const int b0 = (16 * rb + g) * kPoolQStride + t4, b1 = b0 + 8 * kPoolQStride;
        ah[rb][0] = Qh[b0], ah[rb][1] = Qh[b1], ah[rb][2] = Qh[b0 + 4], ah[rb][3] = Qh[b1 + 4];
        al[rb][0] = Ql[b0], al[rb][1] = Ql[b1], al[rb][2] = Ql[b0 + 4], al[rb][3] = Ql[b1 + 4];
      }
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        const int zb = (8 * cb + g) * kPoolZStride + kk + t4;
        const uint32_t bh0 = s_Zh[zb], bh1 = s_Zh[zb + 4], bl0 = s_Zl[zb], bl1 = s_Zl[zb + 4];
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          pool_mma_tf32(acc[rb][cb], ah[rb], bh0, bh1);
          if (MCCBN_POOL_MMA_TERMS >= 2) pool_mma_tf32(acc[rb][cb], al[rb], bh0, bh1);
          if (MCCBN_POOL_MMA_TERMS >= 3) pool_mma_tf32(acc[rb][cb], ah[rb], bl0, bl1);
        }
      }
    }
  }
  flush();
  if (live) {
    out[3] = (double)(k_end - k_begin);
    out[4] = mdl->log_base + (dref > 0 ? (double)dref * mdl->log_r : 0.0);
    if (a.dref) a.dref[i] = dref;
  }
}

// pool_resample = 1: the reference does not average over the whole pool but draws L indices with replacement from
// Discrete(q) and averages d and Z over the drawn rows (src/mcem_hcbn.cpp:477-484, rdiscrete_std :44-54).  The K weights
// of an observation are never stored: L iid indices from Discrete(q) are distributed like the cells that the ORDER
// STATISTICS of L iid uniforms fall into along the cumulative weights, and those can be generated one by one in
// increasing order (the smallest of m iid U(a, 1) is 1 - (1 - a) V^(1/m)), i.e. during ONE more sweep over the pool.
//   PASS 0: qtot[i] = sum_k w'_k in fp64 for the fixed reference distance dref[i] (the total the order statistics scale to)
//   PASS 1: sweep with the running cumulative weight; entry k is drawn cnt_k times; the record's sum w' d and sum w' Z_j
//           become sum w' x (resampled mean), so that finalize_kernel's ratios are the reference's resampled means
//           (sum w' itself, hence obs_llhood = log(sum q / K), is untouched: every drawn row has weight sum q / K, :491)
// Same tiles, sampling-time stream and rank search as estep_pool_kernel; one thread per observation, whole pool.
template <int PMAX, bool TIMES_GIVEN, int PASS>
__global__ void __launch_bounds__(kPoolBlock) estep_pool_resample_kernel(const PoolArgs a,
                                                                         const __grid_constant__ PosetProg pg) {
  __shared__ __align__(16) float s_Tt[kPoolTile * PMAX];
  __shared__ __align__(16) float s_Zt[kPoolTile * PMAX];
  __shared__ uint32_t s_Mt[kPoolTile * (PMAX + 1)];
  __shared__ float s_rtab[kRtabN];
  const DevModel* __restrict__ mdl = a.model;
  const int p = pg.p;
  for (int t = threadIdx.x; t < kRtabN; t += kPoolBlock) s_rtab[t] = mdl->rtab[t];
  const int flip = mdl->flip;
  const uint32_t key1 = iter_key(a.key1, mdl), rkey1 = iter_key(a.rkey1, mdl);
  const int i = blockIdx.x * kPoolBlock + threadIdx.x;
  const bool live = i < a.N;
  const uint32_t yt = live ? to_topo_bits<PMAX>(pg, a.obs_bits[i]) : 0u;
  const float ts_given = (TIMES_GIVEN && live) ? (float)a.times[i] : 0.f;
  const uint32_t gi = (uint32_t)(a.obs_offset + i);
  const float cs = a.scale->cs;
  const int dref = live ? a.dref[i] : 0;

  double C = 0.0;                              // running cumulative weight
  const double S = (PASS == 1 && live) ? a.qtot[i] : 0.0;
  int m = (PASS == 1 && live) ? a.L : 0;       // order statistics still to come
  uint32_t draw = 0;                           // index of the next uniform of the resampling stream
  double rem = 1.0;                            // 1 - (current order statistic)
  double thr = 0.0;                            // S * (current order statistic): the next draw falls into the first
                                               // entry whose cumulative weight exceeds it
  auto next_order_statistic = [&]() {
    const uint4 r = philox4x32_10(draw >> 1, gi, 1u, a.cw, a.key0, rkey1);
    const double v = (draw & 1u) ? u01_open0_d(r.z, r.w) : u01_open0_d(r.x, r.y);
    ++draw;
    rem *= exp(log(v) / (double)m);
    thr = S * (1.0 - rem);
  };
  if (m > 0) next_order_statistic();
  double sd = 0.0;
  double acc[PMAX];  // sums of up to L rows: fp64 (this sweep is bound by the rank searches, not by these adds)
#pragma unroll
  for (int j = 0; j < PMAX; ++j) acc[j] = 0.0;
  int last_k = -1;  // last entry with a positive weight (takes the draws a rounding of S may leave over)

  for (int k0 = 0; k0 < a.K; k0 += kPoolTile) {
    __syncthreads();
    const int nk = min(kPoolTile, a.K - k0);
    for (int t = threadIdx.x; t < nk * PMAX / 4; t += kPoolBlock) {
      reinterpret_cast<float4*>(s_Tt)[t] = reinterpret_cast<const float4*>(a.Tp + (size_t)k0 * PMAX)[t];
      if (PASS == 1) reinterpret_cast<float4*>(s_Zt)[t] = reinterpret_cast<const float4*>(a.Zp + (size_t)k0 * PMAX)[t];
    }
    for (int t = threadIdx.x; t < nk * (PMAX + 1); t += kPoolBlock) s_Mt[t] = a.Mp[(size_t)k0 * (PMAX + 1) + t];
    __syncthreads();
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int kk = 0; kk < nk; ++kk) {
      float Ts = ts_given;
      if (!TIMES_GIVEN) {
        const int k = k0 + kk;
        if ((kk & 3) == 0) r = philox4x32_10((uint32_t)(k >> 2), gi, 0u, a.cw, a.key0, key1);
        const uint32_t bits = (kk & 3) == 0 ? r.x : ((kk & 3) == 1 ? r.y : ((kk & 3) == 2 ? r.z : r.w));
        Ts = fast_lg2(u01_open0(bits)) * cs;
      }
      const float* Tk = s_Tt + kk * PMAX;
      int rank = 0;
#pragma unroll
      for (int step = PMAX / 2; step >= 1; step >>= 1)
        if (Tk[rank + step - 1] <= Ts) rank += step;
      if (Tk[rank] <= Ts) rank += 1;
      const uint32_t X = s_Mt[kk * (PMAX + 1) + rank];
      const int d = __popc(X ^ yt);
      const int de = flip ? p - d : d;
      const float wr = s_rtab[min(max(de - dref, 0), kRtabN - 1)];
      C += (double)wr;
      if (PASS == 1 && wr > 0.f) {
        last_k = k0 + kk;
        int cnt = 0;
        while (m > 0 && thr < C) {  // the order statistics that fall into this entry's cell
          ++cnt;
          --m;
          if (m > 0) next_order_statistic();
        }
        if (cnt > 0) {
          const double fc = (double)cnt;
          sd += (double)(cnt * d);
#pragma unroll
          for (int j = 0; j < PMAX; ++j) acc[j] = fma(fc, (double)s_Zt[kk * PMAX + j], acc[j]);
        }
      }
    }
  }
  if (!live) return;
  if (PASS == 0) {
    a.qtot[i] = C;
    return;
  }
  double* out = a.part + (size_t)i * a.PS;
  const double W = out[0], invL = 1.0 / (double)a.L;
  double left = (double)m;  // > 0 only if S (pass 0) and C (this pass) differ in the last bit
  int d_last = 0;
  if (m > 0 && last_k >= 0) {
    float Ts = ts_given;
    if (!TIMES_GIVEN) {
      const uint4 r = philox4x32_10((uint32_t)(last_k >> 2), gi, 0u, a.cw, a.key0, key1);
      const int q = last_k & 3;
      Ts = fast_lg2(u01_open0(q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w)))) * cs;
    }
    int rank = 0;
    for (int j = 0; j < PMAX; ++j) rank += a.Tp[(size_t)last_k * PMAX + j] <= Ts ? 1 : 0;  // row is sorted, padded with +inf
    d_last = __popc(a.Mp[(size_t)last_k * (PMAX + 1) + rank] ^ yt);
  }
  out[2] = W * (sd + left * (double)d_last) * invL;
#pragma unroll
  for (int j = 0; j < PMAX; ++j)
    if (j < p) {
      const double zl = (m > 0 && last_k >= 0) ? left * (double)a.Zp[(size_t)last_k * PMAX + j] : 0.0;
      out[kStatHdr + j] = W * (acc[j] + zl) * invL;
    }
}

}  // namespace mccbn
