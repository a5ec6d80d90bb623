// kernels_fp64.cuh -- "supplied randomness" kernels: the fp64 parity tier (no RNG on either side).
//
// From host-supplied waiting times Z (L x p) and sampling times T_s (L) these kernels redo, in IEEE fp64 and with
// the reference's association order, exactly the index/weight work of
//   sample_genotypes      src/mcem_hcbn.cpp:160-175  (T_j = Z_j + max(0, max_pa T), X_j = [T_j <= T_s])
//   hamming_dist_mat      src/mcem_hcbn.cpp:326-330
//   forward weights       src/mcem_hcbn.cpp:386-388  (table wtab[d] built on the host with the same libm pow)
//   generate_mutation_times / cbn_density_log   src/add_remove.cpp:156-216  (from supplied uniforms)
// so that X, d and the incompatibility flags are bit-exact and w, sum w, E[d], E[Z], log q are <= 1e-10.
// They are HBM-bound test seams (8 (p+1) bytes read per sample), not the production path.
#pragma once
#include "device_types.cuh"
#include "philox.cuh"

namespace mccbn {

// T_sum[l, ev] for every sample; thread per sample, local array indexed by topological position
__global__ void propagate_times_kernel(const double* __restrict__ Z, double* __restrict__ T_sum, int L,
                                       const DevPoset* pos) {
  const int p = pos->p;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
    double T[64];
    for (int k = 0; k < p; ++k) {
      const int ev = pos->ev[k];
      double tmax = 0.0;
      for (uint64_t m = pos->parent_pos[k]; m; m &= m - 1) {
        const double tu = T[__ffsll((long long)m) - 1];
        tmax = (tu > tmax) ? tu : tmax;
      }
      const double t = Z[(size_t)ev * L + l] + tmax;
      T[k] = t;
      T_sum[(size_t)ev * L + l] = t;
    }
  }
}

// per-sample outputs for one genotype: X (L x p int, optional), dist, w
__global__ void forward_from_times_kernel(const double* __restrict__ T_sum, const double* __restrict__ Ts,
                                          uint64_t yw, int L, int p, const double* __restrict__ wtab,
                                          int* __restrict__ X, int* __restrict__ dist, double* __restrict__ w) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
    const double ts = Ts[l];
    uint64_t x = 0;
    for (int j = 0; j < p; ++j) {
      const bool b = T_sum[(size_t)j * L + l] <= ts;
      x |= (uint64_t)b << j;
      if (X) X[(size_t)j * L + l] = b;
    }
    const int d = __popcll(x ^ yw);
    if (dist) dist[l] = d;
    if (w) w[l] = wtab[d];
  }
}

// E-step from supplied draws shared by all observations: block per observation, threads over samples.
// Writes one absolute-weight partial record per observation: [sum w, sum w^2, sum w d, #{w>0}, 0, sum w Z_k (topo pos k)]
constexpr int kF64Block = 256;
__global__ void __launch_bounds__(kF64Block) estep_from_times_kernel(
    const double* __restrict__ Z, const double* __restrict__ T_sum, const double* __restrict__ Ts,
    const uint64_t* __restrict__ obs_bits, const double* __restrict__ times, int times_given, int N, int L,
    const DevPoset* pos, const double* __restrict__ wtab, double* __restrict__ part, int PS) {
  __shared__ double s_red[kF64Block / 32][kStatHdr + 64];
  const int p = pos->p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    const uint64_t yw = obs_bits[i];
    double acc[kStatHdr + 64];
    for (int s = 0; s < PS; ++s) acc[s] = 0.0;
    for (int l = threadIdx.x; l < L; l += kF64Block) {
      const double ts = times_given ? times[i] : Ts[l];
      uint64_t x = 0;
      for (int j = 0; j < p; ++j) x |= (uint64_t)(T_sum[(size_t)j * L + l] <= ts) << j;
      const int d = __popcll(x ^ yw);
      const double w = wtab[d];
      acc[0] += w;
      acc[1] += w * w;
      acc[2] += w * (double)d;
      acc[3] += (w > 0.0) ? 1.0 : 0.0;
      for (int k = 0; k < p; ++k) acc[kStatHdr + k] += Z[(size_t)pos->ev[k] * L + l] * w;
    }
    for (int s = 0; s < PS; ++s) {
      double v = acc[s];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_red[warp][s] = v;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < PS; s += kF64Block) {
      double t = 0.0;
      for (int w2 = 0; w2 < kF64Block / 32; ++w2) t += s_red[w2][s];
      part[(size_t)i * PS + s] = (s == 4) ? 0.0 : t;  // absolute weights: log scale 0
    }
    __syncthreads();
  }
}

// Pool scheme from a supplied pool (src/mcem_hcbn.cpp:215-236, :326-330, :676): d_pool[i, k] = Hamming distance between
// observation i and the pool genotype X_k = [T_pool(k, j) <= T_s] (T_s = Ts[k], or times[i] when given).
// d_out is K x N column-major (column i = the reference's d_pool vector of observation i).
__global__ void pool_dist_kernel(const double* __restrict__ T_sum, const double* __restrict__ Ts,
                                 const uint64_t* __restrict__ obs_bits, const double* __restrict__ times, int times_given,
                                 int N, int K, int p, int* __restrict__ d_out) {
  const int i = blockIdx.y;
  if (i >= N) return;
  const uint64_t yw = obs_bits[i];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    const double ts = times_given ? times[i] : Ts[k];
    uint64_t x = 0;
    for (int j = 0; j < p; ++j) x |= (uint64_t)(T_sum[(size_t)j * K + k] <= ts) << j;
    d_out[(size_t)i * K + k] = __popcll(x ^ yw);
  }
}

// Source of the randomness of the conditional seam kernel below.  enabled == 0: X, U and Ts are the caller's arrays
// (supplied-randomness tier).  enabled == 1: one genotype `x_word` shared by all samples, 53-bit Philox uniforms
// (counter = (sample l, 0, event >> 1, cw)) and, if draw_ts, T_s ~ Exp(lambda_s) per sample (block 64) -- the legacy
// OT-CBN export `drawHiddenVarsSamples` (src/mcem.cpp:214-248), which returns the draws themselves.
struct SeamRng {
  int enabled, draw_ts;
  uint32_t k0, k1, cw;
  uint64_t x_word;
  double time, lambda_s;
  double* Ts_out;  // optional: the sampling time used by sample l
};

// Conditional proposal (generate_mutation_times, src/add_remove.cpp:156-203; drawHiddenVarsSample, src/mcem.cpp:52-92):
//   m = max_pa T (0 without parents); c = X_j ? T_s - m : +inf; z = -log1p(u expm1(-c lambda))/lambda;
//   T_j = (X_j ? m : max(m, T_s)) + z; Z_j = T_j - m; log q += (log lambda - lambda z) - log(1 - exp(-lambda c))
// plus cbn_density_log (:206-216) and the incompatibility flag any_j Z_j < 0 (src/mcem_hcbn.cpp:524).
__global__ void conditional_from_uniforms_kernel(const int* __restrict__ X, const double* __restrict__ U,
                                                 const double* __restrict__ Ts, int L, const DevPoset* pos,
                                                 const double* __restrict__ lambda, double* __restrict__ Tdiff,
                                                 double* __restrict__ log_q, double* __restrict__ log_pz,
                                                 int* __restrict__ incompatible, const SeamRng rng) {
  const int p = pos->p;
  double sumlog = 0.0;
  for (int j = 0; j < p; ++j) sumlog += log(lambda[j]);  // event-id order, as lambda.array().log().sum()
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
    double T[64], Zev[64];
    double ts;
    if (!rng.enabled) {
      ts = Ts[l];
    } else if (rng.draw_ts) {
      const uint4 r = philox4x32_10((uint32_t)l, 0u, 64u, rng.cw, rng.k0, rng.k1);
      ts = -log(u01_open0_d(r.x, r.y)) / rng.lambda_s;
    } else {
      ts = rng.time;
    }
    if (rng.enabled && rng.Ts_out) rng.Ts_out[l] = ts;
    double dens = 0.0;
    int inc = 0;
    for (int k = 0; k < p; ++k) {
      const int ev = pos->ev[k];
      const double rate = lambda[ev];
      double m = 0.0;
      bool first = true;
      for (uint64_t pm = pos->parent_pos[k]; pm; pm &= pm - 1) {
        const double tu = T[__ffsll((long long)pm) - 1];
        m = first ? tu : fmax(m, tu);  // rowwise().maxCoeff() over the parents (not clamped at 0)
        first = false;
      }
      bool xj;
      double u;
      if (!rng.enabled) {
        xj = X[(size_t)ev * L + l] != 0;
        u = U[(size_t)ev * L + l];
      } else {
        xj = (rng.x_word >> ev) & 1ull;
        const uint4 r = philox4x32_10((uint32_t)l, 0u, (uint32_t)(ev >> 1), rng.cw, rng.k0, rng.k1);
        u = (ev & 1) ? u01_open0_d(r.z, r.w) : u01_open0_d(r.x, r.y);
      }
      const double cutoff = xj ? ts - m : INFINITY;
      const double z = -log1p(u * expm1(-cutoff * rate)) / rate;
      const double base = xj ? m : fmax(m, ts);
      const double tsum = base + z;
      T[k] = tsum;
      const double zd = tsum - m;
      Zev[ev] = zd;
      Tdiff[(size_t)ev * L + l] = zd;
      inc |= zd < 0.0;
      dens += (log(rate) - rate * z) - log(1.0 - exp(-rate * cutoff));
    }
    double dot = 0.0;
    for (int j = 0; j < p; ++j) dot += Zev[j] * lambda[j];
    log_q[l] = dens;
    log_pz[l] = sumlog - dot;
    if (incompatible) incompatible[l] = inc;
  }
}

}  // namespace mccbn
