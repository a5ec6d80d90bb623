// kernels_reduce.cuh -- stages 5 and 6 of the hot path: per-observation finalisation, deterministic reduction of
// the expected sufficient statistics, on-device M-step, and the small indexing kernels (bit-packing,
// compatibility counts, initialize_lambda counts).
//
// Reference arithmetic being replaced:
//   per-observation (src/mcem_hcbn.cpp:683-694): aux = sum w; N_eff += wt_i; obs_llhood += wt_i log(aux / L_eff);
//       expected_dist += wt_i (sum w d)/aux; expected_Tdiff[i,] = (sum w Z)/aux
//   M-step (:707-709, src/model_impl.cpp:21-34): eps = clamp(expected_dist/(N_eff p)); lambda_j = 1/(S_j/N_eff), clamped
#pragma once
#include <cfloat>

#include "device_types.cuh"
#include "kernels_forward.cuh"

namespace mccbn {

constexpr int kFinBlock = 256;
constexpr int kFinWarps = kFinBlock / 32;
constexpr int kMaxPS = kStatHdr + 64;

struct FinArgs {
  const double* part;       // n_items x PS
  const DevPoset* poset;    // this candidate
  const double* weights;    // N
  double* blk_part;         // gridDim.x x PS
  double* w_sum;            // optional per-observation outputs
  double* w_sum_sq;
  double* dist;
  double* Tdiff;            // N x p column-major by event id
  int* L_eff_out;
  int N, chunks, PS, p;
  long long n_items;
  int L_eff;                // > 0: fixed L_eff (forward / bernoulli / pool: L); <= 0: #{w > 0} (backward, :688-689)
  double* log_w;            // optional: natural log of sum w per observation (no underflow; OT-CBN log-likelihood caller)
};

// Partial record (one per observation and chunk): [0]=sum w', [1]=sum w'^2, [2]=sum w' d, [3]=#{w>0},
// [4]=natural-log scale s (true weight w = exp(s) w'), [5+k]=sum w' Z_k (topological position k).
// Chunks are merged with a log-sum-exp over their scales.  Output stats: [0]=LL, [1]=N_eff, [2]=D,
// [3]=#observations whose weights are all zero, [4]=0, [5+k]=S_k.
// One warp per observation, lane s owns record slot s (and s+32): coalesced reads, no shuffles except broadcasts.
__device__ __forceinline__ void finalize_body(const FinArgs& a) {
  __shared__ double s_acc[kFinWarps][kMaxPS];
  const DevPoset* __restrict__ pos = a.poset;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int PS = a.PS;
  const int sA = lane, sB = lane + 32, sC = lane + 64;  // PS <= 69: lanes 0..4 also own a third slot
  double accA = 0.0, accB = 0.0, accC = 0.0;

  for (int i = blockIdx.x * kFinWarps + warp; i < a.N; i += gridDim.x * kFinWarps) {
    const double* rec = a.part + (size_t)i * a.chunks * PS;
    double M = rec[4];
    for (int c = 1; c < a.chunks; ++c) M = fmax(M, rec[(size_t)c * PS + 4]);
    const bool any = M > -INFINITY;
    double vA = 0.0, vB = 0.0, vC = 0.0;
    for (int c = 0; c < a.chunks; ++c) {
      const double* rc = rec + (size_t)c * PS;
      const double s = any ? exp(rc[4] - M) : 0.0;
      if (sA < PS) vA += (sA == 1 ? s * s : (sA == 3 ? 1.0 : s)) * rc[sA];
      if (sB < PS) vB += s * rc[sB];
      if (sC < PS) vC += s * rc[sC];
    }
    const double W = __shfl_sync(0xffffffffu, vA, 0);
    const double npos = __shfl_sync(0xffffffffu, vA, 3);
    const double wt = a.weights[i];
    const double logW = any ? log(W) + M : -INFINITY;
    const double w_true = exp(logW);
    const bool ok = w_true > 0.0;  // the reference aborts when aux == 0 (src/mcem_hcbn.cpp:695-699)
    const double Leff = a.L_eff > 0 ? (double)a.L_eff : npos;
    if (lane == 0) {
      accA += ok ? wt * (logW - log(Leff)) : 0.0;
      if (a.w_sum) a.w_sum[i] = w_true;
      if (a.log_w) a.log_w[i] = logW;
      if (a.L_eff_out) a.L_eff_out[i] = (int)npos;
    } else if (lane == 1) {
      accA += ok ? wt : 0.0;
      if (a.w_sum_sq) a.w_sum_sq[i] = any ? vA * exp(2.0 * M) : 0.0;
    } else if (lane == 2) {
      const double ed = vA / W;
      accA += ok ? wt * ed : 0.0;
      if (a.dist) a.dist[i] = ed;
    } else if (lane == 3) {
      accA += ok ? 0.0 : 1.0;
    } else if (lane > 4 && sA < PS) {
      const double ez = vA / W;
      accA += ok ? wt * ez : 0.0;
      if (a.Tdiff) a.Tdiff[(size_t)pos->ev[sA - kStatHdr] * a.N + i] = ez;
    }
    if (sB < PS) {
      const double ez = vB / W;
      accB += ok ? wt * ez : 0.0;
      if (a.Tdiff) a.Tdiff[(size_t)pos->ev[sB - kStatHdr] * a.N + i] = ez;
    }
    if (sC < PS) {
      const double ez = vC / W;
      accC += ok ? wt * ez : 0.0;
      if (a.Tdiff) a.Tdiff[(size_t)pos->ev[sC - kStatHdr] * a.N + i] = ez;
    }
  }
  if (sA < PS) s_acc[warp][sA] = accA;
  if (sB < PS) s_acc[warp][sB] = accB;
  if (sC < PS) s_acc[warp][sC] = accC;
  __syncthreads();
  for (int s = threadIdx.x; s < PS; s += kFinBlock) {
    double t = 0.0;
    for (int w = 0; w < kFinWarps; ++w) t += s_acc[w][s];
    a.blk_part[(size_t)blockIdx.x * PS + s] = t;
  }
}

__global__ void __launch_bounds__(kFinBlock) finalize_kernel(const FinArgs a) { finalize_body(a); }

// second stage: sum of the block partials -> one value per slot, in a fixed order (lane l adds blocks l, l+32, ...;
// the 32 lane sums are combined by a butterfly): deterministic, and 10x shorter than one thread per slot walking all
// ~300 blocks.  Called by whole warps; returns the total in every lane.
__device__ __forceinline__ double slot_total(const double* __restrict__ blk_part, int n_blocks, int PS, int s) {
  const int lane = threadIdx.x & 31;
  double t = 0.0;
  for (int b = lane; b < n_blocks; b += 32) t += blk_part[(size_t)b * PS + s];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// second stage -> stats[PS]
__global__ void reduce_stats_kernel(const double* __restrict__ blk_part, double* __restrict__ stats, int n_blocks,
                                    int PS) {
  const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int s = warp; s < PS; s += n_warps) {
    const double t = slot_total(blk_part, n_blocks, PS, s);
    if ((threadIdx.x & 31) == 0) stats[s] = t;
  }
}

// ---- peer-memory statistic exchange (replaces ncclAllReduce on one NVLink domain) ---------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// second stage + push: fixed-order sum of the block partials, stored into slot[seq & 1][rank] of every rank's mailbox
// (NVLink peer stores), then flag[rank] = seq with release semantics at system scope.  One block.
__global__ void reduce_push_kernel(const double* __restrict__ blk_part, double* __restrict__ stats, int n_blocks, int PS,
                                   const PeerView pv) {
  const int par = (int)(pv.seq & 1ull);
  const int warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < PS; s += n_warps) {
    const double t = slot_total(blk_part, n_blocks, PS, s);
    if (lane == 0) stats[s] = t;  // the local (unreduced) vector stays visible; mstep overwrites it with the sum
    if (lane < pv.world) pv.box[lane]->slot[par][pv.rank][s] = t;  // lane q stores into rank q's mailbox
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pv.world) st_release_sys(&pv.box[threadIdx.x]->flag[pv.rank], pv.seq);
}

// pull: wait until every rank has published exchange `seq`, then sum the slots in rank order into st_out[PS]
// (shared or global).  Returns false (for all threads) after pv.timeout_clocks without progress: a rank died.
__device__ inline bool peer_pull(const PeerView& pv, int PS, double* st_out) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  Mailbox* mine = pv.box[pv.rank];
  if (threadIdx.x < pv.world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->flag[threadIdx.x]) < pv.seq) {
      if (clock64() - t0 > pv.timeout_clocks) {
        s_ok = 0;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const int par = (int)(pv.seq & 1ull);
  for (int s = threadIdx.x; s < PS; s += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < pv.world; ++r) t += *(volatile double*)&mine->slot[par][r][s];
    st_out[s] = t;
  }
  __syncthreads();
  return s_ok != 0;
}

// Generic sum-exchange of a short vector (n <= kMaxPSlots) over the mailboxes: every rank pushes, then pulls the
// rank-ordered sum back into `vec` (used by the annealer to share candidate scores between GPUs).  One block each.
__global__ void push_vec_kernel(const double* __restrict__ vec, int n, const PeerView pv) {
  const int par = (int)(pv.seq & 1ull);
  for (int s = threadIdx.x; s < n; s += blockDim.x) {
    const double t = vec[s];
    for (int q = 0; q < pv.world; ++q) pv.box[q]->slot[par][pv.rank][s] = t;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pv.world) st_release_sys(&pv.box[threadIdx.x]->flag[pv.rank], pv.seq);
}
__global__ void pull_vec_kernel(double* __restrict__ vec, int n, int* __restrict__ error_flag, const PeerView pv) {
  if (!peer_pull(pv, n, vec) && threadIdx.x == 0) *error_flag = 1;
}

// Derived tables for the next E-step (device function, run by one block).
__device__ inline void prepare_tables(const DevPoset* __restrict__ pos, DevModel* __restrict__ mdl,
                                      FwdScale* __restrict__ sc) {
  const int p = pos->p;
  const double eps = mdl->eps;
  const double le = log(eps), l1 = log1p(-eps);
  const int flip = (eps > 0.5) ? 1 : 0;
  const double rr = flip ? (1.0 - eps) / eps : eps / (1.0 - eps);
  for (int t = threadIdx.x; t < kRtabN; t += blockDim.x) {
    double v = (t <= p) ? pow(rr, (double)t) : 0.0;
    if (t == 0) v = 1.0;
    mdl->rtab_d[t] = v;
    mdl->rtab[t] = (float)v;
  }
  if (threadIdx.x == 0) {
    mdl->flip = flip;
    mdl->log_base = flip ? p * le : p * l1;
    mdl->log_r = flip ? (l1 - le) : (le - l1);
  }
  if (sc && p <= kWideMaxP) {
    const double kLn2 = 0.693147180559945309417;
    for (int k = threadIdx.x; k < kWideMaxP; k += blockDim.x)
      sc->c[k] = k < p ? (float)(-kLn2 / mdl->lambda[pos->ev[k]]) : 0.f;
    if (threadIdx.x == 0) {
      sc->cs = (float)(-kLn2 / mdl->lambda_s);
      sc->flip = flip;
    }
  }
}

__global__ void prepare_kernel(const DevPoset* poset, DevModel* model, FwdScale* scale) {
  prepare_tables(poset, model, scale);
}

// Stage 6: M-step + trace row + tables for the next iteration.  One block.
// stats may have been all-reduced across GPUs in between (identical on every rank => replicated M-step).
// With a PeerView (world > 1) the kernel first pulls the statistic vectors of all ranks out of its mailbox.
__device__ __forceinline__ void mstep_body(double* __restrict__ st, const DevPoset* pos, DevModel* mdl, FwdScale* scale,
                                           double* trace, int trace_stride, int max_trace_rows, int PS, int update_params,
                                           const PeerView& pv) {
  const int p = pos->p;
  if (pv.world > 1) {
    if (!peer_pull(pv, PS, st) && threadIdx.x == 0) mdl->comm_error = 1;
    __threadfence_block();
  }
  const double LL = st[0], Neff = st[1], D = st[2], zero = st[3];
  if (update_params) {
    for (int k = threadIdx.x; k < p; k += blockDim.x) {
      double lam = 1.0 / (st[kStatHdr + k] / Neff);
      if (lam > mdl->max_lambda || !isfinite(lam)) lam = mdl->max_lambda;  // Model::update_lambda
      mdl->lambda[pos->ev[k]] = lam;
    }
    if (threadIdx.x == 0) {
      const double e = D / (Neff * p);
      mdl->eps = fmax(fmin(e, 1.0 - DBL_EPSILON), DBL_EPSILON);  // Model::update_epsilon
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mdl->llhood = LL;
    if (zero > 0.0) mdl->zero_weight = 1;
  }
  const int it = mdl->iter;
  if (trace && it < max_trace_rows) {
    double* row = trace + (size_t)it * trace_stride;
    if (threadIdx.x == 0) {
      row[0] = LL;
      row[1] = mdl->eps;
    }
    for (int j = threadIdx.x; j < p; j += blockDim.x) row[2 + j] = mdl->lambda[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) mdl->iter = it + 1;
  if (update_params) prepare_tables(pos, mdl, scale);
}

__global__ void mstep_kernel(double* __restrict__ st, const DevPoset* pos, DevModel* mdl, FwdScale* scale,
                             double* trace, int trace_stride, int max_trace_rows, int PS, int update_params,
                             const PeerView pv, unsigned long long* rearm) {
  if (threadIdx.x == 0 && rearm) *rearm = 0ull;  // work counter of the E-step kernels, for the next E-step
  mstep_body(st, pos, mdl, scale, trace, trace_stride, max_trace_rows, PS, update_params, pv);
}

// Tail of an EM iteration in ONE launch: per-observation finalize in every block; the block that finishes last (ticket
// counter) sums the block partials, -- when the observations are sharded over ranks -- pushes the vector into every
// rank's mailbox, waits for the vectors of all ranks and sums them in rank order, and runs the M-step.  An EM iteration
// is two launches (E-step + tail) on one GPU and on several.
struct FusedTail {
  double* stats;
  unsigned* counter;  // zero before the first launch; re-armed by the last block
  unsigned long long* rearm;  // work counter of the E-step kernels (FwdArgs / CondArgs::next_item): zeroed for the next one
  PeerView pv;        // world > 1: the last block also exchanges the statistic vector with the other ranks
  const DevPoset* pos;
  DevModel* mdl;
  FwdScale* scale;
  double* trace;
  int trace_stride, max_trace_rows, update_params;
};
__global__ void __launch_bounds__(kFinBlock) finalize_mstep_kernel(const FinArgs a, const FusedTail t) {
  finalize_body(a);
  __shared__ unsigned s_ticket;
  __threadfence();  // this block's partials are visible device-wide before it takes its ticket
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(t.counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  if (threadIdx.x == 0) {
    *t.counter = 0u;
    if (t.rearm) *t.rearm = 0ull;
  }
  __threadfence();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int par = (int)(t.pv.seq & 1ull);
  for (int s = warp; s < a.PS; s += kFinWarps) {  // same order as slot_total; L2 loads: other blocks wrote the data
    double v = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(a.blk_part + (size_t)b * a.PS + s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) t.stats[s] = v;
    // sharded over ranks: lane q stores this rank's value into rank q's mailbox (NVLink peer store / pinned host memory)
    if (t.pv.world > 1 && lane < t.pv.world) t.pv.box[lane]->slot[par][t.pv.rank][s] = v;
  }
  if (t.pv.world > 1) {  // publish: all slot stores of this block before the flags (release at system scope)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < t.pv.world) st_release_sys(&t.pv.box[threadIdx.x]->flag[t.pv.rank], t.pv.seq);
  }
  __syncthreads();
  // world > 1: mstep_body first waits for every rank's flag and sums the slots in rank order (peer_pull)
  mstep_body(t.stats, t.pos, t.mdl, t.scale, t.trace, t.trace_stride, t.max_trace_rows, a.PS, t.update_params, t.pv);
}

// ---- indexing kernels ---------------------------------------------------------------------------------------

__global__ void fill_kernel(double* __restrict__ v, int n, double x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = x;
}

// R's N x p int32 column-major observation matrix -> one 64-bit word per observation (bit j = event j)
__global__ void pack_obs_kernel(const int* __restrict__ obs, uint64_t* __restrict__ bits, int N, int p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    uint64_t w = 0;
    for (int j = 0; j < p; ++j) w |= (uint64_t)(obs[(size_t)j * N + i] != 0) << j;  // coalesced along i
    bits[i] = w;
  }
}

// is_compatible / num_compatible_observations / num_incompatible_events
// (src/compatible_observations.cpp:12-28, :68-76, :51-66) for n_cand posets at once.
// counts[cand*2 + 0] = #compatible observations, counts[cand*2 + 1] = sum over direct edges u->v of #{i: !obs_u & obs_v}
__global__ void compat_kernel(const uint64_t* __restrict__ bits, int N, const DevPoset* posets, int* flags /*N or null*/,
                              unsigned long long* counts) {
  const DevPoset* pos = posets + blockIdx.y;
  __shared__ uint64_t s_par[64];
  const int p = pos->p;
  for (int j = threadIdx.x; j < 64; j += blockDim.x) s_par[j] = j < p ? pos->parent_ev[j] : 0ull;
  __syncthreads();
  unsigned ncomp = 0, ninc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const uint64_t g = bits[i];
    unsigned bad = 0;
    for (uint64_t m = g; m; m &= m - 1) {
      const int j = __ffsll((long long)m) - 1;
      bad += __popcll(s_par[j] & ~g);
    }
    ninc += bad;
    ncomp += bad == 0;
    if (flags && blockIdx.y == 0) flags[i] = bad == 0;
  }
  // block reduction (integers: order-independent, exact)
  for (int o = 16; o > 0; o >>= 1) {
    ncomp += __shfl_xor_sync(0xffffffffu, ncomp, o);
    ninc += __shfl_xor_sync(0xffffffffu, ninc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[blockIdx.y * 2 + 0], (unsigned long long)ncomp);
    atomicAdd(&counts[blockIdx.y * 2 + 1], (unsigned long long)ninc);
  }
}

// counts for initialize_lambda (src/asa.cpp:72-129): per event j, with A_j = all ancestors of j:
//   cnt[cand][j][0] = #{i : A_j subset of obs_i},  cnt[cand][j][1] = #{i : A_j u {j} subset of obs_i}
__global__ void init_lambda_counts_kernel(const uint64_t* __restrict__ bits, int N, const DevPoset* posets,
                                          unsigned long long* cnt) {
  const DevPoset* pos = posets + blockIdx.y;
  const int p = pos->p;
  __shared__ uint64_t s_anc[64];
  __shared__ unsigned s_cnt[64][2];
  for (int j = threadIdx.x; j < 64; j += blockDim.x) {
    s_anc[j] = j < p ? pos->ancestor_ev[j] : 0ull;
    s_cnt[j][0] = s_cnt[j][1] = 0;
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const uint64_t g = bits[i];
    for (int j = 0; j < p; ++j) {
      if ((s_anc[j] & ~g) == 0) {
        atomicAdd(&s_cnt[j][0], 1u);
        if ((g >> j) & 1ull) atomicAdd(&s_cnt[j][1], 1u);
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < p; j += blockDim.x) {
    atomicAdd(&cnt[((size_t)blockIdx.y * 64 + j) * 2 + 0], (unsigned long long)s_cnt[j][0]);
    atomicAdd(&cnt[((size_t)blockIdx.y * 64 + j) * 2 + 1], (unsigned long long)s_cnt[j][1]);
  }
}

}  // namespace mccbn
