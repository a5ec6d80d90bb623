// mccbn_b200.cu -- libmccbn_b200.so: device-resident problem, EM driver and the C ABI (include/mccbn_b200.h).
//
// Host flow of one MCEM run (replaces MCEM_hcbn, src/mcem_hcbn.cpp:571-736):
//   upload + bit-pack observations once  ->  for each EM iteration, all stream-ordered with no host sync:
//     E-step kernel (Philox sampling + weights + per-(observation, chunk) partial sums)
//     finalize kernel (per-observation log-sum-exp combine, E[d], E[Z], first-stage reduction)
//     reduce kernel   (fixed-order second stage -> stats vector of p+4 doubles)
//     [ncclAllReduce of the stats vector when observations are sharded over GPUs]
//     M-step kernel   (lambda/eps update with the reference's clamps, trace row, tables for the next E-step)
//     device-to-device refresh of the __constant__ model tables
//   the host synchronises once per `update_step_size` iterations to apply the reference's batch-average /
//   convergence rule on the trace rows.
#include <algorithm>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <array>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cfloat>
#include <mutex>
#include <sstream>
#include <type_traits>

#include "engine.cuh"
#include "jit.cuh"
#include "kernels_cond.cuh"
#include "kernels_forward.cuh"
#include "kernels_fp64.cuh"
#include "kernels_pool.cuh"
#include "kernels_reduce.cuh"

using namespace mccbn;

namespace mccbn {

enum Scheme { kForward = 0, kAddRemove = 1, kBackward = 2, kBernoulli = 3, kPool = 4, kOtcbn = 5 };

static Scheme parse_scheme(const char* s) {
  if (!s) throw Error(MCCBN_ERR_ARG, "sampling scheme is NULL");
  const std::string v(s);
  if (v == "forward") return kForward;
  if (v == "add-remove") return kAddRemove;
  if (v == "backward") return kBackward;
  if (v == "bernoulli") return kBernoulli;
  if (v == "pool") return kPool;
  throw Error(MCCBN_ERR_ARG, "unknown sampling scheme: " + v);
}

static void flatten_or_throw(const int* poset, int p, Poset& P, FlatPoset& f) {
  if (p < 1 || p > kMaxP) throw Error(MCCBN_ERR_ARG, "p must be in [1, 64]");
  P = Poset::from_adjacency(poset, p);
  for (int v = 0; v < p; ++v)
    if (P.has_edge(v, v)) throw Error(MCCBN_ERR_NOT_ACYCLIC, kMsgNotAcyclic);
  if (!P.sort()) throw Error(MCCBN_ERR_NOT_ACYCLIC, kMsgNotAcyclic);
  P.flatten(f);
}

static void fill_dev_poset(const FlatPoset& f, DevPoset& d) {
  std::memset(&d, 0, sizeof(d));
  d.p = f.p;
  d.n_edges = f.n_edges;
  int e = 0;
  for (int k = 0; k < f.p; ++k) {
    d.ev[k] = (unsigned char)f.topo[k];
    d.parent_pos[k] = f.parent_pos[k];
    if (f.p <= kFastMaxP) {
      d.pstart[k] = (unsigned short)e;
      for (uint64_t m = f.parent_pos[k]; m; m &= m - 1) d.ppos[e++] = (unsigned char)ctz64(m);
    }
  }
  if (f.p <= kFastMaxP)
    for (int k = f.p; k < 68; ++k) d.pstart[k] = (unsigned short)e;
  for (int j = 0; j < f.p; ++j) {
    d.parent_ev[j] = f.parent_ev[j];
    d.ancestor_ev[j] = f.ancestor_ev[j];
  }
  d.has_children_pos = f.has_children_pos;
}

}  // namespace mccbn

// ================================================================================================================
// device-resident problem
// ================================================================================================================
namespace mccbn {

// kernel-parameter program of a poset for a given unroll width and staging element size
template <typename Prog>
static Prog build_prog_t(const FlatPoset& f, int pmax, int elem_bytes) {
  typedef typename Prog::word_t word_t;
  constexpr int kMaxNodes = (int)(sizeof(word_t) * 8);
  Prog g;
  std::memset(&g, 0, sizeof(g));
  g.p = f.p;
  g.store_mask = (word_t)f.has_children_pos;
  g.valid_mask = f.p >= kMaxNodes ? ~(word_t)0 : (((word_t)1 << f.p) - 1);
  g.one_bits = 0x3f800000u;
  const unsigned dummy = (unsigned)pmax * kFwdBlock * elem_bytes;
  int x = 0;
  for (int k = 0; k < kMaxNodes; ++k) {
    g.off[k] = make_uint2(dummy, dummy);
    g.xo[k] = make_uint4(dummy, dummy, dummy, dummy);
    g.xstart[k] = (unsigned short)x;
    if (k >= f.p) continue;
    g.ev[k] = (unsigned char)f.topo[k];
    g.parent_mask[k] = (decltype(g.parent_mask[k] + 0))f.parent_pos[k];
    int n = 0;
    for (uint64_t m = f.parent_pos[k]; m; m &= m - 1, ++n) {
      const unsigned row = (unsigned)ctz64(m);
      const unsigned o = row * kFwdBlock * elem_bytes;
      switch (n) {
        case 0: g.off[k].x = o; break;
        case 1: g.off[k].y = o; break;
        case 2: g.xo[k].x = o; break;
        case 3: g.xo[k].y = o; break;
        case 4: g.xo[k].z = o; break;
        case 5: g.xo[k].w = o; break;
        default:
          if (x >= (int)(sizeof(g.xoff) / sizeof(g.xoff[0]))) throw Error(MCCBN_ERR_ARG, "poset has too many edges");
          g.xoff[x++] = (unsigned short)(row * kFwdBlock * 4);
          break;
      }
    }
    if (n > 2) g.multi_mask |= (word_t)1 << k;
    if (n > 6) g.huge_mask |= (word_t)1 << k;
  }
  g.xstart[kMaxNodes] = g.xstart[kMaxNodes + 1] = (unsigned short)x;
  for (int k = f.p; k < kMaxNodes; ++k) g.xstart[k] = (unsigned short)x;
  return g;
}
static PosetProg build_prog(const FlatPoset& f, int pmax, int elem_bytes) {
  return build_prog_t<PosetProg>(f, pmax, elem_bytes);
}

}  // namespace mccbn

namespace mccbn {
template <typename K>
static void optin_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    MCCBN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}
// cudaFuncSetAttribute is per device: opt in once per (kernel, device), not once per process
#define MCCBN_OPTIN_ONCE(bytes, ...)                                        \
  do {                                                                      \
    static std::atomic<unsigned long long> done__{0};                       \
    const unsigned long long bit__ = 1ull << (ctx->device & 63);            \
    if (!(done__.load() & bit__)) {                                         \
      optin_smem(__VA_ARGS__, bytes);                                       \
      done__.fetch_or(bit__);                                               \
    }                                                                       \
  } while (0)
}  // namespace mccbn

struct mccbn_problem {
  mccbn_problem() = default;
  mccbn_problem(const mccbn_problem&) = delete;  // owns device buffers and graph handles
  mccbn_problem& operator=(const mccbn_problem&) = delete;
  mccbn_ctx* ctx = nullptr;
  int N = 0, p = 0;
  int64_t obs_offset = 0, N_global = 0;
  DevBuf<uint64_t> bits;
  DevBuf<double> times, weights;
  DevBuf<DevPoset> posets;    // kMaxCand candidate slots
  DevBuf<DevModel> models;
  DevBuf<FwdScale> scales;
  DevBuf<double> part, blk, stats, trace;
  DevBuf<double> out_w, out_w2, out_dist, out_T, out_logw;
  DevBuf<int> out_leff;
  DevBuf<uint64_t> ball;  // backward scheme: flip masks of the Hamming ball
  DevBuf<float> pool_T, pool_Z;  // pool scheme: K x PMAX sorted event times / waiting times
  DevBuf<uint32_t> pool_M;       // pool scheme: K x (PMAX + 1) prefix genotypes
  DevBuf<int> pool_dref;         // pool_resample: reference distance per observation
  DevBuf<double> pool_qtot;      // pool_resample: fp64 total weight per observation
  bool pool_resample = false;    // mccbn_em_options.pool_resample
  DevBuf<unsigned long long> counts;  // compatibility / initialize_lambda counters
  int ball_p = -1, ball_k = -1;
  int trace_rows = 0;
  FlatPoset flat[kMaxCand];
  int iter_host[kMaxCand];
  uint32_t eval_id[kMaxCand];
  // candidate-parallel annealing: every rank holds ALL observations and scores its share of the candidate posets, so the
  // EM iterations of this problem must not exchange statistics with the other ranks
  bool local_only = false;
  std::vector<uint64_t> host_bits;  // annealer, moderate N: bit-packed observations on the host (compatibility counts)
  bool test_structural_score = false;  // set by mccbn_test_asa_structural_score only (engine_asa.cuh)
  bool sharded() const { return ctx->world > 1 && !local_only; }
  unsigned long long xchg_seq[kMaxCand] = {};  // exchange number of the candidate's pending reduce_push
  // timeout: sharded EM iterations run in lock-step (30 s covers a first-call NVRTC compile on one rank); the annealer's score exchange waits for the
  // slowest rank's whole MCEM fits (10 min).  MCCBN_PEER_TIMEOUT_S overrides both.
  PeerView peer_view(unsigned long long seq, double timeout_s = 30.0) const {
    PeerView pv;
    std::memset(&pv, 0, sizeof(pv));
    if (const char* e = std::getenv("MCCBN_PEER_TIMEOUT_S")) timeout_s = std::max(1e-3, std::atof(e));
    pv.timeout_clocks = (long long)(timeout_s * 2.0e9);
    for (int q = 0; q < ctx->world && q < kMaxWorld; ++q) pv.box[q] = ctx->peer_box[q];
    pv.rank = ctx->rank;
    pv.world = ctx->world;
    pv.seq = seq;
    return pv;
  }
  int fin_blocks = 0;
  int expected_esteps = 1;  // hint for the JIT policy (set by the EM driver)

  int PS() const { return kStatHdr + p; }
  double* stats_of(int c) { return stats.ptr + (size_t)c * kMaxPS; }
  double* trace_of(int c) { return trace.ptr + (size_t)c * trace_rows * (p + 2); }

  // `ld`: leading dimension of the column-major host matrix `obs` (rows of the full matrix when obs points at a block of
  // rows of it, mccbn_em_options.obs_ld); 0 = N
  DevBuf<int> obs_unpacked;
  bool upload_inflight = false;
  void release_upload_staging() { obs_unpacked.release(); }  // only after the stream has been drained
  void upload(const int* obs, const double* t, const double* w, size_t ld = 0) {
    NvtxRange nvtx("upload+pack observations");
    cudaStream_t s = ctx->stream;
    // the unpacked matrix stays allocated until the problem is destroyed (or release_upload_staging after a sync): its
    // block must not return to the per-device cache -- which other contexts allocate from -- while the copy and the
    // pack kernel are in flight
    upload_inflight = true;
    DevBuf<int>& tmp = obs_unpacked;
    tmp.ensure((size_t)N * p);
    if (ld == 0 || ld == (size_t)N)
      MCCBN_CUDA(cudaMemcpyAsync(tmp.ptr, obs, sizeof(int) * (size_t)N * p, cudaMemcpyHostToDevice, s));
    else  // p columns of N ints each, ld ints apart on the host
      MCCBN_CUDA(cudaMemcpy2DAsync(tmp.ptr, sizeof(int) * (size_t)N, obs, sizeof(int) * ld, sizeof(int) * (size_t)N, (size_t)p,
                                   cudaMemcpyHostToDevice, s));
    bits.ensure(N);
    pack_obs_kernel<<<std::min(1024, (N + 255) / 256), 256, 0, s>>>(tmp.ptr, bits.ptr, N, p);
    ctx->launches++;
    times.ensure(N);
    weights.ensure(N);
    // caller-owned buffers stay valid until the entry that called upload() returns, and every entry drains the stream
    // before it returns: no sync here, the E-step is enqueued while the observations are still in flight.  Absent
    // vectors (times: 0, weights: 1) are filled on the device.
    if (t) MCCBN_CUDA(cudaMemcpyAsync(times.ptr, t, sizeof(double) * N, cudaMemcpyHostToDevice, s));
    else MCCBN_CUDA(cudaMemsetAsync(times.ptr, 0, sizeof(double) * N, s));
    if (w) {
      MCCBN_CUDA(cudaMemcpyAsync(weights.ptr, w, sizeof(double) * N, cudaMemcpyHostToDevice, s));
    } else {
      fill_kernel<<<std::min(1024, (N + 255) / 256), 256, 0, s>>>(weights.ptr, N, 1.0);
      ctx->launches++;
    }
    posets.ensure(kMaxCand);
    models.ensure(kMaxCand);
    scales.ensure(kMaxCand);
    stats.ensure((size_t)kMaxCand * kMaxPS);
    for (int c = 0; c < kMaxCand; ++c) {
      iter_host[c] = 0;
      eval_id[c] = (uint32_t)c;
    }
  }

  // set poset + parameters of candidate slot `c` and derive its device tables
  void set_model(int c, const FlatPoset& f, const double* lambda, double eps, float lambda_s, float max_lambda) {
    cudaStream_t s = ctx->stream;
    flat[c] = f;
    DevPoset dp;
    fill_dev_poset(f, dp);
    DevModel dm;
    std::memset(&dm, 0, sizeof(dm));
    for (int j = 0; j < p; ++j) dm.lambda[j] = lambda[j];
    dm.eps = eps;
    dm.lambda_s = (double)lambda_s;
    dm.max_lambda = (double)max_lambda;
    static_assert(sizeof(DevPoset) + sizeof(DevModel) <= mccbn_ctx::kStageBytes, "staging slot too small");
    int slot = 0;
    if (unsigned char* st = ctx->stage_slot(slot)) {  // pinned staging: asynchronous, no drain of the stream
      std::memcpy(st, &dp, sizeof(dp));
      std::memcpy(st + sizeof(dp), &dm, sizeof(dm));
      MCCBN_CUDA(cudaMemcpyAsync(posets.ptr + c, st, sizeof(dp), cudaMemcpyHostToDevice, s));
      MCCBN_CUDA(cudaMemcpyAsync(models.ptr + c, st + sizeof(dp), sizeof(dm), cudaMemcpyHostToDevice, s));
      MCCBN_CUDA(cudaEventRecord(ctx->stage_ev[slot], s));
    } else {
      MCCBN_CUDA(cudaMemcpyAsync(posets.ptr + c, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
      MCCBN_CUDA(cudaMemcpyAsync(models.ptr + c, &dm, sizeof(dm), cudaMemcpyHostToDevice, s));
      MCCBN_CUDA(cudaStreamSynchronize(s));  // dp/dm are stack objects
    }
    iter_host[c] = 0;
    model_version[c]++;
    prepare_kernel<<<1, 128, 0, s>>>(posets.ptr + c, models.ptr + c, scales.ptr + c);
    ctx->launches++;
    MCCBN_CUDA(cudaGetLastError());
  }
  void ensure_trace(int rows) {
    if (rows > trace_rows || !trace.ptr) {
      trace.ensure((size_t)kMaxCand * rows * (p + 2));
      trace_rows = rows;
    }
  }

  // Poset-specialised kernel (jit.cuh) or nullptr = generic kernel.  Policy of the automatic mode (jit_mode < 0 and no
  // MCCBN_JIT in the environment), by the sampling work the caller announces (samples per E-step x expected E-steps):
  //   kernel already in this context's memory cache or in the on-disk cache      -> use it
  //   announced work >= 5e9 samples (a compilation, ~0.3 s, costs < 10 % of it)  -> compile now
  //   >= 5e7 samples per E-step                                                  -> compile on a background thread, run the
  //                                                                                 generic kernel until it is ready
  //   smaller                                                                    -> generic kernel
  // Both kernels draw the same stream and write identical records (the chunking does not depend on the kernel), so
  // the result of a run does not depend on when -- or whether -- the specialised kernel becomes available.
  int jit_mode_effective() const {
    int mode = ctx->jit_mode;
    if (mode < 0) {
      const char* e = std::getenv("MCCBN_JIT");
      if (e && std::strcmp(e, "0") == 0) mode = 0;
      else if (e && std::strcmp(e, "1") == 0) mode = 1;
    }
    return mode;
  }
  JitCache::Policy jit_policy(int mode, double samples_per_estep) const {
    if (mode > 0) return JitCache::kCompileNow;
    if (samples_per_estep * (double)std::max(1, expected_esteps) >= 5e9) return JitCache::kCompileNow;
    return samples_per_estep >= 5e7 ? JitCache::kBackground : JitCache::kCachedOnly;
  }
  JitKernel* resolve_jit(const FlatPoset& f, bool times_given, int L) {
    const int mode = jit_mode_effective();
    if (mode == 0) return nullptr;
    if (!ctx->jit) ctx->jit = new JitCache();
    return ctx->jit->get(f, times_given, jit_policy(mode, (double)N * (double)L));
  }

  // is the specialised kernel that an E-step of candidate c would use loaded?
  bool jit_ready(int c, int scheme, bool times_given) const;
  // start compiling the specialised kernel of a poset on a background thread (annealer look-ahead); never blocks
  void prefetch_kernel(const Poset& P, int scheme /* Scheme */, bool times_given);

  // the same policy for the conditional-proposal kernels (backward / bernoulli / OT-CBN / add-remove)
  JitKernel* resolve_jit_cond(const FlatPoset& f, bool times_given, int mode, double samples_per_estep,
                              bool required = false) {
    int jm = jit_mode_effective();
    if (required) jm = 1;
    if (jm == 0) return nullptr;
    if (!ctx->jit) ctx->jit = new JitCache();
    return ctx->jit->get_cond(f, times_given, mode, jit_policy(jm, samples_per_estep));
  }

  void launch_forward_jit(JitKernel* jk, const FwdArgs& a, const FlatPoset& f, int c, int grid) {
    cudaStream_t s = ctx->stream;
    MCCBN_CUDA(cudaMemcpyAsync(jk->scale_sym, scales.ptr + c, sizeof(FwdScale), cudaMemcpyDeviceToDevice, s));
    FwdArgs args = a;
    if (f.p > kFastMaxP) {
      PosetProgWide pg = build_prog_t<PosetProgWide>(f, jk->pmax, 4);
      void* params[] = {(void*)&args, (void*)&pg};
      MCCBN_CUDA(cudaLaunchKernel((const void*)jk->kern, dim3(grid), dim3(kFwdBlock), params, jk->smem, s));
    } else {
      PosetProg pg = build_prog(f, jk->pmax, 4);
      void* params[] = {(void*)&args, (void*)&pg};
      MCCBN_CUDA(cudaLaunchKernel((const void*)jk->kern, dim3(grid), dim3(kFwdBlock), params, jk->smem, s));
    }
    ctx->launches++;
    ctx->jit_launches++;
  }

  template <typename real>
  void launch_forward(const FwdArgs& a, const FlatPoset& f, bool times_given, int grid, int c = 0) {
    cudaStream_t s = ctx->stream;
    int max_indeg = 0;  // <= 3: the lean variant of the generic kernel (kernels_forward.cuh: parents_max)
    for (int k = 0; k < f.p; ++k) max_indeg = std::max(max_indeg, (int)__builtin_popcountll(f.parent_pos[k]));
    static const bool lean_ok = [] {  // MCCBN_FWD_LEAN=0: always the general path (A/B runs; same results)
      const char* e = std::getenv("MCCBN_FWD_LEAN");
      return !(e && std::strcmp(e, "0") == 0);
    }();
#define MCCBN_FWD(PM)                                                                                   \
  do {                                                                                                  \
    const PosetProg pg = build_prog(f, PM, (int)sizeof(real));                                          \
    constexpr size_t sm = fwd_smem_bytes<real, PM>();                                                   \
    constexpr int kLean3 = sizeof(real) == 4 ? 3 : 0;                                                   \
    if (kLean3 && lean_ok && max_indeg <= 3 && !times_given) {                                          \
      MCCBN_OPTIN_ONCE(sm, estep_forward_kernel<real, PM, false, kLean3>);                              \
      estep_forward_kernel<real, PM, false, kLean3><<<grid, kFwdBlock, sm, s>>>(a, pg);                 \
    } else if (times_given) {                                                                           \
      MCCBN_OPTIN_ONCE(sm, estep_forward_kernel<real, PM, true>);                  \
      estep_forward_kernel<real, PM, true><<<grid, kFwdBlock, sm, s>>>(a, pg);                          \
    } else {                                                                                            \
      MCCBN_OPTIN_ONCE(sm, estep_forward_kernel<real, PM, false>);                 \
      estep_forward_kernel<real, PM, false><<<grid, kFwdBlock, sm, s>>>(a, pg);                         \
    }                                                                                                   \
  } while (0)
    if constexpr (sizeof(real) == 8) {
      if (p > kFastMaxP) throw Error(MCCBN_ERR_ARG, "the fp64 validation mode supports p <= 32");
      MCCBN_FWD(32);
    } else if (p > kFastMaxP) {
      // 32 < p <= 64: 64-bit genotype words, generic kernel
#define MCCBN_FWD_WIDE(PM)                                                                                  \
  do {                                                                                                      \
    const PosetProgWide pg = build_prog_t<PosetProgWide>(f, PM, 4);                                         \
    constexpr size_t sm = fwd_smem_bytes<float, PM>();                                                      \
    if (times_given) {                                                                                      \
      MCCBN_OPTIN_ONCE(sm, estep_forward_wide_kernel<PM, true>);                       \
      estep_forward_wide_kernel<PM, true><<<grid, kFwdBlock, sm, s>>>(a, pg);                               \
    } else {                                                                                                \
      MCCBN_OPTIN_ONCE(sm, estep_forward_wide_kernel<PM, false>);                      \
      estep_forward_wide_kernel<PM, false><<<grid, kFwdBlock, sm, s>>>(a, pg);                              \
    }                                                                                                       \
  } while (0)
      if (p <= 48) MCCBN_FWD_WIDE(48);
      else MCCBN_FWD_WIDE(64);
#undef MCCBN_FWD_WIDE
    } else {
      if (p <= 4) MCCBN_FWD(4);
      else if (p <= 8) MCCBN_FWD(8);
      else if (p <= 12) MCCBN_FWD(12);
      else if (p <= 16) MCCBN_FWD(16);
      else if (p <= 20) MCCBN_FWD(20);
      else if (p <= 24) MCCBN_FWD(24);
      else if (p <= 28) MCCBN_FWD(28);
      else MCCBN_FWD(32);
    }
#undef MCCBN_FWD
    ctx->launches++;
    MCCBN_CUDA(cudaGetLastError());
  }

  // Chunking of the L samples of one observation into records (one warp per record).  The partition fixes which samples
  // share a per-lane fp32 partial sum, i.e. the rounding of a record, so it must not depend on anything a caller may
  // vary without expecting the result to change: it is a function of (L, GLOBAL number of observations, #SMs) only --
  // not of the kernel that runs (generic / specialised) and not of how the observations are sharded over ranks.
  // Results are therefore identical, up to the fp64 summation order of the final statistic vector (~1e-15), for every
  // --gpus N.  Every record costs an epilogue (fp64 butterfly reduction, ~500 instructions) and restarts the running
  // reference distance, so records are as long as load balance allows (measured on B200, r2h: cfg 3 with 1 / 2 / 3 / 5 /
  // 8 / 12 records per observation = 19.67 / 19.71 / 19.82 / 20.09 / 20.53 / 21.06 ms; cfg 4, L = 1000: 1 / 2 / 3 = 0.184 /
  // 0.198 / 0.210 ms):
  //   * the kernel is issue-bound, so when a warp runs out of items the other warps of its scheduler speed up: balance
  //     is a matter of items per SCHEDULER, not per warp -- 16 per scheduler on each of up to 8 ranks (kShardRanks) keeps
  //     the quantisation loss at a few per cent;
  //   * a record is at least 1024 samples (32 sample iterations: the epilogue stays below ~5 % of it).
  // BASELINE configs 1-4 get one record per observation.  The GRID is one resident wave of the kernel that runs
  // (`blocks_per_sm`; 0: the generic forward kernel's launch bounds); its warps claim the items dynamically.
  static constexpr long long kShardRanks = 8;
  long long want_items_global() const {
    return (long long)ctx->sm_count * 4 /* schedulers */ * 16 * kShardRanks;
  }
  void choose_chunks(int L, int& chunks, int& chunk_len, int& grid_x, int blocks_per_sm = 0) const {
    const int warps_per_block = kFwdBlock / 32;
    if (blocks_per_sm <= 0) blocks_per_sm = p > kFastMaxP ? kFwdWideMinB : (p > MCCBN_FWD_BIGP - 4 ? MCCBN_FWD_MINB_BIG : 6);
    const int resident_blocks = ctx->sm_count * blocks_per_sm;  // blocks of 4 warps
    const long long Ng = std::max<long long>(1, N_global > 0 ? N_global : N);
    long long c = (want_items_global() + Ng - 1) / Ng;
    long long cmax = std::max(1, L / 1024);
    c = std::max(1LL, std::min(c, cmax));
#ifdef MCCBN_DIAG_BUILD  // tools/ profiling builds only
    if (const char* e = std::getenv("MCCBN_DIAG_CHUNKS")) c = std::max(1LL, std::min<long long>(std::atoll(e), std::max(1, L / 32)));
#endif
    chunk_len = (int)(((L + c - 1) / c + 31) / 32 * 32);
    chunks = (L + chunk_len - 1) / chunk_len;
    long long items = (long long)N * chunks;
    long long blocks = (items + warps_per_block - 1) / warps_per_block;
    grid_x = (int)std::max(1LL, std::min(blocks, (long long)resident_blocks));
  }

  // Enqueue one E-step (+ reduction, allreduce, M-step) for candidate slot c.  No host sync.
  void enqueue_iteration(int c, Scheme scheme, int L, unsigned neighborhood_dist, bool times_given, uint32_t seed,
                         int update_params, bool per_obs_out, int precision);

  // n identical iterations.  After the first (eager) one, the launch sequence of one iteration -- scale binding,
  // E-step, finalize, reduction, M-step -- is captured into a CUDA graph and replayed: everything that changes from
  // one iteration to the next (parameters, tables, the iteration field of the Philox key, the trace row) lives in
  // device memory.  Small problems are launch-bound: 29 -> 13 us per EM iteration on BASELINE config 1.
  struct IterGraph {
    cudaGraphExec_t exec = nullptr;
    std::string key;
    uint64_t launches_per_iter = 0, jit_per_iter = 0;
  };
  IterGraph iter_graph[kMaxCand];
  unsigned model_version[kMaxCand] = {};
  ~mccbn_problem() {
    // uploads are asynchronous: nothing may still read the caller's buffers (or the staging block that goes back to the
    // device cache with this object) when the entry that owns this problem returns -- on the error path either
    if (upload_inflight && ctx) cudaStreamSynchronize(ctx->stream);
    for (IterGraph& g : iter_graph)
      if (g.exec) cudaGraphExecDestroy(g.exec);
  }
  void enqueue_iterations(int c, int n, Scheme scheme, int L, unsigned neighborhood_dist, bool times_given,
                          uint32_t seed, int update_params, int precision) {
    NvtxRange nvtx("enqueue EM iterations");
    static const bool enabled = [] {
      const char* e = std::getenv("MCCBN_GRAPHS");
      return !(e && std::strcmp(e, "0") == 0);
    }();
    int done = 0;
    if (enabled && !sharded() && n >= 3) {
      enqueue_iteration(c, scheme, L, neighborhood_dist, times_given, seed, update_params, false, precision);  // eager
      ++done;
      cudaStream_t s = ctx->stream;
      auto graph_key = [&] {  // everything a captured iteration depends on besides device memory
        std::ostringstream k;
        k << (int)scheme << ':' << L << ':' << neighborhood_dist << ':' << times_given << ':' << seed << ':' << update_params
          << ':' << precision << ':' << pool_resample << ':' << N << ':' << p << ':' << model_version[c] << ':' << eval_id[c]
          << ':' << trace_rows << ':' << device_cache().epoch_of_current_device() << ':' << (void*)s << ':' << ctx->jit_mode
          << ':' << jit_ready(c, scheme, times_given);  // a finished background compilation is picked up
        return k.str();
      };
      const std::string key_now = graph_key();
      IterGraph& g = iter_graph[c];
      if (!g.exec || g.key != key_now) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
        const uint64_t l0 = ctx->launches, j0 = ctx->jit_launches;
        const int it0 = iter_host[c];
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
          try {
            enqueue_iteration(c, scheme, L, neighborhood_dist, times_given, seed, update_params, false, precision);
          } catch (...) {
            cudaStreamEndCapture(s, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
          }
          e = cudaStreamEndCapture(s, &graph);
        }
        if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        g.launches_per_iter = ctx->launches - l0;
        g.jit_per_iter = ctx->jit_launches - j0;
        ctx->launches = l0;     // the captured iteration was recorded, not run
        ctx->jit_launches = j0;
        iter_host[c] = it0;
        if (e != cudaSuccess || !g.exec) {  // capture unavailable: fall back to eager launches
          cudaGetLastError();
          g.exec = nullptr;
          g.key.clear();
        } else {
          // allocation epoch as of AFTER the eager iteration and the capture (neither allocates in steady state)
          g.key = graph_key();
        }
      }
      if (g.exec) {
        for (; done < n; ++done) {
          MCCBN_CUDA(cudaGraphLaunch(g.exec, s));
          ctx->launches += g.launches_per_iter;
          ctx->jit_launches += g.jit_per_iter;
          iter_host[c]++;
        }
      }
    }
    for (; done < n; ++done)
      enqueue_iteration(c, scheme, L, neighborhood_dist, times_given, seed, update_params, false, precision);
  }

  // finalize + second-stage reduction + M-step; one fused launch on a single GPU (finalize_mstep_kernel)
  DevBuf<unsigned> fuse_counter;
  DevBuf<unsigned long long> item_counter;  // dynamic work distribution of the conditional kernels
  void run_finalize_mstep(int c, int chunks, long long n_items, int L_eff, bool per_obs_out,
                          int update_params) {
    static const bool fuse = [] {
      const char* e = std::getenv("MCCBN_FUSE_TAIL");
      return !(e && std::strcmp(e, "0") == 0);
    }();
    if (fuse && (!sharded() || ctx->peer_ready)) {  // the NCCL exchange needs a host-enqueued collective in between
      run_finalize(c, chunks, n_items, L_eff, per_obs_out, update_params, /*fused=*/true);
      iter_host[c]++;
    } else {
      run_finalize(c, chunks, n_items, L_eff, per_obs_out);
      run_mstep(c, update_params);
    }
  }

  // work counter of the E-step kernels: zeroed when it is allocated, re-armed by the tail kernel of every iteration
  unsigned long long* armed_item_counter() {
    if (!item_counter.ptr) {
      item_counter.ensure(1);
      MCCBN_CUDA(cudaMemsetAsync(item_counter.ptr, 0, sizeof(unsigned long long), ctx->stream));
    }
    return item_counter.ptr;
  }

  void run_finalize(int c, int chunks, long long n_items, int L_eff, bool per_obs_out,
                    int update_params = 0, bool fused = false) {
    cudaStream_t s = ctx->stream;
    FinArgs fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.part = part.ptr;
    fa.poset = posets.ptr + c;
    fa.weights = weights.ptr;
    fin_blocks = std::max(1, std::min((N + kFinWarps - 1) / kFinWarps, ctx->sm_count * 2));
    fa.blk_part = blk.ensure((size_t)fin_blocks * PS());
    if (per_obs_out) {
      fa.w_sum = out_w.ensure(N);
      fa.w_sum_sq = out_w2.ensure(N);
      fa.dist = out_dist.ensure(N);
      fa.Tdiff = out_T.ensure((size_t)N * p);
      fa.L_eff_out = out_leff.ensure(N);
      fa.log_w = out_logw.ensure(N);
    }
    fa.N = N;
    fa.chunks = chunks;
    fa.PS = PS();
    fa.p = p;
    fa.n_items = n_items;
    fa.L_eff = L_eff;
    if (fused) {
      if (!fuse_counter.ptr) {
        fuse_counter.ensure(1);
        MCCBN_CUDA(cudaMemsetAsync(fuse_counter.ptr, 0, sizeof(unsigned), s));
      }
      FusedTail ft;
      ft.stats = stats_of(c);
      ft.counter = fuse_counter.ptr;
      ft.rearm = item_counter.ptr;
      if (sharded()) {
        xchg_seq[c] = ++ctx->xchg_seq;
        ft.pv = peer_view(xchg_seq[c]);
      } else {
        ft.pv = peer_view(0);
        ft.pv.world = 1;
      }
      ft.pos = posets.ptr + c;
      ft.mdl = models.ptr + c;
      ft.scale = scales.ptr + c;
      ft.trace = trace.ptr ? trace_of(c) : nullptr;
      ft.trace_stride = p + 2;
      ft.max_trace_rows = trace_rows;
      ft.update_params = update_params;
      finalize_mstep_kernel<<<fin_blocks, kFinBlock, 0, s>>>(fa, ft);
      ctx->launches++;
      MCCBN_CUDA(cudaGetLastError());
      return;
    }
    finalize_kernel<<<fin_blocks, kFinBlock, 0, s>>>(fa);
    if (sharded() && ctx->peer_ready) {
      xchg_seq[c] = ++ctx->xchg_seq;
      reduce_push_kernel<<<1, 256, 0, s>>>(blk.ptr, stats_of(c), fin_blocks, PS(), peer_view(xchg_seq[c]));
    } else {
      reduce_stats_kernel<<<1, 256, 0, s>>>(blk.ptr, stats_of(c), fin_blocks, PS());
    }
    ctx->launches += 2;
    MCCBN_CUDA(cudaGetLastError());
  }
  void run_mstep(int c, int update_params) {
    cudaStream_t s = ctx->stream;
    PeerView pv = peer_view(0);
    if (sharded() && ctx->peer_ready) {
      pv = peer_view(xchg_seq[c]);  // the M-step kernel pulls the vectors pushed by reduce_push_kernel
    } else {
      pv.world = 1;
      if (sharded()) {
        if (!ctx->comm) throw Error(MCCBN_ERR_ARG, "world > 1 but neither the peer exchange nor NCCL is initialised");
        ctx->nccl.check(ctx->nccl.AllReduce(stats_of(c), stats_of(c), (size_t)PS(), NcclApi::kDouble, NcclApi::kSum,
                                            ctx->comm, s),
                        "ncclAllReduce");
      }
    }
    mstep_kernel<<<1, 64, 0, s>>>(stats_of(c), posets.ptr + c, models.ptr + c, scales.ptr + c,
                                  trace.ptr ? trace_of(c) : nullptr, p + 2, trace_rows, PS(), update_params, pv,
                                  item_counter.ptr);
    ctx->launches++;
    MCCBN_CUDA(cudaGetLastError());
    iter_host[c]++;
  }
  // element-wise sum over ranks of a host vector (every rank gets the rank-ordered sum): peer mailboxes or NCCL
  DevBuf<double> xbuf;
  DevBuf<int> xerr;
  void exchange_sum(std::vector<double>& v) {
    if (ctx->world <= 1 || v.empty()) return;
    cudaStream_t s = ctx->stream;
    const size_t n = v.size();
    double* d = xbuf.ensure(n);
    int* ef = xerr.ensure(1);
    MCCBN_CUDA(cudaMemcpyAsync(d, v.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
    MCCBN_CUDA(cudaMemsetAsync(ef, 0, sizeof(int), s));
    if (ctx->peer_ready) {
      for (size_t off = 0; off < n; off += kMaxPSlots) {
        const int m = (int)std::min<size_t>(kMaxPSlots, n - off);
        const PeerView pv = peer_view(++ctx->xchg_seq, 600.0);
        push_vec_kernel<<<1, 128, 0, s>>>(d + off, m, pv);
        pull_vec_kernel<<<1, 128, 0, s>>>(d + off, m, ef, pv);
        ctx->launches += 2;
      }
      MCCBN_CUDA(cudaGetLastError());
    } else if (ctx->comm) {
      ctx->nccl.check(ctx->nccl.AllReduce(d, d, n, NcclApi::kDouble, NcclApi::kSum, ctx->comm, s), "ncclAllReduce");
    } else {
      throw Error(MCCBN_ERR_ARG, "world > 1 but neither the peer exchange nor NCCL is initialised");
    }
    int herr = 0;
    MCCBN_CUDA(cudaMemcpyAsync(v.data(), d, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    MCCBN_CUDA(cudaMemcpyAsync(&herr, ef, sizeof(int), cudaMemcpyDeviceToHost, s));
    MCCBN_CUDA(cudaStreamSynchronize(s));
    if (herr) throw Error(MCCBN_ERR_CUDA, kMsgCommTimeout);
  }

  void read_model(int c, DevModel& dm) {
    MCCBN_CUDA(cudaMemcpyAsync(&dm, models.ptr + c, sizeof(dm), cudaMemcpyDeviceToHost, ctx->stream));
    MCCBN_CUDA(cudaStreamSynchronize(ctx->stream));
  }
};

#include "engine_impl.cuh"
