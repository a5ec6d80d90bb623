// engine_impl.cuh -- E-step dispatch, EM driver (MCEM_hcbn host logic) and the extern "C" entry points.
// Included at the end of mccbn_b200.cu (single translation unit: kernels, __constant__ tables and host code).
#pragma once

// ----------------------------------------------------------------------------------------------------------------
// E-step dispatch
// ----------------------------------------------------------------------------------------------------------------
namespace mccbn {

struct BackwardPlan {
  std::vector<uint64_t> masks;
  int n_ball = 0, reps = 0, L_eff_total = 0;
};

static BackwardPlan plan_backward(int p, unsigned neighborhood_dist, unsigned L) {
  BackwardPlan bp;
  bp.masks = hamming_ball_masks((unsigned)p, neighborhood_dist);
  bp.n_ball = (int)bp.masks.size();           // nrows_sum (src/mcem_hcbn.cpp:356-361)
  bp.reps = (int)(L / (unsigned)bp.n_ball);   // :362
  bp.L_eff_total = bp.reps * bp.n_ball;       // :363
  return bp;
}

}  // namespace mccbn

struct CondLauncher {
  template <typename real, int PMAX>
  static void go(const CondArgs& a, const FlatPoset& f, bool times_given, int mode, int grid, cudaStream_t s) {
    const PosetProg pg = build_prog(f, PMAX, (int)sizeof(real));
#define MCCBN_COND(TG, MD) estep_conditional_kernel<real, PMAX, TG, MD><<<grid, kFwdBlock, 0, s>>>(a, pg)
    if (mode == kCondBackward) {
      if (times_given) MCCBN_COND(true, kCondBackward); else MCCBN_COND(false, kCondBackward);
    } else if (mode == kCondBernoulli) {
      if (times_given) MCCBN_COND(true, kCondBernoulli); else MCCBN_COND(false, kCondBernoulli);
    } else if (mode == kCondAddRemove) {
      if (times_given) MCCBN_COND(true, kCondAddRemove); else MCCBN_COND(false, kCondAddRemove);
    } else {
      if (times_given) MCCBN_COND(true, kCondOtcbn); else MCCBN_COND(false, kCondOtcbn);
    }
#undef MCCBN_COND
  }
};

bool mccbn_problem::jit_ready(int c, int scheme, bool times_given) const {
  if (!ctx->jit) return false;
  if (scheme == kForward) return ctx->jit->has(flat[c], times_given, -1);
  if (scheme == kPool) return false;
  const int mode = scheme == kBackward ? kCondBackward
                                       : (scheme == kBernoulli ? kCondBernoulli : (scheme == kAddRemove ? kCondAddRemove : kCondOtcbn));
  return ctx->jit->has(flat[c], times_given, mode);
}

void mccbn_problem::prefetch_kernel(const Poset& P, int scheme, bool times_given) {
  if (jit_mode_effective() == 0) return;
  if (!ctx->jit) ctx->jit = new JitCache();
  FlatPoset f;
  P.flatten(f);
  if (scheme == kForward) {
    ctx->jit->get(f, times_given, JitCache::kBackground);
  } else if (scheme == kBackward || scheme == kBernoulli || scheme == kAddRemove) {
    const int mode = scheme == kBackward ? kCondBackward : (scheme == kBernoulli ? kCondBernoulli : kCondAddRemove);
    ctx->jit->get_cond(f, times_given, mode, JitCache::kBackground);
  }
}

void mccbn_problem::enqueue_iteration(int c, Scheme scheme, int L, unsigned neighborhood_dist, bool times_given,
                                      uint32_t seed, int update_params, bool per_obs_out, int precision) {
  NvtxRange nvtx(scheme == kForward ? "E-step forward" : (scheme == kPool ? "E-step pool" : "E-step conditional"));
  if (p > kWideMaxP) throw Error(MCCBN_ERR_ARG, "the Philox E-step supports p <= 64");
  if (p > kFastMaxP && (scheme == kPool || scheme == kAddRemove || precision == 1))
    throw Error(MCCBN_ERR_ARG, "32 < p <= 64 is supported by the forward, backward and bernoulli schemes (fp32 path)");
  cudaStream_t s = ctx->stream;
  const int iter = iter_host[c];
  int chunks = 1, chunk_len = L, grid_x = 1;
  long long n_items = 0;
  int L_eff = L;

  if (scheme == kForward) {
    JitKernel* jk = precision == 1 ? nullptr : resolve_jit(flat[c], times_given, L);
    choose_chunks(L, chunks, chunk_len, grid_x, jk ? jk->blocks_per_sm : 0);
    n_items = (long long)N * chunks;
    FwdArgs a;
    std::memset(&a, 0, sizeof(a));
    a.obs_bits = bits.ptr;
    a.times = times.ptr;
    a.model = models.ptr + c;
    a.scale = scales.ptr + c;
    a.part = part.ensure((size_t)n_items * PS());
    a.N = N;
    a.L = L;
    a.chunks = chunks;
    a.chunk_len = chunk_len;
    a.n_items = n_items;
    a.obs_offset = obs_offset;
    const PhiloxKey key = make_key(seed, (uint32_t)iter, 0u);
    a.key0 = key.k0;
    a.key1 = key.k1;
    a.cw = eval_id[c];
    a.PS = PS();
    // warps claim items from a counter instead of striding: all items cost the same, but SMs do not finish at the same
    // time (cfg 3: 22.1 -> 20.9 ms); MCCBN_FWD_DYNAMIC=0 restores the static assignment
    static const bool dynamic_items = [] {
      const char* e = std::getenv("MCCBN_FWD_DYNAMIC");
      return !(e && std::strcmp(e, "0") == 0);
    }();
    if (dynamic_items) a.next_item = armed_item_counter();
    if (precision == 1)
      launch_forward<double>(a, flat[c], times_given, grid_x, c);
    else if (jk)
      launch_forward_jit(jk, a, flat[c], c, grid_x);
    else
      launch_forward<float>(a, flat[c], times_given, grid_x, c);
  } else if (scheme == kBackward || scheme == kBernoulli || scheme == kOtcbn || scheme == kAddRemove) {
    CondArgs a;
    std::memset(&a, 0, sizeof(a));
    int mode = scheme == kBackward ? kCondBackward
                                   : (scheme == kBernoulli ? kCondBernoulli : (scheme == kAddRemove ? kCondAddRemove : kCondOtcbn));
    if (scheme == kBackward) {
      BackwardPlan bp = plan_backward(p, neighborhood_dist, (unsigned)L);
      if (bp.reps == 0) throw Error(MCCBN_ERR_ZERO_WEIGHT, kMsgZeroWeight);  // L < nrows_sum: no samples at all
      if (ball_p != p || ball_k != (int)neighborhood_dist) {
        ball.ensure(bp.masks.size());
        MCCBN_CUDA(cudaMemcpyAsync(ball.ptr, bp.masks.data(), sizeof(uint64_t) * bp.masks.size(),
                                   cudaMemcpyHostToDevice, s));
        MCCBN_CUDA(cudaStreamSynchronize(s));
        ball_p = p;
        ball_k = (int)neighborhood_dist;
      }
      a.ball = ball.ptr;
      a.n_ball = bp.n_ball;
      a.reps = bp.reps;
      a.L = bp.L_eff_total;
      // chunk over ball genotypes; keep >= ~256 samples per chunk
      const int per_geno = bp.reps;
      int genos_per_chunk = std::max(1, (256 + per_geno - 1) / per_geno);
      const long long Ng = std::max<long long>(1, N_global > 0 ? N_global : N);  // chunking independent of the sharding
      long long want = (want_items_global() / 4 + Ng - 1) / Ng;  // items of a few ball genotypes are cheap: coarser than the forward rule
      int max_chunks = std::max(1, bp.n_ball / genos_per_chunk);
      chunks = (int)std::max(1LL, std::min<long long>(want, max_chunks));
      chunk_len = (bp.n_ball + chunks - 1) / chunks;
      chunks = (bp.n_ball + chunk_len - 1) / chunk_len;
      L_eff = 0;  // #{w > 0} per observation (src/mcem_hcbn.cpp:688-689)
    } else {
      int gx_unused;
      choose_chunks(L, chunks, chunk_len, gx_unused);
      a.L = L;
      L_eff = L;
    }
    n_items = (long long)N * chunks;
    a.obs_bits = bits.ptr;
    a.times = times.ptr;
    a.model = models.ptr + c;
    a.part = part.ensure((size_t)n_items * PS());
    a.N = N;
    a.chunks = chunks;
    a.chunk_len = chunk_len;
    a.n_items = n_items;
    a.obs_offset = obs_offset;
    const PhiloxKey key = make_key(seed, (uint32_t)iter, scheme == kBackward ? 1u : (scheme == kBernoulli ? 2u : (scheme == kAddRemove ? 6u : 5u)));
    a.key0 = key.k0;
    a.key1 = key.k1;
    a.cw = eval_id[c];
    a.PS = PS();
    a.next_item = armed_item_counter();
    const int warps_per_block = kFwdBlock / 32;
    long long blocks = (n_items + warps_per_block - 1) / warps_per_block;
    // 32 < p <= 64 exists as a specialised kernel only: compile regardless of the work announced
    JitKernel* jkc = precision == 1 ? nullptr
                                    : resolve_jit_cond(flat[c], times_given, mode, (double)N * (double)a.L, p > kFastMaxP);
    if (p > kFastMaxP && !jkc)
      throw Error(MCCBN_ERR_CUDA, "32 < p <= 64 with this scheme needs the NVRTC-specialised kernel (libnvrtc unavailable?)");
    // one resident wave: MCCBN_COND_MINB blocks per SM for the fp32 kernels, 1 for the fp64 validation mode
    const int per_sm = jkc ? jkc->blocks_per_sm : (precision == 1 ? 1 : MCCBN_COND_MINB);
    grid_x = (int)std::max(1LL, std::min<long long>(blocks, (long long)ctx->sm_count * per_sm));
    if (jkc && p > kFastMaxP) {
      PosetProgWide pgc = build_prog_t<PosetProgWide>(flat[c], jkc->pmax, 4);
      CondArgs args = a;
      void* params[] = {(void*)&args, (void*)&pgc};
      MCCBN_CUDA(cudaLaunchKernel((const void*)jkc->kern, dim3(grid_x), dim3(kFwdBlock), params, 0, s));
    } else if (jkc) {
      PosetProg pgc = build_prog(flat[c], jkc->pmax, 4);
      CondArgs args = a;
      void* params[] = {(void*)&args, (void*)&pgc};
      MCCBN_CUDA(cudaLaunchKernel((const void*)jkc->kern, dim3(grid_x), dim3(kFwdBlock), params, 0, s));
    } else if (precision == 1) {
      CondLauncher::go<double, 32>(a, flat[c], times_given, mode, grid_x, s);
    } else if (p <= 16) {
      CondLauncher::go<float, 16>(a, flat[c], times_given, mode, grid_x, s);
    } else if (p <= 24) {
      CondLauncher::go<float, 24>(a, flat[c], times_given, mode, grid_x, s);
    } else {
      CondLauncher::go<float, 32>(a, flat[c], times_given, mode, grid_x, s);
    }
    ctx->launches++;
    if (jkc) ctx->jit_launches++;
    MCCBN_CUDA(cudaGetLastError());
  } else if (scheme == kPool) {
    // pool of K = p * L prior draws shared by all observations, regenerated every EM iteration (:606, :650-654)
    const int K = p * L;
    const int PM = p <= 8 ? 8 : (p <= 16 ? 16 : 32);  // powers of two: the kernel binary-searches the sorted rows
    float* Tp = pool_T.ensure((size_t)K * PM);
    float* Zp = pool_Z.ensure((size_t)K * PM);
    uint32_t* Mp = pool_M.ensure((size_t)K * (PM + 1));
    const PhiloxKey kgen = make_key(seed, (uint32_t)iter, 3u);
    const PhiloxKey kts = make_key(seed, (uint32_t)iter, 4u);
    const int ggen = std::max(1, std::min((K + kFwdBlock - 1) / kFwdBlock, ctx->sm_count * 6));
    const int nb = (N + kPoolBlock - 1) / kPoolBlock;
    const long long nb_global = ((N_global > 0 ? N_global : N) + kPoolBlock - 1) / kPoolBlock;  // independent of the sharding
    long long want = ((long long)ctx->sm_count * 6 * kShardRanks + nb_global - 1) / nb_global;
    chunks = (int)std::max(1LL, std::min<long long>(want, std::max(1, K / (4 * kPoolTile))));
    if (pool_resample) chunks = 1;  // the resampling sweep walks the cumulative weights of the whole pool
    chunk_len = ((K + chunks - 1) / chunks + kPoolTile - 1) / kPoolTile * kPoolTile;
    chunks = (K + chunk_len - 1) / chunk_len;
    n_items = (long long)N * chunks;
    PoolArgs a;
    std::memset(&a, 0, sizeof(a));
    a.obs_bits = bits.ptr;
    a.times = times.ptr;
    a.model = models.ptr + c;
    a.scale = scales.ptr + c;
    a.Tp = Tp;
    a.Mp = Mp;
    a.Zp = Zp;
    a.part = part.ensure((size_t)n_items * PS());
    a.N = N;
    a.K = K;
    a.chunks = chunks;
    a.chunk_len = chunk_len;
    a.obs_offset = obs_offset;
    a.key0 = kts.k0;
    a.key1 = kts.k1;
    a.cw = eval_id[c];
    a.PS = PS();
    if (pool_resample) {
      a.dref = pool_dref.ensure(N);
      a.qtot = pool_qtot.ensure(N);
      a.L = L;
      a.rkey1 = make_key(seed, (uint32_t)iter, 9u).k1;
    }
    const dim3 grid(nb, chunks);
    // weighted sums on the tensor cores (estep_pool_mma_kernel); MCCBN_POOL_MMA=0 keeps the FFMA kernel (A/B runs)
    static const bool pool_mma = [] {
      const char* e = std::getenv("MCCBN_POOL_MMA");
      return !(e && std::strcmp(e, "0") == 0);
    }();
#define MCCBN_POOL(PMX)                                                                                         \
  do {                                                                                                          \
    const PosetProg pg = build_prog(flat[c], PMX, 4);                                                           \
    pool_generate_kernel<PMX><<<ggen, kFwdBlock, fwd_smem_bytes<float, PMX>(), s>>>(Tp, Mp, Zp, K, kgen.k0,     \
                                                                                    kgen.k1, eval_id[c],       \
                                                                                    models.ptr + c,            \
                                                                                    scales.ptr + c, pg);        \
    if (pool_mma) {                                                                                             \
      if (times_given) estep_pool_mma_kernel<PMX, true><<<grid, kPoolBlock, 0, s>>>(a, pg);                     \
      else estep_pool_mma_kernel<PMX, false><<<grid, kPoolBlock, 0, s>>>(a, pg);                                \
    } else if (times_given)                                                                                     \
      estep_pool_kernel<PMX, true><<<grid, kPoolBlock, 0, s>>>(a, pg);                                          \
    else                                                                                                        \
      estep_pool_kernel<PMX, false><<<grid, kPoolBlock, 0, s>>>(a, pg);                                         \
    if (pool_resample) { /* the reference's L-index resampling: two more sweeps over the pool */                \
      if (times_given) {                                                                                        \
        estep_pool_resample_kernel<PMX, true, 0><<<nb, kPoolBlock, 0, s>>>(a, pg);                              \
        estep_pool_resample_kernel<PMX, true, 1><<<nb, kPoolBlock, 0, s>>>(a, pg);                              \
      } else {                                                                                                  \
        estep_pool_resample_kernel<PMX, false, 0><<<nb, kPoolBlock, 0, s>>>(a, pg);                             \
        estep_pool_resample_kernel<PMX, false, 1><<<nb, kPoolBlock, 0, s>>>(a, pg);                             \
      }                                                                                                         \
      ctx->launches += 2;                                                                                       \
    }                                                                                                           \
  } while (0)
    if (PM == 8) MCCBN_POOL(8);
    else if (PM == 16) MCCBN_POOL(16);
    else MCCBN_POOL(32);
#undef MCCBN_POOL
    ctx->launches += 2;
    MCCBN_CUDA(cudaGetLastError());
    L_eff = K;  // obs_llhood_i = log(sum_k q_k / K)  (w_l = sum q / K for all L resampled rows, :491)
  } else {
    throw Error(MCCBN_ERR_ARG, "unknown sampling scheme");
  }
  run_finalize_mstep(c, chunks, n_items, L_eff, per_obs_out, update_params);
}

// ----------------------------------------------------------------------------------------------------------------
// EM driver: the host-side control flow of MCEM_hcbn (src/mcem_hcbn.cpp:571-736) for candidate slot 0..n_cand-1
// ----------------------------------------------------------------------------------------------------------------
namespace mccbn {

struct EmControl {  // ControlEM (src/mcem.hpp:213-226)
  unsigned max_iter = 100, update_step_size = 20;
  double tol = 0.001;
  float max_lambda = 1e6f;
  unsigned neighborhood_dist = 1;
};

struct EmResult {
  std::vector<double> lambda;
  double eps = 0, llhood = 0;
  int n_iter = 0;
};

// Host-side state of one MCEM run (the batch-average / convergence bookkeeping of :583-642, :711-721)
struct EmRun {
  int slot = 0;
  std::vector<double> avg_lambda, avg_lambda_current;
  double avg_eps = 0.0, avg_eps_current = 0.0, avg_llhood = 0.0;
  unsigned boundary = 0, iter = 0;
  bool done = false;
  double* trace_out = nullptr;
};

// Runs MCEM concurrently on several candidate slots (one poset each) of the same observation shard: every round
// enqueues, for every still-active slot, the iterations up to its next batch boundary, then synchronises ONCE and
// applies the reference's batch-average / convergence rule per slot.  With one slot this is exactly MCEM_hcbn.
static std::vector<EmResult> run_mcem_multi(mccbn_problem& pb, const std::vector<int>& slots, Scheme scheme,
                                            unsigned L, const EmControl& ctl, bool times_given, uint32_t seed,
                                            bool verbose, int precision, double* trace_out, int* n_iter_out,
                                            const std::function<void()>& while_device_busy = nullptr) {
  NvtxRange nvtx("MCEM_hcbn");
  const int p = pb.p;
  if (ctl.update_step_size == 0) throw Error(MCCBN_ERR_ARG, "update_step_size must be >= 1");
  const int stride = p + 2;
  pb.ensure_trace((int)std::max(1u, ctl.max_iter));
  pb.expected_esteps = (int)std::max(1u, ctl.max_iter / 2);
  std::vector<EmRun> runs(slots.size());
  for (size_t r = 0; r < slots.size(); ++r) {
    runs[r].slot = slots[r];
    runs[r].avg_lambda.assign(p, 0.0);
    runs[r].avg_lambda_current.assign(p, 0.0);
    runs[r].boundary = ctl.update_step_size;
    runs[r].done = ctl.max_iter == 0;
    runs[r].trace_out = (r == 0) ? trace_out : nullptr;
  }
  if (verbose) {
    DevModel dm;
    pb.read_model(slots[0], dm);
    std::printf("Initial value of the error rate - epsilon: %g\n", dm.eps);
    std::printf("Initial value of the rate parameters - lambda:");
    for (int j = 0; j < p; ++j) std::printf(" %g", dm.lambda[j]);
    std::printf("\n");
  }
  std::vector<std::vector<double>> rows(slots.size());
  std::vector<DevModel> dms(slots.size());
  bool first_round = true;
  for (;;) {
    bool any = false;
    std::vector<unsigned> n_run(slots.size(), 0);
    for (size_t r = 0; r < runs.size(); ++r) {
      EmRun& R = runs[r];
      if (R.done) continue;
      if (R.iter == R.boundary) {  // :622-642
        for (auto& x : R.avg_lambda_current) x /= ctl.update_step_size;
        R.avg_eps_current /= ctl.update_step_size;
        R.avg_llhood /= ctl.update_step_size;
        bool ok = std::abs(R.avg_eps - R.avg_eps_current) <= ctl.tol;
        for (int j = 0; j < p && ok; ++j) ok = std::abs(R.avg_lambda[j] - R.avg_lambda_current[j]) <= ctl.tol;
        if (ok) {
          R.done = true;
          continue;
        }
        R.avg_lambda = R.avg_lambda_current;
        R.avg_eps = R.avg_eps_current;
        R.boundary += ctl.update_step_size;
        std::fill(R.avg_lambda_current.begin(), R.avg_lambda_current.end(), 0.0);
        R.avg_eps_current = 0.0;
        R.avg_llhood = 0.0;
      }
      n_run[r] = std::min(R.boundary, ctl.max_iter) - R.iter;
      pb.enqueue_iterations(R.slot, (int)n_run[r], scheme, (int)L, ctl.neighborhood_dist, times_given, seed, 1, precision);
      rows[r].resize((size_t)n_run[r] * stride);
      MCCBN_CUDA(cudaMemcpyAsync(rows[r].data(), pb.trace_of(R.slot) + (size_t)R.iter * stride,
                                 sizeof(double) * rows[r].size(), cudaMemcpyDeviceToHost, pb.ctx->stream));
      MCCBN_CUDA(cudaMemcpyAsync(&dms[r], pb.models.ptr + R.slot, sizeof(DevModel), cudaMemcpyDeviceToHost,
                                 pb.ctx->stream));
      any = true;
    }
    if (!any) break;
    // the device is busy with the iterations just enqueued: a slice of the caller's host-side work overlaps with them
    (void)first_round;
    if (while_device_busy) while_device_busy();
    if (pb.ctx->jit) pb.ctx->jit->service();  // collect finished background compilations
    {
      NvtxRange nvtx_sync("EM batch: wait for the device");
      MCCBN_CUDA(cudaStreamSynchronize(pb.ctx->stream));  // one host sync per round
    }
    for (size_t r = 0; r < runs.size(); ++r) {
      EmRun& R = runs[r];
      if (R.done || n_run[r] == 0) continue;
      if (dms[r].comm_error) throw Error(MCCBN_ERR_CUDA, kMsgCommTimeout);
      if (dms[r].zero_weight) throw Error(MCCBN_ERR_ZERO_WEIGHT, kMsgZeroWeight);
      for (unsigned t = 0; t < n_run[r]; ++t) {
        const double* row = rows[r].data() + (size_t)t * stride;
        for (int j = 0; j < p; ++j) R.avg_lambda_current[j] += row[2 + j];  // :711-713
        R.avg_eps_current += row[1];
        R.avg_llhood += row[0];
        if (verbose && r == 0) {
          if (R.iter + t == 0) std::printf("llhood\tepsilon\tlambdas\n");
          std::printf("%g\t%g\t", row[0], row[1]);
          for (int j = 0; j < p; ++j) std::printf(" %g", row[2 + j]);
          std::printf("\n");
        }
        if (R.trace_out) std::memcpy(R.trace_out + (size_t)(R.iter + t) * stride, row, sizeof(double) * stride);
      }
      R.iter += n_run[r];
      if (R.iter == ctl.max_iter) {  // :715-721
        const unsigned num_iter = ctl.max_iter - R.boundary + ctl.update_step_size;
        for (auto& x : R.avg_lambda_current) x /= num_iter;
        R.avg_eps_current /= num_iter;
        R.avg_llhood /= num_iter;
        R.done = true;
      }
    }
  }
  std::vector<EmResult> out(slots.size());
  for (size_t r = 0; r < runs.size(); ++r) {
    out[r].lambda = runs[r].avg_lambda_current;  // :731-735: the batch averages are the result
    out[r].eps = runs[r].avg_eps_current;
    out[r].llhood = runs[r].avg_llhood;
    out[r].n_iter = (int)runs[r].iter;
  }
  if (n_iter_out) *n_iter_out = out[0].n_iter;
  return out;
}

static EmResult run_mcem(mccbn_problem& pb, Scheme scheme, unsigned L, const EmControl& ctl, bool times_given,
                         uint32_t seed, bool verbose, int precision, double* trace_out, int* n_iter_out) {
  return run_mcem_multi(pb, std::vector<int>{0}, scheme, L, ctl, times_given, seed, verbose, precision, trace_out,
                        n_iter_out)[0];
}

static std::mutex g_default_mu;
static mccbn_ctx* g_default_ctx = nullptr;

static void ctx_init(mccbn_ctx* c, int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw Error(MCCBN_ERR_CUDA, std::string("no CUDA device available (libmccbn_b200 has no CPU fallback): ") +
                                    cudaGetErrorString(e));
  if (device < 0 || device >= n) throw Error(MCCBN_ERR_ARG, "invalid CUDA device index");
  c->device = device;
  MCCBN_CUDA(cudaSetDevice(device));
  MCCBN_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  cudaDeviceProp prop;
  MCCBN_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
}

// dev aid (MCCBN_SEGV_TRACE=1): native backtrace of the faulting thread on SIGSEGV, then exit
static void segv_trace_handler(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "mccbn: SIGSEGV, native backtrace of the faulting thread:\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) {}
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}
static void maybe_install_segv_trace() {
  static const bool on = [] {
    if (!std::getenv("MCCBN_SEGV_TRACE")) return false;
    std::signal(SIGSEGV, segv_trace_handler);
    return true;
  }();
  (void)on;
}

static mccbn_ctx* get_ctx(mccbn_ctx* c) {
  maybe_install_segv_trace();
  if (c) {
    MCCBN_CUDA(cudaSetDevice(c->device));
    return c;
  }
  std::lock_guard<std::mutex> lk(g_default_mu);
  if (!g_default_ctx) {
    int dev = 0;
    if (const char* lr = std::getenv("LOCAL_RANK")) dev = std::atoi(lr);
    int n = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) dev %= n;
    mccbn_ctx* nc = new mccbn_ctx();
    try {
      ctx_init(nc, dev);
    } catch (...) {
      delete nc;
      throw;
    }
    g_default_ctx = nc;
  }
  MCCBN_CUDA(cudaSetDevice(g_default_ctx->device));
  return g_default_ctx;
}

static int report(const std::exception& e, char* err, size_t errlen) {
  int code = MCCBN_ERR_CUDA;
  if (auto* me = dynamic_cast<const Error*>(&e)) code = me->code;
  if (err && errlen) {
    std::snprintf(err, errlen, "%s", e.what());
  }
  return code;
}

// ---- in-process device team ---------------------------------------------------------------------------------------
// The reference's callers are ONE process (an R session calling .Call): with MCCBN_GPUS=G (or mccbn_set_gpus) the
// reference-facing entries that are given no context and no caller-side sharding split the observations over G
// devices themselves -- one host thread and one context per device, the very code path a rank of a multi-process job
// runs (global observation indices in the Philox counters => the result does not depend on G).  The statistic
// exchange uses the same mailbox protocol as the multi-process path; the mailboxes of a team live in pinned, portable
// host memory, which every device of the process can address without enabling peer access (p + 5 doubles per
// iteration: the few microseconds of PCIe latency are irrelevant next to a millisecond E-step, and no later
// cudaMalloc has to be mapped into the peers).
// One persistent host thread per team member: a call hands every worker its closure and waits for all of them (creating
// and joining G threads per call cost a third of a millisecond at G = 8 -- a tenth of a BASELINE config-3 E-step there).
struct TeamWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, done = false, quit = false;
  void loop() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [&] { return has_job || quit; });
      if (quit) return;
      lk.unlock();
      job();
      lk.lock();
      has_job = false;
      done = true;
      cv.notify_all();
    }
  }
  void start() { th = std::thread([this] { loop(); }); }
  void submit(std::function<void()> f) {
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(f);
    done = false;
    has_job = true;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
  }
};
struct Team {
  std::vector<mccbn_ctx*> ctx;
  Mailbox* boxes = nullptr;  // cudaHostAlloc(portable | mapped), one per member
  std::vector<TeamWorker*> workers;  // intentionally never joined: they park on their condition variables until exit
};
static Team* g_team = nullptr;
static int g_team_want = -1;  // -1: read MCCBN_GPUS on first use
static int g_team_jit_mode = -1;  // mccbn_set_jit(NULL, mode) applies to the team members too

static int team_want_locked() {
  if (g_team_want < 0) {
    g_team_want = 1;
    if (const char* e = std::getenv("MCCBN_GPUS")) {
      int n = 0;
      cudaGetDeviceCount(&n);
      cudaGetLastError();
      g_team_want = std::strcmp(e, "all") == 0 ? std::max(1, n) : std::max(1, std::atoi(e));
    }
  }
  return g_team_want;
}

// devices a call with N observations would be sharded over (1 = no team); g_default_mu held
static int team_size_locked(int N) {
  int G = std::min(team_want_locked(), kMaxWorld);
  if (G > 1) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    cudaGetLastError();
    // one member per device: members that shared a device would deadlock -- a member's tail kernel spins until the
    // other member has published its statistics, and any device-wide synchronising call of that other member's host
    // thread (cudaMalloc, a module load) waits for the spinning kernel.  (Separate PROCESSES on one device are fine:
    // tests/test_gpu_multi.py::test_sharded_em_two_ranks_one_gpu.)
    G = std::min(G, n);
  }
  return (G <= 1 || N < G) ? 1 : G;
}

// the team to shard over, or nullptr (single device).  N = number of observations of the call.
static Team* team_for(int N) {
  std::lock_guard<std::mutex> lk(g_default_mu);
  const int G = team_size_locked(N);
  if (G <= 1) return nullptr;
  if (g_team && (int)g_team->ctx.size() == G) return g_team;
  if (g_team) {  // size changed: the old team's contexts stay alive (a handful of streams) but are no longer used
    for (mccbn_ctx* c : g_team->ctx) c->peer_ready = false;
  }
  int caller_device = 0;  // creating the contexts must not move the calling thread to another device
  cudaGetDevice(&caller_device);
  Team* t = new Team();
  MCCBN_CUDA(cudaHostAlloc((void**)&t->boxes, sizeof(Mailbox) * G, cudaHostAllocPortable | cudaHostAllocMapped));
  std::memset(t->boxes, 0, sizeof(Mailbox) * G);
  for (int r = 0; r < G; ++r) {
    mccbn_ctx* c = new mccbn_ctx();
    ctx_init(c, r);
    c->rank = r;
    c->world = G;
    c->mailbox = &t->boxes[r];
    c->mailbox_host = true;
    for (int q = 0; q < G; ++q) c->peer_box[q] = &t->boxes[q];
    c->peer_ready = true;
    c->jit_mode = g_team_jit_mode;
    t->ctx.push_back(c);
    TeamWorker* w = new TeamWorker();
    w->start();
    t->workers.push_back(w);
  }
  cudaSetDevice(caller_device);
  g_team = t;
  return t;
}

// f(rank, ctx, err, errlen) -> return code, on one host thread per member; the first failing rank's code and message
// are reported (a rank that fails alone leaves the others waiting for its statistics until the peer timeout).
template <class F>
static void team_run(Team& t, F&& f) {
  static std::mutex run_mu;  // one team call at a time: the members' streams, mailboxes and exchange numbers are shared
  std::lock_guard<std::mutex> run_lk(run_mu);
  const int G = (int)t.ctx.size();
  // every member starts the call from the same exchange number
  unsigned long long base = 0;
  for (mccbn_ctx* c : t.ctx) base = std::max(base, c->xchg_seq);
  for (mccbn_ctx* c : t.ctx) c->xchg_seq = base;
  std::vector<int> rc(G, MCCBN_OK);
  std::vector<std::array<char, 512>> msg(G);
  for (int r = 0; r < G; ++r) {
    msg[r][0] = 0;
    t.workers[r]->submit([&, r] { rc[r] = f(r, t.ctx[r], msg[r].data(), msg[r].size()); });
  }
  for (int r = 0; r < G; ++r) t.workers[r]->wait();
  int bad = -1;
  for (int r = 0; r < G; ++r)
    if (rc[r] != MCCBN_OK && (bad < 0 || (std::strstr(msg[bad].data(), "timed out") && !std::strstr(msg[r].data(), "timed out"))))
      bad = r;
  if (bad >= 0) throw Error(rc[bad], msg[bad].data());
}

static inline int shard_begin(int N, int G, int r) { return (int)((long long)N * r / G); }
// a call shards itself only when the caller passed no context and does not shard on its own
static inline bool caller_shards(const mccbn_em_options* opt, int N) {
  return opt && (opt->obs_offset != 0 || (opt->N_global > 0 && opt->N_global != N));
}

// host-built weight table, same libm expression as the reference (src/mcem_hcbn.cpp:387-388)
static std::vector<double> host_wtab(double eps, int p) {
  std::vector<double> w(65, 0.0);
  for (int d = 0; d <= p; ++d) w[d] = std::pow(eps, (double)d) * std::pow(1 - eps, (double)p - (double)d);
  return w;
}

static uint64_t pack_genotype(const int* g, int p) {
  uint64_t w = 0;
  for (int j = 0; j < p; ++j) w |= (uint64_t)(g[j] != 0) << j;
  return w;
}

}  // namespace mccbn

#define MCCBN_TRY try {
#define MCCBN_CATCH                                           \
  }                                                           \
  catch (const std::exception& e) { return report(e, err, errlen); } \
  return MCCBN_OK;

namespace mccbn {
// initialize_lambda from device counts (src/asa.cpp:92-128): float arithmetic reproduced on the host
static void lambda_from_counts(const FlatPoset& f, const unsigned long long* cnt, int N, float lambda_s,
                               float max_lambda, double* lambda_out) {
  for (int j = 0; j < f.p; ++j) {
    unsigned count = 1;
    float mutated = 1e-7f;
    if (f.ancestor_ev[j]) {
      count += (unsigned)cnt[j * 2 + 0];
      mutated += (int)cnt[j * 2 + 1];
    } else {
      count += (unsigned)N;
      mutated += (int)cnt[j * 2 + 1];
    }
    const float aux = mutated / count;
    double lam = aux * lambda_s / (1.0 - aux);
    if (lam > max_lambda || !std::isfinite(lam)) lam = max_lambda;
    lambda_out[j] = lam;
  }
}
}  // namespace mccbn

#include <array>
#include <chrono>
#include <thread>
#include "engine_asa.cuh"

// ================================================================================================================
// C ABI
// ================================================================================================================
extern "C" {

const char* mccbn_version(void) { return "mccbn_b200 0.1 (sm_100a)"; }

int mccbn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int mccbn_ctx_create(mccbn_ctx** out, int device, char* err, size_t errlen) {
  MCCBN_TRY
  mccbn_ctx* c = new mccbn_ctx();
  try {
    ctx_init(c, device);
  } catch (...) {
    delete c;
    throw;
  }
  *out = c;
  MCCBN_CATCH
}

void mccbn_ctx_destroy(mccbn_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  delete c->jit;
  for (int q = 0; q < kMaxWorld; ++q)
    if (!c->mailbox_host && c->peer_box[q] && c->peer_box[q] != c->mailbox) cudaIpcCloseMemHandle(c->peer_box[q]);
  if (c->mailbox && !c->mailbox_host) cudaFree(c->mailbox);
  if (c->comm) c->nccl.CommDestroy(c->comm);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  c->stage_release();
  delete c;
}

int mccbn_ctx_set_stream(mccbn_ctx* c, void* stream) {
  char* err = nullptr;
  size_t errlen = 0;
  MCCBN_TRY
  c = get_ctx(c);
  // work of one context is ordered by ONE stream at a time (its specialised kernels read the context's __constant__
  // scales, its buffers are reused from call to call): drain the old stream before switching
  MCCBN_CUDA(cudaStreamSynchronize(c->stream));
  c->stream = stream ? (cudaStream_t)stream : c->own_stream;
  MCCBN_CATCH
}

int mccbn_nccl_unique_id(unsigned char id[MCCBN_NCCL_ID_BYTES], char* err, size_t errlen) {
  MCCBN_TRY
  NcclApi api;
  api.load();
  NcclApi::unique_id u;
  api.check(api.GetUniqueId(&u), "ncclGetUniqueId");
  std::memcpy(id, u.internal, MCCBN_NCCL_ID_BYTES);
  MCCBN_CATCH
}

int mccbn_ctx_init_comm(mccbn_ctx* c, int rank, int world, const unsigned char id[MCCBN_NCCL_ID_BYTES], char* err,
                        size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (world < 1 || rank < 0 || rank >= world) throw Error(MCCBN_ERR_ARG, "invalid rank/world");
  c->rank = rank;
  c->world = world;
  if (world > 1) {
    c->nccl.load();
    NcclApi::unique_id u;
    std::memcpy(u.internal, id, MCCBN_NCCL_ID_BYTES);
    c->nccl.check(c->nccl.CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank");
  }
  MCCBN_CATCH
}

// ---- peer-memory statistic exchange: bootstrap -------------------------------------------------------------------
int mccbn_ctx_peer_handle(mccbn_ctx* c, unsigned char handle[MCCBN_IPC_HANDLE_BYTES], char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  static_assert(sizeof(cudaIpcMemHandle_t) == MCCBN_IPC_HANDLE_BYTES, "IPC handle size");
  if (!c->mailbox) {
    MCCBN_CUDA(cudaMalloc((void**)&c->mailbox, sizeof(Mailbox)));
    MCCBN_CUDA(cudaMemset(c->mailbox, 0, sizeof(Mailbox)));
    MCCBN_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  MCCBN_CUDA(cudaIpcGetMemHandle(&h, c->mailbox));
  std::memcpy(handle, &h, MCCBN_IPC_HANDLE_BYTES);
  MCCBN_CATCH
}

int mccbn_ctx_init_peers(mccbn_ctx* c, int rank, int world, const unsigned char* handles, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
    throw Error(MCCBN_ERR_ARG, "invalid rank/world for the peer exchange (at most 16 ranks)");
  if (!c->mailbox) throw Error(MCCBN_ERR_ARG, "call mccbn_ctx_peer_handle first");
  if (c->comm && (c->rank != rank || c->world != world)) throw Error(MCCBN_ERR_ARG, "rank/world differ from the NCCL communicator");
  // the mailbox flags are monotone exchange numbers: a context that has already exchanged cannot be re-initialised (its
  // flags would satisfy the waits of a restarted sequence) -- create a new context instead
  if (c->xchg_seq != 0)
    throw Error(MCCBN_ERR_ARG, "this context has already exchanged statistics: create a new context to change the peer group");
  for (int q = 0; q < kMaxWorld; ++q) {  // a repeated bootstrap before the first exchange: drop the old mappings
    if (c->peer_box[q] && c->peer_box[q] != c->mailbox) cudaIpcCloseMemHandle(c->peer_box[q]);
    c->peer_box[q] = nullptr;
  }
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      c->peer_box[q] = c->mailbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)q * MCCBN_IPC_HANDLE_BYTES, MCCBN_IPC_HANDLE_BYTES);
    void* ptr = nullptr;
    MCCBN_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_box[q] = static_cast<Mailbox*>(ptr);
  }
  c->rank = rank;
  c->world = world;
  c->xchg_seq = 0;
  c->peer_ready = world > 1;
  MCCBN_CATCH
}

int mccbn_set_gpus(int n_gpus, char* err, size_t errlen) {
  MCCBN_TRY
  if (n_gpus < 0 || n_gpus > kMaxWorld) throw Error(MCCBN_ERR_ARG, "mccbn_set_gpus: between 0 and 16 devices");
  std::lock_guard<std::mutex> lk(g_default_mu);
  g_team_want = std::max(1, n_gpus);
  MCCBN_CATCH
}

int mccbn_get_gpus(void) {
  std::lock_guard<std::mutex> lk(g_default_mu);
  return team_size_locked(1 << 30);
}

uint64_t mccbn_kernel_launches(mccbn_ctx* c, int reset) {
  try {
    c = get_ctx(c);
  } catch (...) {
    return 0;
  }
  uint64_t v = c->launches;
  if (reset) c->launches = 0;
  return v;
}

uint64_t mccbn_jit_launches(mccbn_ctx* c, int reset) {
  try {
    c = get_ctx(c);
  } catch (...) {
    return 0;
  }
  uint64_t v = c->jit_launches;
  if (reset) c->jit_launches = 0;
  return v;
}

// ---- device-resident problem ------------------------------------------------------------------------------------
int mccbn_problem_create(mccbn_ctx* c, const int* obs, const double* times, const double* weights, int N, int p,
                         int64_t obs_offset, int64_t N_global, mccbn_problem** out, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (N < 1) throw Error(MCCBN_ERR_ARG, "N must be >= 1");
  if (p < 1 || p > kMaxP) throw Error(MCCBN_ERR_ARG, "p must be in [1, 64]");
  mccbn_problem* pb = new mccbn_problem();
  pb->ctx = c;
  pb->N = N;
  pb->p = p;
  pb->obs_offset = obs_offset;
  pb->N_global = N_global > 0 ? N_global : N;
  try {
    pb->upload(obs, times, weights);
    MCCBN_CUDA(cudaStreamSynchronize(c->stream));  // the caller's buffers are free to go when this entry returns
    pb->upload_inflight = false;
    pb->release_upload_staging();
  } catch (...) {
    delete pb;
    throw;
  }
  *out = pb;
  MCCBN_CATCH
}

void mccbn_problem_destroy(mccbn_problem* pb) {
  if (!pb) return;
  cudaSetDevice(pb->ctx->device);
  delete pb;
}

int mccbn_problem_set_model(mccbn_problem* pb, const int* poset, const double* lambda, double eps, float lambda_s,
                            float max_lambda, char* err, size_t errlen) {
  MCCBN_TRY
  get_ctx(pb->ctx);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, pb->p, P, f);
  pb->ensure_trace(4096);
  pb->set_model(0, f, lambda, eps, lambda_s, max_lambda);
  MCCBN_CATCH
}

int mccbn_problem_em_iterations(mccbn_problem* pb, unsigned L, const char* sampling, unsigned neighborhood_dist,
                                int sampling_times_available, int seed, int n_iter, int update_params, char* err,
                                size_t errlen) {
  MCCBN_TRY
  get_ctx(pb->ctx);
  const Scheme sc = parse_scheme(sampling);
  // a device-resident problem exists to run many E-steps on the same poset: announce that to the JIT policy
  pb->expected_esteps = std::max(pb->expected_esteps, std::max(n_iter, 64));
  pb->enqueue_iterations(0, n_iter, sc, (int)L, neighborhood_dist, sampling_times_available != 0, (uint32_t)seed,
                         update_params, 0);
  MCCBN_CATCH
}

int mccbn_problem_read(mccbn_problem* pb, double* lambda, double* eps, double* llhood, char* err, size_t errlen) {
  MCCBN_TRY
  get_ctx(pb->ctx);
  DevModel dm;
  pb->read_model(0, dm);
  if (lambda) std::memcpy(lambda, dm.lambda, sizeof(double) * pb->p);
  if (eps) *eps = dm.eps;
  if (llhood) *llhood = dm.llhood;
  if (dm.comm_error) throw Error(MCCBN_ERR_CUDA, kMsgCommTimeout);
  if (dm.zero_weight) throw Error(MCCBN_ERR_ZERO_WEIGHT, kMsgZeroWeight);
  MCCBN_CATCH
}

void* mccbn_problem_stream(mccbn_problem* pb) { return (void*)pb->ctx->stream; }

// ---- hot-path entry points --------------------------------------------------------------------------------------
int mccbn_MCEM_hcbn(mccbn_ctx* c, const double* ilambda, const int* poset, const int* obs, const double* times,
                    float lambda_s, double eps, const double* weights, unsigned L, const char* sampling,
                    unsigned max_iter, unsigned update_step_size, double tol, float max_lambda,
                    unsigned neighborhood_dist, int sampling_times_available, int thrds, int verbose, int seed, int N,
                    int p, double* lambda_out, double* eps_out, double* llhood_out, const mccbn_em_options* opt,
                    char* err, size_t errlen) {
  (void)thrds;
  MCCBN_TRY
  if (!c && !caller_shards(opt, N)) {
    if (Team* team = team_for(N)) {  // in-process multi-GPU: one thread per device, each running this entry on its shard
      const int G = (int)team->ctx.size();
      std::vector<std::vector<double>> lam_r(G, std::vector<double>(p));
      std::vector<double> eps_r(G), ll_r(G);
      team_run(*team, [&](int r, mccbn_ctx* cr, char* e, size_t el) {
        const int lo = shard_begin(N, G, r), hi = shard_begin(N, G, r + 1);
        mccbn_em_options o;
        if (opt) o = *opt; else std::memset(&o, 0, sizeof(o));
        o.obs_offset = lo;
        o.N_global = N;
        o.obs_ld = (uint32_t)N;  // rows [lo, hi) of the caller's matrix, uploaded by a strided copy
        if (r != 0) o.trace = nullptr, o.n_iter = nullptr;
        return mccbn_MCEM_hcbn(cr, ilambda, poset, obs + lo, times ? times + lo : nullptr, lambda_s, eps,
                               weights ? weights + lo : nullptr, L, sampling, max_iter, update_step_size, tol,
                               max_lambda, neighborhood_dist, sampling_times_available, thrds, r == 0 ? verbose : 0,
                               seed, hi - lo, p, lam_r[r].data(), &eps_r[r], &ll_r[r], &o, e, el);
      });
      std::copy(lam_r[0].begin(), lam_r[0].end(), lambda_out);  // the M-step is replicated: all ranks hold the same fit
      *eps_out = eps_r[0];
      *llhood_out = ll_r[0];
      return MCCBN_OK;
    }
  }
  c = get_ctx(c);
  const Scheme sc = parse_scheme(sampling);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  // Model::update_lambda(ilambda, max_lambda) (src/mcem_hcbn.cpp:837)
  std::vector<double> lam(ilambda, ilambda + p);
  for (auto& l : lam)
    if (l > max_lambda || !std::isfinite(l)) l = max_lambda;
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.obs_offset = opt ? opt->obs_offset : 0;
  pb.N_global = (opt && opt->N_global > 0) ? opt->N_global : N;
  pb.upload(obs, times, weights, opt ? opt->obs_ld : 0);
  if (opt) pb.eval_id[0] = opt->stream_id;
  pb.pool_resample = opt ? opt->pool_resample != 0 : false;
  pb.set_model(0, f, lam.data(), eps, lambda_s, max_lambda);
  EmControl ctl;
  ctl.max_iter = max_iter;
  ctl.update_step_size = update_step_size;
  ctl.tol = tol;
  ctl.max_lambda = max_lambda;
  ctl.neighborhood_dist = neighborhood_dist;
  EmResult r = run_mcem(pb, sc, L, ctl, sampling_times_available != 0, (uint32_t)seed, verbose != 0,
                        opt ? opt->precision : 0, opt ? opt->trace : nullptr, opt ? opt->n_iter : nullptr);
  std::copy(r.lambda.begin(), r.lambda.end(), lambda_out);
  *eps_out = r.eps;
  *llhood_out = r.llhood;
  MCCBN_CATCH
}

static void single_estep(mccbn_ctx* c, mccbn_problem& pb, const int* obs, const int* poset, const double* lambda,
                         double eps, const double* weights, const double* times, unsigned L, const char* sampling,
                         unsigned neighborhood_dist, float lambda_s, int times_available, int seed, int N, int p,
                         const mccbn_em_options* opt, bool per_obs) {
  const Scheme sc = parse_scheme(sampling);
  static const bool trace = std::getenv("MCCBN_TRACE_E2E") != nullptr;  // dev: host time per phase of the one-call path
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::micro>(b - a).count();
  };
  const auto t0 = now();
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.obs_offset = opt ? opt->obs_offset : 0;
  pb.N_global = (opt && opt->N_global > 0) ? opt->N_global : N;
  const auto t1 = now();
  pb.upload(obs, times, weights, opt ? opt->obs_ld : 0);
  const auto t2 = now();
  if (opt) pb.eval_id[0] = opt->stream_id;
  pb.pool_resample = opt ? opt->pool_resample != 0 : false;
  pb.ensure_trace(1);
  pb.set_model(0, f, lambda, eps, lambda_s, 1e6f);
  const auto t3 = now();
  pb.enqueue_iteration(0, sc, (int)L, neighborhood_dist, times_available != 0, (uint32_t)seed, 0, per_obs,
                       opt ? opt->precision : 0);
  if (trace)
    std::fprintf(stderr, "mccbn e2e: flatten %.0f us, upload %.0f us, set_model %.0f us, enqueue %.0f us\n", us(t0, t1),
                 us(t1, t2), us(t2, t3), us(t3, now()));
}

int mccbn_obs_log_likelihood(mccbn_ctx* c, const int* obs, const int* poset, const double* lambda, double eps,
                             const double* weights, const double* times, unsigned L, const char* sampling,
                             unsigned neighborhood_dist, float lambda_s, int sampling_times_available, int thrds,
                             int seed, int N, int p, double* llhood_out, const mccbn_em_options* opt, char* err,
                             size_t errlen) {
  (void)thrds;
  MCCBN_TRY
  if (!c && !caller_shards(opt, N)) {
    if (Team* team = team_for(N)) {
      const int G = (int)team->ctx.size();
      std::vector<double> ll_r(G);
      team_run(*team, [&](int r, mccbn_ctx* cr, char* e, size_t el) {
        const int lo = shard_begin(N, G, r), hi = shard_begin(N, G, r + 1);
        mccbn_em_options o;
        if (opt) o = *opt; else std::memset(&o, 0, sizeof(o));
        o.obs_offset = lo;
        o.N_global = N;
        o.obs_ld = (uint32_t)N;
        return mccbn_obs_log_likelihood(cr, obs + lo, poset, lambda, eps, weights ? weights + lo : nullptr,
                                        times ? times + lo : nullptr, L, sampling, neighborhood_dist, lambda_s,
                                        sampling_times_available, thrds, seed, hi - lo, p, &ll_r[r], &o, e, el);
      });
      *llhood_out = ll_r[0];  // summed over the ranks inside the reduction kernels
      return MCCBN_OK;
    }
  }
  c = get_ctx(c);
  mccbn_problem pb;
  single_estep(c, pb, obs, poset, lambda, eps, weights, times, L, sampling, neighborhood_dist, lambda_s,
               sampling_times_available, seed, N, p, opt, false);
  DevModel dm;
  pb.read_model(0, dm);
  if (dm.comm_error) throw Error(MCCBN_ERR_CUDA, kMsgCommTimeout);
  if (dm.zero_weight) throw Error(MCCBN_ERR_ZERO_WEIGHT, kMsgZeroWeight);
  *llhood_out = dm.llhood;
  MCCBN_CATCH
}

int mccbn_importance_weight(mccbn_ctx* c, const int* obs, unsigned L, const int* poset, const double* lambda,
                            double eps, const double* times, const char* sampling, unsigned neighborhood_dist,
                            float lambda_s, int sampling_times_available, int thrds, int seed, int N, int p,
                            double* w_sum, double* w_sum_sq, int* L_eff, double* dist, double* Tdiff,
                            const mccbn_em_options* opt, char* err, size_t errlen) {
  (void)thrds;
  MCCBN_TRY
  if (!c && !caller_shards(opt, N)) {
    if (Team* team = team_for(N)) {  // per-observation outputs: every rank fills the rows of its shard
      const int G = (int)team->ctx.size();
      team_run(*team, [&](int r, mccbn_ctx* cr, char* e, size_t el) {
        const int lo = shard_begin(N, G, r), hi = shard_begin(N, G, r + 1), n = hi - lo;
        mccbn_em_options o;
        if (opt) o = *opt; else std::memset(&o, 0, sizeof(o));
        o.obs_offset = lo;
        o.N_global = N;
        o.obs_ld = (uint32_t)N;
        std::vector<double> td(Tdiff ? (size_t)n * p : 0);
        const int rc = mccbn_importance_weight(cr, obs + lo, L, poset, lambda, eps, times ? times + lo : nullptr,
                                               sampling, neighborhood_dist, lambda_s, sampling_times_available, thrds,
                                               seed, n, p, w_sum ? w_sum + lo : nullptr,
                                               w_sum_sq ? w_sum_sq + lo : nullptr, L_eff ? L_eff + lo : nullptr,
                                               dist ? dist + lo : nullptr, Tdiff ? td.data() : nullptr, &o, e, el);
        if (rc == MCCBN_OK && Tdiff)
          for (int j = 0; j < p; ++j)
            std::memcpy(Tdiff + (size_t)j * N + lo, td.data() + (size_t)j * n, sizeof(double) * (size_t)n);
        return rc;
      });
      return MCCBN_OK;
    }
  }
  c = get_ctx(c);
  mccbn_problem pb;
  single_estep(c, pb, obs, poset, lambda, eps, nullptr, times, L, sampling, neighborhood_dist, lambda_s,
               sampling_times_available, seed, N, p, opt, true);
  cudaStream_t s = c->stream;
  if (w_sum) MCCBN_CUDA(cudaMemcpyAsync(w_sum, pb.out_w.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (w_sum_sq) MCCBN_CUDA(cudaMemcpyAsync(w_sum_sq, pb.out_w2.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (L_eff) MCCBN_CUDA(cudaMemcpyAsync(L_eff, pb.out_leff.ptr, sizeof(int) * N, cudaMemcpyDeviceToHost, s));
  if (dist) MCCBN_CUDA(cudaMemcpyAsync(dist, pb.out_dist.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, pb.out_T.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  if (parse_scheme(sampling) == kPool) {
    // every resampled row carries the weight sum_k q_k / K (src/mcem_hcbn.cpp:491): sum_l w_l = L sum(q)/K
    const double K = (double)p * (double)L;
    for (int i = 0; i < N; ++i) {
      if (w_sum) w_sum[i] = w_sum[i] * (double)L / K;
      if (w_sum_sq && w_sum) w_sum_sq[i] = w_sum[i] * w_sum[i] / (double)L;
    }
  }
  MCCBN_CATCH
}

int mccbn_importance_weight_genotype(mccbn_ctx* c, const int* genotype, unsigned L, const int* poset,
                                     const double* lambda, double eps, double time, const char* sampling,
                                     unsigned neighborhood_dist, float lambda_s, int sampling_times_available,
                                     int seed, int p, int* L_out, double* w, int* dist, double* Tdiff,
                                     const mccbn_em_options* opt, char* err, size_t errlen) {
  (void)neighborhood_dist;
  MCCBN_TRY
  c = get_ctx(c);
  const Scheme sc = parse_scheme(sampling);
  if (sc != kForward) throw Error(MCCBN_ERR_ARG, "importance_weight_genotype: only 'forward' exposes per-sample output");
  if (p > kFastMaxP) throw Error(MCCBN_ERR_ARG, "the Philox E-step supports p <= 32");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  int one_obs[64];
  for (int j = 0; j < p; ++j) one_obs[j] = genotype[j];
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = 1;
  pb.p = p;
  pb.upload(one_obs, &time, nullptr);
  pb.set_model(0, f, lambda, eps, lambda_s, 1e6f);
  cudaStream_t s = c->stream;
  DevBuf<double> dZ, dTs;
  DevBuf<int> dD;
  DevBuf<uint32_t> dX;
  FwdDumpArgs a;
  std::memset(&a, 0, sizeof(a));
  a.obs_word = pack_genotype(genotype, p);
  a.time = time;
  a.model = pb.models.ptr;
  a.scale = pb.scales.ptr;
  a.Z_out = dZ.ensure((size_t)L * p);
  a.dist_out = dD.ensure(L);
  a.X_out = dX.ensure(L);
  a.Ts_out = dTs.ensure(L);
  a.L = (int)L;
  a.gi = (uint32_t)(opt ? opt->obs_offset : 0);
  const PhiloxKey key = make_key((uint32_t)seed, 0u, 0u);
  a.key0 = key.k0;
  a.key1 = key.k1;
  a.cw = opt ? opt->stream_id : 0u;
  const int grid = (int)std::min<unsigned>((L + kFwdBlock - 1) / kFwdBlock, 4096u);
  const bool tg = sampling_times_available != 0;
  const int prec = opt ? opt->precision : 0;
  if (prec == 1) {
    const PosetProg pg = build_prog(f, 32, 8);
    constexpr size_t sm = fwd_smem_bytes<double, 32>();
    optin_smem(forward_dump_kernel<double, 32, true>, sm);
    optin_smem(forward_dump_kernel<double, 32, false>, sm);
    if (tg) forward_dump_kernel<double, 32, true><<<grid, kFwdBlock, sm, s>>>(a, pg);
    else forward_dump_kernel<double, 32, false><<<grid, kFwdBlock, sm, s>>>(a, pg);
  } else {
    const PosetProg pg = build_prog(f, 32, 4);
    constexpr size_t sm = fwd_smem_bytes<float, 32>();
    if (tg) forward_dump_kernel<float, 32, true><<<grid, kFwdBlock, sm, s>>>(a, pg);
    else forward_dump_kernel<float, 32, false><<<grid, kFwdBlock, sm, s>>>(a, pg);
  }
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  std::vector<int> hd(L);
  MCCBN_CUDA(cudaMemcpyAsync(hd.data(), dD.ptr, sizeof(int) * L, cudaMemcpyDeviceToHost, s));
  if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, dZ.ptr, sizeof(double) * (size_t)L * p, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  const std::vector<double> wt = host_wtab(eps, p);
  for (unsigned l = 0; l < L; ++l) {
    if (dist) dist[l] = hd[l];
    if (w) w[l] = wt[hd[l]];
  }
  if (L_out) *L_out = (int)L;
  MCCBN_CATCH
}

// ---- forward simulators (src/mcem_hcbn.cpp:1012-1120): the generative model on the device -----------------------
// One sample per thread through the fp64 forward sampler (53-bit uniforms): Z ~ Exp(lambda), T_s ~ Exp(lambda_s) unless
// given, T_j = Z_j + max parents, X_j = [T_j <= T_s].  All outputs optional.
// `given_times` != nullptr: the genotypes are formed on the device against the caller's per-sample sampling times.
static void simulate_forward(mccbn_ctx* c, unsigned N, const int* poset, const double* lambda, int p, float lambda_s,
                             int seed, double* T_sampling /* out */, int* samples, double* Tdiff, double* T_sum,
                             const double* given_times = nullptr) {
  if (p > kFastMaxP) throw Error(MCCBN_ERR_ARG, "the Philox samplers support p <= 32");
  if (N == 0) return;
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  int zero_obs[64] = {0};
  double zero_t = 0.0;
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = 1;
  pb.p = p;
  pb.upload(zero_obs, &zero_t, nullptr);
  pb.set_model(0, f, lambda, 0.5, lambda_s, 3.0e38f);
  cudaStream_t s = c->stream;
  DevBuf<double> dZ, dT, dTs;
  DevBuf<int> dD;
  DevBuf<uint32_t> dX;
  FwdDumpArgs a;
  std::memset(&a, 0, sizeof(a));
  a.model = pb.models.ptr;
  a.scale = pb.scales.ptr;
  a.Z_out = dZ.ensure((size_t)N * p);
  a.T_out = (T_sum || given_times) ? dT.ensure((size_t)N * p) : nullptr;
  a.dist_out = dD.ensure(N);
  a.X_out = dX.ensure(N);
  a.Ts_out = dTs.ensure(N);
  a.L = (int)N;
  const PhiloxKey key = make_key((uint32_t)seed, 0u, 7u);
  a.key0 = key.k0;
  a.key1 = key.k1;
  const int grid = (int)std::min<unsigned>((N + kFwdBlock - 1) / kFwdBlock, 4096u);
  const PosetProg pg = build_prog(f, 32, 8);
  constexpr size_t sm = fwd_smem_bytes<double, 32>();
  optin_smem(forward_dump_kernel<double, 32, false>, sm);
  forward_dump_kernel<double, 32, false><<<grid, kFwdBlock, sm, s>>>(a, pg);
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  if (given_times) {  // X_j = [T_j <= t_i] on the device (generate_genotypes_kernel)
    DevBuf<double> dGiven;
    DevBuf<int> dS;
    MCCBN_CUDA(cudaMemcpyAsync(dGiven.ensure(N), given_times, sizeof(double) * N, cudaMemcpyHostToDevice, s));
    dS.ensure((size_t)N * p);
    generate_genotypes_kernel<<<std::min(1024u, (N + 255u) / 256u), 256, 0, s>>>(dT.ptr, dGiven.ptr, dS.ptr, (int)N, p,
                                                                               1.0, 1, 0u, 0u);
    c->launches++;
    MCCBN_CUDA(cudaGetLastError());
    if (samples) MCCBN_CUDA(cudaMemcpyAsync(samples, dS.ptr, sizeof(int) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
    if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, dZ.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
    if (T_sum) MCCBN_CUDA(cudaMemcpyAsync(T_sum, dT.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
    MCCBN_CUDA(cudaStreamSynchronize(s));
    return;
  }
  std::vector<uint32_t> hx(N);
  std::vector<double> hts(N), hT;
  MCCBN_CUDA(cudaMemcpyAsync(hx.data(), dX.ptr, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaMemcpyAsync(hts.data(), dTs.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, dZ.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
  if (T_sum) MCCBN_CUDA(cudaMemcpyAsync(T_sum, dT.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  if (T_sampling) std::copy(hts.begin(), hts.end(), T_sampling);
  if (samples) {
    for (unsigned i = 0; i < N; ++i)
      for (int j = 0; j < p; ++j) samples[(size_t)j * N + i] = (hx[i] >> j) & 1u;
  }
}

int mccbn_sample_genotypes(mccbn_ctx* c, unsigned N, const int* poset, const double* lambda, double* T_sampling,
                           float lambda_s, int sampling_times_available, int seed, int p, int* samples, double* Tdiff,
                           char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (sampling_times_available) {
    // X_j = [T_j <= t_i] with the caller's per-sample times
    if (!T_sampling) throw Error(MCCBN_ERR_ARG, "sampling times expected");
    simulate_forward(c, N, poset, lambda, p, lambda_s, seed, nullptr, samples, Tdiff, nullptr, T_sampling);
  } else {
    simulate_forward(c, N, poset, lambda, p, lambda_s, seed, T_sampling, samples, Tdiff, nullptr);
  }
  MCCBN_CATCH
}

int mccbn_sample_times(mccbn_ctx* c, unsigned N, const int* poset, const double* lambda, int seed, int p,
                       double* mutation_times, double* Tdiff, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  simulate_forward(c, N, poset, lambda, p, 1.0f, seed, nullptr, nullptr, Tdiff, mutation_times);
  MCCBN_CATCH
}

int mccbn_generate_genotypes(mccbn_ctx* c, const double* T_events_sum, const int* poset, double* T_sampling,
                             float lambda_s, int sampling_times_available, int seed, int N, int p, int* samples,
                             char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);  // the reference validates the poset although the comparison does not use it
  if (N <= 0) return MCCBN_OK;
  cudaStream_t s = c->stream;
  DevBuf<double> dT, dTs;
  DevBuf<int> dS;
  dT.ensure((size_t)N * p);
  dTs.ensure(N);
  dS.ensure((size_t)N * p);
  MCCBN_CUDA(cudaMemcpyAsync(dT.ptr, T_events_sum, sizeof(double) * (size_t)N * p, cudaMemcpyHostToDevice, s));
  if (sampling_times_available)
    MCCBN_CUDA(cudaMemcpyAsync(dTs.ptr, T_sampling, sizeof(double) * N, cudaMemcpyHostToDevice, s));
  const PhiloxKey key = make_key((uint32_t)seed, 0u, 8u);
  generate_genotypes_kernel<<<std::min(1024, (N + 255) / 256), 256, 0, s>>>(dT.ptr, dTs.ptr, dS.ptr, N, p, (double)lambda_s,
                                                                          sampling_times_available, key.k0, key.k1);
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  if (samples) MCCBN_CUDA(cudaMemcpyAsync(samples, dS.ptr, sizeof(int) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
  if (T_sampling && !sampling_times_available)
    MCCBN_CUDA(cudaMemcpyAsync(T_sampling, dTs.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  MCCBN_CATCH
}

// ---- indexing seams ---------------------------------------------------------------------------------------------
int mccbn_flatten_poset(const int* poset, int p, int* topo, uint64_t* parent_mask, char* err, size_t errlen) {
  MCCBN_TRY
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  for (int k = 0; k < p; ++k) topo[k] = f.topo[k];
  for (int j = 0; j < p; ++j) parent_mask[j] = f.parent_ev[j];
  MCCBN_CATCH
}

int mccbn_neighbors(int p, int k, const int* genotype, int* out, int* ring, int* nrows_sum, char* err,
                    size_t errlen) {
  MCCBN_TRY
  if (p < 1 || p > kMaxP || k < 0 || k > p) throw Error(MCCBN_ERR_ARG, "invalid p or k");
  std::vector<int> rg;
  std::vector<uint64_t> masks = hamming_ball_masks((unsigned)p, (unsigned)k, &rg);
  const int n = (int)masks.size();
  if (nrows_sum) *nrows_sum = n;
  if (out)
    for (int r = 0; r < n; ++r)
      for (int j = 0; j < p; ++j) out[(size_t)j * n + r] = ((genotype[j] != 0) != (bool)((masks[r] >> j) & 1)) ? 1 : 0;
  if (ring) std::copy(rg.begin(), rg.end(), ring);
  MCCBN_CATCH
}

int mccbn_compatible_observations(mccbn_ctx* c, const int* obs, const int* poset, int N, int p, int* out,
                                  int* num_compatible, int* num_incompatible_events, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(obs, nullptr, nullptr);
  DevPoset dp;
  fill_dev_poset(f, dp);
  cudaStream_t s = c->stream;
  MCCBN_CUDA(cudaMemcpyAsync(pb.posets.ptr, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  DevBuf<int> flags;
  DevBuf<unsigned long long> counts;
  flags.ensure(N);
  counts.ensure(2);
  MCCBN_CUDA(cudaMemsetAsync(counts.ptr, 0, 2 * sizeof(unsigned long long), s));
  compat_kernel<<<dim3(std::min(1024, (N + 255) / 256), 1), 256, 0, s>>>(pb.bits.ptr, N, pb.posets.ptr, flags.ptr,
                                                                          counts.ptr);
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  unsigned long long h[2];
  if (out) MCCBN_CUDA(cudaMemcpyAsync(out, flags.ptr, sizeof(int) * N, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaMemcpyAsync(h, counts.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  if (num_compatible) *num_compatible = (int)h[0];
  if (num_incompatible_events) *num_incompatible_events = (int)h[1];
  MCCBN_CATCH
}


int mccbn_initialize_lambda(mccbn_ctx* c, const int* obs, const int* poset, int N, int p, float lambda_s,
                            float max_lambda, double* lambda_out, double* eps_out, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(obs, nullptr, nullptr);
  DevPoset dp;
  fill_dev_poset(f, dp);
  cudaStream_t s = c->stream;
  MCCBN_CUDA(cudaMemcpyAsync(pb.posets.ptr, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  DevBuf<unsigned long long> cnt, counts;
  cnt.ensure(128);
  counts.ensure(2);
  MCCBN_CUDA(cudaMemsetAsync(cnt.ptr, 0, 128 * sizeof(unsigned long long), s));
  MCCBN_CUDA(cudaMemsetAsync(counts.ptr, 0, 2 * sizeof(unsigned long long), s));
  const int gx = std::min(1024, (N + 255) / 256);
  init_lambda_counts_kernel<<<dim3(gx, 1), 256, 0, s>>>(pb.bits.ptr, N, pb.posets.ptr, cnt.ptr);
  compat_kernel<<<dim3(gx, 1), 256, 0, s>>>(pb.bits.ptr, N, pb.posets.ptr, nullptr, counts.ptr);
  c->launches += 2;
  MCCBN_CUDA(cudaGetLastError());
  unsigned long long hc[128], h2[2];
  MCCBN_CUDA(cudaMemcpyAsync(hc, cnt.ptr, sizeof(hc), cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaMemcpyAsync(h2, counts.ptr, sizeof(h2), cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  lambda_from_counts(f, hc, N, lambda_s, max_lambda, lambda_out);
  if (eps_out) {
    const double e = (double)(int)h2[1] / ((double)N * p);
    *eps_out = std::max(std::min(e, 1 - DBL_EPSILON), DBL_EPSILON);
  }
  MCCBN_CATCH
}

// ---- supplied-randomness seams ----------------------------------------------------------------------------------
int mccbn_forward_from_times(mccbn_ctx* c, const int* genotype, int L, const int* poset, int p, double eps,
                             const double* Z, const double* T_sampling, double* T_sum, int* X, int* dist, double* w,
                             char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  cudaStream_t s = c->stream;
  DevPoset dp;
  fill_dev_poset(f, dp);
  DevBuf<DevPoset> dpos;
  DevBuf<double> dZ, dT, dTs, dW, dWt;
  DevBuf<int> dX, dD;
  dpos.ensure(1);
  const size_t LP = (size_t)L * p;
  MCCBN_CUDA(cudaMemcpyAsync(dpos.ptr, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dZ.ensure(LP), Z, sizeof(double) * LP, cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dTs.ensure(L), T_sampling, sizeof(double) * L, cudaMemcpyHostToDevice, s));
  const std::vector<double> wt = host_wtab(eps, p);
  MCCBN_CUDA(cudaMemcpyAsync(dWt.ensure(65), wt.data(), sizeof(double) * 65, cudaMemcpyHostToDevice, s));
  dT.ensure(LP);
  const int grid = std::min(2048, (L + 127) / 128);
  propagate_times_kernel<<<grid, 128, 0, s>>>(dZ.ptr, dT.ptr, L, dpos.ptr);
  forward_from_times_kernel<<<grid, 128, 0, s>>>(dT.ptr, dTs.ptr, pack_genotype(genotype, p), L, p, dWt.ptr,
                                                 X ? dX.ensure(LP) : nullptr, dD.ensure(L), dW.ensure(L));
  c->launches += 2;
  MCCBN_CUDA(cudaGetLastError());
  if (T_sum) MCCBN_CUDA(cudaMemcpyAsync(T_sum, dT.ptr, sizeof(double) * LP, cudaMemcpyDeviceToHost, s));
  if (X) MCCBN_CUDA(cudaMemcpyAsync(X, dX.ptr, sizeof(int) * LP, cudaMemcpyDeviceToHost, s));
  if (dist) MCCBN_CUDA(cudaMemcpyAsync(dist, dD.ptr, sizeof(int) * L, cudaMemcpyDeviceToHost, s));
  if (w) MCCBN_CUDA(cudaMemcpyAsync(w, dW.ptr, sizeof(double) * L, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  MCCBN_CATCH
}

int mccbn_estep_forward_from_times(mccbn_ctx* c, const int* obs, int N, const int* poset, int p,
                                   const double* lambda, double eps, const double* weights, const double* times,
                                   int sampling_times_available, int L, const double* Z, const double* T_sampling,
                                   float max_lambda, double* w_sum, double* w_sum_sq, double* dist, double* Tdiff,
                                   double* totals, double* eps_new, double* lambda_new, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(obs, times, weights);
  pb.ensure_trace(1);
  pb.set_model(0, f, lambda, eps, 1.0f, max_lambda);
  cudaStream_t s = c->stream;
  DevBuf<double> dZ, dT, dTs, dWt;
  const size_t LP = (size_t)L * p;
  MCCBN_CUDA(cudaMemcpyAsync(dZ.ensure(LP), Z, sizeof(double) * LP, cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dTs.ensure(L), T_sampling, sizeof(double) * L, cudaMemcpyHostToDevice, s));
  const std::vector<double> wt = host_wtab(eps, p);
  MCCBN_CUDA(cudaMemcpyAsync(dWt.ensure(65), wt.data(), sizeof(double) * 65, cudaMemcpyHostToDevice, s));
  dT.ensure(LP);
  propagate_times_kernel<<<std::min(2048, (L + 127) / 128), 128, 0, s>>>(dZ.ptr, dT.ptr, L, pb.posets.ptr);
  pb.part.ensure((size_t)N * pb.PS());
  estep_from_times_kernel<<<std::min(N, c->sm_count * 8), kF64Block, 0, s>>>(
      dZ.ptr, dT.ptr, dTs.ptr, pb.bits.ptr, pb.times.ptr, sampling_times_available, N, L, pb.posets.ptr, dWt.ptr,
      pb.part.ptr, pb.PS());
  c->launches += 2;
  MCCBN_CUDA(cudaGetLastError());
  pb.run_finalize(0, 1, N, L, true);
  // totals before the M-step overwrites nothing: stats = [LL, N_eff, D, zero, 0, S_k (topological order)]
  std::vector<double> st(pb.PS());
  MCCBN_CUDA(cudaMemcpyAsync(st.data(), pb.stats_of(0), sizeof(double) * pb.PS(), cudaMemcpyDeviceToHost, s));
  pb.run_mstep(0, 1);
  DevModel dm;
  pb.read_model(0, dm);
  if (dm.comm_error) throw Error(MCCBN_ERR_CUDA, kMsgCommTimeout);
  if (dm.zero_weight) throw Error(MCCBN_ERR_ZERO_WEIGHT, kMsgZeroWeight);
  if (w_sum) MCCBN_CUDA(cudaMemcpyAsync(w_sum, pb.out_w.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (w_sum_sq) MCCBN_CUDA(cudaMemcpyAsync(w_sum_sq, pb.out_w2.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (dist) MCCBN_CUDA(cudaMemcpyAsync(dist, pb.out_dist.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, pb.out_T.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  if (totals) {
    totals[0] = st[0];
    totals[1] = st[2];
    totals[2] = st[1];
    for (int k = 0; k < p; ++k) totals[3 + f.topo[k]] = st[kStatHdr + k];
  }
  if (eps_new) *eps_new = dm.eps;
  if (lambda_new) std::memcpy(lambda_new, dm.lambda, sizeof(double) * p);
  MCCBN_CATCH
}

// Pool scheme from a supplied pool: the supplied-randomness seam of SURVEY 8a-a9.  The pool branch of importance_weight
// (src/mcem_hcbn.cpp:460-491) is, given the pool, deterministic up to its L-index resampling: q_k = eps^d (1-eps)^(p-d),
// every resampled row carries the weight sum(q)/K, so obs_llhood_i = log(sum_k q_k / K); the resampled means estimate
// sum_k q_k g_k / sum_k q_k, which is what `dist` / `Tdiff` return (fp64, host pow table: same libm expression as :463-464).
int mccbn_pool_from_times(mccbn_ctx* c, const int* obs, int N, const int* poset, int p, double eps,
                          const double* weights, const double* times, int sampling_times_available, int K,
                          const double* T_pool, const double* Tdiff_pool, const double* T_sampling, int* X_pool,
                          int* d_pool, double* q_sum, double* dist, double* Tdiff, double* llhood, char* err,
                          size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (N < 1 || K < 1) throw Error(MCCBN_ERR_ARG, "pool_from_times: N and K must be >= 1");
  if (!sampling_times_available && !T_sampling) throw Error(MCCBN_ERR_ARG, "pool_from_times: sampling times expected");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(obs, times, weights);
  pb.ensure_trace(1);
  std::vector<double> lam1(p, 1.0);
  pb.set_model(0, f, lam1.data(), eps, 1.0f, 1e6f);  // only the structure and eps matter here
  cudaStream_t s = c->stream;
  DevBuf<double> dZ, dT, dTs, dWt;
  DevBuf<int> dD, dX, dD0;
  const size_t KP = (size_t)K * p;
  MCCBN_CUDA(cudaMemcpyAsync(dZ.ensure(KP), Tdiff_pool, sizeof(double) * KP, cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dT.ensure(KP), T_pool, sizeof(double) * KP, cudaMemcpyHostToDevice, s));
  dTs.ensure(K);
  if (T_sampling) MCCBN_CUDA(cudaMemcpyAsync(dTs.ptr, T_sampling, sizeof(double) * K, cudaMemcpyHostToDevice, s));
  const std::vector<double> wt = host_wtab(eps, p);
  MCCBN_CUDA(cudaMemcpyAsync(dWt.ensure(65), wt.data(), sizeof(double) * 65, cudaMemcpyHostToDevice, s));
  const int gx = std::min(2048, (K + 127) / 128);
  if (d_pool) {
    pool_dist_kernel<<<dim3(gx, N), 128, 0, s>>>(dT.ptr, dTs.ptr, pb.bits.ptr, pb.times.ptr, sampling_times_available, N,
                                                 K, p, dD.ensure((size_t)K * N));
    c->launches++;
  }
  if (X_pool) {  // genotypes of the pool at observation 0's sampling times (the same for all observations unless times are given)
    DevBuf<double> dTs0;
    const double* ts0 = dTs.ptr;
    if (sampling_times_available) {
      std::vector<double> t0((size_t)K, times[0]);
      MCCBN_CUDA(cudaMemcpyAsync(dTs0.ensure(K), t0.data(), sizeof(double) * K, cudaMemcpyHostToDevice, s));
      MCCBN_CUDA(cudaStreamSynchronize(s));
      ts0 = dTs0.ptr;
    }
    forward_from_times_kernel<<<gx, 128, 0, s>>>(dT.ptr, ts0, 0ull, K, p, dWt.ptr, dX.ensure(KP), dD0.ensure(K), nullptr);
    c->launches++;
    MCCBN_CUDA(cudaMemcpyAsync(X_pool, dX.ptr, sizeof(int) * KP, cudaMemcpyDeviceToHost, s));
    MCCBN_CUDA(cudaStreamSynchronize(s));
  }
  pb.part.ensure((size_t)N * pb.PS());
  estep_from_times_kernel<<<std::min(N, c->sm_count * 8), kF64Block, 0, s>>>(
      dZ.ptr, dT.ptr, dTs.ptr, pb.bits.ptr, pb.times.ptr, sampling_times_available, N, K, pb.posets.ptr, dWt.ptr,
      pb.part.ptr, pb.PS());
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  pb.run_finalize(0, 1, N, K, true);  // L_eff = K: obs_llhood_i = log(sum q / K)
  pb.run_mstep(0, 0);
  DevModel dm;
  pb.read_model(0, dm);
  if (dm.zero_weight) throw Error(MCCBN_ERR_ZERO_WEIGHT, kMsgZeroWeight);
  if (d_pool) MCCBN_CUDA(cudaMemcpyAsync(d_pool, dD.ptr, sizeof(int) * (size_t)K * N, cudaMemcpyDeviceToHost, s));
  if (q_sum) MCCBN_CUDA(cudaMemcpyAsync(q_sum, pb.out_w.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (dist) MCCBN_CUDA(cudaMemcpyAsync(dist, pb.out_dist.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, pb.out_T.ptr, sizeof(double) * (size_t)N * p, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  if (llhood) *llhood = dm.llhood;
  MCCBN_CATCH
}

int mccbn_conditional_from_uniforms(mccbn_ctx* c, const int* X, int L, const int* poset, int p,
                                    const double* lambda, const double* u, const double* T_sampling, double* Tdiff,
                                    double* log_q, double* log_pz, int* incompatible, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  cudaStream_t s = c->stream;
  DevPoset dp;
  fill_dev_poset(f, dp);
  DevBuf<DevPoset> dpos;
  DevBuf<double> dU, dTs, dLam, dT, dq, dpz;
  DevBuf<int> dX, dInc;
  const size_t LP = (size_t)L * p;
  MCCBN_CUDA(cudaMemcpyAsync(dpos.ensure(1), &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dX.ensure(LP), X, sizeof(int) * LP, cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dU.ensure(LP), u, sizeof(double) * LP, cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dTs.ensure(L), T_sampling, sizeof(double) * L, cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dLam.ensure(p), lambda, sizeof(double) * p, cudaMemcpyHostToDevice, s));
  conditional_from_uniforms_kernel<<<std::min(2048, (L + 127) / 128), 128, 0, s>>>(
      dX.ptr, dU.ptr, dTs.ptr, L, dpos.ptr, dLam.ptr, dT.ensure(LP), dq.ensure(L), dpz.ensure(L), dInc.ensure(L),
      SeamRng{});
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  if (Tdiff) MCCBN_CUDA(cudaMemcpyAsync(Tdiff, dT.ptr, sizeof(double) * LP, cudaMemcpyDeviceToHost, s));
  if (log_q) MCCBN_CUDA(cudaMemcpyAsync(log_q, dq.ptr, sizeof(double) * L, cudaMemcpyDeviceToHost, s));
  if (log_pz) MCCBN_CUDA(cudaMemcpyAsync(log_pz, dpz.ptr, sizeof(double) * L, cudaMemcpyDeviceToHost, s));
  if (incompatible) MCCBN_CUDA(cudaMemcpyAsync(incompatible, dInc.ptr, sizeof(int) * L, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  MCCBN_CATCH
}

// ---- the reference's remaining small export seams ----------------------------------------------------
// Boost's topological_sort (DFS from vertex 0.., out-edges by ascending target, post-order) as the reference uses it
// (src/model_impl.cpp:51-69): returns the vertices in TOPOLOGICAL order (= topo_path.rbegin() .. rend()).
static std::vector<int> boost_like_topological_order(const int* poset, int p) {
  std::vector<std::vector<int>> out(p);
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j)
      if (poset[(size_t)j * p + i] == 1) out[i].push_back(j);
  std::vector<int> post;
  std::vector<char> seen(p, 0);
  std::function<void(int)> dfs = [&](int v) {
    seen[v] = 1;
    for (int w : out[v])
      if (!seen[w]) dfs(w);
    post.push_back(v);
  };
  for (int v = 0; v < p; ++v)
    if (!seen[v]) dfs(v);
  std::reverse(post.begin(), post.end());
  return post;
}

// `_scale_path_to_mutation` (src/add_remove.cpp:17-37, :218-242): host arithmetic, O(p + |E|)
int mccbn_scale_path_to_mutation(const int* poset, int p, const double* lambda, double* out, char* err,
                                 size_t errlen) {
  MCCBN_TRY
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  std::vector<double> sc(p, 0.0);  // by topological position
  for (int k = 0; k < p; ++k) {
    double mx = -1.0;
    for (uint64_t m = f.parent_pos[k]; m; m &= m - 1) mx = std::max(mx, sc[ctz64(m)]);
    sc[k] = 1.0 / lambda[f.topo[k]] + (mx > -1.0 ? mx : 0.0);
    out[f.topo[k]] = sc[k];
  }
  MCCBN_CATCH
}

// `_cbn_density_log` (src/add_remove.cpp:206-216, :323-335): log-density of the waiting times, per row
int mccbn_cbn_density_log(const double* time, int N, int p, const double* lambda, double* out, char* err,
                          size_t errlen) {
  MCCBN_TRY
  double sumlog = 0.0;
  for (int j = 0; j < p; ++j) sumlog += std::log(lambda[j]);
  for (int i = 0; i < N; ++i) {
    double dot = 0.0;
    for (int j = 0; j < p; ++j) dot += time[(size_t)j * N + i] * lambda[j];
    out[i] = sumlog - dot;
  }
  MCCBN_CATCH
}

// `_complete_log_likelihood` (src/mcem_hcbn.cpp:102-136, :738-760): weighted complete-data log-likelihood
int mccbn_complete_log_likelihood(const double* lambda, int p, double eps, const double* Tdiff, int N,
                                  const double* dist, const double* weights, double* out, char* err, size_t errlen) {
  MCCBN_TRY
  double wsum = 0.0, sumlog = 0.0;
  for (int i = 0; i < N; ++i) wsum += weights[i];
  for (int j = 0; j < p; ++j) sumlog += std::log(lambda[j]);
  double llhood = wsum * sumlog;
  for (int j = 0; j < p; ++j) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += weights[i] * Tdiff[(size_t)j * N + i];
    llhood -= lambda[j] * s;
  }
  double wd = 0.0, wpd = 0.0;
  for (int i = 0; i < N; ++i) {
    wd += weights[i] * dist[i];
    wpd += weights[i] * ((double)p - dist[i]);
  }
  if (eps == 0.0) {  // only observations at a non-zero distance contribute (:120-126)
    for (int i = 0; i < N; ++i)
      if (dist[i] != 0.0)
        llhood += std::log(eps + DBL_EPSILON) * weights[i] * dist[i] +
                  std::log(1 - eps - DBL_EPSILON) * weights[i] * ((double)p - dist[i]);
  } else {
    llhood += std::log(eps) * wd + std::log1p(-eps) * wpd;
  }
  *out = llhood;
  MCCBN_CATCH
}

// `_draw_sample` (src/add_remove.cpp:109-141, :244-284): one add / remove / stand-still proposal.  The reference's
// own generator (mt19937 through the ranlux24 -> seed_seq chain, std::discrete_distribution) is used, so the seam is
// bit-identical to the reference for a given seed.
int mccbn_draw_sample(const int* genotype, const int* poset, int p, unsigned move, const double* q_prob, int seed,
                      int* sample, double* q_choice, char* err, size_t errlen) {
  MCCBN_TRY
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  uint64_t g = 0;
  for (int j = 0; j < p; ++j)
    if (genotype[j]) g |= 1ull << j;
  // closures by event id
  std::vector<uint64_t> anc(p, 0), desc(p, 0), par(p, 0);
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j)
      if (poset[(size_t)j * p + i] == 1) par[j] |= 1ull << i;
  for (int k = 0; k < p; ++k) {
    const int v = f.topo[k];
    anc[v] = 1ull << v;
    for (uint64_t m = par[v]; m; m &= m - 1) anc[v] |= anc[ctz64(m)];
  }
  for (int k = p - 1; k >= 0; --k) {
    const int v = f.topo[k];
    desc[v] |= 1ull << v;
    for (uint64_t m = par[v]; m; m &= m - 1) desc[ctz64(m)] |= desc[v];
  }
  bool compatible = true;
  uint64_t up = 0, down = 0;
  for (int v = 0; v < p; ++v)
    if ((g >> v) & 1ull) {
      up |= anc[v];
      if (par[v] & ~g) compatible = false;
      if ((anc[v] & ~g) == 0) down |= 1ull << v;
    }
  std::vector<double> rw(p), aw(p);
  double rs = 0.0, as = 0.0;
  for (int j = 0; j < p; ++j) {
    rw[j] = ((g >> j) & 1ull) ? q_prob[j] : 0.0;
    aw[j] = ((g >> j) & 1ull) ? 0.0 : 1.0 / q_prob[j];
    rs += rw[j];
    as += aw[j];
  }
  std::mt19937 rng = make_std_rng((unsigned)seed);
  std::discrete_distribution<int> Dr(rw.begin(), rw.end()), Da(aw.begin(), aw.end());
  const int idx_remove = Dr(rng);
  const int idx_add = Da(rng);
  uint64_t x = g;
  double q = 1.0;
  if (move == 0) {
    q = aw[idx_add] / as;
    x = (compatible ? g : up) | anc[idx_add];
  } else if (move == 1) {
    q = rw[idx_remove] / rs;
    x = (compatible ? g : down) & ~desc[idx_remove];
  }
  for (int j = 0; j < p; ++j) sample[j] = (int)((x >> j) & 1ull);
  *q_choice = q;
  MCCBN_CATCH
}
}  // extern "C"

namespace mccbn {
// coin_tossing (src/backward_sampling.cpp:64-77) with the E-step's own Philox coins: row l flips bit j with probability eps
__global__ void coin_tossing_kernel(const int* __restrict__ genotype, int* __restrict__ out, int L, int p, uint32_t thr,
                                    uint32_t key0, uint32_t key1) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
    uint4 rc = make_uint4(0, 0, 0, 0);
    for (int j = 0; j < p; ++j) {
      if ((j & 3) == 0) rc = philox4x32_10((uint32_t)l, 0u, 64u + (uint32_t)(j >> 2), 0u, key0, key1);
      const uint32_t bits = (j & 3) == 0 ? rc.x : ((j & 3) == 1 ? rc.y : ((j & 3) == 2 ? rc.z : rc.w));
      out[(size_t)j * L + l] = (bits < thr) ? !genotype[j] : genotype[j];
    }
  }
}
}  // namespace mccbn

extern "C" {
// `_coin_tossing` (src/backward_sampling.cpp:105-123): L noisy copies of a genotype, L x p column-major
int mccbn_coin_tossing(mccbn_ctx* c, double eps, unsigned L, const int* genotype, int p, int seed, int* out, char* err,
                       size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (p < 1 || p > MCCBN_MAX_P) throw Error(MCCBN_ERR_ARG, "1 <= p <= 64 expected");
  if (L == 0) return MCCBN_OK;
  cudaStream_t s = c->stream;
  DevBuf<int> dg, dout;
  MCCBN_CUDA(cudaMemcpyAsync(dg.ensure(p), genotype, sizeof(int) * p, cudaMemcpyHostToDevice, s));
  dout.ensure((size_t)L * p);
  const uint32_t thr = (uint32_t)std::min(std::max(eps, 0.0) * 4294967296.0, 4294967295.0);
  const PhiloxKey key = make_key((uint32_t)seed, 0u, 2u);
  coin_tossing_kernel<<<std::min(1024u, (L + 255u) / 256u), 256, 0, s>>>(dg.ptr, dout.ptr, (int)L, p, thr, key.k0, key.k1);
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  MCCBN_CUDA(cudaMemcpyAsync(out, dout.ptr, sizeof(int) * (size_t)L * p, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  MCCBN_CATCH
}

// `_generate_mutation_times` (src/add_remove.cpp:286-321): conditional proposal for N given genotypes.  The uniforms
// come from the reference's generator in the reference's order (T_s first unless given, then N draws per node in
// topological order); the arithmetic runs in fp64 on the device (conditional_from_uniforms_kernel).
int mccbn_generate_mutation_times(mccbn_ctx* c, const int* obs, const int* poset, const double* lambda, double* time,
                                  float lambda_s, int sampling_times_available, int seed, int N, int p, double* Tdiff,
                                  double* proposal_density, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  if (N <= 0) return MCCBN_OK;
  std::mt19937 rng = make_std_rng((unsigned)seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<double> ts(N), u((size_t)N * p);
  for (int n = 0; n < N; ++n) {
    if (sampling_times_available) {
      ts[n] = time[n];
    } else {
      ts[n] = -std::log1p(U(rng) * std::expm1(-INFINITY * (double)lambda_s)) / (double)lambda_s;  // rexp, rng_utils.cpp:44-56
      time[n] = ts[n];
    }
  }
  for (int v : boost_like_topological_order(poset, p))
    for (int n = 0; n < N; ++n) u[(size_t)v * N + n] = U(rng);
  std::vector<double> lq(N);
  const int rc = mccbn_conditional_from_uniforms(c, obs, N, poset, p, lambda, u.data(), ts.data(), Tdiff, lq.data(),
                                                 nullptr, nullptr, err, errlen);
  if (rc != MCCBN_OK) return rc;
  if (proposal_density) std::copy(lq.begin(), lq.end(), proposal_density);
  MCCBN_CATCH
}

// ---- roofline helper: pure Philox4x32-10 generation rate of this GPU (the INT32 ceiling of the sampling kernels) --
}  // extern "C"
namespace mccbn {
__global__ void __launch_bounds__(256) philox_peak_kernel(uint32_t* __restrict__ out, int iters, uint32_t k0,
                                                          uint32_t k1) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0u;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    const uint4 r = philox4x32_10((uint32_t)i, tid, 0u, 0u, k0, k1);
    acc ^= r.x ^ r.y ^ r.z ^ r.w;
  }
  out[tid] = acc;
}
}  // namespace mccbn
extern "C" {
// returns 32-bit random words generated per second (best of `reps` launches), 0 on failure
double mccbn_microbench_philox(mccbn_ctx* c, int reps) {
  try {
    c = get_ctx(c);
    cudaStream_t s = c->stream;
    const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
    DevBuf<uint32_t> out;
    out.ensure((size_t)blocks * threads);
    cudaEvent_t e0, e1;
    MCCBN_CUDA(cudaEventCreate(&e0));
    MCCBN_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
      MCCBN_CUDA(cudaEventRecord(e0, s));
      philox_peak_kernel<<<blocks, threads, 0, s>>>(out.ptr, iters, 1234u, 5678u + r);
      MCCBN_CUDA(cudaEventRecord(e1, s));
      MCCBN_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      MCCBN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double words = 4.0 * iters * (double)blocks * threads;
      if (r > 0) best = std::max(best, words / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
  } catch (...) {
    return 0.0;
  }
}

// ---- poset-specialised kernels (NVRTC) ---------------------------------------------------------------------------
int mccbn_set_jit(mccbn_ctx* c, int mode) {
  char* err = nullptr;
  size_t errlen = 0;
  MCCBN_TRY
  const int m = mode < 0 ? -1 : (mode > 0 ? 1 : 0);
  if (!c) {  // the default context and the in-process team behind it
    std::lock_guard<std::mutex> lk(g_default_mu);
    g_team_jit_mode = m;
    if (g_team)
      for (mccbn_ctx* t : g_team->ctx) t->jit_mode = m;
  }
  c = get_ctx(c);
  c->jit_mode = m;
  MCCBN_CATCH
}

// host-only: generate + compile the specialised forward kernel of a poset (no device needed); log receives the
// NVRTC / ptxas -v output, source (optional) the generated CUDA text
int mccbn_jit_compile_forward(const int* poset, int p, int times_given, size_t* cubin_bytes, double* compile_ms,
                              char* log, size_t loglen, char* source, size_t sourcelen, char* err, size_t errlen) {
  MCCBN_TRY
  if (p > kWideMaxP) throw Error(MCCBN_ERR_ARG, "the Philox E-step supports p <= 64");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  static NvrtcApi rtc;
  const std::string src = generate_forward_source(f, times_given != 0, p > kFastMaxP ? 2 : jit_default_minb());
  if (source && sourcelen) std::snprintf(source, sourcelen, "%s", src.c_str());
  JitCompiled cc;
  const bool ok = jit_compile(rtc, src, cc);
  if (log && loglen) std::snprintf(log, loglen, "%s", cc.log.c_str());
  if (cubin_bytes) *cubin_bytes = cc.cubin.size();
  if (compile_ms) *compile_ms = cc.compile_ms;
  if (!ok) throw Error(MCCBN_ERR_CUDA, "NVRTC compilation failed: " + cc.log);
  MCCBN_CATCH
}

// the same for the conditional-proposal kernels; mode: 0 backward, 1 bernoulli, 2 OT-CBN, 3 add-remove
int mccbn_jit_compile_conditional(const int* poset, int p, int times_given, int mode, size_t* cubin_bytes,
                                  double* compile_ms, char* log, size_t loglen, char* source, size_t sourcelen,
                                  char* err, size_t errlen) {
  MCCBN_TRY
  if (p > kWideMaxP || (p > kFastMaxP && mode == 3))
    throw Error(MCCBN_ERR_ARG, "the conditional kernels support p <= 64 (add-remove: p <= 32)");
  if (mode < 0 || mode > 3) throw Error(MCCBN_ERR_ARG, "mode must be 0..3");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  static NvrtcApi rtc;
  const std::string src = generate_cond_source(f, times_given != 0, mode, JitCache::cond_minb());
  if (source && sourcelen) std::snprintf(source, sourcelen, "%s", src.c_str());
  JitCompiled cc;
  const bool ok = jit_compile(rtc, src, cc);
  if (log && loglen) std::snprintf(log, loglen, "%s", cc.log.c_str());
  if (cubin_bytes) *cubin_bytes = cc.cubin.size();
  if (compile_ms) *compile_ms = cc.compile_ms;
  if (!ok) throw Error(MCCBN_ERR_CUDA, "NVRTC compilation failed: " + cc.log);
  MCCBN_CATCH
}

// ---- host-only test seam for the poset edit operations of the annealer (src/model_impl.cpp:145-316) --------------
int mccbn_poset_edit(const int* poset, int p, int op, int u, int v, int* poset_out, int* flags, char* err,
                     size_t errlen) {
  MCCBN_TRY
  if (p < 1 || p > kMaxP || u < 0 || v < 0 || u >= p || v >= p) throw Error(MCCBN_ERR_ARG, "invalid p, u or v");
  Poset P = Poset::from_adjacency(poset, p);
  if (!P.sort()) throw Error(MCCBN_ERR_NOT_ACYCLIC, kMsgNotAcyclic);
  P.set_children();
  bool added = false;
  switch (op) {
    case 0: added = P.add_edge(u, v); break;
    case 1: P.remove_edge(u, v); break;
    case 2: added = P.add_relation(u, v); break;
    case 3: P.remove_relation(u, v); break;
    case 4: P.transitive_reduction_dag(); break;
    case 5: P.swap_node(u, v); break;
    default: throw Error(MCCBN_ERR_ARG, "unknown poset edit operation");
  }
  if (flags) *flags = (added ? 1 : 0) | (P.cycle ? 2 : 0) | (P.reduction_flag ? 4 : 0);
  if (poset_out) P.to_adjacency(poset_out);
  MCCBN_CATCH
}

// ---- adaptive simulated annealing --------------------------------------------------------------------------------
}  // extern "C"
extern "C" int mccbn_adaptive_simulated_annealing(mccbn_ctx*, const int*, const int*, const double*, float, const double*,
                                                  unsigned, const char*, unsigned, unsigned, double, float, float, float,
                                                  float, int, unsigned, unsigned, int, const char*, int, int, int, int,
                                                  int, int, double*, double*, int*, double*, char*, size_t);
static int asa_entry(mccbn_ctx* c, const int* poset, const int* obs, const double* times, float lambda_s,
                     const double* weights, unsigned L, const char* sampling, unsigned max_iter_EM,
                     unsigned update_step_size, double tol, float max_lambda, float T0, float adap_rate,
                     float acceptance_rate, int step_size, unsigned max_iter_ASA, unsigned neighborhood_dist,
                     int adaptive, const char* outdir, int sampling_times_available, int thrds, int verbose, int seed,
                     int N, int p, double* lambda_out, double* eps_out, int* poset_out, double* llhood_out, char* err,
                     size_t errlen, bool structural_score) {
  (void)thrds;
  MCCBN_TRY
  if (!c && !structural_score) {
    if (Team* team = team_for(N)) {  // candidate-parallel annealing: every member is given ALL observations
      const int G = (int)team->ctx.size();
      std::vector<std::vector<double>> lam_r(G, std::vector<double>(p));
      std::vector<std::vector<int>> pos_r(G, std::vector<int>((size_t)p * p));
      std::vector<double> eps_r(G), ll_r(G);
      team_run(*team, [&](int r, mccbn_ctx* cr, char* e, size_t el) {
        return mccbn_adaptive_simulated_annealing(cr, poset, obs, times, lambda_s, weights, L, sampling, max_iter_EM,
                                                  update_step_size, tol, max_lambda, T0, adap_rate, acceptance_rate,
                                                  step_size, max_iter_ASA, neighborhood_dist, adaptive, outdir,
                                                  sampling_times_available, thrds, r == 0 ? verbose : 0, seed, N, p,
                                                  lam_r[r].data(), &eps_r[r], pos_r[r].data(), &ll_r[r], e, el);
      });
      std::copy(lam_r[0].begin(), lam_r[0].end(), lambda_out);
      std::copy(pos_r[0].begin(), pos_r[0].end(), poset_out);
      *eps_out = eps_r[0];
      *llhood_out = ll_r[0];
      return MCCBN_OK;
    }
  }
  c = get_ctx(c);
  const Scheme sc = parse_scheme(sampling);
  if (p > kWideMaxP) throw Error(MCCBN_ERR_ARG, "the Philox E-step supports p <= 64");
  Candidate M;
  {
    FlatPoset f;
    flatten_or_throw(poset, p, M.poset, f);
  }
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(obs, times, weights);
  if (N <= 65536) {  // moderate N: compatibility counts of the move generator on the host (engine_asa.cuh::count_compatible)
    pb.host_bits.resize((size_t)N);
    for (int i = 0; i < N; ++i) {
      uint64_t w = 0;
      for (int j = 0; j < p; ++j) w |= (uint64_t)(obs[(size_t)j * N + i] != 0) << j;
      pb.host_bits[i] = w;
    }
  }
  pb.test_structural_score = structural_score;
  // several GPUs: candidate-parallel -- every rank is given ALL observations and fits its share of each speculative batch
  pb.local_only = c->world > 1;
  initialize_candidate(pb, M, lambda_s, max_lambda);  // src/asa.cpp:640-642
  EmControl ctl;
  ctl.max_iter = max_iter_EM;
  ctl.update_step_size = update_step_size;
  ctl.tol = tol;
  ctl.max_lambda = max_lambda;
  ctl.neighborhood_dist = neighborhood_dist;
  AsaControl actl;
  actl.T = T0;
  actl.outdir = outdir ? outdir : "";
  actl.acceptance_rate = acceptance_rate;
  actl.adap_rate = adap_rate;
  actl.step_size = (unsigned)std::max(0, step_size);
  actl.max_iter = max_iter_ASA;
  actl.adaptive = adaptive != 0;
  // Speculative batch size.  A batch only pays when one candidate's E-step cannot fill the GPU (small N * L) and while
  // proposals are mostly rejected; on BASELINE config 4 (1e7 samples per E-step) batches of 1 / 8 / 15 take 5.4 / 10.6 /
  // 18.0 s for 100 steps.  batch <= 0 = automatic: an upper bound from the work per E-step, adapted to the observed
  // acceptance pattern inside simulated_annealing.  The result does not depend on it (tests/test_gpu_asa.py).
  int batch = 0;
  if (const char* b = std::getenv("MCCBN_ASA_BATCH")) batch = std::max(1, std::min(kMaxCand - 1, std::atoi(b)));
  if (batch <= 0) {
    const double work = (double)N * (double)L;  // samples per E-step of one candidate
    batch = -(int)std::max(1.0, std::min((double)(kMaxCand - 1), std::floor(2e6 / std::max(work, 1.0))));
    if (c->world > 1) batch = -std::min(kMaxCand - 1, std::max(-batch, c->world));  // one candidate per GPU at least
  }
  const double ll = simulated_annealing(pb, M, actl, L, sc, ctl, sampling_times_available != 0, (uint32_t)seed,
                                        lambda_s, verbose != 0, batch);
  std::copy(M.lambda.begin(), M.lambda.end(), lambda_out);
  *eps_out = M.eps;
  M.poset.to_adjacency(poset_out);
  *llhood_out = ll;
  MCCBN_CATCH
}
extern "C" {
int mccbn_adaptive_simulated_annealing(mccbn_ctx* c, const int* poset, const int* obs, const double* times,
                                       float lambda_s, const double* weights, unsigned L, const char* sampling,
                                       unsigned max_iter_EM, unsigned update_step_size, double tol, float max_lambda,
                                       float T0, float adap_rate, float acceptance_rate, int step_size,
                                       unsigned max_iter_ASA, unsigned neighborhood_dist, int adaptive,
                                       const char* outdir, int sampling_times_available, int thrds, int verbose,
                                       int seed, int N, int p, double* lambda_out, double* eps_out, int* poset_out,
                                       double* llhood_out, char* err, size_t errlen) {
  return asa_entry(c, poset, obs, times, lambda_s, weights, L, sampling, max_iter_EM, update_step_size, tol, max_lambda,
                   T0, adap_rate, acceptance_rate, step_size, max_iter_ASA, neighborhood_dist, adaptive, outdir,
                   sampling_times_available, thrds, verbose, seed, N, p, lambda_out, eps_out, poset_out, llhood_out, err,
                   errlen, false);
}

// TEST-ONLY seam (tests/test_gpu_asa.py): the annealer with every MCEM score replaced by the deterministic structural
// score  -1000 - num_incompatible_events + 3 |cover relations|  (the oracle's score_mode 1), so that the move generator,
// memo tables, Metropolis rule and temperature schedule can be compared with the oracle without Monte-Carlo noise.
int mccbn_test_asa_structural_score(mccbn_ctx* c, const int* poset, const int* obs, const double* times,
                                    float lambda_s, const double* weights, unsigned L, const char* sampling,
                                    unsigned max_iter_EM, unsigned update_step_size, double tol, float max_lambda,
                                    float T0, float adap_rate, float acceptance_rate, int step_size,
                                    unsigned max_iter_ASA, unsigned neighborhood_dist, int adaptive,
                                    const char* outdir, int sampling_times_available, int thrds, int verbose,
                                    int seed, int N, int p, double* lambda_out, double* eps_out, int* poset_out,
                                    double* llhood_out, char* err, size_t errlen) {
  return asa_entry(c, poset, obs, times, lambda_s, weights, L, sampling, max_iter_EM, update_step_size, tol, max_lambda,
                   T0, adap_rate, acceptance_rate, step_size, max_iter_ASA, neighborhood_dist, adaptive, outdir,
                   sampling_times_available, thrds, verbose, seed, N, p, lambda_out, eps_out, poset_out, llhood_out, err,
                   errlen, true);
}

// ---- legacy OT-CBN MCEM (src/mcem.cpp:149-210, :307-335) ---------------------------------------------------------
int mccbn_MCEM(mccbn_ctx* c, const double* ilambda, const int* poset, const int* O, const double* times, int max_iter,
               float alpha, const int* topo_path, const double* weights, int nrOfSamples, int verbose,
               float max_lambda, float lambda_s, int sampling_time_available, int seed, int N, int p,
               double* lambda_out, double* llhood_out, char* err, size_t errlen) {
  (void)topo_path;  // R passes my.topological.sort(poset) - 1; any valid order is equivalent, we derive our own
  MCCBN_TRY
  c = get_ctx(c);
  if (p > kWideMaxP) throw Error(MCCBN_ERR_ARG, "the Philox E-step supports p <= 64");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  mccbn_problem pb;
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(O, times, weights);
  cudaStream_t s = c->stream;
  // is_all_genotypes_compatible_with_poset (R/mcem_param_est.r:57-59, :75-87)
  {
    DevPoset dp;
    fill_dev_poset(f, dp);
    MCCBN_CUDA(cudaMemcpyAsync(pb.posets.ptr, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
    DevBuf<int> flags;
    flags.ensure(N);
    unsigned long long* cnt = pb.counts.ensure(2);
    MCCBN_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
    compat_kernel<<<dim3(std::min(1024, (N + 255) / 256), 1), 256, 0, s>>>(pb.bits.ptr, N, pb.posets.ptr, flags.ptr,
                                                                            cnt);
    c->launches++;
    std::vector<int> hf(N);
    MCCBN_CUDA(cudaMemcpyAsync(hf.data(), flags.ptr, sizeof(int) * N, cudaMemcpyDeviceToHost, s));
    MCCBN_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < N; ++i)
      if (!hf[i] && (!weights || weights[i] > 0))
        throw Error(MCCBN_ERR_ARG,
                    "Error in the function MCEM: Some genotypes with non-zero weights are incompatible with the poset!");
  }
  std::vector<double> lambda(ilambda, ilambda + p), avg_lambda(p, 0.0), T_colsum(p);
  float W = 0.f;  // `float W = sum(weights)` (src/mcem.cpp:160)
  {
    double sw = 0;
    for (int i = 0; i < N; ++i) sw += weights ? weights[i] : 1.0;
    W = (float)sw;
  }
  float avg_llhood = 0.f, llhood = 0.f;  // float in the reference (:163)
  const int record_iter = std::max((int)((1 - alpha) * max_iter), 1);
  pb.ensure_trace(1);
  pb.set_model(0, f, lambda.data(), 0.5, lambda_s, max_lambda);
  if (verbose) {
    std::printf("Initial value for rate parameters - lambda:");
    for (double l : lambda) std::printf(" %g", l);
    std::printf("\n");
  }
  std::vector<double> st(pb.PS());
  for (int iter = 0; iter < max_iter; ++iter) {
    MCCBN_CUDA(cudaMemcpyAsync(pb.models.ptr, lambda.data(), sizeof(double) * p, cudaMemcpyHostToDevice, s));
    pb.enqueue_iteration(0, kOtcbn, nrOfSamples, 0, sampling_time_available != 0, (uint32_t)seed, 0, false, 0);
    MCCBN_CUDA(cudaMemcpyAsync(st.data(), pb.stats_of(0), sizeof(double) * pb.PS(), cudaMemcpyDeviceToHost, s));
    MCCBN_CUDA(cudaStreamSynchronize(s));
    double sumlog = 0.0, dot = 0.0;
    for (int j = 0; j < p; ++j) {
      T_colsum[j] = st[kStatHdr + f.pos_of_event[j]];  // sum_i wt_i E[Z_ij]
      lambda[j] = W / T_colsum[j];
      if (lambda[j] > max_lambda) lambda[j] = max_lambda;  // :189-191
      sumlog += std::log(lambda[j]);
      dot += lambda[j] * T_colsum[j];
    }
    llhood = (float)(W * sumlog - dot);  // :193
    if (verbose) {
      std::printf("Lambda update:");
      for (double l : lambda) std::printf(" %g", l);
      std::printf(" - Log-likelihood: %g\n", llhood);
    }
    if (iter >= record_iter) {
      for (int j = 0; j < p; ++j) avg_lambda[j] += lambda[j];
      avg_llhood += llhood;
    }
  }
  // The reference's "sum(w) == 0 -> uniform weights" rule (src/mcem.cpp:123-127) has nothing to act on here: every
  // genotype with a non-zero weight was checked to be compatible with the poset above, so each of its draws has a
  // positive weight P(Z)/q(Z), and the per-observation log-sum-exp of the E-step cannot underflow to zero.  (Zero-weight
  // observations may be incompatible; they enter neither W nor T_colsum.)
  for (int j = 0; j < p; ++j) lambda_out[j] = avg_lambda[j] / (max_iter - record_iter);  // :206-207
  *llhood_out = (double)(avg_llhood / (max_iter - record_iter));
  MCCBN_CATCH
}

// ---- the OT-CBN callers of the conditional proposal ------------------------------------------------
// `drawHiddenVarsSamples` (src/mcem.cpp:214-248): N draws of the conditional proposal for ONE genotype; returns the
// waiting times T (N x p column-major) and the log proposal densities.  R's global generator (runif / rexp sugar)
// cannot be mirrored outside R: the draws come from Philox (53-bit uniforms), the arithmetic is the fp64 seam kernel.
int mccbn_drawHiddenVarsSamples(mccbn_ctx* c, int N, const int* O_vec, double time, const double* lambda,
                                const int* poset, const int* topo_path, float lambda_s,
                                int sampling_time_available, int seed, int p, double* T, double* densities,
                                double* time_out, char* err, size_t errlen) {
  (void)topo_path;  // any valid topological order gives the same distribution; we derive our own
  MCCBN_TRY
  c = get_ctx(c);
  if (N < 0) throw Error(MCCBN_ERR_ARG, "drawHiddenVarsSamples: N < 0");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  if (N == 0) return MCCBN_OK;
  cudaStream_t s = c->stream;
  DevPoset dp;
  fill_dev_poset(f, dp);
  DevBuf<DevPoset> dpos;
  DevBuf<double> dLam, dT, dq, dpz, dTs;
  const size_t NP = (size_t)N * p;
  MCCBN_CUDA(cudaMemcpyAsync(dpos.ensure(1), &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  MCCBN_CUDA(cudaMemcpyAsync(dLam.ensure(p), lambda, sizeof(double) * p, cudaMemcpyHostToDevice, s));
  SeamRng rng;
  std::memset(&rng, 0, sizeof(rng));
  rng.enabled = 1;
  rng.draw_ts = sampling_time_available ? 0 : 1;
  const PhiloxKey key = make_key((uint32_t)seed, 0u, 7u);
  rng.k0 = key.k0;
  rng.k1 = key.k1;
  rng.cw = 0u;
  rng.x_word = pack_genotype(O_vec, p);
  rng.time = time;
  rng.lambda_s = (double)lambda_s;
  rng.Ts_out = time_out ? dTs.ensure(N) : nullptr;
  conditional_from_uniforms_kernel<<<std::min(2048, (N + 127) / 128), 128, 0, s>>>(
      nullptr, nullptr, nullptr, N, dpos.ptr, dLam.ptr, dT.ensure(NP), dq.ensure(N), dpz.ensure(N), nullptr, rng);
  c->launches++;
  MCCBN_CUDA(cudaGetLastError());
  if (T) MCCBN_CUDA(cudaMemcpyAsync(T, dT.ptr, sizeof(double) * NP, cudaMemcpyDeviceToHost, s));
  if (densities) MCCBN_CUDA(cudaMemcpyAsync(densities, dq.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  if (time_out) MCCBN_CUDA(cudaMemcpyAsync(time_out, dTs.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  MCCBN_CATCH
}

// one OT-CBN E-step (noise-free model, conditional proposal) over N observations with per-observation outputs
static void otcbn_estep(mccbn_ctx* c, mccbn_problem& pb, const int* obs, const double* times, const double* weights,
                        const int* poset, const double* lambda, int nrOfSamples, float lambda_s,
                        int sampling_time_available, int seed, int N, int p) {
  if (p > kWideMaxP) throw Error(MCCBN_ERR_ARG, "the Philox E-step supports p <= 64");
  if (nrOfSamples < 1) throw Error(MCCBN_ERR_ARG, "nrOfSamples must be >= 1");
  Poset P;
  FlatPoset f;
  flatten_or_throw(poset, p, P, f);
  pb.ctx = c;
  pb.N = N;
  pb.p = p;
  pb.upload(obs, times, weights);
  pb.ensure_trace(1);
  pb.set_model(0, f, lambda, 0.5, lambda_s, 1e6f);
  pb.enqueue_iteration(0, kOtcbn, nrOfSamples, 0, sampling_time_available != 0, (uint32_t)seed, 0, true, 0);
}

// `expectedTimeDifference` (src/mcem.cpp:251-276 -> _expectedTimeDifference :105-146): self-normalised importance
// estimate of E[Z_j | genotype] from nrOfSamples conditional draws.  Same kernel as one observation of `MCEM`.
int mccbn_expectedTimeDifference(mccbn_ctx* c, int nrOfSamples, const int* O_vec, double time, const double* lambda,
                                 const int* poset, const int* topo_path, float lambda_s, int sampling_time_available,
                                 int seed, int p, double* time_difference, char* err, size_t errlen) {
  (void)topo_path;
  MCCBN_TRY
  c = get_ctx(c);
  mccbn_problem pb;
  otcbn_estep(c, pb, O_vec, &time, nullptr, poset, lambda, nrOfSamples, lambda_s, sampling_time_available, seed, 1, p);
  cudaStream_t s = c->stream;
  MCCBN_CUDA(cudaMemcpyAsync(time_difference, pb.out_T.ptr, sizeof(double) * p, cudaMemcpyDeviceToHost, s));
  DevModel dm;
  pb.read_model(0, dm);  // synchronises
  // Weights are kept in log space on the device, so a compatible genotype never loses all its weight to underflow (the
  // case the reference's uniform fallback, src/mcem.cpp:123-127, papers over); weight 0 for every draw means that the
  // genotype is incompatible with the poset -- the reference would return NaN waiting times here.
  if (dm.zero_weight) throw Error(MCCBN_ERR_ZERO_WEIGHT, "expectedTimeDifference: the genotype is incompatible with the poset (all importance weights are 0)");
  MCCBN_CATCH
}

// `loglike_importance_sampling` (R/mcem_genotype_probabilities.r:10-23) for ALL observations in one launch instead of
// one .Call("drawHiddenVarsSamples") per observation from an R loop:
//   probs[i] = weights[i] (logsumexp_l(log P(Z_l) - log q_l) - log nrOfSamples),  approx_loglike = sum_i probs[i]
int mccbn_loglike_importance_sampling(mccbn_ctx* c, const int* poset, const double* lambda, const int* obs_events,
                                      const double* sampling_times, int nrOfSamples, const double* weights,
                                      float lambda_s, int sampling_times_available, int seed, int N, int p,
                                      double* approx_loglike, double* probs, char* err, size_t errlen) {
  MCCBN_TRY
  c = get_ctx(c);
  if (N < 1) throw Error(MCCBN_ERR_ARG, "loglike_importance_sampling: no observations");
  mccbn_problem pb;
  otcbn_estep(c, pb, obs_events, sampling_times, weights, poset, lambda, nrOfSamples, lambda_s,
              sampling_times_available, seed, N, p);
  cudaStream_t s = c->stream;
  std::vector<double> lw(N);  // log sum_l w_il straight from the log-sum-exp of the partial records (no exp / log round trip)
  MCCBN_CUDA(cudaMemcpyAsync(lw.data(), pb.out_logw.ptr, sizeof(double) * N, cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  double total = 0.0;
  const double logn = std::log((double)nrOfSamples);
  for (int i = 0; i < N; ++i) {
    const double v = (weights ? weights[i] : 1.0) * (lw[i] - logn);
    if (probs) probs[i] = v;
    total += v;
  }
  if (approx_loglike) *approx_loglike = total;
  MCCBN_CATCH
}

}  // extern "C"
