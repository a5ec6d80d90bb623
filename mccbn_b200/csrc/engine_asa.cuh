// engine_asa.cuh -- adaptive simulated annealing (src/asa.cpp) and the legacy OT-CBN MCEM entry (src/mcem.cpp)
// on top of the device E-step.  Included by engine_impl.cuh (single translation unit).
#pragma once
#include "asa.cuh"

namespace mccbn {

// ---- device helpers used by the annealer ------------------------------------------------------------------------
// number of observations compatible with poset P (num_compatible_observations, compatible_observations.cpp:68-76)
static int count_compatible(mccbn_problem& pb, const Poset& P, int* num_incompatible_events = nullptr) {
  FlatPoset f;
  P.flatten(f);
  if (!pb.host_bits.empty()) {
    // moderate N: the same integer counts on the host from the bit-packed observations (a device round trip costs as
    // much, and the annealer generates its look-ahead proposals while the device is busy fitting the current candidate)
    int n_comp = 0, n_inc = 0;
    for (const uint64_t g : pb.host_bits) {
      int bad = 0;
      for (uint64_t m = g; m; m &= m - 1) bad += __builtin_popcountll(f.parent_ev[ctz64(m)] & ~g);
      n_inc += bad;
      n_comp += bad == 0;
    }
    if (num_incompatible_events) *num_incompatible_events = n_inc;
    return n_comp;
  }
  DevPoset dp;
  fill_dev_poset(f, dp);
  cudaStream_t s = pb.ctx->stream;
  const int slot = kMaxCand - 1;  // scratch slot
  MCCBN_CUDA(cudaMemcpyAsync(pb.posets.ptr + slot, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  unsigned long long* cnt = pb.counts.ensure(2);
  MCCBN_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s));
  compat_kernel<<<dim3(std::min(1024, (pb.N + 255) / 256), 1), 256, 0, s>>>(pb.bits.ptr, pb.N, pb.posets.ptr + slot,
                                                                              nullptr, cnt);
  pb.ctx->launches++;
  unsigned long long h[2];
  MCCBN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  if (num_incompatible_events) *num_incompatible_events = (int)h[1];
  return (int)h[0];
}

// initialize_lambda + eps start value for a candidate (src/asa.cpp:72-129, :256-258)
static void initialize_candidate(mccbn_problem& pb, Candidate& c, float lambda_s, float max_lambda) {
  FlatPoset f;
  c.poset.flatten(f);
  DevPoset dp;
  fill_dev_poset(f, dp);
  cudaStream_t s = pb.ctx->stream;
  const int slot = kMaxCand - 1;
  MCCBN_CUDA(cudaMemcpyAsync(pb.posets.ptr + slot, &dp, sizeof(dp), cudaMemcpyHostToDevice, s));
  unsigned long long* cnt = pb.counts.ensure(2 + 128);
  MCCBN_CUDA(cudaMemsetAsync(cnt, 0, (2 + 128) * sizeof(unsigned long long), s));
  const int gx = std::min(1024, (pb.N + 255) / 256);
  init_lambda_counts_kernel<<<dim3(gx, 1), 256, 0, s>>>(pb.bits.ptr, pb.N, pb.posets.ptr + slot, cnt + 2);
  compat_kernel<<<dim3(gx, 1), 256, 0, s>>>(pb.bits.ptr, pb.N, pb.posets.ptr + slot, nullptr, cnt);
  pb.ctx->launches += 2;
  unsigned long long h[2 + 128];
  MCCBN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, s));
  MCCBN_CUDA(cudaStreamSynchronize(s));
  c.lambda.assign(pb.p, 0.0);
  lambda_from_counts(f, h + 2, pb.N, lambda_s, max_lambda, c.lambda.data());
  const double e = (double)(int)h[1] / ((double)pb.N * pb.p);
  c.eps = std::max(std::min(e, 1 - DBL_EPSILON), DBL_EPSILON);
}

// scores candidates [0, n) concurrently; results are written back into the candidates (lambda, eps, llhood)
// `while_device_busy` (optional): host work to do once the first EM iterations are enqueued, before waiting for them
static void score_candidates(mccbn_problem& pb, std::vector<Candidate*>& cands, const std::vector<uint32_t>& eval_ids,
                             Scheme scheme, unsigned L, const EmControl& ctl, bool times_given, uint32_t seed,
                             float lambda_s, int precision, const std::function<void()>& while_device_busy = nullptr) {
  // test seam (reachable only through the test-only entry point mccbn_test_asa_structural_score; mirrored by the
  // oracle's score_mode 1): a deterministic score of the poset instead of the MCEM fit, so that the move generator and
  // the annealing bookkeeping can be compared without Monte-Carlo noise.  Not part of the reference.
  if (pb.test_structural_score) {
    for (Candidate* c : cands) {
      int n_inc = 0;
      count_compatible(pb, c->poset, &n_inc);
      c->llhood = -1000.0 - (double)n_inc + 3.0 * (double)c->poset.num_edges();
    }
    if (while_device_busy) while_device_busy();
    return;
  }
  // Several GPUs (candidate-parallel annealing, pb.local_only): rank r fits the candidates q with q % world == r; the
  // fitted (llhood, eps, lambda) of all candidates are then shared by one sum-exchange (each entry is non-zero on exactly
  // one rank, so the sum is exact and every rank continues with bit-identical scores).
  const int world = pb.local_only ? pb.ctx->world : 1, rank = pb.local_only ? pb.ctx->rank : 0;
  if (world > 1 && cands.size() > 1) {
    std::vector<Candidate*> mine;
    std::vector<uint32_t> mine_ids;
    for (size_t q = 0; q < cands.size(); ++q)
      if ((int)(q % (size_t)world) == rank) {
        mine.push_back(cands[q]);
        mine_ids.push_back(eval_ids[q]);
      }
    pb.local_only = false;  // the recursive call must not split again ...
    struct Restore {
      mccbn_problem& pb;
      ~Restore() { pb.local_only = true; }
    } restore{pb};
    const int saved_world = pb.ctx->world;
    pb.ctx->world = 1;      // ... and its EM iterations stay local
    try {
      if (!mine.empty())
        score_candidates(pb, mine, mine_ids, scheme, L, ctl, times_given, seed, lambda_s, precision, while_device_busy);
      else if (while_device_busy)
        while_device_busy();
    } catch (...) {
      pb.ctx->world = saved_world;
      throw;
    }
    pb.ctx->world = saved_world;
    const int stride = pb.p + 2;
    std::vector<double> all(cands.size() * (size_t)stride, 0.0);
    for (size_t q = 0; q < cands.size(); ++q)
      if ((int)(q % (size_t)world) == rank) {
        all[q * stride] = cands[q]->llhood;
        all[q * stride + 1] = cands[q]->eps;
        for (int j = 0; j < pb.p; ++j) all[q * stride + 2 + j] = cands[q]->lambda[j];
      }
    pb.exchange_sum(all);
    for (size_t q = 0; q < cands.size(); ++q) {
      cands[q]->llhood = all[q * stride];
      cands[q]->eps = all[q * stride + 1];
      cands[q]->lambda.assign(all.begin() + q * stride + 2, all.begin() + (q + 1) * stride);
    }
    return;
  }
  size_t done = 0;
  while (done < cands.size()) {
    const size_t n = std::min(cands.size() - done, (size_t)(kMaxCand - 1));
    std::vector<int> slots;
    for (size_t q = 0; q < n; ++q) {
      Candidate& c = *cands[done + q];
      FlatPoset f;
      c.poset.flatten(f);
      pb.eval_id[q] = eval_ids[done + q];
      pb.set_model((int)q, f, c.lambda.data(), c.eps, lambda_s, ctl.max_lambda);
      slots.push_back((int)q);
    }
    std::vector<EmResult> r = run_mcem_multi(pb, slots, scheme, L, ctl, times_given, seed, false, precision, nullptr,
                                             nullptr, while_device_busy);
    for (size_t q = 0; q < n; ++q) {
      Candidate& c = *cands[done + q];
      c.lambda = r[q].lambda;
      c.eps = r[q].eps;
      c.llhood = r[q].llhood;
    }
    done += n;
  }
}

// heuristic_compatibility (src/asa.cpp:133-162)
static bool heuristic_compatibility(mccbn_problem& pb, const Poset& P, double fraction_compatible,
                                    double& fraction_compatible_new, float factor, std::mt19937& rng, bool verbose) {
  std::uniform_real_distribution<> U(0.0, 1.0);
  fraction_compatible_new = (double)count_compatible(pb, P) / pb.N;
  if (verbose) std::printf("Fraction of compatible observations: %g\n", fraction_compatible_new);
  if (fraction_compatible_new < fraction_compatible) {
    const double prob = std::exp((fraction_compatible_new - fraction_compatible) / factor);
    if (prob > U(rng)) {
      if (verbose)
        std::printf("Edge accepted despite of a lower fraction of compatible genotypes (accept. prob. = %g)\n", prob);
      return true;
    }
    return false;
  }
  return true;
}

// The move generator of propose_edge (src/asa.cpp:180-433) without the scoring: picks the move and the pair,
// edits a copy of the current poset and applies the compatibility heuristic.
static Proposal generate_proposal(mccbn_problem& pb, Candidate& M, double fraction_compatible, float factor,
                                  std::mt19937& rng, bool verbose) {
  const unsigned p = (unsigned)M.poset.p;
  Proposal pr;
  const unsigned num_edges = (unsigned)M.poset.num_edges();
  std::vector<int> idxs_pool(num_edges), cover_pool(p * p);
  std::iota(idxs_pool.begin(), idxs_pool.end(), 0);
  std::shuffle(idxs_pool.begin(), idxs_pool.end(), rng);
  std::iota(cover_pool.begin(), cover_pool.end(), 0);
  std::shuffle(cover_pool.begin(), cover_pool.end(), rng);
  std::discrete_distribution<int> dist({0.5, 0.0, 0.4, 0.1});  // :199
  pr.move = dist(rng);
  unsigned i = 0, v1 = 0, v2 = 0;
  Candidate Mn = M;
  double fnew = 0.0;
  bool found = false;
  switch (pr.move) {
    case 0:  // toggle edge (:208-253)
      for (i = 0; i < p * p; ++i) {
        Mn = M;
        v1 = (unsigned)cover_pool[i] / p;
        v2 = (unsigned)cover_pool[i] % p;
        if (v1 == v2) continue;
        if (verbose) std::printf("Testing edge: %d->%d\n", Mn.poset.event_id[v1], Mn.poset.event_id[v2]);
        const bool add = Mn.poset.add_edge((int)v1, (int)v2);
        if (add) {
          if (Mn.poset.cycle) {
            if (verbose) std::printf("Cycle!\n");
            continue;
          }
          if (!Mn.poset.reduction_flag) continue;
          if (verbose) std::printf("Adding new edge: (%d,%d)\n", Mn.poset.event_id[v1], Mn.poset.event_id[v2]);
        } else {
          Mn.poset.remove_edge((int)v1, (int)v2);
          if (verbose) std::printf("Removing edge: (%d,%d)\n", Mn.poset.event_id[v1], Mn.poset.event_id[v2]);
        }
        if (heuristic_compatibility(pb, Mn.poset, fraction_compatible, fnew, factor, rng, verbose)) {
          found = true;
          break;
        }
      }
      if (!found)
        std::printf("Warning: All add/remove moves yielded a poset with lower fraction of compatible observations\n");
      break;
    case 1:  // swap edge: probability 0 in the reference (:199), kept for completeness (:273-317)
      for (i = 0; i < num_edges; ++i) {
        Mn = M;
        int cnt = 0, a = -1, b = -1;
        for (int u = 0; u < (int)p && a < 0; ++u)
          for (uint64_t m = Mn.poset.out[u]; m; m &= m - 1, ++cnt)
            if (cnt == idxs_pool[i]) {
              a = u;
              b = ctz64(m);
              break;
            }
        if (a < 0) continue;
        v1 = (unsigned)a;
        v2 = (unsigned)b;
        Mn.poset.remove_edge(a, b);
        Mn.poset.add_edge(b, a);
        if (Mn.poset.cycle) continue;
        if (heuristic_compatibility(pb, Mn.poset, fraction_compatible, fnew, factor, rng, verbose)) {
          found = true;
          break;
        }
      }
      break;
    case 2:  // toggle cover relation (:322-372)
      for (i = 0; i < p * p; ++i) {
        v1 = (unsigned)cover_pool[i] / p;
        v2 = (unsigned)cover_pool[i] % p;
        if (v1 == v2) continue;
        if (verbose)
          std::printf("Testing edge (preserving cover relations): %d->%d\n", M.poset.event_id[v1], M.poset.event_id[v2]);
        if (M.poset.children_stale) M.poset.set_children();
        Mn = M;
        const bool add = Mn.poset.add_relation((int)v1, (int)v2);
        if (add) {
          if (Mn.poset.cycle) {
            if (verbose) std::printf("Cycle!\n");
            continue;
          }
          if (!Mn.poset.reduction_flag) continue;
          if (verbose) std::printf("Adding new cover relation: (%d,%d)\n", Mn.poset.event_id[v1], Mn.poset.event_id[v2]);
        } else {
          Mn.poset.remove_relation((int)v1, (int)v2);
          if (verbose)
            std::printf("Removing cover relation: (%d,%d)\n", Mn.poset.event_id[v1], Mn.poset.event_id[v2]);
        }
        if (!Mn.poset.sort()) continue;  // keep a valid topological order for the flattening
        if (heuristic_compatibility(pb, Mn.poset, fraction_compatible, fnew, factor, rng, verbose)) {
          found = true;
          break;
        }
      }
      if (!found) std::printf("Warning: All moves yielded a poset with lower fraction of compatible observations\n");
      break;
    case 3:  // swap node labels (:391-416)
      for (i = 0; i < p * p; ++i) {
        Mn = M;
        v1 = (unsigned)cover_pool[i] / p;
        v2 = (unsigned)cover_pool[i] % p;
        if (v1 == v2) continue;
        if (verbose) std::printf("Testing node swap: %d, %d\n", Mn.poset.event_id[v1], Mn.poset.event_id[v2]);
        Mn.poset.swap_node((int)v1, (int)v2);
        if (heuristic_compatibility(pb, Mn.poset, fraction_compatible, fnew, factor, rng, verbose)) {
          found = true;
          break;
        }
      }
      if (!found)
        std::printf("Warning: All swap-node moves yielded a poset with lower fraction of compatible observations\n");
      break;
  }
  pr.v1 = v1;
  pr.v2 = v2;
  pr.valid = found && Mn.poset.sort();
  pr.cand = Mn;
  pr.fraction_compatible_new = fnew;
  pr.rng_after = rng;
  return pr;
}

static void write_poset_file(const Poset& P, const std::string& outdir) {  // write_poset (src/asa.cpp:59-70)
  std::ofstream f(outdir + "poset.txt", std::ofstream::trunc);
  for (int u = 0; u < P.p; ++u)
    for (uint64_t m = P.out[u]; m; m &= m - 1) f << P.event_id[u] << "\t" << P.event_id[ctz64(m)] << std::endl;
}

static void stream_lambda(std::ostream& os, const std::vector<double>& l) {
  for (size_t j = 0; j < l.size(); ++j) os << (j ? " " : "") << l[j];
}

// simulated_annealing (src/asa.cpp:475-594) with speculative batched scoring.  Returns the final llhood; M holds the
// ML poset with the parameters of the final (2x longer) MCEM fit.
static double simulated_annealing(mccbn_problem& pb, Candidate& M, AsaControl& ctl_asa, unsigned L, Scheme scheme,
                                  const EmControl& ctl_em, bool times_given, uint32_t seed, float lambda_s,
                                  bool verbose, int batch) {
  NvtxRange nvtx("simulated_annealing");
  const int p = pb.p;
  const float scaling_const = -std::log(2.0) / std::log(ctl_asa.acceptance_rate);  // :484
  int num_accept = 0;
  std::vector<double> memo_add((size_t)p * p, 0.0), memo_swap((size_t)p * p, 0.0), memo_pres((size_t)p * p, 0.0);
  auto memo_of = [&](int move) -> std::vector<double>* {
    return move == 0 ? &memo_add : (move == 2 ? &memo_pres : (move == 3 ? &memo_swap : nullptr));
  };
  std::mt19937 rng_prop = make_std_rng(seed);
  std::mt19937 rng_acc = make_std_rng(seed ^ 0x9E3779B9u);
  std::uniform_real_distribution<> U(0.0, 1.0);

  // 1. score the initial poset
  {
    std::vector<Candidate*> one{&M};
    score_candidates(pb, one, {0u}, scheme, L, ctl_em, times_given, seed, lambda_s, 0);
  }
  double llhood = M.llhood;
  Candidate M_ML = M;
  double llhood_ML = llhood;
  // 2. fraction of compatible observations with the initial poset
  double fraction_compatible = (double)count_compatible(pb, M.poset) / pb.N;
  if (verbose) {
    std::printf("Initial observed Log-likelihood:%g\nFraction of compatible observations: %g\n", llhood,
                fraction_compatible);
  }
  // candidate-parallel runs: all ranks walk the same trajectory, rank 0 writes the files
  const bool writer = !(pb.local_only && pb.ctx->rank != 0);
  std::ofstream f_temp(writer ? ctl_asa.outdir + "asa.txt" : std::string("/dev/null"));
  std::ofstream f_par(writer ? ctl_asa.outdir + "params.txt" : std::string("/dev/null"));
  if (!f_temp || !f_par) throw Error(MCCBN_ERR_IO, "cannot open output files in '" + ctl_asa.outdir + "'");
  f_temp << "step\t llhood\t temperature\t acceptance rate" << std::endl;
  f_par << "step\t llhood\t epsilon\t lambdas" << std::endl;
  f_par << 0 << "\t" << llhood << "\t" << M.eps << "\t";
  stream_lambda(f_par, M.lambda);
  f_par << std::endl;

  // 3. annealing loop: iter = 1 .. max_iter - 1 (:530)
  // batch > 0: fixed speculative batch size; batch < 0: automatic with upper bound -batch -- start at 2, double after a
  // batch without an accepted move, fall back to (position of the accepted move + 1) after an acceptance
  const unsigned batch_cap = (unsigned)std::max(1, batch < 0 ? -batch : batch);
  // candidate-parallel over several GPUs: a batch smaller than the number of ranks leaves GPUs idle
  const unsigned batch_floor = pb.local_only ? std::min(batch_cap, (unsigned)pb.ctx->world) : 1u;
  unsigned batch_now = batch < 0 ? std::max(batch_floor, std::min(2u, batch_cap)) : batch_cap;
  unsigned iter = 1;
  // Look-ahead: proposals are generated (cheap, host side) further ahead than they are scored, under the same "all
  // rejected" assumption as the scoring batch, so that the NVRTC-specialised kernels of the next candidate posets can be
  // compiled on background host threads while the device fits the current ones (a compilation takes about as long as
  // ten candidate fits of BASELINE config 4).  The proposal sequence, and with it the result, does not depend on the
  // depth: the proposals come from one generator in order, and an acceptance discards everything generated after it.
  std::deque<Proposal> ahead;
  const bool prefetch = pb.jit_mode_effective() != 0 && (double)pb.N * (double)L >= 5e6 && scheme != kPool;
  unsigned depth = 12;
  if (const char* e = std::getenv("MCCBN_ASA_LOOKAHEAD")) depth = (unsigned)std::max(1, std::min(64, std::atoi(e)));
  // kernels are compiled ahead only while proposals are being rejected: at high temperature nearly every move is
  // accepted, the look-ahead is discarded each step and the compilations would be wasted host work
  unsigned steps_since_accept = 0;
  unsigned gate = 3;
  if (const char* e = std::getenv("MCCBN_ASA_PREFETCH_GATE")) gate = (unsigned)std::max(0, std::atoi(e));
  static const bool trace = std::getenv("MCCBN_ASA_TRACE") != nullptr;  // dev: where the host time of a run goes
  double t_gen = 0.0, t_score = 0.0, t_init = 0.0;
  unsigned n_gen = 0;
  auto clk = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
  };
  auto top_up = [&](unsigned want) {
    while (ahead.size() < want) {
      const auto tg = clk();
      ahead.push_back(generate_proposal(pb, M, fraction_compatible, ctl_asa.compatible_fraction_factor, rng_prop, verbose));
      t_gen += ms_since(tg);
      ++n_gen;
      if (prefetch && steps_since_accept >= gate && ahead.back().valid)
        pb.prefetch_kernel(ahead.back().cand.poset, scheme, times_given);
    }
  };
  while (iter < ctl_asa.max_iter) {
    // 3.0 speculative batch: proposals for steps iter, iter+1, ... assuming every earlier one is rejected
    const unsigned B = std::min<unsigned>(batch_now, ctl_asa.max_iter - iter);
    top_up(B);  // usually already there: generated while the device was fitting the previous batch (below)
    std::deque<Proposal>& props = ahead;  // the first B entries form the scoring batch
    std::vector<Candidate*> to_score;
    std::vector<uint32_t> ids;
    for (unsigned b = 0; b < B; ++b) {
      Proposal& pr = props[b];
      if (!pr.valid) continue;
      std::vector<double>* memo = memo_of(pr.move);
      if (memo && (*memo)[(size_t)pr.v2 * p + pr.v1] != 0.0) {  // :254, :374, :418 -- 0.0 means "not computed"
        pr.cached = true;
        pr.cand.llhood = (*memo)[(size_t)pr.v2 * p + pr.v1];
        continue;
      }
      for (unsigned q = 0; q < b; ++q) {
        const Proposal& o = props[q];
        const bool same = o.valid && !o.cached && o.move == pr.move && memo &&
                          ((o.v1 == pr.v1 && o.v2 == pr.v2) || (pr.move == 3 && o.v1 == pr.v2 && o.v2 == pr.v1));
        if (same) {
          pr.same_as = o.same_as >= 0 ? o.same_as : (int)q;
          break;
        }
      }
      if (pr.same_as >= 0) continue;
      const auto ti = clk();
      initialize_candidate(pb, pr.cand, lambda_s, ctl_em.max_lambda);  // :256-258
      t_init += ms_since(ti);
      pr.score_slot = (int)to_score.size();
      to_score.push_back(&pr.cand);
      ids.push_back(iter + b);
    }
    // 3.a score all distinct, un-memoised proposals of the batch concurrently
    // ... and, once their first EM iterations are enqueued, generate the look-ahead proposals and start / collect the
    // compilation of their kernels while the device is busy
    const unsigned want = std::min<unsigned>(B + depth, ctl_asa.max_iter - iter);
    // (a slice per call: it runs once per EM round of the fit, between enqueueing ~20 iterations and waiting for them,
    // and must not take longer than the device needs for those)
    const std::function<void()> look_ahead = [&] {
      if (!prefetch) return;
      const auto t0 = std::chrono::steady_clock::now();
      while (ahead.size() < want) {
        top_up((unsigned)ahead.size() + 1);
        if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > 1.5) break;
      }
    };
    if (!to_score.empty()) {
      NvtxRange nvtx_score("ASA: score candidate batch");
      const auto ts = clk();
      score_candidates(pb, to_score, ids, scheme, L, ctl_em, times_given, seed, lambda_s, 0, look_ahead);
      t_score += ms_since(ts);
    } else {
      look_ahead();
    }
    // 3.b consume them in order with the reference's accept / reject rule (:436-468)
    bool accepted_any = false;
    unsigned n_consumed = B;
    for (unsigned b = 0; b < B && !accepted_any; ++b, ++iter) {
      Proposal& pr = props[b];
      if (verbose) std::printf("Step %u - log-likelihood: %g\n", iter, llhood);
      double llhood_new = llhood;
      bool accept = false;
      if (pr.valid) {
        if (pr.same_as >= 0) pr.cand.llhood = props[pr.same_as].cand.llhood;
        llhood_new = pr.cand.llhood;
        std::vector<double>* memo = memo_of(pr.move);
        if (memo && !pr.cached) {
          (*memo)[(size_t)pr.v2 * p + pr.v1] = llhood_new;
          if (pr.move == 3) (*memo)[(size_t)pr.v1 * p + pr.v2] = llhood_new;  // :428
        }
        if (llhood_new > llhood) {
          accept = true;
        } else if (ctl_asa.T != 0.0f) {
          const double prob = std::exp(-(llhood - llhood_new) / ctl_asa.T);
          accept = prob > U(rng_acc);
        }
      }
      if (accept) {
        num_accept += 1;
        // a memoised (or, within a batch, duplicated) score carries no fitted parameters: the reference copies
        // M_new as it is, i.e. with the parameters of the current model (:438)
        M = pr.cand;
        fraction_compatible = pr.fraction_compatible_new;
        llhood = llhood_new;
        if (verbose) std::printf("Log-likelihood new poset:%g\n", llhood_new);
        std::fill(memo_add.begin(), memo_add.end(), 0.0);
        std::fill(memo_swap.begin(), memo_swap.end(), 0.0);
        std::fill(memo_pres.begin(), memo_pres.end(), 0.0);
        rng_prop = pr.rng_after;  // discard the speculation beyond this step
        // the accepted poset's own kernel serves every label-swap candidate of the following steps (move 3 relabels
        // the events: same structure by topological position, same kernel) and a return to it after a detour
        if (prefetch) pb.prefetch_kernel(M.poset, scheme, times_given);
        accepted_any = true;
        n_consumed = b + 1;
        steps_since_accept = 0;
        if (batch < 0) batch_now = std::max(batch_floor, std::min(batch_cap, b + 1u));
      }
      if (!accept) ++steps_since_accept;
      if (llhood > llhood_ML) {  // :541-545
        llhood_ML = llhood;
        M_ML = M;
        if (writer) write_poset_file(M.poset, ctl_asa.outdir);
      }
      f_par << iter << "\t" << llhood << "\t" << M.eps << "\t";
      stream_lambda(f_par, M.lambda);
      f_par << std::endl;
      // 3.c temperature (:555-579)
      if (!ctl_asa.adaptive) {
        const float rate = (float)num_accept / (iter + 1);
        ctl_asa.T *= 1 - 1.0 / p;
        f_temp << iter << "\t" << llhood << "\t" << ctl_asa.T << "\t" << rate << std::endl;
      } else if (ctl_asa.step_size && !(iter % ctl_asa.step_size)) {
        const float rate = (float)num_accept / ctl_asa.step_size;
        ctl_asa.T *= std::exp((0.5 - std::pow(rate, scaling_const)) * ctl_asa.adap_rate);
        num_accept = 0;
        f_temp << iter << "\t" << llhood << "\t" << ctl_asa.T << "\t" << rate << std::endl;
      }
    }
    if (batch < 0 && !accepted_any) batch_now = std::min(batch_cap, batch_now * 2u);
    // proposals generated after an accepted one belong to the old poset: dropped (the generator was rewound above);
    // without an acceptance only the consumed ones leave the look-ahead queue
    if (accepted_any) ahead.clear();
    else ahead.erase(ahead.begin(), ahead.begin() + n_consumed);
  }
  if (trace)
    std::fprintf(stderr, "mccbn asa: %u proposals generated in %.1f ms (%.2f ms each), candidate initialisation %.1f ms, "
                         "scoring (incl. look-ahead inside) %.1f ms\n", n_gen, t_gen, n_gen ? t_gen / n_gen : 0.0, t_init, t_score);
  // 4. final fit of the ML poset with doubled EM budget; neighborhood_dist falls back to its default 1 (:583-585)
  EmControl last = ctl_em;
  last.max_iter = ctl_em.max_iter * 2;
  last.update_step_size = ctl_em.update_step_size * 2;
  last.neighborhood_dist = 1;
  M = M_ML;
  {
    std::vector<Candidate*> one{&M};
    score_candidates(pb, one, {ctl_asa.max_iter}, scheme, L, last, times_given, seed, lambda_s, 0);
  }
  return M.llhood;
}

}  // namespace mccbn
