/* mccbn_b200.h -- C ABI of libmccbn_b200.so: the B200-native (sm_100a CUDA) replacement for the
 * Monte-Carlo-EM E-step hot path of cbg-ethz/MC-CBN (H-CBN2 model).
 *
 * Drop-in boundary.  The reference is an R package whose R functions reach C++ through
 * `.Call('_name', PACKAGE='mccbn', ...)` on `RcppExport SEXP _name(SEXP...)` symbols (useDynLib(mccbn) without a
 * registration table, NAMESPACE:37, so R resolves by symbol name).  Every entry point below takes the SAME
 * arguments in the SAME order and memory layout as the reference export it replaces (R's column-major int32
 * matrices, double vectors, `float` lambda_s / max_lambda), followed by the dimensions R carries in the SEXP
 * headers (N, p), the outputs the reference returns as an R list, and an error buffer.  INTEGRATION.md shows the
 * ~10-line RcppExport stub per symbol that forwards to these functions.
 *
 * Conventions
 *   - matrices are column-major: element (r, c) of an R x C matrix is a[c*R + r]
 *   - poset[i + j*p] == 1  <=>  cover relation i -> j (i is a parent of j)      (src/model_impl.cpp:343-352)
 *   - obs[i + j*N] != 0    <=>  event j observed in genotype i                   (as<MatrixXb>, src/mcem_hcbn.cpp:817)
 *   - all pointers are HOST pointers unless the name ends in `_dev`
 *   - return value: MCCBN_OK or an MCCBN_ERR_* code; on error `err` (if non-NULL) receives the message the
 *     reference would have raised through Rf_error (src/mcem_hcbn.cpp:30-40)
 *   - `thrds` is accepted for signature compatibility and ignored (the device replaces the OpenMP team)
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns MCCBN_ERR_CUDA
 *
 * Supported sizes: p <= 32 events for the Philox fast path (all BASELINE configs; one 32-bit genotype word),
 * p <= 64 for the indexing entry points.
 */
#ifndef MCCBN_B200_H
#define MCCBN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCCBN_OK 0
#define MCCBN_ERR_NOT_ACYCLIC 1 /* "Poset does not correspond to an acyclic graph" (src/not_acyclic_exception.hpp:18) */
#define MCCBN_ERR_ZERO_WEIGHT 2 /* "ERROR: all samples have weight 0. Consider increasing L" (src/mcem_hcbn.cpp:697) */
#define MCCBN_ERR_CUDA 3        /* CUDA / NCCL failure, or no device */
#define MCCBN_ERR_ARG 4         /* invalid argument (unknown sampling scheme, p too large, ...) */
#define MCCBN_ERR_IO 5          /* ASA output files */

#define MCCBN_MAX_P 64
#define MCCBN_FAST_MAX_P 32 /* all schemes; the forward scheme accepts p <= MCCBN_MAX_P (64-bit genotype words) */
#define MCCBN_NCCL_ID_BYTES 128
#define MCCBN_IPC_HANDLE_BYTES 64

typedef struct mccbn_ctx mccbn_ctx; /* owns the device buffers, stream, NCCL communicator */

/* ---- context -------------------------------------------------------------------------------------------- */
const char* mccbn_version(void);
/* number of visible CUDA devices (0 if none / no driver) */
int mccbn_device_count(void);
/* Create a context on CUDA device `device`.  All entry points accept ctx == NULL and then use a lazily created
 * process-wide context on device 0 (or $LOCAL_RANK if set). */
int mccbn_ctx_create(mccbn_ctx** out, int device, char* err, size_t errlen);
void mccbn_ctx_destroy(mccbn_ctx* ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t); NULL = the context's own non-blocking stream.  The work of one
 * context is ordered by one stream at a time: switching waits for the old stream to drain.  Contexts are independent of
 * each other, also on the same device. */
int mccbn_ctx_set_stream(mccbn_ctx* ctx, void* cuda_stream);
/* Multi-GPU, one process per GPU: observations are sharded over `world` ranks; each EM iteration one
 * sum-allreduce of p+3 doubles (NCCL over NVLink) combines the expected sufficient statistics
 * (replaces the OpenMP reduction of src/mcem_hcbn.cpp:665).  Rank 0 calls mccbn_nccl_unique_id and ships the
 * 128 bytes to the other ranks by any means (torch.distributed broadcast in the Python binding). */
int mccbn_nccl_unique_id(unsigned char id[MCCBN_NCCL_ID_BYTES], char* err, size_t errlen);
int mccbn_ctx_init_comm(mccbn_ctx* ctx, int rank, int world, const unsigned char id[MCCBN_NCCL_ID_BYTES], char* err,
                        size_t errlen);
/* Preferred on one NVLink / NVSwitch domain (<= 16 GPUs of one box): the statistic exchange fused into the reduction
 * and M-step kernels over peer memory.  Every rank owns a mailbox in its HBM; the reduction kernel stores the rank's
 * p+5 doubles into every rank's mailbox with NVLink peer stores and publishes a sequence number, the M-step kernel
 * waits for all ranks and sums in rank order (bit-identical parameters on all ranks, no collective call, no host
 * involvement).  Bootstrap: every rank calls mccbn_ctx_peer_handle, the 64-byte CUDA IPC handles are all-gathered
 * by any means (torch.distributed in the Python binding), then every rank calls mccbn_ctx_init_peers with the
 * world x 64 bytes in rank order.  When both this and the NCCL communicator are initialised, this one is used. */
int mccbn_ctx_peer_handle(mccbn_ctx* ctx, unsigned char handle[MCCBN_IPC_HANDLE_BYTES], char* err, size_t errlen);
int mccbn_ctx_init_peers(mccbn_ctx* ctx, int rank, int world, const unsigned char* handles, char* err, size_t errlen);
/* Counters for the bench contract: kernels launched by this library since the last reset. */
/* In-process multi-GPU ("device team").  The reference's callers are a single process (an R session): with
 * mccbn_set_gpus(G) -- or MCCBN_GPUS=G / MCCBN_GPUS=all in the environment -- the entries below that are called with
 * ctx == NULL and without caller-side sharding (mccbn_em_options.obs_offset / N_global) split the observations over
 * the first G devices themselves: one host thread and one context per device, the statistic vector exchanged through
 * mailboxes inside the reduction / M-step kernels exactly as between the ranks of a multi-process job.  This is the
 * role the reference's `thrds` argument plays for its OpenMP loop (src/mcem_hcbn.cpp:661-665).  Philox counters carry
 * global observation indices: the result does not depend on G.
 *   sharded by observation : mccbn_MCEM_hcbn, mccbn_obs_log_likelihood, mccbn_importance_weight
 *   candidate-parallel     : mccbn_adaptive_simulated_annealing (every device holds all observations)
 * n_gpus <= 1 turns it off; more devices than present are clamped.  mccbn_get_gpus() = devices a call would use. */
int mccbn_set_gpus(int n_gpus, char* err, size_t errlen);
int mccbn_get_gpus(void);

uint64_t mccbn_kernel_launches(mccbn_ctx* ctx, int reset);
/* of those, launches of NVRTC poset-specialised kernels (mccbn_fwd_spec / mccbn_cond_spec, csrc/jit.cuh): lets a caller
 * (tests, bench.py) state which kernel actually ran */
uint64_t mccbn_jit_launches(mccbn_ctx* ctx, int reset);

/* ---- extended options shared by the EM entry points ------------------------------------------------------ */
typedef struct mccbn_em_options {
  /* sharding (multi-GPU): `obs`, `times`, `weights` hold the LOCAL shard of N rows; obs_offset is the global
   * index of local row 0, N_global the global number of observations (0 => N).  Philox counters use global row
   * indices, so results do not depend on the number of ranks. */
  int64_t obs_offset;
  int64_t N_global;
  /* optional per-iteration trace: max_iter x (p+2) doubles [obs_llhood, eps, lambda_0..lambda_{p-1}] row-major
   * per iteration (the `verbose` line of src/mcem_hcbn.cpp:723-728); n_iter receives the number of rows */
  double* trace;
  int* n_iter;
  /* 0 = fp32 sampling / fp64 reduction (default); 1 = fp64 sampling (validation mode, slower) */
  int precision;
  /* pool scheme: 0 = expectations sum_k q_k g_k / sum_k q_k over the whole pool (default: the quantity the reference's
   * resampling estimates, without its extra Monte-Carlo noise), 1 = draw L indices with replacement from Discrete(q) and
   * average over the drawn rows like the reference (src/mcem_hcbn.cpp:477-484); obs_llhood is the same in both modes */
  int pool_resample;
  /* stream id mixed into the Philox key so that independent evaluations (ASA candidates) decorrelate */
  uint32_t stream_id;
  /* leading dimension of `obs`: when the N local rows are rows [r, r + N) of a larger column-major matrix with obs_ld
   * rows, pass obs + r and obs_ld (the device upload is then a strided copy; no host-side gather).  0 => N. */
  uint32_t obs_ld;
  uint32_t reserved[4];
} mccbn_em_options;

/* ---- hot-path entry points (names/arguments follow the reference exports) -------------------------------- */

/* replaces `_MCEM_hcbn` (src/mcem_hcbn.cpp:804-864); returns list(lambda, eps, llhood) through the out pointers */
int mccbn_MCEM_hcbn(mccbn_ctx* ctx, const double* ilambda, const int* poset, const int* obs, const double* times,
                    float lambda_s, double eps, const double* weights, unsigned L, const char* sampling,
                    unsigned max_iter, unsigned update_step_size, double tol, float max_lambda,
                    unsigned neighborhood_dist, int sampling_times_available, int thrds, int verbose, int seed,
                    /* dims */ int N, int p,
                    /* out */ double* lambda_out, double* eps_out, double* llhood_out,
                    /* optional */ const mccbn_em_options* opt, char* err, size_t errlen);

/* replaces `_obs_log_likelihood` (src/mcem_hcbn.cpp:762-797) */
int mccbn_obs_log_likelihood(mccbn_ctx* ctx, const int* obs, const int* poset, const double* lambda, double eps,
                             const double* weights, const double* times, unsigned L, const char* sampling,
                             unsigned neighborhood_dist, float lambda_s, int sampling_times_available, int thrds,
                             int seed, int N, int p, double* llhood_out, const mccbn_em_options* opt, char* err,
                             size_t errlen);

/* replaces `_importance_weight` (src/mcem_hcbn.cpp:914-1010): per-observation sum w, sum w^2, L_eff (#w>0; only
 * meaningful for backward/bernoulli, may be NULL), E[d], E[Z] (N x p column-major) */
int mccbn_importance_weight(mccbn_ctx* ctx, const int* obs, unsigned L, const int* poset, const double* lambda,
                            double eps, const double* times, const char* sampling, unsigned neighborhood_dist,
                            float lambda_s, int sampling_times_available, int thrds, int seed, int N, int p,
                            double* w_sum, double* w_sum_sq, int* L_eff, double* dist, double* Tdiff,
                            const mccbn_em_options* opt, char* err, size_t errlen);

/* replaces `_importance_weight_genotype` (src/mcem_hcbn.cpp:866-912): the raw per-sample quantities for ONE
 * genotype: w[L], dist[L], Tdiff[L x p col-major].  *L_out = rows produced (backward rounds L down to
 * reps * nrows_sum, :356-363).  forward/backward/bernoulli only. */
int mccbn_importance_weight_genotype(mccbn_ctx* ctx, const int* genotype, unsigned L, const int* poset,
                                     const double* lambda, double eps, double time, const char* sampling,
                                     unsigned neighborhood_dist, float lambda_s, int sampling_times_available,
                                     int seed, int p, int* L_out, double* w, int* dist, double* Tdiff,
                                     const mccbn_em_options* opt, char* err, size_t errlen);

/* replaces `_adaptive_simulated_annealing` (src/asa.cpp:596-665); poset_out = p x p adjacency over event ids */
int mccbn_adaptive_simulated_annealing(mccbn_ctx* ctx, const int* poset, const int* obs, const double* times,
                                       float lambda_s, const double* weights, unsigned L, const char* sampling,
                                       unsigned max_iter_EM, unsigned update_step_size, double tol, float max_lambda,
                                       float T0, float adap_rate, float acceptance_rate, int step_size,
                                       unsigned max_iter_ASA, unsigned neighborhood_dist, int adaptive,
                                       const char* outdir, int sampling_times_available, int thrds, int verbose,
                                       int seed, int N, int p, double* lambda_out, double* eps_out, int* poset_out,
                                       double* llhood_out, char* err, size_t errlen);

/* TEST-ONLY seam, not a reference export: the annealer with every MCEM score replaced by the deterministic structural
 * score -1000 - num_incompatible_events + 3 |cover relations|, so that tests can compare the host logic of src/asa.cpp
 * (move generator, memo tables, Metropolis rule, temperature schedule, files) with the oracle without Monte-Carlo noise.
 * Same arguments as mccbn_adaptive_simulated_annealing; always runs on one device. */
int mccbn_test_asa_structural_score(mccbn_ctx* ctx, const int* poset, const int* obs, const double* times,
                                    float lambda_s, const double* weights, unsigned L, const char* sampling,
                                    unsigned max_iter_EM, unsigned update_step_size, double tol, float max_lambda,
                                    float T0, float adap_rate, float acceptance_rate, int step_size,
                                    unsigned max_iter_ASA, unsigned neighborhood_dist, int adaptive,
                                    const char* outdir, int sampling_times_available, int thrds, int verbose,
                                    int seed, int N, int p, double* lambda_out, double* eps_out, int* poset_out,
                                    double* llhood_out, char* err, size_t errlen);

/* replaces `MCEM` (legacy OT-CBN, src/mcem.cpp:307-335, called by estimate_mutation_rates): noise-free model,
 * X == Y, conditional (truncated-exponential) proposal; returns list(lambda, llhood) */
int mccbn_MCEM(mccbn_ctx* ctx, const double* ilambda, const int* poset, const int* O, const double* times,
               int max_iter, float alpha, const int* topo_path, const double* weights, int nrOfSamples, int verbose,
               float max_lambda, float lambda_s, int sampling_time_available, int seed, int N, int p,
               double* lambda_out, double* llhood_out, char* err, size_t errlen);

/* ---- callers of the OT-CBN conditional proposal -----------------------------------------------------
 * replaces `drawHiddenVarsSamples` (src/mcem.cpp:214-248; .Call'ed by R/mcem_genotype_probabilities.r:17,65): N draws of
 * the conditional proposal for ONE genotype O_vec at sampling time `time` (drawn ~ Exp(lambda_s) per sample when
 * !sampling_time_available).  T is N x p column-major (waiting times by event id), densities[N] the log proposal
 * density of each draw; time_out[N] (may be NULL; not in the reference's list) the sampling time each draw used.
 * The reference consumes R's global generator; here the draws are Philox streams of `seed`. */
int mccbn_drawHiddenVarsSamples(mccbn_ctx* ctx, int N, const int* O_vec, double time, const double* lambda,
                                const int* poset, const int* topo_path, float lambda_s, int sampling_time_available,
                                int seed, int p, double* T, double* densities, double* time_out, char* err,
                                size_t errlen);
/* replaces `expectedTimeDifference` (src/mcem.cpp:251-276): E[Z_j | genotype] by self-normalised importance sampling */
int mccbn_expectedTimeDifference(mccbn_ctx* ctx, int nrOfSamples, const int* O_vec, double time, const double* lambda,
                                 const int* poset, const int* topo_path, float lambda_s, int sampling_time_available,
                                 int seed, int p, double* time_difference, char* err, size_t errlen);
/* batched form of `loglike_importance_sampling` (R/mcem_genotype_probabilities.r:10-23), which loops over the
 * observations with one .Call("drawHiddenVarsSamples") each: every (observation, sample) pair in ONE launch.
 * probs[N] (may be NULL) = weights[i] * (logsumexp_l(log P(Z_il) - log q_il) - log nrOfSamples); approx_loglike = sum.
 * All observations must be compatible with the poset (the R caller filters them, :86-91). */
int mccbn_loglike_importance_sampling(mccbn_ctx* ctx, const int* poset, const double* lambda, const int* obs_events,
                                      const double* sampling_times, int nrOfSamples, const double* weights,
                                      float lambda_s, int sampling_times_available, int seed, int N, int p,
                                      double* approx_loglike, double* probs, char* err, size_t errlen);

/* ---- indexing seams (bit-exact tier) --------------------------------------------------------------------- */

/* ---- forward simulators of the generative model (test / benchmark-data seams of the reference) -------------------
 * `_sample_genotypes` (src/mcem_hcbn.cpp:1012-1049): N draws of (X, Z, T_s); samples / Tdiff are N x p column-major,
 * T_sampling is read when sampling_times_available and written otherwise.
 * `_sample_times` (:1052-1083): mutation_times = accumulated event times T, Tdiff = waiting times Z.
 * `_generate_genotypes` (:1086-1120): X = [T_events_sum <= T_s] for given event times. */
int mccbn_sample_genotypes(mccbn_ctx* ctx, unsigned N, const int* poset, const double* lambda, double* T_sampling,
                           float lambda_s, int sampling_times_available, int seed, int p, int* samples, double* Tdiff,
                           char* err, size_t errlen);
int mccbn_sample_times(mccbn_ctx* ctx, unsigned N, const int* poset, const double* lambda, int seed, int p,
                       double* mutation_times, double* Tdiff, char* err, size_t errlen);
int mccbn_generate_genotypes(mccbn_ctx* ctx, const double* T_events_sum, const int* poset, double* T_sampling,
                             float lambda_s, int sampling_times_available, int seed, int N, int p, int* samples,
                             char* err, size_t errlen);

/* ---- the reference's remaining small export seams ------------------------------------------------------------
 * `_scale_path_to_mutation` src/add_remove.cpp:218-242, `_cbn_density_log` :323-335, `_complete_log_likelihood`
 * src/mcem_hcbn.cpp:738-760 (host arithmetic), `_draw_sample` src/add_remove.cpp:244-284 (reference generator: bit-exact
 * for a seed), `_coin_tossing` src/backward_sampling.cpp:105-123 (device, Philox coins), `_generate_mutation_times`
 * src/add_remove.cpp:286-321 (reference generator and draw order, fp64 arithmetic on the device). */
int mccbn_scale_path_to_mutation(const int* poset, int p, const double* lambda, double* out, char* err, size_t errlen);
int mccbn_cbn_density_log(const double* time, int N, int p, const double* lambda, double* out, char* err,
                          size_t errlen);
int mccbn_complete_log_likelihood(const double* lambda, int p, double eps, const double* Tdiff, int N,
                                  const double* dist, const double* weights, double* out, char* err, size_t errlen);
int mccbn_draw_sample(const int* genotype, const int* poset, int p, unsigned move, const double* q_prob, int seed,
                      int* sample, double* q_choice, char* err, size_t errlen);
int mccbn_coin_tossing(mccbn_ctx* ctx, double eps, unsigned L, const int* genotype, int p, int seed, int* out, char* err,
                       size_t errlen);
int mccbn_generate_mutation_times(mccbn_ctx* ctx, const int* obs, const int* poset, const double* lambda, double* time,
                                  float lambda_s, int sampling_times_available, int seed, int N, int p, double* Tdiff,
                                  double* proposal_density, char* err, size_t errlen);

/* Stage 1, poset flattening (adjacency_mat2list + has_cycles + topological_sort + get_direct_predecessors,
 * src/model_impl.cpp:343-352,40-87): topo[p] = event ids in a valid topological order, parent_mask[p] = bitmask of
 * direct parents by event id.  Returns MCCBN_ERR_NOT_ACYCLIC for a cyclic poset.  Host-only (no device needed). */
int mccbn_flatten_poset(const int* poset, int p, int* topo, uint64_t* parent_mask, char* err, size_t errlen);

/* replaces `_compatible_observations` (src/compatible_observations.cpp:91-116): out[i] = is_compatible(obs[i,])
 * num_compatible / num_incompatible_events (may be NULL) follow :68-76 and :51-66.  Device reduction. */
int mccbn_compatible_observations(mccbn_ctx* ctx, const int* obs, const int* poset, int N, int p, int* out,
                                  int* num_compatible, int* num_incompatible_events, char* err, size_t errlen);

/* replaces `_neighbors` (src/backward_sampling.cpp:79-103): the Hamming ball of radius k around `genotype` in the
 * reference's enumeration order; out is nrows_sum x p column-major, ring[nrows_sum] (may be NULL) the distance.
 * Host-only. */
int mccbn_neighbors(int p, int k, const int* genotype, int* out, int* ring, int* nrows_sum, char* err, size_t errlen);

/* initialize_lambda + the ASA start value of eps (src/asa.cpp:72-129, :640-642); device counts, float division */
int mccbn_initialize_lambda(mccbn_ctx* ctx, const int* obs, const int* poset, int N, int p, float lambda_s,
                            float max_lambda, double* lambda_out, double* eps_out, char* err, size_t errlen);

/* ---- supplied-randomness seams (1e-10 fp64 tier; no RNG on either side) ---------------------------------- */

/* forward scheme from host-supplied waiting times: Z is L x p, T_sampling is L.  Per sample: propagated event
 * times T_sum (L x p, may be NULL), hidden genotype X (L x p int, may be NULL), Hamming distance to `genotype`,
 * weight eps^d (1-eps)^(p-d)   (sample_genotypes :160-175 + hamming_dist_mat :326-330 + weights :386-388) */
int mccbn_forward_from_times(mccbn_ctx* ctx, const int* genotype, int L, const int* poset, int p, double eps,
                             const double* Z, const double* T_sampling, double* T_sum, int* X, int* dist, double* w,
                             char* err, size_t errlen);

/* One full E-step + M-step from supplied forward draws shared by all observations (if sampling_times_available
 * the sampling time of observation i is times[i] instead of T_sampling[l]).  Outputs as `_importance_weight`
 * plus totals[3+p] = [obs_llhood, expected_dist, N_eff, Tdiff_colsum...] and the M-step result
 * (src/mcem_hcbn.cpp:683-709). */
int mccbn_estep_forward_from_times(mccbn_ctx* ctx, const int* obs, int N, const int* poset, int p,
                                   const double* lambda, double eps, const double* weights, const double* times,
                                   int sampling_times_available, int L, const double* Z, const double* T_sampling,
                                   float max_lambda, double* w_sum, double* w_sum_sq, double* dist, double* Tdiff,
                                   double* totals, double* eps_new, double* lambda_new, char* err, size_t errlen);

/* pool scheme from a supplied pool (sample_times output T_pool / Tdiff_pool, K x p column-major; T_sampling[K] unless
 * sampling_times_available): per observation the pool genotypes X_k = [T_pool(k,) <= T_s] and their Hamming distances
 * (generate_genotypes :215-236, hamming_dist_mat :326-330; X_pool = K x p for observation 0, d_pool = K x N, both
 * bit-exact), q_sum[i] = sum_k eps^d (1-eps)^(p-d) (so that obs_llhood_i = log(q_sum[i] / K), :491, :690), and the
 * conditional expectations sum_k q_k g_k / sum_k q_k of d and Z that the reference's L-index resampling estimates
 * (:477-484).  All outputs optional. */
int mccbn_pool_from_times(mccbn_ctx* ctx, const int* obs, int N, const int* poset, int p, double eps,
                          const double* weights, const double* times, int sampling_times_available, int K,
                          const double* T_pool, const double* Tdiff_pool, const double* T_sampling, int* X_pool,
                          int* d_pool, double* q_sum, double* dist, double* Tdiff, double* llhood, char* err,
                          size_t errlen);

/* conditional proposal (generate_mutation_times + cbn_density_log, src/add_remove.cpp:156-216) from supplied
 * uniforms u (L x p, indexed by event id) and sampling times: Tdiff (L x p), log q (L), log P(Z) (L),
 * incompatible flag (L; any Z < 0, src/mcem_hcbn.cpp:524) */
int mccbn_conditional_from_uniforms(mccbn_ctx* ctx, const int* X, int L, const int* poset, int p,
                                    const double* lambda, const double* u, const double* T_sampling, double* Tdiff,
                                    double* log_q, double* log_pz, int* incompatible, char* err, size_t errlen);

/* ---- device-resident problem handle (what bench.py times with inputs already in HBM) --------------------- */
typedef struct mccbn_problem mccbn_problem;

/* upload + bit-pack the observation shard once */
int mccbn_problem_create(mccbn_ctx* ctx, const int* obs, const double* times, const double* weights, int N, int p,
                         int64_t obs_offset, int64_t N_global, mccbn_problem** out, char* err, size_t errlen);
void mccbn_problem_destroy(mccbn_problem* pb);
/* set poset + parameters (resets the EM iteration counter) */
int mccbn_problem_set_model(mccbn_problem* pb, const int* poset, const double* lambda, double eps, float lambda_s,
                            float max_lambda, char* err, size_t errlen);
/* enqueue `n_iter` EM iterations (E-step, reduction, allreduce if sharded, M-step) on the context stream without
 * synchronising; parameters stay on the device */
int mccbn_problem_em_iterations(mccbn_problem* pb, unsigned L, const char* sampling, unsigned neighborhood_dist,
                                int sampling_times_available, int seed, int n_iter, int update_params, char* err,
                                size_t errlen);
/* synchronise and read back the current parameters and the last observed log-likelihood */
int mccbn_problem_read(mccbn_problem* pb, double* lambda, double* eps, double* llhood, char* err, size_t errlen);
/* the CUDA stream the problem's work is enqueued on (for event timing by the caller) */
void* mccbn_problem_stream(mccbn_problem* pb);

/* Poset-specialised kernels (NVRTC, csrc/jit.cuh).  mode: -1 automatic (default; env MCCBN_JIT=0/1 overrides), 0 never,
 * 1 always (compile on first use).  Automatic: a kernel found in the context's memory cache or in the on-disk cache
 * ($MCCBN_CACHE_DIR, default ~/.cache/mccbn_b200) is always used; otherwise it is compiled at once when the announced
 * sampling work (samples per E-step x expected E-steps) is >= 5e9 samples, compiled on a background host thread when one
 * E-step has >= 5e7 samples (the generic kernel runs until it is ready), and not at all for smaller problems.  The
 * specialised and the generic kernel draw the same random stream and write identical records, so results do not depend
 * on the mode. */
int mccbn_set_jit(mccbn_ctx* ctx, int mode);
/* host-only: generate and compile (no device needed) the specialised kernel of `poset`; log = NVRTC / ptxas -v
 * output, source = generated CUDA text (either may be NULL) */
int mccbn_jit_compile_forward(const int* poset, int p, int times_given, size_t* cubin_bytes, double* compile_ms,
                              char* log, size_t loglen, char* source, size_t sourcelen, char* err, size_t errlen);
/* the same for the conditional-proposal kernels; mode: 0 backward, 1 bernoulli, 2 OT-CBN, 3 add-remove */
int mccbn_jit_compile_conditional(const int* poset, int p, int times_given, int mode, size_t* cubin_bytes,
                                  double* compile_ms, char* log, size_t loglen, char* source, size_t sourcelen,
                                  char* err, size_t errlen);

/* Host-only test seam for the poset edit operations the annealer uses (Model::add_edge, remove_edge, add_relation,
 * remove_relation, transitive_reduction_dag, swap_node; src/model_impl.cpp:145-316; the reference exposes the
 * debug exports `_remove_test` / `_transitive_reduction_dag`, src/asa.cpp:667-735).  op: 0 add_edge, 1 remove_edge,
 * 2 add_relation, 3 remove_relation, 4 transitive_reduction_dag, 5 swap_node.  *flags: bit 0 = edge inserted,
 * bit 1 = cycle, bit 2 = reduction_flag.  poset_out = adjacency over event ids (adjacency_list2mat). */
int mccbn_poset_edit(const int* poset, int p, int op, int u, int v, int* poset_out, int* flags, char* err,
                     size_t errlen);

/* ---- environment ------------------------------------------------------------------------------------------------
 * None of these changes WHAT is computed (random streams, records, estimators); they select code paths for A/B runs,
 * size caches or print diagnostics.
 *   MCCBN_GPUS=G|all         in-process device team (see mccbn_set_gpus)
 *   MCCBN_JIT=0|1            never / always use poset-specialised kernels (see mccbn_set_jit)
 *   MCCBN_CACHE_DIR          on-disk cubin cache (default ~/.cache/mccbn_b200); MCCBN_JIT_MAX_LOADED: LRU bound on
 *                            loaded modules (256)
 *   MCCBN_DEVICE_CACHE_MB    per-device cache of the one-call entries' buffers (4096; 0 disables it)
 *   MCCBN_ASA_BATCH=B        fixed speculative batch of the annealer (default: adaptive);  MCCBN_ASA_LOOKAHEAD,
 *                            MCCBN_ASA_PREFETCH_GATE: look-ahead depth (12) and the number of consecutive rejections (3)
 *                            after which look-ahead candidates' kernels are compiled in the background
 *   MCCBN_PEER_TIMEOUT_S     give up waiting for a rank's statistic vector after this many seconds (30)
 *   MCCBN_GRAPHS=0, MCCBN_FUSE_TAIL=0, MCCBN_FWD_DYNAMIC=0, MCCBN_POOL_MMA=0
 *                            A/B switches: no CUDA graphs; unfused tail kernels; static item assignment in the forward
 *                            kernel; FFMA instead of TF32-mma weighted sums in the pool kernel (fp32 rounding differs)
 *   MCCBN_JIT_MINB, MCCBN_JIT_RESIDENT, MCCBN_JIT_COND_MINB, MCCBN_JIT_PHILOX_ILP, MCCBN_JIT_PTXAS_O, MCCBN_JIT_VERBOSE
 *                            code-generation knobs of the specialised kernels (tuning runs)
 *   MCCBN_TRACE_E2E, MCCBN_ASA_TRACE
 *                            print host time per phase of the one-call path / of an annealing run to stderr
 *   MCCBN_SEGV_TRACE         install a SIGSEGV handler that prints the faulting thread's native backtrace (dev aid)
 *   MCCBN_FWD_LEAN=0         generic forward kernel: always the general in-degree path (A/B switch)
 */

/* ---- measurement helper ------------------------------------------------------------------------------------ */
/* 32-bit random words per second of a pure Philox4x32-10 generator kernel on this GPU: the INT32 ceiling against
 * which bench.py reports the E-step kernels (the path is instruction-issue bound, not HBM/tensor bound). */
double mccbn_microbench_philox(mccbn_ctx* ctx, int reps);

#ifdef __cplusplus
}
#endif
#endif /* MCCBN_B200_H */
