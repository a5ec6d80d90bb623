// r_shim.cpp -- the R-facing side of the drop-in: every `RcppExport SEXP` symbol of the reference's hot path, with the
// reference's argument lists, forwarding to the C ABI of libmccbn_b200.so (include/mccbn_b200.h).
//
// The reference's R functions reach C++ through `.Call('_name', PACKAGE = 'mccbn', ...)` and `useDynLib(mccbn)` without a
// registration table (NAMESPACE:37), i.e. R resolves the symbols below BY NAME in mccbn.so.  Building the package with
// this file instead of src/mcem_hcbn.cpp, src/asa.cpp, src/add_remove.cpp, src/backward_sampling.cpp,
// src/compatible_observations.cpp and src/mcem.cpp (and linking -lmccbn_b200) keeps every R wrapper untouched:
//     PKG_LIBS = -L$(MCCBN_B200_HOME)/mccbn_b200 -lmccbn_b200        (src/Makevars)
// Each export mirrors its reference twin: same SEXP order, same conversions (`as<float>` for lambda_s / max_lambda, the
// int -> bool copy of `obs`), same names in the returned list, and the reference's error path -- any failure becomes
// Rf_error("c++ exception: <what>") exactly like handle_exceptions() (src/mcem_hcbn.cpp:30-40).
// R is not installed in the build image, so this file is compiled only for syntax there, against the minimal
// stand-in header tests/stubs/Rcpp.h (tests/test_r_shim.py); against the real Rcpp it needs nothing else.
//
// Exports (reference file:line):
//   _MCEM_hcbn src/mcem_hcbn.cpp:804            _obs_log_likelihood :762        _importance_weight :914
//   _importance_weight_genotype :866            _complete_log_likelihood :738   _sample_genotypes :1012
//   _sample_times :1052                         _generate_genotypes :1086
//   _adaptive_simulated_annealing src/asa.cpp:596   _remove_test :667   _transitive_reduction_dag :708
//   _scale_path_to_mutation src/add_remove.cpp:218  _draw_sample :244   _generate_mutation_times :286
//   _cbn_density_log :323                       _neighbors src/backward_sampling.cpp:79      _coin_tossing :105
//   _compatible_observations src/compatible_observations.cpp:91
//   MCEM src/mcem.cpp:307                       drawHiddenVarsSamples :214      expectedTimeDifference :251
#include <Rcpp.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../include/mccbn_b200.h"

namespace {

struct ShimError {
  std::string what;
};

// non-zero return code of the C ABI -> the exception the reference would have thrown
inline void check(int rc, const char* err) {
  if (rc != MCCBN_OK) throw ShimError{std::string(err)};
}

// handle_exceptions() (src/mcem_hcbn.cpp:30-40): rethrow, report through Rf_error (does not return)
inline void raise(const ShimError& e) { Rf_error("c++ exception: %s", e.what.c_str()); }

// the legacy OT-CBN exports draw from R's global generator (runif / rexp sugar, src/mcem.cpp:30-50): the device stream
// is seeded from it, so that set.seed() keeps controlling the result
inline int seed_from_R() {
  Rcpp::RNGScope scope;
  return 1 + (int)(unif_rand() * 2147483646.0);
}

}  // namespace

#define SHIM_TRY try {
#define SHIM_CATCH                \
  }                               \
  catch (const ShimError& e) {    \
    raise(e);                     \
  }                               \
  catch (const std::exception& e) { \
    Rf_error("c++ exception: %s", e.what()); \
  }                               \
  catch (...) {                   \
    Rf_error("c++ exception (unknown reason)"); \
  }                               \
  return R_NilValue;

// ---- src/mcem_hcbn.cpp ---------------------------------------------------------------------------------------------
RcppExport SEXP _MCEM_hcbn(SEXP ilambdaSEXP, SEXP posetSEXP, SEXP obsSEXP, SEXP timesSEXP, SEXP lambda_sSEXP,
                           SEXP epsSEXP, SEXP weightsSEXP, SEXP LSEXP, SEXP samplingSEXP, SEXP max_iterSEXP,
                           SEXP update_step_sizeSEXP, SEXP tolSEXP, SEXP max_lambdaSEXP, SEXP neighborhood_distSEXP,
                           SEXP sampling_times_availableSEXP, SEXP thrdsSEXP, SEXP verboseSEXP, SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::NumericVector ilambda = Rcpp::as<Rcpp::NumericVector>(ilambdaSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::IntegerMatrix obs = Rcpp::as<Rcpp::IntegerMatrix>(obsSEXP);
  Rcpp::NumericVector times = Rcpp::as<Rcpp::NumericVector>(timesSEXP);
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const double eps = Rcpp::as<double>(epsSEXP);
  Rcpp::NumericVector weights = Rcpp::as<Rcpp::NumericVector>(weightsSEXP);
  const unsigned L = Rcpp::as<unsigned int>(LSEXP);
  const std::string sampling = Rcpp::as<std::string>(samplingSEXP);
  const unsigned max_iter = Rcpp::as<unsigned int>(max_iterSEXP);
  const unsigned update_step_size = Rcpp::as<unsigned int>(update_step_sizeSEXP);
  const double tol = Rcpp::as<double>(tolSEXP);
  const float max_lambda = Rcpp::as<float>(max_lambdaSEXP);
  const unsigned neighborhood_dist = Rcpp::as<unsigned int>(neighborhood_distSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int thrds = Rcpp::as<int>(thrdsSEXP);
  const bool verbose = Rcpp::as<bool>(verboseSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int N = obs.nrow(), p = poset.nrow();
  Rcpp::NumericVector lambda(p);
  double eps_out = 0.0, llhood = 0.0;
  char err[512] = {0};
  check(mccbn_MCEM_hcbn(NULL, ilambda.begin(), poset.begin(), obs.begin(), times.begin(), lambda_s, eps, weights.begin(),
                        L, sampling.c_str(), max_iter, update_step_size, tol, max_lambda, neighborhood_dist, avail, thrds,
                        verbose, seed, N, p, lambda.begin(), &eps_out, &llhood, NULL, err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("lambda") = lambda, Rcpp::Named("eps") = eps_out,
                            Rcpp::Named("llhood") = llhood);
  SHIM_CATCH
}

RcppExport SEXP _obs_log_likelihood(SEXP obsSEXP, SEXP posetSEXP, SEXP lambdaSEXP, SEXP epsSEXP, SEXP weightsSEXP,
                                    SEXP timesSEXP, SEXP LSEXP, SEXP samplingSEXP, SEXP neighborhood_distSEXP,
                                    SEXP lambda_sSEXP, SEXP sampling_times_availableSEXP, SEXP thrdsSEXP,
                                    SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix obs = Rcpp::as<Rcpp::IntegerMatrix>(obsSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  const double eps = Rcpp::as<double>(epsSEXP);
  Rcpp::NumericVector weights = Rcpp::as<Rcpp::NumericVector>(weightsSEXP);
  Rcpp::NumericVector times = Rcpp::as<Rcpp::NumericVector>(timesSEXP);
  const unsigned L = Rcpp::as<unsigned int>(LSEXP);
  const std::string sampling = Rcpp::as<std::string>(samplingSEXP);
  const unsigned neighborhood_dist = Rcpp::as<unsigned int>(neighborhood_distSEXP);
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int thrds = Rcpp::as<int>(thrdsSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  double llhood = 0.0;
  char err[512] = {0};
  check(mccbn_obs_log_likelihood(NULL, obs.begin(), poset.begin(), lambda.begin(), eps, weights.begin(), times.begin(), L,
                                 sampling.c_str(), neighborhood_dist, lambda_s, avail, thrds, seed, obs.nrow(),
                                 poset.nrow(), &llhood, NULL, err, sizeof err),
        err);
  return Rcpp::wrap(llhood);
  SHIM_CATCH
}

RcppExport SEXP _importance_weight(SEXP obsSEXP, SEXP LSEXP, SEXP posetSEXP, SEXP lambdaSEXP, SEXP epsSEXP,
                                   SEXP timesSEXP, SEXP samplingSEXP, SEXP neighborhood_distSEXP, SEXP lambda_sSEXP,
                                   SEXP sampling_times_availableSEXP, SEXP thrdsSEXP, SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix obs = Rcpp::as<Rcpp::IntegerMatrix>(obsSEXP);
  const unsigned L = Rcpp::as<unsigned int>(LSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  const double eps = Rcpp::as<double>(epsSEXP);
  Rcpp::NumericVector times = Rcpp::as<Rcpp::NumericVector>(timesSEXP);
  const std::string sampling = Rcpp::as<std::string>(samplingSEXP);
  const unsigned neighborhood_dist = Rcpp::as<unsigned int>(neighborhood_distSEXP);
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int thrds = Rcpp::as<int>(thrdsSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int N = obs.nrow(), p = poset.nrow();
  Rcpp::NumericVector w_sum(N), w_sum_sqrt(N), dist(N);
  Rcpp::IntegerVector L_eff(N);
  Rcpp::NumericMatrix Tdiff(N, p);
  char err[512] = {0};
  check(mccbn_importance_weight(NULL, obs.begin(), L, poset.begin(), lambda.begin(), eps, times.begin(), sampling.c_str(),
                                neighborhood_dist, lambda_s, avail, thrds, seed, N, p, w_sum.begin(), w_sum_sqrt.begin(),
                                L_eff.begin(), dist.begin(), Tdiff.begin(), NULL, err, sizeof err),
        err);
  if (sampling == "backward" || sampling == "bernoulli")  // :999-1005
    return Rcpp::List::create(Rcpp::Named("w") = w_sum, Rcpp::Named("w_sqrt") = w_sum_sqrt,
                              Rcpp::Named("L_eff") = L_eff, Rcpp::Named("dist") = dist, Rcpp::Named("Tdiff") = Tdiff);
  return Rcpp::List::create(Rcpp::Named("w") = w_sum, Rcpp::Named("w_sqrt") = w_sum_sqrt, Rcpp::Named("dist") = dist,
                            Rcpp::Named("Tdiff") = Tdiff);
  SHIM_CATCH
}

// scale_cumulative / d_pool / Tdiff_pool are inputs of the reference's add-remove and pool branches for ONE genotype; the
// device derives the add-remove scales itself, and the per-sample seam is exposed for the forward scheme only (the C
// entry returns MCCBN_ERR_ARG otherwise, which surfaces here as an R error).
RcppExport SEXP _importance_weight_genotype(SEXP genotypeSEXP, SEXP LSEXP, SEXP posetSEXP, SEXP lambdaSEXP,
                                            SEXP epsSEXP, SEXP timeSEXP, SEXP samplingSEXP,
                                            SEXP scale_cumulativeSEXP, SEXP d_poolSEXP, SEXP Tdiff_poolSEXP,
                                            SEXP neighborhood_distSEXP, SEXP lambda_sSEXP,
                                            SEXP sampling_times_availableSEXP, SEXP seedSEXP) {
  SHIM_TRY
  (void)scale_cumulativeSEXP;
  (void)d_poolSEXP;
  (void)Tdiff_poolSEXP;
  Rcpp::IntegerVector genotype = Rcpp::as<Rcpp::IntegerVector>(genotypeSEXP);
  const unsigned L = Rcpp::as<unsigned int>(LSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  const double eps = Rcpp::as<double>(epsSEXP);
  const double time = Rcpp::as<double>(timeSEXP);
  const std::string sampling = Rcpp::as<std::string>(samplingSEXP);
  const unsigned neighborhood_dist = Rcpp::as<unsigned int>(neighborhood_distSEXP);
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int p = poset.nrow();
  std::vector<double> w(L), Tdiff((size_t)L * p);
  std::vector<int> dist(L);
  int L_out = 0;
  char err[512] = {0};
  check(mccbn_importance_weight_genotype(NULL, genotype.begin(), L, poset.begin(), lambda.begin(), eps, time,
                                         sampling.c_str(), neighborhood_dist, lambda_s, avail, seed, p, &L_out, w.data(),
                                         dist.data(), Tdiff.data(), NULL, err, sizeof err),
        err);
  Rcpp::NumericVector w_r(L_out);
  Rcpp::IntegerVector d_r(L_out);
  Rcpp::NumericMatrix T_r(L_out, p);
  for (int l = 0; l < L_out; ++l) {
    w_r[l] = w[l];
    d_r[l] = dist[l];
    for (int j = 0; j < p; ++j) T_r(l, j) = Tdiff[(size_t)j * L + l];
  }
  return Rcpp::List::create(Rcpp::Named("w") = w_r, Rcpp::Named("dist") = d_r, Rcpp::Named("Tdiff") = T_r);
  SHIM_CATCH
}

RcppExport SEXP _complete_log_likelihood(SEXP lambdaSEXP, SEXP epsSEXP, SEXP TdiffSEXP, SEXP distSEXP,
                                         SEXP weightsSEXP) {
  SHIM_TRY
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  const double eps = Rcpp::as<double>(epsSEXP);
  Rcpp::NumericMatrix Tdiff = Rcpp::as<Rcpp::NumericMatrix>(TdiffSEXP);
  Rcpp::NumericVector dist = Rcpp::as<Rcpp::NumericVector>(distSEXP);
  Rcpp::NumericVector weights = Rcpp::as<Rcpp::NumericVector>(weightsSEXP);
  double res = 0.0;
  char err[512] = {0};
  check(mccbn_complete_log_likelihood(lambda.begin(), (int)lambda.size(), eps, Tdiff.begin(), Tdiff.nrow(), dist.begin(),
                                      weights.begin(), &res, err, sizeof err),
        err);
  return Rcpp::wrap(res);
  SHIM_CATCH
}

RcppExport SEXP _sample_genotypes(SEXP NSEXP, SEXP posetSEXP, SEXP lambdaSEXP, SEXP T_eventsSEXP, SEXP T_samplingSEXP,
                                  SEXP lambda_sSEXP, SEXP sampling_times_availableSEXP, SEXP seedSEXP) {
  SHIM_TRY
  (void)T_eventsSEXP;  // output buffer in the reference (N x p, overwritten with the waiting times)
  const unsigned N = Rcpp::as<unsigned int>(NSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  Rcpp::NumericVector T_sampling = Rcpp::clone(Rcpp::as<Rcpp::NumericVector>(T_samplingSEXP));
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int p = poset.nrow();
  Rcpp::LogicalMatrix samples(N, p);
  Rcpp::NumericMatrix Tdiff(N, p);
  char err[512] = {0};
  check(mccbn_sample_genotypes(NULL, N, poset.begin(), lambda.begin(), T_sampling.begin(), lambda_s, avail, seed, p,
                               samples.begin(), Tdiff.begin(), err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("samples") = samples, Rcpp::Named("Tdiff") = Tdiff,
                            Rcpp::Named("sampling_time") = T_sampling);
  SHIM_CATCH
}

RcppExport SEXP _sample_times(SEXP NSEXP, SEXP posetSEXP, SEXP lambdaSEXP, SEXP T_eventsSEXP, SEXP seedSEXP) {
  SHIM_TRY
  (void)T_eventsSEXP;
  const unsigned N = Rcpp::as<unsigned int>(NSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int p = poset.nrow();
  Rcpp::NumericMatrix times(N, p), Tdiff(N, p);
  char err[512] = {0};
  check(mccbn_sample_times(NULL, N, poset.begin(), lambda.begin(), seed, p, times.begin(), Tdiff.begin(), err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("mutation_times") = times, Rcpp::Named("Tdiff") = Tdiff);
  SHIM_CATCH
}

RcppExport SEXP _generate_genotypes(SEXP T_events_sumSEXP, SEXP posetSEXP, SEXP T_samplingSEXP, SEXP lambda_sSEXP,
                                    SEXP sampling_times_availableSEXP, SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::NumericMatrix T_events_sum = Rcpp::as<Rcpp::NumericMatrix>(T_events_sumSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector T_sampling = Rcpp::clone(Rcpp::as<Rcpp::NumericVector>(T_samplingSEXP));
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int N = T_events_sum.nrow(), p = poset.nrow();
  Rcpp::LogicalMatrix samples(N, p);
  char err[512] = {0};
  check(mccbn_generate_genotypes(NULL, T_events_sum.begin(), poset.begin(), T_sampling.begin(), lambda_s, avail, seed, N, p,
                                 samples.begin(), err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("samples") = samples, Rcpp::Named("sampling_time") = T_sampling);
  SHIM_CATCH
}

// ---- src/asa.cpp ----------------------------------------------------------------------------------------------------
RcppExport SEXP _adaptive_simulated_annealing(SEXP posetSEXP, SEXP obsSEXP, SEXP timesSEXP, SEXP lambda_sSEXP,
                                              SEXP weightsSEXP, SEXP LSEXP, SEXP samplingSEXP, SEXP max_iter_EMSEXP,
                                              SEXP update_step_sizeSEXP, SEXP tolSEXP, SEXP max_lambdaSEXP, SEXP T0SEXP,
                                              SEXP adap_rateSEXP, SEXP acceptance_rateSEXP, SEXP step_sizeSEXP,
                                              SEXP max_iter_ASASEXP, SEXP neighborhood_distSEXP, SEXP adaptiveSEXP,
                                              SEXP outdirSEXP, SEXP sampling_times_availableSEXP, SEXP thrdsSEXP,
                                              SEXP verboseSEXP, SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::IntegerMatrix obs = Rcpp::as<Rcpp::IntegerMatrix>(obsSEXP);
  Rcpp::NumericVector times = Rcpp::as<Rcpp::NumericVector>(timesSEXP);
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  Rcpp::NumericVector weights = Rcpp::as<Rcpp::NumericVector>(weightsSEXP);
  const unsigned L = Rcpp::as<unsigned int>(LSEXP);
  const std::string sampling = Rcpp::as<std::string>(samplingSEXP);
  const unsigned max_iter_EM = Rcpp::as<unsigned int>(max_iter_EMSEXP);
  const unsigned update_step_size = Rcpp::as<unsigned int>(update_step_sizeSEXP);
  const double tol = Rcpp::as<double>(tolSEXP);
  const float max_lambda = Rcpp::as<float>(max_lambdaSEXP);
  const float T0 = Rcpp::as<float>(T0SEXP);
  const float adap_rate = Rcpp::as<float>(adap_rateSEXP);
  const float acceptance_rate = Rcpp::as<float>(acceptance_rateSEXP);
  const int step_size = Rcpp::as<int>(step_sizeSEXP);
  const unsigned max_iter_ASA = Rcpp::as<unsigned int>(max_iter_ASASEXP);
  const unsigned neighborhood_dist = Rcpp::as<unsigned int>(neighborhood_distSEXP);
  const bool adaptive = Rcpp::as<bool>(adaptiveSEXP);
  const std::string outdir = Rcpp::as<std::string>(outdirSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int thrds = Rcpp::as<int>(thrdsSEXP);
  const bool verbose = Rcpp::as<bool>(verboseSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int N = obs.nrow(), p = poset.nrow();
  Rcpp::NumericVector lambda(p);
  Rcpp::IntegerMatrix poset_out(p, p);
  double eps = 0.0, llhood = 0.0;
  char err[512] = {0};
  check(mccbn_adaptive_simulated_annealing(NULL, poset.begin(), obs.begin(), times.begin(), lambda_s, weights.begin(), L,
                                           sampling.c_str(), max_iter_EM, update_step_size, tol, max_lambda, T0,
                                           adap_rate, acceptance_rate, step_size, max_iter_ASA, neighborhood_dist,
                                           adaptive, outdir.c_str(), avail, thrds, verbose, seed, N, p, lambda.begin(),
                                           &eps, poset_out.begin(), &llhood, err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("lambda") = lambda, Rcpp::Named("eps") = eps,
                            Rcpp::Named("poset") = poset_out, Rcpp::Named("llhood") = llhood);
  SHIM_CATCH
}

// debug exports of the reference: print, return 0
RcppExport SEXP _transitive_reduction_dag(SEXP posetSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  const int p = poset.nrow();
  std::vector<int> out((size_t)p * p);
  int flags = 0;
  char err[512] = {0};
  check(mccbn_poset_edit(poset.begin(), p, /*transitive_reduction_dag*/ 4, 0, 0, out.data(), &flags, err, sizeof err), err);
  if (!(flags & 4)) {
    std::printf("Not transitively reduced\n");
    for (int u = 0; u < p; ++u)
      for (int v = 0; v < p; ++v)
        if (out[(size_t)v * p + u] == 1) std::printf("%d --> %d\n", u, v);
  }
  return Rcpp::wrap(0);
  SHIM_CATCH
}

RcppExport SEXP _remove_test(SEXP posetSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  const int p = poset.nrow();
  char err[512] = {0};
  for (int u = 0; u < p; ++u)
    for (int v = 0; v < p; ++v) {
      if (poset(u, v) != 1) continue;
      std::printf("Testing edge: %d --> %d\n", u, v);
      std::vector<int> removed((size_t)p * p), reduced((size_t)p * p);
      int flags = 0;
      check(mccbn_poset_edit(poset.begin(), p, /*remove_edge*/ 1, u, v, removed.data(), &flags, err, sizeof err), err);
      check(mccbn_poset_edit(removed.data(), p, /*transitive_reduction_dag*/ 4, 0, 0, reduced.data(), &flags, err, sizeof err),
            err);
      if (!(flags & 4)) std::printf("Not transitively reduced\n");
    }
  return Rcpp::wrap(0);
  SHIM_CATCH
}

// ---- src/add_remove.cpp ---------------------------------------------------------------------------------------------
RcppExport SEXP _scale_path_to_mutation(SEXP posetSEXP, SEXP lambdaSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  Rcpp::NumericVector out(poset.nrow());
  char err[512] = {0};
  check(mccbn_scale_path_to_mutation(poset.begin(), poset.nrow(), lambda.begin(), out.begin(), err, sizeof err), err);
  return out;
  SHIM_CATCH
}

RcppExport SEXP _draw_sample(SEXP genotypeSEXP, SEXP posetSEXP, SEXP moveSEXP, SEXP q_probSEXP, SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::IntegerVector genotype = Rcpp::as<Rcpp::IntegerVector>(genotypeSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  const unsigned move = Rcpp::as<unsigned int>(moveSEXP);
  Rcpp::NumericVector q_prob = Rcpp::as<Rcpp::NumericVector>(q_probSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int p = poset.nrow();
  Rcpp::LogicalVector sample(p);
  double q_choice = 0.0;
  char err[512] = {0};
  check(mccbn_draw_sample(genotype.begin(), poset.begin(), p, move, q_prob.begin(), seed, sample.begin(), &q_choice, err,
                          sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("sample") = sample, Rcpp::Named("q_choice") = q_choice);
  SHIM_CATCH
}

RcppExport SEXP _generate_mutation_times(SEXP obsSEXP, SEXP posetSEXP, SEXP lambdaSEXP, SEXP timeSEXP,
                                         SEXP lambda_sSEXP, SEXP sampling_times_availableSEXP, SEXP seedSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix obs = Rcpp::as<Rcpp::IntegerMatrix>(obsSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  Rcpp::NumericVector time = Rcpp::clone(Rcpp::as<Rcpp::NumericVector>(timeSEXP));
  const float lambda_s = Rcpp::as<float>(lambda_sSEXP);
  const bool avail = Rcpp::as<bool>(sampling_times_availableSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  const int N = obs.nrow(), p = poset.nrow();
  Rcpp::NumericMatrix Tdiff(N, p);
  Rcpp::NumericVector dens(N);
  char err[512] = {0};
  check(mccbn_generate_mutation_times(NULL, obs.begin(), poset.begin(), lambda.begin(), time.begin(), lambda_s, avail, seed,
                                      N, p, Tdiff.begin(), dens.begin(), err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("Tdiff") = Tdiff, Rcpp::Named("proposal_density") = dens);
  SHIM_CATCH
}

RcppExport SEXP _cbn_density_log(SEXP timeSEXP, SEXP lambdaSEXP) {
  SHIM_TRY
  Rcpp::NumericMatrix time = Rcpp::as<Rcpp::NumericMatrix>(timeSEXP);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambdaSEXP);
  Rcpp::NumericVector out(time.nrow());
  char err[512] = {0};
  check(mccbn_cbn_density_log(time.begin(), time.nrow(), (int)lambda.size(), lambda.begin(), out.begin(), err, sizeof err),
        err);
  return out;
  SHIM_CATCH
}

// ---- src/backward_sampling.cpp --------------------------------------------------------------------------------------
RcppExport SEXP _neighbors(SEXP pSEXP, SEXP kSEXP, SEXP genotypeSEXP) {
  SHIM_TRY
  const unsigned p = Rcpp::as<unsigned int>(pSEXP);
  const unsigned k = Rcpp::as<unsigned int>(kSEXP);
  Rcpp::IntegerVector genotype = Rcpp::as<Rcpp::IntegerVector>(genotypeSEXP);
  int n = 0;
  char err[512] = {0};
  check(mccbn_neighbors((int)p, (int)k, genotype.begin(), NULL, NULL, &n, err, sizeof err), err);
  Rcpp::LogicalMatrix samples(n, (int)p);
  check(mccbn_neighbors((int)p, (int)k, genotype.begin(), samples.begin(), NULL, &n, err, sizeof err), err);
  return samples;
  SHIM_CATCH
}

RcppExport SEXP _coin_tossing(SEXP epsSEXP, SEXP LSEXP, SEXP genotypeSEXP, SEXP seedSEXP) {
  SHIM_TRY
  const double eps = Rcpp::as<double>(epsSEXP);
  const unsigned L = Rcpp::as<unsigned int>(LSEXP);
  Rcpp::IntegerVector genotype = Rcpp::as<Rcpp::IntegerVector>(genotypeSEXP);
  const int seed = Rcpp::as<int>(seedSEXP);
  Rcpp::LogicalMatrix samples(L, (int)genotype.size());
  char err[512] = {0};
  check(mccbn_coin_tossing(NULL, eps, L, genotype.begin(), (int)genotype.size(), seed, samples.begin(), err, sizeof err),
        err);
  return samples;
  SHIM_CATCH
}

// ---- src/compatible_observations.cpp ---------------------------------------------------------------------------------
RcppExport SEXP _compatible_observations(SEXP obsSEXP, SEXP posetSEXP) {
  SHIM_TRY
  Rcpp::IntegerMatrix obs = Rcpp::as<Rcpp::IntegerMatrix>(obsSEXP);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(posetSEXP);
  Rcpp::LogicalVector out(obs.nrow());
  char err[512] = {0};
  check(mccbn_compatible_observations(NULL, obs.begin(), poset.begin(), obs.nrow(), poset.nrow(), out.begin(), NULL, NULL,
                                      err, sizeof err),
        err);
  return out;
  SHIM_CATCH
}

// ---- src/mcem.cpp (legacy OT-CBN; no try/catch in the reference -- errors still must not unwind through R) ------------
RcppExport SEXP MCEM(SEXP ilambda_, SEXP poset_, SEXP O_, SEXP times_, SEXP max_iter_, SEXP alpha_, SEXP topo_path_,
                     SEXP weights_, SEXP nrOfSamples_, SEXP verbose_, SEXP max_lambda_, SEXP lambda_s_,
                     SEXP sampling_time_available_) {
  SHIM_TRY
  Rcpp::NumericVector ilambda = Rcpp::as<Rcpp::NumericVector>(ilambda_);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(poset_);
  Rcpp::IntegerMatrix O = Rcpp::as<Rcpp::IntegerMatrix>(O_);
  Rcpp::NumericVector times = Rcpp::as<Rcpp::NumericVector>(times_);
  const int max_iter = Rcpp::as<int>(max_iter_);
  const float alpha = Rcpp::as<float>(alpha_);
  Rcpp::IntegerVector topo_path = Rcpp::as<Rcpp::IntegerVector>(topo_path_);
  Rcpp::NumericVector weights = Rcpp::as<Rcpp::NumericVector>(weights_);
  const int nrOfSamples = Rcpp::as<int>(nrOfSamples_);
  const bool verbose = Rcpp::as<bool>(verbose_);
  const float max_lambda = Rcpp::as<float>(max_lambda_);
  const float lambda_s = Rcpp::as<float>(lambda_s_);
  const bool avail = Rcpp::as<bool>(sampling_time_available_);
  const int N = O.nrow(), p = poset.nrow();
  Rcpp::NumericVector lambda(p);
  double llhood = 0.0;
  char err[512] = {0};
  check(mccbn_MCEM(NULL, ilambda.begin(), poset.begin(), O.begin(), times.begin(), max_iter, alpha, topo_path.begin(),
                   weights.begin(), nrOfSamples, verbose, max_lambda, lambda_s, avail, seed_from_R(), N, p, lambda.begin(),
                   &llhood, err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("lambda") = lambda, Rcpp::Named("llhood") = llhood);
  SHIM_CATCH
}

RcppExport SEXP drawHiddenVarsSamples(SEXP N_, SEXP O_vec_, SEXP time_, SEXP lambda_, SEXP poset_, SEXP topo_path_,
                                      SEXP lambda_s_, SEXP sampling_time_available_) {
  SHIM_TRY
  const int N = Rcpp::as<int>(N_);
  Rcpp::IntegerVector O_vec = Rcpp::as<Rcpp::IntegerVector>(O_vec_);
  const double time = Rcpp::as<double>(time_);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambda_);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(poset_);
  Rcpp::IntegerVector topo_path = Rcpp::as<Rcpp::IntegerVector>(topo_path_);
  const float lambda_s = Rcpp::as<float>(lambda_s_);
  const bool avail = Rcpp::as<bool>(sampling_time_available_);
  const int p = (int)lambda.size();
  Rcpp::NumericMatrix T(N, p);
  Rcpp::NumericVector densities(N);
  char err[512] = {0};
  check(mccbn_drawHiddenVarsSamples(NULL, N, O_vec.begin(), time, lambda.begin(), poset.begin(), topo_path.begin(), lambda_s,
                                    avail, seed_from_R(), p, T.begin(), densities.begin(), NULL, err, sizeof err),
        err);
  return Rcpp::List::create(Rcpp::Named("T") = T, Rcpp::Named("densities") = densities);
  SHIM_CATCH
}

RcppExport SEXP expectedTimeDifference(SEXP nrOfSamples_, SEXP O_vec_, SEXP time_, SEXP lambda_, SEXP poset_,
                                       SEXP topo_path_, SEXP lambda_s_, SEXP sampling_time_available_) {
  SHIM_TRY
  const int nrOfSamples = Rcpp::as<int>(nrOfSamples_);
  Rcpp::IntegerVector O_vec = Rcpp::as<Rcpp::IntegerVector>(O_vec_);
  const double time = Rcpp::as<double>(time_);
  Rcpp::NumericVector lambda = Rcpp::as<Rcpp::NumericVector>(lambda_);
  Rcpp::IntegerMatrix poset = Rcpp::as<Rcpp::IntegerMatrix>(poset_);
  Rcpp::IntegerVector topo_path = Rcpp::as<Rcpp::IntegerVector>(topo_path_);
  const float lambda_s = Rcpp::as<float>(lambda_s_);
  const bool avail = Rcpp::as<bool>(sampling_time_available_);
  const int p = (int)lambda.size();
  Rcpp::NumericVector out(p);
  char err[512] = {0};
  check(mccbn_expectedTimeDifference(NULL, nrOfSamples, O_vec.begin(), time, lambda.begin(), poset.begin(),
                                     topo_path.begin(), lambda_s, avail, seed_from_R(), p, out.begin(), err, sizeof err),
        err);
  return out;
  SHIM_CATCH
}
