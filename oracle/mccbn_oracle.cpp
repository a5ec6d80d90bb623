// mccbn_oracle.cpp -- CPU ORACLE for the MC-CBN H-CBN2 Monte-Carlo-EM hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
// (libmccbn_b200.so) never links, loads or calls anything in this directory.
//
// What it is: a dependency-free C++17/OpenMP restatement of the reference's C++ algorithms
// (no Rcpp / Eigen / Boost -- none of them exists in this image, so the reference itself cannot
// be compiled; see DESIGN.md "Oracle").  Every function cites the reference file:line it follows
// (paths relative to /root/reference).  It keeps the reference's cost structure (std::mt19937
// with 53-bit uniforms, log1p/expm1 per draw, per-observation L x p temporaries, pow per sample,
// "omp parallel for schedule(static)" over observations with per-thread generators reseeded from
// the master every EM iteration) so that it doubles as the timed CPU baseline.
//
// PARITY STATUS: "parity unpinned" at the fixture level -- the reference ships no golden vectors,
// known-answer tests or fixtures for this path and cannot be run here.
// The oracle is instead pinned by model-derived exact answers (tests/test_oracle_exact.py: K1
// closed form on the empty poset, K2 lattice recursion, K3 density identity, K4 neighbourhood
// enumeration) and by the reference's own statistical validation methods.
// and by a second, independent reading of the mathematics: the reference's pure-R prototype hcbn_functions.r
// restated in numpy (tests/rproto.py, tests/test_second_statement.py: forward weights, conditional proposal
// and density, complete-data log-likelihood, M-step agree with this file to 1e-10 .. 1e-12 on random inputs).
//
// Matrices use R / Eigen column-major layout: element (r, c) of an R x C matrix is a[c*R + r].

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace ref {

static const double kInf = std::numeric_limits<double>::infinity();
static const double kDblEps = std::numeric_limits<double>::epsilon();

// ----------------------------------------------------------------------------------------------
// small column-major matrix (stands in for Eigen::Matrix<T, Dynamic, Dynamic>)
// ----------------------------------------------------------------------------------------------
template <class T>
struct Mat {
  int rows = 0, cols = 0;
  std::vector<T> a;
  Mat() {}
  Mat(int r, int c, T v = T()) : rows(r), cols(c), a((size_t)r * (size_t)c, v) {}
  inline T& operator()(int r, int c) { return a[(size_t)c * rows + r]; }
  inline const T& operator()(int r, int c) const { return a[(size_t)c * rows + r]; }
  T* col(int c) { return a.data() + (size_t)c * rows; }
  const T* col(int c) const { return a.data() + (size_t)c * rows; }
};
typedef Mat<double> MatD;
typedef Mat<unsigned char> MatB;  // bool matrix (MatrixXb, src/mcem.hpp:28)

// ----------------------------------------------------------------------------------------------
// RNG: StdRng (src/rng_utils.hpp:87-120) and Context (src/mcem.hpp:49-72)
// ----------------------------------------------------------------------------------------------
struct StdRng {
  typedef std::mt19937::result_type result_type;
  std::mt19937 eng;
  // seeding chain ranlux24_base(seed) -> 624 ints -> seed_seq -> mt19937 (src/rng_utils.hpp:89-98)
  explicit StdRng(unsigned int seed) {
    std::ranlux24_base seeder(seed);
    std::array<int, std::mt19937::state_size> words;
    for (auto& w : words) w = (int)seeder();
    std::seed_seq seq(words.begin(), words.end());
    eng.seed(seq);
  }
  result_type operator()() { return eng(); }
  static constexpr result_type min() { return std::mt19937::min(); }
  static constexpr result_type max() { return std::mt19937::max(); }
};

struct Context {
  StdRng rng;
  bool verbose;
  Context(int seed, bool v = false) : rng((unsigned int)seed), verbose(v) {}
  // one generator per thread, each seeded with one 32-bit master output (src/mcem.hpp:62-68)
  std::vector<std::unique_ptr<StdRng>> get_auxiliary_rngs(int n) {
    std::vector<std::unique_ptr<StdRng>> v;
    v.reserve(n);
    for (int t = 0; t < n; ++t) v.emplace_back(new StdRng(rng()));
    return v;
  }
};

// rtexp<StdRng> (src/rng_utils.cpp:44-56): x = -log1p(u * expm1(-cutoff*rate)) / rate
static inline double rtexp1(double rate, double cutoff, StdRng& rng) {
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double u = U(rng.eng);
  return -std::log1p(u * std::expm1(-cutoff * rate)) / rate;
}
static inline double rtexp_from_u(double rate, double cutoff, double u) {
  return -std::log1p(u * std::expm1(-cutoff * rate)) / rate;
}
// rexp (src/rng_utils.hpp:24-30): rtexp with cutoff = +inf, N sequential draws
static void rexp_fill(double* out, int N, double rate, StdRng& rng) {
  for (int i = 0; i < N; ++i) out[i] = rtexp1(rate, kInf, rng);
}

// rdiscrete_std (src/mcem_hcbn.cpp:44-54)
static std::vector<int> rdiscrete_std(unsigned N, const std::vector<double>& weights, StdRng& rng) {
  std::discrete_distribution<int> D(weights.begin(), weights.end());
  std::vector<int> r(N);
  for (unsigned i = 0; i < N; ++i) r[i] = D(rng);
  return r;
}

// ----------------------------------------------------------------------------------------------
// Model (src/mcem.hpp:74-198, src/model_impl.cpp)
// The poset is a directed graph over VERTICES 0..p-1; vertex v carries event_id[v] (identity until
// an ASA swap_node).  All genotype-column indexing goes through event ids.
// ----------------------------------------------------------------------------------------------
struct Model {
  int p = 0;
  std::vector<std::set<int>> out, in;  // vertex adjacency (ordered sets: deterministic iteration)
  std::vector<int> event_id;           // src/mcem.hpp:36-38,105-108
  std::vector<int> topo_path;          // REVERSE topological order, as Boost writes it (model_impl.cpp:51-69)
  bool cycle = false, reduction_flag = false;
  std::vector<double> lambda;
  float lambda_s = 1.0f;  // float! (src/mcem.hpp:190)
  double epsilon = std::numeric_limits<double>::quiet_NaN();
  double llhood = 0.0;
  std::vector<std::unordered_set<int>> children;  // transitive successors by vertex
  bool update_children_flag = true;
  std::vector<int> node_idx;  // event id -> vertex
  bool update_node_idx_flag = false;

  Model() {}
  // Model(edge_list, p, lambda_s) (src/mcem.hpp:99-110)
  Model(const std::vector<std::pair<int, int>>& edges, int p_, float ls = 1.0f)
      : p(p_), out(p_), in(p_), event_id(p_), lambda(p_, 0.0), lambda_s(ls), children(p_), node_idx(p_) {
    for (auto& e : edges) {
      out[e.first].insert(e.second);
      in[e.second].insert(e.first);
    }
    for (int i = 0; i < p; ++i) event_id[i] = node_idx[i] = i;
  }
  int num_edges() const {
    int n = 0;
    for (auto& s : out) n += (int)s.size();
    return n;
  }
  // update_lambda (model_impl.cpp:21-26): clamp to max_lambda (float) when above it or non-finite
  void update_lambda(const std::vector<double>& l, float max_lambda) {
    lambda = l;
    for (size_t j = 0; j < l.size(); ++j)
      if (l[j] > max_lambda || !std::isfinite(l[j])) lambda[j] = max_lambda;
  }
  // update_epsilon (model_impl.cpp:32-34)
  void update_epsilon(double e, double min_eps) { epsilon = std::max(std::min(e, 1 - min_eps), min_eps); }

  // has_cycles (model_impl.cpp:40-49): cycle <=> number of strongly connected components != p.
  // (Tarjan.)  NB a self-loop keeps #SCC == p; the reference then dies inside Boost's
  // topological_sort ("The graph must be a DAG"); we report it as a cycle as well.
  void has_cycles() {
    std::vector<int> idx(p, -1), low(p, 0), stack;
    std::vector<char> on(p, 0);
    int counter = 0, nscc = 0;
    std::function<void(int)> dfs = [&](int v) {
      idx[v] = low[v] = counter++;
      stack.push_back(v);
      on[v] = 1;
      for (int w : out[v]) {
        if (idx[w] < 0) {
          dfs(w);
          low[v] = std::min(low[v], low[w]);
        } else if (on[w]) {
          low[v] = std::min(low[v], idx[w]);
        }
      }
      if (low[v] == idx[v]) {
        ++nscc;
        for (;;) {
          int w = stack.back();
          stack.pop_back();
          on[w] = 0;
          if (w == v) break;
        }
      }
    };
    for (int v = 0; v < p; ++v)
      if (idx[v] < 0) dfs(v);
    if (nscc != p) cycle = true;
    for (int v = 0; v < p; ++v)
      if (out[v].count(v)) cycle = true;
  }
  // topological_sort (model_impl.cpp:51-69): DFS post-order == reverse topological order
  void topological_sort() {
    topo_path.clear();
    std::vector<char> seen(p, 0);
    std::function<void(int)> dfs = [&](int v) {
      seen[v] = 1;
      for (int w : out[v])
        if (!seen[w]) dfs(w);
      topo_path.push_back(v);
    };
    for (int v = 0; v < p; ++v)
      if (!seen[v]) dfs(v);
  }
  // get_direct_predecessors (model_impl.cpp:72-87): parents as EVENT ids, indexed by event id
  std::vector<std::vector<int>> get_direct_predecessors() const {
    std::vector<std::vector<int>> parents(p);
    for (auto v = topo_path.rbegin(); v != topo_path.rend(); ++v)
      for (int u : in[*v]) parents[event_id[*v]].push_back(event_id[u]);
    return parents;
  }
  void update_node_idx() {  // model_impl.cpp:318-322
    for (int i = 0; i < p; ++i) node_idx[event_id[i]] = i;
    update_node_idx_flag = false;
  }

  // ---- edit operations of the annealer (src/model_impl.cpp:89-316) -------------------------------------------
  bool raw_add_edge(int u, int v) {  // boost::add_edge on a setS out-edge list: false if the edge exists
    if (!out[u].insert(v).second) return false;
    in[v].insert(u);
    return true;
  }
  void raw_remove_edge(int u, int v) {
    out[u].erase(v);
    in[v].erase(u);
  }
  // set_children (:90-107): all successors per vertex, visiting vertices in reverse topological order
  void set_children() {
    for (int u : topo_path) {
      children[u].clear();
      for (int v : out[u]) {
        children[u].insert(v);
        children[u].insert(children[v].begin(), children[v].end());
      }
    }
    update_children_flag = false;
  }
  // update_children (:110-122)
  void update_children(int u) {
    children[u].clear();
    for (int v : out[u]) {
      children[u].insert(v);
      children[u].insert(children[v].begin(), children[v].end());
    }
  }
  // transitive_reduction_dag (:148-190): drop every edge u -> v whose target is already reachable through an
  // earlier (in topological order) direct successor of u; reduction_flag = "nothing had to be dropped"
  void transitive_reduction_dag() {
    reduction_flag = true;
    std::vector<int> rank(p, 0);
    int n = 0;
    for (auto it = topo_path.rbegin(); it != topo_path.rend(); ++it, ++n) rank[*it] = n;
    std::vector<std::vector<int>> succ(p);
    for (int v = 0; v < p; ++v) {
      succ[v].assign(out[v].begin(), out[v].end());
      std::stable_sort(succ[v].begin(), succ[v].end(), [&](int a, int b) { return rank[a] < rank[b]; });
    }
    std::vector<std::unordered_set<int>> all_succ(p);
    for (int u : topo_path) {
      for (int v : succ[u]) {
        if (all_succ[u].count(v)) {
          remove_edge(u, v);
          reduction_flag = false;
        } else {
          all_succ[u].insert(v);
          all_succ[u].insert(all_succ[v].begin(), all_succ[v].end());
        }
      }
    }
  }
  // add_edge (:193-212)
  bool add_edge(int u, int v) {
    const bool added = raw_add_edge(u, v);
    if (added) {
      has_cycles();
      if (!cycle) {
        topological_sort();
        transitive_reduction_dag();
        update_children_flag = true;
      }
    }
    return added;
  }
  // remove_edge (:214-217)
  void remove_edge(int u, int v) {
    raw_remove_edge(u, v);
    update_children_flag = true;
  }
  // remove_redundant_edges (:219-238)
  void remove_redundant_edges(int u, int v) {
    std::vector<int> outs(out[u].begin(), out[u].end());
    for (int w : outs)
      if (w == v || children[v].count(w)) raw_remove_edge(u, w);
    children[u].insert(v);
    children[u].insert(children[v].begin(), children[v].end());
    std::vector<int> ins(in[u].begin(), in[u].end());
    for (int q : ins) remove_redundant_edges(q, v);
  }
  // add_relation (:240-285)
  bool add_relation(int u, int v) {
    const bool added = raw_add_edge(u, v);
    if (added) {
      if (children[u].count(v)) {  // v already reachable from u
        reduction_flag = false;
        return added;
      }
      if (children[v].count(u)) {
        cycle = true;
        return added;
      }
      if (!cycle) {
        std::vector<int> outs(out[u].begin(), out[u].end());
        for (int w : outs)
          if (children[v].count(w)) raw_remove_edge(u, w);
        children[u].insert(v);
        children[u].insert(children[v].begin(), children[v].end());
        std::vector<int> ins(in[u].begin(), in[u].end());
        for (int q : ins) remove_redundant_edges(q, v);
        topological_sort();
        transitive_reduction_dag();
      }
    }
    return added;
  }
  // remove_relation (:287-316)
  void remove_relation(int u, int v) {
    raw_remove_edge(u, v);
    update_children(u);
    std::vector<int> vouts(out[v].begin(), out[v].end());
    for (int w : vouts) {
      if (!children[u].count(w)) {
        raw_add_edge(u, w);
        children[u].insert(w);
        children[u].insert(children[w].begin(), children[w].end());
      }
    }
    std::vector<int> uins(in[u].begin(), in[u].end());
    for (int pa : uins) {
      update_children(pa);
      if (!children[pa].count(v)) {
        raw_add_edge(pa, v);
        children[pa].insert(v);
        children[pa].insert(children[v].begin(), children[v].end());
      }
    }
  }
  // swap_node (:312-316)
  void swap_node(int u, int v) {
    std::swap(event_id[u], event_id[v]);
    update_children_flag = true;
    update_node_idx_flag = true;
  }
};

// adjacency_mat2list (model_impl.cpp:343-352): poset(i,j)==1 => edge i -> j (row-major scan order)
static std::vector<std::pair<int, int>> adjacency_mat2list(const int* poset, int p) {
  std::vector<std::pair<int, int>> e;
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j)
      if (poset[(size_t)j * p + i] == 1) e.emplace_back(i, j);
  return e;
}

struct not_acyclic : std::runtime_error {
  // message of src/not_acyclic_exception.hpp:18-19
  not_acyclic() : std::runtime_error("Poset does not correspond to an acyclic graph") {}
};

static Model build_model(const int* poset, int p, float lambda_s) {
  Model M(adjacency_mat2list(poset, p), p, lambda_s);
  M.has_cycles();
  if (M.cycle) throw not_acyclic();
  M.topological_sort();
  return M;
}

// ----------------------------------------------------------------------------------------------
// compatibility predicates (src/compatible_observations.cpp)
// ----------------------------------------------------------------------------------------------
// is_compatible (:12-28): for every mutated event all direct parents must be mutated
static bool is_compatible(const unsigned char* g /* p, by event id */, const Model& M) {
  for (int v : M.topo_path) {
    if (g[M.event_id[v]])
      for (int u : M.in[v])
        if (!g[M.event_id[u]]) return false;
  }
  return true;
}
// num_incompatible_events (:51-66): sum over direct edges u->v of #{i : obs(i,u)=0 and obs(i,v)=1}
static int num_incompatible_events(const MatB& obs, const Model& M) {
  int count = 0;
  for (int s = 0; s < M.p; ++s)
    for (int t : M.out[s]) {
      int u = M.event_id[s], v = M.event_id[t];
      for (int i = 0; i < obs.rows; ++i) count += ((int)obs(i, u) - (int)obs(i, v)) < 0;
    }
  return count;
}
// num_compatible_observations (:68-76)
static int num_compatible_observations(const MatB& obs, const Model& M, int N) {
  int count = 0;
  std::vector<unsigned char> row(M.p);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < M.p; ++j) row[j] = obs(i, j);
    count += is_compatible(row.data(), M);
  }
  return count;
}

// ----------------------------------------------------------------------------------------------
// src/mcem_hcbn.cpp helpers
// ----------------------------------------------------------------------------------------------
// log_bernoulli_process, scalar overload (:84-100)
static double log_bernoulli_process(unsigned dist, double eps, unsigned p) {
  double lp = 0.0;
  if (eps == 0) {
    if (dist != 0) lp = std::log(eps + kDblEps) * dist + std::log1p(-eps - kDblEps) * (p - dist);
  } else {
    lp = std::log(eps) * dist + std::log1p(-eps) * (p - dist);
  }
  return lp;
}
// vector overload (:56-82); dist arrives as double (cast<double>() at the call sites)
static std::vector<double> log_bernoulli_process_vec(const std::vector<double>& dist, double eps, unsigned p) {
  std::vector<double> lp(dist.size(), 0.0);
  if (eps == 0) {
    for (size_t i = 0; i < dist.size(); ++i)
      if (dist[i] != 0) {
        double le = std::log(eps + kDblEps), l1 = std::log1p(-eps - kDblEps);
        lp[i] = le * dist[i] + l1 * (p - dist[i]);
      }
  } else {
    double le = std::log(eps), l1 = std::log1p(-eps);
    for (size_t i = 0; i < dist.size(); ++i) lp[i] = le * dist[i] + l1 * (p - dist[i]);
  }
  return lp;
}

// complete_log_likelihood (:103-131)
static double complete_log_likelihood(const std::vector<double>& lambda, double eps, const MatD& Tdiff,
                                      const std::vector<double>& dist, const std::vector<double>& weights,
                                      int internal) {
  const unsigned p = (unsigned)lambda.size();
  double W = 0;
  for (double w : weights) W += w;
  double sumlog = 0;
  for (double l : lambda) sumlog += std::log(l);
  double ll = 0;
  if (internal == 0) {
    double dot = 0;
    for (unsigned j = 0; j < p; ++j) {
      double cs = 0;
      for (int i = 0; i < Tdiff.rows; ++i) cs += weights[i] * Tdiff(i, j);
      dot += cs * lambda[j];
    }
    ll = W * sumlog - dot;
  } else {
    ll = W * (sumlog - p);
  }
  if (eps == 0) {
    for (size_t i = 0; i < dist.size(); ++i)
      if (dist[i] != 0)
        ll += std::log(eps + kDblEps) * weights[i] * dist[i] + std::log(1 - eps - kDblEps) * weights[i] * (p - dist[i]);
  } else {
    double a = 0, b = 0;
    for (size_t i = 0; i < dist.size(); ++i) {
      a += weights[i] * dist[i];
      b += weights[i] * (p - dist[i]);
    }
    ll += std::log(eps) * a + std::log1p(-eps) * b;
  }
  return ll;
}

// propagate T_j = Z_j + max(0, max_{u in pa(j)} T_u) in topological order (:160-172, :198-211)
static void propagate_times(const Model& M, const MatD& Z, MatD& Tsum) {
  const int N = Z.rows;
  std::vector<double> tmax(N);
  for (auto v = M.topo_path.rbegin(); v != M.topo_path.rend(); ++v) {
    std::fill(tmax.begin(), tmax.end(), 0.0);
    for (int uv : M.in[*v]) {
      const double* tu = Tsum.col(M.event_id[uv]);
      for (int n = 0; n < N; ++n)
        if (tu[n] > tmax[n]) tmax[n] = tu[n];
    }
    const int j = M.event_id[*v];
    const double* z = Z.col(j);
    double* t = Tsum.col(j);
    for (int n = 0; n < N; ++n) t[n] = z[n] + tmax[n];
  }
}

// sample_genotypes (:138-177).  Draw order: Z column by column in EVENT-ID order, then T_s.
static MatB sample_genotypes(unsigned N, const Model& M, MatD& T_events, std::vector<double>& T_sampling,
                             StdRng& rng, bool times_available) {
  const int p = M.p;
  MatD Tsum((int)N, p, 0.0);
  MatB obs((int)N, p, 0);
  for (int j = 0; j < p; ++j) rexp_fill(T_events.col(j), (int)N, M.lambda[j], rng);
  if (!times_available) rexp_fill(T_sampling.data(), (int)N, M.lambda_s, rng);
  propagate_times(M, T_events, Tsum);
  for (int j = 0; j < p; ++j)
    for (unsigned n = 0; n < N; ++n) obs(n, j) = Tsum(n, j) <= T_sampling[n];
  return obs;
}
// variant with supplied Z and T_s (no RNG): the 1e-10 / bit-exact parity tier
static MatB genotypes_from_times(const Model& M, const MatD& Z, const std::vector<double>& T_sampling, MatD* Tsum_out) {
  MatD Tsum(Z.rows, M.p, 0.0);
  MatB obs(Z.rows, M.p, 0);
  propagate_times(M, Z, Tsum);
  for (int j = 0; j < M.p; ++j)
    for (int n = 0; n < Z.rows; ++n) obs(n, j) = Tsum(n, j) <= T_sampling[n];
  if (Tsum_out) *Tsum_out = Tsum;
  return obs;
}

// sample_times (:184-213)
static MatD sample_times(unsigned N, const Model& M, MatD& T_events, StdRng& rng) {
  MatD Tsum((int)N, M.p, 0.0);
  for (int j = 0; j < M.p; ++j) rexp_fill(T_events.col(j), (int)N, M.lambda[j], rng);
  propagate_times(M, T_events, Tsum);
  return Tsum;
}

// generate_genotypes (:215-236)
static MatB generate_genotypes(const MatD& Tsum, const Model& M, std::vector<double>& T_sampling, StdRng& rng,
                               bool times_available) {
  const int N = Tsum.rows;
  MatB obs(N, M.p, 0);
  if (!times_available) rexp_fill(T_sampling.data(), N, M.lambda_s, rng);
  for (int j = 0; j < M.p; ++j)
    for (int n = 0; n < N; ++n) obs(n, j) = Tsum(n, j) <= T_sampling[n];
  return obs;
}

// hamming_dist_mat (:326-330)
static std::vector<int> hamming_dist_mat(const MatB& x, const unsigned char* y) {
  std::vector<int> d(x.rows, 0);
  for (int j = 0; j < x.cols; ++j)
    for (int n = 0; n < x.rows; ++n) d[n] += (x(n, j) != 0) != (y[j] != 0);
  return d;
}

// ----------------------------------------------------------------------------------------------
// src/backward_sampling.cpp
// ----------------------------------------------------------------------------------------------
// n_choose_k (:20-35) -- Boost binomial_coefficient<double> truncated to unsigned; exact here
static unsigned n_choose_k(unsigned n, unsigned k) {
  if (k > n) return 0;
  long double r = 1;
  unsigned kk = std::min(k, n - k);
  for (unsigned i = 1; i <= kk; ++i) r = r * (n - kk + i) / i;
  return (unsigned)std::llround(r);
}
// neighbors (:37-54): recursive block fill of all genotypes at Hamming distance exactly k.
// `m` is the full matrix, the sub-block starts at (r0, c0) and spans p columns.
static void neighbors(unsigned p, unsigned k, MatB& m, int r0, int c0) {
  if (p >= 1) {
    unsigned nrows = n_choose_k(p - 1, k);
    unsigned nrows_alt = n_choose_k(p - 1, k - 1);  // k == 0 wraps to UINT_MAX -> 0 rows, as in the reference
    if (nrows > 0) neighbors(p - 1, k, m, r0, c0 + 1);
    if (nrows_alt > 0) {
      unsigned char flipped = !m(r0 + (int)nrows, c0);
      for (unsigned r = 0; r < nrows_alt; ++r) m(r0 + (int)nrows + (int)r, c0) = flipped;
      neighbors(p - 1, k - 1, m, r0 + (int)nrows, c0 + 1);
    }
  }
}
// coin_tossing (:64-77): p*N uniforms, u[col*N + row] < eps flips bit `col` of sample `row`
static void coin_tossing(MatB& m, double eps, StdRng& rng) {
  const int p = m.cols, N = m.rows;
  std::vector<double> u((size_t)p * N);
  std::uniform_real_distribution<double> U(0, 1);
  for (auto& x : u) x = U(rng.eng);
  for (int i = 0; i < p; ++i) {
    unsigned char flipped = !m(0, i);
    for (int n = 0; n < N; ++n)
      if (u[(size_t)i * N + n] < eps) m(n, i) = flipped;
  }
}

// ----------------------------------------------------------------------------------------------
// src/add_remove.cpp
// ----------------------------------------------------------------------------------------------
// scale_path_to_mutation (:17-37)
static std::vector<double> scale_path_to_mutation(const Model& M) {
  std::vector<double> s(M.p, 0.0);
  for (auto v = M.topo_path.rbegin(); v != M.topo_path.rend(); ++v) {
    const int j = M.event_id[*v];
    s[j] = 1.0 / M.lambda[j];
    double mx = -1.0;
    for (int u : M.in[*v]) mx = std::max(mx, s[M.event_id[u]]);
    if (mx > -1.0) s[j] += mx;
  }
  return s;
}
static void add_all(std::vector<unsigned char>& g, const Model& M) {  // :39-56
  for (int v : M.topo_path)
    if (g[M.event_id[v]])
      for (int u : M.in[v]) g[M.event_id[u]] = 1;
}
static void add_parents(std::vector<unsigned char>& g, int v, const Model& M) {  // :58-72
  for (int u : M.in[v])
    if (!g[M.event_id[u]]) {
      g[M.event_id[u]] = 1;
      add_parents(g, u, M);
    }
}
static void remove_all(std::vector<unsigned char>& g, const Model& M) {  // :74-91
  for (auto v = M.topo_path.rbegin(); v != M.topo_path.rend(); ++v)
    if (!g[M.event_id[*v]])
      for (int w : M.out[*v]) g[M.event_id[w]] = 0;
}
static void remove_children(std::vector<unsigned char>& g, int v, const Model& M) {  // :93-107
  for (int w : M.out[v])
    if (g[M.event_id[w]]) {
      g[M.event_id[w]] = 0;
      remove_children(g, w, M);
    }
}
// draw_sample (:109-141)
static std::vector<unsigned char> draw_sample(const std::vector<unsigned char>& genotype, const Model& M,
                                              unsigned move, const std::vector<double>& remove_weight,
                                              const std::vector<double>& add_weight, double& q_choice,
                                              int idx_remove, int idx_add, bool compatible) {
  std::vector<unsigned char> s = genotype;
  auto sum = [](const std::vector<double>& v) { double t = 0; for (double x : v) t += x; return t; };
  switch (move) {
    case 0:
      q_choice = add_weight[idx_add] / sum(add_weight);
      s[idx_add] = 1;
      if (compatible) add_parents(s, M.node_idx[idx_add], M);
      else add_all(s, M);
      break;
    case 1:
      q_choice = remove_weight[idx_remove] / sum(remove_weight);
      s[idx_remove] = 0;
      if (compatible) remove_children(s, M.node_idx[idx_remove], M);
      else remove_all(s, M);
      break;
    case 2:
      q_choice = 1;
      break;
  }
  return s;
}

// generate_mutation_times (:156-203) with dexp_log (:144-147) / pexp_log (:150-154).
// `uniforms` == nullptr: draw from rng in the reference order (T_s first, then per node in topological
// order, N draws each).  Otherwise u(n, j) is the uniform consumed by sample n at EVENT j and T_s is
// taken from `sampling_time` (the supplied-randomness parity tier).
static MatD generate_mutation_times(const MatB& obs, const Model& M, std::vector<double>& dens,
                                    std::vector<double>& sampling_time, StdRng* rng, bool times_available,
                                    const MatD* uniforms = nullptr) {
  const int N = obs.rows, p = obs.cols;
  MatD time_events(N, p, 0.0), time_sum(N, p, 0.0);
  if (!times_available && !uniforms) rexp_fill(sampling_time.data(), N, M.lambda_s, *rng);
  std::vector<std::vector<int>> parents = M.get_direct_predecessors();
  std::vector<double> tpm(N), cutoff(N), t(N);
  for (auto v = M.topo_path.rbegin(); v != M.topo_path.rend(); ++v) {
    const int j = M.event_id[*v];
    const double rate = M.lambda[j];
    std::fill(tpm.begin(), tpm.end(), 0.0);
    if (!parents[j].empty()) {
      // rowwise().maxCoeff() over the parents' columns (NOT clamped at 0; times are >= 0 for
      // compatible samples, and negative values only occur in samples that are discarded anyway)
      for (int n = 0; n < N; ++n) {
        double m = time_sum(n, parents[j][0]);
        for (size_t q = 1; q < parents[j].size(); ++q) m = std::max(m, time_sum(n, parents[j][q]));
        tpm[n] = m;
      }
    }
    for (int n = 0; n < N; ++n) cutoff[n] = obs(n, j) ? sampling_time[n] - tpm[n] : kInf;
    if (uniforms)
      for (int n = 0; n < N; ++n) t[n] = rtexp_from_u(rate, cutoff[n], (*uniforms)(n, j));
    else
      for (int n = 0; n < N; ++n) t[n] = rtexp1(rate, cutoff[n], *rng);
    for (int n = 0; n < N; ++n) {
      double base = obs(n, j) ? tpm[n] : std::max(tpm[n], sampling_time[n]);
      time_sum(n, j) = base + t[n];
      time_events(n, j) = time_sum(n, j) - tpm[n];
      // dens += dexp_log(time, rate) - pexp_log(cutoff, rate)
      dens[n] += (std::log(rate) - rate * t[n]) - std::log(1 - std::exp(-rate * cutoff[n]));
    }
  }
  return time_events;
}

// cbn_density_log (:206-216)
static std::vector<double> cbn_density_log(const MatD& time, const std::vector<double>& lambda) {
  double s = 0;
  for (double l : lambda) s += std::log(l);
  std::vector<double> r(time.rows, s);
  for (int n = 0; n < time.rows; ++n) {
    double dot = 0;
    for (int j = 0; j < time.cols; ++j) dot += time(n, j) * lambda[j];
    r[n] -= dot;
  }
  return r;
}

// ----------------------------------------------------------------------------------------------
// importance_weight (src/mcem_hcbn.cpp:343-566)
// ----------------------------------------------------------------------------------------------
struct DataIS {  // DataImportanceSampling (src/mcem.hpp:200-210)
  std::vector<double> w;
  std::vector<int> dist;
  MatD Tdiff;
  DataIS(unsigned L, unsigned p) : w(L), dist(L), Tdiff((int)L, (int)p) {}
};

static MatB replicate_rows(const unsigned char* g, int p, int n) {
  MatB m(n, p);
  for (int j = 0; j < p; ++j)
    for (int r = 0; r < n; ++r) m(r, j) = g[j];
  return m;
}

// Hamming-ball enumeration used by the backward scheme (:497-509): row 0 = Y, then rings 1..k
static MatB backward_ball(const unsigned char* genotype, unsigned p, unsigned k, std::vector<int>& ring) {
  std::vector<unsigned> nrows(k + 1);  // the reference allocates only k entries (OOB write, :354)
  unsigned total = 0;
  for (unsigned i = 0; i <= k; ++i) total += (nrows[i] = n_choose_k(p, i));
  MatB samples = replicate_rows(genotype, (int)p, (int)total);
  ring.assign(total, 0);
  unsigned at = 1;
  for (unsigned i = 1; i <= k; ++i) {
    neighbors(p, i, samples, (int)at, 0);
    for (unsigned r = 0; r < nrows[i]; ++r) ring[at + r] = (int)i;
    at += nrows[i];
  }
  return samples;
}

static DataIS importance_weight(const unsigned char* genotype, unsigned L, const Model& M, double time,
                                const std::string& sampling, const std::vector<double>& scale_cumulative,
                                const std::vector<int>& dist_pool, const MatD& Tdiff_pool, unsigned neighborhood_dist,
                                StdRng& rng, bool times_available) {
  const unsigned p = (unsigned)M.p;
  unsigned reps = 0, L_aux = 1, nrows_sum = 0;
  if (sampling == "backward") {
    for (unsigned i = 0; i <= neighborhood_dist; ++i) nrows_sum += n_choose_k(p, i);
    reps = L / nrows_sum;
    L = reps * nrows_sum;
  } else if (sampling == "bernoulli") {
    reps = 1;
    L_aux = std::max(L / reps, L_aux);
    L = reps * L_aux;
  }
  DataIS is(L, p);
  const double eps = M.epsilon;

  if (sampling == "forward") {
    std::vector<double> T_sampling(L);
    if (times_available) std::fill(T_sampling.begin(), T_sampling.end(), time);
    MatB samples = sample_genotypes(L, M, is.Tdiff, T_sampling, rng, times_available);
    is.dist = hamming_dist_mat(samples, genotype);
    for (unsigned l = 0; l < L; ++l) {
      double d = (double)is.dist[l];
      is.w[l] = std::pow(eps, d) * std::pow(1 - eps, p - d);  // :386-388
    }
  } else if (sampling == "add-remove") {
    std::vector<unsigned char> g(genotype, genotype + p);
    bool compatible = is_compatible(genotype, M);
    unsigned mutations = 0;
    for (unsigned j = 0; j < p; ++j) mutations += g[j] != 0;
    bool wild_type = (mutations == 0), resistant_type = (mutations == p);
    std::vector<int> move(L);
    std::vector<double> log_proposal(L);
    if (compatible) {
      if (wild_type) {
        move = rdiscrete_std(L, {0.5, 0.0, 0.5}, rng);
        std::fill(log_proposal.begin(), log_proposal.end(), std::log(0.5));
      } else if (resistant_type) {
        move = rdiscrete_std(L, {0.0, 0.5, 0.5}, rng);
        std::fill(log_proposal.begin(), log_proposal.end(), std::log(0.5));
      } else {
        move = rdiscrete_std(L, {1.0, 1.0, 1.0}, rng);
        std::fill(log_proposal.begin(), log_proposal.end(), -std::log(3));
      }
    } else {
      move = rdiscrete_std(L, {0.5, 0.5}, rng);
      std::fill(log_proposal.begin(), log_proposal.end(), std::log(0.5));
    }
    std::vector<double> remove_weight(p), add_weight(p);
    for (unsigned j = 0; j < p; ++j) {
      remove_weight[j] = g[j] ? scale_cumulative[j] : 0.0;
      add_weight[j] = g[j] ? 0.0 : 1.0 / scale_cumulative[j];
    }
    std::vector<int> idx_remove = rdiscrete_std(L, remove_weight, rng);
    std::vector<int> idx_add = rdiscrete_std(L, add_weight, rng);
    MatB samples((int)L, (int)p);
    double q_choice = 0;
    for (unsigned l = 0; l < L; ++l) {
      std::vector<unsigned char> s =
          draw_sample(g, M, (unsigned)move[l], remove_weight, add_weight, q_choice, idx_remove[l], idx_add[l], compatible);
      for (unsigned j = 0; j < p; ++j) samples((int)l, (int)j) = s[j];
      log_proposal[l] += std::log(q_choice);
    }
    is.dist = hamming_dist_mat(samples, genotype);
    std::vector<double> T_sampling(L);
    if (times_available) std::fill(T_sampling.begin(), T_sampling.end(), time);
    is.Tdiff = generate_mutation_times(samples, M, log_proposal, T_sampling, &rng, times_available);
    std::vector<double> dd(is.dist.begin(), is.dist.end());
    std::vector<double> lpyx = log_bernoulli_process_vec(dd, eps, p);
    std::vector<double> lpx = cbn_density_log(is.Tdiff, M.lambda);
    for (unsigned l = 0; l < L; ++l) is.w[l] = std::exp(lpyx[l] + lpx[l] - log_proposal[l]);
  } else if (sampling == "pool") {
    const size_t K = dist_pool.size();
    std::vector<double> q(K);
    double qs = 0;
    for (size_t k = 0; k < K; ++k) {
      double d = (double)dist_pool[k];
      q[k] = std::pow(eps, d) * std::pow(1 - eps, p - d);
      qs += q[k];
    }
    bool random = false;
    if (qs == 0) {
      std::fill(q.begin(), q.end(), 1.0);
      random = true;
    }
    double q_sum = 0;
    for (double x : q) q_sum += x;
    for (double& x : q) x /= q_sum;
    std::vector<int> idxs = rdiscrete_std(L, q, rng);
    for (unsigned l = 0; l < L; ++l) {
      int idx = idxs[l];
      is.dist[l] = dist_pool[idx];
      for (unsigned j = 0; j < p; ++j) is.Tdiff((int)l, (int)j) = Tdiff_pool(idx, (int)j);
    }
    if (random) {
      std::vector<double> dd(is.dist.begin(), is.dist.end());
      std::vector<double> lp = log_bernoulli_process_vec(dd, eps, p);
      for (unsigned l = 0; l < L; ++l) is.w[l] = std::exp(lp[l]);
    } else {
      std::fill(is.w.begin(), is.w.end(), q_sum / (double)K);
    }
  } else if (sampling == "backward") {
    std::vector<int> ring;
    MatB samples = backward_ball(genotype, p, neighborhood_dist, ring);
    std::vector<double> lpyx_aux(nrows_sum);
    for (unsigned r = 0; r < nrows_sum; ++r) lpyx_aux[r] = log_bernoulli_process((unsigned)ring[r], eps, p);
    MatB samples_rep((int)L, (int)p);
    std::vector<double> lpyx(L), log_proposal(L, 0.0);
    for (unsigned l = 0; l < L; ++l) {  // replicate(reps, 1): row l = base row (l mod nrows_sum)
      unsigned r = l % nrows_sum;
      for (unsigned j = 0; j < p; ++j) samples_rep((int)l, (int)j) = samples((int)r, (int)j);
      is.dist[l] = ring[r];
      lpyx[l] = lpyx_aux[r];
    }
    std::vector<double> T_sampling(L);
    if (times_available) std::fill(T_sampling.begin(), T_sampling.end(), time);
    is.Tdiff = generate_mutation_times(samples_rep, M, log_proposal, T_sampling, &rng, times_available);
    std::vector<unsigned char> incompatible(L, 0);
    long n_incompatible = 0;
    for (unsigned l = 0; l < L; ++l) {
      for (unsigned j = 0; j < p; ++j)
        if (is.Tdiff((int)l, (int)j) < 0.0) incompatible[l] = 1;
      n_incompatible += incompatible[l];
    }
    // integer division count / reps (:525)
    double corr = reps ? std::log((double)((long)nrows_sum - n_incompatible / (long)reps)) : 0.0;
    std::vector<double> lpx = cbn_density_log(is.Tdiff, M.lambda);
    for (unsigned l = 0; l < L; ++l) {
      double w = std::exp(lpyx[l] + lpx[l] - (log_proposal[l] - corr));
      is.w[l] = incompatible[l] ? 0.0 : w;
    }
  } else if (sampling == "bernoulli") {
    MatB samples = replicate_rows(genotype, (int)p, (int)L_aux);
    coin_tossing(samples, eps, rng);
    std::vector<int> dist = hamming_dist_mat(samples, genotype);
    for (unsigned l = 0; l < L; ++l) is.dist[l] = dist[l % L_aux];
    std::vector<double> log_proposal(L, 0.0), T_sampling(L);
    if (times_available) std::fill(T_sampling.begin(), T_sampling.end(), time);
    is.Tdiff = generate_mutation_times(samples, M, log_proposal, T_sampling, &rng, times_available);
    std::vector<double> lpx = cbn_density_log(is.Tdiff, M.lambda);
    for (unsigned l = 0; l < L; ++l) {
      bool inc = false;
      for (unsigned j = 0; j < p; ++j) inc |= is.Tdiff((int)l, (int)j) < 0.0;
      is.w[l] = inc ? 0.0 : std::exp(lpx[l] - log_proposal[l]);
    }
  } else {
    throw std::runtime_error("unknown sampling scheme: " + sampling);
  }
  return is;
}

static const char* kZeroWeightMsg = "ERROR: all samples have weight 0. Consider increasing L";

struct ControlEM {  // src/mcem.hpp:213-226
  unsigned max_iter = 100, update_step_size = 20;
  double tol = 0.001;
  float max_lambda = 1e6f;
  unsigned neighborhood_dist = 1;
};

struct ObsIS {  // per-observation reductions of one importance_weight call (:683-694, :991-996)
  double aux, wsq, wd;
  long npos;
  std::vector<double> wz;
};
static ObsIS reduce_is(const DataIS& is, unsigned p) {
  ObsIS r;
  r.aux = r.wsq = r.wd = 0;
  r.npos = 0;
  r.wz.assign(p, 0.0);
  const size_t L = is.w.size();
  for (size_t l = 0; l < L; ++l) {
    r.aux += is.w[l];
    r.wsq += is.w[l] * is.w[l];
    r.wd += is.w[l] * (double)is.dist[l];
    r.npos += is.w[l] > 0;
  }
  for (unsigned j = 0; j < p; ++j) {
    double s = 0;
    for (size_t l = 0; l < L; ++l) s += is.Tdiff((int)l, (int)j) * is.w[l];
    r.wz[j] = s;
  }
  return r;
}

// per-observation pool genotypes + distances (:668-677)
static std::vector<int> pool_distances(const MatD& T_pool, const Model& M, double time_i, bool times_available,
                                       const unsigned char* y, StdRng& rng) {
  std::vector<double> T_sampling(T_pool.rows);
  if (times_available) std::fill(T_sampling.begin(), T_sampling.end(), time_i);
  MatB pool = generate_genotypes(T_pool, M, T_sampling, rng, times_available);
  return hamming_dist_mat(pool, y);
}

// MCEM_hcbn (src/mcem_hcbn.cpp:571-736)
static double MCEM_hcbn(Model& model, const MatB& obs, const std::vector<double>& times,
                        const std::vector<double>& weights, unsigned L, const std::string& sampling,
                        const ControlEM& ctl, bool times_available, unsigned thrds, Context& ctx,
                        std::vector<double>* trace = nullptr) {
  const unsigned p = (unsigned)model.p, N = (unsigned)obs.rows;
  double N_eff = 0.0;
  unsigned K = 0, update_step_size = ctl.update_step_size;
  std::vector<double> avg_lambda(p, 0.0), avg_lambda_current(p, 0.0);
  double avg_eps = 0.0, avg_llhood = 0.0, obs_llhood = 0.0, avg_eps_current = 0.0;
  bool tol_comparison = true;
  double expected_dist = 0.0;
  MatD expected_Tdiff((int)N, (int)p, 0.0);
  MatD Tdiff_pool;
  std::vector<double> scale_cumulative;

  if (sampling == "add-remove") {
    scale_cumulative.resize(p);
    if (model.update_node_idx_flag) model.update_node_idx();
  } else if (sampling == "pool") {
    K = p * L;
    Tdiff_pool = MatD((int)K, (int)p);
  }
  if (ctx.verbose) {
    std::printf("Initial value of the error rate - epsilon: %g\n", model.epsilon);
    std::printf("Initial value of the rate parameters - lambda:");
    for (double l : model.lambda) std::printf(" %g", l);
    std::printf("\n");
  }

  for (unsigned iter = 0; iter < ctl.max_iter; ++iter) {
    MatD T_pool;
    if (iter == update_step_size) {  // batch average + convergence test (:622-642)
      for (auto& x : avg_lambda_current) x /= ctl.update_step_size;
      avg_eps_current /= ctl.update_step_size;
      avg_llhood /= ctl.update_step_size;
      if (tol_comparison) {
        bool all_ok = std::abs(avg_eps - avg_eps_current) <= ctl.tol;
        for (unsigned j = 0; j < p && all_ok; ++j) all_ok = std::abs(avg_lambda[j] - avg_lambda_current[j]) <= ctl.tol;
        if (all_ok) break;
      }
      avg_lambda = avg_lambda_current;
      avg_eps = avg_eps_current;
      update_step_size += ctl.update_step_size;
      tol_comparison = true;
      std::fill(avg_lambda_current.begin(), avg_lambda_current.end(), 0.0);
      avg_eps_current = 0.0;
      avg_llhood = 0.0;
    }
    if (sampling == "add-remove") {
      scale_cumulative = scale_path_to_mutation(model);
    } else if (sampling == "pool") {
      T_pool = sample_times(K, model, Tdiff_pool, ctx.rng);  // shared pool, master rng (:650-654)
    }
    N_eff = 0;
    obs_llhood = 0.0;
    expected_dist = 0.0;
#ifdef _OPENMP
    omp_set_num_threads((int)thrds);
#endif
    auto rngs = ctx.get_auxiliary_rngs((int)thrds);
    bool zero_weight = false;

#pragma omp parallel for reduction(+ : obs_llhood) reduction(+ : expected_dist) reduction(+ : N_eff) schedule(static)
    for (unsigned i = 0; i < N; ++i) {
#ifdef _OPENMP
      StdRng& rng_t = *rngs[omp_get_thread_num()];
#else
      StdRng& rng_t = *rngs[0];
#endif
      std::vector<unsigned char> y(p);
      for (unsigned j = 0; j < p; ++j) y[j] = obs((int)i, (int)j);
      std::vector<int> d_pool;
      if (sampling == "pool") d_pool = pool_distances(T_pool, model, times[i], times_available, y.data(), rng_t);
      DataIS is = importance_weight(y.data(), L, model, times[i], sampling, scale_cumulative, d_pool, Tdiff_pool,
                                    ctl.neighborhood_dist, rng_t, times_available);
      ObsIS r = reduce_is(is, p);
      if (r.aux > 0) {
        N_eff += weights[i];
        int L_eff = (int)L;
        if (sampling == "backward") L_eff = (int)r.npos;
        obs_llhood += weights[i] * std::log(r.aux / L_eff);
        expected_dist += weights[i] * r.wd / r.aux;
        for (unsigned j = 0; j < p; ++j) expected_Tdiff((int)i, (int)j) = r.wz[j] / r.aux;
      } else {
#pragma omp critical
        zero_weight = true;  // the reference throws inside the parallel region (:696-698)
      }
    }
    if (zero_weight) throw std::runtime_error(kZeroWeightMsg);

    // M-step (:707-709)
    model.update_epsilon(expected_dist / (N_eff * p), kDblEps);
    std::vector<double> newl(p);
    for (unsigned j = 0; j < p; ++j) {
      double cs = 0;
      for (unsigned i = 0; i < N; ++i) cs += weights[i] * expected_Tdiff((int)i, (int)j);
      newl[j] = 1.0 / (cs / N_eff);
    }
    model.update_lambda(newl, ctl.max_lambda);

    for (unsigned j = 0; j < p; ++j) avg_lambda_current[j] += model.lambda[j];
    avg_eps_current += model.epsilon;
    avg_llhood += obs_llhood;
    if (trace) {
      trace->push_back(obs_llhood);
      trace->push_back(model.epsilon);
      for (double l : model.lambda) trace->push_back(l);
    }
    if (iter + 1 == ctl.max_iter) {  // last batch (:715-721)
      unsigned num_iter = ctl.max_iter - update_step_size + ctl.update_step_size;
      for (auto& x : avg_lambda_current) x /= num_iter;
      avg_eps_current /= num_iter;
      avg_llhood /= num_iter;
    }
    if (ctx.verbose) {
      if (iter == 0) std::printf("llhood\tepsilon\tlambdas\n");
      std::printf("%g\t%g\t", obs_llhood, model.epsilon);
      for (double l : model.lambda) std::printf(" %g", l);
      std::printf("\n");
    }
  }
  model.lambda = avg_lambda_current;
  model.epsilon = avg_eps_current;
  model.llhood = avg_llhood;
  return avg_llhood;
}

// obs_log_likelihood (src/mcem_hcbn.cpp:239-312)
static double obs_log_likelihood(const MatB& obs, const int* poset, int p, const std::vector<double>& lambda, double eps,
                                 const std::vector<double>& weights, const std::vector<double>& times, unsigned L,
                                 const std::string& sampling, unsigned neighborhood_dist, Context& ctx, float lambda_s,
                                 bool times_available, unsigned thrds) {
  const unsigned N = (unsigned)obs.rows;
  double llhood = 0.0;
  Model model = build_model(poset, p, lambda_s);
  model.lambda = lambda;
  model.epsilon = eps;
  unsigned K = 0;
  std::vector<double> scale_cumulative;
  MatD Tdiff_pool, T_pool;
  if (sampling == "add-remove") {
    scale_cumulative = scale_path_to_mutation(model);
    if (model.update_node_idx_flag) model.update_node_idx();
  } else if (sampling == "pool") {
    K = (unsigned)p * L;
    Tdiff_pool = MatD((int)K, p);
    T_pool = sample_times(K, model, Tdiff_pool, ctx.rng);
  }
#ifdef _OPENMP
  omp_set_num_threads((int)thrds);
#endif
  auto rngs = ctx.get_auxiliary_rngs((int)thrds);
  bool zero_weight = false;
#pragma omp parallel for reduction(+ : llhood) schedule(static)
  for (unsigned i = 0; i < N; ++i) {
#ifdef _OPENMP
    StdRng& rng_t = *rngs[omp_get_thread_num()];
#else
    StdRng& rng_t = *rngs[0];
#endif
    std::vector<unsigned char> y(p);
    for (int j = 0; j < p; ++j) y[j] = obs((int)i, j);
    std::vector<int> d_pool;
    if (sampling == "pool") d_pool = pool_distances(T_pool, model, times[i], times_available, y.data(), rng_t);
    DataIS is = importance_weight(y.data(), L, model, times[i], sampling, scale_cumulative, d_pool, Tdiff_pool,
                                  neighborhood_dist, rng_t, times_available);
    double aux = 0;
    long npos = 0;
    for (double w : is.w) {
      aux += w;
      npos += w > 0;
    }
    if (aux > 0) {
      int L_eff = (int)L;
      if (sampling == "backward") L_eff = (int)npos;
      llhood += weights[i] * std::log(aux / L_eff);
    } else {
#pragma omp critical
      zero_weight = true;
    }
  }
  if (zero_weight) throw std::runtime_error(kZeroWeightMsg);
  return llhood;
}

// initialize_lambda (src/asa.cpp:72-129): count-based start values, float division
static void initialize_lambda(Model& model, const MatB& obs, float max_lambda) {
  const unsigned N = (unsigned)obs.rows;
  const int p = model.p;
  std::vector<double> lambda(p);
  std::vector<std::set<int>> all_pred(p);  // by event id, transitive
  for (auto v = model.topo_path.rbegin(); v != model.topo_path.rend(); ++v) {
    const int j = model.event_id[*v];
    for (int uv : model.in[*v]) {
      int u = model.event_id[uv];
      all_pred[j].insert(u);
      all_pred[j].insert(all_pred[u].begin(), all_pred[u].end());
    }
    unsigned count = 1;
    float mutated = 1e-7f;
    int np = (int)all_pred[j].size();
    if (np > 0) {
      unsigned c1 = 0, c2 = 0;
      for (unsigned i = 0; i < N; ++i) {
        int s = 0;
        for (int u : all_pred[j]) s += obs((int)i, u);
        c1 += (s == np);
        c2 += (s + obs((int)i, j) == np + 1);
      }
      count += c1;
      mutated += c2;  // float += int
    } else {
      count += N;
      int s = 0;
      for (unsigned i = 0; i < N; ++i) s += obs((int)i, j);
      mutated += s;
    }
    float aux = mutated / count;
    lambda[j] = aux * model.lambda_s / (1.0 - aux);
  }
  model.update_lambda(lambda, max_lambda);
}


// ----------------------------------------------------------------------------------------------
// adaptive simulated annealing (src/asa.cpp:131-594).  ONE generator (ctx.rng) drives the move proposals, the
// compatibility heuristic, the Metropolis rule and -- through get_auxiliary_rngs -- every MCEM score, exactly as in the
// reference.  `score_mode` 1 replaces the MCEM score by a deterministic function of the proposed poset (test seam for
// the move generator / bookkeeping; not part of the reference).
// ----------------------------------------------------------------------------------------------
struct ControlSA {  // src/mcem.hpp:228-260
  std::string outdir;
  float acceptance_rate = 1.0f / 3, T = 50.0f, adap_rate = 0.3f;
  int step_size = 1;
  unsigned max_iter = 10000;
  bool adaptive = true;
  float compatible_fraction_factor = 0.05f;
};

static void write_poset(const Model& M, const std::string& outdir) {  // :59-70 (boost::edges order: by source vertex)
  std::ofstream f(outdir + "poset.txt", std::ofstream::trunc);
  for (int u = 0; u < M.p; ++u)
    for (int v : M.out[u]) f << M.event_id[u] << "\t" << M.event_id[v] << std::endl;
}

static bool heuristic_compatibility(Model& M, const MatB& obs, double fraction_compatible, double& fraction_new,
                                    float factor, Context& ctx) {  // :133-162
  std::uniform_real_distribution<> rand01(0.0, 1.0);
  const int N = obs.rows;
  fraction_new = (double)num_compatible_observations(obs, M, N) / N;
  if (fraction_new < fraction_compatible) {
    double prob = std::exp((fraction_new - fraction_compatible) / factor);
    if (prob > rand01(ctx.rng)) return true;
  } else {
    return true;
  }
  return false;
}

struct AsaStats {  // what the tests compare
  int num_accept_total = 0, moves[4] = {0, 0, 0, 0}, n_scored = 0;
};

static double asa_score(Model& M_new, const MatB& obs, const std::vector<double>& times,
                        const std::vector<double>& weights, unsigned L, const std::string& sampling,
                        const ControlEM& ctl, bool times_available, unsigned thrds, Context& ctx, int score_mode,
                        AsaStats& st) {
  const int N = obs.rows, p = M_new.p;
  initialize_lambda(M_new, obs, ctl.max_lambda);
  M_new.update_epsilon((double)num_incompatible_events(obs, M_new) / ((double)N * p),
                       std::numeric_limits<double>::epsilon());
  st.n_scored++;
  if (score_mode == 1)  // deterministic pseudo score (negative, never exactly 0.0 = "not computed")
    return -1000.0 - (double)num_incompatible_events(obs, M_new) + 3.0 * (double)M_new.num_edges();
  return MCEM_hcbn(M_new, obs, times, weights, L, sampling, ctl, times_available, thrds, ctx);
}

// propose_edge (:167-470)
static double propose_edge(Model& M, const MatB& obs, double llhood_current, double& fraction_compatible,
                           const std::vector<double>& times, const std::vector<double>& weights, float T,
                           int& num_accept, float factor, unsigned L, const std::string& sampling,
                           const ControlEM& ctl, bool times_available, MatD& ll_addRemove, MatD& ll_swap,
                           MatD& ll_preserve, unsigned thrds, Context& ctx, int score_mode, AsaStats& st) {
  const unsigned p = (unsigned)M.p;
  Model M_new(M);
  double llhood_new = 0.0, fraction_new = 0.0;
  const unsigned num_edges = (unsigned)M.num_edges();
  std::vector<int> idxs_pool(num_edges), cover_pool(p * p);
  std::iota(idxs_pool.begin(), idxs_pool.end(), 0);
  std::shuffle(idxs_pool.begin(), idxs_pool.end(), ctx.rng);
  std::iota(cover_pool.begin(), cover_pool.end(), 0);
  std::shuffle(cover_pool.begin(), cover_pool.end(), ctx.rng);
  std::uniform_real_distribution<> rand01(0.0, 1.0);
  std::discrete_distribution<int> distribution({0.5, 0.0, 0.4, 0.1});
  const int move = distribution(ctx.rng);
  st.moves[move]++;
  unsigned i = 0, v1 = 0, v2 = 0;
  auto score = [&]() {
    return asa_score(M_new, obs, times, weights, L, sampling, ctl, times_available, thrds, ctx, score_mode, st);
  };
  switch (move) {
    case 0:
      for (i = 0; i < p * p; ++i) {
        M_new = M;
        v1 = (unsigned)cover_pool[i] / p;
        v2 = (unsigned)cover_pool[i] % p;
        if (v1 == v2) continue;
        bool add = M_new.add_edge((int)v1, (int)v2);
        if (add) {
          if (M_new.cycle) continue;
          if (!M_new.reduction_flag) continue;
        } else {
          M_new.remove_edge((int)v1, (int)v2);
        }
        if (heuristic_compatibility(M_new, obs, fraction_compatible, fraction_new, factor, ctx)) break;
      }
      if (ll_addRemove((int)v1, (int)v2) == 0.0) {
        llhood_new = score();
        ll_addRemove((int)v1, (int)v2) = llhood_new;
      } else {
        llhood_new = ll_addRemove((int)v1, (int)v2);
      }
      break;
    case 1:
      for (i = 0; i < num_edges; ++i) {
        M_new = M;
        int cnt = 0, a = -1, b = -1;
        for (int u = 0; u < (int)p && a < 0; ++u)
          for (int w : M_new.out[u]) {
            if (cnt == idxs_pool[i]) {
              a = u;
              b = w;
              break;
            }
            ++cnt;
          }
        M_new.remove_edge(a, b);
        M_new.add_edge(b, a);
        if (M_new.cycle) continue;
        if (heuristic_compatibility(M_new, obs, fraction_compatible, fraction_new, factor, ctx)) break;
      }
      llhood_new = score();
      break;
    case 2:
      for (i = 0; i < p * p; ++i) {
        v1 = (unsigned)cover_pool[i] / p;
        v2 = (unsigned)cover_pool[i] % p;
        if (v1 == v2) continue;
        if (M.update_children_flag) M.set_children();
        M_new = M;
        bool add = M_new.add_relation((int)v1, (int)v2);
        if (add) {
          if (M_new.cycle) continue;
          if (!M_new.reduction_flag) continue;
        } else {
          M_new.remove_relation((int)v1, (int)v2);
        }
        if (heuristic_compatibility(M_new, obs, fraction_compatible, fraction_new, factor, ctx)) break;
      }
      if (ll_preserve((int)v1, (int)v2) == 0.0) {
        llhood_new = score();
        ll_preserve((int)v1, (int)v2) = llhood_new;
      } else {
        llhood_new = ll_preserve((int)v1, (int)v2);
      }
      break;
    case 3:
      for (i = 0; i < p * p; ++i) {
        M_new = M;
        v1 = (unsigned)cover_pool[i] / p;
        v2 = (unsigned)cover_pool[i] % p;
        if (v1 == v2) continue;
        M_new.swap_node((int)v1, (int)v2);
        if (heuristic_compatibility(M_new, obs, fraction_compatible, fraction_new, factor, ctx)) break;
      }
      if (ll_swap((int)v1, (int)v2) == 0.0) {
        llhood_new = score();
        ll_swap((int)v1, (int)v2) = ll_swap((int)v2, (int)v1) = llhood_new;
      } else {
        llhood_new = ll_swap((int)v1, (int)v2);
      }
      break;
  }
  auto accept = [&]() {
    num_accept += 1;
    st.num_accept_total += 1;
    M = M_new;
    fraction_compatible = fraction_new;
    std::fill(ll_addRemove.a.begin(), ll_addRemove.a.end(), 0.0);
    std::fill(ll_swap.a.begin(), ll_swap.a.end(), 0.0);
    std::fill(ll_preserve.a.begin(), ll_preserve.a.end(), 0.0);
    return llhood_new;
  };
  if (llhood_new > llhood_current) return accept();
  if (T == 0.0) return llhood_current;
  const double acceptance_prob = std::exp(-(llhood_current - llhood_new) / T);
  if (acceptance_prob > rand01(ctx.rng)) return accept();
  return llhood_current;
}

// simulated_annealing (:475-594)
static double simulated_annealing(Model& poset, const MatB& obs, const std::vector<double>& times,
                                  const std::vector<double>& weights, ControlSA& ctl_asa, unsigned L,
                                  const std::string& sampling, const ControlEM& ctl_em, bool times_available,
                                  unsigned thrds, Context& ctx, int score_mode, AsaStats& st) {
  const int N = obs.rows, p = poset.p;
  const float scaling_const = -std::log(2.0) / std::log(ctl_asa.acceptance_rate);
  int num_accept = 0;
  MatD ll_addRemove(p, p, 0.0), ll_swap(p, p, 0.0), ll_preserve(p, p, 0.0);
  auto full_score = [&](Model& m, const ControlEM& c) {
    if (score_mode == 1) return -1000.0 - (double)num_incompatible_events(obs, m) + 3.0 * (double)m.num_edges();
    return MCEM_hcbn(m, obs, times, weights, L, sampling, c, times_available, thrds, ctx);
  };
  double llhood = full_score(poset, ctl_em);
  Model poset_ML(poset);
  double llhood_ML = llhood;
  double fraction_compatible = (double)num_compatible_observations(obs, poset, N) / N;
  std::ofstream f_temp(ctl_asa.outdir + "asa.txt"), f_par(ctl_asa.outdir + "params.txt");
  f_temp << "step\t llhood\t temperature\t acceptance rate" << std::endl;
  f_par << "step\t llhood\t epsilon\t lambdas" << std::endl;
  auto put_lambda = [&](std::ostream& os, const Model& m) {
    for (int j = 0; j < p; ++j) os << (j ? " " : "") << m.lambda[j];  // Eigen's default row-vector format
  };
  f_par << 0 << "\t" << llhood << "\t" << poset.epsilon << "\t";
  put_lambda(f_par, poset);
  f_par << std::endl;
  for (unsigned iter = 1; iter < ctl_asa.max_iter; ++iter) {
    llhood = propose_edge(poset, obs, llhood, fraction_compatible, times, weights, ctl_asa.T, num_accept,
                          ctl_asa.compatible_fraction_factor, L, sampling, ctl_em, times_available, ll_addRemove,
                          ll_swap, ll_preserve, thrds, ctx, score_mode, st);
    if (llhood > llhood_ML) {
      llhood_ML = llhood;
      poset_ML = poset;
      write_poset(poset, ctl_asa.outdir);
    }
    f_par << iter << "\t" << llhood << "\t" << poset.epsilon << "\t";
    put_lambda(f_par, poset);
    f_par << std::endl;
    if (!ctl_asa.adaptive) {
      const float rate = (float)num_accept / (iter + 1);
      ctl_asa.T *= 1 - 1.0 / p;
      f_temp << iter << "\t" << llhood << "\t" << ctl_asa.T << "\t" << rate << std::endl;
    } else if (!(iter % ctl_asa.step_size)) {
      const float rate = (float)num_accept / ctl_asa.step_size;
      ctl_asa.T *= std::exp((0.5 - std::pow(rate, scaling_const)) * ctl_asa.adap_rate);
      num_accept = 0;
      f_temp << iter << "\t" << llhood << "\t" << ctl_asa.T << "\t" << rate << std::endl;
    }
  }
  ControlEM last = ctl_em;
  last.max_iter = ctl_em.max_iter * 2;
  last.update_step_size = ctl_em.update_step_size * 2;
  last.neighborhood_dist = 1;  // ControlEM's default: the 4-argument constructor is used at :583-585
  poset = poset_ML;
  llhood = full_score(poset, last);
  return llhood;
}

// ----------------------------------------------------------------------------------------------
// legacy OT-CBN MCEM (src/mcem.cpp:30-210).  The reference draws from R's global RNG (runif / rexp sugar);
// here the same formulas are driven by StdRng -- the stream differs, the distribution does not.
// ----------------------------------------------------------------------------------------------
static double ot_runif(StdRng& rng) {
  std::uniform_real_distribution<double> U(0.0, 1.0);
  return U(rng.eng);
}
static double ot_rtexp(double m, double t, StdRng& rng) {  // :30-33
  double u = ot_runif(rng);
  return -std::log(1 - u * (1 - std::exp(-t * m))) / m;
}
static double ot_rexp(double rate, StdRng& rng) { return -std::log1p(-ot_runif(rng)) / rate; }
static double ot_dexp_log(double x, double rate) { return std::log(rate) - rate * x; }
static double ot_pexp_log(double x, double rate) { return std::log1p(-std::exp(-rate * x)); }

// drawHiddenVarsSample (:52-92)
static std::vector<double> drawHiddenVarsSample(const std::vector<int>& O_vec, double time,
                                                const std::vector<double>& lambda, const std::vector<int>& topo_path,
                                                const std::vector<std::vector<int>>& parents, double& dens,
                                                float lambda_s, bool time_available, StdRng& rng) {
  const int p = (int)lambda.size();
  std::vector<double> T(p, 0.0), T_sum(p, 0.0);
  dens = 0.0;
  if (!time_available) time = ot_rexp(lambda_s, rng);
  for (int k = 0; k < p; ++k) {
    const int e = topo_path[k];
    double max_parents_t = 0.0;
    for (int u : parents[e])
      if (T_sum[u] > max_parents_t) max_parents_t = T_sum[u];
    if (O_vec[e] == 1) {
      double tmp = ot_rtexp(lambda[e], time - max_parents_t, rng);
      dens += ot_dexp_log(tmp, lambda[e]) - ot_pexp_log(time - max_parents_t, lambda[e]);
      T_sum[e] = max_parents_t + tmp;
    } else {
      double tmp = ot_rexp(lambda[e], rng);
      T_sum[e] = std::max(time, max_parents_t) + tmp;
      dens += ot_dexp_log(tmp, lambda[e]);
    }
    T[e] = T_sum[e] - max_parents_t;
  }
  return T;
}

// _expectedTimeDifference (:105-146)
static std::vector<double> expectedTimeDifference(const std::vector<int>& O_vec, double time,
                                                  const std::vector<double>& lambda, const std::vector<int>& topo_path,
                                                  const std::vector<std::vector<int>>& parents, int nrOfSamples,
                                                  float lambda_s, bool time_available, StdRng& rng) {
  const int p = (int)lambda.size();
  std::vector<std::vector<double>> samples(nrOfSamples);
  std::vector<double> w(nrOfSamples);
  double wsum = 0;
  for (int i = 0; i < nrOfSamples; ++i) {
    double proposal_density;
    samples[i] = drawHiddenVarsSample(O_vec, time, lambda, topo_path, parents, proposal_density, lambda_s,
                                      time_available, rng);
    double org = 0.0;
    for (int j = 0; j < p; ++j) org += ot_dexp_log(samples[i][j], lambda[j]);  // log_cbn_density_ (:94-102)
    w[i] = std::exp(org - proposal_density);
    wsum += w[i];
  }
  if (wsum == 0)
    for (auto& x : w) x = 1.0 / nrOfSamples;
  else
    for (auto& x : w) x /= wsum;
  std::vector<double> td(p, 0.0);
  for (int j = 0; j < p; ++j)
    for (int i = 0; i < nrOfSamples; ++i) td[j] += samples[i][j] * w[i];
  return td;
}

}  // namespace ref

extern "C" int ref_MCEM_otcbn(const double* ilambda, const int* poset, const int* O, const double* times, int max_iter,
                              float alpha, const int* topo_path, const double* weights, int nrOfSamples,
                              float max_lambda, float lambda_s, int time_available, int seed, int N, int p,
                              double* lambda_out, double* llhood_out) {
  using namespace ref;
  try {
    // _MCEM (:149-210)
    StdRng rng((unsigned)seed);
    std::vector<std::vector<int>> parents(p);
    for (int k = 0; k < p; ++k)
      for (int i = 0; i < p; ++i)
        if (poset[(size_t)k * p + i] == 1) parents[k].push_back(i);
    std::vector<int> topo(topo_path, topo_path + p);
    float W = 0;
    {
      double s = 0;
      for (int i = 0; i < N; ++i) s += weights[i];
      W = (float)s;
    }
    std::vector<double> lambda(ilambda, ilambda + p), avg_lambda(p, 0.0), T_colsum(p);
    float avg_llhood = 0, llhood = 0;
    const int record_iter = std::max((int)((1 - alpha) * max_iter), 1);
    std::vector<int> O_vec(p);
    for (int iter = 0; iter < max_iter; ++iter) {
      std::fill(T_colsum.begin(), T_colsum.end(), 0.0);
      for (int i = 0; i < N; ++i) {  // generateMutationTimes (:280-304)
        for (int j = 0; j < p; ++j) O_vec[j] = O[(size_t)j * N + i];
        std::vector<double> td = expectedTimeDifference(O_vec, times[i], lambda, topo, parents, nrOfSamples, lambda_s,
                                                        time_available != 0, rng);
        for (int j = 0; j < p; ++j) T_colsum[j] += weights[i] * td[j];
      }
      double sumlog = 0, dot = 0;
      for (int j = 0; j < p; ++j) {
        lambda[j] = W / T_colsum[j];
        if (lambda[j] > max_lambda) lambda[j] = max_lambda;
        sumlog += std::log(lambda[j]);
        dot += lambda[j] * T_colsum[j];
      }
      llhood = (float)(W * sumlog - dot);
      if (iter >= record_iter) {
        for (int j = 0; j < p; ++j) avg_lambda[j] += lambda[j];
        avg_llhood += llhood;
      }
    }
    for (int j = 0; j < p; ++j) lambda_out[j] = avg_lambda[j] / (max_iter - record_iter);
    *llhood_out = (double)(avg_llhood / (max_iter - record_iter));
  } catch (const std::exception& e) {
    return 9;
  }
  return 0;
}

static std::vector<std::vector<int>> ot_parents(const int* poset, int p) {  // src/mcem.cpp:232-240
  std::vector<std::vector<int>> parents(p);
  for (int k = 0; k < p; ++k)
    for (int i = 0; i < p; ++i)
      if (poset[(size_t)k * p + i] == 1) parents[k].push_back(i);
  return parents;
}

// drawHiddenVarsSamples (src/mcem.cpp:214-248): T is N x p column-major, densities N
extern "C" int ref_drawHiddenVarsSamples(int N, const int* O_vec, double time, const double* lambda, const int* poset,
                                         const int* topo_path, float lambda_s, int time_available, int seed, int p,
                                         double* T, double* densities) {
  using namespace ref;
  try {
    StdRng rng((unsigned)seed);
    const auto parents = ot_parents(poset, p);
    const std::vector<int> O(O_vec, O_vec + p), topo(topo_path, topo_path + p);
    const std::vector<double> lam(lambda, lambda + p);
    for (int i = 0; i < N; ++i) {
      const std::vector<double> t = drawHiddenVarsSample(O, time, lam, topo, parents, densities[i], lambda_s,
                                                         time_available != 0, rng);
      for (int j = 0; j < p; ++j) T[(size_t)j * N + i] = t[j];
    }
  } catch (const std::exception& e) {
    return 9;
  }
  return 0;
}

// expectedTimeDifference (src/mcem.cpp:251-276)
extern "C" int ref_expectedTimeDifference(int nrOfSamples, const int* O_vec, double time, const double* lambda,
                                          const int* poset, const int* topo_path, float lambda_s, int time_available,
                                          int seed, int p, double* out) {
  using namespace ref;
  try {
    StdRng rng((unsigned)seed);
    const auto parents = ot_parents(poset, p);
    const std::vector<int> O(O_vec, O_vec + p), topo(topo_path, topo_path + p);
    const std::vector<double> lam(lambda, lambda + p);
    const std::vector<double> td = expectedTimeDifference(O, time, lam, topo, parents, nrOfSamples, lambda_s,
                                                          time_available != 0, rng);
    std::copy(td.begin(), td.end(), out);
  } catch (const std::exception& e) {
    return 9;
  }
  return 0;
}

// The proposal density of drawHiddenVarsSample as a function of its OUTPUT (src/mcem.cpp:64-90 with the draw `tmp`
// recovered from T): for given waiting times T (N x p), genotype and per-sample sampling times, the log density the
// sampler reports.  Lets a test check the (T, densities) pairs of another generator to fp64 accuracy.
extern "C" int ref_otcbn_proposal_density(int N, const int* O_vec, const double* times, const double* lambda,
                                          const int* poset, const int* topo_path, int p, const double* T,
                                          double* densities) {
  using namespace ref;
  try {
    const auto parents = ot_parents(poset, p);
    std::vector<double> T_sum(p);
    for (int i = 0; i < N; ++i) {
      double dens = 0.0;
      const double time = times[i];
      for (int k = 0; k < p; ++k) {
        const int e = topo_path[k];
        double max_parents_t = 0.0;
        for (int u : parents[e])
          if (T_sum[u] > max_parents_t) max_parents_t = T_sum[u];
        const double t = T[(size_t)e * N + i];
        T_sum[e] = max_parents_t + t;
        if (O_vec[e] == 1) {
          dens += ot_dexp_log(t, lambda[e]) - ot_pexp_log(time - max_parents_t, lambda[e]);  // tmp = T[e]
        } else {
          const double tmp = T_sum[e] - std::max(time, max_parents_t);
          dens += ot_dexp_log(tmp, lambda[e]);
        }
      }
      densities[i] = dens;
    }
  } catch (const std::exception& e) {
    return 9;
  }
  return 0;
}

// ================================================================================================
// C interface (ctypes).  Return 0 = ok, 1 = not acyclic, 2 = zero weight, 9 = other error.
// ================================================================================================
using namespace ref;

static thread_local std::string g_err;
static int fail(const std::exception& e) {
  g_err = e.what();
  if (dynamic_cast<const not_acyclic*>(&e)) return 1;
  if (g_err == kZeroWeightMsg) return 2;
  return 9;
}
#define REF_TRY try {
#define REF_CATCH \
  }               \
  catch (const std::exception& e) { return fail(e); } \
  return 0;

static MatB obs_from_int(const int* obs, int N, int p) {
  MatB m(N, p);
  for (size_t k = 0; k < (size_t)N * p; ++k) m.a[k] = obs[k] != 0;  // as<MatrixXb>: int -> bool
  return m;
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Stage 1: topological order (parents first, event ids) + parent bitmasks by event id + acyclicity
int ref_flatten_poset(const int* poset, int p, int* topo, uint64_t* parent_mask, int* cyclic) {
  REF_TRY
  Model M(adjacency_mat2list(poset, p), p);
  M.has_cycles();
  *cyclic = M.cycle;
  if (M.cycle) return 0;
  M.topological_sort();
  int k = 0;
  for (auto v = M.topo_path.rbegin(); v != M.topo_path.rend(); ++v) topo[k++] = M.event_id[*v];
  auto par = M.get_direct_predecessors();
  for (int j = 0; j < p; ++j) {
    parent_mask[j] = 0;
    for (int u : par[j]) parent_mask[j] |= (uint64_t)1 << u;
  }
  REF_CATCH
}

int ref_compatible_observations(const int* obs, int N, const int* poset, int p, int* out) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  MatB o = obs_from_int(obs, N, p);
  std::vector<unsigned char> row(p);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < p; ++j) row[j] = o(i, j);
    out[i] = is_compatible(row.data(), M);
  }
  REF_CATCH
}
int ref_num_compatible_observations(const int* obs, int N, const int* poset, int p, int* out) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  *out = num_compatible_observations(obs_from_int(obs, N, p), M, N);
  REF_CATCH
}
int ref_num_incompatible_events(const int* obs, int N, const int* poset, int p, int* out) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  *out = num_incompatible_events(obs_from_int(obs, N, p), M);
  REF_CATCH
}
int ref_hamming_dist_mat(const int* x, int N, int p, const int* y, int* out) {
  REF_TRY
  MatB m = obs_from_int(x, N, p);
  std::vector<unsigned char> yy(p);
  for (int j = 0; j < p; ++j) yy[j] = y[j] != 0;
  auto d = hamming_dist_mat(m, yy.data());
  std::copy(d.begin(), d.end(), out);
  REF_CATCH
}
unsigned ref_n_choose_k(unsigned n, unsigned k) { return n_choose_k(n, k); }
// _neighbors (src/backward_sampling.cpp:79-103): out is nrows_sum x p column-major; ring optional
int ref_neighbors(int p, int k, const int* genotype, int* out, int* ring_out) {
  REF_TRY
  std::vector<unsigned char> g(p);
  for (int j = 0; j < p; ++j) g[j] = genotype[j] != 0;
  std::vector<int> ring;
  MatB s = backward_ball(g.data(), (unsigned)p, (unsigned)k, ring);
  for (size_t t = 0; t < s.a.size(); ++t) out[t] = s.a[t];
  if (ring_out) std::copy(ring.begin(), ring.end(), ring_out);
  REF_CATCH
}
// _coin_tossing (:105-122)
int ref_coin_tossing(double eps, int L, const int* genotype, int p, int seed, int* out) {
  REF_TRY
  Context ctx(seed);
  std::vector<unsigned char> g(p);
  for (int j = 0; j < p; ++j) g[j] = genotype[j] != 0;
  MatB s = replicate_rows(g.data(), p, L);
  coin_tossing(s, eps, ctx.rng);
  for (size_t t = 0; t < s.a.size(); ++t) out[t] = s.a[t];
  REF_CATCH
}

// _sample_genotypes (src/mcem_hcbn.cpp:1012-1050)
int ref_sample_genotypes(int N, const int* poset, int p, const double* lambda, double* T_events /*N x p out*/,
                         double* T_sampling /*N in/out*/, float lambda_s, int times_available, int seed,
                         int* samples /*N x p out*/) {
  REF_TRY
  Model M = build_model(poset, p, lambda_s);
  M.lambda.assign(lambda, lambda + p);
  Context ctx(seed);
  MatD Z(N, p);
  std::vector<double> Ts(T_sampling, T_sampling + N);
  MatB s = sample_genotypes((unsigned)N, M, Z, Ts, ctx.rng, times_available != 0);
  std::copy(Z.a.begin(), Z.a.end(), T_events);
  std::copy(Ts.begin(), Ts.end(), T_sampling);
  for (size_t t = 0; t < s.a.size(); ++t) samples[t] = s.a[t];
  REF_CATCH
}
// _sample_times (:1052-1084)
int ref_sample_times(int N, const int* poset, int p, const double* lambda, int seed, double* T_events, double* T_sum) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  M.lambda.assign(lambda, lambda + p);
  Context ctx(seed);
  MatD Z(N, p);
  MatD T = sample_times((unsigned)N, M, Z, ctx.rng);
  std::copy(Z.a.begin(), Z.a.end(), T_events);
  std::copy(T.a.begin(), T.a.end(), T_sum);
  REF_CATCH
}
// _generate_genotypes (:1086-1120)
int ref_generate_genotypes(const double* T_sum, int N, const int* poset, int p, double* T_sampling, float lambda_s,
                           int times_available, int seed, int* samples) {
  REF_TRY
  Model M = build_model(poset, p, lambda_s);
  Context ctx(seed);
  MatD T(N, p);
  std::copy(T_sum, T_sum + (size_t)N * p, T.a.begin());
  std::vector<double> Ts(T_sampling, T_sampling + N);
  MatB s = generate_genotypes(T, M, Ts, ctx.rng, times_available != 0);
  std::copy(Ts.begin(), Ts.end(), T_sampling);
  for (size_t t = 0; t < s.a.size(); ++t) samples[t] = s.a[t];
  REF_CATCH
}
// supplied-times forward pass: T_sum, hidden genotypes X, Hamming distances to y, weights (no RNG)
int ref_forward_from_times(const int* genotype, int L, const int* poset, int p, double eps, const double* Z,
                           const double* T_sampling, double* T_sum_out, int* X_out, int* dist_out, double* w_out) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  MatD Zm(L, p);
  std::copy(Z, Z + (size_t)L * p, Zm.a.begin());
  std::vector<double> Ts(T_sampling, T_sampling + L);
  MatD Tsum;
  MatB X = genotypes_from_times(M, Zm, Ts, &Tsum);
  std::vector<unsigned char> y(p);
  for (int j = 0; j < p; ++j) y[j] = genotype[j] != 0;
  auto d = hamming_dist_mat(X, y.data());
  if (T_sum_out) std::copy(Tsum.a.begin(), Tsum.a.end(), T_sum_out);
  if (X_out)
    for (size_t t = 0; t < X.a.size(); ++t) X_out[t] = X.a[t];
  for (int l = 0; l < L; ++l) {
    if (dist_out) dist_out[l] = d[l];
    if (w_out) w_out[l] = std::pow(eps, (double)d[l]) * std::pow(1 - eps, p - (double)d[l]);
  }
  REF_CATCH
}

// _generate_mutation_times (src/add_remove.cpp:286-321); uniforms != NULL => supplied randomness
int ref_generate_mutation_times(const int* obs, int N, const int* poset, int p, const double* lambda, double* time,
                                float lambda_s, int times_available, int seed, const double* uniforms, double* Tdiff,
                                double* dens) {
  REF_TRY
  Model M = build_model(poset, p, lambda_s);
  M.lambda.assign(lambda, lambda + p);
  Context ctx(seed);
  MatB o = obs_from_int(obs, N, p);
  std::vector<double> d(N, 0.0), Ts(time, time + N);
  MatD U;
  if (uniforms) {
    U = MatD(N, p);
    std::copy(uniforms, uniforms + (size_t)N * p, U.a.begin());
  }
  MatD T = generate_mutation_times(o, M, d, Ts, &ctx.rng, times_available != 0, uniforms ? &U : nullptr);
  std::copy(T.a.begin(), T.a.end(), Tdiff);
  std::copy(d.begin(), d.end(), dens);
  std::copy(Ts.begin(), Ts.end(), time);
  REF_CATCH
}
// _cbn_density_log (src/add_remove.cpp:323-335)
int ref_cbn_density_log(const double* time, int N, int p, const double* lambda, double* out) {
  REF_TRY
  MatD T(N, p);
  std::copy(time, time + (size_t)N * p, T.a.begin());
  auto r = cbn_density_log(T, std::vector<double>(lambda, lambda + p));
  std::copy(r.begin(), r.end(), out);
  REF_CATCH
}
// _scale_path_to_mutation (src/add_remove.cpp:218-242)
int ref_scale_path_to_mutation(const int* poset, int p, const double* lambda, double* out) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  M.lambda.assign(lambda, lambda + p);
  auto s = scale_path_to_mutation(M);
  std::copy(s.begin(), s.end(), out);
  REF_CATCH
}
// _complete_log_likelihood (src/mcem_hcbn.cpp:738-760)
int ref_complete_log_likelihood(const double* lambda, int p, double eps, const double* Tdiff, int N, const double* dist,
                                const double* weights, double* out) {
  REF_TRY
  MatD T(N, p);
  std::copy(Tdiff, Tdiff + (size_t)N * p, T.a.begin());
  *out = complete_log_likelihood(std::vector<double>(lambda, lambda + p), eps, T, std::vector<double>(dist, dist + N),
                                 std::vector<double>(weights, weights + N), 0);
  REF_CATCH
}
double ref_log_bernoulli_process(unsigned dist, double eps, unsigned p) { return log_bernoulli_process(dist, eps, p); }

// _importance_weight_genotype (src/mcem_hcbn.cpp:866-912).  L_out = number of rows actually produced.
int ref_importance_weight_genotype(const int* genotype, int L, const int* poset, int p, const double* lambda, double eps,
                                   double time, const char* sampling, const double* scale_cumulative, const int* d_pool,
                                   const double* Tdiff_pool, int K, int neighborhood_dist, float lambda_s,
                                   int times_available, int seed, int* L_out, double* w, int* dist, double* Tdiff) {
  REF_TRY
  Model M = build_model(poset, p, lambda_s);
  M.lambda.assign(lambda, lambda + p);
  M.epsilon = eps;
  Context ctx(seed);
  std::vector<unsigned char> y(p);
  for (int j = 0; j < p; ++j) y[j] = genotype[j] != 0;
  std::vector<double> sc;
  if (scale_cumulative) sc.assign(scale_cumulative, scale_cumulative + p);
  std::vector<int> dp;
  MatD Tp;
  if (d_pool && K > 0) {
    dp.assign(d_pool, d_pool + K);
    Tp = MatD(K, p);
    std::copy(Tdiff_pool, Tdiff_pool + (size_t)K * p, Tp.a.begin());
  }
  DataIS is = importance_weight(y.data(), (unsigned)L, M, time, sampling, sc, dp, Tp, (unsigned)neighborhood_dist,
                                ctx.rng, times_available != 0);
  *L_out = (int)is.w.size();
  std::copy(is.w.begin(), is.w.end(), w);
  std::copy(is.dist.begin(), is.dist.end(), dist);
  std::copy(is.Tdiff.a.begin(), is.Tdiff.a.end(), Tdiff);
  REF_CATCH
}

// _importance_weight (src/mcem_hcbn.cpp:914-1010): per-observation sum w, sum w^2, L_eff, E[d], E[Z]
int ref_importance_weight(const int* obs, int N, int L, const int* poset, int p, const double* lambda, double eps,
                          const double* times, const char* sampling_c, int neighborhood_dist, float lambda_s,
                          int times_available, int thrds, int seed, double* w_sum, double* w_sum_sq, int* L_eff,
                          double* e_dist, double* e_Tdiff /* N x p */) {
  REF_TRY
  const std::string sampling = sampling_c;
  MatB o = obs_from_int(obs, N, p);
  Model M = build_model(poset, p, lambda_s);
  M.lambda.assign(lambda, lambda + p);
  M.epsilon = eps;
  Context ctx(seed);
  unsigned K = 0;
  std::vector<double> sc;
  MatD Tdiff_pool, T_pool;
  if (sampling == "add-remove") {
    sc = scale_path_to_mutation(M);
  } else if (sampling == "pool") {
    K = (unsigned)p * (unsigned)L;
    Tdiff_pool = MatD((int)K, p);
    T_pool = sample_times(K, M, Tdiff_pool, ctx.rng);
  }
#ifdef _OPENMP
  omp_set_num_threads(thrds);
#endif
  auto rngs = ctx.get_auxiliary_rngs(thrds);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; ++i) {
#ifdef _OPENMP
    StdRng& rng_t = *rngs[omp_get_thread_num()];
#else
    StdRng& rng_t = *rngs[0];
#endif
    std::vector<unsigned char> y(p);
    for (int j = 0; j < p; ++j) y[j] = o(i, j);
    std::vector<int> d_pool;
    if (sampling == "pool") d_pool = pool_distances(T_pool, M, times[i], times_available != 0, y.data(), rng_t);
    DataIS is = importance_weight(y.data(), (unsigned)L, M, times[i], sampling, sc, d_pool, Tdiff_pool,
                                  (unsigned)neighborhood_dist, rng_t, times_available != 0);
    ObsIS r = reduce_is(is, (unsigned)p);
    if (L_eff) L_eff[i] = (int)r.npos;
    w_sum[i] = r.aux;
    w_sum_sq[i] = r.wsq;
    e_dist[i] = r.wd / r.aux;
    for (int j = 0; j < p; ++j) e_Tdiff[(size_t)j * N + i] = r.wz[j] / r.aux;
  }
  REF_CATCH
}

// _obs_log_likelihood (src/mcem_hcbn.cpp:762-797)
int ref_obs_log_likelihood(const int* obs, int N, const int* poset, int p, const double* lambda, double eps,
                           const double* weights, const double* times, int L, const char* sampling,
                           int neighborhood_dist, float lambda_s, int times_available, int thrds, int seed,
                           double* llhood) {
  REF_TRY
  Context ctx(seed);
  *llhood = obs_log_likelihood(obs_from_int(obs, N, p), poset, p, std::vector<double>(lambda, lambda + p), eps,
                               std::vector<double>(weights, weights + N), std::vector<double>(times, times + N),
                               (unsigned)L, sampling, (unsigned)neighborhood_dist, ctx, lambda_s, times_available != 0,
                               (unsigned)thrds);
  REF_CATCH
}

// _MCEM_hcbn (src/mcem_hcbn.cpp:804-864), argument order as the .Call.
// trace (optional, may be NULL): max_iter x (p+2) doubles [llhood, eps, lambda...] per iteration; n_iter_out = rows
int ref_MCEM_hcbn(const double* ilambda, const int* poset, const int* obs, const double* times, float lambda_s,
                  double eps, const double* weights, unsigned L, const char* sampling, unsigned max_iter,
                  unsigned update_step_size, double tol, float max_lambda, unsigned neighborhood_dist,
                  int times_available, int thrds, int verbose, int seed, int N, int p, double* lambda_out,
                  double* eps_out, double* llhood_out, double* trace, int* n_iter_out) {
  REF_TRY
  Model M = [&] {
    Model m(adjacency_mat2list(poset, p), p, lambda_s);
    m.update_lambda(std::vector<double>(ilambda, ilambda + p), max_lambda);
    m.epsilon = eps;
    m.has_cycles();
    if (m.cycle) throw not_acyclic();
    m.topological_sort();
    return m;
  }();
  ControlEM ctl;
  ctl.max_iter = max_iter;
  ctl.update_step_size = update_step_size;
  ctl.tol = tol;
  ctl.max_lambda = max_lambda;
  ctl.neighborhood_dist = neighborhood_dist;
  Context ctx(seed, verbose != 0);
  std::vector<double> tr;
  double ll = MCEM_hcbn(M, obs_from_int(obs, N, p), std::vector<double>(times, times + N),
                        std::vector<double>(weights, weights + N), L, sampling, ctl, times_available != 0,
                        (unsigned)thrds, ctx, &tr);
  std::copy(M.lambda.begin(), M.lambda.end(), lambda_out);
  *eps_out = M.epsilon;
  *llhood_out = ll;
  if (trace) std::copy(tr.begin(), tr.end(), trace);
  if (n_iter_out) *n_iter_out = (int)(tr.size() / (size_t)(p + 2));
  REF_CATCH
}

// One E-step + M-step from SUPPLIED forward draws (no RNG): every observation is scored against the same
// L supplied samples (Z: L x p, T_s: L; if times_available T_s is replaced per observation by times[i]).
// Follows importance_weight/forward (:371-388) + the per-observation reduction (:683-694) + M-step (:707-709).
// Outputs: per-observation sum w, sum w^2, E[d], E[Z] (N x p); totals obs_llhood, expected_dist, N_eff,
// Tdiff_colsum[p]; updated eps and lambda[p].
int ref_estep_forward_from_times(const int* obs, int N, const int* poset, int p, const double* lambda, double eps,
                                 const double* weights, const double* times, int times_available, int L,
                                 const double* Z, const double* T_sampling, float max_lambda, double* w_sum,
                                 double* w_sum_sq, double* e_dist, double* e_Tdiff, double* totals /*3 + p*/,
                                 double* eps_new, double* lambda_new) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  M.lambda.assign(lambda, lambda + p);
  M.epsilon = eps;
  MatB o = obs_from_int(obs, N, p);
  MatD Zm(L, p);
  std::copy(Z, Z + (size_t)L * p, Zm.a.begin());
  MatD Tsum(L, p, 0.0);
  propagate_times(M, Zm, Tsum);
  double N_eff = 0, obs_llhood = 0, expected_dist = 0;
  MatD eT(N, p, 0.0);
  bool zero = false;
  for (int i = 0; i < N; ++i) {
    DataIS is((unsigned)L, (unsigned)p);
    is.Tdiff = Zm;
    MatB X(L, p);
    for (int j = 0; j < p; ++j)
      for (int l = 0; l < L; ++l) X(l, j) = Tsum(l, j) <= (times_available ? times[i] : T_sampling[l]);
    std::vector<unsigned char> y(p);
    for (int j = 0; j < p; ++j) y[j] = o(i, j);
    is.dist = hamming_dist_mat(X, y.data());
    for (int l = 0; l < L; ++l) {
      double d = (double)is.dist[l];
      is.w[l] = std::pow(eps, d) * std::pow(1 - eps, p - d);
    }
    ObsIS r = reduce_is(is, (unsigned)p);
    w_sum[i] = r.aux;
    w_sum_sq[i] = r.wsq;
    if (r.aux > 0) {
      N_eff += weights[i];
      obs_llhood += weights[i] * std::log(r.aux / L);
      expected_dist += weights[i] * r.wd / r.aux;
      e_dist[i] = r.wd / r.aux;
      for (int j = 0; j < p; ++j) eT(i, j) = r.wz[j] / r.aux;
    } else {
      zero = true;
    }
  }
  if (zero) throw std::runtime_error(kZeroWeightMsg);
  std::copy(eT.a.begin(), eT.a.end(), e_Tdiff);
  totals[0] = obs_llhood;
  totals[1] = expected_dist;
  totals[2] = N_eff;
  M.update_epsilon(expected_dist / (N_eff * p), kDblEps);
  std::vector<double> nl(p);
  for (int j = 0; j < p; ++j) {
    double cs = 0;
    for (int i = 0; i < N; ++i) cs += weights[i] * eT(i, j);
    totals[3 + j] = cs;
    nl[j] = 1.0 / (cs / N_eff);
  }
  M.update_lambda(nl, max_lambda);
  *eps_new = M.epsilon;
  std::copy(M.lambda.begin(), M.lambda.end(), lambda_new);
  REF_CATCH
}

// Conditional-proposal weights from SUPPLIED uniforms (backward / bernoulli core, no RNG).
// X: L x p hidden genotypes (int), u: L x p uniforms (by event id), T_s: L sampling times.
// Outputs Z (L x p), log q (L), log P(Z) (L), incompatible flag (L).
int ref_conditional_from_uniforms(const int* X, int L, const int* poset, int p, const double* lambda, const double* u,
                                  const double* T_sampling, double* Tdiff, double* log_q, double* log_pz,
                                  int* incompatible) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  M.lambda.assign(lambda, lambda + p);
  MatB x = obs_from_int(X, L, p);
  MatD U(L, p);
  std::copy(u, u + (size_t)L * p, U.a.begin());
  std::vector<double> dens(L, 0.0), Ts(T_sampling, T_sampling + L);
  MatD T = generate_mutation_times(x, M, dens, Ts, nullptr, true, &U);
  auto lpz = cbn_density_log(T, M.lambda);
  std::copy(T.a.begin(), T.a.end(), Tdiff);
  for (int l = 0; l < L; ++l) {
    log_q[l] = dens[l];
    log_pz[l] = lpz[l];
    int inc = 0;
    for (int j = 0; j < p; ++j) inc |= T(l, j) < 0.0;
    incompatible[l] = inc;
  }
  REF_CATCH
}

// initialize_lambda + the ASA epsilon start value (src/asa.cpp:72-129, :640-642)
int ref_initialize_lambda(const int* obs, int N, const int* poset, int p, float lambda_s, float max_lambda,
                          double* lambda_out, double* eps_out) {
  REF_TRY
  Model M = build_model(poset, p, lambda_s);
  MatB o = obs_from_int(obs, N, p);
  initialize_lambda(M, o, max_lambda);
  M.update_epsilon((double)num_incompatible_events(o, M) / ((double)N * p), kDblEps);
  std::copy(M.lambda.begin(), M.lambda.end(), lambda_out);
  *eps_out = M.epsilon;
  REF_CATCH
}

// draw `n` exponentials exactly as rexp<StdRng> would from a fresh Context(seed) (RNG pinning test)
// _adaptive_simulated_annealing (src/asa.cpp:596-665); stats_out[6] = {accepted, move0..move3 counts, scored}
int ref_adaptive_simulated_annealing(const int* poset, const int* obs, const double* times, float lambda_s,
                                     const double* weights, unsigned L, const char* sampling, unsigned max_iter_EM,
                                     unsigned update_step_size, double tol, float max_lambda, float T0, float adap_rate,
                                     float acceptance_rate, int step_size, unsigned max_iter_ASA,
                                     unsigned neighborhood_dist, int adaptive, const char* outdir, int times_available,
                                     int thrds, int verbose, int seed, int N, int p, int score_mode, double* lambda_out,
                                     double* eps_out, int* poset_out, double* llhood_out, int* stats_out) {
  REF_TRY
  Model M = build_model(poset, p, lambda_s);
  MatB O = obs_from_int(obs, N, p);
  initialize_lambda(M, O, max_lambda);
  M.update_epsilon((double)num_incompatible_events(O, M) / ((double)N * p), std::numeric_limits<double>::epsilon());
  ControlEM cem;
  cem.max_iter = max_iter_EM;
  cem.update_step_size = update_step_size;
  cem.tol = tol;
  cem.max_lambda = max_lambda;
  cem.neighborhood_dist = neighborhood_dist;
  ControlSA csa;
  csa.outdir = outdir;
  csa.acceptance_rate = acceptance_rate;
  csa.T = T0;
  csa.adap_rate = adap_rate;
  csa.step_size = step_size;
  csa.max_iter = max_iter_ASA;
  csa.adaptive = adaptive != 0;
  Context ctx(seed, verbose != 0);
  AsaStats st;
  double ll = simulated_annealing(M, O, std::vector<double>(times, times + N), std::vector<double>(weights, weights + N),
                                  csa, L, sampling, cem, times_available != 0, (unsigned)thrds, ctx, score_mode, st);
  std::copy(M.lambda.begin(), M.lambda.end(), lambda_out);
  *eps_out = M.epsilon;
  *llhood_out = ll;
  // adjacency_list2mat (src/model_impl.cpp:355-363): entry (event(u), event(v)) = 1, column-major p x p
  for (int i = 0; i < p * p; ++i) poset_out[i] = 0;
  for (int u = 0; u < p; ++u)
    for (int v : M.out[u]) poset_out[(size_t)M.event_id[v] * p + M.event_id[u]] = 1;
  if (stats_out) {
    stats_out[0] = st.num_accept_total;
    for (int k = 0; k < 4; ++k) stats_out[1 + k] = st.moves[k];
    stats_out[5] = st.n_scored;
  }
  REF_CATCH
}

// _draw_sample (src/add_remove.cpp:244-284)
int ref_draw_sample(const int* genotype, const int* poset, int p, unsigned move, const double* q_prob, int seed,
                    int* sample, double* q_choice) {
  REF_TRY
  Model M = build_model(poset, p, 1.0f);
  std::vector<unsigned char> g(p);
  for (int j = 0; j < p; ++j) g[j] = genotype[j] != 0;
  const bool compatible = is_compatible(g.data(), M);
  std::vector<double> rw(p), aw(p);
  for (int j = 0; j < p; ++j) {
    rw[j] = g[j] ? q_prob[j] : 0.0;
    aw[j] = g[j] ? 0.0 : 1.0 / q_prob[j];
  }
  Context ctx(seed);
  const int idx_remove = rdiscrete_std(1, rw, ctx.rng)[0];
  const int idx_add = rdiscrete_std(1, aw, ctx.rng)[0];
  double q = 0;
  std::vector<unsigned char> s = draw_sample(g, M, move, rw, aw, q, idx_remove, idx_add, compatible);
  for (int j = 0; j < p; ++j) sample[j] = s[j];
  *q_choice = q;
  REF_CATCH
}

// Known-answer test of the third-party generator the reference is built on: the C++ standard ([rand.predef]) fixes the
// 10000th output of a default-constructed std::mt19937 to 4123659995.
unsigned ref_mt19937_kat() {
  std::mt19937 g;
  g.discard(9999);
  return (unsigned)g();
}

int ref_rexp(int n, double rate, int seed, double* out) {
  REF_TRY
  Context ctx(seed);
  rexp_fill(out, n, rate, ctx.rng);
  REF_CATCH
}

}  // extern "C"
