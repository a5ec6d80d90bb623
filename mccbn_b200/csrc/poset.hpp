// poset.hpp -- host-side poset model for libmccbn_b200: bitmask adjacency over <= 64 vertices.
//
// Replaces the reference's Boost.Graph `Model` (src/mcem.hpp:74-198, src/model_impl.cpp) for the hot path:
//   * stage 1 "poset flattening": adjacency matrix -> topological order + parent bitmasks by event id
//     (adjacency_mat2list model_impl.cpp:343-352, has_cycles :40-49, topological_sort :51-69,
//      get_direct_predecessors :72-87)
//   * the graph-edit operations the ASA move generator needs (add_edge :187-205, transitive_reduction_dag
//     :145-184, add_relation :234-279, remove_relation :281-310, swap_node :312-316, set_children :90-107)
// Vertices are 0..p-1; vertex v carries event_id[v] (identity until a swap_node).  Genotype columns and lambda
// are indexed by EVENT id, exactly as in the reference.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace mccbn {

constexpr int kMaxP = 64;

static inline int popc64(uint64_t x) { return __builtin_popcountll(x); }
static inline int ctz64(uint64_t x) { return __builtin_ctzll(x); }

// flattened view consumed by the device code
struct FlatPoset {
  int p = 0;
  int n_edges = 0;
  int topo[kMaxP];                 // event id at topological position k (parents first)
  int pos_of_event[kMaxP];         // inverse of topo
  uint64_t parent_ev[kMaxP];       // by event id: bitmask of direct parents (event ids)
  uint64_t ancestor_ev[kMaxP];     // by event id: bitmask of all (transitive) ancestors (event ids)
  uint64_t parent_pos[kMaxP];      // by position: bitmask of direct parents (positions)
  uint64_t has_children_pos = 0;   // bit k: position k has at least one child
};

struct Poset {
  int p = 0;
  uint64_t out[kMaxP];  // out[v]: bitmask of direct successors (vertices)
  uint64_t in[kMaxP];   // in[v] : bitmask of direct predecessors (vertices)
  int event_id[kMaxP];
  std::vector<int> topo_vertices;  // topological order of VERTICES (parents first)
  bool cycle = false;
  bool reduction_flag = false;
  uint64_t children[kMaxP];  // transitive successors by vertex (set_children)
  bool children_stale = true;

  Poset() { clear(0); }
  explicit Poset(int p_) { clear(p_); }
  void clear(int p_) {
    p = p_;
    std::memset(out, 0, sizeof(out));
    std::memset(in, 0, sizeof(in));
    std::memset(children, 0, sizeof(children));
    for (int i = 0; i < kMaxP; ++i) event_id[i] = i;
    topo_vertices.clear();
    cycle = reduction_flag = false;
    children_stale = true;
  }
  // adjacency_mat2list + Model(edge_list, p): poset(i, j) == 1 => edge i -> j (column-major int matrix)
  static Poset from_adjacency(const int* mat, int p) {
    Poset P(p);
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j)
        if (mat[(size_t)j * p + i] == 1) P.insert_edge(i, j);
    return P;
  }
  // adjacency_list2mat (model_impl.cpp:355-363): matrix over EVENT ids
  void to_adjacency(int* mat) const {
    std::memset(mat, 0, sizeof(int) * (size_t)p * p);
    for (int u = 0; u < p; ++u)
      for (uint64_t m = out[u]; m; m &= m - 1) mat[(size_t)event_id[ctz64(m)] * p + event_id[u]] = 1;
  }
  bool has_edge(int u, int v) const { return (out[u] >> v) & 1; }
  void insert_edge(int u, int v) {
    out[u] |= 1ull << v;
    in[v] |= 1ull << u;
  }
  void erase_edge(int u, int v) {
    out[u] &= ~(1ull << v);
    in[v] &= ~(1ull << u);
  }
  int num_edges() const {
    int n = 0;
    for (int v = 0; v < p; ++v) n += popc64(out[v]);
    return n;
  }

  // has_cycles + topological_sort in one pass (Kahn, smallest ready vertex first => deterministic).
  // Sets `cycle`; on success fills topo_vertices.  A self-loop counts as a cycle.
  bool sort() {
    topo_vertices.clear();
    int indeg[kMaxP];
    uint64_t ready = 0;
    for (int v = 0; v < p; ++v) {
      indeg[v] = popc64(in[v]);
      if (!indeg[v]) ready |= 1ull << v;
    }
    while (ready) {
      int v = ctz64(ready);
      ready &= ready - 1;
      topo_vertices.push_back(v);
      for (uint64_t m = out[v]; m; m &= m - 1) {
        int w = ctz64(m);
        if (--indeg[w] == 0) ready |= 1ull << w;
      }
    }
    cycle = (int)topo_vertices.size() != p;
    return !cycle;
  }

  // set_children (model_impl.cpp:90-107): transitive successor sets, reverse topological sweep
  void set_children() {
    for (auto it = topo_vertices.rbegin(); it != topo_vertices.rend(); ++it) {
      int u = *it;
      uint64_t c = 0;
      for (uint64_t m = out[u]; m; m &= m - 1) {
        int v = ctz64(m);
        c |= (1ull << v) | children[v];
      }
      children[u] = c;
    }
    children_stale = false;
  }
  void update_children(int u) {  // model_impl.cpp:110-121
    uint64_t c = 0;
    for (uint64_t m = out[u]; m; m &= m - 1) {
      int v = ctz64(m);
      c |= (1ull << v) | children[v];
    }
    children[u] = c;
  }

  // transitive_reduction_dag (model_impl.cpp:145-184): drop every edge whose head is reachable through another
  // successor; reduction_flag = "nothing had to be removed".  Requires a valid topo_vertices.
  void transitive_reduction_dag() {
    reduction_flag = true;
    int rank[kMaxP];
    for (int k = 0; k < p; ++k) rank[topo_vertices[k]] = k;
    uint64_t all_succ[kMaxP];
    std::memset(all_succ, 0, sizeof(all_succ));
    for (auto it = topo_vertices.rbegin(); it != topo_vertices.rend(); ++it) {
      int u = *it;
      // successors in increasing topological rank
      std::vector<int> succ;
      for (uint64_t m = out[u]; m; m &= m - 1) succ.push_back(ctz64(m));
      for (size_t a = 1; a < succ.size(); ++a)
        for (size_t b = a; b > 0 && rank[succ[b]] < rank[succ[b - 1]]; --b) std::swap(succ[b], succ[b - 1]);
      for (int v : succ) {
        if ((all_succ[u] >> v) & 1) {
          erase_edge(u, v);
          children_stale = true;
          reduction_flag = false;
        } else {
          all_succ[u] |= (1ull << v) | all_succ[v];
        }
      }
    }
  }

  // add_edge (model_impl.cpp:187-205): returns false if the edge already existed
  bool add_edge(int u, int v) {
    if (has_edge(u, v)) return false;
    insert_edge(u, v);
    if (sort()) {
      transitive_reduction_dag();
      children_stale = true;
    }
    return true;
  }
  void remove_edge(int u, int v) {  // :207-210
    erase_edge(u, v);
    children_stale = true;
  }
  // remove_redundant_edges (:212-232)
  void remove_redundant_edges(int u, int v) {
    for (uint64_t m = out[u]; m; m &= m - 1) {
      int w = ctz64(m);
      if (w == v || ((children[v] >> w) & 1)) erase_edge(u, w);
    }
    children[u] |= (1ull << v) | children[v];
    for (uint64_t m = in[u]; m; m &= m - 1) remove_redundant_edges(ctz64(m), v);
  }
  // add_relation (:234-279): add cover relation u -> v keeping the graph transitively reduced
  bool add_relation(int u, int v) {
    if (has_edge(u, v)) return false;
    insert_edge(u, v);
    if ((children[u] >> v) & 1) {  // v already reachable from u
      reduction_flag = false;
      return true;
    }
    if ((children[v] >> u) & 1) {  // would close a cycle
      cycle = true;
      return true;
    }
    if (!cycle) {
      for (uint64_t m = out[u]; m; m &= m - 1) {
        int w = ctz64(m);
        if ((children[v] >> w) & 1) erase_edge(u, w);
      }
      children[u] |= (1ull << v) | children[v];
      for (uint64_t m = in[u]; m; m &= m - 1) remove_redundant_edges(ctz64(m), v);
      sort();
      transitive_reduction_dag();
    }
    return true;
  }
  // remove_relation (:281-310): delete cover relation u -> v, re-adding the order relations it implied
  void remove_relation(int u, int v) {
    erase_edge(u, v);
    update_children(u);
    for (uint64_t m = out[v]; m; m &= m - 1) {
      int w = ctz64(m);
      if (!((children[u] >> w) & 1)) {
        insert_edge(u, w);
        children[u] |= (1ull << w) | children[w];
      }
    }
    for (uint64_t m = in[u]; m; m &= m - 1) {
      int pa = ctz64(m);
      update_children(pa);
      if (!((children[pa] >> v) & 1)) {
        insert_edge(pa, v);
        children[pa] |= (1ull << v) | children[v];
      }
    }
  }
  void swap_node(int u, int v) {  // :312-316
    std::swap(event_id[u], event_id[v]);
    children_stale = true;
  }

  // Stage 1.  Requires sort() == true.
  void flatten(FlatPoset& f) const {
    f.p = p;
    f.n_edges = num_edges();
    f.has_children_pos = 0;
    int pos_of_vertex[kMaxP];
    for (int k = 0; k < p; ++k) {
      int v = topo_vertices[k];
      pos_of_vertex[v] = k;
      f.topo[k] = event_id[v];
      f.pos_of_event[event_id[v]] = k;
    }
    for (int k = 0; k < p; ++k) {
      int v = topo_vertices[k];
      uint64_t pe = 0, pp = 0, anc = 0;
      for (uint64_t m = in[v]; m; m &= m - 1) {
        int u = ctz64(m);
        pe |= 1ull << event_id[u];
        pp |= 1ull << pos_of_vertex[u];
        anc |= (1ull << event_id[u]) | f.ancestor_ev[event_id[u]];
      }
      f.parent_ev[event_id[v]] = pe;
      f.ancestor_ev[event_id[v]] = anc;
      f.parent_pos[k] = pp;
      if (out[v]) f.has_children_pos |= 1ull << k;
    }
  }
};

// n_choose_k (src/backward_sampling.cpp:20-35)
static inline unsigned n_choose_k(unsigned n, unsigned k) {
  if (k > n) return 0;
  unsigned kk = k < n - k ? k : n - k;
  long double r = 1;
  for (unsigned i = 1; i <= kk; ++i) r = r * (n - kk + i) / i;
  return (unsigned)(r + 0.5L);
}

// Flip masks of the Hamming ball of radius k in the reference's enumeration order
// (neighbors(), src/backward_sampling.cpp:37-54, stacked ring by ring as in src/mcem_hcbn.cpp:497-509):
// entry 0 = no flip; bit j of a mask = column (event id) j is flipped.
static inline void ball_ring(unsigned p, unsigned k, unsigned col0, uint64_t prefix, std::vector<uint64_t>& out) {
  if (p == 0) {
    if (k == 0) out.push_back(prefix);
    return;
  }
  if (n_choose_k(p - 1, k) > 0) ball_ring(p - 1, k, col0 + 1, prefix, out);                       // column as observed
  if (k >= 1 && n_choose_k(p - 1, k - 1) > 0) ball_ring(p - 1, k - 1, col0 + 1, prefix | (1ull << col0), out);  // flipped
}
static inline std::vector<uint64_t> hamming_ball_masks(unsigned p, unsigned k, std::vector<int>* ring = nullptr) {
  std::vector<uint64_t> m;
  if (ring) ring->clear();
  for (unsigned i = 0; i <= k; ++i) {
    size_t before = m.size();
    ball_ring(p, i, 0, 0, m);
    if (ring) ring->insert(ring->end(), m.size() - before, (int)i);
  }
  return m;
}

}  // namespace mccbn
