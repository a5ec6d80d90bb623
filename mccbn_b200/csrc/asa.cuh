// asa.cuh -- host driver of (adaptive) simulated annealing over posets (config 4 of BASELINE.json).
//
// Follows src/asa.cpp: propose_edge :167-469, simulated_annealing :475-594, initialize_lambda :72-129,
// heuristic_compatibility :133-162, write_poset :59-70, and the Model edit operations of src/model_impl.cpp
// (poset.hpp).  The move generator, caches, Metropolis rule, temperature schedule and output files stay on the
// host exactly as in the reference; what changes is WHERE the scores come from: every candidate poset is scored by
// a complete MCEM run on the device, and proposals are evaluated speculatively in batches:
//
//   a rejected proposal leaves the current poset unchanged and the reference memoises scores per (move, v1, v2)
//   until the next acceptance (:254-267, reset :442-444), so the next B proposals can be generated ahead of time
//   under the assumption "all rejected", their MCEM runs share the device (and the host synchronisations), and
//   the unchanged accept/reject logic then consumes them in order until the first acceptance; proposals generated
//   after an accepted one are discarded and the proposal RNG is rewound to the state right after it.
//
// RNG: the reference drives everything from one std::mt19937 (interleaved with the per-EM-iteration seeds of the
// OpenMP generators).  Here proposals (shuffles, move type, compatibility heuristic) and the Metropolis uniforms use
// two independent generators with the reference's seeding chain, and every MCEM score uses the Philox stream
// (seed, evaluation id = ASA step), so the result is a deterministic function of the seed and independent of the
// speculation depth.  Exact stream parity with the reference is neither possible nor required.
#pragma once
#include <array>
#include <deque>
#include <fstream>
#include <functional>
#include <numeric>
#include <random>

namespace mccbn {

// StdRng seeding chain (src/rng_utils.hpp:87-98)
static std::mt19937 make_std_rng(unsigned seed) {
  std::ranlux24_base seeder(seed);
  std::array<int, std::mt19937::state_size> words;
  for (auto& w : words) w = (int)seeder();
  std::seed_seq seq(words.begin(), words.end());
  std::mt19937 g;
  g.seed(seq);
  return g;
}

struct Candidate {
  Poset poset;
  std::vector<double> lambda;
  double eps = 0.0;
  double llhood = 0.0;
};

struct Proposal {
  int move = -1;
  unsigned v1 = 0, v2 = 0;
  bool valid = false;     // false: every tested pair was rejected by the heuristic / structure (reference warns)
  Candidate cand;         // proposed model (structure edited, parameters still those of the current model)
  double fraction_compatible_new = 0.0;
  bool cached = false;    // score taken from the memo tables
  int score_slot = -1;    // index into the batch's scoring list
  int same_as = -1;       // earlier proposal of the batch with the same (move, v1, v2)
  std::mt19937 rng_after; // proposal generator state right after this proposal
};

struct AsaControl {  // ControlSA (src/asa.hpp:19-64)
  float T = 50.0f;
  std::string outdir;
  float acceptance_rate = 0.1f, adap_rate = 0.3f;
  unsigned step_size = 20, max_iter = 1000;
  bool adaptive = true;
  float compatible_fraction_factor = 0.05f;
};

}  // namespace mccbn
