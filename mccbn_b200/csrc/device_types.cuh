// device_types.cuh -- data layout shared by host engine and kernels (libmccbn_b200)
//
// HBM layout (per problem = one shard of observations on one GPU):
//   obs_bits   : N x uint64   bit j = event j observed (bit-packed once from R's N x p int32 matrix)
//   times      : N x double   sampling times t_i (only read when sampling_times_available)
//   weights    : N x double   observation weights wt_i
//   part       : n_cand x n_items x PS doubles   per-(observation, chunk) partial sums of one E-step
//   obs_out    : optional N x (3 + p) doubles    per-observation sum w, sum w^2, E[d], E[Z_j]
//   blk_part   : n_cand x n_blocks x PS doubles  deterministic second-stage partials
//   stats      : n_cand x PS doubles             [LL, N_eff, D, zero_weight_count, 0, S_0..S_{p-1}] -- the vector
//                                                 that is all-reduced across GPUs
// The static poset program is a __grid_constant__ kernel parameter (PosetProg); the lambda-derived scales live
// in __constant__ memory (FwdScale), refreshed by a device-to-device copy after every on-device M-step.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#else  // NVRTC (poset-specialised kernels, jit.cuh): no standard headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#endif

namespace mccbn {

constexpr int kFastMaxP = 32;      // one 32-bit genotype word on the Philox fast path
constexpr int kWideMaxP = 64;      // forward scheme only: 64-bit genotype word, generic kernel (PosetProgWide)
constexpr int kMaxCand = 16;       // concurrently scored candidate posets (ASA batching)
constexpr int kMaxEdgesFast = 512; // >= 32*31/2
constexpr int kFwdBlock = 128;     // threads per block of the forward E-step kernel
constexpr int kStatHdr = 5;        // partial record: [sum w', sum w'^2, sum w' d, #{w>0}, log scale]; stats: [LL, N_eff, D, zero_count, -]
constexpr int kMaxWorld = 16;      // ranks of one NVLink domain in the peer-memory statistic exchange
constexpr int kMaxPSlots = 72;     // >= kStatHdr + 64
constexpr int kRtabN = 72;         // relative-weight table r^delta, delta = 0..64, tail = 0

// static structure of one poset, topological-position space (uploaded by the host)
struct DevPoset {
  int p;
  int n_edges;
  unsigned char ev[64];        // event id at topological position k
  unsigned short pstart[68];   // CSR offsets into ppos (fast path, p <= 32)
  unsigned char ppos[kMaxEdgesFast];  // parent positions
  uint64_t parent_pos[64];     // by position: parents (positions)
  uint64_t parent_ev[64];      // by event id: parents (event ids)
  uint64_t ancestor_ev[64];    // by event id: all ancestors (event ids)
  uint64_t has_children_pos;
};

// mutable model parameters + EM bookkeeping of one candidate (device resident)
struct DevModel {
  double lambda[64];  // by event id
  double eps;
  double lambda_s;    // the reference keeps this as float (src/mcem.hpp:190); widened once on the host
  double max_lambda;  // float in the reference (src/mcem.hpp:218)
  double llhood;      // obs_llhood of the last E-step
  // weight tables for the forward scheme: w = base * r^d' with d' = flip ? p - d : d  (r <= 1 always)
  double log_base;    // flip ? p*log(eps) : p*log1p(-eps)
  double log_r;       // log(min(r, 1/r)), r = eps/(1-eps)
  double rtab_d[kRtabN];
  float rtab[kRtabN];
  int flip;
  int iter;           // EM iterations done since set_model
  int zero_weight;    // sticky: some observation had sum w == 0
  int comm_error;     // sticky: the peer-memory statistic exchange timed out waiting for a rank
};

// Peer-memory statistic exchange (kernels_reduce.cuh): every rank owns one mailbox in its HBM; the reduction kernel of
// rank r stores its statistic vector into slot[parity][r] of EVERY rank's mailbox over NVLink and then publishes the
// exchange number in flag[r]; the M-step kernel waits for all flags and sums the slots in rank order, so all ranks
// compute bit-identical parameters without a collective call.
struct Mailbox {
  unsigned long long flag[kMaxWorld];
  double slot[2][kMaxWorld][kMaxPSlots];
};
struct PeerView {
  Mailbox* box[kMaxWorld];  // box[q] = rank q's mailbox mapped into this process (box[rank] is local memory)
  int rank, world;
  unsigned long long seq;   // number of this exchange (1, 2, ...), identical on all ranks
  long long timeout_clocks; // give up waiting for a rank after this many SM clocks (a dead rank must not hang the GPU)
};

// Static "program" of one poset for the sampling kernels.  It is known to the host, so it travels as a
// __grid_constant__ kernel parameter: every field is a warp-uniform constant-bank operand at a fixed offset
// (no per-thread loads, no dynamic indexing).  Nodes are addressed by topological position k; T_k of nodes that
// have children is staged in a per-thread shared-memory column, row PMAX of that column holds 0 so that
// absent parents can be loaded unconditionally (max with 0 is the reference's T_max initialisation).
struct PosetProg {
  typedef uint32_t word_t;  // genotype word of the kernels driven by this program
  int p;
  unsigned store_mask;   // bit k: T_k has children -> stage it
  unsigned multi_mask;   // bit k: node k has more than two parents -> `xo[k]` holds parents 3..6
  unsigned valid_mask;   // (1 << p) - 1
  unsigned huge_mask;    // bit k: node k has more than six parents -> CSR loop over the rest
  unsigned one_bits;     // 0x3f800000 as a run-time value: lets (x & mask) | one compile to a single LOP3
  unsigned pad[2];
  uint2 off[kFastMaxP];            // byte offsets (within the thread's column) of the first two parents' rows
  uint4 xo[kFastMaxP];             // parents 3..6 of nodes with in-degree > 2 (dummy row where absent)
  unsigned char ev[kFastMaxP];     // event id at position k
  unsigned parent_mask[kFastMaxP]; // parents of k as a bitmask over positions (compatibility checks)
  unsigned short xstart[kFastMaxP + 2];  // CSR over `xoff`: seventh and further parents of node k
  unsigned short xoff[kMaxEdgesFast];
};

// The same program for 32 < p <= 64 (forward scheme, generic kernel): 64-bit masks, no compatibility masks.
struct PosetProgWide {
  typedef uint64_t word_t;
  int p;
  unsigned one_bits;
  uint64_t store_mask, multi_mask, valid_mask, huge_mask;
  uint2 off[kWideMaxP];
  uint4 xo[kWideMaxP];
  unsigned char ev[kWideMaxP];
  uint64_t parent_mask[kWideMaxP];  // parents of k as a bitmask over positions (compatibility checks)
  unsigned short xstart[kWideMaxP + 2];
  unsigned short xoff[kMaxEdgesFast];
};

// lambda-derived scales of the exponential draws, refreshed on the device after every M-step and mirrored into
// __constant__ memory so that they are immediate constant-bank operands of the FMULs.
struct FwdScale {
  float c[kWideMaxP];  // -ln2 / lambda at topological position k:  Z = log2(u) * c
  float cs;            // -ln2 / lambda_s
  int flip;            // 1 if eps > 1/2 (weights increase with the Hamming distance)
  int pad[2];
};

}  // namespace mccbn
