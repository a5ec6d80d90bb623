// jit.cuh -- poset-specialised forward E-step kernels through NVRTC.
//
// The poset is a run-time input, but it is constant for a whole MCEM run (hundreds of E-steps over 1e7..1e9 samples
// each).  The generic kernel (kernels_forward.cuh) interprets the poset program per sample: event times live in a
// shared-memory column, every node pays two unconditional LDS + one STS + constant-bank loads + a warp-uniform test
// for further parents (ncu r1b: 1167 instructions per sample for p = 30).  When the sampling work is large enough to
// amortise a compilation, the same hand-written kernel body (estep_forward_body, #included verbatim from the embedded
// headers) is instantiated with a sampler in which the poset is unrolled into straight-line code: event times in
// registers, exactly |E| FMNMX, immediate bit masks, no parent loads / stores / tests.  The random stream, the weight
// handling, the reduction and the output records are those of the generic kernel, so both produce the same
// per-sample values (tests/test_gpu_jit.py compares them bit for bit on the dump seam and to 1e-6 on the sums).
//
// NVRTC (libnvrtc.so.12) and the loaded module are reached through dlopen / the CUDA runtime library-management API;
// if NVRTC is unavailable or compilation fails the generic kernel is used (still the device path, never the CPU).
#pragma once
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <future>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "_embedded_sources.inc"
#include "engine.cuh"
#include "kernels_forward.cuh"
#include "kernels_cond.cuh"

namespace mccbn {

// Background compilations in flight, process-wide.  The default context is never destroyed, so its worker threads may
// still be inside nvrtcCompileProgram when the process exits -- and libnvrtc's own static destructors would then run
// under them (observed: SIGSEGV inside nvrtcCompileProgram at exit after a short annealing run).  An atexit handler that
// runs BEFORE their teardown waits for the workers.
static std::atomic<int> g_jit_bg_inflight{0};
static void jit_wait_background_at_exit() {
  for (int i = 0; i < 3000 && g_jit_bg_inflight.load(std::memory_order_acquire) > 0; ++i)
    std::this_thread::sleep_for(std::chrono::milliseconds(10));  // a compilation takes ~0.3 s; give up after 30 s
}
// atexit handlers and static destructors run in reverse order of registration, and libnvrtc (LLVM) keeps creating
// function-local statics during its first compilations -- each registers its destructor when it is first used.  The
// wait must run before ALL of them, so it is registered again after libnvrtc is loaded and after each of the first 64
// completed compilations (by then every lazily created object exists; the handler returns at once when nothing is in
// flight).
static void jit_register_exit_wait() {
  static std::atomic<int> n{0};
  if (n.fetch_add(1) < 65) std::atexit(jit_wait_background_at_exit);
}

struct NvrtcApi {
  typedef struct _nvrtcProgram* program_t;
  void* lib = nullptr;
  int (*CreateProgram)(program_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(program_t, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(program_t, size_t*) = nullptr;
  int (*GetCUBIN)(program_t, char*) = nullptr;
  int (*GetProgramLogSize)(program_t, size_t*) = nullptr;
  int (*GetProgramLog)(program_t, char*) = nullptr;
  int (*DestroyProgram)(program_t*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  bool tried = false;

  bool load() {
    if (lib) return true;
    if (tried) return false;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
    for (const char* n : names)
      if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib) return false;
    bool ok = true;
    auto sym = [&](const char* s) {
      void* f = dlsym(lib, s);
      if (!f) ok = false;
      return f;
    };
    CreateProgram = (decltype(CreateProgram))sym("nvrtcCreateProgram");
    CompileProgram = (decltype(CompileProgram))sym("nvrtcCompileProgram");
    GetCUBINSize = (decltype(GetCUBINSize))sym("nvrtcGetCUBINSize");
    GetCUBIN = (decltype(GetCUBIN))sym("nvrtcGetCUBIN");
    GetProgramLogSize = (decltype(GetProgramLogSize))sym("nvrtcGetProgramLogSize");
    GetProgramLog = (decltype(GetProgramLog))sym("nvrtcGetProgramLog");
    DestroyProgram = (decltype(DestroyProgram))sym("nvrtcDestroyProgram");
    GetErrorString = (decltype(GetErrorString))sym("nvrtcGetErrorString");
    if (!ok) lib = nullptr;
    else Version = (decltype(Version))dlsym(lib, "nvrtcVersion");
    if (ok) jit_register_exit_wait();
    return ok;
  }
};

// ---- source generation ------------------------------------------------------------------------------------------
static int jit_pmax(int p) { return (p + 3) / 4 * 4; }

static std::string jit_key(const FlatPoset& f, bool times_given, int minb) {
  std::ostringstream k;
  k << f.p << (times_given ? "t" : "s") << minb;
  for (int i = 0; i < f.p; ++i) k << ":" << std::hex << f.parent_pos[i];
  return k.str();
}

// Same arithmetic and random stream as sample_forward (kernels_forward.cuh, fp32 path), poset unrolled:
//   - 24-bit fields packed back to back: field 0 = sampling time, field k + 1 = node k; stream word 4 b + i is state
//     word i of Philox block b (variable a<i>_<b>); blocks are generated two at a time in lock-step (independent
//     multiplies of a round hide the IMAD.WIDE -> LOP3 latency), right before the first node that needs them
//   - event times relative to the sampling time, in registers: t<k> = e<k> c_k + max(parents) (roots: - T_s), the
//     genotype bit is the sign bit of t<k>, shifted into x with one funnel shift (node k -> bit p - 1 - k)
//   - e<k> = log2(u_k) stays in registers (E[k]) for the weighted sums: no shared-memory round trip
static std::string generate_forward_source(const FlatPoset& f, bool times_given, int minb) {
  const int p = f.p, pmax = jit_pmax(p);
  // p > 32: 64-bit genotype words and the wide program
  const char* prog = p > kFastMaxP ? "PosetProgWide" : "PosetProg";
  const char* word_t = p > kFastMaxP ? "uint64_t" : "uint32_t";
  std::ostringstream s;
#ifdef MCCBN_DIAG_BUILD  // tools/ profiling builds only (./build.sh -DMCCBN_DIAG_BUILD): never in the shipped library
  if (const char* e = std::getenv("MCCBN_JIT_DIAG_ROUNDS"))
    s << "#define MCCBN_PHILOX_ROUNDS " << std::max(1, std::min(10, std::atoi(e))) << "\n";
#endif
  s << "#include \"kernels_forward.cuh\"\n"
       "namespace mccbn {\n"
       "// generated for one poset: node k = topological position k, t<k> = event time - sampling time\n"
       "struct SpecSampler {\n"
       "  static constexpr bool kStagesT = false;\n"
       "  static constexpr bool kStagesE = false;\n"
       "  __device__ __forceinline__ void operator()(const " << prog << "& pg, const uint32_t tcol_s, const uint32_t zrow_s,\n"
       "      const uint32_t l, const uint32_t gi, const uint32_t k0, const uint32_t k1, const uint32_t cw,\n"
       "      const float ts_given, const ScaleSrc<float, " << pmax << "> scale, " << word_t << "& X, float& Ts_out,\n"
       "      float (&E)[" << pmax << "]) const {\n"
       "    " << word_t << " x = 0;\n"
       "    const uint32_t one = pg.one_bits;\n";
  const int nblocks = fwd_blocks_for_fields(p + 1);
  auto word = [&](int wi) { return "a" + std::to_string(wi % 4) + "_" + std::to_string(wi / 4); };
  auto field = [&](int f) {  // the expression fast_lg2(field_u01(W, f, one)) folds to
    const int st = 24 * f, wi = st / 32, o = st % 32;
    std::string m;
    if (o == 0) m = "((" + word(wi) + " & 0x7fffffu) | one)";
    else if (o == 8) m = "((" + word(wi) + " >> 9) | 0x3f800000u)";
    else m = "((__funnelshift_r(" + word(wi) + ", " + word(wi + 1) + ", " + std::to_string(o) + ") & 0x7fffffu) | one)";
    return "fast_lg2(2.0f - __uint_as_float(" + m + "))";
  };
  auto philox_init = [&](int b) {
    std::ostringstream o;
    o << "    uint32_t a0_" << b << " = l, a1_" << b << " = gi, a2_" << b << " = " << b << "u, a3_" << b << " = cw;\n";
    return o.str();
  };
  auto philox_round = [&](int b, int r) {  // same arithmetic as philox4x32_10, written out round by round
    std::ostringstream o;
    o << "    { const uint64_t p0 = (uint64_t)0xD2511F53u * a0_" << b << ", p1 = (uint64_t)0xCD9E8D57u * a2_" << b << ";"
      << " a0_" << b << " = (uint32_t)(p1 >> 32) ^ a1_" << b << " ^ (k0 + " << (uint32_t)(0x9E3779B9u * (uint32_t)r) << "u);"
      << " a2_" << b << " = (uint32_t)(p0 >> 32) ^ a3_" << b << " ^ (k1 + " << (uint32_t)(0xBB67AE85u * (uint32_t)r) << "u);"
      << " a1_" << b << " = (uint32_t)p1; a3_" << b << " = (uint32_t)p0; }\n";
    return o.str();
  };
  const int rounds = MCCBN_PHILOX_ROUNDS;
  auto node_stmt = [&](int k) {
    std::ostringstream o;
    o << "    const float e" << k << " = " << field(k + 1) << ";";
    // max over parents, balanced so that the compiler can form FMNMX3
    std::vector<std::string> terms;
    for (uint64_t m = f.parent_pos[k]; m; m &= m - 1) terms.push_back("t" + std::to_string(ctz64(m)));
    while (terms.size() > 1) {
      std::vector<std::string> nxt;
      for (size_t i = 0; i + 1 < terms.size(); i += 2) nxt.push_back("fmaxf(" + terms[i] + ", " + terms[i + 1] + ")");
      if (terms.size() & 1) nxt.push_back(terms.back());
      terms.swap(nxt);
    }
    o << "  const float t" << k << " = fmaf(e" << k << ", c_scale.c[" << k << "], " << (terms.empty() ? "nts" : terms[0])
      << ");  E[" << k << "] = e" << k << ";  ";
    if (p > kFastMaxP) o << "x = (x << 1) | (uint64_t)(__float_as_uint(t" << k << ") >> 31);\n";
    else o << "x = __funnelshift_l(__float_as_uint(t" << k << "), x, 1);\n";
    return o.str();
  };
  // nts = -T_s: the scale c_scale.cs = -ln2/lambda_s is negative, T_s = log2(u) cs >= 0
  const std::string ts_stmt = times_given ? std::string("    const float nts = -ts_given;\n")
                                          : "    const float nts = -(" + field(0) + " * c_scale.cs);\n";
  int G = 2;  // Philox blocks advanced in lock-step
  if (const char* e = std::getenv("MCCBN_JIT_PHILOX_ILP")) G = std::max(1, std::min(8, std::atoi(e)));
  int generated = 0;  // blocks generated so far
  for (int k = 0; k < p; ++k) {
    while (generated <= fwd_field_block(k + 1)) {
      const int b0 = generated, g = std::min(G, nblocks - b0);
      for (int b = b0; b < b0 + g; ++b) s << philox_init(b);
      for (int r = 0; r < rounds; ++r)
        for (int b = b0; b < b0 + g; ++b) s << philox_round(b, r);
      if (b0 == 0) s << ts_stmt;
      generated += g;
    }
    s << node_stmt(k);
  }
  for (int k = p; k < pmax; ++k) s << "    E[" << k << "] = 0.0f;\n";
  s << "    X = x;\n"
       "    Ts_out = -nts;\n"
       "  }\n"
       "};\n"
       "}  // namespace mccbn\n"
       "extern \"C\" __global__ void __launch_bounds__(mccbn::kFwdBlock, "
    << minb
    << ") mccbn_fwd_spec(const mccbn::FwdArgs a,\n"
       "    const __grid_constant__ mccbn::" << prog << " pg) {\n"
       "  mccbn::estep_forward_body<float, "
    << pmax << ", " << (times_given ? "true" : "false")
    << ">(a, pg, mccbn::SpecSampler{});\n"
       "}\n";
  return s.str();
}

// ---- conditional proposal (backward / bernoulli / OT-CBN / add-remove): the same idea for estep_conditional_body ----
// Same random stream and arithmetic as sample_conditional (kernels_cond.cuh): packed 24-bit fields, field 0 = sampling
// time, field k + 1 = node k.
static std::string generate_cond_source(const FlatPoset& f, bool times_given, int mode, int minb) {
  const int p = f.p, pmax = jit_pmax(p);
  // p > 32: 64-bit genotype words and the wide program; these sizes exist as specialised kernels only
  const char* prog = p > kFastMaxP ? "PosetProgWide" : "PosetProg";
  const char* word_t = p > kFastMaxP ? "uint64_t" : "uint32_t";
  std::ostringstream s;
  s << "#include \"kernels_cond.cuh\"\n"
       "namespace mccbn {\n"
       "struct SpecCondSampler {\n"
       "  static constexpr bool kStagesT = false;\n"
       "  __device__ __forceinline__ float operator()(const " << prog << "& pg, const uint32_t tcol_s,\n"
       "      const float* __restrict__ s_lam2, const float* __restrict__ s_inv, const uint32_t l, const uint32_t gi,\n"
       "      const uint32_t k0, const uint32_t k1, const uint32_t cw, const float ts_given, const " << word_t << " Xt,\n"
       "      float (&Z)[" << pmax << "], float& Ts_out) const {\n"
       "    float Ts = ts_given;\n"
       "    float lw2 = 0.0f;\n";
  // stream word 4 b + i is component i of Philox block b (variable r<b>); fields as in generate_forward_source
  const char* comp[4] = {".x", ".y", ".z", ".w"};
  auto word = [&](int wi) { return "r" + std::to_string(wi / 4) + comp[wi % 4]; };
  auto field = [&](int fi) {  // the expression field_u01_open1 folds to: u in [0, 1)
    const int st = 24 * fi, wi = st / 32, o = st % 32;
    std::string m;
    if (o == 0) m = "((" + word(wi) + " & 0x7fffffu) | one)";
    else if (o == 8) m = "((" + word(wi) + " >> 9) | 0x3f800000u)";
    else m = "((__funnelshift_r(" + word(wi) + ", " + word(wi + 1) + ", " + std::to_string(o) + ") & 0x7fffffu) | one)";
    return "(__uint_as_float(" + m + ") - 1.0f)";
  };
  int generated = 0;
  auto need_block = [&](int b) {
    while (generated <= b) {
      s << "    const uint4 r" << generated << " = philox4x32_10(l, gi, " << generated << "u, cw, k0, k1);\n";
      ++generated;
    }
  };
  s << "    const uint32_t one = pg.one_bits;\n";
  need_block(0);
  if (!times_given) s << "    Ts = CondMath<float>::z_of(" << field(0) << ", 1.0f, s_inv[" << pmax << "]);\n";
  for (int k = 0; k < p; ++k) {
    need_block(fwd_field_block(k + 1));
    std::vector<std::string> terms;
    for (uint64_t m = f.parent_pos[k]; m; m &= m - 1) terms.push_back("t" + std::to_string(ctz64(m)));
    while (terms.size() > 1) {
      std::vector<std::string> nxt;
      for (size_t i = 0; i + 1 < terms.size(); i += 2) nxt.push_back("fmaxf(" + terms[i] + ", " + terms[i + 1] + ")");
      if (terms.size() & 1) nxt.push_back(terms.back());
      terms.swap(nxt);
    }
    const std::string K = std::to_string(k);
    s << "    const float m" << K << " = " << (terms.empty() ? "0.0f" : terms[0]) << ";\n"
      << "    const bool x" << K << " = (Xt >> " << K << ") & 1;\n"
      << "    const float a" << K << " = Ts - m" << K << ";\n"
      << "    const float Fx" << K << " = CondMath<float>::F_of(a" << K << ", s_lam2[" << K << "]);\n"
      << "    const float z" << K << " = CondMath<float>::z_of(" << field(k + 1) << ", x" << K << " ? Fx" << K
      << " : 1.0f, s_inv[" << K << "]);\n"
      << "    const float t" << K << " = (x" << K << " ? m" << K << " : fmaxf(m" << K << ", Ts)) + z" << K << ";\n"
      << "    Z[" << K << "] = t" << K << " - m" << K << ";\n"
      << "    lw2 += x" << K << " ? CondMath<float>::lg2(Fx" << K << ") : -s_lam2[" << K << "] * fmaxf(a" << K << ", 0.0f);\n";
  }
  for (int k = p; k < pmax; ++k) s << "    Z[" << k << "] = 0.0f;\n";
  s << "    Ts_out = Ts;\n"
       "    return lw2;\n"
       "  }\n"
       "};\n"
       "}  // namespace mccbn\n"
       "extern \"C\" __global__ void __launch_bounds__(mccbn::kFwdBlock, "
    << minb << ") mccbn_cond_spec(const mccbn::CondArgs a,\n"
       "    const __grid_constant__ mccbn::" << prog << " pg) {\n"
       "  mccbn::estep_conditional_body<float, " << pmax << ", " << (times_given ? "true" : "false") << ", " << mode
    << ">(a, pg, mccbn::SpecCondSampler{});\n"
       "}\n";
  return s.str();
}

struct JitCompiled {
  std::vector<char> cubin;
  std::string log;
  double compile_ms = 0;
};

// compile only (no device needed): returns false + log on failure
static bool jit_compile(NvrtcApi& rtc, const std::string& src, JitCompiled& out) {
  NvtxRange nvtx("NVRTC compile specialised kernel");
  if (!rtc.load()) {
    out.log = "libnvrtc.so.12 could not be loaded";
    return false;
  }
  const char* hdr_src[] = {kSrc_device_types_cuh, kSrc_philox_cuh, kSrc_kernels_forward_cuh, kSrc_kernels_cond_cuh};
  const char* hdr_name[] = {"device_types.cuh", "philox.cuh", "kernels_forward.cuh", "kernels_cond.cuh"};
  NvrtcApi::program_t prog = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = rtc.CreateProgram(&prog, src.c_str(), "mccbn_spec.cu", 4, hdr_src, hdr_name);
  if (rc != 0) {
    out.log = std::string("nvrtcCreateProgram: ") + rtc.GetErrorString(rc);
    return false;
  }
  // ptxas -O1 keeps the generated statement order (needed by the software-pipelined schedule experiment); the default
  // level hoists every Philox round to the front of the sample and leaves the node code at the end -- which is faster
  std::string popt = "--ptxas-options=-O3";
  if (const char* e = std::getenv("MCCBN_JIT_PTXAS_O")) popt = std::string("--ptxas-options=-O") + e;
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--ptxas-options=-v", popt.c_str()};
  rc = rtc.CompileProgram(prog, 5, opts);
  size_t ls = 0;
  rtc.GetProgramLogSize(prog, &ls);
  if (ls > 1) {
    out.log.resize(ls);
    rtc.GetProgramLog(prog, &out.log[0]);
  }
  bool ok = rc == 0;
  jit_register_exit_wait();  // after a compilation (successful or not): in front of whatever it created lazily
  if (ok) {
    size_t n = 0;
    ok = rtc.GetCUBINSize(prog, &n) == 0 && n > 0;
    if (ok) {
      out.cubin.resize(n);
      ok = rtc.GetCUBIN(prog, out.cubin.data()) == 0;
    }
  } else {
    out.log = std::string("nvrtcCompileProgram: ") + rtc.GetErrorString(rc) + "\n" + out.log;
  }
  rtc.DestroyProgram(&prog);
  out.compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return ok;
}

struct JitKernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  void* scale_sym = nullptr;
  size_t smem = 0;
  int pmax = 0;
  int blocks_per_sm = 0;  // resident 128-thread blocks per SM (occupancy query at load time)
  bool ok = false;
  double compile_ms = 0;
  uint64_t last_use = 0;  // LRU stamp (JitCache::tick)
};

static uint64_t fnv1a64(const void* data, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* c = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) {
    h ^= c[i];
    h *= 1099511628211ull;
  }
  return h;
}
static uint64_t fnv1a64(const std::string& s) { return fnv1a64(s.data(), s.size()); }

// On-disk cubin cache ($MCCBN_CACHE_DIR, default ~/.cache/mccbn_b200, "off" disables it): makes the compilation a
// one-time cost per poset.  File name = hash of (generated source, embedded device headers, target architecture, NVRTC
// version, ptxas options); file = 24-byte header {magic, payload size, FNV-1a of the payload} + cubin.  A file whose
// header or checksum does not match is ignored (and recompiled), the directory is created with mode 0700.
struct JitDiskHeader {
  uint64_t magic, size, hash;
};
static constexpr uint64_t kJitDiskMagic = 0x324e4942554343ull;  // "CCUBIN2"

static std::string jit_cache_path(NvrtcApi& rtc, const std::string& src) {
  std::string dir;
  if (const char* e = std::getenv("MCCBN_CACHE_DIR")) dir = e;
  else if (const char* h = std::getenv("HOME")) dir = std::string(h) + "/.cache/mccbn_b200";
  if (dir.empty() || dir == "off") return "";
  static std::string made;  // create the directory once per process, not once per lookup
  static std::mutex made_mu;  // several device threads of one process (in-process team) look kernels up concurrently
  std::lock_guard<std::mutex> made_lk(made_mu);
  if (made != dir) {  // mkdir -p, private to the user
    for (size_t i = 1; i <= dir.size(); ++i)
      if (i == dir.size() || dir[i] == '/') ::mkdir(dir.substr(0, i).c_str(), 0700);
    made = dir;
  }
  int vmaj = 0, vmin = 0;
  if (rtc.load() && rtc.Version) rtc.Version(&vmaj, &vmin);
  std::ostringstream k;
  k << src << kSrc_kernels_forward_cuh << kSrc_kernels_cond_cuh << kSrc_philox_cuh << kSrc_device_types_cuh
    << "|sm_100a|nvrtc" << vmaj << "." << vmin << "|O" << (std::getenv("MCCBN_JIT_PTXAS_O") ? std::getenv("MCCBN_JIT_PTXAS_O") : "3");
  char name[64];
  std::snprintf(name, sizeof name, "/spec_%016llx.cubin", (unsigned long long)fnv1a64(k.str()));
  return dir + name;
}

static bool jit_disk_read(const std::string& path, std::vector<char>& cubin) {
  FILE* fp = path.empty() ? nullptr : std::fopen(path.c_str(), "rb");
  if (!fp) return false;
  JitDiskHeader h;
  bool ok = std::fread(&h, sizeof h, 1, fp) == 1 && h.magic == kJitDiskMagic && h.size > 0 && h.size < (1ull << 28);
  if (ok) {
    cubin.resize((size_t)h.size);
    ok = std::fread(cubin.data(), 1, cubin.size(), fp) == cubin.size() && fnv1a64(cubin.data(), cubin.size()) == h.hash;
  }
  std::fclose(fp);
  if (!ok) cubin.clear();
  return ok;
}

static void jit_disk_write(const std::string& path, const std::vector<char>& cubin) {
  if (path.empty() || cubin.empty()) return;
  // write-then-rename: safe with several ranks / threads compiling the same kernel
  const std::string tmp = path + "." + std::to_string((long)getpid()) + "." +
                          std::to_string(std::hash<std::thread::id>()(std::this_thread::get_id()));
  if (FILE* fp = std::fopen(tmp.c_str(), "wb")) {
    const JitDiskHeader h{kJitDiskMagic, (uint64_t)cubin.size(), fnv1a64(cubin.data(), cubin.size())};
    const bool okw = std::fwrite(&h, sizeof h, 1, fp) == 1 && std::fwrite(cubin.data(), 1, cubin.size(), fp) == cubin.size();
    std::fclose(fp);
    if (okw) std::rename(tmp.c_str(), path.c_str());
    else std::remove(tmp.c_str());
  }
}

// resident blocks per SM the specialised kernel is compiled for (register cap = 65536 / (128 * minb))
static int jit_default_minb() {
  if (const char* e = std::getenv("MCCBN_JIT_MINB")) return std::max(1, std::min(8, std::atoi(e)));
  return 3;  // best of {3,4,5,6} on B200 (profiles/README.md): fewer, fatter warps win on this pipe-bound kernel
}

// Specialised kernels of ONE context.  Lookup order: memory, disk, then -- depending on the caller's policy -- a
// synchronous compilation, a compilation on a background host thread (the generic kernel runs until it is ready; both
// kernels produce identical records, so the switch-over does not change results), or nothing.
struct JitCache {
  enum Policy { kCachedOnly = 0, kBackground = 1, kCompileNow = 2 };
  NvrtcApi rtc;
  std::map<std::string, JitKernel> kernels;
  std::set<std::string> disk_misses;  // keys looked up without compiling and not found on disk
  struct Pending {
    std::future<JitCompiled> fut;
    std::string path;
    const char* kname = nullptr;
    bool forward = true;
    int p = 0;
  };
  std::map<std::string, Pending> pending;  // compilations running on background threads
  struct Ready {
    JitCompiled cc;
    const char* kname = nullptr;
    bool forward = true;
    int p = 0;
  };
  std::map<std::string, Ready> ready;  // compiled (or read from disk) ahead of use, loaded when first requested
  int minb = jit_default_minb();
  uint64_t tick = 0;
  uint64_t tick_loaded = 0;  // modules loaded so far (captured CUDA graphs are re-captured when it changes)
  size_t max_loaded = 256;  // LRU bound on loaded modules (an annealing run visits thousands of posets)

  JitCache() {
    if (const char* e = std::getenv("MCCBN_JIT_MAX_LOADED")) max_loaded = (size_t)std::max(2, std::atoi(e));
  }
  ~JitCache() {
    for (auto& kv : pending)
      if (kv.second.fut.valid()) kv.second.fut.wait();
    for (auto& kv : kernels)
      if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
  }

  JitKernel* get(const FlatPoset& f, bool times_given, Policy pol = kCompileNow) {
    const int mb = f.p > kFastMaxP ? std::min(minb, 2) : minb;  // 64 event times + 64 sums per thread: 2 resident blocks
    return get_kernel(jit_key(f, times_given, mb), [&] { return generate_forward_source(f, times_given, mb); },
                      "mccbn_fwd_spec", /*forward=*/true, f.p, pol);
  }
  // conditional-proposal kernels (mode = CondMode); they keep their scales in shared memory, no __constant__ symbol
  JitKernel* get_cond(const FlatPoset& f, bool times_given, int mode, Policy pol = kCompileNow) {
    const int mb = cond_minb();
    return get_kernel("c" + std::to_string(mode) + jit_key(f, times_given, mb),
                      [&] { return generate_cond_source(f, times_given, mode, mb); }, "mccbn_cond_spec",
                      /*forward=*/false, f.p, pol);
  }
  static int cond_minb() {
    if (const char* e = std::getenv("MCCBN_JIT_COND_MINB")) return std::max(1, std::min(8, std::atoi(e)));
    return 2;  // cfg 5 backward: 2 / 3 / 4 resident blocks = 3.18 / 3.67 / 15.1 ms (4 blocks: 128 registers, spills in the loop)
  }

  // is the specialised kernel of this poset loaded (no side effects)
  bool has(const FlatPoset& f, bool times_given, int cond_mode /* < 0: forward */) const {
    const int mb = cond_mode >= 0 ? cond_minb() : (f.p > kFastMaxP ? std::min(minb, 2) : minb);
    const std::string key = (cond_mode >= 0 ? "c" + std::to_string(cond_mode) : std::string()) + jit_key(f, times_given, mb);
    auto it = kernels.find(key);
    return it != kernels.end() && it->second.ok;
  }

  template <typename SrcFn>
  JitKernel* get_kernel(const std::string& key, SrcFn make_src, const char* kname, bool forward, int p, Policy pol) {
    auto it = kernels.find(key);
    if (it != kernels.end()) {
      it->second.last_use = ++tick;
      return it->second.ok ? &it->second : nullptr;
    }
    JitCompiled cc;
    bool have = false;
    auto rit = ready.find(key);
    if (rit != ready.end()) {  // compiled ahead of use
      if (pol == kBackground) return nullptr;  // a prefetch: nothing more to do
      cc = std::move(rit->second.cc);
      ready.erase(rit);
      return load(key, cc, true, kname, forward, p);
    }
    auto pit = pending.find(key);
    if (pit != pending.end()) {  // a background compilation of this kernel is under way
      if (pol != kCompileNow && pit->second.fut.wait_for(std::chrono::seconds(0)) != std::future_status::ready)
        return nullptr;
      cc = pit->second.fut.get();  // kCompileNow: wait for it instead of compiling twice
      pending.erase(pit);
      have = !cc.cubin.empty();
    } else {
      if (pol == kCachedOnly && disk_misses.count(key)) return nullptr;  // per-iteration path of small problems
      const std::string src = make_src();
      const std::string path = jit_cache_path(rtc, src);
      have = jit_disk_read(path, cc.cubin);
      if (have && pol == kBackground) {  // a prefetch that hit the disk cache: keep the cubin, load on first use
        Ready r;
        r.cc = std::move(cc);
        r.kname = kname;
        r.forward = forward;
        r.p = p;
        if (ready.size() > 256) ready.clear();
        ready.emplace(key, std::move(r));
        return nullptr;
      }
      if (!have && pol == kCachedOnly) {
        if (disk_misses.size() > 4096) disk_misses.clear();
        disk_misses.insert(key);
        return nullptr;
      }
      if (!have && pol == kBackground) {
        // at most half of the host threads compile at once, at low priority: the thread that drives the device must stay responsive
        static const size_t max_bg = (size_t)std::max(1, (int)std::thread::hardware_concurrency() / 2);
        if (!rtc.load() || pending.size() >= std::min<size_t>(64, max_bg)) return nullptr;
        NvrtcApi* api = &rtc;  // the cache outlives its futures (~JitCache waits for them)
        Pending pd;
        pd.path = path;
        pd.kname = kname;
        pd.forward = forward;
        pd.p = p;
        g_jit_bg_inflight.fetch_add(1, std::memory_order_acq_rel);
        pd.fut = std::async(std::launch::async, [api, src, path] {
          struct InFlight {
            ~InFlight() { g_jit_bg_inflight.fetch_sub(1, std::memory_order_acq_rel); }
          } in_flight;
          setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10);  // this thread only (Linux: per-thread nice value)
          JitCompiled out;
          if (jit_compile(*api, src, out)) jit_disk_write(path, out.cubin);
          else out.cubin.clear();
          return out;
        });
        pending.emplace(key, std::move(pd));
        return nullptr;
      }
      if (!have) {
        have = jit_compile(rtc, src, cc);
        if (have) jit_disk_write(path, cc.cubin);
      }
    }
    return load(key, cc, have, kname, forward, p);
  }

  // collect the background compilations that have finished (cheap; the module itself is loaded when the kernel is first
  // requested for a launch -- most look-ahead candidates of an annealing run are never scored)
  void service() {
    for (auto it = pending.begin(); it != pending.end();) {
      if (it->second.fut.wait_for(std::chrono::seconds(0)) != std::future_status::ready) {
        ++it;
        continue;
      }
      Ready r;
      r.cc = it->second.fut.get();
      r.kname = it->second.kname;
      r.forward = it->second.forward;
      r.p = it->second.p;
      if (ready.size() > 256) ready.clear();
      if (!r.cc.cubin.empty()) ready.emplace(it->first, std::move(r));
      it = pending.erase(it);
    }
  }

  JitKernel* load(const std::string& key, JitCompiled& cc, bool have, const char* kname, bool forward, int p) {
    NvtxRange nvtx("load specialised kernel module");
    const auto t_load0 = std::chrono::steady_clock::now();
    JitKernel jk;
    if (have) {
      cudaError_t e = cudaLibraryLoadData(&jk.lib, cc.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
      if (e == cudaSuccess) e = cudaLibraryGetKernel(&jk.kern, jk.lib, kname);
      size_t bytes = sizeof(FwdScale);
      if (e == cudaSuccess && forward) e = cudaLibraryGetGlobal(&jk.scale_sym, &bytes, jk.lib, "_ZN5mccbn7c_scaleE");
      if (e == cudaSuccess && bytes == sizeof(FwdScale)) {
        jk.ok = true;
        jk.pmax = jit_pmax(p);
        jk.smem = 0;  // neither kernel stages anything in dynamic shared memory (event times and log-uniforms in registers)
        jk.compile_ms = cc.compile_ms;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)jk.kern, kFwdBlock, jk.smem) != cudaSuccess ||
            nb <= 0) {
          cudaGetLastError();
          nb = forward ? minb : cond_minb();  // the launch bound the kernel was compiled for
        }
        if (const char* e2 = std::getenv("MCCBN_JIT_RESIDENT")) nb = std::max(1, std::min(nb, std::atoi(e2)));
        jk.blocks_per_sm = nb;
      } else {
        std::fprintf(stderr, "mccbn_b200: loading the specialised kernel failed (%s); using the generic kernel\n",
                     cudaGetErrorString(e));
        cudaGetLastError();
        if (jk.lib) cudaLibraryUnload(jk.lib);
        jk.lib = nullptr;
      }
    } else {
      std::fprintf(stderr, "mccbn_b200: NVRTC specialisation failed, using the generic kernel\n%s\n", cc.log.c_str());
    }
    if (std::getenv("MCCBN_JIT_VERBOSE"))
      std::fprintf(stderr, "mccbn_b200: specialised kernel %s p=%d: %s (compiled in %.0f ms, module loaded in %.2f ms)\n%s\n",
                   kname, p, jk.ok ? "ready" : "FAILED", cc.compile_ms,
                   std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_load0).count(),
                   cc.log.c_str());
    evict_lru();
    ++tick_loaded;
    jk.last_use = ++tick;
    auto res = kernels.emplace(key, jk);
    return res.first->second.ok ? &res.first->second : nullptr;
  }

  // unload the least recently used modules beyond the bound.  Callers hold a JitKernel* only for the launch they are
  // about to enqueue, and a module is unloaded only after the context's stream has drained (cudaLibraryUnload of a
  // module with kernels in flight is undefined), so eviction synchronises the device first -- it is rare by design.
  void evict_lru() {
    if (kernels.size() < max_loaded) return;
    cudaDeviceSynchronize();
    {  // captured CUDA graphs may hold kernels of the modules about to be unloaded: invalidate them
      int dev = 0;
      cudaGetDevice(&dev);
      ++device_cache().epoch[dev & 63];
    }
    while (kernels.size() >= max_loaded) {
      auto victim = kernels.begin();
      for (auto it = kernels.begin(); it != kernels.end(); ++it)
        if (it->second.last_use < victim->second.last_use) victim = it;
      if (victim->second.lib) cudaLibraryUnload(victim->second.lib);
      kernels.erase(victim);
    }
  }
};

}  // namespace mccbn
