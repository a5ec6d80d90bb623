// engine.cuh -- host-side engine of libmccbn_b200: context (stream, NCCL), device-resident problem, EM driver.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: a no-op unless a profiler (nsys / ncu --nvtx) is attached

#include "../../include/mccbn_b200.h"
#include "device_types.cuh"
#include "poset.hpp"

namespace mccbn {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MCCBN_CUDA(expr)                                                                                   \
  do {                                                                                                     \
    cudaError_t e__ = (expr);                                                                              \
    if (e__ != cudaSuccess)                                                                                \
      throw ::mccbn::Error(MCCBN_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + \
                                               __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
  } while (0)

// NVTX range (tracing): `nsys profile --trace=cuda,nvtx` / `ncu --nvtx --nvtx-include "mccbn/"` show
// the phases of a call -- upload, EM batches, E-step enqueues, annealing steps -- on the timeline.  Domain "mccbn".
struct NvtxRange {
  static nvtxDomainHandle_t domain() {
    static nvtxDomainHandle_t d = nvtxDomainCreateA("mccbn");
    return d;
  }
  explicit NvtxRange(const char* name) {
    nvtxEventAttributes_t a = {};
    a.version = NVTX_VERSION;
    a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
    a.messageType = NVTX_MESSAGE_TYPE_ASCII;
    a.message.ascii = name;
    nvtxDomainRangePushEx(domain(), &a);
  }
  ~NvtxRange() { nvtxDomainRangePop(domain()); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

static const char* kMsgNotAcyclic = "Poset does not correspond to an acyclic graph";
static const char* kMsgZeroWeight = "ERROR: all samples have weight 0. Consider increasing L";
static const char* kMsgCommTimeout = "peer-memory statistic exchange timed out: a rank did not publish its statistics";

// ---- NCCL through dlopen (no link-time dependency; reuses the libnccl.so.2 torch already loaded, if any) -----
struct NcclApi {
  typedef struct ncclComm* comm_t;
  struct unique_id {
    char internal[MCCBN_NCCL_ID_BYTES];
  };
  void* lib = nullptr;
  int (*GetUniqueId)(unique_id*) = nullptr;
  int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  static constexpr int kDouble = 8, kSum = 0;

  void load() {
    if (lib) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) throw Error(MCCBN_ERR_CUDA, std::string("cannot load NCCL: ") + dlerror());
    auto sym = [&](const char* s) {
      void* f = dlsym(lib, s);
      if (!f) throw Error(MCCBN_ERR_CUDA, std::string("NCCL symbol missing: ") + s);
      return f;
    };
    GetUniqueId = (int (*)(unique_id*))sym("ncclGetUniqueId");
    CommInitRank = (int (*)(comm_t*, int, unique_id, int))sym("ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, comm_t, cudaStream_t))sym("ncclAllReduce");
    CommDestroy = (int (*)(comm_t))sym("ncclCommDestroy");
    GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  }
  void check(int rc, const char* what) {
    if (rc != 0) throw Error(MCCBN_ERR_CUDA, std::string("NCCL error in ") + what + ": " + GetErrorString(rc));
  }
};

// Device-memory cache.  The reference-facing entry points (mccbn_obs_log_likelihood, mccbn_importance_weight, ...)
// are called once per E-step by the R wrappers and build a fresh problem each time; cudaMalloc / cudaFree of its
// ~10 buffers (tens of MB) would cost more than the upload.  Freed blocks are therefore kept per device and handed
// out again (best fit within 2x).  Every entry point synchronises its stream before its buffers are destroyed, so a
// cached block is never in use by in-flight work.  MCCBN_DEVICE_CACHE_MB caps the cache (default 4096, 0 = off).
struct DeviceCache {
  std::mutex mu;
  std::atomic<unsigned long long> epoch[64] = {};  // per device, bumped by every allocation: captured CUDA graphs hold raw pointers
  unsigned long long epoch_of_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return epoch[dev & 63].load();
  }
  std::multimap<std::pair<int, size_t>, void*> free_blocks;  // (device, bytes) -> block
  std::map<void*, std::pair<int, size_t>> live;               // block -> (device, bytes)
  size_t cached_bytes = 0, cap_bytes = (size_t)4096 << 20;
  DeviceCache() {
    if (const char* e = std::getenv("MCCBN_DEVICE_CACHE_MB")) cap_bytes = (size_t)std::max(0L, std::atol(e)) << 20;
  }
  static size_t round_up(size_t n) {
    const size_t g = n <= (1u << 20) ? 512 : (size_t)1 << 20;  // 512 B below 1 MiB, 1 MiB above
    return (n + g - 1) / g * g;
  }
  void* alloc(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = round_up(std::max<size_t>(bytes, 1));
    {
      std::lock_guard<std::mutex> lk(mu);
      auto it = free_blocks.lower_bound({dev, want});
      if (it != free_blocks.end() && it->first.first == dev && it->first.second <= 2 * want) {
        void* p = it->second;
        ++epoch[dev & 63];
        cached_bytes -= it->first.second;
        live[p] = it->first;
        free_blocks.erase(it);
        return p;
      }
    }
    void* p = nullptr;
    ++epoch[dev & 63];
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
      cudaGetLastError();
      trim();
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) throw Error(MCCBN_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lk(mu);
    live[p] = {dev, want};
    return p;
  }
  void free(void* p) {
    if (!p) return;
    std::unique_lock<std::mutex> lk(mu);
    auto it = live.find(p);
    if (it == live.end()) {
      lk.unlock();
      cudaFree(p);
      return;
    }
    const std::pair<int, size_t> key = it->second;
    live.erase(it);
    if (cached_bytes + key.second <= cap_bytes) {
      free_blocks.emplace(key, p);
      cached_bytes += key.second;
      return;
    }
    lk.unlock();
    cudaFree(p);
  }
  void trim() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
    cached_bytes = 0;
  }
};
inline DeviceCache& device_cache() {
  static DeviceCache* c = new DeviceCache();  // intentionally leaked: outlives every static destructor
  return *c;
}

template <typename T>
struct DevBuf {  // grow-only device buffer
  T* ptr = nullptr;
  size_t cap = 0;
  T* ensure(size_t n) {
    if (n > cap) {
      // growing a buffer that in-flight kernels may still use: wait before its block goes back to the cache (rare:
      // buffers reach their final size in the first iteration)
      if (ptr) cudaDeviceSynchronize();
      if (ptr) device_cache().free(ptr);
      ptr = nullptr;
      cap = 0;
      ptr = static_cast<T*>(device_cache().alloc(n * sizeof(T)));
      cap = n;
    }
    return ptr;
  }
  void release() {
    if (ptr) device_cache().free(ptr);
    ptr = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
};

}  // namespace mccbn

namespace mccbn {
struct JitCache;
}

struct mccbn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int sm_count = 148;
  mccbn::NcclApi nccl;
  mccbn::NcclApi::comm_t comm = nullptr;
  int rank = 0, world = 1;
  uint64_t launches = 0;
  uint64_t jit_launches = 0;  // launches of NVRTC poset-specialised kernels (mccbn_jit_launches)
  // peer-memory statistic exchange (device_types.cuh: Mailbox); preferred over NCCL when initialised
  mccbn::Mailbox* mailbox = nullptr;
  mccbn::Mailbox* peer_box[mccbn::kMaxWorld] = {};
  bool peer_ready = false;
  bool mailbox_host = false;  // member of an in-process team: the mailboxes are pinned host memory owned by the team
  unsigned long long xchg_seq = 0;
  mccbn::JitCache* jit = nullptr;  // poset-specialised kernels (jit.cuh)
  int jit_mode = -1;               // -1 auto (large problems only), 0 off, 1 always
  // Pinned staging ring for the small host -> device parameter blocks (poset + model of a candidate): copies from it
  // are truly asynchronous, so set_model does not have to drain the stream before its stack objects go out of scope.
  // A slot is reused only after the event recorded behind its copies has completed.
  static constexpr int kStageSlots = 32;
  static constexpr size_t kStageBytes = 8192;
  unsigned char* stage = nullptr;
  cudaEvent_t stage_ev[kStageSlots] = {};
  int stage_next = 0;
  unsigned char* stage_slot(int& slot) {
    if (!stage) {
      if (cudaHostAlloc((void**)&stage, kStageSlots * kStageBytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        stage = nullptr;
        return nullptr;
      }
      for (int i = 0; i < kStageSlots; ++i) cudaEventCreateWithFlags(&stage_ev[i], cudaEventDisableTiming);
    }
    slot = stage_next;
    stage_next = (stage_next + 1) % kStageSlots;
    cudaEventSynchronize(stage_ev[slot]);  // completed long ago unless 32 parameter uploads are in flight
    return stage + (size_t)slot * kStageBytes;
  }
  void stage_release() {
    if (!stage) return;
    for (int i = 0; i < kStageSlots; ++i) cudaEventDestroy(stage_ev[i]);
    cudaFreeHost(stage);
    stage = nullptr;
  }
};
