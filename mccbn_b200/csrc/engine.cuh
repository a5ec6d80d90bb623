// engine.cuh -- host-side engine of libmccbn_b200: context (stream, NCCL), device-resident problem, EM driver.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

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

static const char* kMsgNotAcyclic = "Poset does not correspond to an acyclic graph";
static const char* kMsgZeroWeight = "ERROR: all samples have weight 0. Consider increasing L";

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

template <typename T>
struct DevBuf {  // grow-only device buffer
  T* ptr = nullptr;
  size_t cap = 0;
  T* ensure(size_t n) {
    if (n > cap) {
      if (ptr) cudaFree(ptr);
      ptr = nullptr;
      MCCBN_CUDA(cudaMalloc(&ptr, n * sizeof(T)));
      cap = n;
    }
    return ptr;
  }
  void release() {
    if (ptr) cudaFree(ptr);
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
  mccbn::JitCache* jit = nullptr;  // poset-specialised kernels (jit.cuh)
  int jit_mode = -1;               // -1 auto (large problems only), 0 off, 1 always
};
