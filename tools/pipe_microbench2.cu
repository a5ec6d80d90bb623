// pipe_microbench2.cu -- dev tool: pairwise co-issue of IMAD.WIDE with the other instruction classes of the E-step
// kernel.  Each loop body holds NW IMAD.WIDE and NX instructions of class X on independent registers.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_microbench2 tools/pipe_microbench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define WIDE(a, b) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(a) : "r"(b))
#define LOP(a, b, c) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c))
#define FMA(a, b, c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c))
#define LG2(a) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a))
#define EX2(a) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a))
#define IADD(a, b) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b))
#define FADD(a, b) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b))
#define FMNMX(a, b) asm volatile("max.f32 %0, %0, %1;" : "+f"(a) : "f"(b))
#define SHF(a, b) asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(a) : "r"(b))
#define FSET(x, f, g) asm volatile("set.le.u32.f32 %0, %1, %2;" : "=r"(x) : "f"(f), "f"(g))
#define LDS(f, a) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(a))
#define STS(a, f) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(f))
#define LEAHI(a, b) asm volatile("{ .reg .u32 t; shr.u32 t, %1, 9; add.u32 %0, t, 0x3f800000; }" : "=r"(a) : "r"(b))

enum { X_NONE, X_LOP, X_FMA, X_LG2, X_IADD, X_FADD, X_FMNMX, X_SHF, X_FSET, X_LDS, X_STS, X_LEA, X_N };
static const char* kX[X_N] = {"-", "LOP3", "FFMA", "MUFU.LG2", "IADD3", "FADD", "FMNMX", "SHF", "FSET", "LDS", "STS", "LEA.HI"};

template <int NW, int XT, int NX>
__global__ void __launch_bounds__(128) k(uint32_t* out, int iters, uint32_t seed) {
  __shared__ float sm[1024];
  uint32_t x[8], y[8];
  float f[8];
  uint64_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i, y[i] = x[i] * 3, f[i] = 1.5f + i + seed, w[i] = 0;
  sm[threadIdx.x] = 1.f;
  const uint32_t c1 = seed * 3 + 1, c2 = seed * 7 + 5;
  const float g1 = 0.999f + seed, g2 = 0.001f * seed;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // interleave: spread the NW wides evenly among the NX others
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < NW) { WIDE(w[i], x[i]); x[i] = (uint32_t)(w[i] >> 32) ^ (uint32_t)w[i]; }  // + one LOP3-class xor (like Philox)
        if (i < NX) {
          if (XT == X_LOP) LOP(y[i], c1, c2);
          if (XT == X_FMA) FMA(f[i], g1, g2);
          if (XT == X_LG2) LG2(f[i]);
          if (XT == X_IADD) IADD(y[i], c1);
          if (XT == X_FADD) FADD(f[i], g1);
          if (XT == X_FMNMX) FMNMX(f[i], g1);
          if (XT == X_SHF) SHF(y[i], c1);
          if (XT == X_FSET) FSET(y[i], f[i], g1);
          if (XT == X_LDS) LDS(f[i], sa + i * 512);
          if (XT == X_STS) STS(sa + i * 512, f[i]);
          if (XT == X_LEA) LEAHI(y[i], y[i]);
        }
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i] ^ __float_as_uint(f[i]) ^ (uint32_t)w[i];
  if (s == 0x12345678u) out[0] = s;
}

template <int NW, int XT, int NX>
void run(int sms, double ghz, uint32_t* d) {
  const int bps = 4, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<NW, XT, NX><<<sms * bps, 128>>>(d, 100, 1);
  cudaEventRecord(e0);
  k<NW, XT, NX><<<sms * bps, 128>>>(d, iters, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clk = ms * 1e-3 * ghz * 1e9;
  // cycles per loop-body pass per scheduler (4 warps per scheduler, 4 unrolled passes per iteration)
  printf("%d x (IMAD.WIDE + xor) + %d x %-9s : %6.2f cycles per pass per scheduler-warp\n", NW, NX, kX[XT],
         clk / ((double)iters * 4 * bps));
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  const int sms = pr.multiProcessorCount;
  uint32_t* d;
  cudaMalloc(&d, 4);
  printf("%s: cycles per pass (4 resident warps per scheduler; a pass = the listed instructions of ONE warp)\n", pr.name);
  run<4, X_NONE, 0>(sms, ghz, d);
  run<0, X_LOP, 8>(sms, ghz, d);  run<4, X_LOP, 4>(sms, ghz, d);  run<4, X_LOP, 8>(sms, ghz, d);
  run<0, X_FMA, 8>(sms, ghz, d);  run<4, X_FMA, 4>(sms, ghz, d);  run<4, X_FMA, 8>(sms, ghz, d);
  run<0, X_LG2, 2>(sms, ghz, d);  run<4, X_LG2, 1>(sms, ghz, d);  run<4, X_LG2, 2>(sms, ghz, d);
  run<0, X_IADD, 8>(sms, ghz, d); run<4, X_IADD, 4>(sms, ghz, d); run<4, X_IADD, 8>(sms, ghz, d);
  run<0, X_FADD, 8>(sms, ghz, d); run<4, X_FADD, 4>(sms, ghz, d); run<4, X_FADD, 8>(sms, ghz, d);
  run<0, X_FMNMX, 8>(sms, ghz, d); run<4, X_FMNMX, 4>(sms, ghz, d); run<4, X_FMNMX, 8>(sms, ghz, d);
  run<0, X_SHF, 8>(sms, ghz, d);  run<4, X_SHF, 4>(sms, ghz, d);  run<4, X_SHF, 8>(sms, ghz, d);
  run<0, X_FSET, 8>(sms, ghz, d); run<4, X_FSET, 4>(sms, ghz, d); run<4, X_FSET, 8>(sms, ghz, d);
  run<0, X_LDS, 8>(sms, ghz, d);  run<4, X_LDS, 4>(sms, ghz, d);
  run<0, X_STS, 8>(sms, ghz, d);  run<4, X_STS, 4>(sms, ghz, d);
  run<0, X_LEA, 8>(sms, ghz, d);  run<4, X_LEA, 4>(sms, ghz, d);  run<4, X_LEA, 8>(sms, ghz, d);
  return 0;
}
