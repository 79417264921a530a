// Micro-benchmark: ex2 throughput per SM for the three ways a softmax thread can exponentiate a pair of scores
//   f32    : 2 x FADD, 2 x MUFU.EX2 (f32), 1 x F2FP (pack to bf16x2)
//   f16x2  : 2 x FADD, 1 x F2FP (pack to f16x2), 1 x ex2.approx.f16x2        -> P already packed (f16)
//   bf16x2 : 2 x FADD, 1 x F2FP (pack to bf16x2), 1 x ex2.approx.ftz.bf16x2  -> P already packed (bf16)
// and the bare instruction rates.  Answers: does the packed form retire two exponentials per MUFU slot?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/mufu2 scripts/ubench/mufu2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t pack_b2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

constexpr int NP = 16;   // independent score pairs per thread and iteration

// MODE 0 f32 mix, 1 f16x2 mix, 2 bf16x2 mix, 3 bare f32 MUFU, 4 bare f16x2, 5 bare bf16x2
template <int MODE>
__global__ void k(uint32_t* out, int iters, float seed) {
  float s[2 * NP];
#pragma unroll
  for (int i = 0; i < 2 * NP; ++i) s[i] = -seed * (i + 1) * 0.37f - threadIdx.x * 1e-3f;
  uint32_t acc = 0;
  float m = seed;
  uint32_t hx[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) hx[i] = pack_h2(s[2 * i], s[2 * i + 1]);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < NP; ++i) acc ^= pack_b2(ex2f(s[2 * i] - m), ex2f(s[2 * i + 1] - m));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < NP; ++i) acc ^= ex2_h2(pack_h2(s[2 * i] - m, s[2 * i + 1] - m));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < NP; ++i) acc ^= ex2_b2(pack_b2(s[2 * i] - m, s[2 * i + 1] - m));
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 2 * NP; ++i) s[i] = ex2f(s[i]);
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < NP; ++i) hx[i] = ex2_h2(hx[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NP; ++i) hx[i] = ex2_b2(hx[i]);
    }
    m += 1e-6f;   // keeps the subtraction inside the loop
  }
  float fs = 0;
#pragma unroll
  for (int i = 0; i < 2 * NP; ++i) fs += s[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) acc ^= hx[i];
  if (acc == 0x12345678u && fs == 1.0f) out[0] = acc;
}

template <int MODE>
void run(int warps_per_sm, const char* name) {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148, warps_per_sm * 32>>>(d, 100, 0.3f);
  cudaEventRecord(e0);
  k<MODE><<<148, warps_per_sm * 32>>>(d, iters, 0.3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double n = 148.0 * warps_per_sm * 32 * 2.0 * NP * iters;    // exponentials
  double per_clk_sm = n / (ms * 1e-3) / 148.0 / (clk_khz * 1e3);
  printf("%-28s warps/SM %2d: %8.3f ms  %7.2f exp/clk/SM (at %d MHz nominal)  %.2f Texp/s\n", name, warps_per_sm, ms,
         per_clk_sm, clk_khz / 1000, n / ms / 1e9);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>(w, "mix f32 (2 MUFU / pair)");
    run<1>(w, "mix f16x2 (1 MUFU / pair)");
    run<2>(w, "mix bf16x2 (1 MUFU / pair)");
    run<3>(w, "bare MUFU.EX2 f32");
    run<4>(w, "bare ex2 f16x2");
    run<5>(w, "bare ex2 bf16x2");
  }
  return 0;
}
