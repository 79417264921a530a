// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition, alone and mixed with an FMA-pipe exp2 polynomial.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/mufu scripts/ubench/mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^x for x in [-126, 8]: round-to-nearest split, degree-3 minimax on [-0.5, 0.5], exponent add on the bit pattern
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;             // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float j = t - 12582912.0f;
  const float f = x - j;
  float p = fmaf(f, 0.05550410866f, 0.24022650695f);
  p = fmaf(p, f, 0.69314718056f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>   // 0: MUFU only, 1: poly only, 2: 1 poly per 1 MUFU, 3: 1 poly per 2 MUFU, 4: 1 poly per 3 MUFU
__global__ void k(float* out, int iters, float seed) {
  float a[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) a[i] = seed * (i + 1) * 0.01f - threadIdx.x * 1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      bool poly = MODE == 1 || (MODE == 2 && (i & 1)) || (MODE == 3 && (i % 3 == 2)) || (MODE == 4 && (i % 4 == 3));
      a[i] = (poly ? ex2_poly(a[i]) : ex2(a[i])) - 1.5f;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) s += a[i];
  if (s == 12345.f) out[0] = s;
}

template <int MODE>
void run(int warps_per_sm, const char* name) {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148, warps_per_sm * 32>>>(d, 100, 0.3f);
  cudaEventRecord(e0);
  k<MODE><<<148, warps_per_sm * 32>>>(d, iters, 0.3f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double n = 148.0 * warps_per_sm * 32 * 12.0 * iters;
  double per_clk_sm = n / (ms * 1e-3) / 148.0 / (clk_khz * 1e3);
  printf("%-28s warps/SM %2d: %8.3f ms  %7.2f exp/clk/SM (at %d MHz nominal)  %.2f Texp/s\n", name, warps_per_sm, ms, per_clk_sm, clk_khz / 1000, n / ms / 1e9);
  cudaFree(d);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>(w, "MUFU.EX2 only");
    run<1>(w, "poly only");
    run<2>(w, "1 poly : 1 MUFU");
    run<3>(w, "1 poly : 2 MUFU");
    run<4>(w, "1 poly : 3 MUFU");
  }
  return 0;
}
