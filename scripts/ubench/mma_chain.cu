// Micro-benchmark: cycles per tcgen05.mma (bf16, M=128, K=16, cta_group::1) as a function of N, of the operand
// source of A (shared memory / tensor memory) and of how many INDEPENDENT accumulators the issue order rotates over.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -o scripts/ubench/mma_chain scripts/ubench/mma_chain.cu
#include "../../adaprompt_b200/csrc/common.cuh"
namespace af { void set_error(const char*, ...) {} }
using namespace af;

// one CTA per SM; thread 0 issues `iters` rounds of `chain` MMAs; accumulator of MMA i = (i % nacc) * accstride
template <int N, bool TS, int CHAIN, int NACC>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 16384);
    const int accstride = N < 32 ? 32 : N;
    const uint64_t ad0 = umma_desc_sw128(a_addr), bd0 = umma_desc_sw128(b_addr);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < CHAIN; ++i) {
        const int kk = i & 3;
        const uint32_t d = tm + (i % NACC) * accstride;
        if (TS) tc_mma_ts(d, tm + 448 + kk * 8, bd0 + 2 * kk, idesc, 1u);
        else tc_mma_ss(d, ad0 + 2 * kk, bd0 + 2 * kk, idesc, 1u);
      }
      tc_commit(&bar);
      mbar_wait(&bar, ph); ph ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, bool TS, int CHAIN, int NACC>
void run() {
  const int chain = CHAIN, nacc = NACC;
  long long* d; cudaMalloc(&d, 8);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k<N, TS, CHAIN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 200;
  k<N, TS, CHAIN, NACC><<<148, 128, smem>>>(d, 10);
  k<N, TS, CHAIN, NACC><<<148, 128, smem>>>(d, iters);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("N=%3d A=%s chain=%3d accumulators=%d : %7.1f clk/MMA (floor %3d)  %s\n", N, TS ? "tmem" : "smem", chain, nacc,
         double(h) / iters / chain, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
template <int CHAIN>
void all() {
  run<48, false, CHAIN, 1>(); run<48, false, CHAIN, 2>(); run<48, false, CHAIN, 4>();
  run<48, true, CHAIN, 1>(); run<48, true, CHAIN, 2>();
  run<64, false, CHAIN, 1>(); run<64, false, CHAIN, 2>();
  run<128, false, CHAIN, 1>(); run<128, false, CHAIN, 2>();
  run<160, false, CHAIN, 1>(); run<160, false, CHAIN, 2>();
  run<256, false, CHAIN, 1>(); run<256, false, CHAIN, 2>();
}
int main() {
  all<3>();
  all<8>();
  all<64>();
  return 0;
}
