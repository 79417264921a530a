// GroupNorm(32) [+SiLU] and LayerNorm for the NHWC fp32 residual stream -> bf16 GEMM operands.
// Bandwidth-bound kernels: 16-byte vectorised loads, fp32 statistics, warp-shuffle reductions.
//
// Reference semantics:
//   GroupNorm32 (ldm/modules/diffusionmodules/util.py:217-219): nn.GroupNorm(32, C, eps=1e-5) computed in
//   fp32, followed by SiLU in ResBlock.in_layers / out_layers / UNetModel.out (openaimodel.py:205-207,229-231,693-695);
//   Normalize (ldm/modules/attention.py:71-72): nn.GroupNorm(32, C, eps=1e-6), no activation;
//   nn.LayerNorm(C) eps 1e-5 (attention.py:267-269).
// The input may be the channel concat of two tensors (skip connection, openaimodel.py:1019); it is read
// from both sources in place.
#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

constexpr int kGroups = 32;

// ---------------------------------------------------------------------------------------------
// pass 1: partial (sum, sumsq) per (sample, chunk, group)
// grid = (chunks, B); block = 256.  Thread t owns channel quad (t % cq) and strides over pixels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x0, int C0,
                                                       const float* __restrict__ x1, int C1, int HW, int chunks,
                                                       float* __restrict__ partial /* [B, chunks, 32, 2] */) {
  // Deterministic: every thread parks its per-channel partials in shared memory and one thread per group
  // sums them in a fixed order (no floating-point atomics anywhere in the GroupNorm path).
  extern __shared__ float s_part[];  // [slots][8]: 4 channel sums, 4 channel sums of squares
  const int C = C0 + C1;
  const int cq = C >> 2;             // channel quads per pixel
  const int cpg = C / kGroups;
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int pix_per_chunk = (HW + chunks - 1) / chunks;
  const int p_begin = chunk * pix_per_chunk;
  const int p_end = min(HW, p_begin + pix_per_chunk);
  const bool narrow = cq <= 256;
  const int ppb = narrow ? 256 / cq : 1;          // pixels handled per block iteration
  const int slots = narrow ? ppb * cq : cq;

  auto accumulate = [&](int q, int p_first, int p_step, int slot) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    const int c = q * 4;
    const float* src;
    int cs, off;
    if (c < C0) { src = x0; cs = C0; off = c; } else { src = x1; cs = C1; off = c - C0; }
    const float* base = src + static_cast<size_t>(b) * HW * cs + off;
    auto acc = [&](const float4& v) {
      s[0] += v.x; ss[0] += v.x * v.x;
      s[1] += v.y; ss[1] += v.y * v.y;
      s[2] += v.z; ss[2] += v.z * v.z;
      s[3] += v.w; ss[3] += v.w * v.w;
    };
    int p = p_first;
    // four independent 16-byte loads in flight per thread (memory-level parallelism), fixed summation order
    for (; p + 3 * p_step < p_end; p += 4 * p_step) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * cs));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + p_step) * cs));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 2 * p_step) * cs));
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 3 * p_step) * cs));
      acc(v0); acc(v1); acc(v2); acc(v3);
    }
    for (; p < p_end; p += p_step) acc(__ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * cs)));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_part[slot * 8 + e] = s[e];
      s_part[slot * 8 + 4 + e] = ss[e];
    }
  };

  if (narrow) {
    // a fixed thread always sees the same channel quad; slot = pixel lane * cq + quad == threadIdx.x
    if (threadIdx.x < slots) accumulate(threadIdx.x % cq, p_begin + threadIdx.x / cq, ppb, threadIdx.x);
  } else {
    for (int q = threadIdx.x; q < cq; q += 256) accumulate(q, p_begin, 1, q);
  }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    float s = 0.f, ss = 0.f;
    const int lanes = narrow ? ppb : 1;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const int q = c >> 2, e = c & 3;
      for (int l = 0; l < lanes; ++l) {
        s += s_part[(l * cq + q) * 8 + e];
        ss += s_part[(l * cq + q) * 8 + 4 + e];
      }
    }
    float* o = partial + ((static_cast<size_t>(b) * chunks + chunk) * kGroups + g) * 2;
    o[0] = s;
    o[1] = ss;
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: y = silu?((x - mean) * rstd * gamma + beta) -> bf16 (optionally also a raw bf16 copy of x)
// grid = (blocks_per_sample, B); block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x0, int C0,
                                                       const float* __restrict__ x1, int C1, int HW, int chunks,
                                                       const float* __restrict__ partial,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, __nv_bfloat16* __restrict__ y,
                                                       __nv_bfloat16* __restrict__ raw) {
  extern __shared__ float s_ab[];  // [2*C]: scale a[c], shift b[c]
  const int C = C0 + C1;
  const int cpg = C / kGroups;
  const int b = blockIdx.y;
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  if (threadIdx.x < kGroups) {
    float s = 0.f, ss = 0.f;
    const float* pp = partial + (static_cast<size_t>(b) * chunks * kGroups + threadIdx.x) * 2;
    for (int k = 0; k < chunks; ++k) {
      s += pp[static_cast<size_t>(k) * kGroups * 2];
      ss += pp[static_cast<size_t>(k) * kGroups * 2 + 1];
    }
    const float n = static_cast<float>(HW) * cpg;
    const float mean = s / n;
    const float var = fmaxf(ss / n - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    const float a = s_rstd[g] * gamma[c];
    s_ab[c] = a;
    s_ab[C + c] = beta[c] - s_mean[g] * a;
  }
  __syncthreads();

  const int cq = C >> 2;
  const uint32_t total = static_cast<uint32_t>(HW) * cq;   // quads per sample (< 2^31: 32-bit index math)
  const uint32_t per_block = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = blockIdx.x * per_block;
  const uint32_t end = min(total, begin + per_block);
  const float* xb0 = x0 + static_cast<size_t>(b) * HW * C0;
  const float* xb1 = C1 ? x1 + static_cast<size_t>(b) * HW * C1 : nullptr;
  auto src_of = [&](uint32_t i, int& c) -> const float* {
    if (C1 == 0) {  // single source: the input is as dense as the output
      c = static_cast<int>(i % static_cast<uint32_t>(cq)) * 4;
      return xb0 + static_cast<size_t>(i) * 4;
    }
    const uint32_t p = i / static_cast<uint32_t>(cq);
    c = static_cast<int>(i - p * cq) * 4;
    if (c < C0) return xb0 + static_cast<size_t>(p) * C0 + c;
    return xb1 + static_cast<size_t>(p) * C1 + (c - C0);
  };
  auto emit = [&](uint32_t i, int c, const float4& v) {
    float o0 = v.x * s_ab[c] + s_ab[C + c];
    float o1 = v.y * s_ab[c + 1] + s_ab[C + c + 1];
    float o2 = v.z * s_ab[c + 2] + s_ab[C + c + 2];
    float o3 = v.w * s_ab[c + 3] + s_ab[C + c + 3];
    if (silu) {
      o0 = silu_f(o0); o1 = silu_f(o1); o2 = silu_f(o2); o3 = silu_f(o3);
    }
    const size_t oidx = static_cast<size_t>(b) * HW * C + static_cast<size_t>(i) * 4;  // dense concat: same quad index
    uint2 pk;
    pk.x = pack_bf16x2(o0, o1);
    pk.y = pack_bf16x2(o2, o3);
    *reinterpret_cast<uint2*>(y + oidx) = pk;
    if (raw) {
      uint2 rk;
      rk.x = pack_bf16x2(v.x, v.y);
      rk.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(raw + oidx) = rk;
    }
  };
  uint32_t i = begin + threadIdx.x;
  for (; i + 768 < end; i += 1024) {  // four independent 16-byte loads in flight per thread
    int c0, c1, c2, c3;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src_of(i, c0)));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src_of(i + 256, c1)));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(src_of(i + 512, c2)));
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(src_of(i + 768, c3)));
    emit(i, c0, v0); emit(i + 256, c1, v1); emit(i + 512, c2, v2); emit(i + 768, c3, v3);
  }
  for (; i < end; i += 256) {
    int c;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src_of(i, c)));
    emit(i, c, v);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 2048, C % 128 == 0 not required; C % 4 == 0)
// ---------------------------------------------------------------------------------------------
template <int MAXV>  // float4 vectors per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long rows, int C,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        __nv_bfloat16* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      v[i] = __ldg(xr + k);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  s = warp_sum(s);
  const float mean = s / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / C + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* yr = reinterpret_cast<uint2*>(y + row * C);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      const float4 g = __ldg(g4 + k), bb = __ldg(b4 + k);
      uint2 pk;
      pk.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      pk.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      yr[k] = pk;
    }
  }
}

}  // namespace af

using namespace af;

extern "C" size_t af_groupnorm_workspace_bytes(int B) {
  return static_cast<size_t>(B) * AF_GN_MAX_CHUNKS * kGroups * 2 * sizeof(float);
}

extern "C" int af_groupnorm_silu(const float* x0, int C0, const float* x1, int C1, int B, int HW,
                                 const float* gamma, const float* beta, float eps, int silu, void* y_bf16,
                                 void* raw_bf16, float* workspace, cudaStream_t stream) {
  AF_CHECK_ARG(x0 && gamma && beta && y_bf16 && workspace, "af_groupnorm_silu: null pointer");
  const int C = C0 + C1;
  AF_CHECK_ARG(B > 0 && HW > 0 && C0 > 0 && C1 >= 0, "af_groupnorm_silu: bad sizes");
  AF_CHECK_ARG(C % 32 == 0 && C0 % 4 == 0 && C1 % 4 == 0, "af_groupnorm_silu: C0=%d C1=%d need C%%32==0, quads", C0, C1);
  AF_CHECK_ARG(C1 == 0 || x1 != nullptr, "af_groupnorm_silu: x1 null with C1=%d", C1);
  AF_CHECK_ARG(C <= 5120, "af_groupnorm_silu: C=%d too large", C);
  // enough chunks to fill the machine, few enough that pass 2 sums them cheaply
  int chunks = (8 * num_sms() + B - 1) / B;
  if (chunks > HW / 16) chunks = HW / 16;   // keep >= 16 pixels per block
  if (chunks > AF_GN_MAX_CHUNKS) chunks = AF_GN_MAX_CHUNKS;
  if (chunks > HW) chunks = HW;
  if (chunks < 1) chunks = 1;
  const int cq = C / 4;
  const size_t stats_smem = static_cast<size_t>(cq <= 256 ? (256 / cq) * cq : cq) * 8 * sizeof(float);
  gn_stats_kernel<<<dim3(chunks, B), 256, stats_smem, stream>>>(x0, C0, x1, C1, HW, chunks, workspace);
  AF_LAUNCH_CHECK("gn_stats_kernel");
  const size_t total = static_cast<size_t>(HW) * (C / 4);
  int blocks = static_cast<int>((total + 256 * 8 - 1) / (256 * 8));
  const int cap = (16 * num_sms() + B - 1) / B;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  gn_apply_kernel<<<dim3(blocks, B), 256, 2 * C * sizeof(float), stream>>>(
      x0, C0, x1, C1, HW, chunks, workspace, gamma, beta, eps, silu, static_cast<__nv_bfloat16*>(y_bf16),
      static_cast<__nv_bfloat16*>(raw_bf16));
  AF_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}

extern "C" int af_layernorm(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                            void* y_bf16, cudaStream_t stream) {
  AF_CHECK_ARG(x && gamma && beta && y_bf16, "af_layernorm: null pointer");
  AF_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && C <= 2048, "af_layernorm: rows=%lld C=%d (need C%%4==0, C<=2048)", rows, C);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  const int nvec = C / 4;
  if (nvec <= 3 * 32) {
    layernorm_kernel<3><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, static_cast<__nv_bfloat16*>(y_bf16));
  } else if (nvec <= 5 * 32) {
    layernorm_kernel<5><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, static_cast<__nv_bfloat16*>(y_bf16));
  } else if (nvec <= 10 * 32) {
    layernorm_kernel<10><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, static_cast<__nv_bfloat16*>(y_bf16));
  } else {
    layernorm_kernel<16><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, static_cast<__nv_bfloat16*>(y_bf16));
  }
  AF_LAUNCH_CHECK("layernorm_kernel");
  return 0;
}
