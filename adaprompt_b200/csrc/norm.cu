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
// Statistics come in ONE format everywhere: per-channel partial sums
//     stats[(b * slots + slot) * C + c] = (sum, sum of squares) over the pixels of `slot`
// written either by the epilogue of the GEMM / conv that produced the tensor (gemm_tc.cu, one slot per 32 output
// rows) or by gn_stats_kernel below (tensors that no tensor-core kernel produced).  gn_finalize_kernel folds
// them into (mean, rstd) per (sample, group) in a fixed order - no floating-point atomics anywhere, results are
// bit-reproducible run to run - and gn_apply_kernel makes the single normalising pass over the data.
// ---------------------------------------------------------------------------------------------
// grid = (slots, B); block = 256.  Thread t owns channel quad (t % cq) and strides over the slot's pixels.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int C, int HW, int slots,
                                                       float* __restrict__ stats /* [B, slots, C, 2] */) {
  extern __shared__ float s_part[];  // [lanes][cq][8]: 4 channel sums, 4 channel sums of squares
  const int cq = C >> 2;             // channel quads per pixel
  const int b = blockIdx.y;
  const int slot = blockIdx.x;
  const int pix_per_slot = (HW + slots - 1) / slots;
  const int p_begin = slot * pix_per_slot;
  const int p_end = min(HW, p_begin + pix_per_slot);
  const bool narrow = cq <= 256;
  const int ppb = narrow ? 256 / cq : 1;          // pixel lanes per block iteration

  auto accumulate = [&](int q, int p_first, int p_step, int dst) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    const float* base = x + static_cast<size_t>(b) * HW * C + q * 4;
    auto acc = [&](const float4& v) {
      s[0] += v.x; ss[0] = fmaf(v.x, v.x, ss[0]);
      s[1] += v.y; ss[1] = fmaf(v.y, v.y, ss[1]);
      s[2] += v.z; ss[2] = fmaf(v.z, v.z, ss[2]);
      s[3] += v.w; ss[3] = fmaf(v.w, v.w, ss[3]);
    };
    int p = p_first;
    // four independent 16-byte loads in flight per thread (memory-level parallelism), fixed summation order
    for (; p + 3 * p_step < p_end; p += 4 * p_step) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * C));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + p_step) * C));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 2 * p_step) * C));
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p + 3 * p_step) * C));
      acc(v0); acc(v1); acc(v2); acc(v3);
    }
    for (; p < p_end; p += p_step) acc(__ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(p) * C)));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_part[dst * 8 + e] = s[e];
      s_part[dst * 8 + 4 + e] = ss[e];
    }
  };

  if (narrow) {
    if (threadIdx.x < ppb * cq) accumulate(threadIdx.x % cq, p_begin + threadIdx.x / cq, ppb, threadIdx.x);
  } else {
    for (int q = threadIdx.x; q < cq; q += 256) accumulate(q, p_begin, 1, q);
  }
  __syncthreads();
  float* out = stats + (static_cast<size_t>(b) * slots + slot) * C * 2;
  const int lanes = narrow ? ppb : 1;
  for (int q = threadIdx.x; q < cq; q += 256) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < lanes; ++l) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[e] += s_part[(l * cq + q) * 8 + e];
        ss[e] += s_part[(l * cq + q) * 8 + 4 + e];
      }
    }
    float4* o = reinterpret_cast<float4*>(out + q * 8);
    o[0] = make_float4(s[0], ss[0], s[1], ss[1]);
    o[1] = make_float4(s[2], ss[2], s[3], ss[3]);
  }
}

// grid = (32 groups, B); block = 256.  mean_rstd[b][g] = (mean, rstd) over the group's channels of [x0 | x1].
// Latency-bound: every thread keeps four independent 8-byte loads in flight and the block reduces with warp shuffles
// (fixed order: bit-reproducible).
__device__ __forceinline__ void gn_fold_source(const float2* __restrict__ base, int Cs, int lo, int w, int slots,
                                               float& s, float& ss) {
  const int n = w * slots;
  const float2* col = base + lo;
  int i = threadIdx.x;
  for (; i + 3 * 256 < n; i += 4 * 256) {
    int sl[4], c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sl[u] = (i + u * 256) / w;
      c[u] = (i + u * 256) - sl[u] * w;
    }
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(col + static_cast<size_t>(sl[u]) * Cs + c[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s += v[u].x;
      ss += v[u].y;
    }
  }
  for (; i < n; i += 256) {
    const int sl = i / w, c = i - sl * w;
    const float2 v = __ldg(col + static_cast<size_t>(sl) * Cs + c);
    s += v.x;
    ss += v.y;
  }
}
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ st0, int C0, int slots0,
                                                          const float* __restrict__ st1, int C1, int slots1, int HW,
                                                          float eps, float* __restrict__ mean_rstd) {
  __shared__ float red[2][8];
  const int C = C0 + C1;
  const int cpg = C / kGroups;
  const int g = blockIdx.x, b = blockIdx.y;
  const int c_lo = g * cpg, c_hi = c_lo + cpg;
  float s = 0.f, ss = 0.f;
  {
    const int lo = min(c_lo, C0), hi = min(c_hi, C0);
    if (hi > lo)
      gn_fold_source(reinterpret_cast<const float2*>(st0) + static_cast<size_t>(b) * slots0 * C0, C0, lo, hi - lo, slots0, s, ss);
  }
  if (C1 > 0) {
    const int lo = max(c_lo, C0) - C0, hi = max(c_hi, C0) - C0;
    if (hi > lo)
      gn_fold_source(reinterpret_cast<const float2*>(st1) + static_cast<size_t>(b) * slots1 * C1, C1, lo, hi - lo, slots1, s, ss);
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tss = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      ts += red[0][w];
      tss += red[1][w];
    }
    const float n = static_cast<float>(HW) * cpg;
    const float mean = ts / n;
    const float var = fmaxf(tss / n - mean * mean, 0.f);
    mean_rstd[(static_cast<size_t>(b) * kGroups + g) * 2] = mean;
    mean_rstd[(static_cast<size_t>(b) * kGroups + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// ---------------------------------------------------------------------------------------------
// y = silu?((x - mean) * rstd * gamma + beta) -> bf16 (optionally also a raw bf16 copy of x)
// grid = (blocks_per_sample, B); block = 256
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x0, int C0,
                                                       const float* __restrict__ x1, int C1, int HW,
                                                       const float* __restrict__ mean_rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int silu, __nv_bfloat16* __restrict__ y,
                                                       __nv_bfloat16* __restrict__ raw) {
  extern __shared__ float s_ab[];  // [2*C]: scale a[c], shift b[c]
  const int C = C0 + C1;
  const int cpg = C / kGroups;
  const int b = blockIdx.y;
  const float2* mr = reinterpret_cast<const float2*>(mean_rstd) + static_cast<size_t>(b) * kGroups;
  auto prologue = [&]() {   // per-channel scale / shift of this sample; runs UNDER the first streaming loads
    for (int c = threadIdx.x; c < C; c += 256) {
      const float2 m = __ldg(mr + c / cpg);
      const float a = m.y * gamma[c];
      s_ab[c] = a;
      s_ab[C + c] = beta[c] - m.x * a;
    }
    __syncthreads();
  };

  const int cq = C >> 2;
  const uint32_t total = static_cast<uint32_t>(HW) * cq;   // quads per sample (< 2^31: 32-bit index math)
  const uint32_t per_block = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t begin = blockIdx.x * per_block;
  const uint32_t end = min(total, begin + per_block);
  const float* xb0 = x0 + static_cast<size_t>(b) * HW * C0;
  const float* xb1 = C1 ? x1 + static_cast<size_t>(b) * HW * C1 : nullptr;
  // channel quad / pixel of a flat quad index are tracked incrementally (the thread advances by 256 quads per step):
  // no integer division by the run-time channel count in the loop
  const uint32_t ucq = static_cast<uint32_t>(cq);
  const uint32_t step_q = 256u % ucq, step_p = 256u / ucq;
  auto src_at = [&](uint32_t i, uint32_t p, uint32_t q) -> const float* {   // q = i % cq, p = i / cq
    const int c = static_cast<int>(q) * 4;
    if (C1 == 0) return xb0 + static_cast<size_t>(i) * 4;            // single source: as dense as the output
    if (c < C0) return xb0 + static_cast<size_t>(p) * C0 + c;
    return xb1 + static_cast<size_t>(p) * C1 + (c - C0);
  };
  auto advance = [&](uint32_t& p, uint32_t& q) {
    q += step_q;
    p += step_p;
    if (q >= ucq) {
      q -= ucq;
      ++p;
    }
  };
  auto emit = [&](uint32_t i, int c, const float4& v) {
    const float4 a = *reinterpret_cast<const float4*>(s_ab + c);
    const float4 sh = *reinterpret_cast<const float4*>(s_ab + C + c);
    float o0 = fmaf(v.x, a.x, sh.x);
    float o1 = fmaf(v.y, a.y, sh.y);
    float o2 = fmaf(v.z, a.z, sh.z);
    float o3 = fmaf(v.w, a.w, sh.w);
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
  uint32_t p0 = i / ucq, q0 = i - p0 * ucq;
  // Double-buffered stream: the four 16-byte loads of step k+1 are issued before step k is normalised and stored
  // (8 loads in flight per thread), and the very first group is in flight while the block computes its scale / shift.
  float4 v[4];
  uint32_t qv[4];
  auto load4 = [&](uint32_t at, float4 (&vv)[4], uint32_t (&qq)[4]) {   // consumes and advances (p0, q0)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      qq[u] = q0;
      vv[u] = __ldg(reinterpret_cast<const float4*>(src_at(at + 256u * u, p0, q0)));
      advance(p0, q0);
    }
  };
  bool have = i + 768 < end;
  if (have) load4(i, v, qv);
  prologue();
  while (have) {
    const uint32_t ni = i + 1024;
    const bool have_next = ni + 768 < end;
    float4 vn[4];
    uint32_t qn[4];
    if (have_next) load4(ni, vn, qn);
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(i + 256u * u, static_cast<int>(qv[u]) * 4, v[u]);
    if (have_next) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = vn[u];
        qv[u] = qn[u];
      }
    }
    i = ni;
    have = have_next;
  }
  for (; i < end; i += 256) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src_at(i, p0, q0)));
    emit(i, static_cast<int>(q0) * 4, t);
    advance(p0, q0);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (C <= 2048, C % 128 == 0 not required; C % 4 == 0)
// ---------------------------------------------------------------------------------------------
template <int MAXV, bool F32OUT>  // float4 vectors per lane; output bf16 (GEMM operand) or fp32
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long rows, int C,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        void* __restrict__ yv) {
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
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      const float4 g = __ldg(g4 + k), bb = __ldg(b4 + k);
      const float4 o = make_float4((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                                   (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      if (F32OUT) {
        reinterpret_cast<float4*>(static_cast<float*>(yv) + row * C)[k] = o;
      } else {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(yv) + row * C)[k] = pk;
      }
    }
  }
}

}  // namespace af

using namespace af;

static int gn_default_slots(int B, int HW) {
  // enough blocks to fill the machine, few enough that the finalize pass stays cheap
  int slots = (8 * num_sms() + B - 1) / B;
  if (slots > HW / 16) slots = HW / 16;   // keep >= 16 pixels per block
  if (slots > AF_GN_MAX_CHUNKS) slots = AF_GN_MAX_CHUNKS;
  if (slots > HW) slots = HW;
  if (slots < 1) slots = 1;
  return slots;
}

extern "C" int af_groupnorm_stats_slots(int B, int HW) { return gn_default_slots(B, HW); }

extern "C" size_t af_groupnorm_workspace_bytes(int B, int C) {
  // per-channel partials of both sources (<= AF_GN_MAX_CHUNKS slots) + (mean, rstd) per (sample, group)
  return (static_cast<size_t>(B) * AF_GN_MAX_CHUNKS * C * 2 + static_cast<size_t>(B) * kGroups * 2) * sizeof(float);
}

extern "C" int af_groupnorm_stats(const float* x, int C, int B, int HW, float* stats, int slots, cudaStream_t stream) {
  AF_CHECK_ARG(x && stats, "af_groupnorm_stats: null pointer");
  AF_CHECK_ARG(B > 0 && HW > 0 && C > 0 && C % 4 == 0 && C <= 5120, "af_groupnorm_stats: bad sizes B=%d HW=%d C=%d", B, HW, C);
  AF_CHECK_ARG(slots >= 1 && slots <= HW, "af_groupnorm_stats: slots=%d", slots);
  const int cq = C / 4;
  const size_t smem = static_cast<size_t>(cq <= 256 ? (256 / cq) * cq : cq) * 8 * sizeof(float);
  gn_stats_kernel<<<dim3(slots, B), 256, smem, stream>>>(x, C, HW, slots, stats);
  AF_LAUNCH_CHECK("gn_stats_kernel");
  return 0;
}

extern "C" int af_groupnorm_finalize(const float* stats0, int C0, int slots0, const float* stats1, int C1, int slots1,
                                     int B, int HW, float eps, float* mean_rstd, cudaStream_t stream) {
  AF_CHECK_ARG(stats0 && mean_rstd, "af_groupnorm_finalize: null pointer");
  AF_CHECK_ARG(B > 0 && HW > 0 && C0 > 0 && C1 >= 0 && (C0 + C1) % 32 == 0 && slots0 > 0, "af_groupnorm_finalize: bad sizes");
  AF_CHECK_ARG(C1 == 0 || (stats1 && slots1 > 0), "af_groupnorm_finalize: second source needs stats");
  gn_finalize_kernel<<<dim3(kGroups, B), 256, 0, stream>>>(stats0, C0, slots0, stats1, C1, slots1, HW, eps, mean_rstd);
  AF_LAUNCH_CHECK("gn_finalize_kernel");
  return 0;
}

extern "C" int af_groupnorm_apply(const float* x0, int C0, const float* x1, int C1, int B, int HW,
                                  const float* mean_rstd, const float* gamma, const float* beta, int silu,
                                  void* y_bf16, void* raw_bf16, cudaStream_t stream) {
  AF_CHECK_ARG(x0 && mean_rstd && gamma && beta && y_bf16, "af_groupnorm_apply: null pointer");
  const int C = C0 + C1;
  AF_CHECK_ARG(B > 0 && HW > 0 && C0 > 0 && C1 >= 0, "af_groupnorm_apply: bad sizes");
  AF_CHECK_ARG(C % 32 == 0 && C0 % 4 == 0 && C1 % 4 == 0, "af_groupnorm_apply: C0=%d C1=%d need C%%32==0, quads", C0, C1);
  AF_CHECK_ARG(C1 == 0 || x1 != nullptr, "af_groupnorm_apply: x1 null with C1=%d", C1);
  AF_CHECK_ARG(C <= 5120, "af_groupnorm_apply: C=%d too large", C);
  const size_t total = static_cast<size_t>(HW) * (C / 4);
  int blocks = static_cast<int>((total + 256 * 8 - 1) / (256 * 8));
  const int cap = (4 * num_sms() + B - 1) / B;   // one resident wave (4 CTAs x 256 threads x 64 registers per SM): the prologue is paid once
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  gn_apply_kernel<<<dim3(blocks, B), 256, 2 * C * sizeof(float), stream>>>(
      x0, C0, x1, C1, HW, mean_rstd, gamma, beta, silu, static_cast<__nv_bfloat16*>(y_bf16),
      static_cast<__nv_bfloat16*>(raw_bf16));
  AF_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}

// all-in-one: statistics pass(es) + finalize + apply
extern "C" int af_groupnorm_silu(const float* x0, int C0, const float* x1, int C1, int B, int HW,
                                 const float* gamma, const float* beta, float eps, int silu, void* y_bf16,
                                 void* raw_bf16, float* workspace, cudaStream_t stream) {
  AF_CHECK_ARG(x0 && gamma && beta && y_bf16 && workspace, "af_groupnorm_silu: null pointer");
  AF_CHECK_ARG(B > 0 && HW > 0 && C0 > 0 && C1 >= 0 && C0 % 4 == 0 && C1 % 4 == 0, "af_groupnorm_silu: bad sizes");
  const int slots = gn_default_slots(B, HW);
  float* st0 = workspace;
  float* st1 = st0 + static_cast<size_t>(B) * slots * C0 * 2;
  float* mr = workspace + static_cast<size_t>(B) * AF_GN_MAX_CHUNKS * (C0 + C1) * 2;
  int rc = af_groupnorm_stats(x0, C0, B, HW, st0, slots, stream);
  if (rc) return rc;
  if (C1 > 0) {
    AF_CHECK_ARG(x1 != nullptr, "af_groupnorm_silu: x1 null with C1=%d", C1);
    rc = af_groupnorm_stats(x1, C1, B, HW, st1, slots, stream);
    if (rc) return rc;
  }
  rc = af_groupnorm_finalize(st0, C0, slots, C1 > 0 ? st1 : nullptr, C1, slots, B, HW, eps, mr, stream);
  if (rc) return rc;
  return af_groupnorm_apply(x0, C0, x1, C1, B, HW, mr, gamma, beta, silu, y_bf16, raw_bf16, stream);
}

template <bool F32OUT>
static int launch_layernorm(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                            void* y, cudaStream_t stream) {
  AF_CHECK_ARG(x && gamma && beta && y, "af_layernorm: null pointer");
  AF_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && C <= 2048, "af_layernorm: rows=%lld C=%d (need C%%4==0, C<=2048)", rows, C);
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  const int nvec = C / 4;
  if (nvec <= 3 * 32) {
    layernorm_kernel<3, F32OUT><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, y);
  } else if (nvec <= 5 * 32) {
    layernorm_kernel<5, F32OUT><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, y);
  } else if (nvec <= 10 * 32) {
    layernorm_kernel<10, F32OUT><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, y);
  } else {
    layernorm_kernel<16, F32OUT><<<grid, 256, 0, stream>>>(x, rows, C, gamma, beta, eps, y);
  }
  AF_LAUNCH_CHECK("layernorm_kernel");
  return 0;
}

extern "C" int af_layernorm(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                            void* y_bf16, cudaStream_t stream) {
  return launch_layernorm<false>(x, rows, C, gamma, beta, eps, y_bf16, stream);
}

extern "C" int af_layernorm_f32(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps,
                                float* y, cudaStream_t stream) {
  return launch_layernorm<true>(x, rows, C, gamma, beta, eps, y, stream);
}
