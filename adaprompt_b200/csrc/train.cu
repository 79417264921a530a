// Backward kernels of the Stage-1 distillation step (guided_denoise ddpm.py:2483-2532 runs
// the UNet WITH grad; weights are frozen, ddpm.py:783-786, so the UNet needs activation gradients only - they reach the
// trainable SubjBasisGenerator through the cross-attention context).  Everything here is bandwidth-bound glue between
// the tensor-core dgrad GEMMs / convolutions (the forward kernels run on transposed weight packs):
//   GroupNorm(+SiLU) backward, LayerNorm backward, GEGLU forward/backward (unfused training form), quick_gelu,
//   delta = rowsum(dO o O), 77-token CLIP attention backward, conv_out dgrad, 2x2 sum-pool / zero-insert for the
//   resampling convolutions, bf16 transpose (weight-gradient operands), SGD-free helpers.
// Reductions that feed parameters (LayerNorm gamma/beta) use fp32 atomics; activation gradients are deterministic.
#include <math.h>

#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------- GroupNorm backward
// dxhat = dy * silu'(z) * gamma, z = xhat * gamma + beta;  dx = rstd * (dxhat - mean_g(dxhat) - xhat * mean_g(dxhat*xhat))
__device__ __forceinline__ float gn_dxhat(float xhat, float dy, float g, float b, int silu) {
  if (silu) {
    const float z = xhat * g + b;
    const float s = sigmoid_f(z);
    dy *= s * (1.0f + z * (1.0f - s));
  }
  return dy * g;
}

// grid = (chunks, B); block = 256.  Thread t owns channel quad (t % cq) and pixel lane (t / cq) of the chunk's rows:
// 16-byte x loads and 8-byte dy loads, four rows in flight per thread, lanes folded through shared memory in a fixed
// order (bit-reproducible).  (The first version walked one channel per thread with scalar loads in a serial row loop:
// 155 us per call on average, 11 % of the stage-1 step.)
__global__ void __launch_bounds__(256) gn_bwd_partial_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                             const float* __restrict__ mean_rstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int C, int HW, int rows_per_chunk, int silu,
                                                             float* __restrict__ partial) {
  extern __shared__ float s_part[];   // [lanes][cq][8]
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const int cpg = C / 32, cq = C >> 2;
  const int lanes = cq <= 256 ? 256 / cq : 1;
  auto accumulate = [&](int q, int r_first, int r_step, int dst) {
    const int c = q * 4;
    float mean[4], rstd[4], ga[4], be[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int g = (c + e) / cpg;
      mean[e] = mean_rstd[(b * 32 + g) * 2];
      rstd[e] = mean_rstd[(b * 32 + g) * 2 + 1];
      ga[e] = gamma[c + e];
      be[e] = beta[c + e];
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t base = static_cast<size_t>(b) * HW * C + c;
    auto acc = [&](const float4& xv, const uint2& dv) {
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv.x);
      const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv.y);
      const float ds[4] = {__low2float(d01), __high2float(d01), __low2float(d23), __high2float(d23)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (xs[e] - mean[e]) * rstd[e];
        const float d = gn_dxhat(xh, ds[e], ga[e], be[e], silu);
        s1[e] += d;
        s2[e] = fmaf(d, xh, s2[e]);
      }
    };
    int r = r_first;
    for (; r + 3 * r_step < r1; r += 4 * r_step) {
      float4 xv[4];
      uint2 dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t i = base + static_cast<size_t>(r + u * r_step) * C;
        xv[u] = __ldg(reinterpret_cast<const float4*>(x + i));
        dv[u] = __ldg(reinterpret_cast<const uint2*>(dy + i));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc(xv[u], dv[u]);
    }
    for (; r < r1; r += r_step) {
      const size_t i = base + static_cast<size_t>(r) * C;
      acc(__ldg(reinterpret_cast<const float4*>(x + i)), __ldg(reinterpret_cast<const uint2*>(dy + i)));
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_part[dst * 8 + 2 * e] = s1[e];
      s_part[dst * 8 + 2 * e + 1] = s2[e];
    }
  };
  if (cq <= 256) {
    if (threadIdx.x < lanes * cq) accumulate(threadIdx.x % cq, r0 + threadIdx.x / cq, lanes, threadIdx.x);
  } else {
    for (int q = threadIdx.x; q < cq; q += 256) accumulate(q, r0, 1, q);
  }
  __syncthreads();
  float* out = partial + (static_cast<size_t>(b) * chunks + chunk) * C * 2;
  for (int q = threadIdx.x; q < cq; q += 256) {
    float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] += s_part[(l * cq + q) * 8 + e];
    float4* o = reinterpret_cast<float4*>(out + q * 8);
    o[0] = make_float4(t[0], t[1], t[2], t[3]);
    o[1] = make_float4(t[4], t[5], t[6], t[7]);
  }
}

__global__ void __launch_bounds__(128) gn_bwd_finalize_kernel(const float* __restrict__ partial, int C, int chunks,
                                                              float* __restrict__ sums) {
  const int b = blockIdx.x / 32, g = blockIdx.x % 32;
  const int cpg = C / 32;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < chunks * cpg; i += 128) {
    const int chunk = i / cpg, c = g * cpg + i % cpg;
    const float* o = partial + ((static_cast<size_t>(b) * chunks + chunk) * C + c) * 2;
    s1 += o[0];
    s2 += o[1];
  }
  __shared__ float sh[2][4];
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s1;
    sh[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sums[blockIdx.x * 2] = (sh[0][0] + sh[0][1]) + (sh[0][2] + sh[0][3]);
    sums[blockIdx.x * 2 + 1] = (sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3]);
  }
}

// grid = (blocks, B); block = 256: channel quads, 16-byte loads / stores, channel index tracked incrementally
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                           const float* __restrict__ mean_rstd, const float* __restrict__ sums,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ dres, int C, int HW, int silu,
                                                           float* __restrict__ dx) {
  const int b = blockIdx.y;
  const int cpg = C / 32;
  const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
  extern __shared__ float s_tab[];          // [C][6]: mean, rstd, m1, m2, gamma, beta per channel
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cpg;
    s_tab[c * 6 + 0] = mean_rstd[(b * 32 + g) * 2];
    s_tab[c * 6 + 1] = mean_rstd[(b * 32 + g) * 2 + 1];
    s_tab[c * 6 + 2] = sums[(b * 32 + g) * 2] * inv_n;
    s_tab[c * 6 + 3] = sums[(b * 32 + g) * 2 + 1] * inv_n;
    s_tab[c * 6 + 4] = gamma[c];
    s_tab[c * 6 + 5] = beta[c];
  }
  __syncthreads();
  const uint32_t cq = static_cast<uint32_t>(C >> 2);
  const uint32_t total = static_cast<uint32_t>(HW) * cq;
  const uint32_t step = gridDim.x * 256u;
  const uint32_t step_q = step % cq;
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  uint32_t q = i % cq;
  const size_t sample = static_cast<size_t>(b) * HW * C;
  for (; i < total; i += step) {
    const size_t k = sample + static_cast<size_t>(i) * 4;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + k));
    const uint2 dv = __ldg(reinterpret_cast<const uint2*>(dy + k));
    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dres) rv = __ldg(reinterpret_cast<const float4*>(dres + k));
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv.x);
    const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv.y);
    const float ds[4] = {__low2float(d01), __high2float(d01), __low2float(d23), __high2float(d23)};
    const float rs[4] = {rv.x, rv.y, rv.z, rv.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* t = s_tab + (q * 4 + e) * 6;
      const float xh = (xs[e] - t[0]) * t[1];
      const float d = gn_dxhat(xh, ds[e], t[4], t[5], silu);
      o[e] = t[1] * (d - t[2] - xh * t[3]) + rs[e];
    }
    *reinterpret_cast<float4*>(dx + k) = make_float4(o[0], o[1], o[2], o[3]);
    q += step_q;
    if (q >= cq) q -= cq;
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm backward
template <int MAXV, bool DY_F32>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, long long rows, int C,
                                                            const float* __restrict__ gamma, float eps,
                                                            const void* __restrict__ dyv, const float* __restrict__ dres,
                                                            float* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[MAXV], d[MAXV];
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
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / C + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      float4 dy;
      if (DY_F32) {
        dy = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(dyv) + row * C) + k);
      } else {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dyv) + row * C) + k);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        dy = make_float4(bf(lo.x), bf(lo.y), bf(hi.x), bf(hi.y));
      }
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;   // xhat
      if (dgamma) {
        atomicAdd(dgamma + 4 * k, dy.x * v[i].x); atomicAdd(dgamma + 4 * k + 1, dy.y * v[i].y);
        atomicAdd(dgamma + 4 * k + 2, dy.z * v[i].z); atomicAdd(dgamma + 4 * k + 3, dy.w * v[i].w);
        atomicAdd(dbeta + 4 * k, dy.x); atomicAdd(dbeta + 4 * k + 1, dy.y);
        atomicAdd(dbeta + 4 * k + 2, dy.z); atomicAdd(dbeta + 4 * k + 3, dy.w);
      }
      const float4 g = __ldg(g4 + k);
      d[i] = make_float4(dy.x * g.x, dy.y * g.y, dy.z * g.z, dy.w * g.w);
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * v[i].x + d[i].y * v[i].y + d[i].z * v[i].z + d[i].w * v[i].w;
    }
  }
  s1 = warp_sum(s1) / C;
  s2 = warp_sum(s2) / C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int k = lane + i * 32;
    if (k < nvec) {
      float4 o = make_float4(rstd * (d[i].x - s1 - v[i].x * s2), rstd * (d[i].y - s1 - v[i].y * s2),
                             rstd * (d[i].z - s1 - v[i].z * s2), rstd * (d[i].w - s1 - v[i].w * s2));
      if (dres) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(dres + row * C) + k);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dx + row * C)[k] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------- GEGLU / quick_gelu
// proj [T, 2F] bf16 = Linear output (value | gate), h = value * gelu(gate) (attention.py:32-39, exact erf GELU)
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __nv_bfloat16* __restrict__ proj, long long T, int F,
                                                        __nv_bfloat16* __restrict__ h) {
  const long long n = T * F;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const long long t = i / F;
    const int f = static_cast<int>(i % F);
    const float a = bf(proj[t * 2 * F + f]), g = bf(proj[t * 2 * F + F + f]);
    h[i] = __float2bfloat16(a * gelu_erf_f(g));
  }
}
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ proj, const __nv_bfloat16* __restrict__ dh,
                                                        long long T, int F, __nv_bfloat16* __restrict__ dproj) {
  const long long n = T * F;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const long long t = i / F;
    const int f = static_cast<int>(i % F);
    const float a = bf(proj[t * 2 * F + f]), g = bf(proj[t * 2 * F + F + f]), d = bf(dh[i]);
    const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * g * g);
    dproj[t * 2 * F + f] = __float2bfloat16(d * g * cdf);
    dproj[t * 2 * F + F + f] = __float2bfloat16(d * a * (cdf + g * pdf));
  }
}
// quick_gelu(x) = x * sigmoid(1.702 x) (CLIP text MLP)
__global__ void __launch_bounds__(256) quick_gelu_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                         long long n, __nv_bfloat16* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float v = bf(x[i]);
    const float s = sigmoid_f(1.702f * v);
    out[i] = __float2bfloat16(dy ? bf(dy[i]) * s * (1.0f + 1.702f * v * (1.0f - s)) : v * s);
  }
}

// ------------------------------------------------------------------------------------------- attention helpers
// delta[b][h][q] = sum_c dO[b,q,h*d+c] * O[b,q,h*d+c]   (one warp per (row, head))
__global__ void __launch_bounds__(256) rowdot_heads_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b2,
                                                           int B, int N, int heads, int d, float* __restrict__ delta) {
  const long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<long long>(B) * N * heads) return;
  const int h = static_cast<int>(w % heads);
  const long long row = w / heads;
  const __nv_bfloat16* pa = a + row * heads * d + h * d;
  const __nv_bfloat16* pb = b2 + row * heads * d + h * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += bf(pa[c]) * bf(pb[c]);
  s = warp_sum(s);
  if (lane == 0) delta[(static_cast<size_t>(row / N) * heads + h) * N + row % N] = s;
}

// CLIP text attention backward (forward: text.cu attention_small_kernel; adaface/arc2face_models.py:87-173).
// One CTA per (sample, head); everything in shared memory in fp32.  Lk = L * mult keys, key j = (token j / mult ... see
// layout note) - key (t, r) lives in row t of qkv at column k_off + (h*mult + r)*64 and is visible to query i iff t <= i.
__global__ void __launch_bounds__(1024) attention_small_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq, int k_off,
                                                                  int v_off, const __nv_bfloat16* __restrict__ dout, long long ldo,
                                                                  __nv_bfloat16* __restrict__ dqkv, int heads, int L, int mult,
                                                                  float scale, int causal) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int Lk = L * mult;
  float* Q = sm;                 // [L][64]
  float* dO = Q + L * 64;        // [L][64]
  // (1024 threads per (sample, head): only B x heads blocks exist and every phase is a block-wide strided loop)
  // K / V rows are padded to 65 floats: the score loop below walks them with one KEY per thread (stride-64 rows put all
  // 32 lanes on one bank: the kernel took 146 us for a 77-token layer)
  float* K = dO + L * 64;        // [Lk][65]
  float* V = K + Lk * 65;        // [Lk][65]
  float* P = V + Lk * 65;        // [L][Lk]
  float* dS = P + L * Lk;        // [L][Lk]
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * L * ldq;
  for (int i = threadIdx.x; i < L * 64; i += blockDim.x) {
    const int t = i >> 6, c = i & 63;
    Q[i] = bf(base[t * ldq + h * 64 + c]) * scale;
    dO[i] = bf(dout[(static_cast<size_t>(b) * L + t) * ldo + h * 64 + c]);
  }
  for (int i = threadIdx.x; i < Lk * 64; i += blockDim.x) {
    const int j = i >> 6, c = i & 63, t = j / mult, r = j % mult;
    K[j * 65 + c] = bf(base[t * ldq + k_off + (h * mult + r) * 64 + c]);
    V[j * 65 + c] = bf(base[t * ldq + v_off + (h * mult + r) * 64 + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * Lk; i += blockDim.x) {
    const int q = i / Lk, j = i % Lk;
    float s = -INFINITY, dp = 0.f;
    if (!causal || j / mult <= q) {
      s = 0.f;
      for (int c = 0; c < 64; ++c) {
        s += Q[q * 64 + c] * K[j * 65 + c];
        dp += dO[q * 64 + c] * V[j * 65 + c];
      }
    }
    P[i] = s;
    dS[i] = dp;
  }
  __syncthreads();
  for (int q = threadIdx.x >> 5; q < L; q += (blockDim.x >> 5)) {     // one warp per row: softmax, delta, dS
    const int lane = threadIdx.x & 31;
    float m = -INFINITY;
    for (int j = lane; j < Lk; j += 32) m = fmaxf(m, P[q * Lk + j]);
    m = warp_max(m);
    float l = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = __expf(P[q * Lk + j] - m);
      P[q * Lk + j] = e;
      l += e;
    }
    l = warp_sum(l);
    const float inv = 1.0f / l;
    float dl = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float pv = P[q * Lk + j] * inv;
      P[q * Lk + j] = pv;
      dl += pv * dS[q * Lk + j];
    }
    dl = warp_sum(dl);
    for (int j = lane; j < Lk; j += 32) dS[q * Lk + j] = P[q * Lk + j] * (dS[q * Lk + j] - dl);
  }
  __syncthreads();
  __nv_bfloat16* obase = dqkv + static_cast<size_t>(b) * L * ldq;
  for (int i = threadIdx.x; i < L * 64; i += blockDim.x) {    // dQ (carries the forward scale)
    const int q = i >> 6, c = i & 63;
    float s = 0.f;
    for (int j = 0; j < Lk; ++j) s += dS[q * Lk + j] * K[j * 65 + c];
    obase[q * ldq + h * 64 + c] = __float2bfloat16(s * scale);
  }
  for (int i = threadIdx.x; i < Lk * 64; i += blockDim.x) {   // dK, dV
    const int j = i >> 6, c = i & 63, t = j / mult, r = j % mult;
    float sk = 0.f, sv = 0.f;
    for (int q = 0; q < L; ++q) {
      sk += dS[q * Lk + j] * Q[q * 64 + c];            // Q already scaled
      sv += P[q * Lk + j] * dO[q * 64 + c];
    }
    obase[t * ldq + k_off + (h * mult + r) * 64 + c] = __float2bfloat16(sk);
    obase[t * ldq + v_off + (h * mult + r) * 64 + c] = __float2bfloat16(sv);
  }
}

// ------------------------------------------------------------------------------------------- convolution glue
// dy[b,u,v,ci] = sum_{ky,kx,co} W[co,ci,ky,kx] * dout[b,co,u-ky+1,v-kx+1]   (UNetModel.out[2], openaimodel.py:696)
__global__ void __launch_bounds__(256) conv_out_dgrad_kernel(const float* __restrict__ dout_nchw, const float* __restrict__ w,
                                                             int B, int H, int W, int C, int Cout, __nv_bfloat16* __restrict__ dy) {
  const size_t n = static_cast<size_t>(B) * H * W * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256) {
    const int ci = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int v = static_cast<int>(pix % W), u = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    float s = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int y = u - ky + 1;
      if (y < 0 || y >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int x = v - kx + 1;
        if (x < 0 || x >= W) continue;
        for (int co = 0; co < Cout; ++co)
          s += w[((co * C + ci) * 3 + ky) * 3 + kx] * dout_nchw[((static_cast<size_t>(b) * Cout + co) * H + y) * W + x];
      }
    }
    dy[i] = __float2bfloat16(s);
  }
}

// out[b,h,w,c] = sum of the 2x2 block of in[b,2h..2h+1,2w..2w+1,c]  (backward of F.interpolate nearest x2, openaimodel.py:120)
__global__ void __launch_bounds__(256) sumpool2x2_kernel(const float* __restrict__ in, int B, int H, int W, int C, float* __restrict__ out) {
  const size_t n = static_cast<size_t>(B) * H * W * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    const size_t r0 = ((static_cast<size_t>(b) * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
    const size_t r1 = r0 + static_cast<size_t>(2 * W) * C;
    out[i] = (in[r0] + in[r0 + C]) + (in[r1] + in[r1 + C]);
  }
}
// out bf16 [B,2H,2W,C]: out[b,2h,2w,:] = in[b,h,w,:], zero elsewhere (dgrad of the stride-2 Downsample conv, openaimodel.py:155)
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                                            __nv_bfloat16* __restrict__ out) {
  const size_t n = static_cast<size_t>(B) * 2 * H * 2 * W * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int w = static_cast<int>(pix % (2 * W)), h = static_cast<int>((pix / (2 * W)) % (2 * H));
    const int b = static_cast<int>(pix / (static_cast<size_t>(4) * W * H));
    float v = 0.f;
    if (((w | h) & 1) == 0) v = in[((static_cast<size_t>(b) * H + (h >> 1)) * W + (w >> 1)) * C + c];
    out[i] = __float2bfloat16(v);
  }
}
// out [C][ldo] bf16 = in [R][C]^T (fp32 or bf16 source), columns R..ldo-1 zero: operands of the weight-gradient GEMMs
template <typename T>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const T* __restrict__ in, int R, int C, long long ldo,
                                                                __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < C) ? static_cast<float>(in[static_cast<size_t>(r) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < C && r < ldo) out[static_cast<size_t>(c) * ldo + r] = __float2bfloat16(tile[tx][j]);
  }
}

// ------------------------------------------------------------------------------------------- Prodigy (ldm/prodigy.py)
// pass 1 (:177-189): Adam moments scaled by d, s <- beta3 s + s_alpha g, and the two global sums of the d estimate:
// sums[0] += g . (p0 - p), sums[1] += |s_new|  (double accumulators).
__global__ void __launch_bounds__(256) prodigy_moments_kernel(const float* __restrict__ p, const float* __restrict__ grad,
                                                              const float* __restrict__ p0, float* __restrict__ s,
                                                              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                              long long n, float beta1, float beta2, float beta3, float d,
                                                              float s_alpha, float coupled_decay, double* __restrict__ sums) {
  double dot = 0.0, den = 0.0;
  const float c1 = d * (1.0f - beta1), c2 = d * d * (1.0f - beta2);
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float pv = p[i];
    const float g = grad[i] + coupled_decay * pv;
    dot += static_cast<double>(g) * static_cast<double>(p0[i] - pv);
    exp_avg[i] = exp_avg[i] * beta1 + c1 * g;
    exp_avg_sq[i] = exp_avg_sq[i] * beta2 + c2 * g * g;
    const float sn = s[i] * beta3 + s_alpha * g;
    s[i] = sn;
    den += fabs(static_cast<double>(sn));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  __shared__ double sh[2][8];
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = dot;
    sh[1][threadIdx.x >> 5] = den;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += sh[0][w];
      b += sh[1][w];
    }
    atomicAdd(sums, a);
    atomicAdd(sums + 1, b);
  }
}
// pass 2 (:240-248): decoupled decay, then p -= dlr * exp_avg / (sqrt(exp_avg_sq) + d * eps)
__global__ void __launch_bounds__(256) prodigy_apply_kernel(float* __restrict__ p, const float* __restrict__ exp_avg,
                                                            const float* __restrict__ exp_avg_sq, long long n, float dlr,
                                                            float d_eps, float decoupled_decay) {
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    float pv = p[i];
    pv += pv * (-decoupled_decay * dlr);
    pv += -dlr * (exp_avg[i] / (sqrtf(exp_avg_sq[i]) + d_eps));
    p[i] = pv;
  }
}

static inline unsigned grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  return static_cast<unsigned>(g < cap ? (g ? g : 1) : cap);
}

}  // namespace af

using namespace af;

static int gn_bwd_chunks(int B, int HW) {
  int chunks = (4 * num_sms() + B - 1) / B;      // fill the machine ...
  if (chunks > HW / 16) chunks = HW / 16;        // ... with at least 16 rows per block
  if (chunks > 256) chunks = 256;
  if (chunks < 1) chunks = 1;
  return chunks;
}

extern "C" size_t af_groupnorm_bwd_workspace_floats(int B, int C, int HW) {
  (void)HW;
  const int chunks = 256;                        // upper bound of gn_bwd_chunks
  return static_cast<size_t>(B) * chunks * C * 2 + static_cast<size_t>(B) * 64;
}

extern "C" int af_groupnorm_bwd(const float* x, int C, int B, int HW, const float* mean_rstd, const float* gamma,
                                const float* beta, int silu, const void* dy_bf16, const float* dres, float* dx,
                                float* workspace, cudaStream_t stream) {
  AF_CHECK_ARG(x && mean_rstd && gamma && beta && dy_bf16 && dx && workspace, "af_groupnorm_bwd: null pointer");
  AF_CHECK_ARG(C > 0 && C % 32 == 0 && B > 0 && HW > 0 && C <= 5120, "af_groupnorm_bwd: bad sizes");
  AF_CHECK_ARG(static_cast<long long>(HW) * (C / 4) < (1ll << 31), "af_groupnorm_bwd: sample too large for 32-bit quad indices");
  const int chunks = gn_bwd_chunks(B, HW);
  const int rows_per_chunk = (HW + chunks - 1) / chunks;
  float* partial = workspace;
  float* sums = workspace + static_cast<size_t>(B) * chunks * C * 2;
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(dy_bf16);
  const int cq = C / 4;
  const size_t smem_p = static_cast<size_t>(cq <= 256 ? (256 / cq) * cq : cq) * 8 * sizeof(float);
  gn_bwd_partial_kernel<<<dim3(chunks, B), 256, smem_p, stream>>>(x, dy, mean_rstd, gamma, beta, C, HW, rows_per_chunk, silu, partial);
  AF_LAUNCH_CHECK("gn_bwd_partial_kernel");
  gn_bwd_finalize_kernel<<<B * 32, 128, 0, stream>>>(partial, C, chunks, sums);
  AF_LAUNCH_CHECK("gn_bwd_finalize_kernel");
  const size_t nq = static_cast<size_t>(HW) * cq;
  unsigned gx = static_cast<unsigned>((nq + 256 * 4 - 1) / (256 * 4));
  const unsigned cap = static_cast<unsigned>(num_sms() * 8 / B + 1);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const size_t smem_a = static_cast<size_t>(C) * 6 * sizeof(float);
  static size_t configured = 0;
  if (smem_a > 48 * 1024 && smem_a > configured) {
    AF_CUDA(cudaFuncSetAttribute(gn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_a)));
    configured = smem_a;
  }
  gn_bwd_apply_kernel<<<dim3(gx, B), 256, smem_a, stream>>>(x, dy, mean_rstd, sums, gamma, beta, dres, C, HW, silu, dx);
  AF_LAUNCH_CHECK("gn_bwd_apply_kernel");
  return 0;
}

extern "C" int af_layernorm_bwd(const float* x, long long rows, int C, const float* gamma, float eps, const void* dy,
                                int dy_dtype, const float* dres, float* dx, float* dgamma, float* dbeta,
                                cudaStream_t stream) {
  AF_CHECK_ARG(x && gamma && dy && dx, "af_layernorm_bwd: null pointer");
  AF_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0 && C <= 2048, "af_layernorm_bwd: rows=%lld C=%d", rows, C);
  AF_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "af_layernorm_bwd: dgamma / dbeta must come together");
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  const int nvec = C / 4;
  const bool f32 = dy_dtype == AF_DTYPE_F32;
#define AF_LNB(MV)                                                                                                  \
  do {                                                                                                              \
    if (f32) layernorm_bwd_kernel<MV, true><<<grid, 256, 0, stream>>>(x, rows, C, gamma, eps, dy, dres, dx, dgamma, dbeta); \
    else layernorm_bwd_kernel<MV, false><<<grid, 256, 0, stream>>>(x, rows, C, gamma, eps, dy, dres, dx, dgamma, dbeta);    \
  } while (0)
  if (nvec <= 3 * 32) AF_LNB(3);
  else if (nvec <= 6 * 32) AF_LNB(6);
  else if (nvec <= 10 * 32) AF_LNB(10);
  else AF_LNB(16);
#undef AF_LNB
  AF_LAUNCH_CHECK("layernorm_bwd_kernel");
  return 0;
}

extern "C" int af_geglu_fwd(const void* proj, long long T, int F, void* h, cudaStream_t stream) {
  AF_CHECK_ARG(proj && h && T > 0 && F > 0, "af_geglu_fwd: bad arguments");
  geglu_fwd_kernel<<<grid_for(static_cast<size_t>(T) * F), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(proj), T, F,
                                                                             static_cast<__nv_bfloat16*>(h));
  AF_LAUNCH_CHECK("geglu_fwd_kernel");
  return 0;
}
extern "C" int af_geglu_bwd(const void* proj, const void* dh, long long T, int F, void* dproj, cudaStream_t stream) {
  AF_CHECK_ARG(proj && dh && dproj && T > 0 && F > 0, "af_geglu_bwd: bad arguments");
  geglu_bwd_kernel<<<grid_for(static_cast<size_t>(T) * F), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(proj),
                                                                             static_cast<const __nv_bfloat16*>(dh), T, F,
                                                                             static_cast<__nv_bfloat16*>(dproj));
  AF_LAUNCH_CHECK("geglu_bwd_kernel");
  return 0;
}
extern "C" int af_quick_gelu(const void* x, const void* dy, long long n, void* out, cudaStream_t stream) {
  AF_CHECK_ARG(x && out && n > 0, "af_quick_gelu: bad arguments");
  quick_gelu_kernel<<<grid_for(static_cast<size_t>(n)), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                          static_cast<const __nv_bfloat16*>(dy), n,
                                                                          static_cast<__nv_bfloat16*>(out));
  AF_LAUNCH_CHECK("quick_gelu_kernel");
  return 0;
}

extern "C" int af_rowdot_heads(const void* a, const void* b, int B, int N, int heads, int d, float* delta, cudaStream_t stream) {
  AF_CHECK_ARG(a && b && delta && B > 0 && N > 0 && heads > 0 && d > 0, "af_rowdot_heads: bad arguments");
  const long long warps = static_cast<long long>(B) * N * heads;
  rowdot_heads_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(a),
                                                                                  static_cast<const __nv_bfloat16*>(b), B, N, heads, d, delta);
  AF_LAUNCH_CHECK("rowdot_heads_kernel");
  return 0;
}

extern "C" int af_attention_small_bwd(const void* qkv, long long ldq, int k_off, int v_off, const void* dout, long long ldo,
                                      void* dqkv, int B, int heads, int L, int mult, float scale, int causal,
                                      cudaStream_t stream) {
  AF_CHECK_ARG(qkv && dout && dqkv && B > 0 && heads > 0 && L > 0 && mult >= 1, "af_attention_small_bwd: bad arguments");
  const int Lk = L * mult;
  const size_t smem = (static_cast<size_t>(2) * L * 64 + static_cast<size_t>(2) * Lk * 65 + static_cast<size_t>(2) * L * Lk) * 4;
  AF_CHECK_ARG(smem <= 227 * 1024, "af_attention_small_bwd: L=%d mult=%d needs %zu bytes of shared memory", L, mult, smem);
  AF_CUDA(cudaFuncSetAttribute(attention_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  attention_small_bwd_kernel<<<B * heads, 1024, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv), ldq, k_off, v_off,
                                                               static_cast<const __nv_bfloat16*>(dout), ldo,
                                                               static_cast<__nv_bfloat16*>(dqkv), heads, L, mult, scale, causal);
  AF_LAUNCH_CHECK("attention_small_bwd_kernel");
  return 0;
}

extern "C" int af_conv_out_dgrad(const float* dout_nchw, const float* w, int B, int H, int W, int C, int Cout, void* dy_bf16,
                                 cudaStream_t stream) {
  AF_CHECK_ARG(dout_nchw && w && dy_bf16 && B > 0 && H > 0 && W > 0 && C > 0 && Cout > 0, "af_conv_out_dgrad: bad arguments");
  conv_out_dgrad_kernel<<<grid_for(static_cast<size_t>(B) * H * W * C), 256, 0, stream>>>(dout_nchw, w, B, H, W, C, Cout,
                                                                                          static_cast<__nv_bfloat16*>(dy_bf16));
  AF_LAUNCH_CHECK("conv_out_dgrad_kernel");
  return 0;
}

extern "C" int af_sumpool2x2(const float* in, int B, int H, int W, int C, float* out, cudaStream_t stream) {
  AF_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0, "af_sumpool2x2: bad arguments");
  sumpool2x2_kernel<<<grid_for(static_cast<size_t>(B) * H * W * C), 256, 0, stream>>>(in, B, H, W, C, out);
  AF_LAUNCH_CHECK("sumpool2x2_kernel");
  return 0;
}
extern "C" int af_zero_insert2x(const float* in, int B, int H, int W, int C, void* out_bf16, cudaStream_t stream) {
  AF_CHECK_ARG(in && out_bf16 && B > 0 && H > 0 && W > 0 && C > 0, "af_zero_insert2x: bad arguments");
  zero_insert2x_kernel<<<grid_for(static_cast<size_t>(B) * 4 * H * W * C), 256, 0, stream>>>(in, B, H, W, C,
                                                                                             static_cast<__nv_bfloat16*>(out_bf16));
  AF_LAUNCH_CHECK("zero_insert2x_kernel");
  return 0;
}
extern "C" int af_transpose_to_bf16(const void* in, int in_dtype, int R, int C, long long ldo, void* out_bf16, cudaStream_t stream) {
  AF_CHECK_ARG(in && out_bf16 && R > 0 && C > 0 && ldo >= R, "af_transpose_to_bf16: bad arguments");
  dim3 grid((C + 31) / 32, static_cast<unsigned>((ldo + 31) / 32));
  if (in_dtype == AF_DTYPE_F32)
    transpose_to_bf16_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), R, C, ldo, static_cast<__nv_bfloat16*>(out_bf16));
  else
    transpose_to_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), R, C, ldo,
                                                                     static_cast<__nv_bfloat16*>(out_bf16));
  AF_LAUNCH_CHECK("transpose_to_bf16_kernel");
  return 0;
}

extern "C" int af_prodigy_moments(const float* p, const float* grad, const float* p0, float* s, float* exp_avg,
                                  float* exp_avg_sq, long long n, float beta1, float beta2, float beta3, float d,
                                  float s_alpha, float coupled_decay, double* sums, cudaStream_t stream) {
  AF_CHECK_ARG(p && grad && p0 && s && exp_avg && exp_avg_sq && sums && n > 0, "af_prodigy_moments: bad arguments");
  prodigy_moments_kernel<<<grid_for(static_cast<size_t>(n)), 256, 0, stream>>>(p, grad, p0, s, exp_avg, exp_avg_sq, n, beta1,
                                                                               beta2, beta3, d, s_alpha, coupled_decay, sums);
  AF_LAUNCH_CHECK("prodigy_moments_kernel");
  return 0;
}
extern "C" int af_prodigy_apply(float* p, const float* exp_avg, const float* exp_avg_sq, long long n, float dlr, float d_eps,
                                float decoupled_decay, cudaStream_t stream) {
  AF_CHECK_ARG(p && exp_avg && exp_avg_sq && n > 0, "af_prodigy_apply: bad arguments");
  prodigy_apply_kernel<<<grid_for(static_cast<size_t>(n)), 256, 0, stream>>>(p, exp_avg, exp_avg_sq, n, dlr, d_eps,
                                                                             decoupled_decay);
  AF_LAUNCH_CHECK("prodigy_apply_kernel");
  return 0;
}
