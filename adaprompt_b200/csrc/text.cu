// Small fused kernels of the conditioning path (CLIP text transformer + AdaFace token splicing).
//
// Reference call sites:
//   * CLIP self-attention, 12 heads x 64, causal, seq <= 77 (x m keys per token for CLIPAttentionMKV):
//     adaface/arc2face_models.py:87-173 (MKV) and HF CLIPAttention driven from arc2face_models.py:220,
//     ldm/modules/encoders/modules.py:264.
//   * token-embedding lookup + position add: CLIPTextEmbeddings.forward as patched at modules.py:195-223.
//   * placeholder search / splice: ldm/modules/embedding_manager.py:1359,1368,1561-1562 and adaface/util.py:107,184
//     (pure index work and row copies: bit-exact by construction).
//   * weighted sum of the last hidden states: arc2face_models.py:236-246, modules.py:361-368.
#include <math.h>

#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

// ---------------------------------------------------------------------------------------------
// Small-sequence attention.  qkv bf16 [B*L, ldq] with q at column 0, k at k_off, v at v_off; head h of q at
// h*64; the k / v projections may carry `mult` keys per token (MKV): key index j = token*mult + r lives at
// column k_off + (h*mult + r)*64 of row `token` (arc2face_models.py:117-131).  Causal: key token <= query token.
// One block per (head, sample); K / V staged in shared memory as fp32; one warp per query row.
// ---------------------------------------------------------------------------------------------
constexpr int kTD = 64;  // CLIP head dim

__global__ void __launch_bounds__(256) attention_small_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq,
                                                              int k_off, int v_off, int L, int mult, int heads,
                                                              float scale, int causal,
                                                              __nv_bfloat16* __restrict__ out, long long ldo) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const int nk = L * mult;
  float* sk = sm;                       // [nk][kTD + 1]
  float* sv = sk + nk * (kTD + 1);      // [nk][kTD]
  float* sp = sv + nk * kTD;            // [warps][nk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < nk * kTD; i += blockDim.x) {
    const int j = i / kTD, c = i - j * kTD;
    const int tok = j / mult, r = j - tok * mult;
    const __nv_bfloat16* row = qkv + (static_cast<size_t>(b) * L + tok) * ldq;
    sk[j * (kTD + 1) + c] = __bfloat162float(row[k_off + (h * mult + r) * kTD + c]);
    sv[j * kTD + c] = __bfloat162float(row[v_off + (h * mult + r) * kTD + c]);
  }
  __syncthreads();
  float* pw = sp + warp * nk;
  float* sq = sp + nwarps * nk + warp * kTD;       // this warp's query row (scaled)
  // queries are split over the warps of the block AND over gridDim.z blocks (each block stages K / V of its head again:
  // 40 KB from L2); scores: one KEY per lane and 64 FMAs per key - the first version put the 64 channels on the lanes
  // and paid a 5-step shuffle reduction per (query, key) pair, 68 us for a 77-token layer
  for (int i = blockIdx.z * nwarps + warp; i < L; i += nwarps * gridDim.z) {
    const __nv_bfloat16* qrow = qkv + (static_cast<size_t>(b) * L + i) * ldq + h * kTD;
    sq[lane] = __bfloat162float(qrow[lane]) * scale;
    sq[lane + 32] = __bfloat162float(qrow[lane + 32]) * scale;
    __syncwarp();
    const int kmax = causal ? (i + 1) * mult : nk;
    float mx = -INFINITY;
    for (int j = lane; j < kmax; j += 32) {
      const float* kr = sk + j * (kTD + 1);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int c = 0; c < kTD; c += 4) {
        d0 = fmaf(sq[c], kr[c], d0);
        d1 = fmaf(sq[c + 1], kr[c + 1], d1);
        d2 = fmaf(sq[c + 2], kr[c + 2], d2);
        d3 = fmaf(sq[c + 3], kr[c + 3], d3);
      }
      const float d = (d0 + d1) + (d2 + d3);
      pw[j] = d;
      mx = fmaxf(mx, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < kmax; j += 32) {
      const float e = __expf(pw[j] - mx);
      pw[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < kmax; ++j) {
      const float pj = pw[j];
      o0 += pj * sv[j * kTD + lane];
      o1 += pj * sv[j * kTD + lane + 32];
    }
    const float inv = 1.0f / sum;
    __nv_bfloat16* orow = out + (static_cast<size_t>(b) * L + i) * ldo + h * kTD;
    orow[lane] = __float2bfloat16(o0 * inv);
    orow[lane + 32] = __float2bfloat16(o1 * inv);
    __syncwarp();
  }
}

// out[r, :] = table[ids[r], :]  (fp32 rows, exact copies)
__global__ void gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ ids,
                                   float* __restrict__ out, long long rows, int dim, int vocab) {
  const int dq = dim >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * dq;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dq;
    const int c = static_cast<int>(i - r * dq);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(table + id * dim) + c);
  }
}

// x[b, t, :] += pos[t, :]
__global__ void add_pos_kernel(float* __restrict__ x, const float* __restrict__ pos, long long rows, int L, int dim) {
  const int dq = dim >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * dq;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dq;
    const int c = static_cast<int>(i - r * dq);
    const int t = static_cast<int>(r % L);
    float4 v = reinterpret_cast<float4*>(x)[i];
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * dim) + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

// first[r] = first position of `token` in ids[r, :], or -1
__global__ void find_first_token_kernel(const long long* __restrict__ ids, int rows, int L, long long token,
                                        int* __restrict__ first) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int pos = -1;
  for (int t = 0; t < L; ++t)
    if (ids[static_cast<size_t>(r) * L + t] == token) {
      pos = t;
      break;
    }
  first[r] = pos;
}

// dst[r, start[r] + k, :] = src[src_row(r), k, :] for k < K when start[r] >= 0 (exact row copies).
// src_row(r) = (r / rep_outer % src_mod) ... expressed by the caller through src_index[r].
__global__ void splice_rows_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                   const int* __restrict__ start, const int* __restrict__ src_index, int rows, int L,
                                   int K, int dim) {
  const int dq = dim >> 2;
  const long long total = static_cast<long long>(rows) * K * dq;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % dq);
    const int k = static_cast<int>((i / dq) % K);
    const int r = static_cast<int>(i / (static_cast<long long>(dq) * K));
    const int s = start[r];
    if (s < 0 || s + k >= L) continue;
    const int sr = src_index ? src_index[r] : r;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (static_cast<size_t>(sr) * K + k) * dim) + c);
    reinterpret_cast<float4*>(dst + (static_cast<size_t>(r) * L + s + k) * dim)[c] = v;
  }
}

// out = w0*a + w1*b (+ w2*c): hidden-state mixing before the final LayerNorm
__global__ void weighted_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                    const float* __restrict__ c, float w0, float w1, float w2,
                                    float* __restrict__ out, size_t n4) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    // same association as (stack * w).sum(0): ((w0*a + w1*b) + w2*c)
    float4 o = make_float4(__fadd_rn(__fmul_rn(w0, x.x), __fmul_rn(w1, y.x)), __fadd_rn(__fmul_rn(w0, x.y), __fmul_rn(w1, y.y)),
                           __fadd_rn(__fmul_rn(w0, x.z), __fmul_rn(w1, y.z)), __fadd_rn(__fmul_rn(w0, x.w), __fmul_rn(w1, y.w)));
    if (c) {
      const float4 z = reinterpret_cast<const float4*>(c)[i];
      o.x = __fadd_rn(o.x, __fmul_rn(w2, z.x)); o.y = __fadd_rn(o.y, __fmul_rn(w2, z.y));
      o.z = __fadd_rn(o.z, __fmul_rn(w2, z.z)); o.w = __fadd_rn(o.w, __fmul_rn(w2, z.w));
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

static inline int blocks_for(long long items) {
  long long g = (items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace af

using namespace af;

extern "C" int af_attention_small(const void* qkv, long long ldq, int k_off, int v_off, void* out, long long ldo, int B,
                                  int heads, int L, int mult, float scale, int causal, cudaStream_t stream) {
  AF_CHECK_ARG(qkv && out, "af_attention_small: null pointer");
  AF_CHECK_ARG(B > 0 && heads > 0 && L > 0 && mult >= 1, "af_attention_small: bad sizes");
  const int nk = L * mult;
  const size_t smem = (static_cast<size_t>(nk) * (kTD + 1) + static_cast<size_t>(nk) * kTD + 8 * nk + 8 * kTD) * sizeof(float);
  AF_CHECK_ARG(smem <= 200 * 1024, "af_attention_small: L*mult=%d too large", nk);
  static size_t configured = 0;
  if (smem > configured) {
    AF_CUDA(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    configured = smem;
  }
  int qsplit = (2 * num_sms()) / (heads * B);       // fill the machine: heads x B blocks alone leave most SMs idle
  if (qsplit > (L + 7) / 8) qsplit = (L + 7) / 8;
  if (qsplit < 1) qsplit = 1;
  attention_small_kernel<<<dim3(heads, B, qsplit), 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(qkv), ldq, k_off, v_off,
                                                                L, mult, heads, scale, causal,
                                                                static_cast<__nv_bfloat16*>(out), ldo);
  AF_LAUNCH_CHECK("attention_small_kernel");
  return 0;
}

extern "C" int af_gather_rows(const float* table, const long long* ids, float* out, long long rows, int dim, int vocab,
                              cudaStream_t stream) {
  AF_CHECK_ARG(table && ids && out && rows > 0 && dim % 4 == 0 && vocab > 0, "af_gather_rows: bad args");
  gather_rows_kernel<<<blocks_for(rows * (dim / 4)), 256, 0, stream>>>(table, ids, out, rows, dim, vocab);
  AF_LAUNCH_CHECK("gather_rows_kernel");
  return 0;
}

extern "C" int af_add_pos(float* x, const float* pos, long long rows, int L, int dim, cudaStream_t stream) {
  AF_CHECK_ARG(x && pos && rows > 0 && L > 0 && dim % 4 == 0, "af_add_pos: bad args");
  add_pos_kernel<<<blocks_for(rows * (dim / 4)), 256, 0, stream>>>(x, pos, rows, L, dim);
  AF_LAUNCH_CHECK("add_pos_kernel");
  return 0;
}

extern "C" int af_find_first_token(const long long* ids, int rows, int L, long long token, int* first,
                                   cudaStream_t stream) {
  AF_CHECK_ARG(ids && first && rows > 0 && L > 0, "af_find_first_token: bad args");
  find_first_token_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(ids, rows, L, token, first);
  AF_LAUNCH_CHECK("find_first_token_kernel");
  return 0;
}

extern "C" int af_splice_rows(float* dst, const float* src, const int* start, const int* src_index, int rows, int L,
                              int K, int dim, cudaStream_t stream) {
  AF_CHECK_ARG(dst && src && start && rows > 0 && L > 0 && K > 0 && dim % 4 == 0, "af_splice_rows: bad args");
  splice_rows_kernel<<<blocks_for(static_cast<long long>(rows) * K * (dim / 4)), 256, 0, stream>>>(
      dst, src, start, src_index, rows, L, K, dim);
  AF_LAUNCH_CHECK("splice_rows_kernel");
  return 0;
}

extern "C" int af_weighted_sum(const float* a, const float* b, const float* c, float w0, float w1, float w2, float* out,
                               long long n, cudaStream_t stream) {
  AF_CHECK_ARG(a && b && out && n > 0 && n % 4 == 0, "af_weighted_sum: bad args");
  weighted_sum_kernel<<<blocks_for(n / 4), 256, 0, stream>>>(a, b, c, w0, w1, w2, out, static_cast<size_t>(n / 4));
  AF_LAUNCH_CHECK("weighted_sum_kernel");
  return 0;
}
