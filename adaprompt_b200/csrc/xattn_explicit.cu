// Explicit (materialised-score) cross-attention over the <= 96 prompt tokens: the two optional training-time behaviours
// of CrossAttention.forward that need the score matrix itself (ldm/modules/attention.py):
//   * save_attn_vars (:245-255): cache q * sqrt(scale), the scores after replacement (`attnscore`) and the
//     probabilities (`attn`) for the distillation losses (UNetModel.forward collects them, openaimodel.py:947-952,
//     984-988, 1031-1035);
//   * conv attention (:208-216 -> ldm/util.py:700-878 replace_rows_by_conv_attn): the score columns of the first ks*ks
//     subject tokens are replaced by a ks x ks grouped convolution of the query map with those tokens' keys, shifted per
//     token.  Since conv2d(q, k-as-weights) = sum over taps of POINTWISE scores at shifted pixels, the replacement is
//     computed from the pointwise scores of the ks*ks subject columns (af_conv_attn_scores) and fed back as an override.
// The fused kernels (xattn.cu) never materialise scores; these paths are rare, small (N x 77 per head) and run on the
// CUDA cores.  Scores here are in the reference's domain: sim = q.k * scale (Q carries scale * log2 e, divided out).
#include <math.h>
#include <algorithm>

#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

constexpr int kMaxKeys = 96;

// grid (ceil(N / 64), heads, B), block 128 (4 warps, 16 query rows each).  smem: K^T [d][nkp] fp32, V^T [d][nkp] fp32.
__global__ void __launch_bounds__(128) xattn_explicit_kernel(
    const __nv_bfloat16* __restrict__ Q, long long ldq, const __nv_bfloat16* __restrict__ K, long long ldk,
    const __nv_bfloat16* __restrict__ Vt, long long ldvt, int kv_stride, int N, int nk, int d, int dp, int heads,
    const float* __restrict__ override_scores /* [B][heads][N][n_ov] or null */, const int* __restrict__ ov_cols /* [B][n_ov] */,
    int n_ov, __nv_bfloat16* __restrict__ out /* [B*N][heads*d] or null */, float* __restrict__ attn /* [B][heads][N][nk] or null */,
    float* __restrict__ attnscore, float* __restrict__ q_out /* [B][heads][N][d] or null */) {
  extern __shared__ float sm[];
  const int nkp = (nk + 31) / 32 * 32;
  float* kT = sm;                      // [d][nkp]
  float* vT = kT + d * nkp;            // [d][nkp]
  float* rowbuf = vT + d * nkp;        // [4 warps][max(d, nkp)]
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_log2e = 0.69314718055994530942f;
  for (int i = threadIdx.x; i < d * nkp; i += 128) {
    const int c = i / nkp, j = i - c * nkp;
    kT[i] = j < nk ? __bfloat162float(K[(static_cast<size_t>(b) * kv_stride + j) * ldk + h * dp + c]) : 0.f;
    vT[i] = j < nk ? __bfloat162float(Vt[static_cast<size_t>(h * d + c) * ldvt + static_cast<size_t>(b) * kv_stride + j]) : 0.f;
  }
  __syncthreads();
  const int rb = d > nkp ? d : nkp;
  float* myrow = rowbuf + warp * rb;
  const float qscale = rsqrtf(static_cast<float>(d));          // sim scale d^-1/2
  const int row_end = min(N, (static_cast<int>(blockIdx.x) + 1) * 64);
  for (int n = blockIdx.x * 64 + warp; n < row_end; n += 4) {
    const __nv_bfloat16* qrow = Q + (static_cast<size_t>(b) * N + n) * ldq + h * dp;
    for (int c = lane; c < d; c += 32) myrow[c] = __bfloat162float(qrow[c]) * inv_log2e;   // q * scale
    __syncwarp();
    if (q_out) {
      const float f = rsqrtf(qscale);                          // (q * scale) / sqrt(scale) = q * sqrt(scale)
      float* qo = q_out + ((static_cast<size_t>(b) * heads + h) * N + n) * d;
      for (int c = lane; c < d; c += 32) qo[c] = myrow[c] * f;
    }
    float s[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int j = lane + 32 * u;
      float acc = 0.f;
      if (j < nkp)
        for (int c = 0; c < d; ++c) acc = fmaf(myrow[c], kT[c * nkp + j], acc);
      s[u] = j < nk ? acc : -INFINITY;
    }
    if (override_scores) {
      const float* ov = override_scores + ((static_cast<size_t>(b) * heads + h) * N + n) * n_ov;
      for (int m = 0; m < n_ov; ++m) {
        const int col = ov_cols[b * n_ov + m];
        if (col >= 0 && (col & 31) == lane) s[col >> 5] = ov[m];
      }
    }
    float mx = warp_max(fmaxf(s[0], fmaxf(s[1], s[2])));
    float e[3], sum = 0.f;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      e[u] = s[u] == -INFINITY ? 0.f : __expf(s[u] - mx);
      sum += e[u];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int j = lane + 32 * u;
      if (j < nkp) myrow[j] = e[u] * inv;
      if (j < nk) {
        const size_t o = ((static_cast<size_t>(b) * heads + h) * N + n) * nk + j;
        if (attn) attn[o] = e[u] * inv;
        if (attnscore) attnscore[o] = s[u];
      }
    }
    __syncwarp();
    if (out) {
      __nv_bfloat16* orow = out + (static_cast<size_t>(b) * N + n) * (static_cast<size_t>(heads) * d) + h * d;
      for (int c = lane; c < d; c += 32) {
        float acc = 0.f;
        for (int j = 0; j < nk; ++j) acc = fmaf(myrow[j], vT[c * nkp + j], acc);
        orow[c] = __float2bfloat16(acc);
      }
    }
    __syncwarp();
  }
}

// override[b][h][p][m] = (1 / ks^1.5) * sum_{i,j < ks} score[b][h][(y - dy_m + i - pad, x - dx_m + j - pad)][cols[b][ks*i+j]]
// with zero outside the feature map (ldm/util.py:741-766, 790-856); m = (dy + pad) * ks + (dx + pad), pad = (ks - 1) / 2.
__global__ void conv_attn_scores_kernel(const float* __restrict__ score /* [B][heads][N][nk] */, const int* __restrict__ cols,
                                        int B, int heads, int Hf, int Wf, int nk, int ks, float* __restrict__ override) {
  const int N = Hf * Wf, M = ks * ks;
  const size_t total = static_cast<size_t>(B) * heads * N * M;
  const int pad = (ks - 1) / 2;           // pads: ks 2 -> (0,1), 3 -> (1,1), 4 -> (1,2)
  const float inv_norm = 1.0f / powf(static_cast<float>(ks), 1.5f);
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int m = idx % M;
    const int p = (idx / M) % N;
    const int h = (idx / (static_cast<size_t>(M) * N)) % heads;
    const int b = idx / (static_cast<size_t>(M) * N * heads);
    float acc = 0.f;
    if (cols[b * M] >= 0) {
      const int dy = m / ks - pad, dx = m % ks - pad;
      const int y0 = p / Wf - dy, x0 = p % Wf - dx;     // A is read at (y0, x0); zero outside the map
      if (y0 >= 0 && y0 < Hf && x0 >= 0 && x0 < Wf) {
        const float* sb = score + (static_cast<size_t>(b) * heads + h) * N * nk;
        for (int i = 0; i < ks; ++i)
          for (int j = 0; j < ks; ++j) {
            const int yy = y0 + i - pad, xx = x0 + j - pad;
            if (yy >= 0 && yy < Hf && xx >= 0 && xx < Wf) acc += sb[static_cast<size_t>(yy * Wf + xx) * nk + cols[b * M + ks * i + j]];
          }
        acc *= inv_norm;
      }
    }
    override[idx] = acc;
  }
}

}  // namespace af

using namespace af;

extern "C" int af_xattn_explicit(const void* Q, long long ldq, const void* K, long long ldk, const void* Vt, long long ldvt,
                                 int kv_stride, const float* override_scores, const int* ov_cols, int n_ov, void* O,
                                 float* attn, float* attnscore, float* q_out, int B, int heads, int N, int nk, int d,
                                 cudaStream_t stream) {
  AF_CHECK_ARG(Q && K && Vt, "af_xattn_explicit: null pointer");
  AF_CHECK_ARG(d == 40 || d == 80 || d == 160, "af_xattn_explicit: head dim %d unsupported", d);
  AF_CHECK_ARG(B > 0 && heads > 0 && N > 0 && nk > 0 && nk <= kMaxKeys && kv_stride >= nk, "af_xattn_explicit: bad sizes (nk <= %d)", kMaxKeys);
  AF_CHECK_ARG((override_scores == nullptr) == (ov_cols == nullptr) && (override_scores == nullptr || n_ov > 0),
               "af_xattn_explicit: override needs scores, columns and a count");
  const int dp = d == 40 ? 48 : d;
  const int nkp = (nk + 31) / 32 * 32;
  const size_t smem = (2 * static_cast<size_t>(d) * nkp + 4 * static_cast<size_t>(d > nkp ? d : nkp)) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(xattn_explicit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  dim3 grid((N + 63) / 64, heads, B);
  xattn_explicit_kernel<<<grid, 128, smem, stream>>>(static_cast<const __nv_bfloat16*>(Q), ldq, static_cast<const __nv_bfloat16*>(K), ldk,
                                                      static_cast<const __nv_bfloat16*>(Vt), ldvt, kv_stride, N, nk, d, dp, heads,
                                                      override_scores, ov_cols, n_ov, static_cast<__nv_bfloat16*>(O), attn, attnscore,
                                                      q_out);
  AF_LAUNCH_CHECK("xattn_explicit_kernel");
  return 0;
}

extern "C" int af_conv_attn_scores(const float* score, const int* cols, int B, int heads, int Hf, int Wf, int nk, int ks,
                                   float* override_scores, cudaStream_t stream) {
  AF_CHECK_ARG(score && cols && override_scores, "af_conv_attn_scores: null pointer");
  AF_CHECK_ARG(ks >= 2 && ks <= 4 && B > 0 && heads > 0 && Hf > 0 && Wf > 0 && nk > 0, "af_conv_attn_scores: bad sizes (ks in 2..4)");
  const size_t total = static_cast<size_t>(B) * heads * Hf * Wf * ks * ks;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  conv_attn_scores_kernel<<<blocks, 256, 0, stream>>>(score, cols, B, heads, Hf, Wf, nk, ks, override_scores);
  AF_LAUNCH_CHECK("conv_attn_scores_kernel");
  return 0;
}
