/*
 * adaface_b200.h - C ABI of the B200-native AdaFace denoising hot path (libadaface_b200.so).
 *
 * The reference (askerlee/adaprompt) is 100 % Python and has no FFI layer for this path
 * (SURVEY.md section 8(b)); every op below replaces a *PyTorch library call site* inside the
 * reference modules, cited per function.  The Python host mirror of those modules
 * (adaprompt_b200/{unet,attention,ddim,...}.py) binds these symbols with ctypes - see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / ATen types; all pointers are DEVICE pointers
 *     unless a parameter name says host;
 *   - the library never allocates or frees device memory and keeps no state besides
 *     compiled-in tables: the caller (PyTorch) owns inputs, outputs, workspaces, KV caches;
 *   - every launch goes to the caller's stream and is CUDA-graph capturable (no host sync);
 *   - return value: 0 = ok, < 0 = bad argument / unsupported shape, > 0 = cudaError_t;
 *     af_last_error() returns the message (thread local);
 *   - activations are NHWC ("tokens x channels"); the residual stream is fp32, GEMM operands
 *     are bf16, accumulation is fp32 in TMEM.
 *   - sm_100a only.  There is no CPU fallback.
 */
#ifndef ADAFACE_B200_H
#define ADAFACE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AF_VERSION 100
#define AF_DTYPE_F32 0
#define AF_DTYPE_BF16 1
#define AF_GN_MAX_CHUNKS 64

typedef struct CUstream_st* af_stream_t; /* == cudaStream_t */

int af_version(void);
const char* af_last_error(void);
int af_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Fused GEMM epilogue:  out = [geglu]( acc + bias[col] + rowbias[row / rows_per_group, col] + residual[row, col] ) */
typedef struct af_epilogue {
  const float* bias;      /* [N] or NULL (nn.Linear / nn.Conv2d bias) */
  const float* rowbias;   /* [groups, N] or NULL: per-sample time-embedding add, openaimodel.py:268-277 */
  int rows_per_group;     /* rows sharing one rowbias row (H*W); <= 0: derived for conv */
  long long ld_rowbias;   /* row pitch of rowbias in floats; <= 0: N */
  const float* residual;  /* [M, ldr] fp32 or NULL: residual adds attention.py:277,281,283,341 / openaimodel.py:279 */
  long long ldr;          /* <= 0: same as ldo */
  void* out;              /* [M, ldo] */
  long long ldo;          /* row pitch in elements; <= 0: N (N/2 for geglu) */
  int out_dtype;          /* AF_DTYPE_F32 | AF_DTYPE_BF16 */
  int geglu;              /* 1: GEGLU (attention.py:32-39): weight rows packed per 256-col tile as [128 value | 128 gate];
                             out[:, j] = (x.Wv_j + bv_j) * gelu_erf(x.Wg_j + bg_j), N/2 bf16 output columns */
} af_epilogue;

/* D[M,N] = [A0 | A1][M, K0+K1] . Wt[N, K0+K1]^T, bf16 operands (row-major, K contiguous), fp32 accumulate.
 * Replaces nn.Linear (attention.py:35,55,157-165) and 1x1 nn.Conv2d (attention.py:302,313; openaimodel.py:245);
 * the optional second source A1 is the skip-connection concat of openaimodel.py:1019 read in place.
 * bn_hint: 0 = auto, or 64/128/160/256 (N tile). */
int af_gemm_bf16(const void* A0, long long lda0, int K0, const void* A1, long long lda1, int K1, const void* Wt,
                 int M, int N, const af_epilogue* ep, int bn_hint, af_stream_t stream);

/* 3x3 convolution, pad 1, stride 1 or 2, NHWC bf16 input(s) [B,H,W,C0] (+ [B,H,W,C1] concat), weights
 * Wt[Cout][ky][kx][C0+C1] bf16, as an implicit GEMM (no im2col buffer).  Output rows are output pixels
 * (n, oh, ow) row-major.  Replaces nn.Conv2d 3x3 at openaimodel.py:155 (stride 2), :208, :234, :120-122. */
int af_conv3x3_bf16(const void* X0, int C0, const void* X1, int C1, const void* Wt, int B, int H, int W, int Cout,
                    int stride, const af_epilogue* ep, int bn_hint, af_stream_t stream);

/* Flash attention, 8-head SD-1.5 geometry (d in {40,80,160}); replaces attention.py:198-242.
 * Q [B,Nq,ldq], K [B,Nk,ldk] bf16 with head h at column h*DP (DP = 48 for d = 40, zero padded, else d);
 * Q pre-scaled by d^-1/2 * log2(e).  Vt [heads*d, ldvt] bf16 = V transposed.  Sample b's keys start at row
 * b*kv_stride of K and at column b*kv_stride of Vt (kv_stride >= Nk; the cached 77-token context uses 80).
 * key_mask [B,Nk] bytes (1 = attend) or NULL (attention.py:223-232).  O [B,Nq,heads*d] bf16. */
int af_attention_bf16(const void* Q, long long ldq, const void* K, long long ldk, const void* Vt, long long ldvt,
                      int kv_stride, const unsigned char* key_mask, void* O, int B, int heads, int Nq, int Nk, int d,
                      af_stream_t stream);

/* GroupNorm(32) over the channel concat [x0 | x1] of fp32 NHWC tensors, optional SiLU, bf16 output
 * [B, HW, C0+C1]; optional raw bf16 copy of the concat (operand of the 1x1 skip conv).
 * Replaces GroupNorm32+SiLU (util.py:217-219; openaimodel.py:205-207,229-231,693-695) and Normalize
 * (attention.py:71-72,325).  workspace: af_groupnorm_workspace_bytes(B) bytes. */
size_t af_groupnorm_workspace_bytes(int B);
int af_groupnorm_silu(const float* x0, int C0, const float* x1, int C1, int B, int HW, const float* gamma,
                      const float* beta, float eps, int silu, void* y_bf16, void* raw_bf16, float* workspace,
                      af_stream_t stream);

/* nn.LayerNorm(C) (attention.py:267-269) on fp32 rows -> bf16. */
int af_layernorm(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps, void* y_bf16,
                 af_stream_t stream);

/* First / last convolutions of the UNet (4 latent channels), fp32 CUDA-core kernels that also convert
 * between the public NCHW layout and the internal NHWC one (openaimodel.py:527-533, :693-697). */
int af_conv_in(const float* x_nchw, const float* w /*[Cout,Cin,3,3]*/, const float* bias, float* y_nhwc, int B, int Cin,
               int H, int W, int Cout, af_stream_t stream);
int af_conv_out(const void* x_nhwc_bf16, const float* w_packed /*[Cout,3,3,C]*/, const float* bias, float* y_nchw,
                int B, int H, int W, int C, int Cout, af_stream_t stream);

/* timestep_embedding (util.py:154-174): out[b] = [cos(t_b f) | sin(t_b f)], fp32. */
int af_timestep_embedding(const float* t, float* out, int B, int dim, af_stream_t stream);

/* y[M,N] = silu_out?( silu_in?(x)[M,K] . W[N,K]^T + b ), fp32, small M (time_embed, emb_layers). */
int af_linear_small(const float* x, const float* w, const float* bias, float* y, int M, int N, int K, int silu_in,
                    int silu_out, af_stream_t stream);

int af_cast_bf16(const float* x, void* y_bf16, long long n, af_stream_t stream);
/* F.interpolate(scale_factor=2, mode="nearest") (openaimodel.py:120) fused with the bf16 cast. */
int af_upsample2x_cast(const float* x_nhwc, void* y_bf16, int B, int H, int W, int C, af_stream_t stream);

/* CFG combine + DDIM update (ddim.py:260,279,283,295), reference fp32 operation order.
 * coef_table rows of 8 floats [g, sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma*temp, 0, 0];
 * row = *step_idx (device) or 0 when step_idx is NULL.  eps = [cond ; uncond] when has_uncond. */
int af_cfg_ddim_update(const float* x, const float* eps, int has_uncond, const float* noise, const float* coef_table,
                       const int* step_idx, float* x_prev, float* pred_x0, long long n, af_stream_t stream);
int af_advance_step(int* step_idx, const float* t_table, float* t_buf, int B, int num_steps, af_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAFACE_B200_H */
