// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M, N] = A[M, K] * Wt[N, K]^T   (bf16 operands, fp32 accumulation in TMEM)
//
// One kernel template serves every contraction of the SD-1.5 UNet hot path:
//   * Linear / conv1x1 (reference: nn.Linear in ldm/modules/attention.py:35,55,157-165 and the 1x1
//     nn.Conv2d at attention.py:302,313 / openaimodel.py:245): A is a 2-D [M, K] token matrix.
//   * conv3x3 stride 1, pad 1 (openaimodel.py:208,234,530,696 and Upsample :120-122): A is the NHWC
//     activation read through a 4-D TMA map; each of the 9 taps is a shifted box and the zero padding
//     is TMA out-of-bounds fill - no im2col buffer exists anywhere.
//   * conv3x3 stride 2 (Downsample, openaimodel.py:155): the input is viewed as
//     [B, H/2, 2, W/2, 2*C] so every tap is a unit-stride 5-D box.
//   * the skip-connection concat (openaimodel.py:1019) is never materialised: the K loop walks two
//     tensor maps (A0 then A1).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused bias / time-embedding / residual / GEGLU -> HBM).
// smem ring of kStages {A 128x64, B BNx64} 128B-swizzled tiles; 2 TMEM accumulator stages so the
// epilogue of tile i overlaps the mainloop of tile i+1.
#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

struct GemmParams {
  CUtensorMap tmA0;
  CUtensorMap tmA1;
  CUtensorMap tmB;
  int M, N;            // output rows / packed output columns
  int num_kb;          // K blocks (64 wide) in total
  int cpb;             // K blocks per tap (== num_kb for linear)
  int kb_split;        // K blocks per tap served by A0 (rest by A1)
  int amode;           // 0 linear, 1 conv3x3 s1, 2 conv3x3 s2
  int B, H, W;         // conv: output pixel grid
  int C0;              // conv s2: channels of the (single) source
  int bw, bh, nb;      // conv: tile box (bw*bh*nb == 128)
  int tiles_w, tiles_h;
  int m_tiles, n_tiles;
  const float* bias;       // [N] or null
  const float* rowbias;    // [groups, N] or null; group = out_row / rows_per_group
  int rows_per_group;
  long long ld_rowbias;
  const float* residual;   // [M, ldr] fp32 or null
  long long ldr;
  void* out;
  long long ldo;
  int out_bf16;
  int geglu;               // tile cols [0,BN/2) = value, [BN/2,BN) = gate; writes BN/2 cols
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN <= 64 ? 8 : (BN <= 160 ? 6 : 4);
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kAccStride = BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmA1);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int n0 = n_tile * BN;
        int cw = 0, ch = 0, cn = 0;
        if (p.amode != 0) {
          cw = (m_tile % p.tiles_w) * p.bw;
          ch = ((m_tile / p.tiles_w) % p.tiles_h) * p.bh;
          cn = (m_tile / (p.tiles_w * p.tiles_h)) * p.nb;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int tap = kb / p.cpb;
          const int cblk = kb - tap * p.cpb;
          const bool second = cblk >= p.kb_split;
          const CUtensorMap* tmA = second ? &p.tmA1 : &p.tmA0;
          const int kc = (second ? cblk - p.kb_split : cblk) * 64;
          if (p.amode == 0) {
            tma_load_2d(sa, tmA, &full_bar[stage], kc, m_tile * 128);
          } else if (p.amode == 1) {
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d(sa, tmA, &full_bar[stage], kc, cw + kx - 1, ch + ky - 1, cn);
          } else {
            const int ky = tap / 3, kx = tap - ky * 3;
            // input row 2*oh + ky - 1  ->  (coarse row oh + dh, parity ph)
            const int ph = (ky == 1) ? 0 : 1, dh = (ky == 0) ? -1 : 0;
            const int pw = (kx == 1) ? 0 : 1, dw = (kx == 0) ? -1 : 0;
            tma_load_5d(sa, tmA, &full_bar[stage], pw * p.C0 + kc, cw + dw, ph, ch + dh, cn);
          }
          tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * 64, n0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes (16 bf16) along K inside the 128B swizzle atom == +2 in the >>4 address field
            tc_mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      long long grow;
      bool valid;
      if (p.amode == 0) {
        grow = static_cast<long long>(m_tile) * 128 + r;
        valid = grow < p.M;
      } else {
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int rw = r % p.bw;
        const int rh = (r / p.bw) % p.bh;
        const int rn = r / (p.bw * p.bh);
        const int n = tn * p.nb + rn, h = th * p.bh + rh, w = tw * p.bw + rw;
        valid = n < p.B && h < p.H && w < p.W;
        grow = (static_cast<long long>(n) * p.H + h) * p.W + w;
      }
      const float* rb = nullptr;
      if (p.rowbias && valid) rb = p.rowbias + (grow / p.rows_per_group) * p.ld_rowbias;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);

      if (p.geglu) {
        constexpr int HALF = BN / 2;
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
          uint32_t v[32], g[32];
          tmem_ld32(t_acc + c, v);
          tmem_ld32(t_acc + HALF + c, g);
          tmem_ld_wait();
          const int pc = n_tile * BN + c;                  // packed column of the value half
          const int oc = n_tile * HALF + c;                // output column
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (pc + j >= p.N) break;
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float xv = __uint_as_float(v[j + e]);
                float xg = __uint_as_float(g[j + e]);
                if (p.bias) {
                  xv += __ldg(p.bias + pc + j + e);
                  xg += __ldg(p.bias + pc + HALF + j + e);
                }
                o[e] = xv * gelu_erf_f(xg);
              }
              uint4 pk;
              pk.x = pack_bf16x2(o[0], o[1]);
              pk.y = pack_bf16x2(o[2], o[3]);
              pk.z = pack_bf16x2(o[4], o[5]);
              pk.w = pack_bf16x2(o[6], o[7]);
              *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + grow * p.ldo + oc + j) = pk;
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_acc + c, v);
          tmem_ld_wait();
          const int col0 = n_tile * BN + c;
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const int col = col0 + j;
              if (col >= p.N) break;
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[j + e]);
              if (p.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
                o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
              }
              if (rb) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + col + 4));
                o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
                o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
              }
              if (p.residual) {
                const float* rp = p.residual + grow * p.ldr + col;
                const float4 r0 = *reinterpret_cast<const float4*>(rp);
                const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
                o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
                o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
              }
              if (p.out_bf16) {
                uint4 pk;
                pk.x = pack_bf16x2(o[0], o[1]);
                pk.y = pack_bf16x2(o[2], o[3]);
                pk.z = pack_bf16x2(o[4], o[5]);
                pk.w = pack_bf16x2(o[6], o[7]);
                *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + grow * p.ldo + col) = pk;
              } else {
                float* op = static_cast<float*>(p.out) + grow * p.ldo + col;
                *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc_kernel<BN><<<grid, 192, Cfg::kSmemBytes, stream>>>(p);
  AF_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

static int dispatch_gemm(int bn, const GemmParams& p, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_gemm<64>(p, stream);
    case 128: return launch_gemm<128>(p, stream);
    case 160: return launch_gemm<160>(p, stream);
    case 256: return launch_gemm<256>(p, stream);
    default: set_error("unsupported BN %d (64/128/160/256)", bn); return -1;
  }
}

static int pick_bn(int N, int geglu, int bn_hint) {
  if (geglu) return 256;
  if (bn_hint == 64 || bn_hint == 128 || bn_hint == 160 || bn_hint == 256) return bn_hint;
  if (N % 160 == 0) return 160;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 160;
}

static int fill_epilogue(GemmParams& p, const af_epilogue* ep, long long default_ldo) {
  p.bias = ep->bias;
  p.rowbias = ep->rowbias;
  p.rows_per_group = ep->rows_per_group > 0 ? ep->rows_per_group : 1;
  p.ld_rowbias = ep->ld_rowbias > 0 ? ep->ld_rowbias : p.N;
  p.residual = ep->residual;
  p.ldr = ep->ldr > 0 ? ep->ldr : default_ldo;
  p.out = ep->out;
  p.ldo = ep->ldo > 0 ? ep->ldo : default_ldo;
  p.out_bf16 = ep->out_dtype == AF_DTYPE_BF16;
  p.geglu = ep->geglu;
  AF_CHECK_ARG(ep->out != nullptr, "epilogue: out is null");
  AF_CHECK_ARG(ep->out_dtype == AF_DTYPE_BF16 || ep->out_dtype == AF_DTYPE_F32, "epilogue: bad out dtype %d",
               ep->out_dtype);
  AF_CHECK_ARG(!ep->geglu || ep->out_dtype == AF_DTYPE_BF16, "geglu epilogue writes bf16 only");
  AF_CHECK_ARG(!ep->geglu || (!ep->residual && !ep->rowbias), "geglu epilogue: residual / rowbias unsupported");
  AF_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->out) & 15) == 0, "epilogue: out not 16B aligned");
  AF_CHECK_ARG(p.ldo % 8 == 0 && p.ldr % 4 == 0 && p.ld_rowbias % 4 == 0, "epilogue: ldo %lld / ldr %lld / ld_rowbias %lld misaligned", p.ldo, p.ldr, p.ld_rowbias);
  return 0;
}

}  // namespace af

using namespace af;

extern "C" int af_gemm_bf16(const void* A0, long long lda0, int K0, const void* A1, long long lda1, int K1,
                            const void* Wt, int M, int N, const af_epilogue* ep, int bn_hint, cudaStream_t stream) {
  AF_CHECK_ARG(A0 && Wt && ep, "af_gemm_bf16: null pointer");
  AF_CHECK_ARG(M > 0 && N > 0 && K0 > 0 && K1 >= 0, "af_gemm_bf16: bad sizes M=%d N=%d K0=%d K1=%d", M, N, K0, K1);
  AF_CHECK_ARG(N % 8 == 0, "af_gemm_bf16: N=%d must be a multiple of 8", N);
  AF_CHECK_ARG(K1 == 0 || K0 % 64 == 0, "af_gemm_bf16: K0=%d must be a multiple of 64 for a dual-source A", K0);
  AF_CHECK_ARG((K0 + K1) % 8 == 0 && lda0 % 8 == 0 && (K1 == 0 || lda1 % 8 == 0),
               "af_gemm_bf16: K / lda must be multiples of 8 (16-byte TMA strides)");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int K = K0 + K1;
  const int bn = pick_bn(N, ep->geglu, bn_hint);
  AF_CHECK_ARG(!ep->geglu || N % 256 == 0, "geglu: packed N=%d must be a multiple of 256", N);
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K0), static_cast<uint64_t>(M)};
    uint64_t str[1] = {static_cast<uint64_t>(lda0) * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&p.tmA0, A0, 2, dims, str, box);
    if (rc) return rc;
    p.tmA1 = p.tmA0;
  }
  if (K1 > 0) {
    AF_CHECK_ARG(A1 != nullptr, "af_gemm_bf16: A1 null with K1=%d", K1);
    uint64_t dims[2] = {static_cast<uint64_t>(K1), static_cast<uint64_t>(M)};
    uint64_t str[1] = {static_cast<uint64_t>(lda1) * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&p.tmA1, A1, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
    int rc = make_tmap_bf16(&p.tmB, Wt, 2, dims, str, box);
    if (rc) return rc;
  }
  p.M = M;
  p.N = N;
  p.num_kb = (K0 + 63) / 64 + (K1 + 63) / 64;
  p.cpb = p.num_kb;
  p.kb_split = (K0 + 63) / 64;
  p.amode = 0;
  p.m_tiles = (M + 127) / 128;
  p.n_tiles = (N + bn - 1) / bn;
  int rc = fill_epilogue(p, ep, ep->geglu ? N / 2 : N);
  if (rc) return rc;
  return dispatch_gemm(bn, p, stream);
}

extern "C" int af_conv3x3_bf16(const void* X0, int C0, const void* X1, int C1, const void* Wt, int B, int H, int W,
                               int Cout, int stride, const af_epilogue* ep, int bn_hint, cudaStream_t stream) {
  AF_CHECK_ARG(X0 && Wt && ep, "af_conv3x3_bf16: null pointer");
  AF_CHECK_ARG(stride == 1 || stride == 2, "af_conv3x3_bf16: stride %d", stride);
  AF_CHECK_ARG(C0 > 0 && C0 % 64 == 0 && C1 >= 0 && C1 % 64 == 0, "af_conv3x3_bf16: C0=%d C1=%d must be multiples of 64",
               C0, C1);
  AF_CHECK_ARG(Cout % 8 == 0, "af_conv3x3_bf16: Cout=%d must be a multiple of 8", Cout);
  AF_CHECK_ARG(stride == 1 || (C1 == 0 && H % 2 == 0 && W % 2 == 0), "af_conv3x3_bf16: stride 2 needs single source, even H/W");
  AF_CHECK_ARG(!ep->geglu, "af_conv3x3_bf16: geglu epilogue unsupported");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int Cin = C0 + C1;
  const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
  const int bn = pick_bn(Cout, 0, bn_hint);
  // tile box: bw = largest power of two <= min(64, next_pow2(Wo)); rows/images fill up to 128 pixels
  int bw = 1;
  while (bw < Wo && bw < 64) bw <<= 1;
  int bh = 1;
  while (bh < Ho && bw * bh < 128) bh <<= 1;
  int nb = 128 / (bw * bh);
  p.bw = bw; p.bh = bh; p.nb = nb;
  p.tiles_w = (Wo + bw - 1) / bw;
  p.tiles_h = (Ho + bh - 1) / bh;
  p.B = B; p.H = Ho; p.W = Wo; p.C0 = C0;
  const uint64_t es = 2;
  if (stride == 1) {
    uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(nb)};
    {
      uint64_t dims[4] = {static_cast<uint64_t>(C0), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
      uint64_t str[3] = {C0 * es, static_cast<uint64_t>(W) * C0 * es, static_cast<uint64_t>(H) * W * C0 * es};
      int rc = make_tmap_bf16(&p.tmA0, X0, 4, dims, str, box);
      if (rc) return rc;
      p.tmA1 = p.tmA0;
    }
    if (C1 > 0) {
      AF_CHECK_ARG(X1 != nullptr, "af_conv3x3_bf16: X1 null with C1=%d", C1);
      uint64_t dims[4] = {static_cast<uint64_t>(C1), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
      uint64_t str[3] = {C1 * es, static_cast<uint64_t>(W) * C1 * es, static_cast<uint64_t>(H) * W * C1 * es};
      int rc = make_tmap_bf16(&p.tmA1, X1, 4, dims, str, box);
      if (rc) return rc;
    }
    p.amode = 1;
  } else {
    uint32_t box[5] = {64, static_cast<uint32_t>(bw), 1, static_cast<uint32_t>(bh), static_cast<uint32_t>(nb)};
    uint64_t dims[5] = {static_cast<uint64_t>(2 * C0), static_cast<uint64_t>(W / 2), 2, static_cast<uint64_t>(H / 2),
                        static_cast<uint64_t>(B)};
    uint64_t str[4] = {2 * C0 * es, static_cast<uint64_t>(W) * C0 * es, 2 * static_cast<uint64_t>(W) * C0 * es,
                       static_cast<uint64_t>(H) * W * C0 * es};
    int rc = make_tmap_bf16(&p.tmA0, X0, 5, dims, str, box);
    if (rc) return rc;
    p.tmA1 = p.tmA0;
    p.amode = 2;
  }
  {
    const uint64_t K = 9ull * Cin;
    uint64_t dims[2] = {K, static_cast<uint64_t>(Cout)};
    uint64_t str[1] = {K * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(bn)};
    int rc = make_tmap_bf16(&p.tmB, Wt, 2, dims, str, box);
    if (rc) return rc;
  }
  p.M = B * Ho * Wo;
  p.N = Cout;
  p.cpb = Cin / 64;
  p.kb_split = C0 / 64;
  p.num_kb = 9 * p.cpb;
  p.m_tiles = p.tiles_w * p.tiles_h * ((B + nb - 1) / nb);
  p.n_tiles = (Cout + bn - 1) / bn;
  int rc = fill_epilogue(p, ep, Cout);
  if (rc) return rc;
  if (p.rowbias && ep->rows_per_group <= 0) p.rows_per_group = Ho * Wo;
  return dispatch_gemm(bn, p, stream);
}
