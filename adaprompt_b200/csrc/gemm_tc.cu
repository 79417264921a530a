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
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> smem transpose -> fused bias / time-embedding / residual /
// GEGLU -> coalesced HBM stores).
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
  int act;                 // 0 none, 1 quick_gelu x*sigmoid(1.702x) on (acc + bias), before the residual add
  float* gn_stats;         // [slots_total][N][2] per-channel (sum, sumsq) over 32-row quarters of the output, or null
  int gn_slots;            // conv: stat slots per sample (= tiles_w * tiles_h * bw * bh / 32)
};

constexpr int kEpiWarps = 8;                    // 2 warps per TMEM lane quarter, each takes every other 32-col chunk
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kStagingBytes = kEpiWarps * 4096;  // 32 rows x 128 B per epilogue warp

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN <= 64 ? 8 : (BN <= 128 ? 6 : (BN <= 160 ? 5 : 4));
  static constexpr int kAccStride = BN <= 128 ? 128 : 256;
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// GELU for the GEGLU epilogue: x * Phi(x) in its tanh form, 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), one MUFU
// op per element.  |tanh form - erf form| <= 4.8e-4 absolute (rel-L2 2e-4 for unit-variance gates), an order of
// magnitude below the bf16 rounding of the product this epilogue writes (rel-L2 1.7e-3); the erf form cost 3x the
// instructions and left the tensor pipe idle 85 % of the time (profiles/r01_gemm_geglu_before.md).
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = x * x;
  const float u = x * fmaf(0.0356774081f, x2, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// QuickGELU of the CLIP text MLP (transformers QuickGELUActivation): x * sigmoid(1.702 x)
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }

template <int BN, bool GEGLU>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
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
      mbar_init(&tempty_bar[s], kEpiWarps);
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
    // TMEM is read row-per-thread (lane == row).  Global traffic is made coalesced by a transpose through a
    // swizzled per-warp shared-memory slab.
    const int ew = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int half = ew >> 2;        // which 32-column chunks of the tile this warp takes
    uint8_t* stg = staging + ew * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;

    if constexpr (GEGLU) {
      // ---- GEGLU (attention.py:32-39): out[:, j] = (acc[:, j] + bv[j]) * gelu(acc[:, BN/2 + j] + bg[j]), bf16.
      // The math runs in the TMEM layout (one row per lane, 32 consecutive columns); only the bf16 result is
      // transposed (2 KB per chunk) so that every warp store instruction writes 8 rows x 64 contiguous bytes.
      const int orow = lane >> 2;      // row within a group of 8 rows in the coalesced phase
      const int ochk = lane & 3;       // 16-byte chunk within the 64-byte bf16 row
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const long long row0 = static_cast<long long>(m_tile) * 128 + q * 32;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c = half * 32; c < BN / 2; c += 64) {
          uint32_t v[32], g[32];
          tmem_ld32(t_acc + c, v);
          tmem_ld32(t_acc + BN / 2 + c, g);
          const int pc = n_tile * BN + c;              // packed column of the value half (bias index)
          const int col0 = n_tile * (BN / 2) + c;      // first output column of this chunk
          const bool col_ok = col0 < p.N / 2;          // N/2 is a multiple of 128: chunks are all-or-nothing
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
            if (p.bias && col_ok) {
              bv = __ldg(reinterpret_cast<const float4*>(p.bias + pc) + j);
              bg = __ldg(reinterpret_cast<const float4*>(p.bias + pc + BN / 2) + j);
            }
            const float o0 = (__uint_as_float(v[4 * j + 0]) + bv.x) * gelu_tanh(__uint_as_float(g[4 * j + 0]) + bg.x);
            const float o1 = (__uint_as_float(v[4 * j + 1]) + bv.y) * gelu_tanh(__uint_as_float(g[4 * j + 1]) + bg.y);
            const float o2 = (__uint_as_float(v[4 * j + 2]) + bv.z) * gelu_tanh(__uint_as_float(g[4 * j + 2]) + bg.z);
            const float o3 = (__uint_as_float(v[4 * j + 3]) + bv.w) * gelu_tanh(__uint_as_float(g[4 * j + 3]) + bg.w);
            pk[2 * j] = pack_bf16x2(o0, o1);
            pk[2 * j + 1] = pack_bf16x2(o2, o3);
          }
          // slab: [32 rows][64 B], 16-byte chunks XOR-swizzled with (row >> 1) & 3 (conflict-free both ways)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int R = 8 * k + orow;
            const uint4 o = *reinterpret_cast<const uint4*>(stg + R * 64 + ((ochk ^ ((R >> 1) & 3)) << 4));
            const long long gr = row0 + R;
            if (gr < p.M && col_ok)
              *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + gr * p.ldo + col0 + ochk * 8) = o;
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    } else {
      // ---- generic: out = acc + bias[col] + rowbias[row / rows_per_group, col] + residual[row, col]  (fp32 | bf16)
      // Each 32x32 fp32 chunk is transposed through a swizzled 4 KB slab: one warp instruction then covers
      // 4 rows x 128 B (fp32).  The residual / rowbias operands of chunk c+1 are requested before chunk c is
      // processed (and those of a tile's first chunk before its accumulator is complete), so the loads overlap
      // the MMA wait, the TMEM read and the stores.
      const int sub = lane >> 3;       // row within a group of 4 rows in the coalesced phase
      const int cl = lane & 7;         // 16-byte column chunk within the 128-byte row
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        // rows this lane stores in the coalesced phase: tile row q*32 + 4k + sub, k = 0..7 (element offsets fit
        // 32 bits: checked on the host)
        uint32_t grow[8];              // output row index
        uint32_t aoff[8];              // element offset of the row's auxiliary operand (residual, else rowbias)
        uint32_t valid = 0;
        long long stat_slot = -1;      // flat GroupNorm-statistics slot of this warp's 32 rows
        if (p.amode == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            grow[k] = m_tile * 128 + q * 32 + 4 * k + sub;
            if (grow[k] < static_cast<uint32_t>(p.M)) valid |= 1u << k;
          }
          stat_slot = static_cast<long long>(m_tile) * 4 + q;
        } else {
          const int tw = m_tile % p.tiles_w;
          const int th = (m_tile / p.tiles_w) % p.tiles_h;
          const int tn = m_tile / (p.tiles_w * p.tiles_h);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = q * 32 + 4 * k + sub;
            const int rw = r % p.bw;
            const int rh = (r / p.bw) % p.bh;
            const int rn = r / (p.bw * p.bh);
            const int n = tn * p.nb + rn, h = th * p.bh + rh, w = tw * p.bw + rw;
            if (n < p.B && h < p.H && w < p.W) valid |= 1u << k;
            grow[k] = (n * p.H + h) * p.W + w;
          }
          const int per = p.bw * p.bh;                       // pixels of one sample inside the tile (>= 32 if stats)
          const int n0s = tn * p.nb + (q * 32) / per;
          if (n0s < p.B)
            stat_slot = static_cast<long long>(n0s) * p.gn_slots + (th * p.tiles_w + tw) * (per >> 5) + (((q * 32) % per) >> 5);
        }
        // ONE auxiliary fp32 operand is prefetched per row: the residual if present, else the per-sample rowbias.
        // (Both together only occur in tests; the rowbias is then added at consumption time.)
        const float* aux = p.residual ? p.residual : p.rowbias;
        const bool late_rowbias = p.residual != nullptr && p.rowbias != nullptr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (p.residual) aoff[k] = grow[k] * static_cast<uint32_t>(p.ldr);
          else if (p.rowbias) aoff[k] = (grow[k] / static_cast<uint32_t>(p.rows_per_group)) * static_cast<uint32_t>(p.ld_rowbias);
          else aoff[k] = 0;
        }
        const int ncols = p.N;
        auto load_aux = [&](int c, float4 (&rs)[8]) {
          const int col = n_tile * BN + c + 4 * cl;
          const bool col_ok = col < ncols;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const bool ok = ((valid >> k) & 1u) && col_ok;
            const float4* ptr = reinterpret_cast<const float4*>(aux + aoff[k] + (ok ? col : 0));
            rs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) rs[k] = *ptr;
          }
        };
        const bool has_aux = aux != nullptr;
        float4 rs[8];
        if (has_aux) load_aux(half * 32, rs);
        else {
#pragma unroll
          for (int k = 0; k < 8; ++k) rs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }

        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_acc = tmem_base + acc * Cfg::kAccStride + (static_cast<uint32_t>(q * 32) << 16);

        constexpr int kChunkIters = (BN + 63) / 64;
#pragma unroll
        for (int it = 0; it < kChunkIters; ++it) {
          const int c = half * 32 + it * 64;
          if (c >= BN) break;
          const int col = n_tile * BN + c + 4 * cl;  // this lane's 4 output columns
          const bool col_ok = col < ncols;
          uint32_t v[32];
          tmem_ld32(t_acc + c, v);
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          float4 rsn[8];
          const bool more = c + 64 < BN;
          if (has_aux && more) load_aux(c + 64, rsn);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = s4;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int R = 4 * k + sub;
            const float4 a = *reinterpret_cast<const float4*>(stg + R * 128 + ((cl ^ (R & 7)) << 4));
            float4 t4 = make_float4(a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w);
            if (p.act == 1) {
              t4.x = quick_gelu(t4.x); t4.y = quick_gelu(t4.y); t4.z = quick_gelu(t4.z); t4.w = quick_gelu(t4.w);
            }
            float4 o = make_float4(t4.x + rs[k].x, t4.y + rs[k].y, t4.z + rs[k].z, t4.w + rs[k].w);
            if (((valid >> k) & 1u) && col_ok) {
              if (late_rowbias) {
                const float4 rb = __ldg(reinterpret_cast<const float4*>(
                    p.rowbias + static_cast<long long>(grow[k] / static_cast<uint32_t>(p.rows_per_group)) * p.ld_rowbias + col));
                o.x += rb.x; o.y += rb.y; o.z += rb.z; o.w += rb.w;
              }
              if (p.out_bf16) {
                uint2 pk;
                pk.x = pack_bf16x2(o.x, o.y);
                pk.y = pack_bf16x2(o.z, o.w);
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + grow[k] * static_cast<uint32_t>(p.ldo) + col) = pk;
              } else {
                *reinterpret_cast<float4*>(static_cast<float*>(p.out) + grow[k] * static_cast<uint32_t>(p.ldo) + col) = o;
              }
              s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
              q4.x = fmaf(o.x, o.x, q4.x); q4.y = fmaf(o.y, o.y, q4.y);
              q4.z = fmaf(o.z, o.z, q4.z); q4.w = fmaf(o.w, o.w, q4.w);
            }
          }
          if (p.gn_stats) {
            // fixed-order reduction over the 4 row sub-groups (deterministic), lanes 0..7 publish 4 channels each
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              s4.x += __shfl_xor_sync(0xffffffffu, s4.x, o); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, o);
              s4.z += __shfl_xor_sync(0xffffffffu, s4.z, o); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, o);
              q4.x += __shfl_xor_sync(0xffffffffu, q4.x, o); q4.y += __shfl_xor_sync(0xffffffffu, q4.y, o);
              q4.z += __shfl_xor_sync(0xffffffffu, q4.z, o); q4.w += __shfl_xor_sync(0xffffffffu, q4.w, o);
            }
            if (sub == 0 && col_ok && stat_slot >= 0) {
              float4* dst = reinterpret_cast<float4*>(p.gn_stats + (stat_slot * ncols + col) * 2);
              dst[0] = make_float4(s4.x, q4.x, s4.y, q4.y);
              dst[1] = make_float4(s4.z, q4.z, s4.w, q4.w);
            }
          }
          if (has_aux && more) {
#pragma unroll
            for (int k = 0; k < 8; ++k) rs[k] = rsn[k];   // fully unrolled chunk loop: pure register renaming
          }
          __syncwarp();
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
template <int BN, bool GEGLU>
static int launch_gemm(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc_kernel<BN, GEGLU><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(p);
  AF_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

static int dispatch_gemm(int bn, const GemmParams& p, cudaStream_t stream) {
  if (p.geglu) return launch_gemm<256, true>(p, stream);
  switch (bn) {
    case 64: return launch_gemm<64, false>(p, stream);
    case 128: return launch_gemm<128, false>(p, stream);
    case 160: return launch_gemm<160, false>(p, stream);
    case 256: return launch_gemm<256, false>(p, stream);
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
  p.gn_stats = ep->gn_stats;
  p.act = ep->act;
  AF_CHECK_ARG(ep->act == 0 || (ep->act == 1 && !ep->geglu), "epilogue: act %d unsupported", ep->act);
  AF_CHECK_ARG(!ep->gn_stats || (!ep->geglu && (reinterpret_cast<uintptr_t>(ep->gn_stats) & 15) == 0),
               "epilogue: gn_stats needs the generic epilogue and a 16-byte aligned buffer");
  AF_CHECK_ARG(ep->out != nullptr, "epilogue: out is null");
  AF_CHECK_ARG(ep->out_dtype == AF_DTYPE_BF16 || ep->out_dtype == AF_DTYPE_F32, "epilogue: bad out dtype %d",
               ep->out_dtype);
  AF_CHECK_ARG(!ep->geglu || ep->out_dtype == AF_DTYPE_BF16, "geglu epilogue writes bf16 only");
  AF_CHECK_ARG(!ep->geglu || (!ep->residual && !ep->rowbias), "geglu epilogue: residual / rowbias unsupported");
  AF_CHECK_ARG((reinterpret_cast<uintptr_t>(ep->out) & 15) == 0, "epilogue: out not 16B aligned");
  {
    const long long lim = 1ll << 31, rows = p.M;
    const long long groups = ep->rows_per_group > 0 ? rows / p.rows_per_group + 1 : 4096;  // conv: one group per image
    AF_CHECK_ARG(rows * p.ldo < lim && rows * p.ldr < lim && (!p.rowbias || groups * p.ld_rowbias < lim),
                 "epilogue: output / residual / rowbias larger than 2^31 elements");
  }
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

// output-pixel box of one 128-row conv tile: bw = largest power of two <= min(64, next_pow2(Wo)); rows / images
// fill up to 128 pixels
static void conv_tile_box(int Ho, int Wo, int* bw_, int* bh_, int* nb_) {
  int bw = 1;
  while (bw < Wo && bw < 64) bw <<= 1;
  int bh = 1;
  while (bh < Ho && bw * bh < 128) bh <<= 1;
  *bw_ = bw; *bh_ = bh; *nb_ = 128 / (bw * bh);
}

extern "C" int af_conv3x3_gn_slots(int Ho, int Wo) {
  int bw, bh, nb;
  conv_tile_box(Ho, Wo, &bw, &bh, &nb);
  if (bw * bh < 32) return 0;  // a 32-row quarter would straddle samples: use af_groupnorm_stats instead
  return ((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * (bw * bh / 32);
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
  int bw, bh, nb;
  conv_tile_box(Ho, Wo, &bw, &bh, &nb);
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
  if (p.gn_stats) {
    p.gn_slots = af_conv3x3_gn_slots(Ho, Wo);
    AF_CHECK_ARG(p.gn_slots > 0, "af_conv3x3_bf16: gn_stats unsupported for %dx%d outputs (fewer than 32 pixels per tile row group)", Ho, Wo);
  }
  return dispatch_gemm(bn, p, stream);
}
