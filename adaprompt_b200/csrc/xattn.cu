// Cross-attention over a short (<= 128 keys) context: the 77-token CLIP + AdaFace prompt of every
// SpatialTransformer (CrossAttention.forward, ldm/modules/attention.py:172-243 with `context` given).
//
// The scores of one query tile fit a single key block, so there is no online-softmax loop to amortise the per-CTA
// set-up (barrier init, TMEM allocation, K / V fetch) over - with one CTA per 256 queries that set-up dominated
// (profiles/r01_unet_step_by_shape_v3.md: 148 us for 84 MB of traffic at N = 4096).  Here a CTA owns one (sample, head)
// and a contiguous run of query tiles; K and V^T stay resident in shared memory, Q tiles stream through a TMA ring,
// and two softmax warpgroups alternate tiles so that the MMA / TMEM round trips of one tile hide behind the
// exponentials and the output stores of the other:
//
//   warp 0      TMA producer: K, V^T once; Q_i ring
//   warp 1      MMA issuer:   S_0 S_1 | PV_0 S_2 | PV_1 S_3 | ...
//   warps 2-5   softmax + epilogue of even tiles      warps 6-9   odd tiles   (TMEM lane quarter = warp % 4)
//
// Operand conventions are those of af_attention_bf16 (attention.cu): Q pre-scaled by d^-1/2 * log2(e), K as
// [B][kv_stride][ldk], V transposed ([heads*d][ldvt], sample b's keys at column b*kv_stride), bf16 output.
#include <math.h>

#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

struct XAttnParams {
  float* lse;        // optional [B][heads][Nq] log2-sum-exp per query row
  CUtensorMap tmQ;   // 3-D {heads*dp, Nq, B}, box {64, 128, 1}
  CUtensorMap tmK;   // 3-D {heads*dp, Nk, B}, box {64, 128, 1}
  CUtensorMap tmV;   // 2-D {ldvt, heads*d}, box {64, DV}
  int B, heads, Nq, Nk;
  int d, dp, kv_stride;
  int tiles_per_cta;
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  long long ldo;
};

template <int D>
struct XCfg;
template <>
struct XCfg<40> {
  static constexpr int DK = 48, DV = 48, QSTAGES = 3;
};
template <>
struct XCfg<80> {
  static constexpr int DK = 80, DV = 80, QSTAGES = 3;
};

template <int D>
struct XSmem {
  using C = XCfg<D>;
  static constexpr int KA = (C::DK + 63) / 64;
  static constexpr int kQBytes = KA * 128 * 128;
  static constexpr int kKBytes = KA * 128 * 128;
  static constexpr int kVAtomBytes = C::DV * 128;
  static constexpr int kVBytes = ((2 * kVAtomBytes + 1023) / 1024) * 1024;
  static constexpr int kPBytes = 2 * 128 * 128;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + C::QSTAGES * kQBytes;
  static constexpr int kVOff = kKOff + kKBytes;
  static constexpr int kPOff = kVOff + kVBytes;
  static constexpr int kBarOff = kPOff + 2 * kPBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

__device__ __forceinline__ void xa_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int D, int NCH>
__global__ void __launch_bounds__(320, 1) xattn_kernel(const __grid_constant__ XAttnParams p) {
  using C = XCfg<D>;
  using S = XSmem<D>;
  constexpr int DK = C::DK, DV = C::DV, QST = C::QSTAGES, KA = S::KA;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+DV) O1 [384,384+DV)
  constexpr uint32_t kTmemS[2] = {0, 128};
  constexpr uint32_t kTmemO[2] = {256, 384};

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* kv_full = bars;                // 1
  uint64_t* q_full = kv_full + 1;          // QST
  uint64_t* q_empty = q_full + QST;        // QST
  uint64_t* s_full = q_empty + QST;        // 2: S_i in TMEM
  uint64_t* p_full = s_full + 2;           // 2: P_i in shared memory (implies S_i was read)
  uint64_t* pv_done = p_full + 2;          // 2: O_i complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles_all = (p.Nq + 127) / 128;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  const int nt = min(p.tiles_per_cta, n_tiles_all - tile0);
  constexpr int kcols16 = NCH * 16;           // keys that take part in the MMAs (Nk rounded up to 16)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(kv_full, 1);
    for (int s = 0; s < QST; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the tcgen05 issuer

  // 10 warps (no idle warps, no setmaxnreg): 65536 / 320 leaves 200 registers for the 80..128-wide score rows
  if (warp < 2) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      {   // whole warp, convergent; TMA instructions by one elected lane
        if (elect_one()) {
          mbar_arrive_expect_tx(kv_full, S::kKBytes + 2 * S::kVAtomBytes);
#pragma unroll
          for (int a = 0; a < KA; ++a)
            tma_load_3d(smem + S::kKOff + a * 128 * 128, &p.tmK, kv_full, h * p.dp + a * 64, 0, b);
#pragma unroll
          for (int a = 0; a < 2; ++a)
            tma_load_2d(smem + S::kVOff + a * S::kVAtomBytes, &p.tmV, kv_full, b * p.kv_stride + a * 64, h * p.d);
        }
        __syncwarp();
        int slot = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nt; ++i) {
          mbar_wait(&q_empty[slot], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&q_full[slot], S::kQBytes);
#pragma unroll
            for (int a = 0; a < KA; ++a)
              tma_load_3d(smem + S::kQOff + slot * S::kQBytes + a * 128 * 128, &p.tmQ, &q_full[slot],
                          h * p.dp + a * 64, (tile0 + i) * 128, b);
          }
          __syncwarp();
          if (++slot == QST) { slot = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (whole warp, convergent; the
      // tcgen05 instructions are issued by one elected lane: no per-instruction elect / branch waterfall loops)
      {
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, kcols16);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV);
        const uint32_t k_addr = smem_u32(smem + S::kKOff);
        const uint32_t v_addr = smem_u32(smem + S::kVOff);
        const uint32_t p_addr = smem_u32(smem + S::kPOff);
        constexpr int ksteps = NCH;
        auto issue_s = [&](int i) {
          const int slot = i % QST;
          mbar_wait(&q_full[slot], (i / QST) & 1);
          tc_fence_after();
          const uint32_t q_addr = smem_u32(smem + S::kQOff + slot * S::kQBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DK / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(q_addr + (k >> 2) * 128 * 128) + 2 * (k & 3);
              const uint64_t bd = umma_desc_sw128(k_addr + (k >> 2) * 128 * 128) + 2 * (k & 3);
              tc_mma_ss(tmem_base + kTmemS[i & 1], ad, bd, idesc_s, k != 0 ? 1u : 0u);
            }
            tc_commit(&s_full[i & 1]);
            tc_commit(&q_empty[slot]);
          }
          __syncwarp();
        };
        mbar_wait(kv_full, 0);
        issue_s(0);
        if (nt > 1) issue_s(1);
        for (int i = 0; i < nt; ++i) {
          const int w = i & 1;
          mbar_wait(&p_full[w], (i >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = umma_desc_sw128(p_addr + w * S::kPBytes + (k >> 2) * 128 * 128) + 2 * (k & 3);
              const uint64_t bd = umma_desc_sw128(v_addr + (k >> 2) * S::kVAtomBytes) + 2 * (k & 3);
              tc_mma_ss(tmem_base + kTmemO[w], ad, bd, idesc_o, k != 0 ? 1u : 0u);
            }
            tc_commit(&pv_done[w]);
          }
          __syncwarp();
          if (i + 2 < nt) issue_s(i + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warpgroups
    const int w = (warp - 2) >> 2;
    const int qd = warp & 3;                      // TMEM lane quarter (hardware: warp id % 4)
    const int r = qd * 32 + lane;                 // row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint8_t* mrow = p.key_mask ? p.key_mask + static_cast<size_t>(b) * p.Nk : nullptr;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint8_t* p_row = smem + S::kPOff + w * S::kPBytes + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t s_addr = tmem_base + kTmemS[w] + lane_off;
    const uint32_t o_addr = tmem_base + kTmemO[w] + lane_off;
    const int Nk = p.Nk;
    // valid-key bitmap (key < Nk and not masked out), built once: the mask depends on the sample only
    uint32_t kbits[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int key = j * 32 + lane;
      bool ok = key < Nk;
      if (mrow != nullptr && ok) ok = __ldg(mrow + key) != 0;
      kbits[j] = __ballot_sync(0xffffffffu, ok);
    }

    for (int i = w; i < nt; i += 2) {
      const uint32_t par = (i >> 1) & 1;
      mbar_wait(&s_full[w], par);
      tc_fence_after();
      float sc[kcols16];
#pragma unroll
      for (int c = 0; c < kcols16; c += 16) xa_tmem_ld16(s_addr + c, reinterpret_cast<uint32_t*>(sc) + c);
      tmem_ld_wait();
      // keys >= Nk (zero-filled K rows) and masked keys do not take part (warp-uniform test per 16-key chunk)
#pragma unroll
      for (int c = 0; c < kcols16; c += 16)
        if (mrow != nullptr || c + 16 > Nk) {
#pragma unroll
          for (int e = c; e < c + 16; ++e) sc[e] = ((kbits[e >> 5] >> (e & 31)) & 1u) ? sc[e] : -INFINITY;
        }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < kcols16; ++e) mx4[e & 3] = fmaxf(mx4[e & 3], sc[e]);
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_use = (mx == -INFINITY) ? 0.f : mx;
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c8 = 0; c8 < kcols16; c8 += 8) {
            float pe[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              pe[e] = fast_exp2(sc[c8 + e] - m_use);
              l4[e & 3] += pe[e];
            }
            uint4 pk;
            pk.x = pack_bf16x2(pe[0], pe[1]);
            pk.y = pack_bf16x2(pe[2], pe[3]);
            pk.z = pack_bf16x2(pe[4], pe[5]);
            pk.w = pack_bf16x2(pe[6], pe[7]);
            uint8_t* atom = p_row + (c8 >> 6) * (128 * 128);
            const uint32_t chunk = static_cast<uint32_t>((c8 & 63) >> 3);
            sts128(smem_u32(atom) + ((chunk ^ sw) << 4), pk.x, pk.y, pk.z, pk.w);   // STS, not a generic ST.E
      }
      const float l_run = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[w]);

      // epilogue of this tile: O / l -> bf16
      const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
      const int q_row = (tile0 + i) * 128 + r;
      __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.Nq + q_row) * p.ldo + h * p.d;
      if (p.lse != nullptr && q_row < p.Nq)
        p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Nq + q_row] = l_run > 0.f ? m_use + __log2f(l_run) : INFINITY;
      mbar_wait(&pv_done[w], par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t o[16];
        tmem_ld16(o_addr + c, o);
        tmem_ld_wait();
        if (q_row < p.Nq) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (c + g * 8 < D) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
              pk.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
              pk.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
              pk.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c + g * 8) = pk;
            }
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// tiles per CTA: minimise waves * (tiles + set-up) over the (sample, head) x chunk grid
static int pick_tiles_per_cta(int pairs, int n_tiles) {
  const int sms = num_sms();
  int best = n_tiles;
  double best_cost = 1e30;
  for (int tpc = 2; tpc <= n_tiles; ++tpc) {
    const int chunks = (n_tiles + tpc - 1) / tpc;
    const long long ctas = static_cast<long long>(pairs) * chunks;
    const double waves = static_cast<double>((ctas + sms - 1) / sms);
    const double cost = waves * (tpc + 3.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = tpc;
    }
  }
  return n_tiles < 2 ? 1 : best;
}

template <int D, int NCH>
static int launch_xattn_n(const XAttnParams& p, dim3 grid, cudaStream_t stream) {
  using S = XSmem<D>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(xattn_kernel<D, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  xattn_kernel<D, NCH><<<grid, 320, S::kTotal, stream>>>(p);
  AF_LAUNCH_CHECK("xattn_kernel");
  return 0;
}

template <int D>
static int launch_xattn(XAttnParams& p, cudaStream_t stream) {
  const int n_tiles = (p.Nq + 127) / 128;
  p.tiles_per_cta = pick_tiles_per_cta(p.B * p.heads, n_tiles);
  dim3 grid((n_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta, p.heads, p.B);
  switch ((p.Nk + 15) / 16) {
    case 1: return launch_xattn_n<D, 1>(p, grid, stream);
    case 2: return launch_xattn_n<D, 2>(p, grid, stream);
    case 3: return launch_xattn_n<D, 3>(p, grid, stream);
    case 4: return launch_xattn_n<D, 4>(p, grid, stream);
    case 5: return launch_xattn_n<D, 5>(p, grid, stream);
    case 6: return launch_xattn_n<D, 6>(p, grid, stream);
    case 7: return launch_xattn_n<D, 7>(p, grid, stream);
    default: return launch_xattn_n<D, 8>(p, grid, stream);
  }
}

// Called by af_attention_bf16 (attention.cu) for d in {40, 80}, Nk <= 128.  Returns -100 if unsupported.
int xattn_dispatch(const void* Q, long long ldq, const void* K, long long ldk, const void* Vt, long long ldvt,
                   int kv_stride, const unsigned char* key_mask, void* O, int B, int heads, int Nq, int Nk, int d,
                   float* lse, cudaStream_t stream) {
  if (!(d == 40 || d == 80) || Nk > 128) return -100;
  XAttnParams p;
  memset(&p, 0, sizeof(p));
  const int dp = d == 40 ? 48 : d;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(heads) * dp, static_cast<uint64_t>(Nq), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(ldq) * 2, static_cast<uint64_t>(Nq) * ldq * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_bf16(&p.tmQ, Q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(heads) * dp, static_cast<uint64_t>(Nk), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(ldk) * 2, static_cast<uint64_t>(kv_stride) * ldk * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_bf16(&p.tmK, K, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(ldvt), static_cast<uint64_t>(heads) * d};
    uint64_t str[1] = {static_cast<uint64_t>(ldvt) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(dp)};
    int rc = make_tmap_bf16(&p.tmV, Vt, 2, dims, str, box);
    if (rc) return rc;
  }
  p.B = B; p.heads = heads; p.Nq = Nq; p.Nk = Nk; p.d = d; p.dp = dp; p.kv_stride = kv_stride;
  p.key_mask = key_mask;
  p.lse = lse;
  p.out = static_cast<__nv_bfloat16*>(O);
  p.ldo = static_cast<long long>(heads) * d;
  return d == 40 ? launch_xattn<40>(p, stream) : launch_xattn<80>(p, stream);
}

}  // namespace af
