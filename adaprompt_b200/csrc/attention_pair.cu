// Two-query-tile flash attention (256 queries per CTA) for the large-N self-attention layers.
//
// Same operand conventions and reference semantics as attention.cu (CrossAttention.forward,
// ldm/modules/attention.py:172-243); the difference is the schedule.  Softmax on B200 is limited by the
// MUFU (ex2) unit, not by the tensor core, so the CTA keeps TWO 128-row query tiles in flight:
//
//   warp 0      TMA producer (Q0, Q1 once; K_j / V_j ring shared by both tiles)
//   warp 1      MMA issuer:  ... PV0_j, S0_{j+1}, PV1_j, S1_{j+1}, PV0_{j+1} ...   (warps 2-3 idle)
//   warps 4-7   softmax warpgroup of tile 0      warps 8-11  softmax warpgroup of tile 1
//
// While one warpgroup exponentiates its S tile the tensor core works for the other one, every SM sub-partition
// holds two softmax warps (latency hiding), and K / V are fetched once per 256 queries.
// A softmax warpgroup copies its S tile to registers first and immediately hands the TMEM buffer back (s_free), so
// the MMA warp computes S_{j+1} of the tile while the warpgroup is still exponentiating block j: the next scores
// are already waiting when P_j has been written (profiles/r01_attention_pair_before.md: 43 % of the softmax warps'
// time used to be spent waiting for S).  P / O reuse is ordered by pv_done (committed after PV_j).
// The running max is only refreshed when it grows by more than 2^8 (lazy rescale): P stays <= 256 in bf16
// and the O / l correction pass almost never runs after the first block.
#include <math.h>

#include "../../include/adaface_b200.h"
#include "common.cuh"

#ifndef AF_ATTN_STAGGER_CYCLES
#define AF_ATTN_STAGGER_CYCLES 1100
#endif

namespace af {

struct AttnPairParams {
  float* lse;        // optional [B][heads][Nq] log2-sum-exp per query row
  CUtensorMap tmQ;   // 3-D {heads*dp, Nq, B}, box {64, 128, 1}
  CUtensorMap tmK;   // 3-D {heads*dp, Nk, B}, box {64, BLOCK_N, 1}
  CUtensorMap tmV;   // 2-D {ldvt, heads*d}, box {64, DV}
  int B, heads, Nq, Nk;
  int d, dp, kv_stride;
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  long long ldo;
};

template <int D>
struct PairCfg;
template <>
struct PairCfg<40> {
  static constexpr int DK = 48, DV = 48, BLOCK_N = 128, KSTAGES = 3, VSTAGES = 3;
};
template <>
struct PairCfg<80> {
  static constexpr int DK = 80, DV = 80, BLOCK_N = 64, KSTAGES = 3, VSTAGES = 3;
};

template <int D>
struct PairSmem {
  using C = PairCfg<D>;
  static constexpr int KA = (C::DK + 63) / 64;
  static constexpr int PA = C::BLOCK_N / 64;
  static constexpr int kQBytes = KA * 128 * 128;          // one query tile
  static constexpr int kKBytes = KA * C::BLOCK_N * 128;
  static constexpr int kVAtomBytes = C::DV * 128;
  static constexpr int kVBytes = PA * kVAtomBytes;
  static constexpr int kPBytes = PA * 128 * 128;          // one P tile
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + 2 * kQBytes;
  static constexpr int kVOff = kKOff + C::KSTAGES * kKBytes;
  static constexpr int kPOff = kVOff + ((C::VSTAGES * kVBytes + 1023) / 1024) * 1024;
  static constexpr int kBarOff = kPOff + 2 * kPBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

__device__ __forceinline__ void tmem_ld32q(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16q(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int D>
__global__ void __launch_bounds__(384, 1) attention_pair_kernel(const __grid_constant__ AttnPairParams p) {
  using C = PairCfg<D>;
  using S = PairSmem<D>;
  constexpr int DK = C::DK, DV = C::DV, BN = C::BLOCK_N;
  constexpr int KA = S::KA, PA = S::PA;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+DV) O1 [384,384+DV)
  constexpr uint32_t kTmemS[2] = {0, 128};
  constexpr uint32_t kTmemO[2] = {256, 384};
  static_assert(DV <= 128, "two O accumulators must fit next to two S tiles in 512 TMEM columns");
  constexpr float kRescaleThreshold = 8.0f;  // log2 domain

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = q_full + 1;               // KSTAGES
  uint64_t* k_empty = k_full + C::KSTAGES;
  uint64_t* v_full = k_empty + C::KSTAGES;     // VSTAGES
  uint64_t* v_empty = v_full + C::VSTAGES;
  uint64_t* s_full = v_empty + C::VSTAGES;     // 2 (per tile): S_j in TMEM
  uint64_t* p_full = s_full + 2;               // 2 (per tile): P_j in shared memory
  uint64_t* pv_done = p_full + 2;              // 2 (per tile): PV_j complete (P buffer / O accumulator reusable)
  uint64_t* s_free = pv_done + 2;              // 2 (per tile): S_j copied to registers (TMEM buffer reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_blocks = (p.Nk + BN - 1) / BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::KSTAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < C::VSTAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
      mbar_init(&s_free[t], 4);
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
  const uint32_t tmem_base = *tmem_slot;

  // register re-balancing (per warpgroup): the data-movement warpgroup keeps 40 registers per thread, the two
  // softmax warpgroups get 232 (a 128-wide fp32 score row lives in registers)
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * S::kQBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int a = 0; a < KA; ++a)
          tma_load_3d(smem + S::kQOff + t * S::kQBytes + a * 128 * 128, &p.tmQ, q_full, h * p.dp + a * 64,
                      q0 + t * 128, b);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], S::kKBytes);
#pragma unroll
        for (int a = 0; a < KA; ++a)
          tma_load_3d(smem + S::kKOff + ks * S::kKBytes + a * BN * 128, &p.tmK, &k_full[ks], h * p.dp + a * 64,
                      j * BN, b);
        if (++ks == C::KSTAGES) { ks = 0; kph ^= 1; }
        mbar_wait(&v_empty[vs], vph ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], S::kVBytes);
#pragma unroll
        for (int a = 0; a < PA; ++a)
          tma_load_2d(smem + S::kVOff + vs * S::kVBytes + a * S::kVAtomBytes, &p.tmV, &v_full[vs],
                      b * p.kv_stride + j * BN + a * 64, h * p.d);
        if (++vs == C::VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV);
      const uint32_t q_addr = smem_u32(smem + S::kQOff);
      const uint32_t p_addr = smem_u32(smem + S::kPOff);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0, pph = 0;
      // S_t = Q_t . K^T for the K block currently at ring slot `kslot`
      auto issue_s = [&](int t, int kslot) {
        const uint32_t k_addr = smem_u32(smem + S::kKOff + kslot * S::kKBytes);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          const uint64_t ad = umma_desc_sw128(q_addr + t * S::kQBytes + (k >> 2) * 128 * 128) + 2 * (k & 3);
          const uint64_t bd = umma_desc_sw128(k_addr + (k >> 2) * BN * 128) + 2 * (k & 3);
          tc_mma_ss(tmem_base + kTmemS[t], ad, bd, idesc_s, k != 0 ? 1u : 0u);
        }
        tc_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      tc_commit(&k_empty[0]);
      // Event-driven issue: the two query tiles advance independently (an in-order schedule couples them - a
      // warpgroup would wait for the other tile's softmax before its own PV is issued).  Per tile: S_b needs
      // s_free (S_{b-1} copied to registers) and K_b; PV_b needs p_full (P_b written) and V_b.  A K / V ring slot
      // is released when both tiles have consumed it.
      int s_next[2] = {1, 1}, pv_next[2] = {0, 0};
      uint64_t t_start;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
      uint32_t spins = 0;
      while (pv_next[0] < n_blocks || pv_next[1] < n_blocks) {
        bool progress = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int sb = s_next[t];
          if (sb < n_blocks) {
            const int slot = sb % C::KSTAGES;
            if (mbar_test(&s_free[t], (sb - 1) & 1) && mbar_test(&k_full[slot], (sb / C::KSTAGES) & 1)) {
              tc_fence_after();
              issue_s(t, slot);
              if (s_next[1 - t] > sb) tc_commit(&k_empty[slot]);   // the other tile already used K_b
              s_next[t] = sb + 1;
              progress = true;
            }
          }
          const int pb = pv_next[t];
          if (pb < n_blocks) {
            const int slot = pb % C::VSTAGES;
            if (mbar_test(&p_full[t], pb & 1) && mbar_test(&v_full[slot], (pb / C::VSTAGES) & 1)) {
              tc_fence_after();
              const uint32_t v_addr = smem_u32(smem + S::kVOff + slot * S::kVBytes);
#pragma unroll
              for (int k = 0; k < BN / 16; ++k) {
                const uint64_t ad = umma_desc_sw128(p_addr + t * S::kPBytes + (k >> 2) * 128 * 128) + 2 * (k & 3);
                const uint64_t bd = umma_desc_sw128(v_addr + (k >> 2) * S::kVAtomBytes) + 2 * (k & 3);
                tc_mma_ss(tmem_base + kTmemO[t], ad, bd, idesc_o, (pb | k) != 0 ? 1u : 0u);
              }
              tc_commit(&pv_done[t]);
              if (pv_next[1 - t] > pb) tc_commit(&v_empty[slot]);  // the other tile already used V_b
              pv_next[t] = pb + 1;
              progress = true;
            }
          }
        }
        if (!progress && (++spins & 0xfffff) == 0) {                // watchdog (same policy as mbar_wait)
          uint64_t t_now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
          if (t_now - t_start > AF_WATCHDOG_NS) {
            printf("af watchdog: attention MMA issue loop stuck (block %d,%d,%d s %d/%d pv %d/%d)\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, s_next[0], s_next[1], pv_next[0], pv_next[1]);
            __trap();
          }
        }
      }
      (void)ks; (void)vs; (void)kph; (void)vph; (void)pph;
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------------ softmax warpgroups (tile t = 0 / 1)
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;                      // TMEM lane quarter (hardware: warp id % 4)
    const int r = qd * 32 + lane;                 // row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int q_row = q0 + t * 128 + r;
    const uint8_t* mrow = p.key_mask ? p.key_mask + static_cast<size_t>(b) * p.Nk : nullptr;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint8_t* p_row = smem + S::kPOff + t * S::kPBytes + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t s_addr = tmem_base + kTmemS[t] + lane_off;
    const uint32_t o_addr = tmem_base + kTmemO[t] + lane_off;

    // The two warpgroups share the SFU (ex2) units of their SM sub-partitions.  Started together they run in phase
    // - both exponentiating at half rate, then both idle in TMEM loads / maxima - so tile 1 is started half a
    // block late: its load / max phases then fall into tile 0's exponentiation phase and vice versa.
    if (t == 1 && n_blocks > 2) {
      const long long t0 = clock64();
      while (clock64() - t0 < AF_ATTN_STAGGER_CYCLES) {
      }
    }
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const int key0 = j * BN;
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float sc[BN];
#pragma unroll
      for (int c = 0; c < BN; c += 32) tmem_ld32q(s_addr + c, reinterpret_cast<uint32_t*>(sc) + c);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);     // the MMA warp may overwrite S with the next block's scores
      if (key0 + BN > p.Nk || mrow != nullptr) {  // warp-uniform: tail block / explicit key mask only
#pragma unroll
        for (int e = 0; e < BN; ++e) {
          const int key = key0 + e;
          bool ok = key < p.Nk;
          if (mrow != nullptr) ok = ok && (__ldg(mrow + min(key, p.Nk - 1)) != 0);
          sc[e] = ok ? sc[e] : -INFINITY;
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int e = 0; e < BN; e += 4) {
        mx4[0] = fmaxf(mx4[0], sc[e]);
        mx4[1] = fmaxf(mx4[1], sc[e + 1]);
        mx4[2] = fmaxf(mx4[2], sc[e + 2]);
        mx4[3] = fmaxf(mx4[3], sc[e + 3]);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // lazy rescale: keep the old reference unless the max grew by more than 2^8 (first finite max always taken)
      float alpha = 1.0f;
      if (mx > m_ref + kRescaleThreshold) {
        alpha = fast_exp2(m_ref - mx);  // 0 when m_ref == -inf
        m_ref = mx;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      // PV_{j-1} complete: O and the P buffer are ours now
      if (j > 0) {
        mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
        for (int c = 0; c < DV; c += 16) {
          uint32_t o[16];
          tmem_ld16(o_addr + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st16q(o_addr + c, o);
        }
        tmem_st_wait();
      }
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c8 = 0; c8 < BN; c8 += 8) {
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
        *reinterpret_cast<uint4*>(atom + ((chunk ^ sw) << 4)) = pk;
      }
      l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // epilogue: O / l -> bf16
    mbar_wait(&pv_done[t], (n_blocks - 1) & 1);
    tc_fence_after();
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    if (p.lse != nullptr && q_row < p.Nq)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Nq + q_row] = l_run > 0.f ? m_ref + __log2f(l_run) : INFINITY;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.Nq + q_row) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = 0; c < DV; c += 16) {
      uint32_t o[16];
      tmem_ld16(o_addr + c, o);
      tmem_ld_wait();
      if (q_row < p.Nq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D>
static int launch_pair(const AttnPairParams& p, cudaStream_t stream) {
  using S = PairSmem<D>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(attention_pair_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  dim3 grid((p.Nq + 255) / 256, p.heads, p.B);
  attention_pair_kernel<D><<<grid, 384, S::kTotal, stream>>>(p);
  AF_LAUNCH_CHECK("attention_pair_kernel");
  return 0;
}

// Called by af_attention_bf16 (attention.cu) for d in {40, 80} when Nq >= 256.  Returns -100 if unsupported.
int attention_pair_dispatch(const void* Q, long long ldq, const void* K, long long ldk, const void* Vt, long long ldvt,
                            int kv_stride, const unsigned char* key_mask, void* O, int B, int heads, int Nq, int Nk,
                            int d, float* lse, cudaStream_t stream) {
  if (!(d == 40 || d == 80)) return -100;
  AttnPairParams p;
  memset(&p, 0, sizeof(p));
  const int dp = d == 40 ? 48 : d;
  const int bn = d == 40 ? 128 : 64;
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
    uint32_t box[3] = {64, static_cast<uint32_t>(bn), 1};
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
  return d == 40 ? launch_pair<40>(p, stream) : launch_pair<80>(p, stream);
}

}  // namespace af
