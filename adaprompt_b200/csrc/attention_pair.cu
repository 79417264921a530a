// Two-query-tile flash attention (256 queries per CTA) for the large-N self-attention layers.
//
// Same operand conventions and reference semantics as attention.cu (CrossAttention.forward,
// ldm/modules/attention.py:172-243); the difference is the schedule.  Softmax on B200 is limited by the
// MUFU (ex2) unit, not by the tensor core, so the CTA keeps TWO 128-row query tiles in flight:
//
//   warp 0      TMA producer (Q0, Q1 once; K_j / V_j ring shared by both tiles)
//   warps 1, 2  MMA issuers, one per query tile:  S_{j+1}, PV_j, S_{j+2}, PV_{j+1} ...     (warp 3 idle)
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

// 1: P (bf16 probabilities) is handed to the PV MMA through TENSOR MEMORY (tcgen05.st by the softmax thread that owns
//    the row, A-operand-in-TMEM MMA): no P stores to / operand reads from shared memory.  At d = 40 the shared-memory
//    pipe (S and PV operand reads + P stores + TMA fills = 228 KB per 128-key block) was as loaded as the ex2 unit.
// 0: P through 128B-swizzled shared memory (SS-mode MMA).
// Both schedules (and the per-64-key split of the hand-off) stay selectable at run time for A/B measurements:
// af_attention_set_pair_variant(bit 0 = split, bit 1 = P in tensor memory, bit 2 = ping-pong exponentiation,
// bit 3 = the row-split kernel further down, which is the default).
#ifndef AF_ATTN_PAIR_VARIANT
#define AF_ATTN_PAIR_VARIANT 8
#endif

// Every AF_ATTN_POLY_EVERY-th exponential of a row is evaluated on the FMA pipe (round-to-nearest split + degree-3
// minimax polynomial + exponent add, relative error 1e-4 << bf16 rounding of P) instead of the SFU: the ex2 unit
// (16 / clk / SM) is the roofline of the d = 40 / 80 self-attention layers.  0 disables.
#ifndef AF_ATTN_POLY_EVERY
#define AF_ATTN_POLY_EVERY 0
#endif

namespace af {

__device__ __forceinline__ float exp2_poly(float x) {   // x <= ~8 (lazy-rescale slack); flushes below 2^-126
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);      // in [-0.5, 0.5]
  float pl = fmaf(f, 0.05550410866f, 0.24022650695f);
  pl = fmaf(pl, f, 0.69314718056f);
  pl = fmaf(pl, f, 1.0f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}
// i is a compile-time constant after unrolling: the untaken branch folds away
__device__ __forceinline__ float exp2_mix(int i, float x) {
#if AF_ATTN_POLY_EVERY > 0
  if (i % AF_ATTN_POLY_EVERY == AF_ATTN_POLY_EVERY - 1) return exp2_poly(x);
#endif
  return fast_exp2(x);
}

struct AttnPairParams {
  long long* trace;  // optional device timeline of CTA (0,0,0): [4 actors][64 key blocks][8 events] clock64 stamps
  float* lse;        // optional [B][heads][Nq] log2-sum-exp per query row
  CUtensorMap tmQ;   // 3-D {heads*dp, Nq, B}, box {64, 128, 1}
  CUtensorMap tmK;   // 3-D {heads*dp, Nk, B}, box {64, BLOCK_N, 1}
  CUtensorMap tmV;   // 2-D {ldvt, heads*d}, box {64, DV}
  int B, heads, Nq, Nk;
  int d, dp, kv_stride;
  int pingpong;      // 1: the two softmax warps of an SM sub-partition take turns in the exponentiation phase
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

template <int D, bool SPLIT, bool PTMEM>
__global__ void __launch_bounds__(384, 1) attention_pair_kernel(const __grid_constant__ AttnPairParams p) {
  using C = PairCfg<D>;
  using S = PairSmem<D>;
  constexpr int DK = C::DK, DV = C::DV, BN = C::BLOCK_N;
  constexpr int KA = S::KA;
  // P is handed to the PV MMA in PA pieces of HK keys (SPLIT: one piece per 64-key swizzle atom)
  constexpr int PA = SPLIT ? S::PA : 1;
  constexpr int HK = BN / PA;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+DV) O1 [.., +DV) P0 [.., +BN/2) P1 [.., +BN/2) (bf16 pairs)
  constexpr bool kPTmem = PTMEM;
  constexpr uint32_t kTmemS[2] = {0, 128};
  constexpr uint32_t kTmemO[2] = {256, 256 + DV};
  constexpr uint32_t kTmemP[2] = {256 + 2 * DV, 256 + 2 * DV + BN / 2};
  static_assert(256 + 2 * DV + BN <= 512, "S, O and P of both tiles must fit in 512 TMEM columns");
  static_assert(DV % 16 == 0, "accumulator column alignment");
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
  uint64_t* p_full = s_full + 2;               // 2 x PA (tile, 64-key half): that half of P_j is in shared memory
  uint64_t* pv_done = p_full + 2 * PA;         // 2 x PA: the PV MMAs over that half are complete (half reusable)
  uint64_t* s_free = pv_done + 2 * PA;         // 2 (per tile): S_j copied to registers (TMEM buffer reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_blocks = (p.Nk + BN - 1) / BN;
  // timeline probe (af_attention_set_trace): actors 0 / 1 = softmax warps 4 / 8 (same SM sub-partition), 2 / 3 = MMA issuers
  const bool tracing = p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0;
  auto stamp = [&](int actor, int j, int ev) {
    if (tracing && j < 64) p.trace[(actor * 64 + j) * 8 + ev] = clock64();
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::KSTAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);    // one commit per query tile
    }
    for (int s = 0; s < C::VSTAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      for (int hf = 0; hf < PA; ++hf) {
        mbar_init(&p_full[t * PA + hf], 4);
        mbar_init(&pv_done[t * PA + hf], 1);
      }
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the tcgen05 issuers

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
        for (int a = 0; a < S::PA; ++a)
          tma_load_2d(smem + S::kVOff + vs * S::kVBytes + a * S::kVAtomBytes, &p.tmV, &v_full[vs],
                      b * p.kv_stride + j * BN + a * 64, h * p.d);
        if (++vs == C::VSTAGES) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> tile 0, warp 2 -> tile 1
    // One issuing thread per query tile, each walking its own in-order chain S_{j+1}, PV_j with blocking mbarrier
    // waits (a single event-polling thread for both tiles added ~1000 cycles between "P_j written" and "PV_j
    // complete": profiles/r02_attention_pair.md).  K / V ring slots are released by the second of the two commits.
    // The whole warp walks the chain (convergent, warp-uniform waits); one elected lane issues the tcgen05 instructions:
    // under `if (lane == 0)` the compiler wraps every tcgen05.mma / commit in an elect - issue - branch loop.
    {
      const int t = warp - 1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV);
      const uint32_t q_addr = smem_u32(smem + S::kQOff) + t * S::kQBytes;
      const uint32_t p_addr = smem_u32(smem + S::kPOff) + t * S::kPBytes;
      const uint32_t tm_s = tmem_base + kTmemS[t], tm_o = tmem_base + kTmemO[t], tm_p = tmem_base + kTmemP[t];
      auto issue_s = [&](int kslot) {   // S_t = Q_t . K^T for the K block at ring slot `kslot`
        const uint32_t k_addr = smem_u32(smem + S::kKOff + kslot * S::kKBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(q_addr + (k >> 2) * 128 * 128) + 2 * (k & 3);
            const uint64_t bd = umma_desc_sw128(k_addr + (k >> 2) * BN * 128) + 2 * (k & 3);
            tc_mma_ss(tm_s, ad, bd, idesc_s, k != 0 ? 1u : 0u);
          }
          tc_commit(&s_full[t]);
          tc_commit(&k_empty[kslot]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_blocks; ++j) {
        if (j + 1 < n_blocks) {
          const int slot = (j + 1) % C::KSTAGES;
          mbar_wait(&k_full[slot], ((j + 1) / C::KSTAGES) & 1);
          mbar_wait(&s_free[t], j & 1);          // S_j copied to registers
          tc_fence_after();
          stamp(2 + t, j, 0);
          issue_s(slot);
          stamp(2 + t, j, 1);
        }
        const int vslot = j % C::VSTAGES;
        mbar_wait(&v_full[vslot], (j / C::VSTAGES) & 1);
        const uint32_t v_addr = smem_u32(smem + S::kVOff + vslot * S::kVBytes);
#pragma unroll
        for (int hf = 0; hf < PA; ++hf) {
          mbar_wait(&p_full[t * PA + hf], j & 1);
          tc_fence_after();
          stamp(2 + t, j, 2 + 2 * hf);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < HK / 16; ++kk) {
              const int k = hf * (HK / 16) + kk;      // 16-key step within the block
              const uint64_t bd = umma_desc_sw128(v_addr + (k >> 2) * S::kVAtomBytes) + 2 * (k & 3);
              if constexpr (kPTmem) {
                tc_mma_ts(tm_o, tm_p + k * 8, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
              } else {
                const uint64_t ad = umma_desc_sw128(p_addr + (k >> 2) * 128 * 128) + 2 * (k & 3);
                tc_mma_ss(tm_o, ad, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
              }
            }
            tc_commit(&pv_done[t * PA + hf]);
            if (hf == PA - 1) tc_commit(&v_empty[vslot]);
          }
          __syncwarp();
          stamp(2 + t, j, 3 + 2 * hf);
        }
      }
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
    const uint32_t pt_addr = tmem_base + kTmemP[t] + lane_off;

    // The two warpgroups share the SFU (ex2) units of their SM sub-partitions: warp 4+q (tile 0) and warp 8+q (tile 1)
    // sit on sub-partition q.  Left alone they run IN PHASE (measured with af_attention_set_trace: both exponentiate
    // for ~2200 cycles at half rate each, then both spend ~1400 cycles in TMEM loads / maxima / barrier latency with
    // the ex2 unit idle) and the phase offset is only neutrally stable.  Ping-pong: the exponentiation phase is a
    // critical section per sub-partition, handed back and forth through two named barriers (1+q: tile 0's turn,
    // 5+q: tile 1's turn; 64 threads each: one warp syncs, the other arrives) so one warp's non-ex2 work always runs
    // under the other's ex2 phase.
    const bool pingpong = p.pingpong != 0 && n_blocks > 1;
    const int bar_mine = 1 + 4 * t + qd, bar_other = 1 + 4 * (1 - t) + qd;
    if (pingpong && t == 1) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");   // tile 0 goes first
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const int key0 = j * BN;
      const int actor = t;
      const bool tr_warp = qd == 0;
      if (tr_warp) stamp(actor, j, 0);
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (tr_warp) stamp(actor, j, 1);
      float sc[BN];
#pragma unroll
      for (int c = 0; c < BN; c += 32) tmem_ld32q(s_addr + c, reinterpret_cast<uint32_t*>(sc) + c);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);     // the MMA warp may overwrite S with the next block's scores
      if (tr_warp) stamp(actor, j, 2);
      if (key0 + BN > p.Nk || mrow != nullptr) {  // warp-uniform: tail block / explicit key mask only
#pragma unroll
        for (int e = 0; e < BN; ++e) {
          const int key = key0 + e;
          bool ok = key < p.Nk;
          if (mrow != nullptr) ok = ok && (__ldg(mrow + min(key, p.Nk - 1)) != 0);
          sc[e] = ok ? sc[e] : -INFINITY;
        }
      }
      float mx8[8];   // eight independent 3-input max chains (the phase is latency bound)
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(sc[2 * c], sc[2 * c + 1]);
#pragma unroll
      for (int e = 16; e < BN; e += 16)
#pragma unroll
        for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(mx8[c], fmaxf(sc[e + 2 * c], sc[e + 2 * c + 1]));
      const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                             fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      // lazy rescale: keep the old reference unless the max grew by more than 2^8 (first finite max always taken)
      float alpha = 1.0f;
      if (mx > m_ref + kRescaleThreshold) {
        alpha = fast_exp2(m_ref - mx);  // 0 when m_ref == -inf
        m_ref = mx;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      // O may only be corrected once every MMA of PV_{j-1} has landed (rare after the first blocks)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        mbar_wait(&pv_done[t * PA + PA - 1], (j - 1) & 1);
        tc_fence_after();
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
      if (pingpong) asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory");
      if (tr_warp) stamp(actor, j, 3);
#pragma unroll
      for (int hf = 0; hf < PA; ++hf) {
        // this piece of the P buffer is free once the PV_{j-1} MMAs that read it are complete
        if (j > 0) {
          mbar_wait(&pv_done[t * PA + hf], (j - 1) & 1);
          tc_fence_after();
        }
        if (tr_warp) stamp(actor, j, 4 + 2 * hf);
        if constexpr (kPTmem) {
#pragma unroll
          for (int c = hf * HK; c < hf * HK + HK; c += 64) {
            uint32_t pk[32];
#pragma unroll
            for (int e = 0; e < 64; e += 2) {
              const float p0 = exp2_mix(e, sc[c + e] - m_use);
              const float p1 = exp2_mix(e + 1, sc[c + e + 1] - m_use);
              l4[(e >> 1) & 1] += p0;
              l4[2 + ((e >> 1) & 1)] += p1;
              pk[e >> 1] = pack_bf16x2(p0, p1);
            }
            tmem_st32(pt_addr + c / 2, pk);
          }
          tmem_st_wait();
        } else {
#pragma unroll
          for (int c8 = hf * HK; c8 < hf * HK + HK; c8 += 8) {
            float pe[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              pe[e] = exp2_mix(e, sc[c8 + e] - m_use);
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
          fence_async_smem();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t * PA + hf]);
        if (tr_warp) stamp(actor, j, 5 + 2 * hf);
      }
      if (pingpong && !(t == 1 && j == n_blocks - 1)) asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory");
      l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
    }
    // epilogue: O / l -> bf16
    mbar_wait(&pv_done[t * PA + PA - 1], (n_blocks - 1) & 1);
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


// =====================================================================================================================
// Row-split schedule: TWO softmax threads per query row (each owns half of the key block's columns), i.e. four softmax
// warps per SM sub-partition instead of two.
//
// Why (device timeline, scripts/attn_trace.py, profiles/r02_attention_rowsplit.md): with one thread per row the two
// softmax warps of a sub-partition either run in phase - 2 x 128 ex2 per lane saturate the SFU for ~2200 cycles, then
// both spend ~1400 cycles in TMEM loads, maxima and barrier latency with the SFU idle - or, made to take turns, a
// single warp cannot keep the SFU busy on its own (1550 cycles for 128 ex2 instead of 1024).  Either way the kernel
// sits at ~50 % of the ex2 roofline.  With four warps the ex2 demand per key block (2048 SFU cycles per sub-partition)
// exceeds each warp's own latency chain, so the SFU stays saturated.
//
//   warp 0        TMA producer (Q0, Q1 once; K / V ring)         warps 1, 2   MMA issuers of tile 0 / tile 1
//   warps 4-19    softmax: warpgroup wg = (warp - 4) / 4 -> tile wg / 2, column half wg % 2; TMEM lane quarter warp % 4
//
// The two threads of a row exchange their local maxima through shared memory behind a 64-thread named barrier (both
// warps live on the same sub-partition), keep the same running reference, sum their own part of l and combine it
// once at the end.  P goes through tensor memory; each half is handed to the PV MMA on its own (p_full / pv_done per
// half), and only the half-0 thread corrects O on a (rare) reference change - the in-order MMA issuer waits for half 0
// first, so the correction is always complete before any PV MMA of the block is issued.
template <int D>
__global__ void __launch_bounds__(640, 1) attention_rowsplit_kernel(const __grid_constant__ AttnPairParams p) {
  using C = PairCfg<D>;
  using S = PairSmem<D>;
  constexpr int DK = C::DK, DV = C::DV, BN = C::BLOCK_N;
  constexpr int KA = S::KA;
  constexpr int HC = BN / 2;                    // score columns per softmax thread
  static_assert(HC == 32 || HC == 64, "column half must be one or two 32-column TMEM loads");
  constexpr uint32_t kTmemS[2] = {0, 128};
  constexpr uint32_t kTmemO[2] = {256, 256 + DV};
  constexpr uint32_t kTmemP[2] = {256 + 2 * DV, 256 + 2 * DV + BN / 2};
  static_assert(256 + 2 * DV + BN <= 512, "S, O and P of both tiles must fit in 512 TMEM columns");
  constexpr float kRescaleThreshold = 8.0f;     // log2 domain
  // shared memory: Q (2 tiles) | K ring | V ring | exchange floats | barriers   (no P buffer)
  constexpr int kXchOff = S::kPOff;                       // [2 parities][2 tiles][2 halves][128] maxima + [2][2][128] sums
  constexpr int kBarOff = kXchOff + (8 + 4) * 128 * 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* xch = reinterpret_cast<float*>(smem + kXchOff);
  float* lxch = xch + 8 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = q_full + 1;               // KSTAGES
  uint64_t* k_empty = k_full + C::KSTAGES;
  uint64_t* v_full = k_empty + C::KSTAGES;     // VSTAGES
  uint64_t* v_empty = v_full + C::VSTAGES;
  uint64_t* s_full = v_empty + C::VSTAGES;     // 2 (per tile)
  uint64_t* s_free = s_full + 2;               // 2 (per tile): both column halves copied to registers
  uint64_t* p_full = s_free + 2;               // 2 x 2 (tile, half)
  uint64_t* pv_done = p_full + 4;              // 2 x 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_blocks = (p.Nk + BN - 1) / BN;
  const bool tracing = p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0;
  auto stamp = [&](int actor, int j, int ev) {
    if (tracing && j < 64) p.trace[(actor * 64 + j) * 8 + ev] = clock64();
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::KSTAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);
    }
    for (int s = 0; s < C::VSTAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 8);
      for (int hf = 0; hf < 2; ++hf) {
        mbar_init(&p_full[t * 2 + hf], 4);
        mbar_init(&pv_done[t * 2 + hf], 1);
      }
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the tcgen05 issuers

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
          for (int a = 0; a < S::PA; ++a)
            tma_load_2d(smem + S::kVOff + vs * S::kVBytes + a * S::kVAtomBytes, &p.tmV, &v_full[vs],
                        b * p.kv_stride + j * BN + a * 64, h * p.d);
          if (++vs == C::VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp <= 2) {
      // ------------------------------------------------------------------ MMA issuer of tile t (in-order chain)
      {   // whole warp, convergent; tcgen05 instructions by one elected lane (see attention_pair_kernel)
        const int t = warp - 1;
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV);
        const uint32_t q_addr = smem_u32(smem + S::kQOff) + t * S::kQBytes;
        const uint32_t tm_s = tmem_base + kTmemS[t], tm_o = tmem_base + kTmemO[t], tm_p = tmem_base + kTmemP[t];
        auto issue_s = [&](int kslot) {
          const uint32_t k_addr = smem_u32(smem + S::kKOff + kslot * S::kKBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DK / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(q_addr + (k >> 2) * 128 * 128) + 2 * (k & 3);
              const uint64_t bd = umma_desc_sw128(k_addr + (k >> 2) * BN * 128) + 2 * (k & 3);
              tc_mma_ss(tm_s, ad, bd, idesc_s, k != 0 ? 1u : 0u);
            }
            tc_commit(&s_full[t]);
            tc_commit(&k_empty[kslot]);
          }
          __syncwarp();
        };
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        issue_s(0);
        for (int j = 0; j < n_blocks; ++j) {
          if (j + 1 < n_blocks) {
            const int slot = (j + 1) % C::KSTAGES;
            mbar_wait(&k_full[slot], ((j + 1) / C::KSTAGES) & 1);
            mbar_wait(&s_free[t], j & 1);
            tc_fence_after();
            issue_s(slot);
          }
          const int vslot = j % C::VSTAGES;
          mbar_wait(&v_full[vslot], (j / C::VSTAGES) & 1);
          const uint32_t v_addr = smem_u32(smem + S::kVOff + vslot * S::kVBytes);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            mbar_wait(&p_full[t * 2 + hf], j & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < HC / 16; ++kk) {
                const int k = hf * (HC / 16) + kk;
                const uint64_t bd = umma_desc_sw128(v_addr + (k >> 2) * S::kVAtomBytes) + 2 * (k & 3);
                tc_mma_ts(tm_o, tm_p + k * 8, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
              }
              tc_commit(&pv_done[t * 2 + hf]);
              if (hf == 1) tc_commit(&v_empty[vslot]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------------------------------------------ softmax: (tile, column half, lane quarter)
    const int wg = (warp - 4) >> 2;
    const int t = wg >> 1, half = wg & 1;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int q_row = q0 + t * 128 + r;
    const uint8_t* mrow = p.key_mask ? p.key_mask + static_cast<size_t>(b) * p.Nk : nullptr;
    const uint32_t s_addr = tmem_base + kTmemS[t] + lane_off + half * HC;
    const uint32_t o_addr = tmem_base + kTmemO[t] + lane_off;
    const uint32_t pt_addr = tmem_base + kTmemP[t] + lane_off + half * (HC / 2);
    const int bar_id = 1 + t * 4 + qd;            // the two warps that share rows [qd*32, qd*32+32) of tile t
    float* x_mine = xch + (t * 2 + half) * 128 + r;
    float* x_other = xch + (t * 2 + (half ^ 1)) * 128 + r;
    // Ping-pong between the two TILES (bit 2 of the variant): the ex2 phase of a tile (its two column-half warps per
    // sub-partition saturate the SFU together) alternates with the other tile's, whose TMEM loads, maxima and MMA
    // round trips then run underneath.  Named barriers 9 / 10 = "tile 0's / tile 1's turn": the 8 warps of the tile
    // sync, the 8 warps of the other tile arrive (512 threads).
    const bool pingpong = p.pingpong != 0 && n_blocks > 1;
    if (pingpong && t == 1) asm volatile("bar.arrive 9, 512;" ::: "memory");    // tile 0 goes first
    const bool tr_warp = qd == 0 && half == 0;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const int key0 = j * BN + half * HC;
      if (tr_warp) stamp(t, j, 0);
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (tr_warp) stamp(t, j, 1);
      float sc[HC];
#pragma unroll
      for (int c = 0; c < HC; c += 32) tmem_ld32q(s_addr + c, reinterpret_cast<uint32_t*>(sc) + c);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      if (tr_warp) stamp(t, j, 2);
      if (j * BN + BN > p.Nk || mrow != nullptr) {  // warp-uniform: tail block / explicit key mask only
#pragma unroll
        for (int e = 0; e < HC; ++e) {
          const int key = key0 + e;
          bool ok = key < p.Nk;
          if (mrow != nullptr) ok = ok && (__ldg(mrow + min(key, p.Nk - 1)) != 0);
          sc[e] = ok ? sc[e] : -INFINITY;
        }
      }
      float mx8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(sc[2 * c], sc[2 * c + 1]);
#pragma unroll
      for (int e = 16; e < HC; e += 16)
#pragma unroll
        for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(mx8[c], fmaxf(sc[e + 2 * c], sc[e + 2 * c + 1]));
      float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                       fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      // row maximum over both column halves (alternating slots: the partner reads slot j&1 before it can reach the
      // barrier of block j+1, and the slot is rewritten at block j+2)
      x_mine[(j & 1) * 512] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, x_other[(j & 1) * 512]);
      float alpha = 1.0f;
      if (mx > m_ref + kRescaleThreshold) {   // identical decision in both threads of the row
        alpha = fast_exp2(m_ref - mx);
        m_ref = mx;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      if (half == 0 && j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        mbar_wait(&pv_done[t * 2 + 1], (j - 1) & 1);   // every MMA of PV_{j-1} has landed
        tc_fence_after();
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
      if (tr_warp) stamp(t, j, 3);
      if (j > 0) {                                     // my half of the P buffer has been consumed
        mbar_wait(&pv_done[t * 2 + half], (j - 1) & 1);
        tc_fence_after();
      }
      if (tr_warp) stamp(t, j, 4);
      if (pingpong) {
        if (t == 0) asm volatile("bar.sync 9, 512;" ::: "memory");
        else asm volatile("bar.sync 10, 512;" ::: "memory");
      }
      if (tr_warp) stamp(t, j, 5);
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < HC; c += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = exp2_mix(e, sc[c + e] - m_use);
          const float p1 = exp2_mix(e + 1, sc[c + e + 1] - m_use);
          l4[(e >> 1) & 1] += p0;
          l4[2 + ((e >> 1) & 1)] += p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st16q(pt_addr + c / 2, pk);
      }
      // the turn is handed over as soon as the last ex2 has been issued: the TMEM-store drain, fences and the
      // mbarrier arrive below run under the other tile's ex2 phase
      if (pingpong) {
        if (t == 0) asm volatile("bar.arrive 10, 512;" ::: "memory");
        else if (j != n_blocks - 1) asm volatile("bar.arrive 9, 512;" ::: "memory");
      }
      if (tr_warp) stamp(t, j, 6);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t * 2 + half]);
      l_run = l_run * alpha + ((l4[0] + l4[1]) + (l4[2] + l4[3]));
    }
    // epilogue: l = l(half 0) + l(half 1); O / l -> bf16, 16-column chunks alternate between the two threads of the row
    lxch[(t * 2 + half) * 128 + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    const float l_tot = l_run + lxch[(t * 2 + (half ^ 1)) * 128 + r];
    mbar_wait(&pv_done[t * 2 + 1], (n_blocks - 1) & 1);
    tc_fence_after();
    const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    if (half == 0 && p.lse != nullptr && q_row < p.Nq)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Nq + q_row] = l_tot > 0.f ? m_ref + __log2f(l_tot) : INFINITY;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.Nq + q_row) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = half * 16; c < DV; c += 32) {
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

static long long* g_pair_trace = nullptr;
int attention_variant();
static int g_pair_variant = AF_ATTN_PAIR_VARIANT;   // bit 0: split hand-off, bit 1: P in tensor memory, bit 2: ping-pong

template <int D, bool SPLIT, bool PTMEM>
static int launch_pair_v(const AttnPairParams& p, cudaStream_t stream) {
  using S = PairSmem<D>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(attention_pair_kernel<D, SPLIT, PTMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S::kTotal));
    configured = true;
  }
  dim3 grid((p.Nq + 255) / 256, p.heads, p.B);
  attention_pair_kernel<D, SPLIT, PTMEM><<<grid, 384, S::kTotal, stream>>>(p);
  AF_LAUNCH_CHECK("attention_pair_kernel");
  return 0;
}
template <int D>
static int launch_rowsplit(const AttnPairParams& p, cudaStream_t stream) {
  using S = PairSmem<D>;
  constexpr int kSmem = S::kPOff + 12 * 128 * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(attention_rowsplit_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  dim3 grid((p.Nq + 255) / 256, p.heads, p.B);
  attention_rowsplit_kernel<D><<<grid, 640, kSmem, stream>>>(p);
  AF_LAUNCH_CHECK("attention_rowsplit_kernel");
  return 0;
}
template <int D>
static int launch_pair(const AttnPairParams& p, cudaStream_t stream) {
  if (g_pair_variant & 8) return launch_rowsplit<D>(p, stream);
  switch (g_pair_variant & 3) {
    case 0: return launch_pair_v<D, false, false>(p, stream);
    case 1: return launch_pair_v<D, true, false>(p, stream);
    case 2: return launch_pair_v<D, false, true>(p, stream);
    default: return launch_pair_v<D, true, true>(p, stream);
  }
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
  p.trace = g_pair_trace;
  p.pingpong = (g_pair_variant >> 2) & 1;
  p.out = static_cast<__nv_bfloat16*>(O);
  p.ldo = static_cast<long long>(heads) * d;
  return d == 40 ? launch_pair<40>(p, stream) : launch_pair<80>(p, stream);
}

int attention_variant() { return g_pair_variant; }
long long* attention_trace_ptr() { return g_pair_trace; }

}  // namespace af

extern "C" int af_attention_set_trace(long long* device_buffer) {
  af::g_pair_trace = device_buffer;
  return 0;
}

extern "C" int af_attention_set_pair_variant(int variant) {
  const int old = af::g_pair_variant;
  if (variant >= 0) af::g_pair_variant = variant & 0xffff;
  return old;
}
