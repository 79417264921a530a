// Self-attention for the long-sequence layers (N >= 256 keys, head dims 40 / 80): one 128-query tile per CTA,
// TWO (d = 40) CTAs per SM.
//
// Reference semantics: CrossAttention.forward (ldm/modules/attention.py:172-243), same operand conventions as
// attention.cu (Q pre-scaled into the exp2 domain, K-major bf16 tiles, V transposed).
//
// Why this shape (profiles/r02_attention.md).  At d = 40 one exponential buys only 160 FLOP, so the layer is bound by
// the ex2 unit (4 results / clk / SM sub-partition), and what the round-1 kernels lost was not ex2 throughput but the
// serial chain of every softmax warp (S ready -> TMEM load -> row max -> P free -> exponentials -> hand-over: ~1400
// cycles per key block with the ex2 unit idle) while all the warps of the one resident CTA ran that chain IN PHASE.
// Here
//   * two independent CTAs share an SM: nothing couples their phases, so one CTA's chain runs under the other CTA's
//     exponentials (the same latency hiding a plain occupancy-2 kernel gets, with warp-specialised CTAs);
//   * the row sum l is not accumulated by the softmax threads: row D of the V^T tile in shared memory is a constant
//     row of ones (the TMA box covers rows 0..D-1 only and never overwrites it), so column D of the O accumulator IS
//     sum_j P_ij - computed by the tensor core from exactly the bf16 probabilities the numerator uses.  That removes
//     one FADD per score (a quarter of the softmax instruction stream);
//   * no run-time options, trace probes or generic-address shared-memory accesses in the loop: the round-1 row-split
//     kernel issued 580 instructions per warp and key block for ~320 useful ones and spilled the running max and sum.
//
//   warp 0      TMA producer (Q once; K / V ring)          warp 1   MMA issuer (S_{j+1} early, then PV_j)
//   warps 2, 3  idle (setmaxnreg works per warpgroup)      warps 4.. softmax: SPLIT threads per query row, each owning
//                                                                   BLOCK_N / SPLIT score columns (TMEM lane = row)
// P goes to the PV MMA through tensor memory (tcgen05.st, A operand in TMEM).  Lazy rescale: the reference max is only
// moved when the block max exceeds it by more than 2^8, so the O correction almost never runs after the first blocks.
#include <math.h>
#include <type_traits>

#include "../../include/adaface_b200.h"
#include "common.cuh"

namespace af {

struct AttnTileParams {
  long long* trace;  // TRACE instantiation only: [4 actors][64 key blocks][8 events] clock64 stamps (scripts/attn_tile_trace.py)
  float* lse;        // optional [B][heads][Nq] log2-sum-exp per query row
  CUtensorMap tmQ;   // 3-D {heads*dp, Nq, B}, box {64, 128, 1}
  CUtensorMap tmK;   // 3-D {heads*dp, Nk, B}, box {64, BLOCK_N, 1}
  CUtensorMap tmV;   // 2-D {ldvt, heads*d}, box {64, D}   (D rows: rows D.. of the shared-memory tile are constants)
  int B, heads, Nq, Nk;
  int d, dp, kv_stride;
  const uint8_t* key_mask;
  __nv_bfloat16* out;
  long long ldo;
  int first_wave;    // CTAs with a linear id below this start together (2 x SM count)
  int stagger;       // cycles the second CTA of an SM holds its softmax warps back in the first wave (0 = off)
};

// Arrival counter per SM (monotonic across launches: only the parity of consecutive arrivals on one SM is used).
__device__ unsigned int g_tile_sm_arrivals[1024];

template <int D>
struct TileCfg;
template <>
struct TileCfg<40> {
  // TMEM: S [0,128) | O [128,176) | P [192,256)
  static constexpr int DK = 48, DV = 48, BLOCK_N = 128, KSTAGES = 3, VSTAGES = 3;
  static constexpr uint32_t kTmemS = 0, kTmemO = 128, kTmemP = 192, kTmemCols = 256;
};
template <>
struct TileCfg<80> {
  // TMEM: S [0,64) | O [64,160) | P [160,192)
  static constexpr int DK = 80, DV = 96, BLOCK_N = 64, KSTAGES = 3, VSTAGES = 2;
  static constexpr uint32_t kTmemS = 0, kTmemO = 64, kTmemP = 160, kTmemCols = 256;
};

template <int D, int SPLIT>
struct TileSmem {
  using C = TileCfg<D>;
  static constexpr int KA = (C::DK + 63) / 64;            // 64-column swizzle atoms along the head dim
  static constexpr int PA = C::BLOCK_N / 64;              // 64-key swizzle atoms along the key dim
  static constexpr int kQBytes = KA * 128 * 128;
  static constexpr int kKBytes = KA * C::BLOCK_N * 128;
  static constexpr int kVAtomBytes = C::DV * 128;         // DV rows of 64 keys
  static constexpr int kVBytes = PA * kVAtomBytes;
  static constexpr int kVTxBytes = PA * D * 128;          // what the TMA box {64, D} delivers per stage
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + kQBytes;
  static constexpr int kVOff = kKOff + C::KSTAGES * kKBytes;
  static constexpr int kXchOff = kVOff + C::VSTAGES * kVBytes;     // [2 parities][SPLIT][128] row maxima
  static constexpr int kBarOff = kXchOff + 2 * SPLIT * 128 * 4;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static_assert(kVAtomBytes % 1024 == 0 && kKBytes % 1024 == 0 && kQBytes % 1024 == 0, "swizzle atoms are 1024-byte aligned");
};

__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// lean wait for the hot loops: no printf / globaltimer in the instruction stream, a protocol bug still traps
__device__ __forceinline__ void mbar_wait_lean(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(addr, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) { mbar_wait_lean(smem_u32(bar), parity); }
__device__ __forceinline__ void tile_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tile_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 2^x on the FMA pipe: round-to-nearest split, degree-3 minimax on [-0.5, 0.5], exponent add (rel. error 1e-4, far
// below the bf16 rounding of P).  x <= ~8 (lazy-rescale slack); flushes below 2^-126.
__device__ __forceinline__ float tile_exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float pl = fmaf(f, 0.05550410866f, 0.24022650695f);
  pl = fmaf(pl, f, 0.69314718056f);
  pl = fmaf(pl, f, 1.0f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}
template <int POLY>
__device__ __forceinline__ float tile_exp2(int i, float x) {   // i is a compile-time constant after unrolling
  if constexpr (POLY > 0) {
    if (i % POLY == POLY - 1) return tile_exp2_poly(x);
  }
  return fast_exp2(x);
}

// SPLIT: softmax threads per query row (1 or 2).  POLY: every POLY-th exponential on the FMA pipe (0 = none).
// MASKED: instantiation for an explicit key mask or a ragged last key block (the UNet's own shapes need neither).
template <int D, int SPLIT, int POLY, bool MASKED, bool WIDE, bool TRACE = false>
__global__ void __launch_bounds__(128 + 128 * SPLIT, 2) attention_tile_kernel(const __grid_constant__ AttnTileParams p) {
  using C = TileCfg<D>;
  using S = TileSmem<D, SPLIT>;
  constexpr int DK = C::DK, DV = C::DV, BN = C::BLOCK_N;
  constexpr int KA = S::KA, PA = S::PA;
  constexpr int HC = BN / SPLIT;                    // score columns per softmax thread
  static_assert(HC % 32 == 0, "a softmax thread owns whole 32-column TMEM loads");
  static_assert(D % 8 == 0 && DV % 16 == 0 && DV >= D + 8, "V^T tile: D data rows, then the ones row, then zero rows");
  static_assert(C::kTmemO + DV <= C::kTmemP && C::kTmemP + BN / 2 <= C::kTmemCols, "TMEM column map");
  // Lazy reference (log2 domain): the reference only moves when a block maximum exceeds it by more than the window.
  // WIDE: the window is 2^64 and the first reference is 0 whenever the first block's maximum lies inside it - bf16 P
  // and the fp32 accumulators have the exponent range to spare (P <= 2^64, O <= 2^64 * 4096 * |v|), scores more than
  // 2^-62 below the row maximum are irrelevant - so for the UNet's score range the reference stays 0 and the
  // subtraction (one FADD per score, issued before the first MUFU of the block) disappears from the loop.
  constexpr float kRescaleThreshold = WIDE ? 64.0f : 8.0f;
  constexpr int kThreads = 128 + 128 * SPLIT;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = q_full + 1;               // KSTAGES
  uint64_t* k_empty = k_full + C::KSTAGES;
  uint64_t* v_full = k_empty + C::KSTAGES;     // VSTAGES
  uint64_t* v_empty = v_full + C::VSTAGES;
  uint64_t* s_full = v_empty + C::VSTAGES;     // 1: S_j in TMEM
  uint64_t* s_free = s_full + 1;               // 1: S_j copied to registers by every softmax warp
  uint64_t* p_full = s_free + 1;               // SPLIT: that column range of P_j is in TMEM
  uint64_t* pv_done = p_full + SPLIT;          // SPLIT: the PV MMAs over that column range are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + SPLIT);
  uint32_t* arrival_slot = tmem_slot + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_blocks = (p.Nk + BN - 1) / BN;
  // timeline probe (TRACE instantiation, scripts/attn_tile_trace.py): softmax warp 4 of the first 512 CTAs (linear id)
  // -> [cta][64 key blocks][8 events]; then 512 SM ids; then MMA issuer (actor 2) and producer (actor 3) of CTA 0
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  auto stamp = [&](int actor, int j, int ev) {
    if constexpr (TRACE) {
      if (lane == 0 && j < 64) {
        if (actor < 2) {
          if (cta_lin < 512) p.trace[(cta_lin * 64 + j) * 8 + ev] = clock64();
        } else if (cta_lin == 0) {
          p.trace[512 * 64 * 8 + 512 + ((actor - 2) * 64 + j) * 8 + ev] = clock64();
        }
      }
    }
  };
  if constexpr (TRACE) {
    if (threadIdx.x == 0 && cta_lin < 512) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[512 * 64 * 8 + cta_lin] = smid;
    }
  }

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
    mbar_init(s_full, 1);
    mbar_init(s_free, 4 * SPLIT);
    for (int hf = 0; hf < SPLIT; ++hf) {
      mbar_init(&p_full[hf], 4);
      mbar_init(&pv_done[hf], 1);
    }
    mbar_fence_init();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    *arrival_slot = p.stagger > 0 ? atomicAdd(&g_tile_sm_arrivals[smid & 1023], 1u) : 0u;
  }
  // constant rows of every V^T stage: row D = 1.0 (column D of O becomes the row sum), rows D+1 .. DV-1 = 0.
  // A row of equal values is invariant under the 128-byte swizzle; the TMA box never touches these rows.
  {
    constexpr int kConstChunks = (DV - D) * 8;       // 16-byte chunks per atom
    for (int i = threadIdx.x; i < C::VSTAGES * PA * kConstChunks; i += kThreads) {
      const int atom = i / kConstChunks, chunk = i % kConstChunks;
      const uint32_t addr = smem_base + S::kVOff + (atom / PA) * S::kVBytes + (atom % PA) * S::kVAtomBytes + D * 128 + chunk * 16;
      const uint32_t v = chunk < 8 ? 0x3F803F80u : 0u;
      sts128(addr, v, v, v, v);
    }
    fence_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, S::kQBytes);
#pragma unroll
        for (int a = 0; a < KA; ++a)
          tma_load_3d(smem + S::kQOff + a * 128 * 128, &p.tmQ, q_full, h * p.dp + a * 64, q0, b);
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        for (int j = 0; j < n_blocks; ++j) {
          stamp(3, j, 0);
          mbar_wait_lean(&k_empty[ks], kph ^ 1);
          stamp(3, j, 1);
          mbar_arrive_expect_tx(&k_full[ks], S::kKBytes);
#pragma unroll
          for (int a = 0; a < KA; ++a)
            tma_load_3d(smem + S::kKOff + ks * S::kKBytes + a * BN * 128, &p.tmK, &k_full[ks], h * p.dp + a * 64,
                        j * BN, b);
          if (++ks == C::KSTAGES) { ks = 0; kph ^= 1; }
          stamp(3, j, 2);
          mbar_wait_lean(&v_empty[vs], vph ^ 1);
          stamp(3, j, 3);
          mbar_arrive_expect_tx(&v_full[vs], S::kVTxBytes);
#pragma unroll
          for (int a = 0; a < PA; ++a)
            tma_load_2d(smem + S::kVOff + vs * S::kVBytes + a * S::kVAtomBytes, &p.tmV, &v_full[vs],
                        b * p.kv_stride + j * BN + a * 64, h * p.d);
          if (++vs == C::VSTAGES) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer (whole warp walks the chain,
      // one elected lane issues): S_0, then per key block S_{j+1} as soon as S_j has been copied out, then PV_j
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV);
      const uint32_t q_addr = smem_base + S::kQOff;
      const uint32_t tm_s = tmem_base + C::kTmemS, tm_o = tmem_base + C::kTmemO, tm_p = tmem_base + C::kTmemP;
      auto issue_s = [&](int kslot) {
        const uint32_t k_addr = smem_base + S::kKOff + kslot * S::kKBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(q_addr + (k >> 2) * 128 * 128) + 2 * (k & 3);
            const uint64_t bd = umma_desc_sw128(k_addr + (k >> 2) * BN * 128) + 2 * (k & 3);
            tc_mma_ss(tm_s, ad, bd, idesc_s, k != 0 ? 1u : 0u);
          }
          tc_commit(s_full);
          tc_commit(&k_empty[kslot]);
        }
        __syncwarp();
      };
      mbar_wait_lean(q_full, 0);
      mbar_wait_lean(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      int kslot = 0, vslot = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < n_blocks; ++j) {
        if (j + 1 < n_blocks) {
          if (++kslot == C::KSTAGES) { kslot = 0; kph ^= 1; }
          stamp(2, j, 0);
          mbar_wait_lean(&k_full[kslot], kph);
          stamp(2, j, 1);
          mbar_wait_lean(s_free, j & 1);
          tc_fence_after();
          stamp(2, j, 2);
          issue_s(kslot);
          stamp(2, j, 3);
        }
        mbar_wait_lean(&v_full[vslot], vph);
        stamp(2, j, 4);
        const uint32_t v_addr = smem_base + S::kVOff + vslot * S::kVBytes;
#pragma unroll
        for (int hf = 0; hf < SPLIT; ++hf) {
          mbar_wait_lean(&p_full[hf], j & 1);
          tc_fence_after();
          if (hf == 0) stamp(2, j, 5);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < HC / 16; ++kk) {
              const int k = hf * (HC / 16) + kk;      // 16-key step within the block
              const uint64_t bd = umma_desc_sw128(v_addr + (k >> 2) * S::kVAtomBytes) + 2 * (k & 3);
              tc_mma_ts(tm_o, tm_p + k * 8, bd, idesc_o, (j | k) != 0 ? 1u : 0u);
            }
            tc_commit(&pv_done[hf]);
            if (hf == SPLIT - 1) tc_commit(&v_empty[vslot]);
          }
          __syncwarp();
          if (hf == SPLIT - 1) stamp(2, j, 6);
        }
        if (++vslot == C::VSTAGES) { vslot = 0; vph ^= 1; }
      }
    }
  } else {
    if constexpr (SPLIT == 2) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    }
    // ------------------------------------------------------------------ softmax: (column range `half`, lane quarter)
    const int half = (warp - 4) >> 2;
    const int qd = warp & 3;                      // TMEM lane quarter (hardware: warp id % 4)
    const int r = qd * 32 + lane;                 // row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int q_row = q0 + r;
    const uint8_t* mrow = p.key_mask ? p.key_mask + static_cast<size_t>(b) * p.Nk : nullptr;
    const uint32_t s_addr = tmem_base + C::kTmemS + lane_off + half * HC;
    const uint32_t o_addr = tmem_base + C::kTmemO + lane_off;
    const uint32_t pt_addr = tmem_base + C::kTmemP + lane_off + half * (HC / 2);
    const uint32_t a_s_full = smem_u32(s_full), a_s_free = smem_u32(s_free);
    const uint32_t a_p_full = smem_u32(&p_full[half]), a_pv_mine = smem_u32(&pv_done[half]);
    const uint32_t a_pv_last = smem_u32(&pv_done[SPLIT - 1]);
    const uint32_t x_mine = smem_base + S::kXchOff + (half * 128 + r) * 4;
    const uint32_t x_other = smem_base + S::kXchOff + ((half ^ (SPLIT - 1)) * 128 + r) * 4;
    const int bar_id = 1 + qd;                    // the SPLIT warps that share rows [qd*32, qd*32+32)
    // Phase stagger.  The two CTAs of an SM share its ex2 units; started together they stay IN PHASE (both exponentiate
    // at half rate, then both walk their synchronisation chain with the unit idle: profiles/r02_attention.md), and the
    // offset between two such loops is neutrally stable.  So the second arrival of the first wave starts half a key
    // block late, and every later CTA inherits the offset of the CTA whose slot it takes over.
    if (p.stagger > 0 && cta_lin < p.first_wave && (*arrival_slot & 1u)) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger) __nanosleep(64);
    }
    float m_ref = -INFINITY;
    uint32_t par = 0;                             // j & 1
    const int actor = 0;
    const bool tw = TRACE && warp == 4;
    for (int j = 0; j < n_blocks; ++j, par ^= 1) {
      if (tw) stamp(actor, j, 0);
      mbar_wait_lean(a_s_full, par);
      tc_fence_after();
      if (tw) stamp(actor, j, 1);
      float sc[HC];
#pragma unroll
      for (int c = 0; c < HC; c += 32) tile_ld32(s_addr + c, reinterpret_cast<uint32_t*>(sc) + c);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_s_free) : "memory");
      if (tw) stamp(actor, j, 2);
      if constexpr (MASKED) {
        // keep-bits of my HC keys, 32 per ballot (lane l probes key key0 + c + l), then compile-time bit tests
        const int key0 = j * BN + half * HC;
#pragma unroll
        for (int c = 0; c < HC; c += 32) {
          const int key = key0 + c + lane;
          bool ok = key < p.Nk;
          if (ok && mrow != nullptr) ok = __ldg(mrow + key) != 0;
          const uint32_t keep = __ballot_sync(0xffffffffu, ok);
#pragma unroll
          for (int e = 0; e < 32; ++e) sc[c + e] = ((keep >> e) & 1u) ? sc[c + e] : -INFINITY;
        }
      }
      float mx8[8];   // eight independent 3-input max chains
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(sc[2 * c], sc[2 * c + 1]);
#pragma unroll
      for (int e = 16; e < HC; e += 16)
#pragma unroll
        for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(mx8[c], fmaxf(sc[e + 2 * c], sc[e + 2 * c + 1]));
      float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                       fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      if constexpr (SPLIT == 2) {
        // row maximum over both column ranges (alternating slots: the partner reads slot `par` before it can reach the
        // barrier of block j+1, and the slot is rewritten at block j+2)
        sts32f(x_mine + par * (SPLIT * 128 * 4), mx);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        mx = fmaxf(mx, lds32f(x_other + par * (SPLIT * 128 * 4)));
      }
      // lazy rescale: keep the old reference unless the max grew by more than 2^8 (first finite max always taken);
      // identical decision in every thread of the row
      float alpha = 1.0f;
      if (mx > m_ref + kRescaleThreshold) {   // also the first finite maximum (m_ref == -inf)
        const float m_new = (WIDE && m_ref == -INFINITY && fabsf(mx) <= kRescaleThreshold) ? 0.f : mx;
        alpha = fast_exp2(m_ref - m_new);     // 0 when m_ref == -inf
        m_ref = m_new;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      // O (and its row-sum column) is corrected by the range-0 thread once every MMA of PV_{j-1} has landed; the in-order
      // issuer waits for range 0's P first, so the correction is complete before any PV MMA of this block is issued
      if (half == 0 && j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        mbar_wait_lean(a_pv_last, par ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < DV; c += 16) {
          uint32_t o[16];
          tmem_ld16(o_addr + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tile_st16(o_addr + c, o);
        }
        tmem_st_wait();
      }
      if (tw) stamp(actor, j, 3);
      if (j > 0) {                                  // my range of the P buffer has been consumed by PV_{j-1}
        mbar_wait_lean(a_pv_mine, par ^ 1);
        tc_fence_after();
      }
      if (tw) stamp(actor, j, 4);
      auto exp_block = [&](auto sub) {
#pragma unroll
        for (int c = 0; c < HC; c += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = tile_exp2<POLY>(e, decltype(sub)::value ? sc[c + e] - m_use : sc[c + e]);
            const float p1 = tile_exp2<POLY>(e + 1, decltype(sub)::value ? sc[c + e + 1] - m_use : sc[c + e + 1]);
            pk[e >> 1] = pack_bf16x2(p0, p1);
          }
          tile_st16(pt_addr + c / 2, pk);
        }
      };
      if (WIDE && __all_sync(0xffffffffu, m_ref == 0.f)) exp_block(std::false_type{});
      else exp_block(std::true_type{});
      if (tw) stamp(actor, j, 5);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_p_full) : "memory");
      if (tw) stamp(actor, j, 6);
    }
    // epilogue: l = column D of O; O / l -> bf16, 16-column chunks alternate between the threads of the row
    mbar_wait_lean(a_pv_last, (n_blocks - 1) & 1);
    tc_fence_after();
    float l_tot;
    {
      uint32_t lv[8];
      tmem_ld8(o_addr + D, lv);
      tmem_ld_wait();
      l_tot = __uint_as_float(lv[0]);
    }
    const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    if (half == 0 && p.lse != nullptr && q_row < p.Nq)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Nq + q_row] = l_tot > 0.f ? m_ref + __log2f(l_tot) : INFINITY;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.Nq + q_row) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = half * 16; c < D; c += 16 * SPLIT) {
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int D, int SPLIT, int POLY, bool MASKED, bool WIDE, bool TRACE = false>
static int launch_tile_m(const AttnTileParams& p, cudaStream_t stream) {
  using S = TileSmem<D, SPLIT>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(attention_tile_kernel<D, SPLIT, POLY, MASKED, WIDE, TRACE>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    AF_CUDA(cudaFuncSetAttribute(attention_tile_kernel<D, SPLIT, POLY, MASKED, WIDE, TRACE>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  dim3 grid((p.Nq + 127) / 128, p.heads, p.B);
  attention_tile_kernel<D, SPLIT, POLY, MASKED, WIDE, TRACE><<<grid, 128 + 128 * SPLIT, S::kTotal, stream>>>(p);
  AF_LAUNCH_CHECK("attention_tile_kernel");
  return 0;
}

template <int D, int SPLIT, int POLY, bool WIDE>
static int launch_tile(const AttnTileParams& p, cudaStream_t stream) {
  constexpr int BN = TileCfg<D>::BLOCK_N;
  if (p.key_mask != nullptr || p.Nk % BN != 0) return launch_tile_m<D, SPLIT, POLY, true, WIDE>(p, stream);
  if constexpr (SPLIT == 1 && POLY == 0) {
    if (p.trace != nullptr) return launch_tile_m<D, SPLIT, POLY, false, WIDE, true>(p, stream);
  }
  return launch_tile_m<D, SPLIT, POLY, false, WIDE>(p, stream);
}

long long* attention_trace_ptr();  // attention_pair.cu
int attention_variant();

// Called by af_attention_bf16 (attention.cu) for d in {40, 80}, Nq >= 256, Nk > 128.  `variant`: bit 0 = two softmax
// threads per row, bits 1-2 = FMA-pipe share of the exponentials (0 none, 1 every 8th, 2 every 4th, 3 every 3rd),
// bit 3 = wide lazy-reference window (no subtraction in the loop).
int attention_tile_dispatch(const void* Q, long long ldq, const void* K, long long ldk, const void* Vt, long long ldvt,
                            int kv_stride, const unsigned char* key_mask, void* O, int B, int heads, int Nq, int Nk,
                            int d, float* lse, int variant, cudaStream_t stream) {
  if (!(d == 40 || d == 80)) return -100;
  AttnTileParams p;
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
    uint32_t box[2] = {64, static_cast<uint32_t>(d)};
    int rc = make_tmap_bf16(&p.tmV, Vt, 2, dims, str, box);
    if (rc) return rc;
  }
  p.B = B; p.heads = heads; p.Nq = Nq; p.Nk = Nk; p.d = d; p.dp = dp; p.kv_stride = kv_stride;
  p.key_mask = key_mask;
  p.lse = lse;
  p.trace = attention_trace_ptr();
  p.first_wave = 2 * num_sms();
  p.stagger = ((attention_variant() >> 8) & 0xff) * 64;
  p.out = static_cast<__nv_bfloat16*>(O);
  p.ldo = static_cast<long long>(heads) * d;
  const int split = (variant & 1) ? 2 : 1;
  const int poly = (variant >> 1) & 3;
  const bool wide = (variant & 8) != 0;
  if (d == 40) {
    if (split == 2) return wide ? launch_tile<40, 2, 0, true>(p, stream) : launch_tile<40, 2, 0, false>(p, stream);
    if (wide) {
      switch (poly) {
        case 0: return launch_tile<40, 1, 0, true>(p, stream);
        case 1: return launch_tile<40, 1, 8, true>(p, stream);
        case 2: return launch_tile<40, 1, 4, true>(p, stream);
        default: return launch_tile<40, 1, 3, true>(p, stream);
      }
    }
    switch (poly) {
      case 0: return launch_tile<40, 1, 0, false>(p, stream);
      case 1: return launch_tile<40, 1, 8, false>(p, stream);
      case 2: return launch_tile<40, 1, 4, false>(p, stream);
      default: return launch_tile<40, 1, 3, false>(p, stream);
    }
  }
  if (split == 2) return launch_tile<80, 2, 0, false>(p, stream);
  return wide ? launch_tile<80, 1, 0, true>(p, stream) : launch_tile<80, 1, 0, false>(p, stream);
}


// =====================================================================================================================
// Streamed schedule for d = 40 (attention_stream_kernel): 64-key blocks, THREE S buffers in tensor memory, P_j written
// over the first 32 columns of the S buffer it was computed from.
//
// What the device timelines of attention_tile_kernel showed (profiles/r02_attention.md): per 128-key block a softmax
// warp spent ~1550 cycles exponentiating and ~1500 cycles in its serial chain (S-ready wait, TMEM load, row maximum,
// P-free wait, store drain, hand-over), the co-resident CTA's warps drift through every relative phase, and so the ex2
// unit idles ~35 % of the time.  The chain is made of operations whose LATENCY is the cost, so this schedule removes
// them from the chain instead of overlapping CTAs:
//   * MMA issue order ... PV_j, S_{j+3} ... into three S buffers: "S_{j+1} ready" is always two blocks old when the
//     softmax thread asks for it, "S_{j+3} ready" implies PV_j complete (in-order commits), so there is NO s_free and
//     NO P-free barrier at all: one wait (s_full) and one arrive (p_full) per block;
//   * the TMEM load of S_{j+1} is issued before the exponentials of block j (two register tiles, alternating);
//   * the row maximum is off the critical path: with the wide lazy window the reference stays 0, the exponentials of a
//     block do not depend on its maximum, and the maximum is only a guard evaluated next to them (a trip - never seen
//     on UNet scores - re-does the block on the general path: blocking maximum, reference move, O correction).
// Same ones-row trick for the row sum and the same operand conventions as attention_tile_kernel.
// NSB = 3: the schedule above, two CTAs per SM.  NSB = 1: a single S buffer (no prefetch: S_{j+1} is issued right behind
// PV_j, the softmax thread pays the MMA round trip), but only 128 TMEM columns and 128 registers per softmax thread, so
// THREE CTAs share an SM and hide each other's chains.
template <int NSB>
struct StreamCfg {   // d = 40
  static constexpr int D = 40, DK = 48, DV = 48, BN = 64, NSBUF = NSB, KSTAGES = NSB == 1 ? 3 : 4, VSTAGES = NSB == 1 ? 2 : 3;
  static constexpr int kCtasPerSm = NSB == 1 ? 3 : 2;
  static constexpr uint32_t kTmemO = 64 * NSB, kTmemCols = NSB == 1 ? 128 : 256;   // S buffers at columns 0, 64, ..
  static constexpr int kQBytes = 128 * 128, kKBytes = BN * 128, kVBytes = DV * 128, kVTxBytes = D * 128;
  static constexpr int kQOff = 0, kKOff = kQOff + kQBytes, kVOff = kKOff + KSTAGES * kKBytes;
  static constexpr int kBarOff = kVOff + VSTAGES * kVBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

__device__ __forceinline__ void tile_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// PN of every PD exponentials go to the FMA pipe (0 / 1: none).
template <int PN, int PD>
__device__ __forceinline__ float stream_exp2(int i, float x) {
  if constexpr (PN == 9) return x * 0.001f;   // timing ablation only (scripts/attn_tile_check.py prints BAD parity)
  if constexpr (PN > 0) {
    if (i % PD < PN) return tile_exp2_poly(x);
  }
  return fast_exp2(x);
}

// PIPE (NSB = 3): the hand-over of P_{j-1} (store drain, fence, arrive) and the fetch of S_{j+1} (barrier test, TMEM
// load) are issued BETWEEN the two halves of block j's exponentials, i.e. behind 32 queued MUFU instructions per lane:
// the ex2 unit keeps draining its queue while this warp walks the latency-bound part of its chain.
template <int PN, int PD, bool MASKED, bool TRACE, int NSB, bool PIPE = false>
__global__ void __launch_bounds__(256, StreamCfg<NSB>::kCtasPerSm) attention_stream_kernel(const __grid_constant__ AttnTileParams p) {
  using C = StreamCfg<NSB>;
  constexpr int D = C::D, DK = C::DK, DV = C::DV, BN = C::BN;
  constexpr float kWindow = 64.0f;   // lazy-reference window, log2 domain (see attention_tile_kernel, WIDE)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = q_full + 1;               // KSTAGES
  uint64_t* k_empty = k_full + C::KSTAGES;
  uint64_t* v_full = k_empty + C::KSTAGES;     // VSTAGES
  uint64_t* v_empty = v_full + C::VSTAGES;
  uint64_t* s_full = v_empty + C::VSTAGES;     // NSBUF: S_j in buffer j % 3 (use j / 3 of that buffer)
  uint64_t* p_full = s_full + C::NSBUF;        // NSBUF: P_j written over the head of buffer j % 3
  // NSBUF: PV_j complete, on barrier j % 3 (waited on only by the rare O correction and the epilogue).  Per buffer and
  // not one barrier: a waiter may only test a phase parity if the barrier is at most one phase behind, and "S_j seen"
  // guarantees PV_{j-3} (in-order commits), i.e. exactly the previous phase of barrier (j-1) % 3 - nothing newer.
  uint64_t* pv_done = p_full + C::NSBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + C::NSBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_blocks = (p.Nk + BN - 1) / BN;
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  auto stamp = [&](int actor, int j, int ev) {   // same buffer layout as attention_tile_kernel
    if constexpr (TRACE) {
      if (lane == 0 && j < 64) {
        if (actor < 2) {
          if (cta_lin < 512) p.trace[(cta_lin * 64 + j) * 8 + ev] = clock64();
        } else if (cta_lin == 0) {
          p.trace[512 * 64 * 8 + 512 + ((actor - 2) * 64 + j) * 8 + ev] = clock64();
        }
      }
    }
  };
  if constexpr (TRACE) {
    if (threadIdx.x == 0 && cta_lin < 512) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[512 * 64 * 8 + cta_lin] = smid;
    }
  }

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
    for (int s = 0; s < C::NSBUF; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    mbar_fence_init();
  }
  {   // constant rows of every V^T stage: row D = 1.0, rows D+1 .. DV-1 = 0 (never touched by the {64, D} TMA box)
    constexpr int kConstChunks = (DV - D) * 8;
    for (int i = threadIdx.x; i < C::VSTAGES * kConstChunks; i += 256) {
      const int st = i / kConstChunks, chunk = i % kConstChunks;
      const uint32_t v = chunk < 8 ? 0x3F803F80u : 0u;
      sts128(smem_base + C::kVOff + st * C::kVBytes + D * 128 + chunk * 16, v, v, v, v);
    }
    fence_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer: K runs three blocks ahead of V
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, C::kQBytes);
        tma_load_3d(smem + C::kQOff, &p.tmQ, q_full, h * p.dp, q0, b);
        int ks = 0, vs = 0;
        uint32_t kph = 0, vph = 0;
        auto load_k = [&](int j) {
          mbar_wait_lean(&k_empty[ks], kph ^ 1);
          mbar_arrive_expect_tx(&k_full[ks], C::kKBytes);
          tma_load_3d(smem + C::kKOff + ks * C::kKBytes, &p.tmK, &k_full[ks], h * p.dp, j * BN, b);
          if (++ks == C::KSTAGES) { ks = 0; kph ^= 1; }
        };
        for (int j = 0; j < C::NSBUF && j < n_blocks; ++j) load_k(j);
        for (int j = 0; j < n_blocks; ++j) {
          mbar_wait_lean(&v_empty[vs], vph ^ 1);
          mbar_arrive_expect_tx(&v_full[vs], C::kVTxBytes);
          tma_load_2d(smem + C::kVOff + vs * C::kVBytes, &p.tmV, &v_full[vs], b * p.kv_stride + j * BN, h * p.d);
          if (++vs == C::VSTAGES) { vs = 0; vph ^= 1; }
          if (j + C::NSBUF < n_blocks) load_k(j + C::NSBUF);
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer: S_0 S_1 S_2, then PV_j, S_{j+3}
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DV);
      const uint32_t q_addr = smem_base + C::kQOff;
      const uint32_t tm_o = tmem_base + C::kTmemO;
      int kslot = 0, vslot = 0;
      uint32_t kph = 0, vph = 0;
      auto issue_s = [&](int sbuf) {   // next K block of the ring -> S buffer sbuf
        mbar_wait_lean(&k_full[kslot], kph);
        tc_fence_after();
        const uint32_t k_addr = smem_base + C::kKOff + kslot * C::kKBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DK / 16; ++k)
            tc_mma_ss(tmem_base + sbuf * 64, umma_desc_sw128(q_addr) + 2 * k, umma_desc_sw128(k_addr) + 2 * k, idesc_s,
                      k != 0 ? 1u : 0u);
          tc_commit(&s_full[sbuf]);
          tc_commit(&k_empty[kslot]);
        }
        __syncwarp();
        if (++kslot == C::KSTAGES) { kslot = 0; kph ^= 1; }
      };
      mbar_wait_lean(q_full, 0);
      for (int j = 0; j < C::NSBUF && j < n_blocks; ++j) issue_s(j);
      int sbuf = 0;
      uint32_t sph = 0;
      for (int j = 0; j < n_blocks; ++j) {
        stamp(2, j, 0);
        mbar_wait_lean(&v_full[vslot], vph);
        stamp(2, j, 1);
        mbar_wait_lean(&p_full[sbuf], sph);
        tc_fence_after();
        stamp(2, j, 2);
        const uint32_t v_addr = smem_base + C::kVOff + vslot * C::kVBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BN / 16; ++k)
            tc_mma_ts(tm_o, tmem_base + sbuf * 64 + k * 8, umma_desc_sw128(v_addr) + 2 * k, idesc_o,
                      (j | k) != 0 ? 1u : 0u);
          tc_commit(&pv_done[sbuf]);
          tc_commit(&v_empty[vslot]);
        }
        __syncwarp();
        stamp(2, j, 3);
        if (++vslot == C::VSTAGES) { vslot = 0; vph ^= 1; }
        if (j + C::NSBUF < n_blocks) issue_s(sbuf);
        stamp(2, j, 4);
        if (++sbuf == C::NSBUF) { sbuf = 0; sph ^= 1; }
      }
    }
  } else {
    if constexpr (NSB == 1) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    }
    // ------------------------------------------------------------------ softmax: one thread per query row
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int q_row = q0 + r;
    const uint8_t* mrow = p.key_mask ? p.key_mask + static_cast<size_t>(b) * p.Nk : nullptr;
    const uint32_t s_addr = tmem_base + lane_off;
    const uint32_t o_addr = tmem_base + C::kTmemO + lane_off;
    const uint32_t a_s_full = smem_u32(s_full), a_p_full = smem_u32(p_full), a_pv_done = smem_u32(pv_done);
    const bool tw = TRACE && warp == 4;
    float m_ref = -INFINITY;
    int sbuf = 0;         // buffer of the block being exponentiated
    uint32_t sph = 0;     // its use parity

    // S_{j} -> registers (asynchronous: complete after the next tmem_ld_wait)
    auto load_s = [&](float* dst, int buf, uint32_t parity) {
      mbar_wait_lean(a_s_full + 8 * buf, parity);
      tc_fence_after();
      tile_ld32(s_addr + buf * 64, reinterpret_cast<uint32_t*>(dst));
      tile_ld32(s_addr + buf * 64 + 32, reinterpret_cast<uint32_t*>(dst) + 32);
    };
    auto block_max = [&](const float* sc) {
      float mx8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(sc[2 * c], sc[2 * c + 1]);
#pragma unroll
      for (int e = 16; e < BN; e += 16)
#pragma unroll
        for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(mx8[c], fmaxf(sc[e + 2 * c], sc[e + 2 * c + 1]));
      return fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                   fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
    };
    // one key block: `cur` holds S_j (landed), S_{j+1} is fetched into `nxt` underneath the exponentials
    auto step = [&](float* cur, float* nxt, int j) {
      if (tw) stamp(0, j, 0);
      if constexpr (NSB == 1) {
        load_s(cur, 0, sph);
        tmem_ld_wait();
      } else if (j + 1 < n_blocks) {
        const int nb = sbuf + 1 == C::NSBUF ? 0 : sbuf + 1;
        load_s(nxt, nb, nb == 0 ? sph ^ 1 : sph);
      }
      if (tw) stamp(0, j, 1);
      if constexpr (MASKED) {
        const int key0 = j * BN;
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          const int key = key0 + c + lane;
          bool ok = key < p.Nk;
          if (ok && mrow != nullptr) ok = __ldg(mrow + key) != 0;
          const uint32_t keep = __ballot_sync(0xffffffffu, ok);
#pragma unroll
          for (int e = 0; e < 32; ++e) cur[c + e] = ((keep >> e) & 1u) ? cur[c + e] : -INFINITY;
        }
      }
      uint32_t pk[32];
      bool fast = __all_sync(0xffffffffu, m_ref == 0.f);
      if (fast) {
#pragma unroll
        for (int e = 0; e < BN; e += 2)
          pk[e >> 1] = pack_bf16x2(stream_exp2<PN, PD>(e, cur[e]), stream_exp2<PN, PD>(e + 1, cur[e + 1]));
        const float mx = block_max(cur);             // guard only: evaluated next to the exponentials
        fast = !__any_sync(0xffffffffu, mx > kWindow);
      }
      if (!fast) {
        // general path (first block, moved reference, or guard trip): blocking maximum, lazy reference move, O correction
        const float mx = block_max(cur);
        float alpha = 1.0f;
        if (mx > m_ref + kWindow) {   // also the first finite maximum (m_ref == -inf)
          const float m_new = (m_ref == -INFINITY && fabsf(mx) <= kWindow) ? 0.f : mx;
          alpha = fast_exp2(m_ref - m_new);
          m_ref = m_new;
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
          {   // every MMA of PV_{j-1} has landed
            const int pb = sbuf == 0 ? C::NSBUF - 1 : sbuf - 1;
            mbar_wait_lean(a_pv_done + 8 * pb, sbuf == 0 ? sph ^ 1 : sph);
          }
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DV; c += 16) {
            uint32_t o[16];
            tmem_ld16(o_addr + c, o);
            tmem_ld_wait();                           // (also drains the S_{j+1} prefetch: harmless)
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tile_st16(o_addr + c, o);
          }
          tmem_st_wait();
        }
#pragma unroll
        for (int e = 0; e < BN; e += 2)
          pk[e >> 1] = pack_bf16x2(fast_exp2(cur[e] - m_use), fast_exp2(cur[e + 1] - m_use));
      }
      if (tw) stamp(0, j, 2);
      tile_st32(s_addr + sbuf * 64, pk);               // P_j over the head of its own S buffer
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_p_full + 8 * sbuf) : "memory");
      if (tw) stamp(0, j, 3);
      if constexpr (NSB > 1) tmem_ld_wait();           // S_{j+1} is in `nxt`
      if (++sbuf == C::NSBUF) { sbuf = 0; sph ^= 1; }
      if (tw) stamp(0, j, 4);
    };

    // general path of one block (first block, moved reference, or guard trip): blocking maximum, lazy reference move,
    // O correction, exponentials with the subtraction
    auto general_block = [&](const float* cur, uint32_t* pk, int j) {
      const float mx = block_max(cur);
      float alpha = 1.0f;
      if (mx > m_ref + kWindow) {
        const float m_new = (m_ref == -INFINITY && fabsf(mx) <= kWindow) ? 0.f : mx;
        alpha = fast_exp2(m_ref - m_new);
        m_ref = m_new;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        const int pb = sbuf == 0 ? C::NSBUF - 1 : sbuf - 1;
        mbar_wait_lean(a_pv_done + 8 * pb, sbuf == 0 ? sph ^ 1 : sph);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < DV; c += 16) {
          uint32_t o[16];
          tmem_ld16(o_addr + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tile_st16(o_addr + c, o);
        }
        tmem_st_wait();
      }
#pragma unroll
      for (int e = 0; e < BN; e += 2)
        pk[e >> 1] = pack_bf16x2(fast_exp2(cur[e] - m_use), fast_exp2(cur[e + 1] - m_use));
    };
    auto step_pipe = [&](float* cur, float* nxt, int j) {
      if (tw) stamp(0, j, 0);
      if constexpr (MASKED) {
        const int key0 = j * BN;
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          const int key = key0 + c + lane;
          bool ok = key < p.Nk;
          if (ok && mrow != nullptr) ok = __ldg(mrow + key) != 0;
          const uint32_t keep = __ballot_sync(0xffffffffu, ok);
#pragma unroll
          for (int e = 0; e < 32; ++e) cur[c + e] = ((keep >> e) & 1u) ? cur[c + e] : -INFINITY;
        }
      }
      uint32_t pk[32];
      bool fast = __all_sync(0xffffffffu, m_ref == 0.f);
      if (fast) {
#pragma unroll
        for (int e = 0; e < BN / 2; e += 2)
          pk[e >> 1] = pack_bf16x2(stream_exp2<PN, PD>(e, cur[e]), stream_exp2<PN, PD>(e + 1, cur[e + 1]));
      }
      if (tw) stamp(0, j, 1);
      if (j > 0) {   // hand P_{j-1} over (its store was issued at the end of the previous step)
        const int pb = sbuf == 0 ? C::NSBUF - 1 : sbuf - 1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_p_full + 8 * pb) : "memory");
      }
      if (j + 1 < n_blocks) {
        const int nb = sbuf + 1 == C::NSBUF ? 0 : sbuf + 1;
        load_s(nxt, nb, nb == 0 ? sph ^ 1 : sph);
      }
      if (tw) stamp(0, j, 2);
      if (fast) {
#pragma unroll
        for (int e = BN / 2; e < BN; e += 2)
          pk[e >> 1] = pack_bf16x2(stream_exp2<PN, PD>(e, cur[e]), stream_exp2<PN, PD>(e + 1, cur[e + 1]));
        const float mx = block_max(cur);
        fast = !__any_sync(0xffffffffu, mx > kWindow);
      }
      if (!fast) general_block(cur, pk, j);
      if (tw) stamp(0, j, 3);
      tile_st32(s_addr + sbuf * 64, pk);
      tmem_ld_wait();                                  // S_{j+1} is in `nxt`
      if (++sbuf == C::NSBUF) { sbuf = 0; sph ^= 1; }
      if (tw) stamp(0, j, 4);
    };

    if constexpr (PIPE) {
      static_assert(NSB == 3, "the pipelined step needs the S prefetch");
      float sa[BN], sb[BN];
      load_s(sa, 0, 0);
      tmem_ld_wait();
      for (int j = 0; j < n_blocks; j += 2) {
        step_pipe(sa, sb, j);
        if (j + 1 < n_blocks) step_pipe(sb, sa, j + 1);
      }
      {   // hand the last P over
        const int pb = sbuf == 0 ? C::NSBUF - 1 : sbuf - 1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_p_full + 8 * pb) : "memory");
      }
    } else if constexpr (NSB == 1) {
      float sa[BN];
      for (int j = 0; j < n_blocks; ++j) step(sa, sa, j);
    } else {
      float sa[BN], sb[BN];
      load_s(sa, 0, 0);
      tmem_ld_wait();
      for (int j = 0; j < n_blocks; j += 2) {
        step(sa, sb, j);
        if (j + 1 < n_blocks) step(sb, sa, j + 1);
      }
    }
    // epilogue: l = column D of O; O / l -> bf16
    mbar_wait_lean(a_pv_done + 8 * ((n_blocks - 1) % C::NSBUF), ((n_blocks - 1) / C::NSBUF) & 1);
    tc_fence_after();
    float l_tot;
    {
      uint32_t lv[8];
      tmem_ld8(o_addr + D, lv);
      tmem_ld_wait();
      l_tot = __uint_as_float(lv[0]);
    }
    const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
    if (p.lse != nullptr && q_row < p.Nq)
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Nq + q_row] = l_tot > 0.f ? m_ref + __log2f(l_tot) : INFINITY;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.Nq + q_row) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c = 0; c < D; c += 16) {
      uint32_t o[16];
      tmem_ld16(o_addr + c, o);
      tmem_ld_wait();
      if (q_row < p.Nq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < D) {
            uint4 pk4;
            pk4.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
            pk4.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
            pk4.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
            pk4.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + g * 8) = pk4;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int PN, int PD, bool MASKED, int NSB, bool PIPE, bool TRACE = false>
static int launch_stream_m(const AttnTileParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(attention_stream_kernel<PN, PD, MASKED, TRACE, NSB, PIPE>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<NSB>::kTotal));
    AF_CUDA(cudaFuncSetAttribute(attention_stream_kernel<PN, PD, MASKED, TRACE, NSB, PIPE>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  dim3 grid((p.Nq + 127) / 128, p.heads, p.B);
  attention_stream_kernel<PN, PD, MASKED, TRACE, NSB, PIPE><<<grid, 256, StreamCfg<NSB>::kTotal, stream>>>(p);
  AF_LAUNCH_CHECK("attention_stream_kernel");
  return 0;
}
template <int PN, int PD, int NSB, bool PIPE = false>
static int launch_stream(const AttnTileParams& p, cudaStream_t stream) {
  if (p.key_mask != nullptr || p.Nk % 64 != 0) return launch_stream_m<PN, PD, true, NSB, PIPE>(p, stream);
  if constexpr (PN == 0) {
    if (p.trace != nullptr) return launch_stream_m<PN, PD, false, NSB, PIPE, true>(p, stream);
  }
  return launch_stream_m<PN, PD, false, NSB, PIPE>(p, stream);
}

// d = 40 only.  `poly`: 0 none, 1 = 1/4, 2 = 1/3, 3 = 2/5, 4 = 1/2 of the exponentials on the FMA pipe.
int attention_stream_dispatch(const void* Q, long long ldq, const void* K, long long ldk, const void* Vt, long long ldvt,
                              int kv_stride, const unsigned char* key_mask, void* O, int B, int heads, int Nq, int Nk,
                              float* lse, int poly, cudaStream_t stream) {
  AttnTileParams p;
  memset(&p, 0, sizeof(p));
  const int d = 40, dp = 48;
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
    uint32_t box[3] = {64, 64, 1};
    int rc = make_tmap_bf16(&p.tmK, K, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(ldvt), static_cast<uint64_t>(heads) * d};
    uint64_t str[1] = {static_cast<uint64_t>(ldvt) * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(d)};
    int rc = make_tmap_bf16(&p.tmV, Vt, 2, dims, str, box);
    if (rc) return rc;
  }
  p.B = B; p.heads = heads; p.Nq = Nq; p.Nk = Nk; p.d = d; p.dp = dp; p.kv_stride = kv_stride;
  p.key_mask = key_mask;
  p.lse = lse;
  p.trace = attention_trace_ptr();
  p.out = static_cast<__nv_bfloat16*>(O);
  p.ldo = static_cast<long long>(heads) * d;
  switch (poly) {   // bit 2: single S buffer, three CTAs per SM
    case 0: return launch_stream<0, 1, 3>(p, stream);
    case 1: return launch_stream<1, 4, 3>(p, stream);
    case 2: return launch_stream<1, 3, 3>(p, stream);
    case 3: return launch_stream<2, 5, 3>(p, stream);
    case 4: return launch_stream<0, 1, 1>(p, stream);
    case 5: return launch_stream<1, 4, 1>(p, stream);
    case 6: return launch_stream<9, 1, 1>(p, stream);          // ablation: no exponentials
    case 7: return launch_stream<9, 1, 3>(p, stream);          // ablation: no exponentials
    case 8: return launch_stream<0, 1, 3, true>(p, stream);   // bit 3: intra-warp pipelined step
    case 9: return launch_stream<1, 4, 3, true>(p, stream);
    case 10: return launch_stream<1, 3, 3, true>(p, stream);
    default: return launch_stream<2, 5, 3, true>(p, stream);
  }
}

}  // namespace af
