// BLSTM recurrence with the recurrent weights resident in TENSOR MEMORY.
//
// The time loop of torch.nn.LSTM inside RNNP_packed (tssep/train/rnnp.py:87-95, :143-159): gate order
// i,f,g,o, zero initial state, c_t = s(f) c_{t-1} + s(i) tanh(g), h_t = s(o) tanh(c_t), both directions.
//
// One cluster per (NR batch rows, direction), NR = 8, 16 or 32.  Each CTA owns TILES row tiles of 32 hidden units =
// 128 gate rows (row = 4*unit + gate), each an M=128 A tile of Up/2 TMEM columns (two bf16 per 32-bit column),
// written ONCE.  TILES = 2: clusters of ceil(Up/64) CTAs (5 at U = 300), the throughput shape; TILES = 1: clusters
// of 2*ceil(Up/64) CTAs (10 at U = 300, a non-portable cluster size), the latency shape -- one tile per CTA halves
// the MMAs (13-17 cycles each to issue) and the gate math on the dependent chain of a step.
// Per step the pre-activations of a tile are
//
//     D[gate row, batch row] = P . G_t  +  W_hh . h_{t-1}
//
// entirely on the tensor core (tcgen05.mma, A from tensor memory, B from shared memory, fp32 accumulation in
// TMEM).  P is a 128x128 scaled permutation matrix in TMEM that routes the step's input projections G_t -- a TMA
// box of the plain (row, t, dir, gate, unit) bf16 tensor, which lands in shared memory in exactly the K-major
// 128-byte-swizzled layout a B operand needs -- to their gate rows, so the epilogue never touches G (the G term was
// a third of the epilogue's instructions when it was added from shared memory).  The rows of gates i, f, o of both
// P and W_hh carry a factor 1/2 (exact in bf16), so every gate is  k_a * tanh(D) + k_b.
//
//   warp 0 (one thread)  TMA producer: one 5-D box of G per step into an mbarrier ring, several steps ahead;
//   warp 1               MMA issuer (convergent, elect.sync): the P.G MMAs as soon as the accumulators were read
//                        (long before h arrives), then W_hh.h in two K phases: the k-steps fed by the FIRST row
//                        tile of every CTA start while the second tile's epilogue is still running;
//   warps 2..            epilogue (lane = gate row; 8 * NR/NC warps, each NC batch columns of one row tile and TMEM
//                        lane quarter): tcgen05.ld, gates, 4x4 quad transposes, c_t in registers, h_t regrouped to
//                        16-byte chunks and pushed to every CTA's next B operand with st.async (complete_tx on the
//                        destination mbarrier of that K phase) and to H.
#include "../../include/tssep_b200.h"
#include "common.cuh"

#include <type_traits>

namespace tssep {

constexpr int kTsMaxStages = 8;

struct RecTsArgs {
  const uint4* Wimg;  // [dir][cta][tile][kstep][row 128][8 words], gate rows i,f,o pre-scaled by 1/2
  __nv_bfloat16* H;   // (rows, T, 2*Up)
  int rows, T, Up, NA, KS, stages;
  uint2* gates_out;   // SAVE variant (training): (rows, T, 2, Up) x {i, f, g, o} bf16 gate activations, else null
  float* c_out;       // SAVE variant: (rows, T, 2, Up) f32 cell states
  int flags;          // debug builds only: bit 0 = skip the proxy fence of the MMA warp (measurement, NOT correct)
  int* prof;          // debug builds only (TSSEP_DEBUG_KNOBS): per-phase cycle counters of two epilogue warps + MMA warp
};

__device__ __forceinline__ uint64_t ts_desc_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void ts_st_async_v4(uint32_t remote_addr, const uint4& v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   remote_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void ts_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with the shared-memory descriptor as two words (only the low one -- start address -- differs between the
// operands of a kernel): the caller's address arithmetic stays 32 bits wide
__device__ __forceinline__ void tc_mma_bf16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x),
               "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// tanh of the (pre-scaled) pre-activation: MATH 1 = tanh.approx.f32, MATH 0 = exp based
template <int MATH>
__device__ __forceinline__ float ts_tanh(float x) {
  if constexpr (MATH == 1) return tanh_fast(x);
  else return tanh_acc(x);
}

// NR batch rows per sub-batch, NC of them per epilogue warp; MATH: gate arithmetic; SPLIT: two K phases per step (see
// the header); TILES: row tiles per CTA; SUBS: independent sub-batches of NR rows per cluster.  SUBS = 2 is the
// throughput shape: the cluster advances two recurrences in anti-phase -- while one sub-batch's h travels through
// DSMEM and its epilogue warps run, the MMA warp and the tensor pipe work on the other.
// Everything the per-step loops branch on is a template parameter.
// GEPI (the default for two sub-batches of 8 rows): the epilogue adds G_t from shared memory instead of the P . G MMAs;
// the loads are issued while the warp waits for the accumulator anyway.  Measured after the MMA issue was unrolled
// (profiles/r2_rec_ts_microbench.txt): 2 x 8 rows 0.96-1.00 us per step with GEPI against 1.14 without; 2 x 16 rows
// 1.79-1.81 with against 1.48-1.54 without (16 conflicting 2-byte loads per lane cost the epilogue -- which bounds that
// shape -- more than 16 MMAs cost the tensor pipe); 2 x 32 rows spill with it.
// Also measured and dropped: a helper warp that takes the proxy fence of the h exchange (250 cycles under the st.async
// traffic of a ping-pong step) off the MMA warp -- the MMA warp then waits that much longer for h: at 32 rows per cluster
// the step is bound by the DSMEM egress of the exchange (16 KB per CTA and step at ~20 B/clk) plus its latency.
template <int NR, int NC, int MATH, bool SPLIT, int TILES, int SUBS, bool SAVE = false, bool GEPI = (SUBS == 2 && NR == 8)>
__global__ void __launch_bounds__(64 + 128 * TILES * SUBS * (NR / NC), 1)
blstm_rec_ts_kernel(const RecTsArgs a, const __grid_constant__ CUtensorMap gmap) {
  static_assert(TILES == 1 || TILES == 2, "one or two row tiles per CTA");
  static_assert(SUBS == 1 || SUBS == 2, "one or two sub-batches per cluster");
  static_assert((TILES == 2 && SUBS == 1) || !SPLIT, "the K split pairs the two row tiles of a CTA");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NB = NR < 16 ? 16 : NR;   // MMA N (M = 128 needs N % 16 == 0); rows NR..NB-1 of the operands stay zero
  constexpr int NRT = NR * SUBS;          // batch rows per cluster
  constexpr uint32_t kAtomB = NB * 128;   // one 64-k atom of the h operand: NB rows x 128 bytes, 128-byte swizzle
  constexpr uint32_t kAtomG = NRT * 128;  // one 64-k atom of the G operand as the TMA box lays it out (all sub-batches)
  constexpr int EW = NR / NC;             // epilogue warps per (sub-batch, row tile, TMEM lane quarter)
  constexpr int NQ = NC / 4;              // batch-row quads per epilogue warp
  constexpr int kThreads = 64 + 128 * TILES * SUBS * EW;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NA = a.NA, KS = a.KS, GS = a.stages;
  constexpr uint32_t g_stage = 4u * kAtomG;                 // [gate][row][64 units]
  const uint32_t sB = base;                                 // [sub-batch][2 buffers][NA] x kAtomB
  const uint32_t sub_bytes = 2u * NA * kAtomB;
  const uint32_t sG = sB + SUBS * sub_bytes;                // [GS] x g_stage (+ 1 KiB of zeros behind the last stage)
  const uint32_t sT = sG + GS * g_stage + 1024u;            // [epilogue warps] x NC x 16 B
  const uint32_t sBar = sT + 4u * TILES * NRT * 16u;
  const uint32_t hfull0 = sBar;                             // [sub-batch][buffer][phase]
  const uint32_t accfull0 = sBar + 64, accempty0 = sBar + 96;  // [sub-batch][tile]
  const uint32_t gfull0 = sBar + 128, gempty0 = gfull0 + 8 * kTsMaxStages, tptr = gempty0 + 8 * kTsMaxStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t C = cluster_nctarank();
  const int dir = blockIdx.z;
  const int row0 = blockIdx.y * NRT;  // first batch row of this cluster
  const int T = a.T, Up = a.Up;
  // every CTA ships TILES x 32 units x NR rows per step and sub-batch; with SPLIT each row tile completes its own barrier
  const uint32_t tx_bytes = static_cast<uint32_t>(NR) * 64u * (SPLIT ? 1u : static_cast<uint32_t>(TILES)) * C;

  // ---- one-time setup ---------------------------------------------------------------------------
  for (uint32_t i = threadIdx.x; i < (sT - sB) / 16; i += kThreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sB + 16 * i), "r"(0u) : "memory");
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 8; ++i) mbar_init(hfull0 + 8 * i, 1);
      for (int i = 0; i < 4; ++i) {
        mbar_init(accfull0 + 8 * i, 1);
        mbar_init(accempty0 + 8 * i, 4 * EW);
      }
      for (int i = 0; i < GS; ++i) {
        mbar_init(gfull0 + 8 * i, 1);
        mbar_init(gempty0 + 8 * i, GEPI ? 4 * TILES * SUBS * EW : 1);  // GEPI: every epilogue warp reads the stage
      }
      mbar_fence_init();
      for (int sb = 0; sb < SUBS; ++sb)
        for (int bf = 0; bf < 2; ++bf)
          for (int ph = 0; ph < (SPLIT ? 2 : 1); ++ph) mbar_arrive_expect_tx(hfull0 + 8 * ((sb * 2 + bf) * 2 + ph), tx_bytes);
    }
    __syncwarp();
    tc_alloc(tptr, 512);
    tc_relinquish();
  }
  ts_fence_proxy_async();  // zeros written through the generic proxy -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const uint32_t a_tile_cols = (static_cast<uint32_t>(Up) / 2 + 31u) & ~31u;  // columns per A tile (32-aligned)
  const uint32_t p_col = TILES * a_tile_cols;                                  // P behind the A tiles (64 columns)
  const uint32_t acc_col = p_col + 64u;                                        // then the accumulators [sub][tile]

  if (warp >= 2 && warp < 2 + 4 * TILES) {
    // W_hh -> TMEM: lane = gate row of the tile, 8 columns (16 k values) per store
    const int tl = (warp - 2) >> 2, q = warp & 3;
    const uint4* src = a.Wimg + ((static_cast<size_t>(dir) * C + crank) * TILES + tl) * static_cast<size_t>(KS) * 256 +
                       static_cast<size_t>(q * 32 + lane) * 2;
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(tl) * a_tile_cols;
    for (int k = 0; k < KS; ++k) {
      const uint4 w0 = __ldg(src + static_cast<size_t>(k) * 256), w1 = __ldg(src + static_cast<size_t>(k) * 256 + 1);
      tc_st8(t0 + k * 8, w0, w1);
    }
    if (tl == 0 && !GEPI) {
      // P: row m = 4*unit + gate of a tile picks k = gate*32 + unit of the tile's G operand (two k-steps of 16
      // units from each of the four gate atoms), scaled by 1/2 for the sigmoid gates
      const int m = q * 32 + lane, gate = m & 3, unit = m >> 2;
      const int kk = gate * 32 + unit;
      const uint32_t one = (gate == 2 ? 0x3F80u : 0x3F00u) << (16 * (kk & 1));
      const uint32_t tp = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + p_col;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int w = (kk >> 1) - 8 * k;  // word of this k-step holding the non-zero, if any
        uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
        if (w == 0) lo.x = one;
        if (w == 1) lo.y = one;
        if (w == 2) lo.z = one;
        if (w == 3) lo.w = one;
        if (w == 4) hi.x = one;
        if (w == 5) hi.y = one;
        if (w == 6) hi.z = one;
        if (w == 7) hi.w = one;
        tc_st8(tp + k * 8, lo, hi);
      }
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();

  if (warp == 0) {
    // ---- G producer ---------------------------------------------------------------------------------
    // G (rows, T, 2, 4, Up) bf16 viewed as (unit, row, gate, dir, t): one box of 64 units x NRT rows x 4 gates per
    // step = [gate][row][128 bytes], per gate the canonical K-major SWIZZLE_128B operand layout (K = 64 units:
    // k-steps 0,1 belong to the even row tile, k-steps 2,3 to the odd one; with one tile per CTA the two CTAs of a
    // pair load the same box and each uses its half; the sub-batches are row ranges of the box).
    if (lane == 0) {
      tma_prefetch_desc(&gmap);
      int slot = 0;
      uint32_t gph = 0;
      for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        mbar_wait(gempty0 + 8 * slot, gph ^ 1);
        mbar_arrive_expect_tx(gfull0 + 8 * slot, g_stage);
        tma_load_5d(sG + slot * g_stage, &gmap, gfull0 + 8 * slot, static_cast<int>(crank * TILES / 2) * 64, row0, 0, dir, t);
        if (++slot == GS) {
          slot = 0;
          gph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------
    // The whole warp runs the loop convergently and elect.sync guards only the MMA blocks: inside a
    // divergent `if (lane == 0)` ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY waterfall and
    // rebuilds the descriptor through a long uniform-datapath chain (~70 cycles per MMA, measured).
    // The tensor pipe itself needs only max(8, N/2) cycles per M=128, K=16 MMA with A in tensor memory
    // (scripts/probes/mma_probe.cu: 9 cycles at N=16, 16 at N=32) -- what a step pays per MMA is the ISSUE
    // sequence.  So for the product size (KSC = 19 k-steps) the k loop is unrolled completely, every operand address
    // is "base + compile-time constant", and the 64-bit shared-memory descriptor is kept as two words of which only
    // the low one (start address, 16-byte units) ever changes: 2 adds + 2 R2UR per MMA instead of a 64-bit add
    // chain, per-atom loop control and predicates (27-30 cycles per MMA before).  Other sizes take the loop form.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NB >> 3) << 17) |
                           (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t bdesc0 = ts_desc_sw128(sB);
    const uint64_t gdesc0 = ts_desc_sw128(sG);
    const uint32_t desc_hi = static_cast<uint32_t>(bdesc0 >> 32);  // SBO, version, swizzle: the same for every operand
    const uint32_t bdesc_lo0 = static_cast<uint32_t>(bdesc0), gdesc_lo0 = static_cast<uint32_t>(gdesc0);
    const uint32_t buf_step = static_cast<uint32_t>(NA) * (kAtomB >> 4);  // encoded distance of two h buffers
    const uint32_t d0 = tmem_base + acc_col;
    const uint32_t pt = tmem_base + p_col;
#ifdef TSSEP_DEBUG_KNOBS
    const bool mprof = a.prof != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && crank == 0;
    const bool no_fence = (a.flags & 1) != 0;
#else
    constexpr bool mprof = false, no_fence = false;
#endif
    // the whole time loop, instantiated for a compile-time k-step count (ksc > 0) or the run-time one (ksc = 0)
    auto run = [&](auto ksc) {
      constexpr int KSC = decltype(ksc)::value;
      // W_hh . h for one tile and one K phase: phase 0 = k-steps 0,1 of every atom (units of the senders' first
      // row tile), phase 1 = k-steps 2,3; without SPLIT phase 0 covers all four.  Tile after tile, not alternating
      // (measured: alternating makes both tiles' epilogues start together instead of the first one overlapping the
      // second tile's MMAs).  k-step k: A columns 8k of the tile, B = atom k/4, 32-byte slice k%4.
      auto issue_h = [&](uint32_t d, int tile, int phase, uint32_t bd_lo) {
        const uint32_t at = tmem_base + static_cast<uint32_t>(tile) * a_tile_cols;
        const int k_lo = phase * 2, k_hi = SPLIT ? k_lo + 2 : 4;
        if constexpr (KSC > 0) {
#pragma unroll
          for (int k = 0; k < KSC; ++k)
            if ((k & 3) >= k_lo && (k & 3) < k_hi)
              tc_mma_bf16_ts2(d, at + 8 * k, bd_lo + (k >> 2) * (kAtomB >> 4) + 2 * (k & 3), desc_hi, idesc,
                              (GEPI && k == 0) ? 0u : 1u);
        } else {
#pragma unroll 4
          for (int k = 0; k < KS; ++k)
            if ((k & 3) >= k_lo && (k & 3) < k_hi)
              tc_mma_bf16_ts2(d, at + 8 * k, bd_lo + (k >> 2) * (kAtomB >> 4) + 2 * (k & 3), desc_hi, idesc,
                              (GEPI && k == 0) ? 0u : 1u);
        }
      };
      int mc[4] = {0, 0, 0, 0};
      int slot = 0;
      uint32_t gph = 0;
      for (int s = 0; s < T; ++s) {
        const int rb = (s & 1) ^ 1;
        if constexpr (!GEPI) mbar_wait(gfull0 + 8 * slot, gph);
        const uint32_t gd_lo = gdesc_lo0 + static_cast<uint32_t>(slot) * (g_stage >> 4);
#pragma unroll
        for (int sb = 0; sb < SUBS; ++sb) {
          int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
          if (mprof) m0 = clock();
          const uint32_t dsub = d0 + static_cast<uint32_t>(sb * TILES) * NB;
          // P . G_s: needs the accumulators of step s-1 drained -- long before h arrives
          if (s > 0) {
            mbar_wait(accempty0 + 8 * (sb * 2), (s - 1) & 1);
            if constexpr (TILES == 2) mbar_wait(accempty0 + 8 * (sb * 2 + 1), (s - 1) & 1);
          }
          tc_fence_after();
          if (!GEPI && elect_one()) {
#pragma unroll
            for (int tile = 0; tile < TILES; ++tile) {
              // the tile's 32 units are one half (hx) of the 64-unit box: 64 bytes = 4 descriptor units into the row
              const uint32_t gt_lo = gd_lo + static_cast<uint32_t>(sb) * (NR * 8) + 4u * ((crank * TILES + tile) & 1u);
#pragma unroll
              for (int k = 0; k < 8; ++k)  // k-step k of P = gate k/2, half k%2 of the tile's 32 units
                tc_mma_bf16_ts2(dsub + tile * NB, pt + k * 8, gt_lo + (k >> 1) * (kAtomG >> 4) + 2 * (k & 1), desc_hi, idesc,
                                k > 0 ? 1u : 0u);
            }
          }
          __syncwarp();
          if (mprof) m1 = clock();
          const uint32_t bd_lo = bdesc_lo0 + static_cast<uint32_t>(sb * 2 + rb) * buf_step;
          const uint32_t hbar = hfull0 + 8 * ((sb * 2 + rb) * 2);
          if (s > 0) {
            const uint32_t par = ((s - 1) >> 1) & 1;
            mbar_wait(hbar, par);
            if (mprof) m2 = clock();
            if (lane == 0) mbar_arrive_expect_tx(hbar, tx_bytes);  // re-arm for the data of step s+1
            // h arrived through st.async (generic proxy); the MMA reads it through the async proxy
            if (!no_fence) ts_fence_proxy_async();
            tc_fence_after();
            if (mprof) m3 = clock();
            if constexpr (SPLIT) {
              if (elect_one()) {
                issue_h(dsub, 0, 0, bd_lo);
                issue_h(dsub + NB, 1, 0, bd_lo);
              }
              __syncwarp();
              mbar_wait(hbar + 8, par);
              if (lane == 0) mbar_arrive_expect_tx(hbar + 8, tx_bytes);
              ts_fence_proxy_async();
              tc_fence_after();
            }
          }
          if (elect_one()) {
            if (GEPI && s == 0) {  // nothing to multiply at the first step: the epilogue takes G_0 alone
              mbar_arrive(accfull0 + 8 * (sb * 2));
              if constexpr (TILES == 2) mbar_arrive(accfull0 + 8 * (sb * 2 + 1));
            } else {
              if (s > 0) issue_h(dsub, 0, SPLIT ? 1 : 0, bd_lo);
              tc_commit(accfull0 + 8 * (sb * 2));
              if constexpr (TILES == 2) {
                if (s > 0) issue_h(dsub + NB, 1, SPLIT ? 1 : 0, bd_lo);
                tc_commit(accfull0 + 8 * (sb * 2 + 1));
              }
            }
            if (!GEPI && sb == SUBS - 1) tc_commit(gempty0 + 8 * slot);  // the stage is free once the MMAs that read it are done
          }
          __syncwarp();
          if (mprof && s > 0 && sb == 0) {
            const int m4 = clock();
            mc[0] += m1 - m0;  // accumulator-drained waits + P.G issue
            mc[1] += m2 - m1;  // wait for h_{t-1}
            mc[2] += m3 - m2;  // re-arm + proxy fence
            mc[3] += m4 - m3;  // W_hh.h issue + commits
          }
        }
        if (++slot == GS) {
          slot = 0;
          gph ^= 1;
        }
      }
#ifdef TSSEP_DEBUG_KNOBS
      if (mprof && lane == 0)
        for (int i = 0; i < 4; ++i) a.prof[8 + i] = mc[i];
#endif
    };
    if (KS == 19) run(std::integral_constant<int, 19>{});  // U = 300 (Up = 304)
    else run(std::integral_constant<int, 0>{});
  } else {
    // ---- epilogue: gates, cell update, h exchange --------------------------------------------------
    const int wx = (warp - 2) >> 2;            // (tile, column group) of this warp
    const int tl = wx % TILES;                 // row tile handled by this warp
    const int q = warp & 3;                    // TMEM lane quarter
    const int cg = wx / TILES;                 // column group: NC batch rows
    const int sb = cg / EW, half = cg % EW;    // sub-batch and which NC columns of its tile
    const int gate = lane & 3;             // i, f, g, o
    const int ul = lane >> 2;              // unit within the warp's octet
    const bool is_g = gate == 2;
    // pre-activations arrive scaled by 1/2 for i, f, o:  sigmoid(x) = 0.5 tanh(x/2) + 0.5
    const float ka = is_g ? 1.0f : 0.5f, kb = is_g ? 0.0f : 0.5f;
    const uint32_t myT = sT + static_cast<uint32_t>(warp - 2) * (NC * 16);
    const int gtile = static_cast<int>(crank) * TILES + tl;  // row tile of the whole layer: units 32 gtile ...
    const int oc = (gtile & 1) * 4 + q;                      // unit octet = 16-byte chunk inside the 64-unit k-atom
    const int unit0 = gtile * 32 + q * 8;
    const bool oct_ok = unit0 < Up;
    // sender role: lane ships batch row r (of its sub-batch) to CTAs d0, d0 + 32/NC, ...
    constexpr int DG = 32 / NC;
    const int rl = lane % NC, r = half * NC + rl, d0 = lane / NC;
    const uint32_t chunk_off = static_cast<uint32_t>(sb) * sub_bytes + static_cast<uint32_t>(gtile >> 1) * kAtomB +
                               static_cast<uint32_t>(r >> 3) * 1024 + static_cast<uint32_t>(r & 7) * 128 +
                               ((static_cast<uint32_t>(oc) ^ static_cast<uint32_t>(r & 7)) << 4);
    constexpr int ND = ((TILES == 1 ? 16 : 8) + DG - 1) / DG;  // clusters of up to 16 CTAs with one tile per CTA
    uint32_t r_b[ND], r_bar[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const uint32_t d = static_cast<uint32_t>(d0 + j * DG);
      r_b[j] = d < C ? mapa(sB, d) + chunk_off : 0;
      r_bar[j] = d < C ? mapa(hfull0 + 32u * sb + (SPLIT ? 8u * tl : 0u), d) : 0;
    }
    const int grow = row0 + sb * NR + r;  // batch row of the launch
    __nv_bfloat16* hptr = a.H + (static_cast<int64_t>(grow) * T) * (2 * static_cast<int64_t>(Up)) + dir * Up + unit0 +
                          (dir ? static_cast<int64_t>(T - 1) * 2 * Up : 0);  // frame of step 0
    const int64_t h_step = dir ? -2ll * Up : 2ll * Up;
    const bool h_store = oct_ok && d0 == 0 && grow < a.rows;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col + (sb * TILES + tl) * NB + half * NC;
    const uint32_t my_accfull = accfull0 + 8 * (sb * 2 + tl), my_accempty = accempty0 + 8 * (sb * 2 + tl);

    float cst[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) cst[i] = 0.f;
    // GEPI: this lane's G values of a step: gate plane `gate`, box rows sb*NR + half*NC + i, unit (gtile & 1)*32 + q*8 + ul
    // of the 64-unit box; 128-byte rows, 16-byte chunks XOR-swizzled with the row (row base is a multiple of 8)
    const uint32_t g_lane = sG + static_cast<uint32_t>(gate) * kAtomG + static_cast<uint32_t>(sb * NR + half * NC) * 128u +
                            static_cast<uint32_t>(ul) * 2u;
    const float g_scale = is_g ? 1.0f : 0.5f;  // the sigmoid gates work on x / 2 (W_hh rows are pre-scaled)
    int gslot = 0;
    uint32_t gph = 0;

#ifdef TSSEP_DEBUG_KNOBS
    const bool do_prof = a.prof != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && crank == 0 && (warp == 4 || warp == 8);
#else
    constexpr bool do_prof = false;
#endif
    int pc[4] = {0, 0, 0, 0};
    for (int s = 0; s < T; ++s) {
      const int wb = s & 1;
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (do_prof) c0 = clock();
      uint32_t gv[GEPI ? NC / 2 : 1];  // bf16 pairs (rows 2j, 2j + 1)
      if constexpr (GEPI) {
        // G_s arrived steps ago (the ring runs ahead): fetch it while the accumulator is still being computed
        mbar_wait(gfull0 + 8 * gslot, gph);
        const uint32_t gs = g_lane + static_cast<uint32_t>(gslot) * g_stage;
#pragma unroll
        for (int j = 0; j < NC / 2; ++j) {
          uint32_t lo, hi;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(gs + (2 * j) * 128u + ((static_cast<uint32_t>(oc) ^ ((2 * j) & 7u)) << 4)));
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(hi) : "r"(gs + (2 * j + 1) * 128u + ((static_cast<uint32_t>(oc) ^ ((2 * j + 1) & 7u)) << 4)));
          gv[j] = lo | (hi << 16);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gempty0 + 8 * gslot);
        if (++gslot == GS) {
          gslot = 0;
          gph ^= 1;
        }
      }
      mbar_wait(my_accfull, s & 1);
      if (do_prof) c1 = clock();
      tc_fence_after();
      uint32_t v[NC];
      if constexpr (NC == 32) tc_ld32(t_acc, v);
      else if constexpr (NC == 16) tc_ld16(t_acc, v);
      else if constexpr (NC == 8) tc_ld8(t_acc, v);
      else tc_ld4(t_acc, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(my_accempty);  // the next step's P.G MMAs may overwrite the accumulator
      if (do_prof) c2 = clock();

      // gate non-linearity of this lane's row for the warp's NC batch rows
      float act[NC];
      if constexpr (GEPI) {
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const float gx = __uint_as_float((i & 1) ? (gv[i / 2] & 0xFFFF0000u) : (gv[i / 2] << 16));
          const float d = s > 0 ? __uint_as_float(v[i]) : 0.f;  // step 0: h_{-1} = 0, nothing was accumulated
          act[i] = fmaf(ts_tanh<MATH>(fmaf(gx, g_scale, d)), ka, kb);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NC; ++i) act[i] = fmaf(ts_tanh<MATH>(__uint_as_float(v[i])), ka, kb);
      }
      // 4x4 transposes inside lane quads: afterwards act[4i + g] = gate g of batch row 4i + (lane & 3)
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        float* e = act + 4 * i;
        {
          const float x0 = (gate & 2) ? e[0] : e[2], x1 = (gate & 2) ? e[1] : e[3];
          const float y0 = __shfl_xor_sync(0xffffffffu, x0, 2), y1 = __shfl_xor_sync(0xffffffffu, x1, 2);
          if (gate & 2) {
            e[0] = y0;
            e[1] = y1;
          } else {
            e[2] = y0;
            e[3] = y1;
          }
        }
        {
          const float x0 = (gate & 1) ? e[0] : e[1], x1 = (gate & 1) ? e[2] : e[3];
          const float y0 = __shfl_xor_sync(0xffffffffu, x0, 1), y1 = __shfl_xor_sync(0xffffffffu, x1, 1);
          if (gate & 1) {
            e[0] = y0;
            e[2] = y1;
          } else {
            e[1] = y0;
            e[3] = y1;
          }
        }
      }
      // cell update for the NQ batch rows 4i + gate this lane now owns; stage h as bf16 in T[b][ul]
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const float ig = act[4 * i + 0], fg = act[4 * i + 1], gg = act[4 * i + 2], og = act[4 * i + 3];
        const float c = fmaf(fg, cst[i], ig * gg);
        cst[i] = c;
        const float h = og * ts_tanh<MATH>(c);
        const int b = 4 * i + gate;
        if constexpr (SAVE) {
          // training: the backward recurrence (csrc/lstm_bwd.cu) needs the gate activations and c_t of every step
          const int srow = row0 + sb * NR + half * NC + b;
          if (oct_ok && srow < a.rows) {
            const int64_t e = ((static_cast<int64_t>(srow) * T + (dir ? T - 1 - s : s)) * 2 + dir) * Up + unit0 + ul;
            a.gates_out[e] = make_uint2(pack_bf16x2(ig, fg), pack_bf16x2(gg, og));
            a.c_out[e] = c;
          }
        }
        const __nv_bfloat16 hb = __float2bfloat16_rn(h);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(myT + static_cast<uint32_t>(b * 8 + ul) * 2),
                     "h"(*reinterpret_cast<const unsigned short*>(&hb))
                     : "memory");
      }
      __syncwarp();
      if (do_prof) c3 = clock();
      uint4 chunk;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(chunk.x), "=r"(chunk.y), "=r"(chunk.z), "=r"(chunk.w)
                   : "r"(myT + static_cast<uint32_t>(rl) * 16));
      __syncwarp();
      if (s + 1 < T) {
        const uint32_t boff = static_cast<uint32_t>(wb) * NA * kAtomB;
        // st.async: the bytes complete the destination's mbarrier themselves.  (Measured alternatives, all slower:
        // plain st.shared::cluster + one release-arrive per warp, one cp.async.bulk per peer, per-atom barriers.)
#pragma unroll
        for (int j = 0; j < ND; ++j)
          if (static_cast<uint32_t>(d0 + j * DG) < C) ts_st_async_v4(r_b[j] + boff, chunk, r_bar[j] + 16 * wb);
      }
      if (h_store) *reinterpret_cast<uint4*>(hptr) = chunk;
      hptr += h_step;
      if (do_prof) {
        const int c4 = clock();
        pc[0] += c1 - c0;  // wait for the accumulator (h exchange of the cluster + MMAs)
        pc[1] += c2 - c1;  // tcgen05.ld
        pc[2] += c3 - c2;  // gates + transposes + cell update
        pc[3] += c4 - c3;  // chunk regroup + sends + H store
      }
    }
#ifdef TSSEP_DEBUG_KNOBS
    if (do_prof && lane == 0)
      for (int i = 0; i < 4; ++i) a.prof[tl * 4 + i] = pc[i];
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512);
  }
  cluster_sync_all();
}

// weight_hh (4U, U) f32 -> the word image the kernel stores into tensor memory:
// [dir][cta][tile][kstep][row 128][8 words], word j of k-step k = bf16 pair (k*16 + 2j, k*16 + 2j + 1);
// rows of the sigmoid gates (i, f, o) scaled by 1/2 (exact in bf16)
__global__ void pack_whh_ts_kernel(const float* __restrict__ w_fwd, const float* __restrict__ w_bwd, int U, int Up, int C,
                                   int KS, uint32_t* __restrict__ out) {
  const int64_t total = 2ll * C * 2 * KS * 128 * 8;
  for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t rr = o;
    const int j = static_cast<int>(rr % 8);
    rr /= 8;
    const int m = static_cast<int>(rr % 128);
    rr /= 128;
    const int k = static_cast<int>(rr % KS);
    rr /= KS;
    const int tile = static_cast<int>(rr % 2);
    rr /= 2;
    const int cta = static_cast<int>(rr % C);
    const int dir = static_cast<int>(rr / C);
    const int unit = cta * 64 + tile * 32 + (m >> 2), gate = m & 3;
    const int k0 = k * 16 + 2 * j;
    const float* w = dir ? w_bwd : w_fwd;
    const float sc = gate == 2 ? 1.0f : 0.5f;
    const float lo = (unit < U && k0 < U) ? sc * w[(static_cast<size_t>(gate) * U + unit) * U + k0] : 0.f;
    const float hi = (unit < U && k0 + 1 < U) ? sc * w[(static_cast<size_t>(gate) * U + unit) * U + k0 + 1] : 0.f;
    out[o] = pack_bf16x2(lo, hi);
  }
}

template <int NR, int NC, int MATH, bool SPLIT, int TILES, int SUBS, bool SAVE = false>
static cudaError_t prepare_ts(int C, size_t smem) {
  auto* fn = blstm_rec_ts_kernel<NR, NC, MATH, SPLIT, TILES, SUBS, SAVE>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e == cudaSuccess && C > 8) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}

template <int NR, int NC, int MATH, bool SPLIT, int TILES, int SUBS>
static int max_clusters_ts(int C, size_t smem) {
  if (prepare_ts<NR, NC, MATH, SPLIT, TILES, SUBS>(C, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, 64, 2);
  cfg.blockDim = dim3(64 + 128 * TILES * SUBS * (NR / NC));
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, blstm_rec_ts_kernel<NR, NC, MATH, SPLIT, TILES, SUBS>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// shared memory of one CTA and the depth of its G ring; rows = batch rows per cluster = subs sub-batches of rows / subs
static size_t ts_smem(int NA, int rows, int tiles, int subs, int* stages_out, int want_stages) {
  const int NR = rows / subs;
  const int NB = NR < 16 ? 16 : NR;
  const size_t ring_stage = 4ull * rows * 128;
  const size_t fixed_bytes = 1024 + 2ull * subs * NA * NB * 128 + 1024 + 4ull * tiles * rows * 16 + 128 + 16 * kTsMaxStages + 16;
  int stages = static_cast<int>((200 * 1024 - fixed_bytes) / ring_stage);
  stages = stages > kTsMaxStages ? kTsMaxStages : stages;
  if (want_stages >= 2 && want_stages <= stages) stages = want_stages;
  *stages_out = stages;
  // the kernel owns all 512 TMEM columns of its SM: ask for more than half of the shared memory so
  // that no second CTA can be co-resident and block on tcgen05.alloc
  const size_t smem = fixed_bytes + stages * ring_stage;
  return smem < 120 * 1024 ? 120 * 1024 : smem;
}

// CTAs per cluster: every CTA owns `tiles` row tiles of 32 units; the tile count is even (two per 64-unit k-atom)
static int cluster_ctas(int Up, int tiles) { return 2 * ((Up + 63) / 64) / tiles; }

// co-resident clusters of one shape (cached per device: the query costs ~10 us and every launch asks);
// rows = batch rows per cluster (8, 16, 32 or 64) = subs sub-batches of rows / subs (8, 16 or 32)
static int clusters_for(int rows, int Up, int tiles, int subs) {
  static int cache[8][4][2][2][25];  // [device][rows 8/16/32/64][tiles 1/2][subs 1/2][Up / 16]; 0 = not asked yet, -1 = none fit
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 8) dev = 0;
  const int ri = rows == 8 ? 0 : (rows == 16 ? 1 : (rows == 32 ? 2 : 3)), ui = Up / 16;
  int& slot = cache[dev][ri][tiles - 1][subs - 1][ui];
  if (slot == 0) {
    const int C = cluster_ctas(Up, tiles), NA = (Up + 63) / 64, NR = rows / subs;
    int st = 0;
    const size_t smem = ts_smem(NA, rows, tiles, subs, &st, 0);
    int n = 0;
    if (C <= 16 && st >= 2 && (NR == 8 || NR == 16 || NR == 32)) {
      if (subs == 2)  // two tiles per CTA only (with one tile the ping-pong was measured slower than the two-tile shapes)
        n = tiles != 2 ? 0
                       : (NR == 8 ? max_clusters_ts<8, 8, 1, false, 2, 2>(C, smem)
                                  : (NR == 16 ? max_clusters_ts<16, 16, 1, false, 2, 2>(C, smem)
                                              : max_clusters_ts<32, 32, 1, false, 2, 2>(C, smem)));
      else if (tiles == 2)
        n = NR == 8 ? max_clusters_ts<8, 8, 1, false, 2, 1>(C, smem)
                    : (NR == 16 ? max_clusters_ts<16, 16, 1, false, 2, 1>(C, smem) : max_clusters_ts<32, 16, 1, false, 2, 1>(C, smem));
      else
        n = NR == 8 ? max_clusters_ts<8, 8, 1, false, 1, 1>(C, smem)
                    : (NR == 16 ? max_clusters_ts<16, 16, 1, false, 1, 1>(C, smem) : max_clusters_ts<32, 16, 1, false, 1, 1>(C, smem));
    }
    slot = n > 0 ? n : -1;
  }
  return slot > 0 ? slot : 0;
}

template <int NR, int NC, int MATH, bool SPLIT, int TILES, int SUBS, bool SAVE = false>
static int launch_ts(const RecTsArgs& a, const CUtensorMap& gmap, int C, int nsub, size_t smem, cudaStream_t stream) {
  TSSEP_CUDA((prepare_ts<NR, NC, MATH, SPLIT, TILES, SUBS, SAVE>(C, smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, static_cast<unsigned>(nsub), 2);
  cfg.blockDim = dim3(64 + 128 * TILES * SUBS * (NR / NC));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TSSEP_CUDA(cudaLaunchKernelEx(&cfg, blstm_rec_ts_kernel<NR, NC, MATH, SPLIT, TILES, SUBS, SAVE>, a, gmap));
  return check_launch("blstm_rec_ts");
}

// Relative cost of one dependent step per cluster shape (tiles per CTA, rows per cluster, sub-batches), measured at
// U = 300 on B200 (profiles/r2_rec_ts_microbench.txt): ~us per step.  tssep_b200/dist.py::TS_STEP_COST mirrors the best
// shape per rows-per-cluster.  16 and 32 rows per cluster run fastest as two sub-batches in anti-phase (1.0 against
// 1.15-1.25 us and 1.5 against 2.0 us): while the h of one sub-batch travels through DSMEM and its gate math runs, the
// tensor pipe works on the other.
struct TsShape {
  int tiles, rows, subs;  // rows per cluster = subs sub-batches (advanced in anti-phase) of rows / subs
  double cost;
};
static const TsShape kTsShapes[] = {{1, 8, 1, 0.78},  {2, 8, 1, 0.85}, {1, 16, 1, 1.07}, {2, 16, 1, 1.2},  {2, 16, 2, 1.0},
                                    {1, 32, 1, 1.77}, {2, 32, 1, 2.0}, {2, 32, 2, 1.54}, {2, 64, 2, 3.06}};

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_pack_whh_ts(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* Wimg, tssep_stream_t stream) {
  TSSEP_REQUIRE(whh_fwd && whh_bwd && Wimg, "tssep_pack_whh_ts: null pointer");
  TSSEP_REQUIRE(U >= 1 && Up >= U && Up % 16 == 0 && Up <= 384, "tssep_pack_whh_ts: need U <= Up, Up %% 16 == 0, Up <= 384");
  const int C = (Up + 63) / 64, KS = Up / 16;
  const int64_t total = 2ll * C * 2 * KS * 128 * 8;
  const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 32));
  pack_whh_ts_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(whh_fwd, whh_bwd, U, Up, C, KS, Wimg);
  return check_launch("tssep_pack_whh_ts");
}

int tssep_blstm_recurrence_ts_capacity(int Up, int rows_per_cluster, int tiles_per_cta, int sub_batches) {
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 384, "tssep_blstm_recurrence_ts_capacity: bad Up");
  TSSEP_REQUIRE(rows_per_cluster == 8 || rows_per_cluster == 16 || rows_per_cluster == 32 || rows_per_cluster == 64,
                "tssep_blstm_recurrence_ts_capacity: rows_per_cluster must be 8, 16, 32 or 64");
  TSSEP_REQUIRE(tiles_per_cta == 1 || tiles_per_cta == 2, "tssep_blstm_recurrence_ts_capacity: tiles_per_cta must be 1 or 2");
  TSSEP_REQUIRE(sub_batches == 1 || sub_batches == 2, "tssep_blstm_recurrence_ts_capacity: sub_batches must be 1 or 2");
  if ((sub_batches == 2 && (tiles_per_cta == 1 || rows_per_cluster == 8)) || (sub_batches == 1 && rows_per_cluster == 64)) return 0;
  return (clusters_for(rows_per_cluster, Up, tiles_per_cta, sub_batches) / 2) * rows_per_cluster;
}

static int recurrence_ts_impl(const uint16_t* G, const uint32_t* Wimg, uint16_t* H, int64_t rows, int64_t T, int Up,
                              int rows_per_cluster, int tiles_per_cta, int sub_batches, int gate_math, int k_split,
                              uint16_t* gates_out, float* c_out, tssep_stream_t stream) {
  TSSEP_REQUIRE(G && Wimg && H, "tssep_blstm_recurrence_ts: null pointer");
  const bool save = gates_out != nullptr;
  TSSEP_REQUIRE(save == (c_out != nullptr), "tssep_blstm_recurrence_train: gates and cstate go together");
  TSSEP_REQUIRE(!save || (reinterpret_cast<uintptr_t>(gates_out) & 7) == 0, "tssep_blstm_recurrence_train: gates must be 8-byte aligned");
  if (save) {  // the training variant is instantiated for the two-tile shapes of up to 32 rows, one sub-batch, one K phase
    TSSEP_REQUIRE(tiles_per_cta != 1 && k_split != 1 && rows_per_cluster != 64 && sub_batches != 2,
                  "tssep_blstm_recurrence_train: needs two row tiles per CTA, at most 32 rows per cluster, one K phase");
    tiles_per_cta = 2;
    sub_batches = 1;
  }
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 384, "tssep_blstm_recurrence_ts: Up must be a multiple of 16 in [16, 384]");
  TSSEP_REQUIRE(rows >= 0 && T >= 0 && T < (1ll << 30) && (rows + 7) / 8 <= 65535, "tssep_blstm_recurrence_ts: bad extent");
  TSSEP_REQUIRE(rows_per_cluster == 0 || rows_per_cluster == 8 || rows_per_cluster == 16 || rows_per_cluster == 32 ||
                    rows_per_cluster == 64,
                "tssep_blstm_recurrence_ts: rows_per_cluster must be 0 (auto), 8, 16, 32 or 64");
  TSSEP_REQUIRE(tiles_per_cta >= 0 && tiles_per_cta <= 2, "tssep_blstm_recurrence_ts: tiles_per_cta must be 0 (auto), 1 or 2");
  TSSEP_REQUIRE(sub_batches >= 0 && sub_batches <= 2, "tssep_blstm_recurrence_ts: sub_batches must be 0 (auto), 1 or 2");
  TSSEP_REQUIRE(gate_math == 0 || gate_math == 1, "tssep_blstm_recurrence_ts: gate_math must be 0 (exp based) or 1 (tanh.approx)");
  TSSEP_REQUIRE(k_split >= -1 && k_split <= 1, "tssep_blstm_recurrence_ts: k_split must be -1 (default), 0 or 1");
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(H) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wimg) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                "tssep_blstm_recurrence_ts: G, H and Wimg must be 16-byte aligned");
  if (rows == 0 || T == 0) return 0;
  const int NA = (Up + 63) / 64;
  int RPC = rows_per_cluster, tiles = tiles_per_cta, subs = sub_batches;  // RPC: batch rows per cluster
  if (const char* e = debug_env("TSSEP_TS_ROWS")) {
    const int v = atoi(e);
    if (v == 8 || v == 16 || v == 32 || v == 64) RPC = v;
  }
  const bool split = k_split == 1;  // default: one phase (the second barrier + fence cost more than the split hides)
  if (split) {
    TSSEP_REQUIRE(tiles != 1 && subs != 2, "tssep_blstm_recurrence_ts: k_split needs two row tiles per CTA and one sub-batch");
    tiles = 2;
    subs = 1;
  }
  {
    // Fewer rows per cluster and fewer tiles per CTA = shorter step (epilogue math, DSMEM exchange and the MMAs of a
    // step scale with them) but fewer rows per wave of co-resident clusters; a launch that does not fit in one wave
    // runs its waves back to back.  Whatever the caller fixed narrows the candidates.
    double best = 1e30;
    const TsShape* pick = nullptr;
    for (const TsShape& sh : kTsShapes) {
      if ((RPC != 0 && sh.rows != RPC) || (tiles != 0 && sh.tiles != tiles) || (subs != 0 && sh.subs != subs) ||
          (save && sh.subs != 1) || (split && (sh.tiles != 2 || sh.subs != 1)))
        continue;
      const int m = clusters_for(sh.rows, Up, sh.tiles, sh.subs);
      if (m < 2) continue;
      const int64_t n = 2 * ((rows + sh.rows - 1) / sh.rows);
      const double t = sh.cost * static_cast<double>((n + m - 1) / m);
      if (t < best) {
        best = t;
        pick = &sh;
      }
    }
    TSSEP_REQUIRE(pick != nullptr,
                  "tssep_blstm_recurrence_ts: no cluster shape with rows_per_cluster=%d tiles_per_cta=%d sub_batches=%d fits (Up=%d)",
                  rows_per_cluster, tiles_per_cta, sub_batches, Up);
    RPC = pick->rows;
    tiles = pick->tiles;
    subs = pick->subs;
  }
  const int NR = RPC / subs;  // batch rows per sub-batch
  const int C = cluster_ctas(Up, tiles);
  TSSEP_REQUIRE(C <= 16, "tssep_blstm_recurrence_ts: Up=%d needs clusters of %d CTAs with %d tile(s) per CTA (max 16)", Up, C, tiles);
  const uint32_t a_tile_cols = (static_cast<uint32_t>(Up) / 2 + 31u) & ~31u;
  TSSEP_REQUIRE(tiles * a_tile_cols + 64 + tiles * subs * (NR < 16 ? 16 : NR) <= 512,
                "tssep_blstm_recurrence_ts: Up=%d with %d rows per cluster exceeds the 512 tensor-memory columns", Up, RPC);
  RecTsArgs a;
  a.Wimg = reinterpret_cast<const uint4*>(Wimg);
  a.H = reinterpret_cast<__nv_bfloat16*>(H);
  a.rows = static_cast<int>(rows);
  a.T = static_cast<int>(T);
  a.Up = Up;
  a.NA = NA;
  a.KS = Up / 16;
  a.prof = nullptr;
  a.flags = 0;
  a.gates_out = reinterpret_cast<uint2*>(gates_out);
  a.c_out = c_out;
  int want_stages = 0;
#ifdef TSSEP_DEBUG_KNOBS
  if (const char* e = debug_env("TSSEP_REC_PROF")) a.prof = reinterpret_cast<int*>(strtoull(e, nullptr, 0));
  if (const char* e = debug_env("TSSEP_TS_NOFENCE")) a.flags |= atoi(e) ? 1 : 0;
  if (const char* e = debug_env("TSSEP_TS_STAGES")) want_stages = atoi(e);
#endif
  int stages = 0;
  const size_t smem = ts_smem(NA, RPC, tiles, subs, &stages, want_stages);
  TSSEP_REQUIRE(stages >= 2, "tssep_blstm_recurrence_ts: G ring does not fit shared memory");
  a.stages = stages;
  const int nsub = static_cast<int>((rows + RPC - 1) / RPC);
  // batch columns per epilogue warp (debug knob TSSEP_TS_COLS): 16 measured best for 32 rows per cluster
  int NC = NR < 16 ? NR : (subs == 2 ? NR : 16);
  if (const char* e = debug_env("TSSEP_TS_COLS")) {
    const int v = atoi(e);
    if (subs == 1 && (v == NR || (v == NR / 2 && v >= 8))) NC = v;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // G (rows, T, 2, 4, Up) bf16 viewed as (unit, row, gate, dir, t); box = 64 units x NR rows x 4 gates
  CUtensorMap gmap{};
  EncodeTiledFn enc = get_encode_tiled();
  TSSEP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t up = static_cast<cuuint64_t>(Up);
  cuuint64_t dims[5] = {up, static_cast<cuuint64_t>(rows), 4, 2, static_cast<cuuint64_t>(T)};
  cuuint64_t strides[4] = {static_cast<cuuint64_t>(T) * 8 * up * 2, up * 2, 4 * up * 2, 8 * up * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(RPC), 4, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<uint16_t*>(G), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSSEP_REQUIRE(r == CUDA_SUCCESS, "tssep_blstm_recurrence_ts: cuTensorMapEncodeTiled failed with code %d", static_cast<int>(r));

#define TSSEP_TS_CASE(NR_, NC_)                                                                                \
  if (subs == 1 && NR == NR_ && NC == NC_) {                                                                              \
    if (save)                                                                                                  \
      return gate_math ? launch_ts<NR_, NC_, 1, false, 2, 1, true>(a, gmap, C, nsub, smem, st)                 \
                       : launch_ts<NR_, NC_, 0, false, 2, 1, true>(a, gmap, C, nsub, smem, st);                \
    if (tiles == 1)                                                                                            \
      return gate_math ? launch_ts<NR_, NC_, 1, false, 1, 1>(a, gmap, C, nsub, smem, st)                       \
                       : launch_ts<NR_, NC_, 0, false, 1, 1>(a, gmap, C, nsub, smem, st);                      \
    if (gate_math) return split ? launch_ts<NR_, NC_, 1, true, 2, 1>(a, gmap, C, nsub, smem, st)               \
                                : launch_ts<NR_, NC_, 1, false, 2, 1>(a, gmap, C, nsub, smem, st);             \
    return split ? launch_ts<NR_, NC_, 0, true, 2, 1>(a, gmap, C, nsub, smem, st)                              \
                 : launch_ts<NR_, NC_, 0, false, 2, 1>(a, gmap, C, nsub, smem, st);                            \
  }
  TSSEP_TS_CASE(8, 8)
  TSSEP_TS_CASE(16, 8)
  TSSEP_TS_CASE(16, 16)
  TSSEP_TS_CASE(32, 16)
#undef TSSEP_TS_CASE
  // two sub-batches in anti-phase
#define TSSEP_TS_CASE2(NR_, TILES_)                                                                \
  if (subs == 2 && tiles == TILES_ && NR == NR_ && NC == NR_)                                      \
    return gate_math ? launch_ts<NR_, NR_, 1, false, TILES_, 2>(a, gmap, C, nsub, smem, st)        \
                     : launch_ts<NR_, NR_, 0, false, TILES_, 2>(a, gmap, C, nsub, smem, st);
  TSSEP_TS_CASE2(8, 2)
  TSSEP_TS_CASE2(16, 2)
  TSSEP_TS_CASE2(32, 2)
#undef TSSEP_TS_CASE2
  set_error("tssep_blstm_recurrence_ts: no instantiation for %d x %d rows per cluster, %d columns per warp", subs, NR, NC);
  return -1;
}

int tssep_blstm_recurrence_ts(const uint16_t* G, const uint32_t* Wimg, uint16_t* H, int64_t rows, int64_t T, int Up,
                              int rows_per_cluster, int tiles_per_cta, int sub_batches, int gate_math, int k_split,
                              tssep_stream_t stream) {
  return recurrence_ts_impl(G, Wimg, H, rows, T, Up, rows_per_cluster, tiles_per_cta, sub_batches, gate_math, k_split, nullptr,
                            nullptr, stream);
}

int tssep_blstm_recurrence_train(const uint16_t* G, const uint32_t* Wimg, uint16_t* H, uint16_t* gates, float* cstate,
                                 int64_t rows, int64_t T, int Up, int rows_per_cluster, int gate_math, tssep_stream_t stream) {
  TSSEP_REQUIRE(gates && cstate, "tssep_blstm_recurrence_train: null pointer");
  return recurrence_ts_impl(G, Wimg, H, rows, T, Up, rows_per_cluster, 2, 1, gate_math, 0, gates, cstate, stream);
}

}  // extern "C"
