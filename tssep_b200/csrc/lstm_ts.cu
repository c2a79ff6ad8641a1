// BLSTM recurrence with the recurrent weights resident in TENSOR MEMORY.
//
// Same operator as csrc/lstm.cu / csrc/lstm_tc.cu (the time loop of torch.nn.LSTM,
// tssep/train/rnnp.py:87-95, :143-159).  The shared-memory variant (lstm_tc.cu) is bound by the
// tensor core re-reading W_hh from shared memory every step (~78 cycles per M=128,N=32,K=16 MMA);
// here W_hh is written ONCE into TMEM and used as the A operand of tcgen05.mma (A from tensor
// memory, B = h_{t-1} from shared memory), so a step's contraction costs only the issue of
// 2 * Up/16 small MMAs and shared memory is free for a deep ring of input projections.
//
// One cluster of C = ceil(Up/64) CTAs per (NR batch rows, direction), NR = 16 or 32.  Each CTA owns
// 64 hidden units = 256 gate rows (row = 4*unit + gate) as two M=128 A tiles of Up/2 TMEM columns
// each (two bf16 per 32-bit column); the accumulators (2 x NR fp32 columns) sit behind them.
//   warp 0 (one thread)  streams the CTA's slice of G (input projections, GEMM "BT" tile layout)
//                        with cp.async.bulk through an mbarrier ring, several steps ahead;
//   warp 1 (one thread)  waits for h_{t-1} (NR x Up bf16, written by every CTA of the cluster
//                        through DSMEM), issues the MMAs of both tiles and commits each tile;
//   warps 2..            (lane = gate row; 8 * NR/NC of them, each owning NC batch columns of one row
//                        tile and TMEM lane quarter) tcgen05.ld the pre-activations, add G, apply the gates,
//                        transpose 4x4 blocks inside lane quads, update c_t (registers) and h_t,
//                        regroup 8 units into 16-byte chunks and push them to every CTA's next B
//                        operand with st.async (complete_tx on the destination mbarrier) and to H.
#include "../../include/tssep_b200.h"
#include "common.cuh"

#include <cstdlib>

namespace tssep {

constexpr int kTsMaxStages = 8;

struct RecTsArgs {
  const uint8_t* G;   // BT: tiles [group][t][dir][unit octet][b/4][4*(unit%8)+gate][b%4], f32 or bf16 (PLAIN: via gmap)
  const uint4* Wimg;  // [dir][cta][tile][kstep][row 128][8 words]
  __nv_bfloat16* H;   // BT: (groups, T, 32, 2*Up), rows ordered (group, t, b); PLAIN: (rows, T, 2*Up)
  int rows, T, Up, NA, KS, fast, g_bf16, stages;
  int* prof;
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
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// NR batch rows per cluster, NC of them per epilogue warp (8 * NR/NC epilogue warps); FAST: tanh.approx gates;
// GBF16: G stored as bf16.  Everything the per-step loops branch on is a template parameter: the epilogue is
// close to issue bound (ncu: 44 % of all issue slots over the whole step, profiles/r1_ncu_rec_ts.txt).
// PLAIN: G (rows, T, 2, 4, Up) bf16 and H (rows, T, 2*Up), the layouts of the register kernel (csrc/lstm.cu): one
// 5-D TMA box (64 units, 4 gates, NR rows) per step, 128-byte swizzled so that lane (unit, gate) reads without
// bank conflicts.  Otherwise the "BT" tile layout of the GEMM (rows ordered (group, t, b32)).
template <int NR, int NC, bool FAST, bool GBF16, bool PLAIN>
__global__ void __launch_bounds__(64 + 256 * (NR / NC), 1)
blstm_rec_ts_kernel(const RecTsArgs a, const __grid_constant__ CUtensorMap gmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t kAtomB = NR * 128;  // one 64-k atom of the B operand: NR rows x 128 bytes, 128-byte swizzle
  constexpr int EW = NR / NC;            // epilogue warps per (row tile, TMEM lane quarter)
  constexpr int NQ = NC / 4;             // batch-row quads per epilogue warp
  constexpr int SUBS = 32 / NR;          // clusters per 32-row group
  constexpr int kThreads = 64 + 256 * EW;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NA = a.NA, KS = a.KS, GS = a.stages;
  constexpr uint32_t esz = GBF16 ? 2u : 4u;
  const uint32_t oct_bytes = NR * 32u * esz;  // one unit octet of one step: [b/4][lane][b%4]
  const uint32_t g_stage = 8u * oct_bytes;
  const uint32_t sB = base;                         // [2 buffers][NA] x kAtomB
  const uint32_t sG = sB + 2u * NA * kAtomB;        // [GS] x g_stage
  const uint32_t sT = sG + GS * g_stage;            // [8 * EW warps] x NC x 16 B
  const uint32_t sBar = sT + 8u * NR * 16u;
  const uint32_t hfull0 = sBar, accfull0 = sBar + 16, gfull0 = sBar + 32, gempty0 = gfull0 + 8 * kTsMaxStages,
                 tptr = gempty0 + 8 * kTsMaxStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t C = cluster_nctarank();
  const int grp = blockIdx.y / SUBS, sub = blockIdx.y % SUBS, dir = blockIdx.z;
  const int row0 = blockIdx.y * NR;  // PLAIN: first batch row of this cluster
  const int T = a.T, Up = a.Up;
  const uint32_t tx_bytes = static_cast<uint32_t>(NR) * 128u * C;  // every CTA ships 64 units x NR rows
  const int n_oct = Up / 8;
  int oct_valid = n_oct - static_cast<int>(crank) * 8;
  oct_valid = oct_valid < 0 ? 0 : (oct_valid > 8 ? 8 : oct_valid);

  // ---- one-time setup ---------------------------------------------------------------------------
  for (uint32_t i = threadIdx.x; i < 2u * NA * kAtomB / 16; i += kThreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sB + 16 * i), "r"(0u) : "memory");
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(hfull0, 1);
      mbar_init(hfull0 + 8, 1);
      mbar_init(accfull0, 1);
      mbar_init(accfull0 + 8, 1);
      for (int i = 0; i < GS; ++i) {
        mbar_init(gfull0 + 8 * i, 1);
        mbar_init(gempty0 + 8 * i, 8 * EW);
      }
      mbar_fence_init();
      mbar_arrive_expect_tx(hfull0, tx_bytes);
      mbar_arrive_expect_tx(hfull0 + 8, tx_bytes);
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
  const uint32_t acc_col = 2u * a_tile_cols;                                   // accumulators behind the A tiles

  if (warp >= 2 && warp < 10) {
    // W_hh -> TMEM: lane = gate row of the tile, 8 columns (16 k values) per store
    const int tl = (warp - 2) >> 2, q = warp & 3;
    const uint4* src = a.Wimg + ((static_cast<size_t>(dir) * C + crank) * 2 + tl) * static_cast<size_t>(KS) * 256 +
                       static_cast<size_t>(q * 32 + lane) * 2;
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(tl) * a_tile_cols;
    for (int k = 0; k < KS; ++k) {
      const uint4 w0 = __ldg(src + static_cast<size_t>(k) * 256), w1 = __ldg(src + static_cast<size_t>(k) * 256 + 1);
      tc_st8(t0 + k * 8, w0, w1);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();

  if (warp == 0) {
    // ---- G producer ---------------------------------------------------------------------------------
    if (PLAIN) {
      if (lane == 0 && oct_valid > 0) {
        tma_prefetch_desc(&gmap);
        int slot = 0;
        uint32_t gph = 0;
        for (int s = 0; s < T; ++s) {
          const int t = dir ? T - 1 - s : s;
          mbar_wait(gempty0 + 8 * slot, gph ^ 1);
          mbar_arrive_expect_tx(gfull0 + 8 * slot, g_stage);
          tma_load_5d(sG + slot * g_stage, &gmap, gfull0 + 8 * slot, static_cast<int>(crank) * 64, 0, row0, dir, t);
          if (++slot == GS) {
            slot = 0;
            gph ^= 1;
          }
        }
      }
    } else if (lane == 0 && oct_valid > 0) {
      const int64_t tile_bytes = 1024ll * esz;  // one (group, t, dir, octet) tile: 32 rows x 32 columns
      const int64_t t_stride = 2ll * n_oct * tile_bytes;
      const uint8_t* g0 = a.G + ((static_cast<int64_t>(grp) * T * 2 + dir) * n_oct + crank * 8) * tile_bytes +
                          static_cast<int64_t>(sub) * oct_bytes;
      int slot = 0;
      uint32_t gph = 0;
      for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        mbar_wait(gempty0 + 8 * slot, gph ^ 1);
        mbar_arrive_expect_tx(gfull0 + 8 * slot, static_cast<uint32_t>(oct_valid) * oct_bytes);
        const uint8_t* src = g0 + static_cast<int64_t>(t) * t_stride;
        const uint32_t dst = sG + slot * g_stage;
        for (int o = 0; o < oct_valid; ++o) bulk_g2s(dst + o * oct_bytes, src + o * tile_bytes, oct_bytes, gfull0 + 8 * slot);
        if (++slot == GS) {
          slot = 0;
          gph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------
    // The whole warp runs the loop convergently and elect.sync guards only the MMA block: inside a
    // divergent `if (lane == 0)` ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY waterfall and
    // rebuilds the descriptor through a long uniform-datapath chain (~70 cycles per MMA, measured).
    // Descriptors advance by adding a constant to the encoded start address (16-byte units).
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NR >> 3) << 17) |
                           (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t bdesc0 = ts_desc_sw128(sB);
    const uint32_t buf_step = static_cast<uint32_t>(NA) * (kAtomB >> 4);  // encoded distance of the two h buffers
    const uint32_t d0 = tmem_base + acc_col;
    const int full_atoms = KS >> 2, rem = KS & 3;
    for (int s = 0; s < T; ++s) {
      const int rb = (s & 1) ^ 1;
      if (s > 0) {
        mbar_wait(hfull0 + 8 * rb, ((s - 1) >> 1) & 1);
        if (lane == 0) mbar_arrive_expect_tx(hfull0 + 8 * rb, tx_bytes);  // re-arm for the data of step s+1
        // h arrived through st.async (generic proxy); the MMA reads it through the async proxy
        ts_fence_proxy_async();
      }
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bd = bdesc0 + static_cast<uint64_t>(rb ? buf_step : 0u);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const uint32_t d = d0 + tile * NR;
          uint32_t at = tmem_base + static_cast<uint32_t>(tile) * a_tile_cols;
          uint64_t bk = bd;
          for (int atom = 0; atom < full_atoms; ++atom) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              tc_mma_bf16_ts(d, at + k4 * 8, bk + 2 * k4, idesc, (atom > 0 || k4 > 0) ? 1u : 0u);
            at += 32;
            bk += kAtomB >> 4;
          }
#pragma unroll
          for (int k4 = 0; k4 < 3; ++k4)
            if (k4 < rem) tc_mma_bf16_ts(d, at + k4 * 8, bk + 2 * k4, idesc, (full_atoms > 0 || k4 > 0) ? 1u : 0u);
          tc_commit(accfull0 + 8 * tile);
        }
      }
      __syncwarp();
    }
  } else {
    // ---- epilogue: gates, cell update, h exchange --------------------------------------------------
    const int tl = ((warp - 2) >> 2) & 1;  // row tile handled by this warp
    const int q = warp & 3;                // TMEM lane quarter
    const int half = (warp - 2) >> 3;      // which NC columns (batch rows) of the tile
    const int gate = lane & 3;       // i, f, g, o
    const int ul = lane >> 2;        // unit within the warp's octet
    const bool is_g = gate == 2;
    const float sc = FAST ? (is_g ? 1.0f : 0.5f) : (is_g ? 2.0f : 1.0f);
    const float ka = FAST ? (is_g ? 1.0f : 0.5f) : (is_g ? 2.0f : 1.0f);
    const float kb = FAST ? (is_g ? 0.0f : 0.5f) : (is_g ? -1.0f : 0.0f);
    const uint32_t myT = sT + static_cast<uint32_t>(warp - 2) * (NC * 16);
    const int oc = tl * 4 + q;  // unit octet inside the CTA = 16-byte chunk of the CTA's k-atom
    const int unit0 = static_cast<int>(crank) * 64 + oc * 8;
    const bool oct_ok = unit0 < Up;
    // sender role: lane ships batch row r to CTAs d0, d0 + 32/NC, ...
    constexpr int DG = 32 / NC;
    const int rl = lane % NC, r = half * NC + rl, d0 = lane / NC;
    const uint32_t chunk_off = crank * kAtomB + static_cast<uint32_t>(r >> 3) * 1024 + static_cast<uint32_t>(r & 7) * 128 +
                               ((static_cast<uint32_t>(oc) ^ static_cast<uint32_t>(r & 7)) << 4);
    constexpr int ND = (8 + DG - 1) / DG;
    uint32_t r_b[ND], r_bar[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const uint32_t d = static_cast<uint32_t>(d0 + j * DG);
      r_b[j] = d < C ? mapa(sB, d) + chunk_off : 0;
      r_bar[j] = d < C ? mapa(hfull0, d) : 0;
    }
    __nv_bfloat16* hbase =
        PLAIN ? a.H + (static_cast<int64_t>(row0 + r) * T) * (2 * static_cast<int64_t>(Up)) + dir * Up + unit0
              : a.H + ((static_cast<int64_t>(grp) * T) * 32 + sub * NR + r) * (2 * static_cast<int64_t>(Up)) + dir * Up + unit0;
    const int64_t h_tstride = PLAIN ? 2ll * Up : 32ll * 2 * Up;
    __nv_bfloat16* hptr = hbase + (dir ? static_cast<int64_t>(T - 1) * h_tstride : 0);  // frame of step 0
    const int64_t h_step = dir ? -h_tstride : h_tstride;
    int slot = 0;
    uint32_t gph = 0;
    const bool h_store = oct_ok && d0 == 0 && (!PLAIN || row0 + r < a.rows);
    const uint32_t g_lane = static_cast<uint32_t>(oc) * oct_bytes + static_cast<uint32_t>(half * NQ * 32 + lane) * 4u * esz;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col + tl * NR + half * NC;

    float cst[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) cst[i] = 0.f;

    const bool do_prof = a.prof != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && crank == 0 && (warp == 4 || warp == 8);
    int pc[4] = {0, 0, 0, 0};
    for (int s = 0; s < T; ++s) {
      const int wb = s & 1;
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (do_prof) c0 = clock();
      // input projection of this step: shared-memory ring -> registers, slot handed back at once
      float gv[NC];
      mbar_wait(gfull0 + 8 * slot, gph);
      {
        const uint32_t gp = sG + slot * g_stage + g_lane;
        if constexpr (PLAIN) {
          // stage = [row][gate][64 units] bf16, 128-byte lines, 16-byte chunks XOR-swizzled with (line & 7)
#pragma unroll
          for (int i = 0; i < NC; ++i) {
            const uint32_t line = static_cast<uint32_t>((half * NC + i) * 4 + gate);
            unsigned short hv = 0;
            if (oct_ok)
              asm volatile("ld.shared.u16 %0, [%1];"
                           : "=h"(hv)
                           : "r"(sG + slot * g_stage + line * 128u + ((static_cast<uint32_t>(oc) ^ (line & 7u)) << 4) + ul * 2u));
            gv[i] = __uint_as_float(static_cast<uint32_t>(hv) << 16);
          }
        } else {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          if (!oct_ok) {
            gv[4 * i] = gv[4 * i + 1] = gv[4 * i + 2] = gv[4 * i + 3] = 0.f;
          } else if (GBF16) {
            uint32_t x, y;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(gp + i * 256));
            gv[4 * i + 0] = __uint_as_float(x << 16);
            gv[4 * i + 1] = __uint_as_float(x & 0xffff0000u);
            gv[4 * i + 2] = __uint_as_float(y << 16);
            gv[4 * i + 3] = __uint_as_float(y & 0xffff0000u);
          } else {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(gv[4 * i]), "=f"(gv[4 * i + 1]), "=f"(gv[4 * i + 2]), "=f"(gv[4 * i + 3])
                         : "r"(gp + i * 512));
          }
        }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(gempty0 + 8 * slot);
      if (++slot == GS) {
        slot = 0;
        gph ^= 1;
      }
      if (do_prof) c1 = clock();

      mbar_wait(accfull0 + 8 * tl, s & 1);
      if (do_prof) c2 = clock();
      tc_fence_after();
      uint32_t v[NC];
      if constexpr (NC == 32) tc_ld32(t_acc, v);
      else if constexpr (NC == 16) tc_ld16(t_acc, v);
      else tc_ld8(t_acc, v);
      tc_wait_ld();
      tc_fence_before();

      // gate non-linearity of this lane's row for the warp's NC batch rows
      float act[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const float x = (__uint_as_float(v[i]) + gv[i]) * sc;
        const float y = FAST ? tanh_fast(x) : sigmoid_acc(x);
        act[i] = fmaf(y, ka, kb);
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
        const float h = og * (FAST ? tanh_fast(c) : tanh_acc(c));
        const int b = 4 * i + gate;
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
          if (static_cast<uint32_t>(d0 + j * DG) < C) ts_st_async_v4(r_b[j] + boff, chunk, r_bar[j] + 8 * wb);
      }
      if (h_store) *reinterpret_cast<uint4*>(hptr) = chunk;
      hptr += h_step;
      if (do_prof) {
        const int c4 = clock();
        pc[0] += c1 - c0;  // G ring wait + loads
        pc[1] += c2 - c1;  // wait for the accumulator (h exchange of the cluster + MMAs)
        pc[2] += c3 - c2;  // tcgen05.ld + gates + transposes + cell update
        pc[3] += c4 - c3;  // chunk regroup + sends + H store
      }
    }
    if (do_prof && lane == 0)
      for (int i = 0; i < 4; ++i) a.prof[tl * 4 + i] = pc[i];
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
// [dir][cta][tile][kstep][row 128][8 words], word j of k-step k = bf16 pair (k*16 + 2j, k*16 + 2j + 1)
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
    const float lo = (unit < U && k0 < U) ? w[(static_cast<size_t>(gate) * U + unit) * U + k0] : 0.f;
    const float hi = (unit < U && k0 + 1 < U) ? w[(static_cast<size_t>(gate) * U + unit) * U + k0 + 1] : 0.f;
    out[o] = pack_bf16x2(lo, hi);
  }
}

template <int NR, int NC, bool FAST, bool GBF16, bool PLAIN>
static int max_clusters_ts(int C, size_t smem) {
  if (cudaFuncSetAttribute(blstm_rec_ts_kernel<NR, NC, FAST, GBF16, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
      cudaSuccess)
    return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, 64, 2);
  cfg.blockDim = dim3(64 + 256 * (NR / NC));
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, blstm_rec_ts_kernel<NR, NC, FAST, GBF16, PLAIN>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// shared memory of one CTA and the depth of its G ring
static size_t ts_smem(int C, int NR, int g_dtype, int* stages_out) {
  const size_t gsz = g_dtype ? 2 : 4;
  const size_t ring_stage = 8ull * NR * 32 * gsz;
  const size_t fixed_bytes = 1024 + 2ull * C * NR * 128 + 8ull * NR * 16 + 32 + 16 * kTsMaxStages + 16;
  int stages = static_cast<int>((200 * 1024 - fixed_bytes) / ring_stage);
  stages = stages > kTsMaxStages ? kTsMaxStages : stages;
  if (const char* e = getenv("TSSEP_TS_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= stages) stages = v;
  }
  *stages_out = stages;
  // the kernel owns all 512 TMEM columns of its SM: ask for more than half of the shared memory so
  // that no second CTA can be co-resident and block on tcgen05.alloc
  const size_t smem = fixed_bytes + stages * ring_stage;
  return smem < 120 * 1024 ? 120 * 1024 : smem;
}

template <int NR, int NC, bool FAST, bool GBF16, bool PLAIN>
static int launch_ts(const RecTsArgs& a, const CUtensorMap& gmap, int C, int nsub, size_t smem, cudaStream_t stream) {
  TSSEP_CUDA(cudaFuncSetAttribute(blstm_rec_ts_kernel<NR, NC, FAST, GBF16, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, static_cast<unsigned>(nsub), 2);
  cfg.blockDim = dim3(64 + 256 * (NR / NC));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TSSEP_CUDA(cudaLaunchKernelEx(&cfg, blstm_rec_ts_kernel<NR, NC, FAST, GBF16, PLAIN>, a, gmap));
  return check_launch("blstm_rec_ts");
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_pack_whh_ts(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* Wimg, tssep_stream_t stream) {
  TSSEP_REQUIRE(whh_fwd && whh_bwd && Wimg, "tssep_pack_whh_ts: null pointer");
  TSSEP_REQUIRE(U >= 1 && Up >= U && Up % 16 == 0 && Up <= 448, "tssep_pack_whh_ts: need U <= Up, Up %% 16 == 0, Up <= 448");
  const int C = (Up + 63) / 64, KS = Up / 16;
  const int64_t total = 2ll * C * 2 * KS * 128 * 8;
  const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 32));
  pack_whh_ts_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(whh_fwd, whh_bwd, U, Up, C, KS, Wimg);
  return check_launch("tssep_pack_whh_ts");
}

int tssep_blstm_recurrence_ts_capacity(int Up, int rows_per_cluster, int g_dtype) {
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 448, "tssep_blstm_recurrence_ts_capacity: bad Up");
  TSSEP_REQUIRE(rows_per_cluster == 16 || rows_per_cluster == 32,
                "tssep_blstm_recurrence_ts_capacity: rows_per_cluster must be 16 or 32");
  const int C = (Up + 63) / 64;
  int st = 0;
  const int m = rows_per_cluster == 16 ? max_clusters_ts<16, 16, true, true, false>(C, ts_smem(C, 16, g_dtype, &st))
                                       : max_clusters_ts<32, 16, true, true, false>(C, ts_smem(C, 32, g_dtype, &st));
  return (m / 2) * rows_per_cluster;
}

int tssep_blstm_recurrence_ts(const void* G, int g_dtype, const uint32_t* Wimg, uint16_t* H, int64_t rows, int64_t T,
                              int Up, int layout, int rows_per_cluster, int fast_math, tssep_stream_t stream) {
  TSSEP_REQUIRE(G && Wimg && H, "tssep_blstm_recurrence_ts: null pointer");
  TSSEP_REQUIRE(layout == TSSEP_REC_LAYOUT_BT || layout == TSSEP_REC_LAYOUT_ROWS, "tssep_blstm_recurrence_ts: unknown layout %d", layout);
  TSSEP_REQUIRE(layout == TSSEP_REC_LAYOUT_BT || g_dtype == 1, "tssep_blstm_recurrence_ts: the row layout needs bf16 G");
  TSSEP_REQUIRE(g_dtype == 0 || g_dtype == 1, "tssep_blstm_recurrence_ts: g_dtype must be 0 (f32) or 1 (bf16)");
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 448, "tssep_blstm_recurrence_ts: Up must be a multiple of 16 in [16, 448]");
  TSSEP_REQUIRE(rows >= 0 && T >= 0 && T < (1ll << 31) && (rows + 15) / 16 <= 65535, "tssep_blstm_recurrence_ts: bad extent");
  TSSEP_REQUIRE(rows_per_cluster == 0 || rows_per_cluster == 16 || rows_per_cluster == 32,
                "tssep_blstm_recurrence_ts: rows_per_cluster must be 0 (auto), 16 or 32");
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(H) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wimg) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                "tssep_blstm_recurrence_ts: G, H and Wimg must be 16-byte aligned");
  if (rows == 0 || T == 0) return 0;
  const int C = (Up + 63) / 64;
  int NR = rows_per_cluster;
  if (const char* e = getenv("TSSEP_TS_ROWS")) {
    const int v = atoi(e);
    if (v == 16 || v == 32) NR = v;
  }
  if (NR == 0) {
    // 16 rows per cluster has the shortest step (1.3-1.4 us at U=300 vs 2.25 us for 32 rows); a launch that
    // does not fit in one wave of co-resident clusters runs its waves back to back
    int st = 0;
    const int m16 = max_clusters_ts<16, 16, true, true, false>(C, ts_smem(C, 16, g_dtype, &st));
    const int m32 = max_clusters_ts<32, 16, true, true, false>(C, ts_smem(C, 32, g_dtype, &st));
    const int64_t n16 = 2 * ((rows + 15) / 16), n32 = 2 * ((rows + 31) / 32);
    const double t16 = m16 > 0 ? 1.0 * static_cast<double>((n16 + m16 - 1) / m16) : 1e9;
    const double t32 = m32 > 0 ? 1.6 * static_cast<double>((n32 + m32 - 1) / m32) : 2e9;
    NR = t16 <= t32 ? 16 : 32;
  }
  RecTsArgs a;
  a.G = static_cast<const uint8_t*>(G);
  a.Wimg = reinterpret_cast<const uint4*>(Wimg);
  a.H = reinterpret_cast<__nv_bfloat16*>(H);
  a.rows = static_cast<int>(rows);
  a.T = static_cast<int>(T);
  a.Up = Up;
  a.NA = C;
  a.KS = Up / 16;
  a.fast = fast_math & 1;
  a.g_bf16 = g_dtype;
  a.prof = nullptr;
  if (const char* e = getenv("TSSEP_REC_PROF")) a.prof = reinterpret_cast<int*>(strtoull(e, nullptr, 0));
  int stages = 0;
  const size_t smem = ts_smem(C, NR, g_dtype, &stages);
  TSSEP_REQUIRE(stages >= 2, "tssep_blstm_recurrence_ts: G ring does not fit shared memory");
  a.stages = stages;
  const int nsub = static_cast<int>((rows + NR - 1) / NR);
  // batch columns per epilogue warp (TSSEP_TS_COLS): 16 measured best for both cluster widths
  int NC = 16;
  if (const char* e = getenv("TSSEP_TS_COLS")) {
    const int v = atoi(e);
    if (v == NR || v == NR / 2) NC = v;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap gmap{};
  if (layout == TSSEP_REC_LAYOUT_ROWS) {
    // G (rows, T, 2, 4, Up) bf16 viewed as (unit, gate, row, dir, t); box = 64 units x 4 gates x NR rows
    EncodeTiledFn enc = get_encode_tiled();
    TSSEP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t up = static_cast<cuuint64_t>(Up);
    cuuint64_t dims[5] = {up, 4, static_cast<cuuint64_t>(rows), 2, static_cast<cuuint64_t>(T)};
    cuuint64_t strides[4] = {up * 2, static_cast<cuuint64_t>(T) * 8 * up * 2, 4 * up * 2, 8 * up * 2};
    cuuint32_t box[5] = {64, 4, static_cast<cuuint32_t>(NR), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(G), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TSSEP_REQUIRE(r == CUDA_SUCCESS, "tssep_blstm_recurrence_ts: cuTensorMapEncodeTiled failed with code %d", static_cast<int>(r));
  }
#define TSSEP_TS_CASE(NR_, NC_)                                                                                   \
  if (NR == NR_ && NC == NC_) {                                                                                   \
    if (layout == TSSEP_REC_LAYOUT_ROWS)                                                                          \
      return a.fast ? launch_ts<NR_, NC_, true, true, true>(a, gmap, C, nsub, smem, st)                           \
                    : launch_ts<NR_, NC_, false, true, true>(a, gmap, C, nsub, smem, st);                         \
    if (a.fast) return g_dtype ? launch_ts<NR_, NC_, true, true, false>(a, gmap, C, nsub, smem, st)               \
                               : launch_ts<NR_, NC_, true, false, false>(a, gmap, C, nsub, smem, st);             \
    return g_dtype ? launch_ts<NR_, NC_, false, true, false>(a, gmap, C, nsub, smem, st)                          \
                   : launch_ts<NR_, NC_, false, false, false>(a, gmap, C, nsub, smem, st);                        \
  }
  TSSEP_TS_CASE(16, 8)
  TSSEP_TS_CASE(16, 16)
  TSSEP_TS_CASE(32, 16)
  TSSEP_TS_CASE(32, 32)
#undef TSSEP_TS_CASE
  set_error("tssep_blstm_recurrence_ts: no instantiation for %d rows per cluster, %d columns per warp", NR, NC);
  return -1;
}

}  // extern "C"
