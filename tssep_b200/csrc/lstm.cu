// Persistent BLSTM recurrence for the RNNP stack.
//
// Replaces the time loop inside torch.nn.LSTM (tssep/train/rnnp.py:87-95,
// called at rnnp.py:143-159): gate order i,f,g,o, zero initial state,
// c_t = s(f) c_{t-1} + s(i) tanh(g), h_t = s(o) tanh(c_t), both directions.
//
// Layout of one launch
//   grid = (C, ceil(rows/8), 2): one cluster of C CTAs per (batch tile of 8
//   rows, direction); every (tile, direction) recurrence is independent, so
//   all of them run concurrently and the forward and backward passes overlap.
//   Inside a cluster the 4*Up gate rows are split into "unit tiles" of 4
//   hidden units x 4 gates (= one m16 MMA row block).  Each compute warp owns
//   one unit tile and keeps its 16 x Up slice of W_hh in REGISTERS as
//   mma.m16n8k16 A fragments for the whole sequence; per step it multiplies
//   that slice with h_{t-1} (8 batch columns, bf16, shared memory), adds the
//   pre-computed input projection, applies the gates in fp32 (c_t stays in
//   registers) and pushes its 4 x 8 new h values to every CTA of the cluster
//   with st.async through distributed shared memory; the bytes complete an
//   mbarrier in the destination CTA, so one step costs one DSMEM hop and no
//   cluster-wide barrier.  A producer warp streams the input projections
//   G[b, t, dir, gate, unit] with 5-D TMA boxes through a 4-deep ring.
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

constexpr int kGStages = 4;
constexpr int kMaxWarpsCompute = 12;

__device__ __forceinline__ void mma_bf16_16816(float* d, const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}

// compute warps per CTA: large hidden sizes hold many weight fragments per thread and are limited to
// 10 warps (+1 producer) so that ptxas may use up to 184 registers
constexpr int max_compute_warps(int kt) { return kt >= 13 ? 10 : kMaxWarpsCompute; }

template <int KT, int NB>
__global__ void __launch_bounds__(32 * (max_compute_warps(KT) + 1), 1)
blstm_rec_kernel(const __grid_constant__ CUtensorMap gmap, const uint4* __restrict__ Wfrag,
                 __nv_bfloat16* __restrict__ H, int rows, int T, int NT, int fast, int g_bf16,
                 int* __restrict__ prof) {
  constexpr int Up = 16 * KT;
  constexpr int LDH = Up + 8;          // bf16 elements per h row (+8 keeps ldmatrix conflict free)
  constexpr int n_tiles = Up / 4;      // unit tiles per direction
  constexpr int BR = 8 * NB;           // batch rows per cluster
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const int ld_u = 4 * ((NT + 1) & ~1);  // units per staged row (even tile count keeps bf16 rows 16-byte multiples)
  const uint32_t g_elt = g_bf16 ? 2u : 4u;
  const uint32_t g_stage_bytes = static_cast<uint32_t>(4 * BR * ld_u) * g_elt;  // [gate][b][unit] f32 or bf16
  const uint32_t s_gring = sbase;                                             // kGStages stages
  const uint32_t s_hbuf = s_gring + kGStages * g_stage_bytes;                 // [NB][2][8][LDH] bf16
  constexpr uint32_t h_buf_bytes = 8 * LDH * 2;
  const uint32_t s_bar = s_hbuf + NB * 2 * h_buf_bytes;
  const uint32_t hfull0 = s_bar;                    // [NB][2]
  const uint32_t gfull0 = hfull0 + 16 * NB;         // [kGStages]
  const uint32_t gempty0 = gfull0 + 8 * kGStages;   // [kGStages]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t C = cluster_nctarank();
  const int bt = blockIdx.y, dir = blockIdx.z;
  int nvalid = n_tiles - static_cast<int>(crank) * NT;
  nvalid = nvalid < 0 ? 0 : (nvalid > NT ? NT : nvalid);
  const uint32_t tx_bytes = n_tiles * 64u;  // all unit tiles x (8 rows x 4 units x bf16)

  // zero all h buffers (h_{-1} = 0, padded units stay 0)
  for (uint32_t i = threadIdx.x; i < NB * 2 * h_buf_bytes / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_hbuf + 4 * i), "r"(0u) : "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * NB; ++i) mbar_init(hfull0 + 8 * i, 1);
    for (int i = 0; i < kGStages; ++i) {
      mbar_init(gfull0 + 8 * i, 1);
      mbar_init(gempty0 + 8 * i, nvalid > 0 ? nvalid : 1);
    }
    mbar_fence_init();
    for (int i = 0; i < 2 * NB; ++i) mbar_arrive_expect_tx(hfull0 + 8 * i, tx_bytes);
    tma_prefetch_desc(&gmap);
  }
  __syncthreads();
  cluster_sync_all();

  if (warp == NT) {
    // ---- producer: stream G[b-tile, t, dir, :, unit slice] ----------------
    if (lane == 0 && nvalid > 0) {
      for (int s = 0; s < T; ++s) {
        const int gs = s % kGStages;
        const uint32_t gph = (s / kGStages) & 1;
        const int t = dir ? T - 1 - s : s;
        mbar_wait(gempty0 + 8 * gs, gph ^ 1);
        mbar_arrive_expect_tx(gfull0 + 8 * gs, g_stage_bytes);
        tma_load_5d(s_gring + gs * g_stage_bytes, &gmap, gfull0 + 8 * gs, static_cast<int>(crank) * 4 * NT, bt * BR, 0,
                    dir, t);
      }
    }
    __syncwarp();
  } else if (warp < nvalid) {
    // ---- compute warp: one unit tile ---------------------------------------
    const int gt = static_cast<int>(crank) * NT + warp;
    uint4 a[KT];
    {
      const uint4* wp = Wfrag + (static_cast<size_t>(dir) * n_tiles + gt) * KT * 32 + lane;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) a[kt] = __ldg(wp + kt * 32);
    }
    const bool upper = lane >= 16;
    const int u = (lane >> 2) & 3;
    const int n0 = 2 * (lane & 3);
    // G offsets (floats) inside a ring stage: [gate][BR rows][unit]
    const int gA = upper ? 1 : 0, gB = upper ? 3 : 2;
    const int unit_local = warp * 4 + u;
    const int offA = (gA * BR + n0) * ld_u + unit_local;
    const int offB = (gB * BR + n0) * ld_u + unit_local;
    // ldmatrix row address: matrix (lane>>3) covers k offset 8*(lane>>3), row (lane&7)
    const uint32_t ldm_off = static_cast<uint32_t>(((lane & 7) * LDH + (lane >> 3) * 8) * 2);
    // transposed send: this lane ships batch row nn, 4 units, to CTAs dg, dg+4
    const int nn = lane & 7, dg = lane >> 3;
    const int src_base = (nn & 1) * 16 + (nn >> 1);
    const uint32_t send_off = static_cast<uint32_t>((nn * LDH + gt * 4) * 2);
    uint32_t r_hbuf[2], r_bar[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t dst = dg + 4 * j;
      r_hbuf[j] = dst < C ? mapa(s_hbuf, dst) : 0;
      r_bar[j] = dst < C ? mapa(hfull0, dst) : 0;
    }
    const int64_t brow0 = static_cast<int64_t>(bt) * BR + nn;
    __nv_bfloat16* hout = H + (brow0 * T) * (2 * Up) + dir * Up + gt * 4;
    const float kB = upper ? 1.0f : 2.0f;

    float c_state[NB];
    float g_cur[NB][4], g_nxt[NB][4];
#pragma unroll
    for (int b = 0; b < NB; ++b) c_state[b] = 0.f;
    auto load_g = [&](int s, float (*g)[4]) {
      const int gs = s % kGStages;
      mbar_wait(gfull0 + 8 * gs, (s / kGStages) & 1);
      const uint32_t gp = s_gring + gs * g_stage_bytes;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int o[4] = {offA + b * 8 * ld_u, offA + (b * 8 + 1) * ld_u, offB + b * 8 * ld_u, offB + (b * 8 + 1) * ld_u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (g_bf16) {
            unsigned short h;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(gp + 2 * o[j]));
            g[b][j] = __uint_as_float(static_cast<uint32_t>(h) << 16);
          } else {
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g[b][j]) : "r"(gp + 4 * o[j]));
          }
        }
      }
    };
    load_g(0, g_cur);

    const bool do_prof = prof != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && crank == 0 && warp == 0;
    int pc[6] = {0, 0, 0, 0, 0, 0};
    for (int s = 0; s < T; ++s) {
      const int t = dir ? T - 1 - s : s;
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
      if (do_prof) c0 = clock();
      if (do_prof) c1 = clock();
      const int rb = (s & 1) ^ 1;
      // The batch tiles are independent recurrences: while the new h of one tile travels through
      // DSMEM, the warp already works on the other tile.
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (do_prof) c1 = clock();
        // h_{t-1} of batch tile b from every CTA of the cluster.  st.async completes bytes on OUR
        // mbarrier after the data landed in OUR shared memory, so the default (CTA-scope) wait
        // suffices (as for multicast TMA); a cluster-scope acquire costs a CCTL.IVALL per step.
        const uint32_t hbar = hfull0 + 8 * (b * 2 + rb);
        if (s > 0) {
          mbar_wait(hbar, ((s - 1) >> 1) & 1);
          if (warp == 0 && lane == 0) mbar_arrive_expect_tx(hbar, tx_bytes);
        }
        if (do_prof) c2 = clock();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const uint32_t hb = s_hbuf + (b * 2 + rb) * h_buf_bytes + ldm_off;
#pragma unroll
        for (int kt = 0; kt < KT; kt += 2) {
          if (kt + 1 < KT) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(hb + kt * 32, b0, b1, b2, b3);
            mma_bf16_16816(acc[kt & 3], a[kt], b0, b1);
            mma_bf16_16816(acc[(kt + 1) & 3], a[kt + 1], b2, b3);
          } else {
            uint32_t b0, b1;
            ldmatrix_x2(hb + kt * 32, b0, b1);
            mma_bf16_16816(acc[kt & 3], a[kt], b0, b1);
          }
        }
        // prefetch the next step's input projection into registers while the tensor pipe drains
        if (b == 0 && s + 1 < T) load_g(s + 1, g_nxt);
        const float pA0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]) + g_cur[b][0];  // row r,   col n0
        const float pA1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]) + g_cur[b][1];  // row r,   col n0+1
        const float pB0 = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]) + g_cur[b][2];  // row r+8, col n0
        const float pB1 = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]) + g_cur[b][3];  // row r+8, col n0+1
        if (do_prof) c3 = clock() + (__float_as_int(pA0 + pA1 + pB0 + pB1) & 0);

        // lower half-warp holds (i, g), upper half-warp holds (f, o) of unit u, columns n0, n0+1
        float sA0, sA1, sB0, sB1;
        if (fast & 1) {
          sA0 = fmaf(0.5f, tanh_fast(0.5f * pA0), 0.5f);
          sA1 = fmaf(0.5f, tanh_fast(0.5f * pA1), 0.5f);
          sB0 = upper ? fmaf(0.5f, tanh_fast(0.5f * pB0), 0.5f) : tanh_fast(pB0);
          sB1 = upper ? fmaf(0.5f, tanh_fast(0.5f * pB1), 0.5f) : tanh_fast(pB1);
        } else {
          sA0 = sigmoid_acc(pA0);
          sA1 = sigmoid_acc(pA1);
          sB0 = sigmoid_acc(kB * pB0);
          sB1 = sigmoid_acc(kB * pB1);
          if (!upper) {
            sB0 = fmaf(2.0f, sB0, -1.0f);  // tanh(g)
            sB1 = fmaf(2.0f, sB1, -1.0f);
          }
        }
        // lower owns column n0, upper owns column n0+1
        const float ig0 = sA0 * sB0, ig1 = sA1 * sB1;  // meaningful on lower lanes only
        const float x1 = __shfl_xor_sync(0xffffffffu, upper ? sA0 : ig1, 16);  // lower gets s(f0), upper gets ig1
        const float x2 = __shfl_xor_sync(0xffffffffu, sB0, 16);                // lower gets s(o0)
        const float fgate = upper ? sA1 : x1;
        const float inew = upper ? x1 : ig0;
        const float ogate = upper ? sB1 : x2;
        c_state[b] = fmaf(fgate, c_state[b], inew);
        const float hval = ogate * ((fast & 1) ? tanh_fast(c_state[b]) : tanh_acc(c_state[b]));

        // transpose: lane (nn, dg) gathers units 0..3 of batch column nn
        const float v0 = __shfl_sync(0xffffffffu, hval, src_base + 0);
        const float v1 = __shfl_sync(0xffffffffu, hval, src_base + 4);
        const float v2 = __shfl_sync(0xffffffffu, hval, src_base + 8);
        const float v3 = __shfl_sync(0xffffffffu, hval, src_base + 12);
        const uint32_t lo = pack_bf16x2(v0, v1), hi = pack_bf16x2(v2, v3);
        if (do_prof) c4 = clock() + (lo & hi & 0);
        if (s + 1 < T) {
          const uint32_t wb_off = static_cast<uint32_t>(b * 2 + (s & 1)) * h_buf_bytes + send_off;
          const uint32_t bar_off = static_cast<uint32_t>(b * 2 + (s & 1)) * 8;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (dg + 4 * j < static_cast<int>(C)) st_async_v2(r_hbuf[j] + wb_off, lo, hi, r_bar[j] + bar_off);
        }
        if (dg == 0 && brow0 + b * 8 < rows)
          *reinterpret_cast<uint2*>(hout + (static_cast<int64_t>(b) * 8 * T + t) * (2 * Up)) = make_uint2(lo, hi);
        if (do_prof) {
          const int c5 = clock();
          pc[0] += (b == 0) ? c1 - c0 : 0;  // G prefetch
          pc[1] += c2 - c1;                 // h wait
          pc[2] += c3 - c2;                 // ldmatrix + mma
          pc[3] += c4 - c3;                 // gates + shuffles
          pc[4] += c5 - c4;                 // sends + store
          c0 = c1 = c5;
        }
      }
      // the values of stage s were consumed by the gates above: hand the slot back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(gempty0 + 8 * (s % kGStages));
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) g_cur[b][i] = g_nxt[b][i];
      if (do_prof) pc[5] += 1;
    }
    if (do_prof && lane == 0)
      for (int i = 0; i < 6; ++i) prof[i] = pc[i];
  }
  __syncthreads();
  cluster_sync_all();
}

__global__ void pack_whh_kernel(const float* __restrict__ w_fwd, const float* __restrict__ w_bwd, int U, int Up,
                                uint4* __restrict__ out) {
  const int KT = Up / 16, n_tiles = Up / 4;
  const int total = 2 * n_tiles * KT * 32;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < total; o += gridDim.x * blockDim.x) {
    const int lane = o & 31;
    int rest = o >> 5;
    const int kt = rest % KT;
    rest /= KT;
    const int gt = rest % n_tiles;
    const int dir = rest / n_tiles;
    const float* w = dir ? w_bwd : w_fwd;
    const int r0 = lane >> 2, c0 = kt * 16 + (lane & 3) * 2;
    auto get = [&](int r, int c) -> float {
      const int gate = r >> 2, unit = gt * 4 + (r & 3);
      return (unit < U && c < U) ? w[(static_cast<size_t>(gate) * U + unit) * U + c] : 0.f;
    };
    uint4 v;
    v.x = pack_bf16x2(get(r0, c0), get(r0, c0 + 1));
    v.y = pack_bf16x2(get(r0 + 8, c0), get(r0 + 8, c0 + 1));
    v.z = pack_bf16x2(get(r0, c0 + 8), get(r0, c0 + 9));
    v.w = pack_bf16x2(get(r0 + 8, c0 + 8), get(r0 + 8, c0 + 9));
    out[o] = v;
  }
}

template <int KT, int NB>
static int launch_rec(const CUtensorMap& gmap, const uint32_t* Wfrag, uint16_t* H, int64_t rows, int64_t T, int C,
                      int NT, int fast, int g_bf16, int* prof, cudaStream_t stream) {
  constexpr int Up = 16 * KT;
  const size_t smem = static_cast<size_t>(kGStages) * (4 * 8 * NB * 4 * ((NT + 1) & ~1)) * 4 + NB * 2 * 8 * (Up + 8) * 2 + 16 * NB +
                      16 * kGStages + 128;
  TSSEP_CUDA(cudaFuncSetAttribute(blstm_rec_kernel<KT, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, static_cast<unsigned>((rows + 8 * NB - 1) / (8 * NB)), 2);
  cfg.blockDim = dim3(32 * (NT + 1));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TSSEP_CUDA(cudaLaunchKernelEx(&cfg, blstm_rec_kernel<KT, NB>, gmap, reinterpret_cast<const uint4*>(Wfrag),
                                reinterpret_cast<__nv_bfloat16*>(H), static_cast<int>(rows), static_cast<int>(T), NT,
                                fast, g_bf16, prof));
  return check_launch("blstm_rec");
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_pack_whh(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* Wfrag, tssep_stream_t stream) {
  TSSEP_REQUIRE(whh_fwd && whh_bwd && Wfrag, "tssep_pack_whh: null pointer");
  TSSEP_REQUIRE(U >= 1 && Up >= U && Up % 16 == 0 && Up <= 320, "tssep_pack_whh: need U <= Up, Up %% 16 == 0, Up <= 320");
  const int total = 2 * (Up / 4) * (Up / 16) * 32;
  pack_whh_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(whh_fwd, whh_bwd, U, Up,
                                                                                       reinterpret_cast<uint4*>(Wfrag));
  return check_launch("tssep_pack_whh");
}

int tssep_blstm_recurrence(const void* G, int g_dtype, const uint32_t* Wfrag, uint16_t* H, int64_t rows, int64_t T, int Up,
                           int cluster, int fast_math, tssep_stream_t stream) {
  TSSEP_REQUIRE(G && Wfrag && H, "tssep_blstm_recurrence: null pointer");
  TSSEP_REQUIRE(g_dtype == 0 || g_dtype == 1, "tssep_blstm_recurrence: g_dtype must be 0 (f32) or 1 (bf16)");
  const int g_bf16 = g_dtype;
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 320, "tssep_blstm_recurrence: Up must be a multiple of 16 in [16, 320], got %d", Up);
  TSSEP_REQUIRE(rows >= 0 && T >= 0 && T < (1ll << 31) && (rows + 7) / 8 <= 65535, "tssep_blstm_recurrence: bad extent");
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(G) & 15) == 0 && (reinterpret_cast<uintptr_t>(H) & 7) == 0,
                "tssep_blstm_recurrence: G must be 16-byte and H 8-byte aligned");
  if (rows == 0 || T == 0) return 0;
  const int tiles = Up / 4;
  int C = cluster;
  if (C == 0) {
    C = 1;
    while (C < 8 && (tiles + C - 1) / C > 10) C *= 2;
  }
  TSSEP_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, "tssep_blstm_recurrence: cluster must be 0, 1, 2, 4 or 8");
  const int NT = (tiles + C - 1) / C;
  // One batch tile (8 rows) per cluster has the lowest step latency (1.2 us vs 1.7 us measured at U=300) but
  // only ~15 clusters of 8 CTAs are co-resident on 148 SMs; when the tiles of both directions do not fit in
  // one wave, two tiles are interleaved per cluster (one tile's DSMEM hop hides behind the other's math).
  int max_clusters = 148 / C;
  {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
      max_clusters = sms / C;
    if (C == 8) max_clusters -= 3;  // GPC boundaries: 15 clusters of 8 were measured co-resident on B200
  }
  const int btiles = static_cast<int>((rows + 7) / 8);
  int NB = (2 * btiles + max_clusters - 1) / max_clusters;  // smallest NB that fits both directions in one wave
  NB = NB < 1 ? 1 : (NB > 4 ? 4 : NB);
  if (const char* e = debug_env("TSSEP_LSTM_NB")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 4) NB = v;
  }

  // G viewed as (unit, b, gate, dir, t)
  EncodeTiledFn enc = get_encode_tiled();
  TSSEP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap gmap;
  const cuuint64_t up = static_cast<cuuint64_t>(Up);
  cuuint64_t dims[5] = {up, static_cast<cuuint64_t>(rows), 4, 2, static_cast<cuuint64_t>(T)};
  const cuuint64_t esz = g_bf16 ? 2 : 4;
  cuuint64_t strides[4] = {static_cast<cuuint64_t>(T) * 8 * up * esz, up * esz, 4 * up * esz, 8 * up * esz};
  cuuint32_t box[5] = {static_cast<cuuint32_t>(4 * ((NT + 1) & ~1)), static_cast<cuuint32_t>(8 * NB), 4, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&gmap, g_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                   const_cast<void*>(G), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSSEP_REQUIRE(r == CUDA_SUCCESS, "tssep_blstm_recurrence: cuTensorMapEncodeTiled failed with code %d", static_cast<int>(r));

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // debug builds only (-DTSSEP_DEBUG_KNOBS): TSSEP_REC_PROF=<device pointer to 6 ints> makes one warp record
  // per-phase cycle counts; the shipped library never takes an address from the environment
  int* prof = nullptr;
#ifdef TSSEP_DEBUG_KNOBS
  if (const char* e = debug_env("TSSEP_REC_PROF")) prof = reinterpret_cast<int*>(strtoull(e, nullptr, 0));
#endif
  switch (Up / 16) {
#define TSSEP_CASE(kt)                                                                                  \
  case kt:                                                                                              \
    return NB == 1   ? launch_rec<kt, 1>(gmap, Wfrag, H, rows, T, C, NT, fast_math, g_bf16, prof, s)        \
           : NB == 2 ? launch_rec<kt, 2>(gmap, Wfrag, H, rows, T, C, NT, fast_math, g_bf16, prof, s)        \
           : NB == 3 ? launch_rec<kt, 3>(gmap, Wfrag, H, rows, T, C, NT, fast_math, g_bf16, prof, s)        \
                     : launch_rec<kt, 4>(gmap, Wfrag, H, rows, T, C, NT, fast_math, g_bf16, prof, s);
    TSSEP_CASE(1) TSSEP_CASE(2) TSSEP_CASE(3) TSSEP_CASE(4) TSSEP_CASE(5) TSSEP_CASE(6) TSSEP_CASE(7) TSSEP_CASE(8)
    TSSEP_CASE(9) TSSEP_CASE(10) TSSEP_CASE(11) TSSEP_CASE(12) TSSEP_CASE(13) TSSEP_CASE(14) TSSEP_CASE(15)
    TSSEP_CASE(16) TSSEP_CASE(17) TSSEP_CASE(18) TSSEP_CASE(19) TSSEP_CASE(20)
#undef TSSEP_CASE
  }
  set_error("tssep_blstm_recurrence: unsupported Up=%d", Up);
  return -1;
}

}  // extern "C"
