// BLSTM recurrence, first throughput variant: recurrent weights resident in SHARED MEMORY, the
// per-step contraction on tcgen05 tensor cores with TMEM accumulators.
//
// SUPERSEDED by csrc/lstm_ts.cu (weights in tensor memory, convergent MMA issue: 1.3 us/step instead of 3.4 for
// 208 rows); kept selectable (TSSEP_LSTM_KERNEL=tc) as the A/B baseline the profiles refer to.  Its issuer still
// sits in a divergent lane-0 branch, which costs ~70 cycles per MMA (see lstm_ts.cu).
//
// Same operator as csrc/lstm.cu (torch.nn.LSTM time loop, tssep/train/rnnp.py:87-95, :143-159)
// for many batch rows at once.  One cluster of C = ceil(Up/64) CTAs per (32 batch rows,
// direction).  Each CTA owns 64 hidden units = 256 gate rows (row = 4 * unit + gate), kept in
// shared memory for the whole sequence as two 128-row UMMA A operands (K-major, 128-byte
// swizzle).  Per step:
//   warp 1 (one thread)  waits for h_{t-1} (32 rows x Up, bf16, the UMMA B operand, written
//                        by every CTA of the cluster through DSMEM), issues
//                        tcgen05.mma M=128 N=32 K=16 over all of K for both row tiles and
//                        commits each tile to an mbarrier;
//   warps 2-5            (lane = gate row) pull the 128 x 32 fp32 pre-activations out of TMEM
//                        with tcgen05.ld, add the input projection (prefetched one step ahead
//                        straight from global memory), apply the gate non-linearities, transpose
//                        4x4 blocks inside lane quads so that one lane owns all four gates of a
//                        (unit, batch row) cell, update c_t (registers) and h_t, regroup 8 units
//                        into 16-byte chunks and push them to every CTA's next B operand with
//                        st.async (complete_tx on the destination mbarrier) and to H in HBM.
// The epilogue of row tile 0 overlaps the MMAs of row tile 1.
#include "../../include/tssep_b200.h"
#include "common.cuh"

#include <cstdlib>

namespace tssep {

constexpr int kTcN = 32;          // batch rows per cluster
constexpr int kTcThreads = 320;   // warp 0: spare, warp 1: MMA issuer, warps 2-5 / 6-9: epilogue of row tile 0 / 1
constexpr int kAtomA = 128 * 128; // bytes of one 128-row x 64-k swizzle atom
constexpr int kAtomB = kTcN * 128;

struct RecTcArgs {
  const void* G;         // tiles [group][t][dir][unit octet][b/4][4*(unit%8)+gate][b%4], f32 or bf16 (GEMM BT modes)
  const uint4* Wimg;     // [dir][cta][tile][atom] pre-swizzled 16 KB blocks
  __nv_bfloat16* H;      // (groups, T, 32, 2*Up): rows ordered (group, t, b)
  int rows, T, Up, NA, fast, g_bf16;
  int* prof;  // optional: per-phase cycle counters of one epilogue warp (debugging aid)
};

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, const uint4& v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   remote_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kTcThreads, 1) blstm_rec_tc_kernel(const RecTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NA = a.NA;                                  // k-atoms of 64
  const uint32_t sA = base;                             // [2 tiles][NA] x 16 KB
  const uint32_t sB = sA + 2u * NA * kAtomA;            // [2 buffers][NA] x 4 KB
  const uint32_t sT = sB + 2u * NA * kAtomB;            // [8 warps] x 512 B transposition tiles
  const uint32_t sBar = sT + 8 * 512;
  const uint32_t hfull0 = sBar, accfull0 = sBar + 16, accempty0 = sBar + 32, tptr = sBar + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t C = cluster_nctarank();
  const int bg = blockIdx.y, dir = blockIdx.z;
  const int T = a.T, Up = a.Up;
  const uint32_t tx_bytes = static_cast<uint32_t>(kTcN) * 64u * 2u * C;  // every CTA ships 64 units x 32 rows

  // ---- one-time setup: weights -> smem, zero h buffers, barriers, TMEM -------------------------
  {
    const uint4* src = a.Wimg + (static_cast<size_t>(dir) * C + crank) * (2u * NA * kAtomA / 16);
    const uint32_t n16 = 2u * NA * kAtomA / 16;
    for (uint32_t i = threadIdx.x; i < n16; i += kTcThreads) {
      const uint4 v = __ldg(src + i);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + 16 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
    for (uint32_t i = threadIdx.x; i < 2u * NA * kAtomB / 16; i += kTcThreads)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sB + 16 * i), "r"(0u) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(hfull0, 1);
      mbar_init(hfull0 + 8, 1);
      mbar_init(accfull0, 1);
      mbar_init(accfull0 + 8, 1);
      mbar_init(accempty0, 4);
      mbar_init(accempty0 + 8, 4);
      mbar_fence_init();
      mbar_arrive_expect_tx(hfull0, tx_bytes);
      mbar_arrive_expect_tx(hfull0 + 8, tx_bytes);
    }
    __syncwarp();
    tc_alloc(tptr, 64);
    tc_relinquish();
  }
  fence_proxy_async();  // generic-proxy writes of the weights / zeros -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  cluster_sync_all();

  const int ksteps = Up / 16;

  if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTcN >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);
      for (int s = 0; s < T; ++s) {
        const int rb = (s & 1) ^ 1;
        if (s > 0) {
          mbar_wait(hfull0 + 8 * rb, ((s - 1) >> 1) & 1);
          mbar_arrive_expect_tx(hfull0 + 8 * rb, tx_bytes);  // re-arm for the data of step s+1
          fence_proxy_async();  // h arrived through st.async (generic proxy); the MMA reads via the async proxy
        }
        const uint32_t b0 = sB + static_cast<uint32_t>(rb) * NA * kAtomB;
#pragma unroll 1
        for (int tile = 0; tile < 2; ++tile) {
          mbar_wait(accempty0 + 8 * tile, (s & 1) ^ 1);  // epilogue of step s-1 drained this accumulator
          tc_fence_after();
          const uint32_t a0 = sA + static_cast<uint32_t>(tile) * NA * kAtomA;
          const uint32_t d = tmem_base + tile * kTcN;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t atom = k >> 2, k4 = k & 3;
            tc_mma_bf16(d, make_desc_sw128(a0 + atom * kAtomA + k4 * 32), make_desc_sw128(b0 + atom * kAtomB + k4 * 32),
                        idesc, k != 0 ? 1u : 0u);
          }
          tc_commit(accfull0 + 8 * tile);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ---- epilogue: gates, cell update, h exchange (one row tile per warp set) ----------------------
    const int tl = (warp - 2) >> 2;    // row tile handled by this warp
    const int q = warp & 3;            // TMEM lane quarter
    const int gate = lane & 3;         // i, f, g, o
    const int ul = lane >> 2;          // unit within the warp's octet
    const bool is_g = gate == 2;
    const float sc = a.fast ? (is_g ? 1.0f : 0.5f) : (is_g ? 2.0f : 1.0f);
    const float ka = a.fast ? (is_g ? 1.0f : 0.5f) : (is_g ? 2.0f : 1.0f);
    const float kb = a.fast ? (is_g ? 0.0f : 0.5f) : (is_g ? -1.0f : 0.0f);
    const uint32_t myT = sT + static_cast<uint32_t>(warp - 2) * 512;
    // destination of this lane's 16-byte chunk (batch row = lane) inside a B buffer
    const uint32_t chunk_row = static_cast<uint32_t>(lane >> 3) * 1024 + static_cast<uint32_t>(lane & 7) * 128;
    uint32_t r_b[8], r_bar[8];
#pragma unroll
    for (int dcta = 0; dcta < 8; ++dcta) {
      r_b[dcta] = static_cast<uint32_t>(dcta) < C ? mapa(sB, dcta) : 0;
      r_bar[dcta] = static_cast<uint32_t>(dcta) < C ? mapa(hfull0, dcta) : 0;
    }
    const int unit = static_cast<int>(crank) * 64 + tl * 32 + q * 8 + ul;
    const int unit0 = static_cast<int>(crank) * 64 + tl * 32 + q * 8;
    const bool unit_ok = unit < Up;
    // G tile of this warp's unit octet: [b/4][lane][b%4] -> instruction i reads 512 contiguous bytes
    const int octet = static_cast<int>(crank) * 8 + tl * 4 + q;
    const int64_t g_tile = (((static_cast<int64_t>(bg) * T) * 2 + dir) * (static_cast<int64_t>(Up) / 8) + octet);
    const int64_t g_tstride = 2 * (static_cast<int64_t>(Up) / 8);  // tiles per time step
    const float4* gbase32 = reinterpret_cast<const float4*>(a.G) + g_tile * 256 + lane;  // 1024 floats per tile
    const uint2* gbase16 = reinterpret_cast<const uint2*>(a.G) + g_tile * 256 + lane;    // 1024 bf16 per tile
    __nv_bfloat16* hbase = a.H + ((static_cast<int64_t>(bg) * T) * kTcN + lane) * (2 * static_cast<int64_t>(Up)) + dir * Up + unit0;
    const int64_t h_tstride = static_cast<int64_t>(kTcN) * 2 * Up;

    float cst[8];
    uint4 gcur[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cst[i] = 0.f;
    // raw loads only: the values are converted where they are consumed (one step later), otherwise the
    // in-order issue would stall on every load
    auto load_g = [&](int s, uint4* g) {
      const int t = dir ? T - 1 - s : s;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (!unit_ok) {
          g[i] = make_uint4(0u, 0u, 0u, 0u);
        } else if (a.g_bf16) {
          asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0, %1}, [%2];"
                       : "=r"(g[i].x), "=r"(g[i].y)
                       : "l"(gbase16 + static_cast<int64_t>(t) * g_tstride * 256 + i * 32));
        } else {
          asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(g[i].x), "=r"(g[i].y), "=r"(g[i].z), "=r"(g[i].w)
                       : "l"(gbase32 + static_cast<int64_t>(t) * g_tstride * 256 + i * 32));
        }
      }
    };
    load_g(0, gcur);

    const bool do_prof = a.prof != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && crank == 0 && (warp == 2 || warp == 6);
    int pc[4] = {0, 0, 0, 0};
    for (int s = 0; s < T; ++s) {
      const int t = dir ? T - 1 - s : s;
      const int wb = s & 1;
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (do_prof) c0 = clock();
      mbar_wait(accfull0 + 8 * tl, s & 1);
      if (do_prof) c1 = clock();
      tc_fence_after();
      uint32_t v[kTcN];
      tc_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tl * kTcN, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty0 + 8 * tl);

      // gate non-linearity of this lane's row for all 32 batch rows
      float act[kTcN];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gv[4];
        if (a.g_bf16) {
          gv[0] = __uint_as_float(gcur[i].x << 16);
          gv[1] = __uint_as_float(gcur[i].x & 0xffff0000u);
          gv[2] = __uint_as_float(gcur[i].y << 16);
          gv[3] = __uint_as_float(gcur[i].y & 0xffff0000u);
        } else {
          gv[0] = __uint_as_float(gcur[i].x);
          gv[1] = __uint_as_float(gcur[i].y);
          gv[2] = __uint_as_float(gcur[i].z);
          gv[3] = __uint_as_float(gcur[i].w);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float x = (__uint_as_float(v[4 * i + j]) + gv[j]) * sc;
          const float y = a.fast ? tanh_fast(x) : sigmoid_acc(x);
          act[4 * i + j] = fmaf(y, ka, kb);
        }
      }
      if (do_prof) c2 = clock() + (__float_as_int(act[0] + act[31]) & 0);
      // prefetch the next step's input projection while the cell math runs
      if (s + 1 < T) load_g(s + 1, gcur);

      // 4x4 transposes inside lane quads: afterwards act[4i + g] = gate g of batch row 4i + (lane & 3)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
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
      // cell update for the 8 batch rows 4i + gate this lane now owns; stage h as bf16 in T[b][ul]
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ig = act[4 * i + 0], fg = act[4 * i + 1], gg = act[4 * i + 2], og = act[4 * i + 3];
        const float c = fmaf(fg, cst[i], ig * gg);
        cst[i] = c;
        const float h = og * (a.fast ? tanh_fast(c) : tanh_acc(c));
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
                   : "r"(myT + static_cast<uint32_t>(lane) * 16));
      __syncwarp();
      // this CTA's 64 units are k-atom `crank`; the warp's octet is 16-byte chunk tl*4 + q of the row
      const uint32_t cidx = static_cast<uint32_t>(tl * 4 + q);
      const uint32_t off = static_cast<uint32_t>(wb) * NA * kAtomB + crank * kAtomB + chunk_row +
                           ((cidx ^ static_cast<uint32_t>(lane & 7)) << 4);
      if (s + 1 < T) {
#pragma unroll
        for (int dcta = 0; dcta < 8; ++dcta)
          if (static_cast<uint32_t>(dcta) < C) st_async_v4(r_b[dcta] + off, chunk, r_bar[dcta] + 8 * wb);
      }
      if (unit0 < Up) *reinterpret_cast<uint4*>(hbase + static_cast<int64_t>(t) * h_tstride) = chunk;
      if (do_prof) {
        const int c4 = clock();
        pc[0] += c1 - c0;  // wait for the accumulator
        pc[1] += c2 - c1;  // tcgen05.ld + activations
        pc[2] += c3 - c2;  // G prefetch issue + transposes + cell update
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
    tc_dealloc(tmem_base, 64);
  }
  cluster_sync_all();
}

// weight_hh (4U, U) f32 -> the byte image the kernel copies into shared memory
__global__ void pack_whh_tc_kernel(const float* __restrict__ w_fwd, const float* __restrict__ w_bwd, int U, int Up,
                                   int C, int NA, __nv_bfloat16* __restrict__ out) {
  const int64_t per_cta = 2ll * NA * (kAtomA / 2);  // bf16 elements
  const int64_t total = 2ll * C * per_cta;
  for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = o;
    const int e = static_cast<int>(r % 8);
    r /= 8;
    const int chunk_sw = static_cast<int>(r % 8);
    r /= 8;
    const int m = static_cast<int>(r % 128);
    r /= 128;
    const int atom = static_cast<int>(r % NA);
    r /= NA;
    const int tile = static_cast<int>(r % 2);
    r /= 2;
    const int cta = static_cast<int>(r % C);
    const int dir = static_cast<int>(r / C);
    // address inside the atom: (m/8)*1024 + (m%8)*128 + (chunk ^ (m%8))*16 + e*2  ->  invert the swizzle
    const int chunk = chunk_sw ^ (m & 7);
    const int k = atom * 64 + chunk * 8 + e;
    const int unit = cta * 64 + tile * 32 + (m >> 2), gate = m & 3;
    const float* w = dir ? w_bwd : w_fwd;
    const float v = (unit < U && k < U) ? w[(static_cast<size_t>(gate) * U + unit) * U + k] : 0.f;
    out[o] = __float2bfloat16_rn(v);
  }
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_pack_whh_tc(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint16_t* Wimg, tssep_stream_t stream) {
  TSSEP_REQUIRE(whh_fwd && whh_bwd && Wimg, "tssep_pack_whh_tc: null pointer");
  TSSEP_REQUIRE(U >= 1 && Up >= U && Up % 16 == 0 && Up <= 512, "tssep_pack_whh_tc: need U <= Up, Up %% 16 == 0, Up <= 512");
  const int C = (Up + 63) / 64, NA = (Up + 63) / 64;
  const int64_t total = 2ll * C * 2 * NA * (kAtomA / 2);
  const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 32));
  pack_whh_tc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(whh_fwd, whh_bwd, U, Up, C, NA,
                                                                            reinterpret_cast<__nv_bfloat16*>(Wimg));
  return check_launch("tssep_pack_whh_tc");
}

int tssep_blstm_recurrence_tc(const void* G, int g_dtype, const uint16_t* Wimg, uint16_t* H, int64_t rows, int64_t T,
                              int Up, int fast_math, tssep_stream_t stream) {
  TSSEP_REQUIRE(G && Wimg && H, "tssep_blstm_recurrence_tc: null pointer");
  TSSEP_REQUIRE(g_dtype == 0 || g_dtype == 1, "tssep_blstm_recurrence_tc: g_dtype must be 0 (f32) or 1 (bf16)");
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 512, "tssep_blstm_recurrence_tc: Up must be a multiple of 16 in [16, 512]");
  TSSEP_REQUIRE(rows >= 0 && T >= 0 && T < (1ll << 31) && (rows + kTcN - 1) / kTcN <= 65535,
                "tssep_blstm_recurrence_tc: bad extent");
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(H) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wimg) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                "tssep_blstm_recurrence_tc: G, H and Wimg must be 16-byte aligned");
  if (rows == 0 || T == 0) return 0;
  const int C = (Up + 63) / 64, NA = C;
  RecTcArgs a;
  a.G = G;
  a.Wimg = reinterpret_cast<const uint4*>(Wimg);
  a.H = reinterpret_cast<__nv_bfloat16*>(H);
  a.rows = static_cast<int>(rows);
  a.T = static_cast<int>(T);
  a.Up = Up;
  a.NA = NA;
  a.fast = fast_math & 1;
  a.g_bf16 = g_dtype;
  a.prof = nullptr;
  if (const char* e = getenv("TSSEP_REC_PROF")) a.prof = reinterpret_cast<int*>(strtoull(e, nullptr, 0));
  const size_t smem = 1024 + 2ull * NA * kAtomA + 2ull * NA * kAtomB + 8 * 512 + 128;
  TSSEP_REQUIRE(smem <= 227 * 1024, "tssep_blstm_recurrence_tc: Up=%d needs %zu bytes of shared memory", Up, smem);
  TSSEP_CUDA(cudaFuncSetAttribute(blstm_rec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, static_cast<unsigned>((rows + kTcN - 1) / kTcN), 2);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TSSEP_CUDA(cudaLaunchKernelEx(&cfg, blstm_rec_tc_kernel, a));
  return check_launch("blstm_rec_tc");
}

}  // extern "C"
