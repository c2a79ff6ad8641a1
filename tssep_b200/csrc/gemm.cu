// Bulk contractions of the RNNP stack on 5th-gen tensor cores.
//
//   out[z] = act(alpha * A[z / a_div] . B[z]^T + bias[z])
//
// A (rows, K) and B (N, K) are bf16, K-major.  One persistent CTA per SM:
// warp 0 streams 128 x 64 (A) and BN x 64 (B) tiles with TMA (128-byte swizzle)
// through a 4-stage mbarrier ring, one elected thread of warp 1 issues
// tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM accumulator,
// warps 2-5 drain TMEM with tcgen05.ld and run the fused epilogue (bias, tanh,
// cast, or the un-permuting sigmoid head) while the next tile is multiplied.
//
// Replaces torch.nn.Linear / the input half of torch.nn.LSTM in the reference
// (tssep/train/rnnp.py:87-96, tssep/train/net.py:663-666) and the final
// rearrange + trial mean + un-permute + sigmoid (net.py:629-661, :928-986).
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kMaxStages = 4;
// Epilogue warps (template parameter EPIW of the kernel): sets of four, each set covers the 128 TMEM lanes, the sets
// split the 32-column chunks of a tile.  8 warps leave room for a 4-stage operand ring (what the K = 2560 projection
// wants: 1.38 vs 1.22 PFLOP/s); 16 warps (3 stages) are for the GEMMs whose tile time is epilogue time: the head
// (two f32 outputs + sigmoid per accumulator: 1.97 -> 3.49 TB/s of output), the K <= 384 input projections (+7 %) and the
// narrow output projections (N <= 512, +5 %).
constexpr int gemm_threads(int epiw) { return 64 + 32 * epiw; }
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;

enum EpiMode { EPI_F32 = 0, EPI_BF16 = 1, EPI_HEAD = 2 };

struct GemmArgs {
  int64_t M;
  int N, K, batch, a_div, b_mod;
  int64_t a_row_stride;  // rows of A per (z / a_div)
  int64_t b_row_stride;  // rows of B per (z % b_mod)
  const float* bias;
  int64_t bias_stride;
  float alpha;
  int act;
  int mode;
  void* out;
  int64_t ldo, out_stride, out_stride_hi;
  int out_div;
  // head
  float* mask;
  const int* plane_map;
  int n_blocks, row_len;
  // tiling
  int bn, m_tiles, n_tiles, k_blocks, stages;
  int64_t total_tiles;
  // simt only
  const __nv_bfloat16* A;
  const __nv_bfloat16* B;
  int64_t lda, ldb;
};

__device__ __forceinline__ float apply_act(float v, int act) { return act == 1 ? tanh_acc(v) : v; }

// Stores up to 32 consecutive columns [n0, n0+32) of row m of batch z.
__device__ __forceinline__ void epilogue_store(const GemmArgs& g, int z, int64_t m, int n0, int limit, const float* acc) {
  const int nvalid = min(limit, g.N - n0);
  if (nvalid <= 0) return;
  const float* bias = g.bias ? g.bias + static_cast<int64_t>(z % g.b_mod) * g.bias_stride + n0 : nullptr;
  const int64_t zoff = static_cast<int64_t>(z / g.out_div) * g.out_stride_hi + static_cast<int64_t>(z % g.out_div) * g.out_stride;
  if (g.mode == EPI_F32) {
    float* o = static_cast<float*>(g.out) + zoff + m * g.ldo + n0;
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 v;
        v.x = apply_act(g.alpha * acc[i + 0] + (bias ? bias[i + 0] : 0.f), g.act);
        v.y = apply_act(g.alpha * acc[i + 1] + (bias ? bias[i + 1] : 0.f), g.act);
        v.z = apply_act(g.alpha * acc[i + 2] + (bias ? bias[i + 2] : 0.f), g.act);
        v.w = apply_act(g.alpha * acc[i + 3] + (bias ? bias[i + 3] : 0.f), g.act);
        *reinterpret_cast<float4*>(o + i) = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) o[i] = apply_act(g.alpha * acc[i] + (bias ? bias[i] : 0.f), g.act);
    }
  } else if (g.mode == EPI_BF16) {
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(g.out) + zoff + m * g.ldo + n0;
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v;
        uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = apply_act(g.alpha * acc[i + 2 * j] + (bias ? bias[i + 2 * j] : 0.f), g.act);
          const float b = apply_act(g.alpha * acc[i + 2 * j + 1] + (bias ? bias[i + 2 * j + 1] : 0.f), g.act);
          pv[j] = pack_bf16x2(a, b);
        }
        *reinterpret_cast<uint4*>(o + i) = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) o[i] = __float2bfloat16_rn(apply_act(g.alpha * acc[i] + (bias ? bias[i] : 0.f), g.act));
    }
  } else {  // EPI_HEAD
    int q = n0 / g.row_len;
    int f = n0 - q * g.row_len;
    int64_t plane = g.plane_map[z * g.n_blocks + q];
    float* lo = static_cast<float*>(g.out);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i >= nvalid) break;
      const float v = g.alpha * acc[i] + (bias ? bias[i] : 0.f);
      const int64_t idx = (plane * g.M + m) * g.row_len + f;
      if (lo) lo[idx] = v;
      if (g.mask) g.mask[idx] = sigmoid_acc(v);
      if (++f == g.row_len) {
        f = 0;
        ++q;
        if (q < g.n_blocks) plane = g.plane_map[z * g.n_blocks + q];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Coalesced epilogue of the tensor-core kernel.  Each epilogue warp owns 32 tile rows (lane =
// row in TMEM) and receives 32 consecutive columns per tcgen05.ld.  Writing those straight to
// global memory gives 32 scattered 16-byte pieces per store instruction (~0.9 TB/s measured);
// instead the warp transposes the 32x32 block through a private shared-memory tile so that
// consecutive lanes hold consecutive columns and every store instruction writes whole 128-byte
// lines.
//
// The kernel is instantiated per (epilogue mode, activation): with both as run-time values the
// inner loops carried 57 branches and per-chunk integer divisions, the eight epilogue warps were
// busy 88 % of the time and the tensor pipe only 25 % (ncu, profiles/r1_ncu_gemm_b1_inbt.txt).
// Everything that depends only on the tile is computed once per tile (EpiTile).
// ---------------------------------------------------------------------------
constexpr int kStageLd = 36;  // words per staged row: 16-byte aligned rows, conflict-free STS.128 / LDS.128

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
  return ACT == 1 ? tanh_acc(v) : v;
}

// What a lane needs for the columns it writes in the next chunk; loaded one chunk ahead so that the
// global-load latency never sits on the epilogue's critical path.
struct ChunkBias {
  float v[4];
  int plane;  // EPI_HEAD: destination plane of the lane's column
};

template <int MODE>
__device__ __forceinline__ ChunkBias load_chunk_bias(const GemmArgs& g, const float* bias, const int* plane_row, int n0,
                                                     int lane) {
  ChunkBias b;
  b.v[0] = b.v[1] = b.v[2] = b.v[3] = 0.f;
  b.plane = 0;
  if (MODE == EPI_HEAD) {
    const int n = n0 + lane;
    if (n < g.N) {
      if (bias) b.v[0] = __ldg(bias + n);
      b.plane = __ldg(plane_row + n / g.row_len);
    }
  } else if (bias) {
    const int n = n0 + (lane & 7) * 4;
    if (n + 3 < g.N) {
      if ((reinterpret_cast<uintptr_t>(bias + n) & 15) == 0) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(bias + n));
        b.v[0] = t.x, b.v[1] = t.y, b.v[2] = t.z, b.v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) b.v[j] = __ldg(bias + n + j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) b.v[j] = __ldg(bias + n + j);
    }
  }
  return b;
}

// Per (tile, warp) invariants of the epilogue.
struct EpiTile {
  int64_t rows_left;   // rows of this warp's 32-row block that exist
  int64_t base;        // element offset of (row m0 + sub, column c4) in the output
  bool fast;           // vector-aligned full block: one LDS.128 + one vector store per pass
};

template <int MODE>
__device__ __forceinline__ EpiTile make_epi_tile(const GemmArgs& g, int z, int64_t m0, int lane) {
  EpiTile t;
  t.rows_left = g.M - m0;
  t.base = 0;
  t.fast = false;
  const int sub = lane >> 3, c4 = (lane & 7) * 4;
  if (MODE == EPI_F32 || MODE == EPI_BF16) {
    const int64_t zoff = static_cast<int64_t>(z / g.out_div) * g.out_stride_hi + static_cast<int64_t>(z % g.out_div) * g.out_stride;
    t.base = zoff + (m0 + sub) * g.ldo + c4;
    const int esz = MODE == EPI_F32 ? 4 : 2;
    const uintptr_t amask = MODE == EPI_F32 ? 15 : 7;
    t.fast = t.rows_left >= 32 && ((reinterpret_cast<uintptr_t>(g.out) | static_cast<uintptr_t>(zoff * esz) |
                                    static_cast<uintptr_t>(g.ldo * esz)) & amask) == 0;
  }
  return t;
}

__device__ __forceinline__ void stage_rows(uint32_t stage, int lane, const uint32_t* v) {
#pragma unroll
  for (int i = 0; i < 32; i += 4)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (lane * kStageLd + i) * 4), "r"(v[i]),
                 "r"(v[i + 1]), "r"(v[i + 2]), "r"(v[i + 3])
                 : "memory");
  __syncwarp();
}

// EPI_F32 / EPI_BF16: per pass 4 rows x (8 lanes x 4 columns)
template <int MODE, int ACT>
__device__ __forceinline__ void epilogue_rows(const GemmArgs& g, const EpiTile& t, int n0, int limit, const uint32_t* v,
                                              const ChunkBias& cb, uint32_t stage, int lane) {
  const int nvalid = min(limit, g.N - n0);
  if (nvalid <= 0) return;
  stage_rows(stage, lane, v);
  const int sub = lane >> 3, c4 = (lane & 7) * 4;
  const uint32_t lds0 = stage + (sub * kStageLd + c4) * 4;
  if (t.fast && nvalid == 32) {
    const int64_t step = 4 * g.ldo;
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      float4 x;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                   : "r"(lds0 + pass * 4 * kStageLd * 4));
      x.x = act_t<ACT>(fmaf(g.alpha, x.x, cb.v[0]));
      x.y = act_t<ACT>(fmaf(g.alpha, x.y, cb.v[1]));
      x.z = act_t<ACT>(fmaf(g.alpha, x.z, cb.v[2]));
      x.w = act_t<ACT>(fmaf(g.alpha, x.w, cb.v[3]));
      if (MODE == EPI_F32) {
        *reinterpret_cast<float4*>(static_cast<float*>(g.out) + t.base + pass * step + n0) = x;
      } else {
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(g.out) + t.base + pass * step + n0) =
            make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
      }
    }
  } else {
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int r = pass * 4 + sub;
      float x[4];
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3])
                   : "r"(stage + (r * kStageLd + c4) * 4));
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = act_t<ACT>(fmaf(g.alpha, x[j], cb.v[j]));
      const bool row_ok = r < t.rows_left;
      const int64_t off = t.base + pass * 4 * g.ldo + n0;
      if (row_ok && c4 < nvalid) {
        if (MODE == EPI_F32) {
          float* o = static_cast<float*>(g.out) + off;
          if (c4 + 4 <= nvalid && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c4 + j < nvalid) o[j] = x[j];
          }
        } else {
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(g.out) + off;
          if (c4 + 4 <= nvalid && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
            *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c4 + j < nvalid) o[j] = __float2bfloat16_rn(x[j]);
          }
        }
      }
    }
  }
  __syncwarp();  // the staging tile is reused by the next chunk
}

// EPI_HEAD: lane = column; block q, offset f and the destination plane are fixed per lane for the chunk
__device__ __forceinline__ void epilogue_head(const GemmArgs& g, const EpiTile& t, int64_t m0, int n0, int limit,
                                              const uint32_t* v, const ChunkBias& cb, uint32_t stage, int lane) {
  const int nvalid = min(limit, g.N - n0);
  if (nvalid <= 0) return;
  stage_rows(stage, lane, v);
  const int n = n0 + lane;
  const bool col_ok = lane < nvalid;
  const int q = n / g.row_len, f = n - q * g.row_len;
  const int64_t base = col_ok ? (static_cast<int64_t>(cb.plane) * g.M + m0) * g.row_len + f : 0;
  const int rmax = static_cast<int>(t.rows_left < 32 ? t.rows_left : 32);
  const uint32_t lds0 = stage + lane * 4;
  float* lo = static_cast<float*>(g.out);
  float* mk = g.mask;
  if (rmax == 32 && lo != nullptr && mk != nullptr) {
    // all 32 rows of the lane's column at once: 32 independent loads, then the stores walk two running pointers
    float val[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val[r]) : "r"(lds0 + r * kStageLd * 4));
    float* plo = lo + base;
    float* pmk = mk + base;
    const int64_t step = g.row_len;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float x = fmaf(g.alpha, val[r], cb.v[0]);
      if (col_ok) {
        *plo = x;
        *pmk = sigmoid_acc(x);
      }
      plo += step;
      pmk += step;
    }
  } else {
#pragma unroll 1
    for (int r = 0; r < rmax; ++r) {
      float val;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(lds0 + r * kStageLd * 4));
      val = fmaf(g.alpha, val, cb.v[0]);
      if (col_ok) {
        const int64_t idx = base + static_cast<int64_t>(r) * g.row_len;
        if (lo) lo[idx] = val;
        if (mk) mk[idx] = sigmoid_acc(val);
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major, 128-byte swizzle: LBO (ignored) = 1, SBO = 1024 B, version 1, layout type 2
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int MODE, int ACT, int EPIW>
__global__ void __launch_bounds__(gemm_threads(EPIW), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * BK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(g.bn) * BK * 2;
  const uint32_t sA = base;
  const int kStages = g.stages;
  const uint32_t sB = base + kStages * a_bytes;
  const uint32_t sBar = sB + kStages * b_bytes;  // 8-byte aligned (multiples of 1024 before it)
  const uint32_t full0 = sBar, empty0 = sBar + 8 * kMaxStages, tfull0 = sBar + 16 * kMaxStages,
                 tempty0 = tfull0 + 16, tptr = tempty0 + 16;
  const uint32_t sStage = sBar + 128;  // EPIW x 32 rows x kStageLd words
  constexpr int kEpiSets = EPIW / 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full0 + 8 * s, 1);
        mbar_init(empty0 + 8 * s, 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tfull0 + 8 * s, 1);
        mbar_init(tempty0 + 8 * s, EPIW);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tc_alloc(tptr, kTmemCols);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));

  const int64_t tiles_per_z = static_cast<int64_t>(g.m_tiles) * g.n_tiles;

  // Producer and MMA warps run their loops with all 32 lanes converged and guard only the issue block
  // with elect.sync: inside a divergent `if (lane == 0)` ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in
  // an ELECT + BRA.U.ANY waterfall (~70 cycles per MMA, measured on the recurrence kernel).
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const int z = static_cast<int>(tile / tiles_per_z);
      const int64_t rem = tile - z * tiles_per_z;
      const int mt = static_cast<int>(rem / g.n_tiles), nt = static_cast<int>(rem - static_cast<int64_t>(mt) * g.n_tiles);
      const int a_row = static_cast<int>((z / g.a_div) * g.a_row_stride + static_cast<int64_t>(mt) * BM);
      const int b_row = static_cast<int>((z % g.b_mod) * g.b_row_stride + static_cast<int64_t>(nt) * g.bn);
      for (int kb = 0; kb < g.k_blocks; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full0 + 8 * s, a_bytes + b_bytes);
          tma_load_2d(sA + s * a_bytes, &tmA, full0 + 8 * s, kb * BK, a_row);
          tma_load_2d(sB + s * b_bytes, &tmB, full0 + 8 * s, kb * BK, b_row);
        }
        __syncwarp();
        if (++s == kStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(g.bn >> 3) << 17) |
                           (static_cast<uint32_t>(BM >> 4) << 24);
    const uint64_t adesc0 = make_smem_desc(sA), bdesc0 = make_smem_desc(sB);
    const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4;  // encoded distance of two pipeline stages
    const int tail_ksteps = (g.K - (g.k_blocks - 1) * BK + 15) / 16;  // k-steps of the last k-block that hold data
    int s = 0;
    uint32_t ph = 0;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(tempty0 + 8 * as, aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccCols;
      for (int kb = 0; kb < g.k_blocks; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>(s * a_step), bd = bdesc0 + static_cast<uint64_t>(s * b_step);
          // the last k-block may hold fewer than four k-steps of data (K = 513: one): the rest is zero padding of the
          // operand rows and is not multiplied (3 of 36 MMAs at K = 513, 2 of 40 at K = 608)
          const int ksteps = kb == g.k_blocks - 1 ? tail_ksteps : BK / 16;
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4)
            if (k4 < ksteps) tc_mma_bf16(d_tmem, ad + 2 * k4, bd + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
          tc_commit(empty0 + 8 * s);
          if (kb == g.k_blocks - 1) tc_commit(tfull0 + 8 * as);
        }
        __syncwarp();
        if (++s == kStages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const uint32_t stage = sStage + static_cast<uint32_t>(warp - 2) * (32 * kStageLd * 4);
    const int cset = (warp - 2) >> 2;  // column chunks are dealt in turn to the warp sets
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int z = static_cast<int>(tile / tiles_per_z);
      const int64_t rem = tile - z * tiles_per_z;
      const int mt = static_cast<int>(rem / g.n_tiles), nt = static_cast<int>(rem - static_cast<int64_t>(mt) * g.n_tiles);
      const int64_t m0 = static_cast<int64_t>(mt) * BM + q * 32;
      const int nbase = nt * g.bn;
      const float* bias = g.bias ? g.bias + static_cast<int64_t>(z % g.b_mod) * g.bias_stride : nullptr;
      const int* plane_row = MODE == EPI_HEAD ? g.plane_map + static_cast<int64_t>(z) * g.n_blocks : nullptr;
      const EpiTile et = make_epi_tile<MODE>(g, z, m0, lane);
      ChunkBias cb = load_chunk_bias<MODE>(g, bias, plane_row, nbase + cset * 32, lane);
      mbar_wait(tfull0 + 8 * as, aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccCols;
      for (int c0 = cset * 32; c0 < g.bn; c0 += 32 * kEpiSets) {
        uint32_t v[32];
        tc_ld32(t0 + c0, v);
        const ChunkBias cb_next = load_chunk_bias<MODE>(g, bias, plane_row, nbase + c0 + 32 * kEpiSets, lane);  // for the next chunk
        tc_wait_ld();
        if (m0 < g.M) {
          if (MODE == EPI_HEAD) epilogue_head(g, et, m0, nbase + c0, min(32, g.bn - c0), v, cb, stage, lane);
          else epilogue_rows<MODE, ACT>(g, et, nbase + c0, min(32, g.bn - c0), v, cb, stage, lane);
        }
        cb = cb_next;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * as);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc(tmem_base, kTmemCols);
  }
}

// Debug / bisecting reference: one thread per 32-column strip, plain FMAs.
__global__ void gemm_simt_kernel(const GemmArgs g) {
  const int64_t strips = (g.N + 31) / 32;
  const int64_t total = static_cast<int64_t>(g.batch) * g.M * strips;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int strip = static_cast<int>(i % strips);
    const int64_t rest = i / strips;
    const int64_t m = rest % g.M;
    const int z = static_cast<int>(rest / g.M);
    const __nv_bfloat16* a = g.A + ((z / g.a_div) * g.a_row_stride + m) * g.lda;
    float acc[32];
    for (int j = 0; j < 32; ++j) {
      const int n = strip * 32 + j;
      float s = 0.f;
      if (n < g.N) {
        const __nv_bfloat16* b = g.B + ((z % g.b_mod) * g.b_row_stride + n) * g.ldb;
        for (int k = 0; k < g.K; ++k) s = fmaf(__bfloat162float(a[k]), __bfloat162float(b[k]), s);
      }
      acc[j] = s;
    }
    epilogue_store(g, z, m, strip * 32, 32, acc);
  }
}

static int make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t ld_elems,
                       uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  TSSEP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TSSEP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (inner=%llu rows=%llu ld=%llu box_rows=%u)",
                static_cast<int>(r), (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)ld_elems,
                box_rows);
  return 0;
}

static int choose_bn(int N) {
  const int step = 16;
  if (const char* e = debug_env("TSSEP_GEMM_BN")) {
    const int v = atoi(e);
    if (v >= 16 && v <= 256 && v % step == 0) return v;
  }
  if (N <= 256) return ((N + step - 1) / step) * step;
  int best = 256;
  double best_score = 1e9;
  for (int bn = 256; bn >= 128; bn -= step) {
    const int tiles = (N + bn - 1) / bn;
    const double waste = static_cast<double>(tiles) * bn / N - 1.0;
    const double score = waste + 0.0002 * (256 - bn);
    if (score < best_score) {
      best_score = score;
      best = bn;
    }
  }
  return best;
}

static int launch_gemm(GemmArgs g, const uint16_t* A, int64_t lda, int64_t a_stride, const uint16_t* B, int64_t ldb,
                       int64_t b_stride, int impl, int max_ctas, cudaStream_t stream) {
  TSSEP_REQUIRE(A && B, "gemm: null operand");
  TSSEP_REQUIRE(g.M >= 0 && g.N >= 1 && g.K >= 1 && g.batch >= 1 && g.a_div >= 1 && g.b_mod >= 1 && g.out_div >= 1,
                "gemm: bad extent");
  TSSEP_REQUIRE(lda >= g.K && ldb >= g.K, "gemm: leading dimension smaller than K");
  TSSEP_REQUIRE(a_stride % lda == 0 && b_stride % ldb == 0, "gemm: batch strides must be whole rows");
  if (g.M == 0) return 0;
  g.a_row_stride = a_stride / lda;
  g.b_row_stride = b_stride / ldb;
  g.A = reinterpret_cast<const __nv_bfloat16*>(A);
  g.B = reinterpret_cast<const __nv_bfloat16*>(B);
  g.lda = lda;
  g.ldb = ldb;
  if (impl == 1) {
    const int64_t total = static_cast<int64_t>(g.batch) * g.M * ((g.N + 31) / 32);
    const int blocks = static_cast<int>(imin64((total + 127) / 128, 148 * 32));
    gemm_simt_kernel<<<blocks, 128, 0, stream>>>(g);
    return check_launch("gemm_simt");
  }
  TSSEP_REQUIRE(impl == 0, "gemm: unknown impl %d", impl);
  TSSEP_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 (got %lld, %lld)", (long long)lda,
                (long long)ldb);
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                "gemm: operands must be 16-byte aligned");
  g.bn = choose_bn(g.N);
  g.m_tiles = static_cast<int>((g.M + BM - 1) / BM);
  g.n_tiles = (g.N + g.bn - 1) / g.bn;
  g.k_blocks = (g.K + BK - 1) / BK;
  g.total_tiles = static_cast<int64_t>(g.batch) * g.m_tiles * g.n_tiles;
  const int64_t a_rows = ((g.batch - 1) / g.a_div) * g.a_row_stride + g.M;
  const int64_t b_rows = static_cast<int64_t>((g.batch < g.b_mod ? g.batch : g.b_mod) - 1) * g.b_row_stride + g.N;
  TSSEP_REQUIRE(a_rows < (1ll << 31) && b_rows < (1ll << 31), "gemm: too many rows for 32-bit TMA coordinates");
  CUtensorMap tmA, tmB;
  if (int r = make_map_2d(&tmA, A, g.K, a_rows, lda, BM)) return r;
  if (int r = make_map_2d(&tmB, B, g.K, b_rows, ldb, g.bn)) return r;
  int dev = 0, sms = 148;
  TSSEP_CUDA(cudaGetDevice(&dev));
  TSSEP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t stage_bytes = BM * BK * 2 + static_cast<size_t>(g.bn) * BK * 2;
  const int epiw = (g.mode == EPI_HEAD || g.K <= 384 || g.N <= 512) ? 16 : 8;  // measured per shape, profiles/r2_gemm_microbench.txt
  const size_t fixed = 1024 + 128 + static_cast<size_t>(epiw) * 32 * kStageLd * 4;
  g.stages = static_cast<int>(imin64(kMaxStages, (227 * 1024 - fixed) / stage_bytes));
  TSSEP_REQUIRE(g.stages >= 2, "gemm: tile does not fit shared memory");
  const size_t smem = fixed + g.stages * stage_bytes;
  int grid = static_cast<int>(imin64(g.total_tiles, sms));
  // max_ctas: run the persistent kernel on fewer SMs.  Under the board power cap a large GEMM is power limited, and the
  // clocks it leaves behind decide the speed of the latency-bound recurrence that follows
  // (profiles/r1_power_cap_probe.txt).
  if (max_ctas >= 1 && max_ctas < grid) grid = max_ctas;
#define TSSEP_GEMM_CASE_W(MODE_, ACT_, EPIW_)                                                                            \
  if (g.mode == MODE_ && g.act == ACT_ && epiw == EPIW_) {                                                               \
    TSSEP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE_, ACT_, EPIW_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                    static_cast<int>(smem)));                                                           \
    gemm_tc_kernel<MODE_, ACT_, EPIW_><<<grid, gemm_threads(EPIW_), smem, stream>>>(tmA, tmB, g);                        \
    return check_launch("gemm_tc");                                                                                      \
  }
#define TSSEP_GEMM_CASE(MODE_, ACT_) TSSEP_GEMM_CASE_W(MODE_, ACT_, 8) TSSEP_GEMM_CASE_W(MODE_, ACT_, 16)
  TSSEP_GEMM_CASE(EPI_F32, 0)
  TSSEP_GEMM_CASE(EPI_F32, 1)
  TSSEP_GEMM_CASE(EPI_BF16, 0)
  TSSEP_GEMM_CASE(EPI_BF16, 1)
  TSSEP_GEMM_CASE(EPI_HEAD, 0)
#undef TSSEP_GEMM_CASE
#undef TSSEP_GEMM_CASE_W
  set_error("gemm: no tensor-core instantiation for mode %d act %d", g.mode, g.act);
  return -1;
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_gemm(const tssep_gemm_desc* d, tssep_stream_t stream) {
  TSSEP_REQUIRE(d != nullptr, "tssep_gemm: null descriptor");
  TSSEP_REQUIRE(d->mode >= TSSEP_EPI_F32 && d->mode <= TSSEP_EPI_HEAD, "tssep_gemm: unknown epilogue mode %d",
                d->mode);
  TSSEP_REQUIRE(d->act == 0 || d->act == 1, "tssep_gemm: act must be 0 (none) or 1 (tanh)");
  GemmArgs g{};
  g.M = d->M;
  g.N = d->N;
  g.K = d->K;
  g.batch = d->batch;
  g.a_div = d->a_div;
  g.b_mod = d->b_mod;
  g.bias = d->bias;
  g.bias_stride = d->bias_stride;
  g.alpha = d->alpha;
  g.act = d->act;
  g.mode = d->mode;
  g.out = d->out;
  g.ldo = d->ldo;
  g.out_stride = d->out_stride;
  g.out_stride_hi = d->out_stride_hi;
  g.out_div = d->out_div;
  g.mask = d->mask;
  g.plane_map = d->plane_map;
  g.n_blocks = d->n_blocks;
  g.row_len = d->row_len;
  if (d->mode == TSSEP_EPI_HEAD) {
    TSSEP_REQUIRE(d->out || d->mask, "tssep_gemm(head): no output");
    TSSEP_REQUIRE(d->plane_map && d->n_blocks >= 1 && d->row_len >= 1 && d->N == d->n_blocks * d->row_len,
                  "tssep_gemm(head): need plane_map and N == n_blocks * row_len");
    TSSEP_REQUIRE(d->act == 0, "tssep_gemm(head): act must be 0");
  } else {
    TSSEP_REQUIRE(d->out != nullptr, "tssep_gemm: null output");
    TSSEP_REQUIRE(d->ldo >= d->N, "tssep_gemm: ldo < N");
  }
  TSSEP_REQUIRE(d->max_ctas >= 0, "tssep_gemm: max_ctas must be >= 0");
  return launch_gemm(g, d->A, d->lda, d->a_stride, d->B, d->ldb, d->b_stride, d->impl, d->max_ctas,
                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
