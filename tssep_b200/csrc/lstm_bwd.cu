// Backward pass of the BLSTM recurrence (BPTT) for the training step (BASELINE config 5; reference: autograd through
// torch.nn.LSTM inside RNNP_packed, tssep/train/rnnp.py:143-159, driven by tssep/train/loss.py:219-247).
//
// With a_t the gate pre-activations, the reverse-time recurrence is
//     dh_t = dH_t + W_hh^T . da_{t+1}        (t+1 = the step the forward recurrence took AFTER t)
//     dc_t = dc_{t+1} * f_{t+1} + dh_t * o_t * (1 - tanh(c_t)^2)
//     da_t = [dc_t g_t i_t (1 - i_t),  dc_t c_{t-1} f_t (1 - f_t),  dc_t i_t (1 - g_t^2),  dh_t tanh(c_t) o_t (1 - o_t)]
// The forward kernel (csrc/lstm_ts.cu, SAVE variant) stored the gate activations and c_t.
//
// Same cluster decomposition as the forward kernel: one cluster of C = ceil(Up/64) CTAs per (8 batch rows,
// direction), CTA c owns hidden units 64c .. 64c+63, i.e. 256 gate rows.  The contraction W_hh^T . da runs over ALL
// 4*Up gate rows, so it is split along K: every CTA keeps the rows of W_hh it owns -- transposed, (units x own gate
// rows), as ceil(C/2) M=128 tiles of 128 TMEM columns, resident for the whole sequence -- multiplies them with its OWN
// da (which never leaves the CTA: it sits in shared memory as the K-major B operand) and ships the partial sums of
// the units it does not own to their owners through DSMEM (st.async, complete_tx on the owner's mbarrier); the owner
// adds the C partial sums.  Per step and CTA: 16 * ceil(C/2) tcgen05.mma, 2 KB to each peer.
//   warp 0         MMA issuer (convergent, elect.sync)
//   warps 4..      readers: tcgen05.ld of a (tile, lane quarter), partial sums -> owner CTA
//   last 8 warps   pointwise: thread = (unit, two batch rows): sums the partials, gate derivatives, writes da to the
//                  B operand (bf16) and to dG, keeps dc in registers; the saved activations of the next step are
//                  prefetched while the exchange is in flight
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

struct RecBwdArgs {
  const uint4* WTimg;        // [dir][cta][m tile][k-step 16][lane 128][8 words]
  const uint2* gates;        // (rows, T, 2, Up) x {i, f, g, o} bf16
  const float* cstate;       // (rows, T, 2, Up) f32
  const __nv_bfloat16* dH;   // (rows, T, 2*Up) bf16
  uint2* dG;                 // (rows, T, 2, Up) x 4 bf16, [unit][gate]
  int rows, T, Up, MT;
};

constexpr int kBwdRows = 8;                    // batch rows per cluster
constexpr int kBwdNB = 16;                     // MMA N
constexpr uint32_t kBwdAtom = kBwdNB * 128;    // one 64-k atom of the da operand
constexpr int kBwdReaderWarp0 = 4;

__device__ __forceinline__ uint64_t bwd_desc_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void bwd_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void bwd_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x),
               "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void bwd_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void bwd_st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                                uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   remote_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}

__global__ void __launch_bounds__(1024, 1)
blstm_bwd_kernel(const RecBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t C = cluster_nctarank(), crank = cluster_ctarank();
  const int MT = a.MT, T = a.T, Up = a.Up;
  const uint32_t sBop = base;                               // da operand: 4 atoms x (16 rows x 128 B)
  const uint32_t sP = sBop + 4u * kBwdAtom;                 // partial sums [2 buffers][C sources][64 units][8 rows] f32
  const uint32_t p_buf = C * 2048u;
  const uint32_t sBar = sP + 2u * p_buf;
  const uint32_t accfull = sBar, accempty = sBar + 8, daready = sBar + 16, pfull0 = sBar + 24 /* [2] */, tptr = sBar + 40;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_readers = 4 * MT;
  const int pw0 = kBwdReaderWarp0 + n_readers;  // first pointwise warp
  const int dir = blockIdx.z;
  const int row0 = blockIdx.y * kBwdRows;

  for (uint32_t i = threadIdx.x; i < (sBar - sBop) / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sBop + 16 * i), "r"(0u) : "memory");
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(accfull, 1);
      mbar_init(accempty, n_readers);
      mbar_init(daready, 8);
      mbar_init(pfull0, 1);
      mbar_init(pfull0 + 8, 1);
      mbar_fence_init();
      mbar_arrive_expect_tx(pfull0, p_buf);
      mbar_arrive_expect_tx(pfull0 + 8, p_buf);
    }
    __syncwarp();
    tc_alloc(tptr, 512);
    tc_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  const uint32_t acc_col = static_cast<uint32_t>(MT) * 128u;  // accumulators behind the A tiles

  if (warp >= kBwdReaderWarp0 && warp < pw0) {
    // W_hh^T slice -> TMEM: lane = unit of the tile, 8 columns (16 own gate rows) per store
    const int mt = (warp - kBwdReaderWarp0) >> 2, q = warp & 3;
    const uint4* src = a.WTimg + ((static_cast<size_t>(dir) * C + crank) * MT + mt) * 16 * 256 + static_cast<size_t>(q * 32 + lane) * 2;
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt) * 128u;
    for (int k = 0; k < 16; ++k) {
      const uint4 w0 = __ldg(src + static_cast<size_t>(k) * 256), w1 = __ldg(src + static_cast<size_t>(k) * 256 + 1);
      bwd_st8(t0 + k * 8, w0, w1);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();

  if (warp == 0) {
    // ---- MMA issuer: dh_partial(s) = W_own^T . da(s-1) for s >= 1 --------------------------------------------------
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBwdNB >> 3) << 17) |
                           (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t bdesc0 = bwd_desc_sw128(sBop);
    for (int s = 1; s < T; ++s) {
      mbar_wait(daready, (s - 1) & 1);
      if (s > 1) mbar_wait(accempty, s & 1);  // the readers drained the accumulators of step s-1: completion #(s-2)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_after();
      if (elect_one()) {
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t d = tmem_base + acc_col + mt * kBwdNB;
          const uint32_t at = tmem_base + static_cast<uint32_t>(mt) * 128u;
#pragma unroll
          for (int k = 0; k < 16; ++k)
            bwd_mma(d, at + k * 8, bdesc0 + static_cast<uint64_t>((k >> 2) * (kBwdAtom >> 4) + 2 * (k & 3)), idesc, k > 0 ? 1u : 0u);
        }
        tc_commit(accfull);
      }
      __syncwarp();
    }
  } else if (warp >= kBwdReaderWarp0 && warp < pw0) {
    // ---- readers: partial sums of 32 units -> the CTA that owns them ----------------------------------------------
    const int mt = (warp - kBwdReaderWarp0) >> 2, q = warp & 3;
    const int unit_g = 128 * mt + 32 * q + lane;
    const uint32_t owner = static_cast<uint32_t>(unit_g >> 6);
    const bool send = owner < C;
    const uint32_t dst = send ? mapa(sP, owner) + (crank * 64u + static_cast<uint32_t>(unit_g & 63)) * 32u : 0;
    const uint32_t dbar = send ? mapa(pfull0, owner) : 0;
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_col + mt * kBwdNB;
    for (int s = 1; s < T; ++s) {
      mbar_wait(accfull, (s - 1) & 1);
      tc_fence_after();
      uint32_t v[8];
      bwd_ld8(t_acc, v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty);
      if (send) {
        const uint32_t off = static_cast<uint32_t>(s & 1) * p_buf;
        bwd_st_async_v4(dst + off, v[0], v[1], v[2], v[3], dbar + 8 * (s & 1));
        bwd_st_async_v4(dst + off + 16, v[4], v[5], v[6], v[7], dbar + 8 * (s & 1));
      }
    }
  } else if (warp >= pw0 && warp < pw0 + 8) {
    // ---- pointwise: (unit, two batch rows) per thread ------------------------------------------------------------------
    const int tid = threadIdx.x - pw0 * 32;
    const int ul64 = tid >> 2, rp = tid & 3;
    const int unit = static_cast<int>(crank) * 64 + ul64;
    const bool unit_ok = unit < Up;
    float dc[2] = {0.f, 0.f};
    // element (row, t, dir, unit) of the saved tensors; the recurrence of direction 0 ran t = 0..T-1, so its backward
    // pass walks t = T-1..0 and "the step before" is t-1; direction 1 mirrored
    const int64_t t_first = dir ? 0 : T - 1, t_step = dir ? 1 : -1;
    int64_t idx[2];
    bool ok[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = row0 + 2 * rp + j;
      ok[j] = unit_ok && r < a.rows;
      idx[j] = ((static_cast<int64_t>(ok[j] ? r : 0) * T + t_first) * 2 + dir) * Up + (unit_ok ? unit : 0);
    }
    const int64_t idx_step = t_step * 2 * Up;
    uint2 gt[2];
    float ct[2], cp[2], dh_in[2];
    auto load = [&](int s) {
      const int64_t t = t_first + t_step * s;
      const bool has_prev = dir ? (t + 1 < T) : (t > 0);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (ok[j]) {
          const int64_t e = idx[j] + idx_step * s;
          gt[j] = __ldg(a.gates + e);
          ct[j] = __ldg(a.cstate + e);
          cp[j] = has_prev ? __ldg(a.cstate + e + idx_step) : 0.f;  // the forward step before t = the next one of this walk
          // dH (rows, T, 2*Up): same (row, t) -> offset e + dir-independent arithmetic: ((r*T + t)*2 + dir)*Up + unit
          dh_in[j] = __bfloat162float(a.dH[e]);
        } else {
          gt[j] = make_uint2(0, 0);
          ct[j] = cp[j] = dh_in[j] = 0.f;
        }
      }
    };
    load(0);
    const uint32_t my_p = sP + static_cast<uint32_t>(ul64) * 32u + static_cast<uint32_t>(rp) * 8u;
    for (int s = 0; s < T; ++s) {
      float dh[2] = {dh_in[0], dh_in[1]};
      const uint2 g0 = gt[0], g1 = gt[1];
      const float c0 = ct[0], c1 = ct[1], q0 = cp[0], q1 = cp[1];
      if (s + 1 < T) load(s + 1);  // prefetch while the exchange of this step is in flight
      if (s > 0) {
        const uint32_t pb = pfull0 + 8 * (s & 1);
        mbar_wait(pb, ((s - 1) >> 1) & 1);
        const uint32_t pbase = my_p + static_cast<uint32_t>(s & 1) * p_buf;
        for (uint32_t src = 0; src < C; ++src) {
          float x, y;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(pbase + src * 2048u));
          dh[0] += x;
          dh[1] += y;
        }
        // all 256 threads have read the buffer before it is re-armed for step s+2 (named barrier of the 8 warps)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) mbar_arrive_expect_tx(pb, p_buf);
      }
      const uint2 gg[2] = {g0, g1};
      const float cc[2] = {c0, c1}, cpv[2] = {q0, q1};
      uint2 out[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float gi = __uint_as_float(gg[j].x << 16), gf = __uint_as_float(gg[j].x & 0xffff0000u);
        const float gc = __uint_as_float(gg[j].y << 16), go = __uint_as_float(gg[j].y & 0xffff0000u);
        const float th = tanh_acc(cc[j]);
        const float d_o = dh[j] * th;
        const float dcv = fmaf(dh[j] * go, 1.f - th * th, dc[j]);
        const float da_i = dcv * gc * gi * (1.f - gi);
        const float da_f = dcv * cpv[j] * gf * (1.f - gf);
        const float da_g = dcv * gi * (1.f - gc * gc);
        const float da_o = d_o * go * (1.f - go);
        dc[j] = dcv * gf;
        out[j] = make_uint2(pack_bf16x2(da_i, da_f), pack_bf16x2(da_g, da_o));
      }
      // da -> B operand of the next step's MMAs (K-major, 128-byte swizzle: k = 4*unit + gate) and -> dG
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t n = static_cast<uint32_t>(2 * rp + j);
        const uint32_t addr = sBop + static_cast<uint32_t>(ul64 >> 4) * kBwdAtom + n * 128u +
                              (((static_cast<uint32_t>(ul64 & 15) >> 1) ^ n) << 4) + static_cast<uint32_t>(ul64 & 1) * 8u;
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(out[j].x), "r"(out[j].y) : "memory");
        if (ok[j]) a.dG[idx[j] + idx_step * s] = out[j];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(daready);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512);
  }
  cluster_sync_all();
}

// weight_hh (4U, U) f32 -> transposed tensor-memory image of the BPTT kernel:
// [dir][cta][m tile][k-step 16][lane 128][8 words]; lane = unit 128*mt + lane, word j of k-step k = bf16 pair of the CTA's own
// gate rows kk = 16k + 2j, 16k + 2j + 1 with kk = 4 * (unit - 64 cta) + gate
__global__ void pack_whh_bwd_kernel(const float* __restrict__ w_fwd, const float* __restrict__ w_bwd, int U, int C, int MT,
                                    uint32_t* __restrict__ out) {
  const int64_t total = 2ll * C * MT * 16 * 128 * 8;
  for (int64_t o = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; o < total;
       o += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t rr = o;
    const int j = static_cast<int>(rr % 8);
    rr /= 8;
    const int m = static_cast<int>(rr % 128);
    rr /= 128;
    const int k = static_cast<int>(rr % 16);
    rr /= 16;
    const int mt = static_cast<int>(rr % MT);
    rr /= MT;
    const int cta = static_cast<int>(rr % C);
    const int dir = static_cast<int>(rr / C);
    const float* w = dir ? w_bwd : w_fwd;
    const int col = 128 * mt + m;  // the unit this lane accumulates (column of W_hh)
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = 16 * k + 2 * j + h;
      const int unit_own = cta * 64 + (kk >> 2), gate = kk & 3;
      v[h] = (unit_own < U && col < U) ? w[(static_cast<size_t>(gate) * U + unit_own) * U + col] : 0.f;
    }
    out[o] = pack_bf16x2(v[0], v[1]);
  }
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_pack_whh_bwd(const float* whh_fwd, const float* whh_bwd, int U, int Up, uint32_t* WTimg, tssep_stream_t stream) {
  TSSEP_REQUIRE(whh_fwd && whh_bwd && WTimg, "tssep_pack_whh_bwd: null pointer");
  TSSEP_REQUIRE(U >= 1 && Up >= U && Up % 16 == 0 && Up <= 384, "tssep_pack_whh_bwd: need U <= Up, Up %% 16 == 0, Up <= 384");
  const int C = (Up + 63) / 64, MT = (C + 1) / 2;
  const int64_t total = 2ll * C * MT * 16 * 128 * 8;
  const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 32));
  pack_whh_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(whh_fwd, whh_bwd, U, C, MT, WTimg);
  return check_launch("tssep_pack_whh_bwd");
}

int tssep_blstm_recurrence_bwd(const uint16_t* gates, const float* cstate, const uint16_t* dH, const uint32_t* WTimg,
                               uint16_t* dG, int64_t rows, int64_t T, int Up, tssep_stream_t stream) {
  TSSEP_REQUIRE(gates && cstate && dH && WTimg && dG, "tssep_blstm_recurrence_bwd: null pointer");
  TSSEP_REQUIRE(Up >= 16 && Up % 16 == 0 && Up <= 384, "tssep_blstm_recurrence_bwd: Up must be a multiple of 16 in [16, 384]");
  TSSEP_REQUIRE(rows >= 0 && T >= 0 && T < (1ll << 30) && (rows + 7) / 8 <= 65535, "tssep_blstm_recurrence_bwd: bad extent");
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(gates) & 7) == 0 && (reinterpret_cast<uintptr_t>(dG) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(WTimg) & 15) == 0,
                "tssep_blstm_recurrence_bwd: gates / dG must be 8-byte, WTimg 16-byte aligned");
  if (rows == 0 || T == 0) return 0;
  const int C = (Up + 63) / 64, MT = (C + 1) / 2;
  RecBwdArgs a;
  a.WTimg = reinterpret_cast<const uint4*>(WTimg);
  a.gates = reinterpret_cast<const uint2*>(gates);
  a.cstate = cstate;
  a.dH = reinterpret_cast<const __nv_bfloat16*>(dH);
  a.dG = reinterpret_cast<uint2*>(dG);
  a.rows = static_cast<int>(rows);
  a.T = static_cast<int>(T);
  a.Up = Up;
  a.MT = MT;
  // the kernel owns all 512 TMEM columns of its SM: more than half of the shared memory keeps a second CTA away
  const size_t smem = 120 * 1024;
  TSSEP_CUDA(cudaFuncSetAttribute(blstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, static_cast<unsigned>((rows + kBwdRows - 1) / kBwdRows), 2);
  cfg.blockDim = dim3(32 * (kBwdReaderWarp0 + 4 * MT + 8));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TSSEP_CUDA(cudaLaunchKernelEx(&cfg, blstm_bwd_kernel, a));
  return check_launch("blstm_bwd");
}

}  // extern "C"
