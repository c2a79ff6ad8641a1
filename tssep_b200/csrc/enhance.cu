// Conditioning fold, 't'-resolution head expansion, and enhancement
// (mask x mixture STFT -> iSTFT overlap-add).
//
// Reference operators: MaskEstimator_v2 conditioning tssep/train/net.py:862-896,
// final einops Reduce('repeat') net.py:653-659, Masking.__call__
// tssep/train/enhancer.py:73-100, fe.istft (padertorch) called at
// tssep/train/model.py:661-664.
#include "../../include/tssep_b200.h"
#include "common.cuh"
#include "fft.cuh"

namespace tssep {

// ---------------------------------------------------------------------------
// Conditioning folded into the first post_net input projection
// ---------------------------------------------------------------------------
__global__ void fold_mul_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ b,
                                const float* __restrict__ e, int64_t Z, int N, int F, int A,
                                __nv_bfloat16* __restrict__ Wk, int64_t ld_wk, float* __restrict__ bias_k) {
  const int64_t total = Z * N * ld_wk;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t f = i % ld_wk;
    const int64_t zn = i / ld_wk;
    const int64_t n = zn % N, z = zn / N;
    float v = 0.f;
    if (f < F) v = W[n * ldw + f] * e[z * A + f];
    Wk[i] = __float2bfloat16_rn(v);
    if (f == 0) bias_k[zn] = b[n];
  }
}

__global__ void fold_cat_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ b,
                                const float* __restrict__ e, int64_t Z, int N, int F, int A,
                                float* __restrict__ bias_k) {
  const int lane = threadIdx.x & 31;
  const int64_t zn = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (zn >= Z * N) return;
  const int64_t n = zn % N, z = zn / N;
  const float* w = W + n * ldw + F;
  const float* ez = e + z * A;
  float acc = 0.f;
  for (int a = lane; a < A; a += 32) acc = fmaf(w[a], ez[a], acc);
  acc = warp_sum(acc);
  if (lane == 0) bias_k[zn] = b[n] + acc;
}

// ---------------------------------------------------------------------------
// output_resolution 't': broadcast (Z, T, K) logits over frequency
// ---------------------------------------------------------------------------
__global__ void head_expand_t_kernel(const float* __restrict__ small, int64_t Z, int64_t T, int K, int F,
                                     const int* __restrict__ perm, float* __restrict__ logit,
                                     float* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t total = Z * K * T;
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < total;
       r += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
    const int64_t t = r % T;
    const int64_t zq = r / T;
    const int q = static_cast<int>(zq % K);
    const int64_t z = zq / K;
    const int64_t plane = perm[z * K + q];
    const float v = small[(z * T + t) * K + q];
    const float m = sigmoid_acc(v);
    const int64_t o = (plane * T + t) * F;
    for (int f = lane; f < F; f += 32) {
      if (logit) logit[o + f] = v;
      if (mask) mask[o + f] = m;
    }
  }
}

// ---------------------------------------------------------------------------
// mask x X  ->  (stft_estimate)  ->  iSTFT overlap-add
// One CTA walks a contiguous range of output hops of ONE separated signal (item z, speaker k)
// in rounds of 8 frames: every warp inverse-transforms one frame into a ring of
// 8 + OV - 1 shared-memory slots, then the CTA overlap-adds the 8 hops that just became
// complete in gather form (no atomics, every output sample written exactly once).  Only the
// OV - 1 frames before the range are recomputed (a 2-3 % halo); the K CTAs that need the same
// mixture rows run side by side, so X is served from L2.
// ---------------------------------------------------------------------------
constexpr int kEWarps = 8;
constexpr int kEThreads = kEWarps * 32;

__global__ void __launch_bounds__(kEThreads, 2)
mask_istft_kernel(const float2* __restrict__ X, int64_t x_item_stride, const float* __restrict__ mask, int64_t mp, int n_spk,
                  int64_t T, int S, int R, int wl, int trim, const float* __restrict__ synwin,
                  const float2* __restrict__ twiddle, float2* __restrict__ est, float* __restrict__ time_out,
                  int64_t num_samples, int hops) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = S >> 1, F = M + 1, OV = wl / R;
  const int PL = padded_len(M);
  const int NR = kEWarps + OV - 1;  // ring slots
  float2* tw = reinterpret_cast<float2*>(smem_raw);               // M
  float2* ring = tw + M;                                           // NR * PL : windowed time frames
  float2* scratch = ring + NR * PL;                                // kEWarps * PL
  float* syn = reinterpret_cast<float*>(scratch + kEWarps * PL);   // wl

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // speaker is the fastest-varying block index so that the K CTAs sharing mixture rows are co-resident
  const int64_t z = blockIdx.y;
  const int64_t sig = z * n_spk + (blockIdx.x % n_spk);
  const int64_t J = T + OV - 1;
  const int64_t j_begin = static_cast<int64_t>(blockIdx.x / n_spk) * hops;
  const int64_t j_end = imin64(J, j_begin + hops);
  const int64_t t_lo = j_begin - (OV - 1);  // first frame that contributes to this range
  const float inv_m = 1.0f / static_cast<float>(M);
  const bool odd_passes = fft_num_passes(M) & 1;

  for (int i = threadIdx.x; i < M; i += kEThreads) tw[i] = twiddle[i];
  for (int i = threadIdx.x; i < wl; i += kEThreads) syn[i] = synwin[i] * inv_m;
  __syncthreads();

  const float* sf = reinterpret_cast<const float*>(ring);
  int64_t emitted = j_begin;
  for (int64_t tb = t_lo; tb < j_end; tb += kEWarps) {
    // ---- one frame per warp ------------------------------------------------------------------
    const int64_t t = tb + warp;
    if (t < j_end) {
      float2* slot = ring + static_cast<int>((t - t_lo) % NR) * PL;
      if (t < 0 || t >= T) {
        for (int i = lane; i < PL; i += 32) slot[i] = make_float2(0.f, 0.f);
      } else {
        // the FFT ping-pongs between two buffers; start so that the result lands in the slot
        float2* scr = scratch + warp * PL;
        float2* first = odd_passes ? scr : slot;
        float2* other = odd_passes ? slot : scr;
        const int64_t row = (sig * T + t) * F, mrow = (sig * T + t) * mp;  // mask rows are mp floats apart
        const float2* xrow = mask ? X + z * x_item_stride + t * F : X + row;
        const bool own = est != nullptr && t >= j_begin;
#pragma unroll 4
        for (int kk = lane; kk < M; kk += 32) {
          float2 yk = xrow[kk], ym = xrow[M - kk];
          if (mask) {
            const float mk = mask[mrow + kk], mm = mask[mrow + M - kk];
            yk = make_float2(yk.x * mk, yk.y * mk);
            ym = make_float2(ym.x * mm, ym.y * mm);
          }
          if (own) {
            est[row + kk] = yk;
            if (kk == 0) est[row + M] = ym;
          }
          if (kk == 0) {
            yk.y = 0.f;  // c2r ignores the imaginary parts of DC and Nyquist
            ym.y = 0.f;
          }
          first[padi(kk)] = irfft_pack(yk, ym, kk, tw);
        }
        __syncwarp();
        float2* res = warp_fft<true>(first, other, M, tw, S, lane);  // == slot
        for (int n = lane; n < M; n += 32) {
          float2 c = res[padi(n)];
          c.x *= (2 * n < wl) ? syn[2 * n] : 0.f;
          c.y *= (2 * n + 1 < wl) ? syn[2 * n + 1] : 0.f;
          res[padi(n)] = c;
        }
      }
    }
    __syncthreads();
    // ---- overlap-add the hops whose OV frames are all in the ring now ---------------------------
    const int64_t ready = imin64(j_end, tb + kEWarps);  // frames < ready are done
    if (time_out && ready > emitted) {
      const int count = static_cast<int>(ready - emitted) * R;
      for (int idx = threadIdx.x; idx < count; idx += kEThreads) {
        const int hj = idx / R, i = idx - hj * R;
        const int64_t j = emitted + hj;
        const int64_t n = j * R + i - trim;
        if (n >= 0 && n < num_samples) {
          float acc = 0.f;
          for (int o = 0; o < OV; ++o) {
            const int e = o * R + i;
            const int sl = static_cast<int>((j - o - t_lo) % NR);
            acc += sf[2 * (sl * PL + padi(e >> 1)) + (e & 1)];
          }
          time_out[sig * num_samples + n] = acc;
        }
      }
    }
    if (ready > emitted) emitted = ready;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Fast path for the frame geometry of every shipped config (size 1024, shift 256, window 1024).
//
// The 512-point inverse FFT behind the real iFFT is factored 16 x 32 with BOTH factors in registers:
//   pass A  lane k2 holds Z[32 k1 + k2], k1 < 16 (coalesced loads of X and the mask), runs a
//           16-point DFT, multiplies by w512^(k2 n1) and drops Y[n1][k2] into a padded
//           shared-memory tile: the only shared-memory round trip of the transform;
//   pass B  lane l of a half-warp picks up Y[l][0..31], runs a 32-point DFT and owns
//           z[l + 16 n2], n2 < 32, i.e. samples 2 (l + 16 n2) + {0, 1} of the frame.
// With that ownership the four 256-sample quarters of a frame are the register groups n2 / 8, so
// the overlap-add of consecutive frames is a shift of register groups: the accumulator lives in
// registers, every output sample is written once (128-byte rows per half-warp), no ring, no
// atomics.  A warp transforms the SAME frame of two speakers (lanes 0-15 / 16-31 in pass B) and
// loads the mixture row once for both; the warps of a CTA cover up to 8 speakers of one meeting so
// that the mixture row is served from L1/L2.  Each warp walks a range of `hops` hops and recomputes
// only the 3 frames before it.
// ---------------------------------------------------------------------------
constexpr int kFastWarps = 4;   // speaker pairs per CTA
constexpr int kXLd = 33;        // padded row (float2) of the exchange tile: conflict-free both ways

// exp(+2 pi i k / 32)
__device__ __forceinline__ float2 w32(int k) {
  const float c[32] = {1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 6.123233996e-17f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f, -1.000000000e+00f, -9.807852804e-01f, -9.238795325e-01f, -8.314696123e-01f, -7.071067812e-01f, -5.555702330e-01f, -3.826834324e-01f, -1.950903220e-01f, -1.836970199e-16f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f};
  const float sn[32] = {0.000000000e+00f, 1.950903220e-01f, 3.826834324e-01f, 5.555702330e-01f, 7.071067812e-01f, 8.314696123e-01f, 9.238795325e-01f, 9.807852804e-01f, 1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f, 3.826834324e-01f, 1.950903220e-01f, 1.224646799e-16f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f, -1.000000000e+00f, -9.807852804e-01f, -9.238795325e-01f, -8.314696123e-01f, -7.071067812e-01f, -5.555702330e-01f, -3.826834324e-01f, -1.950903220e-01f};
  return make_float2(c[k], sn[k]);
}

// in-place inverse DFT (unnormalised, exp(+i...)) of N register values, natural order in and out
template <int N>
__device__ __forceinline__ void idft_reg(float2 (&v)[N]) {
  if constexpr (N == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else {
    float2 e[N / 2], o[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      e[i] = v[2 * i];
      o[i] = v[2 * i + 1];
    }
    idft_reg<N / 2>(e);
    idft_reg<N / 2>(o);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      float2 t;
      if (k == 0) t = o[0];
      else if (k == N / 4) t = make_float2(-o[k].y, o[k].x);  // * (+i)
      else t = cmul(o[k], w32(k * (32 / N)));
      v[k] = cadd(e[k], t);
      v[k + N / 2] = csub(e[k], t);
    }
  }
}

template <bool ACT>
__global__ void __launch_bounds__(32 * kFastWarps, 3)
mask_istft_1024_kernel(const float2* __restrict__ X, int64_t x_item_stride, const float* __restrict__ mask, int64_t mp, int n_spk,
                       int groups, int64_t T, int trim, const float* __restrict__ synwin, const float2* __restrict__ twiddle,
                       float2* __restrict__ est, float* __restrict__ time_out, int64_t num_samples, int hops,
                       float* __restrict__ activity) {
  constexpr int M = 512, F = 513, R = 256;
  __shared__ float2 tw[M];                              // exp(-2 pi i k / 1024)
  __shared__ float2 twA[16 * 32];                       // exp(+2 pi i k2 n1 / 512) at [n1 * 32 + k2]
  __shared__ float2 syn2[M];                            // synthesis window pairs / 512
  __shared__ float2 xch[kFastWarps][2][16 * kXLd];      // Y[n1][k2] of the warp's two frames

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    tw[i] = twiddle[i];
    syn2[i] = make_float2(synwin[2 * i] * (1.0f / M), synwin[2 * i + 1] * (1.0f / M));
    const int n1 = i >> 5, k2 = i & 31;
    const int idx = 2 * k2 * n1;  // < 1024
    float2 w = twiddle[idx & (M - 1)];
    if (idx >= M) w = make_float2(-w.x, -w.y);
    twA[i] = make_float2(w.x, -w.y);  // conj: exp(+...)
  }
  __syncthreads();

  const int64_t z = blockIdx.y / groups;
  const int spk0 = (blockIdx.y % groups) * (2 * kFastWarps) + 2 * warp;
  if (spk0 >= n_spk) return;
  const bool has_b = spk0 + 1 < n_spk;
  const int64_t sig_a = z * n_spk + spk0, sig_b = sig_a + (has_b ? 1 : 0);
  const int64_t J = T + 3;
  const int64_t j_begin = static_cast<int64_t>(blockIdx.x) * hops;
  const int64_t j_end = imin64(J, j_begin + hops);
  const int half = lane >> 4, l = lane & 15;
  const int64_t sig_mine = half ? sig_b : sig_a;
  const bool store_mine = time_out != nullptr && (half == 0 || has_b);
  float2* xa = xch[warp][0];
  float2* xb = xch[warp][1];
  const float2* mine = half ? xb : xa;

  float2 acc[3][8];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[g][q] = make_float2(0.f, 0.f);

  for (int64_t t = j_begin - 3; t < j_end; ++t) {
    float2 u[32];
    if (t >= 0 && t < T) {
      // ---- pass A: both speakers, element k = 32 k1 + lane ------------------------------------------
      float2 va[16], vb[16];
      const int64_t row_a = (sig_a * T + t) * F, row_b = (sig_b * T + t) * F;
      const int64_t mrow_a = (sig_a * T + t) * mp, mrow_b = (sig_b * T + t) * mp;  // mask rows are mp floats apart
      const float2* xrow_a = mask ? X + z * x_item_stride + t * F : X + row_a;
      const float2* xrow_b = mask ? xrow_a : X + row_b;
      const bool own = est != nullptr && t >= j_begin;
      if constexpr (ACT) {
        // Frame activity mean_f mask[k, t, f] of the two speakers, from a prelude pass over the two mask rows: it
        // doubles as the prefetch of the rows the transform below reads (the sums are dead before it starts, so the
        // register budget of the transform -- 168 registers at 3 CTAs per SM -- is untouched; accumulating inside the
        // transform loop cost 25 % of the kernel's speed).  Halo frames belong to the previous range.
        if (t >= j_begin) {
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int k1 = 0; k1 < 16; ++k1) {
            sa += mask[mrow_a + 32 * k1 + lane];
            sb += mask[mrow_b + 32 * k1 + lane];
          }
          if (lane == 0) {
            sa += mask[mrow_a + M];
            sb += mask[mrow_b + M];
          }
          sa = warp_sum(sa);
          sb = warp_sum(sb);
          if (lane == 0) {
            activity[sig_a * T + t] = sa / static_cast<float>(F);
            if (has_b) activity[sig_b * T + t] = sb / static_cast<float>(F);
          }
        }
      }
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        const int kk = 32 * k1 + lane;
        float2 yk = xrow_a[kk], ym = xrow_a[M - kk];
        float2 yk2 = yk, ym2 = ym;
        if (mask) {
          const float mk = mask[mrow_a + kk], mm = mask[mrow_a + M - kk];
          const float mk2 = mask[mrow_b + kk], mm2 = mask[mrow_b + M - kk];
          yk = make_float2(yk.x * mk, yk.y * mk);
          ym = make_float2(ym.x * mm, ym.y * mm);
          yk2 = make_float2(yk2.x * mk2, yk2.y * mk2);
          ym2 = make_float2(ym2.x * mm2, ym2.y * mm2);
        } else if (has_b) {
          yk2 = xrow_b[kk];
          ym2 = xrow_b[M - kk];
        }
        if (own) {
          est[row_a + kk] = yk;
          if (has_b) est[row_b + kk] = yk2;
          if (kk == 0) {
            est[row_a + M] = ym;
            if (has_b) est[row_b + M] = ym2;
          }
        }
        if (kk == 0) {  // c2r ignores the imaginary parts of DC and Nyquist
          yk.y = ym.y = 0.f;
          yk2.y = ym2.y = 0.f;
        }
        const float2 w = tw[kk];
        va[k1] = irfft_pack_w(yk, ym, w);
        vb[k1] = irfft_pack_w(yk2, ym2, w);
      }
      idft_reg<16>(va);
      idft_reg<16>(vb);
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const float2 w = twA[n1 * 32 + lane];
        xa[n1 * kXLd + lane] = n1 ? cmul(va[n1], w) : va[0];
        xb[n1 * kXLd + lane] = n1 ? cmul(vb[n1], w) : vb[0];
      }
      __syncwarp();
      // ---- pass B: 32-point DFT of row l, then the synthesis window -----------------------------------
#pragma unroll
      for (int k2 = 0; k2 < 32; ++k2) u[k2] = mine[l * kXLd + k2];
      __syncwarp();
      idft_reg<32>(u);
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) {
        const float2 w = syn2[l + 16 * n2];
        u[n2] = make_float2(u[n2].x * w.x, u[n2].y * w.y);
      }
    } else {
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) u[n2] = make_float2(0.f, 0.f);
    }
    // ---- overlap-add in registers: hop t is complete once frame t's first quarter is added -------------
    const bool emit = store_mine && t >= j_begin;
    float* orow = time_out + sig_mine * num_samples;
    const int64_t n0 = t * R + 2 * l - trim;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 o = cadd(acc[0][q], u[q]);
      const int64_t n = n0 + 32 * q;
      if (emit) {
        if (n >= 0 && n + 1 < num_samples && ((reinterpret_cast<uintptr_t>(orow + n) & 7) == 0)) {
          *reinterpret_cast<float2*>(orow + n) = o;
        } else {
          if (n >= 0 && n < num_samples) orow[n] = o.x;
          if (n + 1 >= 0 && n + 1 < num_samples) orow[n + 1] = o.y;
        }
      }
      acc[0][q] = cadd(acc[1][q], u[8 + q]);
      acc[1][q] = cadd(acc[2][q], u[16 + q]);
      acc[2][q] = u[24 + q];
    }
  }
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

// separated audio -> 16-bit PCM (what the evaluation writes to disk, tssep_b200/eval.py::write_wav): round to nearest,
// saturate.  Halves the device-to-host bytes of a serving loop that ships audio (bench.py, e2e_pcm16).
__global__ void pcm16_kernel(const float* __restrict__ x, int64_t n, float scale, int16_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<int16_t>(fminf(fmaxf(rintf(x[i] * scale), -32768.f), 32767.f));
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_fold_embedding(int mode, const float* W, int64_t ldw, const float* b, const float* e, int64_t Z, int N,
                         int F, int A, uint16_t* Wk, int64_t ld_wk, float* bias_k, tssep_stream_t stream) {
  TSSEP_REQUIRE(W && b && e && bias_k, "tssep_fold_embedding: null pointer");
  TSSEP_REQUIRE(mode == 0 || mode == 1, "tssep_fold_embedding: mode must be 0 (mul) or 1 (cat)");
  if (Z == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == 0) {
    TSSEP_REQUIRE(Wk && ld_wk >= F && A == F && ldw >= F, "tssep_fold_embedding(mul): need Wk, ld_wk >= F, A == F");
    const int64_t total = Z * N * ld_wk;
    const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 16));
    fold_mul_kernel<<<blocks, 256, 0, s>>>(W, ldw, b, e, Z, N, F, A, reinterpret_cast<__nv_bfloat16*>(Wk), ld_wk,
                                           bias_k);
  } else {
    TSSEP_REQUIRE(ldw >= F + A, "tssep_fold_embedding(cat): ldw < F + A");
    const int64_t rows = Z * N;
    fold_cat_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(W, ldw, b, e, Z, N, F, A, bias_k);
  }
  return check_launch("tssep_fold_embedding");
}

int tssep_head_expand_t(const float* small, int64_t Z, int64_t T, int n_spk, int F, const int32_t* perm,
                        float* logit, float* mask, tssep_stream_t stream) {
  TSSEP_REQUIRE(small && perm && (logit || mask), "tssep_head_expand_t: null pointer");
  if (Z == 0 || T == 0) return 0;
  const int64_t rows = Z * n_spk * T;
  const int blocks = static_cast<int>(imin64((rows + 7) / 8, 148 * 16));
  head_expand_t_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(small, Z, T, n_spk, F, perm, logit, mask);
  return check_launch("tssep_head_expand_t");
}

int tssep_pcm16(const float* x, int64_t n, float scale, int16_t* out, tssep_stream_t stream) {
  TSSEP_REQUIRE(n >= 0, "tssep_pcm16: bad extent");
  if (n == 0) return 0;
  TSSEP_REQUIRE(x && out, "tssep_pcm16: null pointer");
  const int blocks = static_cast<int>(imin64((n + 1023) / 1024, 148 * 16));
  pcm16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scale, out);
  return check_launch("tssep_pcm16");
}

int tssep_mask_istft(const float* X, int64_t x_item_stride, const float* mask, int64_t mask_pitch, int64_t Z, int n_spk, int64_t T,
                     int size, int shift, int window_length, int fading, const float* synwin, const float* twiddle,
                     float* stft_estimate, float* time, int64_t num_samples, float* activity, tssep_stream_t stream) {
  TSSEP_REQUIRE(X && synwin && twiddle, "tssep_mask_istft: null pointer");
  TSSEP_REQUIRE(stft_estimate || time, "tssep_mask_istft: no output requested");
  TSSEP_REQUIRE(activity == nullptr || mask != nullptr, "tssep_mask_istft: activity needs a mask");
  const int l2 = ilog2_exact(size);
  TSSEP_REQUIRE(l2 >= 3 && size <= 4096, "tssep_mask_istft: size must be a power of two in [8, 4096], got %d", size);
  TSSEP_REQUIRE(window_length <= size && shift >= 1 && window_length % shift == 0,
                "tssep_mask_istft: need window_length <= size and window_length %% shift == 0");
  TSSEP_REQUIRE(Z >= 0 && Z < 65536 && n_spk >= 1 && T >= 0, "tssep_mask_istft: bad extent");
  TSSEP_REQUIRE(mask_pitch == 0 || mask_pitch >= size / 2 + 1, "tssep_mask_istft: mask_pitch must be 0 or >= size / 2 + 1");
  const int64_t mp = mask_pitch > 0 ? mask_pitch : size / 2 + 1;
  if (Z == 0 || T == 0) return 0;
  if (size == 1024 && shift == 256 && window_length == 1024 && debug_env("TSSEP_ISTFT_GENERIC") == nullptr) {
    // (the fast kernel loses nothing by also reducing the mask rows it reads to the frame activity)
    const int64_t J = T + 3;
    int hops = 128;
    const int groups = (n_spk + 2 * kFastWarps - 1) / (2 * kFastWarps);
    while (hops > 16 && ((J + hops - 1) / hops) * Z * groups < 3 * 148) hops /= 2;
    dim3 grid(static_cast<unsigned>((J + hops - 1) / hops), static_cast<unsigned>(Z * groups));
    if (activity != nullptr)
      mask_istft_1024_kernel<true><<<grid, 32 * kFastWarps, 0, static_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const float2*>(X), x_item_stride, mask, mp, n_spk, groups, T, fading ? window_length - shift : 0, synwin,
          reinterpret_cast<const float2*>(twiddle), reinterpret_cast<float2*>(stft_estimate), time, num_samples, hops,
          activity);
    else
      mask_istft_1024_kernel<false><<<grid, 32 * kFastWarps, 0, static_cast<cudaStream_t>(stream)>>>(
          reinterpret_cast<const float2*>(X), x_item_stride, mask, mp, n_spk, groups, T, fading ? window_length - shift : 0, synwin,
          reinterpret_cast<const float2*>(twiddle), reinterpret_cast<float2*>(stft_estimate), time, num_samples, hops,
          activity);
    return check_launch("tssep_mask_istft");
  }
  // generic geometries: the frame activity comes from the stand-alone reduction
  if (activity != nullptr) {
    if (int r = tssep_activity(mask, Z * n_spk, T, size / 2 + 1, mp, activity, stream)) return r;
  }
  const int M = size / 2, OV = window_length / shift;
  const size_t smem = sizeof(float2) * (M + (2 * kEWarps + OV - 1) * padded_len(M)) + sizeof(float) * window_length;
  TSSEP_REQUIRE(smem <= 227 * 1024, "tssep_mask_istft: frame geometry does not fit shared memory");
  TSSEP_CUDA(cudaFuncSetAttribute(mask_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t J = T + OV - 1;
  const int64_t n_sig = Z * n_spk;
  // ranges of 128 hops (a 2-3 % halo), shorter when that would leave SMs idle
  int hops = 128;
  while (hops > 8 && ((J + hops - 1) / hops) * n_sig < 4 * 148) hops /= 2;
  dim3 grid(static_cast<unsigned>(((J + hops - 1) / hops) * n_spk), static_cast<unsigned>(Z));
  mask_istft_kernel<<<grid, kEThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(X), x_item_stride, mask, mp, n_spk, T, size, shift, window_length,
      fading ? window_length - shift : 0, synwin, reinterpret_cast<const float2*>(twiddle),
      reinterpret_cast<float2*>(stft_estimate), time, num_samples, hops);
  return check_launch("tssep_mask_istft");
}

}  // extern "C"
