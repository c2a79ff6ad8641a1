// Feature front end: framed STFT, log-mel/MFCC, log1p max-normalised spectrum,
// instance normalisation.  HBM-bound streaming kernels; one warp owns one frame
// at a time, frames staged through shared memory.
//
// Reference operators: padertorch STFT (call site tssep/train/model.py:504),
// TorchMFCC.stft_to_feature (tssep/train/feature_extractor_torchaudio.py:93-106),
// Log1pMaxNormAbsSTFT.stft_to_feature (tssep/train/feature_extractor.py:233-248),
// ConcaternatedSTFTFeatures.stft_to_feature (feature_extractor.py:352-360),
// InstanceNorm (tssep/train/net.py:250-285).
#include "../../include/tssep_b200.h"
#include "common.cuh"
#include "fft.cuh"

namespace tssep {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

// ---------------------------------------------------------------------------
// STFT
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
stft_kernel(const float* __restrict__ audio, int64_t N, const float* __restrict__ window,
            const float2* __restrict__ twiddle, int S, int R, int wl, int front_pad, int64_t T,
            float2* __restrict__ X, int frames_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = S >> 1;
  const int PL = padded_len(M);
  float2* tw = reinterpret_cast<float2*>(smem_raw);                 // M
  float2* scratch = tw + M;                                          // kWarps * 2 * PL
  float* win = reinterpret_cast<float*>(scratch + kWarps * 2 * PL);  // wl
  float* span = win + wl;                                            // (fpb-1)*R + wl

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t sig = blockIdx.y;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * frames_per_block;
  const int nf = static_cast<int>(imin64(frames_per_block, T - t0));
  const int span_len = (nf - 1) * R + wl;

  for (int i = threadIdx.x; i < M; i += kThreads) tw[i] = twiddle[i];
  for (int i = threadIdx.x; i < wl; i += kThreads) win[i] = window[i];
  const float* a = audio + sig * N;
  const int64_t p0 = t0 * R - front_pad;  // original-signal index of span[0]
  for (int i = threadIdx.x; i < span_len; i += kThreads) {
    const int64_t n = p0 + i;
    span[i] = (n >= 0 && n < N) ? __ldg(a + n) : 0.f;
  }
  __syncthreads();

  float2* za = scratch + warp * 2 * PL;
  float2* zb = za + PL;
  const int F = M + 1;
  for (int f = warp; f < nf; f += kWarps) {
    const float* fr = span + f * R;
    for (int n = lane; n < M; n += 32) {
      const int i0 = 2 * n, i1 = 2 * n + 1;
      const float v0 = i0 < wl ? fr[i0] * win[i0] : 0.f;
      const float v1 = i1 < wl ? fr[i1] * win[i1] : 0.f;
      za[padi(n)] = make_float2(v0, v1);
    }
    __syncwarp();
    const float2* z = warp_fft<false>(za, zb, M, tw, S, lane);
    float2* out = X + (sig * T + t0 + f) * F;
    for (int k = lane; k <= M; k += 32) out[k] = rfft_unpack(z, k, M, tw);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// Feature statistics: max |X| per item, mel dB per frame, max dB per item
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
feature_stats_kernel(const float2* __restrict__ X, int64_t x_item_stride, int64_t T, int F,
                     const float* __restrict__ mel_t, const int* __restrict__ mel_lo, const int* __restrict__ mel_hi,
                     int n_mels, uint32_t* __restrict__ absmax_key, uint32_t* __restrict__ maxdb_key,
                     float* __restrict__ meldb, int frames_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* pw = reinterpret_cast<float*>(smem_raw) + (threadIdx.x >> 5) * F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = blockIdx.y;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * frames_per_block;
  const int nf = static_cast<int>(imin64(frames_per_block, T - t0));
  float pmax = 0.f;
  float dbmax = -INFINITY;
  for (int f = warp; f < nf; f += kWarps) {
    const int64_t t = t0 + f;
    const float2* x = X + item * x_item_stride + t * F;
    for (int k = lane; k < F; k += 32) {
      const float2 v = x[k];
      const float p = v.x * v.x + v.y * v.y;
      pw[k] = p;
      pmax = fmaxf(pmax, p);
    }
    __syncwarp();
    for (int m = lane; m < n_mels; m += 32) {
      const float* w = mel_t + static_cast<size_t>(m) * F;
      float acc = 0.f;
      for (int k = mel_lo[m]; k < mel_hi[m]; ++k) acc = fmaf(pw[k], w[k], acc);
      const float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
      meldb[(item * T + t) * n_mels + m] = db;
      dbmax = fmaxf(dbmax, db);
    }
    __syncwarp();
  }
  pmax = warp_max(pmax);
  dbmax = warp_max(dbmax);
  if (lane == 0) {
    atomicMax(absmax_key + item, float_key(sqrtf(pmax)));
    if (n_mels > 0) atomicMax(maxdb_key + item, float_key(dbmax));
  }
}

// ---------------------------------------------------------------------------
// Feature rows: [ mfcc | log1p(|X| (e-1)/max|X|) ]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
feature_write_kernel(const float2* __restrict__ X, int64_t x_item_stride, int64_t n_items, int64_t T, int F,
                     const uint32_t* __restrict__ absmax_key, const uint32_t* __restrict__ maxdb_key,
                     const float* __restrict__ meldb, const float* __restrict__ dct, int n_mels, int n_mfcc,
                     int with_log1p, float top_db, int couple_batch, float* __restrict__ feat_f32,
                     __nv_bfloat16* __restrict__ feat_bf16, int64_t ld_bf16, int frames_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* mrow = reinterpret_cast<float*>(smem_raw) + (threadIdx.x >> 5) * (n_mels > 0 ? n_mels : 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = blockIdx.y;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * frames_per_block;
  const int nf = static_cast<int>(imin64(frames_per_block, T - t0));
  const int Din = n_mfcc + (with_log1p ? F : 0);

  float cut = -INFINITY;
  if (n_mfcc > 0) {
    float mx = key_float(maxdb_key[item]);
    if (couple_batch)
      for (int64_t i = 0; i < n_items; ++i) mx = fmaxf(mx, key_float(maxdb_key[i]));
    cut = mx - top_db;
  }
  const float scale = with_log1p ? __fdiv_rn(1.71828182845904523536f, key_float(absmax_key[item])) : 0.f;

  for (int f = warp; f < nf; f += kWarps) {
    const int64_t t = t0 + f;
    const int64_t row = item * T + t;
    if (n_mfcc > 0) {
      for (int m = lane; m < n_mels; m += 32) mrow[m] = fmaxf(meldb[row * n_mels + m], cut);
      __syncwarp();
      for (int c = lane; c < n_mfcc; c += 32) {
        float acc = 0.f;
        for (int m = 0; m < n_mels; ++m) acc = fmaf(mrow[m], dct[m * n_mfcc + c], acc);
        if (feat_f32) feat_f32[row * Din + c] = acc;
        if (feat_bf16) feat_bf16[row * ld_bf16 + c] = __float2bfloat16_rn(acc);
      }
      __syncwarp();
    }
    if (with_log1p) {
      const float2* x = X + item * x_item_stride + t * F;
      for (int k = lane; k < F; k += 32) {
        const float2 v = x[k];
        const float a = sqrtf(v.x * v.x + v.y * v.y);
        const float y = log1pf(a * scale);
        if (feat_f32) feat_f32[row * Din + n_mfcc + k] = y;
        if (feat_bf16) feat_bf16[row * ld_bf16 + n_mfcc + k] = __float2bfloat16_rn(y);
      }
    }
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src,
                                 __nv_bfloat16* __restrict__ dst, int64_t ld_dst) {
  const int64_t total = rows * ld_dst;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / ld_dst, c = i - r * ld_dst;
    dst[i] = __float2bfloat16_rn(c < cols ? src[r * ld_src + c] : 0.f);
  }
}

// Normalisation along one axis of a contiguous (outer, cols, inner) tensor (InstanceNorm / InstanceNorm_v2,
// tssep/train/net.py:250-330).  mode 0: (x - mean) / std, mode 1: x - mean, mode 2: x / rms.
// inner == 1: one warp per row (lanes stride the row); inner > 1: one thread per row (adjacent threads = adjacent
// `inner` positions, so every load of the strided walk is coalesced across the warp).
__device__ __forceinline__ float norm_apply(float x, float mean, float inv, int mode) {
  return mode == 1 ? x - mean : (mode == 2 ? x * inv : (x - mean) * inv);
}

__global__ void instance_norm_kernel(const float* __restrict__ src, int64_t outer, int64_t cols, int64_t inner, int mode,
                                     int unbiased, float* __restrict__ dst) {
  const float denom = static_cast<float>(mode == 0 && unbiased ? cols - 1 : cols);
  if (inner == 1) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= outer) return;
    const float* x = src + row * cols;
    float mean = 0.f;
    if (mode != 2) {
      float s = 0.f;
      for (int64_t c = lane; c < cols; c += 32) s += x[c];
      mean = warp_sum(s) / static_cast<float>(cols);
    }
    float inv = 1.f;
    if (mode != 1) {
      float v = 0.f;
      for (int64_t c = lane; c < cols; c += 32) {
        const float d = x[c] - mean;
        v = fmaf(d, d, v);
      }
      inv = 1.0f / sqrtf(warp_sum(v) / denom);
    }
    for (int64_t c = lane; c < cols; c += 32) dst[row * cols + c] = norm_apply(x[c], mean, inv, mode);
    return;
  }
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= outer * inner) return;
  const int64_t o = row / inner, i = row - o * inner;
  const float* x = src + o * cols * inner + i;
  float* y = dst + o * cols * inner + i;
  float mean = 0.f;
  if (mode != 2) {
    float s = 0.f;
    for (int64_t c = 0; c < cols; ++c) s += x[c * inner];
    mean = s / static_cast<float>(cols);
  }
  float inv = 1.f;
  if (mode != 1) {
    float v = 0.f;
    for (int64_t c = 0; c < cols; ++c) {
      const float d = x[c * inner] - mean;
      v = fmaf(d, d, v);
    }
    inv = 1.0f / sqrtf(v / denom);
  }
  for (int64_t c = 0; c < cols; ++c) y[c * inner] = norm_apply(x[c * inner], mean, inv, mode);
}

// log1p(|X|) (Log1pAbsSTFT of padertorch, base of MVNLog1pAbsSTFT / Log1pAbsIPDSTFT, tssep/train/feature_extractor.py:83-168)
__global__ void log1p_abs_kernel(const float2* __restrict__ X, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float2 v = X[i];
    out[i] = log1pf(sqrtf(v.x * v.x + v.y * v.y));
  }
}

// inter-channel phase differences (tssep/train/feature_extractor.py:13-80): z = X[d] conj(X[second[d]]),
// cos = Re z / |z|, sin = Im z / |z|; X (lead, D, TF)
__global__ void ipd_kernel(const float2* __restrict__ X, int64_t lead, int D, int64_t TF, const int* __restrict__ second,
                           float* __restrict__ cos_out, float* __restrict__ sin_out) {
  const int64_t total = lead * D * TF;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = i % TF;
    const int64_t ld = i / TF;
    const int d = static_cast<int>(ld % D);
    const int64_t l = ld / D;
    const float2 a = X[i], b = X[(l * D + second[d]) * TF + e];
    const float re = a.x * b.x + a.y * b.y, im = a.y * b.x - a.x * b.y;
    const float inv = 1.0f / sqrtf(re * re + im * im);
    cos_out[i] = re * inv;
    sin_out[i] = im * inv;
  }
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_stft(const float* audio, int64_t n_signals, int64_t num_samples, const float* window,
               const float* twiddle, int size, int shift, int window_length, int fading, int64_t T, float* X,
               tssep_stream_t stream) {
  TSSEP_REQUIRE(audio && window && twiddle && X, "tssep_stft: null pointer");
  const int l2 = ilog2_exact(size);
  TSSEP_REQUIRE(l2 >= 3 && size <= 4096, "tssep_stft: size must be a power of two in [8, 4096], got %d", size);
  TSSEP_REQUIRE(window_length >= 1 && window_length <= size && shift >= 1 && shift <= window_length,
                "tssep_stft: need 1 <= shift <= window_length <= size (%d, %d, %d)", shift, window_length, size);
  TSSEP_REQUIRE(n_signals >= 0 && n_signals < 65536 && num_samples >= 0 && T >= 0, "tssep_stft: bad extent");
  if (n_signals == 0 || T == 0) return 0;
  const int fpb = 32;
  const int M = size / 2;
  const size_t smem = sizeof(float2) * (M + kWarps * 2 * padded_len(M)) + sizeof(float) * (window_length + (fpb - 1) * shift + window_length);
  TSSEP_REQUIRE(smem <= 227 * 1024, "tssep_stft: frame geometry does not fit shared memory");
  TSSEP_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>((T + fpb - 1) / fpb), static_cast<unsigned>(n_signals));
  stft_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      audio, num_samples, window, reinterpret_cast<const float2*>(twiddle), size, shift, window_length,
      fading ? window_length - shift : 0, T, reinterpret_cast<float2*>(X), fpb);
  return check_launch("tssep_stft");
}

int tssep_feature_stats(const float* X, int64_t n_items, int64_t x_item_stride, int64_t T, int F,
                        const float* mel_t, const int32_t* mel_lo, const int32_t* mel_hi, int n_mels,
                        uint32_t* absmax_key, uint32_t* maxdb_key, float* meldb, tssep_stream_t stream) {
  TSSEP_REQUIRE(X && absmax_key, "tssep_feature_stats: null pointer");
  TSSEP_REQUIRE(n_mels == 0 || (mel_t && mel_lo && mel_hi && maxdb_key && meldb),
                "tssep_feature_stats: mel tables missing");
  TSSEP_REQUIRE(n_items >= 0 && n_items < 65536 && F >= 1 && F <= 4097 && n_mels >= 0 && n_mels <= 256,
                "tssep_feature_stats: bad extent");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_items == 0) return 0;
  TSSEP_CUDA(cudaMemsetAsync(absmax_key, 0, sizeof(uint32_t) * n_items, s));
  if (n_mels > 0) TSSEP_CUDA(cudaMemsetAsync(maxdb_key, 0, sizeof(uint32_t) * n_items, s));
  if (T == 0) return 0;
  const int fpb = 32;
  const size_t smem = sizeof(float) * kWarps * F;
  dim3 grid(static_cast<unsigned>((T + fpb - 1) / fpb), static_cast<unsigned>(n_items));
  feature_stats_kernel<<<grid, kThreads, smem, s>>>(reinterpret_cast<const float2*>(X), x_item_stride, T, F, mel_t,
                                                    mel_lo, mel_hi, n_mels, absmax_key, maxdb_key, meldb, fpb);
  return check_launch("tssep_feature_stats");
}

int tssep_feature_write(const float* X, int64_t n_items, int64_t x_item_stride, int64_t T, int F,
                        const uint32_t* absmax_key, const uint32_t* maxdb_key, const float* meldb,
                        const float* dct, int n_mels, int n_mfcc, int with_log1p, float top_db,
                        int couple_batch, float* feat_f32, uint16_t* feat_bf16, int64_t ld_bf16,
                        tssep_stream_t stream) {
  TSSEP_REQUIRE(X && absmax_key, "tssep_feature_write: null pointer");
  TSSEP_REQUIRE(n_mfcc == 0 || (maxdb_key && meldb && dct && n_mels > 0), "tssep_feature_write: mfcc tables missing");
  TSSEP_REQUIRE(n_mfcc > 0 || with_log1p, "tssep_feature_write: empty feature");
  TSSEP_REQUIRE(feat_f32 || feat_bf16, "tssep_feature_write: no output");
  const int Din = n_mfcc + (with_log1p ? F : 0);
  TSSEP_REQUIRE(!feat_bf16 || ld_bf16 >= Din, "tssep_feature_write: ld_bf16 < feature size");
  if (n_items == 0 || T == 0) return 0;
  const int fpb = 32;
  const size_t smem = sizeof(float) * kWarps * (n_mels > 0 ? n_mels : 1);
  dim3 grid(static_cast<unsigned>((T + fpb - 1) / fpb), static_cast<unsigned>(n_items));
  feature_write_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(X), x_item_stride, n_items, T, F, absmax_key, maxdb_key, meldb, dct, n_mels,
      n_mfcc, with_log1p, top_db, couple_batch, feat_f32, reinterpret_cast<__nv_bfloat16*>(feat_bf16), ld_bf16, fpb);
  return check_launch("tssep_feature_write");
}

int tssep_cast_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, uint16_t* dst, int64_t ld_dst,
                    tssep_stream_t stream) {
  TSSEP_REQUIRE(src && dst && ld_dst >= cols && ld_src >= cols, "tssep_cast_bf16: bad arguments");
  if (rows == 0) return 0;
  const int64_t total = rows * ld_dst;
  const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 16));
  cast_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, rows, cols, ld_src,
                                                                           reinterpret_cast<__nv_bfloat16*>(dst), ld_dst);
  return check_launch("tssep_cast_bf16");
}

int tssep_log1p_abs(const float* X, int64_t n, float* out, tssep_stream_t stream) {
  TSSEP_REQUIRE(X && out && n >= 0, "tssep_log1p_abs: bad arguments");
  if (n == 0) return 0;
  const int blocks = static_cast<int>(imin64((n + 255) / 256, 148 * 16));
  log1p_abs_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(X), n, out);
  return check_launch("tssep_log1p_abs");
}

int tssep_ipd(const float* X, int64_t lead, int D, int64_t TF, const int32_t* second_channel, float* cos_out, float* sin_out,
              tssep_stream_t stream) {
  TSSEP_REQUIRE(X && second_channel && cos_out && sin_out && D >= 2 && lead >= 0 && TF >= 0, "tssep_ipd: bad arguments");
  if (lead == 0 || TF == 0) return 0;
  const int64_t total = lead * D * TF;
  const int blocks = static_cast<int>(imin64((total + 255) / 256, 148 * 16));
  ipd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(X), lead, D, TF, second_channel,
                                                                    cos_out, sin_out);
  return check_launch("tssep_ipd");
}

int tssep_instance_norm(const float* src, int64_t outer, int64_t cols, int64_t inner, int mode, int unbiased, float* dst,
                        tssep_stream_t stream) {
  TSSEP_REQUIRE(src && dst && cols >= 1 && inner >= 1 && outer >= 0, "tssep_instance_norm: bad arguments");
  TSSEP_REQUIRE(mode >= 0 && mode <= 2, "tssep_instance_norm: mode must be 0 (standardise), 1 (centre) or 2 (rms)");
  if (outer == 0) return 0;
  const int64_t rows = outer * inner;
  const int64_t blocks = inner == 1 ? (rows + 7) / 8 : (rows + 255) / 256;
  TSSEP_REQUIRE(blocks < (1ll << 31), "tssep_instance_norm: too many rows");
  instance_norm_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, outer, cols, inner,
                                                                                                    mode, unbiased, dst);
  return check_launch("tssep_instance_norm");
}

}  // extern "C"
