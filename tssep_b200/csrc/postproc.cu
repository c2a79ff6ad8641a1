// Diarization post-processing: frame activity, median smoothing + threshold,
// run-length segment extraction with warp-level scans.
//
// The reference repository has no implementation of this stage (it lives in
// fgnt/tssep_data); the in-repo anchors are the frequency mean used by the VAD
// loss (tssep/train/loss.py:343) and the run-length / index mapping of
// tssep/util/utils.py:11-129.  The spec implemented here is the one written
// down in oracle/tssep_oracle.py::diarize_reference (parity unpinned).
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

__global__ void activity_kernel(const float* __restrict__ mask, int64_t rows, int F, int64_t pitch, float* __restrict__ act) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t r = blockIdx.x * static_cast<int64_t>(wpb) + (threadIdx.x >> 5); r < rows;
       r += static_cast<int64_t>(gridDim.x) * wpb) {
    const float* m = mask + r * pitch;
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s += m[f];
    s = warp_sum(s);
    if (lane == 0) act[r] = s / static_cast<float>(F);
  }
}

__global__ void median_threshold_kernel(const float* __restrict__ act, int64_t n, int64_t T, int width, float thr,
                                        float* __restrict__ smooth, uint8_t* __restrict__ active) {
  const int h = width >> 1;
  const int64_t total = n * T;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i % T;
    const float* a = act + (i - t);
    float v[63];
    for (int j = 0; j < width; ++j) {
      int64_t tt = t - h + j;
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      v[j] = a[tt];
    }
    float med = v[0];
    for (int j = 0; j < width; ++j) {
      int less = 0, leq = 0;
      for (int q = 0; q < width; ++q) {
        less += v[q] < v[j];
        leq += v[q] <= v[j];
      }
      if (less <= h && leq > h) med = v[j];
    }
    if (smooth) smooth[i] = med;
    if (active) active[i] = med > thr ? 1 : 0;
  }
}

__device__ __forceinline__ int frame_to_sample(int64_t frame, int wl, int R, int pad, int64_t num_samples) {
  int64_t s = frame * R - pad + wl / 2 - R / 2;
  s = s < 0 ? 0 : s;
  if (num_samples >= 0 && s > num_samples) s = num_samples;
  return static_cast<int>(s);
}

// Index mapping of stft_vad / istft_vad (tssep/util/utils.py:11-129).  The reference delegates it to paderbox
// (module_stft.sample_index_to_stft_frame_index / stft_frame_index_to_sample_index, paderbox==0.0.8, absent here);
// restated from SURVEY.md App. A:
//   frame(s)  = pad_frames + (s < ceil(wl/2) ? 0 : (s - ceil(wl/2)) / shift + 1),  pad_frames = ceil((wl-shift)/shift) | 0
//   first(f)  = smallest sample mapped to a frame >= f;   last(f) = largest sample mapped to frame f = first(f+1) - 1
__device__ __forceinline__ int64_t vad_first_sample(int64_t frame, int wl, int R, int pad_frames) {
  const int64_t f = frame - pad_frames;
  return f <= 0 ? 0 : (wl + 1) / 2 + (f - 1) * static_cast<int64_t>(R);
}

// frame f is covered by a run of active samples [s, e) with frame(s) <= f < frame(e)  <=>  the sample just below
// first(f + 1) is active: a gather, no scan
__global__ void stft_vad_kernel(const uint8_t* __restrict__ vad, int64_t n, int64_t N, int wl, int R, int pad_frames,
                                int64_t T, uint8_t* __restrict__ out) {
  const int64_t total = n * T;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t f = i % T, sig = i / T;
    const int64_t a = vad_first_sample(f + 1, wl, R, pad_frames);
    out[i] = (a >= 1 && a <= N && vad[sig * N + a - 1]) ? 1 : 0;
  }
}

// one block per signal; starts and ends of runs are compacted with ballot + prefix sums
__global__ void __launch_bounds__(256)
segments_kernel(const uint8_t* __restrict__ active, int64_t T, int wl, int R, int pad, int64_t num_samples,
                int* __restrict__ segments, int* __restrict__ counts, int max_segments, int index_mode) {
  __shared__ int warp_s[8], warp_e[8];
  __shared__ int base_s, base_e;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t sig = blockIdx.x;
  const uint8_t* a = active + sig * T;
  int* seg = segments + sig * max_segments * 2;
  if (threadIdx.x == 0) {
    base_s = 0;
    base_e = 0;
  }
  __syncthreads();
  for (int64_t c0 = 0; c0 < T; c0 += 256) {
    const int64_t t = c0 + threadIdx.x;
    const bool cur = t < T && a[t];
    const bool prev = t > 0 && t < T && a[t - 1];
    const bool next = t + 1 < T && a[t + 1];
    const bool is_start = cur && !prev;
    const bool is_end = cur && !next;  // run ends after frame t
    const unsigned ms = __ballot_sync(0xffffffffu, is_start), me = __ballot_sync(0xffffffffu, is_end);
    if (lane == 0) {
      warp_s[warp] = __popc(ms);
      warp_e[warp] = __popc(me);
    }
    __syncthreads();
    int off_s = base_s, off_e = base_e;
    for (int w = 0; w < warp; ++w) {
      off_s += warp_s[w];
      off_e += warp_e[w];
    }
    const unsigned lt = (1u << lane) - 1;
    if (is_start) {
      const int i = off_s + __popc(ms & lt);
      if (i < max_segments) {
        int64_t v = index_mode ? vad_first_sample(t, wl, R, (pad + R - 1) / R) : frame_to_sample(t, wl, R, pad, num_samples);
        if (num_samples >= 0 && v > num_samples) v = num_samples;
        seg[2 * i] = static_cast<int>(v);
      }
    }
    if (is_end) {
      const int i = off_e + __popc(me & lt);
      if (i < max_segments) {
        // istft_vad maps the exclusive frame end t + 1 with mode 'last' (utils.py:116-122): last(f) = first(f + 1) - 1
        int64_t v = index_mode ? vad_first_sample(t + 2, wl, R, (pad + R - 1) / R) - 1 : frame_to_sample(t + 1, wl, R, pad, num_samples);
        v = v < 0 ? 0 : v;
        if (num_samples >= 0 && v > num_samples) v = num_samples;
        seg[2 * i + 1] = static_cast<int>(v);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int ts = 0, te = 0;
      for (int w = 0; w < 8; ++w) {
        ts += warp_s[w];
        te += warp_e[w];
      }
      base_s += ts;
      base_e += te;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[sig] = base_s;
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int tssep_activity(const float* mask, int64_t n, int64_t T, int F, int64_t mask_pitch, float* activity, tssep_stream_t stream) {
  TSSEP_REQUIRE(mask && activity && F >= 1 && (mask_pitch == 0 || mask_pitch >= F), "tssep_activity: bad arguments");
  const int64_t rows = n * T;
  if (rows == 0) return 0;
  const int blocks = static_cast<int>(imin64((rows + 7) / 8, 148 * 32));
  activity_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, rows, F, mask_pitch > 0 ? mask_pitch : F, activity);
  return check_launch("tssep_activity");
}

int tssep_median_threshold(const float* activity, int64_t n, int64_t T, int width, float threshold, float* smooth,
                           uint8_t* active, tssep_stream_t stream) {
  TSSEP_REQUIRE(activity && (smooth || active), "tssep_median_threshold: null pointer");
  TSSEP_REQUIRE(width >= 1 && width <= 63 && (width & 1), "tssep_median_threshold: width must be odd and in [1, 63]");
  const int64_t total = n * T;
  if (total == 0) return 0;
  const int blocks = static_cast<int>(imin64((total + 127) / 128, 148 * 32));
  median_threshold_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(activity, n, T, width, threshold,
                                                                                 smooth, active);
  return check_launch("tssep_median_threshold");
}

int tssep_stft_vad(const uint8_t* vad, int64_t n, int64_t num_samples, int window_length, int shift, int fading, int64_t T,
                   uint8_t* frames, tssep_stream_t stream) {
  TSSEP_REQUIRE(vad && frames, "tssep_stft_vad: null pointer");
  TSSEP_REQUIRE(shift >= 1 && window_length >= shift && n >= 0 && num_samples >= 0 && T >= 0, "tssep_stft_vad: bad extent");
  if (n == 0 || T == 0) return 0;
  const int blocks = static_cast<int>(imin64((n * T + 255) / 256, 148 * 16));
  stft_vad_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      vad, n, num_samples, window_length, shift, fading ? (window_length - shift + shift - 1) / shift : 0, T, frames);
  return check_launch("tssep_stft_vad");
}

int tssep_segments(const uint8_t* active, int64_t n, int64_t T, int window_length, int shift, int fading,
                   int64_t num_samples, int32_t* segments, int32_t* counts, int max_segments, int index_mode,
                   tssep_stream_t stream) {
  TSSEP_REQUIRE(active && segments && counts && max_segments >= 1, "tssep_segments: bad arguments");
  TSSEP_REQUIRE(index_mode == 0 || index_mode == 1, "tssep_segments: index_mode must be 0 (window centre) or 1 (istft_vad)");
  TSSEP_REQUIRE(shift >= 1 && window_length >= shift, "tssep_segments: bad frame geometry");
  if (n == 0) return 0;
  segments_kernel<<<static_cast<unsigned>(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      active, T, window_length, shift, fading ? window_length - shift : 0, num_samples, segments, counts, max_segments,
      index_mode);
  return check_launch("tssep_segments");
}

}  // extern "C"
