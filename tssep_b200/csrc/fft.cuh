// Warp-cooperative in-place radix-2 FFT over shared memory, plus the real-FFT
// pack / unpack steps used by the STFT and iSTFT kernels.
//
// A real transform of length S is computed as a complex transform of length
// M = S/2 on z[n] = x[2n] + i x[2n+1].  `tw` is the table exp(-2*pi*i*k/S),
// k < S/2, so the half-length twiddles are tw[2k].
#pragma once
#include "common.cuh"

namespace tssep {

__device__ __forceinline__ int bitrev(int i, int bits) { return static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - bits)); }

// s holds M = 1 << log2m complex values in bit-reversed order; result in natural order.
template <bool INVERSE>
__device__ __forceinline__ void warp_fft_inplace(float2* s, int log2m, const float2* tw, int S, int lane) {
  const int half_m = 1 << (log2m - 1);
  for (int st = 0; st < log2m; ++st) {
    const int half = 1 << st;
    const int tstep = S >> (st + 1);
    for (int j = lane; j < half_m; j += 32) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> st) << (st + 1)) + pos;
      const int i1 = i0 + half;
      float2 w = tw[pos * tstep];
      if (INVERSE) w.y = -w.y;
      const float2 a = s[i0];
      const float2 b = s[i1];
      const float bx = b.x * w.x - b.y * w.y;
      const float by = b.x * w.y + b.y * w.x;
      s[i0] = make_float2(a.x + bx, a.y + by);
      s[i1] = make_float2(a.x - bx, a.y - by);
    }
    __syncwarp();
  }
}

// After the forward half-length FFT Z (natural order, length M): X[k], k in [0, M].
__device__ __forceinline__ float2 rfft_unpack(const float2* Z, int k, int M, const float2* tw) {
  const float2 zk = Z[k & (M - 1)];
  float2 zm = Z[(M - k) & (M - 1)];
  zm.y = -zm.y;
  const float2 w = (k < M) ? tw[k] : make_float2(-1.f, 0.f);
  const float sx = zk.x + zm.x, sy = zk.y + zm.y;
  const float dx = zk.x - zm.x, dy = zk.y - zm.y;
  const float wx = w.x * dx - w.y * dy, wy = w.x * dy + w.y * dx;
  return make_float2(0.5f * (sx + wy), 0.5f * (sy - wx));
}

// Before the inverse half-length FFT: Z[k] from the half spectrum Xk = X[k], Xm = X[M-k].
// (The imaginary parts of X[0] and X[M] must already be dropped by the caller.)
__device__ __forceinline__ float2 irfft_pack(float2 xk, float2 xm, int k, const float2* tw) {
  xm.y = -xm.y;
  const float ex = 0.5f * (xk.x + xm.x), ey = 0.5f * (xk.y + xm.y);
  const float dx = 0.5f * (xk.x - xm.x), dy = 0.5f * (xk.y - xm.y);
  float2 w = tw[k];
  w.y = -w.y;  // conj
  const float ox = dx * w.x - dy * w.y, oy = dx * w.y + dy * w.x;
  return make_float2(ex - oy, ey + ox);  // E + i*O
}

}  // namespace tssep
