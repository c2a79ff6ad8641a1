// Warp-cooperative Stockham FFT over shared memory (radix 8 / 4 / 2 passes), plus the
// real-FFT pack / unpack steps used by the STFT and iSTFT kernels.
//
// A real transform of length S is computed as a complex transform of length
// M = S/2 on z[n] = x[2n] + i x[2n+1].  `tw` is the table exp(-2*pi*i*k/S),
// k < S/2.  Buffers are indexed through padi() (one padding element every 16)
// so that the strided writes of the first passes do not serialise on banks.
#pragma once
#include "common.cuh"

namespace tssep {

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ int padded_len(int m) { return m + (m >> 4) + 1; }

__host__ __device__ __forceinline__ int fft_num_passes(int M) {
  int n = 0;
  for (int ns = 1; ns < M; ns *= (M / ns >= 8 ? 8 : M / ns)) ++n;
  return n;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 crot(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// exp(-+ 2 pi i idx / S) from the half table
template <bool INV>
__device__ __forceinline__ float2 twiddle(const float2* tw, int idx, int S) {
  idx &= (S - 1);
  const int half = S >> 1;
  float2 w = tw[idx & (half - 1)];
  if (idx >= half) w = make_float2(-w.x, -w.y);
  if (INV) w.y = -w.y;
  return w;
}

template <bool INV>
__device__ __forceinline__ void dft2(float2* v) {
  const float2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <bool INV>
__device__ __forceinline__ void dft4(float2* v) {
  const float2 b0 = cadd(v[0], v[2]), b2 = csub(v[0], v[2]);
  const float2 b1 = cadd(v[1], v[3]), b3 = crot<INV>(csub(v[1], v[3]));
  v[0] = cadd(b0, b1);
  v[1] = cadd(b2, b3);
  v[2] = csub(b0, b1);
  v[3] = csub(b2, b3);
}
template <bool INV>
__device__ __forceinline__ void dft8(float2* v) {
  constexpr float h = 0.70710678118654752440f;
  const float2 w1 = INV ? make_float2(h, h) : make_float2(h, -h);    // exp(-+ i pi/4)
  const float2 w3 = INV ? make_float2(-h, h) : make_float2(-h, -h);  // exp(-+ 3 i pi/4)
  const float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  const float2 a1 = cadd(v[1], v[5]), a5 = cmul(csub(v[1], v[5]), w1);
  const float2 a2 = cadd(v[2], v[6]), a6 = crot<INV>(csub(v[2], v[6]));
  const float2 a3 = cadd(v[3], v[7]), a7 = cmul(csub(v[3], v[7]), w3);
  const float2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = crot<INV>(csub(a1, a3));
  const float2 c0 = cadd(a4, a6), c2 = csub(a4, a6), c1 = cadd(a5, a7), c3 = crot<INV>(csub(a5, a7));
  v[0] = cadd(b0, b1);
  v[1] = cadd(c0, c1);
  v[2] = cadd(b2, b3);
  v[3] = cadd(c2, c3);
  v[4] = csub(b0, b1);
  v[5] = csub(c0, c1);
  v[6] = csub(b2, b3);
  v[7] = csub(c2, c3);
}

template <bool INV, int R>
__device__ __forceinline__ void stockham_pass(const float2* src, float2* dst, int M, int Ns, const float2* tw, int S,
                                              int lane) {
  const int nb = M / R;
  const int tstep = S / (Ns * R);
  for (int j = lane; j < nb; j += 32) {
    const int k = j & (Ns - 1);
    float2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = src[padi(j + t * nb)];
    if (Ns > 1) {
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul(v[t], twiddle<INV>(tw, t * k * tstep, S));
    }
    if (R == 8) dft8<INV>(v);
    else if (R == 4) dft4<INV>(v);
    else dft2<INV>(v);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) dst[padi(base + t * Ns)] = v[t];
  }
  __syncwarp();
}

// Complex FFT of length M (power of two >= 2) from buffer `a` (natural order, padi-indexed);
// `b` is scratch of the same size.  Unnormalised.  Returns the buffer holding the natural-order
// result (a for an even number of passes, b for an odd number: see fft_num_passes).
template <bool INV>
__device__ __forceinline__ float2* warp_fft(float2* a, float2* b, int M, const float2* tw, int S, int lane) {
  float2* src = a;
  float2* dst = b;
  int Ns = 1;
  while (Ns < M) {
    const int rem = M / Ns;
    if (rem >= 8) {
      stockham_pass<INV, 8>(src, dst, M, Ns, tw, S, lane);
      Ns *= 8;
    } else if (rem == 4) {
      stockham_pass<INV, 4>(src, dst, M, Ns, tw, S, lane);
      Ns *= 4;
    } else {
      stockham_pass<INV, 2>(src, dst, M, Ns, tw, S, lane);
      Ns *= 2;
    }
    float2* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// After the forward half-length FFT Z (natural order, padi-indexed, length M): X[k], k in [0, M].
__device__ __forceinline__ float2 rfft_unpack(const float2* Z, int k, int M, const float2* tw) {
  const float2 zk = Z[padi(k & (M - 1))];
  float2 zm = Z[padi((M - k) & (M - 1))];
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

// Same with the table value w = exp(-2 pi i k / S) passed in.
__device__ __forceinline__ float2 irfft_pack_w(float2 xk, float2 xm, float2 w) {
  xm.y = -xm.y;
  const float ex = 0.5f * (xk.x + xm.x), ey = 0.5f * (xk.y + xm.y);
  const float dx = 0.5f * (xk.x - xm.x), dy = 0.5f * (xk.y - xm.y);
  w.y = -w.y;  // conj
  const float ox = dx * w.x - dy * w.y, oy = dx * w.y + dy * w.x;
  return make_float2(ex - oy, ey + ox);  // E + i*O
}

}  // namespace tssep
