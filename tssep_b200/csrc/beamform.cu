// Mask-based MVDR beamformer (Souden formulation): TorchBF.__call__ of the reference
// (tssep/train/enhancer.py:140-283, "mvdr_souden"), the multi-channel enhancer the LibriCSS pipeline runs on the
// TS-SEP masks.  Three kernels:
//   psd     Phi[z, plane, f] = sum_t w_plane[t, f] Y[z, :, t, f] Y[z, :, t, f]^H   (enhancer.py:237-262)
//           planes = the K * nmask masks, plus the all-ones weight when nmask == 1 (interference = Phi_YY - target,
//           i.e. the einsum with 1 - m); float64 accumulation of float32 products per block, double atomics across time
//           chunks; only the upper triangle is stored (Hermitian);
//   solve   per (speaker, frequency): phi = interference^-1 target by Gaussian elimination with partial pivoting in
//           complex float64, w = phi[:, ref] / max(Re trace(phi), eps)                 (enhancer.py:263-268)
//   apply   enh[z, k, t, f] = sum_d conj(w[z, k, f, d]) Y[z, d, t, f]  (* max(mask, masking_eps))   (enhancer.py:269-281)
// HBM-bound: psd reads Y and the masks once, apply reads Y once and writes K planes.
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

constexpr int kBfMaxD = 8;
constexpr int kBfChunk = 1024;  // frames per block of the PSD kernel

__host__ __device__ constexpr int tri(int D) { return D * (D + 1) / 2; }

// thread (f lane, plane warp): upper triangle of sum_t w Y Y^H over one chunk of frames
template <int D>
__global__ void __launch_bounds__(32 * 17) bf_psd_kernel(const float2* __restrict__ Y, const float* __restrict__ mask, int K, int nmask,
                                                      int64_t T, int F, int planes, double* __restrict__ psd) {
  const int fl = threadIdx.x, plane = threadIdx.y;
  const int f = blockIdx.x * 32 + fl;
  const int64_t z = blockIdx.z;
  const int64_t t0 = static_cast<int64_t>(blockIdx.y) * kBfChunk, t1 = imin64(T, t0 + kBfChunk);
  if (f >= F) return;
  const bool ones = plane >= K * nmask;  // the extra all-ones plane (nmask == 1)
  const float* m = ones ? nullptr : mask + ((z * K * nmask + plane) * T) * F + f;
  const float2* y = Y + (z * D * T) * F + f;
  double acc_re[D][D], acc_im[D][D];  // upper triangle used (every index is a compile-time constant once unrolled)
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) acc_re[a][b] = acc_im[a][b] = 0.0;
  for (int64_t t = t0; t < t1; ++t) {
    float2 v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = y[(d * T + t) * F];
    const float w = ones ? 1.0f : m[t * F];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const float wr = w * v[a].x, wi = w * v[a].y;
#pragma unroll
      for (int b = 0; b < D; ++b) {  // w * Y_a * conj(Y_b)
        if (b >= a) {
          acc_re[a][b] += static_cast<double>(wr * v[b].x + wi * v[b].y);
          acc_im[a][b] += static_cast<double>(wi * v[b].x - wr * v[b].y);
        }
      }
    }
  }
  double* o = psd + (((z * planes + plane) * F + f) * tri(D)) * 2;
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b)
      if (b >= a) {
        const int i = a * D - a * (a - 1) / 2 + (b - a);
        atomicAdd(o + 2 * i, acc_re[a][b]);
        atomicAdd(o + 2 * i + 1, acc_im[a][b]);
      }
}

struct cd {
  double x, y;
};
__device__ __forceinline__ cd cmul_d(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd csub_d(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd cdiv_d(cd a, cd b) {
  const double s = 1.0 / (b.x * b.x + b.y * b.y);
  return {(a.x * b.x + a.y * b.y) * s, (a.y * b.x - a.x * b.y) * s};
}

template <int D>
__global__ void bf_mvdr_kernel(const double* __restrict__ psd, int64_t Z, int K, int nmask, int F, int planes, int ref,
                               double eps, float2* __restrict__ w) {
  const int64_t total = Z * K * F;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(idx % F);
    const int k = static_cast<int>((idx / F) % K);
    const int64_t z = idx / (static_cast<int64_t>(F) * K);
    const double* pt = psd + (((z * planes + k * nmask) * F + f) * tri(D)) * 2;
    const double* pi = psd + (((z * planes + (nmask == 2 ? k * 2 + 1 : K)) * F + f) * tri(D)) * 2;
    cd A[D][2 * D];  // [interference | target]
    int i = 0;
    for (int a = 0; a < D; ++a)
      for (int b = a; b < D; ++b, ++i) {
        cd tg = {pt[2 * i], pt[2 * i + 1]}, in = {pi[2 * i], pi[2 * i + 1]};
        if (nmask == 1) in = csub_d(in, tg);  // sum (1 - m) Y Y^H = Phi_YY - target
        A[a][D + b] = tg;
        A[a][b] = in;
        if (b != a) {
          A[b][D + a] = {tg.x, -tg.y};
          A[b][a] = {in.x, -in.y};
        }
      }
    // Gaussian elimination with partial pivoting (what torch.linalg.solve / LAPACK gesv does)
    for (int c = 0; c < D; ++c) {
      int p = c;
      double best = A[c][c].x * A[c][c].x + A[c][c].y * A[c][c].y;
      for (int r = c + 1; r < D; ++r) {
        const double v = A[r][c].x * A[r][c].x + A[r][c].y * A[r][c].y;
        if (v > best) {
          best = v;
          p = r;
        }
      }
      if (p != c)
        for (int j = 0; j < 2 * D; ++j) {
          const cd tmp = A[c][j];
          A[c][j] = A[p][j];
          A[p][j] = tmp;
        }
      for (int r = c + 1; r < D; ++r) {
        const cd l = cdiv_d(A[r][c], A[c][c]);
        for (int j = c; j < 2 * D; ++j) A[r][j] = csub_d(A[r][j], cmul_d(l, A[c][j]));
      }
    }
    cd X[D][D];
    for (int j = 0; j < D; ++j)
      for (int r = D - 1; r >= 0; --r) {
        cd s = A[r][D + j];
        for (int c = r + 1; c < D; ++c) s = csub_d(s, cmul_d(A[r][c], X[c][j]));
        X[r][j] = cdiv_d(s, A[r][r]);
      }
    double lambda = 0.0;
    for (int d = 0; d < D; ++d) lambda += X[d][d].x;
    const double inv = 1.0 / fmax(lambda, eps);
    float2* o = w + idx * D;
    for (int d = 0; d < D; ++d) o[d] = make_float2(static_cast<float>(X[d][ref].x * inv), static_cast<float>(X[d][ref].y * inv));
  }
}

template <int D>
__global__ void bf_apply_kernel(const float2* __restrict__ Y, const float2* __restrict__ w, const float* __restrict__ mask, int K,
                                int nmask, int64_t T, int F, float masking_eps, float2* __restrict__ out) {
  const int64_t z = blockIdx.z;
  const int64_t total = T * F;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i % F);
    float2 v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = Y[(z * D + d) * total + i];
    for (int k = 0; k < K; ++k) {
      const float2* wk = w + ((z * K + k) * F + f) * D;
      float re = 0.f, im = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {  // conj(w) * y
        const float2 c = __ldg(wk + d);
        re += c.x * v[d].x + c.y * v[d].y;
        im += c.x * v[d].y - c.y * v[d].x;
      }
      if (mask != nullptr) {
        const float g = fmaxf(mask[((z * K + k) * nmask) * total + i], masking_eps);
        re *= g;
        im *= g;
      }
      out[(z * K + k) * total + i] = make_float2(re, im);
    }
  }
}

}  // namespace tssep

using namespace tssep;

#define TSSEP_BF_DISPATCH(D_, CALL)                                                    \
  switch (D_) {                                                                        \
    case 1: { constexpr int DD = 1; CALL; } break;                                     \
    case 2: { constexpr int DD = 2; CALL; } break;                                     \
    case 3: { constexpr int DD = 3; CALL; } break;                                     \
    case 4: { constexpr int DD = 4; CALL; } break;                                     \
    case 5: { constexpr int DD = 5; CALL; } break;                                     \
    case 6: { constexpr int DD = 6; CALL; } break;                                     \
    case 7: { constexpr int DD = 7; CALL; } break;                                     \
    case 8: { constexpr int DD = 8; CALL; } break;                                     \
    default: set_error("beamformer: %d channels (supported: 1..8)", D_); return -1;    \
  }

extern "C" {

int tssep_bf_psd(const float* Y, const float* mask, int64_t Z, int K, int nmask, int D, int64_t T, int F, double* psd,
                 tssep_stream_t stream) {
  TSSEP_REQUIRE(Y && mask && psd, "tssep_bf_psd: null pointer");
  TSSEP_REQUIRE(K >= 1 && (nmask == 1 || nmask == 2) && D >= 1 && D <= kBfMaxD && F >= 1 && Z >= 0 && Z < 65536 && T >= 0,
                "tssep_bf_psd: bad extent");
  const int planes = K * nmask + (nmask == 1 ? 1 : 0);
  TSSEP_REQUIRE(planes <= 17, "tssep_bf_psd: at most 17 weight planes per launch (16 speakers, or 8 with two masks; got %d)", planes);
  if (Z == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TSSEP_CUDA(cudaMemsetAsync(psd, 0, sizeof(double) * 2 * Z * planes * F * tri(D), s));
  if (T == 0) return 0;
  dim3 grid((F + 31) / 32, static_cast<unsigned>((T + kBfChunk - 1) / kBfChunk), static_cast<unsigned>(Z)), block(32, planes);
  TSSEP_BF_DISPATCH(D, (bf_psd_kernel<DD><<<grid, block, 0, s>>>(reinterpret_cast<const float2*>(Y), mask, K, nmask, T, F, planes, psd)));
  return check_launch("tssep_bf_psd");
}

int tssep_bf_mvdr_souden(const double* psd, int64_t Z, int K, int nmask, int D, int F, int reference_channel, double eps,
                         float* w, tssep_stream_t stream) {
  TSSEP_REQUIRE(psd && w, "tssep_bf_mvdr_souden: null pointer");
  TSSEP_REQUIRE(K >= 1 && (nmask == 1 || nmask == 2) && D >= 1 && D <= kBfMaxD && reference_channel >= 0 && reference_channel < D,
                "tssep_bf_mvdr_souden: bad arguments");
  if (Z == 0) return 0;
  const int planes = K * nmask + (nmask == 1 ? 1 : 0);
  const int64_t total = Z * K * F;
  const int blocks = static_cast<int>(imin64((total + 63) / 64, 148 * 16));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TSSEP_BF_DISPATCH(D, (bf_mvdr_kernel<DD><<<blocks, 64, 0, s>>>(psd, Z, K, nmask, F, planes, reference_channel, eps,
                                                                  reinterpret_cast<float2*>(w))));
  return check_launch("tssep_bf_mvdr_souden");
}

int tssep_bf_apply(const float* Y, const float* w, const float* mask, int64_t Z, int K, int nmask, int D, int64_t T, int F,
                   float masking_eps, float* out, tssep_stream_t stream) {
  TSSEP_REQUIRE(Y && w && out, "tssep_bf_apply: null pointer");
  TSSEP_REQUIRE(K >= 1 && D >= 1 && D <= kBfMaxD && Z >= 0 && Z < 65536, "tssep_bf_apply: bad extent");
  if (Z == 0 || T == 0) return 0;
  const int64_t total = T * F;
  dim3 grid(static_cast<unsigned>(imin64((total + 255) / 256, 148 * 8)), 1, static_cast<unsigned>(Z));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TSSEP_BF_DISPATCH(D, (bf_apply_kernel<DD><<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(Y), reinterpret_cast<const float2*>(w),
                                                                  mask, K, nmask, T, F, masking_eps, reinterpret_cast<float2*>(out))));
  return check_launch("tssep_bf_apply");
}

}  // extern "C"
