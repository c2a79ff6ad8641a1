// Weighted prediction error (WPE) dereverberation, the optional pre-processor in front of the STFT features
// (tssep/train/enhancer.py:292-367: WPE / ChannelWiseWPE call nara_wpe.wpe.wpe_v8 -- a third-party package that is
// not in the reference tree; this file restates its published algorithm, see oracle/tssep_oracle.py::wpe).
//
// Per frequency f (all independent), D channels, T frames, K taps, delay Delta, with the stacked delayed observation
//     Ytilde[(tau, d), t] = Y[d, t - Delta - tau]            (zero before the first frame), tau = 0..K-1
// every iteration does
//     lambda_t = mean_d |X[d, t]|^2  (optionally averaged over +-psd_context frames), clamped below at 1e-10 max_t lambda
//     R = sum_t Ytilde_t Ytilde_t^H / lambda_t   (DK x DK),   P = sum_t Ytilde_t Y_t^H / lambda_t   (DK x D)
//     G = R^-1 P,    X = Y - G^H Ytilde
// starting from X = Y.
//
// Layout: the (D, T, F) input is transposed once to (F, D, T) so that a CTA streams the frames of one frequency with
// coalesced loads.  wpe_stats_kernel: one CTA per (frequency, chunk of frames); the chunk (+ its K + Delta - 1 frames
// of history) sits in shared memory, every thread owns 4x4 register tiles of the upper triangle of R (and of P),
// accumulates them in f32 over the chunk and adds them to f64 accumulators in HBM (the sum over 37 503 frames of a
// 10-minute meeting is where f32 would lose the small eigenvalues the solve depends on).  wpe_solve_kernel: one CTA
// per frequency, Gauss-Jordan with partial pivoting in f64 on [R | P] in shared memory.  wpe_apply_kernel: one thread
// per frame, G in shared memory; it also produces lambda for the next iteration.
#include "../../include/tssep_b200.h"
#include "common.cuh"

namespace tssep {

constexpr int kWpeChunk = 256;     // frames per CTA of the statistics / apply kernels
constexpr int kWpeMaxD = 8;

// (A, B, C) -> (C, A, B) for 8-byte elements: in[(a*B + b)*C + c] -> out[(c*A + a)*B + b]; tiles of 32 b x 32 c
__global__ void wpe_to_fdt_kernel(const float2* __restrict__ in, int A, int B, int C, float2* __restrict__ out) {
  __shared__ float2 tile[32][33];
  const int a = blockIdx.z;
  const int b0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, c = c0 + threadIdx.x;
    if (b < B && c < C) tile[i][threadIdx.x] = in[(static_cast<int64_t>(a) * B + b) * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, b = b0 + threadIdx.x;
    if (b < B && c < C) out[(static_cast<int64_t>(c) * A + a) * B + b] = tile[threadIdx.x][i];
  }
}
// the inverse: in (C, A, B) -> out (A, B, C)
__global__ void wpe_from_fdt_kernel(const float2* __restrict__ in, int A, int B, int C, float2* __restrict__ out) {
  __shared__ float2 tile[32][33];
  const int a = blockIdx.z;
  const int b0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, b = b0 + threadIdx.x;
    if (b < B && c < C) tile[i][threadIdx.x] = in[(static_cast<int64_t>(c) * A + a) * B + b];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, c = c0 + threadIdx.x;
    if (b < B && c < C) out[(static_cast<int64_t>(a) * B + b) * C + c] = tile[threadIdx.x][i];
  }
}

__device__ __forceinline__ void block_max_to(float v, unsigned int* dst) {
  // v >= 0: the bit pattern of a non-negative float orders like the float
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(dst, __float_as_uint(v));
}

// lambda[f, t] = mean_d |X[f, d, t]|^2, and its maximum over t (pmax, as float bits) unless a smoothing pass follows
__global__ void wpe_power_kernel(const float2* __restrict__ X, int D, int T, float* __restrict__ power,
                                 unsigned int* __restrict__ pmax) {
  const int f = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  float p = 0.f;
  if (t < T) {
    for (int d = 0; d < D; ++d) {
      const float2 v = X[(static_cast<int64_t>(f) * D + d) * T + t];
      p += v.x * v.x + v.y * v.y;
    }
    p /= static_cast<float>(D);
    power[static_cast<int64_t>(f) * T + t] = p;
  }
  if (pmax) block_max_to(p, pmax + f);
}

// count-normalised mean over [t - ctx, t + ctx] (nara_wpe's window_mean) and the maximum of the result
__global__ void wpe_smooth_kernel(const float* __restrict__ power, int T, int ctx, float* __restrict__ out,
                                  unsigned int* __restrict__ pmax) {
  const int f = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  float p = 0.f;
  if (t < T) {
    const int lo = max(0, t - ctx), hi = min(T - 1, t + ctx);
    double s = 0.0;
    for (int u = lo; u <= hi; ++u) s += power[static_cast<int64_t>(f) * T + u];
    p = static_cast<float>(s / (hi - lo + 1));
    out[static_cast<int64_t>(f) * T + t] = p;
  }
  block_max_to(p, pmax + f);
}

struct WpeDims {
  int D, T, F, taps, delay, DK, t_first;  // t_first: first frame that enters the statistics (0, or delay + taps - 1)
};

// shared memory of the stats / apply kernels: D x W frames (W = chunk + history), then the chunk's 1 / lambda
__device__ __forceinline__ void wpe_load_chunk(const float2* __restrict__ Y, const WpeDims& g, int f, int t0, int W, int hist,
                                               float2* ysm) {
  for (int i = threadIdx.x; i < g.D * W; i += blockDim.x) {
    const int d = i / W, tt = t0 - hist + (i - d * W);
    ysm[i] = (tt >= 0 && tt < g.T) ? Y[(static_cast<int64_t>(f) * g.D + d) * g.T + tt] : make_float2(0.f, 0.f);
  }
}

// acc layout per frequency (doubles, interleaved re/im): R as DK x DK (upper 4x4 tiles written), then P as DK x D
__global__ void __launch_bounds__(256)
wpe_stats_kernel(const float2* __restrict__ Y, const float* __restrict__ power, const unsigned int* __restrict__ pmax,
                 const WpeDims g, double* __restrict__ acc) {
  extern __shared__ __align__(16) uint8_t wpe_smem[];
  const int f = blockIdx.y, t0 = blockIdx.x * kWpeChunk;
  const int hist = g.delay + g.taps - 1, W = kWpeChunk + hist;
  float2* ysm = reinterpret_cast<float2*>(wpe_smem);
  float* inv = reinterpret_cast<float*>(ysm + g.D * W);
  wpe_load_chunk(Y, g, f, t0, W, hist, ysm);
  const float eps = 1e-10f * __uint_as_float(pmax[f]);
  for (int i = threadIdx.x; i < kWpeChunk; i += blockDim.x) {
    const int t = t0 + i;
    inv[i] = (t < g.T && t >= g.t_first) ? 1.0f / fmaxf(power[static_cast<int64_t>(f) * g.T + t], eps) : 0.f;
  }
  __syncthreads();
  const int DK = g.DK, D = g.D;
  const int nt = (DK + 3) / 4;                   // 4-row tiles of R
  const int n_r = nt * (nt + 1) / 2;             // upper-triangle tiles
  const int n_p = nt * ((D + 3) / 4);
  double* accf = acc + static_cast<int64_t>(f) * 2 * (static_cast<int64_t>(DK) * DK + static_cast<int64_t>(DK) * D);
  for (int task = threadIdx.x; task < n_r + n_p; task += blockDim.x) {
    int ti, tj;
    const bool is_p = task >= n_r;
    if (!is_p) {
      // row-major enumeration of the upper triangle: tile row ti holds nt - ti tiles
      int rem = task;
      ti = 0;
      while (rem >= nt - ti) {
        rem -= nt - ti;
        ++ti;
      }
      tj = ti + rem;
    } else {
      ti = (task - n_r) / ((D + 3) / 4);
      tj = (task - n_r) % ((D + 3) / 4);
    }
    // offsets of the 4 rows / 4 columns inside the chunk's frame window: a[(tau, d), t] = ysm[d][hist + i - delay - tau]
    int ro[4], co[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ti * 4 + k;
      ro[k] = r < DK ? (r % D) * W + hist - g.delay - r / D : -1;
      const int c = tj * 4 + k;
      if (!is_p) co[k] = c < DK ? (c % D) * W + hist - g.delay - c / D : -1;
      else co[k] = c < D ? c * W + hist : -1;
    }
    float sr[4][4], si[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) sr[a][b] = si[a][b] = 0.f;
    for (int i = 0; i < kWpeChunk; ++i) {
      const float w = inv[i];
      float2 av[4], bv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        av[k] = ro[k] >= 0 ? ysm[ro[k] + i] : make_float2(0.f, 0.f);
        bv[k] = co[k] >= 0 ? ysm[co[k] + i] : make_float2(0.f, 0.f);
        av[k].x *= w;
        av[k].y *= w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {  // a * conj(b)
          sr[a][b] = fmaf(av[a].x, bv[b].x, fmaf(av[a].y, bv[b].y, sr[a][b]));
          si[a][b] = fmaf(av[a].y, bv[b].x, fmaf(-av[a].x, bv[b].y, si[a][b]));
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = ti * 4 + a, c = tj * 4 + b;
        if (r >= DK) continue;
        if (!is_p) {
          if (c >= DK || c < r) continue;  // the strict lower triangle is rebuilt from the upper one by the solver
          double* dst = accf + 2 * (static_cast<int64_t>(r) * DK + c);
          atomicAdd(dst, static_cast<double>(sr[a][b]));
          atomicAdd(dst + 1, static_cast<double>(si[a][b]));
        } else {
          if (c >= D) continue;
          double* dst = accf + 2 * (static_cast<int64_t>(DK) * DK + static_cast<int64_t>(r) * D + c);
          atomicAdd(dst, static_cast<double>(sr[a][b]));
          atomicAdd(dst + 1, static_cast<double>(si[a][b]));
        }
      }
  }
}

// G = R^-1 P per frequency: Gauss-Jordan with partial pivoting on the augmented [R | P] (DK x (DK + D)) in f64.
// A frequency whose R is singular (an all-zero bin) gets G = 0, i.e. X = Y there.
__global__ void __launch_bounds__(256) wpe_solve_kernel(const double* __restrict__ acc, int DK, int D, float2* __restrict__ G) {
  extern __shared__ __align__(16) uint8_t wpe_smem[];
  double2* M = reinterpret_cast<double2*>(wpe_smem);  // [DK][NCOL]
  __shared__ int piv_row;
  __shared__ int singular;
  const int f = blockIdx.x, NCOL = DK + D;
  const double* accf = acc + static_cast<int64_t>(f) * 2 * (static_cast<int64_t>(DK) * DK + static_cast<int64_t>(DK) * D);
  for (int i = threadIdx.x; i < DK * NCOL; i += blockDim.x) {
    const int r = i / NCOL, c = i - r * NCOL;
    double2 v;
    if (c < DK) {
      if (c >= r) {
        v.x = accf[2 * (static_cast<int64_t>(r) * DK + c)];
        v.y = c == r ? 0.0 : accf[2 * (static_cast<int64_t>(r) * DK + c) + 1];
      } else {  // hermitian
        v.x = accf[2 * (static_cast<int64_t>(c) * DK + r)];
        v.y = -accf[2 * (static_cast<int64_t>(c) * DK + r) + 1];
      }
    } else {
      const double* p = accf + 2 * (static_cast<int64_t>(DK) * DK + static_cast<int64_t>(r) * D + (c - DK));
      v.x = p[0];
      v.y = p[1];
    }
    M[i] = v;
  }
  if (threadIdx.x == 0) singular = 0;
  __syncthreads();
  for (int k = 0; k < DK; ++k) {
    if (threadIdx.x < 32) {
      double best = -1.0;
      int bi = k;
      for (int r = k + threadIdx.x; r < DK; r += 32) {
        const double2 v = M[r * NCOL + k];
        const double m = v.x * v.x + v.y * v.y;
        if (m > best) {
          best = m;
          bi = r;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (threadIdx.x == 0) {
        piv_row = bi;
        if (!(best > 0.0) || !isfinite(best)) singular = 1;
      }
    }
    __syncthreads();
    if (singular) break;
    const int pr = piv_row;
    if (pr != k) {
      for (int c = threadIdx.x; c < NCOL; c += blockDim.x) {
        const double2 t = M[k * NCOL + c];
        M[k * NCOL + c] = M[pr * NCOL + c];
        M[pr * NCOL + c] = t;
      }
    }
    __syncthreads();
    const double2 p = M[k * NCOL + k];
    const double pn = 1.0 / (p.x * p.x + p.y * p.y);
    const double2 pinv = make_double2(p.x * pn, -p.y * pn);
    __syncthreads();
    for (int c = threadIdx.x; c < NCOL; c += blockDim.x) {  // normalise the pivot row
      const double2 v = M[k * NCOL + c];
      M[k * NCOL + c] = make_double2(v.x * pinv.x - v.y * pinv.y, v.x * pinv.y + v.y * pinv.x);
    }
    __syncthreads();
    // eliminate column k from every other row; columns <= k of those rows are never read again
    const int ncol_live = NCOL - (k + 1);
    for (int i = threadIdx.x; i < (DK - 1) * ncol_live; i += blockDim.x) {
      int r = i / ncol_live;
      const int c = k + 1 + (i - r * ncol_live);
      if (r >= k) ++r;
      const double2 m = M[r * NCOL + k], v = M[k * NCOL + c];
      double2 x = M[r * NCOL + c];
      x.x -= m.x * v.x - m.y * v.y;
      x.y -= m.x * v.y + m.y * v.x;
      M[r * NCOL + c] = x;
    }
    __syncthreads();
  }
  const bool bad = singular != 0;
  for (int i = threadIdx.x; i < DK * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D;
    const double2 v = M[r * NCOL + DK + d];
    G[static_cast<int64_t>(f) * DK * D + i] = bad ? make_float2(0.f, 0.f) : make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
  }
}

// X[f, d, t] = Y[f, d, t] - sum_r conj(G[r, d]) Ytilde[r, t]; optionally lambda of X for the next iteration
template <int D>
__global__ void __launch_bounds__(kWpeChunk)
wpe_apply_kernel(const float2* __restrict__ Y, const float2* __restrict__ G, const WpeDims g, float2* __restrict__ X,
                 float* __restrict__ power, unsigned int* __restrict__ pmax) {
  extern __shared__ __align__(16) uint8_t wpe_smem[];
  const int f = blockIdx.y, t0 = blockIdx.x * kWpeChunk;
  const int hist = g.delay + g.taps - 1, W = kWpeChunk + hist;
  float2* ysm = reinterpret_cast<float2*>(wpe_smem);
  float2* gsm = ysm + D * W;
  wpe_load_chunk(Y, g, f, t0, W, hist, ysm);
  for (int i = threadIdx.x; i < g.DK * D; i += blockDim.x) gsm[i] = G[static_cast<int64_t>(f) * g.DK * D + i];
  __syncthreads();
  const int i = threadIdx.x, t = t0 + i;
  float2 acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = ysm[d * W + hist + i];
  for (int tau = 0; tau < g.taps; ++tau) {
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
      const float2 a = ysm[dd * W + hist + i - g.delay - tau];
      const float2* gr = gsm + (tau * D + dd) * D;
#pragma unroll
      for (int d = 0; d < D; ++d) {  // acc -= conj(G) * a
        const float2 q = gr[d];
        acc[d].x -= q.x * a.x + q.y * a.y;
        acc[d].y -= q.x * a.y - q.y * a.x;
      }
    }
  }
  float p = 0.f;
  if (t < g.T) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      X[(static_cast<int64_t>(f) * D + d) * g.T + t] = acc[d];
      p += acc[d].x * acc[d].x + acc[d].y * acc[d].y;
    }
    p /= static_cast<float>(D);
    if (power) power[static_cast<int64_t>(f) * g.T + t] = p;
  }
  if (pmax) block_max_to(p, pmax + f);
}

static size_t wpe_align(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct WpeWorkspace {
  size_t yt, xt, power, power2, pmax, acc, g, total;
};
static WpeWorkspace wpe_layout(int D, int64_t T, int F, int taps) {
  const size_t DK = static_cast<size_t>(D) * taps;
  WpeWorkspace w;
  size_t o = 0;
  w.yt = o;      o += wpe_align(sizeof(float2) * F * D * T);
  w.xt = o;      o += wpe_align(sizeof(float2) * F * D * T);
  w.power = o;   o += wpe_align(sizeof(float) * F * T);
  w.power2 = o;  o += wpe_align(sizeof(float) * F * T);
  w.pmax = o;    o += wpe_align(sizeof(unsigned int) * F);
  w.acc = o;     o += wpe_align(sizeof(double) * 2 * F * (DK * DK + DK * D));
  w.g = o;       o += wpe_align(sizeof(float2) * F * DK * D);
  w.total = o;
  return w;
}

}  // namespace tssep

using namespace tssep;

extern "C" {

int64_t tssep_wpe_workspace_bytes(int D, int64_t T, int F, int taps) {
  if (D < 1 || D > kWpeMaxD || T < 0 || F < 1 || taps < 1) {
    set_error("tssep_wpe_workspace_bytes: bad extent");
    return -1;
  }
  return static_cast<int64_t>(wpe_layout(D, T, F, taps).total);
}

int tssep_wpe(const float* Y, int D, int64_t T, int F, int taps, int delay, int iterations, int psd_context,
              int statistics_mode, float* X, void* workspace, int64_t workspace_bytes, tssep_stream_t stream) {
  TSSEP_REQUIRE(Y && X && workspace, "tssep_wpe: null pointer");
  TSSEP_REQUIRE(D >= 1 && D <= kWpeMaxD && F >= 1 && F <= 65535 && T >= 0 && T < (1ll << 31) - 4096,
                "tssep_wpe: need 1 <= D <= %d channels, 1 <= F <= 65535 (got D=%d F=%d)", kWpeMaxD, D, F);
  TSSEP_REQUIRE(taps >= 1 && delay >= 0 && iterations >= 0 && psd_context >= 0 && (statistics_mode == 0 || statistics_mode == 1),
                "tssep_wpe: bad taps / delay / iterations / psd_context / statistics_mode");
  const int DK = D * taps, NCOL = DK + D;
  const size_t solve_smem = sizeof(double2) * static_cast<size_t>(DK) * NCOL;
  TSSEP_REQUIRE(solve_smem <= 200 * 1024, "tssep_wpe: D * taps = %d is too large for the in-shared-memory solve", DK);
  const int hist = delay + taps - 1, W = kWpeChunk + hist;
  const size_t stats_smem = sizeof(float2) * static_cast<size_t>(D) * W + sizeof(float) * kWpeChunk;
  const size_t apply_smem = sizeof(float2) * (static_cast<size_t>(D) * W + static_cast<size_t>(DK) * D);
  TSSEP_REQUIRE(stats_smem <= 200 * 1024 && apply_smem <= 200 * 1024, "tssep_wpe: delay + taps too large for shared memory");
  const WpeWorkspace w = wpe_layout(D, T, F, taps);
  TSSEP_REQUIRE(workspace_bytes >= static_cast<int64_t>(w.total), "tssep_wpe: workspace of %lld bytes needed, got %lld",
                static_cast<long long>(w.total), static_cast<long long>(workspace_bytes));
  TSSEP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tssep_wpe: workspace must be 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (T == 0) return 0;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float2* Yt = reinterpret_cast<float2*>(ws + w.yt);
  float2* Xt = reinterpret_cast<float2*>(ws + w.xt);
  float* power = reinterpret_cast<float*>(ws + w.power);
  float* power2 = reinterpret_cast<float*>(ws + w.power2);
  unsigned int* pmax = reinterpret_cast<unsigned int*>(ws + w.pmax);
  double* acc = reinterpret_cast<double*>(ws + w.acc);
  float2* G = reinterpret_cast<float2*>(ws + w.g);
  const int Ti = static_cast<int>(T);
  const dim3 tgrid((F + 31) / 32, (Ti + 31) / 32, D), tblock(32, 8);
  if (iterations == 0) {
    TSSEP_CUDA(cudaMemcpyAsync(X, Y, sizeof(float2) * D * T * F, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  wpe_to_fdt_kernel<<<tgrid, tblock, 0, s>>>(reinterpret_cast<const float2*>(Y), D, Ti, F, Yt);
  WpeDims g{D, Ti, F, taps, delay, DK, statistics_mode == 1 ? delay + taps - 1 : 0};
  const dim3 cgrid((Ti + kWpeChunk - 1) / kWpeChunk, F);
  TSSEP_CUDA(cudaFuncSetAttribute(wpe_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stats_smem)));
  TSSEP_CUDA(cudaFuncSetAttribute(wpe_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(solve_smem)));
  // one thread per 4x4 tile of the statistics / per element of a solver row: small problems (ChannelWiseWPE: D = 1,
  // 10 x 10 matrices, thousands of frequencies) get small CTAs, so that many of them share an SM
  const int nt4 = (DK + 3) / 4;
  const int stats_threads = static_cast<int>(imin64(256, ((nt4 * (nt4 + 1) / 2 + nt4 * ((D + 3) / 4) + 31) / 32) * 32));
  const int solve_threads = static_cast<int>(imin64(256, ((static_cast<int64_t>(DK) * NCOL / 4 + 31) / 32) * 32));
  for (int it = 0; it < iterations; ++it) {
    const bool last = it + 1 == iterations;
    if (it == 0) {
      TSSEP_CUDA(cudaMemsetAsync(pmax, 0, sizeof(unsigned int) * F, s));
      wpe_power_kernel<<<cgrid, kWpeChunk, 0, s>>>(Yt, D, Ti, power, psd_context > 0 ? nullptr : pmax);
    }
    const float* lam = power;
    if (psd_context > 0) {
      wpe_smooth_kernel<<<cgrid, kWpeChunk, 0, s>>>(power, Ti, psd_context, power2, pmax);
      lam = power2;
    }
    TSSEP_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * F * (static_cast<size_t>(DK) * DK + static_cast<size_t>(DK) * D), s));
    wpe_stats_kernel<<<cgrid, stats_threads, stats_smem, s>>>(Yt, lam, pmax, g, acc);
    wpe_solve_kernel<<<F, solve_threads, solve_smem, s>>>(acc, DK, D, G);
    if (!last) TSSEP_CUDA(cudaMemsetAsync(pmax, 0, sizeof(unsigned int) * F, s));  // read by the kernels above, rewritten below
    float* pw = last ? nullptr : power;
    unsigned int* pm = (last || psd_context > 0) ? nullptr : pmax;
#define TSSEP_WPE_APPLY(DD)                                                                                               \
  case DD:                                                                                                                \
    TSSEP_CUDA(cudaFuncSetAttribute(wpe_apply_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,                    \
                                    static_cast<int>(apply_smem)));                                                       \
    wpe_apply_kernel<DD><<<cgrid, kWpeChunk, apply_smem, s>>>(Yt, G, g, Xt, pw, pm);                                      \
    break;
    switch (D) {
      TSSEP_WPE_APPLY(1)
      TSSEP_WPE_APPLY(2)
      TSSEP_WPE_APPLY(3)
      TSSEP_WPE_APPLY(4)
      TSSEP_WPE_APPLY(5)
      TSSEP_WPE_APPLY(6)
      TSSEP_WPE_APPLY(7)
      TSSEP_WPE_APPLY(8)
    }
#undef TSSEP_WPE_APPLY
  }
  wpe_from_fdt_kernel<<<tgrid, tblock, 0, s>>>(Xt, D, Ti, F, reinterpret_cast<float2*>(X));
  return check_launch("tssep_wpe");
}

}  // extern "C"
