// Microbenchmark (not part of the library): cycles per tcgen05.mma M=128, K=16 at small N, with the A operand in
// tensor memory, in shared memory, and alternating between the two.  Operands are whatever the memories hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I tssep_b200/csrc -o /tmp/mma_probe scripts/probes/mma_probe.cu
#include <cstdio>
#include "../../tssep_b200/csrc/common.cuh"
using namespace tssep;

namespace tssep {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace tssep

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

// MODE 0: A from TMEM; 1: A from smem; 2: alternating; 3: alternating, two accumulators; 4: A from TMEM, two accumulators
template <int N, int MODE, int NM>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;            // 4 k-atoms of A (64 KB), reused
  const uint32_t sB = base + 65536;    // 5 atoms x N x 128 B
  const uint32_t bar = sB + 5 * N * 128, tptr = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (bar - base) / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + 16 * i), "r"(0x3c003c00u) : "memory");
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tc_alloc(tptr, 512);
    tc_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t ad0 = desc_sw128(sA), bd0 = desc_sw128(sB);
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    for (int r = 0; r < reps + 1; ++r) {
      if (r == 1) t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < NM; ++k) {
          const uint32_t a_t = tmem + 8 * (k % 19);
          const uint64_t a_s = ad0 + ((k % 4) * 1024) + 2 * ((k >> 2) & 3);
          const uint64_t b = bd0 + ((k % 5) * (N * 8)) + 2 * ((k / 5) & 3);
          const bool ss = MODE == 1 || ((MODE == 2 || MODE == 3) && (k & 1));
          const uint32_t d = tmem + 256 + (((MODE == 3 || MODE == 4) && (k & 1)) ? N : 0);
          if (ss) tc_mma_bf16(d, a_s, b, idesc, k > 1 ? 1u : 0u);
          else mma_ts(d, a_t, b, idesc, k > 1 ? 1u : 0u);
        }
        tc_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, ph);
      ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tc_dealloc(tmem, 512);
  }
}

template <int N, int MODE, int NM>
void run1(long long* d_out) {
  const size_t smem = 1024 + 65536 + 5 * N * 128 + 64;
  cudaFuncSetAttribute(probe<N, MODE, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int reps = 2000;
  probe<N, MODE, NM><<<1, 128, smem>>>(reps, d_out);
  long long c = 0;
  cudaError_t e = cudaMemcpy(&c, d_out, sizeof(c), cudaMemcpyDeviceToHost);
  const char* names[5] = {"A in TMEM", "A in smem", "alternating TMEM/smem", "alternating, 2 accumulators", "A in TMEM, 2 accumulators"};
  printf("N=%3d %-28s mmas/batch=%2d: %7.1f cycles per batch  %s\n", N, names[MODE], NM, double(c) / reps,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N>
void run(long long* d_out) {
  run1<N, 0, 19>(d_out); run1<N, 0, 38>(d_out); run1<N, 0, 57>(d_out);
  run1<N, 1, 19>(d_out); run1<N, 1, 38>(d_out); run1<N, 1, 57>(d_out);
  run1<N, 2, 19>(d_out); run1<N, 2, 38>(d_out); run1<N, 2, 57>(d_out);
  run1<N, 3, 38>(d_out); run1<N, 3, 57>(d_out);
  run1<N, 4, 38>(d_out); run1<N, 4, 57>(d_out);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<16>(d_out);
  run<32>(d_out);
  run<64>(d_out);
  run<128>(d_out);
  return 0;
}
