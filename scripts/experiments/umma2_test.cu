// Experiment (not part of the product): one tcgen05.mma.cta_group::2 (CTA pair, M = 256, N = 128, K = 8, tf32) with both operands in the
// K-major no-swizzle layout of csrc/mlp_tc.cu -- checks operand split, TMEM accumulator layout, multicast commit.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma2_test umma2_test.cu && ./umma2_test
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 256, N = 128, K = 8;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) umma2_kernel(const float* __restrict__ A, const float* __restrict__ B, float* out) {
  __shared__ __align__(128) float sA[2 * 128 * 4];   // [chunk 2][row 128][4]
  __shared__ __align__(128) float sB[2 * 64 * 4];    // [chunk 2][row 64][4]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, warp = tid >> 5;
  // this CTA's halves: A rows [128 rank, +128), B rows [64 rank, +64)
  for (int i = tid; i < 128 * K; i += 128) { const int r = i / K, k = i % K; sA[(k / 4) * 128 * 4 + r * 4 + (k % 4)] = A[(size_t)(128 * rank + r) * K + k]; }
  for (int i = tid; i < 64 * K; i += 128) { const int r = i / K, k = i % K; sB[(k / 4) * 64 * 4 + r * 4 + (k % 4)] = B[(size_t)(64 * rank + r) * K + k]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (rank == 0 && tid == 0) {
    const uint64_t ad = smem_desc(s32(sA), 128 * 16, 128), bd = smem_desc(s32(sB), 64 * 16, 128);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad),
                 "l"(bd), "r"(IDESC), "r"(0)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(&bar)), "h"((uint16_t)3)
                 : "memory");
  }
  {
    uint32_t ok;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
    } while (!ok);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tl + c)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[((size_t)rank * 128 + tid) * N + c + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
  std::vector<float> A(M * K), B(N * K), out(M * N, -1.f);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = (float)((m * 3 + k * 5) % 7 - 3);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = (float)((n * 2 + k * 3) % 5 - 2);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, out.size() * 4);
  umma2_kernel<<<2, 128>>>(dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float ref = 0; for (int k = 0; k < K; ++k) ref += A[m * K + k] * B[n * K + k];
    if (out[m * N + n] != ref) { if (bad < 8) printf("mismatch D[%d][%d] = %g, expected %g\n", m, n, out[m * N + n], ref); ++bad; }
  }
  printf("%d mismatches of %d\n", bad, M * N);
  return bad != 0;
}
