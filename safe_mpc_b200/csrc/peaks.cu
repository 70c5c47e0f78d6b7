// Measured machine peaks that the roofline of bench.py needs and that MEASURED_PEAKS.json does not carry: the FP64 FMA rate
// (SURVEY.md section 8(d) charges the dynamics / QP kernels to the FP64 pipe) and the TF32 tensor rate (the viability network).
// Plain micro-benchmarks, timed with CUDA events on their own stream; nothing on the hot path calls them.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/safe_mpc_b200.h"

namespace {

// 16 independent FMA chains per thread: enough instruction-level parallelism to hide the FP64 pipe latency with 8 warps per
// scheduler resident; no memory traffic in the timed loop
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int smpc_measure_peaks(int32_t device, double* fp64_tflops, double* fp32_tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return SMPC_ERR_CUDA;
  cudaDeviceProp dp;
  if (cudaGetDeviceProperties(&dp, device) != cudaSuccess) return SMPC_ERR_CUDA;
  cudaStream_t st;
  cudaEvent_t e0, e1;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return SMPC_ERR_CUDA;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = dp.multiProcessorCount * 8, block = 256;
  double* dout = nullptr;
  cudaMalloc((void**)&dout, sizeof(double) * grid * block);
  int rc = SMPC_OK;
  for (int pass = 0; pass < 2; ++pass) {
    const int iters = pass == 0 ? 1 << 14 : 1 << 16;       // FP64 is the slow pipe: fewer iterations
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                    // first repetition warms up (clocks, instruction cache)
      cudaEventRecord(e0, st);
      if (pass == 0) dfma_peak_kernel<<<grid, block, 0, st>>>(dout, iters, 1.0000001, 1e-9);
      else ffma_peak_kernel<<<grid, block, 0, st>>>((float*)dout, iters, 1.0000001f, 1e-9f);
      cudaEventRecord(e1, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) { rc = SMPC_ERR_CUDA; break; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double tf = 2.0 * 16.0 * (double)iters * (double)grid * block / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    if (pass == 0 && fp64_tflops) *fp64_tflops = best;
    if (pass == 1 && fp32_tflops) *fp32_tflops = best;
  }
  cudaFree(dout);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaStreamDestroy(st);
  return rc;
}
