// MUFU throughput on B200: tanh.approx vs ex2 + rcp formulations (one SM, 1..16 warps), cycles per warp-instruction
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mufu tools/ubench_mufu.cu
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float tanh_a(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_a(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = tanh_a(a[i]);
      else if (MODE == 1) a[i] = ex2_a(a[i]);
      else if (MODE == 2) a[i] = rcp_a(a[i]);
      else a[i] = 1.0f - 2.0f * rcp_a(ex2_a(a[i] * 2.885390f) + 1.0f);   // tanh via ex2 + rcp
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  const char* names[4] = {"tanh.approx", "ex2.approx", "rcp.approx", "tanh via ex2+rcp"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 3) k<3><<<1, warps * 32>>>(out, cyc, iters);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-18s warps=%2d: %.2f cycles per warp-level op per SM sub-partition (%.1f results/clk/SM)\n", names[mode], warps,
             (double)c / (iters * 8.0) / ((warps + 3) / 4), warps * 32.0 * iters * 8 / c);
    }
  return 0;
}
