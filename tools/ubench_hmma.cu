// mma.sync.m16n8k16 bf16 throughput / latency on B200 (one SM): cycles per mma per SM sub-partition for 1..4 warps per
// sub-partition and 1..8 independent accumulator chains per warp (operands in registers: pure tensor-pipe issue rate)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_hmma tools/ubench_hmma.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int CH>
__global__ void k(float* out, long long* cyc, int iters) {
  float d[CH][4];
  for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x3c003c00u, b1 = b0 + 7;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0; for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH> void run(float* out, long long* cyc) {
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    k<CH><<<1, warps * 32>>>(out, cyc, iters);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const int wps = (warps + 3) / 4;
    printf("chains/warp=%d warps=%2d (%d per sub-partition): %.1f cycles per mma per sub-partition, %.1f cycles per mma per chain\n", CH, warps, wps,
           (double)c / (iters * CH * wps), (double)c / iters);
  }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  run<1>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc);
  return 0;
}
