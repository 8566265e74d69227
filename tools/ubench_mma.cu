// mma.sync m16n8k16 bf16 latency / throughput on B200 (diagnostics).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mma(float (&d)[4], uint4 a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
template <int CH>
__global__ void k(long long* out, float* sink, int iters) {
  float d[CH][4];
  for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  uint4 a = make_uint4(threadIdx.x, 2, 3, 4);
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) mma(d[c], a, 0x3f803f80u, 0x3f803f80u);
  }
  long long t1 = clock64();
  float s = 0; for (int c = 0; c < CH; ++c) s += d[c][0] + d[c][3];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  long long* out; float* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 4096 * 4);
  for (int warps : {1, 4, 10, 16}) {
    long long h;
#define RUN(CH) k<CH><<<1, warps * 32>>>(out, sink, 256); cudaDeviceSynchronize(); k<CH><<<1, warps * 32>>>(out, sink, 256); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost); printf("warps=%2d chains=%d: %6.1f cycles per mma per warp (%.1f per SM-wide mma)\n", warps, CH, (double)h / (256.0 * CH), (double)h / (256.0 * CH * warps));
    RUN(1) RUN(2) RUN(4) RUN(8)
  }
  return 0;
}
