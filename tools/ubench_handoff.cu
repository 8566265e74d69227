// Bisecting the producer/consumer hand-off cost of a bulk-copy ring (diagnostics, not product code).
//   relay 0: one warp issues and waits on full itself (4 copies in flight)
//   relay 1: producer warp waits on empty[], consumer warp waits on full[] and arrives on empty[]
//   relay 2: as 1, consumer additionally runs tcgen05.fence::after_thread_sync per unit
//   relay 3: as 1, but producer and consumer timing instrumentation off (no clk in loops) - always off here
#include <cstdio>
#include <vector>
#include "../gst_tacotron_b200/csrc/umma.cuh"
using namespace gstk;
__device__ __forceinline__ long long clk() { long long v; asm volatile("mov.u64 %0, %%clock64;" : "=l"(v) :: "memory"); return v; }
__device__ __forceinline__ void mbar_arrive_l(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

__global__ void __launch_bounds__(128) k(const uint8_t* src, long long* out, int units, int relay, int S, int cwarp, int feat) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[8], empty[8];
  __shared__ long long tt[2];
  __shared__ uint32_t tmem_s;
  const int SB = (feat & 1) ? 24576 : 16384;
  long long acc = 0;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_fence_init(); }
  if ((feat & 4) && tid < 32) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long t0 = clk();
  if (wid == 0) {
    if (relay == 0) {
      int st = 0; uint32_t ph = 0;
      for (int u = 0; u < units + S; ++u) {
        if (u >= S) mbar_wait(&full[st], ph ^ 1u);
        if (u < units && elect_one()) {
          mbar_arrive_expect_tx(&full[st], 16384u);
          bulk_g2s(sm + st * 16384, src + (size_t)((u * 7 + blockIdx.x) & 63) * 16384, 16384u, &full[st]);
        }
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1u; }
      }
    } else {
      int st = 0; uint32_t ph = 0;
      for (int u = 0; u < units; ++u) {
        if (feat & 2) { const long long w0 = clk(); mbar_wait(&empty[st], ph ^ 1u); acc += clk() - w0; } else mbar_wait(&empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[st], 16384u);
          bulk_g2s(sm + st * SB, src + (size_t)((u * 7 + blockIdx.x) & 63) * 16384, 16384u, &full[st]);
        }
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1u; }
      }
    }
    if (lane == 0) tt[0] = clk() - t0 + (acc == 12345);
  } else if (wid == cwarp && relay != 0) {
    int st = 0; uint32_t ph = 0;
    for (int u = 0; u < units; ++u) {
      if (feat & 2) { const long long w0 = clk(); mbar_wait(&full[st], ph); acc += clk() - w0; } else mbar_wait(&full[st], ph);
      if (relay == 2) tc_fence_after();
      if (elect_one()) mbar_arrive_l(&empty[st]);
      __syncwarp();
      if (++st == S) { st = 0; ph ^= 1u; }
    }
    if (lane == 0) tt[1] = clk() - t0 + (acc == 12345);
  }
  __syncthreads();
  if (tid < 2) out[blockIdx.x * 2 + tid] = tt[tid];
  if ((feat & 4) && tid < 32) tmem_dealloc(tmem_s, 512);
}

int main() {
  uint8_t* src; cudaMalloc(&src, 8u << 20); cudaMemset(src, 0, 8u << 20);
  long long* out; cudaMalloc(&out, 148 * 16);
  const int smem = 8 * 24576 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int units = 512;
  for (int grid : {1, 128})
    for (int relay : {1})
      for (int S : {4})
        for (int feat : {0, 1, 2, 4, 8, 9, 7})
        for (int cwarp : {2}) {
          const int smem_l = (feat & 9) ? smem : 8 * 16384 + 1024;
          for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, smem_l>>>(src, out, units, relay, S, cwarp, feat);
          cudaError_t e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> h(grid * 2); cudaMemcpy(h.data(), out, grid * 16, cudaMemcpyDeviceToHost);
          double a = 0, b = 0; for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
          printf("grid=%3d relay=%d stages=%d feat=%d: producer %6.0f ticks/unit (%.1f B/clk), consumer %6.0f ticks/unit\n", grid, relay, S, feat,
                 a / grid / units, 16384.0 * units * grid / a, b / grid / units);
        }
  return 0;
}
