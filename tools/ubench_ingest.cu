// How many bytes per clock can one SM pull from L2, and through which path? (diagnostics, not product code)
//   warps 0..NT-1 : bulk-copy (TMA) streams, each its own 4-stage ring of 16 KB copies (released at once)
//   warps 4..4+NL-1 : LDG.128 streams (8 loads in flight per lane), or LDGSTS (cp.async 16 B) streams
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../gst_tacotron_b200/csrc/umma.cuh"
using namespace gstk;
__device__ __forceinline__ long long clk() { long long v; asm volatile("mov.u64 %0, %%clock64;" : "=l"(v) :: "memory"); return v; }

__global__ void __launch_bounds__(512) k(const uint8_t* src, long long* out, int units, int NT, int NL, int ldgsts, int cbytes) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[4][4];
  __shared__ long long tt[16];
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) { for (int i = 0; i < 16; ++i) mbar_init(&full[i >> 2][i & 3], 1); mbar_fence_init(); }
  if (tid < 16) tt[tid] = 0;
  __syncthreads();
  const long long t0 = clk();
  const size_t region = 4u << 20;  // 4 MB source window (L2 resident)
  if (wid < NT) {
    uint8_t* buf = sm + wid * 4 * cbytes;  // NT * 4 * cbytes <= 128 KB
    // keep 4 copies in flight: issue u+4 when u has landed
    for (int u = 0; u < units + 4; ++u) {
      const int st = u & 3;
      if (u >= 4) mbar_wait(&full[wid][st], ((u - 4) >> 2) & 1);
      if (u < units && elect_one()) {
        mbar_arrive_expect_tx(&full[wid][st], cbytes);
        bulk_g2s(buf + st * cbytes, src + ((size_t)(u * 29 + wid * 7 + blockIdx.x * 3) * 16384) % region, cbytes, &full[wid][st]);
      }
      __syncwarp();
    }
    if (lane == 0) tt[wid] = clk() - t0;
  } else if (wid >= 4 && wid < 4 + NL) {
    const int w = wid - 4;
    if (!ldgsts) {
      uint4 acc = make_uint4(0, 0, 0, 0);
      // each iteration: the warp reads 8 x 512 B = 4 KB
      for (int u = 0; u < units * 4; ++u) {
        const uint8_t* p = src + ((size_t)(u * 13 + w * 5 + blockIdx.x) * 4096) % region + lane * 16;
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcg(reinterpret_cast<const uint4*>(p + j * 512));
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc.x ^= v[j].x; acc.y ^= v[j].y; acc.z ^= v[j].z; acc.w ^= v[j].w; }
      }
      if (acc.x == 0x12345) out[1000] = acc.y;
    } else {
      uint8_t* buf = sm + 131072 + w * 8192;
      for (int u = 0; u < units * 4; ++u) {
        const uint8_t* p = src + ((size_t)(u * 13 + w * 5 + blockIdx.x) * 4096) % region + lane * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + (u & 1) * 4096 + j * 512 + lane * 16)), "l"(p + j * 512) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (lane == 0) tt[4 + w] = clk() - t0;
  }
  __syncthreads();
  if (tid < 16) out[blockIdx.x * 16 + tid] = tt[tid];
}

int main(int argc, char** argv) {
  uint8_t* src; cudaMalloc(&src, 8u << 20); cudaMemset(src, 0, 8u << 20);
  long long* out; cudaMalloc(&out, 148 * 16 * 8 + 16384);
  const int smem = 131072 + 12 * 8192 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int units = 256;
  struct Cfg { int NT, NL, ldgsts, cbytes; };
  const Cfg cfgs[] = {{1, 0, 0, 16384}, {2, 0, 0, 16384}, {4, 0, 0, 8192}, {1, 0, 0, 8192}, {2, 0, 0, 8192}, {4, 0, 0, 4096},
                      {0, 1, 0, 0}, {0, 4, 0, 0}, {0, 8, 0, 0}, {0, 12, 0, 0}, {0, 4, 1, 0}, {0, 8, 1, 0}, {0, 12, 1, 0},
                      {1, 4, 0, 16384}, {1, 8, 0, 16384}, {2, 8, 0, 16384}, {1, 8, 1, 16384}, {2, 8, 1, 16384}};
  std::vector<int> grids = {1, 128, 148};
  if (argc > 1) { grids.clear(); for (int i = 1; i < argc; ++i) grids.push_back(atoi(argv[i])); }   // e.g. 16 32 64 74 96 128 148: per-SM rate vs streaming SMs
  for (int grid : grids)
    for (const Cfg& c : cfgs) {
      for (int rep = 0; rep < 2; ++rep) k<<<grid, 512, smem>>>(src, out, units, c.NT, c.NL, c.ldgsts, c.cbytes);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid * 16); cudaMemcpy(h.data(), out, grid * 128, cudaMemcpyDeviceToHost);
      double tma = 0, ldg = 0, tmax = 0;
      for (int b = 0; b < grid; ++b) {
        for (int w = 0; w < c.NT; ++w) { tma += (double)units * c.cbytes / h[b * 16 + w]; tmax = std::max(tmax, (double)h[b * 16 + w]); }
        for (int w = 0; w < c.NL; ++w) { ldg += (double)units * 4 * 4096 / h[b * 16 + 4 + w]; tmax = std::max(tmax, (double)h[b * 16 + 4 + w]); }
      }
      printf("grid=%3d tma_streams=%d (%5d B) %s_warps=%2d : TMA %6.1f B/clk/SM   LSU %6.1f B/clk/SM   total %6.1f   chip %7.0f B/clk\n", grid, c.NT, c.cbytes,
             c.ldgsts ? "ldgsts" : "ldg   ", c.NL, tma / grid, ldg / grid, (tma + ldg) / grid, tma + ldg);
    }
  return 0;
}
