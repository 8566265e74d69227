// Times fa_dense_mma (phase A dense layers) in isolation (diagnostics).
#include <cstdio>
#include <vector>
#include "../gst_tacotron_b200/csrc/decoder_bf16.cuh"
using namespace gstk;

template <int NU>
__global__ void __launch_bounds__(TC_THREADS, 1) k(const uint4* w, long long* out, int NF, int KT, int reps, int smem_pad) {
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(sm);
  float* part = reinterpret_cast<float*>(sm + 2 * FA_HC * 2);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 2 * FA_HC; i += blockDim.x) act[i] = __float2bfloat16(0.01f * (i % 37));
  __syncthreads();
  if (wid < TC_PA_WARPS) {
    pa_sync<TC_PA_THREADS>();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      fa_dense_mma<NU>(w, NF, KT, act, FA_HC, part, wid, lane);
      pa_sync<TC_PA_THREADS>();
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = (t1 - t0) / reps;
  }
}

int main() {
  const size_t wbytes = (size_t)16 * 72 * 512;
  uint4* w; cudaMalloc(&w, wbytes); cudaMemset(w, 0, wbytes);
  long long* out; cudaMalloc(&out, 148 * 8);
  struct L { const char* n; int NF, KT; } layers[] = {{"projection 96x1152", 6, 72}, {"prenet0 256x80", 16, 5}, {"prenet1 256x256", 16, 16}, {"query 128x256", 8, 16}};
  for (int smem_kb : {20, 190, 226}) {
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    for (int grid : {1, 148}) for (auto& l : layers) {
      long long h1[148], h2[148];
      k<1><<<grid, TC_THREADS, smem_kb * 1024>>>(w, out, l.NF, l.KT, 20, 0); cudaDeviceSynchronize();
      k<1><<<grid, TC_THREADS, smem_kb * 1024>>>(w, out, l.NF, l.KT, 20, 0); cudaDeviceSynchronize();
      cudaMemcpy(h1, out, grid * 8, cudaMemcpyDeviceToHost);
      k<2><<<grid, TC_THREADS, smem_kb * 1024>>>(w, out, l.NF, l.KT, 20, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h2, out, grid * 8, cudaMemcpyDeviceToHost);
      double a1 = 0, a2 = 0; for (int i = 0; i < grid; ++i) { a1 += h1[i]; a2 += h2[i]; }
      printf("smem=%3dKB grid=%3d %-20s NU=1: %7.0f ticks  NU=2: %7.0f ticks   (%d tiles, %.0f KB)\n", smem_kb, grid, l.n, a1 / grid, a2 / grid,
             l.NF * l.KT, l.NF * l.KT * 0.5);
    }
  }
  return 0;
}
