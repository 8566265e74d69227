// Micro-benchmark of the streaming ring used by the LSTM segments (diagnostics, not product code):
// producer warp (bulk copies) -> full[] -> MMA warp (tcgen05.mma, commit) -> empty[] -> producer.
//   mode 0: copy only (consumer releases the stage at once)      mode 1: copy + MMA on the copied stage
//   mode 2: MMA only (no copies; stages are static)               mode 3: copy + MMA reading a static region (no data dependence)
#include <cstdio>
#include <vector>
#include "../gst_tacotron_b200/csrc/umma.cuh"
using namespace gstk;

__device__ __forceinline__ long long clk() { long long v; asm volatile("mov.u64 %0, %%clock64;" : "=l"(v) :: "memory"); return v; }
__device__ __forceinline__ void mbar_arrive_l(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void wait_test(uint64_t* bar, uint32_t parity) {  // non-blocking test_wait spin
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {  // try_wait with a suspend-time hint
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
template <int N>
__global__ void __launch_bounds__(128) k(const uint8_t* src, long long* out, int units, int nstage, int wbytes, int mode, int split, int variant) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int SB = (variant & 8) ? 16384 : 16384 + 8192;
  if (tid == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&done, 1); mbar_fence_init(); }
  if (wid == 0) tmem_alloc(&tmem_s, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = make_idesc_bf16(128, N);
  __shared__ long long tt[4];
  long long t0 = clk();
  long long wp = 0, wc = 0;
  if (wid == 1 && mode != 2) {  // producer
    int st = 0; uint32_t ph = 0;
    for (int u = 0; u < units; ++u) {
      { const long long w0 = clk();
      if (variant & 16) wait_test(&empty[st], ph ^ 1u); else if (variant & 32) wait_hint(&empty[st], ph ^ 1u, 20); else if (variant & 4) mbar_wait_backoff(&empty[st], ph ^ 1u); else mbar_wait(&empty[st], ph ^ 1u);
      wp += clk() - w0; }
      if (elect_one()) {  // same form as tc_produce in decoder_bf16.cuh: straight-line, warp-uniform operands
        mbar_arrive_expect_tx(&full[st], 16384u + (uint32_t)wbytes);
        bulk_g2s(sm + st * SB, src + (size_t)((u * 7 + blockIdx.x) & 63) * 16384, 16384u, &full[st]);
        if (split == 2) bulk_g2s(sm + st * SB + 16384, src + (size_t)(64 + (blockIdx.x & 31)) * 16384 + (u & 1) * 8192, (uint32_t)wbytes, &full[st]);
      }
      __syncwarp();
      if (++st == nstage) { st = 0; ph ^= 1u; }
    }
    if ((tid & 31) == 0) { tt[1] = wp; tt[3] = clk() - t0; }
  } else if (wid == 2) {  // consumer
    int st = 0; uint32_t ph = 0;
    for (int u = 0; u < units; ++u) {
      { const long long w0 = clk();
      if (mode != 2) { if (variant & 16) wait_test(&full[st], ph); else if (variant & 32) wait_hint(&full[st], ph, 20); else if (variant & 2) mbar_wait_backoff(&full[st], ph); else mbar_wait(&full[st], ph); }
      wc += clk() - w0; }
      tc_fence_after();
      if (mode == 0) { if (elect_one()) mbar_arrive_l(&empty[st]); }
      else {
        const uint32_t base = smem_u32(sm) + (mode == 3 ? 7 * SB : st * SB);
        const uint64_t ad = make_desc_sw128(base), bd = make_desc_sw128(base + 16384);
        if (elect_one()) {
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) umma_bf16_ss(tmem, ad + 2 * k2, bd + 2 * k2, idesc, 1u);
          umma_commit(&empty[st]);
          if (u == units - 1) umma_commit(&done);
        }
      }
      __syncwarp();
      if (++st == nstage) { st = 0; ph ^= 1u; }
    }
    if (mode != 0) mbar_wait(&done, 0);
    if ((tid & 31) == 0) { tt[0] = clk() - t0; tt[2] = wc; }
  }
  __syncthreads();
  if (tid < 4) out[blockIdx.x * 4 + tid] = tt[tid];
  if (wid == 0) tmem_dealloc(tmem, 512);
}

int main() {
  uint8_t* src; cudaMalloc(&src, 8u << 20); cudaMemset(src, 0, 8u << 20);
  long long* out; cudaMalloc(&out, 148 * 32);
  const int smem = 8 * (16384 + 8192) + 1024;
  cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int units = 256;
  for (int grid : {1, 128})
    for (int mode : {0, 1})
      for (int variant : {0})
        for (int nstage : {2, 4, 6})
        for (int split : {1, 2}) {
          const int wbytes = split == 2 ? 8192 : 0;
          for (int rep = 0; rep < 2; ++rep) k<64><<<grid, 128, smem>>>(src, out, units, nstage, wbytes, mode, split, variant);
          cudaError_t e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> h(grid * 4); cudaMemcpy(h.data(), out, grid * 32, cudaMemcpyDeviceToHost);
          double av = 0, pw = 0, cw = 0, pt = 0; for (int b = 0; b < grid; ++b) { av += h[b * 4]; pw += h[b * 4 + 1]; cw += h[b * 4 + 2]; pt += h[b * 4 + 3]; }
          av /= grid; pw /= grid; cw /= grid; pt /= grid;
          printf("   producer: total %.0f/unit, waiting %.0f/unit   consumer waiting %.0f/unit\n", pt / units, pw / units, cw / units);
          printf("grid=%3d mode=%d wbytes=%d stages=%d: %7.0f ticks/unit  (%.1f B/clk)\n", grid, mode, wbytes, nstage, av / units, (16384.0 + wbytes) * units / av);
        }
  return 0;
}
