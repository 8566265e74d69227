// Which TMEM lanes does a cta_group::1 tcgen05.mma with M = 64 write?  (diagnostics, not product code)
// A[64 x 64] has row i = (i + 1) in k = 0, B[32 x 64] has row n = 1 in k = 0  =>  D[i][n] = i + 1.
// The whole 128-lane x 32-column TMEM block is pre-filled with -1 and dumped after the MMA.
// Also times one N = 32 MMA stream at M = 64 vs M = 128 (cycles per k-block of 4 MMAs, operands resident).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_m64 tools/ubench_m64.cu
#include <cstdio>
#include <vector>
#include "../gst_tacotron_b200/csrc/umma.cuh"
using namespace gstk;

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* a_img, const __nv_bfloat16* b_img, float* d_out, long long* cyc, int M, int reps) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = sm;            // 16 KB (128 rows x 128 B; M = 64 uses the first 8 KB)
  uint8_t* b_s = sm + 16384;    // 4 KB
  const int tid = threadIdx.x, wid = tid >> 5;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  if (wid == 0) tmem_alloc(&tmem_base_s, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  {  // pre-fill this warp's lane quarter with -1
    uint32_t m1 = __float_as_uint(-1.0f);
    for (int c = 0; c < 32; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tmem + ((uint32_t)(wid * 32) << 16) + c), "r"(m1) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 16384 + 4096);
    bulk_g2s(a_s, a_img, 16384, &bars[0]);
    bulk_g2s(b_s, b_img, 4096, &bars[0]);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(M, 32);
    const uint64_t ad = make_desc_sw128(smem_u32(a_s)), bd = make_desc_sw128(smem_u32(b_s));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem, ad + 2 * kk, bd + 2 * kk, idesc, (r | kk) ? 1u : 0u);
    umma_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(wid * 32) << 16), v);
  for (int i = 0; i < 32; ++i) d_out[(size_t)tid * 32 + i] = v[i];
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem, 32);
}

// tanh throughput: f32 vs packed f16x2 / bf16x2 (results per clock per SM)
template <int MODE>
__global__ void kt(float* out, long long* cyc, int iters) {
  float a[8]; unsigned int h[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; h[i] = 0x38003400u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      else if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      else asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + (float)h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  std::vector<__nv_bfloat16> A(128 * 64, __float2bfloat16(0.f)), B(32 * 64, __float2bfloat16(0.f));
  for (int i = 0; i < 128; ++i) A[sw128_offset_bytes(i, 0) / 2] = __float2bfloat16((float)(i + 1));
  for (int n = 0; n < 32; ++n) B[sw128_offset_bytes(n, 0) / 2] = __float2bfloat16(1.f);
  __nv_bfloat16 *a, *b; float* d; long long* cyc;
  cudaMalloc(&a, A.size() * 2); cudaMalloc(&b, B.size() * 2); cudaMalloc(&d, 128 * 32 * 4); cudaMalloc(&cyc, 1024);
  cudaMemcpy(a, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(b, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024);
  std::vector<float> D(128 * 32);
  for (int M : {64, 128}) {
    k<<<1, 128, 24 * 1024>>>(a, b, d, cyc, M, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d: %s\n", M, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), d, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("M=%d: lane -> D[.][0] (and [.][31]):\n", M);
    for (int l = 0; l < 128; ++l) printf("%s%3d:%g/%g", (l % 8) ? "  " : "\n  ", l, D[l * 32], D[l * 32 + 31]);
    printf("\n");
    for (int reps : {16, 64}) {
      k<<<1, 128, 24 * 1024>>>(a, b, d, cyc, M, reps);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("M=%d N=32: %d k-blocks (4 MMAs each) in %lld cycles = %.1f per k-block\n", M, reps, c, (double)c / reps);
    }
  }
  float* out; cudaMalloc(&out, 1 << 20);
  const int iters = 2000;
  const char* names[3] = {"tanh.approx.f32", "tanh.approx.f16x2", "tanh.approx.bf16x2"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      if (mode == 0) kt<0><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 1) kt<1><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 2) kt<2><<<1, warps * 32>>>(out, cyc, iters);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-20s warps=%2d: %.1f instr-results/clk/SM (x2 values for the packed forms)\n", names[mode], warps, warps * 32.0 * iters * 8 / c);
    }
  return 0;
}
