// Can a tcgen05 SWIZZLE_128B shared-memory descriptor start at a ROW that is not a multiple of 8 (not 1024-byte aligned)?
// (diagnostics, not product code)  The GST conv layers read four taps of the same activation rows (row offsets 0, 1, Wb, Wb + 1):
// if the A operand of tap j can be addressed as "the same shared-memory tile, s rows further down", one TMA box serves all four.
// A image: 192 rows x 64 bf16, swizzled as TMA writes it (16-byte chunk index XOR (row & 7), rows 128 B apart).
// D[128 x 32] = A[s .. s + 128) . B^T for several s, with the descriptor's base-offset field (bits 49-51) = 0 and = s & 7.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_desc_shift tools/ubench_desc_shift.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include "../gst_tacotron_b200/csrc/umma.cuh"
using namespace gstk;

constexpr int ROWS = 192;

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* a_img, const __nv_bfloat16* b_img, float* d_out, int shift, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = sm;                 // 192 x 128 B
  uint8_t* b_s = sm + ROWS * 128;    // 32 x 128 B (1024-aligned: 192 * 128 = 24 KB)
  const int tid = threadIdx.x, wid = tid >> 5;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
  if (wid == 0) tmem_alloc(&tmem_base_s, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], ROWS * 128 + 4096);
    bulk_g2s(a_s, a_img, ROWS * 128, &bars[0]);
    bulk_g2s(b_s, b_img, 4096, &bars[0]);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 32);
    uint64_t ad = make_desc_sw128(smem_u32(a_s + (size_t)shift * 128));
    if (use_base_offset) ad |= (uint64_t)(shift & 7) << 49;
    const uint64_t bd = make_desc_sw128(smem_u32(b_s));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem, ad + 2 * kk, bd + 2 * kk, idesc, kk ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(wid * 32) << 16), v);
  for (int i = 0; i < 32; ++i) d_out[(size_t)tid * 32 + i] = v[i];
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> A((size_t)ROWS * 64), B(32 * 64);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((int)(s >> 20) % 17 - 8) / 8.f; };   // exactly representable
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  std::vector<__nv_bfloat16> a_img((size_t)ROWS * 64), b_img(32 * 64);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < 64; ++c) a_img[sw128_offset_bytes(r, c) / 2] = __float2bfloat16(A[(size_t)r * 64 + c]);
  for (int r = 0; r < 32; ++r)
    for (int c = 0; c < 64; ++c) b_img[sw128_offset_bytes(r, c) / 2] = __float2bfloat16(B[(size_t)r * 64 + c]);
  void *da, *db, *dd;
  cudaMalloc(&da, a_img.size() * 2); cudaMalloc(&db, b_img.size() * 2); cudaMalloc(&dd, 128 * 32 * 4);
  cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = ROWS * 128 + 4096 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> D(128 * 32);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 1, 2, 3, 7, 8, 9, 16, 41, 42, 63}) {
      k<<<1, 128, smem>>>((const __nv_bfloat16*)da, (const __nv_bfloat16*)db, (float*)dd, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset_field %d: %s\n", shift, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
      double err = 0;
      int bad_rows = 0;
      for (int i = 0; i < 128; ++i) {
        double rowerr = 0;
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int c = 0; c < 64; ++c) ref += (double)A[(size_t)(i + shift) * 64 + c] * B[n * 64 + c];
          rowerr = fmax(rowerr, fabs(ref - D[i * 32 + n]));
        }
        err = fmax(err, rowerr);
        bad_rows += rowerr > 1e-3;
      }
      printf("shift %2d  base-offset field %s: max |err| = %.4f, wrong rows %d / 128\n", shift, bo ? "= shift & 7" : "= 0        ", err, bad_rows);
    }
  return 0;
}
