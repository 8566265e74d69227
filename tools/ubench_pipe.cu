// Micro-benchmark of the copy/MMA pipeline primitives, warp-uniform + elect.sync (diagnostics, not product code).
#include <cstdio>
#include <vector>
#include "../gst_tacotron_b200/csrc/umma.cuh"
using namespace gstk;

template <int N>
__device__ __forceinline__ long long mma_burst(uint32_t tmem, uint64_t ad, uint64_t bd, int count, int chains, uint64_t* bar, uint32_t par) {
  const uint32_t idesc = make_idesc_bf16(128, N);
  const long long t0 = clock64();
  if (elect_one()) {
    for (int i = 0; i < count; ++i) umma_bf16_ss(tmem + (uint32_t)(N * (i % chains)), ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
    umma_commit(bar);
  }
  __syncwarp();
  const long long t1 = clock64();
  mbar_wait(bar, par);
  const long long t2 = clock64();
  return ((t1 - t0) << 32) | (t2 - t0);
}

__global__ void __launch_bounds__(128) k(const uint8_t* src, long long* out, int rows) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  if (wid == 0) tmem_alloc(&tmem_s, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_s;
  uint8_t* a_s = sm; uint8_t* b_s = sm + 4 * 16384;  // b: up to 256 rows x 128 B = 32 KB
  if (wid == 1) {
    long long r[16];
    // copy issue cost: 8 x (expect_tx + bulk copy) back to back, then wait all
    long long t0 = clock64();
    if (elect_one()) for (int i = 0; i < 8; ++i) { mbar_arrive_expect_tx(&bars[i], rows * 128); bulk_g2s(a_s + (i & 3) * 16384, src + (size_t)i * 16384, rows * 128, &bars[i]); }
    __syncwarp();
    long long t1 = clock64();
    for (int i = 0; i < 8; ++i) mbar_wait(&bars[i], 0);
    long long t2 = clock64();
    r[0] = t1 - t0; r[1] = t2 - t0;
    // try_wait cost on a completed barrier x 16
    t0 = clock64();
    int acc = 0;
    for (int i = 0; i < 16; ++i) acc += mbar_try_wait(&bars[i & 7], 0);
    t1 = clock64();
    r[2] = (t1 - t0) + (acc == 12345);
    tc_fence_after();
    const uint64_t ad = make_desc_sw128(smem_u32(a_s)), bd = make_desc_sw128(smem_u32(b_s));
    long long v;
    v = mma_burst<32>(tmem, ad, bd, 64, 1, &bars[8], 0);  r[3] = v >> 32; r[4] = v & 0xffffffff;
    v = mma_burst<32>(tmem, ad, bd, 64, 8, &bars[9], 0);  r[5] = v >> 32; r[6] = v & 0xffffffff;
    v = mma_burst<64>(tmem, ad, bd, 64, 4, &bars[10], 0); r[7] = v >> 32; r[8] = v & 0xffffffff;
    v = mma_burst<128>(tmem, ad, bd, 64, 2, &bars[11], 0); r[9] = v >> 32; r[10] = v & 0xffffffff;
    v = mma_burst<256>(tmem, ad, bd, 64, 1, &bars[12], 0); r[11] = v >> 32; r[12] = v & 0xffffffff;
    // commit cost x 8
    t0 = clock64();
    if (elect_one()) for (int i = 0; i < 8; ++i) umma_commit(&bars[13]);
    __syncwarp();
    t1 = clock64();
    r[13] = t1 - t0;
    // clock64 cost
    t0 = clock64(); long long x = 0; for (int i = 0; i < 16; ++i) x += clock64(); t1 = clock64();
    r[14] = (t1 - t0) + (x == 1);
    r[15] = 0;
    if ((tid & 31) == 0) for (int i = 0; i < 16; ++i) out[(size_t)blockIdx.x * 16 + i] = r[i];
  }
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem, 512);
}

int main() {
  uint8_t* src; cudaMalloc(&src, 32 * 16384); cudaMemset(src, 0, 32 * 16384);
  long long* out; cudaMalloc(&out, 148 * 16 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const char* names[16] = {"8x(expect+copy) issue", "8 copies issue+land", "16x try_wait(done)", "64 MMA N=32 1ch issue", "  .. issue+complete",
    "64 MMA N=32 8ch issue", "  .. issue+complete", "64 MMA N=64 4ch issue", "  .. issue+complete", "64 MMA N=128 2ch issue", "  .. issue+complete",
    "64 MMA N=256 1ch issue", "  .. issue+complete", "8x commit issue", "16x clock64", "-"};
  for (int grid : {1, 128}) for (int rows : {1, 128}) {
    for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, 120 * 1024>>>(src, out, rows);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(grid * 16); cudaMemcpy(h.data(), out, grid * 128, cudaMemcpyDeviceToHost);
    printf("grid=%d rows=%d:\n", grid, rows);
    for (int i = 0; i < 15; ++i) { long long mn = 1LL << 60, mx = 0; double av = 0; for (int c = 0; c < grid; ++c) { long long v = h[c * 16 + i]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; av += v; } printf("  %-26s min %7lld avg %9.0f max %7lld\n", names[i], mn, av / grid, mx); }
  }
  return 0;
}
