// Blackwell (sm_100a) primitives used by the bf16 decoder: mbarrier, bulk async copy (TMA engine,
// SASS UBLKCP), tcgen05.mma (SASS UTCHMMA) with shared-memory descriptors, TMEM alloc / tcgen05.ld.
// Descriptor bit layouts follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction
// descriptor" tables (cross-checked against CUTLASS cute/arch/mma_sm100_desc.hpp field lists).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gstk {

// ---- canonical K-major SWIZZLE_128B operand tile ------------------------------------------------
// rows x 64 bf16 (128 B per row), 8-row x 128 B swizzle atoms (1024 B, must be 1024 B aligned),
// 16-byte chunk c of row r is stored at chunk (c ^ (r & 7)).
constexpr int KB_ELEMS = 64;            // bf16 elements per K-block row (128 B)
constexpr int KB_ROW_BYTES = 128;
__host__ __device__ inline size_t sw128_offset_bytes(int row, int k /*0..63*/) {
  const int chunk = (k >> 3) ^ (row & 7);
  return (size_t)row * KB_ROW_BYTES + (size_t)chunk * 16 + (size_t)(k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (elect.sync).  The copy / MMA issuing code runs warp-uniformly and only
// the instruction itself is predicated with this: in a divergent single-lane branch the compiler has to
// wrap every uniform-datapath instruction (UTCHMMA, UBLKCP) in an R2UR vote loop (~50-100 cycles each).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe: try_wait may SUSPEND the thread for a hardware-defined time when the phase is not complete (it is a
// sleeping wait, not a poll) - inside a loop that polls something else at the same time, use this one.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU - trap after ~2 s instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Many-waiter variant for the epilogue warps: one lane polls with a back-off so that 8 spinning warps do not
// flood the mbarrier / shared-memory pipe the copy+MMA warp depends on; the rest of the warp parks at syncwarp.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    if (!mbar_try_wait(bar, parity)) {
      const long long t0 = clock64();
      do {
        __nanosleep(40);
        if (clock64() - t0 > 4000000000LL) __trap();
      } while (!mbar_try_wait(bar, parity));
    }
  }
  __syncwarp();
}

// ---- bulk async copy global -> shared (completes on an mbarrier), bytes % 16 == 0 -----------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- tcgen05 --------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                        // [16,30) leading byte offset >> 4 (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                        // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // [61,64) layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T ; one thread issues
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, descriptors given by their low words (start address | LBO); the constant high word (SBO = 1024 B, version,
// SWIZZLE_128B) is attached inside the asm block so that no 64-bit descriptor arithmetic is generated per instruction.
__device__ __forceinline__ void umma_bf16_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 consecutive fp32 columns (the LSTM cell state of one batch row lives in TMEM between steps)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- self test: D[128 x 32] = A[128 x 64*KB] . B[32 x 64*KB]^T from pre-swizzled global images ----
// A image: [kb][128 rows][128 B], B image: [kb][32 rows][128 B] (both sw128).  One CTA, 128 threads.
__global__ void __launch_bounds__(128) umma_selftest_kernel(const __nv_bfloat16* __restrict__ a_img,
                                                             const __nv_bfloat16* __restrict__ b_img, int KB,
                                                             float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = sm;                              // KB x 16 KB
  uint8_t* b_s = sm + (size_t)KB * 16384;         // KB x 4 KB
  const int tid = threadIdx.x, wid = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (wid == 0) tmem_alloc(&tmem_base_s, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], (uint32_t)KB * (16384 + 4096));
    for (int kb = 0; kb < KB; ++kb) {
      bulk_g2s(a_s + (size_t)kb * 16384, (const uint8_t*)a_img + (size_t)kb * 16384, 16384, &bars[0]);
      bulk_g2s(b_s + (size_t)kb * 4096, (const uint8_t*)b_img + (size_t)kb * 4096, 4096, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 32);
    for (int kb = 0; kb < KB; ++kb) {
      const uint64_t ad = make_desc_sw128(smem_u32(a_s + (size_t)kb * 16384));
      const uint64_t bd = make_desc_sw128(smem_u32(b_s + (size_t)kb * 4096));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(wid * 32) << 16), v);
#pragma unroll
  for (int i = 0; i < 32; ++i) d_out[(size_t)tid * 32 + i] = v[i];
  tc_fence_before();
  __syncthreads();
  if (wid == 0) tmem_dealloc(tmem, 32);
}

}  // namespace gstk
