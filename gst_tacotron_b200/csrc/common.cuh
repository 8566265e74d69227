// Shared device helpers for libgsttaco (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define GSTK_MAX_TV 2048

namespace gstk {

// ---------------------------------------------------------------------------------------------
// Grid-wide barrier for persistent kernels launched with cudaLaunchCooperativeKernel (all CTAs
// co-resident).  Monotonic arrival counter + generation flag in separate 128 B lines so that the
// pollers do not slow the arriving atomics down.  The spin is bounded: a CTA that waits longer
// than ~2 s raises `error` and every CTA leaves the kernel (the host reports GSTK_ETIMEOUT)
// instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
struct GridBarrier {
  unsigned int count;
  unsigned int pad0[31];
  unsigned int flag;
  unsigned int pad1[31];
  unsigned int error;
  unsigned int pad2[31];
  // bf16 decoder: per (m-tile, k-block) arrival counters of the h1 operand image (one 128 B line each): phase C of a CTA starts
  // on a k-block as soon as the 4 CTAs that write it have published it, instead of after a barrier over the whole grid
  unsigned int kbcnt[32][32];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Returns false when the barrier timed out / another CTA flagged an error.  `gen` is the CTA's
// private barrier index (same sequence in every CTA).
__device__ __forceinline__ bool grid_sync(GridBarrier* gb, unsigned int nblocks, unsigned int& gen,
                                          int* ok_s) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int target = ++gen;
    int ok = 1;
    // release: orders this CTA's earlier writes (made visible to thread 0 by the bar.sync above) before the arrival.
    // Arrivals are fire-and-forget reductions and every CTA polls the monotonically increasing counter itself: the
    // critical path is one L2 one-way trip plus one poll, with no second hop through a "last arriver sets a flag" store.
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&gb->count) : "memory");
    const unsigned int want = target * nblocks;
    if (ld_acquire_u32(&gb->count) < want) {
      long long t0 = clock64();
      unsigned int spins = 0;
      while (ld_acquire_u32(&gb->count) < want) {
        if ((++spins & 1023u) == 0u) {
          if (ld_relaxed_u32(&gb->error) != 0u || clock64() - t0 > 4000000000LL) {
            atomicExch(&gb->error, 1u);
            ok = 0;
            break;
          }
        }
      }
    }
    // the acquire load above orders the later reads of this thread; the other threads are ordered by the bar.sync
    // below.  Cross-CTA data is only read through L2 (ld.global.cg / bulk copies), so no L1 invalidation is needed.
    *ok_s = ok;
  }
  __syncthreads();
  return *ok_s != 0;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (must match oracle/reference_port.py:philox4x32_10).
// counter = (item / 4, step, row, stream), key = (seed lo, seed hi).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const unsigned int hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
enum { STREAM_KEEP0 = 0, STREAM_KEEP1 = 1, STREAM_NOISE = 2 };

__device__ __forceinline__ unsigned int philox_word(unsigned long long seed, unsigned int stream,
                                                    unsigned int step, unsigned int row, unsigned int item) {
  uint4 r = philox4x32_10(make_uint4(item >> 2, step, row, stream),
                          make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  const unsigned int w = item & 3u;
  return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w));
}
__device__ __forceinline__ float philox_keep(unsigned long long seed, unsigned int stream, unsigned int step,
                                             unsigned int row, unsigned int item, float rate) {
  const float u = (float)(philox_word(seed, stream, step, row, item) >> 8) * 5.9604644775390625e-08f;
  return u >= rate ? 1.0f : 0.0f;
}
// four N(0,1) values for items 4*blk .. 4*blk+3 (Box-Muller, see oracle philox_normal)
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned int step, unsigned int row,
                                                 unsigned int blk) {
  uint4 r = philox4x32_10(make_uint4(blk, step, row, (unsigned int)STREAM_NOISE),
                          make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  const float s = 5.9604644775390625e-08f;  // 2^-24
  const float r0 = sqrtf(-2.0f * logf(((float)(r.x >> 8) + 1.0f) * s));
  const float r1 = sqrtf(-2.0f * logf(((float)(r.z >> 8) + 1.0f) * s));
  float s0, c0, s1, c1;
  sincospif(2.0f * ((float)(r.y >> 8) * s), &s0, &c0);
  sincospif(2.0f * ((float)(r.w >> 8) * s), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace gstk
