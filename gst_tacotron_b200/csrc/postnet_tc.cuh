// Postnet conv layers on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), for the layers whose input channel count is
// a multiple of 64 (layers 1..4 at the defaults: 95 % of the Postnet's FLOPs; layer 0 with C = 80 stays on the mma.sync
// kernel of postnet.cuh).  Same flat padded activation matrix and the same epilogue contract as postnet.cuh.
//
//   GEMM row g, k-block kb (64 channels of ONE tap: tap = kb / (C/64), c0 = (kb % (C/64)) * 64):
//     A tile [128 rows][64 ch]  = X[g0 - pad_lo + tap .. +128)[c0 .. c0+64)      one 2-D TMA box, SWIZZLE_128B
//     B tile [BN rows][64 k]    = Wt[n0 .. n0+BN)[kb*64 .. +64)  (Wt = folded kernel transposed to [N][K], K-major)
//   rows outside the matrix (g0 - pad_lo < 0, tail tile) are zero-filled by TMA.
//
// Persistent kernel, one CTA per SM, 320 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (M = 128, N = BN <= 256,
// K = 16, fp16 operands, fp32 accumulators), warps 2-9 = epilogue (TMEM -> registers -> + shift, tanh -> fp16 rows of the
// next padded matrix, or + residual -> fp32 post_decodings).  4-stage 48 KB operand ring (full/empty mbarriers, stages
// released by tcgen05.commit) and TWO 256-column accumulators in TMEM, so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Tiles are handed out round-robin with the n-tile index fastest (CTAs that run together share the A tile in L2).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "postnet.cuh"
#include "umma.cuh"

namespace gstk {

struct PostTcParams {
  PostConvParams p;
  int BN;          // accumulator width of a tile (<= 256, % 16 == 0)
  int tiles_n, tiles_m;
  int KB;          // k-blocks per tile = k * C / 64
  int cpb;         // k-blocks per tap = C / 64; 0: C is not a multiple of 64 (layer 0, C = Mel_Dim = 80) - the A tensor map is
                   // then the OVERLAPPING-ROW view [rows][k*C] with row stride C (im2col by strides): k-block kb = columns
                   // [64 kb, 64 kb + 64) of that view, the tail beyond k*C zero-filled by TMA in both operands
  // ---- 2-D convolutions of the GST reference encoder (gst_tc.cuh): the flat matrix is a grid of Hb x Wb space-to-depth blocks
  // per image and the taps are the row offsets toff[] (0, 1, Wb, Wb + 1) instead of consecutive rows
  int ntap;        // 0: Conv1D (tap i = row + i); > 0: number of entries of toff
  int toff[4];
  int gmode;       // 0: Postnet / Encoder epilogues; 1: ReLU -> fp16 pixel scattered into the NEXT layer's block matrix (p.Y, row
                   // stride 4 N); 2: ReLU -> fp32 plain NHWC [b][Ho][Wo][N] (p.out); 3: ReLU -> fp16 plain NHWC (p.Y)
  int shared_rows; // > 0: ONE A box of this many rows (128 + the largest tap offset) per 64-channel block serves every tap - the
                   // taps are descriptor start addresses `offset` rows further down the same swizzled tile (SWIZZLE_128B is a
                   // function of the absolute shared-memory address, so a start row that is not a multiple of 8 needs no
                   // base-offset field: tools/ubench_desc_shift.cu) - and the stage holds the B tiles of all taps
  int Hb, Wb, Ho, Wo;        // block grid of this layer's matrix (Ho + 1, Wo + 1) and its valid outputs
  int nHb, nWb, nsh, nsw;    // gmode 1: block grid of the next matrix and the pixel shift of its odd dimensions
};

constexpr int PT_STAGES = 4, PT_THREADS = 320;
constexpr int PT_A_BYTES = 128 * 128, PT_B_BYTES = 256 * 128, PT_STAGE_BYTES = PT_A_BYTES + PT_B_BYTES;
constexpr size_t PT_SMEM = (size_t)PT_STAGES * PT_STAGE_BYTES + 1024;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void pt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// MUFU.TANH: max relative error 2^-11, the rounding of the fp16 activations it feeds
__device__ __forceinline__ float tanh_mufu(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// {lo, hi} -> f16x2, saturating at +-65504 (one CVT)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// instruction descriptor for kind::f16: fp16 x fp16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(PT_THREADS, 1)
postnet_conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PostTcParams q) {
  extern __shared__ __align__(1024) uint8_t pt_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)pt_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[PT_STAGES], empty_bar[PT_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PostConvParams& p = q.p;
  if (tid == 0) {
    for (int s = 0; s < PT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 8);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int n_tiles = q.tiles_m * q.tiles_n;
  const uint32_t stage_tx = (uint32_t)PT_A_BYTES + (uint32_t)q.BN * 128u;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    const bool leader = elect_one();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int tm = tile / q.tiles_n, tn = tile % q.tiles_n;
      const int row0 = tm * PC_BM - p.pad_lo, n0 = tn * q.BN;
      if (q.shared_rows > 0) {
        const int ntaps = q.KB / q.cpb;
        const uint32_t a_bytes = (uint32_t)q.shared_rows * 128u, a_al = (a_bytes + 1023u) & ~1023u;
        for (int cb = 0; cb < q.cpb; ++cb, ++it) {
          const int s = it % PT_STAGES;
          mbar_wait(&empty_bar[s], ((it / PT_STAGES) & 1) ^ 1);
          if (leader) {
            uint8_t* a = sm + (size_t)s * PT_STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[s], a_bytes + (uint32_t)ntaps * (uint32_t)q.BN * 128u);
            tma_load_2d(a, &tmA, cb * 64, row0, &full_bar[s]);
            for (int tap = 0; tap < ntaps; ++tap)
              tma_load_2d(a + a_al + (size_t)tap * q.BN * 128, &tmB, (tap * q.cpb + cb) * 64, n0, &full_bar[s]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int kb = 0; kb < q.KB; ++kb, ++it) {
        const int s = it % PT_STAGES;
        mbar_wait(&empty_bar[s], ((it / PT_STAGES) & 1) ^ 1);
        if (leader) {
          uint8_t* a = sm + (size_t)s * PT_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], stage_tx);
          if (q.ntap > 0) tma_load_2d(a, &tmA, (kb % q.cpb) * 64, row0 + q.toff[kb / q.cpb], &full_bar[s]);
          else if (q.cpb > 0) tma_load_2d(a, &tmA, (kb % q.cpb) * 64, row0 + kb / q.cpb, &full_bar[s]);
          else tma_load_2d(a, &tmA, kb * 64, row0, &full_bar[s]);   // overlapping-row view [rows][k*C], see below
          tma_load_2d(a + PT_A_BYTES, &tmB, kb * 64, n0, &full_bar[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, q.BN);
    uint32_t it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const uint32_t acc = lt & 1;
      mbar_wait(&acc_empty[acc], ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + acc * 256;
      if (q.shared_rows > 0) {
        const int ntaps = q.KB / q.cpb;
        const uint32_t a_al = ((uint32_t)q.shared_rows * 128u + 1023u) & ~1023u;
        for (int cb = 0; cb < q.cpb; ++cb, ++it) {
          const int s = it % PT_STAGES;
          mbar_wait(&full_bar[s], (it / PT_STAGES) & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t a_addr = smem_u32(sm + (size_t)s * PT_STAGE_BYTES);
            for (int tap = 0; tap < ntaps; ++tap) {
              const int off = q.ntap > 0 ? q.toff[tap] : tap;
              const uint64_t ad = make_desc_sw128(a_addr + (uint32_t)off * 128u);
              const uint64_t bd = make_desc_sw128(a_addr + a_al + (uint32_t)tap * (uint32_t)q.BN * 128u);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ss(d, ad + 2 * k, bd + 2 * k, idesc, (cb | tap | k) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);
            if (cb == q.cpb - 1) umma_commit(&acc_full[acc]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int kb = 0; kb < q.KB; ++kb, ++it) {
        const int s = it % PT_STAGES;
        mbar_wait(&full_bar[s], (it / PT_STAGES) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t a_addr = smem_u32(sm + (size_t)s * PT_STAGE_BYTES);
          const uint64_t ad = make_desc_sw128(a_addr), bd = make_desc_sw128(a_addr + PT_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == q.KB - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue: warps 2..9; warp w owns TMEM lanes 32 (w & 3) .. +32 (the
    // hardware's lane-quarter rule) and one half of the tile's 32-column chunks.  Two warps per scheduler: the
    // per-value chain (add, tanh, convert) is latency-bound with one.
    const int lg = warp & 3, half = (warp - 2) >> 2;
    const int nchunks = (q.BN + 31) / 32, csplit = (nchunks + 1) / 2;
    const int cbeg = (half ? csplit : 0) * 32, cend = (half ? nchunks : csplit) * 32;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int tm = tile / q.tiles_n, tn = tile % q.tiles_n;
      const uint32_t acc = lt & 1;
      mbar_wait_backoff(&acc_full[acc], (lt >> 1) & 1);
      tc_fence_after();
      const long long g = (long long)tm * PC_BM + lg * 32 + lane;
      const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + acc * 256;
      if (q.gmode != 0) {
        // GST conv layer: row g = (image b, block row, block column) = output pixel (ho, wo) when inside the valid grid
        const int per = q.Hb * q.Wb;
        const int b = (int)(g / per), rem = (int)(g - (long long)b * per), ho = rem / q.Wb, wo = rem - ho * q.Wb;
        const bool valid = g < p.Mtotal && ho < q.Ho && wo < q.Wo;
        __half* dsth = nullptr;
        float* dstf = nullptr;
        if (valid) {
          if (q.gmode == 1) {
            const int hs = ho + q.nsh, ws = wo + q.nsw;
            dsth = reinterpret_cast<__half*>(p.Y) + ((size_t)((size_t)b * q.nHb + (hs >> 1)) * q.nWb + (ws >> 1)) * 4 * p.N + (size_t)(((hs & 1) * 2 + (ws & 1)) * p.N);
          } else if (q.gmode == 3) {
            dsth = reinterpret_cast<__half*>(p.Y) + ((size_t)((size_t)b * q.Ho + ho) * q.Wo + wo) * p.N;
          } else {
            dstf = p.out + ((size_t)((size_t)b * q.Ho + ho) * q.Wo + wo) * p.N;
          }
        }
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          const int n = tn * q.BN + c0;
          if (valid && n < p.N) {   // N % 32 == 0 on this path
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + n + 4 * j));
              v[4 * j + 0] = fmaxf(v[4 * j + 0] + sh.x, 0.f); v[4 * j + 1] = fmaxf(v[4 * j + 1] + sh.y, 0.f);
              v[4 * j + 2] = fmaxf(v[4 * j + 2] + sh.z, 0.f); v[4 * j + 3] = fmaxf(v[4 * j + 3] + sh.w, 0.f);
            }
            if (dsth) {
              uint4* d = reinterpret_cast<uint4*>(dsth + n);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                d[j] = make_uint4(pack_f16x2_sat(v[8 * j], v[8 * j + 1]), pack_f16x2_sat(v[8 * j + 2], v[8 * j + 3]), pack_f16x2_sat(v[8 * j + 4], v[8 * j + 5]),
                                  pack_f16x2_sat(v[8 * j + 6], v[8 * j + 7]));
            } else {
              float4* d = reinterpret_cast<float4*>(dstf + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) pt_mbar_arrive(&acc_empty[acc]);
        continue;
      }
      const int r = (int)(g % p.R);
      const bool in_range = g < p.Mtotal;
      const bool valid = in_range && r >= p.PADL && r < p.PADL + p.T;
      const long long bt = (g / p.R) * p.T + (r - p.PADL);
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        const int n = tn * q.BN + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n + 4 * j < p.N) sh = __ldg(reinterpret_cast<const float4*>(p.shift + n + 4 * j));
          v[4 * j + 0] += sh.x; v[4 * j + 1] += sh.y; v[4 * j + 2] += sh.z; v[4 * j + 3] += sh.w;
        }
        if (p.use_tanh == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tanh_mufu(v[j]);
        } else if (p.use_tanh == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.Y) {
          if (in_range) {
            uint4 o[4];
            uint32_t* ou = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int j = 0; j < 16; ++j) ou[j] = valid ? pack_f16x2_sat(v[2 * j], v[2 * j + 1]) : 0u;
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.Y) + (size_t)g * p.N + n);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + 8 * j < p.N && c0 + 8 * j < q.BN) dst[j] = o[j];   // BN % 32 != 0: the last chunk reads past the tile
          }
        } else if (valid && p.ldo) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < p.n_valid && c0 + j < q.BN) p.out[(size_t)bt * p.ldo + n + j] = v[j];
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int nn = n + 4 * j;
            if (nn < p.N && c0 + 4 * j < q.BN) {
              float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.resid) rs = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)bt * p.N + nn));
              *reinterpret_cast<float4*>(p.out + (size_t)bt * p.N + nn) =
                  make_float4(v[4 * j + 0] + rs.x, v[4 * j + 1] + rs.y, v[4 * j + 2] + rs.z, v[4 * j + 3] + rs.w);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) pt_mbar_arrive(&acc_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace gstk
