// Postnet (reference: Modules/Taco2.py:130-149 construction, :230 call):
//   post_decodings = Sequential(5 x [Conv1D(k=5, 'same', no bias) -> BatchNormalization -> tanh (first 3) -> Dropout(off)])(decodings)
//                    + decodings
// as implicit-GEMM kernels.  An activation tensor lives in HBM as ONE flat matrix of zero-padded frames
//   X[b][PADL + T + PADH][C]   (R = PADL + T + PADH rows per utterance, row g = b*R + PADL + t)
// so that the im2col row of output frame g is the CONTIGUOUS run X[g - pad_lo .. g - pad_lo + k) x C: the convolution is a
// plain GEMM  Y[g, :] = A[g, :] . W  with  A row pointer = X + (g - pad_lo) * C,  lda = C,  K = k * C,  W = kernel [k*C, N]
// (BatchNormalization scale folded into W on the host, shift applied in the epilogue).  Rows of the output whose frame index
// falls into the padding are written as zeros by the epilogue, which keeps the 'same' padding of the next layer intact
// without a memset; what such rows read (neighbouring utterance / slack) never reaches a valid output.
//
//   postnet_conv_f16_kernel   the tensor-core mode (handle precision "bf16"): fp16 operands, fp32 accumulation,
//                             mma.sync.m16n8k16 + ldmatrix, 3-stage cp.async ring.  fp16 rather than bf16 operands: the Postnet's
//                             activations are bounded (BatchNormalization + tanh) and its output has magnitude ~8, so the 1e-2
//                             ABSOLUTE tolerance needs the 11-bit mantissa (bf16 operands: 3e-2 max error, fp16: 3e-3; same speed).
//                             Stores saturate at +-65504.
//   postnet_conv_f32_kernel   fp32 FFMA, 8x8 register tile - the exact mode (1e-4 parity)
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace gstk {

struct PostConvParams {
  const void* X;        // padded input activations (row 0 of the flat matrix; PADL slack rows lie before it)
  const void* W;        // [K][N] folded conv kernel (fp16 or fp32)
  const float* shift;   // [N] beta - mean * scale
  void* Y;              // padded output activations [Mtotal][N] (same R), or nullptr on the last layer
  const float* resid;   // last layer: decodings [B][T][N] fp32 added to the result, or nullptr
  float* out;           // last layer: post_decodings [B][T][N] fp32
  long long Mtotal;     // B * R
  int C, K, N;          // input channels, k*C, output channels
  int pad_lo;           // (k-1)/2  ('same', stride 1)
  int R, PADL, T;
  int use_tanh;         // activation code, see postnet_act
  int ldo, n_valid;     // fp32 `out` path with a row stride other than N (Vocoder_Taco1's Dense(513): N is padded to the tile grid):
                        // ldo > 0 -> out[bt * ldo + n] for n < n_valid, scalar stores; ldo == 0 -> row stride N, vector stores
};

constexpr int PC_BM = 128, PC_BN = 128, PC_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ __half2 f16_sat2(float a, float b) {
  return __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
}

// [B][T][C] fp32 -> padded flat matrix (fp32 or fp16), zero rows in the padding
template <typename OutT>
__global__ void postnet_pad_kernel(const float* __restrict__ in, OutT* __restrict__ X, long long Mtotal, int C, int R, int PADL, int T) {
  const long long n4 = Mtotal * (C / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / (C / 4);
    const int c = (int)(i % (C / 4)) * 4;
    const int r = (int)(g % R);
    const long long b = g / R;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= PADL && r < PADL + T) v = __ldg(reinterpret_cast<const float4*>(in + ((size_t)(b * T + (r - PADL))) * C + c));
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(X) + (size_t)g * C + c) = v;
    } else {
      __half2 lo = f16_sat2(v.x, v.y), hi = f16_sat2(v.z, v.w);
      uint2 u;
      u.x = *reinterpret_cast<unsigned*>(&lo);
      u.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(X) + (size_t)g * C + c) = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// fp16 tensor-core implicit GEMM.  CTA tile 128 x 128, BK = 32, 8 warps as 2 (m) x 4 (n): warp tile 64 x 32.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PCB_BK = 32, PCB_STAGES = 3;
constexpr int PCB_A_LD = PCB_BK + 8;    // fp16 elements per A row in shared memory (80 B: ldmatrix conflict-free)
constexpr int PCB_B_LD = PC_BN + 8;     // 272 B
constexpr int PCB_A_STAGE = PC_BM * PCB_A_LD, PCB_B_STAGE = PCB_BK * PCB_B_LD;  // elements
constexpr size_t PCB_SMEM = (size_t)PCB_STAGES * (PCB_A_STAGE + PCB_B_STAGE) * 2;

__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* smem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], const void* smem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// activation code of a layer: 0 none, 1 tanh (Postnet, Taco2.py:145-146), 2 ReLU (Encoder, Taco2.py:35)
__device__ __forceinline__ float postnet_act(float v, int act) { return act == 1 ? tanhf(v) : act == 2 ? fmaxf(v, 0.f) : v; }

__global__ void __launch_bounds__(PC_THREADS) postnet_conv_f16_kernel(const PostConvParams p) {
  extern __shared__ __align__(16) unsigned char pc_smem[];
  __half* As = reinterpret_cast<__half*>(pc_smem);
  __half* Bs = As + PCB_STAGES * PCB_A_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const long long m0 = (long long)blockIdx.y * PC_BM;
  const int n0 = blockIdx.x * PC_BN;
  const __half* X = reinterpret_cast<const __half*>(p.X);
  const __half* W = reinterpret_cast<const __half*>(p.W);
  const int KT = (p.K + PCB_BK - 1) / PCB_BK;

  auto load_stage = [&](int kt, int stage) {
    const int k0 = kt * PCB_BK;
    __half* a = As + stage * PCB_A_STAGE;
    __half* b = Bs + stage * PCB_B_STAGE;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = tid + j * PC_THREADS;
      {  // A: 128 rows x 4 chunks of 8 halves
        const int row = c >> 2, kc = (c & 3) * 8;
        const bool ok = k0 + kc < p.K;
        const __half* src = X + (m0 + row - p.pad_lo) * (long long)p.C + (ok ? k0 + kc : 0);
        cp_async16(a + row * PCB_A_LD + kc, src, ok);
      }
      {  // B: 32 k-rows x 16 chunks of 8 halves
        const int kr = c >> 4, nc = (c & 15) * 8;
        const bool ok = (k0 + kr < p.K) && (n0 + nc < p.N);
        const __half* src = ok ? W + (size_t)(k0 + kr) * p.N + n0 + nc : W;
        cp_async16(b + kr * PCB_B_LD + nc, src, ok);
      }
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

#pragma unroll
  for (int s = 0; s < PCB_STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<PCB_STAGES - 2>();
    __syncthreads();
    {  // refill the stage consumed in the previous iteration
      const int nk = kt + PCB_STAGES - 1;
      if (nk < KT) load_stage(nk, nk % PCB_STAGES);
      cp_async_commit();
    }
    const __half* a = As + (kt % PCB_STAGES) * PCB_A_STAGE;
    const __half* b = Bs + (kt % PCB_STAGES) * PCB_B_STAGE;
#pragma unroll
    for (int kk = 0; kk < PCB_BK / 16; ++kk) {
      unsigned af[4][4], bf[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        ldmatrix_x4(af[i], a + (wm * 64 + i * 16 + (lane & 15)) * PCB_A_LD + kk * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int mat = lane >> 3;
        ldmatrix_x4_trans(bf[jp], b + (kk * 16 + (mat & 1) * 8 + (lane & 7)) * PCB_B_LD + wn * 32 + jp * 16 + (mat >> 1) * 8);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16_16816(acc[i][j], af[i], bf[j >> 1][(j & 1) * 2], bf[j >> 1][(j & 1) * 2 + 1]);
    }
  }
  cp_async_wait<0>();

  // epilogue: + shift, tanh, fp16 store into the next padded matrix (zeros in padding rows) / residual + fp32 store
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const long long g = m0 + wm * 64 + i * 16 + (lane >> 2) + half * 8;
      if (g >= p.Mtotal) continue;
      const int r = (int)(g % p.R);
      const bool valid = r >= p.PADL && r < p.PADL + p.T;
      const long long bt = (g / p.R) * p.T + (r - p.PADL);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn * 32 + j * 8 + (lane & 3) * 2;
        if (n >= p.N) continue;
        float v0 = acc[i][j][half * 2 + 0] + __ldg(p.shift + n);
        float v1 = acc[i][j][half * 2 + 1] + __ldg(p.shift + n + 1);
        v0 = postnet_act(v0, p.use_tanh);
        v1 = postnet_act(v1, p.use_tanh);
        if (p.Y) {
          const __half2 o = valid ? f16_sat2(v0, v1) : __floats2half2_rn(0.f, 0.f);
          *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p.Y) + (size_t)g * p.N + n) = o;
        } else if (valid && p.ldo) {
          if (n < p.n_valid) p.out[(size_t)bt * p.ldo + n] = v0;
          if (n + 1 < p.n_valid) p.out[(size_t)bt * p.ldo + n + 1] = v1;
        } else if (valid) {
          float2 rs = make_float2(0.f, 0.f);
          if (p.resid) rs = __ldg(reinterpret_cast<const float2*>(p.resid + (size_t)bt * p.N + n));
          *reinterpret_cast<float2*>(p.out + (size_t)bt * p.N + n) = make_float2(v0 + rs.x, v1 + rs.y);
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32 FFMA implicit GEMM (exact mode).  CTA tile 128 x 128, BK = 16, thread tile 8 rows (ty + 16 i) x 8 columns
// (tx*4 .. +3 and 64 + tx*4 .. +3).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PCF_BK = 16, PCF_STAGES = 3;
constexpr int PCF_A_LD = PCF_BK + 4;    // floats per A row in shared memory
constexpr int PCF_B_LD = PC_BN;
constexpr int PCF_A_STAGE = PC_BM * PCF_A_LD, PCF_B_STAGE = PCF_BK * PCF_B_LD;
constexpr size_t PCF_SMEM = (size_t)PCF_STAGES * (PCF_A_STAGE + PCF_B_STAGE) * 4;

__global__ void __launch_bounds__(PC_THREADS) postnet_conv_f32_kernel(const PostConvParams p) {
  extern __shared__ __align__(16) unsigned char pc_smem[];
  float* As = reinterpret_cast<float*>(pc_smem);
  float* Bs = As + PCF_STAGES * PCF_A_STAGE;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.y * PC_BM;
  const int n0 = blockIdx.x * PC_BN;
  const float* X = reinterpret_cast<const float*>(p.X);
  const float* W = reinterpret_cast<const float*>(p.W);
  const int KT = (p.K + PCF_BK - 1) / PCF_BK;

  auto load_stage = [&](int kt, int stage) {
    const int k0 = kt * PCF_BK;
    float* a = As + stage * PCF_A_STAGE;
    float* b = Bs + stage * PCF_B_STAGE;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = tid + j * PC_THREADS;
      {  // A: 128 rows x 4 chunks of 4 floats
        const int row = c >> 2, kc = (c & 3) * 4;
        const bool ok = k0 + kc < p.K;
        const float* src = X + (m0 + row - p.pad_lo) * (long long)p.C + (ok ? k0 + kc : 0);
        cp_async16(a + row * PCF_A_LD + kc, src, ok);
      }
      {  // B: 16 k-rows x 32 chunks of 4 floats
        const int kr = c >> 5, nc = (c & 31) * 4;
        const bool ok = (k0 + kr < p.K) && (n0 + nc < p.N);
        const float* src = ok ? W + (size_t)(k0 + kr) * p.N + n0 + nc : W;
        cp_async16(b + kr * PCF_B_LD + nc, src, ok);
      }
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < PCF_STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<PCF_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + PCF_STAGES - 1;
      if (nk < KT) load_stage(nk, nk % PCF_STAGES);
      cp_async_commit();
    }
    const float* a = As + (kt % PCF_STAGES) * PCF_A_STAGE;
    const float* b = Bs + (kt % PCF_STAGES) * PCF_B_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < PCF_BK; k4 += 4) {
      float4 av[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(a + (ty + 16 * i) * PCF_A_LD + k4);
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        const float4 b0 = *reinterpret_cast<const float4*>(b + (k4 + kq) * PCF_B_LD + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(b + (k4 + kq) * PCF_B_LD + 64 + tx * 4);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float aa = kq == 0 ? av[i].x : kq == 1 ? av[i].y : kq == 2 ? av[i].z : av[i].w;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(aa, bv[j], acc[i][j]);
        }
      }
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long g = m0 + ty + 16 * i;
    if (g >= p.Mtotal) continue;
    const int r = (int)(g % p.R);
    const bool valid = r >= p.PADL && r < p.PADL + p.T;
    const long long bt = (g / p.R) * p.T + (r - p.PADL);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      if (n >= p.N) continue;   // N % 4 == 0
      const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + n));
      float4 v;
      v.x = postnet_act(acc[i][h * 4 + 0] + sh.x, p.use_tanh);
      v.y = postnet_act(acc[i][h * 4 + 1] + sh.y, p.use_tanh);
      v.z = postnet_act(acc[i][h * 4 + 2] + sh.z, p.use_tanh);
      v.w = postnet_act(acc[i][h * 4 + 3] + sh.w, p.use_tanh);
      if (p.Y) {
        if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.Y) + (size_t)g * p.N + n) = v;
      } else if (valid && p.ldo) {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < p.n_valid) p.out[(size_t)bt * p.ldo + n + e] = vv[e];
      } else if (valid) {
        float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.resid) rs = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)bt * p.N + n));
        *reinterpret_cast<float4*>(p.out + (size_t)bt * p.N + n) = make_float4(v.x + rs.x, v.y + rs.y, v.z + rs.z, v.w + rs.w);
      }
    }
  }
}

}  // namespace gstk
