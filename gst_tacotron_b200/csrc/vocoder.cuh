// Vocoder_Taco1 (reference: Modules/Taco2.py:234-260; CBHG :285-385, ConvBank :388-414, Highwaynet :416-434) - the pieces that are
// not a Conv1D / Dense layer.  The convolutions (conv bank, projections), the LSTM input projection and the final Dense(513) run
// on the implicit-GEMM kernels of postnet.cuh / postnet_tc.cuh over the same flat padded activation matrix
//   X[b][PADL + T + PADH][C]   (row g = b * R + PADL + t, zero rows in the padding),
// the Bidirectional LSTM on the persistent kernels of encoder.cuh.  Here:
//
//   voc_pool_kernel      MaxPool1D(pool, strides = 1, 'same') (Taco2.py:322-326) on the padded matrix: frame t = max over
//                        x[t - pad_before .. t - pad_before + pool) of the frames that exist (TF pads with -inf)
//   voc_highway_kernel   everything between the last projection conv and the LSTM, per frame (Taco2.py:341-356, 372-374, 430-434):
//                          x = proj . Wd + bd + mel                  Dense(Mel_Dim) of Conv1D_Projection + residual (optional)
//                          x = x . Wh + bh                            Dense(size) of Highwaynet (optional)
//                          x = relu(x Wr + br) * s + x * (1 - s),  s = sigmoid(x Ws + bs)          `count` Highwaynet layers
//                        fp32 FFMA over a 64-frame tile held in shared memory: 0.6 MFLOP per frame, 2 % of the vocoder
//   f32_to_f16_kernel    LSTM outputs -> fp16 operand matrix of the Dense(513) GEMM (tensor-core mode)
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace gstk {

template <typename T>
__device__ __forceinline__ float voc_ld(const T* p) {
  if constexpr (sizeof(T) == 4) return *reinterpret_cast<const float*>(p);
  else return __half2float(*reinterpret_cast<const __half*>(p));
}
template <typename T>
__device__ __forceinline__ void voc_st(T* p, float v) {
  if constexpr (sizeof(T) == 4) *reinterpret_cast<float*>(p) = v;
  else *reinterpret_cast<__half*>(p) = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}

// rows of X and Y: the flat padded matrix (Mtotal = B * R rows of C channels); padding rows of Y are written as zeros
template <typename T>
__global__ void voc_pool_kernel(const T* __restrict__ X, T* __restrict__ Y, long long Mtotal, int C, int R, int PADL, int Tn, int pool,
                                int pad_before) {
  const long long n = Mtotal * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / C;
    const int c = (int)(i - g * C);
    const int t = (int)(g % R) - PADL;
    float v = 0.f;
    if (t >= 0 && t < Tn) {
      v = -INFINITY;
      for (int j = 0; j < pool; ++j) {
        const int tt = t - pad_before + j;
        if (tt >= 0 && tt < Tn) v = fmaxf(v, voc_ld(X + (size_t)(g - pad_before + j) * C + c));
      }
    }
    voc_st(Y + i, v);
  }
}

constexpr int VH_ROWS = 64, VH_THREADS = 256, VH_MAXC = 128, VH_LD = VH_MAXC + 4, VH_MAXL = 12;
constexpr size_t VH_SMEM = (size_t)2 * VH_ROWS * VH_LD * 4;

struct VocHighwayParams {
  const void* X;        // [Mtotal][C0] last projection conv (padded matrix, fp16 or fp32)
  void* Y;              // [Mtotal][C_last]  (same element type): input matrix of the LSTM projection
  const float* resid;   // [B][T][mel] fp32 mels (the residual of Taco2.py:372), see resid_mode
  long long Mtotal;
  int R, PADL, T;
  int C0;               // channels of X
  int n_layers;         // <= VH_MAXL
  int resid_mode;       // mels added: 0 never, 1 after layer 0 (the Dense of Conv1D_Projection), 2 to the input (no such Dense)
  int type[VH_MAXL];    // 0: Dense (W [K][N], b [N]); 1: Highwaynet (W [K][2N] = [Dense_Relu | Dense_Sigmoid], b [2N])
  int N[VH_MAXL];       // output channels (<= VH_MAXC)
  const float* W[VH_MAXL];
  const float* b[VH_MAXL];
};

// One CTA = 64 frames.  Thread (n, half) = (tid % 128, tid / 128) owns output channel n of rows 32 half .. +32: per k one or two
// weight loads (coalesced over n, L1/L2 resident: 0.6 MB of weights in all) and 32 broadcast reads of x[row][k..k+3].
template <typename T>
__global__ void __launch_bounds__(VH_THREADS) voc_highway_kernel(const VocHighwayParams p) {
  extern __shared__ __align__(16) unsigned char vh_raw[];
  float (*xs)[VH_ROWS][VH_LD] = reinterpret_cast<float (*)[VH_ROWS][VH_LD]>(vh_raw);
  const int tid = threadIdx.x, n = tid & 127, half = tid >> 7;
  const long long g0 = (long long)blockIdx.x * VH_ROWS;
  const T* X = reinterpret_cast<const T*>(p.X);
  for (int i = tid; i < VH_ROWS * VH_MAXC; i += VH_THREADS) {
    const int r = i / VH_MAXC, c = i % VH_MAXC;
    const long long g = g0 + r;
    float v = (g < p.Mtotal && c < p.C0) ? voc_ld(X + (size_t)g * p.C0 + c) : 0.f;
    if (p.resid_mode == 2 && g < p.Mtotal && c < p.C0) {
      const int t = (int)(g % p.R) - p.PADL;
      if (t >= 0 && t < p.T) v += __ldg(p.resid + ((size_t)(g / p.R) * p.T + t) * p.C0 + c);
    }
    xs[0][r][c] = v;
  }
  __syncthreads();
  int cur = 0, K = p.C0;
  for (int l = 0; l < p.n_layers; ++l) {
    const int N = p.N[l];
    const bool hw = p.type[l] == 1;
    const int ldw = hw ? 2 * N : N;
    float acc0[32], acc1[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) acc0[r] = acc1[r] = 0.f;
    if (n < N) {
      const float* W = p.W[l];
      for (int k = 0; k < K; k += 4) {   // K % 4 == 0 (checked on the host)
        float w0[4], w1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          w0[e] = __ldg(W + (size_t)(k + e) * ldw + n);
          w1[e] = hw ? __ldg(W + (size_t)(k + e) * ldw + N + n) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[cur][half * 32 + r][k]);
          acc0[r] = fmaf(xv.x, w0[0], fmaf(xv.y, w0[1], fmaf(xv.z, w0[2], fmaf(xv.w, w0[3], acc0[r]))));
          if (hw) acc1[r] = fmaf(xv.x, w1[0], fmaf(xv.y, w1[1], fmaf(xv.z, w1[2], fmaf(xv.w, w1[3], acc1[r]))));
        }
      }
      const float b0 = __ldg(p.b[l] + n), b1 = hw ? __ldg(p.b[l] + N + n) : 0.f;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int row = half * 32 + r;
        float v = acc0[r] + b0;
        if (hw) {
          const float s = 1.f / (1.f + expf(-(acc1[r] + b1)));
          const float xin = xs[cur][row][n];
          v = fmaxf(v, 0.f) * s + xin * (1.f - s);
        } else if (l == 0 && p.resid_mode == 1) {
          const long long g = g0 + row;
          const int t = (int)(g % p.R) - p.PADL;
          if (g < p.Mtotal && t >= 0 && t < p.T) v += __ldg(p.resid + ((size_t)(g / p.R) * p.T + t) * N + n);
        }
        xs[cur ^ 1][row][n] = v;
      }
    }
    __syncthreads();
    cur ^= 1;
    K = N;
  }
  T* Y = reinterpret_cast<T*>(p.Y);
  for (int i = tid; i < VH_ROWS * K; i += VH_THREADS) {
    const int r = i / K, c = i % K;
    const long long g = g0 + r;
    if (g >= p.Mtotal) continue;
    const int t = (int)(g % p.R) - p.PADL;
    voc_st(Y + (size_t)g * K + c, (t >= 0 && t < p.T) ? xs[cur][r][c] : 0.f);
  }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    const __half2 lo = __floats2half2_rn(fminf(fmaxf(v.x, -65504.f), 65504.f), fminf(fmaxf(v.y, -65504.f), 65504.f));
    const __half2 hi = __floats2half2_rn(fminf(fmaxf(v.z, -65504.f), 65504.f), fminf(fmaxf(v.w, -65504.f), 65504.f));
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&lo);
    u.y = *reinterpret_cast<const unsigned*>(&hi);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}

}  // namespace gstk
