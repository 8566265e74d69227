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
#include "postnet.cuh"

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

// rows of X and Y: the flat padded matrix (Mtotal = B * R rows of C channels); padding rows of Y are written as zeros.
// One thread = 16 bytes of one row (8 halves / 4 floats; C is a multiple of that on both paths).
template <typename T>
__global__ void voc_pool_kernel(const T* __restrict__ X, T* __restrict__ Y, long long Mtotal, int C, int R, int PADL, int Tn, int pool,
                                int pad_before) {
  constexpr int V = 16 / sizeof(T);
  const int cv = C / V;
  const long long n = Mtotal * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / cv;
    const int c = (int)(i - g * cv) * V;
    const int t = (int)(g % R) - PADL;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < Tn) {
      bool first = true;
      for (int j = 0; j < pool; ++j) {
        const int tt = t - pad_before + j;
        if (tt < 0 || tt >= Tn) continue;   // TF pads with -inf: frames that do not exist never win
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + (size_t)(g - pad_before + j) * C + c));
        if (first) {
          o = v;
          first = false;
        } else if constexpr (sizeof(T) == 4) {
          float* of = reinterpret_cast<float*>(&o);
          const float* vf = reinterpret_cast<const float*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) of[e] = fmaxf(of[e], vf[e]);
        } else {
          __half2* oh = reinterpret_cast<__half2*>(&o);
          const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __hmax2(oh[e], vh[e]);
        }
      }
    }
    *reinterpret_cast<uint4*>(Y + (size_t)g * C + c) = o;
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

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core form of the same stack (handle precision "bf16": fp16 activations in shared memory, fp32 accumulation, mma.sync
// m16n8k16).  One warp owns 32 frames (two m16 tiles) from the first layer to the last - no block barrier between layers.  A layer
// is processed in chunks of 4 n-tiles: Dense = 32 output channels; Highwaynet = 16 units, tiles 0-1 = Dense_Relu columns, tiles 2-3
// = Dense_Sigmoid columns of the same units, so the gate lands in the same lane and register as the value it gates.  Weights are
// B fragments in global memory, packed on the host in the order the lanes read them ([kt][chunk][pair][lane] uint4, coalesced 512 B
// per warp load; 0.3 MB in all, L1 / L2 resident), biases in packed column order.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int VM_WARPS = 4, VM_THREADS = VM_WARPS * 32, VM_ROWS = VM_WARPS * 32, VM_LD = VH_MAXC + 8;
constexpr size_t VM_SMEM = (size_t)2 * VM_ROWS * VM_LD * 2;

struct VocHighwayMmaParams {
  const __half* X;      // [Mtotal][C0]
  __half* Y;            // [Mtotal][C_last]
  const float* resid;   // [B][T][mel] fp32
  long long Mtotal;
  int R, PADL, T, C0, n_layers, resid_mode;
  int type[VH_MAXL];    // 0 Dense, 1 Highwaynet
  int N[VH_MAXL];       // output channels
  int chunks[VH_MAXL];  // Dense: ceil(N / 32); Highwaynet: N / 16
  const uint4* W[VH_MAXL];
  const float* b[VH_MAXL];   // packed column order: [chunk][32]
};

__global__ void __launch_bounds__(VM_THREADS) voc_highway_mma_kernel(const VocHighwayMmaParams p) {
  extern __shared__ __align__(16) unsigned char vm_raw[];
  __half* xs = reinterpret_cast<__half*>(vm_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  const long long gw = (long long)blockIdx.x * VM_ROWS + warp * 32;   // first row of this warp
  __half* xw[2] = {xs + (size_t)warp * 32 * VM_LD, xs + (size_t)(VM_ROWS + warp * 32) * VM_LD};
  // load the warp's 32 rows (zero beyond C0 / Mtotal), + mels when the projection has no Dense of its own
  for (int i = lane; i < 32 * (VH_MAXC / 2); i += 32) {
    const int r = i / (VH_MAXC / 2), c = (i % (VH_MAXC / 2)) * 2;
    const long long g = gw + r;
    float2 v = make_float2(0.f, 0.f);
    if (g < p.Mtotal && c < p.C0) {
      v = __half22float2(*reinterpret_cast<const __half2*>(p.X + (size_t)g * p.C0 + c));
      if (p.resid_mode == 2) {
        const int t = (int)(g % p.R) - p.PADL;
        if (t >= 0 && t < p.T) {
          const float2 m = __ldg(reinterpret_cast<const float2*>(p.resid + ((size_t)(g / p.R) * p.T + t) * p.C0 + c));
          v.x += m.x;
          v.y += m.y;
        }
      }
    }
    *reinterpret_cast<__half2*>(xw[0] + r * VM_LD + c) = f16_sat2(v.x, v.y);
  }
  __syncwarp();
  int cur = 0, K = p.C0;
  for (int l = 0; l < p.n_layers; ++l) {
    const int N = p.N[l], KT = K >> 4, nch = p.chunks[l];
    const bool hw = p.type[l] == 1;
    const __half* xin = xw[cur];
    __half* xout = xw[cur ^ 1];
    for (int ch = 0; ch < nch; ++ch) {
      float acc[2][4][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
      const uint4* w = p.W[l] + ((size_t)ch * 2) * 32 + lane;   // [kt][chunk][pair][lane]
      for (int kt = 0; kt < KT; ++kt) {
        unsigned af[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) ldmatrix_x4(af[m], xin + (m * 16 + (lane & 15)) * VM_LD + kt * 16 + (lane >> 4) * 8);
        const uint4 b0 = __ldg(w + (size_t)kt * nch * 64), b1 = __ldg(w + (size_t)kt * nch * 64 + 32);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          mma_f16_16816(acc[m][0], af[m], b0.x, b0.y);
          mma_f16_16816(acc[m][1], af[m], b0.z, b0.w);
          mma_f16_16816(acc[m][2], af[m], b1.x, b1.y);
          mma_f16_16816(acc[m][3], af[m], b1.z, b1.w);
        }
      }
      const float* bias = p.b[l] + ch * 32;
      if (hw) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int u = ch * 16 + j * 8 + t4 * 2;   // unit of c0 (c1: u + 1)
          const float bh0 = __ldg(bias + j * 8 + t4 * 2), bh1 = __ldg(bias + j * 8 + t4 * 2 + 1);
          const float bt0 = __ldg(bias + 16 + j * 8 + t4 * 2), bt1 = __ldg(bias + 16 + j * 8 + t4 * 2 + 1);
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int row = m * 16 + g4 + hh * 8;
              const float2 xo = __half22float2(*reinterpret_cast<const __half2*>(xin + row * VM_LD + u));
              const float s0 = 1.f / (1.f + __expf(-(acc[m][j + 2][hh * 2] + bt0)));
              const float s1 = 1.f / (1.f + __expf(-(acc[m][j + 2][hh * 2 + 1] + bt1)));
              const float v0 = fmaxf(acc[m][j][hh * 2] + bh0, 0.f) * s0 + xo.x * (1.f - s0);
              const float v1 = fmaxf(acc[m][j][hh * 2 + 1] + bh1, 0.f) * s1 + xo.y * (1.f - s1);
              *reinterpret_cast<__half2*>(xout + row * VM_LD + u) = f16_sat2(v0, v1);
            }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = ch * 32 + j * 8 + t4 * 2;
          if (n >= N) continue;   // N is even
          const float b0 = __ldg(bias + j * 8 + t4 * 2), b1 = __ldg(bias + j * 8 + t4 * 2 + 1);
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int row = m * 16 + g4 + hh * 8;
              float v0 = acc[m][j][hh * 2] + b0, v1 = acc[m][j][hh * 2 + 1] + b1;
              if (l == 0 && p.resid_mode == 1) {
                const long long g = gw + row;
                const int t = (int)(g % p.R) - p.PADL;
                if (g < p.Mtotal && t >= 0 && t < p.T) {
                  const float2 m2 = __ldg(reinterpret_cast<const float2*>(p.resid + ((size_t)(g / p.R) * p.T + t) * N + n));
                  v0 += m2.x;
                  v1 += m2.y;
                }
              }
              *reinterpret_cast<__half2*>(xout + row * VM_LD + n) = f16_sat2(v0, v1);
            }
        }
      }
    }
    __syncwarp();
    cur ^= 1;
    K = N;
  }
  for (int i = lane; i < 32 * (K / 2); i += 32) {
    const int r = i / (K / 2), c = (i % (K / 2)) * 2;
    const long long g = gw + r;
    if (g >= p.Mtotal) continue;
    const int t = (int)(g % p.R) - p.PADL;
    const __half2 v = (t >= 0 && t < p.T) ? *reinterpret_cast<const __half2*>(xw[cur] + r * VM_LD + c) : __floats2half2_rn(0.f, 0.f);
    *reinterpret_cast<__half2*>(p.Y + (size_t)g * K + c) = v;
  }
}

// fp16 operand matrix of the decoder's value projection: row (b, j) = [gst[b] (S) || enc_text[b][j] (Dt)] (GST_Concated_Encoder order,
// GST.py:121-124), or a plain copy of `encodings` rows when gst is nullptr (S = 0, Dt = E).  4 channels per thread.
__global__ void value_operand_f16_kernel(const float* __restrict__ text, const float* __restrict__ gst, __half* __restrict__ out, long long rows,
                                         int Tv, int S, int Dt) {
  const int E4 = (S + Dt) / 4;
  const long long n = rows * E4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / E4;
    const int c = (int)(i - row * E4) * 4;
    const float4 v = c < S ? __ldg(reinterpret_cast<const float4*>(gst + (size_t)(row / Tv) * S + c))
                           : __ldg(reinterpret_cast<const float4*>(text + (size_t)row * Dt + (c - S)));
    const __half2 lo = f16_sat2(v.x, v.y), hi = f16_sat2(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&lo);
    u.y = *reinterpret_cast<const unsigned*>(&hi);
    *reinterpret_cast<uint2*>(out + (size_t)row * (S + Dt) + c) = u;
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
