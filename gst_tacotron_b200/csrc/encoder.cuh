// Text Encoder (reference: Modules/Taco2.py:12-51):
//   Embedding -> 3 x [Conv1D(k=5,'same',no bias) -> BatchNormalization -> ReLU -> Dropout(off)] -> Bidirectional(LSTM(256)).
// The conv stack and the LSTM input projections (a k = 1 "convolution" onto the 2 x 4u gate pre-activations) run on the
// implicit-GEMM kernels of postnet.cuh / postnet_tc.cuh over the same flat zero-padded token matrix; this file holds the
// embedding gather that builds that matrix and the recurrent half of the two LSTMs.
#pragma once
#include "postnet.cuh"

namespace gstk {

// tokens [B][T] -> padded flat matrix X[b][PADL + T + PADH][E] (fp32 or fp16), zero rows in the padding.  Ids outside
// [0, vocab) give a zero row (what tf.gather does on a GPU; the reference's feeder never produces them, Feeder.py:166-180).
template <typename OutT>
__global__ void encoder_embed_pad_kernel(const int* __restrict__ tokens, const float* __restrict__ table, OutT* __restrict__ X,
                                         long long Mtotal, int E, int R, int PADL, int T, int vocab) {
  const long long n4 = Mtotal * (E / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / (E / 4);
    const int c = (int)(i % (E / 4)) * 4;
    const int r = (int)(g % R);
    const long long b = g / R;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= PADL && r < PADL + T) {
      const int id = __ldg(tokens + b * T + (r - PADL));
      if (id >= 0 && id < vocab) v = __ldg(reinterpret_cast<const float4*>(table + (size_t)id * E + c));
    }
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(X) + (size_t)g * E + c) = v;
    } else {
      __half2 lo = f16_sat2(v.x, v.y), hi = f16_sat2(v.z, v.w);
      uint2 u;
      u.x = *reinterpret_cast<unsigned*>(&lo);
      u.y = *reinterpret_cast<unsigned*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(X) + (size_t)g * E + c) = u;
    }
  }
}

__device__ __forceinline__ float enc_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// Recurrent half of Bidirectional(LSTM(u, return_sequences=True)) (Taco2.py:39-43; Keras LSTMCell: gate blocks i|f|c|o,
// sigmoid recurrent activation, zero initial state, no mask).  xs = x.W + b for both directions, [B][T][2][4u] fp32.
// CTA = (NB utterances, direction); thread j = hidden unit j: its 4 gate columns for NB utterances, c in registers,
// h double-buffered in shared memory.  U (u x 4u fp32) streams from L2 every step, shared by the CTA's NB utterances.
// Measured on B200 (256 x 150 tokens, 128 CTAs): 27 us per step = 4.1 ms, i.e. 128 x 1 MB / 27 us = 4.9 TB/s of L2 -> SM
// traffic for the U stream.  Tried, slower: 32 U loads issued ahead of their FMAs (+12 %), 8 utterances per CTA on 64 CTAs
// (+15 %).  The design that removes the stream is the decoder's: U resident in shared memory, split by hidden unit over the
// grid, h exchanged through L2 with one grid barrier per step.
template <int NB>
__global__ void __launch_bounds__(1024) encoder_bilstm_kernel(const float* __restrict__ xs, const float* __restrict__ Uf,
                                                              const float* __restrict__ Ub, float* __restrict__ out, int B, int T) {
  extern __shared__ __align__(16) float hbuf[];   // [2][NB][u]
  const int u = blockDim.x, j = threadIdx.x;
  const int dir = blockIdx.y, b0 = blockIdx.x * NB;
  const float* __restrict__ U = dir ? Ub : Uf;
  float c[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) {
    c[n] = 0.f;
    hbuf[n * u + j] = 0.f;
  }
  __syncthreads();
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const float* hc = hbuf + (s & 1) * NB * u;
    float* hn = hbuf + ((s + 1) & 1) * NB * u;
    float acc[4][NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      const bool ok = b0 + n < B;
      const float* x = xs + ((size_t)(ok ? b0 + n : b0) * T + t) * 8 * u + (size_t)dir * 4 * u + j;
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g][n] = ok ? __ldg(x + g * u) : 0.f;
    }
#pragma unroll 2
    for (int k = 0; k < u; k += 4) {
      float4 hv[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) hv[n] = *reinterpret_cast<const float4*>(hc + n * u + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float w[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) w[g] = __ldg(U + (size_t)(k + kk) * 4 * u + g * u + j);
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          const float hk = kk == 0 ? hv[n].x : kk == 1 ? hv[n].y : kk == 2 ? hv[n].z : hv[n].w;
#pragma unroll
          for (int g = 0; g < 4; ++g) acc[g][n] = fmaf(hk, w[g], acc[g][n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
      c[n] = enc_sigmoid(acc[1][n]) * c[n] + enc_sigmoid(acc[0][n]) * tanhf(acc[2][n]);
      const float hnew = enc_sigmoid(acc[3][n]) * tanhf(c[n]);
      hn[n * u + j] = hnew;
      if (b0 + n < B) out[((size_t)(b0 + n) * T + t) * 2 * u + (size_t)dir * u + j] = hnew;
    }
    __syncthreads();
  }
}

}  // namespace gstk
