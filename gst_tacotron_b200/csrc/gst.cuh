// GST front end (Modules/GST.py:12-124, Modules/Attention/Layers.py:147-285) and the generic
// fp32 GEMM used for the loop-invariant value projection (Modules/Attention/Steps.py:123).
#pragma once
#include "common.cuh"

namespace gstk {

// ---------------------------------------------------------------------------------------------
// C[m,n] = sum_k A[m*lda + k] * W[k*N + n] + bias[n] + group_bias[(m / rows_per_group)*N + n]
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles.  (Dense: y = x.kernel + bias.)
// ---------------------------------------------------------------------------------------------
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS) sgemm_bias_kernel(
    const float* __restrict__ A, long long lda, const float* __restrict__ W, const float* __restrict__ bias,
    const float* __restrict__ group_bias, int rows_per_group, float* __restrict__ C, int M, int N, int K) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    for (int i = tid; i < SG_BM * SG_BK; i += SG_THREADS) {
      const int mm = i / SG_BK, kk = i % SG_BK;
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? __ldg(A + (size_t)m * lda + k) : 0.f;
    }
    for (int i = tid; i < SG_BK * SG_BN; i += SG_THREADS) {
      const int kk = i / SG_BN, nn = i % SG_BN;
      const int k = k0 + kk, n = n0 + nn;
      Ws[kk][nn] = (k < K && n < N) ? __ldg(W + (size_t)k * N + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      if (group_bias) v += __ldg(group_bias + (size_t)(m / rows_per_group) * N + n);
      C[(size_t)m * N + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Conv2D(3x3, stride 2, 'same', no bias) + BatchNormalization(inference) + ReLU, NHWC
// (GST.py:23-31,55-56).  BN is folded on the host into scale/shift.  One CTA = (batch, 4 output
// rows); the input patch lives in shared memory; thread = (output channel, column group).
// TF 'same' padding for stride 2, k 3: pad_before = 0 for even input, 1 for odd.
// ---------------------------------------------------------------------------------------------
constexpr int CV_THREADS = 256, CV_HT = 4, CV_PW = 5;

__global__ void __launch_bounds__(CV_THREADS) conv3x3s2_bn_relu_kernel(
    const float* __restrict__ in, long long in_batch_stride, const float* __restrict__ w /*[3][3][Cin][Cout]*/,
    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int H, int W,
    int Cin, int Ho, int Wo, int Cout) {
  extern __shared__ __align__(16) float patch[];  // [2*CV_HT+1][W+2][Cin]
  const int b = blockIdx.y, ho0 = blockIdx.x * CV_HT;
  const int pad_h = (H & 1) ? 1 : 0, pad_w = (W & 1) ? 1 : 0;
  const int PR = 2 * CV_HT + 1, PWD = W + 2;
  const int hi0 = 2 * ho0 - pad_h;
  const float* inb = in + (size_t)b * in_batch_stride;
  for (int i = threadIdx.x; i < PR * PWD * Cin; i += CV_THREADS) {
    const int ci = i % Cin, rest = i / Cin;
    const int pw = rest % PWD, ph = rest / PWD;
    const int hi = hi0 + ph, wi = pw - pad_w;
    float v = 0.f;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(inb + ((size_t)hi * W + wi) * Cin + ci);
    patch[i] = v;
  }
  __syncthreads();
  const int WG = CV_THREADS / Cout;  // Cout in {32,64,128,256}
  const int co = threadIdx.x % Cout, wg = threadIdx.x / Cout;
  float acc[CV_HT][CV_PW];
#pragma unroll
  for (int hh = 0; hh < CV_HT; ++hh)
#pragma unroll
    for (int pp = 0; pp < CV_PW; ++pp) acc[hh][pp] = 0.f;
  for (int wbase = 0; wbase < Wo; wbase += WG * CV_PW) {
#pragma unroll
    for (int hh = 0; hh < CV_HT; ++hh)
#pragma unroll
      for (int pp = 0; pp < CV_PW; ++pp) acc[hh][pp] = 0.f;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        for (int ci = 0; ci < Cin; ++ci) {
          const float wv = __ldg(w + ((size_t)(kh * 3 + kw) * Cin + ci) * Cout + co);
#pragma unroll
          for (int hh = 0; hh < CV_HT; ++hh)
#pragma unroll
            for (int pp = 0; pp < CV_PW; ++pp) {
              const int wo = wbase + wg + pp * WG;
              if (wo < Wo)
                acc[hh][pp] = fmaf(patch[((size_t)(2 * hh + kh) * PWD + (2 * wo + kw)) * Cin + ci], wv, acc[hh][pp]);
            }
        }
    const float sc = __ldg(scale + co), sh = __ldg(shift + co);
#pragma unroll
    for (int hh = 0; hh < CV_HT; ++hh) {
      const int ho = ho0 + hh;
      if (ho >= Ho) continue;
#pragma unroll
      for (int pp = 0; pp < CV_PW; ++pp) {
        const int wo = wbase + wg + pp * WG;
        if (wo < Wo)
          out[(((size_t)b * Ho + ho) * Wo + wo) * Cout + co] = fmaxf(fmaf(acc[hh][pp], sc, sh), 0.f);
      }
    }
  }
}

// Register-tiled variant for Cin % 4 == 0, Cout % 4 == 0 (every layer but the first): a thread owns 4 consecutive output
// channels x CV2_PP output columns of one output row, so one 16 B weight load and one 16 B patch load (4 input channels)
// feed 16 FMAs each - the first version issued one shared-memory load per FMA and ran at 5 % of the fp32 FMA peak.
// Thread tiles are enumerated (channel group fastest) so that a warp reads one patch address (broadcast) and 32
// consecutive float4 of weights.
constexpr int CV2_PP = 2;
__global__ void __launch_bounds__(CV_THREADS) conv3x3s2_bn_relu_v2_kernel(
    const float* __restrict__ in, long long in_batch_stride, const float* __restrict__ w /*[3][3][Cin][Cout]*/,
    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int H, int W,
    int Cin, int Ho, int Wo, int Cout) {
  extern __shared__ __align__(16) float patch[];  // [2*CV_HT+1][W+2][Cin]
  const int b = blockIdx.y, ho0 = blockIdx.x * CV_HT;
  const int pad_h = (H & 1) ? 1 : 0, pad_w = (W & 1) ? 1 : 0;
  const int PR = 2 * CV_HT + 1, PWD = W + 2;
  const int hi0 = 2 * ho0 - pad_h;
  const float* inb = in + (size_t)b * in_batch_stride;
  const int cin4 = Cin >> 2;
  for (int i = threadIdx.x; i < PR * PWD * cin4; i += CV_THREADS) {
    const int c4 = i % cin4, rest = i / cin4;
    const int pw = rest % PWD, ph = rest / PWD;
    const int hi = hi0 + ph, wi = pw - pad_w;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)hi * W + wi) * Cin) + c4);
    reinterpret_cast<float4*>(patch)[i] = v;
  }
  __syncthreads();
  const int CG = Cout >> 2, NWG = (Wo + CV2_PP - 1) / CV2_PP;
  for (int tile = threadIdx.x; tile < CG * CV_HT * NWG; tile += CV_THREADS) {
    const int cg = tile % CG, r = tile / CG, hh = r % CV_HT, wo0 = (r / CV_HT) * CV2_PP;
    if (ho0 + hh >= Ho) continue;
    float acc[CV2_PP][4];
#pragma unroll
    for (int pp = 0; pp < CV2_PP; ++pp)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[pp][k] = 0.f;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wp = reinterpret_cast<const float4*>(w + (size_t)(kh * 3 + kw) * Cin * Cout) + cg;   // + ci * CG
        const float4* xp[CV2_PP];
#pragma unroll
        for (int pp = 0; pp < CV2_PP; ++pp) {
          const int wo = min(wo0 + pp, Wo - 1);   // clamped: the surplus column is computed and dropped
          xp[pp] = reinterpret_cast<const float4*>(patch + ((size_t)(2 * hh + kh) * PWD + (2 * wo + kw)) * Cin);
        }
#pragma unroll 2
        for (int c4 = 0; c4 < cin4; ++c4) {
          const float4 w0 = __ldg(wp + (size_t)(4 * c4) * CG), w1 = __ldg(wp + (size_t)(4 * c4 + 1) * CG);
          const float4 w2 = __ldg(wp + (size_t)(4 * c4 + 2) * CG), w3 = __ldg(wp + (size_t)(4 * c4 + 3) * CG);
#pragma unroll
          for (int pp = 0; pp < CV2_PP; ++pp) {
            const float4 x = xp[pp][c4];
            acc[pp][0] = fmaf(x.x, w0.x, acc[pp][0]); acc[pp][1] = fmaf(x.x, w0.y, acc[pp][1]);
            acc[pp][2] = fmaf(x.x, w0.z, acc[pp][2]); acc[pp][3] = fmaf(x.x, w0.w, acc[pp][3]);
            acc[pp][0] = fmaf(x.y, w1.x, acc[pp][0]); acc[pp][1] = fmaf(x.y, w1.y, acc[pp][1]);
            acc[pp][2] = fmaf(x.y, w1.z, acc[pp][2]); acc[pp][3] = fmaf(x.y, w1.w, acc[pp][3]);
            acc[pp][0] = fmaf(x.z, w2.x, acc[pp][0]); acc[pp][1] = fmaf(x.z, w2.y, acc[pp][1]);
            acc[pp][2] = fmaf(x.z, w2.z, acc[pp][2]); acc[pp][3] = fmaf(x.z, w2.w, acc[pp][3]);
            acc[pp][0] = fmaf(x.w, w3.x, acc[pp][0]); acc[pp][1] = fmaf(x.w, w3.y, acc[pp][1]);
            acc[pp][2] = fmaf(x.w, w3.z, acc[pp][2]); acc[pp][3] = fmaf(x.w, w3.w, acc[pp][3]);
          }
        }
      }
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cg), sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
#pragma unroll
    for (int pp = 0; pp < CV2_PP; ++pp) {
      const int wo = wo0 + pp;
      if (wo < Wo)
        reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho0 + hh) * Wo + wo) * Cout)[cg] =
            make_float4(fmaxf(fmaf(acc[pp][0], sc.x, sh.x), 0.f), fmaxf(fmaf(acc[pp][1], sc.y, sh.y), 0.f),
                        fmaxf(fmaf(acc[pp][2], sc.z, sh.z), 0.f), fmaxf(fmaf(acc[pp][3], sc.w, sh.w), 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GRU recurrence (Keras GRU, reset_after=True, gate order z,r,h) over the pre-computed
// input projections xs = x.W + b[0], stopping at the only step the reference keeps
// (gather_nd at ceil(len/compress)-1, GST.py:65-68), then Dense(tanh) (GST.py:70) and - fused -
// the style-token multi-head attention with residual + Layer_Norm (GST.py:100-109,
// Layers.py:172-214, 280-285).  One CTA per reference mel; the recurrent kernel sits in smem.
// ---------------------------------------------------------------------------------------------
struct GruMhaParams {
  const float* xs;     // [B, Tp, 3G]
  const float* U;      // [G, 3G]
  const float* b_rec;  // [3G]  (bias[1])
  const float* Wd;     // [G, D]
  const float* bd;     // [D]
  const float* Wq;     // [D, S]
  const float* bq;     // [S]
  const float* tokkv;  // [NT, S] = tanh(tokens).Wv + bv  (batch invariant)
  const float* ln_g;
  const float* ln_b;
  const int* lengths;  // [B]
  float* out_gst;      // [B,S] or null
  float* out_ref;      // [B,D] or null
  float* out_att;      // [B,NT] or null
  int B, Tp, G, D, S, NT, heads, compress;
};

constexpr int GRU_UPC = 4;   // utterances a CTA runs through the recurrence together (the recurrent kernel is read once for all of them)
__global__ void __launch_bounds__(384) gru_dense_mha_kernel(const GruMhaParams p, int upc) {
  extern __shared__ __align__(16) float sm[];
  const int G = p.G, G3 = 3 * p.G;
  float* U_s = sm;                 // [G][3G]
  float* h4_s = U_s + (size_t)G * G3;    // [G][GRU_UPC]  hidden states, utterance fastest (one 16 B broadcast load per k)
  float* hh_s = h4_s + GRU_UPC * G;      // [GRU_UPC][3G]
  float* h_s = hh_s + GRU_UPC * G3;      // [G]    (tail: the utterance being finished)
  float* ref_s = h_s + G;          // [D]
  float* q_s = ref_s + p.D;        // [S]
  float* y_s = q_s + p.S;          // [S]
  float* pr_s = y_s + p.S;         // [heads][NT]
  float* sc_s = pr_s + p.heads * p.NT;  // [4]
  __shared__ int nsteps_s[GRU_UPC];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b0 = blockIdx.x * upc, nu = min(upc, p.B - b0);
  for (int i = tid; i < G * G3; i += nthr) U_s[i] = __ldg(p.U + i);
  for (int i = tid; i < GRU_UPC * G; i += nthr) h4_s[i] = 0.f;
  if (tid < GRU_UPC) {
    int ns = 0;
    if (tid < nu) {
      const int len = p.lengths[b0 + tid];
      ns = (len + p.compress - 1) / p.compress;
      ns = max(1, min(ns, p.Tp));
    }
    nsteps_s[tid] = ns;
  }
  __syncthreads();
  int nmax = 0;
  for (int j = 0; j < GRU_UPC; ++j) nmax = max(nmax, nsteps_s[j]);
  for (int t = 0; t < nmax; ++t) {
    for (int n = tid; n < G3; n += nthr) {
      float a0[GRU_UPC] = {}, a1[GRU_UPC] = {};
      for (int k = 0; k < G; k += 2) {
        const float u0 = U_s[(size_t)k * G3 + n], u1 = U_s[(size_t)(k + 1) * G3 + n];
        const float4 h0 = *reinterpret_cast<const float4*>(h4_s + k * GRU_UPC), h1 = *reinterpret_cast<const float4*>(h4_s + (k + 1) * GRU_UPC);
        a0[0] = fmaf(h0.x, u0, a0[0]); a0[1] = fmaf(h0.y, u0, a0[1]); a0[2] = fmaf(h0.z, u0, a0[2]); a0[3] = fmaf(h0.w, u0, a0[3]);
        a1[0] = fmaf(h1.x, u1, a1[0]); a1[1] = fmaf(h1.y, u1, a1[1]); a1[2] = fmaf(h1.z, u1, a1[2]); a1[3] = fmaf(h1.w, u1, a1[3]);
      }
      const float br = __ldg(p.b_rec + n);
#pragma unroll
      for (int j = 0; j < GRU_UPC; ++j) hh_s[j * G3 + n] = a0[j] + a1[j] + br;
    }
    __syncthreads();
    for (int i = tid; i < nu * G; i += nthr) {
      const int j = i / G, n = i - j * G;
      if (t < nsteps_s[j]) {   // the reference keeps the state at step ceil(len / compress) - 1 (gather_nd, GST.py:65-68)
        const float* x = p.xs + ((size_t)(b0 + j) * p.Tp + t) * G3;
        const float* hh = hh_s + j * G3;
        const float hp = h4_s[n * GRU_UPC + j];
        const float z = sigmoid_acc(__ldg(x + n) + hh[n]);
        const float r = sigmoid_acc(__ldg(x + G + n) + hh[G + n]);
        const float c = tanhf(__ldg(x + 2 * G + n) + r * hh[2 * G + n]);
        h4_s[n * GRU_UPC + j] = z * hp + (1.0f - z) * c;
      }
    }
    __syncthreads();
  }
  for (int j = 0; j < nu; ++j) {
  const int b = b0 + j;
  __syncthreads();   // the previous utterance's tail is done with h_s / ref_s / q_s / y_s / pr_s
  for (int n = tid; n < G; n += nthr) h_s[n] = h4_s[n * GRU_UPC + j];
  __syncthreads();
  // Dense(tanh)
  for (int n = tid; n < p.D; n += nthr) {
    float a = __ldg(p.bd + n);
    for (int k = 0; k < G; ++k) a = fmaf(h_s[k], __ldg(p.Wd + (size_t)k * p.D + n), a);
    a = tanhf(a);
    ref_s[n] = a;
    if (p.out_ref) p.out_ref[(size_t)b * p.D + n] = a;
  }
  __syncthreads();
  if (!p.out_gst && !p.out_att) continue;
  // query projection
  for (int n = tid; n < p.S; n += nthr) {
    float a = __ldg(p.bq + n);
    for (int k = 0; k < p.D; ++k) a = fmaf(ref_s[k], __ldg(p.Wq + (size_t)k * p.S + n), a);
    q_s[n] = a;
  }
  __syncthreads();
  const int hd = p.S / p.heads;
  // unscaled dot-product scores, one thread per (head, token)
  for (int i = tid; i < p.heads * p.NT; i += nthr) {
    const int h = i / p.NT, tok = i % p.NT;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(q_s[h * hd + d], __ldg(p.tokkv + (size_t)tok * p.S + h * hd + d), a);
    pr_s[i] = a;
  }
  __syncthreads();
  if (tid < p.heads) {  // softmax over tokens (max-subtracted)
    float m = -INFINITY;
    for (int tok = 0; tok < p.NT; ++tok) m = fmaxf(m, pr_s[tid * p.NT + tok]);
    float s = 0.f;
    for (int tok = 0; tok < p.NT; ++tok) {
      const float e = expf(pr_s[tid * p.NT + tok] - m);
      pr_s[tid * p.NT + tok] = e;
      s += e;
    }
    for (int tok = 0; tok < p.NT; ++tok) pr_s[tid * p.NT + tok] /= s;
  }
  __syncthreads();
  for (int n = tid; n < p.S; n += nthr) {
    const int h = n / hd;
    float a = 0.f;
    for (int tok = 0; tok < p.NT; ++tok) a = fmaf(pr_s[h * p.NT + tok], __ldg(p.tokkv + (size_t)tok * p.S + n), a);
    y_s[n] = a + q_s[n];  // residual with the projected query (Layers.py:211)
  }
  if (p.out_att)
    for (int tok = tid; tok < p.NT; tok += nthr) {
      float a = 0.f;
      for (int h = 0; h < p.heads; ++h) a += pr_s[h * p.NT + tok];
      p.out_att[(size_t)b * p.NT + tok] = a / (float)p.heads;
    }
  __syncthreads();
  if (tid < 32) {  // Layer_Norm statistics (biased variance, eps inside sqrt)
    float s = 0.f;
    for (int n = tid; n < p.S; n += 32) s += y_s[n];
    const float mean = warp_sum(s) / (float)p.S;
    float v = 0.f;
    for (int n = tid; n < p.S; n += 32) {
      const float d = y_s[n] - mean;
      v = fmaf(d, d, v);
    }
    const float var = warp_sum(v) / (float)p.S;
    if (tid == 0) {
      sc_s[0] = mean;
      sc_s[1] = 1.0f / sqrtf(var + 1e-8f);
    }
  }
  __syncthreads();
  if (p.out_gst)
    for (int n = tid; n < p.S; n += nthr)
      p.out_gst[(size_t)b * p.S + n] = __ldg(p.ln_g + n) * ((y_s[n] - sc_s[0]) * sc_s[1]) + __ldg(p.ln_b + n);
  }   // utterances of this CTA
}

inline size_t gru_mha_smem_bytes(int G, int D, int S, int NT, int heads) {
  return sizeof(float) * ((size_t)G * 3 * G + GRU_UPC * (G + 3 * G) + G + D + 2 * S + heads * NT + 4);
}

// ---------------------------------------------------------------------------------------------
// Generic MultiHeadAttention.call on [query, value] (Layers.py:172-214): one CTA per (b, tq).
// ---------------------------------------------------------------------------------------------
struct MhaParams {
  const float *query, *value, *Wq, *bq, *Wv, *bv, *ln_g, *ln_b;
  float *out, *out_att;
  int B, tq, tv, dq, dv, S, heads;
};

__global__ void __launch_bounds__(256) mha_generic_kernel(const MhaParams p) {
  extern __shared__ __align__(16) float sm[];
  float* v_s = sm;                      // [tv][S]
  float* q_s = v_s + (size_t)p.tv * p.S;  // [S]
  float* y_s = q_s + p.S;               // [S]
  float* pr_s = y_s + p.S;              // [heads][tv]
  float* sc_s = pr_s + p.heads * p.tv;  // [4]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.x / p.tq, iq = blockIdx.x % p.tq;
  const float* qin = p.query + ((size_t)b * p.tq + iq) * p.dq;
  for (int n = tid; n < p.S; n += nthr) {
    float a = __ldg(p.bq + n);
    for (int k = 0; k < p.dq; ++k) a = fmaf(__ldg(qin + k), __ldg(p.Wq + (size_t)k * p.S + n), a);
    q_s[n] = a;
  }
  for (int i = tid; i < p.tv * p.S; i += nthr) {
    const int j = i / p.S, n = i % p.S;
    const float* vin = p.value + ((size_t)b * p.tv + j) * p.dv;
    float a = __ldg(p.bv + n);
    for (int k = 0; k < p.dv; ++k) a = fmaf(__ldg(vin + k), __ldg(p.Wv + (size_t)k * p.S + n), a);
    v_s[i] = a;
  }
  __syncthreads();
  const int hd = p.S / p.heads;
  for (int i = tid; i < p.heads * p.tv; i += nthr) {
    const int h = i / p.tv, j = i % p.tv;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(q_s[h * hd + d], v_s[(size_t)j * p.S + h * hd + d], a);
    pr_s[i] = a;
  }
  __syncthreads();
  if (tid < p.heads) {
    float m = -INFINITY;
    for (int j = 0; j < p.tv; ++j) m = fmaxf(m, pr_s[tid * p.tv + j]);
    float s = 0.f;
    for (int j = 0; j < p.tv; ++j) {
      const float e = expf(pr_s[tid * p.tv + j] - m);
      pr_s[tid * p.tv + j] = e;
      s += e;
    }
    for (int j = 0; j < p.tv; ++j) pr_s[tid * p.tv + j] /= s;
  }
  __syncthreads();
  for (int n = tid; n < p.S; n += nthr) {
    const int h = n / hd;
    float a = 0.f;
    for (int j = 0; j < p.tv; ++j) a = fmaf(pr_s[h * p.tv + j], v_s[(size_t)j * p.S + n], a);
    y_s[n] = a + q_s[n];
  }
  if (p.out_att)
    for (int j = tid; j < p.tv; j += nthr) {
      float a = 0.f;
      for (int h = 0; h < p.heads; ++h) a += pr_s[h * p.tv + j];
      p.out_att[((size_t)b * p.tq + iq) * p.tv + j] = a / (float)p.heads;
    }
  __syncthreads();
  if (tid < 32) {
    float s = 0.f;
    for (int n = tid; n < p.S; n += 32) s += y_s[n];
    const float mean = warp_sum(s) / (float)p.S;
    float v = 0.f;
    for (int n = tid; n < p.S; n += 32) {
      const float d = y_s[n] - mean;
      v = fmaf(d, d, v);
    }
    const float var = warp_sum(v) / (float)p.S;
    if (tid == 0) {
      sc_s[0] = mean;
      sc_s[1] = 1.0f / sqrtf(var + 1e-8f);
    }
  }
  __syncthreads();
  for (int n = tid; n < p.S; n += nthr)
    p.out[((size_t)b * p.tq + iq) * p.S + n] =
        __ldg(p.ln_g + n) * ((y_s[n] - sc_s[0]) * sc_s[1]) + __ldg(p.ln_b + n);
}

// ---------------------------------------------------------------------------------------------
// Stand-alone step attention (BahdanauMonotonicAttention / StepwiseMonotonicAttention .call,
// Modules/Attention/Steps.py:107-229) on already projected query [B,A], key [B,Tv,A], value [B,Tv,A].
// One CTA (256 threads) per batch row.  type 0 = SMA, 1 = BMA.
// ---------------------------------------------------------------------------------------------
struct AttStepParams {
  const float *q, *key, *value, *prev, *att_v, *noise;
  float score_bias, sigmoid_noise;
  float *ctx, *align;
  int B, Tv, A, type;
};

__global__ void __launch_bounds__(256) attention_step_kernel(const AttStepParams p) {
  extern __shared__ __align__(16) float sm[];
  float* e_s = sm;               // [Tv]
  float* al_s = e_s + p.Tv;      // [Tv]
  float* red = al_s + p.Tv;      // [256]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* q = p.q + (size_t)b * p.A;
  const float* K = p.key + (size_t)b * p.Tv * p.A;
  const float* V = p.value + (size_t)b * p.Tv * p.A;
  const float* prev = p.prev + (size_t)b * p.Tv;
  for (int j = wid; j < p.Tv; j += 8) {
    float acc = 0.f;
    for (int a = lane; a < p.A; a += 32) acc = fmaf(__ldg(p.att_v + a), tanhf(__ldg(q + a) + __ldg(K + (size_t)j * p.A + a)), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += p.score_bias;
      if (p.sigmoid_noise > 0.f && p.noise) acc = fmaf(p.sigmoid_noise, __ldg(p.noise + (size_t)b * p.Tv + j), acc);
      e_s[j] = acc;
    }
  }
  __syncthreads();
  if (p.type == 0) {
    for (int j = tid; j < p.Tv; j += 256) {
      float v = prev[j] * sigmoid_acc(e_s[j]);
      if (j > 0) v += prev[j - 1] * (1.0f - sigmoid_acc(e_s[j - 1]));
      al_s[j] = v;
    }
  } else if (wid == 0) {
    float carry_log = 0.f, carry_sum = 0.f;
    for (int j0 = 0; j0 < p.Tv; j0 += 32) {
      const int j = j0 + lane;
      const bool ok = j < p.Tv;
      const float pj = ok ? sigmoid_acc(e_s[j]) : 0.f;
      const float lg = ok ? logf(fminf(fmaxf(1.0f - pj, 1.17549435e-38f), 1.0f)) : 0.f;
      float inc = lg;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      const float cp = expf(carry_log + inc - lg);
      const float term = ok ? prev[j] / fminf(fmaxf(cp, 1e-10f), 1.0f) : 0.f;
      float cs = term;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, cs, o);
        if (lane >= o) cs += n;
      }
      if (ok) al_s[j] = pj * cp * (carry_sum + cs);
      carry_log += __shfl_sync(0xffffffffu, inc, 31);
      carry_sum += __shfl_sync(0xffffffffu, cs, 31);
    }
  }
  __syncthreads();
  for (int j = tid; j < p.Tv; j += 256) p.align[(size_t)b * p.Tv + j] = al_s[j];
  for (int a = tid; a < p.A; a += 256) {
    float c = 0.f;
    for (int j = 0; j < p.Tv; ++j) c = fmaf(al_s[j], __ldg(V + (size_t)j * p.A + a), c);
    p.ctx[(size_t)b * p.A + a] = c;
  }
  (void)red;
}

// GST_Concated_Encoder.call (GST.py:121-124): out[b,t,:] = [gst[b] || enc[b,t]]
__global__ void concat_encoder_kernel(const float* __restrict__ enc, const float* __restrict__ gst,
                                      float* __restrict__ out, int B, int Tv, int Dt, int Dg) {
  const size_t total = (size_t)B * Tv * (Dt + Dg);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (Dt + Dg));
    const size_t row = i / (Dt + Dg);
    const int b = (int)(row / Tv);
    out[i] = c < Dg ? __ldg(gst + (size_t)b * Dg + c) : __ldg(enc + row * Dt + (c - Dg));
  }
}

__global__ void init_alignment_kernel(float* __restrict__ align, int B, int Tv, int one_hot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * Tv) align[i] = (one_hot && (i % Tv) == 0) ? 1.f : 0.f;
}

}  // namespace gstk
