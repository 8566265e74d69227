// GST reference encoder, convolution stack on the tensor cores (Modules/GST.py:17-31,54-56; SURVEY 8 row NS-b).
//
// Conv2D(3x3, strides 2, 'same') + BatchNormalization + ReLU, six times.  A 3x3 stride-2 window over pixels is a 2x2 stride-1
// window over 2x2 pixel BLOCKS: with the input stored space-to-depth - one row of a flat matrix per block, 4 C channels
// [(sub-row, sub-column, c)] - output pixel (ho, wo) reads the blocks (ho, wo), (ho, wo+1), (ho+1, wo), (ho+1, wo+1), i.e. the
// matrix rows g, g + 1, g + Wb, g + Wb + 1 (Wb = Wo + 1 blocks per block row, the last block column / block row being zero
// padding), against a [16 C] x [N] kernel whose 7 never-touched (sub-)taps are zero.  That is exactly the implicit GEMM of
// postnet_tc.cuh (TMA boxes of 128 rows x 64 channels per tap, tcgen05.mma M=128 x N=Cout, fp16 operands, fp32 accumulators in
// TMEM) with a table of tap row offsets; the epilogue (+ shift, ReLU) scatters every pixel straight into the NEXT layer's
// block matrix.  TF 'same' padding of a stride-2 layer puts one padding row in front of an ODD dimension and none in front of an
// even one: the producer shifts the pixels of an odd dimension by one, so the window always starts on a block boundary.
//
// Layer 0 (one input channel, K = 9) stays a direct convolution: gst_conv0_s2d_kernel reads the fp32 mel and writes fp16 blocks.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "gst.cuh"

namespace gstk {

struct GstGeom {   // one dimension of one layer
  int in, out, pad, shift;   // input size, ceil(in / 2), TF pad_before = ((out-1)*2 + 3 - in) / 2, 1 if `in` is odd
};
__host__ __device__ inline GstGeom gst_geom(int in) {
  GstGeom g;
  g.in = in; g.out = (in + 1) / 2;
  const int total = (g.out - 1) * 2 + 3 - in;
  g.pad = total > 0 ? total / 2 : 0;
  g.shift = in & 1;
  return g;
}

// Layer 0: mel [B][H][W] fp32 (one channel) -> relu(conv * scale + shift) -> fp16 pixels in the block matrix of layer 1.
// K = 9 taps is too thin for tcgen05, and as scalar FMAs the layer is instruction-bound (ncu: 247 M warp instructions for
// 512 x 1000 frames), so it runs on the warp-level tensor cores: mma.sync m16n8k16, A = 16 consecutive output pixels of one row x
// (9 taps, zero-padded to 16), gathered from the fp16 input rows staged in shared memory; B = [16][8 channels] fragments held in
// registers.  The channel columns are PERMUTED when the weights are packed (w0p, shp: column 2t+e of n-tile n of a 32-channel
// group = channel 8t + 2n + e), so the D fragments of the four n-tiles give every lane 8 CONTIGUOUS channels of its two pixels:
// one 16 B store each, no shuffles.  block = (image, strip of G0_HO output rows), warp = (row, 16-pixel tile) tasks.
constexpr int G0_HO = 16;
__device__ __forceinline__ void mma_16816_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// w0p: [co / 32 groups][4 n-tiles][32 lanes][2] packed B fragments (uint32 = two fp16), shp: [co / 32][32 lanes... see host] fp32
template <int NG>   // NG = co / 32 channel groups
__global__ void __launch_bounds__(256) gst_conv0_mma_kernel(const float* __restrict__ x, long long x_bs, const uint32_t* __restrict__ w0p,
                                                            const float* __restrict__ shp, __half* __restrict__ y, int B, int H, int W, int Ho, int Wo,
                                                            int pt, int pl, int nHb, int nWb, int nsh, int nsw) {
  extern __shared__ __half in_h[];   // [2 G0_HO + 1][Wp]: input rows 2 ho0 - pt .., columns -pl .. (zero outside the image)
  constexpr int co = 32 * NG;
  const int mtiles = (Wo + 15) / 16, Wp = 2 * 16 * mtiles + 2;
  const int strips = (Ho + G0_HO - 1) / G0_HO;
  const int b = blockIdx.x / strips, strip = blockIdx.x - b * strips;
  const int ho0 = strip * G0_HO, nrow = 2 * G0_HO + 1;
  const float* xb = x + (size_t)b * x_bs;
  for (int cc = threadIdx.x; cc < Wp; cc += blockDim.x) {   // thread = staged column: walks down the rows with pointer bumps only
    const int ww = cc - pl;
    const bool col_ok = ww >= 0 && ww < W;
    int hh = 2 * ho0 - pt;
    const float* src = xb + (long long)hh * W + ww;
    __half* dst = in_h + cc;
#pragma unroll 4
    for (int rr = 0; rr < nrow; ++rr, ++hh, src += W, dst += Wp) *dst = __float2half_rn((col_ok && hh >= 0 && hh < H) ? __ldg(src) : 0.f);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  uint32_t bw[NG][4][2];
  float sh[NG][8];
#pragma unroll
  for (int q = 0; q < NG; ++q) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(w0p) + (q * 4 + n) * 32 + lane);
      bw[q][n][0] = v.x; bw[q][n][1] = v.y;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) sh[q][c] = __ldg(shp + q * 32 + 8 * t + c);
  }
  __syncthreads();
  // taps 2t, 2t+1 of this lane's A columns (k = kh*3 + kw), tap 8 for t == 0
  const int k0 = 2 * t, k1 = 2 * t + 1;
  const int off0 = (k0 / 3) * Wp + k0 % 3, off1 = (k1 / 3) * Wp + k1 % 3, off8 = 2 * Wp + 2;
  const unsigned short* in_u = reinterpret_cast<const unsigned short*>(in_h);
  for (int task = wid; task < G0_HO * mtiles; task += nwarp) {
    const int j = task / mtiles, mt = task - j * mtiles, ho = ho0 + j;
    if (ho >= Ho) break;
    const int wo_a = mt * 16 + g, wo_b = wo_a + 8;
    const unsigned short* pa = in_u + (2 * j) * Wp + 2 * wo_a;
    const unsigned short* pb = pa + 16;
    const uint32_t a0 = (uint32_t)pa[off0] | ((uint32_t)pa[off1] << 16), a1 = (uint32_t)pb[off0] | ((uint32_t)pb[off1] << 16);
    const uint32_t a2 = t == 0 ? (uint32_t)pa[off8] : 0u, a3 = t == 0 ? (uint32_t)pb[off8] : 0u;
    const int hs = ho + nsh;
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      float d[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        d[n][0] = sh[q][2 * n]; d[n][1] = sh[q][2 * n + 1]; d[n][2] = sh[q][2 * n]; d[n][3] = sh[q][2 * n + 1];
        mma_16816_f16(d[n], a0, a1, a2, a3, bw[q][n][0], bw[q][n][1]);
      }
      // lane: pixel wo_a -> channels 32 q + 8 t + {0..7} = d[n][0], d[n][1] for n = 0..3; pixel wo_b -> d[n][2], d[n][3]
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int wo = half ? wo_b : wo_a;
        if (wo < Wo) {
          __half2 o[4];
#pragma unroll
          for (int n = 0; n < 4; ++n) o[n] = __floats2half2_rn(fmaxf(d[n][2 * half], 0.f), fmaxf(d[n][2 * half + 1], 0.f));
          const int ws = wo + nsw;
          __half* dst = y + ((size_t)((size_t)b * nHb + (hs >> 1)) * nWb + (ws >> 1)) * 4 * co + (size_t)(((hs & 1) * 2 + (ws & 1)) * co) + 32 * q + 8 * t;
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
        }
      }
    }
  }
}

// Zero padding of a block matrix [B][Hb][Wb][4 C] (fp16): the last block row and block column, and - where a dimension is shifted
// (odd input size) - the sub-row / sub-column 0 of the first block row / column.  The producers only write valid pixels.
__global__ void gst_border_zero_kernel(__half* __restrict__ y, int B, int Hb, int Wb, int C, int sh, int sw) {
  const int c8 = C / 8;                               // 16 B chunks per sub-pixel; a block = 4 sub-pixels = 4 c8 chunks
  const int nA = Wb * 4 * c8;                         // last block row
  const int nB = Hb * 4 * c8;                         // last block column
  const int nC = sh ? Wb * 2 * c8 : 0;                // sub-row 0 of the first block row
  const int nD = sw ? Hb * 2 * c8 : 0;                // sub-column 0 of the first block column
  const int per_img = nA + nB + nC + nD;
  const long long total = (long long)B * per_img;
  uint4* out = reinterpret_cast<uint4*>(y);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_img);
    int j = (int)(i - (long long)b * per_img);
    int bh, bw, chunk;
    if (j < nA) { bh = Hb - 1; bw = j / (4 * c8); chunk = j % (4 * c8); }
    else if ((j -= nA) < nB) { bw = Wb - 1; bh = j / (4 * c8); chunk = j % (4 * c8); }
    else if ((j -= nB) < nC) { bh = 0; bw = j / (2 * c8); chunk = j % (2 * c8); }                            // sub-pixels (0, 0), (0, 1)
    else { j -= nC; bw = 0; bh = j / (2 * c8); const int k = j % (2 * c8); chunk = (k / c8) * 2 * c8 + k % c8; }   // sub-pixels (0, 0), (1, 0)
    out[((size_t)((size_t)b * Hb + bh) * Wb + bw) * 4 * c8 + chunk] = make_uint4(0u, 0u, 0u, 0u);
  }
}


// ---------------------------------------------------------------------------------------------
// GRU recurrence + Dense(tanh) + style-token attention (GST.py:58-70,100-109) in the tensor-core mode, G = 128 units.
// Same math as gru_dense_mha_kernel (gst.cuh); the recurrent product hh = h . U runs on mma.sync m16n8k16 (fp16 operands, fp32
// accumulate): warp w owns the units 16 w .. 16 w + 15 of all three gates (z, r, h), its 3 x 8 A fragments of U^T (16 outputs x
// 16 k each) stay in REGISTERS for the whole call (96 registers), the CTA's <= 8 utterances are the N columns, h(t) travels as
// fp16 through a double-buffered shared-memory tile (one block barrier per step) while the fp32 state stays in the registers of
// the lane that owns (unit, utterance) in the D fragments.  Up: image [8 warps][3 gates][8 k-tiles][32 lanes][16 B] (host-packed).
// The tail (Dense, query projection, 4-head token attention, residual, Layer_Norm) is batched over the CTA's utterances.
// ---------------------------------------------------------------------------------------------
constexpr int GM_G = 128, GM_NU = 8, GM_HS = GM_G + 8;   // units, utterance slots of a CTA, fp16 row stride of the h tile
__global__ void __launch_bounds__(256) gru_mma_dense_mha_kernel(const GruMhaParams p, const uint4* __restrict__ Up, int upc) {
  extern __shared__ __align__(16) float gsm[];
  const int D = p.D, S = p.S, NT = p.NT, heads = p.heads;
  __half* hb = reinterpret_cast<__half*>(gsm);              // [2][GM_NU][GM_HS] fp16 h tiles (B operand)
  float* hf = gsm + (2 * GM_NU * GM_HS) / 2;                // [GM_NU][G]  final states (fp32)
  float* ref_s = hf + GM_NU * GM_G;                         // [GM_NU][D]
  float* q_s = ref_s + GM_NU * D;                           // [GM_NU][S]
  float* y_s = q_s + GM_NU * S;                             // [GM_NU][S]
  float* pr_s = y_s + GM_NU * S;                            // [GM_NU][heads][NT]
  float* sc_s = pr_s + GM_NU * heads * NT;                  // [GM_NU][2]
  __shared__ int nsteps_s[GM_NU];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int b0 = blockIdx.x * upc, nu = min(upc, p.B - b0);
  uint4 a[3][8];
#pragma unroll
  for (int gate = 0; gate < 3; ++gate)
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) a[gate][kt] = __ldg(Up + ((size_t)(wid * 3 + gate) * 8 + kt) * 32 + lane);
  for (int i = tid; i < 2 * GM_NU * GM_HS / 2; i += nthr) reinterpret_cast<uint32_t*>(hb)[i] = 0u;
  if (tid < GM_NU) {
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
#pragma unroll
  for (int j = 0; j < GM_NU; ++j) nmax = max(nmax, nsteps_s[j]);
  // D fragment of this lane: units n0 = 16 wid + g and n0 + 8, utterances u0 = 2 t4 and u0 + 1
  const int n0 = 16 * wid + g, u0 = 2 * t4;
  float h[4] = {0.f, 0.f, 0.f, 0.f};   // (n0,u0) (n0,u0+1) (n0+8,u0) (n0+8,u0+1)
  float br[3][2];
#pragma unroll
  for (int gate = 0; gate < 3; ++gate) { br[gate][0] = __ldg(p.b_rec + gate * GM_G + n0); br[gate][1] = __ldg(p.b_rec + gate * GM_G + n0 + 8); }
  int ns_u[2] = {nsteps_s[u0], nsteps_s[u0 + 1]};
  for (int t = 0; t < nmax; ++t) {
    // input projections of this step for the lane's (unit, utterance) pairs: issued first, consumed after the mma chain
    float x[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + 8 * (e >> 1), u = u0 + (e & 1);
      const bool live = u < nu && t < nsteps_s[u];
      const float* xr = p.xs + ((size_t)(b0 + (live ? u : 0)) * p.Tp + (live ? t : 0)) * 3 * GM_G + n;
#pragma unroll
      for (int gate = 0; gate < 3; ++gate) x[gate][e] = live ? __ldg(xr + gate * GM_G) : 0.f;
    }
    const __half* hcur = hb + (size_t)(t & 1) * GM_NU * GM_HS + g * GM_HS + 2 * t4;
    float d[3][4];
#pragma unroll
    for (int gate = 0; gate < 3; ++gate) { d[gate][0] = br[gate][0]; d[gate][1] = br[gate][0]; d[gate][2] = br[gate][1]; d[gate][3] = br[gate][1]; }
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(hcur + kt * 16), bb1 = *reinterpret_cast<const uint32_t*>(hcur + kt * 16 + 8);
#pragma unroll
      for (int gate = 0; gate < 3; ++gate) mma_16816_f16(d[gate], a[gate][kt].x, a[gate][kt].y, a[gate][kt].z, a[gate][kt].w, bb0, bb1);
    }
    __half* hnext = hb + (size_t)((t & 1) ^ 1) * GM_NU * GM_HS;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + 8 * (e >> 1), u = u0 + (e & 1);
      if (t < ns_u[e & 1]) {   // the reference keeps the state at step ceil(len / compress) - 1 (gather_nd, GST.py:65-68)
        const float z = sigmoid_acc(x[0][e] + d[0][e]);
        const float r = sigmoid_acc(x[1][e] + d[1][e]);
        const float c = tanhf(x[2][e] + r * d[2][e]);
        h[e] = z * h[e] + (1.0f - z) * c;
      }
      hnext[u * GM_HS + n] = __float2half_rn(h[e]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) hf[(u0 + (e & 1)) * GM_G + n0 + 8 * (e >> 1)] = h[e];
  __syncthreads();
  // ---- tail, batched over the utterances: Dense(tanh)
  for (int i = tid; i < nu * D; i += nthr) {
    const int u = i / D, n = i - u * D;
    float acc = __ldg(p.bd + n);
    for (int k = 0; k < GM_G; ++k) acc = fmaf(hf[u * GM_G + k], __ldg(p.Wd + (size_t)k * D + n), acc);
    acc = tanhf(acc);
    ref_s[u * D + n] = acc;
    if (p.out_ref) p.out_ref[(size_t)(b0 + u) * D + n] = acc;
  }
  __syncthreads();
  if (!p.out_gst && !p.out_att) return;
  for (int i = tid; i < nu * S; i += nthr) {   // query projection
    const int u = i / S, n = i - u * S;
    float acc = __ldg(p.bq + n);
    for (int k = 0; k < D; ++k) acc = fmaf(ref_s[u * D + k], __ldg(p.Wq + (size_t)k * S + n), acc);
    q_s[i] = acc;
  }
  __syncthreads();
  const int hd = S / heads;
  for (int i = tid; i < nu * heads * NT; i += nthr) {   // unscaled dot-product scores, one thread per (utterance, head, token)
    const int u = i / (heads * NT), r = i - u * heads * NT, hh = r / NT, tok = r - hh * NT;
    float acc = 0.f;
    for (int dd = 0; dd < hd; ++dd) acc = fmaf(q_s[u * S + hh * hd + dd], __ldg(p.tokkv + (size_t)tok * S + hh * hd + dd), acc);
    pr_s[i] = acc;
  }
  __syncthreads();
  if (tid < nu * heads) {   // softmax over tokens (max-subtracted)
    float* pr = pr_s + tid * NT;
    float m = -INFINITY;
    for (int tok = 0; tok < NT; ++tok) m = fmaxf(m, pr[tok]);
    float ssum = 0.f;
    for (int tok = 0; tok < NT; ++tok) {
      const float ex = expf(pr[tok] - m);
      pr[tok] = ex;
      ssum += ex;
    }
    for (int tok = 0; tok < NT; ++tok) pr[tok] /= ssum;
  }
  __syncthreads();
  for (int i = tid; i < nu * S; i += nthr) {
    const int u = i / S, n = i - u * S, hh = n / hd;
    float acc = 0.f;
    for (int tok = 0; tok < NT; ++tok) acc = fmaf(pr_s[(u * heads + hh) * NT + tok], __ldg(p.tokkv + (size_t)tok * S + n), acc);
    y_s[i] = acc + q_s[i];  // residual with the projected query (Layers.py:211)
  }
  if (p.out_att)
    for (int i = tid; i < nu * NT; i += nthr) {
      const int u = i / NT, tok = i - u * NT;
      float acc = 0.f;
      for (int hh = 0; hh < heads; ++hh) acc += pr_s[(u * heads + hh) * NT + tok];
      p.out_att[(size_t)(b0 + u) * NT + tok] = acc / (float)heads;
    }
  __syncthreads();
  if (wid < nu) {  // Layer_Norm statistics (biased variance, eps inside sqrt): one warp per utterance
    const float* y = y_s + wid * S;
    float ssum = 0.f;
    for (int n = lane; n < S; n += 32) ssum += y[n];
    const float mean = warp_sum(ssum) / (float)S;
    float v = 0.f;
    for (int n = lane; n < S; n += 32) {
      const float dlt = y[n] - mean;
      v = fmaf(dlt, dlt, v);
    }
    const float var = warp_sum(v) / (float)S;
    if (lane == 0) {
      sc_s[2 * wid] = mean;
      sc_s[2 * wid + 1] = 1.0f / sqrtf(var + 1e-8f);
    }
  }
  __syncthreads();
  if (p.out_gst)
    for (int i = tid; i < nu * S; i += nthr) {
      const int u = i / S, n = i - u * S;
      p.out_gst[(size_t)(b0 + u) * S + n] = __ldg(p.ln_g + n) * ((y_s[i] - sc_s[2 * u]) * sc_s[2 * u + 1]) + __ldg(p.ln_b + n);
    }
}
inline size_t gru_mma_smem_bytes(int D, int S, int NT, int heads) {
  return (size_t)2 * GM_NU * GM_HS * 2 + sizeof(float) * ((size_t)GM_NU * (GM_G + D + 2 * S + heads * NT + 2));
}

}  // namespace gstk
