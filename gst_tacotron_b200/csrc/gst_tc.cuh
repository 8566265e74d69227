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
constexpr int G0_HO = 8;
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
  for (int i = threadIdx.x; i < nrow * Wp; i += blockDim.x) {
    const int rr = i / Wp, cc = i - rr * Wp;
    const int hh = 2 * ho0 - pt + rr, ww = cc - pl;
    in_h[i] = __float2half_rn((hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xb + (size_t)hh * W + ww) : 0.f);
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

}  // namespace gstk
