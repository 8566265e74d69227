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
// sigmoid recurrent activation, zero initial state, no mask).  xs = x.W + b for both directions, [B][T][2][u][4 gates] fp32
// (the host permutes the columns of the concatenated input kernel so that the 4 gates of a unit are one 16 B load).
// CTA = (NB utterances, direction); thread j = hidden unit j: its 4 gate columns for NB utterances, c in registers,
// h double-buffered in shared memory.  U (u x 4u fp32) streams from L2 every step, shared by the CTA's NB utterances.
// Fall-back for RNN sizes whose persistent grid (below) does not fit the device.  Measured on B200 (256 x 150 tokens, 128
// CTAs): 27 us per step = 4.1 ms, i.e. 128 x 1 MB / 27 us = 4.9 TB/s of L2 -> SM traffic for the U stream.  Tried, slower:
// 32 U loads issued ahead of their FMAs (+12 %), 8 utterances per CTA on 64 CTAs (+15 %).
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
      const float4 x = __ldg(reinterpret_cast<const float4*>(xs + ((size_t)(ok ? b0 + n : b0) * T + t) * 8 * u + (size_t)dir * 4 * u) + j);
      acc[0][n] = ok ? x.x : 0.f; acc[1][n] = ok ? x.y : 0.f; acc[2][n] = ok ? x.z : 0.f; acc[3][n] = ok ? x.w : 0.f;
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

// ---------------------------------------------------------------------------------------------------------------------
// Persistent form of the same recurrence (the decoder's structure): one cooperative launch, 2 x (u / 4) CTAs; CTA = (direction,
// 4 hidden units).  Its 16 gate columns of the recurrent kernel (u x 16 fp32 = 16 KB) stay RESIDENT in shared memory for all
// T steps - nothing of U is re-read - and every CTA processes ALL (<= 256) utterances for its units: per step it stages
// h(t-1) [B][u] from L2 (ld.global.cg, 64-column chunks), accumulates z = xs + h.U_slice with thread = (unit, utterance
// quarter) x 4 utterances, applies the cell update (c in registers) and publishes its 4 columns of h(t) to the other CTAs of
// its direction through a double-buffered global image + one grid barrier per step.
// Measured on B200 (u = 256, T_v = 150): 2.9 ms = 19 us per step for 256 utterances (streaming kernel: 4.1 ms), 1.1 ms = 7 us
// per step for one.  Steps of the way: first version 31 us (h chunks staged with a block barrier either side) -> 2-deep cp.async
// ring 20.5 us -> warps without live utterances skip the FMA loop and only live rows are staged (ncu source page: 2.8 k
// warp-instructions per warp and step were loop overhead of idle warps) 19 us / 7 us.  FFMA-issue-bound (4096 FMA per
// thread and step): this is the exact-mode kernel; the tensor-core mode uses encoder_bilstm_tc_kernel below.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BL_HU = 4, BL_THREADS = 256, BL_ROWS = 256, BL_KC = 64, BL_HS_LD = BL_KC + 4;

struct BilstmParams {
  const float* xs;    // [B][T][2][u][4 gates]
  const float* Uf;    // [u][4u]
  const float* Ub;
  float* out;         // [B][T][2u]
  float* hbuf;        // [2 directions][2 parities][BL_ROWS][u], zeroed before the launch
  GridBarrier* gb;    // zeroed before the launch
  int B, T, u;
};

inline size_t bilstm_persistent_smem(int u) { return (size_t)u * 16 * 4 + (size_t)2 * BL_ROWS * BL_HS_LD * 4; }

__global__ void __launch_bounds__(BL_THREADS) encoder_bilstm_persistent_kernel(const BilstmParams p) {
  extern __shared__ __align__(16) float bl_smem[];
  __shared__ int ok_s;
  float* Us = bl_smem;                 // [u][unit][gate]
  float* hs = bl_smem + p.u * 16;      // [2][BL_ROWS][BL_HS_LD]
  const int tid = threadIdx.x, u = p.u, B = p.B, T = p.T;
  const int nc = u / BL_HU, dir = blockIdx.x / nc, hu0 = (blockIdx.x % nc) * BL_HU;
  const float* __restrict__ U = dir ? p.Ub : p.Uf;
  for (int idx = tid; idx < u * 16; idx += BL_THREADS) {
    const int k = idx >> 4, unit = (idx & 15) >> 2, gate = idx & 3;
    Us[idx] = __ldg(U + (size_t)k * 4 * u + (size_t)gate * u + hu0 + unit);
  }
  const int cg = tid & 3, bq = tid >> 2;   // my hidden unit; utterances bq + 64 i
  const int bqw = (tid >> 5) * 8;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned int gen = 0;
  __syncthreads();
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const float* hin = p.hbuf + (size_t)(dir * 2 + (s & 1)) * BL_ROWS * u;
    float* hout = p.hbuf + (size_t)(dir * 2 + ((s + 1) & 1)) * BL_ROWS * u;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = bq + 64 * i;
      const float4 x = __ldg(reinterpret_cast<const float4*>(p.xs + ((size_t)(b < B ? b : 0) * T + t) * 8 * u + (size_t)dir * 4 * u) + hu0 + cg);
      acc[i][0] = b < B ? x.x : 0.f; acc[i][1] = b < B ? x.y : 0.f; acc[i][2] = b < B ? x.z : 0.f; acc[i][3] = b < B ? x.w : 0.f;
    }
    // h(t-1) arrives in 64-column chunks through a 2-deep cp.async ring (L2 only: .cg), so that the L2 round trip of chunk
    // n+1 overlaps the FMAs of chunk n
    auto stage = [&](int kc, float* dst) {
      for (int idx = tid; idx < B * (BL_KC / 4); idx += BL_THREADS) {   // only the B live rows
        const int row = idx >> 4, c4 = idx & 15;
        cp_async16(dst + row * BL_HS_LD + c4 * 4, hin + (size_t)row * u + kc + c4 * 4, true);
      }
      cp_async_commit();
    };
    const int nchunks = u / BL_KC;
    stage(0, hs);
    for (int ci = 0; ci < nchunks; ++ci) {
      const int kc = ci * BL_KC;
      const float* hcur = hs + (ci & 1) * BL_ROWS * BL_HS_LD;
      if (ci + 1 < nchunks) stage(kc + BL_KC, hs + ((ci + 1) & 1) * BL_ROWS * BL_HS_LD);
      else cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      // warp-uniform guards (bqw = the warp's first utterance): a warp without live utterances skips the chunk, and lanes
      // past B inside a live warp accumulate garbage that is never stored - no per-lane branches in the FMA loop
      if (bqw < B) {
#pragma unroll 4
      for (int k4 = 0; k4 < BL_KC; k4 += 4) {
        float4 w[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Us + (kc + k4 + kk) * 16 + cg * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (bqw + 64 * i < B) {
            const float4 hv = *reinterpret_cast<const float4*>(hcur + (bq + 64 * i) * BL_HS_LD + k4);
            const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              acc[i][0] = fmaf(hk[kk], w[kk].x, acc[i][0]);
              acc[i][1] = fmaf(hk[kk], w[kk].y, acc[i][1]);
              acc[i][2] = fmaf(hk[kk], w[kk].z, acc[i][2]);
              acc[i][3] = fmaf(hk[kk], w[kk].w, acc[i][3]);
            }
          }
        }
      }
      }
      __syncthreads();   // the buffer is refilled two chunks later
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = bq + 64 * i;
      if (b < B) {
        c[i] = enc_sigmoid(acc[i][1]) * c[i] + enc_sigmoid(acc[i][0]) * tanhf(acc[i][2]);
        const float hnew = enc_sigmoid(acc[i][3]) * tanhf(c[i]);
        hout[(size_t)b * u + hu0 + cg] = hnew;
        p.out[((size_t)b * T + t) * 2 * u + (size_t)dir * u + hu0 + cg] = hnew;
      }
    }
    if (s + 1 < T && !grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core form of the persistent recurrence (handle precision "bf16", u = 256): same grid, barrier and h exchange, but
// h travels as fp16 (half the L2 traffic), the CTA's 256 x 16 slice of the recurrent kernel lives in REGISTERS as
// mma.sync.m16n8k16 B fragments (64 registers per thread, loaded once), and z = h.U_slice is 2 m-tiles x 2 n-tiles x 16
// k-tiles of mma.sync per warp and step (warp w = utterances 32w .. 32w+31) with A fragments by ldmatrix from the staged
// h chunks.  Column order of the slice is [unit][gate], so the 4 gates of a (row, unit) sit in two adjacent lanes; one
// shfl_xor pair regroups them (even lane: row r, odd lane: row r + 8).  fp16 (not bf16) operands as for the Postnet:
// |h| < 1 and |U| ~ 0.06 are well inside its range and the 11-bit mantissa keeps the recurrence within the 1e-2 tolerance.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BT_U = 256, BT_LD = BL_KC + 8;   // halves per staged row (144 B: conflict-free ldmatrix)

struct BilstmTcParams {
  const float* xs;
  const float* Uf;
  const float* Ub;
  float* out;
  __half* hbuf;       // [2 directions][2 parities][BL_ROWS][BT_U] fp16, zeroed before the launch
  GridBarrier* gb;
  int B, T;
};

constexpr size_t BT_SMEM = (size_t)2 * BL_ROWS * BT_LD * 2;

__global__ void __launch_bounds__(BL_THREADS) encoder_bilstm_tc_kernel(const BilstmTcParams p) {
  extern __shared__ __align__(16) unsigned char bt_smem[];
  __shared__ int ok_s;
  __half* hs = reinterpret_cast<__half*>(bt_smem);   // [2][BL_ROWS][BT_LD]
  constexpr int u = BT_U;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, B = p.B, T = p.T;
  const int nc = u / BL_HU, dir = blockIdx.x / nc, hu0 = (blockIdx.x % nc) * BL_HU;
  const float* __restrict__ U = dir ? p.Ub : p.Uf;
  // B fragments: bfrag[kt][nt][0] = {U[k][n], U[k+1][n]}, [1] = k + 8; k = 16 kt + 2 (lane % 4), n = 8 nt + lane / 4
  unsigned bfrag[u / 16][2][2];
#pragma unroll
  for (int kt = 0; kt < u / 16; ++kt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int n = nt * 8 + (lane >> 2), unit = n >> 2, gate = n & 3;
      const float* col = U + (size_t)gate * u + hu0 + unit;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = kt * 16 + hh * 8 + 2 * (lane & 3);
        const __half2 v = __floats2half2_rn(__ldg(col + (size_t)k * 4 * u), __ldg(col + (size_t)(k + 1) * 4 * u));
        bfrag[kt][nt][hh] = *reinterpret_cast<const unsigned*>(&v);
      }
    }
  const int q = lane & 3, odd = q & 1;
  const int row_in_tile = (lane >> 2) + odd * 8;       // the row this lane finishes after the regrouping
  float cst[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  unsigned int gen = 0;
  const bool warp_live = warp * 32 < B;
  // gate pre-activations of the (row, unit) pairs this lane finishes: one 16 B load each ([unit][gate] order), fetched one
  // step ahead so that their HBM latency hides behind the grid barrier (ncu: 26 % of the stalls were these loads)
  float4 xn[2][2];
  auto load_x = [&](int t) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int b = warp * 32 + mt * 16 + row_in_tile, unit = nt * 2 + (q >> 1);
        xn[mt][nt] = b < B ? __ldg(reinterpret_cast<const float4*>(p.xs + ((size_t)b * T + t) * 8 * u + (size_t)dir * 4 * u) + hu0 + unit)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  };
  load_x(dir ? T - 1 : 0);
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const __half* hin = p.hbuf + (size_t)(dir * 2 + (s & 1)) * BL_ROWS * u;
    __half* hout = p.hbuf + (size_t)(dir * 2 + ((s + 1) & 1)) * BL_ROWS * u;
    float xr[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        xr[mt][nt][0] = xn[mt][nt].x; xr[mt][nt][1] = xn[mt][nt].y; xr[mt][nt][2] = xn[mt][nt].z; xr[mt][nt][3] = xn[mt][nt].w;
      }
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    auto stage = [&](int kc, __half* dst) {
      for (int idx = tid; idx < B * (BL_KC / 8); idx += BL_THREADS) {
        const int row = idx >> 3, c8 = idx & 7;
        cp_async16(dst + row * BT_LD + c8 * 8, hin + (size_t)row * u + kc + c8 * 8, true);
      }
      cp_async_commit();
    };
    constexpr int nchunks = u / BL_KC;
    stage(0, hs);
#pragma unroll
    for (int ci = 0; ci < nchunks; ++ci) {
      const __half* hcur = hs + (ci & 1) * BL_ROWS * BT_LD;
      if (ci + 1 < nchunks) stage((ci + 1) * BL_KC, hs + ((ci + 1) & 1) * BL_ROWS * BT_LD);
      else cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      if (warp_live) {
#pragma unroll
        for (int kk = 0; kk < BL_KC / 16; ++kk) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (warp * 32 + mt * 16 < B) {
              unsigned af[4];
              ldmatrix_x4(af, hcur + (warp * 32 + mt * 16 + (lane & 15)) * BT_LD + kk * 16 + (lane >> 4) * 8);
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
                mma_f16_16816(acc[mt][nt], af, bfrag[ci * (BL_KC / 16) + kk][nt][0], bfrag[ci * (BL_KC / 16) + kk][nt][1]);
            }
          }
        }
      }
      __syncthreads();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        // even lane: has (i, f) of rows r and r + 8, keeps row r; odd lane: has (c, o), keeps row r + 8
        const float s0 = odd ? acc[mt][nt][0] : acc[mt][nt][2], s1 = odd ? acc[mt][nt][1] : acc[mt][nt][3];
        const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        const float zi = (odd ? r0 : acc[mt][nt][0]) + xr[mt][nt][0], zf = (odd ? r1 : acc[mt][nt][1]) + xr[mt][nt][1];
        const float zc = (odd ? acc[mt][nt][2] : r0) + xr[mt][nt][2], zo = (odd ? acc[mt][nt][3] : r1) + xr[mt][nt][3];
        const int b = warp * 32 + mt * 16 + row_in_tile, unit = nt * 2 + (q >> 1);
        if (b < B) {
          cst[mt][nt] = enc_sigmoid(zf) * cst[mt][nt] + enc_sigmoid(zi) * tanhf(zc);
          const float hnew = enc_sigmoid(zo) * tanhf(cst[mt][nt]);
          hout[(size_t)b * u + hu0 + unit] = __float2half_rn(hnew);
          p.out[((size_t)b * T + t) * 2 * u + (size_t)dir * u + hu0 + unit] = hnew;
        }
      }
    if (s + 1 < T) load_x(dir ? T - 2 - s : s + 1);
    if (s + 1 < T && !grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Cluster form of the recurrence (tensor-core mode, u = 256): the utterances of a batch are INDEPENDENT recurrences, so they are cut
// into groups of 16 and every (group, direction) runs on its own cluster of 4 CTAs - no grid barrier, no h round trip through L2.
// CTA r of a cluster owns units 64 r .. 64 r + 63 of all four gates; warp w owns 8 of them.  Its slice of the recurrent kernel lives
// in REGISTERS for the whole sequence as mma.sync A fragments (gate columns = M: m-tile 0 = gates i | f of the warp's 8 units,
// m-tile 1 = gates c | o; 2 x 16 k-tiles x 4 = 128 registers), the group's 16 utterances are the N columns, h(t-1) [16][256] fp16 is
// the B operand in shared memory.  The four gates of (unit, utterance) land in one lane: cell state and h(t) are computed in
// registers, h(t) is written as fp16 straight into the shared memory of all four CTAs (st.shared::cluster, two units per store) and
// one barrier.cluster per step publishes it.  The input projections (one 16 B load per (unit, utterance), [unit][gate] order) are
// fetched one step ahead.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BC_CL = 4, BC_NB = 16, BC_U = 256, BC_UC = BC_U / BC_CL, BC_THREADS = 256, BC_LD = BC_U + 8;

struct BilstmClParams {
  const float* xs;    // [B][T][8u]: columns dir * 4u + unit * 4 + gate (input projection + bias); or nullptr:
  const __half* xs16; // the same as fp16 rows of the flat padded matrix (row b * R + PADL + t, postnet.cuh) - half the traffic of the
  int R, PADL;        // largest tensor of the Encoder / vocoder
  const float* Uf;    // [u][4u] recurrent kernels (Keras gate blocks i | f | c | o)
  const float* Ub;
  float* out;         // [B][T][2u] = [forward | backward]
  int B, T;
};

__device__ __forceinline__ void st_cluster_u32(uint32_t local_saddr, uint32_t rank, uint32_t v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
// MUFU forms for the tensor-core mode (h is exchanged as fp16; relative error 2^-11)
__device__ __forceinline__ float bc_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bc_sigmoid(float x) { return fmaf(0.5f, bc_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(BC_CL, 1, 1) __launch_bounds__(BC_THREADS, 1) encoder_bilstm_cluster_kernel(const BilstmClParams p) {
  __shared__ __align__(16) __half hs[2][BC_NB][BC_LD];
  constexpr int u = BC_U;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g4 = lane >> 2, t4 = lane & 3;
  const int rank = blockIdx.x, b0 = blockIdx.y * BC_NB, dir = blockIdx.z, T = p.T;
  const int unit = rank * BC_UC + warp * 8 + g4;   // the unit whose gates this lane finishes
  const float* __restrict__ U = dir ? p.Ub : p.Uf;
  // A fragments of m16n8k16 (row-major 16 x 16): a0 = (row g4, k 2 t4 ..+1), a1 = (row g4 + 8, same k), a2 / a3 = k + 8
  unsigned afrag[2][u / 16][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int kt = 0; kt < u / 16; ++kt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int gate = 2 * mt + (r & 1), k = kt * 16 + 2 * t4 + (r >> 1) * 8;
        const float* col = U + (size_t)gate * u + unit;
        const __half2 v = __floats2half2_rn(__ldg(col + (size_t)k * 4 * u), __ldg(col + (size_t)(k + 1) * 4 * u));
        afrag[mt][kt][r] = *reinterpret_cast<const unsigned*>(&v);
      }
  for (int i = tid; i < 2 * BC_NB * BC_LD / 2; i += BC_THREADS) reinterpret_cast<unsigned*>(&hs[0][0][0])[i] = 0u;
  float cst[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  // raw bits of the next step's projections: decoded when they are USED (a conversion at load time would wait for the load
  // at the top of every step - measured: 4.0 instead of 1.7 us per step)
  uint4 xn[2][2];
  auto load_x = [&](int t) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int b = b0 + nt * 8 + 2 * t4 + e;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (b < p.B) {
          if (p.xs16) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.xs16 + ((size_t)b * p.R + p.PADL + t) * 8 * u + (size_t)dir * 4 * u) + unit);
            v.x = raw.x;
            v.y = raw.y;
          } else {
            v = __ldg(reinterpret_cast<const uint4*>(p.xs + ((size_t)b * T + t) * 8 * u + (size_t)dir * 4 * u) + unit);
          }
        }
        xn[nt][e] = v;
      }
  };
  auto decode_x = [&](const uint4& v) {
    if (p.xs16) {
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
  };
  load_x(dir ? T - 1 : 0);
  cluster_sync_all();   // every CTA's buffers are zero before anybody writes into them
  const uint32_t hs_sa = (uint32_t)__cvta_generic_to_shared(&hs[0][0][0]);
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const int cur = s & 1, nxt = cur ^ 1;
    float4 xr[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) xr[nt][e] = decode_x(xn[nt][e]);
    if (s + 1 < T) load_x(dir ? T - 2 - s : s + 1);
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int kt = 0; kt < u / 16; ++kt) {
      // B fragments of both n-tiles: matrices (nt, k half) = (lane >> 4, (lane >> 3) & 1), rows = utterances, 8 k values per row
      unsigned bf[4];
      ldmatrix_x4(bf, &hs[cur][(lane >> 4) * 8 + (lane & 7)][kt * 16 + ((lane >> 3) & 1) * 8]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_f16_16816(acc[mt][0], afrag[mt][kt], bf[0], bf[1]);
        mma_f16_16816(acc[mt][1], afrag[mt][kt], bf[2], bf[3]);
      }
    }
    // gates of (unit, utterance nt * 8 + 2 t4 + e): i = acc[0][nt][e], f = acc[0][nt][2 + e], c~ = acc[1][nt][e], o = acc[1][nt][2 + e]
    float hn[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float x4[4] = {xr[nt][e].x, xr[nt][e].y, xr[nt][e].z, xr[nt][e].w};
        const float zi = acc[0][nt][e] + x4[0], zf = acc[0][nt][2 + e] + x4[1], zc = acc[1][nt][e] + x4[2], zo = acc[1][nt][2 + e] + x4[3];
        cst[nt][e] = bc_sigmoid(zf) * cst[nt][e] + bc_sigmoid(zi) * bc_tanh(zc);
        hn[nt][e] = bc_sigmoid(zo) * bc_tanh(cst[nt][e]);
        const int b = b0 + nt * 8 + 2 * t4 + e;
        if (b < p.B) p.out[((size_t)b * T + t) * 2 * u + (size_t)dir * u + unit] = hn[nt][e];
      }
    if (s + 1 < T) {
      // pair units (g4, g4 ^ 1): the even lane of a pair writes utterance e = 0 of both units, the odd lane utterance e = 1
      const int odd = g4 & 1;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float other = __shfl_xor_sync(0xffffffffu, odd ? hn[nt][0] : hn[nt][1], 4);
        const __half2 v = odd ? __floats2half2_rn(other, hn[nt][1]) : __floats2half2_rn(hn[nt][0], other);
        const uint32_t off = (uint32_t)(((nxt * BC_NB + nt * 8 + 2 * t4 + odd) * BC_LD + (unit & ~1)) * 2);
#pragma unroll
        for (int r = 0; r < BC_CL; ++r) st_cluster_u32(hs_sa + off, (uint32_t)r, *reinterpret_cast<const unsigned*>(&v));
      }
      cluster_sync_all();
    }
  }
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gstk
