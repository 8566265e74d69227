// Small-batch bf16 decoder (batch <= 8, free running, SMA, default widths): the latency kernel (north_star: "warp-shuffle FMA
// below batch 64 ... weights resident in shared memory"; SURVEY row NS-a).
//
// At batch 1 the batch-256 kernel (decoder_bf16.cuh) needs 20 us per step: 6.5 us of operand-ring overhead (16 + 6 pipeline units at
// ~500 clk each, whatever they carry), 5.9 us for one SM to pull the 423 KB of dense weights through its ring, three grid barriers.
// This kernel is built for the latency chain instead:
//
//   * 128 LSTM CTAs x 8 hidden units of BOTH cells (32 gate rows per cell = two m16 tiles).  Every product is a GEMV over <= 8 batch
//     columns on mma.sync m16n8k16 (weights = A fragments, batch = N), K split over the 8 warps, one shared-memory reduction.
//     The fragments that sit on the critical path - W1x (prenet + context -> LSTMCell 0) and W2 (h1 -> LSTMCell 1), 88 KB per CTA -
//     are RESIDENT in shared memory for the whole decode; the recurrent products h1(t-1).U1 and h2(t-1).U2 (128 KB of fragments per
//     CTA and step) stream straight from L2 into registers during the dense / attention window, off the critical path.
//   * one FRONT CTA per utterance (CTAs 128 .. 128 + B): projection -> prenet x2 -> query (Taco2.py:106-118, 270-283; Steps.py:122), then
//     the stepwise-monotonic attention of that utterance (Steps.py:138-166, 215-229).  The 221 KB projection never streams through the
//     front CTA: every LSTM CTA publishes, next to its 8 units of h2, their split-K partial of the projection (8 x 81 FMAs per
//     utterance), the front CTA sums the 128 partials (41 KB, one L2 round trip) and adds the context part, which it computed while the
//     LSTM phases ran.  prenet-1 and query fragments are RESIDENT in its shared memory (192 KB), the prenet-0 fragments are requested
//     into registers before h2 arrives, the layers are split by output features over the warps (no cross-warp reductions), and the
//     projected keys of the utterance stay in registers for the whole decode (kernel templated on the number of key iterations).
//   * three hand-overs per step, each one arrival counter in global memory (no grid barrier): h2 -> front CTAs + U2 products,
//     [p || ctx] -> LSTMCell 0, h1 -> LSTMCell 1.  h1 / h2 / x travel as small bf16 row buffers, double-buffered by step parity.
//   * cell states live in registers of the thread that owns (utterance, unit); dropout flags and attention noise of step t+1 are
//     drawn while the CTA waits.
#pragma once
#include "decoder_bf16_v2.cuh"

namespace gstk {

constexpr int SB_THREADS = 256;                  // eight warps, all of them workers (<= 255 registers: the fragments in flight need them)
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_MAXB = 16;                      // utterances per launch: NT = 1 or 2 mma n-tiles of 8 batch columns (kernel template)
constexpr int SB_UNITS = 8;                      // hidden units per LSTM CTA  => TC_U / SB_UNITS = 128 LSTM CTAs
constexpr int SB_KT_X = TC_KX / 16, SB_KT_H = TC_U / 16;   // 24, 64 k16-tiles
// per-CTA fragment image (tiles of 512 B): [W1x: 2 x 24][W2: 2 x 64] resident, then [U1: 2 x 64][U2: 2 x 64] streamed
constexpr int SB_T_W1X = 0, SB_T_W2 = 2 * SB_KT_X, SB_T_U1 = SB_T_W2 + 2 * SB_KT_H, SB_T_U2 = SB_T_U1 + 2 * SB_KT_H, SB_TILES = SB_T_U2 + 2 * SB_KT_H;
constexpr int SB_RES_BYTES = SB_T_U1 * 512;      // 90,112 B resident
constexpr int SB_HS = TC_U + 8, SB_XS = TC_KX + 8, SB_AS = FA_HC + 8;   // bf16 row strides (4 mod 32 words): conflict-free B fragments

struct SbSync {
  unsigned int h1cnt[32];   // LSTM CTAs that have published h1(t)
  unsigned int h2cnt[32];   // ... h2(t)
  unsigned int xcnt[32];    // front CTAs that have published [p || ctx](t)
};

struct SbParams {
  const uint4* wl;                // [128 CTAs][SB_TILES][32 lanes] LSTM fragments
  const float* bl;                // [128][2 cells][32]  bias, row = gate * 8 + unit
  const uint8_t* wimgA;           // dense fragments (decoder_bf16.cuh: fa_wlayer)
  const __nv_bfloat16* vproj_bf;  // [B][Tv][128]
  __nv_bfloat16* hbuf1;           // [2][SB_MAXB][TC_U]
  __nv_bfloat16* hbuf2;           // [2][SB_MAXB][TC_U]
  __nv_bfloat16* xbuf;            // [2][SB_MAXB][TC_KX]
  float* ppart;                   // [2][SB_MAXB][128 LSTM CTAs][96] split-K partials of the projection: sum over the CTA's 8 units of h2 . P
  SbSync* sync;
  unsigned long long* prof;
};

// 16 B load through L2 (ld.global.cg) that the compiler may not sink to its first use: the fragments requested "before the wait"
// must really be issued there
__device__ __forceinline__ uint4 sb_ldcg_pinned(const uint4* ptr) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}

// partial GEMV of one warp: d[nt][m] += sum_{kt = kt0, kt0 + 8, ...} A(m, kt) . act[8 nt ..][kt]   (A tile (m, kt) at w[(m * NKT + kt) * 32 + lane])
template <bool GLOBAL, int NT>
__device__ __forceinline__ void sb_gemv_warp(float (&d)[NT][2][4], const uint4* __restrict__ w, int NKT, int kt0, const __nv_bfloat16* act_s, int stride, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* brow = act_s + g * stride + 2 * t;
  w += lane;
#pragma unroll 4
  for (int kt = kt0; kt < NKT; kt += SB_WARPS) {
    const uint4 a0 = GLOBAL ? __ldcg(w + (size_t)kt * 32) : w[(size_t)kt * 32];
    const uint4 a1 = GLOBAL ? __ldcg(w + (size_t)(NKT + kt) * 32) : w[(size_t)(NKT + kt) * 32];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const __nv_bfloat16* br = brow + nt * 8 * stride;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(br + kt * 16), b1 = *reinterpret_cast<const uint32_t*>(br + kt * 16 + 8);
      mma_16816_bf16(d[nt][0], a0, b0, b1);
      mma_16816_bf16(d[nt][1], a1, b0, b1);
    }
  }
}
// all warps: dst[nt][row][u] = add[nt][row][u] + bias[row] + sum_k A[row][k] act[8 nt + u][k]  (32 gate rows x 8 NT batch columns); two block barriers
template <bool GLOBAL, int NT>
__device__ __forceinline__ void sb_matvec(float* dst, const float* add, const float* bias, const uint4* w, int NKT, const __nv_bfloat16* act_s, int stride,
                                          float* red, int wid, int lane) {
  float d[NT][2][4] = {};
  sb_gemv_warp<GLOBAL, NT>(d, w, NKT, wid, act_s, stride, lane);
  float4* r4 = reinterpret_cast<float4*>(red) + (wid * NT * 2) * 32 + lane;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    r4[(nt * 2) * 32] = make_float4(d[nt][0][0], d[nt][0][1], d[nt][0][2], d[nt][0][3]);
    r4[(nt * 2 + 1) * 32] = make_float4(d[nt][1][0], d[nt][1][1], d[nt][1][2], d[nt][1][3]);
  }
  pa_sync<SB_THREADS>();
  const int tid = wid * 32 + lane;
  if (tid < 256) {   // (row, batch column): D fragment element (lane = (row % 8) * 4 + col / 2, slot = (row % 16) / 8 * 2 + col % 2) of tile row / 16
    const int row = tid >> 3, u = tid & 7;
    const int m = row >> 4, ln = (row & 7) * 4 + (u >> 1), sl = ((row >> 3) & 1) * 2 + (u & 1);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float s = (add ? add[nt * 256 + tid] : 0.f) + (bias ? bias[row] : 0.f);
#pragma unroll
      for (int w2 = 0; w2 < SB_WARPS; ++w2) s += red[(((w2 * NT + nt) * 2 + m) * 32 + ln) * 4 + sl];
      dst[nt * 256 + tid] = s;
    }
  }
  pa_sync<SB_THREADS>();
}

__device__ __forceinline__ float sb_keep(const DecParams& p, int layer, int t, int b, int n) {
  if (!(p.rng_mode != 0 && p.drop_rate > 0.f)) return 1.f;
  if (p.rng_mode == 1) return __ldg((layer ? p.keep1 : p.keep0) + ((size_t)t * p.rngB + p.rng_b0 + b) * FA_P + n) != 0.f ? p.drop_scale : 0.f;
  return philox_keep(p.seed, layer ? STREAM_KEEP1 : STREAM_KEEP0, p.step_offset + (unsigned int)t, p.row_offset + b, (unsigned int)n, p.drop_rate) != 0.f
             ? p.drop_scale : 0.f;
}

// front CTA shared memory: resident prenet-1 | query fragments, layer inputs (bf16 rows), fp32 scratch, attention scratch
constexpr int SB_FRES_BYTES = (int)(FA_L1.nst() * FA_L1.stride() + FA_LQ.nst() * FA_LQ.stride());   // 192 KB (layers 2, 3 of the dense image are adjacent)
constexpr int SB_FACT_ELEMS = 96 + 2 * (FA_P + 8);                                                    // mel | z0 | z1  (bf16 layer inputs)
constexpr int SB_FRONT_FLOATS = 96 + FA_A + 128 + 2 * FA_P + SB_WARPS * 96;                           // mel | q | ctx | keep scales x2 | projection partials
__host__ __device__ constexpr size_t sb_front_bytes(int Tv) {
  return (size_t)SB_FRES_BYTES + (size_t)SB_FACT_ELEMS * 2 + 4 * ((size_t)SB_FRONT_FLOATS + (size_t)(((4 * Tv + 3) & ~3) + SB_WARPS * 128));
}
__host__ __device__ constexpr size_t sb_lstm_bytes(int NT) {
  return (size_t)SB_RES_BYTES + (size_t)(8 * NT) * (2 * SB_HS + SB_XS) * 2 +
         4 * (size_t)(SB_WARPS * NT * 2 * 32 * 4 + 3 * NT * 256 + 64 + SB_UNITS * 96 + 8 * NT * SB_UNITS);
}

// KVI = key iterations a lane keeps in registers (4 rows x 16 columns each): 3 for key_time <= 96, 5 for <= 160, 8 for <= 256.  The array
// must be no larger than needed: with 8 iterations next to the dense fragments ptxas spills all 64 registers of it (LDL in both passes).
// NT = mma n-tiles of the LSTM GEMVs: 1 for batch <= 8, 2 for batch 9 .. 16 (every weight fragment is then used for both tiles).
template <int KVI, int NT>
__global__ void __launch_bounds__(SB_THREADS, 1) decoder_bf16_sb_kernel(const __grid_constant__ DecParams p, const __grid_constant__ SbParams q) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ unsigned long long prof_sh[PROF_SLOTS + 1];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int cta = blockIdx.x, B = p.B, T = p.T;
  const int NL = TC_U / SB_UNITS;   // 128 LSTM CTAs
  SbSync* sy = q.sync;
  unsigned long long* prof_s = q.prof ? prof_sh : nullptr;
  if (tid == 0) {
    for (int i = 0; i < PROF_SLOTS; ++i) prof_sh[i] = 0;
    prof_sh[PROF_SLOTS] = (unsigned long long)clock64();
  }
  if (cta < NL) {
    // =========================================== LSTM CTA: units [8 cta, 8 cta + 8) of both cells ===========================================
    uint4* wres = reinterpret_cast<uint4*>(sm);
    __nv_bfloat16* h1s = reinterpret_cast<__nv_bfloat16*>(sm + SB_RES_BYTES);
    constexpr int MB = 8 * NT;                 // utterance rows held in shared memory
    __nv_bfloat16* h2s = h1s + MB * SB_HS;
    __nv_bfloat16* xs = h2s + MB * SB_HS;
    float* red = reinterpret_cast<float*>(xs + MB * SB_XS);
    float* P1 = red + SB_WARPS * NT * 2 * 32 * 4;   // [NT][32][8] h1(t-1) . U1
    float* P2 = P1 + NT * 256;                      // [NT][32][8] h2(t-1) . U2
    float* G = P2 + NT * 256;                       // [NT][32][8] gate pre-activations
    float* bias = G + NT * 256;                     // [2][32]
    float* Ps = bias + 64;                     // [8 units][96] rows 8 cta .. 8 cta + 7 of the projection kernel (h2 part)
    float* hv_s = Ps + SB_UNITS * 96;          // [MB][8] h2(t) of this CTA's units (fp32)
    const uint4* wcta = q.wl + (size_t)cta * SB_TILES * 32;
    for (int i = tid; i < SB_T_U1 * 32; i += SB_THREADS) wres[i] = __ldg(wcta + i);
    for (int i = tid; i < MB * (2 * SB_HS + SB_XS) / 2; i += SB_THREADS) reinterpret_cast<uint32_t*>(h1s)[i] = 0u;
    if (tid < 64) bias[tid] = __ldg(q.bl + (size_t)cta * 64 + tid);
    for (int i = tid; i < SB_UNITS * 96; i += SB_THREADS) {
      const int un = i / 96, n = i - un * 96;
      Ps[i] = n < FA_PD ? __ldg(p.Wp + (size_t)(cta * SB_UNITS + un) * FA_PD + n) : 0.f;
    }
    pa_sync<SB_THREADS>();
    // initial hidden states ("step -1" = buffer 1 of p.h1 / p.h2) -> bf16 activation rows
    for (int i = tid; i < B * TC_U; i += SB_THREADS) {
      const int u = i / TC_U, k = i - u * TC_U;
      h1s[u * SB_HS + k] = __float2bfloat16(__ldcg(p.h1 + ((size_t)B + u) * TC_U + k));
      h2s[u * SB_HS + k] = __float2bfloat16(__ldcg(p.h2 + ((size_t)B + u) * TC_U + k));
    }
    // cell states of (utterance u = tid / 8, unit = tid % 8) in registers
    const int cu = tid >> 3, cn = tid & 7;
    const bool cell = tid < SB_UNITS * B;
    float c1 = cell ? __ldcg(p.c1 + (size_t)cu * TC_U + cta * SB_UNITS + cn) : 0.f;
    float c2 = cell ? __ldcg(p.c2 + (size_t)cu * TC_U + cta * SB_UNITS + cn) : 0.f;
    pa_sync<SB_THREADS>();
    for (int t = 0; t < T; ++t) {
      const int cur = t & 1, prv = cur ^ 1;
      // ---- off the critical path: the recurrent products
      if (t > 0) {
        if (tid == 0) v2_poll(&sy->h2cnt[0], (unsigned int)NL * (unsigned int)t);
        pa_sync<SB_THREADS>();
        for (int i = tid; i < B * (TC_U / 8); i += SB_THREADS) {
          const int u = i / (TC_U / 8), c = i - u * (TC_U / 8);
          *reinterpret_cast<uint4*>(h2s + u * SB_HS + 8 * c) = __ldcg(reinterpret_cast<const uint4*>(q.hbuf2 + ((size_t)prv * SB_MAXB + u) * TC_U) + c);
        }
        pa_sync<SB_THREADS>();
      }
      prof_tick(prof_s, 0);
      sb_matvec<true, NT>(P2, nullptr, bias + 32, wcta + (size_t)SB_T_U2 * 32, SB_KT_H, h2s, SB_HS, red, wid, lane);
      sb_matvec<true, NT>(P1, nullptr, bias, wcta + (size_t)SB_T_U1 * 32, SB_KT_H, h1s, SB_HS, red, wid, lane);
      prof_tick(prof_s, 1);
      // ---- LSTMCell 0: [p || ctx](t) . W1x (resident) + P1
      if (tid == 0) v2_poll(&sy->xcnt[0], (unsigned int)B * (unsigned int)(t + 1));
      pa_sync<SB_THREADS>();
      prof_tick(prof_s, 2);
      for (int i = tid; i < B * (TC_KX / 8); i += SB_THREADS) {
        const int u = i / (TC_KX / 8), c = i - u * (TC_KX / 8);
        *reinterpret_cast<uint4*>(xs + u * SB_XS + 8 * c) = __ldcg(reinterpret_cast<const uint4*>(q.xbuf + ((size_t)cur * SB_MAXB + u) * TC_KX) + c);
      }
      pa_sync<SB_THREADS>();
      sb_matvec<false, NT>(G, P1, nullptr, wres + (size_t)SB_T_W1X * 32, SB_KT_X, xs, SB_XS, red, wid, lane);
      if (cell) {   // rows: tile 0 = gates i (0-7), f (8-15); tile 1 = g (16-23), o (24-31)
        const float* Gc = G + (cu >> 3) * 256 + (cu & 7);
        const float zi = Gc[cn * 8], zf = Gc[(8 + cn) * 8], zg = Gc[(16 + cn) * 8], zo = Gc[(24 + cn) * 8];
        c1 = sigmoid_fast(zf) * c1 + sigmoid_fast(zi) * tanh_fast(zg);
        const float hv = sigmoid_fast(zo) * tanh_fast(c1);
        q.hbuf1[((size_t)cur * SB_MAXB + cu) * TC_U + cta * SB_UNITS + cn] = __float2bfloat16(hv);
        if (t == T - 1) p.h1[((size_t)cur * B + cu) * TC_U + cta * SB_UNITS + cn] = hv;
      }
      pa_sync<SB_THREADS>();
      if (tid == 0) v2_signal(&sy->h1cnt[0], 1u);
      prof_tick(prof_s, 3);
      // ---- LSTMCell 1: h1(t) . W2 (resident) + P2
      if (tid == 0) v2_poll(&sy->h1cnt[0], (unsigned int)NL * (unsigned int)(t + 1));
      pa_sync<SB_THREADS>();
      prof_tick(prof_s, 4);
      for (int i = tid; i < B * (TC_U / 8); i += SB_THREADS) {
        const int u = i / (TC_U / 8), c = i - u * (TC_U / 8);
        *reinterpret_cast<uint4*>(h1s + u * SB_HS + 8 * c) = __ldcg(reinterpret_cast<const uint4*>(q.hbuf1 + ((size_t)cur * SB_MAXB + u) * TC_U) + c);
      }
      pa_sync<SB_THREADS>();
      sb_matvec<false, NT>(G, P2, nullptr, wres + (size_t)SB_T_W2 * 32, SB_KT_H, h1s, SB_HS, red, wid, lane);
      if (cell) {
        const float* Gc = G + (cu >> 3) * 256 + (cu & 7);
        const float zi = Gc[cn * 8], zf = Gc[(8 + cn) * 8], zg = Gc[(16 + cn) * 8], zo = Gc[(24 + cn) * 8];
        c2 = sigmoid_fast(zf) * c2 + sigmoid_fast(zi) * tanh_fast(zg);
        const float hv = sigmoid_fast(zo) * tanh_fast(c2);
        q.hbuf2[((size_t)cur * SB_MAXB + cu) * TC_U + cta * SB_UNITS + cn] = __float2bfloat16(hv);
        hv_s[cu * SB_UNITS + cn] = hv;
        if (t == T - 1) p.h2[((size_t)cur * B + cu) * TC_U + cta * SB_UNITS + cn] = hv;
      }
      pa_sync<SB_THREADS>();
      // split-K partial of the projection of this step (Taco2.py:112-118): this CTA's 8 units of h2(t) against its 8 rows of P
      for (int i = tid; i < B * 96; i += SB_THREADS) {
        const int u = i / 96, n = i - u * 96;
        float a = 0.f;
#pragma unroll
        for (int un = 0; un < SB_UNITS; ++un) a = fmaf(hv_s[u * SB_UNITS + un], Ps[un * 96 + n], a);
        q.ppart[(((size_t)cur * SB_MAXB + u) * (TC_U / SB_UNITS) + cta) * 96 + n] = a;
      }
      pa_sync<SB_THREADS>();
      if (tid == 0) v2_signal(&sy->h2cnt[0], 1u);
      prof_tick(prof_s, 5);
    }
    if (cell) {
      p.c1[(size_t)cu * TC_U + cta * SB_UNITS + cn] = c1;
      p.c2[(size_t)cu * TC_U + cta * SB_UNITS + cn] = c2;
    }
  } else if (cta - NL < B) {
    // =========================================== front CTA: utterance b ===========================================
    // mma.sync with the utterance in batch column 0 (the other seven columns read the same activation row: their results are
    // ignored).  Weight fragments: prenet-1 and query RESIDENT in shared memory; prenet-0 and the first projection batch are
    // requested into registers BEFORE h2 arrives (weights do not depend on it); the rest of the projection streams from L2.
    // Projection: K split over the warps + one reduction; the other layers: features split over the warps, no reduction.
    const int b = cta - NL, Tv = p.Tv;
    const uint4* wres = reinterpret_cast<const uint4*>(sm);                                   // prenet-1 | query fragments
    __nv_bfloat16* act_mel = reinterpret_cast<__nv_bfloat16*>(sm + SB_FRES_BYTES);            // [96]
    __nv_bfloat16* act_z0 = act_mel + 96;                                                     // [256 + 8]
    __nv_bfloat16* act_z1 = act_z0 + FA_P + 8;                                                // [256 + 8]
    float* mel_s = reinterpret_cast<float*>(act_z1 + FA_P + 8);   // [96]
    float* q_s = mel_s + 96;           // [128]
    float* ctx_s = q_s + FA_A;         // [128] context of the previous step (fp32)
    float* keep_s = ctx_s + 128;       // [2][256] dropout scale (0 or 1/(1-rate)) of the two prenet layers for the coming step
    float* red = keep_s + 2 * FA_P;    // [SB_WARPS][96] projection partial sums
    float* pc_s = mel_s;               // [96] ctx(t-1) . P[1024:, :]  (computed right after the attention of step t-1)
    float* alig = red + SB_WARPS * 96; // [2][Tv] alignments by step parity | pbuf [Tv] | nzbuf [Tv] | ctxp [SB_WARPS][128]
    float* pbuf = alig + 2 * Tv;
    float* nzbuf = pbuf + Tv;
    float* ctxp = alig + ((4 * Tv + 3) & ~3);
    __shared__ __align__(16) float attv_s[128];
    const FaW L0 = FA_L0, L1 = FA_L1;
    {
      const uint4* src = reinterpret_cast<const uint4*>(q.wimgA + L1.base);
      uint4* dst = reinterpret_cast<uint4*>(sm);
      for (int i = tid; i < SB_FRES_BYTES / 16; i += SB_THREADS) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < SB_FACT_ELEMS / 2; i += SB_THREADS) reinterpret_cast<uint32_t*>(act_mel)[i] = 0u;
    if (tid < 128) { attv_s[tid] = __ldg(p.att_v + tid); ctx_s[tid] = 0.f; }
    if (tid < 96) pc_s[tid] = 0.f;
    for (int i = tid; i < Tv; i += SB_THREADS) alig[Tv + i] = __ldcg(p.align + ((size_t)B + b) * Tv + i);   // "step -1" = parity 1
    const bool noisy = p.rng_mode != 0 && p.sigmoid_noise > 0.f;
    const float sb_bias = __ldg(p.att_sb);
    auto draw = [&](int t) {   // dropout scales and attention noise of step t (off the critical path)
      for (int n = tid; n < 2 * FA_P; n += SB_THREADS) keep_s[n] = sb_keep(p, n / FA_P, t, b, n % FA_P);
      if (noisy) for (int j = tid; j < Tv; j += SB_THREADS) nzbuf[j] = att_noise(p, t, b, j);
    };
    draw(0);
    const int g = lane >> 2, t4 = lane & 3;
    // the projected keys of the utterance never change: they stay in registers for the whole decode.  Rows are spread evenly over
    // the warps (RW = 4 ceil(Tv / 32) rows each, NIT = RW / 4 iterations of 4 row groups); lane = (row group rg, 16-column slab cs)
    const int rg = lane >> 3, cs = lane & 7;
    const int NIT = (Tv + 4 * SB_WARPS - 1) / (4 * SB_WARPS), RW = 4 * NIT;
    uint4 kv[KVI][2];
    {
      const uint4* Vl = reinterpret_cast<const uint4*>(q.vproj_bf + (size_t)b * Tv * 128) + 2 * cs;
#pragma unroll
      for (int i = 0; i < KVI; ++i) {
        const int j = wid * RW + 4 * i + rg;
        if (i < NIT && j < Tv) { kv[i][0] = __ldg(Vl + (size_t)j * 16); kv[i][1] = __ldg(Vl + (size_t)j * 16 + 1); }
        else { kv[i][0] = make_uint4(0u, 0u, 0u, 0u); kv[i][1] = kv[i][0]; }
      }
    }
    // fragment addresses: projection tile (ft, kt) (fa_wlayer 0: stages of [6 ft][8 kt]), prenet-0 tile (ft, kt) (stages of [16 ft][4 | 1 kt])
    const uint4* w0 = reinterpret_cast<const uint4*>(q.wimgA + L0.base) + lane;
    auto pre0_tile = [&](int kt, int f) { return kt < 4 ? w0 + ((size_t)f * 4 + kt) * 32 : w0 + ((size_t)64 + f) * 32; };
    pa_sync<SB_THREADS>();
    for (int t = 0; t <= T; ++t) {
      const int cur = t & 1, prv = cur ^ 1;
      // requested before the wait: this warp's prenet-0 fragments (features 32 wid .. 32 wid + 31)
      uint4 a0[2][5];
      if (t < T) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
          for (int kt = 0; kt < 5; ++kt) a0[s2][kt] = sb_ldcg_pinned(pre0_tile(kt, 2 * wid + s2));
      }
      if (t > 0) {
        // ---- projection of step t-1: [h2(t-1) || ctx(t-1)] . P + bP  (Taco2.py:112-118).  The h2 part arrives as 128 split-K
        // partials from the LSTM CTAs (41 KB, one round trip), the ctx part (pc_s) was computed right after the attention of step t-1
        if (tid == 0) v2_poll(&sy->h2cnt[0], (unsigned int)NL * (unsigned int)t);
        pa_sync<SB_THREADS>();
        prof_tick(prof_s, 8);
        if (lane < 24) {   // warp w sums the partials of LSTM CTAs w, w + 8, ...: 16 independent 16 B loads per lane (4 outputs each)
          const float4* src = reinterpret_cast<const float4*>(q.ppart + ((size_t)prv * SB_MAXB + b) * (TC_U / SB_UNITS) * 96) + lane;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __ldcg(src + (size_t)(wid + i * SB_WARPS) * 24);
#pragma unroll
          for (int i = 0; i < 16; ++i) { acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w; }
          reinterpret_cast<float4*>(red + wid * 96)[lane] = acc;
        }
        pa_sync<SB_THREADS>();
        if (tid < FA_PD) {
          float v = pc_s[tid] + __ldg(p.bp + tid);
#pragma unroll
          for (int w2 = 0; w2 < SB_WARPS; ++w2) v += red[w2 * 96 + tid];
          if (tid < FA_PD - 1) {
            if (p.out_mel) p.out_mel[((size_t)b * p.To + (t - 1)) * (FA_PD - 1) + tid] = v;
            act_mel[tid] = __float2bfloat16(v);        // free running: the projected frame is the next decoder input (Taco2.py:183-187)
          } else {
            if (p.out_stop) p.out_stop[(size_t)b * p.To + (t - 1)] = v;
            note_stop(p, b, t - 1, v);
          }
        }
        prof_tick(prof_s, 9);
      } else if (tid < FA_MEL) {
        act_mel[tid] = __float2bfloat16(p.init_mel ? __ldg(p.init_mel + (size_t)b * FA_MEL + tid) : 0.f);
      }
      if (t == T) break;
      pa_sync<SB_THREADS>();
      // ---- prenet (Taco2.py:270-283, dropout always on) and query projection (Steps.py:122): warp = 32 (16) output features
      {
        float d[2][4] = {};
        const __nv_bfloat16* brow = act_mel + 2 * t4;
#pragma unroll
        for (int kt = 0; kt < 5; ++kt) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow + kt * 16), b1 = *reinterpret_cast<const uint32_t*>(brow + kt * 16 + 8);
          mma_16816_bf16(d[0], a0[0][kt], b0, b1);
          mma_16816_bf16(d[1], a0[1][kt], b0, b1);
        }
        if (t4 == 0)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
              const int n = (2 * wid + s2) * 16 + g + 8 * hi;
              act_z0[n] = __float2bfloat16(fmaxf(d[s2][2 * hi] + __ldg(p.b0 + n), 0.f) * keep_s[n]);
            }
      }
      pa_sync<SB_THREADS>();
      prof_tick(prof_s, 12);
      {
        float d[2][4] = {};
        const __nv_bfloat16* brow = act_z0 + 2 * t4;
        const uint4* w1 = wres + lane;   // fa_wlayer 2: 4 stages of [16 ft][4 kt]
#pragma unroll
        for (int kt = 0; kt < 16; ++kt) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow + kt * 16), b1 = *reinterpret_cast<const uint32_t*>(brow + kt * 16 + 8);
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) mma_16816_bf16(d[s2], w1[(size_t)(((kt >> 2) * 16 + 2 * wid + s2) * 4 + (kt & 3)) * 32], b0, b1);
        }
        if (t4 == 0)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int hi = 0; hi < 2; ++hi) {
              const int n = (2 * wid + s2) * 16 + g + 8 * hi;
              const __nv_bfloat16 v = __float2bfloat16(fmaxf(d[s2][2 * hi] + __ldg(p.b1 + n), 0.f) * keep_s[FA_P + n]);
              act_z1[n] = v;
              q.xbuf[((size_t)cur * SB_MAXB + b) * TC_KX + n] = v;   // p(t): LSTMCell-0 input
            }
      }
      pa_sync<SB_THREADS>();
      prof_tick(prof_s, 13);
      {
        float d[4] = {};
        const __nv_bfloat16* brow = act_z1 + 2 * t4;
        const uint4* wq = wres + V2_WRES_LQ / 16 + lane;   // fa_wlayer 3: 2 stages of [8 ft][8 kt]
#pragma unroll
        for (int kt = 0; kt < 16; ++kt) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow + kt * 16), b1 = *reinterpret_cast<const uint32_t*>(brow + kt * 16 + 8);
          mma_16816_bf16(d, wq[(size_t)(((kt >> 3) * 8 + wid) * 8 + (kt & 7)) * 32], b0, b1);
        }
        if (t4 == 0) {
          q_s[wid * 16 + g] = d[0] + __ldg(p.bq + wid * 16 + g);
          q_s[wid * 16 + 8 + g] = d[2] + __ldg(p.bq + wid * 16 + 8 + g);
        }
      }
      pa_sync<SB_THREADS>();
      prof_tick(prof_s, 10);
      // ---- stepwise monotonic attention of this utterance (Steps.py:138-166, 215-229)
      {
        float qr[16], vr[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { qr[c] = q_s[16 * cs + c]; vr[c] = attv_s[16 * cs + c]; }
        float e[KVI];
#pragma unroll
        for (int i = 0; i < KVI; ++i) {
          e[i] = 0.f;
          if (i < NIT) {
            const uint32_t w[8] = {kv[i][0].x, kv[i][0].y, kv[i][0].z, kv[i][0].w, kv[i][1].x, kv[i][1].y, kv[i][1].z, kv[i][1].w};
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              s0 = fmaf(vr[2 * c], tanh_fast(qr[2 * c] + bf16lo(w[c])), s0);
              s1 = fmaf(vr[2 * c + 1], tanh_fast(qr[2 * c + 1] + bf16hi(w[c])), s1);
            }
            e[i] = s0 + s1;
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
#pragma unroll
          for (int i = 0; i < KVI; ++i) e[i] += __shfl_xor_sync(0xffffffffu, e[i], o);
        float er = e[0];
#pragma unroll
        for (int i = 1; i < KVI; ++i) er = cs == i ? e[i] : er;
        {   // lane cs == i finishes row 4 i + rg of the warp
          const int j = wid * RW + 4 * cs + rg;
          if (cs < NIT && j < Tv) {
            er += sb_bias;
            if (noisy) er = fmaf(p.sigmoid_noise, nzbuf[j], er);
            pbuf[j] = sigmoid_fast(er);
          }
        }
        pa_sync<SB_THREADS>();
        prof_tick(prof_s, 14);
        for (int j = tid; j < Tv; j += SB_THREADS) {   // alignment recurrence (Steps.py:223-229)
          const float* prev_s = alig + prv * Tv;
          float a = prev_s[j] * pbuf[j];
          if (j > 0) a = fmaf(prev_s[j - 1], 1.0f - pbuf[j - 1], a);
          alig[cur * Tv + j] = a;
          if (p.out_align) p.out_align[((size_t)b * p.To + t) * Tv + j] = a;
          if (t == T - 1) p.align[((size_t)cur * B + b) * Tv + j] = a;
        }
        pa_sync<SB_THREADS>();
        prof_tick(prof_s, 15);
        float cx[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) cx[c] = 0.f;
#pragma unroll
        for (int i = 0; i < KVI; ++i) {
          if (i < NIT) {
            const int j = wid * RW + 4 * i + rg;
            const float a = j < Tv ? alig[cur * Tv + j] : 0.f;
            const uint32_t w[8] = {kv[i][0].x, kv[i][0].y, kv[i][0].z, kv[i][0].w, kv[i][1].x, kv[i][1].y, kv[i][1].z, kv[i][1].w};
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              cx[2 * c] = fmaf(a, bf16lo(w[c]), cx[2 * c]);
              cx[2 * c + 1] = fmaf(a, bf16hi(w[c]), cx[2 * c + 1]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          cx[c] += __shfl_xor_sync(0xffffffffu, cx[c], 8);
          cx[c] += __shfl_xor_sync(0xffffffffu, cx[c], 16);
        }
        if (rg == 0) {
          float4* dst = reinterpret_cast<float4*>(ctxp + wid * 128 + 16 * cs);
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = make_float4(cx[4 * c], cx[4 * c + 1], cx[4 * c + 2], cx[4 * c + 3]);
        }
        pa_sync<SB_THREADS>();
        if (tid < 128) {
          float c = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < SB_WARPS; ++w2) c += ctxp[w2 * 128 + tid];
          ctx_s[tid] = c;
          q.xbuf[((size_t)cur * SB_MAXB + b) * TC_KX + FA_P + tid] = __float2bfloat16(c);
          if (t == T - 1) {
            p.xin[(size_t)b * (p.P1 + p.A) + p.P1 + tid] = c;
            if (p.out_ctx) p.out_ctx[(size_t)b * p.A + tid] = c;
          }
        }
      }
      pa_sync<SB_THREADS>();
      if (tid == 0) v2_signal(&sy->xcnt[0], 1u);
      prof_tick(prof_s, 11);
      // off the critical path (the LSTM phases of this step run now): draws of step t+1 and the ctx part of this step's projection
      if (t + 1 < T) draw(t + 1);
      if (tid < FA_PD) {
        float a = 0.f;
        const float* Pc = p.Wp + (size_t)TC_U * FA_PD + tid;
#pragma unroll 8
        for (int k = 0; k < 128; ++k) a = fmaf(ctx_s[k], __ldg(Pc + (size_t)k * FA_PD), a);
        pc_s[tid] = a;
      }
      pa_sync<SB_THREADS>();
    }
  }
  __syncthreads();
  if (q.prof && tid == 0)
    for (int i = 0; i < PROF_SLOTS; ++i) q.prof[(size_t)cta * PROF_SLOTS + i] = prof_sh[i];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SbState {
  uint4* wl = nullptr;
  float* bl = nullptr;
  __nv_bfloat16* bufs = nullptr;   // hbuf1 | hbuf2 | xbuf
  float* ppart = nullptr;
  SbSync* sync = nullptr;
  bool ready = false;
};
inline void sb_release(SbState& s) {
  cudaFree(s.wl);
  cudaFree(s.bl);
  cudaFree(s.bufs);
  cudaFree(s.ppart);
  cudaFree(s.sync);
  s = SbState();
}
inline int sb_prepare(SbState& s, const std::map<std::string, std::vector<float>>& hw, std::string& err) {
  if (s.ready) return GSTK_OK;
  const std::string d = "Decoder/Decoder_Step/RNN/";
  const std::vector<float>* mats[4] = {&hw.at(d + "cell_0/kernel"), &hw.at(d + "cell_1/kernel"), &hw.at(d + "cell_0/recurrent_kernel"),
                                       &hw.at(d + "cell_1/recurrent_kernel")};   // image order: W1x, W2, U1, U2
  const int NKT[4] = {SB_KT_X, SB_KT_H, SB_KT_H, SB_KT_H}, T0[4] = {SB_T_W1X, SB_T_W2, SB_T_U1, SB_T_U2};
  const std::vector<float>& b0 = hw.at(d + "cell_0/bias");
  const std::vector<float>& b1 = hw.at(d + "cell_1/bias");
  const int NL = TC_U / SB_UNITS;
  std::vector<__nv_bfloat16> img((size_t)NL * SB_TILES * 256);
  std::vector<float> bias((size_t)NL * 64);
  for (int c = 0; c < NL; ++c) {
    for (int mi = 0; mi < 4; ++mi) {
      const std::vector<float>& W = *mats[mi];   // [K][4 U], Keras gate order i, f, c, o
      // A(row r of tile m, column k): gate = 2 m + r / 8, unit = 8 c + r % 8
      auto at = [&](int m, int r, int k) { return __float2bfloat16(W[(size_t)k * 4 * TC_U + (size_t)(2 * m + r / 8) * TC_U + c * SB_UNITS + r % 8]); };
      for (int m = 0; m < 2; ++m)
        for (int kt = 0; kt < NKT[mi]; ++kt)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3, k0 = kt * 16;
            __nv_bfloat16* o = img.data() + ((size_t)c * SB_TILES + T0[mi] + m * NKT[mi] + kt) * 256 + (size_t)lane * 8;
            o[0] = at(m, g, k0 + 2 * t);         o[1] = at(m, g, k0 + 2 * t + 1);
            o[2] = at(m, g + 8, k0 + 2 * t);     o[3] = at(m, g + 8, k0 + 2 * t + 1);
            o[4] = at(m, g, k0 + 2 * t + 8);     o[5] = at(m, g, k0 + 2 * t + 9);
            o[6] = at(m, g + 8, k0 + 2 * t + 8); o[7] = at(m, g + 8, k0 + 2 * t + 9);
          }
    }
    for (int row = 0; row < 32; ++row) {
      bias[(size_t)c * 64 + row] = b0[(size_t)(row / 8) * TC_U + c * SB_UNITS + row % 8];
      bias[(size_t)c * 64 + 32 + row] = b1[(size_t)(row / 8) * TC_U + c * SB_UNITS + row % 8];
    }
  }
  auto fail = [&](const char* m) { err = m; return GSTK_ECUDA; };
  if (cudaMalloc((void**)&s.wl, img.size() * 2) != cudaSuccess) return fail("cudaMalloc(sb weights) failed");
  if (cudaMalloc((void**)&s.bl, bias.size() * 4) != cudaSuccess) return fail("cudaMalloc(sb bias) failed");
  if (cudaMalloc((void**)&s.bufs, (size_t)2 * SB_MAXB * (2 * TC_U + TC_KX) * 2) != cudaSuccess) return fail("cudaMalloc(sb buffers) failed");
  if (cudaMalloc((void**)&s.ppart, (size_t)2 * SB_MAXB * (TC_U / SB_UNITS) * 96 * 4) != cudaSuccess) return fail("cudaMalloc(sb partials) failed");
  if (cudaMalloc((void**)&s.sync, sizeof(SbSync)) != cudaSuccess) return fail("cudaMalloc(sb sync) failed");
  if (cudaMemcpy(s.wl, img.data(), img.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  if (cudaMemcpy(s.bl, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  s.ready = true;
  return GSTK_OK;
}
inline size_t sb_smem_bytes(const DecParams& p) {
  const size_t f = sb_front_bytes(p.Tv);
  const size_t l = sb_lstm_bytes(p.B > 8 ? 2 : 1);
  return 128 + (f > l ? f : l);
}
inline bool sb_usable(const Bf16State& st, const DecParams& p, int num_sms) {
  return st.fast_a && p.mode == 0 && p.T >= 1 && p.B <= SB_MAXB && p.Tv <= 32 * SB_WARPS && num_sms >= TC_U / SB_UNITS + p.B &&
         sb_smem_bytes(p) <= 227 * 1024;   // key_time <= 256: the keys of an utterance stay in its front CTA's registers
}
inline int sb_decode(Bf16State& st, SbState& s, DecParams& p, int num_sms, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1, int64_t& launches,
                     std::string& err) {
  auto fail = [&](int code, const std::string& m) { err = m; return code; };
  cudaError_t e;
  SbParams q;
  q.wl = s.wl;
  q.bl = s.bl;
  q.wimgA = st.wimgA;
  q.hbuf1 = s.bufs;
  q.hbuf2 = q.hbuf1 + (size_t)2 * SB_MAXB * TC_U;
  q.xbuf = q.hbuf2 + (size_t)2 * SB_MAXB * TC_U;
  q.ppart = s.ppart;
  q.sync = s.sync;
  p.MT = 1;
  p.actX = nullptr;
  const size_t nv = (size_t)p.B * p.Tv * 128;
  if (st.vproj_elems < nv) {
    cudaFree(st.vproj_bf);
    st.vproj_bf = nullptr;
    if ((e = cudaMalloc((void**)&st.vproj_bf, nv * 2)) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
    st.vproj_elems = nv;
  }
  f32_to_bf16_kernel<<<num_sms * 2, 256, 0, stream>>>(p.vproj, st.vproj_bf, nv);
  launches += 1;
  q.vproj_bf = st.vproj_bf;
  if (!st.prof) {
    if ((e = cudaMalloc((void**)&st.prof, (size_t)num_sms * PROF_SLOTS * sizeof(unsigned long long))) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
    st.prof_ctas = num_sms;
  }
  q.prof = (p.debug_flags & 8) ? st.prof : nullptr;
  if ((e = cudaMemsetAsync(s.sync, 0, sizeof(SbSync), stream)) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
  if ((e = cudaMemsetAsync(s.bufs, 0, (size_t)2 * SB_MAXB * (2 * TC_U + TC_KX) * 2, stream)) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
  const size_t smem = sb_smem_bytes(p);
  void* kern = p.B > 8 ? (p.Tv <= 96 ? (void*)decoder_bf16_sb_kernel<3, 2> : p.Tv <= 160 ? (void*)decoder_bf16_sb_kernel<5, 2> : (void*)decoder_bf16_sb_kernel<8, 2>)
                       : (p.Tv <= 96 ? (void*)decoder_bf16_sb_kernel<3, 1> : p.Tv <= 160 ? (void*)decoder_bf16_sb_kernel<5, 1> : (void*)decoder_bf16_sb_kernel<8, 1>);
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  const int grid = TC_U / SB_UNITS + p.B;
  void* args[] = {&p, &q};
  cudaEventRecord(ev0, stream);
  if ((e = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(SB_THREADS), args, smem, stream)) != cudaSuccess)
    return fail(GSTK_ECUDA, std::string("cooperative launch failed: ") + cudaGetErrorString(e));
  cudaEventRecord(ev1, stream);
  launches += 1;
  return GSTK_OK;
}

}  // namespace gstk
