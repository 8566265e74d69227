// Persistent bf16 decoder, DATAFLOW version ("v2"): the free-running fast path (SMA, default widths, Step_Reduction 1).
//
// Same tiling of the two LSTMCells as decoder_bf16.cuh (128 LSTM CTAs = batch m-tile x 16 hidden units of both cells, tcgen05
// M=128 x N=64 products out of a bulk-copy operand ring, accumulators + cell state in TMEM) - what changes is everything
// BETWEEN the LSTM phases (Modules/Taco2.py:106-118):
//
//   * no grid barrier at all.  Every producer -> consumer edge of the step is a monotonic arrival counter in global memory
//     (V2Sync): h2 k-blocks -> fold / U2 streams, z0 slices -> prenet-1, p / ctx -> LSTMCell-0 stream, h1 k-blocks -> LSTMCell-1
//     stream.  Only the 128 LSTM CTAs work; the two batch m-tiles are independent pipelines coupled by the attention CTAs.
//   * the projection -> prenet-0 chain is folded OFFLINE into one matrix (free-running mode feeds the projected frame straight
//     back, Taco2.py:183-187, and Step_Reduction is 1):  relu(W0^T (P^T [h2||ctx] + bP) + b0) = relu(Wf^T [h2||ctx] + bf)  with
//     Wf = P[:, :80] W0  [1152 x 256].  That GEMM (and the 81-column projection itself, now off the critical path) runs on the
//     TENSOR CORES of 22 LSTM CTAs per m-tile as an extra operand stream: M=64 (half an m-tile) x N=32 output columns x K=1152.
//   * prenet-1 (256->256) and the query projection (256->128) run on the CTA that also runs the utterance's attention
//     (utterances c and c + 128): mma.sync with the weights as A fragments straight from L2 (192 KB per CTA and step) - no
//     dense CTAs, no query hand-over, the queries never leave shared memory.
//   * the operand ring is fed by TWO warps (weights | activations): a lone copy warp needs ~500 clk of dependent issue work
//     per unit, which bounded every stream (measured at batch 1, where no stream carries data worth mentioning).
//
// Hazard analysis (why single buffers are safe where they are single): h1 and ctx are double-buffered by
// step parity; h2, p and z0 are single buffers whose next write transitively depends on every reader of the old value.
#pragma once
#include "decoder_bf16.cuh"

namespace gstk {

constexpr int V2_THREADS = 416;                               // 13 warps: 0-9 phase A / epilogues, 10 weight copies, 11 MMA, 12 activation copies
constexpr int V2_NSTAGE = 7;
constexpr int V2_STAGE_W = TC_A_BYTES;                        // weight block offset inside a stage
constexpr int V2_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;       // 24 KB: one activation tile + one 64-row weight block
constexpr int V2_FOLD_N = 32;                                 // output columns of one fold slice
constexpr int V2_FOLD_ZSLICES = FA_P / V2_FOLD_N;             // 8 slices of prenet-0 features
constexpr int V2_FOLD_SLICES = V2_FOLD_ZSLICES + 3;           // + 96 >= 81 projection columns
constexpr int V2_FOLD_NKB = TC_NKB_H + 2;                     // 16 k-blocks of h2 + 2 of ctx
constexpr int V2_FOLD_B_BYTES = V2_FOLD_N * 128;              // 4 KB
constexpr int V2_FOLD_IMG_BYTES = V2_FOLD_NKB * V2_FOLD_B_BYTES;
constexpr int V2_FOLD_CTAS = 2 * V2_FOLD_SLICES;              // unit groups 0..21 of an m-tile: (slice, row half)
constexpr int V2_FOLD_UNITS = TC_NKB_H / 2;                   // fold units of two h2 k-blocks (+ one unit of the two ctx k-blocks)
constexpr uint32_t V2_D2 = 0, V2_D1 = 64, V2_C1 = 128, V2_C2 = 144, V2_Z = 160;
constexpr int V2_ACT_STRIDE = FA_P + 8;                       // bf16 row stride of the prenet layer inputs (4 mod 32 words)
constexpr int V2_WRES_LQ = (int)(FA_L1.nst() * FA_L1.stride());   // byte offset of the query kernel behind the prenet-1 kernel
constexpr int V2_ZS_ELEMS = 8 * V2_ACT_STRIDE;                // one [8 utterance slots][264] bf16 layer-input buffer
constexpr int V2_DENSE_SCRATCH = 2 * V2_ZS_ELEMS * 2 + 2 * FA_P;   // z0s | z1s | keep flags of prenet layer 1 (bytes)

// arrival counters (one 128 B line each); zeroed by the host before every launch
struct V2Sync {
  unsigned int kb_h1[2 * TC_NKB_H][32];   // [m-tile][k-block]: CTAs that have published their 16 units of h1(t)   (4 per step)
  unsigned int kb_h2[2 * TC_NKB_H][32];   // same for h2(t)
  unsigned int zcnt[2][32];               // [m-tile]: fold CTAs that have published their slice of z0(t)
  unsigned int pcnt[2][32];               // [m-tile]: batch rows whose p(t) (prenet output) is in the operand image
  unsigned int ctxcnt[2][32];             // [m-tile]: warps (4 per batch row) that have stored their part of ctx(t)
};

struct V2Params {
  const __nv_bfloat16* wimg;      // LSTM weight blocks (layout of decoder_bf16.cuh)
  const float* bias;              // [TC_UG][2][64]
  const uint8_t* fold_img;        // [V2_FOLD_SLICES][V2_FOLD_NKB][32 rows x 128 B] SWIZZLE_128B blocks of [Wf | P]
  const float* fold_bias;         // [V2_FOLD_SLICES][32]
  const uint8_t* wres;            // prenet-1 | query kernels in mma.sync fragment order (decoder_bf16.cuh: fa_wlayer 2, 3)
  __nv_bfloat16* actP;            // [4][MT][128][64]        p(t)
  __nv_bfloat16* actC;            // [2][2][MT][128][64]     ctx, by step parity
  __nv_bfloat16* actH1;           // [2][16][MT][128][64]    h1, by step parity
  __nv_bfloat16* actH2;           // [16][MT][128][64]
  __nv_bfloat16* z0buf;           // [B][256] prenet-0 output (after ReLU + dropout)
  const __nv_bfloat16* vproj_bf;  // [B][Tv][128]
  V2Sync* sync;
  unsigned long long* prof;
};

// ---- global-flag waits (bounded: a protocol bug traps instead of hanging the GPU) ------------------------------------
__device__ __forceinline__ void v2_poll(const unsigned int* cnt, unsigned int want) {
  if (ld_relaxed_u32(cnt) >= want) return;
  const long long t0 = clock64();
  while (ld_relaxed_u32(cnt) < want)
    if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void v2_signal(unsigned int* cnt, unsigned int inc) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(inc) : "memory");
}

// ---- operand ring -------------------------------------------------------------------------------------------------------
// One segment = NU pipeline units; a unit = activation tile(s) + weight block(s) of one k-block (rotated by `rot`).  The ring
// position carries over from segment to segment (round robin).  Three warps walk the same unit sequence, each with its own
// cursor: the WEIGHT warp arms the stage's full barrier (expect_tx of the whole unit) and copies the weight block - weights are
// constants, so it runs up to V2_NSTAGE units ahead; the ACTIVATION warp waits for the unit's gate (arrival counters of the
// tile's writers) and copies the tile; the MMA warp consumes.  A complete_tx that lands before the stage is armed only drives
// the transaction count negative - the phase cannot complete before the arming arrive.
template <int NU_>
__device__ __noinline__ TcRing v2_prod_w(TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* wsrc, uint32_t wstride, uint32_t wbytes, uint32_t total,
                                         int rot) {
  uint64_t* empty = full + V2_NSTAGE;
  int kb = rot;
  const uint8_t* w = wsrc + (size_t)kb * wstride;
#pragma unroll 1
  for (int i = 0; i < NU_; ++i) {
    mbar_wait(&empty[r.stage], r.phase() ^ 1u);
    if (elect_one()) {
      mbar_arrive_expect_tx(&full[r.stage], total);
      bulk_g2s(stages + (size_t)r.stage * V2_STAGE_BYTES + V2_STAGE_W, w, wbytes, &full[r.stage]);
    }
    __syncwarp();
    r.template advance<V2_NSTAGE>();
    if (++kb == NU_) { kb = 0; w = wsrc; }
    else w += wstride;
  }
  return r;
}
//   GATE 0: none;  GATE 1: one counter (gate[0] >= want);  GATE 2: one counter per k-block (gate[kb * 32] >= want).
//   pf (diagnostics): pf[0] += ticks waiting for a free stage, pf[1] += ticks polling the gate
template <int NU_, int GATE>
__device__ __noinline__ TcRing v2_prod_a(TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* act, uint32_t astride, uint32_t abytes, int rot,
                                         const unsigned int* gate, unsigned int want, unsigned long long* pf) {
  static_assert(NU_ <= 32, "one lane per k-block");
  uint64_t* empty = full + V2_NSTAGE;
  const int lane = threadIdx.x & 31;
  uint32_t ready = GATE == 0 ? 0xffffffffu : 0u;
  int kb = rot;
  const uint8_t* a = act + (size_t)kb * astride;
#pragma unroll 1
  for (int i = 0; i < NU_; ++i) {
    const long long te = pf ? clock64() : 0;
    mbar_wait(&empty[r.stage], r.phase() ^ 1u);
    if (pf && lane == 0) pf[0] += (unsigned long long)(clock64() - te);
    if (GATE != 0 && !((ready >> kb) & 1u)) {
      const long long t0 = clock64();
      for (;;) {
        bool ok;
        if (GATE == 1) ok = ld_relaxed_u32(gate) >= want;
        else ok = lane < NU_ ? ld_relaxed_u32(gate + lane * 32) >= want : true;
        ready = GATE == 1 ? (__all_sync(0xffffffffu, ok) ? 0xffffffffu : 0u) : __ballot_sync(0xffffffffu, ok);
        if ((ready >> kb) & 1u) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      fence_proxy_async();   // the tile was written with generic-proxy stores by other CTAs; the copy engine reads it
      if (pf && lane == 0) pf[1] += (unsigned long long)(clock64() - t0);
    }
    if (elect_one()) bulk_g2s(stages + (size_t)r.stage * V2_STAGE_BYTES, a, abytes, &full[r.stage]);
    __syncwarp();
    r.template advance<V2_NSTAGE>();
    if (++kb == NU_) { kb = 0; a = act; }
    else a += astride;
  }
  return r;
}
// fold stream, activation side: units of TWO k-blocks (the MMA warp needs ~250 clk of issue work per unit whatever its size, and a
// fold k-block is only 12 KB): stage = [A half-tile kb | A half-tile kb+1 | W kb (4 KB) | W kb+1 (4 KB)].  Units 0..7 = h2 k-block
// pairs (rotated, gated by both h2 counters), unit 8 = the two ctx k-blocks (complete by construction).
__device__ __noinline__ TcRing v2_prod_a_fold(TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* h2_half, const uint8_t* ctx_half, uint32_t astride,
                                              uint32_t hbytes, int rot, const unsigned int* kb_h2, unsigned int want, unsigned long long* pf) {
  uint64_t* empty = full + V2_NSTAGE;
  const int lane = threadIdx.x & 31;
  uint32_t ready = 0u;
  int u = rot;
#pragma unroll 1
  for (int i = 0; i < V2_FOLD_UNITS + 1; ++i) {
    const bool ctx = i == V2_FOLD_UNITS;
    const int kb = 2 * u;
    uint8_t* st = stages + (size_t)r.stage * V2_STAGE_BYTES;
    const long long te = pf ? clock64() : 0;
    mbar_wait(&empty[r.stage], r.phase() ^ 1u);
    if (pf && lane == 0) pf[0] += (unsigned long long)(clock64() - te);
    if (!ctx && ((ready >> kb) & 3u) != 3u) {
      const long long t0 = clock64();
      for (;;) {
        const bool ok = lane < TC_NKB_H ? ld_relaxed_u32(kb_h2 + lane * 32) >= want : true;
        ready = __ballot_sync(0xffffffffu, ok);
        if (((ready >> kb) & 3u) == 3u) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      fence_proxy_async();
      if (pf && lane == 0) pf[1] += (unsigned long long)(clock64() - t0);
    }
    if (elect_one()) {
      const uint8_t* a = ctx ? ctx_half : h2_half + (size_t)kb * astride;
      bulk_g2s(st, a, hbytes, &full[r.stage]);
      bulk_g2s(st + 8192, a + astride, hbytes, &full[r.stage]);
    }
    __syncwarp();
    r.template advance<V2_NSTAGE>();
    if (++u == V2_FOLD_UNITS) u = 0;
  }
  return r;
}

// MMA warp: D[M x N] (+)= A[M x 64] . B[N x 64]^T per unit (4 x K=16), B = the block at stage offset V2_STAGE_W
template <int NU_, bool FRESH, int M, int N>
__device__ __noinline__ TcRing v2_seg_consume(TcRing r, uint64_t* full, uint32_t stages_sa, uint32_t tmem_d, uint64_t* commit_done) {
  uint64_t* empty = full + V2_NSTAGE;
  constexpr uint32_t idesc = make_idesc_bf16(M, N);
#pragma unroll 1
  for (int i = 0; i < NU_; ++i) {
    const uint32_t acc = (FRESH && i == 0) ? 0u : 1u;
    mbar_wait(&full[r.stage], r.phase());
    tc_fence_after();
    const uint32_t st_sa = stages_sa + r.stage * (uint32_t)V2_STAGE_BYTES;
    const uint32_t ad = tc_desc_lo(st_sa), bd = ad + (V2_STAGE_W >> 4);
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (k > 0) ? 1u : acc);
      if (commit_done && i == NU_ - 1) umma_commit(commit_done);
      umma_commit(&empty[r.stage]);
    }
    __syncwarp();
    r.template advance<V2_NSTAGE>();
  }
  return r;
}
__device__ __noinline__ TcRing v2_fold_consume(TcRing r, uint64_t* full, uint32_t stages_sa, uint32_t tmem_d, uint64_t* commit_done, unsigned long long* pf) {
  uint64_t* empty = full + V2_NSTAGE;
  constexpr uint32_t idesc = make_idesc_bf16(64, V2_FOLD_N);
#pragma unroll 1
  for (int i = 0; i < V2_FOLD_UNITS + 1; ++i) {
    const long long tf = pf ? clock64() : 0;
    mbar_wait(&full[r.stage], r.phase());
    if (pf && (threadIdx.x & 31) == 0) pf[0] += (unsigned long long)(clock64() - tf);
    tc_fence_after();
    const uint32_t ad = tc_desc_lo(stages_sa + r.stage * (uint32_t)V2_STAGE_BYTES), bd = ad + (V2_STAGE_W >> 4);
    if (elect_one()) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_lo(tmem_d, ad + j * (8192 >> 4) + 2 * k, bd + j * (V2_FOLD_B_BYTES >> 4) + 2 * k, idesc, (i | j | k) ? 1u : 0u);
      if (i == V2_FOLD_UNITS) umma_commit(commit_done);
      umma_commit(&empty[r.stage]);
    }
    __syncwarp();
    r.template advance<V2_NSTAGE>();
  }
  return r;
}

// ---- fold epilogue (warps 0-3 of a fold CTA): z0 = dropout(relu(acc + bf)) -> z0buf, or projection columns -> out_mel / out_stop
// An M = 64 accumulator occupies TMEM lanes 0-15 of every lane quarter: warp w, lane l < 16 owns row 16 w + l of the CTA's
// 64-row half (measured: tools/ubench_m64.cu).  `t` = the decoder step this z0 feeds; the projected frame is that of step t-1.
__device__ __noinline__ void v2_fold_epilogue(const DecParams& p, const V2Params& q, uint64_t* z_full, uint32_t parity, uint32_t t_z,
                                              const float* fbias_s, int slice, int row, bool row_ok, int t) {
  const bool zslice = slice < V2_FOLD_ZSLICES;
  const bool drop = p.rng_mode != 0 && p.drop_rate > 0.f;
  uint32_t keep = 0xffffffffu;
  if (zslice && drop && row_ok && t < p.T) {   // drawn before the wait: the Philox rounds hide behind the operand stream
    if (p.rng_mode == 1) {
      const float4* kp = reinterpret_cast<const float4*>(p.keep0 + ((size_t)t * p.rngB + p.rng_b0 + row) * FA_P + slice * V2_FOLD_N);
      keep = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 m = __ldg(kp + i);
        keep |= (m.x != 0.f ? 1u : 0u) << (4 * i) | (m.y != 0.f ? 2u : 0u) << (4 * i) | (m.z != 0.f ? 4u : 0u) << (4 * i) | (m.w != 0.f ? 8u : 0u) << (4 * i);
      }
    } else {
      keep = 0u;
      const float sc = 5.9604644775390625e-08f;
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const uint4 rr = philox4x32_10(make_uint4((unsigned int)(slice * 8 + i), p.step_offset + (unsigned int)t, p.row_offset + (unsigned int)row,
                                                  (unsigned int)STREAM_KEEP0), make_uint2((unsigned int)p.seed, (unsigned int)(p.seed >> 32)));
        keep |= (((float)(rr.x >> 8) * sc >= p.drop_rate ? 1u : 0u) | ((float)(rr.y >> 8) * sc >= p.drop_rate ? 2u : 0u) |
                 ((float)(rr.z >> 8) * sc >= p.drop_rate ? 4u : 0u) | ((float)(rr.w >> 8) * sc >= p.drop_rate ? 8u : 0u)) << (4 * i);
      }
    }
  }
  mbar_wait_backoff(z_full, parity);
  tc_fence_after();
  float v[32];
  tmem_ld32(t_z, v);
  if (row_ok) {
    if (zslice) {
      if (t < p.T) {
        const float dscale = p.drop_scale;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = fmaxf(v[2 * i] + fbias_s[2 * i], 0.f), b = fmaxf(v[2 * i + 1] + fbias_s[2 * i + 1], 0.f);
          if (drop) {
            a = ((keep >> (2 * i)) & 1u) ? a * dscale : 0.f;
            b = ((keep >> (2 * i + 1)) & 1u) ? b * dscale : 0.f;
          }
          const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(q.z0buf + (size_t)row * FA_P + slice * V2_FOLD_N);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
    } else {
      // projection of step t-1 (Taco2.py:112-118): mel frame + stop logit
      const int c0 = (slice - V2_FOLD_ZSLICES) * V2_FOLD_N;
      if (p.out_mel) {
        float* dst = p.out_mel + ((size_t)row * p.To + (t - 1)) * (FA_PD - 1) + c0;
        if (c0 + V2_FOLD_N <= FA_PD - 1 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i] + fbias_s[4 * i], v[4 * i + 1] + fbias_s[4 * i + 1], v[4 * i + 2] + fbias_s[4 * i + 2],
                                                            v[4 * i + 3] + fbias_s[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < V2_FOLD_N; ++i)
            if (c0 + i < FA_PD - 1) dst[i] = v[i] + fbias_s[i];
        }
      }
      if (c0 <= FA_PD - 1 && FA_PD - 1 < c0 + V2_FOLD_N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V2_FOLD_N; ++i) s = (c0 + i == FA_PD - 1) ? v[i] + fbias_s[i] : s;
        if (p.out_stop) p.out_stop[(size_t)row * p.To + (t - 1)] = s;
        note_stop(p, row, t - 1, s);
      }
    }
  }
  tc_fence_before();
}


// ---- prenet layer 1 + query projection (Taco2.py:270-283, Steps.py:122) for the CTA's own utterances b0 and b0 + 128 ---------
// mma.sync m16n8k16 with the weights as A fragments (16 features x 16 k, fragment-ordered image of decoder_bf16.cuh, read
// straight from L2: 16 B per lane and tile) and the utterances as the N columns (2 of 8 used).  Warps 0-7: two feature tiles of
// prenet-1 each, then one feature tile of the query layer each; warps 8-9 store p(t) into the LSTMCell-0 operand image.
// scratch: ... | qs [2][128] fp32 (queries) | ... behind the attention scratch: z0s | z1s [8][264] bf16 | keep1 [2][256] bytes
template <int NU>
__device__ __forceinline__ void v2_keep1_fill(const DecParams& p, uint8_t* keep_s, int b0, int t) {
  if (!(p.rng_mode != 0 && p.drop_rate > 0.f)) return;
  uint32_t* keep = reinterpret_cast<uint32_t*>(keep_s);
  const unsigned int step_id = p.step_offset + (unsigned int)t;
  for (int i = threadIdx.x; i < NU * (FA_P / 4); i += TC_PA_THREADS) {
    const int u = i / (FA_P / 4), n4 = (i - u * (FA_P / 4)) * 4, b = b0 + u * 128;
    uint32_t f;
    if (p.rng_mode == 1) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(p.keep1 + ((size_t)t * p.rngB + p.rng_b0 + b) * FA_P + n4));
      f = (m.x != 0.f ? 1u : 0u) | (m.y != 0.f ? 0x100u : 0u) | (m.z != 0.f ? 0x10000u : 0u) | (m.w != 0.f ? 0x1000000u : 0u);
    } else {
      const uint4 rr = philox4x32_10(make_uint4((unsigned int)n4 >> 2, step_id, p.row_offset + b, (unsigned int)STREAM_KEEP1),
                                     make_uint2((unsigned int)p.seed, (unsigned int)(p.seed >> 32)));
      const float sc = 5.9604644775390625e-08f;
      f = ((float)(rr.x >> 8) * sc >= p.drop_rate ? 1u : 0u) | ((float)(rr.y >> 8) * sc >= p.drop_rate ? 0x100u : 0u) |
          ((float)(rr.z >> 8) * sc >= p.drop_rate ? 0x10000u : 0u) | ((float)(rr.w >> 8) * sc >= p.drop_rate ? 0x1000000u : 0u);
    }
    keep[u * (FA_P / 4) + (n4 >> 2)] = f;
  }
}

template <int NU>
__device__ __noinline__ void v2_dense_front(const DecParams& p, const V2Params& q, float* scratch, int b0, int t, unsigned int zneed0, unsigned int zneed1,
                                            unsigned long long* prof) {
  const int Tv = p.Tv;
  float* qs = scratch + ((8 * Tv + 3) & ~3);
  __nv_bfloat16* z0s = reinterpret_cast<__nv_bfloat16*>(scratch + att_scratch_floats(Tv));
  __nv_bfloat16* z1s = z0s + V2_ZS_ELEMS;
  uint8_t* keep1 = reinterpret_cast<uint8_t*>(z1s + V2_ZS_ELEMS);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool drop = p.rng_mode != 0 && p.drop_rate > 0.f;
  prof_tick(prof, 0);
  if (t == 0) {
    // first step of the launch: padding utterance slots zero; the decoder input is given (zero frame / init_mel,
    // Taco2.py:163-166,183-187), prenet layer 0 straight from the fp32 kernel
    for (int i = tid; i < V2_ZS_ELEMS; i += TC_PA_THREADS) reinterpret_cast<uint32_t*>(z0s)[i] = 0u;   // both buffers
    pa_sync<TC_PA_THREADS>();
    v2_keep1_fill<NU>(p, keep1, b0, 0);
    for (int i = tid; i < NU * FA_P; i += TC_PA_THREADS) {
      const int u = i / FA_P, f = i - u * FA_P, b = b0 + u * 128;
      float a = __ldg(p.b0 + f);
      if (p.init_mel) {
        const float* x = p.init_mel + (size_t)b * FA_MEL;
        for (int k = 0; k < FA_MEL; ++k) a = fmaf(__ldg(x + k), __ldg(p.W0 + (size_t)k * FA_P + f), a);
      }
      a = fmaxf(a, 0.f);
      if (drop) {
        const float kf = p.rng_mode == 1 ? __ldg(p.keep0 + ((size_t)p.rng_b0 + b) * FA_P + f)
                                         : philox_keep(p.seed, STREAM_KEEP0, p.step_offset, p.row_offset + b, (unsigned int)f, p.drop_rate);
        a = a * kf * p.drop_scale;
      }
      z0s[u * V2_ACT_STRIDE + f] = __float2bfloat16(a);
    }
  } else {
    if (tid == 0) {
      v2_poll(&q.sync->zcnt[b0 >> 7][0], zneed0 * (unsigned int)t);
      if (NU == 2) v2_poll(&q.sync->zcnt[(b0 + 128) >> 7][0], zneed1 * (unsigned int)t);
    }
    pa_sync<TC_PA_THREADS>();
    prof_tick(prof, 1);
    if (tid < NU * 32) {
      const int u = tid >> 5, c = tid & 31;
      *reinterpret_cast<uint4*>(z0s + u * V2_ACT_STRIDE + 8 * c) = __ldcg(reinterpret_cast<const uint4*>(q.z0buf + (size_t)(b0 + u * 128) * FA_P) + c);
    }
  }
  pa_sync<TC_PA_THREADS>();
  const int g = lane >> 2, tq = lane & 3;
  if (wid < 8) {
    // prenet layer 1: feature tiles 2 wid, 2 wid + 1 (fa_wlayer 2: 4 stages of [16 ft][4 kt][32 lanes][16 B])
    const __nv_bfloat16* brow = z0s + g * V2_ACT_STRIDE + 2 * tq;
    const uint4* w1 = reinterpret_cast<const uint4*>(q.wres) + lane;
    float d[2][4] = {};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 a[2][8];
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int kt = 8 * h + k;
          a[s][k] = __ldg(w1 + (size_t)(((kt >> 2) * 16 + 2 * wid + s) * 4 + (kt & 3)) * 32);
        }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(brow + (8 * h + k) * 16), bb1 = *reinterpret_cast<const uint32_t*>(brow + (8 * h + k) * 16 + 8);
#pragma unroll
        for (int s = 0; s < 2; ++s) mma_16816_bf16(d[s], a[s][k], bb0, bb1);
      }
    }
    if (tq == 0) {   // D fragment: rows = features g, g + 8; columns 0, 1 = the CTA's utterances
      const float dscale = p.drop_scale;
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int f = (2 * wid + s) * 16 + g + 8 * hi;
          const float bf = __ldg(p.b1 + f);
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            float y = fmaxf(d[s][2 * hi + u] + bf, 0.f);
            if (drop) y = keep1[u * FA_P + f] ? y * dscale : 0.f;
            z1s[u * V2_ACT_STRIDE + f] = __float2bfloat16(y);
          }
        }
    }
  }
  pa_sync<TC_PA_THREADS>();
  if (wid >= 8) {
    // p(t) -> operand image of LSTMCell 0: one 16 B swizzle chunk per lane, published per utterance
    const int u = wid - 8;
    if (u < NU) {
      const int b = b0 + u * 128;
      *reinterpret_cast<uint4*>(q.actP + act_elem_index(p.MT, b, 8 * lane)) = *reinterpret_cast<const uint4*>(z1s + u * V2_ACT_STRIDE + 8 * lane);
      __syncwarp();
      if (lane == 0) v2_signal(&q.sync->pcnt[b >> 7][0], 1u);
    }
  } else {
    // query projection: feature tile wid (fa_wlayer 3: 2 stages of [8 ft][8 kt][32 lanes][16 B])
    const __nv_bfloat16* brow = z1s + g * V2_ACT_STRIDE + 2 * tq;
    const uint4* wq = reinterpret_cast<const uint4*>(q.wres + V2_WRES_LQ) + lane;
    float d[4] = {};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = __ldg(wq + (size_t)((h * 8 + wid) * 8 + k) * 32);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(brow + (8 * h + k) * 16), bb1 = *reinterpret_cast<const uint32_t*>(brow + (8 * h + k) * 16 + 8);
        mma_16816_bf16(d, a[k], bb0, bb1);
      }
    }
    if (tq == 0) {
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int f = wid * 16 + g + 8 * hi;
        const float bq = __ldg(p.bq + f);
#pragma unroll
        for (int u = 0; u < NU; ++u) qs[u * FA_A + f] = d[2 * hi + u] + bq;
      }
    }
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(prof, 5);
}

// ---- attention (Steps.py:138-166, 215-229) for the CTA's utterances b0 and b0 + 128: decoder_bf16.cuh's attention_a with the
// queries taken from shared memory (v2_dense_front) and ctx written into the parity image + published per warp
template <int NU>
__device__ __forceinline__ void v2_noise_fill(const DecParams& p, float* scratch, int b0, int t) {
  if (!(p.rng_mode != 0 && p.sigmoid_noise > 0.f)) return;
  float* nzbuf = scratch + 6 * p.Tv;
  for (int i = threadIdx.x; i < NU * p.Tv; i += TC_PA_THREADS) {
    const int u = i / p.Tv, j = i - u * p.Tv;
    nzbuf[i] = att_noise(p, t, b0 + u * 128, j);
  }
}

template <int NU>
__device__ __noinline__ void v2_attention(const DecParams& p, const V2Params& q, float* scratch, const float* attv_s, int b0, int t, unsigned long long* prof) {
  constexpr int WPU = FA_WARPS / NU;
  const int Tv = p.Tv;
  float* alig = scratch;
  float* pbuf = scratch + 4 * Tv;
  const float* nzbuf = scratch + 6 * Tv;
  const float* qs = scratch + ((8 * Tv + 3) & ~3);
  float* ctxp = scratch + ((8 * Tv + 3) & ~3) + 256;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int cur = t & 1, prv = cur ^ 1;
  const int XW = p.P1 + p.A;
  int bs[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) bs[u] = b0 + u * 128;
  const float sb = __ldg(p.att_sb);
  const bool noisy = p.rng_mode != 0 && p.sigmoid_noise > 0.f;
  const int uw = NU == 2 ? wid / WPU : 0, wl = wid - uw * WPU;
  const int bw = bs[NU == 2 ? uw : 0];
  const int rg = lane >> 3, cs = lane & 7;
  const uint4* Vl = reinterpret_cast<const uint4*>(q.vproj_bf + (size_t)bw * Tv * 128) + 2 * cs;
  const bool single = Tv <= WPU * 32;
  uint4 kv[8][2];
  auto load_rows = [&](int base) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = base + 4 * i + rg;
      if (j < Tv) { kv[i][0] = __ldg(Vl + (size_t)j * 16); kv[i][1] = __ldg(Vl + (size_t)j * 16 + 1); }
      else { kv[i][0] = make_uint4(0u, 0u, 0u, 0u); kv[i][1] = kv[i][0]; }
    }
  };
  load_rows(wl * 32);
  float qr[16], vr[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 b = *reinterpret_cast<const float4*>(attv_s + 16 * cs + 4 * c);
    vr[4 * c] = b.x; vr[4 * c + 1] = b.y; vr[4 * c + 2] = b.z; vr[4 * c + 3] = b.w;
    const float4 a = *reinterpret_cast<const float4*>(qs + uw * FA_A + 16 * cs + 4 * c);
    qr[4 * c] = a.x; qr[4 * c + 1] = a.y; qr[4 * c + 2] = a.z; qr[4 * c + 3] = a.w;
  }
  if (t == 0) {
    for (int i = tid; i < NU * Tv; i += TC_PA_THREADS) {
      const int u = i / Tv, j = i - u * Tv;
      alig[(u * 2 + prv) * Tv + j] = __ldcg(p.align + ((size_t)prv * p.B + bs[u]) * Tv + j);
    }
    v2_noise_fill<NU>(p, scratch, b0, 0);
    pa_sync<TC_PA_THREADS>();
  }
  // ---- pass 1: energies
  for (int base = wl * 32; base < Tv; base += WPU * 32) {
    if (base != wl * 32) load_rows(base);
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t w[8] = {kv[i][0].x, kv[i][0].y, kv[i][0].z, kv[i][0].w, kv[i][1].x, kv[i][1].y, kv[i][1].z, kv[i][1].w};
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        a0 = fmaf(vr[2 * c], tanh_fast(qr[2 * c] + bf16lo(w[c])), a0);
        a1 = fmaf(vr[2 * c + 1], tanh_fast(qr[2 * c + 1] + bf16hi(w[c])), a1);
      }
      e[i] = a0 + a1;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] += __shfl_xor_sync(0xffffffffu, e[i], o);
    float er = e[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) er = cs == i ? e[i] : er;
    const int j = base + 4 * cs + rg;
    if (j < Tv) {
      er += sb;
      if (noisy) er = fmaf(p.sigmoid_noise, nzbuf[uw * Tv + j], er);
      pbuf[uw * Tv + j] = sigmoid_fast(er);
    }
  }
  pa_sync<TC_PA_THREADS>();
  // ---- pass 2: alignment recurrence (Steps.py:223-229)
  for (int i = tid; i < NU * Tv; i += TC_PA_THREADS) {
    const int u = i / Tv, j = i - u * Tv, b = bs[u];
    const float* prev_s = alig + (u * 2 + prv) * Tv;
    const float* ps = pbuf + u * Tv;
    float a = prev_s[j] * ps[j];
    if (j > 0) a = fmaf(prev_s[j - 1], 1.0f - ps[j - 1], a);
    alig[(u * 2 + cur) * Tv + j] = a;
    if (p.out_align) p.out_align[((size_t)b * p.To + t) * Tv + j] = a;
    if (t == p.T - 1) p.align[((size_t)cur * p.B + b) * Tv + j] = a;
  }
  pa_sync<TC_PA_THREADS>();
  // ---- pass 3: context = alignment . V' (Steps.py:164)
  {
    const float* cur_s = alig + (uw * 2 + cur) * Tv;
    float cx[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) cx[c] = 0.f;
    for (int base = wl * 32; base < Tv; base += WPU * 32) {
      if (!single) load_rows(base);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = base + 4 * i + rg;
        const float a = j < Tv ? cur_s[j] : 0.f;
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
  }
  pa_sync<TC_PA_THREADS>();
  if (tid < 128 * NU) {   // whole warps: 4 per utterance
    const int u = tid >> 7, n = tid & 127, b = bs[u];
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < WPU; ++w) c += ctxp[(u * WPU + w) * 128 + n];
    __nv_bfloat16* img = q.actC + (size_t)cur * 2 * p.MT * 128 * 64;
    img[act_elem_index(p.MT, b, n)] = __float2bfloat16(c);
    if (t == p.T - 1) {
      p.xin[(size_t)b * XW + p.P1 + n] = c;
      if (p.out_ctx) p.out_ctx[(size_t)b * p.A + n] = c;
    }
    __syncwarp();
    if (lane == 0) v2_signal(&q.sync->ctxcnt[b >> 7][0], 1u);
  }
  prof_tick(prof, 2);
}

constexpr size_t V2_RING_BYTES = (size_t)V2_NSTAGE * V2_STAGE_BYTES;

__global__ void __launch_bounds__(V2_THREADS, 1) decoder_bf16_v2_kernel(const __grid_constant__ DecParams p, const __grid_constant__ V2Params q) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * V2_NSTAGE + 3];
  __shared__ float bias_s[128];
  __shared__ float fbias_s[V2_FOLD_N];
  __shared__ __align__(16) float attv_s[128];
  __shared__ __align__(16) DecParams p_sh;
  __shared__ __align__(16) V2Params q_sh;
  __shared__ unsigned long long prof_sh[PROF_SLOTS + 1];
  const int cta = blockIdx.x;
  if (cta >= TC_LSTM_CTAS) return;   // the dataflow kernel only uses the 128 LSTM CTAs
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int MT = p.MT;
  const int ug = cta >> 1, mt_c = cta & 1;
  const bool lstm_act = mt_c < MT;
  const int rows_mt = lstm_act ? min(128, p.B - mt_c * 128) : 0;
  // fold duty: unit groups 0..21 of an m-tile = (slice, 64-row half)
  const int f_slice = ug >> 1, f_half = ug & 1;
  const int rows_half = max(0, min(64, rows_mt - 64 * f_half));
  const bool fold_cta = lstm_act && ug < V2_FOLD_CTAS && rows_half > 0;
  // attention + prenet-1 / query duty: utterances cta and cta + 128
  const int att_nu = cta + 128 < p.B ? 2 : (cta < p.B ? 1 : 0);
  auto z_need = [&](int mt) { return (unsigned int)V2_FOLD_ZSLICES * (min(128, p.B - mt * 128) > 64 ? 2u : 1u); };
  const unsigned int zneed0 = z_need(0), zneed1 = MT > 1 ? z_need(1) : 0u;

  uint8_t* stages = sm;
  float* scratch = reinterpret_cast<float*>(sm + V2_RING_BYTES);   // attention + prenet scratch
  uint64_t* full = bars;
  uint64_t* d1_full = bars + 2 * V2_NSTAGE;
  uint64_t* d2_full = bars + 2 * V2_NSTAGE + 1;
  uint64_t* z_full = bars + 2 * V2_NSTAGE + 2;

  for (int i = tid; i < (int)(sizeof(DecParams) / 4); i += V2_THREADS) reinterpret_cast<uint32_t*>(&p_sh)[i] = reinterpret_cast<const uint32_t*>(&p)[i];
  for (int i = tid; i < (int)(sizeof(V2Params) / 4); i += V2_THREADS) reinterpret_cast<uint32_t*>(&q_sh)[i] = reinterpret_cast<const uint32_t*>(&q)[i];
  unsigned long long* prof_s = q.prof ? prof_sh : nullptr;
  if (tid == 0) {
    for (int i = 0; i < PROF_SLOTS; ++i) prof_sh[i] = 0;
    prof_sh[PROF_SLOTS] = (unsigned long long)clock64();
    for (int i = 0; i < 2 * V2_NSTAGE + 3; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (tid < 128) {
    bias_s[tid] = __ldg(q.bias + (size_t)ug * 128 + tid);
    attv_s[tid] = __ldg(p.att_v + tid);
  }
  if (fold_cta && tid < V2_FOLD_N) fbias_s[tid] = __ldg(q.fold_bias + f_slice * V2_FOLD_N + tid);
  if (wid == 0) tmem_alloc(&tmem_base_s, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t stages_sa = smem_u32(stages);
  V2Sync* sy = q.sync;
  const int rot_h = ug % TC_NKB_H;
  const uint32_t astride = (uint32_t)MT * TC_A_BYTES;
  const uint32_t abytes = (uint32_t)rows_mt * 128u, hbytes = (uint32_t)rows_half * 128u;

  if (wid == TC_PA_WARPS) {
    // ================= weight-copy warp =================
    if (lstm_act) {
      TcRing ring;
      ring.stage = 0; ring.bits = 0;
      const uint8_t* wimg_cta = reinterpret_cast<const uint8_t*>(q.wimg) + (size_t)ug * TC_IMG_BYTES;
      const uint8_t* fold_s = q.fold_img + (size_t)f_slice * V2_FOLD_IMG_BYTES;
      const uint32_t tot = abytes + TC_B_BYTES, tot_f = 2u * hbytes + 2u * V2_FOLD_B_BYTES;
      for (int t = 0; t <= p.T; ++t) {
        if (fold_cta && t > 0) {
          ring = v2_prod_w<V2_FOLD_UNITS>(ring, full, stages, fold_s, 2 * V2_FOLD_B_BYTES, 2 * V2_FOLD_B_BYTES, tot_f, ug & 7);
          ring = v2_prod_w<1>(ring, full, stages, fold_s + (size_t)TC_NKB_H * V2_FOLD_B_BYTES, 0u, 2 * V2_FOLD_B_BYTES, tot_f, 0);
        }
        if (t == p.T) break;
        ring = v2_prod_w<TC_NKB_H>(ring, full, stages, wimg_cta + TC_IMG_U2, TC_B_BYTES, TC_B_BYTES, tot, rot_h);                         // U2
        ring = v2_prod_w<TC_NKB_H>(ring, full, stages, wimg_cta + TC_IMG_WU + TC_B_BYTES, 2 * TC_B_BYTES, TC_B_BYTES, tot, rot_h);       // U1
        ring = v2_prod_w<4>(ring, full, stages, wimg_cta + TC_IMG_W1X, TC_B_BYTES, TC_B_BYTES, tot, ug & 3);                             // W1x (p)
        ring = v2_prod_w<2>(ring, full, stages, wimg_cta + TC_IMG_W1X + 4 * TC_B_BYTES, TC_B_BYTES, TC_B_BYTES, tot, ug & 1);            // W1x (ctx)
        ring = v2_prod_w<TC_NKB_H>(ring, full, stages, wimg_cta + TC_IMG_WU, 2 * TC_B_BYTES, TC_B_BYTES, tot, rot_h);                    // W2
      }
    }
  } else if (wid == TC_PA_WARPS + 2) {
    // ================= activation-copy warp =================
    if (lstm_act) {
      TcRing ring;
      ring.stage = 0; ring.bits = 0;
      const size_t tile_off = (size_t)mt_c * TC_A_BYTES;
      const uint8_t* actP_b = reinterpret_cast<const uint8_t*>(q.actP) + tile_off;
      const uint8_t* actC_b = reinterpret_cast<const uint8_t*>(q.actC) + tile_off;      // + parity * 2 * astride
      const uint8_t* actH1_b = reinterpret_cast<const uint8_t*>(q.actH1) + tile_off;    // + parity * 16 * astride
      const uint8_t* actH2_b = reinterpret_cast<const uint8_t*>(q.actH2) + tile_off;
      const unsigned int* kb_h1 = &sy->kb_h1[mt_c * TC_NKB_H][0];
      const unsigned int* kb_h2 = &sy->kb_h2[mt_c * TC_NKB_H][0];
      const unsigned int urows = (unsigned int)rows_mt;
      for (int t = 0; t <= p.T; ++t) {
        const unsigned int ut = (unsigned int)t;
        const size_t par_c = (size_t)(t & 1) * 2 * astride, par_p = (size_t)((t & 1) ^ 1) * 2 * astride;
        const size_t h1_c = (size_t)(t & 1) * TC_NKB_H * astride, h1_p = (size_t)((t & 1) ^ 1) * TC_NKB_H * astride;
        if (fold_cta && t > 0)   // fold stream: Z[64 x 32] = [h2(t-1) || ctx(t-1)] (this CTA's row half) . [Wf | P] slice
          ring = v2_prod_a_fold(ring, full, stages, actH2_b + (size_t)f_half * 8192, actC_b + par_p + (size_t)f_half * 8192, astride, hbytes, ug & 7, kb_h2,
                                4u * ut, prof_s ? prof_s + 8 : nullptr);
        if (t == p.T) break;
        // D2 = h2(t-1) . U2.  The first MMA of this stream overwrites D2, which this CTA's own LSTMCell-1 epilogue of step t-1
        // reads: wait for the CTA's OWN h2 k-block counter first (the epilogue publishes it after its TMEM loads).
        if (t > 0) {
          if (lane == 0) v2_poll(kb_h2 + (size_t)(ug >> 2) * 32, 4u * ut);
          __syncwarp();
        }
        ring = v2_prod_a<TC_NKB_H, 2>(ring, full, stages, actH2_b, astride, abytes, rot_h, kb_h2, 4u * ut, nullptr);
        // D1 = h1(t-1) . U1 (complete since the LSTMCell-1 stream of step t-1 consumed it)
        ring = v2_prod_a<TC_NKB_H, 0>(ring, full, stages, actH1_b + h1_p, astride, abytes, rot_h, nullptr, 0u, nullptr);
        // D1 += p(t) . W1x[0:256]  |  D1 += ctx(t) . W1x[256:384]
        ring = v2_prod_a<4, 1>(ring, full, stages, actP_b, astride, abytes, ug & 3, &sy->pcnt[mt_c][0], urows * (ut + 1u), nullptr);
        ring = v2_prod_a<2, 1>(ring, full, stages, actC_b + par_c, astride, abytes, ug & 1, &sy->ctxcnt[mt_c][0], 4u * urows * (ut + 1u), nullptr);
        // D2 += h1(t) . W2, k-block by k-block as the LSTMCell-0 epilogues of this m-tile publish them
        ring = v2_prod_a<TC_NKB_H, 2>(ring, full, stages, actH1_b + h1_c, astride, abytes, rot_h, kb_h1, 4u * (ut + 1u), prof_s ? prof_s + 10 : nullptr);
      }
    }
  } else if (wid == TC_PA_WARPS + 1) {
    // ================= MMA warp =================
    if (lstm_act) {
      TcRing ring;
      ring.stage = 0; ring.bits = 0;
      // per-segment issue time of the MMA warp (diagnostics: slots 6, 7, 12-14 of the phase profile; 15 = fold)
      long long last = clock64();
      auto seg_tick = [&](int slot) {
        if (prof_s && lane == 0) {
          const long long now = clock64();
          prof_s[slot] += (unsigned long long)(now - last);
          last = now;
        }
      };
      for (int t = 0; t <= p.T; ++t) {
        if (fold_cta && t > 0) {
          ring = v2_fold_consume(ring, full, stages_sa, tmem + V2_Z, z_full, nullptr);
          seg_tick(15);
        }
        if (t == p.T) break;
        ring = v2_seg_consume<TC_NKB_H, true, 128, 64>(ring, full, stages_sa, tmem + V2_D2, nullptr);
        seg_tick(6);
        ring = v2_seg_consume<TC_NKB_H, true, 128, 64>(ring, full, stages_sa, tmem + V2_D1, nullptr);
        seg_tick(7);
        ring = v2_seg_consume<4, false, 128, 64>(ring, full, stages_sa, tmem + V2_D1, nullptr);
        seg_tick(12);
        ring = v2_seg_consume<2, false, 128, 64>(ring, full, stages_sa, tmem + V2_D1, d1_full);
        seg_tick(13);
        ring = v2_seg_consume<TC_NKB_H, false, 128, 64>(ring, full, stages_sa, tmem + V2_D2, d2_full);
        seg_tick(14);
      }
    }
  } else {
    // ================= phase-A / epilogue warps 0-9 =================
    const bool epi = lstm_act && wid < 8;
    const int erow = mt_c * 128 + (wid & 3) * 32 + lane;
    const int ub = 2 * ug + (wid >> 2);
    const bool erow_ok = epi && erow < p.B;
    const uint32_t t_row = tmem + ((uint32_t)((wid & 3) * 32) << 16);
    const uint32_t t_half = (uint32_t)(wid >> 2);
    const bool zwarp = fold_cta && wid < 4;
    const int zrow = mt_c * 128 + f_half * 64 + wid * 16 + lane;
    const bool zrow_ok = zwarp && lane < 16 && zrow < p.B;
    const unsigned int z_pub = (f_slice < V2_FOLD_ZSLICES) ? 1u : 0u;
    uint8_t* keep1 = reinterpret_cast<uint8_t*>(scratch + att_scratch_floats(p.Tv)) + 2 * V2_ZS_ELEMS * 2;
    if (epi) {
      float c1[8], c2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        c1[u] = erow_ok ? __ldcg(p.c1 + (size_t)erow * TC_U + ub * 8 + u) : 0.f;
        c2[u] = erow_ok ? __ldcg(p.c2 + (size_t)erow * TC_U + ub * 8 + u) : 0.f;
      }
      tmem_st8(t_row + V2_C1 + t_half * 8u, c1);
      tmem_st8(t_row + V2_C2 + t_half * 8u, c2);
    }
    const size_t h1_par = (size_t)TC_NKB_H * MT * 128 * 64;   // elements between the two h1 parity images
    for (int t = 0; t <= p.T; ++t) {
      if (zwarp && t > 0) {
        v2_fold_epilogue(p_sh, q_sh, z_full, (uint32_t)(t - 1) & 1u, t_row + V2_Z, fbias_s, f_slice, zrow, zrow_ok, t);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (tid == 0 && z_pub && t < p.T) v2_signal(&sy->zcnt[mt_c][0], 1u);
      }
      if (t == p.T) break;
      if (att_nu == 2) {
        v2_dense_front<2>(p_sh, q_sh, scratch, cta, t, zneed0, zneed1, prof_s);
        v2_attention<2>(p_sh, q_sh, scratch, attv_s, cta, t, prof_s);
      } else if (att_nu == 1) {
        v2_dense_front<1>(p_sh, q_sh, scratch, cta, t, zneed0, zneed1, prof_s);
        v2_attention<1>(p_sh, q_sh, scratch, attv_s, cta, t, prof_s);
      }
      // ---------------- LSTMCell 0 epilogue -> h1(t) (parity image t & 1) -------------------------------------------
      const bool h32 = t == p.T - 1;   // the fp32 copies of h1 / h2 are only read after the last step
      if (epi)
        lstm_epilogue(d1_full, (uint32_t)t & 1u, t_row + V2_D1 + t_half * 32u, t_row + V2_C1 + t_half * 8u, bias_s + t_half * 32u, erow_ok, erow, ub, MT,
                      q.actH1 + (size_t)(t & 1) * h1_par, h32 ? p.h1 + ((size_t)(t & 1) * p.B + erow) * TC_U : nullptr);
      pa_sync<TC_PA_THREADS>();
      if (tid == 0 && lstm_act) v2_signal(&sy->kb_h1[mt_c * TC_NKB_H + (ug >> 2)][0], 1u);
      prof_tick(prof_s, 3);
      if (t + 1 < p.T) {   // off the critical path: the attention noise and the prenet-1 dropout flags of step t+1
        if (att_nu == 2) { v2_noise_fill<2>(p_sh, scratch, cta, t + 1); v2_keep1_fill<2>(p_sh, keep1, cta, t + 1); }
        else if (att_nu == 1) { v2_noise_fill<1>(p_sh, scratch, cta, t + 1); v2_keep1_fill<1>(p_sh, keep1, cta, t + 1); }
      }
      // ---------------- LSTMCell 1 epilogue -> h2(t) ------------------------------------------------------------------
      if (epi)
        lstm_epilogue(d2_full, (uint32_t)t & 1u, t_row + V2_D2 + t_half * 32u, t_row + V2_C2 + t_half * 8u, bias_s + 64 + t_half * 32u, erow_ok, erow, ub,
                      MT, q.actH2, h32 ? p.h2 + ((size_t)(t & 1) * p.B + erow) * TC_U : nullptr);
      pa_sync<TC_PA_THREADS>();
      if (tid == 0 && lstm_act) v2_signal(&sy->kb_h2[mt_c * TC_NKB_H + (ug >> 2)][0], 1u);
      prof_tick(prof_s, 4);
    }
    if (epi) {
      float c1[8], c2[8];
      tmem_ld8(t_row + V2_C1 + t_half * 8u, c1);
      tmem_ld8(t_row + V2_C2 + t_half * 8u, c2);
      if (erow_ok) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          p.c1[(size_t)erow * TC_U + ub * 8 + u] = c1[u];
          p.c2[(size_t)erow * TC_U + ub * 8 + u] = c2[u];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (q.prof && tid == 0)
    for (int i = 0; i < PROF_SLOTS; ++i) q.prof[(size_t)cta * PROF_SLOTS + i] = prof_sh[i];
  if (wid == 0) tmem_dealloc(tmem_base_s, TC_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct V2State {
  uint8_t* fold_img = nullptr;
  float* fold_bias = nullptr;
  __nv_bfloat16* act = nullptr;   // actP | actC | actH1 | actH2 for MT = 2
  __nv_bfloat16* z0buf = nullptr;
  V2Sync* sync = nullptr;
  bool ready = false;
};

inline void v2_release(V2State& v) {
  cudaFree(v.fold_img);
  cudaFree(v.fold_bias);
  cudaFree(v.act);
  cudaFree(v.z0buf);
  cudaFree(v.sync);
  v = V2State();
}

inline int v2_prepare(V2State& v, const GstkConfig& c, const std::map<std::string, std::vector<float>>& hw, std::string& err) {
  if (v.ready) return GSTK_OK;
  const std::string dd = "Decoder/Decoder_Step/";
  const std::vector<float>& P = hw.at(dd + "Projection/kernel");      // [1152][81]
  const std::vector<float>& bP = hw.at(dd + "Projection/bias");
  const std::vector<float>& W0 = hw.at(dd + "Prenet/dense/kernel");   // [80][256]
  const std::vector<float>& b0 = hw.at(dd + "Prenet/dense/bias");
  const int K = FA_HC, PD = FA_PD, mel = FA_MEL, F = FA_P;
  // column n of the fused [K x (256 + 96)] matrix: n < 256 -> Wf[:, n] = P[:, :mel] . W0[:, n]; n >= 256 -> P[:, n - 256]
  std::vector<float> M((size_t)K * V2_FOLD_SLICES * V2_FOLD_N, 0.f), bias((size_t)V2_FOLD_SLICES * V2_FOLD_N, 0.f);
  const int NC = V2_FOLD_SLICES * V2_FOLD_N;
  for (int k = 0; k < K; ++k) {
    for (int f = 0; f < F; ++f) {
      double a = 0.0;
      for (int m = 0; m < mel; ++m) a += (double)P[(size_t)k * PD + m] * (double)W0[(size_t)m * F + f];
      M[(size_t)k * NC + f] = (float)a;
    }
    for (int j = 0; j < PD; ++j) M[(size_t)k * NC + F + j] = P[(size_t)k * PD + j];
  }
  for (int f = 0; f < F; ++f) {
    double a = b0[f];
    for (int m = 0; m < mel; ++m) a += (double)bP[m] * (double)W0[(size_t)m * F + f];
    bias[f] = (float)a;
  }
  for (int j = 0; j < PD; ++j) bias[F + j] = bP[j];
  std::vector<__nv_bfloat16> img((size_t)V2_FOLD_SLICES * V2_FOLD_IMG_BYTES / 2, __float2bfloat16(0.f));
  for (int s = 0; s < V2_FOLD_SLICES; ++s)
    for (int kb = 0; kb < V2_FOLD_NKB; ++kb) {
      __nv_bfloat16* blk = img.data() + ((size_t)s * V2_FOLD_IMG_BYTES + (size_t)kb * V2_FOLD_B_BYTES) / 2;
      for (int n = 0; n < V2_FOLD_N; ++n)
        for (int k = 0; k < 64; ++k) blk[sw128_offset_bytes(n, k) / 2] = __float2bfloat16(M[(size_t)(kb * 64 + k) * NC + s * V2_FOLD_N + n]);
    }
  auto fail = [&](const char* m) { err = m; return GSTK_ECUDA; };
  if (cudaMalloc((void**)&v.fold_img, img.size() * 2) != cudaSuccess) return fail("cudaMalloc(fold_img) failed");
  if (cudaMalloc((void**)&v.fold_bias, bias.size() * 4) != cudaSuccess) return fail("cudaMalloc(fold_bias) failed");
  const size_t act_elems = (size_t)(4 + 4 + 2 * TC_NKB_H + TC_NKB_H) * 2 * 128 * 64;
  if (cudaMalloc((void**)&v.act, act_elems * 2) != cudaSuccess) return fail("cudaMalloc(v2 act) failed");
  if (cudaMalloc((void**)&v.z0buf, (size_t)TC_MAX_B * FA_P * 2) != cudaSuccess) return fail("cudaMalloc(z0buf) failed");
  if (cudaMalloc((void**)&v.sync, sizeof(V2Sync)) != cudaSuccess) return fail("cudaMalloc(v2 sync) failed");
  if (cudaMemcpy(v.fold_img, img.data(), img.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  if (cudaMemcpy(v.fold_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  (void)c;
  v.ready = true;
  return GSTK_OK;
}

inline size_t v2_smem_bytes(const DecParams& p) {
  return 1024 + V2_RING_BYTES + 4 * (size_t)att_scratch_floats(p.Tv) + (size_t)V2_DENSE_SCRATCH;
}

// can this launch take the dataflow kernel?  (fast-path shapes, free-running, enough dense CTAs, shared memory)
inline bool v2_usable(const Bf16State& st, const DecParams& p, int num_sms) {
  if (!st.fast_a || p.mode != 0 || p.T < 1 || p.B > TC_MAX_B || num_sms < TC_LSTM_CTAS) return false;
  return v2_smem_bytes(p) <= 227 * 1024;
}

inline int v2_decode(Bf16State& st, V2State& v, const GstkConfig& c, DecParams& p, int num_sms, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1,
                     int64_t& launches, std::string& err) {
  auto fail = [&](int code, const std::string& m) { err = m; return code; };
  (void)c;
  const int MT = (p.B + 127) / 128;
  const size_t smem = v2_smem_bytes(p);
  cudaError_t e;
  V2Params q;
  q.wimg = st.wimg;
  q.bias = st.bias;
  q.fold_img = v.fold_img;
  q.fold_bias = v.fold_bias;
  q.wres = st.wimgA + FA_L1.base;
  const size_t tile = (size_t)MT * 128 * 64;
  q.actP = v.act;
  q.actC = q.actP + 4 * tile;
  q.actH1 = q.actC + 4 * tile;
  q.actH2 = q.actH1 + 2 * TC_NKB_H * tile;
  q.z0buf = v.z0buf;
  q.sync = v.sync;
  p.MT = MT;
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
  // operand images: p / ctx zero (rows >= B must stay finite), h1(-1) into parity image 1, h2(-1)
  if ((e = cudaMemsetAsync(q.actP, 0, 8 * tile * 2, stream)) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
  if ((e = cudaMemsetAsync(v.sync, 0, sizeof(V2Sync), stream)) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h1 + (size_t)p.B * TC_U, q.actH1 + TC_NKB_H * tile, p.B, MT);
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h1 + (size_t)p.B * TC_U, q.actH1, p.B, MT);   // rows >= B of parity image 0 zero-filled as well
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h2 + (size_t)p.B * TC_U, q.actH2, p.B, MT);
  launches += 3;
  if ((e = cudaFuncSetAttribute(decoder_bf16_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  int occ = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decoder_bf16_v2_kernel, V2_THREADS, smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  if (occ < 1) return fail(GSTK_EINVAL, "bf16 dataflow decoder kernel does not fit on an SM");
  void* args[] = {&p, &q};
  cudaEventRecord(ev0, stream);
  if ((e = cudaLaunchCooperativeKernel((void*)decoder_bf16_v2_kernel, dim3(num_sms), dim3(V2_THREADS), args, smem, stream)) != cudaSuccess)
    return fail(GSTK_ECUDA, std::string("cooperative launch failed: ") + cudaGetErrorString(e));
  cudaEventRecord(ev1, stream);
  launches += 1;
  return GSTK_OK;
}

}  // namespace gstk
