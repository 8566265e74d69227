// Persistent bf16 tensor-core decoder (tcgen05 / TMEM / bulk-async copies), the throughput mode.
//
// Same loop structure as decoder_fp32.cuh (one cooperative launch for the whole decode, three grid
// barriers per step, phase A = projection/prenet/attention per utterance in fp32), but the two
// LSTMCells (98 % of the FLOPs, Modules/Taco2.py:77-85,111) run on the 5th-gen tensor cores:
//
//   * LSTM CTA c < 128 owns batch m-tile (c & 1) (128 rows) x hidden units [16 (c >> 1), +16) of BOTH cells = 64 gate
//     columns per cell.  A tcgen05.mma costs ~70-90 cycles here whatever its N (the 128 x 16 A slice is re-fetched from shared
//     memory for every instruction), so the tile is as wide in N as the work allows: one N=64 (or N=128 for W2|U1) MMA per
//     k16 step instead of two m-tiles x N=32.  The weight slice of a unit group is packed once (host) as K-blocks of
//     [64 rows x 64 k] bf16 in the canonical K-major SWIZZLE_128B layout and streamed with the activations.
//   * activations (p || ctx, h1, h2) live in global memory as bf16 *pre-swizzled UMMA operand
//     images* [k-block][m-tile][128 rows][128 B]: the epilogue thread that owns (row, 8 units) stores
//     exactly one 16-byte swizzle chunk, and a consumer brings a tile in with ONE bulk async copy
//     (cp.async.bulk, SASS UBLKCP) that completes on an mbarrier - no tensor map needed.
//   * D[batch tile 128, 64 gate cols] (+)= A[128, 64] . B[64, 64]^T with tcgen05.mma kind::f16 (K=16 x4 per k-block), fp32
//     accumulators in TMEM, issued by one elected lane of the MMA warp; the copy warp runs the operand ring
//     (TC_NSTAGE stages, full/empty mbarriers, tcgen05.commit).
//   * epilogue warps 0-7 (lane quarter x unit half) read their row's 32 accumulator columns with tcgen05.ld, apply the LSTM
//     point-wise update with the cell state kept in TMEM for the whole decode, and publish h.
//   * the recurrent halves are taken off the critical path: h2(t-1).U2 and h1(t-1).U1 stream through the tensor core during
//     the dense-layer / attention phases of step t, when the LSTM pipeline would otherwise idle.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/gstk.h"
#include "common.cuh"
#include "decoder_fp32.cuh"
#include "umma.cuh"

namespace gstk {

constexpr int TC_U = 1024;          // LSTM units per cell (both cells)
constexpr int TC_KX = 384;          // prenet + attention size
constexpr int TC_NKB_X = TC_KX / 64;   // 6
constexpr int TC_NKB_H = TC_U / 64;    // 16
constexpr int TC_LSTM_CTAS = 128;   // (m-tile, unit group of 16) pairs
constexpr int TC_UG = TC_U / 16;    // 64 unit groups
constexpr int TC_NSTAGE = 5;      // 32 KB stages of the LSTM operand ring
constexpr int TC_NSTAGE_BC = 5;   // (phases B and C use the same ring)
constexpr int TC_A_BYTES = 128 * 128;   // one activation tile (128 rows x 64 bf16)
constexpr int TC_HB_BYTES = 32 * 128;   // weight block of one 8-unit half (32 gate rows x 64 k)
constexpr int TC_B_BYTES = 2 * TC_HB_BYTES;   // one weight block of a unit group (64 gate rows: [half][gate][unit])
// one pipeline unit = one k-block: the activation tile of the CTA's m-tile + up to two weight blocks (W2 | U1 of that k-block)
constexpr int TC_STAGE_W = TC_A_BYTES;
constexpr int TC_STAGE_BYTES = TC_STAGE_W + 2 * TC_B_BYTES;  // 32 KB
// per-unit-group weight image: [W1x: 6 blocks][per k-block: W2 | U1 (adjacent => one N=128 B operand)][U2: 16 blocks]
constexpr int TC_IMG_W1X = 0, TC_IMG_WU = TC_NKB_X * TC_B_BYTES, TC_IMG_U2 = TC_IMG_WU + TC_NKB_H * 2 * TC_B_BYTES;
constexpr int TC_IMG_BYTES = TC_IMG_U2 + TC_NKB_H * TC_B_BYTES;  // 432 KB
constexpr int TC_THREADS = 384;     // 12 warps => up to 168 registers per thread
constexpr int TC_PA_THREADS = TC_THREADS - 64;  // warps 0-9 run phase A; warp 10 = copy producer, warp 11 = MMA issuer
constexpr int TC_PA_WARPS = TC_PA_THREADS / 32;
// TMEM columns: [D2 (64) | D1 (64)] so that W2|U1 can be one N=128 MMA; cell states c1, c2 (2 halves x 8 units) behind them
constexpr int TC_TMEM_COLS = 256;
constexpr uint32_t TC_D2 = 0, TC_D1 = 64, TC_C1 = 128, TC_C2 = 144;
constexpr int TC_MAX_B = 256;
// Phase-A dense layers (projection | prenet x2 | query) run on the DENSE CTAs (blockIdx >= TC_LSTM_CTAS, the SMs the LSTM
// tiling leaves free): each handles <= DA_MAXU utterances with mma.sync (weights = A operand, utterances = the N columns)
// and streams the fragment-ordered weights through a DA_WSTAGES-deep ring of bulk-copied stages (<= FA_TPS tiles of 512 B).
// The ring runs DA_WSTAGES stages ahead across the grid barriers, so most of the projection weights of step t+1 are already
// in shared memory when h2(t) is published.
constexpr int FA_TPS = 64, FA_WSTAGE_BYTES = FA_TPS * 512;   // 32 KB weight stages
constexpr int DA_WSTAGES = 4;
constexpr int DA_MAXU = 16;   // utterances per dense CTA: two n-tiles of mma.m16n8k16
constexpr int DA_MMA_WARPS = 8;   // warps issuing mma.sync in a dense CTA: two per SM sub-partition, equal tile counts (HMMA issue is the bound)
// k16 tiles per stage: prenet 4, query 8; the projection (6 feature tiles) takes 8 so that its four k-slices are 2 tiles each
__host__ __device__ constexpr int fa_kts(int NF) { return NF == 6 ? 8 : (FA_TPS / NF > 0 ? FA_TPS / NF : 1); }

struct Bf16Params {
  const __nv_bfloat16* wimg;  // [TC_UG][TC_IMG_BYTES] per-unit-group swizzled weight blocks (TC_IMG_*)
  const float* bias;          // [TC_UG][2 cells][64]  (half*32 + gate*8 + u)
  __nv_bfloat16* actX;        // [6][MT][128][64]
  __nv_bfloat16* actH1;       // [16][MT][128][64]
  __nv_bfloat16* actH2;       // [16][MT][128][64]
  // phase A fast path (SMA): fragment-ordered bf16 weights and the bf16 copy of V'
  const uint8_t* wimgA;           // stage-ordered fragment images of Projection | Prenet0 | Prenet1 | Query (fa_wlayer)
  const __nv_bfloat16* vproj_bf;  // [B][Tv][128]
  float* qbuf;                    // [B][128] projected queries: dense CTAs -> attention CTAs
  unsigned long long* prof;       // [grid][PROF_SLOTS] accumulated clock64 ticks per phase, or null
};

// ring position of one pipeline role (producer and MMA warp each keep their own copy; both walk the
// same sequence of units, so the copies stay in step)
struct TcRing {
  uint32_t stage, bits;  // bit s of `bits` = parity of the number of completed uses of stage s
  __device__ __forceinline__ uint32_t phase() const { return (bits >> stage) & 1u; }
  template <int NS>
  __device__ __forceinline__ void advance() {
    bits ^= 1u << stage;
    if (++stage == NS) stage = 0;
  }
};

// copy warp: wait until this CTA has passed grid barrier number `need`; false = the barrier failed (leave the kernel)
__device__ __forceinline__ bool tc_wait_gen(const unsigned int* gen_s, unsigned int need) {
  unsigned int v;
  for (;;) {
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(gen_s)) : "memory");
    if (v >= need) break;
    __nanosleep(20);
  }
  return v != 0xFFFFFFFFu;
}

// Producer warp: walks the k-blocks (rotated by `rot`) of one segment and issues the bulk copies as stages free up: the
// activation tile of the CTA's m-tile plus the weight block(s) of that k-block (wbytes = 8 KB, or 16 KB for W2|U1).
// A lone warp issues dependent instructions every ~6-10 cycles, so the per-unit instruction count IS the pipeline's
// throughput limit (measured: tools/ubench_handoff.cu): units are as large as shared memory allows and the loop body is
// kept to pointer bumps.  `act` points at the CTA's m-tile of k-block 0, astride = bytes between k-blocks.
template <int NKB, int NS>
__device__ __forceinline__ void tc_produce(TcRing& r, uint64_t* full, uint64_t* empty, uint8_t* stages, const uint8_t* act, uint32_t astride,
                                           const uint8_t* wsrc, uint32_t wstride, uint32_t wbytes, uint32_t abytes, int rot) {
  const uint32_t total = abytes + wbytes;
  int kb = rot;
  const uint8_t* a = act + (size_t)kb * astride;
  const uint8_t* w = wsrc + (size_t)kb * wstride;
  r.stage = 0;  // every segment starts at stage 0 (producer and MMA warp agree; the per-stage parities carry over)
  for (int i = 0; i < NKB; ++i) {
    mbar_wait(&empty[r.stage], r.phase() ^ 1u);
    uint8_t* st = stages + (size_t)r.stage * TC_STAGE_BYTES;
    if (elect_one()) {
      mbar_arrive_expect_tx(&full[r.stage], total);
      bulk_g2s(st, a, abytes, &full[r.stage]);
      bulk_g2s(st + TC_STAGE_W, w, wbytes, &full[r.stage]);
    }
    __syncwarp();
    r.template advance<NS>();
    if (++kb == NKB) { kb = 0; a = act; w = wsrc; }
    else { a += astride; w += wstride; }
  }
}

// Same, for a segment whose ACTIVATIONS only exist once this CTA has passed grid barrier `need` (published in gen_s), while
// its weight blocks are constants: the weight copies of the first min(NKB, NS) units are issued as soon as their stages
// are free - typically while the epilogue warps sit in that barrier - and only the activation tiles wait for it.
// Returns false when the barrier failed.
template <int NKB, int NS>
__device__ __forceinline__ bool tc_produce_ew(TcRing& r, uint64_t* full, uint64_t* empty, uint8_t* stages, const uint8_t* act, uint32_t astride,
                                              const uint8_t* wsrc, uint32_t wstride, uint32_t wbytes, uint32_t abytes, int rot,
                                              const unsigned int* gen_s, unsigned int need) {
  constexpr int NPRE = NKB < NS ? NKB : NS;
  const uint32_t total = abytes + wbytes;
  r.stage = 0;
  {
    TcRing e = r;
    int kb = rot;
#pragma unroll 1
    for (int i = 0; i < NPRE; ++i) {
      mbar_wait(&empty[e.stage], e.phase() ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[e.stage], total);
        bulk_g2s(stages + (size_t)e.stage * TC_STAGE_BYTES + TC_STAGE_W, wsrc + (size_t)kb * wstride, wbytes, &full[e.stage]);
      }
      __syncwarp();
      e.template advance<NS>();
      if (++kb == NKB) kb = 0;
    }
  }
  if (!tc_wait_gen(gen_s, need)) return false;
  fence_proxy_async();
  int kb = rot;
#pragma unroll 1
  for (int i = 0; i < NKB; ++i) {
    uint8_t* st = stages + (size_t)r.stage * TC_STAGE_BYTES;
    if (i < NPRE) {
      if (elect_one()) bulk_g2s(st, act + (size_t)kb * astride, abytes, &full[r.stage]);
    } else {
      mbar_wait(&empty[r.stage], r.phase() ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[r.stage], total);
        bulk_g2s(st, act + (size_t)kb * astride, abytes, &full[r.stage]);
        bulk_g2s(st + TC_STAGE_W, wsrc + (size_t)kb * wstride, wbytes, &full[r.stage]);
      }
    }
    __syncwarp();
    r.template advance<NS>();
    if (++kb == NKB) kb = 0;
  }
  return true;
}

// Phase C producer: like tc_produce_ew, but the activation tile of k-block kb is copied as soon as counter kbc[kb] has reached
// `want` (the CTAs that write that k-block of the h1 image have published it).  16 lanes poll the 16 counters at once, units are
// still issued in ring order.  The polls are relaxed L2 loads followed by a control dependency and a proxy fence (see grid_sync_pa).
template <int NKB, int NS>
__device__ __forceinline__ void tc_produce_kb(TcRing& r, uint64_t* full, uint64_t* empty, uint8_t* stages, const uint8_t* act, uint32_t astride,
                                              const uint8_t* wsrc, uint32_t wstride, uint32_t wbytes, uint32_t abytes, int rot,
                                              const unsigned int* kbc, unsigned int want) {
  static_assert(NKB <= 32, "one lane per k-block");
  constexpr int NPRE = NKB < NS ? NKB : NS;
  const uint32_t total = abytes + wbytes;
  const int lane = threadIdx.x & 31;
  r.stage = 0;
  {
    TcRing e = r;
    int kb = rot;
#pragma unroll 1
    for (int i = 0; i < NPRE; ++i) {
      mbar_wait(&empty[e.stage], e.phase() ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[e.stage], total);
        bulk_g2s(stages + (size_t)e.stage * TC_STAGE_BYTES + TC_STAGE_W, wsrc + (size_t)kb * wstride, wbytes, &full[e.stage]);
      }
      __syncwarp();
      e.template advance<NS>();
      if (++kb == NKB) kb = 0;
    }
  }
  uint32_t ready = 0u;
  int kb = rot;
#pragma unroll 1
  for (int i = 0; i < NKB; ++i) {
    if (!((ready >> kb) & 1u)) {
      const long long t0 = clock64();
      for (;;) {
        const bool ok = lane < NKB ? ld_relaxed_u32(kbc + lane * 32) >= want : true;
        ready = __ballot_sync(0xffffffffu, ok);
        if ((ready >> kb) & 1u) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
      fence_proxy_async();
    }
    uint8_t* st = stages + (size_t)r.stage * TC_STAGE_BYTES;
    if (i < NPRE) {
      if (elect_one()) bulk_g2s(st, act + (size_t)kb * astride, abytes, &full[r.stage]);
    } else {
      mbar_wait(&empty[r.stage], r.phase() ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[r.stage], total);
        bulk_g2s(st, act + (size_t)kb * astride, abytes, &full[r.stage]);
        bulk_g2s(st + TC_STAGE_W, wsrc + (size_t)kb * wstride, wbytes, &full[r.stage]);
      }
    }
    __syncwarp();
    r.template advance<NS>();
    if (++kb == NKB) kb = 0;
  }
}

__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t saddr) {  // low word of make_desc_sw128 (high word: umma_bf16_ss_lo)
  return ((saddr >> 4) & 0x3FFFu) | 0x10000u;
}

// MMA warp, one N=64 product per k16 step: D (+)= A . B^T with B = the 8 KB block at stage offset TC_STAGE_W.
template <int NKB, bool FRESH, int NS>
__device__ __forceinline__ void tc_consume(TcRing& r, uint64_t* full, uint64_t* empty, uint32_t stages_sa, uint32_t tmem_d,
                                           uint64_t* commit_done) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64);
  r.stage = 0;
  for (int i = 0; i < NKB; ++i) {
    const uint32_t acc = (FRESH && i == 0) ? 0u : 1u;
    mbar_wait(&full[r.stage], r.phase());
    tc_fence_after();
    const uint32_t st_sa = stages_sa + r.stage * (uint32_t)TC_STAGE_BYTES;
    const uint32_t ad = tc_desc_lo(st_sa), bd = ad + (TC_STAGE_W >> 4);
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (k > 0) ? 1u : acc);
      if (commit_done && i == NKB - 1) umma_commit(commit_done);
      umma_commit(&empty[r.stage]);
    }
    __syncwarp();
    r.template advance<NS>();
  }
}

// out of line (own register allocation), ring state by value
template <int NKB, int NS>
__device__ __noinline__ TcRing seg_produce(TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* act, uint32_t astride, const uint8_t* wsrc,
                                           uint32_t wstride, uint32_t wbytes, uint32_t abytes, int rot) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  tc_produce<NKB, NS>(r, full, empty, stages, act, astride, wsrc, wstride, wbytes, abytes, rot);
  return r;
}
// early-weights variant; ring.stage == 0xFFFF on return = the awaited grid barrier failed
template <int NKB, int NS>
__device__ __noinline__ TcRing seg_produce_ew(TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* act, uint32_t astride, const uint8_t* wsrc,
                                              uint32_t wstride, uint32_t wbytes, uint32_t abytes, int rot, const unsigned int* gen_s, unsigned int need) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  if (!tc_produce_ew<NKB, NS>(r, full, empty, stages, act, astride, wsrc, wstride, wbytes, abytes, rot, gen_s, need)) r.stage = 0xFFFFu;
  return r;
}
template <int NKB, int NS>
__device__ __noinline__ TcRing seg_produce_kb(TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* act, uint32_t astride, const uint8_t* wsrc,
                                              uint32_t wstride, uint32_t wbytes, uint32_t abytes, int rot, const unsigned int* kbc, unsigned int want) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  tc_produce_kb<NKB, NS>(r, full, empty, stages, act, astride, wsrc, wstride, wbytes, abytes, rot, kbc, want);
  return r;
}
template <int NKB, bool FRESH, int NS>
__device__ __noinline__ TcRing seg_consume(TcRing r, uint64_t* full, uint32_t stages_sa, uint32_t tmem_d, uint64_t* commit_done) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  tc_consume<NKB, FRESH, NS>(r, full, empty, stages_sa, tmem_d, commit_done);
  return r;
}

// Operand images are written with generic-proxy global stores and read by other CTAs' bulk copies (async proxy).  The
// READER always executes fence.proxy.async between passing the grid barrier and issuing its copies; the WRITER-side proxy fence
// before the barrier is redundant on this path (the barrier's release already orders the stores at gpu scope, the copy
// engine reads L2) and costs a store drain (~1 k cycles) per phase, so it is off by default.  -DGSTK_WRITER_PROXY_FENCE=1
// restores it; tests/test_decoder_bf16_gpu.py::test_bf16_repeatable checks 256 x 300 free-running steps bit for bit.
#ifndef GSTK_WRITER_PROXY_FENCE
#define GSTK_WRITER_PROXY_FENCE 0
#endif
__device__ __forceinline__ void writer_proxy_fence() {
#if GSTK_WRITER_PROXY_FENCE
  fence_proxy_async();
#endif
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(x) = 0.5 + 0.5 tanh(x / 2): ONE MUFU op (MUFU.TANH runs at the full 16 results/clk/SM, tools/ubench_mufu.cu) instead of
// ex2 + rcp; |error| <= 2.5e-4, inside the bf16 mode's budget (the fp32 kernel keeps expf)
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// LSTM point-wise update for one batch row and this CTA's 8 units of one cell; v[gate*8+u] = x.W + h.U
__device__ __forceinline__ void tc_epilogue_row(const float (&v)[32], const float* bias_s, float (&c)[8], int row,
                                                int cta, int MT, __nv_bfloat16* act_out, float* h_out /*[B][U] row base*/) {
  float h[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float zi = v[u] + bias_s[u], zf = v[8 + u] + bias_s[8 + u];
    const float zg = v[16 + u] + bias_s[16 + u], zo = v[24 + u] + bias_s[24 + u];
    c[u] = sigmoid_fast(zf) * c[u] + sigmoid_fast(zi) * tanh_fast(zg);
    h[u] = sigmoid_fast(zo) * tanh_fast(c[u]);
  }
  // bf16 operand image: k-block = (8*cta)/64, chunk = cta % 8 (swizzled with the row)
  __nv_bfloat162 pk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) pk[i] = __floats2bfloat162_rn(h[2 * i], h[2 * i + 1]);
  const int kb = cta >> 3, mt = row >> 7, r = row & 127;
  const int chunk = (cta & 7) ^ (r & 7);
  uint4* dst = reinterpret_cast<uint4*>(act_out + ((size_t)(kb * MT + mt) * 128 + r) * 64 + chunk * 8);
  *dst = *reinterpret_cast<const uint4*>(pk);
  if (h_out) {   // fp32 copy: only the generic phase A and the final-state hand-over read it
    float4* hf = reinterpret_cast<float4*>(h_out + cta * 8);
    hf[0] = make_float4(h[0], h[1], h[2], h[3]);
    hf[1] = make_float4(h[4], h[5], h[6], h[7]);
  }
}


// ---------------------------------------------------------------------------------------------
// Phase A fast path (SMA, attention size 128, <= 2 utterances per CTA, handled together).
// Dense layers: W^T stored [N][Kp] bf16; warp w owns columns w, w+15, ...; a lane reads 8
// consecutive k (16 B) so that one warp load covers 256 k of one column (fully coalesced), CB columns
// x KC k-chunks are in flight per warp; fp32 accumulate; 5-step shuffle reduction per column.
// ---------------------------------------------------------------------------------------------
constexpr int FA_WARPS = TC_PA_THREADS / 32;  // 15
constexpr int PROF_SLOTS = 16;
__device__ __forceinline__ void prof_tick(unsigned long long* prof_s, int slot) {
  // prof_s[PROF_SLOTS] = last timestamp; only thread 0 of the CTA records
  if (prof_s && threadIdx.x == 0) {
    const unsigned long long now = (unsigned long long)clock64();
    prof_s[slot] += now - prof_s[PROF_SLOTS];
    prof_s[PROF_SLOTS] = now;
  }
}

// ---- dense layers of phase A on warp-level tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32) ----------
// y[n][u] = sum_k W[k][n] * act[u][k]:  A operand = 16 output features x 16 k (weights), B operand = 16 k x 8
// "columns" of which the first NU are the CTA's utterances (the rest are zero), D = 16 features x 8.
// The weights are pre-packed on the host in FRAGMENT ORDER: tile (ft, kt) is 32 lanes x 16 B, lane (g = lane/4,
// t = lane%4) holds {W(g,2t) W(g,2t+1) W(g+8,2t) W(g+8,2t+1) W(g,2t+8) W(g,2t+9) W(g+8,2t+8) W(g+8,2t+9)}
// (row = feature within the tile, col = k within the tile), i.e. registers a0..a3 of the PTX fragment layout.
// One LDG.128 per lane fetches a whole tile (512 B per warp, coalesced); 16 tiles are in flight per warp.
__device__ __forceinline__ void mma_16816_bf16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// Phase-A dense weights: a ring stage holds, for ALL feature tiles of a layer, KTS consecutive k16 tiles:
// [ft][kt_in_stage][32 lanes][16 B].  Warp w owns feature tiles w and w + 10 and accumulates them over the
// stages in registers, so no cross-warp reduction is needed.
struct FaW {
  uint32_t base;     // byte offset of the layer in the image
  int NF, KT, KTS;   // feature tiles, k16 tiles, k16 tiles per stage
  __host__ __device__ constexpr int nst() const { return (KT + KTS - 1) / KTS; }
  __host__ __device__ constexpr uint32_t stride() const { return (uint32_t)NF * KTS * 512u; }
};
enum { FA_L_PROJ = 0, FA_L_PRE0 = 1, FA_L_PRE1 = 2, FA_L_QUERY = 3 };
__host__ __device__ constexpr FaW fa_wlayer(int layer, int PD, int mel, int P0, int P1, int A, int HC) {
  const int NFs[4] = {(PD + 15) / 16, P0 / 16, P1 / 16, A / 16};
  const int KTs[4] = {HC / 16, (mel + 15) / 16, P0 / 16, P1 / 16};
  FaW L{};
  uint32_t base = 0;
  for (int l = 0; l < 4; ++l) {
    L.NF = NFs[l]; L.KT = KTs[l]; L.KTS = fa_kts(NFs[l]); L.base = base;
    if (l == layer) break;
    base += (uint32_t)L.nst() * L.stride();
  }
  return L;
}
// the fast path only runs with the reference's default widths (bf16_fast_a): its layer table is a compile-time constant
constexpr int FA_PD = 81, FA_MEL = 80, FA_P = 256, FA_A = 128;
constexpr FaW FA_LP = fa_wlayer(0, FA_PD, FA_MEL, FA_P, FA_P, FA_A, TC_U + 128), FA_L0 = fa_wlayer(1, FA_PD, FA_MEL, FA_P, FA_P, FA_A, TC_U + 128),
              FA_L1 = fa_wlayer(2, FA_PD, FA_MEL, FA_P, FA_P, FA_A, TC_U + 128), FA_LQ = fa_wlayer(3, FA_PD, FA_MEL, FA_P, FA_P, FA_A, TC_U + 128);
constexpr int FA_NST_P = FA_LP.nst(), FA_NST_REST = FA_L0.nst() + FA_L1.nst() + FA_LQ.nst();
// number of weight stages consumed before the phase-A call of step t (projection runs for t > 0, the rest for t < T)
__device__ __forceinline__ uint32_t fa_stages_before(int t, int nst_p, int nst_rest) {
  return (uint32_t)t * (uint32_t)nst_rest + (uint32_t)(t > 0 ? t - 1 : 0) * (uint32_t)nst_p;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- producer side of the dense CTAs' weight ring -------------------------------------------------------------
// The weight stages of one decoder step form the fixed sequence [projection (of step t-1) | prenet0 | prenet1 | query];
// step 0 has no projection, step T only the projection.  The copy warp walks that sequence for the WHOLE decode in
// one go and is throttled only by the ring's empty barriers, i.e. it always runs DA_WSTAGES stages ahead of the
// consumers - across the grid barriers too.
constexpr int FA_NST_0 = FA_L0.nst(), FA_NST_1 = FA_L1.nst();
struct FaStage { uint32_t off, bytes; };
template <int LAYER>
__device__ __forceinline__ FaStage fa_layer_stage(int si) {
  constexpr FaW L = fa_wlayer(LAYER, FA_PD, FA_MEL, FA_P, FA_P, FA_A, TC_U + 128);
  constexpr int NST = L.nst(), LASTK = L.KT - (NST - 1) * L.KTS;
  const int kts = (si == NST - 1) ? LASTK : L.KTS;
  return FaStage{L.base + (uint32_t)si * L.stride(), (uint32_t)(L.NF * kts) * 512u};
}
__device__ __forceinline__ FaStage fa_step_stage(int j) {  // j in [0, FA_NST_P + FA_NST_REST)
  if (j < FA_NST_P) return fa_layer_stage<FA_L_PROJ>(j);
  j -= FA_NST_P;
  if (j < FA_NST_0) return fa_layer_stage<FA_L_PRE0>(j);
  j -= FA_NST_0;
  if (j < FA_NST_1) return fa_layer_stage<FA_L_PRE1>(j);
  return fa_layer_stage<FA_L_QUERY>(j - FA_NST_1);
}
// `exit_s` (early stop): the consumers stop taking stages once every utterance has stopped; the producer then drains the copies
// it has in flight (a CTA must not exit with bulk copies into its shared memory pending) and leaves.
__device__ __noinline__ void da_produce_all(int T, uint64_t* wfull, uint64_t* wempty, uint8_t* wstages, const uint8_t* wimg, const volatile int* exit_s) {
  uint32_t cnt = 0;
  for (int t = 0; t <= T; ++t) {
    const int j0 = t > 0 ? 0 : FA_NST_P, j1 = t < T ? FA_NST_P + FA_NST_REST : FA_NST_P;
    for (int j = j0; j < j1; ++j, ++cnt) {
      const uint32_t st = cnt % DA_WSTAGES, ph = (cnt / DA_WSTAGES) & 1u;
      const FaStage g = fa_step_stage(j);
      if (!mbar_try_wait(&wempty[st], ph ^ 1u)) {
        const long long t0 = clock64();
        while (!mbar_try_wait(&wempty[st], ph ^ 1u)) {
          if (*exit_s) {
            for (uint32_t c = cnt > DA_WSTAGES ? cnt - DA_WSTAGES : 0u; c < cnt; ++c) mbar_wait(&wfull[c % DA_WSTAGES], (c / DA_WSTAGES) & 1u);
            return;
          }
          if (clock64() - t0 > 4000000000LL) __trap();
        }
      }
      if (elect_one()) {
        mbar_arrive_expect_tx(&wfull[st], g.bytes);
        bulk_g2s(wstages + (size_t)st * FA_WSTAGE_BYTES, wimg + g.off, g.bytes, &wfull[st]);
      }
      __syncwarp();
    }
  }
}

// ---- consumer side: y[n][u] = sum_k W[k][n] act[u][k] for the <= 16 utterances of this dense CTA ----------------
constexpr int FA_HC = TC_U + 128;   // [h2 || ctx]
constexpr int DA_HCP = FA_HC + 8;   // padded row stride (bf16) of the layer-input buffer: conflict-free B-fragment loads
// A layer's work is cut into NF * KSPLIT units (feature tile, k-slice of every stage), UPW = NF * KSPLIT / 8 per warp: warp
// w < DA_MMA_WARPS owns k-slice w / (8 / KSPLIT) and feature tiles (w % (8 / KSPLIT)) * UPW + [0, UPW), so the four SM
// sub-partitions issue the same number of mma.sync, the B fragments (activations) are loaded once per k-tile for all of the
// warp's feature tiles, and every loop bound is a compile-time constant (loads hoisted, mma chains interleaved).
// NT = utterance n-tiles (1: <= 8 utterances, 2: <= 16).
// out: fp32 [KSPLIT][DA_MAXU utterance slots][PSTR] partial sums (PSTR = 4 mod 32: conflict-free fragment stores).
template <int NT, int NF, int KN /* k-tiles in this stage */, int KSPLIT, int UPW>
__device__ __forceinline__ void da_stage_mma(float (&d)[UPW][NT][4], const uint4* __restrict__ tiles, const __nv_bfloat16* bp /* + kt0 */,
                                             int qk, int ft0, int lane) {
  constexpr int KQ = (KN + KSPLIT - 1) / KSPLIT;   // k-tiles of one unit in this stage
  const int k0 = qk * KQ;
  uint32_t b[KQ][NT][2];
#pragma unroll
  for (int ki = 0; ki < KQ; ++ki)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const __nv_bfloat16* bq = bp + (k0 + ki) * 16 + nt * 8 * DA_HCP;
      b[ki][nt][0] = *reinterpret_cast<const uint32_t*>(bq);
      b[ki][nt][1] = *reinterpret_cast<const uint32_t*>(bq + 8);
    }
  uint4 a[UPW][KQ];
#pragma unroll
  for (int s = 0; s < UPW; ++s)
#pragma unroll
    for (int ki = 0; ki < KQ; ++ki) a[s][ki] = tiles[(size_t)((ft0 + s) * KN + k0 + ki) * 32 + lane];
#pragma unroll
  for (int ki = 0; ki < KQ; ++ki)
#pragma unroll
    for (int s = 0; s < UPW; ++s)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_16816_bf16(d[s][nt], a[s][ki], b[ki][nt][0], b[ki][nt][1]);   // UPW * NT independent chains
}

// epi(k-slice, feature, utterance slot, value) consumes the accumulators (fused layer epilogue)
template <int NT, int NF, int KT, int KSPLIT, class Epi>
__device__ __forceinline__ void da_consume_layer(uint32_t& cnt, uint64_t* wfull, uint64_t* wempty, const uint8_t* wstages,
                                                 const __nv_bfloat16* act_s, int wid, int lane, Epi epi) {
  constexpr int KTS = fa_kts(NF);
  constexpr int NST = (KT + KTS - 1) / KTS, LASTK = KT - (NST - 1) * KTS;
  constexpr int WPK = DA_MMA_WARPS / KSPLIT, UPW = NF / WPK;   // warps per k-slice, feature tiles per warp
  static_assert(NF % WPK == 0 && KTS % KSPLIT == 0 && (KSPLIT == 1 || LASTK == KTS), "dense layer does not tile evenly");
  const int g = lane >> 2, t = lane & 3;
  const int qk = wid / WPK, ft0 = (wid - qk * WPK) * UPW;
  const __nv_bfloat16* arow = act_s + g * DA_HCP + 2 * t;
  float d[UPW][NT][4] = {};
#pragma unroll 1
  for (int si = 0; si < NST; ++si, ++cnt) {
    const uint32_t st = cnt % DA_WSTAGES, ph = (cnt / DA_WSTAGES) & 1u;
    mbar_wait(&wfull[st], ph);
    const uint4* tiles = reinterpret_cast<const uint4*>(wstages + (size_t)st * FA_WSTAGE_BYTES);   // [ft][kn][32 lanes][16 B]
    if (si < NST - 1 || LASTK == KTS) da_stage_mma<NT, NF, KTS, KSPLIT, UPW>(d, tiles, arow + si * KTS * 16, qk, ft0, lane);
    else da_stage_mma<NT, NF, LASTK, KSPLIT, UPW>(d, tiles, arow + si * KTS * 16, qk, ft0, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&wempty[st]);  // this warp is done with the stage
  }
  // D fragment: rows = features g, g + 8 of the tile; columns = utterances nt*8 + 2t, 2t + 1
#pragma unroll
  for (int s = 0; s < UPW; ++s)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int f = (ft0 + s) * 16 + g, u = nt * 8 + 2 * t;
      epi(qk, f, u, d[s][nt][0]);
      epi(qk, f, u + 1, d[s][nt][1]);
      epi(qk, f + 8, u, d[s][nt][2]);
      epi(qk, f + 8, u + 1, d[s][nt][3]);
    }
}

// ReLU + dropout (keep mask: external tensor or one Philox call per 4 consecutive units) for features n4..n4+3
__device__ __forceinline__ void fa_relu_dropout4(float (&v)[4], const DecParams& p, const float* keep_ext, int stream, unsigned int step_id,
                                                 unsigned int row_id, int n4, bool drop) {
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
  if (!drop) return;
  float k[4];
  if (p.rng_mode == 1) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(keep_ext + n4));
    k[0] = m.x; k[1] = m.y; k[2] = m.z; k[3] = m.w;
  } else {
    const uint4 r = philox4x32_10(make_uint4((unsigned int)n4 >> 2, step_id, row_id, (unsigned int)stream),
                                  make_uint2((unsigned int)p.seed, (unsigned int)(p.seed >> 32)));
    const float sc = 5.9604644775390625e-08f;
    k[0] = (float)(r.x >> 8) * sc >= p.drop_rate ? 1.f : 0.f;
    k[1] = (float)(r.y >> 8) * sc >= p.drop_rate ? 1.f : 0.f;
    k[2] = (float)(r.z >> 8) * sc >= p.drop_rate ? 1.f : 0.f;
    k[3] = (float)(r.w >> 8) * sc >= p.drop_rate ? 1.f : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = v[i] * k[i] * p.drop_scale;
}

constexpr int FA_B_P = 0, FA_B_0 = 96, FA_B_1 = 96 + 256, FA_B_Q = 96 + 512, FA_B_V = 96 + 512 + 128, FA_BIAS_N = 96 + 512 + 256;
// dense CTA scratch: layer input (bf16 [DA_MAXU][DA_HCP]) | layer output sums | biases
constexpr int DA_PSTR_P = 100, DA_KSPLIT_P = 4;   // projection: row stride of the partial sums (4 mod 32), k-slices
constexpr int DA_COL_P0 = 256, DA_COL_P1 = 512;   // layer outputs go to other columns of the (dead) [h2 || ctx] rows: no in-place hazard
constexpr int DA_ACT_BYTES = DA_MAXU * DA_HCP * 2;
constexpr int DA_PART_BYTES = DA_MAXU * 4 * DA_KSPLIT_P * DA_PSTR_P;
constexpr int DA_KEEP_BYTES = 2 * DA_MAXU * FA_P;   // pre-drawn dropout keep flags (one byte per unit) of the two prenet layers
constexpr int DA_SCRATCH_BYTES = DA_ACT_BYTES + DA_PART_BYTES + FA_BIAS_N * 4 + DA_KEEP_BYTES;
static_assert(DA_ACT_BYTES % 16 == 0, "dense activation buffer must keep 16 B alignment");
// generic (BMA / LSA / other widths) phase A: fp32 path of decoder_fp32.cuh on the 10 phase-A warps
__device__ __noinline__ void phase_a_generic(const DecParams& p, float* scratch, int b, int t) {
  PhaseASmem s;
  float* f = scratch;
  auto take = [&](int n) { float* r = f; f += (n + 3) & ~3; return r; };
  s.x = take(p.mel); s.y = take(p.PD); s.hc = take(p.U1 + p.A); s.p0 = take(p.P0); s.p1 = take(p.P1);
  s.q = take(p.A); s.e = take(p.Tv); s.al = take(p.Tv); s.prev = take(p.Tv); s.src = take(p.Tv);
  s.red = take(DEC_THREADS); s.scal = take(8);
  phase_a_utt<TC_PA_THREADS>(p, s, b, t);
}

// ---------------------------------------------------------------------------------------------
// Phase A1 on a dense CTA: projection of step t-1 (Taco2.py:112-118) -> decoder input (Taco2.py:183-187) -> prenet
// (Taco2.py:270-283, dropout always on) -> query projection (Steps.py:122) for utterances [b0, b0 + nu).
// Inputs come from the bf16 operand images the LSTM phases use (h2(t-1): actH2, ctx(t-1): k-blocks 4,5 of actX), outputs
// go to out_mel / out_stop, the p part of actX (LSTMCell-0 input) and qbuf (queries for the attention CTAs).
// ---------------------------------------------------------------------------------------------
// Draw the prenet dropout keep flags of step t for this dense CTA's utterances into shared memory (one Philox call per 4
// units, same counters as fa_relu_dropout4).  Runs off the critical path (phase B of step t-1): the draws are counter-based.
__device__ __forceinline__ void dense_keep_fill(const DecParams& p, uint8_t* scratch, int b0, int nu, int t) {
  if (!(p.rng_mode != 0 && p.drop_rate > 0.f)) return;
  uint32_t* keep = reinterpret_cast<uint32_t*>(scratch + DA_ACT_BYTES + DA_PART_BYTES + FA_BIAS_N * 4);
  const unsigned int step_id = p.step_offset + (unsigned int)t;
  for (int i = threadIdx.x; i < 2 * nu * (FA_P / 4); i += TC_PA_THREADS) {
    const int layer = i / (nu * (FA_P / 4)), r = i - layer * nu * (FA_P / 4), u = r / (FA_P / 4), n4 = (r - u * (FA_P / 4)) * 4, b = b0 + u;
    uint32_t f;
    if (p.rng_mode == 1) {
      const float* kp = (layer ? p.keep1 : p.keep0) + ((size_t)t * p.rngB + p.rng_b0 + b) * FA_P + n4;
      const float4 m = __ldg(reinterpret_cast<const float4*>(kp));
      f = (m.x != 0.f ? 1u : 0u) | (m.y != 0.f ? 0x100u : 0u) | (m.z != 0.f ? 0x10000u : 0u) | (m.w != 0.f ? 0x1000000u : 0u);
    } else {
      const uint4 rr = philox4x32_10(make_uint4((unsigned int)n4 >> 2, step_id, p.row_offset + b, (unsigned int)(layer ? STREAM_KEEP1 : STREAM_KEEP0)),
                                     make_uint2((unsigned int)p.seed, (unsigned int)(p.seed >> 32)));
      const float sc = 5.9604644775390625e-08f;
      f = ((float)(rr.x >> 8) * sc >= p.drop_rate ? 1u : 0u) | ((float)(rr.y >> 8) * sc >= p.drop_rate ? 0x100u : 0u) |
          ((float)(rr.z >> 8) * sc >= p.drop_rate ? 0x10000u : 0u) | ((float)(rr.w >> 8) * sc >= p.drop_rate ? 0x1000000u : 0u);
    }
    keep[(layer * DA_MAXU + u) * (FA_P / 4) + (n4 >> 2)] = f;
  }
}
// ReLU + dropout with pre-drawn keep flags (4 units)
__device__ __forceinline__ void da_relu_keep4(float (&v)[4], uint32_t f, float scale, bool drop) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = fmaxf(v[i], 0.f);
    if (drop) v[i] = ((f >> (8 * i)) & 1u) ? v[i] * scale : 0.f;
  }
}

template <int NT>
__device__ __noinline__ void dense_a(const DecParams& p, const Bf16Params& q, uint8_t* scratch, unsigned long long* prof, uint64_t* wfull,
                                     const uint8_t* wstages, int b0, int nu, int t) {
  __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(scratch);
  float* part = reinterpret_cast<float*>(scratch + DA_ACT_BYTES);
  const float* bias = reinterpret_cast<const float*>(scratch + DA_ACT_BYTES + DA_PART_BYTES);
  const uint32_t* keep = reinterpret_cast<const uint32_t*>(scratch + DA_ACT_BYTES + DA_PART_BYTES + FA_BIAS_N * 4);
  uint64_t* wempty = wfull + DA_WSTAGES;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool mma_w = wid < DA_MMA_WARPS;
  if (t == 0) dense_keep_fill(p, scratch, b0, nu, 0);   // later steps: drawn during phase B of the step before
  const int melp = (p.mel + 15) & ~15;
  uint32_t wcnt = fa_stages_before(t, FA_NST_P, FA_NST_REST);  // position in the weight ring
  const unsigned int step_id = p.step_offset + (unsigned int)t;
  const bool drop = p.rng_mode != 0 && p.drop_rate > 0.f;
  if (t > 0) {
    // ---- [h2(t-1) || ctx(t-1)] rows of this CTA's utterances: 18 k-blocks x 8 swizzled 16 B chunks per utterance,
    // up to 6 loads in flight per thread (one L2 round trip for <= 13 utterances)
    const int nchunk = nu * 144;
#pragma unroll 1
    for (int i0 = tid; i0 < nchunk; i0 += 6 * TC_PA_THREADS) {
      uint4 v[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int i = i0 + k * TC_PA_THREADS;
        if (i < nchunk) {
          const int u = i / 144, c = i - u * 144, kbx = c >> 3, ch = c & 7;
          const int b = b0 + u, mt = b >> 7, r = b & 127;
          const __nv_bfloat16* img = kbx < TC_NKB_H ? q.actH2 + ((size_t)(kbx * p.MT + mt) * 128 + r) * 64
                                                    : q.actX + ((size_t)((kbx - TC_NKB_H + 4) * p.MT + mt) * 128 + r) * 64;
          v[k] = __ldcg(reinterpret_cast<const uint4*>(img) + ch);
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int i = i0 + k * TC_PA_THREADS;
        if (i < nchunk) {
          const int u = i / 144, c = i - u * 144, kbx = c >> 3, ch = c & 7;
          const int r = (b0 + u) & 127;
          *reinterpret_cast<uint4*>(act + u * DA_HCP + kbx * 64 + ((ch ^ (r & 7)) << 3)) = v[k];
        }
      }
    }
    pa_sync<TC_PA_THREADS>();
    prof_tick(prof, 8);
    if (mma_w)
      da_consume_layer<NT, 6, FA_HC / 16, DA_KSPLIT_P>(wcnt, wfull, wempty, wstages, act, wid, lane,
                                                       [&](int qk, int f, int u, float v) { part[(qk * DA_MAXU + u) * DA_PSTR_P + f] = v; });
    pa_sync<TC_PA_THREADS>();
    prof_tick(prof, 9);
    // output pass; in free-running mode the last of the r frames is also the next decoder input (Taco2.py:183-187)
#pragma unroll 1
    for (int i0 = tid; i0 < nu * 96; i0 += 4 * TC_PA_THREADS) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = min(i0 + k * TC_PA_THREADS, nu * 96 - 1), u = i / 96, n = i - u * 96;
        float a = bias[FA_B_P + n];
#pragma unroll
        for (int kk = 0; kk < DA_KSPLIT_P; ++kk) a += part[(kk * DA_MAXU + u) * DA_PSTR_P + n];
        v[k] = a;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * TC_PA_THREADS, u = i / 96, n = i - u * 96;
        if (i < nu * 96 && n < FA_PD) {
          if (n < FA_PD - 1) {
            if (p.out_mel) p.out_mel[((size_t)(b0 + u) * p.To + (t - 1)) * (FA_PD - 1) + n] = v[k];
            if (p.mode == 0) act[u * DA_HCP + n] = __float2bfloat16(v[k]);
          } else {
            if (p.out_stop) p.out_stop[(size_t)(b0 + u) * p.To + (t - 1)] = v[k];
            note_stop(p, b0 + u, t - 1, v[k]);
          }
        }
      }
    }
  }
  if (t == p.T) return;
  // ---- decoder input (Taco2.py:183-187) when it is not the projected frame: teacher frame, or the initial frame at t == 0
  if (p.mode == 1 || t == 0) {
    for (int i = tid; i < nu * FA_MEL; i += TC_PA_THREADS) {
      const int u = i / FA_MEL, n = i - u * FA_MEL;
      float x = 0.f;
      if (p.mode == 1) x = __ldg(p.teacher + (size_t)(b0 + u) * p.ts_b + (size_t)t * p.ts_t + n);
      else if (p.init_mel) x = __ldg(p.init_mel + (size_t)(b0 + u) * p.mel + n);
      act[u * DA_HCP + n] = __float2bfloat16(x);
    }
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(prof, 10);
  // ---- prenet layers: bias + ReLU + dropout (pre-drawn keep flags) fused into the mma epilogue, bf16 output into the columns
  // the next layer reads
  const uint8_t* keepb = reinterpret_cast<const uint8_t*>(keep);
  const float dscale = p.drop_scale;
  if (mma_w)
    da_consume_layer<NT, 16, 5, 1>(wcnt, wfull, wempty, wstages, act, wid, lane, [&](int, int f, int u, float v) {
      float y = fmaxf(v + bias[FA_B_0 + f], 0.f);
      if (drop) y = keepb[u * FA_P + f] ? y * dscale : 0.f;
      act[u * DA_HCP + DA_COL_P0 + f] = __float2bfloat16(y);
    });
  pa_sync<TC_PA_THREADS>();
  prof_tick(prof, 11);
  if (mma_w)
    da_consume_layer<NT, 16, 16, 1>(wcnt, wfull, wempty, wstages, act + DA_COL_P0, wid, lane, [&](int, int f, int u, float v) {
      float y = fmaxf(v + bias[FA_B_1 + f], 0.f);
      if (drop) y = keepb[(DA_MAXU + u) * FA_P + f] ? y * dscale : 0.f;
      act[u * DA_HCP + DA_COL_P1 + f] = __float2bfloat16(y);
    });
  pa_sync<TC_PA_THREADS>();
  prof_tick(prof, 13);
  // ---- p1 -> actX image (LSTMCell-0 input), 16 B swizzle chunks, by the warps that do not issue mma while the query layer runs
  if (!mma_w) {
    for (int i = tid - DA_MMA_WARPS * 32; i < nu * (FA_P / 8); i += TC_PA_THREADS - DA_MMA_WARPS * 32) {
      const int u = i / (FA_P / 8), c = i - u * (FA_P / 8);
      const uint4 w = *reinterpret_cast<const uint4*>(act + u * DA_HCP + DA_COL_P1 + 8 * c);
      *reinterpret_cast<uint4*>(p.actX + act_elem_index(p.MT, b0 + u, 8 * c)) = w;
    }
  }
  // ---- query projection (Steps.py:122) -> qbuf
  if (mma_w)
    da_consume_layer<NT, 8, 16, 1>(wcnt, wfull, wempty, wstages, act + DA_COL_P1, wid, lane, [&](int, int f, int u, float v) {
      if (u < nu) q.qbuf[(size_t)(b0 + u) * FA_A + f] = v + bias[FA_B_Q + f];
    });
  prof_tick(prof, 15);
}

// LSTM epilogue of one cell for one (batch row, 8-unit half): wait for the accumulators, point-wise update with the cell state
// in TMEM, publish h.  Out of line so that its 50-odd live registers do not inflate the kernel body (a spill there is an L1
// miss after every grid barrier).
__device__ __noinline__ void lstm_epilogue(uint64_t* d_full, uint32_t parity, uint32_t t_acc, uint32_t t_c, const float* bias, bool row_ok,
                                           int row, int ub, int MT, __nv_bfloat16* act_out, float* h_out) {
  mbar_wait_backoff(d_full, parity);
  tc_fence_after();
  float v[32], c[8];
  tmem_ld32(t_acc, v);
  tmem_ld8(t_c, c);
  if (row_ok) tc_epilogue_row(v, bias, c, row, ub, MT, act_out, h_out);
  tmem_st8(t_c, c);
  tc_fence_before();
  writer_proxy_fence();
}

// Grid barrier executed by the phase-A / epilogue warps only (threads [0, TC_PA_THREADS)).  The copy and MMA warps never
// join it: they follow the barrier generation thread 0 publishes in shared memory (tc_wait_gen), so a long operand
// stream (h2 . U2) keeps running while the other warps cross barriers.
__device__ __forceinline__ bool grid_sync_pa(GridBarrier* gb, unsigned int nblocks, unsigned int& gen, int* ok_s, unsigned int* gen_s) {
  pa_sync<TC_PA_THREADS>();
  if (threadIdx.x == 0) {
    const unsigned int target = ++gen;
    int ok = 1;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&gb->count) : "memory");
    // The poll is a RELAXED (strong, L2) load on purpose: an acquire makes ptxas emit CCTL.IVALL, and with L1 invalidated every
    // stack slot / spilled register touched after the barrier costs an L2 round trip.  Nothing produced by another CTA is ever
    // read through L1 in this kernel (ld.global.cg, bulk copies, constant inputs via ld.global.nc), the poll is followed by a
    // control dependency and a bar.sync, and loads issue in program order - so no L1 invalidation is needed for correctness.
    const unsigned int want = target * nblocks;
    if (ld_relaxed_u32(&gb->count) < want) {
      long long t0 = clock64();
      unsigned int spins = 0;
      while (ld_relaxed_u32(&gb->count) < want) {
        if ((++spins & 1023u) == 0u) {
          if (ld_relaxed_u32(&gb->error) != 0u || clock64() - t0 > 4000000000LL) {
            atomicExch(&gb->error, 1u);
            ok = 0;
            break;
          }
        }
      }
    }
    *ok_s = ok;
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(gen_s)), "r"(ok ? target : 0xFFFFFFFFu) : "memory");
  }
  pa_sync<TC_PA_THREADS>();
  return *ok_s != 0;
}

// ---------------------------------------------------------------------------------------------
// Phase A2 on every CTA: stepwise-monotonic attention for the CTA's <= 2 utterances (Steps.py:138-166, 215-229).
// A warp handles 32 consecutive key rows as 8 iterations of 4 rows: lane = (row group rg = lane / 8, column slab
// cs = lane % 8 of 16 columns), so one 16 B load per lane fetches four fully used 128 B segments and the projected keys
// V' stay in REGISTERS (64 per lane) from the energy pass to the context pass - V' is read once per step.
//   pass 1 (energy): per-lane partial v . tanh(q + V'[j]) over 16 columns (q, v slabs in registers), 3-step shuffle
//           reduction inside the 8-lane row group, + noise (pre-drawn into shared memory), sigmoid -> p[j] in shared memory
//   pass 2 (alignment recurrence, one thread per row): a_t[j] = a_{t-1}[j] p[j] + a_{t-1}[j-1] (1 - p[j-1])
//   pass 3 (context): per-lane a_t[j] * V'[j][slab] over the warp's rows, 2-step shuffle reduction over the row groups,
//           cross-warp sum through shared memory
// scratch (floats): alig [2 utt][2 (step parity)][Tv] resident for the whole decode | pbuf [2][Tv] | nzbuf [2][Tv] |
//                   qs [2][128] | ctxp [FA_WARPS][128]
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int att_scratch_floats(int Tv) { return ((8 * Tv + 3) & ~3) + 256 + FA_WARPS * 128; }

// N(0,1) draw for (utterance row b, key row j, step): half of a Philox block, fast-math Box-Muller (oracle: philox_normal)
__device__ __forceinline__ float att_noise(const DecParams& p, int t, int b, int j) {
  if (p.rng_mode == 1) return __ldg(p.noise + ((size_t)t * p.rngB + p.rng_b0 + b) * p.Tv + j);
  const uint4 rr = philox4x32_10(make_uint4((unsigned int)j >> 2, p.step_offset + (unsigned int)t, p.row_offset + b, (unsigned int)STREAM_NOISE),
                                 make_uint2((unsigned int)p.seed, (unsigned int)(p.seed >> 32)));
  const int w = j & 3;
  const unsigned int ua = (w & 2) ? rr.z : rr.x, ub = (w & 2) ? rr.w : rr.y;
  const float sc = 5.9604644775390625e-08f;
  const float rad = sqrtf(-2.0f * __logf(((float)(ua >> 8) + 1.0f) * sc));
  float sn, cs;
  __sincosf(6.283185307179586f * ((float)(ub >> 8) * sc), &sn, &cs);
  return rad * ((w & 1) ? sn : cs);
}
// draw the sigmoid noise of step t for the CTA's utterances into shared memory (runs off the critical path, during phase B
// of step t-1; the draws are counter-based, so when they are made does not matter)
template <int NU>
__device__ __forceinline__ void att_noise_fill(const DecParams& p, float* scratch, int b0, int t) {
  if (!(p.rng_mode != 0 && p.sigmoid_noise > 0.f)) return;
  float* nzbuf = scratch + 6 * p.Tv;
  for (int i = threadIdx.x; i < NU * p.Tv; i += TC_PA_THREADS) {
    const int u = i / p.Tv, j = i - u * p.Tv;
    nzbuf[i] = att_noise(p, t, b0 + u * (int)gridDim.x, j);
  }
}

__device__ __forceinline__ float bf16lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf16hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

// The function contains grid barrier 0 (dense layers -> attention): the projected keys are constant, so their loads are
// issued BEFORE the barrier and the L2 latency hides behind the wait.  Returns false when the barrier failed.
template <int NU, bool EARLY_LOADS>
__device__ __noinline__ bool attention_a(const DecParams& p, const Bf16Params& q, float* scratch, const float* attv_s, int b0, int t,
                                         unsigned int gen /* by value: a reference would put the caller's counter in local memory,
                                         one L1 miss per barrier */, int* ok_s, unsigned int* gen_s, unsigned long long* prof) {
  constexpr int WPU = FA_WARPS / NU;   // warps per utterance
  const int Tv = p.Tv;
  float* alig = scratch;
  float* pbuf = scratch + 4 * Tv;
  const float* nzbuf = scratch + 6 * Tv;
  float* qs = scratch + ((8 * Tv + 3) & ~3);
  float* ctxp = qs + 256;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int cur = t & 1, prv = cur ^ 1;
  const int XW = p.P1 + p.A;
  int bs[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) bs[u] = b0 + u * (int)gridDim.x;
  const float sb = __ldg(p.att_sb);
  const bool noisy = p.rng_mode != 0 && p.sigmoid_noise > 0.f;
  const int uw = NU == 2 ? wid / WPU : 0, wl = wid - uw * WPU;   // utterance and warp-within-utterance of this warp
  const int bw = bs[NU == 2 ? uw : 0];
  const int rg = lane >> 3, cs = lane & 7;
  const uint4* Vl = reinterpret_cast<const uint4*>(q.vproj_bf + (size_t)bw * Tv * 128) + 2 * cs;   // + 16 * row
  const bool single = Tv <= WPU * 32;   // one sweep: V' stays in registers between the passes
  uint4 kv[8][2];
  auto load_rows = [&](int base) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = base + 4 * i + rg;
      if (j < Tv) { kv[i][0] = __ldg(Vl + (size_t)j * 16); kv[i][1] = __ldg(Vl + (size_t)j * 16 + 1); }
      else { kv[i][0] = make_uint4(0u, 0u, 0u, 0u); kv[i][1] = kv[i][0]; }
    }
  };
  // early loads only where the CTA waits at the barrier anyway: the barrier's release fence drains the loads first, which
  // would delay the arrival of the dense CTAs (the last arrivers) by an L2 round trip
  if (EARLY_LOADS) load_rows(wl * 32);
  float qr[16], vr[16];   // this lane's 16-column slab of the query and of attention_v
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 b = *reinterpret_cast<const float4*>(attv_s + 16 * cs + 4 * c);
    vr[4 * c] = b.x; vr[4 * c + 1] = b.y; vr[4 * c + 2] = b.z; vr[4 * c + 3] = b.w;
  }
  prof_tick(prof, 0);
  if (!grid_sync_pa(p.gb, gridDim.x, gen, ok_s, gen_s)) return false;
  prof_tick(prof, 1);
  if (!EARLY_LOADS) load_rows(wl * 32);
  // the query slab straight from L2 into registers (no shared-memory staging, no block barrier on steps t > 0)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(q.qbuf + (size_t)bw * FA_A + 16 * cs) + c);
    qr[4 * c] = a.x; qr[4 * c + 1] = a.y; qr[4 * c + 2] = a.z; qr[4 * c + 3] = a.w;
  }
  if (t == 0) {  // first step of this launch: initial alignments into their resident buffers, noise of step 0
    for (int i = tid; i < NU * Tv; i += TC_PA_THREADS) {
      const int u = i / Tv, j = i - u * Tv;
      alig[(u * 2 + prv) * Tv + j] = __ldcg(p.align + ((size_t)prv * p.B + bs[u]) * Tv + j);
    }
    att_noise_fill<NU>(p, scratch, b0, 0);
    pa_sync<TC_PA_THREADS>();   // alignments, noise visible
  }
  // ---- pass 1: energies
  {
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
      // lane cs == i finishes row 4 i + rg: noise, sigmoid
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
    if (t == p.T - 1) p.align[((size_t)cur * p.B + b) * Tv + j] = a;  // final alignment for state hand-over
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
  if (tid < 128 * NU) {
    const int u = tid >> 7, n = tid & 127, b = bs[u];
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < WPU; ++w) c += ctxp[(u * WPU + w) * 128 + n];
    p.actX[act_elem_index(p.MT, b, p.P1 + n)] = __float2bfloat16(c);
    if (t == p.T - 1) {   // the fp32 context is only read after the last step (state hand-over / out_context)
      p.xin[(size_t)b * XW + p.P1 + n] = c;
      if (p.out_ctx) p.out_ctx[(size_t)b * p.A + n] = c;
    }
  }
  return true;
}

constexpr size_t TC_SCRATCH_OFF = (size_t)TC_NSTAGE * TC_STAGE_BYTES > (size_t)DA_WSTAGES * FA_WSTAGE_BYTES + DA_SCRATCH_BYTES
                                      ? (size_t)TC_NSTAGE * TC_STAGE_BYTES : (size_t)DA_WSTAGES * FA_WSTAGE_BYTES + DA_SCRATCH_BYTES;

__global__ void __launch_bounds__(TC_THREADS, 1) decoder_bf16_kernel(const __grid_constant__ DecParams p, const __grid_constant__ Bf16Params q) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ int ok_s;
  __shared__ int exit_s;   // early stop: set at barrier 0 of the step at which every utterance has stopped; all warps leave after that step
  __shared__ unsigned int gen_s;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * TC_NSTAGE_BC + 3 + 2 * DA_WSTAGES];
  __shared__ float bias_s[128];   // [cell][half*32 + gate*8 + u] of this CTA's unit group
  __shared__ __align__(16) float attv_s[128];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform (role dispatch stays on the uniform datapath)
  const int cta = blockIdx.x;
  const bool lstm_cta = cta < TC_LSTM_CTAS;
  const int ug = cta >> 1, mt_c = cta & 1;             // LSTM CTAs: unit group (16 units) and batch m-tile
  const bool lstm_act = lstm_cta && mt_c < p.MT;       // B <= 128: only the m-tile-0 CTAs have LSTM work
  const bool fast_a = q.wimgA != nullptr;
  // dense CTAs (the SMs the LSTM tiling leaves free) run the phase-A dense layers of the fast path for `nu_d` utterances each
  const int n_dense = (int)gridDim.x - TC_LSTM_CTAS;
  const int upd = fast_a ? (p.B + n_dense - 1) / n_dense : 0;
  const int bd0 = (cta - TC_LSTM_CTAS) * upd;
  const int nu_d = (fast_a && !lstm_cta) ? max(0, min(upd, p.B - bd0)) : 0;
  // shared memory: LSTM CTAs = operand ring | attention scratch; dense CTAs = weight ring | dense scratch
  uint8_t* stages = sm;                                             // TC_NSTAGE x 40 KB
  uint8_t* wstages = sm;                                            // DA_WSTAGES x 28 KB (dense CTAs)
  uint8_t* scratch_d = sm + (size_t)DA_WSTAGES * FA_WSTAGE_BYTES;
  // phase-A scratch (attention / generic) sits behind both role-specific regions: dense CTAs run the attention too
  float* scratch = reinterpret_cast<float*>(sm + TC_SCRATCH_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_NSTAGE_BC;
  uint64_t* d1_full = bars + 2 * TC_NSTAGE_BC;
  uint64_t* d2_full = bars + 2 * TC_NSTAGE_BC + 1;
  uint64_t* u1_done = bars + 2 * TC_NSTAGE_BC + 2;   // this CTA's h1(t-1) . U1 stream has read its last h1 tile
  uint64_t* wfull = bars + 2 * TC_NSTAGE_BC + 3;   // [DA_WSTAGES] full, then [DA_WSTAGES] empty

  if (nu_d > 0) {
    float* bias_c = reinterpret_cast<float*>(scratch_d + DA_ACT_BYTES + DA_PART_BYTES);
    for (int i = tid; i < FA_BIAS_N; i += TC_THREADS) {
      float v = 0.f;
      if (i < FA_B_0) { if (i < p.PD) v = __ldg(p.bp + i); }
      else if (i < FA_B_1) { if (i - FA_B_0 < p.P0) v = __ldg(p.b0 + i - FA_B_0); }
      else if (i < FA_B_Q) v = __ldg(p.b1 + i - FA_B_1);
      else if (i < FA_B_V) v = __ldg(p.bq + i - FA_B_Q);
      else v = __ldg(p.att_v + i - FA_B_V);
      bias_c[i] = v;
    }
    for (int i = tid; i < DA_ACT_BYTES / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(scratch_d)[i] = 0u;
  }
  if (fast_a && tid < 128) attv_s[tid] = __ldg(p.att_v + tid);
  // Shared-memory copies of the parameter blocks for the out-of-line phase functions: through a reference they would read
  // the kernel parameters with generic loads, and every grid barrier (acquire) invalidates L1, so each first touch after a
  // barrier cost an L2 round trip (ncu: long-scoreboard stalls on p.* / q.* reads, several per sub-phase and dependent)
  __shared__ __align__(16) DecParams p_sh;
  __shared__ __align__(16) Bf16Params q_sh;
  for (int i = tid; i < (int)(sizeof(DecParams) / 4); i += TC_THREADS) reinterpret_cast<uint32_t*>(&p_sh)[i] = reinterpret_cast<const uint32_t*>(&p)[i];
  for (int i = tid; i < (int)(sizeof(Bf16Params) / 4); i += TC_THREADS) reinterpret_cast<uint32_t*>(&q_sh)[i] = reinterpret_cast<const uint32_t*>(&q)[i];
  __shared__ unsigned long long prof_sh[PROF_SLOTS + 1];
  unsigned long long* prof_s = q.prof ? prof_sh : nullptr;
  if (tid == 0) {
    for (int i = 0; i < PROF_SLOTS; ++i) prof_sh[i] = 0;
    prof_sh[PROF_SLOTS] = (unsigned long long)clock64();
    gen_s = 0u;
    exit_s = 0;
  }
  auto prof_mark = [&](int slot) { prof_tick(prof_s, slot); };
  if (tid == 0) {
    for (int i = 0; i < TC_NSTAGE_BC; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(d1_full, 1);
    mbar_init(d2_full, 1);
    mbar_init(u1_done, 1);
    for (int i = 0; i < DA_WSTAGES; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wfull[DA_WSTAGES + i], DA_MMA_WARPS);  // every mma warp of a dense CTA releases a stage
    }
    mbar_fence_init();
  }
  if (lstm_cta && tid < 128) bias_s[tid] = __ldg(q.bias + (size_t)ug * 128 + tid);
  if (lstm_cta && wid == 0) tmem_alloc(&tmem_base_s, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int MT = p.MT;
  const bool copy_warp = wid == TC_PA_WARPS;                  // bulk-copy producer warp
  const bool prod_warp = lstm_act && copy_warp;               // LSTM operand tiles on LSTM CTAs, dense-layer weights on dense CTAs
  const bool mma_warp = lstm_act && wid == TC_PA_WARPS + 1;   // tcgen05.mma issuer
  TcRing ring;
  ring.stage = 0; ring.bits = 0;
  const uint8_t* wimg_cta = reinterpret_cast<const uint8_t*>(q.wimg) + (size_t)ug * TC_IMG_BYTES;
  const uint32_t tmem = lstm_cta ? tmem_base_s : 0u;
  const uint32_t stages_sa = smem_u32(stages);
  const int rot_x = ug % TC_NKB_X, rot_h = ug % TC_NKB_H;  // per-unit-group k-block rotation (spreads the L2 hot spot)
  // this CTA's m-tile of k-block 0 in the three operand images; k-blocks are astride apart
  const uint8_t* actX_b = reinterpret_cast<const uint8_t*>(q.actX) + (size_t)mt_c * TC_A_BYTES;
  const uint8_t* actH1_b = reinterpret_cast<const uint8_t*>(q.actH1) + (size_t)mt_c * TC_A_BYTES;
  const uint8_t* actH2_b = reinterpret_cast<const uint8_t*>(q.actH2) + (size_t)mt_c * TC_A_BYTES;
  const uint32_t astride = (uint32_t)MT * TC_A_BYTES;
  const uint32_t abytes = (uint32_t)max(0, min(128, p.B - mt_c * 128)) * 128u;
  // grid barriers per step: fast path = after A1 (dense layers), A2 (attention), B, C; generic path = after A, B, C
  // fast path: the barrier after phase B is replaced by per-k-block flags (GridBarrier::kbcnt) => 3 barriers per step as well
  const unsigned int NB = 3u;

  // ================= copy warp: runs its whole schedule without joining the grid barriers =================
  if (copy_warp) {
    if (nu_d > 0) da_produce_all(p.T, wfull, wfull + DA_WSTAGES, wstages, q.wimgA, &exit_s);
    if (prod_warp) {
      bool ok = true;
      // the h2 . U2 stream starts as soon as h2(t-1) is published (measured: delaying it until after the dense layers, so that it
      // does not compete in L2 with the dense CTAs, pushes it into the attention and costs 1.7 us per step; GSTK_DEBUG bit 2 = late)
      const bool u2_late = fast_a && (p.debug_flags & 4);
      for (int t = 0; t < p.T && ok; ++t) {
        const unsigned int g0 = (unsigned int)t * NB;
        // D1 = h1(t-1) . U1 (U1 = second half of W2|U1): its input has been complete since the barrier after phase B of step t-1,
        // so it runs in the dense-layer / attention window behind the h2 . U2 stream (started any earlier it delays the dense
        // CTAs' critical loads right after barrier 3).  (Shared memory bandwidth - every operand byte goes in by bulk copy and out to the
        // tensor core - bounds the LSTM phases, so everything that can move out of phases B and C does.)
        // D2 = h2(t-1) . U2: needs the h2 image of step t-1 (barrier after phase C of step t-1)
        if (!(p.debug_flags & 1)) {
          ring = seg_produce_ew<TC_NKB_H, TC_NSTAGE>(ring, full, stages, actH2_b, astride, wimg_cta + TC_IMG_U2, TC_B_BYTES, TC_B_BYTES, abytes, rot_h,
                                                     &gen_s, g0 + (u2_late ? 1u : 0u));
          if (!(ok = ring.stage != 0xFFFFu)) break;
        }
        ring = seg_produce<TC_NKB_H, TC_NSTAGE>(ring, full, stages, actH1_b, astride, wimg_cta + TC_IMG_WU + TC_B_BYTES, 2 * TC_B_BYTES, TC_B_BYTES, abytes, rot_h);
        // phase B: D1 += [p || ctx](t) . W1x.  Fast path: the four p k-blocks are ready after the dense layers (barrier 0), only
        // the two ctx k-blocks wait for the attention (barrier 1); the MMA warp splits its segment the same way.
        if (fast_a) {
          ring = seg_produce_ew<4, TC_NSTAGE_BC>(ring, full, stages, actX_b, astride, wimg_cta + TC_IMG_W1X, TC_B_BYTES, TC_B_BYTES, abytes, ug & 3,
                                                 &gen_s, g0 + 1);
          if (!(ok = ring.stage != 0xFFFFu)) break;
          ring = seg_produce_ew<2, TC_NSTAGE_BC>(ring, full, stages, actX_b + (size_t)4 * astride, astride, wimg_cta + TC_IMG_W1X + 4 * TC_B_BYTES,
                                                 TC_B_BYTES, TC_B_BYTES, abytes, ug & 1, &gen_s, g0 + 2);
          if (!(ok = ring.stage != 0xFFFFu)) break;
        } else {
          if (!(ok = tc_wait_gen(&gen_s, g0 + NB - 2))) break;
          fence_proxy_async();
          ring = seg_produce<TC_NKB_X, TC_NSTAGE_BC>(ring, full, stages, actX_b, astride, wimg_cta + TC_IMG_W1X, TC_B_BYTES, TC_B_BYTES, abytes, rot_x);
        }
        // phase C: D2 += h1(t) . W2: needs h1(t), i.e. the phase-B epilogues of the CTAs of this m-tile
        if (fast_a) {
          ring = seg_produce_kb<TC_NKB_H, TC_NSTAGE_BC>(ring, full, stages, actH1_b, astride, wimg_cta + TC_IMG_WU, 2 * TC_B_BYTES, TC_B_BYTES, abytes,
                                                        rot_h, &p.gb->kbcnt[mt_c * TC_NKB_H][0], 4u * (unsigned int)(t + 1));
        } else {
          ring = seg_produce_ew<TC_NKB_H, TC_NSTAGE_BC>(ring, full, stages, actH1_b, astride, wimg_cta + TC_IMG_WU, 2 * TC_B_BYTES, TC_B_BYTES, abytes,
                                                        rot_h, &gen_s, g0 + NB - 1);
          if (!(ok = ring.stage != 0xFFFFu)) break;
        }
        if (*(volatile int*)&exit_s) break;   // early stop: step t was the last one
      }
    }
  } else if (mma_warp) {
    // ================= MMA warp: follows the operand ring (full barriers) =================
    for (int t = 0; t < p.T; ++t) {
      if (!(p.debug_flags & 1)) ring = seg_consume<TC_NKB_H, true, TC_NSTAGE>(ring, full, stages_sa, tmem + TC_D2, nullptr);
      ring = seg_consume<TC_NKB_H, true, TC_NSTAGE>(ring, full, stages_sa, tmem + TC_D1, u1_done);
      if (fast_a) {   // same 4 + 2 split as the copy warp (every segment restarts at ring stage 0)
        ring = seg_consume<4, false, TC_NSTAGE_BC>(ring, full, stages_sa, tmem + TC_D1, nullptr);
        ring = seg_consume<2, false, TC_NSTAGE_BC>(ring, full, stages_sa, tmem + TC_D1, d1_full);
      } else {
        ring = seg_consume<TC_NKB_X, false, TC_NSTAGE_BC>(ring, full, stages_sa, tmem + TC_D1, d1_full);
      }
      ring = seg_consume<TC_NKB_H, false, TC_NSTAGE_BC>(ring, full, stages_sa, tmem + TC_D2, d2_full);
      if (*(volatile int*)&exit_s) break;   // early stop (written before this step's post-attention barrier, see below)
    }
  } else if (wid < TC_PA_WARPS) {
    // ================= phase-A / epilogue warps =================
    // cell state of (row, this CTA's 8 units) lives in TMEM; epilogue warps 0..4*MT-1 own one batch row per thread
    // epilogue thread = (batch row of the CTA's m-tile, 8-unit half of the unit group); ub = index of that 8-unit block
    const bool epi = lstm_act && wid < 8;
    const int erow = mt_c * 128 + (wid & 3) * 32 + lane;
    const int ub = 2 * ug + (wid >> 2);
    const bool erow_ok = epi && erow < p.B;
    // TMEM address of this epilogue thread's row: lane quarter of the warp, m-tile selects the column block
    const uint32_t t_row = tmem + ((uint32_t)((wid & 3) * 32) << 16);
    const uint32_t t_half = (uint32_t)(wid >> 2);
    if (epi) {  // initial cell states -> TMEM (they stay there for the whole decode)
      float c1[8], c2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        c1[u] = erow_ok ? __ldcg(p.c1 + (size_t)erow * TC_U + ub * 8 + u) : 0.f;
        c2[u] = erow_ok ? __ldcg(p.c2 + (size_t)erow * TC_U + ub * 8 + u) : 0.f;
      }
      tmem_st8(t_row + TC_C1 + t_half * 8u, c1);
      tmem_st8(t_row + TC_C2 + t_half * 8u, c2);
    }
    unsigned int gen = 0;
    bool alive = true;
    for (int t = 0; t <= p.T && alive; ++t) {
      if (fast_a) {
        // ---------------- phase A1 (dense CTAs): projection of step t-1, prenet, query ----------------------------
        if (nu_d > 0) {
          if (nu_d > 8) dense_a<2>(p_sh, q_sh, scratch_d, prof_s, wfull, wstages, bd0, nu_d, t);
          else dense_a<1>(p_sh, q_sh, scratch_d, prof_s, wfull, wstages, bd0, nu_d, t);
          writer_proxy_fence();  // actX stores (generic proxy) -> later bulk copies (async proxy)
        }
        if (t == p.T) break;
        // ---------------- barrier 0 + phase A2 (all CTAs): attention of the owned utterances cta, cta + grid -------
        if (cta + (int)gridDim.x < p.B) { alive = attention_a<2, true>(p_sh, q_sh, scratch, attv_s, cta, t, gen, &ok_s, &gen_s, prof_s); ++gen; }
        else if (cta < p.B && nu_d == 0) { alive = attention_a<1, true>(p_sh, q_sh, scratch, attv_s, cta, t, gen, &ok_s, &gen_s, prof_s); ++gen; }
        else if (cta < p.B) { alive = attention_a<1, false>(p_sh, q_sh, scratch, attv_s, cta, t, gen, &ok_s, &gen_s, prof_s); ++gen; }
        else {
          prof_mark(0);
          alive = grid_sync_pa(p.gb, gridDim.x, gen, &ok_s, &gen_s);
          prof_mark(1);
        }
        if (!alive) break;
        // Early stop: the stop logits of step t-1 were produced by the dense layers before barrier 0 and the counter only moves
        // there, so every CTA reads the same value here.  The step is finished (its phases are already in flight on the copy /
        // MMA warps) and everybody leaves after it.
        if (p.early_stop && tid == 0 && ld_relaxed_u32(p.stop_state) >= (unsigned int)p.B) exit_s = 1;
      } else {
        for (int b = cta; b < p.B; b += gridDim.x) phase_a_generic(p_sh, scratch, b, t);
        if (t == p.T) break;
      }
      writer_proxy_fence();  // actX stores (generic proxy) -> later bulk copies (async proxy)
      // The h1 image is overwritten by the phase-B epilogues, i.e. by any CTA that has passed this barrier.  Readers of h1(t-1)
      // are the U1 streams, which run decoupled from the barriers: a CTA only arrives here once its own stream has consumed its
      // last h1 tile (always long done in practice - the stream starts a whole A1 + A2 earlier - but now also by construction).
      // The h2 image needs no such guard: its readers (U2 streams) precede phase B's MMAs in every CTA's in-order ring and
      // every CTA finishes phase B before anyone passes the next barrier.
      if (epi && wid == 0) mbar_wait_backoff(u1_done, (uint32_t)t & 1u);
      prof_mark(2);
      if (!grid_sync_pa(p.gb, gridDim.x, gen, &ok_s, &gen_s)) { alive = false; break; }
      prof_mark(3);
      // ---------------- phase B: LSTMCell 0 epilogue -------------------------------------------------
      const bool h32 = !fast_a || t == p.T - 1;   // fast path: the fp32 copies of h1 / h2 are only read after the last step
      if (nu_d > 0 && t + 1 < p.T) dense_keep_fill(p_sh, scratch_d, bd0, nu_d, t + 1);   // dropout flags of step t+1
      if (fast_a && t + 1 < p.T) {   // off the critical path: pre-draw the attention noise of step t+1
        if (cta + (int)gridDim.x < p.B) att_noise_fill<2>(p_sh, scratch, cta, t + 1);
        else if (cta < p.B) att_noise_fill<1>(p_sh, scratch, cta, t + 1);
      }
      if (epi)
        lstm_epilogue(d1_full, (uint32_t)t & 1u, t_row + TC_D1 + t_half * 32u, t_row + TC_C1 + t_half * 8u, bias_s + t_half * 32u, erow_ok, erow, ub,
                      MT, q.actH1, h32 ? p.h1 + ((size_t)(t & 1) * p.B + erow) * TC_U : nullptr);
      prof_mark(4);
      if (fast_a) {
        // publish this CTA's 16 units of h1(t) (k-block ug / 4 of its m-tile): no grid barrier between phases B and C
        pa_sync<TC_PA_THREADS>();
        if (tid == 0 && lstm_act) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&p.gb->kbcnt[mt_c * TC_NKB_H + (ug >> 2)][0]) : "memory");
      } else if (!grid_sync_pa(p.gb, gridDim.x, gen, &ok_s, &gen_s)) { alive = false; break; }
      prof_mark(5);
      // ---------------- phase C: LSTMCell 1 epilogue -------------------------------------------------
      if (epi)
        lstm_epilogue(d2_full, (uint32_t)t & 1u, t_row + TC_D2 + t_half * 32u, t_row + TC_C2 + t_half * 8u, bias_s + 64 + t_half * 32u, erow_ok, erow,
                      ub, MT, q.actH2, h32 ? p.h2 + ((size_t)(t & 1) * p.B + erow) * TC_U : nullptr);
      prof_mark(6);
      if (!grid_sync_pa(p.gb, gridDim.x, gen, &ok_s, &gen_s)) { alive = false; break; }
      prof_mark(7);
      if (exit_s) {   // (block-uniform: written before several bar.syncs of this step)
        if (cta == 0 && tid == 0) p.stop_state[1] = (unsigned int)(p.t_base + t);   // frames 0 .. t-1 are valid
        break;
      }
    }
    if (alive) {
      if (q.prof && tid == 0)
        for (int i = 0; i < PROF_SLOTS; ++i) q.prof[(size_t)cta * PROF_SLOTS + i] = prof_sh[i];
      // final cell states (h is already in p.h1 / p.h2)
      if (epi) {
        float c1[8], c2[8];
        tmem_ld8(t_row + TC_C1 + t_half * 8u, c1);
        tmem_ld8(t_row + TC_C2 + t_half * 8u, c2);
        if (erow_ok) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            p.c1[(size_t)erow * TC_U + ub * 8 + u] = c1[u];
            p.c2[(size_t)erow * TC_U + ub * 8 + u] = c2[u];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (lstm_cta && wid == 0) tmem_dealloc(tmem_base_s, TC_TMEM_COLS);
}

// fp32 [B, 1024] (row-major) -> bf16 operand image [16][MT][128][64]; rows >= B are zero-filled
__global__ void pack_act_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int MT) {
  const size_t total = (size_t)TC_NKB_H * MT * 128 * 64;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i & 63);
    const int r = (int)((i >> 6) & 127);
    const int mt = (int)((i >> 13) % MT);
    const int kb = (int)((i >> 13) / MT);
    const int chunk = e >> 3, logical = ((chunk ^ (r & 7)) << 3) | (e & 7);
    const int b = mt * 128 + r;
    dst[i] = __float2bfloat16(b < B ? src[(size_t)b * TC_U + kb * 64 + logical] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Bf16State {
  __nv_bfloat16* wimg = nullptr;
  float* bias = nullptr;
  __nv_bfloat16* act = nullptr;  // actX | actH1 | actH2 for MT = 2
  uint8_t* wimgA = nullptr;
  __nv_bfloat16* vproj_bf = nullptr;
  size_t vproj_elems = 0;
  float* qbuf = nullptr;  // [TC_MAX_B][128]
  unsigned long long* prof = nullptr;  // [num_sms][PROF_SLOTS]
  int prof_ctas = 0;
  bool fast_a = false;
  bool ready = false;
};

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

inline bool bf16_fast_a(const GstkConfig& c) {
  // compile-time layer shapes of fa_consume_layer: the reference's default widths (Hyper_Parameters.json:4,109-121)
  return c.attention_type == GSTK_ATT_SMA && c.attention_size == 128 && c.prenet0 == 256 && c.prenet1 == 256 && c.mel_dim == 80 &&
         c.step_reduction == 1;
}

inline size_t bf16_smem_bytes(const DecParams& p, bool fast_a) {
  auto r4 = [](int n) { return (size_t)((n + 3) & ~3); };
  const size_t generic = 4 * (r4(p.mel) + r4(p.PD) + r4(p.U1 + p.A) + r4(p.P0) + r4(p.P1) + r4(p.A) + 4 * r4(p.Tv) +
                              DEC_THREADS + 8);
  const size_t attention = 4 * (size_t)att_scratch_floats(p.Tv);
  return 1024 + TC_SCRATCH_OFF + (fast_a ? attention : generic);
}

inline bool bf16_config_supported(const GstkConfig& c, std::string& why) {
  if (c.lstm0 != TC_U || c.lstm1 != TC_U) { why = "needs Tacotron2.Decoder.RNN.Size == [1024, 1024]"; return false; }
  if (c.prenet1 + c.attention_size != TC_KX) { why = "needs prenet size + attention size == 384"; return false; }
  return true;
}

inline int bf16_prepare(Bf16State& st, const GstkConfig& c, const std::map<std::string, std::vector<float>>& hw,
                        std::string& err) {
  if (st.ready) return GSTK_OK;
  const std::string d = "Decoder/Decoder_Step/RNN/";
  const std::vector<float>* src[4] = {&hw.at(d + "cell_0/kernel"), &hw.at(d + "cell_0/recurrent_kernel"),
                                      &hw.at(d + "cell_1/kernel"), &hw.at(d + "cell_1/recurrent_kernel")};
  const std::vector<float>& b0 = hw.at(d + "cell_0/bias");
  const std::vector<float>& b1 = hw.at(d + "cell_1/bias");
  std::vector<__nv_bfloat16> img((size_t)TC_UG * TC_IMG_BYTES / 2);
  std::vector<float> bias((size_t)TC_UG * 128);
  // one [32 gate rows x 64 k] SWIZZLE_128B half block: row n = gate*8 + u <-> column gate*U + ub*8 + u of the Keras kernel
  // (ub = 8-unit block index); a unit group's 64-row block = the half blocks of ub = 2 ug and 2 ug + 1 back to back
  auto put_half = [&](__nv_bfloat16* blk, const std::vector<float>& W, int ub, int kb) {
    for (int n = 0; n < 32; ++n) {
      const int gate = n >> 3, u = n & 7;
      const size_t col = (size_t)gate * TC_U + ub * 8 + u;
      for (int k = 0; k < 64; ++k) blk[sw128_offset_bytes(n, k) / 2] = __float2bfloat16(W[(size_t)(kb * 64 + k) * 4 * TC_U + col]);
    }
  };
  auto put_block = [&](__nv_bfloat16* blk, const std::vector<float>& W, int g, int kb) {
    put_half(blk, W, 2 * g, kb);
    put_half(blk + TC_HB_BYTES / 2, W, 2 * g + 1, kb);
  };
  for (int g = 0; g < TC_UG; ++g) {
    __nv_bfloat16* base = img.data() + (size_t)g * TC_IMG_BYTES / 2;
    for (int kb = 0; kb < TC_NKB_X; ++kb) put_block(base + (TC_IMG_W1X + kb * TC_B_BYTES) / 2, *src[0], g, kb);
    for (int kb = 0; kb < TC_NKB_H; ++kb) {
      put_block(base + (TC_IMG_WU + kb * 2 * TC_B_BYTES) / 2, *src[2], g, kb);               // W2 (cell_1/kernel)
      put_block(base + (TC_IMG_WU + kb * 2 * TC_B_BYTES + TC_B_BYTES) / 2, *src[1], g, kb);  // U1 (cell_0/recurrent_kernel)
      put_block(base + (TC_IMG_U2 + kb * TC_B_BYTES) / 2, *src[3], g, kb);                   // U2 (cell_1/recurrent_kernel)
    }
    for (int n = 0; n < 64; ++n) {
      const int half = n >> 5, gate = (n >> 3) & 3, u = n & 7;
      bias[(size_t)g * 128 + n] = b0[(size_t)gate * TC_U + (2 * g + half) * 8 + u];
      bias[(size_t)g * 128 + 64 + n] = b1[(size_t)gate * TC_U + (2 * g + half) * 8 + u];
    }
  }
  auto fail = [&](const char* m) { err = m; return GSTK_ECUDA; };
  if (cudaMalloc((void**)&st.wimg, img.size() * 2) != cudaSuccess) return fail("cudaMalloc(wimg) failed");
  if (cudaMalloc((void**)&st.bias, bias.size() * 4) != cudaSuccess) return fail("cudaMalloc(bias) failed");
  const size_t act_elems = (size_t)(TC_NKB_X + 2 * TC_NKB_H) * 2 * 128 * 64;
  if (cudaMalloc((void**)&st.act, act_elems * 2) != cudaSuccess) return fail("cudaMalloc(act) failed");
  if (cudaMemcpy(st.wimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  if (cudaMemcpy(st.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  st.fast_a = bf16_fast_a(c);
  if (st.fast_a) {
    // stage-ordered, fragment-ordered bf16 image of the four phase-A dense kernels (see FaW / fa_consume_layer; Keras layout is [K][N])
    const std::string dd = "Decoder/Decoder_Step/";
    const int PD = c.mel_dim * c.step_reduction + 1;
    const char* names[4] = {"Projection/kernel", "Prenet/dense/kernel", "Prenet/dense_1/kernel", "Attention/Query/kernel"};
    const int Ks[4] = {TC_U + 128, c.mel_dim, c.prenet0, c.prenet1};
    const int Ns[4] = {PD, c.prenet0, c.prenet1, c.attention_size};
    const FaW last = fa_wlayer(FA_L_QUERY, PD, c.mel_dim, c.prenet0, c.prenet1, c.attention_size, TC_U + 128);
    std::vector<__nv_bfloat16> F(((size_t)last.base + (size_t)last.nst() * last.stride()) / 2, __float2bfloat16(0.f));
    for (int l = 0; l < 4; ++l) {
      const std::vector<float>& W = hw.at(dd + names[l]);
      const FaW L = fa_wlayer(l, PD, c.mel_dim, c.prenet0, c.prenet1, c.attention_size, TC_U + 128);
      const int K = Ks[l], N = Ns[l];
      auto at = [&](int n, int k) { return (n < N && k < K) ? W[(size_t)k * N + n] : 0.f; };
      for (int si = 0; si < L.nst(); ++si) {
        const int kts = std::min(L.KTS, L.KT - si * L.KTS);
        for (int ft = 0; ft < L.NF; ++ft)
          for (int ki = 0; ki < kts; ++ki)
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane >> 2, t = lane & 3, n0 = ft * 16, k0 = (si * L.KTS + ki) * 16;
              __nv_bfloat16* o = F.data() + ((size_t)L.base + (size_t)si * L.stride()) / 2 + ((size_t)(ft * kts + ki) * 32 + lane) * 8;
              o[0] = __float2bfloat16(at(n0 + g, k0 + 2 * t));         o[1] = __float2bfloat16(at(n0 + g, k0 + 2 * t + 1));
              o[2] = __float2bfloat16(at(n0 + g + 8, k0 + 2 * t));     o[3] = __float2bfloat16(at(n0 + g + 8, k0 + 2 * t + 1));
              o[4] = __float2bfloat16(at(n0 + g, k0 + 2 * t + 8));     o[5] = __float2bfloat16(at(n0 + g, k0 + 2 * t + 9));
              o[6] = __float2bfloat16(at(n0 + g + 8, k0 + 2 * t + 8)); o[7] = __float2bfloat16(at(n0 + g + 8, k0 + 2 * t + 9));
            }
      }
    }
    if (cudaMalloc((void**)&st.wimgA, F.size() * 2) != cudaSuccess) return fail("cudaMalloc(wimgA) failed");
    if (cudaMemcpy(st.wimgA, F.data(), F.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return fail("memcpy failed");
  }
  st.ready = true;
  return GSTK_OK;
}

inline void bf16_release(Bf16State& st) {
  cudaFree(st.wimg);
  cudaFree(st.bias);
  cudaFree(st.act);
  cudaFree(st.wimgA);
  cudaFree(st.vproj_bf);
  cudaFree(st.qbuf);
  cudaFree(st.prof);
  st = Bf16State();
}

// One launch for a batch chunk of <= 256 rows.  p.h1/p.h2 (index 1 = "step -1"), p.c1/p.c2 must hold
// the initial states (fp32); they receive the final ones.
inline int bf16_decode(Bf16State& st, const GstkConfig& c, DecParams& p, int num_sms, cudaStream_t stream,
                       cudaEvent_t ev0, cudaEvent_t ev1, int64_t& launches, std::string& err) {
  auto fail = [&](int code, const std::string& m) { err = m; return code; };
  if (p.B > TC_MAX_B) return fail(GSTK_EINVAL, "bf16 decoder chunk larger than 256 rows");
  if (num_sms < TC_LSTM_CTAS) return fail(GSTK_ENODEVICE, "bf16 decoder needs at least 128 SMs");
  const int MT = (p.B + 127) / 128;
  const size_t smem = bf16_smem_bytes(p, st.fast_a);
  if (st.fast_a && (num_sms - TC_LSTM_CTAS) * DA_MAXU < p.B)
    return fail(GSTK_ENODEVICE, "bf16 decoder needs (SMs - 128) * 16 >= batch chunk for the phase-A dense CTAs");
  if (smem > 227 * 1024) return fail(GSTK_EINVAL, "key_time too large for the bf16 decoder's shared-memory budget");
  Bf16Params q;
  q.wimg = st.wimg;
  q.bias = st.bias;
  q.actX = st.act;
  q.actH1 = st.act + (size_t)TC_NKB_X * MT * 128 * 64;
  q.actH2 = q.actH1 + (size_t)TC_NKB_H * MT * 128 * 64;
  p.actX = q.actX;
  p.MT = MT;
  cudaError_t e;
  q.wimgA = nullptr; q.vproj_bf = nullptr; q.qbuf = nullptr;
  if (st.fast_a) {
    if (!st.qbuf && (e = cudaMalloc((void**)&st.qbuf, (size_t)TC_MAX_B * FA_A * sizeof(float))) != cudaSuccess)
      return fail(GSTK_ECUDA, cudaGetErrorString(e));
    q.qbuf = st.qbuf;
    const size_t nv = (size_t)p.B * p.Tv * 128;
    if (st.vproj_elems < nv) {
      cudaFree(st.vproj_bf);
      st.vproj_bf = nullptr;
      if ((e = cudaMalloc((void**)&st.vproj_bf, nv * 2)) != cudaSuccess) return fail(GSTK_ECUDA, cudaGetErrorString(e));
      st.vproj_elems = nv;
    }
    f32_to_bf16_kernel<<<num_sms * 2, 256, 0, stream>>>(p.vproj, st.vproj_bf, nv);
    launches += 1;
    q.wimgA = st.wimgA; q.vproj_bf = st.vproj_bf;
  }
  if (!st.prof) {
    if ((e = cudaMalloc((void**)&st.prof, (size_t)num_sms * PROF_SLOTS * sizeof(unsigned long long))) != cudaSuccess)
      return fail(GSTK_ECUDA, cudaGetErrorString(e));
    st.prof_ctas = num_sms;
  }
  q.prof = (p.debug_flags & 8) ? st.prof : nullptr;   // per-phase clock64 timers: opt-in (GSTK_DEBUG bit 3), thread 0 leads every barrier
  // operand images of the initial hidden states (rows >= B zero so that unused tile rows stay finite)
  if ((e = cudaMemsetAsync(q.actX, 0, (size_t)TC_NKB_X * MT * 128 * 64 * 2, stream)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h1 + (size_t)p.B * TC_U, q.actH1, p.B, MT);
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h2 + (size_t)p.B * TC_U, q.actH2, p.B, MT);
  launches += 2;
  if ((e = cudaFuncSetAttribute(decoder_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  int occ = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decoder_bf16_kernel, TC_THREADS, smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  if (occ < 1) return fail(GSTK_EINVAL, "bf16 decoder kernel does not fit on an SM");
  void* args[] = {&p, &q};
  cudaEventRecord(ev0, stream);
  if ((e = cudaLaunchCooperativeKernel((void*)decoder_bf16_kernel, dim3(num_sms), dim3(TC_THREADS), args, smem, stream)) !=
      cudaSuccess)
    return fail(GSTK_ECUDA, std::string("cooperative launch failed: ") + cudaGetErrorString(e));
  cudaEventRecord(ev1, stream);
  launches += 1;
  return GSTK_OK;
}

}  // namespace gstk
