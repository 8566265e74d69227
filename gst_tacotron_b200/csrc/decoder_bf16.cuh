// Persistent bf16 tensor-core decoder (tcgen05 / TMEM / bulk-async copies), the throughput mode.
//
// Same loop structure as decoder_fp32.cuh (one cooperative launch for the whole decode, three grid
// barriers per step, phase A = projection/prenet/attention per utterance in fp32), but the two
// LSTMCells (98 % of the FLOPs, Modules/Taco2.py:77-85,111) run on the 5th-gen tensor cores:
//
//   * CTA c < U/8 owns hidden units [8c, 8c+8) of BOTH cells = 32 gate columns per cell.  Its weight
//     slice is packed once (host) as 54 K-blocks of [32 rows x 64 k] bf16 in the canonical K-major
//     SWIZZLE_128B layout; RES_WB of them stay resident in shared memory for the whole decode, the
//     rest are streamed with the activations.
//   * activations (p || ctx, h1, h2) live in global memory as bf16 *pre-swizzled UMMA operand
//     images* [k-block][m-tile][128 rows][128 B]: the epilogue thread that owns (row, 8 units) stores
//     exactly one 16-byte swizzle chunk, and a consumer brings a tile in with ONE bulk async copy
//     (cp.async.bulk, SASS UBLKCP) that completes on an mbarrier - no tensor map needed.
//   * D[batch tile 128, 32 gate cols] (+)= A[128, 64] . B[32, 64]^T with tcgen05.mma kind::f16
//     (M=128, N=32, K=16 x4 per k-block), fp32 accumulators in TMEM, issued by one thread (warp 15),
//     which also runs the copy pipeline (NSTAGE-deep ring, full/empty mbarriers, tcgen05.commit).
//   * epilogue warps 0-7 read their row's 32 accumulator columns with tcgen05.ld, apply the LSTM
//     point-wise update with the cell state kept in REGISTERS for the whole decode, and publish h.
//   * the recurrent halves are taken off the critical path: h1(t).U1 is accumulated right after
//     h1(t) is published (same A tiles as h1(t).W2), h2(t-1).U2 during phase A of step t.
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
constexpr int TC_LSTM_CTAS = TC_U / 8;  // 128
constexpr int TC_NSTAGE = 3;      // 40 KB stages of the LSTM operand ring while phase A runs (h2.U2 segment)
constexpr int TC_NSTAGE_BC = 4;   // phases B and C borrow the (then idle) phase-A weight ring as a fourth stage: bytes in flight bound the stream
constexpr int TC_A_BYTES = 128 * 128;   // one activation tile (128 rows x 64 bf16)
constexpr int TC_B_BYTES = 32 * 128;    // one weight block (32 gate rows x 64 k)
// one pipeline unit = one k-block: the activation tiles of both m-tiles + up to two weight blocks (W2 | U1 of that k-block)
constexpr int TC_STAGE_W = 2 * TC_A_BYTES;
constexpr int TC_STAGE_BYTES = TC_STAGE_W + 2 * TC_B_BYTES;  // 40 KB
// per-CTA weight image: [W1x: 6 blocks][per k-block: W2 | U1 (adjacent => one N=64 B operand)][U2: 16 blocks]
constexpr int TC_IMG_W1X = 0, TC_IMG_WU = TC_NKB_X * TC_B_BYTES, TC_IMG_U2 = TC_IMG_WU + TC_NKB_H * 2 * TC_B_BYTES;
constexpr int TC_IMG_BYTES = TC_IMG_U2 + TC_NKB_H * TC_B_BYTES;  // 216 KB
constexpr int TC_THREADS = 384;     // 12 warps => up to 168 registers per thread
constexpr int TC_PA_THREADS = TC_THREADS - 64;  // warps 0-9 run phase A; warp 10 = copy producer, warp 11 = MMA issuer
constexpr int TC_PA_WARPS = TC_PA_THREADS / 32;
// TMEM columns: per m-tile [D2 (32) | D1 (32)] so that W2|U1 can be one N=64 MMA; cell states c1, c2 behind them
constexpr int TC_TMEM_COLS = 256;
constexpr uint32_t TC_DSTRIDE = 64, TC_D2 = 0, TC_D1 = 32, TC_C1 = 128, TC_C2 = 144;
constexpr int TC_MAX_B = 256;
// phase-A dense weights stream through their own ring of bulk-copied stages (<= FA_TPS mma.sync fragment tiles of 512 B).
// Stages are as large as shared memory allows: the copy warp's per-stage instruction path (~150-450 cycles for a lone
// warp), not bandwidth, bounds this stream, so fewer and bigger stages win (tools/ubench_handoff.cu).
constexpr int FA_WSTAGES = 2, FA_TPS = 56, FA_WSTAGE_BYTES = FA_TPS * 512;
__host__ __device__ constexpr int fa_kts(int NF) { return FA_TPS / NF > 0 ? FA_TPS / NF : 1; }  // k16 tiles per stage

static_assert(FA_WSTAGES * FA_WSTAGE_BYTES >= (TC_NSTAGE_BC - TC_NSTAGE) * TC_STAGE_BYTES, "borrowed LSTM stage must fit in the phase-A weight ring");

struct Bf16Params {
  const __nv_bfloat16* wimg;  // [TC_LSTM_CTAS][TC_IMG_BYTES] per-CTA swizzled weight blocks (TC_IMG_*)
  const float* bias;          // [TC_LSTM_CTAS][2][32]  (gate*8+u)
  __nv_bfloat16* actX;        // [6][MT][128][64]
  __nv_bfloat16* actH1;       // [16][MT][128][64]
  __nv_bfloat16* actH2;       // [16][MT][128][64]
  // phase A fast path (SMA): fragment-ordered bf16 weights and the bf16 copy of V'
  const uint8_t* wimgA;           // stage-ordered fragment images of Projection | Prenet0 | Prenet1 | Query (fa_wlayer)
  const __nv_bfloat16* vproj_bf;  // [B][Tv][128]
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

// Producer warp: walks the k-blocks (rotated by `rot`) of one segment and issues the bulk copies as stages free up: the
// activation tile of every m-tile plus the weight block(s) of that k-block (wbytes = 4 KB, or 8 KB for W2|U1).
// A lone warp issues dependent instructions every ~6-10 cycles, so the per-unit instruction count IS the pipeline's
// throughput limit (measured: tools/ubench_handoff.cu): units are as large as shared memory allows and the loop body is
// kept to pointer bumps.
template <int NKB, int MT, int NS>
__device__ __forceinline__ void tc_produce(TcRing& r, uint64_t* full, uint64_t* empty, uint8_t* stages, const uint8_t* act,
                                           const uint8_t* wsrc, uint32_t wstride, uint32_t wbytes, int B, int rot, unsigned long long* pw = nullptr) {
  const uint32_t a0 = (uint32_t)min(128, B) * 128u, a1 = MT == 2 ? (uint32_t)(B - 128) * 128u : 0u;
  const uint32_t total = a0 + a1 + wbytes;
  int kb = rot;
  const uint8_t* a = act + (size_t)kb * MT * TC_A_BYTES;
  const uint8_t* w = wsrc + (size_t)kb * wstride;
  r.stage = 0;  // every segment starts at stage 0 (producer and MMA warp agree; the per-stage parities carry over)
  for (int i = 0; i < NKB; ++i) {
    if (pw) { const long long w0 = clock64(); mbar_wait(&empty[r.stage], r.phase() ^ 1u); if ((threadIdx.x & 31) == 0) pw[0] += clock64() - w0; }
    else mbar_wait(&empty[r.stage], r.phase() ^ 1u);
    uint8_t* st = stages + (size_t)r.stage * TC_STAGE_BYTES;
    if (elect_one()) {
      mbar_arrive_expect_tx(&full[r.stage], total);
      bulk_g2s(st, a, a0, &full[r.stage]);
      if (MT == 2) bulk_g2s(st + TC_A_BYTES, a + TC_A_BYTES, a1, &full[r.stage]);
      bulk_g2s(st + TC_STAGE_W, w, wbytes, &full[r.stage]);
    }
    __syncwarp();
    r.template advance<NS>();
    if (++kb == NKB) { kb = 0; a = act; w = wsrc; }
    else { a += (size_t)MT * TC_A_BYTES; w += wstride; }
  }
}

__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t saddr) {  // low word of make_desc_sw128 (high word: umma_bf16_ss_lo)
  return ((saddr >> 4) & 0x3FFFu) | 0x10000u;
}

// MMA warp, one N=32 product per m-tile: D[mt] (+)= A[mt] . B^T with B = the 4 KB block at stage offset TC_STAGE_W.
template <int NKB, bool FRESH, int MT, int NS>
__device__ __forceinline__ void tc_consume(TcRing& r, uint64_t* full, uint64_t* empty, uint32_t stages_sa, uint32_t tmem_d,
                                           uint64_t* commit_done, unsigned long long* pw = nullptr) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 32);
  r.stage = 0;
  for (int i = 0; i < NKB; ++i) {
    const uint32_t acc = (FRESH && i == 0) ? 0u : 1u;
    if (pw) { const long long w0 = clock64(); mbar_wait(&full[r.stage], r.phase()); if ((threadIdx.x & 31) == 0) pw[1] += clock64() - w0; }
    else mbar_wait(&full[r.stage], r.phase());
    tc_fence_after();
    const uint32_t st_sa = stages_sa + r.stage * (uint32_t)TC_STAGE_BYTES;
    const uint32_t ad = tc_desc_lo(st_sa), bd = ad + (TC_STAGE_W >> 4);
    if (elect_one()) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss_lo(tmem_d + (uint32_t)mt * TC_DSTRIDE, ad + (uint32_t)(mt * (TC_A_BYTES >> 4) + 2 * k), bd + 2 * k, idesc, (k > 0) ? 1u : acc);
      if (commit_done && i == NKB - 1) umma_commit(commit_done);
      umma_commit(&empty[r.stage]);
    }
    __syncwarp();
    r.template advance<NS>();
  }
}

// MMA warp, LSTMCell-1 segment: the stage holds the h1 tiles + [W2 | U1] (64 gate rows).  D2 += h1.W2 (accumulates onto the
// pre-computed h2.U2) and D1 = h1.U1 (fresh, for the next step).  D2 and D1 are adjacent in TMEM, so from the second
// k-block on one N=64 MMA does both (N=64 costs 48 cycles vs 2 x 40 for two N=32 instructions).
template <int NKB, int MT, int NS>
__device__ __forceinline__ void tc_consume_wu(TcRing& r, uint64_t* full, uint64_t* empty, uint32_t stages_sa, uint32_t tmem,
                                              uint64_t* commit_done, unsigned long long* pw = nullptr) {
  constexpr uint32_t idesc32 = make_idesc_bf16(128, 32), idesc64 = make_idesc_bf16(128, 64);
  r.stage = 0;
  for (int i = 0; i < NKB; ++i) {
    if (pw) { const long long w0 = clock64(); mbar_wait(&full[r.stage], r.phase()); if ((threadIdx.x & 31) == 0) pw[1] += clock64() - w0; }
    else mbar_wait(&full[r.stage], r.phase());
    tc_fence_after();
    const uint32_t st_sa = stages_sa + r.stage * (uint32_t)TC_STAGE_BYTES;
    const uint32_t ad = tc_desc_lo(st_sa), bd = ad + (TC_STAGE_W >> 4), bdu = bd + (TC_B_BYTES >> 4);
    if (elect_one()) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t dc = tmem + (uint32_t)mt * TC_DSTRIDE;
        const uint32_t am = ad + (uint32_t)(mt * (TC_A_BYTES >> 4));
        if (i == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(dc + TC_D2, am + 2 * k, bd + 2 * k, idesc32, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(dc + TC_D1, am + 2 * k, bdu + 2 * k, idesc32, (k > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo(dc + TC_D2, am + 2 * k, bd + 2 * k, idesc64, 1u);
        }
      }
      if (i == NKB - 1) umma_commit(commit_done);
      umma_commit(&empty[r.stage]);
    }
    __syncwarp();
    r.template advance<NS>();
  }
}

// run-time m-tile count -> compile-time loop shape; out of line (own register allocation), ring state by value
template <int NKB, int NS>
__device__ __noinline__ TcRing seg_produce(int MT, TcRing r, uint64_t* full, uint8_t* stages, const uint8_t* act, const uint8_t* wsrc,
                                           uint32_t wstride, uint32_t wbytes, int B, int rot, unsigned long long* pw = nullptr) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  if (MT == 2) tc_produce<NKB, 2, NS>(r, full, empty, stages, act, wsrc, wstride, wbytes, B, rot, pw);
  else tc_produce<NKB, 1, NS>(r, full, empty, stages, act, wsrc, wstride, wbytes, B, rot, pw);
  return r;
}
template <int NKB, bool FRESH, int NS>
__device__ __noinline__ TcRing seg_consume(int MT, TcRing r, uint64_t* full, uint32_t stages_sa, uint32_t tmem_d, uint64_t* commit_done) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  if (MT == 2) tc_consume<NKB, FRESH, 2, NS>(r, full, empty, stages_sa, tmem_d, commit_done);
  else tc_consume<NKB, FRESH, 1, NS>(r, full, empty, stages_sa, tmem_d, commit_done);
  return r;
}
template <int NKB, int NS>
__device__ __noinline__ TcRing seg_consume_wu(int MT, TcRing r, uint64_t* full, uint32_t stages_sa, uint32_t tmem, uint64_t* commit_done,
                                              unsigned long long* pw = nullptr) {
  uint64_t* empty = full + TC_NSTAGE_BC;
  if (MT == 2) tc_consume_wu<NKB, 2, NS>(r, full, empty, stages_sa, tmem, commit_done, pw);
  else tc_consume_wu<NKB, 1, NS>(r, full, empty, stages_sa, tmem, commit_done, pw);
  return r;
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
  float4* hf = reinterpret_cast<float4*>(h_out + cta * 8);
  hf[0] = make_float4(h[0], h[1], h[2], h[3]);
  hf[1] = make_float4(h[4], h[5], h[6], h[7]);
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

// Phase-A dense weights are streamed by the producer warp with bulk-async copies into a 2-stage shared-memory
// ring (FA_WSTAGES x 16 KB).  A stage holds, for ALL feature tiles of a layer, KTS consecutive k16 tiles:
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

template <int NF, int KT>
__device__ __forceinline__ void fa_produce_layer(uint32_t& cnt, uint64_t* wfull, uint64_t* wempty, uint8_t* wstages, const uint8_t* src) {
  constexpr int KTS = fa_kts(NF), NST = (KT + KTS - 1) / KTS, LASTK = KT - (NST - 1) * KTS;
  constexpr uint32_t FULLB = (uint32_t)NF * KTS * 512u, LASTB = (uint32_t)NF * LASTK * 512u;
#pragma unroll 1
  for (int si = 0; si < NST; ++si, ++cnt, src += FULLB) {
    const uint32_t st = cnt % FA_WSTAGES, ph = (cnt / FA_WSTAGES) & 1u;
    const uint32_t bytes = (si == NST - 1) ? LASTB : FULLB;
    mbar_wait(&wempty[st], ph ^ 1u);
    if (elect_one()) {
      mbar_arrive_expect_tx(&wfull[st], bytes);
      bulk_g2s(wstages + (size_t)st * FA_WSTAGE_BYTES, src, bytes, &wfull[st]);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// y[n][u] = sum_k W[k][n] act[u][k] for the NU utterances of this CTA; out: fp32 [NF*16][2].
// Layer shape is a compile-time constant (the fast path is only taken for the reference's default widths), so
// the per-stage tile loop is fully unrolled: ~6 instructions per mma instead of ~40 with run-time bounds.
template <int NU, int NF, int KT, int KTS, int KN /* k-tiles in this stage */>
__device__ __forceinline__ void fa_stage_mma(float (&d)[2][2][4], const uint4* __restrict__ tiles, const __nv_bfloat16* arow, int kt0, int wid,
                                             int lane, bool has_b) {
#pragma unroll
  for (int sl = 0; sl < 2; ++sl) {
    if (sl * FA_WARPS >= NF) break;
    const int ft = wid + sl * FA_WARPS;
    if (ft < NF) {
      const uint4* tp = tiles + (size_t)(ft * KN) * 32 + lane;
#pragma unroll
      for (int ki = 0; ki < KN; ++ki) {
        const uint4 a = tp[ki * 32];
        uint32_t b0 = 0, b1 = 0;
        if (has_b) {
          b0 = *reinterpret_cast<const uint32_t*>(arow + (kt0 + ki) * 16);
          b1 = *reinterpret_cast<const uint32_t*>(arow + (kt0 + ki) * 16 + 8);
        }
        mma_16816_bf16(d[sl][ki & 1], a, b0, b1);  // two independent accumulator chains per feature tile
      }
    }
  }
}

template <int NU, int NF, int KT>
__device__ __forceinline__ void fa_consume_layer(uint32_t& cnt, uint64_t* wfull, uint64_t* wempty, const uint8_t* wstages,
                                                 const __nv_bfloat16* act_s, int kstride, float* out, int wid, int lane) {
  constexpr int KTS = fa_kts(NF);
  constexpr int NST = (KT + KTS - 1) / KTS, LASTK = KT - (NST - 1) * KTS;
  const int g = lane >> 2, t = lane & 3;
  const bool has_b = g < NU;
  const __nv_bfloat16* arow = act_s + (has_b ? g : 0) * kstride + 2 * t;
  float d[2][2][4] = {};
#pragma unroll 1
  for (int si = 0; si < NST; ++si, ++cnt) {
    const uint32_t st = cnt % FA_WSTAGES, ph = (cnt / FA_WSTAGES) & 1u;
    mbar_wait(&wfull[st], ph);
    const uint4* tiles = reinterpret_cast<const uint4*>(wstages + (size_t)st * FA_WSTAGE_BYTES);
    if (si < NST - 1 || LASTK == KTS) fa_stage_mma<NU, NF, KT, KTS, KTS>(d, tiles, arow, si * KTS, wid, lane, has_b);
    else fa_stage_mma<NU, NF, KT, KTS, LASTK>(d, tiles, arow, si * KTS, wid, lane, has_b);
    __syncwarp();
    if (lane == 0) mbar_arrive(&wempty[st]);  // this warp is done with the stage
  }
  if (t == 0) {  // columns 0,1 of D = utterances 0,1
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int ft = wid + sl * FA_WARPS;
      if (ft < NF) {
        float* o = out + (size_t)(ft * 16 + g) * 2;
        o[0] = d[sl][0][0] + d[sl][1][0]; o[1] = d[sl][0][1] + d[sl][1][1];
        o[16] = d[sl][0][2] + d[sl][1][2]; o[17] = d[sl][0][3] + d[sl][1][3];  // feature g + 8
      }
    }
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

constexpr int FA_HC = TC_U + 128;   // [h2 || ctx]
constexpr int FA_PARTF = 256 * 2;    // dense-layer output sums [features <= 256][2 utterances]

struct FaSmem {
  __nv_bfloat16* act;  // [2][FA_HC] bf16 layer input (hc / x / p0 / p1)
  float* part;         // [FA_PARTF] output sums of the current dense layer
  float* y;            // [2][PDp] projection output (the decoder input is taken from it in free mode)
  float* qv;           // [2][128] projected query
  float* alig;  // [2 utterances][2 (step parity)][Tv] alignments, resident for the whole decode
  float* ctxp;  // [FA_WARPS][128]
  const float* bias;   // [FA_BIAS_N] bp | b0 | b1 | bq | attention_v, staged once per launch
  unsigned long long* prof;  // shared-memory phase timers (null when profiling is off)
};
constexpr int FA_B_P = 0, FA_B_0 = 96, FA_B_1 = 96 + 256, FA_B_Q = 96 + 512, FA_B_V = 96 + 512 + 128, FA_BIAS_N = 96 + 512 + 256;

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

// shared-memory carve-up of the fast path (kept inside the callee: the kernel body only keeps one pointer live)
__device__ __forceinline__ FaSmem fa_carve(float* scratch, const DecParams& p, unsigned long long* prof, const uint8_t* wstages = nullptr) {
  FaSmem fs;
  float* f = scratch;
  auto take = [&](int n) { float* r = f; f += (n + 3) & ~3; return r; };
  fs.act = reinterpret_cast<__nv_bfloat16*>(take(FA_HC));  // 2 x FA_HC bf16
  fs.part = take(FA_PARTF);
  fs.y = take(2 * ((p.PD + 3) & ~3)); fs.qv = take(256);
  fs.alig = take(4 * p.Tv);
  // context partials live in weight-ring stage 0: every stage of this step has been consumed before the attention starts
  fs.ctxp = reinterpret_cast<float*>(const_cast<uint8_t*>(wstages));
  fs.bias = take(FA_BIAS_N);
  fs.prof = prof;
  return fs;
}

template <int NU>
__device__ __noinline__ void phase_a_fast(const DecParams& p, const Bf16Params& q, float* scratch, unsigned long long* prof,
                                          uint64_t* wfull, const uint8_t* wstages, int b0, int t) {
  const FaSmem s = fa_carve(scratch, p, prof, wstages);
  uint64_t* wempty = wfull + FA_WSTAGES;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int cur = t & 1, prv = cur ^ 1;
  const int XW = p.P1 + p.A;
  const int PDp = (p.PD + 3) & ~3;
  const int NFP = (p.PD + 15) >> 4;   // feature tiles of the projection
  int bs[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) bs[u] = b0 + u * (int)gridDim.x;
  const int melp = (p.mel + 15) & ~15;
  uint32_t wcnt = fa_stages_before(t, FA_NST_P, FA_NST_REST);  // position in the weight ring
  const unsigned int step_id = p.step_offset + (unsigned int)t;
  const bool drop = p.rng_mode != 0 && p.drop_rate > 0.f;
  if (t > 0) {
    // ---- projection of step t-1 (Taco2.py:112-118): input [h2 || ctx] rounded to bf16
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const float* h2 = p.h2 + ((size_t)prv * p.B + bs[u]) * TC_U;
      for (int i = tid; i < TC_U / 4; i += TC_PA_THREADS) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(h2) + i);
        __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(s.act + u * FA_HC + 4 * i);
        dst[0] = __floats2bfloat162_rn(v.x, v.y);
        dst[1] = __floats2bfloat162_rn(v.z, v.w);
      }
      for (int i = tid; i < p.A; i += TC_PA_THREADS)
        s.act[u * FA_HC + TC_U + i] = __float2bfloat16(__ldcg(p.xin + (size_t)bs[u] * XW + p.P1 + i));
    }
    pa_sync<TC_PA_THREADS>();
    fa_consume_layer<NU, 6, FA_HC / 16>(wcnt, wfull, wempty, wstages, s.act, FA_HC, s.part, wid, lane);
    pa_sync<TC_PA_THREADS>();
    // output pass; in free-running mode the last of the r frames is also the next decoder input (Taco2.py:183-187)
    for (int i = tid; i < NU * p.PD; i += TC_PA_THREADS) {
      const int u = i / p.PD, n = i - u * p.PD;
      const float v = s.bias[FA_B_P + n] + s.part[n * 2 + u];
      if (n < p.PD - 1) {
        if (p.out_mel) p.out_mel[((size_t)bs[u] * p.T + (t - 1)) * (p.PD - 1) + n] = v;
        const int f = n - (p.r - 1) * p.mel;
        if (p.mode == 0 && f >= 0) s.act[u * FA_HC + f] = __float2bfloat16(v);
      } else if (p.out_stop) {
        p.out_stop[(size_t)bs[u] * p.T + (t - 1)] = v;
      }
    }
  }
  prof_tick(s.prof, 6);
  if (t == p.T) return;
  // ---- decoder input (Taco2.py:183-187), rounded to bf16 for the tensor cores; zero padding up to melp
  for (int i = tid; i < NU * melp; i += TC_PA_THREADS) {
    const int u = i / melp, n = i - u * melp;
    if (n >= p.mel) s.act[u * FA_HC + n] = __float2bfloat16(0.f);
    else if (p.mode == 1) s.act[u * FA_HC + n] = __float2bfloat16(__ldg(p.teacher + (size_t)bs[u] * p.ts_b + (size_t)t * p.ts_t + n));
    else if (t == 0) s.act[u * FA_HC + n] = __float2bfloat16(p.init_mel ? __ldg(p.init_mel + (size_t)bs[u] * p.mel + n) : 0.f);
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 7);
  // ---- prenet layer 0 (Taco2.py:270-283, dropout always on)
  fa_consume_layer<NU, 16, 5>(wcnt, wfull, wempty, wstages, s.act, FA_HC, s.part, wid, lane);
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 8);
  for (int i = tid; i < NU * (p.P0 / 4); i += TC_PA_THREADS) {
    const int u = i / (p.P0 / 4), n4 = (i - u * (p.P0 / 4)) * 4;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = s.bias[FA_B_0 + n4 + k] + s.part[(n4 + k) * 2 + u];
    fa_relu_dropout4(v, p, p.keep0 + ((size_t)t * p.rngB + p.rng_b0 + bs[u]) * p.P0, STREAM_KEEP0, step_id, p.row_offset + bs[u], n4, drop);
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(s.act + u * FA_HC + n4);
    dst[0] = __floats2bfloat162_rn(v[0], v[1]);
    dst[1] = __floats2bfloat162_rn(v[2], v[3]);
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 9);
  // ---- prenet layer 1
  fa_consume_layer<NU, 16, 16>(wcnt, wfull, wempty, wstages, s.act, FA_HC, s.part, wid, lane);
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 10);
  for (int i = tid; i < NU * (p.P1 / 4); i += TC_PA_THREADS) {
    const int u = i / (p.P1 / 4), n4 = (i - u * (p.P1 / 4)) * 4;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = s.bias[FA_B_1 + n4 + k] + s.part[(n4 + k) * 2 + u];
    fa_relu_dropout4(v, p, p.keep1 + ((size_t)t * p.rngB + p.rng_b0 + bs[u]) * p.P1, STREAM_KEEP1, step_id, p.row_offset + bs[u], n4, drop);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(s.act + u * FA_HC + n4);
    dst[0] = lo;
    dst[1] = hi;
    __nv_bfloat162* gdst = reinterpret_cast<__nv_bfloat162*>(p.actX + act_elem_index(p.MT, bs[u], n4));  // 4 units stay in one 16 B chunk
    gdst[0] = lo;
    gdst[1] = hi;
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 11);
  // ---- query projection (Steps.py:122)
  fa_consume_layer<NU, 8, 16>(wcnt, wfull, wempty, wstages, s.act, FA_HC, s.part, wid, lane);
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 12);
  for (int i = tid; i < NU * p.A; i += TC_PA_THREADS) {
    const int u = i / p.A, n = i - u * p.A;
    s.qv[u * 128 + n] = s.bias[FA_B_Q + n] + s.part[n * 2 + u];
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 13);
  // ---- fused stepwise-monotonic attention, one pass over V' (Steps.py:138-166, 215-229).
  // Warp w owns rows [j0, j1); a batch is 16 rows starting one row early (row j needs p[j-1]).  The 128-wide
  // energy dot products are reduce-SCATTERED (16 shuffles for 16 rows) so that lane l ends up owning row l>>1:
  // noise, sigmoid and the alignment recurrence run lane-parallel instead of warp-redundantly.
  const float sb = __ldg(p.att_sb);
  const bool noisy = p.rng_mode != 0 && p.sigmoid_noise > 0.f;
  const int jw = (p.Tv + FA_WARPS - 1) / FA_WARPS;
  if (t == 0) {  // first step of this launch: fetch the initial alignments into their resident buffers
    for (int i = tid; i < NU * p.Tv; i += TC_PA_THREADS) {
      const int u = i / p.Tv, j = i - u * p.Tv;
      s.alig[(u * 2 + prv) * p.Tv + j] = __ldcg(p.align + ((size_t)prv * p.B + bs[u]) * p.Tv + j);
    }
    pa_sync<TC_PA_THREADS>();
  }
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int b = bs[u];
    const float* prev_s = s.alig + (u * 2 + prv) * p.Tv;   // alignment of step t-1, resident in shared memory
    float* cur_s = s.alig + (u * 2 + cur) * p.Tv;
    const __nv_bfloat16* V = q.vproj_bf + (size_t)b * p.Tv * 128;
    const float4 q4 = *reinterpret_cast<const float4*>(s.qv + u * 128 + 4 * lane);
    const float4 v4 = *reinterpret_cast<const float4*>(s.bias + FA_B_V + 4 * lane);
    const int j0 = wid * jw, j1 = min(p.Tv, j0 + jw);
    float4 ctx = make_float4(0.f, 0.f, 0.f, 0.f);
    float p_carry = 0.f;  // sigmoid of the row before the batch
    const int r = lane >> 1;  // row of the batch this lane owns after the reduce-scatter
    for (int jj = j0 - 1; jj < j1; jj += 15) {  // 16 rows, the first one only provides p[j-1]
      uint2 kraw[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = jj + i;
        kraw[i] = (j >= 0 && j < j1) ? __ldg(reinterpret_cast<const uint2*>(V + (size_t)j * 128) + lane) : make_uint2(0, 0);
      }
      // noise of this lane's row (counter-based: any lane can compute any row)
      const int jr = jj + r;
      float nz = 0.f;
      if (noisy && jr >= 0 && jr < j1) {
        if (p.rng_mode == 1) nz = __ldg(p.noise + ((size_t)t * p.rngB + p.rng_b0 + b) * p.Tv + jr);
        else {
          // one normal per lane: only the half of the Philox block this row needs, fast-math Box-Muller
          const uint4 rr = philox4x32_10(make_uint4((unsigned int)jr >> 2, step_id, p.row_offset + b, (unsigned int)STREAM_NOISE),
                                         make_uint2((unsigned int)p.seed, (unsigned int)(p.seed >> 32)));
          const int w = jr & 3;
          const unsigned int ua = (w & 2) ? rr.z : rr.x, ub = (w & 2) ? rr.w : rr.y;
          const float sc = 5.9604644775390625e-08f;
          const float rad = sqrtf(-2.0f * __logf(((float)(ua >> 8) + 1.0f) * sc));
          float sn, cs;
          __sincosf(6.283185307179586f * ((float)(ub >> 8) * sc), &sn, &cs);
          nz = rad * ((w & 1) ? sn : cs);
        }
      }
      float e[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&kraw[i]);
        const float2 ka = __bfloat1622float2(kh[0]), kb = __bfloat1622float2(kh[1]);
        float a = v4.x * tanh_fast(q4.x + ka.x);
        a = fmaf(v4.y, tanh_fast(q4.y + ka.y), a);
        a = fmaf(v4.z, tanh_fast(q4.z + kb.x), a);
        a = fmaf(v4.w, tanh_fast(q4.w + kb.y), a);
        e[i] = a;
      }
      // reduce-scatter: after offsets 16,8,4,2 lane l holds the partial sum of row (l>>1)&15, then offset 1 completes it
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool hi = lane & 16;
        const float send = hi ? e[i] : e[i + 8];
        const float keep = hi ? e[i + 8] : e[i];
        e[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool hi = lane & 8;
        const float send = hi ? e[i] : e[i + 4];
        const float keep = hi ? e[i + 4] : e[i];
        e[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool hi = lane & 4;
        const float send = hi ? e[i] : e[i + 2];
        const float keep = hi ? e[i + 2] : e[i];
        e[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      {
        const bool hi = lane & 2;
        const float send = hi ? e[0] : e[1];
        const float keep = hi ? e[1] : e[0];
        e[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      float er = e[0] + __shfl_xor_sync(0xffffffffu, e[0], 1) + sb;  // energy of row jr (both lanes of the pair)
      if (noisy) er = fmaf(p.sigmoid_noise, nz, er);
      const float pr = sigmoid_fast(er);
      float p_before = __shfl_up_sync(0xffffffffu, pr, 2);  // p of row jr - 1
      if (r == 0) p_before = p_carry;
      float a = 0.f;
      if (jr >= j0 && jr < j1 && r > 0) {
        a = prev_s[jr] * pr;
        if (jr > 0) a = fmaf(prev_s[jr - 1], 1.0f - p_before, a);
        if ((lane & 1) == 0) cur_s[jr] = a;
      }
      p_carry = __shfl_sync(0xffffffffu, pr, 30);  // row 15 of this batch = row "-1" of the next one
#pragma unroll
      for (int i = 1; i < 16; ++i) {
        const float ai = __shfl_sync(0xffffffffu, a, 2 * i);
        const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&kraw[i]);
        const float2 ka = __bfloat1622float2(kh[0]), kb = __bfloat1622float2(kh[1]);
        ctx.x = fmaf(ai, ka.x, ctx.x); ctx.y = fmaf(ai, ka.y, ctx.y);
        ctx.z = fmaf(ai, kb.x, ctx.z); ctx.w = fmaf(ai, kb.y, ctx.w);
      }
    }
    reinterpret_cast<float4*>(s.ctxp + (u * FA_WARPS + wid) * 128)[lane] = ctx;
  }
  pa_sync<TC_PA_THREADS>();
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int b = bs[u];
    const float* cur_s = s.alig + (u * 2 + cur) * p.Tv;
    if (p.out_align)
      for (int j = tid; j < p.Tv; j += TC_PA_THREADS) p.out_align[((size_t)b * p.T + t) * p.Tv + j] = cur_s[j];
    if (t == p.T - 1) {  // publish the final alignment for state hand-over
      float* al_g = p.align + ((size_t)cur * p.B + b) * p.Tv;
      for (int j = tid; j < p.Tv; j += TC_PA_THREADS) al_g[j] = cur_s[j];
    }
  }
  if (tid < 128 * NU) {
    const int u = tid >> 7, n = tid & 127, b = bs[u];
    float c = 0.f;
#pragma unroll
    for (int w = 0; w < FA_WARPS; ++w) c += s.ctxp[(u * FA_WARPS + w) * 128 + n];
    p.xin[(size_t)b * XW + p.P1 + n] = c;
    p.actX[act_elem_index(p.MT, b, p.P1 + n)] = __float2bfloat16(c);
    if (p.out_ctx && t == p.T - 1) p.out_ctx[(size_t)b * p.A + n] = c;
  }
  pa_sync<TC_PA_THREADS>();
  prof_tick(s.prof, 14);
}

__global__ void __launch_bounds__(TC_THREADS, 1) decoder_bf16_kernel(const __grid_constant__ DecParams p, const __grid_constant__ Bf16Params q) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ int ok_s;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * TC_NSTAGE_BC + 3 + 2 * FA_WSTAGES];
  __shared__ float bias_s[64];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform (role dispatch stays on the uniform datapath)
  const int cta = blockIdx.x;
  const bool lstm_cta = cta < TC_LSTM_CTAS;
  uint8_t* stages = sm;                                             // TC_NSTAGE x 40 KB
  uint8_t* wstages = stages + (size_t)TC_NSTAGE * TC_STAGE_BYTES;   // FA_WSTAGES x 16 KB phase-A weight ring
  float* scratch = reinterpret_cast<float*>(wstages + (size_t)FA_WSTAGES * FA_WSTAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_NSTAGE_BC;
  uint64_t* d1_full = bars + 2 * TC_NSTAGE_BC;
  uint64_t* d2_full = bars + 2 * TC_NSTAGE_BC + 1;
  uint64_t* wres_full = bars + 2 * TC_NSTAGE_BC + 2;
  uint64_t* wfull = bars + 2 * TC_NSTAGE_BC + 3;   // [FA_WSTAGES] full, then [FA_WSTAGES] empty

  // generic (BMA / LSA) and fast (SMA) phase-A scratch share the same region; both are carved inside the callees
  const bool fast_a = q.wimgA != nullptr;
  if (fast_a) {
    float* bias_c = const_cast<float*>(fa_carve(scratch, p, nullptr).bias);
    for (int i = tid; i < FA_BIAS_N; i += TC_THREADS) {
      float v = 0.f;
      if (i < FA_B_0) { if (i < p.PD) v = __ldg(p.bp + i); }
      else if (i < FA_B_1) { if (i - FA_B_0 < p.P0) v = __ldg(p.b0 + i - FA_B_0); }
      else if (i < FA_B_Q) v = __ldg(p.b1 + i - FA_B_1);
      else if (i < FA_B_V) v = __ldg(p.bq + i - FA_B_Q);
      else v = __ldg(p.att_v + i - FA_B_V);
      bias_c[i] = v;
    }
  }
  __shared__ unsigned long long prof_sh[PROF_SLOTS + 1];
  unsigned long long* prof_s = q.prof ? prof_sh : nullptr;
  if (tid == 0) {
    for (int i = 0; i < PROF_SLOTS; ++i) prof_sh[i] = 0;
    prof_sh[PROF_SLOTS] = (unsigned long long)clock64();
  }
  auto prof_mark = [&](int slot) { prof_tick(prof_s, slot); };
  // GSTK_DEBUG bit 1: slots 6..9 = phase-C producer wait-on-empty, MMA wait-on-full, producer total, MMA total (phase-A sub-timers off)
  unsigned long long* dbg_pw = (q.prof && (p.debug_flags & 2)) ? prof_sh + 6 : nullptr;
  if (tid == 0) {
    for (int i = 0; i < TC_NSTAGE_BC; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(d1_full, 1);
    mbar_init(d2_full, 1);
    mbar_init(wres_full, 1);
    for (int i = 0; i < FA_WSTAGES; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wfull[FA_WSTAGES + i], FA_WARPS);  // every phase-A warp releases a stage
    }
    mbar_fence_init();
  }
  if (lstm_cta && tid < 64) bias_s[tid] = __ldg(q.bias + (size_t)cta * 64 + tid);
  if (lstm_cta && wid == 0) tmem_alloc(&tmem_base_s, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int MT = p.MT;
  const bool copy_warp = wid == TC_PA_WARPS;                  // bulk-copy producer warp (every CTA: phase-A weights)
  const bool prod_warp = lstm_cta && copy_warp;               // ... and the LSTM operand tiles on LSTM CTAs
  const bool mma_warp = lstm_cta && wid == TC_PA_WARPS + 1;   // tcgen05.mma issuer
  TcRing ring;
  ring.stage = 0; ring.bits = 0;
  const uint8_t* wimg_cta = reinterpret_cast<const uint8_t*>(q.wimg) + (size_t)cta * TC_IMG_BYTES;
  const uint32_t tmem = lstm_cta ? tmem_base_s : 0u;
  const uint32_t stages_sa = smem_u32(stages);
  const int rot_x = cta % TC_NKB_X, rot_h = cta % TC_NKB_H;  // per-CTA k-block rotation (spreads the L2 hot spot)
  const uint8_t* actX_b = reinterpret_cast<const uint8_t*>(q.actX);
  const uint8_t* actH1_b = reinterpret_cast<const uint8_t*>(q.actH1);
  const uint8_t* actH2_b = reinterpret_cast<const uint8_t*>(q.actH2);

  // cell state of (row, this CTA's 8 units) lives in registers of epilogue warps 0..4*MT-1
  const bool epi = lstm_cta && wid < 4 * MT;
  const int erow = (wid >> 2) * 128 + (wid & 3) * 32 + lane;  // batch row of this epilogue thread
  const bool erow_ok = epi && erow < p.B;
  // TMEM address of this epilogue thread's row: lane quarter of the warp, m-tile selects the column block
  const uint32_t t_row = tmem + ((uint32_t)((wid & 3) * 32) << 16);
  const uint32_t t_mt = (uint32_t)(wid >> 2);
  if (epi) {  // initial cell states -> TMEM (they stay there for the whole decode)
    float c1[8], c2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      c1[u] = erow_ok ? __ldcg(p.c1 + (size_t)erow * TC_U + cta * 8 + u) : 0.f;
      c2[u] = erow_ok ? __ldcg(p.c2 + (size_t)erow * TC_U + cta * 8 + u) : 0.f;
    }
    tmem_st8(t_row + TC_C1 + t_mt * 8u, c1);
    tmem_st8(t_row + TC_C2 + t_mt * 8u, c2);
  }

  if (prod_warp) {
    // prologue: D1 = h1(-1) . U1 (the images of the initial states were packed by the host-side kernel); U1 = second half of W2|U1
    ring = seg_produce<TC_NKB_H, TC_NSTAGE>(MT, ring, full, stages, actH1_b, wimg_cta + TC_IMG_WU + TC_B_BYTES, 2 * TC_B_BYTES, TC_B_BYTES, p.B, rot_h);
  } else if (mma_warp) {
    ring = seg_consume<TC_NKB_H, true, TC_NSTAGE>(MT, ring, full, stages_sa, tmem + TC_D1, nullptr);
  }

  unsigned int gen = 0;
  for (int t = 0; t <= p.T; ++t) {
    // ---------------- phase A (+ overlapped: D2 = h2(t-1) . U2) --------------------------------
    if (wid < TC_PA_WARPS) {
      if (fast_a) {
        // owned utterances: cta, cta + grid (chunks are <= 256 rows, so at most two)
        unsigned long long* prof_a = dbg_pw ? nullptr : prof_s;
        if (cta + (int)gridDim.x < p.B) phase_a_fast<2>(p, q, scratch, prof_a, wfull, wstages, cta, t);
        else if (cta < p.B) phase_a_fast<1>(p, q, scratch, prof_a, wfull, wstages, cta, t);
      } else {
        for (int b = cta; b < p.B; b += gridDim.x) phase_a_generic(p, scratch, b, t);
      }
      fence_proxy_async();  // actX stores (generic proxy) -> later bulk copies (async proxy)
    } else if (copy_warp) {
      if (fast_a && cta < p.B) {
        // stream this step's dense-layer weights (projection of step t-1, then prenet x2, query) to the phase-A warps
        uint32_t wcnt = fa_stages_before(t, FA_NST_P, FA_NST_REST);
        if (t > 0) fa_produce_layer<FA_LP.NF, FA_LP.KT>(wcnt, wfull, wfull + FA_WSTAGES, wstages, q.wimgA + FA_LP.base);
        if (t < p.T) {
          fa_produce_layer<FA_L0.NF, FA_L0.KT>(wcnt, wfull, wfull + FA_WSTAGES, wstages, q.wimgA + FA_L0.base);
          fa_produce_layer<FA_L1.NF, FA_L1.KT>(wcnt, wfull, wfull + FA_WSTAGES, wstages, q.wimgA + FA_L1.base);
          fa_produce_layer<FA_LQ.NF, FA_LQ.KT>(wcnt, wfull, wfull + FA_WSTAGES, wstages, q.wimgA + FA_LQ.base);
        }
      }
      if (prod_warp && t < p.T && !(p.debug_flags & 1)) {
        fence_proxy_async();
        ring = seg_produce<TC_NKB_H, TC_NSTAGE>(MT, ring, full, stages, actH2_b, wimg_cta + TC_IMG_U2, TC_B_BYTES, TC_B_BYTES, p.B, rot_h);
      }
    } else if (mma_warp && t < p.T && !(p.debug_flags & 1)) {
      tc_fence_after();
      ring = seg_consume<TC_NKB_H, true, TC_NSTAGE>(MT, ring, full, stages_sa, tmem + TC_D2, nullptr);
    }
    if (t == p.T) break;
    prof_mark(0);
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    prof_mark(1);
    // ---------------- phase B: LSTMCell 0 -------------------------------------------------------
    if (prod_warp) {
      fence_proxy_async();
      ring = seg_produce<TC_NKB_X, TC_NSTAGE_BC>(MT, ring, full, stages, actX_b, wimg_cta + TC_IMG_W1X, TC_B_BYTES, TC_B_BYTES, p.B, rot_x);
    } else if (mma_warp) {
      ring = seg_consume<TC_NKB_X, false, TC_NSTAGE_BC>(MT, ring, full, stages_sa, tmem + TC_D1, d1_full);
    }
    if (epi) {
      mbar_wait_backoff(d1_full, (uint32_t)t & 1u);
      tc_fence_after();
      float v[32], c[8];
      tmem_ld32(t_row + TC_D1 + t_mt * TC_DSTRIDE, v);
      tmem_ld8(t_row + TC_C1 + t_mt * 8u, c);
      if (erow_ok)
        tc_epilogue_row(v, bias_s, c, erow, cta, MT, q.actH1, p.h1 + ((size_t)(t & 1) * p.B + erow) * TC_U);
      tmem_st8(t_row + TC_C1 + t_mt * 8u, c);
      tc_fence_before();
      fence_proxy_async();
    }
    prof_mark(2);
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    prof_mark(3);
    // ---------------- phase C: LSTMCell 1 (+ D1 = h1(t) . U1 for the next step) ------------------
    if (prod_warp) {
      fence_proxy_async();
      const long long s0 = clock64();
      ring = seg_produce<TC_NKB_H, TC_NSTAGE_BC>(MT, ring, full, stages, actH1_b, wimg_cta + TC_IMG_WU, 2 * TC_B_BYTES, 2 * TC_B_BYTES, p.B, rot_h, dbg_pw);
      if (dbg_pw && lane == 0) dbg_pw[2] += clock64() - s0;
    } else if (mma_warp) {
      tc_fence_after();
      const long long s0 = clock64();
      ring = seg_consume_wu<TC_NKB_H, TC_NSTAGE_BC>(MT, ring, full, stages_sa, tmem, d2_full, dbg_pw);
      if (dbg_pw && lane == 0) dbg_pw[3] += clock64() - s0;
    }
    if (epi) {
      mbar_wait_backoff(d2_full, (uint32_t)t & 1u);
      tc_fence_after();
      float v[32], c[8];
      tmem_ld32(t_row + TC_D2 + t_mt * TC_DSTRIDE, v);
      tmem_ld8(t_row + TC_C2 + t_mt * 8u, c);
      if (erow_ok)
        tc_epilogue_row(v, bias_s + 32, c, erow, cta, MT, q.actH2, p.h2 + ((size_t)(t & 1) * p.B + erow) * TC_U);
      tmem_st8(t_row + TC_C2 + t_mt * 8u, c);
      tc_fence_before();
      fence_proxy_async();
    }
    prof_mark(4);
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    prof_mark(5);
    if (mma_warp) tc_fence_after();
  }
  if (q.prof && tid == 0)
    for (int i = 0; i < PROF_SLOTS; ++i) q.prof[(size_t)cta * PROF_SLOTS + i] = prof_sh[i];
  // final cell states (h is already in p.h1 / p.h2)
  if (epi) {
    float c1[8], c2[8];
    tmem_ld8(t_row + TC_C1 + t_mt * 8u, c1);
    tmem_ld8(t_row + TC_C2 + t_mt * 8u, c2);
    if (erow_ok) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        p.c1[(size_t)erow * TC_U + cta * 8 + u] = c1[u];
        p.c2[(size_t)erow * TC_U + cta * 8 + u] = c2[u];
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

inline size_t bf16_smem_bytes(const DecParams& p) {
  auto r4 = [](int n) { return (size_t)((n + 3) & ~3); };
  const size_t generic = 4 * (r4(p.mel) + r4(p.PD) + r4(p.U1 + p.A) + r4(p.P0) + r4(p.P1) + r4(p.A) + 4 * r4(p.Tv) +
                              DEC_THREADS + 8);
  const size_t fast = 4 * (r4(FA_HC) + r4(FA_PARTF) + r4(2 * ((p.PD + 3) & ~3)) + 256 + r4(4 * p.Tv) + r4(FA_BIAS_N));
  const size_t scratch = (p.att_type == 0 && p.A == 128) ? fast : generic;
  return 1024 + (size_t)TC_NSTAGE * TC_STAGE_BYTES + (size_t)FA_WSTAGES * FA_WSTAGE_BYTES + scratch;
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
  std::vector<__nv_bfloat16> img((size_t)TC_LSTM_CTAS * TC_IMG_BYTES / 2);
  std::vector<float> bias((size_t)TC_LSTM_CTAS * 64);
  // one [32 gate rows x 64 k] SWIZZLE_128B block: row n = gate*8 + u <-> column gate*U + cta*8 + u of the Keras kernel
  auto put_block = [&](__nv_bfloat16* blk, const std::vector<float>& W, int cta, int kb) {
    for (int n = 0; n < 32; ++n) {
      const int gate = n >> 3, u = n & 7;
      const size_t col = (size_t)gate * TC_U + cta * 8 + u;
      for (int k = 0; k < 64; ++k) blk[sw128_offset_bytes(n, k) / 2] = __float2bfloat16(W[(size_t)(kb * 64 + k) * 4 * TC_U + col]);
    }
  };
  for (int cta = 0; cta < TC_LSTM_CTAS; ++cta) {
    __nv_bfloat16* base = img.data() + (size_t)cta * TC_IMG_BYTES / 2;
    for (int kb = 0; kb < TC_NKB_X; ++kb) put_block(base + (TC_IMG_W1X + kb * TC_B_BYTES) / 2, *src[0], cta, kb);
    for (int kb = 0; kb < TC_NKB_H; ++kb) {
      put_block(base + (TC_IMG_WU + kb * 2 * TC_B_BYTES) / 2, *src[2], cta, kb);               // W2 (cell_1/kernel)
      put_block(base + (TC_IMG_WU + kb * 2 * TC_B_BYTES + TC_B_BYTES) / 2, *src[1], cta, kb);  // U1 (cell_0/recurrent_kernel)
      put_block(base + (TC_IMG_U2 + kb * TC_B_BYTES) / 2, *src[3], cta, kb);                   // U2 (cell_1/recurrent_kernel)
    }
    for (int n = 0; n < 32; ++n) {
      const int gate = n >> 3, u = n & 7;
      bias[(size_t)cta * 64 + n] = b0[(size_t)gate * TC_U + cta * 8 + u];
      bias[(size_t)cta * 64 + 32 + n] = b1[(size_t)gate * TC_U + cta * 8 + u];
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
  const size_t smem = bf16_smem_bytes(p);
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
  q.wimgA = nullptr; q.vproj_bf = nullptr;
  if (st.fast_a) {
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
  q.prof = st.prof;
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
