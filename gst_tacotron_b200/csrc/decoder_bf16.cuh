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
constexpr int TC_NWB = TC_NKB_X + 3 * TC_NKB_H;  // 54 weight blocks per CTA
constexpr int TC_WB_W1X = 0, TC_WB_U1 = TC_NKB_X, TC_WB_W2 = TC_NKB_X + TC_NKB_H, TC_WB_U2 = TC_NKB_X + 2 * TC_NKB_H;
constexpr int TC_LSTM_CTAS = TC_U / 8;  // 128
constexpr int TC_RES_WB = 38;       // resident weight blocks: W1x (6) + W2 (16) + U1 (16); U2 is streamed
constexpr int TC_NSTAGE = 3;
constexpr int TC_A_BYTES = 128 * 128;   // one activation tile (128 rows x 64 bf16)
constexpr int TC_B_BYTES = 32 * 128;    // one weight block
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_PA_THREADS = 480;  // warps 0-14 run phase A; warp 15 is the copy/MMA warp
constexpr int TC_TMEM_COLS = 128;   // D1: cols [0,64) (2 m-tiles x 32), D2: cols [64,128)
constexpr int TC_MAX_B = 256;

// resident slot of weight block wb, or -1 when it is streamed
__host__ __device__ constexpr int tc_res_slot(int wb) {
  // order of residency: W1x, W2, U1 (critical-path operands first)
  return wb < TC_NKB_X ? wb
         : (wb >= TC_WB_W2 && wb < TC_WB_U2) ? TC_NKB_X + (wb - TC_WB_W2)
         : (wb >= TC_WB_U1 && wb < TC_WB_W2) ? TC_NKB_X + TC_NKB_H + (wb - TC_WB_U1)
                                              : -1;
}

static_assert(tc_res_slot(TC_WB_W2) >= 0 && tc_res_slot(TC_WB_U1 + TC_NKB_H - 1) >= 0 && tc_res_slot(TC_WB_W1X) >= 0,
              "the operands of the two-MMA segment (W2, U1) and W1x must be resident");
static_assert(TC_NKB_X + 2 * TC_NKB_H == TC_RES_WB, "residency table and TC_RES_WB disagree");

struct Bf16Params {
  const __nv_bfloat16* wimg;  // [TC_LSTM_CTAS][TC_NWB][32][64] swizzled
  const float* bias;          // [TC_LSTM_CTAS][2][32]  (gate*8+u)
  __nv_bfloat16* actX;        // [6][MT][128][64]
  __nv_bfloat16* actH1;       // [16][MT][128][64]
  __nv_bfloat16* actH2;       // [16][MT][128][64]
};

struct TcPipe {
  uint64_t* full;
  uint64_t* empty;
  uint8_t* stages;
  const uint8_t* wres;
  const uint8_t* wimg_cta;  // this CTA's 54-block image in global memory
  uint32_t it;              // units issued so far (ring position)
  uint32_t tmem;
  int MT, B;
};

// One segment of the copy/MMA pipeline.  Unit u = (mt, kb): A tile = act[kb][mt]; up to two MMAs
// use it: (wb0 -> D column d0, accumulate flag acc0_first for kb == 0) and optionally (wb1 -> d1).
__device__ __forceinline__ void tc_segment(TcPipe& pp, const __nv_bfloat16* act, int nkb, int wb0_base, uint32_t d0_col,
                                           bool acc0_first, int wb1_base, uint32_t d1_col, bool acc1_first,
                                           uint64_t* commit_after_wb0) {
  const int n = pp.MT * nkb;
  const uint32_t idesc = make_idesc_bf16(128, 32);
  int issued = 0;
  for (int done = 0; done < n; ++done) {
    while (issued < n && issued < done + TC_NSTAGE) {
      const uint32_t g = pp.it + issued;
      const uint32_t s = g % TC_NSTAGE;
      mbar_wait(&pp.empty[s], ((g / TC_NSTAGE) & 1u) ^ 1u);
      const int mt = issued / nkb, kb = issued % nkb;
      const int rows = min(128, pp.B - mt * 128);
      uint32_t bytes = (uint32_t)rows * 128u;
      const bool s0 = tc_res_slot(wb0_base + kb) < 0;
      const bool s1 = !s0 && wb1_base >= 0 && tc_res_slot(wb1_base + kb) < 0;  // one streamed block per unit
      if (s0) bytes += TC_B_BYTES;
      if (s1) bytes += TC_B_BYTES;
      mbar_arrive_expect_tx(&pp.full[s], bytes);
      uint8_t* st = pp.stages + (size_t)s * TC_STAGE_BYTES;
      bulk_g2s(st, act + ((size_t)(kb * pp.MT + mt) * 128) * 64, (uint32_t)rows * 128u, &pp.full[s]);
      // at most one of the two weight blocks of a unit is streamed with the default residency
      if (s0) bulk_g2s(st + TC_A_BYTES, pp.wimg_cta + (size_t)(wb0_base + kb) * TC_B_BYTES, TC_B_BYTES, &pp.full[s]);
      else if (s1) bulk_g2s(st + TC_A_BYTES, pp.wimg_cta + (size_t)(wb1_base + kb) * TC_B_BYTES, TC_B_BYTES, &pp.full[s]);
      ++issued;
    }
    const uint32_t g = pp.it + done;
    const uint32_t s = g % TC_NSTAGE;
    mbar_wait(&pp.full[s], (g / TC_NSTAGE) & 1u);
    tc_fence_after();
    const int mt = done / nkb, kb = done % nkb;
    uint8_t* st = pp.stages + (size_t)s * TC_STAGE_BYTES;
    const uint64_t ad = make_desc_sw128(smem_u32(st));
    {
      const int slot = tc_res_slot(wb0_base + kb);
      const uint64_t bd = make_desc_sw128(slot >= 0 ? smem_u32(pp.wres + (size_t)slot * TC_B_BYTES) : smem_u32(st + TC_A_BYTES));
      const uint32_t dcol = pp.tmem + d0_col + (uint32_t)mt * 32u;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(dcol, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0 || !acc0_first) ? 1u : 0u);
    }
    if (commit_after_wb0 && done == n - 1) umma_commit(commit_after_wb0);
    if (wb1_base >= 0) {
      const int slot = tc_res_slot(wb1_base + kb);
      const uint64_t bd = make_desc_sw128(slot >= 0 ? smem_u32(pp.wres + (size_t)slot * TC_B_BYTES) : smem_u32(st + TC_A_BYTES));
      const uint32_t dcol = pp.tmem + d1_col + (uint32_t)mt * 32u;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(dcol, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0 || !acc1_first) ? 1u : 0u);
    }
    umma_commit(&pp.empty[s]);
  }
  pp.it += n;
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

__global__ void __launch_bounds__(DEC_THREADS, 1) decoder_bf16_kernel(const DecParams p, const Bf16Params q) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  __shared__ int ok_s;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * TC_NSTAGE + 3];
  __shared__ float bias_s[64];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cta = blockIdx.x;
  const bool lstm_cta = cta < TC_LSTM_CTAS;
  uint8_t* wres = sm;                                              // TC_RES_WB x 4 KB
  uint8_t* stages = sm + (size_t)TC_RES_WB * TC_B_BYTES;            // TC_NSTAGE x 20 KB
  float* scratch = reinterpret_cast<float*>(stages + (size_t)TC_NSTAGE * TC_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + TC_NSTAGE;
  uint64_t* d1_full = bars + 2 * TC_NSTAGE;
  uint64_t* d2_full = bars + 2 * TC_NSTAGE + 1;
  uint64_t* wres_full = bars + 2 * TC_NSTAGE + 2;

  PhaseASmem s;
  {
    float* f = scratch;
    auto take = [&](int n) { float* r = f; f += (n + 3) & ~3; return r; };
    s.x = take(p.mel); s.y = take(p.PD); s.hc = take(p.U1 + p.A); s.p0 = take(p.P0); s.p1 = take(p.P1);
    s.q = take(p.A); s.e = take(p.Tv); s.al = take(p.Tv); s.prev = take(p.Tv); s.src = take(p.Tv);
    s.red = take(DEC_THREADS); s.scal = take(8);
  }
  if (tid == 0) {
    for (int i = 0; i < TC_NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(d1_full, 1);
    mbar_init(d2_full, 1);
    mbar_init(wres_full, 1);
    mbar_fence_init();
  }
  if (lstm_cta && tid < 64) bias_s[tid] = __ldg(q.bias + (size_t)cta * 64 + tid);
  if (lstm_cta && wid == 0) tmem_alloc(&tmem_base_s, TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int MT = p.MT;
  const bool tensor_thread = lstm_cta && wid == 15 && lane == 0;
  TcPipe pp;
  pp.full = full; pp.empty = empty; pp.stages = stages; pp.wres = wres;
  pp.wimg_cta = reinterpret_cast<const uint8_t*>(q.wimg) + (size_t)cta * TC_NWB * TC_B_BYTES;
  pp.it = 0; pp.tmem = lstm_cta ? tmem_base_s : 0; pp.MT = MT; pp.B = p.B;

  // cell state of (row, this CTA's 8 units) lives in registers of epilogue warps 0..4*MT-1
  const bool epi = lstm_cta && wid < 4 * MT;
  const int erow = (wid >> 2) * 128 + (wid & 3) * 32 + lane;  // batch row of this epilogue thread
  const bool erow_ok = epi && erow < p.B;
  float c1[8], c2[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    c1[u] = erow_ok ? __ldcg(p.c1 + (size_t)erow * TC_U + cta * 8 + u) : 0.f;
    c2[u] = erow_ok ? __ldcg(p.c2 + (size_t)erow * TC_U + cta * 8 + u) : 0.f;
  }

  if (tensor_thread) {
    // resident weights: one barrier, TC_RES_WB bulk copies
    mbar_arrive_expect_tx(wres_full, (uint32_t)TC_RES_WB * TC_B_BYTES);
    for (int wb = 0; wb < TC_NWB; ++wb) {
      const int slot = tc_res_slot(wb);
      if (slot >= 0) bulk_g2s(wres + (size_t)slot * TC_B_BYTES, pp.wimg_cta + (size_t)wb * TC_B_BYTES, TC_B_BYTES, wres_full);
    }
    mbar_wait(wres_full, 0);
    fence_proxy_async();
    // prologue: D1 = h1(-1) . U1 (the images of the initial states were packed by the host-side kernel)
    tc_segment(pp, q.actH1, TC_NKB_H, TC_WB_U1, 0u, true, -1, 0u, false, nullptr);
  }

  unsigned int gen = 0;
  for (int t = 0; t <= p.T; ++t) {
    // ---------------- phase A (+ overlapped: D2 = h2(t-1) . U2) --------------------------------
    if (wid < 15) {
      for (int b = cta; b < p.B; b += gridDim.x) phase_a_utt<TC_PA_THREADS>(p, s, b, t);
      fence_proxy_async();  // actX stores (generic proxy) -> later bulk copies (async proxy)
    } else if (tensor_thread && t < p.T) {
      fence_proxy_async();
      tc_segment(pp, q.actH2, TC_NKB_H, TC_WB_U2, 64u, true, -1, 0u, false, nullptr);
    }
    if (t == p.T) break;
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    // ---------------- phase B: LSTMCell 0 -------------------------------------------------------
    if (tensor_thread) {
      fence_proxy_async();
      tc_segment(pp, q.actX, TC_NKB_X, TC_WB_W1X, 0u, false, -1, 0u, false, d1_full);
    }
    if (epi) {
      mbar_wait(d1_full, (uint32_t)t & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld32(pp.tmem + ((uint32_t)((wid & 3) * 32) << 16) + 0u + (uint32_t)(wid >> 2) * 32u, v);
      if (erow_ok)
        tc_epilogue_row(v, bias_s, c1, erow, cta, MT, q.actH1, p.h1 + ((size_t)(t & 1) * p.B + erow) * TC_U);
      tc_fence_before();
      fence_proxy_async();
    }
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    // ---------------- phase C: LSTMCell 1 (+ D1 = h1(t) . U1 for the next step) ------------------
    if (tensor_thread) {
      fence_proxy_async();
      tc_fence_after();
      tc_segment(pp, q.actH1, TC_NKB_H, TC_WB_W2, 64u, false, TC_WB_U1, 0u, true, d2_full);
    }
    if (epi) {
      mbar_wait(d2_full, (uint32_t)t & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld32(pp.tmem + ((uint32_t)((wid & 3) * 32) << 16) + 64u + (uint32_t)(wid >> 2) * 32u, v);
      if (erow_ok)
        tc_epilogue_row(v, bias_s + 32, c2, erow, cta, MT, q.actH2, p.h2 + ((size_t)(t & 1) * p.B + erow) * TC_U);
      tc_fence_before();
      fence_proxy_async();
    }
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    if (tensor_thread) tc_fence_after();
  }
  // final cell states (h is already in p.h1 / p.h2)
  if (erow_ok) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      p.c1[(size_t)erow * TC_U + cta * 8 + u] = c1[u];
      p.c2[(size_t)erow * TC_U + cta * 8 + u] = c2[u];
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
  bool ready = false;
};

inline size_t bf16_smem_bytes(const DecParams& p) {
  auto r4 = [](int n) { return (size_t)((n + 3) & ~3); };
  const size_t scratch = 4 * (r4(p.mel) + r4(p.PD) + r4(p.U1 + p.A) + r4(p.P0) + r4(p.P1) + r4(p.A) + 4 * r4(p.Tv) +
                              DEC_THREADS + 8);
  return 1024 + (size_t)TC_RES_WB * TC_B_BYTES + (size_t)TC_NSTAGE * TC_STAGE_BYTES + scratch;
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
  std::vector<__nv_bfloat16> img((size_t)TC_LSTM_CTAS * TC_NWB * 32 * 64);
  std::vector<float> bias((size_t)TC_LSTM_CTAS * 64);
  for (int cta = 0; cta < TC_LSTM_CTAS; ++cta) {
    for (int wb = 0; wb < TC_NWB; ++wb) {
      int m, kb;
      if (wb < TC_WB_U1) { m = 0; kb = wb; }
      else if (wb < TC_WB_W2) { m = 1; kb = wb - TC_WB_U1; }
      else if (wb < TC_WB_U2) { m = 2; kb = wb - TC_WB_W2; }
      else { m = 3; kb = wb - TC_WB_U2; }
      const std::vector<float>& W = *src[m];
      __nv_bfloat16* blk = img.data() + ((size_t)cta * TC_NWB + wb) * 32 * 64;
      for (int n = 0; n < 32; ++n) {
        const int gate = n >> 3, u = n & 7;
        const size_t col = (size_t)gate * TC_U + cta * 8 + u;
        for (int k = 0; k < 64; ++k)
          blk[sw128_offset_bytes(n, k) / 2] = __float2bfloat16(W[(size_t)(kb * 64 + k) * 4 * TC_U + col]);
      }
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
  st.ready = true;
  return GSTK_OK;
}

inline void bf16_release(Bf16State& st) {
  cudaFree(st.wimg);
  cudaFree(st.bias);
  cudaFree(st.act);
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
  // operand images of the initial hidden states (rows >= B zero so that unused tile rows stay finite)
  if ((e = cudaMemsetAsync(q.actX, 0, (size_t)TC_NKB_X * MT * 128 * 64 * 2, stream)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h1 + (size_t)p.B * TC_U, q.actH1, p.B, MT);
  pack_act_kernel<<<num_sms, 256, 0, stream>>>(p.h2 + (size_t)p.B * TC_U, q.actH2, p.B, MT);
  launches += 2;
  if ((e = cudaFuncSetAttribute(decoder_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  int occ = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decoder_bf16_kernel, DEC_THREADS, smem)) != cudaSuccess)
    return fail(GSTK_ECUDA, cudaGetErrorString(e));
  if (occ < 1) return fail(GSTK_EINVAL, "bf16 decoder kernel does not fit on an SM");
  void* args[] = {&p, &q};
  cudaEventRecord(ev0, stream);
  if ((e = cudaLaunchCooperativeKernel((void*)decoder_bf16_kernel, dim3(num_sms), dim3(DEC_THREADS), args, smem, stream)) !=
      cudaSuccess)
    return fail(GSTK_ECUDA, std::string("cooperative launch failed: ") + cudaGetErrorString(e));
  cudaEventRecord(ev1, stream);
  launches += 1;
  return GSTK_OK;
}

}  // namespace gstk
