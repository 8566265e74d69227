// Persistent fp32 ("exact") decoder kernel.
//
// Replaces the reference's tf.while_loop over Decoder_Step.call (Modules/Taco2.py:96-120,153-228)
// with ONE cooperative launch: every CTA stays resident for the whole decode and the steps are
// separated by three grid-wide barriers.
//
//   phase A (per utterance, owner CTA = b mod grid):   projection of the previous step
//       (Taco2.py:113-118) -> prenet (Taco2.py:282-283) -> query (Steps.py:122) -> energies
//       (Steps.py:138-152 / Layers.py:393-407) -> alignment update (Steps.py:168-180, 215-229 /
//       Layers.py:409-424) -> context (Steps.py:164)
//   phase B (LSTMCell 0) and phase C (LSTMCell 1): each CTA owns groups of 8 hidden units (32 gate
//       columns) and streams its packed weight slice once per batch tile (weights stay in L2:
//       57.8 MB fp32 < 126 MB), fp32 FFMA, pointwise cell update fused (Keras LSTMCell).
//
// All math is fp32 with accurate expf/tanhf; this is the 1e-4 parity mode.  The bf16 tensor-core
// kernel (decoder_bf16.cuh) is the throughput mode.
#pragma once
#include "common.cuh"

namespace gstk {

constexpr int DEC_THREADS = 512;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int LSTM_HU = 8;  // hidden units per weight group => 32 gate columns

struct DecParams {
  int B, Tv, T, mode, rng_mode, att_type;
  int mel, r, P0, P1, A, U0, U1, PD;
  int lsa_filters, lsa_kernel, lsa_cumulate, lsa_smoothing;
  float drop_rate, drop_scale, sigmoid_noise;
  unsigned long long seed;
  unsigned int step_offset, row_offset;
  // weights
  const float *W0, *b0, *W1, *b1, *Wq, *bq, *att_v, *att_sb;
  const float *lsa_cw, *lsa_cb, *lsa_dw, *lsa_db, *lsa_bias;
  const float *Wp, *bp;
  const float *L1pk, *L1b, *L2pk, *L2b;
  // inputs
  const float* vproj;  // [B,Tv,A]
  const float* teacher;
  long long ts_b, ts_t;
  const float *keep0, *keep1, *noise, *init_mel;
  // state (workspace)
  float* xin;    // [B, P1 + A]
  float* h1;     // [2][B][U0]
  float* h2;     // [2][B][U1]
  float* c1;     // [B][U0]
  float* c2;     // [B][U1]
  float* align;  // [2][B][Tv]
  float* cum;    // [B][Tv]  (LSA)
  // outputs (device; may be null)
  float *out_mel, *out_stop, *out_align, *out_ctx;
  GridBarrier* gb;
  // [T, rngB, .] layout of keep0/keep1/noise and the first row this launch handles (batch chunking)
  int rngB, rng_b0;
  // bf16 tensor-core path only: K-major SWIZZLE_128B images of the LSTM inputs (see decoder_bf16.cuh)
  __nv_bfloat16* actX;
  int MT;
  int debug_flags;  // diagnostics only (GSTK_DEBUG env): bit 0 = skip the h2.U2 pre-accumulation segment
  int To;           // row count (steps) of the output tensors: == T unless a decode is split into several launches over time
  // early stop (Model.py:380 cuts every utterance at argmax(stop < 0)).  stop_index[b] = first step whose stop logit
  // is negative (STOP_UNSET until then; steps are counted from the start of the call: t_base + local step);
  // stop_state[0] = rows of this batch chunk that have stopped, stop_state[1] = steps whose outputs are valid when the kernel left.
  int early_stop, t_base;
  int* stop_index;
  unsigned int* stop_state;
};
constexpr int STOP_UNSET = 0x7fffffff;

// stop bookkeeping for one (row, step) stop logit; single writer per row
__device__ __forceinline__ void note_stop(const DecParams& p, int b, int step, float logit) {
  if (p.stop_index && logit < 0.f && p.stop_index[b] == STOP_UNSET) {
    p.stop_index[b] = p.t_base + step;
    atomicAdd(&p.stop_state[0], 1u);
  }
}

// CTA-subset barrier: NT == DEC_THREADS -> __syncthreads, otherwise named barrier 1 over threads [0, NT)
template <int NT>
__device__ __forceinline__ void pa_sync() {
  if constexpr (NT == DEC_THREADS) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// element k of batch row b inside a [kb][MT][128 rows][64] bf16 SWIZZLE_128B activation image
__device__ __forceinline__ size_t act_elem_index(int MT, int b, int k) {
  const int kb = k >> 6, kk = k & 63, mt = b >> 7, r = b & 127;
  const int chunk = (kk >> 3) ^ (r & 7);
  return ((size_t)(kb * MT + mt) * 128 + r) * 64 + (size_t)chunk * 8 + (kk & 7);
}

// out_s[n] = sum_k in_s[k] * W[k*N + n] for n < N (N <= DEC_THREADS). red_s: DEC_THREADS floats.
template <int NT>
__device__ __forceinline__ void cta_gemv(const float* __restrict__ W, int K, int N, const float* in_s,
                                         float* out_s, float* red_s) {
  const int NP = (N + 31) & ~31;
  const int KG = NT / NP;
  const int n = threadIdx.x % NP, kg = threadIdx.x / NP;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (kg < KG && n < N) {
    int k = kg;
    for (; k + 3 * KG < K; k += 4 * KG) {
      const float w0 = __ldg(W + (size_t)k * N + n);
      const float w1 = __ldg(W + (size_t)(k + KG) * N + n);
      const float w2 = __ldg(W + (size_t)(k + 2 * KG) * N + n);
      const float w3 = __ldg(W + (size_t)(k + 3 * KG) * N + n);
      a0 = fmaf(in_s[k], w0, a0);
      a1 = fmaf(in_s[k + KG], w1, a1);
      a2 = fmaf(in_s[k + 2 * KG], w2, a2);
      a3 = fmaf(in_s[k + 3 * KG], w3, a3);
    }
    for (; k < K; k += KG) a0 = fmaf(in_s[k], __ldg(W + (size_t)k * N + n), a0);
  }
  red_s[threadIdx.x] = (a0 + a1) + (a2 + a3);
  pa_sync<NT>();
  if (threadIdx.x < N) {
    float s = 0.f;
    for (int g = 0; g < KG; ++g) s += red_s[g * NP + threadIdx.x];
    out_s[threadIdx.x] = s;
  }
  pa_sync<NT>();
}

struct PhaseASmem {
  float* x;     // [mel]
  float* y;     // [PD]
  float* hc;    // [U1 + A]
  float* p0;    // [P0]
  float* p1;    // [P1]
  float* q;     // [A]
  float* e;     // [Tv]
  float* al;    // [Tv]
  float* prev;  // [Tv]
  float* src;   // [Tv]  (LSA location source)
  float* red;   // [DEC_THREADS]
  float* scal;  // [8]
};

template <int NT>
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red, float* scal) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  if (lane == 0) red[wid] = v;
  pa_sync<NT>();
  if (wid == 0) {
    float x = lane < NT / 32 ? red[lane] : (is_max ? -INFINITY : 0.f);
    x = is_max ? warp_max(x) : warp_sum(x);
    if (lane == 0) scal[0] = x;
  }
  pa_sync<NT>();
  const float out = scal[0];
  pa_sync<NT>();
  return out;
}

// One utterance, iteration t in [0, T]: projection of step t-1, then (t < T) the front end of step t.
template <int NT>
__device__ __noinline__ void phase_a_utt(const DecParams& p, const PhaseASmem s, int b, int t) {
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cur = t & 1, prv = cur ^ 1;
  const int XW = p.P1 + p.A;
  if (t > 0) {
    // ---- projection of step t-1 : [h2 || ctx] . Wp + bp  (Taco2.py:112-118)
    const float* h2 = p.h2 + ((size_t)prv * p.B + b) * p.U1;
    for (int i = tid; i < p.U1; i += NT) s.hc[i] = __ldcg(h2 + i);
    for (int i = tid; i < p.A; i += NT) s.hc[p.U1 + i] = __ldcg(p.xin + (size_t)b * XW + p.P1 + i);
    pa_sync<NT>();
    cta_gemv<NT>(p.Wp, p.U1 + p.A, p.PD, s.hc, s.y, s.red);
    if (tid < p.PD) {
      const float v = s.y[tid] + __ldg(p.bp + tid);
      s.y[tid] = v;
      if (tid < p.PD - 1) {
        if (p.out_mel) p.out_mel[((size_t)b * p.To + (t - 1)) * (p.PD - 1) + tid] = v;
      } else {
        if (p.out_stop) p.out_stop[(size_t)b * p.To + (t - 1)] = v;
        note_stop(p, b, t - 1, v);
      }
    }
    pa_sync<NT>();
  }
  if (t == p.T) return;
  // ---- decoder input (Taco2.py:183-187)
  if (tid < p.mel) {
    float v;
    if (p.mode == 1) {
      v = __ldg(p.teacher + (size_t)b * p.ts_b + (size_t)t * p.ts_t + tid);
    } else if (t == 0) {
      v = p.init_mel ? __ldg(p.init_mel + (size_t)b * p.mel + tid) : 0.f;
    } else {
      v = s.y[(p.r - 1) * p.mel + tid];  // last of the r frames (decodings[:, -1])
    }
    s.x[tid] = v;
  }
  pa_sync<NT>();
  const unsigned int step_id = p.step_offset + (unsigned int)t, row_id = p.row_offset + (unsigned int)b;
  const bool drop = p.rng_mode != 0 && p.drop_rate > 0.f;
  // ---- prenet layer 0
  cta_gemv<NT>(p.W0, p.mel, p.P0, s.x, s.p0, s.red);
  if (tid < p.P0) {
    float v = fmaxf(s.p0[tid] + __ldg(p.b0 + tid), 0.f);
    if (drop) {
      const float keep = p.rng_mode == 1 ? __ldg(p.keep0 + ((size_t)t * p.rngB + p.rng_b0 + b) * p.P0 + tid)
                                         : philox_keep(p.seed, STREAM_KEEP0, step_id, row_id, tid, p.drop_rate);
      v = v * keep * p.drop_scale;
    }
    s.p0[tid] = v;
  }
  pa_sync<NT>();
  // ---- prenet layer 1
  cta_gemv<NT>(p.W1, p.P0, p.P1, s.p0, s.p1, s.red);
  if (tid < p.P1) {
    float v = fmaxf(s.p1[tid] + __ldg(p.b1 + tid), 0.f);
    if (drop) {
      const float keep = p.rng_mode == 1 ? __ldg(p.keep1 + ((size_t)t * p.rngB + p.rng_b0 + b) * p.P1 + tid)
                                         : philox_keep(p.seed, STREAM_KEEP1, step_id, row_id, tid, p.drop_rate);
      v = v * keep * p.drop_scale;
    }
    s.p1[tid] = v;
    p.xin[(size_t)b * XW + tid] = v;
    if (p.actX) p.actX[act_elem_index(p.MT, b, tid)] = __float2bfloat16(v);
  }
  pa_sync<NT>();
  // ---- query projection (Steps.py:122)
  cta_gemv<NT>(p.Wq, p.P1, p.A, s.p1, s.q, s.red);
  if (tid < p.A) s.q[tid] += __ldg(p.bq + tid);
  // ---- previous alignment (and LSA location source)
  const float* prev_g = p.align + ((size_t)prv * p.B + b) * p.Tv;
  for (int j = tid; j < p.Tv; j += NT) {
    s.prev[j] = __ldcg(prev_g + j);
    if (p.att_type == 2) s.src[j] = p.lsa_cumulate ? __ldcg(p.cum + (size_t)b * p.Tv + j) : s.prev[j];
  }
  pa_sync<NT>();
  // ---- energies: one warp per memory position
  const float* V = p.vproj + (size_t)b * p.Tv * p.A;
  const bool noisy = p.rng_mode != 0 && p.sigmoid_noise > 0.f && p.att_type != 2;
  for (int j = wid; j < p.Tv; j += NW) {
    float acc = 0.f;
    if (p.att_type != 2) {
      for (int a = lane; a < p.A; a += 32)
        acc = fmaf(__ldg(p.att_v + a), tanhf(s.q[a] + __ldg(V + (size_t)j * p.A + a)), acc);
    } else {
      // location features: Dense_A(Conv1D(src))  (Layers.py:362-364), 'same' padding, stride 1
      float cf = 0.f;
      if (lane < p.lsa_filters) {
        const int pad = (p.lsa_kernel - 1) / 2;
        cf = __ldg(p.lsa_cb + lane);
        for (int k = 0; k < p.lsa_kernel; ++k) {
          const int jj = j + k - pad;
          if (jj >= 0 && jj < p.Tv) cf = fmaf(s.src[jj], __ldg(p.lsa_cw + k * p.lsa_filters + lane), cf);
        }
      }
      for (int a = lane; a < p.A; a += 32) {
        float loc = __ldg(p.lsa_db + a);
        for (int f = 0; f < p.lsa_filters; ++f)
          loc = fmaf(__shfl_sync(0xffffffffu, cf, f), __ldg(p.lsa_dw + f * p.A + a), loc);
        acc += tanhf(s.q[a] + __ldg(V + (size_t)j * p.A + a) + loc + __ldg(p.lsa_bias + a));
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (p.att_type != 2) acc += __ldg(p.att_sb);
      if (noisy) {
        float nz;
        if (p.rng_mode == 1) {
          nz = __ldg(p.noise + ((size_t)t * p.rngB + p.rng_b0 + b) * p.Tv + j);
        } else {
          const float4 z = philox_normal4(p.seed, step_id, row_id, (unsigned int)j >> 2);
          const int w = j & 3;
          nz = w == 0 ? z.x : (w == 1 ? z.y : (w == 2 ? z.z : z.w));
        }
        acc = fmaf(p.sigmoid_noise, nz, acc);
      }
      s.e[j] = acc;
    }
  }
  pa_sync<NT>();
  // ---- alignment update
  if (p.att_type == 0) {  // SMA, Steps.py:215-229
    for (int j = tid; j < p.Tv; j += NT) {
      const float pj = sigmoid_acc(s.e[j]);
      float v = s.prev[j] * pj;
      if (j > 0) v += s.prev[j - 1] * (1.0f - sigmoid_acc(s.e[j - 1]));
      s.al[j] = v;
    }
  } else if (p.att_type == 1) {  // BMA, Steps.py:168-198 : two prefix scans, warp 0 with carries
    if (wid == 0) {
      float carry_log = 0.f, carry_sum = 0.f;
      for (int j0 = 0; j0 < p.Tv; j0 += 32) {
        const int j = j0 + lane;
        const bool ok = j < p.Tv;
        const float pj = ok ? sigmoid_acc(s.e[j]) : 0.f;
        const float lg = ok ? logf(fminf(fmaxf(1.0f - pj, 1.17549435e-38f), 1.0f)) : 0.f;
        float inc = lg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += n;
        }
        const float cp = expf(carry_log + inc - lg);  // exclusive cumsum
        const float term = ok ? s.prev[j] / fminf(fmaxf(cp, 1e-10f), 1.0f) : 0.f;
        float cs = term;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float n = __shfl_up_sync(0xffffffffu, cs, o);
          if (lane >= o) cs += n;
        }
        if (ok) s.al[j] = pj * cp * (carry_sum + cs);
        carry_log += __shfl_sync(0xffffffffu, inc, 31);
        carry_sum += __shfl_sync(0xffffffffu, cs, 31);
      }
    }
  } else {  // LSA, Layers.py:409-444
    if (!p.lsa_smoothing) {
      float m = -INFINITY;
      for (int j = tid; j < p.Tv; j += NT) m = fmaxf(m, s.e[j]);
      m = block_reduce<NT>(m, true, s.red, s.scal);
      float sum = 0.f;
      for (int j = tid; j < p.Tv; j += NT) {
        const float v = expf(s.e[j] - m);
        s.al[j] = v;
        sum += v;
      }
      sum = block_reduce<NT>(sum, false, s.red, s.scal);
      for (int j = tid; j < p.Tv; j += NT) s.al[j] = s.al[j] / sum;
    } else {
      float sum = 0.f;
      for (int j = tid; j < p.Tv; j += NT) {
        const float v = sigmoid_acc(s.e[j]);
        s.al[j] = v;
        sum += v;
      }
      sum = block_reduce<NT>(sum, false, s.red, s.scal);
      for (int j = tid; j < p.Tv; j += NT) s.al[j] = s.al[j] / sum;
    }
  }
  pa_sync<NT>();
  float* al_g = p.align + ((size_t)cur * p.B + b) * p.Tv;
  for (int j = tid; j < p.Tv; j += NT) {
    const float v = s.al[j];
    al_g[j] = v;
    if (p.out_align) p.out_align[((size_t)b * p.To + t) * p.Tv + j] = v;
    if (p.att_type == 2) p.cum[(size_t)b * p.Tv + j] = s.src[j] * (p.lsa_cumulate ? 1.f : 0.f) + v;
  }
  // ---- context = alignment . V'  (Steps.py:164)
  {
    const int JG = NT / p.A;  // A <= NT
    const int a = tid % p.A, jg = tid / p.A;
    float acc = 0.f;
    if (jg < JG)
      for (int j = jg; j < p.Tv; j += JG) acc = fmaf(s.al[j], __ldg(V + (size_t)j * p.A + a), acc);
    s.red[tid] = acc;
    pa_sync<NT>();
    if (tid < p.A) {
      float c = 0.f;
      for (int g = 0; g < JG; ++g) c += s.red[g * p.A + tid];
      p.xin[(size_t)b * XW + p.P1 + tid] = c;
      if (p.actX) p.actX[act_elem_index(p.MT, b, p.P1 + tid)] = __float2bfloat16(c);
      if (p.out_ctx && t == p.T - 1) p.out_ctx[(size_t)b * p.A + tid] = c;
    }
    pa_sync<NT>();
  }
}

// LSTMCell `layer` for all batch rows; CTA g handles unit groups g, g+grid, ...
// Packed weights: Wpk[group][k4][col(32)][4] with col = gate*8 + unit_in_group and k over
// [kernel rows ; recurrent_kernel rows].
template <int BT>
__device__ void lstm_phase(const DecParams& p, int layer, int t, float* smem) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cur = t & 1, prv = cur ^ 1;
  const int U = layer == 0 ? p.U0 : p.U1;
  const int Kx = layer == 0 ? p.P1 + p.A : p.U0;
  const int K = Kx + U, K4 = K >> 2;
  const float* xsrc = layer == 0 ? p.xin : p.h1 + (size_t)cur * p.B * p.U0;
  const float* hprev = (layer == 0 ? p.h1 : p.h2) + (size_t)prv * p.B * U;
  float* hnew = (layer == 0 ? p.h1 : p.h2) + (size_t)cur * p.B * U;
  float* cst = layer == 0 ? p.c1 : p.c2;
  const float* Wpk = layer == 0 ? p.L1pk : p.L2pk;
  const float* bias = layer == 0 ? p.L1b : p.L2b;
  float* in_s = smem;                       // [BT][K]
  float* red_s = smem + (size_t)BT * K;     // [DEC_WARPS][BT][32]
  const int groups = U / LSTM_HU;
  const int k4_per_warp = (K4 + DEC_WARPS - 1) / DEC_WARPS;
  const int k4_lo = wid * k4_per_warp, k4_hi = min(K4, k4_lo + k4_per_warp);
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const float4* Wg = reinterpret_cast<const float4*>(Wpk) + (size_t)g * K4 * 32;
    for (int b0 = 0; b0 < p.B; b0 += BT) {
      const int nb = min(BT, p.B - b0);
      for (int i = tid; i < BT * K; i += DEC_THREADS) {
        const int bb = i / K, k = i - bb * K;
        float v = 0.f;
        if (bb < nb) {
          const int b = b0 + bb;
          v = k < Kx ? __ldcg(xsrc + (size_t)b * Kx + k) : __ldcg(hprev + (size_t)b * U + (k - Kx));
        }
        in_s[i] = v;
      }
      __syncthreads();
      float acc[BT];
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) acc[bb] = 0.f;
#pragma unroll 4
      for (int k4 = k4_lo; k4 < k4_hi; ++k4) {
        const float4 w = __ldg(Wg + (size_t)k4 * 32 + lane);
#pragma unroll
        for (int bb = 0; bb < BT; ++bb) {
          const float4 a = *reinterpret_cast<const float4*>(in_s + (size_t)bb * K + (k4 << 2));
          acc[bb] = fmaf(a.x, w.x, acc[bb]);
          acc[bb] = fmaf(a.y, w.y, acc[bb]);
          acc[bb] = fmaf(a.z, w.z, acc[bb]);
          acc[bb] = fmaf(a.w, w.w, acc[bb]);
        }
      }
#pragma unroll
      for (int bb = 0; bb < BT; ++bb) red_s[(wid * BT + bb) * 32 + lane] = acc[bb];
      __syncthreads();
      if (tid < nb * LSTM_HU) {
        const int bb = tid / LSTM_HU, u = tid % LSTM_HU;
        const int b = b0 + bb, unit = g * LSTM_HU + u;
        float z[4];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          float sgm = 0.f;
          for (int w = 0; w < DEC_WARPS; ++w) sgm += red_s[(w * BT + bb) * 32 + gate * LSTM_HU + u];
          z[gate] = sgm + __ldg(bias + gate * U + unit);
        }
        const float c_old = cst[(size_t)b * U + unit];
        const float c_new = sigmoid_acc(z[1]) * c_old + sigmoid_acc(z[0]) * tanhf(z[2]);
        const float h_new = sigmoid_acc(z[3]) * tanhf(c_new);
        cst[(size_t)b * U + unit] = c_new;
        hnew[(size_t)b * U + unit] = h_new;
      }
      __syncthreads();
    }
  }
}

template <int BT>
__global__ void __launch_bounds__(DEC_THREADS, 1) decoder_fp32_kernel(const __grid_constant__ DecParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int ok_s;
  PhaseASmem s;
  {
    float* q = smem;
    auto take = [&](int n) { float* r = q; q += (n + 3) & ~3; return r; };
    s.x = take(p.mel);
    s.y = take(p.PD);
    s.hc = take(p.U1 + p.A);
    s.p0 = take(p.P0);
    s.p1 = take(p.P1);
    s.q = take(p.A);
    s.e = take(p.Tv);
    s.al = take(p.Tv);
    s.prev = take(p.Tv);
    s.src = take(p.Tv);
    s.red = take(DEC_THREADS);
    s.scal = take(8);
  }
  unsigned int gen = 0;
  for (int t = 0; t <= p.T; ++t) {
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) phase_a_utt<DEC_THREADS>(p, s, b, t);
    if (t == p.T) break;
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    lstm_phase<BT>(p, 0, t, smem);
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
    lstm_phase<BT>(p, 1, t, smem);
    if (!grid_sync(p.gb, gridDim.x, gen, &ok_s)) return;
  }
}

inline size_t decoder_fp32_smem_bytes(const DecParams& p, int BT) {
  auto r4 = [](int n) { return (size_t)((n + 3) & ~3); };
  const size_t a = r4(p.mel) + r4(p.PD) + r4(p.U1 + p.A) + r4(p.P0) + r4(p.P1) + r4(p.A) + 4 * r4(p.Tv) +
                   DEC_THREADS + 8;
  const int Kmax = max(p.P1 + p.A + p.U0, p.U0 + p.U1);
  const size_t b = (size_t)BT * Kmax + (size_t)DEC_WARPS * BT * 32;
  return 4 * (a > b ? a : b);
}

}  // namespace gstk
