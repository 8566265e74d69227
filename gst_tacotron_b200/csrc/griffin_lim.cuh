// Audio.inv_spectrogram (reference: Audio.py:23-27) = de-normalise -> dB to amplitude -> ^power -> Griffin-Lim (Audio.py:57-68, on
// librosa.stft / istft with center=True, reflect padding, periodic Hann window, win_length == n_fft) -> inverse pre-emphasis
// (Audio.py:14-15), for a batch of utterances of different lengths at once.  Layout: spectrogram / magnitudes S [B][T][F]
// (the vocoder's own layout; the reference transposes one utterance at a time, Model.py:413), F = n_fft / 2 + 1.
//
//   gl_magnitude_kernel   S = (10 ^ ((denorm(spec) + ref_level_db) / 20)) ^ power
//   gl_frames_kernel      one CTA = TWO frames (t0, t0 + 1) of one utterance packed into ONE complex FFT of n_fft points:
//                           iteration 0: unit phases exp(2 pi i u) from the caller's uniforms (np.random.rand, Audio.py:61)
//                           later:       z = w * (y_pad[t0 hop ..] + i y_pad[(t0 + 1) hop ..]) -> FFT -> split by Hermitian symmetry
//                                        -> unit phases  X / |X|   (np.angle(0) = 0 -> 1)
//                           then W = S0 e^{i a0} + i S1 e^{i a1} (Hermitian extension of both) -> inverse FFT -> real part = frame t0,
//                           imaginary part = frame t0 + 1 -> * window -> frames[b][t][n_fft]
//                         Stockham radix-4 (+ one radix-2 stage when log2(n_fft) is odd) in shared memory, n_fft / 4 threads.
//   gl_overlap_add_kernel y[n] = sum_t frames[t][n + n_fft/2 - t hop] / sum_t w^2[..]  (librosa's window-sum-square normalisation,
//                         n_fft / 2 samples dropped at both ends): hop (T_b - 1) samples per utterance
//   gl_deemphasis_kernel  out[n] = y[n] + c out[n - 1]  (scipy.signal.lfilter([1], [1, -c])): one CTA per utterance, the linear
//                         recurrence as a block scan of affine maps
#pragma once
#include "common.cuh"

namespace gstk {

struct GlParams {
  const float* S;          // [B][T][F] magnitudes
  const float* uniform;    // [B][T][F] or nullptr (iteration > 0)
  const float* y;          // [B][Lmax] current signal (iteration > 0)
  float* frames;           // [B][T][N]
  const float* window;     // [N]
  const float2* tw;        // [N] e^{-2 pi i k / N}
  const int* lengths;      // [B] frames per utterance, or nullptr (= T)
  int B, T, F, N, hop, Lmax, log2n;
  int init;                // 1: first pass - phases from `uniform` (or Philox when it is nullptr) instead of from the STFT of y
  int row_offset;          // Philox row of utterance 0
  unsigned long long seed;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place-by-ping-pong Stockham FFT of N points, forward sign, N / 4 threads; returns the buffer holding the result
__device__ float2* gl_fft(float2* a, float2* b, const float2* __restrict__ tw, int N, int log2n) {
  const int j = threadIdx.x, T4 = N >> 2;
  int Ns = 1;
  for (int s = 0; s + 2 <= log2n; s += 2) {
    const int k = j & (Ns - 1);
    const int tstep = N / (Ns * 4);   // twiddle index of e^{-2 pi i k / (4 Ns)} = k * tstep
    float2 v0 = a[j], v1 = a[j + T4], v2 = a[j + 2 * T4], v3 = a[j + 3 * T4];
    if (k) {
      v1 = cmul(v1, tw[k * tstep]);
      v2 = cmul(v2, tw[2 * k * tstep]);
      v3 = cmul(v3, tw[3 * k * tstep]);
    }
    const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
    const int j0 = ((j - k) << 2) + k;
    b[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
    b[j0 + Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);       // d02 - i d13
    b[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
    b[j0 + 3 * Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
    __syncthreads();
    float2* t = a; a = b; b = t;
    Ns <<= 2;
  }
  if (log2n & 1) {   // last stage radix 2: N / 2 butterflies, two per thread
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int jj = j + h * T4, k = jj & (Ns - 1);
      float2 v0 = a[jj], v1 = a[jj + (N >> 1)];
      if (k) v1 = cmul(v1, tw[k * (N / (Ns * 2))]);
      const int j0 = ((jj - k) << 1) + k;
      b[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
      b[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
    }
    __syncthreads();
    float2* t = a; a = b; b = t;
  }
  return a;
}

// np.pad(y, N / 2, mode='reflect') at padded position m (y has L >= 2 samples)
__device__ __forceinline__ float gl_reflect(const float* __restrict__ y, int L, int m, int halfN) {
  int i = m - halfN;
  const int period = 2 * (L - 1);
  i %= period;
  if (i < 0) i += period;
  if (i >= L) i = period - i;
  return __ldg(y + i);
}

__global__ void gl_magnitude_kernel(const float* __restrict__ spec, float* __restrict__ S, long long n, float max_abs, float ref_db, float power) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(spec + i);
    // Audio._denormalize / _symmetric_denormalize (Audio.py:96-100), min_level_db = -100
    const float db = max_abs > 0.f ? (fminf(fmaxf(v, -max_abs), max_abs) + max_abs) / (2.f * max_abs) * 100.f - 100.f
                                   : fminf(fmaxf(v, 0.f), 1.f) * 100.f - 100.f;
    // Audio._db_to_amp (:90-91) and "** power" (:26) as one exponential: (10^(x / 20))^p = 10^(x p / 20)
    S[i] = exp10f((db + ref_db) * (0.05f * power));
  }
}

constexpr unsigned int STREAM_GRIFFIN_LIM = 3;   // Philox stream of the initial phases (common.cuh: 0..2 are the decoder's)

inline size_t gl_frames_smem(int N) { return ((size_t)3 * N + 2) * sizeof(float2); }

__global__ void gl_frames_kernel(const GlParams p) {
  extern __shared__ __align__(16) unsigned char gl_raw[];
  const int tid = threadIdx.x, nt = blockDim.x, N = p.N, half = N >> 1;
  float2* a = reinterpret_cast<float2*>(gl_raw);   // FFT ping
  float2* b = a + N;                               // FFT pong
  float2* sp0 = b + N;                             // unit phases of frame t0, bins 0 .. N/2
  float2* sp1 = sp0 + half + 1;                    // ... of frame t0 + 1
  const int bi = blockIdx.y, t0 = 2 * blockIdx.x, t1 = t0 + 1;
  const int Tb = p.lengths ? min(max(__ldg(p.lengths + bi), 0), p.T) : p.T;
  if (t0 >= Tb || Tb < 2) return;
  const bool has1 = t1 < Tb;
  const float* S0 = p.S + ((size_t)bi * p.T + t0) * p.F;
  const float* S1 = S0 + p.F;
  if (p.init) {
    const float* u0 = p.uniform ? p.uniform + ((size_t)bi * p.T + t0) * p.F : nullptr;
    for (int k = tid; k <= half; k += nt) {
      float ua, ub;
      if (u0) {
        ua = __ldg(u0 + k);
        ub = has1 ? __ldg(u0 + p.F + k) : 0.f;
      } else {
        ua = (float)(philox_word(p.seed, STREAM_GRIFFIN_LIM, (unsigned)t0, (unsigned)(p.row_offset + bi), (unsigned)k) >> 8) * 5.9604644775390625e-08f;
        ub = (float)(philox_word(p.seed, STREAM_GRIFFIN_LIM, (unsigned)t1, (unsigned)(p.row_offset + bi), (unsigned)k) >> 8) * 5.9604644775390625e-08f;
      }
      float sn, cs;
      sincospif(2.f * ua, &sn, &cs);
      sp0[k] = make_float2(cs, sn);
      sincospif(2.f * ub, &sn, &cs);
      sp1[k] = make_float2(cs, sn);
    }
  } else {
    const int L = p.hop * (Tb - 1);
    const float* y = p.y + (size_t)bi * p.Lmax;
    // frames that lie inside the signal (all but the first and last n_fft / (2 hop) of an utterance) need no reflection
    const bool inside = t0 * p.hop >= half && (has1 ? t1 : t0) * p.hop + N - half <= L;
    if (inside) {
      const float* ya_p = y + t0 * p.hop - half;
      for (int i = tid; i < N; i += nt) {
        const float w = __ldg(p.window + i);
        a[i] = make_float2(w * __ldg(ya_p + i), has1 ? w * __ldg(ya_p + p.hop + i) : 0.f);
      }
    } else {
      for (int i = tid; i < N; i += nt) {
        const float w = __ldg(p.window + i);
        const float ya = gl_reflect(y, L, t0 * p.hop + i, half);
        const float yb = has1 ? gl_reflect(y, L, t1 * p.hop + i, half) : 0.f;
        a[i] = make_float2(w * ya, w * yb);
      }
    }
    __syncthreads();
    const float2* X = gl_fft(a, b, p.tw, N, p.log2n);
    for (int k = tid; k <= half; k += nt) {
      const float2 z = X[k], zc = X[(N - k) & (N - 1)];
      // Xa = (Z[k] + conj(Z[N-k])) / 2,  Xb = (Z[k] - conj(Z[N-k])) / (2 i)
      const float2 xa = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
      const float2 xb = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
      // exp(1j * np.angle(X)) = X / |X|, angle(0) = 0 (magnitudes below 1e-19 square to zero and count as 0)
      const float ma = xa.x * xa.x + xa.y * xa.y, mb = xb.x * xb.x + xb.y * xb.y;
      const float ia = rsqrtf(ma), ib = rsqrtf(mb);
      sp0[k] = ma > 0.f ? make_float2(xa.x * ia, xa.y * ia) : make_float2(1.f, 0.f);
      sp1[k] = mb > 0.f ? make_float2(xb.x * ib, xb.y * ib) : make_float2(1.f, 0.f);
    }
  }
  __syncthreads();
  // conj(W), W = Ya + i Yb with the Hermitian extension of both (irfft ignores the imaginary parts of bins 0 and N/2)
  for (int k = tid; k <= half; k += nt) {
    const float s0 = __ldg(S0 + k), s1 = has1 ? __ldg(S1 + k) : 0.f;
    const float2 ya = make_float2(s0 * sp0[k].x, s0 * sp0[k].y);
    const float2 yb = make_float2(s1 * sp1[k].x, s1 * sp1[k].y);
    if (k == 0 || k == half) {
      a[k] = make_float2(ya.x, -yb.x);
    } else {
      a[k] = make_float2(ya.x - yb.y, -(ya.y + yb.x));
      a[N - k] = make_float2(ya.x + yb.y, -(yb.x - ya.y));
    }
  }
  __syncthreads();
  const float2* z = gl_fft(a, b, p.tw, N, p.log2n);   // = N * conj(ifft(W))
  const float inv = 1.f / (float)N;
  float* f0 = p.frames + ((size_t)bi * p.T + t0) * N;
  for (int i = tid; i < N; i += nt) {
    const float w = __ldg(p.window + i) * inv;
    f0[i] = w * z[i].x;
    if (has1) f0[N + i] = -w * z[i].y;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// n_fft = 1024 (the reference's Frame_Length): the same frame-pair scheme with the FFT as THREE Stockham stages of radix 16, 16, 4
// (64 threads, 16 points per thread in registers) instead of five radix-4 stages over 256 threads: the radix-4 kernel is bound by
// its instruction count (profiles/r2_ncu_gl_frames_summary.txt), and this form executes less than half of them.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gl_dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
  a0 = make_float2(s02.x + s13.x, s02.y + s13.y);
  a1 = make_float2(d02.x + d13.y, d02.y - d13.x);   // d02 - i d13
  a2 = make_float2(s02.x - s13.x, s02.y - s13.y);
  a3 = make_float2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
}
// 16-point forward DFT of v[n] (n = 4 n1 + n2) in registers; on return v[4 k1 + k2] = X[k1 + 4 k2]
__device__ __forceinline__ void gl_dft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) gl_dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // over n1 -> v[4 k1 + n2]
  // twiddles W16^(n2 k1), W16 = exp(-2 pi i / 16)
  v[5] = cmul(v[5], make_float2(C1, -S1));    // k1 = 1, n2 = 1: W^1
  v[6] = cmul(v[6], make_float2(C2, -C2));    //         n2 = 2: W^2
  v[7] = cmul(v[7], make_float2(S1, -C1));    //         n2 = 3: W^3
  v[9] = cmul(v[9], make_float2(C2, -C2));    // k1 = 2, n2 = 1: W^2
  v[10] = make_float2(v[10].y, -v[10].x);     //         n2 = 2: W^4 = -i
  v[11] = cmul(v[11], make_float2(-C2, -C2)); //         n2 = 3: W^6
  v[13] = cmul(v[13], make_float2(S1, -C1));  // k1 = 3, n2 = 1: W^3
  v[14] = cmul(v[14], make_float2(-C2, -C2)); //         n2 = 2: W^6
  v[15] = cmul(v[15], make_float2(-C1, S1));  //         n2 = 3: W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) gl_dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // over n2 -> v[4 k1 + k2]
}
// shared-memory index of FFT element i: one pad element per 16 (the radix-16 stages store with a stride of 16 elements between
// threads: unpadded that is a 32-way bank conflict)
__device__ __forceinline__ int glp(int i) { return i + (i >> 4); }
constexpr int GL1K_PADDED = 1024 + 64;
// forward FFT of 1024 points, 64 threads: x -> y -> x -> y, result in y (both buffers indexed through glp)
__device__ __forceinline__ void gl_fft1024(float2* x, float2* y, const float2* __restrict__ tw) {
  const int j = threadIdx.x;
#pragma unroll 1
  for (int st = 0; st < 2; ++st) {   // radix 16, Ns = 1 then 16
    const int Ns = st ? 16 : 1, k = j & (Ns - 1);
    const float2* in = st ? y : x;
    float2* out = st ? x : y;
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = in[glp(j + 64 * r)];
    if (k) {
#pragma unroll
      for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], __ldg(tw + 4 * r * k));   // exp(-2 pi i r k / 256)
    }
    gl_dft16(v);
    const int o = ((j - k) << 4) + k;
#pragma unroll
    for (int idx = 0; idx < 16; ++idx) out[glp(o + ((idx >> 2) + 4 * (idx & 3)) * Ns)] = v[idx];
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {   // radix 4, Ns = 256: butterflies j' = j + 64 m
    const int jj = j + 64 * m;
    float2 v0 = x[glp(jj)], v1 = x[glp(jj + 256)], v2 = x[glp(jj + 512)], v3 = x[glp(jj + 768)];
    if (jj) {
      v1 = cmul(v1, __ldg(tw + jj));
      v2 = cmul(v2, __ldg(tw + 2 * jj));
      v3 = cmul(v3, __ldg(tw + 3 * jj));
    }
    gl_dft4(v0, v1, v2, v3);
    y[glp(jj)] = v0; y[glp(jj + 256)] = v1; y[glp(jj + 512)] = v2; y[glp(jj + 768)] = v3;
  }
  __syncthreads();
}

constexpr int GL1K_THREADS = 64;

// shared memory: two padded FFT buffers (17 KB per frame pair: 13 pairs per SM).  a: windowed signal -> (FFT) -> b = X; the unit
// phases of both frames go back into a ([0, 513) and [513, 1026)), conj(W) into b, (FFT) -> a = N conj(ifft(W)).
__global__ void __launch_bounds__(GL1K_THREADS) gl_frames1024_kernel(const GlParams p) {
  __shared__ __align__(16) float2 a[GL1K_PADDED], b[GL1K_PADDED];
  constexpr int N = 1024, half = 512;
  float2* sp0 = a;
  float2* sp1 = a + half + 1;
  const int tid = threadIdx.x, nt = GL1K_THREADS;
  const int bi = blockIdx.y, t0 = 2 * blockIdx.x, t1 = t0 + 1;
  const int Tb = p.lengths ? min(max(__ldg(p.lengths + bi), 0), p.T) : p.T;
  if (t0 >= Tb || Tb < 2) return;
  const bool has1 = t1 < Tb;
  const float* S0 = p.S + ((size_t)bi * p.T + t0) * p.F;
  const float* S1 = S0 + p.F;
  if (p.init) {
    const float* u0 = p.uniform ? p.uniform + ((size_t)bi * p.T + t0) * p.F : nullptr;
    for (int k = tid; k <= half; k += nt) {
      float ua, ub;
      if (u0) {
        ua = __ldg(u0 + k);
        ub = has1 ? __ldg(u0 + p.F + k) : 0.f;
      } else {
        ua = (float)(philox_word(p.seed, STREAM_GRIFFIN_LIM, (unsigned)t0, (unsigned)(p.row_offset + bi), (unsigned)k) >> 8) * 5.9604644775390625e-08f;
        ub = (float)(philox_word(p.seed, STREAM_GRIFFIN_LIM, (unsigned)t1, (unsigned)(p.row_offset + bi), (unsigned)k) >> 8) * 5.9604644775390625e-08f;
      }
      float sn, cs;
      sincospif(2.f * ua, &sn, &cs);
      sp0[k] = make_float2(cs, sn);
      sincospif(2.f * ub, &sn, &cs);
      sp1[k] = make_float2(cs, sn);
    }
  } else {
    const int L = p.hop * (Tb - 1);
    const float* y = p.y + (size_t)bi * p.Lmax;
    const bool inside = t0 * p.hop >= half && (has1 ? t1 : t0) * p.hop + N - half <= L;
    if (inside) {
      const float* ya_p = y + t0 * p.hop - half;
      for (int i = tid; i < N; i += nt) {
        const float w = __ldg(p.window + i);
        a[glp(i)] = make_float2(w * __ldg(ya_p + i), has1 ? w * __ldg(ya_p + p.hop + i) : 0.f);
      }
    } else {
      for (int i = tid; i < N; i += nt) {
        const float w = __ldg(p.window + i);
        const float ya = gl_reflect(y, L, t0 * p.hop + i, half);
        const float yb = has1 ? gl_reflect(y, L, t1 * p.hop + i, half) : 0.f;
        a[glp(i)] = make_float2(w * ya, w * yb);
      }
    }
    __syncthreads();
    gl_fft1024(a, b, p.tw);   // X in b; a is free
    for (int k = tid; k <= half; k += nt) {
      const float2 z = b[glp(k)], zc = b[glp((N - k) & (N - 1))];
      const float2 xa = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
      const float2 xb = make_float2(0.5f * (z.y + zc.y), -0.5f * (z.x - zc.x));
      const float ma = xa.x * xa.x + xa.y * xa.y, mb = xb.x * xb.x + xb.y * xb.y;
      const float ia = rsqrtf(ma), ib = rsqrtf(mb);
      sp0[k] = ma > 0.f ? make_float2(xa.x * ia, xa.y * ia) : make_float2(1.f, 0.f);
      sp1[k] = mb > 0.f ? make_float2(xb.x * ib, xb.y * ib) : make_float2(1.f, 0.f);
    }
  }
  __syncthreads();   // phases complete (and every read of X done) before b is overwritten
  for (int k = tid; k <= half; k += nt) {
    const float s0 = __ldg(S0 + k), s1 = has1 ? __ldg(S1 + k) : 0.f;
    const float2 ya = make_float2(s0 * sp0[k].x, s0 * sp0[k].y);
    const float2 yb = make_float2(s1 * sp1[k].x, s1 * sp1[k].y);
    if (k == 0 || k == half) {
      b[glp(k)] = make_float2(ya.x, -yb.x);
    } else {
      b[glp(k)] = make_float2(ya.x - yb.y, -(ya.y + yb.x));
      b[glp(N - k)] = make_float2(ya.x + yb.y, -(yb.x - ya.y));
    }
  }
  __syncthreads();
  gl_fft1024(b, a, p.tw);   // a = N * conj(ifft(W))
  const float inv = 1.f / (float)N;
  float* f0 = p.frames + ((size_t)bi * p.T + t0) * N;
  for (int i = tid; i < N; i += nt) {
    const float w = __ldg(p.window + i) * inv;
    f0[i] = w * a[glp(i)].x;
    if (has1) f0[N + i] = -w * a[glp(i)].y;
  }
}

// VEC = 4: four consecutive samples per thread (hop, n_fft and the sample index are multiples of 4, so the four samples lie in the
// same frames at consecutive offsets: 16 B loads); VEC = 1: any hop
template <int VEC>
__global__ void gl_overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ window, const int* __restrict__ lengths,
                                      float* __restrict__ y, int T, int N, int hop, int Lmax) {
  const int bi = blockIdx.y;
  const int Tb = lengths ? min(max(__ldg(lengths + bi), 0), T) : T;
  const int L = Tb > 0 ? hop * (Tb - 1) : 0;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < Lmax; i += gridDim.x * blockDim.x * VEC) {
    float v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = 0.f;
    if (i < L) {   // L is a multiple of hop: with VEC = 4 the whole group is inside or outside
      const int n = i + (N >> 1);
      int tlo = (n - N + hop) / hop;   // first frame that covers sample n: ceil((n - N + 1) / hop)
      if (n - N + 1 <= 0) tlo = 0;
      const int thi = min(Tb - 1, n / hop);
      float acc[VEC], wss[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] = wss[e] = 0.f;
      for (int t = tlo; t <= thi; ++t) {
        const int o = n - t * hop;
        if constexpr (VEC == 4) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(window + o));
          const float4 f = __ldg(reinterpret_cast<const float4*>(frames + ((size_t)bi * T + t) * N + o));
          acc[0] += f.x; acc[1] += f.y; acc[2] += f.z; acc[3] += f.w;
          wss[0] = fmaf(w.x, w.x, wss[0]); wss[1] = fmaf(w.y, w.y, wss[1]); wss[2] = fmaf(w.z, w.z, wss[2]); wss[3] = fmaf(w.w, w.w, wss[3]);
        } else {
          const float w = __ldg(window + o);
          acc[0] += __ldg(frames + ((size_t)bi * T + t) * N + o);
          wss[0] = fmaf(w, w, wss[0]);
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = wss[e] > 1.17549435e-38f ? acc[e] / wss[e] : acc[e];   // librosa.util.tiny of float32
    }
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(y + (size_t)bi * Lmax + i) = make_float4(v[0], v[1], v[2], v[3]);
    else y[(size_t)bi * Lmax + i] = v[0];
  }
}

// out[n] = y[n] + c * out[n-1]: thread i of the CTA owns 8 consecutive samples of a 2048-sample chunk; its affine map
// carry -> c^8 carry + local_last composes over the block by a scan, the chunk's last value carries into the next chunk.
constexpr int GLD_THREADS = 256, GLD_PER = 8;
__global__ void __launch_bounds__(GLD_THREADS) gl_deemphasis_kernel(const float* __restrict__ y, const int* __restrict__ lengths, float* __restrict__ out,
                                                                    int T, int hop, int Lmax, float c) {
  __shared__ float wP[GLD_THREADS / 32], wL[GLD_THREADS / 32];
  __shared__ float carry_s;
  const int bi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int Tb = lengths ? min(max(__ldg(lengths + bi), 0), T) : T;
  const int L = Tb > 0 ? hop * (Tb - 1) : 0;
  float cp[GLD_PER + 1];
  cp[0] = 1.f;
#pragma unroll
  for (int e = 1; e <= GLD_PER; ++e) cp[e] = cp[e - 1] * c;
  if (tid == 0) carry_s = 0.f;
  __syncthreads();
  for (int base = 0; base < Lmax; base += GLD_THREADS * GLD_PER) {
    const int i0 = base + tid * GLD_PER;
    float v[GLD_PER];
    float run = 0.f;
#pragma unroll
    for (int e = 0; e < GLD_PER; ++e) {
      const int i = i0 + e;
      const float x = i < L ? __ldg(y + (size_t)bi * Lmax + i) : 0.f;
      run = fmaf(c, run, x);
      v[e] = run;
    }
    // inclusive scan of (P, Lc): carry_out = P * carry_in + Lc
    float P = cp[GLD_PER], Lc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float Pp = __shfl_up_sync(0xffffffffu, P, d), Lp = __shfl_up_sync(0xffffffffu, Lc, d);
      if (lane >= d) {
        Lc = fmaf(P, Lp, Lc);
        P *= Pp;
      }
    }
    if (lane == 31) {
      wP[wid] = P;
      wL[wid] = Lc;
    }
    __syncthreads();
    float cin = carry_s;   // carry into this warp
    for (int w = 0; w < wid; ++w) cin = fmaf(wP[w], cin, wL[w]);
    // carry into this thread = exclusive prefix of the warp applied to cin
    const float Pe = __shfl_up_sync(0xffffffffu, P, 1), Le = __shfl_up_sync(0xffffffffu, Lc, 1);
    const float mine = lane ? fmaf(Pe, cin, Le) : cin;
#pragma unroll
    for (int e = 0; e < GLD_PER; ++e) {
      const int i = i0 + e;
      if (i < Lmax) out[(size_t)bi * Lmax + i] = i < L ? fmaf(cp[e + 1], mine, v[e]) : 0.f;
    }
    __syncthreads();
    if (tid == GLD_THREADS - 1) carry_s = fmaf(P, cin, Lc);
    __syncthreads();
  }
}

}  // namespace gstk
