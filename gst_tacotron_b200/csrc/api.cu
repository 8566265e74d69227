// libgsttaco.so - C ABI (include/gstk.h) over the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/gstk.h"
#include "common.cuh"
#include "decoder_bf16.cuh"
#include "decoder_bf16_v2.cuh"
#include "decoder_bf16_sb.cuh"
#include "decoder_fp32.cuh"
#include "gst.cuh"
#include "gst_tc.cuh"
#include "postnet.cuh"
#include "postnet_tc.cuh"
#include "encoder.cuh"
#include "vocoder.cuh"
#include "griffin_lim.cuh"
#include "umma.cuh"

using namespace gstk;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

enum Slot {
  SL_ENC = 0, SL_ENC_TEXT, SL_GST_IN, SL_TEACHER, SL_KEEP0, SL_KEEP1, SL_NOISE, SL_INIT_MEL, SL_INIT_ALIGN,
  SL_INIT_CUM, SL_INIT_STATES, SL_OUT_MEL, SL_OUT_STOP, SL_OUT_ALIGN, SL_OUT_STATES, SL_OUT_CUM, SL_OUT_CTX, SL_LASTMEL,
  SL_VPROJ, SL_GBIAS, SL_XIN, SL_H1, SL_H2, SL_C1, SL_C2, SL_ALIGN, SL_CUM,
  SL_MELS, SL_LENGTHS, SL_ACT0, SL_ACT1, SL_XS, SL_OUT_GST, SL_OUT_REF, SL_OUT_ATT,
  SL_MHA0, SL_MHA1, SL_MHA2, SL_MHA3, SL_MHA4, SL_MHA5, SL_MHA6, SL_MHA7, SL_MHA_OUT, SL_MHA_ATT,
  SL_CAT0, SL_CAT1, SL_CAT_OUT, SL_AT0, SL_AT1, SL_AT2, SL_AT3, SL_AT4, SL_AT5, SL_AT6, SL_AT7, SL_AT8, SL_AT9, SL_AT10,
  SL_AT11, SL_AT12, SL_AT_Q, SL_AT_K, SL_AT_V, SL_AT_CTX, SL_AT_AL, SL_STOP_IDX, SL_STOP_STATE, SL_GST_BLK0, SL_GST_BLK1, SL_BF16_A, SL_BF16_B, SL_BF16_C, SL_BF16_D, SL_BF16_E, SL_BF16_F,
  SL_POST_IN, SL_POST_OUT, SL_POST_A, SL_POST_B, SL_ENC_TOK, SL_ENC_OUT, SL_ENC_XS, SL_ENC_H,
  SL_VOC_IN, SL_VOC_OUT, SL_VOC_RNN, SL_VOC_RNN16, SL_VPROJ_A16, SL_PRE_IN, SL_PRE_K0, SL_PRE_K1, SL_PRE_MID, SL_PRE_OUT, SL_GL_SPEC, SL_GL_LEN, SL_GL_UNI, SL_GL_S, SL_GL_FRAMES, SL_GL_Y, SL_GL_OUT,
  SL_COUNT
};

struct PendingCopy {
  void* dst;
  const void* src;
  size_t bytes;
};

}  // namespace

struct GstkHandle {
  GstkConfig cfg;
  std::string err;
  int num_sms = 0;
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, DevBuf> dev_w;
  std::map<std::string, DevBuf> derived;
  bool dec_ready = false, gst_ready = false;
  std::string post_key, enc_key, voc_key;  // layer description the folded Postnet weights were prepared for
  DevBuf slots[SL_COUNT];
  GridBarrier* gb = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  bool ev_valid = false;
  cudaStream_t ev_stream = nullptr;
  int64_t launches = 0;
  std::vector<PendingCopy> pending;
  Bf16State bf16;
  V2State v2;   // dataflow variant of the bf16 decoder (free-running fast path)
  SbState sb;   // small-batch latency kernel (batch <= 16, free-running fast path)
  // time-chunked decode with overlapped device->host copies (host output buffers only)
  cudaStream_t st_copy = nullptr;
  cudaEvent_t ev_chunk = nullptr, ev_copied = nullptr;
  // host INPUT buffers are copied on a stream of their own: a call's host->device copies do not queue behind kernels that are
  // already on the caller's stream (an earlier call of the same step, or - with two handles on one GPU - whatever the device is
  // busy with), only behind the previous use of the same staging slot (slot_ev, recorded when the call that used it ends)
  cudaStream_t st_in = nullptr;
  cudaEvent_t ev_in = nullptr;
  cudaEvent_t slot_ev[SL_COUNT] = {};
  std::vector<int> staged;
};

namespace {

int fail(GstkHandle* h, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

// Scoped current-device switch: every entry point runs on the handle's device and restores the caller's current device on
// exit (a process that drives several GPUs, or PyTorch work on another device, is not disturbed).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    else if (err == cudaSuccess) prev = -1;   // nothing to restore
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define DEVICE_GUARD(h_, dev_)                                                                     \
  DeviceGuard dev_guard_(dev_);                                                                    \
  if (dev_guard_.err != cudaSuccess)                                                               \
    return fail(h_, GSTK_ECUDA, "cudaSetDevice(%d) failed: %s", (dev_), cudaGetErrorString(dev_guard_.err))

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, GSTK_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

bool is_device_ptr(const void* p, int* device = nullptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (device) *device = a.device;
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// borrowed device tensors must live on the handle's GPU (managed memory migrates, so it is accepted from anywhere)
// GSTK_TRACE=1: host-side timeline of the API calls on stderr (microseconds since the first event, handle, label)
void trace(GstkHandle* h, const char* what) {
  static const bool on = getenv("GSTK_TRACE") && atoi(getenv("GSTK_TRACE"));
  if (!on) return;
  static const auto t0 = std::chrono::steady_clock::now();
  const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stderr, "[gstk %p] %10.0f us  %s\n", (void*)h, us, what);
}

int check_borrowed(GstkHandle* h, const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return GSTK_OK;
  }
  if (a.type == cudaMemoryTypeDevice && a.device != h->cfg.device)
    return fail(h, GSTK_EINVAL, "device tensor lives on GPU %d but the handle was created for GPU %d", a.device, h->cfg.device);
  return GSTK_OK;
}

int slot_reserve(GstkHandle* h, int slot, size_t bytes, void** out) {
  DevBuf& b = h->slots[slot];
  if (b.bytes < bytes) {
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
    size_t want = std::max(bytes, (size_t)256);
    CK(cudaMalloc(&b.p, want));
    b.bytes = want;
  }
  *out = b.p;
  return GSTK_OK;
}

// input tensor: device pointers are borrowed, host pointers are staged (H2D on `st`)
int stage_in(GstkHandle* h, int slot, const void* src, size_t bytes, cudaStream_t st, const void** out) {
  if (!src) {
    *out = nullptr;
    return GSTK_OK;
  }
  if (is_device_ptr(src)) {
    *out = src;
    return check_borrowed(h, src);
  }
  void* d;
  int rc = slot_reserve(h, slot, bytes, &d);
  if (rc) return rc;
  if (!h->st_in) {
    CK(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  }
  if (h->slot_ev[slot]) CK(cudaStreamWaitEvent(h->st_in, h->slot_ev[slot], 0));   // the slot's previous readers are done
  CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, h->st_in));
  CK(cudaEventRecord(h->ev_in, h->st_in));
  CK(cudaStreamWaitEvent(st, h->ev_in, 0));                                        // this call's kernels see the copy
  h->staged.push_back(slot);
  *out = d;
  return GSTK_OK;
}

// output tensor: host pointers get a device staging buffer + a pending D2H copy
int stage_out(GstkHandle* h, int slot, void* dst, size_t bytes, void** out) {
  if (!dst) {
    *out = nullptr;
    return GSTK_OK;
  }
  if (is_device_ptr(dst)) {
    *out = dst;
    return check_borrowed(h, dst);
  }
  void* d;
  int rc = slot_reserve(h, slot, bytes, &d);
  if (rc) return rc;
  h->pending.push_back({dst, d, bytes});
  *out = d;
  return GSTK_OK;
}

// GridBarrier::error is sticky: launches never clear it (reset_barrier leaves it alone), so a time-out in any chunk of a call -
// or in an earlier call whose outputs all stayed on the device - is still there when the flag is finally read here.
int check_barrier_error(GstkHandle* h) {
  unsigned int e = 0;
  CK(cudaMemcpy(&e, &h->gb->error, sizeof(e), cudaMemcpyDeviceToHost));
  if (e) {
    CK(cudaMemset(&h->gb->error, 0, sizeof(e)));
    return fail(h, GSTK_ETIMEOUT, "persistent decoder kernel: grid barrier timed out");
  }
  return GSTK_OK;
}
int reset_barrier(GstkHandle* h, cudaStream_t st) {
  const size_t e0 = offsetof(GridBarrier, error), e1 = offsetof(GridBarrier, pad2);
  CK(cudaMemsetAsync(h->gb, 0, e0, st));
  CK(cudaMemsetAsync((char*)h->gb + e1, 0, sizeof(GridBarrier) - e1, st));
  return GSTK_OK;
}

int flush_pending(GstkHandle* h, cudaStream_t st, bool check_barrier) {
  trace(h, "flush: enter");
  for (int slot : h->staged) {   // end of the call: everything that reads this call's staged inputs is on `st` by now
    if (!h->slot_ev[slot]) CK(cudaEventCreateWithFlags(&h->slot_ev[slot], cudaEventDisableTiming));
    CK(cudaEventRecord(h->slot_ev[slot], st));
  }
  h->staged.clear();
  if (h->pending.empty()) return GSTK_OK;
  for (auto& c : h->pending)
    if (c.bytes) CK(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, st));
  h->pending.clear();
  CK(cudaStreamSynchronize(st));
  trace(h, "flush: synchronized");
  if (getenv("GSTK_TRACE") && atoi(getenv("GSTK_TRACE")) && h->ev_valid) {   // device-side position of the decoder launches
    static cudaEvent_t ref = nullptr;
    if (!ref) {
      cudaEventCreate(&ref);
      cudaEventRecord(ref, st);
      cudaEventSynchronize(ref);
    }
    float a0 = 0.f, a1 = 0.f;
    if (cudaEventElapsedTime(&a0, ref, h->ev0) == cudaSuccess && cudaEventElapsedTime(&a1, ref, h->ev1) == cudaSuccess)
      fprintf(stderr, "[gstk %p] device: first decoder launch at %.2f ms, last ends at %.2f ms (since the reference event)\n", (void*)h, a0, a1);
    cudaGetLastError();
  }
  if (check_barrier) return check_barrier_error(h);
  return GSTK_OK;
}

const std::vector<float>* hw(GstkHandle* h, const std::string& name) {
  auto it = h->host_w.find(name);
  return it == h->host_w.end() ? nullptr : &it->second;
}
const float* dw(GstkHandle* h, const std::string& name) {
  auto it = h->dev_w.find(name);
  return it == h->dev_w.end() ? nullptr : (const float*)it->second.p;
}

int upload_derived(GstkHandle* h, const std::string& name, const void* data, size_t bytes) {
  DevBuf& b = h->derived[name];
  if (b.bytes < bytes) {
    if (b.p) CK(cudaFree(b.p));
    CK(cudaMalloc(&b.p, bytes));
    b.bytes = bytes;
  }
  CK(cudaMemcpy(b.p, data, bytes, cudaMemcpyHostToDevice));
  return GSTK_OK;
}
const float* dd(GstkHandle* h, const std::string& name) { return (const float*)h->derived[name].p; }

const char* DEC = "Decoder/Decoder_Step";
const char* GSTP = "Style_Token_Layer";

// Re-pack [kernel ; recurrent_kernel] of one LSTMCell into [group][k4][32 cols][4] (see lstm_phase).
std::vector<float> pack_lstm(const std::vector<float>& kx, const std::vector<float>& kh, int Kx, int U) {
  const int K = Kx + U, K4 = K / 4, groups = U / LSTM_HU;
  std::vector<float> out((size_t)groups * K4 * 32 * 4);
  for (int g = 0; g < groups; ++g)
    for (int k = 0; k < K; ++k) {
      const float* row = k < Kx ? &kx[(size_t)k * 4 * U] : &kh[(size_t)(k - Kx) * 4 * U];
      for (int gate = 0; gate < 4; ++gate)
        for (int u = 0; u < LSTM_HU; ++u) {
          const int col = gate * LSTM_HU + u;
          out[(((size_t)g * K4 + k / 4) * 32 + col) * 4 + (k % 4)] = row[(size_t)gate * U + g * LSTM_HU + u];
        }
    }
  return out;
}

// rows x K row-major floats -> bf16 K-major SWIZZLE_128B image [kb][rows][64] (umma.cuh)
void pack_sw128(const float* src, int ld, int rows, int KB, __nv_bfloat16* dst) {
  for (int kb = 0; kb < KB; ++kb)
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < 64; ++k) {
        const size_t off = (size_t)kb * rows * 128 + sw128_offset_bytes(r, k);
        dst[off / 2] = __float2bfloat16(src[(size_t)r * ld + kb * 64 + k]);
      }
}

int need(GstkHandle* h, const std::string& name, size_t count) {
  auto* v = hw(h, name);
  if (!v) return fail(h, GSTK_ENOWEIGHTS, "variable %s has not been loaded", name.c_str());
  if (v->size() != count)
    return fail(h, GSTK_EINVAL, "variable %s has %zu elements, expected %zu", name.c_str(), v->size(), count);
  return GSTK_OK;
}

int upload_folded_conv(GstkHandle* h, const std::string& tag, const std::vector<float>& w, const std::vector<float>& sc,
                       const std::vector<float>& sh, size_t K, int co);

int prepare_decoder(GstkHandle* h) {
  if (h->dec_ready) return GSTK_OK;
  const GstkConfig& c = h->cfg;
  const std::string d = DEC;
  const int PD = c.mel_dim * c.step_reduction + 1;
  int rc;
#define NEED(n, cnt) if ((rc = need(h, d + n, (size_t)(cnt)))) return rc
  NEED("/Prenet/dense/kernel", c.mel_dim * c.prenet0);
  NEED("/Prenet/dense/bias", c.prenet0);
  NEED("/Prenet/dense_1/kernel", c.prenet0 * c.prenet1);
  NEED("/Prenet/dense_1/bias", c.prenet1);
  NEED("/Attention/Query/kernel", c.prenet1 * c.attention_size);
  NEED("/Attention/Query/bias", c.attention_size);
  NEED("/Attention/Value/kernel", c.enc_dim * c.attention_size);
  NEED("/Attention/Value/bias", c.attention_size);
  if (c.attention_type != GSTK_ATT_LSA) {
    NEED("/Attention/attention_v", c.attention_size);
    NEED("/Attention/attention_score_bias", 1);
  } else {
    NEED("/Attention/Alignment_Conv/kernel", c.lsa_kernel * c.lsa_filters);
    NEED("/Attention/Alignment_Conv/bias", c.lsa_filters);
    NEED("/Attention/Alignment_Dense/kernel", c.lsa_filters * c.attention_size);
    NEED("/Attention/Alignment_Dense/bias", c.attention_size);
    NEED("/Attention/bias", c.attention_size);
  }
  NEED("/RNN/cell_0/kernel", (c.prenet1 + c.attention_size) * 4 * c.lstm0);
  NEED("/RNN/cell_0/recurrent_kernel", c.lstm0 * 4 * c.lstm0);
  NEED("/RNN/cell_0/bias", 4 * c.lstm0);
  NEED("/RNN/cell_1/kernel", c.lstm0 * 4 * c.lstm1);
  NEED("/RNN/cell_1/recurrent_kernel", c.lstm1 * 4 * c.lstm1);
  NEED("/RNN/cell_1/bias", 4 * c.lstm1);
  NEED("/Projection/kernel", (c.lstm1 + c.attention_size) * PD);
  NEED("/Projection/bias", PD);
#undef NEED
  {
    auto pk = pack_lstm(*hw(h, d + "/RNN/cell_0/kernel"), *hw(h, d + "/RNN/cell_0/recurrent_kernel"),
                        c.prenet1 + c.attention_size, c.lstm0);
    if ((rc = upload_derived(h, "L1pk", pk.data(), pk.size() * 4))) return rc;
    pk = pack_lstm(*hw(h, d + "/RNN/cell_1/kernel"), *hw(h, d + "/RNN/cell_1/recurrent_kernel"), c.lstm0, c.lstm1);
    if ((rc = upload_derived(h, "L2pk", pk.data(), pk.size() * 4))) return rc;
  }
  if (c.precision == GSTK_PREC_BF16) {
    if ((rc = bf16_prepare(h->bf16, c, h->host_w, h->err))) return rc;
    // value projection as a k = 1 layer of the tcgen05 conv kernel: fp16 kernel [E][A] + its [A][E] transpose, shift = bias
    std::vector<float> one((size_t)c.attention_size, 1.f);
    if ((rc = upload_folded_conv(h, "decv", *hw(h, d + "/Attention/Value/kernel"), one, *hw(h, d + "/Attention/Value/bias"),
                                 (size_t)c.enc_dim, c.attention_size))) return rc;
  }
  h->dec_ready = true;
  return GSTK_OK;
}

int prepare_gst(GstkHandle* h) {
  if (h->gst_ready) return GSTK_OK;
  const GstkConfig& c = h->cfg;
  if (!c.gst_use) return fail(h, GSTK_ENOTIMPL, "GST is not used");
  const std::string g = GSTP, r = g + "/Reference_Encoder";
  int rc;
  int cin = 1, melw = c.mel_dim;
  for (int i = 0; i < c.ref_layers; ++i) {
    const std::string base = r + "/Conv2D_" + std::to_string(i);
    const int co = c.ref_filters[i];
    if ((rc = need(h, base + "/conv2d/kernel", (size_t)9 * cin * co))) return rc;
    for (const char* n : {"gamma", "beta", "moving_mean", "moving_variance"})
      if ((rc = need(h, base + "/batch_normalization/" + n, co))) return rc;
    const auto& ga = *hw(h, base + "/batch_normalization/gamma");
    const auto& be = *hw(h, base + "/batch_normalization/beta");
    const auto& mu = *hw(h, base + "/batch_normalization/moving_mean");
    const auto& va = *hw(h, base + "/batch_normalization/moving_variance");
    std::vector<float> sc(co), sh(co);
    for (int k = 0; k < co; ++k) {
      // BatchNormalization inference: gamma*(x-mean)/sqrt(var+1e-3)+beta
      const double s = (double)ga[k] / std::sqrt((double)va[k] + 1e-3);
      sc[k] = (float)s;
      sh[k] = (float)((double)be[k] - (double)mu[k] * s);
    }
    if ((rc = upload_derived(h, "conv_scale" + std::to_string(i), sc.data(), co * 4))) return rc;
    if ((rc = upload_derived(h, "conv_shift" + std::to_string(i), sh.data(), co * 4))) return rc;
    if (c.precision == GSTK_PREC_BF16) {
      // tensor-core mode (gst_tc.cuh): BatchNormalization scale folded into the kernel; layer 0 = [9][co] fp32 for the direct
      // kernel, layers >= 1 = the 2x2-block kernel [co][16 cin] fp16, K index = (block tap dh*2+dw, sub-pixel sh*2+sw, c),
      // value = w[2 dh + sh][2 dw + sw][c][n] where that tap exists (<= 2), zero otherwise
      const auto& w = *hw(h, base + "/conv2d/kernel");   // [3][3][cin][co]
      if (i == 0) {
        // mma.sync B fragments of gst_conv0_mma_kernel: [co/32 groups][4 n-tiles][32 lanes][2 x (fp16, fp16)], K = 9 taps (rows 9..15
        // zero); column j of n-tile n of group q = channel 32 q + 8 (j / 2) + 2 n + (j % 2)  (see the kernel: contiguous channels per lane)
        std::vector<__half> w0((size_t)(co / 32) * 4 * 32 * 4);
        auto wv = [&](int k, int ch) { return k < 9 ? (float)((double)w[(size_t)k * co + ch] * (double)sc[ch]) : 0.f; };
        for (int q = 0; q < co / 32; ++q)
          for (int n = 0; n < 4; ++n)
            for (int lane = 0; lane < 32; ++lane) {
              const int g = lane >> 2, t = lane & 3;          // B fragment: b0 = {B[2t][g], B[2t+1][g]}, b1 = {B[2t+8][g], B[2t+9][g]}
              const int ch = 32 * q + 8 * (g / 2) + 2 * n + (g % 2);
              __half* o = w0.data() + ((size_t)((q * 4 + n) * 32 + lane)) * 4;
              o[0] = __float2half_rn(wv(2 * t, ch));     o[1] = __float2half_rn(wv(2 * t + 1, ch));
              o[2] = __float2half_rn(wv(2 * t + 8, ch)); o[3] = __float2half_rn(wv(2 * t + 9, ch));
            }
        if ((rc = upload_derived(h, "gst_w0p", w0.data(), w0.size() * 2))) return rc;
      } else {
        const size_t K = (size_t)16 * cin;
        std::vector<__half> wt((size_t)co * K, __float2half_rn(0.f));
        for (int dh = 0; dh < 2; ++dh)
          for (int dw = 0; dw < 2; ++dw)
            for (int s2 = 0; s2 < 2; ++s2)
              for (int t2 = 0; t2 < 2; ++t2) {
                const int kh = 2 * dh + s2, kw = 2 * dw + t2;
                if (kh > 2 || kw > 2) continue;
                for (int ci = 0; ci < cin; ++ci)
                  for (int n = 0; n < co; ++n) {
                    const float v = (float)((double)w[((size_t)(kh * 3 + kw) * cin + ci) * co + n] * (double)sc[n]);
                    wt[(size_t)n * K + (size_t)((dh * 2 + dw) * 4 + (s2 * 2 + t2)) * cin + ci] = __float2half_rn(std::min(std::max(v, -65504.f), 65504.f));
                  }
              }
        if ((rc = upload_derived(h, "gst_wt" + std::to_string(i), wt.data(), wt.size() * 2))) return rc;
      }
    }
    cin = co;
    melw = (melw + 1) / 2;
  }
  const int gin = melw * cin, G = c.ref_gru, D = c.ref_dense, S = c.style_size;
  if ((rc = need(h, r + "/RNN/kernel", (size_t)gin * 3 * G))) return rc;
  if (c.precision == GSTK_PREC_BF16) {   // [3G][gin] fp16, K-major: the GRU input projection as one more tcgen05 GEMM
    const auto& wk = *hw(h, r + "/RNN/kernel");
    std::vector<__half> wt((size_t)3 * G * gin);
    for (int k = 0; k < gin; ++k)
      for (int n = 0; n < 3 * G; ++n) wt[(size_t)n * gin + k] = __float2half_rn(wk[(size_t)k * 3 * G + n]);
    if ((rc = upload_derived(h, "gst_rnn_wt", wt.data(), wt.size() * 2))) return rc;
    if (G == GM_G) {
      // recurrent kernel as mma.sync A fragments (gru_mma_dense_mha_kernel): [8 warps][3 gates][8 k-tiles][32 lanes][8 fp16],
      // tile row i = output gate * G + 16 w + i, tile column = k;  lane (g, t): rows g, g + 8, columns 2t, 2t+1, 2t+8, 2t+9
      const auto& U = *hw(h, r + "/RNN/recurrent_kernel");   // [G][3G]
      std::vector<__half> up((size_t)8 * 3 * 8 * 32 * 8);
      auto at = [&](int n, int k) { return __float2half_rn(U[(size_t)k * 3 * G + n]); };
      for (int w = 0; w < 8; ++w)
        for (int gate = 0; gate < 3; ++gate)
          for (int kt = 0; kt < 8; ++kt)
            for (int lane = 0; lane < 32; ++lane) {
              const int gg = lane >> 2, t = lane & 3, nb = gate * G + 16 * w, kb = 16 * kt;
              __half* o = up.data() + ((size_t)((w * 3 + gate) * 8 + kt) * 32 + lane) * 8;
              o[0] = at(nb + gg, kb + 2 * t);         o[1] = at(nb + gg, kb + 2 * t + 1);
              o[2] = at(nb + gg + 8, kb + 2 * t);     o[3] = at(nb + gg + 8, kb + 2 * t + 1);
              o[4] = at(nb + gg, kb + 2 * t + 8);     o[5] = at(nb + gg, kb + 2 * t + 9);
              o[6] = at(nb + gg + 8, kb + 2 * t + 8); o[7] = at(nb + gg + 8, kb + 2 * t + 9);
            }
      if ((rc = upload_derived(h, "gst_gru_up", up.data(), up.size() * 2))) return rc;
    }
  }
  if ((rc = need(h, r + "/RNN/recurrent_kernel", (size_t)G * 3 * G))) return rc;
  if ((rc = need(h, r + "/RNN/bias", (size_t)2 * 3 * G))) return rc;
  if ((rc = need(h, r + "/Dense/kernel", (size_t)G * D))) return rc;
  if ((rc = need(h, r + "/Dense/bias", D))) return rc;
  if ((rc = need(h, g + "/Attention/Query/kernel", (size_t)D * S))) return rc;
  if ((rc = need(h, g + "/Attention/Query/bias", S))) return rc;
  if ((rc = need(h, g + "/Attention/Value/kernel", (size_t)c.token_dim * S))) return rc;
  if ((rc = need(h, g + "/Attention/Value/bias", S))) return rc;
  if ((rc = need(h, g + "/Attention/Layer_Normalization/gamma", S))) return rc;
  if ((rc = need(h, g + "/Attention/Layer_Normalization/beta", S))) return rc;
  if ((rc = need(h, g + "/gst_tokens", (size_t)c.n_tokens * c.token_dim))) return rc;
  {
    // batch-invariant key/value rows: tanh(gst_tokens).Value + bias  (GST.py:100-103, Layers.py:175)
    const auto& tok = *hw(h, g + "/gst_tokens");
    const auto& Wv = *hw(h, g + "/Attention/Value/kernel");
    const auto& bv = *hw(h, g + "/Attention/Value/bias");
    std::vector<float> kv((size_t)c.n_tokens * S);
    for (int t = 0; t < c.n_tokens; ++t)
      for (int n = 0; n < S; ++n) {
        double a = bv[n];
        for (int k = 0; k < c.token_dim; ++k)
          a += (double)std::tanh((double)tok[(size_t)t * c.token_dim + k]) * (double)Wv[(size_t)k * S + n];
        kv[(size_t)t * S + n] = (float)a;
      }
    if ((rc = upload_derived(h, "tokkv", kv.data(), kv.size() * 4))) return rc;
  }
  h->gst_ready = true;
  return GSTK_OK;
}

int launch_sgemm(GstkHandle* h, const float* A, long long lda, const float* W, const float* bias,
                 const float* gbias, int rpg, float* C, int M, int N, int K, cudaStream_t st) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  sgemm_bias_kernel<<<grid, SG_THREADS, 0, st>>>(A, lda, W, bias, gbias, rpg, C, M, N, K);
  h->launches++;
  CK(cudaGetLastError());
  return GSTK_OK;
}

const char* POSTP = "Decoder/Postnet";

// Postnet variables (Taco2.py:130-147): BatchNormalization (inference form, eps 1e-3) folded into the bias-free Conv1D
// kernels as a per-output-channel scale; the shift stays for the epilogue.  Uploaded in the handle's precision.
int prepare_postnet(GstkHandle* h, const GstkPostnetArgs* a) {
  const GstkConfig& c = h->cfg;
  std::string key = std::to_string(c.precision) + ":";
  for (int i = 0; i < a->n_layers; ++i) key += std::to_string(a->filters[i]) + "x" + std::to_string(a->kernel[i]) + ",";
  if (h->post_key == key) return GSTK_OK;
  int rc, cin = c.mel_dim;
  for (int i = 0; i < a->n_layers; ++i) {
    const int co = a->filters[i], k = a->kernel[i];
    const std::string conv = std::string(POSTP) + "/conv1d_" + std::to_string(i) + "/kernel";
    const std::string bn = std::string(POSTP) + "/batch_normalization_" + std::to_string(i) + "/";
    if ((rc = need(h, conv, (size_t)k * cin * co))) return rc;
    for (const char* n : {"gamma", "beta", "moving_mean", "moving_variance"})
      if ((rc = need(h, bn + n, co))) return rc;
    const auto& w = *hw(h, conv);
    const auto& ga = *hw(h, bn + "gamma");
    const auto& be = *hw(h, bn + "beta");
    const auto& mu = *hw(h, bn + "moving_mean");
    const auto& va = *hw(h, bn + "moving_variance");
    std::vector<float> sc(co), sh(co), wf(w.size());
    for (int n = 0; n < co; ++n) {
      const double s = (double)ga[n] / std::sqrt((double)va[n] + 1e-3);
      sc[n] = (float)s;
      sh[n] = (float)((double)be[n] - (double)mu[n] * s);
    }
    for (size_t e = 0; e < w.size(); ++e) wf[e] = (float)((double)w[e] * (double)sc[e % co]);
    if ((rc = upload_derived(h, "post_shift" + std::to_string(i), sh.data(), (size_t)co * 4))) return rc;
    if (c.precision == GSTK_PREC_BF16) {
      std::vector<__half> wb(wf.size());   // fp16 operands, see postnet.cuh
      for (size_t e = 0; e < wf.size(); ++e) wb[e] = __float2half_rn(std::min(std::max(wf[e], -65504.f), 65504.f));
      if ((rc = upload_derived(h, "post_w" + std::to_string(i), wb.data(), wb.size() * 2))) return rc;
      // K-major copy [N][K] for the tcgen05 kernel (postnet_tc.cuh)
      const size_t K = (size_t)k * cin;
      std::vector<__half> wt(wb.size());
      for (size_t kk = 0; kk < K; ++kk)
        for (int n = 0; n < co; ++n) wt[(size_t)n * K + kk] = wb[kk * co + n];
      if ((rc = upload_derived(h, "post_wt" + std::to_string(i), wt.data(), wt.size() * 2))) return rc;
    } else {
      if ((rc = upload_derived(h, "post_w" + std::to_string(i), wf.data(), wf.size() * 4))) return rc;
    }
    cin = co;
  }
  h->post_key = key;
  return GSTK_OK;
}

// 2-D fp16 tensor map (row-major matrix, `inner` contiguous elements per row), SWIZZLE_128B boxes of 64 x box_rows
int encode_tmap_f16(GstkHandle* h, CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_bytes,
                    uint32_t box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    if (!sym || qres != cudaDriverEntryPointSuccess) return fail(h, GSTK_ECUDA, "cuTensorMapEncodeTiled is not available");
    fn = (EncodeFn)sym;
  }
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {row_bytes};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, GSTK_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return GSTK_OK;
}

// One Conv1D('same', stride 1) + folded BatchNormalization + activation layer on the flat padded activation matrix
// (postnet.cuh): tcgen05 kernel where the shapes allow it (tensor-core mode), mma.sync / FFMA kernels otherwise.
// `wt` = the [N][K] fp16 copy of the folded kernel for the tcgen05 path.
int launch_conv_layer(GstkHandle* h, const PostConvParams& p, bool bf16, int k, int padh, const void* wt, cudaStream_t st) {
  int rc;
  const int cin = p.C, co = p.N;
  const long long Mtotal = p.Mtotal;
  if (bf16) CK(cudaFuncSetAttribute(postnet_conv_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCB_SMEM));
  else CK(cudaFuncSetAttribute(postnet_conv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCF_SMEM));
    dim3 grid((co + PC_BN - 1) / PC_BN, (unsigned)((Mtotal + PC_BM - 1) / PC_BM));
    if (grid.y > 65535) return fail(h, GSTK_EINVAL, "conv layer: batch * frames too large for one launch");
    // tcgen05 path: input channels a multiple of 64 (one k-block = 64 channels of one tap), N a multiple of 16 that is
    // <= 256 or a multiple of 256.  GSTK_POSTNET_TC=0 keeps every layer on the mma.sync kernel (A/B measurements).
    static const bool tc_on = !(getenv("GSTK_POSTNET_TC") && atoi(getenv("GSTK_POSTNET_TC")) == 0);
    // accumulator width of a tile: the widest divisor of N that is a multiple of 16 and <= 256
    int tiles_n = 0;
    for (int tn = (co + 255) / 256; tn <= 64 && !tiles_n; ++tn)
      if (co % tn == 0 && (co / tn) % 16 == 0 && co / tn <= 256) tiles_n = tn;
    bool tc = bf16 && tc_on && cin % 8 == 0 && tiles_n > 0;
    CUtensorMap tmA, tmB;
    PostTcParams q;
    memset(&q, 0, sizeof(q));
    if (tc) {
      q.p = p;
      q.BN = co / tiles_n;
      q.tiles_n = tiles_n;
      q.tiles_m = (int)((Mtotal + PC_BM - 1) / PC_BM);
      q.cpb = cin % 64 == 0 ? cin / 64 : 0;
      q.KB = (k * cin + 63) / 64;
      // cpb > 0: plain [rows][C] matrix, one k-block = 64 channels of one tap; cpb == 0: overlapping-row view [rows][k*C]
      // with row stride C.  If the driver refuses the overlapping view the layer stays on the mma.sync kernel.
      rc = encode_tmap_f16(h, &tmA, p.X, q.cpb ? (uint64_t)cin : (uint64_t)k * cin, (uint64_t)Mtotal + padh, (uint64_t)cin * 2, PC_BM);
      if (rc && q.cpb == 0) tc = false;
      else if (rc) return rc;
    }
    if (tc) {
      if ((rc = encode_tmap_f16(h, &tmB, const_cast<void*>(wt), (uint64_t)k * cin, (uint64_t)co,
                                (uint64_t)k * cin * 2, (uint32_t)q.BN))) return rc;
      CK(cudaFuncSetAttribute(postnet_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
      const int ctas = std::min(h->num_sms, q.tiles_m * q.tiles_n);
      postnet_conv_tc_kernel<<<ctas, PT_THREADS, PT_SMEM, st>>>(tmA, tmB, q);
    } else if (bf16) postnet_conv_f16_kernel<<<grid, PC_THREADS, PCB_SMEM, st>>>(p);
    else postnet_conv_f32_kernel<<<grid, PC_THREADS, PCF_SMEM, st>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  return GSTK_OK;
}

const char* ENCP = "Encoder";

// fold BatchNormalization into a bias-free conv kernel [k*cin][co] and upload it (fp32, or fp16 + its [co][k*cin]
// transpose for the tcgen05 kernel) together with the shift vector
int upload_folded_conv(GstkHandle* h, const std::string& tag, const std::vector<float>& w, const std::vector<float>& sc,
                       const std::vector<float>& sh, size_t K, int co) {
  int rc;
  std::vector<float> wf(w.size());
  for (size_t e = 0; e < w.size(); ++e) wf[e] = (float)((double)w[e] * (double)sc[e % co]);
  if ((rc = upload_derived(h, tag + "_shift", sh.data(), (size_t)co * 4))) return rc;
  if (h->cfg.precision == GSTK_PREC_BF16) {
    std::vector<__half> wb(wf.size()), wt(wf.size());
    for (size_t e = 0; e < wf.size(); ++e) wb[e] = __float2half_rn(std::min(std::max(wf[e], -65504.f), 65504.f));
    for (size_t kk = 0; kk < K; ++kk)
      for (int n = 0; n < co; ++n) wt[(size_t)n * K + kk] = wb[kk * co + n];
    if ((rc = upload_derived(h, tag + "_w", wb.data(), wb.size() * 2))) return rc;
    return upload_derived(h, tag + "_wt", wt.data(), wt.size() * 2);
  }
  return upload_derived(h, tag + "_w", wf.data(), wf.size() * 4);
}

// Input kernels and biases of the forward and backward LSTM cells of a Bidirectional wrapper as ONE [cin][8u] projection (a
// k = 1 conv layer, tag_w / tag_wt / tag_shift).  Keras column n = gate * u + unit -> projection column dir * 4u + unit * 4 + gate
// (the recurrent kernels read the 4 gates of a unit with one 16 B load).
int upload_bilstm_proj(GstkHandle* h, const std::string& tag, const std::string& prefix, int cin, int u) {
  int rc;
  std::vector<float> wx((size_t)cin * 8 * u), bx((size_t)8 * u), one((size_t)8 * u, 1.f);
  int d = 0;
  for (const char* dn : {"forward_lstm", "backward_lstm"}) {
    const std::string base = prefix + dn + "/lstm_cell/";
    if ((rc = need(h, base + "kernel", (size_t)cin * 4 * u))) return rc;
    if ((rc = need(h, base + "recurrent_kernel", (size_t)u * 4 * u))) return rc;
    if ((rc = need(h, base + "bias", (size_t)4 * u))) return rc;
    const auto& kx = *hw(h, base + "kernel");
    const auto& b = *hw(h, base + "bias");
    for (int n = 0; n < 4 * u; ++n) {
      const size_t dst = (size_t)d * 4 * u + (size_t)(n % u) * 4 + n / u;
      for (int r = 0; r < cin; ++r) wx[(size_t)r * 8 * u + dst] = kx[(size_t)r * 4 * u + n];
      bx[dst] = b[n];
    }
    ++d;
  }
  return upload_folded_conv(h, tag, wx, one, bx, (size_t)cin, 8 * u);
}

// Encoder variables (Taco2.py:16-45, weights.encoder_spec): conv kernels with BatchNormalization folded in; the input kernels
// and biases of the forward and backward LSTM cells concatenated into one [cin][8u] projection (a k = 1 conv layer).
int prepare_encoder(GstkHandle* h, const GstkEncoderArgs* a) {
  const GstkConfig& c = h->cfg;
  std::string key = std::to_string(c.precision) + ":" + std::to_string(a->vocab) + ":" + std::to_string(a->embedding) + ":" +
                    std::to_string(a->rnn_size) + ":";
  for (int i = 0; i < a->n_layers; ++i) key += std::to_string(a->filters[i]) + "x" + std::to_string(a->kernel[i]) + ",";
  if (h->enc_key == key) return GSTK_OK;
  int rc, cin = a->embedding;
  const std::string e = ENCP;
  if ((rc = need(h, e + "/embedding/embeddings", (size_t)a->vocab * a->embedding))) return rc;
  for (int i = 0; i < a->n_layers; ++i) {
    const int co = a->filters[i], k = a->kernel[i];
    const std::string conv = e + "/conv1d_" + std::to_string(i) + "/kernel";
    const std::string bn = e + "/batch_normalization_" + std::to_string(i) + "/";
    if ((rc = need(h, conv, (size_t)k * cin * co))) return rc;
    for (const char* n : {"gamma", "beta", "moving_mean", "moving_variance"})
      if ((rc = need(h, bn + n, co))) return rc;
    const auto& ga = *hw(h, bn + "gamma");
    const auto& be = *hw(h, bn + "beta");
    const auto& mu = *hw(h, bn + "moving_mean");
    const auto& va = *hw(h, bn + "moving_variance");
    std::vector<float> sc(co), sh(co);
    for (int n = 0; n < co; ++n) {
      const double s = (double)ga[n] / std::sqrt((double)va[n] + 1e-3);
      sc[n] = (float)s;
      sh[n] = (float)((double)be[n] - (double)mu[n] * s);
    }
    if ((rc = upload_folded_conv(h, "enc" + std::to_string(i), *hw(h, conv), sc, sh, (size_t)k * cin, co))) return rc;
    cin = co;
  }
  if ((rc = upload_bilstm_proj(h, "encx", e + "/bidirectional/", cin, a->rnn_size))) return rc;
  h->enc_key = key;
  return GSTK_OK;
}

const char* VOCP = "Vocoder_Taco1";
constexpr int VOC_DENSE_ALIGN = 16;   // Dense(Spectrogram_Dim): N padded to the next multiple (513 -> 528 = 3 tiles of 176)

// Vocoder_Taco1 variables (weights.vocoder_spec): BatchNormalization folded into the conv kernels; the conv bank as ONE conv layer
// of kernel size bank_count whose column block i holds kernel i (size i + 1) at the taps its own 'same' padding selects, zeros
// elsewhere; Dense / Highwaynet kernels in fp32 for voc_highway_kernel; the LSTM input projection and the final Dense (N padded
// with zero columns) as k = 1 conv layers.
int prepare_vocoder(GstkHandle* h, const GstkVocoderArgs* a) {
  const GstkConfig& c = h->cfg;
  std::string key = std::to_string(c.precision) + ":" + std::to_string(a->bank_count) + ":" + std::to_string(a->bank_filters) + ":" +
                    std::to_string(a->highway_count) + ":" + std::to_string(a->highway_size) + ":" + std::to_string(a->rnn_size) + ":" +
                    std::to_string(a->spectrogram_dim) + ":";
  for (int i = 0; i < a->n_proj; ++i) key += std::to_string(a->proj_filters[i]) + "x" + std::to_string(a->proj_kernel[i]) + ",";
  if (h->voc_key == key) return GSTK_OK;
  int rc;
  const int mel = c.mel_dim, KB = a->bank_count, F = a->bank_filters, CB = KB * F;
  const std::string cb = std::string(VOCP) + "/CBHG/";
  auto fold = [&](const std::string& bn, int co, std::vector<float>& sc, std::vector<float>& sh) -> int {
    int r;
    for (const char* n : {"gamma", "beta", "moving_mean", "moving_variance"})
      if ((r = need(h, bn + n, co))) return r;
    const auto& ga = *hw(h, bn + "gamma");
    const auto& be = *hw(h, bn + "beta");
    const auto& mu = *hw(h, bn + "moving_mean");
    const auto& va = *hw(h, bn + "moving_variance");
    sc.resize(co);
    sh.resize(co);
    for (int n = 0; n < co; ++n) {
      const double s = (double)ga[n] / std::sqrt((double)va[n] + 1e-3);
      sc[n] = (float)s;
      sh[n] = (float)((double)be[n] - (double)mu[n] * s);
    }
    return GSTK_OK;
  };
  {  // conv bank: combined kernel [KB taps][mel][CB]
    std::vector<float> w((size_t)KB * mel * CB, 0.f), sc(CB), sh(CB), s1, h1;
    const int pad_c = (KB - 1) / 2;
    for (int i = 0; i < KB; ++i) {
      const int ki = i + 1, pad_i = (ki - 1) / 2;
      const std::string conv = cb + "ConvBank_" + std::to_string(i) + "/conv1d/kernel";
      if ((rc = need(h, conv, (size_t)ki * mel * F))) return rc;
      if ((rc = fold(cb + "ConvBank_" + std::to_string(i) + "/batch_normalization/", F, s1, h1))) return rc;
      const auto& wi = *hw(h, conv);
      for (int j = 0; j < ki; ++j) {
        const int tap = j - pad_i + pad_c;   // 0 <= tap < KB: pad_i <= pad_c and ki - 1 - pad_i <= KB - 1 - pad_c
        for (int ch = 0; ch < mel; ++ch)
          for (int f = 0; f < F; ++f) w[((size_t)tap * mel + ch) * CB + (size_t)i * F + f] = wi[((size_t)j * mel + ch) * F + f];
      }
      for (int f = 0; f < F; ++f) {
        sc[(size_t)i * F + f] = s1[f];
        sh[(size_t)i * F + f] = h1[f];
      }
    }
    if ((rc = upload_folded_conv(h, "vocb", w, sc, sh, (size_t)KB * mel, CB))) return rc;
  }
  int cin = CB;
  for (int i = 0; i < a->n_proj; ++i) {
    const int co = a->proj_filters[i], k = a->proj_kernel[i];
    const std::string conv = cb + "Conv1D_Projection/conv1d_" + std::to_string(i) + "/kernel";
    std::vector<float> sc, sh;
    if ((rc = need(h, conv, (size_t)k * cin * co))) return rc;
    if ((rc = fold(cb + "Conv1D_Projection/batch_normalization_" + std::to_string(i) + "/", co, sc, sh))) return rc;
    if ((rc = upload_folded_conv(h, "vocp" + std::to_string(i), *hw(h, conv), sc, sh, (size_t)k * cin, co))) return rc;
    cin = co;
  }
  // per-frame stack of voc_highway_kernel: [Dense(mel)] [Dense(size)] Highwaynet x count, fp32; in tensor-core mode also the
  // B-fragment images of voc_highway_mma_kernel ("vocm_w{l}": [kt][chunk][pair][lane] uint4, "vocm_b{l}": [chunk][32])
  int l = 0;
  auto pack_mma = [&](int layer, const float* W, const float* bias, int K, int ldw, int N, bool highway) -> int {
    if (c.precision != GSTK_PREC_BF16) return GSTK_OK;
    const int nch = highway ? N / 16 : (N + 31) / 32, KT = (K + 15) / 16;
    auto col_of = [&](int ch, int cp) {   // packed column cp of chunk ch -> column of W, or -1
      if (highway) return cp < 16 ? ch * 16 + cp : N + ch * 16 + (cp - 16);
      const int n = ch * 32 + cp;
      return n < N ? n : -1;
    };
    std::vector<__half> img((size_t)KT * nch * 2 * 32 * 8);
    std::vector<float> bp((size_t)nch * 32, 0.f);
    for (int kt = 0; kt < KT; ++kt)
      for (int ch = 0; ch < nch; ++ch)
        for (int pair = 0; pair < 2; ++pair)
          for (int lane = 0; lane < 32; ++lane)
            for (int r = 0; r < 4; ++r)
              for (int e = 0; e < 2; ++e) {
                const int tile = pair * 2 + r / 2, k = kt * 16 + (lane % 4) * 2 + e + 8 * (r % 2);
                const int col = col_of(ch, tile * 8 + lane / 4);
                const float v = (col >= 0 && k < K) ? W[(size_t)k * ldw + col] : 0.f;
                img[(((((size_t)kt * nch + ch) * 2 + pair) * 32 + lane) * 4 + r) * 2 + e] = __float2half_rn(v);
              }
    for (int ch = 0; ch < nch; ++ch)
      for (int cp = 0; cp < 32; ++cp) {
        const int col = col_of(ch, cp);
        if (col >= 0) bp[(size_t)ch * 32 + cp] = bias[col];
      }
    int r;
    if ((r = upload_derived(h, "vocm_w" + std::to_string(layer), img.data(), img.size() * 2))) return r;
    return upload_derived(h, "vocm_b" + std::to_string(layer), bp.data(), bp.size() * 4);
  };
  auto dense = [&](const std::string& base, int K, int N) -> int {
    int r;
    if ((r = need(h, base + "kernel", (size_t)K * N))) return r;
    if ((r = need(h, base + "bias", (size_t)N))) return r;
    if ((r = upload_derived(h, "voch_w" + std::to_string(l), hw(h, base + "kernel")->data(), (size_t)K * N * 4))) return r;
    if ((r = upload_derived(h, "voch_b" + std::to_string(l), hw(h, base + "bias")->data(), (size_t)N * 4))) return r;
    if ((r = pack_mma(l, hw(h, base + "kernel")->data(), hw(h, base + "bias")->data(), K, N, N, false))) return r;
    ++l;
    return GSTK_OK;
  };
  if (cin != mel && (rc = dense(cb + "Conv1D_Projection/dense/", cin, mel))) return rc;
  const int hs = a->highway_size;
  if (mel != hs && (rc = dense(cb + "Highwaynet/dense/", mel, hs))) return rc;
  for (int i = 0; i < a->highway_count; ++i) {
    const std::string base = cb + "Highwaynet/highwaynet_" + std::to_string(i) + "/";
    for (const char* nm : {"Dense_Relu/", "Dense_Sigmoid/"}) {
      if ((rc = need(h, base + nm + "kernel", (size_t)hs * hs))) return rc;
      if ((rc = need(h, base + nm + "bias", (size_t)hs))) return rc;
    }
    const auto &wr = *hw(h, base + "Dense_Relu/kernel"), &ws = *hw(h, base + "Dense_Sigmoid/kernel");
    const auto &br = *hw(h, base + "Dense_Relu/bias"), &bs = *hw(h, base + "Dense_Sigmoid/bias");
    std::vector<float> w((size_t)hs * 2 * hs), b((size_t)2 * hs);
    for (int k = 0; k < hs; ++k)
      for (int n = 0; n < hs; ++n) {
        w[(size_t)k * 2 * hs + n] = wr[(size_t)k * hs + n];
        w[(size_t)k * 2 * hs + hs + n] = ws[(size_t)k * hs + n];
      }
    for (int n = 0; n < hs; ++n) {
      b[n] = br[n];
      b[hs + n] = bs[n];
    }
    if ((rc = upload_derived(h, "voch_w" + std::to_string(l), w.data(), w.size() * 4))) return rc;
    if ((rc = upload_derived(h, "voch_b" + std::to_string(l), b.data(), b.size() * 4))) return rc;
    if ((rc = pack_mma(l, w.data(), b.data(), hs, 2 * hs, hs, true))) return rc;
    ++l;
  }
  const int u = a->rnn_size;
  if ((rc = upload_bilstm_proj(h, "vocx", cb + "RNN/", hs, u))) return rc;
  {
    const int N = a->spectrogram_dim, Np = (N + VOC_DENSE_ALIGN - 1) / VOC_DENSE_ALIGN * VOC_DENSE_ALIGN;
    const std::string base = std::string(VOCP) + "/Dense/";
    if ((rc = need(h, base + "kernel", (size_t)2 * u * N))) return rc;
    if ((rc = need(h, base + "bias", (size_t)N))) return rc;
    const auto &w = *hw(h, base + "kernel"), &b = *hw(h, base + "bias");
    std::vector<float> wp((size_t)2 * u * Np, 0.f), bp((size_t)Np, 0.f), one((size_t)Np, 1.f);
    for (int k = 0; k < 2 * u; ++k)
      for (int n = 0; n < N; ++n) wp[(size_t)k * Np + n] = w[(size_t)k * N + n];
    for (int n = 0; n < N; ++n) bp[n] = b[n];
    if ((rc = upload_folded_conv(h, "vocd", wp, one, bp, (size_t)2 * u, Np))) return rc;
  }
  h->voc_key = key;
  return GSTK_OK;
}

// Bidirectional(LSTM(u, return_sequences=True)) over xs [B][T][8u] (input projections + biases of both directions, gate-interleaved
// columns, see upload_bilstm_proj) -> out [B][T][2u] = [forward | backward]  (Encoder: Taco2.py:39-43; CBHG: Taco2.py:358-362)
// true when run_bilstm will take the cluster kernel, which also reads the projections as fp16 rows of the padded matrix
bool bilstm_cluster(bool bf16, int u) { return bf16 && u == BC_U && !getenv("GSTK_ENC_BILSTM"); }

int run_bilstm(GstkHandle* h, const float* xs, const float* uf, const float* ub, float* o_enc, int B, int T, int u, bool bf16,
               cudaStream_t st, const __half* xs16 = nullptr, int R = 0, int PADL = 0) {
  int rc;
  // tensor-core mode, u = 256: independent clusters of 4 CTAs per (16 utterances, direction), recurrent kernel in registers, h through
  // distributed shared memory (encoder.cuh).  GSTK_ENC_BILSTM=stream|persistent|ffma|tc selects one of the older kernels.
  if (bilstm_cluster(bf16, u)) {
    BilstmClParams cp;
    cp.xs = xs16 ? nullptr : xs; cp.xs16 = xs16; cp.R = R; cp.PADL = PADL; cp.Uf = uf; cp.Ub = ub; cp.out = o_enc; cp.B = B; cp.T = T;
    encoder_bilstm_cluster_kernel<<<dim3(BC_CL, (B + BC_NB - 1) / BC_NB, 2), BC_THREADS, 0, st>>>(cp);
    h->launches++;
    CK(cudaGetLastError());
    return GSTK_OK;
  }
  // persistent kernel (recurrent kernels resident in shared memory over 2 * u/4 co-resident CTAs, one grid barrier per step)
  // whenever its grid fits the device: measured 8 us + 0.045 us per utterance per step against a flat 27 us for the
  // streaming kernel (B200, u = 256).  GSTK_ENC_BILSTM=stream|persistent forces one (A/B measurements).
  const char* force = getenv("GSTK_ENC_BILSTM");
  const size_t psm = bilstm_persistent_smem(u);
  const bool fits = u % BL_KC == 0 && 2 * (u / BL_HU) <= h->num_sms && psm <= 200 * 1024;
  bool persistent = fits;
  if (force && !strcmp(force, "stream")) persistent = false;
  if (force && !strcmp(force, "persistent")) persistent = fits;
  // tensor-core mode, u = 256: mma.sync form of the persistent kernel (fp16 h exchange, recurrent slice in registers)
  const bool tcl = persistent && bf16 && u == BT_U && !(force && !strcmp(force, "ffma"));   // "tc" or unset
  if (persistent) {
    void* hb;
    const size_t hbytes = (size_t)4 * BL_ROWS * u * (tcl ? 2 : 4);
    if ((rc = slot_reserve(h, SL_ENC_H, hbytes, &hb))) return rc;
    if (tcl) CK(cudaFuncSetAttribute(encoder_bilstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BT_SMEM));
    else CK(cudaFuncSetAttribute(encoder_bilstm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
    for (int b0 = 0; b0 < B; b0 += BL_ROWS) {   // <= 256 utterances per launch
      CK(cudaMemsetAsync(hb, 0, hbytes, st));
      if ((rc = reset_barrier(h, st))) return rc;
      const float* xs0 = (const float*)xs + (size_t)b0 * T * 8 * u;
      float* out0 = (float*)o_enc + (size_t)b0 * T * 2 * u;
      const int Bc = std::min(BL_ROWS, B - b0);
      if (tcl) {
        BilstmTcParams bp;
        bp.xs = xs0; bp.Uf = uf; bp.Ub = ub; bp.out = out0; bp.hbuf = (__half*)hb; bp.gb = h->gb; bp.B = Bc; bp.T = T;
        void* args[] = {&bp};
        CK(cudaLaunchCooperativeKernel((void*)encoder_bilstm_tc_kernel, dim3(2 * (u / BL_HU)), dim3(BL_THREADS), args, BT_SMEM, st));
      } else {
        BilstmParams bp;
        bp.xs = xs0; bp.Uf = uf; bp.Ub = ub; bp.out = out0; bp.hbuf = (float*)hb; bp.gb = h->gb; bp.B = Bc; bp.T = T; bp.u = u;
        void* args[] = {&bp};
        CK(cudaLaunchCooperativeKernel((void*)encoder_bilstm_persistent_kernel, dim3(2 * (u / BL_HU)), dim3(BL_THREADS), args, psm, st));
      }
      h->launches++;
    }
  } else {
    constexpr int NB = 4;
    dim3 grid((B + NB - 1) / NB, 2);
    const size_t smem = (size_t)2 * NB * u * 4;
    encoder_bilstm_kernel<NB><<<grid, u, smem, st>>>((const float*)xs, uf, ub, (float*)o_enc, B, T);
    h->launches++;
  }
  CK(cudaGetLastError());
  return GSTK_OK;
}

template <int BT>
int launch_decoder_fp32(GstkHandle* h, DecParams& p, cudaStream_t st) {
  const size_t smem = decoder_fp32_smem_bytes(p, BT);
  if (smem > 227 * 1024) return fail(h, GSTK_EINVAL, "decoder needs %zu B of shared memory (key_time too large)", smem);
  CK(cudaFuncSetAttribute(decoder_fp32_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decoder_fp32_kernel<BT>, DEC_THREADS, smem));
  if (occ < 1) return fail(h, GSTK_EINVAL, "decoder kernel does not fit on an SM");
  const int grid = h->num_sms;
  void* args[] = {&p};
  CK(cudaEventRecord(h->ev0, st));
  CK(cudaLaunchCooperativeKernel((void*)decoder_fp32_kernel<BT>, dim3(grid), dim3(DEC_THREADS), args, smem, st));
  CK(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  h->ev_stream = st;
  h->launches++;
  return GSTK_OK;
}

}  // namespace

extern "C" {

int gstk_version(void) { return GSTK_VERSION; }

const char* gstk_last_error(GstkHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int gstk_create(const GstkConfig* cfg, GstkHandle** out) {
  GstkHandle* h = nullptr;
  if (!cfg || !out) return fail(h, GSTK_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->version != GSTK_VERSION) return fail(h, GSTK_EINVAL, "GstkConfig.version %d != %d", cfg->version, GSTK_VERSION);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, GSTK_ENODEVICE, "no CUDA device available (libgsttaco has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(h, GSTK_ENODEVICE, "device %d out of range (%d devices)", cfg->device, ndev);
  const GstkConfig& c = *cfg;
  // reference raises ValueError for unsupported attention types (Taco2.py:75)
  if (c.attention_type < GSTK_ATT_SMA || c.attention_type > GSTK_ATT_LSA)
    return fail(h, GSTK_EINVAL, "Unsupported attention type: %d", c.attention_type);
  if (c.style_heads > 0 && c.style_size % c.style_heads != 0)
    return fail(h, GSTK_EINVAL, "size must be divisible by num_heads. ('%d' %% '%d' != 0)", c.style_size, c.style_heads);
  const int PD = c.mel_dim * c.step_reduction + 1;
  if (c.mel_dim < 1 || c.step_reduction < 1 || c.mel_dim > DEC_THREADS || PD > DEC_THREADS || c.prenet0 > DEC_THREADS ||
      c.prenet1 > DEC_THREADS || c.attention_size > DEC_THREADS || c.prenet0 < 1 || c.prenet1 < 1 || c.attention_size < 1)
    return fail(h, GSTK_EINVAL, "decoder layer widths must be in [1,%d]", DEC_THREADS);
  if ((c.prenet1 + c.attention_size) % 4 || c.lstm0 % LSTM_HU || c.lstm1 % LSTM_HU || c.lstm0 < LSTM_HU || c.lstm1 < LSTM_HU)
    return fail(h, GSTK_EINVAL, "LSTM sizes must be multiples of %d and prenet+attention size a multiple of 4", LSTM_HU);
  if (c.attention_type == GSTK_ATT_LSA && (c.lsa_filters < 1 || c.lsa_filters > 32 || c.lsa_kernel < 1))
    return fail(h, GSTK_EINVAL, "LSA conv filters must be in [1,32]");
  if (c.enc_dim < 1 || (c.gst_use && c.enc_dim <= c.style_size)) return fail(h, GSTK_EINVAL, "bad enc_dim");
  if (c.gst_use) {
    if (c.ref_layers < 1 || c.ref_layers > 8) return fail(h, GSTK_EINVAL, "1..8 reference-encoder conv layers supported");
    for (int i = 0; i < c.ref_layers; ++i) {
      if (c.ref_kernel[i] != 3 || c.ref_stride[i] != 2)
        return fail(h, GSTK_EINVAL, "reference-encoder convs must be 3x3 stride 2");
      const int f = c.ref_filters[i];
      if (!(f == 32 || f == 64 || f == 128 || f == 256)) return fail(h, GSTK_EINVAL, "conv filters must be 32/64/128/256");
    }
    if (c.ref_gru < 2 || c.ref_gru > 128 || (c.ref_gru & 1)) return fail(h, GSTK_EINVAL, "GRU size must be even and <= 128");
  }
  if (c.precision != GSTK_PREC_FP32 && c.precision != GSTK_PREC_BF16) return fail(h, GSTK_EINVAL, "bad precision");
  if (cudaSetDevice(c.device) != cudaSuccess) return fail(h, GSTK_ECUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess) return fail(h, GSTK_ECUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(h, GSTK_ENODEVICE, "device %d is sm_%d%d; libgsttaco is built for sm_100a only", c.device, prop.major, prop.minor);
  if (c.precision == GSTK_PREC_BF16) {
    std::string why;
    if (!bf16_config_supported(c, why)) return fail(h, GSTK_EINVAL, "bf16 tensor-core path: %s", why.c_str());
  }
  h = new GstkHandle();
  h->cfg = c;
  h->num_sms = prop.multiProcessorCount;
  if (cudaMalloc((void**)&h->gb, sizeof(GridBarrier)) != cudaSuccess || cudaEventCreate(&h->ev0) != cudaSuccess ||
      cudaEventCreate(&h->ev1) != cudaSuccess || cudaEventCreate(&h->ev2) != cudaSuccess) {
    delete h;
    h = nullptr;
    return fail(h, GSTK_ECUDA, "allocation failed in gstk_create");
  }
  cudaMemset(h->gb, 0, sizeof(GridBarrier));
  *out = h;
  return GSTK_OK;
}

int gstk_destroy(GstkHandle* h) {
  if (!h) return GSTK_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (auto& kv : h->dev_w) cudaFree(kv.second.p);
  for (auto& kv : h->derived) cudaFree(kv.second.p);
  for (auto& s : h->slots) cudaFree(s.p);
  bf16_release(h->bf16);
  v2_release(h->v2);
  sb_release(h->sb);
  cudaFree(h->gb);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev2);
  if (h->st_in) {
    cudaStreamDestroy(h->st_in);
    cudaEventDestroy(h->ev_in);
  }
  for (cudaEvent_t e : h->slot_ev)
    if (e) cudaEventDestroy(e);
  if (h->st_copy) {
    cudaStreamDestroy(h->st_copy);
    cudaEventDestroy(h->ev_chunk);
    cudaEventDestroy(h->ev_copied);
  }
  delete h;
  return GSTK_OK;
}

int gstk_load_weights(GstkHandle* h, const GstkTensorDesc* tensors, int32_t n) {
  if (!h || (!tensors && n > 0)) return fail(h, GSTK_EINVAL, "null argument");
  DEVICE_GUARD(h, h->cfg.device);
  for (int i = 0; i < n; ++i) {
    const GstkTensorDesc& t = tensors[i];
    if (!t.name || !t.data || t.ndim < 0 || t.ndim > 4) return fail(h, GSTK_EINVAL, "bad tensor descriptor %d", i);
    size_t cnt = 1;
    for (int k = 0; k < t.ndim; ++k) cnt *= (size_t)t.shape[k];
    std::vector<float> v(cnt);
    if (is_device_ptr(t.data)) CK(cudaMemcpy(v.data(), t.data, cnt * 4, cudaMemcpyDeviceToHost));
    else memcpy(v.data(), t.data, cnt * 4);
    DevBuf& b = h->dev_w[t.name];
    if (b.bytes < cnt * 4) {
      if (b.p) CK(cudaFree(b.p));
      CK(cudaMalloc(&b.p, std::max(cnt * 4, (size_t)16)));
      b.bytes = cnt * 4;
    }
    CK(cudaMemcpy(b.p, v.data(), cnt * 4, cudaMemcpyHostToDevice));
    h->host_w[t.name] = std::move(v);
  }
  h->dec_ready = false;
  h->gst_ready = false;
  h->post_key.clear();
  h->enc_key.clear();
  h->voc_key.clear();
  // the bf16 decoder keeps its own packed images of the LSTM / dense kernels: drop them so that prepare_decoder rebuilds
  // them from the new weights (a stale image would silently mix two checkpoints)
  CK(cudaDeviceSynchronize());
  bf16_release(h->bf16);
  v2_release(h->v2);
  sb_release(h->sb);
  return GSTK_OK;
}

namespace {
__global__ void fill_i32_kernel(int* p, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
// stop indices that were never set -> `steps` (GstkDecodeArgs::out_stop_index)
__global__ void stop_index_finish_kernel(int* idx, int n, int steps) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (idx[i] == STOP_UNSET) idx[i] = steps;
}
// One launch of the bf16 tensor-core decoder for a batch chunk.  Default: the barrier-phased kernel (decoder_bf16.cuh).
// GSTK_DECODER=dataflow selects the barrier-free variant (decoder_bf16_v2.cuh) wherever it applies (free-running, SMA, default
// widths): parity-green, but at 29.0 us vs 26.9 us per step (batch 256) it is not the faster one - see DESIGN.md 3.1b.
int run_bf16_decoder(GstkHandle* h, DecParams& p, int kernel, cudaStream_t st, cudaEvent_t e0) {
  if (kernel == GSTK_KERNEL_AUTO) {
    const char* which = getenv("GSTK_DECODER");   // read per call: tools and tests flip it
    if (which && !strcmp(which, "barrier")) kernel = GSTK_KERNEL_BATCH;
    else if (which && !strcmp(which, "dataflow")) kernel = v2_usable(h->bf16, p, h->num_sms) ? GSTK_KERNEL_DATAFLOW : GSTK_KERNEL_BATCH;
    else if (which && *which) return fail(h, GSTK_EINVAL, "GSTK_DECODER=%s: expected barrier or dataflow", which);
    else kernel = (!p.early_stop && sb_usable(h->bf16, p, h->num_sms)) ? GSTK_KERNEL_SMALL : GSTK_KERNEL_BATCH;
  }
  switch (kernel) {
    case GSTK_KERNEL_SMALL: {   // batch <= 16, free running, SMA, default widths: the latency kernel (decoder_bf16_sb.cuh)
      if (p.early_stop || !sb_usable(h->bf16, p, h->num_sms))
        return fail(h, GSTK_EINVAL, "GSTK_KERNEL_SMALL: needs a free-running SMA decode of batch <= %d, key_time <= 256, no early_stop", SB_MAXB);
      int rc = sb_prepare(h->sb, h->host_w, h->err);
      if (rc) return rc;
      return sb_decode(h->bf16, h->sb, p, h->num_sms, st, e0, h->ev1, h->launches, h->err);
    }
    case GSTK_KERNEL_DATAFLOW: {
      if (!v2_usable(h->bf16, p, h->num_sms)) return fail(h, GSTK_EINVAL, "GSTK_KERNEL_DATAFLOW: needs a free-running SMA decode at the default widths");
      int rc = v2_prepare(h->v2, h->cfg, h->host_w, h->err);
      if (rc) return rc;
      return v2_decode(h->bf16, h->v2, h->cfg, p, h->num_sms, st, e0, h->ev1, h->launches, h->err);
    }
    case GSTK_KERNEL_BATCH:
      return bf16_decode(h->bf16, h->cfg, p, h->num_sms, st, e0, h->ev1, h->launches, h->err);
  }
  return fail(h, GSTK_EINVAL, "bad kernel selector %d", kernel);
}
}  // namespace

int gstk_decode(GstkHandle* h, const GstkDecodeArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  const GstkConfig& c = h->cfg;
  DEVICE_GUARD(h, c.device);
  trace(h, "decode: enter");
  int rc = prepare_decoder(h);
  if (rc) return rc;
  const int B = a->batch, Tv = a->key_time, T = a->steps;
  if (B < 1 || Tv < 1 || T < 0) return fail(h, GSTK_EINVAL, "batch/key_time/steps must be positive");
  if (Tv > GSTK_MAX_TV) return fail(h, GSTK_EINVAL, "key_time %d > %d", Tv, GSTK_MAX_TV);
  if (a->mode != GSTK_MODE_FREE && a->mode != GSTK_MODE_TEACHER) return fail(h, GSTK_EINVAL, "bad mode");
  if (a->mode == GSTK_MODE_TEACHER && !a->teacher_mels && T > 0) return fail(h, GSTK_EINVAL, "teacher mode needs teacher_mels");
  if (a->rng_mode < GSTK_RNG_NONE || a->rng_mode > GSTK_RNG_PHILOX) return fail(h, GSTK_EINVAL, "bad rng_mode");
  if (a->early_stop && a->mode != GSTK_MODE_FREE) return fail(h, GSTK_EINVAL, "early_stop applies to free-running decodes only");
  if (a->kernel < GSTK_KERNEL_AUTO || a->kernel > GSTK_KERNEL_DATAFLOW) return fail(h, GSTK_EINVAL, "bad kernel selector %d", a->kernel);
  const bool need_noise = c.sigmoid_noise > 0.f && c.attention_type != GSTK_ATT_LSA;
  if (a->rng_mode == GSTK_RNG_EXTERNAL &&
      ((c.prenet_dropout > 0.f && (!a->keep0 || !a->keep1)) || (need_noise && !a->noise)) && T > 0)
    return fail(h, GSTK_EINVAL, "rng_mode EXTERNAL needs keep0/keep1%s", need_noise ? "/noise" : "");
  if (!a->encodings && !(a->enc_text && a->gst && c.gst_use))
    return fail(h, GSTK_EINVAL, "pass encodings, or enc_text + gst on a GST-enabled handle");
  cudaStream_t st = (cudaStream_t)a->stream;
  const int A = c.attention_size, U0 = c.lstm0, U1 = c.lstm1, mel = c.mel_dim, r = c.step_reduction;
  const int PD = mel * r + 1, E = c.enc_dim, S = c.style_size, Dt = E - (c.gst_use ? S : 0);
  h->pending.clear();
  const std::string d = DEC;

  // ---- inputs
  const void *enc = nullptr, *enc_text = nullptr, *gst = nullptr, *teacher = nullptr, *keep0 = nullptr, *keep1 = nullptr,
             *noise = nullptr, *init_mel = nullptr, *init_align = nullptr, *init_cum = nullptr, *init_states = nullptr;
  if (a->encodings) {
    if ((rc = stage_in(h, SL_ENC, a->encodings, (size_t)B * Tv * E * 4, st, &enc))) return rc;
  } else {
    if ((rc = stage_in(h, SL_ENC_TEXT, a->enc_text, (size_t)B * Tv * Dt * 4, st, &enc_text))) return rc;
    if ((rc = stage_in(h, SL_GST_IN, a->gst, (size_t)B * S * 4, st, &gst))) return rc;
  }
  long long ts_b = a->teacher_stride_b ? a->teacher_stride_b : (long long)T * mel;
  long long ts_t = a->teacher_stride_t ? a->teacher_stride_t : mel;
  if (a->mode == GSTK_MODE_TEACHER && T > 0) {
    if (is_device_ptr(a->teacher_mels)) {
      teacher = a->teacher_mels;
    } else {
      if (ts_t != mel && T > 1) return fail(h, GSTK_EINVAL, "host teacher_mels must have contiguous frames");
      // copy the strided [B, T, mel] view row by row into a dense staging buffer
      void* dbuf;
      if ((rc = slot_reserve(h, SL_TEACHER, (size_t)B * T * mel * 4, &dbuf))) return rc;
      CK(cudaMemcpy2DAsync(dbuf, (size_t)T * mel * 4, a->teacher_mels, (size_t)ts_b * 4, (size_t)T * mel * 4, B,
                           cudaMemcpyHostToDevice, st));
      teacher = dbuf;
      ts_b = (long long)T * mel;
      ts_t = mel;
    }
  }
  if (a->rng_mode == GSTK_RNG_EXTERNAL && T > 0) {
    if ((rc = stage_in(h, SL_KEEP0, a->keep0, (size_t)T * B * c.prenet0 * 4, st, &keep0))) return rc;
    if ((rc = stage_in(h, SL_KEEP1, a->keep1, (size_t)T * B * c.prenet1 * 4, st, &keep1))) return rc;
    if (need_noise && (rc = stage_in(h, SL_NOISE, a->noise, (size_t)T * B * Tv * 4, st, &noise))) return rc;
  }
  if ((rc = stage_in(h, SL_INIT_MEL, a->init_mel, (size_t)B * mel * 4, st, &init_mel))) return rc;
  if ((rc = stage_in(h, SL_INIT_ALIGN, a->init_alignment, (size_t)B * Tv * 4, st, &init_align))) return rc;
  if ((rc = stage_in(h, SL_INIT_CUM, a->init_cum_alignment, (size_t)B * Tv * 4, st, &init_cum))) return rc;
  if ((rc = stage_in(h, SL_INIT_STATES, a->init_states, (size_t)2 * B * (U0 + U1) * 4, st, &init_states))) return rc;

  // ---- outputs
  void *o_mel, *o_stop, *o_align, *o_states, *o_cum, *o_ctx;
  if ((rc = stage_out(h, SL_OUT_MEL, a->out_mel, (size_t)B * T * r * mel * 4, &o_mel))) return rc;
  if ((rc = stage_out(h, SL_OUT_STOP, a->out_stop, (size_t)B * T * 4, &o_stop))) return rc;
  if ((rc = stage_out(h, SL_OUT_ALIGN, a->out_alignment, (size_t)B * T * Tv * 4, &o_align))) return rc;
  if ((rc = stage_out(h, SL_OUT_STATES, a->out_states, (size_t)2 * B * (U0 + U1) * 4, &o_states))) return rc;
  if ((rc = stage_out(h, SL_OUT_CUM, a->out_cum_alignment, (size_t)B * Tv * 4, &o_cum))) return rc;
  if ((rc = stage_out(h, SL_OUT_CTX, a->out_context, (size_t)B * A * 4, &o_ctx))) return rc;

  // ---- workspace
  void *vproj, *gbias, *xin, *h1, *h2, *c1, *c2, *align, *cum;
  if ((rc = slot_reserve(h, SL_VPROJ, (size_t)B * Tv * A * 4, &vproj))) return rc;
  if ((rc = slot_reserve(h, SL_GBIAS, (size_t)B * A * 4, &gbias))) return rc;
  if ((rc = slot_reserve(h, SL_XIN, (size_t)B * (c.prenet1 + A) * 4, &xin))) return rc;
  if ((rc = slot_reserve(h, SL_H1, (size_t)2 * B * U0 * 4, &h1))) return rc;
  if ((rc = slot_reserve(h, SL_H2, (size_t)2 * B * U1 * 4, &h2))) return rc;
  if ((rc = slot_reserve(h, SL_C1, (size_t)B * U0 * 4, &c1))) return rc;
  if ((rc = slot_reserve(h, SL_C2, (size_t)B * U1 * 4, &c2))) return rc;
  if ((rc = slot_reserve(h, SL_ALIGN, (size_t)2 * B * Tv * 4, &align))) return rc;
  if ((rc = slot_reserve(h, SL_CUM, (size_t)B * Tv * 4, &cum))) return rc;

  // ---- stop bookkeeping (early stop / stop indices, Model.py:380)
  const bool want_stop = a->early_stop || a->out_stop_index || a->out_steps_done;
  void *stop_idx = nullptr, *stop_state = nullptr;
  int steps_done = T;
  if (want_stop && T > 0) {
    if ((rc = slot_reserve(h, SL_STOP_IDX, (size_t)B * 4, &stop_idx))) return rc;
    if ((rc = slot_reserve(h, SL_STOP_STATE, 16, &stop_state))) return rc;
    fill_i32_kernel<<<(B + 255) / 256, 256, 0, st>>>((int*)stop_idx, B, STOP_UNSET);
    h->launches++;
    if (a->early_stop) {   // rows of the device output tensors beyond the early exit read as zero
      if (o_mel) CK(cudaMemsetAsync(o_mel, 0, (size_t)B * T * r * mel * 4, st));
      if (o_stop) CK(cudaMemsetAsync(o_stop, 0, (size_t)B * T * 4, st));
      if (o_align) CK(cudaMemsetAsync(o_align, 0, (size_t)B * T * Tv * 4, st));
    }
  }

  // ---- loop-invariant value projection V' = Dense_V(encodings)  (Steps.py:123, hoisted)
  const float* Wv = dw(h, d + "/Attention/Value/kernel");
  const float* bv = dw(h, d + "/Attention/Value/bias");
  // tensor-core handles: one tcgen05 GEMM [B T_v, E] x [E, A] (the k = 1 form of the conv layers; fp16 operand rows [gst || text])
  const bool vproj_tc = c.precision == GSTK_PREC_BF16 && E % 16 == 0 && S % 4 == 0 && A % 16 == 0 && A <= 256 && T > 0 && h->derived.count("decv_wt");
  if (vproj_tc) {
    const long long rows = (long long)B * Tv;
    void* a16;
    if ((rc = slot_reserve(h, SL_VPROJ_A16, (size_t)(rows + PC_BM) * E * 2, &a16))) return rc;
    const long long n4 = rows * (E / 4);
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16);
    if (enc) value_operand_f16_kernel<<<blocks, 256, 0, st>>>((const float*)enc, nullptr, (__half*)a16, rows, Tv, 0, E);
    else value_operand_f16_kernel<<<blocks, 256, 0, st>>>((const float*)enc_text, (const float*)gst, (__half*)a16, rows, Tv, S, Dt);
    h->launches++;
    CK(cudaGetLastError());
    PostConvParams q{};
    q.X = a16;
    q.W = h->derived["decv_w"].p;
    q.shift = dd(h, "decv_shift");
    q.out = (float*)vproj;
    q.Mtotal = rows;
    q.C = E; q.K = E; q.N = A;
    q.R = Tv; q.PADL = 0; q.T = Tv;
    if ((rc = launch_conv_layer(h, q, true, 1, 0, h->derived["decv_wt"].p, st))) return rc;
  } else if (enc) {
    if ((rc = launch_sgemm(h, (const float*)enc, E, Wv, bv, nullptr, 1, (float*)vproj, B * Tv, A, E, st))) return rc;
  } else {
    // [gst || enc_text] . Wv = enc_text . Wv[S:] + gst . Wv[:S]   (GST_Concated_Encoder folded, GST.py:121-124)
    if ((rc = launch_sgemm(h, (const float*)gst, S, Wv, nullptr, nullptr, 1, (float*)gbias, B, A, S, st))) return rc;
    if ((rc = launch_sgemm(h, (const float*)enc_text, Dt, Wv + (size_t)S * A, bv, (const float*)gbias, Tv,
                           (float*)vproj, B * Tv, A, Dt, st))) return rc;
  }

  DecParams p;
  memset(&p, 0, sizeof(p));
  p.Tv = Tv; p.T = T; p.To = T; p.mode = a->mode; p.rng_mode = a->rng_mode; p.att_type = c.attention_type;
  p.mel = mel; p.r = r; p.P0 = c.prenet0; p.P1 = c.prenet1; p.A = A; p.U0 = U0; p.U1 = U1; p.PD = PD;
  p.lsa_filters = c.lsa_filters; p.lsa_kernel = c.lsa_kernel; p.lsa_cumulate = c.lsa_cumulate; p.lsa_smoothing = c.lsa_smoothing;
  p.drop_rate = c.prenet_dropout;
  p.drop_scale = c.prenet_dropout > 0.f ? 1.0f / (1.0f - c.prenet_dropout) : 1.0f;
  p.sigmoid_noise = c.sigmoid_noise;
  p.seed = a->seed; p.step_offset = a->step_offset;
  p.W0 = dw(h, d + "/Prenet/dense/kernel"); p.b0 = dw(h, d + "/Prenet/dense/bias");
  p.W1 = dw(h, d + "/Prenet/dense_1/kernel"); p.b1 = dw(h, d + "/Prenet/dense_1/bias");
  p.Wq = dw(h, d + "/Attention/Query/kernel"); p.bq = dw(h, d + "/Attention/Query/bias");
  p.att_v = dw(h, d + "/Attention/attention_v"); p.att_sb = dw(h, d + "/Attention/attention_score_bias");
  p.lsa_cw = dw(h, d + "/Attention/Alignment_Conv/kernel"); p.lsa_cb = dw(h, d + "/Attention/Alignment_Conv/bias");
  p.lsa_dw = dw(h, d + "/Attention/Alignment_Dense/kernel"); p.lsa_db = dw(h, d + "/Attention/Alignment_Dense/bias");
  p.lsa_bias = dw(h, d + "/Attention/bias");
  p.Wp = dw(h, d + "/Projection/kernel"); p.bp = dw(h, d + "/Projection/bias");
  p.L1pk = dd(h, "L1pk"); p.L1b = dw(h, d + "/RNN/cell_0/bias");
  p.L2pk = dd(h, "L2pk"); p.L2b = dw(h, d + "/RNN/cell_1/bias");
  p.ts_b = ts_b; p.ts_t = ts_t;
  p.keep0 = (const float*)keep0; p.keep1 = (const float*)keep1; p.noise = (const float*)noise;
  p.xin = (float*)xin; p.h1 = (float*)h1; p.h2 = (float*)h2; p.c1 = (float*)c1; p.c2 = (float*)c2;
  p.align = (float*)align; p.cum = (float*)cum;
  p.gb = h->gb;
  p.rngB = B;
  p.early_stop = a->early_stop ? 1 : 0;
  p.stop_state = (unsigned int*)stop_state;
  { const char* dbg = getenv("GSTK_DEBUG"); p.debug_flags = dbg ? atoi(dbg) : 0; }
  int steps_done_max = 0;   // over the batch chunks

  // The fp32 kernel takes the whole batch in one launch; the bf16 tensor-core kernel works on chunks of
  // <= 256 utterances (two 128-row MMA tiles).  Utterances are independent, so chunks run back to back.
  const int chunk = c.precision == GSTK_PREC_BF16 ? TC_MAX_B : B;
  bool first_launch = true;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int Bc = std::min(chunk, B - b0);
    p.B = Bc;
    p.rng_b0 = b0;
    p.row_offset = a->row_offset + (unsigned int)b0;
    p.vproj = (const float*)vproj + (size_t)b0 * Tv * A;
    p.teacher = teacher ? (const float*)teacher + (size_t)b0 * ts_b : nullptr;
    p.init_mel = init_mel ? (const float*)init_mel + (size_t)b0 * mel : nullptr;
    p.out_mel = o_mel ? (float*)o_mel + (size_t)b0 * T * r * mel : nullptr;
    p.out_stop = o_stop ? (float*)o_stop + (size_t)b0 * T : nullptr;
    p.out_align = o_align ? (float*)o_align + (size_t)b0 * T * Tv : nullptr;
    p.out_ctx = o_ctx ? (float*)o_ctx + (size_t)b0 * A : nullptr;
    p.stop_index = stop_idx ? (int*)stop_idx + b0 : nullptr;
    p.t_base = 0;
    int launched_end = T;   // last step covered by a launch of this batch chunk (time-chunked decodes may stop launching early)
    if (stop_state) {   // [0] rows stopped = 0, [1] valid steps = T until a kernel leaves early
      const unsigned int init[2] = {0u, (unsigned int)T};
      CK(cudaMemcpyAsync(stop_state, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    // ---- initial state: buffers with index 1 hold "step -1"
    float* h1p = (float*)h1 + (size_t)Bc * U0;
    float* h2p = (float*)h2 + (size_t)Bc * U1;
    if (init_states) {
      const float* s = (const float*)init_states;  // [h1 | c1 | h2 | c2], each [B, U]
      CK(cudaMemcpyAsync(h1p, s + (size_t)b0 * U0, (size_t)Bc * U0 * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(c1, s + (size_t)B * U0 + (size_t)b0 * U0, (size_t)Bc * U0 * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(h2p, s + (size_t)2 * B * U0 + (size_t)b0 * U1, (size_t)Bc * U1 * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(c2, s + (size_t)2 * B * U0 + (size_t)B * U1 + (size_t)b0 * U1, (size_t)Bc * U1 * 4,
                         cudaMemcpyDeviceToDevice, st));
    } else {
      CK(cudaMemsetAsync(h1p, 0, (size_t)Bc * U0 * 4, st));
      CK(cudaMemsetAsync(h2p, 0, (size_t)Bc * U1 * 4, st));
      CK(cudaMemsetAsync(c1, 0, (size_t)Bc * U0 * 4, st));
      CK(cudaMemsetAsync(c2, 0, (size_t)Bc * U1 * 4, st));
    }
    float* alignp = (float*)align + (size_t)Bc * Tv;
    if (init_align) {
      CK(cudaMemcpyAsync(alignp, (const float*)init_align + (size_t)b0 * Tv, (size_t)Bc * Tv * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      // initial_alignment_fn: one-hot at 0 (Steps.py:201-206); zeros for LSA (Layers.py:356)
      init_alignment_kernel<<<(Bc * Tv + 255) / 256, 256, 0, st>>>(alignp, Bc, Tv, c.attention_type != GSTK_ATT_LSA);
      h->launches++;
    }
    if (init_cum) CK(cudaMemcpyAsync(cum, (const float*)init_cum + (size_t)b0 * Tv, (size_t)Bc * Tv * 4, cudaMemcpyDeviceToDevice, st));
    else CK(cudaMemsetAsync(cum, 0, (size_t)Bc * Tv * 4, st));
    if ((rc = reset_barrier(h, st))) return rc;

    if (T > 0) {
      // Host output buffers + a long decode on the fast path: run the steps as a few launches with in-place state hand-over (even
      // chunk lengths keep the "step -1" state in buffer 1) and copy each chunk's outputs to the host on a second stream
      // while the next chunk decodes - the 237 MB of alignments / mels no longer serialise behind the kernel.
      const bool tchunk = c.precision == GSTK_PREC_BF16 && bf16_fast_a(c) && T >= 256 && r == 1 && a->out_mel && !is_device_ptr(a->out_mel) &&
                          !getenv("GSTK_NO_TCHUNK");
      if (tchunk) {
        if (!h->st_copy) {
          CK(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
          CK(cudaEventCreateWithFlags(&h->ev_chunk, cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
        }
        void* lastmel;
        if ((rc = slot_reserve(h, SL_LASTMEL, (size_t)Bc * mel * 4, &lastmel))) return rc;
        // Chunk lengths.  A chunk's outputs go to the host while the NEXT chunk decodes, and the copy is ~5x faster than the decode
        // (0.24 MB against 27 us per step), so the chunks shrink geometrically: what stays exposed is the copy of the LAST chunk, and
        // that one is small (1000 steps: 750 + 188 + 62 -> 0.3 ms of tail instead of 1.2 ms with four equal chunks, one launch less).
        // GSTK_TCHUNKS=n: n equal chunks (A/B measurements).  Every chunk but the last has an even length (state parity, see above).
        std::vector<int> chunk_len;
        if (getenv("GSTK_TCHUNKS")) {
          const int nch = std::max(1, atoi(getenv("GSTK_TCHUNKS")));
          const int Tc = (((T + nch - 1) / nch) + 1) & ~1;
          for (int t0 = 0; t0 < T; t0 += Tc) chunk_len.push_back(std::min(Tc, T - t0));
        } else {
          int left = T;
          for (int i = 0; i < 2 && left >= 128; ++i) {
            const int c = ((left * 3 / 4) + 1) & ~1;
            chunk_len.push_back(c);
            left -= c;
          }
          if (left > 0) chunk_len.push_back(left);
        }
        int t0 = 0;
        for (size_t ci = 0; ci < chunk_len.size(); t0 += chunk_len[ci], ++ci) {
          const int Tn = chunk_len[ci];
          DecParams pc = p;
          pc.T = Tn;
          pc.t_base = t0;
          launched_end = t0 + Tn;
          pc.step_offset = p.step_offset + (unsigned int)t0;
          if (pc.out_mel) pc.out_mel += (size_t)t0 * mel;
          if (pc.out_stop) pc.out_stop += t0;
          if (pc.out_align) pc.out_align += (size_t)t0 * Tv;
          if (pc.teacher) pc.teacher += (size_t)t0 * ts_t;
          if (pc.keep0) pc.keep0 += (size_t)t0 * B * c.prenet0;
          if (pc.keep1) pc.keep1 += (size_t)t0 * B * c.prenet1;
          if (pc.noise) pc.noise += (size_t)t0 * B * Tv;
          if (t0 > 0) {
            // free-running hand-over: the next decoder input is the last frame of the previous chunk (Taco2.py:183-187)
            CK(cudaMemcpy2DAsync(lastmel, (size_t)mel * 4, p.out_mel + (size_t)(t0 - 1) * mel, (size_t)T * mel * 4, (size_t)mel * 4, Bc,
                                 cudaMemcpyDeviceToDevice, st));
            pc.init_mel = (const float*)lastmel;
            if ((rc = reset_barrier(h, st))) return rc;
          }
          cudaEvent_t e0 = first_launch ? h->ev0 : h->ev2;
          trace(h, "decode: launching chunk");
          rc = run_bf16_decoder(h, pc, a->kernel, st, e0);
          trace(h, "decode: chunk launched");
          if (rc) return rc;
          first_launch = false;
          CK(cudaEventRecord(h->ev_chunk, st));
          CK(cudaStreamWaitEvent(h->st_copy, h->ev_chunk, 0));
          auto d2h = [&](void* host, const float* dev, size_t row) -> cudaError_t {   // [Bc rows of T * row floats], columns [t0, t0 + Tn)
            if (!host || is_device_ptr(host)) return cudaSuccess;
            return cudaMemcpy2DAsync((float*)host + ((size_t)b0 * T + t0) * row, (size_t)T * row * 4, dev + (size_t)t0 * row, (size_t)T * row * 4,
                                     (size_t)Tn * row * 4, Bc, cudaMemcpyDeviceToHost, h->st_copy);
          };
          CK(d2h(a->out_mel, p.out_mel, mel));
          CK(d2h(a->out_stop, p.out_stop, 1));
          CK(d2h(a->out_alignment, p.out_align, Tv));
          if (p.early_stop) {   // every utterance of the chunk stopped: the remaining time chunks are not launched
            unsigned int ss[2];
            CK(cudaMemcpyAsync(ss, stop_state, sizeof(ss), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (ss[0] >= (unsigned int)Bc) break;
          }
        }
        CK(cudaEventRecord(h->ev_copied, h->st_copy));
        CK(cudaStreamWaitEvent(st, h->ev_copied, 0));
        h->ev_valid = true;
        h->ev_stream = st;
      } else if (c.precision == GSTK_PREC_BF16) {
        cudaEvent_t e0 = first_launch ? h->ev0 : h->ev2;
        rc = run_bf16_decoder(h, p, a->kernel, st, e0);
        if (rc) return rc;
        h->ev_valid = true;
        h->ev_stream = st;
      } else {
        if (Bc >= 16) rc = launch_decoder_fp32<16>(h, p, st);
        else if (Bc > 4) rc = launch_decoder_fp32<8>(h, p, st);
        else if (Bc > 2) rc = launch_decoder_fp32<4>(h, p, st);
        else if (Bc > 1) rc = launch_decoder_fp32<2>(h, p, st);
        else rc = launch_decoder_fp32<1>(h, p, st);
        if (rc) return rc;
      }
      first_launch = false;
    }
    if (p.early_stop && T > 0) {
      unsigned int ss[2];
      CK(cudaMemcpyAsync(ss, stop_state, sizeof(ss), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      steps_done_max = std::max(steps_done_max, (int)std::min(ss[1], (unsigned int)launched_end));
    }
    // ---- final state out
    const int last = T > 0 ? ((T - 1) & 1) : 1;
    if (o_states) {
      float* s = (float*)o_states;
      CK(cudaMemcpyAsync(s + (size_t)b0 * U0, (float*)h1 + (size_t)last * Bc * U0, (size_t)Bc * U0 * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(s + (size_t)B * U0 + (size_t)b0 * U0, c1, (size_t)Bc * U0 * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(s + (size_t)2 * B * U0 + (size_t)b0 * U1, (float*)h2 + (size_t)last * Bc * U1, (size_t)Bc * U1 * 4,
                         cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(s + (size_t)2 * B * U0 + (size_t)B * U1 + (size_t)b0 * U1, c2, (size_t)Bc * U1 * 4,
                         cudaMemcpyDeviceToDevice, st));
    }
    if (o_cum) CK(cudaMemcpyAsync((float*)o_cum + (size_t)b0 * Tv, cum, (size_t)Bc * Tv * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (want_stop && T > 0) {
    if (p.early_stop) steps_done = steps_done_max;
    stop_index_finish_kernel<<<(B + 255) / 256, 256, 0, st>>>((int*)stop_idx, B, T);
    h->launches++;
    if (a->out_stop_index)
      CK(cudaMemcpyAsync(a->out_stop_index, stop_idx, (size_t)B * 4, is_device_ptr(a->out_stop_index) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (a->out_steps_done) {
      if (is_device_ptr(a->out_steps_done)) CK(cudaMemcpyAsync(a->out_steps_done, &steps_done, 4, cudaMemcpyHostToDevice, st));
      else *a->out_steps_done = steps_done;
    }
    if (a->out_stop_index && !is_device_ptr(a->out_stop_index)) CK(cudaStreamSynchronize(st));
  } else {
    if (a->out_steps_done && !is_device_ptr(a->out_steps_done)) *a->out_steps_done = T;
  }
  if (T >= 256 && c.precision == GSTK_PREC_BF16 && bf16_fast_a(c) && r == 1 && a->out_mel && !is_device_ptr(a->out_mel) && !getenv("GSTK_NO_TCHUNK")) {
    // these went out chunk by chunk on the copy stream
    auto& pd = h->pending;
    pd.erase(std::remove_if(pd.begin(), pd.end(), [&](const PendingCopy& cpy) {
               return cpy.dst == a->out_mel || cpy.dst == a->out_stop || cpy.dst == a->out_alignment; }), pd.end());
    if (pd.empty()) {
      pd.push_back({nullptr, nullptr, 0});   // nothing left to copy: flush_pending still ends the call (slot events, synchronise)
    }
  }
  return flush_pending(h, st, true);
}

int gstk_gst(GstkHandle* h, const GstkGstArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  trace(h, "gst: enter");
  const GstkConfig& c = h->cfg;
  DEVICE_GUARD(h, c.device);
  int rc = prepare_gst(h);
  if (rc) return rc;
  const int B = a->batch, mel = c.mel_dim;
  const int H0 = a->frames - (a->drop_first ? 1 : 0);
  if (B < 1 || H0 < 1) return fail(h, GSTK_EINVAL, "batch and frames must be positive");
  if (!a->mels || !a->lengths) return fail(h, GSTK_EINVAL, "mels and lengths are required");
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const void *mels, *lengths;
  if ((rc = stage_in(h, SL_MELS, a->mels, (size_t)B * a->frames * mel * 4, st, &mels))) return rc;
  if ((rc = stage_in(h, SL_LENGTHS, a->lengths, (size_t)B * 4, st, &lengths))) return rc;
  void *o_gst, *o_ref, *o_att;
  if ((rc = stage_out(h, SL_OUT_GST, a->out_gst, (size_t)B * c.style_size * 4, &o_gst))) return rc;
  if ((rc = stage_out(h, SL_OUT_REF, a->out_ref, (size_t)B * c.ref_dense * 4, &o_ref))) return rc;
  if ((rc = stage_out(h, SL_OUT_ATT, a->out_attention, (size_t)B * c.n_tokens * 4, &o_att))) return rc;
  // activation ping-pong buffers
  size_t max_act = 0;
  {
    int H = H0, W = mel;
    for (int i = 0; i < c.ref_layers; ++i) {
      H = (H + 1) / 2;
      W = (W + 1) / 2;
      max_act = std::max(max_act, (size_t)B * H * W * c.ref_filters[i] * 4);
    }
  }
  void *act0 = nullptr, *act1 = nullptr;   // fp32 ping-pong activations of the FFMA kernels (reserved below, only if they run)
  const std::string r = std::string(GSTP) + "/Reference_Encoder";
  CK(cudaEventRecord(h->ev0, st));
  const float* in = (const float*)mels + (a->drop_first ? mel : 0);  // mels[:, 1:] (GST.py:98)
  long long in_bs = (long long)a->frames * mel;
  int H = H0, W = mel, cin = 1;
  // Tensor-core mode (handle precision bf16): layer 0 direct, layers >= 1 as implicit GEMMs on tcgen05 (gst_tc.cuh), where the
  // channel counts allow it (input channels % 16, filters % 32, <= 256).  GSTK_GST_TC=0 keeps the fp32 FFMA kernels.
  int gru_bn = 0;   // > 0: accumulator width of the tcgen05 GEMM that computes the GRU input projections
  bool conv_tc = c.precision == GSTK_PREC_BF16 && c.ref_layers >= 2 && c.ref_filters[0] % 32 == 0 && c.ref_filters[0] <= 256 &&
                 !(getenv("GSTK_GST_TC") && atoi(getenv("GSTK_GST_TC")) == 0);
  for (int i = 1; i < c.ref_layers && conv_tc; ++i) conv_tc = c.ref_filters[i] % 32 == 0 && c.ref_filters[i] <= 256 && c.ref_filters[i - 1] % 16 == 0;
  if (conv_tc) {
    int rc2;
    // block matrices of layers 1..L-1 (fp16, ping-pong) and the fp32 NHWC output of the last layer
    std::vector<GstGeom> gh(c.ref_layers), gw(c.ref_layers);
    size_t max_blk = 0;
    {
      int hh = H0, ww = mel;
      for (int i = 0; i < c.ref_layers; ++i) {
        gh[i] = gst_geom(hh); gw[i] = gst_geom(ww);
        hh = gh[i].out; ww = gw[i].out;
        if (i >= 1) max_blk = std::max(max_blk, (size_t)B * (gh[i].out + 1) * (gw[i].out + 1) * 4 * c.ref_filters[i - 1] * 2);
      }
    }
    void *blk0, *blk1, *outf;
    if ((rc2 = slot_reserve(h, SL_GST_BLK0, max_blk + 1024, &blk0))) return rc2;
    if ((rc2 = slot_reserve(h, SL_GST_BLK1, max_blk + 1024, &blk1))) return rc2;
    const int L = c.ref_layers, HoL = gh[L - 1].out, WoL = gw[L - 1].out, coL = c.ref_filters[L - 1];
    if ((rc2 = slot_reserve(h, SL_ACT0, (size_t)B * HoL * WoL * coL * 4, &outf))) return rc2;
    // GRU input projection on the tensor cores too when its shapes fit one or two accumulator tiles: the last conv layer then
    // leaves fp16 rows [B * T'][W' * C] for it
    const int ginL = WoL * coL, G3 = 3 * c.ref_gru;
    gru_bn = ginL % 64 == 0 && G3 % 16 == 0 ? (G3 <= 256 ? G3 : ((G3 / 2) % 16 == 0 && G3 / 2 <= 256 ? G3 / 2 : 0)) : 0;
    auto zero_border = [&](void* y, int i) -> int {   // block matrix that feeds layer i
      const int Hb = gh[i].out + 1, Wb = gw[i].out + 1, C = c.ref_filters[i - 1];
      const long long n = (long long)B * ((Hb + Wb) * 4 + (gh[i].shift ? Wb * 2 : 0) + (gw[i].shift ? Hb * 2 : 0)) * (C / 8);
      gst_border_zero_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 16), 256, 0, st>>>((__half*)y, B, Hb, Wb, C, gh[i].shift, gw[i].shift);
      h->launches++;
      CK(cudaGetLastError());
      return GSTK_OK;
    };
    // layer 0: mel -> block matrix of layer 1
    {
      const int co = c.ref_filters[0];
      if ((rc2 = zero_border(blk0, 1))) return rc2;
      const int strips = (gh[0].out + G0_HO - 1) / G0_HO, mtiles = (gw[0].out + 15) / 16;
      const size_t smem0 = (size_t)(2 * G0_HO + 1) * (32 * mtiles + 2) * 2;
      const uint32_t* w0p = (const uint32_t*)h->derived["gst_w0p"].p;
      const float* sh0 = dd(h, "conv_shift0");
      __half* y0 = (__half*)blk0;
#define GST_CONV0(NG) gst_conv0_mma_kernel<NG><<<(unsigned)(B * strips), 256, smem0, st>>>(in, in_bs, w0p, sh0, y0, B, H0, mel, gh[0].out, gw[0].out, \
          gh[0].pad, gw[0].pad, \
          gh[1].out + 1, gw[1].out + 1, gh[1].shift, gw[1].shift)
      if (co == 32) { GST_CONV0(1); } else if (co == 64) { GST_CONV0(2); } else if (co == 128) { GST_CONV0(4); } else { GST_CONV0(8); }
#undef GST_CONV0
      h->launches++;
      CK(cudaGetLastError());
    }
    CK(cudaFuncSetAttribute(postnet_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
    for (int i = 1; i < L; ++i) {
      void* xin_blk = (i & 1) ? blk0 : blk1;
      void* yout_blk = (i & 1) ? blk1 : blk0;
      const int ci = c.ref_filters[i - 1], co = c.ref_filters[i];
      const int Hb = gh[i].out + 1, Wb = gw[i].out + 1;
      PostTcParams q;
      memset(&q, 0, sizeof(q));
      q.p.X = xin_blk;
      q.p.shift = dd(h, "conv_shift" + std::to_string(i));
      q.p.Mtotal = (long long)B * Hb * Wb;
      q.p.C = 4 * ci; q.p.K = 16 * ci; q.p.N = co;
      q.p.use_tanh = 2;
      q.BN = co;
      q.tiles_n = 1;
      q.tiles_m = (int)((q.p.Mtotal + PC_BM - 1) / PC_BM);
      q.cpb = 4 * ci / 64;
      q.KB = 4 * q.cpb;
      q.ntap = 4;
      q.toff[0] = 0; q.toff[1] = 1; q.toff[2] = Wb; q.toff[3] = Wb + 1;
      q.Hb = Hb; q.Wb = Wb; q.Ho = gh[i].out; q.Wo = gw[i].out;
      if (i + 1 < L) {
        q.gmode = 1;
        q.p.Y = yout_blk;
        q.nHb = gh[i + 1].out + 1; q.nWb = gw[i + 1].out + 1; q.nsh = gh[i + 1].shift; q.nsw = gw[i + 1].shift;
        if ((rc2 = zero_border(yout_blk, i + 1))) return rc2;
      } else if (gru_bn) {
        q.gmode = 3;
        q.p.Y = outf;
      } else {
        q.gmode = 2;
        q.p.out = (float*)outf;
      }
      CUtensorMap tmA, tmB;
      // one A box of 128 + Wb + 1 rows per 64-channel block for all four taps when box + the taps' B tiles fit a ring stage
      // (layer 1 at the defaults, the memory-bound one: a quarter of the shared-memory fill); GSTK_GST_SHARED_BOX=0 for A/B runs
      static const bool shared_on = !(getenv("GSTK_GST_SHARED_BOX") && atoi(getenv("GSTK_GST_SHARED_BOX")) == 0);
      const int srows = PC_BM + Wb + 1;
      if (shared_on && srows <= 256 && (((size_t)srows * 128 + 1023) & ~(size_t)1023) + (size_t)4 * co * 128 <= (size_t)PT_STAGE_BYTES)
        q.shared_rows = srows;
      // rows beyond the matrix (taps of the last tile) are zero-filled by TMA
      if ((rc2 = encode_tmap_f16(h, &tmA, xin_blk, (uint64_t)4 * ci, (uint64_t)q.p.Mtotal, (uint64_t)4 * ci * 2,
                                 q.shared_rows ? (uint32_t)q.shared_rows : (uint32_t)PC_BM))) return rc2;
      if ((rc2 = encode_tmap_f16(h, &tmB, h->derived["gst_wt" + std::to_string(i)].p, (uint64_t)16 * ci, (uint64_t)co, (uint64_t)16 * ci * 2, (uint32_t)co))) return rc2;
      postnet_conv_tc_kernel<<<std::min(h->num_sms, q.tiles_m), PT_THREADS, PT_SMEM, st>>>(tmA, tmB, q);
      h->launches++;
      CK(cudaGetLastError());
    }
    in = (const float*)outf;
    H = HoL; W = WoL; cin = coL;
    in_bs = (long long)H * W * cin;
  }
  if (!conv_tc) {
    if ((rc = slot_reserve(h, SL_ACT0, max_act, &act0))) return rc;
    if ((rc = slot_reserve(h, SL_ACT1, max_act, &act1))) return rc;
  }
  for (int i = 0; i < c.ref_layers && !conv_tc; ++i) {
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, co = c.ref_filters[i];
    float* out = (float*)((i & 1) ? act1 : act0);
    const size_t smem = (size_t)(2 * CV_HT + 1) * (W + 2) * cin * 4;
    if (smem > 200 * 1024) return fail(h, GSTK_EINVAL, "reference-encoder conv patch too large");
    CK(cudaFuncSetAttribute(conv3x3s2_bn_relu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((Ho + CV_HT - 1) / CV_HT, B);
    // register-tiled kernel wherever the channel counts allow 16 B accesses (all layers but the first, and batch strides that
    // keep the rows 16 B aligned); the scalar kernel otherwise
    const float* cw = dw(h, r + "/Conv2D_" + std::to_string(i) + "/conv2d/kernel");
    const float *csc = dd(h, "conv_scale" + std::to_string(i)), *csh = dd(h, "conv_shift" + std::to_string(i));
    const bool v2 = cin % 4 == 0 && co % 4 == 0 && in_bs % 4 == 0 &&
                    (((uintptr_t)in | (uintptr_t)out | (uintptr_t)cw | (uintptr_t)csc | (uintptr_t)csh) & 15) == 0;
    if (v2) {
      CK(cudaFuncSetAttribute(conv3x3s2_bn_relu_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      conv3x3s2_bn_relu_v2_kernel<<<grid, CV_THREADS, smem, st>>>(
          in, in_bs, dw(h, r + "/Conv2D_" + std::to_string(i) + "/conv2d/kernel"),
          dd(h, "conv_scale" + std::to_string(i)), dd(h, "conv_shift" + std::to_string(i)), out, H, W, cin, Ho, Wo, co);
    } else
    conv3x3s2_bn_relu_kernel<<<grid, CV_THREADS, smem, st>>>(
        in, in_bs, dw(h, r + "/Conv2D_" + std::to_string(i) + "/conv2d/kernel"),
        dd(h, "conv_scale" + std::to_string(i)), dd(h, "conv_shift" + std::to_string(i)), out, H, W, cin, Ho, Wo, co);
    h->launches++;
    CK(cudaGetLastError());
    in = out;
    H = Ho; W = Wo; cin = co;
    in_bs = (long long)H * W * cin;
  }
  // GRU input projections xs = x.W + b[0] for all (b, t)
  const int G = c.ref_gru, gin = W * cin, Tp = H;
  void* xs;
  if ((rc = slot_reserve(h, SL_XS, (size_t)B * Tp * 3 * G * 4, &xs))) return rc;
  const float* gb = dw(h, r + "/RNN/bias");
  if (gru_bn) {
    PostTcParams q;
    memset(&q, 0, sizeof(q));
    q.p.X = in;                       // fp16 [B * Tp][gin] from the last conv layer
    q.p.shift = gb;                   // bias[0]
    q.p.out = (float*)xs;
    q.p.Mtotal = (long long)B * Tp;
    q.p.C = gin; q.p.K = gin; q.p.N = 3 * G;
    q.p.R = 1; q.p.PADL = 0; q.p.T = 1;   // every row is a valid output row
    q.BN = gru_bn;
    q.tiles_n = 3 * G / gru_bn;
    q.tiles_m = (int)((q.p.Mtotal + PC_BM - 1) / PC_BM);
    q.cpb = gin / 64;
    q.KB = q.cpb;
    CUtensorMap tmA, tmB;
    if ((rc = encode_tmap_f16(h, &tmA, in, (uint64_t)gin, (uint64_t)q.p.Mtotal, (uint64_t)gin * 2, PC_BM))) return rc;
    if ((rc = encode_tmap_f16(h, &tmB, h->derived["gst_rnn_wt"].p, (uint64_t)gin, (uint64_t)3 * G, (uint64_t)gin * 2, (uint32_t)gru_bn))) return rc;
    postnet_conv_tc_kernel<<<std::min(h->num_sms, q.tiles_m * q.tiles_n), PT_THREADS, PT_SMEM, st>>>(tmA, tmB, q);
    h->launches++;
    CK(cudaGetLastError());
  } else if ((rc = launch_sgemm(h, in, gin, dw(h, r + "/RNN/kernel"), gb, nullptr, 1, (float*)xs, B * Tp, 3 * G, gin, st))) return rc;
  GruMhaParams gp;
  gp.xs = (const float*)xs; gp.U = dw(h, r + "/RNN/recurrent_kernel"); gp.b_rec = gb + 3 * G;
  gp.Wd = dw(h, r + "/Dense/kernel"); gp.bd = dw(h, r + "/Dense/bias");
  const std::string at = std::string(GSTP) + "/Attention";
  gp.Wq = dw(h, at + "/Query/kernel"); gp.bq = dw(h, at + "/Query/bias");
  gp.tokkv = dd(h, "tokkv");
  gp.ln_g = dw(h, at + "/Layer_Normalization/gamma"); gp.ln_b = dw(h, at + "/Layer_Normalization/beta");
  gp.lengths = (const int*)lengths;
  gp.out_gst = (float*)o_gst; gp.out_ref = (float*)o_ref; gp.out_att = (float*)o_att;
  gp.B = B; gp.Tp = Tp; gp.G = G; gp.D = c.ref_dense; gp.S = c.style_size; gp.NT = c.n_tokens; gp.heads = c.style_heads;
  gp.compress = 1;
  for (int i = 0; i < c.ref_layers; ++i) gp.compress *= c.ref_stride[i];
  const size_t gsm = gru_mha_smem_bytes(G, c.ref_dense, c.style_size, c.n_tokens, c.style_heads);
  CK(cudaFuncSetAttribute(gru_dense_mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
  if (conv_tc && G == GM_G && c.n_tokens * c.style_heads <= 1024) {
    // tensor-core mode: recurrence on mma.sync with the recurrent kernel in registers, up to GM_NU utterances per CTA
    const int upc = std::min(GM_NU, std::max(1, (B + h->num_sms - 1) / h->num_sms));
    const size_t msm = gru_mma_smem_bytes(c.ref_dense, c.style_size, c.n_tokens, c.style_heads);
    CK(cudaFuncSetAttribute(gru_mma_dense_mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
    gru_mma_dense_mha_kernel<<<(B + upc - 1) / upc, 256, msm, st>>>(gp, (const uint4*)h->derived["gst_gru_up"].p, upc);
  } else {
    // up to GRU_UPC utterances per CTA once the batch exceeds one wave of CTAs (the 196 KB recurrent kernel allows one CTA per SM)
    const int upc = std::min(GRU_UPC, std::max(1, (B + h->num_sms - 1) / h->num_sms));
    gru_dense_mha_kernel<<<(B + upc - 1) / upc, 384, gsm, st>>>(gp, upc);
  }
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  h->ev_stream = st;
  return flush_pending(h, st, false);
}

int gstk_postnet(GstkHandle* h, const GstkPostnetArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  const GstkConfig& c = h->cfg;
  DEVICE_GUARD(h, c.device);
  const int B = a->batch, T = a->frames, mel = c.mel_dim, L = a->n_layers;
  if (B < 1 || T < 1) return fail(h, GSTK_EINVAL, "batch and frames must be positive");
  if (!a->decodings || !a->out_post) return fail(h, GSTK_EINVAL, "decodings and out_post are required");
  if (L < 1 || L > 8) return fail(h, GSTK_EINVAL, "1..8 Postnet layers supported");
  const bool bf16 = c.precision == GSTK_PREC_BF16;   // tensor-core mode: fp16 operands for the Postnet (postnet.cuh)
  const int align = bf16 ? 8 : 4;   // 16-byte cp.async chunks
  int cmax = mel, padl = 0, padh = 0;
  if (mel % align) return fail(h, GSTK_EINVAL, "Postnet: Mel_Dim must be a multiple of %d", align);
  for (int i = 0; i < L; ++i) {
    if (a->kernel[i] < 1 || a->filters[i] < 1 || a->filters[i] % align)
      return fail(h, GSTK_EINVAL, "Postnet layer %d: filters must be a positive multiple of %d", i, align);
    cmax = std::max(cmax, a->filters[i]);
    padl = std::max(padl, (a->kernel[i] - 1) / 2);
    padh = std::max(padh, a->kernel[i] - 1 - (a->kernel[i] - 1) / 2);
  }
  // the residual add of Taco2.py:230 needs the last layer to produce Mel_Dim channels
  if (a->filters[L - 1] != mel) return fail(h, GSTK_EINVAL, "Postnet: the last layer must have Mel_Dim filters");
  int rc = prepare_postnet(h, a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const size_t io_bytes = (size_t)B * T * mel * 4;
  const void* dec;
  void* o_post;
  if ((rc = stage_in(h, SL_POST_IN, a->decodings, io_bytes, st, &dec))) return rc;
  if ((rc = stage_out(h, SL_POST_OUT, a->out_post, io_bytes, &o_post))) return rc;
  const int R = padl + T + padh;
  const long long Mtotal = (long long)B * R;
  const size_t elt = bf16 ? 2 : 4;
  // slack: padl rows before row 0 (first tile reads from g - pad_lo), one CTA tile + padh rows after the last row
  const size_t rows_alloc = (size_t)padl + (size_t)Mtotal + PC_BM + padh + 8;
  void* buf[2];
  if ((rc = slot_reserve(h, SL_POST_A, rows_alloc * cmax * elt, &buf[0]))) return rc;
  if ((rc = slot_reserve(h, SL_POST_B, rows_alloc * cmax * elt, &buf[1]))) return rc;
  auto row0 = [&](int which, int ch) { return (void*)((char*)buf[which] + (size_t)padl * ch * elt); };
  CK(cudaEventRecord(h->ev0, st));
  {
    const long long n4 = Mtotal * (mel / 4);
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16);
    if (bf16) postnet_pad_kernel<__half><<<blocks, 256, 0, st>>>((const float*)dec, (__half*)row0(0, mel), Mtotal, mel, R, padl, T);
    else postnet_pad_kernel<float><<<blocks, 256, 0, st>>>((const float*)dec, (float*)row0(0, mel), Mtotal, mel, R, padl, T);
    h->launches++;
    CK(cudaGetLastError());
  }
  int cin = mel;
  for (int i = 0; i < L; ++i) {
    const int co = a->filters[i], k = a->kernel[i];
    PostConvParams p{};
    p.X = row0(i & 1, cin);
    p.W = h->derived["post_w" + std::to_string(i)].p;
    p.shift = dd(h, "post_shift" + std::to_string(i));
    const bool last = i == L - 1;
    p.Y = last ? nullptr : row0((i + 1) & 1, co);
    p.resid = (const float*)dec;
    p.out = (float*)o_post;
    p.Mtotal = Mtotal;
    p.C = cin; p.K = k * cin; p.N = co;
    p.pad_lo = (k - 1) / 2;
    p.R = R; p.PADL = padl; p.T = T;
    p.use_tanh = a->use_tanh[i] ? 1 : 0;
    if ((rc = launch_conv_layer(h, p, bf16, k, padh, h->derived["post_wt" + std::to_string(i)].p, st))) return rc;
    cin = co;
  }
  CK(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  h->ev_stream = st;
  return flush_pending(h, st, false);
}

int gstk_encoder(GstkHandle* h, const GstkEncoderArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  const GstkConfig& c = h->cfg;
  DEVICE_GUARD(h, c.device);
  const int B = a->batch, T = a->key_time, E = a->embedding, L = a->n_layers, u = a->rnn_size;
  if (B < 1 || T < 1) return fail(h, GSTK_EINVAL, "batch and key_time must be positive");
  if (!a->tokens || !a->out) return fail(h, GSTK_EINVAL, "tokens and out are required");
  if (L < 0 || L > 8) return fail(h, GSTK_EINVAL, "0..8 Encoder conv layers supported");
  if (a->vocab < 1) return fail(h, GSTK_EINVAL, "bad vocabulary size");
  if (u < 32 || u > 1024 || u % 32) return fail(h, GSTK_EINVAL, "Encoder RNN size must be a multiple of 32 in [32,1024]");
  const bool bf16 = c.precision == GSTK_PREC_BF16;   // tensor-core mode: fp16 operands, as for the Postnet
  const int align = bf16 ? 16 : 4;
  if (E < 1 || E % align) return fail(h, GSTK_EINVAL, "Encoder: embedding size must be a multiple of %d", align);
  int cmax = E, padl = 0, padh = 0;
  for (int i = 0; i < L; ++i) {
    if (a->kernel[i] < 1 || a->filters[i] < 1 || a->filters[i] % align)
      return fail(h, GSTK_EINVAL, "Encoder conv layer %d: filters must be a positive multiple of %d", i, align);
    cmax = std::max(cmax, a->filters[i]);
    padl = std::max(padl, (a->kernel[i] - 1) / 2);
    padh = std::max(padh, a->kernel[i] - 1 - (a->kernel[i] - 1) / 2);
  }
  int rc = prepare_encoder(h, a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const void* tok;
  void* o_enc;
  if ((rc = stage_in(h, SL_ENC_TOK, a->tokens, (size_t)B * T * 4, st, &tok))) return rc;
  if ((rc = stage_out(h, SL_ENC_OUT, a->out, (size_t)B * T * 2 * u * 4, &o_enc))) return rc;
  const int R = padl + T + padh;
  const long long Mtotal = (long long)B * R;
  const size_t elt = bf16 ? 2 : 4;
  const size_t rows_alloc = (size_t)padl + (size_t)Mtotal + PC_BM + padh + 8;
  void *buf[2], *xs;
  if ((rc = slot_reserve(h, SL_POST_A, rows_alloc * cmax * elt, &buf[0]))) return rc;
  if ((rc = slot_reserve(h, SL_POST_B, rows_alloc * cmax * elt, &buf[1]))) return rc;
  // gate pre-activations of the recurrence: fp32 [B][T][8u], or - for the cluster kernel - fp16 rows of the padded matrix
  const bool x16 = bilstm_cluster(bf16, u);
  if ((rc = slot_reserve(h, SL_ENC_XS, x16 ? (size_t)Mtotal * 8 * u * 2 : (size_t)B * T * 8 * u * 4, &xs))) return rc;
  auto row0 = [&](int which, int ch) { return (void*)((char*)buf[which] + (size_t)padl * ch * elt); };
  CK(cudaEventRecord(h->ev0, st));
  {
    const long long n4 = Mtotal * (E / 4);
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16);
    const float* table = dw(h, std::string(ENCP) + "/embedding/embeddings");
    if (bf16) encoder_embed_pad_kernel<__half><<<blocks, 256, 0, st>>>((const int*)tok, table, (__half*)row0(0, E), Mtotal, E, R, padl, T, a->vocab);
    else encoder_embed_pad_kernel<float><<<blocks, 256, 0, st>>>((const int*)tok, table, (float*)row0(0, E), Mtotal, E, R, padl, T, a->vocab);
    h->launches++;
    CK(cudaGetLastError());
  }
  int cin = E;
  for (int i = 0; i <= L; ++i) {   // i == L: the LSTM input projections of both directions, a k = 1 layer onto fp32 xs
    const bool proj = i == L;
    const int co = proj ? 8 * u : a->filters[i], k = proj ? 1 : a->kernel[i];
    const std::string tag = proj ? std::string("encx") : "enc" + std::to_string(i);
    PostConvParams p{};
    p.X = row0(i & 1, cin);
    p.W = h->derived[tag + "_w"].p;
    p.shift = dd(h, tag + "_shift");
    p.Y = proj ? (x16 ? xs : nullptr) : row0((i + 1) & 1, co);
    p.resid = nullptr;
    p.out = x16 ? nullptr : (float*)xs;
    p.Mtotal = Mtotal;
    p.C = cin; p.K = k * cin; p.N = co;
    p.pad_lo = (k - 1) / 2;
    p.R = R; p.PADL = padl; p.T = T;
    p.use_tanh = proj ? 0 : 2;   // ReLU (Taco2.py:35)
    if ((rc = launch_conv_layer(h, p, bf16, k, padh, bf16 ? h->derived[tag + "_wt"].p : nullptr, st))) return rc;
    cin = co;
  }
  {
    const std::string base = std::string(ENCP) + "/bidirectional/";
    if ((rc = run_bilstm(h, (const float*)xs, dw(h, base + "forward_lstm/lstm_cell/recurrent_kernel"),
                         dw(h, base + "backward_lstm/lstm_cell/recurrent_kernel"), (float*)o_enc, B, T, u, bf16, st,
                         x16 ? (const __half*)xs : nullptr, R, padl))) return rc;
  }
  CK(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  h->ev_stream = st;
  return flush_pending(h, st, true);
}

int gstk_vocoder(GstkHandle* h, const GstkVocoderArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  const GstkConfig& c = h->cfg;
  DEVICE_GUARD(h, c.device);
  const int B = a->batch, T = a->frames, mel = c.mel_dim, KB = a->bank_count, F = a->bank_filters, CB = KB * F, NP = a->n_proj;
  const int hs = a->highway_size, u = a->rnn_size, NS = a->spectrogram_dim;
  if (B < 1 || T < 1) return fail(h, GSTK_EINVAL, "batch and frames must be positive");
  if (!a->mels || !a->out) return fail(h, GSTK_EINVAL, "mels and out are required");
  if (KB < 1 || KB > 32 || F < 1) return fail(h, GSTK_EINVAL, "Vocoder: bad conv bank");
  if (NP < 1 || NP > 8) return fail(h, GSTK_EINVAL, "Vocoder: 1..8 projection conv layers supported");
  if (a->pool_size < 1 || a->pool_strides != 1)
    return fail(h, GSTK_EINVAL, "Vocoder: MaxPool1D strides must be 1 (other strides break the residual add of Taco2.py:372)");
  if (u < 32 || u > 1024 || u % 32) return fail(h, GSTK_EINVAL, "Vocoder RNN size must be a multiple of 32 in [32,1024]");
  if (NS < 1) return fail(h, GSTK_EINVAL, "bad Spectrogram_Dim");
  const bool bf16 = c.precision == GSTK_PREC_BF16;   // tensor-core mode: fp16 operands, as for the Postnet
  const int align = bf16 ? 16 : 4;
  if (mel % align || mel > VH_MAXC) return fail(h, GSTK_EINVAL, "Vocoder: Mel_Dim must be a multiple of %d and <= %d", align, VH_MAXC);
  if (F % align) return fail(h, GSTK_EINVAL, "Vocoder: Conv_Bank.Filters must be a multiple of %d", align);
  if (hs % align || hs > VH_MAXC) return fail(h, GSTK_EINVAL, "Vocoder: Highwaynet.Size must be a multiple of %d and <= %d", align, VH_MAXC);
  if (a->highway_count < 0 || a->highway_count > VH_MAXL - 2) return fail(h, GSTK_EINVAL, "Vocoder: at most %d Highwaynet layers", VH_MAXL - 2);
  int cmax = std::max(std::max(mel, CB), hs), padl = (KB - 1) / 2, padh = KB - 1 - (KB - 1) / 2;
  for (int i = 0; i < NP; ++i) {
    if (a->proj_kernel[i] < 1 || a->proj_filters[i] < 1 || a->proj_filters[i] % align)
      return fail(h, GSTK_EINVAL, "Vocoder projection conv %d: filters must be a positive multiple of %d", i, align);
    cmax = std::max(cmax, a->proj_filters[i]);
    padl = std::max(padl, (a->proj_kernel[i] - 1) / 2);
    padh = std::max(padh, a->proj_kernel[i] - 1 - (a->proj_kernel[i] - 1) / 2);
  }
  if (a->proj_filters[NP - 1] > VH_MAXC) return fail(h, GSTK_EINVAL, "Vocoder: the last projection conv may have at most %d filters", VH_MAXC);
  const int pool_before = (a->pool_size - 1) / 2;
  padl = std::max(padl, pool_before);
  padh = std::max(padh, a->pool_size - 1 - pool_before);
  int rc = prepare_vocoder(h, a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const void* mels;
  void* o_spec;
  if ((rc = stage_in(h, SL_VOC_IN, a->mels, (size_t)B * T * mel * 4, st, &mels))) return rc;
  if ((rc = stage_out(h, SL_VOC_OUT, a->out, (size_t)B * T * NS * 4, &o_spec))) return rc;
  const int R = padl + T + padh;
  const long long Mtotal = (long long)B * R;
  const size_t elt = bf16 ? 2 : 4;
  const size_t rows_alloc = (size_t)padl + (size_t)Mtotal + PC_BM + padh + 8;
  void *buf[2], *xs, *rnn, *rnn16 = nullptr;
  if ((rc = slot_reserve(h, SL_POST_A, rows_alloc * cmax * elt, &buf[0]))) return rc;
  if ((rc = slot_reserve(h, SL_POST_B, rows_alloc * cmax * elt, &buf[1]))) return rc;
  const bool x16 = bilstm_cluster(bf16, u);   // fp16 gate pre-activations in the padded row layout (see gstk_encoder)
  if ((rc = slot_reserve(h, SL_ENC_XS, x16 ? (size_t)Mtotal * 8 * u * 2 : (size_t)B * T * 8 * u * 4, &xs))) return rc;
  if ((rc = slot_reserve(h, SL_VOC_RNN, ((size_t)B * T + PC_BM) * 2 * u * 4, &rnn))) return rc;
  if (bf16 && (rc = slot_reserve(h, SL_VOC_RNN16, ((size_t)B * T + PC_BM) * 2 * u * 2, &rnn16))) return rc;
  auto row0 = [&](int which, int ch) { return (void*)((char*)buf[which] + (size_t)padl * ch * elt); };
  CK(cudaEventRecord(h->ev0, st));
  {
    const long long n4 = Mtotal * (mel / 4);
    const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16);
    if (bf16) postnet_pad_kernel<__half><<<blocks, 256, 0, st>>>((const float*)mels, (__half*)row0(0, mel), Mtotal, mel, R, padl, T);
    else postnet_pad_kernel<float><<<blocks, 256, 0, st>>>((const float*)mels, (float*)row0(0, mel), Mtotal, mel, R, padl, T);
    h->launches++;
    CK(cudaGetLastError());
  }
  auto conv = [&](const std::string& tag, int src, int cin, int co, int k, int act, void* y, float* out, int ldo, int nvalid, const void* xptr,
                  long long mtot, int r_, int padl_, int padh_) -> int {
    PostConvParams p{};
    p.X = xptr ? xptr : row0(src, cin);
    p.W = h->derived[tag + "_w"].p;
    p.shift = dd(h, tag + "_shift");
    p.Y = y;
    p.out = out;
    p.Mtotal = mtot;
    p.C = cin; p.K = k * cin; p.N = co;
    p.pad_lo = (k - 1) / 2;
    p.R = r_; p.PADL = padl_; p.T = T;
    p.use_tanh = act;
    p.ldo = ldo; p.n_valid = nvalid;
    return launch_conv_layer(h, p, bf16, k, padh_, bf16 ? h->derived[tag + "_wt"].p : nullptr, st);
  };
  // conv bank (ReLU) -> buffer 1
  if ((rc = conv("vocb", 0, mel, CB, KB, 2, row0(1, CB), nullptr, 0, 0, nullptr, Mtotal, R, padl, padh))) return rc;
  {  // max pool -> buffer 0
    const long long n = Mtotal * CB / (16 / (long long)elt);
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)h->num_sms * 32);
    if (bf16) voc_pool_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)row0(1, CB), (__half*)row0(0, CB), Mtotal, CB, R, padl, T, a->pool_size, pool_before);
    else voc_pool_kernel<float><<<blocks, 256, 0, st>>>((const float*)row0(1, CB), (float*)row0(0, CB), Mtotal, CB, R, padl, T, a->pool_size, pool_before);
    h->launches++;
    CK(cudaGetLastError());
  }
  int cur = 0, cin = CB;
  for (int i = 0; i < NP; ++i) {
    const int co = a->proj_filters[i];
    if ((rc = conv("vocp" + std::to_string(i), cur, cin, co, a->proj_kernel[i], i < NP - 1 ? 2 : 0, row0(cur ^ 1, co), nullptr, 0, 0, nullptr,
                   Mtotal, R, padl, padh))) return rc;
    cur ^= 1;
    cin = co;
  }
  {  // Dense(mel) + mels, Dense(size), Highwaynet x count -> the other buffer
    VocHighwayParams q;
    memset(&q, 0, sizeof(q));
    q.X = row0(cur, cin);
    q.Y = row0(cur ^ 1, hs);
    q.resid = (const float*)mels;
    q.Mtotal = Mtotal;
    q.R = R; q.PADL = padl; q.T = T;
    q.C0 = cin;
    int l = 0;
    auto add = [&](int type, int n) {
      q.type[l] = type;
      q.N[l] = n;
      q.W[l] = dd(h, "voch_w" + std::to_string(l));
      q.b[l] = dd(h, "voch_b" + std::to_string(l));
      ++l;
    };
    if (cin != mel) {
      add(0, mel);
      q.resid_mode = 1;   // residual after the Dense
    } else {
      q.resid_mode = 2;   // residual on the input
    }
    if (mel != hs) add(0, hs);
    for (int i = 0; i < a->highway_count; ++i) add(1, hs);
    q.n_layers = l;
    const int blocks = (int)((Mtotal + VH_ROWS - 1) / VH_ROWS);
    // tensor-core mode: the mma.sync form (GSTK_VOC_HIGHWAY=ffma keeps the FFMA kernel on fp16 matrices for A/B measurements)
    const char* force = getenv("GSTK_VOC_HIGHWAY");
    if (bf16 && !(force && !strcmp(force, "ffma"))) {
      VocHighwayMmaParams m;
      memset(&m, 0, sizeof(m));
      m.X = (const __half*)q.X; m.Y = (__half*)q.Y; m.resid = q.resid; m.Mtotal = Mtotal;
      m.R = R; m.PADL = padl; m.T = T; m.C0 = cin; m.n_layers = l; m.resid_mode = q.resid_mode;
      for (int i = 0; i < l; ++i) {
        m.type[i] = q.type[i];
        m.N[i] = q.N[i];
        m.chunks[i] = q.type[i] == 1 ? q.N[i] / 16 : (q.N[i] + 31) / 32;
        m.W[i] = (const uint4*)h->derived["vocm_w" + std::to_string(i)].p;
        m.b[i] = dd(h, "vocm_b" + std::to_string(i));
      }
      CK(cudaFuncSetAttribute(voc_highway_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VM_SMEM));
      voc_highway_mma_kernel<<<(int)((Mtotal + VM_ROWS - 1) / VM_ROWS), VM_THREADS, VM_SMEM, st>>>(m);
    } else if (bf16) {
      CK(cudaFuncSetAttribute(voc_highway_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VH_SMEM));
      voc_highway_kernel<__half><<<blocks, VH_THREADS, VH_SMEM, st>>>(q);
    } else {
      CK(cudaFuncSetAttribute(voc_highway_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VH_SMEM));
      voc_highway_kernel<float><<<blocks, VH_THREADS, VH_SMEM, st>>>(q);
    }
    h->launches++;
    CK(cudaGetLastError());
    cur ^= 1;
  }
  // LSTM input projections of both directions (k = 1, fp32 xs [B][T][8u]) and the recurrence
  if ((rc = conv("vocx", cur, hs, 8 * u, 1, 0, x16 ? xs : nullptr, x16 ? nullptr : (float*)xs, 0, 0, nullptr, Mtotal, R, padl, padh))) return rc;
  {
    const std::string base = std::string(VOCP) + "/CBHG/RNN/";
    if ((rc = run_bilstm(h, (const float*)xs, dw(h, base + "forward_lstm/lstm_cell/recurrent_kernel"),
                         dw(h, base + "backward_lstm/lstm_cell/recurrent_kernel"), (float*)rnn, B, T, u, bf16, st,
                         x16 ? (const __half*)xs : nullptr, R, padl))) return rc;
  }
  // Dense(Spectrogram_Dim) on the plain [B * T][2u] matrix (k = 1: no padding rows), N padded, output row stride Spectrogram_Dim
  {
    const long long rows = (long long)B * T;
    const void* x = rnn;
    if (bf16) {
      const long long n4 = rows * (2 * u) / 4;
      const int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)h->num_sms * 16);
      f32_to_f16_kernel<<<blocks, 256, 0, st>>>((const float*)rnn, (__half*)rnn16, n4);
      h->launches++;
      CK(cudaGetLastError());
      x = rnn16;
    }
    const int Np = (NS + VOC_DENSE_ALIGN - 1) / VOC_DENSE_ALIGN * VOC_DENSE_ALIGN;
    if ((rc = conv("vocd", 0, 2 * u, Np, 1, 0, nullptr, (float*)o_spec, NS, NS, x, rows, T, 0, 0))) return rc;
  }
  CK(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  h->ev_stream = st;
  return flush_pending(h, st, true);
}

int gstk_griffin_lim(GstkHandle* h, const GstkGriffinLimArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  DEVICE_GUARD(h, h->cfg.device);
  const int B = a->batch, T = a->frames, F = a->num_freq, hop = a->hop_length, N = 2 * (F - 1);
  if (B < 1 || T < 2) return fail(h, GSTK_EINVAL, "Griffin-Lim: batch >= 1 and frames >= 2 are required");
  if (!a->spectrogram || !a->out_wav) return fail(h, GSTK_EINVAL, "spectrogram and out_wav are required");
  int log2n = 0;
  while ((1 << log2n) < N) ++log2n;
  if (F < 2 || (1 << log2n) != N || N < 64 || N > 4096) return fail(h, GSTK_EINVAL, "Griffin-Lim: n_fft = 2 (num_freq - 1) must be a power of two in [64, 4096]");
  if (a->win_length != N) return fail(h, GSTK_EINVAL, "Griffin-Lim: win_length must equal n_fft = %d", N);
  if (hop < 2 || N % hop) return fail(h, GSTK_EINVAL, "Griffin-Lim: hop_length must divide n_fft");
  if (a->iters < 0) return fail(h, GSTK_EINVAL, "Griffin-Lim: negative iteration count");
  if (a->rng_mode != GSTK_RNG_EXTERNAL && a->rng_mode != GSTK_RNG_PHILOX) return fail(h, GSTK_EINVAL, "Griffin-Lim: rng_mode must be EXTERNAL or PHILOX");
  if (a->rng_mode == GSTK_RNG_EXTERNAL && !a->init_uniform) return fail(h, GSTK_EINVAL, "rng_mode EXTERNAL needs init_uniform");
  int rc;
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const std::string wkey = "gl_window" + std::to_string(N), tkey = "gl_tw" + std::to_string(N);
  if (!h->derived.count(wkey)) {
    std::vector<float> w(N);
    std::vector<float2> tw(N);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < N; ++i) {
      w[i] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * i / N));   // scipy.signal.get_window('hann', N, fftbins=True)
      tw[i] = make_float2((float)std::cos(2.0 * pi * i / N), (float)-std::sin(2.0 * pi * i / N));
    }
    if ((rc = upload_derived(h, wkey, w.data(), (size_t)N * 4))) return rc;
    if ((rc = upload_derived(h, tkey, tw.data(), (size_t)N * 8))) return rc;
  }
  const int Lmax = hop * (T - 1);
  const size_t nspec = (size_t)B * T * F;
  const void *spec, *lengths = nullptr, *uni = nullptr;
  void *o_wav, *S, *frames, *y;
  if ((rc = stage_in(h, SL_GL_SPEC, a->spectrogram, nspec * 4, st, &spec))) return rc;
  if (a->lengths && (rc = stage_in(h, SL_GL_LEN, a->lengths, (size_t)B * 4, st, &lengths))) return rc;
  if (a->rng_mode == GSTK_RNG_EXTERNAL && (rc = stage_in(h, SL_GL_UNI, a->init_uniform, nspec * 4, st, &uni))) return rc;
  if ((rc = stage_out(h, SL_GL_OUT, a->out_wav, (size_t)B * Lmax * 4, &o_wav))) return rc;
  if ((rc = slot_reserve(h, SL_GL_S, nspec * 4, &S))) return rc;
  if ((rc = slot_reserve(h, SL_GL_FRAMES, (size_t)B * T * N * 4, &frames))) return rc;
  if ((rc = slot_reserve(h, SL_GL_Y, (size_t)B * Lmax * 4, &y))) return rc;
  CK(cudaEventRecord(h->ev0, st));
  {
    const int blocks = (int)std::min<size_t>((nspec + 255) / 256, (size_t)h->num_sms * 16);
    gl_magnitude_kernel<<<blocks, 256, 0, st>>>((const float*)spec, (float*)S, (long long)nspec, a->max_abs_value, a->ref_level_db, a->power);
    h->launches++;
    CK(cudaGetLastError());
  }
  GlParams p;
  memset(&p, 0, sizeof(p));
  p.S = (const float*)S;
  p.frames = (float*)frames;
  p.window = dd(h, wkey);
  p.tw = (const float2*)h->derived[tkey].p;
  p.lengths = (const int*)lengths;
  p.B = B; p.T = T; p.F = F; p.N = N; p.hop = hop; p.Lmax = Lmax; p.log2n = log2n;
  p.row_offset = a->row_offset;
  p.seed = a->seed;
  const size_t smem = gl_frames_smem(N);
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(gl_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 fgrid((T + 1) / 2, B), ogrid(std::min((Lmax / (hop % 4 == 0 ? 4 : 1) + 255) / 256, 4 * h->num_sms), B);
  for (int it = 0; it <= a->iters; ++it) {   // pass 0: the initial phases (Audio.py:61-63); then `iters` rounds (Audio.py:65-67)
    p.init = it == 0;
    p.uniform = it == 0 ? (const float*)uni : nullptr;
    p.y = it == 0 ? nullptr : (const float*)y;
    // n_fft = 1024: radix-16 form (GSTK_GL_RADIX4=1 keeps the generic radix-4 kernel for A/B measurements)
    static const bool r16 = !(getenv("GSTK_GL_RADIX4") && atoi(getenv("GSTK_GL_RADIX4")));
    if (N == 1024 && r16) gl_frames1024_kernel<<<fgrid, GL1K_THREADS, 0, st>>>(p);
    else gl_frames_kernel<<<fgrid, N / 4, smem, st>>>(p);
    if (hop % 4 == 0) gl_overlap_add_kernel<4><<<ogrid, 256, 0, st>>>((const float*)frames, p.window, p.lengths, (float*)y, T, N, hop, Lmax);
    else gl_overlap_add_kernel<1><<<ogrid, 256, 0, st>>>((const float*)frames, p.window, p.lengths, (float*)y, T, N, hop, Lmax);
    h->launches += 2;
  }
  CK(cudaGetLastError());
  gl_deemphasis_kernel<<<B, GLD_THREADS, 0, st>>>((const float*)y, p.lengths, (float*)o_wav, T, hop, Lmax, a->preemphasis);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, st));
  h->ev_valid = true;
  h->ev_stream = st;
  return flush_pending(h, st, false);
}

// ReLU + dropout of one prenet layer in place: keep from the caller's mask, from the decoder's Philox stream, or always (no dropout)
__global__ void prenet_relu_dropout_kernel(float* __restrict__ x, const float* __restrict__ keep, long long rows, int n, int rng_mode, float rate,
                                           float scale, unsigned long long seed, unsigned int stream_id, unsigned int step, int row_offset) {
  const long long total = rows * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int c = (int)(i - r * n);
    float v = fmaxf(x[i], 0.f);
    if (rng_mode != GSTK_RNG_NONE && rate > 0.f) {
      const float k = rng_mode == GSTK_RNG_EXTERNAL ? __ldg(keep + i) : philox_keep(seed, stream_id, step, (unsigned int)(row_offset + r), (unsigned int)c, rate);
      v = k != 0.f ? v * scale : 0.f;
    }
    x[i] = v;
  }
}

int gstk_prenet(GstkHandle* h, const GstkPrenetArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  const GstkConfig& c = h->cfg;
  DEVICE_GUARD(h, c.device);
  int rc = prepare_decoder(h);
  if (rc) return rc;
  const int R = a->rows, mel = c.mel_dim;
  if (R < 1) return fail(h, GSTK_EINVAL, "rows must be positive");
  if (!a->inputs || !a->out) return fail(h, GSTK_EINVAL, "inputs and out are required");
  if (a->rng_mode < GSTK_RNG_NONE || a->rng_mode > GSTK_RNG_PHILOX) return fail(h, GSTK_EINVAL, "bad rng_mode");
  if (a->rng_mode == GSTK_RNG_EXTERNAL && c.prenet_dropout > 0.f && (!a->keep0 || !a->keep1)) return fail(h, GSTK_EINVAL, "rng_mode EXTERNAL needs keep0 / keep1");
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const std::string d = DEC;
  const void *x, *k0 = nullptr, *k1 = nullptr;
  void *mid, *o;
  if ((rc = stage_in(h, SL_PRE_IN, a->inputs, (size_t)R * mel * 4, st, &x))) return rc;
  if (a->rng_mode == GSTK_RNG_EXTERNAL) {
    if ((rc = stage_in(h, SL_PRE_K0, a->keep0, (size_t)R * c.prenet0 * 4, st, &k0))) return rc;
    if ((rc = stage_in(h, SL_PRE_K1, a->keep1, (size_t)R * c.prenet1 * 4, st, &k1))) return rc;
  }
  if ((rc = slot_reserve(h, SL_PRE_MID, (size_t)R * c.prenet0 * 4, &mid))) return rc;
  if ((rc = stage_out(h, SL_PRE_OUT, a->out, (size_t)R * c.prenet1 * 4, &o))) return rc;
  const float scale = c.prenet_dropout > 0.f ? 1.f / (1.f - c.prenet_dropout) : 1.f;
  auto act = [&](void* buf, const void* keep, int n, unsigned int stream_id) {
    const long long total = (long long)R * n;
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8);
    prenet_relu_dropout_kernel<<<blocks, 256, 0, st>>>((float*)buf, (const float*)keep, R, n, a->rng_mode, c.prenet_dropout, scale, a->seed, stream_id,
                                                       a->step, a->row_offset);
    h->launches++;
  };
  if ((rc = launch_sgemm(h, (const float*)x, mel, dw(h, d + "/Prenet/dense/kernel"), dw(h, d + "/Prenet/dense/bias"), nullptr, 1, (float*)mid, R,
                         c.prenet0, mel, st))) return rc;
  act(mid, k0, c.prenet0, STREAM_KEEP0);
  if ((rc = launch_sgemm(h, (const float*)mid, c.prenet0, dw(h, d + "/Prenet/dense_1/kernel"), dw(h, d + "/Prenet/dense_1/bias"), nullptr, 1, (float*)o, R,
                         c.prenet1, c.prenet0, st))) return rc;
  act(o, k1, c.prenet1, STREAM_KEEP1);
  CK(cudaGetLastError());
  return flush_pending(h, st, false);
}

int gstk_mha(GstkHandle* h, const GstkMhaArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  DEVICE_GUARD(h, h->cfg.device);
  if (a->heads < 1 || a->size % a->heads != 0)
    return fail(h, GSTK_EINVAL, "size must be divisible by num_heads. ('%d' %% '%d' != 0)", a->size, a->heads);
  if (a->batch < 1 || a->tq < 1 || a->tv < 1) return fail(h, GSTK_EINVAL, "bad shape");
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  int rc;
  MhaParams p;
  const void* t;
  if ((rc = stage_in(h, SL_MHA0, a->query, (size_t)a->batch * a->tq * a->dq * 4, st, &t))) return rc; p.query = (const float*)t;
  if ((rc = stage_in(h, SL_MHA1, a->value, (size_t)a->batch * a->tv * a->dv * 4, st, &t))) return rc; p.value = (const float*)t;
  if ((rc = stage_in(h, SL_MHA2, a->q_kernel, (size_t)a->dq * a->size * 4, st, &t))) return rc; p.Wq = (const float*)t;
  if ((rc = stage_in(h, SL_MHA3, a->q_bias, (size_t)a->size * 4, st, &t))) return rc; p.bq = (const float*)t;
  if ((rc = stage_in(h, SL_MHA4, a->v_kernel, (size_t)a->dv * a->size * 4, st, &t))) return rc; p.Wv = (const float*)t;
  if ((rc = stage_in(h, SL_MHA5, a->v_bias, (size_t)a->size * 4, st, &t))) return rc; p.bv = (const float*)t;
  if ((rc = stage_in(h, SL_MHA6, a->ln_gamma, (size_t)a->size * 4, st, &t))) return rc; p.ln_g = (const float*)t;
  if ((rc = stage_in(h, SL_MHA7, a->ln_beta, (size_t)a->size * 4, st, &t))) return rc; p.ln_b = (const float*)t;
  void *o, *oa;
  if ((rc = stage_out(h, SL_MHA_OUT, a->out, (size_t)a->batch * a->tq * a->size * 4, &o))) return rc;
  if ((rc = stage_out(h, SL_MHA_ATT, a->out_attention, (size_t)a->batch * a->tq * a->tv * 4, &oa))) return rc;
  if (!o) return fail(h, GSTK_EINVAL, "out is required");
  p.out = (float*)o; p.out_att = (float*)oa;
  p.B = a->batch; p.tq = a->tq; p.tv = a->tv; p.dq = a->dq; p.dv = a->dv; p.S = a->size; p.heads = a->heads;
  const size_t smem = sizeof(float) * ((size_t)a->tv * a->size + 2 * a->size + (size_t)a->heads * a->tv + 4);
  if (smem > 220 * 1024) return fail(h, GSTK_EINVAL, "value sequence too long for the generic attention kernel");
  CK(cudaFuncSetAttribute(mha_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mha_generic_kernel<<<a->batch * a->tq, 256, smem, st>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  return flush_pending(h, st, false);
}

int gstk_attention_step(GstkHandle* h, const GstkAttentionArgs* a) {
  if (!h || !a) return fail(h, GSTK_EINVAL, "null argument");
  DEVICE_GUARD(h, h->cfg.device);
  if (a->type != GSTK_ATT_SMA && a->type != GSTK_ATT_BMA) return fail(h, GSTK_EINVAL, "Unsupported attention type: %d", a->type);
  if (a->batch < 1 || a->key_time < 1 || a->size < 1 || a->key_time > GSTK_MAX_TV) return fail(h, GSTK_EINVAL, "bad shape");
  if (!a->query || !a->value || !a->prev_alignment || !a->q_kernel || !a->q_bias || !a->v_kernel || !a->v_bias || !a->attention_v ||
      !a->attention_score_bias || !a->out_context || !a->out_alignment)
    return fail(h, GSTK_EINVAL, "Unexpected input length");  /* Steps.py:119-120 */
  if (a->key && (!a->k_kernel || !a->k_bias)) return fail(h, GSTK_EINVAL, "4-input form needs the Key Dense variables");
  cudaStream_t st = (cudaStream_t)a->stream;
  h->pending.clear();
  const int B = a->batch, Tv = a->key_time, A = a->size;
  int rc;
  const void *q, *v, *k = nullptr, *prev, *noise, *wq, *bq, *wv, *bv, *wk = nullptr, *bk = nullptr, *av, *sb;
  if ((rc = stage_in(h, SL_AT0, a->query, (size_t)B * a->query_dim * 4, st, &q))) return rc;
  if ((rc = stage_in(h, SL_AT1, a->value, (size_t)B * Tv * a->value_dim * 4, st, &v))) return rc;
  if (a->key && (rc = stage_in(h, SL_AT2, a->key, (size_t)B * Tv * a->key_dim * 4, st, &k))) return rc;
  if ((rc = stage_in(h, SL_AT3, a->prev_alignment, (size_t)B * Tv * 4, st, &prev))) return rc;
  if ((rc = stage_in(h, SL_AT4, a->noise, (size_t)B * Tv * 4, st, &noise))) return rc;
  if ((rc = stage_in(h, SL_AT5, a->q_kernel, (size_t)a->query_dim * A * 4, st, &wq))) return rc;
  if ((rc = stage_in(h, SL_AT6, a->q_bias, (size_t)A * 4, st, &bq))) return rc;
  if ((rc = stage_in(h, SL_AT7, a->v_kernel, (size_t)a->value_dim * A * 4, st, &wv))) return rc;
  if ((rc = stage_in(h, SL_AT8, a->v_bias, (size_t)A * 4, st, &bv))) return rc;
  if (a->key) {
    if ((rc = stage_in(h, SL_AT9, a->k_kernel, (size_t)a->key_dim * A * 4, st, &wk))) return rc;
    if ((rc = stage_in(h, SL_AT10, a->k_bias, (size_t)A * 4, st, &bk))) return rc;
  }
  if ((rc = stage_in(h, SL_AT11, a->attention_v, (size_t)A * 4, st, &av))) return rc;
  float score_bias = 0.f;
  if (is_device_ptr(a->attention_score_bias)) CK(cudaMemcpy(&score_bias, a->attention_score_bias, 4, cudaMemcpyDeviceToHost));
  else score_bias = *a->attention_score_bias;
  (void)sb;
  void *qp, *kp = nullptr, *vp, *octx, *oal;
  if ((rc = slot_reserve(h, SL_AT_Q, (size_t)B * A * 4, &qp))) return rc;
  if ((rc = slot_reserve(h, SL_AT_V, (size_t)B * Tv * A * 4, &vp))) return rc;
  if (a->key && (rc = slot_reserve(h, SL_AT_K, (size_t)B * Tv * A * 4, &kp))) return rc;
  if ((rc = stage_out(h, SL_AT_CTX, a->out_context, (size_t)B * A * 4, &octx))) return rc;
  if ((rc = stage_out(h, SL_AT_AL, a->out_alignment, (size_t)B * Tv * 4, &oal))) return rc;
  if ((rc = launch_sgemm(h, (const float*)q, a->query_dim, (const float*)wq, (const float*)bq, nullptr, 1, (float*)qp, B, A, a->query_dim, st))) return rc;
  if ((rc = launch_sgemm(h, (const float*)v, a->value_dim, (const float*)wv, (const float*)bv, nullptr, 1, (float*)vp, B * Tv, A, a->value_dim, st))) return rc;
  if (a->key && (rc = launch_sgemm(h, (const float*)k, a->key_dim, (const float*)wk, (const float*)bk, nullptr, 1, (float*)kp, B * Tv, A, a->key_dim, st))) return rc;
  AttStepParams p;
  p.q = (const float*)qp; p.value = (const float*)vp; p.key = a->key ? (const float*)kp : (const float*)vp;
  p.prev = (const float*)prev; p.att_v = (const float*)av; p.noise = (const float*)noise;
  p.score_bias = score_bias; p.sigmoid_noise = a->sigmoid_noise;
  p.ctx = (float*)octx; p.align = (float*)oal; p.B = B; p.Tv = Tv; p.A = A; p.type = a->type;
  attention_step_kernel<<<B, 256, (size_t)(2 * Tv + 256) * 4, st>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  return flush_pending(h, st, false);
}

int gstk_concat_encoder(GstkHandle* h, const float* enc_text, const float* gst, float* out, int32_t batch,
                        int32_t key_time, void* stream) {
  if (!h || !enc_text || !gst || !out) return fail(h, GSTK_EINVAL, "null argument");
  const GstkConfig& c = h->cfg;
  if (!c.gst_use) return fail(h, GSTK_ENOTIMPL, "GST is not used");
  DEVICE_GUARD(h, c.device);
  cudaStream_t st = (cudaStream_t)stream;
  h->pending.clear();
  const int S = c.style_size, Dt = c.enc_dim - S;
  int rc;
  const void *e, *g;
  void* o;
  if ((rc = stage_in(h, SL_CAT0, enc_text, (size_t)batch * key_time * Dt * 4, st, &e))) return rc;
  if ((rc = stage_in(h, SL_CAT1, gst, (size_t)batch * S * 4, st, &g))) return rc;
  if ((rc = stage_out(h, SL_CAT_OUT, out, (size_t)batch * key_time * c.enc_dim * 4, &o))) return rc;
  concat_encoder_kernel<<<h->num_sms * 4, 256, 0, st>>>((const float*)e, (const float*)g, (float*)o, batch, key_time, Dt, S);
  h->launches++;
  CK(cudaGetLastError());
  return flush_pending(h, st, false);
}

int gstk_synchronize(GstkHandle* h, void* stream) {
  if (!h) return GSTK_EINVAL;
  DEVICE_GUARD(h, h->cfg.device);
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return check_barrier_error(h);
}

int64_t gstk_launch_count(GstkHandle* h) { return h ? h->launches : 0; }

float gstk_last_kernel_ms(GstkHandle* h) {
  if (!h || !h->ev_valid) return -1.f;
  cudaSetDevice(h->cfg.device);
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.f;
  return ms;
}

int gstk_get_phase_profile(GstkHandle* h, uint64_t* out, int32_t max_ctas, int32_t* n_ctas) {
  if (!h || !out || !n_ctas) return fail(h, GSTK_EINVAL, "null argument");
  DEVICE_GUARD(h, h->cfg.device);
  const int n = std::min(max_ctas, h->bf16.prof ? h->bf16.prof_ctas : 0);
  *n_ctas = n;
  if (n > 0) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, h->bf16.prof, (size_t)n * PROF_SLOTS * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  return GSTK_OK;
}

int gstk_selftest_umma(GstkHandle* h, const float* A, const float* B, int32_t K, float* D) {
  if (!h || !A || !B || !D) return fail(h, GSTK_EINVAL, "null argument");
  if (K < 64 || K % 64 || K > 512) return fail(h, GSTK_EINVAL, "K must be a multiple of 64 in [64,512]");
  DEVICE_GUARD(h, h->cfg.device);
  const int KB = K / 64;
  std::vector<__nv_bfloat16> a_img((size_t)KB * 128 * 64), b_img((size_t)KB * 32 * 64);
  pack_sw128(A, K, 128, KB, a_img.data());
  pack_sw128(B, K, 32, KB, b_img.data());
  void *da, *db, *dout;
  CK(cudaMalloc(&da, a_img.size() * 2));
  CK(cudaMalloc(&db, b_img.size() * 2));
  CK(cudaMalloc(&dout, 128 * 32 * 4));
  CK(cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)KB * (16384 + 4096) + 1024;
  CK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem>>>((const __nv_bfloat16*)da, (const __nv_bfloat16*)db, KB, (float*)dout);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D, dout, 128 * 32 * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return GSTK_OK;
}

}  // extern "C"
