/*
 * gstk.h - C ABI of libgsttaco.so: the B200 (sm_100a) implementation of GST_Tacotron's
 * inference hot path (Tacotron2 decoder loop + GST front end).
 *
 * The reference has no FFI: its "operator API" for this path is the Keras-layer call contract
 * (paths relative to /root/reference).  Each entry point below names the reference interface it
 * replaces; the Python classes in gst_tacotron_b200/Modules keep those call signatures and call
 * this library through ctypes.  INTEGRATION.md shows the binding a reference maintainer adds.
 *
 * Conventions
 *  - every function returns 0 on success, a GSTK_E* code otherwise; gstk_last_error() gives text.
 *  - all tensors are dense, row-major, float32 unless stated; pointers may be DEVICE pointers
 *    (borrowed for the duration of the call, e.g. from DLPack) or HOST pointers (the library
 *    stages them through its own device workspace; pinned host memory makes that asynchronous).
 *    The library classifies each pointer with cudaPointerGetAttributes.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous with respect to the host when every pointer is a device pointer; outputs are
 *    valid after the stream is synchronised.  With host output pointers the call returns after
 *    the device-to-host copies have completed.
 *  - a handle belongs to one device; calls on one handle must be serialised by the caller.
 *  - there is no CPU fallback: without a CUDA device gstk_create fails with GSTK_ENODEVICE.
 */
#ifndef GSTK_H_
#define GSTK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSTK_VERSION 1

enum {
  GSTK_OK = 0,
  GSTK_EINVAL = 1,      /* bad argument / unsupported configuration (Python shim raises ValueError) */
  GSTK_ENODEVICE = 2,   /* no usable CUDA device */
  GSTK_ECUDA = 3,       /* CUDA runtime error, text in gstk_last_error */
  GSTK_ENOWEIGHTS = 4,  /* a required variable has not been loaded */
  GSTK_ENOTIMPL = 5,    /* e.g. GST entry points on a handle created with gst_use = 0
                           (reference raises NotImplementedError, Model.py:258-259) */
  GSTK_ETIMEOUT = 6     /* the persistent kernel's bounded grid barrier gave up */
};

enum { GSTK_ATT_SMA = 0, GSTK_ATT_BMA = 1, GSTK_ATT_LSA = 2 };   /* Taco2.py:66-75 (+LSA ext.) */
enum { GSTK_PREC_FP32 = 0, GSTK_PREC_BF16 = 1 };
enum { GSTK_RNG_NONE = 0,      /* no dropout, no noise (deterministic debugging mode)         */
       GSTK_RNG_EXTERNAL = 1,  /* caller passes keep masks / noise tensors                     */
       GSTK_RNG_PHILOX = 2 };  /* in-kernel Philox4x32-10, counters (idx/4, step, row, stream) */
enum { GSTK_MODE_FREE = 0,     /* training=False: input = last produced frame (Taco2.py:183-187) */
       GSTK_MODE_TEACHER = 1 };/* training=True : input = mels[:, step]                         */

/* Hyper-parameters the path reads from Hyper_Parameters.json (reference lines 4,13-38,109-132). */
typedef struct GstkConfig {
  int32_t version;          /* GSTK_VERSION */
  int32_t device;           /* CUDA ordinal */
  int32_t mel_dim;          /* Sound.Mel_Dim (80) */
  int32_t step_reduction;   /* Step_Reduction (1) */
  int32_t prenet0, prenet1; /* Tacotron2.Decoder.Prenet.Size (256,256) */
  int32_t attention_size;   /* Tacotron2.Decoder.Attention.Size (128) */
  int32_t attention_type;   /* GSTK_ATT_* */
  int32_t lstm0, lstm1;     /* Tacotron2.Decoder.RNN.Size (1024,1024) */
  int32_t enc_dim;          /* channels of `encodings`: 2*Encoder.RNN.Size + (GST ? style size : 0) */
  int32_t gst_use;          /* GST.Use */
  int32_t ref_layers;       /* len(GST.Reference_Encoder.Conv.Filters) (<= 8) */
  int32_t ref_filters[8];
  int32_t ref_kernel[8];    /* only 3 supported */
  int32_t ref_stride[8];    /* only 2 supported */
  int32_t ref_gru;          /* GST.Reference_Encoder.RNN.Size (128) */
  int32_t ref_dense;        /* GST.Reference_Encoder.Dense.Size (128) */
  int32_t n_tokens;         /* GST.Style_Token.Size (16) */
  int32_t token_dim;        /* GST.Style_Token.Embedding.Size (256) */
  int32_t style_heads;      /* GST.Style_Token.Attention.Head (4) */
  int32_t style_size;       /* GST.Style_Token.Attention.Size (128) */
  int32_t lsa_filters, lsa_kernel, lsa_cumulate, lsa_smoothing; /* step-form LSA (Layers.py:289-320) */
  int32_t precision;        /* GSTK_PREC_* */
  float prenet_dropout;     /* Tacotron2.Decoder.Prenet.Dropout_Rate (0.5), always applied (Taco2.py:283) */
  float sigmoid_noise;      /* 2.0 for SMA (Steps.py:212), 0.0 for BMA (Steps.py:58) */
  int32_t reserved[8];
} GstkConfig;

typedef struct GstkHandle GstkHandle;

/* A named variable; `name` is the canonical path of gst_tacotron_b200/weights.py, e.g.
 * "Decoder/Decoder_Step/RNN/cell_0/recurrent_kernel" (Keras variable layout). */
typedef struct GstkTensorDesc {
  const char* name;
  const float* data;        /* host or device, float32, row-major */
  int32_t ndim;
  int64_t shape[4];
} GstkTensorDesc;

/* Replaces: Decoder.call's tf.while_loop (Taco2.py:153-228) including Decoder_Step.call
 * (Taco2.py:96-120), Prenet (262-283), the step attentions (Steps.py:107-229) and the per-step
 * tf.concat accumulation (203-205).  The Postnet (Taco2.py:230) is a separate call: gstk_postnet. */
typedef struct GstkDecodeArgs {
  int32_t batch;            /* B */
  int32_t key_time;         /* T_v */
  int32_t steps;            /* trip count: shape(mels[:,0:-1:r])[1] (teacher) or Max_Step//r (free) */
  int32_t mode;             /* GSTK_MODE_* */
  int32_t rng_mode;         /* GSTK_RNG_* */
  int32_t pad0;
  uint64_t seed;            /* GSTK_RNG_PHILOX */
  uint32_t step_offset;     /* philox step counter of the first step (continuing a decode) */
  uint32_t row_offset;      /* philox row counter of batch row 0 (utterance-sharded callers) */
  /* inputs */
  const float* encodings;   /* [B,T_v,enc_dim] (GST channels first, GST.py:121-124), or NULL with: */
  const float* enc_text;    /* [B,T_v,enc_dim-style_size]  + */
  const float* gst;         /* [B,style_size]: the concat of GST_Concated_Encoder folded into V' */
  const float* teacher_mels;/* [B,steps,mel_dim]: mels[:,0:-1:r] already sliced; teacher mode only */
  int64_t teacher_stride_b; /* element stride between batch rows of teacher_mels (0 = steps*mel_dim) */
  int64_t teacher_stride_t; /* element stride between steps (0 = mel_dim) */
  const float* keep0;       /* [steps,B,prenet0] in {0,1}; GSTK_RNG_EXTERNAL */
  const float* keep1;       /* [steps,B,prenet1] */
  const float* noise;       /* [steps,B,T_v] ~N(0,1), scaled by sigmoid_noise inside (Steps.py:220-221) */
  /* optional state in (NULL = reference initial state: zeros / initial_alignment_fn) */
  const float* init_mel;        /* [B,mel_dim] last produced frame (free mode) */
  const float* init_alignment;  /* [B,T_v] */
  const float* init_cum_alignment; /* [B,T_v] LSA only */
  const float* init_states;     /* [4,B,lstm] order h1,c1,h2,c2 */
  /* outputs (any may be NULL) */
  float* out_mel;           /* [B,steps*r,mel_dim]  decodings[:,1:] */
  float* out_stop;          /* [B,steps]            stop logits */
  float* out_alignment;     /* [B,steps,T_v]        alignments[:,1:] */
  float* out_states;        /* [4,B,lstm] final h1,c1,h2,c2 */
  float* out_cum_alignment; /* [B,T_v] LSA only */
  float* out_context;       /* [B,attention_size] context of the last step */
  void* stream;
  /* Early stop.  The reference always runs Max_Step//r steps and the CALLER cuts every utterance at
   * np.argmax(stop < 0) (Model.py:380,413).  early_stop = 1 (free-running mode): the loop ends after the first step at which
   * EVERY utterance of the call has produced a negative stop logit; outputs up to each utterance's stop index are bit-identical
   * to the full-length decode, rows of the DEVICE output tensors beyond *out_steps_done are zero (the alignment carries one more
   * valid row: the exit step's attention has run), host output buffers are only written up to the exit, out_states is undefined.
   * out_stop_index / out_steps_done may be requested with early_stop = 0 too (the decode then runs all `steps`). */
  int32_t early_stop;
  /* Which bf16 decoder kernel runs (ignored by fp32 handles).  GSTK_KERNEL_AUTO: the small-batch latency kernel for free-running
   * SMA decodes of batch <= 16 and key_time <= 256 without early_stop, else the batch-256 kernel.  The two kernels agree within the
   * bf16 tolerance, not bit for bit, so a caller that splits one job into calls of different batch sizes and wants the pieces
   * bit-identical to the whole (gst_tacotron_b200/shard.py) pins GSTK_KERNEL_BATCH.  GSTK_KERNEL_SMALL / _DATAFLOW return
   * GSTK_EINVAL when the call is outside that kernel's domain.  The environment variable GSTK_DECODER=barrier|dataflow overrides
   * AUTO (profiling tools). */
  int32_t kernel;
  int32_t* out_stop_index;  /* [B] first step whose stop logit is negative, `steps` if there is none; or NULL */
  int32_t* out_steps_done;  /* [1] number of steps whose outputs are valid (== steps unless the decode stopped early); or NULL */
  int32_t reserved[2];
} GstkDecodeArgs;

enum { GSTK_KERNEL_AUTO = 0, GSTK_KERNEL_BATCH = 1, GSTK_KERNEL_SMALL = 2, GSTK_KERNEL_DATAFLOW = 3 };

/* Replaces: Style_Token_Layer.call (GST.py:91-109) = Reference_Encoder.call (GST.py:47-70) +
 * MultiHeadAttention.call / Layer_Norm (Layers.py:172-214, 280-285). */
typedef struct GstkGstArgs {
  int32_t batch;            /* B */
  int32_t frames;           /* T' = shape(mels_for_gst)[1] (including the initial frame if drop_first) */
  int32_t drop_first;       /* 1: Style_Token_Layer semantics (mels[:,1:], GST.py:98); 0: Reference_Encoder */
  int32_t pad0;
  const float* mels;        /* [B,frames,mel_dim] */
  const int32_t* lengths;   /* [B] mel_lengths (un-shifted, GST.py:38-40,67) */
  float* out_gst;           /* [B,style_size]  Style_Token_Layer output, or NULL */
  float* out_ref;           /* [B,ref_dense]   Reference_Encoder output, or NULL */
  float* out_attention;     /* [B,n_tokens]    head-averaged distribution (Layers.py:212), or NULL */
  void* stream;
  int32_t reserved[8];
} GstkGstArgs;

/* Replaces: MultiHeadAttention.call on arbitrary 2-input [query, value] (Layers.py:172-214). */
typedef struct GstkMhaArgs {
  int32_t batch, tq, tv, dq, dv, size, heads, pad0;
  const float* query;       /* [B,tq,dq] */
  const float* value;       /* [B,tv,dv] */
  const float* q_kernel; const float* q_bias;   /* [dq,size],[size] */
  const float* v_kernel; const float* v_bias;   /* [dv,size],[size] */
  const float* ln_gamma; const float* ln_beta;  /* [size] */
  float* out;               /* [B,tq,size] */
  float* out_attention;     /* [B,tq,tv] or NULL */
  void* stream;
} GstkMhaArgs;

/* Replaces: BahdanauMonotonicAttention.call / StepwiseMonotonicAttention.call (Steps.py:107-136, 3- and
 * 4-input forms) as a stand-alone layer with caller-owned variables (Steps.py:65-86). */
typedef struct GstkAttentionArgs {
  int32_t batch, key_time, query_dim, value_dim, key_dim, size;
  int32_t type;             /* GSTK_ATT_SMA or GSTK_ATT_BMA */
  int32_t pad0;
  float sigmoid_noise;      /* Steps.py:58,212 */
  float pad1;
  const float* query;       /* [B,query_dim] */
  const float* value;       /* [B,T_v,value_dim] */
  const float* key;         /* [B,T_v,key_dim] or NULL: key = projected value (3-input form, Steps.py:124) */
  const float* prev_alignment; /* [B,T_v] */
  const float* noise;       /* [B,T_v] ~N(0,1) or NULL (no noise) */
  const float* q_kernel; const float* q_bias;   /* Query Dense [query_dim,size],[size] */
  const float* v_kernel; const float* v_bias;   /* Value Dense */
  const float* k_kernel; const float* k_bias;   /* Key Dense (4-input form only) */
  const float* attention_v;                     /* [size] */
  const float* attention_score_bias;            /* [] */
  float* out_context;       /* [B,size] */
  float* out_alignment;     /* [B,T_v] */
  void* stream;
} GstkAttentionArgs;

/* Replaces: Decoder.call's Postnet residual (Taco2.py:230) = the tf.keras.Sequential built at Taco2.py:130-149:
 * per layer Conv1D(filters, kernel_size, strides=1, padding='same', use_bias=False) -> BatchNormalization (moving
 * statistics) -> tanh where `index < len(Filters) - 1` -> Dropout (identity at inference); post = Sequential(x) + x.
 * Variables: "Decoder/Postnet/conv1d_{i}/kernel" [k,in,out], "Decoder/Postnet/batch_normalization_{i}/{gamma,beta,
 * moving_mean,moving_variance}" [out], loaded through gstk_load_weights.  The handle's precision selects the kernels
 * (GSTK_PREC_FP32: fp32 FFMA; GSTK_PREC_BF16: tensor cores with fp16 operands and fp32 accumulation - the outputs have
 * magnitude ~8, and the 1e-2 absolute tolerance needs the 11-bit mantissa).  Strides other than 1 break the reference's own residual add and are not accepted. */
typedef struct GstkPostnetArgs {
  int32_t batch;            /* B */
  int32_t frames;           /* T*r = shape(decodings)[1] */
  int32_t n_layers;         /* len(Conv.Filters) + 1 (<= 8) */
  int32_t pad0;
  int32_t filters[8];       /* Conv.Filters + [Mel_Dim] */
  int32_t kernel[8];        /* Conv.Kernel_Size + [5] */
  int32_t use_tanh[8];      /* 1 for index < len(Conv.Filters) - 1 (Taco2.py:145-146) */
  const float* decodings;   /* [B,frames,mel_dim] */
  float* out_post;          /* [B,frames,mel_dim]  post_decodings */
  void* stream;
  int32_t reserved[8];
} GstkPostnetArgs;

/* Replaces: Encoder.call (Taco2.py:47-51) = the tf.keras.Sequential built at Taco2.py:16-45: Embedding -> per conv layer
 * Conv1D(filters, kernel_size, strides=1, padding='same', use_bias=False) -> BatchNormalization (moving statistics) -> ReLU ->
 * Dropout (identity at inference) -> Bidirectional(LSTM(rnn_size, return_sequences=True)), outputs [forward | backward].
 * Variables (gstk_load_weights): "Encoder/embedding/embeddings" [vocab,E], "Encoder/conv1d_{i}/kernel" [k,in,out],
 * "Encoder/batch_normalization_{i}/{gamma,beta,moving_mean,moving_variance}", "Encoder/bidirectional/{forward,backward}_lstm/
 * lstm_cell/{kernel [in,4u], recurrent_kernel [u,4u], bias [4u]}".  No padding mask, like the reference. */
typedef struct GstkEncoderArgs {
  int32_t batch;            /* B */
  int32_t key_time;         /* T_v = shape(tokens)[1] */
  int32_t vocab;            /* len(token_Index_Dict) (Taco2.py:19) */
  int32_t embedding;        /* Tacotron2.Encoder.Embedding.Size (512) */
  int32_t n_layers;         /* len(Tacotron2.Encoder.Conv.Filters) (<= 8) */
  int32_t rnn_size;         /* Tacotron2.Encoder.RNN.Size (256): output channels = 2 * rnn_size */
  int32_t filters[8];       /* Tacotron2.Encoder.Conv.Filters */
  int32_t kernel[8];        /* Tacotron2.Encoder.Conv.Kernel_Size (strides must be 1) */
  const int32_t* tokens;    /* [B,T_v] token ids */
  float* out;               /* [B,T_v,2*rnn_size] */
  void* stream;
  int32_t reserved[8];
} GstkEncoderArgs;

/* Replaces: Vocoder_Taco1.call (Taco2.py:258-260) = Dense(Spectrogram_Dim)(CBHG(mels)), called on the Postnet output
 * (Model.py:126-129).  CBHG (Taco2.py:285-385): ConvBank = concat_k ReLU(BatchNormalization(Conv1D(bank_filters, k, 'same', no
 * bias))), k = 1 .. bank_count (:388-414) -> MaxPool1D(pool_size, pool_strides, 'same') -> Conv1D_Projection (per entry Conv1D
 * 'same' no bias -> BatchNormalization -> ReLU on all but the last; then Dense(mel_dim) when the last filter count differs from
 * mel_dim) -> + mels -> Highwaynet (Dense(highway_size) when mel_dim differs from it, then highway_count layers
 * relu(xWr+br)*s + x*(1-s), s = sigmoid(xWs+bs), :416-434) -> Bidirectional(LSTM(rnn_size, return_sequences=True)).
 * Variables (gstk_load_weights; gst_tacotron_b200/weights.py:vocoder_spec): "Vocoder_Taco1/CBHG/ConvBank_{i}/{conv1d/kernel,
 * batch_normalization/*}", ".../Conv1D_Projection/{conv1d_{i}/kernel, batch_normalization_{i}/*, dense/{kernel,bias}}",
 * ".../Highwaynet/{dense/*, highwaynet_{i}/{Dense_Relu,Dense_Sigmoid}/{kernel,bias}}", ".../RNN/{forward,backward}_lstm/lstm_cell/
 * {kernel,recurrent_kernel,bias}", "Vocoder_Taco1/Dense/{kernel,bias}".  pool_strides other than 1 change the frame count and
 * break the reference's own residual add (:372): rejected.  Precision as for the Postnet. */
typedef struct GstkVocoderArgs {
  int32_t batch;            /* B */
  int32_t frames;           /* T = shape(mels)[1] */
  int32_t bank_count;       /* Vocoder_Taco1.CBHG.Conv_Bank.Stack_Count (8) */
  int32_t bank_filters;     /* ....Conv_Bank.Filters (256) */
  int32_t pool_size;        /* ....Pool.Pool_Size (2) */
  int32_t pool_strides;     /* ....Pool.Strides (1) */
  int32_t n_proj;           /* len(....Conv1D.Filters) (<= 8) */
  int32_t highway_count;    /* ....Highwaynet.Count (4) */
  int32_t highway_size;     /* ....Highwaynet.Size (128) */
  int32_t rnn_size;         /* ....RNN.Size (256) */
  int32_t spectrogram_dim;  /* Sound.Spectrogram_Dim (513) */
  int32_t pad0;
  int32_t proj_filters[8];  /* ....Conv1D.Filters */
  int32_t proj_kernel[8];   /* ....Conv1D.Kernel_Size */
  const float* mels;        /* [B,frames,mel_dim] */
  float* out;               /* [B,frames,spectrogram_dim] */
  void* stream;
  int32_t reserved[8];
} GstkVocoderArgs;

/* Replaces: Audio.inv_spectrogram (Audio.py:23-27) as Export_Inference calls it (Model.py:412-420), for a whole batch:
 * de-normalise (max_abs_value > 0: _symmetric_denormalize, else _denormalize; Audio.py:96-100) -> 10^((S + ref_level_db)/20) ->
 * ^power -> Griffin-Lim (Audio.py:57-68: random initial phases, `iters` rounds of librosa.stft / istft with n_fft = win_length =
 * 2 (num_freq - 1), hop_length, periodic Hann window, center=True with reflect padding) -> inverse pre-emphasis (Audio.py:14-15).
 * The reference transposes ONE utterance at a time to [num_freq, frames_b]; here the vocoder's [B, frames, num_freq] tensor goes
 * in as it is with the per-utterance frame counts in `lengths` (Model.py:413: max(1, stop index) * Step_Reduction).
 * Utterance b yields hop_length * (lengths[b] - 1) samples; the rest of its out_wav row is zero (lengths[b] < 2: all zero -
 * librosa cannot frame an empty signal).  rng_mode GSTK_RNG_EXTERNAL: init_uniform stands for np.random.rand(*S.shape)
 * (Audio.py:61); GSTK_RNG_PHILOX: the library draws them (Philox stream 3, step = frame, row = row_offset + b, item = bin). */
typedef struct GstkGriffinLimArgs {
  int32_t batch;            /* B */
  int32_t frames;           /* T (padded length) */
  int32_t num_freq;         /* Sound.Spectrogram_Dim: n_fft = 2 (num_freq - 1), a power of two in [64, 4096] */
  int32_t hop_length;       /* Sound.Frame_Shift; must divide n_fft */
  int32_t win_length;       /* Sound.Frame_Length; must equal n_fft (as in the reference's configuration) */
  int32_t iters;            /* Vocoder_Taco1.Griffin-Lim_Iter (60) */
  int32_t rng_mode;         /* GSTK_RNG_EXTERNAL or GSTK_RNG_PHILOX */
  int32_t row_offset;       /* Philox row of utterance 0 (sharded jobs) */
  float ref_level_db;       /* 20 */
  float power;              /* 1.5 */
  float max_abs_value;      /* Sound.Max_Abs_Mel (4); <= 0: None */
  float preemphasis;        /* 0.97 */
  uint64_t seed;
  const float* spectrogram; /* [B,frames,num_freq] normalised (vocoder output) */
  const int32_t* lengths;   /* [B] frames per utterance, or NULL (= frames) */
  const float* init_uniform;/* [B,frames,num_freq] in [0,1), rng_mode EXTERNAL */
  float* out_wav;           /* [B, hop_length * (frames - 1)] */
  void* stream;
  int32_t reserved[8];
} GstkGriffinLimArgs;

/* Replaces: Prenet.call (Taco2.py:282-283) as a stand-alone layer: Sequential[Dense(relu) -> Dropout] x 2 with the dropout ALWAYS on
 * (`training= True   #Always true`), on rows of Mel_Dim channels, with the handle's Decoder_Step/Prenet variables.  (Inside gstk_decode
 * the same two layers are part of the persistent kernel.)  rng_mode: GSTK_RNG_NONE = no dropout, GSTK_RNG_EXTERNAL = keep0 / keep1
 * given ({0,1} floats), GSTK_RNG_PHILOX = the decoder's streams: row i draws (step, row_offset + i). */
typedef struct GstkPrenetArgs {
  int32_t rows;             /* number of input rows (batch, or batch * time) */
  int32_t rng_mode;
  uint64_t seed;
  uint32_t step;            /* Philox step index */
  int32_t row_offset;       /* Philox row of input row 0 */
  const float* inputs;      /* [rows, mel_dim] */
  const float* keep0;       /* [rows, prenet0] or NULL */
  const float* keep1;       /* [rows, prenet1] or NULL */
  float* out;               /* [rows, prenet1] */
  void* stream;
  int32_t reserved[4];
} GstkPrenetArgs;

int gstk_version(void);
int gstk_create(const GstkConfig* cfg, GstkHandle** out);          /* model construction (Taco2.py:59-94, GST.py:12-89) */
int gstk_destroy(GstkHandle* h);
int gstk_load_weights(GstkHandle* h, const GstkTensorDesc* tensors, int32_t n); /* Checkpoint.restore (Model.py:267-276) */
int gstk_decode(GstkHandle* h, const GstkDecodeArgs* args);        /* Decoder.call loop / Decoder_Step.call */
int gstk_gst(GstkHandle* h, const GstkGstArgs* args);              /* Style_Token_Layer.call / Reference_Encoder.call */
int gstk_postnet(GstkHandle* h, const GstkPostnetArgs* args);  /* Postnet(decodings) + decodings (Taco2.py:230) */
int gstk_encoder(GstkHandle* h, const GstkEncoderArgs* args);  /* Encoder.call (Taco2.py:47-51) */
int gstk_vocoder(GstkHandle* h, const GstkVocoderArgs* args);  /* Vocoder_Taco1.call (Taco2.py:258-260) */
int gstk_griffin_lim(GstkHandle* h, const GstkGriffinLimArgs* args); /* Audio.inv_spectrogram (Audio.py:23-27) */
int gstk_prenet(GstkHandle* h, const GstkPrenetArgs* args);    /* Prenet.call (Taco2.py:282-283) */
int gstk_mha(GstkHandle* h, const GstkMhaArgs* args);              /* MultiHeadAttention.call */
int gstk_attention_step(GstkHandle* h, const GstkAttentionArgs* args); /* Bahdanau/StepwiseMonotonicAttention.call */
int gstk_concat_encoder(GstkHandle* h, const float* enc_text, const float* gst, float* out,
                        int32_t batch, int32_t key_time, void* stream); /* GST_Concated_Encoder.call (GST.py:115-124) */
int gstk_synchronize(GstkHandle* h, void* stream);
/* number of kernels this library launched on the handle since creation (bench.py's gpu_launches) */
int64_t gstk_launch_count(GstkHandle* h);
/* device time (ms, CUDA events on the call's stream) of the main kernel of the last
 * gstk_decode / gstk_gst call; synchronises the stream. */
float gstk_last_kernel_ms(GstkHandle* h);
const char* gstk_last_error(GstkHandle* h);                        /* h may be NULL (create errors) */
/* Diagnostics: one tcgen05 tile D[128x32] = bf16(A[128xK]) . bf16(B[32xK])^T (K % 64 == 0, K <= 512) through the
 * same operand packing, bulk copies, descriptors and TMEM loads the bf16 decoder uses. Host pointers. */
/* Diagnostics: per-CTA clock64 ticks the last bf16 decode launch spent in
 * {phase A, barrier 1, phase B, barrier 2, phase C, barrier 3, -, -}; out is [n_ctas][16] uint64 (slots 6-9: sub-phases of A); returns
 * the number of CTAs through *n_ctas (0 when no bf16 decode has run). */
int gstk_get_phase_profile(GstkHandle* h, uint64_t* out, int32_t max_ctas, int32_t* n_ctas);
int gstk_selftest_umma(GstkHandle* h, const float* A, const float* B, int32_t K, float* D);

#ifdef __cplusplus
}
#endif
#endif /* GSTK_H_ */
