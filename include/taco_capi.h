/*
 * taco_capi.h — C ABI of the B200-native Tacotron hot path (libtaco_b200.so).
 *
 * The reference (GSByeon/multi-speaker-tacotron-tensorflow) has no FFI boundary: its hot
 * path is reached through a Python object protocol (models/__init__.py:6 create_model,
 * models/tacotron.py:21 Tacotron.initialize, :274 add_loss, :305 add_optimizer) that
 * builds TensorFlow stock ops.  This header is the boundary inserted underneath that
 * Python surface: every entry point names the reference lines whose arithmetic it
 * replaces.  The reference-side binding (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  Every function returns
 *     0 on success or a negative TACO_E* code; text via taco_last_error() (thread-local).
 *   - All tensor pointers are DEVICE pointers borrowed from the caller (the Python side
 *     passes torch tensors' data_ptr()); the library allocates nothing persistent except
 *     cuFFT plans inside a Griffin-Lim handle.
 *   - All work is enqueued on the caller's cudaStream_t (passed as void*); no hidden
 *     synchronisation except where stated.
 *   - Layouts are TensorFlow's: activations row-major [N,T,C]; dense W[in,out]; conv
 *     W[k,in,out]; GRU gates W[(x;h),(r|u)], candidate W[(x;h),h].
 */
#ifndef TACO_CAPI_H_
#define TACO_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TACO_OK        0
#define TACO_EINVAL   (-1)
#define TACO_ESHAPE   (-2)
#define TACO_ECUDA    (-3)
#define TACO_ENOMEM   (-4)
#define TACO_ESTATE   (-5)

#define TACO_ABI_VERSION 4

/* attention_type (reference: models/tacotron.py:132-152; only these three are reachable) */
#define TACO_ATT_BAH_MON  0
#define TACO_ATT_BAH      1
#define TACO_ATT_BAH_NORM 2

/* speaker injection mode (reference: models/tacotron.py:41-94) */
#define TACO_SPK_NONE            0
#define TACO_SPK_SIMPLE          1
#define TACO_SPK_DEEPVOICE       2   /* five dense(16->d, softsign) sites */
#define TACO_SPK_DEEPVOICE_TABLE 3   /* speaker_embedding_size == 1: five lookup tables */

/* compute precision of the contraction kernels (state, statistics, scans stay fp32):
 *   FP32: every contraction in fp32 FMA (exact-parity mode)
 *   TF32: GEMM-shaped work on tcgen05 tensor cores (kind::tf32, fp32 accumulate in TMEM), fast tanh/sigmoid in the recurrences
 *   BF16: as TF32, but every large contraction reads bf16 operands (kind::f16, fp32 accumulate in TMEM): the producing
 *         kernels write bf16 mirrors of the activations / gradients that feed a GEMM, parameters get a bf16 mirror per step
 *         (BASELINE.json configs[1] names bf16 as the training precision) */
#define TACO_PREC_FP32 0
#define TACO_PREC_TF32 1
#define TACO_PREC_BF16 2

/* POD mirror of the hparams the hot path reads (reference: hparams.py:31-69,83-94). */
typedef struct taco_config {
    int32_t abi_version;           /* must be TACO_ABI_VERSION */
    int32_t num_symbols;           /* text/symbols.py:13 -> 80 */
    int32_t embedding_size;        /* 256 */
    int32_t num_speakers;
    int32_t speaker_mode;          /* TACO_SPK_* */
    int32_t speaker_embedding_size;/* 16 */
    int32_t enc_prenet_sizes[2];   /* 256,128 */
    int32_t enc_bank_size;         /* 16 */
    int32_t enc_bank_channels;     /* 128 */
    int32_t enc_proj_sizes[2];     /* 128,128 */
    int32_t enc_proj_width;        /* 3 */
    int32_t enc_highway_depth;     /* 4 */
    int32_t enc_rnn_size;          /* 128 */
    int32_t attention_type;        /* TACO_ATT_* */
    int32_t attention_size;        /* 256 */
    int32_t attention_state_size;  /* 256 */
    int32_t dec_prenet_sizes[2];   /* 256,128 */
    int32_t dec_layer_num;         /* 2 */
    int32_t dec_rnn_size;          /* 256 */
    int32_t post_bank_size;        /* 8 */
    int32_t post_bank_channels;    /* 256 */
    int32_t post_proj_sizes[2];    /* 256,80 */
    int32_t post_proj_width;       /* 3 */
    int32_t post_highway_depth;    /* 4 */
    int32_t post_rnn_size;         /* 256 */
    int32_t num_mels;              /* 80 */
    int32_t num_freq;              /* 1025 */
    int32_t reduction_factor;      /* r */
    int32_t precision;             /* TACO_PREC_* */
    int32_t device;                /* CUDA ordinal */
    int32_t prioritize_loss;       /* models/tacotron.py:283-295 */
    int32_t priority_lo, priority_hi; /* linear-bin range of the prioritised band */
} taco_config;

typedef struct taco_model_s* taco_model;

/* One named tensor inside a flat buffer: offset/numel in elements. */
typedef struct taco_param_entry {
    const char* name;
    int64_t     offset;
    int64_t     numel;
    int32_t     trainable;         /* 1: lives in the trainable buffer, 0: BN-state buffer */
} taco_param_entry;

/* One batch, device pointers.  reference: the feed of train.py:217-219 / synthesizer.py:166 */
typedef struct taco_batch {
    int32_t N, T_in, T_out;        /* T_out = r * decoder steps; ignored when targets are NULL */
    const int32_t* inputs;         /* [N,T_in] token ids */
    const int32_t* input_lengths;  /* [N] */
    const int32_t* speaker_id;     /* [N] or NULL */
    const float*   mel_targets;    /* [N,T_out,num_mels] or NULL */
    const float*   linear_targets; /* [N,T_out,num_freq] or NULL  (non-NULL => is_training, tacotron.py:26) */
    const float*   loss_coeff;     /* [N] or NULL (=> ones) */
    const float*   manual_alignments; /* [N,T_dec,T_in] or NULL (rnn_wrappers.py:313-317) */
    int32_t        decoder_steps;  /* inference: number of steps (max_iters); training: 0 => T_out/r */
    int32_t        rnn_decoder_test_mode; /* helpers.py:63-66 */
    int32_t        linear_targets_bf16;   /* 1: linear_targets points at bf16 values (same [N,T_out,num_freq] layout): halves the
                                           * host->device bytes of a step (105 -> 52 MB at batch 32 x 800 frames); the loss is then
                                           * taken against the bf16-rounded targets */
} taco_batch;

/* Scalars produced by a step (device-resident copy lives in the workspace; this is the host copy). */
typedef struct taco_step_scalars {
    float loss, mel_loss, linear_loss, loss_without_coeff;  /* tacotron.py:274-302 */
    float grad_norm;                                        /* global norm before clipping (:331) */
    float learning_rate;                                    /* :314-326 */
} taco_step_scalars;

const char* taco_last_error(void);
int taco_abi_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t taco_launch_count(void);

/* ---- model lifetime -------------------------------------------------------------- */
/* replaces: models/__init__.py:6 create_model + the graph construction of tacotron.py:21-251 */
int taco_create(taco_model* out, const taco_config* cfg);
int taco_destroy(taco_model m);

/* Parameter table bound by NAME (see params.py).  Buffers are flat fp32, caller-owned:
 * params/grads/adam_m/adam_v have n_trainable elements, bn_state n_state elements. */
int taco_bind_params(taco_model m, const taco_param_entry* table, int32_t n_entries,
                     float* params, float* grads, float* adam_m, float* adam_v,
                     float* bn_state, int64_t n_trainable, int64_t n_state);

/* Workspace sizing: bytes needed for activations/stashes for a batch shape.
 * training: 0 = inference forward only, 1 = forward+backward stashes. */
int taco_workspace_bytes(taco_model m, int32_t N, int32_t T_in, int32_t T_out_or_steps,
                         int32_t training, size_t* bytes);
int taco_bind_workspace(taco_model m, void* ws, size_t bytes);
/* Named region lookup inside the bound workspace (for tests / zero-copy output views).
 * Valid after a forward with the same shape.  offset in bytes; dims up to 4.  *ndim encodes the element type: n = fp32,
 * -n = double, 100 + n = bf16 (the "<name>.h" mirrors of the bf16 precision mode). */
int taco_ws_region(taco_model m, const char* name, size_t* offset_bytes, int64_t* numel,
                   int64_t dims[4], int64_t strides[4], int32_t* ndim);

/* ---- hot path ---------------------------------------------------------------------- */
/* replaces: Tacotron.initialize forward, tacotron.py:29-251 (encoder CBHG, attention
 * decoder, post CBHG, linear projection).  Outputs land in workspace regions
 * "mel_outputs" [N,T_out,M], "linear_outputs" [N,T_out,F], "alignments" [N,T_in,T_dec]. */
int taco_forward(taco_model m, const taco_batch* b, void* stream);

/* replaces: add_loss (tacotron.py:274-302) + TF autodiff (:328) — fills the grads buffer
 * (overwrites it) and the device scalars; call after taco_forward with training targets. */
int taco_backward(taco_model m, const taco_batch* b, void* stream);

/* replaces: clip_by_global_norm(1.0) + AdamOptimizer.apply_gradients + BN UPDATE_OPS
 * (tacotron.py:327-336).  global_step is the value BEFORE the update (step = global_step+1) and drives the learning-rate
 * schedule; adam_step is the number of Adam updates already applied to these moments (TF keeps it as beta1_power /
 * beta2_power) and drives the bias correction.  The two differ after `--initialize_path`, which restores the optimizer
 * state and resets only global_step (train.py:194-205).  grad_scale multiplies gradients first (1/world after an
 * all-reduce SUM). */
int taco_optimizer_step(taco_model m, int64_t global_step, int64_t adam_step, int32_t is_randomly_initialized,
                        float initial_learning_rate, int32_t decay_mode,
                        float beta1, float beta2, float grad_scale, void* stream);

/* ---- block level: the model's own code paths, one block at a time, for operator-level parity tests -------------
 * All tensors are device fp32, row-major; the model's workspace must be planned for the batch (taco_workspace_bytes) and
 * bound.  Parameter gradients ACCUMULATE into the bound gradient buffer (zero it first).
 *
 * taco_cbhg_forward  replaces models/modules.py:27-96 cbhg(): conv1d bank (:35-44,123-131) -> max-pool (:47-51) -> two
 *   projections (:54-59) -> residual (+ before_highway, :62-69) -> [dense] -> 4x highwaynet (:72-77,105-120) -> bidirectional GRU
 *   (:82-96).  which: 0 = encoder CBHG (inputs [N,T,enc_prenet_sizes[1]]), 1 = post-net CBHG (inputs [N,T,num_mels]);
 *   input_lengths [N] or NULL (the post-net passes none); before_highway [N,proj_sizes[1]] / rnn_init_state [N,2*rnn] or NULL;
 *   outputs [N,T,2*rnn].  is_training selects batch statistics (tf.layers.batch_normalization(training=...)).
 * taco_cbhg_backward: d_outputs [N,T,2*rnn] -> d_inputs [N,T,Cin] (+ optional d_before_highway, d_rnn_init_state); call after
 *   taco_cbhg_forward(is_training = 1) on the same inputs. */
int taco_cbhg_forward(taco_model m, int32_t which, const float* inputs, const int32_t* input_lengths, const float* before_highway,
                      const float* rnn_init_state, int32_t N, int32_t T, int32_t is_training, float* outputs, void* stream);
int taco_cbhg_backward(taco_model m, int32_t which, const float* d_outputs, const int32_t* input_lengths, int32_t N, int32_t T,
                       float* d_inputs, float* d_before_highway, float* d_rnn_init_state, void* stream);
/* taco_decoder_forward replaces models/tacotron.py:127-214 (attention mechanism, DecoderPrenetWrapper / AttentionWrapper /
 *   ConcatOutputAndAttentionWrapper / 2x ResidualWrapper(GRUCell) / r-frame projection of models/rnn_wrappers.py:218-415, the
 *   helpers of models/helpers.py) on a given encoder memory [N,T_in,2*enc_rnn]: the batch supplies the token shape, the targets
 *   (teacher forcing) or decoder_steps (free running), speaker ids and manual alignments exactly as for taco_forward.  Results:
 *   workspace regions "mel_outputs" [N,T_out,M] and "alignments" [N,T_in,T_dec].
 * taco_decoder_backward: d_mel_outputs [N,T_out,M] -> d_memory [N,T_in,2*enc_rnn], decoder parameter gradients accumulated. */
int taco_decoder_forward(taco_model m, const taco_batch* b, const float* memory, void* stream);
int taco_decoder_backward(taco_model m, const taco_batch* b, const float* d_mel_outputs, float* d_memory, void* stream);
/* Element-wise halves of highwaynet (models/modules.py:105-120: y = H*T + x*(1-T); the two dense layers are taco_gemm calls with
 * relu / sigmoid epilogues) and conv1d (:123-131: tf.layers.batch_normalization over the last axis, eps 1e-3; the convolution is
 * a taco_gemm call with tap addressing).  scratch: 4*C doubles of device memory. */
int taco_highway_combine(const float* H, const float* T, const float* x, float* y, int64_t rows, int32_t C, void* stream);
int taco_batch_norm(const float* x, const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                    int32_t N, int32_t T, int32_t C, int32_t is_training, float* y, float* batch_mean, float* batch_var,
                    void* scratch, void* stream);

/* ---- data parallel (replaces nothing in the reference, which is single-device; SURVEY.md 8e: ONE gradient all-reduce per step) ----
 * The flat gradient is reduced in two buckets so that the larger one overlaps the encoder's backward pass:
 *   bucket 0 ("early"): decoder, post-net and linear-projection gradients = the tail [offset, offset + numel) of the flat buffer,
 *                       complete when the decoder's backward pass and its weight-gradient leaves are;
 *   bucket 1 ("late") : embedding, speaker and encoder gradients = the head of the buffer, complete when taco_backward is.
 * taco_dp_wait_bucket makes `stream` (the caller's communication stream) wait for a bucket of the backward pass that was
 * enqueued last; the caller then issues its all-reduce (NCCL) on that stream and orders taco_optimizer_step behind it. */
int taco_dp_bucket(taco_model m, int32_t bucket, int64_t* offset, int64_t* numel);
int taco_dp_wait_bucket(taco_model m, int32_t bucket, void* stream);

/* Copies the device scalars to host (synchronises the stream). */
int taco_read_scalars(taco_model m, taco_step_scalars* out, void* stream);

/* The same read-back without the synchronisation, for training loops that log step k while step k+1 is already
 * enqueued (train.py:217-226 fetches the loss of every step): taco_copy_scalars_async enqueues the device->host copy of
 * the raw accumulators into a caller-owned PINNED buffer of TACO_SCALARS_RAW_BYTES bytes; once the caller's own event
 * after it has completed, taco_finish_scalars turns the raw values into the step scalars (host arithmetic only). */
#define TACO_SCALARS_RAW_BYTES 128
int taco_copy_scalars_async(taco_model m, void* pinned_raw, void* stream);
int taco_finish_scalars(taco_model m, const void* pinned_raw, taco_step_scalars* out);

/* ---- Griffin-Lim (replaces audio/__init__.py:54-56,76-84,99-106,149,158-165) -------- */
typedef struct taco_gl_s* taco_gl;
int taco_gl_create(taco_gl* out, int32_t n_fft, int32_t hop, int32_t win, int32_t max_frames, int32_t device);
int taco_gl_destroy(taco_gl g);
size_t taco_gl_workspace_bytes(taco_gl g);
/* linear_spec: [T,num_freq] normalised-dB spectrogram (device); init_phase: [T,num_freq] in [0,1) or NULL (zeros);
 * wav_out: [hop*(T-1)] device floats. */
int taco_gl_inv_spectrogram(taco_gl g, const float* linear_spec, const float* init_phase, int32_t T,
                            int32_t n_iters, float power, float min_level_db, float ref_level_db,
                            float preemphasis, float* wav_out, void* ws, void* stream);

/* Analysis front end (replaces audio/__init__.py:48-51 spectrogram, :64-67 melspectrogram, :99-101 _stft,
 * :127-131 _linear_to_mel, :145-146 _amp_to_db, :155-156 _preemphasis, :161-162 _normalize): pre-emphasis, centred
 * reflect-padded STFT (cuFFT), magnitude, optional mel projection, dB, normalisation.  wav: [n_samples] device floats;
 * T = 1 + n_samples / hop frames (<= max_frames of the handle).  linear_out: [T, 1 + n_fft/2] or NULL; mel_out:
 * [T, num_mels] or NULL (then mel_basis[num_mels, 1 + n_fft/2], device, is required).  Note the reference subtracts
 * ref_level_db for the linear spectrogram only. */
int taco_audio_spectrogram(taco_gl g, const float* wav, int32_t n_samples, float preemphasis, float ref_level_db,
                           float min_level_db, const float* mel_basis, int32_t num_mels, float* linear_out,
                           float* mel_out, void* stream);

/* Host utility: CRC-32C (Castagnoli) of n bytes, continuing from `crc` (0 to start) — the checksum TensorFlow's checkpoint
 * format uses (replaces the tf.train.Saver dependency of train.py:175,242-244 / synthesizer.py:56-58 when reference
 * checkpoints are imported or exported by tf_checkpoint.py).  No GPU involved. */
uint32_t taco_crc32c(const void* data, size_t n, uint32_t crc);

/* Per-class device timing with CUDA events on the launching stream (bench.py's roofline leg).  enable=1 starts collecting;
 * enable=0 stops, synchronises the device and returns totals per class: 0 GEMMs, 1 GRU recurrences, 2 attention recurrences. */
int taco_profile(int32_t enable, double ms_out[4], int64_t count_out[4]);

/* ---- operator-level entry points (unit parity tests; same kernels the model uses) ---- */
typedef struct taco_gemm_desc {
    const void* A; const void* B; float* C;
    int32_t M, N, K;               /* C[M,N] = alpha * sum_k A(m,k) B(k,n) (+ bias, act, ...) */
    int32_t lda, ldb, ldc;
    int32_t transA;                /* 0: A(m,k)=A[(m+k/ctap)*lda + k%ctap]; 1: A(m,k)=A[(k+m/ctap)*lda + m%ctap] */
    int32_t ctap;                  /* tap width for implicit-conv addressing; 0 = none */
    int32_t transB;                /* 0: B(k,n)=B[k*ldb+n]; 1: B(k,n)=B[n*ldb+k] */
    float   alpha;
    int32_t accumulate;            /* 1: C += result (atomic when split_k>1) */
    const float* bias;             /* [N] or NULL */
    int32_t act;                   /* 0 none, 1 relu, 2 sigmoid, 3 tanh, 4 softsign */
    int32_t mask_period, mask_lo, mask_hi; /* rows with (m % period) outside [lo,hi) are written as 0; period 0 = off */
    int32_t remap_period;          /* 0 = off; else out row ptr = C + (m/period)*remap_outer + (m%period)*remap_inner */
    int64_t remap_outer, remap_inner;
    double* colsum; double* colsumsq; /* optional per-column statistics of the stored values over unmasked rows */
    int32_t split_k;               /* >=1 */
    /* Optional per-k-tile tap table (tensor-core path only, transA = 0, K % 32 == 0): device array of K/32 (col, row)
     * int32 pairs; k-tile i of A row m is A[(m + row_i) * lda + col_i .. + 32).  Lets one GEMM walk several convolutions
     * of different widths that read column blocks of one activation matrix (the conv-bank data gradient).  tap_rows =
     * number of rows addressable from A (rows beyond it read as zero). */
    const int32_t* tap_table; int32_t tap_rows;
    /* bf16 path (precision TACO_PREC_BF16; kind::f16 tensor-core kernel): A16 / B16 are bf16 mirrors of A / B with the same
     * logical layout and pitches (lda / ldb in elements, multiples of 8); when both are given and 16-byte aligned the bf16
     * kernel runs (tap tables then hold one pair per 64-wide k-tile, K % 64 == 0), otherwise the fp32 operands are used.
     * C16 (optional) receives a bf16 copy of the stored values, laid out like C with row pitch ldc16 (0: ldc); C may then be
     * NULL (bf16-only output).  Atomic accumulation (split_k > 1, accumulate == 2) writes fp32 only. */
    const void* A16; const void* B16; void* C16; int32_t ldc16;
} taco_gemm_desc;
int taco_gemm(const taco_gemm_desc* d, int32_t n_problems, int32_t precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TACO_CAPI_H_ */
