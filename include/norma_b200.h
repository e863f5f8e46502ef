/*
 * norma_b200.h — C ABI of the B200-native (sm_100a) hot path of MikeIvanichev/norma.
 *
 * This is the boundary a maintainer binds from Rust (`norma-b200-sys`) in place of the
 * candle calls norma makes on its hot path.  Every entry point cites the reference interface it replaces
 * (paths are under the reference tree, e.g. src/models/whisper/model.rs).
 *
 * Conventions
 *   - every function returns an nb200_status (0 = ok); the message is available from nb200_last_error().
 *     Nothing aborts, exits or throws across this boundary.
 *   - one nb200_ctx = one GPU ordinal = one private CUDA stream; calls on one ctx are serialised by the
 *     caller (norma drives a model from exactly one thread, src/lib.rs:377,464); different ctxs are
 *     independent and may be driven concurrently from different threads (the 8-GPU mode).
 *   - the ctx owns all device memory; the caller owns every host pointer; host inputs are consumed and host
 *     outputs are complete when the call returns.
 *   - there is no CPU fallback: a device that is not compute capability 10.x fails nb200_create with
 *     NB200_ARCH_MISMATCH.
 */
#ifndef NORMA_B200_H
#define NORMA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define NB200_API __attribute__((visibility("default")))
#else
#define NB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_ctx nb200_ctx;

typedef enum {
    NB200_OK = 0,
    NB200_INVALID_ARG = 1,
    NB200_CUDA_ERROR = 2,
    NB200_OOM = 3,
    NB200_NOT_LOADED = 4,
    NB200_UNSUPPORTED_SHAPE = 5,
    NB200_ARCH_MISMATCH = 6
} nb200_status;

/* replaces candle_core::DType as returned by norma's `DType::to_dtype` (src/dtype.rs:15,23) */
typedef enum { NB200_F32 = 0, NB200_BF16 = 1, NB200_F16 = 2, NB200_F64 = 3, NB200_U8 = 4, NB200_U32 = 5 } nb200_dtype;

/* replaces candle_transformers::models::whisper::Config as parsed at src/models/whisper/monolingual.rs:347 */
typedef struct {
    int32_t num_mel_bins;            /* 80 | 128 (monolingual.rs:351-362) */
    int32_t max_source_positions;    /* 1500 */
    int32_t d_model;
    int32_t encoder_attention_heads;
    int32_t encoder_layers;
    int32_t vocab_size;
    int32_t max_target_positions;    /* 448 */
    int32_t decoder_attention_heads;
    int32_t decoder_layers;
    int32_t max_batch;               /* windows processed per call by the *_batch entry points (>= 1) */
} nb200_config;

/* the ids norma looks up from tokenizer.json at load (monolingual.rs:376-384, 419-420) */
typedef struct {
    uint32_t sot, eot, task, lang /* UINT32_MAX = no language token in the prompt */;
    uint32_t no_speech, no_timestamps, ts_zero /* <|0.00|> */, ts_one /* <|1.00|> */;
} nb200_special_tokens;

typedef enum {
    NB200_Q_N_FRAMES = 0,        /* frames kept per window (3000) */
    NB200_Q_ENC_LEN = 1,         /* encoder positions per window (1500) */
    NB200_Q_D_MODEL = 2,
    NB200_Q_VOCAB = 3,
    NB200_Q_MAX_BATCH = 4,
    NB200_Q_KERNEL_LAUNCHES = 5, /* kernels launched by this ctx since creation (monotonic) */
    NB200_Q_DEVICE_BYTES = 6,    /* device bytes owned by this ctx */
    NB200_Q_COMPUTE_DTYPE = 7,
    NB200_Q_MAX_TARGET_POSITIONS = 8
} nb200_query_key;

/* kernel classes for the live per-kernel timers (bench.py roofline) */
typedef enum {
    NB200_K_MEL = 0, NB200_K_MEL_NORM = 1, NB200_K_GEMM = 2, NB200_K_ATTN = 3, NB200_K_LAYERNORM = 4,
    NB200_K_DECODE_GEMV = 5, NB200_K_DECODE_ATTN = 6, NB200_K_DECODE_SELECT = 7, NB200_K_MISC = 8, NB200_K_COUNT = 9
} nb200_kernel_class;

/* ---- device / context: replaces `SelectedDevice::Cuda(n)` -> `candle_core::Device::new_cuda(n)`
 *      (src/models/mod.rs:38-55) ------------------------------------------------------------------------- */
NB200_API int nb200_device_count(int *out);
NB200_API int nb200_create(int ordinal, const nb200_config *cfg, nb200_dtype compute /* NB200_BF16 | NB200_F32 */, nb200_ctx **out);
NB200_API void nb200_destroy(nb200_ctx *ctx);
/* ctx may be NULL: returns the calling thread's last error from a ctx-less call (create / device_count) */
NB200_API const char *nb200_last_error(nb200_ctx *ctx);
NB200_API int nb200_query(nb200_ctx *ctx, nb200_query_key key, int64_t *out);
NB200_API int nb200_sync(nb200_ctx *ctx);

/* ---- load: replaces `VarBuilder::from_mmaped_safetensors(.., m::DTYPE, &device)` + `Whisper::load`
 *      (monolingual.rs:371-373).  `hf_name` is the safetensors tensor name ("model.encoder.conv1.weight", ...);
 *      `model.encoder.embed_positions.weight` is accepted and ignored (sinusoids are recomputed). --------- */
NB200_API int nb200_load_tensor(nb200_ctx *ctx, const char *hf_name, const void *host, nb200_dtype dtype, const int64_t *shape, int rank);
NB200_API int nb200_finalize_weights(nb200_ctx *ctx);
/* replaces the `mel_filters: Vec<f32>` handed to pcm_to_mel (monolingual.rs:351-362): [n_mel][201] f32 */
NB200_API int nb200_set_mel_filters(nb200_ctx *ctx, const float *filters, int n_mel);
/* replaces the token-id fields and the four vocab-sized mask tensors of `Model` (monolingual.rs:376-430) */
NB200_API int nb200_set_tokens(nb200_ctx *ctx, const nb200_special_tokens *tok);
NB200_API int nb200_set_suppress(nb200_ctx *ctx, const uint32_t *ids, size_t n); /* Config::suppress_tokens */

/* ---- seam (1): replaces `audio::pcm_to_mel(&config, data_slice, &mel_filters)` (model.rs:74).
 *      pcm: n <= 480000 f32 samples of one window.  mel_out (nullable): [n_mel][*n_len] f32, mel-major, ALL
 *      n_len frames exactly as candle returns them (norma narrows to min(3000, n_len), model.rs:87-88).
 *      Pass mel_out = NULL to query *n_len.  The kept frames stay resident on the device as window 0. -------- */
NB200_API int nb200_pcm_to_mel(nb200_ctx *ctx, const float *pcm, size_t n, float *mel_out, size_t *n_len);
/* batch form: n_windows <= max_batch windows, window w at pcm + w*stride, lens[w] samples (NULL = stride).
 * mel_out (nullable): [n_windows][n_mel][3000] f32 (the kept frames). */
NB200_API int nb200_pcm_to_mel_batch(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens, float *mel_out);

/* ---- seam (2): replaces `Type::encoder_forward(mel, flush)` -> candle `AudioEncoder::forward`
 *      (model.rs:455-464).  mel: host [n_windows][n_mel][3000] f32 or NULL = use the mel left resident by
 *      nb200_pcm_to_mel*.  out (nullable): host [n_windows][1500][d_model] f32.  The encoder output stays
 *      resident as the `audio_features` of windows 0..n_windows-1. ---------------------------------------- */
NB200_API int nb200_encoder_forward(nb200_ctx *ctx, const float *mel, size_t n_windows, float *out);

/* ---- fused front half for throughput configs (BASELINE configs 2, 3, 5): PCM -> log-mel -> encoder in one
 *      call; H2D of PCM, both stages and (if out != NULL) the D2H of the features happen inside. ---------- */
NB200_API int nb200_transcode_batch(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens, float *out);
/* pipelined form: submit enqueues H2D (own stream) -> log-mel + encoder -> D2H (own stream) for one batch and returns; at most two
 * batches may be in flight; collect blocks until the OLDEST submitted batch's features are in its `out`.  `pcm` and `out` must stay
 * valid (and should be pinned) until that batch is collected.  Batch k+1's upload and batch k-1's download overlap batch k's compute. */
NB200_API int nb200_transcode_submit(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens, float *out);
NB200_API int nb200_transcode_collect(nb200_ctx *ctx);
/* the same on inputs already resident in HBM (bench `value`): stage once, run many times */
NB200_API int nb200_stage_pcm(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens);
NB200_API int nb200_run_resident(nb200_ctx *ctx, size_t n_windows, int do_mel, int do_encoder);
/* copy the first `n` floats of window w's resident encoder output / mel to the host */
NB200_API int nb200_fetch_features(nb200_ctx *ctx, size_t window, float *out, size_t n);
NB200_API int nb200_fetch_mel(nb200_ctx *ctx, size_t window, float *out, size_t n);

/* ---- seam (3): replaces `Type::decoder_forward(tokens, xa, flush)` -> candle `TextDecoder::forward`
 *      (model.rs:466-476).  tokens: ALL n tokens so far of window `window` (the reference keeps no
 *      self-attention cache); xa = that window's resident audio features; flush != 0 rebuilds the
 *      cross-attention K/V cache.  hidden_out (nullable): [n][d_model] f32. ------------------------------- */
NB200_API int nb200_decoder_forward(nb200_ctx *ctx, size_t window, const uint32_t *tokens, size_t n, int flush, float *hidden_out);
/* ---- seam (4): replaces `Type::decoder_final_linear(x)` (model.rs:478-483): hidden [d] -> logits [vocab] */
NB200_API int nb200_final_linear(nb200_ctx *ctx, const float *hidden, float *logits_out);
/* ---- seam (5): replaces `Type::reset_kv_cache()` (model.rs:485-490) ----------------------------------- */
NB200_API int nb200_reset_kv_cache(nb200_ctx *ctx);

/* ---- replaces norma's `Model::decode(audio_features, t = 0.0)` (model.rs:279-390) including the
 *      suppression rules (model.rs:212-277), kept entirely on the device: prompt [sot, lang?, task],
 *      no-speech probability, greedy loop with last-index tie-break, logprob sum, stop rule, trailing
 *      timestamp strip.  Decodes windows [0, n_windows) in lock-step.  tokens_out: [n_windows][max_target_positions]
 *      u32; n_tokens/avg_logprob/no_speech_prob: [n_windows].  max_new_tokens = 0 reproduces the reference
 *      stop rule only (model.rs:367-370); > 0 additionally stops (pushing eot) after that many sampled
 *      tokens (bounded tests). ----------------------------------------------------------------------------- */
NB200_API int nb200_decode_greedy(nb200_ctx *ctx, size_t n_windows, size_t max_new_tokens, uint32_t *tokens_out, size_t *n_tokens,
                        double *avg_logprob, double *no_speech_prob);

/* the same loop at any temperature: t = 0 is nb200_decode_greedy; t > 0 replaces `softmax(p / t)` + `WeightedIndex::sample`
 * (model.rs:340-348) by an inverse-CDF draw on the device from a counter-based generator seeded with `seed` (the reference
 * seeds StdRng from entropy, monolingual.rs:439, so only the distribution is comparable). */
NB200_API int nb200_decode(nb200_ctx *ctx, size_t n_windows, float temperature, uint64_t seed, size_t max_new_tokens, uint32_t *tokens_out,
                           size_t *n_tokens, double *avg_logprob, double *no_speech_prob);

/* ---- host side of norma's whisper `Model` (C++ mirror of model.rs:55-191): buffering, 30 s slicing,
 *      temperature fallback, timestamp segmentation and seek.  ctx = NULL creates a model over a SCRIPTED backend whose
 *      decoding results are pushed with nb200_model_script_push (host-logic tests without a GPU). ------------------- */
typedef struct nb200_model nb200_model;
NB200_API int nb200_model_create(nb200_ctx *ctx, const nb200_special_tokens *tok, size_t max_chunk_len, uint64_t seed, nb200_model **out);
NB200_API void nb200_model_destroy(nb200_model *m);
NB200_API const char *nb200_model_last_error(nb200_model *m);
NB200_API int nb200_model_set_vocab(nb200_model *m, uint32_t id, const char *bytes, size_t n);
/* replaces `Model::transcribe(&mut self, data: &mut Vec<f32>, final_chunk) -> Result<String, _>` (model.rs:55-160).
 * text_out: NUL-terminated (truncated to text_cap); seg_out: [n_segments, len_0, tokens_0.., len_1, tokens_1.., ..] */
NB200_API int nb200_model_transcribe(nb200_model *m, const float *data, size_t n, int final_chunk, char *text_out, size_t text_cap,
                                     size_t *text_len, uint32_t *seg_out, size_t seg_cap, size_t *seg_len);
NB200_API int nb200_model_state(nb200_model *m, size_t *buffered, size_t *n_encodes, size_t *n_decodes, size_t *n_resets);
NB200_API int nb200_model_script_push(nb200_model *m, double avg_logprob, double no_speech_prob, const uint32_t *tokens, size_t n);
NB200_API int nb200_model_script_log(nb200_model *m, size_t i, size_t *encode_len, double *decode_temp);

/* ---- streaming front half (BASELINE config 4): what `Packer` + the re-mel of the whole buffer on every chunk
 *      (src/lib.rs:224-262, model.rs:68-74) become on the device.  Window 0 only.  push appends PCM and recomputes only the mel
 *      frames the new samples touch; drain drops samples from the front (norma's seek, model.rs:110,126-127); features
 *      normalises with the global max over the buffered audio (exactly pcm_to_mel of those samples) and runs the encoder. --- */
NB200_API int nb200_stream_reset(nb200_ctx *ctx);
NB200_API int nb200_stream_push(nb200_ctx *ctx, const float *chunk, size_t n);
NB200_API int nb200_stream_drain(nb200_ctx *ctx, size_t n);
NB200_API int nb200_stream_features(nb200_ctx *ctx, int run_encoder, float *mel_out /* [n_mel][3000] or NULL */, float *features_out /* or NULL */);

/* ---- measurement helpers (bench.py): CUDA events on the ctx stream ------------------------------------ */
NB200_API int nb200_timer_start(nb200_ctx *ctx);
NB200_API int nb200_timer_stop(nb200_ctx *ctx, float *ms);          /* records, synchronises, returns elapsed ms */
NB200_API int nb200_profile_enable(nb200_ctx *ctx, int on);         /* per-kernel-class event timing on/off */
NB200_API int nb200_profile_read(nb200_ctx *ctx, float *ms_by_class /*[NB200_K_COUNT]*/, int64_t *launches_by_class /*[NB200_K_COUNT]*/, double *flops_gemm);
NB200_API int nb200_profile_reset(nb200_ctx *ctx);
/* write >= 192 MiB of device memory on the ctx stream (L2 flush between timed iterations) */
NB200_API int nb200_flush_l2(nb200_ctx *ctx);

/* ---- standalone kernel self-tests (tests/ only): C[M,N] = A[M,K] . W[N,K]^T (+bias, act) on the device.
 *      a, w are host bf16 (as uint16) or f32 by `compute`; c_out host f32. ------------------------------- */
NB200_API int nb200_test_gemm(nb200_ctx *ctx, const void *a, const void *w, const float *bias, int M, int N, int K, int act_gelu, float *c_out);
/* GEMM microbenchmark on zero-filled device buffers: epi_kind 0 = bias -> bf16, 1 = bias + GELU -> bf16,
 * 2 = bias + f32 residual in place -> f32; returns the average ms per launch over `iters` launches */
NB200_API int nb200_test_gemm_perf(nb200_ctx *ctx, int M, int N, int K, int epi_kind, int iters, float *ms_out);
/* tcgen05 attention microbenchmark: average ms per launch over `iters` launches on the given q|k|v */
NB200_API int nb200_test_attention_perf(nb200_ctx *ctx, const float *qkv, int B, int T, int n_heads, int iters, float *ms_out);
/* attention self-test: qkv host f32 [B*T][3*d] (q|k|v), h heads of 64 -> ctx_out [B*T][d] f32 */
NB200_API int nb200_test_attention(nb200_ctx *ctx, const float *qkv, int B, int T, int n_heads, float *ctx_out);

#ifdef __cplusplus
}
#endif
#endif /* NORMA_B200_H */
