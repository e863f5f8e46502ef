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
    NB200_ARCH_MISMATCH = 6,
    NB200_IO_ERROR = 7,     /* whisper::Error::Io: a checkpoint file cannot be read */
    NB200_PARSE_ERROR = 8,  /* whisper::Error::{Json, LoadTokenizer} / candle's safetensors errors */
    NB200_NOT_FOUND = 9,    /* whisper::Error::TokenId: the tokenizer has no such token (src/models/whisper/mod.rs:86-90) */
    NB200_BUFFER_TOO_SMALL = 10 /* a caller-allocated output buffer cannot hold the result; sizes are reported, nothing partial is written */
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

/* the `xa` / `audio_features` ARGUMENT of the reference's decoder calls (model.rs:279, 466): host [n_windows][1500][d_model] f32 become the
 * resident audio features of windows [0, n_windows) without running the encoder (a maintainer who keeps candle's encoder, or a test that
 * feeds the decoder directly) */
NB200_API int nb200_set_audio_features(nb200_ctx *ctx, const float *xa, size_t n_windows);

/* ---- seam (3): replaces `Type::decoder_forward(tokens, xa, flush)` -> candle `TextDecoder::forward`
 *      (model.rs:466-476).  tokens: ALL n tokens so far of window `window` (the reference keeps no
 *      self-attention cache); xa = that window's resident audio features; flush != 0 rebuilds the
 *      cross-attention K/V cache.  hidden_out (nullable): [n][d_model] f32, all n positions.
 *      Cost: the reference recomputes all n positions on every call (O(n^2) steps per window over norma's loop, model.rs:317-322).
 *      Here a call with flush == 0 whose first k tokens are the tokens of the previous call on the same window runs only positions
 *      [k, n) — one step per token when the five-method seam is bound literally — and returns the cached rows for [0, k). ---------- */
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

/* the same loop opened up (what the reference does per iteration of model.rs:317-371 is visible between the calls):
 *   begin    prompt positions, no-speech probability, first sampled token (model.rs:285-315 and the first pass of the loop)
 *   advance  up to n_steps more positions for all windows; *all_done (nullable) = 1 if every window had already finished (nothing ran)
 *   peek     logits [vocab] of `window` at the last computed position = what `decoder_final_linear` returned there (model.rs:324-329)
 *   end      the DecodingResults, as nb200_decode returns them.  nb200_decode == begin; advance(16) until done; end. */
NB200_API int nb200_decode_begin(nb200_ctx *ctx, size_t n_windows, float temperature, uint64_t seed, size_t max_new_tokens);
NB200_API int nb200_decode_advance(nb200_ctx *ctx, size_t n_steps, int *all_done);
NB200_API int nb200_decode_peek_logits(nb200_ctx *ctx, size_t window, float *logits_out);
NB200_API int nb200_decode_end(nb200_ctx *ctx, uint32_t *tokens_out, size_t *n_tokens, double *avg_logprob, double *no_speech_prob);

/* how the greedy steady state runs: NB200_DECODE_AUTO (default) = the fused cooperative step kernel when the context supports it (bf16),
 * NB200_DECODE_SEPARATE = the per-operation kernels replayed as a CUDA graph (always used for t > 0 and in F32 mode).  Same results up to the
 * rounding of the GEMV inputs (the fused kernel feeds its tensor-core GEMVs bf16 activations). */
typedef enum { NB200_DECODE_AUTO = 0, NB200_DECODE_SEPARATE = 1 } nb200_decode_mode;
NB200_API int nb200_set_decode_mode(nb200_ctx *ctx, nb200_decode_mode mode);

/* the same loop at any temperature: t = 0 is nb200_decode_greedy; t > 0 replaces `softmax(p / t)` + `WeightedIndex::sample`
 * (model.rs:340-348) by an inverse-CDF draw on the device from a counter-based generator seeded with `seed` (the reference
 * seeds StdRng from entropy, monolingual.rs:439, so only the distribution is comparable). */
NB200_API int nb200_decode(nb200_ctx *ctx, size_t n_windows, float temperature, uint64_t seed, size_t max_new_tokens, uint32_t *tokens_out,
                           size_t *n_tokens, double *avg_logprob, double *no_speech_prob);

/* ---- replaces `Model::detect_language(audio_features)` (model.rs:194-210), multilingual checkpoints only: one decoder pass over
 *      [sot] with flush = true, the logits of the `n_langs` language tokens (norma: the 99 `Language` variants in declaration
 *      order, multilingual.rs:251-254), softmax over those alone, and the most probable one (the earlier token on a tie, as the
 *      reference's stable descending sort).  probs_out (nullable): [n_langs] f32.  n_langs <= 1024. ---------------------- */
NB200_API int nb200_detect_language(nb200_ctx *ctx, size_t window, const uint32_t *lang_tokens, size_t n_langs, uint32_t *token_out, float *probs_out);

/* ---- host side of norma's whisper `Model` (C++ mirror of model.rs:55-191): buffering, 30 s slicing,
 *      temperature fallback, timestamp segmentation and seek.  ctx = NULL creates a model over a SCRIPTED backend whose
 *      decoding results are pushed with nb200_model_script_push (host-logic tests without a GPU). ------------------- */
typedef struct nb200_model nb200_model;
NB200_API int nb200_model_create(nb200_ctx *ctx, const nb200_special_tokens *tok, size_t max_chunk_len, uint64_t seed, nb200_model **out);
NB200_API void nb200_model_destroy(nb200_model *m);
NB200_API const char *nb200_model_last_error(nb200_model *m);
NB200_API int nb200_model_set_vocab(nb200_model *m, uint32_t id, const char *bytes, size_t n);
/* replaces `Model::transcribe(&mut self, data: &mut Vec<f32>, final_chunk) -> Result<String, _>` (model.rs:55-160).
 * text_out: NUL-terminated UTF-8; seg_out: [n_segments, len_0, tokens_0.., len_1, tokens_1.., ..].  *text_len (bytes, without the NUL)
 * and *seg_len (words) always report the full sizes.  If a buffer is too small NOTHING partial is written and the call returns
 * NB200_BUFFER_TOO_SMALL: the audio has been consumed, the result is kept and nb200_model_last_result delivers it into larger buffers.
 * One deliberate divergence: a decoding result without a drainable segment (the silent window: prompt only, avg_logprob 0) makes the
 * reference loop forever on the same slice (model.rs:68-150 never drains it); here the slice is dropped like the no-speech skip of
 * model.rs:95-98 and counted in nb200_model_no_progress_windows. */
NB200_API int nb200_model_transcribe(nb200_model *m, const float *data, size_t n, int final_chunk, char *text_out, size_t text_cap,
                                     size_t *text_len, uint32_t *seg_out, size_t seg_cap, size_t *seg_len);
NB200_API int nb200_model_last_result(nb200_model *m, char *text_out, size_t text_cap, size_t *text_len, uint32_t *seg_out, size_t seg_cap,
                                      size_t *seg_len);
NB200_API int nb200_model_no_progress_windows(nb200_model *m, size_t *n);
NB200_API int nb200_model_state(nb200_model *m, size_t *buffered, size_t *n_encodes, size_t *n_decodes, size_t *n_resets);
NB200_API int nb200_model_script_push(nb200_model *m, double avg_logprob, double no_speech_prob, const uint32_t *tokens, size_t n);
NB200_API int nb200_model_script_log(nb200_model *m, size_t i, size_t *encode_len, double *decode_temp);

/* `LanguageState::Detect` (model.rs:392-440; multilingual.rs:319-322): the language is detected with nb200_detect_language on the
 * first window of every transcription and forgotten on final_chunk (model.rs:154,170-173).  Without this call the model is
 * `LanguageState::ConstLang(tok.lang)` (monolingual.rs:449).  nb200_model_language: current language token (UINT32_MAX = none yet). */
NB200_API int nb200_model_set_language_detection(nb200_model *m, const uint32_t *lang_tokens, size_t n);
NB200_API int nb200_model_language(nb200_model *m, uint32_t *token, size_t *n_detects);
NB200_API int nb200_model_script_push_language(nb200_model *m, uint32_t token);

/* ---- checkpoint files: what `Definition::blocking_try_to_model` does once hf-hub has fetched config.json,
 *      tokenizer.json and model.safetensors (monolingual.rs:347-430, multilingual.rs:225-300).  The download is out of scope.
 *      These are host-only; ctx-less failures leave their message in nb200_last_error(NULL). ----------------------------- */
typedef enum { NB200_TASK_TRANSCRIBE = 0, NB200_TASK_TRANSLATE = 1 } nb200_task; /* multilingual::Task (multilingual.rs:19-25) */
/* replaces `serde_json::from_str::<Config>` (monolingual.rs:347).  max_batch is set to 1.  suppress_out (nullable) receives up to
 * suppress_cap ids of Config::suppress_tokens, *n_suppress their count. */
NB200_API int nb200_config_from_file(const char *path, nb200_config *cfg, uint32_t *suppress_out, size_t suppress_cap, size_t *n_suppress);
/* replaces `include_bytes!("./whisper_mel_bytes/{80,128}.bytes")` (monolingual.rs:351-362): out = [n_mel][201] f32 */
NB200_API int nb200_mel_filters(int n_mel, float *out);
/* replaces `tokenizers::Tokenizer::from_file` / `token_to_id` / `decode(ids, skip_special_tokens)` (monolingual.rs:348; mod.rs:86-90;
 * model.rs:147): byte-level BPE vocabulary + added tokens; decode output is NUL-terminated UTF-8, truncated to cap, *len = full length */
typedef struct nb200_tokenizer nb200_tokenizer;
NB200_API int nb200_tokenizer_from_file(const char *path, nb200_tokenizer **out);
NB200_API void nb200_tokenizer_destroy(nb200_tokenizer *t);
NB200_API int nb200_tokenizer_token_to_id(const nb200_tokenizer *t, const char *token, uint32_t *id);
NB200_API int nb200_tokenizer_decode(const nb200_tokenizer *t, const uint32_t *ids, size_t n, int skip_special_tokens, char *out, size_t cap, size_t *len);
/* the id lookups of monolingual.rs:376-384,419-420 / multilingual.rs:239-249; language_token e.g. "<|en|>" or NULL (no language token) */
NB200_API int nb200_tokenizer_special_tokens(const nb200_tokenizer *t, const char *language_token, int task, nb200_special_tokens *out);
/* ids of the 99 `Language` tokens in declaration order (languages.rs:7-107; multilingual.rs:251-254): out[99] */
NB200_API int nb200_tokenizer_language_tokens(const nb200_tokenizer *t, uint32_t *out);
/* replaces `VarBuilder::from_mmaped_safetensors(&[weights_file], m::DTYPE, &device)` (monolingual.rs:371-372): mmaps the file and
 * hands every `model.*` tensor (F32 / F16 / BF16 / F64, converted to f32 like candle) to nb200_load_tensor; finalize separately */
NB200_API int nb200_load_safetensors(nb200_ctx *ctx, const char *path, size_t *n_tensors);
/* replaces `quantized_var_builder::VarBuilder::from_gguf` + `quantized_model::Whisper::load` (monolingual.rs:364-369) for the `Quantized*`
 * model types: GGUF v2 / v3 with F32, F16 and Q8_0 tensors, every `model.*` tensor DEQUANTISED to f32 and handed to nb200_load_tensor (candle
 * keeps q8_0 weights and quantises activations per block: parity with it is within that quantisation error, see loader.h).
 * nb200_model_from_files recognises a GGUF weights file by its magic. */
NB200_API int nb200_load_gguf(nb200_ctx *ctx, const char *path, size_t *n_tensors);
NB200_API int nb200_gguf_read(const char *path, const char *name, float *out, size_t cap, int64_t *shape, int *rank, int *type);
/* ctx-less read of one tensor converted to f32 (tests, tools): out (nullable) receives min(cap, numel) values, shape up to 8 dims */
NB200_API int nb200_safetensors_read(const char *path, const char *name, float *out, size_t cap, int64_t *shape, int *rank);
/* `transcribe` then detokenizes with the tokenizer (a copy is kept) instead of the nb200_model_set_vocab table */
NB200_API int nb200_model_set_tokenizer(nb200_model *m, const nb200_tokenizer *t);
/* all of the above in `blocking_try_to_model` order: language_token = "<|en|>".. -> monolingual / MultiAsMono (ConstLang);
 * NULL -> multilingual (Detect).  On success the caller owns *ctx_out and *model_out (destroy the model first). */
NB200_API int nb200_model_from_files(int ordinal, const char *config_json, const char *tokenizer_json, const char *safetensors, nb200_dtype compute,
                                     const char *language_token, int task, size_t max_chunk_len, uint64_t seed, nb200_ctx **ctx_out, nb200_model **model_out);

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
