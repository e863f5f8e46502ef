// common.cuh — shared declarations of libnorma_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/norma_b200.h"

typedef __nv_bfloat16 bf16;

// Whisper constants norma takes from candle (`m::N_FFT`, `m::HOP_LENGTH`, `m::N_SAMPLES`, `m::N_FRAMES`;
// used at /root/reference/src/models/whisper/model.rs:69,88)
constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_BINS = 201;
constexpr int N_SAMPLES = 480000;
constexpr int N_FRAMES = 3000;   // frames kept per window
constexpr int HEAD_DIM = 64;
constexpr int MEL_PAD_FRAMES = 1500;  // candle pads n_len by one 15 s chunk of zeros
constexpr int LN_MAX_SLOTS = 20;      // fused LayerNorm: (d_model / 128 n-tiles) x 2 epilogue warp groups at most (d_model <= 1280)
constexpr float LN_EPS = 1e-5f;       // candle_nn LayerNorm eps

// ---------------------------------------------------------------------------------------------------------
// GEMM epilogue description shared by the SIMT fp32 and the tcgen05 bf16 kernels.
//   v = acc + bias[n];  if (n < n_scale) v *= scale;  if (act) v = gelu_tanh(v);
//   if (residual) v += residual[b*res_bs + r*ldr + n];   out[b*out_bs + r*ldo + n] = v   (f32 or bf16)
// Rows are addressed as (batch b, row r < rows_per_batch): this is what lets the conv stem run as a GEMM
// over overlapping row windows of a time-major buffer without an im2col copy.
// ---------------------------------------------------------------------------------------------------------
struct Epilogue {
    const float *bias;      // [N] or nullptr
    const float *residual;  // f32 or nullptr (may alias out when out is f32)
    void *out;
    long long ldr, res_bs;  // residual row / batch stride (elements)
    long long ldo, out_bs;  // out row / batch stride (elements)
    float scale;            // applied to columns [0, n_scale)
    int n_scale;
    int act;                // 0 none, 1 gelu(tanh)
    int out_bf16;           // 1: out is bf16, 0: f32
    // ---- LayerNorm folded into the adjacent GEMMs (tcgen05 bf16 path only; DESIGN.md §4 "fused LayerNorm") ----
    // producer (f32 out + residual = the new residual-stream rows): also writes a bf16 copy of the rows and per-row partial (mean, M2)
    // slots, one per (n tile, epilogue warp group)
    bf16 *xb_out;           // [rows][ldo] bf16 copy of out (same row addressing as out), or nullptr
    float2 *stats_out;      // [slots][stats_ld] partial (mean, M2) per row, or nullptr
    long long stats_ld;     // rows per slot plane
    // consumer (A = the bf16 copy, W = W . diag(gamma) . (I - 11^T / K): gamma and the centring live in the weight):
    // out = rstd . acc + bias[n], bias = b + W . beta, rstd from the row's slots
    const float2 *stats_in; // the producer's slots for A's rows, or nullptr
    int stats_slots;        // slots per row
    int stats_cols;         // columns per slot
};

struct GemmShape {
    int rows_per_batch;  // M per batch
    int batch;
    int N, K;
    long long lda, a_bs;  // A row / batch stride (elements)
};

struct EncLayer {
    void *wqkv, *wo, *w1, *w2;  // compute dtype; with ctx->ln_fold: wqkv = Wqkv . diag(ln1 gamma) . (I - 11^T / d), w1 likewise with ln2
    float *bqkv, *bo, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b;  // with ctx->ln_fold: bqkv = b + Wqkv . beta1, b1 = b + W1 . beta2
};
struct DecLayer {
    void *wqkv, *wo, *cwq, *cwkv, *cwo, *w1, *w2;
    float *bqkv, *bo, *cbq, *cbkv, *cbo, *b1, *b2;
    float *ln1g, *ln1b, *lncg, *lncb, *ln2g, *ln2b;
};

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

// debug / A-B switches, read from the environment ONCE per context at nb200_create (no process-wide mutable state on the launch path:
// different contexts are driven from different threads, include/norma_b200.h "Conventions")
struct CtxOptions {
    int gemm_mode = 2;        // NB200_GEMM: 2 = CTA-pair 256 x 256 tiles (default), 1 = "1cta", 3 = "1cta128"
    int gemm_epi_tma = 1;     // NB200_EPI=direct -> 0: thread-per-row stores
    int gemm_nofit = 0;       // NB200_GEMM_NOFIT: keep the pair tile for one window's worth of rows
    int gemm_debug = 0;       // NB200_GEMM_DEBUG: microbenchmark switches of gemm_tc_kernel
    int gemm_np3 = 0;         // NB200_GEMM_NP3=1: a third staging patch per epilogue warp (4 pipeline stages) for the short-K f32-residual GEMM: out-proj 153.6 -> 156.0 us, off
    int gemm_wide = 1;        // NB200_GEMM_WIDE=0: no single-round wide tiles for one window's worth of rows (gemm_wide_kernel)
    int attn_tc = 1;          // NB200_ATTN=simt -> 0: CUDA-core attention
    int decode_fused = 1;     // NB200_DECODE_FUSED=0 -> per-operation decode kernels
    int decode_graph = 1;     // NB200_DECODE_NOGRAPH -> 0
    int encoder_graph = 1;    // NB200_ENCODER_NOGRAPH -> 0: launch the encoder kernel by kernel
    int pdl = 7;              // NB200_PDL: bit mask of the kernels launched with programmatic dependent launch (1 GEMM, 2 attention, 4 LayerNorm;
                              // 0 = off).  Every kernel of the chain owns its SM's shared memory, so only the ~1 us prologue (barrier init, TMEM
                              // allocation, tensor-map prefetch) overlaps the previous kernel's tail: 3.19 -> 3.12 ms per window, nothing at 25
                              // windows.  (It looked unstable earlier in round 2; that was the attention race, profiles/r2_notes.md sections 1, 3.)
    int ln_fused = 1;         // NB200_LN_FUSED=0 -> standalone LayerNorm kernels in the encoder
    int prof_dump = 0;        // NB200_PROF_DUMP
};

// a captured launch sequence and what it stands for in the context's counters
struct CapturedGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;                     // kernels in the graph (ctx->launches is advanced by this on every replay)
    int64_t by_class[NB200_K_COUNT] = {0};
};

struct nb200_ctx {
    int ordinal = 0;
    CtxOptions opt;
    nb200_config cfg{};
    nb200_dtype compute = NB200_BF16;
    cudaStream_t stream = nullptr;
    std::string err;
    std::vector<void *> allocs;
    size_t device_bytes = 0;
    int64_t launches = 0;
    int sm_count = 148;

    // load state
    std::map<std::string, HostTensor> host_tensors;
    bool finalized = false;
    bool has_filters = false, has_tokens = false;
    bool has_decoder = false;

    // weights
    void *conv1_w = nullptr, *conv2_w = nullptr;  // [d][3*n_mel], [d][3*d] (k-major taps, compute dtype)
    float *conv1_b = nullptr, *conv2_b = nullptr;
    float *pos = nullptr;  // sinusoids [1500][d] f32
    std::vector<EncLayer> enc;
    float *lnpost_g = nullptr, *lnpost_b = nullptr;
    void *embed = nullptr;       // [V][d] compute dtype
    float *embed_pos = nullptr;  // [448][d] f32
    std::vector<DecLayer> dec;
    float *lndec_g = nullptr, *lndec_b = nullptr;

    // mel
    float *filt_vals = nullptr;  // banded filterbank: [n_mel][MEL_BAND] values
    int *filt_start = nullptr;   // [n_mel] first non-zero bin
    int *filt_len = nullptr;     // [n_mel] mel row of slot entry e (slot-ordered filterbank, see mel_setup_filters)
    int *mel_slot_len = nullptr; // [n_mel / 8] warp-uniform trip count per slot
    float *mel_tables = nullptr;  // hann[400] | tw200[8*25*2] | tw400[201*2]
    float *pcm = nullptr;         // [max_batch][N_SAMPLES]
    int *pcm_len = nullptr;       // [max_batch]
    float *logmel = nullptr;      // [max_batch][n_mel][N_FRAMES] un-normalised log10
    unsigned *mel_max = nullptr;  // [max_batch] ordered-uint encoding of the running max
    float *mel_norm = nullptr;    // [max_batch][n_mel][N_FRAMES] normalised f32 (reference layout)
    void *melT = nullptr;         // [max_batch][N_FRAMES+2][n_mel] time-major, zero pad rows 0 and N_FRAMES+1
    size_t *host_lens = nullptr;
    // pipelined batches (nb200_transcode_submit / _collect): two slots of PCM and feature buffers, copies on their own streams
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    float *pipe_pcm[2] = {nullptr, nullptr};
    int *pipe_len[2] = {nullptr, nullptr};
    float *pipe_out[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_mel[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    std::vector<int> pipe_lens_host[2];
    long long pipe_submitted = 0, pipe_collected = 0;
    int stream_len = 0;           // streaming (window 0): samples currently buffered on the device
    float *stream_tmp = nullptr;  // scratch for the seek shift: [N_SAMPLES] + [n_mel][N_FRAMES]

    // encoder activations
    void *y1 = nullptr;     // [max_batch][N_FRAMES+1][d] conv1 out, row 0 = zero pad (compute dtype)
    float *x = nullptr;     // [max_batch*1500][d] residual stream f32
    void *h = nullptr;      // [M][d] LN out (compute dtype); with ln_fold: the bf16 copy of the residual stream
    bool ln_fold = false;   // LayerNorm folded into the encoder GEMMs (bf16 mode, d % 256 == 0, NB200_LN_FUSED != 0); decided at finalize
    float2 *ln_stats = nullptr;  // [LN_MAX_SLOTS][max_batch*1500] per-row partial (mean, M2)
    int ln_slots = 0, ln_slot_cols = 0;  // what the last producer launch wrote
    void *qkv = nullptr;    // [M][3d]
    void *attn = nullptr;   // [M][d]
    void *ff = nullptr;     // [M][4d]
    float *enc_out = nullptr;  // [M][d] f32 (audio_features)
    void *enc_out_c = nullptr; // [M][d] compute dtype copy (A operand of the cross-K/V GEMM); == enc_out in f32 mode
    int n_resident = 0;        // windows with valid features

    // decoder state
    void *cross_kv = nullptr;  // [L][max_batch][1500][2d]  (k | v), compute dtype
    void *self_kv = nullptr;   // [L][max_batch][448][2d]
    bool cross_valid = false;
    float *dx = nullptr, *dh = nullptr, *dqkv = nullptr, *dattn = nullptr, *dff = nullptr, *dq = nullptr;  // [max_batch][..] f32
    float *dhid = nullptr;   // [max_batch][d] final LN output
    float *logits = nullptr; // [max_batch][V]
    uint32_t *d_tokens = nullptr;  // [max_batch][max_target_positions]
    int *d_len = nullptr, *d_last_ts = nullptr, *d_done = nullptr, *d_nsampled = nullptr;
    double *d_sumlp = nullptr;
    float *d_nospeech = nullptr;
    void *d_dec_layers = nullptr;  // device copy of `dec` (fused decoder step)
    void *d_fused_sel_ws = nullptr;    // greedy-select partials of the fused step, one vocabulary chunk per CTA
    float *d_fused_attn_ws = nullptr;  // split-K partials of the fused step's attention [max_batch][heads][64][66]
    void *d_fused_sync = nullptr;  // monotonic grid-barrier counter (u32, own 128 B line) | attention tickets [max_batch][heads]
    int fused_ctas = 0;            // cooperative grid of the fused decoder step (0 = not sized yet)
    bool decode_separate = false;  // nb200_set_decode_mode(NB200_DECODE_SEPARATE)
    bool fused_failed = false;     // a cooperative launch was refused once: this context stays on the separate kernels
    void *d_lang = nullptr;      // detect_language scratch: [NB200_MAX_LANGS] u32 ids | [NB200_MAX_LANGS] f32 probs | i32 best
    void *d_sel_ws = nullptr;    // greedy select partials: [max_batch][32] float2 + [max_batch][32] SelCand
    float *d_attn_ws = nullptr;  // split-K decode attention partials [max_batch][heads][8][66]
    void *d_dyn = nullptr;  // DecodeDyn (decoder.cu): device-resident position / temperature / seed / token budget
    std::map<int, cudaGraphExec_t> step_graphs;  // captured decode step (embed .. logits .. select) per batch size
    std::map<long long, CapturedGraph> front_graphs;  // captured log-mel / encoder passes per (windows, stages, buffer slot)
    // decode in progress (nb200_decode_begin .. _end)
    struct DecodeRun { bool active = false; int B = 0, plen = 0, pos = 0, greedy = 1; bool use_fused = false; cudaGraphExec_t gexec = nullptr; } run;
    // seam (3) incremental state: the window whose self-attention K/V cache and hidden rows cover `seam_tokens`
    int seam_window = -1;
    std::vector<uint32_t> seam_tokens;
    float *seam_hidden = nullptr;  // [max_target_positions][d] f32, allocated on first use
    float *suppress = nullptr;  // [V] additive mask (0 / -inf): Config::suppress_tokens U {no_timestamps}
    nb200_special_tokens tok{};
    std::vector<uint32_t> suppress_ids;

    // misc
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    bool profiling = false;
    struct ProfRec { int cls; cudaEvent_t a, b; long long tag; };
    std::map<long long, std::pair<double, long long>> prof_by_tag;  // debug: per (class, shape) time
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> ev_pool;
    float prof_ms[NB200_K_COUNT] = {0};
    int64_t prof_launches[NB200_K_COUNT] = {0};
    double prof_gemm_flops = 0;
    void *host_pinned = nullptr;
    size_t host_pinned_bytes = 0;
    CUtensorMap *tmap_scratch = nullptr;
};

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
int nb200_fail(nb200_ctx *ctx, int code, const char *fmt, ...);
#define CUDA_TRY(ctx, expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return nb200_fail(ctx, _e == cudaErrorMemoryAllocation ? NB200_OOM : NB200_CUDA_ERROR,      \
                              "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define NB_TRY(expr)               \
    do {                           \
        int _s = (expr);           \
        if (_s != NB200_OK) return _s; \
    } while (0)

// per-kernel-class timing scope (events only when profiling is enabled)
struct KernelScope {
    nb200_ctx *ctx;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    long long tag = 0;
    KernelScope(nb200_ctx *c, int k, long long tag_ = 0);
    ~KernelScope();
};

// ---------------------------------------------------------------------------------------------------------
// kernel launchers (one per .cu file)
// ---------------------------------------------------------------------------------------------------------
// mel.cu
int launch_mel(nb200_ctx *ctx, int n_windows);            // pcm -> logmel (+ running max)
int launch_mel_norm(nb200_ctx *ctx, int n_windows);       // logmel -> mel_norm (f32, reference layout) + melT
int launch_mel_from_host_layout(nb200_ctx *ctx, int n_windows);  // mel_norm (already normalised) -> melT
int mel_setup_tables(nb200_ctx *ctx);
int mel_stream_reset(nb200_ctx *ctx);
int mel_stream_update(nb200_ctx *ctx, int f_lo, int f_hi);
int mel_stream_window_max(nb200_ctx *ctx);
int mel_stream_shift(nb200_ctx *ctx, int ns, int new_len, float *tmp_pcm, float *tmp_lm);
int mel_setup_filters(nb200_ctx *ctx, const float *filters, int n_mel);
// simt.cu
int launch_gemm_f32(nb200_ctx *ctx, const float *A, const float *W, const GemmShape &s, const Epilogue &e);
int launch_layernorm(nb200_ctx *ctx, const float *x, const float *g, const float *b, int rows, int d, void *out, int out_bf16,
                     float *out2_f32);
int launch_attention_simt(nb200_ctx *ctx, const void *qkv, void *out, int B, int T, int n_heads, int is_bf16);
int launch_f32_to_bf16(nb200_ctx *ctx, const float *in, bf16 *out, size_t n);
// gemm_tcgen05.cu
int launch_gemm_bf16(nb200_ctx *ctx, const bf16 *A, const bf16 *W, const GemmShape &s, const Epilogue &e);
// attn_tcgen05.cu
int launch_attention_tc(nb200_ctx *ctx, const bf16 *qkv, bf16 *out, int B, int T, int n_heads);
// decoder.cu
int decoder_build_cross_kv(nb200_ctx *ctx, int n_windows);
int decoder_step(nb200_ctx *ctx, int w0, int n_windows, int pos, int want_logits);  // windows [w0, w0+n); scratch rows 0..n
int decoder_copy_hidden(nb200_ctx *ctx, float *dst, int d);
int decoder_init(nb200_ctx *ctx);
int simt_init(nb200_ctx *ctx);
int gemm_tc_init(nb200_ctx *ctx);
int attn_tc_init(nb200_ctx *ctx);
int tmap_encode_bf16(nb200_ctx *ctx, CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box);
int decoder_select(nb200_ctx *ctx, int n_windows, int greedy);  // also advances the device-resident position
int decoder_set_dyn(nb200_ctx *ctx, int pos, int max_new, float temperature, unsigned long long seed, int set_params);
int decoder_nospeech(nb200_ctx *ctx, int n_windows);
int decoder_step_fused(nb200_ctx *ctx, int n_windows, int n_steps);  // n_steps greedy steps in one cooperative launch (device-resident position)
bool decoder_fused_supported(const nb200_ctx *ctx);
int decoder_fused_max_windows();
int decoder_fused_prepare(nb200_ctx *ctx);
int decoder_fused_ws_floats(const nb200_ctx *ctx);  // one greedy step, one cooperative launch (device-resident position)
constexpr int NB200_MAX_LANGS = 1024;
int decoder_language(nb200_ctx *ctx, int n_langs);  // softmax over logits[ids] of row 0, first-index argmax -> d_lang
int decoder_init_state(nb200_ctx *ctx, int n_windows);
// api.cu
int encoder_run(nb200_ctx *ctx, int n_windows);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch of a kernel on the encoder chain: optional cluster shape, and (ctx->opt.pdl) programmatic dependent launch, so the kernel's
// prologue overlaps the tail of its predecessor.  ONLY for kernels that execute ptx::griddep_wait() before their first global access.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(nb200_ctx *ctx, int pdl_bit, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster_x, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster_x;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (ctx->opt.pdl & pdl_bit) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
static inline size_t dtype_size(nb200_dtype t) { return t == NB200_F32 ? 4 : 2; }

// tanh-approximation GELU exactly as candle's `Tensor::gelu`
__device__ __forceinline__ float gelu_tanh_precise(float x) {
    const float k0 = 0.7978845608028654f;  // sqrt(2/pi)
    const float k1 = 0.044715f;
    float u = k0 * x * (1.0f + k1 * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
    const float k0 = 0.7978845608028654f;
    const float k1 = 0.044715f;
    float u = k0 * x * (1.0f + k1 * x * x);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * x * (1.0f + t);
}
