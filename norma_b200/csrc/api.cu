// api.cu — the C ABI (include/norma_b200.h): context, weight loading, encoder orchestration, decode loop.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

static thread_local std::string g_err;
int decoder_logits_rows(nb200_ctx *ctx, int n);
static void destroy_step_graphs(nb200_ctx *ctx) {
    for (auto &g : ctx->step_graphs) cudaGraphExecDestroy(g.second);
    ctx->step_graphs.clear();
    ctx->run.gexec = nullptr;
}

int nb200_fail(nb200_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_err = buf;
    return code;
}

KernelScope::KernelScope(nb200_ctx *c, int k, long long tag_) : ctx(c), cls(k), tag(tag_) {
    ctx->launches++;
    ctx->prof_launches[cls]++;
    if (ctx->profiling) {
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->ev_pool.empty()) { e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        a = get();
        b = get();
        cudaEventRecord(a, ctx->stream);
    }
}
KernelScope::~KernelScope() {
    if (a) {
        cudaEventRecord(b, ctx->stream);
        ctx->prof_recs.push_back({cls, a, b, tag});
    }
}

namespace {

int dev_alloc(nb200_ctx *ctx, size_t bytes, void **out, bool zero = false) {
    if (bytes == 0) bytes = 16;
    CUDA_TRY(ctx, cudaMalloc(out, bytes));
    ctx->allocs.push_back(*out);
    ctx->device_bytes += bytes;
    if (zero) CUDA_TRY(ctx, cudaMemsetAsync(*out, 0, bytes, ctx->stream));
    return NB200_OK;
}
template <typename T>
int dev_alloc_t(nb200_ctx *ctx, size_t n, T **out, bool zero = false) { return dev_alloc(ctx, n * sizeof(T), (void **)out, zero); }

int set_device(nb200_ctx *ctx) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->ordinal));
    return NB200_OK;
}

int ensure_pinned(nb200_ctx *ctx, size_t bytes) {
    if (ctx->host_pinned_bytes >= bytes) return NB200_OK;
    if (ctx->host_pinned) cudaFreeHost(ctx->host_pinned);
    ctx->host_pinned = nullptr;
    ctx->host_pinned_bytes = 0;
    CUDA_TRY(ctx, cudaMallocHost(&ctx->host_pinned, bytes));
    ctx->host_pinned_bytes = bytes;
    return NB200_OK;
}

int upload_f32(nb200_ctx *ctx, const std::vector<float> &v, float **out) {
    NB_TRY(dev_alloc_t(ctx, v.size(), out));
    CUDA_TRY(ctx, cudaMemcpyAsync(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

// upload in the compute dtype (f32 -> bf16 conversion on the device)
int upload_compute(nb200_ctx *ctx, const std::vector<float> &v, void **out) {
    if (ctx->compute == NB200_F32) return upload_f32(ctx, v, (float **)out);
    float *tmp = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&tmp, v.size() * 4));
    cudaError_t e = cudaMemcpyAsync(tmp, v.data(), v.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    int st = NB200_OK;
    if (e == cudaSuccess) {
        st = dev_alloc(ctx, v.size() * 2, out);
        if (st == NB200_OK) st = launch_f32_to_bf16(ctx, tmp, (bf16 *)*out, v.size());
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(tmp);
    if (e != cudaSuccess) return nb200_fail(ctx, NB200_CUDA_ERROR, "weight upload failed: %s", cudaGetErrorString(e));
    return st;
}

const HostTensor *find(nb200_ctx *ctx, const std::string &name) {
    auto it = ctx->host_tensors.find(name);
    return it == ctx->host_tensors.end() ? nullptr : &it->second;
}

int need(nb200_ctx *ctx, const std::string &name, size_t numel, const HostTensor **out) {
    const HostTensor *t = find(ctx, name);
    if (!t) return nb200_fail(ctx, NB200_NOT_LOADED, "tensor '%s' was not loaded", name.c_str());
    if (t->data.size() != numel)
        return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "tensor '%s' has %zu elements, expected %zu", name.c_str(), t->data.size(), numel);
    *out = t;
    return NB200_OK;
}

int up_vec(nb200_ctx *ctx, const std::string &name, size_t n, float **out) {
    const HostTensor *t = nullptr;
    NB_TRY(need(ctx, name, n, &t));
    return upload_f32(ctx, t->data, out);
}
int up_mat(nb200_ctx *ctx, const std::string &name, size_t n, void **out) {
    const HostTensor *t = nullptr;
    NB_TRY(need(ctx, name, n, &t));
    return upload_compute(ctx, t->data, out);
}

// Fold the LayerNorm in front of a linear layer into it (ctx->ln_fold).  LN(x) = (x - mean(x)) . rstd . gamma + beta and the centring is
// itself linear (x - mean(x) = x . (I - 11^T / K)), so with W' = W . diag(gamma) . (I - 11^T / K) — every row of W . diag(gamma) minus its own
// mean — and b' = b + W . beta:   LN(x) . W^T + b = rstd . (x . W'^T) + b'.
// The consuming GEMM multiplies the bf16 copy of x with W' and applies the row's rstd in its epilogue (gemm_tcgen05.cu).
int fold_layernorm(nb200_ctx *ctx, const std::string &ln, int N, int K, std::vector<float> &W, std::vector<float> &B) {
    const HostTensor *g = nullptr, *bt = nullptr;
    NB_TRY(need(ctx, ln + ".weight", K, &g));
    NB_TRY(need(ctx, ln + ".bias", K, &bt));
    for (int n = 0; n < N; ++n) {
        float *w = W.data() + (size_t)n * K;
        double acc_b = 0.0, acc_s = 0.0;
        for (int k = 0; k < K; ++k) {
            acc_b += (double)w[k] * bt->data[k];
            w[k] *= g->data[k];
            acc_s += (double)w[k];
        }
        const float m = (float)(acc_s / K);
        for (int k = 0; k < K; ++k) w[k] -= m;
        B[n] += (float)acc_b;
    }
    return NB200_OK;
}

// concatenated [q | k | v] weight and bias ([3d][d], [3d] with zeros for the bias-less k_proj); `fold_ln`: the LayerNorm to fold in
int up_qkv(nb200_ctx *ctx, const std::string &p, int d, void **w, float **b, const std::string &fold_ln = "") {
    const HostTensor *q = nullptr, *k = nullptr, *v = nullptr, *bq = nullptr, *bv = nullptr;
    NB_TRY(need(ctx, p + "q_proj.weight", (size_t)d * d, &q));
    NB_TRY(need(ctx, p + "k_proj.weight", (size_t)d * d, &k));
    NB_TRY(need(ctx, p + "v_proj.weight", (size_t)d * d, &v));
    NB_TRY(need(ctx, p + "q_proj.bias", d, &bq));
    NB_TRY(need(ctx, p + "v_proj.bias", d, &bv));
    std::vector<float> W((size_t)3 * d * d), B(3 * d, 0.f);
    memcpy(W.data(), q->data.data(), (size_t)d * d * 4);
    memcpy(W.data() + (size_t)d * d, k->data.data(), (size_t)d * d * 4);
    memcpy(W.data() + (size_t)2 * d * d, v->data.data(), (size_t)d * d * 4);
    memcpy(B.data(), bq->data.data(), d * 4);
    memcpy(B.data() + 2 * d, bv->data.data(), d * 4);
    if (!fold_ln.empty()) NB_TRY(fold_layernorm(ctx, fold_ln, 3 * d, d, W, B));
    NB_TRY(upload_compute(ctx, W, w));
    return upload_f32(ctx, B, b);
}

// conv weight [co][ci][3] -> [co][k*ci_n + ci] (tap-major) so a GEMM row is 3 consecutive time-major input rows
int up_conv(nb200_ctx *ctx, const std::string &name, int co, int ci, void **out) {
    const HostTensor *t = nullptr;
    NB_TRY(need(ctx, name, (size_t)co * ci * 3, &t));
    std::vector<float> W((size_t)co * ci * 3);
    for (int o = 0; o < co; ++o)
        for (int i = 0; i < ci; ++i)
            for (int k = 0; k < 3; ++k) W[((size_t)o * 3 + k) * ci + i] = t->data[((size_t)o * ci + i) * 3 + k];
    return upload_compute(ctx, W, out);
}

float bf16_bits_to_float(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
float f16_bits_to_float(uint16_t h) {
    uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1f, man = h & 0x3ff, u;
    if (exp == 0) {
        if (man == 0) u = sign << 31;
        else {
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400));
            u = (sign << 31) | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) u = (sign << 31) | 0x7f800000u | (man << 13);
    else u = (sign << 31) | ((exp - 15 + 127) << 23) | (man << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int alloc_decoder(nb200_ctx *ctx) {
    const nb200_config &c = ctx->cfg;
    const size_t es = dtype_size(ctx->compute), B = c.max_batch, d = c.d_model;
    NB_TRY(dev_alloc(ctx, (size_t)c.decoder_layers * B * c.max_source_positions * 2 * d * es, &ctx->cross_kv));
    NB_TRY(dev_alloc(ctx, (size_t)c.decoder_layers * B * c.max_target_positions * 2 * d * es, &ctx->self_kv, true));
    NB_TRY(dev_alloc_t(ctx, B * d, &ctx->dx));
    NB_TRY(dev_alloc_t(ctx, B * d, &ctx->dh));
    NB_TRY(dev_alloc_t(ctx, B * 3 * d, &ctx->dqkv));
    NB_TRY(dev_alloc_t(ctx, B * d, &ctx->dattn));
    NB_TRY(dev_alloc_t(ctx, B * 4 * d, &ctx->dff));
    NB_TRY(dev_alloc_t(ctx, B * d, &ctx->dq));
    NB_TRY(dev_alloc_t(ctx, std::max<size_t>(B, c.max_target_positions) * d, &ctx->dhid));
    NB_TRY(dev_alloc_t(ctx, B * c.vocab_size, &ctx->logits));
    NB_TRY(dev_alloc_t(ctx, B * c.max_target_positions, &ctx->d_tokens, true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->d_len, true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->d_last_ts, true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->d_done, true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->d_nsampled, true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->d_sumlp, true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->d_nospeech, true));
    NB_TRY(dev_alloc(ctx, 64, &ctx->d_dyn, true));
    NB_TRY(dev_alloc(ctx, B * 32 * (8 + 32), &ctx->d_sel_ws, true));
    NB_TRY(dev_alloc(ctx, (size_t)NB200_MAX_LANGS * 8 + 16, &ctx->d_lang, true));
    NB_TRY(dev_alloc(ctx, 128 + (size_t)B * c.decoder_attention_heads * 4, &ctx->d_fused_sync, true));
    NB_TRY(dev_alloc(ctx, (size_t)c.decoder_layers * sizeof(DecLayer), &ctx->d_dec_layers));
    NB_TRY(dev_alloc_t(ctx, (size_t)decoder_fused_ws_floats(ctx), &ctx->d_fused_attn_ws, true));
    NB_TRY(dev_alloc_t(ctx, B * (size_t)c.decoder_attention_heads * 8 * 66, &ctx->d_attn_ws, true));
    NB_TRY(dev_alloc_t(ctx, (size_t)c.vocab_size, &ctx->suppress, true));
    return NB200_OK;
}

int upload_suppress(nb200_ctx *ctx) {
    if (!ctx->suppress) return NB200_OK;
    std::vector<float> m(ctx->cfg.vocab_size, 0.f);
    for (uint32_t t : ctx->suppress_ids)
        if (t < (uint32_t)ctx->cfg.vocab_size) m[t] = -INFINITY;
    if (ctx->has_tokens && ctx->tok.no_timestamps < (uint32_t)ctx->cfg.vocab_size) m[ctx->tok.no_timestamps] = -INFINITY;  // monolingual.rs:388
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->suppress, m.data(), m.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int stage_pcm(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens) {
    if (!pcm || n_windows == 0 || n_windows > (size_t)ctx->cfg.max_batch)
        return nb200_fail(ctx, NB200_INVALID_ARG, "stage_pcm: n_windows=%zu (max_batch %d)", n_windows, ctx->cfg.max_batch);
    std::vector<int> l(n_windows);
    for (size_t w = 0; w < n_windows; ++w) {
        size_t n = lens ? lens[w] : stride;
        if (n > (size_t)N_SAMPLES) return nb200_fail(ctx, NB200_INVALID_ARG, "window %zu has %zu samples (> %d)", w, n, N_SAMPLES);
        l[w] = (int)n;
        if (n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pcm + w * N_SAMPLES, pcm + w * stride, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pcm_len, l.data(), n_windows * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `l` and possibly pageable `pcm` are consumed
    return NB200_OK;
}

int check_ready(nb200_ctx *ctx, bool need_weights, bool need_filters) {
    if (!ctx) return NB200_INVALID_ARG;
    NB_TRY(set_device(ctx));
    if (need_weights && !ctx->finalized) return nb200_fail(ctx, NB200_NOT_LOADED, "weights not finalized (call nb200_finalize_weights)");
    if (need_filters && !ctx->has_filters) return nb200_fail(ctx, NB200_NOT_LOADED, "mel filters not set (call nb200_set_mel_filters)");
    return NB200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// encoder: candle `AudioEncoder::forward` on windows [0, B)
// ---------------------------------------------------------------------------------------------------------
int encoder_run(nb200_ctx *ctx, int B) {
    const nb200_config &c = ctx->cfg;
    const int d = c.d_model, T = c.max_source_positions, n_mel = c.num_mel_bins, H = c.encoder_attention_heads;
    const bool bf = ctx->compute == NB200_BF16;
    const int M = B * T;
    const float qk_scale = powf((float)HEAD_DIM, -0.25f);
    auto gemm = [&](const void *A, const void *W, const GemmShape &s, const Epilogue &e) {
        return bf ? launch_gemm_bf16(ctx, (const bf16 *)A, (const bf16 *)W, s, e) : launch_gemm_f32(ctx, (const float *)A, (const float *)W, s, e);
    };
    const size_t es = dtype_size(ctx->compute);
    // conv1 (k3, s1, p1) + GELU as a GEMM over overlapping rows of the time-major mel buffer
    {
        GemmShape s{N_FRAMES, B, d, 3 * n_mel, n_mel, (long long)(N_FRAMES + 2) * n_mel};
        Epilogue e{};
        e.bias = ctx->conv1_b; e.act = 1;
        e.out = (char *)ctx->y1 + (size_t)d * es;  // row 0 of every window is the zero pad
        e.ldo = d; e.out_bs = (long long)(N_FRAMES + 1) * d; e.out_bf16 = bf;
        NB_TRY(gemm(ctx->melT, ctx->conv1_w, s, e));
    }
    // conv2 (k3, s2, p1) + GELU + sinusoidal positions -> residual stream x [B*1500][d] f32
    // With ctx->ln_fold the LayerNorms before QKV and fc1 do not exist as kernels: every GEMM that writes the residual stream (conv2, out-proj,
    // fc2) also writes its bf16 copy into `h` and per-row partial (mean, M2); the consuming GEMM multiplies the copy with W . diag(gamma) and
    // finishes the normalisation in its epilogue (gemm_tcgen05.cu).
    const bool fold = ctx->ln_fold;
    const long long Mmax = (long long)c.max_batch * T;
    auto produce = [&](Epilogue &e) {
        if (!fold) return;
        e.xb_out = (bf16 *)ctx->h;
        e.stats_out = ctx->ln_stats;
        e.stats_ld = Mmax;
    };
    auto consume = [&](Epilogue &e) {
        if (!fold) return;
        e.stats_in = ctx->ln_stats;
        e.stats_ld = Mmax;
        e.stats_slots = ctx->ln_slots;
        e.stats_cols = ctx->ln_slot_cols;
    };
    {
        GemmShape s{T, B, d, 3 * d, 2 * d, (long long)(N_FRAMES + 1) * d};
        Epilogue e{};
        e.bias = ctx->conv2_b; e.act = 1;
        e.residual = ctx->pos; e.ldr = d; e.res_bs = 0;
        e.out = ctx->x; e.ldo = d; e.out_bs = (long long)T * d; e.out_bf16 = 0;
        produce(e);
        NB_TRY(gemm(ctx->y1, ctx->conv2_w, s, e));
    }
    for (int l = 0; l < c.encoder_layers; ++l) {
        const EncLayer &w = ctx->enc[l];
        if (!fold) NB_TRY(launch_layernorm(ctx, ctx->x, w.ln1g, w.ln1b, M, d, ctx->h, bf, nullptr));
        {
            GemmShape s{M, 1, 3 * d, d, d, (long long)M * d};
            Epilogue e{};
            e.bias = w.bqkv; e.scale = qk_scale; e.n_scale = 2 * d;  // q and k each scaled by hd^-0.25
            e.out = ctx->qkv; e.ldo = 3 * d; e.out_bf16 = bf;
            consume(e);
            NB_TRY(gemm(ctx->h, w.wqkv, s, e));
        }
        if (bf && ctx->opt.attn_tc) NB_TRY(launch_attention_tc(ctx, (const bf16 *)ctx->qkv, (bf16 *)ctx->attn, B, T, H));
        else NB_TRY(launch_attention_simt(ctx, ctx->qkv, ctx->attn, B, T, H, bf));
        {
            GemmShape s{M, 1, d, d, d, (long long)M * d};
            Epilogue e{};
            e.bias = w.bo; e.residual = ctx->x; e.ldr = d;
            e.out = ctx->x; e.ldo = d; e.out_bf16 = 0;
            produce(e);
            NB_TRY(gemm(ctx->attn, w.wo, s, e));
        }
        if (!fold) NB_TRY(launch_layernorm(ctx, ctx->x, w.ln2g, w.ln2b, M, d, ctx->h, bf, nullptr));
        {
            GemmShape s{M, 1, 4 * d, d, d, (long long)M * d};
            Epilogue e{};
            e.bias = w.b1; e.act = 1;
            e.out = ctx->ff; e.ldo = 4 * d; e.out_bf16 = bf;
            consume(e);
            NB_TRY(gemm(ctx->h, w.w1, s, e));
        }
        {
            GemmShape s{M, 1, d, 4 * d, 4 * d, (long long)M * 4 * d};
            Epilogue e{};
            e.bias = w.b2; e.residual = ctx->x; e.ldr = d;
            e.out = ctx->x; e.ldo = d; e.out_bf16 = 0;
            if (l + 1 < c.encoder_layers) produce(e);  // ln_post stays a kernel: nothing consumes the last layer's copy
            NB_TRY(gemm(ctx->ff, w.w2, s, e));
        }
    }
    // ln_post -> audio features (f32) and, in bf16 mode, the bf16 copy the cross-K/V GEMM consumes
    if (bf) NB_TRY(launch_layernorm(ctx, ctx->x, ctx->lnpost_g, ctx->lnpost_b, M, d, ctx->enc_out_c, 1, ctx->enc_out));
    else NB_TRY(launch_layernorm(ctx, ctx->x, ctx->lnpost_g, ctx->lnpost_b, M, d, ctx->enc_out, 0, nullptr));
    ctx->n_resident = B;
    ctx->cross_valid = false;
    ctx->seam_window = -1;
    return NB200_OK;
}

// log-mel and / or encoder over windows [0, B) replayed as ONE CUDA graph per (B, stages, buffer slot).  A 32-layer encoder is ~200
// launches: at one window per call (BASELINE configs 2 and 4) issuing them one by one costs more host time than the kernels take.  The
// tensor maps are encoded once, at capture.  Profiling runs (per-kernel events) and NB200_ENCODER_NOGRAPH launch kernel by kernel.
int run_front(nb200_ctx *ctx, int B, bool do_mel, bool do_enc, int slot) {
    auto direct = [&]() -> int {
        if (do_mel) {
            NB_TRY(launch_mel(ctx, B));
            NB_TRY(launch_mel_norm(ctx, B));
        }
        if (do_enc) NB_TRY(encoder_run(ctx, B));
        return NB200_OK;
    };
    if (ctx->profiling || !ctx->opt.encoder_graph) return direct();
    const long long key = ((long long)B << 8) | (do_mel ? 1 : 0) | (do_enc ? 2 : 0) | ((long long)slot << 2);
    auto it = ctx->front_graphs.find(key);
    if (it == ctx->front_graphs.end()) {
        CapturedGraph cg;
        const int64_t l0 = ctx->launches;
        int64_t c0[NB200_K_COUNT];
        for (int i = 0; i < NB200_K_COUNT; ++i) c0[i] = ctx->prof_launches[i];
        const double f0 = ctx->prof_gemm_flops;
        const int res0 = ctx->n_resident;
        const bool cv0 = ctx->cross_valid;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int st = direct();
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        cg.launches = ctx->launches - l0;  // nothing ran yet: the counters describe the graph, every replay adds them
        ctx->launches = l0;
        for (int i = 0; i < NB200_K_COUNT; ++i) {
            cg.by_class[i] = ctx->prof_launches[i] - c0[i];
            ctx->prof_launches[i] = c0[i];
        }
        ctx->prof_gemm_flops = f0;
        ctx->n_resident = res0;
        ctx->cross_valid = cv0;
        if (st != NB200_OK || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (st != NB200_OK) return st;
            return nb200_fail(ctx, NB200_CUDA_ERROR, "front graph capture failed: %s", cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&cg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return nb200_fail(ctx, NB200_CUDA_ERROR, "front graph instantiate failed: %s", cudaGetErrorString(ie));
        it = ctx->front_graphs.emplace(key, cg).first;
    }
    CUDA_TRY(ctx, cudaGraphLaunch(it->second.exec, ctx->stream));
    ctx->launches += it->second.launches;
    for (int i = 0; i < NB200_K_COUNT; ++i) ctx->prof_launches[i] += it->second.by_class[i];
    if (do_enc) {
        ctx->n_resident = B;
        ctx->cross_valid = false;
        ctx->seam_window = -1;
    }
    return NB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int nb200_device_count(int *out) {
    if (!out) return nb200_fail(nullptr, NB200_INVALID_ARG, "device_count: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out = 0;
        return nb200_fail(nullptr, NB200_CUDA_ERROR, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *out = n;
    return NB200_OK;
}

const char *nb200_last_error(nb200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int nb200_create(int ordinal, const nb200_config *cfg, nb200_dtype compute, nb200_ctx **out) {
    if (!cfg || !out) return nb200_fail(nullptr, NB200_INVALID_ARG, "create: NULL argument");
    *out = nullptr;
    if (compute != NB200_BF16 && compute != NB200_F32) return nb200_fail(nullptr, NB200_INVALID_ARG, "create: compute dtype must be BF16 or F32");
    const nb200_config &c = *cfg;
    if (c.num_mel_bins != 80 && c.num_mel_bins != 128)
        return nb200_fail(nullptr, NB200_UNSUPPORTED_SHAPE, "Unexpected number of mel bins (num_mel_bins), got: %d", c.num_mel_bins);  // whisper::Error::MelBins
    if (c.d_model % 128 != 0 || c.d_model > 1280 || c.d_model < 128 || c.encoder_attention_heads * HEAD_DIM != c.d_model ||
        c.decoder_attention_heads * HEAD_DIM != c.d_model)
        return nb200_fail(nullptr, NB200_UNSUPPORTED_SHAPE, "d_model=%d heads=%d/%d: need d_model %% 128 == 0, <= 1280, head_dim 64", c.d_model,
                          c.encoder_attention_heads, c.decoder_attention_heads);
    if (c.max_source_positions != 1500 || c.max_target_positions < 8 || c.max_target_positions > 448 || c.max_batch < 1 || c.encoder_layers < 1 ||
        c.decoder_layers < 0 || c.vocab_size < 1)
        return nb200_fail(nullptr, NB200_UNSUPPORTED_SHAPE, "unsupported config (max_source_positions must be 1500, max_target_positions in [8, 448])");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return nb200_fail(nullptr, NB200_CUDA_ERROR, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    if (ordinal < 0 || ordinal >= ndev) return nb200_fail(nullptr, NB200_INVALID_ARG, "create: ordinal %d out of range (%d devices)", ordinal, ndev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, ordinal);
    if (e != cudaSuccess) return nb200_fail(nullptr, NB200_CUDA_ERROR, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return nb200_fail(nullptr, NB200_ARCH_MISMATCH, "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", ordinal, prop.major,
                          prop.minor);
    nb200_ctx *ctx = new nb200_ctx();
    ctx->ordinal = ordinal;
    ctx->cfg = c;
    ctx->compute = compute;
    ctx->sm_count = prop.multiProcessorCount;
    {   // A/B and debug switches: read once, here, into the context
        auto env = [](const char *k) { return getenv(k); };
        auto is = [&](const char *k, const char *v) { const char *e = env(k); return e && !strcmp(e, v); };
        CtxOptions &o = ctx->opt;
        o.gemm_mode = is("NB200_GEMM", "1cta") ? 1 : is("NB200_GEMM", "1cta128") ? 3 : is("NB200_GEMM", "2cta128") ? 4 : 2;
        o.gemm_epi_tma = is("NB200_EPI", "direct") ? 0 : 1;
        o.gemm_nofit = env("NB200_GEMM_NOFIT") != nullptr;
        o.gemm_debug = env("NB200_GEMM_DEBUG") ? atoi(env("NB200_GEMM_DEBUG")) : 0;
        o.gemm_np3 = (env("NB200_GEMM_NP3") && env("NB200_GEMM_NP3")[0] == '1') ? 1 : 0;
        o.gemm_wide = (env("NB200_GEMM_WIDE") && env("NB200_GEMM_WIDE")[0] == '0') ? 0 : 1;
        o.attn_tc = is("NB200_ATTN", "simt") ? 0 : 1;
        o.decode_fused = (env("NB200_DECODE_FUSED") && env("NB200_DECODE_FUSED")[0] == '0') ? 0 : 1;
        o.decode_graph = env("NB200_DECODE_NOGRAPH") ? 0 : 1;
        o.encoder_graph = env("NB200_ENCODER_NOGRAPH") ? 0 : 1;
        if (env("NB200_PDL")) o.pdl = atoi(env("NB200_PDL"));
        o.ln_fused = (env("NB200_LN_FUSED") && env("NB200_LN_FUSED")[0] == '0') ? 0 : 1;
        o.prof_dump = env("NB200_PROF_DUMP") != nullptr;
    }
    int st = [&]() -> int {
        NB_TRY(set_device(ctx));
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreate(&ctx->ev_start));
        CUDA_TRY(ctx, cudaEventCreate(&ctx->ev_stop));
        const size_t B = c.max_batch, d = c.d_model, T = c.max_source_positions, nm = c.num_mel_bins, es = dtype_size(compute);
        const size_t M = B * T;
        NB_TRY(dev_alloc_t(ctx, (size_t)nm * 16, &ctx->filt_vals, true));
        NB_TRY(dev_alloc_t(ctx, nm, &ctx->filt_start, true));
        NB_TRY(dev_alloc_t(ctx, nm, &ctx->filt_len, true));
        NB_TRY(dev_alloc_t(ctx, nm, &ctx->mel_slot_len, true));
        NB_TRY(dev_alloc_t(ctx, 1280, &ctx->mel_tables, true));
        NB_TRY(dev_alloc_t(ctx, B * N_SAMPLES, &ctx->pcm, true));
        NB_TRY(dev_alloc_t(ctx, B, &ctx->pcm_len, true));
        NB_TRY(dev_alloc_t(ctx, B * nm * N_FRAMES, &ctx->logmel));
        NB_TRY(dev_alloc_t(ctx, B, &ctx->mel_max, true));
        NB_TRY(dev_alloc_t(ctx, B * nm * N_FRAMES, &ctx->mel_norm));
        NB_TRY(dev_alloc(ctx, B * (N_FRAMES + 2) * nm * es, &ctx->melT, true));
        NB_TRY(dev_alloc(ctx, B * (N_FRAMES + 1) * d * es, &ctx->y1, true));
        NB_TRY(dev_alloc_t(ctx, M * d, &ctx->x));
        NB_TRY(dev_alloc(ctx, M * d * es, &ctx->h));
        NB_TRY(dev_alloc(ctx, M * 3 * d * es, &ctx->qkv));
        NB_TRY(dev_alloc(ctx, M * d * es, &ctx->attn));
        NB_TRY(dev_alloc(ctx, M * 4 * d * es, &ctx->ff));
        NB_TRY(dev_alloc_t(ctx, M * d, &ctx->enc_out));
        if (compute == NB200_BF16) NB_TRY(dev_alloc(ctx, M * d * 2, &ctx->enc_out_c));
        else ctx->enc_out_c = ctx->enc_out;
        ctx->flush_bytes = 256ull << 20;
        NB_TRY(dev_alloc(ctx, ctx->flush_bytes, &ctx->flush_buf));
        NB_TRY(mel_setup_tables(ctx));
        NB_TRY(simt_init(ctx));
        NB_TRY(gemm_tc_init(ctx));
        NB_TRY(attn_tc_init(ctx));
        NB_TRY(decoder_init(ctx));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return NB200_OK;
    }();
    if (st != NB200_OK) {
        g_err = ctx->err;
        nb200_destroy(ctx);
        return st;
    }
    *out = ctx;
    return NB200_OK;
}

void nb200_destroy(nb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->ordinal);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &g : ctx->step_graphs) cudaGraphExecDestroy(g.second);
    for (auto &g : ctx->front_graphs) cudaGraphExecDestroy(g.second.exec);
    for (void *p : ctx->allocs) cudaFree(p);
    for (auto &r : ctx->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->host_pinned) cudaFreeHost(ctx->host_pinned);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_mel[i]) cudaEventDestroy(ctx->ev_mel[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int nb200_query(nb200_ctx *ctx, nb200_query_key key, int64_t *out) {
    if (!ctx || !out) return nb200_fail(ctx, NB200_INVALID_ARG, "query: NULL argument");
    switch (key) {
        case NB200_Q_N_FRAMES: *out = N_FRAMES; break;
        case NB200_Q_ENC_LEN: *out = ctx->cfg.max_source_positions; break;
        case NB200_Q_D_MODEL: *out = ctx->cfg.d_model; break;
        case NB200_Q_VOCAB: *out = ctx->cfg.vocab_size; break;
        case NB200_Q_MAX_BATCH: *out = ctx->cfg.max_batch; break;
        case NB200_Q_KERNEL_LAUNCHES: *out = ctx->launches; break;
        case NB200_Q_DEVICE_BYTES: *out = (int64_t)ctx->device_bytes; break;
        case NB200_Q_COMPUTE_DTYPE: *out = ctx->compute; break;
        case NB200_Q_MAX_TARGET_POSITIONS: *out = ctx->cfg.max_target_positions; break;
        default: return nb200_fail(ctx, NB200_INVALID_ARG, "query: unknown key %d", (int)key);
    }
    return NB200_OK;
}

int nb200_sync(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_load_tensor(nb200_ctx *ctx, const char *hf_name, const void *host, nb200_dtype dtype, const int64_t *shape, int rank) {
    if (!ctx || !hf_name || !host || !shape || rank < 1 || rank > 4) return nb200_fail(ctx, NB200_INVALID_ARG, "load_tensor: bad argument");
    if (ctx->finalized) return nb200_fail(ctx, NB200_INVALID_ARG, "load_tensor: weights already finalized");
    size_t n = 1;
    for (int i = 0; i < rank; ++i) {
        if (shape[i] <= 0) return nb200_fail(ctx, NB200_INVALID_ARG, "load_tensor: non-positive dimension");
        n *= (size_t)shape[i];
    }
    if (!strcmp(hf_name, "model.encoder.embed_positions.weight")) return NB200_OK;  // candle recomputes sinusoids
    HostTensor t;
    t.shape.assign(shape, shape + rank);
    t.data.resize(n);
    switch (dtype) {
        case NB200_F32: memcpy(t.data.data(), host, n * 4); break;
        case NB200_BF16: for (size_t i = 0; i < n; ++i) t.data[i] = bf16_bits_to_float(((const uint16_t *)host)[i]); break;
        case NB200_F16: for (size_t i = 0; i < n; ++i) t.data[i] = f16_bits_to_float(((const uint16_t *)host)[i]); break;
        case NB200_F64: for (size_t i = 0; i < n; ++i) t.data[i] = (float)((const double *)host)[i]; break;
        default: return nb200_fail(ctx, NB200_INVALID_ARG, "load_tensor: unsupported dtype %d for '%s'", (int)dtype, hf_name);
    }
    ctx->host_tensors[hf_name] = std::move(t);
    return NB200_OK;
}

int nb200_finalize_weights(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    if (ctx->finalized) return nb200_fail(ctx, NB200_INVALID_ARG, "finalize_weights: already finalized");
    const nb200_config &c = ctx->cfg;
    const int d = c.d_model, nm = c.num_mel_bins;
    const std::string E = "model.encoder.", D = "model.decoder.";
    NB_TRY(up_conv(ctx, E + "conv1.weight", d, nm, &ctx->conv1_w));
    NB_TRY(up_vec(ctx, E + "conv1.bias", d, &ctx->conv1_b));
    NB_TRY(up_conv(ctx, E + "conv2.weight", d, d, &ctx->conv2_w));
    NB_TRY(up_vec(ctx, E + "conv2.bias", d, &ctx->conv2_b));
    {   // candle `sinusoids` in f32: [sin | cos] halves
        const int half = d / 2, T = c.max_source_positions;
        std::vector<float> pos((size_t)T * d);
        const float inc = logf(10000.0f) / (float)(half - 1);
        std::vector<float> inv(half);
        for (int i = 0; i < half; ++i) inv[i] = expf((float)i * (-inc));
        for (int t = 0; t < T; ++t)
            for (int i = 0; i < half; ++i) {
                float a = (float)t * inv[i];
                pos[(size_t)t * d + i] = sinf(a);
                pos[(size_t)t * d + half + i] = cosf(a);
            }
        NB_TRY(upload_f32(ctx, pos, &ctx->pos));
    }
    // LayerNorm folded into the encoder's QKV / fc1 GEMMs: the tcgen05 path only, on shapes whose N = d_model tiles are whole
    ctx->ln_fold = ctx->compute == NB200_BF16 && ctx->opt.ln_fused && ctx->opt.gemm_epi_tma && (d % 256 == 0 || (d % 128 == 0 && d <= 1024)) &&
                   2 * (d / 128) <= LN_MAX_SLOTS;
    if (ctx->ln_fold) {
        const size_t rows = (size_t)c.max_batch * c.max_source_positions;
        NB_TRY(dev_alloc_t(ctx, (size_t)LN_MAX_SLOTS * rows, &ctx->ln_stats, true));
    }
    ctx->enc.resize(c.encoder_layers);
    for (int l = 0; l < c.encoder_layers; ++l) {
        const std::string p = E + "layers." + std::to_string(l) + ".";
        EncLayer &w = ctx->enc[l];
        if (ctx->ln_fold) {
            NB_TRY(up_qkv(ctx, p + "self_attn.", d, &w.wqkv, &w.bqkv, p + "self_attn_layer_norm"));
            const HostTensor *w1 = nullptr, *b1 = nullptr;
            NB_TRY(need(ctx, p + "fc1.weight", (size_t)4 * d * d, &w1));
            NB_TRY(need(ctx, p + "fc1.bias", 4 * d, &b1));
            std::vector<float> W = w1->data, Bv = b1->data;
            NB_TRY(fold_layernorm(ctx, p + "final_layer_norm", 4 * d, d, W, Bv));
            NB_TRY(upload_compute(ctx, W, &w.w1));
            NB_TRY(upload_f32(ctx, Bv, &w.b1));
        } else {
            NB_TRY(up_qkv(ctx, p + "self_attn.", d, &w.wqkv, &w.bqkv));
            NB_TRY(up_mat(ctx, p + "fc1.weight", (size_t)4 * d * d, &w.w1));
            NB_TRY(up_vec(ctx, p + "fc1.bias", 4 * d, &w.b1));
        }
        NB_TRY(up_mat(ctx, p + "self_attn.out_proj.weight", (size_t)d * d, &w.wo));
        NB_TRY(up_vec(ctx, p + "self_attn.out_proj.bias", d, &w.bo));
        NB_TRY(up_vec(ctx, p + "self_attn_layer_norm.weight", d, &w.ln1g));
        NB_TRY(up_vec(ctx, p + "self_attn_layer_norm.bias", d, &w.ln1b));
        NB_TRY(up_mat(ctx, p + "fc2.weight", (size_t)4 * d * d, &w.w2));
        NB_TRY(up_vec(ctx, p + "fc2.bias", d, &w.b2));
        NB_TRY(up_vec(ctx, p + "final_layer_norm.weight", d, &w.ln2g));
        NB_TRY(up_vec(ctx, p + "final_layer_norm.bias", d, &w.ln2b));
    }
    NB_TRY(up_vec(ctx, E + "layer_norm.weight", d, &ctx->lnpost_g));
    NB_TRY(up_vec(ctx, E + "layer_norm.bias", d, &ctx->lnpost_b));
    // decoder is optional (encoder-only throughput configs): present iff embed_tokens was loaded
    ctx->has_decoder = c.decoder_layers > 0 && find(ctx, D + "embed_tokens.weight") != nullptr;
    if (ctx->has_decoder) {
        NB_TRY(alloc_decoder(ctx));
        NB_TRY(up_mat(ctx, D + "embed_tokens.weight", (size_t)c.vocab_size * d, &ctx->embed));
        NB_TRY(up_vec(ctx, D + "embed_positions.weight", (size_t)c.max_target_positions * d, &ctx->embed_pos));
        ctx->dec.resize(c.decoder_layers);
        for (int l = 0; l < c.decoder_layers; ++l) {
            const std::string p = D + "layers." + std::to_string(l) + ".";
            DecLayer &w = ctx->dec[l];
            NB_TRY(up_qkv(ctx, p + "self_attn.", d, &w.wqkv, &w.bqkv));
            NB_TRY(up_mat(ctx, p + "self_attn.out_proj.weight", (size_t)d * d, &w.wo));
            NB_TRY(up_vec(ctx, p + "self_attn.out_proj.bias", d, &w.bo));
            NB_TRY(up_vec(ctx, p + "self_attn_layer_norm.weight", d, &w.ln1g));
            NB_TRY(up_vec(ctx, p + "self_attn_layer_norm.bias", d, &w.ln1b));
            NB_TRY(up_mat(ctx, p + "encoder_attn.q_proj.weight", (size_t)d * d, &w.cwq));
            NB_TRY(up_vec(ctx, p + "encoder_attn.q_proj.bias", d, &w.cbq));
            {   // [k | v] concatenation for the one-shot cross K/V GEMM
                const HostTensor *k, *v, *bv;
                NB_TRY(need(ctx, p + "encoder_attn.k_proj.weight", (size_t)d * d, &k));
                NB_TRY(need(ctx, p + "encoder_attn.v_proj.weight", (size_t)d * d, &v));
                NB_TRY(need(ctx, p + "encoder_attn.v_proj.bias", d, &bv));
                std::vector<float> W((size_t)2 * d * d), Bv(2 * d, 0.f);
                memcpy(W.data(), k->data.data(), (size_t)d * d * 4);
                memcpy(W.data() + (size_t)d * d, v->data.data(), (size_t)d * d * 4);
                memcpy(Bv.data() + d, bv->data.data(), d * 4);
                NB_TRY(upload_compute(ctx, W, &w.cwkv));
                NB_TRY(upload_f32(ctx, Bv, &w.cbkv));
            }
            NB_TRY(up_mat(ctx, p + "encoder_attn.out_proj.weight", (size_t)d * d, &w.cwo));
            NB_TRY(up_vec(ctx, p + "encoder_attn.out_proj.bias", d, &w.cbo));
            NB_TRY(up_vec(ctx, p + "encoder_attn_layer_norm.weight", d, &w.lncg));
            NB_TRY(up_vec(ctx, p + "encoder_attn_layer_norm.bias", d, &w.lncb));
            NB_TRY(up_mat(ctx, p + "fc1.weight", (size_t)4 * d * d, &w.w1));
            NB_TRY(up_vec(ctx, p + "fc1.bias", 4 * d, &w.b1));
            NB_TRY(up_mat(ctx, p + "fc2.weight", (size_t)4 * d * d, &w.w2));
            NB_TRY(up_vec(ctx, p + "fc2.bias", d, &w.b2));
            NB_TRY(up_vec(ctx, p + "final_layer_norm.weight", d, &w.ln2g));
            NB_TRY(up_vec(ctx, p + "final_layer_norm.bias", d, &w.ln2b));
        }
        NB_TRY(up_vec(ctx, D + "layer_norm.weight", d, &ctx->lndec_g));
        NB_TRY(up_vec(ctx, D + "layer_norm.bias", d, &ctx->lndec_b));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_dec_layers, ctx->dec.data(), ctx->dec.size() * sizeof(DecLayer), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        NB_TRY(upload_suppress(ctx));
    }
    ctx->host_tensors.clear();
    ctx->finalized = true;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_set_mel_filters(nb200_ctx *ctx, const float *filters, int n_mel) {
    NB_TRY(check_ready(ctx, false, false));
    if (!filters) return nb200_fail(ctx, NB200_INVALID_ARG, "set_mel_filters: NULL filters");
    if (n_mel != ctx->cfg.num_mel_bins)
        return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "Unexpected number of mel bins (num_mel_bins), got: %d", n_mel);
    NB_TRY(mel_setup_filters(ctx, filters, n_mel));
    ctx->has_filters = true;
    return NB200_OK;
}

int nb200_set_tokens(nb200_ctx *ctx, const nb200_special_tokens *tok) {
    NB_TRY(check_ready(ctx, false, false));
    if (!tok) return nb200_fail(ctx, NB200_INVALID_ARG, "set_tokens: NULL");
    const uint32_t V = (uint32_t)ctx->cfg.vocab_size;
    if (tok->sot >= V || tok->eot >= V || tok->task >= V || tok->no_speech >= V || tok->no_timestamps >= V || tok->ts_zero >= V || tok->ts_one >= V ||
        (tok->lang != UINT32_MAX && tok->lang >= V) || tok->ts_zero > tok->ts_one)
        return nb200_fail(ctx, NB200_INVALID_ARG, "set_tokens: token id out of range for vocab %u", V);
    ctx->tok = *tok;
    ctx->has_tokens = true;
    destroy_step_graphs(ctx);  // the captured select kernels hold the token ids by value
    return upload_suppress(ctx);
}

int nb200_set_suppress(nb200_ctx *ctx, const uint32_t *ids, size_t n) {
    NB_TRY(check_ready(ctx, false, false));
    if (n && !ids) return nb200_fail(ctx, NB200_INVALID_ARG, "set_suppress: NULL ids");
    ctx->suppress_ids.assign(ids, ids + n);
    return upload_suppress(ctx);
}

int nb200_pcm_to_mel_batch(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens, float *mel_out) {
    NB_TRY(check_ready(ctx, false, true));
    NB_TRY(stage_pcm(ctx, pcm, n_windows, stride, lens));
    NB_TRY(launch_mel(ctx, (int)n_windows));
    NB_TRY(launch_mel_norm(ctx, (int)n_windows));
    if (mel_out)
        CUDA_TRY(ctx, cudaMemcpyAsync(mel_out, ctx->mel_norm, n_windows * ctx->cfg.num_mel_bins * N_FRAMES * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_pcm_to_mel(nb200_ctx *ctx, const float *pcm, size_t n, float *mel_out, size_t *n_len) {
    NB_TRY(check_ready(ctx, false, true));
    if (n > (size_t)N_SAMPLES) return nb200_fail(ctx, NB200_INVALID_ARG, "pcm_to_mel: %zu samples exceed one window (%d)", n, N_SAMPLES);
    // candle frame-count rule (SURVEY §8 c-1 rule 3)
    size_t nl = n / HOP;
    if (nl % MEL_PAD_FRAMES != 0) nl = (nl / MEL_PAD_FRAMES + 1) * MEL_PAD_FRAMES;
    nl += MEL_PAD_FRAMES;
    if (n_len) *n_len = nl;
    if (!mel_out && !pcm) return NB200_OK;  // pure size query
    static const float zero = 0.f;
    size_t len1 = n;
    NB_TRY(stage_pcm(ctx, n ? pcm : &zero, 1, n ? n : 1, &len1));
    NB_TRY(launch_mel(ctx, 1));
    NB_TRY(launch_mel_norm(ctx, 1));
    if (mel_out) {
        // frames [0, min(3000, nl)) are computed; frames beyond the data are the normalised value of log10(1e-10)
        const int nm = ctx->cfg.num_mel_bins;
        const size_t keep = std::min<size_t>(N_FRAMES, nl);
        NB_TRY(ensure_pinned(ctx, (size_t)nm * N_FRAMES * 4 + 16));
        float *hp = (float *)ctx->host_pinned;
        unsigned *hmax = (unsigned *)(hp + (size_t)nm * N_FRAMES);
        CUDA_TRY(ctx, cudaMemcpyAsync(hp, ctx->mel_norm, (size_t)nm * N_FRAMES * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(hmax, ctx->mel_max, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        unsigned u = *hmax;
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        float mmax;
        memcpy(&mmax, &u, 4);
        const float padv = std::max(-10.0f, mmax - 8.0f) / 4.0f + 1.0f;
        for (int m = 0; m < nm; ++m) {
            memcpy(mel_out + (size_t)m * nl, hp + (size_t)m * N_FRAMES, keep * 4);
            for (size_t f = keep; f < nl; ++f) mel_out[(size_t)m * nl + f] = padv;
        }
    } else {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return NB200_OK;
}

int nb200_encoder_forward(nb200_ctx *ctx, const float *mel, size_t n_windows, float *out) {
    NB_TRY(check_ready(ctx, true, false));
    if (n_windows == 0 || n_windows > (size_t)ctx->cfg.max_batch)
        return nb200_fail(ctx, NB200_INVALID_ARG, "encoder_forward: n_windows=%zu (max_batch %d)", n_windows, ctx->cfg.max_batch);
    if (mel) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->mel_norm, mel, n_windows * ctx->cfg.num_mel_bins * N_FRAMES * 4, cudaMemcpyHostToDevice, ctx->stream));
        NB_TRY(launch_mel_from_host_layout(ctx, (int)n_windows));
    }
    NB_TRY(run_front(ctx, (int)n_windows, false, true, 0));
    if (out)
        CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->enc_out, n_windows * ctx->cfg.max_source_positions * ctx->cfg.d_model * 4, cudaMemcpyDeviceToHost,
                                      ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_transcode_batch(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens, float *out) {
    NB_TRY(check_ready(ctx, true, true));
    NB_TRY(stage_pcm(ctx, pcm, n_windows, stride, lens));
    NB_TRY(run_front(ctx, (int)n_windows, true, true, 0));
    if (out)
        CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->enc_out, n_windows * ctx->cfg.max_source_positions * ctx->cfg.d_model * 4, cudaMemcpyDeviceToHost,
                                      ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

// ---- pipelined batches: batch k+1's PCM goes up and batch k-1's features come down while batch k computes --------------
static int pipe_init(nb200_ctx *ctx) {
    if (ctx->h2d_stream) return NB200_OK;
    const size_t B = ctx->cfg.max_batch, per_out = (size_t)ctx->cfg.max_source_positions * ctx->cfg.d_model;
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    ctx->pipe_pcm[0] = ctx->pcm;  // slot 0 reuses the synchronous path's buffers
    ctx->pipe_len[0] = ctx->pcm_len;
    ctx->pipe_out[0] = ctx->enc_out;
    NB_TRY(dev_alloc_t(ctx, B * N_SAMPLES, &ctx->pipe_pcm[1], true));
    NB_TRY(dev_alloc_t(ctx, B, &ctx->pipe_len[1], true));
    NB_TRY(dev_alloc_t(ctx, B * per_out, &ctx->pipe_out[1]));
    for (int s = 0; s < 2; ++s) {
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d[s], cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_mel[s], cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[s], cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_d2h[s], cudaEventDisableTiming));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_transcode_submit(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens, float *out) {
    NB_TRY(check_ready(ctx, true, true));
    if (!pcm || !out || n_windows == 0 || n_windows > (size_t)ctx->cfg.max_batch)
        return nb200_fail(ctx, NB200_INVALID_ARG, "transcode_submit: n_windows=%zu (max_batch %d)", n_windows, ctx->cfg.max_batch);
    if (ctx->pipe_submitted - ctx->pipe_collected >= 2)
        return nb200_fail(ctx, NB200_INVALID_ARG, "transcode_submit: two batches are already in flight; call nb200_transcode_collect first");
    NB_TRY(pipe_init(ctx));
    const int s = (int)(ctx->pipe_submitted & 1);
    std::vector<int> &l = ctx->pipe_lens_host[s];
    l.resize(n_windows);
    for (size_t w = 0; w < n_windows; ++w) {
        const size_t n = lens ? lens[w] : stride;
        if (n > (size_t)N_SAMPLES) return nb200_fail(ctx, NB200_INVALID_ARG, "window %zu has %zu samples (> %d)", w, n, N_SAMPLES);
        l[w] = (int)n;
    }
    // H2D on its own stream, after the mel kernel that last read this slot's PCM (two submits ago)
    if (ctx->pipe_submitted >= 2) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_mel[s], 0));
    for (size_t w = 0; w < n_windows; ++w)
        if (l[w]) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pipe_pcm[s] + w * N_SAMPLES, pcm + w * stride, (size_t)l[w] * 4, cudaMemcpyHostToDevice, ctx->h2d_stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pipe_len[s], l.data(), n_windows * 4, cudaMemcpyHostToDevice, ctx->h2d_stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_h2d[s], ctx->h2d_stream));
    // compute: needs the PCM, and the feature buffer of this slot must have been drained by the D2H of two submits ago
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[s], 0));
    if (ctx->pipe_submitted >= 2) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[s], 0));
    float *save_pcm = ctx->pcm, *save_out = ctx->enc_out;
    int *save_len = ctx->pcm_len;
    ctx->pcm = ctx->pipe_pcm[s];
    ctx->pcm_len = ctx->pipe_len[s];
    ctx->enc_out = ctx->pipe_out[s];
    if (ctx->compute == NB200_F32) ctx->enc_out_c = ctx->enc_out;
    int st = run_front(ctx, (int)n_windows, true, true, s);
    if (st == NB200_OK) cudaEventRecord(ctx->ev_mel[s], ctx->stream);  // this slot's PCM has been consumed (its next upload is two submits away)
    ctx->pcm = save_pcm;
    ctx->pcm_len = save_len;
    ctx->enc_out = save_out;
    if (ctx->compute == NB200_F32) ctx->enc_out_c = ctx->enc_out;
    if (s != 0) ctx->n_resident = 0;  // the resident-features slot (decoder input) is slot 0 only
    if (st != NB200_OK) return st;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_comp[s], ctx->stream));
    // D2H on its own stream
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_comp[s], 0));
    CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->pipe_out[s], n_windows * ctx->cfg.max_source_positions * ctx->cfg.d_model * 4, cudaMemcpyDeviceToHost,
                                  ctx->d2h_stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_d2h[s], ctx->d2h_stream));
    ctx->pipe_submitted++;
    return NB200_OK;
}

int nb200_transcode_collect(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    if (ctx->pipe_collected >= ctx->pipe_submitted) return nb200_fail(ctx, NB200_INVALID_ARG, "transcode_collect: nothing in flight");
    const int s = (int)(ctx->pipe_collected & 1);
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_d2h[s]));
    ctx->pipe_collected++;
    return NB200_OK;
}

int nb200_stage_pcm(nb200_ctx *ctx, const float *pcm, size_t n_windows, size_t stride, const size_t *lens) {
    NB_TRY(check_ready(ctx, false, false));
    return stage_pcm(ctx, pcm, n_windows, stride, lens);
}

int nb200_run_resident(nb200_ctx *ctx, size_t n_windows, int do_mel, int do_encoder) {
    NB_TRY(check_ready(ctx, do_encoder != 0, do_mel != 0));
    if (n_windows == 0 || n_windows > (size_t)ctx->cfg.max_batch) return nb200_fail(ctx, NB200_INVALID_ARG, "run_resident: n_windows=%zu", n_windows);
    NB_TRY(run_front(ctx, (int)n_windows, do_mel != 0, do_encoder != 0, 0));
    return NB200_OK;  // asynchronous on the ctx stream: pair with nb200_timer_stop / nb200_sync
}

int nb200_fetch_features(nb200_ctx *ctx, size_t window, float *out, size_t n) {
    NB_TRY(check_ready(ctx, true, false));
    const size_t per = (size_t)ctx->cfg.max_source_positions * ctx->cfg.d_model;
    if (!out || window >= (size_t)ctx->cfg.max_batch || n > per) return nb200_fail(ctx, NB200_INVALID_ARG, "fetch_features: bad argument");
    CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->enc_out + window * per, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_fetch_mel(nb200_ctx *ctx, size_t window, float *out, size_t n) {
    NB_TRY(check_ready(ctx, false, false));
    const size_t per = (size_t)ctx->cfg.num_mel_bins * N_FRAMES;
    if (!out || window >= (size_t)ctx->cfg.max_batch || n > per) return nb200_fail(ctx, NB200_INVALID_ARG, "fetch_mel: bad argument");
    CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->mel_norm + window * per, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

static int check_decoder(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, true, false));
    if (!ctx->has_decoder) return nb200_fail(ctx, NB200_NOT_LOADED, "decoder weights were not loaded");
    return NB200_OK;
}

// `audio_features` as an argument (model.rs:279, 466: decode / decoder_forward take `xa`): host [n_windows][1500][d_model] f32 become the
// resident features of windows [0, n_windows), as if the encoder had produced them
int nb200_set_audio_features(nb200_ctx *ctx, const float *xa, size_t n_windows) {
    NB_TRY(check_ready(ctx, true, false));
    if (!xa || n_windows == 0 || n_windows > (size_t)ctx->cfg.max_batch)
        return nb200_fail(ctx, NB200_INVALID_ARG, "set_audio_features: n_windows=%zu (max_batch %d)", n_windows, ctx->cfg.max_batch);
    const size_t n = n_windows * ctx->cfg.max_source_positions * ctx->cfg.d_model;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->enc_out, xa, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->compute == NB200_BF16) NB_TRY(launch_f32_to_bf16(ctx, ctx->enc_out, (bf16 *)ctx->enc_out_c, n));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_resident = (int)n_windows;
    ctx->cross_valid = false;
    ctx->seam_window = -1;
    return NB200_OK;
}

int nb200_decoder_forward(nb200_ctx *ctx, size_t window, const uint32_t *tokens, size_t n, int flush, float *hidden_out) {
    NB_TRY(check_decoder(ctx));
    const nb200_config &c = ctx->cfg;
    if (!tokens || n == 0 || n > (size_t)c.max_target_positions || window >= (size_t)c.max_batch)
        return nb200_fail(ctx, NB200_INVALID_ARG, "decoder_forward: n=%zu window=%zu", n, window);
    if ((int)window >= ctx->n_resident) return nb200_fail(ctx, NB200_NOT_LOADED, "decoder_forward: window %zu has no resident audio features", window);
    for (size_t i = 0; i < n; ++i)
        if (tokens[i] >= (uint32_t)c.vocab_size) return nb200_fail(ctx, NB200_INVALID_ARG, "decoder_forward: token %u out of vocab", tokens[i]);
    const int P = c.max_target_positions, d = c.d_model;
    const bool rebuild = flush || !ctx->cross_valid;
    if (rebuild) {
        NB_TRY(decoder_build_cross_kv(ctx, ctx->n_resident));
        ctx->seam_window = -1;
    }
    if (ctx->run.active) ctx->run.active = false;  // a decode in progress shares the self-attention cache: it is abandoned
    if (!ctx->seam_hidden) NB_TRY(dev_alloc_t(ctx, (size_t)P * d, &ctx->seam_hidden));
    // The reference recomputes every position on every call (candle keeps no self-attention cache), so norma's loop costs O(n^2) steps per
    // window.  Here the self-attention K/V rows and the hidden rows of the positions computed by the previous call are still valid when
    // this call extends the same token prefix with the same cross K/V (flush = 0): only the new positions run.
    size_t first = 0;
    if (ctx->seam_window == (int)window && ctx->seam_tokens.size() <= n && std::equal(ctx->seam_tokens.begin(), ctx->seam_tokens.end(), tokens))
        first = ctx->seam_tokens.size();
    ctx->seam_window = -1;  // until this call has completed
    int ln = (int)n;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_tokens + window * P, tokens, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_len + window, &ln, 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `tokens` and `ln` are consumed
    for (size_t pos = first; pos < n; ++pos) {
        NB_TRY(decoder_step(ctx, (int)window, 1, (int)pos, 0));
        NB_TRY(decoder_copy_hidden(ctx, ctx->seam_hidden + pos * d, d));
    }
    if (hidden_out) CUDA_TRY(ctx, cudaMemcpyAsync(hidden_out, ctx->seam_hidden, n * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->seam_window = (int)window;
    ctx->seam_tokens.assign(tokens, tokens + n);
    return NB200_OK;
}

int nb200_final_linear(nb200_ctx *ctx, const float *hidden, float *logits_out) {
    NB_TRY(check_decoder(ctx));
    if (!logits_out) return nb200_fail(ctx, NB200_INVALID_ARG, "final_linear: NULL output");
    const int d = ctx->cfg.d_model, V = ctx->cfg.vocab_size;
    if (hidden) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dhid, hidden, d * 4, cudaMemcpyHostToDevice, ctx->stream));
    // logits row 0 = dhid row 0 . embed^T (tied embedding, no bias)
    {
        // reuse the step's logits path for one row
        NB_TRY(decoder_logits_rows(ctx, 1));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(logits_out, ctx->logits, (size_t)V * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_detect_language(nb200_ctx *ctx, size_t window, const uint32_t *lang_tokens, size_t n_langs, uint32_t *token_out, float *probs_out) {
    NB_TRY(check_decoder(ctx));
    const nb200_config &c = ctx->cfg;
    if (!ctx->has_tokens) return nb200_fail(ctx, NB200_NOT_LOADED, "special tokens not set (call nb200_set_tokens)");
    if (!lang_tokens || n_langs == 0 || n_langs > (size_t)NB200_MAX_LANGS || !token_out || window >= (size_t)c.max_batch)
        return nb200_fail(ctx, NB200_INVALID_ARG, "detect_language: n_langs=%zu window=%zu", n_langs, window);
    if ((int)window >= ctx->n_resident) return nb200_fail(ctx, NB200_NOT_LOADED, "detect_language: window %zu has no resident audio features", window);
    for (size_t i = 0; i < n_langs; ++i)
        if (lang_tokens[i] >= (uint32_t)c.vocab_size) return nb200_fail(ctx, NB200_INVALID_ARG, "detect_language: token %u out of vocab", lang_tokens[i]);
    // `decoder_forward([[sot]], audio_features, flush = true)` then `final_linear(ys[..1])` (model.rs:195-197)
    NB_TRY(decoder_build_cross_kv(ctx, ctx->n_resident));
    ctx->seam_window = -1;
    ctx->run.active = false;
    const int P = c.max_target_positions;
    const int one = 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_tokens + window * P, &ctx->tok.sot, 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_len + window, &one, 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_lang, lang_tokens, n_langs * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // the three sources above are pageable / stack memory
    NB_TRY(decoder_step(ctx, (int)window, 1, 0, 1));
    NB_TRY(decoder_language(ctx, (int)n_langs));
    int best = -1;
    const float *d_probs = (const float *)((const uint32_t *)ctx->d_lang + NB200_MAX_LANGS);
    CUDA_TRY(ctx, cudaMemcpyAsync(&best, d_probs + NB200_MAX_LANGS, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (probs_out) CUDA_TRY(ctx, cudaMemcpyAsync(probs_out, d_probs, n_langs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (best < 0 || best >= (int)n_langs) return nb200_fail(ctx, NB200_CUDA_ERROR, "detect_language: no maximum (non-finite logits)");
    *token_out = lang_tokens[best];
    return NB200_OK;
}

int nb200_reset_kv_cache(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    ctx->cross_valid = false;
    return NB200_OK;
}

int nb200_set_decode_mode(nb200_ctx *ctx, nb200_decode_mode mode) {
    NB_TRY(check_ready(ctx, false, false));
    if (mode != NB200_DECODE_AUTO && mode != NB200_DECODE_SEPARATE) return nb200_fail(ctx, NB200_INVALID_ARG, "set_decode_mode: %d", (int)mode);
    ctx->decode_separate = mode == NB200_DECODE_SEPARATE;
    return NB200_OK;
}

int nb200_decode_greedy(nb200_ctx *ctx, size_t n_windows, size_t max_new_tokens, uint32_t *tokens_out, size_t *n_tokens, double *avg_logprob,
                        double *no_speech_prob) {
    return nb200_decode(ctx, n_windows, 0.0f, 0, max_new_tokens, tokens_out, n_tokens, avg_logprob, no_speech_prob);
}

// ---- norma's `Model::decode` (model.rs:279-390) opened up into begin / advance / end so that a caller (and the parity tests) can look at
// every step's logits, exactly what `decoder_final_linear` returns inside the reference loop (model.rs:324-329) --------------------------
static int ensure_step_graph(nb200_ctx *ctx) {
    nb200_ctx::DecodeRun &r = ctx->run;
    if (r.gexec || ctx->profiling || !ctx->opt.decode_graph) return NB200_OK;
    const int gkey = r.B * 2 + r.greedy;
    auto it = ctx->step_graphs.find(gkey);
    if (it != ctx->step_graphs.end()) { r.gexec = it->second; return NB200_OK; }
    cudaGraph_t graph = nullptr;
    CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int st = decoder_step(ctx, 0, r.B, -1, 1);
    if (st == NB200_OK) st = decoder_select(ctx, r.B, r.greedy);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (st != NB200_OK || ce != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (st != NB200_OK) return st;
        return nb200_fail(ctx, NB200_CUDA_ERROR, "decode graph capture failed: %s", cudaGetErrorString(ce));
    }
    cudaGraphExec_t gexec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return nb200_fail(ctx, NB200_CUDA_ERROR, "decode graph instantiate failed: %s", cudaGetErrorString(ie));
    ctx->step_graphs[gkey] = gexec;
    r.gexec = gexec;
    return NB200_OK;
}

int nb200_decode_begin(nb200_ctx *ctx, size_t n_windows, float temperature, uint64_t seed, size_t max_new_tokens) {
    NB_TRY(check_decoder(ctx));
    if (!(temperature >= 0.0f)) return nb200_fail(ctx, NB200_INVALID_ARG, "decode: temperature must be >= 0");
    const nb200_config &c = ctx->cfg;
    if (!ctx->has_tokens) return nb200_fail(ctx, NB200_NOT_LOADED, "special tokens not set (call nb200_set_tokens)");
    if (n_windows == 0 || (int)n_windows > ctx->n_resident)
        return nb200_fail(ctx, NB200_INVALID_ARG, "decode: n_windows=%zu but %d windows have resident audio features", n_windows, ctx->n_resident);
    nb200_ctx::DecodeRun &r = ctx->run;
    r = nb200_ctx::DecodeRun{};
    ctx->seam_window = -1;  // the self-attention caches are about to be overwritten
    const int B = (int)n_windows;
    const int plen = ctx->tok.lang != UINT32_MAX ? 3 : 2;
    // `decoder_forward(prompt, audio_features, flush = true)` (model.rs:297-299): rebuild the cross K/V cache
    NB_TRY(decoder_build_cross_kv(ctx, ctx->n_resident));
    NB_TRY(decoder_init_state(ctx, B));
    NB_TRY(decoder_set_dyn(ctx, 0, (int)max_new_tokens, temperature, seed, 1));
    for (int pos = 0; pos < plen; ++pos) {
        NB_TRY(decoder_step(ctx, 0, B, pos, pos == 0 || pos == plen - 1));
        if (pos == 0) NB_TRY(decoder_nospeech(ctx, B));
    }
    r.greedy = temperature == 0.0f ? 1 : 0;
    NB_TRY(decoder_select(ctx, B, r.greedy));  // first sampled token; advances the device position to plen
    r.B = B;
    r.plen = plen;
    r.pos = plen;
    // greedy steady state: up to 16 steps (embed .. logits .. select each) are one cooperative kernel; NB200_DECODE_FUSED=0 or
    // nb200_set_decode_mode fall back to the ~25 separate kernels per step replayed as a CUDA graph (the only path for t > 0, whose
    // sampler is a single-block kernel, and for F32 contexts)
    r.use_fused = r.greedy && ctx->opt.decode_fused && !ctx->decode_separate && !ctx->fused_failed && decoder_fused_supported(ctx) &&
                  B <= decoder_fused_max_windows() && (int)c.max_target_positions >= plen;
    if (r.use_fused && decoder_fused_prepare(ctx) != NB200_OK) r.use_fused = false;
    if (!r.use_fused) NB_TRY(ensure_step_graph(ctx));
    r.active = true;
    return NB200_OK;
}

// up to n_steps more positions for every window of the run (model.rs:317-371; the position, temperature and token budget live on the
// device).  The done flags are read back first: *all_done = 1 means every window had already finished and nothing was launched.
int nb200_decode_advance(nb200_ctx *ctx, size_t n_steps, int *all_done) {
    NB_TRY(check_decoder(ctx));
    nb200_ctx::DecodeRun &r = ctx->run;
    if (!r.active) return nb200_fail(ctx, NB200_INVALID_ARG, "decode_advance: no decode in progress (call nb200_decode_begin)");
    const int B = r.B, P = ctx->cfg.max_target_positions;
    std::vector<int> done(B);
    CUDA_TRY(ctx, cudaMemcpyAsync(done.data(), ctx->d_done, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    bool all = r.pos >= P;
    if (!all) {
        all = true;
        for (int b = 0; b < B; ++b) all &= done[b] != 0;
    }
    if (all_done) *all_done = all ? 1 : 0;
    if (all || n_steps == 0) return NB200_OK;
    int left = (int)std::min<size_t>(n_steps, (size_t)(P - r.pos));
    const bool batchable = r.use_fused || r.gexec != nullptr;
    while (left > 0) {
        const int n = batchable ? std::min(left, 16) : 1;
        if (r.use_fused) {  // one cooperative launch for the next n positions: the CTAs stay resident between steps
            if (decoder_step_fused(ctx, B, n) == NB200_OK) {
                r.pos += n;
                left -= n;
                continue;
            }
            // the cooperative launch was refused (e.g. the SMs are shared with another process and 148 CTAs cannot be co-resident):
            // nothing ran, the decoding state is untouched; this context uses the separate kernels from now on
            cudaGetLastError();
            ctx->fused_failed = true;
            r.use_fused = false;
            NB_TRY(ensure_step_graph(ctx));
        }
        for (int i = 0; i < n; ++i) {
            if (r.gexec) CUDA_TRY(ctx, cudaGraphLaunch(r.gexec, ctx->stream));
            else {
                NB_TRY(decoder_step(ctx, 0, B, -1, 1));
                NB_TRY(decoder_select(ctx, B, r.greedy));
            }
        }
        r.pos += n;
        left -= n;
    }
    return NB200_OK;
}

// logits [vocab] of `window` at the LAST computed position: after nb200_decode_begin the ones the first sampled token was chosen from,
// after k single-step advances the ones of sampled token k + 1 (what `decoder_final_linear` returned there, model.rs:324-329)
int nb200_decode_peek_logits(nb200_ctx *ctx, size_t window, float *logits_out) {
    NB_TRY(check_decoder(ctx));
    if (!ctx->run.active || (int)window >= ctx->run.B || !logits_out) return nb200_fail(ctx, NB200_INVALID_ARG, "decode_peek_logits: bad argument or no decode in progress");
    const size_t V = ctx->cfg.vocab_size;
    CUDA_TRY(ctx, cudaMemcpyAsync(logits_out, ctx->logits + window * V, V * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_decode_end(nb200_ctx *ctx, uint32_t *tokens_out, size_t *n_tokens, double *avg_logprob, double *no_speech_prob) {
    NB_TRY(check_decoder(ctx));
    nb200_ctx::DecodeRun &r = ctx->run;
    if (!r.active || !tokens_out || !n_tokens) return nb200_fail(ctx, NB200_INVALID_ARG, "decode_end: bad argument or no decode in progress");
    const int B = r.B, P = ctx->cfg.max_target_positions, plen = r.plen;
    r.active = false;
    std::vector<uint32_t> toks((size_t)B * P);
    std::vector<int> len(B), done(B);
    std::vector<double> slp(B);
    std::vector<float> nsp(B);
    CUDA_TRY(ctx, cudaMemcpyAsync(toks.data(), ctx->d_tokens, toks.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(len.data(), ctx->d_len, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(slp.data(), ctx->d_sumlp, B * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(nsp.data(), ctx->d_nospeech, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(done.data(), ctx->d_done, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t nts = ctx->tok.no_timestamps;
    for (int b = 0; b < B; ++b) {
        uint32_t *t = toks.data() + (size_t)b * P;
        int l = len[b];
        double avg;
        if (done[b] == 2) {  // no-speech early return: prompt only, avg_logprob = 0 (model.rs:308-315)
            l = plen;
            avg = 0.0;
        } else {
            avg = slp[b] / (double)l;                      // model.rs:373 (length before the strip)
            while (l >= 2 && t[l - 2] > nts) {             // model.rs:375-381
                memmove(t + l - 2, t + l - 1, 4);
                --l;
            }
        }
        memcpy(tokens_out + (size_t)b * P, t, (size_t)l * 4);
        for (int i = l; i < P; ++i) tokens_out[(size_t)b * P + i] = 0;
        n_tokens[b] = (size_t)l;
        if (avg_logprob) avg_logprob[b] = avg;
        if (no_speech_prob) no_speech_prob[b] = (double)nsp[b];
    }
    return NB200_OK;
}

int nb200_decode(nb200_ctx *ctx, size_t n_windows, float temperature, uint64_t seed, size_t max_new_tokens, uint32_t *tokens_out, size_t *n_tokens,
                 double *avg_logprob, double *no_speech_prob) {
    if (!tokens_out || !n_tokens) return nb200_fail(ctx, NB200_INVALID_ARG, "decode: NULL output");
    NB_TRY(nb200_decode_begin(ctx, n_windows, temperature, seed, max_new_tokens));
    // the host only polls the done flags every 16 steps (no per-token sync or transfer)
    for (;;) {
        int all = 0;
        const int st = nb200_decode_advance(ctx, 16, &all);
        if (st != NB200_OK) { ctx->run.active = false; return st; }
        if (all) break;
    }
    return nb200_decode_end(ctx, tokens_out, n_tokens, avg_logprob, no_speech_prob);
}

// ---- streaming front half (BASELINE config 4): the window-0 PCM lives on the device, small chunks are
// appended as they arrive and only the mel frames they touch are recomputed -------------------------------------------
int nb200_stream_reset(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, true));
    ctx->stream_len = 0;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->pcm, 0, (size_t)N_SAMPLES * 4, ctx->stream));
    return mel_stream_reset(ctx);
}

int nb200_stream_push(nb200_ctx *ctx, const float *chunk, size_t n) {
    NB_TRY(check_ready(ctx, false, true));
    if ((n && !chunk) || ctx->stream_len + n > (size_t)N_SAMPLES)
        return nb200_fail(ctx, NB200_INVALID_ARG, "stream_push: %zu samples do not fit the 30 s window (%d buffered)", n, ctx->stream_len);
    if (n == 0) return NB200_OK;
    const int n_old = ctx->stream_len, n_new = n_old + (int)n;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pcm + n_old, chunk, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pcm_len, &n_new, 4, cudaMemcpyHostToDevice, ctx->stream));
    // frame i = samples [160 i, 160 i + 400): frames that were still zero-filled at n_old change, frames starting before n_new appear
    const int f_lo = n_old >= N_FFT ? (n_old - N_FFT) / HOP + 1 : 0;
    const int f_hi = std::min(N_FRAMES, (n_new + HOP - 1) / HOP);
    NB_TRY(mel_stream_update(ctx, f_lo, f_hi));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `chunk` and `n_new` are consumed
    ctx->stream_len = n_new;
    return NB200_OK;
}

int nb200_stream_drain(nb200_ctx *ctx, size_t n) {
    NB_TRY(check_ready(ctx, false, true));
    if (n > (size_t)ctx->stream_len) return nb200_fail(ctx, NB200_INVALID_ARG, "stream_drain: %zu > %d buffered samples", n, ctx->stream_len);
    if (n == 0) return NB200_OK;
    const int n_new = ctx->stream_len - (int)n;
    if (!ctx->stream_tmp) NB_TRY(dev_alloc_t(ctx, (size_t)N_SAMPLES + (size_t)ctx->cfg.num_mel_bins * N_FRAMES, &ctx->stream_tmp));
    NB_TRY(mel_stream_shift(ctx, (int)n, n_new, ctx->stream_tmp, ctx->stream_tmp + N_SAMPLES));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pcm_len, &n_new, 4, cudaMemcpyHostToDevice, ctx->stream));
    // frames keep their values when the shift is a whole number of hops; otherwise (and for the frames that now reach past
    // the end) recompute
    const int f_lo = (n % HOP == 0) ? (n_new >= N_FFT ? (n_new - N_FFT) / HOP + 1 : 0) : 0;
    NB_TRY(mel_stream_update(ctx, f_lo, N_FRAMES));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream_len = n_new;
    return NB200_OK;
}

// log-mel of the buffered audio exactly as pcm_to_mel would return it for those samples (first 3000 frames), then the encoder
int nb200_stream_features(nb200_ctx *ctx, int run_encoder, float *mel_out, float *features_out) {
    NB_TRY(check_ready(ctx, run_encoder != 0, true));
    NB_TRY(mel_stream_window_max(ctx));
    NB_TRY(launch_mel_norm(ctx, 1));
    if (mel_out) CUDA_TRY(ctx, cudaMemcpyAsync(mel_out, ctx->mel_norm, (size_t)ctx->cfg.num_mel_bins * N_FRAMES * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (run_encoder) {
        NB_TRY(run_front(ctx, 1, false, true, 0));
        if (features_out)
            CUDA_TRY(ctx, cudaMemcpyAsync(features_out, ctx->enc_out, (size_t)ctx->cfg.max_source_positions * ctx->cfg.d_model * 4, cudaMemcpyDeviceToHost,
                                          ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int nb200_timer_start(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return NB200_OK;
}

int nb200_timer_stop(nb200_ctx *ctx, float *ms) {
    NB_TRY(check_ready(ctx, false, false));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_stop));
    if (ms) CUDA_TRY(ctx, cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return NB200_OK;
}

int nb200_profile_enable(nb200_ctx *ctx, int on) {
    NB_TRY(check_ready(ctx, false, false));
    ctx->profiling = on != 0;
    return NB200_OK;
}

static int profile_collect(nb200_ctx *ctx) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ctx->prof_ms[r.cls] += ms;
            auto &t = ctx->prof_by_tag[(long long)r.cls * 1000000000000LL + r.tag];
            t.first += ms;
            t.second += 1;
        }
        ctx->ev_pool.push_back(r.a);
        ctx->ev_pool.push_back(r.b);
    }
    ctx->prof_recs.clear();
    return NB200_OK;
}

int nb200_profile_read(nb200_ctx *ctx, float *ms_by_class, int64_t *launches_by_class, double *flops_gemm) {
    NB_TRY(check_ready(ctx, false, false));
    NB_TRY(profile_collect(ctx));
    for (int i = 0; i < NB200_K_COUNT; ++i) {
        if (ms_by_class) ms_by_class[i] = ctx->prof_ms[i];
        if (launches_by_class) launches_by_class[i] = ctx->prof_launches[i];
    }
    if (flops_gemm) *flops_gemm = ctx->prof_gemm_flops;
    if (ctx->opt.prof_dump) {
        for (auto &kv : ctx->prof_by_tag)
            fprintf(stderr, "[nb200 prof] class %lld tag %lld: %.3f ms over %lld launches (%.1f us each)\n", kv.first / 1000000000000LL,
                    kv.first % 1000000000000LL, kv.second.first, kv.second.second, 1e3 * kv.second.first / (double)kv.second.second);
    }
    return NB200_OK;
}

int nb200_profile_reset(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    NB_TRY(profile_collect(ctx));
    for (int i = 0; i < NB200_K_COUNT; ++i) { ctx->prof_ms[i] = 0.f; ctx->prof_launches[i] = 0; }
    ctx->prof_gemm_flops = 0;
    ctx->prof_by_tag.clear();
    return NB200_OK;
}

int nb200_flush_l2(nb200_ctx *ctx) {
    NB_TRY(check_ready(ctx, false, false));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    return NB200_OK;
}

int nb200_test_gemm(nb200_ctx *ctx, const void *a, const void *w, const float *bias, int M, int N, int K, int act_gelu, float *c_out) {
    // act_gelu: bit 0 = tanh-GELU; bit 1 (bf16 contexts) = take the bf16-output epilogue (the QKV / fc1 path, incl. the single-round wide
    // tiles at M <= 1536 and N > 1536) and widen the result to f32 on the way out
    NB_TRY(check_ready(ctx, false, false));
    if (!a || !w || !c_out || M < 1 || N < 1 || K < 1) return nb200_fail(ctx, NB200_INVALID_ARG, "test_gemm: bad argument");
    const size_t es = dtype_size(ctx->compute);
    const bool out16 = (act_gelu & 2) && ctx->compute == NB200_BF16;
    void *dA = nullptr, *dW = nullptr;
    float *dB = nullptr, *dC = nullptr;
    int st = [&]() -> int {
        CUDA_TRY(ctx, cudaMalloc(&dA, (size_t)M * K * es));
        CUDA_TRY(ctx, cudaMalloc(&dW, (size_t)N * K * es));
        CUDA_TRY(ctx, cudaMalloc(&dC, (size_t)M * N * 4));
        CUDA_TRY(ctx, cudaMemcpyAsync(dA, a, (size_t)M * K * es, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(dW, w, (size_t)N * K * es, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(dC, 0xff, (size_t)M * N * 4, ctx->stream));
        if (bias) {
            CUDA_TRY(ctx, cudaMalloc(&dB, (size_t)N * 4));
            CUDA_TRY(ctx, cudaMemcpyAsync(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        GemmShape s{M, 1, N, K, K, (long long)M * K};
        Epilogue e{};
        e.bias = dB; e.act = act_gelu & 1; e.out = dC; e.ldo = N; e.out_bf16 = out16 ? 1 : 0;
        if (ctx->compute == NB200_BF16) NB_TRY(launch_gemm_bf16(ctx, (const bf16 *)dA, (const bf16 *)dW, s, e));
        else NB_TRY(launch_gemm_f32(ctx, (const float *)dA, (const float *)dW, s, e));
        if (out16) {
            std::vector<uint16_t> h((size_t)M * N);
            CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), dC, h.size() * 2, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            for (size_t i = 0; i < h.size(); ++i) c_out[i] = bf16_bits_to_float(h[i]);
            return NB200_OK;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(c_out, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return NB200_OK;
    }();
    cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dC);
    return st;
}

int nb200_test_gemm_perf(nb200_ctx *ctx, int M, int N, int K, int epi_kind, int iters, float *ms_out) {
    NB_TRY(check_ready(ctx, false, false));
    if (ctx->compute != NB200_BF16 || !ms_out || iters < 1) return nb200_fail(ctx, NB200_INVALID_ARG, "test_gemm_perf: bf16 ctx required");
    bf16 *dA = nullptr, *dW = nullptr;
    float *dB = nullptr;
    void *dC = nullptr;
    int st = [&]() -> int {
        CUDA_TRY(ctx, cudaMalloc(&dA, (size_t)M * K * 2));
        CUDA_TRY(ctx, cudaMalloc(&dW, (size_t)N * K * 2));
        CUDA_TRY(ctx, cudaMalloc(&dB, (size_t)N * 4));
        CUDA_TRY(ctx, cudaMalloc(&dC, (size_t)M * N * 4));
        CUDA_TRY(ctx, cudaMemsetAsync(dA, 0, (size_t)M * K * 2, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(dW, 0, (size_t)N * K * 2, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(dB, 0, (size_t)N * 4, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(dC, 0, (size_t)M * N * 4, ctx->stream));
        GemmShape s{M, 1, N, K, K, (long long)M * K};
        Epilogue e{};
        e.bias = dB; e.out = dC; e.ldo = N;
        if (epi_kind == 0) { e.out_bf16 = 1; }
        else if (epi_kind == 1) { e.out_bf16 = 1; e.act = 1; }
        else { e.out_bf16 = 0; e.residual = (const float *)dC; e.ldr = N; }
        for (int i = 0; i < 3; ++i) NB_TRY(launch_gemm_bf16(ctx, dA, dW, s, e));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
        for (int i = 0; i < iters; ++i) NB_TRY(launch_gemm_bf16(ctx, dA, dW, s, e));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_stop));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        *ms_out = ms / iters;
        return NB200_OK;
    }();
    cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dC);
    return st;
}

int nb200_test_attention_perf(nb200_ctx *ctx, const float *qkv, int B, int T, int n_heads, int iters, float *ms_out) {
    NB_TRY(check_ready(ctx, false, false));
    if (ctx->compute != NB200_BF16 || !qkv || !ms_out || iters < 1) return nb200_fail(ctx, NB200_INVALID_ARG, "test_attention_perf: bf16 ctx required");
    const int d = n_heads * HEAD_DIM;
    const size_t nq = (size_t)B * T * 3 * d, no = (size_t)B * T * d;
    float *dq32 = nullptr;
    bf16 *dq16 = nullptr, *do16 = nullptr;
    int st = [&]() -> int {
        CUDA_TRY(ctx, cudaMalloc(&dq32, nq * 4));
        CUDA_TRY(ctx, cudaMalloc(&dq16, nq * 2));
        CUDA_TRY(ctx, cudaMalloc(&do16, no * 2));
        CUDA_TRY(ctx, cudaMemcpyAsync(dq32, qkv, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
        NB_TRY(launch_f32_to_bf16(ctx, dq32, dq16, nq));
        for (int i = 0; i < 3; ++i) NB_TRY(launch_attention_tc(ctx, dq16, do16, B, T, n_heads));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
        for (int i = 0; i < iters; ++i) NB_TRY(launch_attention_tc(ctx, dq16, do16, B, T, n_heads));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_stop));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        *ms_out = ms / iters;
        return NB200_OK;
    }();
    cudaFree(dq32); cudaFree(dq16); cudaFree(do16);
    return st;
}

int nb200_test_attention(nb200_ctx *ctx, const float *qkv, int B, int T, int n_heads, float *ctx_out) {
    NB_TRY(check_ready(ctx, false, false));
    if (!qkv || !ctx_out || B < 1 || T < 1 || n_heads < 1) return nb200_fail(ctx, NB200_INVALID_ARG, "test_attention: bad argument");
    const int d = n_heads * HEAD_DIM;
    const size_t nq = (size_t)B * T * 3 * d, no = (size_t)B * T * d;
    const bool bf = ctx->compute == NB200_BF16;
    float *dq32 = nullptr, *do32 = nullptr;
    bf16 *dq16 = nullptr, *do16 = nullptr;
    int st = [&]() -> int {
        CUDA_TRY(ctx, cudaMalloc(&dq32, nq * 4));
        CUDA_TRY(ctx, cudaMalloc(&do32, no * 4));
        CUDA_TRY(ctx, cudaMemcpyAsync(dq32, qkv, nq * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (bf) {
            CUDA_TRY(ctx, cudaMalloc(&dq16, nq * 2));
            CUDA_TRY(ctx, cudaMalloc(&do16, no * 2));
            NB_TRY(launch_f32_to_bf16(ctx, dq32, dq16, nq));
            if (!ctx->opt.attn_tc) NB_TRY(launch_attention_simt(ctx, dq16, do16, B, T, n_heads, 1));
            else NB_TRY(launch_attention_tc(ctx, dq16, do16, B, T, n_heads));
            std::vector<uint16_t> h(no);
            CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), do16, no * 2, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            for (size_t i = 0; i < no; ++i) ctx_out[i] = bf16_bits_to_float(h[i]);
        } else {
            NB_TRY(launch_attention_simt(ctx, dq32, do32, B, T, n_heads, 0));
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx_out, do32, no * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return NB200_OK;
    }();
    cudaFree(dq32); cudaFree(do32); cudaFree(dq16); cudaFree(do16);
    return st;
}

}  // extern "C"
