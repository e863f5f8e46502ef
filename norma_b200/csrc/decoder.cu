// decoder.cu — K7-K9: the KV-cached decoder step and norma's greedy decode loop kept on the device.
//   * cross-attention K/V of the audio features are built once per window (flush = true semantics of candle's
//     MultiHeadAttention cache) with the big GEMM kernels;
//   * every decoder position is ONE token per window: weight-streaming skinny GEMMs (HBM-bound), single-query
//     attention over the self K/V cache (the reference recomputes the whole prefix each step; a cache is
//     numerically equivalent) and over the cross K/V;
//   * tied-embedding logits + softmax + norma's suppression rules + arg-max (last index wins ties) + logprob
//     accumulation run in one kernel per step (`decode_select_kernel`), replacing the 4 D2H syncs and vocab-sized
//     transfers per token at /root/reference/src/models/whisper/model.rs:263-270, 350, 364.
// Replaces `Type::decoder_forward` / `decoder_final_linear` (model.rs:466-483) and `Model::decode` at t = 0
// (model.rs:279-390) with the rules of model.rs:212-277 and the masks of monolingual.rs:386-430.
#include "common.cuh"

namespace {

struct SkinnyEpi {
    const float *bias;
    const float *residual;  // [Bd][ldr] or nullptr (may alias out)
    float *out;             // [Bd][ldo]
    int ldr, ldo;
    float scale;
    int n_scale;
    int act;
};

constexpr int SK_KC = 1024;
constexpr int SK_MB = 8;

template <typename WT>
__device__ __forceinline__ float4 ld_w4(const WT *p) {
    if constexpr (sizeof(WT) == 4) {
        return __ldg((const float4 *)p);
    } else {
        uint2 u = __ldg((const uint2 *)p);
        __nv_bfloat162 a = *(__nv_bfloat162 *)&u.x, b = *(__nv_bfloat162 *)&u.y;
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
}

// out[m][n] = epilogue( sum_k x[m][k] * W[n][k] ),  m < Bd <= 8.  One warp owns NC output columns and streams
// their weight rows once; activations are staged in shared memory in K chunks.
template <typename WT, int NC>
__global__ void __launch_bounds__(256)
skinny_gemm_kernel(const float *__restrict__ x, int ldx, const WT *__restrict__ W, int N, int K, int Bd, SkinnyEpi e) {
    __shared__ __align__(16) float xs[SK_MB][SK_KC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_base = (blockIdx.x * 8 + warp) * NC;
    float acc[NC][SK_MB];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int m = 0; m < SK_MB; ++m) acc[c][m] = 0.f;
    for (int k0 = 0; k0 < K; k0 += SK_KC) {
        const int kc = min(SK_KC, K - k0);
        __syncthreads();
        for (int i = tid * 4; i < Bd * kc; i += 256 * 4) {
            int m = i / kc, k = i - m * kc;
            *(float4 *)&xs[m][k] = *(const float4 *)(x + (size_t)m * ldx + k0 + k);
        }
        __syncthreads();
        if (n_base < N) {
#pragma unroll 4
            for (int kk = lane * 4; kk < kc; kk += 128) {
                float4 w[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    w[c] = (n_base + c < N) ? ld_w4(W + (size_t)(n_base + c) * K + k0 + kk) : make_float4(0, 0, 0, 0);
#pragma unroll
                for (int m = 0; m < SK_MB; ++m) {
                    if (m < Bd) {
                        float4 xv = *(const float4 *)&xs[m][kk];
#pragma unroll
                        for (int c = 0; c < NC; ++c)
                            acc[c][m] += xv.x * w[c].x + xv.y * w[c].y + xv.z * w[c].z + xv.w * w[c].w;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int m = 0; m < SK_MB; ++m) {
            float v = acc[c][m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[c][m] = v;
        }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int n = n_base + c;
            if (n >= N) continue;
#pragma unroll
            for (int m = 0; m < SK_MB; ++m) {
                if (m >= Bd) continue;
                float v = acc[c][m];
                if (e.bias) v += e.bias[n];
                if (n < e.n_scale) v *= e.scale;
                if (e.act) v = gelu_tanh_precise(v);
                if (e.residual) v += e.residual[(size_t)m * e.ldr + n];
                e.out[(size_t)m * e.ldo + n] = v;
            }
        }
    }
}

template <typename ET>
__global__ void embed_kernel(const uint32_t *__restrict__ tokens, const int *__restrict__ len, int max_pos, int pos, const ET *__restrict__ embed,
                             const float *__restrict__ embed_pos, int V, int d, float *__restrict__ out) {
    const int b = blockIdx.x;
    uint32_t tok = pos < len[b] ? tokens[(size_t)b * max_pos + pos] : 0u;
    if (tok >= (uint32_t)V) tok = 0;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float ev;
        if constexpr (sizeof(ET) == 4) ev = embed[(size_t)tok * d + c];
        else ev = __bfloat162float(embed[(size_t)tok * d + c]);
        out[(size_t)b * d + c] = ev + embed_pos[(size_t)pos * d + c];
    }
}

template <typename KT>
__device__ __forceinline__ float dot64(const float *q, const KT *k) {
    float s = 0.f;
    if constexpr (sizeof(KT) == 4) {
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            float4 kv = *(const float4 *)(k + i);
            s += q[i] * kv.x + q[i + 1] * kv.y + q[i + 2] * kv.z + q[i + 3] * kv.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
            uint4 u = *(const uint4 *)(k + i);
            float2 a = __bfloat1622float2(*(__nv_bfloat162 *)&u.x), b = __bfloat1622float2(*(__nv_bfloat162 *)&u.y);
            float2 c = __bfloat1622float2(*(__nv_bfloat162 *)&u.z), dd = __bfloat1622float2(*(__nv_bfloat162 *)&u.w);
            s += q[i] * a.x + q[i + 1] * a.y + q[i + 2] * b.x + q[i + 3] * b.y + q[i + 4] * c.x + q[i + 5] * c.y + q[i + 6] * dd.x +
                 q[i + 7] * dd.y;
        }
    }
    return s;
}

// single-query attention for one (head, window): q [B][ldq] f32 (pre-scaled), cache [B][Tmax][2d] (k | v).
// If newkv != nullptr the block first appends this position's k, v (f32, already scaled) at row n_keys-1.
template <typename KT>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const float *__restrict__ q, int ldq, KT *__restrict__ cache, int Tmax, int d, int n_keys,
                   const float *__restrict__ newkv, int ldkv, int koff, int voff, float *__restrict__ out, int ldo) {
    extern __shared__ float dsm[];
    float *qs = dsm, *sc = dsm + 64, *red = sc + ((n_keys + 3) & ~3);  // red: 128 floats
    const int tid = threadIdx.x, h = blockIdx.x, b = blockIdx.y;
    KT *cb = cache + (size_t)b * Tmax * 2 * d;
    if (newkv) {
        const float *src = newkv + (size_t)b * ldkv + (tid < 64 ? koff : voff) + h * HEAD_DIM + (tid & 63);
        KT *dst = cb + (size_t)(n_keys - 1) * 2 * d + (tid < 64 ? 0 : d) + h * HEAD_DIM + (tid & 63);
        if constexpr (sizeof(KT) == 4) *dst = *src;
        else *dst = __float2bfloat16(*src);
    }
    if (tid < 64) qs[tid] = q[(size_t)b * ldq + h * HEAD_DIM + tid];
    __syncthreads();
    float lmax = -INFINITY;
    for (int j = tid; j < n_keys; j += 128) {
        float s = dot64<KT>(qs, cb + (size_t)j * 2 * d + h * HEAD_DIM);
        sc[j] = s;
        lmax = fmaxf(lmax, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    __syncthreads();
    const float mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float lsum = 0.f;
    for (int j = tid; j < n_keys; j += 128) {
        float e = expf(sc[j] - mx);
        sc[j] = e;
        lsum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((tid & 31) == 0) red[tid >> 5] = lsum;
    __syncthreads();
    const float total = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    const int half = tid >> 6, dim = tid & 63;
    float acc = 0.f;
    const KT *vb = cb + d + h * HEAD_DIM + dim;
    for (int j = half; j < n_keys; j += 2) {
        float vv;
        if constexpr (sizeof(KT) == 4) vv = vb[(size_t)j * 2 * d];
        else vv = __bfloat162float(vb[(size_t)j * 2 * d]);
        acc += sc[j] * vv;
    }
    red[tid] = acc;
    __syncthreads();
    if (tid < 64) out[(size_t)b * ldo + h * HEAD_DIM + tid] = (red[tid] + red[64 + tid]) / total;
}

// ---- block reductions for the 1024-thread select kernel ----------------------------------------------------
__device__ __forceinline__ float block_max(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
    return r;
}
__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
    return r;
}

struct SelectParams {
    float *logits;  // [B][V]; overwritten with exp(x - max)
    const float *suppress;
    uint32_t *tokens;
    int *len, *last_ts, *done, *nsampled;
    double *sumlp;
    int V, max_pos, max_new;
    float temperature;          // 0: greedy arg-max; > 0: sample from softmax(p_masked / t)
    unsigned long long seed;
    uint32_t eot, nts, ts_zero, ts_one;
};

// p = softmax(logits); apply norma's rules ON PROBABILITIES; arg-max with last-index tie-break; update state.
__global__ void __launch_bounds__(1024)
decode_select_kernel(SelectParams sp) {
    __shared__ float red[32];
    __shared__ int red_i[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (sp.done[b]) return;
    float *x = sp.logits + (size_t)b * sp.V;
    const int V = sp.V;
    const int nts = (int)sp.nts;
    float lmax = -INFINITY;
    for (int i = tid; i < V; i += 1024) lmax = fmaxf(lmax, x[i]);
    const float mx = block_max(lmax, red);
    float lsum = 0.f;
    for (int i = tid; i < V; i += 1024) {
        float e = expf(x[i] - mx);
        x[i] = e;
        lsum += e;
    }
    const float sum = block_sum(lsum, red);
    const int len = sp.len[b];
    const int last_ts = sp.last_ts[b];
    // modes: 0 first token (timestamps in [<|0.00|>, <|1.00|>]); 1 text only; 2 timestamps > last only;
    //        3 anything but timestamps <= last
    int mode;
    if (last_ts < 0) {
        mode = 0;
    } else {
        const uint32_t l_tok = sp.tokens[(size_t)b * sp.max_pos + len - 1];
        const bool has_sl = len >= 2;
        const uint32_t sl_tok = has_sl ? sp.tokens[(size_t)b * sp.max_pos + len - 2] : 0u;
        if ((int)l_tok > nts) {
            mode = (has_sl && sl_tok >= sp.eot) ? 1 : 2;
        } else {
            float ts = 0.f, mt = -INFINITY;
            for (int i = tid; i < V; i += 1024) {
                float p = x[i] / sum + sp.suppress[i];
                if (i > nts) ts += p;
                else if (i < nts) mt = fmaxf(mt, p);
            }
            const float sum_ts = block_sum(ts, red);
            const float max_text = block_max(mt, red);
            mode = (sum_ts >= max_text) ? 2 : 3;
        }
    }
    auto masked_p = [&](int i) -> float {
        bool masked;
        if (mode == 0) masked = i < (int)sp.ts_zero || i > (int)sp.ts_one;
        else {
            masked = sp.suppress[i] != 0.f;
            if (mode == 1) masked |= i > nts;
            else if (mode == 2) masked |= i <= nts || i <= last_ts;
            else masked |= (i > nts && i <= last_ts);
        }
        return masked ? -INFINITY : x[i] / sum;
    };
    float best = -INFINITY;
    int best_i = -1;
    if (sp.temperature > 0.f) {
        // model.rs:340-348: prs = softmax(p_masked / t); next = WeightedIndex(prs).sample(rng); all-NaN (everything masked) -> eot.
        // Inverse CDF over contiguous per-thread ranges; the uniform comes from a counter-based hash of (seed, window, step).
        __shared__ float s_part[1024];
        __shared__ int s_pick;
        const float inv_t = 1.0f / sp.temperature;
        float pm = -INFINITY;
        for (int i = tid; i < V; i += 1024) pm = fmaxf(pm, masked_p(i));
        const float pmax = block_max(pm, red);
        const int per = (V + 1023) / 1024, lo = tid * per, hi = min(V, lo + per);
        float local = 0.f;
        for (int i = lo; i < hi; ++i) local += expf((masked_p(i) - pmax) * inv_t);  // exp(-inf) = 0 for masked entries
        s_part[tid] = local;
        if (tid == 0) s_pick = -1;
        __syncthreads();
        if (tid == 0 && pmax > -INFINITY) {
            float total = 0.f;
            for (int t = 0; t < 1024; ++t) total += s_part[t];
            unsigned long long z = sp.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(b * 65536 + sp.nsampled[b] + 1);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;  // splitmix64
            const float target = (float)(z >> 40) * (1.0f / 16777216.0f) * total;
            float acc = 0.f;
            int t = 0;
            for (; t < 1023; ++t) {
                if (acc + s_part[t] > target) break;
                acc += s_part[t];
            }
            int pick = -1;
            const int l2 = t * per, h2 = min(V, l2 + per);
            for (int i = l2; i < h2; ++i) {
                const float q = expf((masked_p(i) - pmax) * inv_t);
                if (q > 0.f) pick = i;  // last candidate with mass: fallback against rounding at the range end
                acc += q;
                if (acc > target && q > 0.f) break;
            }
            s_pick = pick;
        }
        __syncthreads();
        best_i = s_pick;
        if (best_i >= 0) best = masked_p(best_i);
        if (tid == 0 && best_i < 0) {  // every candidate masked: the reference pushes eot and stops (model.rs:343-346)
            int l = len;
            sp.tokens[(size_t)b * sp.max_pos + l] = sp.eot;
            sp.len[b] = l + 1;
            sp.done[b] = 1;
        }
        if (best_i < 0) return;
        // fall through to the shared state update below with (best, best_i); skip the arg-max reduction
        if (tid == 0) { red[0] = best; red_i[0] = best_i; }
        for (int i = 1; i < 32; ++i)
            if (tid == 0) { red[i] = -INFINITY; red_i[i] = -1; }
        __syncthreads();
        goto update_state;
    }
    // arg-max of the masked probabilities; among equal maxima the LAST index wins (Rust `max_by`)
    for (int i = tid; i < V; i += 1024) {
        float p = masked_p(i);
        if (p >= best) { best = p; best_i = i; }  // ascending i: >= keeps the last
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi > best_i)) { best = ob; best_i = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = best; red_i[tid >> 5] = best_i; }
    __syncthreads();
update_state:
    if (tid == 0) {
        best = red[0];
        best_i = red_i[0];
        for (int i = 1; i < 32; ++i)
            if (red[i] > best || (red[i] == best && red_i[i] > best_i)) { best = red[i]; best_i = red_i[i]; }
        const uint32_t next = (uint32_t)best_i;
        int l = len;
        if ((int)next > nts) sp.last_ts[b] = (int)next;
        sp.tokens[(size_t)b * sp.max_pos + l] = next;
        ++l;
        sp.sumlp[b] += log((double)best);
        const int ns = ++sp.nsampled[b];
        if (l >= sp.max_pos - 1 || (sp.max_new > 0 && ns >= sp.max_new)) {
            sp.tokens[(size_t)b * sp.max_pos + l] = sp.eot;
            ++l;
            sp.done[b] = 1;
        } else if (next == sp.eot) {
            sp.done[b] = 1;
        }
        sp.len[b] = l;
    }
}

// no_speech_prob = softmax(logits at prompt position 0)[no_speech]; > 0.6 ends the window (model.rs:293-315)
__global__ void __launch_bounds__(1024)
nospeech_kernel(const float *__restrict__ logits, int V, uint32_t no_speech, float *__restrict__ out, int *__restrict__ done, float threshold) {
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *x = logits + (size_t)b * V;
    float lmax = -INFINITY;
    for (int i = tid; i < V; i += 1024) lmax = fmaxf(lmax, x[i]);
    const float mx = block_max(lmax, red);
    float lsum = 0.f;
    for (int i = tid; i < V; i += 1024) lsum += expf(x[i] - mx);
    const float sum = block_sum(lsum, red);
    if (tid == 0) {
        float p = expf(x[no_speech] - mx) / sum;
        out[b] = p;
        if ((double)p > (double)threshold) done[b] = 2;  // 2 = ended by the no-speech gate
    }
}

__global__ void init_state_kernel(uint32_t *tokens, int max_pos, int *len, int *last_ts, int *done, int *nsampled, double *sumlp, float *nospeech,
                                  uint32_t t0, uint32_t t1, uint32_t t2, int plen) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        tokens[(size_t)b * max_pos + 0] = t0;
        tokens[(size_t)b * max_pos + 1] = t1;
        if (plen > 2) tokens[(size_t)b * max_pos + 2] = t2;
        len[b] = plen;
        last_ts[b] = -1;
        done[b] = 0;
        nsampled[b] = 0;
        sumlp[b] = 0.0;
        nospeech[b] = 0.f;
    }
}

__global__ void copy_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

}  // namespace

static int skinny(nb200_ctx *ctx, const float *x, int ldx, const void *W, int N, int K, int Bd, const SkinnyEpi &e) {
    if (K % 4 != 0 || ldx % 4 != 0) return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "skinny gemm: K=%d ldx=%d", K, ldx);
    KernelScope ks(ctx, NB200_K_DECODE_GEMV);
    const bool bf = ctx->compute == NB200_BF16;
    for (int m0 = 0; m0 < Bd; m0 += SK_MB) {
        const int mb = Bd - m0 < SK_MB ? Bd - m0 : SK_MB;
        SkinnyEpi ee = e;
        ee.out = e.out + (size_t)m0 * e.ldo;
        if (e.residual) ee.residual = e.residual + (size_t)m0 * e.ldr;
        const float *xx = x + (size_t)m0 * ldx;
        if (N > 2048) {
            const int blocks = ceil_div(N, 8 * 2);
            if (bf) skinny_gemm_kernel<bf16, 2><<<blocks, 256, 0, ctx->stream>>>(xx, ldx, (const bf16 *)W, N, K, mb, ee);
            else skinny_gemm_kernel<float, 2><<<blocks, 256, 0, ctx->stream>>>(xx, ldx, (const float *)W, N, K, mb, ee);
        } else {
            const int blocks = ceil_div(N, 8);
            if (bf) skinny_gemm_kernel<bf16, 1><<<blocks, 256, 0, ctx->stream>>>(xx, ldx, (const bf16 *)W, N, K, mb, ee);
            else skinny_gemm_kernel<float, 1><<<blocks, 256, 0, ctx->stream>>>(xx, ldx, (const float *)W, N, K, mb, ee);
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

static int dec_attn(nb200_ctx *ctx, const float *q, int ldq, void *cache, int Tmax, int n_keys, const float *newkv, int ldkv, int koff, int voff,
                    float *out, int ldo, int n_windows) {
    KernelScope ks(ctx, NB200_K_DECODE_ATTN);
    const int d = ctx->cfg.d_model, heads = ctx->cfg.decoder_attention_heads;
    const size_t smem = (64 + ((n_keys + 3) & ~3) + 128) * sizeof(float);
    dim3 grid(heads, n_windows);
    if (ctx->compute == NB200_BF16)
        decode_attn_kernel<bf16><<<grid, 128, smem, ctx->stream>>>(q, ldq, (bf16 *)cache, Tmax, d, n_keys, newkv, ldkv, koff, voff, out, ldo);
    else
        decode_attn_kernel<float><<<grid, 128, smem, ctx->stream>>>(q, ldq, (float *)cache, Tmax, d, n_keys, newkv, ldkv, koff, voff, out, ldo);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int decoder_build_cross_kv(nb200_ctx *ctx, int n_windows) {
    const int d = ctx->cfg.d_model, T = ctx->cfg.max_source_positions, L = ctx->cfg.decoder_layers;
    const size_t es = dtype_size(ctx->compute);
    const float kscale = powf((float)HEAD_DIM, -0.25f);
    for (int l = 0; l < L; ++l) {
        GemmShape s{T * n_windows, 1, 2 * d, d, d, (long long)T * n_windows * d};
        Epilogue e{};
        e.bias = ctx->dec[l].cbkv;
        e.out = (char *)ctx->cross_kv + (size_t)l * ctx->cfg.max_batch * T * 2 * d * es;
        e.ldo = 2 * d;
        e.out_bs = 0;
        e.scale = kscale;
        e.n_scale = d;  // k columns
        e.out_bf16 = ctx->compute == NB200_BF16;
        if (ctx->compute == NB200_BF16) NB_TRY(launch_gemm_bf16(ctx, (const bf16 *)ctx->enc_out_c, (const bf16 *)ctx->dec[l].cwkv, s, e));
        else NB_TRY(launch_gemm_f32(ctx, (const float *)ctx->enc_out_c, (const float *)ctx->dec[l].cwkv, s, e));
    }
    ctx->cross_valid = true;
    return NB200_OK;
}

int decoder_init_state(nb200_ctx *ctx, int n_windows) {
    KernelScope ks(ctx, NB200_K_MISC);
    const bool has_lang = ctx->tok.lang != UINT32_MAX;
    init_state_kernel<<<n_windows, 32, 0, ctx->stream>>>(ctx->d_tokens, ctx->cfg.max_target_positions, ctx->d_len, ctx->d_last_ts, ctx->d_done,
                                                         ctx->d_nsampled, ctx->d_sumlp, ctx->d_nospeech, ctx->tok.sot,
                                                         has_lang ? ctx->tok.lang : ctx->tok.task, ctx->tok.task, has_lang ? 3 : 2);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

// one decoder position `pos` for windows [0, n_windows): reads token ids from d_tokens, leaves the final LayerNorm
// output in dhid and (optionally) the logits in ctx->logits
int decoder_step(nb200_ctx *ctx, int w0, int n_windows, int pos, int want_logits) {
    const nb200_config &c = ctx->cfg;
    const int d = c.d_model, T = c.max_source_positions, P = c.max_target_positions, V = c.vocab_size;
    const size_t es = dtype_size(ctx->compute);
    const float qscale = powf((float)HEAD_DIM, -0.25f);
    const int B = n_windows;
    {
        KernelScope ks(ctx, NB200_K_MISC);
        if (ctx->compute == NB200_BF16)
            embed_kernel<bf16><<<B, 256, 0, ctx->stream>>>(ctx->d_tokens + (size_t)w0 * P, ctx->d_len + w0, P, pos, (const bf16 *)ctx->embed, ctx->embed_pos, V, d, ctx->dx);
        else
            embed_kernel<float><<<B, 256, 0, ctx->stream>>>(ctx->d_tokens + (size_t)w0 * P, ctx->d_len + w0, P, pos, (const float *)ctx->embed, ctx->embed_pos, V, d, ctx->dx);
    }
    for (int l = 0; l < c.decoder_layers; ++l) {
        const DecLayer &w = ctx->dec[l];
        char *skv = (char *)ctx->self_kv + ((size_t)l * c.max_batch + w0) * P * 2 * d * es;
        char *ckv = (char *)ctx->cross_kv + ((size_t)l * c.max_batch + w0) * T * 2 * d * es;
        // self attention (q, k scaled by hd^-0.25 as candle does at attention time; k has no bias)
        NB_TRY(launch_layernorm(ctx, ctx->dx, w.ln1g, w.ln1b, B, d, ctx->dh, 0, nullptr));
        SkinnyEpi e{};
        e.bias = w.bqkv; e.out = ctx->dqkv; e.ldo = 3 * d; e.scale = qscale; e.n_scale = 2 * d;
        NB_TRY(skinny(ctx, ctx->dh, d, w.wqkv, 3 * d, d, B, e));
        NB_TRY(dec_attn(ctx, ctx->dqkv, 3 * d, skv, P, pos + 1, ctx->dqkv, 3 * d, d, 2 * d, ctx->dattn, d, B));
        e = SkinnyEpi{};
        e.bias = w.bo; e.out = ctx->dx; e.ldo = d; e.residual = ctx->dx; e.ldr = d;
        NB_TRY(skinny(ctx, ctx->dattn, d, w.wo, d, d, B, e));
        // cross attention over the cached K/V of the audio features
        NB_TRY(launch_layernorm(ctx, ctx->dx, w.lncg, w.lncb, B, d, ctx->dh, 0, nullptr));
        e = SkinnyEpi{};
        e.bias = w.cbq; e.out = ctx->dq; e.ldo = d; e.scale = qscale; e.n_scale = d;
        NB_TRY(skinny(ctx, ctx->dh, d, w.cwq, d, d, B, e));
        NB_TRY(dec_attn(ctx, ctx->dq, d, ckv, T, T, nullptr, 0, 0, 0, ctx->dattn, d, B));
        e = SkinnyEpi{};
        e.bias = w.cbo; e.out = ctx->dx; e.ldo = d; e.residual = ctx->dx; e.ldr = d;
        NB_TRY(skinny(ctx, ctx->dattn, d, w.cwo, d, d, B, e));
        // MLP
        NB_TRY(launch_layernorm(ctx, ctx->dx, w.ln2g, w.ln2b, B, d, ctx->dh, 0, nullptr));
        e = SkinnyEpi{};
        e.bias = w.b1; e.out = ctx->dff; e.ldo = 4 * d; e.act = 1;
        NB_TRY(skinny(ctx, ctx->dh, d, w.w1, 4 * d, d, B, e));
        e = SkinnyEpi{};
        e.bias = w.b2; e.out = ctx->dx; e.ldo = d; e.residual = ctx->dx; e.ldr = d;
        NB_TRY(skinny(ctx, ctx->dff, 4 * d, w.w2, d, 4 * d, B, e));
    }
    NB_TRY(launch_layernorm(ctx, ctx->dx, ctx->lndec_g, ctx->lndec_b, B, d, ctx->dhid, 0, nullptr));
    if (want_logits) {
        SkinnyEpi e{};
        e.out = ctx->logits; e.ldo = V;
        NB_TRY(skinny(ctx, ctx->dhid, d, ctx->embed, V, d, B, e));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int decoder_nospeech(nb200_ctx *ctx, int n_windows) {
    KernelScope ks(ctx, NB200_K_DECODE_SELECT);
    nospeech_kernel<<<n_windows, 1024, 0, ctx->stream>>>(ctx->logits, ctx->cfg.vocab_size, ctx->tok.no_speech, ctx->d_nospeech, ctx->d_done, 0.6f);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int decoder_select(nb200_ctx *ctx, int n_windows, int max_new_tokens, float temperature, unsigned long long seed) {
    KernelScope ks(ctx, NB200_K_DECODE_SELECT);
    SelectParams sp;
    sp.logits = ctx->logits; sp.suppress = ctx->suppress; sp.tokens = ctx->d_tokens;
    sp.len = ctx->d_len; sp.last_ts = ctx->d_last_ts; sp.done = ctx->d_done; sp.nsampled = ctx->d_nsampled; sp.sumlp = ctx->d_sumlp;
    sp.V = ctx->cfg.vocab_size; sp.max_pos = ctx->cfg.max_target_positions; sp.max_new = max_new_tokens;
    sp.temperature = temperature; sp.seed = seed;
    sp.eot = ctx->tok.eot; sp.nts = ctx->tok.no_timestamps; sp.ts_zero = ctx->tok.ts_zero; sp.ts_one = ctx->tok.ts_one;
    decode_select_kernel<<<n_windows, 1024, 0, ctx->stream>>>(sp);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int decoder_copy_hidden(nb200_ctx *ctx, float *dst, int d) {
    KernelScope ks(ctx, NB200_K_MISC);
    copy_rows_kernel<<<ceil_div(d, 256), 256, 0, ctx->stream>>>(ctx->dhid, dst, d);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int decoder_init(nb200_ctx *ctx) {
    // decode attention: scores for up to 1500 keys in dynamic smem (< 48 KB: no opt-in needed)
    (void)ctx;
    return NB200_OK;
}

// logits rows [0, n) = dhid rows [0, n) . embed^T (tied embedding, no bias): `TextDecoder::final_linear`
int decoder_logits_rows(nb200_ctx *ctx, int n) {
    SkinnyEpi e{};
    e.out = ctx->logits;
    e.ldo = ctx->cfg.vocab_size;
    return skinny(ctx, ctx->dhid, ctx->cfg.d_model, ctx->embed, ctx->cfg.vocab_size, ctx->cfg.d_model, n, e);
}
