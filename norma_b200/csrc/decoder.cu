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

// Per-decode dynamic state kept on the device so that ONE captured CUDA graph serves every step of every decode call:
// the position advances on the device (decode_select_kernel), temperature / seed / token budget are read from here.
struct DecodeDyn {
    int pos;
    int max_new;
    float temperature;
    unsigned long long seed;
};

struct SkinnyEpi {
    const float *bias;
    const float *residual;  // [Bd][ldr] or nullptr (may alias out)
    float *out;             // [Bd][ldo]
    int ldr, ldo;
    float scale;
    int n_scale;
    int act;
};

constexpr int SK_MB = 8;      // activation rows per pass
constexpr int SK_KC_MAX = 1280;  // K chunk staged in shared memory (whole row when K = d_model: LayerNorm can be fused)

template <typename WT>
__device__ __forceinline__ void ld_w8(const WT *p, float *w) {
    if constexpr (sizeof(WT) == 4) {
        const float4 a = __ldg((const float4 *)p), b = __ldg((const float4 *)p + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
        const uint4 u = __ldg((const uint4 *)p);  // 8 bf16 = 16 B per lane, 512 B per warp instruction
        const float2 f0 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.x), f1 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.y);
        const float2 f2 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.z), f3 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.w);
        w[0] = f0.x; w[1] = f0.y; w[2] = f1.x; w[3] = f1.y; w[4] = f2.x; w[5] = f2.y; w[6] = f3.x; w[7] = f3.y;
    }
}

// out[m][n] = epilogue( sum_k LN?(x)[m][k] * W[n][k] ),  m < Bd <= 8: the decoder's weight-streaming GEMV-class GEMM.
// Persistent blocks: the activation rows are staged in shared memory ONCE per block (with the LayerNorm fused when
// ln_g != nullptr, K = d_model), then the block walks column groups; one warp owns NC output columns per group and streams
// their weight rows with 16-byte loads (5 x NC independent loads in flight per lane at K = 1280).  K > kc_max (fc2) is
// handled in K chunks and then needs one group per block.  Block 0 optionally writes the normalised rows to ln_out.
template <typename WT, int NC>
__global__ void __launch_bounds__(256)
skinny_gemm_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ ln_g, const float *__restrict__ ln_b, float *__restrict__ ln_out,
                   const WT *__restrict__ W, int N, int K, int kc_max, int Bd, int n_groups, SkinnyEpi e) {
    extern __shared__ __align__(16) float xs[];  // [SK_MB][kc_max]
    __shared__ float red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_chunks = (K + kc_max - 1) / kc_max;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int n_base = (g * 8 + warp) * NC;
        float acc[NC][SK_MB];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int m = 0; m < SK_MB; ++m) acc[c][m] = 0.f;
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int k0 = ch * kc_max, kc = min(kc_max, K - k0);
            if (n_chunks > 1 || g == (int)blockIdx.x) {  // single chunk: staged once, reused by every group of this block
                __syncthreads();
                for (int i = tid * 4; i < Bd * kc; i += 256 * 4) {
                    const int m = i / kc, k = i - m * kc;
                    *(float4 *)&xs[m * kc_max + k] = *(const float4 *)(x + (size_t)m * ldx + k0 + k);
                }
                __syncthreads();
                if (ln_g) {  // candle_nn LayerNorm, eps 1e-5, over the whole row (K == kc)
                    for (int m = 0; m < Bd; ++m) {
                        float s1 = 0.f;
                        for (int k = tid; k < kc; k += 256) s1 += xs[m * kc_max + k];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        if (lane == 0) red[warp] = s1;
                        __syncthreads();
                        const float mean = (red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7]) / (float)kc;
                        __syncthreads();
                        float s2 = 0.f;
                        for (int k = tid; k < kc; k += 256) {
                            const float dv = xs[m * kc_max + k] - mean;
                            s2 += dv * dv;
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        if (lane == 0) red[warp] = s2;
                        __syncthreads();
                        const float rstd = rsqrtf((red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7]) / (float)kc + 1e-5f);
                        for (int k = tid; k < kc; k += 256) {
                            const float y = (xs[m * kc_max + k] - mean) * rstd * __ldg(ln_g + k) + __ldg(ln_b + k);
                            xs[m * kc_max + k] = y;
                            if (ln_out && blockIdx.x == 0) ln_out[(size_t)m * kc + k] = y;
                        }
                        __syncthreads();
                    }
                }
            }
            if (n_base < N) {
#pragma unroll 5
                for (int kk = lane * 8; kk < kc; kk += 256) {
                    float w[NC][8];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (n_base + c < N) ld_w8(W + (size_t)(n_base + c) * K + k0 + kk, w[c]);
                        else {
#pragma unroll
                            for (int t = 0; t < 8; ++t) w[c][t] = 0.f;
                        }
                    }
#pragma unroll
                    for (int m = 0; m < SK_MB; ++m) {
                        if (m < Bd) {
                            const float4 x0 = *(const float4 *)&xs[m * kc_max + kk], x1 = *(const float4 *)&xs[m * kc_max + kk + 4];
#pragma unroll
                            for (int c = 0; c < NC; ++c)
                                acc[c][m] += x0.x * w[c][0] + x0.y * w[c][1] + x0.z * w[c][2] + x0.w * w[c][3] + x1.x * w[c][4] + x1.y * w[c][5] +
                                             x1.z * w[c][6] + x1.w * w[c][7];
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
}

template <typename ET>
__global__ void embed_kernel(const uint32_t *__restrict__ tokens, const int *__restrict__ len, int max_pos, const DecodeDyn *__restrict__ dyn,
                             const ET *__restrict__ embed,
                             const float *__restrict__ embed_pos, int V, int d, float *__restrict__ out) {
    const int b = blockIdx.x;
    const int pos = dyn->pos;
    uint32_t tok = pos < len[b] ? tokens[(size_t)b * max_pos + pos] : 0u;
    if (tok >= (uint32_t)V) tok = 0;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float ev;
        if constexpr (sizeof(ET) == 4) ev = embed[(size_t)tok * d + c];
        else ev = __bfloat162float(embed[(size_t)tok * d + c]);
        out[(size_t)b * d + c] = ev + embed_pos[(size_t)pos * d + c];
    }
}

// Split-K single-query attention (flash-decoding): grid (heads, windows, splits).  8 lanes share one key: every lane
// loads 16 B / 32 B of the K (then V) row, so a key row is one coalesced 128 B / 256 B access; each 8-lane group runs its own
// online softmax over every 16th key of the split, groups are merged with shuffles and shared memory, and the split's
// (m, l, o[64]) goes to a workspace that decode_attn_merge_kernel folds.  For self attention the split that owns the
// newest position takes k, v from the QKV GEMV output (f32) and appends them to the cache.
constexpr int ATT_WS = HEAD_DIM + 2;

template <typename KT>
__device__ __forceinline__ void ld_kv8(const KT *p, float *v) {
    if constexpr (sizeof(KT) == 4) {
        const float4 a = *(const float4 *)p, b = *((const float4 *)p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = *(const uint4 *)p;
        const float2 f0 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.x), f1 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.y);
        const float2 f2 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.z), f3 = __bfloat1622float2(*(const __nv_bfloat162 *)&u.w);
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
    }
}

__device__ __forceinline__ void osm_merge(float &m, float &l, float *acc, float m2, float l2, const float *acc2) {
    const float mn = fmaxf(m, m2);
    const float a = (m == -INFINITY) ? 0.f : expf(m - mn), b = (m2 == -INFINITY) ? 0.f : expf(m2 - mn);
    l = l * a + l2 * b;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = acc[t] * a + acc2[t] * b;
    m = mn;
}

template <typename KT>
__global__ void __launch_bounds__(128)
decode_attn_part_kernel(const float *__restrict__ q, int ldq, KT *__restrict__ cache, int Tmax, int d, int n_keys_static,
                        const DecodeDyn *__restrict__ dyn, const float *__restrict__ newkv, int ldkv, int koff, int voff,
                        float *__restrict__ ws, int S) {
    __shared__ float sm[4][8][10];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l8 = lane & 7, grp = lane >> 3;
    const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, H = gridDim.x;
    const int n_keys = dyn ? dyn->pos + 1 : n_keys_static;
    const int chunk = (n_keys + S - 1) / S, k0 = sp * chunk, k1 = min(n_keys, k0 + chunk);
    KT *cb = cache + (size_t)b * Tmax * 2 * d + h * HEAD_DIM + l8 * 8;
    float qr[8];
    {
        const float *qp = q + (size_t)b * ldq + h * HEAD_DIM + l8 * 8;
        const float4 a = *(const float4 *)qp, c = *((const float4 *)qp + 1);
        qr[0] = a.x; qr[1] = a.y; qr[2] = a.z; qr[3] = a.w; qr[4] = c.x; qr[5] = c.y; qr[6] = c.z; qr[7] = c.w;
    }
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    const unsigned gmask = 0xffu << (grp * 8);
    for (int j = k0 + warp * 4 + grp; j < k1; j += 16) {
        float kv[8], vv[8];
        if (newkv && j == n_keys - 1) {  // the position being decoded: k, v come from the QKV GEMV and join the cache
            const float *kp = newkv + (size_t)b * ldkv + koff + h * HEAD_DIM + l8 * 8, *vp = newkv + (size_t)b * ldkv + voff + h * HEAD_DIM + l8 * 8;
            KT *kd = cb + (size_t)j * 2 * d, *vd = kd + d;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                kv[t] = kp[t];
                vv[t] = vp[t];
                if constexpr (sizeof(KT) == 4) { kd[t] = kv[t]; vd[t] = vv[t]; }
                else {
                    kd[t] = __float2bfloat16(kv[t]); vd[t] = __float2bfloat16(vv[t]);
                    kv[t] = __bfloat162float(kd[t]); vv[t] = __bfloat162float(vd[t]);  // what every later step will read
                }
            }
        } else {
            ld_kv8<KT>(cb + (size_t)j * 2 * d, kv);
            ld_kv8<KT>(cb + (size_t)j * 2 * d + d, vv);
        }
        float sdot = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) sdot = fmaf(qr[t], kv[t], sdot);
        sdot += __shfl_xor_sync(gmask, sdot, 1);  // group-local mask: the four groups of a warp run different trip counts
        sdot += __shfl_xor_sync(gmask, sdot, 2);
        sdot += __shfl_xor_sync(gmask, sdot, 4);
        const float mn = fmaxf(m, sdot);
        const float a = expf(m - mn), p = expf(sdot - mn);  // m = -inf on the first key: a = 0
        l = l * a + p;
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = acc[t] * a + p * vv[t];
        m = mn;
    }
    // merge the four 8-lane groups of the warp, then the four warps
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o), a2[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) a2[t] = __shfl_xor_sync(0xffffffffu, acc[t], o);
        osm_merge(m, l, acc, m2, l2, a2);
    }
    if (grp == 0) {
        sm[warp][l8][0] = m;
        sm[warp][l8][1] = l;
#pragma unroll
        for (int t = 0; t < 8; ++t) sm[warp][l8][2 + t] = acc[t];
    }
    __syncthreads();
    if (warp == 0 && grp == 0) {
        for (int w2 = 1; w2 < 4; ++w2) osm_merge(m, l, acc, sm[w2][l8][0], sm[w2][l8][1], &sm[w2][l8][2]);
        float *o = ws + (((size_t)b * H + h) * S + sp) * ATT_WS;
        if (l8 == 0) { o[0] = m; o[1] = l; }
#pragma unroll
        for (int t = 0; t < 8; ++t) o[2 + l8 * 8 + t] = acc[t];
    }
}

__global__ void __launch_bounds__(64)
decode_attn_merge_kernel(const float *__restrict__ ws, int S, float *__restrict__ out, int ldo) {
    const int h = blockIdx.x, b = blockIdx.y, H = gridDim.x, t = threadIdx.x;
    const float *p = ws + ((size_t)b * H + h) * S * ATT_WS;
    float M = -INFINITY;
    for (int s = 0; s < S; ++s) M = fmaxf(M, p[s * ATT_WS]);
    float l = 0.f, o = 0.f;
    for (int s = 0; s < S; ++s) {
        const float ms = p[s * ATT_WS];
        if (ms == -INFINITY) continue;  // empty split
        const float a = expf(ms - M);
        l += p[s * ATT_WS + 1] * a;
        o += p[s * ATT_WS + 2 + t] * a;
    }
    out[(size_t)b * ldo + h * HEAD_DIM + t] = o / l;
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
    int V, max_pos;
    DecodeDyn *dyn;  // temperature 0: greedy arg-max; > 0: sample from softmax(p_masked / t)
    uint32_t eot, nts, ts_zero, ts_one;
};

// p = softmax(logits); apply norma's rules ON PROBABILITIES; arg-max with last-index tie-break; update state.
__global__ void __launch_bounds__(1024)
decode_select_kernel(SelectParams sp) {
    __shared__ float red[32];
    __shared__ int red_i[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float temperature = sp.dyn->temperature;
    const int max_new = sp.dyn->max_new;
    const unsigned long long seed = sp.dyn->seed;
    __syncthreads();  // everyone has read dyn before block 0 advances the position below
    if (b == 0 && tid == 0) sp.dyn->pos += 1;  // the step is over once this kernel runs: the next graph replay sees pos + 1
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
    if (temperature > 0.f) {
        // model.rs:340-348: prs = softmax(p_masked / t); next = WeightedIndex(prs).sample(rng); all-NaN (everything masked) -> eot.
        // Inverse CDF over contiguous per-thread ranges; the uniform comes from a counter-based hash of (seed, window, step).
        __shared__ float s_part[1024];
        __shared__ int s_pick;
        const float inv_t = 1.0f / temperature;
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
            unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(b * 65536 + sp.nsampled[b] + 1);
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
        if (l >= sp.max_pos - 1 || (max_new > 0 && ns >= max_new)) {
            sp.tokens[(size_t)b * sp.max_pos + l] = sp.eot;
            ++l;
            sp.done[b] = 1;
        } else if (next == sp.eot) {
            sp.done[b] = 1;
        }
        sp.len[b] = l;
    }
}

// ---- greedy (t = 0) select in three short multi-block phases (the single 1024-thread block above costs ~58 us per step
// at V = 51 866; it remains the path for t > 0) --------------------------------------------------------------------
constexpr int SEL_CH = 32;  // vocabulary chunks per window
struct SelCand { float sum_ts, max_text, best_a, best_b; int idx_a, idx_b; };

// phase A: per-chunk (max, sum exp(x - max))
__global__ void __launch_bounds__(256)
select_stats_kernel(const float *__restrict__ logits, int V, float2 *__restrict__ ws_a) {
    __shared__ float red[32];
    const int b = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int per = (V + SEL_CH - 1) / SEL_CH, lo = c * per, hi = min(V, lo + per);
    const float *x = logits + (size_t)b * V;
    float lm = -INFINITY;
    for (int i = lo + tid; i < hi; i += 256) lm = fmaxf(lm, x[i]);
    const float mx = block_max(lm, red);
    float ls = 0.f;
    for (int i = lo + tid; i < hi; i += 256) ls += expf(x[i] - mx);
    const float sm = block_sum(ls, red);
    if (tid == 0) ws_a[b * SEL_CH + c] = make_float2(mx, sm);
}

// phase B: with the global (max, sum) every chunk yields its share of sum_ts / max_text and the arg-max candidates of the
// rule(s) that can still apply: mode 0 / 1 are known from the token state; otherwise both 2 and 3 are prepared
__global__ void __launch_bounds__(256)
select_cand_kernel(SelectParams sp, const float2 *__restrict__ ws_a, SelCand *__restrict__ ws_b) {
    __shared__ float red[32];
    __shared__ int red_i[32];
    const int b = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    if (sp.done[b]) return;
    const int V = sp.V, nts = (int)sp.nts;
    float M = -INFINITY;
    for (int k = 0; k < SEL_CH; ++k) M = fmaxf(M, ws_a[b * SEL_CH + k].x);
    float S = 0.f;
    for (int k = 0; k < SEL_CH; ++k) S += ws_a[b * SEL_CH + k].y * expf(ws_a[b * SEL_CH + k].x - M);
    const int len = sp.len[b], last_ts = sp.last_ts[b];
    int mode_a, mode_b = -1;
    if (last_ts < 0) mode_a = 0;
    else {
        const uint32_t l_tok = sp.tokens[(size_t)b * sp.max_pos + len - 1];
        const bool has_sl = len >= 2;
        const uint32_t sl_tok = has_sl ? sp.tokens[(size_t)b * sp.max_pos + len - 2] : 0u;
        if ((int)l_tok > nts) mode_a = (has_sl && sl_tok >= sp.eot) ? 1 : 2;
        else { mode_a = 2; mode_b = 3; }
    }
    const int per = (V + SEL_CH - 1) / SEL_CH, lo = c * per, hi = min(V, lo + per);
    const float *x = sp.logits + (size_t)b * V;
    float ts = 0.f, mt = -INFINITY, ba = -INFINITY, bb = -INFINITY;
    int ia = -1, ib = -1;
    auto masked = [&](int mode, int i, float sup) -> bool {
        if (mode == 0) return i < (int)sp.ts_zero || i > (int)sp.ts_one;
        bool mk = sup != 0.f;
        if (mode == 1) mk |= i > nts;
        else if (mode == 2) mk |= i <= nts || i <= last_ts;
        else mk |= (i > nts && i <= last_ts);
        return mk;
    };
    for (int i = lo + tid; i < hi; i += 256) {
        const float p = expf(x[i] - M) / S, sup = sp.suppress[i];
        if (mode_b >= 0) {
            const float ps = p + sup;
            if (i > nts) ts += ps;
            else if (i < nts) mt = fmaxf(mt, ps);
        }
        const float pa = masked(mode_a, i, sup) ? -INFINITY : p;
        if (pa >= ba) { ba = pa; ia = i; }
        if (mode_b >= 0) {
            const float pb = masked(mode_b, i, sup) ? -INFINITY : p;
            if (pb >= bb) { bb = pb; ib = i; }
        }
    }
    auto arg_reduce = [&](float &bv, int &bi) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();
        if ((tid & 31) == 0) { red[tid >> 5] = bv; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0)
            for (int k = 1; k < 8; ++k)
                if (red[k] > bv || (red[k] == bv && red_i[k] > bi)) { bv = red[k]; bi = red_i[k]; }
    };
    arg_reduce(ba, ia);
    if (mode_b >= 0) arg_reduce(bb, ib);
    const float sum_ts = block_sum(ts, red);
    const float max_text = block_max(mt, red);
    if (tid == 0) ws_b[b * SEL_CH + c] = SelCand{sum_ts, max_text, ba, bb, ia, ib};
}

// phase C: one warp per window folds the chunk results, applies norma's rule and updates the decoding state
__global__ void __launch_bounds__(32)
select_final_kernel(SelectParams sp, const SelCand *__restrict__ ws_b) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int max_new = sp.dyn->max_new;
    __syncwarp();
    if (b == 0 && lane == 0) sp.dyn->pos += 1;
    if (sp.done[b]) return;
    const SelCand cd = ws_b[b * SEL_CH + lane];  // SEL_CH == 32: one chunk per lane
    float ts = cd.sum_ts, mt = cd.max_text;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
    }
    const int len = sp.len[b], last_ts = sp.last_ts[b], nts = (int)sp.nts;
    bool use_b = false;
    if (last_ts >= 0) {
        const uint32_t l_tok = sp.tokens[(size_t)b * sp.max_pos + len - 1];
        if ((int)l_tok <= nts) use_b = !(ts >= mt);  // sum_prob_timestamp >= prob_non_timestamp -> timestamps only (model.rs:272)
    }
    float best = use_b ? cd.best_b : cd.best_a;
    int best_i = use_b ? cd.idx_b : cd.idx_a;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi > best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) {
        const uint32_t next = (uint32_t)best_i;
        int l = len;
        if ((int)next > nts) sp.last_ts[b] = (int)next;
        sp.tokens[(size_t)b * sp.max_pos + l] = next;
        ++l;
        sp.sumlp[b] += log((double)best);
        const int ns = ++sp.nsampled[b];
        if (l >= sp.max_pos - 1 || (max_new > 0 && ns >= max_new)) {
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

// `detect_language` (model.rs:198-207): softmax over the logits of the language tokens only, then the FIRST maximum in
// `Language` order (a stable descending sort by total_cmp keeps the earlier of equal probabilities first)
__global__ void __launch_bounds__(1024)
language_kernel(const float *__restrict__ logits, const uint32_t *__restrict__ ids, int n, float *__restrict__ probs, int *__restrict__ best) {
    __shared__ float red[32];
    __shared__ int redi[32];
    const int tid = threadIdx.x;
    const float x = tid < n ? logits[ids[tid]] : -INFINITY;
    const float mx = block_max(x, red);
    const float e = tid < n ? expf(x - mx) : 0.f;
    const float sum = block_sum(e, red);
    const float p = e / sum;
    if (tid < n) probs[tid] = p;
    // arg-max of p with the lowest index winning ties
    float bp = tid < n ? p : -1.f;
    int bi = tid < n ? tid : 0x7fffffff;
    for (int o = 16; o; o >>= 1) {
        const float op = __shfl_xor_sync(0xffffffffu, bp, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (op > bp || (op == bp && oi < bi)) { bp = op; bi = oi; }
    }
    __syncthreads();  // everyone is done reading red[] in block_sum
    if ((tid & 31) == 0) { red[tid >> 5] = bp; redi[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
        bp = red[tid];
        bi = redi[tid];
        for (int o = 16; o; o >>= 1) {
            const float op = __shfl_xor_sync(0xffffffffu, bp, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (op > bp || (op == bp && oi < bi)) { bp = op; bi = oi; }
        }
        if (tid == 0) *best = bi;
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

__global__ void set_dyn_kernel(DecodeDyn *dyn, int pos, int max_new, float temperature, unsigned long long seed, int set_params) {
    dyn->pos = pos;
    if (set_params) {
        dyn->max_new = max_new;
        dyn->temperature = temperature;
        dyn->seed = seed;
    }
}

__global__ void copy_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

}  // namespace

static int skinny_ln(nb200_ctx *ctx, const float *x, int ldx, const float *ln_g, const float *ln_b, float *ln_out, const void *W, int N, int K,
                     int Bd, const SkinnyEpi &e) {
    if (K % 8 != 0 || ldx % 4 != 0) return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "skinny gemm: K=%d ldx=%d", K, ldx);
    const int kc = K < SK_KC_MAX ? K : SK_KC_MAX;
    if (ln_g && K != kc) return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "skinny gemm: fused LayerNorm needs K <= %d", SK_KC_MAX);
    KernelScope ks(ctx, NB200_K_DECODE_GEMV);
    const bool bf = ctx->compute == NB200_BF16;
    const size_t smem = (size_t)SK_MB * kc * sizeof(float);
    for (int m0 = 0; m0 < Bd; m0 += SK_MB) {
        const int mb = Bd - m0 < SK_MB ? Bd - m0 : SK_MB;
        SkinnyEpi ee = e;
        ee.out = e.out + (size_t)m0 * e.ldo;
        if (e.residual) ee.residual = e.residual + (size_t)m0 * e.ldr;
        const float *xx = x + (size_t)m0 * ldx;
        float *lo = ln_out ? ln_out + (size_t)m0 * K : nullptr;
        const int max_blocks = 4 * ctx->sm_count;
#define SK_LAUNCH(WT_, NC_)                                                                                                                 \
    {                                                                                                                                       \
        const int groups = ceil_div(N, 8 * NC_);                                                                                            \
        const int blocks = (K > kc || groups < max_blocks) ? groups : max_blocks;                                                           \
        skinny_gemm_kernel<WT_, NC_><<<blocks, 256, smem, ctx->stream>>>(xx, ldx, ln_g, ln_b, lo, (const WT_ *)W, N, K, kc, mb, groups, ee); \
    }
        if (N > 8192) {
            if (bf) SK_LAUNCH(bf16, 4) else SK_LAUNCH(float, 4)
        } else if (N > 2048) {
            if (bf) SK_LAUNCH(bf16, 2) else SK_LAUNCH(float, 2)
        } else {
            if (bf) SK_LAUNCH(bf16, 1) else SK_LAUNCH(float, 1)
        }
#undef SK_LAUNCH
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

static int skinny(nb200_ctx *ctx, const float *x, int ldx, const void *W, int N, int K, int Bd, const SkinnyEpi &e) {
    return skinny_ln(ctx, x, ldx, nullptr, nullptr, nullptr, W, N, K, Bd, e);
}

static int dec_attn(nb200_ctx *ctx, const float *q, int ldq, void *cache, int Tmax, int n_keys, bool dyn_keys, const float *newkv, int ldkv, int koff,
                    int voff, float *out, int ldo, int n_windows) {
    KernelScope ks(ctx, NB200_K_DECODE_ATTN);
    const int d = ctx->cfg.d_model, heads = ctx->cfg.decoder_attention_heads;
    const DecodeDyn *dyn = dyn_keys ? (const DecodeDyn *)ctx->d_dyn : nullptr;
    const int S = dyn_keys ? 2 : 8;  // self attention: <= 448 keys; cross attention: 1500
    dim3 grid(heads, n_windows, S);
    if (ctx->compute == NB200_BF16)
        decode_attn_part_kernel<bf16><<<grid, 128, 0, ctx->stream>>>(q, ldq, (bf16 *)cache, Tmax, d, n_keys, dyn, newkv, ldkv, koff, voff, ctx->d_attn_ws, S);
    else
        decode_attn_part_kernel<float><<<grid, 128, 0, ctx->stream>>>(q, ldq, (float *)cache, Tmax, d, n_keys, dyn, newkv, ldkv, koff, voff, ctx->d_attn_ws, S);
    decode_attn_merge_kernel<<<dim3(heads, n_windows), 64, 0, ctx->stream>>>(ctx->d_attn_ws, S, out, ldo);
    ctx->launches++;
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
int decoder_set_dyn(nb200_ctx *ctx, int pos, int max_new, float temperature, unsigned long long seed, int set_params) {
    KernelScope ks(ctx, NB200_K_MISC);
    set_dyn_kernel<<<1, 1, 0, ctx->stream>>>((DecodeDyn *)ctx->d_dyn, pos, max_new, temperature, seed, set_params);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

// pos >= 0: the host names the position (prompt / seam calls); pos < 0: use the device-resident position (graph replay)
int decoder_step(nb200_ctx *ctx, int w0, int n_windows, int pos, int want_logits) {
    if (pos >= 0) NB_TRY(decoder_set_dyn(ctx, pos, 0, 0.f, 0, 0));
    const nb200_config &c = ctx->cfg;
    const int d = c.d_model, T = c.max_source_positions, P = c.max_target_positions, V = c.vocab_size;
    const size_t es = dtype_size(ctx->compute);
    const float qscale = powf((float)HEAD_DIM, -0.25f);
    const int B = n_windows;
    {
        KernelScope ks(ctx, NB200_K_MISC);
        if (ctx->compute == NB200_BF16)
            embed_kernel<bf16><<<B, 256, 0, ctx->stream>>>(ctx->d_tokens + (size_t)w0 * P, ctx->d_len + w0, P, (const DecodeDyn *)ctx->d_dyn, (const bf16 *)ctx->embed, ctx->embed_pos, V, d, ctx->dx);
        else
            embed_kernel<float><<<B, 256, 0, ctx->stream>>>(ctx->d_tokens + (size_t)w0 * P, ctx->d_len + w0, P, (const DecodeDyn *)ctx->d_dyn, (const float *)ctx->embed, ctx->embed_pos, V, d, ctx->dx);
    }
    for (int l = 0; l < c.decoder_layers; ++l) {
        const DecLayer &w = ctx->dec[l];
        char *skv = (char *)ctx->self_kv + ((size_t)l * c.max_batch + w0) * P * 2 * d * es;
        char *ckv = (char *)ctx->cross_kv + ((size_t)l * c.max_batch + w0) * T * 2 * d * es;
        // self attention (q, k scaled by hd^-0.25 as candle does at attention time; k has no bias)
        SkinnyEpi e{};
        e.bias = w.bqkv; e.out = ctx->dqkv; e.ldo = 3 * d; e.scale = qscale; e.n_scale = 2 * d;
        NB_TRY(skinny_ln(ctx, ctx->dx, d, w.ln1g, w.ln1b, nullptr, w.wqkv, 3 * d, d, B, e));
        NB_TRY(dec_attn(ctx, ctx->dqkv, 3 * d, skv, P, 0, true, ctx->dqkv, 3 * d, d, 2 * d, ctx->dattn, d, B));
        e = SkinnyEpi{};
        e.bias = w.bo; e.out = ctx->dx; e.ldo = d; e.residual = ctx->dx; e.ldr = d;
        NB_TRY(skinny(ctx, ctx->dattn, d, w.wo, d, d, B, e));
        // cross attention over the cached K/V of the audio features
        e = SkinnyEpi{};
        e.bias = w.cbq; e.out = ctx->dq; e.ldo = d; e.scale = qscale; e.n_scale = d;
        NB_TRY(skinny_ln(ctx, ctx->dx, d, w.lncg, w.lncb, nullptr, w.cwq, d, d, B, e));
        NB_TRY(dec_attn(ctx, ctx->dq, d, ckv, T, T, false, nullptr, 0, 0, 0, ctx->dattn, d, B));
        e = SkinnyEpi{};
        e.bias = w.cbo; e.out = ctx->dx; e.ldo = d; e.residual = ctx->dx; e.ldr = d;
        NB_TRY(skinny(ctx, ctx->dattn, d, w.cwo, d, d, B, e));
        // MLP
        e = SkinnyEpi{};
        e.bias = w.b1; e.out = ctx->dff; e.ldo = 4 * d; e.act = 1;
        NB_TRY(skinny_ln(ctx, ctx->dx, d, w.ln2g, w.ln2b, nullptr, w.w1, 4 * d, d, B, e));
        e = SkinnyEpi{};
        e.bias = w.b2; e.out = ctx->dx; e.ldo = d; e.residual = ctx->dx; e.ldr = d;
        NB_TRY(skinny(ctx, ctx->dff, 4 * d, w.w2, d, 4 * d, B, e));
    }
    if (want_logits) {  // final LayerNorm fused into the tied-embedding logits GEMV; block 0 leaves the hidden state in dhid
        SkinnyEpi e{};
        e.out = ctx->logits; e.ldo = V;
        NB_TRY(skinny_ln(ctx, ctx->dx, d, ctx->lndec_g, ctx->lndec_b, ctx->dhid, ctx->embed, V, d, B, e));
    } else {
        NB_TRY(launch_layernorm(ctx, ctx->dx, ctx->lndec_g, ctx->lndec_b, B, d, ctx->dhid, 0, nullptr));
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

int decoder_language(nb200_ctx *ctx, int n_langs) {
    KernelScope ks(ctx, NB200_K_DECODE_SELECT);
    uint32_t *ids = (uint32_t *)ctx->d_lang;
    float *probs = (float *)(ids + NB200_MAX_LANGS);
    int *best = (int *)(probs + NB200_MAX_LANGS);
    language_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->logits, ids, n_langs, probs, best);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int decoder_select(nb200_ctx *ctx, int n_windows, int greedy) {
    KernelScope ks(ctx, NB200_K_DECODE_SELECT);
    SelectParams sp;
    sp.logits = ctx->logits; sp.suppress = ctx->suppress; sp.tokens = ctx->d_tokens;
    sp.len = ctx->d_len; sp.last_ts = ctx->d_last_ts; sp.done = ctx->d_done; sp.nsampled = ctx->d_nsampled; sp.sumlp = ctx->d_sumlp;
    sp.V = ctx->cfg.vocab_size; sp.max_pos = ctx->cfg.max_target_positions;
    sp.dyn = (DecodeDyn *)ctx->d_dyn;
    sp.eot = ctx->tok.eot; sp.nts = ctx->tok.no_timestamps; sp.ts_zero = ctx->tok.ts_zero; sp.ts_one = ctx->tok.ts_one;
    if (greedy) {
        static_assert(SEL_CH == 32, "select_final_kernel maps one chunk per lane");
        select_stats_kernel<<<dim3(n_windows, SEL_CH), 256, 0, ctx->stream>>>(ctx->logits, sp.V, (float2 *)ctx->d_sel_ws);
        SelCand *wb = (SelCand *)((float2 *)ctx->d_sel_ws + (size_t)ctx->cfg.max_batch * SEL_CH);
        select_cand_kernel<<<dim3(n_windows, SEL_CH), 256, 0, ctx->stream>>>(sp, (const float2 *)ctx->d_sel_ws, wb);
        select_final_kernel<<<n_windows, 32, 0, ctx->stream>>>(sp, wb);
        ctx->launches += 2;
    } else {
        decode_select_kernel<<<n_windows, 1024, 0, ctx->stream>>>(sp);
    }
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
