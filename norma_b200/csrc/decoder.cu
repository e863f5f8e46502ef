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
#include "ptx.cuh"

#include <algorithm>

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

// ---- block reductions.  BAR = 0: the whole block (__syncthreads, blockDim.x threads); BAR > 0: the first NT threads of the block
// through named barrier BAR (the fused step's consumer threads)
template <int BAR, int NT>
__device__ __forceinline__ void block_sync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
}
template <int BAR = 0, int NT = 0>
__device__ __forceinline__ float block_max(float v, float *red) {
    const int nw = (NT ? NT : (int)blockDim.x) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    block_sync<BAR, NT>();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    block_sync<BAR, NT>();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
    return r;
}
template <int BAR = 0, int NT = 0>
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int nw = (NT ? NT : (int)blockDim.x) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    block_sync<BAR, NT>();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    block_sync<BAR, NT>();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += red[i];
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
struct __align__(8) SelCand { float sum_ts, max_text, best_a, best_b; int idx_a, idx_b; };

// phase A: per-chunk (max, sum exp(x - max)).  Block-uniform (b, c); BAR / NT as in block_max.
template <int BAR = 0, int NT = 0>
__device__ __forceinline__ void select_stats_body(const float *__restrict__ logits, int V, float2 *__restrict__ ws_a, int b, int c, int nch, float *red) {
    const int tid = threadIdx.x, nt = NT ? NT : (int)blockDim.x;
    const int per = (V + nch - 1) / nch, lo = c * per, hi = min(V, lo + per);
    const float *x = logits + (size_t)b * V;
    float lm = -INFINITY;
#pragma unroll 1
    for (int i = lo + tid; i < hi; i += nt) lm = fmaxf(lm, __ldcg(x + i));
    const float mx = block_max<BAR, NT>(lm, red);
    float ls = 0.f;
#pragma unroll 1
    for (int i = lo + tid; i < hi; i += nt) ls += expf(__ldcg(x + i) - mx);
    const float sm = block_sum<BAR, NT>(ls, red);
    if (tid == 0) ws_a[b * nch + c] = make_float2(mx, sm);
}
__global__ void __launch_bounds__(256)
select_stats_kernel(const float *__restrict__ logits, int V, float2 *__restrict__ ws_a) {
    __shared__ float red[32];
    select_stats_body(logits, V, ws_a, blockIdx.x, blockIdx.y, SEL_CH, red);
}

// which rule(s) of norma's token selection can apply to window b (mode 0 / 1 are known from the token state; otherwise both 2 and 3 are prepared)
__device__ __forceinline__ void cand_modes(const SelectParams &sp, int b, int len, int last_ts, int &mode_a, int &mode_b) {
    const int nts = (int)sp.nts;
    mode_b = -1;
    if (last_ts < 0) mode_a = 0;
    else {
        const uint32_t l_tok = __ldcg(sp.tokens + (size_t)b * sp.max_pos + len - 1);
        const bool has_sl = len >= 2;
        const uint32_t sl_tok = has_sl ? __ldcg(sp.tokens + (size_t)b * sp.max_pos + len - 2) : 0u;
        if ((int)l_tok > nts) mode_a = (has_sl && sl_tok >= sp.eot) ? 1 : 2;
        else { mode_a = 2; mode_b = 3; }
    }
}
// one vocabulary entry (probability p, suppression mask sup) folded into a thread's share of sum_ts / max_text and its arg-max candidates
__device__ __forceinline__ void cand_update(const SelectParams &sp, int mode_a, int mode_b, int last_ts, int i, float p, float sup, float &ts, float &mt,
                                            float &ba, float &bb, int &ia, int &ib) {
    const int nts = (int)sp.nts;
    auto masked = [&](int mode) -> bool {
        if (mode == 0) return i < (int)sp.ts_zero || i > (int)sp.ts_one;
        bool mk = sup != 0.f;
        if (mode == 1) mk |= i > nts;
        else if (mode == 2) mk |= i <= nts || i <= last_ts;
        else mk |= (i > nts && i <= last_ts);
        return mk;
    };
    if (mode_b >= 0) {
        const float ps = p + sup;
        if (i > nts) ts += ps;
        else if (i < nts) mt = fmaxf(mt, ps);
    }
    const float pa = masked(mode_a) ? -INFINITY : p;
    if (pa >= ba) { ba = pa; ia = i; }
    if (mode_b >= 0) {
        const float pb = masked(mode_b) ? -INFINITY : p;
        if (pb >= bb) { bb = pb; ib = i; }
    }
}

// phase B: with the global (max, sum) every chunk yields its share of sum_ts / max_text and the arg-max candidates of the
// rule(s) that can still apply: mode 0 / 1 are known from the token state; otherwise both 2 and 3 are prepared
template <int BAR = 0, int NT = 0>
__device__ __forceinline__ void select_cand_body(const SelectParams &sp, const float2 *__restrict__ ws_a, SelCand *__restrict__ ws_b, int b, int c, int nch,
                                                 float *red, int *red_i) {
    const int tid = threadIdx.x, nt = NT ? NT : (int)blockDim.x;
    if (__ldcg(sp.done + b)) return;  // (state and logits change every step of a multi-step launch: read them from L2)
    const int V = sp.V;
    // global (max, sum) from the chunk partials, folded by the whole block (one partial per thread: nch can be as large as the grid)
    float pm = -INFINITY;
#pragma unroll 1
    for (int k = tid; k < nch; k += nt) pm = fmaxf(pm, __ldcg(&ws_a[b * nch + k].x));
    const float M = block_max<BAR, NT>(pm, red);
    float ps = 0.f;
#pragma unroll 1
    for (int k = tid; k < nch; k += nt) ps += __ldcg(&ws_a[b * nch + k].y) * expf(__ldcg(&ws_a[b * nch + k].x) - M);
    const float S = block_sum<BAR, NT>(ps, red);
    const int len = __ldcg(sp.len + b), last_ts = __ldcg(sp.last_ts + b);
    int mode_a, mode_b;
    cand_modes(sp, b, len, last_ts, mode_a, mode_b);
    const int per = (V + nch - 1) / nch, lo = c * per, hi = min(V, lo + per);
    const float *x = sp.logits + (size_t)b * V;
    float ts = 0.f, mt = -INFINITY, ba = -INFINITY, bb = -INFINITY;
    int ia = -1, ib = -1;
#pragma unroll 1
    for (int i = lo + tid; i < hi; i += nt) cand_update(sp, mode_a, mode_b, last_ts, i, expf(__ldcg(x + i) - M) / S, sp.suppress[i], ts, mt, ba, bb, ia, ib);
    auto arg_reduce = [&](float &bv, int &bi) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
        }
        block_sync<BAR, NT>();
        if ((tid & 31) == 0) { red[tid >> 5] = bv; red_i[tid >> 5] = bi; }
        block_sync<BAR, NT>();
        if (tid == 0)
#pragma unroll 1
            for (int k = 1; k < (nt >> 5); ++k)
                if (red[k] > bv || (red[k] == bv && red_i[k] > bi)) { bv = red[k]; bi = red_i[k]; }
    };
    arg_reduce(ba, ia);
    if (mode_b >= 0) arg_reduce(bb, ib);
    const float sum_ts = block_sum<BAR, NT>(ts, red);
    const float max_text = block_max<BAR, NT>(mt, red);
    if (tid == 0) ws_b[b * nch + c] = SelCand{sum_ts, max_text, ba, bb, ia, ib};
}
__global__ void __launch_bounds__(256)
select_cand_kernel(SelectParams sp, const float2 *__restrict__ ws_a, SelCand *__restrict__ ws_b) {
    __shared__ float red[32];
    __shared__ int red_i[32];
    select_cand_body(sp, ws_a, ws_b, blockIdx.x, blockIdx.y, SEL_CH, red, red_i);
}

// phase C: one warp per window folds the chunk results, applies norma's rule and updates the decoding state
// one warp, warp-uniform b; `advance` = this warp also moves the device-resident position to the next step
__device__ __forceinline__ void select_final_body(const SelectParams &sp, const SelCand *__restrict__ ws_b, int b, int lane, bool advance, int nch) {
    const int max_new = sp.dyn->max_new;
    __syncwarp();
    if (advance && lane == 0) sp.dyn->pos += 1;
    if (__ldcg(sp.done + b)) return;
    SelCand cd{0.f, -INFINITY, -INFINITY, -INFINITY, -1, -1};
#pragma unroll 1
    for (int c = lane; c < nch; c += 32) {  // chunks of this lane, in ascending order
        SelCand o;
        {
            const float2 *pc = (const float2 *)(ws_b + b * nch + c);
            const float2 u0 = __ldcg(pc), u1 = __ldcg(pc + 1), u2 = __ldcg(pc + 2);
            o.sum_ts = u0.x; o.max_text = u0.y; o.best_a = u1.x; o.best_b = u1.y; o.idx_a = __float_as_int(u2.x); o.idx_b = __float_as_int(u2.y);
        }
        cd.sum_ts += o.sum_ts;
        cd.max_text = fmaxf(cd.max_text, o.max_text);
        if (o.best_a > cd.best_a || (o.best_a == cd.best_a && o.idx_a > cd.idx_a)) { cd.best_a = o.best_a; cd.idx_a = o.idx_a; }
        if (o.best_b > cd.best_b || (o.best_b == cd.best_b && o.idx_b > cd.idx_b)) { cd.best_b = o.best_b; cd.idx_b = o.idx_b; }
    }
    float ts = cd.sum_ts, mt = cd.max_text;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
    }
    const int len = __ldcg(sp.len + b), last_ts = __ldcg(sp.last_ts + b), nts = (int)sp.nts;
    bool use_b = false;
    if (last_ts >= 0) {
        const uint32_t l_tok = __ldcg(sp.tokens + (size_t)b * sp.max_pos + len - 1);
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

__global__ void __launch_bounds__(32)
select_final_kernel(SelectParams sp, const SelCand *__restrict__ ws_b) {
    select_final_body(sp, ws_b, blockIdx.x, threadIdx.x, blockIdx.x == 0, SEL_CH);
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

// =====================================================================================================================
// Fused decoder step (round 1d, bf16).  One decoder position used to be ~25 kernels (11 weight-streaming GEMVs worth 0.5 - 20 us of
// HBM time each, 8 attention launches, 3 select launches): at B = 1 the step took 264 us against a 34 us HBM floor, every kernel
// paying launch, drain and prologue latency.  Here the whole step is ONE cooperative kernel, one CTA per SM:
//   per layer: LN1+QKV | self-attn | out+res | LNc+q | cross-attn | out+res | LN2+fc1+GELU | fc2+res || LN+logits | stats | cand | final
// * a producer warp streams every weight matrix of the step, in order, through a 4 x 40 KB shared-memory ring with 1-D bulk copies
//   (rows of a [N][K] matrix are contiguous: a tile of R rows is one copy, no tensor map).  The weights do not depend on the
//   activations, so the producer never waits at a grid barrier: while the CTAs synchronise and stage the next phase's input the ring
//   is already filling with its weights, and HBM keeps streaming across phase boundaries (160 KB in flight per SM).
// * 16 consumer warps take the tiles out of the ring (one row, or a K-slice of a row, per warp), phases are separated by a grid
//   barrier (one release-add + acquire-poll on a monotonic counter; cooperative launch guarantees co-residency).
// * the split-K attention partials are merged by whichever worker finishes a (window, head) last (atomic ticket): no extra barrier.
// Same arithmetic as the separate kernels (shared device code for attention and select; the GEMV sums in a different order).
// The phase bodies are __noinline__ on purpose: inlined nine times the kernel was 316 KB of SASS and every phase ran out of a cold
// instruction cache (the first tile of a phase took 8 us, the following ones 2 us).
// =====================================================================================================================
constexpr int FS_WARPS = 16;                    // consumer warps
constexpr int FS_THREADS = (FS_WARPS + 1) * 32;  // + the producer warp
constexpr int FS_CTHREADS = FS_WARPS * 32;
constexpr int FS_MMA_WARPS = 4;                 // consumer warps that run the GEMV tiles (tensor cores: one warp outruns HBM)
constexpr int FS_STAGES = 4;
constexpr int FS_ROW_PAD = 16;                  // bytes added to a tile row so that ldmatrix rows fall in different banks
constexpr int FS_TILE_BYTES = 16 * (SK_KC_MAX * 2 + FS_ROW_PAD);  // 16 weight rows x d_model columns (one K chunk)
constexpr int FS_ATT_ST = 3;                    // stages (of four keys) a warp of fs_attn keeps in flight: 3 KB per warp in the staging buffer
constexpr int FS_XS_PAD = 8;                    // bf16 elements of padding per staged row: see xs_at
constexpr int FS_XS_BYTES = 16 * FS_ATT_ST * 1024;  // staged activations in bf16: 8 rows at K <= 2560, 4 rows at K = 5120 (fc2); the attention FIFOs
static_assert(FS_XS_BYTES >= 4 * (4 * SK_KC_MAX + FS_XS_PAD) * 2, "fc2 stages four rows per pass");
constexpr int FS_OFF_XS = FS_STAGES * FS_TILE_BYTES;
constexpr int FS_OFF_BAR = FS_OFF_XS + FS_XS_BYTES;
constexpr int FS_SMEM = FS_OFF_BAR + 2 * FS_STAGES * 8;
constexpr int FS_BAR_ID = 5;  // named barrier of the 512 consumer threads
constexpr int FS_SELF_SPLITS = 8, FS_MAX_SPLITS = 64;  // split-K of the decode attention (workspace: FS_MAX_SPLITS partials per window and head)

__device__ __forceinline__ void cbar() { asm volatile("bar.sync %0, %1;" ::"n"(FS_BAR_ID), "n"(FS_CTHREADS) : "memory"); }

// grid barrier over the consumer threads of every CTA: `target` = arrivals expected so far on the monotonic counter
__device__ __forceinline__ void grid_sync(unsigned *counter, unsigned &target) {
    cbar();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        unsigned seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while ((int)(seen - target) < 0);
    }
    cbar();
}

struct FusedArgs {
    int B, d, V, P, T, H, L, max_batch;
    const DecLayer *layers;  // device copy of ctx->dec
    const bf16 *embed;
    const float *embed_pos, *lndec_g, *lndec_b;
    uint32_t *tokens;
    int *len;
    DecodeDyn *dyn;
    float *dx, *dqkv, *dattn, *dq, *dff, *dhid, *logits, *attn_ws;
    unsigned *attn_cnt;  // [max_batch][H] tickets, zero between uses
    bf16 *self_kv, *cross_kv;
    SelectParams sp;
    float2 *sel_a;
    SelCand *sel_b;
    unsigned *sync_counter;
    unsigned *sync_epoch;  // value the counter has when a launch starts: read by everyone at entry, advanced by CTA 0 at exit (no host-side
                           // argument changes between steps, so a step can be replayed from a CUDA graph)
    int n_steps;         // decoder positions handled by this launch (the CTAs stay resident: starting 148 CTAs of 206 KB costs ~60 us)
    int cross_splits;    // split-K of the cross attention: enough (window, head, split) items for every warp of the grid
    float qscale;
    unsigned long long *tdbg;  // NB200_DECODE_TIMING: %globaltimer of CTA 0 after every phase
};

#ifdef NB200_DECODE_TIMING
#define DEC_STAMP() do { if (a.tdbg && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); a.tdbg[stamp_i] = t_; a.tdbg[32 + stamp_i] = (unsigned long long)clock64(); ++stamp_i; } } while (0)
#define DEC_CLK(x) x = clock64()
#define DEC_FINE(tag) do { if (a.tdbg && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); const unsigned long long k_ = a.tdbg[63]; if (k_ < 400) { a.tdbg[64 + 2 * k_] = t_; a.tdbg[65 + 2 * k_] = (tag); a.tdbg[63] = k_ + 1; } } } while (0)
#else
#define DEC_STAMP()
#define DEC_FINE(tag)
#define DEC_CLK(x)
#endif

// the weight matrices of one step in streaming order: 6 per layer, then the tied embedding
struct FsJob { const bf16 *W; int N, K; };
__device__ __forceinline__ FsJob fs_job(const FusedArgs &a, int j) {
    const int d = a.d;
    if (j == 6 * a.L) return FsJob{a.embed, a.V, d};
    const DecLayer &w = a.layers[j / 6];
    switch (j % 6) {
        case 0: return FsJob{(const bf16 *)w.wqkv, 3 * d, d};
        case 1: return FsJob{(const bf16 *)w.wo, d, d};
        case 2: return FsJob{(const bf16 *)w.cwq, d, d};
        case 3: return FsJob{(const bf16 *)w.cwo, d, d};
        case 4: return FsJob{(const bf16 *)w.w1, 4 * d, d};
        default: return FsJob{(const bf16 *)w.w2, d, 4 * d};
    }
}
__device__ __forceinline__ int fs_row_cap(int K) { return min(SK_MB, FS_XS_BYTES / (2 * (K + FS_XS_PAD))); }  // activation rows staged per pass

// producer: every tile this CTA will consume, in consumption order.  A tile = 16 consecutive weight rows x one K chunk of d_model
// columns; rows are copied one by one (2 d_model bytes each) to a padded pitch so that ldmatrix is bank-conflict free.
__device__ __noinline__ void fs_produce(const FusedArgs &a, uint32_t ring, uint32_t full0, uint32_t empty0, volatile int *go) {
    unsigned cnt = 0;
    const int KC = a.d;
    const uint32_t pitch = (uint32_t)KC * 2 + FS_ROW_PAD;
#pragma unroll 1
    for (int step = 0; step < a.n_steps; ++step) {
        // the consumers decide after the previous step's select whether there is another step (every window done: stop); no tile of a
        // step is requested before that, so nothing is in flight when the CTA exits
        int g;
        while ((g = *go) <= step && g >= 0) __nanosleep(64);
        if (g < 0) return;
#pragma unroll 1
        for (int j = 0; j <= 6 * a.L; ++j) {
            const FsJob job = fs_job(a, j);
            const int n_groups = (job.N + 15) / 16, n_chunks = job.K / KC;
            const int cap = fs_row_cap(job.K), passes = (a.B + cap - 1) / cap;
#pragma unroll 1
            for (int p = 0; p < passes; ++p)
#pragma unroll 1
                for (int t = blockIdx.x; t < n_groups; t += gridDim.x)
#pragma unroll 1
                    for (int c = 0; c < n_chunks; ++c, ++cnt) {
                        const unsigned s = cnt % FS_STAGES;
                        if (cnt >= FS_STAGES) ptx::mbar_wait(empty0 + 8 * s, ((cnt / FS_STAGES) & 1) ^ 1);
                        const int rows = min(16, job.N - t * 16);
                        ptx::mbar_expect_tx(full0 + 8 * s, (uint32_t)rows * KC * 2);
                        const bf16 *src = job.W + (size_t)t * 16 * job.K + (size_t)c * KC;
#pragma unroll 1
                        for (int r = 0; r < rows; ++r) ptx::bulk_load_1d(ring + s * FS_TILE_BYTES + r * pitch, src + (size_t)r * job.K, (uint32_t)KC * 2, full0 + 8 * s);
                    }
        }
    }
}

// rows [0, Bd) of the phase input -> xs[Bd][K] in bf16 (the B operand of the tensor-core GEMV).  ln_g != nullptr: LayerNorm over the
// row (K = d_model); from_tokens: the row is the embedding of the token at `pos` plus the positional row (first phase of the step;
// CTA 0 also leaves it in dx for the residual).  This code runs once per phase, i.e. from a cold instruction cache, so it has to stay
// short: FS_STAGE_NV pieces per lane is the only unrolled dimension.
// 16-byte pieces a thread keeps in flight while staging.  The LayerNorm version holds them from the load to the normalised store, and its
// register count is NOT its own business: with 5 pieces per lane (all eight rows at once) ptxas' interprocedural allocation left the key
// loop of fs_attn short of registers, its prefetched K / V rows were spilled — i.e. waited for — and both attentions of a layer got
// 3 - 5 us slower at B = 1, 23 us at B = 8.  3 pieces = four warps per row = four rows per round is the most that leaves fs_attn alone.
constexpr int FS_STAGE_NV = 3;
constexpr int FS_STAGE_NV_PLAIN = 3;  // (the plain copy's register count matters in the same way: 4 already costs fs_attn 8 spills)
// element index of (row m, column k) in the staging buffer.  The row pitch is 2 K + 16 bytes: at 2 K (a multiple of 128) the eight window
// rows a B fragment reads lie in the same four banks, an 8-way conflict on every fragment load that made the logits 1.5x slower at B = 8
__device__ __forceinline__ int xs_at(int m, int k, int K) { return m * (K + FS_XS_PAD) + k; }

__device__ __noinline__ void fs_stage_plain(const float *__restrict__ x, int ldx, int K, int Bd, bf16 *xs) {
    const int tid = threadIdx.x;
    cbar();  // the previous phase is done with xs
    // plain f32 -> bf16 copy.  Every thread requests FS_STAGE_NV_PLAIN independent 16-byte pieces before it touches the first one: a rolled
    // load-convert-store loop pays one L2 round trip (~0.7 us) per iteration, which made a 4-row pass of fc2 cost 7 us
    const int n4row = K >> 2, n4 = Bd * n4row;
    const unsigned rcp = 0xffffffffu / (unsigned)n4row + 1u;  // idx / n4row == umulhi(idx, rcp) for idx < 2^16
#pragma unroll 1
    for (int base = tid; base < n4; base += FS_CTHREADS * FS_STAGE_NV_PLAIN) {
        float4 v[FS_STAGE_NV_PLAIN];
#pragma unroll
        for (int j = 0; j < FS_STAGE_NV_PLAIN; ++j) {
            const int idx = base + j * FS_CTHREADS;
            if (idx < n4) {
                const int m = (int)__umulhi((unsigned)idx, rcp);
                v[j] = __ldcg((const float4 *)(x + (size_t)m * ldx) + (idx - m * n4row));  // written by other SMs during this launch: L2, not L1
            }
        }
#pragma unroll
        for (int j = 0; j < FS_STAGE_NV_PLAIN; ++j) {
            const int idx = base + j * FS_CTHREADS;
            if (idx < n4) {
                const int m = (int)__umulhi((unsigned)idx, rcp);
                __nv_bfloat162 lo = __floats2bfloat162_rn(v[j].x, v[j].y), hi = __floats2bfloat162_rn(v[j].z, v[j].w);
                *(uint2 *)&xs[xs_at(m, 4 * (idx - m * n4row), K)] = make_uint2(*(uint32_t *)&lo, *(uint32_t *)&hi);
            }
        }
    }
    cbar();
}

// LayerNorm (candle_nn: mean, then the centred variance, eps 1e-5), up to four rows per round: 16 / pow2(rows) >= 4 warps share a row,
// every lane keeps its <= FS_STAGE_NV float4 of the row in registers from the load to the normalised store (one L2 round trip and two
// barriers per round).  The first version went through a rolled load loop — one L2 round trip per iteration — and cost 5 us + 1.5 us per row.
// One call = one round of Bd <= 4 rows (the caller shifts the pointers).  Keep this function's register count low: see FS_STAGE_NV.
__device__ __noinline__ void fs_stage(const FusedArgs &a, const float *__restrict__ x, int ldx, int K, int m_first, int Bd, const float *__restrict__ ln_g,
                                      const float *__restrict__ ln_b, float *__restrict__ ln_out, bool from_tokens, int pos, bf16 *xs, float *stats) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cbar();  // the previous phase is done with xs, the previous round with the statistics
    int p2 = 1;
    while (p2 < Bd) p2 <<= 1;
    const int wpr = FS_WARPS / p2, r = warp / wpr, n4 = K >> 2, stride = wpr * 32, q0 = (warp - r * wpr) * 32 + lane;
    {
    const int nj = (r < Bd && q0 < n4) ? (n4 - q0 + stride - 1) / stride : 0;  // float4 pieces of this lane
    // the row: x, or (first phase of the step) the positional row plus the embedding of the token at `pos`
    const float4 *xp = (const float4 *)(from_tokens ? a.embed_pos + (size_t)pos * K : x + (size_t)r * ldx) + q0;
    float4 v[FS_STAGE_NV];
#pragma unroll
    for (int j = 0; j < FS_STAGE_NV; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nj) v[j] = __ldcg(xp + j * stride);  // (x is written by other SMs during this launch: L2, not L1)
    }
    if (from_tokens) {
        const int b = m_first + r;
        uint32_t tok = 0u;
        if (nj > 0 && pos < __ldcg(a.len + b)) tok = __ldcg(a.tokens + (size_t)b * a.P + pos);
        if (tok >= (uint32_t)a.V) tok = 0;
        const uint2 *er = (const uint2 *)(a.embed + (size_t)tok * K) + q0;
        float4 *dxp = (float4 *)(a.dx + (size_t)b * K) + q0;
#pragma unroll
        for (int j = 0; j < FS_STAGE_NV; ++j)
            if (j < nj) {
                const uint2 eb = __ldg(er + j * stride);
                const float2 e01 = __bfloat1622float2(*(const __nv_bfloat162 *)&eb.x), e23 = __bfloat1622float2(*(const __nv_bfloat162 *)&eb.y);
                v[j].x += e01.x; v[j].y += e01.y; v[j].z += e23.x; v[j].w += e23.y;
                if (blockIdx.x == 0) dxp[j * stride] = v[j];
            }
    }
    float s1 = 0.f;
#pragma unroll
    for (int j = 0; j < FS_STAGE_NV; ++j)
        if (j < nj) s1 += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    if (lane == 0) stats[warp] = s1;
    cbar();
    float mean = 0.f;
#pragma unroll 1
    for (int w = 0; w < wpr; ++w) mean += stats[r * wpr + w];
    mean /= (float)K;
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < FS_STAGE_NV; ++j)
        if (j < nj) {
            const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
            s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) stats[FS_WARPS + warp] = s2;
    cbar();
    float var = 0.f;
#pragma unroll 1
    for (int w = 0; w < wpr; ++w) var += stats[FS_WARPS + r * wpr + w];
    const float rstd = rsqrtf(var / (float)K + 1e-5f);
    const float4 *gp = (const float4 *)ln_g + q0, *bp = (const float4 *)ln_b + q0;
    bf16 *xo = xs + xs_at(r, 4 * q0, K);
    float4 *lo_p = (ln_out && blockIdx.x == 0) ? (float4 *)(ln_out + (size_t)r * K) + q0 : nullptr;
#pragma unroll
    for (int j = 0; j < FS_STAGE_NV; ++j)
        if (j < nj) {
            const float4 g = __ldg(gp + j * stride), bt = __ldg(bp + j * stride);
            const float4 y = make_float4((v[j].x - mean) * rstd * g.x + bt.x, (v[j].y - mean) * rstd * g.y + bt.y,
                                         (v[j].z - mean) * rstd * g.z + bt.z, (v[j].w - mean) * rstd * g.w + bt.w);
            __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
            *(uint2 *)(xo + 4 * j * stride) = make_uint2(*(uint32_t *)&lo, *(uint32_t *)&hi);
            if (lo_p) lo_p[j * stride] = y;
        }
    }
    cbar();
}

// consumer side of one weight matrix: out[m][n] = epilogue(sum_k x[m][k] W[n][k]) for every window row, on the tensor cores:
// D[16 weight rows][8 window rows] += A[16 x 16 from the ring tile, ldmatrix] . B[16 x 8 from xs] (mma.sync m16n8k16, fp32 accumulate).
// The activations are rounded to bf16 like every other GEMM input of the bf16 build.  A 16-row group (all its K chunks) belongs to
// one of FS_MMA_WARPS warps; the first version did this with FFMA on 16 warps and spent ~300 issue slots per weight row, more than
// the 0.9 us a 40 KB tile takes to arrive from HBM.
__device__ __noinline__ void fs_gemv(const FusedArgs &a, int j, const float *x, int ldx, const float *ln_g, const float *ln_b, float *ln_out, bool from_tokens,
                                     int pos, SkinnyEpi e, uint8_t *smem, uint32_t full0, uint32_t empty0, unsigned &cnt, float *stats, float *kpart) {
    const FsJob job = fs_job(a, j);
    const int K = job.K, N = job.N, KC = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = (N + 15) / 16, n_chunks = K / KC, cap = fs_row_cap(K);
    const uint32_t pitch = (uint32_t)KC * 2 + FS_ROW_PAD;
    bf16 *xs = (bf16 *)(smem + FS_OFF_XS);
    const uint32_t ring = ptx::smem_u32(smem);
#pragma unroll 1
    for (int m0 = 0; m0 < a.B; m0 += cap) {
        const int Bd = min(cap, a.B - m0);
        if (ln_g) {
#pragma unroll 1
            for (int r0 = 0; r0 < Bd; r0 += 4)  // four rows per round
                fs_stage(a, x ? x + (size_t)(m0 + r0) * ldx : nullptr, ldx, K, m0 + r0, min(4, Bd - r0), ln_g, ln_b, ln_out ? ln_out + (size_t)(m0 + r0) * K : nullptr,
                         from_tokens, pos, xs + xs_at(r0, 0, K), stats);
        } else {
            fs_stage_plain(x + (size_t)m0 * ldx, ldx, K, Bd, xs);
        }
        int gi = 0;
        if (n_chunks == FS_MMA_WARPS) {
            // K = 4 d_model (fc2): a group's four K chunks are four consecutive ring tiles; MMA warp w takes chunk w of EVERY group, so the four
            // tiles are consumed concurrently instead of one after the other, and warp 0 folds the four partial fragments through `kpart`
#pragma unroll 1
            for (int t = blockIdx.x; t < n_groups; t += gridDim.x, ++gi, cnt += n_chunks) {
                if (warp >= FS_MMA_WARPS) continue;
                float acc[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
                float ebias[2] = {0.f, 0.f}, eres[4] = {0.f, 0.f, 0.f, 0.f};
                if (warp == 0) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int n = t * 16 + (lane >> 2) + (r >> 1) * 8, m = (lane & 3) * 2 + (r & 1);
                        if (n < N && m < Bd) {
                            if (e.bias && (r & 1) == 0) ebias[r >> 1] = e.bias[n];
                            if (e.residual) eres[r] = __ldcg(e.residual + (size_t)(m0 + m) * e.ldr + n);
                        }
                    }
                }
                const int bn = lane >> 2;
                const bf16 *xk = xs + xs_at(bn < Bd ? bn : 0, (lane & 3) * 2 + warp * KC, K);
                const unsigned ct = cnt + warp, st = ct % FS_STAGES;
                ptx::mbar_wait(full0 + 8 * st, (ct / FS_STAGES) & 1);
                const uint32_t arow = ring + st * FS_TILE_BYTES + (lane & 15) * pitch + (lane >> 4) * 16;
                // four k-steps per trip with the accumulator index a compile-time constant (d_model is a multiple of 64): with `acc[ks & 3]` in
                // a partially unrolled loop the remainder loop indexed the fragments dynamically and ptxas put all sixteen accumulators in
                // local memory — every mma.sync of the step sat between an LDL and an STL (ncu: 14 M local loads per 16 positions)
#pragma unroll 1
                for (int k4 = 0; k4 < KC / 64; ++k4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ks = k4 * 4 + u;
                        uint32_t a0, a1, a2, a3, b0 = 0u, b1 = 0u;
                        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(arow + ks * 32));
                        if (bn < Bd) {
                            b0 = *(const uint32_t *)(xk + ks * 16);
                            b1 = *(const uint32_t *)(xk + ks * 16 + 8);
                        }
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                     : "+f"(acc[u][0]), "+f"(acc[u][1]), "+f"(acc[u][2]), "+f"(acc[u][3])
                                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(empty0 + 8 * st);
                float4 *kp = (float4 *)kpart + (gi & 1) * 3 * 32;  // double-buffered by group parity
                const float4 mine = make_float4(acc[0][0] + acc[1][0] + acc[2][0] + acc[3][0], acc[0][1] + acc[1][1] + acc[2][1] + acc[3][1],
                                                acc[0][2] + acc[1][2] + acc[2][2] + acc[3][2], acc[0][3] + acc[1][3] + acc[2][3] + acc[3][3]);
                if (warp > 0) kp[(warp - 1) * 32 + lane] = mine;
                asm volatile("bar.sync 6, 128;" ::: "memory");  // the four MMA warps
                if (warp == 0) {
                    const float4 p1 = kp[lane], p2 = kp[32 + lane], p3 = kp[64 + lane];
                    const float vv[4] = {mine.x + p1.x + p2.x + p3.x, mine.y + p1.y + p2.y + p3.y, mine.z + p1.z + p2.z + p3.z, mine.w + p1.w + p2.w + p3.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int n = t * 16 + (lane >> 2) + (r >> 1) * 8, m = (lane & 3) * 2 + (r & 1);
                        if (n < N && m < Bd) {
                            float v = vv[r] + ebias[r >> 1];
                            if (n < e.n_scale) v *= e.scale;
                            if (e.act) v = gelu_tanh_precise(v);
                            v += eres[r];
                            e.out[(size_t)(m0 + m) * e.ldo + n] = v;
                        }
                    }
                }
            }
            continue;
        }
#pragma unroll 1
        for (int t = blockIdx.x; t < n_groups; t += gridDim.x, ++gi) {
            if (warp != (gi % FS_MMA_WARPS)) {  // not this warp's group (warps >= FS_MMA_WARPS own none): just keep the tile count
                cnt += n_chunks;
                continue;
            }
            float acc[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
            // bias and residual of this lane's four outputs, requested now so that their L2 round trips hide behind the MMA loop
            float ebias[2] = {0.f, 0.f}, eres[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int n = t * 16 + (lane >> 2) + (r >> 1) * 8, m = (lane & 3) * 2 + (r & 1);
                if (n < N && m < Bd) {
                    if (e.bias && (r & 1) == 0) ebias[r >> 1] = e.bias[n];
                    if (e.residual) eres[r] = __ldcg(e.residual + (size_t)(m0 + m) * e.ldr + n);
                }
            }
            const int bn = lane >> 2;  // window row (B operand column) this lane feeds
            const bf16 *xrow = xs + xs_at(bn < Bd ? bn : 0, (lane & 3) * 2, K);
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c, ++cnt) {
                const unsigned s = cnt % FS_STAGES;
                ptx::mbar_wait(full0 + 8 * s, (cnt / FS_STAGES) & 1);
                const uint32_t arow = ring + s * FS_TILE_BYTES + (lane & 15) * pitch + (lane >> 4) * 16;
                const bf16 *xk = xrow + c * KC;
#pragma unroll 1
                for (int k4 = 0; k4 < KC / 64; ++k4) {  // accumulator index static, see above
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ks = k4 * 4 + u;
                        uint32_t a0, a1, a2, a3, b0 = 0u, b1 = 0u;
                        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(arow + ks * 32));
                        if (bn < Bd) {
                            b0 = *(const uint32_t *)(xk + ks * 16);
                            b1 = *(const uint32_t *)(xk + ks * 16 + 8);
                        }
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                     : "+f"(acc[u][0]), "+f"(acc[u][1]), "+f"(acc[u][2]), "+f"(acc[u][3])
                                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(empty0 + 8 * s);  // the tile has been read: the stage can be refilled
            }
            // D fragment: rows lane / 4 and lane / 4 + 8 of the group, window rows 2 (lane % 4) and 2 (lane % 4) + 1
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float v = acc[0][r] + acc[1][r] + acc[2][r] + acc[3][r];
                const int n = t * 16 + (lane >> 2) + (r >> 1) * 8, m = (lane & 3) * 2 + (r & 1);
                if (n < N && m < Bd) {
                    v += ebias[r >> 1];
                    if (n < e.n_scale) v *= e.scale;
                    if (e.act) v = gelu_tanh_precise(v);
                    v += eres[r];
                    e.out[(size_t)(m0 + m) * e.ldo + n] = v;
                }
            }
        }
    }
}

// Self attention of the fused step when there are no more (window, head) pairs than CTAs (B <= 7 at 20 heads): one pair per CTA, its
// (at most 448) keys split over the 16 consumer warps, the partial (m, l, o[64]) folded through shared memory — no global workspace,
// fences or tickets as in the split-K version below (6.9 -> 4.8 us at B = 1).  Eight lanes share a key (16 bytes of K and of
// V per lane), so a warp advances four keys per iteration; the next iteration's rows are requested before this one's softmax update.
__device__ __forceinline__ void osm_merge_fast(float &m, float &l, float *acc, float m2, float l2, const float *acc2) {
    const float mn = fmaxf(m, m2);
    const float a = (m == -INFINITY) ? 0.f : __expf(m - mn), b = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
    l = l * a + l2 * b;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = acc[t] * a + acc2[t] * b;
    m = mn;
}

// (few arguments on purpose, here and in fs_attn: the phase functions get the registers the kernel body does not keep live across the
// call, and the key loop spills its prefetched rows — which serialises the prefetch — as soon as it is a handful short.  `append`: self
// attention, q = the [B][3 d] QKV rows whose k and v parts join the cache at position n_keys - 1.)
__device__ __noinline__ void fs_attn_cta(const float *__restrict__ q, int ldq, bf16 *__restrict__ cache, int Tmax, int d, int B, int n_keys, bool append,
                                         float *__restrict__ out, float *sm) {
    const int H = d / HEAD_DIM, ldo = d, ldkv = ldq, koff = d, voff = 2 * d;
    const float *newkv = q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, l8 = lane & 7, grp = lane >> 3;
    const unsigned gmask = 0xffu << (grp * 8);
#pragma unroll 1
    for (int it = blockIdx.x; it < B * H; it += gridDim.x) {
        const int h = it % H, b = it / H;
        bf16 *cb = cache + (size_t)b * Tmax * 2 * d + h * HEAD_DIM + l8 * 8;
        float qr[8];
        {
            const float *qp = q + (size_t)b * ldq + h * HEAD_DIM + l8 * 8;
            const float4 qa = __ldcg((const float4 *)qp), qc = __ldcg((const float4 *)qp + 1);
            qr[0] = qa.x; qr[1] = qa.y; qr[2] = qa.z; qr[3] = qa.w; qr[4] = qc.x; qr[5] = qc.y; qr[6] = qc.z; qr[7] = qc.w;
        }
        const int chunk = (n_keys + FS_WARPS - 1) / FS_WARPS, k0 = warp * chunk, k1 = min(n_keys, k0 + chunk);
        float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.f;
        uint4 kn = make_uint4(0u, 0u, 0u, 0u), vn = kn;
        auto fetch = [&](int jj) {  // K / V row jj for this lane; the position being decoded comes from the QKV GEMV and joins the cache
            if (append && jj == n_keys - 1) {
                const float *kp = newkv + (size_t)b * ldkv + koff + h * HEAD_DIM + l8 * 8, *vp = newkv + (size_t)b * ldkv + voff + h * HEAD_DIM + l8 * 8;
                const float4 k0f = __ldcg((const float4 *)kp), k1f = __ldcg((const float4 *)kp + 1), v0f = __ldcg((const float4 *)vp), v1f = __ldcg((const float4 *)vp + 1);
                __nv_bfloat162 t0 = __floats2bfloat162_rn(k0f.x, k0f.y), t1 = __floats2bfloat162_rn(k0f.z, k0f.w), t2 = __floats2bfloat162_rn(k1f.x, k1f.y),
                               t3 = __floats2bfloat162_rn(k1f.z, k1f.w);
                kn = make_uint4(*(uint32_t *)&t0, *(uint32_t *)&t1, *(uint32_t *)&t2, *(uint32_t *)&t3);
                t0 = __floats2bfloat162_rn(v0f.x, v0f.y); t1 = __floats2bfloat162_rn(v0f.z, v0f.w); t2 = __floats2bfloat162_rn(v1f.x, v1f.y);
                t3 = __floats2bfloat162_rn(v1f.z, v1f.w);
                vn = make_uint4(*(uint32_t *)&t0, *(uint32_t *)&t1, *(uint32_t *)&t2, *(uint32_t *)&t3);
                *(uint4 *)(cb + (size_t)jj * 2 * d) = kn;
                *(uint4 *)(cb + (size_t)jj * 2 * d + d) = vn;
            } else {
                kn = *(const uint4 *)(cb + (size_t)jj * 2 * d);
                vn = *(const uint4 *)(cb + (size_t)jj * 2 * d + d);
            }
        };
        int j = k0 + grp;
        if (j < k1) fetch(j);
#pragma unroll 1
        for (; j < k1; j += 4) {
            const uint4 ku = kn, vu = vn;
            if (j + 4 < k1) fetch(j + 4);
            const float2 kf0 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.x), kf1 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.y);
            const float2 kf2 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.z), kf3 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.w);
            float sd = qr[0] * kf0.x + qr[1] * kf0.y + qr[2] * kf1.x + qr[3] * kf1.y + qr[4] * kf2.x + qr[5] * kf2.y + qr[6] * kf3.x + qr[7] * kf3.y;
            sd += __shfl_xor_sync(gmask, sd, 1);  // group-local mask: the four groups of a warp run different trip counts
            sd += __shfl_xor_sync(gmask, sd, 2);
            sd += __shfl_xor_sync(gmask, sd, 4);
            const float mn = fmaxf(m, sd);
            const float al = __expf(m - mn), pw = __expf(sd - mn);  // m = -inf on the first key: al = 0
            const float2 vf0 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.x), vf1 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.y);
            const float2 vf2 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.z), vf3 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.w);
            l = l * al + pw;
            acc[0] = acc[0] * al + pw * vf0.x; acc[1] = acc[1] * al + pw * vf0.y; acc[2] = acc[2] * al + pw * vf1.x; acc[3] = acc[3] * al + pw * vf1.y;
            acc[4] = acc[4] * al + pw * vf2.x; acc[5] = acc[5] * al + pw * vf2.y; acc[6] = acc[6] * al + pw * vf3.x; acc[7] = acc[7] * al + pw * vf3.y;
            m = mn;
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {  // the four key groups of the warp
            float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o), a2[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) a2[t] = __shfl_xor_sync(0xffffffffu, acc[t], o);
            osm_merge_fast(m, l, acc, m2, l2, a2);
        }
        if (grp == 0) {
            float *o = sm + (warp * 8 + l8) * 10;
            o[0] = m; o[1] = l;
#pragma unroll
            for (int t = 0; t < 8; ++t) o[2 + t] = acc[t];
        }
        cbar();
        if (warp == 0 && grp == 0) {  // the 16 warps
#pragma unroll 1
            for (int w2 = 1; w2 < FS_WARPS; ++w2) {
                const float *o = sm + (w2 * 8 + l8) * 10;
                osm_merge_fast(m, l, acc, o[0], o[1], o + 2);
            }
            float *op = out + (size_t)b * ldo + h * HEAD_DIM + l8 * 8;
            const float inv = 1.0f / l;
            *(float4 *)op = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
            *((float4 *)op + 1) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
        }
        cbar();  // sm is reused by the next pair
    }
}

// Split-K single-query attention of the fused step: one (head, window, split) item per WARP.  A lane owns two of the 64 head
// dimensions, so a K (or V) row is one coalesced 128-byte access; the warp walks the keys of its split two at a time with an online
// softmax, writes (m, l, o[64]) to the workspace, and the warp that takes the last ticket of a (window, head) folds the S partials
// into `out` (no grid barrier between the splits and the merge).  Rolled loops, ~2 KB of code: the 128-thread version shared with the
// stand-alone kernels was 17 KB that each phase executed once, out of a cold instruction cache.
__device__ __noinline__ void fs_attn(const float *__restrict__ q, int ldq, bf16 *__restrict__ cache, int Tmax, int d, int B, int n_keys, bool append,
                                     float *__restrict__ ws, unsigned *__restrict__ cnt, int S, float *__restrict__ out, uint8_t *stage) {
    const int H = d / HEAD_DIM, ldo = d;
    const int lane = threadIdx.x & 31, l8 = lane & 7, grp = lane >> 3, gw = blockIdx.x * FS_WARPS + (threadIdx.x >> 5), n_warps = gridDim.x * FS_WARPS;
    const unsigned gmask = 0xffu << (grp * 8);
    const int n_items = H * B * S;
    // this lane's FIFO slots in the (idle) activation staging buffer: FS_ATT_ST stages of [K 16 B | V 16 B] per lane.  Eight lanes share a
    // key (16 bytes of the K and of the V row each), so one stage is four keys of the warp; the rows travel by cp.async, FS_ATT_ST stages
    // ahead — in registers (the first version) four keys per warp were all that could be in flight, 2.4 MB over the whole GPU, which is
    // 2 TB/s at HBM latency: 30 us for the cross attention of eight windows.
    const uint32_t slot = ptx::smem_u32(stage) + (threadIdx.x >> 5) * (FS_ATT_ST * 1024) + lane * 16;
#pragma unroll 1
    for (int it = gw; it < n_items; it += n_warps) {
        const int sp = it % S, h = (it / S) % H, b = it / (S * H);
        const int chunk = (n_keys + S - 1) / S, k0 = sp * chunk;
        int k1 = min(n_keys, k0 + chunk);
        bf16 *cb = cache + (size_t)b * Tmax * 2 * d + h * HEAD_DIM + l8 * 8;
        float qr[8];
        {
            const float *qp = q + (size_t)b * ldq + h * HEAD_DIM + l8 * 8;
            const float4 qa = __ldcg((const float4 *)qp), qc = __ldcg((const float4 *)qp + 1);
            qr[0] = qa.x; qr[1] = qa.y; qr[2] = qa.z; qr[3] = qa.w; qr[4] = qc.x; qr[5] = qc.y; qr[6] = qc.z; qr[7] = qc.w;
        }
        auto issue = [&](int i4) {  // stage i4 % FS_ATT_ST <- keys k0 + 4 i4 .. + 3 (always one group, possibly empty: the wait below counts groups)
            const int jj = k0 + 4 * i4 + grp;
            if (jj < k1) {
                const bf16 *src = cb + (size_t)jj * 2 * d;
                const uint32_t dst = slot + (i4 % FS_ATT_ST) * 1024;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512), "l"(src + d) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.f;
        const bool own_new = append && k1 == n_keys && k1 > k0;  // this split owns the position being decoded
        if (own_new) k1 = n_keys - 1;                           // the staged loop covers the cached keys only
#pragma unroll 1
        for (int i4 = 0; i4 < FS_ATT_ST; ++i4) issue(i4);
        if (own_new && grp == 0) {  // k, v of the new position come from the QKV GEMV (q = the [B][3 d] rows) and join the cache
            const float *kp = q + (size_t)b * ldq + d + h * HEAD_DIM + l8 * 8, *vp = kp + d;
            const float4 k0f = __ldcg((const float4 *)kp), k1f = __ldcg((const float4 *)kp + 1), v0f = __ldcg((const float4 *)vp), v1f = __ldcg((const float4 *)vp + 1);
            const __nv_bfloat162 t0 = __floats2bfloat162_rn(k0f.x, k0f.y), t1 = __floats2bfloat162_rn(k0f.z, k0f.w), t2 = __floats2bfloat162_rn(k1f.x, k1f.y),
                                 t3 = __floats2bfloat162_rn(k1f.z, k1f.w);
            const __nv_bfloat162 u0 = __floats2bfloat162_rn(v0f.x, v0f.y), u1 = __floats2bfloat162_rn(v0f.z, v0f.w), u2 = __floats2bfloat162_rn(v1f.x, v1f.y),
                                 u3 = __floats2bfloat162_rn(v1f.z, v1f.w);
            *(uint4 *)(cb + (size_t)(n_keys - 1) * 2 * d) = make_uint4(*(const uint32_t *)&t0, *(const uint32_t *)&t1, *(const uint32_t *)&t2, *(const uint32_t *)&t3);
            *(uint4 *)(cb + (size_t)(n_keys - 1) * 2 * d + d) = make_uint4(*(const uint32_t *)&u0, *(const uint32_t *)&u1, *(const uint32_t *)&u2, *(const uint32_t *)&u3);
            const float2 a = __bfloat1622float2(t0), c = __bfloat1622float2(t1), e = __bfloat1622float2(t2), g = __bfloat1622float2(t3);  // what every later step reads
            float sd = qr[0] * a.x + qr[1] * a.y + qr[2] * c.x + qr[3] * c.y + qr[4] * e.x + qr[5] * e.y + qr[6] * g.x + qr[7] * g.y;
            sd += __shfl_xor_sync(gmask, sd, 1);
            sd += __shfl_xor_sync(gmask, sd, 2);
            sd += __shfl_xor_sync(gmask, sd, 4);
            const float2 v0 = __bfloat1622float2(u0), v1 = __bfloat1622float2(u1), v2 = __bfloat1622float2(u2), v3 = __bfloat1622float2(u3);
            m = sd; l = 1.f;
            acc[0] = v0.x; acc[1] = v0.y; acc[2] = v1.x; acc[3] = v1.y; acc[4] = v2.x; acc[5] = v2.y; acc[6] = v3.x; acc[7] = v3.y;
        }
        const int n4 = (k1 - k0 + 3) >> 2;  // (<= 0: an empty split)
#pragma unroll 1
        for (int i4 = 0; i4 < n4; ++i4) {
            asm volatile("cp.async.wait_group %0;" ::"n"(FS_ATT_ST - 1) : "memory");  // the oldest stage has landed (each lane reads its own copies)
            if (k0 + 4 * i4 + grp < k1) {
                uint4 ku, vu;
                const uint32_t src = slot + (i4 % FS_ATT_ST) * 1024;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ku.x), "=r"(ku.y), "=r"(ku.z), "=r"(ku.w) : "r"(src));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(vu.x), "=r"(vu.y), "=r"(vu.z), "=r"(vu.w) : "r"(src + 512));
                const float2 kf0 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.x), kf1 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.y);
                const float2 kf2 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.z), kf3 = __bfloat1622float2(*(const __nv_bfloat162 *)&ku.w);
                float sd = qr[0] * kf0.x + qr[1] * kf0.y + qr[2] * kf1.x + qr[3] * kf1.y + qr[4] * kf2.x + qr[5] * kf2.y + qr[6] * kf3.x + qr[7] * kf3.y;
                sd += __shfl_xor_sync(gmask, sd, 1);  // group-local mask: the last stage may not have a key for every group
                sd += __shfl_xor_sync(gmask, sd, 2);
                sd += __shfl_xor_sync(gmask, sd, 4);
                const float mn = fmaxf(m, sd);
                const float al = __expf(m - mn), pw = __expf(sd - mn);  // m = -inf on the first key: al = 0
                const float2 vf0 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.x), vf1 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.y);
                const float2 vf2 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.z), vf3 = __bfloat1622float2(*(const __nv_bfloat162 *)&vu.w);
                l = l * al + pw;
                acc[0] = acc[0] * al + pw * vf0.x; acc[1] = acc[1] * al + pw * vf0.y; acc[2] = acc[2] * al + pw * vf1.x; acc[3] = acc[3] * al + pw * vf1.y;
                acc[4] = acc[4] * al + pw * vf2.x; acc[5] = acc[5] * al + pw * vf2.y; acc[6] = acc[6] * al + pw * vf3.x; acc[7] = acc[7] * al + pw * vf3.y;
                m = mn;
            }
            issue(i4 + FS_ATT_ST);  // refill the stage just read
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // (only empty groups are left)
        __syncwarp();
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {  // the four key groups of the warp
            float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o), a2[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) a2[t] = __shfl_xor_sync(0xffffffffu, acc[t], o);
            osm_merge_fast(m, l, acc, m2, l2, a2);
        }
        float *wsb = ws + ((size_t)b * H + h) * S * ATT_WS;
        {
            float *o = wsb + (size_t)sp * ATT_WS;
            if (lane == 0) { o[0] = m; o[1] = l; }
            if (grp == 0) {
#pragma unroll
                for (int t = 0; t < 8; t += 2) *(float2 *)(o + 2 + l8 * 8 + t) = make_float2(acc[t], acc[t + 1]);
            }
        }
        __threadfence();
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) last = atomicAdd(cnt + b * H + h, 1u) == (unsigned)(S - 1);
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {  // every split of this (window, head) has landed: fold them
            __threadfence();
            // lane s holds (m, l) of splits s, s + 32: the weights 2^(m_s - M) come out of two shuffles per split, and the S loads of
            // the partial outputs below do not depend on each other
            float m_a = -INFINITY, l_a = 0.f, m_b = -INFINITY, l_b = 0.f;
            if (lane < S) { m_a = __ldcg(wsb + lane * ATT_WS); l_a = __ldcg(wsb + lane * ATT_WS + 1); }
            if (lane + 32 < S) { m_b = __ldcg(wsb + (lane + 32) * ATT_WS); l_b = __ldcg(wsb + (lane + 32) * ATT_WS + 1); }
            float M = fmaxf(m_a, m_b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
            const float w_a = m_a == -INFINITY ? 0.f : expf(m_a - M), w_b = m_b == -INFINITY ? 0.f : expf(m_b - M);  // empty splits weigh nothing
            float ll = l_a * w_a + l_b * w_b;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
            float o0 = 0.f, o1 = 0.f;
#pragma unroll 16
            for (int s2 = 0; s2 < S; ++s2) {
                const float al = __shfl_sync(0xffffffffu, s2 < 32 ? w_a : w_b, s2 & 31);
                const float2 ov = __ldcg((const float2 *)(wsb + s2 * ATT_WS + 2 + lane * 2));
                o0 += ov.x * al;
                o1 += ov.y * al;
            }
            *(float2 *)(out + (size_t)b * ldo + h * HEAD_DIM + lane * 2) = make_float2(o0 / ll, o1 / ll);
            if (lane == 0) cnt[b * H + h] = 0;  // ticket back to zero for the next attention of this launch
        }
    }
}

// ---- select of the fused step (B <= FS_SEL_MAXB windows): one window per WARP (vocabulary chunk = this CTA), so the windows of a
// batch are folded in parallel and without block barriers (the block-wide bodies take them one after the other: 14 + 48 us at B = 8).
// The statistics phase leaves the chunk's suppression mask (slot 0) and every window's logits (slot b + 1) in the idle activation staging
// buffer for the candidates phase of the same CTA.
constexpr int FS_SEL_PER = 384;  // largest vocabulary chunk (V / grid) this path caches: 12 values per lane
constexpr int FS_SEL_NP = 5;     // chunk partials per lane: grid <= 160
constexpr int FS_SEL_MAXB = FS_XS_BYTES / (FS_SEL_PER * 4) - 1;

__device__ __noinline__ void fs_select_stats_w(const FusedArgs &a, float *sm) {
    const int tid = threadIdx.x, lane = tid & 31, nch = gridDim.x, V = a.V;
    const int per = (V + nch - 1) / nch, lo = blockIdx.x * per, hi = min(V, lo + per);
    if (lo + tid < hi) sm[tid] = a.sp.suppress[lo + tid];  // per <= FS_SEL_PER <= FS_CTHREADS
#pragma unroll 1
    for (int b = tid >> 5; b < a.B; b += FS_WARPS) {
        const float *x = a.logits + (size_t)b * V + lo + lane;
        float *xl = sm + (b + 1) * FS_SEL_PER + lane;
        float v[FS_SEL_PER / 32], lm = -INFINITY;
#pragma unroll
        for (int j = 0; j < FS_SEL_PER / 32; ++j) {  // all the loads of the lane before the first use
            v[j] = lo + lane + 32 * j < hi ? __ldcg(x + 32 * j) : -INFINITY;
            lm = fmaxf(lm, v[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
        float ls = 0.f;
#pragma unroll
        for (int j = 0; j < FS_SEL_PER / 32; ++j)
            if (lo + lane + 32 * j < hi) {
                ls += expf(v[j] - lm);
                xl[32 * j] = v[j];
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
        if (lane == 0) a.sel_a[b * nch + blockIdx.x] = make_float2(lm, ls);
    }
}

__device__ __noinline__ void fs_select_cand_w(const FusedArgs &a, const float *sm) {
    const SelectParams &sp = a.sp;
    const int tid = threadIdx.x, lane = tid & 31, nch = gridDim.x, V = a.V;
    const int per = (V + nch - 1) / nch, lo = blockIdx.x * per, n = min(V, lo + per) - lo;
#pragma unroll 1
    for (int b = tid >> 5; b < a.B; b += FS_WARPS) {
        if (__ldcg(sp.done + b)) continue;
        const int len = __ldcg(sp.len + b), last_ts = __ldcg(sp.last_ts + b);
        float2 u[FS_SEL_NP];
        float M = -INFINITY;
#pragma unroll
        for (int j = 0; j < FS_SEL_NP; ++j) {
            u[j] = lane + 32 * j < nch ? __ldcg(a.sel_a + b * nch + lane + 32 * j) : make_float2(-INFINITY, 0.f);
            M = fmaxf(M, u[j].x);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
        float S = 0.f;
#pragma unroll
        for (int j = 0; j < FS_SEL_NP; ++j) S += u[j].y * expf(u[j].x - M);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
        int mode_a, mode_b;
        cand_modes(sp, b, len, last_ts, mode_a, mode_b);
        const float *xl = sm + (b + 1) * FS_SEL_PER;
        float ts = 0.f, mt = -INFINITY, ba = -INFINITY, bb = -INFINITY;
        int ia = -1, ib = -1;
#pragma unroll 1
        for (int i = lane; i < n; i += 32) cand_update(sp, mode_a, mode_b, last_ts, lo + i, expf(xl[i] - M) / S, sm[i], ts, mt, ba, bb, ia, ib);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, ba, o), pv = __shfl_xor_sync(0xffffffffu, bb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ia, o), pi = __shfl_xor_sync(0xffffffffu, ib, o);
            if (ov > ba || (ov == ba && oi > ia)) { ba = ov; ia = oi; }
            if (pv > bb || (pv == bb && pi > ib)) { bb = pv; ib = pi; }
            ts += __shfl_xor_sync(0xffffffffu, ts, o);
            mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
        }
        if (lane == 0) a.sel_b[b * nch + blockIdx.x] = SelCand{ts, mt, ba, bb, ia, ib};
    }
}

__global__ void __launch_bounds__(FS_THREADS, 1)
decoder_step_fused_kernel(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float red[32];
    __shared__ int red_i[32];
    __shared__ float ln_stats[32];
    __shared__ float att_sm[FS_WARPS * 8 * 10];  // per-warp (m, l, o[8]) partials of fs_attn_cta
    __shared__ __align__(16) float kpart[2 * 3 * 32 * 4];  // K-split partial fragments of fc2 (fs_gemv)
    __shared__ volatile int go_step;
    __shared__ DecLayer w_sm;
    const int tid = threadIdx.x;
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t full0 = sbase + FS_OFF_BAR, empty0 = full0 + 8 * FS_STAGES;
    if (tid == 0) {
        for (int s = 0; s < FS_STAGES; ++s) {
            ptx::mbar_init(full0 + 8 * s, 1);
            ptx::mbar_init(empty0 + 8 * s, 1);
        }
        ptx::fence_barrier_init();
        go_step = 0;
    }
    __syncthreads();
    if (tid >= FS_CTHREADS) {  // producer warp: free-running inside a step, never at a grid barrier
        if (tid == FS_CTHREADS) fs_produce(a, sbase, full0, empty0, &go_step);
        return;
    }
#ifdef NB200_DECODE_TIMING
    int stamp_i = 0;
#endif
    const int d = a.d, B = a.B, V = a.V, P = a.P, T = a.T, H = a.H;
    const int nch = gridDim.x;  // vocabulary chunks per window in the select: one per CTA
    unsigned cnt = 0, target = __ldcg(a.sync_epoch);
#pragma unroll 1
    for (int step = 0; step < a.n_steps; ++step) {
    if (step > 0) grid_sync(a.sync_counter, target);  // the previous step's tokens, lengths, done flags and position are visible
    bool all_done = true;
#pragma unroll 1
    for (int b = 0; b < B; ++b) all_done &= __ldcg(a.sp.done + b) != 0;
    const int pos = __ldcg(&a.dyn->pos);
    if (all_done || pos >= P) {  // grid-uniform: every CTA reads the same flags after the same barrier
        if (tid == 0) go_step = -1;
        break;
    }
    if (tid == 0) go_step = step + 1;
#ifdef NB200_DECODE_TIMING
    stamp_i = 0;
#endif
    DEC_STAMP();
    for (int l = 0; l < a.L; ++l) {
        // the layer's 20 pointers live in shared memory, not in registers: whatever the kernel body keeps live across the calls below is
        // taken from the register window of the (non-inlined) phase functions, and the attention loop spilled once it lost ~20 registers
        cbar();
        if (tid < (int)(sizeof(DecLayer) / 8)) ((unsigned long long *)&w_sm)[tid] = ((const unsigned long long *)(a.layers + l))[tid];
        cbar();
        const DecLayer &w = w_sm;
        bf16 *skv = a.self_kv + (size_t)l * a.max_batch * P * 2 * d;
        bf16 *ckv = a.cross_kv + (size_t)l * a.max_batch * T * 2 * d;
        SkinnyEpi e{};
        e.bias = w.bqkv; e.out = a.dqkv; e.ldo = 3 * d; e.scale = a.qscale; e.n_scale = 2 * d;
        fs_gemv(a, 6 * l + 0, l == 0 ? nullptr : a.dx, d, w.ln1g, w.ln1b, nullptr, l == 0, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        if (B * H <= (int)gridDim.x) fs_attn_cta(a.dqkv, 3 * d, skv, P, d, B, pos + 1, true, a.dattn, att_sm);
        else fs_attn(a.dqkv, 3 * d, skv, P, d, B, pos + 1, true, a.attn_ws, a.attn_cnt, FS_SELF_SPLITS, a.dattn, smem + FS_OFF_XS);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        e = SkinnyEpi{};
        e.bias = w.bo; e.out = a.dx; e.ldo = d; e.residual = a.dx; e.ldr = d;
        fs_gemv(a, 6 * l + 1, a.dattn, d, nullptr, nullptr, nullptr, false, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        e = SkinnyEpi{};
        e.bias = w.cbq; e.out = a.dq; e.ldo = d; e.scale = a.qscale; e.n_scale = d;
        fs_gemv(a, 6 * l + 2, a.dx, d, w.lncg, w.lncb, nullptr, false, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        // (1500 keys per pair are too many for one CTA's 16 loads in flight: 19.5 us against 14.6 us split over the grid)
        fs_attn(a.dq, d, ckv, T, d, B, T, false, a.attn_ws, a.attn_cnt, a.cross_splits, a.dattn, smem + FS_OFF_XS);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        e = SkinnyEpi{};
        e.bias = w.cbo; e.out = a.dx; e.ldo = d; e.residual = a.dx; e.ldr = d;
        fs_gemv(a, 6 * l + 3, a.dattn, d, nullptr, nullptr, nullptr, false, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        e = SkinnyEpi{};
        e.bias = w.b1; e.out = a.dff; e.ldo = 4 * d; e.act = 1;
        fs_gemv(a, 6 * l + 4, a.dx, d, w.ln2g, w.ln2b, nullptr, false, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
        e = SkinnyEpi{};
        e.bias = w.b2; e.out = a.dx; e.ldo = d; e.residual = a.dx; e.ldr = d;
        fs_gemv(a, 6 * l + 5, a.dff, 4 * d, nullptr, nullptr, nullptr, false, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
        grid_sync(a.sync_counter, target);
        DEC_STAMP();
    }
    {   // final LayerNorm fused into the tied-embedding logits; CTA 0 leaves the hidden state in dhid
        SkinnyEpi e{};
        e.out = a.logits; e.ldo = V;
        fs_gemv(a, 6 * a.L, a.dx, d, a.lndec_g, a.lndec_b, a.dhid, false, pos, e, smem, full0, empty0, cnt, ln_stats, kpart);
    }
    grid_sync(a.sync_counter, target);
    DEC_STAMP();
    // ---- greedy select (the three phases of select_*_kernel; FS_CTHREADS threads through the consumers' named barrier) ----
    fs_select_stats_w(a, (float *)(smem + FS_OFF_XS));
    grid_sync(a.sync_counter, target);
    DEC_STAMP();
    fs_select_cand_w(a, (const float *)(smem + FS_OFF_XS));
    grid_sync(a.sync_counter, target);
    DEC_STAMP();
    if (tid < 32)
        for (int b = blockIdx.x; b < B; b += gridDim.x) select_final_body(a.sp, a.sel_b, b, tid, b == 0, nch);
    DEC_STAMP();
    }  // step
    if (blockIdx.x == 0 && tid == 0) *a.sync_epoch = target;  // every CTA read the old value before its first barrier
#ifdef NB200_DECODE_TIMING
    if (a.tdbg && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); atomicMax(a.tdbg + 52, t_); }
#endif
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

// n_steps greedy decoder steps for windows [0, n_windows) as a single cooperative launch (embed .. logits .. select per step); the
// position is the device-resident one (DecodeDyn) and advances after every step; the launch ends early once every window is done.
// bf16 contexts only.
bool decoder_fused_supported(const nb200_ctx *ctx) {
    const nb200_config &c = ctx->cfg;
    const int nch = ctx->sm_count;  // one vocabulary chunk per CTA in the select (fs_select_*_w)
    if (nch < 1 || nch > 32 * FS_SEL_NP || (c.vocab_size + nch - 1) / nch > FS_SEL_PER) return false;
    return ctx->compute == NB200_BF16 && c.d_model % 16 == 0 && c.d_model <= SK_KC_MAX && SK_KC_MAX <= FS_STAGE_NV * 512 && c.decoder_attention_heads * HEAD_DIM == c.d_model;
}

int decoder_fused_max_windows() { return FS_SEL_MAXB; }  // windows whose logits chunk fits the staging buffer during the select

int decoder_fused_ws_floats(const nb200_ctx *ctx) { return ctx->cfg.max_batch * ctx->cfg.decoder_attention_heads * FS_MAX_SPLITS * ATT_WS; }

// one-time sizing of the cooperative grid and its workspaces (not capturable: called before any graph capture of the step)
int decoder_fused_prepare(nb200_ctx *ctx) {
    const nb200_config &c = ctx->cfg;
    if (!decoder_fused_supported(ctx)) return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "fused decoder step: d_model=%d, compute=%d", c.d_model, (int)ctx->compute);
    if (ctx->fused_ctas != 0) return NB200_OK;
    void *kern = (void *)decoder_step_fused_kernel;
    int per_sm = 0;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FS_THREADS, FS_SMEM));
    if (per_sm < 1) return nb200_fail(ctx, NB200_CUDA_ERROR, "fused decoder step does not fit an SM");
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_fused_sel_ws, (size_t)c.max_batch * ctx->sm_count * (sizeof(float2) + sizeof(SelCand))));
    ctx->allocs.push_back(ctx->d_fused_sel_ws);
    ctx->fused_ctas = ctx->sm_count;  // one CTA per SM; the cooperative launch fails rather than hangs if they cannot all be resident
    return NB200_OK;
}

int decoder_step_fused(nb200_ctx *ctx, int n_windows, int n_steps) {
    const nb200_config &c = ctx->cfg;
    NB_TRY(decoder_fused_prepare(ctx));
    if (n_windows < 1 || n_windows > FS_SEL_MAXB) return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "fused decoder step: %d windows (at most %d)", n_windows, FS_SEL_MAXB);
    void *kern = (void *)decoder_step_fused_kernel;
    FusedArgs a{};
    a.B = n_windows; a.d = c.d_model; a.V = c.vocab_size; a.P = c.max_target_positions; a.T = c.max_source_positions;
    a.H = c.decoder_attention_heads; a.L = c.decoder_layers; a.max_batch = c.max_batch;
    a.layers = (const DecLayer *)ctx->d_dec_layers;
    a.embed = (const bf16 *)ctx->embed; a.embed_pos = ctx->embed_pos; a.lndec_g = ctx->lndec_g; a.lndec_b = ctx->lndec_b;
    a.tokens = ctx->d_tokens; a.len = ctx->d_len; a.dyn = (DecodeDyn *)ctx->d_dyn;
    a.dx = ctx->dx; a.dqkv = ctx->dqkv; a.dattn = ctx->dattn; a.dq = ctx->dq; a.dff = ctx->dff; a.dhid = ctx->dhid; a.logits = ctx->logits;
    a.attn_ws = ctx->d_fused_attn_ws; a.attn_cnt = (unsigned *)ctx->d_fused_sync + 32;
    a.cross_splits = std::max(1, std::min(FS_MAX_SPLITS, (ctx->fused_ctas * FS_WARPS) / std::max(1, n_windows * c.decoder_attention_heads)));
    a.self_kv = (bf16 *)ctx->self_kv; a.cross_kv = (bf16 *)ctx->cross_kv;
    SelectParams &sp = a.sp;
    sp.logits = ctx->logits; sp.suppress = ctx->suppress; sp.tokens = ctx->d_tokens;
    sp.len = ctx->d_len; sp.last_ts = ctx->d_last_ts; sp.done = ctx->d_done; sp.nsampled = ctx->d_nsampled; sp.sumlp = ctx->d_sumlp;
    sp.V = c.vocab_size; sp.max_pos = c.max_target_positions; sp.dyn = (DecodeDyn *)ctx->d_dyn;
    sp.eot = ctx->tok.eot; sp.nts = ctx->tok.no_timestamps; sp.ts_zero = ctx->tok.ts_zero; sp.ts_one = ctx->tok.ts_one;
    a.sel_a = (float2 *)ctx->d_fused_sel_ws;
    a.sel_b = (SelCand *)((float2 *)ctx->d_fused_sel_ws + (size_t)c.max_batch * ctx->fused_ctas);
    a.sync_counter = (unsigned *)ctx->d_fused_sync;
    a.n_steps = n_steps;
    a.sync_epoch = (unsigned *)ctx->d_fused_sync + 16;  // its own 64 B, away from the counter's line
    a.qscale = powf((float)HEAD_DIM, -0.25f);
#ifdef NB200_DECODE_TIMING
    static unsigned long long *tdbg = nullptr;
    if (!tdbg) cudaMalloc(&tdbg, 1024 * 8);
    cudaMemsetAsync(tdbg + 63, 0, 8, ctx->stream);
    cudaMemsetAsync(tdbg + 50, 0xff, 8, ctx->stream);
    cudaMemsetAsync(tdbg + 51, 0, 16, ctx->stream);
    a.tdbg = tdbg;
#endif
    KernelScope ks(ctx, NB200_K_DECODE_GEMV);
    void *args[] = {(void *)&a};
    CUDA_TRY(ctx, cudaLaunchCooperativeKernel(kern, dim3(ctx->fused_ctas), dim3(FS_THREADS), args, FS_SMEM, ctx->stream));
#ifdef NB200_DECODE_TIMING
    {
        static int calls = 0;
        if (++calls % 4 == 0) {
            unsigned long long h[1024];
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpy(h, tdbg, sizeof h, cudaMemcpyDeviceToHost);
            const int n = 1 + 8 * c.decoder_layers + 4;
            fprintf(stderr, "[fused step] phases (us):");
            for (int i = 1; i < n; ++i) fprintf(stderr, " %.1f", (double)(h[i] - h[i - 1]) * 1e-3);
            fprintf(stderr, " | CTA starts spread over %.1f us, first start -> last end %.1f us", (double)(h[51] - h[50]) * 1e-3, (double)(h[52] - h[50]) * 1e-3);
            fprintf(stderr, " | total %.1f us, SM clock %.0f MHz\n", (double)(h[n - 1] - h[0]) * 1e-3, (double)(h[32 + n - 1] - h[32]) / ((double)(h[n - 1] - h[0]) * 1e-3));
        }
    }
#endif
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
