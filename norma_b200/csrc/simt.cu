// simt.cu — fp32 CUDA-core kernels: the F32 compute mode of the encoder (parity target 1e-4, SURVEY §8 d) and
// the bandwidth-bound stages shared by both modes (LayerNorm, dtype conversion).
//   * sgemm_kernel        C = A . W^T (+ fused bias / q,k scale / GELU / residual epilogue), fp32 FMA
//   * layernorm_kernel    candle_nn LayerNorm, eps 1e-5, one warp per row, f32 in, f32|bf16 out (HBM-bound)
//   * attention_simt      flash-style (online softmax) non-causal attention, fp32 math, f32|bf16 I/O
// These replace candle's CPU `gemm` + unfused elementwise ops reached from
// /root/reference/src/models/whisper/model.rs:455-464 (`Type::encoder_forward`).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// fp32 GEMM: 128x128x16 tiles, 256 threads, 8x8 register tile per thread, register-prefetched global loads
// ---------------------------------------------------------------------------------------------------------
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16;

__device__ __forceinline__ void epilogue_store(const Epilogue &e, int b, int r, int n, float v) {
    if (e.bias) v += e.bias[n];
    if (n < e.n_scale) v *= e.scale;
    if (e.act) v = gelu_tanh_precise(v);
    if (e.residual) v += e.residual[(long long)b * e.res_bs + (long long)r * e.ldr + n];
    long long o = (long long)b * e.out_bs + (long long)r * e.ldo + n;
    if (e.out_bf16) ((bf16 *)e.out)[o] = __float2bfloat16(v);
    else ((float *)e.out)[o] = v;
}

__global__ void __launch_bounds__(256)
sgemm_kernel(const float *__restrict__ A, const float *__restrict__ W, GemmShape s, Epilogue e) {
    __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int Mtot = s.batch * s.rows_per_batch;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader: rows lrow, lrow+64; k offset lk
    const int tx = tid & 15, ty = tid >> 4;

    const float *arow[2];
    const float *brow[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int m = m0 + lrow + 64 * i;
        if (m < Mtot) {
            int b = m / s.rows_per_batch, r = m - b * s.rows_per_batch;
            arow[i] = A + (long long)b * s.a_bs + (long long)r * s.lda;
        } else arow[i] = nullptr;
        int n = n0 + lrow + 64 * i;
        brow[i] = n < s.N ? W + (long long)n * s.K : nullptr;
    }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 pa[2], pb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            pa[i] = arow[i] ? *(const float4 *)(arow[i] + k0 + lk) : make_float4(0, 0, 0, 0);
            pb[i] = brow[i] ? __ldg((const float4 *)(brow[i] + k0 + lk)) : make_float4(0, 0, 0, 0);
        }
    };
    gload(0);
    for (int k0 = 0; k0 < s.K; k0 += SG_BK) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = lrow + 64 * i;
            As[lk + 0][r] = pa[i].x; As[lk + 1][r] = pa[i].y; As[lk + 2][r] = pa[i].z; As[lk + 3][r] = pa[i].w;
            Bs[lk + 0][r] = pb[i].x; Bs[lk + 1][r] = pb[i].y; Bs[lk + 2][r] = pb[i].z; Bs[lk + 3][r] = pb[i].w;
        }
        __syncthreads();
        if (k0 + SG_BK < s.K) gload(k0 + SG_BK);
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float4 a0 = *(const float4 *)&As[k][ty * 4], a1 = *(const float4 *)&As[k][64 + ty * 4];
            float4 b0 = *(const float4 *)&Bs[k][tx * 4], b1 = *(const float4 *)&Bs[k][64 + tx * 4];
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= Mtot) continue;
        int b = m / s.rows_per_batch, r = m - b * s.rows_per_batch;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n < s.N) epilogue_store(e, b, r, n, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass in registers.  d % 128 == 0, d <= 1280.
// ---------------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float *__restrict__ x, const float *__restrict__ g, const float *__restrict__ bta, int rows, int d,
                 OutT *__restrict__ out, float *__restrict__ out2) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // programmatic dependent launch (common.cuh launch_chain): x comes from the previous kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp >= rows) return;
    const int nch = d >> 7;  // float4 chunks per lane
    const float4 *xr = (const float4 *)(x + (size_t)warp * d);
    float4 v[10];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c)
        if (c < nch) {
            v[c] = xr[c * 32 + lane];
            sum += v[c].x + v[c].y + v[c].z + v[c].w;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c)
        if (c < nch) {
            float a = v[c].x - mean, b = v[c].y - mean, cc = v[c].z - mean, dd = v[c].w - mean;
            sq += a * a + b * b + cc * cc + dd * dd;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)d + 1e-5f);
    const float4 *g4 = (const float4 *)g, *b4 = (const float4 *)bta;
#pragma unroll
    for (int c = 0; c < 10; ++c)
        if (c < nch) {
            float4 gg = __ldg(g4 + c * 32 + lane), bb = __ldg(b4 + c * 32 + lane);
            float4 y;
            y.x = (v[c].x - mean) * rstd * gg.x + bb.x;
            y.y = (v[c].y - mean) * rstd * gg.y + bb.y;
            y.z = (v[c].z - mean) * rstd * gg.z + bb.z;
            y.w = (v[c].w - mean) * rstd * gg.w + bb.w;
            size_t idx = (size_t)warp * d + (size_t)(c * 32 + lane) * 4;
            if constexpr (sizeof(OutT) == 4) {
                *(float4 *)((float *)out + idx) = y;
            } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
                uint2 pk;
                pk.x = *(unsigned *)&lo;
                pk.y = *(unsigned *)&hi;
                *(uint2 *)((bf16 *)out + idx) = pk;
            }
            if (out2) *(float4 *)(out2 + idx) = y;
        }
}

// ---------------------------------------------------------------------------------------------------------
// SIMT attention (fp32 math).  grid (q tiles of 64, heads, B); 256 threads as 16x16, 4x4 outputs per thread.
// q and k arrive pre-scaled by head_dim^-0.25 each (fused in the QKV GEMM epilogue), as candle scales both.
// ---------------------------------------------------------------------------------------------------------
constexpr int AT_B = 64, AT_S = 65;

template <typename T>
__device__ __forceinline__ float ld_f(const T *p) {
    if constexpr (sizeof(T) == 4) return *p;
    else return __bfloat162float(*p);
}

template <typename T>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const T *__restrict__ qkv, T *__restrict__ out, int T_len, int d) {
    extern __shared__ float sm[];
    float *Qs = sm, *Ks = Qs + AT_B * AT_S, *Vs = Ks + AT_B * AT_S, *Ps = Vs + AT_B * AT_S;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * AT_B, h = blockIdx.y, b = blockIdx.z;
    const size_t row0 = (size_t)b * T_len;
    const int ld = 3 * d;
    const T *qb = qkv + h * HEAD_DIM, *kb = qkv + d + h * HEAD_DIM, *vb = qkv + 2 * d + h * HEAD_DIM;

    for (int i = tid; i < AT_B * HEAD_DIM; i += 256) {
        int r = i >> 6, c = i & 63;
        Qs[r * AT_S + c] = (q0 + r < T_len) ? ld_f(qb + (row0 + q0 + r) * ld + c) : 0.f;
    }
    float m_i[4], l_i[4], o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_i[i] = -INFINITY;
        l_i[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    }
    for (int k0 = 0; k0 < T_len; k0 += AT_B) {
        __syncthreads();
        for (int i = tid; i < AT_B * HEAD_DIM; i += 256) {
            int r = i >> 6, c = i & 63;
            bool ok = k0 + r < T_len;
            Ks[r * AT_S + c] = ok ? ld_f(kb + (row0 + k0 + r) * ld + c) : 0.f;
            Vs[r * AT_S + c] = ok ? ld_f(vb + (row0 + k0 + r) * ld + c) : 0.f;
        }
        __syncthreads();
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll 8
        for (int c = 0; c < HEAD_DIM; ++c) {
            float qv[4], kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = Qs[(ty + 16 * i) * AT_S + c];
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = Ks[(tx + 16 * j) * AT_S + c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sc[i][j] = fmaf(qv[i], kv[j], sc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx + 16 * j >= T_len) sc[i][j] = -INFINITY;
                mx = fmaxf(mx, sc[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float m_new = fmaxf(m_i[i], mx);
            float alpha = __expf(m_i[i] - m_new);  // m_i = -inf on the first tile -> 0
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float p = __expf(sc[i][j] - m_new);
                Ps[(ty + 16 * i) * AT_S + tx + 16 * j] = p;
                ps += p;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
            l_i[i] = l_i[i] * alpha + ps;
            m_i[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < AT_B; ++kk) {
            float pv[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = Ps[(ty + 16 * i) * AT_S + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = Vs[kk * AT_S + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(pv[i], vv[j], o[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int q = q0 + ty + 16 * i;
        if (q >= T_len) continue;
        float inv = 1.0f / l_i[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = o[i][j] * inv;
            size_t idx = (row0 + q) * (size_t)d + h * HEAD_DIM + tx + 16 * j;
            if constexpr (sizeof(T) == 4) out[idx] = v;
            else out[idx] = __float2bfloat16(v);
        }
    }
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ in, bf16 *__restrict__ out, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *(const float4 *)(in + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *(unsigned *)&lo;
        pk.y = *(unsigned *)&hi;
        *(uint2 *)(out + i) = pk;
    } else {
        for (; i < n; ++i) out[i] = __float2bfloat16(in[i]);
    }
}

}  // namespace

int simt_init(nb200_ctx *ctx) {
    const int smem = 4 * AT_B * AT_S * 4;
    CUDA_TRY(ctx, cudaFuncSetAttribute(attention_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(ctx, cudaFuncSetAttribute(attention_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return NB200_OK;
}

int launch_gemm_f32(nb200_ctx *ctx, const float *A, const float *W, const GemmShape &s, const Epilogue &e) {
    if (s.K % SG_BK != 0 || s.lda % 4 != 0 || s.a_bs % 4 != 0)
        return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "sgemm: K=%d lda=%lld a_bs=%lld not aligned", s.K, s.lda, s.a_bs);
    KernelScope ks(ctx, NB200_K_GEMM);
    ctx->prof_gemm_flops += 2.0 * s.batch * s.rows_per_batch * (double)s.N * s.K;
    dim3 grid(ceil_div(s.N, SG_BN), ceil_div(s.batch * s.rows_per_batch, SG_BM));
    sgemm_kernel<<<grid, 256, 0, ctx->stream>>>(A, W, s, e);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int launch_layernorm(nb200_ctx *ctx, const float *x, const float *g, const float *b, int rows, int d, void *out, int out_bf16,
                     float *out2_f32) {
    if (d % 128 != 0 || d > 1280) return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "layernorm: d=%d (need d %% 128 == 0, d <= 1280)", d);
    KernelScope ks(ctx, NB200_K_LAYERNORM);
    int blocks = ceil_div(rows, 8);
    if (out_bf16) CUDA_TRY(ctx, launch_chain(ctx, 4, layernorm_kernel<bf16>, dim3(blocks), dim3(256), 0, 1, x, g, b, rows, d, (bf16 *)out, out2_f32));
    else CUDA_TRY(ctx, launch_chain(ctx, 4, layernorm_kernel<float>, dim3(blocks), dim3(256), 0, 1, x, g, b, rows, d, (float *)out, out2_f32));
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int launch_attention_simt(nb200_ctx *ctx, const void *qkv, void *out, int B, int T, int n_heads, int is_bf16) {
    KernelScope ks(ctx, NB200_K_ATTN);
    const int smem = 4 * AT_B * AT_S * 4;
    const int d = n_heads * HEAD_DIM;
    dim3 grid(ceil_div(T, AT_B), n_heads, B);
    if (is_bf16) attention_simt_kernel<bf16><<<grid, 256, smem, ctx->stream>>>((const bf16 *)qkv, (bf16 *)out, T, d);
    else attention_simt_kernel<float><<<grid, 256, smem, ctx->stream>>>((const float *)qkv, (float *)out, T, d);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int launch_f32_to_bf16(nb200_ctx *ctx, const float *in, bf16 *out, size_t n) {
    KernelScope ks(ctx, NB200_K_MISC);
    size_t threads = (n + 3) / 4;
    f32_to_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(in, out, n);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}
