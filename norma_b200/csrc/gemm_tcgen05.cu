// gemm_tcgen05.cu — K5: bf16 x bf16 -> f32 GEMM on the 5th-generation tensor cores, C = A . W^T with a fused
// epilogue (bias, q/k scale, tanh-GELU, f32 residual add, f32|bf16 store).  Used for the conv stem (as an
// im2col-free GEMM over overlapping rows of a time-major buffer), QKV / out / fc1 / fc2 projections and the
// decoder's cross-attention K/V build.  Replaces candle's cuBLAS-SGEMM + separate bias/GELU/add kernels reached
// from /root/reference/src/models/whisper/model.rs:455-464.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   TMA producer: A tile 128 x 64 and W tile BN x 64 (bf16, K-major, SWIZZLE_128B) per stage
//   warp 1   MMA issuer:   one elected thread, tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16,
//                          accumulators in TMEM (2 stages x BN columns: epilogue of tile i overlaps mainloop of i+1)
//   warp 2   TMEM alloc / dealloc
//   warps 4-7 epilogue:    tcgen05.ld 32x32b (thread = accumulator row), fused math, vectorised global stores
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int GEMM_THREADS = 256;

struct GemmTcParams {
    int m_tiles_per_batch, n_tiles, total_tiles, k_blocks;
    int rows_per_batch, N;
    int vec_ok;  // all epilogue pointers / strides allow 16-byte vector access
    Epilogue epi;
};

template <int BN>
struct GemmCfg {
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void epi_chunk(const Epilogue &e, const uint32_t *acc, int b, int r, int n0, int N, bool row_ok, bool vec_ok) {
    if (!row_ok) return;
    float v[32];
    const bool full = vec_ok && n0 + 32 <= N;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    if (e.bias) {
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 bb = __ldg((const float4 *)(e.bias + n0 + j));
                v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) v[j] += __ldg(e.bias + n0 + j);
        }
    }
    if (n0 < e.n_scale) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < e.n_scale) v[j] *= e.scale;
    }
    if (e.act) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
    }
    if (e.residual) {
        const float *rp = e.residual + (long long)b * e.res_bs + (long long)r * e.ldr + n0;
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 rr = *(const float4 *)(rp + j);
                v[j] += rr.x; v[j + 1] += rr.y; v[j + 2] += rr.z; v[j + 3] += rr.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) v[j] += rp[j];
        }
    }
    const long long o = (long long)b * e.out_bs + (long long)r * e.ldo + n0;
    if (e.out_bf16) {
        bf16 *op = (bf16 *)e.out + o;
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                uint4 pk;
                pk.x = *(unsigned *)&p0; pk.y = *(unsigned *)&p1; pk.z = *(unsigned *)&p2; pk.w = *(unsigned *)&p3;
                *(uint4 *)(op + j) = pk;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) op[j] = __float2bfloat16(v[j]);
        }
    } else {
        float *op = (float *)e.out + o;
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *(float4 *)(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) op[j] = v[j];
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcParams p) {
    using C = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base, sB = smem_base + C::STAGES * C::A_BYTES;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    // barrier layout (8 B each): full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem_ptr (4 B)
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
    volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n_idx = tile % p.n_tiles, mb = tile / p.n_tiles;
                const int b = mb / p.m_tiles_per_batch, mt = mb - b * p.m_tiles_per_batch;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    ptx::mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    ptx::tma_load_3d(sA + stage * C::A_BYTES, &tmA, full_bar(stage), kb * BK, mt * BM, b);
                    ptx::tma_load_2d(sB + stage * C::B_BYTES, &tmB, full_bar(stage), kb * BK, n_idx * BN);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_sw128_desc(sA + stage * C::A_BYTES, 16, 1024);
                    const uint64_t db = ptx::make_sw128_desc(sB + stage * C::B_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)  // +32 B per UMMA_K step inside the 128 B swizzle atom
                        ptx::mma_bf16_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    ptx::mma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                ptx::mma_commit(tfull_bar(as));  // accumulator complete
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int n_idx = tile % p.n_tiles, mb = tile / p.n_tiles;
            const int b = mb / p.m_tiles_per_batch, mt = mb - b * p.m_tiles_per_batch;
            const int r = mt * BM + q * 32 + lane;
            const bool row_ok = r < p.rows_per_batch;
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int n0 = n_idx * BN + c * 32;
                if (n0 >= p.N) break;  // warp-uniform
                uint32_t acc[32];
                ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), acc);
                ptx::tmem_ld_wait();
                epi_chunk(p.epi, acc, b, r, n0, p.N, row_ok, p.vec_ok != 0);
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(as));
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled g_encode = nullptr;

}  // namespace

int tmap_encode_bf16(nb200_ctx *ctx, CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                     const uint32_t *box) {
    if (!g_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return nb200_fail(ctx, NB200_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
        g_encode = (PFN_tmapEncodeTiled)fn;
    }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return nb200_fail(ctx, NB200_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu strides %llu,%llu box %u,%u", (int)r,
                          rank, (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
                          (unsigned long long)strides_bytes[0], rank > 2 ? (unsigned long long)strides_bytes[1] : 0ull, box[0], box[1]);
    return NB200_OK;
}

int gemm_tc_init(nb200_ctx *ctx) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::SMEM_BYTES));
    return NB200_OK;
}

int launch_gemm_bf16(nb200_ctx *ctx, const bf16 *A, const bf16 *W, const GemmShape &s, const Epilogue &e) {
    if (s.lda % 8 != 0 || s.a_bs % 8 != 0 || s.K % 8 != 0 || ((uintptr_t)A & 15) || ((uintptr_t)W & 15))
        return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "gemm_bf16: operands must be 16-byte aligned (K=%d lda=%lld a_bs=%lld)", s.K, s.lda,
                          s.a_bs);
    // BN = 256 unless N is small enough that 128-wide tiles give a better wave fit
    const int m_tiles = ceil_div(s.rows_per_batch, BM) * s.batch;
    const bool use128 = (s.N % 256 != 0 && s.N % 128 == 0 && s.N <= 1024) || s.N <= 128;
    const int BN = use128 ? 128 : 256;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[3] = {(uint64_t)s.K, (uint64_t)s.rows_per_batch, (uint64_t)s.batch};
        uint64_t str[2] = {(uint64_t)s.lda * 2, (uint64_t)(s.batch > 1 ? s.a_bs : (long long)s.lda * s.rows_per_batch) * 2};
        uint32_t box[3] = {BK, BM, 1};
        NB_TRY(tmap_encode_bf16(ctx, &tmA, A, 3, dims, str, box));
    }
    {
        uint64_t dims[2] = {(uint64_t)s.K, (uint64_t)s.N};
        uint64_t str[1] = {(uint64_t)s.K * 2};
        uint32_t box[2] = {BK, (uint32_t)BN};
        NB_TRY(tmap_encode_bf16(ctx, &tmB, W, 2, dims, str, box));
    }
    GemmTcParams p;
    p.m_tiles_per_batch = ceil_div(s.rows_per_batch, BM);
    p.n_tiles = ceil_div(s.N, BN);
    p.total_tiles = m_tiles * p.n_tiles;
    p.k_blocks = ceil_div(s.K, BK);
    p.rows_per_batch = s.rows_per_batch;
    p.N = s.N;
    p.epi = e;
    const int oa = e.out_bf16 ? 8 : 4;
    p.vec_ok = (e.ldo % oa == 0) && (e.out_bs % oa == 0) && (((uintptr_t)e.out) % 16 == 0) && (!e.bias || ((uintptr_t)e.bias) % 16 == 0) &&
               (!e.residual || (e.ldr % 4 == 0 && e.res_bs % 4 == 0 && ((uintptr_t)e.residual) % 16 == 0));
    const int grid = p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count;
    KernelScope ks(ctx, NB200_K_GEMM);
    ctx->prof_gemm_flops += 2.0 * s.batch * s.rows_per_batch * (double)s.N * s.K;
    if (BN == 256)
        gemm_tc_kernel<256><<<grid, GEMM_THREADS, GemmCfg<256>::SMEM_BYTES, ctx->stream>>>(tmA, tmB, p);
    else
        gemm_tc_kernel<128><<<grid, GEMM_THREADS, GemmCfg<128>::SMEM_BYTES, ctx->stream>>>(tmA, tmB, p);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}
