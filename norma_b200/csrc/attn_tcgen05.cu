// attn_tcgen05.cu — K6: fused non-causal attention for the Whisper encoder on the 5th-generation tensor cores.
// Replaces candle's materialised [B, h, 1500, 1500] f32 score tensor + softmax_last_dim + second batched matmul
// (reached from /root/reference/src/models/whisper/model.rs:455-464; SURVEY §2b "attention").
//
// One CTA = one 128-row query tile of one (window, head); two CTAs are resident per SM so one CTA's softmax
// overlaps the other's tensor-core work.  q and k arrive pre-scaled by head_dim^-0.25 each (QKV GEMM epilogue).
//   warp 0      TMA producer: Q tile once, then K_j / V_j tiles (128 keys x 64, SWIZZLE_128B) double-buffered
//   warp 1      MMA issuer (one thread):  S = Q . K_j^T  -> TMEM cols [0,128)   (UMMA 128x128x16, K-major A and B)
//                                         O += P_j . V_j -> TMEM cols [128,192) (UMMA 128x64x16, V is the MN-major B)
//                                         L += P_j . 1   -> TMEM cols [192,208) (UMMA 128x16x16 against a block of ones):
//                                         the softmax row sums come off the tensor core, not a serial FADD chain
//   warps 2-5   softmax: thread = query row (TMEM lane).  tcgen05.ld the S row, running max in fp32, then
//               ex2.approx.ftz.bf16x2: two exponentials per MUFU op whose result IS the bf16 P operand; P -> shared memory
//               in the UMMA K-major SWIZZLE_128B layout; when a row maximum grows the O and L accumulator rows are
//               rescaled in TMEM (tcgen05.ld / st); final O / l -> bf16 -> global.
// Measured dead ends (profiles/r1c_attention_notes.md): moving a fraction of the fp32 exponentials to an FMA-pipe
// polynomial made the kernel slower (it trades MUFU cycles for issue slots about 1:1).
// S(j+1) is issued as soon as the softmax warps have pulled S(j) into registers, so QK^T overlaps the exp phase.
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace {

constexpr int AT_BM = 128, AT_BN = 128;
constexpr int AT_THREADS = 192;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: a [128 x 64] bf16 tile
// shared memory map (base must be 1024-aligned)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;       // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;   // 2 stages
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;   // 2 sub-tiles of [128 x 64]
constexpr int OFF_ONES = OFF_P + 2 * TILE_BYTES;  // 768 B of bf16 1.0 (the B operand of the row-sum MMA reads 512 B)
constexpr int OFF_BAR = OFF_ONES + 768;
constexpr int AT_SMEM = OFF_BAR + 128;  // 2 x (AT_SMEM + 1 KB system reserve) must stay <= 228 KB: two CTAs per SM
static_assert(2 * (AT_SMEM + 1024) <= 233472, "two attention CTAs must fit one SM");
constexpr int AT_TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct SoftmaxCtx {
    uint32_t tS, tO, lane_off, s_full, s_empty, p_full, o_done;
    uint8_t *p_row;
    int rx, lane;
};

// One K/V tile of the online softmax for this thread's query row.  MASK is a template parameter so the tail masking
// (keys >= T exist only in the last tile) is not if-converted into per-element selects on every tile.
// two exponentials per MUFU op; the packed bf16 result is the P operand as it goes to shared memory
__device__ __forceinline__ uint32_t ex2_bf16x2(float x0, float x1) {
    uint32_t packed, y;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(packed) : "f"(x1), "f"(x0));  // low half <- x0
    asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(packed));
    return y;
}

// One K/V tile of the online softmax for this thread's query row.  MASK is a template parameter: keys >= T exist only
// in the last tile.
template <bool MASK, bool BF16EXP, bool ONES>
__device__ __forceinline__ void softmax_tile(const SoftmaxCtx &c, int j, int T, float &m, float &l) {
    ptx::mbar_wait(c.s_full, (uint32_t)(j & 1));
    ptx::tc_fence_after();
    uint32_t sv[128];
    ptx::tmem_ld_32x32b_x32(c.tS + c.lane_off + 0, sv);
    ptx::tmem_ld_32x32b_x32(c.tS + c.lane_off + 32, sv + 32);
    ptx::tmem_ld_32x32b_x32(c.tS + c.lane_off + 64, sv + 64);
    ptx::tmem_ld_32x32b_x32(c.tS + c.lane_off + 96, sv + 96);
    ptx::tmem_ld_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (c.lane == 0) ptx::mbar_arrive(c.s_empty);
    float mx = -INFINITY;
    if (MASK) {
        const int valid = T - j * AT_BN;  // >= 1
#pragma unroll
        for (int i = 0; i < 128; ++i) {
            float s = (i < valid) ? __uint_as_float(sv[i]) : -INFINITY;
            sv[i] = __float_as_uint(s);
            mx = fmaxf(mx, s);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
    }
    const float m_new = fmaxf(m, mx);
    const float mb = m_new * LOG2E;
    uint32_t pk[64];
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        const float x0 = fmaf(__uint_as_float(sv[2 * i]), LOG2E, -mb), x1 = fmaf(__uint_as_float(sv[2 * i + 1]), LOG2E, -mb);
        if (BF16EXP) {
            pk[i] = ex2_bf16x2(x0, x1);
            if (!ONES) {
                rs0 += __uint_as_float(pk[i] << 16);
                rs1 += __uint_as_float(pk[i] & 0xffff0000u);
            }
        } else {
            const float p0 = ex2(x0), p1 = ex2(x1);
            if (!ONES) { rs0 += p0; rs1 += p1; }
            __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
            pk[i] = *(uint32_t *)&t;
        }
    }
    if (!ONES) l = l * ex2((m - m_new) * LOG2E) + (rs0 + rs1);
    if (j > 0) {
        ptx::mbar_wait(c.o_done, (uint32_t)((j - 1) & 1));  // P.V(j-1) retired: P buffer free, O / L readable
        ptx::tc_fence_after();
    }
    // P row -> smem, K-major SWIZZLE_128B: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            uint4 v = make_uint4(pk[t * 32 + cc * 4], pk[t * 32 + cc * 4 + 1], pk[t * 32 + cc * 4 + 2], pk[t * 32 + cc * 4 + 3]);
            *(uint4 *)(c.p_row + t * TILE_BYTES + ((cc ^ c.rx) << 4)) = v;
        }
    if (j > 0 && __any_sync(0xffffffffu, m_new > m)) {
        // rescale this warp's 32 rows of O (64 columns) and L (16 columns; the chunk's other 16 are unused TMEM) by
        // alpha = 2^(m_old - m_new) (1 for rows whose maximum did not move)
        const float alpha = ex2((m - m_new) * LOG2E);
#pragma unroll
        for (int hh = 0; hh < (ONES ? 3 : 2); ++hh) {
            uint32_t ov[32];
            ptx::tmem_ld_32x32b_x32(c.tO + c.lane_off + hh * 32, ov);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            ptx::tmem_st_32x32b_x32(c.tO + c.lane_off + hh * 32, ov);
        }
        ptx::tmem_st_wait();
    }
    m = m_new;
    ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    ptx::tc_fence_before();
    __syncwarp();
    if (c.lane == 0) ptx::mbar_arrive(c.p_full);
}

template <bool BF16EXP, bool ONES>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16 *__restrict__ out, int T, int d) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = ptx::smem_u32(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * AT_BM;
    const int nkv = (T + AT_BN - 1) / AT_BN;

    const uint32_t sQ = sbase + OFF_Q, sK = sbase + OFF_K, sV = sbase + OFF_V, sP = sbase + OFF_P;
    const uint32_t bar = sbase + OFF_BAR;
    // K and V have separate 2-deep rings: the K stage is free as soon as S = Q.K^T has been computed, the V stage only after
    // P.V retires, so K(j+1) can be in flight two tiles ahead and S(j+1) never waits behind P.V(j-1) -> TMA -> K(j+1)
    const uint32_t q_full = bar, k_full0 = bar + 8, k_empty0 = bar + 24, v_full0 = bar + 40, v_empty0 = bar + 56, s_full = bar + 72,
                   s_empty = bar + 80, p_full = bar + 88, o_done = bar + 96, tmem_slot = bar + 104;
    volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(smem_raw + OFF_BAR + 104);

    if (threadIdx.x == 0 && (sbase & 1023u)) __trap();  // SWIZZLE_128B tiles need a 1024-byte aligned base
    ((uint32_t *)(smem_raw + OFF_ONES))[threadIdx.x] = 0x3F803F80u;  // 192 threads x 4 B = 768 B of bf16 1.0
    ptx::fence_proxy_async();
    if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmQKV);
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(k_full0 + 8 * s, 1);
            ptx::mbar_init(k_empty0 + 8 * s, 1);
            ptx::mbar_init(v_full0 + 8 * s, 1);
            ptx::mbar_init(v_empty0 + 8 * s, 1);
        }
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(s_empty, 4);
        ptx::mbar_init(p_full, 4);
        ptx::mbar_init(o_done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, AT_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tS = tmem_base, tO = tmem_base + 128, tL = tmem_base + 192;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_expect_tx(q_full, TILE_BYTES);
            ptx::tma_load_3d(sQ, &tmQKV, q_full, h * HEAD_DIM, q0, b);
            // K runs ahead of V: K(j) is requested as soon as S(j-2) has been computed
            for (int j = 0; j < nkv + 1; ++j) {
                if (j < nkv) {
                    const int s = j & 1;
                    if (j >= 2) ptx::mbar_wait(k_empty0 + 8 * s, (uint32_t)(((j >> 1) & 1) ^ 1));
                    ptx::mbar_expect_tx(k_full0 + 8 * s, TILE_BYTES);
                    ptx::tma_load_3d(sK + s * TILE_BYTES, &tmQKV, k_full0 + 8 * s, d + h * HEAD_DIM, j * AT_BN, b);
                }
                if (j >= 1) {
                    const int jv = j - 1, s = jv & 1;
                    if (jv >= 2) ptx::mbar_wait(v_empty0 + 8 * s, (uint32_t)(((jv >> 1) & 1) ^ 1));
                    ptx::mbar_expect_tx(v_full0 + 8 * s, TILE_BYTES);
                    ptx::tma_load_3d(sV + s * TILE_BYTES, &tmQKV, v_full0 + 8 * s, 2 * d + h * HEAD_DIM, jv * AT_BN, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(AT_BM, AT_BN, 0, 0);
            constexpr uint32_t idesc_o = ptx::make_idesc_bf16(AT_BM, HEAD_DIM, 0, 1);  // B (= V) is MN-major
            constexpr uint32_t idesc_l = ptx::make_idesc_bf16(AT_BM, 16, 0, 0);
            const uint64_t d_ones = ptx::make_nosw_desc(sbase + OFF_ONES, 128, 256);  // 16 (N) x 16 (K) block of ones
            auto issue_s = [&](int j) {
                const uint32_t kb = sK + (j & 1) * TILE_BYTES;
#pragma unroll
                for (int k = 0; k < HEAD_DIM / 16; ++k)
                    ptx::mma_bf16_ss(tS, ptx::make_sw128_desc(sQ + k * 32, 16, 1024), ptx::make_sw128_desc(kb + k * 32, 16, 1024), idesc_s, k != 0);
                ptx::mma_commit(k_empty0 + 8 * (j & 1));  // K stage reusable once S(j) is computed
                ptx::mma_commit(s_full);
            };
            ptx::mbar_wait(q_full, 0);
            ptx::mbar_wait(k_full0, 0);
            ptx::tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                if (j + 1 < nkv) {
                    ptx::mbar_wait(k_full0 + 8 * ((j + 1) & 1), (uint32_t)(((j + 1) >> 1) & 1));
                    ptx::mbar_wait(s_empty, (uint32_t)(j & 1));  // softmax has S(j) in registers
                    ptx::tc_fence_after();
                    issue_s(j + 1);
                }
                ptx::mbar_wait(v_full0 + 8 * (j & 1), (uint32_t)((j >> 1) & 1));
                ptx::mbar_wait(p_full, (uint32_t)(j & 1));  // P(j) in smem, O rescaled
                ptx::tc_fence_after();
                const uint32_t vb = sV + (j & 1) * TILE_BYTES;
#pragma unroll
                for (int ks = 0; ks < AT_BN / 16; ++ks) {
                    const uint64_t da = ptx::make_sw128_desc(sP + (ks >> 2) * TILE_BYTES + (ks & 3) * 32, 16, 1024);
                    const uint64_t db = ptx::make_sw128_desc(vb + ks * 2048, 1024, 1024);
                    ptx::mma_bf16_ss(tO, da, db, idesc_o, (j | ks) != 0);
                    if (ONES) ptx::mma_bf16_ss(tL, da, d_ones, idesc_l, (j | ks) != 0);  // row sums of the bf16 P actually used
                }
                ptx::mma_commit(v_empty0 + 8 * (j & 1));
                ptx::mma_commit(o_done);
            }
        }
    } else {
        // ================= softmax warps: thread <-> query row / TMEM lane =================
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        float m = -INFINITY, l = 0.f;
        uint8_t *p_row = smem_raw + OFF_P + r * 128;
        const int rx = r & 7;
        SoftmaxCtx sc{tS, tO, lane_off, s_full, s_empty, p_full, o_done, p_row, rx, lane};
        for (int j = 0; j < nkv - 1; ++j) softmax_tile<false, BF16EXP, ONES>(sc, j, T, m, l);
        softmax_tile<true, BF16EXP, ONES>(sc, nkv - 1, T, m, l);  // only the last K/V tile can hold keys >= T
        ptx::mbar_wait(o_done, (uint32_t)((nkv - 1) & 1));
        ptx::tc_fence_after();
        float inv = 1.0f / l;
        if (ONES) {
            uint32_t lv[32];
            ptx::tmem_ld_32x32b_x32(tL + lane_off, lv);
            ptx::tmem_ld_wait();
            inv = 1.0f / __uint_as_float(lv[0]);
        }
        const int q = q0 + r;
        bf16 *orow = out + ((size_t)b * T + q) * d + h * HEAD_DIM;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t ov[32];
            ptx::tmem_ld_32x32b_x32(tO + lane_off + hh * 32, ov);
            ptx::tmem_ld_wait();
            if (q < T) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 v;
                    __nv_bfloat162 a0 = __floats2bfloat162_rn(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
                    __nv_bfloat162 a1 = __floats2bfloat162_rn(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
                    __nv_bfloat162 a2 = __floats2bfloat162_rn(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
                    __nv_bfloat162 a3 = __floats2bfloat162_rn(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
                    v.x = *(uint32_t *)&a0; v.y = *(uint32_t *)&a1; v.z = *(uint32_t *)&a2; v.w = *(uint32_t *)&a3;
                    *(uint4 *)(orow + hh * 32 + i) = v;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, AT_TMEM_COLS);
    }
}

}  // namespace

int attn_tc_init(nb200_ctx *ctx) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(attn_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    CUDA_TRY(ctx, cudaFuncSetAttribute(attn_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    CUDA_TRY(ctx, cudaFuncSetAttribute(attn_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    CUDA_TRY(ctx, cudaFuncSetAttribute(attn_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    return NB200_OK;
}

// qkv: [B*T][3d] bf16 (q | k | v, heads of 64 inside each third); out: [B*T][d] bf16
int launch_attention_tc(nb200_ctx *ctx, const bf16 *qkv, bf16 *out, int B, int T, int n_heads) {
    const int d = n_heads * HEAD_DIM;
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
    uint32_t box[3] = {HEAD_DIM, 128, 1};
    NB_TRY(tmap_encode_bf16(ctx, &tm, qkv, 3, dims, str, box));
    KernelScope ks(ctx, NB200_K_ATTN);
    dim3 grid(ceil_div(T, AT_BM), n_heads, B);
    static int variant = -1;  // NB200_ATTN_VARIANT: bit0 = bf16x2 exp2, bit1 = row sums on the tensor core (A/B experiments)
    if (variant < 0) {
        const char *ev = getenv("NB200_ATTN_VARIANT");
        variant = ev ? atoi(ev) : 0;
    }
    switch (variant & 3) {
        case 0: attn_tc_kernel<false, false><<<grid, AT_THREADS, AT_SMEM, ctx->stream>>>(tm, out, T, d); break;
        case 1: attn_tc_kernel<true, false><<<grid, AT_THREADS, AT_SMEM, ctx->stream>>>(tm, out, T, d); break;
        case 2: attn_tc_kernel<false, true><<<grid, AT_THREADS, AT_SMEM, ctx->stream>>>(tm, out, T, d); break;
        default: attn_tc_kernel<true, true><<<grid, AT_THREADS, AT_SMEM, ctx->stream>>>(tm, out, T, d); break;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}
