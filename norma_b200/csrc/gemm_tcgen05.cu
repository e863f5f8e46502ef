// gemm_tcgen05.cu — K5: bf16 x bf16 -> f32 GEMM on the 5th-generation tensor cores, C = A . W^T with a fused
// epilogue (bias, q/k scale, tanh-GELU, f32 residual add, f32|bf16 store).  Used for the conv stem (as an
// im2col-free GEMM over overlapping rows of a time-major buffer), QKV / out / fc1 / fc2 projections and the
// decoder's cross-attention K/V build.  Replaces candle's cuBLAS-SGEMM + separate bias/GELU/add kernels reached
// from /root/reference/src/models/whisper/model.rs:455-464.
//
// Structure (persistent, warp-specialised, one CTA per SM, 384 threads):
//   warp 0     TMA producer: A tile 128 x 64 and W tile (bf16, K-major, SWIZZLE_128B) per stage
//   warp 1     MMA issuer: one thread, tcgen05.mma kind::f16, accumulators in TMEM, 2 stages x BN columns so the
//              epilogue of tile i overlaps the mainloop of tile i+1
//   warp 2     TMEM alloc / dealloc
//   warps 4-11 epilogue (2 per scheduler, alternate column chunks): thread = accumulator row (TMEM lane).  tcgen05.ld -> fused math -> the warp's 32-row x 128-byte
//              chunk is written to a swizzled smem patch and leaves with ONE bulk tensor store (TMA); the f32
//              residual chunk arrives the same way (TMA load, first chunk issued before the accumulator is ready,
//              later chunks one ahead).  Thread-per-row global stores were measured to cap at ~1.2 TB/s of 16-byte
//              L2 write requests and made every K = 1280 GEMM epilogue-bound (profiles/).
// CTA2 = true is the cta_group::2 variant: a CTA pair (cluster 2x1) owns a 256 x BN tile; each CTA loads its own
// 128 A rows and HALF of the W tile, the leader issues UMMA 256 x BN x 16 over both CTAs' shared memory (operand
// traffic per FLOP / 1.5 against the 128 x 256 single-CTA tile), accumulator rows [128r, 128r+128) live in CTA r.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

namespace {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + 32 * EPI_WARPS;
constexpr int STG_BYTES = 4096;  // one staging patch: 32 rows x 128 B

struct GemmTcParams {
    int m_tiles_per_batch, n_tiles, total_tiles, k_blocks;
    int rows_per_batch, N;
    int vec_ok;   // direct epilogue: all pointers / strides allow 16-byte vector access
    int use_tma;  // TMA-store epilogue (needs 16-byte aligned rows); 0 = direct thread-per-row stores
    int res_b0;   // residual has no batch dimension (positional table): always batch coordinate 0
    int debug;    // microbenchmark switches: 1 = epilogue touches no global memory, 2 = no TMA / MMA mainloop
    long long *dbg_clk;  // NB200_GEMM_DEBUG & 256 (gemm_wide_kernel): per-CTA clock64 stamps of the phases, [grid][8]
    Epilogue epi;
};

// NP = staging patches per epilogue warp.  Two are enough without a residual (one being filled, one being stored).  With an f32
// residual every chunk's residual is a TMA load the chunk has to wait for: NP patches keep NP loads in flight per warp, the first NP
// issued while the mainloop of the tile still runs — with two, three of a 256-wide tile's four chunks each exposed most of an HBM round
// trip (the out-proj GEMM ran at 3.3 TB/s of memory traffic, bound by neither pipe).  The third patch costs the pipeline one stage.
template <int BN, bool CTA2, int NP>
struct GemmCfg {
    static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;  // W rows this CTA loads per stage
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGING = EPI_WARPS * NP * STG_BYTES;
    static constexpr int BAR_BYTES = 384;
    static constexpr int LN_BYTES = BM * 4;        // 1 / std of the tile's rows (fused LayerNorm consumer)
    static constexpr int BIAS_BYTES = 2 * BN * 4;  // the tile's bias, per accumulator stage
    static constexpr int SMEM_LIMIT = 232448;      // the 227 KB a CTA may opt in to
    static constexpr int FIT = (SMEM_LIMIT - STAGING - BAR_BYTES - LN_BYTES - BIAS_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = FIT > 5 ? 5 : FIT;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int TILE_M = CTA2 ? 256 : 128;
    // no alignment slack: the dynamic shared memory of a kernel without static shared memory starts 1024-byte aligned (checked, trap)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING + BAR_BYTES + LN_BYTES + BIAS_BYTES;
    static_assert(STAGES >= 3, "pipeline too shallow");
    static_assert((2 * STAGES + 9 + EPI_WARPS * NP) * 8 <= BAR_BYTES, "barrier block");
    static_assert(SMEM_BYTES <= SMEM_LIMIT, "exceeds the 227 KB of shared memory a CTA may opt in to");
};

__device__ __forceinline__ float epi_math(float acc, float bias, float cs, int act) {
    float v = (acc + bias) * cs;
    if (act) v = gelu_tanh_fast(v);
    return v;
}

// ---- fallback epilogue: thread-per-row direct global access (arbitrary alignment) -------------------------------
__device__ __forceinline__ void epi_chunk_direct(const Epilogue &e, const uint32_t *acc, int b, int r, int n0, int N, bool row_ok, bool vec_ok) {
    if (!row_ok) return;
    const bool full = vec_ok && n0 + 32 <= N;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float bias = (e.bias && n0 + j < N) ? __ldg(e.bias + n0 + j) : 0.f;
        v[j] = epi_math(__uint_as_float(acc[j]), bias, (n0 + j < e.n_scale) ? e.scale : 1.0f, e.act);
    }
    if (e.residual) {
        const float *rp = e.residual + (long long)b * e.res_bs + (long long)r * e.ldr + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < N) v[j] += rp[j];
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

// bias for 4 consecutive columns (guarded at the N edge)
__device__ __forceinline__ float4 bias4(const Epilogue &e, int n, int N, bool full) {
    if (!e.bias) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (full) return __ldg((const float4 *)(e.bias + n));
    float4 b;
    b.x = n + 0 < N ? __ldg(e.bias + n + 0) : 0.f;
    b.y = n + 1 < N ? __ldg(e.bias + n + 1) : 0.f;
    b.z = n + 2 < N ? __ldg(e.bias + n + 2) : 0.f;
    b.w = n + 3 < N ? __ldg(e.bias + n + 3) : 0.f;
    return b;
}

// ---- fused LayerNorm, producer side ---------------------------------------------------------------------------------
// (mean, M2) of 32 values held in registers (two passes: exact to rounding whatever the mean), merged into the running partial
// of this thread's row with Chan's update
__device__ __forceinline__ void ln_stats_chunk(const uint32_t *v, float &cnt, float &mean, float &m2) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        s0 += __uint_as_float(v[i]); s1 += __uint_as_float(v[i + 1]); s2 += __uint_as_float(v[i + 2]); s3 += __uint_as_float(v[i + 3]);
    }
    const float cm = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float d0 = __uint_as_float(v[i]) - cm, d1 = __uint_as_float(v[i + 1]) - cm, d2 = __uint_as_float(v[i + 2]) - cm, d3 = __uint_as_float(v[i + 3]) - cm;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
    }
    const float cq = (q0 + q1) + (q2 + q3);
    const float tot = cnt + 32.0f, delta = cm - mean, w = __fdividef(32.0f, tot);
    mean = fmaf(delta, w, mean);
    m2 += cq + delta * delta * cnt * w;
    cnt = tot;
}

// bf16 copy of this thread's 32 values (one row of a 32-column chunk): 64 contiguous bytes as two 256-bit stores — whole 32-byte
// sectors, no shuffles (16-byte thread-per-row stores leave every sector half-written per instruction and were measured to cap at
// ~1.2 TB/s)
__device__ __forceinline__ void ln_store_bf16_row(const uint32_t *v, bf16 *dst) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        pk[i] = *(uint32_t *)&t;
    }
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]),
                 "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                 : "memory");
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16), "r"(pk[8]), "r"(pk[9]), "r"(pk[10]), "r"(pk[11]),
                 "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                 : "memory");
}

// the same copy for a 32-row x 32-column chunk with 16-byte stores: packed to 4 x 16 bytes per row, transposed inside each lane quad
// with two butterfly stages so that in store k the four lanes of a quad write the 64 contiguous bytes of row (quad base + k)
__device__ __forceinline__ void ln_store_bf16_chunk(const uint32_t *v, bf16 *dst_row0 /* row (quad base) of this chunk, column n0 */, long long ld, int lane,
                                                    int rows_left /* valid rows from the quad base on */) {
    uint4 a[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[8 * p + 0]), __uint_as_float(v[8 * p + 1]));
        __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[8 * p + 2]), __uint_as_float(v[8 * p + 3]));
        __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[8 * p + 4]), __uint_as_float(v[8 * p + 5]));
        __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[8 * p + 6]), __uint_as_float(v[8 * p + 7]));
        a[p].x = *(uint32_t *)&t0; a[p].y = *(uint32_t *)&t1; a[p].z = *(uint32_t *)&t2; a[p].w = *(uint32_t *)&t3;
    }
    const bool b0 = lane & 1, b1 = lane & 2;
    auto xchg = [&](uint4 &lo, uint4 &hi, bool bit, int mask) {  // lanes with the bit clear keep lo and trade hi, the others the reverse
        uint4 snd = bit ? lo : hi, rcv;
        rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, mask); rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, mask);
        rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, mask); rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, mask);
        if (bit) lo = rcv; else hi = rcv;
    };
    xchg(a[0], a[1], b0, 1); xchg(a[2], a[3], b0, 1);
    xchg(a[0], a[2], b1, 2); xchg(a[1], a[3], b1, 2);
    // now a[k] = piece (lane & 3) of the row of quad lane k
    bf16 *d = dst_row0 + 8 * (lane & 3);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < rows_left) *(uint4 *)(d + (long long)k * ld) = a[k];
}

template <int BN, bool CTA2, int NP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const GemmTcParams p) {
    using C = GemmCfg<BN, CTA2, NP>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_raw_u = ptx::smem_u32(smem_raw);
    if (threadIdx.x == 0 && (smem_raw_u & 1023u)) __trap();  // the swizzled tiles need it; there is no slack to align by hand
    const uint32_t smem_base = smem_raw_u;
    const uint32_t sA = smem_base, sB = smem_base + C::STAGES * C::A_BYTES;
    const uint32_t stg_base = smem_base + C::STAGES * C::STAGE_BYTES;  // 1024-aligned
    const uint32_t bars = stg_base + C::STAGING;
    // barrier layout (8 B each): full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr | res_full[EPI_WARPS][2]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
    auto res_bar = [&](int w, int buf) { return bars + 8u * (2 * C::STAGES + 5 + NP * w + buf); };
    // fused LayerNorm (consumer): warp 3 merges the rows' partial statistics of every tile and publishes 1 / std per accumulator stage
    auto lnfull_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 5 + NP * EPI_WARPS + s); };
    auto lnempty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + 7 + NP * EPI_WARPS + s); };
    float *ln_smem = (float *)(smem_raw + (bars + C::BAR_BYTES - smem_raw_u));  // [BM]
    // the tile's bias, staged once per tile by the epilogue warps while the mainloop runs: bias loads straight from global memory inside the
    // chunk loop were one L2 round trip per 8 columns, in program order (4 800 clk of a 64-column chunk's 5 000, measured with clock64)
    float *bias_smem = ln_smem + BM;  // [2][BN]
    volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(smem_raw + (tmem_slot - smem_raw_u));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CTA2 ? ptx::cluster_ctarank() : 0u;
    const int first_tile = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // t-th tile of this CTA (pair) -> (row block mb, n tile); false when the CTA has run out of tiles.  All three roles walk the same sequence.
    auto tile_at = [&](int t, int &mb, int &n_idx) -> bool {
        const int tile = first_tile + t * tile_step;
        n_idx = tile % p.n_tiles;
        mb = tile / p.n_tiles;
        return tile < p.total_tiles;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        if (p.use_tma) {
            ptx::prefetch_tmap(&tmOut);
            if (p.epi.residual) ptx::prefetch_tmap(&tmRes);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), CTA2 ? 2 * EPI_WARPS : EPI_WARPS);  // CTA2: both CTAs' warps arrive on the LEADER
        }
        for (int w = 0; w < EPI_WARPS; ++w) {
            for (int k = 0; k < NP; ++k) ptx::mbar_init(res_bar(w, k), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(lnfull_bar(s), 1);
            ptx::mbar_init(lnempty_bar(s), EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if (CTA2) {
            ptx::tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
            ptx::tmem_relinquish_2sm();
        } else {
            ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync();  // both CTAs' barriers are initialised before any remote signal
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int k_blocks = (p.debug & 2) ? 0 : p.k_blocks;
    ptx::griddep_wait();    // everything above overlapped the previous kernel's tail; from here on its results are needed
    ptx::griddep_launch();  // AFTER the wait: the next kernel's CTAs may take this SM as soon as this CTA leaves it, but never start while the
                            // kernel before this one still runs (a trigger in front of the wait let three kernels overlap and broke bit-identity)

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int mb, n_idx;
            for (int t = 0; tile_at(t, mb, n_idx); ++t) {
                const int b = mb / p.m_tiles_per_batch, mt = mb - b * p.m_tiles_per_batch;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    if (CTA2) {
                        if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);  // bytes of both CTAs
                        const uint32_t lead_full = ptx::mapa(full_bar(stage), 0);
                        ptx::tma_load_3d_2sm(sA + stage * C::A_BYTES, &tmA, lead_full, kb * BK, mt * 256 + (int)rank * BM, b);
                        ptx::tma_load_2d_2sm(sB + stage * C::B_BYTES, &tmB, lead_full, kb * BK, n_idx * BN + (int)rank * (BN / 2));
                    } else {
                        ptx::mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                        ptx::tma_load_3d(sA + stage * C::A_BYTES, &tmA, full_bar(stage), kb * BK, mt * BM, b);
                        ptx::tma_load_2d(sB + stage * C::B_BYTES, &tmB, full_bar(stage), kb * BK, n_idx * BN);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only when paired) =================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(C::TILE_M, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            int mb_, n_idx_;
            for (int t = 0; tile_at(t, mb_, n_idx_); ++t) {
                ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::make_sw128_desc(sA + stage * C::A_BYTES, 16, 1024);
                    const uint64_t db = ptx::make_sw128_desc(sB + stage * C::B_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {  // +32 B per UMMA_K step inside the 128 B swizzle atom
                        if (CTA2) ptx::mma_bf16_ss_2sm(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        else ptx::mma_bf16_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    if (CTA2) ptx::mma_commit_2sm_mc(empty_bar(stage), 3);  // frees the stage in both CTAs
                    else ptx::mma_commit(empty_bar(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                if (CTA2) ptx::mma_commit_2sm_mc(tfull_bar(as), 3);  // accumulator complete
                else ptx::mma_commit(tfull_bar(as));
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp == 3) {
        // ================= fused LayerNorm statistics (consumer side only) =================
        // For every tile of this CTA, in order: merge the producer's partial (mean, M2) slots of the tile's 128 rows (Chan's update, equal
        // column counts) and publish 1 / std in shared memory for the epilogue warps.  Lane l owns rows l, l + 32, l + 64, l + 96: every
        // load is one coalesced 256-byte request and four rows' chains run side by side.  A whole mainloop is available per tile, so none of
        // this is on the epilogue's critical path (merging the slots in the epilogue warps cost the QKV / fc1 GEMMs 15 %).
        const Epilogue &e = p.epi;
        if (e.stats_in != nullptr && p.use_tma && e.out_bf16) {
            int as = 0;
            uint32_t aphase = 0;
            const float cols = (float)e.stats_cols;
            int mb, n_idx;
            for (int t = 0; tile_at(t, mb, n_idx); ++t) {
                const int b = mb / p.m_tiles_per_batch, mt = mb - b * p.m_tiles_per_batch;
                const int r0 = mt * C::TILE_M + (int)rank * BM;
                ptx::mbar_wait(lnempty_bar(0), (uint32_t)(t & 1) ^ 1u);  // the epilogue warps hold the previous tile's values in registers
                const float2 *sp = e.stats_in + (long long)b * p.rows_per_batch + r0 + lane;
                float cnt = 0.f, mean[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
                bool ok[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ok[k] = r0 + lane + 32 * k < p.rows_per_batch;
#pragma unroll 2
                for (int i = 0; i < e.stats_slots; ++i) {
                    float2 pm[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) pm[k] = ok[k] ? __ldcg(sp + (long long)i * e.stats_ld + 32 * k) : make_float2(0.f, 1.f);
                    const float tot = cnt + cols, w = __fdividef(cols, tot);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float delta = pm[k].x - mean[k];
                        mean[k] = fmaf(delta, w, mean[k]);
                        m2[k] += pm[k].y + delta * delta * cnt * w;
                    }
                    cnt = tot;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) ln_smem[lane + 32 * k] = rsqrtf(m2[k] / cnt + LN_EPS);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(lnfull_bar(0));  // release: the stores above are ordered before the arrival
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: 8 warps = 2 per scheduler; warps (4+q) and (8+q) share TMEM lane quarter q and take
        // alternate column chunks =================
        const int ew = warp - 4;
        const int q = warp & 3;      // TMEM lane quarter this warp may access
        const int grp = ew >> 2;     // 0 / 1: even / odd chunks
        const Epilogue &e = p.epi;
        const bool has_res = e.residual != nullptr;
        const bool no_mem = (p.debug & 1) != 0;
        const uint32_t stg = stg_base + (uint32_t)(ew * NP * STG_BYTES);
        uint8_t *stg_ptr = smem_raw + (stg - smem_raw_u);
        const int sw = lane & 7;
        uint32_t rphase = 0;  // bit b: parity of res_bar(ew, b)
        int as = 0;
        uint32_t aphase = 0;
        // timing build of a launch (NB200_GEMM_DEBUG & 256): clock64 deltas of this warp's phases, summed over its tiles
        float bias_next = 0.f;
        static_assert(BN <= 32 * EPI_WARPS, "one bias value per epilogue thread");
        const bool dbg_on = p.dbg_clk != nullptr;
        long long d_wait = 0, d_res = 0, d_math = 0, d_st = 0, d_ln = 0, d_tiles = 0, d_chunks = 0, d_t0 = dbg_on ? clock64() : 0;
        int mb, n_idx;
        for (int t = 0; tile_at(t, mb, n_idx); ++t) {
            const int b = mb / p.m_tiles_per_batch, mt = mb - b * p.m_tiles_per_batch;
            const int r0 = mt * C::TILE_M + (int)rank * BM + q * 32;  // this warp's first row
            const int nt0 = n_idx * BN;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            const int rb = p.res_b0 ? 0 : b;
            // the tile's bias -> shared memory (all 8 epilogue warps, one L2 round trip in the shadow of the mainloop).  The named barrier also
            // orders the reuse of the buffer two tiles on: every warp has finished tile t - 1, so nobody still reads tile t - 2's values.
            float *bs = bias_smem + as * BN;
            if (t == 0) {
                for (int k = ew * 32 + lane; k < BN; k += 32 * EPI_WARPS) bs[k] = (e.bias && nt0 + k < p.N) ? __ldg(e.bias + nt0 + k) : 0.f;
            } else {
                if (ew * 32 + lane < BN) bs[ew * 32 + lane] = bias_next;  // requested a tile ago (BN <= 256: one value per epilogue thread)
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            {   // the NEXT tile's bias: the load is in flight while this tile's chunks are processed
                int mb2, n2;
                bias_next = 0.f;
                if (tile_at(t + 1, mb2, n2)) {
                    const int k = n2 * BN + ew * 32 + lane;
                    if (e.bias && ew * 32 + lane < BN && k < p.N) bias_next = __ldg(e.bias + k);
                }
            }
            if (p.use_tma && !no_mem) {
                if (e.out_bf16) {
                    // ---------- bf16 out: 64-column chunks (128 B rows), no residual ----------
                    // fused LayerNorm (consumer): A is the bf16 copy of the residual stream and W carries gamma AND the centring
                    // (W . diag(gamma) . (I - 11^T / K)), so all that is left of the LayerNorm is this row's 1 / std
                    float ln_rstd = 1.0f;
                    if (e.stats_in != nullptr) {  // published by the statistics warp (warp 3) for this accumulator stage
                        ptx::mbar_wait(lnfull_bar(0), (uint32_t)(t & 1));
                        ln_rstd = ln_smem[q * 32 + lane];
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(lnempty_bar(0));
                    }
                    ptx::mbar_wait(tfull_bar(as), aphase);
                    ptx::tc_fence_after();
#pragma unroll 1
                    for (int ci = 0; ci < BN / 128; ++ci) {
                        const int c = 2 * ci + grp;
                        const int n0 = nt0 + c * 64;
                        if (n0 >= p.N) continue;  // warp-uniform
                        const uint32_t sb = stg + (uint32_t)((ci & 1) * STG_BYTES);
                        uint8_t *sbp = stg_ptr + (ci & 1) * STG_BYTES + lane * 128;
                        uint32_t acc[64];
                        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64), acc);
                        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + 32), acc + 32);
                        if (lane == 0) ptx::bulk_wait_read<1>();  // the store that used this patch two chunks ago is done reading
                        __syncwarp();
                        const bool full = n0 + 64 <= p.N;
                        const bool mixed = n0 < e.n_scale && n0 + 64 > e.n_scale;
                        const float cs = (n0 + 64 <= e.n_scale) ? e.scale : 1.0f;
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {  // 16-byte group j = columns [8j, 8j+8)
                            const float4 b0 = *(const float4 *)(bs + c * 64 + 8 * j), b1 = *(const float4 *)(bs + c * 64 + 8 * j + 4);
                            float v[8];
                            v[0] = fmaf(__uint_as_float(acc[8 * j + 0]), ln_rstd, b0.x) * cs; v[1] = fmaf(__uint_as_float(acc[8 * j + 1]), ln_rstd, b0.y) * cs;
                            v[2] = fmaf(__uint_as_float(acc[8 * j + 2]), ln_rstd, b0.z) * cs; v[3] = fmaf(__uint_as_float(acc[8 * j + 3]), ln_rstd, b0.w) * cs;
                            v[4] = fmaf(__uint_as_float(acc[8 * j + 4]), ln_rstd, b1.x) * cs; v[5] = fmaf(__uint_as_float(acc[8 * j + 5]), ln_rstd, b1.y) * cs;
                            v[6] = fmaf(__uint_as_float(acc[8 * j + 6]), ln_rstd, b1.z) * cs; v[7] = fmaf(__uint_as_float(acc[8 * j + 7]), ln_rstd, b1.w) * cs;
                            if (mixed) {  // a chunk straddling the scaled-column boundary (not hit by the Whisper shapes)
#pragma unroll
                                for (int t = 0; t < 8; ++t)
                                    if (n0 + 8 * j + t < e.n_scale) v[t] *= e.scale;
                            }
                            if (e.act) {
#pragma unroll
                                for (int t = 0; t < 8; ++t) v[t] = gelu_tanh_fast(v[t]);
                            }
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                            uint4 pk;
                            pk.x = *(unsigned *)&p0; pk.y = *(unsigned *)&p1; pk.z = *(unsigned *)&p2; pk.w = *(unsigned *)&p3;
                            *(uint4 *)(sbp + ((j ^ sw) << 4)) = pk;
                        }
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_3d(&tmOut, sb, n0, r0, b);
                            ptx::bulk_commit();
                        }
                    }
                } else {
                    // ---------- f32 out: 32-column chunks (128 B rows), optional f32 residual via TMA ----------
                    constexpr int NCI = BN / 64;  // chunks per warp
                    const bool ln_out = e.stats_out != nullptr;  // fused LayerNorm (producer): bf16 copy + per-row partial statistics
                    float ln_cnt = 0.f, ln_mean = 0.f, ln_m2 = 0.f;  // (count, mean, M2) of this thread's row over this warp's columns of the tile
                    if (has_res && lane == 0) {  // the residual of the first NP chunks is fetched while the mainloop of this tile still runs
                        ptx::bulk_wait_read<0>();  // the previous tile's stores no longer read the patches
#pragma unroll
                        for (int k = 0; k < (NP < NCI ? NP : NCI); ++k) {
                            const int nk = nt0 + (2 * k + grp) * 32;
                            if (nk < p.N) {
                                ptx::mbar_expect_tx(res_bar(ew, k), STG_BYTES);
                                ptx::tma_load_3d(stg + (uint32_t)(k * STG_BYTES), &tmRes, res_bar(ew, k), nk, r0, rb);
                            }
                        }
                    }
                    const long long k0 = dbg_on ? clock64() : 0;
                    ptx::mbar_wait(tfull_bar(as), aphase);
                    ptx::tc_fence_after();
                    if (dbg_on) { d_wait += clock64() - k0; ++d_tiles; }
#pragma unroll 1
                    for (int ci = 0; ci < NCI; ++ci) {
                        const int c = 2 * ci + grp;
                        const int n0 = nt0 + c * 32;
                        if (n0 >= p.N) continue;  // warp-uniform
                        const long long k1 = dbg_on ? clock64() : 0;
                        const int buf = has_res ? ci % NP : (ci & 1);
                        const uint32_t sb = stg + (uint32_t)(buf * STG_BYTES);
                        uint8_t *sbp = stg_ptr + buf * STG_BYTES + lane * 128;
                        uint32_t acc[32];
                        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), acc);
                        if (has_res) {
                            // the patch the previous chunk was stored from takes the residual of chunk ci - 1 + NP
                            if (lane == 0 && ci >= 1 && ci - 1 + NP < NCI && n0 + (NP - 1) * 64 < p.N) {
                                const int pb = (ci - 1) % NP;
                                ptx::bulk_wait_read<0>();  // that store no longer reads the patch
                                ptx::mbar_expect_tx(res_bar(ew, pb), STG_BYTES);
                                ptx::tma_load_3d(stg + (uint32_t)(pb * STG_BYTES), &tmRes, res_bar(ew, pb), n0 + (NP - 1) * 64, r0, rb);
                            }
                            ptx::mbar_wait(res_bar(ew, buf), (rphase >> buf) & 1u);
                            rphase ^= 1u << buf;
                        } else {
                            if (lane == 0) ptx::bulk_wait_read<1>();
                            __syncwarp();
                        }
                        const bool full = n0 + 32 <= p.N;
                        const bool mixed = n0 < e.n_scale && n0 + 32 > e.n_scale;
                        const float cs = (n0 + 32 <= e.n_scale) ? e.scale : 1.0f;
                        ptx::tmem_ld_wait();
                        const long long k2 = dbg_on ? clock64() : 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {  // 16-byte group j = columns [4j, 4j+4)
                            const float4 bb = *(const float4 *)(bs + c * 32 + 4 * j);
                            float v[4];
                            v[0] = (__uint_as_float(acc[4 * j + 0]) + bb.x) * cs; v[1] = (__uint_as_float(acc[4 * j + 1]) + bb.y) * cs;
                            v[2] = (__uint_as_float(acc[4 * j + 2]) + bb.z) * cs; v[3] = (__uint_as_float(acc[4 * j + 3]) + bb.w) * cs;
                            if (mixed) {
#pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    if (n0 + 4 * j + t < e.n_scale) v[t] *= e.scale;
                            }
                            if (e.act) {
#pragma unroll
                                for (int t = 0; t < 4; ++t) v[t] = gelu_tanh_fast(v[t]);
                            }
                            float4 *sp = (float4 *)(sbp + ((j ^ sw) << 4));
                            if (has_res) {
                                const float4 rr = *sp;
                                v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
                            }
                            *sp = make_float4(v[0], v[1], v[2], v[3]);
                            if (ln_out) {  // keep the finished values: statistics and the bf16 copy follow the store
                                acc[4 * j + 0] = __float_as_uint(v[0]); acc[4 * j + 1] = __float_as_uint(v[1]);
                                acc[4 * j + 2] = __float_as_uint(v[2]); acc[4 * j + 3] = __float_as_uint(v[3]);
                            }
                        }
                        const long long k3 = dbg_on ? clock64() : 0;
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_3d(&tmOut, sb, n0, r0, b);
                            ptx::bulk_commit();
                        }
                        const long long k4 = dbg_on ? clock64() : 0;
                        if (ln_out) {
                            if (!(p.debug & 16)) ln_stats_chunk(acc, ln_cnt, ln_mean, ln_m2);
                            if (p.debug & 8) {
                            } else if (p.debug & 32) {
                                if (r0 + lane < p.rows_per_batch) ln_store_bf16_row(acc, e.xb_out + (long long)b * e.out_bs + (long long)(r0 + lane) * e.ldo + n0);
                            } else {
                                const int qr = r0 + (lane & ~3);  // first row of this lane's quad
                                ln_store_bf16_chunk(acc, e.xb_out + (long long)b * e.out_bs + (long long)qr * e.ldo + n0, e.ldo, lane, p.rows_per_batch - qr);
                            }
                        }
                        if (dbg_on) { const long long k5 = clock64(); d_res += k2 - k1; d_math += k3 - k2; d_st += k4 - k3; d_ln += k5 - k4; ++d_chunks; }
                    }
                    // one partial per (row, n tile, epilogue warp group); the consuming GEMM's statistics warp merges a row's slots
                    if (ln_out && r0 + lane < p.rows_per_batch)
                        e.stats_out[(long long)(n_idx * 2 + grp) * e.stats_ld + (long long)b * p.rows_per_batch + r0 + lane] = make_float2(ln_mean, ln_m2);
                }
            } else {
                // ---------- direct thread-per-row epilogue (unaligned shapes; also the no-memory microbenchmark mode) ----------
                const int r = r0 + lane;
                const bool row_ok = r < p.rows_per_batch && !no_mem;
                ptx::mbar_wait(tfull_bar(as), aphase);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c = grp; c < BN / 32; c += 2) {
                    const int n0 = nt0 + c * 32;
                    if (n0 >= p.N) break;
                    uint32_t acc[32];
                    ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), acc);
                    ptx::tmem_ld_wait();
                    epi_chunk_direct(e, acc, b, r, n0, p.N, row_ok, p.vec_ok != 0);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) ptx::mbar_arrive_cluster(ptx::mapa(tempty_bar(as), 0));
                else ptx::mbar_arrive(tempty_bar(as));
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
        if (lane == 0) ptx::bulk_wait_all();  // every bulk store of this thread has been written
        if (dbg_on && ew == 0 && lane == 0) {
            long long *o = p.dbg_clk + (size_t)blockIdx.x * 16;
            o[0] = d_wait; o[1] = d_res; o[2] = d_math; o[3] = d_st; o[4] = d_ln; o[5] = d_tiles; o[6] = d_chunks; o[7] = clock64() - d_t0;
        }
    }
    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync();  // nobody leaves (or frees TMEM) while the peer can still signal into this CTA
    else __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}


// =====================================================================================================================
// gemm_wide_kernel — the bf16-out GEMMs (QKV, fc1) for ONE window's worth of rows (M <= 1536: BASELINE config 2 as written, streaming).
// With 256 x 256 tiles, M = 1500 is 6 x 15 (QKV) or 6 x 20 (fc1) tiles on 74 CTA pairs: two rounds, the second nearly empty, and
// every round pays a pipeline fill and an epilogue (22.7 / 26.0 us against 10.4 / 13.9 us of tensor time).  Here every CTA pair computes
// exactly ONE 256 x BN tile, BN = NSUB x UN chosen so that 6 x ceil(N / BN) <= 74 pairs: one round.
//   * BN up to 448 columns: NSUB UMMAs (256 x UN x 16, cta_group::2) per k-step into one TMEM accumulator (no double buffer: one tile);
//   * one tile per CTA means the epilogue never overlaps a mainloop, so its staging patches ALIAS the first pipeline stages and the ring
//     gets the whole shared memory (6 stages of 36 KB at BN = 320, 5 of 44 KB at BN = 448 — the loop is bound by bytes in flight);
//   * same fused epilogue as gemm_tc_kernel's bf16 branch (bias, q/k scale, GELU, folded-LayerNorm 1/std from the statistics warp).
// =====================================================================================================================
template <int UN, int NSUB>
struct WideCfg {
    static constexpr int BN = UN * NSUB;
    static constexpr int A_BYTES = BM * BK * 2;            // this CTA's 128 rows
    static constexpr int B_SUB_BYTES = (UN / 2) * BK * 2;  // this CTA's half of one UMMA's W rows
    static constexpr int B_BYTES = NSUB * B_SUB_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int LN_BYTES = BM * 4;
    static constexpr int SMEM_LIMIT = 232448;
    static constexpr int BIAS_BYTES = BN * 4;
    static constexpr int STAGES = (SMEM_LIMIT - 1024 - BAR_BYTES - LN_BYTES - BIAS_BYTES) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + LN_BYTES + BIAS_BYTES;
    static constexpr int TMEM_COLS = 512;
    static_assert(UN % 16 == 0 && UN <= 256 && BN <= 512, "UMMA shape");
    static_assert(B_SUB_BYTES % 1024 == 0, "sub-tiles start on a swizzle atom");
    static_assert(STAGES * A_BYTES >= EPI_WARPS * 2 * STG_BYTES, "the epilogue patches alias the A stages");
    static_assert((2 * STAGES + 3) * 8 + 4 <= BAR_BYTES, "barrier block");
};

template <int UN, int NSUB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const GemmTcParams p) {
    using C = WideCfg<UN, NSUB>;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_raw_u = ptx::smem_u32(smem_raw);
    const uint32_t smem_base = (smem_raw_u + 1023u) & ~1023u;
    const uint32_t sA = smem_base, sB = smem_base + C::STAGES * C::A_BYTES;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    const uint32_t tfull_bar = bars + 8u * (2 * C::STAGES), lnfull_bar = bars + 8u * (2 * C::STAGES + 1);
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 2);
    volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(smem_raw + (tmem_slot - smem_raw_u));
    float *ln_smem = (float *)(smem_raw + (bars + C::BAR_BYTES - smem_raw_u));
    float *bias_smem = ln_smem + BM;  // the tile's bias, staged while the mainloop runs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int tile = (int)(blockIdx.x >> 1);
    const int n_idx = tile % p.n_tiles, mt = tile / p.n_tiles;  // batch == 1 on this path
    const Epilogue &e = p.epi;
    long long *dbg = p.dbg_clk ? p.dbg_clk + (size_t)blockIdx.x * 16 : nullptr;  // phase stamps (timing build of the launch, not the product path)
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        ptx::prefetch_tmap(&tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::mbar_init(lnfull_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    ptx::griddep_wait();
    ptx::griddep_launch();
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();  // set-up done

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                if (rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                const uint32_t lead_full = ptx::mapa(full_bar(stage), 0);
                ptx::tma_load_3d_2sm(sA + stage * C::A_BYTES, &tmA, lead_full, kb * BK, mt * 256 + (int)rank * BM, 0);
#pragma unroll
                for (int j = 0; j < NSUB; ++j)  // W rows of UMMA j: this CTA's half (rows beyond N arrive as zeros)
                    ptx::tma_load_2d_2sm(sB + stage * C::B_BYTES + j * C::B_SUB_BYTES, &tmB, lead_full, kb * BK, n_idx * BN + j * UN + (int)rank * (UN / 2));
                if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(256, UN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                if (dbg && kb == 0) dbg[2] = clock64();  // first operands have landed
                if (dbg && kb == p.k_blocks - 1) dbg[3] = clock64();  // last operands have landed
                const uint64_t da = ptx::make_sw128_desc(sA + stage * C::A_BYTES, 16, 1024);
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const uint64_t db = ptx::make_sw128_desc(sB + stage * C::B_BYTES + j * C::B_SUB_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        ptx::mma_bf16_ss_2sm(tmem_base + (uint32_t)(j * UN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                }
                ptx::mma_commit_2sm_mc(empty_bar(stage), 3);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            }
            ptx::mma_commit_2sm_mc(tfull_bar, 3);
        }
    } else if (warp == 3) {
        // folded LayerNorm: 1 / std of this CTA's 128 rows from the producer's partial slots (as in gemm_tc_kernel)
        if (e.stats_in != nullptr) {
            const int r0 = mt * 256 + (int)rank * BM;
            const float cols = (float)e.stats_cols;
            const float2 *sp = e.stats_in + r0 + lane;
            float cnt = 0.f, mean[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) ok[k] = r0 + lane + 32 * k < p.rows_per_batch;
#pragma unroll 2
            for (int i = 0; i < e.stats_slots; ++i) {
                float2 pm[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) pm[k] = ok[k] ? __ldcg(sp + (long long)i * e.stats_ld + 32 * k) : make_float2(0.f, 1.f);
                const float tot = cnt + cols, w = __fdividef(cols, tot);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float delta = pm[k].x - mean[k];
                    mean[k] = fmaf(delta, w, mean[k]);
                    m2[k] += pm[k].y + delta * delta * cnt * w;
                }
                cnt = tot;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) ln_smem[lane + 32 * k] = rsqrtf(m2[k] / cnt + LN_EPS);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(lnfull_bar);
        }
    } else if (warp >= 4) {
        const int ew = warp - 4, q = warp & 3, grp = ew >> 2;
        const uint32_t stg = sA + (uint32_t)(ew * 2 * STG_BYTES);  // aliases the A stages: nothing reads them once the accumulator is complete
        uint8_t *stg_ptr = smem_raw + (stg - smem_raw_u);
        const int sw = lane & 7;
        const int r0 = mt * 256 + (int)rank * BM + q * 32;
        const int nt0 = n_idx * BN;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int k = ew * 32 + lane; k < BN; k += 32 * EPI_WARPS) bias_smem[k] = (e.bias && nt0 + k < p.N) ? __ldg(e.bias + nt0 + k) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        float ln_rstd = 1.0f;
        if (e.stats_in != nullptr) {
            ptx::mbar_wait(lnfull_bar, 0);
            ln_rstd = ln_smem[q * 32 + lane];
        }
        ptx::mbar_wait(tfull_bar, 0);
        ptx::tc_fence_after();
        if (dbg && warp == 4 && lane == 0) dbg[4] = clock64();  // accumulator complete
        int it = 0;
        long long e_ld = 0, e_math = 0, e_st = 0, e_n = 0;
#pragma unroll 1
        for (int c = grp; c < BN / 64; c += 2, ++it) {
            const int n0 = nt0 + c * 64;
            if (n0 >= p.N) break;  // warp-uniform
            const uint32_t sb = stg + (uint32_t)((it & 1) * STG_BYTES);
            uint8_t *sbp = stg_ptr + (it & 1) * STG_BYTES + lane * 128;
            uint32_t acc[64];
            const long long c0 = dbg ? clock64() : 0;
            ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64), acc);
            ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 64 + 32), acc + 32);
            if (lane == 0) ptx::bulk_wait_read<1>();
            __syncwarp();
            const bool full = n0 + 64 <= p.N;
            const bool mixed = n0 < e.n_scale && n0 + 64 > e.n_scale;
            const float cs = (n0 + 64 <= e.n_scale) ? e.scale : 1.0f;
            ptx::tmem_ld_wait();
            const long long c1 = dbg ? clock64() : 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b0 = *(const float4 *)(bias_smem + c * 64 + 8 * j), b1 = *(const float4 *)(bias_smem + c * 64 + 8 * j + 4);
                float v[8];
                v[0] = fmaf(__uint_as_float(acc[8 * j + 0]), ln_rstd, b0.x) * cs; v[1] = fmaf(__uint_as_float(acc[8 * j + 1]), ln_rstd, b0.y) * cs;
                v[2] = fmaf(__uint_as_float(acc[8 * j + 2]), ln_rstd, b0.z) * cs; v[3] = fmaf(__uint_as_float(acc[8 * j + 3]), ln_rstd, b0.w) * cs;
                v[4] = fmaf(__uint_as_float(acc[8 * j + 4]), ln_rstd, b1.x) * cs; v[5] = fmaf(__uint_as_float(acc[8 * j + 5]), ln_rstd, b1.y) * cs;
                v[6] = fmaf(__uint_as_float(acc[8 * j + 6]), ln_rstd, b1.z) * cs; v[7] = fmaf(__uint_as_float(acc[8 * j + 7]), ln_rstd, b1.w) * cs;
                if (mixed) {
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (n0 + 8 * j + t < e.n_scale) v[t] *= e.scale;
                }
                if (e.act) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[t] = gelu_tanh_fast(v[t]);
                }
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                uint4 pk;
                pk.x = *(unsigned *)&p0; pk.y = *(unsigned *)&p1; pk.z = *(unsigned *)&p2; pk.w = *(unsigned *)&p3;
                *(uint4 *)(sbp + ((j ^ sw) << 4)) = pk;
            }
            const long long c2 = dbg ? clock64() : 0;
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&tmOut, sb, n0, r0, 0);  // rows beyond M and columns beyond N are clipped by the tensor map
                ptx::bulk_commit();
            }
            if (dbg) { const long long c3 = clock64(); e_ld += c1 - c0; e_math += c2 - c1; e_st += c3 - c2; ++e_n; }
        }
        if (dbg && warp == 4 && lane == 0) { dbg[5] = clock64(); dbg[8] = e_ld; dbg[9] = e_math; dbg[10] = e_st; dbg[11] = e_n; }  // last store issued
        if (lane == 0) ptx::bulk_wait_all();
        if (dbg && warp == 4 && lane == 0) dbg[6] = clock64();  // stores written
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    }
    if (dbg && threadIdx.x == 0) dbg[7] = clock64();
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// the driver entry point is resolved once per process, thread-safely (contexts are created and driven from different threads)
PFN_tmapEncodeTiled tmap_encode_fn() {
    static std::once_flag once;
    static PFN_tmapEncodeTiled fn_ = nullptr;
    std::call_once(once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn) fn_ = (PFN_tmapEncodeTiled)fn;
    });
    return fn_;
}

int tmap_encode(nb200_ctx *ctx, CUtensorMap *out, CUtensorMapDataType dt, const void *base, int rank, const uint64_t *dims,
                const uint64_t *strides_bytes, const uint32_t *box) {
    const PFN_tmapEncodeTiled g_encode = tmap_encode_fn();
    if (!g_encode) return nb200_fail(ctx, NB200_CUDA_ERROR, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = g_encode(out, dt, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return nb200_fail(ctx, NB200_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu strides %llu,%llu box %u,%u", (int)r,
                          rank, (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
                          (unsigned long long)strides_bytes[0], rank > 2 ? (unsigned long long)strides_bytes[1] : 0ull, box[0], box[1]);
    return NB200_OK;
}

}  // namespace

int tmap_encode_bf16(nb200_ctx *ctx, CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                     const uint32_t *box) {
    return tmap_encode(ctx, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}

int gemm_tc_init(nb200_ctx *ctx) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<256, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, false, 2>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<128, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, false, 2>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<256, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, true, 2>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<128, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, true, 2>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_tc_kernel<256, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, true, 3>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_wide_kernel<160, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideCfg<160, 2>::SMEM_BYTES));
    CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_wide_kernel<224, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WideCfg<224, 2>::SMEM_BYTES));
    return NB200_OK;
}

int launch_gemm_bf16(nb200_ctx *ctx, const bf16 *A, const bf16 *W, const GemmShape &s, const Epilogue &e) {
    if (s.lda % 8 != 0 || s.a_bs % 8 != 0 || s.K % 8 != 0 || ((uintptr_t)A & 15) || ((uintptr_t)W & 15))
        return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "gemm_bf16: operands must be 16-byte aligned (K=%d lda=%lld a_bs=%lld)", s.K, s.lda,
                          s.a_bs);
    const int mode = ctx->opt.gemm_mode, epi = ctx->opt.gemm_epi_tma;  // NB200_GEMM / NB200_EPI, read at nb200_create (A/B comparison)
    // Tile configuration: the CTA-pair 256 x 256 tile wins everywhere except for ONE window's worth of rows (streaming, BASELINE configs 4 / 5 at
    // B = 1) on the N = d_model GEMMs, where 6 x 5 pair tiles occupy 30 of 74 pairs; 128 x 128 single-CTA tiles (120 of 148 CTAs) are then
    // faster (scripts/gpu_gemm_fit.py, M = 1500: out-proj 18.6 -> 14.4 us, fc2 36.5 -> 28.5 us; from M = 3000 on the pair tile is as fast or faster)
    bool pair = (mode == 2 && s.N >= 256) || (mode == 4 && s.N >= 128);
    bool use128 = mode == 4 ? pair : (!pair && (mode == 3 || (s.N % 256 != 0 && s.N % 128 == 0 && s.N <= 1024) || s.N <= 128));
    if (mode == 2 && (long long)s.rows_per_batch * s.batch <= 2048 && s.N <= 1536 && s.N % 128 == 0 && !ctx->opt.gemm_nofit) {
        pair = false;
        use128 = true;
    }
    if (e.stats_out && s.N % 256 != 0 && s.N % 128 == 0 && !use128) {  // fused LayerNorm statistics are per whole n tile
        pair = false;
        use128 = true;
    }
    const int BN = use128 ? 128 : 256;
    const int tile_m = pair ? 256 : BM;
    GemmTcParams p;
    p.m_tiles_per_batch = ceil_div(s.rows_per_batch, tile_m);
    p.n_tiles = ceil_div(s.N, BN);
    p.total_tiles = p.m_tiles_per_batch * s.batch * p.n_tiles;
    p.k_blocks = ceil_div(s.K, BK);
    p.rows_per_batch = s.rows_per_batch;
    p.N = s.N;
    p.epi = e;
    p.debug = ctx->opt.gemm_debug;
    const int oes = e.out_bf16 ? 2 : 4;
    const int oa = 16 / oes;
    p.vec_ok = (e.ldo % oa == 0) && (e.out_bs % oa == 0) && (((uintptr_t)e.out) % 16 == 0) && (!e.bias || ((uintptr_t)e.bias) % 16 == 0) &&
               (!e.residual || (e.ldr % 4 == 0 && e.res_bs % 4 == 0 && ((uintptr_t)e.residual) % 16 == 0));
    // the TMA epilogue handles (bf16 out, no residual) and (f32 out, optional f32 residual) on 16-byte aligned rows
    p.use_tma = epi == 1 && p.vec_ok && !(e.out_bf16 && e.residual) && e.ldo > 0;
    p.res_b0 = (e.residual && (s.batch == 1 || e.res_bs == 0)) ? 1 : 0;
    if (e.stats_out) {  // fused LayerNorm, producer: whole tiles only, one slot per (n tile, epilogue warp group)
        if (!p.use_tma || e.out_bf16 || !e.xb_out || s.N % BN != 0 || p.n_tiles * 2 > LN_MAX_SLOTS || e.ldo % 16 != 0 || ((uintptr_t)e.xb_out & 31))
            return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "gemm_bf16: fused LayerNorm statistics need f32 TMA output and N %% %d == 0 (N=%d)", BN, s.N);
        ctx->ln_slots = p.n_tiles * 2;
        ctx->ln_slot_cols = BN / 2;
    }
    if (e.stats_in && (!p.use_tma || !e.out_bf16 || e.stats_slots <= 0 || e.stats_slots > LN_MAX_SLOTS))
        return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "gemm_bf16: fused LayerNorm input needs bf16 TMA output (N=%d)", s.N);

    // One window's worth of rows on a wide GEMM (QKV, fc1 at d_model 1280): one 256 x BN tile per CTA pair, a single round (gemm_wide_kernel)
    if (ctx->opt.gemm_wide && p.use_tma && e.out_bf16 && !e.residual && s.batch == 1 && s.rows_per_batch <= 1536 && s.N > 1536 && s.N % 64 == 0 &&
        mode == 2) {
        const int pairs = ctx->sm_count / 2, mt = ceil_div(s.rows_per_batch, 256);
        int bn = 0;
        for (int cand : {320, 448}) {
            const int tiles = mt * ceil_div(s.N, cand);
            if (tiles <= pairs && tiles * 10 >= pairs * 8) { bn = cand; break; }
        }
        if (bn) {
            p.m_tiles_per_batch = mt;
            p.n_tiles = ceil_div(s.N, bn);
            p.total_tiles = mt * p.n_tiles;
            CUtensorMap tmA, tmB, tmOut;
            {
                uint64_t dims[3] = {(uint64_t)s.K, (uint64_t)s.rows_per_batch, 1};
                uint64_t str[2] = {(uint64_t)s.lda * 2, (uint64_t)s.lda * s.rows_per_batch * 2};
                uint32_t box[3] = {BK, BM, 1};
                NB_TRY(tmap_encode_bf16(ctx, &tmA, A, 3, dims, str, box));
            }
            {
                uint64_t dims[2] = {(uint64_t)s.K, (uint64_t)s.N};
                uint64_t str[1] = {(uint64_t)s.K * 2};
                uint32_t box[2] = {BK, (uint32_t)(bn / 4)};  // UN / 2 rows: this CTA's half of one UMMA's W rows
                NB_TRY(tmap_encode_bf16(ctx, &tmB, W, 2, dims, str, box));
            }
            {
                uint64_t dims[3] = {(uint64_t)s.N, (uint64_t)s.rows_per_batch, 1};
                uint64_t str[2] = {(uint64_t)e.ldo * 2, (uint64_t)e.ldo * s.rows_per_batch * 2};
                uint32_t box[3] = {64, 32, 1};
                NB_TRY(tmap_encode_bf16(ctx, &tmOut, e.out, 3, dims, str, box));
            }
            KernelScope ks(ctx, NB200_K_GEMM, (long long)s.N * 100000 + s.K);
            ctx->prof_gemm_flops += 2.0 * s.rows_per_batch * (double)s.N * s.K;
            p.dbg_clk = nullptr;
            static thread_local long long *dbg_buf = nullptr;  // timing build of a launch only (NB200_GEMM_DEBUG & 256): never set on the product path
            if (p.debug & 256) {
                if (!dbg_buf) cudaMalloc(&dbg_buf, 2 * 74 * 16 * sizeof(long long));
                p.dbg_clk = dbg_buf;
            }
            if (bn == 320) CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_wide_kernel<160, 2>, dim3(2 * p.total_tiles), dim3(GEMM_THREADS), WideCfg<160, 2>::SMEM_BYTES, 2, tmA, tmB, tmOut, p));
            else CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_wide_kernel<224, 2>, dim3(2 * p.total_tiles), dim3(GEMM_THREADS), WideCfg<224, 2>::SMEM_BYTES, 2, tmA, tmB, tmOut, p));
            CUDA_TRY(ctx, cudaGetLastError());
            if (p.debug & 256) {  // where a CTA's cycles go, averaged over the grid (stderr)
                std::vector<long long> h((size_t)2 * p.total_tiles * 16);
                cudaStreamSynchronize(ctx->stream);
                cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                double d[7] = {0, 0, 0, 0, 0, 0, 0}, ee[4] = {0, 0, 0, 0};
                int nl = 0;
                for (int c = 0; c < 2 * p.total_tiles; c += 2, ++nl) {  // leader CTAs (they own the MMA stamps)
                    const long long *t = &h[(size_t)c * 16];
                    d[0] += t[1] - t[0]; d[1] += t[2] - t[1]; d[2] += t[3] - t[2]; d[3] += t[4] - t[3]; d[4] += t[5] - t[4]; d[5] += t[6] - t[5]; d[6] += t[7] - t[0];
                    ee[0] += t[8]; ee[1] += t[9]; ee[2] += t[10]; ee[3] += t[11];
                }
                fprintf(stderr, "   epilogue warp 4, clk per 64-column chunk: tcgen05.ld + wait %.0f | bias + math + pack + st.shared %.0f | fence + store issue %.0f (%.1f chunks)\n",
                        ee[0] / ee[3], ee[1] / ee[3], ee[2] / ee[3], ee[3] / nl);
                fprintf(stderr, "[gemm_wide N=%d K=%d] clk per CTA: set-up %.0f | first operands %.0f | mainloop (first -> last operands) %.0f | last operands -> accumulator "
                                "%.0f | epilogue %.0f | store drain %.0f || whole CTA %.0f\n", s.N, s.K, d[0] / nl, d[1] / nl, d[2] / nl, d[3] / nl, d[4] / nl, d[5] / nl, d[6] / nl);
            }
            return NB200_OK;
        }
    }
    p.dbg_clk = nullptr;

    CUtensorMap tmA, tmB, tmOut, tmRes;
    {
        uint64_t dims[3] = {(uint64_t)s.K, (uint64_t)s.rows_per_batch, (uint64_t)s.batch};
        uint64_t str[2] = {(uint64_t)s.lda * 2, (uint64_t)(s.batch > 1 ? s.a_bs : (long long)s.lda * s.rows_per_batch) * 2};
        uint32_t box[3] = {BK, BM, 1};
        NB_TRY(tmap_encode_bf16(ctx, &tmA, A, 3, dims, str, box));
    }
    {
        uint64_t dims[2] = {(uint64_t)s.K, (uint64_t)s.N};
        uint64_t str[1] = {(uint64_t)s.K * 2};
        uint32_t box[2] = {BK, (uint32_t)(pair ? BN / 2 : BN)};  // a CTA of a pair loads half of the W tile
        NB_TRY(tmap_encode_bf16(ctx, &tmB, W, 2, dims, str, box));
    }
    tmOut = tmA;
    tmRes = tmA;
    if (p.use_tma) {
        {
            uint64_t dims[3] = {(uint64_t)s.N, (uint64_t)s.rows_per_batch, (uint64_t)s.batch};
            uint64_t str[2] = {(uint64_t)e.ldo * oes, (uint64_t)(s.batch > 1 ? e.out_bs : e.ldo * (long long)s.rows_per_batch) * oes};
            uint32_t box[3] = {(uint32_t)(128 / oes), 32, 1};
            NB_TRY(tmap_encode(ctx, &tmOut, e.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, e.out, 3, dims, str, box));
        }
        if (e.residual) {
            const bool b0 = p.res_b0 != 0;
            uint64_t dims[3] = {(uint64_t)s.N, (uint64_t)s.rows_per_batch, (uint64_t)(b0 ? 1 : s.batch)};
            uint64_t str[2] = {(uint64_t)e.ldr * 4, (uint64_t)(b0 ? e.ldr * (long long)s.rows_per_batch : e.res_bs) * 4};
            uint32_t box[3] = {32, 32, 1};
            NB_TRY(tmap_encode(ctx, &tmRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, e.residual, 3, dims, str, box));
        }
    }
    KernelScope ks(ctx, NB200_K_GEMM, (long long)s.N * 100000 + s.K);
    ctx->prof_gemm_flops += 2.0 * s.batch * s.rows_per_batch * (double)s.N * s.K;
    static thread_local long long *dbg_main = nullptr;  // timing build of a launch only (NB200_GEMM_DEBUG & 256)
    const bool dbg_main_on = (p.debug & 256) && p.use_tma && !e.out_bf16;
    if (dbg_main_on) {
        if (!dbg_main) cudaMalloc(&dbg_main, 148 * 16 * sizeof(long long));
        cudaMemsetAsync(dbg_main, 0, 148 * 16 * sizeof(long long), ctx->stream);
        p.dbg_clk = dbg_main;
    }
    if (pair) {
        const int max_cl = ctx->sm_count / 2;
        const int clusters = p.total_tiles < max_cl ? p.total_tiles : max_cl;
        // (more than two staging patches per epilogue warp — deeper residual prefetch at the price of pipeline stages — were measured and
        // dropped: out-proj 147.6 / 144.1 / 152.9 us and fc2 358 / 394 / 460 us with 2 / 3 / 4 patches, profiles/r2_notes.md)
        // f32 residual with a short K (out-proj): the epilogue is the critical path and every chunk waits for its residual tile; a third staging
        // patch per warp keeps one more residual load in flight, paid for with a pipeline stage (NB200_GEMM_NP3=0 disables)
        const bool np3 = ctx->opt.gemm_np3 && e.residual && !e.out_bf16 && p.use_tma && s.K <= 2048 && BN == 256;
        if (BN == 128) CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_tc_kernel<128, true, 2>, dim3(2 * clusters), dim3(GEMM_THREADS), GemmCfg<128, true, 2>::SMEM_BYTES, 2, tmA, tmB, tmOut, tmRes, p));
        else if (np3) CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_tc_kernel<256, true, 3>, dim3(2 * clusters), dim3(GEMM_THREADS), GemmCfg<256, true, 3>::SMEM_BYTES, 2, tmA, tmB, tmOut, tmRes, p));
        else CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_tc_kernel<256, true, 2>, dim3(2 * clusters), dim3(GEMM_THREADS), GemmCfg<256, true, 2>::SMEM_BYTES, 2, tmA, tmB, tmOut, tmRes, p));
    } else {
        const int grid = p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count;
        if (BN == 256) CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_tc_kernel<256, false, 2>, dim3(grid), dim3(GEMM_THREADS), GemmCfg<256, false, 2>::SMEM_BYTES, 1, tmA, tmB, tmOut, tmRes, p));
        else CUDA_TRY(ctx, launch_chain(ctx, 1, gemm_tc_kernel<128, false, 2>, dim3(grid), dim3(GEMM_THREADS), GemmCfg<128, false, 2>::SMEM_BYTES, 1, tmA, tmB, tmOut, tmRes, p));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if (dbg_main_on) {  // where an epilogue warp's cycles go (warp 4 of every CTA, averaged; stderr)
        std::vector<long long> h(148 * 16);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h.data(), dbg_main, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 148; ++c)
            for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)c * 16 + i];
        if (a[5] > 0 && a[6] > 0)
            fprintf(stderr, "[gemm_tc f32 epilogue N=%d K=%d ln=%d] warp 4, clk: accumulator wait %.0f per tile | per 32-column chunk: tcgen05.ld + residual wait %.0f | math + "
                            "st.shared %.0f | fence + store issue %.0f | statistics + bf16 copy %.0f | %.1f chunks per tile, whole kernel %.0f clk for %.1f tiles per CTA\n",
                    s.N, s.K, e.stats_out != nullptr, a[0] / a[5], a[1] / a[6], a[2] / a[6], a[3] / a[6], a[4] / a[6], a[6] / a[5], a[7] / 148.0, a[5] / 148.0);
    }
    return NB200_OK;
}
