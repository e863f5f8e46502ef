// attn_tcgen05.cu — K6: fused non-causal attention for the Whisper encoder on the 5th-generation tensor cores.
// Replaces candle's materialised [B, h, 1500, 1500] f32 score tensor + softmax_last_dim + second batched matmul
// (reached from /root/reference/src/models/whisper/model.rs:455-464; SURVEY §2b "attention").
//
// Persistent kernel: 2 CTAs per SM stay resident and walk the (window, head, 128-row query tile) items round-robin.  q and k
// arrive pre-scaled by head_dim^-0.25 each (QKV GEMM epilogue).  Per CTA:
//   warp 0      TMA producer: Q tile of the item (double-buffered across items), then K_j / V_j tiles (64 keys x 64, SWIZZLE_128B)
//               through a 4-stage ring
//   warp 1      MMA issuer (one thread):  S_j = Q . K_j^T -> one of three 64-column score buffers in TMEM (UMMA 128x64x16, SS)
//                                         O += P_j . V_j  -> TMEM columns [192, 256)  (UMMA 128x64x16, A = P_j read from TMEM,
//                                                            V is the MN-major B operand)
//   warps 2-5   softmax: thread = query row (TMEM lane).  tcgen05.ld the S row, online softmax in fp32 (ex2.approx), and the bf16
//               probabilities go straight back into the TMEM columns the scores came from (tcgen05.st): P never touches shared
//               memory.  The running maximum only moves when a row outgrows it by 2^8 (lazy rescale), so the O accumulator is
//               rarely touched; final O / l -> bf16 -> global.
// The softmax loop does not wait for the current tile's P.V: the scores of the next two tiles are already in the other S buffers, and the
// tensor core computes the first scores of the next item while the softmax warps write this item's output.  It does wait, just before it
// stores P(j), for P(j-1).V to have retired (see softmax_tile: a correctness requirement found in round 2, normally satisfied long before).
//
// How it got here, with the measurements (profiles/r1d_attention_notes.md): the bound for head_dim 64 on B200 is the XU pipe
// (16 ex2 / clk / SM: 1024 clk per 128 x 128 tile, measured by scripts/probes/mufu_probe.cu), not the tensor core (512 clk); the
// first kernel (one P buffer in shared memory, one CTA per item) sat at ~1840 clk per tile because (a) ptxas sinks the exponentials
// of tile j+1 below the wait for P(j).V(j) whatever the source order, (b) "some row's maximum grew" is true for most tiles when a
// warp owns 32 rows, so the O rescale and its wait ran almost every tile, and (c) every CTA paid ~5 us of launch / TMEM alloc /
// pipeline fill / write-out.  (a) is removed by aliasing P onto S with three S buffers, (b) by the lazy rescale, (c) by persistence.
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

#include <algorithm>

namespace {

constexpr int AT_BM = 128;
constexpr int AT_THREADS = 192;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: a [128 x 64] bf16 tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_LOG2 = 8.0f;  // lazy-rescale threshold in log2 units

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

#ifdef NB200_ATTN_TIMING
// Empty asm that consumes 16 registers: everything they depend on is computed before the next asm volatile as far as nvcc is
// concerned (ptxas may still sink it).  Only the timing build uses it, to keep the clock64 brackets honest.
__device__ __forceinline__ void pin16(const uint32_t *r) {
    asm volatile("" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
#endif

#ifdef NB200_ATTN_TIMING
#define AT_CLK(x) x = clock64()
#else
#define AT_CLK(x)
#endif
template <bool MASK>
__device__ __forceinline__ void softmax_tile(uint32_t tSj, uint32_t tO, uint32_t s_full, uint32_t s_par, uint32_t p_full, uint32_t o_done, uint32_t o_par_prev,
                                              int lane, bool first, bool have_prev, int valid, float &m, float &l, long long *tm = nullptr) {
#ifdef NB200_ATTN_TIMING
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#endif
    AT_CLK(c0);
    ptx::mbar_wait(s_full, s_par);
    ptx::tc_fence_after();
    AT_CLK(c1);
    uint32_t sv[64];
    ptx::tmem_ld_32x32b_x32(tSj, sv);
    ptx::tmem_ld_32x32b_x32(tSj + 32, sv + 32);
    ptx::tmem_ld_wait();
    AT_CLK(c2);
    if (MASK) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
            if (i >= valid) sv[i] = __float_as_uint(-INFINITY);
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
    }
    const float m_cand = fmaxf(fmaxf(m, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
    // lazy rescale: the reference point of the exponentials only moves when some row of this warp outgrew it by more than 2^8
    // (RESCALE_LOG2); until then p = 2^(s - m_stale) <= 256, exact enough in bf16 / fp32, and l, O stay consistent with m_stale
    const bool moved = __any_sync(0xffffffffu, (m_cand - m) * LOG2E > RESCALE_LOG2);
    const float m_new = moved ? m_cand : m;
    const float mb = m_new * LOG2E;
    // scale / subtract and the two row-sum chains as packed f32x2 operations (fma.rn.f32x2 / add.rn.f32x2: one issue slot per PAIR, the
    // same roundings as the scalar forms): the XU pipe is the bound, but the softmax stream alone reaches only 78 % of it and every issue
    // slot taken out of the loop helps the two warps of a scheduler interleave (scripts/probes/mufu_probe2.cu: V0 1 310 -> V3 1 195 clk per tile)
    const uint64_t c2 = pack2(LOG2E, LOG2E), nmb2 = pack2(-mb, -mb);
    uint64_t rs2 = pack2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float a, b;
        unpack2(fma2(pack2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nmb2), a, b);
        const float p0 = ex2(a), p1 = ex2(b);
        rs2 = add2(rs2, pack2(p0, p1));
        __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
        sv[i] = *(uint32_t *)&t;
    }
    float rs0, rs1;
    unpack2(rs2, rs0, rs1);
    const float alpha = ex2((m - m_new) * LOG2E);
    l = l * alpha + (rs0 + rs1);
#ifdef NB200_ATTN_TIMING
    pin16(sv); pin16(sv + 16);
#endif
    AT_CLK(c3);
    // P(j-1).V must have RETIRED before P(j) is stored.  Nothing in the dataflow asks for it (P(j) goes to its own score buffer, and only a
    // rescale touches O), but with two CTAs per SM a tcgen05.st issued while the previous TS-form MMA is still reading ITS A operand from
    // tensor memory corrupted single rows of that product: the last windows of a 25-window batch differed from run to run by up to 0.3 in a
    // few hundred rows (scripts/gpu_attn_determinism.py: 6 of 6 repetitions; none with this wait; the same wait in front of the
    // tcgen05.ld costs 11 % instead of 3 %, issuing S one tile later 30 %).  The barrier has normally completed long before: P(j-1).V is
    // issued when the softmax of tile j starts.  `have_prev`: false only for the CTA's very first tile.
    if (have_prev) {
        ptx::mbar_wait(o_done, o_par_prev);
        ptx::tc_fence_after();
    }
    ptx::tmem_st_32x32b_x32(tSj, sv);  // P(j): 64 bf16 = 32 packed columns, over the first half of S(j)
    AT_CLK(c4);
    if (!first && moved) {  // (P(j-1).V has retired, see above)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t ov[32];
            ptx::tmem_ld_32x32b_x32(tO + hh * 32, ov);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            ptx::tmem_st_32x32b_x32(tO + hh * 32, ov);
        }
    }
#ifdef NB200_ATTN_TIMING
    if (tm && !first && moved) tm[6] += 1;
#endif
    AT_CLK(c5);
    m = m_new;
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(p_full);
#ifdef NB200_ATTN_TIMING
    if (tm) {
        const long long c6 = clock64();
        tm[0] += c1 - c0; tm[1] += c2 - c1; tm[2] += c3 - c2; tm[3] += c4 - c3; tm[4] += c5 - c4; tm[5] += c6 - c5; tm[7] += 1;
    }
#endif
}

struct AtCfg {
    static constexpr int BN = 64;
    static constexpr int KV_BYTES = BN * 64 * 2;  // 8 KB
    static constexpr int ST = 4;                  // K / V stages
    static constexpr int NS = 3;                  // S buffers in TMEM
    static constexpr int OFF_Q = 0;               // 2 x 16 KB
    static constexpr int OFF_K = 2 * TILE_BYTES;
    static constexpr int OFF_V = OFF_K + ST * KV_BYTES;
    static constexpr int OFF_BAR = OFF_V + ST * KV_BYTES;
    static constexpr int SMEM = OFF_BAR + 256;
    static constexpr int TMEM_COLS = 256;  // S0 | S1 | S2 | O
    // barrier block (8 B each): q_full[2] q_empty[2] k_full[ST] k_empty[ST] v_full[ST] v_empty[ST] s_full[NS] p_full[NS] o_done, then the TMEM slot
    static constexpr int B_QF = 0, B_QE = 16, B_KF = 32, B_KE = B_KF + 8 * ST, B_VF = B_KE + 8 * ST, B_VE = B_VF + 8 * ST, B_SF = B_VE + 8 * ST,
                         B_PF = B_SF + 8 * NS, B_OD = B_PF + 8 * NS, B_SLOT = B_OD + 8;
    static_assert(B_SLOT + 4 <= 256, "barrier block");
};
static_assert(2 * (AtCfg::SMEM + 1024) <= 233472, "two attention CTAs must fit one SM");

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, bf16 *__restrict__ out, int T, int d, int n_heads, int n_items,
                unsigned long long *dbg) {
    using L = AtCfg;
    constexpr int BN = L::BN, ST = L::ST, NS = L::NS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t sbase = ptx::smem_u32(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = (T + AT_BM - 1) / AT_BM;
    const int nkv = (T + BN - 1) / BN;
    const uint32_t sQ = sbase + L::OFF_Q, sK = sbase + L::OFF_K, sV = sbase + L::OFF_V;
    const uint32_t bar = sbase + L::OFF_BAR;
    volatile uint32_t *tmem_slot_ptr = (volatile uint32_t *)(smem_raw + L::OFF_BAR + L::B_SLOT);
    // items this CTA owns: blockIdx.x, blockIdx.x + gridDim.x, ...   item -> (query tile fastest, then head, then window)
    const int my_items = n_items > (int)blockIdx.x ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0 && (sbase & 1023u)) __trap();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmKV);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(bar + L::B_QF + 8 * s, 1);
            ptx::mbar_init(bar + L::B_QE + 8 * s, 1);
        }
        for (int s = 0; s < ST; ++s) {
            ptx::mbar_init(bar + L::B_KF + 8 * s, 1);
            ptx::mbar_init(bar + L::B_KE + 8 * s, 1);
            ptx::mbar_init(bar + L::B_VF + 8 * s, 1);
            ptx::mbar_init(bar + L::B_VE + 8 * s, 1);
        }
        for (int s = 0; s < NS; ++s) {
            ptx::mbar_init(bar + L::B_SF + 8 * s, 1);
            ptx::mbar_init(bar + L::B_PF + 8 * s, 4);
        }
        ptx::mbar_init(bar + L::B_OD, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(bar + L::B_SLOT, L::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tO = tmem_base + NS * 64;
    ptx::griddep_wait();  // q | k | v come from the previous kernel; the set-up above overlapped its tail
    ptx::griddep_launch();

    if (warp == 0) {
        if (lane == 0) {
            int g = 0;  // running K/V tile count of this CTA
            for (int it = 0; it < my_items; ++it) {
                const int item = (int)blockIdx.x + it * (int)gridDim.x;
                const int qt = item % nq, hh = (item / nq) % n_heads, b = item / (nq * n_heads);
                const int qb = it & 1;
                if (it >= 2) ptx::mbar_wait(bar + L::B_QE + 8 * qb, (uint32_t)(((it >> 1) & 1) ^ 1));  // every S = Q.K^T of item it-2 has retired
                ptx::mbar_expect_tx(bar + L::B_QF + 8 * qb, TILE_BYTES);
                ptx::tma_load_3d(sQ + qb * TILE_BYTES, &tmQ, bar + L::B_QF + 8 * qb, hh * HEAD_DIM, qt * AT_BM, b);
                for (int j = 0; j < nkv; ++j, ++g) {
                    const int s = g % ST;
                    const uint32_t ph = (uint32_t)(((g / ST) & 1) ^ 1);
                    if (g >= ST) ptx::mbar_wait(bar + L::B_KE + 8 * s, ph);
                    ptx::mbar_expect_tx(bar + L::B_KF + 8 * s, L::KV_BYTES);
                    ptx::tma_load_3d(sK + s * L::KV_BYTES, &tmKV, bar + L::B_KF + 8 * s, d + hh * HEAD_DIM, j * BN, b);
                    if (g >= ST) ptx::mbar_wait(bar + L::B_VE + 8 * s, ph);
                    ptx::mbar_expect_tx(bar + L::B_VF + 8 * s, L::KV_BYTES);
                    ptx::tma_load_3d(sV + s * L::KV_BYTES, &tmKV, bar + L::B_VF + 8 * s, 2 * d + hh * HEAD_DIM, j * BN, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(AT_BM, BN, 0, 0);
            constexpr uint32_t idesc_o = ptx::make_idesc_bf16(AT_BM, HEAD_DIM, 0, 1);  // B (= V) is MN-major
            const int G = my_items * nkv;
            // S(g): g-th score tile of this CTA (item g / nkv, key tile g % nkv) -> S buffer g % NS
            int s_it = 0, s_j = 0;  // (item, tile) of the next S to issue
            auto issue_s = [&](int g) {
                const int qb = s_it & 1;
                if (s_j == 0) ptx::mbar_wait(bar + L::B_QF + 8 * qb, (uint32_t)((s_it >> 1) & 1));
                const int s = g % ST;
                ptx::mbar_wait(bar + L::B_KF + 8 * s, (uint32_t)((g / ST) & 1));
                ptx::tc_fence_after();
                const uint32_t qa = sQ + qb * TILE_BYTES, kb = sK + s * L::KV_BYTES, ts = tmem_base + (g % NS) * 64;
#pragma unroll
                for (int k = 0; k < HEAD_DIM / 16; ++k)
                    ptx::mma_bf16_ss(ts, ptx::make_sw128_desc(qa + k * 32, 16, 1024), ptx::make_sw128_desc(kb + k * 32, 16, 1024), idesc_s, k != 0);
                ptx::mma_commit(bar + L::B_KE + 8 * s);
                ptx::mma_commit(bar + L::B_SF + 8 * (g % NS));
                if (++s_j == nkv) {
                    ptx::mma_commit(bar + L::B_QE + 8 * qb);  // the Q buffer may be refilled
                    s_j = 0;
                    ++s_it;
                }
            };
            for (int g = 0; g < NS - 1 && g < G; ++g) issue_s(g);
            int j = 0;
            for (int g = 0; g < G; ++g) {
                const int s = g % ST, sb = g % NS;
                ptx::mbar_wait(bar + L::B_VF + 8 * s, (uint32_t)((g / ST) & 1));
                ptx::mbar_wait(bar + L::B_PF + 8 * sb, (uint32_t)((g / NS) & 1));  // P(g) sits in the columns of S(g); O rescaled (or, j = 0, read out)
                ptx::tc_fence_after();
                const uint32_t vb = sV + s * L::KV_BYTES, tp = tmem_base + sb * 64;
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks)
                    ptx::mma_bf16_ts(tO, tp + ks * 8, ptx::make_sw128_desc(vb + ks * 2048, 1024, 1024), idesc_o, (j | ks) != 0);
                ptx::mma_commit(bar + L::B_VE + 8 * s);
                ptx::mma_commit(bar + L::B_OD);
                if (++j == nkv) j = 0;
                if (g + NS - 1 < G) issue_s(g + NS - 1);  // into the buffer S(g-1)/P(g-1) used: P(g-1).V was issued one iteration ago
            }
        }
    } else {
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        int g = 0;
        long long *tm = nullptr;
#ifdef NB200_ATTN_TIMING
        long long tm_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, e0 = 0, e1 = 0, e2 = 0;
        const long long k0 = clock64();
        tm = tm_acc;
#endif
        for (int it = 0; it < my_items; ++it) {
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            const int qt = item % nq, hh = (item / nq) % n_heads, b = item / (nq * n_heads);
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < nkv - 1; ++j, ++g) {
                const int sb = g % NS;
                softmax_tile<false>(tmem_base + lane_off + sb * 64, tO + lane_off, bar + L::B_SF + 8 * sb, (uint32_t)((g / NS) & 1), bar + L::B_PF + 8 * sb,
                                     bar + L::B_OD, (uint32_t)((g - 1) & 1), lane, j == 0, g > 0, 64, m, l, tm);
            }
            {
                const int sb = g % NS;
                softmax_tile<true>(tmem_base + lane_off + sb * 64, tO + lane_off, bar + L::B_SF + 8 * sb, (uint32_t)((g / NS) & 1), bar + L::B_PF + 8 * sb,
                                    bar + L::B_OD, (uint32_t)((g - 1) & 1), lane, nkv == 1, g > 0, T - (nkv - 1) * BN, m, l, tm);
                ++g;
            }
            AT_CLK(e0);
            ptx::mbar_wait(bar + L::B_OD, (uint32_t)((g - 1) & 1));  // last P.V of this item
            ptx::tc_fence_after();
            AT_CLK(e1);
            const float inv = 1.0f / l;
            const int q = qt * AT_BM + r;
            bf16 *orow = out + ((size_t)b * T + q) * d + hh * HEAD_DIM;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t ov[32];
                ptx::tmem_ld_32x32b_x32(tO + lane_off + hf * 32, ov);
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
                        *(uint4 *)(orow + hf * 32 + i) = v;
                    }
                }
            }
            ptx::tc_fence_before();  // O is in registers: the next item's first P.V (ordered behind this warp's next p_full arrive) may overwrite it
#ifdef NB200_ATTN_TIMING
            e2 = clock64();
            tm[8] += e1 - e0; tm[9] += e2 - e1; tm[10] += 1;
#endif
        }
#ifdef NB200_ATTN_TIMING
        if (dbg && lane == 0) {
            tm[11] = clock64() - k0;
            for (int i = 0; i < 12; ++i) atomicAdd(dbg + i, (unsigned long long)tm[i]);
            atomicAdd(dbg + 12, 1ull);
        }
#endif
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

}  // namespace

int attn_tc_init(nb200_ctx *ctx) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg::SMEM));
    return NB200_OK;
}

// qkv: [B*T][3d] bf16 (q | k | v, heads of 64 inside each third); out: [B*T][d] bf16
int launch_attention_tc(nb200_ctx *ctx, const bf16 *qkv, bf16 *out, int B, int T, int n_heads) {
    const int d = n_heads * HEAD_DIM;
    CUtensorMap tq, tkv;
    uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
    uint32_t box_q[3] = {HEAD_DIM, AT_BM, 1}, box_kv[3] = {HEAD_DIM, AtCfg::BN, 1};
    NB_TRY(tmap_encode_bf16(ctx, &tq, qkv, 3, dims, str, box_q));
    NB_TRY(tmap_encode_bf16(ctx, &tkv, qkv, 3, dims, str, box_kv));
    KernelScope ks(ctx, NB200_K_ATTN);
    const int n_items = ceil_div(T, AT_BM) * n_heads * B;   // (window, head, query tile), query tile fastest: CTAs that run together share K / V in L2
    int ctas = std::min(n_items, 2 * ctx->sm_count);  // persistent: two resident CTAs per SM
    if (const char *e = getenv("NB200_ATTN_CTAS")) ctas = std::max(1, std::min(n_items, atoi(e)));  // scripts/gpu_attn_determinism.py: grid size
    unsigned long long *dbg = nullptr;
#ifdef NB200_ATTN_TIMING
    static unsigned long long *dbg_buf = nullptr;
    if (!dbg_buf) cudaMalloc(&dbg_buf, 128);
    cudaMemsetAsync(dbg_buf, 0, 128, ctx->stream);
    dbg = dbg_buf;
#endif
    CUDA_TRY(ctx, launch_chain(ctx, 2, attn_tc_kernel, dim3(ctas), dim3(AT_THREADS), AtCfg::SMEM, 1, tq, tkv, out, T, d, n_heads, n_items, dbg));
    CUDA_TRY(ctx, cudaGetLastError());
#ifdef NB200_ATTN_TIMING  // scripts/probes: where a softmax warp's cycles go (clock64 deltas kept in registers, one atomicAdd per warp at the end)
    {
        unsigned long long hst[16];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hst, dbg_buf, 128, cudaMemcpyDeviceToHost);
        const double n = (double)hst[7], w = (double)hst[12], it = (double)hst[10];
        fprintf(stderr, "[attn timing] per warp per 64-key tile, clk: s_full wait %.0f | tmem ld %.0f | max+exp+pack %.0f | P st issue %.0f | rescale %.0f | "
                        "st wait+arrive %.0f | rescales %.3f/tile || per item: o_done wait %.0f, O write-out %.0f || per warp: total %.0f clk, %.1f tiles, %.1f items\n",
                hst[0] / n, hst[1] / n, hst[2] / n, hst[3] / n, hst[4] / n, hst[5] / n, hst[6] / n, hst[8] / it, hst[9] / it, hst[11] / w, n / w, it / w);
    }
#endif
    return NB200_OK;
}
