// mel.cu — K1: 16 kHz PCM -> Whisper log-mel, replacing candle's CPU-only `audio::pcm_to_mel`
// (called at /root/reference/src/models/whisper/model.rs:74; spec SURVEY.md §8 c-1).
//
// One CTA = 32 consecutive frames of one window (8 warps x 4 frames, 8 lanes per frame).  Samples are
// staged once into shared memory with float4 coalesced loads (each sample is read from HBM once although
// frames overlap 2.5x).  The 400-point real FFT is a 200-point complex FFT (z[n] = x[2n] + i x[2n+1]),
// factored 200 = 25 x 8: every lane does a 25-point DFT (5 x 5, Winograd-style radix-5) entirely in
// registers, then the radix-8 step runs ACROSS the 8 lanes with warp shuffles.  Power spectrum, the
// `p[j] += p[400-j]` fold (x2 on bins 1..199), the banded (<= 16 non-zeros per row) mel filterbank, log10 and
// the running global max follow in the same kernel; the output tile goes through shared memory so global
// stores are 128-byte row segments.  All arithmetic is fp32 FMA with fp64-derived twiddle tables (H2).
#include "common.cuh"

#include <math.h>

#include <algorithm>

namespace {

constexpr int FR_PER_CTA = 32;
constexpr int MEL_THREADS = 256;
constexpr int MEL_BAND = 16;       // max non-zeros per filter row we store (80: <= 14, 128: <= 9)
constexpr int BAND_STRIDE = 17;    // padded to dodge bank conflicts
constexpr int POW_STRIDE = 232;    // == 8 (mod 32): the four frames of a warp hit disjoint banks
constexpr int OUT_STRIDE = 36;
constexpr int PCM_SMEM = 176 * (FR_PER_CTA - 1) + 400 + 32 + 16;  // padded index space, see pcm_idx()

__constant__ float W25R[17] = {1.f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                               0.53582679497899655f, 0.30901699437494745f, 0.062790519529313527f,
                               -0.1873813145857246f, -0.42577929156507272f, -0.63742398974868975f,
                               -0.80901699437494734f, -0.92977648588825135f, -0.99211470131447776f,
                               -0.99211470131447788f, -0.92977648588825146f, -0.80901699437494778f,
                               -0.63742398974868952f};
__constant__ float W25I[17] = {-0.f, -0.24868988716485479f, -0.48175367410171532f, -0.68454710592868862f,
                               -0.84432792550201508f, -0.95105651629515353f, -0.99802672842827156f,
                               -0.98228725072868872f, -0.90482705246601947f, -0.77051324277578925f,
                               -0.58778525229247325f, -0.36812455268467814f, -0.12533323356430454f,
                               0.12533323356430429f, 0.36812455268467792f, 0.58778525229247269f,
                               0.77051324277578936f};

// sample s (relative to the CTA's first sample) -> padded smem index: +16 floats per hop so that the four
// frames a warp works on land in different bank halves for the 8-byte loads
__device__ __forceinline__ int pcm_idx(int s) { return s + 16 * (s / HOP); }

__device__ __forceinline__ unsigned enc_max(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_max(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// forward 5-point DFT in place: (r[i*S], q[i*S]) i = 0..4
template <int S>
__device__ __forceinline__ void dft5(float *r, float *q) {
    const float C1 = 0.30901699437494745f, S1 = 0.9510565162951535f;
    const float C2 = -0.8090169943749473f, S2 = 0.5877852522924732f;
    float a0r = r[0], a0i = q[0];
    float t1r = r[S] + r[4 * S], t1i = q[S] + q[4 * S];
    float t2r = r[2 * S] + r[3 * S], t2i = q[2 * S] + q[3 * S];
    float t3r = r[S] - r[4 * S], t3i = q[S] - q[4 * S];
    float t4r = r[2 * S] - r[3 * S], t4i = q[2 * S] - q[3 * S];
    float m1r = a0r + C1 * t1r + C2 * t2r, m1i = a0i + C1 * t1i + C2 * t2i;
    float m2r = a0r + C2 * t1r + C1 * t2r, m2i = a0i + C2 * t1i + C1 * t2i;
    float s1r = S1 * t3r + S2 * t4r, s1i = S1 * t3i + S2 * t4i;
    float s2r = S2 * t3r - S1 * t4r, s2i = S2 * t3i - S1 * t4i;
    r[0] = a0r + t1r + t2r;
    q[0] = a0i + t1i + t2i;
    // X1 = m1 - i*s1, X4 = m1 + i*s1, X2 = m2 - i*s2, X3 = m2 + i*s2   (-i*(a+ib) = b - ia)
    r[S] = m1r + s1i;      q[S] = m1i - s1r;
    r[4 * S] = m1r - s1i;  q[4 * S] = m1i + s1r;
    r[2 * S] = m2r + s2i;  q[2 * S] = m2i - s2r;
    r[3 * S] = m2r - s2i;  q[3 * S] = m2i + s2r;
}

__global__ void mel_init_max_kernel(unsigned *mel_max, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mel_max[i] = enc_max(-10.0f);  // the zero-pad frames candle appends are exactly log10(1e-10)
}

// Persistent: each CTA loops over (window, 32-frame tile) work items; tables and filters are staged once per CTA, and the
// samples of the NEXT tile are prefetched into registers while the current tile is transformed.
__global__ void __launch_bounds__(MEL_THREADS, 2)
mel_kernel(const float *__restrict__ pcm, const int *__restrict__ pcm_len, const float *__restrict__ tables,
           const float *__restrict__ filt_vals, const int *__restrict__ filt_start, const int *__restrict__ filt_row,
           const int *__restrict__ slot_len, int n_mel, int n_tiles, int tiles_per_window, int tile0, float *__restrict__ logmel,
           unsigned *__restrict__ mel_max) {
    extern __shared__ __align__(16) float smem[];
    float *s_pcm = smem;                                  // PCM_SMEM (aliased by s_out after the FFT loads)
    float *s_out = smem;                                  // [n_mel][OUT_STRIDE]
    const int a_sz = (PCM_SMEM > n_mel * OUT_STRIDE) ? PCM_SMEM : n_mel * OUT_STRIDE;
    float *s_pow = smem + a_sz;                           // [32][POW_STRIDE]
    float *s_hann = s_pow + FR_PER_CTA * POW_STRIDE;      // [400]
    float2 *s_tw200 = (float2 *)(s_hann + 400);           // [25][8]
    float2 *s_tw400 = s_tw200 + 200;                      // [201] (+1 pad)
    float *s_fv = (float *)(s_tw400 + 202);               // [n_mel][BAND_STRIDE]  (slot-major: entry s*8 + lane)
    int *s_fstart = (int *)(s_fv + n_mel * BAND_STRIDE);  // [n_mel]
    int *s_frow = s_fstart + n_mel;                       // [n_mel]
    int *s_slen = s_frow + n_mel;                         // [n_mel / 8]
    __shared__ float s_wmax[MEL_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- once per CTA: tables and the slot-ordered banded filterbank ----------------------------------------
    for (int i = tid; i < 400; i += MEL_THREADS) s_hann[i] = tables[i];
    for (int i = tid; i < 400; i += MEL_THREADS) ((float *)s_tw200)[i] = tables[400 + i];
    for (int i = tid; i < 402; i += MEL_THREADS) ((float *)s_tw400)[i] = tables[800 + i];
    for (int i = tid; i < n_mel * MEL_BAND; i += MEL_THREADS)
        s_fv[(i / MEL_BAND) * BAND_STRIDE + (i % MEL_BAND)] = filt_vals[i];
    for (int i = tid; i < n_mel; i += MEL_THREADS) {
        s_fstart[i] = filt_start[i];
        s_frow[i] = filt_row[i];
    }
    for (int i = tid; i < n_mel / 8; i += MEL_THREADS) s_slen[i] = slot_len[i];

    constexpr int N_S = (FR_PER_CTA - 1) * HOP + N_FFT;  // 5360 samples = 1340 float4
    constexpr int N_V4 = N_S / 4;
    constexpr int PF = (N_V4 + MEL_THREADS - 1) / MEL_THREADS;  // float4 registers per thread
    float4 pf[PF];
    auto prefetch = [&](int tile) {
        const int b = tile / tiles_per_window, f0 = (tile - b * tiles_per_window) * FR_PER_CTA;
        const int len = pcm_len[b];
        const float *wpcm = pcm + (size_t)b * N_SAMPLES;
        const int s_base = f0 * HOP;
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = (tid + u * MEL_THREADS) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < N_S) {
                const int g = s_base + i;
                if (g + 3 < len) v = __ldg((const float4 *)(wpcm + g));
                else {
                    if (g + 0 < len) v.x = __ldg(wpcm + g + 0);
                    if (g + 1 < len) v.y = __ldg(wpcm + g + 1);
                    if (g + 2 < len) v.z = __ldg(wpcm + g + 2);
                    if (g + 3 < len) v.w = __ldg(wpcm + g + 3);
                }
            }
            pf[u] = v;
        }
    };
    int tile = tile0 + blockIdx.x;
    if (tile < n_tiles) prefetch(tile);

    const int n2 = lane & 7, grp = lane >> 3;
    const int fi = warp * 4 + grp;
    for (; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_window, f0 = (tile - b * tiles_per_window) * FR_PER_CTA;
        __syncthreads();  // previous tile's output store no longer reads s_out (aliases s_pcm); tables visible on the first pass
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int i = (tid + u * MEL_THREADS) * 4;
            if (i < N_S) *(float4 *)(s_pcm + pcm_idx(i)) = pf[u];
        }
        __syncthreads();
        if (tile + (int)gridDim.x < n_tiles) prefetch(tile + gridDim.x);  // in flight while this tile is transformed

        // ---- per-lane 25-point DFT over n1 (element n1 of lane n2 is z[8*n1 + n2]) -----------------------------
        float zr[25], zi[25];
#pragma unroll
        for (int n1 = 0; n1 < 25; ++n1) {
            int sidx = fi * HOP + 16 * n1 + 2 * n2;
            float2 v = *(const float2 *)(s_pcm + pcm_idx(sidx));
            float2 hw = *(const float2 *)(s_hann + 16 * n1 + 2 * n2);
            zr[n1] = v.x * hw.x;
            zi[n1] = v.y * hw.y;
        }
        // n1 = 5a + b: DFT over a for each b (stride 5), twiddle W25^(b*c), DFT over b for each c (stride 1)
#pragma unroll
        for (int bb = 0; bb < 5; ++bb) dft5<5>(zr + bb, zi + bb);  // now index 5c + b holds Y_b[c]
#pragma unroll
        for (int c = 1; c < 5; ++c) {
#pragma unroll
            for (int bb = 1; bb < 5; ++bb) {
                float wr = W25R[bb * c], wi = W25I[bb * c];
                float xr = zr[5 * c + bb], xi = zi[5 * c + bb];
                zr[5 * c + bb] = xr * wr - xi * wi;
                zi[5 * c + bb] = xr * wi + xi * wr;
            }
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) dft5<1>(zr + 5 * c, zi + 5 * c);  // index 5c + e holds X[c + 5e]
        // reorder to k1 order and apply the 200-point twiddle W200^(n2*k1)
        float yr[25], yi[25];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
#pragma unroll
            for (int e = 0; e < 5; ++e) {
                const int k1 = c + 5 * e;
                float2 w = s_tw200[k1 * 8 + n2];
                float xr = zr[5 * c + e], xi = zi[5 * c + e];
                yr[k1] = xr * w.x - xi * w.y;
                yi[k1] = xr * w.y + xi * w.x;
            }
        }
        // ---- radix-8 across the 8 lanes (DIF; lane n2 ends up holding k2 = bitrev3(n2)) ------------------------
        {
            const float R = 0.70710678118654752f;
            const int j = n2 & 3;
            const bool up4 = n2 & 4, up2 = n2 & 2, up1 = n2 & 1;
            // W8^j for the upper half of stage 1, identity otherwise
            float w1r = !up4 ? 1.f : (j == 0 ? 1.f : (j == 1 ? R : (j == 2 ? 0.f : -R)));
            float w1i = !up4 ? 0.f : (j == 0 ? 0.f : (j == 1 ? -R : (j == 2 ? -1.f : -R)));
            const float s4 = up4 ? -1.f : 1.f, s2 = up2 ? -1.f : 1.f, s1 = up1 ? -1.f : 1.f;
            const bool rot2 = up2 && (n2 & 1);  // multiply by W4^1 = -i
#pragma unroll
            for (int k = 0; k < 25; ++k) {
                float pr = __shfl_xor_sync(0xffffffffu, yr[k], 4), pi = __shfl_xor_sync(0xffffffffu, yi[k], 4);
                float tr = pr + s4 * yr[k], ti = pi + s4 * yi[k];
                float ar = tr * w1r - ti * w1i, ai = tr * w1i + ti * w1r;
                pr = __shfl_xor_sync(0xffffffffu, ar, 2);
                pi = __shfl_xor_sync(0xffffffffu, ai, 2);
                tr = pr + s2 * ar;
                ti = pi + s2 * ai;
                ar = rot2 ? ti : tr;
                ai = rot2 ? -tr : ti;
                pr = __shfl_xor_sync(0xffffffffu, ar, 1);
                pi = __shfl_xor_sync(0xffffffffu, ai, 1);
                yr[k] = pr + s1 * ar;
                yi[k] = pi + s1 * ai;
            }
        }
        // ---- real-FFT post-processing + power ---------------------------------------------------------------
        {
            const int k2 = ((n2 & 1) << 2) | (n2 & 2) | ((n2 & 4) >> 2);
            const int pk2 = (8 - k2) & 7;
            const int src0 = (lane & ~7) | (((pk2 & 1) << 2) | (pk2 & 2) | ((pk2 & 4) >> 2));
            float *prow = s_pow + fi * POW_STRIDE;
#pragma unroll
            for (int k1 = 0; k1 < 25; ++k1) {
                float qr, qi;
                if (k1 == 0) {
                    qr = __shfl_sync(0xffffffffu, yr[0], src0);
                    qi = __shfl_sync(0xffffffffu, yi[0], src0);
                } else {
                    qr = __shfl_xor_sync(0xffffffffu, yr[25 - k1], 7);
                    qi = __shfl_xor_sync(0xffffffffu, yi[25 - k1], 7);
                }
                const int k = k1 + 25 * k2;
                float er = 0.5f * (yr[k1] + qr), ei = 0.5f * (yi[k1] - qi);
                float orr = 0.5f * (yi[k1] + qi), oi = -0.5f * (yr[k1] - qr);
                float2 w = s_tw400[k];
                float xr = er + w.x * orr - w.y * oi;
                float xi = ei + w.x * oi + w.y * orr;
                float p = xr * xr + xi * xi;
                prow[k] = (k == 0) ? p : 2.0f * p;  // candle: p[j] += p[400-j] for j = 1..199
                if (k1 == 0 && k2 == 0) {
                    float x200 = yr[0] - yi[0];
                    prow[200] = x200 * x200;
                }
            }
        }
        __syncthreads();  // all warps done reading s_pcm (aliased by s_out); power rows visible

        // ---- banded mel projection (slot-ordered: the 8 lanes of a slot share one trip count), log10, running max ---
        float vmax = -10.0f;
        {
            const float *prow = s_pow + fi * POW_STRIDE;
            const bool valid = (f0 + fi) < N_FRAMES;
            const int n_slots = n_mel >> 3;
            for (int sl = 0; sl < n_slots; ++sl) {
                const int e = sl * 8 + n2;
                const int st = s_fstart[e], ln = s_slen[sl];  // ln is warp-uniform
                const float *fv = s_fv + e * BAND_STRIDE;
                float sum = 0.f;
                for (int j = 0; j < ln; ++j) sum = fmaf(prow[st + j], fv[j], sum);
                // log10(x) = log2(x) * log10(2); lg2.approx is within 2^-22 absolute on this range (1e-7 after /4)
                float v = __log2f(fmaxf(sum, 1e-10f)) * 0.30102999566398120f;
                s_out[s_frow[e] * OUT_STRIDE + fi] = v;
                if (valid) vmax = fmaxf(vmax, v);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if (lane == 0) s_wmax[warp] = vmax;
        __syncthreads();
        if (tid == 0) {
            float m = s_wmax[0];
            for (int i = 1; i < MEL_THREADS / 32; ++i) m = fmaxf(m, s_wmax[i]);
            atomicMax(mel_max + b, enc_max(m));
        }
        // ---- coalesced store: rows of 32 frames = 128 B ---------------------------------------------------------
        float *obase = logmel + (size_t)b * n_mel * N_FRAMES;
        for (int i = tid; i < n_mel * (FR_PER_CTA / 4); i += MEL_THREADS) {
            int m = i >> 3, seg = (i & 7) * 4;
            int f = f0 + seg;
            if (f < N_FRAMES) {
                float4 v = *(const float4 *)(s_out + m * OUT_STRIDE + seg);
                *(float4 *)(obase + (size_t)m * N_FRAMES + f) = v;
            }
        }
    }
}

// logmel [b][m][f] -> normalised `max(x, max-8)/4+1` -> (optional) mel_norm [b][m][f] f32 and the time-major
// conv1 operand melT [b][1+f][m] (T = float | bf16).  `raw` = 0: input is already normalised (mel supplied by
// the host through nb200_encoder_forward).  32x32 smem transpose tiles.
template <typename T>
__global__ void mel_norm_kernel(const float *__restrict__ in, const unsigned *__restrict__ mel_max, int raw,
                                int n_mel, float *__restrict__ mel_norm, T *__restrict__ melT) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int f0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    const float floorv = raw ? dec_max(mel_max[b]) - 8.0f : 0.f;
    const float *ib = in + (size_t)b * n_mel * N_FRAMES;
    for (int r = ty; r < 32; r += 8) {
        int m = m0 + r, f = f0 + tx;
        float v = 0.f;
        if (m < n_mel && f < N_FRAMES) {
            v = ib[(size_t)m * N_FRAMES + f];
            if (raw) {
                v = fmaxf(v, floorv) / 4.0f + 1.0f;
                if (mel_norm) mel_norm[(size_t)b * n_mel * N_FRAMES + (size_t)m * N_FRAMES + f] = v;
            }
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    T *ob = melT + (size_t)b * (N_FRAMES + 2) * n_mel;
    for (int r = ty; r < 32; r += 8) {
        int f = f0 + r, m = m0 + tx;
        if (m < n_mel && f < N_FRAMES) ob[(size_t)(1 + f) * n_mel + m] = (T)tile[tx][r];
    }
}

}  // namespace

static size_t mel_smem_bytes(int n_mel) {
    size_t a = (PCM_SMEM > n_mel * OUT_STRIDE) ? PCM_SMEM : n_mel * OUT_STRIDE;
    size_t fl = a + FR_PER_CTA * POW_STRIDE + 400 + 400 + 404 + (size_t)n_mel * BAND_STRIDE + 2 * n_mel + n_mel / 8 + 8;
    return fl * 4;
}

int mel_setup_tables(nb200_ctx *ctx) {
    // per-device opt-in to > 48 KB dynamic shared memory
    CUDA_TRY(ctx, cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    std::vector<float> t(800 + 402 + 2, 0.f);
    // periodic Hann in f32 as candle computes it: 0.5*(1 - cos(2*pi*i/400))
    const float two_pi = 3.14159265358979323846f + 3.14159265358979323846f;
    for (int i = 0; i < 400; ++i) t[i] = 0.5f * (1.0f - cosf((two_pi * (float)i) / 400.0f));
    // fp64-derived twiddles
    for (int k1 = 0; k1 < 25; ++k1)
        for (int n2 = 0; n2 < 8; ++n2) {
            double a = -2.0 * M_PI * (double)(n2 * k1) / 200.0;
            t[400 + 2 * (k1 * 8 + n2)] = (float)cos(a);
            t[400 + 2 * (k1 * 8 + n2) + 1] = (float)sin(a);
        }
    for (int k = 0; k <= 200; ++k) {
        double a = -2.0 * M_PI * (double)k / 400.0;
        t[800 + 2 * k] = (float)cos(a);
        t[800 + 2 * k + 1] = (float)sin(a);
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->mel_tables, t.data(), t.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int mel_setup_filters(nb200_ctx *ctx, const float *filters, int n_mel) {
    // Banded rows, re-ordered into SLOTS of 8 rows of similar length (one row per lane of a frame's 8-lane group) so the
    // projection loop has a warp-uniform trip count: entry e = slot*8 + lane holds (mel row, first bin, padded weights).
    std::vector<int> lo(n_mel, 0), len(n_mel, 0);
    for (int m = 0; m < n_mel; ++m) {
        int l = -1, h = -1;
        for (int k = 0; k < N_BINS; ++k)
            if (filters[(size_t)m * N_BINS + k] != 0.0f) {
                if (l < 0) l = k;
                h = k;
            }
        if (l < 0) continue;
        if (h - l + 1 > MEL_BAND)
            return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "mel filter row %d spans %d bins (> %d)", m, h - l + 1, MEL_BAND);
        lo[m] = l;
        len[m] = h - l + 1;
    }
    std::vector<int> order(n_mel);
    for (int m = 0; m < n_mel; ++m) order[m] = m;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] < len[b]; });
    std::vector<float> vals((size_t)n_mel * MEL_BAND, 0.f);
    std::vector<int> start(n_mel, 0), row(n_mel, 0), slen(n_mel / 8, 0);
    for (int s = 0; s < n_mel / 8; ++s) {
        int L = 1;
        for (int l = 0; l < 8; ++l) L = std::max(L, len[order[s * 8 + l]]);
        slen[s] = L;
        for (int l = 0; l < 8; ++l) {
            const int e = s * 8 + l, m = order[e];
            int st = std::min(lo[m], N_BINS - L);  // keep every read inside the 201 valid power bins
            if (st < 0) st = 0;
            row[e] = m;
            start[e] = st;
            for (int k = lo[m]; k < lo[m] + len[m]; ++k) vals[(size_t)e * MEL_BAND + (k - st)] = filters[(size_t)m * N_BINS + k];
        }
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->filt_vals, vals.data(), vals.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->filt_start, start.data(), n_mel * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->filt_len, row.data(), n_mel * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->mel_slot_len, slen.data(), (n_mel / 8) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return NB200_OK;
}

int launch_mel(nb200_ctx *ctx, int n_windows) {
    const int n_mel = ctx->cfg.num_mel_bins;
    size_t smem = mel_smem_bytes(n_mel);
    {
        KernelScope ks(ctx, NB200_K_MISC);
        mel_init_max_kernel<<<ceil_div(n_windows, 128), 128, 0, ctx->stream>>>(ctx->mel_max, n_windows);
    }
    {
        KernelScope ks(ctx, NB200_K_MEL);
        const int tpw = ceil_div(N_FRAMES, FR_PER_CTA), n_tiles = tpw * n_windows;
        const int grid = n_tiles < 2 * ctx->sm_count ? n_tiles : 2 * ctx->sm_count;  // persistent: 2 CTAs per SM
        mel_kernel<<<grid, MEL_THREADS, smem, ctx->stream>>>(ctx->pcm, ctx->pcm_len, ctx->mel_tables, ctx->filt_vals, ctx->filt_start,
                                                             ctx->filt_len, ctx->mel_slot_len, n_mel, n_tiles, tpw, 0, ctx->logmel, ctx->mel_max);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

static int launch_norm_impl(nb200_ctx *ctx, int n_windows, const float *in, int raw, float *mel_norm) {
    const int n_mel = ctx->cfg.num_mel_bins;
    KernelScope ks(ctx, NB200_K_MEL_NORM);
    dim3 grid(ceil_div(N_FRAMES, 32), ceil_div(n_mel, 32), n_windows), block(32, 8);
    if (ctx->compute == NB200_BF16)
        mel_norm_kernel<bf16><<<grid, block, 0, ctx->stream>>>(in, ctx->mel_max, raw, n_mel, mel_norm, (bf16 *)ctx->melT);
    else
        mel_norm_kernel<float><<<grid, block, 0, ctx->stream>>>(in, ctx->mel_max, raw, n_mel, mel_norm, (float *)ctx->melT);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

namespace {
// streaming helpers (window 0): running max over the frames that hold data, frame shift when norma seeks forward
__global__ void mel_window_max_kernel(const float *__restrict__ logmel, int n_mel, unsigned *__restrict__ mel_max) {
    __shared__ float red[32];
    float m = -10.0f;  // candle's zero-pad frames
    for (int i = threadIdx.x; i < n_mel * N_FRAMES; i += blockDim.x) m = fmaxf(m, logmel[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
        mel_max[0] = enc_max(m);
    }
}
__global__ void fill_kernel(float *p, float v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// out-of-place shift of window 0: pcm by `ns` samples, logmel by `nf` frames (tail filled with silence)
__global__ void stream_shift_kernel(const float *__restrict__ pcm_in, float *__restrict__ pcm_out, int ns, int new_len, const float *__restrict__ lm_in,
                                    float *__restrict__ lm_out, int nf, int n_mel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N_SAMPLES) pcm_out[i] = (i < new_len) ? pcm_in[i + ns] : 0.f;
    if (i < n_mel * N_FRAMES) {
        const int m = i / N_FRAMES, f = i - m * N_FRAMES;
        lm_out[i] = (f + nf < N_FRAMES) ? lm_in[m * N_FRAMES + f + nf] : -10.0f;
    }
}
}  // namespace

int mel_stream_reset(nb200_ctx *ctx) {
    KernelScope ks(ctx, NB200_K_MISC);
    const size_t n = (size_t)ctx->cfg.num_mel_bins * N_FRAMES;
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->logmel, -10.0f, n);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->pcm_len, 0, 4, ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

// recompute the 32-frame tiles of window 0 that contain frames [f_lo, f_hi)
int mel_stream_update(nb200_ctx *ctx, int f_lo, int f_hi) {
    const int n_mel = ctx->cfg.num_mel_bins;
    const int tpw = ceil_div(N_FRAMES, FR_PER_CTA);
    int t0 = f_lo / FR_PER_CTA, t1 = ceil_div(f_hi, FR_PER_CTA);
    if (t1 > tpw) t1 = tpw;
    if (t0 >= t1) return NB200_OK;
    KernelScope ks(ctx, NB200_K_MEL);
    mel_kernel<<<t1 - t0, MEL_THREADS, mel_smem_bytes(n_mel), ctx->stream>>>(ctx->pcm, ctx->pcm_len, ctx->mel_tables, ctx->filt_vals, ctx->filt_start,
                                                                            ctx->filt_len, ctx->mel_slot_len, n_mel, t1, tpw, t0, ctx->logmel, ctx->mel_max);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int mel_stream_window_max(nb200_ctx *ctx) {
    KernelScope ks(ctx, NB200_K_MEL_NORM);
    mel_window_max_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->logmel, ctx->cfg.num_mel_bins, ctx->mel_max);
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

// drop `ns` samples (and ns / 160 frames) from the front of window 0; scratch = mel_norm (pcm) and logmel row of window 1.. is
// not available for max_batch = 1, so the shift goes through a temporary allocation owned by the ctx
int mel_stream_shift(nb200_ctx *ctx, int ns, int new_len, float *tmp_pcm, float *tmp_lm) {
    const int n_mel = ctx->cfg.num_mel_bins;
    const int nf = ns / HOP;
    KernelScope ks(ctx, NB200_K_MISC);
    const int n = N_SAMPLES > n_mel * N_FRAMES ? N_SAMPLES : n_mel * N_FRAMES;
    stream_shift_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(ctx->pcm, tmp_pcm, ns, new_len, ctx->logmel, tmp_lm, nf, n_mel);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pcm, tmp_pcm, (size_t)N_SAMPLES * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->logmel, tmp_lm, (size_t)n_mel * N_FRAMES * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    return NB200_OK;
}

int launch_mel_norm(nb200_ctx *ctx, int n_windows) { return launch_norm_impl(ctx, n_windows, ctx->logmel, 1, ctx->mel_norm); }
int launch_mel_from_host_layout(nb200_ctx *ctx, int n_windows) { return launch_norm_impl(ctx, n_windows, ctx->mel_norm, 0, nullptr); }
