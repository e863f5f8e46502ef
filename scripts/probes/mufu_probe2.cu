// mufu_probe2.cu — round 2 probes for the attention softmax stream: does anything lift the f32 ex2 floor of
// 16 / clk / SM (1024 clk per 128 x 128 tile)?  Each kernel runs only the per-element instruction stream of one softmax variant
// (no TMEM, barriers or UMMA) and reports clk per 128-row x 128-column tile per SM.
//   V0  baseline: 4-chain max, FFMA (scale, subtract), ex2.approx.ftz.f32, FADD row sums, F2FP pack            (round 1 stream)
//   V1  row sums by MMA: V0 without the FADD chain
//   V2  ex2.approx.ftz.bf16x2 on packed scores: FFMA x2 -> F2FP pack -> ONE bf16x2 ex2; no FADD (row sums by MMA)
//   V3  packed fma.rn.f32x2 for scale / subtract + f32 ex2 + pack, no FADD
//   V4  V3 with a Cody-Waite degree-3 polynomial 2^x on the FMA pipe (fma.rn.f32x2 / add.rn.f32x2) for 1 pair in every PERIOD
//   V5  ex2.approx.f16x2 (comparison only)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe2 mufu_probe2.cu && ./mufu_probe2
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// 2^x for x <= 0 (|x| < 126): round-to-nearest split with the 1.5 * 2^23 constant, degree-3 minimax polynomial on [-0.5, 0.5], exponent add
__device__ __forceinline__ void exp2_poly2(uint64_t x, float &o0, float &o1) {
    const uint64_t MAGIC = pack2(12582912.f, 12582912.f), NMAGIC = pack2(-12582912.f, -12582912.f);
    const uint64_t t = add2(x, MAGIC);           // integer part in the low mantissa bits
    const uint64_t xi = add2(t, NMAGIC);         // rounded x
    const uint64_t NEG1 = pack2(-1.f, -1.f);
    const uint64_t f = fma2(xi, NEG1, x);        // x - xi in [-0.5, 0.5]
    const uint64_t C3 = pack2(0.0555041f, 0.0555041f), C2 = pack2(0.2402265f, 0.2402265f), C1 = pack2(0.6931472f, 0.6931472f), C0 = pack2(1.f, 1.f);
    uint64_t p = fma2(C3, f, C2);
    p = fma2(p, f, C1);
    p = fma2(p, f, C0);
    float p0, p1, t0, t1;
    unpack2(p, p0, p1);
    unpack2(t, t0, t1);
    o0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    o1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

template <int N, int V, int PERIOD>
__global__ void __launch_bounds__(128) probe(float *out, int tiles, float c) {
    float s[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = -(float)((threadIdx.x * 7 + i) & 63) * 0.05f;
    float m = 0.f, l = 0.f;
    uint32_t acc = 0;
    for (int t = 0; t < tiles; ++t) {
        float mx0 = m, mx1 = m, mx2 = m, mx3 = m;
#pragma unroll
        for (int i = 0; i < N; i += 8) {
            mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1])); mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
            mx2 = fmaxf(mx2, fmaxf(s[i + 4], s[i + 5])); mx3 = fmaxf(mx3, fmaxf(s[i + 6], s[i + 7]));
        }
        m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * 1e-3f;
        const float mb = m * c;
        float r0 = 0.f, r1 = 0.f;
        const uint64_t c2 = pack2(c, c), nmb2 = pack2(-mb, -mb);
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            float a = s[i], b = s[i + 1];
            uint32_t pk;
            if (V == 0 || V == 1) {
                a = ex2(fmaf(a, c, -mb)); b = ex2(fmaf(b, c, -mb));
                if (V == 0) { r0 += a; r1 += b; }
                __nv_bfloat162 p = __floats2bfloat162_rn(a, b); pk = *(uint32_t *)&p;
            } else if (V == 2) {
                __nv_bfloat162 p = __floats2bfloat162_rn(fmaf(a, c, -mb), fmaf(b, c, -mb));
                pk = ex2_bf16x2(*(uint32_t *)&p);
            } else if (V == 5) {
                __half2 p = __floats2half2_rn(fmaf(a, c, -mb), fmaf(b, c, -mb));
                pk = ex2_f16x2(*(uint32_t *)&p);
            } else {  // V3 / V4
                const uint64_t x = fma2(pack2(a, b), c2, nmb2);
                if (V == 4 && ((i / 2) % PERIOD) == 0) exp2_poly2(x, a, b);
                else { unpack2(x, a, b); a = ex2(a); b = ex2(b); }
                __nv_bfloat162 p = __floats2bfloat162_rn(a, b); pk = *(uint32_t *)&p;
            }
            acc ^= pk;
            // next tile's "scores": cheap data-dependent refresh (1 LOP3-class op per pair would be ideal; keep 2 FFMA like probe 1)
            const float fa = __uint_as_float((pk << 16) | 0x3f000000u) , fb = __uint_as_float((pk & 0xffff0000u) | 0x3f000000u);
            s[i] = fa * -0.5f; s[i + 1] = fb * -0.5f;
        }
        l += r0 + r1;
    }
    float z = l + m;
#pragma unroll
    for (int i = 0; i < N; ++i) z += s[i];
    if (z == 123.456f || acc == 0xdeadbeef) out[0] = z;
}

template <int N, int V, int PERIOD>
void run(int ctas_per_sm, const char *name) {
    float *out; cudaMalloc(&out, 4);
    int tiles = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<N, V, PERIOD><<<148 * ctas_per_sm, 128>>>(out, 10, 1.44f);
    cudaEventRecord(a);
    probe<N, V, PERIOD><<<148 * ctas_per_sm, 128>>>(out, tiles, 1.44f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double tiles128 = (double)ctas_per_sm * tiles * N / 128.0;
    double ns_per_tile = ms * 1e6 / tiles128;
    printf("%-44s N=%3d warps/SM=%2d: %7.1f ns per 128x128 tile per SM (= %6.0f clk @1.965 GHz; f32 MUFU floor 1024 clk)\n", name, N, ctas_per_sm * 4,
           ns_per_tile, ns_per_tile * 1.965);
    cudaFree(out);
}

// accuracy of the variants on [-24, 0]
__global__ void acc_kernel(float *err) {
    float e2 = 0.f, e4 = 0.f;
    for (int i = threadIdx.x; i < 24000; i += blockDim.x) {
        const float x = -(float)i * 1e-3f;
        const float ref = exp2f(x);
        __nv_bfloat162 p = __floats2bfloat162_rn(x, x);
        uint32_t y = ex2_bf16x2(*(uint32_t *)&p);
        const float xb = __bfloat162float(__low2bfloat16(p));
        const float got2 = __uint_as_float(y << 16);
        e2 = fmaxf(e2, fabsf(got2 - exp2f(xb)) / exp2f(xb));  // error of the instruction itself (input already rounded)
        float a, b;
        exp2_poly2(pack2(x, x), a, b);
        e4 = fmaxf(e4, fabsf(a - ref) / ref);
    }
    atomicMax((int *)err, __float_as_int(e2));
    atomicMax((int *)err + 1, __float_as_int(e4));
}

int main() {
    for (int w : {2, 4}) {
        run<64, 0, 1>(w, "V0 max+FFMA+MUFU+FADD+F2FP (round 1)");
        run<64, 1, 1>(w, "V1 no FADD (row sums by MMA)");
        run<64, 2, 1>(w, "V2 ex2.bf16x2 packed, no FADD");
        run<64, 5, 1>(w, "V5 ex2.f16x2 packed, no FADD");
        run<64, 3, 1>(w, "V3 fma.f32x2 + f32 ex2, no FADD");
        run<64, 4, 4>(w, "V4 poly on 1/4 of the pairs (f32x2)");
        run<64, 4, 3>(w, "V4 poly on 1/3 of the pairs (f32x2)");
        run<64, 4, 2>(w, "V4 poly on 1/2 of the pairs (f32x2)");
        run<64, 4, 1>(w, "V4 poly on all pairs (f32x2)");
    }
    float *err; cudaMalloc(&err, 8); cudaMemset(err, 0, 8);
    acc_kernel<<<1, 256>>>(err);
    float h[2]; cudaMemcpy(h, err, 8, cudaMemcpyDeviceToHost);
    printf("max rel err on [-24, 0]: ex2.bf16x2 (vs exact of the bf16 input) %.3e; degree-3 polynomial (vs exp2f) %.3e\n", h[0], h[1]);
    return 0;
}
