// mufu_probe.cu — what the softmax instruction stream of the attention kernel can reach on its own (no TMEM, barriers or UMMA):
// per thread and "tile": N scores in registers -> max -> 2^(s*c - m) -> sum -> bf16 pack.  Reports clk per 128-row x 128-column tile per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu && ./mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int N, int MODE>  // MODE 0: MUFU only; 1: + FFMA; 2: + FADD sums; 3: + bf16 pack; 4: + max chain
__global__ void __launch_bounds__(128) probe(float *out, int tiles, float c) {
    float s[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = (float)(threadIdx.x * 7 + i) * 1e-3f * c;
    float m = 0.f, l = 0.f;
    uint32_t acc = 0;
    for (int t = 0; t < tiles; ++t) {
        if (MODE >= 4) {
            float mx0 = m, mx1 = m, mx2 = m, mx3 = m;
#pragma unroll
            for (int i = 0; i < N; i += 8) {
                mx0 = fmaxf(mx0, fmaxf(s[i], s[i + 1])); mx1 = fmaxf(mx1, fmaxf(s[i + 2], s[i + 3]));
                mx2 = fmaxf(mx2, fmaxf(s[i + 4], s[i + 5])); mx3 = fmaxf(mx3, fmaxf(s[i + 6], s[i + 7]));
            }
            m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * 1e-3f;
        }
        const float mb = m * c;
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            float a = s[i], b = s[i + 1];
            if (MODE >= 1) { a = fmaf(a, c, -mb); b = fmaf(b, c, -mb); }
            a = ex2(a); b = ex2(b);
            if (MODE >= 2) { r0 += a; r1 += b; }
            if (MODE >= 3) { __nv_bfloat162 p = __floats2bfloat162_rn(a, b); acc ^= *(uint32_t *)&p; }
            s[i] = a * 0.5f - 3.f; s[i + 1] = b * 0.5f - 3.f;  // keep the values in a sane range (2 extra FFMA per pair)
        }
        l += r0 + r1;
    }
    float z = l + m;
#pragma unroll
    for (int i = 0; i < N; ++i) z += s[i];
    if (z == 123.456f || acc == 0xdeadbeef) out[0] = z;
}

template <int N, int MODE>
void run(int ctas_per_sm, const char *name) {
    float *out; cudaMalloc(&out, 4);
    int tiles = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<N, MODE><<<148 * ctas_per_sm, 128>>>(out, 10, 1.44f);
    cudaEventRecord(a);
    probe<N, MODE><<<148 * ctas_per_sm, 128>>>(out, tiles, 1.44f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    // work per SM: ctas_per_sm * tiles * 128 threads * N elements; express as time per 128x128-element tile
    double tiles128 = (double)ctas_per_sm * tiles * N / 128.0;
    double ns_per_tile = ms * 1e6 / tiles128;
    printf("%-28s N=%3d warps/SM=%2d: %7.1f ns per 128x128 tile per SM (= %6.0f clk @1.9 GHz; MUFU floor 1024 clk)\n", name, N, ctas_per_sm * 4, ns_per_tile,
           ns_per_tile * 1.9);
    cudaFree(out);
}

int main() {
    for (int w : {2, 4, 8}) {
        run<128, 0>(w, "MUFU only");
        run<128, 1>(w, "FFMA+MUFU");
        run<128, 2>(w, "FFMA+MUFU+FADD");
        run<128, 3>(w, "FFMA+MUFU+FADD+F2FP");
        run<128, 4>(w, "max+FFMA+MUFU+FADD+F2FP");
    }
    for (int w : {2, 4, 8}) run<64, 4>(w, "max+FFMA+MUFU+FADD+F2FP");
    return 0;
}
