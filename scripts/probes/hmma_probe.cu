// hmma_probe.cu — what one warp per SM sub-partition gets out of the legacy tensor path (mma.sync.m16n8k16 bf16), alone and with the operand
// loads of the fused decoder step's GEMV loop (ldmatrix.x4 for the 16 x 16 weight fragment, two 32-bit shared loads for the activations).
// Reports clk per k-step (one m16n8k16) per warp.  The GEMV loop of decoder.cu was measured at ~120-155 clk per k-step inside the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu && ./hmma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int PITCH = 2560 + 16;  // bytes per weight row in the ring tile (decoder.cu: FS_ROW_PAD)
constexpr int KS = 80;            // k-steps per tile (d_model 1280)

template <int MODE, int NACC, int WARPS>  // MODE 0: HMMA only; 1: + ldmatrix; 2: + ldmatrix + 2 LDS (the real loop)
__global__ void __launch_bounds__(WARPS * 32) probe(float *out, long long *clk, int tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16 * PITCH + 8 * PITCH) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3c003c00u + i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t arow = base + (lane & 15) * PITCH + (lane >> 4) * 16;
    const uint8_t *xs = smem + 16 * PITCH + (lane >> 2) * PITCH + (lane & 3) * 4;
    float acc[NACC][4];
#pragma unroll
    for (int q = 0; q < NACC; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = 0.f;
    uint32_t a0 = 0x3c003c00u + lane, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
#pragma unroll 4
        for (int ks = 0; ks < KS; ++ks) {
            if (MODE >= 1) asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(arow + ks * 32));
            if (MODE >= 2) {
                b0 = *(const volatile uint32_t *)(xs + ks * 32);
                b1 = *(const volatile uint32_t *)(xs + ks * 32 + 16);
            }
            float *d = acc[ks % NACC];
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NACC; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) s += acc[q][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    (void)warp;
}

template <int MODE, int NACC, int WARPS>
static void run(const char *what) {
    float *out;
    long long *clk;
    cudaMalloc(&out, 148 * WARPS * 32 * 4);
    cudaMalloc(&clk, 148 * 8);
    const int tiles = 200, smem = 24 * PITCH;
    cudaFuncSetAttribute(probe<MODE, NACC, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<MODE, NACC, WARPS><<<148, WARPS * 32, smem>>>(out, clk, 10);
    probe<MODE, NACC, WARPS><<<148, WARPS * 32, smem>>>(out, clk, tiles);
    long long h[148];
    cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    printf("%-44s warps/CTA %d  accumulators %d : %7.1f clk per k-step per warp  (%s)\n", what, WARPS, NACC, avg / ((double)tiles * KS),
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0, 4, 4>("HMMA only");
    run<0, 1, 4>("HMMA only, one dependent chain");
    run<1, 4, 4>("ldmatrix.x4 + HMMA");
    run<2, 4, 4>("ldmatrix.x4 + 2 LDS.32 + HMMA (GEMV loop)");
    run<2, 4, 8>("GEMV loop, 8 warps");
    run<2, 4, 16>("GEMV loop, 16 warps");
    run<2, 2, 4>("GEMV loop, two accumulators");
    return 0;
}
