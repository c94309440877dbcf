// Micro-benchmark: issue rate of legacy warp-level mma.sync (TF32 m16n8k8, BF16 m16n8k16) vs FP32 FFMA on sm_100a.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_tf32(float* out, int iters) {
    float c[4][4] = {};
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u};
    unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_bf16(float* out, int iters) {
    float c[4][4] = {};
    unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f003f00u, 0x3e803e80u, 0x3f403f40u};
    unsigned b[2] = {0x3f803f80u, 0x3f003f00u + threadIdx.x};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, int iters) {
    float c[16];
    for (int q = 0; q < 16; ++q) c[q] = threadIdx.x * 1e-3f + q;
    float a = 1.0001f, b = 0.9999f + threadIdx.x * 1e-6f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 16; ++q) c[q] = fmaf(c[q], a, b);
    }
    float s = 0.f;
    for (int q = 0; q < 16; ++q) s += c[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
float time_kernel(K k, float* out, int grid, int block, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, block>>>(out, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<grid, block>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        int block = 256, grid = sms * warps * 32 / block;
        float t1 = time_kernel(k_tf32, out, grid, block, iters);
        float t2 = time_kernel(k_bf16, out, grid, block, iters);
        float t3 = time_kernel(k_ffma, out, grid, block, iters);
        double cyc = (double)clk * 1e3;   // Hz
        double mma_tf32 = (double)grid * (block / 32) * iters * 4 / (t1 * 1e-3) / cyc / sms;   // warp-MMAs per clk per SM
        double mma_bf16 = (double)grid * (block / 32) * iters * 4 / (t2 * 1e-3) / cyc / sms;
        double ffma = (double)grid * (block / 32) * iters * 16 / (t3 * 1e-3) / cyc / sms;      // warp-FFMAs per clk per SM
        printf("warps/SM %2d: tf32 m16n8k8 %.3f mma/clk/SM (%.0f MAC/clk/SM, %.1f TFLOP/s) | bf16 m16n8k16 %.3f mma/clk/SM (%.0f MAC/clk/SM) | ffma %.2f warp-inst/clk/SM (%.0f MAC/clk/SM)\n",
               warps, mma_tf32, mma_tf32 * 1024, mma_tf32 * 1024 * 2 * cyc * sms / 1e12, mma_bf16, mma_bf16 * 2048, ffma, ffma * 32);
    }
    return 0;
}
