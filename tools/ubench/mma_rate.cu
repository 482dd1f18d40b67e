// Microbenchmark: per-SM throughput of mma.sync.m16n8k8 TF32, mma.sync m16n8k16 BF16 and FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void k_mma_tf32(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma_bf16(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, b0 = 4, b1 = 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters, float x, float y) {
    float c[16];
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], x, y);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FFMA with smem operand loads in the 4x4 register-tile pattern (16 FMA : 2 LDS.128)
__global__ void k_ffma_lds(float* out, int iters) {
    __shared__ float As[64 * 128];
    __shared__ float Ws[64 * 32];
    for (int i = threadIdx.x; i < 64 * 128; i += blockDim.x) As[i] = i * 1e-4f;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) Ws[i] = i * 1e-4f;
    __syncthreads();
    const int to = threadIdx.x & 7, te = (threadIdx.x >> 3) & 31;
    float acc[4][4] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k * 128 + te * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Ws[k * 32 + to * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    float s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* out; cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024);
    int dev_clock; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32, blocks = 148;
        float ms = timeit([&] { k_mma_tf32<<<blocks, threads>>>(out, iters); });
        double mac = (double)blocks * warps * iters * 8 * 16 * 8 * 8;
        printf("mma tf32 m16n8k8  warps/SM=%2d: %.3f ms  %.1f TMAC/s  %.0f MAC/clk/SM @%.0f MHz\n", warps, ms, mac / ms / 1e9, mac / (ms * 1e-3) / 148 / (dev_clock * 1e3), dev_clock / 1e3);
        ms = timeit([&] { k_mma_bf16<<<blocks, threads>>>(out, iters); });
        mac = (double)blocks * warps * iters * 8 * 16 * 8 * 16;
        printf("mma bf16 m16n8k16 warps/SM=%2d: %.3f ms  %.1f TMAC/s  %.0f MAC/clk/SM\n", warps, ms, mac / ms / 1e9, mac / (ms * 1e-3) / 148 / (dev_clock * 1e3));
        ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, iters * 8, 1.0001f, 0.5f); });
        mac = (double)blocks * threads * iters * 8.0 * 16;
        printf("ffma              warps/SM=%2d: %.3f ms  %.1f TMAC/s  %.0f MAC/clk/SM\n", warps, ms, mac / ms / 1e9, mac / (ms * 1e-3) / 148 / (dev_clock * 1e3));
        ms = timeit([&] { k_ffma_lds<<<blocks, threads>>>(out, iters / 20); });
        mac = (double)blocks * threads * (iters / 20) * 64.0 * 16;
        printf("ffma+lds 4x4 tile warps/SM=%2d: %.3f ms  %.1f TMAC/s  %.0f MAC/clk/SM\n", warps, ms, mac / ms / 1e9, mac / (ms * 1e-3) / 148 / (dev_clock * 1e3));
    }
    return 0;
}
