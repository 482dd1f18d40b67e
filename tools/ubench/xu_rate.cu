// Throughput of the transcendental (XU) pipe and of the fp16 pack / unpack conversions on one B200 SM, in lanes per clock per SM.
// 148 CTAs x 512 threads, 8 independent chains per thread, clock64() deltas of CTA 0.   nvcc -arch=sm_100a -O3 -o xu_rate xu_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, float seed, float* out, long long* clk) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) v[i] = __sinf(v[i]);
            if (OP == 1) v[i] = exp2f(v[i]) * 0.25f;                       // MUFU.EX2 + FMUL
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) {                                                  // F2FP.F16.F32.PACK_AB (one per 2 values) + unpack
                __half2 h = __floats2half2_rn(v[i], v[i] + 1.0f);
                float2 f = __half22float2(h);
                v[i] = f.x + f.y * 0.5f;
            }
            if (OP == 4) v[i] = fmaf(v[i], 0.999f, 0.001f);
            if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); }
            if (OP == 6) {                                                  // pack only: cvt.rn.f16x2.f32, result reinterpreted
                __half2 h = __floats2half2_rn(v[i], v[i]);
                v[i] = __uint_as_float(*reinterpret_cast<uint32_t*>(&h) | 0x3c003c00u);
            }
            if (OP == 7) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * 512 + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int OP>
void run(const char* name, double ops_per_iter_elem) {
    float* o; long long* c;
    cudaMalloc(&o, 148 * 512 * 4); cudaMalloc(&c, 8);
    const int iters = 4000;
    k<OP><<<148, 512>>>(10, 0.5f, o, c);
    k<OP><<<148, 512>>>(iters, 0.5f, o, c);
    cudaDeviceSynchronize();
    long long clk = 0; cudaMemcpy(&clk, c, 8, cudaMemcpyDeviceToHost);
    const double elems = 512.0 * 8 * iters;
    printf("%-34s %8.2f lane-ops/clk/SM   (%.2f clk per warp instruction per SMSP)\n", name, elems * ops_per_iter_elem / clk,
           clk / (elems * ops_per_iter_elem / 32 / 4));
    cudaFree(o); cudaFree(c);
}
int main() {
    run<4>("FFMA", 1);
    run<0>("MUFU.SIN (__sinf)", 1);
    run<5>("MUFU.EX2 (ex2.approx.ftz)", 1);
    run<2>("MUFU.RCP (rcp.approx.ftz)", 1);
    run<7>("MUFU.TANH (tanh.approx)", 1);
    run<1>("exp2f (EX2 + range handling)", 1);
    run<6>("F2FP.F16.F32.PACK_AB alone", 1);
    run<3>("pack + unpack half2 + 2 FP32", 1);
    return 0;
}
