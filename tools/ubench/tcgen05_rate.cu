// tcgen05.mma issue-rate microbenchmark: one CTA per SM issues `iters` x 16 back-to-back kind::tf32 / kind::f16 MMAs
// (M=128, N in {64,128,256}, K=8 tf32 / K=16 bf16) on zeroed shared-memory operands; reports clocks per MMA and MAC/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) |
           (static_cast<uint64_t>(1) << 46);
}
template <int N, int BF16>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 64 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t fmt = BF16 ? 1u : 2u;  // a/b format: 1 = BF16 (kind::f16), 2 = TF32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(smem), 2048, 128), db = make_desc(smem_u32(smem) + 128 * 64, (N / 8) * 128, 128);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (BF16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        for (long long w = 0; w < (1ll << 28) && !done; ++w)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
template <int N, int BF16>
void run(const char* name, int kdim) {
    long long* d; cudaMalloc(&d, 8);
    const int iters = 2000, smem = (128 + 256) * 64;
    cudaFuncSetAttribute(rate_kernel<N, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N, BF16><<<148, 128, smem>>>(10, d);
    rate_kernel<N, BF16><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long clk = 0; cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    const double per = double(clk) / (iters * 16.0);
    printf("%s M=128 N=%3d K=%2d: %7.1f clk/MMA  %8.0f MAC/clk/SM  (%s)\n", name, N, kdim, per, 128.0 * N * kdim / per, cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<64, 0>("tf32", 8); run<128, 0>("tf32", 8); run<256, 0>("tf32", 8);
    run<16, 1>("bf16", 16); run<32, 1>("bf16", 16);
    run<64, 1>("bf16", 16); run<128, 1>("bf16", 16); run<256, 1>("bf16", 16);
    return 0;
}
