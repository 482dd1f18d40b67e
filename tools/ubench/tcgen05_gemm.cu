// Standalone bring-up of a tcgen05 TF32 GEMM tile: D[128x128] = A[128x64] . B[128x64]^T, both operands K-major in
// shared memory in the canonical no-swizzle core-matrix layout, accumulator in TMEM, 3xTF32 split for fp32 accuracy.
// Checks against an fp64 CPU product.  (Descriptor formats per cute/arch/mma_sm100_desc.hpp.)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 128, K = 64;
// canonical K-major no-swizzle layout: element (r, k) at  (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4  bytes
constexpr int SBO = 128;             // next 8-row group
constexpr int LBO = (M / 8) * 128;   // next 16-byte K chunk (M == N here)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= static_cast<uint64_t>(LBO >> 4) << 16;                    // leading byte offset, bits [16,30)
    d |= static_cast<uint64_t>(SBO >> 4) << 32;                    // stride byte offset, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                           // version = 1 (Blackwell)
    return d;                                                       // layout_type = 0 (no swizzle)
}

__global__ void __launch_bounds__(128, 1)
gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int split3, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sAh = reinterpret_cast<float*>(smem);                    // 32 KB each
    float* sAl = sAh + M * K;
    float* sBh = sAl + M * K;
    float* sBl = sBh + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        const int off = ((k >> 2) * LBO + (r >> 3) * SBO + (r & 7) * 16 + (k & 3) * 4) >> 2;
        const float a = A[i], b = B[i];
        const float ah = __uint_as_float(__float_as_uint(a) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
        sAh[off] = split3 ? ah : a; sAl[off] = a - ah;
        sBh[off] = split3 ? bh : b; sBl[off] = b - bh;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        // instruction descriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
        const int nterm = split3 ? 3 : 1;
        int first = 1;
        for (int term = 0; term < nterm; ++term) {
            const float* pa = (term == 0 && split3) ? sAl : sAh;     // lo*hi, hi*lo, hi*hi
            const float* pb = (term == 1) ? sBl : sBh;
            for (int kb = 0; kb < K / 8; ++kb) {
                const uint64_t da = make_desc(smem_u32(pa) + kb * 2 * LBO);
                const uint64_t db = make_desc(smem_u32(pb) + kb * 2 * LBO);
                const uint32_t acc = first ? 0u : 1u;
                first = 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait on the mbarrier (phase 0)
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    if (!done) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
}

int main() {
    std::vector<float> hA(M * K), hB(N * K), hD(M * N);
    srand(1);
    for (auto& v : hA) v = (rand() / (float)RAND_MAX) * 2 - 1;
    for (auto& v : hB) v = (rand() / (float)RAND_MAX) * 2 - 1;
    float *dA, *dB, *dD; int* dS;
    cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dB, sizeof(float) * N * K); cudaMalloc(&dD, sizeof(float) * M * N); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), sizeof(float) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), sizeof(float) * N * K, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(float) * (2 * M * K + 2 * N * K) + 1024;
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int split3 = 0; split3 < 2; ++split3) {
        cudaMemset(dD, 0, sizeof(float) * M * N); cudaMemset(dS, 0, 4);
        gemm_kernel<<<1, 128, smem>>>(dA, dB, dD, split3, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hD.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)hA[i * K + k] * hB[j * K + k];
            maxerr = fmax(maxerr, fabs(s - hD[i * N + j])); maxref = fmax(maxref, fabs(s));
        }
        printf("split3=%d cuda=%s status=%d max_abs_err=%.3e (max |ref| %.3f) rel=%.3e  D[0][0]=%f D[5][77]=%f\n", split3, cudaGetErrorString(e), st,
               maxerr, maxref, maxerr / maxref, hD[0], hD[5 * N + 77]);
    }
    return 0;
}
