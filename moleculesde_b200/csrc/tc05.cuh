// tcgen05 / TMEM / mbarrier primitives for kernels built on the "quad" scheme (128 threads own one 128-row tile, thread = row =
// TMEM lane; see sde2d3d.cu for the full description): split-fp16 GEMM issue against canonical K-major no-swizzle operand tiles,
// accumulator row loads, bounded mbarrier waits.  sm_100a only.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace molsde {
namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// bounded parity wait (a descriptor / protocol bug must not hang the GPU); returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 16) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    return done != 0;
}
__device__ __forceinline__ void group_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 B; LBO = byte distance between core matrices
// adjacent in K, SBO = 128 B between 8-row groups
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo_bytes) {
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
           (static_cast<uint64_t>(128u >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
}
// D[tmem, 128 lanes x N columns] (+)= A[smem, 128 x 16] . B[smem, N x 16]^T, fp16 inputs, fp32 accumulate; issued by ONE thread
template <int N>
__device__ __forceinline__ void mma_f16_m128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    constexpr uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(N) >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// split-fp16 GEMM (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, small terms first) over KSTEPS K = 16 steps; A tile [128 x 16*KSTEPS] with
// k-chunk (8 halves) stride 2048 B, B tile [N x 16*KSTEPS] with k-chunk stride N*16 B
template <int N, int KSTEPS>
__device__ __forceinline__ void mma_split_f16(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t accumulate) {
    constexpr uint32_t LBO_A = 2048, LBO_B = N * 16;
    const uint64_t dah = desc(a_hi, LBO_A), dal = desc(a_lo, LBO_A), dbh = desc(b_hi, LBO_B), dbl = desc(b_lo, LBO_B);
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint64_t da = (term == 0) ? dal : dah, db = (term == 1) ? dbl : dbh;
#pragma unroll
        for (int kb = 0; kb < KSTEPS; ++kb) {
            mma_f16_m128<N>(tmem_d, da + static_cast<uint64_t>((kb * 2 * LBO_A) >> 4), db + static_cast<uint64_t>((kb * 2 * LBO_B) >> 4), accumulate);
            accumulate = 1;
        }
    }
}
// accumulator columns [col, col + 32) of this thread's TMEM lane (all 32 lanes of the warp must call)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&f)[32]) {
    uint32_t v[32];
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
}
// two fp32 values -> packed fp16 hi parts and packed fp16 lo parts (x = hi + lo to 22 significant bits)
__device__ __forceinline__ void split2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(e0, e1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(e0 - hf.x, e1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 consecutive-k values of operand row r -> one 16 B chunk of the hi tile and of the lo tile (k-chunk stride `kstride` bytes)
__device__ __forceinline__ void store_chunk(uint8_t* hi, uint8_t* lo, int r, int kc, int kstride, const float* v) {
    uint4 h, l;
    split2(v[0], v[1], h.x, l.x);
    split2(v[2], v[3], h.y, l.y);
    split2(v[4], v[5], h.z, l.z);
    split2(v[6], v[7], h.w, l.w);
    *reinterpret_cast<uint4*>(hi + kc * kstride + r * 16) = h;
    *reinterpret_cast<uint4*>(lo + kc * kstride + r * 16) = l;
}

}  // namespace tc05
}  // namespace molsde
