// Shared device/host helpers for the molsde_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/molsde_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "molsde_b200 targets sm_100a only"
#endif

namespace molsde {

void set_last_error(const char* msg);
int check_launch(const char* what);  // returns MOLSDE_OK or MOLSDE_ERR_CUDA (records message)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte async global->shared copy (LDGSTS); both addresses 16B aligned.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float silu_f(float x) {
    // x * sigmoid(x) as torch computes it in fp32: x / (1 + exp(-x))
    return x / (1.0f + expf(-x));
}
__device__ __forceinline__ float softplus_f(float x) {
    // F.softplus(beta=1, threshold=20): x > 20 ? x : log1p(exp(x))
    return x > 20.0f ? x : log1pf(expf(x));
}

}  // namespace molsde
