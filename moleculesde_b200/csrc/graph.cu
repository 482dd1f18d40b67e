// Graph-construction kernels: CSR-sorted integer work, bit-exact by construction.
//   K1  radius graph     (replaces torch_cluster.radius_graph, Geom3D/models/schnet.py:91)
//   K1b extended graph   (replaces spspmm/coalesce in Geom3D/datasets/dataset_3D.py:12-35)
//   CSR-by-target view   (the accumulation order of MessagePassing.propagate)
// One warp owns one molecule (molecules are tiny and independent); rows of the boolean
// adjacency live in shared memory as 128-bit masks.
#include "common.cuh"

namespace molsde {

constexpr int kWarpsPerCta = 4;
constexpr int kMaxMol = MOLSDE_MAX_MOL_NODES;  // 128 -> 4 words per row
constexpr int kWords = kMaxMol / 32;

// ---------------------------------------------------------------------------------------
// segment_ptr: lower_bound per segment over an ascending key sequence
// ---------------------------------------------------------------------------------------
__global__ void segment_ptr_kernel(const int64_t* __restrict__ keys, const int64_t* __restrict__ indirect,
                                   int64_t M, int32_t num_segments, int32_t* __restrict__ ptr) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > num_segments) return;
    int64_t lo = 0, hi = M;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int64_t k = indirect ? keys[indirect[mid]] : keys[mid];
        if (k < s) lo = mid + 1; else hi = mid;
    }
    ptr[s] = static_cast<int32_t>(lo);
}

// ---------------------------------------------------------------------------------------
// exclusive scan (single CTA, 1024 threads, contiguous slice per thread)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ in, int64_t n,
                                                              int32_t* __restrict__ out) {
    __shared__ int32_t partial[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t b = t * per, e = min(n, b + per);
    int32_t s = 0;
    for (int64_t i = b; i < e; ++i) s += in[i];
    partial[t] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        int32_t v = (t >= off) ? partial[t - off] : 0;
        __syncthreads();
        partial[t] += v;
        __syncthreads();
    }
    int32_t run = partial[t] - s;
    for (int64_t i = b; i < e; ++i) { out[i] = run; run += in[i]; }
    if (t == 1023) out[n] = partial[1023];
}

// ---------------------------------------------------------------------------------------
// K1b: extended graph (<= 4-hop closure)
// ---------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
extend_graph_kernel(const int64_t* __restrict__ edge_index, int64_t E_b, const int32_t* __restrict__ node_ptr,
                    const int32_t* __restrict__ edge_ptr, int32_t B, int32_t* __restrict__ deg,
                    const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                    int64_t* __restrict__ ext_edge_index, int64_t E_x) {
    __shared__ uint32_t sA[kWarpsPerCta][kMaxMol][kWords];
    __shared__ uint32_t sR[kWarpsPerCta][kMaxMol][kWords];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kWarpsPerCta + w;
    if (g >= B) return;
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    const int e0 = edge_ptr[g], e1 = edge_ptr[g + 1];
    if (n > kMaxMol) return;  // host wrappers reject such batches before launching
    uint32_t(*A)[kWords] = sA[w];
    uint32_t(*R)[kWords] = sR[w];
    for (int i = lane; i < n * kWords; i += 32) (&A[0][0])[i] = 0u;
    __syncwarp();
    for (int e = e0 + lane; e < e1; e += 32) {
        int r = static_cast<int>(edge_index[e]) - n0;
        int c = static_cast<int>(edge_index[E_b + e]) - n0;
        atomicOr(&A[r][c >> 5], 1u << (c & 31));
    }
    __syncwarp();
    // R2[i] = A[i] | ((OR_{k in A[i]} A[k]) & ~bit_i)      dataset_3D.py:18-24
    for (int i = lane; i < n; i += 32) {
        uint32_t acc[kWords];
#pragma unroll
        for (int q = 0; q < kWords; ++q) acc[q] = 0u;
#pragma unroll
        for (int q = 0; q < kWords; ++q) {
            uint32_t m = A[i][q];
            while (m) {
                int k = (q << 5) + __ffs(m) - 1;
                m &= m - 1;
#pragma unroll
                for (int p = 0; p < kWords; ++p) acc[p] |= A[k][p];
            }
        }
        acc[i >> 5] &= ~(1u << (i & 31));
#pragma unroll
        for (int q = 0; q < kWords; ++q) R[i][q] = A[i][q] | acc[q];
    }
    __syncwarp();
    // R4[i] = R2[i] | ((OR_{k in R2[i]} R2[k]) & ~bit_i)   dataset_3D.py:28-34
    for (int i = lane; i < n; i += 32) {
        uint32_t acc[kWords];
#pragma unroll
        for (int q = 0; q < kWords; ++q) acc[q] = 0u;
#pragma unroll
        for (int q = 0; q < kWords; ++q) {
            uint32_t m = R[i][q];
            while (m) {
                int k = (q << 5) + __ffs(m) - 1;
                m &= m - 1;
#pragma unroll
                for (int p = 0; p < kWords; ++p) acc[p] |= R[k][p];
            }
        }
        acc[i >> 5] &= ~(1u << (i & 31));
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < kWords; ++q) { acc[q] |= R[i][q]; cnt += __popc(acc[q]); }
        if (!FILL) {
            deg[n0 + i] = cnt;
        } else {
            int o = rowptr[n0 + i];
#pragma unroll
            for (int q = 0; q < kWords; ++q) {
                uint32_t m = acc[q];
                while (m) {
                    int k = (q << 5) + __ffs(m) - 1;
                    m &= m - 1;
                    col[o] = n0 + k;
                    if (ext_edge_index) {
                        ext_edge_index[o] = n0 + i;
                        ext_edge_index[E_x + o] = n0 + k;
                    }
                    ++o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K1: radius graph
// ---------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
radius_graph_kernel(const float* __restrict__ pos, const int32_t* __restrict__ node_ptr, int32_t B, float r2,
                    int32_t cap, int32_t* __restrict__ deg, const int32_t* __restrict__ rowptr,
                    int32_t* __restrict__ col, int64_t* __restrict__ edge_index, int64_t E_r) {
    __shared__ float sp[kWarpsPerCta][kMaxMol * 3];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kWarpsPerCta + w;
    if (g >= B) return;
    const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
    float* P = sp[w];
    if (n > kMaxMol) return;  // host wrappers reject such batches before launching
    for (int i = lane; i < n * 3; i += 32) P[i] = pos[static_cast<int64_t>(n0) * 3 + i];
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const float xi = P[3 * i], yi = P[3 * i + 1], zi = P[3 * i + 2];
        int found = 0, kept = 0;
        int o = FILL ? rowptr[n0 + i] : 0;
        for (int j = 0; j < n && found < cap; ++j) {
            // fixed, FMA-free order ((dx*dx + dy*dy) + dz*dz): matches oracle/ref_ops.sq_dist_f32
            const float dx = __fsub_rn(xi, P[3 * j]), dy = __fsub_rn(yi, P[3 * j + 1]), dz = __fsub_rn(zi, P[3 * j + 2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 < r2) {
                ++found;
                if (j != i) {
                    if (FILL) {
                        col[o + kept] = n0 + j;
                        if (edge_index) {
                            edge_index[o + kept] = n0 + j;        // row 0: source
                            edge_index[E_r + o + kept] = n0 + i;  // row 1: target
                        }
                    }
                    ++kept;
                }
            }
        }
        if (!FILL) deg[n0 + i] = kept;
    }
}

// ---------------------------------------------------------------------------------------
// CSR-by-target of a generic edge list (stable in input order)
// ---------------------------------------------------------------------------------------
__global__ void csr_count_kernel(const int64_t* __restrict__ edge_index, int64_t E, int32_t* __restrict__ deg) {
    int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (e < E) atomicAdd(&deg[edge_index[E + e]], 1);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
csr_fill_kernel(const int64_t* __restrict__ edge_index, int64_t E, const int32_t* __restrict__ node_ptr,
                const int32_t* __restrict__ edge_ptr, int32_t B, const int32_t* __restrict__ rowptr,
                int32_t* __restrict__ src, int32_t* __restrict__ perm) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kWarpsPerCta + w;
    if (g >= B) return;
    const int n0 = node_ptr[g], n1 = node_ptr[g + 1];
    const int e0 = edge_ptr[g], e1 = edge_ptr[g + 1];
    for (int i = n0 + lane; i < n1; i += 32) {
        int o = rowptr[i];
        for (int e = e0; e < e1; ++e) {
            if (static_cast<int>(edge_index[E + e]) == i) {
                src[o] = static_cast<int32_t>(edge_index[e]);
                if (perm) perm[o] = e;
                ++o;
            }
        }
    }
}

}  // namespace molsde

using namespace molsde;

extern "C" {

int molsde_segment_ptr(const int64_t* keys, const int64_t* indirect, int64_t M, int32_t num_segments,
                       int32_t* ptr, void* stream) {
    if (!keys && M > 0) return MOLSDE_ERR_INVALID;
    if (!ptr || num_segments < 0) return MOLSDE_ERR_INVALID;
    int threads = 128, blocks = (num_segments + 1 + threads - 1) / threads;
    segment_ptr_kernel<<<blocks, threads, 0, as_stream(stream)>>>(keys, indirect, M, num_segments, ptr);
    return check_launch("segment_ptr");
}

int molsde_exclusive_scan_i32(const int32_t* counts, int64_t n, int32_t* out, void* stream) {
    if (!out || n < 0 || (n > 0 && !counts)) return MOLSDE_ERR_INVALID;
    exclusive_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(counts, n, out);
    return check_launch("exclusive_scan");
}

int molsde_extend_graph_count(const int64_t* edge_index, int64_t E_b, const int32_t* node_ptr,
                              const int32_t* edge_ptr, int32_t B, int32_t* deg, void* stream) {
    if (!node_ptr || !edge_ptr || !deg || B < 0) return MOLSDE_ERR_INVALID;
    if (B == 0) return MOLSDE_OK;
    int blocks = (B + kWarpsPerCta - 1) / kWarpsPerCta;
    extend_graph_kernel<false><<<blocks, kWarpsPerCta * 32, 0, as_stream(stream)>>>(
        edge_index, E_b, node_ptr, edge_ptr, B, deg, nullptr, nullptr, nullptr, 0);
    return check_launch("extend_graph_count");
}

int molsde_extend_graph_fill(const int64_t* edge_index, int64_t E_b, const int32_t* node_ptr,
                             const int32_t* edge_ptr, int32_t B, const int32_t* rowptr, int64_t E_x,
                             int32_t* col, int64_t* ext_edge_index, void* stream) {
    if (E_x == 0) return MOLSDE_OK;  // nothing to emit (e.g. a batch of isolated atoms)
    if (!node_ptr || !edge_ptr || !rowptr || !col || B < 0) return MOLSDE_ERR_INVALID;
    if (B == 0) return MOLSDE_OK;
    int blocks = (B + kWarpsPerCta - 1) / kWarpsPerCta;
    extend_graph_kernel<true><<<blocks, kWarpsPerCta * 32, 0, as_stream(stream)>>>(
        edge_index, E_b, node_ptr, edge_ptr, B, nullptr, rowptr, col, ext_edge_index, E_x);
    return check_launch("extend_graph_fill");
}

int molsde_radius_graph_count(const float* pos, const int32_t* node_ptr, int32_t B, float r,
                              int32_t max_num_neighbors, int32_t* deg, void* stream) {
    if (!pos || !node_ptr || !deg || B < 0) return MOLSDE_ERR_INVALID;
    if (B == 0) return MOLSDE_OK;
    int blocks = (B + kWarpsPerCta - 1) / kWarpsPerCta;
    const float r2 = static_cast<float>(static_cast<double>(r) * static_cast<double>(r));
    radius_graph_kernel<false><<<blocks, kWarpsPerCta * 32, 0, as_stream(stream)>>>(
        pos, node_ptr, B, r2, max_num_neighbors + 1, deg, nullptr, nullptr, nullptr, 0);
    return check_launch("radius_graph_count");
}

int molsde_radius_graph_fill(const float* pos, const int32_t* node_ptr, int32_t B, float r,
                             int32_t max_num_neighbors, const int32_t* rowptr, int64_t E_r, int32_t* col,
                             int64_t* edge_index, void* stream) {
    if (E_r == 0) return MOLSDE_OK;
    if (!pos || !node_ptr || !rowptr || !col || B < 0) return MOLSDE_ERR_INVALID;
    if (B == 0) return MOLSDE_OK;
    int blocks = (B + kWarpsPerCta - 1) / kWarpsPerCta;
    const float r2 = static_cast<float>(static_cast<double>(r) * static_cast<double>(r));
    radius_graph_kernel<true><<<blocks, kWarpsPerCta * 32, 0, as_stream(stream)>>>(
        pos, node_ptr, B, r2, max_num_neighbors + 1, nullptr, rowptr, col, edge_index, E_r);
    return check_launch("radius_graph_fill");
}

int molsde_csr_by_target_count(const int64_t* edge_index, int64_t E, int64_t N, int32_t* deg, void* stream) {
    if (!deg || E < 0 || (E > 0 && !edge_index)) return MOLSDE_ERR_INVALID;
    cudaError_t err = cudaMemsetAsync(deg, 0, sizeof(int32_t) * N, as_stream(stream));
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    if (E == 0) return MOLSDE_OK;
    int threads = 256;
    int64_t blocks = (E + threads - 1) / threads;
    csr_count_kernel<<<static_cast<unsigned>(blocks), threads, 0, as_stream(stream)>>>(edge_index, E, deg);
    return check_launch("csr_count");
}

int molsde_csr_by_target_fill(const int64_t* edge_index, int64_t E, const int32_t* node_ptr,
                              const int32_t* edge_ptr, int32_t B, const int32_t* rowptr, int32_t* src,
                              int32_t* perm, void* stream) {
    if (B == 0 || E == 0) return MOLSDE_OK;
    if (!node_ptr || !edge_ptr || !rowptr || !src || B < 0) return MOLSDE_ERR_INVALID;
    int blocks = (B + kWarpsPerCta - 1) / kWarpsPerCta;
    csr_fill_kernel<<<blocks, kWarpsPerCta * 32, 0, as_stream(stream)>>>(edge_index, E, node_ptr, edge_ptr, B,
                                                                       rowptr, src, perm);
    return check_launch("csr_fill");
}

}  // extern "C"
