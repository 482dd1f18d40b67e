// Library plumbing: version, error reporting, device check.
#include <string.h>

#include "common.cuh"

namespace molsde {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* msg) {
    strncpy(g_last_error, msg ? msg : "", sizeof(g_last_error) - 1);
    g_last_error[sizeof(g_last_error) - 1] = 0;
}

int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) return MOLSDE_OK;
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(err));
    set_last_error(buf);
    return MOLSDE_ERR_CUDA;
}

}  // namespace molsde

extern "C" {

const char* molsde_version(void) { return "molsde_b200 0.1.0 (sm_100a)"; }

const char* molsde_last_error_string(void) { return molsde::g_last_error; }

int molsde_check_device(int device) {
    cudaDeviceProp prop;
    cudaError_t err = cudaGetDeviceProperties(&prop, device);
    if (err != cudaSuccess) {
        molsde::set_last_error(cudaGetErrorString(err));
        return MOLSDE_ERR_CUDA;
    }
    if (prop.major != 10) {
        char buf[128];
        snprintf(buf, sizeof(buf), "device %d is sm_%d%d; molsde_b200 is built for sm_100a only", device, prop.major,
                 prop.minor);
        molsde::set_last_error(buf);
        return MOLSDE_ERR_UNSUPPORTED;
    }
    return MOLSDE_OK;
}

/* Test hook for the calling-convention shortcut of csrc/fastcall.c: echoes its (deliberately interleaved, stack-spilling) arguments. */
int molsde_debug_echo(int64_t a, float x, int32_t b, const void* p, float y, int64_t c, int32_t d, uint64_t e, int64_t f, float z,
                      int64_t g, int32_t h, double* out) {
    if (!out) return MOLSDE_ERR_INVALID;
    out[0] = static_cast<double>(a); out[1] = x; out[2] = b; out[3] = static_cast<double>(reinterpret_cast<uintptr_t>(p)); out[4] = y;
    out[5] = static_cast<double>(c); out[6] = d; out[7] = static_cast<double>(e); out[8] = static_cast<double>(f); out[9] = z;
    out[10] = static_cast<double>(g); out[11] = h;
    return 7;
}

/* Host-side chunk / tile plan of a batch (the bookkeeping behind molsde_plan): chunks = runs of whole molecules (<= max_nodes atoms,
 * greedy under an edge budget of 0.6 * max_tiles * tile_edges) or the caller's fixed groups; tiles = runs of whole target nodes with
 * <= tile_edges incoming edges (greedy, <= tile_edges targets).  Plain sequential C over HOST arrays: the Python loop it replaces cost
 * ~15 ms for a 1 M-edge batch.  Outputs: chunk_tile_ptr [num_chunks + 1], tile_tgt_ptr [num_tiles + 1] (capacities: B + 2 / N + 2);
 * counts_out = {num_chunks, num_tiles, max_chunk_tiles, offending index}.  Returns MOLSDE_OK, or MOLSDE_ERR_UNSUPPORTED with
 * counts_out[3] = the node with more than tile_edges incoming edges (code -2), the chunk beyond max_nodes (-3) or max_tiles (-4)
 * in counts_out[2]. */
int molsde_build_plan_host(const int64_t* rowptr, const int64_t* node_ptr, int32_t B, const int64_t* groups, int32_t G, int32_t tile_edges,
                           int32_t max_nodes, int32_t max_tiles, int32_t* chunk_tile_ptr, int32_t* tile_tgt_ptr, int64_t* counts_out) {
    if (!rowptr || !node_ptr || !chunk_tile_ptr || !tile_tgt_ptr || !counts_out || B < 0) return MOLSDE_ERR_INVALID;
    const int64_t N = node_ptr[B];
    int64_t nchunks = 0, ntiles = 0, max_ct = 0;
    chunk_tile_ptr[0] = 0;
    const int64_t edge_budget = static_cast<int64_t>(static_cast<double>(max_tiles) * tile_edges * 0.6);
    int32_t m = 0, gi = 0;
    while (groups ? gi < G : m < B) {
        int64_t a, b;
        if (groups) {
            a = node_ptr[groups[gi]]; b = node_ptr[groups[gi + 1]]; ++gi;
        } else {
            int32_t end = m;
            while (end < B && node_ptr[end + 1] - node_ptr[m] <= max_nodes &&
                   rowptr[node_ptr[end + 1]] - rowptr[node_ptr[m]] <= edge_budget) ++end;
            if (end == m) end = m + 1;   // a single big molecule: the exact checks below decide
            a = node_ptr[m]; b = node_ptr[end]; m = end;
        }
        if (b - a > max_nodes) { counts_out[2] = -3; counts_out[3] = nchunks; return MOLSDE_ERR_UNSUPPORTED; }
        int64_t ct = 0, i = a;
        while (i < b) {
            tile_tgt_ptr[ntiles++] = static_cast<int32_t>(i);
            ++ct;
            // largest j with rowptr[j] <= rowptr[i] + tile_edges  (rowptr ascending)
            int64_t lo = i, hi = N;   // invariant: rowptr[lo] <= limit
            const int64_t limit = rowptr[i] + tile_edges;
            while (hi - lo > 0) {
                const int64_t mid = lo + (hi - lo + 1) / 2;
                if (rowptr[mid] <= limit) lo = mid; else hi = mid - 1;
            }
            if (lo <= i) { counts_out[2] = -2; counts_out[3] = i; return MOLSDE_ERR_UNSUPPORTED; }
            int64_t nx = lo < b ? lo : b;
            if (i + tile_edges < nx) nx = i + tile_edges;
            i = nx;
        }
        if (ct > max_tiles) { counts_out[2] = -4; counts_out[3] = nchunks; counts_out[1] = ct; return MOLSDE_ERR_UNSUPPORTED; }
        if (ct > max_ct) max_ct = ct;
        chunk_tile_ptr[++nchunks] = static_cast<int32_t>(ntiles);
    }
    tile_tgt_ptr[ntiles] = static_cast<int32_t>(N);
    counts_out[0] = nchunks; counts_out[1] = ntiles; counts_out[2] = max_ct; counts_out[3] = 0;
    return MOLSDE_OK;
}

}  // extern "C"
