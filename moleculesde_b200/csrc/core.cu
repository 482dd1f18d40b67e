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

}  // extern "C"
