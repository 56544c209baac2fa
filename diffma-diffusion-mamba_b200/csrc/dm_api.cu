// Library-level entry points of the diffma_b200 C-ABI (include/diffma_b200.h): version, status strings.
#include "dm_common.cuh"

namespace dm {
thread_local int g_last_cuda_error = 0;
}

extern "C" {

int dm_version(void) { return DM_ABI_VERSION; }

const char* dm_status_string(int status) {
    switch (status) {
        case DM_OK: return "ok";
        case DM_ERR_INVALID_ARG: return "invalid argument (null pointer, bad size, misaligned pointer or stride)";
        case DM_ERR_UNSUPPORTED: return "unsupported shape / dtype for this build";
        case DM_ERR_CUDA: return "CUDA runtime error (see dm_last_cuda_error)";
        default: return "unknown status";
    }
}

int dm_last_cuda_error(void) { return dm::g_last_cuda_error; }

#define DM_STR2(x) #x
#define DM_STR(x) DM_STR2(x)
const char* dm_build_info(void) {
    return "diffma_b200 abi " DM_STR(DM_ABI_VERSION) " sm_100a nvcc " DM_STR(__CUDACC_VER_MAJOR__) "." DM_STR(
        __CUDACC_VER_MINOR__) "." DM_STR(__CUDACC_VER_BUILD__) " built " __DATE__;
}

}  // extern "C"
