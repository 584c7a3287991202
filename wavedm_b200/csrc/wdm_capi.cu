// wdm_capi.cu -- library identification + status strings of the C ABI (include/wavedm_b200.h).
#include "wdm_common.cuh"

extern "C" int wdm_version(void) { return 100 * 0 + 1; }

extern "C" const char* wdm_build_arch(void) { return "sm_100a"; }

extern "C" const char* wdm_status_string(int status) {
    switch (status) {
        case WDM_OK: return "ok";
        case WDM_ERR_BAD_SHAPE: return "unsupported shape";
        case WDM_ERR_BAD_ALIGN: return "pointer not 16-byte aligned";
        case WDM_ERR_BAD_ARG: return "bad argument";
        case WDM_ERR_UNSUPPORTED: return "not supported by this engine";
        case WDM_ERR_WORKSPACE: return "workspace too small";
        case WDM_ERR_NO_DEVICE: return "no sm_100 device";
        default: break;
    }
    if (status <= WDM_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(WDM_ERR_CUDA_BASE - status));
    return "unknown status";
}

#include <atomic>
static std::atomic<long long> g_launches{0};
extern "C" long long wdm_launch_counter_add(long long n) { return g_launches.fetch_add(n) + n; }
extern "C" long long wdm_launch_counter(void) { return g_launches.load(); }
