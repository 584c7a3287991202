/*
 * wavedm_b200.h -- C ABI of libwavedm_b200.so (hand-written sm_100a kernels for WaveDM's sampling hot path).
 *
 * The reference (Easquel/WaveDM) is pure Python/PyTorch and has NO FFI / plugin boundary of its own
 * (SURVEY.md 8b): every entry point below is net-new and replaces a *PyTorch call site* of the reference,
 * cited per function as reference file:line (paths relative to the reference checkout). The Python host
 * side (wavedm_b200/*.py) mirrors the reference's class API and binds these symbols with ctypes; the
 * stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every data pointer is DEVICE memory valid on `stream`;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - returns WDM_OK (0) or a negative wdm_status; never throws, never synchronises, never frees
 *     caller memory; device scratch is passed in (`workspace`, sized by the matching *_workspace_bytes);
 *   - re-entrant for distinct streams / distinct handles;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns a CUDA error.
 */
#ifndef WAVEDM_B200_H_
#define WAVEDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define WDM_API __attribute__((visibility("default")))
#else
#define WDM_API
#endif

typedef enum wdm_status {
    WDM_OK = 0,
    WDM_ERR_BAD_SHAPE = -1,   /* dimension not supported (e.g. H or W not a multiple of 4) */
    WDM_ERR_BAD_ALIGN = -2,   /* pointer not aligned as required (16 B) */
    WDM_ERR_BAD_ARG = -3,     /* null pointer / unknown flag / inconsistent arguments */
    WDM_ERR_UNSUPPORTED = -4, /* valid request the engine does not implement */
    WDM_ERR_WORKSPACE = -5,   /* workspace too small */
    WDM_ERR_NO_DEVICE = -6,   /* device is not sm_100 */
    WDM_ERR_CUDA_BASE = -1000 /* CUDA runtime error e is returned as WDM_ERR_CUDA_BASE - e */
} wdm_status;

/* Library / build identification. */
WDM_API int wdm_version(void);               /* 100*major + minor */
WDM_API const char* wdm_build_arch(void);    /* "sm_100a" */
WDM_API const char* wdm_status_string(int status);

/* ------------------------------------------------------------------------------------------------
 * 2-level Haar-packet ("c2", scale=2) wavelet transform.
 * Replaces  models/wavelet.py:36-43  (WaveletTransform.forward, dec=True: Conv2d(3,48,k=4,s=4,groups=3)
 * with the fixed rec4 weights + view/transpose/contiguous permute) and  models/wavelet.py:44-49
 * (dec=False: inverse permute + ConvTranspose2d). One kernel each, no intermediate tensor.
 *   x : [n, 3, H, W]      fp32 NCHW contiguous, H % 4 == W % 4 == 0
 *   y : [n, 48, H/4, W/4] fp32 NCHW contiguous, channel = 3*k + colour  (sub-band-major)
 * flags:
 *   WDM_DWT_PRE_2XM1    apply data_transform 2x-1 on load        (models/restoration.py:8-9)
 *   WDM_IWT_POST_CLAMP  apply clamp((x+1)/2, 0, 1) on store      (models/restoration.py:12-13)
 *   WDM_WT_IMPL_*       force a kernel variant (testing / benchmarking); default AUTO
 * ------------------------------------------------------------------------------------------------ */
#define WDM_DWT_PRE_2XM1 0x1
#define WDM_IWT_POST_CLAMP 0x1
#define WDM_WT_IMPL_AUTO 0x00
#define WDM_WT_IMPL_DIRECT 0x10 /* register-only kernel: vector loads, coalesced stores */
#define WDM_WT_IMPL_TMA 0x20    /* TMA-staged shared-memory tiles (needs W % 16 == 0) */
#define WDM_WT_IMPL_MASK 0xF0

WDM_API int wdm_dwt4x4_fwd(const float* x, float* y, int n, int H, int W, int flags, void* stream);
WDM_API int wdm_iwt4x4_fwd(const float* y, float* x, int n, int h, int w, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVEDM_B200_H_ */
