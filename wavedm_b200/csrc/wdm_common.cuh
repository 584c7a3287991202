// wdm_common.cuh -- shared helpers for the wavedm_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/wavedm_b200.h"

#define WDM_CHECK_CUDA(expr)                                   \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return wdm_cuda_error((int)_e); \
    } while (0)

// CUDA errors are passed through as WDM_ERR_CUDA_BASE - cudaError (always < WDM_ERR_CUDA_BASE).
static inline int wdm_cuda_error(int e) { return WDM_ERR_CUDA_BASE - e; }

// every kernel launcher ends with wdm_launch_status(): it also feeds the launch counter (wdm_launch_counter)
extern "C" long long wdm_launch_counter_add(long long n);
static inline int wdm_launch_status() {
    wdm_launch_counter_add(1);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return wdm_cuda_error((int)e);
    }
    return WDM_OK;
}

#ifdef __CUDACC__
// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains (its prologue
// -- barrier init, TMEM allocation, descriptor prefetch -- overlaps the predecessor's tail); it must execute
// wdm_grid_dependency_wait() before touching anything the predecessor wrote.
template <typename... KArgs, typename... Args>
static inline cudaError_t wdm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// Called first by every PDL kernel: lets the NEXT kernel's CTAs be scheduled as soon as SM resources allow (they park
// in wdm_grid_dependency_wait() until this grid has completed). The trigger fires once all CTAs of this grid have
// executed it, i.e. when the whole grid is resident -- later waves are never starved by parked dependents.
__device__ __forceinline__ void wdm_grid_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void wdm_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

static inline bool wdm_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

static inline int wdm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
__device__ __forceinline__ float4 wdm_ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float wdm_ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void wdm_stg_stream(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void wdm_stg_stream(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// x * sigmoid(x) = 0.5 x (1 + tanh(x / 2)): one MUFU op (tanh.approx, rel. err ~2^-11, below bf16 resolution) instead of
// two (ex2 + rcp) -- the bf16 GroupNorm+SiLU pass is MUFU-throughput bound otherwise.
__device__ __forceinline__ float wdm_silu(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
#endif
