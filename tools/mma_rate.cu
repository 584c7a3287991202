// Micro-benchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (M=128, N in {128,192,256}, K=16) on operands that
// already sit in shared memory (no TMA in the loop). Prints cycles per MMA instruction per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../wavedm_b200/csrc/wdm_ptx.cuh"
using namespace wdm;

__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out, int kblocks_distinct) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1 && lane == 0) {
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t sa = ptx::smem_u32(smem + (it % kblocks_distinct) * 49152);
            const uint64_t da = make_desc(sa), db = make_desc(sa + 16384);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ptx::umma_f16_ss(tm, da + 2 * kk, db + 2 * kk, idesc(128, N), (it | kk) ? 1u : 0u);
        }
        ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N>
void run(const char* name, int grid) {
    long long* d; cudaMalloc(&d, 8);
    const int smem = 4 * 49152 + 1024;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int kd : {1, 4}) {
        k<N><<<grid, 128, smem>>>(iters, d, kd);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%s grid=%d distinct_stages=%d: %.1f cycles per MMA (M=128,N=%d,K=16)  err=%s\n", name, grid, kd,
               (double)h / (iters * 4.0), N, cudaGetErrorString(e));
    }
    cudaFree(d);
}

// The real kernel's issue pattern: every G MMAs the issuer waits on an (already complete) mbarrier, fences and commits.
template <int N, int G>
__global__ void __launch_bounds__(128, 1) k2(int groups, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, ready, sink[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1); ptx::mbar_init(&ready, 1);
        for (int i = 0; i < 8; ++i) ptx::mbar_init(&sink[i], 1);
        ptx::fence_mbar_init();
        ptx::mbar_arrive(&ready);  // phase 0 complete: wait(parity 0) returns at once
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            ptx::mbar_wait(&ready, 0);
            ptx::tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < G / 4; ++j) {
                    const uint32_t sa = ptx::smem_u32(smem + ((g * (G / 4) + j) % 4) * 49152);
                    const uint64_t da = make_desc(sa), db = make_desc(sa + 16384);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) ptx::umma_f16_ss(tm, da + 2 * kk, db + 2 * kk, idesc(128, N), (g | j | kk) ? 1u : 0u);
                    ptx::umma_commit(&sink[(g * (G / 4) + j) % 8]);
                }
            }
            __syncwarp();
        }
        if (lane == 0) { ptx::umma_commit(&bar); }
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N, int G>
void run2(int grid) {
    long long* d; cudaMalloc(&d, 8);
    const int smem = 4 * 49152 + 1024;
    cudaFuncSetAttribute(k2<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int groups = 8000 / G;
    k2<N, G><<<grid, 128, smem>>>(groups, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("wait+fence+commit every %2d MMAs, N=%d, grid=%d: %.1f cycles per MMA  err=%s\n", G, N, grid,
           (double)h / (groups * (double)G), cudaGetErrorString(e));
    cudaFree(d);
}

// Cost of the pieces of the issue-group protocol for ONE thread (the MMA issuer): try_wait on an already-complete
// mbarrier, tcgen05.fence::after_thread_sync, tcgen05.commit (no MMAs pending), and a wait+fence+commit round.
__global__ void __launch_bounds__(128, 1) k3(int iters, long long* out) {
    __shared__ uint64_t ready, sink[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { ptx::tmem_alloc(&slot, 64); ptx::tmem_relinquish(); }
    if (threadIdx.x == 0) {
        ptx::mbar_init(&ready, 1);
        for (int i = 0; i < 8; ++i) ptx::mbar_init(&sink[i], 1);
        ptx::fence_mbar_init();
        ptx::mbar_arrive(&ready);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1 && lane == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) ptx::mbar_wait(&ready, 0);
        long long t1 = clock64();
        for (int i = 0; i < iters; ++i) ptx::tc_fence_after();
        long long t2 = clock64();
        for (int i = 0; i < iters; ++i) ptx::umma_commit(&sink[i & 7]);
        long long t3 = clock64();
        for (int i = 0; i < iters; ++i) {
            ptx::mbar_wait(&ready, 0);
            ptx::mbar_wait(&ready, 0);
            ptx::tc_fence_after();
            ptx::umma_commit(&sink[i & 7]);
            ptx::umma_commit(&sink[(i + 1) & 7]);
        }
        long long t4 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t1, out[2] = t3 - t2, out[3] = t4 - t3;
    }
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(slot, 64);
}

void run3() {
    long long* d; cudaMalloc(&d, 32);
    const int iters = 2000;
    k3<<<148, 128>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4] = {0, 0, 0, 0}; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("issuer protocol pieces (cycles each, one thread): try_wait(complete) %.1f, fence::after_thread_sync %.1f, "
           "tcgen05.commit %.1f, [2 waits + fence + 2 commits] %.1f  err=%s\n", h[0] / (double)iters, h[1] / (double)iters,
           h[2] / (double)iters, h[3] / (double)iters, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run3();
    run2<256, 4>(148); run2<256, 8>(148); run2<256, 16>(148);
    run2<128, 4>(148); run2<128, 8>(148); run2<128, 16>(148);
    run<256>("cg1", 1); run<256>("cg1", 148);
    run<192>("cg1", 148);
    run<128>("cg1", 148);
    run<64>("cg1", 148);
    return 0;
}
