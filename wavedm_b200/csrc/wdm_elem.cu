// wdm_elem.cu -- the HBM-bound / small kernels around the contractions:
//   GroupNorm statistics + normalise/affine/SiLU (models/unet.py:36-37,31-33), row softmax (unet.py:182),
//   nearest x2 upsample (unet.py:52-53), timestep embedding + temb projections (unet.py:10-28,354-357,125),
//   patch gather (ddm_wavelet.py:467-478), fused overlap-average + DDIM update (ddm_wavelet.py:485-503),
//   weight packing.
#include <stdlib.h>

#include "wdm_common.cuh"
#include "wdm_engine.h"

namespace wdm {
namespace {

// ---------------------------------------------------------------------------------------------- helpers
template <typename T, int V>
struct Vec;
template <>
struct Vec<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Vec<float, 8> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
template <>
struct Vec<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        uint2 q = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(q.x << 16), v[1] = __uint_as_float(q.x & 0xffff0000u);
        v[2] = __uint_as_float(q.y << 16), v[3] = __uint_as_float(q.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        uint2 q;
        q.x = pack_bf16(v[0], v[1]), q.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = q;
    }
};
template <>
struct Vec<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 q = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[2 * i] = __uint_as_float(w[i] << 16), v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 q;
        q.x = pack_bf16(v[0], v[1]), q.y = pack_bf16(v[2], v[3]), q.z = pack_bf16(v[4], v[5]), q.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = q;
    }
};

template <typename T>
struct AccT {
    typedef double type;  // fp32 engine: double statistics (sum / sum of squares without cancellation)
};
template <>
struct AccT<__nv_bfloat16> {
    typedef float type;
};

__device__ __forceinline__ float silu_precise(float x) { return x * (1.0f / (1.0f + expf(-x))); }

constexpr int kMaxSlabs = 32;

// ---------------------------------------------------------------------------------------------- GroupNorm
// grid (S, P); block = ppi * nvec threads (thread = fixed V-channel vector, ppi pixels in flight).
// partial[p][s][g] = (sum, sumsq) in double.
template <typename T, int V>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ s0, int C0, const T* __restrict__ s1,
                                                       int C1, int HW, int S, double2* __restrict__ partial) {
    __shared__ double red[32][2];
    const int C = C0 + C1, nvec = C / V, cpg = C / 32;
    const int p = blockIdx.y, s = blockIdx.x;
    const int vec = threadIdx.x % nvec, lane_pix = threadIdx.x / nvec, ppi = blockDim.x / nvec;
    const int c = vec * V, g = c / cpg;
    if (threadIdx.x < 64) red[threadIdx.x >> 1][threadIdx.x & 1] = 0.0;
    __syncthreads();
    const int pix_per_slab = HW / S;
    const int px0 = s * pix_per_slab, px1 = px0 + pix_per_slab;
    const T* base;
    int ld, co;
    if (c < C0)
        base = s0, ld = C0, co = c;
    else
        base = s1, ld = C1, co = c - C0;
    base += (long long)p * HW * ld + co;
    typename AccT<T>::type sum = 0, sq = 0;
    for (int px = px0 + lane_pix; px < px1; px += ppi) {
        float v[V];
        Vec<T, V>::load(base + (long long)px * ld, v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            sum += v[i];
            sq += (typename AccT<T>::type)v[i] * v[i];
        }
    }
    atomicAdd(&red[g][0], (double)sum);
    atomicAdd(&red[g][1], (double)sq);
    __syncthreads();
    if (threadIdx.x < 32) partial[((long long)p * S + s) * 32 + threadIdx.x] = make_double2(red[threadIdx.x][0], red[threadIdx.x][1]);
}

// Finalise partial sums -> (mean, rstd) per (p, g). One thread per (p, g).
__global__ void gn_finalize_kernel(const double2* __restrict__ partial, int S, int n_pg, double inv_n, double eps,
                                   float* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pg) return;
    const int p = i / 32, g = i % 32;
    double sum = 0, sq = 0;
    for (int s = 0; s < S; ++s) {
        double2 t = partial[((long long)p * S + s) * 32 + g];
        sum += t.x, sq += t.y;
    }
    const double mean = sum * inv_n;
    double var = sq * inv_n - mean * mean;
    if (var < 0) var = 0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

template <typename T, int V, bool kSilu, bool kPrecise>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ s0, int C0, const T* __restrict__ s1,
                                                       int C1, int HW, int pix_per_cta,
                                                       const float* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, T* __restrict__ out) {
    wdm_grid_launch_dependents();
    wdm_grid_dependency_wait();
    const int C = C0 + C1, nvec = C / V, cpg = C / 32;
    const int p = blockIdx.y;
    const int vec = threadIdx.x % nvec, lane_pix = threadIdx.x / nvec, ppi = blockDim.x / nvec;
    const int c = vec * V;
    float a[V], b[V];
    {
        // V = 8 channels with C/32 a multiple of 4: the vector touches at most two groups (first / last 4 channels)
        const int g0 = c / cpg, g1 = (c + V - 1) / cpg;
        const float2 st0 = *reinterpret_cast<const float2*>(stats + (p * 32 + g0) * 2);
        const float2 st1 = *reinterpret_cast<const float2*>(stats + (p * 32 + g1) * 2);
        float ga[V], be[V];
#pragma unroll
        for (int i = 0; i < V; i += 4) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c + i));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c + i));
            ga[i] = g4.x, ga[i + 1] = g4.y, ga[i + 2] = g4.z, ga[i + 3] = g4.w;
            be[i] = b4.x, be[i + 1] = b4.y, be[i + 2] = b4.z, be[i + 3] = b4.w;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const bool second = (c + i) / cpg != g0;
            const float mean = second ? st1.x : st0.x, rstd = second ? st1.y : st0.y;
            a[i] = rstd * ga[i];
            b[i] = be[i] - mean * a[i];
        }
    }
    const T* base;
    int ld, co;
    if (c < C0)
        base = s0, ld = C0, co = c;
    else
        base = s1, ld = C1, co = c - C0;
    base += (long long)p * HW * ld + co;
    T* o = out + (long long)p * HW * C + c;
    const int px0 = blockIdx.x * pix_per_cta;
    const int px1 = min(HW, px0 + pix_per_cta);
    // All of this thread's loads are issued before the first use (raw 16-byte registers, converted on the fly), so one
    // CTA pass keeps U x 16 B per thread in flight; the launcher sizes CTAs to exactly one such pass.
    constexpr int U = 8;
    constexpr int R = sizeof(T) * V / 16;  // 16-byte registers per vector (1 for bf16, 2 for fp32)
    uint4 raw[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int px = px0 + lane_pix + u * ppi;
        if (px < px1) {
#pragma unroll
            for (int r = 0; r < R; ++r) raw[u][r] = __ldg(reinterpret_cast<const uint4*>(base + (long long)px * ld) + r);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int px = px0 + lane_pix + u * ppi;
        if (px < px1) {
            float v[V];
            if (sizeof(T) == 2) {
                const uint32_t w[4] = {raw[u][0].x, raw[u][0].y, raw[u][0].z, raw[u][0].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) v[2 * i] = __uint_as_float(w[i] << 16), v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    v[4 * r] = __uint_as_float(raw[u][r].x), v[4 * r + 1] = __uint_as_float(raw[u][r].y);
                    v[4 * r + 2] = __uint_as_float(raw[u][r].z), v[4 * r + 3] = __uint_as_float(raw[u][r].w);
                }
            }
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float y = fmaf(v[i], a[i], b[i]);
                if (kSilu) y = kPrecise ? silu_precise(y) : wdm_silu(y);
                v[i] = y;
            }
            Vec<T, V>::store(o + (long long)px * C, v);
        }
    }
}

// (mean, rstd) per (patch, group) from the side-car partial sums the tensor-core epilogue wrote next to the
// tensor(s): sc[M/32][C/4][2]. One warp per (patch, group); double accumulation.
__global__ void __launch_bounds__(256) gn_finalize_sidecar_kernel(const float* __restrict__ sc0, int C0,
                                                                  const float* __restrict__ sc1, int C1, int HW,
                                                                  int P, double eps, float* __restrict__ stats) {
    wdm_grid_launch_dependents();
    wdm_grid_dependency_wait();
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= P * 32) return;
    const int lane = threadIdx.x & 31;
    const int p = wid >> 5, g = wid & 31;
    const int cpg = (C0 + C1) / 32, nb = cpg >> 2, fb = (g * cpg) >> 2, nrg = HW >> 5;
    const int b0n = C0 >> 2, b1n = C1 >> 2;
    double sum = 0, sq = 0;
    for (int it = lane; it < nrg * nb; it += 32) {
        const int rgi = it / nb, b = fb + (it - rgi * nb);
        const float2 t = b < b0n
            ? *reinterpret_cast<const float2*>(sc0 + (((long long)p * nrg + rgi) * b0n + b) * 2)
            : *reinterpret_cast<const float2*>(sc1 + (((long long)p * nrg + rgi) * b1n + (b - b0n)) * 2);
        sum += t.x, sq += t.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
        const double inv_n = 1.0 / ((double)HW * cpg);
        const double mean = sum * inv_n;
        double var = sq * inv_n - mean * mean;
        if (var < 0) var = 0;
        stats[2 * wid] = (float)mean;
        stats[2 * wid + 1] = (float)(1.0 / sqrt(var + eps));
    }
}


// GroupNorm normalise (+SiLU) with the statistics FINALISED IN THE KERNEL from the side-car partial sums of the producing
// tensor-core kernel(s) (sc[M/32][C/4][2], see gn_finalize_sidecar_kernel): one launch instead of two per GroupNorm -- the
// separate finalize launch was ~4 us of pure dependency latency, 51 times per UNet call. Every CTA of a patch reduces that
// patch's side-car (<= 96 KiB, L2-resident: the producer wrote it microseconds ago) in a fixed order, so all CTAs obtain
// bit-identical (mean, rstd); a CTA then makes `passes` passes of U x ppi pixels so the prologue is amortised.
// Prologue mapping: column = 4-channel block b, `ngrp` thread groups split the 32-row groups; coalesced float2 reads,
// double accumulation, a fixed-order second stage per GroupNorm group.
// `reverse`: CTAs walk patches / pixel chunks from the END of the tensor: the producing kernel wrote the tensor front to
// back, so its tail is what the L2 still holds (and this kernel's own output is then consumed front to back, again most
// recently written first).
template <bool kSilu>
__global__ void __launch_bounds__(256, 3) gn_apply_sc_kernel(const __nv_bfloat16* __restrict__ s0, int C0, const float* __restrict__ sc0,
                                                          const __nv_bfloat16* __restrict__ s1, int C1, const float* __restrict__ sc1,
                                                          int HW, int pix_per_cta, double eps, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                          int reverse) {
    constexpr int V = 8;
    __shared__ double2 part[384];
    __shared__ float2 s_stat[32];
    wdm_grid_launch_dependents();
    const int C = C0 + C1, nvec = C / V, cpg = C / 32;
    const int p = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    const int chunk = reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const int T = blockDim.x, tid = threadIdx.x;
    const int vec = tid % nvec, lane_pix = tid / nvec, ppi = T / nvec;
    const int c = vec * V;
    float ga[V], be[V];
#pragma unroll
    for (int i = 0; i < V; i += 4) {   // parameters do not depend on the predecessor: load before the dependency wait
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c + i));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c + i));
        ga[i] = g4.x, ga[i + 1] = g4.y, ga[i + 2] = g4.z, ga[i + 3] = g4.w;
        be[i] = b4.x, be[i + 1] = b4.y, be[i + 2] = b4.z, be[i + 3] = b4.w;
    }
    wdm_grid_dependency_wait();
    {
        const int nblk = C >> 2, nb = cpg >> 2, nrg = HW >> 5, b0n = C0 >> 2, b1n = C1 >> 2;
        const int ngrp = T >= nblk ? T / nblk : 1;
        const int ncol = T >= nblk ? 1 : (nblk + T - 1) / T;   // columns per thread when the CTA is narrower than the side-car
        for (int j = 0; j < ncol; ++j) {
            const int b = T >= nblk ? tid % nblk : tid + j * T;
            const int sub = T >= nblk ? tid / nblk : 0;
            if (b < nblk && sub < ngrp) {
                const float* base = b < b0n ? sc0 + ((long long)p * nrg * b0n + b) * 2 : sc1 + ((long long)p * nrg * b1n + (b - b0n)) * 2;
                const long long pitch = (b < b0n ? b0n : b1n) * 2;
                double sum = 0, sq = 0;
                int rg = sub;
                for (; rg + 7 * ngrp < nrg; rg += 8 * ngrp) {
                    float2 t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = *reinterpret_cast<const float2*>(base + (long long)(rg + u * ngrp) * pitch);
#pragma unroll
                    for (int u = 0; u < 8; ++u) sum += t[u].x, sq += t[u].y;
                }
                for (; rg < nrg; rg += ngrp) {
                    const float2 t = *reinterpret_cast<const float2*>(base + (long long)rg * pitch);
                    sum += t.x, sq += t.y;
                }
                part[sub * nblk + b] = make_double2(sum, sq);
            }
        }
        __syncthreads();
        if (tid < 32) {
            double sum = 0, sq = 0;
            for (int sub = 0; sub < ngrp; ++sub)
                for (int bi = 0; bi < nb; ++bi) {
                    const double2 t = part[sub * nblk + tid * nb + bi];
                    sum += t.x, sq += t.y;
                }
            const double inv_n = 1.0 / ((double)HW * cpg);
            const double mean = sum * inv_n;
            double var = sq * inv_n - mean * mean;
            if (var < 0) var = 0;
            s_stat[tid] = make_float2((float)mean, (float)(1.0 / sqrt(var + eps)));
        }
        __syncthreads();
    }
    float a[V], b[V];
    {
        const int g0 = c / cpg, g1 = (c + V - 1) / cpg;
        const float2 st0 = s_stat[g0], st1 = s_stat[g1];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const bool second = (c + i) / cpg != g0;
            const float mean = second ? st1.x : st0.x, rstd = second ? st1.y : st0.y;
            a[i] = rstd * ga[i];
            b[i] = be[i] - mean * a[i];
        }
    }
    const __nv_bfloat16* base;
    int ld, co;
    if (c < C0)
        base = s0, ld = C0, co = c;
    else
        base = s1, ld = C1, co = c - C0;
    base += (long long)p * HW * ld + co;
    __nv_bfloat16* o = out + (long long)p * HW * C + c;
    const int px_begin = chunk * pix_per_cta;
    const int px_end = min(HW, px_begin + pix_per_cta);
    constexpr int U = 8;
    for (int px0 = px_begin; px0 < px_end; px0 += U * ppi) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = px0 + lane_pix + u * ppi;
            if (px < px_end) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + (long long)px * ld));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = px0 + lane_pix + u * ppi;
            if (px < px_end) {
                float v[V];
                const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) v[2 * i] = __uint_as_float(w[i] << 16), v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float y = fmaf(v[i], a[i], b[i]);
                    if (kSilu) y = wdm_silu(y);
                    v[i] = y;
                }
                Vec<__nv_bfloat16, V>::store(o + (long long)px * C, v);
            }
        }
    }
}

struct GnGeom {
    int V, nvec, threads;
};
inline bool gn_geom_apply(int C0, int C1, GnGeom* g) {
    const int C = C0 + C1;
    if ((C % 32) || ((C / 32) % 4) || (C0 % 8) || (C % 8) || C / 8 > 256) return false;
    g->V = 8;
    g->nvec = C / 8;
    g->threads = (256 / g->nvec) * g->nvec;
    return true;
}
inline bool gn_geom(int C0, int C1, GnGeom* g) {
    const int C = C0 + C1;
    if (C % 32) return false;
    int V = 4;
    if (C / 4 > 256) V = 8;
    if ((C % V) || (C0 % V) || ((C / 32) % V) || C / V > 256) return false;
    g->V = V;
    g->nvec = C / V;
    g->threads = (256 / g->nvec) * g->nvec;
    return true;
}

// ---------------------------------------------------------------------------------------------- upsample
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ src, int P, int H, int W, int C8, T* __restrict__ out) {
    // one thread per 16-byte vector of the OUTPUT
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P * 4 * H * W * C8;
    if (i >= total) return;
    const int cv = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % (2 * W));
    r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int p = (int)(r / (2 * H));
    const uint4* s = reinterpret_cast<const uint4*>(src) + (((long long)p * H + (oy >> 1)) * W + (ox >> 1)) * C8 + cv;
    reinterpret_cast<uint4*>(out)[i] = __ldg(s);
}

// ---------------------------------------------------------------------------------------------- softmax
// one warp per row, L <= 1024, L % 32 == 0
// seg > 0: block-diagonal attention over groups of L/seg patches packed in one row (the 8x8 mid-block runs two
// 64-token patches per 128-row tile): row r only attends to columns [seg*((r/seg) % (L/seg)), +seg); the others get 0.
template <typename TO>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, long long rows, int L, int seg,
                                                           TO* __restrict__ out) {
    wdm_grid_launch_dependents();
    wdm_grid_dependency_wait();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* s = S + row * L;
    float v[32];
    const int n = L >> 5;
    const int c0 = seg > 0 ? (int)((row / seg) % (L / seg)) * seg : 0, c1 = seg > 0 ? c0 + seg : L;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) {
            const int col = i * 32 + lane;
            v[i] = (col >= c0 && col < c1) ? s[col] : -INFINITY;
            mx = fmaxf(mx, v[i]);
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) {
            v[i] = expf(v[i] - mx);
            sum += v[i];
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    TO* d = out + row * L;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < n) {
            const float y = v[i] / sum;
            if (sizeof(TO) == 4)
                reinterpret_cast<float*>(d)[i * 32 + lane] = y;
            else
                reinterpret_cast<__nv_bfloat16*>(d)[i * 32 + lane] = __float2bfloat16(y);
        }
}

// ---------------------------------------------------------------------------------------------- temb
// out[t][n] = act( sum_k in[t][k] * W[n][k] + b[n] ); one warp per (t, n).
// mode 0: in = sinusoidal embedding of t computed on the fly from freqs (K = 2*half)
// act_silu: apply SiLU to the OUTPUT (the consumers all take silu(.) of it)
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, const float* __restrict__ tvals,
                                                     const float* __restrict__ freqs, int half,
                                                     const float* __restrict__ W, const float* __restrict__ b, int T,
                                                     int N, int K, int act_silu, float* __restrict__ out) {
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (long long)T * N) return;
    const int t = (int)(wid / N), n = (int)(wid % N);
    const int lane = threadIdx.x & 31;
    const float* w = W + (long long)n * K;
    float acc = 0.f;
    if (tvals) {
        const float tv = tvals[t];
        for (int k = lane; k < K; k += 32) {
            float e;
            if (k < half)
                e = sinf(__fmul_rn(tv, freqs[k]));
            else if (k < 2 * half)
                e = cosf(__fmul_rn(tv, freqs[k - half]));
            else
                e = 0.f;
            acc = fmaf(e, w[k], acc);
        }
    } else {
        const float* x = in + (long long)t * K;
        for (int k = lane; k < K; k += 32) acc = fmaf(x[k], w[k], acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        float y = acc + b[n];
        if (act_silu) y = silu_precise(y);
        out[(long long)t * N + n] = y;
    }
}

// ---------------------------------------------------------------------------------------------- gather
// grid (R rows, P patches). The row of R pixels x Cpad channels is staged in OUTPUT layout ([x][Cpad], output dtype) in
// 32-bit words with an odd row pitch: the scatter of one channel over consecutive x and the row-contiguous copy-out (16-byte
// global stores) are both bank-conflict free. (The first version staged fp32 [c][x] and wrote 2-byte global stores:
// 116 us per 64-patch step = 22 % of the HBM roofline.)
template <typename TO>
__global__ void __launch_bounds__(256) gather_patches_kernel(const GatherParams p) {
    extern __shared__ __align__(16) unsigned char smem_g[];
    constexpr int kPerWord = 4 / (int)sizeof(TO);
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem_g);
    const int y = blockIdx.x, pi = blockIdx.y;
    const int R = p.R, Cpad = p.Cpad;
    const int wpr = Cpad / kPerWord, pitch = wpr | 1;
    const int img = p.patches[pi * 3], hi = p.patches[pi * 3 + 1], wi = p.patches[pi * 3 + 2];
    const int c01 = p.Cs[0] + p.Cs[1];
    const int ctot = c01 + (p.nsrc > 2 ? p.Cs[2] : 0);
    auto src_of = [&](int c, int x) -> const float* {
        const float* s;
        int cs, Cs;
        if (c < p.Cs[0])
            s = p.src[0], cs = c, Cs = p.Cs[0];
        else if (c < c01)
            s = p.src[1], cs = c - p.Cs[0], Cs = p.Cs[1];
        else
            s = p.src[2], cs = c - c01, Cs = p.Cs[2];
        return s + (((long long)img * Cs + cs) * p.h + (hi + y)) * p.w + (wi + x);
    };
    // thread = (x, channel lane): no per-element division when R divides the block (R = 16, 64, 256 in practice); kU
    // independent loads in flight per thread before the first shared-memory store
    constexpr int kU = 8;
    if ((blockDim.x % R) == 0) {
        const int x = threadIdx.x % R, cl = threadIdx.x / R, cstep = blockDim.x / R;
        TO* col = reinterpret_cast<TO*>(tile + x * pitch);
        for (int c0 = cl; c0 < ctot; c0 += kU * cstep) {
            float v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int c = c0 + u * cstep;
                v[u] = c < ctot ? __ldg(src_of(c, x)) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int c = c0 + u * cstep;
                if (c < ctot) col[c] = TO(v[u]);
            }
        }
    } else {
        for (int i = threadIdx.x; i < ctot * R; i += blockDim.x) {
            const int c = i / R, x = i - c * R;
            reinterpret_cast<TO*>(tile + x * pitch)[c] = TO(__ldg(src_of(c, x)));
        }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<TO*>(p.out) + ((long long)pi * R + y) * R * Cpad);
    const int valid_words = (ctot + kPerWord - 1) / kPerWord;  // channels >= ctot are zero padding
    const bool odd_tail = kPerWord == 2 && (ctot & 1);          // last valid word holds one real channel
    const int qpr = wpr >> 2;
    for (int e = threadIdx.x; e < R * qpr; e += blockDim.x) {
        const int x = e / qpr, cw = (e - x * qpr) * 4;
        const uint32_t* t = tile + x * pitch + cw;
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            w[k] = cw + k < valid_words ? t[k] : 0u;
            if (odd_tail && cw + k == valid_words - 1) w[k] &= 0x0000ffffu;
        }
        dst[e] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Refresh of ONE source's channels inside an already gathered tensor: out[p, y, x, c_off + c] = src[img_p, c, hi_p+y, wi_p+x],
// every other channel untouched. The sampler's conditioning sources (x_cond, x_other: 93 of the 96 channels) do not
// change between DDIM steps, only x_t (3 channels) does: steps 2..S move 6 B per pixel instead of re-gathering 256 B.
template <typename TO>
__global__ void __launch_bounds__(256) gather_update_kernel(const float* __restrict__ src, int C, int c_off, int h, int w,
                                                            const int* __restrict__ patches, int R, int Cpad,
                                                            TO* __restrict__ out) {
    const int pi = blockIdx.y;
    const int img = patches[pi * 3], hi = patches[pi * 3 + 1], wi = patches[pi * 3 + 2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pixel of the patch, x fastest
    if (i >= R * R) return;
    const int y = i / R, x = i - y * R;
    TO* o = out + ((long long)pi * R * R + i) * Cpad + c_off;
    const float* s = src + ((long long)img * C * h + (hi + y)) * w + (wi + x);
    for (int c = 0; c < C; ++c) o[c] = TO(__ldg(s + (long long)c * h * w));
}

// ---------------------------------------------------------------------------------------------- DDIM step
// One thread per image element. Every arithmetic step is an explicitly rounded fp32 op in the order torch
// eager evaluates models/ddm_wavelet.py:496-502 (no FMA contraction) so the update is bit-identical to
// the oracle given the same eps.
__global__ void __launch_bounds__(256) ddim_step_kernel(const DdimParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)p.B * p.Cp * p.h * p.w;
    if (i >= total) return;
    const int x = (int)(i % p.w);
    long long r = i / p.w;
    const int y = (int)(r % p.h);
    r /= p.h;
    const int c = (int)(r % p.Cp);
    const int b = (int)(r / p.Cp);
    float sum = 0.f;
    float cnt = 0.f;
    const int q0 = p.img_first[b], q1 = p.img_first[b + 1];
    for (int q = q0; q < q1; ++q) {
        const int hi = p.patches[q * 3 + 1], wi = p.patches[q * 3 + 2];
        const int py = y - hi, px = x - wi;
        if (py >= 0 && py < p.R && px >= 0 && px < p.R) {
            sum = __fadd_rn(sum, p.eps[(((long long)q * p.Cp + c) * p.R + py) * p.R + px]);
            cnt += 1.f;
        }
    }
    const float et = __fdiv_rn(sum, cnt);
    const float xt = p.xt[i];
    const float s1 = __fsqrt_rn(__fsub_rn(1.0f, p.at));
    const float x0 = __fdiv_rn(__fsub_rn(xt, __fmul_rn(et, s1)), __fsqrt_rn(p.at));
    p.x0_out[i] = x0;
    const float c2 = __fsqrt_rn(__fsub_rn(1.0f, p.at_next));
    // at_next.sqrt()*x0 + c1*randn (c1 = 0 for eta = 0) + c2*et
    const float xn = __fadd_rn(__fadd_rn(__fmul_rn(__fsqrt_rn(p.at_next), x0), 0.0f), __fmul_rn(c2, et));
    p.xt_next[i] = xn;
}

// ---------------------------------------------------------------------------------------------- packing
template <typename TO>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int Cin_pad,
                                        TO* __restrict__ out, long long ldk, long long k_off) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)Cout * taps * Cin_pad;
    if (i >= total) return;
    const int ci = (int)(i % Cin_pad);
    long long r = i / Cin_pad;
    const int tap = (int)(r % taps);
    const int co = (int)(r / taps);
    const float v = ci < Cin ? w[((long long)co * Cin + ci) * taps + tap] : 0.f;
    TO* d = out + (long long)co * ldk + k_off + (long long)tap * Cin_pad + ci;
    if (sizeof(TO) == 4)
        *reinterpret_cast<float*>(d) = v;
    else
        *reinterpret_cast<__nv_bfloat16*>(d) = __float2bfloat16(v);
}

__global__ void vec_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

// nearest-x2-upsample followed by a 3x3 conv == four 2x2 convs on the source grid, one per output phase (py, px):
//   out[2i+py][2j+px] = sum_{ty,tx} Wp[py][px][ty][tx] . x[i + py - 1 + ty][j + px - 1 + tx]
//   Wp[0][.][0] = W[0], Wp[0][.][1] = W[1] + W[2];  Wp[1][.][0] = W[0] + W[1], Wp[1][.][1] = W[2]   (rows; same for columns)
// w: OIHW fp32 [Cout][Cin][3][3]  ->  out[ph][Cout][t][Cin], ph = 2 py + px, t = 2 ty + tx (summed in fp32, then rounded)
template <typename TO>
__global__ void pack_subpix_weight_kernel(const float* __restrict__ w, int Cout, int Cin, TO* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = 16LL * Cout * Cin;
    if (i >= total) return;
    const int ci = (int)(i % Cin);
    long long r = i / Cin;
    const int t = (int)(r % 4);
    r /= 4;
    const int co = (int)(r % Cout);
    const int ph = (int)(r / Cout);
    const int py = ph >> 1, px = ph & 1, ty = t >> 1, tx = t & 1;
    const int y0 = py == 0 ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), y1 = py == 0 ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
    const int x0 = px == 0 ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), x1 = px == 0 ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
    const float* wp = w + ((long long)co * Cin + ci) * 9;
    float acc = 0.f;
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) acc += wp[y * 3 + x];
    if (sizeof(TO) == 4)
        reinterpret_cast<float*>(out)[i] = acc;
    else
        reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16(acc);
}

}  // namespace

// Pack-time products of two C x C fp32 matrices (the algebraic attention fusion in wdm_unet.cu): one thread per output,
// fp32 FMA chain over c in ascending order (deterministic).  mode 0: out = X^T Y,  mode 1: out = X Y.
__global__ void __launch_bounds__(256) matmul_cc_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                       float* __restrict__ out, int C, int mode) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= C) return;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(mode == 0 ? X[(long long)c * C + a] : X[(long long)a * C + c], Y[(long long)c * C + b], acc);
    out[(long long)a * C + b] = acc;
}
// out[a] = sum_c (mode 0: X[c][a], mode 1: X[a][c]) * v[c] (+ add[a])
__global__ void __launch_bounds__(256) matvec_c_kernel(const float* __restrict__ X, const float* __restrict__ v,
                                                      const float* __restrict__ add, float* __restrict__ out, int C, int mode) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= C) return;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(mode == 0 ? X[(long long)c * C + a] : X[(long long)a * C + c], v[c], acc);
    out[a] = acc + (add ? add[a] : 0.f);
}

// ================================================================================================ launchers
int launch_matmul_cc(const float* X, const float* Y, float* out, int C, int mode, cudaStream_t s) {
    matmul_cc_kernel<<<dim3((C + 255) / 256, C), 256, 0, s>>>(X, Y, out, C, mode);
    return wdm_launch_status();
}
int launch_matvec_c(const float* X, const float* v, const float* add, float* out, int C, int mode, cudaStream_t s) {
    matvec_c_kernel<<<(C + 255) / 256, 256, 0, s>>>(X, v, add, out, C, mode);
    return wdm_launch_status();
}

int launch_vec_add(const float* a, const float* b, float* out, int n, cudaStream_t s) {
    vec_add_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, out, n);
    return wdm_launch_status();
}

int launch_pack_subpix_weight(const float* w, int Cout, int Cin, void* out, int out_dtype, cudaStream_t s) {
    const long long total = 16LL * Cout * Cin;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (out_dtype == DT_F32)
        pack_subpix_weight_kernel<float><<<grid, 256, 0, s>>>(w, Cout, Cin, reinterpret_cast<float*>(out));
    else
        pack_subpix_weight_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(out));
    return wdm_launch_status();
}

static int pick_slabs(int P, int HW) {
    int S = 1;
    while (S < kMaxSlabs && P * S < 2 * 148 && HW / (S * 2) >= 16 && HW % (S * 2) == 0) S *= 2;
    return S;
}

size_t gn_partial_bytes(int P) { return (size_t)P * kMaxSlabs * 32 * sizeof(double2); }

int launch_gn_stats(const void* src0, int C0, const void* src1, int C1, int dtype, int P, int HW, float eps,
                    float* stats, cudaStream_t s) {
    // `stats` buffer layout: [P*32*2 floats (mean, rstd)] followed (256-byte aligned) by the double2 partials.
    GnGeom g;
    if (!gn_geom(C0, C1, &g)) return WDM_ERR_BAD_SHAPE;
    const int S = pick_slabs(P, HW);
    double2* partial = reinterpret_cast<double2*>(reinterpret_cast<char*>(stats) + (((size_t)P * 64 * 4 + 255) & ~(size_t)255));
    dim3 grid(S, P);
#define WDM_GN_STATS(T, V)                                                                                     \
    gn_stats_kernel<T, V><<<grid, g.threads, 0, s>>>(reinterpret_cast<const T*>(src0), C0,                      \
                                                     reinterpret_cast<const T*>(src1), C1, HW, S, partial)
    if (dtype == DT_F32) {
        if (g.V == 4) WDM_GN_STATS(float, 4); else WDM_GN_STATS(float, 8);
    } else {
        if (g.V == 4) WDM_GN_STATS(__nv_bfloat16, 4); else WDM_GN_STATS(__nv_bfloat16, 8);
    }
#undef WDM_GN_STATS
    int st = wdm_launch_status();
    if (st != WDM_OK) return st;
    const int n_pg = P * 32;
    const double inv_n = 1.0 / ((double)HW * ((C0 + C1) / 32));
    gn_finalize_kernel<<<(n_pg + 127) / 128, 128, 0, s>>>(partial, S, n_pg, inv_n, (double)eps, stats);
    return wdm_launch_status();
}

size_t gn_stats_bytes(int P) { return (((size_t)P * 64 * 4 + 255) & ~(size_t)255) + gn_partial_bytes(P); }

int launch_gn_apply(const void* src0, int C0, const void* src1, int C1, int dtype, int P, int HW, const float* stats,
                    const float* gamma, const float* beta, int silu, void* out, cudaStream_t s) {
    GnGeom g;
    if (!gn_geom_apply(C0, C1, &g)) return WDM_ERR_BAD_SHAPE;
    const int ppi = g.threads / g.nvec;
    int pix_per_cta = ppi * 8;  // one pass of U = 8 vectors per thread (see gn_apply_kernel)
    if (pix_per_cta > HW) pix_per_cta = HW;
    dim3 grid((HW + pix_per_cta - 1) / pix_per_cta, P);
#define WDM_GN_APPLY(T, SILU, PREC)                                                                               \
    wdm_launch_pdl(gn_apply_kernel<T, 8, SILU, PREC>, grid, dim3(g.threads), 0, s, reinterpret_cast<const T*>(src0), C0, \
                   reinterpret_cast<const T*>(src1), C1, HW, pix_per_cta, stats, gamma, beta, reinterpret_cast<T*>(out))
    if (dtype == DT_F32) {
        if (silu) WDM_GN_APPLY(float, true, true); else WDM_GN_APPLY(float, false, true);
    } else {
        if (silu) WDM_GN_APPLY(__nv_bfloat16, true, false); else WDM_GN_APPLY(__nv_bfloat16, false, false);
    }
#undef WDM_GN_APPLY
    return wdm_launch_status();
}

int launch_gn_apply_sidecar(const void* src0, int C0, const float* sc0, const void* src1, int C1, const float* sc1, int P, int HW,
                            float eps, const float* gamma, const float* beta, int silu, void* out, cudaStream_t s) {
    GnGeom g;
    if (!gn_geom_apply(C0, C1, &g)) return WDM_ERR_BAD_SHAPE;
    if (((C0 + C1) % 128) || (C0 % 4) || (C1 % 4) || (HW % 32) || !sc0 || (C1 && !sc1) || (C0 + C1) / 4 > 384) return WDM_ERR_BAD_SHAPE;
    static const int reverse = []() {
        const char* e = getenv("WDM_GN_REVERSE");
        return e ? atoi(e) : 0;  // measured: no effect (5.19 vs 5.17 ms)
    }();
    const int ppi = g.threads / g.nvec;
    const int pass_px = ppi * 8;                          // pixels of one pass (U = 8 vectors per thread)
    int max_chunks = (HW + pass_px - 1) / pass_px;
    // enough CTAs to fill the chip about twice, as few as possible beyond that: each CTA re-reduces its patch's side-car
    int chunks = (2 * 148 + P - 1) / P;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int passes = (max_chunks + chunks - 1) / chunks;
    const int pix_per_cta = passes * pass_px;
    dim3 grid((HW + pix_per_cta - 1) / pix_per_cta, P);
    const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(src0);
    const __nv_bfloat16* a1 = reinterpret_cast<const __nv_bfloat16*>(src1);
    cudaError_t e;
    if (silu)
        e = wdm_launch_pdl(gn_apply_sc_kernel<true>, grid, dim3(g.threads), 0, s, a0, C0, sc0, a1, C1, sc1, HW, pix_per_cta,
                           (double)eps, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), reverse);
    else
        e = wdm_launch_pdl(gn_apply_sc_kernel<false>, grid, dim3(g.threads), 0, s, a0, C0, sc0, a1, C1, sc1, HW, pix_per_cta,
                           (double)eps, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), reverse);
    if (e != cudaSuccess) return wdm_cuda_error((int)e);
    return wdm_launch_status();
}

int launch_gn_finalize_sidecar(const float* sc0, int C0, const float* sc1, int C1, int P, int HW, float eps,
                               float* stats, cudaStream_t s) {
    if (((C0 + C1) % 128) || (C0 % 4) || (C1 % 4) || (HW % 32) || !sc0 || (C1 && !sc1)) return WDM_ERR_BAD_SHAPE;
    const int warps = P * 32;
    wdm_launch_pdl(gn_finalize_sidecar_kernel, dim3((warps + 7) / 8), dim3(256), 0, s, sc0, C0, sc1, C1, HW, P, (double)eps,
                   stats);
    return wdm_launch_status();
}

int launch_upsample2x(const void* src, int dtype, int P, int H, int W, int C, void* out, cudaStream_t s) {
    const int per16 = dtype == DT_F32 ? 4 : 8;
    if (C % per16) return WDM_ERR_BAD_SHAPE;
    const int C8 = C / per16;
    const long long total = (long long)P * 4 * H * W * C8;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == DT_F32)
        upsample2x_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(src), P, H, W, C8,
                                                       reinterpret_cast<float*>(out));
    else
        upsample2x_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), P, H, W, C8,
                                                               reinterpret_cast<__nv_bfloat16*>(out));
    return wdm_launch_status();
}

int launch_softmax_rows(const float* S, int rows, int L, void* out, int out_dtype, cudaStream_t s, int seg) {
    if (L % 32 || L > 1024 || (seg > 0 && (L % seg))) return WDM_ERR_BAD_SHAPE;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (out_dtype == DT_F32)
        wdm_launch_pdl(softmax_rows_kernel<float>, dim3(grid), dim3(256), 0, s, S, (long long)rows, L, seg, reinterpret_cast<float*>(out));
    else
        wdm_launch_pdl(softmax_rows_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, S, (long long)rows, L, seg,
                       reinterpret_cast<__nv_bfloat16*>(out));
    return wdm_launch_status();
}

int launch_temb(const TembParams& p, cudaStream_t s) {
    const int tc = 4 * p.ch;
    float* h1 = p.scratch;                        // silu(dense0(emb(t)))     [T][tc]
    float* h2 = p.scratch + (long long)p.T * tc;  // silu(dense1(h1))         [T][tc]
    auto grid = [&](long long warps) { return (unsigned)((warps + 7) / 8); };
    int st;
    linear_kernel<<<grid((long long)p.T * tc), 256, 0, s>>>(nullptr, p.t, p.freqs, p.ch / 2, p.w0, p.b0, p.T, tc, p.ch,
                                                           1, h1);
    if ((st = wdm_launch_status()) != WDM_OK) return st;
    linear_kernel<<<grid((long long)p.T * tc), 256, 0, s>>>(h1, nullptr, nullptr, 0, p.w1, p.b1, p.T, tc, tc, 1, h2);
    if ((st = wdm_launch_status()) != WDM_OK) return st;
    linear_kernel<<<grid((long long)p.T * p.total), 256, 0, s>>>(h2, nullptr, nullptr, 0, p.wp, p.bp, p.T, p.total, tc,
                                                                0, p.out);
    return wdm_launch_status();
}

int launch_gather_patches(const GatherParams& p, cudaStream_t s) {
    if (p.P <= 0) return WDM_OK;
    if (p.nsrc < 1 || p.nsrc > 3) return WDM_ERR_BAD_ARG;
    if (p.Cpad % (p.out_dtype == DT_F32 ? 4 : 8)) return WDM_ERR_BAD_SHAPE;  // 16-byte copy-out units
    const size_t wpr = p.out_dtype == DT_F32 ? (size_t)p.Cpad : (size_t)p.Cpad / 2;
    const size_t smem = (size_t)p.R * (wpr | 1) * 4;
    if (smem > 96 * 1024) return WDM_ERR_BAD_SHAPE;
    dim3 grid(p.R, p.P);
    if (p.out_dtype == DT_F32) {
        cudaFuncSetAttribute(gather_patches_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        gather_patches_kernel<float><<<grid, 256, smem, s>>>(p);
    } else {
        cudaFuncSetAttribute(gather_patches_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             96 * 1024);
        gather_patches_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(p);
    }
    return wdm_launch_status();
}

int launch_gather_update(const float* src, int C, int c_off, int h, int w, const int* patches, int P, int R, int Cpad,
                         void* out, int out_dtype, cudaStream_t s) {
    if (P <= 0 || C <= 0) return WDM_OK;
    dim3 grid((R * R + 255) / 256, P);
    if (out_dtype == DT_F32)
        gather_update_kernel<float><<<grid, 256, 0, s>>>(src, C, c_off, h, w, patches, R, Cpad, reinterpret_cast<float*>(out));
    else
        gather_update_kernel<__nv_bfloat16>
            <<<grid, 256, 0, s>>>(src, C, c_off, h, w, patches, R, Cpad, reinterpret_cast<__nv_bfloat16*>(out));
    return wdm_launch_status();
}

int launch_ddim_step(const DdimParams& p, cudaStream_t s) {
    const long long total = (long long)p.B * p.Cp * p.h * p.w;
    if (total <= 0) return WDM_OK;
    ddim_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
    return wdm_launch_status();
}

int launch_pack_conv_weight(const float* w, int Cout, int Cin, int taps, int Cin_pad, void* out, int out_dtype,
                            long long ldk, long long k_off, cudaStream_t s) {
    const long long total = (long long)Cout * taps * Cin_pad;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (out_dtype == DT_F32)
        pack_conv_weight_kernel<float><<<grid, 256, 0, s>>>(w, Cout, Cin, taps, Cin_pad, reinterpret_cast<float*>(out),
                                                           ldk, k_off);
    else
        pack_conv_weight_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(w, Cout, Cin, taps, Cin_pad,
                                                                   reinterpret_cast<__nv_bfloat16*>(out), ldk, k_off);
    return wdm_launch_status();
}

// ------------------------------------------------------------------------------------------------ tc32 operand split
// fp32 -> three bf16 pieces: p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) (the subtractions are exact in fp32),
// x = p0 + p1 + p2 up to 2^-24 |x|. The tensor-core kernel accumulates the six products
//   x0 w0 + x0 w1 + x1 w0 + x1 w1 + x0 w2 + x2 w0
// in fp32 (every bf16 x bf16 product is exact in fp32); the dropped terms x1 w2, x2 w1, x2 w2 are <= 2^-24 relative, i.e.
// at the rounding level of an fp32 FFMA chain. See GemmParams::a_split3 / a_chunk() in wdm_gemm_tc.cu.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
    p0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(p0);
    p1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(p1);
    p2 = __float2bfloat16_rn(r2);
}

// out[row][piece * C + c] for the channel concat of up to two fp32 sources, C = C0 + C1 (both multiples of 4)
__global__ void __launch_bounds__(256) split3_act_kernel(const float* __restrict__ s0, int C0, const float* __restrict__ s1, int C1,
                                                        long long rows, __nv_bfloat16* __restrict__ out) {
    const int C = C0 + C1, cq = C >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cq) return;
    const long long row = i / cq;
    const int c = (int)(i - row * cq) * 4;
    const float4 v = c < C0 ? *reinterpret_cast<const float4*>(s0 + row * C0 + c) : *reinterpret_cast<const float4*>(s1 + row * C1 + (c - C0));
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 q[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3(x[j], q[0][j], q[1][j], q[2][j]);
#pragma unroll
    for (int pc = 0; pc < 3; ++pc)
        *reinterpret_cast<uint2*>(out + row * 3 * C + pc * C + c) = *reinterpret_cast<const uint2*>(q[pc]);
}

// w: fp32 [N][taps][C] (K-major packed rows) -> out bf16 [N][taps][6][C], product order (w0, w1, w0, w1, w2, w0);
// with out_main: the dominant product apart -- out_main [N][taps][C] = w0 and out [N][taps][5][C] = (w1, w0, w1, w2, w0)
__global__ void __launch_bounds__(256) split3_weight_kernel(const float* __restrict__ w, long long total, int C,
                                                           __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_main) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long nt = i / C;  // (n, tap) row
    const int c = (int)(i - nt * C);
    __nv_bfloat16 q[3];
    split3(w[i], q[0], q[1], q[2]);
    if (out_main) {
        out_main[i] = q[0];
        __nv_bfloat16* o = out + nt * 5 * C + c;
        o[0] = q[1], o[C] = q[0], o[2 * C] = q[1], o[3 * C] = q[2], o[4 * C] = q[0];
    } else {
        __nv_bfloat16* o = out + nt * 6 * C + c;
        o[0] = q[0], o[C] = q[1], o[2 * C] = q[0], o[3 * C] = q[1], o[4 * C] = q[2], o[5 * C] = q[0];
    }
}

int launch_split3_act(const float* src0, int C0, const float* src1, int C1, long long rows, void* out, cudaStream_t s) {
    if (!src0 || !out || C0 <= 0 || (C0 % 4) || (C1 % 4) || (C1 && !src1)) return WDM_ERR_BAD_SHAPE;
    if (rows <= 0) return WDM_OK;
    const long long n = rows * ((C0 + C1) / 4);
    split3_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src0, C0, src1, C1, rows, reinterpret_cast<__nv_bfloat16*>(out));
    return wdm_launch_status();
}

int launch_split3_weight(const float* w, int N, int taps, int C, void* out, void* out_main, cudaStream_t s) {
    if (!w || !out || N <= 0 || taps <= 0 || C <= 0) return WDM_ERR_BAD_SHAPE;
    const long long total = (long long)N * taps * C;
    split3_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, total, C, reinterpret_cast<__nv_bfloat16*>(out),
                                                                        reinterpret_cast<__nv_bfloat16*>(out_main));
    return wdm_launch_status();
}

__global__ void __launch_bounds__(256) rows_to_nchw_kernel(const float* __restrict__ y, int ld, long long total, int HW, int C,
                                                          float* __restrict__ x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output index (p, c, pix)
    if (i >= total) return;
    const int pix = (int)(i % HW);
    const long long pc = i / HW;
    const int c = (int)(pc % C);
    const long long pp = pc / C;
    x[i] = y[(pp * HW + pix) * ld + c];
}

int launch_rows_to_nchw(const float* y, int ld, int P, int HW, int C, float* x, cudaStream_t s) {
    if (!y || !x || C <= 0 || C > ld) return WDM_ERR_BAD_SHAPE;
    const long long total = (long long)P * C * HW;
    if (total <= 0) return WDM_OK;
    rows_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(y, ld, total, HW, C, x);
    return wdm_launch_status();
}

}  // namespace wdm
