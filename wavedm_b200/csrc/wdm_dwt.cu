// wdm_dwt.cu -- 2-level Haar-packet DWT / IWT (scale=2, "c2") as single HBM-bound sm_100a kernels.
//
// Replaces models/wavelet.py:36-50 of the reference (grouped stride-4 conv + permute copy, and its
// transposed twin). Closed form:
//   y[n, 3k+g, i, j] = sum_{r,c} 0.25 (-1)^(b0 c_hi + b1 r_hi + b2 c_lo + b3 r_lo) x[n, g, 4i+r, 4j+c]
// evaluated as a 4-stage butterfly in exactly the order of oracle/dwt_oracle.c (wht16_fwd / wht16_inv),
// so results are bit-identical to the oracle's lifting form.
//
// Two variants per direction:
//   DIRECT : one 4x4 block per thread, 4 x LDG.128 (512 B contiguous per warp per row) -> 16 x STG.32
//            (128 B contiguous per warp per sub-band plane); registers only.
//   TMA    : persistent CTAs; cp.async.bulk.tensor tiles (3-D map of the image planes, 5-D map of the
//            sub-band tensor) staged in shared memory through a 3-stage mbarrier ring, butterflies from
//            smem, results staged in smem and written back with TMA stores (bulk groups).
// Algorithmic HBM bytes: 2 * 4 B * 3*H*W per image (read once, write once).
#include "wdm_common.cuh"
#include "wdm_ptx.cuh"
#include "wdm_tmap.h"

namespace {

using namespace wdm;

// ---- butterflies (must mirror oracle/dwt_oracle.c exactly) ------------------------------------------
__device__ __forceinline__ void wht16_fwd(const float (&v)[4][4], float (&o)[16]) {
    float u[4][2][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float s00 = v[r][0] + v[r][1];
        float s01 = v[r][0] - v[r][1];
        float s10 = v[r][2] + v[r][3];
        float s11 = v[r][2] - v[r][3];
        u[r][0][0] = s00 + s10;
        u[r][1][0] = s00 - s10;
        u[r][0][1] = s01 + s11;
        u[r][1][1] = s01 - s11;
    }
#pragma unroll
    for (int b0 = 0; b0 < 2; ++b0)
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
            float t00 = u[0][b0][b2] + u[1][b0][b2];
            float t01 = u[0][b0][b2] - u[1][b0][b2];
            float t10 = u[2][b0][b2] + u[3][b0][b2];
            float t11 = u[2][b0][b2] - u[3][b0][b2];
            o[b0 + 0 + 4 * b2 + 0] = (t00 + t10) * 0.25f;
            o[b0 + 2 + 4 * b2 + 0] = (t00 - t10) * 0.25f;
            o[b0 + 0 + 4 * b2 + 8] = (t01 + t11) * 0.25f;
            o[b0 + 2 + 4 * b2 + 8] = (t01 - t11) * 0.25f;
        }
}

__device__ __forceinline__ void wht16_inv(const float (&in)[16], float (&v)[4][4]) {
    float u[2][2][4];
#pragma unroll
    for (int b1 = 0; b1 < 2; ++b1)
#pragma unroll
        for (int b3 = 0; b3 < 2; ++b3) {
            const int p = 2 * b1 + 8 * b3;
            float s00 = in[p + 0] + in[p + 4];
            float s01 = in[p + 0] - in[p + 4];
            float s10 = in[p + 1] + in[p + 5];
            float s11 = in[p + 1] - in[p + 5];
            u[b1][b3][0] = s00 + s10;
            u[b1][b3][2] = s00 - s10;
            u[b1][b3][1] = s01 + s11;
            u[b1][b3][3] = s01 - s11;
        }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float t00 = u[0][0][c] + u[0][1][c];
        float t01 = u[0][0][c] - u[0][1][c];
        float t10 = u[1][0][c] + u[1][1][c];
        float t11 = u[1][0][c] - u[1][1][c];
        v[0][c] = (t00 + t10) * 0.25f;
        v[2][c] = (t00 - t10) * 0.25f;
        v[1][c] = (t01 + t11) * 0.25f;
        v[3][c] = (t01 - t11) * 0.25f;
    }
}

__device__ __forceinline__ float pre_2xm1(float t) { return 2.0f * t - 1.0f; }
__device__ __forceinline__ float post_clamp(float t) {
    t = (t + 1.0f) / 2.0f;
    return fminf(fmaxf(t, 0.0f), 1.0f);
}

// ---- DIRECT variants ----------------------------------------------------------------------------------
template <bool kPre>
__global__ void __launch_bounds__(256) dwt4x4_direct_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            long long nblocks, int h, int w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nblocks) return;
    const int j = (int)(idx % w);
    const long long t1 = idx / w;
    const int i = (int)(t1 % h);
    const long long p = t1 / h;  // plane = n*3 + g
    const int W = 4 * w;
    const float* src = x + ((p * (4LL * h) + 4 * i) * W + 4 * j);
    float v[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float4 q = wdm_ldg_stream(reinterpret_cast<const float4*>(src + (long long)r * W));
        v[r][0] = q.x, v[r][1] = q.y, v[r][2] = q.z, v[r][3] = q.w;
    }
    if (kPre) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) v[r][c] = pre_2xm1(v[r][c]);
    }
    float o[16];
    wht16_fwd(v, o);
    const long long n = p / 3;
    const int g = (int)(p - n * 3);
    const long long plane = (long long)h * w;
    float* dst = y + ((n * 48 + g) * plane + (long long)i * w + j);
#pragma unroll
    for (int k = 0; k < 16; ++k) wdm_stg_stream(dst + 3LL * k * plane, o[k]);
}

template <bool kPost>
__global__ void __launch_bounds__(256) iwt4x4_direct_kernel(const float* __restrict__ y, float* __restrict__ x,
                                                            long long nblocks, int h, int w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nblocks) return;
    const int j = (int)(idx % w);
    const long long t1 = idx / w;
    const int i = (int)(t1 % h);
    const long long p = t1 / h;
    const long long n = p / 3;
    const int g = (int)(p - n * 3);
    const long long plane = (long long)h * w;
    const float* src = y + ((n * 48 + g) * plane + (long long)i * w + j);
    float in[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) in[k] = wdm_ldg_stream(src + 3LL * k * plane);
    float v[4][4];
    wht16_inv(in, v);
    const int W = 4 * w;
    float* dst = x + ((p * (4LL * h) + 4 * i) * W + 4 * j);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float4 q;
        if (kPost)
            q = make_float4(post_clamp(v[r][0]), post_clamp(v[r][1]), post_clamp(v[r][2]), post_clamp(v[r][3]));
        else
            q = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        wdm_stg_stream(reinterpret_cast<float4*>(dst + (long long)r * W), q);
    }
}

// ---- restore() epilogue (models/restoration.py:111-135): torch.cat([latent[:, :pred], hf[:, pred:]]) -> wavelet_rec ->
// inverse_data_transform as ONE kernel: sub-band channel c comes from `lo` when c < Clo, else from `hi` (channel c of a
// 48-channel tensor, or channel c - Clo of a tensor that holds only the remaining bands when hi_has_all == 0).
template <bool kPost>
__global__ void __launch_bounds__(256) iwt4x4_cat_kernel(const float* __restrict__ lo, int Clo, const float* __restrict__ hi,
                                                         int Chi, int hi_off, float* __restrict__ x, long long nblocks,
                                                         int h, int w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nblocks) return;
    const int j = (int)(idx % w);
    const long long t1 = idx / w;
    const int i = (int)(t1 % h);
    const long long p = t1 / h;
    const long long n = p / 3;
    const int g = (int)(p - n * 3);
    const long long plane = (long long)h * w;
    const long long pix = (long long)i * w + j;
    float in[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int c = 3 * k + g;
        in[k] = c < Clo ? wdm_ldg_stream(lo + (n * Clo + c) * plane + pix)
                        : wdm_ldg_stream(hi + (n * Chi + (c - hi_off)) * plane + pix);
    }
    float v[4][4];
    wht16_inv(in, v);
    const int W = 4 * w;
    float* dst = x + ((p * (4LL * h) + 4 * i) * W + 4 * j);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float4 q;
        if (kPost)
            q = make_float4(post_clamp(v[r][0]), post_clamp(v[r][1]), post_clamp(v[r][2]), post_clamp(v[r][3]));
        else
            q = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        wdm_stg_stream(reinterpret_cast<float4*>(dst + (long long)r * W), q);
    }
}

// ---- wavelet_in_unet variants (models/unet.py:338-350, 393-394) -----------------------------------------------
// DiffusionUNet(wavelet_in_unet=True) applies the DWT to each 3-channel half of its pixel-domain input INSIDE forward()
// and the IWT to its 48-channel output, i.e. once per DDIM step. Here the forward transform is fused with the sampler's
// crop + concat + NHWC conversion (ddm_wavelet.py:467-478) and the inverse reads the conv_out result in place:
//   dwt_gather : fp32 NCHW image-level sources [B,3,H,W] x nsrc, cropped at (hi, wi) in PIXELS, side 4R
//                -> UNet input [P, R, R, Cpad] NHWC (channel s*48 + 3k + g, engine dtype), pad channels zero.
//   iwt_nhwc   : conv_out result [P*R*R, ld] fp32 (first 48 columns) -> eps [P, 3, 4R, 4R] fp32 NCHW.
// One CTA per (sub-band row i, patch): thread (plane = s*3+g, j) owns one 4x4 pixel block; the row of R pixels x Cpad
// channels is staged in shared memory so that the global store is one contiguous run.
template <typename TO>
__global__ void __launch_bounds__(384) dwt_gather_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                                        int nsrc, int H, int W, const int* __restrict__ patches, int R,
                                                        int Cpad, TO* __restrict__ out) {
    // staging tile [R][Cpad] in 32-bit words with an ODD row pitch: the sub-band scatter (fixed channel, consecutive j)
    // and the row-contiguous copy-out are both bank-conflict free
    extern __shared__ __align__(16) unsigned char smem_dg[];
    constexpr int kPerWord = 4 / (int)sizeof(TO);  // 1 (fp32) or 2 (bf16) channels per word
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem_dg);
    const int wpr = Cpad / kPerWord;  // words per output pixel
    const int pitch = wpr | 1;
    const int i = blockIdx.x, pi = blockIdx.y;
    const int img = patches[pi * 3], hi = patches[pi * 3 + 1], wi = patches[pi * 3 + 2];
    const int nplanes = nsrc * 3;
    for (int u = threadIdx.x; u < nplanes * R; u += blockDim.x) {
        const int plane = u / R, j = u - plane * R;
        const int sidx = plane / 3, g = plane - 3 * sidx;
        const float* src = (sidx == 0 ? src0 : src1) + (((long long)img * 3 + g) * H + (hi + 4 * i)) * W + (wi + 4 * j);
        float v[4][4];
        const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((W & 3) == 0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (vec) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(src + (long long)r * W));
                v[r][0] = q.x, v[r][1] = q.y, v[r][2] = q.z, v[r][3] = q.w;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[r][c] = __ldg(src + (long long)r * W + c);
            }
        }
        float o[16];
        wht16_fwd(v, o);
        TO* row = reinterpret_cast<TO*>(tile + j * pitch);
#pragma unroll
        for (int k = 0; k < 16; ++k) row[sidx * 48 + 3 * k + g] = TO(o[k]);
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + ((long long)pi * R + i) * R * Cpad);
    const int valid_words = nplanes * 16 / kPerWord;  // channels >= nplanes*16 are zero padding
    const int qpr = wpr >> 2;                         // 16-byte units per output pixel (Cpad % 8 == 0)
    for (int e = threadIdx.x; e < R * qpr; e += blockDim.x) {
        const int j = e / qpr, cw = (e - j * qpr) * 4;
        const uint32_t* t = tile + j * pitch + cw;
        uint4 q;
        q.x = cw + 0 < valid_words ? t[0] : 0u, q.y = cw + 1 < valid_words ? t[1] : 0u;
        q.z = cw + 2 < valid_words ? t[2] : 0u, q.w = cw + 3 < valid_words ? t[3] : 0u;
        dst[e] = q;
    }
}

__global__ void __launch_bounds__(192) iwt_nhwc_kernel(const float* __restrict__ y, int ld, int R, float* __restrict__ x) {
    // the row of R pixels x ld floats is staged with coalesced 16-byte loads; odd pitch -> the stride-3 sub-band reads of
    // consecutive pixels fall into different banks
    extern __shared__ __align__(16) unsigned char smem_in[];
    float* tile = reinterpret_cast<float*>(smem_in);
    const int pitch = ld | 1;
    const int i = blockIdx.x, pi = blockIdx.y;
    const float4* src4 = reinterpret_cast<const float4*>(y + ((long long)pi * R + i) * R * ld);
    const int ld4 = ld >> 2;
    for (int e = threadIdx.x; e < R * 12; e += blockDim.x) {  // 12 float4 = the 48 sub-band columns of a pixel
        const int j = e / 12, c4 = e - j * 12;
        const float4 q = __ldg(src4 + j * ld4 + c4);
        float* t = tile + j * pitch + c4 * 4;
        t[0] = q.x, t[1] = q.y, t[2] = q.z, t[3] = q.w;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < 3 * R; u += blockDim.x) {
        const int g = u / R, j = u - g * R;
        const float* t = tile + j * pitch + g;
        float in[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) in[k] = t[3 * k];
        float v[4][4];
        wht16_inv(in, v);
        const int Wp = 4 * R;
        float* dst = x + (((long long)pi * 3 + g) * Wp + 4 * i) * Wp + 4 * j;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            *reinterpret_cast<float4*>(dst + (long long)r * Wp) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    }
}

// ---- TMA variants -------------------------------------------------------------------------------------
constexpr int kTW = 128;         // pixel tile width
constexpr int kTH = 32;          // pixel tile height
constexpr int kTBW = kTW / 4;    // 32 blocks per tile row (one warp)
constexpr int kTBH = kTH / 4;    // 8 block rows
constexpr int kThreads = kTBW * kTBH;  // 256: one 4x4 block per thread per tile
constexpr int kPixStages = 3;    // ring depth of the side that is LOADED
constexpr int kOutStages = 2;    // ring depth of the side that is STORED
constexpr int kTileBytes = kTW * kTH * 4;  // 16 KiB (pixel tile == sub-band tile in bytes)
constexpr int kSmemBytes = (kPixStages + kOutStages) * kTileBytes + 128;

struct TileCoord {
    int tx, ty, g, n;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_x, int tiles_y) {
    TileCoord c;
    c.tx = tile % tiles_x;
    int t = tile / tiles_x;
    c.ty = t % tiles_y;
    int p = t / tiles_y;
    c.n = p / 3;
    c.g = p - 3 * c.n;
    return c;
}

// pix_map: 3-D (W, H, n*3) box (kTW, kTH, 1); sub_map: 5-D (w, h, 3, 16, n) box (kTBW, kTBH, 1, 16, 1).
template <bool kPre>
__global__ void __launch_bounds__(kThreads) dwt4x4_tma_kernel(const __grid_constant__ CUtensorMap pix_map,
                                                              const __grid_constant__ CUtensorMap sub_map, int tiles_x,
                                                              int tiles_y, int ntiles) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    float* s_in = reinterpret_cast<float*>(smem);                               // [kPixStages][kTH][kTW]
    float* s_out = reinterpret_cast<float*>(smem + kPixStages * kTileBytes);    // [kOutStages][16][kTBH][kTBW]
    __shared__ __align__(8) uint64_t full[kPixStages];

    const int tid = threadIdx.x;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        ptx::prefetch_tmap(&pix_map);
        ptx::prefetch_tmap(&sub_map);
        for (int s = 0; s < kPixStages; ++s) ptx::mbar_init(&full[s], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < kPixStages && s < n_my; ++s) {
            TileCoord c = tile_coord(first + s * stride, tiles_x, tiles_y);
            ptx::mbar_arrive_expect_tx(&full[s], kTileBytes);
            ptx::tma_load_3d(s_in + s * (kTileBytes / 4), &pix_map, &full[s], c.tx * kTW, c.ty * kTH, c.n * 3 + c.g);
        }
    }
    const int bi = tid / kTBW, bj = tid % kTBW;
    for (int it = 0; it < n_my; ++it) {
        const int s = it % kPixStages;
        const int os = it % kOutStages;
        ptx::mbar_wait(&full[s], (it / kPixStages) & 1);
        const float* tin = s_in + s * (kTileBytes / 4) + (4 * bi) * kTW + 4 * bj;
        float v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float4 q = *reinterpret_cast<const float4*>(tin + r * kTW);
            v[r][0] = q.x, v[r][1] = q.y, v[r][2] = q.z, v[r][3] = q.w;
        }
        if (kPre) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) v[r][c] = pre_2xm1(v[r][c]);
        }
        float o[16];
        wht16_fwd(v, o);
        // the store that last used s_out[os] (iteration it - kOutStages) must have finished reading smem
        if (tid == 0) ptx::bulk_wait_group_read<kOutStages - 1>();
        __syncthreads();
        float* tout = s_out + os * (kTileBytes / 4) + bi * kTBW + bj;
#pragma unroll
        for (int k = 0; k < 16; ++k) tout[k * (kTBH * kTBW)] = o[k];
        ptx::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            TileCoord c = tile_coord(first + it * stride, tiles_x, tiles_y);
            ptx::tma_store_5d(&sub_map, s_out + os * (kTileBytes / 4), c.tx * kTBW, c.ty * kTBH, c.g, 0, c.n);
            ptx::bulk_commit_group();
            if (it + kPixStages < n_my) {
                TileCoord d = tile_coord(first + (it + kPixStages) * stride, tiles_x, tiles_y);
                ptx::mbar_arrive_expect_tx(&full[s], kTileBytes);
                ptx::tma_load_3d(s_in + s * (kTileBytes / 4), &pix_map, &full[s], d.tx * kTW, d.ty * kTH,
                                 d.n * 3 + d.g);
            }
        }
    }
    if (tid == 0) ptx::bulk_wait_group_read<0>();
}

template <bool kPost>
__global__ void __launch_bounds__(kThreads) iwt4x4_tma_kernel(const __grid_constant__ CUtensorMap sub_map,
                                                              const __grid_constant__ CUtensorMap pix_map, int tiles_x,
                                                              int tiles_y, int ntiles) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    float* s_in = reinterpret_cast<float*>(smem);                               // [kPixStages][16][kTBH][kTBW]
    float* s_out = reinterpret_cast<float*>(smem + kPixStages * kTileBytes);    // [kOutStages][kTH][kTW]
    __shared__ __align__(8) uint64_t full[kPixStages];

    const int tid = threadIdx.x;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_my = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        ptx::prefetch_tmap(&pix_map);
        ptx::prefetch_tmap(&sub_map);
        for (int s = 0; s < kPixStages; ++s) ptx::mbar_init(&full[s], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < kPixStages && s < n_my; ++s) {
            TileCoord c = tile_coord(first + s * stride, tiles_x, tiles_y);
            ptx::mbar_arrive_expect_tx(&full[s], kTileBytes);
            ptx::tma_load_5d(s_in + s * (kTileBytes / 4), &sub_map, &full[s], c.tx * kTBW, c.ty * kTBH, c.g, 0, c.n);
        }
    }
    const int bi = tid / kTBW, bj = tid % kTBW;
    for (int it = 0; it < n_my; ++it) {
        const int s = it % kPixStages;
        const int os = it % kOutStages;
        ptx::mbar_wait(&full[s], (it / kPixStages) & 1);
        const float* tin = s_in + s * (kTileBytes / 4) + bi * kTBW + bj;
        float in[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) in[k] = tin[k * (kTBH * kTBW)];
        float v[4][4];
        wht16_inv(in, v);
        if (tid == 0) ptx::bulk_wait_group_read<kOutStages - 1>();
        __syncthreads();
        float* tout = s_out + os * (kTileBytes / 4) + (4 * bi) * kTW + 4 * bj;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float4 q;
            if (kPost)
                q = make_float4(post_clamp(v[r][0]), post_clamp(v[r][1]), post_clamp(v[r][2]), post_clamp(v[r][3]));
            else
                q = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
            *reinterpret_cast<float4*>(tout + r * kTW) = q;
        }
        ptx::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            TileCoord c = tile_coord(first + it * stride, tiles_x, tiles_y);
            ptx::tma_store_3d(&pix_map, s_out + os * (kTileBytes / 4), c.tx * kTW, c.ty * kTH, c.n * 3 + c.g);
            ptx::bulk_commit_group();
            if (it + kPixStages < n_my) {
                TileCoord d = tile_coord(first + (it + kPixStages) * stride, tiles_x, tiles_y);
                ptx::mbar_arrive_expect_tx(&full[s], kTileBytes);
                ptx::tma_load_5d(s_in + s * (kTileBytes / 4), &sub_map, &full[s], d.tx * kTBW, d.ty * kTBH, d.g, 0,
                                 d.n);
            }
        }
    }
    if (tid == 0) ptx::bulk_wait_group_read<0>();
}

// ---- host side ----------------------------------------------------------------------------------------
int num_sms() {
    static int sms = []() {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        return v;
    }();
    return sms;
}

bool tma_ok(int n, int H, int W) { return (W % 16 == 0) && W >= kTW && H >= kTH && n > 0; }

int make_wt_maps(CUtensorMap* pix, CUtensorMap* sub, const float* x, const float* y, int n, int H, int W) {
    const uint64_t h = H / 4, w = W / 4;
    uint64_t pd[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)n * 3};
    uint64_t ps[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
    uint32_t pb[3] = {kTW, kTH, 1};
    int r = make_tmap(pix, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, pd, ps, pb, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r) return r;
    uint64_t sd[5] = {w, h, 3, 16, (uint64_t)n};
    uint64_t ss[4] = {w * 4, h * w * 4, 3 * h * w * 4, 48 * h * w * 4};
    uint32_t sb[5] = {kTBW, kTBH, 1, 16, 1};
    return make_tmap(sub, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, y, sd, ss, sb, CU_TENSOR_MAP_SWIZZLE_NONE);
}

template <typename K>
int set_smem(K kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    return e == cudaSuccess ? WDM_OK : wdm_cuda_error((int)e);
}

int wt_common_checks(const void* a, const void* b, int n, int H, int W, int flags, int flag_mask) {
    if (n < 0 || H < 0 || W < 0) return WDM_ERR_BAD_SHAPE;
    if ((!a || !b) && n != 0 && H != 0 && W != 0) return WDM_ERR_BAD_ARG;
    if ((H % 4) || (W % 4)) return WDM_ERR_BAD_SHAPE;
    if (flags & ~(flag_mask | WDM_WT_IMPL_MASK)) return WDM_ERR_BAD_ARG;
    const int impl = flags & WDM_WT_IMPL_MASK;
    if (impl != WDM_WT_IMPL_AUTO && impl != WDM_WT_IMPL_DIRECT && impl != WDM_WT_IMPL_TMA) return WDM_ERR_BAD_ARG;
    if (!wdm_aligned(a, 16) || !wdm_aligned(b, 16)) return WDM_ERR_BAD_ALIGN;
    return WDM_OK;
}

}  // namespace

extern "C" int wdm_dwt4x4_fwd(const float* x, float* y, int n, int H, int W, int flags, void* stream_) {
    int st = wt_common_checks(x, y, n, H, W, flags, WDM_DWT_PRE_2XM1);
    if (st != WDM_OK) return st;
    if (n == 0 || H == 0 || W == 0) return WDM_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool pre = flags & WDM_DWT_PRE_2XM1;
    int impl = flags & WDM_WT_IMPL_MASK;
    if (impl == WDM_WT_IMPL_TMA && !tma_ok(n, H, W)) return WDM_ERR_UNSUPPORTED;
    if (impl == WDM_WT_IMPL_AUTO) impl = WDM_WT_IMPL_DIRECT;
    if (impl == WDM_WT_IMPL_TMA) {
        CUtensorMap pix, sub;
        int r = make_wt_maps(&pix, &sub, x, y, n, H, W);
        if (r) return r < 0 ? WDM_ERR_UNSUPPORTED : wdm_cuda_error(r);
        const int tiles_x = wdm_cdiv(W, kTW), tiles_y = wdm_cdiv(H, kTH);
        const long long ntiles = (long long)tiles_x * tiles_y * n * 3;
        if (ntiles > 0x7fffffffLL) return WDM_ERR_BAD_SHAPE;
        const int grid = (int)(ntiles < 2LL * num_sms() ? ntiles : 2LL * num_sms());
        if (pre) {
            if ((st = set_smem(dwt4x4_tma_kernel<true>)) != WDM_OK) return st;
            dwt4x4_tma_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(pix, sub, tiles_x, tiles_y, (int)ntiles);
        } else {
            if ((st = set_smem(dwt4x4_tma_kernel<false>)) != WDM_OK) return st;
            dwt4x4_tma_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(pix, sub, tiles_x, tiles_y, (int)ntiles);
        }
        return wdm_launch_status();
    }
    const int h = H / 4, w = W / 4;
    const long long nblocks = (long long)n * 3 * h * w;
    const long long grid = (nblocks + 255) / 256;
    if (grid > 0x7fffffffLL) return WDM_ERR_BAD_SHAPE;
    if (pre)
        dwt4x4_direct_kernel<true><<<(unsigned)grid, 256, 0, stream>>>(x, y, nblocks, h, w);
    else
        dwt4x4_direct_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(x, y, nblocks, h, w);
    return wdm_launch_status();
}

extern "C" int wdm_iwt4x4_fwd(const float* y, float* x, int n, int h, int w, int flags, void* stream_) {
    if (h > (1 << 28) || w > (1 << 28)) return WDM_ERR_BAD_SHAPE;
    const int H = 4 * h, W = 4 * w;
    int st = wt_common_checks(y, x, n, H, W, flags, WDM_IWT_POST_CLAMP);
    if (st != WDM_OK) return st;
    if (n == 0 || h == 0 || w == 0) return WDM_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const bool post = flags & WDM_IWT_POST_CLAMP;
    int impl = flags & WDM_WT_IMPL_MASK;
    if (impl == WDM_WT_IMPL_TMA && !tma_ok(n, H, W)) return WDM_ERR_UNSUPPORTED;
    if (impl == WDM_WT_IMPL_AUTO) impl = WDM_WT_IMPL_DIRECT;
    if (impl == WDM_WT_IMPL_TMA) {
        CUtensorMap pix, sub;
        int r = make_wt_maps(&pix, &sub, x, y, n, H, W);
        if (r) return r < 0 ? WDM_ERR_UNSUPPORTED : wdm_cuda_error(r);
        const int tiles_x = wdm_cdiv(W, kTW), tiles_y = wdm_cdiv(H, kTH);
        const long long ntiles = (long long)tiles_x * tiles_y * n * 3;
        if (ntiles > 0x7fffffffLL) return WDM_ERR_BAD_SHAPE;
        const int grid = (int)(ntiles < 2LL * num_sms() ? ntiles : 2LL * num_sms());
        if (post) {
            if ((st = set_smem(iwt4x4_tma_kernel<true>)) != WDM_OK) return st;
            iwt4x4_tma_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(sub, pix, tiles_x, tiles_y, (int)ntiles);
        } else {
            if ((st = set_smem(iwt4x4_tma_kernel<false>)) != WDM_OK) return st;
            iwt4x4_tma_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(sub, pix, tiles_x, tiles_y, (int)ntiles);
        }
        return wdm_launch_status();
    }
    const long long nblocks = (long long)n * 3 * h * w;
    const long long grid = (nblocks + 255) / 256;
    if (grid > 0x7fffffffLL) return WDM_ERR_BAD_SHAPE;
    if (post)
        iwt4x4_direct_kernel<true><<<(unsigned)grid, 256, 0, stream>>>(y, x, nblocks, h, w);
    else
        iwt4x4_direct_kernel<false><<<(unsigned)grid, 256, 0, stream>>>(y, x, nblocks, h, w);
    return wdm_launch_status();
}


// ---- wavelet_in_unet entry points ---------------------------------------------------------------------
namespace wdm {
int launch_dwt_gather(const float* src0, const float* src1, int nsrc, int B, int H, int W, const int* patches, int P, int R,
                      int Cpad, void* out, int out_dtype, cudaStream_t s) {
    if (P <= 0) return WDM_OK;
    if (!src0 || nsrc < 1 || nsrc > 2 || (nsrc == 2 && !src1) || !patches || !out) return WDM_ERR_BAD_ARG;
    if (R <= 0 || 4 * R > H || 4 * R > W || Cpad < nsrc * 48 || (Cpad % 8) || B <= 0) return WDM_ERR_BAD_SHAPE;
    if (!wdm_aligned(out, 16)) return WDM_ERR_BAD_ALIGN;
    const size_t wpr = out_dtype == 0 ? (size_t)Cpad : (size_t)Cpad / 2;
    const size_t smem = (size_t)R * (wpr | 1) * 4;
    if (smem > 96 * 1024) return WDM_ERR_BAD_SHAPE;
    dim3 grid(R, P);
    if (out_dtype == 0) {
        cudaFuncSetAttribute(dwt_gather_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        dwt_gather_kernel<float><<<grid, 384, smem, s>>>(src0, src1, nsrc, H, W, patches, R, Cpad, reinterpret_cast<float*>(out));
    } else {
        cudaFuncSetAttribute(dwt_gather_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        dwt_gather_kernel<__nv_bfloat16><<<grid, 384, smem, s>>>(src0, src1, nsrc, H, W, patches, R, Cpad,
                                                                  reinterpret_cast<__nv_bfloat16*>(out));
    }
    return wdm_launch_status();
}

int launch_iwt_nhwc(const float* y, int ld, int P, int R, float* x, cudaStream_t s) {
    if (P <= 0) return WDM_OK;
    if (!y || !x || ld < 48 || (ld & 3) || R <= 0) return WDM_ERR_BAD_ARG;
    if (!wdm_aligned(x, 16) || !wdm_aligned(y, 16)) return WDM_ERR_BAD_ALIGN;
    const size_t smem = (size_t)R * (ld | 1) * sizeof(float);
    if (smem > 96 * 1024) return WDM_ERR_BAD_SHAPE;
    cudaFuncSetAttribute(iwt_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    iwt_nhwc_kernel<<<dim3(R, P), 192, smem, s>>>(y, ld, R, x);
    return wdm_launch_status();
}
}  // namespace wdm

extern "C" int wdm_gather_patches_dwt(const float* src0, const float* src1, int nsrc, int B, int H, int W, const int* patches,
                                      int P, int R, int Cpad, void* out, int out_dtype, void* stream) {
    if (out_dtype != WDM_PREC_FP32 && out_dtype != WDM_PREC_BF16) return WDM_ERR_BAD_ARG;
    return wdm::launch_dwt_gather(src0, src1, nsrc, B, H, W, patches, P, R, Cpad, out, out_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_iwt4x4_nhwc(const float* y, int ld, int P, int R, float* x, void* stream) {
    return wdm::launch_iwt_nhwc(y, ld, P, R, x, static_cast<cudaStream_t>(stream));
}

// lo: [n, Clo, h, w] supplies sub-band channels [0, Clo); hi: [n, Chi, h, w] supplies the rest -- Chi == 48 (a full
// wavelet tensor, its first Clo channels are skipped) or Chi == 48 - Clo (only the remaining bands).
extern "C" int wdm_iwt4x4_cat(const float* lo, int Clo, const float* hi, int Chi, float* x, int n, int h, int w, int flags,
                              void* stream) {
    if (!lo || !hi || !x) return WDM_ERR_BAD_ARG;
    if (flags & ~WDM_IWT_POST_CLAMP) return WDM_ERR_BAD_ARG;
    if (n < 0 || h <= 0 || w <= 0 || Clo < 1 || Clo > 47 || (Chi != 48 && Chi != 48 - Clo)) return WDM_ERR_BAD_SHAPE;
    if (!wdm_aligned(x, 16)) return WDM_ERR_BAD_ALIGN;
    if (n == 0) return WDM_OK;
    const long long nblocks = (long long)n * 3 * h * w;
    const unsigned grid = (unsigned)((nblocks + 255) / 256);
    const int hi_off = Chi == 48 ? 0 : Clo;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (flags & WDM_IWT_POST_CLAMP)
        iwt4x4_cat_kernel<true><<<grid, 256, 0, s>>>(lo, Clo, hi, Chi, hi_off, x, nblocks, h, w);
    else
        iwt4x4_cat_kernel<false><<<grid, 256, 0, s>>>(lo, Clo, hi, Chi, hi_off, x, nblocks, h, w);
    return wdm_launch_status();
}
