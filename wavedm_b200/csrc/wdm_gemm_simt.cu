// wdm_gemm_simt.cu -- CUDA-core (FFMA) implicit-GEMM: the fp32 "parity mode" contraction kernel, and the
// generic fallback for shapes the tcgen05 kernel does not tile. fp32 accumulate, fp32 or bf16 storage.
//
// Covers, through GemmParams (wdm_engine.h):
//   3x3 s1 conv (models/unet.py:91,100,45,233), 3x3 s2 conv with right/bottom zero pad (unet.py:65-75),
//   1x1 conv (unet.py:113,147-162), nearest-x2 upsample folded into the conv addressing (unet.py:52-53),
//   channel concat of two sources without a copy (unet.py:379-380), attention matmuls q.k^T and p.v
//   (unet.py:176-189) as batched GEMMs; epilogue bias + temb row + residual (unet.py:125,138,193).
// Tile 128x128x16, 256 threads, 8x8 outputs per thread (split 4+4 to keep LDS.128 conflict-free),
// register-prefetch double buffering.
#include "wdm_common.cuh"
#include "wdm_engine.h"

namespace wdm {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PITCH = 132;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 q;
    q.x = *reinterpret_cast<uint32_t*>(&lo);
    q.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = q;
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 q = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(q.x << 16), v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16), v[3] = __uint_as_float(q.y & 0xffff0000u);
}

// plain (non-__ldg) 8-element loads, safe for shared memory
__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

template <typename TA, typename TB, typename TO, int kBLayout>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[2][BK][PITCH];
    __shared__ __align__(16) float Bs[2][BK][PITCH];

    const int tid = threadIdx.x;
    const int rows_per_batch = p.Hout * p.Wout;
    // M tiling: tiles never straddle a batch when B is per-batch
    const int group_rows = p.b_batch_stride ? rows_per_batch : p.M;
    const int tiles_per_group = (group_rows + BM - 1) / BM;
    const int group = blockIdx.x / tiles_per_group;
    const int tile_row0 = (blockIdx.x % tiles_per_group) * BM;
    const long long m_base = (long long)group * group_rows + tile_row0;
    const int n0 = blockIdx.y * BN;

    // ---- A loader: thread -> (row lr, 8-channel half kh)
    const int lr = tid & 127, kh = tid >> 7;
    const bool a_row_ok = (tile_row0 + lr) < group_rows;
    const long long am = m_base + lr;
    const int ab = (a_row_ok && !p.a_shared) ? (int)(am / rows_per_batch) : 0;
    const int ar = a_row_ok ? (int)(am % rows_per_batch) : 0;
    const int aoy = ar / p.Wout, aox = ar % p.Wout;
    const int Ct = p.tail_1x1 ? p.C0 : p.C0 + p.C1;        // channels per tap of the main part
    const int cblocks = Ct / BK;
    const int main_kb = p.taps * cblocks;
    const int nkb = main_kb + (p.tail_1x1 ? (p.C1 + p.C2) / BK : 0);
    const int Hv = p.ups ? 2 * p.Hin : p.Hin, Wv = p.ups ? 2 * p.Win : p.Win;  // virtual (post-upsample) size

    // ---- B loader
    // BL_NK: thread -> (n row = tid & 127, k half = tid >> 7), 8 contiguous k
    // BL_KN: thread -> (k row = tid >> 4, n chunk = (tid & 15) * 8), 8 contiguous n
    const TB* Bbase = reinterpret_cast<const TB*>(p.B) + (p.b_batch_stride ? (long long)group * p.b_batch_stride : 0);

    float ra[8], rb[8];
    auto load_tiles = [&](int kb) {
        const bool tail = kb >= main_kb;
        const int tap = tail ? 0 : kb / cblocks;
        const int c = tail ? (kb - main_kb) * BK : (kb - tap * cblocks) * BK;
        // A
        bool ok = a_row_ok;
        int iy = aoy, ix = aox;
        if (tail) {
            // 1x1 tail sources: centre pixel of the output position
        } else if (p.taps == 9) {
            iy = aoy * p.stride + tap / 3 - p.pad;
            ix = aox * p.stride + tap % 3 - p.pad;
            ok = ok && iy >= 0 && iy < Hv && ix >= 0 && ix < Wv;
        } else if (p.stride != 1) {
            iy = aoy * p.stride;
            ix = aox * p.stride;
        }
        if (p.ups) iy >>= 1, ix >>= 1;
        if (ok) {
            const long long pix = ((long long)ab * p.Hin + iy) * p.Win + ix;
            if (tail) {
                if (c < p.C1)
                    load8<TA>(reinterpret_cast<const TA*>(p.src1) + pix * p.ld1 + c + kh * 8, ra);
                else
                    load8<TA>(reinterpret_cast<const TA*>(p.src2) + pix * p.ld2 + (c - p.C1) + kh * 8, ra);
            } else if (c < p.C0)
                load8<TA>(reinterpret_cast<const TA*>(p.src0) + pix * p.ld0 + c + kh * 8, ra);
            else
                load8<TA>(reinterpret_cast<const TA*>(p.src1) + pix * p.ld1 + (c - p.C0) + kh * 8, ra);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = 0.f;
        }
        // B
        const int k0 = tail ? p.taps * Ct + c : tap * Ct + c;
        if (kBLayout == BL_NK) {
            const int n = n0 + (tid & 127);
            if (n < p.N)
                load8<TB>(Bbase + (long long)n * p.ldb + k0 + (tid >> 7) * 8, rb);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) rb[i] = 0.f;
            }
        } else {
            const int n = n0 + (tid & 15) * 8;
            if (n < p.N)  // N is a multiple of 8 in every caller
                load8<TB>(Bbase + (long long)(k0 + (tid >> 4)) * p.ldb + n, rb);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) rb[i] = 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][kh * 8 + i][lr] = ra[i];
        if (kBLayout == BL_NK) {
#pragma unroll
            for (int i = 0; i < 8; ++i) Bs[buf][(tid >> 7) * 8 + i][tid & 127] = rb[i];
        } else {
            float* d = &Bs[buf][tid >> 4][(tid & 15) * 8];
            *reinterpret_cast<float4*>(d) = make_float4(rb[0], rb[1], rb[2], rb[3]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(rb[4], rb[5], rb[6], rb[7]);
        }
    };

    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 thread grid
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nkb) load_tiles(kb + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            a[0] = a0.x, a[1] = a0.y, a[2] = a0.z, a[3] = a0.w, a[4] = a1.x, a[5] = a1.y, a[6] = a1.z, a[7] = a1.w;
            b[0] = b0.x, b[1] = b0.y, b[2] = b0.z, b[3] = b0.w, b[4] = b1.x, b[5] = b1.y, b[6] = b1.z, b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kb + 1 < nkb) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue
    TO* out = reinterpret_cast<TO*>(p.out);
    const TO* res = reinterpret_cast<const TO*>(p.residual);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
        if (tile_row0 + rl >= group_rows) continue;
        const long long m = m_base + rl;
        const int trow = (p.temb && p.temb_rows > 1) ? (int)(m / rows_per_batch) : 0;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + jh * 64 + tx * 4;
            if (n >= p.N) continue;  // N is a multiple of 4 in every caller
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][jh * 4 + j] * p.alpha;
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += __ldg(p.bias + n + j);
            }
            if (p.temb) {
                const float* tr = p.temb + (long long)trow * p.temb_ld + n;
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += __ldg(tr + j);
            }
            if (res) {
                float r4[4];
                load4(res + m * p.ldr + n, r4);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += r4[j];
            }
            store4(out + m * p.ldo + n, v[0], v[1], v[2], v[3]);
        }
    }
}

template <typename TA, typename TB, typename TO>
int launch_t(const GemmParams& p, cudaStream_t s) {
    const int rows_per_batch = p.Hout * p.Wout;
    const int group_rows = p.b_batch_stride ? rows_per_batch : p.M;
    const int ngroups = p.b_batch_stride ? p.M / rows_per_batch : 1;
    dim3 grid(((group_rows + BM - 1) / BM) * ngroups, (p.N + BN - 1) / BN);
    if (p.b_layout == BL_NK)
        gemm_simt_kernel<TA, TB, TO, BL_NK><<<grid, 256, 0, s>>>(p);
    else
        gemm_simt_kernel<TA, TB, TO, BL_KN><<<grid, 256, 0, s>>>(p);
    return wdm_launch_status();
}

// ---- conv_out: Cout <= 4, NHWC in, NCHW fp32 out -------------------------------------------------------
// One CTA per (patch, strip of kStripRows output rows): a rolling window of three input rows lives in shared memory
// (each input row is loaded once per strip, coalesced 16-byte loads, zero rows/columns for the padding), weights
// [Cout][9][C] in shared memory as fp32. Thread = (4 adjacent pixels, one 8-channel slice); the C/8 slices of a pixel
// group sit in adjacent lanes and are shuffle-reduced.
constexpr int kStripRows = 8;

template <typename T>
__global__ void __launch_bounds__(256, 2) conv_small_cout_kernel(const T* __restrict__ src, int P, int H, int W, int C,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias, int Cout,
                                                                 float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_cs[];
    const int y0 = blockIdx.x * kStripRows, p = blockIdx.y;
    const int Wp = W + 2;
    T* rows = reinterpret_cast<T*>(smem_cs);                                         // ring [3][W+2][C]
    float* sw = reinterpret_cast<float*>(smem_cs + (size_t)3 * Wp * C * sizeof(T));  // [Cout][9][C]
    for (int i = threadIdx.x; i < Cout * 9 * C; i += blockDim.x) sw[i] = w[i];
    constexpr int VE = 16 / sizeof(T);  // elements per 16-byte vector
    const int vec_per_row = Wp * C / VE;
    auto load_row = [&](int iy) {  // input row iy -> ring slot (iy + 3) % 3
        uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)((iy + 3) % 3) * Wp * C);
        for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
            const int px = (v * VE) / C - 1, c = (v * VE) % C;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (iy >= 0 && iy < H && px >= 0 && px < W)
                q = __ldg(reinterpret_cast<const uint4*>(src + (((long long)p * H + iy) * W + px) * C + c));
            dst[v] = q;
        }
    };
    load_row(y0 - 1);
    load_row(y0);
    const int nslice = C / 8;  // 16 for C = 128 (must divide 32)
    const int s = threadIdx.x % nslice;
    const int groups_per_pass = blockDim.x / nslice;
    const int yend = min(H, y0 + kStripRows);
    for (int y = y0; y < yend; ++y) {
        load_row(y + 1);
        __syncthreads();
        for (int pg = threadIdx.x / nslice; pg * 4 < W; pg += groups_per_pass) {
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int o = 0; o < 4; ++o) acc[a][o] = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap % 3;
                const T* rrow = rows + (size_t)((y + dy - 1 + 3) % 3) * Wp * C;
                float wv[4][8];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    if (o < Cout) {
                        const float4 w0 = *reinterpret_cast<const float4*>(sw + (o * 9 + tap) * C + s * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(sw + (o * 9 + tap) * C + s * 8 + 4);
                        wv[o][0] = w0.x, wv[o][1] = w0.y, wv[o][2] = w0.z, wv[o][3] = w0.w;
                        wv[o][4] = w1.x, wv[o][5] = w1.y, wv[o][6] = w1.z, wv[o][7] = w1.w;
                    }
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int px = pg * 4 + a + dx;  // index in the padded row (pixel x + dx - 1 + 1)
                    float v[8];
                    lds8(rrow + (size_t)px * C + s * 8, v);
#pragma unroll
                    for (int o = 0; o < 4; ++o)
                        if (o < Cout) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) acc[a][o] = fmaf(v[i], wv[o][i], acc[a][o]);
                        }
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (o < Cout) {
                        for (int off = nslice >> 1; off; off >>= 1)
                            acc[a][o] += __shfl_xor_sync(0xffffffffu, acc[a][o], off);
                    }
            if (s == 0) {
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (o < Cout) {
                        float* d = out + (((long long)p * Cout + o) * H + y) * W + pg * 4;
                        const float b = bias ? bias[o] : 0.f;
                        *reinterpret_cast<float4*>(d) =
                            make_float4(acc[0][o] + b, acc[1][o] + b, acc[2][o] + b, acc[3][o] + b);
                    }
            }
        }
        __syncthreads();  // the next iteration overwrites the slot of row y - 1
    }
}

}  // namespace

int launch_gemm_simt(const GemmParams& p, cudaStream_t s) {
    if (p.row_scale || p.row_scale_out || p.fuse_softmax) return WDM_ERR_UNSUPPORTED;  // tensor-core epilogue features
    if (p.M <= 0 || p.N <= 0) return WDM_OK;
    if ((p.C0 % BK) || (p.C1 % BK) || (p.N % 8)) return WDM_ERR_BAD_SHAPE;
    if (p.tail_1x1) {
        if ((p.C2 % BK) || p.K != p.taps * p.C0 + p.C1 + p.C2 || p.stride != 1 || p.ups || (p.C2 && !p.src2) || !p.src1)
            return WDM_ERR_BAD_SHAPE;
    } else if (p.K != p.taps * (p.C0 + p.C1)) {
        return WDM_ERR_BAD_SHAPE;
    }
    if (p.taps != 1 && p.taps != 9) return WDM_ERR_BAD_SHAPE;
    if (p.a_dtype == DT_F32 && p.b_dtype == DT_F32 && p.out_dtype == DT_F32) return launch_t<float, float, float>(p, s);
    if (p.a_dtype == DT_BF16 && p.b_dtype == DT_BF16 && p.out_dtype == DT_BF16)
        return launch_t<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(p, s);
    if (p.a_dtype == DT_BF16 && p.b_dtype == DT_BF16 && p.out_dtype == DT_F32)
        return launch_t<__nv_bfloat16, __nv_bfloat16, float>(p, s);
    return WDM_ERR_UNSUPPORTED;
}

int launch_conv_small_cout(const void* src, int dtype, int P, int H, int W, int C, const float* w, const float* bias,
                           int Cout, float* out, cudaStream_t s) {
    const int nslice = C / 8;
    if (Cout < 1 || Cout > 4 || (C % 8) || nslice > 32 || (32 % nslice) || (W % 4)) return WDM_ERR_BAD_SHAPE;
    const size_t es = dtype == DT_F32 ? 4 : 2;
    const size_t smem = (size_t)3 * (W + 2) * C * es + (size_t)Cout * 9 * C * sizeof(float);
    if (smem > 200 * 1024) return WDM_ERR_BAD_SHAPE;
    dim3 grid((H + kStripRows - 1) / kStripRows, P);
    if (dtype == DT_F32) {
        cudaFuncSetAttribute(conv_small_cout_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_small_cout_kernel<float><<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(src), P, H, W, C, w, bias,
                                                               Cout, out);
    } else {
        cudaFuncSetAttribute(conv_small_cout_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
        conv_small_cout_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), P,
                                                                       H, W, C, w, bias, Cout, out);
    }
    return wdm_launch_status();
}

}  // namespace wdm
