// wdm_hfrm.cu -- the HFRM engine: the one-shot high-frequency refinement CNN that restore() runs once per image before
// the sampling loop (reference models/arch.py:158-253, call site models/restoration.py:94). SURVEY.md 8(f)-1.
//
// Activations are NHWC ([B, H, W, C], channels innermost) in the engine's storage type: bf16 (throughput mode: warp-level
// tensor-core MMAs, fp32 accumulate) or fp32 (parity mode: FFMA). A ResidualBlock (arch.py:158-204) is SIX launches:
//   1. pw  : LayerNorm2d (norm1, folded) -> conv1 1x1 C->2C                            -> t1 [px, 2C]
//   2. dw  : conv2 depthwise 3x3 + SimpleGate (x1*x2) + per-tile channel sums          -> t2 [px, C], partial sums
//   3. chan: global average pool + chan_conv 1x1 (channel attention weights)           -> s  [B, C]
//   4. pw  : (t2 * s) -> conv3 1x1 (beta folded) + residual                            -> x1 = x + beta*conv3(.)
//   5. pw  : LayerNorm2d (norm2, folded) -> conv4 1x1 C->2C -> SimpleGate in the epilogue -> t3 [px, C]
//   6. pw  : conv5 1x1 (gamma folded) + residual                                       -> y = x1 + gamma*conv5(.)
// "Folded" = done once at pack time: the LayerNorm affine goes into the following 1x1 conv (W' = W diag(ln_w),
// b' = b + W ln_b), beta / gamma scale the rows of conv3 / conv5, the SimpleGate partners (c, c+C) of conv4 are interleaved
// in 16-row groups so both land in the same thread's accumulators, and the PixelShuffle of the up path is a row permutation
// of the 1x1 conv + a strided store. The HBM-bound part (C = 32 / 64 at full / half resolution) therefore moves 14 C bytes
// per pixel and block instead of the 34 C of the op-by-op form.
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "wdm_common.cuh"
#include "wdm_engine.h"

namespace wdm {
namespace {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ 8-element vectors
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f(float x) { return x; }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }

// ------------------------------------------------------------------------------------------------ pointwise (1x1) conv
// out = epilogue( prologue(A)[M x K] . W[N x K]^T ), one CTA = 128 pixels x 64 GEMM columns, K in chunks of 32 through
// shared memory (register prefetch of the next chunk). bf16: mma.sync.m16n8k16 (each warp 16 rows x 64 columns), fp32: FFMA
// (each thread 8 rows x 4 columns). Both give a thread, per 16-column group, the columns (2s, 2s+1) and (8+2s, 9+2s): with
// the gate-interleaved row order of conv4 those are the SimpleGate partners.
enum { PRO_NONE = 0, PRO_LN = 1, PRO_SCALE = 2 };
enum { EPI_BIAS = 0, EPI_GATE = 1, EPI_RES = 2, EPI_SHUFFLE = 3 };
enum { AM_PLAIN = 0, AM_S2D = 1 };

struct PwParams {
    const void* A;
    int lda;
    int M, K, N;
    int a_mode;      // AM_S2D: rows run over the OUTPUT grid [B, H, W] of a 2x2 stride-2 conv, K = 4*Cin, k = (dy*2+dx)*Cin + c
    int H, W;        // AM_S2D: output grid; EPI_SHUFFLE: input grid (the output is 2H x 2W); PRO_SCALE: H*W pixels per image
    int Cin;
    const void* Wt;  // [N][K] storage type
    const float* bias;
    const float* wsum;  // ring kernel, PRO_LN: sum_k Wt[n][k] (of the rounded storage values), see hfrm_pw_ring_kernel
    int pro;
    float eps;
    const float* scale;  // PRO_SCALE: [B][K]
    int epi;
    const void* res;  // EPI_RES: [M][ldr]; EPI_SHUFFLE: the skip tensor on the output grid
    int ldr;
    void* out;
    int ldo;
};

constexpr int kPwTM = 128, kPwTN = 64, kPwTK = 32, kPwThreads = 256;
constexpr int kCsLd = 65;
template <typename T>
struct PwLd {
    static constexpr int v = std::is_same<T, float>::value ? 33 : 40;  // smem row pitch in elements
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__global__ void __launch_bounds__(kPwThreads) hfrm_pw_kernel(const PwParams p) {
    constexpr bool kF32 = std::is_same<T, float>::value;
    constexpr int LD = PwLd<T>::v;
    constexpr int kABytes = kPwTM * LD * (int)sizeof(T), kWBytes = kPwTN * LD * (int)sizeof(T);
    constexpr int kCsBytes = kPwTM * kCsLd * 4;
    constexpr int kSm = kABytes + kWBytes > kCsBytes ? kABytes + kWBytes : kCsBytes;
    __shared__ __align__(16) unsigned char smem[kSm];
    __shared__ float mu_s[kPwTM], rstd_s[kPwTM];
    T* As = reinterpret_cast<T*>(smem);
    T* Ws = reinterpret_cast<T*>(smem + kABytes);
    float* Cs = reinterpret_cast<float*>(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.x * kPwTM;
    const int n0 = blockIdx.y * kPwTN;
    const T* A = reinterpret_cast<const T*>(p.A);
    const T* Wt = reinterpret_cast<const T*>(p.Wt);

    // ---- LayerNorm2d statistics of this CTA's rows (arch.py:6-17: biased variance over the channels of one pixel).
    // A row is K/8 16-byte vectors: min(32, K/8) lanes per row, 32/that rows per pass, the row stays in registers between the
    // mean and the variance pass (K <= 512; longer rows are re-read).
    if (p.pro == PRO_LN) {
        const int lpr = p.K >= 256 ? 32 : (p.K >> 3);  // lanes per row: 4 / 8 / 16 / 32 (K is a power-of-two multiple of 32)
        const int rpp = 32 / lpr, sub = lane / lpr, sl = lane - sub * lpr;
        const int nvec = p.K / (8 * lpr);
        for (int r0 = 0; r0 < 16; r0 += rpp) {
            const int r = warp * 16 + r0 + sub;
            const long long m = m0 + r;
            const bool ok = m < p.M;
            const T* row = A + (ok ? m : 0) * p.lda;
            float c0[8], c1[8];
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) c0[j] = c1[j] = 0.f;
            if (ok) {
                load8(row + sl * 8, c0);
                if (nvec > 1) load8(row + (lpr + sl) * 8, c1);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += c0[j] + c1[j];
                for (int i = 2; i < nvec; ++i) {
                    float v[8];
                    load8(row + (i * lpr + sl) * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) s += v[j];
                }
            }
            for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)p.K;
            float q = 0.f;
            if (ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d0 = c0[j] - mean;
                    q = fmaf(d0, d0, q);
                }
                if (nvec > 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float d1 = c1[j] - mean;
                        q = fmaf(d1, d1, q);
                    }
                }
                for (int i = 2; i < nvec; ++i) {
                    float v[8];
                    load8(row + (i * lpr + sl) * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float d = v[j] - mean;
                        q = fmaf(d, d, q);
                    }
                }
            }
            for (int o = lpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            if (sl == 0) {
                mu_s[r] = ok ? mean : 0.f;
                rstd_s[r] = ok ? 1.0f / sqrtf(q / (float)p.K + p.eps) : 0.f;
            }
        }
        __syncthreads();
    }

    // ---- per-thread staging coordinates: A vectors v = tid, tid + 256 (row = v >> 2, 8 elements at (v & 3) * 8), W vector tid
    long long arow[2];   // source row index (elements offset = arow * lda) or -1
    int ar[2], akv[2];
    long long aimg[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int v = tid + i * kPwThreads;
        ar[i] = v >> 2, akv[i] = (v & 3) * 8;
        const long long m = m0 + ar[i];
        arow[i] = -1, aimg[i] = 0;
        if (m < p.M) {
            if (p.a_mode == AM_S2D) {
                const long long hw = (long long)p.H * p.W;
                const long long img = m / hw, rem = m - img * hw;
                const int oy = (int)(rem / p.W), ox = (int)(rem - (long long)oy * p.W);
                arow[i] = (img * 2 * p.H + 2 * oy) * (2 * p.W) + 2 * ox;
            } else {
                arow[i] = m;
            }
            if (p.pro == PRO_SCALE) aimg[i] = m / ((long long)p.H * p.W);
        }
    }
    const int wr = tid >> 2, wkv = (tid & 3) * 8;
    const bool wvalid = n0 + wr < p.N;

    float areg[2][8], wreg[8];
    auto fetch = [&](int kc) {
        const int k0 = kc * kPwTK;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (arow[i] >= 0) {
                long long row = arow[i];
                int k = k0 + akv[i];
                if (p.a_mode == AM_S2D) {
                    const int tap = k0 / p.Cin;
                    row += (long long)(tap >> 1) * (2 * p.W) + (tap & 1);
                    k -= tap * p.Cin;
                }
                load8(A + row * p.lda + k, areg[i]);
                if (p.pro == PRO_LN) {
                    const float mu = mu_s[ar[i]], rs = rstd_s[ar[i]];
#pragma unroll
                    for (int j = 0; j < 8; ++j) areg[i][j] = (areg[i][j] - mu) * rs;
                } else if (p.pro == PRO_SCALE) {
                    float sc[8];
                    load8(p.scale + aimg[i] * p.K + k0 + akv[i], sc);
#pragma unroll
                    for (int j = 0; j < 8; ++j) areg[i][j] *= sc[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) areg[i][j] = 0.f;
            }
        }
        if (wvalid) {
            load8(Wt + (long long)(n0 + wr) * p.K + k0 + wkv, wreg);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) wreg[j] = 0.f;
        }
    };
    auto stage = [&]() {
        if constexpr (kF32) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) As[ar[i] * LD + akv[i] + j] = areg[i][j];
#pragma unroll
            for (int j = 0; j < 8; ++j) Ws[wr * LD + wkv + j] = wreg[j];
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) store8(As + ar[i] * LD + akv[i], areg[i]);
            store8(Ws + wr * LD + wkv, wreg);
        }
    };

    // accumulators: bf16 path acc[np][h][4] (np = 16-column group, h = half, c0..c3 of the mma); fp32 path acc[i][j]
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;

    const int nk = p.K / kPwTK;
    fetch(0);
    for (int kc = 0; kc < nk; ++kc) {
        stage();
        __syncthreads();
        if (kc + 1 < nk) fetch(kc + 1);
        if constexpr (kF32) {
            const int ty = tid >> 4, tx = tid & 15, q = tx >> 2, s = tx & 3;
            const int nc[4] = {16 * q + 2 * s, 16 * q + 2 * s + 1, 16 * q + 8 + 2 * s, 16 * q + 9 + 2 * s};
#pragma unroll 4
            for (int k = 0; k < kPwTK; ++k) {
                float a[8], b[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = As[(ty * 8 + i) * LD + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Ws[nc[j] * LD + k];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fmaf(a[i], b[j], acc[i * 4 + j]);
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                ldmatrix_x4(a, As + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + ks * 16 + ((lane >> 4) & 1) * 8);
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t b[4];
                    ldmatrix_x4(b, Ws + (np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
                    float(&c0)[4] = *reinterpret_cast<float(*)[4]>(&acc[(np * 2 + 0) * 4]);
                    float(&c1)[4] = *reinterpret_cast<float(*)[4]>(&acc[(np * 2 + 1) * 4]);
                    mma_bf16(c0, a, b[0], b[1]);
                    mma_bf16(c1, a, b[2], b[3]);
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue, stage 1: (+ bias, SimpleGate) -> Cs[128][65] fp32 (aliases the operand tiles: all reads are done)
    const bool gate = p.epi == EPI_GATE;
    auto put = [&](int row, int q, int s2, float f0, float f1, float g0, float g1) {
        // f = columns (16q + s2, +1), g = columns (16q + 8 + s2, +1) of this CTA's 64-column tile
        const int c = n0 + 16 * q + s2;
        float bf0 = 0.f, bf1 = 0.f, bg0 = 0.f, bg1 = 0.f;
        if (p.bias && c < p.N) bf0 = p.bias[c], bf1 = p.bias[c + 1], bg0 = p.bias[c + 8], bg1 = p.bias[c + 9];
        if (gate) {
            Cs[row * kCsLd + 8 * q + s2] = (f0 + bf0) * (g0 + bg0);
            Cs[row * kCsLd + 8 * q + s2 + 1] = (f1 + bf1) * (g1 + bg1);
        } else {
            Cs[row * kCsLd + 16 * q + s2] = f0 + bf0;
            Cs[row * kCsLd + 16 * q + s2 + 1] = f1 + bf1;
            Cs[row * kCsLd + 16 * q + 8 + s2] = g0 + bg0;
            Cs[row * kCsLd + 16 * q + 8 + s2 + 1] = g1 + bg1;
        }
    };
    if constexpr (kF32) {
        const int ty = tid >> 4, tx = tid & 15, q = tx >> 2, s = tx & 3;
#pragma unroll
        for (int i = 0; i < 8; ++i) put(ty * 8 + i, q, 2 * s, acc[i * 4], acc[i * 4 + 1], acc[i * 4 + 2], acc[i * 4 + 3]);
    } else {
        const int g = lane >> 2, s = lane & 3;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            const float* c0 = &acc[(np * 2 + 0) * 4];
            const float* c1 = &acc[(np * 2 + 1) * 4];
            put(warp * 16 + g, np, 2 * s, c0[0], c0[1], c1[0], c1[1]);
            put(warp * 16 + g + 8, np, 2 * s, c0[2], c0[3], c1[2], c1[3]);
        }
    }
    __syncthreads();

    // ---- stage 2: coalesced 8-element stores (+ residual / PixelShuffle scatter + skip)
    const int tno = gate ? kPwTN / 2 : kPwTN, vpr = tno / 8;
    const int nout = gate ? p.N / 2 : p.N, nbase = gate ? n0 / 2 : n0;
    T* out = reinterpret_cast<T*>(p.out);
    const T* res = reinterpret_cast<const T*>(p.res);
    for (int idx = tid; idx < kPwTM * vpr; idx += kPwThreads) {
        const int row = idx / vpr, cv = (idx - row * vpr) * 8;
        const long long m = m0 + row;
        const int ncol = nbase + cv;
        if (m >= p.M || ncol >= nout) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = Cs[row * kCsLd + cv + j];
        if (p.epi == EPI_SHUFFLE) {
            // PixelShuffle(2) (arch.py:228): GEMM column ph*Cq + c (rows permuted at pack time) -> pixel (2i+dy, 2j+dx), channel c
            const int cq = p.N / 4, ph = ncol / cq, c = ncol - ph * cq;
            const long long hw = (long long)p.H * p.W;
            const long long img = m / hw, rem = m - img * hw;
            const int i = (int)(rem / p.W), j2 = (int)(rem - (long long)i * p.W);
            const long long orow = (img * 2 * p.H + 2 * i + (ph >> 1)) * (2 * p.W) + 2 * j2 + (ph & 1);
            float r[8];
            load8(res + orow * p.ldr + c, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += r[j];
            store8(out + orow * p.ldo + c, v);
        } else {
            if (p.epi == EPI_RES) {
                float r[8];
                load8(res + m * p.ldr + ncol, r);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += r[j];
            }
            store8(out + m * p.ldo + ncol, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ bf16 fast path
// Same tile (128 pixels x 64 GEMM columns, K chunks of 32) and the same epilogues as hfrm_pw_kernel, re-organised for
// memory-level parallelism -- the generic kernel moves 24 KiB per CTA behind three dependent global round trips (statistics,
// operands, residual) and ran at 1 TB/s on the full-resolution tensors and at 60 TFLOP/s on the deep levels:
//   * persistent CTAs walk (pixel tile, column tile) work items; operand chunks stream through a 4-stage cp.async ring that
//     runs THREE chunks ahead across tile boundaries, so the loads of the next tiles are in flight during the MMAs, the
//     epilogue and the stores of the current one;
//   * LayerNorm2d is applied in the epilogue: with the affine folded into W' at pack time,
//       sum_k W'[n][k] (x_k - mu) rstd = rstd (acc[n] - mu wsum[n]),   wsum[n] = sum_k W'[n][k],
//     so the MMAs consume the raw tensor straight from the ring (no transform pass, no statistics pre-pass); mu and E[x^2]
//     are accumulated per row from the chunks as they land (each thread re-reads the two 16-byte vectors it copied);
//   * the channel-attention scale (conv3) is applied to the landed chunk in place by the thread that copied it.
constexpr int kRingStages = 4;
constexpr int kRingLd = 40;                                            // bf16 elements per smem row (80 bytes)
constexpr int kRingStageBytes = (kPwTM + kPwTN) * kRingLd * 2;         // 15 360
constexpr int kRingSmem = kRingStages * kRingStageBytes + kPwTM * kCsLd * 4 + 2 * kPwTM * 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kPwThreads, 2) hfrm_pw_ring_kernel(const PwParams p, int m_tiles, int n_tiles) {
    extern __shared__ __align__(128) unsigned char rsm[];
    float* Cs = reinterpret_cast<float*>(rsm + kRingStages * kRingStageBytes);
    float* mu_s = Cs + kPwTM * kCsLd;
    float* rstd_s = mu_s + kPwTM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bf16* A = reinterpret_cast<const bf16*>(p.A);
    const bf16* Wt = reinterpret_cast<const bf16*>(p.Wt);
    const int nk = p.K / kPwTK;
    const int total = m_tiles * n_tiles;
    const int my_tiles = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long my_chunks = (long long)my_tiles * nk;
    const long long hw = (long long)p.H * p.W;

    // staging coordinates of this thread: A vectors (row ar[i], 8 elements at akv[i]), W vector (row wr, wkv)
    int ar[2], akv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) ar[i] = (tid + i * kPwThreads) >> 2, akv[i] = ((tid + i * kPwThreads) & 3) * 8;
    const int wr = tid >> 2, wkv = (tid & 3) * 8;

    // issue cursor
    long long ic = 0;       // chunks issued
    int i_tile = 0, i_kc = 0;
    auto issue = [&]() {
        if (ic < my_chunks) {
            const int w = blockIdx.x + i_tile * gridDim.x;
            const int mt = w / n_tiles, nt = w - mt * n_tiles;
            const long long m0 = (long long)mt * kPwTM;
            const int n0 = nt * kPwTN, k0 = i_kc * kPwTK;
            bf16* As = reinterpret_cast<bf16*>(rsm + (ic % kRingStages) * kRingStageBytes);
            bf16* Ws = As + kPwTM * kRingLd;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const long long m = m0 + ar[i];
                const bool ok = m < p.M;
                long long row = ok ? m : 0;
                int k = k0 + akv[i];
                if (p.a_mode == AM_S2D) {
                    const long long img = row / hw, rem = row - img * hw;
                    const int oy = (int)(rem / p.W), ox = (int)(rem - (long long)oy * p.W);
                    const int tap = k0 / p.Cin;
                    row = (img * 2 * p.H + 2 * oy + (tap >> 1)) * (2 * p.W) + 2 * ox + (tap & 1);
                    k -= tap * p.Cin;
                }
                cp_async16(As + ar[i] * kRingLd + akv[i], A + row * p.lda + k, ok);
            }
            const bool wok = n0 + wr < p.N;
            cp_async16(Ws + wr * kRingLd + wkv, Wt + (long long)(wok ? n0 + wr : 0) * p.K + k0 + wkv, wok);
            if (++i_kc == nk) i_kc = 0, ++i_tile;
        }
        cp_async_commit();  // one group per call, empty past the end: the wait arithmetic stays uniform
        ++ic;
    };

    for (int i = 0; i < kRingStages - 1; ++i) issue();

    long long cc = 0;  // chunks consumed
    for (int t = 0; t < my_tiles; ++t) {
        const int w = blockIdx.x + t * gridDim.x;
        const int mt = w / n_tiles, nt = w - mt * n_tiles;
        const long long m0 = (long long)mt * kPwTM;
        const int n0 = nt * kPwTN;
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
        for (int kc = 0; kc < nk; ++kc, ++cc) {
            cp_async_wait<kRingStages - 2>();  // this thread's copies of chunk cc have landed
            bf16* As = reinterpret_cast<bf16*>(rsm + (cc % kRingStages) * kRingStageBytes);
            bf16* Ws = As + kPwTM * kRingLd;
            if (p.pro != PRO_NONE) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float v[8];
                    load8(As + ar[i] * kRingLd + akv[i], v);
                    if (p.pro == PRO_LN) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) s1[i] += v[j], s2[i] = fmaf(v[j], v[j], s2[i]);
                    } else {
                        const long long m = m0 + ar[i];
                        if (m < p.M) {
                            float sc[8];
                            load8(p.scale + (m / hw) * p.K + kc * kPwTK + akv[i], sc);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= sc[j];
                            store8(As + ar[i] * kRingLd + akv[i], v);
                        }
                    }
                }
            }
            __syncthreads();  // chunk cc visible to all; every thread is done with chunk cc - 1, whose stage is refilled now
            issue();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                ldmatrix_x4(a, As + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kRingLd + ks * 16 + ((lane >> 4) & 1) * 8);
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t b[4];
                    ldmatrix_x4(b, Ws + (np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * kRingLd + ks * 16 + ((lane >> 3) & 1) * 8);
                    float(&c0)[4] = *reinterpret_cast<float(*)[4]>(&acc[(np * 2 + 0) * 4]);
                    float(&c1)[4] = *reinterpret_cast<float(*)[4]>(&acc[(np * 2 + 1) * 4]);
                    mma_bf16(c0, a, b[0], b[1]);
                    mma_bf16(c1, a, b[2], b[3]);
                }
            }
        }
        // ---- epilogue
        if (p.pro == PRO_LN) {
            // a row's K elements were summed by the 4 threads that copied it (adjacent lanes)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 1), s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], 1);
                s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 2), s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], 2);
                if ((tid & 3) == 0) {
                    const float mean = s1[i] / (float)p.K;
                    const float var = fmaxf(s2[i] / (float)p.K - mean * mean, 0.f);
                    mu_s[ar[i]] = mean;
                    rstd_s[ar[i]] = 1.0f / sqrtf(var + p.eps);
                }
            }
            __syncthreads();
        }
        const bool gate = p.epi == EPI_GATE;
        {
            const int g = lane >> 2, s = lane & 3;
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int c = n0 + 16 * np + 2 * s;
                float bv[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
                if (c < p.N) {
                    if (p.bias) bv[0] = p.bias[c], bv[1] = p.bias[c + 1], bv[2] = p.bias[c + 8], bv[3] = p.bias[c + 9];
                    if (p.pro == PRO_LN) wv[0] = p.wsum[c], wv[1] = p.wsum[c + 1], wv[2] = p.wsum[c + 8], wv[3] = p.wsum[c + 9];
                }
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const int row = warp * 16 + g + 8 * hr;
                    float f0 = acc[(np * 2) * 4 + 2 * hr], f1 = acc[(np * 2) * 4 + 2 * hr + 1];
                    float g0 = acc[(np * 2 + 1) * 4 + 2 * hr], g1 = acc[(np * 2 + 1) * 4 + 2 * hr + 1];
                    if (p.pro == PRO_LN) {
                        const float mu = mu_s[row], rs = rstd_s[row];
                        f0 = rs * (f0 - mu * wv[0]), f1 = rs * (f1 - mu * wv[1]);
                        g0 = rs * (g0 - mu * wv[2]), g1 = rs * (g1 - mu * wv[3]);
                    }
                    f0 += bv[0], f1 += bv[1], g0 += bv[2], g1 += bv[3];
                    if (gate) {
                        Cs[row * kCsLd + 8 * np + 2 * s] = f0 * g0;
                        Cs[row * kCsLd + 8 * np + 2 * s + 1] = f1 * g1;
                    } else {
                        Cs[row * kCsLd + 16 * np + 2 * s] = f0;
                        Cs[row * kCsLd + 16 * np + 2 * s + 1] = f1;
                        Cs[row * kCsLd + 16 * np + 8 + 2 * s] = g0;
                        Cs[row * kCsLd + 16 * np + 8 + 2 * s + 1] = g1;
                    }
                }
            }
        }
        __syncthreads();
        // stage 2: coalesced 16-byte stores; the residual / skip loads of all of this thread's vectors are issued first
        const int tno = gate ? kPwTN / 2 : kPwTN, vpr = tno / 8;
        const int nout = gate ? p.N / 2 : p.N, nbase = gate ? n0 / 2 : n0;
        bf16* out = reinterpret_cast<bf16*>(p.out);
        const bf16* res = reinterpret_cast<const bf16*>(p.res);
        const int nit = (kPwTM * vpr) / kPwThreads;  // 4 (2 with the gate)
        long long orow[4];
        int ocol[4], crow[4], ccol[4];
        uint4 rv[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            orow[it] = -1;
            if (it >= nit) continue;
            const int idx = tid + it * kPwThreads;
            const int row = idx / vpr, cv = (idx - row * vpr) * 8;
            const long long m = m0 + row;
            const int ncol = nbase + cv;
            if (m >= p.M || ncol >= nout) continue;
            crow[it] = row, ccol[it] = cv;
            if (p.epi == EPI_SHUFFLE) {
                const int cq = p.N / 4, ph = ncol / cq;
                const long long img = m / hw, rem = m - img * hw;
                const int i = (int)(rem / p.W), j2 = (int)(rem - (long long)i * p.W);
                orow[it] = (img * 2 * p.H + 2 * i + (ph >> 1)) * (2 * p.W) + 2 * j2 + (ph & 1);
                ocol[it] = ncol - ph * cq;
            } else {
                orow[it] = m, ocol[it] = ncol;
            }
            if (p.epi == EPI_SHUFFLE || p.epi == EPI_RES) rv[it] = *reinterpret_cast<const uint4*>(res + orow[it] * p.ldr + ocol[it]);
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            if (orow[it] < 0) continue;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = Cs[crow[it] * kCsLd + ccol[it] + j];
            if (p.epi == EPI_SHUFFLE || p.epi == EPI_RES) {
                const uint32_t wv[4] = {rv[it].x, rv[it].y, rv[it].z, rv[it].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[2 * i] += __uint_as_float(wv[i] << 16);
                    v[2 * i + 1] += __uint_as_float(wv[i] & 0xffff0000u);
                }
            }
            store8(out + orow[it] * p.ldo + ocol[it], v);
        }
        // the next tile writes Cs / mu_s only after at least one __syncthreads of its K loop
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ bf16, C = 32 / 64
// The full- and half-resolution 1x1 convs (K = 32 / 64, N <= 128) are pure streaming: 64..256 bytes per pixel against
// 2..16 K MACs. Here the whole weight matrix stays in shared memory and every WARP runs its own pipeline over 16-pixel strips
// -- a private 4-stage cp.async ring for the activations, a private fp32 patch for the coalesced stores, __syncwarp only (no
// CTA-wide barrier after the weight load). LayerNorm in the epilogue as in hfrm_pw_ring_kernel; the row statistics never
// leave registers (the lanes that copied a row are the lanes whose accumulators need it, or one shuffle away).
constexpr int kStripPatchLd = 72;  // 72 % 32 == 8: the 64-bit fragment stores of a half-warp hit 32 distinct banks
template <int K>
struct StripCfg {
    static constexpr int kLd = K + 8;                                   // bf16 elements per smem row
    static constexpr int kStages = K <= 32 ? 4 : 3;                     // K = 64, N = 128: 107 KiB, two CTAs per SM
    static constexpr int kRingBytes = kStages * 16 * kLd * 2;           // per warp
    static constexpr int kPatchBytes = 16 * kStripPatchLd * 4;          // per warp
    static constexpr int kWarpBytes = kRingBytes + kPatchBytes;
    static int smem_bytes(int N) { return ((N * kLd * 2 + 2 * N * 4 + 127) / 128) * 128 + 8 * kWarpBytes; }
};

// PRO / EPI are template parameters: the four combinations a ResidualBlock uses compile to straight-line code (the runtime
// switches of the generic kernels cost this one a third of its issue slots)
template <int K, int PRO, int EPI>
__global__ void __launch_bounds__(256) hfrm_pw_strip_kernel(const PwParams p, long long n_strips) {
    using C = StripCfg<K>;
    constexpr int LD = C::kLd, VPR = K / 8, NV = 16 * VPR / 32, RPS = 32 / VPR;  // vectors / row, vectors / lane, rows / slot
    constexpr int kStripStages = C::kStages;
    extern __shared__ __align__(128) unsigned char ssm[];
    bf16* Wsm = reinterpret_cast<bf16*>(ssm);
    float* bias_s = reinterpret_cast<float*>(ssm + p.N * LD * 2);
    float* wsum_s = bias_s + p.N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* wbase = ssm + ((p.N * LD * 2 + 2 * p.N * 4 + 127) / 128) * 128 + warp * C::kWarpBytes;
    bf16* ring = reinterpret_cast<bf16*>(wbase);
    float* patch = reinterpret_cast<float*>(wbase + C::kRingBytes);
    const bf16* A = reinterpret_cast<const bf16*>(p.A);
    const long long hw = (long long)p.H * p.W;

    const long long first = (long long)blockIdx.x * 8 + warp, stride = (long long)gridDim.x * 8;
    const long long mine = first < n_strips ? (n_strips - first + stride - 1) / stride : 0;
    long long ic = 0;
    auto issue = [&]() {
        if (ic < mine) {
            const long long m0 = (first + ic * stride) * 16;
            bf16* st = ring + (ic % kStripStages) * 16 * LD;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int vec = lane + 32 * i, r = vec / VPR, kv = (vec - r * VPR) * 8;
                const bool ok = m0 + r < p.M;
                cp_async16(st + r * LD + kv, A + (ok ? m0 + r : 0) * p.lda + kv, ok);
            }
        }
        cp_async_commit();
        ++ic;
    };
    for (int i = 0; i < kStripStages - 1; ++i) issue();  // in flight while the weights are staged

    {
        const bf16* Wt = reinterpret_cast<const bf16*>(p.Wt);
        for (int v = tid; v < p.N * VPR; v += 256) {
            const int r = v / VPR, kv = (v - r * VPR) * 8;
            *reinterpret_cast<uint4*>(Wsm + r * LD + kv) = *reinterpret_cast<const uint4*>(Wt + (long long)r * K + kv);
        }
        for (int i = tid; i < p.N; i += 256) {
            bias_s[i] = p.bias ? p.bias[i] : 0.f;
            wsum_s[i] = (PRO == PRO_LN && p.wsum) ? p.wsum[i] : 0.f;
        }
    }
    __syncthreads();

    constexpr bool gate = EPI == EPI_GATE;
    const int g = lane >> 2, s = lane & 3;
    const int ngroups = (p.N + 63) / 64;
    bf16* out = reinterpret_cast<bf16*>(p.out);
    const bf16* res = reinterpret_cast<const bf16*>(p.res);
    for (long long j = 0; j < mine; ++j) {
        const long long m0 = (first + j * stride) * 16;
        bf16* st = ring + (j % kStripStages) * 16 * LD;
        cp_async_wait<kStripStages - 2>();
        // ---- row statistics / channel-attention scale on the vectors this lane copied
        float mean[NV], rstd[NV];
        if (PRO != PRO_NONE) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int vec = lane + 32 * i, r = vec / VPR, kv = (vec - r * VPR) * 8;
                float v[8];
                load8(st + r * LD + kv, v);
                if (PRO == PRO_LN) {
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int q = 0; q < 8; ++q) a += v[q], b = fmaf(v[q], v[q], b);
#pragma unroll
                    for (int o = 1; o < VPR; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o), b += __shfl_xor_sync(0xffffffffu, b, o);
                    mean[i] = a / (float)K;
                    rstd[i] = 1.0f / sqrtf(fmaxf(b / (float)K - mean[i] * mean[i], 0.f) + p.eps);
                } else if (m0 + r < p.M) {
                    float sc[8];
                    load8(p.scale + ((m0 + r) / hw) * K + kv, sc);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] *= sc[q];
                    store8(st + r * LD + kv, v);
                }
            }
        }
        __syncwarp();
        // statistics of the rows whose accumulators this lane holds: g and g + 8. Row r sits in slot r / RPS of the lanes
        // (r % RPS) * VPR ..; with VPR == 4 that is this lane's own group
        float mu0 = 0.f, rs0 = 1.f, mu1 = 0.f, rs1 = 1.f;
        if (PRO == PRO_LN) {
            if (VPR == 4) {
                mu0 = mean[0], rs0 = rstd[0], mu1 = mean[NV - 1], rs1 = rstd[NV - 1];
            } else {
                const int src = (g % RPS) * VPR;
                float t0[NV], t1[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) t0[i] = __shfl_sync(0xffffffffu, mean[i], src), t1[i] = __shfl_sync(0xffffffffu, rstd[i], src);
                const int sl = g / RPS;  // 0 or 1; row g + 8 is slot sl + 8 / RPS
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (i == sl) mu0 = t0[i], rs0 = t1[i];
                    if (i == sl + 8 / RPS) mu1 = t0[i], rs1 = t1[i];
                }
            }
        }
        uint32_t a[K / 16][4];
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks)
            ldmatrix_x4(a[ks], st + ((lane & 7) + ((lane >> 3) & 1) * 8) * LD + ks * 16 + ((lane >> 4) & 1) * 8);
        __syncwarp();
        issue();  // refills the stage strip j - 1 used
        for (int ng = 0; ng < ngroups; ++ng) {
            const int n0 = ng * 64;
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                if (n0 + 16 * np >= p.N) continue;
#pragma unroll
                for (int ks = 0; ks < K / 16; ++ks) {
                    uint32_t b[4];
                    ldmatrix_x4(b, Wsm + (n0 + np * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
                    float(&c0)[4] = *reinterpret_cast<float(*)[4]>(&acc[(np * 2 + 0) * 4]);
                    float(&c1)[4] = *reinterpret_cast<float(*)[4]>(&acc[(np * 2 + 1) * 4]);
                    mma_bf16(c0, a[ks], b[0], b[1]);
                    mma_bf16(c1, a[ks], b[2], b[3]);
                }
            }
            // ---- fragments -> this warp's fp32 patch (+ LayerNorm algebra, bias, SimpleGate), 64-bit shared accesses
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int c = n0 + 16 * np + 2 * s;
                if (c >= p.N) continue;
                const float2 bf = *reinterpret_cast<const float2*>(bias_s + c), bg = *reinterpret_cast<const float2*>(bias_s + c + 8);
                const float2 wf = *reinterpret_cast<const float2*>(wsum_s + c), wg = *reinterpret_cast<const float2*>(wsum_s + c + 8);
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const int row = g + 8 * hr;
                    float f0 = acc[(np * 2) * 4 + 2 * hr], f1 = acc[(np * 2) * 4 + 2 * hr + 1];
                    float g0 = acc[(np * 2 + 1) * 4 + 2 * hr], g1 = acc[(np * 2 + 1) * 4 + 2 * hr + 1];
                    if (PRO == PRO_LN) {
                        const float mu = hr ? mu1 : mu0, rs = hr ? rs1 : rs0;
                        f0 = rs * (f0 - mu * wf.x), f1 = rs * (f1 - mu * wf.y);
                        g0 = rs * (g0 - mu * wg.x), g1 = rs * (g1 - mu * wg.y);
                    }
                    f0 += bf.x, f1 += bf.y, g0 += bg.x, g1 += bg.y;
                    if (gate) {
                        *reinterpret_cast<float2*>(patch + row * kStripPatchLd + 8 * np + 2 * s) = make_float2(f0 * g0, f1 * g1);
                    } else {
                        *reinterpret_cast<float2*>(patch + row * kStripPatchLd + 16 * np + 2 * s) = make_float2(f0, f1);
                        *reinterpret_cast<float2*>(patch + row * kStripPatchLd + 16 * np + 8 + 2 * s) = make_float2(g0, g1);
                    }
                }
            }
            __syncwarp();
            // ---- coalesced 16-byte stores (+ residual, loaded first)
            constexpr int tno = gate ? 32 : 64, vpr = tno / 8;
            const int nout = gate ? p.N / 2 : p.N, nbase = gate ? n0 / 2 : n0;
            constexpr int nit = 16 * vpr / 32;  // 4 (2 with the gate)
            long long om[4];
            int oc[4], pr[4], pc[4];
            uint4 rv[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                om[it] = -1;
                if (it >= nit) continue;
                const int idx = lane + 32 * it, row = idx / vpr, cv = (idx - row * vpr) * 8;
                if (m0 + row >= p.M || nbase + cv >= nout) continue;
                om[it] = m0 + row, oc[it] = nbase + cv, pr[it] = row, pc[it] = cv;
                if (EPI == EPI_RES) rv[it] = *reinterpret_cast<const uint4*>(res + om[it] * p.ldr + oc[it]);
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                if (om[it] < 0) continue;
                float v[8];
                load8(patch + pr[it] * kStripPatchLd + pc[it], v);
                if (EPI == EPI_RES) {
                    const uint32_t w4[4] = {rv[it].x, rv[it].y, rv[it].z, rv[it].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[2 * i] += __uint_as_float(w4[i] << 16);
                        v[2 * i + 1] += __uint_as_float(w4[i] & 0xffff0000u);
                    }
                }
                store8(out + om[it] * p.ldo + oc[it], v);
            }
            __syncwarp();  // the patch is rewritten by the next column group / strip
        }
    }
    cp_async_wait<0>();
}

template <int K, int PRO, int EPI>
int launch_pw_strip_t(const PwParams& p, cudaStream_t s, int sms) {
    using C = StripCfg<K>;
    const int smem = C::smem_bytes(p.N);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(hfrm_pw_strip_kernel<K, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::smem_bytes(128));
        if (e != cudaSuccess) return wdm_cuda_error((int)e);
        attr = true;
    }
    const long long n_strips = ((long long)p.M + 15) / 16;
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const long long want = (n_strips + 7) / 8;
    const int grid = (int)(want < (long long)per_sm * sms ? want : (long long)per_sm * sms);
    hfrm_pw_strip_kernel<K, PRO, EPI><<<grid, 256, smem, s>>>(p, n_strips);
    return wdm_launch_status();
}

// the (prologue, epilogue) pairs of a ResidualBlock; anything else goes to the ring kernel
template <int K>
int launch_pw_strip(const PwParams& p, cudaStream_t s, int sms, bool* handled) {
    *handled = true;
    if (p.pro == PRO_LN && p.epi == EPI_BIAS) return launch_pw_strip_t<K, PRO_LN, EPI_BIAS>(p, s, sms);
    if (p.pro == PRO_LN && p.epi == EPI_GATE) return launch_pw_strip_t<K, PRO_LN, EPI_GATE>(p, s, sms);
    if (p.pro == PRO_SCALE && p.epi == EPI_RES) return launch_pw_strip_t<K, PRO_SCALE, EPI_RES>(p, s, sms);
    if (p.pro == PRO_NONE && p.epi == EPI_RES) return launch_pw_strip_t<K, PRO_NONE, EPI_RES>(p, s, sms);
    *handled = false;
    return WDM_OK;
}

int g_pw_ring = -1;  // WDM_HFRM_RING=0 keeps the generic kernel in bf16 mode (A/B testing)

template <typename T>
int launch_pw(const PwParams& p, cudaStream_t s) {
    if (p.M <= 0) return WDM_OK;
    if ((p.K % kPwTK) || (p.N % 16) || (p.lda % 8) || (p.ldo % 8) || !p.A || !p.Wt || !p.out) return WDM_ERR_BAD_SHAPE;
    if (p.a_mode == AM_S2D && ((p.Cin % kPwTK) || p.K != 4 * p.Cin || p.pro != PRO_NONE)) return WDM_ERR_BAD_SHAPE;
    if (p.epi == EPI_SHUFFLE && ((p.N / 4) % 8 || !p.res)) return WDM_ERR_BAD_SHAPE;
    if (p.epi == EPI_RES && (!p.res || (p.ldr % 8))) return WDM_ERR_BAD_SHAPE;
    const int m_tiles = (p.M + kPwTM - 1) / kPwTM, n_tiles = (p.N + kPwTN - 1) / kPwTN;
    if (std::is_same<T, bf16>::value) {
        if (g_pw_ring < 0) {
            const char* e = getenv("WDM_HFRM_RING");
            g_pw_ring = e ? atoi(e) : 1;
        }
        if (g_pw_ring && (p.pro != PRO_LN || p.wsum)) {
            static int sms = 0;
            if (!sms) {
                int dev = 0;
                cudaGetDevice(&dev);
                if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
                cudaError_t e = cudaFuncSetAttribute(hfrm_pw_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem);
                if (e != cudaSuccess) return wdm_cuda_error((int)e);
            }
            // C = 32 / 64 (full / half resolution, HBM bound): resident weights, one pipeline per warp
            if (g_pw_ring != 2 && (p.K == 32 || p.K == 64) && p.N <= 128 && p.a_mode == AM_PLAIN && p.epi != EPI_SHUFFLE) {
                bool handled = false;
                const int r = p.K == 32 ? launch_pw_strip<32>(p, s, sms, &handled) : launch_pw_strip<64>(p, s, sms, &handled);
                if (handled) return r;
            }
            const long long total = (long long)m_tiles * n_tiles;
            const int grid = (int)(total < 2LL * sms ? total : 2LL * sms);
            hfrm_pw_ring_kernel<<<grid, kPwThreads, kRingSmem, s>>>(p, m_tiles, n_tiles);
            return wdm_launch_status();
        }
    }
    dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
    hfrm_pw_kernel<T><<<grid, kPwThreads, 0, s>>>(p);
    return wdm_launch_status();
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3 + SimpleGate
// conv2 (arch.py:163-165, groups = 2C, padding 1) on t1 [B, H, W, 2C], SimpleGate x1 * x2 (arch.py:132-141) -> out
// [B, H, W, C], and the per-CTA channel sums of the gated tensor (the global average pool of ChannelAttn, arch.py:143-155, is
// finished by hfrm_chan_kernel in a fixed order: deterministic).
// Thread = one pixel COLUMN x one channel pair (c, c+1) and its SimpleGate partners (C+c, C+c+1): the 2 x 2 x 9 weights live
// in registers and the thread slides down kDwTH rows keeping three rows of accumulators, so every input row is loaded once
// per column (3 x 2 small loads per output pixel instead of 9 x 2; a warp's loads are 128 contiguous bytes).
// CTA = 256 threads = (256 / PP) columns x PP channel pairs, PP = min(C, 64) / 2; grid (tiles, B, C / 64).
constexpr int kDwTH = 32, kDwCS = 64;
__host__ __device__ constexpr int dw_tile_w(int C) { return 256 / ((C < kDwCS ? C : kDwCS) / 2); }

__device__ __forceinline__ void load2(const bf16* p, float& a, float& b) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ void load2(const float* p, float& a, float& b) {
    const float2 u = *reinterpret_cast<const float2*>(p);
    a = u.x, b = u.y;
}
__device__ __forceinline__ void store2(bf16* p, float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<__nv_bfloat162*>(p) = t;
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// one input row (3 columns x {first pair, partner pair}) into the three live output rows; PH = input row mod 3
template <int PH>
__device__ __forceinline__ void dw_accumulate(float (&acc)[3][4], const float (&v)[3][4], const float (&w)[9][4]) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[(PH + 1) % 3][e] = fmaf(v[dx][e], w[0 * 3 + dx][e], acc[(PH + 1) % 3][e]);  // output row r + 1: tap dy = 0
            acc[PH][e] = fmaf(v[dx][e], w[1 * 3 + dx][e], acc[PH][e]);                      // output row r:     dy = 1
            acc[(PH + 2) % 3][e] = fmaf(v[dx][e], w[2 * 3 + dx][e], acc[(PH + 2) % 3][e]);  // output row r - 1: dy = 2
        }
}

// CT > 0: the channel count as a compile-time constant (C = 32 / 64, the full / half resolution levels that carry 75 % of
// this kernel's time): neighbour-pixel and partner-half offsets become immediates of the load instructions -- the generic
// form spent two thirds of its issue slots on address arithmetic.
template <typename T, int CT>
__global__ void __launch_bounds__(256, 2) hfrm_dw_gate_kernel(const T* __restrict__ in, const float* __restrict__ w9 /*[9][2C]*/,
                                                          const float* __restrict__ bias /*[2C]*/, T* __restrict__ out,
                                                          float* __restrict__ partial /*[B][tiles][C]*/, int C_rt, int H, int W,
                                                          int tiles_x) {
    const int C = CT ? CT : C_rt;
    __shared__ float red[256 * 2];
    const int CS = C < kDwCS ? C : kDwCS, c0 = blockIdx.z * CS, PP = CS >> 1, TW = 256 / PP;
    const int tid = threadIdx.x, cp = tid % PP, xl = tid / PP;
    const int c = c0 + 2 * cp;
    const long long img = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * kDwTH, x = (blockIdx.x % tiles_x) * TW + xl;
    const bool active = x < W;
    float w[9][4], b4[4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float2 a = *reinterpret_cast<const float2*>(w9 + t * 2 * C + c), d = *reinterpret_cast<const float2*>(w9 + t * 2 * C + C + c);
        w[t][0] = a.x, w[t][1] = a.y, w[t][2] = d.x, w[t][3] = d.y;
    }
    b4[0] = bias[c], b4[1] = bias[c + 1], b4[2] = bias[C + c], b4[3] = bias[C + c + 1];
    {
        // fold the column validity of x - 1, x, x + 1 (and of the whole thread, x >= W) into the weights
        const int xg = (blockIdx.x % tiles_x) * TW + xl;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int xx = xg + dx - 1;
            const float mk = (xx >= 0 && xx < W && xg < W) ? 1.f : 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int e = 0; e < 4; ++e) w[dy * 3 + dx][e] *= mk;
        }
    }
    float acc[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = b4[e];
    float sum0 = 0.f, sum1 = 0.f;
    const T* base = in + img * H * W * 2 * C;
    T* obase = out + img * H * W * C;
    const int r_last = ty0 + kDwTH < H ? ty0 + kDwTH : H;  // last input row needed (row H is the zero padding row)
    // one running pointer to (row, x) of this thread's channel pair; the neighbours x - 1 / x + 1 and the partner half sit at
    // fixed element offsets from it (-2C, +2C, +C). Out-of-image columns are predicated off (their weights are zero and the
    // registers keep a finite earlier value); rows outside the image are not loaded at all (r is uniform over the CTA)
    const bool vx0 = active && x >= 1, vx1 = active, vx2 = active && x + 1 < W;
    const int rowpitch = W * 2 * C, opitch = W * C;
    const T* rp = base + ((long long)(ty0 - 1) * rowpitch + (long long)(active ? x : 0) * 2 * C + c);  // may point before row 0: not
    int rf = ty0 - 1;                                                                                // dereferenced then
    int ooff = (ty0 - 1) * opitch + x * C + c;  // offset of output row (r - 1) for the step of input row r = ty0
    auto fetch = [&](float (&v)[3][4]) {
        if (rf >= 0 && rf < H && rf <= r_last) {
            if (vx0) load2(rp - 2 * C, v[0][0], v[0][1]), load2(rp - C, v[0][2], v[0][3]);
            if (vx1) load2(rp, v[1][0], v[1][1]), load2(rp + C, v[1][2], v[1][3]);
            if (vx2) load2(rp + 2 * C, v[2][0], v[2][1]), load2(rp + 3 * C, v[2][2], v[2][3]);
        }
        ++rf, rp += rowpitch;
    };
    auto step = [&](int r, float (&v)[3][4], auto ph) {
        constexpr int PH = decltype(ph)::value;
        if (r > r_last) return;
        if (r >= 0 && r < H) dw_accumulate<PH>(acc, v, w);
        // output row r - 1 has now seen its three input rows
        constexpr int DONE = (PH + 2) % 3;
        if (r - 1 >= ty0 && active) {
            const float g0 = acc[DONE][0] * acc[DONE][2], g1 = acc[DONE][1] * acc[DONE][3];
            store2(obase + ooff, g0, g1);
            sum0 += g0, sum1 += g1;
        }
        if (r >= ty0) ooff += opitch;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[DONE][e] = b4[e];
    };
    // input rows ty0 - 1 .. r_last; the phase of a row is (row - (ty0 - 1)) mod 3 = the register buffer it was fetched into
    float va[3][4] = {}, vb[3][4] = {}, vc[3][4] = {};
    fetch(va);
    fetch(vb);
    for (int i = 0; i < kDwTH + 2; i += 3) {
        const int r = ty0 - 1 + i;
        fetch(vc);
        step(r, va, std::integral_constant<int, 0>());
        fetch(va);
        step(r + 1, vb, std::integral_constant<int, 1>());
        fetch(vb);
        step(r + 2, vc, std::integral_constant<int, 2>());
    }
    red[tid * 2] = sum0, red[tid * 2 + 1] = sum1;
    __syncthreads();
    for (int ch = tid; ch < CS; ch += 256) {
        float sacc = 0.f;
        for (int q = 0; q < TW; ++q) sacc += red[(q * PP + (ch >> 1)) * 2 + (ch & 1)];  // fixed order: deterministic
        partial[(img * gridDim.x + blockIdx.x) * C + c0 + ch] = sacc;
    }
}

// The same computation with the input staged through shared memory: the register-streaming kernel above keeps only
// ~12 KiB of unique DRAM bytes in flight per SM (16 warps x 3 rows x 2 pixels) and runs at 1.3 TB/s whatever its instruction
// count. Here all 256 threads cp.async whole bands of kDwRB rows x (TW + 2) pixels x 2 x CS channels into a 4-deep ring
// (three bands = ~40 KiB per CTA in flight), zero-filling everything outside the image (so the compute needs no masks), and
// the column threads read their 3 x 2 values per row from the ring.
constexpr int kDwRB = 4, kDwNB = 4;
template <typename T>
__host__ __device__ constexpr int dw_px_stride(int CS) {  // bytes per staged pixel; C = 32: +64 so that the two pixels of a warp
    return 2 * CS * (int)sizeof(T) + ((2 * CS * (int)sizeof(T)) % 256 == 0 ? 0 : 64);  // hit different banks
}
template <typename T>
__host__ __device__ constexpr int dw_band_bytes(int CS) { return kDwRB * (256 / (CS / 2) + 2) * dw_px_stride<T>(CS); }

template <typename T>
__global__ void __launch_bounds__(256, 2) hfrm_dw_gate_smem_kernel(const T* __restrict__ in, const float* __restrict__ w9,
                                                               const float* __restrict__ bias, T* __restrict__ out,
                                                               float* __restrict__ partial, int C, int H, int W, int tiles_x) {
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    __shared__ float red[256 * 2];
    const int CS = C < kDwCS ? C : kDwCS, c0 = blockIdx.z * CS, PP = CS >> 1, TW = 256 / PP;
    const int PS = dw_px_stride<T>(CS), band_bytes = kDwRB * (TW + 2) * PS;
    const int tid = threadIdx.x, cp = tid % PP, xl = tid / PP;
    const int c = c0 + 2 * cp;
    const long long img = blockIdx.y;
    const int ty0 = (blockIdx.x / tiles_x) * kDwTH, tx0 = (blockIdx.x % tiles_x) * TW, x = tx0 + xl;
    const bool active = x < W;
    float w[9][4], b4[4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float2 a = *reinterpret_cast<const float2*>(w9 + t * 2 * C + c), d = *reinterpret_cast<const float2*>(w9 + t * 2 * C + C + c);
        w[t][0] = a.x, w[t][1] = a.y, w[t][2] = d.x, w[t][3] = d.y;
    }
    b4[0] = bias[c], b4[1] = bias[c + 1], b4[2] = bias[C + c], b4[3] = bias[C + c + 1];
    float acc[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = b4[e];
    float sum0 = 0.f, sum1 = 0.f;
    const T* base = in + img * H * W * 2 * C;
    T* obase = out + img * H * W * C;
    const int r_last = ty0 + kDwTH < H ? ty0 + kDwTH : H;
    constexpr int EPV = 16 / (int)sizeof(T);          // elements per 16-byte vector
    const int vph = CS / EPV;                         // vectors per half of a pixel
    const int vpb = kDwRB * (TW + 2) * 2 * vph;       // vectors per band
    const int nbands = (kDwTH + 2 + kDwRB - 1) / kDwRB;
    auto issue = [&](int b) {
        if (b < nbands) {
            unsigned char* dst = dsm_raw + (b % kDwNB) * band_bytes;
            for (int v = tid; v < vpb; v += 256) {
                const int q = v % vph, hh = (v / vph) & 1, pp = (v / (2 * vph)) % (TW + 2), rr = v / (2 * vph * (TW + 2));
                const int r = ty0 - 1 + b * kDwRB + rr, xx = tx0 - 1 + pp;
                const bool ok = r >= 0 && r < H && xx >= 0 && xx < W;
                const T* src = base + ((long long)(ok ? r : 0) * W + (ok ? xx : 0)) * 2 * C + hh * C + c0 + q * EPV;
                cp_async16(dst + (rr * (TW + 2) + pp) * PS + (hh * CS + q * EPV) * (int)sizeof(T), src, ok);
            }
        }
        cp_async_commit();
    };
    int ooff = (ty0 - 1) * (W * C) + x * C + c;
    auto row = [&](const unsigned char* band, int rr, int r, auto ph) {
        constexpr int PH = decltype(ph)::value;
        if (r > r_last) return;
        float v[3][4];
        const unsigned char* p0 = band + (rr * (TW + 2) + xl) * PS + 2 * cp * (int)sizeof(T);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            load2(reinterpret_cast<const T*>(p0 + dx * PS), v[dx][0], v[dx][1]);
            load2(reinterpret_cast<const T*>(p0 + dx * PS) + CS, v[dx][2], v[dx][3]);
        }
        dw_accumulate<PH>(acc, v, w);  // rows / columns outside the image were zero-filled by the loader
        constexpr int DONE = (PH + 2) % 3;
        if (r - 1 >= ty0 && active) {
            const float g0 = acc[DONE][0] * acc[DONE][2], g1 = acc[DONE][1] * acc[DONE][3];
            store2(obase + ooff, g0, g1);
            sum0 += g0, sum1 += g1;
        }
        if (r >= ty0) ooff += W * C;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[DONE][e] = b4[e];
    };
    for (int b = 0; b < kDwNB - 1; ++b) issue(b);
    // three bands (12 rows) per pass: the accumulator phase of a row, (row index) mod 3, is then a compile-time constant
    for (int b0 = 0; b0 < nbands; b0 += 3) {
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const int b = b0 + bb;
            cp_async_wait<kDwNB - 2>();
            __syncthreads();          // band b has landed for every thread; band b - 1 is free again
            issue(b + kDwNB - 1);
            if (b < nbands) {
                const unsigned char* band = dsm_raw + (b % kDwNB) * band_bytes;
                const int r = ty0 - 1 + b * kDwRB;
                if (bb == 0) {
                    row(band, 0, r, std::integral_constant<int, 0>()), row(band, 1, r + 1, std::integral_constant<int, 1>());
                    row(band, 2, r + 2, std::integral_constant<int, 2>()), row(band, 3, r + 3, std::integral_constant<int, 0>());
                } else if (bb == 1) {
                    row(band, 0, r, std::integral_constant<int, 1>()), row(band, 1, r + 1, std::integral_constant<int, 2>());
                    row(band, 2, r + 2, std::integral_constant<int, 0>()), row(band, 3, r + 3, std::integral_constant<int, 1>());
                } else {
                    row(band, 0, r, std::integral_constant<int, 2>()), row(band, 1, r + 1, std::integral_constant<int, 0>());
                    row(band, 2, r + 2, std::integral_constant<int, 1>()), row(band, 3, r + 3, std::integral_constant<int, 2>());
                }
            }
        }
    }
    cp_async_wait<0>();
    red[tid * 2] = sum0, red[tid * 2 + 1] = sum1;
    __syncthreads();
    for (int ch = tid; ch < CS; ch += 256) {
        float sacc = 0.f;
        for (int q = 0; q < TW; ++q) sacc += red[(q * PP + (ch >> 1)) * 2 + (ch & 1)];  // fixed order: deterministic
        partial[(img * gridDim.x + blockIdx.x) * C + c0 + ch] = sacc;
    }
}

// ChannelAttn (arch.py:143-155): mean over the image of the gated tensor, then chan_conv 1x1: s[b][c] = bc[c] + Wc[c][:] . mean.
// grid (B, C / 64): every CTA reduces the tile sums of all C channels (fixed order: deterministic) and produces 64 outputs.
__global__ void __launch_bounds__(256) hfrm_chan_kernel(const float* __restrict__ partial, int tiles, int C, float inv_hw,
                                                       const float* __restrict__ Wc, const float* __restrict__ bc,
                                                       float* __restrict__ s_out) {
    extern __shared__ float csm[];  // mean[C] + red[256]
    float* mean = csm;
    float* red = csm + C;
    const long long img = blockIdx.x;
    const int t = threadIdx.x;
    if (C >= 256) {
        for (int c = t; c < C; c += 256) {
            float s = 0.f;
            for (int tl = 0; tl < tiles; ++tl) s += partial[(img * tiles + tl) * C + c];
            mean[c] = s * inv_hw;
        }
    } else {
        const int G = 256 / C, tg = t / C, c = t - tg * C;  // G groups of tiles per channel
        float s = 0.f;
        for (int tl = tg; tl < tiles; tl += G) s += partial[(img * tiles + tl) * C + c];
        red[t] = s;
        __syncthreads();
        if (t < C) {
            float a = 0.f;
            for (int g = 0; g < G; ++g) a += red[g * C + t];
            mean[t] = a * inv_hw;
        }
    }
    __syncthreads();
    const int warp = t >> 5, lane = t & 31;
    const int cbeg = blockIdx.y * 64, cend = cbeg + 64 < C ? cbeg + 64 : C;
    for (int c = cbeg + warp; c < cend; c += 8) {
        float s = 0.f;
        for (int k = lane; k < C; k += 32) s = fmaf(Wc[(long long)c * C + k], mean[k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_out[img * C + c] = s + bc[c];
    }
}

// ------------------------------------------------------------------------------------------------ conv_in / conv_out
// conv_in (arch.py:210): 3x3, 3 -> Cd, NCHW fp32 image -> NHWC. Thread = one pixel, 32 output channels at a time.
template <typename T>
__global__ void __launch_bounds__(256) hfrm_conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w /*[27][Cd]*/,
                                                          const float* __restrict__ bias, T* __restrict__ out, long long npix,
                                                          int H, int W, int Cd) {
    extern __shared__ float ws[];  // [27][Cd] + [Cd]
    for (int i = threadIdx.x; i < 27 * Cd; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < Cd; i += blockDim.x) ws[27 * Cd + i] = bias[i];
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long hw = (long long)H * W;
    const long long img = pix / hw, rem = pix - img * hw;
    const int y = (int)(rem / W), xq = (int)(rem - (long long)y * W);
    float in[27];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int yy = y + dy - 1, xx = xq + dx - 1;
                in[ci * 9 + dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[(img * 3 + ci) * hw + (long long)yy * W + xx] : 0.f;
            }
    for (int c0 = 0; c0 < Cd; c0 += 8) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = ws[27 * Cd + c0 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(in[k], ws[k * Cd + c0 + j], a[j]);
        store8(out + pix * Cd + c0, a);
    }
}

// conv_out (arch.py:232) + the input residual (arch.py:250): NHWC -> NCHW fp32, y = conv3x3(h) + b + x
template <typename T>
__global__ void __launch_bounds__(256) hfrm_conv_out_kernel(const T* __restrict__ h, const float* __restrict__ w /*[9][Cd][3]*/,
                                                           const float* __restrict__ bias, const float* __restrict__ x,
                                                           float* __restrict__ y, long long npix, int H, int W, int Cd) {
    extern __shared__ float ws[];  // [9][Cd][3]
    for (int i = threadIdx.x; i < 27 * Cd; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const long long hw = (long long)H * W;
    const long long img = pix / hw, rem = pix - img * hw;
    const int yq = (int)(rem / W), xq = (int)(rem - (long long)yq * W);
    float a0 = bias[0], a1 = bias[1], a2 = bias[2];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = yq + dy - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int xx = xq + dx - 1;
            if (xx < 0 || xx >= W) continue;
            const T* px = h + (img * hw + (long long)yy * W + xx) * Cd;
            const float* wt = ws + (dy * 3 + dx) * Cd * 3;
            for (int c0 = 0; c0 < Cd; c0 += 8) {
                float v[8];
                load8(px + c0, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a0 = fmaf(v[j], wt[(c0 + j) * 3], a0);
                    a1 = fmaf(v[j], wt[(c0 + j) * 3 + 1], a1);
                    a2 = fmaf(v[j], wt[(c0 + j) * 3 + 2], a2);
                }
            }
        }
    }
    const long long o = img * 3 * hw + rem;
    y[o] = a0 + x[o];
    y[o + hw] = a1 + x[o + hw];
    y[o + 2 * hw] = a2 + x[o + 2 * hw];
}

// ------------------------------------------------------------------------------------------------ deep levels (C >= 128)
// From C = 128 on the 1x1 convs are real GEMMs (M = B*h*w rows, K = C, N = C or 2C): they run on the tcgen05 kernel
// (launch_gemm_tc, wdm_gemm_tc.cu: bias and residual in its epilogue) and only the pieces that kernel has no epilogue for
// stay separate, as one-pass bf16 row kernels: the LayerNorm normalisation (affine folded into the conv), the
// channel-attention scale (in place) and the SimpleGate.
// mode 0: y = (x - mean) * rstd per row;  mode 1: x *= scale[img][c] in place;  mode 2: y[c] = x[c] * x[C + c]
__global__ void __launch_bounds__(256) hfrm_row_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long M, int C, int mode,
                                                      float eps, const float* __restrict__ scale, long long hw) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + warp;
    if (m >= M) return;
    if (mode == 0) {
        const bf16* row = x + m * C;
        float v[2][8];
        float s = 0.f;
        const int nv = C >> 3;  // <= 64 vectors: at most two per lane
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
            if (lane + 32 * i < nv) load8(row + (lane + 32 * i) * 8, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (lane + 32 * i < nv) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[i][j] - mean;
                    q = fmaf(d, d, q);
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = 1.0f / sqrtf(q / (float)C + eps);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (lane + 32 * i < nv) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i][j] = (v[i][j] - mean) * rstd;
                store8(y + m * C + (lane + 32 * i) * 8, v[i]);
            }
    } else if (mode == 1) {
        const float* sc = scale + (m / hw) * C;
        for (int k = lane * 8; k < C; k += 256) {
            float v[8], f[8];
            load8(x + m * C + k, v);
            load8(sc + k, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= f[j];
            store8(y + m * C + k, v);
        }
    } else {
        for (int k = lane * 8; k < C; k += 256) {
            float a[8], b[8];
            load8(x + m * 2 * C + k, a);
            load8(x + m * 2 * C + C + k, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] *= b[j];
            store8(y + m * C + k, a);
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
enum { PERM_NONE = 0, PERM_GATE = 1, PERM_SHUFFLE = 2 };
// one CTA per packed row r: Wout[r][k] = rowscale[n] * W[n][k] * colscale[k], bout[r] = rowscale[n] * (b[n] + sum_k W[n][k] lnb[k])
// with n = perm(r). src_down: W is [N][Cin][2][2] and k = (dy*2+dx)*Cin + c.
template <typename T>
__global__ void __launch_bounds__(128) hfrm_pack_pw_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                          const float* __restrict__ colscale, const float* __restrict__ lnb,
                                                          const float* __restrict__ rowscale, int N, int K, int perm, int src_down,
                                                          int Cin, T* __restrict__ Wout, float* __restrict__ bout,
                                                          float* __restrict__ wsum_out) {
    __shared__ float red[128], red2[128];
    const int r = blockIdx.x;
    int n = r;
    if (perm == PERM_GATE) {
        const int q = r >> 4, j = r & 15, C = N >> 1;
        n = j < 8 ? 8 * q + j : C + 8 * q + (j - 8);
    } else if (perm == PERM_SHUFFLE) {
        const int cq = N >> 2, ph = r / cq, c = r - ph * cq;
        n = c * 4 + ph;
    }
    const float rs = rowscale ? rowscale[n] : 1.f;
    float fold = 0.f, wsum = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        long long si = (long long)n * K + k;
        if (src_down) {
            const int ph = k / Cin, c = k - ph * Cin;
            si = ((long long)n * Cin + c) * 4 + ph;
        }
        const float w = W[si];
        if (lnb) fold = fmaf(w, lnb[k], fold);
        const T wq = from_f<T>(w * (colscale ? colscale[k] : 1.f) * rs);
        Wout[(long long)r * K + k] = wq;
        wsum += to_f(wq);  // of the STORED values: what the MMAs multiply the raw tensor with
    }
    red[threadIdx.x] = fold, red2[threadIdx.x] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, s2 = 0.f;
        for (int i = 0; i < 128; ++i) s += red[i], s2 += red2[i];
        if (bout) bout[r] = ((b ? b[n] : 0.f) + s) * rs;
        if (wsum_out) wsum_out[r] = s2;
    }
}
// small fp32 re-layouts: mode 0 = dw [2C][9] -> [9][2C]; 1 = conv_in [Cd][27] -> [27][Cd]; 2 = conv_out [3][Cd][9] -> [9][Cd][3]
__global__ void hfrm_pack_misc_kernel(const float* __restrict__ src, float* __restrict__ dst, int mode, int C, int total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    if (mode == 0) {
        const int tap = i / C, c = i - tap * C;  // C = 2C channels
        dst[i] = src[c * 9 + tap];
    } else if (mode == 1) {
        const int k = i / C, c = i - k * C;
        dst[i] = src[c * 27 + k];
    } else {
        const int o = i % 3, c = (i / 3) % C, tap = i / (3 * C);
        dst[i] = src[(o * C + c) * 9 + tap];
    }
}

// ------------------------------------------------------------------------------------------------ model
struct PRef {
    std::string name;
    long long numel = 0, off = 0;
};
struct PwW {
    void* w = nullptr;      // packed [N][K]
    float* b = nullptr;     // packed bias [N] (null: no bias)
    float* wsum = nullptr;  // row sums of the packed matrix (convs behind a LayerNorm: the ring kernel's epilogue form)
    int N = 0, K = 0;
};
struct BlockSpec {
    int C = 0;
    int beta, gamma, c1[2], c2[2], c3[2], cc[2], c4[2], c5[2], n1[2], n2[2];  // param indices
    PwW conv1, conv3, conv4, conv5;
    PwW conv4n;  // bf16, C >= 128: conv4 in natural row order (the tcgen05 path gates in a separate pass)
    float *dw_w = nullptr, *dw_b = nullptr, *cc_w = nullptr, *cc_b = nullptr;
};
struct HModel {
    wdm_hfrm_config cfg;
    std::vector<PRef> params;
    int conv_in[2], conv_out[2];
    std::vector<std::vector<BlockSpec>> enc, dec;
    std::vector<BlockSpec> mid;
    std::vector<int> down_w, down_b, up_w;
    std::vector<PwW> downs, ups;
    float *cin_w = nullptr, *cin_b = nullptr, *cout_w = nullptr, *cout_b = nullptr;

    int add(const std::string& n, long long numel) {
        PRef r;
        r.name = n, r.numel = numel;
        r.off = params.empty() ? 0 : params.back().off + params.back().numel;
        params.push_back(r);
        return (int)params.size() - 1;
    }
    void block(BlockSpec& b, const std::string& n, int C) {
        b.C = C;
        b.beta = add(n + ".beta", C);
        b.gamma = add(n + ".gamma", C);
        b.c1[0] = add(n + ".conv1.weight", 2LL * C * C), b.c1[1] = add(n + ".conv1.bias", 2 * C);
        b.c2[0] = add(n + ".conv2.weight", 2LL * C * 9), b.c2[1] = add(n + ".conv2.bias", 2 * C);
        b.c3[0] = add(n + ".conv3.weight", (long long)C * C), b.c3[1] = add(n + ".conv3.bias", C);
        b.cc[0] = add(n + ".channel_attn.chan_conv.weight", (long long)C * C), b.cc[1] = add(n + ".channel_attn.chan_conv.bias", C);
        b.c4[0] = add(n + ".conv4.weight", 2LL * C * C), b.c4[1] = add(n + ".conv4.bias", 2 * C);
        b.c5[0] = add(n + ".conv5.weight", (long long)C * C), b.c5[1] = add(n + ".conv5.bias", C);
        b.n1[0] = add(n + ".norm1.weight", C), b.n1[1] = add(n + ".norm1.bias", C);
        b.n2[0] = add(n + ".norm2.weight", C), b.n2[1] = add(n + ".norm2.bias", C);
    }
};

// arch.py:206-232 -- same module / state-dict names (the host packs the flat buffer by NAME from wdm_hfrm_param_info)
int build_hmodel(const wdm_hfrm_config& cfg, HModel* m) {
    // dim a power of two >= 32: the depthwise kernel maps 256 threads onto (pixels, 8-channel vectors)
    if (cfg.in_channel != 3 || cfg.dim < 32 || (cfg.dim & (cfg.dim - 1)) || cfg.n_levels < 1 || cfg.n_levels > 6 || cfg.mid_blk_num < 0)
        return WDM_ERR_BAD_ARG;
    if ((cfg.dim << cfg.n_levels) > 2048) return WDM_ERR_BAD_ARG;  // 256 threads x 8 channels per pixel in the dw kernel
    m->cfg = cfg;
    char buf[96];
    const int L = cfg.n_levels;
    m->conv_in[0] = m->add("conv_in.weight", 27LL * cfg.dim), m->conv_in[1] = m->add("conv_in.bias", cfg.dim);
    m->enc.resize(L), m->dec.resize(L), m->downs.resize(L), m->ups.resize(L);
    int dim = cfg.dim;
    for (int l = 0; l < L; ++l) {
        if (cfg.enc_blk_nums[l] < 0 || cfg.dec_blk_nums[l] < 0) return WDM_ERR_BAD_ARG;
        m->enc[l].resize(cfg.enc_blk_nums[l]);
        for (int i = 0; i < cfg.enc_blk_nums[l]; ++i) {
            snprintf(buf, sizeof buf, "encoders.%d.%d", l, i);
            m->block(m->enc[l][i], buf, dim);
        }
        snprintf(buf, sizeof buf, "downs.%d", l);
        m->down_w.push_back(m->add(std::string(buf) + ".weight", 2LL * dim * dim * 4));
        m->down_b.push_back(m->add(std::string(buf) + ".bias", 2 * dim));
        dim *= 2;
    }
    m->mid.resize(cfg.mid_blk_num);
    for (int i = 0; i < cfg.mid_blk_num; ++i) {
        snprintf(buf, sizeof buf, "mid_blks.%d", i);
        m->block(m->mid[i], buf, dim);
    }
    for (int l = 0; l < L; ++l) {
        snprintf(buf, sizeof buf, "ups.%d.0.weight", l);
        m->up_w.push_back(m->add(buf, 2LL * dim * dim));
        dim /= 2;
        m->dec[l].resize(cfg.dec_blk_nums[l]);
        for (int i = 0; i < cfg.dec_blk_nums[l]; ++i) {
            snprintf(buf, sizeof buf, "decoders.%d.%d", l, i);
            m->block(m->dec[l][i], buf, dim);
        }
    }
    m->conv_out[0] = m->add("conv_out.weight", 27LL * cfg.dim), m->conv_out[1] = m->add("conv_out.bias", 3);
    return WDM_OK;
}

}  // namespace
}  // namespace wdm

struct wdm_hfrm {
    wdm::HModel model;
    int dt = wdm::DT_BF16;
    char* packed = nullptr;
    size_t packed_bytes = 0;
};

namespace wdm {
namespace {

constexpr float kLnEps = 1e-6f;  // arch.py:37

// Walks the packed arena; with flat == nullptr only sizes are accumulated.
int pack_hmodel(wdm_hfrm* net, const float* flat, cudaStream_t s, size_t* total) {
    HModel& m = net->model;
    const bool dry = flat == nullptr;
    size_t off = 0;
    const size_t esz = dtype_size(net->dt);
    auto take = [&](size_t bytes) -> char* {
        char* p = net->packed ? net->packed + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    auto P = [&](int idx) { return flat + m.params[idx].off; };
    int st = WDM_OK;
    auto pack_pw = [&](PwW& o, int widx, int bidx, int N, int K, const float* colscale, const float* lnb, const float* rowscale,
                       int perm, int src_down, int Cin, bool has_bias, bool ln = false) {
        o.N = N, o.K = K;
        o.w = take((size_t)N * K * esz);
        o.b = has_bias ? reinterpret_cast<float*>(take((size_t)N * 4)) : nullptr;
        o.wsum = ln ? reinterpret_cast<float*>(take((size_t)N * 4)) : nullptr;
        if (dry || st != WDM_OK) return;
        if (net->dt == DT_F32)
            hfrm_pack_pw_kernel<float><<<N, 128, 0, s>>>(P(widx), bidx >= 0 ? P(bidx) : nullptr, colscale, lnb, rowscale, N, K, perm,
                                                         src_down, Cin, reinterpret_cast<float*>(o.w), o.b, o.wsum);
        else
            hfrm_pack_pw_kernel<bf16><<<N, 128, 0, s>>>(P(widx), bidx >= 0 ? P(bidx) : nullptr, colscale, lnb, rowscale, N, K, perm,
                                                        src_down, Cin, reinterpret_cast<bf16*>(o.w), o.b, o.wsum);
        st = wdm_launch_status();
    };
    auto misc = [&](float*& dst, int idx, int mode, int C, int total_el) {
        dst = reinterpret_cast<float*>(take((size_t)total_el * 4));
        if (dry || st != WDM_OK) return;
        hfrm_pack_misc_kernel<<<wdm_cdiv(total_el, 256), 256, 0, s>>>(P(idx), dst, mode, C, total_el);
        st = wdm_launch_status();
    };
    auto copy = [&](float*& dst, int idx, int n) {
        dst = reinterpret_cast<float*>(take((size_t)n * 4));
        if (dry || st != WDM_OK) return;
        cudaError_t e = cudaMemcpyAsync(dst, P(idx), (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) st = wdm_cuda_error((int)e);
    };
    auto block = [&](BlockSpec& b) {
        const int C = b.C;
        const float* n1w = dry ? nullptr : P(b.n1[0]);
        const float* n1b = dry ? nullptr : P(b.n1[1]);
        const float* n2w = dry ? nullptr : P(b.n2[0]);
        const float* n2b = dry ? nullptr : P(b.n2[1]);
        pack_pw(b.conv1, b.c1[0], b.c1[1], 2 * C, C, n1w, n1b, nullptr, PERM_NONE, 0, 0, true, true);
        misc(b.dw_w, b.c2[0], 0, 2 * C, 18 * C);
        copy(b.dw_b, b.c2[1], 2 * C);
        copy(b.cc_w, b.cc[0], C * C);
        copy(b.cc_b, b.cc[1], C);
        pack_pw(b.conv3, b.c3[0], b.c3[1], C, C, nullptr, nullptr, dry ? nullptr : P(b.beta), PERM_NONE, 0, 0, true);
        pack_pw(b.conv4, b.c4[0], b.c4[1], 2 * C, C, n2w, n2b, nullptr, PERM_GATE, 0, 0, true, true);
        if (net->dt == DT_BF16 && C >= 128) pack_pw(b.conv4n, b.c4[0], b.c4[1], 2 * C, C, n2w, n2b, nullptr, PERM_NONE, 0, 0, true, true);
        pack_pw(b.conv5, b.c5[0], b.c5[1], C, C, nullptr, nullptr, dry ? nullptr : P(b.gamma), PERM_NONE, 0, 0, true);
    };
    const int L = m.cfg.n_levels;
    misc(m.cin_w, m.conv_in[0], 1, m.cfg.dim, 27 * m.cfg.dim);
    copy(m.cin_b, m.conv_in[1], m.cfg.dim);
    int dim = m.cfg.dim;
    for (int l = 0; l < L; ++l) {
        for (auto& b : m.enc[l]) block(b);
        pack_pw(m.downs[l], m.down_w[l], m.down_b[l], 2 * dim, 4 * dim, nullptr, nullptr, nullptr, PERM_NONE, 1, dim, true);
        dim *= 2;
    }
    for (auto& b : m.mid) block(b);
    for (int l = 0; l < L; ++l) {
        pack_pw(m.ups[l], m.up_w[l], -1, 2 * dim, dim, nullptr, nullptr, nullptr, PERM_SHUFFLE, 0, 0, false);
        dim /= 2;
        for (auto& b : m.dec[l]) block(b);
    }
    misc(m.cout_w, m.conv_out[0], 2, m.cfg.dim, 27 * m.cfg.dim);
    copy(m.cout_b, m.conv_out[1], 3);
    if (total) *total = off;
    return st;
}

struct HArena {
    char* base = nullptr;
    size_t cap = 0, off = 0;
    bool failed = false;
    char* take(size_t bytes) {
        const size_t o = off;
        off += (bytes + 255) / 256 * 256;
        if (base && off > cap) failed = true;
        return base ? base + o : nullptr;
    }
};

template <typename T>
int hfrm_forward_t(wdm_hfrm* net, HArena& ar, const float* x, int B, int H, int W, float* y, cudaStream_t s) {
    HModel& m = net->model;
    const int L = m.cfg.n_levels, d0 = m.cfg.dim;
    const bool dry = ar.base == nullptr;
    const long long px0 = (long long)B * H * W;
    const size_t S0 = (size_t)px0 * d0 * sizeof(T);  // bytes of one [B, H, W, dim] tensor; every level's C*pixels is S0 / 2^l
    T* xa = reinterpret_cast<T*>(ar.take(S0));
    T* xb = reinterpret_cast<T*>(ar.take(S0));
    T* t1 = reinterpret_cast<T*>(ar.take(2 * S0));
    T* t2 = reinterpret_cast<T*>(ar.take(S0));
    T* t3 = reinterpret_cast<T*>(ar.take(S0));
    std::vector<T*> skip(L);
    for (int l = 0; l < L; ++l) skip[l] = reinterpret_cast<T*>(ar.take(S0 >> l));
    size_t part_el = 0;  // per image: max over the levels of (256-pixel tiles) x channels
    for (int l = 0; l <= L; ++l) {
        const size_t e = (size_t)wdm_cdiv(H >> l, kDwTH) * wdm_cdiv(W >> l, dw_tile_w(d0 << l)) * (size_t)(d0 << l);
        part_el = e > part_el ? e : part_el;
    }
    float* partial = reinterpret_cast<float*>(ar.take((size_t)B * part_el * 4));
    float* sbuf = reinterpret_cast<float*>(ar.take((size_t)B * (d0 << L) * 4));
    if (dry) return WDM_OK;
    if (ar.failed) return WDM_ERR_WORKSPACE;

    int st = WDM_OK;
    auto pw = [&](const T* A, int lda, long long M, const PwW& w, int pro, int epi, const T* res, T* out, int ldo, int a_mode, int gh,
                  int gw, int Cin, const float* scale) {
        if (st != WDM_OK) return;
        PwParams p;
        memset(&p, 0, sizeof p);
        p.A = A, p.lda = lda, p.M = (int)M, p.K = w.K, p.N = w.N;
        p.a_mode = a_mode, p.H = gh, p.W = gw, p.Cin = Cin;
        p.Wt = w.w, p.bias = w.b, p.wsum = w.wsum, p.pro = pro, p.eps = kLnEps, p.scale = scale, p.epi = epi;
        p.res = res, p.ldr = ldo, p.out = out, p.ldo = ldo;
        st = launch_pw<T>(p, s);
    };
    static const int tc_enabled = []() {
        const char* e = getenv("WDM_HFRM_TC");
        return e ? atoi(e) : 1;
    }();
    // 1x1 conv on the tcgen05 kernel: plain [M x K] . [N x K]^T with bias (+ residual), M a multiple of 128
    auto tc = [&](const T* A, long long M, const PwW& w, const T* res, T* out) {
        if (st != WDM_OK) return;
        GemmParams g;
        memset(&g, 0, sizeof g);
        g.src0 = A, g.C0 = w.K, g.ld0 = w.K, g.Hin = g.Hout = (int)(M / 128), g.Win = g.Wout = 128, g.taps = 1, g.stride = 1;
        g.B = w.w, g.ldb = w.K, g.b_layout = BL_NK, g.M = (int)M, g.N = w.N, g.K = w.K, g.alpha = 1.f;
        g.bias = w.b, g.residual = res, g.ldr = w.N, g.out = out, g.ldo = w.N;
        g.a_dtype = g.b_dtype = g.out_dtype = DT_BF16;
        st = launch_gemm_tc(g, s);
    };
    auto rowop = [&](const T* x, T* y, long long M, int C, int mode, const float* scale, long long hw) {
        if (st != WDM_OK) return;
        hfrm_row_kernel<<<wdm_cdiv(M, 8), 256, 0, s>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), M, C, mode, kLnEps,
                                                        scale, hw);
        st = wdm_launch_status();
    };
    // one ResidualBlock (arch.py:185-204): reads cur, writes dst (dst may be cur). x1 = x + beta*conv3(.) overwrites cur in
    // place (every output element only depends on the residual at the same position; the GEMM operand is t2)
    auto block = [&](const BlockSpec& b, T* cur, T* dst, int h, int w) {
        const int C = b.C;
        const long long M = (long long)B * h * w;
        const bool use_tc = std::is_same<T, bf16>::value && tc_enabled && C >= 128 && C <= 512 && (M % 128) == 0 && b.conv4n.w;
        if (use_tc) {
            rowop(cur, t3, M, C, 0, nullptr, 0);           // LayerNorm2d (norm1)
            tc(t3, M, b.conv1, nullptr, t1);
        } else {
            pw(cur, C, M, b.conv1, PRO_LN, EPI_BIAS, nullptr, t1, 2 * C, AM_PLAIN, h, w, 0, nullptr);
        }
        if (st != WDM_OK) return;
        const int tiles_x = wdm_cdiv(w, dw_tile_w(C)), tiles = tiles_x * wdm_cdiv(h, kDwTH);
        if ((long long)h * w * 2 * C >= (1LL << 31)) {  // the depthwise kernel indexes one image with 32-bit offsets
            st = WDM_ERR_BAD_SHAPE;
            return;
        }
        const dim3 dgrid(tiles, B, C > kDwCS ? C / kDwCS : 1);
        static const int dw_smem = []() {
            const char* e = getenv("WDM_HFRM_DW_SMEM");
            return e ? atoi(e) : 1;
        }();
        if (dw_smem) {
            const int ring = kDwNB * dw_band_bytes<T>(C < kDwCS ? C : kDwCS);
            static bool attr = false;
            if (!attr) {
                cudaError_t e = cudaFuncSetAttribute(hfrm_dw_gate_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
                if (e != cudaSuccess) {
                    st = wdm_cuda_error((int)e);
                    return;
                }
                attr = true;
            }
            hfrm_dw_gate_smem_kernel<T><<<dgrid, 256, ring, s>>>(t1, b.dw_w, b.dw_b, t2, partial, C, h, w, tiles_x);
        } else if (C == 32)
            hfrm_dw_gate_kernel<T, 32><<<dgrid, 256, 0, s>>>(t1, b.dw_w, b.dw_b, t2, partial, C, h, w, tiles_x);
        else if (C == 64)
            hfrm_dw_gate_kernel<T, 64><<<dgrid, 256, 0, s>>>(t1, b.dw_w, b.dw_b, t2, partial, C, h, w, tiles_x);
        else
            hfrm_dw_gate_kernel<T, 0><<<dgrid, 256, 0, s>>>(t1, b.dw_w, b.dw_b, t2, partial, C, h, w, tiles_x);
        st = wdm_launch_status();
        if (st != WDM_OK) return;
        hfrm_chan_kernel<<<dim3(B, wdm_cdiv(C, 64)), 256, (C + 256) * sizeof(float), s>>>(partial, tiles, C, 1.0f / (float)(h * w),
                                                                                         b.cc_w, b.cc_b, sbuf);
        st = wdm_launch_status();
        if (use_tc) {
            rowop(t2, t2, M, C, 1, sbuf, (long long)h * w);  // channel attention, in place
            tc(t2, M, b.conv3, cur, cur);                   // x1 = x + beta * conv3(.)
            rowop(cur, t3, M, C, 0, nullptr, 0);           // LayerNorm2d (norm2)
            tc(t3, M, b.conv4n, nullptr, t1);
            rowop(t1, t3, M, C, 2, nullptr, 0);            // SimpleGate
            tc(t3, M, b.conv5, cur, dst);                   // y = x1 + gamma * conv5(.)
            return;
        }
        pw(t2, C, M, b.conv3, PRO_SCALE, EPI_RES, cur, cur, C, AM_PLAIN, h, w, 0, sbuf);
        pw(cur, C, M, b.conv4, PRO_LN, EPI_GATE, nullptr, t3, C, AM_PLAIN, h, w, 0, nullptr);
        pw(t3, C, M, b.conv5, PRO_NONE, EPI_RES, cur, dst, C, AM_PLAIN, h, w, 0, nullptr);
    };

    {
        const int thr = 256;
        hfrm_conv_in_kernel<T><<<wdm_cdiv(px0, thr), thr, (27 * d0 + d0) * sizeof(float), s>>>(x, m.cin_w, m.cin_b, xa, px0, H, W, d0);
        st = wdm_launch_status();
    }
    int h = H, w = W, dim = d0;
    T* cur = xa;
    for (int l = 0; l < L && st == WDM_OK; ++l) {
        const int nb = (int)m.enc[l].size();
        for (int i = 0; i < nb; ++i) {
            T* dst = (i == nb - 1) ? skip[l] : cur;
            block(m.enc[l][i], cur, dst, h, w);
            cur = dst;
        }
        if (nb == 0) {  // no encoder block at this level: the skip is the level input itself
            if (st == WDM_OK) {
                cudaError_t e = cudaMemcpyAsync(skip[l], cur, (size_t)B * h * w * dim * sizeof(T), cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess) st = wdm_cuda_error((int)e);
            }
            cur = skip[l];
        }
        // downs[l] (arch.py:221): 2x2 stride-2 conv dim -> 2 dim, read straight from the skip tensor (space-to-depth addressing)
        pw(cur, dim, (long long)B * (h / 2) * (w / 2), m.downs[l], PRO_NONE, EPI_BIAS, nullptr, xa, 2 * dim, AM_S2D, h / 2, w / 2, dim,
           nullptr);
        cur = xa;
        h /= 2, w /= 2, dim *= 2;
    }
    for (size_t i = 0; i < m.mid.size() && st == WDM_OK; ++i) block(m.mid[i], cur, cur, h, w);
    for (int l = 0; l < L && st == WDM_OK; ++l) {
        // ups[l] (arch.py:228) = 1x1 conv dim -> 2 dim (no bias) + PixelShuffle(2), + the encoder skip (arch.py:246-247)
        T* nxt = cur == xa ? xb : xa;  // the shuffle writes a 2h x 2w tensor while other CTAs still read `cur`
        pw(cur, dim, (long long)B * h * w, m.ups[l], PRO_NONE, EPI_SHUFFLE, skip[L - 1 - l], nxt, dim / 2, AM_PLAIN, h, w, 0, nullptr);
        h *= 2, w *= 2, dim /= 2;
        cur = nxt;
        for (auto& b : m.dec[l]) block(b, cur, cur, h, w);
    }
    if (st != WDM_OK) return st;
    {
        const int thr = 256;
        hfrm_conv_out_kernel<T><<<wdm_cdiv(px0, thr), thr, 27 * d0 * sizeof(float), s>>>(cur, m.cout_w, m.cout_b, x, y, px0, H, W, d0);
        st = wdm_launch_status();
    }
    return st;
}

int hfrm_forward(wdm_hfrm* net, HArena& ar, const float* x, int B, int H, int W, float* y, cudaStream_t s) {
    return net->dt == DT_F32 ? hfrm_forward_t<float>(net, ar, x, B, H, W, y, s) : hfrm_forward_t<bf16>(net, ar, x, B, H, W, y, s);
}

}  // namespace
}  // namespace wdm

// ================================================================================================ C ABI
using namespace wdm;

extern "C" int wdm_hfrm_param_count(const wdm_hfrm_config* cfg) {
    if (!cfg) return WDM_ERR_BAD_ARG;
    HModel m;
    const int st = build_hmodel(*cfg, &m);
    return st != WDM_OK ? st : (int)m.params.size();
}

extern "C" int wdm_hfrm_param_info(const wdm_hfrm_config* cfg, int i, char* name, int cap, long long* numel) {
    if (!cfg) return WDM_ERR_BAD_ARG;
    HModel m;
    const int st = build_hmodel(*cfg, &m);
    if (st != WDM_OK) return st;
    if (i < 0 || i >= (int)m.params.size()) return WDM_ERR_BAD_ARG;
    if (name && cap > 0) {
        strncpy(name, m.params[i].name.c_str(), cap - 1);
        name[cap - 1] = 0;
    }
    if (numel) *numel = m.params[i].numel;
    return WDM_OK;
}

extern "C" size_t wdm_hfrm_packed_bytes(const wdm_hfrm_config* cfg, int precision) {
    if (!cfg || (precision != WDM_PREC_FP32 && precision != WDM_PREC_BF16)) return 0;
    wdm_hfrm net;
    if (build_hmodel(*cfg, &net.model) != WDM_OK) return 0;
    net.dt = precision == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
    size_t total = 0;
    pack_hmodel(&net, nullptr, 0, &total);
    return total;
}

extern "C" int wdm_hfrm_create(const wdm_hfrm_config* cfg, int precision, const float* flat_params, long long flat_numel,
                               void* packed, size_t packed_bytes, void* stream, wdm_hfrm_t** out) {
    if (!cfg || !flat_params || !packed || !out) return WDM_ERR_BAD_ARG;
    if (precision != WDM_PREC_FP32 && precision != WDM_PREC_BF16) return WDM_ERR_BAD_ARG;
    if (!wdm_aligned(packed, 256) || !wdm_aligned(flat_params, 16)) return WDM_ERR_BAD_ALIGN;
    wdm_hfrm* net = new wdm_hfrm();
    int st = build_hmodel(*cfg, &net->model);
    if (st == WDM_OK) {
        const PRef& last = net->model.params.back();
        if (flat_numel != last.off + last.numel) st = WDM_ERR_BAD_ARG;
    }
    if (st == WDM_OK) {
        net->dt = precision == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
        size_t need = 0;
        pack_hmodel(net, nullptr, 0, &need);
        if (packed_bytes < need) st = WDM_ERR_WORKSPACE;
    }
    if (st == WDM_OK) {
        net->packed = reinterpret_cast<char*>(packed);
        net->packed_bytes = packed_bytes;
        size_t total = 0;
        st = pack_hmodel(net, flat_params, static_cast<cudaStream_t>(stream), &total);
    }
    if (st != WDM_OK) {
        delete net;
        return st;
    }
    *out = net;
    return WDM_OK;
}

extern "C" void wdm_hfrm_destroy(wdm_hfrm_t* net) { delete net; }

extern "C" size_t wdm_hfrm_workspace_bytes(const wdm_hfrm_t* net, int B, int H, int W) {
    if (!net || B <= 0 || H <= 0 || W <= 0) return 0;
    HArena ar;
    hfrm_forward(const_cast<wdm_hfrm*>(net), ar, nullptr, B, H, W, nullptr, 0);
    return ar.off;
}

extern "C" int wdm_hfrm_forward(wdm_hfrm_t* net, const float* x, int B, int H, int W, float* y, void* workspace,
                                size_t workspace_bytes, void* stream) {
    if (!net || !x || !y || !workspace) return WDM_ERR_BAD_ARG;
    const int g = 1 << net->model.cfg.n_levels;
    if (B <= 0 || H <= 0 || W <= 0 || (H % g) || (W % g)) return WDM_ERR_BAD_SHAPE;  // arch.py has no padding path either
    if ((long long)B * H * W >= (1LL << 31)) return WDM_ERR_BAD_SHAPE;
    if (!wdm_aligned(workspace, 256) || !wdm_aligned(x, 16) || !wdm_aligned(y, 16)) return WDM_ERR_BAD_ALIGN;
    HArena ar;
    ar.base = reinterpret_cast<char*>(workspace);
    ar.cap = workspace_bytes;
    return hfrm_forward(net, ar, x, B, H, W, y, static_cast<cudaStream_t>(stream));
}
