// wdm_engine.h -- internal (C++) interface between the UNet executor (wdm_unet.cu) and the kernels.
// Activations are NHWC ([P, H, W, C], channels innermost) in the engine's storage type: fp32
// (WDM_PREC_FP32, parity mode) or bf16 (WDM_PREC_BF16, tensor-core mode). All statistics, biases,
// softmax and the DDIM update are fp32 in both modes.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wavedm_b200.h"

namespace wdm {

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };
inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

enum BLayout : int { BL_NK = 0 /* B[n][k], k contiguous (weights, K matrix) */, BL_KN = 1 /* B[k][n] (V matrix) */ };

// Generic implicit-GEMM:  out[m][n] = alpha * sum_k A[m][k] * B[k][n] (+ bias[n] + temb[row(m)][n] + residual[m][n])
// Row m = (b, oy, ox), b = m / (Hout*Wout). A[m][k], k = tap*(C0+C1) + c, is gathered from up to two NHWC
// sources concatenated along channels (c < C0 -> src0 else src1):
//   taps=9: (dy,dx) = (tap/3, tap%3); iy = oy*stride + dy - pad, ix likewise; out-of-range -> 0
//   ups=1 : the source is read through a virtual nearest-neighbour x2 upsample (src row = iy >> 1)
//   taps=1: plain row-major matrix rows (1x1 conv / attention matmuls)
// (the struct itself is part of the C ABI: wdm_gemm_params in include/wavedm_b200.h)
using GemmParams = ::wdm_gemm_params;

int launch_gemm_simt(const GemmParams& p, cudaStream_t stream);
// tcgen05 / TMA path (wdm_gemm_tc.cu); returns WDM_ERR_UNSUPPORTED for shapes it does not tile.
bool gemm_tc_supported(const GemmParams& p);
int launch_gemm_tc(const GemmParams& p, cudaStream_t stream);
int gemm_tc_ksplit_plan(const GemmParams& p);  // split-K factor worth using for p (1 = none), see wdm_gemm_params::ksplit

// conv_out (models/unet.py:303-307, Cout <= 4): NHWC in -> NCHW fp32 out [P, Cout, H, W].
int launch_conv_small_cout(const void* src, int dtype, int P, int H, int W, int C, const float* w /*[Cout][9][C]*/,
                           const float* bias, int Cout, float* out_nchw, cudaStream_t stream);

// GroupNorm(32 groups, eps) statistics over up to two concatenated NHWC sources -> stats[P][32][2] = (mean, rstd)
int launch_gn_stats(const void* src0, int C0, const void* src1, int C1, int dtype, int P, int HW, float eps,
                    float* stats, cudaStream_t stream);
// same result from the side-car partial sums written by the tensor-core epilogue (wdm_gemm_params::stats_out) of the
// kernel(s) that produced the source tensor(s): sc[P*HW/32][C/4][2]
int launch_gn_finalize_sidecar(const float* sc0, int C0, const float* sc1, int C1, int P, int HW, float eps,
                               float* stats, cudaStream_t stream);
// one launch: (mean, rstd) reduced from the side-car(s) inside the normalise kernel (bf16 storage only)
int launch_gn_apply_sidecar(const void* src0, int C0, const float* sc0, const void* src1, int C1, const float* sc1, int P, int HW,
                            float eps, const float* gamma, const float* beta, int silu, void* out, cudaStream_t stream);
size_t gn_stats_bytes(int P);  // size of the `stats` scratch buffer (mean/rstd + double partial sums)
// y = ((x - mean) * rstd * gamma + beta), optionally * sigmoid(.)  -> out [P, HW, C0+C1] (materialises the concat)
int launch_gn_apply(const void* src0, int C0, const void* src1, int C1, int dtype, int P, int HW, const float* stats,
                    const float* gamma, const float* beta, int silu, void* out, cudaStream_t stream);
// nearest x2 upsample NHWC (bf16 path only; the fp32 path folds it into the conv addressing)
int launch_upsample2x(const void* src, int dtype, int P, int H, int W, int C, void* out, cudaStream_t stream);
// row softmax: S fp32 [rows][L] -> probabilities in out_dtype [rows][L]
int launch_softmax_rows(const float* S, int rows, int L, void* out, int out_dtype, cudaStream_t stream, int seg = 0);

// timestep path (models/unet.py:10-28,354-357,125): temb[T][4ch] then all per-block projections
//   out[T][total] = concat_b( Linear_b(silu(temb)) + conv1_b.bias )
struct TembParams {
    const float* t;       // [T] device
    const float* freqs;   // [ch/2] exp(-i ln(1e4)/(ch/2-1)), computed by the host exactly as the reference does
    int T, ch;            // ch = model.ch (embedding dim), temb_ch = 4*ch
    const float *w0, *b0; // dense.0 [4ch][ch], [4ch]
    const float *w1, *b1; // dense.1 [4ch][4ch], [4ch]
    const float* wp;      // all temb_proj weights stacked [total][4ch]
    const float* bp;      // stacked temb_proj.bias + conv1.bias [total]
    int total;
    float* scratch;       // [T][2*4ch]
    float* out;           // [T][total]
};
int launch_temb(const TembParams& p, cudaStream_t stream);

// Patch gather: builds the UNet input [P, R, R, Cpad] (NHWC, engine dtype) from up to three fp32 NCHW image-level
// tensors [B, Cs, h, w] cropped at the patch corners (models/ddm_wavelet.py:467-478); channels >= sum(Cs) are zero.
struct GatherParams {
    const float* src[3];
    int Cs[3];
    int nsrc;
    int B, h, w;
    const int* patches;  // [P][3] = (image, hi, wi) device
    int P, R, Cpad;
    void* out;
    int out_dtype;
};
int launch_gather_patches(const GatherParams& p, cudaStream_t stream);
// one source's channels rewritten in place at channel offset c_off (x_t between DDIM steps)
int launch_gather_update(const float* src, int C, int c_off, int h, int w, const int* patches, int P, int R, int Cpad,
                         void* out, int out_dtype, cudaStream_t stream);
// wavelet_in_unet (wdm_dwt.cu): fused crop + DWT + concat + NHWC gather, and the IWT of the NHWC conv_out result
int launch_dwt_gather(const float* src0, const float* src1, int nsrc, int B, int H, int W, const int* patches, int P, int R,
                      int Cpad, void* out, int out_dtype, cudaStream_t stream);
int launch_iwt_nhwc(const float* y, int ld, int P, int R, float* x, cudaStream_t stream);
// the first C columns of a row-major fp32 matrix [P*HW][ld] -> NCHW [P][C][HW] (conv_out with more than 4 channels: use_window)
int launch_rows_to_nchw(const float* y, int ld, int P, int HW, int C, float* x, cudaStream_t stream);

// Fused overlap-average + DDIM update (models/ddm_wavelet.py:485-503, eta = 0), per image pixel:
//   et = sum_{patches covering the pixel} eps_patch / count ; x0 = (xt - et*sqrt(1-at))/sqrt(at)
//   xt_next = sqrt(at_next)*x0 + sqrt(1-at_next)*et
struct DdimParams {
    const float* eps;      // [P][Cp][R][R] patch outputs (NCHW fp32)
    const int* patches;    // [P][3]
    const int* img_first;  // [B+1] patch range of each image (patches sorted by image)
    int P, B, Cp, R, h, w;
    const float* xt;       // [B][Cp][h][w]
    float* x0_out;         // [B][Cp][h][w]
    float* xt_next;        // [B][Cp][h][w] (may alias xt)
    float at, at_next;
};
int launch_ddim_step(const DdimParams& p, cudaStream_t stream);

// weight packing: OIHW fp32 [Cout][Cin][kh][kw] -> [Cout][kh*kw][Cin_pad] in dtype, at column offset k_off of a
// packed matrix with row pitch ldk (so several convs can share one K-concatenated matrix)
int launch_pack_conv_weight(const float* w, int Cout, int Cin, int taps, int Cin_pad, void* out, int out_dtype,
                            long long ldk, long long k_off, cudaStream_t stream);

// tc32 (fp32 contractions on the tensor cores through a 3-way bf16 split, see wdm_elem.cu / GemmParams::a_split3)
int launch_split3_act(const float* src0, int C0, const float* src1, int C1, long long rows, void* out, cudaStream_t stream);
// out_main == null: [N][taps][6][C]; else the dominant product apart: out_main [N][taps][C], out [N][taps][5][C]
int launch_split3_weight(const float* w, int N, int taps, int C, void* out, void* out_main, cudaStream_t stream);

int launch_vec_add(const float* a, const float* b, float* out, int n, cudaStream_t stream);
// pack-time C x C fp32 products: mode 0 out = X^T Y, mode 1 out = X Y; matvec: out = X^T v / X v (+ add)
int launch_matmul_cc(const float* X, const float* Y, float* out, int C, int mode, cudaStream_t stream);
int launch_matvec_c(const float* X, const float* v, const float* add, float* out, int C, int mode, cudaStream_t stream);
// sub-pixel decomposition of (nearest x2 upsample -> 3x3 conv): [4 phases][Cout][4 taps][Cin] (see wdm_elem.cu)
int launch_pack_subpix_weight(const float* w, int Cout, int Cin, void* out, int out_dtype, cudaStream_t stream);

}  // namespace wdm
