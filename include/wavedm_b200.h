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
/* number of CUDA kernels this library has launched in this process (bench.py's "gpu_launches") */
WDM_API long long wdm_launch_counter(void);

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
/* models/restoration.py:111-135 in one kernel: IWT of cat([lo[:, :Clo], hi-bands]) (+ WDM_IWT_POST_CLAMP) without the
 * concatenated tensor. lo: [n, >=Clo.. exactly Clo channels used, tensor has Clo channels]; hi: [n, Chi, h, w] with
 * Chi == 48 (full wavelet tensor, channels [Clo, 48) used) or Chi == 48 - Clo (only the remaining bands). */
WDM_API int wdm_iwt4x4_cat(const float* lo, int Clo, const float* hi, int Chi, float* x, int n, int h, int w, int flags,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conditional diffusion UNet engine.
 * Replaces  models/unet.py:196-395  (DiffusionUNet.__init__/forward; identical network:
 * models/unet_wav.py:10-155) with use_window=False, wavelet_in_unet=False: sinusoidal timestep embedding
 * (:10-28), ResnetBlock (:81-138), AttnBlock (:141-193), Downsample (:59-78), Upsample (:40-56),
 * GroupNorm(32, eps=1e-6)+SiLU (:31-37).
 *
 * Precision modes
 *   WDM_PREC_FP32 : fp32 storage + fp32 FFMA contractions (parity mode: |d| < 1e-3 per pixel gate)
 *   WDM_PREC_BF16 : bf16 storage, tcgen05 tensor-core contractions with fp32 accumulation in TMEM,
 *                   fp32 statistics / softmax / timestep path (throughput mode: PSNR gate)
 *
 * Parameters are passed as ONE flat fp32 device buffer holding the reference's state_dict tensors
 * (models/unet.py state-dict names, OIHW conv weights) concatenated in the canonical order reported by
 * wdm_unet_param_info(), followed by the extra tensor "temb.freqs" (ch/2 floats, the frequency table of
 * unet.py:19-21 computed by the host exactly as the reference computes it). wdm_unet_create() re-packs
 * them into `packed` (caller-allocated, wdm_unet_packed_bytes()) and never touches `flat_params` again.
 * ------------------------------------------------------------------------------------------------ */
#define WDM_PREC_FP32 0
#define WDM_PREC_BF16 1
/* engine flags (wdm_unet_create) */
#define WDM_ENGINE_NO_TC 0x1 /* bf16 mode: use the CUDA-core GEMM for every contraction (debug / A-B testing) */
/* bf16 mode: allow a contraction whose shape the tcgen05 kernel does not tile to run on the CUDA-core kernel. Without
 * this flag such a shape makes wdm_unet_forward return WDM_ERR_UNSUPPORTED -- there is no silent fallback. (Odd patch
 * counts are NOT such a shape: the engine pads them internally.) */
#define WDM_ENGINE_ALLOW_SIMT 0x2
/* fp32 mode: run the contractions on the tcgen05 kernel through the 3-way bf16 operand split (six bf16 products per
 * element pair accumulate in fp32 in TMEM: x*w = (x0+x1+x2)(w0+w1+w2) minus the three terms below 2^-24) instead of
 * CUDA-core FFMA. Storage, GroupNorm, softmax, residuals stay fp32. Shapes the tensor-core kernel does not tile (conv_in's
 * 96 channels, the attention bmm's, conv_out) stay on the FFMA kernel. */
#define WDM_ENGINE_TC32 0x4

typedef struct wdm_unet_config {
    int ch;             /* model.ch */
    int n_levels;       /* len(model.ch_mult) <= 8 */
    int ch_mult[8];
    int num_res_blocks; /* model.num_res_blocks */
    int n_attn_res;     /* len(model.attn_resolutions) <= 8 */
    int attn_res[8];
    int resolution;     /* data.image_size (patch side R) */
    int in_channels;    /* UNet input channels (models/unet.py:212), 96 for raindrop_wavelet.yml */
    int out_ch;         /* model.out_ch (<= 4; 48 with wavelet_in_unet) */
    int wavelet_in_unet; /* data.wavelet_in_unet (models/unet.py:203-206,349-350,393-394): the network consumes the DWT of
                            its pixel-domain input (built by wdm_gather_patches_dwt) and wdm_unet_forward applies the IWT
                            to the 48-channel conv_out result: eps_out is [P, 3, 4R, 4R] */
} wdm_unet_config;

typedef struct wdm_unet wdm_unet_t;

WDM_API int wdm_unet_param_count(const wdm_unet_config* cfg);
/* i-th parameter tensor in canonical order: its state_dict name (NUL-terminated into name[0..cap)) and numel. */
WDM_API int wdm_unet_param_info(const wdm_unet_config* cfg, int i, char* name, int cap, long long* numel);
/* upper bound over the engine flags (WDM_PREC_FP32: includes the WDM_ENGINE_TC32 split weights); _flags: exact */
WDM_API size_t wdm_unet_packed_bytes(const wdm_unet_config* cfg, int precision);
WDM_API size_t wdm_unet_packed_bytes_flags(const wdm_unet_config* cfg, int precision, int flags);
WDM_API int wdm_unet_create(const wdm_unet_config* cfg, int precision, int flags, const float* flat_params,
                            long long flat_numel, void* packed, size_t packed_bytes, void* stream,
                            wdm_unet_t** out);
WDM_API void wdm_unet_destroy(wdm_unet_t* net);
/* channels of the NHWC input tensor (in_channels rounded up to the engine's K granule) */
WDM_API int wdm_unet_input_channels_padded(const wdm_unet_t* net);
WDM_API size_t wdm_unet_workspace_bytes(const wdm_unet_t* net, int P);
/* x: [P, R, R, Cpad] NHWC in the engine's storage type (as produced by wdm_gather_patches);
 * t: device fp32 [T], T == 1 (one timestep for all patches, the sampler's case: ddm_wavelet.py:457) or T == P;
 * eps_out: [P, out_ch, R, R] fp32 NCHW ([P, 3, 4R, 4R] with wavelet_in_unet). */
WDM_API int wdm_unet_forward(wdm_unet_t* net, const void* x, const float* t, int T, int P, float* eps_out,
                             void* workspace, size_t workspace_bytes, void* stream);
/* Per-kernel-class timing of the contraction (conv / GEMM) launches for the roofline report: while enabled,
 * wdm_unet_forward brackets every contraction launch with CUDA events on the launch stream.
 * wdm_unet_profile_read synchronises those events, returns the totals since the last read and resets:
 *   tc_*   : tcgen05 tensor-core kernel launches;  simt_* : CUDA-core kernel launches. */
WDM_API int wdm_unet_profile_enable(wdm_unet_t* net, int on);
WDM_API int wdm_unet_profile_read(wdm_unet_t* net, double* tc_ms, double* tc_flops, long long* tc_launches,
                                  double* simt_ms, double* simt_flops, long long* simt_launches);
/* contraction launches since wdm_unet_create, by kernel class: tcgen05 tensor-core / CUDA-core (always counted, also
 * without profiling). In WDM_PREC_BF16 mode without WDM_ENGINE_NO_TC / WDM_ENGINE_ALLOW_SIMT, *simt stays 0. */
WDM_API int wdm_unet_counters(const wdm_unet_t* net, long long* tc_launches, long long* simt_launches);
/* algorithmic HBM bytes (each operand / result tensor counted once) of the tensor-core launches since the last call */
WDM_API double wdm_unet_profile_tc_bytes(wdm_unet_t* net);

/* ------------------------------------------------------------------------------------------------
 * restore() metrics. Replaces the three per-image PSNR reductions of  models/restoration.py:142-146  (utils/metrics.py:
 * 7-11 torchPSNR, :43-51 calculate_psnr_in_GPU(.., True), :53-86 calculate_psnr(.., True)) by one batched launch:
 *   a, b : [B, 3, H, W] fp32 NCHW;  sse : [B][3] doubles (device) =
 *   { sum (clamp01(a)-clamp01(b))^2 over 3HW,  sum (Y(a)-Y(b))^2 over HW,  sum (Y(clamp01 a)-Y(clamp01 b))^2 over HW },
 *   Y = (24.966 c0 + 128.553 c1 + 65.481 c2 + 16) / 255. PSNR = 10 log10(n / sse). Deterministic (fixed-order reduction).
 * ------------------------------------------------------------------------------------------------ */
WDM_API int wdm_psnr_stats(const float* a, const float* b, int B, int H, int W, double* sse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HFRM engine -- the one-shot high-frequency refinement CNN of restore().
 * Replaces  models/arch.py:206-253  (HFRM.__init__/forward: conv_in, encoder / middle / decoder ResidualBlocks
 * arch.py:158-204 = LayerNorm2d + 1x1 + depthwise 3x3 + SimpleGate + channel attention + 1x1, 2x2 stride-2 downs,
 * 1x1 + PixelShuffle ups, conv_out + input residual) at its call site  models/restoration.py:94
 * (`self.diffusion.generator(x[:, :3])`, also models/ddm_wavelet.py:367).
 *   x, y : [B, 3, H, W] fp32 NCHW, H and W multiples of 2^n_levels (the reference has no padding path either)
 * Precision: WDM_PREC_FP32 (fp32 storage, FFMA: parity mode) / WDM_PREC_BF16 (bf16 storage, tensor-core MMAs with fp32
 * accumulation; LayerNorm statistics, depthwise conv, pooling and channel attention in fp32).
 * Parameters: ONE flat fp32 device buffer holding the reference's state_dict tensors concatenated in the order reported
 * by wdm_hfrm_param_info() (names = the reference module's state_dict keys, e.g. "encoders.0.1.conv4.weight").
 * ------------------------------------------------------------------------------------------------ */
typedef struct wdm_hfrm_config {
    int in_channel;      /* 3 */
    int dim;             /* 32: channels at full resolution, a power of two >= 32 */
    int mid_blk_num;     /* 6 */
    int n_levels;        /* len(enc_blk_nums) == len(dec_blk_nums), <= 6 */
    int enc_blk_nums[8]; /* [2, 2, 2, 4] in models/ddm_wavelet.py:137 */
    int dec_blk_nums[8]; /* [2, 2, 2, 2] */
} wdm_hfrm_config;

typedef struct wdm_hfrm wdm_hfrm_t;

WDM_API int wdm_hfrm_param_count(const wdm_hfrm_config* cfg);
WDM_API int wdm_hfrm_param_info(const wdm_hfrm_config* cfg, int i, char* name, int cap, long long* numel);
WDM_API size_t wdm_hfrm_packed_bytes(const wdm_hfrm_config* cfg, int precision);
WDM_API int wdm_hfrm_create(const wdm_hfrm_config* cfg, int precision, const float* flat_params, long long flat_numel,
                            void* packed, size_t packed_bytes, void* stream, wdm_hfrm_t** out);
WDM_API void wdm_hfrm_destroy(wdm_hfrm_t* net);
WDM_API size_t wdm_hfrm_workspace_bytes(const wdm_hfrm_t* net, int B, int H, int W);
WDM_API int wdm_hfrm_forward(wdm_hfrm_t* net, const float* x, int B, int H, int W, float* y, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training-step parameter update. Replaces  utils/optimize.py:6-8  (torch.optim.Adam: betas as given,
 * L2 weight decay, amsgrad off) followed by  models/ddm_wavelet.py:48-53  (EMAHelper.update: shadow = (1 - mu) * param +
 * mu * shadow on the UPDATED parameters) by one launch over all parameter tensors, with the rounding points of torch's
 * CUDA foreach implementation (csrc/wdm_optim.cu).
 *   segs      : device int64 [T][6] = (param, grad, exp_avg, exp_avg_sq, ema_shadow, numel), fp32 contiguous tensors;
 *   cta_first : device int32 [T + 1], prefix sums of ceil(numel / wdm_optim_chunk()); n_ctas = cta_first[T];
 *   do_adam / do_ema select the two halves (Adam only: optimizer.step();  EMA only: ema_helper.update(); both: fused);
 *   step      : the 1-based Adam step count of THIS update (bias corrections 1 - beta^step).
 * ------------------------------------------------------------------------------------------------ */
WDM_API int wdm_optim_chunk(void);
WDM_API int wdm_adam_ema_step(const void* segs, const int* cta_first, int T, int n_ctas, int do_adam, int do_ema, double lr,
                              double beta1, double beta2, double eps, double weight_decay, long long step, double mu,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampler kernels.
 * wdm_gather_patches replaces the crop + cat of  models/ddm_wavelet.py:467-478  (and the NCHW->NHWC
 * conversion of the module-level forward): out[p, y, x, c] = concat_s(src_s)[img_p, c, hi_p+y, wi_p+x],
 * channels >= sum(Cs) zero-filled up to Cpad. srcs are fp32 NCHW [B, Cs, h, w]; patches is device int32
 * [P][3] = (image, hi, wi). out_dtype: WDM_PREC_FP32 / WDM_PREC_BF16.
 * wdm_ddim_step replaces  models/ddm_wavelet.py:485-503  (scatter-add of patch outputs, division by the
 * overlap count, x0 prediction, eta=0 DDIM update) in one kernel; img_first is device int32 [B+1] with the
 * patch range of every image (patches sorted by image, reference corner order within an image).
 * ------------------------------------------------------------------------------------------------ */
WDM_API int wdm_gather_patches(const float* src0, int C0, const float* src1, int C1, const float* src2, int C2, int B,
                               int h, int w, const int* patches, int P, int R, int Cpad, void* out, int out_dtype,
                               void* stream);
/* In-place refresh of one source inside a tensor wdm_gather_patches produced: out[p, y, x, c_off + c] = src[img_p, c, ..]
 * for c < C, all other channels kept. Between DDIM steps only x_t changes (models/ddm_wavelet.py:470-478 re-crops and
 * re-concatenates x_cond / x_other every step although they are loop invariants): the sampler gathers once and then
 * rewrites the C = 3 channels of x_t at c_off = channels(x_cond). */
WDM_API int wdm_gather_patches_update(const float* src, int C, int c_off, int B, int h, int w, const int* patches, int P,
                                      int R, int Cpad, void* out, int out_dtype, void* stream);
/* wavelet_in_unet mode: crop + DWT + concat + NHWC in one kernel. src0 / src1: fp32 NCHW [B, 3, H, W] pixel-domain images
 * (already data_transform'ed), patches: (image, hi, wi) in PIXELS, patch side 4R; out: [P, R, R, Cpad], channel
 * s*48 + 3k + colour (models/unet.py:338-344 all_wavlet_dec on the crops of models/ddm_wavelet.py:467-478). */
WDM_API int wdm_gather_patches_dwt(const float* src0, const float* src1, int nsrc, int B, int H, int W, const int* patches,
                                   int P, int R, int Cpad, void* out, int out_dtype, void* stream);
/* IWT of a row-major fp32 matrix [P*R*R, ld] whose first 48 columns are the sub-bands (3k + colour) -> [P, 3, 4R, 4R] */
WDM_API int wdm_iwt4x4_nhwc(const float* y, int ld, int P, int R, float* x, void* stream);
WDM_API int wdm_ddim_step(const float* eps, const int* patches, const int* img_first, int P, int B, int Cp, int R,
                          int h, int w, const float* xt, float* x0_out, float* xt_next, float at, float at_next,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-kernel entry points (unit tests / micro-benchmarks). Semantics in csrc/wdm_engine.h.
 * ------------------------------------------------------------------------------------------------ */
typedef struct wdm_gemm_params {
    const void* src0;
    const void* src1;
    int C0, C1, ld0, ld1, Hin, Win, Hout, Wout, taps, stride, pad, ups;
    const void* B;
    long long b_batch_stride;
    int ldb, b_layout, M, N, K;
    float alpha;
    const float* bias;
    const float* temb;
    int temb_rows, temb_ld;
    const void* residual;
    int ldr;
    void* out;
    int ldo, a_dtype, b_dtype, out_dtype;
    /* optional GroupNorm side-car (tensor-core path only): per (32-row group, 4-column block) partial sums of the
     * values written to `out`: stats_out[M/32][N/4][2] = (sum, sum of squares), fp32. NULL = not produced. */
    float* stats_out;
    /* != 0: the A operand is ONE matrix shared by every batch (rows m % (Hout*Wout)); used with a per-batch B to
     * compute V^T = Wv * h^T for attention. */
    int a_shared;
    /* != 0: sources 1 and 2 are 1x1 "tails" appended after the main taps instead of channel-concatenated inputs:
     *   K = taps*C0 + C1 + C2,  A[m][taps*C0 + c] = (c < C1 ? src1 : src2)[pixel m][c]   (centre tap, stride 1)
     * This is conv2 + nin_shortcut of a ResnetBlock (models/unet.py:128-138) as ONE contraction:
     *   x + conv2(h) with x = nin(cat[x1, x2])  ==  [W2 | Wnin] . [im2col(h) ; x1 ; x2] + (b2 + bnin). */
    int tail_1x1;
    const void* src2;
    int C2, ld2;
    /* tensor-core path only, N == one tile (64/128/256): the epilogue applies a row softmax to alpha*acc and stores the
     * probabilities (bf16) instead of the scores (models/unet.py:180-182). softmax_seg > 0: block-diagonal, row r only
     * attends to columns [seg*((r/seg) % (N/seg)), +seg), the rest get probability 0. */
    int fuse_softmax, softmax_seg;
    /* tensor-core path only: > 0 = store only the first `out_nchw_valid` (<= 32) output columns, as fp32 NCHW planes
     * out[(patch*valid + n)*Hout*Wout + pixel] (conv_out with Cout = 3 zero-padded to a 64-wide N tile). */
    int out_nchw_valid;
    /* tensor-core path only. row_scale_out (with fuse_softmax): the probabilities are stored UNNORMALISED
     * (exp(s - max), in (0, 1]) and 1 / sum of row m goes to row_scale_out[m] -- one pass over the scores instead of two,
     * shared by both epilogue warp groups. row_scale: the epilogue multiplies accumulator row m by row_scale[m] before
     * bias / residual (the P.V product that consumes those probabilities). */
    float* row_scale_out;
    const float* row_scale;
    /* tensor-core path, fp32 emulation ("tc32"): a_split3 != 0 means src0 is the 3-way bf16 split of an fp32 tensor, laid out
     * [rows][3*C0] = [hi | mid | lo] with x ~= hi + mid + lo (each piece the bf16 rounding of the remainder), ld0 its row
     * pitch, and B the matching 6-product arrangement of the split weights [N][taps][6][C0] =
     * (w_hi, w_mid, w_hi, w_mid, w_lo, w_hi) that pair with the A pieces (hi, hi, mid, mid, hi, lo): K = taps*6*C0,
     * out_dtype must be fp32. Built by wdm_split3_act / wdm_split3_weight. a_split3 == 2: the same without the dominant
     * (hi, hi) product, B = [N][taps][5][C0], K = taps*5*C0 -- the executor runs that product as a second plain launch whose
     * epilogue adds this result (the accumulator truncates on every add: five sixths of the truncations then happen on a
     * sum 2^-8 of the result). B3 / B3m are used by the UNet executor only (the split weights of a convolution whose `B`
     * holds the fp32 weights: five small products / dominant product). */
    int a_split3;
    const void* B3;
    const void* B3m;
    /* tensor-core path only, small problems (few output tiles, deep K: single-image latency): ksplit = S > 1 runs the K loop
     * on S CTAs per output tile, raw fp32 partial results go to ksplit_scratch ([S][M][N] floats, 16-byte aligned) and a
     * second launch reduces them in split order and applies alpha / bias / temb / residual / stats_out. S must not exceed
     * wdm_gemm_ksplit_plan(p) (1 = not applicable: the plain launch). 0 / 1: off. */
    int ksplit;
    void* ksplit_scratch;
} wdm_gemm_params;
#define WDM_GEMM_IMPL_SIMT 0
#define WDM_GEMM_IMPL_TC 1
WDM_API int wdm_gemm(const wdm_gemm_params* p, int impl, void* stream);
WDM_API int wdm_gemm_ksplit_plan(const wdm_gemm_params* p);
/* tc32 helpers (unit tests; the executor calls the same kernels): split an fp32 matrix (two channel-concatenated sources)
 * into [rows][3*(C0+C1)] bf16 pieces / a packed fp32 weight matrix [N][taps][C] into [N][taps][6][C] bf16. */
WDM_API int wdm_split3_act(const float* src0, int C0, const float* src1, int C1, long long rows, void* out, void* stream);
/* out_main == NULL: out = [N][taps][6][C]; else out_main = [N][taps][C] (w_hi) and out = [N][taps][5][C] */
WDM_API int wdm_split3_weight(const float* w, int N, int taps, int C, void* out, void* out_main, void* stream);
WDM_API size_t wdm_groupnorm_scratch_bytes(int P);
WDM_API int wdm_groupnorm_silu(const void* src0, int C0, const void* src1, int C1, int dtype, int P, int HW,
                               float eps, const float* gamma, const float* beta, int silu, void* out,
                               void* scratch, void* stream);
/* bf16 only: GroupNorm(32, eps)(+SiLU) of the concat of two NHWC tensors whose statistics come from the side-cars the
 * tensor-core epilogue writes (wdm_gemm_params::stats_out, sc[P*HW/32][C/4][2] = per 32-row group and 4-channel block (sum,
 * sum of squares)): finalise + normalise in ONE launch, the form the UNet engine uses (models/unet.py:36-37,31-33). */
WDM_API int wdm_groupnorm_silu_sidecar(const void* src0, int C0, const float* sc0, const void* src1, int C1, const float* sc1,
                                       int P, int HW, float eps, const float* gamma, const float* beta, int silu, void* out,
                                       void* stream);
WDM_API int wdm_softmax_rows(const float* S, int rows, int L, void* out, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVEDM_B200_H_ */
