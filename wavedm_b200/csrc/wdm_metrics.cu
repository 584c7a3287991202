// wdm_metrics.cu -- the PSNR statistics restore() prints per image (reference models/restoration.py:142-146 calling
// utils/metrics.py:7-11 torchPSNR, :43-51 calculate_psnr_in_GPU(test_y_channel=True), :53-86 calculate_psnr(test_y_channel=
// True) on the [0,255] clamp), batched: ONE launch per image pair set instead of three full-image reductions + two host
// round trips per image. SURVEY.md 8(f)-2.
//   sse[b][0] = sum over 3*H*W of (clamp01(a) - clamp01(b))^2                      -> torchPSNR
//   sse[b][1] = sum over H*W of (Y(a) - Y(b))^2, Y = (24.966 r + 128.553 g + 65.481 b + 16) / 255  (unclamped inputs)
//   sse[b][2] = the same on clamp01 inputs (the numpy variant works on clamp(x*255, 0, 255))
// One CTA per image, fixed-order tree reduction in double: deterministic.
#include "wdm_common.cuh"

namespace {
constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads) psnr_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, int HW,
                                                             double* __restrict__ sse) {
    __shared__ double red[3][kThreads / 32];
    const long long img = blockIdx.x;
    const float* pa = a + img * 3 * HW;
    const float* pb = b + img * 3 * HW;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < HW; i += kThreads) {
        float da[3], db[3], ca[3], cb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            da[c] = pa[c * HW + i], db[c] = pb[c * HW + i];
            ca[c] = fminf(fmaxf(da[c], 0.f), 1.f), cb[c] = fminf(fmaxf(db[c], 0.f), 1.f);
            const float d = ca[c] - cb[c];
            s0 += (double)(d * d);
        }
        const float ya = (24.966f * da[0] + 128.553f * da[1] + 65.481f * da[2] + 16.0f) / 255.f;
        const float yb = (24.966f * db[0] + 128.553f * db[1] + 65.481f * db[2] + 16.0f) / 255.f;
        const float yca = (24.966f * ca[0] + 128.553f * ca[1] + 65.481f * ca[2] + 16.0f) / 255.f;
        const float ycb = (24.966f * cb[0] + 128.553f * cb[1] + 65.481f * cb[2] + 16.0f) / 255.f;
        s1 += (double)((ya - yb) * (ya - yb));
        s2 += (double)((yca - ycb) * (yca - ycb));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[0][warp] = s0, red[1][warp] = s1, red[2][warp] = s2;
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[threadIdx.x][w];
        sse[img * 3 + threadIdx.x] = s;
    }
}
}  // namespace

extern "C" int wdm_psnr_stats(const float* a, const float* b, int B, int H, int W, double* sse, void* stream) {
    if (!a || !b || !sse) return WDM_ERR_BAD_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || (long long)H * W > 0x7fffffffLL / 3) return WDM_ERR_BAD_SHAPE;
    psnr_stats_kernel<<<B, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, b, H * W, sse);
    return wdm_launch_status();
}
