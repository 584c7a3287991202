// wdm_optim.cu -- the parameter update of the training step: Adam (utils/optimize.py:6-8 ->
// torch.optim.Adam, betas (0.9, 0.999), L2 weight decay, no amsgrad) and the EMA shadow update
// (models/ddm_wavelet.py:48-53: shadow = (1 - mu) * param + mu * shadow) over EVERY parameter tensor in ONE launch.
// The reference runs ~10 foreach passes over the 156 M UNet parameters for Adam and then a Python loop of three small
// kernels per parameter tensor (~700 tensors) for the EMA; fused, an element is read once (p, g, m, v, ema = 20 B) and
// written once (p, m, v, ema = 16 B): 36 B per parameter, HBM-bound.
//
// Arithmetic: the operation order and the rounding points of torch's CUDA foreach implementation (each foreach pass rounds
// to fp32; inside one pass the functor is contracted to an FMA by nvcc), so the update is reproducible against
// torch.optim.Adam step for step:
//   g' = fma(wd, p, g)                      (_foreach_add(grads, params, alpha = wd); skipped when wd == 0)
//   m  = fma(1 - beta1, g' - m, m)          (_foreach_lerp_, weight < 0.5 branch)
//   v  = fma(1 - beta2, g' * g', v * beta2) (_foreach_mul_ then _foreach_addcmul_: a + value * (b * c))
//   d  = sqrt(v) / sqrt(1 - beta2^t) + eps  (_foreach_sqrt, _foreach_div_, _foreach_add_)
//   p  = fma(-lr / (1 - beta1^t), m / d, p) (_foreach_addcdiv_: a + value * (b / c))
//   e  = (1 - mu) * p + mu * e              (three eager kernels in the reference: both products rounded, then the sum)
// The host passes the scalars as doubles computed the way torch's Python computes them; they are cast to fp32 here exactly
// as the foreach kernels cast their Python-number arguments.
#include "wdm_common.cuh"

#include <cstdint>

namespace {
constexpr int kThreads = 256;
constexpr int kChunk = 8192;  // elements per CTA: 8 float4 per thread

struct Seg {  // one parameter tensor; mirrored by the int64 [T][6] table the host uploads
    float* p;
    const float* g;
    float* m;
    float* v;
    float* ema;
    long long n;
};

struct OptScalars {
    float wd, w1, beta2, w2, bc2_sqrt, eps, neg_step, mu, one_minus_mu;
    int adam, ema;
};

__device__ __forceinline__ void update_one(float& p, float g, float& m, float& v, float& e, const OptScalars& s) {
    if (s.adam) {
        if (s.wd != 0.f) g = __fmaf_rn(s.wd, p, g);
        m = __fmaf_rn(s.w1, __fsub_rn(g, m), m);
        v = __fmaf_rn(s.w2, __fmul_rn(g, g), __fmul_rn(v, s.beta2));
        const float d = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
        p = __fmaf_rn(s.neg_step, __fdiv_rn(m, d), p);
    }
    if (s.ema) e = __fadd_rn(__fmul_rn(s.one_minus_mu, p), __fmul_rn(s.mu, e));
}

__global__ void __launch_bounds__(kThreads) adam_ema_kernel(const Seg* __restrict__ segs, const int* __restrict__ cta_first,
                                                            int T, const OptScalars s) {
    // CTA -> (tensor, chunk): cta_first[t] = first CTA of tensor t (prefix sums of ceil(n / kChunk)), cta_first[T] = grid
    int lo = 0, hi = T;
    const int b = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cta_first[mid] <= b)
            lo = mid;
        else
            hi = mid;
    }
    const Seg sg = segs[lo];
    const long long base = (long long)(b - cta_first[lo]) * kChunk;
    const long long left = sg.n - base;
    const int cnt = left < kChunk ? (int)left : kChunk;
    float* p = sg.p + base;
    const float* g = sg.g ? sg.g + base : nullptr;
    float* m = sg.m ? sg.m + base : nullptr;
    float* v = sg.v ? sg.v + base : nullptr;
    float* e = sg.ema ? sg.ema + base : nullptr;
    const uintptr_t align = (uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)e;
    const int nvec = (align & 15) == 0 ? cnt >> 2 : 0;
    // kU vectors of every operand are requested before the first dependent op: kU x 5 16-byte loads in flight per thread
    constexpr int kU = 4;
    for (int i0 = threadIdx.x; i0 < nvec; i0 += kThreads * kU) {
        float4 P[kU], G[kU] = {}, M[kU] = {}, V[kU] = {}, E[kU] = {};
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * kThreads;
            if (i < nvec) {
                P[u] = reinterpret_cast<const float4*>(p)[i];
                if (s.adam) {
                    G[u] = __ldcs(reinterpret_cast<const float4*>(g) + i);
                    M[u] = reinterpret_cast<const float4*>(m)[i];
                    V[u] = reinterpret_cast<const float4*>(v)[i];
                }
                if (s.ema) E[u] = reinterpret_cast<const float4*>(e)[i];
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * kThreads;
            if (i < nvec) {
                update_one(P[u].x, G[u].x, M[u].x, V[u].x, E[u].x, s);
                update_one(P[u].y, G[u].y, M[u].y, V[u].y, E[u].y, s);
                update_one(P[u].z, G[u].z, M[u].z, V[u].z, E[u].z, s);
                update_one(P[u].w, G[u].w, M[u].w, V[u].w, E[u].w, s);
                if (s.adam) {
                    reinterpret_cast<float4*>(p)[i] = P[u];
                    reinterpret_cast<float4*>(m)[i] = M[u];
                    reinterpret_cast<float4*>(v)[i] = V[u];
                }
                if (s.ema) reinterpret_cast<float4*>(e)[i] = E[u];
            }
        }
    }
    for (int i = nvec * 4 + threadIdx.x; i < cnt; i += kThreads) {  // tail / unaligned tensors
        float P = p[i], G = 0.f, M = 0.f, V = 0.f, E = 0.f;
        if (s.adam) G = g[i], M = m[i], V = v[i];
        if (s.ema) E = e[i];
        update_one(P, G, M, V, E, s);
        if (s.adam) p[i] = P, m[i] = M, v[i] = V;
        if (s.ema) e[i] = E;
    }
}
}  // namespace

extern "C" int wdm_optim_chunk(void) { return kChunk; }

// segs: device int64 [T][6] = (param, grad, exp_avg, exp_avg_sq, ema shadow, numel) -- fp32 tensors, contiguous;
// cta_first: device int32 [T + 1] prefix sums of ceil(numel / wdm_optim_chunk()).
// do_adam: grad / exp_avg / exp_avg_sq must be non-null; do_ema: the shadow must be non-null.
extern "C" int wdm_adam_ema_step(const void* segs, const int* cta_first, int T, int n_ctas, int do_adam, int do_ema,
                                 double lr, double beta1, double beta2, double eps, double weight_decay, long long step,
                                 double mu, void* stream) {
    if (!segs || !cta_first) return WDM_ERR_BAD_ARG;
    if (T <= 0 || n_ctas <= 0) return WDM_OK;
    if (!do_adam && !do_ema) return WDM_ERR_BAD_ARG;
    if (do_adam && (step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0))) return WDM_ERR_BAD_ARG;
    OptScalars s;
    memset(&s, 0, sizeof s);
    s.adam = do_adam ? 1 : 0, s.ema = do_ema ? 1 : 0;
    if (do_adam) {
        // torch/optim/adam.py (_multi_tensor_adam, non-capturable): Python-double scalars, cast to fp32 by the foreach kernels
        const double bc1 = 1.0 - pow(beta1, (double)step);
        const double bc2 = 1.0 - pow(beta2, (double)step);
        s.wd = (float)weight_decay;
        s.w1 = (float)(1.0 - beta1);
        s.beta2 = (float)beta2;
        s.w2 = (float)(1.0 - beta2);
        s.bc2_sqrt = (float)pow(bc2, 0.5);  // bias_correction2 ** 0.5
        s.eps = (float)eps;
        s.neg_step = (float)((lr / bc1) * -1.0);
    }
    if (do_ema) {
        s.mu = (float)mu;
        s.one_minus_mu = (float)(1.0 - mu);
    }
    adam_ema_kernel<<<n_ctas, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const Seg*>(segs), cta_first,
                                                                               T, s);
    return wdm_launch_status();
}
