/*
 * oracle/dwt_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's wavelet transform for scale=2:
 *   /root/reference/models/wavelet.py:6-50   (WaveletTransform: stride-4 4x4 groups=3 conv with the
 *                                             fixed `rec4` weights + the channel permute at :40-43 / :45-48)
 *   /root/reference/models/wavelet_weights_c2.pkl ['rec4']  (48,1,4,4) fp32, every |w| = 0.25
 *
 * Closed form (re-verified against the pickle by oracle/make_golden.py):
 *   sub-band k = b0 + 2 b1 + 4 b2 + 8 b3, block pixel (r, c), r = 2 r_hi + r_lo, c = 2 c_hi + c_lo
 *   Wk[r][c] = 0.25 * (-1)^(b0 c_hi + b1 r_hi + b2 c_lo + b3 r_lo)
 *   DWT: y[n, 3k+g, i, j] = sum_{r,c} Wk[r][c] x[n, g, 4i+r, 4j+c]
 *   IWT: x[n, g, 4i+r, 4j+c] = sum_k Wk[r][c] y[n, 3k+g, i, j]
 *
 * Two evaluations are provided:
 *   *_direct   : the 16-term dot product in (r, c) raster order, exactly what a conv does per output
 *                (order of summation of a conv backend is unspecified -> compared with a tolerance);
 *   *_lifting  : the 4-stage butterfly order that the CUDA kernels use (c_lo, c_hi, r_lo, r_hi, then an
 *                exact *0.25). The CUDA kernels must match this one BIT-EXACTLY.
 *
 * Parity pinning: the reference has no tests / golden vectors. This oracle is pinned
 * against outputs of the reference module itself generated in the build container by
 * oracle/make_golden.py and committed under tests/golden/.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stddef.h>
#include <stdint.h>

#define WDM_ORACLE_API __attribute__((visibility("default")))

static inline float wsign(int k, int r, int c) {
    int b0 = k & 1, b1 = (k >> 1) & 1, b2 = (k >> 2) & 1, b3 = (k >> 3) & 1;
    int e = b0 * (c >> 1) + b1 * (r >> 1) + b2 * (c & 1) + b3 * (r & 1);
    return (e & 1) ? -0.25f : 0.25f;
}

/* ---- direct (conv-like) form -------------------------------------------------------------------- */

WDM_ORACLE_API void wdm_oracle_dwt4x4_direct(const float* x, float* y, int n, int H, int W) {
    const int h = H / 4, w = W / 4;
    for (int b = 0; b < n; ++b)
        for (int g = 0; g < 3; ++g)
            for (int k = 0; k < 16; ++k)
                for (int i = 0; i < h; ++i)
                    for (int j = 0; j < w; ++j) {
                        float acc = 0.f;
                        for (int r = 0; r < 4; ++r)
                            for (int c = 0; c < 4; ++c)
                                acc += wsign(k, r, c) *
                                       x[(((size_t)b * 3 + g) * H + (4 * i + r)) * W + (4 * j + c)];
                        y[(((size_t)b * 48 + (3 * k + g)) * h + i) * w + j] = acc;
                    }
}

WDM_ORACLE_API void wdm_oracle_iwt4x4_direct(const float* y, float* x, int n, int h, int w) {
    const int H = 4 * h, W = 4 * w;
    for (int b = 0; b < n; ++b)
        for (int g = 0; g < 3; ++g)
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < w; ++j)
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) {
                            float acc = 0.f;
                            for (int k = 0; k < 16; ++k)
                                acc += wsign(k, r, c) *
                                       y[(((size_t)b * 48 + (3 * k + g)) * h + i) * w + j];
                            x[(((size_t)b * 3 + g) * H + (4 * i + r)) * W + (4 * j + c)] = acc;
                        }
}

/* ---- lifting (butterfly) form: the order the CUDA kernels follow --------------------------------- */

/* 16-point 4-D Walsh-Hadamard butterfly on v[r][c] -> out[k], k = b0 + 2 b1 + 4 b2 + 8 b3.
 * Stage order: c_lo (-> b2), c_hi (-> b0), r_lo (-> b3), r_hi (-> b1), then * 0.25. */
static inline void wht16_fwd(const float v[4][4], float out[16]) {
    float u[4][2][2]; /* u[r][b0][b2] */
    for (int r = 0; r < 4; ++r) {
        float s00 = v[r][0] + v[r][1]; /* c_hi=0, b2=0 */
        float s01 = v[r][0] - v[r][1]; /* c_hi=0, b2=1 */
        float s10 = v[r][2] + v[r][3]; /* c_hi=1, b2=0 */
        float s11 = v[r][2] - v[r][3]; /* c_hi=1, b2=1 */
        u[r][0][0] = s00 + s10;
        u[r][1][0] = s00 - s10;
        u[r][0][1] = s01 + s11;
        u[r][1][1] = s01 - s11;
    }
    for (int b0 = 0; b0 < 2; ++b0)
        for (int b2 = 0; b2 < 2; ++b2) {
            float t00 = u[0][b0][b2] + u[1][b0][b2]; /* r_hi=0, b3=0 */
            float t01 = u[0][b0][b2] - u[1][b0][b2]; /* r_hi=0, b3=1 */
            float t10 = u[2][b0][b2] + u[3][b0][b2]; /* r_hi=1, b3=0 */
            float t11 = u[2][b0][b2] - u[3][b0][b2]; /* r_hi=1, b3=1 */
            out[b0 + 0 + 4 * b2 + 0] = (t00 + t10) * 0.25f; /* b1=0,b3=0 */
            out[b0 + 2 + 4 * b2 + 0] = (t00 - t10) * 0.25f; /* b1=1,b3=0 */
            out[b0 + 0 + 4 * b2 + 8] = (t01 + t11) * 0.25f; /* b1=0,b3=1 */
            out[b0 + 2 + 4 * b2 + 8] = (t01 - t11) * 0.25f; /* b1=1,b3=1 */
        }
}

/* Inverse: in[k] -> v[r][c]. Stage order: b2 (-> c_lo), b0 (-> c_hi), b3 (-> r_lo), b1 (-> r_hi), * 0.25. */
static inline void wht16_inv(const float in[16], float v[4][4]) {
    float u[2][2][4]; /* u[b1][b3][c] */
    for (int b1 = 0; b1 < 2; ++b1)
        for (int b3 = 0; b3 < 2; ++b3) {
            const float* p = in + 2 * b1 + 8 * b3; /* p[b0 + 4 b2] */
            float s00 = p[0] + p[4];               /* b0=0, c_lo=0 */
            float s01 = p[0] - p[4];               /* b0=0, c_lo=1 */
            float s10 = p[1] + p[5];               /* b0=1, c_lo=0 */
            float s11 = p[1] - p[5];               /* b0=1, c_lo=1 */
            u[b1][b3][0] = s00 + s10;              /* c_hi=0, c_lo=0 */
            u[b1][b3][2] = s00 - s10;              /* c_hi=1, c_lo=0 */
            u[b1][b3][1] = s01 + s11;              /* c_hi=0, c_lo=1 */
            u[b1][b3][3] = s01 - s11;              /* c_hi=1, c_lo=1 */
        }
    for (int c = 0; c < 4; ++c) {
        float t00 = u[0][0][c] + u[0][1][c]; /* b1=0, r_lo=0 */
        float t01 = u[0][0][c] - u[0][1][c]; /* b1=0, r_lo=1 */
        float t10 = u[1][0][c] + u[1][1][c]; /* b1=1, r_lo=0 */
        float t11 = u[1][0][c] - u[1][1][c]; /* b1=1, r_lo=1 */
        v[0][c] = (t00 + t10) * 0.25f;       /* r_hi=0, r_lo=0 */
        v[2][c] = (t00 - t10) * 0.25f;       /* r_hi=1, r_lo=0 */
        v[1][c] = (t01 + t11) * 0.25f;       /* r_hi=0, r_lo=1 */
        v[3][c] = (t01 - t11) * 0.25f;       /* r_hi=1, r_lo=1 */
    }
}

/* flags mirror include/wavedm_b200.h: bit0 on DWT = apply data_transform 2x-1 on load
 * (restoration.py:8-9); bit0 on IWT = apply inverse_data_transform clamp((x+1)/2,0,1) on store
 * (restoration.py:12-13). */
WDM_ORACLE_API void wdm_oracle_dwt4x4(const float* x, float* y, int n, int H, int W, int flags) {
    const int h = H / 4, w = W / 4;
    for (int b = 0; b < n; ++b)
        for (int g = 0; g < 3; ++g)
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < w; ++j) {
                    float v[4][4], o[16];
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) {
                            float t = x[(((size_t)b * 3 + g) * H + (4 * i + r)) * W + (4 * j + c)];
                            v[r][c] = (flags & 1) ? (2.0f * t - 1.0f) : t;
                        }
                    wht16_fwd(v, o);
                    for (int k = 0; k < 16; ++k)
                        y[(((size_t)b * 48 + (3 * k + g)) * h + i) * w + j] = o[k];
                }
}

WDM_ORACLE_API void wdm_oracle_iwt4x4(const float* y, float* x, int n, int h, int w, int flags) {
    const int H = 4 * h, W = 4 * w;
    for (int b = 0; b < n; ++b)
        for (int g = 0; g < 3; ++g)
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < w; ++j) {
                    float in[16], v[4][4];
                    for (int k = 0; k < 16; ++k)
                        in[k] = y[(((size_t)b * 48 + (3 * k + g)) * h + i) * w + j];
                    wht16_inv(in, v);
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) {
                            float t = v[r][c];
                            if (flags & 1) {
                                t = (t + 1.0f) / 2.0f;
                                t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
                            }
                            x[(((size_t)b * 3 + g) * H + (4 * i + r)) * W + (4 * j + c)] = t;
                        }
                }
}

/* The 48x16 weight table in the reference's `rec4` layout [g*16+k][r][c] (all three colour groups are
 * identical), for checking the closed form against the pickle. */
WDM_ORACLE_API void wdm_oracle_rec4(float* w48x16) {
    for (int g = 0; g < 3; ++g)
        for (int k = 0; k < 16; ++k)
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) w48x16[((g * 16 + k) * 4 + r) * 4 + c] = wsign(k, r, c);
}
