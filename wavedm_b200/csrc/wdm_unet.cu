// wdm_unet.cu -- the UNet executor: parameter table, weight packing, a straight-line forward schedule over a
// caller-provided workspace (first-fit arena, deterministic addresses -> CUDA-graph capturable), and the
// C ABI around it. Mirrors models/unet.py:196-395 of the reference (see include/wavedm_b200.h).
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "wdm_common.cuh"
#include "wdm_engine.h"

namespace wdm {
namespace {

constexpr float kGnEps = 1e-6f;

// ------------------------------------------------------------------------------------------------ specs
struct ParamRef {
    std::string name;
    long long numel = 0;
    long long off = 0;  // offset in the flat fp32 parameter buffer
};

struct ConvSpec {
    int Cin = 0, Cout = 0, taps = 0;
    int Cin_pad = 0;          // packed channels per tap
    int w = -1, b = -1;       // param indices (weight, bias)
    void* pw = nullptr;       // packed weights [Cout][taps*Cin_pad] in engine dtype
    float* pb = nullptr;      // bias fp32 [Cout]
    float* pw32 = nullptr;    // fp32 copy (conv_out only)
    float* pb_pad = nullptr;  // conv_out on the tensor-core path: bias zero-padded to 64
    void* pw_subpix = nullptr;  // upsample convs, bf16 mode: [4 phases][Cout][4 taps][Cin] (sub-pixel decomposition)
    // tc32 (fp32 mode on the tensor cores, see wdm_elem.cu): the 3-way bf16 split of the weights, dominant product apart:
    void* pw3 = nullptr;        //   the five small products [Cout][taps][5][Cin]
    void* pw3m = nullptr;       //   the (hi, hi) product    [Cout][taps][Cin]
    void* pw_subpix3 = nullptr, *pw_subpix3m = nullptr;  // the same for the sub-pixel weights [4 phases][Cout][4 taps][..][Cin]
};
struct GnSpec {
    int C = 0, w = -1, b = -1;
    float *gamma = nullptr, *beta = nullptr;
};
struct ResSpec {
    int Cin = 0, Cout = 0;
    GnSpec norm1, norm2;
    ConvSpec conv1, conv2, nin;
    ConvSpec conv2f;  // bf16 mode, has_nin: conv2 and nin_shortcut K-concatenated [Cout][9*Cout + Cin], bias = b2 + bnin
    bool has_nin = false;
    int temb_w = -1, temb_b = -1;
    int temb_off = 0;
};
struct AttnSpec {
    int C = 0;
    GnSpec norm;
    int q[2], k[2], v[2];
    ConvSpec qkv, proj;  // qkv: fused [3C][C]
    // Algebraic fusion (bf16 tensor-core mode): softmax_j(q_i.k_j) only sees  h_i^T (Wq^T Wk) h_j + (Wk^T bq).h_j , and
    // proj(P v) = P (h (Wp Wv)^T) + (Wp bv + bp)  because the rows of P sum to one. With G = Wk^T Wq, u = Wk^T bq,
    // Wpv = Wp Wv, bo = Wp bv + bp (fp32 products at pack time, then bf16) the block is FOUR contractions --
    // g = h G^T + u;  w^T = Wpv h^T;  P = softmax(g h^T C^-1/2);  out = P w + bo + x -- instead of five, with 40 % fewer
    // FLOPs (no k, no separate proj_out).
    void* gw = nullptr;     // [C][C] G
    float* gb = nullptr;    // [C]    u
    void* wpv = nullptr;    // [C][C] Wp Wv
    float* bo = nullptr;    // [C]    Wp bv + bp
};
struct LevelSpec {
    std::vector<ResSpec> blocks;
    std::vector<AttnSpec> attns;
    bool has_resample = false;
    ConvSpec resample;
};

struct Model {
    wdm_unet_config cfg;
    std::vector<ParamRef> params;
    std::map<std::string, int> index;
    int temb_d0[2], temb_d1[2];
    int freqs = -1;
    ConvSpec conv_in, conv_out;
    std::vector<LevelSpec> down, up;
    ResSpec mid1, mid2;
    AttnSpec mid_attn;
    GnSpec norm_out;
    int temb_total = 0;
    std::vector<ResSpec*> res_order;  // every ResnetBlock, in temb_off order

    int add(const std::string& n, long long numel) {
        ParamRef r;
        r.name = n;
        r.numel = numel;
        r.off = params.empty() ? 0 : params.back().off + params.back().numel;
        params.push_back(r);
        index[n] = (int)params.size() - 1;
        return (int)params.size() - 1;
    }
    void conv(ConvSpec& c, const std::string& n, int ci, int co, int k) {
        c.Cin = ci, c.Cout = co, c.taps = k * k;
        c.w = add(n + ".weight", (long long)co * ci * k * k);
        c.b = add(n + ".bias", co);
    }
    void gn(GnSpec& g, const std::string& n, int C) {
        g.C = C;
        g.w = add(n + ".weight", C);
        g.b = add(n + ".bias", C);
    }
    void res(ResSpec& r, const std::string& n, int ci, int co) {
        const int tc = cfg.ch * 4;
        r.Cin = ci, r.Cout = co;
        gn(r.norm1, n + ".norm1", ci);
        conv(r.conv1, n + ".conv1", ci, co, 3);
        r.temb_w = add(n + ".temb_proj.weight", (long long)co * tc);
        r.temb_b = add(n + ".temb_proj.bias", co);
        gn(r.norm2, n + ".norm2", co);
        conv(r.conv2, n + ".conv2", co, co, 3);
        r.has_nin = ci != co;
        if (r.has_nin) conv(r.nin, n + ".nin_shortcut", ci, co, 1);
        r.temb_off = temb_total;
        temb_total += co;
    }
    void attn(AttnSpec& a, const std::string& n, int C) {
        a.C = C;
        gn(a.norm, n + ".norm", C);
        ConvSpec t;
        conv(t, n + ".q", C, C, 1), a.q[0] = t.w, a.q[1] = t.b;
        conv(t, n + ".k", C, C, 1), a.k[0] = t.w, a.k[1] = t.b;
        conv(t, n + ".v", C, C, 1), a.v[0] = t.w, a.v[1] = t.b;
        conv(a.proj, n + ".proj_out", C, C, 1);
        a.qkv.Cin = C, a.qkv.Cout = 3 * C, a.qkv.taps = 1;
    }
};

bool attn_at(const wdm_unet_config& c, int res) {
    for (int i = 0; i < c.n_attn_res; ++i)
        if (c.attn_res[i] == res) return true;
    return false;
}

// models/unet.py:196-307 -- same construction order, so the parameter list is the reference's module order.
int build_model(const wdm_unet_config& cfg, Model* m) {
    if (cfg.ch <= 0 || cfg.ch % 128 || cfg.n_levels < 1 || cfg.n_levels > 8 || cfg.num_res_blocks < 1 ||
        cfg.n_attn_res < 0 || cfg.n_attn_res > 8 || cfg.in_channels < 1 || cfg.out_ch < 1)
        return WDM_ERR_BAD_ARG;
    // wavelet_in_unet: the DWT of the two 3-channel halves feeds the network, the IWT consumes 48 output channels
    // out_ch <= 4: the reference configs; up to 64: data.use_window (out_ch = 3 * window_size^2), conv_out through the GEMM path
    if (cfg.wavelet_in_unet ? (cfg.out_ch != 48 || cfg.in_channels != 96) : cfg.out_ch > 64) return WDM_ERR_BAD_ARG;
    if (cfg.resolution < (1 << (cfg.n_levels - 1)) * 2 || (cfg.resolution % (1 << (cfg.n_levels - 1))))
        return WDM_ERR_BAD_SHAPE;
    for (int i = 0; i < cfg.n_levels; ++i)
        if (cfg.ch_mult[i] < 1) return WDM_ERR_BAD_ARG;
    m->cfg = cfg;
    const int ch = cfg.ch, tc = 4 * ch, L = cfg.n_levels, nrb = cfg.num_res_blocks;
    char buf[128];
    m->temb_d0[0] = m->add("temb.dense.0.weight", (long long)tc * ch);
    m->temb_d0[1] = m->add("temb.dense.0.bias", tc);
    m->temb_d1[0] = m->add("temb.dense.1.weight", (long long)tc * tc);
    m->temb_d1[1] = m->add("temb.dense.1.bias", tc);
    m->conv(m->conv_in, "conv_in", cfg.in_channels, ch, 3);
    int cur = cfg.resolution;
    int block_in = ch;
    m->down.resize(L);
    m->up.resize(L);
    // reserve so that ResSpec addresses stay valid
    for (int lv = 0; lv < L; ++lv) {
        m->down[lv].blocks.resize(nrb);
        m->up[lv].blocks.resize(nrb + 1);
    }
    for (int lv = 0; lv < L; ++lv) {
        block_in = ch * (lv == 0 ? 1 : cfg.ch_mult[lv - 1]);
        const int block_out = ch * cfg.ch_mult[lv];
        for (int ib = 0; ib < nrb; ++ib) {
            snprintf(buf, sizeof buf, "down.%d.block.%d", lv, ib);
            m->res(m->down[lv].blocks[ib], buf, block_in, block_out);
            m->res_order.push_back(&m->down[lv].blocks[ib]);
            block_in = block_out;
            if (attn_at(cfg, cur)) {
                snprintf(buf, sizeof buf, "down.%d.attn.%d", lv, ib);
                m->down[lv].attns.emplace_back();
                m->attn(m->down[lv].attns.back(), buf, block_in);
            }
        }
        if (lv != L - 1) {
            snprintf(buf, sizeof buf, "down.%d.downsample.conv", lv);
            m->down[lv].has_resample = true;
            m->conv(m->down[lv].resample, buf, block_in, block_in, 3);
            cur /= 2;
        }
    }
    m->res(m->mid1, "mid.block_1", block_in, block_in);
    m->res_order.push_back(&m->mid1);
    m->attn(m->mid_attn, "mid.attn_1", block_in);
    m->res(m->mid2, "mid.block_2", block_in, block_in);
    m->res_order.push_back(&m->mid2);
    for (int lv = L - 1; lv >= 0; --lv) {
        const int block_out = ch * cfg.ch_mult[lv];
        int skip_in = ch * cfg.ch_mult[lv];
        for (int ib = 0; ib < nrb + 1; ++ib) {
            if (ib == nrb) skip_in = ch * (lv == 0 ? 1 : cfg.ch_mult[lv - 1]);
            snprintf(buf, sizeof buf, "up.%d.block.%d", lv, ib);
            m->res(m->up[lv].blocks[ib], buf, block_in + skip_in, block_out);
            m->res_order.push_back(&m->up[lv].blocks[ib]);
            block_in = block_out;
            if (attn_at(cfg, cur)) {
                snprintf(buf, sizeof buf, "up.%d.attn.%d", lv, ib);
                m->up[lv].attns.emplace_back();
                m->attn(m->up[lv].attns.back(), buf, block_in);
            }
        }
        if (lv != 0) {
            snprintf(buf, sizeof buf, "up.%d.upsample.conv", lv);
            m->up[lv].has_resample = true;
            m->conv(m->up[lv].resample, buf, block_in, block_in, 3);
            cur *= 2;
        }
    }
    m->gn(m->norm_out, "norm_out", block_in);
    m->conv(m->conv_out, "conv_out", block_in, cfg.out_ch, 3);
    m->freqs = m->add("temb.freqs", ch / 2);
    return WDM_OK;
}

// ------------------------------------------------------------------------------------------------ arena
// First-fit allocator over a caller-provided buffer. In `dry` mode nothing is dereferenced or launched and
// only the peak is tracked (wdm_unet_workspace_bytes). The allocation sequence is a pure function of the
// schedule, so addresses are identical on every call with the same workspace (graph-capture safe).
struct Arena {
    char* base = nullptr;
    size_t cap = 0;
    size_t peak = 0;
    bool dry = false;
    bool failed = false;
    struct Blk {
        size_t off, size;
    };
    std::vector<Blk> used;  // sorted by off
    static size_t align(size_t n) { return (n + 1023) & ~(size_t)1023; }
    void* alloc(size_t n) {
        n = align(n ? n : 1);
        size_t pos = 0;
        size_t i = 0;
        for (; i < used.size(); ++i) {
            if (used[i].off - pos >= n) break;
            pos = used[i].off + used[i].size;
        }
        if (!dry && pos + n > cap) {
            failed = true;
            return nullptr;
        }
        used.insert(used.begin() + i, Blk{pos, n});
        if (pos + n > peak) peak = pos + n;
        return base + pos;
    }
    void free(void* p) {
        if (!p && !dry) return;
        const size_t off = (size_t)((char*)p - base);
        for (size_t i = 0; i < used.size(); ++i)
            if (used[i].off == off) {
                used.erase(used.begin() + i);
                return;
            }
    }
};

struct Act {
    void* p = nullptr;
    int H = 0, W = 0, C = 0;
    float* stats = nullptr;  // GroupNorm side-car [P*H*W/32][C/4][2] written by the producing tensor-core kernel
};

}  // namespace
}  // namespace wdm

using namespace wdm;

struct wdm_unet {
    Model model;
    int dt = DT_F32;  // storage dtype of activations / packed weights
    int flags = 0;
    int cin_pad = 0;
    char* packed = nullptr;
    size_t packed_bytes = 0;
    // packed small tensors
    float *w0 = nullptr, *b0 = nullptr, *w1 = nullptr, *b1 = nullptr, *wp = nullptr, *bp = nullptr, *freqs = nullptr;
    // profiling (wdm_unet_profile_*)
    bool profile = false;
    struct Span {
        cudaEvent_t a, b;
        double flops;
        bool tc;
        int M, N, K, taps, W, tag;  // shape of the launch (WDM_PROFILE_DUMP)
    };
    std::vector<Span> spans;
    double tc_bytes = 0;  // algorithmic operand/result bytes of the profiled tensor-core launches
    long long n_tc = 0, n_simt = 0;  // contraction launches since creation, by kernel class (wdm_unet_counters)
    ~wdm_unet() {
        for (auto& s : spans) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
    }
};

namespace wdm {
namespace {

int k_granule(int dt) { return dt == DT_BF16 ? 64 : 16; }

// Lays out (and, when `net->packed` is set, fills) the packed arena. Returns total bytes via *total.
int pack_model(wdm_unet* net, const float* flat, cudaStream_t s, size_t* total) {
    Model& m = net->model;
    const int dt = net->dt;
    const size_t es = dtype_size(dt);
    const bool fill = flat != nullptr;
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* {
        char* p = net->packed ? net->packed + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    int st = WDM_OK;
    auto copy_f32 = [&](int param, float** dst) {
        const ParamRef& r = m.params[param];
        *dst = reinterpret_cast<float*>(take((size_t)r.numel * 4));
        if (fill && st == WDM_OK) {
            cudaError_t e = cudaMemcpyAsync(*dst, flat + r.off, (size_t)r.numel * 4, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) st = wdm_cuda_error((int)e);
        }
    };
    // tc32: 3-way bf16 split of a packed fp32 matrix [rows][taps][C] -> [rows][taps][6][C]
    const bool tc32 = dt == DT_F32 && (net->flags & WDM_ENGINE_TC32);
    auto split_w = [&](void** dst, void** dst_main, const void* src_f32, int rows, int taps, int C) {
        *dst = *dst_main = nullptr;
        if (!tc32 || (C % 64) || (rows % 64)) return;
        *dst = take((size_t)rows * taps * 5 * C * 2);
        *dst_main = take((size_t)rows * taps * C * 2);
        if (fill && st == WDM_OK)
            st = launch_split3_weight(reinterpret_cast<const float*>(src_f32), rows, taps, C, *dst, *dst_main, s);
    };
    auto pack_conv = [&](ConvSpec& c, int cin_pad) {
        c.Cin_pad = cin_pad;
        const long long ldk = (long long)c.taps * cin_pad;
        c.pw = take((size_t)c.Cout * ldk * es);
        if (fill && st == WDM_OK)
            st = launch_pack_conv_weight(flat + m.params[c.w].off, c.Cout, c.Cin, c.taps, cin_pad, c.pw, dt, ldk, 0, s);
        copy_f32(c.b, &c.pb);
        split_w(&c.pw3, &c.pw3m, c.pw, c.Cout, c.taps, cin_pad);
    };
    auto pack_gn = [&](GnSpec& g) {
        copy_f32(g.w, &g.gamma);
        copy_f32(g.b, &g.beta);
    };
    auto pack_res = [&](ResSpec& r) {
        pack_gn(r.norm1);
        pack_conv(r.conv1, r.Cin);
        pack_gn(r.norm2);
        pack_conv(r.conv2, r.Cout);
        if (r.has_nin) pack_conv(r.nin, r.Cin);
        if (r.has_nin && dt == DT_BF16) {
            ConvSpec& f = r.conv2f;
            f.Cin = r.Cout, f.Cout = r.Cout, f.taps = 9, f.Cin_pad = r.Cout;
            const long long ldk = 9LL * r.Cout + r.Cin;
            f.pw = take((size_t)r.Cout * ldk * es);
            f.pb = reinterpret_cast<float*>(take((size_t)r.Cout * 4));
            if (fill && st == WDM_OK)
                st = launch_pack_conv_weight(flat + m.params[r.conv2.w].off, r.Cout, r.Cout, 9, r.Cout, f.pw, dt, ldk, 0, s);
            if (fill && st == WDM_OK)
                st = launch_pack_conv_weight(flat + m.params[r.nin.w].off, r.Cout, r.Cin, 1, r.Cin, f.pw, dt, ldk,
                                             9LL * r.Cout, s);
            if (fill && st == WDM_OK)
                st = launch_vec_add(flat + m.params[r.conv2.b].off, flat + m.params[r.nin.b].off, f.pb, r.Cout, s);
        }
    };
    auto pack_attn = [&](AttnSpec& a) {
        pack_gn(a.norm);
        const int C = a.C;
        a.qkv.Cin_pad = C;
        a.qkv.pw = take((size_t)3 * C * C * es);
        a.qkv.pb = reinterpret_cast<float*>(take((size_t)3 * C * 4));
        if (fill && st == WDM_OK) {
            const int* src[3] = {a.q, a.k, a.v};
            for (int i = 0; i < 3 && st == WDM_OK; ++i) {
                st = launch_pack_conv_weight(flat + m.params[src[i][0]].off, C, C, 1, C,
                                             (char*)a.qkv.pw + (size_t)i * C * C * es, dt, C, 0, s);
                cudaError_t e = cudaMemcpyAsync(a.qkv.pb + i * C, flat + m.params[src[i][1]].off, (size_t)C * 4,
                                                cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess && st == WDM_OK) st = wdm_cuda_error((int)e);
            }
        }
        split_w(&a.qkv.pw3, &a.qkv.pw3m, a.qkv.pw, 3 * C, 1, C);
        pack_conv(a.proj, C);
        if (dt == DT_BF16) {
            a.gw = take((size_t)C * C * es);
            a.gb = reinterpret_cast<float*>(take((size_t)C * 4));
            a.wpv = take((size_t)C * C * es);
            a.bo = reinterpret_cast<float*>(take((size_t)C * 4));
            float* tmp = reinterpret_cast<float*>(take((size_t)C * C * 4));  // fp32 product before the bf16 rounding
            if (fill && st == WDM_OK) {
                const float* Wq = flat + m.params[a.q[0]].off;
                const float* bq = flat + m.params[a.q[1]].off;
                const float* Wk = flat + m.params[a.k[0]].off;
                const float* Wv = flat + m.params[a.v[0]].off;
                const float* bv = flat + m.params[a.v[1]].off;
                const float* Wp = flat + m.params[a.proj.w].off;
                const float* bp = flat + m.params[a.proj.b].off;
                st = launch_matmul_cc(Wk, Wq, tmp, C, 0, s);                                        // G = Wk^T Wq
                if (st == WDM_OK) st = launch_pack_conv_weight(tmp, C, C, 1, C, a.gw, dt, C, 0, s);
                if (st == WDM_OK) st = launch_matvec_c(Wk, bq, nullptr, a.gb, C, 0, s);              // u = Wk^T bq
                if (st == WDM_OK) st = launch_matmul_cc(Wp, Wv, tmp, C, 1, s);                      // Wpv = Wp Wv
                if (st == WDM_OK) st = launch_pack_conv_weight(tmp, C, C, 1, C, a.wpv, dt, C, 0, s);
                if (st == WDM_OK) st = launch_matvec_c(Wp, bv, bp, a.bo, C, 1, s);                  // bo = Wp bv + bp
            }
        }
    };

    const int g = k_granule(dt);
    net->cin_pad = (m.cfg.in_channels + g - 1) / g * g;
    copy_f32(m.temb_d0[0], &net->w0);
    copy_f32(m.temb_d0[1], &net->b0);
    copy_f32(m.temb_d1[0], &net->w1);
    copy_f32(m.temb_d1[1], &net->b1);
    copy_f32(m.freqs, &net->freqs);
    // stacked temb projections
    const int tc = 4 * m.cfg.ch;
    net->wp = reinterpret_cast<float*>(take((size_t)m.temb_total * tc * 4));
    net->bp = reinterpret_cast<float*>(take((size_t)m.temb_total * 4));
    if (fill) {
        for (ResSpec* r : m.res_order) {
            if (st != WDM_OK) break;
            cudaError_t e = cudaMemcpyAsync(net->wp + (size_t)r->temb_off * tc, flat + m.params[r->temb_w].off,
                                            (size_t)r->Cout * tc * 4, cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(net->bp + r->temb_off, flat + m.params[r->temb_b].off, (size_t)r->Cout * 4,
                                    cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) st = wdm_cuda_error((int)e);
        }
    }
    pack_conv(m.conv_in, net->cin_pad);
    for (auto& lv : m.down) {
        for (auto& r : lv.blocks) pack_res(r);
        for (auto& a : lv.attns) pack_attn(a);
        if (lv.has_resample) pack_conv(lv.resample, lv.resample.Cin);
    }
    pack_res(m.mid1);
    pack_attn(m.mid_attn);
    pack_res(m.mid2);
    for (auto& lv : m.up) {
        for (auto& r : lv.blocks) pack_res(r);
        for (auto& a : lv.attns) pack_attn(a);
        if (lv.has_resample) {
            pack_conv(lv.resample, lv.resample.Cin);
            if (dt == DT_BF16) {
                ConvSpec& c = lv.resample;
                c.pw_subpix = take((size_t)16 * c.Cout * c.Cin * es);
                if (fill && st == WDM_OK)
                    st = launch_pack_subpix_weight(flat + m.params[c.w].off, c.Cout, c.Cin, c.pw_subpix, dt, s);
            } else if (tc32) {
                // the pre-summed phase weights in fp32 (scratch inside the packed arena), then their split
                ConvSpec& c = lv.resample;
                float* tmp = reinterpret_cast<float*>(take((size_t)16 * c.Cout * c.Cin * 4));
                if (fill && st == WDM_OK)
                    st = launch_pack_subpix_weight(flat + m.params[c.w].off, c.Cout, c.Cin, tmp, DT_F32, s);
                split_w(&c.pw_subpix3, &c.pw_subpix3m, tmp, 4 * c.Cout, 4, c.Cin);
            }
        }
    }
    pack_gn(m.norm_out);
    // conv_out: fp32 [Cout][9][C] for the small-Cout kernel
    {
        ConvSpec& c = m.conv_out;
        c.Cin_pad = c.Cin;
        // rows (and bias) zero-padded to a multiple of 8: the CUDA-core GEMM tiles N in eights (Cout > 4: data.use_window)
        const int cout8 = (c.Cout + 7) / 8 * 8;
        c.pw32 = reinterpret_cast<float*>(take((size_t)cout8 * 9 * c.Cin * 4));
        float* pb8 = reinterpret_cast<float*>(take((size_t)cout8 * 4));
        if (fill && st == WDM_OK) {
            cudaError_t e = cudaMemsetAsync(c.pw32, 0, (size_t)cout8 * 9 * c.Cin * 4, s);
            if (e == cudaSuccess) e = cudaMemsetAsync(pb8, 0, (size_t)cout8 * 4, s);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(pb8, flat + m.params[c.b].off, (size_t)c.Cout * 4, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) st = wdm_cuda_error((int)e);
        }
        if (fill && st == WDM_OK)
            st = launch_pack_conv_weight(flat + m.params[c.w].off, c.Cout, c.Cin, 9, c.Cin, c.pw32, DT_F32, 9LL * c.Cin, 0, s);
        c.pb = pb8;
        if (dt == DT_BF16 && c.Cout <= 64 && (c.Cin % 64) == 0) {
            // tensor-core conv_out: Cout zero-padded to one 64-wide N tile
            c.pw = take((size_t)64 * 9 * c.Cin * es);
            float* pb64 = reinterpret_cast<float*>(take(64 * 4));
            if (fill && st == WDM_OK) {
                cudaError_t e = cudaMemsetAsync(c.pw, 0, (size_t)64 * 9 * c.Cin * es, s);
                if (e == cudaSuccess) e = cudaMemsetAsync(pb64, 0, 64 * 4, s);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(pb64, flat + m.params[c.b].off, (size_t)c.Cout * 4, cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess) st = wdm_cuda_error((int)e);
                if (st == WDM_OK)
                    st = launch_pack_conv_weight(flat + m.params[c.w].off, c.Cout, c.Cin, 9, c.Cin, c.pw, dt, 9LL * c.Cin, 0, s);
            }
            c.pb_pad = pb64;
        }
    }
    *total = off;
    return st;
}

// ------------------------------------------------------------------------------------------------ forward
struct Ctx {
    wdm_unet* net;
    Arena* ar;
    cudaStream_t s;
    int P, T;
    // Odd patch counts (the real RainDrop image is 120x180 -> 45 patches): every activation is ALLOCATED for Pa = P + 1
    // patches, so the two places that need an even count -- the 8x8 attention (two 64-token patches share a 128-row tile)
    // and the sub-pixel upsample-conv from an 8x8 grid (phase-major m-tiles) -- run on Pa patches; the slack patch holds
    // don't-care values that never reach a real patch (rows of a contraction are independent, the softmax is
    // block-diagonal). No CUDA-core fallback for such shapes.
    int Pa;
    int st = WDM_OK;
    float* temb = nullptr;   // [T][temb_total]
    float* gn_scratch = nullptr;
    bool dry() const { return ar->dry; }
    void fail(int e) {
        if (st == WDM_OK && e != WDM_OK) st = e;
    }
};

Act new_act(Ctx& c, int H, int W, int C) {
    Act a;
    a.H = H, a.W = W, a.C = C;
    a.p = c.ar->alloc((size_t)c.Pa * H * W * C * dtype_size(c.net->dt));
    if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
    return a;
}
void free_act(Ctx& c, Act& a) {
    c.ar->free(a.p);
    a.p = nullptr;
    if (a.stats) {
        c.ar->free(a.stats);
        a.stats = nullptr;
    }
}

bool will_use_tc(const Ctx& c, const GemmParams& p) {
    return c.net->dt == DT_BF16 && !(c.net->flags & WDM_ENGINE_NO_TC) && gemm_tc_supported(p);
}

int run_gemm_impl(Ctx& c, const GemmParams& p);

// tc32: the fp32 contraction p on the tcgen05 kernel -- split the A operand (the channel concat of src0 / src1) into its
// three bf16 pieces in arena scratch and contract with the pre-split weights p.B3 (six products per tap). tc32_plan returns
// false when the shape stays on the CUDA-core kernel (conv_in's 96 channels, batched attention products, virtual
// upsampling); q = the tensor-core problem with the A pointer still to be filled in.
bool tc32_plan(const Ctx& c, const GemmParams& p, GemmParams* q) {
    if (c.net->dt != DT_F32 || !(c.net->flags & WDM_ENGINE_TC32) || (c.net->flags & WDM_ENGINE_NO_TC) || !p.B3 || !p.B3m) return false;
    if (p.b_batch_stride || p.a_shared || p.tail_1x1 || p.ups == 1 || p.out_dtype != DT_F32 || p.fuse_softmax) return false;
    const int C = p.C0 + p.C1;
    *q = p;
    q->src1 = nullptr, q->C1 = 0, q->ld1 = 0;
    q->C0 = C, q->ld0 = 3 * C;
    q->B = p.B3, q->ldb = p.taps * 5 * C, q->K = p.taps * 5 * C;
    q->a_dtype = q->b_dtype = DT_BF16, q->a_split3 = 2;
    q->src0 = reinterpret_cast<void*>(16);  // alignment placeholder for the support check
    return gemm_tc_supported(*q);
}

bool run_gemm_tc32(Ctx& c, const GemmParams& p) {
    GemmParams q;
    if (!tc32_plan(c, p, &q)) return false;
    const int C = p.C0 + p.C1;
    const long long rows = (long long)((p.M + p.Hout * p.Wout - 1) / (p.Hout * p.Wout)) * p.Hin * p.Win;
    void* sp = c.ar->alloc((size_t)rows * 3 * C * 2);
    float* tmp = reinterpret_cast<float*>(c.ar->alloc((size_t)p.M * p.N * sizeof(float)));
    if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
    if (!c.dry() && c.st == WDM_OK) {
        c.fail(launch_split3_act(reinterpret_cast<const float*>(p.src0), p.C0, reinterpret_cast<const float*>(p.src1), p.C1, rows, sp,
                                 c.s));
        // 1. the five small products (+ bias, timestep row, residual) -> tmp
        q.src0 = sp, q.out = tmp, q.ldo = p.N;
        if (c.st == WDM_OK) run_gemm_impl(c, q);
        // 2. the dominant (hi, hi) product: a plain bf16 contraction over the first C channels of the split tensor, + tmp
        GemmParams m = q;
        m.a_split3 = 0, m.B = p.B3m, m.ldb = p.taps * C, m.K = p.taps * C;
        m.bias = nullptr, m.temb = nullptr, m.temb_rows = 0;
        m.residual = tmp, m.ldr = p.N, m.out = p.out, m.ldo = p.ldo;
        if (c.st == WDM_OK) run_gemm_impl(c, m);
    }
    c.ar->free(tmp);
    c.ar->free(sp);
    return true;
}

int run_gemm(Ctx& c, const GemmParams& p) {
    if (run_gemm_tc32(c, p)) return c.st;
    if (will_use_tc(c, p)) {
        // few output tiles under a deep K loop (single-image latency): split-K over otherwise idle SMs (wdm_gemm_tc.cu)
        const int S = gemm_tc_ksplit_plan(p);
        if (S > 1) {
            GemmParams q = p;
            q.ksplit = S;
            q.ksplit_scratch = c.ar->alloc((size_t)S * p.M * p.N * sizeof(float));
            if (c.ar->failed) {
                c.fail(WDM_ERR_WORKSPACE);
                return c.st;
            }
            run_gemm_impl(c, q);
            c.ar->free(q.ksplit_scratch);
            return c.st;
        }
    }
    return run_gemm_impl(c, p);
}

int run_gemm_impl(Ctx& c, const GemmParams& p) {
    if (c.dry() || c.st != WDM_OK) return c.st;
    int st;
    const bool tc = (p.a_split3 || (c.net->dt == DT_F32 && p.a_dtype == DT_BF16)) ? true : will_use_tc(c, p);
    if (!tc && c.net->dt == DT_BF16 && !(c.net->flags & (WDM_ENGINE_NO_TC | WDM_ENGINE_ALLOW_SIMT))) {
        // bf16 mode never drops to the CUDA-core kernel silently: a shape the tcgen05 kernel does not tile is an error
        c.fail(WDM_ERR_UNSUPPORTED);
        return c.st;
    }
    (tc ? c.net->n_tc : c.net->n_simt) += 1;
    wdm_unet::Span sp;
    if (c.net->profile) {
        cudaEventCreate(&sp.a);
        cudaEventCreate(&sp.b);
        sp.flops = p.a_split3 == 2 ? 0.0 : 2.0 * p.M * p.N * p.K / (p.a_split3 ? 6 : 1);  // tc32: six products per algorithmic MAC,
                                                                                           // counted on the (hi, hi) launch
        sp.tc = tc;
        sp.M = p.M, sp.N = p.N, sp.K = p.K, sp.taps = p.taps, sp.W = p.Wout;
        sp.tag = (p.tail_1x1 ? 1 : 0) | (p.ups == 2 ? 2 : 0) | (p.fuse_softmax ? 4 : 0) | (p.b_batch_stride ? 8 : 0) |
                 (p.residual ? 16 : 0) | (p.stride == 2 ? 32 : 0);
        if (tc) {
            const double es = 2.0, eo = p.out_dtype == DT_F32 ? 4.0 : 2.0;
            const double rows_in = p.a_shared ? (double)p.Hin * p.Win : (double)p.M / ((double)p.Hout * p.Wout) * p.Hin * p.Win;
            double b = rows_in * (p.C0 + (p.tail_1x1 ? 0 : p.C1)) * es;                   // main input tensor(s)
            if (p.tail_1x1) b += (double)p.M * (p.C1 + p.C2) * es;                        // shortcut inputs
            b += (double)p.N * p.K * es * (p.b_batch_stride ? (double)p.M / ((double)p.Hout * p.Wout) : (p.ups == 2 ? 4.0 : 1.0));
            b += (double)p.M * (p.out_nchw_valid ? p.out_nchw_valid : p.N) * eo;          // result
            if (p.residual) b += (double)p.M * p.N * eo;
            c.net->tc_bytes += b;
        }
        cudaEventRecord(sp.a, c.s);
    }
    st = tc ? launch_gemm_tc(p, c.s) : launch_gemm_simt(p, c.s);
    if (c.net->profile) {
        cudaEventRecord(sp.b, c.s);
        c.net->spans.push_back(sp);
    }
    c.fail(st);
    return st;
}

// conv over one or two (channel-concatenated) sources
Act conv_op(Ctx& c, const Act& a, const Act* a2, const ConvSpec& w, int stride, int ups, const float* temb_row,
            const Act* residual, bool want_stats = false) {
    const int dt = c.net->dt;
    int Hout = a.H, Wout = a.W;
    if (ups) Hout *= 2, Wout *= 2;
    if (stride == 2) Hout /= 2, Wout /= 2;
    Act o = new_act(c, Hout, Wout, w.Cout);
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.src0 = a.p, p.C0 = a.C, p.ld0 = a.C;
    if (a2) p.src1 = a2->p, p.C1 = a2->C, p.ld1 = a2->C;
    p.Hin = a.H, p.Win = a.W, p.Hout = Hout, p.Wout = Wout;
    p.taps = w.taps, p.stride = stride, p.pad = (w.taps == 9 && stride == 1) ? 1 : 0, p.ups = ups;
    p.B = w.pw, p.B3 = w.pw3, p.B3m = w.pw3m, p.ldb = w.taps * w.Cin_pad, p.b_layout = BL_NK;
    p.M = c.P * Hout * Wout, p.N = w.Cout, p.K = w.taps * (p.C0 + p.C1);
    p.alpha = 1.f, p.bias = w.pb;
    if (temb_row) p.temb = temb_row, p.temb_rows = c.T, p.temb_ld = c.net->model.temb_total;
    if (residual) p.residual = residual->p, p.ldr = residual->C;
    p.out = o.p, p.ldo = w.Cout;
    p.a_dtype = p.b_dtype = p.out_dtype = dt;
    if (p.C0 + p.C1 != w.Cin_pad) c.fail(WDM_ERR_BAD_SHAPE);
    if (want_stats && (p.M % 32) == 0 && (Hout * Wout) % 32 == 0 && will_use_tc(c, p)) {
        // tensor-core epilogue also emits the GroupNorm partial sums of what it stores (no separate stats pass)
        o.stats = reinterpret_cast<float*>(c.ar->alloc((size_t)(p.M / 32) * (p.N / 4) * 2 * sizeof(float)));
        if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
        p.stats_out = o.stats;
    }
    run_gemm(c, p);
    return o;
}

Act gn_op(Ctx& c, const Act& a, const Act* a2, const GnSpec& g, int silu) {
    const int C = a.C + (a2 ? a2->C : 0);
    Act o = new_act(c, a.H, a.W, C);
    if (C != g.C) c.fail(WDM_ERR_BAD_SHAPE);
    // Measured (round 2, P = 64, 30 calls): 5.05 ms per UNet call with the separate finalize launch, 5.19 ms with the
    // statistics reduced inside the normalise kernel -- under programmatic dependent launch the 4 us finalize kernels hide
    // behind the producer's tail, while the in-kernel reduction puts ~1.5 us in front of EVERY normalise CTA. Walking the
    // tensor back to front (WDM_GN_REVERSE) made no difference either way. Kept for the small-batch path experiments.
    static const int fused_finalize = []() {
        const char* e = getenv("WDM_GN_FUSED_FINALIZE");
        return e ? atoi(e) : 0;
    }();
    if (!c.dry() && c.st == WDM_OK && fused_finalize && c.net->dt == DT_BF16 && a.stats && (!a2 || a2->stats) && (C % 128) == 0 &&
        C / 4 <= 384 && ((a.H * a.W) % 32) == 0) {
        // statistics finalised inside the normalise kernel from the producers' side-cars: ONE launch per GroupNorm
        c.fail(launch_gn_apply_sidecar(a.p, a.C, a.stats, a2 ? a2->p : nullptr, a2 ? a2->C : 0, a2 ? a2->stats : nullptr, c.P,
                                       a.H * a.W, kGnEps, g.gamma, g.beta, silu, o.p, c.s));
        return o;
    }
    if (!c.dry() && c.st == WDM_OK) {
        if (a.stats && (!a2 || a2->stats))
            c.fail(launch_gn_finalize_sidecar(a.stats, a.C, a2 ? a2->stats : nullptr, a2 ? a2->C : 0, c.P, a.H * a.W,
                                              kGnEps, c.gn_scratch, c.s));
        else
            c.fail(launch_gn_stats(a.p, a.C, a2 ? a2->p : nullptr, a2 ? a2->C : 0, c.net->dt, c.P, a.H * a.W, kGnEps,
                                   c.gn_scratch, c.s));
        if (c.st == WDM_OK)
            c.fail(launch_gn_apply(a.p, a.C, a2 ? a2->p : nullptr, a2 ? a2->C : 0, c.net->dt, c.P, a.H * a.W,
                                   c.gn_scratch, g.gamma, g.beta, silu, o.p, c.s));
    }
    return o;
}

// models/unet.py:51-56 on the tensor-core path: nearest x2 upsample + 3x3 conv as four 2x2 phase convs on the source
// grid (2.25x fewer FLOPs, no materialised 4x tensor). Returns an empty Act (C == 0) when the shape is not supported.
Act upsample_conv_subpix(Ctx& c, const Act& a, const ConvSpec& w) {
    Act none;
    const bool t32 = c.net->dt == DT_F32 && w.pw_subpix3;  // tc32: the same decomposition with split operands
    if (!w.pw_subpix && !t32) return none;
    GemmParams p;
    memset(&p, 0, sizeof p);
    p.src0 = a.p, p.C0 = a.C, p.ld0 = a.C;
    p.Hin = a.H, p.Win = a.W, p.Hout = 2 * a.H, p.Wout = 2 * a.W;
    p.taps = 4, p.stride = 1, p.pad = 0, p.ups = 2;
    p.B = w.pw_subpix, p.B3 = w.pw_subpix3, p.B3m = w.pw_subpix3m, p.ldb = 4 * w.Cin, p.b_layout = BL_NK;
    // phase-major m-tiles of 128 source pixels: a source grid of 64 pixels pairs two patches per tile -> even count (see Ctx::Pa)
    const int Prun = (a.H * a.W < 128) ? c.Pa : c.P;
    p.M = Prun * p.Hout * p.Wout, p.N = w.Cout, p.K = 4 * a.C;
    p.alpha = 1.f, p.bias = w.pb;
    p.ldo = w.Cout;
    p.a_dtype = p.b_dtype = p.out_dtype = c.net->dt;
    p.out = reinterpret_cast<void*>(16);  // placeholder for the support check (alignment only)
    if (t32) {
        GemmParams q;
        if (!tc32_plan(c, p, &q)) return none;
        Act o = new_act(c, p.Hout, p.Wout, w.Cout);
        p.out = o.p;
        run_gemm(c, p);
        return o;
    }
    if (!will_use_tc(c, p)) return none;
    Act o = new_act(c, p.Hout, p.Wout, w.Cout);
    p.out = o.p;
    o.stats = reinterpret_cast<float*>(c.ar->alloc((size_t)(p.M / 32) * (p.N / 4) * 2 * sizeof(float)));
    if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
    p.stats_out = o.stats;
    run_gemm(c, p);
    return o;
}

// models/unet.py:119-138. x2 != null: the block input is cat([x, x2], dim=1) (unet.py:379-380).
Act resblock_op(Ctx& c, const Act& x, const Act* x2, const ResSpec& r) {
    Act n1 = gn_op(c, x, x2, r.norm1, 1);
    Act h1 = conv_op(c, n1, nullptr, r.conv1, 1, 0, c.temb + r.temb_off, nullptr, true);
    free_act(c, n1);
    Act n2 = gn_op(c, h1, nullptr, r.norm2, 1);
    free_act(c, h1);
    Act out;
    if (r.has_nin && r.conv2f.pw) {
        // tensor-core path: x + conv2(h), x = nin(cat[x, x2]), as ONE contraction over K = 9*Cout + Cin
        GemmParams p;
        memset(&p, 0, sizeof p);
        p.src0 = n2.p, p.C0 = n2.C, p.ld0 = n2.C;
        p.src1 = x.p, p.C1 = x.C, p.ld1 = x.C;
        if (x2) p.src2 = x2->p, p.C2 = x2->C, p.ld2 = x2->C;
        p.tail_1x1 = 1;
        p.Hin = p.Hout = n2.H, p.Win = p.Wout = n2.W;
        p.taps = 9, p.stride = 1, p.pad = 1;
        p.B = r.conv2f.pw, p.ldb = 9 * r.Cout + r.Cin, p.b_layout = BL_NK;
        p.M = c.P * n2.H * n2.W, p.N = r.Cout, p.K = 9 * r.Cout + r.Cin;
        p.alpha = 1.f, p.bias = r.conv2f.pb;
        p.ldo = r.Cout;
        p.a_dtype = p.b_dtype = p.out_dtype = c.net->dt;
        p.out = reinterpret_cast<void*>(16);
        if (will_use_tc(c, p)) {
            out = new_act(c, n2.H, n2.W, r.Cout);
            p.out = out.p;
            if ((p.M % 32) == 0 && (n2.H * n2.W) % 32 == 0) {
                out.stats = reinterpret_cast<float*>(c.ar->alloc((size_t)(p.M / 32) * (p.N / 4) * 2 * sizeof(float)));
                if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
                p.stats_out = out.stats;
            }
            run_gemm(c, p);
            free_act(c, n2);
            return out;
        }
    }
    if (r.has_nin) {
        Act sc = conv_op(c, x, x2, r.nin, 1, 0, nullptr, nullptr);
        out = conv_op(c, n2, nullptr, r.conv2, 1, 0, nullptr, &sc, true);
        free_act(c, sc);
    } else {
        if (x2) c.fail(WDM_ERR_BAD_SHAPE);
        out = conv_op(c, n2, nullptr, r.conv2, 1, 0, nullptr, &x, true);
    }
    free_act(c, n2);
    return out;
}

// Tensor-core attention (L % 128 == 0): every contraction on the tcgen05 kernel with K-major operands only.
//   qk  = h Wqk^T + b            [P*L][2C]
//   vT  = Wv h^T                 [P][C][L]   (A = Wv shared by all patches, B = h per patch)  -- V transposed so that
//   S   = (q k^T) C^-1/2         [P*L][L]     P.V has a K-major right operand; b_v is added after P.V (rows of
//   Pm  = softmax(S)                          the softmax sum to one)
//   O   = Pm vT^T + b_v          [P*L][C]
//   out = O Wproj^T + b + x
Act attn_op_tc(Ctx& c, const Act& x, const AttnSpec& a, int G) {
    // G patches share one 128-row tile when a patch has fewer than 128 tokens (the 8x8 mid block: G = 2): the matmuls run
    // on groups of G*L rows and the softmax is block-diagonal (cross-patch scores are masked to zero probability).
    const int C = a.C, L = x.H * x.W, Lg = G * L, Hg = G * x.H;
    const int P = G > 1 ? c.Pa : c.P;  // G = 2 pairs patches: odd counts run with the slack patch (see Ctx::Pa)
    const size_t es = 2;
    static const int fused_enabled = []() {
        const char* e = getenv("WDM_ATTN_FUSED");
        return e ? atoi(e) : 1;
    }();
    const bool fused = fused_enabled && a.gw && a.wpv;  // the four-contraction form (see AttnSpec)
    Act n = gn_op(c, x, nullptr, a.norm, 0);
    if (P != c.P && !c.dry() && c.st == WDM_OK) {
        // the slack patch of the normalised input feeds the keys / values of its tile partner's GEMMs as B-operand columns
        // that the block-diagonal softmax weights with exactly 0: they must be finite (0 * NaN would poison the real patch)
        cudaError_t e = cudaMemsetAsync((char*)n.p + (size_t)c.P * L * C * es, 0, (size_t)(P - c.P) * L * C * es, c.s);
        if (e != cudaSuccess) c.fail(wdm_cuda_error((int)e));
    }
    GemmParams p;
    // fused: g = n G^T + u  [P*L][C];  unfused: qk = n Wqk^T + b  [P*L][2C]
    const int ldq = fused ? C : 2 * C;
    Act qk = new_act(c, x.H, x.W, ldq);
    memset(&p, 0, sizeof p);
    p.src0 = n.p, p.C0 = C, p.ld0 = C, p.Hin = p.Hout = x.H, p.Win = p.Wout = x.W, p.taps = 1, p.stride = 1;
    p.B = fused ? a.gw : a.qkv.pw, p.ldb = C, p.b_layout = BL_NK, p.M = P * L, p.N = ldq, p.K = C, p.alpha = 1.f;
    p.bias = fused ? a.gb : a.qkv.pb;
    p.out = qk.p, p.ldo = ldq, p.a_dtype = p.b_dtype = p.out_dtype = DT_BF16;
    run_gemm(c, p);
    // vT = Wv n^T (fused: w^T = (Wp Wv) n^T)   [P][C][L]: A = weights shared by all patches, B = n per patch
    void* vT = c.ar->alloc((size_t)P * C * L * es);
    if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
    memset(&p, 0, sizeof p);
    p.src0 = fused ? a.wpv : (void*)((char*)a.qkv.pw + (size_t)2 * C * C * es), p.C0 = C, p.ld0 = C, p.a_shared = 1;
    p.Hin = p.Hout = C / 128, p.Win = p.Wout = 128, p.taps = 1, p.stride = 1;
    p.B = n.p, p.ldb = C, p.b_batch_stride = (long long)Lg * C, p.b_layout = BL_NK;
    p.M = (P / G) * C, p.N = Lg, p.K = C, p.alpha = 1.f;
    p.out = vT, p.ldo = Lg, p.a_dtype = p.b_dtype = p.out_dtype = DT_BF16;
    run_gemm(c, p);
    // S -> softmax fused in the score GEMM's epilogue (probabilities straight to bf16; the fp32 scores stay in TMEM).
    // fused: keys = n itself (K-major rows of the normalised input), unfused: the k half of qk
    void* Pm = c.ar->alloc((size_t)P * L * Lg * es);
    if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
    memset(&p, 0, sizeof p);
    p.src0 = qk.p, p.C0 = C, p.ld0 = ldq, p.Hin = p.Hout = Hg, p.Win = p.Wout = x.W, p.taps = 1, p.stride = 1;
    if (fused)
        p.B = n.p, p.b_batch_stride = (long long)Lg * C, p.ldb = C;
    else
        p.B = (char*)qk.p + (size_t)C * es, p.b_batch_stride = (long long)Lg * 2 * C, p.ldb = 2 * C;
    p.b_layout = BL_NK;
    p.M = P * L, p.N = Lg, p.K = C, p.alpha = (float)(1.0 / sqrt((double)C));
    p.out = Pm, p.ldo = Lg, p.a_dtype = p.b_dtype = DT_BF16, p.out_dtype = DT_BF16;
    p.fuse_softmax = 1, p.softmax_seg = G > 1 ? L : 0;
    float* S = nullptr;
    // deferred normalisation: the score epilogue stores exp(s - max) and 1/sum per row, the P.V epilogue applies it
    float* rscale = nullptr;
    if (will_use_tc(c, p)) {
        rscale = reinterpret_cast<float*>(c.ar->alloc((size_t)P * L * sizeof(float)));
        if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
        p.row_scale_out = rscale ? rscale : reinterpret_cast<float*>(16);
        run_gemm(c, p);
    } else {
        p.fuse_softmax = 0, p.softmax_seg = 0, p.out_dtype = DT_F32;
        S = reinterpret_cast<float*>(c.ar->alloc((size_t)P * L * Lg * 4));
        if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
        p.out = S;
        run_gemm(c, p);
        if (!c.dry() && c.st == WDM_OK) c.fail(launch_softmax_rows(S, P * L, Lg, Pm, DT_BF16, c.s, G > 1 ? L : 0));
    }
    free_act(c, qk);
    free_act(c, n);
    // O = Pm vT^T + b_v  (fused: the block output  Pm w + bo + x  with the GroupNorm side-car of the next block)
    Act O = new_act(c, x.H, x.W, C);
    memset(&p, 0, sizeof p);
    p.src0 = Pm, p.C0 = Lg, p.ld0 = Lg, p.Hin = p.Hout = Hg, p.Win = p.Wout = x.W, p.taps = 1, p.stride = 1;
    p.B = vT, p.b_batch_stride = (long long)C * Lg, p.ldb = Lg, p.b_layout = BL_NK;
    p.M = P * L, p.N = C, p.K = Lg, p.alpha = 1.f, p.bias = fused ? a.bo : a.qkv.pb + 2 * C;
    p.out = O.p, p.ldo = C, p.a_dtype = p.b_dtype = p.out_dtype = DT_BF16;
    if (rscale) p.row_scale = rscale;
    if (fused) {
        p.residual = x.p, p.ldr = x.C;
        if ((p.M % 32) == 0 && (Hg * x.W) % 32 == 0) {
            O.stats = reinterpret_cast<float*>(c.ar->alloc((size_t)(p.M / 32) * (p.N / 4) * 2 * sizeof(float)));
            if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
            p.stats_out = O.stats;
        }
    }
    run_gemm(c, p);
    if (S) c.ar->free(S);
    if (rscale) c.ar->free(rscale);
    c.ar->free(Pm);
    c.ar->free(vT);
    if (fused) return O;
    Act out = conv_op(c, O, nullptr, a.proj, 1, 0, nullptr, &x, true);
    free_act(c, O);
    return out;
}

// models/unet.py:168-193
Act attn_op(Ctx& c, const Act& x, const AttnSpec& a) {
    const int dt = c.net->dt;
    const int C = a.C, L = x.H * x.W, P = c.P;
    if (dt == DT_BF16 && !(c.net->flags & WDM_ENGINE_NO_TC) && (C % 128) == 0) {
        if ((L % 128) == 0) return attn_op_tc(c, x, a, 1);
        if (L == 64) return attn_op_tc(c, x, a, 2);
    }
    Act n = gn_op(c, x, nullptr, a.norm, 0);
    Act qkv = conv_op(c, n, nullptr, a.qkv, 1, 0, nullptr, nullptr);  // [P*L][3C]
    free_act(c, n);
    float* S = reinterpret_cast<float*>(c.ar->alloc((size_t)P * L * L * 4));
    void* Pm = c.ar->alloc((size_t)P * L * L * dtype_size(dt));
    if (c.ar->failed) c.fail(WDM_ERR_WORKSPACE);
    GemmParams p;
    memset(&p, 0, sizeof p);
    // S[b] = (q k^T) * C^-0.5
    p.src0 = qkv.p, p.C0 = C, p.ld0 = 3 * C, p.Hin = x.H, p.Win = x.W, p.Hout = x.H, p.Wout = x.W;
    p.taps = 1, p.stride = 1;
    p.B = (char*)qkv.p + (size_t)C * dtype_size(dt), p.b_batch_stride = (long long)L * 3 * C, p.ldb = 3 * C;
    p.b_layout = BL_NK;
    p.M = P * L, p.N = L, p.K = C;
    p.alpha = (float)(1.0 / sqrt((double)C));
    p.out = S, p.ldo = L;
    p.a_dtype = p.b_dtype = dt, p.out_dtype = DT_F32;
    run_gemm(c, p);
    if (!c.dry() && c.st == WDM_OK) c.fail(launch_softmax_rows(S, P * L, L, Pm, dt, c.s));
    // O[b] = P[b] V[b]
    Act O = new_act(c, x.H, x.W, C);
    memset(&p, 0, sizeof p);
    p.src0 = Pm, p.C0 = L, p.ld0 = L, p.Hin = x.H, p.Win = x.W, p.Hout = x.H, p.Wout = x.W;
    p.taps = 1, p.stride = 1;
    p.B = (char*)qkv.p + (size_t)2 * C * dtype_size(dt), p.b_batch_stride = (long long)L * 3 * C, p.ldb = 3 * C;
    p.b_layout = BL_KN;
    p.M = P * L, p.N = C, p.K = L;
    p.alpha = 1.f;
    p.out = O.p, p.ldo = C;
    p.a_dtype = p.b_dtype = p.out_dtype = dt;
    run_gemm(c, p);
    c.ar->free(S);
    c.ar->free(Pm);
    free_act(c, qkv);
    Act out = conv_op(c, O, nullptr, a.proj, 1, 0, nullptr, &x, true);
    free_act(c, O);
    return out;
}

// models/unet.py:346-395
int forward_impl(wdm_unet* net, Arena* ar, const void* x, const float* t, int T, int P, float* eps_out,
                 cudaStream_t s) {
    Model& m = net->model;
    Ctx c;
    c.net = net, c.ar = ar, c.s = s, c.P = P, c.T = T;
    c.Pa = P + (P & 1);
    const int R = m.cfg.resolution, L = m.cfg.n_levels, nrb = m.cfg.num_res_blocks, tc = 4 * m.cfg.ch;
    c.temb = reinterpret_cast<float*>(ar->alloc((size_t)T * m.temb_total * 4));
    float* temb_scratch = reinterpret_cast<float*>(ar->alloc((size_t)T * 2 * tc * 4));
    c.gn_scratch = reinterpret_cast<float*>(ar->alloc(gn_stats_bytes(P)));
    if (ar->failed) return WDM_ERR_WORKSPACE;
    if (!c.dry()) {
        TembParams tp;
        tp.t = t, tp.freqs = net->freqs, tp.T = T, tp.ch = m.cfg.ch;
        tp.w0 = net->w0, tp.b0 = net->b0, tp.w1 = net->w1, tp.b1 = net->b1, tp.wp = net->wp, tp.bp = net->bp;
        tp.total = m.temb_total, tp.scratch = temb_scratch, tp.out = c.temb;
        c.fail(launch_temb(tp, s));
    }
    Act xin;
    xin.p = const_cast<void*>(x), xin.H = R, xin.W = R, xin.C = net->cin_pad;
    std::vector<Act> hs;
    hs.push_back(conv_op(c, xin, nullptr, m.conv_in, 1, 0, nullptr, nullptr, true));
    for (int lv = 0; lv < L; ++lv) {
        for (int ib = 0; ib < nrb; ++ib) {
            Act h = resblock_op(c, hs.back(), nullptr, m.down[lv].blocks[ib]);
            if (!m.down[lv].attns.empty()) {
                Act h2 = attn_op(c, h, m.down[lv].attns[ib]);
                free_act(c, h);
                h = h2;
            }
            hs.push_back(h);
        }
        if (m.down[lv].has_resample) hs.push_back(conv_op(c, hs.back(), nullptr, m.down[lv].resample, 2, 0, nullptr, nullptr, true));
    }
    Act h = resblock_op(c, hs.back(), nullptr, m.mid1);
    {
        Act h2 = attn_op(c, h, m.mid_attn);
        free_act(c, h);
        h = resblock_op(c, h2, nullptr, m.mid2);
        free_act(c, h2);
    }
    const bool fold_ups = (net->dt == DT_F32) || (net->flags & WDM_ENGINE_NO_TC);
    for (int lv = L - 1; lv >= 0; --lv) {
        for (int ib = 0; ib < nrb + 1; ++ib) {
            Act skip = hs.back();
            hs.pop_back();
            Act h2 = resblock_op(c, h, &skip, m.up[lv].blocks[ib]);
            free_act(c, h);
            free_act(c, skip);
            h = h2;
            if (!m.up[lv].attns.empty()) {
                Act h3 = attn_op(c, h, m.up[lv].attns[ib]);
                free_act(c, h);
                h = h3;
            }
        }
        if (m.up[lv].has_resample) {
            Act h2;
            if (Act sp = upsample_conv_subpix(c, h, m.up[lv].resample); sp.p || (c.dry() && sp.C)) {
                h2 = sp;  // bf16 tensor-core mode, or fp32 with WDM_ENGINE_TC32
            } else if (fold_ups) {
                h2 = conv_op(c, h, nullptr, m.up[lv].resample, 1, 1, nullptr, nullptr, true);
            } else {
                Act u = new_act(c, h.H * 2, h.W * 2, h.C);
                if (!c.dry() && c.st == WDM_OK) c.fail(launch_upsample2x(h.p, net->dt, P, h.H, h.W, h.C, u.p, s));
                h2 = conv_op(c, u, nullptr, m.up[lv].resample, 1, 0, nullptr, nullptr, true);
                free_act(c, u);
            }
            free_act(c, h);
            h = h2;
        }
    }
    Act n = gn_op(c, h, nullptr, m.norm_out, 1);
    free_act(c, h);
    bool out_done = false;
    if (m.cfg.wavelet_in_unet || m.conv_out.Cout > 4) {
        // conv_out -> [P*R*R][ld] fp32 (NHWC rows, 48 valid columns) -> IWT -> eps_out [P, 3, 4R, 4R]  (unet.py:393-394)
        const bool tc_out = net->dt == DT_BF16 && m.conv_out.pw && m.conv_out.pb_pad;  // bf16: Cout zero-padded to 64
        const int ld = tc_out ? 64 : (m.conv_out.Cout + 7) / 8 * 8;
        float* tmp = reinterpret_cast<float*>(ar->alloc((size_t)P * R * R * ld * sizeof(float)));
        if (ar->failed) c.fail(WDM_ERR_WORKSPACE);
        GemmParams p;
        memset(&p, 0, sizeof p);
        p.src0 = n.p, p.C0 = n.C, p.ld0 = n.C, p.Hin = p.Hout = R, p.Win = p.Wout = R;
        p.taps = 9, p.stride = 1, p.pad = 1;
        p.b_layout = BL_NK, p.ldb = 9 * n.C;
        p.M = P * R * R, p.K = 9 * n.C, p.alpha = 1.f;
        p.out = tmp ? (void*)tmp : reinterpret_cast<void*>(16), p.ldo = ld, p.out_dtype = DT_F32;
        if (tc_out) {
            p.B = m.conv_out.pw, p.N = 64, p.bias = m.conv_out.pb_pad, p.a_dtype = p.b_dtype = DT_BF16;
        } else {
            p.B = m.conv_out.pw32, p.N = ld, p.bias = m.conv_out.pb, p.a_dtype = net->dt, p.b_dtype = DT_F32;
        }
        run_gemm(c, p);
        if (!c.dry() && c.st == WDM_OK)
            c.fail(m.cfg.wavelet_in_unet ? launch_iwt_nhwc(tmp, ld, P, R, eps_out, s)
                                         : launch_rows_to_nchw(tmp, ld, P, R * R, m.conv_out.Cout, eps_out, s));
        ar->free(tmp);
        out_done = true;
    }
    if (!out_done && m.conv_out.pw && m.conv_out.pb_pad) {
        GemmParams p;
        memset(&p, 0, sizeof p);
        p.src0 = n.p, p.C0 = n.C, p.ld0 = n.C, p.Hin = p.Hout = R, p.Win = p.Wout = R;
        p.taps = 9, p.stride = 1, p.pad = 1;
        p.B = m.conv_out.pw, p.ldb = 9 * n.C, p.b_layout = BL_NK;
        p.M = P * R * R, p.N = 64, p.K = 9 * n.C, p.alpha = 1.f, p.bias = m.conv_out.pb_pad;
        p.out = eps_out ? (void*)eps_out : reinterpret_cast<void*>(16), p.ldo = 64, p.out_nchw_valid = m.conv_out.Cout;
        p.a_dtype = p.b_dtype = DT_BF16, p.out_dtype = DT_F32;
        if (will_use_tc(c, p)) {
            run_gemm(c, p);
            out_done = true;
        }
    }
    if (!out_done && !c.dry() && c.st == WDM_OK)
        c.fail(launch_conv_small_cout(n.p, net->dt, P, R, R, n.C, m.conv_out.pw32, m.conv_out.pb, m.conv_out.Cout,
                                      eps_out, s));
    free_act(c, n);
    if (!hs.empty()) c.fail(WDM_ERR_BAD_ARG);  // schedule bug: every skip must be consumed
    return c.st;
}

}  // namespace
}  // namespace wdm

// ================================================================================================ C ABI
extern "C" int wdm_unet_param_count(const wdm_unet_config* cfg) {
    if (!cfg) return WDM_ERR_BAD_ARG;
    Model m;
    int st = build_model(*cfg, &m);
    return st != WDM_OK ? st : (int)m.params.size();
}

extern "C" int wdm_unet_param_info(const wdm_unet_config* cfg, int i, char* name, int cap, long long* numel) {
    if (!cfg) return WDM_ERR_BAD_ARG;
    Model m;
    int st = build_model(*cfg, &m);
    if (st != WDM_OK) return st;
    if (i < 0 || i >= (int)m.params.size()) return WDM_ERR_BAD_ARG;
    if (name && cap > 0) {
        strncpy(name, m.params[i].name.c_str(), cap - 1);
        name[cap - 1] = 0;
    }
    if (numel) *numel = m.params[i].numel;
    return WDM_OK;
}

extern "C" size_t wdm_unet_packed_bytes(const wdm_unet_config* cfg, int precision) {
    if (!cfg || (precision != WDM_PREC_FP32 && precision != WDM_PREC_BF16)) return 0;
    wdm_unet net;
    if (build_model(*cfg, &net.model) != WDM_OK) return 0;
    net.dt = precision == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
    net.flags = precision == WDM_PREC_FP32 ? WDM_ENGINE_TC32 : 0;  // upper bound over the engine flags: tc32 adds the split weights
    size_t total = 0;
    pack_model(&net, nullptr, 0, &total);
    return total;
}

extern "C" size_t wdm_unet_packed_bytes_flags(const wdm_unet_config* cfg, int precision, int flags) {
    if (!cfg || (precision != WDM_PREC_FP32 && precision != WDM_PREC_BF16)) return 0;
    wdm_unet net;
    if (build_model(*cfg, &net.model) != WDM_OK) return 0;
    net.dt = precision == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
    net.flags = flags;
    size_t total = 0;
    pack_model(&net, nullptr, 0, &total);
    return total;
}

extern "C" int wdm_unet_create(const wdm_unet_config* cfg, int precision, int flags, const float* flat_params,
                               long long flat_numel, void* packed, size_t packed_bytes, void* stream,
                               wdm_unet_t** out) {
    if (!cfg || !flat_params || !packed || !out) return WDM_ERR_BAD_ARG;
    if (precision != WDM_PREC_FP32 && precision != WDM_PREC_BF16) return WDM_ERR_BAD_ARG;
    if (!wdm_aligned(packed, 256) || !wdm_aligned(flat_params, 16)) return WDM_ERR_BAD_ALIGN;
    wdm_unet* net = new wdm_unet();
    int st = build_model(*cfg, &net->model);
    if (st != WDM_OK) {
        delete net;
        return st;
    }
    const ParamRef& last = net->model.params.back();
    if (flat_numel != last.off + last.numel) {
        delete net;
        return WDM_ERR_BAD_ARG;
    }
    net->dt = precision == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
    net->flags = flags;
    net->packed = reinterpret_cast<char*>(packed);
    net->packed_bytes = packed_bytes;
    size_t need = 0;
    {
        wdm_unet probe;
        probe.model = net->model;
        probe.dt = net->dt;
        probe.flags = flags;
        pack_model(&probe, nullptr, 0, &need);
    }
    if (packed_bytes < need) {
        delete net;
        return WDM_ERR_WORKSPACE;
    }
    size_t total = 0;
    st = pack_model(net, flat_params, static_cast<cudaStream_t>(stream), &total);
    if (st != WDM_OK) {
        delete net;
        return st;
    }
    *out = net;
    return WDM_OK;
}

extern "C" void wdm_unet_destroy(wdm_unet_t* net) { delete net; }

extern "C" int wdm_unet_input_channels_padded(const wdm_unet_t* net) { return net ? net->cin_pad : WDM_ERR_BAD_ARG; }

extern "C" size_t wdm_unet_workspace_bytes(const wdm_unet_t* net, int P) {
    if (!net || P <= 0) return 0;
    Arena ar;
    ar.dry = true;
    forward_impl(const_cast<wdm_unet*>(net), &ar, nullptr, nullptr, 1, P, nullptr, 0);
    // T == P needs more temb rows than T == 1: account for it
    Arena ar2;
    ar2.dry = true;
    forward_impl(const_cast<wdm_unet*>(net), &ar2, nullptr, nullptr, P, P, nullptr, 0);
    return (ar.peak > ar2.peak ? ar.peak : ar2.peak);
}

extern "C" int wdm_unet_forward(wdm_unet_t* net, const void* x, const float* t, int T, int P, float* eps_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (!net || !x || !t || !eps_out || !workspace) return WDM_ERR_BAD_ARG;
    if (P <= 0 || (T != 1 && T != P)) return WDM_ERR_BAD_ARG;
    if (!wdm_aligned(workspace, 1024) || !wdm_aligned(x, 16) || !wdm_aligned(eps_out, 16)) return WDM_ERR_BAD_ALIGN;
    Arena ar;
    ar.base = reinterpret_cast<char*>(workspace);
    ar.cap = workspace_bytes;
    return forward_impl(net, &ar, x, t, T, P, eps_out, static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_unet_profile_enable(wdm_unet_t* net, int on) {
    if (!net) return WDM_ERR_BAD_ARG;
    net->profile = on != 0;
    return WDM_OK;
}

extern "C" int wdm_unet_profile_read(wdm_unet_t* net, double* tc_ms, double* tc_flops, long long* tc_launches,
                                     double* simt_ms, double* simt_flops, long long* simt_launches) {
    if (!net) return WDM_ERR_BAD_ARG;
    double ms[2] = {0, 0}, fl[2] = {0, 0};
    long long n[2] = {0, 0};
    // WDM_PROFILE_DUMP=<path>: one CSV row per profiled launch (tools/profile_unet.py --spans)
    FILE* dump = nullptr;
    if (const char* path = getenv("WDM_PROFILE_DUMP")) dump = fopen(path, "a");
    for (auto& s : net->spans) {
        cudaError_t e = cudaEventSynchronize(s.b);
        if (e != cudaSuccess) return wdm_cuda_error((int)e);
        float t = 0.f;
        cudaEventElapsedTime(&t, s.a, s.b);
        const int k = s.tc ? 0 : 1;
        ms[k] += t, fl[k] += s.flops, n[k] += 1;
        if (dump) fprintf(dump, "%d,%d,%d,%d,%d,%d,%d,%.3f,%.6e\n", s.tc ? 1 : 0, s.M, s.N, s.K, s.taps, s.W, s.tag, t * 1e3, s.flops);
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    net->spans.clear();
    if (dump) fclose(dump);
    if (tc_ms) *tc_ms = ms[0];
    if (tc_flops) *tc_flops = fl[0];
    if (tc_launches) *tc_launches = n[0];
    if (simt_ms) *simt_ms = ms[1];
    if (simt_flops) *simt_flops = fl[1];
    if (simt_launches) *simt_launches = n[1];
    return WDM_OK;
}

extern "C" int wdm_unet_counters(const wdm_unet_t* net, long long* tc_launches, long long* simt_launches) {
    if (!net) return WDM_ERR_BAD_ARG;
    if (tc_launches) *tc_launches = net->n_tc;
    if (simt_launches) *simt_launches = net->n_simt;
    return WDM_OK;
}

extern "C" double wdm_unet_profile_tc_bytes(wdm_unet_t* net) {
    if (!net) return 0.0;
    const double b = net->tc_bytes;
    net->tc_bytes = 0;
    return b;
}

extern "C" int wdm_gather_patches(const float* src0, int C0, const float* src1, int C1, const float* src2, int C2,
                                  int B, int h, int w, const int* patches, int P, int R, int Cpad, void* out,
                                  int out_dtype, void* stream) {
    if (!src0 || !patches || !out || C0 <= 0) return WDM_ERR_BAD_ARG;
    if (out_dtype != WDM_PREC_FP32 && out_dtype != WDM_PREC_BF16) return WDM_ERR_BAD_ARG;
    GatherParams g;
    memset(&g, 0, sizeof g);
    g.src[0] = src0, g.Cs[0] = C0, g.nsrc = 1;
    if (src1 && C1 > 0) g.src[1] = src1, g.Cs[1] = C1, g.nsrc = 2;
    if (src2 && C2 > 0) {
        if (g.nsrc != 2) return WDM_ERR_BAD_ARG;
        g.src[2] = src2, g.Cs[2] = C2, g.nsrc = 3;
    }
    if (C0 + g.Cs[1] + g.Cs[2] > Cpad || R > h || R > w) return WDM_ERR_BAD_SHAPE;
    g.B = B, g.h = h, g.w = w, g.patches = patches, g.P = P, g.R = R, g.Cpad = Cpad, g.out = out;
    g.out_dtype = out_dtype == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
    return launch_gather_patches(g, static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_gather_patches_update(const float* src, int C, int c_off, int B, int h, int w, const int* patches,
                                         int P, int R, int Cpad, void* out, int out_dtype, void* stream) {
    (void)B;
    if (!src || !patches || !out || C <= 0 || c_off < 0) return WDM_ERR_BAD_ARG;
    if (out_dtype != WDM_PREC_FP32 && out_dtype != WDM_PREC_BF16) return WDM_ERR_BAD_ARG;
    if (c_off + C > Cpad || R > h || R > w) return WDM_ERR_BAD_SHAPE;
    return launch_gather_update(src, C, c_off, h, w, patches, P, R, Cpad, out,
                                out_dtype == WDM_PREC_FP32 ? DT_F32 : DT_BF16, static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_ddim_step(const float* eps, const int* patches, const int* img_first, int P, int B, int Cp, int R,
                             int h, int w, const float* xt, float* x0_out, float* xt_next, float at, float at_next,
                             void* stream) {
    if (!eps || !patches || !img_first || !xt || !x0_out || !xt_next) return WDM_ERR_BAD_ARG;
    if (at <= 0.f || at > 1.f || at_next <= 0.f || at_next > 1.f) return WDM_ERR_BAD_ARG;
    DdimParams d;
    d.eps = eps, d.patches = patches, d.img_first = img_first, d.P = P, d.B = B, d.Cp = Cp, d.R = R, d.h = h, d.w = w;
    d.xt = xt, d.x0_out = x0_out, d.xt_next = xt_next, d.at = at, d.at_next = at_next;
    return launch_ddim_step(d, static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_gemm(const wdm_gemm_params* p, int impl, void* stream) {
    if (!p || !p->src0 || !p->B || !p->out) return WDM_ERR_BAD_ARG;
    if (impl == WDM_GEMM_IMPL_SIMT) return launch_gemm_simt(*p, static_cast<cudaStream_t>(stream));
    if (impl == WDM_GEMM_IMPL_TC) {
        if (!gemm_tc_supported(*p)) return WDM_ERR_UNSUPPORTED;
        return launch_gemm_tc(*p, static_cast<cudaStream_t>(stream));
    }
    return WDM_ERR_BAD_ARG;
}

extern "C" int wdm_gemm_ksplit_plan(const wdm_gemm_params* p) { return p ? gemm_tc_ksplit_plan(*p) : 1; }

extern "C" size_t wdm_groupnorm_scratch_bytes(int P) { return gn_stats_bytes(P); }

extern "C" int wdm_groupnorm_silu(const void* src0, int C0, const void* src1, int C1, int dtype, int P, int HW,
                                  float eps, const float* gamma, const float* beta, int silu, void* out,
                                  void* scratch, void* stream) {
    if (!src0 || !gamma || !beta || !out || !scratch) return WDM_ERR_BAD_ARG;
    if (dtype != WDM_PREC_FP32 && dtype != WDM_PREC_BF16) return WDM_ERR_BAD_ARG;
    const int dt = dtype == WDM_PREC_FP32 ? DT_F32 : DT_BF16;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int st = launch_gn_stats(src0, C0, src1, C1, dt, P, HW, eps, reinterpret_cast<float*>(scratch), s);
    if (st != WDM_OK) return st;
    return launch_gn_apply(src0, C0, src1, C1, dt, P, HW, reinterpret_cast<const float*>(scratch), gamma, beta, silu,
                           out, s);
}

extern "C" int wdm_groupnorm_silu_sidecar(const void* src0, int C0, const float* sc0, const void* src1, int C1, const float* sc1,
                                          int P, int HW, float eps, const float* gamma, const float* beta, int silu, void* out,
                                          void* stream) {
    if (!src0 || !sc0 || !gamma || !beta || !out || (C1 && (!src1 || !sc1))) return WDM_ERR_BAD_ARG;
    return launch_gn_apply_sidecar(src0, C0, sc0, src1, C1, sc1, P, HW, eps, gamma, beta, silu, out,
                                   static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_softmax_rows(const float* S, int rows, int L, void* out, int out_dtype, void* stream) {
    if (!S || !out) return WDM_ERR_BAD_ARG;
    return launch_softmax_rows(S, rows, L, out, out_dtype == WDM_PREC_FP32 ? DT_F32 : DT_BF16,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_split3_act(const float* src0, int C0, const float* src1, int C1, long long rows, void* out, void* stream) {
    return launch_split3_act(src0, C0, src1, C1, rows, out, static_cast<cudaStream_t>(stream));
}

extern "C" int wdm_split3_weight(const float* w, int N, int taps, int C, void* out, void* out_main, void* stream) {
    return launch_split3_weight(w, N, taps, C, out, out_main, static_cast<cudaStream_t>(stream));
}
