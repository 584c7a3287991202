"""GPU parity of the UNet engine and its kernels against the torch-fp32 oracle (oracle/unet_oracle.py, itself
pinned to the reference module by tests/golden/*) -- through the C ABI."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from oracle import unet_oracle as O
from wavedm_b200 import _lib, engine

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_cfg():
    return O.default_config(data__image_size=16, model__ch=128, model__ch_mult=[1, 2], model__num_res_blocks=1,
                            model__attn_resolutions=[8])


class GemmParams(ctypes.Structure):
    _fields_ = [("src0", ctypes.c_void_p), ("src1", ctypes.c_void_p)] + \
        [(n, ctypes.c_int) for n in ("C0", "C1", "ld0", "ld1", "Hin", "Win", "Hout", "Wout", "taps", "stride", "pad", "ups")] + \
        [("B", ctypes.c_void_p), ("b_batch_stride", ctypes.c_longlong)] + \
        [(n, ctypes.c_int) for n in ("ldb", "b_layout", "M", "N", "K")] + \
        [("alpha", ctypes.c_float), ("bias", ctypes.c_void_p), ("temb", ctypes.c_void_p), ("temb_rows", ctypes.c_int),
         ("temb_ld", ctypes.c_int), ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int), ("out", ctypes.c_void_p)] + \
        [(n, ctypes.c_int) for n in ("ldo", "a_dtype", "b_dtype", "out_dtype")] + [("stats_out", ctypes.c_void_p), ("a_shared", ctypes.c_int), ("tail_1x1", ctypes.c_int),
                                                                                ("src2", ctypes.c_void_p), ("C2", ctypes.c_int), ("ld2", ctypes.c_int),
                                                                                ("fuse_softmax", ctypes.c_int), ("softmax_seg", ctypes.c_int), ("out_nchw_valid", ctypes.c_int),
                                                                                ("row_scale_out", ctypes.c_void_p), ("row_scale", ctypes.c_void_p),
                                                                                ("a_split3", ctypes.c_int), ("B3", ctypes.c_void_p), ("B3m", ctypes.c_void_p),
                                                                                ("ksplit", ctypes.c_int), ("ksplit_scratch", ctypes.c_void_p)]


def run_conv(x, x2, w, bias, stride, ups, temb, residual, dtype, impl, want_stats=False, ksplit=None):
    """x, x2: NCHW fp32 torch (cpu); w OIHW. Runs the C-ABI gemm on NHWC tensors, returns NCHW fp32."""
    lib = _lib.load()
    td = torch.float32 if dtype == 0 else torch.bfloat16
    P, C0, H, W = x.shape
    C1 = 0 if x2 is None else x2.shape[1]
    Cout, Cin, k, _ = w.shape
    taps = k * k
    Hout, Wout = (H * 2, W * 2) if ups else ((H // 2, W // 2) if stride == 2 else (H, W))
    xa = x.permute(0, 2, 3, 1).contiguous().to(DEV, td)
    xb = None if x2 is None else x2.permute(0, 2, 3, 1).contiguous().to(DEV, td)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, taps * Cin).contiguous().to(DEV, td)
    out = torch.empty(P, Hout, Wout, Cout, device=DEV, dtype=td)
    p = GemmParams()
    p.src0, p.C0, p.ld0 = xa.data_ptr(), C0, C0
    if xb is not None:
        p.src1, p.C1, p.ld1 = xb.data_ptr(), C1, C1
    p.Hin, p.Win, p.Hout, p.Wout = H, W, Hout, Wout
    p.taps, p.stride, p.pad, p.ups = taps, stride, (1 if taps == 9 and stride == 1 else 0), ups
    p.B, p.ldb, p.b_layout = wp.data_ptr(), taps * Cin, 0
    p.M, p.N, p.K = P * Hout * Wout, Cout, taps * Cin
    p.alpha = 1.0
    keep = [xa, xb, wp]
    if bias is not None:
        b = bias.to(DEV)
        keep.append(b)
        p.bias = b.data_ptr()
    if temb is not None:
        tb = temb.contiguous().to(DEV)
        keep.append(tb)
        p.temb, p.temb_rows, p.temb_ld = tb.data_ptr(), tb.shape[0], tb.shape[1]
    if residual is not None:
        r = residual.permute(0, 2, 3, 1).contiguous().to(DEV, td)
        keep.append(r)
        p.residual, p.ldr = r.data_ptr(), Cout
    p.out, p.ldo = out.data_ptr(), Cout
    p.a_dtype = p.b_dtype = p.out_dtype = dtype
    stats = None
    if want_stats:
        stats = torch.full((p.M // 32, Cout // 4, 2), float("nan"), device=DEV)
        p.stats_out = stats.data_ptr()
    if ksplit is not None:   # split-K: "plan" = what the engine would use, or an explicit factor
        plan = lib.wdm_gemm_ksplit_plan(ctypes.byref(p))
        S = plan if ksplit == "plan" else int(ksplit)
        assert 2 <= S <= plan, (S, plan)
        scratch = torch.full((S, p.M, p.N), float("nan"), device=DEV)
        keep.append(scratch)
        p.ksplit, p.ksplit_scratch = S, scratch.data_ptr()
    st = lib.wdm_gemm(ctypes.byref(p), impl, torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "wdm_gemm")
    torch.cuda.synchronize()
    res = out.float().permute(0, 3, 1, 2).cpu()
    return (res, stats.cpu()) if want_stats else res


def ref_conv(x, x2, w, bias, stride, ups, temb, residual):
    xx = x if x2 is None else torch.cat([x, x2], 1)
    if ups:
        xx = F.interpolate(xx, scale_factor=2.0, mode="nearest")
    k = w.shape[2]
    if stride == 2:
        y = F.conv2d(F.pad(xx, (0, 1, 0, 1)), w, bias, stride=2)
    else:
        y = F.conv2d(xx, w, bias, padding=k // 2)
    if temb is not None:
        t = temb if temb.shape[0] > 1 else temb.expand(x.shape[0], -1)
        y = y + t[:, :, None, None]
    if residual is not None:
        y = y + residual
    return y


CONV_CASES = [
    # P, C0, C1, Cout, H, k, stride, ups, temb_rows, residual
    (2, 128, 0, 128, 16, 3, 1, 0, 0, False),
    (3, 96, 0, 128, 16, 3, 1, 0, 1, False),     # conv_in-like K, broadcast temb
    (2, 128, 128, 256, 8, 3, 1, 0, 2, True),    # concat + per-patch temb + residual
    (2, 256, 128, 128, 8, 1, 1, 0, 0, True),    # nin_shortcut over a concat
    (2, 128, 0, 128, 16, 3, 2, 0, 0, False),    # downsample (pad right/bottom)
    (2, 128, 0, 128, 8, 3, 1, 1, 0, False),     # upsample folded into addressing
    (1, 256, 0, 768, 8, 3, 1, 0, 0, False),     # N not a multiple of 128*k tile edge
    (5, 128, 0, 256, 4, 1, 1, 0, 0, False),     # M = 80: partial M tile
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_gemm_simt_fp32_conv_cases(case):
    P, C0, C1, Cout, H, k, stride, ups, trows, has_res = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(P, C0, H, H, generator=g)
    x2 = torch.randn(P, C1, H, H, generator=g) if C1 else None
    w = torch.randn(Cout, C0 + C1, k, k, generator=g) / (k * (C0 + C1) ** 0.5)
    bias = torch.randn(Cout, generator=g)
    temb = torch.randn(trows, Cout, generator=g) if trows else None
    Ho = H * 2 if ups else (H // 2 if stride == 2 else H)
    res = torch.randn(P, Cout, Ho, Ho, generator=g) if has_res else None
    ref = ref_conv(x, x2, w, bias, stride, ups, temb, res)
    out = run_conv(x, x2, w, bias, stride, ups, temb, res, 0, _lib.WDM_GEMM_IMPL_SIMT)
    # fp32 FFMA vs MKL-DNN: summation-order differences only
    assert (out - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    outb = run_conv(x, x2, w, bias, stride, ups, temb, res, 1, _lib.WDM_GEMM_IMPL_SIMT)
    # bf16 storage: inputs/outputs rounded to 8 mantissa bits
    assert (outb - ref).abs().max().item() <= 4e-2 * max(1.0, ref.abs().max().item())


TC_CASES = [
    # P, C0, C1, Cout, H, k, stride, ups, temb_rows, residual
    (2, 128, 0, 128, 16, 3, 1, 0, 0, False),    # BN=128, 2 rows of 64.. (16x16: 8 rows per tile)
    (3, 128, 0, 256, 16, 3, 1, 0, 1, True),     # BN=256, broadcast temb, residual, M = 768 (6 tiles)
    (2, 128, 0, 256, 8, 3, 1, 0, 2, True),      # per-patch temb + residual, 8x8 -> 2 patches per tile
    (4, 256, 128, 128, 8, 1, 1, 0, 0, True),    # 1x1 over a concat
    (2, 128, 0, 128, 16, 3, 2, 0, 0, False),    # stride-2 via TMA element strides (out 8x8)
    (2, 128, 0, 128, 64, 3, 2, 0, 0, False),    # stride-2 64 -> 32
    (1, 64, 0, 64, 64, 3, 1, 0, 0, False),      # BN=64, 64-wide rows
    (3, 192, 0, 768, 8, 3, 1, 0, 0, False),     # odd patch count: M = 192 -> partial last tile (8x8)
    (2, 1536, 0, 768, 8, 3, 1, 0, 1, True),     # deepest level shape, K = 13824
    (1, 128, 0, 128, 32, 1, 1, 0, 0, False),    # 1x1, 32-wide rows
    (5, 128, 0, 256, 4, 1, 1, 0, 0, True),      # M = 80: rows 64..79 valid in one warp only (M % 32 != 0), residual
    (3, 128, 0, 128, 4, 3, 1, 0, 1, False),     # M = 48, 4x4 patches (8 per tile), 3x3
]


@pytest.mark.parametrize("case", TC_CASES)
def test_gemm_tc_conv_cases(case):
    """tcgen05 implicit-GEMM vs torch fp32 conv on the SAME bf16-rounded inputs (differences: accumulation order
    and the bf16 rounding of the stored output, 2^-9 relative)."""
    P, C0, C1, Cout, H, k, stride, ups, trows, has_res = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    rb = lambda t: t.bfloat16().float()
    x = rb(torch.randn(P, C0, H, H, generator=g))
    x2 = rb(torch.randn(P, C1, H, H, generator=g)) if C1 else None
    w = rb(torch.randn(Cout, C0 + C1, k, k, generator=g) / (k * (C0 + C1) ** 0.5))
    bias = torch.randn(Cout, generator=g)
    temb = torch.randn(trows, Cout, generator=g) if trows else None
    Ho = H // 2 if stride == 2 else H
    res = rb(torch.randn(P, Cout, Ho, Ho, generator=g)) if has_res else None
    ref = ref_conv(x, x2, w, bias, stride, ups, temb, res)
    out = run_conv(x, x2, w, bias, stride, ups, temb, res, 1, _lib.WDM_GEMM_IMPL_TC)
    scale = max(1.0, ref.abs().max().item())
    err = (out - ref).abs().max().item()
    assert err <= 6e-3 * scale, (err, scale)
    # and it agrees with the CUDA-core kernel on the same data
    outs = run_conv(x, x2, w, bias, stride, ups, temb, res, 1, _lib.WDM_GEMM_IMPL_SIMT)
    assert (out - outs).abs().max().item() <= 1e-2 * scale


SPLITK_CASES = [
    # P, C0, C1, Cout, H, k, stride, ups, temb_rows, residual, ksplit
    (1, 1536, 0, 768, 8, 3, 1, 0, 1, True, "plan"),    # the deepest P = 1 shape: M = 64 (half a tile), 216 k-blocks
    (1, 512, 0, 512, 16, 3, 1, 0, 1, False, "plan"),   # 16x16 level, M = 256
    (2, 768, 0, 768, 8, 3, 1, 0, 2, True, "plan"),     # two patches in one tile, per-patch temb rows
    (1, 1024, 0, 512, 16, 3, 1, 0, 0, True, 2),        # smallest factor
    (1, 768, 0, 768, 8, 3, 1, 0, 0, False, 5),         # 108 k-blocks over 5 splits: uneven ranges
    (1, 512, 0, 256, 16, 3, 2, 0, 0, False, "plan"),   # stride-2 (out 8x8), M = 64
    (16, 768, 0, 768, 8, 3, 1, 0, 16, True, "plan"),   # 16 patches at 8x8: 96 64-wide tiles are too many -> 128-wide tiles, S = 3
]


@pytest.mark.parametrize("case", SPLITK_CASES)
def test_gemm_tc_splitk_cases(case):
    """Split-K form of the tcgen05 contraction (single-image latency path: few output tiles, deep K): the same result as the
    one-CTA-per-tile launch up to fp32 summation order, GroupNorm side-car included."""
    P, C0, C1, Cout, H, k, stride, ups, trows, has_res, ks = case
    g = torch.Generator().manual_seed(31)
    rb = lambda t: t.bfloat16().float()
    x = rb(torch.randn(P, C0, H, H, generator=g))
    w = rb(torch.randn(Cout, C0, k, k, generator=g) / (k * C0 ** 0.5))
    bias = torch.randn(Cout, generator=g)
    temb = torch.randn(trows, Cout, generator=g) if trows else None
    Ho = H // 2 if stride == 2 else H
    res = rb(torch.randn(P, Cout, Ho, Ho, generator=g)) if has_res else None
    ref = ref_conv(x, None, w, bias, stride, ups, temb, res)
    lib = _lib.load()
    n0 = lib.wdm_launch_counter()
    out, stats = run_conv(x, None, w, bias, stride, ups, temb, res, 1, _lib.WDM_GEMM_IMPL_TC, want_stats=True, ksplit=ks)
    assert lib.wdm_launch_counter() - n0 == 2          # contraction + reduce / epilogue
    plain, stats_plain = run_conv(x, None, w, bias, stride, ups, temb, res, 1, _lib.WDM_GEMM_IMPL_TC, want_stats=True)
    scale = max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() <= 6e-3 * scale
    assert (out - plain).abs().max().item() <= 2 ** -7 * scale          # one bf16 ulp where the fp32 sums round differently
    assert float((out != plain).float().mean()) < 0.02
    assert not torch.isnan(stats).any()
    assert (stats - stats_plain).abs().max().item() <= 1e-4 * max(1.0, stats_plain.abs().max().item())


@pytest.mark.parametrize("case", [(2, 256, 128, 128, 16), (2, 768, 768, 512, 8), (3, 128, 256, 0, 32), (2, 512, 512, 256, 16)])
@pytest.mark.parametrize("impl", [0, 1])
def test_gemm_fused_conv2_plus_shortcut(case, impl):
    """conv2 (3x3) and nin_shortcut (1x1 over cat[x1, x2]) as ONE contraction (tail_1x1): K = 9*C + C1 + C2."""
    P, C, C1, C2, H = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(31)
    rb = lambda t: t.bfloat16().float()
    h = rb(torch.randn(P, C, H, H, generator=g))
    x1 = rb(torch.randn(P, C1, H, H, generator=g))
    x2 = rb(torch.randn(P, C2, H, H, generator=g)) if C2 else None
    w2 = rb(torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5))
    wn = rb(torch.randn(C, C1 + C2, 1, 1, generator=g) / (C1 + C2) ** 0.5)
    b = torch.randn(C, generator=g)
    xx = x1 if x2 is None else torch.cat([x1, x2], 1)
    ref = F.conv2d(h, w2, None, padding=1) + F.conv2d(xx, wn, None) + b[None, :, None, None]
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    hd, x1d = nhwc(h), nhwc(x1)
    x2d = nhwc(x2) if x2 is not None else None
    wf = torch.cat([w2.permute(0, 2, 3, 1).reshape(C, 9 * C), wn.reshape(C, C1 + C2)], 1).contiguous().to(DEV, torch.bfloat16)
    out = torch.empty(P, H, H, C, device=DEV, dtype=torch.bfloat16)
    bd = b.to(DEV)
    p = GemmParams()
    p.src0, p.C0, p.ld0 = hd.data_ptr(), C, C
    p.src1, p.C1, p.ld1 = x1d.data_ptr(), C1, C1
    if x2d is not None:
        p.src2, p.C2, p.ld2 = x2d.data_ptr(), C2, C2
    p.tail_1x1 = 1
    p.Hin = p.Hout = p.Win = p.Wout = H
    p.taps, p.stride, p.pad = 9, 1, 1
    p.B, p.ldb, p.b_layout = wf.data_ptr(), 9 * C + C1 + C2, 0
    p.M, p.N, p.K = P * H * H, C, 9 * C + C1 + C2
    p.alpha, p.bias = 1.0, bd.data_ptr()
    p.out, p.ldo = out.data_ptr(), C
    p.a_dtype = p.b_dtype = p.out_dtype = 1
    _lib.check(lib.wdm_gemm(ctypes.byref(p), impl, torch.cuda.current_stream().cuda_stream), "wdm_gemm")
    torch.cuda.synchronize()
    res = out.float().permute(0, 3, 1, 2).cpu()
    assert (res - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(2, 256, 256, 16), (64, 768, 768, 8), (2, 128, 128, 32), (4, 64, 256, 8)])
def test_gemm_tc_subpixel_upsample_conv(case):
    """nearest-x2 upsample + 3x3 conv evaluated as four 2x2 phase convs on the source grid (ups = 2) equals
    F.interpolate + conv2d; its GroupNorm side-car sums to the sums of the output."""
    P, C, Cout, H = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(23)
    rb = lambda t: t.bfloat16().float()
    x = rb(torch.randn(P, C, H, H, generator=g))
    w = torch.randn(Cout, C, 3, 3, generator=g) / (3 * C ** 0.5)
    bias = torch.randn(Cout, generator=g)
    # phase weights, summed in fp32 then rounded to bf16 (what wdm_unet_create packs)
    rows = {0: [[0], [1, 2]], 1: [[0, 1], [2]]}
    wp = torch.zeros(4, Cout, 4, C)
    for py in (0, 1):
        for px in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = torch.zeros(Cout, C)
                    for yy in rows[py][ty]:
                        for xx in rows[px][tx]:
                            acc += w[:, :, yy, xx]
                    wp[py * 2 + px, :, ty * 2 + tx, :] = acc
    wp = rb(wp)
    # reference: the same (rounded) phase weights applied as 2x2 convs == upsample + conv with the unrounded sum
    ref = torch.zeros(P, Cout, 2 * H, 2 * H)
    xp = F.pad(x, (1, 1, 1, 1))
    for py in (0, 1):
        for px in (0, 1):
            k = wp[py * 2 + px].reshape(Cout, 2, 2, C).permute(0, 3, 1, 2)   # [Cout][C][ty][tx]
            y = F.conv2d(xp[:, :, py:py + H + 1, px:px + H + 1], k, bias)    # rows i+py-1+ty  (padded index +1)
            ref[:, :, py::2, px::2] = y
    direct = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, bias, padding=1)
    assert (ref - direct).abs().max().item() <= 3e-2 * max(1.0, direct.abs().max().item())  # bf16 weight rounding only
    xa = x.permute(0, 2, 3, 1).contiguous().to(DEV, torch.bfloat16)
    wd = wp.reshape(4 * Cout, 4 * C).contiguous().to(DEV, torch.bfloat16)
    out = torch.empty(P, 2 * H, 2 * H, Cout, device=DEV, dtype=torch.bfloat16)
    bd = bias.to(DEV)
    M = P * 4 * H * H
    stats = torch.full((M // 32, Cout // 4, 2), float("nan"), device=DEV)
    p = GemmParams()
    p.src0, p.C0, p.ld0 = xa.data_ptr(), C, C
    p.Hin, p.Win, p.Hout, p.Wout = H, H, 2 * H, 2 * H
    p.taps, p.stride, p.pad, p.ups = 4, 1, 0, 2
    p.B, p.ldb, p.b_layout = wd.data_ptr(), 4 * C, 0
    p.M, p.N, p.K = M, Cout, 4 * C
    p.alpha, p.bias = 1.0, bd.data_ptr()
    p.out, p.ldo = out.data_ptr(), Cout
    p.a_dtype = p.b_dtype = p.out_dtype = 1
    p.stats_out = stats.data_ptr()
    _lib.check(lib.wdm_gemm(ctypes.byref(p), _lib.WDM_GEMM_IMPL_TC, torch.cuda.current_stream().cuda_stream), "wdm_gemm")
    torch.cuda.synchronize()
    res = out.float().permute(0, 3, 1, 2).cpu()
    scale = max(1.0, ref.abs().max().item())
    assert (res - ref).abs().max().item() <= 6e-3 * scale
    st = stats.cpu()
    assert not torch.isnan(st).any()
    # per patch, per 4-channel block: the side-car sums equal the sums over the patch's output pixels
    per_patch = st.reshape(P, (4 * H * H) // 32, Cout // 4, 2).sum(1)
    rsum = ref.reshape(P, Cout // 4, 4, -1).sum(dim=(2, 3))
    rsq = (ref ** 2).reshape(P, Cout // 4, 4, -1).sum(dim=(2, 3))
    assert (per_patch[..., 0] - rsum).abs().max().item() <= 2e-3 * max(1.0, rsum.abs().max().item())
    assert (per_patch[..., 1] - rsq).abs().max().item() <= 2e-3 * max(1.0, rsq.abs().max().item())


@pytest.mark.parametrize("case", [(2, 128, 0, 256, 16, 3, 1, 0, 1, True), (4, 128, 0, 768, 8, 1, 1, 0, 0, False),
                                  (1, 128, 0, 128, 64, 3, 1, 0, 0, False)])
def test_gemm_tc_groupnorm_sidecar(case):
    """The tensor-core epilogue's GroupNorm side-car = per (32-row group, 4-channel block) sums of the stored tensor."""
    P, C0, C1, Cout, H, k, stride, ups, trows, has_res = case
    g = torch.Generator().manual_seed(17)
    rb = lambda t: t.bfloat16().float()
    x = rb(torch.randn(P, C0, H, H, generator=g))
    w = rb(torch.randn(Cout, C0, k, k, generator=g) / (k * C0 ** 0.5))
    bias = torch.randn(Cout, generator=g)
    temb = torch.randn(trows, Cout, generator=g) if trows else None
    res = rb(torch.randn(P, Cout, H, H, generator=g)) if has_res else None
    ref = ref_conv(x, None, w, bias, stride, ups, temb, res)          # [P, Cout, H, W] fp32 (pre-rounding values)
    out, stats = run_conv(x, None, w, bias, stride, ups, temb, res, 1, _lib.WDM_GEMM_IMPL_TC, want_stats=True)
    rows = ref.permute(0, 2, 3, 1).reshape(-1, Cout)                   # [M, Cout]
    blk = rows.reshape(rows.shape[0] // 32, 32, Cout // 4, 4)
    s_ref = torch.stack([blk.sum(dim=(1, 3)), (blk ** 2).sum(dim=(1, 3))], dim=-1)
    assert not torch.isnan(stats).any()
    assert (stats - s_ref).abs().max().item() <= 2e-3 * max(1.0, s_ref.abs().max().item())


@pytest.mark.parametrize("dtype", [0, 1])
@pytest.mark.parametrize("shape", [(3, 128, 0, 64), (2, 256, 128, 16), (2, 768, 768, 64), (1, 1280, 0, 256), (2, 512, 0, 256)])
def test_groupnorm_silu(dtype, shape):
    P, C0, C1, HW = shape
    lib = _lib.load()
    td = torch.float32 if dtype == 0 else torch.bfloat16
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(P, HW, C0, generator=g) * 3 + 1.5).to(td)
    b = (torch.randn(P, HW, C1, generator=g) - 2).to(td) if C1 else None
    C = C0 + C1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    full = a.float() if b is None else torch.cat([a.float(), b.float()], 2)
    for silu in (0, 1):
        ref = F.group_norm(full.permute(0, 2, 1).reshape(P, C, HW, 1), 32, gamma, beta, eps=1e-6)
        if silu:
            ref = ref * torch.sigmoid(ref)
        ref = ref.reshape(P, C, HW).permute(0, 2, 1)
        ad, bd = a.to(DEV), (b.to(DEV) if b is not None else None)
        out = torch.empty(P, HW, C, device=DEV, dtype=td)
        scratch = torch.empty(lib.wdm_groupnorm_scratch_bytes(P), dtype=torch.uint8, device=DEV)
        gd, bed = gamma.to(DEV), beta.to(DEV)
        st = lib.wdm_groupnorm_silu(ad.data_ptr(), C0, bd.data_ptr() if bd is not None else None, C1, dtype, P, HW, 1e-6,
                                    gd.data_ptr(), bed.data_ptr(), silu, out.data_ptr(), scratch.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "wdm_groupnorm_silu")
        err = (out.float().cpu() - ref).abs().max().item()
        assert err <= (2e-5 if dtype == 0 else 6e-2), err


@pytest.mark.parametrize("shape", [(3, 128, 0, 4096), (2, 256, 128, 1024), (2, 768, 768, 64), (1, 1280, 0, 256), (2, 512, 0, 256),
                                   (5, 128, 128, 4096), (70, 256, 0, 1024), (1, 768, 0, 64)])
def test_groupnorm_silu_from_sidecar_one_launch(shape):
    """GroupNorm + SiLU with the statistics reduced INSIDE the normalise kernel from the producers' side-cars (the engine's
    form: one launch per GroupNorm), vs F.group_norm; side-cars built here exactly as the tensor-core epilogue defines them."""
    P, C0, C1, HW = shape
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(P, HW, C0, generator=g) * 3 + 1.5).bfloat16()
    b = (torch.randn(P, HW, C1, generator=g) - 2).bfloat16() if C1 else None
    C = C0 + C1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)

    def sidecar(t):
        f = t.float().reshape(P * HW // 32, 32, t.shape[2] // 4, 4)
        return torch.stack([f.sum(dim=(1, 3)), (f ** 2).sum(dim=(1, 3))], dim=-1).contiguous()
    full = a.float() if b is None else torch.cat([a.float(), b.float()], 2)
    ad, bd = a.to(DEV), (b.to(DEV) if b is not None else None)
    sa, sb = sidecar(a).to(DEV), (sidecar(b).to(DEV) if b is not None else None)
    gd, bed = gamma.to(DEV), beta.to(DEV)
    for silu in (0, 1):
        ref = F.group_norm(full.permute(0, 2, 1).reshape(P, C, HW, 1), 32, gamma, beta, eps=1e-6)
        if silu:
            ref = ref * torch.sigmoid(ref)
        ref = ref.reshape(P, C, HW).permute(0, 2, 1)
        out = torch.full((P, HW, C), float("nan"), device=DEV, dtype=torch.bfloat16)
        st = lib.wdm_groupnorm_silu_sidecar(ad.data_ptr(), C0, sa.data_ptr(), bd.data_ptr() if bd is not None else None, C1,
                                            sb.data_ptr() if sb is not None else None, P, HW, 1e-6, gd.data_ptr(), bed.data_ptr(),
                                            silu, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "wdm_groupnorm_silu_sidecar")
        err = (out.float().cpu() - ref).abs().max().item()
        assert err <= 6e-2, err


def test_softmax_rows():
    lib = _lib.load()
    g = torch.Generator().manual_seed(4)
    for L in (64, 256):
        S = torch.randn(37, L, generator=g) * 4
        out = torch.empty(37, L, device=DEV)
        Sd = S.to(DEV)
        _lib.check(lib.wdm_softmax_rows(Sd.data_ptr(), 37, L, out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream), "softmax")
        assert (out.cpu() - torch.softmax(S, 1)).abs().max().item() <= 1e-6


def _engine(cfg, sd, precision, flags=0):
    return engine.UNetEngine(cfg, sd, DEV, precision=precision, flags=flags)


def test_unet_small_fp32_vs_reference_golden():
    g = golden("unet_small.npz")
    cfg = small_cfg()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    eng = _engine(cfg, sd, "fp32")
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    out = eng.forward(x.to(DEV), t.to(DEV)).cpu()
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    assert err <= 5e-5 * max(1.0, ref.abs().max().item()), err
    out1 = eng.forward(x.to(DEV), t[:1].to(DEV)).cpu()  # broadcast-t form used by the sampler
    assert (out1 - torch.from_numpy(g["out_t1"])).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())


def test_unet_small_bf16_simt_vs_oracle():
    g = golden("unet_small.npz")
    cfg = small_cfg()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    eng = _engine(cfg, sd, "bf16", flags=_lib.WDM_ENGINE_NO_TC)
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    out = eng.forward(x.to(DEV), t.to(DEV)).cpu()
    ref = torch.from_numpy(g["out"])
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= 3e-2, rel


def test_unet_small_bf16_tc_vs_oracle():
    g = golden("unet_small.npz")
    cfg = small_cfg()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    eng = _engine(cfg, sd, "bf16")
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    out = eng.forward(x.to(DEV), t.to(DEV)).cpu()
    ref = torch.from_numpy(g["out"])
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= 3e-2, rel
    eng.profile(True)
    eng.forward(x.to(DEV), t.to(DEV))
    tc_ms, tc_fl, tc_n, s_ms, s_fl, s_n = eng.profile_read()
    eng.profile(False)
    assert tc_n > 0, "the bf16 engine did not launch the tcgen05 kernel"


def test_unet_full_bf16_tc_vs_reference_golden():
    g = golden("unet_full.npz")
    cfg = O.default_config()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    eng = _engine(cfg, sd, "bf16")
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    x = torch.randn(2, 96, 64, 64, generator=gen)
    out = eng.forward(x.to(DEV), torch.from_numpy(g["t"]).to(DEV)).cpu()
    ref = torch.from_numpy(g["out"])
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel <= 3e-2, rel
    eng.profile(True)
    eng.forward(x.to(DEV), torch.from_numpy(g["t"]).to(DEV))
    tc_ms, tc_fl, tc_n, s_ms, s_fl, s_n = eng.profile_read()
    eng.profile(False)
    assert tc_n >= 70 and tc_fl > 0.95 * (tc_fl + s_fl), (tc_n, s_n, tc_fl, s_fl)


def test_unet_full_bf16_single_patch_runs_splitk_and_matches_the_batched_call():
    """The single-image operating point (BASELINE configs[0]: one latent patch per DDIM step): the engine splits the deep-K
    contractions over idle SMs (reduce launches on top of the 181 of a batched call) and patch 0 of the reference golden comes
    out the same as in the two-patch call (same parity gate)."""
    from wavedm_b200 import _lib
    g = golden("unet_full.npz")
    cfg = O.default_config()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    eng = _engine(cfg, sd, "bf16")
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    x = torch.randn(2, 96, 64, 64, generator=gen)
    t = torch.from_numpy(g["t"])
    lib = _lib.load()
    eng.forward(x[:1].to(DEV), t.to(DEV))
    n0 = lib.wdm_launch_counter()
    out1 = eng.forward(x[:1].to(DEV), t.to(DEV)).cpu()
    n1 = lib.wdm_launch_counter() - n0
    n0 = lib.wdm_launch_counter()
    out64 = eng.forward(x.repeat(32, 1, 1, 1).to(DEV), t.to(DEV)).cpu()     # one timestep for all patches (the sampler's case)
    n64 = lib.wdm_launch_counter() - n0
    assert n1 >= n64 + 20, (n1, n64)          # split-K: one reduce / epilogue launch per split contraction
    ref = torch.from_numpy(g["out"])[:1]
    rel = ((out1 - ref).norm() / ref.norm()).item()
    assert rel <= 3e-2, rel
    rel64 = ((out64[:1] - ref).norm() / ref.norm()).item()
    assert rel64 <= 3e-2 and abs(rel - rel64) <= 5e-3, (rel, rel64)


def test_unet_full_fp32_vs_reference_golden():
    g = golden("unet_full.npz")
    cfg = O.default_config()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(g["weight_sum"])) < 1e-6
    eng = _engine(cfg, sd, "fp32")
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    x = torch.randn(2, 96, 64, 64, generator=gen)
    out = eng.forward(x.to(DEV), torch.from_numpy(g["t"]).to(DEV)).cpu()
    ref = torch.from_numpy(g["out"])
    err = (out - ref).abs().max().item()
    assert err <= 5e-5 * max(1.0, ref.abs().max().item()), err


def test_workspace_too_small_is_reported():
    cfg = small_cfg()
    sd = O.init_state_dict(cfg, seed=1)
    eng = _engine(cfg, sd, "fp32")
    x = torch.zeros(2, 16, 16, eng.cin_pad, device=DEV)
    t = torch.zeros(1, device=DEV)
    out = torch.empty(2, 3, 16, 16, device=DEV)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=DEV)
    wp = (ws.data_ptr() + 1023) // 1024 * 1024
    st = eng.lib.wdm_unet_forward(eng.handle, x.data_ptr(), t.data_ptr(), 1, 2, out.data_ptr(), wp, 4096, None)
    assert st == _lib.WDM_ERR_WORKSPACE
    with pytest.raises(KeyError):
        engine.UNetEngine(cfg, {k: v for k, v in sd.items() if k != "conv_in.bias"}, DEV, precision="fp32")


_VARIANT_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {tests!r}); sys.path.insert(0, {repo!r})
from test_unet_gpu import run_conv, ref_conv
from wavedm_b200 import _lib
g = torch.Generator().manual_seed(5)
rb = lambda t: t.bfloat16().float()
P, C, H = {P}, 128, 64
x = rb(torch.randn(P, C, H, H, generator=g)); w = rb(torch.randn(128, C, 3, 3, generator=g) / (3 * C ** 0.5))
bias = torch.randn(128, generator=g); temb = torch.randn(1, 128, generator=g); res = rb(torch.randn(P, 128, H, H, generator=g))
ref = ref_conv(x, None, w, bias, 1, 0, temb, res)
out, stats = run_conv(x, None, w, bias, 1, 0, temb, res, 1, _lib.WDM_GEMM_IMPL_TC, want_stats=True)
scale = max(1.0, ref.abs().max().item())
assert (out - ref).abs().max().item() <= 6e-3 * scale
rows = ref.permute(0, 2, 3, 1).reshape(-1, 128); blk = rows.reshape(rows.shape[0] // 32, 32, 32, 4)
s_ref = torch.stack([blk.sum(dim=(1, 3)), (blk ** 2).sum(dim=(1, 3))], dim=-1)
assert not torch.isnan(stats).any() and (stats - s_ref).abs().max().item() <= 2e-3 * max(1.0, s_ref.abs().max().item())
print("variant ok")
"""


@pytest.mark.parametrize("env,P", [({"WDM_TC_SWAP": "1"}, 3), ({"WDM_TC_PAIR128X2": "1"}, 19), ({"WDM_TC_EPI_TMA": "0"}, 3)])
def test_gemm_tc_level0_variants(env, P):
    """The env-gated level-0 kernel variants (swap-AB accumulator, CTA pairs with two m-tiles per CTA, direct-store
    epilogue) stay bit-for-tolerance equal to the default path: Cout = 128 conv @64x64 with bias, temb, residual and the
    GroupNorm side-car. The switches are read once per process, hence the subprocess. P = 19: 304 m-tiles (>= 4 per CTA
    pair, an odd count of 4-tile super-tiles)."""
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    src = _VARIANT_SCRIPT.format(tests=os.path.dirname(os.path.abspath(__file__)), repo=REPO, P=P)
    r = subprocess.run([sys.executable, "-c", src], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout + r.stderr


# ---------------------------------------------------------------------------------------------- wavelet_in_unet
def wiu_cfg():
    return O.default_config(data__image_size=16, data__patch_size=64, data__wavelet_in_unet=True, model__ch=128,
                            model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8],
                            model__use_other_channels=False, model__in_channels=93, model__out_ch=48)


@pytest.mark.parametrize("dtype", [0, 1])
def test_gather_patches_dwt_vs_oracle(dtype):
    """Fused crop + DWT + concat + NHWC (wdm_gather_patches_dwt) == the C oracle's lifting-form DWT of the crops, bit for
    bit in fp32 (bf16: the same values rounded once); unaligned pixel corners; pad channels are zero."""
    from oracle import dwt_oracle as DO
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    B, H, W, R, Cpad = 2, 44, 52, 8, 128 if dtype else 96
    a, b = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
    pats = torch.tensor([[0, 0, 0], [1, 12, 20], [0, 5, 3], [1, 7, 17]], dtype=torch.int32)
    td = torch.bfloat16 if dtype else torch.float32
    out = torch.full((len(pats), R, R, Cpad), 7.0, dtype=td, device=DEV)
    ad, bd, pd = a.to(DEV), b.to(DEV), pats.to(DEV)
    st = lib.wdm_gather_patches_dwt(ad.data_ptr(), bd.data_ptr(), 2, B, H, W, pd.data_ptr(), len(pats), R, Cpad,
                                    out.data_ptr(), dtype, torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "wdm_gather_patches_dwt")
    torch.cuda.synchronize()
    for q, (n, hi, wi) in enumerate(pats.tolist()):
        crops = [t[n:n + 1, :, hi:hi + 4 * R, wi:wi + 4 * R].contiguous().numpy() for t in (a, b)]
        ref = np.concatenate([DO.dwt(c) for c in crops], axis=1)[0]            # [96, R, R]
        got = out[q].float().cpu().numpy().transpose(2, 0, 1)                     # [Cpad, R, R]
        if dtype == 0:
            assert np.array_equal(got[:96], ref)
        else:
            assert np.array_equal(got[:96], torch.from_numpy(ref).bfloat16().float().numpy())
        assert not got[96:].any()
    # error behaviour: patch larger than the image
    assert lib.wdm_gather_patches_dwt(ad.data_ptr(), bd.data_ptr(), 2, B, H, W, pd.data_ptr(), 1, 16, Cpad,
                                      out.data_ptr(), dtype, 0) == _lib.WDM_ERR_BAD_SHAPE


@pytest.mark.parametrize("ld", [48, 64])
def test_iwt_nhwc_vs_oracle(ld):
    from oracle import dwt_oracle as DO
    lib = _lib.load()
    g = torch.Generator().manual_seed(4)
    P, R = 3, 8
    y = torch.randn(P, R, R, ld, generator=g)
    x = torch.empty(P, 3, 4 * R, 4 * R, device=DEV)
    yd = y.to(DEV)
    _lib.check(lib.wdm_iwt4x4_nhwc(yd.data_ptr(), ld, P, R, x.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "wdm_iwt4x4_nhwc")
    torch.cuda.synchronize()
    ref = DO.iwt(y[..., :48].permute(0, 3, 1, 2).contiguous().numpy())
    assert np.array_equal(x.cpu().numpy(), ref)


def test_unet_wavelet_in_unet_vs_reference_golden():
    """DiffusionUNet(wavelet_in_unet=True): pixel-domain [P,6,64,64] -> [P,3,64,64] against the reference module's own
    output (golden): fp32 engine <= 5e-5 relative, bf16 tensor-core engine <= 3e-2 relative L2."""
    g = golden("unet_wiu.npz")
    cfg = wiu_cfg()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    x, t = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["t"]).to(DEV)
    ref = torch.from_numpy(g["out"])
    scale = ref.abs().max().item()
    e32 = engine.UNetEngine(cfg, sd, DEV, precision="fp32")
    out = e32.forward(x, t).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 5e-5 * max(1.0, scale)
    e16 = engine.UNetEngine(cfg, sd, DEV, precision="bf16")
    n0 = _lib.load().wdm_launch_counter()
    outb = e16.forward(x, t).cpu()
    assert _lib.load().wdm_launch_counter() > n0
    rel = ((outb - ref).pow(2).sum() / ref.pow(2).sum()).sqrt().item()
    assert rel <= 3e-2, rel
    # the broadcast-t form the sampler uses
    out1 = e32.forward(x, t[:1]).cpu()
    with torch.no_grad():
        ref1 = O.unet_forward(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"][:1]))
    assert (out1 - ref1).abs().max().item() <= 5e-5 * max(1.0, ref1.abs().max().item())


@pytest.mark.parametrize("case", [(4, 256, 512, 0), (6, 64, 768, 64)])   # (patches, tokens L, C, softmax_seg): 16x16 and the 8x8 pair form
@pytest.mark.parametrize("deferred", [False, True])
def test_gemm_tc_fused_softmax_epilogue(case, deferred):
    """Score GEMM with the row softmax in the epilogue (models/unet.py:176-182): normalised probabilities (single-group
    form) or, with row_scale_out, unnormalised exp(s - max) + 1/sum per row (two-group form); both against torch."""
    P, L, C, seg = case
    G = 2 if seg else 1
    Lg = G * L
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(P * L, C, generator=g) * 0.7).bfloat16()
    k = (torch.randn(P * L, C, generator=g) * 0.7).bfloat16()
    scale = C ** -0.5
    qd, kd = q.to(DEV), k.to(DEV)
    out = torch.full((P * L, Lg), float("nan"), device=DEV, dtype=torch.bfloat16)
    rs = torch.full((P * L,), float("nan"), device=DEV)
    p = GemmParams()
    p.src0, p.C0, p.ld0 = qd.data_ptr(), C, C
    hw = int(round(L ** 0.5))
    p.Hin = p.Hout = G * hw
    p.Win = p.Wout = hw
    p.taps, p.stride = 1, 1
    p.B, p.ldb, p.b_batch_stride, p.b_layout = kd.data_ptr(), C, Lg * C, 0
    p.M, p.N, p.K = P * L, Lg, C
    p.alpha = scale
    p.out, p.ldo = out.data_ptr(), Lg
    p.a_dtype = p.b_dtype = p.out_dtype = 1
    p.fuse_softmax, p.softmax_seg = 1, seg
    if deferred:
        p.row_scale_out = rs.data_ptr()
    _lib.check(lib.wdm_gemm(ctypes.byref(p), _lib.WDM_GEMM_IMPL_TC, torch.cuda.current_stream().cuda_stream), "wdm_gemm")
    torch.cuda.synchronize()
    qf, kf = q.float().view(P, L, C), k.float().view(P, L, C)
    s = torch.einsum("plc,pmc->plm", qf, kf) * scale                       # per patch [L, L]
    ref = torch.softmax(s, dim=2)
    got = out.float().cpu().view(P // G, G, L, G, L)
    if deferred:
        got = got * rs.cpu().view(P // G, G, L, 1, 1)
        assert torch.isfinite(rs).all()
    for a in range(G):
        for b in range(G):
            blk = got[:, a, :, b, :].reshape(P // G, L, L)
            if a == b:
                want = ref.view(P // G, G, L, L)[:, a]
                assert (blk - want).abs().max().item() <= 6e-3, (a, b)
            else:
                assert not blk.any()   # block-diagonal: other patches of the group get probability zero


# ---------------------------------------------------------------------------------------------- tc32 (fp32 on the tensor cores)
TC32_CASES = [
    # P, C, Cout, H, k, stride, subpix, temb_rows, residual
    (2, 128, 128, 16, 3, 1, False, 2, True),     # halo macro-stage kernel (Cout = 128 3x3)
    (2, 256, 256, 16, 3, 1, False, 1, False),    # CTA-pair kernel
    (2, 128, 128, 16, 3, 2, False, 0, False),    # downsample
    (3, 192, 384, 8, 1, 1, False, 0, True),      # 1x1 (odd patch count, 192-wide pair tiles)
    (2, 128, 256, 8, 3, 1, True, 0, False),      # nearest x2 upsample + 3x3 as four sub-pixel phases
]


@pytest.mark.parametrize("case", TC32_CASES)
def test_gemm_tc32_split_matches_fp32_conv(case):
    """WDM_ENGINE_TC32's contraction: 3-way bf16 split of both fp32 operands, six products per tap on the tcgen05 kernel with
    fp32 accumulation -- against the fp64 convolution, next to what the CUDA-core fp32 kernel achieves on the same problem."""
    P, C, Cout, H, k, stride, subpix, trows, use_res = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(1234 + C + Cout)
    x = torch.randn(P, C, H, H, generator=g)
    w = torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    temb = torch.randn(trows, Cout, generator=g) if trows else None
    Ho = H * 2 if subpix else (H // 2 if stride == 2 else H)
    res = torch.randn(P, Cout, Ho, Ho, generator=g) if use_res else None
    ref = ref_conv(x.double(), None, w.double(), bias.double(), stride, 1 if subpix else 0, None if temb is None else temb.double(),
                   None if res is None else res.double())
    st = torch.cuda.current_stream().cuda_stream
    xa = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    xs = torch.empty(P * H * H, 3 * C, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.wdm_split3_act(xa.data_ptr(), C, 0, 0, P * H * H, xs.data_ptr(), st), "wdm_split3_act")
    # the three pieces reproduce the fp32 values to 2^-24
    rec = xs.float().view(-1, 3, C).sum(1)
    assert float((rec - xa.view(-1, C)).abs().max()) <= 2.0 ** -23 * float(xa.abs().max())
    if subpix:
        # pre-summed phase weights (fp32): phase (py, px) tap (ty, tx) sums the 3x3 taps that land on the same source pixel
        sets = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
        wp = torch.zeros(4, Cout, 4, C)
        for py in range(2):
            for px in range(2):
                for ty in range(2):
                    for tx in range(2):
                        for dy in sets[py][ty]:
                            for dx in sets[px][tx]:
                                wp[py * 2 + px, :, ty * 2 + tx] += w[:, :, dy, dx]
        wflat, rows, taps = wp.reshape(4 * Cout, 4 * C).contiguous().to(DEV), 4 * Cout, 4
    else:
        taps = k * k
        wflat, rows = w.permute(0, 2, 3, 1).reshape(Cout, taps * C).contiguous().to(DEV), Cout
    ws = torch.empty(rows, taps * 6 * C, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.wdm_split3_weight(wflat.data_ptr(), rows, taps, C, ws.data_ptr(), 0, st), "wdm_split3_weight")
    ws5 = torch.empty(rows, taps * 5 * C, dtype=torch.bfloat16, device=DEV)
    wsm = torch.empty(rows, taps * C, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.wdm_split3_weight(wflat.data_ptr(), rows, taps, C, ws5.data_ptr(), wsm.data_ptr(), st), "wdm_split3_weight")
    b = bias.to(DEV)
    keep = [b]
    tb = r = None
    if temb is not None and not subpix:
        tb = temb.contiguous().to(DEV)
    if res is not None and not subpix:
        r = res.permute(0, 2, 3, 1).contiguous().to(DEV)

    def launch(B, split, kf, a_c0, out_t, bias_t, temb_t, res_t):
        p = GemmParams()
        p.src0, p.C0, p.ld0 = xs.data_ptr(), a_c0, 3 * C
        p.Hin, p.Win, p.Hout, p.Wout = H, H, Ho, Ho
        p.taps, p.stride, p.pad, p.ups = taps, stride, (1 if k == 3 and stride == 1 and not subpix else 0), (2 if subpix else 0)
        p.B, p.ldb, p.b_layout = B.data_ptr(), taps * kf * C, 0
        p.M, p.N, p.K, p.alpha = P * Ho * Ho, Cout, taps * kf * C, 1.0
        if bias_t is not None:
            p.bias = bias_t.data_ptr()
        if temb_t is not None:
            p.temb, p.temb_rows, p.temb_ld = temb_t.data_ptr(), trows, Cout
        if res_t is not None:
            p.residual, p.ldr = res_t.data_ptr(), Cout
        p.out, p.ldo = out_t.data_ptr(), Cout
        p.a_dtype = p.b_dtype = 1
        p.out_dtype = 0
        p.a_split3 = split
        _lib.check(lib.wdm_gemm(ctypes.byref(p), _lib.WDM_GEMM_IMPL_TC, st), f"wdm_gemm tc32 split={split}")

    if subpix:
        ref = ref_conv(x.double(), None, w.double(), bias.double(), 1, 1, None, None)
    # (a) all six products in one accumulator
    out6 = torch.empty(P, Ho, Ho, Cout, device=DEV)
    launch(ws, 1, 6, C, out6, b, tb, r)
    # (b) the engine's form: five small products (+ bias, temb, residual), then the dominant product + that result
    tmp = torch.empty(P, Ho, Ho, Cout, device=DEV)
    out = torch.empty(P, Ho, Ho, Cout, device=DEV)
    launch(ws5, 2, 5, C, tmp, b, tb, r)
    launch(wsm, 0, 1, C, out, None, None, tmp)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    err6 = float((out6.permute(0, 3, 1, 2).cpu().double() - ref).abs().max()) / scale
    err = float((out.permute(0, 3, 1, 2).cpu().double() - ref).abs().max()) / scale
    # the CUDA-core fp32 kernel on the same problem
    simt = run_conv(x, None, w, bias, stride, 1 if subpix else 0, None if subpix else temb, None if subpix else res, 0,
                    _lib.WDM_GEMM_IMPL_SIMT).double()
    err_simt = float((simt - ref).abs().max()) / scale
    print(f"tc32 {case}: max rel err two launches {err:.2e}, one accumulator {err6:.2e} (CUDA-core fp32 kernel {err_simt:.2e})")
    assert err6 <= 3e-5, err6
    assert err <= 4e-6, err


def test_unet_fp32_engine_uses_tensor_cores_and_matches_ffma_engine():
    """precision='fp32' runs WDM_ENGINE_TC32 by default: tcgen05 launches in the parity mode, output equal to the pure FFMA
    engine (WDM_ENGINE_NO_TC) to fp32 rounding, and the reference golden still holds (test_unet_full_fp32_vs_reference_golden)."""
    cfg = O.default_config()
    sd = O.init_state_dict(cfg, seed=61)
    x = torch.randn(3, 96, 64, 64, generator=torch.Generator().manual_seed(4))
    t = torch.tensor([321.0])
    e32 = engine.UNetEngine(cfg, sd, DEV, precision="fp32")
    assert e32.flags & _lib.WDM_ENGINE_TC32
    y = e32.forward(x, t).cpu()
    tc, simt = e32.counters()
    assert tc >= 60 and simt <= 16, (tc, simt)   # conv_in, conv_out's fallback and the attention bmm's stay on CUDA cores
    eff = engine.UNetEngine(cfg, sd, DEV, precision="fp32", flags=_lib.WDM_ENGINE_NO_TC)
    yf = eff.forward(x, t).cpu()
    assert eff.counters()[0] == 0
    rel = float((y - yf).abs().max()) / float(yf.abs().max())
    print(f"fp32 engine: tc32 vs FFMA max rel diff {rel:.2e}; launches tc {tc} simt {simt}")
    assert rel <= 2e-5


def test_unet_use_window_module_vs_reference_golden():
    """data.use_window through the public module (DiffusionUNet.forward): tile-major space-to-depth around the engine, conv_out
    with 3 p^2 = 12 channels; fp32 (tc32) and bf16 engines against the reference module's output."""
    from wavedm_b200.unet import DiffusionUNet
    g = golden("unet_win.npz")
    cfg = O.default_config(data__image_size=16, data__use_window=True, data__window_size=2, model__ch=128,
                           model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8],
                           model__use_other_channels=False, model__in_channels=21, model__out_ch=12)
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    net = DiffusionUNet(cfg).eval()
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).requires_grad_(False)
    x, t = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["t"]).to(DEV)
    ref = torch.from_numpy(g["out"])
    with torch.no_grad():
        net.engine_precision = "fp32"
        y32 = net(x, t).cpu()
        net.engine_precision = "bf16"
        y16 = net(x, t).cpu()
    assert y32.shape == ref.shape == (3, 3, 32, 32)
    e32 = float((y32 - ref).abs().max() / ref.abs().max())
    e16 = float((y16 - ref).norm() / ref.norm())
    print(f"use_window: fp32 max rel {e32:.2e}, bf16 rel L2 {e16:.2e}")
    assert e32 <= 5e-5 and e16 <= 3e-2
