"""WaveletTransform -- drop-in for the reference's ``models/wavelet.py:6-50`` backed by the sm_100a
DWT / IWT kernels (``wdm_dwt4x4_fwd`` / ``wdm_iwt4x4_fwd``, csrc/wdm_dwt.cu).

Same constructor signature and the same ``.conv`` attribute (frozen fixed weights, so ``state_dict()`` and
``.to(device)`` behave as in the reference); ``forward`` does not run the conv, it calls the CUDA kernel
through the C ABI. The weights are the closed form of the pickle's ``rec4`` table
(``W[k][r][c] = 0.25 (-1)^(b0 c_hi + b1 r_hi + b2 c_lo + b3 r_lo)``) so the 839 KB
``wavelet_weights_c2.pkl`` is not a dependency; ``params_path`` is accepted and ignored.

Differences from the reference, all loud:
  * only ``scale=2`` (the only scale the reference ever instantiates: ddm_wavelet.py:134-135,
    unet.py:205-206) and ``transpose=True`` are implemented -> ``NotImplementedError`` otherwise
    (the reference itself crashes for dec=False, transpose=False: wavelet.py:45-49);
  * input must be a CUDA float32 tensor; there is no CPU path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib


def haar_packet_weights(scale: int = 2) -> torch.Tensor:
    """[3*ks*ks, 1, ks, ks] fp32 in the reference's `rec{ks}` layout (group-major: g*ks*ks + k)."""
    ks = 2 ** scale
    nb = scale  # bits per axis
    w = torch.empty(ks * ks, ks, ks, dtype=torch.float32)
    for k in range(ks * ks):
        for r in range(ks):
            for c in range(ks):
                e = 0
                # bit pairs, coarse first: (b0,b1) <-> (c_hi, r_hi), (b2,b3) <-> (c_lo, r_lo), ...
                for lvl in range(nb):
                    cb = (c >> (nb - 1 - lvl)) & 1
                    rb = (r >> (nb - 1 - lvl)) & 1
                    e += ((k >> (2 * lvl)) & 1) * cb + ((k >> (2 * lvl + 1)) & 1) * rb
                w[k, r, c] = (1.0 / ks) * (-1.0) ** e
    return w.repeat(3, 1, 1).unsqueeze(1).contiguous()


def _check_input(x: torch.Tensor, channels: int) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("wavedm_b200.WaveletTransform needs a CUDA tensor (no CPU fallback)")
    if x.dtype != torch.float32:
        raise TypeError(f"WaveletTransform expects float32, got {x.dtype}")
    if x.dim() != 4 or x.shape[1] != channels:
        raise ValueError(f"expected [N,{channels},H,W], got {tuple(x.shape)}")
    return x.contiguous()


def dwt4x4(x: torch.Tensor, pre_2xm1: bool = False, impl: int = _lib.WDM_WT_IMPL_AUTO) -> torch.Tensor:
    """[N,3,H,W] fp32 -> [N,48,H/4,W/4], channel 3k+g.  (models/wavelet.py:37-43)"""
    x = _check_input(x, 3)
    n, _, H, W = x.shape
    if H % 4 or W % 4:
        raise ValueError(f"H and W must be multiples of 4, got {H}x{W}")
    y = torch.empty((n, 48, H // 4, W // 4), dtype=torch.float32, device=x.device)
    if y.numel() == 0:
        return y
    flags = (_lib.WDM_DWT_PRE_2XM1 if pre_2xm1 else 0) | impl
    with torch.cuda.device(x.device):
        st = _lib.load().wdm_dwt4x4_fwd(x.data_ptr(), y.data_ptr(), n, H, W, flags, _lib.current_stream_ptr(x.device))
    _lib.check(st, "wdm_dwt4x4_fwd")
    return y


def iwt4x4(y: torch.Tensor, post_clamp: bool = False, impl: int = _lib.WDM_WT_IMPL_AUTO) -> torch.Tensor:
    """[N,48,h,w] fp32 -> [N,3,4h,4w].  (models/wavelet.py:44-49)"""
    y = _check_input(y, 48)
    n, _, h, w = y.shape
    x = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=y.device)
    if x.numel() == 0:
        return x
    flags = (_lib.WDM_IWT_POST_CLAMP if post_clamp else 0) | impl
    with torch.cuda.device(y.device):
        st = _lib.load().wdm_iwt4x4_fwd(y.data_ptr(), x.data_ptr(), n, h, w, flags, _lib.current_stream_ptr(y.device))
    _lib.check(st, "wdm_iwt4x4_fwd")
    return x


def iwt4x4_cat(lo: torch.Tensor, hi: torch.Tensor, post_clamp: bool = False) -> torch.Tensor:
    """IWT of ``torch.cat([lo, hi-bands], 1)`` without materialising the concatenation (models/restoration.py:111-135):
    ``lo`` [N, Clo, h, w] supplies the first Clo sub-band channels; ``hi`` is either a full 48-channel wavelet tensor
    (its channels [Clo, 48) are used) or holds exactly the remaining 48 - Clo bands."""
    for t in (lo, hi):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4:
            raise TypeError("iwt4x4_cat expects CUDA float32 NCHW tensors (no CPU fallback)")
    n, clo, h, w = lo.shape
    if hi.shape[0] != n or hi.shape[2:] != (h, w) or hi.shape[1] not in (48, 48 - clo):
        raise ValueError(f"incompatible shapes {tuple(lo.shape)} / {tuple(hi.shape)}")
    lo, hi = lo.contiguous(), hi.contiguous()
    x = torch.empty((n, 3, 4 * h, 4 * w), dtype=torch.float32, device=lo.device)
    if x.numel() == 0:
        return x
    with torch.cuda.device(lo.device):
        st = _lib.load().wdm_iwt4x4_cat(lo.data_ptr(), clo, hi.data_ptr(), hi.shape[1], x.data_ptr(), n, h, w,
                                        _lib.WDM_IWT_POST_CLAMP if post_clamp else 0, _lib.current_stream_ptr(lo.device))
    _lib.check(st, "wdm_iwt4x4_cat")
    return x


class _DwtFn(torch.autograd.Function):
    """The packet basis is orthonormal, so the adjoint of the DWT is the IWT (and vice versa)."""

    @staticmethod
    def forward(ctx, x):
        return dwt4x4(x)

    @staticmethod
    def backward(ctx, gy):
        return iwt4x4(gy)


class _IwtFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        return iwt4x4(y)

    @staticmethod
    def backward(ctx, gx):
        return dwt4x4(gx)


class WaveletTransform(nn.Module):
    def __init__(self, scale=1, dec=True, params_path='./models/wavelet_weights_c2.pkl', transpose=True):
        super().__init__()
        self.scale = scale
        self.dec = dec
        self.transpose = transpose
        ks = int(math.pow(2, self.scale))
        nc = 3 * ks * ks
        if dec:
            self.conv = nn.Conv2d(in_channels=3, out_channels=nc, kernel_size=ks, stride=ks, padding=0, groups=3,
                                  bias=False)
        else:
            self.conv = nn.ConvTranspose2d(in_channels=nc, out_channels=3, kernel_size=ks, stride=ks, padding=0,
                                           groups=3, bias=False)
        if scale not in (1, 2, 3):
            raise NotImplementedError(f"no closed form for scale={scale} (the reference's rec16 is irregular)")
        self.conv.weight.data = haar_packet_weights(scale)
        self.conv.weight.requires_grad = False

    def forward(self, x):
        if self.scale != 2:
            raise NotImplementedError("wavedm_b200 implements the scale=2 transform only "
                                      "(the only one the reference instantiates)")
        if not self.transpose:
            raise NotImplementedError("transpose=False is not implemented (never used by the reference; "
                                      "dec=False/transpose=False crashes there: models/wavelet.py:45-49)")
        if self.dec:
            return _DwtFn.apply(x) if x.requires_grad else dwt4x4(x)
        return _IwtFn.apply(x) if x.requires_grad else iwt4x4(x)
