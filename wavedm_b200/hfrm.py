"""HFRM -- the one-shot high-frequency refinement CNN (reference ``models/arch.py:132-253``).

Runs ONCE per image at full resolution, before the per-timestep loop (restoration.py:94). Inference
(no autograd graph) runs on the sm_100a HFRM engine (``csrc/wdm_hfrm.cu`` through ``wdm_hfrm_*`` of the C ABI): NHWC
bf16 (or fp32 in the parity mode), LayerNorm / SimpleGate / channel attention / PixelShuffle folded into the 1x1-conv
kernels. There is no CPU / PyTorch fallback for it: a CPU module under ``no_grad`` raises. The ``nn.Module`` definition
below keeps the reference's parameter names (``DenoisingDiffusion_Wavelet.__init__`` strict-loads the checkpoint,
ddm_wavelet.py:137-143) and is the differentiable definition autograd uses when a graph is being recorded.
"""
import ctypes
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class _HfrmConfig(ctypes.Structure):
    _fields_ = [("in_channel", ctypes.c_int), ("dim", ctypes.c_int), ("mid_blk_num", ctypes.c_int),
                ("n_levels", ctypes.c_int), ("enc_blk_nums", ctypes.c_int * 8), ("dec_blk_nums", ctypes.c_int * 8)]


class HfrmEngine:
    """The packed CUDA engine of one HFRM parameter set (``wdm_hfrm_create`` / ``wdm_hfrm_forward``)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, precision: str = "bf16", in_channel=3, dim=32,
                 mid_blk_num=6, enc_blk_nums=(2, 2, 2, 4), dec_blk_nums=(2, 2, 2, 2)):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("wavedm_b200.HfrmEngine needs a CUDA device (no CPU fallback)")
        if len(enc_blk_nums) != len(dec_blk_nums):
            raise ValueError("enc_blk_nums and dec_blk_nums must have the same length (arch.py:218-230)")
        self.precision = precision
        self.prec = {"fp32": _lib.WDM_PREC_FP32, "fp32_ffma": _lib.WDM_PREC_FP32, "bf16": _lib.WDM_PREC_BF16}[precision]
        cfg = _HfrmConfig()
        cfg.in_channel, cfg.dim, cfg.mid_blk_num, cfg.n_levels = in_channel, dim, mid_blk_num, len(enc_blk_nums)
        for i, (a, b) in enumerate(zip(enc_blk_nums, dec_blk_nums)):
            cfg.enc_blk_nums[i], cfg.dec_blk_nums[i] = int(a), int(b)
        self.cfg = cfg
        self.granule = 1 << cfg.n_levels
        n = self.lib.wdm_hfrm_param_count(ctypes.byref(cfg))
        if n < 0:
            raise _lib.WdmError(n, "wdm_hfrm_param_count")
        buf = ctypes.create_string_buffer(256)
        numel = ctypes.c_longlong()
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        chunks, names = [], set()
        for i in range(n):
            _lib.check(self.lib.wdm_hfrm_param_info(ctypes.byref(cfg), i, buf, 256, ctypes.byref(numel)), "wdm_hfrm_param_info")
            name = buf.value.decode()
            if name not in sd:
                raise KeyError(f"HFRM state_dict is missing '{name}'")
            if sd[name].numel() != int(numel.value):
                raise ValueError(f"'{name}': expected {int(numel.value)} elements, got {tuple(sd[name].shape)}")
            chunks.append(sd[name].detach().to(device=self.device, dtype=torch.float32).reshape(-1))
            names.add(name)
        extra = set(sd) - names
        if extra:
            raise KeyError(f"unexpected keys in the HFRM state_dict: {sorted(extra)[:5]} ...")
        flat = torch.cat(chunks)
        nbytes = self.lib.wdm_hfrm_packed_bytes(ctypes.byref(cfg), self.prec)
        if nbytes == 0:
            raise _lib.WdmError(_lib.WDM_ERR_BAD_ARG, "wdm_hfrm_packed_bytes")
        self.packed = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        pptr = (self.packed.data_ptr() + 255) // 256 * 256
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            st = self.lib.wdm_hfrm_create(ctypes.byref(cfg), self.prec, flat.data_ptr(), flat.numel(), pptr, nbytes,
                                          _lib.current_stream_ptr(self.device), ctypes.byref(handle))
            _lib.check(st, "wdm_hfrm_create")
            torch.cuda.current_stream(self.device).synchronize()   # flat may be freed after packing
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.wdm_hfrm_destroy(h)
            except Exception:
                pass
            self.handle = None

    def forward(self, x: torch.Tensor, max_batch: Optional[int] = None) -> torch.Tensor:
        """x: [B, 3, H, W] in the value range the reference feeds (restoration.py:94: the [0, 1] image). H and W must be
        multiples of 2^n_levels (arch.py has no padding path either). ``max_batch`` bounds the workspace."""
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        B, C, H, W = x.shape
        if C != self.cfg.in_channel:
            raise ValueError(f"HFRM input has {C} channels, expected {self.cfg.in_channel}")
        y = torch.empty_like(x)
        step = B if not max_batch else max(1, int(max_batch))
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            need = self.lib.wdm_hfrm_workspace_bytes(self.handle, n, H, W)
            if need == 0:
                raise _lib.WdmError(_lib.WDM_ERR_BAD_SHAPE, "wdm_hfrm_workspace_bytes")
            if self._ws is None or self._ws.numel() < need + 256:
                self._ws = None
                self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            wptr = (self._ws.data_ptr() + 255) // 256 * 256
            with torch.cuda.device(self.device):
                st = self.lib.wdm_hfrm_forward(self.handle, x[b0:b0 + n].data_ptr(), n, H, W, y[b0:b0 + n].data_ptr(), wptr,
                                               self._ws.numel() - (wptr - self._ws.data_ptr()),
                                               _lib.current_stream_ptr(self.device))
            _lib.check(st, "wdm_hfrm_forward")
        return y


class LayerNorm2d(nn.Module):
    """Per-pixel LayerNorm over channels (arch.py:6-43; the custom autograd function there is only a
    memory optimisation -- plain autograd gives the same values)."""

    def __init__(self, channels, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.eps = eps

    def forward(self, x):
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        y = (x - mu) / (var + self.eps).sqrt()
        return self.weight.view(1, -1, 1, 1) * y + self.bias.view(1, -1, 1, 1)


class SpatialAttn(nn.Module):
    """SimpleGate: product of the two channel halves (arch.py:132-141)."""

    def __init__(self, mid_dim):
        super().__init__()
        self.mid_dim = mid_dim

    def forward(self, x):
        return x[:, :self.mid_dim] * x[:, self.mid_dim:]


class ChannelAttn(nn.Module):
    """Global-average-pooled 1x1 gate (arch.py:143-155)."""

    def __init__(self, chan_dim):
        super().__init__()
        self.pool2d = nn.AdaptiveAvgPool2d(1)
        self.chan_conv = nn.Conv2d(chan_dim, chan_dim, kernel_size=1, bias=True)

    def forward(self, x):
        return x * self.chan_conv(self.pool2d(x))


class ResidualBlock(nn.Module):
    """arch.py:158-204."""

    def __init__(self, dim):
        super().__init__()
        self.conv1 = nn.Conv2d(dim, dim * 2, kernel_size=1)
        self.conv2 = nn.Conv2d(dim * 2, dim * 2, kernel_size=3, padding=1, groups=dim * 2)
        self.conv3 = nn.Conv2d(dim, dim, kernel_size=1)
        self.spatial_attn = SpatialAttn(mid_dim=dim)
        self.channel_attn = ChannelAttn(chan_dim=dim)
        self.conv4 = nn.Conv2d(dim, dim * 2, kernel_size=1)
        self.conv5 = nn.Conv2d(dim, dim, kernel_size=1)
        self.norm1 = LayerNorm2d(dim)
        self.norm2 = LayerNorm2d(dim)
        self.beta = nn.Parameter(torch.zeros((1, dim, 1, 1)))
        self.gamma = nn.Parameter(torch.zeros((1, dim, 1, 1)))

    def forward(self, x):
        y = self.conv2(self.conv1(self.norm1(x)))
        y = self.conv3(self.channel_attn(self.spatial_attn(y)))
        x = x + y * self.beta
        y = self.conv5(self.spatial_attn(self.conv4(self.norm2(x))))
        return x + y * self.gamma


class HFRM(nn.Module):
    """arch.py:206-253."""

    def __init__(self, in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=(2, 2, 2, 2), dec_blk_nums=(2, 2, 2, 2)):
        super().__init__()
        self._arch = dict(in_channel=in_channel, dim=dim, mid_blk_num=mid_blk_num, enc_blk_nums=tuple(enc_blk_nums),
                          dec_blk_nums=tuple(dec_blk_nums))
        self._engines, self._engine_key, self._engine_gen = {}, None, 0
        self.conv_in = nn.Conv2d(in_channel, dim, kernel_size=3, padding=1)
        self.encoders, self.decoders = nn.ModuleList(), nn.ModuleList()
        self.mid_blks = nn.ModuleList()
        self.ups, self.downs = nn.ModuleList(), nn.ModuleList()
        for num in enc_blk_nums:
            self.encoders.append(nn.Sequential(*[ResidualBlock(dim) for _ in range(num)]))
            self.downs.append(nn.Conv2d(dim, 2 * dim, 2, 2))
            dim *= 2
        self.mid_blks = nn.Sequential(*[ResidualBlock(dim) for _ in range(mid_blk_num)])
        for num in dec_blk_nums:
            self.ups.append(nn.Sequential(nn.Conv2d(dim, dim * 2, 1, bias=False), nn.PixelShuffle(2)))
            dim //= 2
            self.decoders.append(nn.Sequential(*[ResidualBlock(dim) for _ in range(num)]))
        self.conv_out = nn.Conv2d(dim, in_channel, kernel_size=3, padding=1)

    #: storage / arithmetic of the CUDA engine: "bf16" (tensor cores) or "fp32" (parity mode); DenoisingDiffusion_Wavelet
    #: sets it from config.model.engine_precision like the UNet's
    engine_precision = "bf16"
    #: images per engine call (bounds the activation workspace: ~33 MB per 256x256 image in bf16)
    engine_max_batch = 64

    def _param_version(self):
        return (getattr(self, "_engine_gen", 0),) + tuple(p._version for p in self.parameters()) + (next(self.parameters()).device,)

    def invalidate_engine(self):
        """Drop the packed CUDA engine (call after writing parameters through ``param.data``)."""
        self._engine_gen = getattr(self, "_engine_gen", 0) + 1
        self._engines = {}

    def engine(self, precision=None) -> HfrmEngine:
        precision = precision or self.engine_precision
        key = self._param_version()
        if key != getattr(self, "_engine_key", None):
            self._engines, self._engine_key = {}, key
        eng = self._engines.get(precision)
        if eng is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("HFRM inference needs the module on a CUDA device (wavedm_b200 has no CPU / PyTorch "
                                   "fallback on the restore() path)")
            eng = HfrmEngine(self.state_dict(), dev, precision=precision, **self._arch)
            self._engines[precision] = eng
        return eng

    def forward(self, x):
        # autograd whenever a graph is being recorded; otherwise the CUDA engine (same dispatch as DiffusionUNet.forward)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_autograd(x)
        return self.engine().forward(x, max_batch=self.engine_max_batch)

    def _forward_autograd(self, x):
        """Differentiable PyTorch definition (arch.py:234-253)."""
        inp = x
        H, W = x.shape[2:]
        x = self.conv_in(x)
        skips = []
        for enc, down in zip(self.encoders, self.downs):
            x = enc(x)
            skips.append(x)
            x = down(x)
        x = self.mid_blks(x)
        for dec, up, skip in zip(self.decoders, self.ups, reversed(skips)):
            x = dec(up(x) + skip)
        return (self.conv_out(x) + inp)[:, :, :H, :W]
