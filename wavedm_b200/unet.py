"""DiffusionUNet -- drop-in for the reference's ``models/unet.py:196-395`` (same network as
``models/unet_wav.py:10-155``): identical constructor, parameter names (the 332 state-dict keys) and
initialisation order (so a seed reproduces the reference's weights), but ``forward`` under ``torch.no_grad()``
runs the sm_100a engine (``wdm_unet_forward``) instead of ~430 ATen launches.

Inference has NO PyTorch fallback: without the CUDA library / a CUDA device ``forward`` raises.
With autograd enabled (``train_diffusion.py``) ``forward`` runs the differentiable PyTorch definition of the
same modules -- training is SURVEY.md 8(f)-3 ("next"): API-complete, not accelerated, and never used by
sampling, tests of the hot path, or bench.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine as _engine


def get_timestep_embedding(timesteps, embedding_dim):
    """models/unet.py:10-28 (host/PyTorch form; the engine evaluates the same table on the device)."""
    assert len(timesteps.shape) == 1
    half_dim = embedding_dim // 2
    emb = math.log(10000) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, dtype=torch.float32) * -emb).to(device=timesteps.device)
    emb = timesteps.float()[:, None] * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def nonlinearity(x):
    return x * torch.sigmoid(x)


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x):
        if self.with_conv:
            return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))
        return F.avg_pool2d(x, kernel_size=2, stride=2)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, temb):
        h = self.conv1(nonlinearity(self.norm1(x)))
        h = h + self.temb_proj(nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(nonlinearity(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1)

    def forward(self, x):
        h_ = self.norm(x)
        q, k, v = self.q(h_), self.k(h_), self.v(h_)
        b, c, h, w = q.shape
        w_ = torch.bmm(q.reshape(b, c, h * w).permute(0, 2, 1), k.reshape(b, c, h * w)) * (int(c) ** (-0.5))
        w_ = F.softmax(w_, dim=2)
        h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
        return x + self.proj_out(h_)


class DiffusionUNet(nn.Module):
    """Constructor mirrors models/unet.py:197-307 statement for statement in *module creation order*."""

    #: engine precision used by the no-grad forward ("bf16" = tcgen05 throughput mode, "fp32" = parity mode)
    engine_precision = "bf16"

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.use_window = config.data.use_window
        self.window_size = config.data.window_size
        self.use_wavelet_in_unet = config.data.wavelet_in_unet
        if self.use_window and self.use_wavelet_in_unet:
            raise NotImplementedError("use_window together with wavelet_in_unet: the reference applies both re-layouts "
                                      "in sequence (unet.py:347-350) and no channel count satisfies both; unused")
        if self.use_wavelet_in_unet:
            # models/unet.py:203-206 -- created first, as in the reference (module order = state-dict / RNG order)
            from .wavelet import WaveletTransform
            self.wavelet_dec = WaveletTransform(scale=2, dec=True)
            self.wavelet_rec = WaveletTransform(scale=2, dec=False)
        m = config.model
        ch, out_ch, ch_mult = m.ch, m.out_ch, tuple(m.ch_mult)
        in_channels = _engine.unet_in_channels(config)
        self.ch, self.temb_ch = ch, ch * 4
        self.dropout_p = float(m.dropout)
        self.num_resolutions, self.num_res_blocks = len(ch_mult), m.num_res_blocks
        self.resolution, self.in_channels = config.data.image_size, in_channels
        precision = getattr(m, "engine_precision", None)
        if precision is not None:
            self.engine_precision = precision

        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, self.temb_ch), nn.Linear(self.temb_ch, self.temb_ch)])
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = self.resolution
        in_ch_mult = (1,) + ch_mult
        self.down = nn.ModuleList()
        block_in = None
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch,
                                         dropout=m.dropout))
                block_in = block_out
                if curr_res in m.attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, m.resamp_with_conv)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=m.dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch,
                                       dropout=m.dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out, skip_in = ch * ch_mult[i_level], ch * ch_mult[i_level]
            for i_block in range(self.num_res_blocks + 1):
                if i_block == self.num_res_blocks:
                    skip_in = ch * in_ch_mult[i_level]
                block.append(ResnetBlock(in_channels=block_in + skip_in, out_channels=block_out,
                                         temb_channels=self.temb_ch, dropout=m.dropout))
                block_in = block_out
                if curr_res in m.attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, m.resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        if not m.resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False is not implemented by the engine")
        self._engines = {}
        self._engine_key = None
        self._engine_gen = 0

    # ------------------------------------------------------------------------------------------ engine
    def _param_version(self):
        # version counters catch in-place updates (optimizer.step, load_state_dict, no_grad copy_); writes through
        # ``param.data`` do NOT bump them, hence the explicit generation counter (invalidate_engine) as well
        return (self._engine_gen,) + tuple(p._version for p in self.parameters()) + (next(self.parameters()).device,)

    def invalidate_engine(self):
        """Drop the packed CUDA engine(s): the next inference forward re-packs the current parameter values. Call after
        writing parameters through ``param.data`` (EMAHelper.ema and load_ddm_ckpt do)."""
        self._engine_gen += 1
        self._engines = {}
        self._engine_key = None

    def engine(self, precision=None) -> "_engine.UNetEngine":
        """The CUDA engine for the current parameter values (re-packed when parameters change)."""
        precision = precision or self.engine_precision
        key = self._param_version()
        if key != self._engine_key:
            self._engines = {}
            self._engine_key = key
        eng = self._engines.get(precision)
        if eng is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("DiffusionUNet inference needs the module on a CUDA device "
                                   "(wavedm_b200 has no CPU / PyTorch fallback for the sampling path)")
            eng = _engine.UNetEngine(self.config, self.state_dict(), dev, precision=precision)
            self._engines[precision] = eng
        return eng

    # ------------------------------------------------------------------------------------------ use_window
    # data.use_window (models/unet.py:309-336,347-348,391-392): the 3-channel halves of the input are cut into p x p
    # tiles of side R that become channels (tile-major space-to-depth), the network runs on [P, 6 p^2, R, R], and its
    # [P, 3 p^2, R, R] output is pasted back to [P, 3, pR, pR]. Pure index re-layouts around the same engine.
    @staticmethod
    def to_win(x, p):
        B, C, H, W = x.shape
        return x.view(B, C, p, H // p, p, W // p).permute(0, 1, 2, 4, 3, 5).contiguous().view(B, -1, H // p, W // p)

    @staticmethod
    def win_back(x, p):
        B, C, H, W = x.shape
        return x.view(B, C // (p * p), p, p, H, W).permute(0, 1, 2, 4, 3, 5).contiguous().view(B, C // (p * p), H * p, W * p)

    def convert_image_to_patches(self, x):
        p = self.window_size
        return torch.cat([self.to_win(x[:, :3], p), self.to_win(x[:, 3:], p)], dim=1)

    def convert_patches_to_image(self, x):
        return self.win_back(x, self.window_size)

    def forward(self, x, t):
        if self.use_window:
            # pixel-domain [P, 6, pR, pR] in, [P, 3, pR, pR] out (the reference asserts after the re-layout, :351)
            assert x.shape[2] == x.shape[3] == self.resolution * self.window_size
            xw = self.convert_image_to_patches(x)
            if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
                return self.convert_patches_to_image(self._forward_autograd(xw, t))
            return self.convert_patches_to_image(self.engine().forward(xw, t))
        # wavelet_in_unet: pixel-domain [P, 6, 4R, 4R] in, [P, 3, 4R, 4R] out (the reference asserts after its DWT, :351)
        side = self.resolution * (4 if self.use_wavelet_in_unet else 1)
        assert x.shape[2] == x.shape[3] == side
        # autograd whenever a graph is being recorded (train_diffusion.py, but also eval-mode validation losses or guidance
        # that differentiate through the network); the engine (bf16 tensor-core by default -- config.model.engine_precision,
        # "fp32" for the parity mode) only under no_grad / for tensors that need no gradient
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_autograd(x, t)
        if self.training and self.dropout_p > 0.0:
            raise NotImplementedError("DiffusionUNet in train() mode under no_grad with dropout > 0: the CUDA engine has no "
                                      "dropout; call .eval() for inference")
        return self.engine().forward(x, t)

    # ------------------------------------------------------------------------------------------ training path
    def _forward_autograd(self, x, t):
        """Differentiable PyTorch definition (models/unet.py:353-389) -- training only, see module docstring."""
        if self.use_wavelet_in_unet:  # all_wavlet_dec, models/unet.py:338-344
            x = torch.cat([self.wavelet_dec(x[:, :3].contiguous()), self.wavelet_dec(x[:, 3:].contiguous())], dim=1)
        temb = get_timestep_embedding(t, self.ch)
        temb = self.temb.dense[1](nonlinearity(self.temb.dense[0](temb)))
        hs = [self.conv_in(x)]
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block](hs[-1], temb)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block](h)
                hs.append(h)
            if i_level != self.num_resolutions - 1:
                hs.append(self.down[i_level].downsample(hs[-1]))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(hs[-1], temb)), temb)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](torch.cat([h, hs.pop()], dim=1), temb)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        h = self.conv_out(nonlinearity(self.norm_out(h)))
        return self.wavelet_rec(h) if self.use_wavelet_in_unet else h
