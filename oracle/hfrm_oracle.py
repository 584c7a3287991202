"""CPU oracle of the HFRM (TEST INFRASTRUCTURE ONLY -- nothing on the product path imports this file).

Functional torch-fp32 restatement of the reference's ``models/arch.py:132-253`` driven by a reference-format
``state_dict`` (same keys as ``models.arch.HFRM(...).state_dict()``): no ``nn.Module`` of this repo is involved, so a
bug in ``wavedm_b200/hfrm.py`` or in the CUDA engine cannot hide in the checker. Pinned against the output of the
unmodified reference module by ``tests/golden/hfrm.npz`` (``oracle/make_golden.py:golden_hfrm``, parameters filled
from seed 71, input from seed 72); ``tests/test_oracle_pinning.py`` re-checks that on every CPU run.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def fill_params(shapes: Dict[str, Sequence[int]], seed: int) -> Dict[str, torch.Tensor]:
    """oracle/make_golden.py:fill_params_deterministically -- the same pseudo-random values, in key order."""
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(tuple(s), generator=g) * 0.1 for k, s in shapes.items()}


def default_shapes(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=(2, 2, 2, 4), dec_blk_nums=(2, 2, 2, 2)):
    """State-dict keys and shapes of models.arch.HFRM(...) (arch.py:206-232; per block arch.py:159-183), in the module's
    own order: parameters of a module before those of its children."""
    sh: Dict[str, Sequence[int]] = {}

    def block(p, C):
        sh[p + "beta"], sh[p + "gamma"] = [1, C, 1, 1], [1, C, 1, 1]
        for name, co, ci, k in (("conv1", 2 * C, C, 1), ("conv2", 2 * C, 1, 3), ("conv3", C, C, 1),
                                ("channel_attn.chan_conv", C, C, 1), ("conv4", 2 * C, C, 1), ("conv5", C, C, 1)):
            sh[p + name + ".weight"], sh[p + name + ".bias"] = [co, ci, k, k], [co]
        for name in ("norm1", "norm2"):
            sh[p + name + ".weight"], sh[p + name + ".bias"] = [C], [C]

    sh["conv_in.weight"], sh["conv_in.bias"] = [dim, in_channel, 3, 3], [dim]
    d = dim
    for l, n in enumerate(enc_blk_nums):
        for i in range(n):
            block(f"encoders.{l}.{i}.", d)
        d *= 2
    dec = {}
    for l, n in enumerate(dec_blk_nums):
        d //= 2
        dec[l] = d
    # registration order of arch.py:212-232: conv_in, encoders, decoders, mid_blks, ups, downs, conv_out
    d = dim * 2 ** len(enc_blk_nums)
    for l, n in enumerate(dec_blk_nums):
        for i in range(n):
            block(f"decoders.{l}.{i}.", dec[l])
    for i in range(mid_blk_num):
        block(f"mid_blks.{i}.", d)
    for l in range(len(dec_blk_nums)):
        sh[f"ups.{l}.0.weight"] = [2 * d, d, 1, 1]
        d //= 2
    d = dim
    for l in range(len(enc_blk_nums)):
        sh[f"downs.{l}.weight"], sh[f"downs.{l}.bias"] = [2 * d, d, 2, 2], [2 * d]
        d *= 2
    sh["conv_out.weight"], sh["conv_out.bias"] = [in_channel, dim, 3, 3], [in_channel]
    return sh


def layer_norm2d(x, w, b, eps=1e-6):
    """arch.py:6-17 (LayerNormFunction.forward): per-pixel statistics over the channel axis, biased variance."""
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    y = (x - mu) / (var + eps).sqrt()
    return w.view(1, -1, 1, 1) * y + b.view(1, -1, 1, 1)


def residual_block(x, sd, p):
    """arch.py:185-204."""
    C = x.shape[1]
    inp = x
    x = layer_norm2d(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    x = F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"])
    x = F.conv2d(x, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1, groups=2 * C)
    x = x[:, :C] * x[:, C:]                                                   # SpatialAttn, arch.py:132-141
    s = F.conv2d(F.adaptive_avg_pool2d(x, 1), sd[p + "channel_attn.chan_conv.weight"],
                 sd[p + "channel_attn.chan_conv.bias"])                       # ChannelAttn, arch.py:143-155
    x = x * s
    x = F.conv2d(x, sd[p + "conv3.weight"], sd[p + "conv3.bias"])
    y = inp + x * sd[p + "beta"]
    x = F.conv2d(layer_norm2d(y, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), sd[p + "conv4.weight"], sd[p + "conv4.bias"])
    x = x[:, :C] * x[:, C:]
    x = F.conv2d(x, sd[p + "conv5.weight"], sd[p + "conv5.bias"])
    return y + x * sd[p + "gamma"]


@torch.no_grad()
def hfrm_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, mid_blk_num=6, enc_blk_nums=(2, 2, 2, 4),
                 dec_blk_nums=(2, 2, 2, 2)) -> torch.Tensor:
    """arch.py:234-253."""
    inp = x
    H, W = x.shape[2:]
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    encs = []
    for l, n in enumerate(enc_blk_nums):
        for i in range(n):
            x = residual_block(x, sd, f"encoders.{l}.{i}.")
        encs.append(x)
        x = F.conv2d(x, sd[f"downs.{l}.weight"], sd[f"downs.{l}.bias"], stride=2)
    for i in range(mid_blk_num):
        x = residual_block(x, sd, f"mid_blks.{i}.")
    for l, (n, skip) in enumerate(zip(dec_blk_nums, encs[::-1])):
        x = F.pixel_shuffle(F.conv2d(x, sd[f"ups.{l}.0.weight"]), 2)
        x = x + skip
        for i in range(n):
            x = residual_block(x, sd, f"decoders.{l}.{i}.")
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    return (x + inp)[:, :, :H, :W]
