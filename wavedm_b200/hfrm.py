"""HFRM -- the one-shot high-frequency refinement CNN (reference ``models/arch.py:132-253``).

Runs ONCE per image at full resolution, outside the per-timestep loop, so it is not a kernel target this
round ("next"). It is kept as a plain PyTorch module because
``DenoisingDiffusion_Wavelet.__init__`` builds and strict-loads it (ddm_wavelet.py:137-143) and
``restore()`` calls it (restoration.py:94): parameter names match the reference checkpoint layout.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


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

    def forward(self, x):
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
