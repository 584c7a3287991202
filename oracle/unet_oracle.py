"""oracle/unet_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU/torch-fp32 oracle; never on the product path).

A plain-PyTorch fp32 *functional* restatement of the reference's conditional diffusion UNet and DDIM
overlapping-patch sampler, driven directly by a reference-format ``state_dict`` (the 332 keys of
``/root/reference/models/unet.py:196-307``). Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.

Every function cites the reference lines it follows. The arithmetic lives in torch (the reference pins
torch==1.8.0, ``requirements.txt:8``; this image has 2.11 -- the semantics of conv2d / group_norm /
linear / bmm / softmax / interpolate(nearest) / cumprod used here are unchanged).

Parity pinning: the reference has no tests or golden vectors. This restatement is
pinned against the reference modules themselves, imported from /root/reference in the build container by
``oracle/make_golden.py``; the resulting vectors are committed under ``tests/golden/``.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------
# config helpers
# ----------------------------------------------------------------------------------------------------

def default_config(**over) -> SimpleNamespace:
    """configs/raindrop_wavelet.yml:1-63 as a namespace tree (only the keys the hot path reads)."""
    cfg = SimpleNamespace(
        data=SimpleNamespace(dataset="RainDrop", image_size=64, patch_size=256, lap=False, global_attn=False,
                             wavelet=True, wavelet_in_unet=False, use_window=False, window_size=2,
                             begin_from_noise=True, num_workers=32, data_dir="/data1/weather/",
                             conditional=True),
        model=SimpleNamespace(pred_channels=3, use_other_channels=True, other_channels_begin=3,
                              use_gt_in_train=True, in_channels=48, out_ch=3, ch=128, ch_mult=[1, 2, 4, 6],
                              num_res_blocks=2, attn_resolutions=[16], dropout=0.0, ema_rate=0.999, ema=True,
                              resamp_with_conv=True),
        diffusion=SimpleNamespace(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                                  num_diffusion_timesteps=1000),
        training=SimpleNamespace(use_mse=False, patch_n=8, batch_size=1, n_epochs=38000, n_iters=2000000,
                                 snapshot_freq=3000, validation_freq=3000),
        sampling=SimpleNamespace(batch_size=1, last_only=True),
        optim=SimpleNamespace(weight_decay=0.0, optimizer="Adam", lr=0.00004, amsgrad=False, eps=1e-8),
    )
    for k, v in over.items():
        sec, key = k.split("__")
        setattr(getattr(cfg, sec), key, v)
    return cfg


def unet_in_channels(cfg) -> int:
    """models/unet.py:212."""
    m = cfg.model
    if m.use_other_channels:
        return m.in_channels * 2 + m.pred_channels - m.other_channels_begin
    return m.in_channels + m.pred_channels


# ----------------------------------------------------------------------------------------------------
# UNet pieces
# ----------------------------------------------------------------------------------------------------

def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    """models/unet.py:10-28."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e).to(t.device)
    e = t.float()[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, (0, 1, 0, 0))
    return e


def swish(x: Tensor) -> Tensor:
    """models/unet.py:31-33."""
    return x * torch.sigmoid(x)


def group_norm(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """models/unet.py:36-37 (32 groups, eps 1e-6, affine)."""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def conv(x: Tensor, sd, p: str, stride=1, padding=0) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(x: Tensor, temb: Tensor, sd, p: str) -> Tensor:
    """models/unet.py:119-138 (dropout p=0 in the shipped config -> identity)."""
    h = conv(swish(group_norm(x, sd, p + ".norm1")), sd, p + ".conv1", padding=1)
    h = h + F.linear(swish(temb), sd[p + ".temb_proj.weight"], sd[p + ".temb_proj.bias"])[:, :, None, None]
    h = conv(swish(group_norm(h, sd, p + ".norm2")), sd, p + ".conv2", padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(x, sd, p + ".nin_shortcut")
    elif (p + ".conv_shortcut.weight") in sd:
        x = conv(x, sd, p + ".conv_shortcut", padding=1)
    return x + h


def attn_block(x: Tensor, sd, p: str) -> Tensor:
    """models/unet.py:168-193."""
    h_ = group_norm(x, sd, p + ".norm")
    q, k, v = conv(h_, sd, p + ".q"), conv(h_, sd, p + ".k"), conv(h_, sd, p + ".v")
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + conv(h_, sd, p + ".proj_out")


def downsample(x: Tensor, sd, p: str) -> Tensor:
    """models/unet.py:70-78 (resamp_with_conv=True: pad right/bottom by one, 3x3 stride 2)."""
    return conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0), sd, p + ".conv", stride=2)


def upsample(x: Tensor, sd, p: str) -> Tensor:
    """models/unet.py:51-56."""
    return conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd, p + ".conv", padding=1)


def haar_conv_weight() -> Tensor:
    """The `rec4` table of models/wavelet_weights_c2.pkl in closed form: [48, 1, 4, 4], row g*16 + k (SURVEY.md A.1;
    equality with the pickle is asserted by oracle/make_golden.py)."""
    return torch.from_numpy(np.tile(haar_packet_matrix(), (3, 1, 1))[:, None].copy())


def dwt_torch(x: Tensor) -> Tensor:
    """models/wavelet.py:37-43 statement for statement (grouped stride-4 conv, then the channel transpose)."""
    out = F.conv2d(x, haar_conv_weight(), stride=4, groups=3)
    osz = out.size()
    return out.view(osz[0], 3, -1, osz[2], osz[3]).transpose(1, 2).contiguous().view(osz)


def iwt_torch(y: Tensor) -> Tensor:
    """models/wavelet.py:44-49 (inverse channel transpose, then the grouped transposed conv)."""
    sz = y.size()
    xx = y.view(sz[0], -1, 3, sz[2], sz[3]).transpose(1, 2).contiguous().view(sz)
    return F.conv_transpose2d(xx, haar_conv_weight(), stride=4, groups=3)


def to_win(x: Tensor, p: int) -> Tensor:
    """models/unet.py:309-314."""
    B, C, H, W = x.shape
    x = x.view(B, C, p, H // p, p, W // p).permute(0, 1, 2, 4, 3, 5).contiguous()
    return x.view(B, -1, H // p, W // p)


def win_back(x: Tensor, p: int) -> Tensor:
    """models/unet.py:316-321."""
    B, C, H, W = x.shape
    x = x.view(B, C // (p ** 2), p, p, H, W).permute(0, 1, 2, 4, 3, 5).contiguous()
    return x.view(B, C // (p ** 2), H * p, W * p)


def unet_forward(sd: Dict[str, Tensor], cfg, x: Tensor, t: Tensor) -> Tensor:
    """models/unet.py:346-395 with use_window=False (identical network to models/unet_wav.py:115-155).
    data.wavelet_in_unet: the DWT of each 3-channel half on the way in (:338-344,349-350), the IWT on the way
    out (:393-394)."""
    m = cfg.model
    ch, ch_mult, nrb = m.ch, tuple(m.ch_mult), m.num_res_blocks
    nres = len(ch_mult)
    res = cfg.data.image_size
    wiu = bool(getattr(cfg.data, "wavelet_in_unet", False))
    win = bool(getattr(cfg.data, "use_window", False))
    if win:  # models/unet.py:323-331,347-348: each 3-channel half cut into p x p tiles that become channels
        x = torch.cat([to_win(x[:, :3], cfg.data.window_size), to_win(x[:, 3:], cfg.data.window_size)], dim=1)
    if wiu:
        x = torch.cat([dwt_torch(x[:, :3]), dwt_torch(x[:, 3:])], dim=1)
    assert x.shape[2] == x.shape[3] == res
    temb = timestep_embedding(t, ch)
    temb = F.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    temb = F.linear(swish(temb), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])

    hs = [conv(x, sd, "conv_in", padding=1)]
    cur = res
    for lv in range(nres):
        for ib in range(nrb):
            h = resnet_block(hs[-1], temb, sd, f"down.{lv}.block.{ib}")
            if cur in m.attn_resolutions:
                h = attn_block(h, sd, f"down.{lv}.attn.{ib}")
            hs.append(h)
        if lv != nres - 1:
            hs.append(downsample(hs[-1], sd, f"down.{lv}.downsample"))
            cur //= 2
    h = hs[-1]
    h = resnet_block(h, temb, sd, "mid.block_1")
    h = attn_block(h, sd, "mid.attn_1")
    h = resnet_block(h, temb, sd, "mid.block_2")
    for lv in reversed(range(nres)):
        for ib in range(nrb + 1):
            h = resnet_block(torch.cat([h, hs.pop()], dim=1), temb, sd, f"up.{lv}.block.{ib}")
            if cur in m.attn_resolutions:
                h = attn_block(h, sd, f"up.{lv}.attn.{ib}")
        if lv != 0:
            h = upsample(h, sd, f"up.{lv}.upsample")
            cur *= 2
    h = swish(group_norm(h, sd, "norm_out"))
    h = conv(h, sd, "conv_out", padding=1)
    if win:  # models/unet.py:333-336,391-392
        h = win_back(h, cfg.data.window_size)
    return iwt_torch(h) if wiu else h


def init_state_dict(cfg, seed: int = 61) -> Dict[str, Tensor]:
    """A reference-format state_dict with PyTorch default initialisation, built WITHOUT importing the
    reference: same module construction order as models/unet.py:196-307 so that, under the same
    ``torch.manual_seed``, it reproduces ``DiffusionUNet(config).state_dict()`` bit for bit
    (checked by oracle/make_golden.py and tests/test_oracle_pinning.py)."""
    import torch.nn as nn

    m = cfg.model
    ch, out_ch, ch_mult, nrb = m.ch, m.out_ch, tuple(m.ch_mult), m.num_res_blocks
    nres = len(ch_mult)
    temb_ch = ch * 4
    cin = unet_in_channels(cfg)
    g = torch.Generator().manual_seed(seed)  # noqa: F841  (kept for doc; torch.manual_seed is what nn uses)
    torch.manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def put(name, mod):
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()

    def resblock(name, ci, co):
        put(name + ".norm1", nn.GroupNorm(32, ci, eps=1e-6))
        put(name + ".conv1", nn.Conv2d(ci, co, 3, 1, 1))
        put(name + ".temb_proj", nn.Linear(temb_ch, co))
        put(name + ".norm2", nn.GroupNorm(32, co, eps=1e-6))
        put(name + ".conv2", nn.Conv2d(co, co, 3, 1, 1))
        if ci != co:
            put(name + ".nin_shortcut", nn.Conv2d(ci, co, 1, 1, 0))

    def attn(name, c):
        put(name + ".norm", nn.GroupNorm(32, c, eps=1e-6))
        for s in ("q", "k", "v", "proj_out"):
            put(name + "." + s, nn.Conv2d(c, c, 1, 1, 0))

    if getattr(cfg.data, "wavelet_in_unet", False):
        # unet.py:203-206: the two WaveletTransform modules are built first; their default conv init consumes the RNG
        # before the frozen Haar weights replace it (wavelet.py:19-34)
        nn.Conv2d(3, 48, 4, 4, 0, groups=3, bias=False)
        nn.ConvTranspose2d(48, 3, 4, 4, 0, groups=3, bias=False)
        sd["wavelet_dec.conv.weight"] = haar_conv_weight()
        sd["wavelet_rec.conv.weight"] = haar_conv_weight()
    put("temb.dense.0", nn.Linear(ch, temb_ch))
    put("temb.dense.1", nn.Linear(temb_ch, temb_ch))
    put("conv_in", nn.Conv2d(cin, ch, 3, 1, 1))
    cur = cfg.data.image_size
    in_ch_mult = (1,) + ch_mult
    block_in = None
    for lv in range(nres):
        block_in = ch * in_ch_mult[lv]
        block_out = ch * ch_mult[lv]
        # the reference builds `block` and `attn` ModuleLists interleaved per i_block (unet.py:243-253)
        for ib in range(nrb):
            resblock(f"down.{lv}.block.{ib}", block_in, block_out)
            block_in = block_out
            if cur in m.attn_resolutions:
                attn(f"down.{lv}.attn.{ib}", block_in)
        if lv != nres - 1:
            put(f"down.{lv}.downsample.conv", nn.Conv2d(block_in, block_in, 3, 2, 0))
            cur //= 2
    resblock("mid.block_1", block_in, block_in)
    attn("mid.attn_1", block_in)
    resblock("mid.block_2", block_in, block_in)
    for lv in reversed(range(nres)):
        block_out = ch * ch_mult[lv]
        skip_in = ch * ch_mult[lv]
        for ib in range(nrb + 1):
            if ib == nrb:
                skip_in = ch * in_ch_mult[lv]
            resblock(f"up.{lv}.block.{ib}", block_in + skip_in, block_out)
            block_in = block_out
            if cur in m.attn_resolutions:
                attn(f"up.{lv}.attn.{ib}", block_in)
        if lv != 0:
            put(f"up.{lv}.upsample.conv", nn.Conv2d(block_in, block_in, 3, 1, 1))
            cur *= 2
    put("norm_out", nn.GroupNorm(32, block_in, eps=1e-6))
    put("conv_out", nn.Conv2d(block_in, out_ch, 3, 1, 1))
    return sd


# ----------------------------------------------------------------------------------------------------
# diffusion schedule + sampler
# ----------------------------------------------------------------------------------------------------

def beta_schedule(cfg) -> Tensor:
    """models/ddm_wavelet.py:87-105 ('linear' branch) then the float32 cast at :177."""
    d = cfg.diffusion
    if d.beta_schedule != "linear":
        raise NotImplementedError(d.beta_schedule)
    b = np.linspace(d.beta_start, d.beta_end, d.num_diffusion_timesteps, dtype=np.float64)
    return torch.from_numpy(b).float()


def compute_alpha(beta: Tensor, t: Tensor) -> Tensor:
    """utils/sampling.py:10-13."""
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def overlapping_grid_indices(h: int, w: int, output_size: int, r: Optional[int] = None):
    """models/ddm_wavelet.py:426-435 (= models/restoration.py:187-196)."""
    r = 16 if r is None else r
    h_list = [i for i in range(0, h - output_size + 1, r)]
    w_list = [i for i in range(0, w - output_size + 1, r)]
    if h_list[-1] + output_size < h:
        h_list.append(h - output_size)
    if w_list[-1] + output_size < w:
        w_list.append(w - output_size)
    return h_list, w_list


def sampling_seq(num_timesteps: int, sampling_timesteps: int) -> List[int]:
    """models/ddm_wavelet.py:296-297."""
    skip = num_timesteps // sampling_timesteps
    return list(range(0, num_timesteps, skip))


def ddim_sample_overlapping(model_fn, x: Tensor, x_cond: Tensor, x_other: Optional[Tensor], seq: Sequence[int],
                            betas: Tensor, corners: Sequence[Tuple[int, int]], p_size: int,
                            batch_patches: int = 8) -> Tuple[List[Tensor], List[Tensor]]:
    """models/ddm_wavelet.py:437-506 with eta=0, begin_from_noise=True, use_global=False, generalised to a
    batch of B images by running the reference's batch-1 semantics independently per image (SURVEY.md
    fact 7: the reference writes only batch row 0, so B>1 means B independent B=1 runs).
    ``model_fn(x[P,Cin,p,p], t[1]) -> [P,3,p,p]``. Returns (xs, x0_preds) like the reference."""
    B = x.shape[0]
    seq = list(seq)
    seq_next = [-1] + seq[:-1]
    xs = [x]
    x0_preds: List[Tensor] = []
    mask = torch.zeros_like(x)
    for (hi, wi) in corners:
        mask[:, :, hi:hi + p_size, wi:wi + p_size] += 1
    for i_t, j_t in zip(reversed(seq), reversed(seq_next)):
        t = torch.ones(1) * i_t
        at = compute_alpha(betas, (torch.ones(1) * i_t).long())
        at_next = compute_alpha(betas, (torch.ones(1) * j_t).long())
        xt = xs[-1]
        et_out = torch.zeros_like(x)
        for b in range(B):
            pats = []
            for (hi, wi) in corners:
                parts = [x_cond[b:b + 1, :, hi:hi + p_size, wi:wi + p_size],
                         xt[b:b + 1, :, hi:hi + p_size, wi:wi + p_size]]
                if x_other is not None:
                    parts.append(x_other[b:b + 1, :, hi:hi + p_size, wi:wi + p_size])
                pats.append(torch.cat(parts, dim=1))
            pats = torch.cat(pats, dim=0)
            for i in range(0, len(corners), batch_patches):
                out = model_fn(pats[i:i + batch_patches], t)
                for idx, (hi, wi) in enumerate(corners[i:i + batch_patches]):
                    et_out[b, :, hi:hi + p_size, wi:wi + p_size] += out[idx]
        et = torch.div(et_out, mask)
        x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
        x0_preds.append(x0_t)
        c2 = (1 - at_next).sqrt()  # eta = 0 -> c1 = 0 (ddm_wavelet.py:500-501)
        xs.append(at_next.sqrt() * x0_t + c2 * et)
    return xs, x0_preds


def torch_psnr(tar: Tensor, prd: Tensor) -> Tensor:
    """utils/metrics.py:7-11."""
    d = torch.clamp(prd, 0, 1) - torch.clamp(tar, 0, 1)
    return 20 * torch.log10(1 / (d ** 2).mean().sqrt())


# ----------------------------------------------------------------------------------------------------
# DWT / IWT in numpy (closed form; models/wavelet.py:36-50)
# ----------------------------------------------------------------------------------------------------

def haar_packet_matrix() -> np.ndarray:
    """W[k, r, c] = 0.25 (-1)^(b0 c_hi + b1 r_hi + b2 c_lo + b3 r_lo)  (SURVEY.md A.1)."""
    W = np.zeros((16, 4, 4), np.float32)
    for k in range(16):
        b0, b1, b2, b3 = k & 1, (k >> 1) & 1, (k >> 2) & 1, (k >> 3) & 1
        for r in range(4):
            for c in range(4):
                W[k, r, c] = 0.25 * (-1) ** (b0 * (c >> 1) + b1 * (r >> 1) + b2 * (c & 1) + b3 * (r & 1))
    return W


def dwt_np(x: np.ndarray) -> np.ndarray:
    """[N,3,H,W] -> [N,48,H/4,W/4], channel 3k+g (models/wavelet.py:39-43)."""
    n, g, H, Wd = x.shape
    blk = x.reshape(n, g, H // 4, 4, Wd // 4, 4)
    y = np.einsum("krc,ngirjc->nkgij", haar_packet_matrix(), blk, optimize=False).astype(np.float32)
    return y.reshape(n, 16 * g, H // 4, Wd // 4)


def iwt_np(y: np.ndarray) -> np.ndarray:
    """[N,48,h,w] -> [N,3,4h,4w] (models/wavelet.py:45-49)."""
    n, c, h, w = y.shape
    yy = y.reshape(n, 16, c // 16, h, w)
    x = np.einsum("krc,nkgij->ngirjc", haar_packet_matrix(), yy, optimize=False).astype(np.float32)
    return x.reshape(n, c // 16, 4 * h, 4 * w)
