"""DenoisingDiffusion_Wavelet -- drop-in for the reference's ``models/ddm_wavelet.py:127-506``.

Same constructor contract, attributes (``.model .wavelet_dec .wavelet_rec .generator .ema_helper
.optimizer .betas .num_timesteps .start_epoch .step``), checkpoint dict format and method signatures, so
``train_diffusion.py`` / ``eval_diffusion.py`` run unmodified against this package.
The sampling path (``sample_image`` / ``generalized_steps_overlapping`` / ``diffusive_restoration``) runs on
the sm_100a engine (DWT/IWT kernels, UNet engine, fused DDIM step); it has no PyTorch fallback.
``train`` keeps the reference's training step over the same parameters with PyTorch autograd (SURVEY.md
8f-3: API-complete, not accelerated).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import logging as wlogging
from . import optimize as woptimize
from .hfrm import HFRM
from .metrics import torchPSNR
from .sampler import DdimSampler
from .sampling import compute_alpha, generalized_steps  # noqa: F401  (re-exported like the reference module)
from .unet import DiffusionUNet
from .wavelet import WaveletTransform


def data_transform(X):
    return 2 * X - 1.0


def inverse_data_transform(X):
    return torch.clamp((X + 1.0) / 2.0, 0.0, 1.0)


class EMAHelper(object):
    """ddm_wavelet.py:35-84."""

    def __init__(self, mu=0.9999):
        self.mu = mu
        self.shadow = {}

    @staticmethod
    def _unwrap(module):
        return module.module if hasattr(module, "module") and isinstance(getattr(module, "module"), nn.Module) else module

    def register(self, module):
        for name, param in self._unwrap(module).named_parameters():
            if param.requires_grad:
                self.shadow[name] = param.data.clone()

    def update(self, module):
        """ddm_wavelet.py:48-53. CUDA parameters: one launch of ``wdm_adam_ema_step`` (EMA half) over all tensors, in place
        -- or nothing at all when the attached ``FusedAdam.step()`` already updated the shadow in its own launch."""
        if getattr(self, "_fused_update_done", False):
            self._fused_update_done = False
            return
        named = [(n, p) for n, p in self._unwrap(module).named_parameters() if p.requires_grad]
        if named and all(p.is_cuda for _, p in named):
            from .optimize import SegmentTable
            dev = named[0][1].device
            rows = []
            for n, p in named:
                sh = self.shadow[n]
                if sh.dtype != torch.float32 or p.dtype != torch.float32 or not sh.is_contiguous() or not p.is_contiguous() \
                        or sh.device != dev or p.device != dev:
                    raise TypeError("EMAHelper.update: dense contiguous float32 parameters / shadows on one device expected")
                if p.numel():
                    rows.append((p.data_ptr(), 0, 0, 0, sh.data_ptr(), p.numel()))
            tab = self.__dict__.get("_table")
            if tab is None or tab.device != dev:
                tab = self._table = SegmentTable(dev)
            tab.update(rows).launch(False, True, mu=self.mu)
            return
        for name, param in named:
            self.shadow[name].data = (1. - self.mu) * param.data + self.mu * self.shadow[name].data

    def ema(self, module):
        """ddm_wavelet.py:63-66. Written through ``param.copy_`` under no_grad (not ``param.data``) so the tensors' version
        counters move, and the packed CUDA engine of the module is dropped explicitly: the next forward re-packs."""
        inner = self._unwrap(module)
        with torch.no_grad():
            for name, param in inner.named_parameters():
                if param.requires_grad:
                    param.copy_(self.shadow[name].data)
        if hasattr(inner, "invalidate_engine"):
            inner.invalidate_engine()

    def ema_copy(self, module):
        inner = self._unwrap(module)
        module_copy = type(inner)(inner.config).to(inner.config.device)
        module_copy.load_state_dict(inner.state_dict())
        self.ema(module_copy)
        return module_copy

    def state_dict(self):
        return self.shadow

    def load_state_dict(self, state_dict):
        self.shadow = state_dict


def get_beta_schedule(beta_schedule, *, beta_start, beta_end, num_diffusion_timesteps):
    """ddm_wavelet.py:87-105 (float64, cast to float32 by the caller at :177)."""
    n = num_diffusion_timesteps
    if beta_schedule == "quad":
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float64) ** 2
    elif beta_schedule == "linear":
        betas = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    elif beta_schedule == "const":
        betas = beta_end * np.ones(n, dtype=np.float64)
    elif beta_schedule == "jsd":
        betas = 1.0 / np.linspace(n, 1, n, dtype=np.float64)
    elif beta_schedule == "sigmoid":
        betas = 1 / (np.exp(-np.linspace(-6, 6, n)) + 1) * (beta_end - beta_start) + beta_start
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (n,)
    return betas


def noise_estimation_loss(model, x0, t, e, b, total=None, use_global=False, inp_channels=3, pred_channels=3,
                          use_other_channels=False):
    """ddm_wavelet.py:108-124 (training objective; autograd path)."""
    if use_global:
        raise NotImplementedError("global_attn is out of scope")
    a = (1 - b).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
    x_inp = x0[:, :inp_channels]
    x_tar = x0[:, inp_channels:inp_channels + pred_channels]
    xt = x_tar * a.sqrt() + e * (1.0 - a).sqrt()
    x = torch.cat([xt, x0[:, inp_channels + pred_channels:]], dim=1) if use_other_channels else xt
    output = model(torch.cat([x_inp, x], dim=1), t.float())
    x0_pred = (xt - output * (1 - a).sqrt()) / a.sqrt()
    simple_loss = (e - output).square().sum(dim=(1, 2, 3))
    mse_loss = (x_tar - x0_pred).square().sum(dim=(1, 2, 3))
    return simple_loss.mean(dim=0), output, x0_pred, mse_loss.mean(dim=0)


class _ModuleHolder(nn.Module):
    """Stands in for DistributedDataParallel when no process group exists (single-process use): exposes the
    same ``.module`` attribute and forwards calls. With an initialised group the real DDP is used."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


class DenoisingDiffusion_Wavelet(object):
    HFRM_CKPT = "saved_models/raindrop/lastest.pth"  # ddm_wavelet.py:143 (cwd-relative, as in the reference)

    def __init__(self, args, config):
        super().__init__()
        self.args = args
        self.config = config
        self.device = config.device
        if getattr(config.data, "global_attn", False):
            raise NotImplementedError("global_attn (DiffusionUNet_Global) is out of scope (SURVEY.md 2.1 #11)")

        self.wavelet_dec = WaveletTransform(scale=2, dec=True).to(self.device)
        self.wavelet_rec = WaveletTransform(scale=2, dec=False).to(self.device)

        self.generator = HFRM(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=[2, 2, 2, 4],
                              dec_blk_nums=[2, 2, 2, 2]).to(self.device).eval()
        hfrm_path = getattr(args, "hfrm_ckpt", None) or self.HFRM_CKPT
        self.generator.load_state_dict(torch.load(hfrm_path, map_location=self.device), strict=True)
        self.generator.requires_grad_(False)
        self.generator.engine_precision = getattr(config.model, "engine_precision", None) or "bf16"

        self.model = DiffusionUNet(config)
        self.model.to(self.device)
        print("Total_params_model_real: {}M".format(sum(p.numel() for p in self.model.parameters()) / 1000000.0))

        self.ema_helper = EMAHelper()
        self.ema_helper.register(self.model)
        self.optimizer = woptimize.get_optimizer(self.config, self.model.parameters())
        if hasattr(self.optimizer, "attach_ema"):   # CUDA: Adam + EMA shadow update in one launch per training step
            self.optimizer.attach_ema(self.ema_helper, self.model)
        self.start_epoch, self.step = 0, 0

        print("my local rank", self.args.local_rank)
        if os.path.isfile(self.args.resume):
            self.load_ddm_ckpt(self.args.resume)

        if dist.is_available() and dist.is_initialized():
            if torch.device(self.device).type == "cuda":
                self.model = nn.parallel.DistributedDataParallel(self.model, device_ids=[self.args.local_rank],
                                                                 output_device=self.args.local_rank)
            else:
                self.model = nn.parallel.DistributedDataParallel(self.model)
        else:
            self.model = _ModuleHolder(self.model)

        betas = get_beta_schedule(beta_schedule=config.diffusion.beta_schedule, beta_start=config.diffusion.beta_start,
                                  beta_end=config.diffusion.beta_end,
                                  num_diffusion_timesteps=config.diffusion.num_diffusion_timesteps)
        betas = self.betas = torch.from_numpy(betas).float().to(self.device)
        self.num_timesteps = betas.shape[0]

    # ------------------------------------------------------------------------------------------ checkpoints
    def load_ddm_ckpt(self, load_path, ema=False):
        checkpoint = wlogging.load_checkpoint(load_path, self.device)
        self.start_epoch = checkpoint['epoch']
        self.step = checkpoint['step']
        net = self.model.module if hasattr(self.model, "module") else self.model
        net.load_state_dict(checkpoint['state_dict'], strict=True)
        if hasattr(net, "invalidate_engine"):
            net.invalidate_engine()
        self.optimizer.load_state_dict(checkpoint['optimizer'])
        self.ema_helper.load_state_dict(checkpoint['ema_helper'])
        if ema:
            self.ema_helper.ema(self.model)
        print("=> loaded checkpoint '{}' (epoch {}, step {})".format(load_path, checkpoint['epoch'], self.step))

    def all_wavlet_dec(self, x):
        return torch.cat([self.wavelet_dec(x[:, :3].contiguous()), self.wavelet_dec(x[:, 3:].contiguous())], dim=1)

    # ------------------------------------------------------------------------------------------ sampling
    def _unet(self):
        return self.model.module if hasattr(self.model, "module") else self.model

    def sample_image(self, x_cond, x, x_other=None, last=True, patch_locs=None, patch_size=None, total=None,
                     use_global=False, use_other=False):
        """ddm_wavelet.py:295-309."""
        skip = self.config.diffusion.num_diffusion_timesteps // self.args.sampling_timesteps
        seq = range(0, self.config.diffusion.num_diffusion_timesteps, skip)
        if patch_locs is not None:
            xs = self.generalized_steps_overlapping(x, x_cond, seq, self.model, self.betas, eta=0., corners=patch_locs,
                                                    p_size=patch_size, total=total, use_global=use_global,
                                                    x_other=x_other, use_other=use_other)
        else:
            xs = generalized_steps(x, x_cond, seq, self.model, self.betas, eta=0.)
        if last:
            xs = xs[0][-1]
        return xs

    def generalized_steps_overlapping(self, x, x_cond, seq, model, b, eta=0., corners=None, p_size=None,
                                      manual_batching=True, total=None, x_other=None, use_global=False,
                                      use_other=False):
        """ddm_wavelet.py:437-506 on the engine; returns (xs, x0_preds) with the reference's contract
        (xs[0] is the start tensor, all other entries are CPU tensors)."""
        if use_global:
            raise NotImplementedError("global_attn is out of scope")
        with torch.no_grad():
            if not self.config.data.begin_from_noise:
                # ddm_wavelet.py:445-447 ("not work" per the source; kept for signature completeness)
                a = (1 - b).cumprod(dim=0).index_select(0, torch.tensor(self.num_timesteps - 1).to(b.device)).view(-1, 1, 1, 1)
                x = x_cond[:, :, :, :] * a.sqrt() + x * (1.0 - a).sqrt()
            net = model.module if hasattr(model, "module") else model
            sampler = DdimSampler(net.engine(), max_patches=getattr(self.args, "max_patches", None))
            print("patch num :", len(corners))
            return sampler.sample_lists(x, x_cond, x_other if use_other else None, seq, b, corners, p_size, eta=eta)

    def diffusive_restoration(self, x_cond, x_other=None, r=None, last=True, total=None, use_global=False,
                              use_other=False):
        """ddm_wavelet.py:413-424."""
        p_size = self.config.data.patch_size if self.config.data.wavelet_in_unet else self.config.data.image_size
        h_list, w_list = self.overlapping_grid_indices(x_cond, output_size=p_size, r=r)
        corners = [(i, j) for i in h_list for j in w_list]
        x = torch.randn((x_cond.shape[0], self.config.model.pred_channels, x_cond.shape[2], x_cond.shape[3]),
                        device=self.device)
        return self.sample_image(x_cond, x, x_other=x_other, patch_locs=corners, last=last, patch_size=p_size,
                                 total=total, use_global=use_global, use_other=use_other)

    def overlapping_grid_indices(self, x_cond, output_size, r=None):
        """ddm_wavelet.py:426-435."""
        _, c, h, w = x_cond.shape
        r = 16 if r is None else r
        h_list = [i for i in range(0, h - output_size + 1, r)]
        w_list = [i for i in range(0, w - output_size + 1, r)]
        if h_list[-1] + output_size < h:
            h_list.append(h - output_size)
        if w_list[-1] + output_size < w:
            w_list.append(w - output_size)
        return h_list, w_list

    # ------------------------------------------------------------------------------------------ validation
    def restore(self, val_loader, validation='snow', r=None, epoch=0):
        """In-training validation (ddm_wavelet.py:340-409): first two images, PNG grid + PSNR print."""
        from einops import rearrange
        from torchvision.utils import make_grid
        image_folder = os.path.join(self.args.image_folder, self.config.data.dataset, validation)
        cfgm = self.config.model
        with torch.no_grad():
            all_samples = []
            y = None
            for i, (x, y, total) in enumerate(val_loader):
                print(f"starting processing from image {y}")
                x = x.flatten(start_dim=0, end_dim=1) if x.ndim == 5 else x
                x_all = data_transform(x.to(self.device))
                wiu = bool(getattr(self.config.data, "wavelet_in_unet", False))
                x_cond, x_gt = x_all[:, :3].contiguous(), x_all[:, 3:].contiguous()
                x_other, wd_wav = None, None
                if not wiu:   # ddm_wavelet.py:359-370: with wavelet_in_unet the network does the DWT / IWT itself
                    x_cond = self.wavelet_dec(x_cond)
                    x_gt = self.wavelet_dec(x_gt)
                    if cfgm.use_other_channels:
                        wd = self.generator(x[:, :3].to(self.device))
                        wd_wav = self.wavelet_dec(data_transform(wd))
                        x_other = wd_wav[:, cfgm.other_channels_begin:].contiguous()
                out_list = self.diffusive_restoration(x_cond, x_other=x_other, r=r, total=total, last=False,
                                                      use_other=cfgm.use_other_channels)
                x_output = out_list[1][-5].to(self.device)
                # (the reference leaves x_output_hrgt_cat undefined in wavelet_in_unet mode and raises NameError at :398;
                # here that panel of the grid shows the restored image again)
                x_hrgt = x_output
                if not wiu and cfgm.pred_channels < cfgm.in_channels and wd_wav is not None:   # :380-384
                    x_hrgt = torch.cat([x_output[:, :cfgm.pred_channels], x_gt[:, cfgm.pred_channels:]], dim=1)
                    x_output = torch.cat([x_output[:, :cfgm.pred_channels], wd_wav[:, cfgm.pred_channels:]], dim=1)
                if not wiu:                                                                      # :386-390
                    x_output = self.wavelet_rec(x_output.contiguous())
                    x_cond = self.wavelet_rec(x_cond)
                    x_hrgt = self.wavelet_rec(x_hrgt.contiguous())
                x_output = inverse_data_transform(x_output)
                x_cond_img = inverse_data_transform(x_cond)
                x_hrgt = inverse_data_transform(x_hrgt)
                gt = x[:, 3:]
                print("psnr", torchPSNR(gt.to(self.device), x_output))
                all_samples += [x_cond_img.cpu(), x_hrgt.cpu(), x_output.cpu(), gt]
                if i == 1:
                    break
            grid = rearrange(torch.stack(all_samples, 0), 'n b c h w -> (n b) c h w')
            grid = make_grid(grid, nrow=4)
            wlogging.save_image(grid, os.path.join(image_folder, f"{y}_output" + "_epoch" + str(epoch) + ".png"))

    # ------------------------------------------------------------------------------------------ training
    def train(self, DATASET):
        """ddm_wavelet.py:200-292: the reference's training step over the same parameters. On CUDA the DWT, the HFRM call and
        the parameter update (Adam + EMA shadow: one launch, csrc/wdm_optim.cu) run on this library's kernels; the UNet
        forward / backward is PyTorch autograd."""
        cfg, cfgm = self.config, self.config.model
        train_loader, _ = DATASET.get_loaders()
        num_of_pixel = cfgm.pred_channels * cfg.data.image_size ** 2
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        for epoch in range(self.start_epoch, cfg.training.n_epochs):
            print('epoch: ', epoch)
            data_start, data_time = time.time(), 0
            for i, (x, y, total) in enumerate(train_loader):
                x = x.flatten(start_dim=0, end_dim=1) if x.ndim == 5 else x
                n = x.size(0)
                data_time += time.time() - data_start
                self.model.train()
                self.step += 1
                x = x.to(self.device)
                x_all = data_transform(x)
                wiu = bool(getattr(cfg.data, "wavelet_in_unet", False))
                if not wiu:   # ddm_wavelet.py:227: with wavelet_in_unet the model consumes pixel-domain halves (half = 3)
                    x_all = self.all_wavlet_dec(x_all)
                half = x_all.shape[1] // 2
                if cfgm.use_other_channels:
                    if wiu and not cfgm.use_gt_in_train:
                        raise NotImplementedError("wavelet_in_unet with use_other_channels and use_gt_in_train=False is "
                                                  "undefined in the reference (x_output_wdnet_wav is never set, :247)")
                    if cfgm.use_gt_in_train:
                        hf = x_all[:, half:][:, cfgm.other_channels_begin:]
                    else:
                        with torch.no_grad():
                            wd = self.generator(x[:, :3])
                        hf = self.wavelet_dec(data_transform(wd))[:, cfgm.other_channels_begin:]
                    x_for_pred = torch.cat([x_all[:, :half + cfgm.pred_channels], hf], dim=1)
                else:
                    x_for_pred = x_all[:, :half + cfgm.pred_channels]
                e = torch.randn_like(x_for_pred[:, half:half + cfgm.pred_channels])
                # antithetic timestep sampling
                t = torch.randint(low=0, high=self.num_timesteps, size=(n // 2 + 1,)).to(self.device)
                t = torch.cat([t, self.num_timesteps - t - 1], dim=0)[:n]
                loss, e_pred, x0_pred, mse_loss = noise_estimation_loss(
                    self.model, x_for_pred, t, e, self.betas, inp_channels=half, pred_channels=cfgm.pred_channels,
                    use_other_channels=cfgm.use_other_channels)
                if self.step % 10 == 0:
                    print(f"step: {self.step}, loss: {loss.item()}, loss mean: {loss.item() / num_of_pixel}, "
                          f"mse loss mean: {mse_loss.item() / num_of_pixel}, data time: {data_time / (i + 1)} \n")
                self.optimizer.zero_grad()
                (mse_loss if cfg.training.use_mse else loss).backward()
                self.optimizer.step()
                self.ema_helper.update(self.model)
                data_start = time.time()
                if (world > 1 and self.step % cfg.training.validation_freq == 0) or (world == 1 and self.step % 10 == 0):
                    if rank == 0:
                        self.model.eval()
                        _, val_loader = DATASET.get_loaders(parse_patches=False, validation=self.args.test_set)
                        self.restore(val_loader, validation=self.args.test_set, r=self.args.grid_r, epoch=epoch)
                if self.step % cfg.training.snapshot_freq == 0 or self.step == 1:
                    if rank == 0:
                        wlogging.save_checkpoint({
                            'epoch': epoch + 1, 'step': self.step, 'state_dict': self._unet().state_dict(),
                            'optimizer': self.optimizer.state_dict(), 'ema_helper': self.ema_helper.state_dict(),
                            'params': self.args, 'config': self.config},
                            filename=os.path.join(cfg.data.data_dir, 'ckpts', cfg.data.dataset + '_epoch' + str(epoch + 1) + '_ddpm'))
