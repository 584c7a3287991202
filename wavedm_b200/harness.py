"""Builds the public objects exactly the way ``eval_diffusion.py:93-98`` does, for synthetic runs (bench.py,
smoke, tests): seeded default-init UNet weights (no trained checkpoint exists) and a seeded,
synthesised HFRM checkpoint at the path the constructor loads."""
import argparse
import os
import tempfile

import torch


def synth_hfrm_checkpoint(seed: int = 61) -> str:
    from .hfrm import HFRM
    path = os.path.join(tempfile.gettempdir(), f"wavedm_b200_hfrm_seed{seed}_{os.getpid()}.pth")
    if not os.path.isfile(path):
        rng = torch.random.get_rng_state()
        torch.manual_seed(seed)
        net = HFRM(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=[2, 2, 2, 4], dec_blk_nums=[2, 2, 2, 2])
        torch.save(net.state_dict(), path)
        torch.random.set_rng_state(rng)
    return path


def build_restorer(config, device, sampling_timesteps=50, max_patches=64, seed=61, grid_r=16, broadcast=False,
                   image_folder="results/images"):
    """Returns a DiffusiveRestoration whose UNet has PyTorch-default-init weights under ``seed``. With an
    initialised process group the constructor wraps the UNet in DistributedDataParallel, which broadcasts rank 0's
    weights over NCCL (the reference's only inference-time collective, ddm_wavelet.py:168)."""
    from .ddm_wavelet import DenoisingDiffusion_Wavelet
    from .restoration import DiffusiveRestoration
    config.device = device
    args = argparse.Namespace(resume="", local_rank=torch.device(device).index or 0, sampling_timesteps=sampling_timesteps,
                              grid_r=grid_r, image_folder=image_folder, hfrm_ckpt=synth_hfrm_checkpoint(seed),
                              test_set="raindrop", max_patches=max_patches, seed=seed)
    rng = torch.random.get_rng_state()
    torch.manual_seed(seed)
    diffusion = DenoisingDiffusion_Wavelet(args, config)
    # the UNet weights every golden vector was made with: torch.manual_seed(seed); DiffusionUNet(config) ALONE (the
    # constructor above builds the HFRM first, which advances the generator)
    seeded_unet_weights(diffusion, seed)
    torch.random.set_rng_state(rng)
    diffusion.model.eval()
    return DiffusiveRestoration(diffusion, args, config)


def seeded_unet_weights(diffusion, seed: int = 61) -> None:
    """Loads ``torch.manual_seed(seed); DiffusionUNet(config).state_dict()`` (PyTorch default init, the reference's own
    module order) into ``diffusion.model`` -- identical on every rank, so it is consistent with the DDP broadcast."""
    from .unet import DiffusionUNet
    rng = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sd = DiffusionUNet(diffusion.config).state_dict()
    torch.random.set_rng_state(rng)
    net = diffusion.model.module if hasattr(diffusion.model, "module") else diffusion.model
    net.load_state_dict(sd, strict=True)
    net.invalidate_engine()
