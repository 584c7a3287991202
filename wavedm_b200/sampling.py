"""``utils/sampling.py`` of the reference (10-107), engine-backed.

``compute_alpha`` / ``data_transform`` / ``inverse_data_transform`` are the same few torch expressions;
``generalized_steps`` (whole-image DDIM, :23-44) and ``generalized_steps_overlapping`` (:47-107, the
non-wavelet twin of ddm_wavelet.py:437-506) run on the CUDA engine through ``wavedm_b200.sampler``.
"""
import torch

from .sampler import DdimSampler


def compute_alpha(beta, t):
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    a = (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)
    return a


def data_transform(X):
    return 2 * X - 1.0


def inverse_data_transform(X):
    return torch.clamp((X + 1.0) / 2.0, 0.0, 1.0)


def _engine_of(model):
    net = model.module if hasattr(model, "module") else model
    if not hasattr(net, "engine"):
        raise TypeError("model must be a wavedm_b200 DiffusionUNet (or a wrapper exposing .module)")
    return net.engine()


def generalized_steps(x, x_cond, seq, model, b, eta=0.):
    """utils/sampling.py:23-44: whole-image DDIM; the image must be exactly the UNet resolution."""
    eng = _engine_of(model)
    return DdimSampler(eng).sample_lists(x, x_cond, None, seq, b, [(0, 0)], eng.patch, eta=eta)


def generalized_steps_overlapping(x, x_cond, seq, model, b, eta=0., corners=None, p_size=None, manual_batching=True,
                                  total=None, use_global=False, use_FFT=False):
    """utils/sampling.py:47-107 (no x_other). use_global / use_FFT variants are out of scope."""
    if use_global or use_FFT:
        raise NotImplementedError("use_global / use_FFT sampling variants are not implemented")
    eng = _engine_of(model)
    return DdimSampler(eng).sample_lists(x, x_cond, None, seq, b, corners, p_size, eta=eta)
