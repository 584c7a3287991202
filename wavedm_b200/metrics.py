"""The parity metric and the PSNR variants ``restore()`` prints (``utils/metrics.py:7-11,43-86``).

``psnr_batch`` computes all three variants for a whole batch of image pairs in ONE kernel launch
(``wdm_psnr_stats``, csrc/wdm_metrics.cu) -- what ``DiffusiveRestoration.restore`` uses; the torch / numpy definitions
below are the reference-shaped API (``utils.torchPSNR`` etc.) and what the tests compare the kernel with."""
import math

import numpy as np
import torch


def psnr_batch(a: torch.Tensor, b: torch.Tensor):
    """a, b: [B, 3, H, W] fp32 CUDA tensors. Returns three float lists (one entry per image):
    torchPSNR(a, b), calculate_psnr_in_GPU(a, b, True), calculate_psnr(u8(a), u8(b), True) -- restoration.py:142-146."""
    from . import _lib
    assert a.shape == b.shape and a.ndim == 4 and a.shape[1] == 3
    if not a.is_cuda:
        raise RuntimeError("psnr_batch needs CUDA tensors (no CPU fallback; use torchPSNR / calculate_psnr on the host)")
    a = a.to(torch.float32).contiguous()
    b = b.to(device=a.device, dtype=torch.float32).contiguous()
    B, _, H, W = a.shape
    sse = torch.empty(B, 3, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        st = _lib.load().wdm_psnr_stats(a.data_ptr(), b.data_ptr(), B, H, W, sse.data_ptr(), _lib.current_stream_ptr(a.device))
    _lib.check(st, "wdm_psnr_stats")
    s = sse.cpu().tolist()

    def db(x, n):
        return float("inf") if x == 0 else 10.0 * math.log10(n / x)
    return ([db(r[0], 3 * H * W) for r in s], [db(r[1], H * W) for r in s], [db(r[2], H * W) for r in s])


def torchPSNR(tar_img, prd_img):
    """utils/metrics.py:7-11."""
    imdff = torch.clamp(prd_img, 0, 1) - torch.clamp(tar_img, 0, 1)
    rmse = (imdff ** 2).mean().sqrt()
    return 20 * torch.log10(1 / rmse)


def _y_weights(device):
    return torch.tensor([24.966, 128.553, 65.481], device=device)[None, :, None, None]


def to_y_channel_in_GPU(img):
    """utils/metrics.py:23-41 (y_only branch; the reference feeds RGB tensors through BGR weights)."""
    y = ((img * _y_weights(img.device)).sum(dim=1) + 16.0) / 255
    return y[:, None, :, :]


def calculate_psnr_in_GPU(img1, img2, test_y_channel=False):
    """utils/metrics.py:43-51."""
    if test_y_channel:
        img1, img2 = to_y_channel_in_GPU(img1), to_y_channel_in_GPU(img2)
    mse = torch.mean((img1 - img2) ** 2)
    return (20. * torch.log10(1. / torch.sqrt(mse))).cpu()


def to_y_channel(img):
    """utils/metrics.py (numpy Y channel of a [0,255] HWC image; BT.601, BGR weight order as the reference)."""
    img = img.astype(np.float32) / 255.
    y = np.dot(img, [24.966, 128.553, 65.481]) + 16.0
    return (y / 255.)[..., None] * 255.


def calculate_psnr(img1, img2, test_y_channel=False):
    """utils/metrics.py:53-86."""
    assert img1.shape == img2.shape and img1.shape[2] == 3
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    if test_y_channel:
        img1, img2 = to_y_channel(img1), to_y_channel(img2)
    mse = np.mean((img1 - img2) ** 2)
    if mse == 0:
        return float('inf')
    return 20. * np.log10(255. / np.sqrt(mse))
