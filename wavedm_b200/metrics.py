"""The parity metric and the PSNR variants ``restore()`` prints (``utils/metrics.py:7-11,43-86``).
Not accelerated (scalar reductions once per image)."""
import numpy as np
import torch


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
