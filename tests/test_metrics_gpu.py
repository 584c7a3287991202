"""wdm_psnr_stats (csrc/wdm_metrics.cu) against the reference-shaped metric definitions
(utils/metrics.py:7-11, :43-51, :53-86 as mirrored in wavedm_b200/metrics.py and pinned by tests/test_host_cpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W", [(1, 256, 256), (3, 120, 180), (2, 8, 4)])
def test_psnr_batch_matches_reference_definitions(B, H, W):
    from wavedm_b200 import metrics
    g = torch.Generator().manual_seed(B * 1000 + H)
    gt = torch.rand(B, 3, H, W, generator=g)
    out = gt + 0.05 * torch.randn(B, 3, H, W, generator=g)
    out[:, :, : H // 4] = out[:, :, : H // 4] * 2.0 - 0.4   # values outside [0, 1]: the clamps matter
    t, yg, ynp = metrics.psnr_batch(gt.cuda(), out.cuda())
    for b in range(B):
        ref_t = float(metrics.torchPSNR(gt[b:b + 1], out[b:b + 1]))
        ref_g = float(metrics.calculate_psnr_in_GPU(gt[b:b + 1], out[b:b + 1], True))

        def u8(x):
            return torch.clamp(x[0] * 255, 0, 255).numpy().transpose((1, 2, 0))
        ref_n = float(metrics.calculate_psnr(u8(gt[b:b + 1]), u8(out[b:b + 1]), True))
        assert abs(t[b] - ref_t) < 1e-3, (t[b], ref_t)
        assert abs(yg[b] - ref_g) < 1e-3, (yg[b], ref_g)
        assert abs(ynp[b] - ref_n) < 1e-3, (ynp[b], ref_n)


def test_psnr_batch_identical_images_and_errors():
    from wavedm_b200 import metrics
    x = torch.rand(2, 3, 16, 16).cuda()
    t, yg, ynp = metrics.psnr_batch(x, x.clone())
    assert t == [float("inf")] * 2 and yg == [float("inf")] * 2 and ynp == [float("inf")] * 2
    with pytest.raises(RuntimeError):
        metrics.psnr_batch(x.cpu(), x.cpu())
