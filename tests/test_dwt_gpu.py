"""GPU parity: the sm_100a DWT/IWT kernels (through the C ABI / WaveletTransform) against the C oracle
(bit-exact vs its lifting form), the committed reference goldens, and size-independent properties at
BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import dwt_oracle as DO
from wavedm_b200 import _lib
from wavedm_b200.wavelet import WaveletTransform, dwt4x4, iwt4x4

pytestmark = pytest.mark.gpu

IMPLS = [_lib.WDM_WT_IMPL_DIRECT, _lib.WDM_WT_IMPL_TMA]


def _dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("shape", [(2, 3, 8, 12), (1, 3, 4, 4), (3, 3, 64, 64), (1, 3, 480, 720), (2, 3, 36, 100)])
def test_direct_bit_exact_vs_oracle(shape):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)
    y = dwt4x4(x.to(_dev()), impl=_lib.WDM_WT_IMPL_DIRECT).cpu().numpy()
    assert np.array_equal(y, DO.dwt(x.numpy()))
    yy = torch.randn(shape[0], 48, shape[2] // 4, shape[3] // 4, generator=g)
    xr = iwt4x4(yy.to(_dev()), impl=_lib.WDM_WT_IMPL_DIRECT).cpu().numpy()
    assert np.array_equal(xr, DO.iwt(yy.numpy()))


@pytest.mark.parametrize("shape", [(2, 3, 128, 128), (1, 3, 480, 720), (3, 3, 256, 256), (1, 3, 32, 144), (1, 3, 100, 400)])
def test_tma_bit_exact_vs_oracle(shape):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(shape, generator=g)
    y = dwt4x4(x.to(_dev()), impl=_lib.WDM_WT_IMPL_TMA).cpu().numpy()
    assert np.array_equal(y, DO.dwt(x.numpy()))
    yy = torch.randn(shape[0], 48, shape[2] // 4, shape[3] // 4, generator=g)
    xr = iwt4x4(yy.to(_dev()), impl=_lib.WDM_WT_IMPL_TMA).cpu().numpy()
    assert np.array_equal(xr, DO.iwt(yy.numpy()))


@pytest.mark.parametrize("impl", IMPLS)
def test_fused_transforms_bit_exact(impl):
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 64, 256, generator=g)
    y = dwt4x4(x.to(_dev()), pre_2xm1=True, impl=impl).cpu().numpy()
    assert np.array_equal(y, DO.dwt(x.numpy(), flags=1))
    yy = torch.randn(2, 48, 16, 64, generator=g)
    xr = iwt4x4(yy.to(_dev()), post_clamp=True, impl=impl).cpu().numpy()
    assert np.array_equal(xr, DO.iwt(yy.numpy(), flags=1))
    assert xr.min() >= 0.0 and xr.max() <= 1.0


def test_module_vs_reference_golden():
    g = golden("dwt_kat.npz")
    dec = WaveletTransform(scale=2, dec=True).to(_dev())
    rec = WaveletTransform(scale=2, dec=False).to(_dev())
    assert np.array_equal(dec.conv.weight.cpu().numpy(), g["rec4"])
    y = dec(torch.from_numpy(g["x"]).to(_dev())).cpu().numpy()
    # tolerance: summation order of the reference's conv backend (16 terms of O(1))
    assert np.abs(y - g["dwt_x"]).max() <= 2e-6
    x = rec(torch.from_numpy(g["y"]).to(_dev())).cpu().numpy()
    assert np.abs(x - g["iwt_y"]).max() <= 2e-6
    # integer-valued data is exact in any order: bit-exact index layout vs the reference module
    assert np.array_equal(dec(torch.from_numpy(g["xi"]).to(_dev())).cpu().numpy(), g["dwt_xi"])
    assert np.array_equal(rec(torch.from_numpy(g["dwt_xi"]).to(_dev())).cpu().numpy(), g["xi"])


@pytest.mark.parametrize("impl", IMPLS)
def test_full_size_properties(impl):
    """BASELINE config sizes (B=64 @256^2 and 512^2): round trip, linearity, energy (orthonormality), LL = 4*mean."""
    for shape in [(64, 3, 256, 256), (8, 3, 512, 512)]:
        g = torch.Generator(device="cuda").manual_seed(11)
        x = torch.randn(shape, device=_dev(), generator=g)
        x2 = torch.randn(shape, device=_dev(), generator=g)
        y = dwt4x4(x, impl=impl)
        assert y.shape == (shape[0], 48, shape[2] // 4, shape[3] // 4)
        xr = iwt4x4(y, impl=impl)
        assert (xr - x).abs().max().item() <= 2e-6 * 4
        rel = abs(y.double().pow(2).sum().item() / x.double().pow(2).sum().item() - 1.0)
        assert rel < 1e-6
        y2 = dwt4x4(x2, impl=impl)
        ysum = dwt4x4(x + x2, impl=impl)
        assert (ysum - (y + y2)).abs().max().item() <= 4e-6
        ll = y.view(shape[0], 16, 3, shape[2] // 4, shape[3] // 4)[:, 0]
        blockmean = torch.nn.functional.avg_pool2d(x, 4) * 4.0
        assert (ll - blockmean).abs().max().item() <= 4e-6
        # the two variants agree bit for bit
        assert torch.equal(y, dwt4x4(x, impl=_lib.WDM_WT_IMPL_DIRECT))


def test_edge_cases_and_errors():
    dec = WaveletTransform(scale=2, dec=True).to(_dev())
    assert dec(torch.zeros(0, 3, 8, 8, device=_dev())).shape == (0, 48, 2, 2)
    with pytest.raises(ValueError):
        dec(torch.zeros(1, 3, 6, 8, device=_dev()))
    with pytest.raises(TypeError):
        dec(torch.zeros(1, 3, 8, 8, device=_dev(), dtype=torch.float16))
    # non-contiguous channel slice of a 6-channel loader tensor (models/restoration.py:84-85)
    x6 = torch.rand(2, 6, 16, 16, device=_dev())
    assert torch.equal(dec(x6[:, 3:]), dec(x6[:, 3:].contiguous()))
    with pytest.raises(_lib.WdmError):
        dwt4x4(torch.zeros(1, 3, 8, 8, device=_dev()), impl=_lib.WDM_WT_IMPL_TMA)


def test_autograd_adjoint():
    dec = WaveletTransform(scale=2, dec=True).to(_dev())
    x = torch.randn(1, 3, 8, 8, device=_dev(), requires_grad=True)
    gy = torch.randn(1, 48, 2, 2, device=_dev())
    (dec(x) * gy).sum().backward()
    assert (x.grad - iwt4x4(gy)).abs().max().item() == 0.0


@pytest.mark.parametrize("full_hi", [False, True])
def test_iwt_cat_equals_cat_then_iwt(full_hi):
    """restore() epilogue kernel: IWT(cat([lo, hi bands])) (+clamp) without the concatenated tensor, bit-identical to the
    two-step form and to the C oracle."""
    from wavedm_b200.wavelet import iwt4x4_cat
    DEV = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(8)
    lo = torch.randn(3, 3, 10, 14, generator=g)
    full = torch.randn(3, 48, 10, 14, generator=g)
    hi = full if full_hi else full[:, 3:].contiguous()
    ref = DO.iwt(torch.cat([lo, full[:, 3:]], 1).numpy(), flags=1)
    out = iwt4x4_cat(lo.to(DEV), hi.to(DEV), post_clamp=True)
    assert np.array_equal(out.cpu().numpy(), ref)
    out2 = iwt4x4(torch.cat([lo, full[:, 3:]], 1).to(DEV), post_clamp=True)
    assert torch.equal(out, out2)
    with pytest.raises(ValueError):
        iwt4x4_cat(lo.to(DEV), full[:, :10].contiguous().to(DEV))
