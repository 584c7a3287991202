"""HFRM engine (csrc/wdm_hfrm.cu) against the oracle (oracle/hfrm_oracle.py, pinned to the reference module by
tests/golden/hfrm.npz) and against that golden itself. fp32 engine = parity mode (FFMA), bf16 engine = throughput mode."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

ARCH = dict(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=(2, 2, 2, 4), dec_blk_nums=(2, 2, 2, 2))  # ddm_wavelet.py:137


def _shapes(arch):
    from wavedm_b200.hfrm import HFRM
    return {k: list(v.shape) for k, v in HFRM(**arch).state_dict().items()}


def _engine(sd, precision, arch=ARCH):
    from wavedm_b200.hfrm import HfrmEngine
    return HfrmEngine(sd, torch.device("cuda", 0), precision=precision, **arch)


def _rel(a, b):
    return float((a - b).norm() / b.norm()), float((a - b).abs().max())


def test_hfrm_fp32_and_bf16_vs_reference_golden():
    """The reference module's own output (golden, 2 x 3 x 32 x 48: non-square, 2 x 3 pixels at the deepest level)."""
    from oracle import hfrm_oracle as HO
    g = golden("hfrm.npz")
    shapes = {str(k): [int(v) for v in str(s).split(",")] for k, s in zip(g["keys"], g["shapes"])}
    sd = HO.fill_params(shapes, int(g["param_seed"]))
    x = torch.from_numpy(g["x"])
    ref = torch.from_numpy(g["y"])
    y32 = _engine(sd, "fp32").forward(x.cuda()).cpu()
    r32, m32 = _rel(y32, ref)
    y16 = _engine(sd, "bf16").forward(x.cuda()).cpu()
    r16, m16 = _rel(y16, ref)
    print(f"HFRM vs reference golden: fp32 rel-L2 {r32:.3e} max|d| {m32:.3e}; bf16 rel-L2 {r16:.3e} max|d| {m16:.3e}")
    assert m32 <= 2e-5 * float(ref.abs().max())
    assert r16 <= 2e-2


@pytest.mark.parametrize("B,H,W", [(1, 16, 16), (3, 64, 96), (2, 256, 256), (1, 480, 720)])
def test_hfrm_vs_oracle_shapes(B, H, W):
    """Seeded parameters / inputs at sizes the CPU oracle finishes in seconds: the smallest legal image (one pixel at the
    deepest level), a ragged batch of non-square images, the BASELINE 256 x 256 size, and the real RainDrop geometry
    (480 x 720: 30 x 45 pixels at the deepest level, row counts that are no multiple of the 128-row tensor-core tile)."""
    from oracle import hfrm_oracle as HO
    sd = HO.fill_params(_shapes(ARCH), 71)
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(5))
    ref = HO.hfrm_forward(sd, x, ARCH["mid_blk_num"], ARCH["enc_blk_nums"], ARCH["dec_blk_nums"])
    y32 = _engine(sd, "fp32").forward(x.cuda()).cpu()
    y16 = _engine(sd, "bf16").forward(x.cuda()).cpu()
    r32, m32 = _rel(y32, ref)
    r16, m16 = _rel(y16, ref)
    print(f"HFRM {B}x{H}x{W}: fp32 rel-L2 {r32:.3e} max|d| {m32:.3e}; bf16 rel-L2 {r16:.3e} max|d| {m16:.3e}")
    assert m32 <= 5e-5 * float(ref.abs().max())
    assert r16 <= 2e-2


def test_hfrm_residual_only_vs_oracle():
    """The part of the network the input residual does not mask: y - x (the refinement itself) against the oracle, with
    larger parameters so every branch (channel attention, both gates, beta / gamma) carries signal."""
    from oracle import hfrm_oracle as HO
    sd = {k: v * 3.0 for k, v in HO.fill_params(_shapes(ARCH), 9).items()}
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(6))
    ref = HO.hfrm_forward(sd, x) - x
    y32 = _engine(sd, "fp32").forward(x.cuda()).cpu() - x
    y16 = _engine(sd, "bf16").forward(x.cuda()).cpu() - x
    r32, _ = _rel(y32, ref)
    r16, _ = _rel(y16, ref)
    print(f"HFRM refinement only: fp32 rel-L2 {r32:.3e}, bf16 rel-L2 {r16:.3e}; |ref| max {float(ref.abs().max()):.3f}")
    assert r32 <= 1e-4
    assert r16 <= 5e-2


@pytest.mark.parametrize("arch", [
    dict(in_channel=3, dim=32, mid_blk_num=1, enc_blk_nums=(1, 1), dec_blk_nums=(1, 1)),
    dict(in_channel=3, dim=64, mid_blk_num=2, enc_blk_nums=(0, 2, 1), dec_blk_nums=(1, 0, 2)),
])
def test_hfrm_other_architectures(arch):
    """Constructor arguments other than the raindrop ones (arch.py:208), incl. levels without blocks."""
    from oracle import hfrm_oracle as HO
    sd = HO.fill_params(_shapes(arch), 3)
    x = torch.rand(2, 3, 32, 40, generator=torch.Generator().manual_seed(7))
    ref = HO.hfrm_forward(sd, x, arch["mid_blk_num"], arch["enc_blk_nums"], arch["dec_blk_nums"])
    y32 = _engine(sd, "fp32", arch).forward(x.cuda()).cpu()
    assert float((y32 - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
    y16 = _engine(sd, "bf16", arch).forward(x.cuda()).cpu()
    assert _rel(y16, ref)[0] <= 2e-2


def test_hfrm_bad_shapes_and_chunked_batches():
    from oracle import hfrm_oracle as HO
    from wavedm_b200 import _lib
    sd = HO.fill_params(_shapes(ARCH), 71)
    eng = _engine(sd, "fp32")
    with pytest.raises(_lib.WdmError):
        eng.forward(torch.rand(1, 3, 40, 64).cuda())      # 40 is not a multiple of 16
    with pytest.raises(ValueError):
        eng.forward(torch.rand(1, 4, 32, 32).cuda())
    x = torch.rand(5, 3, 32, 32, generator=torch.Generator().manual_seed(8)).cuda()
    assert torch.equal(eng.forward(x), eng.forward(x, max_batch=2))   # images are independent: chunking changes nothing


def test_hfrm_module_dispatch_runs_the_engine():
    """HFRM.forward under no_grad on a CUDA module = the engine (kernel launches of this library, no PyTorch fallback)."""
    from oracle import hfrm_oracle as HO
    from wavedm_b200 import _lib
    from wavedm_b200.hfrm import HFRM
    net = HFRM(**ARCH).eval()
    sd = HO.fill_params(_shapes(ARCH), 71)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().requires_grad_(False)
    net.engine_precision = "fp32"
    x = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(9))
    lib = _lib.load()
    n0 = lib.wdm_launch_counter()
    with torch.no_grad():
        y = net(x.cuda()).cpu()
    assert lib.wdm_launch_counter() - n0 >= 24 * 6
    ref = HO.hfrm_forward(sd, x)
    assert float((y - ref).abs().max()) <= 5e-5 * float(ref.abs().max())
    # parameters written in place (optimizer / load_state_dict) re-pack the engine
    with torch.no_grad():
        net.conv_out.bias.add_(0.5)
        y2 = net(x.cuda()).cpu()
    assert float((y2 - y).abs().max()) > 0.4
