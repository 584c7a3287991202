"""CPU: the oracles (oracle/dwt_oracle.c, oracle/unet_oracle.py) against the committed golden vectors that
oracle/make_golden.py produced from the reference's own modules (the reference ships no tests/fixtures)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import dwt_oracle as DO
from oracle import unet_oracle as O


def test_rec4_closed_form_matches_pickle_weights():
    g = golden("dwt_kat.npz")
    assert np.array_equal(DO.rec4(), g["rec4"])
    assert np.array_equal(O.haar_packet_matrix().reshape(16, 16), g["rec4"][:16].reshape(16, 16))


def test_haar_packet_weights_module_layout():
    from wavedm_b200.wavelet import haar_packet_weights
    g = golden("dwt_kat.npz")
    assert np.array_equal(haar_packet_weights(2).numpy(), g["rec4"])


@pytest.mark.parametrize("direct", [False, True])
def test_dwt_oracle_vs_reference_module(direct):
    g = golden("dwt_kat.npz")
    y = DO.dwt(g["x"], direct=direct)
    assert y.shape == g["dwt_x"].shape
    # tolerance: fp32 summation order of a 16-term dot product of O(1) values (conv backend order unknown)
    assert np.abs(y - g["dwt_x"]).max() <= 2e-6
    x = DO.iwt(g["y"], direct=direct)
    assert np.abs(x - g["iwt_y"]).max() <= 2e-6
    # integer-valued data: exact in every summation order -> bit-exact layout check
    assert np.array_equal(DO.dwt(g["xi"], direct=direct), g["dwt_xi"])
    assert np.array_equal(DO.iwt(g["dwt_xi"], direct=direct), g["iwt_dwt_xi"])
    assert np.array_equal(g["iwt_dwt_xi"], g["xi"])  # perfect reconstruction on the reference itself


def test_dwt_oracle_flags_and_numpy_twin():
    rng = np.random.default_rng(0)
    x = rng.random((2, 3, 16, 20), dtype=np.float32)
    assert np.array_equal(DO.dwt(x, flags=1), DO.dwt(2.0 * x - 1.0))
    y = DO.dwt(2.0 * x - 1.0)
    assert np.array_equal(DO.iwt(y, flags=1), np.clip((DO.iwt(y) + 1.0) / 2.0, 0.0, 1.0).astype(np.float32))
    assert np.abs(O.dwt_np(x) - DO.dwt(x)).max() <= 2e-6
    assert np.abs(O.iwt_np(y) - DO.iwt(y)).max() <= 4e-6
    # round trip: orthonormal basis
    assert np.abs(DO.iwt(DO.dwt(x)) - x).max() <= 1e-6
    # empty batch
    assert DO.dwt(np.zeros((0, 3, 8, 8), np.float32)).shape == (0, 48, 2, 2)


def test_unet_oracle_small_vs_reference_golden():
    g = golden("unet_small.npz")
    cfg = O.default_config(data__image_size=16, model__ch=128, model__ch_mult=[1, 2], model__num_res_blocks=1,
                           model__attn_resolutions=[8])
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(g["weight_sum"])) < 1e-9
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    with torch.no_grad():
        out = O.unet_forward(sd, cfg, x, t)
        out1 = O.unet_forward(sd, cfg, x, t[:1])
    # same torch build + same CPU kernels as the generating run -> equal up to thread-count dependent
    # reduction order
    assert (out - torch.from_numpy(g["out"])).abs().max() <= 2e-5
    assert (out1 - torch.from_numpy(g["out_t1"])).abs().max() <= 2e-5


def test_ddim_oracle_small_vs_reference_golden():
    g = golden("ddim_small.npz")
    cfg = O.default_config(data__image_size=16, model__ch=128, model__ch_mult=[1, 2], model__num_res_blocks=1,
                           model__attn_resolutions=[8])
    betas = O.beta_schedule(cfg)
    assert np.array_equal(betas.numpy(), g["betas"])
    assert np.array_equal(O.compute_alpha(betas, torch.arange(-1, 1000)).flatten().numpy(), g["alphas"])
    assert g["alphas"][0] == 1.0 and abs(g["alphas"][981] - 5.90375e-5) < 1e-9  # SURVEY A.2 samples
    h, w, p = g["x"].shape[2], g["x"].shape[3], int(g["p_size"])
    hl, wl = O.overlapping_grid_indices(h, w, p, int(g["r"]))
    corners = [(i, j) for i in hl for j in wl]
    assert np.array_equal(np.array(corners, np.int32), g["corners"])
    assert O.sampling_seq(1000, 6) == list(g["seq"]) and len(g["seq"]) == 7
    sd = O.init_state_dict(cfg, seed=61)
    with torch.no_grad():
        xs, x0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), torch.from_numpy(g["x"]),
                                            torch.from_numpy(g["x_cond"]), torch.from_numpy(g["x_other"]),
                                            list(g["seq"]), betas, corners, p)
    ref = torch.from_numpy(g["x0_preds"])
    scale = ref.abs().max()
    assert (torch.stack(x0p) - ref).abs().max() <= 1e-4 * scale
    assert (xs[-1] - torch.from_numpy(g["xs_last"])).abs().max() <= 1e-4 * scale
    assert (xs[1] - torch.from_numpy(g["xs_1"])).abs().max() <= 1e-4 * scale


def wiu_cfg():
    return O.default_config(data__image_size=16, data__patch_size=64, data__wavelet_in_unet=True, model__ch=128,
                            model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8],
                            model__use_other_channels=False, model__in_channels=93, model__out_ch=48)


def test_unet_oracle_wavelet_in_unet_vs_reference_golden():
    """data.wavelet_in_unet (models/unet.py:203-206,338-350,393-394): pixel-domain in / out, DWT and IWT inside forward;
    and the pixel-domain overlapping-patch DDIM run (restoration.py:171-172)."""
    g = golden("unet_wiu.npz")
    cfg = wiu_cfg()
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    assert len(sd) == int(g["nkeys"]) and "wavelet_dec.conv.weight" in sd
    # the torch restatement of the transform agrees with the C lifting form (summation order only)
    x6 = torch.from_numpy(g["x"])
    assert np.abs(O.dwt_torch(x6[:, :3]).numpy() - DO.dwt(g["x"][:, :3].copy())).max() <= 4e-6
    with torch.no_grad():
        out = O.unet_forward(sd, cfg, x6, torch.from_numpy(g["t"]))
    assert out.shape == (3, 3, 64, 64)
    assert (out - torch.from_numpy(g["out"])).abs().max() <= 2e-5 * max(1.0, float(np.abs(g["out"]).max()))
    corners = [tuple(int(v) for v in c) for c in g["corners"]]
    hl, wl = O.overlapping_grid_indices(80, 96, 64, 16)
    assert corners == [(i, j) for i in hl for j in wl]
    with torch.no_grad():
        xs, x0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), torch.from_numpy(g["x_noise"]),
                                            torch.from_numpy(g["x_cond"]), None, list(g["seq"]), O.beta_schedule(cfg),
                                            corners, 64)
    ref = torch.from_numpy(g["x0_preds"])
    assert (torch.stack(x0p) - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert (xs[-1] - torch.from_numpy(g["xs_last"])).abs().max() <= 1e-4 * ref.abs().max()


def test_grid_indices_cases():
    # 64^2 -> 1 corner; 128^2 -> 5x5; 120x180 -> 5x9
    assert O.overlapping_grid_indices(64, 64, 64, 16) == ([0], [0])
    hl, wl = O.overlapping_grid_indices(128, 128, 64, 16)
    assert len(hl) == 5 and len(wl) == 5
    hl, wl = O.overlapping_grid_indices(120, 180, 64, 16)
    assert (len(hl), len(wl)) == (5, 9) and hl[-1] == 56 and wl[-1] == 116


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference checkout not present")
def test_oracle_state_dict_is_bit_identical_to_reference_init():
    """Only in the build container: re-import the reference and compare (what make_golden.py asserts)."""
    import subprocess, sys
    code = (
        "import sys, types, os, torch\n"
        "for n in ('skimage','skimage.color'): sys.modules.setdefault(n, types.ModuleType(n))\n"
        "sys.modules['skimage'].color = sys.modules['skimage.color']\n"
        "repo = sys.argv[1]\n"
        "sys.path.insert(0, '/root/reference'); os.chdir('/root/reference')\n"
        "import models.unet as U\n"
        "sys.path.insert(0, repo)\n"
        "from oracle import unet_oracle as O\n"
        "cfg = O.default_config(data__image_size=16, model__ch=128, model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8])\n"
        "torch.manual_seed(7); net = U.DiffusionUNet(cfg)\n"
        "sd = O.init_state_dict(cfg, seed=7)\n"
        "ref = net.state_dict()\n"
        "assert sorted(sd) == sorted(ref)\n"
        "assert all(torch.equal(sd[k], ref[k]) for k in sd)\n"
        "x = torch.randn(2, 96, 16, 16); t = torch.tensor([3., 700.])\n"
        "with torch.no_grad(): assert torch.equal(net(x, t), O.unet_forward(sd, cfg, x, t))\n"
        "print('OK')\n")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code, repo], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_hfrm_mirror_vs_reference_golden():
    """wavedm_b200/hfrm.py (the plain-PyTorch mirror of models/arch.py:132-253 that the constructor strict-loads and
    restore() calls once per image): same state-dict keys / shapes as the reference module and the same
    output on a seeded input with deterministically filled parameters."""
    from wavedm_b200.hfrm import HFRM
    g = golden("hfrm.npz")
    net = HFRM(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=[2, 2, 2, 4], dec_blk_nums=[2, 2, 2, 2]).eval()
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert sum(v.numel() for v in sd.values()) == int(g["nparams"])
    gen = torch.Generator().manual_seed(int(g["param_seed"]))
    with torch.no_grad():
        for v in sd.values():
            v.copy_(torch.randn(v.shape, generator=gen) * 0.1)
        y = net._forward_autograd(torch.from_numpy(g["x"]))   # the differentiable definition (inference runs the CUDA engine)
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 1e-4 * ref.abs().max()
    # inference on a CPU module must fail loudly: there is no PyTorch fallback on the restore() path
    net.requires_grad_(False)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        net(torch.from_numpy(g["x"]))


def test_hfrm_oracle_vs_reference_golden():
    """oracle/hfrm_oracle.py (functional restatement of models/arch.py:132-253, the checker of the CUDA HFRM engine) against
    the output of the unmodified reference module (tests/golden/hfrm.npz, generated by oracle/make_golden.py)."""
    from oracle import hfrm_oracle as HO
    g = golden("hfrm.npz")
    shapes = {str(k): [int(v) for v in str(s).split(",")] for k, s in zip(g["keys"], g["shapes"])}
    sd = HO.fill_params(shapes, int(g["param_seed"]))
    y = HO.hfrm_forward(sd, torch.from_numpy(g["x"]))
    ref = torch.from_numpy(g["y"])
    assert (y - ref).abs().max() <= 2e-5 * ref.abs().max()


def test_use_window_oracle_vs_reference_golden():
    """data.use_window (models/unet.py:309-336,347-348,391-392): the oracle's re-layout + network against the output of the
    reference module built with the same config (tests/golden/unet_win.npz, oracle/make_golden.py --only-win)."""
    g = golden("unet_win.npz")
    cfg = O.default_config(data__image_size=16, data__use_window=True, data__window_size=2, model__ch=128,
                           model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8],
                           model__use_other_channels=False, model__in_channels=21, model__out_ch=12)
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    assert len(sd) == int(g["nkeys"])
    with torch.no_grad():
        out = O.unet_forward(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    ref = torch.from_numpy(g["out"])
    assert (out - ref).abs().max() <= 2e-5 * ref.abs().max()


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_optim_oracle_vs_torch_adam_and_reference_ema_loop(wd):
    """oracle/optim_oracle.py against torch.optim.Adam itself (what utils/optimize.py:7-8 constructs) and the EMA update loop
    of ddm_wavelet.py:48-53, six steps on CPU, one tensor without a gradient. Bit-identical."""
    from oracle.optim_oracle import AdamEmaOracle
    g = torch.Generator().manual_seed(5)
    params = [torch.nn.Parameter(torch.randn(n, generator=g)) for n in (1, 7, 1000, 8193)]
    shadow = [p.data.clone() for p in params]
    opt = torch.optim.Adam(params, lr=4e-5, weight_decay=wd, betas=(0.9, 0.999), amsgrad=False, eps=1e-8)
    orc = AdamEmaOracle([p.data for p in params], 4e-5, (0.9, 0.999), 1e-8, wd, mu=0.9999)
    for it in range(6):
        grads = [torch.randn(p.shape, generator=g) * 10.0 ** (it - 3) for p in params]
        grads[1] = None
        for p, gr in zip(params, grads):
            p.grad = gr
        opt.step()
        for i, p in enumerate(params):
            shadow[i] = (1. - 0.9999) * p.data + 0.9999 * shadow[i]
        orc.step(grads)
    for i, p in enumerate(params):
        assert torch.equal(p.data, orc.p[i]), i
        assert torch.equal(shadow[i], orc.shadow[i]), i
        if i != 1:
            assert torch.equal(opt.state[p]["exp_avg"], orc.m[i]) and torch.equal(opt.state[p]["exp_avg_sq"], orc.v[i])


def test_get_optimizer_on_cpu_is_torch_adam_and_fused_adam_refuses_cpu_parameters():
    """utils/optimize.py:5-14 factory: a model on the CPU gets torch.optim.Adam (the reference's object); the fused CUDA
    optimizer never steps CPU tensors through some other implementation."""
    from types import SimpleNamespace as NS
    from wavedm_b200 import optimize
    cfg = NS(optim=NS(optimizer="Adam", lr=4e-5, weight_decay=0.0, amsgrad=False, eps=1e-8))
    lin = torch.nn.Linear(3, 2)
    opt = optimize.get_optimizer(cfg, lin.parameters())
    assert type(opt) is torch.optim.Adam
    lin(torch.ones(1, 3)).sum().backward()
    with pytest.raises(RuntimeError):
        optimize.FusedAdam(lin.parameters(), lr=1e-3).step()
    with pytest.raises(NotImplementedError):
        optimize.FusedAdam(lin.parameters(), lr=1e-3, amsgrad=True)
