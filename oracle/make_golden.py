"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates the committed golden vectors under tests/golden/ by importing the UNMODIFIED reference from
/root/reference (read-only) in the build container and running its own modules on seeded inputs:

  * models/wavelet.py:6-50          WaveletTransform(scale=2) dec / rec      -> dwt_kat.npz
  * models/wavelet_weights_c2.pkl   rec4 weights                             -> dwt_kat.npz['rec4']
  * models/unet.py:196-395          DiffusionUNet (small + full config)      -> unet_small.npz, unet_full.npz
  * utils/sampling.py:10-13         compute_alpha                            -> ddim_small.npz['alphas']
  * models/ddm_wavelet.py:437-506   generalized_steps_overlapping            -> ddim_small.npz
  * models/ddm_wavelet.py:87-105    get_beta_schedule                        -> ddim_small.npz['betas']
  * models/unet.py:203-206,338-350  DiffusionUNet(wavelet_in_unet=True) + the pixel-domain sampler -> unet_wiu.npz
  * models/arch.py:132-253          HFRM (plain-PyTorch mirror in wavedm_b200/hfrm.py)                 -> hfrm.npz
  * models/restoration.py:63-168    the DWT -> sample -> x0_preds[-5] -> IWT -> clamp sandwich (config #1,
                                    with HFRM bypassed: x_other = HF bands of the DWT of the synthetic
                                    gt)                                      -> sandwich_full.npz
  * the same sandwich at 50 DDIM steps (BASELINE configs[1], two independent B = 1 runs) -> sandwich_s50.npz

It also checks, while it runs, that the oracle restatements (oracle/unet_oracle.py, oracle/dwt_oracle.c)
agree with the reference. /root/reference does not exist on the GPU box, so nothing at test time imports
this file; tests read only the .npz files. Run:  python oracle/make_golden.py
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from oracle import unet_oracle as O  # noqa: E402


def import_reference():
    """Shims of SURVEY.md 8(c): stub skimage (utils/metrics.py:4), cwd so that the cwd-relative
    './models/wavelet_weights_c2.pkl' (models/wavelet.py:7) resolves."""
    for name in ("skimage", "skimage.color"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].color = sys.modules["skimage.color"]
    os.chdir(REF)
    sys.path.insert(0, REF)
    import models.unet as ref_unet  # noqa
    import models.wavelet as ref_wavelet  # noqa
    import models.ddm_wavelet as ref_ddm  # noqa
    import utils.sampling as ref_sampling  # noqa
    import utils.metrics as ref_metrics  # noqa
    return ref_unet, ref_wavelet, ref_ddm, ref_sampling, ref_metrics


def small_cfg():
    return O.default_config(data__image_size=16, model__ch=128, model__ch_mult=[1, 2], model__num_res_blocks=1,
                            model__attn_resolutions=[8])


def wiu_cfg():
    """data.wavelet_in_unet: True needs out_ch 48, use_other_channels False and in_channels + pred_channels = 96
    (SURVEY.md A.5); small network, 16x16 sub-band patches = 64x64 pixel patches."""
    return O.default_config(data__image_size=16, data__patch_size=64, data__wavelet_in_unet=True, model__ch=128,
                            model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8],
                            model__use_other_channels=False, model__in_channels=93, model__out_ch=48)


def win_cfg():
    """data.use_window: True (models/unet.py:309-336,347-348,391-392) needs in_channels + pred_channels = 6 p^2 and
    out_ch = 3 p^2 ("needs its own config"); p = 2, 16x16 tiles of a 32x32 input."""
    return O.default_config(data__image_size=16, data__use_window=True, data__window_size=2, model__ch=128,
                            model__ch_mult=[1, 2], model__num_res_blocks=1, model__attn_resolutions=[8],
                            model__use_other_channels=False, model__in_channels=21, model__out_ch=12)


def golden_window(ref_unet):
    """The use_window variant of the reference module on a seeded input -> unet_win.npz"""
    cfg = win_cfg()
    torch.manual_seed(61)
    net = ref_unet.DiffusionUNet(cfg).eval()
    sd_ref = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd = O.init_state_dict(cfg, seed=61)
    assert sorted(sd.keys()) == sorted(sd_ref.keys()) and all(torch.equal(sd[k], sd_ref[k]) for k in sd)
    g = torch.Generator().manual_seed(67)
    xin = torch.randn(3, 6, 32, 32, generator=g)
    t = torch.tensor([12.0, 700.0, 333.0])
    with torch.no_grad():
        ref_out = net(xin, t)
        ora_out = O.unet_forward(sd, cfg, xin, t)
    assert ref_out.shape == (3, 3, 32, 32)
    assert torch.equal(ref_out, ora_out), float((ref_out - ora_out).abs().max())
    np.savez(os.path.join(OUT, "unet_win.npz"), x=xin.numpy(), t=t.numpy(), out=ref_out.numpy(), seed=61, nkeys=len(sd_ref))
    print("unet_win.npz ok; keys", len(sd_ref), "out range", float(ref_out.min()), float(ref_out.max()))


def golden_wavelet_in_unet(ref_unet, ref_ddm):
    """models/unet.py:203-206,338-350,393-394 + the pixel-domain sampler (restoration.py:171-172) -> unet_wiu.npz"""
    cfg = wiu_cfg()
    torch.manual_seed(61)
    net = ref_unet.DiffusionUNet(cfg).eval()
    sd_ref = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd = O.init_state_dict(cfg, seed=61)
    assert sorted(sd.keys()) == sorted(sd_ref.keys()), "state-dict keys differ (wavelet_in_unet)"
    assert all(torch.equal(sd[k], sd_ref[k]) for k in sd), "init_state_dict (wavelet_in_unet) is not bit-identical"
    g = torch.Generator().manual_seed(66)
    xin = torch.randn(3, 6, 64, 64, generator=g)
    t = torch.tensor([37.0, 980.0, 500.0])
    with torch.no_grad():
        ref_out = net(xin, t)
        ora_out = O.unet_forward(sd, cfg, xin, t)
    assert ref_out.shape == (3, 3, 64, 64)
    assert torch.equal(ref_out, ora_out), float((ref_out - ora_out).abs().max())
    # the sampler in the pixel domain: one 80x96 image, 64-pixel patches, r = 16 -> 2 x 3 corners, 4 DDIM steps
    stub = types.SimpleNamespace(config=cfg, num_timesteps=1000, device=torch.device("cpu"))
    betas = O.beta_schedule(cfg)
    xc = torch.randn(1, 3, 80, 96, generator=g)
    x0 = torch.randn(1, 3, 80, 96, generator=g)
    hl, wl = ref_ddm.DenoisingDiffusion_Wavelet.overlapping_grid_indices(stub, xc, output_size=64, r=16)
    corners = [(i, j) for i in hl for j in wl]
    seq = range(0, 1000, 250)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        xs, x0p = ref_ddm.DenoisingDiffusion_Wavelet.generalized_steps_overlapping(
            stub, x0, xc, seq, net, betas, eta=0., corners=corners, p_size=64, x_other=None, use_other=False)
    oxs, ox0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), x0, xc, None, list(seq), betas,
                                          corners, 64)
    assert all(torch.equal(a, b) for a, b in zip(xs, oxs)) and all(torch.equal(a, b) for a, b in zip(x0p, ox0p))
    np.savez(os.path.join(OUT, "unet_wiu.npz"), x=xin.numpy(), t=t.numpy(), out=ref_out.numpy(), seed=61, x_seed=66,
             x_cond=xc.numpy(), x_noise=x0.numpy(), corners=np.array(corners, np.int32), seq=np.array(list(seq)),
             xs_last=xs[-1].numpy(), x0_preds=torch.stack(x0p).numpy(), nkeys=len(sd_ref))
    print("unet_wiu.npz ok; keys", len(sd_ref), "corners", len(corners), "steps", len(x0p))


def fill_params_deterministically(module, seed):
    """The same pseudo-random parameter values for any module with the same state-dict key order and shapes (the
    default init leaves HFRM's beta / gamma at zero, which would hide most of the network from the comparison)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, v in module.state_dict().items():
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)


def golden_hfrm():
    """models/arch.py:132-253 (HFRM, the one-shot high-frequency refinement CNN) -> hfrm.npz: state-dict
    layout + the output of the reference module on a seeded input with deterministically filled parameters."""
    import models.arch as ref_arch
    kw = dict(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=[2, 2, 2, 4], dec_blk_nums=[2, 2, 2, 2])  # ddm_wavelet.py:137
    net = ref_arch.HFRM(**kw).eval()
    fill_params_deterministically(net, 71)
    x = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(72))
    with torch.no_grad():
        y = net(x)
    keys = list(net.state_dict().keys())
    shapes = [list(v.shape) for v in net.state_dict().values()]
    np.savez(os.path.join(OUT, "hfrm.npz"), x=x.numpy(), y=y.numpy(), keys=np.array(keys), nparams=sum(int(np.prod(s_)) for s_ in shapes),
             shapes=np.array([",".join(map(str, s_)) for s_ in shapes]), param_seed=71)
    print("hfrm.npz ok;", len(keys), "tensors", sum(int(np.prod(s_)) for s_ in shapes), "parameters; out range",
          float(y.min()), float(y.max()))


def golden_sandwich_s50(ref_unet, ref_wavelet, ref_ddm, ref_metrics):
    """BASELINE.json configs[1] (the north_star's own parity config: 256x256, **50** DDIM steps, fp32) -> sandwich_s50.npz.
    The reference sampler is batch-1 only (ddm_wavelet.py:486), so "batch 16" = independent B = 1 runs; two of
    them are stored (seeds 61 and 62 for the inputs, the SAME seed-61 weights) and the GPU test places them at different
    slots of a 16-image batch. Per image: x0_preds[-5] (restoration.py:108, the prediction made at t = 80), the full clamped
    256x256 output of restoration.py:111-135 with the HFRM bypassed (x_other = HF bands of DWT(gt), the `if 0:` branch at
    :99-100), torchPSNR against the synthetic gt, and the last x_t. The oracle restatement is run on image 0 and must be
    bit-equal."""
    cfgF = O.default_config()
    torch.manual_seed(61)
    netF = ref_unet.DiffusionUNet(cfgF).eval()
    sdF = O.init_state_dict(cfgF, seed=61)
    assert all(torch.equal(sdF[k], v) for k, v in netF.state_dict().items())
    dec = ref_wavelet.WaveletTransform(scale=2, dec=True)
    rec = ref_wavelet.WaveletTransform(scale=2, dec=False)
    betas = O.beta_schedule(cfgF)
    stubF = types.SimpleNamespace(config=cfgF, num_timesteps=1000, device=torch.device("cpu"))
    seq = range(0, 1000, 1000 // 50)
    lats, outs, psnrs, xlast, seeds = [], [], [], [], [61, 62]
    for n, seed in enumerate(seeds):
        g = torch.Generator().manual_seed(seed)
        ximg = torch.rand(1, 6, 256, 256, generator=g)
        noise = torch.randn(1, 3, 64, 64, generator=g)
        with torch.no_grad():
            xa = 2 * ximg - 1.0
            x_cond = dec(xa[:, :3].contiguous())
            x_other = dec(xa[:, 3:].contiguous())[:, 3:]
            with contextlib.redirect_stdout(io.StringIO()):
                xs, x0p = ref_ddm.DenoisingDiffusion_Wavelet.generalized_steps_overlapping(
                    stubF, noise, x_cond, seq, netF, betas, eta=0., corners=[(0, 0)], p_size=64, x_other=x_other,
                    use_other=True)
            assert len(x0p) == 50
            lat = x0p[-5]
            out = torch.clamp((rec(torch.cat([lat[:, :3], x_other], dim=1)) + 1.0) / 2.0, 0.0, 1.0)
            psnr = ref_metrics.torchPSNR(ximg[:, 3:], out)
            if n == 0:
                oxs, ox0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sdF, cfgF, a, tt), noise, x_cond, x_other,
                                                      list(seq), betas, [(0, 0)], 64)
                assert all(torch.equal(a, b) for a, b in zip(x0p, ox0p)), "oracle DDIM (50 steps) != reference"
        lats.append(lat.numpy()), outs.append(out.numpy()), psnrs.append(float(psnr)), xlast.append(xs[-1].numpy())
        print(f"  s50 image {n}: psnr {float(psnr):.6f} latent range [{float(lat.min()):.1f}, {float(lat.max()):.1f}] "
              f"saturated pixels {float(((out == 0) | (out == 1)).float().mean()):.3f}")
    np.savez(os.path.join(OUT, "sandwich_s50.npz"), seeds=np.array(seeds), steps=50, latent_m5=np.concatenate(lats),
             out=np.concatenate(outs), psnr=np.array(psnrs, np.float32), xs_last=np.concatenate(xlast))
    print("sandwich_s50.npz ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref_unet, ref_wavelet, ref_ddm, ref_sampling, ref_metrics = import_reference()
    if "--only-wiu" in sys.argv:   # regenerate just the wavelet_in_unet vectors
        golden_wavelet_in_unet(ref_unet, ref_ddm)
        return
    if "--only-win" in sys.argv:
        golden_window(ref_unet)
        return
    if "--only-hfrm" in sys.argv:
        golden_hfrm()
        return
    if "--only-s50" in sys.argv:
        golden_sandwich_s50(ref_unet, ref_wavelet, ref_ddm, ref_metrics)
        return

    # ---------------------------------------------------------------- DWT / IWT
    import pickle
    with open(os.path.join(REF, "models", "wavelet_weights_c2.pkl"), "rb") as f:
        u = pickle._Unpickler(f)
        u.encoding = "latin1"
        rec4 = u.load()["rec4"].astype(np.float32)
    dec = ref_wavelet.WaveletTransform(scale=2, dec=True)
    rec = ref_wavelet.WaveletTransform(scale=2, dec=False)
    g = torch.Generator().manual_seed(61)
    x = torch.randn(2, 3, 8, 12, generator=g)
    y = torch.randn(2, 48, 3, 5, generator=g)
    # integer-valued input: every summation order gives the same exactly representable result
    xi = torch.randint(-64, 64, (1, 3, 12, 8), generator=g).float()
    with torch.no_grad():
        ref_y, ref_x, ref_yi = dec(x), rec(y), dec(xi)
        ref_xi = rec(ref_yi)
    assert np.array_equal(O.haar_packet_matrix().reshape(16, 16), rec4[:16].reshape(16, 16))
    assert np.allclose(O.dwt_np(x.numpy()), ref_y.numpy(), atol=2e-6)
    assert np.allclose(O.iwt_np(y.numpy()), ref_x.numpy(), atol=2e-6)
    assert np.array_equal(O.dwt_np(xi.numpy()), ref_yi.numpy())
    np.savez(os.path.join(OUT, "dwt_kat.npz"), rec4=rec4, x=x.numpy(), dwt_x=ref_y.numpy(), y=y.numpy(),
             iwt_y=ref_x.numpy(), xi=xi.numpy(), dwt_xi=ref_yi.numpy(), iwt_dwt_xi=ref_xi.numpy())
    print("dwt_kat.npz ok")

    # ---------------------------------------------------------------- UNet, small config
    cfg = small_cfg()
    torch.manual_seed(61)
    net = ref_unet.DiffusionUNet(cfg).eval()
    sd_ref = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd = O.init_state_dict(cfg, seed=61)
    assert sorted(sd.keys()) == sorted(sd_ref.keys()), "state-dict keys differ"
    assert all(torch.equal(sd[k], sd_ref[k]) for k in sd), "init_state_dict is not bit-identical to the reference"
    g = torch.Generator().manual_seed(62)
    xin = torch.randn(3, 96, 16, 16, generator=g)
    t = torch.tensor([37.0, 980.0, 500.0])
    with torch.no_grad():
        ref_out = net(xin, t)
        ora_out = O.unet_forward(sd, cfg, xin, t)
        ref_out_t1 = net(xin, t[:1])  # the sampler's broadcast-t form (n=1 against P patches)
    assert torch.equal(ref_out, ora_out), float((ref_out - ora_out).abs().max())
    wsum = float(sum(v.double().sum() for v in sd.values()))
    np.savez(os.path.join(OUT, "unet_small.npz"), x=xin.numpy(), t=t.numpy(), out=ref_out.numpy(),
             out_t1=ref_out_t1.numpy(), weight_sum=np.float64(wsum), seed=61, x_seed=62)
    print("unet_small.npz ok; weight_sum", wsum)

    # ---------------------------------------------------------------- DDIM sampler, small config
    stub = types.SimpleNamespace(config=cfg, num_timesteps=1000, device=torch.device("cpu"))
    betas_np = ref_ddm.get_beta_schedule(beta_schedule="linear", beta_start=1e-4, beta_end=0.02,
                                         num_diffusion_timesteps=1000)
    betas = torch.from_numpy(betas_np).float()
    assert torch.equal(betas, O.beta_schedule(cfg))
    alphas = ref_sampling.compute_alpha(betas, torch.arange(-1, 1000)).flatten()
    assert torch.equal(alphas, O.compute_alpha(betas, torch.arange(-1, 1000)).flatten())
    g = torch.Generator().manual_seed(63)
    B, h, w, p = 1, 24, 40, 16
    xc = torch.randn(B, 48, h, w, generator=g)
    xo = torch.randn(B, 45, h, w, generator=g)
    x0 = torch.randn(B, 3, h, w, generator=g)
    h_list, w_list = ref_ddm.DenoisingDiffusion_Wavelet.overlapping_grid_indices(stub, xc, output_size=p, r=8)
    assert (h_list, w_list) == O.overlapping_grid_indices(h, w, p, 8)
    corners = [(i, j) for i in h_list for j in w_list]
    seq = range(0, 1000, 1000 // 6)  # 6 -> skip 166 -> 7 entries: covers len(seq) = S+1 (SURVEY hard parts)
    with contextlib.redirect_stdout(io.StringIO()):
        xs, x0p = ref_ddm.DenoisingDiffusion_Wavelet.generalized_steps_overlapping(
            stub, x0, xc, seq, net, betas, eta=0., corners=corners, p_size=p, x_other=xo, use_other=True)
    oxs, ox0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), x0, xc, xo, list(seq),
                                          betas, corners, p)
    for a, b in zip(xs, oxs):
        assert torch.equal(a, b)
    for a, b in zip(x0p, ox0p):
        assert torch.equal(a, b)
    np.savez(os.path.join(OUT, "ddim_small.npz"), betas=betas.numpy(), alphas=alphas.numpy(), x_cond=xc.numpy(),
             x_other=xo.numpy(), x=x0.numpy(), corners=np.array(corners, np.int32), seq=np.array(list(seq)),
             xs_last=xs[-1].numpy(), x0_preds=torch.stack(x0p).numpy(), xs_1=xs[1].numpy(), p_size=p, r=8)
    print("ddim_small.npz ok; corners", len(corners), "steps", len(x0p))

    # ---------------------------------------------------------------- UNet, full config, one patch
    cfgF = O.default_config()
    torch.manual_seed(61)
    netF = ref_unet.DiffusionUNet(cfgF).eval()
    sdF = O.init_state_dict(cfgF, seed=61)
    assert all(torch.equal(sdF[k], v) for k, v in netF.state_dict().items())
    nparams = sum(v.numel() for v in sdF.values())
    assert nparams == 156492675, nparams
    g = torch.Generator().manual_seed(64)
    xin = torch.randn(2, 96, 64, 64, generator=g)
    t = torch.tensor([500.0])
    with torch.no_grad():
        ref_out = netF(xin, t)
        ora_out = O.unet_forward(sdF, cfgF, xin, t)
    assert torch.equal(ref_out, ora_out)
    np.savez(os.path.join(OUT, "unet_full.npz"), t=t.numpy(), out=ref_out.numpy(), seed=61, x_seed=64,
             weight_sum=np.float64(sum(v.double().sum() for v in sdF.values())))
    print("unet_full.npz ok")

    # ---------------------------------------------------------------- the restore() sandwich, config #1
    # restoration.py:73-135 with the HFRM bypassed (x_other := HF bands of DWT(gt), the `if 0:` branch at
    # :99-100), 10 DDIM steps, one 256x256 image -> one 64x64 patch.
    g = torch.Generator().manual_seed(61)
    ximg = torch.rand(1, 6, 256, 256, generator=g)
    noise = torch.randn(1, 3, 64, 64, generator=g)
    stubF = types.SimpleNamespace(config=cfgF, num_timesteps=1000, device=torch.device("cpu"))
    with torch.no_grad():
        xa = 2 * ximg - 1.0
        x_cond = dec(xa[:, :3].contiguous())
        x_gt = dec(xa[:, 3:].contiguous())
        x_other = x_gt[:, 3:]
        seq = range(0, 1000, 1000 // 10)
        with contextlib.redirect_stdout(io.StringIO()):
            xs, x0p = ref_ddm.DenoisingDiffusion_Wavelet.generalized_steps_overlapping(
                stubF, noise, x_cond, seq, netF, betas, eta=0., corners=[(0, 0)], p_size=64, x_other=x_other,
                use_other=True)
        lat = x0p[-5]
        out = torch.clamp((rec(torch.cat([lat[:, :3], x_other], dim=1)) + 1.0) / 2.0, 0.0, 1.0)
        psnr = ref_metrics.torchPSNR(ximg[:, 3:], out)
    np.savez(os.path.join(OUT, "sandwich_full.npz"), seed=61, steps=10, latent_m5=lat.numpy(),
             out_crop=out[:, :, 96:160, 96:160].numpy(), out_mean=np.float64(out.double().mean()),
             psnr=np.float32(psnr), xs_last=xs[-1].numpy())
    # ---------------------------------------------------------------- overlapping-patch sampling at 512x512 (config #5 shape)
    # wavelet domain 128x128, 64x64 patches, grid_r = 16 -> 5x5 = 25 patches per image, 5 DDIM steps (of the 100-step
    # schedule's spacing this is coarser, the arithmetic is the same).
    g = torch.Generator().manual_seed(65)
    xc5 = torch.randn(1, 48, 128, 128, generator=g)
    xo5 = torch.randn(1, 45, 128, 128, generator=g)
    xn5 = torch.randn(1, 3, 128, 128, generator=g)
    hl, wl = ref_ddm.DenoisingDiffusion_Wavelet.overlapping_grid_indices(stubF, xc5, output_size=64, r=16)
    corners5 = [(i, j) for i in hl for j in wl]
    assert len(corners5) == 25
    seq5 = range(0, 1000, 200)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        xs5, x0p5 = ref_ddm.DenoisingDiffusion_Wavelet.generalized_steps_overlapping(
            stubF, xn5, xc5, seq5, netF, betas, eta=0., corners=corners5, p_size=64, x_other=xo5, use_other=True)
    np.savez(os.path.join(OUT, "patched_full.npz"), seed=65, x0_first=x0p5[0].numpy(), x0_last=x0p5[-1].numpy(),
             xs_last=xs5[-1].numpy(), ncorners=len(corners5), steps=len(x0p5))
    print("patched_full.npz ok; latent range", float(x0p5[-1].min()), float(x0p5[-1].max()))

    print("sandwich_full.npz ok; psnr", float(psnr), "latent range", float(lat.min()), float(lat.max()))
    golden_wavelet_in_unet(ref_unet, ref_ddm)
    golden_hfrm()
    golden_window(ref_unet)
    golden_sandwich_s50(ref_unet, ref_wavelet, ref_ddm, ref_metrics)


if __name__ == "__main__":
    main()
