"""Engine part of __graft_entry__.smoke(): a tiny UNet forward + DDIM sampling run on cuda:0, checked against
the CPU oracle (oracle/unet_oracle.py). Test infrastructure use of the oracle only -- nothing here ships."""
import torch


def run(dev) -> None:
    from oracle import unet_oracle as O
    from wavedm_b200 import engine
    from wavedm_b200.sampler import DdimSampler
    cfg = O.default_config(data__image_size=16, model__ch=128, model__ch_mult=[1, 2], model__num_res_blocks=1,
                           model__attn_resolutions=[8])
    sd = O.init_state_dict(cfg, seed=61)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 96, 16, 16, generator=g)
    t = torch.tensor([500.0])
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t)
    for prec, tol in (("fp32", 1e-4), ("bf16", 5e-2)):
        eng = engine.UNetEngine(cfg, sd, dev, precision=prec)
        out = eng.forward(x.to(dev), t.to(dev)).cpu()
        rel = ((out - ref).norm() / ref.norm()).item()
        assert rel < tol, f"UNet {prec} mismatch vs oracle: rel {rel}"
    # 3-step DDIM over a 2x2 patch grid, fp32 engine vs oracle
    eng = engine.UNetEngine(cfg, sd, dev, precision="fp32")
    xc, xo, x0 = torch.randn(1, 48, 24, 24, generator=g), torch.randn(1, 45, 24, 24, generator=g), torch.randn(1, 3, 24, 24, generator=g)
    corners = [(0, 0), (0, 8), (8, 0), (8, 8)]
    seq = [0, 333, 666]
    betas = O.beta_schedule(cfg)
    with torch.no_grad():
        _, x0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), x0, xc, xo, seq, betas, corners, 16)
    _, x0h = DdimSampler(eng).sample(x0, xc, xo, seq, betas, corners, 16)
    err = (x0h[-1].cpu() - x0p[-1]).abs().max().item()
    assert err < 1e-3 * x0p[-1].abs().max().item(), f"DDIM mismatch vs oracle: {err}"

    # HFRM engine (restore()'s once-per-image refinement CNN) vs its oracle, and the batched PSNR statistics kernel
    from oracle import hfrm_oracle as HO
    from wavedm_b200 import metrics
    from wavedm_b200.hfrm import HfrmEngine
    hsd = {k: v * 0.5 for k, v in HO.fill_params(HO.default_shapes(), 3).items()}
    xi = torch.rand(2, 3, 32, 48, generator=g)
    href = HO.hfrm_forward(hsd, xi)
    for prec, tol in (("fp32", 1e-5), ("bf16", 3e-2)):
        y = HfrmEngine(hsd, dev, precision=prec).forward(xi.to(dev)).cpu()
        rel = ((y - href).norm() / href.norm()).item()
        assert rel < tol, f"HFRM {prec} mismatch vs oracle: rel {rel}"
    ps = metrics.psnr_batch(xi.to(dev), href.to(dev))[0]
    for b in range(2):
        assert abs(ps[b] - float(metrics.torchPSNR(xi[b:b + 1], href[b:b + 1]))) < 1e-3, "PSNR kernel mismatch"
