"""GPU timeline of one DDIM step of the sampler at the bench shape (64 images x 1 patch): CUDA events between the gather,
the UNet call, the DDIM update and the RNG mirror. Stream-ordered events measure device time incl. any idle gaps."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from wavedm_b200.harness import build_restorer  # noqa: E402
from wavedm_b200.sampler import alpha_table, make_patch_table  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = bench.make_cfg("bf16", dev)
    restorer = build_restorer(cfg, dev, sampling_timesteps=50, max_patches=64, seed=61)
    eng = restorer.diffusion.model.module.engine()
    B = 64
    x, noise = bench.synth_inputs(B, 0, device=dev)
    x_cond = restorer.diffusion.wavelet_dec(2 * x[:, :3].contiguous() - 1.0).contiguous()
    xo = restorer.diffusion.wavelet_dec(2 * x[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
    patches, first = make_patch_table(B, [(0, 0)], dev)
    xin = torch.empty((B, eng.R, eng.R, eng.cin_pad), dtype=eng.dtype, device=dev)
    eps = torch.empty((B, 3, 64, 64), device=dev)
    xt = noise.clone()
    x0 = torch.empty_like(xt)
    xn = torch.empty_like(xt)
    t = torch.tensor([500.0], device=dev)
    alphas = alpha_table(restorer.diffusion.betas)
    at, an = float(alphas[501]), float(alphas[481])
    S = 30
    ev = [[torch.cuda.Event(True) for _ in range(5)] for _ in range(S)]
    for k in range(S + 3):
        e = ev[max(0, k - 3)]
        e[0].record()
        if k == 0:
            eng.gather([x_cond, xt, xo], patches, out=xin)              # step 1: all 96 channels (54-59 us)
        else:
            eng.gather_update(xt, x_cond.shape[1], patches, xin)        # steps 2..S: the 3 channels of x_t, as the sampler does
        e[1].record()
        eng.forward_nhwc(xin, t, out=eps)
        e[2].record()
        eng.ddim_step(eps, patches, first, xt, x0, xn, at, an)
        e[3].record()
        torch.randn_like(xt)
        e[4].record()
    torch.cuda.synchronize()
    names = ["gather_upd", "unet", "ddim_step", "randn_like"]
    tot = 0.0
    for i, n in enumerate(names):
        ms = sum(e[i].elapsed_time(e[i + 1]) for e in ev[3:]) / (S - 3)
        tot += ms
        print(f"{n:12s} {ms * 1e3:9.1f} us")
    step = sum(ev[k][0].elapsed_time(ev[k + 1][0]) for k in range(3, S - 1)) / (S - 4)
    print(f"sum {tot * 1e3:.1f} us; step-to-step {step * 1e3:.1f} us")


if __name__ == "__main__":
    main()


def phases():
    """restore_batch vs its sampling loop: where the time outside the 50 DDIM steps goes."""
    from wavedm_b200.sampler import DdimSampler
    dev = torch.device("cuda", 0)
    cfg = bench.make_cfg("bf16", dev)
    restorer = build_restorer(cfg, dev, sampling_timesteps=50, max_patches=64, seed=61)
    eng = restorer.diffusion.model.module.engine()
    B = 64
    x, noise = bench.synth_inputs(B, 0, device=dev)
    x_cond = restorer.diffusion.wavelet_dec(2 * x[:, :3].contiguous() - 1.0).contiguous()
    xo = restorer.diffusion.wavelet_dec(2 * x[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
    seq = range(0, 1000, 20)
    smp = DdimSampler(eng, max_patches=64)

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_s = timed(lambda: smp.sample(noise, x_cond, xo, seq, restorer.diffusion.betas, [(0, 0)], 64, keep_last=5))
    t_r = timed(lambda: restorer.restore_batch(x, r=16, noise=noise, x_other=xo))
    t_h = timed(lambda: restorer.restore_batch(x, r=16, noise=noise))
    print(f"sampler.sample (50 steps): {t_s:.2f} ms; restore_batch(x_other given): {t_r:.2f} ms; restore_batch with HFRM: {t_h:.2f} ms")


if __name__ == "__main__" and len(sys.argv) > 1:
    phases()
