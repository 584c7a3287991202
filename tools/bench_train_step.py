"""Training step of configs/raindrop_wavelet.yml (batch_size 1 x patch_n 8 = 8 latent patches of 64 x 64, fp32 autograd,
ddm_wavelet.py:200-270) with the parameter update done the reference's way (torch.optim.Adam + the EMAHelper.update loop,
ddm_wavelet.py:48-53) and through the fused kernel (csrc/wdm_optim.cu: FusedAdam with the EMA attached): CUDA events around
the update alone and around the whole step; roofline of the fused launch = 36 B per parameter / its duration.
    python tools/bench_train_step.py [--steps 10] [--patches 8] [--json out.json]
The forward / backward here is PyTorch autograd (cuDNN / cuBLAS), not this repo's kernels: SURVEY 8(f)-3 status in DESIGN.md 8."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import unet_oracle as O  # noqa: E402  (config helper only)
from wavedm_b200 import _lib  # noqa: E402
from wavedm_b200.ddm_wavelet import EMAHelper, get_beta_schedule, noise_estimation_loss  # noqa: E402
from wavedm_b200.optimize import FusedAdam  # noqa: E402
from wavedm_b200.unet import DiffusionUNet  # noqa: E402


class ReferenceEma:
    """ddm_wavelet.py:35-53 as the reference runs it: three eager kernels per parameter tensor, new shadow tensors."""

    def __init__(self, module, mu=0.9999):
        self.mu = mu
        self.shadow = {n: p.data.clone() for n, p in module.named_parameters() if p.requires_grad}

    def update(self, module):
        for n, p in module.named_parameters():
            if p.requires_grad:
                self.shadow[n].data = (1. - self.mu) * p.data + self.mu * self.shadow[n].data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--patches", type=int, default=8)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = O.default_config()
    cfg.device = dev
    betas = torch.from_numpy(get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000)).float().to(dev)
    res = {"patches": a.patches, "steps": a.steps}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    for arm in ("reference", "fused"):
        torch.manual_seed(7)
        net = DiffusionUNet(cfg).to(dev).train()
        n_params = sum(p.numel() for p in net.parameters())
        if arm == "fused":
            ema = EMAHelper()
            ema.register(net)
            opt = FusedAdam(net.parameters(), lr=4e-5, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8)
            opt.attach_ema(ema, net)
        else:
            ema = ReferenceEma(net)
            opt = torch.optim.Adam(net.parameters(), lr=4e-5, weight_decay=0.0, betas=(0.9, 0.999), amsgrad=False, eps=1e-8)
        g = torch.Generator(device=dev).manual_seed(8)
        x0 = torch.randn(a.patches, 96, 64, 64, device=dev, generator=g)     # [x_cond 48 | x_gt 3 (+45 other)] wavelet bands
        ev = [torch.cuda.Event(True) for _ in range(4)]
        upd_ms = step_ms = 0.0
        lib = _lib.load()
        launches = 0
        for it in range(a.warmup + a.steps):
            e = torch.randn(a.patches, 3, 64, 64, device=dev, generator=g)
            t = torch.randint(0, 1000, (a.patches,), device=dev, generator=g)
            ev[0].record()
            loss, _, _, _ = noise_estimation_loss(net, x0, t, e, betas, inp_channels=48, pred_channels=3, use_other_channels=True)
            opt.zero_grad()
            loss.backward()
            ev[1].record()
            n0 = lib.wdm_launch_counter()
            opt.step()
            ema.update(net)
            launches = lib.wdm_launch_counter() - n0
            ev[2].record()
            torch.cuda.synchronize()
            if it >= a.warmup:
                upd_ms += ev[1].elapsed_time(ev[2])
                step_ms += ev[0].elapsed_time(ev[2])
        upd_ms /= a.steps
        step_ms /= a.steps
        if arm == "fused":
            r_tab = sum(b["table"].uploads for pl in opt._plans.values() for b in pl["buckets"])
        r = {"update_ms": round(upd_ms, 4), "step_ms": round(step_ms, 3), "params": n_params, "loss": float(loss.detach()),
             "kernel_launches_of_this_library_in_update": int(launches)}
        if arm == "fused":
            # the launch alone: back-to-back updates on the last gradients (GPU-bound: the host side is ~0.4 ms per call)
            ev[0].record()
            for _ in range(10):
                opt.step()
                ema.update(net)
            ev[1].record()
            torch.cuda.synchronize()
            k_ms = ev[0].elapsed_time(ev[1]) / 10
            r["kernel_ms"] = round(k_ms, 4)
            r["pointer_table_uploads"] = int(r_tab)   # 1 = the gradient tensors kept their addresses over all steps
            upd_ms_in_step, upd_ms = upd_ms, k_ms
            alg = 36.0 * n_params
            peak = float(peaks.get("hbm_gbs", 0) or 0)
            r["roofline"] = {"bound": "hbm", "algorithmic_bytes": alg, "achieved": round(alg / upd_ms / 1e6, 1), "unit": "GB/s",
                             "peak": peak or None, "frac": round(alg / upd_ms / 1e6 / peak, 3) if peak else None,
                             "note": "36 B per parameter / kernel_ms (back-to-back launches); update_ms is the same update inside the "
                                     "launch-bound training step, host side of step() included"}
            upd_ms = upd_ms_in_step
        res[arm] = r
        print(f"{arm:9s}: update (optimizer.step + ema.update) {upd_ms:.3f} ms, whole training step {step_ms:.2f} ms, "
              f"{n_params / 1e6:.1f} M parameters, loss {float(loss):.4f}", flush=True)
        del net, opt, ema
        torch.cuda.empty_cache()
    res["update_speedup"] = round(res["reference"]["update_ms"] / res["fused"]["update_ms"], 2)
    res["step_speedup"] = round(res["reference"]["step_ms"] / res["fused"]["step_ms"], 3)
    print(json.dumps(res))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
