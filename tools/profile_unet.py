"""Runs N UNet engine forwards at P patches (random weights/input) -- the target for ncu launch lists / captures.
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv \
        python tools/profile_unet.py --patches 64 --iters 2
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavedm_b200 import engine  # noqa: E402
from wavedm_b200.configs import default_config  # noqa: E402
from wavedm_b200.unet import DiffusionUNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = default_config()
    torch.manual_seed(61)
    net = DiffusionUNet(cfg)
    eng = engine.UNetEngine(cfg, net.state_dict(), dev, precision=a.precision, max_patches=a.patches)
    del net
    x = torch.randn(a.patches, 64, 64, eng.cin_pad, device=dev).to(eng.dtype)
    t = torch.tensor([500.0], device=dev)
    out = torch.empty(a.patches, 3, 64, 64, device=dev)
    eng.forward_nhwc(x, t, out=out)
    torch.cuda.synchronize()
    if a.time:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.iters):
            eng.forward_nhwc(x, t, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"P={a.patches} {a.precision}: {ms:.3f} ms/forward, {79.945e9 * a.patches / ms / 1e9:.1f} TFLOP/s algorithmic")
        eng.profile(True)
        eng.forward_nhwc(x, t, out=out)
        print("profile (tc_ms, tc_flops, tc_n, simt_ms, simt_flops, simt_n):", eng.profile_read())
        eng.profile(False)
    else:
        for _ in range(a.iters):
            eng.forward_nhwc(x, t, out=out)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
