"""Times the HFRM engine (csrc/wdm_hfrm.cu) alone: CUDA events around K calls, per precision / batch.
    python tools/bench_hfrm.py [--batch 64] [--size 256]
Algorithmic HBM bytes of the fused schedule: 14 C bytes-per-element units per pixel and ResidualBlock."""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavedm_b200.hfrm import HFRM, HfrmEngine  # noqa: E402

ARCH = dict(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=(2, 2, 2, 4), dec_blk_nums=(2, 2, 2, 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precisions", default="bf16,fp32")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(61)
    net = HFRM(**ARCH)
    with torch.no_grad():
        for p in net.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.1)
    sd = net.state_dict()
    for prec in a.precisions.split(","):
        B = a.batch if prec == "bf16" else max(1, a.batch // 4)
        x = torch.rand(B, 3, a.size, a.size, device=dev)
        eng = HfrmEngine(sd, dev, precision=prec, **ARCH)
        eng.forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.iters):
            eng.forward(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        es = 2 if prec == "bf16" else 4
        px = B * a.size * a.size
        # blocks per level (enc + dec, mid at the deepest): pixels / 4^l, C = 32 * 2^l -> px*C halves per level
        blocks = [4, 4, 4, 6, 6]
        alg = sum(nb * 14 * (px * 32 / 2 ** l) * es for l, nb in enumerate(blocks))
        flops = sum(nb * 2 * 6 * (px / 4 ** l) * (32 * 2 ** l) ** 2 for l, nb in enumerate(blocks))
        print(f"HFRM {prec} B={B} {a.size}x{a.size}: {ms:.3f} ms/call = {ms / B * 1e3:.1f} us/image; schedule bytes {alg / 1e9:.2f} GB "
              f"-> {alg / ms / 1e6:.0f} GB/s; 1x1-conv flops {flops / 1e9:.1f} GF -> {flops / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
