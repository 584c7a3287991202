"""A/B: the per-step (gather -> UNet) pair as a replayed CUDA graph at the bench's patch count (64) vs eager launches.
    python tools/graph_ab.py [--batch 64] [--steps 3]"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from wavedm_b200.harness import build_restorer  # noqa: E402
from wavedm_b200.sampler import DdimSampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = bench.make_cfg("bf16", dev)
    restorer = build_restorer(cfg, dev, sampling_timesteps=50, max_patches=64, seed=61)
    x, noise = bench.synth_inputs(a.batch, 0, device=dev)
    xo = restorer.diffusion.wavelet_dec(2 * x[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
    for gmax in (16, 64, 16, 64):
        DdimSampler.GRAPH_MAX_PATCHES = gmax
        restorer.restore_batch(x, r=16, noise=noise, x_other=xo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.steps):
            restorer.restore_batch(x, r=16, noise=noise, x_other=xo)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(f"GRAPH_MAX_PATCHES={gmax}: {ms:.2f} ms per restore_batch ({a.batch} images, HFRM bypassed) = {a.batch / ms * 1e3:.1f} images/s")


if __name__ == "__main__":
    main()
