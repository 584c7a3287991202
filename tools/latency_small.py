"""Single-image latency (BASELINE.json configs[0] shape: batch 1, 256x256, 10 DDIM steps, one 64x64 patch per step):
the per-step (gather -> UNet) pair launched eagerly (~190 launches per call, host-bound) vs replayed as a CUDA graph."""
import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavedm_b200 import engine  # noqa: E402
from wavedm_b200.configs import default_config  # noqa: E402
from wavedm_b200.sampler import DdimSampler  # noqa: E402
from wavedm_b200.unet import DiffusionUNet  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = default_config()
    torch.manual_seed(61)
    net = DiffusionUNet(cfg)
    eng = engine.UNetEngine(cfg, net.state_dict(), dev, precision="bf16")
    del net
    betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64).float()
    for B, hw, steps in ((1, 64, 10), (1, 64, 50), (4, 64, 10), (1, 128, 10)):
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, 3, hw, hw, generator=g).to(dev)
        xc = torch.randn(B, 48, hw, hw, generator=g).to(dev)
        xo = torch.randn(B, 45, hw, hw, generator=g).to(dev)
        r = range(0, hw - 64 + 1, 16)
        corners = [(i, j) for i in r for j in r]
        seq = list(range(0, 1000, 1000 // steps))
        for use_graph in (False, True):
            if use_graph and B * len(corners) > DdimSampler.GRAPH_MAX_PATCHES:
                continue
            smp = DdimSampler(eng, use_graph=use_graph)
            for _ in range(2):
                smp.sample(x, xc, xo, seq, betas, corners, 64, keep_history=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                smp.sample(x, xc, xo, seq, betas, corners, 64, keep_history=False)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            print(f"B={B} {4*hw}x{4*hw} ({B*len(corners)} patches/step) {steps} DDIM steps, graph={use_graph}: "
                  f"{dt*1e3:8.2f} ms per image batch, {dt/steps*1e3:6.3f} ms per step")


if __name__ == "__main__":
    main()
