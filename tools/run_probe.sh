#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from wavedm_b200 import engine
from wavedm_b200.configs import default_config
from wavedm_b200.unet import DiffusionUNet
from wavedm_b200.sampler import make_patch_table
dev = torch.device("cuda", 0)
cfg = default_config(); torch.manual_seed(61)
net = DiffusionUNet(cfg); eng = engine.UNetEngine(cfg, net.state_dict(), dev, precision="bf16")
B = 64
xc = torch.randn(B, 48, 64, 64, device=dev); xt = torch.randn(B, 3, 64, 64, device=dev); xo = torch.randn(B, 45, 64, 64, device=dev)
patches, first = make_patch_table(B, [(0, 0)], dev)
xin = torch.empty(B, 64, 64, eng.cin_pad, device=dev, dtype=eng.dtype)
eps = torch.randn(B, 3, 64, 64, device=dev); x0 = torch.empty_like(xt); xn = torch.empty_like(xt)
for name, fn in (("gather", lambda: eng.gather([xc, xt, xo], patches, out=xin)), ("ddim_step", lambda: eng.ddim_step(eps, patches, first, xt, x0, xn, 0.5, 0.6))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "us per call:", e0.elapsed_time(e1) / 50 * 1e3)
PY
