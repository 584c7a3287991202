#!/bin/bash
# GPU-box driver (run through gpurun)
cd $GRAFT_REPO_ROOT
O=gpurun_out/probe.log
: > $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) >> $O
timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time 2>&1 | grep -v "^profile" >> $O
WDM_ATTN_FUSED=0 timeout 200 python tools/profile_unet.py --patches 64 --iters 10 --time 2>&1 | grep -v "^profile" >> $O
python - >> $O 2>&1 <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import golden
from oracle import unet_oracle as O
from wavedm_b200 import engine
g = golden("unet_full.npz")
cfg = O.default_config()
sd = O.init_state_dict(cfg, seed=61)
x = torch.randn(2, 96, 64, 64, generator=torch.Generator().manual_seed(int(g["x_seed"])))
ref = torch.from_numpy(g["out"])
e = engine.UNetEngine(cfg, sd, torch.device("cuda", 0), precision="bf16")
out = e.forward(x.cuda(), torch.from_numpy(g["t"]).cuda()).cpu()
print("bf16 full UNet rel L2 vs reference golden (WDM_ATTN_FUSED=%s): %.5f" % (os.environ.get("WDM_ATTN_FUSED", "1"), ((out-ref).pow(2).sum()/ref.pow(2).sum()).sqrt().item()))
PY
cat $O
