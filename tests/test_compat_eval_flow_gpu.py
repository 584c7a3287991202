"""GPU: the reference's eval script flow (eval_diffusion.py:57-103) run against the compat/ import shims.

The driver below is eval_diffusion.py's main() statement for statement -- `import models, datasets, utils`, YAML ->
dict2namespace, seeds, `dist.init_process_group(backend='nccl')`, `datasets.__dict__[config.data.dataset](args, config)`,
`get_loaders(parse_patches=False, ...)`, `DenoisingDiffusion_Wavelet(args, config)`, `DiffusiveRestoration(...).restore(...)`
-- with only the command line replaced by an argparse.Namespace and the dataset root pointing at two synthetic 720x480
"raindrop" pairs. It runs in a subprocess with compat/ first on PYTHONPATH, i.e. exactly what INTEGRATION.md tells a
reference user to do. 120x180 wavelet image -> 45 overlapping patches (odd), HFRM branch, x0_preds[-5], PNG side effects."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu

DRIVER = textwrap.dedent('''
    import argparse, os, sys, yaml
    import numpy as np
    import torch
    import torch.distributed as dist
    import models
    import datasets
    import utils
    from models import DenoisingDiffusion, DenoisingDiffusion_Wavelet, DiffusiveRestoration

    def dict2namespace(config):
        namespace = argparse.Namespace()
        for key, value in config.items():
            setattr(namespace, key, dict2namespace(value) if isinstance(value, dict) else value)
        return namespace

    root = sys.argv[1]
    args = argparse.Namespace(config="raindrop_wavelet.yml", resume=os.path.join(root, "ddpm.pth.tar"), grid_r=16,
                              sampling_timesteps=5, test_set="raindrop", image_folder=os.path.join(root, "results"), seed=61,
                              init_method="env://", rank=0, world_size=1, local_rank=0,
                              hfrm_ckpt=os.path.join(root, "hfrm.pth"))
    with open(os.path.join(root, "configs", args.config), "r") as f:
        config = dict2namespace(yaml.safe_load(f))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = '5678'
    os.environ["RANK"] = "0"
    os.environ['WORLD_SIZE'] = '1'
    device = torch.device("cuda", args.local_rank)
    config.device = device
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    torch.cuda.manual_seed_all(args.seed)
    torch.backends.cudnn.benchmark = True
    dist.init_process_group(backend='nccl')
    DATASET = datasets.__dict__[config.data.dataset](args, config)
    _, val_loader = DATASET.get_loaders(parse_patches=False, validation=args.test_set)
    diffusion = DenoisingDiffusion_Wavelet(args, config)
    model = DiffusiveRestoration(diffusion, args, config)
    model.restore(val_loader, validation=args.test_set, r=args.grid_r)
    eng = diffusion.model.module.engine()
    print("ENGINE", eng.precision, *eng.counters())
    dist.destroy_process_group()
''')


def test_eval_diffusion_flow_through_compat(tmp_path):
    import PIL.Image
    import torch
    import yaml
    from wavedm_b200.configs import RAINDROP_WAVELET
    from wavedm_b200.hfrm import HFRM
    from wavedm_b200.unet import DiffusionUNet
    from wavedm_b200.configs import default_config
    root = str(tmp_path)
    cfg = {k: dict(v) for k, v in RAINDROP_WAVELET.items()}
    cfg["data"]["data_dir"] = root
    cfg["data"]["num_workers"] = 0
    os.makedirs(os.path.join(root, "configs"))
    with open(os.path.join(root, "configs", "raindrop_wavelet.yml"), "w") as f:
        yaml.safe_dump(cfg, f)
    rng = np.random.default_rng(0)
    for split in ("train", "raindrop_test"):
        for sub in ("input", "gt"):
            os.makedirs(os.path.join(root, "raindrop", split, sub))
    for i in range(2):
        for split in ("train", "raindrop_test"):
            img = rng.integers(0, 256, (480, 720, 3), dtype=np.uint8)
            PIL.Image.fromarray(img).save(os.path.join(root, "raindrop", split, "input", f"{i}_rain.png"))
            PIL.Image.fromarray(255 - img).save(os.path.join(root, "raindrop", split, "gt", f"{i}_clean.png"))
    # a reference-format checkpoint (ddm_wavelet.py:284-292) of seeded default-init weights + the HFRM checkpoint
    torch.manual_seed(61)
    net = DiffusionUNet(default_config())
    sd = net.state_dict()
    opt = torch.optim.Adam(net.parameters(), lr=4e-5, eps=1e-8)
    torch.save({"epoch": 1, "step": 1, "state_dict": sd, "optimizer": opt.state_dict(),
                "ema_helper": {k: v.clone() for k, v in net.named_parameters()}, "params": None, "config": None},
               os.path.join(root, "ddpm.pth.tar"))
    torch.manual_seed(5)
    torch.save(HFRM(in_channel=3, dim=32, mid_blk_num=6, enc_blk_nums=[2, 2, 2, 4], dec_blk_nums=[2, 2, 2, 2]).state_dict(),
               os.path.join(root, "hfrm.pth"))
    drv = os.path.join(root, "eval_driver.py")
    with open(drv, "w") as f:
        f.write(DRIVER)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "compat"), REPO, env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, drv, root], env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    out = r.stdout
    assert "psnr all torch" in out and out.count("psnr this") == 2
    eng = [l for l in out.splitlines() if l.startswith("ENGINE")][0].split()
    assert eng[1] == "bf16" and int(eng[2]) > 0 and int(eng[3]) == 0, eng   # tensor cores only, odd patch count included
    res = os.path.join(root, "results", "RainDrop", "raindrop")
    names = os.listdir(res)
    for i in range(2):
        # the id is the DataLoader-collated list (f"{y}" == "['0_rain']"), exactly as in the reference (restoration.py:156-163)
        for suffix in ("output", "cond", "gt", "all_wdnet", "lrdiff_hrgt", "lrgt_hrcond", "lrgt_hrwdnet"):
            hits = [n for n in names if f"{i}_rain" in n and n.endswith(f"_{suffix}.png")]
            assert len(hits) == 1, (i, suffix, names)
        out_png = [n for n in names if f"{i}_rain" in n and n.endswith("_output.png")][0]
        im = np.asarray(PIL.Image.open(os.path.join(res, out_png)))
        assert im.shape == (480, 720, 3)
