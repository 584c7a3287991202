"""RainDrop loader (reference ``datasets/raindrop.py:14-150``): host-side data loading, OUT OF SCOPE for kernels
 and kept only so ``train_diffusion.py`` / ``eval_diffusion.py`` run unmodified. The contract the hot
path consumes is the batch tuple ``(x[B,6,H,W] in [0,1] (input || gt), image id, total_image)``."""
import os
import random
import re

import numpy as np
import torch
import torch.utils.data
import torch.utils.data.distributed as distributed


class RainDropDataset(torch.utils.data.Dataset):
    def __init__(self, dir, patch_size, n, transforms, filelist=None, parse_patches=True):
        super().__init__()
        if filelist is None:
            inp_dir = os.path.join(dir, 'input')
            images = [f for f in os.listdir(inp_dir) if os.path.isfile(os.path.join(inp_dir, f))]
            input_names = [os.path.join(inp_dir, i) for i in images]
            gt_names = [os.path.join(dir, 'gt', i.replace('rain', 'clean')) for i in images]
            print(len(input_names))
            order = list(range(len(input_names)))
            random.shuffle(order)
            input_names = [input_names[i] for i in order]
            gt_names = [gt_names[i] for i in order]
            self.dir = None
        else:
            self.dir = dir
            with open(os.path.join(dir, filelist)) as f:
                input_names = [l.strip() for l in f.readlines()]
            gt_names = [i.replace('input', 'gt') for i in input_names]
        self.input_names, self.gt_names = input_names, gt_names
        self.patch_size, self.transforms, self.n, self.parse_patches = patch_size, transforms, n, parse_patches

    def _open(self, name, rgb=False):
        import PIL.Image
        img = PIL.Image.open(os.path.join(self.dir, name) if self.dir else name)
        return img.convert('RGB') if rgb else img

    def get_images(self, index):
        import PIL.Image
        input_name, gt_name = self.input_names[index], self.gt_names[index]
        img_id = re.split('/', input_name)[-1][:-4]
        input_img = self._open(input_name)
        try:
            gt_img = self._open(gt_name)
        except Exception:
            gt_img = self._open(gt_name, rgb=True)
        if self.parse_patches:
            w, h = input_img.size
            th = tw = self.patch_size
            if w == tw and h == th:
                ii, jj = [0] * self.n, [0] * self.n
            else:
                ii = [random.randint(0, h - th) for _ in range(self.n)]
                jj = [random.randint(0, w - tw) for _ in range(self.n)]
            total = self.transforms(input_img.resize((720, 480), PIL.Image.LANCZOS)).repeat(self.n, 1, 1, 1)
            outs = []
            for i, j in zip(ii, jj):
                box = (j, i, j + tw, i + th)
                outs.append(torch.cat([self.transforms(input_img.crop(box)), self.transforms(gt_img.crop(box))], dim=0))
            return torch.stack(outs, dim=0), img_id, total
        # whole-image restoration: 720x480, capped at 1024, rounded up to multiples of 16
        input_img = input_img.resize((720, 480), PIL.Image.LANCZOS)
        wd, ht = input_img.size
        if ht > wd and ht > 1024:
            wd, ht = int(np.ceil(wd * 1024 / ht)), 1024
        elif ht <= wd and wd > 1024:
            ht, wd = int(np.ceil(ht * 1024 / wd)), 1024
        wd, ht = int(16 * np.ceil(wd / 16.0)), int(16 * np.ceil(ht / 16.0))
        input_img = input_img.resize((wd, ht), PIL.Image.LANCZOS)
        gt_img = gt_img.resize((wd, ht), PIL.Image.LANCZOS)
        return torch.cat([self.transforms(input_img), self.transforms(gt_img)], dim=0), img_id, self.transforms(input_img)

    def __getitem__(self, index):
        return self.get_images(index)

    def __len__(self):
        return len(self.input_names)


class RainDrop:
    def __init__(self, args, config):
        import torchvision
        self.args, self.config = args, config
        self.transforms = torchvision.transforms.Compose([torchvision.transforms.ToTensor()])

    def get_loaders(self, parse_patches=True, validation='raindrop'):
        print("=> evaluating raindrop test set...")
        cfg = self.config
        root = os.path.join(cfg.data.data_dir, 'raindrop')
        train_ds = RainDropDataset(os.path.join(root, 'train'), n=cfg.training.patch_n, patch_size=cfg.data.patch_size,
                                   transforms=self.transforms, filelist=None, parse_patches=parse_patches)
        val_ds = RainDropDataset(os.path.join(root, 'raindrop_test'), n=cfg.training.patch_n,
                                 patch_size=cfg.data.patch_size, transforms=self.transforms, parse_patches=parse_patches)
        if not parse_patches:
            cfg.sampling.batch_size = 1
        samp = lambda ds: distributed.DistributedSampler(ds, num_replicas=self.args.world_size, rank=self.args.rank)
        train_loader = torch.utils.data.DataLoader(train_ds, batch_size=cfg.training.batch_size, sampler=samp(train_ds),
                                                   num_workers=cfg.data.num_workers, pin_memory=True)
        val_loader = torch.utils.data.DataLoader(val_ds, batch_size=cfg.sampling.batch_size, shuffle=False,
                                                 sampler=samp(val_ds), num_workers=cfg.data.num_workers, pin_memory=True)
        return train_loader, val_loader
