"""``utils/logging.py:9-29`` of the reference: image dump + checkpoint I/O (same dict format / file naming)."""
import os

import torch


def save_image(img, file_directory, normalize=False):
    import torchvision.utils as tvu
    d = os.path.dirname(file_directory)
    if d:
        os.makedirs(d, exist_ok=True)   # restore() calls this from several PNG worker threads
    tvu.save_image(img, file_directory, normalize=normalize)


def save_checkpoint(state, filename):
    d = os.path.dirname(filename)
    if d and not os.path.exists(d):
        os.makedirs(d)
    torch.save(state, filename + '.pth.tar')


def load_checkpoint(path, device):
    # the reference pickles argparse.Namespace objects ('params', 'config') inside the checkpoint
    if device is None:
        return torch.load(path, map_location='cpu', weights_only=False)
    print("load to this device:", device)
    return torch.load(path, map_location=device, weights_only=False)
