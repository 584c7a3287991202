"""Micro-benchmark of the DWT / IWT kernels (HBM-bound): CUDA-event timing, working set >> L2 by rotating
over several buffers. Prints GB/s per variant against MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavedm_b200 import _lib  # noqa: E402
from wavedm_b200.wavelet import dwt4x4, iwt4x4  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    peak = 6533.2
    try:
        peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    lib = _lib.load()
    for (B, H, W) in [(64, 256, 256), (256, 256, 256), (64, 512, 512)]:
        nbuf = max(2, int(1.5e9 // (B * 3 * H * W * 4)))  # rotate over >= 1.5 GB of inputs
        nbuf = min(nbuf, 16)
        xs = [torch.randn(B, 3, H, W, device=dev) for _ in range(nbuf)]
        ys = [torch.empty(B, 48, H // 4, W // 4, device=dev) for _ in range(nbuf)]
        bytes_per = 2 * 4 * B * 3 * H * W
        st = torch.cuda.current_stream().cuda_stream
        for name, impl in (("direct", _lib.WDM_WT_IMPL_DIRECT), ("tma", _lib.WDM_WT_IMPL_TMA)):
            for kind in ("dwt", "iwt"):
                def run(i):
                    if kind == "dwt":
                        r = lib.wdm_dwt4x4_fwd(xs[i % nbuf].data_ptr(), ys[i % nbuf].data_ptr(), B, H, W, impl, st)
                    else:
                        r = lib.wdm_iwt4x4_fwd(ys[i % nbuf].data_ptr(), xs[i % nbuf].data_ptr(), B, H // 4, W // 4,
                                               impl, st)
                    assert r == 0, r
                for i in range(5):
                    run(i)
                iters = 40
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(iters):
                    run(i)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                gbs = bytes_per / ms / 1e6
                print(f"{kind} {name:6s} B={B} {H}x{W} nbuf={nbuf}: {ms*1e3:8.1f} us  {gbs:8.1f} GB/s  "
                      f"{gbs/peak:.3f} of measured peak {peak}")
    # torch copy for calibration on this box
    a = torch.empty(1 << 28, device=dev)
    b = torch.empty(1 << 28, device=dev)
    for _ in range(3):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"torch copy 1 GiB fp32: {2*4*(1<<28)/ms/1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
