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
    # wavelet_in_unet per-step kernels: fused crop + DWT + concat + NHWC(bf16, 128 ch) gather, and the IWT of the NHWC result
    P, R = 128, 64
    nbuf = 4
    srcs = [(torch.randn(P, 3, 4 * R, 4 * R, device=dev), torch.randn(P, 3, 4 * R, 4 * R, device=dev)) for _ in range(nbuf)]
    outs = [torch.empty(P, R, R, 128, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    pats = torch.zeros(P, 3, dtype=torch.int32, device=dev)
    pats[:, 0] = torch.arange(P, dtype=torch.int32, device=dev)
    ys = [torch.randn(P, R, R, 64, device=dev) for _ in range(nbuf)]
    xo = [torch.empty(P, 3, 4 * R, 4 * R, device=dev) for _ in range(nbuf)]
    st = torch.cuda.current_stream().cuda_stream
    cases = {
        "dwt_gather (2x3ch fp32 -> 96(+32 pad)ch bf16 NHWC)":
            (lambda i: lib.wdm_gather_patches_dwt(srcs[i % nbuf][0].data_ptr(), srcs[i % nbuf][1].data_ptr(), 2, P, 4 * R, 4 * R,
                                                  pats.data_ptr(), P, R, 128, outs[i % nbuf].data_ptr(), 1, st),
             P * (6 * 16 * R * R * 4 + R * R * 128 * 2)),
        "iwt_nhwc (48 of 64 fp32 columns -> 3ch fp32 NCHW)":
            (lambda i: lib.wdm_iwt4x4_nhwc(ys[i % nbuf].data_ptr(), 64, P, R, xo[i % nbuf].data_ptr(), st),
             P * (R * R * 48 * 4 + 3 * 16 * R * R * 4)),
    }
    for name, (fn, nbytes) in cases.items():
        for i in range(5):
            assert fn(i) == 0
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(40):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        print(f"{name} P={P}: {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s algorithmic  {nbytes/ms/1e6/peak:.3f} of measured peak")
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
