"""Runs N UNet engine forwards at P patches (random weights/input) -- the target for ncu launch lists / captures.
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv \
        python tools/profile_unet.py --patches 64 --iters 2
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavedm_b200 import engine  # noqa: E402
from wavedm_b200.configs import default_config  # noqa: E402
from wavedm_b200.unet import DiffusionUNet  # noqa: E402


def span_table(eng, x, t, out, iters):
    import collections
    import tempfile
    path = os.path.join(tempfile.gettempdir(), f"wdm_spans_{os.getpid()}.csv")
    os.environ["WDM_PROFILE_DUMP"] = path
    eng.profile(True)
    for _ in range(iters):
        eng.forward_nhwc(x, t, out=out)
    eng.profile_read()
    eng.profile(False)
    del os.environ["WDM_PROFILE_DUMP"]
    agg = collections.OrderedDict()
    for line in open(path):
        tc, M, N, K, taps, W, tag, us, fl = line.strip().split(",")
        key = (int(tc), int(M), int(N), int(K), int(taps), int(W), int(tag))
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += float(us)
        a[2] += float(fl)
    os.remove(path)
    tot = sum(v[1] for v in agg.values()) / iters
    print(f"contraction launches per forward: {sum(v[0] for v in agg.values()) // iters}, {tot:.1f} us (event-bracketed, no PDL overlap)")
    print("  tc       M     N      K taps   W tag    n   us/launch  TFLOP/s  share")
    for k, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[0]:2d} {k[1]:7d} {k[2]:5d} {k[3]:6d} {k[4]:4d} {k[5]:3d} {k[6]:3d} {n // iters:4d} {us / n:10.1f} {fl / us / 1e6:8.1f} {us / iters / tot * 100:5.1f}%")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--spans", action="store_true", help="per-shape table of the contraction launches (CUDA events)")
    ap.add_argument("--split", type=int, default=0, help="time the batch as N sub-batches on N streams (tail backfill)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = default_config()
    torch.manual_seed(61)
    net = DiffusionUNet(cfg)
    eng_sd = net.state_dict()
    eng = engine.UNetEngine(cfg, eng_sd, dev, precision=a.precision, max_patches=a.patches)
    del net
    x = torch.randn(a.patches, 64, 64, eng.cin_pad, device=dev).to(eng.dtype)
    t = torch.tensor([500.0], device=dev)
    out = torch.empty(a.patches, 3, 64, 64, device=dev)
    eng.forward_nhwc(x, t, out=out)
    torch.cuda.synchronize()
    if a.split > 1:
        n = a.split
        Ps = a.patches // n
        engs = [eng] + [engine.UNetEngine(cfg, eng_sd, dev, precision=a.precision, max_patches=Ps) for _ in range(n - 1)]
        streams = [torch.cuda.Stream() for _ in range(n)]
        xs = [x[i * Ps:(i + 1) * Ps].contiguous() for i in range(n)]
        outs = [torch.empty(Ps, 3, 64, 64, device=dev) for _ in range(n)]
        torch.cuda.synchronize()
        def run_split():
            cur = torch.cuda.current_stream()
            for i in range(n):
                streams[i].wait_stream(cur)
                with torch.cuda.stream(streams[i]):
                    engs[i].forward_nhwc(xs[i], t, out=outs[i])
            for i in range(n):
                cur.wait_stream(streams[i])
        for _ in range(2):
            run_split()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.iters):
            run_split()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"P={a.patches} as {n} x {Ps} on {n} streams: {ms:.3f} ms/forward, {79.945e9 * a.patches / ms / 1e9:.1f} TFLOP/s algorithmic")
        return
    if a.time:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(a.iters):
            eng.forward_nhwc(x, t, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"P={a.patches} {a.precision}: {ms:.3f} ms/forward, {79.945e9 * a.patches / ms / 1e9:.1f} TFLOP/s algorithmic")
        eng.profile(True)
        eng.forward_nhwc(x, t, out=out)
        print("profile (tc_ms, tc_flops, tc_n, simt_ms, simt_flops, simt_n):", eng.profile_read())
        eng.profile(False)
        if a.spans:
            span_table(eng, x, t, out, a.iters)
    else:
        for _ in range(a.iters):
            eng.forward_nhwc(x, t, out=out)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
