#!/usr/bin/env python
"""bench.py -- restored images/sec @256x256, 50-step DDIM (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W                (N > 1: launched under torchrun by the driver)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic input: DWT(cond) / DWT(gt-> HF bands) ->
50 DDIM steps of the conditional UNet over every 64x64 wavelet-domain patch -> pick x0_preds[-5] -> concat HF ->
IWT -> clamp, for B images per GPU (workload = BASELINE.json configs[2]: batch 64, 256x256, 50 DDIM steps, bf16
tensor-core UNet). Weak scaling: every rank restores its own B images; NCCL is used only for the initial weight
broadcast and the gather of restored images to rank 0 (inside the timed region).

  value : images/s with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max over ranks)
  e2e   : the same through the public API (DiffusiveRestoration.restore_batch) from pinned HOST buffers, with
          the H2D copy of the inputs and the D2H copy of the restored images inside the timed region
  roofline     : the dominant kernel class (UNet contraction kernels) -- algorithmic FLOPs / event-timed duration
                 of exactly those launches, measured live in a profiled pass of the same workload
  roofline_dwt : the DWT kernel's HBM GB/s on >L2 working sets (the metric's second half)
  cpu_baseline : the CPU oracle port of the same path timed on this box's host cores (bounded sample)

The HFRM (one-shot high-frequency CNN, SURVEY.md 8f-1 "next") is bypassed in BOTH arms: the 45 high-frequency
channels fed to the UNet are the HF bands of the DWT of the synthetic ground truth (the reference's own
`if 0:` branch, restoration.py:99-100), so the two arms time the same computation.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

UNET_GFLOP_PER_PATCH = 79.945  # algorithmic 2*MAC per 96x64x64 patch per UNet call
H = W = 256
DDIM_STEPS = 50
SEED = 61


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ddim-steps", type=int, default=DDIM_STEPS)
    ap.add_argument("--max-patches", type=int, default=64)
    ap.add_argument("--wavelet-in-unet", action="store_true",
                    help="NOT the BASELINE config: data.wavelet_in_unet (DWT / IWT inside the network at every DDIM step, "
                         "pixel-domain sampler, out_ch 48); the line is labelled accordingly")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-seconds", type=float, default=15.0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        # under-load samples: upper half of the sorted clocks is dominated by busy samples; report the median of all
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ synthetic data
def synth_inputs(batch, rank, device=None, pinned=False):
    """generator seed 61 (+rank), x = rand(B,6,H,W) in [0,1) (cond || gt), noise = randn(B,3,H/4,W/4)
    drawn on the CPU and passed in."""
    import torch
    g = torch.Generator().manual_seed(SEED + rank)
    x = torch.rand(batch, 6, H, W, generator=g)
    noise = torch.randn(batch, 3, H // 4, W // 4, generator=g)
    if pinned:
        x, noise = x.pin_memory(), noise.pin_memory()
    if device is not None:
        x, noise = x.to(device), noise.to(device)
    return x, noise


def make_cfg(precision, device, wavelet_in_unet=False):
    from wavedm_b200.configs import default_config
    cfg = default_config()
    cfg.device = device
    cfg.model.engine_precision = precision
    if wavelet_in_unet:  # the only self-consistent setting of this mode
        cfg.data.wavelet_in_unet = True
        cfg.model.use_other_channels, cfg.model.in_channels, cfg.model.out_ch = False, 93, 48
    return cfg


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_restore_sample(n_images, ddim_steps, threads):
    """The CPU oracle port (oracle/unet_oracle.py + oracle/dwt_oracle.c) of the same path on `n_images` images
    for `ddim_steps` DDIM steps of the 50-step schedule. Returns seconds."""
    import torch
    from oracle import dwt_oracle as DO
    from oracle import unet_oracle as O
    torch.set_num_threads(threads)
    cfg = O.default_config()
    if not hasattr(cpu_restore_sample, "sd"):
        cpu_restore_sample.sd = O.init_state_dict(cfg, seed=SEED)
    sd = cpu_restore_sample.sd
    x, noise = synth_inputs(n_images, 0)
    betas = O.beta_schedule(cfg)
    seq = O.sampling_seq(1000, DDIM_STEPS)
    seq_run = seq[len(seq) - ddim_steps:]  # the first `ddim_steps` iterations of the descending schedule
    t0 = time.perf_counter()
    with torch.no_grad():
        x_cond = torch.from_numpy(DO.dwt(x[:, :3].numpy(), flags=1))
        x_gt = torch.from_numpy(DO.dwt(x[:, 3:].numpy(), flags=1))
        x_other = x_gt[:, 3:]
        xs, x0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), noise, x_cond, x_other,
                                            seq_run, betas, [(0, 0)], 64)
        lat = x0p[-5] if len(x0p) >= 5 else x0p[-1]
        out = DO.iwt(torch.cat([lat[:, :3], x_other], 1).numpy(), flags=1)
    dt = time.perf_counter() - t0
    assert out.shape == (n_images, 3, H, W)
    return dt


def cpu_baseline(target_seconds, threads):
    """Bounded sample: one image, as many of the 50 DDIM steps as fit in ~target_seconds (probe with 2)."""
    probe_steps = 2
    t_probe = cpu_restore_sample(1, probe_steps, threads)
    per_step = t_probe / probe_steps
    steps = int(max(2, min(DDIM_STEPS, target_seconds / max(per_step, 1e-3))))
    t = cpu_restore_sample(1, steps, threads)
    ips = (steps / DDIM_STEPS) / t  # images/s normalised to the 50-step schedule
    return {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"1 image 256x256, {steps} of {DDIM_STEPS} DDIM steps (UNet 1 patch/step) + DWT/IWT in {t:.2f} s, "
                      f"scaled to {DDIM_STEPS} steps; torch {threads} threads (oracle/unet_oracle.py + dwt_oracle.c)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    per_step_target = float(os.environ.get("WDM_BENCH_REF_STEP_SECONDS", "8.0"))  # CPU seconds of work per timed step
    cpu_restore_sample(1, 2, threads)  # build weights, warm caches
    probe = cpu_restore_sample(1, 2, threads) / 2
    n = int(max(2, min(DDIM_STEPS, per_step_target / max(probe, 1e-3))))
    for _ in range(min(args.warmup, 1)):
        cpu_restore_sample(1, n, threads)
    ts = [cpu_restore_sample(1, n, threads) for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    val = (n / DDIM_STEPS) / t
    sample = (f"each step = 1 image 256x256, {n} of {DDIM_STEPS} DDIM steps + DWT/IWT, scaled to {DDIM_STEPS} steps; "
              f"CPU oracle port, torch {threads} threads")
    line = {"impl": "reference", "metric": "restored images/sec @256x256, 50-step DDIM", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, torch.__version__),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, torch_version):
    if getattr(args, "wavelet_in_unet", False):
        return {"workload": f"NOT a BASELINE config -- data.wavelet_in_unet variant: batch {args.batch}/GPU, 256x256, "
                            f"{args.ddim_steps} DDIM steps, {args.precision} UNet with the DWT (2x3 ch) and IWT (48 ch) inside "
                            f"every step, pixel-domain sampler (1 256x256 patch/image), seed-61 default-init weights",
                "global_batch": args.batch * args.gpus, "images_per_gpu": args.batch, "ddim_steps": args.ddim_steps,
                "l2": "inputs+activations per step far exceed L2 (126 MB); weights 313 MB bf16", "torch": torch_version,
                "hfrm": "not part of this mode (x_other = None, restoration.py:98-104)"}
    return {"workload": f"BASELINE.json configs[2]: batch {args.batch}/GPU, 256x256, {args.ddim_steps} DDIM steps, "
                        f"{args.precision} UNet (1 64x64 wavelet patch/image), raindrop_wavelet.yml, seed-61 default-init weights",
            "global_batch": args.batch * args.gpus, "images_per_gpu": args.batch, "ddim_steps": args.ddim_steps,
            "l2": "inputs+activations per step far exceed L2 (126 MB); weights 313 MB bf16", "torch": torch_version,
            "hfrm": "bypassed in both arms (x_other = HF bands of DWT(gt))"}


# ------------------------------------------------------------------------------------------------ our arm
def dwt_roofline(dev, peak):
    import torch
    from wavedm_b200 import _lib
    lib = _lib.load()
    B = 256
    nbuf = 3
    xs = [torch.randn(B, 3, H, W, device=dev) for _ in range(nbuf)]
    ys = [torch.empty(B, 48, H // 4, W // 4, device=dev) for _ in range(nbuf)]
    st = torch.cuda.current_stream().cuda_stream
    for i in range(3):
        lib.wdm_dwt4x4_fwd(xs[i % nbuf].data_ptr(), ys[i % nbuf].data_ptr(), B, H, W, 0, st)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    iters = 30
    e0.record()
    for i in range(iters):
        lib.wdm_dwt4x4_fwd(xs[i % nbuf].data_ptr(), ys[i % nbuf].data_ptr(), B, H, W, 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_per = 2 * 4 * B * 3 * H * W
    ach = bytes_per / ms / 1e6
    return {"kernel": "dwt4x4_direct_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": None,
            "note": f"B={B} 256x256 fp32, {nbuf} rotating buffer pairs ({nbuf * bytes_per / 1e6:.0f} MB > L2), "
                    f"algorithmic bytes/launch {bytes_per}"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from wavedm_b200 import _lib
    from wavedm_b200.harness import build_restorer

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: wavedm_b200 has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise RuntimeError("launch with torchrun --nproc-per-node N for --gpus N")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks.update(json.load(f))
            peaks["src"] = "measured"
    except Exception:
        pass

    cfg = make_cfg(args.precision, dev, args.wavelet_in_unet)
    restorer = build_restorer(cfg, dev, sampling_timesteps=args.ddim_steps, max_patches=args.max_patches, seed=SEED,
                              broadcast=world > 1)
    lib = _lib.load()
    B = args.batch
    x_dev, noise_dev = synth_inputs(B, rank, device=dev)
    x_pin, noise_pin = synth_inputs(B, rank, pinned=True)
    x_gt_hf = None

    if args.wavelet_in_unet:  # the sampler runs in the pixel domain: the initial noise is image-sized
        g = torch.Generator().manual_seed(SEED + 100 + rank)
        noise_pin = torch.randn(B, 3, H, W, generator=g).pin_memory()
        noise_dev = noise_pin.to(dev)

    def step_device():
        if args.wavelet_in_unet:
            return restorer.restore_batch(x_dev, r=16, noise=noise_dev)["output"]
        xo = restorer.diffusion.wavelet_dec(2 * x_dev[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
        res = restorer.restore_batch(x_dev, r=16, noise=noise_dev, x_other=xo)
        return res["output"]

    gathered = None
    if world > 1:
        gathered = [torch.empty(B, 3, H, W, device=dev) for _ in range(world)] if rank == 0 else None

    def step_full():
        out = step_device()
        if world > 1:
            dist.gather(out, gathered, dst=0)
        return out

    out_pin = torch.empty(B, 3, H, W).pin_memory()

    def step_e2e():
        xh = x_pin.to(dev, non_blocking=True)
        nh = noise_pin.to(dev, non_blocking=True)
        if args.wavelet_in_unet:
            out = restorer.restore_batch(xh, r=16, noise=nh)["output"]
        else:
            xo = restorer.diffusion.wavelet_dec(2 * xh[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
            out = restorer.restore_batch(xh, r=16, noise=nh, x_other=xo)["output"]
        if world > 1:
            dist.gather(out, gathered, dst=0)
        out_pin.copy_(out, non_blocking=True)  # pinned destination; the timed region ends with a device synchronize
        return out_pin

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_full()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = lib.wdm_launch_counter()
    ms_total = timed(step_full, args.steps)
    launches = lib.wdm_launch_counter() - l0
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = B * world / (ms_per_step / 1e3)

    # e2e: host buffers in, host images out
    step_e2e()
    ms_e2e = timed(step_e2e, max(1, min(args.steps, 3))) / max(1, min(args.steps, 3))
    e2e_val = B * world / (ms_e2e / 1e3)

    # roofline of the dominant kernel class: one profiled UNet-heavy pass of the same workload
    eng = restorer.diffusion.model.module.engine()
    eng.profile(True)
    step_device()
    prof = eng.profile_read()
    eng.profile(False)
    tc_ms, tc_fl, tc_n, s_ms, s_fl, s_n = prof
    if tc_n > 0:
        kname, k_ms, k_fl, k_n = "gemm_tc_kernel (tcgen05 implicit-GEMM conv / 1x1 / attention)", tc_ms, tc_fl, tc_n
    else:
        kname, k_ms, k_fl, k_n = "gemm_simt_kernel (CUDA-core implicit-GEMM)", s_ms, s_fl, s_n
    peak_tf = peaks["bf16_tflops_sustained"]
    # achieved = ALGORITHMIC flops of the contraction work (79.945 GFLOP per 64x64 patch and UNet call, the
    # reference's own operation count) / CUDA-event time of exactly those launches; the executed count is lower (sub-pixel
    # upsample-conv -6.0, four-contraction attention -1.5 GFLOP per patch and call) and is reported next to it
    n_patch_calls = B * args.ddim_steps
    alg_fl = UNET_GFLOP_PER_PATCH * 1e9 * n_patch_calls
    ach_tf = alg_fl / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
    exe_tf = k_fl / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]  # dram read+write per launch from the committed ncu capture
    except Exception:
        pass
    roof = {"kernel": kname, "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": ach_tf / peak_tf, "traffic": traffic,
            "executed_tflops": exe_tf, "executed_frac": exe_tf / peak_tf,
            "algorithmic_flops_per_launch": alg_fl / max(k_n, 1), "executed_flops_per_launch": k_fl / max(k_n, 1),
            "algorithmic_bytes_per_launch": (getattr(eng, "last_tc_bytes", 0.0) / max(tc_n, 1)) if tc_n > 0 else None,
            "launches": k_n, "avg_launch_ms": k_ms / max(k_n, 1),
            "share_of_step": k_ms / ms_per_step, "other_contraction_ms": (s_ms if tc_n > 0 else 0.0),
            "peak_src": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a long step)",
            "algorithmic_gflop_per_patch_call": UNET_GFLOP_PER_PATCH,
            "note": "achieved = algorithmic 79.945 GFLOP/patch/call x patch-calls of the step / CUDA-event time of the "
                    "contraction launches (all tensor-core launches of the UNet calls of one profiled step); executed_* counts "
                    "2*M*N*K of what is launched; whole_step_* divides the same algorithmic flops by the whole step time"}
    e2e_alg_tf = alg_fl / (ms_per_step / 1e3) / 1e12
    roof["whole_step_algorithmic_tflops"] = e2e_alg_tf
    roof["whole_step_frac"] = e2e_alg_tf / peak_tf

    out = None
    if rank == 0:
        rdwt = dwt_roofline(dev, peaks["hbm_gbs"])
        cpu = None
        if not args.no_cpu_baseline and not args.wavelet_in_unet:
            cpu = cpu_baseline(args.cpu_sample_seconds, os.cpu_count() or 1)
        out = {"metric": "restored images/sec @256x256, 50-step DDIM", "value": value, "unit": "images/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
               "config": workload_config(args, torch.__version__), "clocks": clk,
               "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(x_pin.numel() * 4 + noise_pin.numel() * 4) * world,
                       "d2h_bytes_per_step": int(B * 3 * H * W * 4) * world, "ms_per_step": ms_e2e,
                       "note": "every rank copies its own inputs H2D from pinned memory and its restored images D2H into pinned memory"},
               "gpu_launches": int(launches), "roofline": roof, "roofline_dwt": rdwt, "cpu_baseline": cpu}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
