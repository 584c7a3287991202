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

  parity       : (N = 1) the north_star's own parity config -- BASELINE.json configs[1]: batch 16, 256x256, 50 DDIM steps --
                 run on the fp32 engine AND on the benched bf16 engine against the golden vectors the reference itself
                 produced (tests/golden/sandwich_s50.npz): achieved latent / image / PSNR errors in the line
  gpu_eager_baseline : (N = 1) the reference's PyTorch path (oracle port, eager + cuDNN) on THIS B200, fp32 and bf16
                 autocast, bounded sample -- the honest GPU comparator; a reported baseline, not the product

  --config 5   : BASELINE.json configs[4] (512x512, grid_r = 16 -> 25 overlapping patches per image, 100 DDIM steps,
                 32 images per GPU); lines kept under profiles/
  --precision fp32 : the parity-mode engine (tensor-core contractions through a 3-way bf16 split); fp32_ffma: CUDA cores

  single_image : (N = 1, default config) BASELINE.json configs[0]'s operating point on the GPU: batch 1, one latent patch per
                 DDIM step, ms per step of the sampling loop (split-K contractions, graph replay); reported beside the metric

The HFRM (one-shot high-frequency CNN) runs in BOTH arms (csrc/wdm_hfrm.cu here, oracle/hfrm_oracle.py in the
CPU arms): `value` / `e2e` time the whole restore() computation. `--bypass-hfrm` restores the round-1 definition (the 45
high-frequency channels fed to the UNet are the HF bands of the DWT of the synthetic ground truth, the reference's own
`if 0:` branch, restoration.py:99-100).
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
H = W = 256          # overwritten by --config 5 (512)
DDIM_STEPS = 50      # the schedule the metric is quoted on; --config 5 uses 100
SEED = 61
GRID_R = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 5],
                    help="BASELINE.json configs index + 0: 2 = batch 64/GPU, 256x256, 50 DDIM steps (the metric's config); "
                         "5 = configs[4]: 512x512, 25 patches/image (grid_r 16), 100 DDIM steps, 32 images/GPU")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 64; 32 for --config 5)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32_ffma"],
                    help="bf16 = tcgen05 throughput mode; fp32 = parity mode on the tensor cores (3-way bf16 split); fp32_ffma = parity "
                         "mode on CUDA cores")
    ap.add_argument("--ddim-steps", type=int, default=None)
    ap.add_argument("--max-patches", type=int, default=None, help="patches per UNet call (default 64; 148 for --config 5)")
    ap.add_argument("--wavelet-in-unet", action="store_true",
                    help="NOT the BASELINE config: data.wavelet_in_unet (DWT / IWT inside the network at every DDIM step, "
                         "pixel-domain sampler, out_ch 48); the line is labelled accordingly")
    ap.add_argument("--bypass-hfrm", action="store_true",
                    help="round-1 behaviour: skip the HFRM (models/arch.py) in both arms, x_other = HF bands of DWT(gt)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-seconds", type=float, default=15.0)
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (N = 1 only anyway)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch-eager GPU comparator (N = 1 only anyway)")
    args = ap.parse_args()
    global H, W, DDIM_STEPS
    if args.config == 5:
        H = W = 512
        DDIM_STEPS = 100
    if args.batch is None:
        args.batch = 32 if args.config == 5 else 64
    if args.ddim_steps is None:
        args.ddim_steps = DDIM_STEPS
    if args.max_patches is None:
        # config 5: 800 patches per DDIM step in calls of 148 = one patch per SM: every level's tile count is a whole number
        # of waves (32x32: 4 pair tiles / patch, 16x16: 2, 8x8: 1/2 with 192-wide N tiles, 64x64: 16 tiles of 256 rows)
        args.max_patches = 148 if args.config == 5 else 64
    return args


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
BYPASS_HFRM = False  # set from --bypass-hfrm


def cpu_restore_sample(n_images, ddim_steps, threads):
    """The CPU oracle port (oracle/unet_oracle.py + oracle/dwt_oracle.c) of the same path on `n_images` images
    for `ddim_steps` DDIM steps of the DDIM_STEPS-step schedule (overlapping 64x64 patches on the grid_r grid when the
    wavelet-domain image is larger than one patch). Returns seconds."""
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
        if BYPASS_HFRM:
            x_other = x_gt[:, 3:]
        else:  # restoration.py:94-102: HFRM on the [0, 1] conditioning image, its DWT supplies the high-frequency bands
            from oracle import hfrm_oracle as HO
            if not hasattr(cpu_restore_sample, "hsd"):
                cpu_restore_sample.hsd = HO.fill_params(HO.default_shapes(), SEED)
            wd = HO.hfrm_forward(cpu_restore_sample.hsd, x[:, :3].contiguous())
            x_other = torch.from_numpy(DO.dwt(wd.numpy(), flags=1))[:, 3:].contiguous()
        hl, wl = O.overlapping_grid_indices(H // 4, W // 4, 64, GRID_R)
        xs, x0p = O.ddim_sample_overlapping(lambda a, tt: O.unet_forward(sd, cfg, a, tt), noise, x_cond, x_other,
                                            seq_run, betas, [(i, j) for i in hl for j in wl], 64)
        lat = x0p[-5] if len(x0p) >= 5 else x0p[-1]
        out = DO.iwt(torch.cat([lat[:, :3], x_other], 1).numpy(), flags=1)
    dt = time.perf_counter() - t0
    assert out.shape == (n_images, 3, H, W)
    return dt


def patches_per_image():
    n = lambda d: len(range(0, d // 4 - 64 + 1, GRID_R)) + (1 if (d // 4 - 64) % GRID_R else 0)
    return n(H) * n(W)


def cpu_baseline(target_seconds, threads):
    """Bounded sample: one image, as many of the 50 DDIM steps as fit in ~target_seconds (probe with 2)."""
    probe_steps = 2
    t_probe = cpu_restore_sample(1, probe_steps, threads)
    per_step = t_probe / probe_steps
    steps = int(max(2, min(DDIM_STEPS, target_seconds / max(per_step, 1e-3))))
    t = cpu_restore_sample(1, steps, threads)
    ips = (steps / DDIM_STEPS) / t  # images/s normalised to the 50-step schedule
    return {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"1 image {H}x{W}, {steps} of {DDIM_STEPS} DDIM steps (UNet {patches_per_image()} patch(es)/step) + DWT/IWT in {t:.2f} s, "
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
    sample = (f"each step = 1 image {H}x{W}, {n} of {DDIM_STEPS} DDIM steps + DWT/IWT, scaled to {DDIM_STEPS} steps; "
              f"CPU oracle port, torch {threads} threads")
    line = {"impl": "reference", "metric": metric_name(), "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, torch.__version__),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def metric_name():
    return f"restored images/sec @{H}x{W}, {DDIM_STEPS}-step DDIM"


def workload_config(args, torch_version):
    if getattr(args, "wavelet_in_unet", False):
        return {"workload": f"NOT a BASELINE config -- data.wavelet_in_unet variant: batch {args.batch}/GPU, 256x256, "
                            f"{args.ddim_steps} DDIM steps, {args.precision} UNet with the DWT (2x3 ch) and IWT (48 ch) inside "
                            f"every step, pixel-domain sampler (1 256x256 patch/image), seed-61 default-init weights",
                "global_batch": args.batch * args.gpus, "images_per_gpu": args.batch, "ddim_steps": args.ddim_steps,
                "l2": "inputs+activations per step far exceed L2 (126 MB); weights 313 MB bf16", "torch": torch_version,
                "hfrm": "not part of this mode (x_other = None, restoration.py:98-104)"}
    which = "configs[4] (512x512 patched sampling, grid_r 16)" if args.config == 5 else "configs[2]"
    return {"workload": f"BASELINE.json {which}: batch {args.batch}/GPU, {H}x{W}, {args.ddim_steps} DDIM steps, "
                        f"{args.precision} UNet ({patches_per_image()} 64x64 wavelet patch(es)/image), raindrop_wavelet.yml, "
                        f"seed-61 default-init weights",
            "global_batch": args.batch * args.gpus, "images_per_gpu": args.batch, "ddim_steps": args.ddim_steps,
            "l2": "inputs+activations per step far exceed L2 (126 MB); weights 313 MB bf16", "torch": torch_version,
            "hfrm": ("bypassed in both arms (x_other = HF bands of DWT(gt))" if getattr(args, "bypass_hfrm", False) else
                     "on in both arms (models/arch.py HFRM once per image at full resolution -> DWT -> x_other and the "
                     "high-frequency bands of the output, restoration.py:94-135): value and e2e time the whole "
                     "restore_batch; seeded synthetic HFRM weights")}


# ------------------------------------------------------------------------------------------------ our arm
def parity_block(dev, precisions=("fp32", "fp32_ffma", "bf16"), batch=16, slots=(3, 12)):
    """The north_star's own parity config (BASELINE.json configs[1]: batch 16, 256x256, 50 DDIM steps, fp32; gates
    |PSNR_new - PSNR_ref| < 0.01 dB and per-pixel |d| < 1e-3 on the element restore() returns, restoration.py:106-135)
    through the public API, against tests/golden/sandwich_s50.npz -- two independent B = 1 runs of the UNMODIFIED reference
    (oracle/make_golden.py --only-s50; the reference sampler is batch-1 only), placed at two slots of a 16-image batch whose
    other images are random. Returns the ACHIEVED errors per engine precision (tests/test_parity_s50_gpu.py asserts on the
    same numbers). No oracle code runs here: the checker is the committed golden file."""
    import numpy as np
    import torch
    from wavedm_b200.harness import build_restorer
    from wavedm_b200.metrics import torchPSNR
    g = np.load(os.path.join(REPO, "tests", "golden", "sandwich_s50.npz"))
    seeds = [int(v) for v in g["seeds"]]
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand(batch, 6, 256, 256, generator=gen)
    noise = torch.randn(batch, 3, 64, 64, generator=gen)
    for slot, seed in zip(slots, seeds):
        gs = torch.Generator().manual_seed(seed)
        x[slot] = torch.rand(1, 6, 256, 256, generator=gs)[0]
        noise[slot] = torch.randn(1, 3, 64, 64, generator=gs)[0]
    lat_ref = torch.from_numpy(g["latent_m5"])
    out_ref = torch.from_numpy(g["out"])
    res = {"config": f"BASELINE.json configs[1]: batch {batch}, 256x256, {int(g['steps'])} DDIM steps; golden = the reference's own "
                     f"classes on CPU fp32 (tests/golden/sandwich_s50.npz, images at batch slots {list(slots)})",
           "gates": {"image_max_abs": 1e-3, "psnr_db": 0.01}}
    for prec in precisions:
        cfg = make_cfg(prec, dev)
        restorer = build_restorer(cfg, dev, sampling_timesteps=int(g["steps"]), max_patches=64, seed=SEED)
        xd = x.to(dev)
        x_other = restorer.diffusion.wavelet_dec(2 * xd[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
        t0 = time.perf_counter()
        r = restorer.restore_batch(xd, r=GRID_R, noise=noise.to(dev), x_other=x_other)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # the same call again, timed on the device: throughput of this precision mode at the parity batch (HFRM bypassed)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        nd = noise.to(dev)
        e0.record()
        restorer.restore_batch(xd, r=GRID_R, noise=nd, x_other=x_other)
        e1.record()
        torch.cuda.synchronize()
        ips = batch / (e0.elapsed_time(e1) / 1e3)
        lat = r["latent"].cpu()[list(slots)]
        out = r["output"].cpu()[list(slots)]
        dps = [abs(float(torchPSNR(x[s_:s_ + 1, 3:], out[i:i + 1])) - float(g["psnr"][i])) for i, s_ in enumerate(slots)]
        unsat = (out_ref > 0) & (out_ref < 1)
        tc_n, simt_n = restorer.diffusion.model.module.engine().counters()
        res[prec] = {"latent_max_abs": float((lat - lat_ref).abs().max()),
                     "latent_max_rel": float((lat - lat_ref).abs().max() / lat_ref.abs().max()),
                     "latent_rel_l2": float((lat - lat_ref).norm() / lat_ref.norm()),
                     "image_max_abs": float((out - out_ref).abs().max()),
                     "image_mean_abs": float((out - out_ref).abs().mean()),
                     "image_max_abs_unsaturated_px": float((out - out_ref).abs()[unsat].max()) if unsat.any() else 0.0,
                     "unsaturated_px_frac": float(unsat.float().mean()),
                     "psnr_abs_diff_db": max(dps), "psnr_ref_db": [float(v) for v in g["psnr"]],
                     "tc_launches": tc_n, "simt_launches": simt_n, "seconds": dt, "images_per_s": ips}
        res[prec]["pass"] = bool(res[prec]["image_max_abs"] < 1e-3 and res[prec]["psnr_abs_diff_db"] < 0.01)
        del restorer
        torch.cuda.empty_cache()
    return res


def single_image_latency(eng, betas, dev, steps=50):
    """BASELINE.json configs[0]'s operating point on the GPU (batch 1, 256 x 256 = one 64 x 64 latent patch per DDIM step): the
    sampling loop alone through DdimSampler (gather refresh -> UNet engine with split-K contractions -> DDIM update, replayed
    as a CUDA graph), CUDA events around 3 runs of `steps` steps. Not the headline metric: reported beside it."""
    import torch
    from wavedm_b200.sampler import DdimSampler
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 64, 64, generator=g).to(dev)
    xc = torch.randn(1, 48, 64, 64, generator=g).to(dev)
    xo = torch.randn(1, 45, 64, 64, generator=g).to(dev)
    seq = list(range(0, 1000, 1000 // steps))
    smp = DdimSampler(eng)
    for _ in range(2):
        smp.sample(x, xc, xo, seq, betas, [(0, 0)], 64, keep_history=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3):
        smp.sample(x, xc, xo, seq, betas, [(0, 0)], 64, keep_history=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return {"ms_per_ddim_step": ms / steps, "ms_per_image_sampling": ms, "ddim_steps": steps, "patches_per_step": 1,
            "what": "batch 1, 256x256: DdimSampler.sample alone (graph replay), bf16; HFRM / DWT / IWT not included"}


def gpu_eager_baseline(dev, batch, ddim_sample_steps=3):
    """The reference's own PyTorch path on THIS GPU (SURVEY.md 8d / BASELINE.md 4: "the honest GPU comparator"): the oracle
    port of models/unet.py + the DDIM loop (bit-equal to the reference modules on CPU, oracle/make_golden.py) run with CUDA
    tensors -- PyTorch eager + cuDNN/cuBLAS. A reported baseline like cpu_baseline: bounded sample (`ddim_sample_steps` of
    the 50-step schedule, scaled), three variants: fp32 with torch's default TF32-conv setting, strict fp32 (TF32 off), and
    bf16 autocast; batched the way OUR sampler batches (all images' patches in one UNet call -- kinder to eager than the
    reference's batch-1 loop), plus the reference's real batch-1 semantics for one image."""
    import torch
    from oracle import unet_oracle as O
    cfg = O.default_config()
    sd = {k: v.to(dev) for k, v in O.init_state_dict(cfg, seed=SEED).items()}
    betas = O.beta_schedule(cfg).to(dev)
    seq = O.sampling_seq(1000, 50)
    seq_run = seq[len(seq) - ddim_sample_steps:]
    out = {"what": "oracle port of the reference's PyTorch modules on cuda (eager + cuDNN), same sampler shape as the bench "
                   f"step ({batch} images x 1 patch per UNet call); {ddim_sample_steps} of 50 DDIM steps timed after one warm-up "
                   "pass, scaled to 50", "unit": "images/s", "torch": torch.__version__}

    def run(n_img, mode):
        g = torch.Generator().manual_seed(SEED)
        xc = torch.randn(n_img, 48, 64, 64, generator=g).to(dev)
        xo = torch.randn(n_img, 45, 64, 64, generator=g).to(dev)
        xn = torch.randn(n_img, 3, 64, 64, generator=g).to(dev)

        def model(a, tt):
            if mode == "bf16_autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return O.unet_forward(sd, cfg, a, tt).float()
            return O.unet_forward(sd, cfg, a, tt)
        tf32 = mode != "fp32_strict"
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False if mode == "fp32_strict" else old[1]
        try:
            with torch.no_grad():
                # the oracle's sampler is per image (reference semantics); batch it over images like our sampler does
                def once():
                    xt = xn
                    for i_t in reversed(seq_run):
                        t = torch.full((1,), float(i_t), device=dev)   # one timestep for all patches (ddm_wavelet.py:457)
                        et = model(torch.cat([xc, xt, xo], 1), t)
                        at = O.compute_alpha(betas, torch.full((1,), i_t, device=dev, dtype=torch.long))
                        at_next = O.compute_alpha(betas, torch.full((1,), max(i_t - 20, -1), device=dev, dtype=torch.long))
                        x0 = (xt - et * (1 - at).sqrt()) / at.sqrt()
                        xt = at_next.sqrt() * x0 + (1 - at_next).sqrt() * et
                    return xt
                once()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                once()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / 1e3
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old

    for mode in ("fp32_tf32conv_default", "fp32_strict", "bf16_autocast"):
        try:
            t = run(batch, mode)
            out[mode] = {"value": batch * (ddim_sample_steps / 50.0) / t, "ms_per_unet_call": t / ddim_sample_steps * 1e3,
                         "images_per_call": batch}
        except Exception as e:  # an OOM of the eager path must not take the bench line down
            out[mode] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
    try:
        t1 = run(1, "fp32_tf32conv_default")
        out["batch1_fp32_tf32conv_default"] = {"value": (ddim_sample_steps / 50.0) / t1, "ms_per_unet_call": t1 / ddim_sample_steps * 1e3,
                                               "note": "the reference's real operating point: one image per UNet call (ddm_wavelet.py:486)"}
    except Exception as e:
        out["batch1_fp32_tf32conv_default"] = {"error": repr(e)[:200]}
    del sd
    torch.cuda.empty_cache()
    return out


def dram_traffic_per_launch():
    """roofline.traffic: dram__bytes_read + write per contraction launch from the committed ncu capture of ONE UNet call,
    valid only for the build it was taken on. Builds are identified by the hash of the library SOURCES (`_lib.source_id()`:
    nvcc output is not byte-reproducible, so a rebuilt .so of the same sources has another file hash); captures of older
    rounds that only recorded the .so hash match on that."""
    import hashlib
    from wavedm_b200 import _lib
    try:
        with open(os.path.join(REPO, "wavedm_b200", "libwavedm_b200.so"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()[:16]
    except Exception:
        sha = None
    try:
        src = _lib.source_id()
    except Exception:
        src = None
    best = None
    pdir = os.path.join(REPO, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if name.endswith("_traffic.json"):
            try:
                with open(os.path.join(pdir, name)) as f:
                    d = json.load(f)
            except Exception:
                continue
            d["file"] = "profiles/" + name
            if (src and d.get("src_sha16") == src) or (sha and d.get("lib_sha16") == sha):
                return d["traffic_bytes_per_launch"], {"file": d["file"], "src_sha16": d.get("src_sha16"),
                                                       "lib_sha16": d.get("lib_sha16"), "same_build": True}
            best = d
    if best is not None:
        return None, {"file": best["file"], "src_sha16": best.get("src_sha16"), "lib_sha16": best.get("lib_sha16"),
                      "same_build": False, "stale_value": best.get("traffic_bytes_per_launch"), "this_src_sha16": src,
                      "this_lib_sha16": sha}
    return None, {"this_src_sha16": src, "this_lib_sha16": sha}


def dwt_roofline(dev, peak):
    import torch
    from wavedm_b200 import _lib
    lib = _lib.load()
    H = W = 256  # the DWT half of the metric is quoted at 256x256 whatever --config says
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
    global BYPASS_HFRM
    BYPASS_HFRM = bool(args.bypass_hfrm or args.wavelet_in_unet)
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
            return restorer.restore_batch(x_dev, r=GRID_R, noise=noise_dev)["output"]
        if not args.bypass_hfrm:   # the real restore() path: HFRM engine -> DWT -> x_other / output HF bands
            return restorer.restore_batch(x_dev, r=GRID_R, noise=noise_dev)["output"]
        xo = restorer.diffusion.wavelet_dec(2 * x_dev[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
        res = restorer.restore_batch(x_dev, r=GRID_R, noise=noise_dev, x_other=xo)
        return res["output"]

    # the final gather of restored images: one all-gather into a preallocated [world*B, 3, H, W] tensor --
    # every rank receives in parallel over NVSwitch (a gather to rank 0 serialises 7 receives on one GPU)
    gathered = torch.empty(world * B, 3, H, W, device=dev) if world > 1 else None

    def step_full():
        out = step_device()
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.contiguous())
        return out

    out_pin = torch.empty(B, 3, H, W).pin_memory()

    def step_e2e():
        xh = x_pin.to(dev, non_blocking=True)
        nh = noise_pin.to(dev, non_blocking=True)
        if args.wavelet_in_unet:
            out = restorer.restore_batch(xh, r=GRID_R, noise=nh)["output"]
        elif not args.bypass_hfrm:
            out = restorer.restore_batch(xh, r=GRID_R, noise=nh)["output"]
        else:
            xo = restorer.diffusion.wavelet_dec(2 * xh[:, 3:].contiguous() - 1.0)[:, 3:].contiguous()
            out = restorer.restore_batch(xh, r=GRID_R, noise=nh, x_other=xo)["output"]
        if world > 1:
            dist.all_gather_into_tensor(gathered, out.contiguous())
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
    n_patch_calls = B * patches_per_image() * args.ddim_steps
    alg_fl = UNET_GFLOP_PER_PATCH * 1e9 * n_patch_calls
    ach_tf = alg_fl / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
    exe_tf = k_fl / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
    traffic, traffic_src = dram_traffic_per_launch()
    roof = {"kernel": kname, "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": ach_tf / peak_tf, "traffic": traffic, "traffic_src": traffic_src,
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

    # the HFRM engine alone (once per image, before the sampling loop): CUDA events around K calls on this batch
    hfrm = None
    if not BYPASS_HFRM:
        gen = restorer.diffusion.generator
        cond01 = x_dev[:, :3].contiguous()
        gen(cond01)
        torch.cuda.synchronize()
        l1 = lib.wdm_launch_counter()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            gen(cond01)
        e1.record()
        torch.cuda.synchronize()
        h_ms = e0.elapsed_time(e1) / 3
        hfrm = {"ms_per_batch": h_ms, "images": B, "share_of_step": h_ms / ms_per_step,
                "launches_per_batch": int((lib.wdm_launch_counter() - l1) // 3), "precision": args.precision,
                "kernel": "hfrm_pw_kernel / hfrm_dw_gate_kernel (csrc/wdm_hfrm.cu)",
                "note": "models/arch.py HFRM(dim 32, enc [2,2,2,4], mid 6, dec [2,2,2,2]) on the [B,3,H,W] conditioning image"}

    out = None
    if rank == 0:
        rdwt = dwt_roofline(dev, peaks["hbm_gbs"])
        cpu = parity = gpu_base = None
        # baselines and the parity block run at N = 1 only: at N > 1 the other ranks would spin in the closing barrier on the
        # host cores the CPU sample is timed on (and rank 0's GPU would idle through it in the driver's utilisation record)
        single = None
        if world == 1 and not args.wavelet_in_unet and args.config == 2 and args.precision == "bf16":
            single = single_image_latency(eng, restorer.diffusion.betas, dev)
        if world == 1 and not args.wavelet_in_unet:
            del restorer, eng
            torch.cuda.empty_cache()
            if not args.no_parity:
                parity = parity_block(dev)
            if not args.no_gpu_baseline:
                gpu_base = gpu_eager_baseline(dev, B if args.config == 2 else 64)
            if not args.no_cpu_baseline:
                cpu = cpu_baseline(args.cpu_sample_seconds, os.cpu_count() or 1)
        out = {"metric": metric_name(), "value": value, "unit": "images/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
               "config": workload_config(args, torch.__version__), "clocks": clk,
               "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(x_pin.numel() * 4 + noise_pin.numel() * 4) * world,
                       "d2h_bytes_per_step": int(B * 3 * H * W * 4) * world, "ms_per_step": ms_e2e,
                       "note": "every rank copies its own inputs H2D from pinned memory and its restored images D2H into pinned memory"},
               "gpu_launches": int(launches), "roofline": roof, "roofline_dwt": rdwt, "cpu_baseline": cpu,
               "parity": parity, "gpu_eager_baseline": gpu_base, "hfrm": hfrm, "single_image": single}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
