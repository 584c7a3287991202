"""Batched DDIM (eta = 0) overlapping-patch sampler on the CUDA engine.

This is the device-side restatement of ``models/ddm_wavelet.py:437-506`` (and of the whole-image twin
``utils/sampling.py:23-44``): per timestep ONE gather kernel per patch chunk (crop + concat + NHWC),
the UNet engine, and ONE fused scatter-average + DDIM-update kernel; ``xs`` / ``x0_preds`` stay on the
device and are copied out once at the end (the reference does two D2H copies and four ``.item()`` syncs
per step). The reference's sampler is batch-1 only (it writes ``et_output[0]``, ddm_wavelet.py:486); here
every image of the batch is an independent reference run.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .engine import UNetEngine


def alpha_table(betas: torch.Tensor) -> torch.Tensor:
    """``compute_alpha`` (utils/sampling.py:10-13) for every t in [-1, T): float32 cumprod in the same
    (sequential, CPU) order; entry [t+1] is alpha_bar_t and entry [0] = 1."""
    b = betas.detach().to("cpu", torch.float32)
    return (1 - torch.cat([torch.zeros(1), b], dim=0)).cumprod(dim=0)


def make_patch_table(n_images: int, corners: Sequence[Tuple[int, int]], device) -> Tuple[torch.Tensor, torch.Tensor]:
    """[(image, hi, wi)] sorted by image with the reference's corner order inside an image, + ranges."""
    rows = [(b, int(hi), int(wi)) for b in range(n_images) for (hi, wi) in corners]
    patches = torch.tensor(rows, dtype=torch.int32).reshape(-1, 3).to(device)
    first = (torch.arange(n_images + 1, dtype=torch.int32) * len(corners)).to(device)
    return patches, first


class _StepGraph:
    """One captured CUDA graph of (patch gather -> UNet forward) over persistent buffers: a P <= 16 UNet call is ~180
    dependent launches of a few microseconds each; replaying the captured sequence takes the host (ctypes calls,
    tensor-map encoding) out of the loop. Measured gain 3 % (2.43 -> 2.35 ms per step at P = 1): the small-batch call is
    bound by the serial depth of the K loops, not by the launch rate. Cached on the engine per problem
    shape; bit-identical to eager launches."""

    def __init__(self, eng, x_cond, x, x_other, patches):
        dev = eng.device
        self.x_cond = x_cond.clone()
        self.xt = x.clone()
        self.x_other = None if x_other is None else x_other.clone()
        self.patches = patches.clone()
        P = patches.shape[0]
        self.t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.xin = torch.empty((P, eng.R, eng.R, eng.cin_pad), dtype=eng.dtype, device=dev)
        self.eps = torch.empty((P, eng.out_ch, eng.patch, eng.patch), dtype=torch.float32, device=dev)
        self.eng = eng
        self.srcs = [self.x_cond, self.xt] + ([self.x_other] if self.x_other is not None else [])
        # x_cond / x_other are loop invariants of the DDIM loop: gathered once per image batch (gather_invariants), the
        # captured step only rewrites the channels of x_t. (wavelet_in_unet gathers DWT coefficients: whole gather per step.)
        self.partial = not eng.wavelet_in_unet

        def body():
            if self.partial:
                eng.gather_update(self.xt, self.x_cond.shape[1], self.patches, self.xin)
            else:
                eng.gather(self.srcs, self.patches, out=self.xin)
            eng.forward_nhwc(self.xin, self.t, out=self.eps)
        self.gather_invariants()
        with torch.cuda.device(dev):    # capture on the engine's device whatever the caller's current device is
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):   # warm-up outside capture: workspace allocation, function attributes, descriptors
                body()
            cur.wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                body()

    def load(self, x_cond, x, x_other, patches):
        self.x_cond.copy_(x_cond)
        self.xt.copy_(x)
        if self.x_other is not None:
            self.x_other.copy_(x_other)
        self.patches.copy_(patches)
        self.gather_invariants()

    def gather_invariants(self):
        if self.partial:
            self.eng.gather(self.srcs, self.patches, out=self.xin)


class DdimSampler:
    #: patch counts up to this run the per-step (gather, UNet) pair as a replayed CUDA graph
    GRAPH_MAX_PATCHES = 16
    #: captured graphs (with their persistent buffers) kept per engine, least recently used evicted
    GRAPH_CACHE_ENTRIES = 4

    def __init__(self, engine: UNetEngine, max_patches: Optional[int] = None, use_graph: Optional[bool] = None,
                 mirror_rng: bool = True):
        self.engine = engine
        self.max_patches = int(max_patches or engine.max_patches)
        self.use_graph = use_graph
        #: the reference evaluates ``c1 * torch.randn_like(x)`` with c1 = 0 at every step (ddm_wavelet.py:502): the value
        #: is irrelevant but the device's default generator advances, which decides the initial noise of the NEXT image
        #: (restoration.py:177). Drawing the same throw-away tensor keeps seeded multi-image runs in step with it.
        self.mirror_rng = mirror_rng

    @torch.no_grad()
    def sample(self, x: torch.Tensor, x_cond: torch.Tensor, x_other: Optional[torch.Tensor], seq: Sequence[int],
               betas: torch.Tensor, corners: Sequence[Tuple[int, int]], p_size: int, eta: float = 0.0,
               keep_history: bool = True, keep_last: Optional[int] = None):
        """Returns (xs_hist [S, B, C, h, w], x0_hist [S, B, C, h, w]) device tensors (history of x_{t-1} and
        of the x0 predictions, in sampling order). With keep_history=False only the last entries are kept
        ([1, B, C, h, w]); with keep_last=n only the last n steps ([min(n, S), ...], so that ``x0_hist[-5]`` -- the element
        restoration.py:108 reads -- needs 5 slots instead of S)."""
        eng = self.engine
        if eta != 0.0:
            raise NotImplementedError("only eta = 0 (deterministic DDIM) is implemented; the reference never "
                                      "passes another value (ddm_wavelet.py:302)")
        if p_size != eng.patch:
            raise ValueError(f"patch size {p_size} != the engine's patch side {eng.patch}")
        dev = eng.device
        x = x.to(dev, torch.float32).contiguous()
        x_cond = x_cond.to(dev, torch.float32).contiguous()
        srcs_tail = []
        if x_other is not None:
            srcs_tail = [x_other.to(dev, torch.float32).contiguous()]
        B, Cp, h, w = x.shape
        if x_cond.shape[1] + Cp + (srcs_tail[0].shape[1] if srcs_tail else 0) != eng.in_channels:
            raise ValueError("x_cond / x / x_other channel counts do not add up to the UNet's input channels")
        seq = list(seq)
        seq_next = [-1] + seq[:-1]
        S = len(seq)
        alphas = alpha_table(betas)
        patches, first = make_patch_table(B, corners, dev)
        P = patches.shape[0]
        tvals = torch.tensor(list(reversed(seq)), dtype=torch.float32).to(dev)
        nh = S if keep_history else 1
        if keep_last is not None and keep_history:
            nh = max(1, min(int(keep_last), S))
        xs_hist = torch.empty((nh, B, Cp, h, w), dtype=torch.float32, device=dev)
        x0_hist = torch.empty((nh, B, Cp, h, w), dtype=torch.float32, device=dev)
        eps = torch.empty((P, eng.out_ch, eng.patch, eng.patch), dtype=torch.float32, device=dev)
        chunk = min(self.max_patches, P)
        use_graph = (P <= self.GRAPH_MAX_PATCHES) if self.use_graph is None else bool(self.use_graph)
        if use_graph and P <= chunk and not getattr(eng, "_profiling", False):
            return self._sample_graph(x, x_cond, srcs_tail[0] if srcs_tail else None, seq, seq_next, alphas, patches, first,
                                      tvals, xs_hist, x0_hist, keep_history)
        xin = torch.empty((chunk, eng.R, eng.R, eng.cin_pad), dtype=eng.dtype, device=dev)
        xt = x
        # one chunk holds every patch: the gathered tensor stays resident over the DDIM loop and steps after the first only
        # rewrite the x_t channels (x_cond / x_other do not change: 6 instead of 256 B per pixel per step)
        resident = P <= chunk and not eng.wavelet_in_unet
        for k, (i_t, j_t) in enumerate(zip(reversed(seq), reversed(seq_next))):
            at = float(alphas[i_t + 1])
            at_next = float(alphas[j_t + 1])
            t = tvals[k:k + 1]
            for p0 in range(0, P, chunk):
                n = min(chunk, P - p0)
                if k > 0 and resident:
                    eng.gather_update(xt, x_cond.shape[1], patches, xin)
                else:
                    eng.gather([x_cond, xt] + srcs_tail, patches[p0:p0 + n], out=xin[:n])
                eng.forward_nhwc(xin[:n], t, out=eps[p0:p0 + n])
            slot = self._slot(k, S, nh)
            eng.ddim_step(eps, patches, first, xt, x0_hist[slot], xs_hist[slot], at, at_next)
            if self.mirror_rng:
                torch.randn_like(xt)
            xt = xs_hist[slot]
        return xs_hist, x0_hist

    @staticmethod
    def _slot(k: int, S: int, nh: int) -> int:
        """History slot of step k when only the last nh of S steps are kept: the kept window is written in order, the steps
        before it rotate through the same slots (each is overwritten before anyone reads it; x_t of step k lives in the
        slot step k - 1 wrote, which differs from step k's own slot whenever nh >= 2; nh == 1 updates in place)."""
        return k if nh == S else (k - (S - nh)) % nh

    def _sample_graph(self, x, x_cond, x_other, seq, seq_next, alphas, patches, first, tvals, xs_hist, x0_hist, keep_history):
        eng = self.engine
        key = (tuple(x.shape), tuple(x_cond.shape), None if x_other is None else tuple(x_other.shape), patches.shape[0])
        cache = eng.__dict__.setdefault("_step_graphs", {})
        sg = cache.pop(key, None)
        if sg is None:
            while len(cache) >= self.GRAPH_CACHE_ENTRIES:   # least recently used shape goes first
                cache.pop(next(iter(cache)))
            sg = _StepGraph(eng, x_cond, x, x_other, patches)
        else:
            sg.load(x_cond, x, x_other, patches)
        cache[key] = sg
        for k, (i_t, j_t) in enumerate(zip(reversed(seq), reversed(seq_next))):
            sg.t.copy_(tvals[k:k + 1])
            sg.graph.replay()
            slot = self._slot(k, len(seq), x0_hist.shape[0])
            eng.ddim_step(sg.eps, sg.patches, first, sg.xt, x0_hist[slot], xs_hist[slot], float(alphas[i_t + 1]),
                          float(alphas[j_t + 1]))
            if self.mirror_rng:
                torch.randn_like(sg.xt)
            sg.xt.copy_(xs_hist[slot])
        return xs_hist, x0_hist

    @torch.no_grad()
    def sample_lists(self, x, x_cond, x_other, seq, betas, corners, p_size, eta=0.0):
        """Reference return contract (ddm_wavelet.py:449,498,503,506): ``xs`` = [x] + S CPU tensors,
        ``x0_preds`` = S CPU tensors."""
        xs_hist, x0_hist = self.sample(x, x_cond, x_other, seq, betas, corners, p_size, eta=eta)
        xs_cpu = xs_hist.to("cpu")
        x0_cpu = x0_hist.to("cpu")
        xs: List[torch.Tensor] = [x] + [xs_cpu[i] for i in range(xs_cpu.shape[0])]
        x0_preds: List[torch.Tensor] = [x0_cpu[i] for i in range(x0_cpu.shape[0])]
        return xs, x0_preds
