"""Host-side handle of the CUDA UNet engine (``wdm_unet_*`` in include/wavedm_b200.h).

``UNetEngine`` owns the packed weight arena and the per-batch-size workspaces (torch tensors, i.e. device
memory only -- PyTorch is plumbing here) and exposes the three device-level operations the sampler needs:
``gather`` (patch crop + concat + NCHW->NHWC), ``forward`` (the UNet) and ``ddim_step``.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

#: "fp32"      parity mode: fp32 storage, contractions on the tcgen05 kernel through the 3-way bf16 operand split (tc32)
#: "fp32_ffma" the same with every contraction on the CUDA-core FFMA kernel (round-to-nearest accumulation: tightest parity)
#: "bf16"      throughput mode: bf16 storage, tcgen05 contractions
PREC = {"fp32": _lib.WDM_PREC_FP32, "fp32_ffma": _lib.WDM_PREC_FP32, "bf16": _lib.WDM_PREC_BF16}


class WdmUnetConfig(ctypes.Structure):
    _fields_ = [("ch", ctypes.c_int), ("n_levels", ctypes.c_int), ("ch_mult", ctypes.c_int * 8),
                ("num_res_blocks", ctypes.c_int), ("n_attn_res", ctypes.c_int), ("attn_res", ctypes.c_int * 8),
                ("resolution", ctypes.c_int), ("in_channels", ctypes.c_int), ("out_ch", ctypes.c_int),
                ("wavelet_in_unet", ctypes.c_int)]


def unet_in_channels(config) -> int:
    """models/unet.py:212."""
    m = config.model
    if m.use_other_channels:
        return m.in_channels * 2 + m.pred_channels - m.other_channels_begin
    return m.in_channels + m.pred_channels


def make_config_struct(config) -> WdmUnetConfig:
    m = config.model
    c = WdmUnetConfig()
    c.ch = int(m.ch)
    mult = list(m.ch_mult)
    if len(mult) > 8 or len(m.attn_resolutions) > 8:
        raise ValueError("at most 8 levels / attention resolutions")
    c.n_levels = len(mult)
    for i, v in enumerate(mult):
        c.ch_mult[i] = int(v)
    c.num_res_blocks = int(m.num_res_blocks)
    c.n_attn_res = len(m.attn_resolutions)
    for i, v in enumerate(m.attn_resolutions):
        c.attn_res[i] = int(v)
    c.resolution = int(config.data.image_size)
    c.in_channels = unet_in_channels(config)
    c.out_ch = int(m.out_ch)
    c.wavelet_in_unet = 1 if getattr(config.data, "wavelet_in_unet", False) else 0
    return c


def param_table(cstruct: WdmUnetConfig) -> List[Tuple[str, int]]:
    lib = _lib.load()
    n = lib.wdm_unet_param_count(ctypes.byref(cstruct))
    if n < 0:
        raise _lib.WdmError(n, "wdm_unet_param_count")
    out = []
    buf = ctypes.create_string_buffer(256)
    numel = ctypes.c_longlong()
    for i in range(n):
        _lib.check(lib.wdm_unet_param_info(ctypes.byref(cstruct), i, buf, 256, ctypes.byref(numel)), "wdm_unet_param_info")
        out.append((buf.value.decode(), int(numel.value)))
    return out


def temb_freqs(ch: int) -> torch.Tensor:
    """The frequency table exactly as the reference computes it (models/unet.py:19-21), on the host."""
    half = ch // 2
    e = math.log(10000) / (half - 1)
    return torch.exp(torch.arange(half, dtype=torch.float32) * -e)


class UNetEngine:
    def __init__(self, config, state_dict: Dict[str, torch.Tensor], device, precision: str = "bf16", flags: int = 0,
                 max_patches: int = 64):
        # data.use_window is a re-layout DiffusionUNet.forward applies around the engine (the engine sees [P, 6 p^2, R, R])
        if getattr(config.data, "global_attn", False):
            raise NotImplementedError("global_attn (DiffusionUNet_Global) is out of scope")
        if float(getattr(config.model, "dropout", 0.0)) != 0.0:
            pass  # dropout is the identity at inference
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("wavedm_b200.UNetEngine needs a CUDA device (no CPU fallback)")
        # fp32 (parity mode): contractions on the tcgen05 kernel through the 3-way bf16 operand split (WDM_ENGINE_TC32, fp32
        # accumulation in TMEM; storage / GroupNorm / softmax stay fp32). WDM_TC32=0 or flags |= WDM_ENGINE_NO_TC keeps every
        # contraction on the CUDA-core FFMA kernel (the round-1 parity mode).
        if precision == "fp32_ffma":
            flags |= _lib.WDM_ENGINE_NO_TC
        if precision == "fp32" and os.environ.get("WDM_TC32", "1") != "0" and not (flags & _lib.WDM_ENGINE_NO_TC):
            flags |= _lib.WDM_ENGINE_TC32
        self.flags = flags
        self.precision = precision
        self.prec = PREC[precision]
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.cstruct = make_config_struct(config)
        self.R = int(config.data.image_size)
        # wavelet_in_unet (models/unet.py:203-206): the module consumes pixel-domain [P, 6, 4R, 4R] and returns [P, 3, 4R, 4R]
        self.wavelet_in_unet = bool(self.cstruct.wavelet_in_unet)
        self.patch = 4 * self.R if self.wavelet_in_unet else self.R      # side of a sampler patch / of eps
        self.out_ch = 3 if self.wavelet_in_unet else int(config.model.out_ch)
        self.in_channels = 6 if self.wavelet_in_unet else self.cstruct.in_channels  # channels the CALLER concatenates
        self.max_patches = int(max_patches)
        table = param_table(self.cstruct)
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        # the frozen wavelet filters of a wavelet_in_unet module (unet.py:205-206) are not engine parameters
        sd = {k: v for k, v in sd.items() if not k.startswith(("wavelet_dec.", "wavelet_rec."))}
        chunks = []
        for name, numel in table:
            if name == "temb.freqs":
                t = temb_freqs(self.cstruct.ch)
            else:
                if name not in sd:
                    raise KeyError(f"state_dict is missing '{name}'")
                t = sd[name]
            if t.numel() != numel:
                raise ValueError(f"'{name}': expected {numel} elements, got {tuple(t.shape)}")
            chunks.append(t.detach().to(device=self.device, dtype=torch.float32).reshape(-1))
        extra = set(sd) - {n for n, _ in table}
        if extra:
            raise KeyError(f"unexpected keys in state_dict: {sorted(extra)[:5]} ...")
        flat = torch.cat(chunks)
        nbytes = self.lib.wdm_unet_packed_bytes_flags(ctypes.byref(self.cstruct), self.prec, flags)
        if nbytes == 0:
            raise _lib.WdmError(_lib.WDM_ERR_BAD_ARG, "wdm_unet_packed_bytes")
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            st = self.lib.wdm_unet_create(ctypes.byref(self.cstruct), self.prec, flags, flat.data_ptr(), flat.numel(),
                                          self.packed.data_ptr(), nbytes, _lib.current_stream_ptr(self.device),
                                          ctypes.byref(handle))
            _lib.check(st, "wdm_unet_create")
            torch.cuda.current_stream(self.device).synchronize()  # flat may be freed after packing
        self.handle = handle
        del flat
        self.cin_pad = self.lib.wdm_unet_input_channels_padded(self.handle)
        self._ws: Dict[int, torch.Tensor] = {}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.wdm_unet_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ------------------------------------------------------------------------------------------ device ops
    def workspace(self, P: int) -> torch.Tensor:
        ws = self._ws.get(P)
        if ws is None:
            n = self.lib.wdm_unet_workspace_bytes(self.handle, P)
            if n == 0:
                raise _lib.WdmError(_lib.WDM_ERR_BAD_ARG, "wdm_unet_workspace_bytes")
            ws = torch.empty(n + 1024, dtype=torch.uint8, device=self.device)
            self._ws[P] = ws
        return ws

    @staticmethod
    def _aligned_ptr(t: torch.Tensor, a: int = 1024) -> int:
        return (t.data_ptr() + a - 1) // a * a

    def gather(self, srcs: Sequence[torch.Tensor], patches: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """srcs: up to three fp32 NCHW tensors [B, Cs, h, w]; patches: int32 device [P, 3] (image, hi, wi).
        Returns the UNet input [P, R, R, cin_pad] in the engine dtype."""
        P = patches.shape[0]
        B, _, h, w = srcs[0].shape
        if out is None:
            out = torch.empty((P, self.R, self.R, self.cin_pad), dtype=self.dtype, device=self.device)
        if self.wavelet_in_unet:
            # crop (pixel coordinates, side 4R) + DWT of every 3-channel source + concat + NHWC in one kernel
            assert 1 <= len(srcs) <= 2
            for s_ in srcs:
                assert s_.is_cuda and s_.dtype == torch.float32 and s_.is_contiguous() and s_.shape == (B, 3, h, w)
            with torch.cuda.device(self.device):
                st = self.lib.wdm_gather_patches_dwt(srcs[0].data_ptr(), srcs[1].data_ptr() if len(srcs) > 1 else 0,
                                                     len(srcs), B, h, w, patches.data_ptr(), P, self.R, self.cin_pad,
                                                     out.data_ptr(), self.prec, _lib.current_stream_ptr(self.device))
            _lib.check(st, "wdm_gather_patches_dwt")
            return out
        ptr = [0, 0, 0]
        cs = [0, 0, 0]
        for i, s in enumerate(srcs):
            assert s.is_cuda and s.dtype == torch.float32 and s.is_contiguous() and s.shape[0] == B and s.shape[2:] == (h, w)
            ptr[i], cs[i] = s.data_ptr(), s.shape[1]
        with torch.cuda.device(self.device):
            st = self.lib.wdm_gather_patches(ptr[0], cs[0], ptr[1], cs[1], ptr[2], cs[2], B, h, w, patches.data_ptr(), P,
                                             self.R, self.cin_pad, out.data_ptr(), self.prec,
                                             _lib.current_stream_ptr(self.device))
        _lib.check(st, "wdm_gather_patches")
        return out

    def gather_update(self, src: torch.Tensor, c_off: int, patches: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Rewrites the channels [c_off, c_off + C) of a tensor ``gather`` produced from ``src`` [B, C, h, w] (the sampler's
        x_t between DDIM steps; the conditioning channels are loop invariants of ddm_wavelet.py:467-478)."""
        if self.wavelet_in_unet:
            raise RuntimeError("gather_update: the wavelet_in_unet input is a DWT of the sources, gather it whole")
        P = patches.shape[0]
        B, C, h, w = src.shape
        assert src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()
        assert out.dtype == self.dtype and out.is_contiguous() and out.shape == (P, self.R, self.R, self.cin_pad)
        with torch.cuda.device(self.device):
            st = self.lib.wdm_gather_patches_update(src.data_ptr(), C, int(c_off), B, h, w, patches.data_ptr(), P, self.R,
                                                    self.cin_pad, out.data_ptr(), self.prec,
                                                    _lib.current_stream_ptr(self.device))
        _lib.check(st, "wdm_gather_patches_update")
        return out

    def forward_nhwc(self, x: torch.Tensor, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: [P, R, R, cin_pad] engine dtype; t: fp32 device [1] or [P]. Returns eps [P, out_ch, patch, patch] fp32
        (patch = R, or 4R in wavelet_in_unet mode where the IWT is applied to the 48-channel conv_out result)."""
        P = x.shape[0]
        assert x.dtype == self.dtype and x.is_contiguous() and x.shape[1:] == (self.R, self.R, self.cin_pad)
        assert t.dtype == torch.float32 and t.is_cuda and t.numel() in (1, P)
        if out is None:
            out = torch.empty((P, self.out_ch, self.patch, self.patch), dtype=torch.float32, device=self.device)
        ws = self.workspace(P)
        wptr = self._aligned_ptr(ws)
        with torch.cuda.device(self.device):
            st = self.lib.wdm_unet_forward(self.handle, x.data_ptr(), t.data_ptr(), t.numel(), P, out.data_ptr(), wptr,
                                           ws.numel() - (wptr - ws.data_ptr()), _lib.current_stream_ptr(self.device))
        _lib.check(st, "wdm_unet_forward")
        return out

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Module-level entry: x [P, Cin, R, R] fp32 NCHW, t [1] or [P] (models/unet.py:346)."""
        P = x.shape[0]
        assert x.shape[1] == self.in_channels and x.shape[2] == x.shape[3] == self.patch
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        outs = []
        for i in range(0, P, self.max_patches):
            xc = x[i:i + self.max_patches]
            n = xc.shape[0]
            patches = torch.zeros((n, 3), dtype=torch.int32, device=self.device)
            patches[:, 0] = torch.arange(n, dtype=torch.int32, device=self.device)
            xin = self.gather([xc[:, :3].contiguous(), xc[:, 3:].contiguous()] if self.wavelet_in_unet else [xc], patches)
            tc = t if t.numel() == 1 else t[i:i + n].contiguous()
            outs.append(self.forward_nhwc(xin, tc))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def ddim_step(self, eps: torch.Tensor, patches: torch.Tensor, img_first: torch.Tensor, xt: torch.Tensor,
                  x0_out: torch.Tensor, xt_next: torch.Tensor, at: float, at_next: float) -> None:
        P = eps.shape[0]
        B, Cp, h, w = xt.shape
        with torch.cuda.device(self.device):
            st = self.lib.wdm_ddim_step(eps.data_ptr(), patches.data_ptr(), img_first.data_ptr(), P, B, Cp, self.patch, h, w,
                                        xt.data_ptr(), x0_out.data_ptr(), xt_next.data_ptr(), ctypes.c_float(at),
                                        ctypes.c_float(at_next), _lib.current_stream_ptr(self.device))
        _lib.check(st, "wdm_ddim_step")

    def counters(self):
        """(tensor-core, CUDA-core) contraction launches since the engine was created."""
        a, b = ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(self.lib.wdm_unet_counters(self.handle, ctypes.byref(a), ctypes.byref(b)), "wdm_unet_counters")
        return int(a.value), int(b.value)

    # ------------------------------------------------------------------------------------------ profiling
    def profile(self, on: bool) -> None:
        self._profiling = bool(on)  # event-bracketed launches cannot be captured into a CUDA graph
        _lib.check(self.lib.wdm_unet_profile_enable(self.handle, 1 if on else 0), "wdm_unet_profile_enable")

    def profile_read(self):
        """(tc_ms, tc_flops, tc_launches, simt_ms, simt_flops, simt_launches) since the last read."""
        d = [ctypes.c_double() for _ in range(4)]
        n = [ctypes.c_longlong() for _ in range(2)]
        st = self.lib.wdm_unet_profile_read(self.handle, ctypes.byref(d[0]), ctypes.byref(d[1]), ctypes.byref(n[0]),
                                            ctypes.byref(d[2]), ctypes.byref(d[3]), ctypes.byref(n[1]))
        _lib.check(st, "wdm_unet_profile_read")
        self.last_tc_bytes = float(self.lib.wdm_unet_profile_tc_bytes(self.handle))
        return d[0].value, d[1].value, n[0].value, d[2].value, d[3].value, n[1].value
