"""ctypes binding of libwavedm_b200.so (the C ABI declared in include/wavedm_b200.h).

There is no CPU fallback: if the shared library is missing, or an entry point returns a non-zero
``wdm_status``, a ``WdmError`` is raised. ``load()`` only dlopens the library (works without a GPU, which
is what the CPU test-suite checks); every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwavedm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "wavedm_b200.h")

# status codes (include/wavedm_b200.h)
WDM_OK = 0
WDM_ERR_BAD_SHAPE = -1
WDM_ERR_BAD_ALIGN = -2
WDM_ERR_BAD_ARG = -3
WDM_ERR_UNSUPPORTED = -4
WDM_ERR_WORKSPACE = -5
WDM_ERR_NO_DEVICE = -6
WDM_ERR_CUDA_BASE = -1000

# flags
WDM_DWT_PRE_2XM1 = 0x1
WDM_IWT_POST_CLAMP = 0x1
WDM_WT_IMPL_AUTO = 0x00
WDM_WT_IMPL_DIRECT = 0x10
WDM_WT_IMPL_TMA = 0x20
WDM_PREC_FP32 = 0
WDM_PREC_BF16 = 1
WDM_ENGINE_NO_TC = 0x1
WDM_ENGINE_ALLOW_SIMT = 0x2
WDM_ENGINE_TC32 = 0x4
WDM_GEMM_IMPL_SIMT = 0
WDM_GEMM_IMPL_TC = 1


class WdmError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        try:
            msg = load().wdm_status_string(status).decode()
        except Exception:  # pragma: no cover
            msg = "?"
        super().__init__(f"{where} failed: wdm_status {status} ({msg})")


_lib = None


def declared_symbols() -> List[str]:
    """Every WDM_API symbol declared in include/wavedm_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    return re.findall(r"WDM_API\s+[\w\s\*]+?\b(wdm_\w+)\s*\(", src)


def source_id() -> str:
    """sha256[:16] over the CUDA / C sources the library is built from (csrc/*.cu, *.cuh, *.h and the C ABI header), in
    name order. The build id measurement files are keyed on: nvcc output is not byte-reproducible, the sources are."""
    import hashlib
    h = hashlib.sha256()
    src = os.path.join(_HERE, "csrc")
    files = sorted(f for f in os.listdir(src) if f.endswith((".cu", ".cuh", ".h")))
    for f in files:
        with open(os.path.join(src, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    with open(HEADER_PATH, "rb") as fh:
        h.update(b"wavedm_b200.h\0" + fh.read())
    return h.hexdigest()[:16]


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C wavedm_b200/csrc`). wavedm_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    c_int, c_void_p, c_char_p, c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t
    c_longlong, c_float = ctypes.c_longlong, ctypes.c_float
    sigs: Dict[str, tuple] = {
        "wdm_version": (c_int, []),
        "wdm_build_arch": (c_char_p, []),
        "wdm_status_string": (c_char_p, [c_int]),
        "wdm_launch_counter": (c_longlong, []),
        "wdm_unet_profile_enable": (c_int, [c_void_p, c_int]),
        "wdm_unet_profile_read": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        "wdm_unet_profile_tc_bytes": (ctypes.c_double, [c_void_p]),
        "wdm_unet_counters": (c_int, [c_void_p, c_void_p, c_void_p]),
        "wdm_dwt4x4_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
        "wdm_iwt4x4_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
        "wdm_iwt4x4_cat": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
        "wdm_unet_param_count": (c_int, [c_void_p]),
        "wdm_unet_param_info": (c_int, [c_void_p, c_int, c_char_p, c_int, c_void_p]),
        "wdm_unet_packed_bytes": (c_size_t, [c_void_p, c_int]),
        "wdm_unet_packed_bytes_flags": (c_size_t, [c_void_p, c_int, c_int]),
        "wdm_unet_create": (c_int, [c_void_p, c_int, c_int, c_void_p, c_longlong, c_void_p, c_size_t, c_void_p,
                                    c_void_p]),
        "wdm_unet_destroy": (None, [c_void_p]),
        "wdm_unet_input_channels_padded": (c_int, [c_void_p]),
        "wdm_unet_workspace_bytes": (c_size_t, [c_void_p, c_int]),
        "wdm_unet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
        "wdm_split3_act": (c_int, [c_void_p, c_int, c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
        "wdm_split3_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
        "wdm_psnr_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
        "wdm_hfrm_param_count": (c_int, [c_void_p]),
        "wdm_hfrm_param_info": (c_int, [c_void_p, c_int, c_char_p, c_int, c_void_p]),
        "wdm_hfrm_packed_bytes": (c_size_t, [c_void_p, c_int]),
        "wdm_hfrm_create": (c_int, [c_void_p, c_int, c_void_p, c_longlong, c_void_p, c_size_t, c_void_p, c_void_p]),
        "wdm_hfrm_destroy": (None, [c_void_p]),
        "wdm_hfrm_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
        "wdm_hfrm_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
        "wdm_optim_chunk": (c_int, []),
        "wdm_adam_ema_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_double, c_longlong, ctypes.c_double, c_void_p]),
        "wdm_gather_patches": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                       c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
        "wdm_gather_patches_update": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                              c_void_p, c_int, c_void_p]),
        "wdm_gather_patches_dwt": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                           c_void_p, c_int, c_void_p]),
        "wdm_iwt4x4_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
        "wdm_ddim_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_float, c_float, c_void_p]),
        "wdm_gemm": (c_int, [c_void_p, c_int, c_void_p]),
        "wdm_gemm_ksplit_plan": (c_int, [c_void_p]),
        "wdm_groupnorm_scratch_bytes": (c_size_t, [c_int]),
        "wdm_groupnorm_silu": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                       c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
        "wdm_groupnorm_silu_sidecar": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p,
                                               c_void_p, c_int, c_void_p, c_void_p]),
        "wdm_softmax_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._wdm_sigs = sigs
    _lib = lib
    return lib


def check(status: int, where: str) -> None:
    if status != WDM_OK:
        raise WdmError(status, where)


def current_stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
