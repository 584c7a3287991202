"""Does the GroupNorm normalise pass run faster when its input is still in L2? (stats + apply through wdm_groupnorm_silu on
bf16 [P, 4096, 128] tensors; `hot`: the tensor was written by a device copy right before, `cold`: a 512 MB fill in between)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavedm_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
C, HW = 128, 4096
gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
thrash = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for P in (16, 27, 37, 64):
    src = torch.randn(P, HW, C, device=dev).bfloat16()
    x = torch.empty_like(src)
    out = torch.empty_like(src)
    scratch = torch.empty(lib.wdm_groupnorm_scratch_bytes(P), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for mode in ("hot", "cold"):
        tot = 0.0
        n = 20
        for i in range(n + 3):
            x.copy_(src)
            if mode == "cold":
                thrash.fill_(i & 255)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            rc = lib.wdm_groupnorm_silu(x.data_ptr(), C, None, 0, 1, P, HW, 1e-6, gamma.data_ptr(), beta.data_ptr(), 1, out.data_ptr(),
                                        scratch.data_ptr(), st)
            e1.record()
            assert rc == 0
            torch.cuda.synchronize()
            if i >= 3:
                tot += e0.elapsed_time(e1)
        mb = x.numel() * 2 / 1e6
        print(f"P={P} tensor {mb:.1f} MB {mode}: {tot / n * 1e3:.1f} us (stats + apply)")
