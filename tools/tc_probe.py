"""Times single conv shapes through wdm_gemm (tensor-core path). WDM_TC_DBG=1|2 and WDM_TC_PAIR=0|1 select probes."""
import ctypes, os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from wavedm_b200 import _lib
from test_unet_gpu import GemmParams
DEV = torch.device("cuda", 0)
lib = _lib.load()
def run(P, C, Cout, H, taps, iters=20):
    x = torch.randn(P, H, H, C, device=DEV).bfloat16()
    w = (torch.randn(Cout, taps * C, device=DEV) * 0.02).bfloat16()
    out = torch.empty(P, H, H, Cout, device=DEV, dtype=torch.bfloat16)
    p = GemmParams()
    p.src0, p.C0, p.ld0 = x.data_ptr(), C, C
    p.Hin = p.Win = p.Hout = p.Wout = H
    p.taps, p.stride, p.pad = taps, 1, 1 if taps == 9 else 0
    p.B, p.ldb = w.data_ptr(), taps * C
    p.M, p.N, p.K = P * H * H, Cout, taps * C
    p.alpha = 1.0
    p.out, p.ldo = out.data_ptr(), Cout
    p.a_dtype = p.b_dtype = p.out_dtype = 1
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): assert lib.wdm_gemm(ctypes.byref(p), 1, st) == 0
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): lib.wdm_gemm(ctypes.byref(p), 1, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * p.M * p.N * p.K
    print(f"P={P} C={C}->{Cout} @{H}x{H} taps={taps}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  dbg={os.environ.get('WDM_TC_DBG','0')} pair={os.environ.get('WDM_TC_PAIR','1')}")
for shape in [(64, 256, 256, 32, 9), (64, 512, 512, 16, 9), (64, 768, 768, 8, 9), (64, 128, 128, 64, 9), (64, 512, 512, 16, 1)]:
    run(*shape)
