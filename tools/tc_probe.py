"""Times single conv / 1x1 shapes through wdm_gemm (tensor-core path). Probes (env): WDM_TC_DBG=1 no TMA loads, 2 no MMAs,
3 no epilogue stores, 5 no epilogue; WDM_TC_PAIR=0|1. `full` rows add bias + residual + GroupNorm side-car statistics."""
import ctypes, os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from wavedm_b200 import _lib
from test_unet_gpu import GemmParams
DEV = torch.device("cuda", 0)
lib = _lib.load()
def run(P, C, Cout, H, taps, full=0, iters=20):
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
    keep = []
    if full:
        b = torch.randn(Cout, device=DEV); r = torch.randn(P, H, H, Cout, device=DEV).bfloat16()
        st = torch.empty(p.M // 32 * (Cout // 4) * 2, device=DEV)
        keep = [b, r, st]
        p.bias, p.residual, p.ldr, p.stats_out = b.data_ptr(), r.data_ptr(), Cout, st.data_ptr()
    st_ = torch.cuda.current_stream().cuda_stream
    for _ in range(3): assert lib.wdm_gemm(ctypes.byref(p), 1, st_) == 0
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): lib.wdm_gemm(ctypes.byref(p), 1, st_)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * p.M * p.N * p.K
    print(f"P={P} C={C}->{Cout} @{H}x{H} taps={taps} full={full}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  "
          f"dbg={os.environ.get('WDM_TC_DBG','0')} pair={os.environ.get('WDM_TC_PAIR','1')}")
SHAPES = [(64, 256, 256, 32, 9, 0), (64, 512, 512, 16, 9, 0), (64, 512, 512, 16, 9, 1), (64, 768, 768, 8, 9, 0), (64, 128, 128, 64, 9, 0),
          (64, 128, 128, 64, 9, 1), (64, 512, 512, 16, 1, 0), (64, 512, 512, 16, 1, 1), (64, 512, 1024, 16, 1, 0)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    SHAPES = SHAPES[-3:] + [SHAPES[5]]
if len(sys.argv) > 1 and sys.argv[1] == "ncu":   # two launches each of the two epilogue-heavy shapes
    for shape in (SHAPES[7], SHAPES[5]):
        run(*shape, iters=1)
    sys.exit(0)
for shape in SHAPES:
    run(*shape)
