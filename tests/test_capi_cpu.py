"""CPU: the C-ABI shared library loads and exports every symbol include/wavedm_b200.h declares, and argument
validation (which happens before any CUDA call) behaves as documented. No compute calls here."""
import ctypes

import pytest

from wavedm_b200 import _lib


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 5 and len(set(names)) == len(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wavedm_b200.h but not exported"
    # every exported entry point has a ctypes signature bound in _lib.py
    assert set(names) == set(lib._wdm_sigs.keys())
    assert lib.wdm_build_arch() == b"sm_100a"
    assert lib.wdm_version() >= 1
    assert lib.wdm_status_string(0) == b"ok"
    assert b"shape" in lib.wdm_status_string(_lib.WDM_ERR_BAD_SHAPE)


def test_argument_validation_without_a_device():
    lib = _lib.load()
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    p16 = (p + 15) & ~15
    assert lib.wdm_dwt4x4_fwd(None, p16, 1, 8, 8, 0, None) == _lib.WDM_ERR_BAD_ARG
    assert lib.wdm_dwt4x4_fwd(p16, p16, 1, 6, 8, 0, None) == _lib.WDM_ERR_BAD_SHAPE
    assert lib.wdm_dwt4x4_fwd(p16, p16, 1, 8, 10, 0, None) == _lib.WDM_ERR_BAD_SHAPE
    assert lib.wdm_dwt4x4_fwd(p16 + 4, p16, 1, 8, 8, 0, None) == _lib.WDM_ERR_BAD_ALIGN
    assert lib.wdm_dwt4x4_fwd(p16, p16, 1, 8, 8, 0x4, None) == _lib.WDM_ERR_BAD_ARG
    assert lib.wdm_dwt4x4_fwd(p16, p16, 0, 8, 8, 0, None) == _lib.WDM_OK  # empty batch: nothing to launch
    assert lib.wdm_iwt4x4_fwd(p16, p16, 0, 2, 2, 0, None) == _lib.WDM_OK
    assert lib.wdm_iwt4x4_fwd(p16, None, 1, 2, 2, 0, None) == _lib.WDM_ERR_BAD_ARG
    # TMA variant refuses shapes it cannot tile instead of silently switching
    assert lib.wdm_dwt4x4_fwd(p16, p16, 1, 8, 8, _lib.WDM_WT_IMPL_TMA, None) == _lib.WDM_ERR_UNSUPPORTED


def test_no_cpu_fallback():
    import torch
    from wavedm_b200.wavelet import WaveletTransform
    m = WaveletTransform(scale=2, dec=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 8, 8))
    with pytest.raises(NotImplementedError):
        WaveletTransform(scale=1, dec=True)(torch.zeros(1, 3, 8, 8))


def test_unet_config_validation_and_parameter_table_without_a_device():
    """wdm_unet_param_count / _info only build the host-side model description: the 332 reference state-dict tensors
    (+ temb.freqs) for raindrop_wavelet.yml, the same table for data.wavelet_in_unet (the frozen wavelet filters are not
    engine parameters), and loud refusal of inconsistent configs."""
    from wavedm_b200 import engine
    from wavedm_b200.configs import default_config
    lib = _lib.load()
    cfg = default_config()
    cs = engine.make_config_struct(cfg)
    assert cs.in_channels == 96 and cs.out_ch == 3 and cs.wavelet_in_unet == 0
    table = engine.param_table(cs)
    assert len(table) == 333 and table[-1][0] == "temb.freqs" and table[0][0] == "temb.dense.0.weight"
    assert sum(n for name, n in table if name != "temb.freqs") == 156492675
    cfg.data.wavelet_in_unet = True
    cfg.model.use_other_channels, cfg.model.in_channels, cfg.model.out_ch = False, 93, 48
    cw = engine.make_config_struct(cfg)
    assert cw.wavelet_in_unet == 1 and cw.in_channels == 96 and cw.out_ch == 48
    tw = engine.param_table(cw)
    assert [n for n, _ in tw] == [n for n, _ in table]
    assert dict(tw)["conv_out.weight"] == 48 * 128 * 9
    cw.out_ch = 3   # wavelet_in_unet needs 48 output channels
    assert lib.wdm_unet_param_count(ctypes.byref(cw)) == _lib.WDM_ERR_BAD_ARG
    cs.out_ch = 65  # and the plain mode at most 64 (out_ch = 3 p^2 of data.use_window)
    assert lib.wdm_unet_param_count(ctypes.byref(cs)) == _lib.WDM_ERR_BAD_ARG
    # the fused restore() epilogue kernel validates shapes before touching the device
    buf = (ctypes.c_float * 64)()
    p16 = (ctypes.addressof(buf) + 15) & ~15
    assert lib.wdm_iwt4x4_cat(p16, 3, p16, 44, p16, 1, 2, 2, 0, None) == _lib.WDM_ERR_BAD_SHAPE
    assert lib.wdm_iwt4x4_cat(p16, 3, p16, 45, p16, 0, 2, 2, 0, None) == _lib.WDM_OK
    assert lib.wdm_iwt4x4_nhwc(None, 48, 1, 2, p16, None) == _lib.WDM_ERR_BAD_ARG
