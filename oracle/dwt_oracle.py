"""ctypes wrapper of oracle/dwt_oracle.c -- TEST INFRASTRUCTURE ONLY (see the header of that file)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libdwt_oracle.so")
        if not os.path.isfile(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dwt(x: np.ndarray, flags: int = 0, direct: bool = False) -> np.ndarray:
    """[N,3,H,W] fp32 -> [N,48,H/4,W/4]. direct=True: 16-term dot products; else the lifting order."""
    x = np.ascontiguousarray(x, np.float32)
    n, c, H, W = x.shape
    assert c == 3 and H % 4 == 0 and W % 4 == 0
    y = np.empty((n, 48, H // 4, W // 4), np.float32)
    if direct:
        assert flags == 0
        _lib().wdm_oracle_dwt4x4_direct(_p(x), _p(y), n, H, W)
    else:
        _lib().wdm_oracle_dwt4x4(_p(x), _p(y), n, H, W, flags)
    return y


def iwt(y: np.ndarray, flags: int = 0, direct: bool = False) -> np.ndarray:
    y = np.ascontiguousarray(y, np.float32)
    n, c, h, w = y.shape
    assert c == 48
    x = np.empty((n, 3, 4 * h, 4 * w), np.float32)
    if direct:
        assert flags == 0
        _lib().wdm_oracle_iwt4x4_direct(_p(y), _p(x), n, h, w)
    else:
        _lib().wdm_oracle_iwt4x4(_p(y), _p(x), n, h, w, flags)
    return x


def rec4() -> np.ndarray:
    w = np.empty((48, 1, 4, 4), np.float32)
    _lib().wdm_oracle_rec4(_p(w))
    return w
