"""wavedm_b200 -- B200-native (sm_100a) implementation of WaveDM's wavelet-diffusion sampling hot path.

Host side: Python mirroring the reference's class API (WaveletTransform, DiffusionUNet,
DenoisingDiffusion_Wavelet, DiffusiveRestoration); device side: hand-written CUDA behind the C ABI of
include/wavedm_b200.h (libwavedm_b200.so, bound with ctypes in wavedm_b200/_lib.py).
"""
__version__ = "0.1"
