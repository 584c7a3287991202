"""Import shim: makes ``from models import DenoisingDiffusion, DenoisingDiffusion_Wavelet, DiffusiveRestoration``
(train_diffusion.py:14, eval_diffusion.py:13) resolve to wavedm_b200. Put this directory on PYTHONPATH."""
from wavedm_b200.ddm_wavelet import *  # noqa: F401,F403
from wavedm_b200.ddm_wavelet import DenoisingDiffusion_Wavelet, EMAHelper, get_beta_schedule, noise_estimation_loss
from wavedm_b200.restoration import DiffusiveRestoration
from wavedm_b200.unet import DiffusionUNet
from wavedm_b200.wavelet import WaveletTransform
from wavedm_b200.hfrm import HFRM


class DenoisingDiffusion(object):
    """models/ddm.py:124 -- the non-wavelet pixel-space model. Bit-rotted in the reference itself (configs/raindrop.yml
    lacks keys models/unet.py dereferences); kept importable, not implemented."""

    def __init__(self, args, config):
        raise NotImplementedError("DenoisingDiffusion (non-wavelet) is out of scope; use a config with data.wavelet: True")
