"""Import shim for ``import utils`` (utils/__init__.py:1-4 star-imports logging, sampling, optimize, metrics)."""
from wavedm_b200 import logging, metrics, optimize, sampling  # noqa: F401  (utils.logging.save_image etc.)
from wavedm_b200.logging import *  # noqa: F401,F403
from wavedm_b200.logging import load_checkpoint, save_checkpoint, save_image
from wavedm_b200.metrics import calculate_psnr, calculate_psnr_in_GPU, torchPSNR
from wavedm_b200.optimize import get_optimizer, weights_init
from wavedm_b200.sampling import (compute_alpha, data_transform, generalized_steps, generalized_steps_overlapping,
                                  inverse_data_transform)
