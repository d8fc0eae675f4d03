"""Drop-in for the hot-path entry points of the reference's `train` module (src/snr_denoising/train.py)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from diffusion_models_for_gravitational_waveform_reconstruction_b200.train import (  # noqa: E402,F401
    FusedTrainStep, _element_loss, _match_batch, _predict_x0_norm, _sample_timesteps_stratified,
    _compute_meta_scale, make_warmup_cosine_scheduler, train_diffusion, update_ema, warmup_cosine_lambda)
from diffusion_models_for_gravitational_waveform_reconstruction_b200.models import CustomDiffusion, UNet1D  # noqa: E402,F401
