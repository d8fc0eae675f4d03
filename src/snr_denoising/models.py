"""Drop-in for the reference's flat-imported `models` module (src/snr_denoising/models.py)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from diffusion_models_for_gravitational_waveform_reconstruction_b200.models import (  # noqa: E402,F401
    CustomDiffusion, TimeEmbedding, UNet1D, cosine_beta_schedule)
