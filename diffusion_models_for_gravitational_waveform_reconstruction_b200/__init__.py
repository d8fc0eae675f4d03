"""B200-native (sm_100a) implementation of the snr_denoising hot path.

Public surface mirrors the reference's `models`, `inference` and `train` modules; see DESIGN.md.
"""
from . import _cabi  # noqa: F401
from .engine import ModelSpec, SamplerPlan, UNetEngine  # noqa: F401
from .models import CustomDiffusion, TimeEmbedding, UNet1D, cosine_beta_schedule  # noqa: F401

__version__ = "0.1.0"
