"""Drop-in for the evaluation core of the reference's `grid_infer` module: batched per-index metrics (`eval_indices`)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from diffusion_models_for_gravitational_waveform_reconstruction_b200.sweep import Batch, eval_indices  # noqa: E402,F401
from diffusion_models_for_gravitational_waveform_reconstruction_b200.scoring import objective as _objective  # noqa: E402,F401
