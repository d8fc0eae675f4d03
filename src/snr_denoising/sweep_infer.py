"""Drop-in for the evaluation core of the reference's `sweep_infer` module: batched `eval_combo`, grid search and random sweep."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from diffusion_models_for_gravitational_waveform_reconstruction_b200.sweep import (  # noqa: E402,F401
    Batch, best_command, eval_combo, grid_search, random_sweep, sample_combo)
from diffusion_models_for_gravitational_waveform_reconstruction_b200.scoring import objective as _objective  # noqa: E402,F401
