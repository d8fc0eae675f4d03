"""Drop-in for the reference's `dataloader` module (src/snr_denoising/dataloader.py): same names, GPU-side whitening / sigma."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from diffusion_models_for_gravitational_waveform_reconstruction_b200.dataloader import (  # noqa: E402,F401
    BatchLoader, NoisyWaveDataset, _mad_std, make_dataloader, open_h5, pad_collate, resolve_h5_path)
