"""Drop-in for the hot-path entry points of the reference's `inference` module."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from diffusion_models_for_gravitational_waveform_reconstruction_b200.inference import (  # noqa: E402,F401
    _build_t_schedule, _cfg_weight, _dewhiten_model, _dewhiten_train_like, _dewhiten_welch, _interp_psd_for_length,
    _load_measurement_from_h5, _load_measurement_from_npy, _mad_std, _meta_to_stack, _pick_sigma, _reduce_to_one_channel, _stats,
    _whiten_pair_model, _whiten_pair_train_like, _whiten_pair_welch, ddim_sample, load_checkpoint, make_sampler_plan,
    one_step_proxy_like_test_infer, philox_normal, snr_from_alpha_bar, t_for_target_snr)
from diffusion_models_for_gravitational_waveform_reconstruction_b200.models import CustomDiffusion, UNet1D  # noqa: E402,F401
