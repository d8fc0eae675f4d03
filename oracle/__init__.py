"""CPU oracle for the hot path (TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product path
(`diffusion_models_for_gravitational_waveform_reconstruction_b200`) never does.
"""
from .unet_oracle import *  # noqa: F401,F403
