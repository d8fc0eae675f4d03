from setuptools import setup, find_packages

setup(
    name="snr_denoising",
    version="0.1.0",
    description="B200-native drop-in for the snr_denoising diffusion hot path",
    packages=find_packages("src") + ["diffusion_models_for_gravitational_waveform_reconstruction_b200"],
    package_dir={"": "src", "diffusion_models_for_gravitational_waveform_reconstruction_b200":
                 "diffusion_models_for_gravitational_waveform_reconstruction_b200"},
)
