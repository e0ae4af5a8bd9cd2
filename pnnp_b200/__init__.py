"""pnnp_b200 — B200-native (sm_100a) implementation of the PNNP data-parallel hot path:
physics-based noise synthesis over packed Bayer tensors → UNet denoiser forward.

`from pnnp_b200 import *` provides the names the reference's trainers resolve through
globals() (trainer_SID.py:17,48,69): raw2bayer, bayer2raw, sample_params, sample_params_max,
generate_noisy_obs, generate_noisy_torch, (UNetSeeInDark, ResUnet, initialize_weights, ...).
"""
from . import _lib  # noqa: F401
from .isp_ops import raw2bayer, bayer2raw, bayer2rggb, rggb2bayer
from .noise_params import (Dual_ISO_Cameras, HALF_CLIP, ParamTable, get_camera_noisy_params,
                           get_specific_noise_params, sample_params, sample_params_max)
from .noise import (generate_noisy_obs, generate_noisy_torch, noise_code_bits, replay_batch,
                    synthesize_batch)
from .rng import PhiloxGenerator, default_generator, manual_seed
from .archs import ResUnet, UNetSeeInDark, initialize_weights

__all__ = [
    "raw2bayer", "bayer2raw", "bayer2rggb", "rggb2bayer",
    "Dual_ISO_Cameras", "HALF_CLIP", "ParamTable", "get_camera_noisy_params", "get_specific_noise_params",
    "sample_params", "sample_params_max",
    "generate_noisy_obs", "generate_noisy_torch", "noise_code_bits", "replay_batch", "synthesize_batch",
    "PhiloxGenerator", "default_generator", "manual_seed",
    "UNetSeeInDark", "ResUnet", "initialize_weights",
]
