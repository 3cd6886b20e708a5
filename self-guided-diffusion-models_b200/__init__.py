"""sgdm_b200 — B200-native guided reverse-diffusion hot path.

Drop-in for the reference's `dynamic` UNet modules (unet_fast / unetca_fast) and the
sampling entry points of diffusion/ddpm.py + diffusion/sampler, backed by a C-ABI
CUDA library (csrc/ -> libsgdm_b200.so, sm_100a only).  There is no CPU fallback:
any compute call raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
