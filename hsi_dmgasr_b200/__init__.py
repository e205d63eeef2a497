"""hsi_dmgasr_b200 - B200-native inference hot path of HSI-DMGASR (SR3 denoising loop in GAE latent space).

Python here is the host-side mirror of the reference's object API; all arithmetic happens in
``libhsidm_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/hsidm.h``).
"""
from .spec import GAE_PRESETS, GAEGeometry, UNetConfig  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not need torch.cuda or the shared library
    import importlib
    table = {"UNet": "unet", "GaussianDiffusion": "diffusion", "define_G": "networks", "DDPM": "model",
             "create_model": "model", "GAE": "gae", "load_gae": "gae", "save_gae_state": "gae", "SRPipeline": "pipeline",
             "set_default_precision": "unet"}
    if name in table:
        return getattr(importlib.import_module("." + table[name], __name__), name)
    raise AttributeError(name)
