"""Makes the reference's own import names resolve to this package, so an unmodified driver such as ``sr_gae.py``
(``import model as Model``, ``from AE import *``, ``torch.load('GAE_4_Pav.pth')`` of ``__main__.GAE`` pickles) runs
on the native path.  Call ``install()`` before those imports."""
from __future__ import annotations

import sys
import types


def install(main_module=None) -> None:
    from . import diffusion, gae, model, networks, unet

    def alias(name, mod):
        sys.modules[name] = mod

    pkg_model = types.ModuleType("model")
    pkg_model.create_model = model.create_model
    pkg_model.__path__ = []            # mark as package so `import model.networks` works
    alias("model", pkg_model)
    alias("model.model", model)
    alias("model.networks", networks)
    pkg_model.networks, pkg_model.model = networks, model
    sr3 = types.ModuleType("model.sr3_modules")
    sr3.__path__ = []
    sr3.unet, sr3.diffusion = unet, diffusion
    alias("model.sr3_modules", sr3)
    alias("model.sr3_modules.unet", unet)
    alias("model.sr3_modules.diffusion", diffusion)
    pkg_model.sr3_modules = sr3
    alias("AE", gae)
    common = types.ModuleType("common")
    for name in ("ResBlock", "ResAttentionBlock", "CALayer", "Upsampler", "default_conv"):
        setattr(common, name, getattr(gae, name))
    alias("common", common)
    # whole-module GAE pickles name their classes as __main__.* (AE.py:637 was run as a script)
    main = main_module or sys.modules.get("__main__")
    if main is not None:
        for name in ("GAE", "Encoder", "Decoder", "BranchUnit", "SSPN", "SSB"):
            if not hasattr(main, name):
                setattr(main, name, getattr(gae, name))
