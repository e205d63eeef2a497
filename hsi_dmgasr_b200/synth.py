"""Seeded synthetic inputs: weights, hyperspectral cubes and noise tapes.

No UNet checkpoint and no HSI data ship with the reference, so tests and ``bench.py``
use random-init weights of the named architecture and synthetic cubes of the named shapes.  Everything is
drawn from numpy's PCG64 streams so the same seed gives the same bytes on every box.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .spec import GAEGeometry, UNetConfig, gae_param_shapes, unet_param_shapes


def _fill(shapes: "OrderedDict[str, Tuple[int, ...]]", seed: int, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    rng = np.random.default_rng(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in shapes.items():
        if len(shape) == 1:
            is_norm_scale = key.endswith("weight")           # 1-D weights only occur in GroupNorm
            v = rng.uniform(-0.1, 0.1, size=shape)
            if is_norm_scale:
                v = 1.0 + v
        else:
            fan_in = int(np.prod(shape[1:]))
            b = gain * np.sqrt(3.0 / fan_in)                  # unit-variance-preserving uniform init
            v = rng.uniform(-b, b, size=shape)
        sd[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def unet_state_dict(cfg: UNetConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random weights keyed like ``UNet.state_dict()`` (keys relative to ``denoise_fn.``)."""
    return _fill(unet_param_shapes(cfg), seed, gain=1.0)


def gae_state_dict(geom: GAEGeometry, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random weights keyed like ``GAE.state_dict()``."""
    return _fill(gae_param_shapes(geom), seed, gain=1.0)


def sr_cube(batch: int, bands: int, size: int, seed: int = 0, scale: int = 4) -> torch.Tensor:
    """Bicubic-upsampled synthetic LR cube in [0,1] (sr_gae.py:72, HStest.py:59-60)."""
    rng = np.random.default_rng(seed)
    lr = torch.from_numpy(rng.random((batch, bands, size // scale, size // scale), dtype=np.float32))
    sr = F.interpolate(lr, scale_factor=scale, mode="bicubic", align_corners=False)
    return sr.clamp_(0.0, 1.0).contiguous()


def noise_tape(images: int, steps: int, channels: int, h: int, w: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x_T [images,C,H,W], tape [images, steps-1, C,H,W]); tape[:, j] is used at loop index i = steps-1-j.

    Image-major ("group-major") like the reference driver, which finishes all T draws of one group before
    the next.
    """
    rng = np.random.default_rng(seed)
    x_T = torch.from_numpy(rng.standard_normal((images, channels, h, w), dtype=np.float32))
    tape = torch.from_numpy(rng.standard_normal((images, max(steps - 1, 0), channels, h, w), dtype=np.float32))
    return x_T, tape
