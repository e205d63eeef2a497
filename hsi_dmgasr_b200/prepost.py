"""The steps either side of the sampling path, on the GPU.

``bicubic_upsample`` replaces the dataset code's ``torch.nn.functional.interpolate(img_LR, scale_factor=4,
mode='bicubic')`` (reference sr_gae.py:72, :118); ``quality_metrics`` replaces the CPU ``compare_mpsnr`` /
``compare_sam`` of the validation loop (eval_hsi.py:110-121, :47-65; the latter is a Python double loop over pixels in
the reference).  Both call the C ABI (hsidm_bicubic_upsample / hsidm_quality_metrics); there is no CPU fallback here -
``metrics.py`` keeps the numpy forms the parity tests are stated in.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


def bicubic_upsample(lr: torch.Tensor, scale: int = 4, clamp01: bool = False) -> torch.Tensor:
    """lr [N,C,h,w] (or [C,h,w]) fp32 CUDA -> [N,C,h*scale,w*scale]; ``clamp01`` applies the dataset's clamp to [0,1]."""
    squeeze = lr.dim() == 3
    x = _lib.require_cuda_f32(lr.unsqueeze(0) if squeeze else lr, "lr")
    if x.dim() != 4:
        raise _lib.HsidmError(-1, f"lr must be [N,C,h,w] or [C,h,w], got {tuple(lr.shape)}")
    n, c, h, w = x.shape
    out = torch.empty((n, c, h * scale, w * scale), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().hsidm_bicubic_upsample(x.data_ptr(), out.data_ptr(), n, c, h, w, int(scale), int(clamp01),
                                                  _lib.stream_ptr(x.device)))
    return out[0] if squeeze else out


def quality_metrics(truth: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """truth / pred [N,C,H,W] fp32 CUDA cubes -> [N,2] = (MPSNR dB, SAM degrees) per cube, both clamped to [0,1] first."""
    a = _lib.require_cuda_f32(truth, "truth")
    b = _lib.require_cuda_f32(pred, "pred")
    if a.dim() != 4 or tuple(a.shape) != tuple(b.shape):
        raise _lib.HsidmError(-1, f"truth {tuple(truth.shape)} and pred {tuple(pred.shape)} must be equal [N,C,H,W] shapes")
    n, c, h, w = a.shape
    out = torch.empty((n, 2), device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().hsidm_quality_metrics(a.data_ptr(), b.data_ptr(), n, c, h, w, out.data_ptr(), _lib.stream_ptr(a.device)))
    return out


def mean_metrics(truth: torch.Tensor, pred: torch.Tensor) -> Tuple[float, float]:
    """Validation-loop style averages over the cubes of a batch: (mean MPSNR, mean SAM)."""
    m = quality_metrics(truth, pred).mean(dim=0)
    return float(m[0]), float(m[1])
