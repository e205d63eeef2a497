"""The steps either side of the sampling path, on the GPU.

``bicubic_upsample`` replaces the dataset code's ``torch.nn.functional.interpolate(img_LR, scale_factor=4,
mode='bicubic')`` (reference sr_gae.py:72, :118); ``quality_metrics`` replaces the CPU ``compare_mpsnr`` /
``compare_sam`` of the validation loop (eval_hsi.py:110-121, :47-65; the latter is a Python double loop over pixels in
the reference).  ``imresize`` is the MATLAB-style resize of the dataset classes (GAE/imsize.py:116-158, called by
HStest.py:44-45 / HStrain.py:61-63 - a different bicubic from torch's: a = -0.5, antialiased when shrinking, mirrored
borders); ``quality_assessment`` returns all six indices of eval_hsi.py:217-238.  Everything calls the C ABI
(hsidm_bicubic_upsample / hsidm_imresize / hsidm_quality_metrics / hsidm_quality_assessment); there is no CPU fallback
here - ``metrics.py`` keeps the numpy forms the parity tests are stated in.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

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


_METHODS = {"bicubic": 0, "bilinear": 1}


def imresize(img: torch.Tensor, scalar_scale: Optional[float] = None, method: str = "bicubic",
             output_shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    """imsize.imresize on the device: img [N,C,h,w], [C,h,w] or [h,w] fp32 CUDA (band planes; the reference's HWC arrays
    are resized band by band) -> the same rank with the two spatial axes resized.  Same argument rules as the reference:
    exactly one of ``scalar_scale`` / ``output_shape`` (ValueError otherwise), output size ``ceil(scale * size)``."""
    if method not in _METHODS:
        raise ValueError("unidentified kernel method supplied")
    if (scalar_scale is None) == (output_shape is None):
        raise ValueError("either scalar_scale OR output_shape should be defined")
    x = _lib.require_cuda_f32(img, "img")
    if x.dim() not in (2, 3, 4):
        raise _lib.HsidmError(-1, f"img must be [N,C,h,w], [C,h,w] or [h,w], got {tuple(img.shape)}")
    h, w = x.shape[-2:]
    if scalar_scale is not None:
        import math
        sc = float(scalar_scale)
        oh, ow = int(math.ceil(sc * h)), int(math.ceil(sc * w))
    else:
        sc = 0.0                                   # the library derives out / in per axis
        oh, ow = int(output_shape[0]), int(output_shape[1])
    planes = int(x.numel() // (h * w))
    out = torch.empty(tuple(x.shape[:-2]) + (oh, ow), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().hsidm_imresize(x.data_ptr(), out.data_ptr(), planes, h, w, oh, ow, sc, sc, _METHODS[method],
                                          _lib.stream_ptr(x.device)))
    return out


def degrade_and_preupsample(gt: torch.Tensor, n_scale: int = 4) -> Tuple[torch.Tensor, torch.Tensor]:
    """HStest.py:44-45 on the device: ms = imresize(gt, gt_size // n_scale), lms = imresize(ms, gt_size)."""
    h, w = gt.shape[-2:]
    ms = imresize(gt, output_shape=(h // n_scale, w // n_scale))
    return ms, imresize(ms, output_shape=(h, w))


ASSESSMENT_KEYS = ("MPSNR", "MSSIM", "ERGAS", "SAM", "CrossCorrelation", "RMSE")


def quality_assessment(truth: torch.Tensor, pred: torch.Tensor, ratio: float = 4.0) -> torch.Tensor:
    """truth / pred [N,C,H,W] fp32 CUDA cubes -> [N,6] in the key order of eval_hsi.quality_assessment
    (``ASSESSMENT_KEYS``), data_range 1, both cubes clamped to [0,1] first like the validation driver does."""
    a = _lib.require_cuda_f32(truth, "truth")
    b = _lib.require_cuda_f32(pred, "pred")
    if a.dim() != 4 or tuple(a.shape) != tuple(b.shape):
        raise _lib.HsidmError(-1, f"truth {tuple(truth.shape)} and pred {tuple(pred.shape)} must be equal [N,C,H,W] shapes")
    n, c, h, w = a.shape
    out = torch.empty((n, 6), device=a.device, dtype=torch.float32)
    if c > 65535:
        raise _lib.HsidmError(-1, f"quality_assessment: {c} bands exceed the 65535 planes of one launch")
    step = max(1, 65535 // c)                       # one launch takes at most 65535 (cube, band) planes
    lib, st = _lib.load(), _lib.stream_ptr(a.device)
    for n0 in range(0, n, step):
        cnt = min(step, n - n0)
        _lib.check(lib.hsidm_quality_assessment(a[n0:n0 + cnt].data_ptr(), b[n0:n0 + cnt].data_ptr(), cnt, c, h, w, float(ratio),
                                                out[n0:n0 + cnt].data_ptr(), st))
    return out


def quality_assessment_dict(truth: torch.Tensor, pred: torch.Tensor, ratio: float = 4.0) -> Dict[str, float]:
    """The reference's return shape for ONE cube pair ([C,H,W] or [1,C,H,W]): {'MPSNR': ..., 'MSSIM': ..., ...}."""
    t = truth if truth.dim() == 4 else truth.unsqueeze(0)
    p = pred if pred.dim() == 4 else pred.unsqueeze(0)
    row = quality_assessment(t, p, ratio)[0].tolist()
    return dict(zip(ASSESSMENT_KEYS, row))


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
