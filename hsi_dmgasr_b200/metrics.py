"""MPSNR and SAM exactly as the reference's validation computes them (eval_hsi.py:110-121 and :47-65), vectorised.

Inputs are HWC arrays after the driver's clamp to [0,1] (sr_gae.py:474-475, 483-484). These two are the metrics the
parity gates are stated in; all six indices of quality_assessment run on the device through
``prepost.quality_assessment`` (hsidm_quality_assessment)."""
from __future__ import annotations

import numpy as np


def mpsnr(x_true: np.ndarray, x_pred: np.ndarray, data_range: float = 1.0) -> float:
    """Mean over bands of 10*log10(R^2 / MSE_band) (skimage.metrics.peak_signal_noise_ratio per band)."""
    a = np.asarray(x_true, dtype=np.float32).astype(np.float64)
    b = np.asarray(x_pred, dtype=np.float32).astype(np.float64)
    err = np.mean((a - b) ** 2, axis=(0, 1))
    return float(np.mean(10.0 * np.log10((data_range ** 2) / err)))


def sam_degrees(x_true: np.ndarray, x_pred: np.ndarray) -> float:
    """Mean spectral angle in degrees over the pixels whose two spectra are both non-zero."""
    t = np.asarray(x_true, dtype=np.float32).reshape(-1, x_true.shape[-1])
    p = np.asarray(x_pred, dtype=np.float32).reshape(-1, x_pred.shape[-1])
    nt, npd = np.linalg.norm(t, axis=1), np.linalg.norm(p, axis=1)
    keep = (nt != 0) & (npd != 0)
    cosine = np.einsum("ij,ij->i", p[keep], t[keep]) / (nt[keep] * npd[keep])
    return float(np.arccos(np.clip(cosine, -1.0, 1.0)).sum() / keep.sum() * 180.0 / np.pi)
