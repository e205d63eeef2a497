"""Generate tests/golden/prepost.npz from the UNMODIFIED reference's GAE/imsize.py and GAE/eval_hsi.py.

TEST INFRASTRUCTURE.  Run once in the build container (``python oracle/make_golden_prepost.py``).  Inputs are re-derived
from seeds on both sides; the file holds only reference outputs:
  * imresize: the dataset code's own calls (HStest.py:44-45 - x4 degradation then pre-upsampling with output_shape; an
    anisotropic and a bilinear case besides);
  * quality_assessment's ERGAS / SAM / CrossCorrelation / RMSE (eval_hsi.py:18-96) on clamped HWC cubes.  MPSNR and
    MSSIM call skimage, which this image does not have: they are not in the file (MPSNR is pinned through the long-chain
    goldens, MSSIM is restated from skimage's published algorithm and marked unpinned).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden as MG  # noqa: E402

IMRESIZE_CASES = {   # name: (seed, (H, W, C), output_shape of the first resize, output_shape of the second or None, method)
    "hstest_x4": (11, (64, 64, 6), (16, 16), (64, 64), "bicubic"),
    "odd_x4": (12, (36, 52, 3), (9, 13), (36, 52), "bicubic"),
    "aniso": (13, (40, 24, 2), (10, 36), None, "bicubic"),
    "bilinear": (14, (32, 32, 2), (8, 8), (32, 32), "bilinear"),
    "down3": (15, (45, 33, 1), (15, 11), (45, 33), "bicubic"),
}
ASSESS_CASES = {"a": (21, (2, 7, 24, 20)), "b": (22, (1, 31, 32, 32))}   # name: (seed, NCHW)


def imresize_input(seed, shape):
    return np.random.default_rng(seed).random(shape, dtype=np.float32)


def assess_inputs(seed, shape):
    rng = np.random.default_rng(seed)
    truth = rng.random(shape, dtype=np.float32)
    pred = truth + 0.05 * rng.standard_normal(shape, dtype=np.float32)     # leaves [0,1]: the clamp matters
    truth[0, :, 0, 0] = 0.0                                                # a zero spectrum: excluded from SAM
    return truth, pred


def main():
    _, eval_hsi, _, _ = MG.import_reference()
    sys.path.insert(0, os.path.join(MG.REF, "GAE"))
    import imsize  # noqa: E402  (the reference's GAE/imsize.py; HStest.py imports it the same way)
    out = {}
    for name, (seed, shape, first, second, method) in IMRESIZE_CASES.items():
        x = imresize_input(seed, shape)
        ms = imsize.imresize(x, output_shape=first, method=method)
        out[f"imresize.{name}.first"] = ms
        if second is not None:
            # HStest.py:45 feeds the float64 result of the first call straight into the second
            out[f"imresize.{name}.second"] = imsize.imresize(ms, output_shape=second, method=method)
            # what a device pipeline sees: the fp32-rounded intermediate
            out[f"imresize.{name}.second_from_f32"] = imsize.imresize(ms.astype(np.float32), output_shape=second, method=method)
    x = imresize_input(16, (20, 28, 2))
    out["imresize.scalar_scale.first"] = imsize.imresize(x, scalar_scale=0.3)
    for name, (seed, shape) in ASSESS_CASES.items():
        truth, pred = assess_inputs(seed, shape)
        rows = []
        for t, p in zip(truth, pred):
            th, ph = np.clip(t, 0, 1).transpose(1, 2, 0), np.clip(p, 0, 1).transpose(1, 2, 0)
            rows.append([eval_hsi.compare_ergas(th, ph, 4), eval_hsi.compare_sam(th, ph), eval_hsi.compare_corr(th, ph),
                         eval_hsi.compare_rmse(th, ph)])
        out[f"assess.{name}"] = np.asarray(rows, dtype=np.float64)      # [N][ERGAS, SAM, CC, RMSE]
    np.savez_compressed(os.path.join(MG.OUT, "prepost.npz"), **out)
    print("wrote prepost.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
