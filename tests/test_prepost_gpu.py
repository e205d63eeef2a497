"""SURVEY 8f row N3 on the B200: bicubic x4 pre-upsampling against the reference's own call (torch bicubic on the CPU)
and the device-side MPSNR / SAM against the restated eval_hsi.py metrics."""
import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import prepost
from hsi_dmgasr_b200._lib import HsidmError
from oracle import hsidm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,scale", [((1, 31, 32, 32), 4), ((2, 128, 32, 32), 4), ((1, 102, 17, 23), 4), ((3, 5, 8, 8), 2),
                                         ((1, 3, 1, 7), 4)])
def test_bicubic_matches_torch(shape, scale):
    lr = torch.from_numpy(np.random.default_rng(7).random(shape, dtype=np.float32))
    want = O.bicubic_pre_upsample(lr, scale)
    got = prepost.bicubic_upsample(lr.cuda(), scale).cpu()
    assert got.shape == want.shape
    # same formula, different summation order: a few ulp of values in [-0.3, 1.3]
    assert float((got - want).abs().max()) < 2e-6
    clamped = prepost.bicubic_upsample(lr.cuda(), scale, clamp01=True).cpu()
    assert torch.equal(clamped, got.clamp(0, 1))


def test_bicubic_3d_input_and_errors():
    lr = torch.rand(4, 8, 8)
    got = prepost.bicubic_upsample(lr.cuda(), 4)
    assert got.shape == (4, 32, 32)
    with pytest.raises(HsidmError):
        prepost.bicubic_upsample(lr, 4)                  # CPU tensor: no fallback
    with pytest.raises(HsidmError):
        prepost.bicubic_upsample(torch.rand(8, 8).cuda(), 4)


@pytest.mark.parametrize("shape", [(2, 31, 128, 128), (1, 128, 64, 48), (3, 4, 16, 16)])
def test_metrics_match_eval_hsi(shape):
    rng = np.random.default_rng(3)
    truth = torch.from_numpy(rng.random(shape, dtype=np.float32))
    pred = (truth + 0.05 * torch.from_numpy(rng.standard_normal(shape, dtype=np.float32)))     # leaves [0,1]: clamp matters
    truth[0, :, 0, 0] = 0.0                                                                       # a zero spectrum: excluded from SAM
    want = O.cube_metrics(truth, pred)
    got = prepost.quality_metrics(truth.cuda(), pred.cuda()).cpu()
    for n, (m, s) in enumerate(want):
        assert abs(float(got[n, 0]) - m) < 1e-3, (n, float(got[n, 0]), m)      # dB
        assert abs(float(got[n, 1]) - s) < 1e-3, (n, float(got[n, 1]), s)      # degrees
    again = prepost.quality_metrics(truth.cuda(), pred.cuda()).cpu()
    assert torch.equal(got, again)                                             # fixed-order folds: same bits


def test_lr_to_metrics_pipeline_shapes():
    """LR cube -> bicubic x4 (clamped) -> metrics against the HR cube, all on the device."""
    hr = torch.rand(2, 31, 64, 64).cuda()
    lr = torch.nn.functional.avg_pool2d(hr, 4)
    sr = prepost.bicubic_upsample(lr, 4, clamp01=True)
    m, s = prepost.mean_metrics(hr, sr)
    assert sr.shape == hr.shape and 5.0 < m < 60.0 and 0.0 < s < 90.0


# ---- MATLAB-style imresize (GAE/imsize.py) and the full quality_assessment (eval_hsi.py:217-238) ------------------------
from oracle import make_golden_prepost as GP  # noqa: E402  (case tables and seeded inputs; the reference is not needed here)


def _planes(hwc: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(hwc.transpose(2, 0, 1)))


@pytest.mark.parametrize("name", list(GP.IMRESIZE_CASES))
def test_imresize_matches_reference_vectors(golden, name):
    """hsidm_imresize against the outputs of the unmodified GAE/imsize.py (tests/golden/prepost.npz) for the dataset
    code's own call pattern: x4 degradation with output_shape, then pre-upsampling of the result (HStest.py:44-45).
    The reference works in float64; the device result is its fp32 rounding (values in [-0.2, 1.2]: 6e-8 per ulp)."""
    g = golden("prepost.npz")
    seed, shape, first, second, method = GP.IMRESIZE_CASES[name]
    x = _planes(GP.imresize_input(seed, shape)).cuda()
    ms = prepost.imresize(x, output_shape=first, method=method)
    ref = _planes(g[f"imresize.{name}.first"])
    assert ms.shape == ref.shape and ms.dtype == torch.float32
    assert float((ms.cpu().double() - ref).abs().max()) < 2e-7
    if second is not None:
        lms = prepost.imresize(ms, output_shape=second, method=method)
        # the device chain feeds the fp32 intermediate: compare with the reference run on that same fp32 intermediate
        # (the oracle reproduces it from the device's ms), and with the reference's all-float64 chain more loosely
        want = O.imresize_matlab(ms.cpu().numpy().transpose(1, 2, 0), second, method)
        assert float((lms.cpu().double() - _planes(want)).abs().max()) < 2e-7
        assert float((lms.cpu().double() - _planes(g[f"imresize.{name}.second"])).abs().max()) < 1e-6


def test_imresize_scalar_scale_batched_and_errors(golden):
    g = golden("prepost.npz")
    x = _planes(GP.imresize_input(16, (20, 28, 2))).cuda()
    got = prepost.imresize(x, scalar_scale=0.3)
    ref = _planes(g["imresize.scalar_scale.first"])
    assert got.shape == ref.shape == (2, 6, 9)
    assert float((got.cpu().double() - ref).abs().max()) < 2e-7
    # [N,C,h,w] and [h,w] inputs resize plane by plane
    b = prepost.imresize(torch.stack([x, x]), scalar_scale=0.3)
    assert b.shape == (2, 2, 6, 9) and torch.equal(b[1], got)
    assert torch.equal(prepost.imresize(x[0], scalar_scale=0.3), got[0])
    ms, lms = prepost.degrade_and_preupsample(torch.rand(3, 5, 32, 32).cuda(), 4)
    assert ms.shape == (3, 5, 8, 8) and lms.shape == (3, 5, 32, 32)
    with pytest.raises(ValueError):
        prepost.imresize(x)
    with pytest.raises(ValueError):
        prepost.imresize(x, scalar_scale=2, output_shape=(4, 4))
    with pytest.raises(ValueError):
        prepost.imresize(x, scalar_scale=2, method="lanczos")
    with pytest.raises(HsidmError):
        prepost.imresize(x.cpu(), scalar_scale=2)


@pytest.mark.parametrize("name", list(GP.ASSESS_CASES))
def test_quality_assessment_matches_reference_vectors(golden, name):
    """ERGAS / SAM / CrossCorrelation / RMSE against eval_hsi.py's own outputs; MPSNR and MSSIM (skimage calls in the
    reference) against the oracle's restatements."""
    g = golden("prepost.npz")
    truth, pred = GP.assess_inputs(*GP.ASSESS_CASES[name])
    t, p = torch.from_numpy(truth), torch.from_numpy(pred)
    got = prepost.quality_assessment(t.cuda(), p.cuda(), ratio=4.0).cpu().double().numpy()
    want = O.cube_assessment(t, p, 4.0)
    for n, ref in enumerate(g[f"assess.{name}"]):
        assert abs(got[n, 2] - ref[0]) < 2e-4 * ref[0], ("ERGAS", got[n, 2], ref[0])
        assert abs(got[n, 3] - ref[1]) < 1e-3, ("SAM", got[n, 3], ref[1])
        assert abs(got[n, 4] - ref[2]) < 1e-5, ("CC", got[n, 4], ref[2])
        assert abs(got[n, 5] - ref[3]) < 1e-6, ("RMSE", got[n, 5], ref[3])
        assert abs(got[n, 0] - want[n][0]) < 1e-3, ("MPSNR", got[n, 0], want[n][0])
        assert abs(got[n, 1] - want[n][1]) < 1e-5, ("MSSIM", got[n, 1], want[n][1])
    two = prepost.quality_metrics(t.cuda(), p.cuda()).cpu().double().numpy()
    assert np.array_equal(two[:, 0], got[:, 0]) and np.array_equal(two[:, 1], got[:, 3])   # same kernels
    again = prepost.quality_assessment(t.cuda(), p.cuda(), ratio=4.0).cpu().double().numpy()
    assert np.array_equal(got, again)


def test_quality_assessment_bench_shape_and_dict():
    """16 Chikusei-shaped cubes (the C2 batch): identical cubes give the ideal values; the dict form has the reference's keys."""
    hr = torch.rand(16, 128, 128, 128).cuda()
    same = prepost.quality_assessment(hr, hr).cpu()
    assert torch.all(same[:, 1] > 1 - 1e-6) and torch.all(same[:, 2] == 0) and torch.all(same[:, 5] == 0)
    assert torch.all((same[:, 4] - 1).abs() < 1e-6) and torch.isinf(same[:, 0]).all()
    d = prepost.quality_assessment_dict(hr[0], (hr[0] + 0.01).clamp(0, 1))
    assert list(d) == ["MPSNR", "MSSIM", "ERGAS", "SAM", "CrossCorrelation", "RMSE"]
    assert 39.0 < d["MPSNR"] < 41.0 and abs(d["RMSE"] - 0.01) < 1e-3
    tiny = prepost.quality_assessment(hr[:1, :4, :5, :5], hr[:1, :4, :5, :5]).cpu()
    assert torch.isnan(tiny[0, 1])                      # no 7x7 window fits: MSSIM undefined
