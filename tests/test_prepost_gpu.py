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
