"""Single-kernel parity on the B200: every CUDA kernel against the same op in plain PyTorch fp32.

Tolerances: fp32 kernels 2e-5 relative L2 (summation-order noise); bf16 tensor-core kernels are compared with a
torch fp32 conv fed the SAME bf16-rounded operands, so only accumulation order and the final bf16 rounding of the
output differ (<= 6e-3 relative L2)."""
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import conv2d, groupnorm, randn, rel_l2, tc_flag

pytestmark = pytest.mark.gpu
SIMT, TC, AUTO = 0, 1, 2


def bf(x):
    return None if x is None else x.to(torch.bfloat16).float()


def ref_conv(x0, x1, w, b, *, stride=1, up=0, nbias=None, act=0, scale=1.0, resid=None):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    if up:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(x, w, b, stride=stride, padding=w.shape[-1] // 2)
    if nbias is not None:
        y = y + nbias[:, :, None, None]
    if act:
        y = F.leaky_relu(y, 0.01)
    y = y * scale
    if resid is not None:
        y = y + resid
    return y


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


SIMT_CASES = [
    # n, c0, c1, h, w, cout, k, stride, up
    (2, 6, 0, 16, 16, 64, 3, 1, 0),
    (2, 3, 3, 16, 16, 32, 3, 1, 0),
    (1, 64, 0, 16, 16, 3, 3, 1, 0),
    (2, 32, 0, 16, 16, 32, 3, 2, 0),
    (2, 32, 0, 8, 8, 32, 3, 1, 1),
    (3, 64, 32, 8, 8, 64, 1, 1, 0),
    (1, 31, 0, 12, 20, 33, 3, 1, 0),
    (1, 64, 0, 15, 15, 64, 3, 2, 0),
]


@pytest.mark.parametrize("case", SIMT_CASES)
def test_conv_simt_fp32(case):
    n, c0, c1, h, w, cout, k, stride, up = case
    x0 = randn((n, c0, h, w), 1)
    x1 = randn((n, c1, h, w), 2) if c1 else None
    wt = randn((cout, c0 + c1, k, k), 3, scale=0.1)
    b = randn((cout,), 4)
    nb = randn((n, cout), 5)
    want = ref_conv(x0, x1, wt, b, stride=stride, up=up, nbias=nb, act=1, scale=0.5)
    resid = randn(tuple(want.shape), 6)
    want = want + resid
    got = conv2d(SIMT, "fp32", x0, x1, wt, b, ksize=k, stride=stride, up=up, nbias=nb, act=1, scale=0.5, resid=resid)
    assert rel_l2(got, want) < 2e-5


def test_conv_simt_nchw_io():
    x0, x1 = randn((2, 3, 16, 16), 1), randn((2, 3, 16, 16), 2)
    wt, b = randn((64, 6, 3, 3), 3, scale=0.2), randn((64,), 4)
    got = conv2d(SIMT, "fp32", x0, x1, wt, b, ksize=3, src_nchw=True)
    assert rel_l2(got, ref_conv(x0, x1, wt, b)) < 2e-5
    wt2, b2 = randn((3, 64, 3, 3), 5, scale=0.1), randn((3,), 6)
    x = randn((2, 64, 16, 16), 7)
    got = conv2d(SIMT, "fp32", x, None, wt2, b2, ksize=3, out_nchw=True)
    assert rel_l2(got, ref_conv(x, None, wt2, b2)) < 2e-5


def test_conv_simt_bf16_storage():
    x0 = randn((2, 64, 16, 16), 1)
    wt, b = randn((64, 64, 3, 3), 3, scale=0.05), randn((64,), 4)
    got = conv2d(SIMT, "bf16", x0, None, wt, b, ksize=3)
    assert rel_l2(got, ref_conv(bf(x0), None, wt, b)) < 6e-3


TC_CASES = [
    # n, c0, c1, h, w, cout, k          tile geometry exercised
    (1, 64, 0, 8, 16, 64, 1),         # one tile, one k-block
    (2, 64, 0, 16, 16, 64, 3),        # 16x8 boxes, 9 taps
    (3, 64, 0, 8, 8, 128, 3),         # two images per tile, odd N -> image tail
    (2, 128, 0, 32, 32, 256, 3),      # 32x4 boxes, BN=256
    (1, 64, 0, 128, 128, 64, 3),      # full rows
    (1, 64, 0, 64, 64, 128, 3),       # 64x2 boxes
    (2, 128, 64, 16, 16, 128, 3),     # two sources (virtual concat)
    (2, 256, 128, 16, 16, 128, 1),    # two sources, 1x1 (res_conv)
    (2, 512, 0, 16, 16, 1536, 1),     # qkv projection, 6 n-tiles
    (5, 512, 0, 8, 8, 512, 3),        # deep stage, long K
    (1, 64, 0, 16, 256, 64, 3),       # W > 128
]


TC_CASES += [
    (2, 64, 0, 32, 32, 64, 3),        # halo kernel <4,64>, one tile per image
    (3, 64, 64, 16, 32, 128, 3),      # halo kernel <2,128>, two sources, two tiles per row
    (1, 192, 0, 48, 64, 64, 3),       # three chunks, H = 3 tiles
    # more tiles than SMs: every CTA walks several tiles, so all barrier rings wrap (persistent-loop phases)
    (10, 64, 0, 128, 128, 64, 3),     # <4,64>: 320 tiles
    (24, 128, 64, 64, 64, 128, 3),    # <2,128>: 384 tiles, 3 chunks (odd ring occupancy per tile)
    (40, 64, 0, 32, 32, 64, 1),       # centre-tap form of a short-K 1x1: 160 tiles
    (40, 256, 0, 16, 16, 512, 1),     # per-tap kernel, 80 x 2 tiles, 4 k-blocks
    # 8x8 images on the halo kernel: two images per tile with interleaved rows (odd batches leave half a tile empty)
    (301, 64, 64, 8, 8, 128, 3),      # 151 image pairs > SMs: rings wrap; two sources; odd N
    (7, 128, 0, 8, 8, 64, 3),         # N = 64 tile
]


@pytest.fixture(params=["halo", "pertap"])
def conv_mode(request):
    from hsi_dmgasr_b200 import _lib
    lib = _lib.load()
    lib.hsidm_debug_conv_mode(1 if request.param == "pertap" else 0, 0)
    yield request.param
    lib.hsidm_debug_conv_mode(0, 0)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_matches_torch(case, conv_mode):
    n, c0, c1, h, w, cout, k = case
    x0 = randn((n, c0, h, w), 11)
    x1 = randn((n, c1, h, w), 12) if c1 else None
    wt = randn((cout, c0 + c1, k, k), 13, scale=(1.0 / ((c0 + c1) * k * k)) ** 0.5)
    b = randn((cout,), 14)
    nb = randn((n, cout), 15)
    resid = randn((n, cout, h, w), 16)
    got = conv2d(TC, "bf16", x0, x1, wt, b, ksize=k, nbias=nb, resid=resid)
    assert tc_flag() == 0, "tensor-core kernel hit a barrier timeout"
    want = ref_conv(bf(x0), bf(x1), bf(wt), b, nbias=nb, resid=bf(resid))
    assert rel_l2(got, want) < 6e-3
    # and against the CUDA-core kernel on identical bf16 inputs
    simt = conv2d(SIMT, "bf16", x0, x1, bf(wt), b, ksize=k, nbias=nb, resid=resid)
    assert rel_l2(got, simt) < 6e-3


def test_conv_tc_epilogue_variants(conv_mode):
    x = randn((2, 64, 16, 16), 21)
    wt, b = randn((64, 64, 3, 3), 22, scale=0.05), randn((64,), 23)
    got = conv2d(TC, "bf16", x, None, wt, b, ksize=3, act=1, scale=0.1, resid=x)
    want = ref_conv(bf(x), None, bf(wt), b, act=1, scale=0.1, resid=bf(x))
    assert tc_flag() == 0 and rel_l2(got, want) < 6e-3
    got = conv2d(TC, "bf16", x, None, wt, None, ksize=3)
    assert tc_flag() == 0 and rel_l2(got, ref_conv(bf(x), None, bf(wt), None)) < 6e-3


def test_conv_tc_small_cout_nchw_out(conv_mode):
    x = randn((2, 64, 32, 32), 31)
    wt, b = randn((3, 64, 3, 3), 32, scale=0.05), randn((3,), 33)
    got = conv2d(TC, "bf16", x, None, wt, b, ksize=3, out_nchw=True)
    assert tc_flag() == 0 and rel_l2(got, ref_conv(bf(x), None, bf(wt), b)) < 2e-3


@pytest.mark.parametrize("shape", [(2, 128, 16, 16, 128), (1, 64, 32, 64, 64), (3, 256, 16, 32, 256)])
def test_upsample_conv_subpixel_form(shape, conv_mode):
    """nearest-2x + 3x3 conv: halo mode runs the four sub-pixel 2x2 convs with pre-summed weights, per-tap mode the
    materialised upsample; both must match torch (the weight sums are rounded to bf16 once, hence the looser bound)."""
    n, c, h, w, co = shape
    x = randn((n, c, h, w), 61)
    wt, b = randn((co, c, 3, 3), 62, scale=(1.0 / (c * 9)) ** 0.5), randn((co,), 63)
    got = conv2d(AUTO, "bf16", x, None, wt, b, ksize=3, up=1)
    assert tc_flag() == 0 and got.shape == (n, co, 2 * h, 2 * w)
    assert rel_l2(got, ref_conv(bf(x), None, bf(wt), b, up=1)) < 8e-3


@pytest.mark.parametrize("shape", [(2, 64, 64, 64, 64), (3, 128, 32, 64, 128), (20, 256, 32, 32, 256), (1, 64, 128, 128, 64),
                                   (2, 64, 32, 32, 128)])
def test_downsample_conv_phase_lattice_form(shape):
    """Downsample (3x3, stride 2, pad 1) runs on the halo kernel over the four phase lattices of the input (no im2col
    buffer) whenever the OUTPUT is a multiple of 16 rows; (20, 256, 32, 32) has more tiles than SMs."""
    n, c, h, w, cout = shape
    x = randn((n, c, h, w), 61)
    wt, b = randn((cout, c, 3, 3), 62, scale=(1.0 / (9 * c)) ** 0.5), randn((cout,), 63)
    got = conv2d(AUTO, "bf16", x, None, wt, b, ksize=3, stride=2)
    assert tc_flag() == 0
    want = ref_conv(bf(x), None, bf(wt), b, stride=2)
    assert got.shape == want.shape and rel_l2(got, want) < 6e-3


@pytest.mark.parametrize("shape", [(2, 256, 32, 32, 256), (4, 512, 16, 16, 512), (6, 256, 32, 32, 512)])
def test_pair_tile_is_bitwise_single_cta(shape):
    """Cout % 256 == 0 runs on CTA pairs (cta_group::2, M = 256, N = 256, weights split over the pair); the developer
    switch 32 keeps the same conv on the single-CTA <2,128> tile.  Same K order per output -> the same bits."""
    from hsi_dmgasr_b200 import _lib
    lib = _lib.load()
    n, c, h, w, cout = shape
    x = randn((n, c, h, w), 71)
    wt, b = randn((cout, c, 3, 3), 72, scale=(1.0 / (9 * c)) ** 0.5), randn((cout,), 73)
    resid = randn((n, cout, h, w), 74)
    try:
        outs = []
        for variant in (0, 32):
            lib.hsidm_debug_conv_mode(0, variant)
            outs.append(conv2d(TC, "bf16", x, None, wt, b, ksize=3, resid=resid))
    finally:
        lib.hsidm_debug_conv_mode(0, 0)
    assert tc_flag() == 0
    assert torch.equal(outs[0], outs[1])
    assert rel_l2(outs[0], ref_conv(bf(x), None, bf(wt), b, resid=bf(resid))) < 6e-3


@pytest.mark.parametrize("kind", ["down", "up"])
def test_conv_dispatch_lowerings(kind):
    x = randn((2, 128, 16, 16), 41)
    wt, b = randn((128, 128, 3, 3), 42, scale=0.03), randn((128,), 43)
    if kind == "down":
        got = conv2d(AUTO, "bf16", x, None, wt, b, ksize=3, stride=2)
        want = ref_conv(bf(x), None, bf(wt), b, stride=2)
    else:
        got = conv2d(AUTO, "bf16", x, None, wt, b, ksize=3, up=1)
        want = ref_conv(bf(x), None, bf(wt), b, up=1)
    assert tc_flag() == 0 and rel_l2(got, want) < 6e-3


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 8e-3)])
@pytest.mark.parametrize("c0,c1,groups,hw", [(64, 0, 32, 16), (128, 64, 32, 8), (256, 128, 32, 16), (512, 256, 32, 8),
                                             (32, 0, 8, 16), (1024, 512, 16, 8), (64, 0, 32, 128)])
def test_groupnorm_swish(prec, tol, c0, c1, groups, hw):
    x0 = randn((2, c0, hw, hw), 51) * 1.7 + 0.3
    x1 = randn((2, c1, hw, hw), 52) if c1 else None
    gamma, beta = 1 + 0.1 * randn((c0 + c1,), 53), 0.1 * randn((c0 + c1,), 54)
    src0, src1 = (bf(x0), bf(x1)) if prec == "bf16" else (x0, x1)
    x = src0 if x1 is None else torch.cat([src0, src1], 1)
    for swish in (True, False):
        want = F.group_norm(x, groups, gamma, beta, eps=1e-5)
        if swish:
            want = want * torch.sigmoid(want)
        got = groupnorm(prec, x0, x1, groups, gamma, beta, swish)
        assert rel_l2(got, want) < tol
