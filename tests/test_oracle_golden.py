"""Pins oracle/hsidm_oracle.py against vectors produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import synth
from oracle import hsidm_oracle as O
from tests.cfgs import GAE_CASES, SMALL, UNET_CASES


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32))


@pytest.mark.parametrize("name,T", [("cosine", 20), ("cosine", 50), ("cosine", 2000), ("linear", 30), ("quad", 30),
                                    ("warmup10", 30), ("warmup50", 30), ("const", 10), ("jsd", 10)])
def test_schedule_tables_bit_exact(golden, name, T):
    g = golden("schedules.npz")
    with np.errstate(divide="ignore"):
        tab = O.schedule_tables(O.beta_schedule(name, T, 1e-6, 1e-2))
    for k, v in tab.items():
        ref = g[f"{name}{T}.{k}"]
        assert v.dtype == ref.dtype and v.shape == ref.shape
        assert np.array_equal(v, ref, equal_nan=True), (name, T, k)


def test_schedule_spot_values():
    # SURVEY.md 8c spot values measured on the reference at T=50
    tab = O.schedule_tables(O.beta_schedule("cosine", 50, 1e-6, 1e-2))
    assert abs(tab["betas"][0] - 0.0017475) < 1e-7 and abs(tab["betas"][49] - 0.999) < 1e-7
    assert abs(tab["sqrt_alphas_cumprod_prev"][1] - 0.99912586) < 1e-7
    assert abs(tab["posterior_log_variance_clipped"][0] + 46.0517) < 1e-3


@pytest.mark.parametrize("tag", list(UNET_CASES))
def test_unet_forward_matches_reference(golden, tag):
    cfg, seed, n, hw, lvls = UNET_CASES[tag]
    g = golden("unet_forward.npz")
    sd = synth.unet_state_dict(cfg, seed)
    x = rand((n, 6, hw, hw), 1000 + seed)
    lv = torch.tensor(lvls, dtype=torch.float32).view(n, 1)
    taps = {}
    with torch.no_grad():
        y = O.unet_forward(sd, cfg.as_dict(), x, lv, taps)
    assert rel(y.numpy(), g[f"{tag}.eps"]) < 2e-6
    for k, v in taps.items():
        if f"{tag}.tap.{k}" in g:
            assert rel(v.numpy(), g[f"{tag}.tap.{k}"]) < 2e-6, k
        else:
            fp = g[f"{tag}.fp.{k}"]
            mine = np.array([v.mean().item(), v.std().item(), v.abs().max().item()])
            assert np.allclose(mine, fp, rtol=1e-4, atol=1e-6), k


def test_sample_loop_matches_reference(golden):
    g = golden("sample_loop.npz")
    T, n, hw = int(g["T"]), 2, 16
    sd = synth.unet_state_dict(SMALL, 21)
    tab = O.schedule_tables(O.beta_schedule("cosine", T, 1e-6, 1e-2))
    cond = rand((n, 3, hw, hw), 31)
    x_T, tape = synth.noise_tape(n, T, 3, hw, hw, seed=32)
    eps_l, x_l = [], []
    with torch.no_grad():
        ret = O.sample_loop(sd, SMALL.as_dict(), tab, cond, x_T, lambda i: tape[:, T - 1 - i], continous=True,
                            record=lambda i, e, x: (eps_l.append(e), x_l.append(x)))
        last = O.sample_loop(sd, SMALL.as_dict(), tab, cond, x_T, lambda i: tape[:, T - 1 - i], continous=False)
    assert rel(torch.stack(eps_l).numpy(), g["eps"]) < 5e-6
    assert rel(torch.stack(x_l).numpy(), g["x"]) < 5e-6
    assert ret.shape == g["ret_all"].shape and rel(ret.numpy(), g["ret_all"]) < 5e-6
    assert last.shape == g["ret_last"].shape == (3, hw, hw) and rel(last.numpy(), g["ret_last"]) < 5e-6


@pytest.mark.parametrize("tag", list(GAE_CASES))
def test_gae_matches_reference(golden, tag):
    geom, seed, hw = GAE_CASES[tag]
    g = golden("gae.npz")
    G, start, end = O.gae_groups(geom.n_colors, geom.n_subs, geom.n_ovls)
    assert start == list(g[f"{tag}.start"]) and end == list(g[f"{tag}.end"]) and G == geom.G
    assert (start, end) == geom.groups()
    sd = synth.gae_state_dict(geom, seed)
    x = synth.sr_cube(2, geom.n_colors, hw, seed=seed + 100)
    with torch.no_grad():
        zs = O.gae_encode(sd, geom.as_dict(), x)
        y = O.gae_decode(sd, geom.as_dict(), x, zs)
    assert rel(torch.stack(zs).numpy(), g[f"{tag}.z"]) < 2e-6
    assert rel(y.numpy(), g[f"{tag}.dec"]) < 2e-6


@pytest.mark.parametrize("tag", ["e2e", "e2e_full"])
def test_end_to_end_cube_and_metrics(golden, tag):
    from hsi_dmgasr_b200.spec import GAEGeometry
    from tests.cfgs import FULL
    g = golden(tag + ".npz")
    cfg = SMALL if tag == "e2e" else FULL
    geom, T, hw = GAEGeometry(31, 8, 2), int(g["T"]), int(g["hw"])
    gsd = synth.gae_state_dict(geom, 51)
    usd = synth.unet_state_dict(cfg, 52)
    tab = O.schedule_tables(O.beta_schedule("cosine", T, 1e-6, 1e-2))
    sr = synth.sr_cube(1, 31, hw, seed=53)
    hr = synth.sr_cube(1, 31, hw, seed=54)
    x_T, tape = synth.noise_tape(geom.G, T, 3, hw, hw, seed=55)
    with torch.no_grad():
        cube = O.sr_cube(usd, cfg.as_dict(), tab, gsd, geom.as_dict(), sr,
                         [x_T[i:i + 1] for i in range(geom.G)], lambda gi, i: tape[gi:gi + 1, T - 1 - i])
    assert rel(cube.numpy(), g["cube"]) < 1e-5
    pred = cube[0].permute(1, 2, 0).numpy()
    true = hr[0].permute(1, 2, 0).numpy()
    assert abs(O.sam_deg(true, pred) - float(g["sam"])) < 1e-3          # reference compare_sam (python loop)
    assert abs(O.mpsnr(true, pred) - float(g["mpsnr"])) < 1e-4


@pytest.mark.parametrize("loss_type", ["l1", "l2"])
def test_train_step_matches_reference(golden, loss_type):
    """SURVEY 8f row N2 (next scope row): loss and gradients of one training step of the unmodified reference
    (p_losses + optimize_parameters' normalisation) with the timestep, the per-sample noise levels and the noise
    injected - the pin a CUDA backward will have to meet.  Vectors: oracle/make_golden_train.py."""
    from hsi_dmgasr_b200.spec import UNetConfig
    g = golden("train_step.npz")
    cfg = UNetConfig(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2), attn_res=(8,),
                     res_blocks=1, dropout=0.0, image_size=16)
    sd = synth.unet_state_dict(cfg, 51)
    hr, sr, noise = rand((3, 3, 16, 16), 61), rand((3, 3, 16, 16), 62), rand((3, 3, 16, 16), 63)
    loss_sum, loss, grads = O.train_step(sd, cfg.as_dict(), hr, sr, noise, g[f"{loss_type}.levels"], loss_type)
    assert abs(loss_sum - float(g[f"{loss_type}.loss_sum"])) < 1e-4 * abs(float(g[f"{loss_type}.loss_sum"]))
    assert abs(loss - float(g[f"{loss_type}.loss"])) < 1e-5
    checked = 0
    for k, gr in grads.items():
        want_norm = float(g[f"{loss_type}.gnorm.{k}"])
        assert abs(float(gr.double().norm()) - want_norm) <= 2e-4 * want_norm + 1e-9, k
        if f"{loss_type}.grad.{k}" in g:
            assert rel(gr.numpy(), g[f"{loss_type}.grad.{k}"]) < 2e-4, k
            checked += 1
    assert len(grads) == 124 and checked > 50


# ---- the steps either side of the path: MATLAB-style imresize and quality_assessment ----------------
from oracle import make_golden_prepost as GP  # noqa: E402  (case tables and seeded inputs only; never touches the reference here)


@pytest.mark.parametrize("name", list(GP.IMRESIZE_CASES))
def test_imresize_matches_reference_imsize(golden, name):
    """oracle.imresize_matlab against GAE/imsize.py's outputs for the dataset code's own calls (HStest.py:44-45)."""
    g = golden("prepost.npz")
    seed, shape, first, second, method = GP.IMRESIZE_CASES[name]
    x = GP.imresize_input(seed, shape)
    ms = O.imresize_matlab(x, first, method)
    ref = g[f"imresize.{name}.first"]
    assert ms.shape == ref.shape and ms.dtype == np.float64
    assert float(np.abs(ms - ref).max()) < 1e-13
    if second is not None:
        lms = O.imresize_matlab(ms, second, method)
        assert float(np.abs(lms - g[f"imresize.{name}.second"]).max()) < 1e-13
        lms32 = O.imresize_matlab(ms.astype(np.float32), second, method)
        assert float(np.abs(lms32 - g[f"imresize.{name}.second_from_f32"]).max()) < 1e-13


def test_imresize_scalar_scale_shape_rule(golden):
    g = golden("prepost.npz")
    x = GP.imresize_input(16, (20, 28, 2))
    ref = g["imresize.scalar_scale.first"]
    assert ref.shape == (6, 9, 2)                       # ceil(0.3 * 20), ceil(0.3 * 28)  (imsize.py:3-7)
    assert float(np.abs(O.imresize_matlab(x, scalar_scale=0.3) - ref).max()) < 1e-13


@pytest.mark.parametrize("name", list(GP.ASSESS_CASES))
def test_assessment_indices_match_reference_eval_hsi(golden, name):
    """ERGAS / SAM / CrossCorrelation / RMSE of the oracle against eval_hsi.py's own functions (float32 sums there)."""
    g = golden("prepost.npz")
    truth, pred = GP.assess_inputs(*GP.ASSESS_CASES[name])
    rows = O.cube_assessment(torch.from_numpy(truth), torch.from_numpy(pred), 4.0)
    for (m, ss, e, s, c, r), ref in zip(rows, g[f"assess.{name}"]):
        assert abs(e - ref[0]) < 2e-4 * ref[0]
        assert abs(s - ref[1]) < 1e-3
        assert abs(c - ref[2]) < 1e-5
        assert abs(r - ref[3]) < 1e-6
        assert 0.0 < ss < 1.0 and m > 10.0


def test_mssim_restatement_properties():
    """MSSIM is restated from skimage's algorithm (skimage is absent: unpinned); check what the definition implies."""
    rng = np.random.default_rng(5)
    a = rng.random((20, 18, 3), dtype=np.float32)
    assert abs(O.mssim(a, a) - 1.0) < 1e-12
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    assert O.mssim(a, b) < 1.0 and abs(O.mssim(a, b) - O.mssim(b, a)) < 1e-12
    # direct evaluation of one window against the formula
    x, y = a[:7, :7, 0].astype(np.float64), b[:7, :7, 0].astype(np.float64)
    ux, uy = x.mean(), y.mean()
    vx, vy, vxy = x.var(ddof=1), y.var(ddof=1), ((x - ux) * (y - uy)).sum() / 48
    s = ((2 * ux * uy + 1e-4) * (2 * vxy + 9e-4)) / ((ux * ux + uy * uy + 1e-4) * (vx + vy + 9e-4))
    one = O.mssim(a[:7, :7, :1], b[:7, :7, :1])
    assert abs(one - s) < 1e-12


def test_compiled_reference_modules_agree_with_the_oracle():
    """oracle/_ref holds the unmodified reference's unet.py / diffusion.py as sourceless bytecode (oracle/build_ref.py,
    run by __graft_entry__.build() where /root/reference exists); bench.py's reference arm times them.  They and the
    oracle port must be the same function."""
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    unet_mod, diff_mod = ref
    cfg = SMALL
    sd = synth.unet_state_dict(cfg, 3)
    net = unet_mod.UNet(in_channel=cfg.in_channel, out_channel=cfg.out_channel, norm_groups=cfg.norm_groups,
                        inner_channel=cfg.inner_channel, channel_mults=list(cfg.channel_mults), attn_res=list(cfg.attn_res),
                        res_blocks=cfg.res_blocks, dropout=cfg.dropout, image_size=cfg.image_size).eval()
    net.load_state_dict(sd, strict=True)
    x, lv = rand((2, 6, 16, 16), 5), torch.tensor([[0.3], [0.9]])
    with torch.no_grad():
        want = net(x, lv)
        got = O.unet_forward(sd, cfg.as_dict(), x, lv)
    assert rel(got, want) < 2e-6
    gd = diff_mod.GaussianDiffusion(net, image_size=16, channels=3, conditional=True)
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=20, linear_start=1e-6, linear_end=1e-2), "cpu")
    tab = O.schedule_tables(O.beta_schedule("cosine", 20, 1e-6, 1e-2))
    assert np.array_equal(gd.betas.numpy(), tab["betas"].astype(np.float32))


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_oracle_prepost_against_live_reference_on_random_shapes():
    """Where the reference tree is mounted (the build container) the restatements meet the reference itself, not only
    the committed vectors: imresize over random sizes / scales / both kernels, ERGAS / CC / RMSE / SAM on random cubes."""
    import sys
    from oracle import make_golden as MG
    _, eval_hsi, _, _ = MG.import_reference()
    sys.path.insert(0, __import__("os").path.join(MG.REF, "GAE"))
    import imsize
    rng = np.random.default_rng(2024)
    for _ in range(12):
        h, w, c = int(rng.integers(5, 40)), int(rng.integers(5, 40)), int(rng.integers(1, 4))
        oh, ow = int(rng.integers(2, 60)), int(rng.integers(2, 60))
        method = ["bicubic", "bilinear"][int(rng.integers(0, 2))]
        x = rng.random((h, w, c), dtype=np.float32)
        assert float(np.abs(O.imresize_matlab(x, (oh, ow), method) - imsize.imresize(x, output_shape=(oh, ow), method=method)).max()) < 1e-12
        sc = float(rng.uniform(0.2, 3.0))
        assert float(np.abs(O.imresize_matlab(x, scalar_scale=sc, method=method) - imsize.imresize(x, scalar_scale=sc, method=method)).max()) < 1e-12
    x2 = rng.random((9, 11), dtype=np.float32)                                   # 2-D input keeps its rank (imsize.py:147-156)
    assert O.imresize_matlab(x2, (18, 5)).shape == imsize.imresize(x2, output_shape=(18, 5)).shape == (18, 5)
    for _ in range(4):
        h, w, c = int(rng.integers(8, 24)), int(rng.integers(8, 24)), int(rng.integers(2, 12))
        t = rng.random((h, w, c), dtype=np.float32)
        p = np.clip(t + 0.1 * rng.standard_normal((h, w, c)).astype(np.float32), 0, 1)
        assert abs(O.ergas(t, p, 4) - eval_hsi.compare_ergas(t, p, 4)) < 3e-4 * eval_hsi.compare_ergas(t, p, 4)
        assert abs(O.cross_correlation(t, p) - eval_hsi.compare_corr(t, p)) < 1e-5
        assert abs(O.rmse(t, p) - eval_hsi.compare_rmse(t, p)) < 1e-6
        assert abs(O.sam_deg(t, p) - eval_hsi.compare_sam(t, p)) < 1e-3
