"""Sampling-loop parity (diffusion.py:142-201) on the B200 with an injected noise tape."""
import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth
from tests.cfgs import SMALL
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu
SCHED = dict(schedule="cosine", linear_start=1e-6, linear_end=1e-2)


def make(precision, T, seed=21):
    net = UNet(in_channel=6, out_channel=3, inner_channel=SMALL.inner_channel, norm_groups=SMALL.norm_groups,
               channel_mults=SMALL.channel_mults, attn_res=SMALL.attn_res, res_blocks=SMALL.res_blocks,
               dropout=SMALL.dropout, image_size=SMALL.image_size, precision=precision)
    net.load_state_dict(synth.unet_state_dict(SMALL, seed))
    gd = GaussianDiffusion(net, image_size=16, channels=3, conditional=True).cuda().eval()
    gd.set_new_noise_schedule(dict(SCHED, n_timestep=T), torch.device("cuda"))
    return gd


def inputs(n, T, hw):
    cond = torch.from_numpy(np.random.default_rng(31).standard_normal((n, 3, hw, hw), dtype=np.float32)).cuda()
    x_T, tape = synth.noise_tape(n, T, 3, hw, hw, seed=32)
    return cond, x_T.cuda(), tape.cuda()


def test_schedule_buffers_match_reference(golden):
    g = golden("schedules.npz")
    gd = make("fp32", 50)
    for k, v in gd.state_dict().items():
        if not k.startswith("denoise_fn."):
            assert np.array_equal(v.cpu().numpy(), g[f"cosine50.{k}"], equal_nan=True), k
    assert np.array_equal(gd.sqrt_alphas_cumprod_prev, g["cosine50.sqrt_alphas_cumprod_prev"])


def test_stepwise_loop_matches_reference(golden):
    g = golden("sample_loop.npz")
    T, n, hw = int(g["T"]), 2, 16
    gd = make("fp32", T)
    cond, x, tape = inputs(n, T, hw)
    for j, i in enumerate(reversed(range(T))):
        eps = gd.predict_noise(x, i, cond)
        assert rel_l2(eps, torch.from_numpy(g["eps"][j])) < 1e-4, f"eps at loop index {i}"
        x = gd.p_sample(x, i, condition_x=cond, noise=tape[:, j] if i > 0 else None)
        assert rel_l2(x, torch.from_numpy(g["x"][j])) < 1e-4, f"x after loop index {i}"


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_graph_sampler_matches_reference(golden, precision, tol):
    g = golden("sample_loop.npz")
    T, n, hw = int(g["T"]), 2, 16
    gd = make(precision, T)
    cond, x_T, tape = inputs(n, T, hw)
    ret_all = gd.super_resolution(cond, continous=True, x_T=x_T, noise_tape=tape)
    assert ret_all.shape == g["ret_all"].shape
    assert rel_l2(ret_all, torch.from_numpy(g["ret_all"])) < tol
    last = gd.super_resolution(cond, continous=False, x_T=x_T, noise_tape=tape)
    assert last.shape == (3, hw, hw)                                  # the reference's ret_img[-1] quirk
    assert rel_l2(last, torch.from_numpy(g["ret_last"])) < tol
    full = gd.super_resolution(cond, continous=False, x_T=x_T, noise_tape=tape, return_all=True)
    assert full.shape == cond.shape and torch.equal(full[-1], last)
    # replaying the cached graph gives the same bits
    again = gd.super_resolution(cond, continous=False, x_T=x_T, noise_tape=tape, return_all=True)
    assert torch.equal(again, full)


def test_builtin_noise_generator_is_standard_normal_and_seeded():
    T = 8
    gd = make("fp32", T)
    cond, x_T, _ = inputs(4, T, 16)
    a = gd.super_resolution(cond, x_T=x_T, return_all=True, seed=7)
    b = gd.super_resolution(cond, x_T=x_T, return_all=True, seed=7)
    c = gd.super_resolution(cond, x_T=x_T, return_all=True, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.isfinite(a).all()
    # zero weights -> eps == 0; with x_T == 0 the first snapshot is exactly sigma_{T-1} * z, exposing the raw draws
    with torch.no_grad():
        for p in gd.denoise_fn.parameters():
            p.zero_()
    big = torch.zeros(64, 3, 64, 64, device="cuda")
    snaps = gd.super_resolution(big, continous=True, x_T=torch.zeros_like(big), seed=123)
    first = snaps[64:128]
    sigma = float(torch.exp(0.5 * gd.posterior_log_variance_clipped[T - 1]))
    z = (first / sigma).flatten()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05                       # kurtosis of a standard normal
    assert abs(float((z[:-1] * z[1:]).mean())) < 5e-3                     # neighbouring draws are uncorrelated


def test_posterior_step_formula():
    T = 50
    gd = make("fp32", T)
    x = torch.randn(2, 3, 16, 16, device="cuda") * 2
    eps = torch.randn_like(x)
    z = torch.randn_like(x)
    from hsi_dmgasr_b200 import _lib
    h = gd._native(x.device)
    for t in (0, 1, 17, T - 1):
        out = torch.empty_like(x)
        _lib.check(_lib.load().hsidm_posterior_step(h.ptr, t, x.data_ptr(), eps.data_ptr(), z.data_ptr() if t else None,
                                                    out.data_ptr(), x.numel(), _lib.stream_ptr(x.device)))
        x0 = (gd.sqrt_recip_alphas_cumprod[t] * x - gd.sqrt_recipm1_alphas_cumprod[t] * eps).clamp(-1, 1)
        mean, logvar = gd.q_posterior(x0, x, t)
        want = mean + (z if t else 0) * (0.5 * logvar).exp()
        assert torch.allclose(out, want, rtol=1e-5, atol=1e-6), t


def test_full_batch_sampling_is_reproducible_and_tracks_fp32():
    """BASELINE config C2's per-step batch (176 group latents @128x128, the production UNet) through the graph sampler
    for a short schedule: (1) two runs with the same seed give the same bits - with 148 persistent CTAs walking dozens
    of tiles each, any barrier-phase race in the tensor-core kernels shows up as a difference here; (2) the same
    sampling with GroupNorm as a separate pass (developer switch, variant 8) gives the same bits as the fused load
    path; (3) a 16-latent slice tracks this library's fp32 CUDA-core path through the whole reverse process."""
    from hsi_dmgasr_b200 import _lib
    from hsi_dmgasr_b200.spec import UNetConfig
    lib = _lib.load()
    full = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                      attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
    T, n = 6, 176

    def sampler(precision):
        net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                   attn_res=[16], res_blocks=2, dropout=0.2, image_size=128, precision=precision)
        net.load_state_dict(synth.unet_state_dict(full, 3))
        gd = GaussianDiffusion(net, image_size=128, channels=3, conditional=True).cuda().eval()
        gd.set_new_noise_schedule(dict(SCHED, n_timestep=T), torch.device("cuda"))
        return gd

    cond = torch.from_numpy(np.random.default_rng(41).standard_normal((n, 3, 128, 128), dtype=np.float32)).cuda()
    x0 = torch.from_numpy(np.random.default_rng(42).standard_normal((n, 3, 128, 128), dtype=np.float32)).cuda()
    gd = sampler("bf16")
    a = gd.p_sample_loop(cond, False, x_T=x0, return_all=True, seed=9).clone()   # step noise: the library's seeded generator
    b = gd.p_sample_loop(cond, False, x_T=x0, return_all=True, seed=9).clone()
    assert torch.isfinite(a).all()
    assert torch.equal(a, b), f"replay differs: max |diff| {float((a - b).abs().max()):.3e}"
    del gd
    try:
        lib.hsidm_debug_conv_mode(0, 8)
        gd = sampler("bf16")
        c = gd.p_sample_loop(cond, False, x_T=x0, return_all=True, seed=9).clone()
        del gd
    finally:
        lib.hsidm_debug_conv_mode(0, 0)
    assert torch.equal(a, c), f"fused GroupNorm differs from the two-pass form: max |diff| {float((a - c).abs().max()):.3e}"
    gd = sampler("fp32")
    x_T, tape = synth.noise_tape(16, T, 3, 128, 128, seed=10)
    ref = gd.super_resolution(cond[:16], continous=False, x_T=x_T.cuda(), noise_tape=tape.cuda(), return_all=True).clone()
    del gd
    torch.cuda.empty_cache()
    gd = sampler("bf16")
    got = gd.super_resolution(cond[:16], continous=False, x_T=x_T.cuda(), noise_tape=tape.cuda(), return_all=True)
    err = rel_l2(got, ref)
    print(f"176-latent replay identical; 16-latent bf16 vs fp32 over T={T}: rel-L2 {err:.3e}")
    assert err < 3e-2
