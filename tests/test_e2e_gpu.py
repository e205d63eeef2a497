"""End-to-end cube parity: encode -> batched T-step sampling of all groups -> decode -> clamp, against the
reference's sequential val loop (tests/golden/e2e.npz, produced by oracle/make_golden.py section 5).
Gates from BASELINE.json: |dMPSNR| <= 0.05 dB, |dSAM| <= 0.01 degrees."""
import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import GAE, GaussianDiffusion, SRPipeline, UNet, synth
from hsi_dmgasr_b200.metrics import mpsnr, sam_degrees
from hsi_dmgasr_b200.spec import GAEGeometry
from tests.cfgs import FULL, SMALL
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def build(precision, T, cfg=SMALL):
    geom = GAEGeometry(31, 8, 2)
    gae = GAE(n_subs=8, n_ovls=2, n_colors=31, n_feats=64)
    gae.load_state_dict(synth.gae_state_dict(geom, 51))
    net = UNet(in_channel=6, out_channel=3, inner_channel=cfg.inner_channel, norm_groups=cfg.norm_groups,
               channel_mults=cfg.channel_mults, attn_res=cfg.attn_res, res_blocks=cfg.res_blocks,
               dropout=cfg.dropout, image_size=cfg.image_size, precision=precision)
    net.load_state_dict(synth.unet_state_dict(cfg, 52))
    gd = GaussianDiffusion(net, image_size=16, channels=3, conditional=True).cuda().eval()
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), torch.device("cuda"))
    return SRPipeline(gd, gae.cuda().eval()), geom


SETUPS = {"e2e": SMALL, "e2e_full": FULL}      # oracle/make_golden.py section 5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", list(SETUPS))
def test_cube_and_metric_gates(golden, tag, precision):
    """fp32 mode: cube within 1e-4 of the reference and both metric gates with orders of margin.
    bf16 mode: dMPSNR <= 0.05 dB always; dSAM <= 0.01 deg on the full-UNet setup.  On the tiny random-weight setup the
    UNMODIFIED reference under torch bf16 autocast itself drifts 0.0187 deg (recorded in the fixture), so there the
    bar is "no worse than PyTorch's own bf16 path of the reference"."""
    g = golden(tag + ".npz")
    T, hw = int(g["T"]), int(g["hw"])
    pipe, geom = build(precision, T, SETUPS[tag])
    sr = synth.sr_cube(1, 31, hw, seed=53).cuda()
    hr = synth.sr_cube(1, 31, hw, seed=54)
    x_T, tape = synth.noise_tape(geom.G, T, 3, hw, hw, seed=55)
    cube, lat = pipe.super_resolve(sr, x_T=x_T.cuda(), noise_tape=tape.cuda(), return_latents=True)
    want = torch.from_numpy(g["cube"])
    true = hr[0].permute(1, 2, 0).numpy()
    pred = cube[0].permute(1, 2, 0).cpu().numpy()
    d_psnr = abs(mpsnr(true, pred) - float(g["mpsnr"]))
    d_sam = abs(sam_degrees(true, pred) - float(g["sam"]))
    print(f"{tag} {precision}: cube rel-L2 {rel_l2(cube, want):.3e} latents {rel_l2(lat, torch.from_numpy(g['latents'])):.3e} "
          f"dMPSNR {d_psnr:.5f} dB dSAM {d_sam:.5f} deg  (reference under torch bf16 autocast: "
          f"{float(g['autocast_dpsnr']):.5f} dB, {float(g['autocast_dsam']):.5f} deg)")
    assert d_psnr <= 0.05
    if precision == "fp32":
        assert rel_l2(cube, want) < 1e-4 and d_sam <= 1e-3
    else:
        assert d_sam <= max(0.01, float(g["autocast_dsam"]))
        assert rel_l2(cube, want) <= max(3e-2, 1.5 * float(g["autocast_cube_rel"]))
        if tag == "e2e_full":
            assert d_sam <= 0.01


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,cfg", [("small", SMALL), ("full", FULL)])
def test_headline_schedule_T2000_metric_gates(golden, tag, cfg, precision):
    """The same gates at the schedule length the bench is quoted on: T = 2000 cosine steps with an injected noise tape,
    against the unmodified reference's sequential val loop (oracle/make_golden_r2.py section `long`).  No autocast escape
    hatch: |dMPSNR| <= 0.05 dB and |dSAM| <= 0.01 deg in both precisions."""
    g = golden(f"e2e_T2000_{tag}.npz")
    T, hw = int(g["T"]), int(g["hw"])
    assert T == 2000
    pipe, geom = build(precision, T, cfg)
    sr = synth.sr_cube(1, 31, hw, seed=53).cuda()
    hr = synth.sr_cube(1, 31, hw, seed=54)
    draws = [synth.noise_tape(1, T, 3, hw, hw, seed=5500 + k) for k in range(geom.G)]      # per group, like the generator
    x_T = torch.cat([d[0] for d in draws]).cuda()
    tape = torch.cat([d[1] for d in draws]).cuda()
    cube, lat = pipe.super_resolve(sr, x_T=x_T, noise_tape=tape, return_latents=True)
    true = hr[0].permute(1, 2, 0).numpy()
    pred = cube[0].permute(1, 2, 0).cpu().numpy()
    d_psnr = abs(mpsnr(true, pred) - float(g["mpsnr"]))
    d_sam = abs(sam_degrees(true, pred) - float(g["sam"]))
    ctx = ""
    if "autocast_dsam" in g.files:
        ctx = f" (reference under torch bf16 autocast: {float(g['autocast_dpsnr']):.5f} dB, {float(g['autocast_dsam']):.5f} deg)"
    print(f"T=2000 {tag} {precision}: cube rel-L2 {rel_l2(cube, torch.from_numpy(g['cube'])):.3e} latents "
          f"{rel_l2(lat, torch.from_numpy(g['latents'])):.3e} dMPSNR {d_psnr:.5f} dB dSAM {d_sam:.5f} deg{ctx}")
    assert d_psnr <= 0.05 and d_sam <= 0.01


def test_host_buffer_call_and_batch_independence():
    pipe, geom = build("fp32", 5)
    sr = synth.sr_cube(3, 31, 16, seed=60)
    x_T, tape = synth.noise_tape(3 * geom.G, 5, 3, 16, 16, seed=61)
    out = pipe.super_resolve_host(sr, torch.device("cuda"), x_T=x_T.cuda(), noise_tape=tape.cuda())
    assert out.shape == sr.shape and not out.is_cuda and float(out.min()) >= 0 and float(out.max()) <= 1
    one = pipe.super_resolve(sr[1:2].cuda(), x_T=x_T[geom.G:2 * geom.G].cuda(), noise_tape=tape[geom.G:2 * geom.G].cuda())
    assert rel_l2(one, out[1:2]) < 1e-5
    # chunked sampling (max_latents) is the same computation
    pipe.max_latents = 4
    chunked = pipe.super_resolve(sr.cuda(), x_T=x_T.cuda(), noise_tape=tape.cuda())
    assert rel_l2(chunked, out) < 1e-5


def test_validation_driver_matches_per_cube_metrics():
    """pipeline.validate restates the reference's val loop (sr_gae.py:436-497): averages of the device-side MPSNR / SAM
    over the cubes equal the numpy eval_hsi metrics of the individually super-resolved, clamped cubes."""
    from hsi_dmgasr_b200 import metrics
    from hsi_dmgasr_b200.pipeline import validate
    from oracle import hsidm_oracle as O
    pipe, geom = build("fp32", 5)
    sr = synth.sr_cube(3, 31, 16, seed=70)
    hr = (sr + 0.03 * torch.from_numpy(np.random.default_rng(71).standard_normal(tuple(sr.shape), dtype=np.float32))).clamp(0, 1)
    x_T, tape = synth.noise_tape(3 * geom.G, 5, 3, 16, 16, seed=72)
    loader = [{"HR": hr[0:2], "SR": sr[0:2]}, {"HR": hr[2], "SR": sr[2]}]           # a batched and an unbatched entry
    draws = iter([(x_T[:2 * geom.G], tape[:2 * geom.G]), (x_T[2 * geom.G:], tape[2 * geom.G:])])

    # per-cube reference numbers, batch by batch with the matching slice of the injected noise
    want_m, want_s, want_all = [], [], []
    for b, (xt, tp) in zip(loader, draws):
        s = b["SR"] if b["SR"].dim() == 4 else b["SR"].unsqueeze(0)
        h = b["HR"] if b["HR"].dim() == 4 else b["HR"].unsqueeze(0)
        y = pipe.super_resolve(s.cuda(), x_T=xt.cuda(), noise_tape=tp.cuda()).cpu()
        for i in range(y.shape[0]):
            want_m.append(metrics.mpsnr(h[i].permute(1, 2, 0).numpy(), y[i].permute(1, 2, 0).numpy()))
            want_s.append(metrics.sam_degrees(h[i].permute(1, 2, 0).numpy(), y[i].permute(1, 2, 0).numpy()))
        want_all += O.cube_assessment(h, y, 4.0)          # (MPSNR, MSSIM, ERGAS, SAM, CrossCorrelation, RMSE) per cube
    got = validate(pipe, [{"HR": hr, "SR": sr}], torch.device("cuda"), x_T=x_T.cuda(), noise_tape=tape.cuda())
    assert got["cubes"] == 3
    assert abs(got["MPSNR"] - float(np.mean(want_m))) < 1e-3 and abs(got["SAM"] - float(np.mean(want_s))) < 1e-3
    # the reference loop sums every index of quality_assessment(gt, y, data_range=1., ratio=4) (sr_gae.py:487-491)
    mean_all = np.mean(np.asarray(want_all, dtype=np.float64), axis=0)
    for i, key in enumerate(("MPSNR", "MSSIM", "ERGAS", "SAM", "CrossCorrelation", "RMSE")):
        assert abs(got[key] - mean_all[i]) < 1e-3 * max(1.0, abs(mean_all[i])), (key, got[key], mean_all[i])
