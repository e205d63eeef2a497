"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE.  Run once in the build container (``python oracle/make_golden.py``); the GPU box never
sees /root/reference, only the committed vectors.  Weights and inputs are NOT stored: they are re-derived from
seeds by ``hsi_dmgasr_b200.synth`` on both sides, so the fixtures hold only reference outputs.

Import shims follow SURVEY.md 8c: empty stubs for modules missing from this image, the GAE loops of
AE.py:283-324 re-driven by hand because the originals hard-code 'cuda:0'.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HSIDM_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from hsi_dmgasr_b200 import synth  # noqa: E402
from hsi_dmgasr_b200.spec import GAEGeometry, UNetConfig  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SMALL = UNetConfig(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2),
                   attn_res=(8,), res_blocks=1, dropout=0.2, image_size=16)
# config/sr_sr3_16_128ae.json "model.unet" + "model.diffusion.image_size"
FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
# config/sr_sr3_64_512.json (norm_groups 16, no attn_res, res_blocks 1)
WIDE = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_mults=(1, 2, 4, 8, 16),
                  attn_res=(), res_blocks=1, dropout=0.0, image_size=128)


def import_reference():
    sys.path.insert(0, REF)
    for m in ["sewar", "h5py", "matplotlib", "matplotlib.pyplot", "turtle", "skimage", "skimage.metrics",
              "tensorboardX"]:
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["turtle"].forward = None
    sys.modules["skimage.metrics"].peak_signal_noise_ratio = None
    sys.modules["skimage.metrics"].structural_similarity = None
    import AE  # noqa
    import eval_hsi  # noqa
    from model.sr3_modules import diffusion, unet  # noqa
    return AE, eval_hsi, unet, diffusion


def ref_unet(unet_mod, cfg: UNetConfig, seed: int):
    net = unet_mod.UNet(in_channel=cfg.in_channel, out_channel=cfg.out_channel, norm_groups=cfg.norm_groups,
                        inner_channel=cfg.inner_channel, channel_mults=list(cfg.channel_mults),
                        attn_res=list(cfg.attn_res), res_blocks=cfg.res_blocks, dropout=cfg.dropout,
                        image_size=cfg.image_size)
    sd = synth.unet_state_dict(cfg, seed)
    assert list(net.state_dict().keys()) == list(sd.keys()), "state_dict key order drifted from the reference"
    net.load_state_dict(sd, strict=True)
    return net.eval()


def rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32))


def main():
    os.makedirs(OUT, exist_ok=True)
    AE, eval_hsi, unet_mod, diff_mod = import_reference()
    torch.set_grad_enabled(False)
    torch.manual_seed(0)

    # ---- 1. schedules ---------------------------------------------------------------------------------
    sched = {}
    for name, T in [("cosine", 20), ("cosine", 50), ("cosine", 2000), ("linear", 30), ("quad", 30),
                    ("warmup10", 30), ("warmup50", 30), ("const", 10), ("jsd", 10)]:
        gd = diff_mod.GaussianDiffusion(torch.nn.Identity(), image_size=16, channels=3)
        gd.set_new_noise_schedule(dict(schedule=name, n_timestep=T, linear_start=1e-6, linear_end=1e-2), "cpu")
        for k, v in gd.state_dict().items():
            sched[f"{name}{T}.{k}"] = v.numpy()
        sched[f"{name}{T}.sqrt_alphas_cumprod_prev"] = np.asarray(gd.sqrt_alphas_cumprod_prev, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **sched)

    # ---- 2. UNet forward, three configs ------------------------------------------------------------------
    out = {}
    for tag, cfg, seed, n, hw, lvls in [("small", SMALL, 11, 3, 16, [0.9, 0.35, 0.01]),
                                        ("full32", FULL, 12, 2, 32, [0.71, 0.05]),
                                        ("full128", FULL, 12, 1, 128, [0.5]),
                                        ("wide64", WIDE, 13, 1, 64, [0.2])]:
        net = ref_unet(unet_mod, cfg, seed)
        x = rand((n, 6, hw, hw), 1000 + seed)
        lv = torch.tensor(lvls, dtype=torch.float32).view(n, 1)
        taps = {}
        hooks = []
        for nm, mod in list(net.downs.named_children()):
            hooks.append(mod.register_forward_hook(lambda m, i, o, nm=nm: taps.__setitem__(f"downs.{nm}", o)))
        for nm, mod in list(net.mid.named_children()):
            hooks.append(mod.register_forward_hook(lambda m, i, o, nm=nm: taps.__setitem__(f"mid.{nm}", o)))
        for nm, mod in list(net.ups.named_children()):
            hooks.append(mod.register_forward_hook(lambda m, i, o, nm=nm: taps.__setitem__(f"ups.{nm}", o)))
        y = net(x, lv)
        for h in hooks:
            h.remove()
        out[f"{tag}.eps"] = y.numpy()
        out[f"{tag}.level"] = lv.numpy()
        if tag == "small":
            for k, v in taps.items():
                out[f"{tag}.tap.{k}"] = v.numpy()
        else:  # keep fixtures small: per-layer mean/std/abs-max fingerprints
            for k, v in taps.items():
                out[f"{tag}.fp.{k}"] = np.array([v.mean().item(), v.std().item(), v.abs().max().item()], np.float64)
        print(tag, "eps", tuple(y.shape), float(y.abs().mean()))
    np.savez_compressed(os.path.join(OUT, "unet_forward.npz"), **out)

    # ---- 3. sampling loop with an injected noise tape (small config, T=6, two images) ---------------------
    T, n, hw = 6, 2, 16
    net = ref_unet(unet_mod, SMALL, 21)
    gd = diff_mod.GaussianDiffusion(net, image_size=16, channels=3, conditional=True)
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), "cpu")
    gd.eval()
    cond = rand((n, 3, hw, hw), 31)
    x_T, tape = synth.noise_tape(n, T, 3, hw, hw, seed=32)
    draws = [x_T] + [tape[:, j] for j in range(T - 1)]
    rec = {"eps": [], "x": []}
    orig_randn, orig_like, orig_dn = torch.randn, torch.randn_like, gd.denoise_fn.forward

    def dn(x, lvl):
        e = orig_dn(x, lvl)
        rec["eps"].append(e.clone())
        return e
    it = iter(draws)
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    gd.denoise_fn.forward = dn
    orig_ps = gd.p_sample

    def ps(x, t, **kw):
        r = orig_ps(x, t, **kw)
        rec["x"].append(r.clone())
        return r
    gd.p_sample = ps
    try:
        ret_all = gd.super_resolution(cond, continous=True)
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like
    # continous=False path re-run to pin the "last element of last snapshot" quirk
    first = {k: list(v) for k, v in rec.items()}
    it = iter(draws)
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    try:
        ret_last = gd.super_resolution(cond, continous=False)
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like
    rec = first
    np.savez_compressed(os.path.join(OUT, "sample_loop.npz"), T=T, eps=torch.stack(rec["eps"]).numpy(),
                        x=torch.stack(rec["x"]).numpy(), ret_all=ret_all.numpy(), ret_last=ret_last.numpy())
    print("sample_loop", tuple(ret_all.shape), tuple(ret_last.shape))

    # ---- 4. GAE encode / decode with synthetic weights (all four band geometries) -------------------------
    g_out = {}
    for tag, geom, seed, hw in [("Cav", GAEGeometry(31, 8, 2), 41, 16), ("Chi", GAEGeometry(128, 16, 4), 42, 16),
                                ("Pav", GAEGeometry(102, 16, 4), 43, 16)]:
        gae = AE.GAE(AE.Encoder, AE.Decoder, n_subs=geom.n_subs, n_ovls=geom.n_ovls, n_colors=geom.n_colors,
                     n_feats=geom.n_feats)
        sd = synth.gae_state_dict(geom, seed)
        assert list(gae.state_dict().keys()) == list(sd.keys())
        gae.load_state_dict(sd, strict=True)
        gae.eval()
        assert (gae.G, gae.start_idx, gae.end_idx) == (geom.G, *geom.groups())
        x = synth.sr_cube(2, geom.n_colors, hw, seed=seed + 100)
        zs = [gae.Encoder(x[:, s:e]) for s, e in zip(gae.start_idx, gae.end_idx)]          # AE.py:316-324
        y = torch.zeros_like(x)
        cnt = torch.zeros(geom.n_colors)
        for g in range(gae.G):                                                              # AE.py:288-295
            s, e = gae.start_idx[g], gae.end_idx[g]
            y[:, s:e] += gae.Decoder(zs[g])
            cnt[s:e] = cnt[s:e] + 1
        y = y / cnt.unsqueeze(1).unsqueeze(2)
        y = gae.final(gae.trunk(y)) + y                                                     # AE.py:302-307
        g_out[f"{tag}.z"] = torch.stack(zs).numpy()
        g_out[f"{tag}.dec"] = y.numpy()
        g_out[f"{tag}.start"] = np.array(gae.start_idx)
        g_out[f"{tag}.end"] = np.array(gae.end_idx)
        print("gae", tag, gae.G, float(y.abs().mean()))
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **g_out)

    # ---- 5. end to end (val-loop restatement, sr_gae.py:456-475) + metrics ----------------------------------
    # Two setups. "small": tiny UNet, T=5 (strict fp32 parity).  "full": the 16_128ae UNet at 32x32, T=10 (bf16 gate).
    # For both, the same loop is ALSO run with the reference UNet under torch.autocast(bfloat16): how far PyTorch's own
    # bf16 path of the unmodified reference drifts from its fp32 result on these synthetic weights (context for the
    # 0.05 dB / 0.01 deg gates, which the reference meets on real checkpoints with 2x margin).
    def val_loop(cfg, T, hw, autocast):
        geom = GAEGeometry(31, 8, 2)
        gae = AE.GAE(AE.Encoder, AE.Decoder, n_subs=8, n_ovls=2, n_colors=31, n_feats=64)
        gae.load_state_dict(synth.gae_state_dict(geom, 51), strict=True)
        gae.eval()
        net = ref_unet(unet_mod, cfg, 52)
        gd = diff_mod.GaussianDiffusion(net, image_size=16, channels=3, conditional=True)
        gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), "cpu")
        gd.eval()
        if autocast:
            plain = gd.denoise_fn.forward

            def low_precision(x, t):
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    return plain(x, t).float()
            gd.denoise_fn.forward = low_precision
        sr = synth.sr_cube(1, 31, hw, seed=53)
        x_T, tape = synth.noise_tape(geom.G, T, 3, hw, hw, seed=55)
        zs = [gae.Encoder(sr[:, s:e]) for s, e in zip(gae.start_idx, gae.end_idx)]
        outs = []
        for g in range(gae.G):
            it = iter([x_T[g:g + 1]] + [tape[g:g + 1, j] for j in range(T - 1)])
            torch.randn = lambda *a, **k: next(it).clone()
            torch.randn_like = lambda *a, **k: next(it).clone()
            try:
                r = gd.super_resolution(zs[g], continous=False)
            finally:
                torch.randn, torch.randn_like = orig_randn, orig_like
            outs.append(r.unsqueeze(0))
        y = torch.zeros_like(sr)
        cnt = torch.zeros(31)
        for g in range(gae.G):
            s, e = gae.start_idx[g], gae.end_idx[g]
            y[:, s:e] += gae.Decoder(outs[g])
            cnt[s:e] = cnt[s:e] + 1
        y = y / cnt.unsqueeze(1).unsqueeze(2)
        y = gae.final(gae.trunk(y)) + y
        y[-1][y[-1] < 0] = 0
        y[-1][y[-1] > 1] = 1.0
        return y, torch.cat(outs)

    def metrics(y, hw):
        hr = synth.sr_cube(1, 31, hw, seed=54)
        pred = y[0].permute(1, 2, 0).numpy()
        true = hr[0].permute(1, 2, 0).numpy()
        # compare_mpsnr needs skimage (absent); its definition is 10*log10(R^2/MSE) per band, averaged
        mse = ((true.astype(np.float64) - pred.astype(np.float64)) ** 2).mean(axis=(0, 1))
        return float(eval_hsi.compare_sam(true, pred)), float(np.mean(10 * np.log10(1.0 / mse)))

    for tag, cfg, T, hw in [("e2e", SMALL, 5, 16), ("e2e_full", FULL, 10, 32)]:
        y, lat = val_loop(cfg, T, hw, autocast=False)
        sam, psnr = metrics(y, hw)
        y16, _ = val_loop(cfg, T, hw, autocast=True)
        sam16, psnr16 = metrics(y16, hw)
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), cube=y.numpy(), latents=lat.numpy(), sam=np.float64(sam),
                            mpsnr=np.float64(psnr), T=T, hw=hw, autocast_dsam=np.float64(abs(sam16 - sam)),
                            autocast_dpsnr=np.float64(abs(psnr16 - psnr)),
                            autocast_cube_rel=np.float64(float((y16 - y).norm() / y.norm())))
        print(tag, "sam", sam, "mpsnr", psnr, "| reference under bf16 autocast: dSAM", abs(sam16 - sam), "dPSNR", abs(psnr16 - psnr))


if __name__ == "__main__":
    main()
