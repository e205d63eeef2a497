"""Round-2 golden vectors, produced by running the UNMODIFIED reference from /root/reference on CPU.

TEST INFRASTRUCTURE (like make_golden.py; same import shims).  Sections, selectable on the command line:

  ckpt     the four shipped GAE checkpoints re-saved as tensors-only files (hsi_dmgasr_b200.save_gae_state format:
           state_dict + band-group geometry, loads with weights_only=True) -> tests/golden/gae_ckpt/GAE_4_*.state.pth
  gae128   reference Encoder / Decoder / trunk of those checkpoints on 128x128 cubes (B = 2 for Cav) -> gae128.npz
  dropin   the reference's own call pattern of the validation loop (sr_gae.py:444-475): Model.create_model(opt) with
           config/sr_sr3_16_128ae.json, per band group feed_data / test / get_current_visuals, GAE_4_Cav.pth decode,
           clamp - on one 31-band 32x32 cube, val schedule (T = 20), injected noise -> dropin.npz
  c4       one forward of the sr_sr3_64_512.json UNet at 512x512 (BASELINE configs[3]) -> c4_512.npz
  long     the same loop at the HEADLINE schedule length (T = 2000, cosine) with injected noise, for the tiny UNet at
           16x16 and the full 16_128ae UNet at 32x32, plus the reference's own drift under torch bf16 autocast for
           the tiny one -> e2e_T2000_small.npz, e2e_T2000_full.npz

  python oracle/make_golden_r2.py ckpt gae128 dropin long
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import make_golden as MG  # noqa: E402  (shims, ref_unet, SMALL/FULL)
from hsi_dmgasr_b200 import synth  # noqa: E402
from hsi_dmgasr_b200.spec import GAE_PRESETS, GAEGeometry  # noqa: E402

REF = MG.REF
OUT = MG.OUT
CKPT_DIR = os.path.join(OUT, "gae_ckpt")


def load_reference_gae(AE, name):
    """torch.load of the whole-module pickle with '__main__' resolved to the reference's AE module."""
    import io
    import pickle

    class U(pickle.Unpickler):
        def find_class(self, module, cls):
            if module == "__main__":
                return getattr(AE, cls)
            return super().find_class(module, cls)

    pm = types.SimpleNamespace(Unpickler=U, load=lambda f, **kw: U(f, **kw).load(),
                               loads=lambda b, **kw: U(io.BytesIO(b), **kw).load(), dump=pickle.dump, dumps=pickle.dumps,
                               Pickler=pickle.Pickler, PickleError=pickle.PickleError, UnpicklingError=pickle.UnpicklingError,
                               __name__="pickle")
    gae = torch.load(os.path.join(REF, "GAE_pretrained", f"GAE_4_{name}.pth"), map_location="cpu", weights_only=False,
                     pickle_module=pm)
    return gae.eval()


def ref_encode(gae, x):
    return [gae.Encoder(x[:, s:e]) for s, e in zip(gae.start_idx, gae.end_idx)]                   # AE.py:316-324


def ref_decode(gae, x, zs):
    y = torch.zeros_like(x)
    cnt = torch.zeros(x.shape[1])
    for g in range(gae.G):                                                                          # AE.py:288-295
        s, e = gae.start_idx[g], gae.end_idx[g]
        y[:, s:e] += gae.Decoder(zs[g])
        cnt[s:e] = cnt[s:e] + 1
    y = y / cnt.unsqueeze(1).unsqueeze(2)
    return gae.final(gae.trunk(y)) + y                                                              # AE.py:302-307


def section_ckpt(AE):
    os.makedirs(CKPT_DIR, exist_ok=True)
    for name, geom in GAE_PRESETS.items():
        gae = load_reference_gae(AE, name)
        assert (gae.G, list(gae.start_idx), list(gae.end_idx)) == (geom.G, *geom.groups())
        torch.save({"format": "hsidm-gae-state-v1",
                    "geometry": {"n_colors": geom.n_colors, "n_subs": geom.n_subs, "n_ovls": geom.n_ovls, "n_feats": geom.n_feats},
                    "state_dict": {k: v.detach().cpu().clone() for k, v in gae.state_dict().items()}},
                   os.path.join(CKPT_DIR, f"GAE_4_{name}.state.pth"))
        print("ckpt", name, sum(p.numel() for p in gae.parameters()))


# (name, batch, cube seed, spatial stride at which z / dec are stored)
GAE128 = [("Cav", 2, 301, 1, 2), ("Har", 1, 302, 2, 2), ("Chi", 1, 303, 2, 2), ("Pav", 1, 304, 2, 2)]


def section_gae128(AE):
    out = {}
    for name, b, seed, zs_stride, dec_stride in GAE128:
        gae = load_reference_gae(AE, name)
        geom = GAE_PRESETS[name]
        x = synth.sr_cube(b, geom.n_colors, 128, seed=seed)
        t0 = time.time()
        zs = ref_encode(gae, x)
        y = ref_decode(gae, x, zs)
        z = torch.stack(zs)                                  # [G,B,3,128,128]
        out[f"{name}.z"] = z[..., ::zs_stride, ::zs_stride].contiguous().numpy()
        out[f"{name}.dec"] = y[..., ::dec_stride, ::dec_stride].contiguous().numpy()
        out[f"{name}.z_mean"] = z.double().mean(dim=(-1, -2)).numpy()         # full-resolution fingerprints
        out[f"{name}.z_rms"] = z.double().pow(2).mean(dim=(-1, -2)).sqrt().numpy()
        out[f"{name}.dec_mean"] = y.double().mean(dim=(-1, -2)).numpy()
        out[f"{name}.dec_rms"] = y.double().pow(2).mean(dim=(-1, -2)).sqrt().numpy()
        out[f"{name}.meta"] = np.array([b, seed, zs_stride, dec_stride])
        print("gae128", name, tuple(z.shape), tuple(y.shape), f"{time.time() - t0:.1f}s", float(z.abs().max()), float(y.abs().max()))
    np.savez_compressed(os.path.join(OUT, "gae128.npz"), **out)


def patched_randn(draws):
    it = iter(draws)
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()


ORIG = (torch.randn, torch.randn_like)


def restore_randn():
    torch.randn, torch.randn_like = ORIG


DROPIN = dict(hw=32, sr_seed=401, hr_seed=402, noise_seed=403, unet_seed=404)


def section_dropin(AE):
    """sr_gae.py:444-475 with the reference's own Model.create_model / DDPM, its config file and GAE_4_Cav.pth."""
    sys.path.insert(0, REF)
    import core.logger as Logger
    import model as Model
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="hsidm_dropin_")
    os.chdir(tmp)                                   # Logger.parse creates experiments/... under the cwd
    try:
        args = types.SimpleNamespace(phase="val", config=os.path.join(REF, "config", "sr_sr3_16_128ae.json"), gpu_ids=None,
                                     enable_wandb=False, debug=False)
        keep = os.environ.get("CUDA_VISIBLE_DEVICES")
        opt = Logger.parse(args)
        if keep is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = keep
        opt = Logger.dict_to_nonedict(opt)
        opt["gpu_ids"] = None                        # CPU container
        # the UNet checkpoint the config points at does not ship: write seeded weights where load_network looks
        from hsi_dmgasr_b200.schedule import diffusion_buffers, make_beta_schedule
        sd = {"denoise_fn." + k: v for k, v in synth.unet_state_dict(MG.FULL, DROPIN["unet_seed"]).items()}
        opt["path"]["resume_state"] = os.path.join(tmp, "I0_E0")
        torch.save(sd, opt["path"]["resume_state"] + "_gen.pth")
        torch.manual_seed(0)
        diffusion = Model.create_model(opt)
        diffusion.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], schedule_phase="val")      # sr_gae.py:206-207
        T = diffusion.netG.num_timesteps
        gae = load_reference_gae(AE, "Cav")
        hw = DROPIN["hw"]
        sr = synth.sr_cube(1, 31, hw, seed=DROPIN["sr_seed"])
        hr = synth.sr_cube(1, 31, hw, seed=DROPIN["hr_seed"])
        x_T, tape = synth.noise_tape(gae.G, T, 3, hw, hw, seed=DROPIN["noise_seed"])
        val_data = {"HR": hr, "SR": sr.clone()}
        row_data = val_data["SR"]
        zs = ref_encode(gae, val_data["SR"])
        new_list = []
        for i in range(len(zs)):
            val_data["SR"] = zs[i]
            diffusion.feed_data(val_data)
            patched_randn([x_T[i:i + 1]] + [tape[i:i + 1, j] for j in range(T - 1)])
            try:
                diffusion.test(continous=False)
            finally:
                restore_randn()
            visuals = diffusion.get_current_visuals()
            visuals["SR"] = torch.unsqueeze(visuals["SR"], 0)
            new_list.append(visuals["SR"])
        with torch.no_grad():
            y = ref_decode(gae, row_data, new_list)
        y[-1][y[-1] < 0] = 0
        y[-1][y[-1] > 1] = 1.0
        # the dropped-key behaviour of load_network (model.py:189-192) is part of what this golden pins: the first conv
        # weight and the last conv stay at their torch.manual_seed(0) default init; store them so the GPU side can match
        own = diffusion.netG.state_dict()
        np.savez_compressed(os.path.join(OUT, "dropin.npz"), cube=y.detach().numpy(), latents=torch.cat(new_list).numpy(), T=T,
                            hw=hw, first_w=own["denoise_fn.downs.0.weight"].numpy(),
                            last_w=own["denoise_fn.final_conv.block.3.weight"].numpy(),
                            last_b=own["denoise_fn.final_conv.block.3.bias"].numpy(),
                            inf=visuals["INF"].numpy(), hr=visuals["HR"].numpy())
        print("dropin", tuple(y.shape), "T", T, float(y.mean()))
    finally:
        os.chdir(cwd)


def section_long(AE, eval_hsi, unet_mod, diff_mod):
    def val_loop(cfg, T, hw, autocast, threads=None):
        geom = GAEGeometry(31, 8, 2)
        gae = AE.GAE(AE.Encoder, AE.Decoder, n_subs=8, n_ovls=2, n_colors=31, n_feats=64)
        gae.load_state_dict(synth.gae_state_dict(geom, 51), strict=True)
        gae.eval()
        net = MG.ref_unet(unet_mod, cfg, 52)
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
        zs = ref_encode(gae, sr)
        outs = []
        for g in range(gae.G):
            x_T, tape = synth.noise_tape(1, T, 3, hw, hw, seed=5500 + g)        # per group, so the tape stays small
            patched_randn([x_T] + [tape[:, j] for j in range(T - 1)])
            try:
                r = gd.super_resolution(zs[g], continous=False)
            finally:
                restore_randn()
            outs.append(r.unsqueeze(0))
        y = ref_decode(gae, sr, outs)
        y[-1][y[-1] < 0] = 0
        y[-1][y[-1] > 1] = 1.0
        return y, torch.cat(outs)

    def metrics(y, hw):
        hr = synth.sr_cube(1, 31, hw, seed=54)
        pred = y[0].permute(1, 2, 0).numpy()
        true = hr[0].permute(1, 2, 0).numpy()
        mse = ((true.astype(np.float64) - pred.astype(np.float64)) ** 2).mean(axis=(0, 1))
        return float(eval_hsi.compare_sam(true, pred)), float(np.mean(10 * np.log10(1.0 / mse)))

    which = os.environ.get("HSIDM_LONG", "small,full").split(",")
    for tag, cfg, T, hw, with_autocast in [("small", MG.SMALL, 2000, 16, True), ("full", MG.FULL, 2000, 32, True)]:
        if tag not in which:
            continue
        t0 = time.time()
        y, lat = val_loop(cfg, T, hw, False)
        sam, psnr = metrics(y, hw)
        print(f"long {tag}: fp32 {time.time() - t0:.0f}s sam {sam} mpsnr {psnr}", flush=True)
        extra = {}
        if with_autocast and os.environ.get("HSIDM_LONG_AUTOCAST", "1") == "1":
            t0 = time.time()
            y16, _ = val_loop(cfg, T, hw, True)
            sam16, psnr16 = metrics(y16, hw)
            extra = dict(autocast_dsam=np.float64(abs(sam16 - sam)), autocast_dpsnr=np.float64(abs(psnr16 - psnr)),
                         autocast_cube_rel=np.float64(float((y16 - y).norm() / y.norm())))
            print(f"long {tag}: autocast {time.time() - t0:.0f}s dSAM {abs(sam16 - sam)} dPSNR {abs(psnr16 - psnr)}", flush=True)
        np.savez_compressed(os.path.join(OUT, f"e2e_T2000_{tag}.npz"), cube=y.numpy(), latents=lat.numpy(), sam=np.float64(sam),
                            mpsnr=np.float64(psnr), T=T, hw=hw, **extra)


def section_c4(unet_mod):
    """BASELINE configs[3]: the sr_sr3_64_512.json UNet (inner 64, mults 1-2-4-8-16, norm_groups 16, mid-block attention only)
    on one 512x512 latent: eps of the unmodified reference, stored on the stride-4 pixel lattice plus full-resolution
    per-channel mean / rms and per-row-block fingerprints -> c4_512.npz."""
    net = MG.ref_unet(unet_mod, MG.WIDE, 13)
    x = MG.rand((1, 6, 512, 512), 2013)
    lv = torch.tensor([[0.37]], dtype=torch.float32)
    t0 = time.time()
    y = net(x, lv)
    print(f"c4 512x512 forward {time.time() - t0:.0f}s", tuple(y.shape), float(y.abs().mean()))
    np.savez_compressed(os.path.join(OUT, "c4_512.npz"), eps_s4=y[..., ::4, ::4].contiguous().numpy(), level=lv.numpy(),
                        mean=y.double().mean(dim=(-1, -2)).numpy(), rms=y.double().pow(2).mean(dim=(-1, -2)).sqrt().numpy(),
                        block_rms=y.double().pow(2).view(1, 3, 16, 32, 16, 32).mean(dim=(3, 5)).sqrt().numpy())


def main():
    sections = sys.argv[1:] or ["ckpt", "gae128", "dropin", "long", "c4"]
    AE, eval_hsi, unet_mod, diff_mod = MG.import_reference()
    torch.set_grad_enabled(False)
    if "ckpt" in sections:
        section_ckpt(AE)
    if "gae128" in sections:
        section_gae128(AE)
    if "dropin" in sections:
        section_dropin(AE)
    if "long" in sections:
        section_long(AE, eval_hsi, unet_mod, diff_mod)
    if "c4" in sections:
        section_c4(unet_mod)


if __name__ == "__main__":
    main()
