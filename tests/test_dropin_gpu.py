"""The drop-in claim, exercised on the GPU: the reference's OWN call pattern of the validation loop
(sr_gae.py:444-475) - ``import model as Model`` / ``Model.create_model(opt)`` / ``torch.load('GAE_4_*.pth')`` /
per band group ``feed_data`` -> ``test`` -> ``get_current_visuals`` / ``model_GAE.decode`` / clamp - run unmodified on
this package through ``compat.install()``, with the REAL shipped GAE checkpoint (tests/golden/gae_ckpt, written by
oracle/make_golden_r2.py from GAE_pretrained/GAE_4_Cav.pth) and compared with what the unmodified reference produced
for the same weights, inputs and injected noise (tests/golden/dropin.npz)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import synth
from tests.cfgs import FULL
from tests.conftest import GOLDEN
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu

# config/sr_sr3_16_128ae.json: the keys the model layer reads ("datasets", "train", "wandb" are the driver's)
OPT = {
    "name": "Pav3-srae", "phase": "val", "gpu_ids": [0], "distributed": False,
    "path": {"log": "logs", "tb_logger": "tb_logger", "results": "results", "checkpoint": "checkpoint", "resume_state": None},
    "model": {
        "which_model_G": "sr3", "finetune_norm": False,
        "unet": {"in_channel": 6, "out_channel": 3, "inner_channel": 64, "channel_multiplier": [1, 2, 4, 8, 8],
                 "attn_res": [16], "res_blocks": 2, "dropout": 0.2},
        "beta_schedule": {"train": {"schedule": "cosine", "n_timestep": 20, "linear_start": 1e-6, "linear_end": 1e-2},
                          "val": {"schedule": "cosine", "n_timestep": 20, "linear_start": 1e-6, "linear_end": 1e-2}},
        "diffusion": {"image_size": 128, "channels": 3, "conditional": True}},
}
DROPIN = dict(hw=32, sr_seed=401, hr_seed=402, noise_seed=403, unet_seed=404)      # oracle/make_golden_r2.py


@pytest.fixture()
def reference_names():
    """compat.install() with a stand-in for the driver's __main__; restores sys.modules afterwards."""
    from hsi_dmgasr_b200 import compat
    saved = {k: sys.modules.get(k) for k in ("model", "model.model", "model.networks", "model.sr3_modules",
                                              "model.sr3_modules.unet", "model.sr3_modules.diffusion", "AE", "common")}
    main = sys.modules["__main__"]
    had = {n: getattr(main, n, None) for n in ("GAE", "Encoder", "Decoder", "BranchUnit", "SSPN", "SSB")}
    compat.install()
    yield
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for n, v in had.items():
        if v is None and hasattr(main, n):
            delattr(main, n)


def _patched_randn(draws):
    it = iter(draws)
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_reference_val_loop_runs_unmodified_on_the_native_path(reference_names, tmp_path, golden, precision, tol):
    from hsi_dmgasr_b200 import config, gae as G, load_gae, set_default_precision
    import model as Model                       # the reference's import line (sr_gae.py:9), resolved by compat.install()

    g = golden("dropin.npz")
    # -- files the driver expects on disk: the UNet checkpoint prefix and the whole-module GAE pickle ----------------------
    sd = {"denoise_fn." + k: v for k, v in synth.unet_state_dict(FULL, DROPIN["unet_seed"]).items()}
    torch.save(sd, str(tmp_path / "I0_E0_gen.pth"))
    real = load_gae(os.path.join(GOLDEN, "gae_ckpt", "GAE_4_Cav.state.pth"))
    pickle_path = str(tmp_path / "GAE_4_Cav.pth")
    G.save_reference_style_pickle(real, pickle_path)             # classes named __main__.* / common.* like AE.py:637
    opt = config.dict_to_nonedict({**OPT, "path": {**OPT["path"], "resume_state": str(tmp_path / "I0_E0")}})
    set_default_precision(precision)
    orig = (torch.randn, torch.randn_like)
    try:
        diffusion = Model.create_model(opt)
        diffusion.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], schedule_phase="val")     # sr_gae.py:206-207
        # load_network dropped three keys (model.py:189-192): those stay at default init in the reference; copy the
        # values its run ended up with, through .data like a driver would (exercises the in-place edit tracking)
        net = diffusion.netG.denoise_fn
        net.downs["0"].weight.data.copy_(torch.from_numpy(g["first_w"]))
        net.final_conv["block"]["3"].weight.data.copy_(torch.from_numpy(g["last_w"]))
        net.final_conv["block"]["3"].bias.data.copy_(torch.from_numpy(g["last_b"]))
        diffusion.netG.rng_mode = "torch"       # draw through torch.randn / randn_like in the reference's order
        T, hw = diffusion.netG.num_timesteps, DROPIN["hw"]
        assert T == int(g["T"])
        device = torch.device("cuda:0")
        # ---- sr_gae.py:444-467, verbatim apart from the checkpoint path and the noise injection ----------------------------
        model_GAE = torch.load(pickle_path, map_location="cuda:0", weights_only=False)
        model_GAE = model_GAE.to(device)
        val_data = {"HR": synth.sr_cube(1, 31, hw, seed=DROPIN["hr_seed"]), "SR": synth.sr_cube(1, 31, hw, seed=DROPIN["sr_seed"])}
        x_T, tape = synth.noise_tape(5, T, 3, hw, hw, seed=DROPIN["noise_seed"])
        x_T, tape = x_T.cuda(), tape.cuda()
        row_data = val_data["SR"]
        row_data = row_data.to(device)
        val_data["SR"] = val_data["SR"].to(device)
        zSRval_list = model_GAE.encode(val_data["SR"])
        new_list = []
        for i in range(len(zSRval_list)):
            val_data["SR"] = zSRval_list[i]
            diffusion.feed_data(val_data)
            _patched_randn([x_T[i:i + 1]] + [tape[i:i + 1, j] for j in range(T - 1)])
            try:
                diffusion.test(continous=False)
            finally:
                torch.randn, torch.randn_like = orig
            visuals = diffusion.get_current_visuals()
            visuals["SR"] = visuals["SR"].to(device)
            visuals["SR"] = torch.unsqueeze(visuals["SR"], 0)
            new_list.append(visuals["SR"])
        visuals["SR"] = model_GAE.decode(row_data, new_list)
        visuals["SR"][-1][visuals["SR"][-1] < 0] = 0
        visuals["SR"][-1][visuals["SR"][-1] > 1] = 1.
    finally:
        torch.randn, torch.randn_like = orig
        set_default_precision("bf16")
    assert isinstance(model_GAE, G.GAE) and len(zSRval_list) == 5
    lat = torch.cat(new_list).cpu()
    e_lat, e_cube = rel_l2(lat, torch.from_numpy(g["latents"])), rel_l2(visuals["SR"], torch.from_numpy(g["cube"]))
    print(f"drop-in val loop {precision}: latents rel-L2 {e_lat:.3e}, cube rel-L2 {e_cube:.3e}")
    assert visuals["SR"].shape == (1, 31, hw, hw) and visuals["HR"].shape == (1, 31, hw, hw)
    assert torch.equal(visuals["INF"], torch.from_numpy(g["inf"])) or rel_l2(visuals["INF"], torch.from_numpy(g["inf"])) < tol
    assert e_lat < tol and e_cube < tol


def test_default_rng_path_through_the_same_api(reference_names, tmp_path):
    """Same objects with the default generator (CUDA-graph loop + Philox): shapes, ranges and the reference's 3-D return."""
    from hsi_dmgasr_b200 import config, load_gae
    import model as Model
    opt = config.dict_to_nonedict(OPT)
    diffusion = Model.create_model(opt)
    diffusion.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], schedule_phase="val")
    gae = load_gae(os.path.join(GOLDEN, "gae_ckpt", "GAE_4_Cav.state.pth")).cuda()
    sr = synth.sr_cube(1, 31, 32, seed=1).cuda()
    zs = gae.encode(sr)
    outs = []
    for z in zs:
        diffusion.feed_data({"HR": sr, "SR": z})
        diffusion.test(continous=False)
        v = diffusion.get_current_visuals()
        assert v["SR"].shape == (3, 32, 32) and not v["SR"].is_cuda and torch.isfinite(v["SR"]).all()
        outs.append(v["SR"].unsqueeze(0).cuda())
    y = gae.decode(sr, outs)
    assert y.shape == sr.shape and torch.isfinite(y).all()
    diffusion.test(continous=True)
    assert diffusion.SR.shape[0] == 1 + (20 - 1) // (1 | 2) + 1     # condition + snapshots (diffusion.py:193-197)
