"""Generate tests/golden/train_step.npz by running the UNMODIFIED reference training step on CPU
(GaussianDiffusion.p_losses, diffusion.py:222-250, + DDPM.optimize_parameters' normalisation, model.py:49-59).

TEST INFRASTRUCTURE for SURVEY 8f row N2 (the next scope row): pins the loss and the gradients of one training step so
that the oracle restatement (oracle.hsidm_oracle.train_step) - and later a CUDA backward - have a reference to meet.
The reference draws t and the per-sample noise level from numpy's global RNG; both are injected here.  Dropout is 0 in
this configuration so that train() and eval() agree (the reference trains with dropout 0.2; its mask stream cannot be
reproduced outside torch's generator and is not part of the arithmetic being pinned).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from hsi_dmgasr_b200 import synth  # noqa: E402
from hsi_dmgasr_b200.spec import UNetConfig  # noqa: E402
from make_golden import OUT, import_reference, rand, ref_unet  # noqa: E402

TRAIN = UNetConfig(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2), attn_res=(8,),
                   res_blocks=1, dropout=0.0, image_size=16)
SEED, T, B, HW, STEP_T = 51, 20, 3, 16, 7
LEVEL_U = [0.25, 0.5, 0.9]          # position of each sample's noise level inside [lvl[t-1], lvl[t]] (np.random.uniform)


def main():
    _, _, unet_mod, diff_mod = import_reference()
    out = {}
    for loss_type in ("l1", "l2"):
        net = ref_unet(unet_mod, TRAIN, SEED).train()
        gd = diff_mod.GaussianDiffusion(net, image_size=HW, channels=3, loss_type=loss_type, conditional=True)
        gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), "cpu")
        gd.set_loss("cpu")
        hr, sr, noise = rand((B, 3, HW, HW), 61), rand((B, 3, HW, HW), 62), rand((B, 3, HW, HW), 63)
        lo, hi = gd.sqrt_alphas_cumprod_prev[STEP_T - 1], gd.sqrt_alphas_cumprod_prev[STEP_T]
        levels = np.asarray([lo + u * (hi - lo) for u in LEVEL_U], dtype=np.float64)
        orig_randint, orig_uniform = np.random.randint, np.random.uniform
        np.random.randint = lambda a, b=None, *k, **kw: STEP_T
        np.random.uniform = lambda a, b, size=None: levels.copy()
        try:
            with torch.enable_grad():
                l_pix = gd({"HR": hr, "SR": sr}, noise=noise)
                loss = l_pix.sum() / int(B * 3 * HW * HW)            # model.py:53-55
                loss.backward()
        finally:
            np.random.randint, np.random.uniform = orig_randint, orig_uniform
        out[f"{loss_type}.loss_sum"] = np.float64(l_pix.item())
        out[f"{loss_type}.loss"] = np.float64(loss.item())
        out[f"{loss_type}.levels"] = levels
        for k, p in net.named_parameters():
            g = p.grad
            out[f"{loss_type}.gnorm.{k}"] = np.float64(g.double().norm().item())
            if g.numel() <= 4096:                                     # small tensors in full, the rest by norm
                out[f"{loss_type}.grad.{k}"] = g.numpy().copy()
        print(loss_type, "loss", float(loss.detach()), "params with grads", sum(1 for _ in net.parameters()))
    out["t"] = np.int64(STEP_T)
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **out)


if __name__ == "__main__":
    main()
