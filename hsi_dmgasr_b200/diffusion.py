"""Drop-in for ``model/sr3_modules/diffusion.py::GaussianDiffusion`` (sampling side).

``super_resolution`` -> ``p_sample_loop`` runs the whole T-step conditional loop inside the native library
(``hsidm_sample``: one CUDA graph of UNet forward + fused posterior step, replayed T times), batched over all
latent images it is given.  ``p_sample`` / ``p_mean_variance`` keep the reference's step-wise semantics through
``hsidm_unet_forward`` + ``hsidm_posterior_step`` and are what the per-step parity tests drive.
"""
from __future__ import annotations

import hashlib

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .schedule import BUFFER_NAMES, diffusion_buffers, make_beta_schedule


class GaussianDiffusion(nn.Module):
    def __init__(self, denoise_fn, image_size, channels=31, loss_type="l1", conditional=True, schedule_opt=None):
        super().__init__()
        self.channels = channels
        self.image_size = image_size
        self.denoise_fn = denoise_fn
        self.loss_type = loss_type
        self.conditional = conditional
        self.num_timesteps = 0
        self._betas64: Optional[np.ndarray] = None
        # Where the random draws of p_sample_loop come from when the caller injects none:
        #   "philox" (default): x_T from torch.randn, per-step noise from the library's counter-based generator inside the
        #            CUDA-graph loop (hsidm_sample);
        #   "torch":  exactly the reference's call sequence - torch.randn(shape) once, then torch.randn_like(x) for
        #            t = T-1 .. 1 (diffusion.py:174, 192) - through the step-wise path, so a driver that seeds or patches
        #            torch's generators sees the same stream of draws as with the reference.
        self.rng_mode = os.environ.get("HSIDM_RNG", "philox")
        # like the reference (diffusion.py:82-84) the schedule is NOT installed here; DDPM.__init__ does it.

    # ---- schedule -------------------------------------------------------------------------------------------
    def set_loss(self, device):
        if self.loss_type == "l1":
            self.loss_func = nn.L1Loss(reduction="sum").to(device)
        elif self.loss_type == "l2":
            self.loss_func = nn.MSELoss(reduction="sum").to(device)
        else:
            raise NotImplementedError()

    def set_new_noise_schedule(self, schedule_opt, device):
        """diffusion.py:93-140: float64 tables -> 12 fp32 buffers + float64 ``sqrt_alphas_cumprod_prev``."""
        betas = make_beta_schedule(schedule=schedule_opt["schedule"], n_timestep=schedule_opt["n_timestep"],
                                   linear_start=schedule_opt["linear_start"], linear_end=schedule_opt["linear_end"])
        tabs = diffusion_buffers(betas)
        self.sqrt_alphas_cumprod_prev = tabs["sqrt_alphas_cumprod_prev"]
        self.num_timesteps = int(betas.shape[0])
        self._betas64 = np.ascontiguousarray(betas, dtype=np.float64)
        for name in BUFFER_NAMES:
            self.register_buffer(name, torch.tensor(tabs[name], dtype=torch.float32, device=device))

    def _native(self, device):
        h = self.denoise_fn.native(device)
        if self._betas64 is None:
            raise _lib.HsidmError(-7, "set_new_noise_schedule has not been called")
        # content, not address: setting the same schedule again (the reference's drivers do, per cube) must not cost a
        # table rebuild and a graph re-capture
        sig = (self.num_timesteps, hashlib.sha1(self._betas64.tobytes()).hexdigest())
        if h.schedule_sig != sig:
            lib = _lib.load()
            _lib.check(lib.hsidm_set_schedule(h.ptr, self._betas64.ctypes.data_as(C.POINTER(C.c_double)),
                                              self.num_timesteps))
            h.schedule_sig = sig
        return h

    # ---- reference helper formulas (diffusion.py:142-150), kept for API parity -------------------------------------
    def predict_start_from_noise(self, x_t, t, noise):
        return self.sqrt_recip_alphas_cumprod[t] * x_t - self.sqrt_recipm1_alphas_cumprod[t] * noise

    def q_posterior(self, x_start, x_t, t):
        mean = self.posterior_mean_coef1[t] * x_start + self.posterior_mean_coef2[t] * x_t
        return mean, self.posterior_log_variance_clipped[t]

    # ---- step-wise path -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def predict_noise(self, x, t, condition_x=None):
        """eps = denoise_fn(cat([condition_x, x]), noise_level_t) (diffusion.py:154-161) without materialising the cat."""
        x = _lib.require_cuda_f32(x, "x")
        h = self._native(x.device)
        n, c, hh, ww = x.shape
        level = torch.full((1,), float(self.sqrt_alphas_cumprod_prev[t + 1]), dtype=torch.float32, device=x.device)
        eps = torch.empty((n, self.denoise_fn.cfg.out_channel, hh, ww), device=x.device, dtype=torch.float32)
        lib = _lib.load()
        st = _lib.stream_ptr(x.device)
        if condition_x is not None:
            cond = _lib.require_cuda_f32(condition_x, "condition_x")
            _lib.check(lib.hsidm_unet_forward(h.ptr, cond.data_ptr(), cond.shape[1], x.data_ptr(), c, level.data_ptr(), 0,
                                              eps.data_ptr(), n, hh, ww, st))
        else:
            _lib.check(lib.hsidm_unet_forward(h.ptr, x.data_ptr(), c, None, 0, level.data_ptr(), 0, eps.data_ptr(), n, hh,
                                              ww, st))
        return eps

    @torch.no_grad()
    def p_sample(self, x, t, clip_denoised=True, condition_x=None, noise=None):
        """One reverse step (diffusion.py:170-175). ``noise`` defaults to ``torch.randn_like(x)`` for t > 0."""
        if not clip_denoised:
            raise NotImplementedError("clip_denoised=False is never used by the reference drivers")
        eps = self.predict_noise(x, t, condition_x)
        h = self._native(x.device)
        if t > 0 and noise is None:
            noise = torch.randn_like(x)
        if t == 0:
            noise = None
        if noise is not None:
            noise = _lib.require_cuda_f32(noise, "noise")
            if noise.shape != x.shape:
                raise _lib.HsidmError(-1, f"noise shape {tuple(noise.shape)} != x shape {tuple(x.shape)}")
        x = x.contiguous()
        out = torch.empty_like(x)
        _lib.check(_lib.load().hsidm_posterior_step(h.ptr, int(t), x.data_ptr(), eps.data_ptr(), _lib.ptr(noise),
                                                    out.data_ptr(), x.numel(), _lib.stream_ptr(x.device)))
        return out

    # ---- whole loop -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def p_sample_loop(self, x_in, continous=False, *, x_T=None, noise_tape=None, return_all=False, seed=None, first_image=0):
        """diffusion.py:177-201.

        Conditional: ``x_in`` is the condition [N,c,H,W]; all N images are sampled as one batch.
        ``x_T`` [N,c,H,W] / ``noise_tape`` [N,T-1,c,H,W] inject the random draws (parity tests); otherwise x_T comes
        from ``torch.randn`` and the per-step noise from the library's counter-based generator seeded from torch's
        default generator.  Return shapes follow the reference: ``continous`` -> cat([x_in, snapshots...], 0);
        else the LAST batch element of the final image as a 3-D tensor (diffusion.py:198-201), unless
        ``return_all`` asks for the whole batch [N,c,H,W].

        ``seed`` (explicit) makes every draw - x_T included - come from the library's counter-based generator, keyed
        by (seed, position of the image in the caller's list): with ``first_image`` = index of this batch's first image
        in that list, an image gets the same noise whatever batch or GPU it is sampled in.
        """
        device = self.betas.device
        if not self.conditional:
            shape = x_in
            img = torch.randn(shape, device=device) if x_T is None else x_T
            ret = img
            inter = 1 | (self.num_timesteps // 10)
            for i in reversed(range(self.num_timesteps)):
                img = self.p_sample(img, i)
                if i % inter == 0:
                    ret = torch.cat([ret, img], dim=0)
            return ret if continous else (img if return_all else ret[-1])
        cond = _lib.require_cuda_f32(x_in, "x_in")
        n, c, hh, ww = cond.shape
        if self.rng_mode == "torch" and x_T is None and noise_tape is None:
            # the reference's own loop, draw for draw (diffusion.py:188-201)
            img = torch.randn(cond.shape, device=device)
            ret_img = cond
            inter = 1 | (self.num_timesteps // 10)
            for i in reversed(range(self.num_timesteps)):
                img = self.p_sample(img, i, condition_x=cond)
                if i % inter == 0:
                    ret_img = torch.cat([ret_img, img], dim=0)
            if continous:
                return ret_img
            return img if return_all else ret_img[-1]
        if self.rng_mode not in ("philox", "torch"):
            raise ValueError(f"rng_mode must be 'philox' or 'torch', not {self.rng_mode!r}")
        h = self._native(cond.device)
        lib = _lib.load()
        if x_T is None and seed is not None:
            x_T = torch.empty_like(cond)
            _lib.check(lib.hsidm_randn(x_T.data_ptr(), x_T.numel(), int(seed), int(first_image) * c * hh * ww,
                                       _lib.stream_ptr(cond.device)))
        x_T = torch.randn(cond.shape, device=cond.device) if x_T is None else _lib.require_cuda_f32(x_T, "x_T")
        if tuple(x_T.shape) != tuple(cond.shape):
            raise _lib.HsidmError(-1, f"x_T shape {tuple(x_T.shape)} != condition shape {tuple(cond.shape)}")
        tape_ptr, s_img, s_step = None, 0, 0
        if noise_tape is not None:
            noise_tape = _lib.require_cuda_f32(noise_tape, "noise_tape")
            want = (n, max(self.num_timesteps - 1, 0), c, hh, ww)
            if tuple(noise_tape.shape) != want:
                raise _lib.HsidmError(-1, f"noise_tape shape {tuple(noise_tape.shape)} != {want}")
            tape_ptr, s_img, s_step = noise_tape.data_ptr(), noise_tape.stride(0), noise_tape.stride(1)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        out = torch.empty_like(cond)
        snaps = None
        if continous:
            n_snap = lib.hsidm_snapshot_count(h.ptr)
            snaps = torch.empty((n_snap,) + tuple(cond.shape), device=cond.device, dtype=torch.float32)
        _lib.check(lib.hsidm_sample_at(h.ptr, cond.data_ptr(), x_T.data_ptr(), tape_ptr, s_img, s_step, seed, int(first_image),
                                       out.data_ptr(), _lib.ptr(snaps), n, hh, ww, _lib.stream_ptr(cond.device)))
        if continous:
            return torch.cat([cond, snaps.reshape(-1, c, hh, ww)], dim=0)
        return out if return_all else out[-1]

    @torch.no_grad()
    def sample(self, batch_size=1, continous=False):
        return self.p_sample_loop((batch_size, self.channels, self.image_size, self.image_size), continous)

    @torch.no_grad()
    def super_resolution(self, x_in, continous=False, **kw):
        return self.p_sample_loop(x_in, continous, **kw)

    # ---- training side -------------------------------------------------------------------------------
    def q_sample(self, x_start, continuous_sqrt_alpha_cumprod, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return continuous_sqrt_alpha_cumprod * x_start + (1 - continuous_sqrt_alpha_cumprod ** 2).sqrt() * noise

    def p_losses(self, x_in, noise=None):
        """diffusion.py:222-250: one shared integer t from ``np.random.randint``, per-sample continuous noise levels from
        ``np.random.uniform`` between the neighbouring schedule levels, q_sample, UNet on cat([SR, x_noisy]) and the summed
        L1 / L2 loss.  The forward and the hand-written backward run in the native library (hsidm_train_forward /
        hsidm_train_backward, fp32); the returned scalar carries an autograd node, so ``loss.backward()`` followed by a stock
        ``torch.optim`` step works exactly as in ``DDPM.optimize_parameters`` (model.py:49-59)."""
        if not self.conditional:
            raise NotImplementedError("unconditional training is not used by HSI-DMGASR (every config sets conditional: true)")
        x_start = _lib.require_cuda_f32(x_in["HR"], "x_in['HR']")
        cond = _lib.require_cuda_f32(x_in["SR"], "x_in['SR']")
        b = x_start.shape[0]
        t = np.random.randint(1, self.num_timesteps + 1)
        levels = np.random.uniform(self.sqrt_alphas_cumprod_prev[t - 1], self.sqrt_alphas_cumprod_prev[t], size=b)
        levels = np.ascontiguousarray(np.asarray(levels, dtype=np.float32).reshape(-1))      # torch.FloatTensor(...) in the reference
        noise = torch.randn_like(x_start) if noise is None else _lib.require_cuda_f32(noise, "noise")
        if self.loss_type not in ("l1", "l2"):
            raise NotImplementedError()
        h = self._native(x_start.device)
        named = dict(self.denoise_fn.named_parameters())
        params = [named[k] for k in h.keys()]
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if self.denoise_fn.cfg.dropout > 0 else 0
        return _PLosses.apply(self, h, x_start, cond, noise, levels, seed, *params)

    def forward(self, x, *args, **kwargs):
        return self.p_losses(x, *args, **kwargs)


class _PLosses(torch.autograd.Function):
    """Autograd node around the native training step: forward = hsidm_train_forward (saves activations inside the library),
    backward = hsidm_train_backward; the parameter gradients are views of ONE contiguous slab (``gd.last_grad_slab``), which is
    what a data-parallel job all-reduces."""

    @staticmethod
    def forward(ctx, gd, h, hr, sr, noise, levels, seed, *params):
        lib = _lib.load()
        n, c, hh, ww = hr.shape
        if tuple(sr.shape) != tuple(hr.shape) or tuple(noise.shape) != tuple(hr.shape):
            raise _lib.HsidmError(-1, f"HR {tuple(hr.shape)}, SR {tuple(sr.shape)} and noise {tuple(noise.shape)} must agree")
        slab = torch.empty(lib.hsidm_train_grad_numel(h.ptr), device=hr.device, dtype=torch.float32)
        loss = torch.empty(1, device=hr.device, dtype=torch.float32)
        _lib.check(lib.hsidm_train_forward(h.ptr, hr.data_ptr(), sr.data_ptr(), noise.data_ptr(), levels.ctypes.data, n, hh, ww,
                                           0 if gd.loss_type == "l1" else 1, seed, slab.data_ptr(), loss.data_ptr(),
                                           _lib.stream_ptr(hr.device)))
        ctx.h, ctx.slab, ctx.count, ctx.device = h, slab, n * c * hh * ww, hr.device
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.offsets = [lib.hsidm_train_grad_offset(h.ptr, i) for i in range(len(params))]
        gd.last_grad_slab = slab
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        upstream = float(grad_out) * ctx.count          # the library differentiates loss_sum / count
        _lib.check(lib.hsidm_train_backward(ctx.h.ptr, upstream, _lib.stream_ptr(ctx.device)))
        grads = []
        for shape, off in zip(ctx.shapes, ctx.offsets):
            numel = 1
            for d in shape:
                numel *= d
            grads.append(ctx.slab[off:off + numel].view(shape))
        return (None,) * 7 + tuple(grads)


def allreduce_gradients(gd: "GaussianDiffusion", world: int, dist=None) -> None:
    """Data-parallel gradient averaging for one process per GPU (BASELINE configs[4]; the reference wraps netG in
    nn.DataParallel instead, networks.py:113-115): ONE all-reduce of the contiguous gradient slab of the last backward
    (NCCL over NVLink on GPUs), then 1/world.  Every ``p.grad`` is a view of that slab."""
    if world <= 1:
        return
    if dist is None:
        import torch.distributed as dist
    slab = getattr(gd, "last_grad_slab", None)
    if slab is None:
        raise _lib.HsidmError(-7, "no backward pass to all-reduce")
    dist.all_reduce(slab)
    slab.mul_(1.0 / world)
