"""CPU oracle for the HSI-DMGASR inference hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 PyTorch on the CPU and in a *functional* style over flat
``state_dict`` mappings, the algorithm of the reference's hot path:

  * beta schedules + schedule buffers      reference model/sr3_modules/diffusion.py:11-49, 93-140
  * noise-level embedding, UNet forward    reference model/sr3_modules/unet.py:18-31, 34-50, 80-159, 162-263
  * p_mean_variance / p_sample / loop      reference model/sr3_modules/diffusion.py:142-201
  * GAE group layout / encode / decode     reference AE.py:256-324, common.py:163-182, 231-271
  * MPSNR / SAM                            reference eval_hsi.py:47-65, 110-121
  * ERGAS / CC / RMSE / MSSIM              reference eval_hsi.py:18-35, 58-70, 88-96, 124-135 (MSSIM restates the
                                           published algorithm of skimage.metrics.structural_similarity, a dependency
                                           absent from this image: that one index is "parity unpinned")
  * MATLAB-style imresize                  reference GAE/imsize.py:35-59, 116-158

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; it is the checker, never the product.  The product package
(``hsi_dmgasr_b200``) does not import anything from ``oracle/`` and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the *unmodified reference modules* imported in the build container; the generating
script is ``oracle/make_golden.py`` and the vectors live in ``tests/golden/``
(``tests/test_oracle_golden.py`` checks them on every CPU run).  The dense arithmetic itself
(conv / group_norm / softmax / matmul) lives in PyTorch ATen, exactly as it does for the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------
# schedule  (diffusion.py:11-49, 93-140)
# --------------------------------------------------------------------------------------------------
def beta_schedule(schedule: str, n_timestep: int, linear_start: float = 1e-4, linear_end: float = 2e-2,
                  cosine_s: float = 8e-3) -> np.ndarray:
    """float64 betas; one branch per schedule name of diffusion.py:19-49."""
    T = int(n_timestep)
    if schedule == "quad":
        b = np.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=np.float64) ** 2
    elif schedule == "linear":
        b = np.linspace(linear_start, linear_end, T, dtype=np.float64)
    elif schedule in ("warmup10", "warmup50"):
        frac = 0.1 if schedule == "warmup10" else 0.5
        b = linear_end * np.ones(T, dtype=np.float64)
        w = int(T * frac)
        b[:w] = np.linspace(linear_start, linear_end, w, dtype=np.float64)
    elif schedule == "const":
        b = linear_end * np.ones(T, dtype=np.float64)
    elif schedule == "jsd":
        b = 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    elif schedule == "cosine":
        # the reference evaluates this branch with torch float64 ops (diffusion.py:36-45)
        s = torch.arange(T + 1, dtype=torch.float64) / T + cosine_s
        a = torch.cos(s / (1 + cosine_s) * math.pi / 2).pow(2)
        a = a / a[0]
        b = (1 - a[1:] / a[:-1]).clamp(max=0.999).numpy()
    else:
        raise NotImplementedError(schedule)
    return np.asarray(b, dtype=np.float64)


def schedule_tables(betas: np.ndarray) -> Dict[str, np.ndarray]:
    """The 12 fp32 buffers + the float64 noise-level table of diffusion.py:93-140."""
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(post_var),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
        # float64, length T+1; entry t+1 is the noise level fed to the UNet at loop index t
        "sqrt_alphas_cumprod_prev": np.sqrt(np.append(1.0, ac)),
    }


# --------------------------------------------------------------------------------------------------
# UNet  (unet.py)
# --------------------------------------------------------------------------------------------------
def unet_topology(inner_channel: int, channel_mults: Sequence[int], attn_res: Sequence[int], res_blocks: int,
                  image_size: int):
    """Module list implied by UNet.__init__ (unet.py:190-236).

    Returns (downs, mid, ups); each entry is ("conv",) | ("res", cin, cout, attn) | ("down", c) | ("up", c).
    ``attn`` is decided against the *config* image_size, not the runtime size (unet.py:195,200).
    """
    attn_res = list(attn_res) if attn_res is not None else []
    pre = inner_channel
    feat = [pre]
    now = image_size
    downs = [("conv",)]
    L = len(channel_mults)
    for i in range(L):
        cm = inner_channel * channel_mults[i]
        for _ in range(res_blocks):
            downs.append(("res", pre, cm, now in attn_res))
            feat.append(cm)
            pre = cm
        if i != L - 1:
            downs.append(("down", pre))
            feat.append(pre)
            now //= 2
    mid = [("res", pre, pre, True), ("res", pre, pre, False)]
    ups = []
    for i in reversed(range(L)):
        cm = inner_channel * channel_mults[i]
        for _ in range(res_blocks + 1):
            ups.append(("res", pre + feat.pop(), cm, now in attn_res))
            pre = cm
        if i >= 1:
            ups.append(("up", pre))
            now *= 2
    return downs, mid, ups


def noise_embedding(sd: SD, level: torch.Tensor, dim: int) -> torch.Tensor:
    """PositionalEncoding + 2-layer MLP (unet.py:18-31, 182-187). level: [N,1] -> [N,1,dim]."""
    half = dim // 2
    step = torch.arange(half, dtype=level.dtype, device=level.device) / half
    enc = level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
    h = F.linear(enc, sd["noise_level_mlp.1.weight"], sd["noise_level_mlp.1.bias"])
    h = h * torch.sigmoid(h)
    return F.linear(h, sd["noise_level_mlp.3.weight"], sd["noise_level_mlp.3.bias"])


def _gn_swish_conv(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """Block = GroupNorm -> Swish -> (Dropout=identity in eval) -> Conv3x3 (unet.py:80-91)."""
    h = F.group_norm(x, groups, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], eps=1e-5)
    h = h * torch.sigmoid(h)
    return F.conv2d(h, sd[p + ".block.3.weight"], sd[p + ".block.3.bias"], padding=1)


def _res_block(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, groups: int) -> torch.Tensor:
    """ResnetBlock.forward (unet.py:105-111) with use_affine_level=False (unet.py:49)."""
    h = _gn_swish_conv(sd, p + ".block1", x, groups)
    nb = F.linear(temb, sd[p + ".noise_func.noise_func.0.weight"], sd[p + ".noise_func.noise_func.0.bias"])
    h = h + nb.view(x.shape[0], -1, 1, 1)
    h = _gn_swish_conv(sd, p + ".block2", h, groups)
    if (p + ".res_conv.weight") in sd:
        x = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    return h + x


def _self_attention(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """SelfAttention.forward with n_head=1 (unet.py:124-143)."""
    n, c, hh, ww = x.shape
    nrm = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    qkv = F.conv2d(nrm, sd[p + ".qkv.weight"], None)
    q, k, v = qkv.view(n, 3, c, hh * ww).unbind(1)             # channel chunks [0:C],[C:2C],[2C:3C]
    att = torch.softmax(torch.einsum("ncs,nct->nst", q, k) / math.sqrt(c), dim=-1)
    o = torch.einsum("nst,nct->ncs", att, v).reshape(n, c, hh, ww)
    o = F.conv2d(o, sd[p + ".out.weight"], sd[p + ".out.bias"])
    return o + x


def unet_forward(sd: SD, cfg: dict, x: torch.Tensor, level: torch.Tensor,
                 taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """UNet.forward (unet.py:239-263). ``sd`` keys are relative to ``denoise_fn.``; x [N,Cin,H,W], level [N,1]."""
    groups = cfg.get("norm_groups") or 32
    downs, mid, ups = unet_topology(cfg["inner_channel"], cfg["channel_multiplier"], cfg["attn_res"],
                                    cfg["res_blocks"], cfg["image_size"])
    temb = noise_embedding(sd, level, cfg["inner_channel"])

    def run_res(prefix, spec, h):
        h = _res_block(sd, prefix + ".res_block", h, temb, groups)
        if spec[3]:
            h = _self_attention(sd, prefix + ".attn", h, groups)
        return h

    feats: List[torch.Tensor] = []
    h = x
    for i, spec in enumerate(downs):
        if spec[0] == "conv":
            h = F.conv2d(h, sd["downs.0.weight"], sd["downs.0.bias"], padding=1)
        elif spec[0] == "res":
            h = run_res(f"downs.{i}", spec, h)
        else:
            h = F.conv2d(h, sd[f"downs.{i}.conv.weight"], sd[f"downs.{i}.conv.bias"], stride=2, padding=1)
        feats.append(h)
        if taps is not None:
            taps[f"downs.{i}"] = h
    for i, spec in enumerate(mid):
        h = run_res(f"mid.{i}", spec, h)
        if taps is not None:
            taps[f"mid.{i}"] = h
    for i, spec in enumerate(ups):
        if spec[0] == "res":
            h = run_res(f"ups.{i}", spec, torch.cat((h, feats.pop()), dim=1))
        else:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[f"ups.{i}.conv.weight"], sd[f"ups.{i}.conv.bias"], padding=1)
        if taps is not None:
            taps[f"ups.{i}"] = h
    return _gn_swish_conv(sd, "final_conv", h, groups)


# --------------------------------------------------------------------------------------------------
# sampling  (diffusion.py:142-201)
# --------------------------------------------------------------------------------------------------
def posterior_step(tab: Dict[str, np.ndarray], t: int, x_t: torch.Tensor, eps: torch.Tensor,
                   noise: Optional[torch.Tensor]) -> torch.Tensor:
    """predict_start_from_noise -> clamp -> q_posterior -> add noise (diffusion.py:142-175)."""
    f = lambda k: torch.tensor(tab[k][t], dtype=torch.float32)
    x0 = f("sqrt_recip_alphas_cumprod") * x_t - f("sqrt_recipm1_alphas_cumprod") * eps
    x0 = x0.clamp(-1.0, 1.0)
    mean = f("posterior_mean_coef1") * x0 + f("posterior_mean_coef2") * x_t
    if noise is None:
        noise = torch.zeros_like(x_t)
    return mean + noise * (0.5 * f("posterior_log_variance_clipped")).exp()


def sample_loop(sd: SD, cfg: dict, tab: Dict[str, np.ndarray], cond: torch.Tensor, x_T: torch.Tensor,
                step_noise, continous: bool = False, record=None) -> torch.Tensor:
    """Conditional branch of p_sample_loop (diffusion.py:188-201).

    ``step_noise(i)`` returns the N(0,1) tensor used at loop index i (i = T-1 .. 1); index 0 uses zeros.
    Returns what the reference returns: ``ret_img`` (continous) or ``ret_img[-1]`` (a 3-D tensor).
    """
    T = len(tab["betas"])
    inter = 1 | (T // 10)
    img = x_T
    ret = cond
    n = cond.shape[0]
    for i in reversed(range(T)):
        level = torch.full((n, 1), float(tab["sqrt_alphas_cumprod_prev"][i + 1]), dtype=torch.float32)
        eps = unet_forward(sd, cfg, torch.cat([cond, img], dim=1), level)
        img = posterior_step(tab, i, img, eps, step_noise(i) if i > 0 else None)
        if record is not None:
            record(i, eps, img)
        if i % inter == 0:
            ret = torch.cat([ret, img], dim=0)
    return ret if continous else ret[-1]


# --------------------------------------------------------------------------------------------------
# GAE  (AE.py:145-324, common.py:163-182, 231-271)
# --------------------------------------------------------------------------------------------------
def gae_groups(n_colors: int, n_subs: int, n_ovls: int):
    """Group start/end indices (AE.py:264-280)."""
    G = math.ceil((n_colors - n_ovls) / (n_subs - n_ovls))
    start, end = [], []
    for g in range(G):
        s = (n_subs - n_ovls) * g
        e = s + n_subs
        if e > n_colors:
            e = n_colors
            s = n_colors - n_subs
        start.append(s)
        end.append(e)
    return G, start, end


def _lrelu(x):
    return F.leaky_relu(x, 0.01)


def _branch(sd: SD, p: str, x: torch.Tensor, n_blocks: int, res_scale: float = 0.1) -> torch.Tensor:
    """BranchUnit(use_tail=False, up_scale=1) = head conv -> SSPN(n_blocks x SSB) + skip (AE.py:120-165)."""
    y = F.conv2d(x, sd[p + ".head.weight"], sd[p + ".head.bias"], padding=1)
    h = y
    for b in range(n_blocks):
        q = f"{p}.body.net.{b}"
        # ResBlock k=3 (common.py:163-182): x + s*conv(act(conv(x)))
        r = F.conv2d(h, sd[q + ".spa.body.0.weight"], sd[q + ".spa.body.0.bias"], padding=1)
        r = F.conv2d(_lrelu(r), sd[q + ".spa.body.2.weight"], sd[q + ".spa.body.2.bias"], padding=1)
        h = r * res_scale + h
        # ResAttentionBlock k=1 + CALayer(reduction 3) (common.py:231-271)
        r = F.conv2d(h, sd[q + ".spc.body.0.weight"], sd[q + ".spc.body.0.bias"])
        r = F.conv2d(_lrelu(r), sd[q + ".spc.body.2.weight"], sd[q + ".spc.body.2.bias"])
        w = r.mean(dim=(2, 3), keepdim=True)
        w = F.relu(F.conv2d(w, sd[q + ".spc.body.3.conv_du.0.weight"], sd[q + ".spc.body.3.conv_du.0.bias"]))
        w = torch.sigmoid(F.conv2d(w, sd[q + ".spc.body.3.conv_du.2.weight"], sd[q + ".spc.body.3.conv_du.2.bias"]))
        h = (r * w) * res_scale + h
    return h + y


def _count_blocks(sd: SD, p: str) -> int:
    n = 0
    while f"{p}.body.net.{n}.spa.body.0.weight" in sd:
        n += 1
    return n


def gae_coder(sd: SD, which: str, x: torch.Tensor) -> torch.Tensor:
    """Encoder.forward / Decoder.forward (AE.py:195-199, 236-242): final(branch(x))."""
    h = _branch(sd, which + ".branch", x, _count_blocks(sd, which + ".branch"))
    return F.conv2d(h, sd[which + ".final.weight"], sd[which + ".final.bias"], padding=1)


def gae_encode(sd: SD, geom: dict, x: torch.Tensor) -> List[torch.Tensor]:
    """GAE.encode (AE.py:310-324) without the hard-coded 'cuda:0'."""
    _, start, end = gae_groups(geom["n_colors"], geom["n_subs"], geom["n_ovls"])
    return [gae_coder(sd, "Encoder", x[:, s:e]) for s, e in zip(start, end)]


def gae_decode(sd: SD, geom: dict, x: torch.Tensor, z_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """GAE.decode (AE.py:283-308): overlap-accumulate, divide by band count, residual trunk."""
    b, c, h, w = x.shape
    _, start, end = gae_groups(geom["n_colors"], geom["n_subs"], geom["n_ovls"])
    y = torch.zeros(b, c, h, w)
    cnt = torch.zeros(c)
    for g, (s, e) in enumerate(zip(start, end)):
        y[:, s:e] += gae_coder(sd, "Decoder", z_list[g])
        cnt[s:e] = cnt[s:e] + 1
    y = y / cnt.unsqueeze(1).unsqueeze(2)
    t = _branch(sd, "trunk", y, _count_blocks(sd, "trunk"))
    t = F.conv2d(t, sd["final.weight"], sd["final.bias"], padding=1)
    return t + y


# --------------------------------------------------------------------------------------------------
# val driver (sr_gae.py:456-475) and metrics (eval_hsi.py:47-65, 110-121)
# --------------------------------------------------------------------------------------------------
def sr_cube(unet_sd: SD, cfg: dict, tab, gae_sd: SD, geom: dict, sr: torch.Tensor, x_T: Sequence[torch.Tensor],
            step_noise) -> torch.Tensor:
    """One cube through encode -> per-group sampling (groups sequential, batch 1) -> decode -> clamp[0,1].

    ``x_T[g]`` / ``step_noise(g, i)`` is the group-major noise tape of SURVEY.md 8d.
    """
    assert sr.shape[0] == 1
    zs = gae_encode(gae_sd, geom, sr)
    outs = []
    for g, z in enumerate(zs):
        r = sample_loop(unet_sd, cfg, tab, z, x_T[g], lambda i, g=g: step_noise(g, i), continous=False)
        outs.append(r.unsqueeze(0))
    return gae_decode(gae_sd, geom, sr, outs).clamp(0.0, 1.0)


def mpsnr(x_true: np.ndarray, x_pred: np.ndarray, data_range: float = 1.0) -> float:
    """compare_mpsnr over HWC arrays; skimage's PSNR is 10*log10(R^2/MSE) with float64 MSE."""
    a = x_true.astype(np.float32).astype(np.float64)
    b = x_pred.astype(np.float32).astype(np.float64)
    mse = ((a - b) ** 2).mean(axis=(0, 1))
    return float(np.mean(10.0 * np.log10(data_range ** 2 / mse)))


def sam_deg(x_true: np.ndarray, x_pred: np.ndarray) -> float:
    """compare_sam over HWC arrays: mean spectral angle (degrees) over pixels with non-zero norms."""
    a = x_true.astype(np.float32).reshape(-1, x_true.shape[2])
    b = x_pred.astype(np.float32).reshape(-1, x_pred.shape[2])
    na = np.linalg.norm(a, axis=1)
    nb = np.linalg.norm(b, axis=1)
    ok = (na != 0) & (nb != 0)
    cos = (a[ok] * b[ok]).sum(axis=1) / (na[ok] * nb[ok])
    # the reference lets arccos produce nan for |cos|>1 by rounding; clip only by 1 ulp to stay defined
    ang = np.arccos(np.clip(cos, -1.0, 1.0))
    return float(ang.sum() / ok.sum() * 180.0 / np.pi)


# ----------------------------------------------------------------------------------------------------------------------
# the steps either side of the path
def bicubic_pre_upsample(lr: torch.Tensor, scale: int = 4) -> torch.Tensor:
    """The dataset code's pre-upsampling, verbatim call (sr_gae.py:72, :118): torch bicubic, align_corners=False."""
    return torch.nn.functional.interpolate(lr, scale_factor=scale, mode="bicubic")


def imresize_matlab(img: np.ndarray, output_shape=None, method: str = "bicubic", scalar_scale=None) -> np.ndarray:
    """imsize.imresize(I, output_shape=... | scalar_scale=...) (imsize.py:116-158) for an HWC or HW array, float64 result.
    With scalar_scale the scale is the given number and the size ceil(scale * n) (imsize.py:3-7, 131-134); with
    output_shape the scale is out / in per axis (imsize.py:10-14, 135-137).  Per axis
    (`contributions`, imsize.py:35-59): kernel stretched by 1/scale when shrinking, P = ceil(width) + 2 taps from
    floor(u - width/2) - 1, weights normalised, taps mirrored with period 2*length; the axis with the smaller scale goes
    first (imsize.py:141-152)."""
    def cubic(x):
        ax = np.abs(x)
        return (1.5 * ax ** 3 - 2.5 * ax ** 2 + 1) * (ax <= 1) + (-0.5 * ax ** 3 + 2.5 * ax ** 2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2))

    def triangle(x):
        return (x + 1) * ((x >= -1) & (x < 0)) + (1 - x) * ((x <= 1) & (x >= 0))

    kern = {"bicubic": cubic, "bilinear": triangle}[method]
    a = np.asarray(img)
    b = a[..., None] if a.ndim == 2 else a
    if scalar_scale is not None:
        scales = [float(scalar_scale)] * 2
        output_shape = [int(math.ceil(sc * n)) for sc, n in zip(scales, b.shape[:2])]
    else:
        scales = [1.0 * out / n for out, n in zip(output_shape, b.shape[:2])]
    for dim in np.argsort(np.array(scales)):
        scale, n_in, n_out = scales[dim], b.shape[dim], int(output_shape[dim])
        width = 4.0 / scale if scale < 1 else 4.0
        u = np.arange(1, n_out + 1, dtype=np.float64) / scale + 0.5 * (1 - 1 / scale)
        left = np.floor(u - width / 2)
        taps = (left[:, None] + np.arange(int(math.ceil(width)) + 2) - 1).astype(np.int64)
        arg = u[:, None] - taps - 1
        wts = scale * kern(scale * arg) if scale < 1 else kern(arg)
        wts = wts / wts.sum(axis=1, keepdims=True)
        mirror = np.concatenate([np.arange(n_in), np.arange(n_in - 1, -1, -1)])
        taps = mirror[np.mod(taps, mirror.size)]
        moved = np.moveaxis(b.astype(np.float64), dim, 0)                  # [n_in, ...]
        b = np.moveaxis(np.einsum("op,op...->o...", wts, moved[taps]), 0, dim)
    return b[..., 0] if a.ndim == 2 else b


def ergas(x_true: np.ndarray, x_pred: np.ndarray, ratio: float) -> float:
    """compare_ergas (eval_hsi.py:18-35) over HWC arrays: (100/ratio) sqrt(mean_bands(MSE_b / mean(true_b)^2))."""
    a = x_true.astype(np.float32).astype(np.float64).reshape(-1, x_true.shape[2])
    b = x_pred.astype(np.float32).astype(np.float64).reshape(-1, x_pred.shape[2])
    return float(100.0 / ratio * np.sqrt(np.mean(((a - b) ** 2).mean(axis=0) / a.mean(axis=0) ** 2)))


def cross_correlation(x_true: np.ndarray, x_pred: np.ndarray) -> float:
    """compare_corr (eval_hsi.py:58-70): Pearson correlation per band, mean over bands."""
    a = x_true.astype(np.float32).astype(np.float64).reshape(-1, x_true.shape[2])
    b = x_pred.astype(np.float32).astype(np.float64).reshape(-1, x_pred.shape[2])
    a = a - a.mean(axis=0)
    b = b - b.mean(axis=0)
    return float(((a * b).sum(axis=0) / np.sqrt((a * a).sum(axis=0) * (b * b).sum(axis=0))).mean())


def rmse(x_true: np.ndarray, x_pred: np.ndarray) -> float:
    """compare_rmse (eval_hsi.py:88-96): ||true - pred||_F / sqrt(H W C)."""
    d = x_true.astype(np.float32).astype(np.float64) - x_pred.astype(np.float32).astype(np.float64)
    return float(np.sqrt((d ** 2).sum() / d.size))


def mssim(x_true: np.ndarray, x_pred: np.ndarray, data_range: float = 1.0) -> float:
    """compare_mssim (eval_hsi.py:124-135): mean over bands of skimage.metrics.structural_similarity(im1, im2,
    data_range=...) with skimage's defaults for float images - win_size 7, uniform filter, K1 = 0.01, K2 = 0.03,
    use_sample_covariance=True, mean of S over the image cropped by (win_size-1)//2 (where the window is inside the
    image, so the filter's border mode does not matter).  skimage is absent here: restated from its published algorithm,
    PARITY UNPINNED for this one index."""
    from scipy.ndimage import uniform_filter
    win, c1, c2 = 7, (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    cov = win * win / (win * win - 1.0)
    pad = (win - 1) // 2
    vals = []
    for k in range(x_true.shape[2]):
        x = x_true[:, :, k].astype(np.float32).astype(np.float64)
        y = x_pred[:, :, k].astype(np.float32).astype(np.float64)
        ux, uy = uniform_filter(x, win), uniform_filter(y, win)
        vx = cov * (uniform_filter(x * x, win) - ux * ux)
        vy = cov * (uniform_filter(y * y, win) - uy * uy)
        vxy = cov * (uniform_filter(x * y, win) - ux * uy)
        S = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        vals.append(S[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))


def cube_assessment(truth: torch.Tensor, pred: torch.Tensor, ratio: float = 4.0):
    """Per-cube rows (MPSNR, MSSIM, ERGAS, SAM, CrossCorrelation, RMSE) of NCHW batches: the validation loop's clamp to
    [0,1] and HWC layout (sr_gae.py:474-475, 483-484), then quality_assessment's indices (eval_hsi.py:217-238)."""
    out = []
    for t, p in zip(truth, pred):
        th = t.clamp(0, 1).permute(1, 2, 0).numpy()
        ph = p.clamp(0, 1).permute(1, 2, 0).numpy()
        out.append((mpsnr(th, ph), mssim(th, ph), ergas(th, ph, ratio), sam_deg(th, ph), cross_correlation(th, ph), rmse(th, ph)))
    return out


def cube_metrics(truth: torch.Tensor, pred: torch.Tensor):
    """Per-cube (MPSNR, SAM) of NCHW batches the way the validation loop computes them: clamp to [0,1], HWC, then
    compare_mpsnr / compare_sam (sr_gae.py:474-475, eval_hsi.py:110-121, :47-65)."""
    out = []
    for t, p in zip(truth, pred):
        th = t.clamp(0, 1).permute(1, 2, 0).numpy()
        ph = p.clamp(0, 1).permute(1, 2, 0).numpy()
        out.append((mpsnr(th, ph), sam_deg(th, ph)))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# training step (SURVEY 8f row N2 - the next scope row; pinned by tests/golden/train_step.npz)
def q_sample(x_start: torch.Tensor, level: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """diffusion.py:213-220: level is the continuous sqrt(alpha_bar) drawn per sample, shape [B,1,1,1]."""
    return level * x_start + (1 - level ** 2).sqrt() * noise


def train_step(sd: SD, cfg: dict, hr: torch.Tensor, sr: torch.Tensor, noise: torch.Tensor, levels: np.ndarray,
               loss_type: str = "l1"):
    """One optimisation step's loss and gradients: p_losses (diffusion.py:222-250) with the per-sample noise levels
    given, then DDPM.optimize_parameters' normalisation loss = sum / (b*c*h*w) (model.py:49-55).  Dropout is the
    identity (the pinned configuration trains with dropout 0).  Returns (loss_sum, loss, {param: grad})."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lv = torch.as_tensor(np.asarray(levels), dtype=torch.float32).view(-1, 1)   # FloatTensor(...) in the reference
    x_noisy = q_sample(hr, lv.view(-1, 1, 1, 1), noise)
    with torch.enable_grad():
        eps = unet_forward(params, cfg, torch.cat([sr, x_noisy], dim=1), lv)
        diff = noise - eps
        loss_sum = diff.abs().sum() if loss_type == "l1" else (diff ** 2).sum()
        b, c, h, w = hr.shape
        loss = loss_sum / int(b * c * h * w)
        loss.backward()
    return float(loss_sum.detach()), float(loss.detach()), {k: v.grad for k, v in params.items()}

