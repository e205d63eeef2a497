"""Static description of the two networks on the hot path: SR3 UNet and the group autoencoder (GAE).

Everything here is host-side bookkeeping (layer lists, parameter names/shapes, band-group geometry).
It mirrors what the reference derives in ``UNet.__init__`` (model/sr3_modules/unet.py:163-236) and
``GAE.__init__`` (AE.py:256-280) so that ``state_dict`` keys and checkpoint layouts stay drop-in.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


@dataclass(frozen=True)
class UNetConfig:
    """Constructor arguments of the reference UNet (unet.py:163-176), as passed by define_G (networks.py:91-101)."""
    in_channel: int = 6
    out_channel: int = 3
    inner_channel: int = 32
    norm_groups: int = 32
    channel_mults: Tuple[int, ...] = (1, 2, 4, 8, 8)
    attn_res: Tuple[int, ...] = (8,)
    res_blocks: int = 3
    dropout: float = 0.0
    image_size: int = 128

    @staticmethod
    def from_opt(model_opt: dict) -> "UNetConfig":
        u = model_opt["unet"]
        groups = u.get("norm_groups") if isinstance(u, dict) else u["norm_groups"]
        return UNetConfig(
            in_channel=u["in_channel"], out_channel=u["out_channel"], inner_channel=u["inner_channel"],
            norm_groups=32 if groups is None else int(groups),          # networks.py:89-90
            channel_mults=tuple(u["channel_multiplier"]), attn_res=tuple(u["attn_res"] or ()),
            res_blocks=u["res_blocks"], dropout=u["dropout"] or 0.0,
            image_size=model_opt["diffusion"]["image_size"])

    def as_dict(self) -> dict:
        return dict(inner_channel=self.inner_channel, channel_multiplier=list(self.channel_mults),
                    attn_res=list(self.attn_res), res_blocks=self.res_blocks, image_size=self.image_size,
                    norm_groups=self.norm_groups, in_channel=self.in_channel, out_channel=self.out_channel)


@dataclass(frozen=True)
class Layer:
    """One entry of ``downs`` / ``mid`` / ``ups``."""
    kind: str                 # "conv" | "res" | "down" | "up"
    name: str                 # state_dict prefix, e.g. "downs.4"
    cin: int = 0              # for "res": channels of the (possibly concatenated) input
    cout: int = 0
    attn: bool = False
    skip: int = 0             # for "res" in ups: channels that come from the popped skip tensor


def unet_layers(cfg: UNetConfig) -> Tuple[List[Layer], List[Layer], List[Layer]]:
    """Layer list with the attention placement rule of unet.py:195-233 (attn_res vs *config* image_size)."""
    ic = cfg.inner_channel
    pre, now = ic, cfg.image_size
    feat = [pre]
    downs = [Layer("conv", "downs.0", cfg.in_channel, ic)]
    nlev = len(cfg.channel_mults)
    for lev, mult in enumerate(cfg.channel_mults):
        for _ in range(cfg.res_blocks):
            downs.append(Layer("res", f"downs.{len(downs)}", pre, ic * mult, now in cfg.attn_res))
            pre = ic * mult
            feat.append(pre)
        if lev + 1 < nlev:
            downs.append(Layer("down", f"downs.{len(downs)}", pre, pre))
            feat.append(pre)
            now //= 2
    mid = [Layer("res", "mid.0", pre, pre, True), Layer("res", "mid.1", pre, pre, False)]
    ups: List[Layer] = []
    for lev in reversed(range(nlev)):
        mult = cfg.channel_mults[lev]
        for _ in range(cfg.res_blocks + 1):
            sk = feat.pop()
            ups.append(Layer("res", f"ups.{len(ups)}", pre + sk, ic * mult, now in cfg.attn_res, sk))
            pre = ic * mult
        if lev > 0:
            ups.append(Layer("up", f"ups.{len(ups)}", pre, pre))
            now *= 2
    return downs, mid, ups


def unet_param_shapes(cfg: UNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameter names (relative to ``denoise_fn.``) and shapes, in the reference's state_dict order."""
    ic = cfg.inner_channel
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def lin(p, i, o):
        out[p + ".weight"] = (o, i)
        out[p + ".bias"] = (o,)

    def conv(p, i, o, k, bias=True):
        out[p + ".weight"] = (o, i, k, k)
        if bias:
            out[p + ".bias"] = (o,)

    def gn(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    lin("noise_level_mlp.1", ic, 4 * ic)
    lin("noise_level_mlp.3", 4 * ic, ic)
    downs, mid, ups = unet_layers(cfg)
    final_c = ic
    for L in downs + mid + ups:
        if L.kind == "conv":
            conv(L.name, L.cin, L.cout, 3)
        elif L.kind in ("down", "up"):
            conv(L.name + ".conv", L.cin, L.cout, 3)
        else:
            r = L.name + ".res_block"
            lin(r + ".noise_func.noise_func.0", ic, L.cout)
            gn(r + ".block1.block.0", L.cin)
            conv(r + ".block1.block.3", L.cin, L.cout, 3)
            gn(r + ".block2.block.0", L.cout)
            conv(r + ".block2.block.3", L.cout, L.cout, 3)
            if L.cin != L.cout:
                conv(r + ".res_conv", L.cin, L.cout, 1)
            if L.attn:
                a = L.name + ".attn"
                gn(a + ".norm", L.cout)
                conv(a + ".qkv", L.cout, 3 * L.cout, 1, bias=False)
                conv(a + ".out", L.cout, L.cout, 1)
        final_c = L.cout
    gn("final_conv.block.0", final_c)
    conv("final_conv.block.3", final_c, cfg.out_channel if cfg.out_channel is not None else cfg.in_channel, 3)
    return out


# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class GAEGeometry:
    """Band-group geometry of one GAE checkpoint (AE.py:256-280)."""
    n_colors: int
    n_subs: int
    n_ovls: int
    n_feats: int = 64
    trunk_feats: int = 32
    enc_blocks: int = 3
    trunk_blocks: int = 2
    latent: int = 3

    @property
    def G(self) -> int:
        return math.ceil((self.n_colors - self.n_ovls) / (self.n_subs - self.n_ovls))

    def groups(self) -> Tuple[List[int], List[int]]:
        start, end = [], []
        step = self.n_subs - self.n_ovls
        for g in range(self.G):
            s, e = step * g, step * g + self.n_subs
            if e > self.n_colors:                       # last group is pulled back inside the cube
                s, e = self.n_colors - self.n_subs, self.n_colors
            start.append(s)
            end.append(e)
        return start, end

    def band_counts(self) -> List[int]:
        cnt = [0] * self.n_colors
        for s, e in zip(*self.groups()):
            for c in range(s, e):
                cnt[c] += 1
        return cnt

    def as_dict(self) -> dict:
        return dict(n_colors=self.n_colors, n_subs=self.n_subs, n_ovls=self.n_ovls)


# Geometry of the four shipped checkpoints GAE_pretrained/GAE_4_{Cav,Har,Chi,Pav}.pth.
GAE_PRESETS: Dict[str, GAEGeometry] = {
    "Cav": GAEGeometry(31, 8, 2), "Har": GAEGeometry(31, 8, 2),
    "Chi": GAEGeometry(128, 16, 4), "Pav": GAEGeometry(102, 16, 4),
}


def _branch_shapes(out, p, cin, feats, blocks):
    out[p + ".head.weight"] = (feats, cin, 3, 3)
    out[p + ".head.bias"] = (feats,)
    red = feats // 3                                            # CALayer(n_feats, 3), common.py:267
    for b in range(blocks):
        q = f"{p}.body.net.{b}"
        for j in (0, 2):
            out[f"{q}.spa.body.{j}.weight"] = (feats, feats, 3, 3)
            out[f"{q}.spa.body.{j}.bias"] = (feats,)
        for j in (0, 2):
            out[f"{q}.spc.body.{j}.weight"] = (feats, feats, 1, 1)
            out[f"{q}.spc.body.{j}.bias"] = (feats,)
        out[f"{q}.spc.body.3.conv_du.0.weight"] = (red, feats, 1, 1)
        out[f"{q}.spc.body.3.conv_du.0.bias"] = (red,)
        out[f"{q}.spc.body.3.conv_du.2.weight"] = (feats, red, 1, 1)
        out[f"{q}.spc.body.3.conv_du.2.bias"] = (feats,)


def gae_param_shapes(g: GAEGeometry) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameter names/shapes of GAE.state_dict() (Encoder, Decoder, trunk, final)."""
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _branch_shapes(out, "Encoder.branch", g.n_subs, g.n_feats, g.enc_blocks)
    out["Encoder.final.weight"] = (g.latent, g.n_feats, 3, 3)
    out["Encoder.final.bias"] = (g.latent,)
    _branch_shapes(out, "Decoder.branch", g.latent, g.n_feats, g.enc_blocks)
    out["Decoder.final.weight"] = (g.n_subs, g.n_feats, 3, 3)
    out["Decoder.final.bias"] = (g.n_subs,)
    _branch_shapes(out, "trunk", g.n_colors, g.trunk_feats, g.trunk_blocks)
    out["final.weight"] = (g.n_colors, g.trunk_feats, 3, 3)
    out["final.bias"] = (g.n_colors,)
    return out
