"""Drop-in for ``model/networks.py::define_G`` (networks.py:83-116): builds UNet + GaussianDiffusion from ``opt``."""
from __future__ import annotations

import logging

import torch
from torch import nn
from torch.nn import init

from .diffusion import GaussianDiffusion
from .spec import UNetConfig
from .unet import UNet

logger = logging.getLogger("base")


def init_weights(net: nn.Module, init_type: str = "kaiming", scale: float = 1, std: float = 0.02) -> None:
    """Weight initialisers of networks.py:13-74 (train phase only; the val phase keeps default init)."""
    logger.info("Initialization method [{:s}]".format(init_type))
    for m in net.modules():
        is_conv = isinstance(m, nn.Conv2d)
        is_lin = isinstance(m, nn.Linear)
        if not (is_conv or is_lin):
            continue
        if init_type == "normal":
            init.normal_(m.weight.data, 0.0, std)
        elif init_type == "kaiming":
            init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            m.weight.data *= scale
        elif init_type == "orthogonal":
            init.orthogonal_(m.weight.data, gain=1)
        else:
            raise NotImplementedError("initialization method [{:s}] not implemented".format(init_type))
        if m.bias is not None:
            m.bias.data.zero_()
    for m in net.modules():        # the edits above go through .data: make the native weight copies stale explicitly
        if hasattr(m, "invalidate_native"):
            m.invalidate_native()


def define_G(opt):
    model_opt = opt["model"]
    if model_opt["which_model_G"] != "sr3":
        raise NotImplementedError("only which_model_G == 'sr3' is on the HSI-DMGASR hot path "
                                  "(the 'ddpm' BatchNorm variant is used by no shipped config)")
    if ("norm_groups" not in model_opt["unet"]) or model_opt["unet"]["norm_groups"] is None:
        model_opt["unet"]["norm_groups"] = 32
    cfg = UNetConfig.from_opt(model_opt)
    precision = model_opt["unet"].get("precision") if hasattr(model_opt["unet"], "get") else None
    model = UNet(in_channel=cfg.in_channel, out_channel=cfg.out_channel, norm_groups=cfg.norm_groups,
                 inner_channel=cfg.inner_channel, channel_mults=cfg.channel_mults, attn_res=cfg.attn_res,
                 res_blocks=cfg.res_blocks, dropout=cfg.dropout, image_size=cfg.image_size, precision=precision)
    netG = GaussianDiffusion(model, image_size=model_opt["diffusion"]["image_size"],
                             channels=model_opt["diffusion"]["channels"], loss_type="l1",
                             conditional=model_opt["diffusion"]["conditional"],
                             schedule_opt=model_opt["beta_schedule"]["train"])
    if opt["phase"] == "train":
        init_weights(netG, init_type="orthogonal")
    if opt["gpu_ids"] and opt["distributed"]:
        # The reference wraps in nn.DataParallel here (networks.py:113-115) but *bypasses* it at inference
        # (model.py:64-66). Inference shards by cube/tile with one process per GPU instead (pipeline.py).
        assert torch.cuda.is_available()
    return netG
