"""Drop-in for ``model/__init__.py::create_model`` and ``model/model.py::DDPM`` (inference side).

Same methods and attributes as the reference wrapper: ``feed_data``, ``test``, ``sample``, ``set_loss``,
``set_new_noise_schedule``, ``get_current_visuals``, ``save_network``, ``load_network`` (incl. its habit of
dropping the first/last conv weights, model.py:189-192).  ``optimize_parameters`` belongs to the training row
 and raises.
"""
from __future__ import annotations

import logging
import os
from collections import OrderedDict

import torch
from torch import nn

from . import _lib, networks

logger = logging.getLogger("base")

_DROPPED_ON_LOAD = ("denoise_fn.downs.0.weight", "denoise_fn.final_conv.block.3.weight",
                    "denoise_fn.final_conv.block.3.bias")


class DDPM:
    def __init__(self, opt):
        self.opt = opt
        self.device = torch.device("cuda" if opt["gpu_ids"] is not None else "cpu")   # base_model.py:14-15
        self.begin_step = 0
        self.begin_epoch = 0
        self.netG = networks.define_G(opt).to(self.device)
        self.schedule_phase = None
        self.set_loss()
        self.set_new_noise_schedule(opt["model"]["beta_schedule"]["train"], schedule_phase="train")
        if opt["phase"] == "train":
            self.netG.train()
            # model.py:26-44: which parameters train, and the Adam optimiser over them
            if opt["model"]["finetune_norm"]:
                optim_params = []
                for k, v in self.netG.named_parameters():
                    v.requires_grad = False
                    if k.find("transformer") >= 0:
                        v.requires_grad = True
                        v.data.zero_()
                        optim_params.append(v)
                        logger.info("Params [{:s}] initialized to 0 and will optimize.".format(k))
            else:
                optim_params = list(self.netG.parameters())
            self.optG = torch.optim.Adam(optim_params, lr=opt["train"]["optimizer"]["lr"])
            self.log_dict = OrderedDict()
        self.load_network()

    # base_model.py:34-45
    def set_device(self, x):
        if isinstance(x, dict):
            for key, item in x.items():
                if item is not None:
                    x[key] = item.to(self.device)
        elif isinstance(x, list):
            x = [item.to(self.device) if item is not None else None for item in x]
        else:
            x = x.to(self.device)
        return x

    def feed_data(self, data):
        self.data = self.set_device(data)

    def optimize_parameters(self, world: int = 1):
        """model.py:49-59.  ``world`` > 1 (one process per GPU) averages the gradients over the ranks with one NCCL
        all-reduce between backward and the optimiser step (the reference uses nn.DataParallel in one process)."""
        self.optG.zero_grad()
        l_pix = self.netG(self.data)
        b, c, h, w = self.data["HR"].shape
        l_pix = l_pix.sum() / int(b * c * h * w)
        l_pix.backward(retain_graph=True)
        if world > 1:
            from .diffusion import allreduce_gradients
            allreduce_gradients(self.netG, world)
        self.optG.step()
        self.log_dict["l_pix"] = l_pix.item()

    def test(self, continous=False):
        self.netG.eval()
        with torch.no_grad():
            self.SR = self.netG.super_resolution(self.data["SR"], continous)
        _lib.check_health(self.SR.device)
        self.netG.train()

    def sample(self, batch_size=1, continous=False):
        self.netG.eval()
        with torch.no_grad():
            self.SR = self.netG.sample(batch_size, continous)
        self.netG.train()

    def set_loss(self):
        self.netG.set_loss(self.device)

    def set_new_noise_schedule(self, schedule_opt, schedule_phase="train"):
        if self.schedule_phase is None or self.schedule_phase != schedule_phase:
            self.schedule_phase = schedule_phase
            self.netG.set_new_noise_schedule(schedule_opt, self.device)

    def get_current_log(self):
        return self.log_dict

    def get_current_visuals(self, need_LR=True, sample=False):
        out = OrderedDict()
        if sample:
            out["SAM"] = self.SR.detach().float().cpu()
            return out
        out["SR"] = self.SR.detach().float().cpu()
        out["INF"] = self.data["SR"].detach().float().cpu()
        out["HR"] = self.data["HR"].detach().float().cpu()
        out["LR"] = self.data["LR"].detach().float().cpu() if (need_LR and "LR" in self.data) else out["INF"]
        return out

    def get_network_description(self, network):
        return str(network), sum(p.numel() for p in network.parameters())

    def print_network(self):
        s, n = self.get_network_description(self.netG)
        logger.info("Network G structure: {}, with parameters: {:,d}".format(self.netG.__class__.__name__, n))
        logger.info(s)

    def save_network(self, epoch, iter_step):
        """Writes ``I{iter}_E{epoch}_gen.pth`` = GaussianDiffusion.state_dict() on CPU (model.py:125-145)."""
        gen_path = os.path.join(self.opt["path"]["checkpoint"], "I{}_E{}_gen.pth".format(iter_step, epoch))
        state = OrderedDict((k, v.cpu()) for k, v in self.netG.state_dict().items())
        torch.save(state, gen_path)
        logger.info("Saved model in [{:s}] ...".format(gen_path))
        return gen_path

    def load_network(self):
        """``path.resume_state`` is a prefix; ``_gen.pth`` is appended; three keys are dropped; strict=False."""
        load_path = self.opt["path"]["resume_state"]
        if load_path is None:
            return
        logger.info("Loading pretrained model for G [{:s}] ...".format(load_path))
        ckpt = torch.load("{}_gen.pth".format(load_path), map_location="cpu")
        kept = {k: v for k, v in ckpt.items() if k not in _DROPPED_ON_LOAD}
        self.netG.load_state_dict(kept, strict=False)
        if self.opt["phase"] == "train":
            self.begin_step = 0
            self.begin_epoch = 0


def create_model(opt):
    m = DDPM(opt)
    logger.info("Model [{:s}] is created.".format(m.__class__.__name__))
    return m
