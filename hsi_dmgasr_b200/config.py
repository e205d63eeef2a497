"""``config/*.json`` schema support: the parts of ``core/logger.py`` the hot path depends on
(logger.py:21-94 ``parse`` minus directory creation / env mutation, and :97-112 ``NoneDict``)."""
from __future__ import annotations

import json
from collections import OrderedDict


class NoneDict(dict):
    def __missing__(self, key):
        return None


def dict_to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict(**{k: dict_to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [dict_to_nonedict(v) for v in opt]
    return opt


def load_json_with_comments(path: str) -> OrderedDict:
    """The reference strips everything after '//' on each line before json.loads (logger.py:26-32)."""
    text = ""
    with open(path, "r") as f:
        for line in f:
            text += line.split("//")[0] + "\n"
    return json.loads(text, object_pairs_hook=OrderedDict)


def parse(config_path: str, phase: str = "val", gpu_ids=None, debug: bool = False) -> NoneDict:
    """opt dict with the keys the model layer reads: phase, gpu_ids, distributed, model.*, path.resume_state.

    Unlike the reference it neither creates experiment directories nor exports CUDA_VISIBLE_DEVICES; device
    placement is the launcher's job (one process per GPU)."""
    opt = load_json_with_comments(config_path)
    opt["phase"] = phase
    if debug:
        opt["name"] = "debug_{}".format(opt["name"])
    if gpu_ids is not None:
        opt["gpu_ids"] = [int(i) for i in str(gpu_ids).split(",")]
    gpu_list = ",".join(str(x) for x in (opt.get("gpu_ids") or []))
    opt["distributed"] = len(gpu_list) > 1                       # string length, exactly as logger.py:56-59
    if "debug" in opt["name"]:
        opt["model"]["beta_schedule"]["train"]["n_timestep"] = 10
        opt["model"]["beta_schedule"]["val"]["n_timestep"] = 10
    return dict_to_nonedict(opt)
