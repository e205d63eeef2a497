"""ctypes binding of libhsidm_b200.so (the C ABI declared in include/hsidm.h).

There is deliberately no fallback here: if the shared library is missing or a CUDA device is absent, the
product path raises.  ``load()`` only dlopens (safe on a CPU box); compute entry points need a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhsidm_b200.so")

F32, BF16 = 0, 1
MAX_LEVELS = 8

STATUS_NAMES = {0: "OK", -1: "BAD_SHAPE", -2: "BAD_DTYPE", -3: "UNSUPPORTED_CFG", -4: "CUDA_ERROR",
                -5: "OOM_WORKSPACE", -6: "BAD_ARG", -7: "BAD_STATE"}


class HsidmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hsidm {STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


class UNetCfg(C.Structure):
    _fields_ = [("in_channel", C.c_int32), ("out_channel", C.c_int32), ("inner_channel", C.c_int32),
                ("norm_groups", C.c_int32), ("n_mults", C.c_int32), ("channel_mults", C.c_int32 * MAX_LEVELS),
                ("n_attn_res", C.c_int32), ("attn_res", C.c_int32 * MAX_LEVELS), ("res_blocks", C.c_int32),
                ("dropout", C.c_float), ("image_size", C.c_int32), ("precision", C.c_int32)]


class GAECfg(C.Structure):
    _fields_ = [("n_colors", C.c_int32), ("n_subs", C.c_int32), ("n_ovls", C.c_int32), ("n_feats", C.c_int32),
                ("trunk_feats", C.c_int32), ("n_blocks", C.c_int32), ("trunk_blocks", C.c_int32),
                ("latent", C.c_int32), ("precision", C.c_int32)]


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
# name -> (restype, argtypes); mirrors include/hsidm.h and include/hsidm_debug.h one to one
SIGNATURES = {
    "hsidm_version": (C.c_int, []),
    "hsidm_last_error": (C.c_char_p, []),
    "hsidm_launch_count": (C.c_int64, []),
    "hsidm_prof_enable": (C.c_int, [C.c_int]),
    "hsidm_prof_dump": (C.c_int, [C.c_char_p]),
    "hsidm_prof_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "hsidm_ctx_create": (C.c_int, [C.POINTER(UNetCfg), C.c_int, C.POINTER(_P)]),
    "hsidm_ctx_destroy": (C.c_int, [_P]),
    "hsidm_unet_param_count": (C.c_int, [_P]),
    "hsidm_unet_param_name": (C.c_char_p, [_P, C.c_int]),
    "hsidm_unet_set_param": (C.c_int, [_P, C.c_char_p, _P, _I64P, C.c_int]),
    "hsidm_unet_commit": (C.c_int, [_P]),
    "hsidm_unet_params_changed": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int), _P]),
    "hsidm_gae_params_changed": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int), _P]),
    "hsidm_check_health": (C.c_int, []),
    "hsidm_set_schedule": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int]),
    "hsidm_unet_forward": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P]),
    "hsidm_posterior_step": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int64, _P]),
    "hsidm_sample": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_uint64, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "hsidm_sample_at": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_uint64, C.c_int64, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "hsidm_randn": (C.c_int, [_P, C.c_int64, C.c_uint64, C.c_int64, _P]),
    "hsidm_blend_tiles": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hsidm_train_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, _P, _P, _P]),
    "hsidm_train_backward": (C.c_int, [_P, C.c_float, _P]),
    "hsidm_train_grad_numel": (C.c_int64, [_P]),
    "hsidm_train_grad_offset": (C.c_int64, [_P, C.c_int]),
    "hsidm_snapshot_count": (C.c_int, [_P]),
    "hsidm_num_timesteps": (C.c_int, [_P]),
    "hsidm_ctx_bytes": (C.c_int64, [_P]),
    "hsidm_gae_create": (C.c_int, [C.POINTER(GAECfg), C.c_int, C.POINTER(_P)]),
    "hsidm_gae_destroy": (C.c_int, [_P]),
    "hsidm_gae_param_count": (C.c_int, [_P]),
    "hsidm_gae_param_name": (C.c_char_p, [_P, C.c_int]),
    "hsidm_gae_set_param": (C.c_int, [_P, C.c_char_p, _P, _I64P, C.c_int]),
    "hsidm_gae_commit": (C.c_int, [_P]),
    "hsidm_gae_groups": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hsidm_gae_encode": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "hsidm_gae_decode": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "hsidm_debug_conv2d": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P, C.c_int64, C.c_int, C.c_float,
                                     _P, _P, C.c_int]),
    "hsidm_debug_groupnorm": (C.c_int, [C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                                        C.c_float, C.c_int, _P]),
    "hsidm_bicubic_upsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hsidm_imresize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "hsidm_quality_assessment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "hsidm_quality_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hsidm_debug_conv_mode": (C.c_int, [C.c_int, C.c_int]),
    "hsidm_debug_halo_timing": (C.c_int, [C.c_void_p]),
    "hsidm_debug_tc_error_flag": (C.c_int, [C.POINTER(C.c_int)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library (build it first with ``python -m hsi_dmgasr_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsidmError(-7, f"{LIB_PATH} is missing: run `python -m hsi_dmgasr_b200.build` (needs nvcc). "
                             "There is no fallback implementation.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here means header and library drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().hsidm_last_error()
        raise HsidmError(status, msg.decode("utf-8", "replace") if msg else "")


def check_health(device) -> None:
    """Raise if any tensor-core launch on `device` since the last check timed out on a pipeline barrier (its output
    would be incomplete).  Synchronises the device."""
    import torch
    with torch.cuda.device(device):
        check(load().hsidm_check_health())


def precision_code(precision) -> int:
    if precision in (F32, "f32", "fp32", "float32"):
        return F32
    if precision in (BF16, "bf16", "bfloat16"):
        return BF16
    raise ValueError(f"unknown precision {precision!r} (use 'fp32' or 'bf16')")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_f32(t, name: str):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HsidmError(-6, f"{name} must be a CUDA tensor: the hsidm hot path has no CPU implementation")
    if t.dtype != torch.float32:
        raise HsidmError(-2, f"{name} must be float32 (got {t.dtype})")
    return t.contiguous()
