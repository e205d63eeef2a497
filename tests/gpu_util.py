"""Helpers shared by the -m gpu tests: raw-kernel launchers through the C ABI test hooks and error metrics."""
import ctypes as C

import numpy as np
import torch

from hsi_dmgasr_b200 import _lib


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def randn(shape, seed, device="cuda", scale=1.0):
    t = torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)) * scale
    return t.to(device)


def act_dtype(prec):
    return torch.bfloat16 if _lib.precision_code(prec) == _lib.BF16 else torch.float32


def to_nhwc(x: torch.Tensor, prec) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous().to(act_dtype(prec))


def from_nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.float().permute(0, 3, 1, 2).contiguous()


def conv2d(backend, prec, x0, x1, weight, bias, *, ksize, stride=1, up=0, nbias=None, act=0, scale=1.0, resid=None,
           out_nchw=False, src_nchw=False):
    """x0/x1: NCHW fp32 torch tensors (converted to the kernel's layout here). Returns NCHW fp32."""
    lib = _lib.load()
    p = _lib.precision_code(prec)
    n, c0, h, w = x0.shape
    c1 = 0 if x1 is None else x1.shape[1]
    if src_nchw:
        s0, s1 = x0.contiguous(), None if x1 is None else x1.contiguous()
    else:
        s0, s1 = to_nhwc(x0, prec), None if x1 is None else to_nhwc(x1, prec)
    he, we = (2 * h, 2 * w) if up else (h, w)
    ho, wo = ((he + 1) // 2, (we + 1) // 2) if stride == 2 else (he, we)
    cout = weight.shape[0]
    r = None if resid is None else to_nhwc(resid, prec)
    if out_nchw:
        out = torch.full((n, cout, ho, wo), float("nan"), device=x0.device, dtype=torch.float32)
    else:
        out = torch.full((n, ho, wo, cout), float("nan"), device=x0.device, dtype=torch.float32).to(act_dtype(prec))
    wt = weight.contiguous().float()
    nb = None if nbias is None else nbias.contiguous().float()
    _lib.check(lib.hsidm_debug_conv2d(backend, p, s0.data_ptr(), c0, _lib.ptr(s1), c1, 1 if src_nchw else 0, n, h, w, up,
                                      stride, wt.data_ptr(), _lib.ptr(None if bias is None else bias.contiguous().float()),
                                      cout, ksize, _lib.ptr(nb), 0 if nb is None else nb.shape[1], act, float(scale),
                                      _lib.ptr(r), out.data_ptr(), 1 if out_nchw else 0))
    return out if out_nchw else from_nhwc(out)


def tc_flag() -> int:
    v = C.c_int(0)
    _lib.check(_lib.load().hsidm_debug_tc_error_flag(C.byref(v)))
    return v.value


def groupnorm(prec, x0, x1, groups, gamma, beta, swish, eps=1e-5):
    lib = _lib.load()
    p = _lib.precision_code(prec)
    n, c0, h, w = x0.shape
    c1 = 0 if x1 is None else x1.shape[1]
    s0, s1 = to_nhwc(x0, prec), None if x1 is None else to_nhwc(x1, prec)
    out = torch.empty((n, h, w, c0 + c1), device=x0.device, dtype=act_dtype(prec))
    _lib.check(lib.hsidm_debug_groupnorm(p, s0.data_ptr(), c0, _lib.ptr(s1), c1, n, h * w, groups,
                                         gamma.contiguous().data_ptr(), beta.contiguous().data_ptr(), eps, int(swish),
                                         out.data_ptr()))
    return from_nhwc(out)
