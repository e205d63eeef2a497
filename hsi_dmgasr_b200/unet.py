"""Drop-in for ``model/sr3_modules/unet.py::UNet``: same constructor, same ``state_dict`` keys, same
``forward(x, time)`` contract - but forward is one call into the native library (``hsidm_unet_forward``).

The module tree below holds parameters only; its leaves (``nn.Conv2d`` / ``nn.Linear`` / ``nn.GroupNorm``) are
never called.  They exist so that ``state_dict()`` / ``load_state_dict()`` / ``.to()`` / optimizers see exactly
the tensors the reference exposes (``denoise_fn.downs.{i}.res_block.block1.block.{0,3}.weight`` ...).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib
from .spec import Layer, UNetConfig, unet_layers

_DEFAULT_PRECISION = "bf16"


def set_default_precision(precision: str) -> None:
    """'bf16' (tensor cores, parity gate 2e-2) or 'fp32' (CUDA cores, parity gate 1e-4)."""
    global _DEFAULT_PRECISION
    _lib.precision_code(precision)
    _DEFAULT_PRECISION = precision


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


def _holder(**children) -> nn.ModuleDict:
    return nn.ModuleDict({k: v for k, v in children.items() if v is not None})


def _norm_act_conv(groups: int, cin: int, cout: int) -> nn.ModuleDict:
    # indices 0 (GroupNorm) and 3 (Conv2d) of the reference's Sequential; 1 = Swish, 2 = Dropout have no state
    return _holder(block=nn.ModuleDict({"0": nn.GroupNorm(groups, cin), "3": nn.Conv2d(cin, cout, 3, padding=1)}))


def _res_layer(L: Layer, emb: int, groups: int) -> nn.ModuleDict:
    rb = nn.ModuleDict()
    rb["noise_func"] = _holder(noise_func=nn.ModuleDict({"0": nn.Linear(emb, L.cout)}))
    rb["block1"] = _norm_act_conv(groups, L.cin, L.cout)
    rb["block2"] = _norm_act_conv(groups, L.cout, L.cout)
    if L.cin != L.cout:
        rb["res_conv"] = nn.Conv2d(L.cin, L.cout, 1)
    out = nn.ModuleDict({"res_block": rb})
    if L.attn:
        out["attn"] = _holder(norm=nn.GroupNorm(groups, L.cout), qkv=nn.Conv2d(L.cout, 3 * L.cout, 1, bias=False),
                              out=nn.Conv2d(L.cout, L.cout, 1))
    return out


def _stack(layers: Sequence[Layer], emb: int, groups: int) -> nn.ModuleDict:
    mods = nn.ModuleDict()
    for i, L in enumerate(layers):
        if L.kind == "conv":
            mods[str(i)] = nn.Conv2d(L.cin, L.cout, 3, padding=1)
        elif L.kind == "res":
            mods[str(i)] = _res_layer(L, emb, groups)
        elif L.kind == "down":
            mods[str(i)] = _holder(conv=nn.Conv2d(L.cin, L.cout, 3, 2, 1))
        else:
            mods[str(i)] = _holder(conv=nn.Conv2d(L.cin, L.cout, 3, padding=1))
    return mods


class NativeHandle:
    """Owns one ``hsidm_ctx`` and keeps it in sync with the torch parameters it mirrors."""

    def __init__(self, cfg: UNetConfig, precision: str, device: torch.device):
        lib = _lib.load()
        c = _lib.UNetCfg()
        c.in_channel, c.out_channel, c.inner_channel = cfg.in_channel, cfg.out_channel, cfg.inner_channel
        c.norm_groups, c.res_blocks, c.dropout, c.image_size = cfg.norm_groups, cfg.res_blocks, cfg.dropout, cfg.image_size
        c.n_mults = len(cfg.channel_mults)
        for i, m in enumerate(cfg.channel_mults):
            c.channel_mults[i] = m
        c.n_attn_res = len(cfg.attn_res)
        for i, r in enumerate(cfg.attn_res):
            c.attn_res[i] = r
        c.precision = _lib.precision_code(precision)
        self.device = device
        self.precision = precision
        self.ptr = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(lib.hsidm_ctx_create(C.byref(c), idx, C.byref(self.ptr)))
        self.param_sig = None
        self.schedule_sig = None
        self._ptr_table = None

    def keys(self):
        lib = _lib.load()
        return [lib.hsidm_unet_param_name(self.ptr, i).decode() for i in range(lib.hsidm_unet_param_count(self.ptr))]

    def upload(self, state: dict) -> None:
        lib = _lib.load()
        # set_param copies with cudaMemcpy on the legacy stream: whatever produced the tensors on torch's current
        # (possibly non-blocking) stream has to be complete first
        torch.cuda.current_stream(self.device).synchronize()
        for key in self.keys():
            t = state[key].detach()
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=self.device, dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(lib.hsidm_unet_set_param(self.ptr, key.encode(), t.data_ptr(), shape, t.dim()))
        _lib.check(lib.hsidm_unet_commit(self.ptr))

    def data_changed(self, params) -> bool:
        """True if any parameter's bytes differ from the copy the library holds: catches in-place edits through
        ``.data`` (init_weights, finetune_norm, EMA swaps), which bump no autograd version counter."""
        ptrs = tuple(p.data_ptr() for p in params)
        if self._ptr_table is None or self._ptr_table[0] != ptrs:
            self._ptr_table = (ptrs, torch.tensor(ptrs, dtype=torch.int64, device=self.device))
        changed = C.c_int(0)
        _lib.check(_lib.load().hsidm_unet_params_changed(self.ptr, self._ptr_table[1].data_ptr(), len(ptrs), C.byref(changed),
                                                         _lib.stream_ptr(self.device)))
        return changed.value != 0

    def close(self) -> None:
        if self.ptr:
            try:
                _lib.load().hsidm_ctx_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = C.c_void_p()

    def __del__(self):
        self.close()


class UNet(nn.Module):
    """SR3 noise-level-conditioned UNet (reference unet.py:162-263), executed by hand-written sm_100a kernels."""

    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                 attn_res=(8), res_blocks=3, dropout=0, with_noise_level_emb=True, image_size=128,
                 precision: Optional[str] = None):
        super().__init__()
        if not with_noise_level_emb:
            raise NotImplementedError("with_noise_level_emb=False is not on the HSI-DMGASR path (no config uses it)")
        if isinstance(attn_res, int):      # the reference default `(8)` is an int and breaks `in`; accept it
            attn_res = (attn_res,)
        self.cfg = UNetConfig(in_channel=in_channel, out_channel=out_channel if out_channel is not None else in_channel,
                              inner_channel=inner_channel, norm_groups=norm_groups,
                              channel_mults=tuple(channel_mults), attn_res=tuple(attn_res or ()),
                              res_blocks=res_blocks, dropout=float(dropout or 0), image_size=image_size)
        self.precision = precision or _DEFAULT_PRECISION
        emb = inner_channel
        self.noise_level_mlp = nn.ModuleDict({"1": nn.Linear(emb, 4 * emb), "3": nn.Linear(4 * emb, emb)})
        downs, mid, ups = unet_layers(self.cfg)
        self.downs = _stack(downs, emb, norm_groups)
        self.mid = _stack(mid, emb, norm_groups)
        self.ups = _stack(ups, emb, norm_groups)
        self.final_conv = _norm_act_conv(norm_groups, ups[-1].cout, self.cfg.out_channel)
        self._native: Optional[NativeHandle] = None
        # Compare the library's weight copies with the live tensors on the device before every native call (one small
        # kernel + a 4-byte read, ~0.1 ms for 98 M parameters).  Switch off only if weights are never edited in place.
        self.track_data_edits = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_native() and None)

    # ---- native context management --------------------------------------------------------------------
    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def native(self, device: Optional[torch.device] = None) -> NativeHandle:
        """The hsidm_ctx for this module, (re)uploading weights if any parameter changed since the last call."""
        dev = device or next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.HsidmError(-4, "UNet parameters are not on a CUDA device; the hsidm hot path has no CPU fallback")
        h = self._native
        if h is None or h.device != dev or h.precision != self.precision:
            if h is not None:
                h.close()
            h = self._native = NativeHandle(self.cfg, self.precision, dev)
        sig = self._signature()
        if h.param_sig != sig or (self.track_data_edits and self._params_ok_for_compare(dev) and h.data_changed(self._plist())):
            h.upload(dict(self.state_dict()))
            h.param_sig = sig
        return h

    def _plist(self):
        named = dict(self.named_parameters())
        return [named[k] for k in self._native.keys()]      # the library's parameter order

    def _params_ok_for_compare(self, dev) -> bool:
        return all(p.device == dev and p.dtype == torch.float32 and p.is_contiguous() for p in self.parameters())

    def invalidate_native(self) -> "UNet":
        """Force a re-upload of every weight on the next call (use after editing parameters through ``.data``; with
        ``track_data_edits`` left on this is detected anyway)."""
        if self._native is not None:
            self._native.param_sig = None
        return self

    def set_precision(self, precision: str) -> "UNet":
        _lib.precision_code(precision)
        self.precision = precision
        return self

    # ---- reference API -------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        """x [N, in_channel, H, W], time (noise level) [N, 1] -> eps [N, out_channel, H, W] (unet.py:239-263)."""
        if self.training and self.cfg.dropout > 0 and torch.is_grad_enabled():
            raise NotImplementedError("training-mode forward (dropout + autograd) is not part of the inference hot path; "
                                      "call .eval() / torch.no_grad() as DDPM.test does (model/model.py:61-70)")
        x = _lib.require_cuda_f32(x, "x")
        time = _lib.require_cuda_f32(time, "time").reshape(-1)
        n, c, hh, ww = x.shape
        if time.numel() != n:
            raise _lib.HsidmError(-1, f"time has {time.numel()} entries for a batch of {n}")
        h = self.native(x.device)
        out = torch.empty((n, self.cfg.out_channel, hh, ww), device=x.device, dtype=torch.float32)
        lib = _lib.load()
        _lib.check(lib.hsidm_unet_forward(h.ptr, x.data_ptr(), c, None, 0, time.data_ptr(), 1, out.data_ptr(), n, hh, ww,
                                          _lib.stream_ptr(x.device)))
        return out
