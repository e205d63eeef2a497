"""Drop-in for the group autoencoder of ``AE.py`` (GAE / Encoder / Decoder / BranchUnit / SSPN / SSB) and the
``common.py`` blocks it is built from (ResBlock, ResAttentionBlock, CALayer, Upsampler).

The classes carry parameters with the reference's names (so ``GAE.state_dict()`` and the whole-module pickles
``GAE_pretrained/GAE_4_*.pth`` stay loadable); ``GAE.encode`` / ``GAE.decode`` call ``hsidm_gae_encode`` /
``hsidm_gae_decode``, which push all B*G band groups through the shared Encoder/Decoder as one batch.
Unlike the reference (AE.py:285, 313) nothing is pinned to 'cuda:0': the tensors' own device is used.
"""
from __future__ import annotations

import ctypes as C
import io
import math
import pickle
from typing import List, Optional

import torch
from torch import nn

from . import _lib
from .spec import GAEGeometry


def default_conv(in_channels, out_channels, kernel_size, bias=True, dilation=1):
    if dilation != 1:
        raise NotImplementedError("dilated convolutions are not used by the GAE")
    return nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2, bias=bias)


class CALayer(nn.Module):          # common.py:231-247
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv_du = nn.Sequential(nn.Conv2d(channel, channel // reduction, 1), nn.ReLU(inplace=False),
                                     nn.Conv2d(channel // reduction, channel, 1), nn.Sigmoid())


class ResBlock(nn.Module):         # common.py:163-182
    def __init__(self, conv, n_feats, kernel_size, bias=True, bn=False, act=None, res_scale=1):
        super().__init__()
        if bn:
            raise NotImplementedError("BatchNorm variants are not used by the GAE")
        self.body = nn.Sequential(conv(n_feats, n_feats, kernel_size, bias=bias), act or nn.ReLU(True),
                                  conv(n_feats, n_feats, kernel_size, bias=bias))
        self.res_scale = res_scale


class ResAttentionBlock(nn.Module):  # common.py:250-271
    def __init__(self, conv, n_feats, kernel_size, bias=True, bn=False, act=None, res_scale=1):
        super().__init__()
        if bn:
            raise NotImplementedError("BatchNorm variants are not used by the GAE")
        self.body = nn.Sequential(conv(n_feats, n_feats, kernel_size, bias=bias), act or nn.ReLU(True),
                                  conv(n_feats, n_feats, kernel_size, bias=bias), CALayer(n_feats, 3))
        self.res_scale = res_scale


class Upsampler(nn.Sequential):    # common.py:184-211; the GAE only ever builds scale=1 (an empty Sequential)
    def __init__(self, conv, scale, n_feats, bn=False, act=False, bias=True):
        if scale != 1:
            raise NotImplementedError("GAE branches use up_scale=1 (AE.py:192,225,268)")
        super().__init__()


class SSB(nn.Module):              # AE.py:102-109
    def __init__(self, n_feats, kernel_size, act, res_scale, conv=default_conv):
        super().__init__()
        self.spa = ResBlock(conv, n_feats, kernel_size, act=act, res_scale=res_scale)
        self.spc = ResAttentionBlock(conv, n_feats, 1, act=act, res_scale=res_scale)


class SSPN(nn.Module):             # AE.py:120-141
    def __init__(self, n_feats, n_blocks, act, res_scale):
        super().__init__()
        self.net = nn.Sequential(*[SSB(n_feats, 3, act=act, res_scale=res_scale) for _ in range(n_blocks)])


class BranchUnit(nn.Module):       # AE.py:145-165
    def __init__(self, n_colors, n_feats, n_blocks, act, res_scale, up_scale, use_tail=True, conv=default_conv):
        super().__init__()
        if use_tail:
            raise NotImplementedError("GAE branches are built with use_tail=False")
        self.head = nn.Conv2d(n_colors, n_feats, kernel_size=3, padding=1)
        self.body = SSPN(n_feats, n_blocks, act, res_scale)
        self.upsample = Upsampler(conv, up_scale, n_feats)
        self.tail = None


class _Coder(nn.Module):
    def __init__(self, input_channel, out_channel, n_feats=128):
        super().__init__()
        self.input_channel = input_channel
        self.out_channel = out_channel
        self.branch = BranchUnit(input_channel, n_feats=n_feats, n_blocks=3, act=nn.LeakyReLU(), res_scale=0.1,
                                 use_tail=False, up_scale=1, conv=default_conv)
        self.final = nn.Conv2d(n_feats, out_channel, kernel_size=3, padding=1)


class Encoder(_Coder):             # AE.py:168-199
    pass


class Decoder(_Coder):             # AE.py:202-242
    pass


class _GAEHandle:
    def __init__(self, geom: GAEGeometry, precision: str, device: torch.device):
        lib = _lib.load()
        c = _lib.GAECfg(geom.n_colors, geom.n_subs, geom.n_ovls, geom.n_feats, geom.trunk_feats, geom.enc_blocks,
                        geom.trunk_blocks, geom.latent, _lib.precision_code(precision))
        self.ptr = C.c_void_p()
        self.device, self.precision, self.geom = device, precision, geom
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(lib.hsidm_gae_create(C.byref(c), idx, C.byref(self.ptr)))
        self.param_sig = None
        self._ptr_table = None

    def keys(self):
        lib = _lib.load()
        return [lib.hsidm_gae_param_name(self.ptr, i).decode() for i in range(lib.hsidm_gae_param_count(self.ptr))]

    def data_changed(self, named_params: dict) -> bool:
        """Bitwise comparison of the library's weight copies with the live tensors (in-place ``.data`` edits)."""
        params = [named_params[k] for k in self.keys()]
        if not all(p.device == self.device and p.dtype == torch.float32 and p.is_contiguous() for p in params):
            return True
        ptrs = tuple(p.data_ptr() for p in params)
        if self._ptr_table is None or self._ptr_table[0] != ptrs:
            self._ptr_table = (ptrs, torch.tensor(ptrs, dtype=torch.int64, device=self.device))
        changed = C.c_int(0)
        _lib.check(_lib.load().hsidm_gae_params_changed(self.ptr, self._ptr_table[1].data_ptr(), len(ptrs), C.byref(changed),
                                                        _lib.stream_ptr(self.device)))
        return changed.value != 0

    def upload(self, state: dict) -> None:
        lib = _lib.load()
        torch.cuda.current_stream(self.device).synchronize()   # set_param copies on the legacy stream
        for i in range(lib.hsidm_gae_param_count(self.ptr)):
            key = lib.hsidm_gae_param_name(self.ptr, i).decode()
            t = state[key].detach().to(device=self.device, dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(lib.hsidm_gae_set_param(self.ptr, key.encode(), t.data_ptr(), shape, t.dim()))
        _lib.check(lib.hsidm_gae_commit(self.ptr))

    def close(self):
        if self.ptr:
            try:
                _lib.load().hsidm_gae_destroy(self.ptr)
            except Exception:
                pass
            self.ptr = C.c_void_p()

    def __del__(self):
        self.close()


class GAE(nn.Module):
    """Group autoencoder (AE.py:256-361): ``encode(x) -> list[G] of [B,3,H,W]``, ``decode(x, z_list) -> [B,C,H,W]``.

    ``precision`` defaults to fp32: the codec runs once per cube (either side of the T-step loop), costs <0.1 % of a
    full-sampling patch and its output feeds the MPSNR/SAM gates directly."""

    def __init__(self, Encoder=Encoder, Decoder=Decoder, n_subs=8, n_ovls=2, n_colors=31, n_feats=128,
                 precision: str = "fp32"):
        super().__init__()
        self.Encoder = Encoder(n_subs, 3, n_feats)
        self.Decoder = Decoder(3, n_subs, n_feats)
        self.device = "cuda:0"                         # attribute kept for pickle parity; never used for placement
        self.G = math.ceil((n_colors - n_ovls) / (n_subs - n_ovls))
        self.trunk = BranchUnit(n_colors, n_feats=32, n_blocks=2, act=nn.LeakyReLU(), res_scale=0.1, up_scale=1,
                                conv=default_conv, use_tail=False)
        self.final = nn.Conv2d(32, n_colors, kernel_size=3, padding=1)
        geom = GAEGeometry(n_colors, n_subs, n_ovls, n_feats)
        self.start_idx, self.end_idx = geom.groups()
        self.precision = precision

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_native", None)          # the hsidm_gae handle is process-local
        return state

    # ---- geometry is re-derived from the tensors so that unpickled reference objects (no __init__) work too ----------
    def geometry(self) -> GAEGeometry:
        head = self.Encoder.branch.head.weight
        n_subs, n_feats = head.shape[1], head.shape[0]
        n_colors = self.final.weight.shape[0]
        trunk_feats = self.trunk.head.weight.shape[0]
        starts = list(self.start_idx)
        n_ovls = n_subs - (starts[1] - starts[0]) if len(starts) > 1 else 0
        geom = GAEGeometry(n_colors, n_subs, n_ovls, n_feats, trunk_feats, len(self.Encoder.branch.body.net),
                           len(self.trunk.body.net), self.Encoder.final.weight.shape[0])
        if geom.groups() != (starts, list(self.end_idx)):
            raise _lib.HsidmError(-3, f"band-group layout {starts}/{list(self.end_idx)} does not follow AE.py:264-280")
        return geom

    def _handle(self, device: torch.device) -> _GAEHandle:
        if device.type != "cuda":
            raise _lib.HsidmError(-4, "GAE tensors are not on a CUDA device; the hsidm hot path has no CPU fallback")
        precision = self.__dict__.get("precision", "fp32")
        h: Optional[_GAEHandle] = self.__dict__.get("_native")
        if h is None or h.device != device or h.precision != precision:
            if h is not None:
                h.close()
            h = _GAEHandle(self.geometry(), precision, device)
            self.__dict__["_native"] = h
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if h.param_sig != sig or (self.__dict__.get("track_data_edits", True) and h.data_changed(dict(self.named_parameters()))):
            h.upload(dict(self.state_dict()))
            h.param_sig = sig
        return h

    def invalidate_native(self) -> "GAE":
        """Force a re-upload of every weight on the next encode/decode."""
        h = self.__dict__.get("_native")
        if h is not None:
            h.param_sig = None
        return self

    @torch.no_grad()
    def encode_batched(self, x: torch.Tensor) -> torch.Tensor:
        """[B,C,H,W] -> [B*G,3,H,W]; latent of (cube b, group g) at index b*G+g."""
        x = _lib.require_cuda_f32(x, "x")
        h = self._handle(x.device)
        b, c, hh, ww = x.shape
        if c != h.geom.n_colors:
            raise _lib.HsidmError(-1, f"cube has {c} bands, GAE expects {h.geom.n_colors}")
        z = torch.empty((b * h.geom.G, h.geom.latent, hh, ww), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().hsidm_gae_encode(h.ptr, x.data_ptr(), z.data_ptr(), b, hh, ww, _lib.stream_ptr(x.device)))
        return z

    @torch.no_grad()
    def decode_batched(self, z: torch.Tensor, clamp01: bool = False) -> torch.Tensor:
        """[B*G,3,H,W] -> [B,C,H,W] (overlap average + residual trunk; optional driver clamp sr_gae.py:474-475)."""
        z = _lib.require_cuda_f32(z, "z")
        h = self._handle(z.device)
        n, c, hh, ww = z.shape
        if n % h.geom.G or c != h.geom.latent:
            raise _lib.HsidmError(-1, f"latent batch {tuple(z.shape)} is not a multiple of G={h.geom.G} groups x {h.geom.latent} channels")
        b = n // h.geom.G
        y = torch.empty((b, h.geom.n_colors, hh, ww), device=z.device, dtype=torch.float32)
        _lib.check(_lib.load().hsidm_gae_decode(h.ptr, z.data_ptr(), y.data_ptr(), b, hh, ww, int(clamp01),
                                                _lib.stream_ptr(z.device)))
        return y

    # ---- reference API (AE.py:283-361) -----------------------------------------------------------------------------------
    def encode(self, x: torch.Tensor) -> List[torch.Tensor]:
        z = self.encode_batched(x)
        b = x.shape[0]
        g = z.shape[0] // b
        z = z.view(b, g, *z.shape[1:])
        return [z[:, k].contiguous() for k in range(g)]

    def decode(self, x: torch.Tensor, z_list) -> torch.Tensor:
        """``x`` is only used for its shape in the reference (AE.py:284); same here."""
        z = torch.stack([t.to(x.device) for t in z_list], dim=1)       # [B,G,3,H,W]
        return self.decode_batched(z.reshape(-1, *z.shape[2:]))

    def forward(self, x: torch.Tensor):
        z_list = self.encode(x)
        return self.decode(x, z_list), z_list


# ---- checkpoint loading ---------------------------------------------------------------------------------------------
_PICKLE_CLASSES = {
    ("__main__", "GAE"): GAE, ("__main__", "Encoder"): Encoder, ("__main__", "Decoder"): Decoder,
    ("__main__", "BranchUnit"): BranchUnit, ("__main__", "SSPN"): SSPN, ("__main__", "SSB"): SSB,
    ("AE", "GAE"): GAE, ("AE", "Encoder"): Encoder, ("AE", "Decoder"): Decoder, ("AE", "BranchUnit"): BranchUnit,
    ("AE", "SSPN"): SSPN, ("AE", "SSB"): SSB, ("SSPSR", "BranchUnit"): BranchUnit, ("SSPSR", "SSPN"): SSPN,
    ("SSPSR", "SSB"): SSB,
    ("common", "ResBlock"): ResBlock, ("common", "ResAttentionBlock"): ResAttentionBlock,
    ("common", "CALayer"): CALayer, ("common", "Upsampler"): Upsampler,
}


# Exact (module, name) pairs a whole-module GAE pickle may reference besides the remapped classes above: what the four
# shipped GAE_4_*.pth name (inspected with pickletools) plus the storage / rebuild helpers newer torch versions emit.
# Nothing under builtins / numpy / torch.hub is reachable: no eval, exec, getattr, __import__ or os.system.
_PICKLE_GLOBALS = {
    ("collections", "OrderedDict"), ("__builtin__", "set"), ("builtins", "set"),
    ("torch", "device"), ("torch", "Size"), ("torch", "FloatStorage"), ("torch", "HalfStorage"), ("torch", "BFloat16Storage"),
    ("torch", "DoubleStorage"), ("torch", "LongStorage"), ("torch", "IntStorage"), ("torch", "BoolStorage"),
    ("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch.nn.parameter", "Parameter"),
    ("torch.nn.modules.activation", "LeakyReLU"), ("torch.nn.modules.activation", "ReLU"),
    ("torch.nn.modules.activation", "Sigmoid"), ("torch.nn.modules.container", "Sequential"),
    ("torch.nn.modules.conv", "Conv2d"), ("torch.nn.modules.pooling", "AdaptiveAvgPool2d"),
}


class _RemapUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        hit = _PICKLE_CLASSES.get((module, name))
        if hit is not None:
            return hit
        if (module, name) not in _PICKLE_GLOBALS:
            raise pickle.UnpicklingError(f"refusing to import {module}.{name} while loading a GAE checkpoint "
                                         "(not on the exact allow-list of hsi_dmgasr_b200.gae)")
        return super().find_class(module, name)


class _PickleModule:
    """``pickle_module`` for torch.load that maps the reference's ``__main__.*`` / ``common.*`` classes onto this file."""
    __name__ = "hsi_dmgasr_b200.gae"
    Unpickler = _RemapUnpickler
    load = staticmethod(lambda f, **kw: _RemapUnpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _RemapUnpickler(io.BytesIO(b), **kw).load())
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    Pickler = pickle.Pickler
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def save_reference_style_pickle(gae: "GAE", path: str) -> str:
    """Write `gae` as a whole-module pickle laid out like the reference's ``torch.save(model, 'GAE_4_*.pth')`` from
    ``AE.py`` run as a script (AE.py:637): classes named ``__main__.{GAE,Encoder,Decoder,BranchUnit,SSPN,SSB}`` and
    ``common.{ResBlock,ResAttentionBlock,CALayer,Upsampler}``.  Used to exercise the pickle loader on boxes that do not
    have the reference tree, and to hand a GAE back to reference-side tooling."""
    import sys
    import types
    remap = {cls: (mod, name) for (mod, name), cls in _PICKLE_CLASSES.items() if mod in ("__main__", "common")}
    saved_mod = {cls: cls.__module__ for cls in remap}
    saved_sys = {m: sys.modules.get(m) for m in ("common",)}
    main = sys.modules["__main__"]
    saved_main = {name: getattr(main, name, None) for cls, (mod, name) in remap.items() if mod == "__main__"}
    common = types.ModuleType("common")
    try:
        for cls, (mod, name) in remap.items():
            cls.__module__ = mod
            setattr(main if mod == "__main__" else common, name, cls)
        sys.modules["common"] = common
        torch.save(gae, path)
    finally:
        for cls, mod in saved_mod.items():
            cls.__module__ = mod
        for name, old in saved_main.items():
            if old is None:
                if hasattr(main, name):
                    delattr(main, name)
            else:
                setattr(main, name, old)
        for m, old in saved_sys.items():
            if old is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = old
    return path


GAE_STATE_FORMAT = "hsidm-gae-state-v1"


def save_gae_state(gae: GAE, path: str) -> str:
    """the GAE as a self-describing ``state_dict`` file (tensors + band-group geometry, no pickled
    classes), so it loads with ``weights_only=True`` and needs none of the reference's class names in ``__main__``.
    ``load_gae`` reads it back; the reference's whole-module pickles stay importable (import-only compatibility)."""
    g = gae.geometry()
    torch.save({"format": GAE_STATE_FORMAT,
                "geometry": {"n_colors": g.n_colors, "n_subs": g.n_subs, "n_ovls": g.n_ovls, "n_feats": g.n_feats},
                "state_dict": {k: v.detach().cpu() for k, v in gae.state_dict().items()}}, path)
    return path


def load_gae(path: str, map_location="cpu", precision: str = "fp32", allow_pickle: bool = True) -> GAE:
    """Load ``GAE_pretrained/GAE_4_*.pth`` (whole-module pickle, AE.py:637), a ``save_gae_state`` file, or a plain
    ``state_dict`` file.

    Tensors-only files load with ``weights_only=True``.  Only when that is refused because the file names classes
    (an ``UnpicklingError``; any other failure propagates) and ``allow_pickle`` is true does the remapping unpickler
    run; it resolves an exact allow-list of (module, name) pairs - the GAE classes of this module and the handful of
    ``torch`` / ``collections`` globals a module pickle needs - and refuses everything else, ``builtins`` included."""
    try:   # tensors-only files first: no unpickling of arbitrary globals
        obj = torch.load(path, map_location=map_location, weights_only=True)
    except pickle.UnpicklingError:
        if not allow_pickle:
            raise
        obj = torch.load(path, map_location=map_location, weights_only=False, pickle_module=_PickleModule)
    if isinstance(obj, dict) and obj.get("format") == GAE_STATE_FORMAT:
        geo = obj["geometry"]
        gae = GAE(n_subs=geo["n_subs"], n_ovls=geo["n_ovls"], n_colors=geo["n_colors"], n_feats=geo["n_feats"], precision=precision)
        gae.load_state_dict(obj["state_dict"], strict=True)
        return gae
    if isinstance(obj, dict):
        sd = obj
        n_subs, n_feats = sd["Encoder.branch.head.weight"].shape[1], sd["Encoder.branch.head.weight"].shape[0]
        n_colors = sd["final.weight"].shape[0]
        n_ovls = {8: 2, 16: 4}.get(n_subs, n_subs // 4)
        gae = GAE(n_subs=n_subs, n_ovls=n_ovls, n_colors=n_colors, n_feats=n_feats, precision=precision)
        gae.load_state_dict(sd, strict=True)
        return gae
    if not isinstance(obj, GAE):
        raise _lib.HsidmError(-6, f"{path} does not hold a GAE module or state_dict")
    obj.__dict__["precision"] = precision
    return obj
