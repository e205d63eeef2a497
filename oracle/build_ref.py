"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, compiled where they lie.

TEST / BASELINE INFRASTRUCTURE.  The reference is pure Python, so "compiling" it means byte-compiling
``model/sr3_modules/unet.py`` and ``model/sr3_modules/diffusion.py`` from /root/reference into sourceless bytecode files
(``*.bc``: the ``.pyc`` format under another extension, so that tools which skip ``*.pyc`` still ship them) under ``oracle/_ref/`` (git-ignored, so no reference source enters the history; not gpurun-ignored, so the files travel to
the GPU box, which has this same interpreter but no /root/reference).  ``bench.py --impl reference`` and the
``cpu_baseline`` leg load them (``load()`` below) and time the reference's own ``UNet.forward`` /
``GaussianDiffusion.super_resolution`` on the host cores (``kind: "reference"``); when the directory is missing they fall
back to the oracle port (``kind: "port"``).  Nothing in the product package imports this file.

AE.py (the GAE codec) is not compiled: it imports training-only modules absent from this image (sewar, GELIN, HStest ...)
and its encode/decode loops hard-code 'cuda:0' (AE.py:283-324); the codec share of the CPU baseline (0.01 % of a patch)
stays with the oracle port, which is pinned to AE.py's outputs by tests/golden/gae*.npz.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULES = {"unet": "model/sr3_modules/unet.py", "diffusion": "model/sr3_modules/diffusion.py"}


def build(reference: str = "/root/reference") -> list:
    """Byte-compile the reference modules into oracle/_ref/.  Returns the files written ([] when the tree is absent)."""
    if not os.path.isdir(reference):
        return []
    os.makedirs(OUT, exist_ok=True)
    written = []
    for name, rel in MODULES.items():
        src = os.path.join(reference, rel)
        dst = os.path.join(OUT, f"{name}.bc")
        # unchecked-hash pyc: valid without the source file next to it; dfile keeps reference paths in tracebacks
        py_compile.compile(src, cfile=dst, dfile=f"<reference>/{rel}", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        written.append(dst)
    return written


def load():
    """(unet module, diffusion module) of the unmodified reference from oracle/_ref/, or None when it was not built (or was
    built by another interpreter version)."""
    mods = []
    for name in MODULES:
        path = os.path.join(OUT, f"{name}.bc")
        if not os.path.isfile(path):
            return None
        try:
            qual = f"hsidm_reference_{name}"
            loader = importlib.machinery.SourcelessFileLoader(qual, path)
            spec = importlib.util.spec_from_loader(qual, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
        except (ImportError, EOFError, ValueError):
            return None
        mods.append(mod)
    return tuple(mods)


if __name__ == "__main__":
    print("\n".join(build(*sys.argv[1:])) or "reference tree not found: nothing built")
