"""Developer probe: halo-kernel correctness under both descriptor base-offset modes, printed per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from hsi_dmgasr_b200 import _lib
from tests.gpu_util import conv2d, randn, rel_l2, tc_flag

torch.backends.cudnn.allow_tf32 = False
lib = _lib.load()
bf = lambda t: t.to(torch.bfloat16).float()
cases = [(1, 64, 16, 32, 64), (2, 64, 32, 32, 64), (1, 64, 16, 16, 128), (2, 128, 32, 32, 256), (1, 64, 128, 128, 64),
         (1, 192, 48, 64, 64), (2, 64, 32, 32, 3)]
for mode in (0, 1):
    for (n, c, h, w, co) in cases:
        lib.hsidm_debug_conv_mode(0, mode)
        x = randn((n, c, h, w), 1)
        wt = randn((co, c, 3, 3), 2, scale=(1.0 / (c * 9)) ** 0.5)
        b = randn((co,), 3)
        try:
            got = conv2d(1, "bf16", x, None, wt, b, ksize=3, out_nchw=(co == 3))
            flag = tc_flag()
            want = F.conv2d(bf(x), bf(wt), b, padding=1)
            print(f"mode {mode} case {(n, c, h, w, co)}: rel {rel_l2(got, want):.3e} flag {flag}", flush=True)
        except Exception as e:
            print(f"mode {mode} case {(n, c, h, w, co)}: EXC {e}", flush=True)
            sys.exit(1)
lib.hsidm_debug_conv_mode(0, 0)
