"""Developer probe: a conv through the CTA-pair tile <1,256> against the single-CTA tile <2,128> on identical inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from hsi_dmgasr_b200 import _lib
from tests.gpu_util import conv2d, randn, rel_l2, tc_flag

lib = _lib.load()
bf = lambda t: t.to(torch.bfloat16).float()
for (n, c, h, w, co) in [(2, 256, 32, 32, 256), (4, 512, 16, 16, 512), (2, 128, 16, 16, 256), (6, 256, 32, 32, 512)]:
    x = randn((n, c, h, w), 1)
    wt = randn((co, c, 3, 3), 2, scale=(1.0 / (c * 9)) ** 0.5)
    b = randn((co,), 3)
    outs = {}
    for variant in (0, 32):
        lib.hsidm_debug_conv_mode(0, variant)
        outs[variant] = conv2d(1, "bf16", x, None, wt, b, ksize=3)
    lib.hsidm_debug_conv_mode(0, 0)
    want = F.conv2d(bf(x), bf(wt), b, padding=1)
    d = (outs[0] - outs[32]).abs()
    print(f"{(n, c, h, w, co)}: pair vs single max |diff| {float(d.max()):.3e} (differing {int((d > 0).sum())} of {d.numel()});"
          f" vs torch: pair {rel_l2(outs[0], want):.3e} single {rel_l2(outs[32], want):.3e} flag {tc_flag()}", flush=True)
