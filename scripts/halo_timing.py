"""Developer probe: where the halo kernel's warp roles wait (cycle counters from hsidm_debug_halo_timing).

Runs single 3x3 convs of the shapes the 176-latent denoise step uses and prints, per shape, the wall time and the mean
per-CTA share of the MMA issuer's loop spent waiting on each barrier.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hsi_dmgasr_b200 import _lib
if os.environ.get("HSIDM_AB_LIB"):   # A/B against another build of the library (developer runs only)
    _lib.LIB_PATH = os.environ["HSIDM_AB_LIB"]
from tests.gpu_util import conv2d, randn, tc_flag

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=176)
ap.add_argument("--variant", type=int, default=0)
args = ap.parse_args()
lib = _lib.load()
lib.hsidm_debug_conv_mode(0, args.variant)
# (cin, H, W, cout) at n images
only_last = os.environ.get("HALO_TIMING_LAST")   # just the network's last conv (64 -> 3, fp32 NCHW output, <4,16> tile)
cases = [(64, 128, 128, 3)] if only_last else [(64, 128, 128, 64), (128, 128, 128, 64), (128, 64, 64, 128), (256, 64, 64, 128), (256, 32, 32, 256), (512, 32, 32, 256),
         (512, 16, 16, 512), (1024, 16, 16, 512), (512, 8, 8, 512)]
dbg = torch.zeros(8 * 148, dtype=torch.int64, device="cuda")
names = ["a_empty", "b_empty", "t_empty", "a_full", "b_full", "mma_total", "t_full", "epi_total"]
for (c, h, w, co) in cases:
    n = args.n
    x = randn((n, c, h, w), 1)
    wt = randn((co, c, 3, 3), 2, scale=(1.0 / (c * 9)) ** 0.5)
    b = randn((co,), 3)
    # conv2d converts layouts on every call; time only the kernel through the profiler API
    lib.hsidm_debug_halo_timing(dbg.data_ptr())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(3):
        conv2d(1, "bf16", x, None, wt, b, ksize=3, out_nchw=co < 16)
    torch.cuda.synchronize()
    lib.hsidm_debug_halo_timing(None)
    d = dbg.cpu().double().view(148, 8)
    tot = d[:, 5].clamp_min(1)
    flops = 2.0 * n * h * w * co * 9 * c
    line = " ".join(f"{names[i]} {100 * float((d[:, i] / tot).mean()):5.1f}%" for i in (0, 1, 2, 3, 4, 6))
    line += f"  epi_total {float(d[:, 7].mean()) / 1e3:.1f} kclk"
    mma_clk = float(tot.mean())
    mmas = n * h * w / 128 * max(1.0, co / (128 if co % 128 == 0 else 64)) * 9 * c / 16 / 148   # MMAs per CTA
    print(f"cin{c} {h}x{w} cout{co} n{n}: issuer loop {mma_clk / 1e3:8.1f} kclk = {mma_clk / mmas:6.1f} clk/MMA  waits: {line}  flag {tc_flag()}", flush=True)
    dbg.zero_()
