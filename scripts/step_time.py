"""Times one UNet denoise step (forward + posterior) on the GPU for a few batch sizes. Developer tool."""
import argparse
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth
from hsi_dmgasr_b200.spec import UNetConfig

FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batches", default="1,5,11,44")
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
               res_blocks=2, dropout=0.2, image_size=128, precision=a.precision)
    net.load_state_dict(synth.unet_state_dict(FULL, 0))
    gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=2000, linear_start=1e-6, linear_end=1e-2), dev)
    for n in [int(b) for b in a.batches.split(",")]:
        cond = torch.randn(n, 3, a.hw, a.hw, device=dev)
        x = torch.randn_like(cond)
        for _ in range(2):
            gd.predict_noise(x, 1000, cond)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            gd.predict_noise(x, 1000, cond)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        tf = 92.353e9 * n * (a.hw / 128) ** 2 / (ms * 1e-3) / 1e12
        print(f"{a.precision} N={n:4d} {a.hw}x{a.hw}: {ms:9.3f} ms/step  {tf:8.1f} TFLOP/s  ({tf / 1367.4 * 100:5.1f}% of sustained bf16 peak)",
              flush=True)


if __name__ == "__main__":
    main()
