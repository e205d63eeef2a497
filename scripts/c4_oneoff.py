"""Developer probe: where a short C4 pass (5 latents of 512x512) spends its one-off time after a schedule change."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth
from hsi_dmgasr_b200.spec import UNetConfig

CFG = dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_mults=(1, 2, 4, 8, 16), attn_res=(), res_blocks=1,
           dropout=0.2, image_size=128)
HW, N = int(os.environ.get("C4_HW", "512")), int(os.environ.get("C4_N", "5"))
dev = torch.device("cuda:0")
net = UNet(**{**CFG, "attn_res": []}, precision="bf16")
net.load_state_dict(synth.unet_state_dict(UNetConfig(**CFG), 0))
gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
cond = torch.randn(N, 3, HW, HW, device=dev)


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print(f"{label:44s} {1e3 * (time.perf_counter() - t0):9.2f} ms", flush=True)
    return r


for T in (3, 20, 20, 21, 40, 20):
    timed(f"set_new_noise_schedule(T={T})", lambda: gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), dev))
    timed(f"  super_resolution, T={T}, return_all", lambda: gd.super_resolution(cond, return_all=True, seed=1))
    timed(f"  super_resolution again", lambda: gd.super_resolution(cond, return_all=True, seed=2))
    timed(f"  super_resolution, return_all off", lambda: gd.super_resolution(cond, seed=2))
