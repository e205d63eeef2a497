"""Developer probe: bitwise repeatability of one eager bf16 UNet forward and of the graph sampler at --n latents."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, _lib, synth
from hsi_dmgasr_b200.spec import UNetConfig
from tests.gpu_util import tc_flag

FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=176)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--no-halo", type=int, default=0)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
lib = _lib.load()
lib.hsidm_debug_conv_mode(a.no_halo, a.variant)
dev = torch.device("cuda:0")
net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
           res_blocks=2, dropout=0.2, image_size=128, precision="bf16")
net.load_state_dict(synth.unet_state_dict(FULL, 0))
gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=6, linear_start=1e-6, linear_end=1e-2), dev)
g = torch.Generator(device="cpu").manual_seed(1)
cond = torch.randn(a.n, 3, 128, 128, generator=g).to(dev)
x = torch.randn(a.n, 3, 128, 128, generator=g).to(dev)
outs = [gd.predict_noise(x, 3, cond).clone() for _ in range(a.reps)]
torch.cuda.synchronize()
for i in range(1, a.reps):
    d = (outs[i] - outs[0]).abs()
    bad = (d.flatten(1).max(dim=1).values > 0).nonzero().flatten().tolist()
    print(f"forward rep {i}: max |diff| {float(d.max()):.3e}  images differing: {bad[:12]}{'...' if len(bad) > 12 else ''} ({len(bad)})", flush=True)
print("flag", tc_flag(), flush=True)
s = [gd.p_sample_loop(cond, False, return_all=True, seed=9).clone() for _ in range(3)]
for i in range(1, 3):
    d = (s[i] - s[0]).abs()
    bad = (d.flatten(1).max(dim=1).values > 0).nonzero().flatten().tolist()
    print(f"sampler rep {i}: max |diff| {float(d.max()):.3e}  images differing: {bad[:12]} ({len(bad)})", flush=True)
print("flag", tc_flag(), flush=True)
