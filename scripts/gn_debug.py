"""Developer probe: one eager bf16 UNet forward at --n images with per-launch times and the barrier-timeout code."""
import argparse, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, _lib, synth
from hsi_dmgasr_b200.spec import UNetConfig
from tests.gpu_util import tc_flag

FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=12)
a = ap.parse_args()
dev = torch.device("cuda:0")
net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
           res_blocks=2, dropout=0.2, image_size=128, precision="bf16")
net.load_state_dict(synth.unet_state_dict(FULL, 0))
gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=2000, linear_start=1e-6, linear_end=1e-2), dev)
cond = torch.randn(a.n, 3, 128, 128, device=dev)
x = torch.randn_like(cond)
lib = _lib.load()
lib.hsidm_prof_enable(1)
gd.predict_noise(x, 1000, cond)
torch.cuda.synchronize()
print("flag", tc_flag(), flush=True)
out = "gpurun_out/gn_debug.csv"
os.makedirs("gpurun_out", exist_ok=True)
_lib.check(lib.hsidm_prof_dump(out.encode()))
for i, r in enumerate(csv.DictReader(open(out))):
    if float(r["ms"]) > 5.0:
        print(i, r["tag"], r["ms"], flush=True)
