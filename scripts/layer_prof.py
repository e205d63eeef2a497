"""Developer tool: per-launch CUDA-event times of one eager UNet step (hsidm_prof_dump), aggregated per layer shape."""
import argparse, collections, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, _lib, synth
from hsi_dmgasr_b200.spec import UNetConfig

FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=176)
ap.add_argument("--out", default="gpurun_out/layer_prof.csv")
ap.add_argument("--variant", type=int, default=0, help="hsidm_debug_conv_mode variant bits")
ap.add_argument("--top", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda:0")
net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
           res_blocks=2, dropout=0.2, image_size=128, precision="bf16")
net.load_state_dict(synth.unet_state_dict(FULL, 0))
gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=2000, linear_start=1e-6, linear_end=1e-2), dev)
_lib.load().hsidm_debug_conv_mode(0, a.variant)
cond = torch.randn(a.n, 3, 128, 128, device=dev)
x = torch.randn_like(cond)
for _ in range(2):
    gd.predict_noise(x, 1000, cond)
lib = _lib.load()
lib.hsidm_prof_enable(1)
gd.predict_noise(x, 1000, cond)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
_lib.check(lib.hsidm_prof_dump(a.out.encode()))
lib.hsidm_prof_enable(0)
agg = collections.OrderedDict()
for r in csv.DictReader(open(a.out)):
    k = (r["kind"], r["tag"])
    v = agg.setdefault(k, [0, 0.0, 0.0])
    v[0] += 1; v[1] += float(r["ms"]); v[2] += float(r["work"])
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.3f} ms")
for (kind, tag), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    rate = v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0
    unit = "TFLOP/s" if kind in ("0", "1", "4") else "TB/s"
    print(f"kind {kind} {tag:55s} x{v[0]:3d} {v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%  {rate:8.1f} {unit}")
