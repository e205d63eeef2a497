"""SURVEY 8(d) step-ms protocol: CUDA events around one CUDA-graph replay of (UNet forward + posterior step) at
N in {1, 5, 11, 176} latent images of 128x128; median of 50 replays after warm-up.  Prints one JSON line per N."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth, _lib
if os.environ.get("HSIDM_AB_LIB"):   # A/B against another build of the library (developer runs only)
    _lib.LIB_PATH = os.environ["HSIDM_AB_LIB"]
from hsi_dmgasr_b200.spec import UNetConfig

FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
dev = torch.device("cuda:0")
T = 64   # one sampling pass = T graph replays; per-step time = pass time / T, median over passes
for n in [int(v) for v in os.environ.get("STEP_LAT_N", "1,5,11,176").split(",")]:
    net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
               res_blocks=2, dropout=0.2, image_size=128, precision="bf16")
    net.load_state_dict(synth.unet_state_dict(FULL, 0))
    gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), dev)
    cond = torch.randn(n, 3, 128, 128, device=dev)
    passes = 3 if n == 176 else 8
    times = []
    for i in range(passes + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gd.p_sample_loop(cond, False, return_all=True, seed=i)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1) / T)
    ms = statistics.median(times)
    print(json.dumps({"latents": n, "ms_per_step": round(ms, 4), "tflops": round(92.353 * n / ms, 1),
                      "frac_of_sustained_bf16_peak": round(92.353 * n / ms / 1367.4, 3), "passes": passes, "steps_per_pass": T}),
          flush=True)
    del gd, net
    torch.cuda.empty_cache()
