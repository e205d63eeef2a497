"""Developer probe: CUDA-graph sampling latency of the C4 workload (64_512 UNet, 5 latents of 512x512), optionally with
another build of the library (HSIDM_AB_LIB; symbols that build lacks are dropped from the ctypes table first)."""
import ctypes, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth, _lib
if os.environ.get("HSIDM_AB_LIB"):
    _lib.LIB_PATH = os.environ["HSIDM_AB_LIB"]
    probe = ctypes.CDLL(_lib.LIB_PATH)
    for k in list(_lib.SIGNATURES):
        if not hasattr(probe, k):
            del _lib.SIGNATURES[k]
from hsi_dmgasr_b200.spec import UNetConfig

CFG = dict(in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_mults=(1, 2, 4, 8, 16), attn_res=(), res_blocks=1,
           dropout=0.2, image_size=128)
HW = int(os.environ.get("C4_HW", "512"))
N = int(os.environ.get("C4_N", "5"))
T = int(os.environ.get("C4_T", "16"))
dev = torch.device("cuda:0")
net = UNet(**{**CFG, "attn_res": []}, precision="bf16")
net.load_state_dict(synth.unet_state_dict(UNetConfig(**CFG), 0))
gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), dev)
cond = torch.randn(N, 3, HW, HW, device=dev)
times = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gd.p_sample_loop(cond, False, return_all=bool(int(os.environ.get("C4_RETURN_ALL", "1"))), seed=i)
    e1.record()
    torch.cuda.synchronize()
    times.append(round(e0.elapsed_time(e1) / T, 3))
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "latents": N, "hw": HW, "T": T, "ms_per_step_each_pass": times,
                  "median_after_warmup": statistics.median(times[2:])}), flush=True)
