"""Developer probe: does running two half-batches on two streams overlap the HBM-bound and tensor-bound phases?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth
from hsi_dmgasr_b200.spec import UNetConfig

FULL = UNetConfig(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                  attn_res=(16,), res_blocks=2, dropout=0.2, image_size=128)
dev = torch.device("cuda:0")


def make():
    net = UNet(in_channel=6, out_channel=3, inner_channel=64, norm_groups=32, channel_mults=(1, 2, 4, 8, 8), attn_res=[16],
               res_blocks=2, dropout=0.2, image_size=128, precision="bf16")
    net.load_state_dict(synth.unet_state_dict(FULL, 0))
    gd = GaussianDiffusion(net, image_size=128, channels=3).to(dev).eval()
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=2000, linear_start=1e-6, linear_end=1e-2), dev)
    return gd


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


n = int(sys.argv[1]) if len(sys.argv) > 1 else 176
a, b = make(), make()
cond = torch.randn(n, 3, 128, 128, device=dev)
x = torch.randn_like(cond)
h = n // 2
print(f"one stream, N={n}: {timeit(lambda: a.predict_noise(x, 1000, cond)):.3f} ms")
print(f"one stream, N={h} x2 sequential: {timeit(lambda: (a.predict_noise(x[:h], 1000, cond[:h]), a.predict_noise(x[h:], 1000, cond[h:]))):.3f} ms")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
c1, c2, x1, x2 = cond[:h].contiguous(), cond[h:].contiguous(), x[:h].contiguous(), x[h:].contiguous()


def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur), s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        a.predict_noise(x1, 1000, c1)
    with torch.cuda.stream(s2):
        b.predict_noise(x2, 1000, c2)
    cur.wait_stream(s1), cur.wait_stream(s2)


print(f"two streams, N={h} each: {timeit(two):.3f} ms")
