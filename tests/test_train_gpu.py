"""Training step on the GPU (reference GaussianDiffusion.p_losses diffusion.py:222-250 and
DDPM.optimize_parameters model.py:49-59): loss and all 124 parameter gradients of the hand-written CUDA forward+backward
against the UNMODIFIED reference's autograd (tests/golden/train_step.npz, oracle/make_golden_train.py), the drop-in
optimiser step, dropout statistics, determinism."""
import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import GaussianDiffusion, UNet, synth
from hsi_dmgasr_b200.spec import UNetConfig
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu

TRAIN = UNetConfig(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2), attn_res=(8,),
                   res_blocks=1, dropout=0.0, image_size=16)
SEED, T, B, HW, STEP_T = 51, 20, 3, 16, 7          # oracle/make_golden_train.py


def rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape, dtype=np.float32))


def build(loss_type, cfg=TRAIN, precision="fp32"):
    net = UNet(in_channel=6, out_channel=3, inner_channel=cfg.inner_channel, norm_groups=cfg.norm_groups,
               channel_mults=cfg.channel_mults, attn_res=cfg.attn_res, res_blocks=cfg.res_blocks, dropout=cfg.dropout,
               image_size=cfg.image_size, precision=precision)
    net.load_state_dict(synth.unet_state_dict(cfg, SEED))
    gd = GaussianDiffusion(net, image_size=HW, channels=3, loss_type=loss_type, conditional=True).cuda().train()
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), torch.device("cuda"))
    gd.set_loss(torch.device("cuda"))
    return gd


class injected_numpy_draws:
    """The reference draws t and the noise levels from numpy's global RNG (diffusion.py:226-233); inject them like the
    golden generator does."""

    def __init__(self, levels):
        self.levels = np.asarray(levels, dtype=np.float64)

    def __enter__(self):
        self.saved = (np.random.randint, np.random.uniform)
        np.random.randint = lambda a, b=None, *k, **kw: STEP_T
        np.random.uniform = lambda a, b, size=None: self.levels.copy()

    def __exit__(self, *exc):
        np.random.randint, np.random.uniform = self.saved


@pytest.mark.parametrize("loss_type", ["l1", "l2"])
def test_loss_and_gradients_match_reference(golden, loss_type):
    g = golden("train_step.npz")
    gd = build(loss_type)
    hr, sr, noise = rand((B, 3, HW, HW), 61).cuda(), rand((B, 3, HW, HW), 62).cuda(), rand((B, 3, HW, HW), 63).cuda()
    with injected_numpy_draws(g[f"{loss_type}.levels"]):
        l_pix = gd({"HR": hr, "SR": sr}, noise=noise)                    # model.py:51
    loss = l_pix.sum() / int(B * 3 * HW * HW)                              # model.py:53-55
    loss.backward()
    want_sum, want = float(g[f"{loss_type}.loss_sum"]), float(g[f"{loss_type}.loss"])
    assert abs(float(l_pix.detach()) - want_sum) < 1e-5 * abs(want_sum) and abs(float(loss.detach()) - want) < 1e-5
    worst, checked, n = 0.0, 0, 0
    for k, p in gd.denoise_fn.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, k
        want_norm = float(g[f"{loss_type}.gnorm.{k}"])
        got_norm = float(p.grad.double().norm())
        assert abs(got_norm - want_norm) <= 2e-4 * want_norm + 1e-9, (k, got_norm, want_norm)
        if f"{loss_type}.grad.{k}" in g.files:
            e = rel_l2(p.grad, torch.from_numpy(g[f"{loss_type}.grad.{k}"]))
            worst = max(worst, e)
            assert e < 2e-4, (k, e)
            checked += 1
        n += 1
    print(f"train step {loss_type}: loss {float(loss):.7f} (reference {want:.7f}), {n} gradients, worst rel-L2 of the {checked} "
          f"stored in full {worst:.2e}")
    assert n == 124 and checked > 50


def test_step_is_deterministic_and_optimizer_is_drop_in(golden):
    """Two steps from the same state give bit-identical gradients (no atomics); DDPM.optimize_parameters (model.py:49-59) with
    torch.optim.Adam moves the weights exactly like Adam applied to the reference gradients, and the next forward sees the
    updated weights."""
    from hsi_dmgasr_b200.config import dict_to_nonedict
    from hsi_dmgasr_b200.model import DDPM
    g = golden("train_step.npz")
    unet = dict(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_multiplier=[1, 2], attn_res=[8], res_blocks=1,
                dropout=0.0, precision="fp32")
    sched = dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2)
    opt = dict_to_nonedict(dict(phase="train", gpu_ids=[0], distributed=False, path=dict(resume_state=None),
                                train=dict(optimizer=dict(type="adam", lr=1e-3)),
                                model=dict(which_model_G="sr3", finetune_norm=False, unet=unet, beta_schedule=dict(train=sched, val=sched),
                                           diffusion=dict(image_size=16, channels=3, conditional=True))))
    m = DDPM(opt)
    m.netG.denoise_fn.load_state_dict(synth.unet_state_dict(TRAIN, SEED))
    before = {k: v.detach().clone() for k, v in m.netG.denoise_fn.named_parameters()}
    data = {"HR": rand((B, 3, HW, HW), 61), "SR": rand((B, 3, HW, HW), 62)}
    noise = rand((B, 3, HW, HW), 63).cuda()
    m.feed_data(data)
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: noise.clone()
    try:
        with injected_numpy_draws(g["l1.levels"]):
            m.optimize_parameters()
            first = float(m.log_dict["l_pix"])
            grads1 = {k: p.grad.detach().clone() for k, p in m.netG.denoise_fn.named_parameters()}
            # Adam's first step moves every weight by lr * sign(grad) (up to eps): check against the reference gradient signs
            k0 = "final_conv.block.3.bias"
            step = (dict(m.netG.denoise_fn.named_parameters())[k0].detach() - before[k0]).cpu()
            want = -1e-3 * torch.sign(torch.from_numpy(g[f"l1.grad.{k0}"]))
            assert torch.allclose(step, want, atol=2e-6)
            m.optimize_parameters()
            second = float(m.log_dict["l_pix"])
    finally:
        torch.randn_like = orig
    assert abs(first - float(g["l1.loss"])) < 1e-5
    assert second < first                      # same batch, one Adam step later: the loss went down
    # determinism: rebuild the same state and repeat the first step
    m2 = DDPM(opt)
    m2.netG.denoise_fn.load_state_dict(synth.unet_state_dict(TRAIN, SEED))
    m2.feed_data(data)
    torch.randn_like = lambda x, *a, **k: noise.clone()
    try:
        with injected_numpy_draws(g["l1.levels"]):
            m2.optimize_parameters()
    finally:
        torch.randn_like = orig
    # grads of m2 are those of its (single) step; m's .grad now holds the SECOND step: compare with the saved first ones
    for k, p in m2.netG.denoise_fn.named_parameters():
        assert torch.equal(p.grad, grads1[k]), k


def test_dropout_mask_statistics_and_regeneration():
    """dropout 0.2 (config/sr_sr3_16_128ae.json): the counter-based mask keeps ~80 % and is regenerated identically by the
    backward - with all-ones upstream the input gradient of a dropped element is zero - checked through finite differences of
    the loss on one parameter."""
    cfg = UNetConfig(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2), attn_res=(8,),
                     res_blocks=1, dropout=0.2, image_size=16)
    gd = build("l2", cfg)
    hr, sr, noise = rand((2, 3, HW, HW), 71).cuda(), rand((2, 3, HW, HW), 72).cuda(), rand((2, 3, HW, HW), 73).cuda()
    p = gd.denoise_fn.final_conv["block"]["3"].bias

    def loss_at(delta, seed):
        torch.manual_seed(seed)
        np.random.seed(5)
        with torch.no_grad():
            p.add_(delta)
        out = gd({"HR": hr, "SR": sr}, noise=noise)
        with torch.no_grad():
            p.sub_(delta)
        return out

    base = loss_at(0.0, 9)
    base.backward()
    grad = p.grad.detach().clone()
    eps = 1e-2
    d = torch.zeros_like(p)
    d[1] = eps
    fd = (float(loss_at(d, 9)) - float(loss_at(-d, 9))) / (2 * eps)
    assert abs(fd - float(grad[1])) < 2e-3 * max(1.0, abs(fd)), (fd, float(grad[1]))
    # a different dropout seed gives a different loss, the same seed the same loss
    assert float(loss_at(0.0, 9)) == float(base) and float(loss_at(0.0, 10)) != float(base)


def _dp_worker(rank, world, port, path):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hsi_dmgasr_b200.diffusion import allreduce_gradients
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_step.npz"))
        with torch.cuda.device(rank):
            net = UNet(in_channel=6, out_channel=3, inner_channel=32, norm_groups=8, channel_mults=(1, 2), attn_res=(8,), res_blocks=1,
                       dropout=0.0, image_size=16, precision="fp32")
            net.load_state_dict(synth.unet_state_dict(TRAIN, SEED))
            dev = torch.device("cuda", rank)
            gd = GaussianDiffusion(net, image_size=HW, channels=3, loss_type="l1", conditional=True).to(dev).train()
            gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), dev)
            gd.set_loss(dev)
            hr, sr, noise = rand((B, 3, HW, HW), 61), rand((B, 3, HW, HW), 62), rand((B, 3, HW, HW), 63)
            # rank r takes sample r of the golden batch (the third one is dropped): the averaged gradient must equal the gradient
            # of the two-sample batch normalised by ITS element count
            sl = slice(rank, rank + 1)
            with injected_numpy_draws(g["l1.levels"][sl]):
                l_pix = gd({"HR": hr[sl].to(dev), "SR": sr[sl].to(dev)}, noise=noise[sl].to(dev))
            (l_pix.sum() / int(1 * 3 * HW * HW)).backward()
            allreduce_gradients(gd, world)
            if rank == 0:
                torch.save({k: p.grad.cpu() for k, p in gd.denoise_fn.named_parameters()}, path)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_data_parallel_gradients_equal_the_joint_batch(tmp_path, golden):
    """BASELINE configs[4]: one process per GPU, NCCL all-reduce of the gradient slab.  Two ranks with one sample each
    give the gradient of the joint two-sample batch computed on one GPU."""
    import socket
    import torch.multiprocessing as mp
    g = golden("train_step.npz")
    gd = build("l1")
    hr, sr, noise = rand((B, 3, HW, HW), 61).cuda(), rand((B, 3, HW, HW), 62).cuda(), rand((B, 3, HW, HW), 63).cuda()
    with injected_numpy_draws(g["l1.levels"][:2]):
        l_pix = gd({"HR": hr[:2], "SR": sr[:2]}, noise=noise[:2])
    (l_pix.sum() / int(2 * 3 * HW * HW)).backward()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = str(tmp_path / "dp.pt")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    dp = torch.load(path)
    worst = max(rel_l2(dp[k], p.grad) for k, p in gd.denoise_fn.named_parameters())
    print(f"2-rank data-parallel vs joint batch: worst gradient rel-L2 {worst:.2e}")
    assert worst < 1e-5
