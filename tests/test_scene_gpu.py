"""Overlapping-tile scene driver on the GPU (BASELINE configs[2]): the blend kernel against its host
restatement, independence of the result from batch size, and - when the box has two GPUs - bit-equality of the scene
super-resolved on one GPU and sharded over two, including the NCCL gather of device tensors."""
import os
import socket

import numpy as np
import pytest
import torch

from hsi_dmgasr_b200 import GAE, GaussianDiffusion, SRPipeline, UNet, synth
from hsi_dmgasr_b200.pipeline import blend_tiles, super_resolve_scene, tile_scene
from hsi_dmgasr_b200.spec import GAEGeometry
from tests.cfgs import SMALL
from tests.host_ref import blend_tiles_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,h,w,tile,ov", [(5, 300, 260, 128, 16), (3, 128, 128, 128, 16), (2, 1096, 715, 128, 16), (4, 90, 75, 32, 0)])
def test_blend_kernel_bit_equal_to_host_reference(c, h, w, tile, ov):
    scene = torch.from_numpy(np.random.default_rng(5).random((c, h, w), dtype=np.float32))
    tiles, pos = tile_scene(scene, tile, ov)
    tiles = tiles + 0.01 * torch.from_numpy(np.random.default_rng(6).standard_normal(tuple(tiles.shape), dtype=np.float32))
    want = blend_tiles_ref(tiles, pos, h, w, ov)                    # on the CPU
    got = blend_tiles(tiles.cuda(), pos, h, w, ov)
    assert got.is_cuda and torch.equal(got.cpu(), want)
    back = blend_tiles(tile_scene(scene, tile, ov)[0].cuda(), pos, h, w, ov).cpu()
    assert torch.allclose(back, scene, atol=1e-6)                   # tile -> blend is the identity


def build(precision, device, T=4):
    geom = GAEGeometry(31, 8, 2)
    gae = GAE(n_subs=8, n_ovls=2, n_colors=31, n_feats=64)
    gae.load_state_dict(synth.gae_state_dict(geom, 81))
    net = UNet(in_channel=6, out_channel=3, inner_channel=SMALL.inner_channel, norm_groups=SMALL.norm_groups,
               channel_mults=SMALL.channel_mults, attn_res=SMALL.attn_res, res_blocks=SMALL.res_blocks, dropout=SMALL.dropout,
               image_size=SMALL.image_size, precision=precision)
    net.load_state_dict(synth.unet_state_dict(SMALL, 82))
    gd = GaussianDiffusion(net, image_size=16, channels=3, conditional=True).to(device).eval()
    gd.set_new_noise_schedule(dict(schedule="cosine", n_timestep=T, linear_start=1e-6, linear_end=1e-2), device)
    return SRPipeline(gd, gae.to(device).eval())


def scene_input():
    return synth.sr_cube(1, 31, 96, seed=83)[0][:, :80, :72].contiguous()       # 3 x 3 tiles of 32 with overlap 8


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_scene_result_does_not_depend_on_batching(precision):
    dev = torch.device("cuda")
    pipe = build(precision, dev)
    scene = scene_input()
    a = super_resolve_scene(pipe, scene, dev, tile=32, overlap=8, batch=9, seed=11)
    b = super_resolve_scene(pipe, scene, dev, tile=32, overlap=8, batch=2, seed=11)
    c = super_resolve_scene(pipe, scene, dev, tile=32, overlap=8, batch=4, seed=12)
    assert a.shape == scene.shape and a.is_cuda and torch.isfinite(a).all()
    assert torch.equal(a, b)                       # same seed: every tile sees the same noise whatever batch it is in
    assert not torch.equal(a, c)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, precision, path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        pipe = build(precision, dev)
        out = super_resolve_scene(pipe, scene_input(), dev, tile=32, overlap=8, batch=3, rank=rank, world=world, seed=11)
        if rank == 0:
            torch.save(out.cpu(), path)
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_gpu_scene_bit_equal_to_one_gpu(tmp_path, precision):
    import torch.multiprocessing as mp
    dev = torch.device("cuda", 0)
    one = super_resolve_scene(build(precision, dev), scene_input(), dev, tile=32, overlap=8, batch=9, seed=11).cpu()
    path = str(tmp_path / "two.pt")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    two = torch.load(path)
    assert torch.equal(one, two)
