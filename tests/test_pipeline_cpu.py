"""Host-side logic of the sharded pipeline and the tile driver, on CPU (gloo, world_size 2)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hsi_dmgasr_b200.pipeline import feather_window, run_sharded, shard_bounds, super_resolve_scene, tile_scene, tile_starts
from tests.host_ref import blend_tiles_ref as blend_tiles     # the package blends on the GPU only; this is the checker


class FakePipeline:
    """Stands in for SRPipeline on a CPU box: the 'SR' of a cube is 2*cube + band index (checks routing, not arithmetic)."""

    def super_resolve_host(self, cubes, device, **kw):
        bands = torch.arange(cubes.shape[1], dtype=cubes.dtype).view(1, -1, 1, 1)
        return cubes * 2 + bands


def test_tile_starts_and_pavia_tile_count():
    assert tile_starts(128, 128, 16) == [0]
    assert tile_starts(240, 128, 16) == [0, 112]
    assert tile_starts(250, 128, 16) == [0, 112, 122]
    ys, xs = tile_starts(1096, 128, 16), tile_starts(715, 128, 16)
    assert (len(ys), len(xs), len(ys) * len(xs)) == (10, 7, 70)          # 70 tiles for Pavia Centre
    assert ys[-1] == 1096 - 128 and xs[-1] == 715 - 128
    with pytest.raises(ValueError):
        tile_starts(100, 128, 16)


def test_tile_then_blend_is_identity():
    scene = torch.rand(5, 300, 260)
    tiles, pos = tile_scene(scene, 128, 16)
    assert tiles.shape == (len(pos), 5, 128, 128)
    back = blend_tiles(tiles, pos, 300, 260, 16)
    assert torch.allclose(back, scene, atol=1e-6)
    w = feather_window(128, 16)
    assert float(w.min()) > 0 and float(w.max()) == 1.0 and torch.equal(w, w.t())


def test_scene_driver_single_rank():
    scene = torch.rand(4, 200, 150)
    out = super_resolve_scene(FakePipeline(), scene, torch.device("cpu"), batch=3, blend_fn=blend_tiles, keep_on_device=False)
    want = scene * 2 + torch.arange(4.0).view(-1, 1, 1)
    assert torch.allclose(out, want, atol=1e-5)


def test_package_blend_has_no_host_path():
    from hsi_dmgasr_b200 import _lib, pipeline
    tiles, pos = tile_scene(torch.rand(2, 140, 140), 128, 16)
    with pytest.raises(_lib.HsidmError):
        pipeline.blend_tiles(tiles, pos, 140, 140, 16)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        cubes = torch.rand(7, 3, 8, 8)                       # 7 work items over 2 ranks -> 4 + 3
        lo, hi = shard_bounds(7, rank, world)
        got = run_sharded(FakePipeline(), cubes, torch.device("cpu"), rank, world, batch=2, gather=True)
        scene = torch.rand(2, 240, 250)
        blended = super_resolve_scene(FakePipeline(), scene, torch.device("cpu"), batch=2, rank=rank, world=world,
                                      blend_fn=blend_tiles, keep_on_device=False)
        if rank == 0:
            want = cubes * 2 + torch.arange(3.0).view(1, -1, 1, 1)
            ret["cubes_ok"] = bool(torch.allclose(got, want))
            ret["scene_ok"] = bool(torch.allclose(blended, scene * 2 + torch.arange(2.0).view(-1, 1, 1), atol=1e-5))
        else:
            ret["other_none"] = got is None and blended is None
        ret[f"slice{rank}"] = (lo, hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_gather_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret["cubes_ok"] and ret["scene_ok"] and ret["other_none"]
    assert ret["slice0"] == (0, 4) and ret["slice1"] == (4, 7)


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hsi_dmgasr_b200.diffusion import allreduce_gradients

        class Holder:
            pass
        h = Holder()
        h.last_grad_slab = torch.arange(6, dtype=torch.float32) * (rank + 1)
        view = h.last_grad_slab[2:4]                       # a parameter's .grad is a view of the slab
        allreduce_gradients(h, world, dist)
        ret[f"slab{rank}"] = h.last_grad_slab.tolist()
        ret[f"view{rank}"] = view.tolist()
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_averages_the_slab_gloo():
    """Data-parallel training (BASELINE configs[4]): one all-reduce of the contiguous gradient slab, then 1/world; the
    per-parameter .grad views see the averaged values."""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = [1.5 * i for i in range(6)]                     # mean of i*1 and i*2
    assert ret["slab0"] == want and ret["slab1"] == want and ret["view0"] == want[2:4]
