"""Batched restatement of the reference's validation loop (sr_gae.py:436-475) and its multi-GPU sharding.

Reference: per cube -> ``encode`` -> for each of the G groups, sequentially at batch 1: ``feed_data`` / ``test`` /
``get_current_visuals`` (a D2H + H2D round trip per group) -> ``decode`` -> clamp to [0,1].
Here: all B*G latent images of a batch of cubes are denoised together (the UNet weights are shared by every group),
nothing leaves the device between encode and decode, and cubes are sharded across GPUs with no per-step collective
(SURVEY.md 8e): one process per GPU, a contiguous slice of the work list each, one gather at the very end.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib
from .diffusion import GaussianDiffusion
from .gae import GAE


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of a work list for `rank` of `world` (first n%world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class SRPipeline:
    """encode -> T-step conditional sampling of every group latent -> decode, for a batch of cubes on one GPU."""

    def __init__(self, diffusion: GaussianDiffusion, gae: GAE, max_latents: Optional[int] = None):
        self.diffusion = diffusion
        self.gae = gae
        self.max_latents = max_latents      # cap on latent images per sampling call (None = all B*G at once)

    @torch.no_grad()
    def super_resolve(self, sr: torch.Tensor, *, x_T: Optional[torch.Tensor] = None,
                      noise_tape: Optional[torch.Tensor] = None, seed: Optional[int] = None, first_cube: int = 0,
                      clamp: bool = True, return_latents: bool = False):
        """sr: bicubic-upsampled cubes [B,C,H,W] on the GPU -> SR cubes [B,C,H,W] (clamped to [0,1] like sr_gae.py:474-475).

        x_T [B*G,3,H,W] / noise_tape [B*G,T-1,3,H,W] inject the random draws in (cube, group)-major order.
        With an explicit `seed` the draws are keyed by (seed, first_cube + position of the cube): a cube's result does not
        depend on which batch or rank it is processed in."""
        z = self.gae.encode_batched(sr)
        n = z.shape[0]
        g = n // sr.shape[0]
        step = self.max_latents or n
        outs = []
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            outs.append(self.diffusion.super_resolution(
                z[lo:hi], False, return_all=True, x_T=None if x_T is None else x_T[lo:hi],
                noise_tape=None if noise_tape is None else noise_tape[lo:hi],
                seed=seed, first_image=first_cube * g + lo))
        lat = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        y = self.gae.decode_batched(lat, clamp01=clamp)
        _lib.check_health(y.device)        # results are about to leave the library: a barrier timeout must not pass silently
        return (y, lat) if return_latents else y

    @torch.no_grad()
    def super_resolve_lr(self, lr: torch.Tensor, scale: int = 4, **kw):
        """From the LOW-RESOLUTION cubes [B,C,h,w] on the GPU: the dataset code's bicubic x`scale` pre-upsampling and clamp
        (sr_gae.py:72, HStest.py:59-60) run on the device (prepost.bicubic_upsample), then `super_resolve`."""
        from . import prepost
        return self.super_resolve(prepost.bicubic_upsample(lr, scale, clamp01=True), **kw)

    @torch.no_grad()
    def super_resolve_host(self, sr_host: torch.Tensor, device: torch.device, **kw) -> torch.Tensor:
        """End-to-end call with HOST buffers: pinned H2D of the cubes, the pipeline, D2H of the result."""
        if sr_host.is_cuda:
            raise ValueError("super_resolve_host expects a host tensor")
        src = sr_host if sr_host.is_pinned() else sr_host.pin_memory()
        dev = src.to(device, non_blocking=True)
        y = self.super_resolve(dev, **kw)
        out = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
        out.copy_(y, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return out


@torch.no_grad()
def validate(pipeline: SRPipeline, loader, device: torch.device, ratio: float = 4.0, **kw) -> dict:
    """The reference's validation loop (sr_gae.py:436-497) over an iterable of ``{'HR': [B,C,H,W], 'SR': [B,C,H,W]}``
    batches: encode -> sample -> decode -> clamp to [0,1] -> ``quality_assessment(gt, y, data_range=1., ratio=4)``
    (sr_gae.py:487-491) with all six indices computed on the device (hsidm_quality_assessment) and averaged over cubes
    like ``sum_dict`` / ``idx`` do.  MPSNR and SAM are the indices the parity gates are stated in.  The GAE is loaded once
    (the reference re-reads the pickle per cube, sr_gae.py:444); results stay on the device until the final averages."""
    from . import prepost
    total = torch.zeros(len(prepost.ASSESSMENT_KEYS), device=device, dtype=torch.float64)
    count = 0
    for batch in loader:
        sr = batch["SR"].to(device, non_blocking=True)
        hr = batch["HR"].to(device, non_blocking=True)
        if sr.dim() == 3:
            sr, hr = sr.unsqueeze(0), hr.unsqueeze(0)
        y = pipeline.super_resolve(sr.float().contiguous(), clamp=True, **kw)
        total += prepost.quality_assessment(hr.float().contiguous(), y, ratio=ratio).double().sum(dim=0)
        count += sr.shape[0]
    if count == 0:
        return {**{k: float("nan") for k in prepost.ASSESSMENT_KEYS}, "cubes": 0}
    mean = (total / count).cpu()
    return {**{k: float(mean[i]) for i, k in enumerate(prepost.ASSESSMENT_KEYS)}, "cubes": count}


def run_sharded(pipeline: SRPipeline, cubes_host: torch.Tensor, device: torch.device, rank: int, world: int,
                batch: int, gather: bool = False, keep_on_device: bool = False, **kw) -> Optional[torch.Tensor]:
    """Each rank super-resolves its contiguous slice of `cubes_host` in batches of `batch` cubes.

    No data-path collective.  With gather=True the per-rank results are collected on rank 0 once at the end with one
    torch.distributed gather of equally padded tensors (NCCL on device tensors, gloo on the CPU tests); other ranks
    return None.  keep_on_device=True leaves the results on the GPU (pipeline.super_resolve on pinned H2D copies) so that
    the gather and the blend never touch host memory.  An explicit `seed` in **kw keys the noise by global cube index."""
    lo, hi = shard_bounds(cubes_host.shape[0], rank, world)
    mine = _run_range(pipeline, cubes_host[lo:hi], lo, cubes_host.shape[1:], cubes_host.dtype, device, batch, keep_on_device, **kw)
    if not gather or world == 1:
        return mine
    import torch.distributed as dist
    return gather_rows(mine, cubes_host.shape[0], rank, world, dist)


def _run_range(pipeline: SRPipeline, mine_host: torch.Tensor, first: int, cube_shape, dtype, device: torch.device, batch: int,
               keep_on_device: bool, **kw) -> torch.Tensor:
    """The cubes `mine_host` (global indices first, first+1, ...) through the pipeline in batches of `batch`."""
    parts: List[torch.Tensor] = []
    for b0 in range(0, mine_host.shape[0], batch):
        chunk = mine_host[b0:b0 + batch]
        extra = dict(kw, first_cube=first + b0) if kw.get("seed") is not None else kw
        if keep_on_device:
            src = chunk if chunk.is_pinned() else chunk.pin_memory()
            parts.append(pipeline.super_resolve(src.to(device, non_blocking=True), **extra))
        else:
            parts.append(pipeline.super_resolve_host(chunk, device, **extra))
    if parts:
        return torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]
    return torch.zeros((0,) + tuple(cube_shape), dtype=dtype, device=device if keep_on_device else "cpu")


def gather_rows(mine: torch.Tensor, n_total: int, rank: int, world: int, dist) -> Optional[torch.Tensor]:
    """Row blocks of the contiguous shards (shard_bounds) -> the full [n_total, ...] tensor on rank 0, through ONE
    dist.gather of tensors padded to the largest shard (device tensors travel over NCCL/NVLink, no pickling)."""
    most = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    send = mine
    if mine.shape[0] < most:
        send = torch.zeros((most,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        send[:mine.shape[0]] = mine
    send = send.contiguous()
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, bufs, dst=0)
    if rank != 0:
        return None
    rows = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        rows.append(bufs[r][:hi - lo])
    return torch.cat(rows, dim=0)


# ---- overlapping-tile scene driver (BASELINE config 3) ------------------------------------------------
# The reference only crops non-overlapping 128x128 blocks offline (GAE/crop.py:12-36, HStest.py:33-45); full-scene SR with
# overlapping tiles and a feathered blend is the caller-side step the multi-GPU throughput config needs.
def tile_starts(size: int, tile: int, overlap: int) -> List[int]:
    """Tile origins along one axis: stride tile-overlap, last tile pulled back flush with the edge."""
    if size < tile:
        raise ValueError(f"scene side {size} is smaller than the tile {tile}")
    stride = tile - overlap
    starts = list(range(0, size - tile + 1, stride))
    if starts[-1] != size - tile:
        starts.append(size - tile)
    return starts


def tile_positions(height: int, width: int, tile: int = 128, overlap: int = 16) -> List[Tuple[int, int]]:
    """Tile origins (y0, x0) in row-major tile order."""
    return [(y, x) for y in tile_starts(height, tile, overlap) for x in tile_starts(width, tile, overlap)]


def tile_scene(scene: torch.Tensor, tile: int = 128, overlap: int = 16, lo: int = 0, hi: Optional[int] = None,
               pin: bool = False) -> Tuple[torch.Tensor, List[Tuple[int, int]]]:
    """scene [C,H,W] -> (tiles [T,C,tile,tile], [(y0,x0)...] of ALL tiles) in row-major tile order.  lo / hi restrict the
    returned tiles to that slice of the order (a rank's shard: the other tiles are never copied); pin=True writes them
    straight into pinned memory (no second staging copy before the H2D transfer)."""
    _, h, w = scene.shape
    pos = tile_positions(h, w, tile, overlap)
    hi = len(pos) if hi is None else hi
    tiles = torch.empty((max(0, hi - lo), scene.shape[0], tile, tile), dtype=scene.dtype, pin_memory=pin and torch.cuda.is_available())
    for i, (y, x) in enumerate(pos[lo:hi]):
        tiles[i].copy_(scene[:, y:y + tile, x:x + tile])
    return tiles, pos


def feather_window(tile: int, overlap: int, device=None) -> torch.Tensor:
    """Separable weight [tile,tile]: linear ramp over `overlap` pixels at every border, 1 inside (never zero)."""
    ramp = torch.ones(tile, device=device)
    if overlap > 0:
        edge = (torch.arange(overlap, device=device, dtype=torch.float32) + 1.0) / (overlap + 1.0)
        ramp[:overlap] = edge
        ramp[tile - overlap:] = edge.flip(0)
    return ramp[:, None] * ramp[None, :]


def blend_tiles(tiles: torch.Tensor, pos: List[Tuple[int, int]], height: int, width: int, overlap: int = 16) -> torch.Tensor:
    """Weighted overlap-add of [T,C,t,t] CUDA tiles (row-major grid `pos` from tile_scene) back into a [C,H,W] scene on the
    device (hsidm_blend_tiles: gather form, weights of `feather_window`, normalised per pixel, deterministic).  There is no
    host implementation in the package; tests/ keep a torch restatement as the checker."""
    tiles = _lib.require_cuda_f32(tiles, "tiles")
    ys = sorted({y for y, _ in pos})
    xs = sorted({x for _, x in pos})
    if [(y, x) for y in ys for x in xs] != list(pos) or tiles.shape[0] != len(pos):
        raise _lib.HsidmError(-1, "blend_tiles expects the row-major tile grid produced by tile_scene")
    t = tiles.shape[-1]
    ys_d = torch.tensor(ys, dtype=torch.int32, device=tiles.device)
    xs_d = torch.tensor(xs, dtype=torch.int32, device=tiles.device)
    out = torch.empty((tiles.shape[1], height, width), dtype=torch.float32, device=tiles.device)
    _lib.check(_lib.load().hsidm_blend_tiles(tiles.data_ptr(), ys_d.data_ptr(), len(ys), xs_d.data_ptr(), len(xs), tiles.shape[1],
                                             t, overlap, height, width, out.data_ptr(), _lib.stream_ptr(tiles.device)))
    return out


def super_resolve_scene(pipeline: SRPipeline, sr_scene: torch.Tensor, device: torch.device, tile: int = 128, overlap: int = 16,
                        batch: int = 8, rank: int = 0, world: int = 1, blend_fn=None, keep_on_device: bool = True,
                        **kw) -> Optional[torch.Tensor]:
    """Full-scene SR (BASELINE configs[2]): tile the bicubic-upsampled scene [C,H,W] (host), shard the tiles over `world`
    ranks (contiguous slices, no per-step collective), super-resolve them in batches, gather the tiles on rank 0 (one
    NCCL gather of device tensors) and blend them there on the GPU.  Returns the [C,H,W] device tensor on rank 0 (None
    elsewhere).  Pass an explicit seed=... for results that do not depend on `world` or `batch`."""
    n_tiles = len(tile_positions(sr_scene.shape[1], sr_scene.shape[2], tile, overlap))
    lo, hi = shard_bounds(n_tiles, rank, world)
    tiles, pos = tile_scene(sr_scene, tile, overlap, lo, hi, pin=keep_on_device)     # this rank's tiles only
    mine = _run_range(pipeline, tiles, lo, tiles.shape[1:], tiles.dtype, device, batch, keep_on_device, **kw)
    if world > 1:
        import torch.distributed as dist
        mine = gather_rows(mine, n_tiles, rank, world, dist)
    if rank != 0:
        return None
    return (blend_fn or blend_tiles)(mine, pos, sr_scene.shape[1], sr_scene.shape[2], overlap)
