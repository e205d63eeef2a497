"""Host-side (torch) restatements used as CHECKERS by the tests; nothing in the package imports this."""
from typing import List, Tuple

import torch

from hsi_dmgasr_b200.pipeline import feather_window


def blend_tiles_ref(tiles: torch.Tensor, pos: List[Tuple[int, int]], height: int, width: int, overlap: int = 16) -> torch.Tensor:
    """Weighted overlap-add of [T,C,t,t] tiles back into a [C,H,W] scene (weights normalised per pixel); the CUDA kernel
    behind pipeline.blend_tiles performs the same fp32 operations in the same order."""
    t = tiles.shape[-1]
    win = feather_window(t, overlap, tiles.device)
    acc = torch.zeros((tiles.shape[1], height, width), device=tiles.device, dtype=torch.float32)
    wsum = torch.zeros((height, width), device=tiles.device, dtype=torch.float32)
    for k, (y, x) in enumerate(pos):
        acc[:, y:y + t, x:x + t] += tiles[k].float() * win
        wsum[y:y + t, x:x + t] += win
    return acc / wsum
