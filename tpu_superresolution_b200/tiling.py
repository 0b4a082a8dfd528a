"""Overlapping-tile inference of large LR images, sharded over the GPUs of one box (BASELINE.json configs[4]).

The reference has no tiling (`validate` feeds whole images, finetune_swinir.py:195); this is the
upstream-SwinIR tiling rule the north star asks for: cut the LR image into `tile` x `tile` tiles every
`tile - overlap` pixels (the last tile of an axis is flush with the border), super-resolve the tiles independently,
sum them into E, count them into W and return E / W.

Determinism: tiles are processed class by class, class = (parity-or-last of the tile row, parity-or-last of the
tile column).  Two tiles of one class never overlap (stride >= tile / 2), so one stitch launch handles a whole batch
without atomics, and every output pixel receives its (at most four) contributions in a fixed global order.  A rank
that owns an HR row band computes every tile touching the band (seam tiles are computed twice, SURVEY.md §8e option i)
and keeps only its own rows -- the sharded result is therefore bit-identical to the single-GPU result, and the only
communication is the final gather of the bands (no collective inside the hot path).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def axis_positions(size: int, tile: int, overlap: int) -> List[int]:
    """Start offsets of the tiles along one axis: range(0, size - tile, tile - overlap) + [size - tile]."""
    if size < tile:
        raise ValueError(f"image side {size} smaller than the tile {tile}")
    if not (0 <= overlap <= tile // 2):
        raise ValueError("overlap must be in [0, tile/2]")
    stride = tile - overlap
    pos = list(range(0, size - tile, stride))
    pos.append(size - tile)
    return pos


def _axis_class(i: int, n: int) -> int:
    return 2 if (i == n - 1 and n > 1) else (i & 1)


@dataclass(frozen=True)
class Tile:
    iy: int
    ix: int
    y0: int
    x0: int
    cls: int        # 0..8: 3 * row class + column class


def plan_tiles(height: int, width: int, tile: int, overlap: int) -> List[Tile]:
    """All tiles in processing order: by class, then row-major."""
    ys, xs = axis_positions(height, tile, overlap), axis_positions(width, tile, overlap)
    tiles = [Tile(iy, ix, y, x, 3 * _axis_class(iy, len(ys)) + _axis_class(ix, len(xs)))
             for iy, y in enumerate(ys) for ix, x in enumerate(xs)]
    return sorted(tiles, key=lambda t: (t.cls, t.iy, t.ix))


def assign_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous LR row bands [r0, r1), balanced to +-1 row."""
    base, rem = divmod(height, world)
    bands, r = [], 0
    for k in range(world):
        n = base + (1 if k < rem else 0)
        bands.append((r, r + n))
        r += n
    return bands


def tiles_for_band(tiles: Sequence[Tile], tile: int, band: Tuple[int, int]) -> List[Tile]:
    """Tiles whose LR rows intersect the band, in the global processing order."""
    r0, r1 = band
    return [t for t in tiles if t.y0 < r1 and t.y0 + tile > r0]


def batches_of(tiles: Sequence[Tile], batch: int) -> List[List[Tile]]:
    """Chunk into batches that never mix classes (tiles of one stitch launch must be disjoint)."""
    out: List[List[Tile]] = []
    cur: List[Tile] = []
    for t in tiles:
        if cur and (len(cur) == batch or cur[-1].cls != t.cls):
            out.append(cur)
            cur = []
        cur.append(t)
    if cur:
        out.append(cur)
    return out


def _cuda_accumulate(sr_tiles: torch.Tensor, E: torch.Tensor, Wt: torch.Tensor, yx: torch.Tensor) -> None:
    from . import _lib as L
    n, c, th, tw = sr_tiles.shape
    L.stitch_accumulate(sr_tiles.contiguous(), E, Wt, yx, channels=c, tile_h=th, tile_w=tw, out_h=E.shape[1], out_w=E.shape[2])


def _cuda_normalize(E: torch.Tensor, Wt: torch.Tensor) -> None:
    from . import _lib as L
    L.stitch_normalize(E, Wt, channels=E.shape[0], pixels=E.shape[1] * E.shape[2])


class TiledSuperResolver:
    """sr = TiledSuperResolver(model, scale)(lr)  with lr (1, C, H, W) on the GPU.

    `run_tiles`, `accumulate`, `normalize` are injection points for the CPU tests of the host logic (tests pass
    plain-torch stand-ins); the product path uses the model and the libsrk stitch kernels and has no fallback.
    """

    def __init__(self, model: Optional[torch.nn.Module], scale: int, tile: int = 64, overlap: int = 8, batch: int = 16,
                 run_tiles: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 accumulate: Callable = _cuda_accumulate, normalize: Callable = _cuda_normalize):
        self.scale, self.tile, self.overlap, self.batch = scale, tile, overlap, batch
        self.run_tiles = run_tiles if run_tiles is not None else model
        self.accumulate, self.normalize = accumulate, normalize

    @torch.no_grad()
    def band(self, lr: torch.Tensor, rank: int = 0, world: int = 1) -> Tuple[torch.Tensor, Tuple[int, int]]:
        """Stitched HR rows of this rank's band: ((C, rows * scale, W * scale), (lr_row0, lr_row1))."""
        _, C, H, W = lr.shape
        s, T = self.scale, self.tile
        tiles = plan_tiles(H, W, T, self.overlap)
        r0, r1 = assign_bands(H, world)[rank]
        mine = tiles_for_band(tiles, T, (r0, r1))
        E = torch.zeros(C, (r1 - r0) * s, W * s, device=lr.device, dtype=torch.float32)
        Wt = torch.zeros((r1 - r0) * s, W * s, device=lr.device, dtype=torch.float32)
        for chunk in batches_of(mine, self.batch):
            x = torch.stack([lr[0, :, t.y0:t.y0 + T, t.x0:t.x0 + T] for t in chunk])
            y = self.run_tiles(x)
            yx = torch.tensor([[(t.y0 - r0) * s, t.x0 * s] for t in chunk], dtype=torch.int32, device=lr.device)
            self.accumulate(y.float(), E, Wt, yx)
        self.normalize(E, Wt)
        return E, (r0, r1)

    @torch.no_grad()
    def __call__(self, lr: torch.Tensor, rank: int = 0, world: int = 1, gather: bool = True) -> Optional[torch.Tensor]:
        """Whole stitched image (1, C, H*s, W*s) on rank 0 (None elsewhere); with gather=False every rank returns its band."""
        E, _ = self.band(lr, rank, world)
        if world == 1 or not gather:
            return E.unsqueeze(0)
        import torch.distributed as dist
        _, C, H, W = lr.shape
        bands = assign_bands(H, world)
        rows = max(b[1] - b[0] for b in bands) * self.scale
        pad = torch.zeros(C, rows, W * self.scale, device=E.device, dtype=E.dtype)
        pad[:, :E.shape[1]] = E
        out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, out, dst=0)          # the single exchange of the sharded run (NVLink / NVSwitch under NCCL)
        if rank != 0:
            return None
        return torch.cat([o[:, :(b[1] - b[0]) * self.scale] for o, b in zip(out, bands)], dim=1).unsqueeze(0)
