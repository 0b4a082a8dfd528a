"""Overlapping-tile inference of large LR images, sharded over the GPUs of one box (BASELINE.json configs[4]).

The reference has no tiling (`validate` feeds whole images, finetune_swinir.py:195); this is the
upstream-SwinIR tiling rule the north star asks for: cut the LR image into `tile` x `tile` tiles every
`tile - overlap` pixels (the last tile of an axis is flush with the border), super-resolve the tiles independently,
sum them into E, count them into W and return E / W.  An image side shorter than the tile is one tile of that side
(the model reflect-pads it to a window multiple itself, network_swinir.py:783-788).

Determinism: tiles are ordered class by class, class = (parity-or-last of the tile row, parity-or-last of the tile
column).  Two tiles of one class never overlap (stride >= tile / 2), so one stitch launch handles a class segment of a
batch without atomics, and every output pixel receives its (at most four) contributions in a fixed global order.  A
rank that owns an HR row band computes every tile touching the band (seam tiles are computed twice, SURVEY.md §8e
option i) and keeps only its own rows -- the sharded result is therefore bit-identical to the single-GPU result, and
the only communication is the final gather of the bands (no collective inside the hot path).

Hot loop (per batch of `batch` tiles, nothing crosses the host): one gather kernel cuts the tiles out of the rank's LR
band by a device-resident coordinate table (`srk_gather_tiles`), the model runs as one CUDA-graph replay (ragged last
batch: padded with repeats of its last tile), and one stitch launch per class segment adds the outputs into E
(`srk_stitch_accumulate_strided`).  W is not accumulated: the cover count is separable, W[y, x] = cnt_y[y] * cnt_x[x],
and `srk_stitch_finalize` divides by it while converting to the output dtype (fp32, bf16 or uint8).  The gather sends
each band straight into its rows of the final image on rank 0 (one send per channel plane, no padding, no staging).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

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


def plan_tiles(height: int, width: int, tile, overlap: int) -> List[Tile]:
    """All tiles in processing order: by class, then row-major.  `tile`: side or (tile_h, tile_w)."""
    th, tw = (tile, tile) if isinstance(tile, int) else tile
    ys, xs = axis_positions(height, th, min(overlap, th // 2)), axis_positions(width, tw, min(overlap, tw // 2))
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
    """Plain chunks of `batch` tiles over the class-ordered list (only the last one may be ragged)."""
    return [list(tiles[i:i + batch]) for i in range(0, len(tiles), batch)]


def class_segments(chunk: Sequence[Tile]) -> List[Tuple[int, int]]:
    """[a, b) index ranges of a batch with one class each: the tiles of one stitch launch must be disjoint."""
    segs, a = [], 0
    for i in range(1, len(chunk) + 1):
        if i == len(chunk) or chunk[i].cls != chunk[a].cls:
            segs.append((a, i))
            a = i
    return segs


def cover_counts(size: int, tile: int, overlap: int, scale: int) -> torch.Tensor:
    """Number of tiles covering each HR row (or column) of an axis: W[y, x] = cnt_y[y] * cnt_x[x]."""
    cnt = torch.zeros(size * scale, dtype=torch.float32)
    for p in axis_positions(size, tile, min(overlap, tile // 2)):
        cnt[p * scale:(p + tile) * scale] += 1
    return cnt


def seam_recompute_fraction(height: int, width: int, tile: int, overlap: int, world: int) -> float:
    """Extra tiles the sharded run computes (seam tiles run on both neighbours) / tiles of the single-GPU run."""
    th, tw = min(tile, height), min(tile, width)
    tiles = plan_tiles(height, width, (th, tw), overlap)
    total = sum(len(tiles_for_band(tiles, th, b)) for b in assign_bands(height, world) if b[1] > b[0])
    return total / len(tiles) - 1.0


# ---- injection points (the CPU tests of the host logic pass plain-torch stand-ins; the product path has no fallback)
def _cuda_gather(slab: torch.Tensor, src_yx: torch.Tensor, out: torch.Tensor) -> None:
    from . import _lib as L
    L.gather_tiles(slab, src_yx, out)


def _cuda_accumulate(sr_tiles: torch.Tensor, E: torch.Tensor, dst_yx: torch.Tensor) -> None:
    from . import _lib as L
    L.stitch_accumulate_strided(sr_tiles if sr_tiles.dtype == torch.float32 else sr_tiles.float(), E, dst_yx)


def _cuda_finalize(E: torch.Tensor, cnt_y: torch.Tensor, cnt_x: torch.Tensor, out: torch.Tensor) -> None:
    from . import _lib as L
    L.stitch_finalize(E, cnt_y, cnt_x, out)


@dataclass
class _Plan:
    band: Tuple[int, int]               # LR rows [r0, r1) this rank owns
    rows: Tuple[int, int]               # LR rows [y_lo, y_hi) its tiles read (band + halo)
    tile_hw: Tuple[int, int]
    batches: List[Tuple[int, int, List[Tuple[int, int]]]]      # (first tile, count, class segments) per batch
    n_tiles: int
    src_yx: torch.Tensor                # (n_padded, 2) int32, relative to rows[0]; padded with repeats of the last tile
    dst_yx: torch.Tensor                # (n_padded, 2) int32 HR coordinates relative to the band
    cnt_y: torch.Tensor                 # (band rows * s,) cover counts of the band's HR rows
    cnt_x: torch.Tensor                 # (W * s,)


class TiledSuperResolver:
    """sr = TiledSuperResolver(model, scale)(lr)  with lr (1, C, H, W), float32 in [0, 1] or uint8, on the GPU or in
    (pinned) host memory -- a rank only ever copies the rows its own tiles read.

    out_dtype: torch.float32 (default), torch.bfloat16 or torch.uint8 (round(clamp(v, 0, 1) * 255)): the dtype of the
    stitched image, converted inside the finalize kernel; the cross-GPU gather moves that dtype (uint8 = 1/4 of the bytes).
    graph: replay the model as a CUDA graph (`GraphedModel`); every batch then has exactly `batch` tiles.
    `run_tiles`, `gather`, `accumulate`, `finalize` are injection points for the CPU tests of the host logic.
    """

    def __init__(self, model: Optional[torch.nn.Module], scale: int, tile: int = 64, overlap: int = 8, batch: int = 16,
                 run_tiles: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, out_dtype: torch.dtype = torch.float32,
                 graph: bool = True, gather: Callable = _cuda_gather, accumulate: Callable = _cuda_accumulate,
                 finalize: Callable = _cuda_finalize):
        self.scale, self.tile, self.overlap, self.batch, self.out_dtype = scale, tile, overlap, batch, out_dtype
        self.pad_batches = False
        if run_tiles is not None:
            self.run_tiles = run_tiles
        elif graph and model is not None:
            from .graphs import GraphedModel
            self.run_tiles = GraphedModel(model)
            self.pad_batches = True             # one graph for every batch: the ragged last one is padded with repeats
        else:
            self.run_tiles = model
        self.gather, self.accumulate, self.finalize = gather, accumulate, finalize
        self._plans: Dict[Tuple, _Plan] = {}
        self.last_stats: Dict[str, float] = {}

    # ------------------------------------------------------------------------------------------ planning (host, cached)
    def plan(self, H: int, W: int, rank: int, world: int, device) -> _Plan:
        key = (H, W, rank, world, str(device))
        p = self._plans.get(key)
        if p is not None:
            return p
        s = self.scale
        th, tw = min(self.tile, H), min(self.tile, W)
        tiles = plan_tiles(H, W, (th, tw), self.overlap)
        r0, r1 = assign_bands(H, world)[rank]
        mine = tiles_for_band(tiles, th, (r0, r1)) if r1 > r0 else []
        y_lo = min((t.y0 for t in mine), default=r0)
        y_hi = max((t.y0 + th for t in mine), default=r0)
        chunks = batches_of(mine, self.batch)
        batches, first = [], 0
        for c in chunks:
            batches.append((first, len(c), class_segments(c)))
            first += len(c)
        padded = list(mine)
        if mine and self.pad_batches and len(mine) % self.batch:
            padded += [mine[-1]] * (self.batch - len(mine) % self.batch)
        src = torch.tensor([[t.y0 - y_lo, t.x0] for t in padded], dtype=torch.int32).reshape(-1, 2)
        dst = torch.tensor([[(t.y0 - r0) * s, t.x0 * s] for t in padded], dtype=torch.int32).reshape(-1, 2)
        cnt_y = cover_counts(H, th, self.overlap, s)[r0 * s:r1 * s].contiguous()
        cnt_x = cover_counts(W, tw, self.overlap, s)
        p = _Plan((r0, r1), (y_lo, y_hi), (th, tw), batches, len(mine), src.to(device), dst.to(device), cnt_y.to(device), cnt_x.to(device))
        self._plans[key] = p
        return p

    # ------------------------------------------------------------------------------------------ one rank's band
    @torch.no_grad()
    def band(self, lr: torch.Tensor, rank: int = 0, world: int = 1, device=None, out: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, Tuple[int, int]]:
        """Stitched HR rows of this rank's band: ((C, rows * scale, W * scale) of `out_dtype`, (lr_row0, lr_row1)).
        `out`: optional destination view with contiguous planes (e.g. the band's rows of the final image)."""
        _, C, H, W = lr.shape
        if device is None:
            device = lr.device
        device = torch.device(device)
        s = self.scale
        p = self.plan(H, W, rank, world, device)
        r0, r1 = p.band
        th, tw = p.tile_hw
        # the LR rows this rank reads; a host image is copied band-wise (pinned memory: asynchronously)
        if lr.device == device:
            slab = lr[0, :, p.rows[0]:p.rows[1], :]
        else:                                   # one contiguous (rows x W) copy per channel plane: asynchronous from pinned memory
            slab = torch.empty(C, p.rows[1] - p.rows[0], W, device=device, dtype=lr.dtype)
            for c in range(C):
                slab[c].copy_(lr[0, c, p.rows[0]:p.rows[1], :], non_blocking=True)
        if slab.dtype == torch.uint8:
            slab = slab.float().div_(255.0)
        slab = slab.float().contiguous()
        rows = (r1 - r0) * s
        fp32_out = self.out_dtype == torch.float32
        if out is not None and (tuple(out.shape) != (C, rows, W * s) or out.dtype != self.out_dtype):
            raise ValueError("band(out=...): wrong shape / dtype")
        if fp32_out and out is not None:
            E = out
            E.zero_()
        else:
            E = torch.zeros(C, rows, W * s, device=device, dtype=torch.float32)
        x = torch.empty(self.batch, C, th, tw, device=device, dtype=torch.float32)
        for first, n, segs in p.batches:
            nb = self.batch if self.pad_batches else n
            xb = x[:nb]
            self.gather(slab, p.src_yx[first:first + nb], xb)
            y = self.run_tiles(xb)
            for a, b in segs:                 # one launch per class segment: disjoint tiles, fixed accumulation order
                self.accumulate(y[a:b], E, p.dst_yx[first + a:first + b])
        if out is None:
            out = E if fp32_out else torch.empty(C, rows, W * s, device=device, dtype=self.out_dtype)
        if rows > 0:
            self.finalize(E, p.cnt_y, p.cnt_x, out)
        return out, (r0, r1)

    # ------------------------------------------------------------------------------------------ whole image
    @torch.no_grad()
    def __call__(self, lr: torch.Tensor, rank: int = 0, world: int = 1, gather: bool = True, device=None) -> Optional[torch.Tensor]:
        """Whole stitched image (1, C, H*s, W*s) on rank 0 (None elsewhere); with gather=False every rank returns its band."""
        if world == 1 or not gather:
            return self.band(lr, rank, world, device)[0].unsqueeze(0)
        import torch.distributed as dist
        _, C, H, W = lr.shape
        s = self.scale
        if device is None:
            device = lr.device
        bands = assign_bands(H, world)
        if rank == 0:
            full = torch.empty(1, C, H * s, W * s, device=device, dtype=self.out_dtype)
            r0, r1 = bands[0]
            self.band(lr, 0, world, device, out=full[0, :, r0 * s:r1 * s, :])        # rank 0's band is stitched in place
            ops = [dist.P2POp(dist.irecv, full[0, c, b0 * s:b1 * s, :], src)
                   for src, (b0, b1) in enumerate(bands) if src != 0 and b1 > b0 for c in range(C)]
        else:
            mine, (b0, b1) = self.band(lr, rank, world, device)
            ops = [dist.P2POp(dist.isend, mine[c], 0) for c in range(C)] if b1 > b0 else []
        # the single exchange of the sharded run: every band goes straight into its rows of the final image (NVLink under NCCL)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return full if rank == 0 else None
