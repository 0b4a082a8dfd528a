"""Host logic of the tiler / sharder on the CPU, incl. a world_size-2 gloo run (BASELINE configs[4])."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from tpu_superresolution_b200 import tiling


def test_positions_match_survey_config5():
    pos = tiling.axis_positions(4096, 64, 8)
    assert pos == list(range(0, 4096 - 64, 56)) + [4032] and len(pos) == 73          # 73 x 73 = 5329 tiles
    assert len(tiling.plan_tiles(4096, 4096, 64, 8)) == 5329
    assert len(tiling.axis_positions(4096, 64, 32)) == 127
    assert tiling.axis_positions(64, 64, 8) == [0]
    with pytest.raises(ValueError):
        tiling.axis_positions(40, 64, 8)
    # seam tiles computed twice by the sharded run (<= 9.6 % at N = 8)
    assert tiling.seam_recompute_fraction(4096, 4096, 64, 8, 1) == 0.0
    assert 0.0 < tiling.seam_recompute_fraction(4096, 4096, 64, 8, 8) <= 0.096 + 1e-9


@pytest.mark.parametrize("H,W,tile,ov", [(200, 136, 64, 8), (64, 64, 64, 8), (129, 70, 64, 32), (500, 90, 64, 0)])
def test_plan_covers_image_and_classes_are_disjoint(H, W, tile, ov):
    tiles = tiling.plan_tiles(H, W, tile, ov)
    cover = torch.zeros(H, W)
    for t in tiles:
        cover[t.y0:t.y0 + tile, t.x0:t.x0 + tile] += 1
    assert cover.min() >= 1
    chunks = tiling.batches_of(tiles, 16)
    assert all(len(c) == 16 for c in chunks[:-1]) and 1 <= len(chunks[-1]) <= 16
    for chunk in chunks:
        segs = tiling.class_segments(chunk)
        assert segs[0][0] == 0 and segs[-1][1] == len(chunk) and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        for a, b in segs:                       # one stitch launch per class segment
            assert len({t.cls for t in chunk[a:b]}) == 1
            seen = torch.zeros(H, W)
            for t in chunk[a:b]:
                seen[t.y0:t.y0 + tile, t.x0:t.x0 + tile] += 1
            assert seen.max() <= 1, "tiles of one stitch launch overlap"
    assert [t for c in chunks for t in c] == tiles
    # the cover count of the E / W rule is separable
    cy, cx = tiling.cover_counts(H, tile, ov, 1), tiling.cover_counts(W, tile, ov, 1)
    assert torch.equal(cover, cy[:, None] * cx[None, :])


def test_bands_partition_rows():
    for H, world in [(4096, 8), (200, 3), (64, 4), (5, 8)]:
        bands = tiling.assign_bands(H, world)
        assert bands[0][0] == 0 and bands[-1][1] == H and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        assert max(b[1] - b[0] for b in bands) - min(b[1] - b[0] for b in bands) <= 1


# ---- plain-torch stand-ins for the model and the CUDA stitch kernels (test doubles, CPU only)
def _fake_sr(x, scale=2):
    y = F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False)
    return y * 0.9 + 0.1 * y.mean(dim=(2, 3), keepdim=True)       # depends on the whole tile -> seams are visible


def _cpu_gather(slab, src_yx, out):
    n, c, th, tw = out.shape
    for k in range(n):
        y0, x0 = int(src_yx[k, 0]), int(src_yx[k, 1])
        out[k] = slab[:, y0:y0 + th, x0:x0 + tw]


def _cpu_accumulate(sr, E, dst_yx):
    n, c, th, tw = sr.shape
    for k in range(n):
        y0, x0 = int(dst_yx[k, 0]), int(dst_yx[k, 1])
        ys, ye = max(y0, 0), min(y0 + th, E.shape[1])
        if ye <= ys:
            continue
        E[:, ys:ye, x0:x0 + tw] += sr[k, :, ys - y0:ye - y0]


def _cpu_finalize(E, cnt_y, cnt_x, out):
    v = E / (cnt_y[:, None] * cnt_x[None, :])
    if out.dtype == torch.uint8:
        v = (v.clamp(0, 1) * 255).round()
    out.copy_(v.to(out.dtype))


def _resolver(batch=16, **kw):
    return tiling.TiledSuperResolver(None, scale=2, tile=64, overlap=8, batch=batch, run_tiles=_fake_sr, gather=_cpu_gather,
                                     accumulate=_cpu_accumulate, finalize=_cpu_finalize, **kw)


def test_single_rank_matches_direct_stitch_and_logical_ranks_are_bit_identical():
    lr = torch.rand(1, 3, 150, 200, generator=torch.Generator().manual_seed(0))
    full = _resolver()(lr)
    assert full.shape == (1, 3, 300, 400) and torch.isfinite(full).all()
    # direct E / W evaluation in plain row-major order is the same up to fp32 summation order
    E, Wt = torch.zeros(3, 300, 400), torch.zeros(300, 400)
    for t in sorted(tiling.plan_tiles(150, 200, 64, 8), key=lambda t: (t.iy, t.ix)):
        y = _fake_sr(lr[:, :, t.y0:t.y0 + 64, t.x0:t.x0 + 64])[0]
        E[:, 2 * t.y0:2 * t.y0 + 128, 2 * t.x0:2 * t.x0 + 128] += y
        Wt[2 * t.y0:2 * t.y0 + 128, 2 * t.x0:2 * t.x0 + 128] += 1
    assert (full[0] - E / Wt).abs().max() < 1e-6
    # N logical ranks on one device: concatenated bands are bit-identical to the 1-rank result, any batch size
    for world in (2, 3, 4, 8):
        bands = [_resolver(batch=5).band(lr, r, world)[0] for r in range(world)]
        assert torch.equal(torch.cat(bands, dim=1), full[0])


def test_small_images_uint8_input_and_output_dtypes():
    """Image sides shorter than the tile are one tile of that side; uint8 LR is scaled to [0, 1]; bf16 / uint8 outputs."""
    lr = torch.rand(1, 3, 40, 100, generator=torch.Generator().manual_seed(1))
    full = _resolver()(lr)
    assert full.shape == (1, 3, 80, 200)
    ref = torch.zeros(3, 80, 200)
    cnt = torch.zeros(80, 200)
    for t in tiling.plan_tiles(40, 100, (40, 64), 8):
        ref[:, :, 2 * t.x0:2 * t.x0 + 128] += _fake_sr(lr[:, :, :, t.x0:t.x0 + 64])[0]
        cnt[:, 2 * t.x0:2 * t.x0 + 128] += 1
    assert (full[0] - ref / cnt).abs().max() < 1e-6
    tiny = torch.rand(1, 3, 24, 17, generator=torch.Generator().manual_seed(2))
    assert torch.equal(_resolver()(tiny)[0], _fake_sr(tiny)[0])
    lr8 = (lr * 255).round().to(torch.uint8)
    y8 = _resolver(out_dtype=torch.uint8)(lr8)
    yf = _resolver()(lr8.float() / 255)
    assert y8.dtype == torch.uint8 and (y8.float() - (yf.clamp(0, 1) * 255)).abs().max() <= 0.5 + 1e-4
    yb = _resolver(out_dtype=torch.bfloat16)(lr8)
    assert yb.dtype == torch.bfloat16 and (yb.float() - yf).abs().max() <= 4e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lr = torch.rand(1, 3, 150, 200, generator=torch.Generator().manual_seed(0))
    out = _resolver()(lr, rank=rank, world=world)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gather_is_bit_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    lr = torch.rand(1, 3, 150, 200, generator=torch.Generator().manual_seed(0))
    assert torch.equal(out, _resolver()(lr))
