"""CPU-side checks of the HAT / DAT drop-in boundary: state_dict manifests of the real reference, strided bias tables against
the reference's index formulas, weight-stream packing against the plain Linear math.  No GPU, no kernel launches."""
import json
import os

import numpy as np
import pytest
import torch

import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L, packing
from oracle import synth
from oracle import hat_oracle as HO
from oracle import dat_oracle as DO

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("family,name", [("hat", "hat_x4"), ("dat", "dat_x2")])
def test_state_dict_matches_reference_manifest(family, name):
    man = json.load(open(os.path.join(GOLDEN, f"{name}_manifest.json")))
    cfg = (synth.HAT_CONFIGS if family == "hat" else synth.DAT_CONFIGS)[name]
    m = (srk.HAT if family == "hat" else srk.DAT)(**cfg.as_kwargs())
    sd = m.state_dict()
    assert list(sd) == [e[0] for e in man]
    for k, shape, dtype in man:
        assert list(sd[k].shape) == shape and str(sd[k].dtype).replace("torch.", "") == dtype, k


def _lookup(tab, c0, sy, qh, qw, kh, kw):
    """tab[h][c0 + sy * (yi - yj) + (xi - xj)] for every query i (qh x qw) and key j (kh x kw) -> (heads, Nq, Nk)."""
    ti, tj = torch.arange(qh * qw), torch.arange(kh * kw)
    idx = c0 + sy * ((ti // qw)[:, None] - (tj // kw)[None, :]) + ((ti % qw)[:, None] - (tj % kw)[None, :])
    assert int(idx.min()) >= 0 and int(idx.max()) < tab.shape[1]
    return tab[:, idx.reshape(-1)].reshape(tab.shape[0], qh * qw, kh * kw)


def test_bias_tables_reproduce_the_reference_gathers():
    rng = np.random.default_rng(3)
    # HAT W-MSA: table[rpi_sa] (hat_arch.py:184-186); kernel constants of capi.cu: SY 48, c0 = 15 * 48 + 15
    t = torch.from_numpy(rng.normal(size=(961, 6)).astype(np.float32))
    ref = t[HO.rpi_sa(16).reshape(-1)].reshape(256, 256, 6).permute(2, 0, 1) * packing.LOG2E
    assert torch.allclose(_lookup(packing.pack_bias_table_wmsa(t), 15 * 48 + 15, 48, 16, 16, 16, 16), ref)
    # HAT OCAB: negative indices wrap (hat_arch.py:424); SY 48, c0 = 23 * 48 + 23
    t = torch.from_numpy(rng.normal(size=(1521, 6)).astype(np.float32))
    ref = t[HO.rpi_oca(16, 0.5).reshape(-1)].reshape(256, 576, 6).permute(2, 0, 1) * packing.LOG2E     # python indexing wraps negatives
    assert torch.allclose(_lookup(packing.pack_bias_table_ocab(t), 23 * 48 + 23, 48, 16, 16, 24, 24), ref)
    # DAT: dynamic position bias gathered by the rectangular index (dat_arch.py:221-225)
    for hs, ws, sy in ((8, 32, 64), (32, 8, 24)):
        pos = torch.from_numpy(rng.normal(size=((2 * hs - 1) * (2 * ws - 1), 3)).astype(np.float32))
        ref = pos[DO.rect_relative_position_index(hs, ws).reshape(-1)].reshape(256, 256, 3).permute(2, 0, 1) * packing.LOG2E
        got = _lookup(packing.pack_bias_table_rect(pos, hs, ws, sy), (hs - 1) * sy + ws - 1, sy, hs, ws, hs, ws)
        assert torch.allclose(got[:3], ref) and float(got[3].abs().max()) == 0.0
    lib = L.load()
    assert [lib.srk_window_attention_table_floats(k) for k in range(4)] == [31 * 48, 39 * 48, 15 * 64, 63 * 24]


def _unpack_stream(ws, n_chunks, k_atoms):
    slabs = ws.view(torch.bfloat16).reshape(n_chunks, k_atoms, 192, 64)
    rows = [torch.cat([packing.unswizzle_slab(slabs[c, ka]) for ka in range(k_atoms)], 1) for c in range(n_chunks)]
    return torch.cat(rows, 0).float()                                  # (n_chunks * 192, k_atoms * 64)


def test_qkv_plane_packing_matches_linear_math():
    rng = np.random.default_rng(4)
    w = torch.from_numpy(rng.normal(0, 0.1, size=(540, 180)).astype(np.float32))
    b = torch.from_numpy(rng.normal(0, 0.1, size=(540,)).astype(np.float32))
    g = torch.from_numpy(rng.uniform(0.5, 1.5, size=(180,)).astype(np.float32))
    be = torch.from_numpy(rng.normal(0, 0.1, size=(180,)).astype(np.float32))
    x = torch.from_numpy(rng.normal(size=(7, 180)).astype(np.float32))
    xhat = (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    ref = (xhat * g + be) @ w.T + b
    ref[:, :180] *= 30 ** -0.5 * packing.LOG2E
    # HAT layout: q | k | v, heads padded to 32
    ws, bias = packing.pack_qkv_planes(w, b, g, be)
    got = (torch.nn.functional.pad(xhat, (0, 12)) @ _unpack_stream(ws, 3, 3).T + bias).view(7, 3, 6, 32)
    assert torch.allclose(got[..., :30].reshape(7, 540), ref, atol=2e-2)
    assert bool((got[:, 2, :, 30] == 1.0).all()) and float(got[:, :2, :, 30:].abs().max()) == 0.0 and float(got[:, 2, :, 31].abs().max()) == 0.0
    # DAT layout: per part the slots (h0, h1, h2, -, h3, h4, h5, -)
    ws, bias = packing.pack_dat_qkv_planes(w, b, g, be)
    got = (torch.nn.functional.pad(xhat, (0, 12)) @ _unpack_stream(ws, 4, 3).T + bias).view(7, 3, 8, 32)
    heads = got[:, :, [0, 1, 2, 4, 5, 6], :30].reshape(7, 540)
    assert torch.allclose(heads, ref, atol=2e-2)
    assert float(got[:, :, [3, 7]].abs().max()) == 0.0                   # padding head slots are all-zero (no ones column either)
    # fp32-row layout: 180-column chunks
    ws, bias = packing.pack_rows_linear(w, b, g, be)
    got = (torch.nn.functional.pad(xhat, (0, 12)) @ _unpack_stream(ws, 3, 3).T + bias).view(7, 3, 192)[..., :180].reshape(7, 540)
    ref2 = (xhat * g + be) @ w.T + b
    assert torch.allclose(got, ref2, atol=2e-2)


def test_pad_pages_layout():
    pages = packing.make_pad_pages("cpu")
    assert pages.numel() == 8192 and int(pages[:4096].abs().max()) == 0
    rows = packing.unswizzle_planes(pages[4096:].view(torch.bfloat16).reshape(1, 32, 64), 0)[0].float()
    expect = torch.zeros(32, 64)
    expect[:, 30] = 1.0
    expect[:, 62] = 1.0
    assert torch.equal(rows, expect)
