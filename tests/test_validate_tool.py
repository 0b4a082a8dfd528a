"""tools/validate.py: the reference's checkpoint + evaluation loop (finetune_swinir.py:69-74, :182-207, :283-285; sr_datasets.py
:31-74) on the fused path.  CPU: checkpoint loading, pairing, metrics.  GPU: the whole loop on synthetic PNG pairs."""
import importlib.util
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("srk_validate", os.path.join(ROOT, "tools", "validate.py"))
V = importlib.util.module_from_spec(spec)
spec.loader.exec_module(V)


def test_metrics_and_pairing(tmp_path):
    from PIL import Image
    a = torch.rand(2, 3, 40, 48, generator=torch.Generator().manual_seed(0))
    assert torch.allclose(V.ssim(a, a), torch.ones(2, dtype=torch.float64))
    assert V.ssim(a, (a + 0.1 * torch.randn(a.shape, generator=torch.Generator().manual_seed(1))).clamp(0, 1)).max() < 0.95
    b = (a + 0.01).clamp(0, 1)
    mse = ((a.clamp(0, 1) - b) ** 2).flatten(1).mean(1)
    assert torch.allclose(V.batch_psnr(a, b), 20 * torch.log10(1 / torch.sqrt(mse + 1e-8)))          # finetune_swinir.py:69-74
    (tmp_path / "lr").mkdir(); (tmp_path / "hr").mkdir()
    for stem in ("a", "b"):
        Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(tmp_path / "hr" / f"{stem}.png")
    Image.fromarray(np.zeros((2, 2, 3), np.uint8)).save(tmp_path / "lr" / "a_x4.png")
    Image.fromarray(np.zeros((2, 2, 3), np.uint8)).save(tmp_path / "lr" / "bx4.png")
    Image.fromarray(np.zeros((2, 2, 3), np.uint8)).save(tmp_path / "lr" / "orphan_x4.png")
    pairs = V.pair_files(str(tmp_path / "lr"), str(tmp_path / "hr"), 4)
    assert [(p.name, h.name) for p, h in pairs] == [("a_x4.png", "a.png"), ("bx4.png", "b.png")]


def test_checkpoint_loading_is_strict(tmp_path):
    import tpu_superresolution_b200 as srk
    from tpu_superresolution_b200 import synth
    cfg = synth.CONFIGS["swinir_x4_d2"]
    sd = synth.make_swinir_state_dict(cfg, seed=3, kind="init")
    torch.save({"params": sd}, tmp_path / "a.pth")
    torch.save(sd, tmp_path / "b.pth")
    for name in ("a.pth", "b.pth"):
        m = srk.SwinIR(**cfg.as_kwargs()).eval()
        V.load_checkpoint(m, str(tmp_path / name))
        assert torch.equal(m.layers[0].residual_group.blocks[0].attn.qkv.weight, sd["layers.0.residual_group.blocks.0.attn.qkv.weight"])
    bad = dict(sd)
    bad.pop("conv_last.bias")
    torch.save({"params": bad}, tmp_path / "c.pth")
    with pytest.raises(RuntimeError):
        V.load_checkpoint(srk.SwinIR(**cfg.as_kwargs()), str(tmp_path / "c.pth"))


@pytest.mark.gpu
def test_selftest_whole_vs_tiled_on_the_gpu():
    r = V.selftest(torch.device("cuda"))
    assert r["pairs"] == 2
    (l1a, pa, sa), (l1b, pb, sb) = r["whole"], r["tiled"]
    assert np.isfinite([l1a, pa, sa, l1b, pb, sb]).all()
    assert abs(pa - pb) < 0.05 and abs(sa - sb) < 5e-3           # tiling only changes pixels near tile seams (reflect pad vs neighbours)
