"""Host-side contracts of the drop-in models on the GPU: packed-weight cache invalidation after in-place ``.data`` edits (which
bump neither ``_version`` nor ``data_ptr``), and a model living on a device that is not the current one (every launch runs under
a device guard on the tensors' device and stream)."""
import pytest
import torch

import tpu_superresolution_b200 as srk
from tpu_superresolution_b200 import _lib as L
from oracle import synth

pytestmark = pytest.mark.gpu

FAMILIES = {
    "swinir": (lambda: synth.CONFIGS["swinir_x2_d2"], synth.make_swinir_state_dict, "SwinIR"),
    "hat": (lambda: synth.HAT_CONFIGS["hat_x4_d2"], synth.make_hat_state_dict, "HAT"),
    "dat": (lambda: synth.DAT_CONFIGS["dat_x2_d3"], synth.make_dat_state_dict, "DAT"),
}


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L.load()
    with torch.no_grad():
        yield


def _build(name, seed, device="cuda"):
    cfg_of, make_sd, cls = FAMILIES[name]
    cfg = cfg_of()
    m = getattr(srk, cls)(**cfg.as_kwargs()).eval()
    m.load_state_dict(make_sd(cfg, seed=seed, kind="stress"), strict=True)
    return m.to(device), cfg


@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_invalidate_packed_after_in_place_data_edit(name):
    """EMA-style ``p.data.mul_().add_()``: without invalidate_packed() the packed images are stale by design (documented);
    after it the model must equal a freshly built model holding the edited weights, bit for bit."""
    m, cfg = _build(name, 11)
    lr = synth.make_lr_batch(1, 64, 64, seed=3).cuda()
    y0 = m(lr).clone()
    g = torch.Generator().manual_seed(5)
    for p in m.parameters():
        if p.dim() >= 2:
            p.data.mul_(0.9).add_(0.01 * torch.randn(p.shape, generator=g).to(p.device) * p.data.abs().mean())
    m.invalidate_packed()
    y1 = m(lr).clone()
    assert (y1 - y0).abs().max().item() > 1e-4                      # the edit is visible
    fresh = getattr(srk, FAMILIES[name][2])(**cfg.as_kwargs()).eval()
    fresh.load_state_dict(m.state_dict(), strict=True)
    assert torch.equal(fresh.cuda()(lr), y1)
    srk.invalidate_packed()                                         # the package-level form used with patched reference containers
    assert torch.equal(m(lr), y1)


@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_model_on_a_non_current_device(name):
    """the binding must launch on the tensors' device and that device's current stream, not on the current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    m0, _ = _build(name, 21, "cuda:0")
    m1, _ = _build(name, 21, "cuda:1")
    lr = synth.make_lr_batch(2, 64, 64, seed=4)
    torch.cuda.set_device(0)
    y0 = m0(lr.to("cuda:0"))
    y1 = m1(lr.to("cuda:1"))                                       # current device is still cuda:0
    assert y1.device == torch.device("cuda:1")
    assert torch.equal(y0.cpu(), y1.cpu())
