"""GPU parity on the 512 px token pyramid (utils/arg_util.py:246-247: patch_nums (1,2,3,4,6,9,13,18,24,32), L = 2240,
32x32 latent) and assorted edge cases (B = 1, label range, argument errors)."""
import numpy as np
import pytest
import torch

from helpers import quant_oracle_of, sd_cpu, var_cfg_of
from oracle import var_oracle as VO
from var_b200 import build_vae_var
from var_b200.init_utils import dense_init_
from var_b200.lib import VarB200Error

pytestmark = pytest.mark.gpu
DEV = "cuda"
PN512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)
_models = {}


def _models512():
    if "m" not in _models:
        vae, var = build_vae_var("cpu", depth=2, patch_nums=PN512)
        dense_init_(vae, seed=3); dense_init_(var, seed=4)
        var.eval(); vae.eval(); var.cond_drop_rate = 0
        _models["m"] = (vae.to(DEV), var.to(DEV))
    return _models["m"]


def test_quantizer_512px_bit_exact():
    vae, var = _models512()
    qo = quant_oracle_of(vae)
    g = torch.Generator().manual_seed(5)
    f = (torch.randn(2, 32, 32, 32, generator=g) * 1.5).numpy()
    ref = qo.f_to_idxBl_or_fhat(f, to_fhat=False)
    got = vae.quantize.f_to_idxBl_or_fhat(torch.from_numpy(f).to(DEV), to_fhat=False)
    assert [tuple(t.shape) for t in got] == [(2, p * p) for p in PN512]
    assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(got, ref))
    vin = vae.quantize.idxBl_to_var_input(got)
    assert vin.shape == (2, 2239, 32) and np.array_equal(vin.cpu().numpy(), qo.idxBl_to_var_input(ref))
    assert np.array_equal(vae.quantize.idxBl_to_fhat(got, last_one=True).cpu().numpy(), qo.idxBl_to_fhat(ref, last_one=True))


def test_var_forward_512px_vs_oracle():
    vae, var = _models512()
    qo = quant_oracle_of(vae)
    g = torch.Generator().manual_seed(6)
    f = (torch.randn(2, 32, 32, 32, generator=g) * 1.5).numpy()
    idx = qo.f_to_idxBl_or_fhat(f, to_fhat=False)
    vin = torch.from_numpy(qo.idxBl_to_var_input(idx))
    labels = torch.tensor([11, 1000])
    ref = VO.var_forward(sd_cpu(var), var_cfg_of(var), labels, vin)
    got = var(labels.to(DEV), vin.to(DEV))
    assert got.shape == (2, 2240, 4096)
    err = (got.cpu() - ref).abs().max().item()
    print(f"512px logits max-abs err {err:.4f}")
    assert err < 6e-2


def test_sampling_512px_runs_and_is_consistent():
    vae, var = _models512()
    labels = torch.tensor([5], device=DEV)
    f_hat, tr = var.autoregressive_infer_cfg(1, labels, g_seed=1, cfg=1.5, top_k=900, top_p=0.95, return_trace=True,
                                             decode=False)
    assert f_hat.shape == (1, 32, 32, 32) and bool(torch.isfinite(f_hat).all())
    again = vae.quantize.idxBl_to_fhat(tr["idx"], last_one=True)
    assert torch.equal(again, f_hat)  # embed_to_fhat identity
    vin = vae.quantize.idxBl_to_var_input(tr["idx"])
    tf = var(labels, vin)
    assert (tf[:, :1] - tr["logits"][0]).abs().max().item() < 2e-2


def test_edge_cases_and_errors():
    from helpers import seeded_models
    vae, var = seeded_models(device=DEV)
    # B = 1, int label, unconditional label (< 0 -> class 1000, var.py:149)
    img = var.autoregressive_infer_cfg(1, 3, g_seed=0, top_k=0, top_p=0.0)
    assert img.shape == (1, 3, 256, 256)
    img = var.autoregressive_infer_cfg(2, -1, g_seed=0, top_k=5)
    assert img.shape == (2, 3, 256, 256)
    img = var.autoregressive_infer_cfg(2, None, g_seed=0, top_k=900, top_p=0.96)  # labels drawn from the generator
    assert bool(torch.isfinite(img).all())
    img = var.autoregressive_infer_cfg(1, 3, g_seed=0, more_smooth=True)  # Gumbel soft-embedding path (var.py:178-180)
    assert img.shape == (1, 3, 256, 256) and bool(torch.isfinite(img).all())
    # CPU parameters: no fallback path
    from var_b200 import build_vae_var as b
    vae_c, var_c = b("cpu", depth=2)
    var_c.eval()
    with pytest.raises(VarB200Error):
        var_c(torch.tensor([1]), torch.zeros(1, 679, 32))
    with pytest.raises(VarB200Error):
        vae_c.quantize.f_to_idxBl_or_fhat(torch.zeros(1, 32, 16, 16), to_fhat=False)
    # label out of range is rejected on the host when the labels live on the CPU (IndexError like nn.Embedding)
    with pytest.raises(IndexError):
        var(torch.tensor([1001, 0, 0]), torch.zeros(3, 679, 32))
