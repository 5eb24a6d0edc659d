"""CPU tier: the oracle against the reference's own outputs at the HEADLINE depth (VAR-d16) and at BASELINE
configs[1] (B=64 tokenisation). Fixtures: oracle/gen_golden_d16.py (the unmodified reference, run in the build
container). The GPU tier (tests/test_parity_d16_gpu.py) compares the kernels with the same fixtures."""
import sys
from pathlib import Path

import numpy as np
import torch

from helpers import golden, quant_oracle_of, replay_noise, sd_cpu, seeded_models, split_scales, var_cfg_of
from oracle import var_oracle as VO

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))


def _inputs():
    from gen_golden_d16 import quant_b64_inputs, score_image_f
    return quant_b64_inputs, score_image_f


def test_quant_oracle_b64_matches_reference_golden():
    """BASELINE configs[1]: 2 x 43 520 indices, sigma in {1, 3}: the C oracle equals the reference index for index."""
    g = golden("quant_b64.npz")
    quant_b64_inputs, _ = _inputs()
    vae, _ = seeded_models()
    qo = quant_oracle_of(vae)
    for sigma in (1.0, 3.0):
        got = np.concatenate(qo.f_to_idxBl_or_fhat(quant_b64_inputs(sigma).numpy(), to_fhat=False), axis=1)
        ref = g[f"idx_s{int(sigma)}"].astype(np.int64)
        assert got.shape == (64, 680)
        n_bad = int((got != ref).sum())
        assert n_bad == 0, f"sigma={sigma}: {n_bad} of 43520 indices differ from the reference"


def test_var_oracle_d16_forward_and_scores_match_reference_golden():
    g = golden("d16_forward.npz")
    _, score_image_f = _inputs()
    vae, var = seeded_models(depth=16)
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    qo = quant_oracle_of(vae)
    idx_np = qo.f_to_idxBl_or_fhat(score_image_f().numpy(), to_fhat=False)
    assert np.array_equal(np.concatenate(idx_np, axis=1), g["idx"].astype(np.int64))
    vin = torch.from_numpy(qo.idxBl_to_var_input(idx_np))
    labels = torch.from_numpy(g["labels"])
    logits, acts = VO.var_forward(sd, cfg, labels, vin.expand(3, -1, -1), return_blocks=True)
    assert (logits[:, ::7, ::29] - torch.from_numpy(g["logits_sub"])).abs().max().item() < 1e-3
    assert (torch.logsumexp(logits, -1) - torch.from_numpy(g["lse"])).abs().max().item() < 1e-3
    for j, bi in enumerate((0, 7, 15)):
        ref = torch.from_numpy(g["block_sub"][j])
        assert (acts[bi][:, ::7, ::5] - ref).abs().max().item() < 1e-3 * max(1.0, float(g["block_absmax"][j]))
    # 8 of the 40 scored classes (the rest is covered on the GPU tier): the oracle's score restatement at d16
    lab = torch.from_numpy(g["score_labels"][-8:])
    lg = VO.var_forward(sd, cfg, lab, vin.expand(8, -1, -1))
    sc = VO.class_scores(lg, torch.from_numpy(g["idx"].astype(np.int64)))
    assert (sc - torch.from_numpy(g["scores"][-8:])).abs().max().item() < 0.05  # |score| ~ 6000, fp32 sums of 680 terms


def test_var_oracle_d16_ar_matches_reference_golden():
    """KV-cached CFG sampling at d16, B=2: the oracle, fed the same Exp(1) noise, draws the reference's 1360 tokens."""
    g = golden("d16_ar.npz")
    vae, var = seeded_models(depth=16)
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    out = VO.ar_infer(sd, cfg, quant_oracle_of(vae), torch.from_numpy(g["labels"]), replay_noise(1234, B=2), cfg_scale=1.5,
                      top_k=900)
    got = np.concatenate([i.numpy() for i in out["idx"]], axis=1)
    assert np.array_equal(got, g["idx"].astype(np.int64))
    mixed = torch.cat([lg[:, :, ::29] for lg in out["logits"]], dim=1)
    assert (mixed - torch.from_numpy(g["mixed_sub"])).abs().max().item() < 2e-3
    assert (out["f_hat"] - torch.from_numpy(g["f_hat"])).abs().max().item() < 5e-5
