"""ORACLE tooling: headline-depth golden vectors from the UNMODIFIED reference (/root/reference) on CPU.

Run here (the reference cannot travel to the GPU box):   python oracle/gen_golden_d16.py
Same recipe as oracle/gen_golden.py; the model is VAR-d16 (C=1024, 16 heads, 16 blocks), the
depth BASELINE.json configs[2]/[3] are quoted on. Both sides hold the same weights through the per-name seeded
`dense_init_` (vae seed 1, var seed 2), so only inputs' seeds and sub-sampled outputs are stored.

Writes tests/golden/d16_forward.npz  (teacher-forced logits / lse / per-block activations, 40-class scores of one image)
       tests/golden/d16_ar.npz       (KV-cached CFG sampling, B=2: the reference's tokens and per-scale mixed logits)
       tests/golden/quant_b64.npz    (BASELINE configs[1]: 2 x 43 520 indices of f ~ N(0, sigma^2), sigma in {1, 3})
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
OUT = ROOT / "tests" / "golden"

from gen_golden import import_reference  # noqa: E402
from var_b200.init_utils import dense_init_  # noqa: E402

# the candidate classes of the scoring fixture: 39 labels spread over [0, 1000) plus the unconditional label 1000
SCORE_LABELS = sorted({(37 * i + 11) % 1000 for i in range(39)}) + [1000]
BLOCKS_KEPT = (0, 7, 15)


def quant_b64_inputs(sigma: float) -> torch.Tensor:
    """BASELINE configs[1] / SURVEY 8d Config 2: f ~ N(0, sigma^2) [64,32,16,16] (shared with the tests)."""
    g = torch.Generator().manual_seed(6400 + int(sigma * 10))
    return (torch.randn(64, 32, 16, 16, generator=g) * sigma).contiguous()


def score_image_f() -> torch.Tensor:
    g = torch.Generator().manual_seed(1601)
    return (torch.randn(1, 32, 16, 16, generator=g) * 1.5).contiguous()


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    torch.set_num_threads(os.cpu_count() or 8)
    OUT.mkdir(parents=True, exist_ok=True)
    build, ref_var, _ = import_reference()
    t0 = time.time()
    vae, var = build(depth=16)
    dense_init_(vae, seed=1); dense_init_(var, seed=2)
    q = vae.quantize
    print(f"reference d16 built in {time.time() - t0:.0f} s")

    # ---------------------------------------------------------------- config 2: B=64 tokenisation, sigma in {1, 3}
    out = {}
    with torch.no_grad():
        for sigma in (1.0, 3.0):
            idx = q.f_to_idxBl_or_fhat(quant_b64_inputs(sigma), to_fhat=False)
            out[f"idx_s{int(sigma)}"] = np.concatenate([i.numpy().astype(np.int16) for i in idx], axis=1)
    np.savez_compressed(OUT / "quant_b64.npz", **out)

    # ---------------------------------------------------------------- teacher-forced forward + 40-class scores (d16)
    with torch.no_grad():
        idx = q.f_to_idxBl_or_fhat(score_image_f(), to_fhat=False)
        var_in = q.idxBl_to_var_input(idx)
        gt = torch.cat(idx, dim=1)
        labels = torch.tensor([3, 999, 1000])
        acts = []
        hooks = [var.blocks[i].register_forward_hook(lambda m, i_, o: acts.append(o.detach().clone())) for i in BLOCKS_KEPT]
        t0 = time.time()
        logits = var(labels, var_in.expand(3, -1, -1))
        print(f"d16 forward B=3: {time.time() - t0:.1f} s, logits std {logits.std():.3f} absmax {logits.abs().max():.2f}")
        for h in hooks:
            h.remove()
        lab_s = torch.tensor(SCORE_LABELS)
        scores, lse_s = [], []
        for lo in range(0, len(SCORE_LABELS), 8):  # eval_prob.py:436-463, eight classes per forward
            lg = var(lab_s[lo:lo + 8], var_in.expand(len(lab_s[lo:lo + 8]), -1, -1))
            lp = torch.log_softmax(lg, dim=-1).gather(-1, gt.expand(lg.shape[0], -1).unsqueeze(-1)).squeeze(-1)
            scores.append(lp.sum(1)); lse_s.append(torch.logsumexp(lg, -1)[:, ::17])
        scores = torch.cat(scores)
    order = torch.argsort(scores, descending=True)
    print("scores: top-5 labels", lab_s[order[:5]].tolist(), "values", scores[order[:5]].tolist(),
          "spread", float(scores.max() - scores.min()))
    np.savez_compressed(
        OUT / "d16_forward.npz", idx=gt.numpy().astype(np.int16), labels=labels.numpy(),
        logits_sub=logits[:, ::7, ::29].numpy(), lse=torch.logsumexp(logits, -1).numpy(),
        block_sub=np.stack([a[:, ::7, ::5].numpy() for a in acts]), block_absmax=np.array([float(a.abs().max()) for a in acts]),
        score_labels=lab_s.numpy(), scores=scores.numpy(), score_lse_sub=torch.cat(lse_s).numpy())

    # ---------------------------------------------------------------- KV-cached CFG sampling B=2 (var.py:126-190), d16
    rec = []
    orig = ref_var.sample_with_top_k_top_p_

    def spy(logits_BlV, **kw):
        before = logits_BlV.detach().clone()  # CFG-mixed logits before the in-place top-k mask (var.py:173-175)
        r = orig(logits_BlV, **kw)
        rec.append((before, r[:, :, 0].clone()))
        return r
    ref_var.sample_with_top_k_top_p_ = spy
    labels_ar = torch.tensor([207, 980])
    with torch.no_grad():
        t0 = time.time()
        img = var.autoregressive_infer_cfg(B=2, label_B=labels_ar, g_seed=1234, cfg=1.5, top_k=900, top_p=0.0)
        print(f"d16 AR B=2: {time.time() - t0:.1f} s")
    ref_var.sample_with_top_k_top_p_ = orig
    with torch.no_grad():
        ms = [t for _, t in rec]
        f_hat = q.embed_to_fhat([q.embedding(t).transpose(1, 2).reshape(2, 32, pn, pn) for t, pn in zip(ms, var.patch_nums)],
                                all_to_max_scale=True, last_one=True)
    np.savez_compressed(
        OUT / "d16_ar.npz", labels=labels_ar.numpy(), idx=np.concatenate([t.numpy().astype(np.int16) for _, t in rec], axis=1),
        f_hat=f_hat.numpy(), img_sub=img[:, :, ::8, ::8].numpy(),
        mixed_sub=np.concatenate([lg[:, :, ::29].numpy() for lg, _ in rec], axis=1),  # [2, 680, 142]
        mixed_lse=np.concatenate([torch.logsumexp(lg, -1).numpy() for lg, _ in rec], axis=1))
    for n in ("quant_b64.npz", "d16_forward.npz", "d16_ar.npz"):
        print(f"  {n}: {(OUT / n).stat().st_size / 1024:.0f} KB")


if __name__ == "__main__":
    main()
