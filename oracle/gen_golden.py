"""ORACLE tooling: generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run here (the reference cannot travel to the GPU box):   python oracle/gen_golden.py
Recipe: SURVEY.md Appendix B (torch.Optional shim, undo the process-wide reset_parameters patch, seeded dense init,
eval mode, cond_drop_rate = 0, TF32 off). The same `var_b200.init_utils.dense_init_` is applied to the reference
modules here and to our modules in the tests, so both sides hold bit-identical weights without storing them.
"""
from __future__ import annotations

import json
import os
import sys
import typing
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
REF = "/root/reference"
OUT = ROOT / "tests" / "golden"

from var_b200.init_utils import dense_init_  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    torch.Optional = typing.Optional  # models/var.py:241-242 needs it (absent in torch 2.11)
    saved = {c: c.reset_parameters for c in (nn.Linear, nn.LayerNorm, nn.BatchNorm2d, nn.SyncBatchNorm, nn.Conv1d,
                                             nn.Conv2d, nn.ConvTranspose1d, nn.ConvTranspose2d)}
    import models  # noqa
    from models import build_vae_var
    import models.var as ref_var
    import models.helpers as ref_helpers

    def build(**kw):
        vae, var = build_vae_var(device="cpu", flash_if_available=False, fused_if_available=False, **kw)
        for c, f in saved.items():
            c.reset_parameters = f  # undo models/__init__.py:24-25
        var.eval(); vae.eval(); var.cond_drop_rate = 0
        return vae, var
    return build, ref_var, ref_helpers


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    build, ref_var, ref_helpers = import_reference()

    # ---------------------------------------------------------------- state_dict contracts
    shapes = {}
    for depth, shared in ((2, False), (2, True)):
        vae, var = build(depth=depth, shared_aln=shared)
        shapes[f"var_d{depth}_shared{int(shared)}"] = {k: list(v.shape) for k, v in var.state_dict().items()}
        if not shared:
            shapes["vqvae"] = {k: list(v.shape) for k, v in vae.state_dict().items()}
    (OUT / "state_dict_shapes.json").write_text(json.dumps(shapes, indent=0, sort_keys=True))

    # ---------------------------------------------------------------- G1/G2: quantizer + teacher-forced forward (depth 2)
    vae, var = build(depth=2)
    dense_init_(vae, seed=1); dense_init_(var, seed=2)
    q = vae.quantize
    g = torch.Generator().manual_seed(11)
    B = 3
    f = (torch.randn(B, 32, 16, 16, generator=g) * 1.5).contiguous()
    with torch.no_grad():
        idx = q.f_to_idxBl_or_fhat(f, to_fhat=False)
        fhats = q.f_to_idxBl_or_fhat(f, to_fhat=True)
        var_in = q.idxBl_to_var_input(idx)
        # fp64 margin of every argmin (near-tie report): recompute distances in double
        f_rest = f.double().clone(); margins = []
        E = q.embedding.weight.double()
        # custom patch grid incl. non-square scales (quant.py:143)
        vpn = [(1, 1), (2, 3), (4, 4), (5, 8), (16, 16)]
        idx_ns = q.f_to_idxBl_or_fhat(f[:2], to_fhat=False, v_patch_nums=vpn)
        labels = torch.tensor([3, 999, 1000])
        logits = var(labels, var_in)
        gt = torch.cat(idx, dim=1)
        logp = torch.log_softmax(logits, dim=-1).gather(-1, gt.unsqueeze(-1)).squeeze(-1)
        # per-block activations via forward hooks
        acts = []
        hooks = [b.register_forward_hook(lambda m, i, o: acts.append(o.detach().clone())) for b in var.blocks]
        var(labels, var_in)
        for h in hooks:
            h.remove()
        # CFG-mixed scoring exactly as var_analysis.py:320-346,437-466 does it around VAR.forward
        pn = var.patch_nums
        unc = var(torch.tensor(1000), var_in[:1])
        ratio = torch.cat([torch.full((p * p,), si / (len(pn) - 1)) for si, p in enumerate(pn)])
        t_cfg = 1.5 * ratio.unsqueeze(0).unsqueeze(-1)
        lg_c = var(torch.tensor([3, 999, 17]), var_in[:1].expand(3, -1, -1))
        mixed = (1 + t_cfg) * lg_c - t_cfg * unc
        glp = torch.nn.functional.log_softmax(mixed, dim=-1).gather(-1, gt[:1].expand(3, -1).unsqueeze(-1)).squeeze(-1)
        cfg_scale_sums, st_ = [], 0
        for p in pn:
            cfg_scale_sums.append(glp[:, st_:st_ + p * p].sum(-1)); st_ += p * p
        cfg_scale_sums = torch.stack(cfg_scale_sums, 1)
        # embed_to_fhat / idxBl_to_img path (decoder output pins the boundary function)
        img = vae.idxBl_to_img(idx, same_shape=True, last_one=True)
        f_enc = vae.quant_conv(vae.encoder(torch.rand(1, 3, 256, 256, generator=g) * 2 - 1))
    np.savez_compressed(
        OUT / "quant_forward_d2.npz",
        f=f.numpy(), idx=np.concatenate([i.numpy().astype(np.int16) for i in idx], axis=1),
        fhat_last=fhats[-1].numpy(), fhat_s3=fhats[3].numpy(), var_input=var_in.numpy(),
        idx_nonsquare=np.concatenate([i.numpy().astype(np.int16) for i in idx_ns], axis=1),
        labels=labels.numpy(), logits_sub=logits[:, ::23, ::29].numpy(), lse=torch.logsumexp(logits, -1).numpy(),
        logp=logp.numpy(), scores=logp.sum(1).numpy(),
        block_sub=np.stack([a[:, ::23, ::7].numpy() for a in acts]),
        img_sub=img[:, :, ::8, ::8].numpy(), f_enc_sub=f_enc[:, :, ::2, ::2].numpy(),
        cfg_scale_sums=cfg_scale_sums.numpy(), cfg_tok_logp=glp.numpy())

    # ---------------------------------------------------------------- G2b: --mode l2_dist scores (var_analysis.py:252-256,468-524)
    # computed from the reference model's own logits with the script's statements (the script has no importable function)
    with torch.no_grad():
        dists_ref = torch.cdist(q.embedding.weight, q.embedding.weight, p=2)
        probs_ref = torch.nn.functional.softmax(mixed, dim=-1)
        gt_d_ref = dists_ref[gt[:1].expand(3, -1)]
        avg_all = (gt_d_ref * probs_ref).sum(dim=-1)
        tkp, tki = torch.topk(probs_ref, k=50, dim=-1)
        tkd = torch.gather(gt_d_ref, dim=-1, index=tki)
        tkp = tkp / tkp.sum(dim=-1, keepdim=True)
        avg_k50 = (tkd * tkp).sum(dim=-1)
        probs_nocfg = torch.nn.functional.softmax(lg_c, dim=-1)
        avg_nocfg = (gt_d_ref * probs_nocfg).sum(dim=-1)
    np.savez_compressed(OUT / "l2dist_d2.npz", neg_all=(-avg_all).numpy(), neg_k50=(-avg_k50).numpy(),
                        neg_nocfg=(-avg_nocfg).numpy())

    # ---------------------------------------------------------------- G3: sampler op (helpers.py:6-19) with replayed noise
    gs = torch.Generator().manual_seed(5)
    lg = (torch.randn(2, 9, 4096, generator=gs) * 2.0)
    lg[0, 0, 100:110] = lg[0, 0, 100]  # exact ties inside the row
    out = {}
    for name, (k, p) in dict(k900=(900, 0.0), k900p95=(900, 0.95), k0=(0, 0.0), p50=(0, 0.5)).items():
        rng = torch.Generator().manual_seed(77)
        tok = ref_helpers.sample_with_top_k_top_p_(lg.clone(), top_k=k, top_p=p, rng=rng, num_samples=1)[:, :, 0]
        out["tok_" + name] = tok.numpy().astype(np.int16)
    rng = torch.Generator().manual_seed(77)
    qn = torch.empty(18, 4096).exponential_(1, generator=rng)
    np.savez_compressed(OUT / "sampler.npz", logits=lg.numpy(), q=qn.numpy(), **out)

    # ---------------------------------------------------------------- G4: KV-cached CFG sampling (var.py:126-190), depth 2
    rec = []
    orig = ref_var.sample_with_top_k_top_p_

    def spy(logits_BlV, **kw):
        r = orig(logits_BlV, **kw)
        rec.append((logits_BlV.detach().clone(), r[:, :, 0].clone()))  # logits AFTER in-place masking, tokens
        return r
    ref_var.sample_with_top_k_top_p_ = spy
    labels_ar = torch.tensor([7, 481])
    with torch.no_grad():
        img_ar = var.autoregressive_infer_cfg(B=2, label_B=labels_ar, g_seed=123, cfg=1.5, top_k=900, top_p=0.0)
    ref_var.sample_with_top_k_top_p_ = orig
    ar_idx = np.concatenate([t.numpy().astype(np.int16) for _, t in rec], axis=1)
    with torch.no_grad():  # identity: teacher-forced logits == per-scale cached logits
        ms = [t for _, t in rec]
        f_hat_ar = q.embed_to_fhat([q.embedding(t).transpose(1, 2).reshape(2, 32, pn, pn) for t, pn in zip(ms, var.patch_nums)],
                                   all_to_max_scale=True, last_one=True)
    np.savez_compressed(OUT / "ar_d2.npz", labels=labels_ar.numpy(), idx=ar_idx, f_hat=f_hat_ar.numpy(),
                        img_sub=img_ar[:, :, ::8, ::8].numpy(),
                        logit_max=np.stack([lg_.amax(-1).numpy().reshape(-1)[:2] for lg_, _ in rec]),
                        kept=np.array([int(torch.isfinite(lg_).sum()) for lg_, _ in rec]))
    # ---------------------------------------------------------------- G5: VAR.inpainting (var.py:236-364), depth 2
    # keep mask: scales 0-2 entirely (exercises the skip branch, which consumes no noise), then the left half of
    # every later token map. gt tokens = the quantizer's tokens of the random feature map above.
    gt_tok = torch.cat(idx, dim=1)[:2].clone()
    keep = torch.zeros_like(gt_tok, dtype=torch.bool)
    off = 0
    for si_, pn_ in enumerate(var.patch_nums):
        m_ = torch.zeros(pn_, pn_, dtype=torch.bool)
        if si_ <= 2:
            m_[:] = True
        else:
            m_[:, : pn_ // 2] = True
        keep[:, off:off + pn_ * pn_] = m_.reshape(-1)
        off += pn_ * pn_
    rec2 = []

    def spy2(logits_BlV, **kw):
        r = orig(logits_BlV, **kw)
        rec2.append((logits_BlV.detach().clone(), r[:, :, 0].clone()))
        return r
    ref_var.sample_with_top_k_top_p_ = spy2
    emb_calls = []
    emb_orig = q.embedding.forward

    def emb_spy(t):
        emb_calls.append(t.detach().clone())
        return emb_orig(t)
    q.embedding.forward = emb_spy
    with torch.no_grad():
        img_inp = var.inpainting(gt_tok, keep, label=labels_ar, g_seed=321, cfg=1.5, top_k=900, top_p=0.0)
    q.embedding.forward = emb_orig
    ref_var.sample_with_top_k_top_p_ = orig
    final_tok = np.concatenate([t.numpy().astype(np.int16) for t in emb_calls], axis=1)
    with torch.no_grad():
        f_hat_inp = q.embed_to_fhat([q.embedding(t).transpose(1, 2).reshape(2, 32, pn, pn) for t, pn in zip(emb_calls, var.patch_nums)],
                                    all_to_max_scale=True, last_one=True)
    np.savez_compressed(OUT / "inpaint_d2.npz", labels=labels_ar.numpy(), gt_tokens=gt_tok.numpy().astype(np.int16),
                        keep=keep.numpy(), final_tokens=final_tok, f_hat=f_hat_inp.numpy(),
                        img_sub=img_inp[:, :, ::8, ::8].numpy(), n_sampled_scales=np.array(len(rec2)),
                        logit_max=np.stack([lg_.amax(-1).numpy().reshape(-1)[:2] for lg_, _ in rec2]))
    # ---------------------------------------------------------------- G6: VAR.smooth_sampling (var.py:366-575), depth 2
    sm = {}
    for tag, kw in (("cnt", dict(n=8)), ("thr", dict(n=8, neighbor_threshold=0.9))):
        emb_calls = []
        q.embedding.forward = lambda t, _c=emb_calls: (_c.append(t.detach().clone()), emb_orig(t))[1]
        with torch.no_grad():
            img_s, sll, sdll = var.smooth_sampling(gt_tok, label=labels_ar, g_seed=1, cfg=1.5, **kw)
        q.embedding.forward = emb_orig
        sm[f"tok_{tag}"] = np.concatenate([t.numpy().astype(np.int16) for t in emb_calls], axis=1)
        sm[f"sum_ll_{tag}"] = np.array(float(sll))
        sm[f"sum_dll_{tag}"] = np.array(float(sdll))
        sm[f"img_sub_{tag}"] = img_s[:, :, ::8, ::8].numpy()
    np.savez_compressed(OUT / "smooth_d2.npz", labels=labels_ar.numpy(), gt_tokens=gt_tok.numpy().astype(np.int16), **sm)
    # ---------------------------------------------------------------- G7: more_smooth sampling (var.py:178-180), depth 2
    with torch.no_grad():
        f_hats = []
        gni = q.get_next_autoregressive_input

        def gni_spy(si_, SN_, f_hat_, h_):
            r = gni(si_, SN_, f_hat_, h_)
            if si_ == SN_ - 1:
                f_hats.append(r[0].detach().clone())
            return r
        q.get_next_autoregressive_input = gni_spy
        img_ms = var.autoregressive_infer_cfg(B=2, label_B=labels_ar, g_seed=77, cfg=1.5, top_k=900, top_p=0.0, more_smooth=True)
        q.get_next_autoregressive_input = gni
    np.savez_compressed(OUT / "more_smooth_d2.npz", labels=labels_ar.numpy(), f_hat=f_hats[0].numpy(),
                        img_sub=img_ms[:, :, ::8, ::8].numpy())
    print("golden fixtures written to", OUT)
    for p in sorted(OUT.iterdir()):
        print(f"  {p.name}: {p.stat().st_size / 1024:.0f} KB")


if __name__ == "__main__":
    main()
