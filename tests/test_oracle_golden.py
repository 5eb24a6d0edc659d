"""CPU tier: pin the oracle (oracle/) to the reference through the committed golden vectors, and check the host-side
contracts (state_dict keys, C-ABI symbols). No GPU, no /root/reference needed."""
import json

import numpy as np
import pytest
import torch

from helpers import (replay_noise_more_smooth, GOLDEN, PATCH_NUMS, golden, quant_oracle_of, replay_noise, sd_cpu, seeded_models, split_scales,
                     var_cfg_of)
from oracle import var_oracle as VO


def test_state_dict_keys_and_shapes_match_reference():
    ref = json.loads((GOLDEN / "state_dict_shapes.json").read_text())
    for shared in (False, True):
        vae, var = seeded_models(depth=2, shared_aln=shared)
        ours = {k: list(v.shape) for k, v in var.state_dict().items()}
        assert ours == ref[f"var_d2_shared{int(shared)}"]
    ours = {k: list(v.shape) for k, v in vae.state_dict().items()}
    assert ours == ref["vqvae"]


def test_capi_exports_every_declared_symbol():
    from var_b200 import lib as L
    lib = L.load()
    syms = L.exported_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.var_b200_last_error() is not None
    assert lib.var_b200_gemm_tile_n(1920) == 192 and lib.var_b200_gemm_tile_n(4096) == 256


def test_binding_declares_the_signature_of_every_entry_point():
    """Every function include/var_b200.h declares has its ctypes signature declared by var_b200/lib.py with as many
    arguments as the header's prototype (a C entry point bound without argtypes would truncate 64-bit pointers /
    sizes silently)."""
    import re
    from pathlib import Path
    from var_b200 import lib as L
    lib = L.load()
    hdr = (Path(__file__).resolve().parent.parent / "include" / "var_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"VAR_B200_API[^;(]*?\b(var_b200_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S))
    assert set(protos) == set(L.exported_symbols())
    for name, args in protos.items():
        args = args.strip()
        n_args = 0 if args in ("", "void") else len(args.split(","))
        fn = getattr(lib, name)
        assert fn.argtypes is not None, f"{name}: no argtypes declared in var_b200/lib.py"
        assert len(fn.argtypes) == n_args, f"{name}: header has {n_args} arguments, the binding {len(fn.argtypes)}"


def test_custom_op_layer_registers_cuda_only_ops():
    """var_b200/ops.py: every C entry point of the path is a `var_b200::` torch.library op with a CUDA kernel only;
    CPU tensors are refused by the dispatcher (no fallback)."""
    import var_b200  # noqa: F401
    from var_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.var_b200, name), name
    x = torch.randn(4, 64)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.var_b200.ln_modulate(x, torch.zeros(1, 64), torch.zeros(1, 64), 64, 4, 1e-6)


def test_header_constants_and_struct_layouts_match_the_binding():
    """The ctypes mirror must agree with include/var_b200.h and the kernels: constants, field order of the structs."""
    import ctypes as C
    import re
    from pathlib import Path
    from var_b200 import lib as L
    root = Path(__file__).resolve().parent.parent
    hdr = (root / "include" / "var_b200.h").read_text()
    gemm_h = (root / "var_b200" / "csrc" / "gemm.h").read_text()
    assert int(re.search(r"#define VAR_B200_GEMM_EPI_PARTS (\d+)", hdr).group(1)) == L.GEMM_EPI_PARTS
    assert int(re.search(r"constexpr int GEMM_EPI_SUB = (\d+);", gemm_h).group(1)) == L.GEMM_EPI_PARTS
    assert int(re.search(r"#define VAR_B200_MAX_SCALES (\d+)", hdr).group(1)) == L.MAX_SCALES

    def fields_of(struct_name):
        body = re.search(r"typedef struct " + struct_name + r" \{(.*?)\} ", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.search(r"(\w+)\s*(\[[^\]]*\])?\s*$", part.strip()).group(1))
        return names

    assert fields_of("var_b200_gemm_args") == [f[0] for f in L.GemmArgs._fields_]
    assert fields_of("var_b200_model") == [f[0] for f in L.ModelDesc._fields_]
    assert fields_of("var_b200_quant") == [f[0] for f in L.QuantDesc._fields_]
    assert fields_of("var_b200_block_weights") == [f[0] for f in L.BlockWeights._fields_]
    # the attention entry point takes (.., level_end, max_score, q_log2, stream)
    assert L.load().var_b200_attention.argtypes[-3:] == [C.c_float, C.c_int, C.c_void_p]


def test_quant_oracle_indices_match_reference_golden():
    g = golden("quant_forward_d2.npz")
    vae, _ = seeded_models()
    qo = quant_oracle_of(vae)
    idx = qo.f_to_idxBl_or_fhat(g["f"], to_fhat=False)
    got = np.concatenate(idx, axis=1)
    mism = np.argwhere(got != g["idx"].astype(np.int64))
    assert mism.size == 0, f"{len(mism)} index mismatches vs reference, first at {mism[:5].tolist()}"
    fh = qo.f_to_idxBl_or_fhat(g["f"], to_fhat=True)
    assert np.abs(fh[-1] - g["fhat_last"]).max() < 2e-5
    assert np.abs(fh[3] - g["fhat_s3"]).max() < 2e-5


def test_quant_oracle_nonsquare_patch_grid():
    g = golden("quant_forward_d2.npz")
    vae, _ = seeded_models()
    vpn = [(1, 1), (2, 3), (4, 4), (5, 8), (16, 16)]
    idx = quant_oracle_of(vae).f_to_idxBl_or_fhat(g["f"][:2], to_fhat=False, v_patch_nums=vpn)
    assert np.array_equal(np.concatenate(idx, axis=1), g["idx_nonsquare"].astype(np.int64))
    with pytest.raises(AssertionError):
        quant_oracle_of(vae).f_to_idxBl_or_fhat(g["f"][:2], to_fhat=False, v_patch_nums=[1, 2, 8])


def test_quant_oracle_var_input_and_fhat_decode():
    g = golden("quant_forward_d2.npz")
    vae, _ = seeded_models()
    qo = quant_oracle_of(vae)
    idx = split_scales(g["idx"])
    assert np.abs(qo.idxBl_to_var_input(idx) - g["var_input"]).max() < 2e-5
    assert np.abs(qo.idxBl_to_fhat(idx, last_one=True) - g["fhat_last"]).max() < 2e-5
    # step-wise decode == whole decode (quant.py:187-196 vs :169-184)
    f_hat = np.zeros_like(g["fhat_last"])
    nxts = []
    for si in range(len(PATCH_NUMS)):
        nxt = qo.get_next_autoregressive_input(si, f_hat, idx[si])
        if nxt is not None:
            nxts.append(nxt.reshape(nxt.shape[0], 32, -1).transpose(0, 2, 1))
    assert np.array_equal(np.concatenate(nxts, axis=1), qo.idxBl_to_var_input(idx))
    assert np.array_equal(f_hat, qo.idxBl_to_fhat(idx, last_one=True))


def test_var_oracle_forward_matches_reference_golden():
    g = golden("quant_forward_d2.npz")
    _, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    logits, acts = VO.var_forward(sd, cfg, torch.from_numpy(g["labels"]), torch.from_numpy(g["var_input"]), return_blocks=True)
    assert (logits[:, ::23, ::29] - torch.from_numpy(g["logits_sub"])).abs().max() < 5e-4
    assert (torch.logsumexp(logits, -1) - torch.from_numpy(g["lse"])).abs().max() < 5e-4
    gt = torch.from_numpy(g["idx"].astype(np.int64))
    logp = VO.token_log_probs(logits, gt)
    assert (logp - torch.from_numpy(g["logp"])).abs().max() < 1e-3
    assert (logp.sum(1) - torch.from_numpy(g["scores"])).abs().max() < 2e-2
    for i, a in enumerate(acts):
        ref = torch.from_numpy(g["block_sub"][i])
        assert (a[:, ::23, ::7] - ref).abs().max() < 1e-3 * max(1.0, ref.abs().max().item()), f"block {i}"
    # the golden weights are dense: the transformer body must matter
    assert float(np.abs(g["block_sub"][1] - g["block_sub"][0]).max()) > 0.1


def test_cfg_scoring_oracle_matches_reference_golden():
    """var_analysis.py:320-346,437-466: CFG-mixed teacher-forced per-scale log-likelihoods."""
    g = golden("quant_forward_d2.npz")
    _, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    vin = torch.from_numpy(g["var_input"][:1])
    gt = torch.from_numpy(g["idx"][:1].astype(np.int64))
    unc = VO.var_forward(sd, cfg, torch.tensor([1000]), vin)
    lc = VO.var_forward(sd, cfg, torch.tensor([3, 999, 17]), vin.expand(3, -1, -1))
    total, per_scale, tok = VO.cfg_class_scores(lc, unc, gt, 1.5, PATCH_NUMS)
    assert (tok - torch.from_numpy(g["cfg_tok_logp"])).abs().max() < 2e-3
    assert (per_scale - torch.from_numpy(g["cfg_scale_sums"])).abs().max() < 2e-2
    assert (total - per_scale.sum(1)).abs().max() < 1e-2


def test_sampler_oracle_matches_reference_golden():
    g = golden("sampler.npz")
    lg, q = torch.from_numpy(g["logits"]), torch.from_numpy(g["q"])
    for name, (k, p) in dict(k900=(900, 0.0), k900p95=(900, 0.95), k0=(0, 0.0), p50=(0, 0.5)).items():
        tok = VO.sample_top_k_top_p(lg, q, top_k=k, top_p=p)
        assert torch.equal(tok, torch.from_numpy(g["tok_" + name].astype(np.int64))), name


def test_ar_oracle_matches_reference_golden():
    g = golden("ar_d2.npz")
    vae, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    noise = replay_noise(123, B=2)
    out = VO.ar_infer(sd, cfg, quant_oracle_of(vae), torch.from_numpy(g["labels"]), noise, cfg_scale=1.5, top_k=900)
    got = torch.cat(out["idx"], dim=1).numpy()
    ref = g["idx"].astype(np.int64)
    assert (got != ref).sum() == 0, f"{(got != ref).sum()} sampled tokens differ from the reference"
    assert (out["f_hat"] - torch.from_numpy(g["f_hat"])).abs().max() < 5e-5


def test_inpainting_oracle_matches_reference_golden():
    """VAR.inpainting (var.py:236-364): kept tokens override the samples, fully kept scales skip the sampler."""
    g = golden("inpaint_d2.npz")
    vae, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    keep = torch.from_numpy(g["keep"])
    gt = torch.from_numpy(g["gt_tokens"].astype(np.int64))
    assert int(g["n_sampled_scales"]) == 7
    noise = replay_noise(321, B=2, skip_scales=(0, 1, 2))
    out = VO.ar_infer(sd, cfg, quant_oracle_of(vae), torch.from_numpy(g["labels"]), noise, cfg_scale=1.5, top_k=900,
                      gt_tokens=gt, keep_mask=keep)
    got = torch.cat(out["idx"], dim=1).numpy()
    ref = g["final_tokens"].astype(np.int64)
    assert (got != ref).sum() == 0, f"{(got != ref).sum()} tokens differ from the reference"
    assert (got[g["keep"]] == g["gt_tokens"].astype(np.int64)[g["keep"]]).all()
    assert all(l is None for l in out["logits"][:3]) and all(l is not None for l in out["logits"][3:])
    assert (out["f_hat"] - torch.from_numpy(g["f_hat"])).abs().max() < 5e-5


def test_expected_dist_oracle_matches_reference_golden():
    """var_analysis.py --mode l2_dist (:252-256,468-524): expected codebook distance scores, full and top-k renormalised."""
    g, g2 = golden("quant_forward_d2.npz"), golden("l2dist_d2.npz")
    vae, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    vin = torch.from_numpy(g["var_input"][:1])
    gt = torch.from_numpy(g["idx"][:1].astype(np.int64))
    E = vae.quantize.embedding.weight.detach()
    unc = VO.var_forward(sd, cfg, torch.tensor([1000]), vin)
    lc = VO.var_forward(sd, cfg, torch.tensor([3, 999, 17]), vin.expand(3, -1, -1))
    for key, kw, c in (("neg_all", {}, 1.5), ("neg_k50", dict(top_k=50), 1.5), ("neg_nocfg", {}, 0.0)):
        total, per_scale, tok = VO.expected_dist_scores(lc, unc, gt, c, PATCH_NUMS, E, **kw)
        assert (tok - torch.from_numpy(g2[key])).abs().max() < 2e-3, key
        assert (total - per_scale.sum(1)).abs().max() < 1e-2


@pytest.mark.parametrize("tag,kw", [("cnt", {}), ("thr", dict(neighbor_threshold=0.9))])
def test_smooth_sampling_oracle_matches_reference_golden(tag, kw):
    """VAR.smooth_sampling (var.py:366-575): neighbour-restricted arg-max tokens and both accumulated likelihoods."""
    g = golden("smooth_d2.npz")
    vae, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    gt = torch.from_numpy(g["gt_tokens"].astype(np.int64))
    out = VO.smooth_infer(sd, cfg, quant_oracle_of(vae), torch.from_numpy(g["labels"]), gt, 8,
                          vae.quantize.embedding.weight.detach(), cfg_scale=1.5, **kw)
    got = torch.cat(out["idx"], dim=1).numpy()
    ref = g[f"tok_{tag}"].astype(np.int64)
    assert (got != ref).sum() == 0, f"{(got != ref).sum()} tokens differ from the reference"
    assert float(out["sum_ll"]) == float(g[f"sum_ll_{tag}"])          # integer-truncated terms (var.py:536)
    assert abs(float(out["sum_dll"]) - float(g[f"sum_dll_{tag}"])) < 1e-2 * max(1.0, abs(float(g[f"sum_dll_{tag}"])))


def test_more_smooth_oracle_matches_reference_golden():
    """autoregressive_infer_cfg(more_smooth=True) (var.py:178-180): the Gumbel soft embeddings drive f_hat."""
    g = golden("more_smooth_d2.npz")
    vae, var = seeded_models()
    sd, cfg = sd_cpu(var), var_cfg_of(var)
    qs, gs = replay_noise_more_smooth(77, B=2)
    out = VO.ar_infer(sd, cfg, quant_oracle_of(vae), torch.from_numpy(g["labels"]), qs, cfg_scale=1.5, top_k=900,
                      g_noise=gs, codebook=vae.quantize.embedding.weight.detach())
    err = (out["f_hat"] - torch.from_numpy(g["f_hat"])).abs().max().item()
    assert err < 2e-3 * float(np.abs(g["f_hat"]).max()), err


def test_embed_to_fhat_and_get_logits_oracle_match_reference_golden():
    """quant.py:107-121 on arbitrary per-scale maps and var.py:118-124, against outputs of the imported reference
    (oracle/gen_golden_embed.py; same seeded inputs from helpers.py)."""
    from helpers import embed_inputs, logits_inputs
    g = golden("embed_get_logits_d2.npz")
    vae, var = seeded_models(depth=2)
    qo = quant_oracle_of(vae)
    hs = [h.numpy() for h in embed_inputs()]
    f_hat = np.zeros((2, 32, 16, 16), np.float32)
    for si in range(10):
        qo.get_next_autoregressive_input(si, f_hat, None, h=hs[si])
        if si == 3:
            assert np.abs(f_hat - g["fhat_s3"]).max() < 2e-5
        if si == 8:
            assert np.abs(f_hat[:, ::4] - g["fhat_s8_sub"]).max() < 2e-5
    assert np.abs(f_hat - g["fhat_last"]).max() < 2e-5
    h, labels = logits_inputs(var.C)
    sd = sd_cpu(var)
    logits = VO.get_logits(sd, var_cfg_of(var), h, sd["class_emb.weight"][labels])
    assert (logits[:, :, ::8] - torch.from_numpy(g["logits_sub"])).abs().max().item() < 5e-4
