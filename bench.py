#!/usr/bin/env python
"""bench.py — throughput of the VAR next-scale-prediction hot path on B200 (contract: task statement ④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sample_d30|score_d16|sample_d16] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input:
  sample_d30 (default, BASELINE.json configs[4]): VAR-d30 autoregressive_infer_cfg, 256 px, B=256 per GPU, cfg=1.5,
              top_k=900, bf16 GEMMs / fp32 residual, preallocated KV cache. Metric: images/sec.
  score_d16  (configs[3]): VAR-d16 1000-class likelihood scoring of one image per step per GPU... (classes sharded
              over the ranks with one all-gather when N > 1). Metric: images/sec.
  sample_d16 (configs[2]): VAR-d16 sampling, B=64.
  sample_d36_512: the 512 px variant (2240-token pyramid, d36 with shared adaLN), B=32; GPU arm only.
The JSON line carries `value` (device-timed, inputs resident, to f_hat), `e2e` (public API with host labels in, images
out, CNN decoder under bf16 autocast), `roofline` (dominant kernel = the tcgen05 GEMM, timed live at the step's
largest shape), `cpu_baseline` (the oracle port on this box's host cores, bounded sample) and `secondary` (the other
headline workload). `--impl reference` times the CPU oracle port only (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PATCH_NUMS = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
PATCH_NUMS_512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)  # utils/arg_util.py:246-247
L_SEQ = sum(p * p for p in PATCH_NUMS)
V, CVAE = 4096, 32
PROFILE_STEP = os.environ.get("VAR_B200_PROFILE_STEP") == "1"


def flops_per_seq(depth: int, patch_nums=PATCH_NUMS) -> float:
    """Algorithmic FLOPs of one token-pyramid sequence; attention counted on visible pairs only."""
    C = 64 * depth
    vis, cum = 0, 0
    for p in patch_nums:
        cum += p * p
        vis += p * p * cum
    L = cum
    return (24 * C * C * depth * L + 4 * C * depth * vis + 2 * C * V * L + (12 * C * C * depth + 4 * C * C)
            + 2 * CVAE * C * (L - 1))


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"],
                    sustained_mhz=(d.get("clocks_under_load") or {}).get("sm_mhz_median"), src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, sustained_mhz=None, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, reasons=reasons)


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sampling_step(depth: int, n_img: int, threads: int):
    """One bounded CPU sample of the same workload through the oracle port: n_img images, all 10 scales."""
    from oracle import var_oracle as VO
    from oracle.quant_oracle import QuantOracle
    torch.set_num_threads(threads)
    cfg = VO.VarCfg(depth=depth)
    sd, quant = _cpu_state(depth)
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, 1000, (n_img,), generator=g)
    noise = [torch.empty(n_img * p * p, V).exponential_(1, generator=g) for p in PATCH_NUMS]
    t0 = time.perf_counter()
    VO.ar_infer(sd, cfg, quant, labels, noise, cfg_scale=1.5, top_k=900)
    return time.perf_counter() - t0


def cpu_scoring_step(depth: int, n_cls: int, threads: int):
    from oracle import var_oracle as VO
    torch.set_num_threads(threads)
    cfg = VO.VarCfg(depth=depth)
    sd, quant = _cpu_state(depth)
    g = torch.Generator().manual_seed(0)
    idx = [torch.randint(0, V, (1, p * p), generator=g).numpy() for p in PATCH_NUMS]
    t0 = time.perf_counter()
    vin = torch.from_numpy(quant.idxBl_to_var_input(idx))
    logits = VO.var_forward(sd, cfg, torch.arange(n_cls), vin.expand(n_cls, -1, -1))
    VO.class_scores(logits, torch.from_numpy(__import__("numpy").concatenate(idx, axis=1)))
    return time.perf_counter() - t0


_CPU_STATE = {}


def _cpu_state(depth: int):
    if depth not in _CPU_STATE:
        from oracle.quant_oracle import QuantOracle
        from var_b200 import build_vae_var
        from var_b200.init_utils import dense_init_
        import numpy as np
        # VQVAE quantizer weights only (the CNN is outside the hot path): tiny ch to keep construction cheap
        vae, var = build_vae_var("cpu", depth=depth, ch=32)
        dense_init_(var, seed=2)
        dense_init_(vae.quantize, seed=1)
        sd = {k: v.detach() for k, v in var.state_dict().items()}
        phis = vae.quantize.quant_resi.phis()
        quant = QuantOracle(vae.quantize.embedding.weight.detach().numpy(), np.stack([p.weight.detach().numpy() for p in phis]),
                            np.stack([p.bias.detach().numpy() for p in phis]), PATCH_NUMS)
        _CPU_STATE[depth] = (sd, quant)
    return _CPU_STATE[depth]


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


REF_IMAGES = 4    # images per step of the reference arm's sampling sample (CFG batch 8)
REF_CLASSES = 8   # classes per step of the reference arm's scoring sample


_REF_STATE = {}


def _reference_models(depth: int):
    """The UNMODIFIED reference (oracle/_ref = byte copies of /root/reference/models + dist.py, oracle/ref_loader.py)
    on CPU with the same seeded dense init as the GPU arm."""
    if depth not in _REF_STATE:
        from oracle import ref_loader
        from var_b200.init_utils import dense_init_
        import contextlib
        build, _, _ = ref_loader.load()
        with contextlib.redirect_stdout(sys.stderr):  # the reference prints its constructor banner: keep stdout = one JSON line
            vae, var = build(depth=depth)
        dense_init_(vae, seed=1)
        dense_init_(var, seed=2)
        _REF_STATE.clear()  # one model at a time (d30 is 8 GB in fp32)
        _REF_STATE[depth] = (vae, var)
    return _REF_STATE[depth]


def ref_sampling_step(depth: int, n_img: int, threads: int):
    """models/var.py:126-190 as shipped: labels -> images [n_img,3,256,256] (CNN decoder included), fp32, CPU."""
    torch.set_num_threads(threads)
    vae, var = _reference_models(depth)
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, 1000, (n_img,), generator=g)
    t0 = time.perf_counter()
    with torch.inference_mode():
        img = var.autoregressive_infer_cfg(B=n_img, label_B=labels, g_seed=0, cfg=1.5, top_k=900, top_p=0.0)
    assert img.shape == (n_img, 3, 256, 256)
    return time.perf_counter() - t0


def ref_scoring_step(depth: int, n_cls: int, threads: int):
    """The class loop of eval_prob.py:423-465 around the reference's own modules (eval_prob.py itself needs clip /
    matplotlib, absent here): one class per forward, idxBl_to_var_input recomputed per class as the script does."""
    torch.set_num_threads(threads)
    vae, var = _reference_models(depth)
    g = torch.Generator().manual_seed(0)
    gt_idx = [torch.randint(0, V, (1, p * p), generator=g) for p in PATCH_NUMS]
    gt = torch.cat(gt_idx, dim=1)
    t0 = time.perf_counter()
    scores = []
    with torch.inference_mode():
        for c in range(n_cls):
            x_in = vae.quantize.idxBl_to_var_input(gt_idx)                     # eval_prob.py:436
            logits = var(torch.tensor([c]), x_in)                                # eval_prob.py:441
            lp = torch.log_softmax(logits, dim=-1).gather(-1, gt.unsqueeze(-1))  # eval_prob.py:446-463
            scores.append(lp.sum())
    return time.perf_counter() - t0


def cpu_arm(workload):
    """(step function, units per step, kind, sample description) of the CPU arm for this workload."""
    from oracle import ref_loader
    threads = host_threads()
    depth = workload["depth"]
    if ref_loader.available():
        if workload["kind"] == "sample":
            return (lambda: ref_sampling_step(depth, REF_IMAGES, threads), float(REF_IMAGES), "reference",
                    f"{REF_IMAGES} images per step (CFG batch {2 * REF_IMAGES}), all 10 scales, cfg=1.5, top_k=900, fp32: the "
                    "unmodified reference VAR.autoregressive_infer_cfg incl. the CNN decoder, labels -> images")
        return (lambda: ref_scoring_step(depth, REF_CLASSES, threads), REF_CLASSES / 1000.0, "reference",
                f"{REF_CLASSES} of the 1000 classes of one image per step, fp32: the unmodified reference VAR.forward in the "
                "class loop of eval_prob.py:423-465 (one class per forward)")
    if workload["kind"] == "sample":
        return (lambda: cpu_sampling_step(depth, 1, threads), 1.0, "port",
                "1 image per step, all 10 scales, fp32 oracle port of autoregressive_infer_cfg to f_hat (oracle/_ref absent)")
    return (lambda: cpu_scoring_step(depth, 4, threads), 4 / 1000.0, "port",
            "4 of 1000 classes of one image per step, fp32 oracle port of VAR.forward + score (oracle/_ref absent)")


def run_reference_arm(args, workload):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores (the unmodified
    reference from oracle/_ref; the oracle port only if that copy is absent), each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    fn, unit_per_step, kind, sample = cpu_arm(workload)
    for _ in range(args.warmup):
        fn()
    ts = [fn() for _ in range(args.steps)]
    total = sum(ts)
    value = unit_per_step * args.steps / total
    line = dict(impl="reference", metric=workload["metric"], value=value, unit="images/sec", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * total / args.steps, higher_is_better=True,
                scaling=workload["scaling"], vs_baseline=None, dtype="f32", data="synthetic", config=workload["config"],
                sample=dict(what=sample, units_per_step=unit_per_step, note="config is the GPU arm's (the driver pairs the "
                            "two lines on it); THIS arm runs the bounded sample stated here, not the full batch"),
                cpu_baseline=dict(value=value, unit="images/sec", cores=threads, kind=kind, sample=sample),
                e2e=dict(value=value, unit="images/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def time_gemm(M, N, K, epi, reps=20):
    """Live CUDA-event timing of the dominant kernel (tcgen05 GEMM) at the step's largest shape."""
    import ctypes as C
    from var_b200 import lib as L
    lib = L.load()
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.05).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    a = L.GemmArgs()
    a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, epi
    a.bias, a.out = bias.data_ptr(), out.data_ptr()
    st = torch.cuda.current_stream()
    for _ in range(3):
        L.check(lib.var_b200_gemm_bf16(C.byref(a), st.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        L.check(lib.var_b200_gemm_bf16(C.byref(a), st.cuda_stream))
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="var_b200")
    ap.add_argument("--workload", default="sample_d30", choices=["sample_d30", "sample_d16", "score_d16", "sample_d36_512"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (sampling) / classes (scoring)")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    wl = dict(
        sample_d30=dict(kind="sample", depth=30, batch=256,
                        metric="images/sec: VAR-d30 256px CFG sampling (autoregressive_infer_cfg, cfg=1.5, top_k=900)"),
        sample_d16=dict(kind="sample", depth=16, batch=64,
                        metric="images/sec: VAR-d16 256px CFG sampling (autoregressive_infer_cfg, cfg=1.5, top_k=900)"),
        score_d16=dict(kind="score", depth=16, batch=1000,
                       metric="images/sec: VAR-d16 1000-class likelihood scoring (eval_prob bayesian)"),
        # the 512 px variant: 2240-token pyramid, d36 with shared adaLN (GPU arm only)
        sample_d36_512=dict(kind="sample", depth=36, batch=32, px=512, patch_nums=PATCH_NUMS_512, shared_aln=True,
                            metric="images/sec: VAR-d36 512px CFG sampling (autoregressive_infer_cfg, cfg=1.5, top_k=900)"),
    )[args.workload]
    pns = wl.get("patch_nums", PATCH_NUMS)
    px = wl.get("px", 256)
    L_wl = sum(p * p for p in pns)
    if args.batch:
        wl["batch"] = args.batch
    # sampling: per-GPU batch fixed (weak); scoring: the 1000 classes of one image are split over the ranks (strong)
    wl["scaling"] = "weak" if wl["kind"] == "sample" else "strong"
    wl["config"] = dict(workload=args.workload, depth=wl["depth"], px=px, tokens=L_wl,
                        per_gpu_batch=wl["batch"] if wl["kind"] == "sample" else 1,
                        classes=1000 if wl["kind"] == "score" else None,
                        sampler="cfg=1.5,top_k=900,top_p=0" if wl["kind"] == "sample" else None,
                        l2="activations per step >> 126 MB L2 (inputs larger than L2)", parallelism=f"dp{args.gpus}")

    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""  # CPU arm: the reference picks its device from torch.cuda.is_available()
        if px != 256:
            print(json.dumps(dict(impl="reference", unavailable="the CPU arm is built for the 256 px headline workloads")))
            return
        run_reference_arm(args, wl)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from var_b200 import build_vae_var, lib as L
    from var_b200.init_utils import dense_init_
    lib = L.load()
    torch.backends.cudnn.benchmark = True  # CNN encoder/decoder (boundary helpers, cuDNN): let cuDNN pick its algorithms
    pk = peaks()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(obj):
        """obj of every rank, in rank order (diagnostics only, outside every timed region)."""
        if world == 1:
            return [obj]
        out = [None] * world
        torch.distributed.all_gather_object(out, obj)
        return out

    def build(depth, **kw):
        vae, var = build_vae_var(dev, depth=depth, **kw)
        dense_init_(var, seed=2)
        dense_init_(vae, seed=1)
        var.eval(); vae.eval(); var.cond_drop_rate = 0
        return vae, var

    def sampling_runner(vae, var, B, cuda_graph=False):
        g = torch.Generator(device="cpu").manual_seed(rank)
        labels_host = torch.randint(0, 1000, (B,), generator=g).pin_memory()
        labels_dev = labels_host.to(dev)

        def step_hot():  # inputs resident, hot path to f_hat
            var.autoregressive_infer_cfg(B, labels_dev, g_seed=0, cfg=1.5, top_k=900, top_p=0.0, decode=False,
                                         cuda_graph=cuda_graph)

        img_host = torch.empty((B, 3, px, px), dtype=torch.float32).pin_memory()

        vae.decoder_dtype = torch.bfloat16  # CNN decoder as the 16-bit NHWC plan (own tcgen05 convolutions + GroupNorm glue)

        def step_e2e():  # public API: host labels in, images out
            lab = labels_host.to(dev, non_blocking=True)
            img = var.autoregressive_infer_cfg(B, lab, g_seed=0, cfg=1.5, top_k=900, top_p=0.0, cuda_graph=cuda_graph)
            img_host.copy_(img, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return img_host
        return step_hot, step_e2e, B, labels_host.numel() * 8, B * 3 * px * px * 4

    def scoring_runner(vae, var, K):
        from var_b200.scoring import class_log_likelihoods, gather_class_scores, shard_range
        g = torch.Generator(device="cpu").manual_seed(0)
        img_host = (torch.rand(1, 3, 256, 256, generator=g) * 2 - 1).pin_memory()
        lo, hi = shard_range(K, rank, world)
        labels = torch.arange(lo, hi, device=dev)
        gt_idx = vae.img_to_idxBl(img_host.to(dev))

        def step_hot():  # tokens resident; class-sharded forward + fused score + all-gather
            local_scores = class_log_likelihoods(var, gt_idx, labels, class_batch=125)
            return gather_class_scores(local_scores, K, rank, world)

        def step_e2e():  # image on the host -> encoder -> tokens -> scores -> prediction on the host
            img = img_host.to(dev, non_blocking=True)
            idx = vae.img_to_idxBl(img)
            s = gather_class_scores(class_log_likelihoods(var, idx, labels, class_batch=125), K, rank, world)
            return int(torch.argmax(s).item())
        return step_hot, step_e2e, 1.0 / world, img_host.numel() * 4, 8

    def measure(step, steps, warmup):
        """-> (max over ranks of the device time of `steps` steps [s], launches of this rank, this rank's own time)."""
        for _ in range(warmup):
            step()
        barrier()
        st = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.var_b200_launch_count()
        e0.record(st)
        for i in range(steps):
            # VAR_B200_PROFILE_STEP=1: cudaProfilerStart/Stop around the last timed step, for
            # `ncu --profile-from-start off` launch lists (a number printed by a run under ncu is never a bench value)
            prof = PROFILE_STEP and i == steps - 1
            if prof:
                if wl["kind"] == "sample":
                    from var_b200 import var as _var_mod
                    _var_mod._PROFILE_ARMED[0] = True  # the AR loop starts the profiler at VAR_B200_PROFILE_FROM_SCALE
                else:
                    torch.cuda.profiler.start()
            step()
            if prof:
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
        e1.record(st)
        barrier()
        launches = lib.var_b200_launch_count() - n0
        own = e0.elapsed_time(e1) * 1e-3
        return max_over_ranks(own), launches, own

    def kernel_split(step):
        """One extra step with the library's per-kernel CUDA events (recorded on the launching stream around every launch)."""
        barrier()
        with L.kernel_profile() as kp:
            step()
        return kp

    def family_ms(kp):
        fam = dict(gemm=sum(v for k, v in kp.ms.items() if k.startswith("gemm")), attn=kp.ms.get("attn", 0.0),
                   ln_modulate=kp.ms.get("ln_modulate", 0.0))
        fam["other"] = sum(kp.ms.values()) - sum(fam.values())
        return {k: round(v, 2) for k, v in fam.items()}

    shard_lo, shard_hi = 0, 0
    vae, var = build(wl["depth"], **({"patch_nums": pns, "shared_aln": True} if wl.get("shared_aln") else {}))
    if wl["kind"] == "sample":
        hot, e2e, units, h2d, d2h = sampling_runner(vae, var, wl["batch"])
    else:
        from var_b200.scoring import shard_range
        shard_lo, shard_hi = shard_range(wl["batch"], rank, world)
        hot, e2e, units, h2d, d2h = scoring_runner(vae, var, wl["batch"])

    clocks = ClockSampler(local)
    clocks.start()
    t_hot, launches, own_hot = measure(hot, args.steps, max(args.warmup, 3))
    clk = clocks.stop()
    t_e2e, _, _ = measure(e2e, args.steps, 1)
    value = units * world * args.steps / t_hot
    value_e2e = units * world * args.steps / t_e2e

    # ---- roofline of the dominant kernel family (tcgen05 GEMM = 94-97 % of the FLOPs) ----
    # (a) inside the step: one extra hot step with the library's per-kernel CUDA events (recorded on the launching
    #     stream around every launch); achieved = algorithmic GEMM FLOPs of the step / summed GEMM kernel time,
    #     against the SUSTAINED measured peak (kernel timed inside a long step);
    # (b) alone: fc1 at the step's largest shape, 20 launches, against the BURST measured peak.
    depth = wl["depth"]
    C_ = 64 * depth
    ncu_path = ROOT / "profiles" / "r02_ncu_gemm_lnf.json"  # ncu --set full of the d30 GEMM flavours (tools/gemm_prof_lnf.py)
    ncu_traffic = json.loads(ncu_path.read_text()).get("fc2", {}) if ncu_path.exists() else {}
    n_seq_step = (2 * wl["batch"]) if wl["kind"] == "sample" else (shard_hi - shard_lo)
    gemm_flops = n_seq_step * (24.0 * C_ * C_ * depth * L_wl + 2.0 * C_ * V * L_wl + 12.0 * C_ * C_ * depth + 4.0 * C_ * C_)
    kp = kernel_split(hot)
    gemm_ms = sum(v for k, v in kp.ms.items() if k.startswith("gemm"))
    all_ms = sum(kp.ms.values())
    gemm_tf_step = gemm_flops / (gemm_ms * 1e-3) / 1e12
    M_big = (2 * wl["batch"] * pns[-1] ** 2) if wl["kind"] == "sample" else min(125, shard_hi - shard_lo) * L_SEQ
    t_g = time_gemm(M_big, 4 * C_, C_, L.EPI_GELU_BF16)
    gemm_tf = 2.0 * M_big * 4 * C_ * C_ / t_g / 1e12
    fl_img = (2 if wl["kind"] == "sample" else 1000) * flops_per_seq(depth, pns)
    step_tf = value / world * fl_img / 1e12
    roofline = dict(bound="tensor", kernel="gemm_bf16_kernel<BN,EPI,2> (all fused epilogues of the step)",
                    achieved=gemm_tf_step, peak=pk["sustained"], unit="TFLOP/s", frac=gemm_tf_step / pk["sustained"],
                    traffic=ncu_traffic.get("traffic_GB") if wl["kind"] == "sample" and depth == 30 else None,
                    traffic_note=("GB per launch of the GEMM with the largest share of the step (" + ncu_traffic.get("kernel", "") +
                                  "), ncu --set full dram read+write, vs %.3f GB algorithmic; tensor pipe %.1f %% active under ncu. "
                                  "Writes equal the algorithmic bytes; the activation operand is re-read 1.6-3.8x depending on the "
                                  "physical GPU's SM-to-die map (profiles/r02_gemm_dram_probe.txt, r02_ncu_gemm_lnf.json)"
                                  % (ncu_traffic.get("algorithmic_GB", 0.0), ncu_traffic.get("tensor_active_pct", 0.0)))
                    if ncu_traffic else None,
                    peak_source=f"{pk['src']} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
                    # the sustained peak was measured at pk["sustained_mhz"]; the step ran at clk["sm_mhz"] under the same
                    # power cap: the same fraction against the peak scaled to the step's clock
                    frac_at_step_clock=(gemm_tf_step / (pk["sustained"] * clk["sm_mhz"] / pk["sustained_mhz"])
                                        if clk.get("sm_mhz") and pk.get("sustained_mhz") else None),
                    gemm_share_of_kernel_time=gemm_ms / all_ms,
                    kernel_ms={k: round(v, 3) for k, v in sorted(kp.ms.items(), key=lambda kv: -kv[1])},
                    kernel_launches=kp.n,
                    isolated=dict(kernel=f"gemm_bf16_kernel<BN,GELU,2> M={M_big} N={4 * C_} K={C_}", achieved=gemm_tf,
                                  peak=pk["burst"], frac=gemm_tf / pk["burst"]),
                    step_achieved=step_tf, step_peak=pk["sustained"], step_frac=step_tf / pk["sustained"],
                    step_note="whole-step algorithmic FLOPs (SURVEY 8d; attention on visible pairs only) / step time")
    # every rank's own step time, kernel-family split and clocks: attributes a scaling loss to a GPU or to the host
    per_rank = all_ranks(dict(rank=rank, ms_per_step=round(1e3 * own_hot / args.steps, 2), kernel_ms=family_ms(kp),
                              host_gap_ms=round(1e3 * own_hot / args.steps - all_ms, 2), sm_mhz=clk["sm_mhz"],
                              reasons=clk["reasons"]))

    line = dict(metric=wl["metric"], value=value, unit="images/sec", n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=1e3 * t_hot / args.steps, higher_is_better=True, scaling=wl["scaling"],
                vs_baseline=None, dtype="bf16", data="synthetic (random labels/images, seeded dense random-init weights)",
                config=wl["config"], clocks=clk,
                e2e=dict(value=value_e2e, unit="images/sec", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         note="public API call with host buffers: pinned H2D of the inputs, the whole call (sampling: incl. the bf16 NHWC CNN decoder on the tcgen05 implicit-GEMM convolutions), D2H of the result"),
                gpu_launches=int(launches), roofline=roofline, per_rank=per_rank)

    sec_steps = max(10, args.steps)
    if not args.no_secondary and args.workload == "sample_d30":
        # (1) BASELINE configs[4] as written: 256 images TOTAL, 256/N per GPU ("strong"), the whole 10-scale loop replayed
        #     as one CUDA graph (at 32 images per GPU the small scales are launch-bound in eager mode)
        B_total = wl["batch"]
        if B_total % world == 0:
            Bs = B_total // world
            hot_s, e2e_s, _, _, _ = sampling_runner(vae, var, Bs, cuda_graph=True)
            ts, _, own_s = measure(hot_s, args.steps, 2)
            tse, _, _ = measure(e2e_s, args.steps, 1)
            kps = kernel_split(lambda: var.autoregressive_infer_cfg(Bs, torch.zeros(Bs, dtype=torch.long, device=dev), g_seed=0,
                                                                    cfg=1.5, top_k=900, decode=False))
            fams = all_ranks(dict(rank=rank, ms_per_step=round(1e3 * own_s / args.steps, 2), eager_kernel_ms=family_ms(kps)))
            v_s = B_total * args.steps / ts
            line["strong_config5"] = dict(
                metric="images/sec: VAR-d30 256px CFG sampling, 256 images total batch-sharded over the GPUs (BASELINE configs[4])",
                scaling="strong", total_batch=B_total, per_gpu_batch=Bs, cuda_graph=True, value=v_s, unit="images/sec",
                ms_per_step=1e3 * ts / args.steps, e2e=B_total * args.steps / tse,
                step_frac=v_s / world * fl_img / 1e12 / pk["sustained"], per_rank=fams,
                limiting_kernel=max(fams[0]["eager_kernel_ms"].items(), key=lambda kv: kv[1])[0])
        # (2) the other headline metric in the same run: d16 1000-class scoring of one image per step, the classes
        #     sharded over the ranks, one all-gather of the per-class scores inside the timed region
        del var, vae, hot, e2e
        if B_total % world == 0:
            del hot_s, e2e_s
        torch.cuda.empty_cache()
        vae2, var2 = build(16)
        hot2, e2e2, units2, _, _ = scoring_runner(vae2, var2, 1000)
        t2, _, own2 = measure(hot2, sec_steps, 2)
        t2e, _, _ = measure(e2e2, sec_steps, 1)
        kp2 = kernel_split(hot2)
        v2 = units2 * world * sec_steps / t2
        line["secondary"] = dict(metric="images/sec: VAR-d16 1000-class likelihood scoring (classes sharded over the GPUs, "
                                        "one all-gather of the scores)", value=v2, unit="images/sec", steps=sec_steps,
                                 scaling="strong", e2e=units2 * world * sec_steps / t2e, pairs_per_sec=v2 * 1000,
                                 ms_per_step=1e3 * t2 / sec_steps,
                                 step_frac=v2 / world * 1000 * flops_per_seq(16) / 1e12 / pk["sustained"],
                                 per_rank=all_ranks(dict(rank=rank, ms_per_step=round(1e3 * own2 / sec_steps, 2),
                                                         kernel_ms=family_ms(kp2))))
    if rank == 0 and world == 1 and not args.no_cpu and px == 256:
        line["cpu_baseline"] = cpu_baseline_subprocess(args.workload)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_baseline_subprocess(workload: str):
    """The CPU arm (`--impl reference`, one step, no warm-up) in a child process without CUDA: the reference chooses its
    device from torch.cuda.is_available() at import time, and its fp32 d30 copy (8 GB) is gone when the child exits."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
                            "--warmup", "0"], env=env, capture_output=True, text=True, timeout=900)
        ref = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        return ref["cpu_baseline"]
    except Exception as e:  # noqa: BLE001 - a failed baseline must not lose the GPU measurement
        return dict(value=None, unit="images/sec", cores=host_threads(), kind="unavailable", sample=f"CPU arm failed: {e!r}"[:300])


if __name__ == "__main__":
    main()
