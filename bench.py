#!/usr/bin/env python
"""bench.py — FairLoRA training throughput on B200 (BASELINE.json metric; default = configs[1], ViT-B/16, batch 64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`
(one rank per GPU = one simulated site, NCCL).  Rank 0 prints ONE JSON line.

A "step" = one GLP_OT_SVLoRA training iteration of one site on one batch of 64 synthetic images:
pixel normalisation -> CLIP image encoder with the fused FairLoRA linears -> text encoder (4 prompts) -> GLP_OT Sinkhorn
head -> cross-entropy -> backward -> SGD stepped twice (reference quirk F6).  The per-round FedAvg (all-reduce of U, V,
s_g over NVLink + EMA / shared-half-S epilogue) runs once inside the timed region.
  value    : images/s with the batches already resident in HBM (device timed, CUDA events, max over ranks)
  e2e      : the same metric through the trainer's PUBLIC `train(idx=...)` — what federated_main.py:622 calls — over a
             loader of pinned HOST batches: every step's H2D copy, the per-step loss / accuracy / training-AUC bookkeeping
             and their read-back are inside the timed region
  roofline : the dominant kernel (fused SVLoRA tcgen05 GEMM): algorithmic FLOPs per launch / CUDA-event duration per
             launch measured live (ffm_profile_*, events on the launching stream) in a second short profiled region
  roofline_hbm : the persistent Sinkhorn kernel on a stress shape (65 536 problems, read K + write T) against measured HBM
  cpu_baseline / reference_gpu_eager : the reference's own modules (oracle/_ref staged archive, else the oracle port)
             timed on this box's host cores / on this GPU in eager PyTorch fp32 — reported beside, not the target
  agg_parity : (N > 1) one extra round after the timed region: the NCCL-aggregated buffer against the reference formula
             evaluated in fp64 on the gathered client buffers, and whether every rank holds identical bytes
--impl reference times the reference's CPU implementation of the same step with all host threads as the reference arm.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "train_images_per_s"
UNIT = "images/s"
BATCH = 64

# BASELINE.json configs[1..4] (configs[0] is the reference's own CPU-runnable case: a parity-test shape, not a bench line)
CONFIGS = {
    2: dict(tag="configs[1]", desc="CLIP ViT-B/16, synthetic 2D SLO 224x224", dataset="FairFedMed", modality="slo_fundus",
            attributes=["race"], attr_type="race", backbone="ViT-B/16", rank=12, alpha=2.0, channels=3),
    3: dict(tag="configs[2]", desc="CLIP ViT-B/16, synthetic 3D OCT B-scan stacks [32,224,224] = 4 slice-images/sample "
            "(T = 50 432 token rows)", dataset="FairFedMed", modality="oct_bscans",
            attributes=["race", "gender", "ethnicity", "language"], attr_type="race", backbone="ViT-B/16", rank=12,
            alpha=2.0, channels=32),
    4: dict(tag="configs[3]", desc="CLIP ResNet50 backbone (rank-32 FairLoRA on the 1x1 convs, LoRA attention pool, "
            "trainable BatchNorm), synthetic 2D SLO", dataset="FairFedMed", modality="slo_fundus",
            attributes=["race"], attr_type="race", backbone="RN50", rank=32, alpha=8.0, channels=3),
    5: dict(tag="configs[4]", desc="CLIP ViT-B/16, synthetic FedChexMimic chest X-ray 224x224", dataset="FedChexMimic",
            modality="slo_fundus", attributes=["race", "gender", "age"], attr_type="race", backbone="ViT-B/16", rank=12,
            alpha=2.0, channels=3),
}


def make_cfg(world: int, batch: int, ot: str, c: dict):
    from fairfedmed_b200.config import get_cfg_default
    cfg = get_cfg_default()
    cfg.DATASET.merge_from_dict(dict(USERS=max(world, 1), NUM_TRAIN_PER_CLIENT=batch * 2, NUM_TEST_PER_CLIENT=batch,
                                     NAME=c["dataset"], MODALITY_TYPE=c["modality"], ATTRIBUTES=c["attributes"],
                                     ATTRIBUTE_TYPE=c["attr_type"], DIM_PER_3D_SLICE=8, SYNTHETIC=True))
    cfg.DATALOADER.TRAIN_X.BATCH_SIZE = batch
    cfg.MODEL.BACKBONE.NAME = c["backbone"]
    cfg.TRAINER.GLP_OT.OT = ot
    cfg.TRAINER.GLP_OT_LORA.merge_from_dict(dict(RANK=c["rank"], ALPHA=c["alpha"]))
    return cfg


def workload_config(world, batch, ot, c):
    from fairfedmed_b200.config import ATTRIBUTE_GROUPS
    G = len(ATTRIBUTE_GROUPS[c["dataset"]][c["attr_type"]])
    return {
        "workload": f"{c['tag']}: FairLoRA (GLP_OT_SVLoRA) {c['desc']}, one site per GPU, batch {batch}/GPU, "
                    f"{c['attr_type']} attribute ({G} groups), rank {c['rank']} alpha {c['alpha']:g}, OT={ot}, "
                    "FedAvg of U,V,s_g per round",
        "global_batch": batch * world, "batch_per_gpu": batch, "sites": world, "ot": ot,
        "parallelism": f"sites{world}", "optimizer": "SGD lr1e-3 m0.9 wd5e-4, stepped twice per iteration (reference F6)",
        "l2_policy": "no explicit flush: one step touches >2 GB of activations/weights (>> 126 MB L2)",
    }


def synthetic_batch(gen, batch, c, groups):
    """FairFedMed-shaped host batch: uint8-valued pixels as float32 0..255, balanced labels, uniform groups."""
    if c["channels"] == 3:
        img = torch.randint(0, 256, (batch, 1, 224, 224), generator=gen).float().repeat(1, 3, 1, 1)
    else:
        img = torch.randint(0, 256, (batch, c["channels"], 224, 224), generator=gen).float()
    lab = (torch.arange(batch) % 2).long()
    att = torch.stack([torch.randint(0, g, (batch,), generator=gen) for g in groups], dim=1)
    return {"img": img, "label": lab, "attrs": att}


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
def reference_step_fn(batch: int, ot: str, c: dict, device: str = "cpu", seed: int = 1):
    """step() -> loss of ONE training iteration through the reference's own modules (real reference when the tree or the
    staged archive oracle/_ref/reference.zip is there — CustomCLIP + apply_lora_to_model + torch SGD stepped twice, as
    trainers/GLP_OT_SVLoRA.py:804-975 does — else the restated oracle port).  Returns (step, kind)."""
    import torch.nn.functional as F
    from fairfedmed_b200.config import ATTRIBUTE_GROUPS
    from oracle import shim
    G = len(ATTRIBUTE_GROUPS[c["dataset"]][c["attr_type"]])
    g = torch.Generator().manual_seed(seed)
    b = synthetic_batch(g, batch, c, [G])
    image, label, attr = b["img"].to(device), b["label"].to(device), b["attrs"][:, 0].contiguous()
    is_oct = c["modality"] == "oct_bscans"
    if shim.available():
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):     # the reference prints while it builds: keep stdout to ONE JSON line
            return _reference_modules_step(batch, ot, c, device, seed, G, g, image, label, attr), "reference"
    return _port_step(batch, ot, c, device, seed, G, g, image, label, attr, is_oct), "port"


def _reference_modules_step(batch, ot, c, device, seed, G, g, image, label, attr):
    import torch.nn.functional as F
    from oracle import shim
    if True:
        T, CM, _, _ = shim.modules()
        cfg = shim.make_cfg(modality=c["modality"], ot=ot, dataset=c["dataset"], dim_per_3d_slice=8)
        torch.manual_seed(seed)
        dd = {"trainer": "GLP_OT", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0}
        if c["backbone"] == "RN50":
            clip_model = CM.CLIP(1024, 224, (3, 4, 6, 3), 64, None, 77, 49408, 512, 8, 12, dd).float()
        else:
            clip_model = CM.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12, dd).float()
        classes = ["No Finding", "Finding"] if c["dataset"] == "FedChexMimic" else ["NOT Glaucoma", "Glaucoma"]
        model = T.CustomCLIP(cfg, classes, clip_model)
        for n_, p_ in model.named_parameters():
            bn = ".bn" in n_ or "downsample.1" in n_
            p_.requires_grad_("prompt_learner" in n_ or "proj_per_3d_slice" in n_ or (c["backbone"] == "RN50" and bn))
        T.apply_lora_to_model(model, True, rank=c["rank"], alpha=c["alpha"], lora_type="FairLoRA", global_s=False,
                              num_attrs=G)
        with torch.no_grad():
            for n_, p_ in model.named_parameters():
                if "lora_A" in n_:
                    p_.copy_(0.02 * torch.randn(p_.shape, generator=g))
        model.to(device).train()
        params = [p_ for p_ in model.parameters() if p_.requires_grad]
        optim = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)

        def step():
            logits = model(image, attr)                    # attr stays a CPU tensor, as upstream (:988-994)
            loss = F.cross_entropy(logits, label)
            optim.zero_grad()
            loss.backward()
            optim.step(); optim.step()                     # one optimizer registered under two names (F6)
            return float(loss.detach())

        return step


def _port_step(batch, ot, c, device, seed, G, g, image, label, attr, is_oct):
    """restated port (no reference tree, no staged archive)"""
    import torch.nn.functional as F
    from fairfedmed_b200.clip_model import CustomCLIP
    from fairfedmed_b200.modules import apply_lora_to_model
    from oracle import ref_port as rp
    if c["backbone"] == "RN50" or is_oct:
        raise RuntimeError("the port arm covers the 2-D ViT configs; stage the reference (python oracle/stage_ref.py)")
    torch.manual_seed(seed)
    model = CustomCLIP(ot=ot)
    for n, p in model.named_parameters():
        p.requires_grad_("prompt_learner" in n)
    apply_lora_to_model(model, True, rank=c["rank"], alpha=c["alpha"], lora_type="FairLoRA", num_attrs=G)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    params = {k: v.detach().clone().float().to(device) for k, v in model.state_dict().items()}
    for k in names:
        if "lora_A" in k:
            params[k] = (0.02 * torch.randn(params[k].shape, generator=g)).to(device)
        params[k].requires_grad_(True)
    eot = model.prompt_learner.eot_index.clone().to(device)
    del model
    bufs = [None] * len(names)
    attr_d = attr.to(device)

    def step():
        logits = rp.custom_clip_forward(image, attr_d, params, eot, ot=ot, scaling=c["alpha"] / c["rank"])
        loss = F.cross_entropy(logits, label)
        grads = torch.autograd.grad(loss, [params[k] for k in names])
        rp.sgd_double_step([params[k] for k in names], grads, bufs, lr=1e-3)
        return float(loss.detach())

    return step


def time_reference(steps: int, warmup: int, batch: int, ot: str, c: dict, budget_s: float, device: str = "cpu"):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, kind = reference_step_fn(batch, ot, c, device)
    sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)
    t0 = time.perf_counter()
    step(); sync()                                    # first call also pays one-off allocation
    probe = time.perf_counter() - t0
    done_w = 1
    while done_w < warmup and (time.perf_counter() - t0) < budget_s * 0.3:
        step(); done_w += 1
    sync()
    if done_w > 1:
        t = time.perf_counter(); step(); sync(); probe = time.perf_counter() - t; done_w += 1
    n = max(1, min(steps, int((budget_s * 0.7) / max(probe, 1e-4))))
    t1 = time.perf_counter()
    for _ in range(n):
        step()
    sync()
    dt = time.perf_counter() - t1
    return {"images_per_s": batch * n / dt, "ms_per_step": 1e3 * dt / n, "steps_run": n, "warmup_run": done_w,
            "threads": threads, "batch": batch, "kind": kind}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    c = CONFIGS[args.config]
    res = time_reference(args.steps, args.warmup, batch=args.cpu_batch, ot=args.ot, c=c, budget_s=args.cpu_budget)
    what = ("the UNMODIFIED reference modules (trainers/GLP_OT_SVLoRA.py CustomCLIP + apply_lora_to_model, clip/model.py; "
            "staged by oracle/stage_ref.py)" if res["kind"] == "reference" else "oracle port (oracle/ref_port.py)")
    sample = (f"{what}, torch-CPU fp32, the same training step: batch {res['batch']} images/step, {res['steps_run']} "
              f"timed steps after {res['warmup_run']} warm-up (requested {args.steps}/{args.warmup}, capped to a "
              f"{args.cpu_budget:.0f}s budget), {res['threads']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": res["images_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": res["steps_run"], "warmup": res["warmup_run"], "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(max(args.gpus, 1), res["batch"], args.ot, c),
        "cpu_baseline": {"value": res["images_per_s"], "unit": UNIT, "cores": res["threads"], "kind": res["kind"],
                         "sample": sample},
        "e2e": {"value": res["images_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------- our arm
def sinkhorn_stress(dev, peaks):
    """Persistent Sinkhorn on 65 536 problems x 196 x 2 (205 MB read K + write T), L2 flushed between runs."""
    from fairfedmed_b200 import ops
    P, M, N = 65536, 196, 2
    g = torch.Generator(device=dev).manual_seed(3)
    sim = torch.rand((P, M, N), device=dev, generator=g) * 0.4 + 0.3
    K = torch.exp(-(1.0 - sim) / 0.1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times, iters = [], 0
    for i in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, status = ops.sinkhorn(K, mode="Sinkhorn", thresh=1e-3, max_iter=100)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
        iters = int(status[0].item())
    ms = statistics.median(times)
    byts = 2.0 * P * M * N * 4
    peak = peaks.get("hbm_gbs") or 6650.0
    ach = byts / (ms * 1e-3) / 1e9
    return {"kernel": "ffm::sinkhorn_kernel (all iterations in one persistent launch)", "bound": "hbm",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6.65 TB/s",
            "shape": f"{P} problems x {M} x {N} fp32 (read K + write T = {byts / 1e6:.0f} MB)", "iterations": iters,
            "us": 1e3 * ms, "algorithmic_bytes": "2 * P * M * N * 4 (SURVEY 8d stand-alone Sinkhorn)"}


def aggregation_parity(tr, agg, dist, world, rank, dev):
    """One extra round: NCCL result vs the reference formula (utils/fed_utils.py:42-100) in fp64 on the gathered buffers."""
    spec = tr.flat_spec
    G, r = spec.G, spec.r
    n_k = 1000 + 37 * rank
    n_kg = [300 + 11 * rank, 200 + 5 * rank, n_k - 500 - 16 * rank][:G] if G >= 3 else [n_k // 2, n_k - n_k // 2][:G]
    local = tr.get_flat().clone()
    prev = local.clone()
    dist.broadcast(prev, src=0)
    epoch, max_epoch, beta = 3, 50, 0.999
    new = agg.aggregate(local, prev, n_k, n_kg, True, epoch, max_epoch, beta=beta, shared_half_s=True)
    flats = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(flats, local)
    cnt = torch.tensor([n_k] + n_kg, dtype=torch.float64, device=dev)
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    news = [torch.empty_like(new) for _ in range(world)]
    dist.all_gather(news, new)
    identical = all(bool(torch.equal(news[0], t)) for t in news)
    if rank != 0:
        return None
    C = torch.stack(cnts).cpu()                                  # [world, 1 + G]
    w_s = C[:, 0] / C[:, 0].sum()
    w_g = C[:, 1:] / C[:, 1:].sum(0, keepdim=True)
    F64 = torch.stack([f.cpu().double() for f in flats])          # [world, P]
    avg = torch.zeros(F64.shape[1], dtype=torch.float64)
    for key, off, shp, kind in zip(spec.keys, spec.offsets, spec.shapes, spec.kinds):
        n = int(torch.Size(shp).numel())
        seg = F64[:, off:off + n]
        if kind == 1:
            s = (seg.view(world, G, r) * w_g[:, :, None]).sum(0)
            s[:, : r // 2] = s[:, : r // 2].mean(0, keepdim=True)
            avg[off:off + n] = s.reshape(-1)
        else:
            avg[off:off + n] = (seg * w_s[:, None]).sum(0)
    bd = beta * epoch / max_epoch
    ref = (1 - bd) * avg + bd * prev.cpu().double()
    err = float((news[0].cpu().double() - ref).abs().max() / ref.abs().max())
    return {"max_rel_err": err, "identical_across_ranks": identical, "ranks": world, "elements": int(ref.numel()),
            "reference": "utils/fed_utils.py:42-100 formula in fp64 on the all-gathered client buffers"}


def run_ours(args):
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the fairfedmed_b200 path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from fairfedmed_b200 import _cabi
    import fairfedmed_b200.trainer  # noqa: F401
    from fairfedmed_b200.config import ATTRIBUTE_GROUPS
    from fairfedmed_b200.data import CachedLoader
    from fairfedmed_b200.fed_utils import FederatedAggregator
    from fairfedmed_b200.registry import build_trainer

    lib = _cabi.load()
    c = CONFIGS[args.config]
    cfg = make_cfg(world, BATCH, args.ot, c)
    cfg.SEED = 1
    tr = build_trainer(cfg)
    tr.sync_metrics = False
    tr.step_auc = False
    tr.model.check_nan = False           # NaN check of the plan is a host sync; status is read back once per region
    tr.batch_idx, tr.num_batches = 0, 10 ** 9
    with torch.no_grad():
        g = torch.Generator().manual_seed(7)
        for n_, p_ in tr.model.named_parameters():
            if "lora_A" in n_:           # non-zero A so every gradient path does real work
                p_.copy_((0.02 * torch.randn(p_.shape, generator=g)).to(dev))
            if ".bn3.weight" in n_:      # CLIP zero-initialises the last BatchNorm of every bottleneck: open the branch
                p_.fill_(1.0)
    agg = FederatedAggregator(tr.flat_spec)
    n_k = BATCH * 20
    G = tr.num_groups
    n_kg = [n_k // G] * (G - 1) + [n_k - (G - 1) * (n_k // G)]

    # synthetic pool: pinned host batches (e2e) and device-resident copies (value)
    pool = 4
    gen = torch.Generator().manual_seed(100 + rank)
    groups = [len(ATTRIBUTE_GROUPS[c["dataset"]][a]) for a in c["attributes"]]
    host, devb = [], []
    for _ in range(pool):
        host.append({k: v.pin_memory() for k, v in synthetic_batch(gen, BATCH, c, groups).items()})
        devb.append({k: v.to(dev) for k, v in host[-1].items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fed_round(epoch):
        prev = tr.get_flat().clone()
        new = agg.aggregate(tr.get_flat(), prev, n_k, n_kg, True, epoch, 50, shared_half_s=True)
        tr.set_flat(new)

    def timed(body, warm):
        warm()
        fed_round(0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.ffm_launch_count(1)
        e0.record()
        body()
        fed_round(1)
        e1.record()
        barrier()
        launches = lib.ffm_launch_count(0)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), int(launches)

    # one eager step to count this library's launches per step, then capture the whole step into a CUDA graph
    for i in range(2):
        tr.forward_backward(devb[i % pool])
    torch.cuda.synchronize()
    lib.ffm_launch_count(1)
    tr.forward_backward(devb[0])
    torch.cuda.synchronize()
    launches_per_step = int(lib.ffm_launch_count(0))
    use_graph = not args.no_graph
    graph_note = "eager launches"
    if use_graph:
        try:
            tr.capture_step_graph(devb[0])
            graph_note = "whole step (fwd+bwd+2xSGD) replayed from one CUDA graph"
        except Exception as e:                       # capture is an optimisation, not a correctness requirement
            use_graph = False
            graph_note = f"eager launches (graph capture failed: {type(e).__name__}: {e})"[:300]
            torch.cuda.synchronize()
    tr.use_cuda_graph = use_graph
    run_step = tr.forward_backward_graphed if use_graph else tr.forward_backward

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(lambda: [run_step(devb[i % pool]) for i in range(args.steps)],
                               lambda: [run_step(devb[i % pool]) for i in range(args.warmup)])
    if use_graph:       # replayed kernels do not pass through the library's launch counter: add them back
        launches += launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = BATCH * world * args.steps / (ms_total / 1e3)

    # ---- e2e: the trainer's public train(idx=0) over a loader of pinned HOST batches ----
    tr.step_auc = True                                 # per-step training AUC like the reference (:964-970)
    tr.step_metrics = "epoch" if use_graph else "step"
    tr.sync_metrics = True
    ds0 = tr.fed_train_loader_x_dict[0].dataset
    import gc
    gc.collect()
    gc.disable()                                       # a collector pause inside a 20-step region is a 1 ms/step artefact
    last = {}
    try:
        def e2e_epoch(n):
            tr.fed_train_loader_x_dict[0] = CachedLoader(host, n, ds0)
            last["summary"] = tr.train(idx=0, global_epoch=0, is_fed=True, is_last_client=True)
        ms_e2e, _ = timed(lambda: e2e_epoch(args.steps), lambda: e2e_epoch(max(args.warmup, 3)))
    finally:
        gc.enable()
    e2e_value = BATCH * world * args.steps / (ms_e2e / 1e3)
    final = last.get("summary") or {}
    tr.step_auc = False
    tr.sync_metrics = False

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream ----
    lib.ffm_profile_enable(1)
    prof_steps = 3
    for i in range(prof_steps):
        tr.forward_backward(devb[i % pool])
    torch.cuda.synchronize()
    lib.ffm_profile_enable(0)
    cap = 256 * prof_steps + 8
    ms_buf = (ctypes.c_float * cap)()
    tkn_buf = (ctypes.c_int * (3 * cap))()
    nrec = lib.ffm_profile_read(ctypes.cast(ms_buf, ctypes.c_void_p), ctypes.cast(tkn_buf, ctypes.c_void_p), cap)
    R = c["rank"]
    # records with N < 0 are launches of the adapter-free build (frozen in_proj / out_proj): no low-rank terms
    flops = sum(2.0 * tkn_buf[3 * i] * tkn_buf[3 * i + 1] * abs(tkn_buf[3 * i + 2]) +
                (2.0 * tkn_buf[3 * i] * R * (tkn_buf[3 * i + 1] + tkn_buf[3 * i + 2]) if tkn_buf[3 * i + 2] > 0 else 0.0)
                for i in range(nrec))
    gemm_ms = sum(ms_buf[i] for i in range(nrec))
    shapes = sorted({(tkn_buf[3 * i], tkn_buf[3 * i + 1], tkn_buf[3 * i + 2]) for i in range(nrec)})
    n_adapted = sum(1 for i in range(nrec) if tkn_buf[3 * i + 2] > 0)
    status_nan = int(tr.model.last_status[1].item()) if tr.model.last_status is not None else 0

    parity = aggregation_parity(tr, agg, dist, world, rank, dev) if world > 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long training step)"
    if peak is None:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists() and args.config == 2:
        try:
            traffic = json.loads(tf.read_text()).get("svlora_gemm_dram_bytes_per_launch")
        except Exception:
            traffic = None
    achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {
        "kernel": "ffm::svlora_gemm_kernel / pair::svlora_gemm_pair_kernel (fused FairLoRA linear fwd / dX and the frozen "
                  "attention projections on the same tcgen05 + TMA kernel), FLOP-weighted over every launch of a step",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": "ncu --set full dram__bytes_read+write per launch (profiles/roofline_traffic.json)"
        if traffic else None, "peak_source": peak_src, "launches_timed": nrec,
        "avg_launch_us": 1e3 * gemm_ms / max(nrec, 1),
        "share_of_step": (gemm_ms / prof_steps) / (ms_total / args.steps),
        "launches_adapted": n_adapted,
        "algorithmic_flops_per_launch": f"2*T*K*N + 2*T*r*(K+N), r={R} (adapted linears; N < 0 marks the adapter-free launches of "
                                        f"the frozen in_proj / out_proj: 2*T*K*|N|); (T,K,N) launched: {shapes[:6]}"
                                        + (" ..." if len(shapes) > 6 else ""),
    }
    roofline_hbm = None
    if not args.no_stress:
        try:
            roofline_hbm = sinkhorn_stress(dev, peaks)
        except Exception as e:
            roofline_hbm = {"error": f"{type(e).__name__}: {e}"[:200]}

    cpu = gpu_eager = None
    if world == 1 and not args.no_cpu_baseline:
        res = time_reference(3, 1, batch=args.cpu_batch, ot=args.ot, c=c, budget_s=args.cpu_budget)
        cpu = {"value": res["images_per_s"], "unit": UNIT, "cores": res["threads"], "kind": res["kind"],
               "sample": f"the same training step through the reference's modules on the host: batch {res['batch']}, "
                         f"{res['steps_run']} timed steps after {res['warmup_run']} warm-up, torch-CPU fp32, "
                         f"{res['threads']} threads"}
        try:
            torch.cuda.empty_cache()
            rg = time_reference(10, 3, batch=BATCH, ot=args.ot, c=c, budget_s=60.0, device=str(dev))
            gpu_eager = {"value": rg["images_per_s"], "unit": UNIT, "kind": rg["kind"], "dtype": "f32 (TF32 off)",
                         "ms_per_step": rg["ms_per_step"],
                         "sample": f"the reference's modules moved to this B200, eager PyTorch, batch {rg['batch']}, "
                                   f"{rg['steps_run']} timed steps after {rg['warmup_run']} warm-up"}
        except Exception as e:
            gpu_eager = {"error": f"{type(e).__name__}: {e}"[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(world, BATCH, args.ot, c), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8 + 8 * BATCH,
                "ms_per_step": ms_e2e / args.steps,
                "api": "GLP_OT_SVLoRA.train(idx=0, is_fed=True) over a loader of pinned host batches (per-step loss / acc "
                       "/ training AUC bookkeeping, read back once per epoch)"},
        "gpu_launches": launches, "launches_per_step": launches_per_step, "step_submission": graph_note,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu, "reference_gpu_eager": gpu_eager,
        "agg_parity": parity, "final_loss": final.get("loss"), "final_train_auc": final.get("auc"),
        "plan_nan": status_nan,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config: 2 = configs[1] (headline), 3 = OCT, 4 = RN50, 5 = FedChexMimic")
    ap.add_argument("--ot", default="Sinkhorn", choices=["None", "Sinkhorn", "COT"])
    ap.add_argument("--cpu-batch", type=int, default=BATCH, help="images per step of the CPU arm (the labelled batch)")
    ap.add_argument("--cpu-budget", type=float, default=170.0, help="seconds of CPU work allowed for the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the Sinkhorn stress-shape roofline")
    ap.add_argument("--no-graph", action="store_true", help="submit every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if not args.no_cpu_baseline:
        args.cpu_budget = min(args.cpu_budget, 30.0)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
