#!/usr/bin/env python
"""bench.py — FairLoRA ViT-B/16 training throughput on B200 (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`
(one rank per GPU = one simulated site, NCCL).  Rank 0 prints ONE JSON line.

A "step" = one GLP_OT_SVLoRA training iteration of one site on one batch of 64 synthetic SLO images:
pixel normalisation -> CLIP ViT-B/16 image encoder with 24 fused FairLoRA linears -> text encoder (4 prompts) ->
GLP_OT Sinkhorn head -> cross-entropy -> backward -> SGD stepped twice (reference quirk F6).  The per-round FedAvg
(all-reduce of U, V, s_g over NVLink + EMA / shared-half-S epilogue) runs once inside the timed region.
  value : images/s with the batches already resident in HBM (device timed, CUDA events, max over ranks)
  e2e   : same metric through the trainer's public forward_backward() with HOST (pinned) batches: the H2D copy of
          every batch and a D2H read of the step's loss are inside the timed region
  roofline : the dominant kernel (fused SVLoRA tcgen05 GEMM): algorithmic FLOPs per launch / CUDA-event duration per
          launch measured live (ffm_profile_*, events on the launching stream) in a second short profiled region
  cpu_baseline : the oracle port (oracle/ref_port.py, the reference's algorithm in torch-CPU fp32) timed on this box's
          host cores on a bounded sample (config-1 shape, batch 8) — reported beside, not the target
--impl reference times that same oracle port with all host threads as the reference arm.
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
RANK_R, GROUPS = 12, 3


def make_cfg(world: int, batch: int, ot: str):
    from fairfedmed_b200.config import get_cfg_default
    cfg = get_cfg_default()
    cfg.DATASET.USERS = max(world, 1)
    cfg.DATASET.NUM_TRAIN_PER_CLIENT = batch * 2     # loaders are bypassed by the bench (fixed synthetic pool)
    cfg.DATASET.NUM_TEST_PER_CLIENT = batch
    cfg.DATALOADER.TRAIN_X.BATCH_SIZE = batch
    cfg.TRAINER.GLP_OT.OT = ot
    return cfg


def workload_config(world, batch, ot):
    return {
        "workload": "configs[1]: FairLoRA (GLP_OT_SVLoRA) CLIP ViT-B/16, synthetic 2D SLO 224x224, one site per GPU, "
                    f"batch {batch}/GPU, race attribute (3 groups), rank 12 alpha 2, OT={ot}, FedAvg of U,V,s_g per round",
        "global_batch": batch * world, "batch_per_gpu": batch, "sites": world, "ot": ot,
        "parallelism": f"sites{world}", "optimizer": "SGD lr1e-3 m0.9 wd5e-4, stepped twice per iteration (reference F6)",
        "l2_policy": "no explicit flush: one step touches >2 GB of activations/weights (>> 126 MB L2)",
    }


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


# --------------------------------------------------------------------------------------------- oracle (CPU) arm
def oracle_step_fn(batch: int, ot: str, seed: int = 1):
    """Build the oracle's parameters (random-init ViT-B/16 + FairLoRA adapters) and return step() -> loss."""
    import torch.nn.functional as F
    from fairfedmed_b200.clip_model import CustomCLIP
    from fairfedmed_b200.modules import apply_lora_to_model
    from oracle import ref_port as rp
    torch.manual_seed(seed)
    model = CustomCLIP(ot=ot)
    for n, p in model.named_parameters():
        p.requires_grad_("prompt_learner" in n)
    apply_lora_to_model(model, True, rank=RANK_R, alpha=2.0, lora_type="FairLoRA", num_attrs=GROUPS)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    params = {k: v.detach().clone().float() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    for k in names:
        if "lora_A" in k:
            params[k] = 0.02 * torch.randn(params[k].shape, generator=g)
        params[k].requires_grad_(True)
    eot = model.prompt_learner.eot_index.clone()
    del model
    image = torch.randint(0, 256, (batch, 1, 224, 224), generator=g).float().repeat(1, 3, 1, 1)
    label = (torch.arange(batch) % 2).long()
    attr = torch.randint(0, GROUPS, (batch,), generator=g)
    bufs = [None] * len(names)

    def step():
        logits = rp.custom_clip_forward(image, attr, params, eot, ot=ot, scaling=2.0 / RANK_R)
        loss = F.cross_entropy(logits, label)
        grads = torch.autograd.grad(loss, [params[k] for k in names])
        rp.sgd_double_step([params[k] for k in names], grads, bufs, lr=1e-3)
        return float(loss)

    return step


def time_oracle(steps: int, warmup: int, batch: int, ot: str, budget_s: float):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step = oracle_step_fn(batch, ot)
    t0 = time.perf_counter()
    step()                                            # first call also pays one-off allocation
    probe = time.perf_counter() - t0
    done_w = 1
    while done_w < warmup and (time.perf_counter() - t0) < budget_s * 0.3:
        step(); done_w += 1
    per = probe
    n = max(1, min(steps, int((budget_s * 0.7) / max(per, 1e-3))))
    t1 = time.perf_counter()
    for _ in range(n):
        step()
    dt = time.perf_counter() - t1
    return {"images_per_s": batch * n / dt, "ms_per_step": 1e3 * dt / n, "steps_run": n, "warmup_run": done_w,
            "threads": threads, "batch": batch}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ot = args.ot
    res = time_oracle(args.steps, args.warmup, batch=args.cpu_batch, ot=ot, budget_s=args.cpu_budget)
    sample = (f"oracle port (oracle/ref_port.py, torch-CPU fp32) of the same training step, bounded sample: batch "
              f"{res['batch']} images/step, {res['steps_run']} timed steps after {res['warmup_run']} warm-up "
              f"(requested {args.steps}/{args.warmup}, capped to a {args.cpu_budget:.0f}s budget), {res['threads']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": res["images_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": res["steps_run"], "warmup": res["warmup_run"], "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(max(args.gpus, 1), BATCH, ot),
        "cpu_baseline": {"value": res["images_per_s"], "unit": UNIT, "cores": res["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": res["images_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------- our arm
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
    from fairfedmed_b200.fed_utils import FederatedAggregator
    from fairfedmed_b200.registry import build_trainer

    lib = _cabi.load()
    cfg = make_cfg(world, BATCH, args.ot)
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
    agg = FederatedAggregator(tr.flat_spec)
    n_k = BATCH * 20
    n_kg = [n_k // 3, n_k // 3, n_k - 2 * (n_k // 3)]

    # synthetic pool: pinned host batches (e2e) and device-resident copies (value)
    pool = 4
    gen = torch.Generator().manual_seed(100 + rank)
    host, devb = [], []
    for _ in range(pool):
        img = torch.randint(0, 256, (BATCH, 1, 224, 224), generator=gen).float().repeat(1, 3, 1, 1).pin_memory()
        lab = (torch.arange(BATCH) % 2).long().pin_memory()
        att = torch.randint(0, GROUPS, (BATCH, 1), generator=gen).pin_memory()
        host.append({"img": img, "label": lab, "attrs": att})
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

    def timed(fn_step, steps, warmup, with_round=True, finalize=None):
        for i in range(warmup):
            fn_step(i)
        if finalize is not None:
            finalize()
        if with_round:
            fed_round(0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.ffm_launch_count(1)
        e0.record()
        for i in range(steps):
            fn_step(i)
        if with_round:
            fed_round(1)
        if finalize is not None:
            finalize()
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
    run_step = tr.forward_backward_graphed if use_graph else tr.forward_backward

    # ---- value: inputs resident in HBM ----
    def step_device(i):
        run_step(devb[i % pool])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_device, args.steps, args.warmup)
    if use_graph:       # replayed kernels do not pass through the library's launch counter: add them back
        launches += launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = BATCH * world * args.steps / (ms_total / 1e3)

    # ---- e2e: host batches through the public trainer API, H2D inside, loss read back every step ----
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}
    # three rotating device staging slots (allocated once: no allocator traffic, no cudaMalloc inside the timed region)
    n_slots = 3
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(n_slots)]
    slot_free = [None] * n_slots                       # event: the step that last read the slot has consumed it

    def stage(i):
        sl = i % n_slots
        with torch.cuda.stream(copy_stream):
            if slot_free[sl] is not None:
                copy_stream.wait_event(slot_free[sl])
            for k, v in host[i % pool].items():
                slots[sl][k].copy_(v, non_blocking=True)                # H2D from pinned memory, every step
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (slots[sl], ev)

    losses = []
    loss_ring = torch.zeros(8, dtype=torch.float32).pin_memory()
    in_flight = []                                     # (slot, event) of losses whose D2H copy has been issued

    def drain(keep):
        while len(in_flight) > keep:
            slot, ev = in_flight.pop(0)
            ev.synchronize()
            losses.append(float(loss_ring[slot]))      # the host reads EVERY step's loss inside the timed region

    host_ts = []

    def step_host(i):
        host_ts.append(time.perf_counter())
        if i not in staged:
            stage(i)
        batch, ev = staged.pop(i)
        stage(i + 1)                                   # prefetch the next batch while this one computes
        torch.cuda.current_stream().wait_event(ev)
        out = run_step(batch)
        consumed = torch.cuda.Event()
        consumed.record()
        slot_free[i % n_slots] = consumed
        slot = i % 8
        loss_ring[slot:slot + 1].copy_(out["loss"].detach().reshape(1), non_blocking=True)   # D2H, every step
        done = torch.cuda.Event()
        done.record()
        in_flight.append((slot, done))
        drain(keep=1)                                  # read step i-1's loss while step i runs (no pipeline bubble)

    staged.clear()
    import gc
    gc.collect()
    gc.disable()                                       # a collector pause inside a 20-step region is a 1 ms/step artefact
    try:
        ms_e2e, _ = timed(step_host, args.steps, max(args.warmup, 3), finalize=lambda: drain(keep=0))
    finally:
        gc.enable()
    staged.clear()
    e2e_value = BATCH * world * args.steps / (ms_e2e / 1e3)
    if rank == 0 and len(host_ts) > args.steps:      # diagnostics only (stderr): host-side gaps between e2e steps
        gaps = [1e3 * (b - a) for a, b in zip(host_ts[-args.steps:-1], host_ts[-args.steps + 1:])]
        if gaps:
            print(f"[bench] e2e host step gaps ms: median {statistics.median(gaps):.2f} max {max(gaps):.2f} "
                  f"(n={len(gaps)})", file=sys.stderr)

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream ----
    lib.ffm_profile_enable(1)
    prof_steps = 3
    for i in range(prof_steps):
        tr.forward_backward(devb[i % pool])
    torch.cuda.synchronize()
    lib.ffm_profile_enable(0)
    cap = 96 * prof_steps + 8
    ms_buf = (ctypes.c_float * cap)()
    tkn_buf = (ctypes.c_int * (3 * cap))()
    nrec = lib.ffm_profile_read(ctypes.cast(ms_buf, ctypes.c_void_p), ctypes.cast(tkn_buf, ctypes.c_void_p), cap)
    flops = sum(2.0 * tkn_buf[3 * i] * tkn_buf[3 * i + 1] * tkn_buf[3 * i + 2] +
                2.0 * tkn_buf[3 * i] * RANK_R * (tkn_buf[3 * i + 1] + tkn_buf[3 * i + 2]) for i in range(nrec))
    gemm_ms = sum(ms_buf[i] for i in range(nrec))
    status_nan = int(tr.model.last_status[1].item()) if tr.model.last_status is not None else 0

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
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("svlora_gemm_dram_bytes_per_launch")
        except Exception:
            traffic = None
    achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {
        "kernel": "ffm::svlora_gemm_kernel (fused FairLoRA linear fwd / dX, tcgen05 + TMA)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "launches_timed": nrec,
        "avg_launch_us": 1e3 * gemm_ms / max(nrec, 1),
        "share_of_step": (gemm_ms / prof_steps) / (ms_total / args.steps),
        "algorithmic_flops_per_launch": "2*T*K*N + 2*T*r*(K+N), T=12608, (K,N) in {(768,3072),(3072,768)}, r=12",
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        res = time_oracle(3, 1, batch=args.cpu_batch, ot=args.ot, budget_s=args.cpu_budget)
        cpu = {"value": res["images_per_s"], "unit": UNIT, "cores": res["threads"], "kind": "port",
               "sample": f"oracle port of the same step at config-1 shape: batch {res['batch']}, {res['steps_run']} timed "
                         f"steps after {res['warmup_run']} warm-up, torch-CPU fp32, {res['threads']} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": workload_config(world, BATCH, args.ot), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "launches_per_step": launches_per_step, "step_submission": graph_note,
        "roofline": roofline, "cpu_baseline": cpu,
        "final_loss": losses[-1] if losses else None, "plan_nan": status_nan,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ot", default="Sinkhorn", choices=["None", "Sinkhorn", "COT"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds of CPU work allowed for the oracle arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="submit every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "ours" and not args.no_cpu_baseline:
        args.cpu_budget = min(args.cpu_budget, 30.0)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
