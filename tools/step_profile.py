"""Per-kernel time breakdown of one training step (torch.profiler / CUPTI), to decide what to optimise next.
Usage on the B200 box:  python tools/step_profile.py [--steps 3] > gpurun_out/step_profile.txt"""
import argparse, sys, collections
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench  # noqa: E402
import step_ablation  # noqa: E402

ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--ot", default="Sinkhorn")
a = ap.parse_args()
dev = torch.device("cuda:0")
tr, _c = step_ablation.build("baseline")       # configs[1], batch 64, Sinkhorn head (eager launches are profiled here)
g = torch.Generator().manual_seed(0)
batch = {"img": torch.randint(0, 256, (bench.BATCH, 1, 224, 224), generator=g).float().repeat(1, 3, 1, 1).to(dev),
         "label": (torch.arange(bench.BATCH) % 2).to(dev), "attrs": torch.randint(0, 3, (bench.BATCH, 1), generator=g).to(dev)}
for _ in range(5): tr.forward_backward(batch)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(a.steps): tr.forward_backward(batch)
t_cpu = (time.perf_counter() - t0) / a.steps
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / a.steps
print(f"wall per step: {t_all*1e3:.2f} ms; CPU enqueue time per step: {t_cpu*1e3:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps): tr.forward_backward(batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        agg[ev.name][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"total GPU kernel time per step: {tot/a.steps/1e3:.2f} ms over {sum(v[1] for v in agg.values())//a.steps} launches")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{us/a.steps/1e3:8.3f} ms {100*us/tot:5.1f}% {n//a.steps:5d}x {us/n:8.1f} us  {name[:110]}")
