"""Time the GLP_OT head kernels at the config-2 shape (B=64: 12 544 tokens x 512, 128 Sinkhorn problems) with CUPTI.
Usage on the B200 box:  [FFM_SK_GRID=n] python tools/head_bench.py"""
import collections, sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import ops

dev = "cuda:0"
g = torch.Generator().manual_seed(11)
feats = torch.randn(64, 197, 512, generator=g).to(dev).bfloat16().requires_grad_(True)   # batch-first, as the tower stores it
txt = torch.randn(4, 512, generator=g).to(dev).requires_grad_(True)
ls = torch.tensor(float(np.log(1 / 0.07)), device=dev, requires_grad=True)
dl = torch.randn(64, 2, generator=g).to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def step():
    flush.zero_()
    logits, status, T = ops.ot_head(feats, txt, ls, n_cls=2, ot="Sinkhorn", batch_first=True)
    (logits * dl).sum().backward()
    return status


for _ in range(3):
    st = step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
n = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and "ffm::" in ev.name:
        agg[ev.name][0] += ev.device_time
        agg[ev.name][1] += 1
print("iterations", st.cpu().tolist())
row_bytes = 64 * 197 * 512 * 2
for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    per = us / cnt
    extra = ""
    if "sim_kernel" in name:
        extra = f"  {row_bytes / per / 1e3:.0f} GB/s (features read once)"
    if "head_bwd" in name:
        extra = f"  {2 * row_bytes / per / 1e3:.0f} GB/s (features read + gradient written)"
    print(f"{per:8.1f} us x{cnt // n}  {name[:90]}{extra}")
