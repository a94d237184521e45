"""Idle gaps inside the replayed step graph: CUPTI kernel records (torch.profiler) of graph replays, per stream.
Prints, for one step: wall, busy time per stream, union of all streams, and the gaps between consecutive kernels of the
busiest (critical) stream.   python tools/graph_gaps.py
"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import step_ablation  # noqa: E402
from fairfedmed_b200.config import ATTRIBUTE_GROUPS  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

tr, c = step_ablation.build("baseline")
gen = torch.Generator().manual_seed(100)
groups = [len(ATTRIBUTE_GROUPS[c["dataset"]][a]) for a in c["attributes"]]
pool = [{k: v.to("cuda:0") for k, v in bench.synthetic_batch(gen, bench.BATCH, c, groups).items()} for _ in range(2)]
for i in range(2):
    tr.forward_backward(pool[i])
tr.capture_step_graph(pool[0])
for i in range(5):
    tr.forward_backward_graphed(pool[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        tr.forward_backward_graphed(pool[i % 2])
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
        ev.append((e.time_range.start, e.time_range.end, getattr(e, "stream", None) or getattr(e, "device_resource_id", 0), e.name))
ev.sort()
sgd = [i for i, x in enumerate(ev) if "sgd" in x[3].lower()]
print("kernels recorded", len(ev), "sgd markers", len(sgd))
a, b = sgd[0] + 1, sgd[1] + 1
step = ev[a:b]
t0, t1 = step[0][0], max(x[1] for x in step)
print(f"step wall {t1 - t0:.1f} us, {len(step)} kernels")
per = collections.defaultdict(list)
for s, e, st, nm in step:
    per[st].append((s, e, nm))
busy = {st: sum(e - s for s, e, _ in v) for st, v in per.items()}
for st, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print(f"  stream {st}: {len(per[st])} kernels, busy {v:.1f} us")
# union
iv = sorted((s, e) for s, e, _, _ in step)
u, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        u += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
u += cur_e - cur_s
print(f"union of all streams busy {u:.1f} us -> {t1 - t0 - u:.1f} us with nothing running")
main = max(busy, key=busy.get)
m = sorted(per[main])
gaps = [(m[i + 1][0] - m[i][1], m[i][2][:50], m[i + 1][2][:50]) for i in range(len(m) - 1)]
tot = sum(g for g, _, _ in gaps if g > 0)
print(f"critical stream {main}: sum of gaps {tot:.1f} us over {len(gaps)} hand-offs, median {sorted(g for g, _, _ in gaps)[len(gaps) // 2]:.2f} us")
for g, x, y in sorted(gaps, reverse=True)[:12]:
    print(f"   {g:7.2f} us  after {x}  before {y}")

agg = collections.defaultdict(lambda: [0, 0.0])
for s_, e_, st, nm in step:
    k = nm.split("(")[0].replace("void ", "")[:70]
    agg[k][0] += 1
    agg[k][1] += e_ - s_
print("per-kernel device time inside the replayed step (streams overlap, so the sum exceeds the wall):")
for k, (n_, t_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {t_:8.1f} us  x{n_:3d}  avg {t_ / n_:6.1f}  {k}")
