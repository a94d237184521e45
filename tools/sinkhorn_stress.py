"""Persistent Sinkhorn on stress shapes (stand-alone, read K + write T) against the measured HBM bandwidth; the driver for
`ncu -k regex:sinkhorn`.  Usage on the B200 box: python tools/sinkhorn_stress.py [P ...]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import ops

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for P in [int(a) for a in sys.argv[1:]] or [128, 8192, 16384, 65536]:
    M, N = 196, 2
    g = torch.Generator(device=dev).manual_seed(3)
    sim = torch.rand((P, M, N), device=dev, generator=g) * 0.4 + 0.3
    K = torch.exp(-(1.0 - sim) / 0.1)
    ts = []
    for i in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        T, status = ops.sinkhorn(K, mode="Sinkhorn", thresh=1e-3, max_iter=100)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    us = 1e3 * sorted(ts)[len(ts) // 2]
    byts = 2.0 * P * M * N * 4
    print(f"P={P:6d}: {us:8.1f} us, {int(status[0])} iterations, {byts / us / 1e3:7.1f} GB/s "
          f"({byts / us / 1e3 / 6552:.3f} of measured HBM), marginals ok: "
          f"{bool(torch.allclose(T.sum(1), torch.full((P, N), 1.0 / N, device=dev), rtol=1e-3))}")
