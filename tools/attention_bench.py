"""Attention core at the image-tower shape (B=64, L=197, H=12, head dim 64): own kernels (csrc/attention.cu) against
torch SDPA on the same packed qkv, forward and backward, CUDA events, L2 flushed (write + read pass) between repetitions.
Usage on the B200 box:  python tools/attention_bench.py   (also the driver for `ncu -k regex:attention_`)"""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import ops

dev = "cuda:0"
B, L, H, hd = 64, 197, 12, 64
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, L, 3 * H * hd, generator=g).to(dev).to(torch.bfloat16).requires_grad_(True)
d_out = torch.randn(B, L, H * hd, generator=g).to(dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def own():
    return ops.attention(qkv, H, False, True)


def lib():
    q, k, v = qkv.view(B, L, 3, H, hd).unbind(2)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return o.transpose(1, 2).reshape(B, L, H * hd)


def timed(fn, reps=10):
    tf = tb = 0.0
    for i in range(reps + 3):
        flush.zero_(); flush.view(torch.int32).sum()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); o = fn(); e[1].record(); o.backward(d_out); e[2].record()
        torch.cuda.synchronize()
        qkv.grad = None
        if i >= 3:
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    return 1e3 * tf / reps, 1e3 * tb / reps


flops_f = 4.0 * B * H * L * L * hd
for name, fn in (("own (csrc/attention.cu)", own), ("lib (torch SDPA)", lib)):
    f, b = timed(fn)
    print(f"{name:26s} forward {f:7.1f} us ({flops_f / f / 1e6:6.1f} TFLOP/s useful)   backward {b:7.1f} us "
          f"(eager: includes launch gaps and, for lib, the cat of dq/dk/dv)")
