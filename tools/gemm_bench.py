"""Launch the fused SVLoRA GEMM a few times on one shape (driver for ncu / timing on the B200 box)."""
import argparse, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=12608); ap.add_argument("--K", type=int, default=768)
ap.add_argument("--N", type=int, default=3072); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--act", type=int, default=0); ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0"); r, B = 12, 64
T, K, N = a.T, a.K, a.N
x = torch.randn(T, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
bias = torch.randn(N, device=dev); A = torch.randn(K, r, device=dev) * 0.05; Bm = torch.randn(r, N, device=dev)
s_eff = torch.rand(B, r, device=dev); y = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
ypre = torch.empty_like(y) if a.act else None
h = torch.zeros(T, 16, device=dev); zz = torch.zeros(T, 16, device=dev, dtype=torch.bfloat16); lib = _cabi.load()
wsb = lib.ffm_svlora_fwd_workspace_bytes(T, K, N, B); ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: 0 if t is None else t.data_ptr()
def fwd():
    _cabi.call("ffm_svlora_fwd", p(x), p(W), p(bias), p(A), p(Bm), p(s_eff), p(y), p(ypre), p(h), p(zz), p(ws), wsb,
               T, K, N, r, B, B, 1, 1, 1.0 / 6, a.act, st)
dy = torch.randn(T, N, device=dev).bfloat16(); Wt = W.t().contiguous(); dx = torch.empty_like(x)
dA = torch.zeros(K, r, device=dev); dB = torch.zeros(r, N, device=dev); dse = torch.zeros(B, r, device=dev)
bwsb = lib.ffm_svlora_bwd_workspace_bytes(T, K, N, B); bws = torch.empty(bwsb, device=dev, dtype=torch.uint8)
def bwd():
    _cabi.call("ffm_svlora_bwd", p(dy), p(x), p(Wt), p(A), p(Bm), p(s_eff), p(h), p(zz), p(ws), 0, p(dx), p(dA), p(dB), p(dse),
               p(bws), bwsb, T, K, N, r, B, B, 1, 1, 1.0 / 6, st)
fn = bwd if a.bwd else fwd
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import ctypes
lib.ffm_profile_enable(1)
e0.record()
for _ in range(a.iters): fn()
e1.record(); torch.cuda.synchronize()
lib.ffm_profile_enable(0)
ms_arr = (ctypes.c_float * 4096)(); tkn = (ctypes.c_int * (3 * 4096))()
n = lib.ffm_profile_read(ms_arr, tkn, 4096)
kern = sorted(ms_arr[i] for i in range(n))
ms = e0.elapsed_time(e1) / a.iters
kmed = kern[len(kern) // 2] if n > 0 else float('nan')
print(f"  GEMM kernel only (events around the launch): median {kmed*1e3:.1f} us min {kern[0]*1e3:.1f} us -> {2.0*T*K*N/kmed/1e9:.1f} TFLOP/s")
# cuBLAS reference on the same box / same clocks (plain x @ W^T, no adapter, no bias)
for _ in range(3): torch.matmul(x, W.t())
torch.cuda.synchronize(); e0.record()
for _ in range(a.iters): torch.matmul(x, W.t())
e1.record(); torch.cuda.synchronize()
cms = e0.elapsed_time(e1) / a.iters
print(f"  cuBLAS x@W^T on this box: {cms*1e3:.1f} us -> {2.0*T*K*N/cms/1e9:.1f} TFLOP/s")
print(f"T={T} K={K} N={N} {'bwd' if a.bwd else 'fwd'} act={a.act}: {ms*1e3:.1f} us  {2.0*T*K*N/ms/1e9:.1f} TFLOP/s (base GEMM flops)")
