"""Bring-up check of the fused SVLoRA GEMM through the C ABI (run on the B200 box via gpurun).

Compares ffm_svlora_fwd / ffm_svlora_bwd against plain torch fp32 math on the same bf16-rounded inputs
and prints a per-(32-row, 64-column) block error map when something is off, so layout / swizzle /
descriptor bugs can be localised from one run.  Not part of the product; tests/ holds the real parity
tests.
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import _cabi  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def sample_of_rows(T, b_prime, num_slices, device):
    t = torch.arange(T, device=device)
    return (t % b_prime) // num_slices


def quick_gelu(u):
    return u * torch.sigmoid(1.702 * u)


def quick_gelu_grad(u):
    s = torch.sigmoid(1.702 * u)
    return s * (1 + 1.702 * u * (1 - s))


def block_map(err, tag, rb=32, cb=64, limit=12):
    T, N = err.shape
    rows = (T + rb - 1) // rb
    cols = (N + cb - 1) // cb
    pad = torch.zeros(rows * rb, cols * cb, device=err.device)
    pad[:T, :N] = err
    m = pad.view(rows, rb, cols, cb).amax(dim=(1, 3)).cpu()
    print(f"  [{tag}] block max-abs-err map ({rb} rows x {cb} cols per cell), first {limit} row blocks:")
    for i in range(min(rows, limit)):
        print("   r%03d " % i + " ".join(f"{v:8.2e}" for v in m[i, : min(cols, 12)].tolist()))
    if rows > limit:
        i = rows - 1
        print("   r%03d " % i + " ".join(f"{v:8.2e}" for v in m[i, : min(cols, 12)].tolist()))


def run_case(T, K, N, r, b_prime, num_slices, act, seed=0, verbose=True, timing=False):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    nS = b_prime // num_slices
    x = (torch.randn(T, K, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev, generator=g) * (K ** -0.5)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    A = torch.randn(K, r, device=dev, generator=g) * 0.05
    B = torch.randn(r, N, device=dev, generator=g)
    s_eff = torch.rand(nS, r, device=dev, generator=g) + 0.1
    scaling = 2.0 / r
    y = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
    y_pre = torch.empty(T, N, device=dev, dtype=torch.bfloat16) if act else None
    RP = 16 if r <= 16 else 32
    h = torch.zeros(T, RP, device=dev)
    zz = torch.zeros(T, RP, device=dev, dtype=torch.bfloat16)
    lib = _cabi.load()
    ws_bytes = lib.ffm_svlora_fwd_workspace_bytes(T, K, N, nS)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)

    def fwd():
        _cabi.call("ffm_svlora_fwd", ptr(x), ptr(W), ptr(bias), ptr(A), ptr(B), ptr(s_eff), ptr(y), ptr(y_pre),
                   ptr(h), ptr(zz), ptr(ws), ws_bytes, T, K, N, r, nS, b_prime, num_slices, 1, scaling, act, stream())

    fwd()
    torch.cuda.synchronize()

    # ---- reference (fp32 math on the same rounded operands) ----
    samp = sample_of_rows(T, b_prime, num_slices, dev)
    A16 = A.to(torch.bfloat16).float()
    B16 = B.to(torch.bfloat16).float()
    xf = x.float()
    h_ref = xf @ A16
    z_ref = (h_ref * (scaling * s_eff)[samp]).to(torch.bfloat16).float()
    u_ref = xf @ W.float().t() + bias + z_ref @ B16
    y_ref = quick_gelu(u_ref) if act else u_ref

    ok = True
    eh = (h[:, :r] - h_ref).abs()
    tol_h = 1e-3 * max(1.0, h_ref.abs().max().item())
    if verbose:
        print(f"case T={T} K={K} N={N} r={r} b'={b_prime} slices={num_slices} act={act}")
        print(f"  h   max err {eh.max().item():.3e} (tol {tol_h:.1e}); pad cols max {(h[:, r:].abs().max().item() if r < RP else 0.0):.2e}")
    if eh.max().item() > tol_h or not torch.isfinite(h).all():
        ok = False
        block_map(eh, "h", rb=32, cb=RP)
    ey = (y.float() - y_ref).abs()
    scale = y_ref.abs().max().item()
    rel = (ey / (2.0 ** -7 * y_ref.abs() + 2e-3 * scale)).max().item() * 2e-2
    if verbose:
        print(f"  y   max abs err {ey.max().item():.3e} (|y|max {scale:.2f}) rel-ish {rel:.3e}")
    if rel > 2e-2 or not torch.isfinite(y.float()).all():
        ok = False
        block_map(ey, "y")
    if act:
        ep = (y_pre.float() - quick_gelu_grad(u_ref)).abs()
        relp = (ep / 1.5e-2).max().item() * 2e-2
        if verbose:
            print(f"  pre max abs err {ep.max().item():.3e} rel-ish {relp:.3e}")
        if relp > 2e-2:
            ok = False
            block_map(ep, "y_pre")

    # ---- backward ----
    dy = (torch.randn(T, N, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    Wt = W.t().contiguous()
    dx = torch.empty(T, K, device=dev, dtype=torch.bfloat16)
    dA = torch.zeros(K, r, device=dev)
    dB = torch.zeros(r, N, device=dev)
    dse = torch.zeros(nS, r, device=dev)
    bws_bytes = lib.ffm_svlora_bwd_workspace_bytes(T, K, N, nS)
    bws = torch.empty(bws_bytes, device=dev, dtype=torch.uint8)
    gelu_pre = (torch.rand(T, K, device=dev, generator=g)).to(torch.bfloat16) if act else None

    def bwd():
        _cabi.call("ffm_svlora_bwd", ptr(dy), ptr(x), ptr(Wt), ptr(A), ptr(B), ptr(s_eff), ptr(h), ptr(zz), ptr(ws),
                   ptr(gelu_pre), ptr(dx), ptr(dA), ptr(dB), ptr(dse), ptr(bws), bws_bytes, T, K, N, r, nS, b_prime, num_slices,
                   1, scaling, stream())

    bwd()
    torch.cuda.synchronize()
    dyf = dy.float()
    dzu_ref = dyf @ B16.t()
    dh_ref = dzu_ref * (scaling * s_eff)[samp]
    dx_ref = dyf @ W.float() + dh_ref.to(torch.bfloat16).float() @ A16.t()
    if act:
        dx_ref = dx_ref * gelu_pre.float()
    dA_ref = xf.t() @ dh_ref
    dB_ref = (h_ref * (scaling * s_eff)[samp]).t() @ dyf
    dse_ref = torch.zeros(nS, r, device=dev).index_add_(0, samp, scaling * dzu_ref * h_ref)

    def rel_err(a, b):
        return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()

    e_dx = (dx.float() - dx_ref).abs()
    r_dx = (e_dx / (2.0 ** -7 * dx_ref.abs() + 2e-3 * dx_ref.abs().max())).max().item() * 2e-2
    r_dA, r_dB, r_ds = rel_err(dA, dA_ref), rel_err(dB, dB_ref), rel_err(dse, dse_ref)
    if verbose:
        print(f"  dx rel-ish {r_dx:.3e}  dA rel {r_dA:.3e}  dB rel {r_dB:.3e}  ds_eff rel {r_ds:.3e}")
    if r_dx > 2e-2 or not torch.isfinite(dx.float()).all():
        ok = False
        block_map(e_dx, "dx")
    if r_dA > 5e-3 or r_dB > 5e-3 or r_ds > 5e-3:
        ok = False
        print("  !! adapter gradient mismatch")

    if timing:
        for name, fn, flops in (
            ("fwd", fwd, 2.0 * T * K * N + 2.0 * T * r * (K + N)),
            ("bwd", bwd, 2.0 * T * K * N + 4.0 * T * r * (K + N) + 2.0 * T * r),
        ):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 20
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"  {name}: {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s (incl. prep/small kernels)")
        # torch reference timing for the plain GEMM
        for _ in range(3):
            torch.matmul(x, W.t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(x, W.t())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"  cuBLAS x@W^T: {ms*1e3:8.1f} us  {2.0*T*K*N/ms/1e9:8.1f} TFLOP/s")
    print("  ->", "OK" if ok else "FAIL")
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--timing", action="store_true")
    args = ap.parse_args()
    t0 = time.time()
    print("lib version", _cabi.load().ffm_version(), "device", torch.cuda.get_device_name(0))
    cases = [
        # T, K, N, r, b_prime, num_slices, act
        (128, 64, 192, 12, 8, 1, 0),       # one tile, one k block
        (128, 128, 192, 12, 8, 1, 0),      # two k blocks
        (256, 256, 384, 12, 8, 1, 0),      # 2x2 tiles
        (200, 192, 400, 12, 8, 1, 0),      # ragged M / N tails
        (1576, 768, 3072, 12, 8, 1, 0),    # cfg1 c_fc
        (1576, 3072, 768, 12, 8, 1, 0),    # cfg1 c_proj
        (1576, 768, 3072, 12, 8, 1, 1),    # fused QuickGELU (+ grad in bwd)
        (1576, 768, 3072, 12, 8, 4, 0),    # OCT style slices (2 samples x 4 slices)
        (788, 768, 3072, 12, 4, 4, 0),     # attr=None style (single sample row)
        (128, 64, 192, 32, 8, 1, 0),       # rank 32 (RN50 recipe): SW64 Z / Bside tiles, two fix-up UMMAs
        (392, 256, 64, 32, 8, 1, 0),       # RN50 layer1 conv3-like 1x1 conv (K=256 -> N=64), 8 images x 7x7
        (1568, 1024, 2048, 32, 8, 1, 0),   # RN50 layer4-like
        (1576, 768, 3072, 20, 8, 1, 1),    # rank 20 -> padded 32, fused QuickGELU
    ]
    if not args.quick:
        cases += [
            (12608, 768, 3072, 12, 64, 1, 0),
            (12608, 3072, 768, 12, 64, 1, 0),
            (12608, 768, 3072, 12, 64, 1, 1),      # c_fc forward as used in the model (fused QuickGELU, dual store)
            (12608, 3072, 768, 12, 64, 1, 1),      # c_proj backward as used in the model (QuickGELU' on dx)
        ]
    all_ok = True
    for c in cases:
        big = c[0] >= 12608
        try:
            ok = run_case(*c, timing=args.timing and big)
        except Exception as e:  # keep going: later cases may still be informative
            print(f"case {c} raised {type(e).__name__}: {e}")
            ok = False
            if "CUDA" in str(e) or "launch" in str(e):
                print("CUDA context likely poisoned; stopping")
                all_ok = False
                break
        all_ok &= ok
    print(f"ALL OK: {all_ok}  ({time.time()-t0:.1f}s)")
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
