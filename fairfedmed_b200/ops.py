"""torch custom ops over the C ABI (libffm_b200.so) + their autograd wiring.

Every op passes raw device pointers, sizes and the current CUDA stream to the shared library through ctypes
(`_cabi`).  PyTorch only provides memory, streams and autograd bookkeeping here; there is no eager fallback:
a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _cabi

def padded_rank(r: int) -> int:
    """Row length of the h / z side outputs for adapter rank r (ffm_svlora_padded_rank: 16 for r <= 16, else 32)."""
    if r < 1 or r > 32:
        raise _cabi.FfmError(f"adapter rank {r} not supported by the fused kernel (1..32)")
    return 16 if r <= 16 else 32

OT_MODES = {"None": 0, "Sinkhorn": 1, "COT": 2}


def _ptr(t: Optional[Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _cabi.FfmError("fairfedmed_b200 ops run on CUDA tensors only (no CPU fallback)")


# =====================================================================================================
# raw ops
# =====================================================================================================
@torch.library.custom_op("ffm::svlora_fwd", mutates_args=())
def svlora_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor], lora_a: Tensor, lora_b: Tensor, s_eff: Tensor,
               scaling: float, b_prime: int, num_slices: int, act: int,
               row_div: int = 1, prepared: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """y, y_dact, h, z, tiles = fused FairLoRA linear.  x [T,K] bf16, w [N,K] bf16, lora_a [K,r], lora_b [r,N],
    s_eff [nS,r].  y_dact = QuickGELU'(u) when act = 1 (else empty); h = x·A (f32 [T,rp]); z = bf16(scaling·h⊙s_eff)
    ([T,rp], rp = padded_rank(r)); tiles = the call's workspace holding the prepared bf16 adapter tiles (hand it to svlora_bwd).
    Row t of x belongs to sample ((t // row_div) % b_prime) // num_slices (row_div = 1: sequence-first rows).
    `prepared`: workspace already filled by svlora_prepare for these (lora_a, lora_b, s_eff): no preparation launch;
    the returned `tiles` is then empty (keep using `prepared`)."""
    _need_cuda(x, w, lora_a, lora_b, s_eff)
    T, K = x.shape
    N = w.shape[0]
    r = lora_a.shape[1]
    nS = s_eff.shape[0]
    y = torch.empty((T, N), device=x.device, dtype=torch.bfloat16)
    y_dact = torch.empty((T, N), device=x.device, dtype=torch.bfloat16) if act else y.new_empty((0,))
    rp = padded_rank(r)
    h = torch.empty((T, rp), device=x.device, dtype=torch.float32)
    z = torch.empty((T, rp), device=x.device, dtype=torch.bfloat16)
    lib = _cabi.load()
    ws_bytes = lib.ffm_svlora_fwd_workspace_bytes(T, K, N, nS)
    if prepared is not None:
        if prepared.numel() < ws_bytes or not prepared.is_cuda:
            raise _cabi.FfmError("svlora_fwd: prepared workspace too small")
        _cabi.call("ffm_svlora_fwd", _ptr(x), _ptr(w), _ptr(bias), 0, 0, 0, _ptr(y), _ptr(y_dact) if act else 0, _ptr(h),
                   _ptr(z), _ptr(prepared), prepared.numel(), T, K, N, r, nS, b_prime, num_slices, int(row_div),
                   float(scaling), int(act), _stream())
        return y, y_dact, h, z, y.new_empty((0,), dtype=torch.uint8)
    ws = torch.empty((ws_bytes,), device=x.device, dtype=torch.uint8)
    _cabi.call("ffm_svlora_fwd", _ptr(x), _ptr(w), _ptr(bias), _ptr(lora_a), _ptr(lora_b), _ptr(s_eff), _ptr(y),
               _ptr(y_dact) if act else 0, _ptr(h), _ptr(z), _ptr(ws), ws_bytes, T, K, N, r, nS, b_prime, num_slices,
               int(row_div), float(scaling), int(act), _stream())
    return y, y_dact, h, z, ws


@torch.library.custom_op("ffm::svlora_prepare", mutates_args=())
def svlora_prepare(lora_a: Tensor, lora_b: Tensor, s_eff: Tensor, scaling: float) -> Tensor:
    """Workspace with the bf16 adapter tiles of both directions + scaled singular values (ffm_svlora_prepare).  Depends on
    parameters and attribute rows only: issue it ahead of the layer (side stream) and pass it to svlora_fwd(prepared=)."""
    _need_cuda(lora_a, lora_b, s_eff)
    K, r = lora_a.shape
    N = lora_b.shape[1]
    nS = s_eff.shape[0]
    lib = _cabi.load()
    ws_bytes = lib.ffm_svlora_fwd_workspace_bytes(0, K, N, nS)
    ws = torch.empty((ws_bytes,), device=lora_a.device, dtype=torch.uint8)
    _cabi.call("ffm_svlora_prepare", _ptr(lora_a), _ptr(lora_b), _ptr(s_eff), _ptr(ws), ws_bytes, K, N, r, nS,
               float(scaling), _stream())
    return ws


@svlora_prepare.register_fake
def _(lora_a, lora_b, s_eff, scaling):
    return lora_a.new_empty((1,), dtype=torch.uint8)


@svlora_fwd.register_fake
def _(x, w, bias, lora_a, lora_b, s_eff, scaling, b_prime, num_slices, act, row_div=1, prepared=None):
    T, N = x.shape[0], w.shape[0]
    y = x.new_empty((T, N))
    rp = padded_rank(lora_a.shape[1])
    return (y, (x.new_empty((T, N)) if act else x.new_empty((0,))), x.new_empty((T, rp), dtype=torch.float32),
            x.new_empty((T, rp)), x.new_empty((1,), dtype=torch.uint8))


@torch.library.custom_op("ffm::svlora_bwd", mutates_args=())
def svlora_bwd(dy: Tensor, x: Tensor, w_t: Tensor, lora_a: Tensor, lora_b: Tensor, s_eff: Tensor, h: Tensor,
               z: Tensor, tiles: Optional[Tensor], gelu_dact: Optional[Tensor], scaling: float, b_prime: int,
               num_slices: int, row_div: int = 1) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """dx, d_lora_a, d_lora_b, d_s_eff.  dy [T,N] bf16, x [T,K] bf16, w_t [K,N] bf16 (transposed frozen weight);
    h, z, tiles are the side outputs of svlora_fwd (tiles may be None: the adapter tiles are then prepared again)."""
    _need_cuda(dy, x, w_t)
    T, N = dy.shape
    K = x.shape[1]
    r = lora_a.shape[1]
    nS = s_eff.shape[0]
    dx = torch.empty((T, K), device=x.device, dtype=torch.bfloat16)
    dA = torch.empty((K, r), device=x.device, dtype=torch.float32)
    dB = torch.empty((r, N), device=x.device, dtype=torch.float32)
    dse = torch.empty((nS, r), device=x.device, dtype=torch.float32)
    lib = _cabi.load()
    ws_bytes = lib.ffm_svlora_bwd_workspace_bytes(T, K, N, nS)
    ws = torch.empty((ws_bytes,), device=x.device, dtype=torch.uint8)
    _cabi.call("ffm_svlora_bwd", _ptr(dy), _ptr(x), _ptr(w_t), _ptr(lora_a), _ptr(lora_b), _ptr(s_eff), _ptr(h),
               _ptr(z), _ptr(tiles), _ptr(gelu_dact), _ptr(dx), _ptr(dA), _ptr(dB), _ptr(dse), _ptr(ws), ws_bytes,
               T, K, N, r, nS, b_prime, num_slices, int(row_div), float(scaling), _stream())
    return dx, dA, dB, dse


@svlora_bwd.register_fake
def _(dy, x, w_t, lora_a, lora_b, s_eff, h, z, tiles, gelu_dact, scaling, b_prime, num_slices, row_div=1):
    return (x.new_empty(x.shape), lora_a.new_empty(lora_a.shape), lora_b.new_empty(lora_b.shape),
            s_eff.new_empty(s_eff.shape))


@torch.library.custom_op("ffm::svlora_bwd_into", mutates_args=("dA", "dB"))
def svlora_bwd_into(dy: Tensor, x: Tensor, w_t: Tensor, lora_a: Tensor, lora_b: Tensor, s_eff: Tensor, h: Tensor,
                    z: Tensor, tiles: Optional[Tensor], gelu_dact: Optional[Tensor], dA: Tensor, dB: Tensor,
                    scaling: float, b_prime: int, num_slices: int, row_div: int = 1) -> Tuple[Tensor, Tensor]:
    """svlora_bwd with the adapter gradients OVERWRITING caller-owned fp32 buffers dA [K,r] / dB [r,N] (the trainer's
    flat gradient buffer: no allocation, no accumulate kernel).  Returns dx, d_s_eff."""
    _need_cuda(dy, x, w_t, dA, dB)
    T, N = dy.shape
    K = x.shape[1]
    r = lora_a.shape[1]
    nS = s_eff.shape[0]
    if not (dA.is_contiguous() and dB.is_contiguous() and dA.dtype == torch.float32 and dB.dtype == torch.float32
            and tuple(dA.shape) == (K, r) and tuple(dB.shape) == (r, N)):
        raise _cabi.FfmError("svlora_bwd_into: dA / dB must be contiguous fp32 [K,r] / [r,N]")
    dx = torch.empty((T, K), device=x.device, dtype=torch.bfloat16)
    dse = torch.empty((nS, r), device=x.device, dtype=torch.float32)
    lib = _cabi.load()
    ws_bytes = lib.ffm_svlora_bwd_workspace_bytes(T, K, N, nS)
    ws = torch.empty((ws_bytes,), device=x.device, dtype=torch.uint8)
    _cabi.call("ffm_svlora_bwd", _ptr(dy), _ptr(x), _ptr(w_t), _ptr(lora_a), _ptr(lora_b), _ptr(s_eff), _ptr(h),
               _ptr(z), _ptr(tiles), _ptr(gelu_dact), _ptr(dx), _ptr(dA), _ptr(dB), _ptr(dse), _ptr(ws), ws_bytes,
               T, K, N, r, nS, b_prime, num_slices, int(row_div), float(scaling), _stream())
    return dx, dse


@svlora_bwd_into.register_fake
def _(dy, x, w_t, lora_a, lora_b, s_eff, h, z, tiles, gelu_dact, dA, dB, scaling, b_prime, num_slices, row_div=1):
    return x.new_empty(x.shape), s_eff.new_empty(s_eff.shape)


_PENDING_DIRECT_WRITES: list = []     # events behind in-place gradient writes that autograd does not know about


def join_direct_grad_writes() -> None:
    """Make the current stream wait for every in-place gradient write issued by the last backward() (call it after
    loss.backward() and before the optimizer reads the flat gradient buffer).  A no-op without direct gradients."""
    if _PENDING_DIRECT_WRITES:
        cur = torch.cuda.current_stream()
        for ev in _PENDING_DIRECT_WRITES:
            cur.wait_event(ev)
        _PENDING_DIRECT_WRITES.clear()


def _direct_grad(param: Tensor) -> Optional[Tensor]:
    """Gradient buffer a trainer registered for direct writes (`param._ffm_direct_grad = view of its flat gradient
    buffer`), or None: then the gradient is returned to autograd as usual."""
    return getattr(param, "_ffm_direct_grad", None)


def _svlora_bwd_dispatch(dy, x, w_t, lora_a, lora_b, s_eff, h, z, tiles, gelu_dact, scaling, b_prime, num_slices,
                         row_div):
    """-> dx, dA or None, dB or None, d_s_eff.  With registered direct-gradient buffers on BOTH adapter matrices the
    kernels write there (overwrite semantics: one backward per zero_grad, which is what the trainer does) and autograd
    gets None, which removes two fp32 accumulate launches per layer."""
    ga, gb = _direct_grad(lora_a), _direct_grad(lora_b)
    if ga is not None and gb is not None:
        dx, dse = svlora_bwd_into(dy, x, w_t, lora_a, lora_b, s_eff, h, z, tiles, gelu_dact, ga, gb, scaling, b_prime,
                                  num_slices, row_div)
        return dx, None, None, dse
    return svlora_bwd(dy, x, w_t, lora_a, lora_b, s_eff, h, z, tiles, gelu_dact, scaling, b_prime, num_slices, row_div)


@torch.library.custom_op("ffm::seff", mutates_args=())
def seff_op(attr: Optional[Tensor], S: Tensor, S_global: Optional[Tensor], lam: float) -> Tensor:
    """s_eff [nS, r] = pi(attr) @ S (+ S_global);  attr int64 [nS] on the device or None (=> nS = 1, pi = 1/G)."""
    _need_cuda(S, attr, S_global)
    G, r = S.shape
    nS = 1 if attr is None else attr.shape[0]
    out = torch.empty((nS, r), device=S.device, dtype=torch.float32)
    _cabi.call("ffm_seff", _ptr(attr), _ptr(S), _ptr(S_global), _ptr(out), nS, G, r, float(lam), _stream())
    return out


@seff_op.register_fake
def _(attr, S, S_global, lam):
    return S.new_empty((1 if attr is None else attr.shape[0], S.shape[1]))


@torch.library.custom_op("ffm::ds", mutates_args=())
def ds_op(attr: Optional[Tensor], ds_eff: Tensor, G: int, lam: float, want_global: bool) -> Tuple[Tensor, Tensor]:
    nS, r = ds_eff.shape
    dS = torch.empty((G, r), device=ds_eff.device, dtype=torch.float32)
    dSg = torch.empty((r,), device=ds_eff.device, dtype=torch.float32) if want_global else dS.new_empty((0,))
    _cabi.call("ffm_ds", _ptr(attr), _ptr(ds_eff), _ptr(dS), _ptr(dSg) if want_global else 0, nS, G, r, float(lam),
               _stream())
    return dS, dSg


@ds_op.register_fake
def _(attr, ds_eff, G, lam, want_global):
    r = ds_eff.shape[1]
    return ds_eff.new_empty((G, r)), ds_eff.new_empty((r,) if want_global else (0,))


# =====================================================================================================
# autograd wiring
# =====================================================================================================
class _SEff(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, S, S_global, lam):
        ctx.attr, ctx.lam, ctx.G = attr, lam, S.shape[0]
        ctx.set_materialize_grads(False)      # a consumer that finished dS itself hands back None, not zeros
        ctx.has_global = S_global is not None
        # trainer-registered gradient views (see _direct_grad): dS is then written in place, autograd gets None
        ctx.direct = (_direct_grad(S), None if S_global is None else _direct_grad(S_global))
        ctx.sg_shape = None if S_global is None else S_global.shape
        sg = None if S_global is None else S_global.reshape(-1).contiguous()
        out = seff_op(attr, S.contiguous(), sg, lam)
        gS, gSg = ctx.direct
        if gS is not None and (S_global is None or gSg is not None):
            # everything a consumer needs to finish dS itself (the split backward of the fused MLP does, on its
            # parameter-gradient stream; it then returns no gradient for s_eff and this node's backward never runs)
            out._ffm_ds_ctx = (attr, float(lam), int(S.shape[0]), gS, gSg)
        return out

    @staticmethod
    def backward(ctx, ds_eff):
        if ds_eff is None:                    # dS already written by the consumer (split MLP backward)
            return None, None, None, None
        gS, gSg = ctx.direct
        if gS is not None and (not ctx.has_global or gSg is not None):
            ds_eff = ds_eff.contiguous()
            nS, r = ds_eff.shape
            _cabi.call("ffm_ds", _ptr(ctx.attr), _ptr(ds_eff), _ptr(gS), _ptr(gSg) if ctx.has_global else 0, nS, ctx.G,
                       r, float(ctx.lam), _stream())
            # No AccumulateGrad node runs for S, so the autograd engine does not join this node's stream (the forward
            # may have run on a side stream, clip_model.Transformer._prepare_adapters) with the caller's at the end of
            # backward(): leave an event for join_direct_grad_writes().
            ev = torch.cuda.Event()
            ev.record()
            _PENDING_DIRECT_WRITES.append(ev)
            return None, None, None, None
        dS, dSg = ds_op(ctx.attr, ds_eff.contiguous(), ctx.G, ctx.lam, ctx.has_global)
        return None, dS, (dSg.reshape(ctx.sg_shape) if ctx.has_global else None), None


def effective_singular_values(attr: Optional[Tensor], S: Tensor, S_global: Optional[Tensor] = None,
                              lam: float = 0.7) -> Tensor:
    """Differentiable s_eff (trainers/GLP_OT_SVLoRA.py:453-467)."""
    return _SEff.apply(attr, S, S_global, lam)


class _SVLoRALinear(torch.autograd.Function):
    """y = x W^T + b + scaling ((x A) ⊙ s_eff[sample]) B — one fused kernel each way."""

    @staticmethod
    def forward(ctx, x2d, w, w_t, bias, lora_a, lora_b, s_eff, scaling, b_prime, num_slices, row_div):
        y, _, h, z, tiles = svlora_fwd(x2d, w, bias, lora_a, lora_b, s_eff, scaling, b_prime, num_slices, 0, row_div)
        ctx.save_for_backward(x2d, w_t, lora_a, lora_b, s_eff, h, z, tiles)
        ctx.cfg = (scaling, b_prime, num_slices, row_div)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, w_t, lora_a, lora_b, s_eff, h, z, tiles = ctx.saved_tensors
        scaling, b_prime, num_slices, row_div = ctx.cfg
        dx, dA, dB, dse = _svlora_bwd_dispatch(dy.contiguous(), x2d, w_t, lora_a, lora_b, s_eff, h, z, tiles, None,
                                               scaling, b_prime, num_slices, row_div)
        return dx, None, None, None, dA, dB, dse, None, None, None, None


def svlora_linear(x2d: Tensor, w: Tensor, w_t: Tensor, bias: Optional[Tensor], lora_a: Tensor, lora_b: Tensor,
                  s_eff: Tensor, scaling: float, b_prime: int, num_slices: int, row_div: int = 1) -> Tensor:
    return _SVLoRALinear.apply(x2d, w, w_t, bias, lora_a, lora_b, s_eff, scaling, b_prime, num_slices, row_div)


@torch.library.custom_op("ffm::frozen_linear", mutates_args=())
def frozen_linear_op(x2d: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tensor:
    """y [T, N] = x2d [T, K] · w [N, K]^T + bias (bf16 in / out, fp32 bias): ffm_frozen_linear."""
    _need_cuda(x2d, w, bias)
    T, K = x2d.shape
    N = w.shape[0]
    y = torch.empty((T, N), device=x2d.device, dtype=torch.bfloat16)
    _cabi.call("ffm_frozen_linear", _ptr(x2d), _ptr(w), _ptr(bias), _ptr(y), T, K, N, _stream())
    return y


@frozen_linear_op.register_fake
def _(x2d, w, bias):
    return x2d.new_empty((x2d.shape[0], w.shape[0]))


class _FrozenLinear(torch.autograd.Function):
    """A frozen projection on the library's own GEMM: forward x W^T + b, backward dx = dy W (the same kernel on W^T)."""

    @staticmethod
    def forward(ctx, x2d, w, w_t, bias):
        ctx.save_for_backward(w_t)
        return frozen_linear_op(x2d, w, bias)

    @staticmethod
    def backward(ctx, dy):
        (w_t,) = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        return frozen_linear_op(dy, w_t, None), None, None, None


def frozen_linear(x: Tensor, w: Tensor, w_t: Tensor, bias: Optional[Tensor]) -> Tensor:
    """x [..., K] bf16, w [N, K] / w_t [K, N] bf16 copies of a frozen weight, bias fp32 or None -> [..., N] bf16."""
    lead = x.shape[:-1]
    y = _FrozenLinear.apply(x.reshape(-1, x.shape[-1]).contiguous(), w, w_t, bias)
    return y.reshape(*lead, -1)


# adapter-gradient kernels of the fused MLP leave the dX critical path (see _SVLoRAMLP.backward); FFM_PARAMS_SIDE=0: A/B
PARAMS_ON_SIDE_STREAM = os.environ.get("FFM_PARAMS_SIDE", "1") != "0"
_PARAM_STREAMS: dict = {}


def _param_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index
    if key not in _PARAM_STREAMS:
        _PARAM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _PARAM_STREAMS[key]


def _bwd_phase(phases, dy, x, w_t, a, b, s_eff, h, z, tiles, dact, dx, gA, gB, dse, ws, scaling, b_prime, num_slices,
               row_div):
    T, N = dy.shape
    K = x.shape[1]
    _cabi.call("ffm_svlora_bwd_phase", _ptr(dy), _ptr(x), _ptr(w_t), _ptr(a), _ptr(b), _ptr(s_eff), _ptr(h), _ptr(z),
               _ptr(tiles), _ptr(dact), _ptr(dx), _ptr(gA), _ptr(gB), _ptr(dse), _ptr(ws), ws.numel(), T, K, N,
               a.shape[1], s_eff.shape[0], b_prime, num_slices, int(row_div), float(scaling), int(phases), _stream())


def _split_layer_backward(main, side, dy, x, w_t, a, b, s_eff, h, z, tiles, dact, ds_ctx, scaling, b_prime, num_slices,
                          row_div):
    """dX GEMM on the current stream, adapter gradients + dS on the parameter-gradient stream.  Returns dx."""
    T, N = dy.shape
    K = x.shape[1]
    nS, r = s_eff.shape
    lib = _cabi.load()
    ws = torch.empty((lib.ffm_svlora_bwd_workspace_bytes(T, K, N, nS),), device=x.device, dtype=torch.uint8)
    dx = torch.empty((T, K), device=x.device, dtype=torch.bfloat16)
    dse = torch.empty((nS, r), device=x.device, dtype=torch.float32)
    gA, gB = _direct_grad(a), _direct_grad(b)
    args = (dy, x, w_t, a, b, s_eff, h, z, tiles, dact, dx, gA, gB, dse, ws, scaling, b_prime, num_slices, row_div)
    _bwd_phase(1, *args)
    ev = torch.cuda.Event()
    ev.record(main)
    attr, lam, G, gS, gSg = ds_ctx
    with torch.cuda.stream(side):
        side.wait_event(ev)
        _bwd_phase(2, *args)
        _cabi.call("ffm_ds", _ptr(attr), _ptr(dse), _ptr(gS), _ptr(gSg), nS, G, r, lam, _stream())
    for t in (dy, x, h, z, ws, dse, s_eff):          # allocated on the main stream, still read on the side stream
        t.record_stream(side)
    return dx


class _SVLoRAMLP(torch.autograd.Function):
    """c_proj(QuickGELU(c_fc(x))) with both adapters (clip/model.py:325-332): QuickGELU is fused into the c_fc
    epilogue (dual store: the activation and its derivative) and the multiplication by that derivative into the
    c_proj backward epilogue."""

    @staticmethod
    def forward(ctx, x2d, w1, w1_t, b1, a1, bb1, s1, w2, w2_t, b2, a2, bb2, s2, scaling, b_prime, num_slices,
                row_div, p1=None, p2=None, dsc1=None, dsc2=None):
        ctx.ds_ctx = (dsc1, dsc2)
        g, u, h1, z1, t1 = svlora_fwd(x2d, w1, b1, a1, bb1, s1, scaling, b_prime, num_slices, 1, row_div, p1)
        y, _, h2, z2, t2 = svlora_fwd(g, w2, b2, a2, bb2, s2, scaling, b_prime, num_slices, 0, row_div, p2)
        if p1 is not None:
            t1 = p1
        if p2 is not None:
            t2 = p2
        ctx.save_for_backward(x2d, g, u, h1, h2, w1_t, a1, bb1, s1, w2_t, a2, bb2, s2, z1, t1, z2, t2)
        ctx.cfg = (scaling, b_prime, num_slices, row_div)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2d, g, u, h1, h2, w1_t, a1, bb1, s1, w2_t, a2, bb2, s2, z1, t1, z2, t2 = ctx.saved_tensors
        scaling, b_prime, num_slices, row_div = ctx.cfg
        dsc1, dsc2 = ctx.ds_ctx
        if PARAMS_ON_SIDE_STREAM and dsc1 is not None and dsc2 is not None and \
                all(_direct_grad(t) is not None for t in (a1, bb1, a2, bb2)):
            # Only dx feeds the blocks below: the adapter-gradient contractions, their fold and dS (6 launches, 54 us
            # per block) run on a side stream behind the dX GEMMs and overlap the attention backward that follows.
            main = torch.cuda.current_stream()
            side = _param_stream(dy.device)
            dy = dy.contiguous()
            du = _split_layer_backward(main, side, dy, g, w2_t, a2, bb2, s2, h2, z2, t2, u, dsc2, scaling, b_prime,
                                       num_slices, row_div)
            dx = _split_layer_backward(main, side, du, x2d, w1_t, a1, bb1, s1, h1, z1, t1, None, dsc1, scaling, b_prime,
                                       num_slices, row_div)
            ev = torch.cuda.Event()
            ev.record(side)
            _PENDING_DIRECT_WRITES.append(ev)
            return (dx,) + (None,) * 20
        du, dA2, dB2, ds2 = _svlora_bwd_dispatch(dy.contiguous(), g, w2_t, a2, bb2, s2, h2, z2, t2, u, scaling,
                                                 b_prime, num_slices, row_div)
        dx, dA1, dB1, ds1 = _svlora_bwd_dispatch(du, x2d, w1_t, a1, bb1, s1, h1, z1, t1, None, scaling, b_prime,
                                                 num_slices, row_div)
        return (dx, None, None, None, dA1, dB1, ds1, None, None, None, dA2, dB2, ds2, None, None, None, None, None, None,
                None, None)


def svlora_mlp(x2d, fc, proj, scaling: float, b_prime: int, num_slices: int, row_div: int = 1,
               prepared: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """fc / proj = (w, w_t, bias, lora_a, lora_b, s_eff) tuples; prepared = (svlora_prepare of fc, of proj) or None."""
    p1, p2 = prepared if prepared is not None else (None, None)
    dsc1, dsc2 = getattr(fc[5], "_ffm_ds_ctx", None), getattr(proj[5], "_ffm_ds_ctx", None)
    return _SVLoRAMLP.apply(x2d, *fc, *proj, scaling, b_prime, num_slices, row_div, p1, p2, dsc1, dsc2)


# =====================================================================================================
# average pooling of the ResNet trunk on channels-last activations
# =====================================================================================================
@torch.library.custom_op("ffm::avgpool_nhwc_fwd", mutates_args=())
def avgpool_nhwc_fwd_op(x: Tensor, k: int) -> Tensor:
    _need_cuda(x)
    b, c, h, w = x.shape
    y = torch.empty((b, c, h // k, w // k), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    _cabi.call("ffm_avgpool_nhwc_fwd", _ptr(x), _ptr(y), b, h, w, c, int(k), int(x.dtype == torch.bfloat16), _stream())
    return y


@avgpool_nhwc_fwd_op.register_fake
def _(x, k):
    b, c, h, w = x.shape
    return torch.empty((b, c, h // k, w // k), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)


@torch.library.custom_op("ffm::avgpool_nhwc_bwd", mutates_args=())
def avgpool_nhwc_bwd_op(dy: Tensor, k: int) -> Tensor:
    _need_cuda(dy)
    b, c, ho, wo = dy.shape
    dx = torch.empty((b, c, ho * k, wo * k), device=dy.device, dtype=dy.dtype, memory_format=torch.channels_last)
    _cabi.call("ffm_avgpool_nhwc_bwd", _ptr(dy), _ptr(dx), b, ho * k, wo * k, c, int(k), int(dy.dtype == torch.bfloat16),
               _stream())
    return dx


@avgpool_nhwc_bwd_op.register_fake
def _(dy, k):
    b, c, ho, wo = dy.shape
    return torch.empty((b, c, ho * k, wo * k), device=dy.device, dtype=dy.dtype, memory_format=torch.channels_last)


class _AvgPoolNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        ctx.k = k
        return avgpool_nhwc_fwd_op(x, k)

    @staticmethod
    def backward(ctx, dy):
        return avgpool_nhwc_bwd_op(dy.contiguous(memory_format=torch.channels_last), ctx.k), None


def avgpool_nhwc_supported(x: Tensor, k: int) -> bool:
    per = 8 if x.dtype == torch.bfloat16 else 4
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] % per == 0 and
            x.shape[2] % k == 0 and x.shape[3] % k == 0 and x.is_contiguous(memory_format=torch.channels_last))


def avgpool_nhwc(x: Tensor, k: int) -> Tensor:
    """nn.AvgPool2d(k) for channels-last fp32 activations (clip/model.py:30, :42, :108) with its own backward."""
    return _AvgPoolNHWC.apply(x, int(k))


# =====================================================================================================
# training-mode BatchNorm (+ ReLU) of the ResNet trunk on channels-last fp32 activations
# =====================================================================================================
@torch.library.custom_op("ffm::bn_relu_fwd", mutates_args=("running_mean", "running_var"))
def bn_relu_fwd_op(x: Tensor, gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor, momentum: float,
                   eps: float, relu: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(x, gamma, beta, running_mean, running_var)
    b, c, h, w = x.shape
    y = torch.empty_like(x, memory_format=torch.channels_last)
    mean = torch.empty((c,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((c,), device=x.device, dtype=torch.float32)
    nbytes = int(_cabi.load().ffm_bn_ws_bytes(c))
    ws = torch.empty((nbytes // 4,), device=x.device, dtype=torch.float32)
    _cabi.call("ffm_bn_relu_fwd", _ptr(x), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(y),
               _ptr(mean), _ptr(rstd), _ptr(ws), nbytes, b * h * w, c, float(momentum), float(eps), int(bool(relu)), _stream())
    return y, mean, rstd


@bn_relu_fwd_op.register_fake
def _(x, gamma, beta, running_mean, running_var, momentum, eps, relu):
    c = x.shape[1]
    return (torch.empty_like(x, memory_format=torch.channels_last), x.new_empty((c,)), x.new_empty((c,)))


@torch.library.custom_op("ffm::bn_relu_bwd", mutates_args=())
def bn_relu_bwd_op(x: Tensor, dy: Tensor, gamma: Tensor, beta: Tensor, mean: Tensor, rstd: Tensor,
                   relu: bool) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(x, dy, gamma, beta, mean, rstd)
    b, c, h, w = x.shape
    dx = torch.empty_like(x, memory_format=torch.channels_last)
    dgamma = torch.empty((c,), device=x.device, dtype=torch.float32)
    dbeta = torch.empty((c,), device=x.device, dtype=torch.float32)
    nbytes = int(_cabi.load().ffm_bn_ws_bytes(c))
    ws = torch.empty((nbytes // 4,), device=x.device, dtype=torch.float32)
    _cabi.call("ffm_bn_relu_bwd", _ptr(x), _ptr(dy), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dgamma),
               _ptr(dbeta), _ptr(ws), nbytes, b * h * w, c, int(bool(relu)), _stream())
    return dx, dgamma, dbeta


@bn_relu_bwd_op.register_fake
def _(x, dy, gamma, beta, mean, rstd, relu):
    c = x.shape[1]
    return torch.empty_like(x, memory_format=torch.channels_last), x.new_empty((c,)), x.new_empty((c,))


class _BatchNormReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu):
        y, mean, rstd = bn_relu_fwd_op(x, gamma, beta, running_mean, running_var, momentum, eps, relu)
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.relu = relu
        ctx.mark_non_differentiable(mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous(memory_format=torch.channels_last)
        if dy.dtype != torch.float32:
            dy = dy.float()
        dx, dgamma, dbeta = bn_relu_bwd_op(x, dy, gamma.detach(), beta.detach(), mean, rstd, ctx.relu)
        return dx, dgamma, dbeta, None, None, None, None, None


def batchnorm_relu_supported(x: Tensor) -> bool:
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and x.shape[1] % 4 == 0 and
            x.is_contiguous(memory_format=torch.channels_last))


def batchnorm_relu(x: Tensor, gamma: Tensor, beta: Tensor, running_mean: Tensor, running_var: Tensor, momentum: float,
                   eps: float, relu: bool) -> Tensor:
    """Training-mode nn.BatchNorm2d followed (relu=True) by ReLU on channels-last fp32 activations (clip/model.py:18-58): batch
    statistics, running-statistics update as torch does, gradients for x / gamma / beta; the ReLU mask is recomputed in the
    backward, so only x is kept."""
    return _BatchNormReLU.apply(x, gamma.float().contiguous(), beta.float().contiguous(), running_mean, running_var,
                                float(momentum), float(eps), bool(relu))


@torch.library.custom_op("ffm::add_relu", mutates_args=())
def add_relu_op(a: Tensor, b: Tensor) -> Tensor:
    _need_cuda(a, b)
    y = torch.empty_like(a)
    _cabi.call("ffm_add_relu", _ptr(a), _ptr(b), _ptr(y), a.numel(), _stream())
    return y


@add_relu_op.register_fake
def _(a, b):
    return torch.empty_like(a)


@torch.library.custom_op("ffm::relu_mask", mutates_args=())
def relu_mask_op(dy: Tensor, y: Tensor) -> Tensor:
    _need_cuda(dy, y)
    g = torch.empty_like(y)
    _cabi.call("ffm_relu_mask", _ptr(dy), _ptr(y), _ptr(g), y.numel(), _stream())
    return g


@relu_mask_op.register_fake
def _(dy, y):
    return torch.empty_like(y)


class _AddReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        y = add_relu_op(a, b)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        if dy.dtype != torch.float32 or dy.stride() != y.stride():
            dy = dy.float().contiguous(memory_format=torch.channels_last) if y.dim() == 4 and \
                y.is_contiguous(memory_format=torch.channels_last) else dy.float().contiguous()
        g = relu_mask_op(dy, y)
        return g, g


def add_relu(a: Tensor, b: Tensor) -> Tensor:
    """relu(a + b) for two fp32 CUDA tensors of identical shape and memory layout (clip/model.py:56-58); torch otherwise."""
    if (a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape and
            a.stride() == b.stride() and a.numel() >= 4 and a.numel() % 4 == 0 and
            (a.is_contiguous() or (a.dim() == 4 and a.is_contiguous(memory_format=torch.channels_last)))):
        return _AddReLU.apply(a, b)
    return torch.relu(a + b)


@torch.library.custom_op("ffm::widen_bf16", mutates_args=())
def widen_bf16_op(x: Tensor) -> Tensor:
    _need_cuda(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    _cabi.call("ffm_widen_bf16", _ptr(x), _ptr(out), x.numel(), _stream())
    return out


@widen_bf16_op.register_fake
def _(x):
    return torch.empty(x.shape, device=x.device, dtype=torch.float32)


class _WidenBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return widen_bf16_op(x)

    @staticmethod
    def backward(ctx, dy):
        return dy.to(torch.bfloat16)


def widen_bf16(x: Tensor) -> Tensor:
    """bf16 -> fp32 of a contiguous CUDA tensor (vectorised; differentiable); anything else falls back to Tensor.to."""
    if x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() >= 8 and x.numel() % 8 == 0:
        return _WidenBf16.apply(x)
    return x.to(torch.float32)


# =====================================================================================================
# merged weight of a plain LoRA projection (RN50 attention pool)
# =====================================================================================================
@torch.library.custom_op("ffm::lora_merged_weight", mutates_args=())
def lora_merged_weight_op(w: Tensor, a: Tensor, b: Tensor, scaling: float) -> Tensor:
    _need_cuda(w, a, b)
    out_f, in_f = w.shape
    out = torch.empty_like(w)
    _cabi.call("ffm_lora_merged_weight", _ptr(w), _ptr(a), _ptr(b), _ptr(out), out_f, in_f, int(a.shape[1]),
               float(scaling), _stream())
    return out


@lora_merged_weight_op.register_fake
def _(w, a, b, scaling):
    return torch.empty_like(w)


@torch.library.custom_op("ffm::lora_merged_weight_bwd", mutates_args=())
def lora_merged_weight_bwd_op(d_wm: Tensor, a: Tensor, b: Tensor, scaling: float) -> Tuple[Tensor, Tensor]:
    _need_cuda(d_wm, a, b)
    out_f, in_f = d_wm.shape
    nbytes = int(_cabi.load().ffm_lora_merged_weight_ws_bytes(out_f, in_f))
    ws = torch.empty((nbytes // 4,), device=d_wm.device, dtype=torch.float32)
    d_a, d_b = torch.empty_like(a), torch.empty_like(b)
    _cabi.call("ffm_lora_merged_weight_bwd", _ptr(d_wm), _ptr(a), _ptr(b), _ptr(d_a), _ptr(d_b), _ptr(ws), nbytes, out_f,
               in_f, int(a.shape[1]), float(scaling), _stream())
    return d_a, d_b


@lora_merged_weight_bwd_op.register_fake
def _(d_wm, a, b, scaling):
    return torch.empty_like(a), torch.empty_like(b)


class _LoraMergedWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, a, b, scaling):
        ctx.save_for_backward(a, b)
        ctx.scaling = scaling
        return lora_merged_weight_op(w, a, b, scaling)

    @staticmethod
    def backward(ctx, d_wm):
        a, b = ctx.saved_tensors
        d_a, d_b = lora_merged_weight_bwd_op(d_wm.float().contiguous(), a, b, ctx.scaling)
        return None, d_a, d_b, None


def lora_merged_weight(w: Tensor, a: Tensor, b: Tensor, scaling: float) -> Tensor:
    """W + scaling * (A @ B)^T in one pass, with its own deterministic backward for A and B (the frozen W gets no gradient).
    w f32 [out, in] frozen, a = lora_A.weight f32 [in, r], b = lora_B.weight f32 [r, out]
    (LoRALinear.weight, trainers/GLP_OT_SVLoRA.py:235-239)."""
    if w.requires_grad:
        raise _cabi.FfmError("lora_merged_weight: the base weight must be frozen")
    if w.dtype != torch.float32 or a.dtype != torch.float32 or b.dtype != torch.float32:
        raise _cabi.FfmError("lora_merged_weight: fp32 operands expected")
    return _LoraMergedWeight.apply(w.detach().contiguous(), a.contiguous(), b.contiguous(), float(scaling))


# =====================================================================================================
# residual add + LayerNorm (frozen glue between the adapted MLP and attention)
# =====================================================================================================
@torch.library.custom_op("ffm::add_layernorm_fwd", mutates_args=())
def add_layernorm_fwd(x: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor,
                      eps: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """s, ln, mean, rstd: s = bf16(x + res) (empty when res is None: s is x itself), ln = LayerNorm(s)."""
    _need_cuda(x, res, gamma, beta)
    C = x.shape[-1]
    rows = x.numel() // C
    s = torch.empty_like(x) if res is not None else x.new_empty((0,))
    ln = torch.empty_like(x)
    mean = torch.empty((rows,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((rows,), device=x.device, dtype=torch.float32)
    _cabi.call("ffm_add_layernorm_fwd", _ptr(x), _ptr(res), _ptr(gamma), _ptr(beta), _ptr(s) if res is not None else 0,
               _ptr(ln), _ptr(mean), _ptr(rstd), rows, C, float(eps), _stream())
    return s, ln, mean, rstd


@add_layernorm_fwd.register_fake
def _(x, res, gamma, beta, eps):
    rows = x.numel() // x.shape[-1]
    return (torch.empty_like(x) if res is not None else x.new_empty((0,)), torch.empty_like(x),
            x.new_empty((rows,), dtype=torch.float32), x.new_empty((rows,), dtype=torch.float32))


@torch.library.custom_op("ffm::add_layernorm_bwd", mutates_args=())
def add_layernorm_bwd(d_ln: Tensor, d_res: Optional[Tensor], s: Tensor, gamma: Tensor, mean: Tensor,
                      rstd: Tensor) -> Tensor:
    _need_cuda(d_ln, d_res, s)
    C = s.shape[-1]
    dx = torch.empty_like(s)
    _cabi.call("ffm_add_layernorm_bwd", _ptr(d_ln), _ptr(d_res), _ptr(s), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dx),
               s.numel() // C, C, _stream())
    return dx


@add_layernorm_bwd.register_fake
def _(d_ln, d_res, s, gamma, mean, rstd):
    return torch.empty_like(s)


class _AddLayerNorm(torch.autograd.Function):
    """(x, res) -> (x + res, LayerNorm(x + res)); the LayerNorm parameters are frozen (no gradient)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        s, ln, mean, rstd = add_layernorm_fwd(x, res, gamma, beta, eps)
        if res is None:
            s = x
        ctx.save_for_backward(s, gamma, mean, rstd)
        ctx.has_res = res is not None
        return s, ln

    @staticmethod
    def backward(ctx, d_s, d_ln):
        s, gamma, mean, rstd = ctx.saved_tensors
        if d_ln is None:
            dx = d_s
        else:
            dx = add_layernorm_bwd(d_ln.contiguous(), None if d_s is None else d_s.contiguous(), s, gamma, mean, rstd)
        return dx, (dx if ctx.has_res else None), None, None, None


def add_layernorm(x: Tensor, res: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float = 1e-5):
    """Residual update + LayerNorm in one pass: returns (x + res, LayerNorm(x + res)); res=None: (x, LayerNorm(x)).
    bf16 activations [..., C], fp32 frozen gamma/beta (clip/model.py:304-310, :354-357)."""
    if gamma.requires_grad or beta.requires_grad:
        raise _cabi.FfmError("add_layernorm: LayerNorm parameters must be frozen (FairLoRA trains adapters and prompts only)")
    x = x.contiguous()
    if res is not None:
        res = res.contiguous()
    return _AddLayerNorm.apply(x, res, gamma, beta, eps)


# =====================================================================================================
# frozen attention core (packed q/k/v in, packed dq/dk/dv out)
# =====================================================================================================
ATTENTION_MAX_LEN = 208
ATTENTION_HEAD_DIM = 64


@torch.library.custom_op("ffm::attention_fwd", mutates_args=())
def attention_fwd(qkv: Tensor, n_head: int, causal: bool, batch_first: bool) -> Tuple[Tensor, Tensor]:
    """out, lse.  qkv bf16 [B, L, 3*C] (batch_first) or [L, B, 3*C] as produced by the packed in_proj; out has the same
    leading dims and C = n_head * 64 columns; lse f32 [B*n_head, L] (base 2)."""
    _need_cuda(qkv)
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.dim() != 3:
        raise _cabi.FfmError("attention_fwd: qkv must be contiguous bf16 [B, L, 3C] / [L, B, 3C]")
    d0, d1, c3 = qkv.shape
    B, L = (d0, d1) if batch_first else (d1, d0)
    C = c3 // 3
    if C != n_head * ATTENTION_HEAD_DIM or L > ATTENTION_MAX_LEN:
        raise _cabi.FfmError(f"attention_fwd: head dim must be {ATTENTION_HEAD_DIM} and L <= {ATTENTION_MAX_LEN}")
    out = torch.empty((d0, d1, C), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((B * n_head, L), device=qkv.device, dtype=torch.float32)
    _cabi.call("ffm_attention_fwd", _ptr(qkv), _ptr(out), _ptr(lse), B, L, n_head, ATTENTION_HEAD_DIM, int(bool(causal)),
               int(bool(batch_first)), _stream())
    return out, lse


@attention_fwd.register_fake
def _(qkv, n_head, causal, batch_first):
    d0, d1, c3 = qkv.shape
    B, L = (d0, d1) if batch_first else (d1, d0)
    return qkv.new_empty((d0, d1, c3 // 3)), qkv.new_empty((B * n_head, L), dtype=torch.float32)


@torch.library.custom_op("ffm::attention_bwd", mutates_args=())
def attention_bwd(qkv: Tensor, out: Tensor, d_out: Tensor, lse: Tensor, n_head: int, causal: bool,
                  batch_first: bool) -> Tensor:
    _need_cuda(qkv, out, d_out, lse)
    d0, d1, c3 = qkv.shape
    B, L = (d0, d1) if batch_first else (d1, d0)
    d_qkv = torch.empty_like(qkv)
    _cabi.call("ffm_attention_bwd", _ptr(qkv), _ptr(out), _ptr(d_out), _ptr(lse), _ptr(d_qkv), B, L, n_head,
               ATTENTION_HEAD_DIM, int(bool(causal)), int(bool(batch_first)), _stream())
    return d_qkv


@attention_bwd.register_fake
def _(qkv, out, d_out, lse, n_head, causal, batch_first):
    return torch.empty_like(qkv)


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, n_head, causal, batch_first):
        out, lse = attention_fwd(qkv, n_head, causal, batch_first)
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (n_head, causal, batch_first)
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        n_head, causal, batch_first = ctx.cfg
        d_out = d_out.contiguous()
        if d_out.dtype != torch.bfloat16:
            d_out = d_out.to(torch.bfloat16)
        return attention_bwd(qkv, out, d_out, lse, n_head, causal, batch_first), None, None, None


def attention(qkv: Tensor, n_head: int, causal: bool = False, batch_first: bool = True) -> Tensor:
    """Self-attention core on the packed in_proj output (clip/model.py:350-352): returns [.., C] ready for out_proj."""
    return _Attention.apply(qkv.contiguous(), n_head, bool(causal), bool(batch_first))


def attention_supported(width: int, n_head: int, seq_len: int) -> bool:
    return width == n_head * ATTENTION_HEAD_DIM and 1 <= seq_len <= ATTENTION_MAX_LEN


# =====================================================================================================
# ViT input side: /255 + mean/std + im2col in one pass, class token + positions + ln_pre + first ln_1 in another
# =====================================================================================================
@torch.library.custom_op("ffm::patchify_normalize", mutates_args=())
def patchify_normalize(image: Tensor, mean: Tensor, std: Tensor, patch: int, div255: bool) -> Tensor:
    """image f32 [B', C, H, W] -> bf16 [B', G, C*patch*patch] = im2col of bf16(((image / 255) - mean) / std)
    (trainers/GLP_OT_SVLoRA.py:679-693 + clip/model.py:431-433)."""
    _need_cuda(image, mean, std)
    if image.dtype != torch.float32 or not image.is_contiguous():
        raise _cabi.FfmError("patchify_normalize: image must be contiguous fp32 [B', C, H, W]")
    bp, c, h, w = image.shape
    out = torch.empty((bp, (h // patch) * (w // patch), c * patch * patch), device=image.device, dtype=torch.bfloat16)
    _cabi.call("ffm_patchify_normalize", _ptr(image), _ptr(out), _ptr(mean), _ptr(std), bp, c, h, w, int(patch),
               int(bool(div255)), _stream())
    return out


@patchify_normalize.register_fake
def _(image, mean, std, patch, div255):
    bp, c, h, w = image.shape
    return image.new_empty((bp, (h // patch) * (w // patch), c * patch * patch), dtype=torch.bfloat16)


@torch.library.custom_op("ffm::oct_slice_conv_fwd", mutates_args=())
def oct_slice_conv_fwd_op(x: Tensor, w: Tensor, bias: Tensor, in_scale: float) -> Tensor:
    _need_cuda(x, w, bias)
    bp, cin, h, wd = x.shape
    cout = w.shape[0]
    y = torch.empty((bp, cout, h, wd), device=x.device, dtype=torch.float32)
    _cabi.call("ffm_oct_slice_conv_fwd", _ptr(x), _ptr(w), _ptr(bias), _ptr(y), bp, cin, cout, h, wd, float(in_scale),
               _stream())
    return y


@oct_slice_conv_fwd_op.register_fake
def _(x, w, bias, in_scale):
    return x.new_empty((x.shape[0], w.shape[0], x.shape[2], x.shape[3]))


@torch.library.custom_op("ffm::oct_slice_conv_wgrad", mutates_args=())
def oct_slice_conv_wgrad_op(x: Tensor, dy: Tensor, cout: int, in_scale: float) -> Tuple[Tensor, Tensor]:
    _need_cuda(x, dy)
    bp, cin, h, wd = x.shape
    dw = torch.empty((cout, cin, 5, 5), device=x.device, dtype=torch.float32)
    db = torch.empty((cout,), device=x.device, dtype=torch.float32)
    nbytes = int(_cabi.load().ffm_oct_slice_conv_wgrad_ws_bytes(cin))
    ws = torch.empty((nbytes // 4,), device=x.device, dtype=torch.float32)
    _cabi.call("ffm_oct_slice_conv_wgrad", _ptr(x), _ptr(dy), _ptr(dw), _ptr(db), _ptr(ws), nbytes, bp, cin, cout, h, wd,
               float(in_scale), _stream())
    return dw, db


@oct_slice_conv_wgrad_op.register_fake
def _(x, dy, cout, in_scale):
    return x.new_empty((cout, x.shape[1], 5, 5)), x.new_empty((cout,))


class _OctSliceConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, in_scale):
        ctx.save_for_backward(x)
        ctx.cfg = (w.shape[0], in_scale)
        return oct_slice_conv_fwd_op(x, w, bias, in_scale)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        cout, in_scale = ctx.cfg
        dw, db = oct_slice_conv_wgrad_op(x, dy.float().contiguous(), cout, in_scale)
        return None, dw, db, None


def oct_slice_conv_supported(x: Tensor, w: Tensor, padding: int) -> bool:
    return (x.is_cuda and x.dim() == 4 and w.dim() == 4 and tuple(w.shape[2:]) == (5, 5) and padding == 2 and
            w.shape[0] <= 4 and w.shape[1] <= 32 and x.shape[3] % 4 == 0 and x.shape[0] <= 65535 and not x.requires_grad)


def oct_slice_conv(x: Tensor, w: Tensor, bias: Tensor, in_scale: float = 1.0 / 255.0) -> Tensor:
    """proj_per_3d_slice(image * in_scale) for the raw slices x f32 [B', Cin, H, W] (trainers/GLP_OT_SVLoRA.py:684): own
    forward and weight-gradient kernels (the input is data and gets no gradient)."""
    return _OctSliceConv.apply(x.float().contiguous(), w.float().contiguous(), bias.float().contiguous(), float(in_scale))


@torch.library.custom_op("ffm::oct_minmax_patchify", mutates_args=())
def oct_minmax_patchify_op(y: Tensor, mean: Tensor, std: Tensor, patch: int) -> Tuple[Tensor, Tensor, Tensor]:
    """patches, lo, hi: per-slice min-max scaling + mean/std + bf16 im2col of y f32 [B', C, H, W] (ffm_oct_minmax_patchify)."""
    _need_cuda(y, mean, std)
    if y.dtype != torch.float32 or not y.is_contiguous():
        raise _cabi.FfmError("oct_minmax_patchify: y must be contiguous fp32 [B', C, H, W]")
    bp, c, h, w = y.shape
    out = torch.empty((bp, (h // patch) * (w // patch), c * patch * patch), device=y.device, dtype=torch.bfloat16)
    lo = torch.empty((bp,), device=y.device, dtype=torch.float32)
    hi = torch.empty((bp,), device=y.device, dtype=torch.float32)
    _cabi.call("ffm_oct_minmax_patchify", _ptr(y), _ptr(lo), _ptr(hi), _ptr(out), _ptr(mean), _ptr(std), bp, c, h, w,
               int(patch), _stream())
    return out, lo, hi


@oct_minmax_patchify_op.register_fake
def _(y, mean, std, patch):
    bp, c, h, w = y.shape
    return (y.new_empty((bp, (h // patch) * (w // patch), c * patch * patch), dtype=torch.bfloat16), y.new_empty((bp,)),
            y.new_empty((bp,)))


@torch.library.custom_op("ffm::oct_input_bwd", mutates_args=())
def oct_input_bwd_op(d_patches: Tensor, y: Tensor, lo: Tensor, hi: Tensor, std: Tensor, patch: int) -> Tensor:
    bp, c, h, w = y.shape
    d_y = torch.empty_like(y)
    nbytes = int(_cabi.load().ffm_oct_input_bwd_ws_bytes(bp))
    ws = torch.empty((nbytes // 4,), device=y.device, dtype=torch.float32)
    _cabi.call("ffm_oct_input_bwd", _ptr(d_patches), _ptr(y), _ptr(lo), _ptr(hi), _ptr(std), _ptr(d_y), _ptr(ws), nbytes,
               bp, c, h, w, int(patch), _stream())
    return d_y


@oct_input_bwd_op.register_fake
def _(d_patches, y, lo, hi, std, patch):
    return torch.empty_like(y)


class _OctMinMaxPatchify(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, mean, std, patch):
        patches, lo, hi = oct_minmax_patchify_op(y, mean, std, patch)
        ctx.save_for_backward(y, lo, hi, std)
        ctx.patch = patch
        return patches

    @staticmethod
    def backward(ctx, d_patches):
        y, lo, hi, std = ctx.saved_tensors
        d_patches = d_patches.contiguous()
        if d_patches.dtype != torch.bfloat16:
            d_patches = d_patches.to(torch.bfloat16)
        return oct_input_bwd_op(d_patches, y, lo, hi, std, ctx.patch), None, None, None


def oct_minmax_patchify(y: Tensor, mean: Tensor, std: Tensor, patch: int) -> Tensor:
    """OCT input side after the slice projection (trainers/GLP_OT_SVLoRA.py:686-693): per-slice min-max scaling, CLIP
    mean/std, bf16 cast and im2col in two passes over y, with the matching fused backward (differentiable in y)."""
    return _OctMinMaxPatchify.apply(y.contiguous(), mean.float().contiguous(), std.float().contiguous(), int(patch))


@torch.library.custom_op("ffm::vit_embed_ln", mutates_args=())
def vit_embed_ln(patch_emb: Tensor, cls: Tensor, pos: Tensor, g_pre: Tensor, b_pre: Tensor, g_1: Tensor, b_1: Tensor,
                 eps_pre: float, eps_1: float) -> Tuple[Tensor, Tensor]:
    """x0 = LN_pre(cat(cls, patch_emb) + pos), h0 = LN_1(x0): bf16 [B', G+1, C] each (clip/model.py:434-440, :354).
    patch_emb bf16 [B', G, C]; the tables and LayerNorm parameters are the frozen fp32 masters."""
    _need_cuda(patch_emb, cls, pos, g_pre, b_pre, g_1, b_1)
    if patch_emb.dtype != torch.bfloat16 or not patch_emb.is_contiguous():
        raise _cabi.FfmError("vit_embed_ln: patch_emb must be contiguous bf16 [B', G, C]")
    bp, G, C = patch_emb.shape
    x0 = torch.empty((bp, G + 1, C), device=patch_emb.device, dtype=torch.bfloat16)
    h0 = torch.empty_like(x0)
    _cabi.call("ffm_vit_embed_ln", _ptr(patch_emb), _ptr(cls), _ptr(pos), _ptr(g_pre), _ptr(b_pre), _ptr(g_1),
               _ptr(b_1), _ptr(x0), _ptr(h0), 0, 0, bp, G, C, float(eps_pre), float(eps_1), _stream())
    return x0, h0


@vit_embed_ln.register_fake
def _(patch_emb, cls, pos, g_pre, b_pre, g_1, b_1, eps_pre, eps_1):
    bp, G, C = patch_emb.shape
    return patch_emb.new_empty((bp, G + 1, C)), patch_emb.new_empty((bp, G + 1, C))


@torch.library.custom_op("ffm::ot_head_fwd", mutates_args=())
def ot_head_fwd(img: Tensor, txt: Tensor, logit_scale: Tensor, num_slices: int, mode: int, eps: float, thresh: float,
                max_iter: int, top_percent: float,
                batch_first: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """logits, T, sim, inv_norm, status, workspace.  img [M+1, Bp, D] (or [Bp, M+1, D] with batch_first; bf16/f32),
    txt [N, n_cls, D] f32."""
    _need_cuda(img, txt, logit_scale)
    if batch_first:
        Bp, Mp1, D = img.shape
    else:
        Mp1, Bp, D = img.shape
    M = Mp1 - 1
    N, n_cls, _ = txt.shape
    P = Bp * n_cls
    dev = img.device
    logits = torch.empty((Bp // num_slices, n_cls), device=dev, dtype=torch.float32)
    T = torch.empty((P, M, N), device=dev, dtype=torch.float32) if mode else logits.new_empty((0,))
    sim = torch.empty((P, M, N), device=dev, dtype=torch.float32)
    inv_norm = torch.empty((M, Bp), device=dev, dtype=torch.float32)
    status = torch.empty((2,), device=dev, dtype=torch.int32)
    lib = _cabi.load()
    ws_bytes = lib.ffm_ot_head_workspace_bytes(M, Bp, D, N, n_cls)
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    _cabi.call("ffm_ot_head_fwd", _ptr(img), int(img.dtype == torch.bfloat16), int(bool(batch_first)), _ptr(txt),
               _ptr(logit_scale),
               _ptr(logits), _ptr(T) if mode else 0, _ptr(sim), _ptr(inv_norm), _ptr(status), _ptr(ws), ws_bytes, M,
               Bp, D, N, n_cls, num_slices, mode, float(eps), float(thresh), int(max_iter), float(top_percent),
               _stream())
    return logits, T, sim, inv_norm, status, ws


@ot_head_fwd.register_fake
def _(img, txt, logit_scale, num_slices, mode, eps, thresh, max_iter, top_percent, batch_first=False):
    Mp1, Bp, D = (img.shape[1], img.shape[0], img.shape[2]) if batch_first else img.shape
    N, n_cls, _ = txt.shape
    f = dict(device=img.device, dtype=torch.float32)
    P = Bp * n_cls
    return (torch.empty((Bp // num_slices, n_cls), **f), torch.empty((P, Mp1 - 1, N) if mode else (0,), **f),
            torch.empty((P, Mp1 - 1, N), **f), torch.empty((Mp1 - 1, Bp), **f),
            torch.empty((2,), device=img.device, dtype=torch.int32),
            torch.empty((1,), device=img.device, dtype=torch.uint8))


@torch.library.custom_op("ffm::ot_head_bwd", mutates_args=())
def ot_head_bwd(img: Tensor, txt: Tensor, logit_scale: Tensor, d_logits: Tensor, T: Tensor, sim: Tensor,
                inv_norm: Tensor, ws: Tensor, num_slices: int, mode: int,
                batch_first: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    if batch_first:
        Bp, Mp1, D = img.shape
    else:
        Mp1, Bp, D = img.shape
    N, n_cls, _ = txt.shape
    d_img = torch.empty_like(img)
    d_txt = torch.empty_like(txt)
    d_ls = torch.empty((), device=img.device, dtype=torch.float32)
    _cabi.call("ffm_ot_head_bwd", _ptr(img), int(img.dtype == torch.bfloat16), int(bool(batch_first)), _ptr(txt),
               _ptr(logit_scale),
               _ptr(d_logits), _ptr(T) if mode else 0, _ptr(sim), _ptr(inv_norm), _ptr(d_img), _ptr(d_txt), _ptr(d_ls),
               _ptr(ws), ws.numel(), Mp1 - 1, Bp, D, N, n_cls, num_slices, mode, _stream())
    return d_img, d_txt, d_ls


@ot_head_bwd.register_fake
def _(img, txt, logit_scale, d_logits, T, sim, inv_norm, ws, num_slices, mode, batch_first=False):
    return torch.empty_like(img), torch.empty_like(txt), logit_scale.new_empty(())


class _OTHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, logit_scale, num_slices, mode, eps, thresh, max_iter, top_percent, batch_first=False):
        logits, T, sim, inv_norm, status, ws = ot_head_fwd(img, txt, logit_scale, num_slices, mode, eps, thresh,
                                                           max_iter, top_percent, batch_first)
        ctx.save_for_backward(img, txt, logit_scale, T, sim, inv_norm, ws)
        ctx.cfg = (num_slices, mode, batch_first)
        ctx.mark_non_differentiable(status, T)
        return logits, status, T

    @staticmethod
    def backward(ctx, d_logits, _d_status, _d_T):
        img, txt, logit_scale, T, sim, inv_norm, ws = ctx.saved_tensors
        num_slices, mode, batch_first = ctx.cfg
        d_img, d_txt, d_ls = ot_head_bwd(img, txt, logit_scale, d_logits.contiguous().float(), T, sim, inv_norm, ws,
                                         num_slices, mode, batch_first)
        return d_img, d_txt, d_ls.reshape(logit_scale.shape), None, None, None, None, None, None, None


def ot_head(image_features: Tensor, text_features: Tensor, logit_scale: Tensor, *, n_cls: int, num_slices: int = 1,
            ot: str = "Sinkhorn", eps: float = 0.1, thresh: float = 1e-3, max_iter: int = 100,
            top_percent: float = 0.8, batch_first: bool = False):
    """Head of CustomCLIP.forward (trainers/GLP_OT_SVLoRA.py:696-757).

    image_features [M+1, Bp, D] like the reference (or [Bp, M+1, D] with batch_first=True; bf16 or f32),
    text_features [N*n_cls, D] prompt-major. Returns
    (logits [Bp/num_slices, n_cls] f32, status int32[2] = {iterations, nan flag}, T)."""
    img = image_features.contiguous()
    if img.dtype not in (torch.bfloat16, torch.float32):
        img = img.float()
    D = img.shape[-1]
    txt = text_features.float().contiguous().view(-1, n_cls, D)
    ls = logit_scale.float().reshape(1) if logit_scale.dim() == 0 else logit_scale.float()
    return _OTHead.apply(img, txt, ls, num_slices, OT_MODES[ot], eps, thresh, max_iter, top_percent, bool(batch_first))


def sinkhorn(K: Tensor, *, mode: str = "Sinkhorn", v_mass: float = 1.0, thresh: float = 1e-3, max_iter: int = 100):
    """Stand-alone persistent Sinkhorn / COT on K [P, M, N] f32 with u = 1/M, v = v_mass/N. Returns (T, status)."""
    _need_cuda(K)
    K = K.float().contiguous()
    P, M, N = K.shape
    T = torch.empty_like(K)
    status = torch.empty((2,), device=K.device, dtype=torch.int32)
    lib = _cabi.load()
    ws_bytes = lib.ffm_sinkhorn_workspace_bytes(P, M, N)
    ws = torch.empty((ws_bytes,), device=K.device, dtype=torch.uint8)
    _cabi.call("ffm_sinkhorn", _ptr(K), _ptr(T), _ptr(status), _ptr(ws), ws_bytes, P, M, N, OT_MODES[mode],
               float(v_mass), float(thresh), int(max_iter), _stream())
    return T, status


# =====================================================================================================
# flat-buffer utilities (aggregation / SGD) and metrics
# =====================================================================================================
def fedavg_scale(flat_in: Tensor, seg_kind: Tensor, seg_off: Tensor, seg_len: Tensor, w_scalar: float,
                 w_group: Tensor, G: int, r: int, out: Optional[Tensor] = None) -> Tensor:
    _need_cuda(flat_in, seg_kind, seg_off, seg_len, w_group)
    out = torch.empty_like(flat_in) if out is None else out
    _cabi.call("ffm_fedavg_scale", _ptr(flat_in), _ptr(out), _ptr(seg_kind), _ptr(seg_off), _ptr(seg_len),
               seg_kind.numel(), flat_in.numel(), float(w_scalar), _ptr(w_group), G, r, _stream())
    return out


def fedavg_epilogue(avg: Tensor, prev_global: Tensor, seg_kind: Tensor, seg_off: Tensor, seg_len: Tensor,
                    beta_decay: float, shared_half_s: bool, G: int, r: int) -> Tensor:
    _need_cuda(avg, prev_global)
    out = torch.empty_like(avg)
    _cabi.call("ffm_fedavg_epilogue", _ptr(avg), _ptr(prev_global), _ptr(out), _ptr(seg_kind), _ptr(seg_off),
               _ptr(seg_len), seg_kind.numel(), avg.numel(), float(beta_decay), int(bool(shared_half_s)), G, r,
               _stream())
    return out


def sgd_step_(param: Tensor, grad: Tensor, momentum_buf: Tensor, lr: float, momentum: float, weight_decay: float,
              n_steps: int, first_step: bool) -> None:
    _need_cuda(param, grad, momentum_buf)
    _cabi.call("ffm_sgd_step", _ptr(param), _ptr(grad), _ptr(momentum_buf), param.numel(), float(lr), float(momentum),
               float(weight_decay), int(n_steps), int(bool(first_step)), _stream())


def sgd_step_dev_lr_(param: Tensor, grad: Tensor, momentum_buf: Tensor, lr_dev: Tensor, momentum: float,
                     weight_decay: float, n_steps: int) -> None:
    """sgd_step_ with the learning rate in device memory (graph-replayable; momentum buffer zero-initialised)."""
    _need_cuda(param, grad, momentum_buf, lr_dev)
    _cabi.call("ffm_sgd_step_dev_lr", _ptr(param), _ptr(grad), _ptr(momentum_buf), param.numel(), _ptr(lr_dev),
               float(momentum), float(weight_decay), int(n_steps), _stream())


def group_auc_counts(prob: Tensor, label: Tensor, attrs: Optional[Tensor], max_groups: int) -> Tensor:
    """uint64-as-int64 counts [n_slots, 8] = {gt0, eq0, gt1, eq1, tp, fp, tn, fn}; see include/ffm_b200.h."""
    _need_cuda(prob, label, attrs)
    prob = prob.float().contiguous()
    label = label.to(torch.int32).contiguous()
    N = prob.shape[0]
    n_attr = 0 if attrs is None else attrs.shape[0]
    attrs_i = None if attrs is None else attrs.to(torch.int32).contiguous()
    n_slots = 1 + n_attr * (max_groups + 1)
    counts = torch.empty((n_slots, 8), device=prob.device, dtype=torch.int64)
    lib = _cabi.load()
    ws_bytes = lib.ffm_group_auc_workspace_bytes(N, n_attr, max_groups)
    ws = torch.empty((ws_bytes,), device=prob.device, dtype=torch.uint8)
    _cabi.call("ffm_group_auc", _ptr(prob), _ptr(label), _ptr(attrs_i), _ptr(counts), _ptr(ws), ws_bytes, N, n_attr,
               max_groups, _stream())
    return counts
