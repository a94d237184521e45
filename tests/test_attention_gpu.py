"""GPU parity of the attention core (scope row f1, csrc/attention.cu) against an fp32 restatement of what
nn.MultiheadAttention computes inside ResidualAttentionBlock (clip/model.py:350-352): softmax(q k^T / sqrt(d)) v per
head on the packed in_proj output, optional causal mask (text tower), forward and backward, both memory layouts."""
from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(qkv32, n_head, causal, batch_first):
    """The oracle's attention core (oracle/ref_port.attention_core, pinned to torch's nn.MultiheadAttention by
    tests/test_oracle_golden.py) on fp32 qkv [B, L, 3C] / [L, B, 3C] -> same leading dims, C columns."""
    from oracle import ref_port as rp
    x = qkv32.transpose(0, 1) if batch_first else qkv32                       # the oracle speaks sequence-first
    L = x.shape[0]
    mask = torch.full((L, L), float("-inf"), device=x.device).triu_(1) if causal else None
    o = rp.attention_core(x, n_head, mask)
    return o.transpose(0, 1) if batch_first else o


@pytest.mark.parametrize("B,L,H,causal,batch_first", [
    (2, 197, 12, False, True),      # image tower, config shapes
    (4, 77, 8, True, True),         # text tower: causal
    (3, 50, 2, False, False),       # reference layout [L, B, 3C] (RN50-style token count)
    (2, 208, 1, False, True),       # the longest supported sequence
    (2, 16, 2, True, False),        # exactly one tile, causal, sequence-first
    (1, 5, 3, False, True),         # shorter than a tile
    (2, 100, 4, True, True),        # several key blocks under the causal mask
])
def test_attention_matches_fp32_reference(B, L, H, causal, batch_first):
    from fairfedmed_b200 import ops
    g = torch.Generator().manual_seed(L * 31 + H)
    C = H * 64
    shape = (B, L, 3 * C) if batch_first else (L, B, 3 * C)
    qkv = (1.5 * torch.randn(shape, generator=g)).to(DEV).to(torch.bfloat16)
    d_out = torch.randn(shape[0], shape[1], C, generator=g).to(DEV).to(torch.bfloat16)
    q_g = qkv.clone().requires_grad_(True)
    out = ops.attention(q_g, H, causal, batch_first)
    out.backward(d_out)
    q_r = qkv.float().requires_grad_(True)
    ref = _reference(q_r, H, causal, batch_first)
    ref.backward(d_out.float())
    assert out.shape == ref.shape and out.dtype == torch.bfloat16
    # bf16 output rounding (2^-8 relative) + bf16 probabilities inside P·V
    assert float((out.float() - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
    gmax = float(q_r.grad.abs().max())
    err = (q_g.grad.float() - q_r.grad).abs()
    assert float(err.max()) <= 3e-2 * gmax, float(err.max()) / gmax
    # per part (dq, dk, dv): relative Frobenius error
    Cn = C
    for part in range(3):
        a = q_g.grad.float()[..., part * Cn:(part + 1) * Cn]
        b = q_r.grad[..., part * Cn:(part + 1) * Cn]
        assert float((a - b).norm() / b.norm()) <= 1.5e-2, part


def test_attention_is_deterministic_and_matches_library_path():
    """Two runs are bit-identical (no atomics), and the block's own-kernel path agrees with the torch SDPA path."""
    from fairfedmed_b200 import clip_model, ops
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(8, 197, 3 * 768, generator=g).to(DEV).to(torch.bfloat16).requires_grad_(True)
    d_out = torch.randn(8, 197, 768, generator=g).to(DEV).to(torch.bfloat16)
    grads = []
    for _ in range(2):
        qkv.grad = None
        o = ops.attention(qkv, 12, False, True)
        o.backward(d_out)
        grads.append((o.detach().clone(), qkv.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    blk = clip_model.ResidualAttentionBlock(768, 12, batch_first=True).to(DEV)
    x = torch.randn(4, 197, 768, generator=g).to(DEV).to(torch.bfloat16)
    outs = []
    for own in (True, False):
        clip_model.OWN_ATTENTION = own
        try:
            outs.append(blk.attention(x).float())
        finally:
            clip_model.OWN_ATTENTION = False
    assert float((outs[0] - outs[1]).abs().max()) <= 2e-2 * float(outs[1].abs().max()) + 1e-3
