"""CLIP ViT image encoder, text encoder, prompt learner and CustomCLIP with the reference's module interfaces.

Class names, attribute names, call signatures (`MLP.forward(x, attr)`, `ResidualAttentionBlock.forward(x, attr)`,
`ModifiedVisionTransformer.forward(x, attr)` returning ALL tokens sequence-first, `CustomCLIP.forward(image, attr)`)
and state-dict keys follow clip/model.py:304-449 and trainers/GLP_OT_SVLoRA.py:46-200,575-763 of the reference, so
checkpoints and the federated key conventions carry over.  What differs is how it runs:

  * the adapted MLP (c_fc -> QuickGELU -> c_proj) is two fused tcgen05 kernels forward and two backward
    (ops.svlora_mlp), QuickGELU and its derivative living in the GEMM epilogues;
  * the GLP_OT head (normalise, similarities, Sinkhorn / COT, logits) is the fused head op (ops.ot_head);
  * frozen weights are consumed from cached bf16 copies (fp32 masters stay in the state dict); activations are
    bf16, sequence-first [L, B', C] like the reference so the row -> sample mapping of the adapters holds.
Frozen attention / LayerNorm / patch embedding use PyTorch library kernels (bf16 cuBLAS, SDPA) — they are the
"next" row of the scope table, not the hot path built here.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .modules import FairLoRALinear, LoRALinear, SVLoRALinear, _AdapterBase, _attr_on
from .resnet_model import ModifiedResNet_GLP_OT

PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)
PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)
# attention core: "own" = csrc/attention.cu (scope row f1: tcgen05 / TMEM forward + backward, packed dq/dk/dv,
# deterministic), "lib" = torch SDPA (cuDNN) kept as an A/B switch.
OWN_ATTENTION = os.environ.get("FFM_ATTENTION", "own") == "own"
# in_proj / out_proj of the image tower: "own" = ffm_frozen_linear (the fused GEMM built without the adapter side product),
# "lib" = cuBLAS through F.linear
OWN_PROJ = os.environ.get("FFM_PROJ", "own") == "own"
_SIDE_STREAMS: dict = {}     # (device index, role) -> side stream (module level: models stay picklable)


def _side_stream(device, role: str):
    key = (torch.device(device).index, role)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class _Bf16Cache:
    """bf16 compute copies of frozen fp32 parameters, refreshed when the master changes (load_state_dict, .to())."""

    def _bf(self, name: str, tensor: Optional[torch.Tensor], dtype=torch.bfloat16):
        if tensor is None:
            return None
        store = self.__dict__.setdefault("_bf_store", {})
        key = (tensor.data_ptr(), tensor._version, tensor.device, dtype)
        hit = store.get(name)
        if hit is None or hit[0] != key:
            store[name] = (key, tensor.detach().to(dtype).contiguous())
            hit = store[name]
        return hit[1]


    def _bf_t(self, name: str, tensor: torch.Tensor):
        """bf16 copy of the TRANSPOSE of a frozen 2-D weight (the dX operand of ops.frozen_linear)."""
        store = self.__dict__.setdefault("_bf_store", {})
        key = (tensor.data_ptr(), tensor._version, tensor.device, "t")
        hit = store.get(name)
        if hit is None or hit[0] != key:
            store[name] = (key, tensor.detach().t().to(torch.bfloat16).contiguous())
            hit = store[name]
        return hit[1]


class LayerNorm(nn.LayerNorm, _Bf16Cache):
    """fp32 statistics on bf16 activations (clip/model.py:304-310 computes in fp32 and casts back)."""

    def forward(self, x: torch.Tensor):
        if x.dtype == torch.float32:
            return super().forward(x)
        return F.layer_norm(x, self.normalized_shape, self._bf("w", self.weight, x.dtype),
                            self._bf("b", self.bias, x.dtype), self.eps)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class MLP(nn.Module, _Bf16Cache):
    def __init__(self, d_model: int, batch_first: bool = False):
        super().__init__()
        self.c_fc = nn.Linear(d_model, d_model * 4)
        self.gelu = QuickGELU()
        self.c_proj = nn.Linear(d_model * 4, d_model)
        self.batch_first = batch_first    # rows of x are [B', L, C] instead of the reference's [L, B', C]

    def _adapter_operands(self, layer: _AdapterBase, attr, device):
        w, w_t, bias = layer._operands()
        if isinstance(layer, FairLoRALinear):
            s_eff = layer._s_eff(attr, device)
        elif isinstance(layer, SVLoRALinear):
            s_eff = layer.lora_S.weight.reshape(1, -1)
            if layer.global_s:
                s_eff = s_eff + layer.lora_S_global.weight.reshape(1, -1)
            s_eff = s_eff.contiguous()
        else:
            s_eff = torch.ones((1, layer.rank), device=device, dtype=torch.float32)
        return (w, w_t, bias, layer.lora_A.weight, layer.lora_B.weight, s_eff)

    def _fusable(self) -> bool:
        return isinstance(self.c_fc, _AdapterBase) and isinstance(self.c_proj, _AdapterBase) \
            and self.c_fc.scaling == self.c_proj.scaling

    def prepare(self, attr, device):
        """Everything of the fused MLP that depends on parameters and `attr` only (s_eff of both adapters, their bf16
        tiles): the caller may run this ahead of time on another stream and hand the result to forward(prepared=)."""
        fc = self._adapter_operands(self.c_fc, attr, device)
        pj = self._adapter_operands(self.c_proj, attr, device)
        t1 = ops.svlora_prepare(fc[3].detach(), fc[4].detach(), fc[5].detach(), self.c_fc.scaling)
        t2 = ops.svlora_prepare(pj[3].detach(), pj[4].detach(), pj[5].detach(), self.c_proj.scaling)
        return fc, pj, t1, t2

    def forward(self, x: torch.Tensor, attr: Optional[torch.Tensor] = None, prepared=None):
        fused = self._fusable() and x.is_cuda and x.dim() == 3
        if fused:
            if self.batch_first:
                bp, L, c = x.shape
                row_div = L          # row t = column * L + position: sample = (t // L) // num_slices
            else:
                L, bp, c = x.shape
                row_div = 1          # row t = position * B' + column (reference layout)
            if prepared is not None:
                fc, pj, t1, t2 = prepared
                tiles = (t1, t2)
            else:
                fc = self._adapter_operands(self.c_fc, attr, x.device)
                pj = self._adapter_operands(self.c_proj, attr, x.device)
                tiles = None
            n_samples = fc[5].shape[0]
            x2d = x.reshape(L * bp, c)
            if x2d.dtype != torch.bfloat16:
                x2d = x2d.to(torch.bfloat16)
            y = ops.svlora_mlp(x2d.contiguous(), fc, pj, self.c_fc.scaling, bp, bp // n_samples, row_div, tiles)
            return y.reshape(x.shape[0], x.shape[1], -1).to(x.dtype)
        if isinstance(self.c_fc, _AdapterBase):
            if self.batch_first:     # the stand-alone adapter modules speak the reference's sequence-first layout
                return self.c_proj(self.gelu(self.c_fc(x.transpose(0, 1), attr)), attr).transpose(0, 1)
            return self.c_proj(self.gelu(self.c_fc(x, attr)), attr)
        # un-adapted (text tower): plain frozen linears on bf16 copies
        h = F.linear(x, self._bf("fc_w", self.c_fc.weight, x.dtype), self._bf("fc_b", self.c_fc.bias, x.dtype))
        h = self.gelu(h)
        return F.linear(h, self._bf("pj_w", self.c_proj.weight, x.dtype), self._bf("pj_b", self.c_proj.bias, x.dtype))


class ResidualAttentionBlock(nn.Module, _Bf16Cache):
    def __init__(self, d_model: int, n_head: int, attn_mask: Optional[torch.Tensor] = None,
                 batch_first: bool = False):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)   # parameter container (same state-dict keys)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = MLP(d_model, batch_first=batch_first)
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.n_head = n_head
        self.batch_first = batch_first
        # opt-in FairLoRA adapters on in_proj / out_proj (modules.apply_lora_to_model(adapt_attention=True)); None = the
        # reference's configuration (MLP adapters only)
        self.attn_in_lora = None
        self.attn_out_lora = None

    def _proj(self, adapter, name, x, weight, bias, attr):
        if adapter is not None:            # raises on CPU tensors: no silent un-adapted fallback
            return adapter(x, attr, batch_first=self.batch_first)
        if OWN_PROJ and x.is_cuda and x.dtype == torch.bfloat16 and self.attn_mask is None \
                and weight.shape[0] % 8 == 0 and weight.shape[1] % 8 == 0 and not weight.requires_grad:
            # image tower (no causal mask): frozen projections on the package's own tcgen05 GEMM
            w = self._bf(name + "_w", weight, torch.bfloat16)
            w_t = self._bf_t(name + "_wt", weight)
            return ops.frozen_linear(x, w, w_t, None if bias is None else self._bf(name + "_bf", bias, torch.float32))
        return F.linear(x, self._bf(name + "_w", weight, x.dtype), self._bf(name + "_b", bias, x.dtype))

    def attention(self, x: torch.Tensor, attr: Optional[torch.Tensor] = None):
        """Self-attention with nn.MultiheadAttention semantics (clip/model.py:350-352).

        batch_first: x is [B, L, C]; q/k/v are strided views of the in_proj output in "bshd" memory order, which the
        fused attention kernels consume without copies.  Otherwise x is the reference's [L, B, C]."""
        a = self.attn
        causal = self.attn_mask is not None
        qkv = self._proj(self.attn_in_lora, "in", x, a.in_proj_weight, a.in_proj_bias, attr)
        seq_len = x.shape[1] if self.batch_first else x.shape[0]
        if OWN_ATTENTION and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 3 \
                and ops.attention_supported(x.shape[-1], self.n_head, seq_len):
            # own kernels: q/k/v read from the packed projection in place, dq/dk/dv written packed (no cat in backward)
            out = ops.attention(qkv, self.n_head, causal, self.batch_first)
            return self._proj(self.attn_out_lora, "out", out, a.out_proj.weight, a.out_proj.bias, attr)
        if self.batch_first:
            bn, L, c = x.shape
            hd = c // self.n_head
            q, k, v = qkv.view(bn, L, 3, self.n_head, hd).unbind(2)                       # each [B, L, H, hd]
            out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                                 is_causal=causal)                        # [B, H, L, hd]
            out = out.transpose(1, 2).reshape(bn, L, c)
        else:
            L, bn, c = x.shape
            hd = c // self.n_head
            qkv = qkv.view(L, bn, 3, self.n_head, hd).permute(2, 1, 3, 0, 4)              # [3, B, H, L, hd]
            out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], is_causal=causal)
            out = out.permute(2, 0, 1, 3).reshape(L, bn, c)
        return self._proj(self.attn_out_lora, "out", out, a.out_proj.weight, a.out_proj.bias, attr)

    def forward(self, x: torch.Tensor, attr: Optional[torch.Tensor] = None):
        x = x + self.attention(self.ln_1(x), attr)
        x = x + self.mlp(self.ln_2(x), attr=attr)
        return x


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: Optional[torch.Tensor] = None,
                 text_layer=False, design_details=None, batch_first: bool = False):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, attn_mask, batch_first)
                                        for _ in range(layers)])

    hoist_adapter_prep = True     # issue every block's s_eff / adapter-tile kernels up front on a side stream

    def _prepare_adapters(self, attr, device):
        """The 4 tiny launches per block that only depend on parameters and `attr` (group mixing of the singular values
        and the bf16 adapter tiles, for c_fc and c_proj) leave the critical path: they are issued for all blocks at
        once on a side stream, each block waits for its own event.  Autograd runs the matching backward nodes (dS) on
        that stream too."""
        if not any(blk.mlp._fusable() for blk in self.resblocks):
            return None
        cur = torch.cuda.current_stream(device)
        side = _side_stream(device, "prep")
        side.wait_stream(cur)
        out = []
        with torch.cuda.stream(side):
            for blk in self.resblocks:
                if not blk.mlp._fusable():
                    out.append(None)
                    continue
                pre = blk.mlp.prepare(attr, device)
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((pre, ev))
        return out

    def _fusable(self, x: torch.Tensor) -> bool:
        if not (x.is_cuda and x.dtype == torch.bfloat16 and self.width % 256 == 0 and self.width <= 1024):
            return False
        # the fused kernel has no LayerNorm-parameter gradients: FairLoRA freezes them (trainers/GLP_OT_SVLoRA.py:822-829)
        return not any(p.requires_grad for blk in self.resblocks for ln in (blk.ln_1, blk.ln_2) for p in ln.parameters())

    def forward(self, x: torch.Tensor, attr=None, final_ln: Optional[nn.LayerNorm] = None,
                h0: Optional[torch.Tensor] = None):
        """Same computation as the chain of ResidualAttentionBlock.forward (clip/model.py:370-374), optionally followed
        by `final_ln` (ln_post / ln_final of the caller).  On the device every `x = x + branch; h = LayerNorm(x)` pair
        runs as ONE kernel (ops.add_layernorm): the residual stream is read and written once per half-block.
        `h0`: ln_1 of the first block already applied to x by the caller (ops.vit_embed_ln)."""
        if not self._fusable(x) or (final_ln is not None and any(p.requires_grad for p in final_ln.parameters())):
            assert h0 is None
            for block in self.resblocks:
                x = block(x, attr)
            return x if final_ln is None else final_ln(x)
        blocks = self.resblocks
        prepared = self._prepare_adapters(attr, x.device) if self.hoist_adapter_prep else None
        ln = blocks[0].ln_1
        if h0 is None:
            x, h = ops.add_layernorm(x, None, ln.weight, ln.bias, ln.eps)
        else:
            h = h0
        for i, blk in enumerate(blocks):
            a = blk.attention(h, attr)
            x, h = ops.add_layernorm(x, a, blk.ln_2.weight, blk.ln_2.bias, blk.ln_2.eps)
            if prepared is not None and prepared[i] is not None:
                pre, ev = prepared[i]
                cur = torch.cuda.current_stream(x.device)
                cur.wait_event(ev)
                for t in (pre[0][5], pre[1][5], pre[2], pre[3]):     # s_eff x2, tiles x2: allocated on the side stream
                    t.record_stream(cur)
                m = blk.mlp(h, attr=attr, prepared=pre)
            else:
                m = blk.mlp(h, attr=attr)
            nxt = blocks[i + 1].ln_1 if i + 1 < len(blocks) else final_ln
            if nxt is None:
                return x + m
            x, h = ops.add_layernorm(x, m, nxt.weight, nxt.bias, nxt.eps)
        return h


class ModifiedVisionTransformer(nn.Module, _Bf16Cache):
    """Returns all 197 tokens, sequence-first [L, B', output_dim] (clip/model.py:413-449)."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int,
                 design_details=None):
        super().__init__()
        self.input_resolution, self.output_dim, self.patch_size = input_resolution, output_dim, patch_size
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        # internal activations are batch-first [B', L, C] (no transposes around attention); the adapters are told
        # through row_div, and the result is handed back sequence-first like the reference
        self.transformer = Transformer(width, layers, heads, design_details=design_details, batch_first=True)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor, attr: Optional[torch.Tensor] = None):
        dt = x.dtype
        # stride = kernel = patch: the convolution is a GEMM over non-overlapping patches (clip/model.py:431-433)
        bp, ch, hh, ww = x.shape
        ps = self.patch_size
        gh, gw = hh // ps, ww // ps
        patches = x.view(bp, ch, gh, ps, gw, ps).permute(0, 2, 4, 1, 3, 5).reshape(bp, gh * gw, ch * ps * ps)
        return self.forward_patches(patches, attr)

    def _frozen_input_side(self) -> bool:
        tr = self.transformer
        params = [self.conv1.weight, self.class_embedding, self.positional_embedding, *self.ln_pre.parameters(),
                  *self.ln_post.parameters(), *tr.resblocks[0].ln_1.parameters()]
        return not any(p.requires_grad for p in params)

    def forward_patches(self, patches: torch.Tensor, attr: Optional[torch.Tensor] = None, batch_first: bool = False):
        """`forward` from the im2col'ed, normalised image on (patches [B', G, 3*P*P], e.g. ops.patchify_normalize).
        batch_first=True returns the tokens as they are stored, [B', L, output_dim] (ops.ot_head(batch_first=True)
        reads that layout directly); the default is the reference's sequence-first view."""
        dt = patches.dtype
        x = patches @ self._bf("conv1", self.conv1.weight, dt).flatten(1).t()                  # [B', g*g, width]
        tr = self.transformer
        if x.is_cuda and dt == torch.bfloat16 and not x.requires_grad and tr._fusable(x) and self._frozen_input_side():
            # class token + positions + ln_pre + the first block's ln_1 in one pass over the tokens (forward only:
            # nothing on this side of the residual stream is trainable)
            ln1 = tr.resblocks[0].ln_1
            x0, h0 = ops.vit_embed_ln(x.contiguous(), self.class_embedding.detach(), self.positional_embedding.detach(),
                                      self.ln_pre.weight.detach(), self.ln_pre.bias.detach(), ln1.weight.detach(),
                                      ln1.bias.detach(), self.ln_pre.eps, ln1.eps)
            x = tr(x0, attr=attr, final_ln=self.ln_post, h0=h0)
            x = x @ self._bf("proj", self.proj, dt)
            return x if batch_first else x.transpose(0, 1)
        cls = self._bf("cls", self.class_embedding, dt).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self._bf("pos", self.positional_embedding, dt)
        x = self.ln_pre(x)
        x = self.transformer(x, attr=attr, final_ln=self.ln_post)
        x = x @ self._bf("proj", self.proj, dt)                                                # [B', L, output_dim]
        return x if batch_first else x.transpose(0, 1)                                         # [L, B', output_dim]


class TextEncoder(nn.Module, _Bf16Cache):
    """trainers/GLP_OT_SVLoRA.py:46-66."""

    def __init__(self, width=512, layers=12, heads=8, context_length=77, embed_dim=512):
        super().__init__()
        mask = torch.full((context_length, context_length), float("-inf")).triu_(1)
        self.transformer = Transformer(width, layers, heads, attn_mask=mask, batch_first=True)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, width))
        self.ln_final = LayerNorm(width)
        self.text_projection = nn.Parameter(torch.empty(width, embed_dim))
        self.compute_dtype = torch.bfloat16

    def forward(self, prompts: torch.Tensor, eot_index: torch.Tensor):
        dt = self.compute_dtype if prompts.is_cuda else prompts.dtype
        x = prompts.to(dt) + self._bf("pos", self.positional_embedding, dt)
        x = self.transformer(x, final_ln=self.ln_final)   # prompts are already [n_prompts * n_cls, 77, width]
        x = x[torch.arange(x.shape[0], device=x.device), eot_index]
        return x.float() @ self.text_projection


class PromptLearner(nn.Module):
    """Generic learnable context, class token at the end (trainers/GLP_OT_SVLoRA.py:68-152).

    The tokenizer / token-embedding lookup that produces `token_prefix` / `token_suffix` and the EOT positions is
    host-side glue outside the hot path: pass them in (from the reference's PromptLearner — see INTEGRATION.md — or
    `synthetic_prompt_buffers` for offline runs with random-init weights)."""

    def __init__(self, n_prompts: int, n_ctx: int, ctx_dim: int, n_cls: int, token_prefix: torch.Tensor,
                 token_suffix: torch.Tensor, eot_index: torch.Tensor):
        super().__init__()
        ctx = torch.empty(n_prompts, n_ctx, ctx_dim)
        nn.init.normal_(ctx, std=0.02)
        self.ctx = nn.Parameter(ctx)
        self.register_buffer("token_prefix", token_prefix.clone())      # [N*n_cls, 1, D]
        self.register_buffer("token_suffix", token_suffix.clone())      # [N*n_cls, 77-1-n_ctx, D]
        self.register_buffer("eot_index", eot_index.clone().long(), persistent=False)
        self.N, self.n_cls, self.n_ctx = n_prompts, n_cls, n_ctx

    def forward(self):
        ctx = self.ctx.unsqueeze(0).expand(self.n_cls, -1, -1, -1).permute(1, 0, 2, 3)
        ctx = ctx.reshape(self.N * self.n_cls, self.n_ctx, -1)
        return torch.cat([self.token_prefix, ctx.to(self.token_prefix.dtype), self.token_suffix], dim=1)


def synthetic_prompt_buffers(n_prompts: int, n_cls: int, n_ctx: int, ctx_dim: int, context_length: int = 77,
                             name_lens: Sequence[int] = (3, 2), seed: int = 1):
    """Random stand-ins for the frozen token embeddings of "X X X X <classname>." (std 0.02 like
    CLIP.initialize_parameters) and the matching EOT positions: SOS + n_ctx + name tokens + '.' then EOT."""
    g = torch.Generator().manual_seed(seed)
    n = n_prompts * n_cls
    prefix = 0.02 * torch.randn(n_cls, 1, ctx_dim, generator=g)
    suffix = 0.02 * torch.randn(n_cls, context_length - 1 - n_ctx, ctx_dim, generator=g)
    eot = torch.tensor([1 + n_ctx + name_lens[c % len(name_lens)] + 1 for c in range(n_cls)])
    return prefix.repeat(n_prompts, 1, 1), suffix.repeat(n_prompts, 1, 1), eot.repeat(n_prompts)[:n]


class CustomCLIP(nn.Module):
    """FairLoRA model: image encoder (adapted), text encoder, prompt learner and the GLP_OT head (:575-763)."""

    def __init__(self, *, classnames: Sequence[str] = ("NOT Glaucoma", "Glaucoma"), n_prompts: int = 2, n_ctx: int = 4,
                 ot: str = "None", eps: float = 0.1, thresh: float = 1e-3, max_iter: int = 100,
                 top_percent: float = 0.8, image_resolution: int = 224, vision_layers=12,
                 vision_width: int = 768, vision_patch_size: int = 16, embed_dim: int = 512, text_width: int = 512,
                 text_layers: int = 12, text_heads: int = 8, context_length: int = 77,
                 dim_per_3d_slice: Optional[int] = None, prompt_buffers=None, dataset: str = "FairFedMed",
                 seed: int = 1):
        super().__init__()
        self.n_cls = len(classnames)
        self.N = n_prompts
        self.OT, self.eps, self.thresh, self.max_iter, self.top_percent = ot, eps, thresh, max_iter, top_percent
        self.dataset = dataset
        self.is_3d_input = dim_per_3d_slice is not None
        self.dim_per_3d_slice = dim_per_3d_slice
        if self.is_3d_input:
            self.proj_per_3d_slice = nn.Conv2d(dim_per_3d_slice, 3, kernel_size=5, padding=2)
            nn.init.normal_(self.proj_per_3d_slice.weight, std=dim_per_3d_slice ** -0.5)
            nn.init.zeros_(self.proj_per_3d_slice.bias)
        if prompt_buffers is None:
            prompt_buffers = synthetic_prompt_buffers(n_prompts, self.n_cls, n_ctx, text_width, context_length,
                                                      seed=seed)
        self.prompt_learner = PromptLearner(n_prompts, n_ctx, text_width, self.n_cls, *prompt_buffers)
        if isinstance(vision_layers, (tuple, list)):
            # CLIP ResNet (clip/model.py:484-492): heads = width * 32 / 64, attention pool returns all tokens
            self.image_encoder = ModifiedResNet_GLP_OT(tuple(vision_layers), embed_dim, vision_width * 32 // 64,
                                                       image_resolution, vision_width)
        else:
            self.image_encoder = ModifiedVisionTransformer(image_resolution, vision_patch_size, vision_width,
                                                           vision_layers, vision_width // 64, embed_dim)
        self.text_encoder = TextEncoder(text_width, text_layers, text_heads, context_length, embed_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.compute_dtype = torch.bfloat16
        self.register_buffer("pixel_mean", torch.tensor(PIXEL_MEAN).reshape(1, -1, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(PIXEL_STD).reshape(1, -1, 1, 1), persistent=False)
        self.last_status = None
        self.initialize_parameters()

    def initialize_parameters(self):
        """Random init in the spirit of CLIP.initialize_parameters (clip/model.py:533-560); no checkpoints offline."""
        te, ve = self.text_encoder, self.image_encoder
        nn.init.normal_(te.positional_embedding, std=0.01)
        towers = (te.transformer, ve.transformer) if hasattr(ve, "transformer") else (te.transformer,)
        if not hasattr(ve, "transformer"):           # ResNet: attention-pool init as in CLIP.initialize_parameters :536-547
            std = ve.attnpool.c_proj.in_features ** -0.5
            for lin in (ve.attnpool.q_proj, ve.attnpool.k_proj, ve.attnpool.v_proj, ve.attnpool.c_proj):
                nn.init.normal_(lin.weight, std=std)
            for layer in (ve.layer1, ve.layer2, ve.layer3, ve.layer4):
                for block in layer:
                    nn.init.zeros_(block.bn3.weight)
        for tower in towers:
            proj_std = (tower.width ** -0.5) * ((2 * tower.layers) ** -0.5)
            attn_std = tower.width ** -0.5
            fc_std = (2 * tower.width) ** -0.5
            for block in tower.resblocks:
                nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
                nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
                nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
                nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(te.text_projection, std=te.transformer.width ** -0.5)

    def preprocess(self, image: torch.Tensor):
        """/255, OCT slice projection + per-item min-max, mean/std normalisation (:679-693)."""
        b, c, h, w = image.shape
        image = image / 255.0
        if self.is_3d_input:
            image = image.reshape(-1, self.dim_per_3d_slice, h, w)
            image = self.proj_per_3d_slice(image)
            lo = image.amin(dim=(1, 2, 3), keepdim=True)
            hi = image.amax(dim=(1, 2, 3), keepdim=True)
            image = (image - lo) / (hi - lo + 1e-5)
        return (image - self.pixel_mean) / self.pixel_std

    def forward(self, image: torch.Tensor, attr: Optional[torch.Tensor] = None):
        b = image.shape[0]
        # The text tower (4 prompts x 77 tokens: ~300 launches of tiny kernels forward + backward) does not depend on
        # the image: fork it onto a side stream so it fills the gaps of the image tower instead of extending the
        # critical path.  Autograd replays the backward of each node on its forward stream, so the text backward
        # overlaps the image backward the same way; under CUDA-graph capture the fork/join become graph edges.
        fork = image.is_cuda and self.overlap_text
        if fork:
            cur = torch.cuda.current_stream(image.device)
            side = self._side_stream(image.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                prompts = self.prompt_learner()
                txt = self.text_encoder(prompts, self.prompt_learner.eot_index)
        dt = self.compute_dtype if image.is_cuda else torch.float32
        attr_dev = _attr_on(image.device, attr)
        ve = self.image_encoder
        fast_input = (image.is_cuda and not self.is_3d_input and isinstance(ve, ModifiedVisionTransformer)
                      and image.shape[1] == self.pixel_mean.shape[1] and ve.patch_size % 8 == 0
                      and image.shape[2] % ve.patch_size == 0 and image.shape[3] % ve.patch_size == 0
                      and image.shape[3] % 4 == 0)
        if isinstance(self.image_encoder, ModifiedResNet_GLP_OT):
            # conv trunk with training-mode BatchNorm over small batches: keep fp32 activations (bf16 only inside the
            # fused adapter kernels), batch statistics amplify bf16 rounding
            feats = self.image_encoder(self.preprocess(image.float()), attr=attr_dev)
        elif (image.is_cuda and self.is_3d_input and isinstance(ve, ModifiedVisionTransformer) and ve.patch_size % 8 == 0
              and image.shape[2] % ve.patch_size == 0 and image.shape[3] % ve.patch_size == 0 and image.shape[3] % 4 == 0):
            # OCT volumes (:681-693): the trainable slice projection with the /255 folded into its weights (a library
            # convolution), then min-max scaling + mean/std + bf16 cast + im2col as two fused passes (own backward)
            fast_input = True
            slices = image.float().reshape(-1, self.dim_per_3d_slice, image.shape[2], image.shape[3])
            pw, pb = self.proj_per_3d_slice.weight, self.proj_per_3d_slice.bias
            if ops.oct_slice_conv_supported(slices, pw, 2) and pb is not None:
                y = ops.oct_slice_conv(slices, pw, pb, 1.0 / 255.0)            # own forward + weight-gradient kernels
            else:
                y = F.conv2d(slices, pw / 255.0, pb, padding=2)
            patches = ops.oct_minmax_patchify(y, self.pixel_mean.reshape(-1), self.pixel_std.reshape(-1), ve.patch_size)
            feats = ve.forward_patches(patches, attr=attr_dev, batch_first=True)
        elif fast_input:
            # /255, mean/std, bf16 cast and im2col in one pass over the raw image
            ve = self.image_encoder
            patches = ops.patchify_normalize(image.float().contiguous(), self.pixel_mean.reshape(-1),
                                             self.pixel_std.reshape(-1), ve.patch_size, True)
            feats = ve.forward_patches(patches, attr=attr_dev, batch_first=True)       # [B', M+1, D] as stored
        else:
            x = self.preprocess(image.float())
            feats = self.image_encoder(x.to(dt), attr=attr_dev)                    # [M+1, B', D]
        if fork:
            cur.wait_stream(side)
            txt.record_stream(cur)
        else:
            prompts = self.prompt_learner()
            txt = self.text_encoder(prompts, self.prompt_learner.eot_index)       # [N*n_cls, D] fp32
        num_slices = feats.shape[0 if fast_input else 1] // b
        logits, status, _ = ops.ot_head(feats, txt, self.logit_scale, n_cls=self.n_cls, num_slices=num_slices,
                                        ot=self.OT, eps=self.eps, thresh=self.thresh, max_iter=self.max_iter,
                                        top_percent=self.top_percent, batch_first=fast_input)
        self.last_status = status
        if self.OT != "None" and self.check_nan and int(status[1].item()) != 0:
            return None                                                            # reference :738-743
        return logits

    check_nan = True
    overlap_text = True      # run the text tower on a side stream (CUDA only)

    @staticmethod
    def _side_stream(device):
        return _side_stream(device, "text")
