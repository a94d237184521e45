"""Drop-in adapter modules: FairLoRALinear / SVLoRALinear / LoRALinear and apply_lora_to_model.

Same constructor signatures, parameter names / shapes and call conventions as the reference
(trainers/GLP_OT_SVLoRA.py:203-573), so state dicts, the federated key conventions (`lora_S` substring,
`original_linear.*`) and `clip/model.py`'s `MLP.forward(x, attr)` call sites keep working — but forward and
backward run as ONE fused sm_100a kernel each (fairfedmed_b200/csrc/svlora_gemm.cu).

Numerics: activations and the frozen weight are consumed in bf16 with fp32 accumulation; the adapter
parameters stay fp32 masters (nn.Embedding weights, as upstream).  Inputs of another dtype are cast on the
way in and the result is cast back, so the modules also drop into an fp32 / fp16 reference model.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

LAMBDA_GROUP = 0.7  # trainers/GLP_OT_SVLoRA.py:459


def _attr_on(device, attr: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The reference hands every layer the same CPU int64 tensor (never moved, :988-994).  The device copy is cached
    under the tensor's CONTENT (a few dozen group ids), so all layers of a step share one copy and an in-place edit
    of the host tensor can never return stale ids."""
    if attr is None:
        return None
    if attr.device == device and attr.dtype == torch.int64:
        return attr
    if attr.is_cuda:
        return attr.to(device=device, dtype=torch.int64).contiguous()
    key = (str(device), tuple(attr.shape), tuple(attr.reshape(-1).tolist()))
    cache = getattr(_attr_on, "_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    dev_attr = attr.to(device=device, dtype=torch.int64, non_blocking=False).contiguous()
    _attr_on._cache = (key, dev_attr)
    return dev_attr


class FrozenLinearView:
    """A frozen projection that lives as bare parameters of another module (nn.MultiheadAttention's packed
    `in_proj_weight` / `in_proj_bias`, or its `out_proj`), presented with the nn.Linear attributes the adapters read.
    Deliberately NOT an nn.Module: wrapping it must not register the frozen tensors under a second state-dict key."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        self.weight, self.bias = weight, bias
        self.out_features, self.in_features = weight.shape

    def parameters(self):
        return [p for p in (self.weight, self.bias) if p is not None]


class _AdapterBase(nn.Module):
    """Shared plumbing: frozen base layer, bf16 operand cache, [L, B', C] / 1x1-conv layouts."""

    def _init_base(self, original_linear: nn.Module, rank: int, alpha: float):
        self.original_linear = original_linear
        self.rank = rank
        self.alpha = alpha
        self.scaling = self.alpha / self.rank
        w = original_linear.weight
        if w.dim() == 2:
            self.is_1x1_conv = False
            self.in_features, self.out_features = original_linear.in_features, original_linear.out_features
        else:
            if tuple(w.shape[-2:]) != (1, 1):
                raise ValueError("only 1x1 convolutions can be wrapped")
            self.is_1x1_conv = True
            self.out_features, self.in_features = w.shape[:2]
        ops.padded_rank(rank)          # raises for ranks the fused kernel was not built for (1..32)
        for p in self.original_linear.parameters():
            p.requires_grad = False
        self._cache_key = None
        self._w = self._w_t = self._bias = None

    def _new_embedding(self, n: int, d: int) -> nn.Embedding:
        w = self.original_linear.weight
        emb = nn.Embedding(n, d)
        emb.weight.data = emb.weight.data.to(dtype=torch.float32, device=w.device)
        return emb

    def _operands(self):
        """bf16 copies of the frozen weight ([N,K] and its transpose [K,N]) + fp32 bias, refreshed on change."""
        w = self.original_linear.weight
        b = self.original_linear.bias
        key = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        if key != self._cache_key:
            w2 = w.detach().reshape(self.out_features, self.in_features)
            self._w = w2.to(torch.bfloat16).contiguous()
            self._w_t = w2.t().to(torch.bfloat16).contiguous()
            self._bias = None if b is None else b.detach().float().contiguous()
            self._cache_key = key
        return self._w, self._w_t, self._bias

    def bias(self):
        return self.original_linear.bias

    def _tokens_in(self, x):
        """-> (x2d bf16 [T, C_in], restore(y2d) -> reference layout, B')."""
        if self.is_1x1_conv:
            b, c, h, w = x.shape
            tok = x.reshape(b, c, h * w).permute(2, 0, 1)            # [hw, b, c]  (:470-471)

            def restore(y2d):
                return y2d.reshape(h, w, b, -1).permute(2, 3, 0, 1)  # (:479-480)
            bp = b
        else:
            tok = x
            lead = x.shape[:-1]

            def restore(y2d):
                return y2d.reshape(*lead, -1)
            bp = x.shape[1] if x.dim() == 3 else 1
        x2d = tok.reshape(-1, tok.shape[-1])
        if x2d.dtype != torch.bfloat16:
            x2d = x2d.to(torch.bfloat16)
        return x2d.contiguous(), restore, bp

    def _run(self, x, s_eff, batch_first: bool = False):
        """batch_first: x is [B', L, C] (the layout inside this package's ViT tower) instead of the reference's
        sequence-first [L, B', C]; the kernel is told through row_div which sample a row belongs to."""
        if not x.is_cuda:
            raise RuntimeError("fairfedmed_b200 adapters need CUDA tensors: there is no CPU fallback "
                               "(the CPU oracle lives under oracle/ and is test infrastructure only)")
        w, w_t, bias = self._operands()
        row_div = 1
        if self.is_1x1_conv and x.dim() == 4:
            # a 1x1 convolution is a linear layer over the B*H*W pixels.  The reference walks them as [hw, b, c]
            # (:469-471, a permute copy each way); in channels-last memory the pixels of one image are already rows of
            # a [b, hw, c] matrix, so the kernel reads / writes the activations in place and is told through row_div
            # which sample a row belongs to: no layout copies around the 32 adapted convolutions of the ResNet trunk.
            b, c, h, wd = x.shape
            xc = x.contiguous(memory_format=torch.channels_last)
            x2d = xc.permute(0, 2, 3, 1).reshape(b * h * wd, c)
            x2d = (x2d if x2d.dtype == torch.bfloat16 else x2d.to(torch.bfloat16)).contiguous()
            bp, row_div = b, h * wd

            def restore(y2d):
                return y2d.view(b, h, wd, -1).permute(0, 3, 1, 2)      # logical NCHW, channels-last memory
        elif batch_first:
            if self.is_1x1_conv or x.dim() != 3:
                raise ValueError("batch_first adapters take [B', L, C] token tensors")
            bp, row_div = x.shape[0], x.shape[1]
            lead = x.shape[:-1]
            x2d = x.reshape(-1, x.shape[-1])
            x2d = (x2d if x2d.dtype == torch.bfloat16 else x2d.to(torch.bfloat16)).contiguous()

            def restore(y2d):
                return y2d.reshape(*lead, -1)
        else:
            x2d, restore, bp = self._tokens_in(x)
        n_samples = s_eff.shape[0]
        if bp % n_samples != 0:
            raise ValueError(f"batch columns {bp} not divisible by the number of attribute rows {n_samples}")
        num_slices = bp // n_samples                                   # OCT slices per sample (:473-475)
        y2d = ops.svlora_linear(x2d, w, w_t, bias, self.lora_A.weight, self.lora_B.weight, s_eff, self.scaling, bp,
                                num_slices, row_div)
        if y2d.dtype != x.dtype:           # on the contiguous [T, N] matrix, before the layout view
            y2d = ops.widen_bf16(y2d) if x.dtype == torch.float32 else y2d.to(x.dtype)
        return restore(y2d)


class FairLoRALinear(_AdapterBase):
    """Group-conditioned SVD-factored adapter (reference :333-500)."""

    def __init__(self, original_linear, rank=4, alpha=0.4, global_s=False, num_attrs=1):
        super().__init__()
        assert num_attrs > 0, "Number of attributes must be provided!"
        self._init_base(original_linear, rank, alpha)
        self.global_s = global_s
        self.num_attrs = num_attrs
        self.lora_A = self._new_embedding(self.in_features, rank)
        self.lora_S = self._new_embedding(num_attrs, rank)
        if self.global_s:
            self.lora_S_global = self._new_embedding(1, rank)
        self.lora_B = self._new_embedding(rank, self.out_features)
        self.reset_parameters()

    def reset_parameters(self, init_type="same+cycle"):
        """A = 0, B ~ N(0,1), S per the reference's 'same+cycle' / 'same' / 'cycle_shift' schemes (:380-423)."""
        nn.init.zeros_(self.lora_A.weight)
        r, G = self.rank, self.num_attrs
        dev = self.lora_S.weight.device
        if init_type in {"same", "cycle_shift"}:
            base = torch.linspace(1, 0.1, steps=r, device=dev)
            if init_type == "same":
                S = base.unsqueeze(0).repeat(G, 1)
            else:
                assert r >= G
                S = torch.stack([torch.roll(base, -i * (r // G)) for i in range(G)])
        else:
            assert r % 2 == 0 and r >= G
            half = r // 2
            base = torch.linspace(0.5, 0.1, steps=half, device=dev)
            cyc = torch.stack([torch.roll(base, -i * (half // G)) for i in range(G)])
            S = torch.cat([base.unsqueeze(0).repeat(G, 1), 0.2 * cyc], dim=1)
        self.lora_S.weight.data = S.to(torch.float32)
        if self.global_s:
            # upstream leaves this 1-D after reset_parameters (:419-422); state-dict shape is part of the API
            self.lora_S_global.weight.data = torch.linspace(1, 0.1, steps=r, device=dev).to(torch.float32)
        nn.init.normal_(self.lora_B.weight)

    def _s_eff(self, attr, device, lam=LAMBDA_GROUP):
        a = _attr_on(device, attr)
        sg = self.lora_S_global.weight if self.global_s else None
        return ops.effective_singular_values(a, self.lora_S.weight, sg, lam)

    def forward(self, x, attr=None, batch_first: bool = False):
        return self._run(x, self._s_eff(attr, x.device), batch_first)

    def weight(self, x, attr=None):
        """Per-sample merged weight [B', out, in] with a HARD one-hot mixture (:425-445). RN50 attention pool only."""
        S = self.lora_S.weight
        if attr is not None:
            pi = F.one_hot(attr.to(x.device), num_classes=self.num_attrs).to(S.dtype)
        else:
            pi = torch.full((1, self.num_attrs), 1.0 / self.num_attrs, device=x.device, dtype=S.dtype)
        s_eff = pi @ S
        if self.global_s:
            s_eff = s_eff + self.lora_S_global.weight.reshape(1, -1)
        num_slices = x.shape[1] // s_eff.shape[0]
        s_rows = s_eff.repeat_interleave(num_slices, dim=0)
        dw = (self.lora_A.weight.unsqueeze(0) * s_rows.unsqueeze(1)) @ self.lora_B.weight
        w = self.original_linear.weight.reshape(self.out_features, self.in_features)
        return w.unsqueeze(0) + self.scaling * dw.transpose(1, 2)


class SVLoRALinear(_AdapterBase):
    """One global singular-value vector: the G = 1 special case (reference :255-331)."""

    def __init__(self, original_linear, rank=4, alpha=0.4, global_s=False):
        super().__init__()
        self._init_base(original_linear, rank, alpha)
        self.global_s = global_s
        self.lora_A = self._new_embedding(self.in_features, rank)
        self.lora_S = self._new_embedding(rank, 1)
        if self.global_s:
            self.lora_S_global = self._new_embedding(rank, 1)
        self.lora_B = self._new_embedding(rank, self.out_features)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.zeros_(self.lora_A.weight)
        dev = self.lora_S.weight.device
        self.lora_S.weight.data = torch.linspace(1, 0.1, steps=self.rank, device=dev)          # 1-D, as upstream
        if self.global_s:
            self.lora_S_global.weight.data = torch.linspace(1, 0.1, steps=self.rank, device=dev)
        nn.init.normal_(self.lora_B.weight)

    def forward(self, x, attr=None):
        s = self.lora_S.weight.reshape(1, -1)
        if self.global_s:
            s = s + self.lora_S_global.weight.reshape(1, -1)
        return self._run(x, s.contiguous())


class LoRALinear(_AdapterBase):
    """Plain LoRA: no singular values (reference :203-252)."""

    def __init__(self, original_linear, rank=4, alpha=0.04):
        super().__init__()
        self._init_base(original_linear, rank, alpha)
        self.lora_A = self._new_embedding(self.in_features, rank)
        self.lora_B = self._new_embedding(rank, self.out_features)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.zeros_(self.lora_A.weight)
        nn.init.normal_(self.lora_B.weight)

    def weight(self, x=None, attr=None):
        """Merged weight W + scaling (A B)^T (:235-239): one fused pass with its own backward on the GPU."""
        w = self.original_linear.weight
        if not w.is_cuda:
            raise RuntimeError("fairfedmed_b200 adapters need CUDA tensors: there is no CPU fallback")
        return ops.lora_merged_weight(w.reshape(self.out_features, self.in_features), self.lora_A.weight,
                                      self.lora_B.weight, self.scaling)

    def forward(self, x, attr=None):
        ones = torch.ones((1, self.rank), device=x.device, dtype=torch.float32)
        return self._run(x, ones)


def apply_lora_to_model(model, unfreeze_image_encoder, rank=4, alpha=0.04, lora_type="loRA", global_s=False,
                        num_attrs=1, adapt_attention=False):
    """Module surgery with the reference's selection rules (:503-573): ViT -> the `.mlp.` linears of
    `image_encoder.*`; ResNet -> 1x1 convs named conv* under layer1-4 (FairLoRA) and the attention-pool
    linears (plain LoRA).

    adapt_attention (opt-in, OFF for parity with the reference, which adapts the MLP only — SURVEY F2): additionally puts
    a FairLoRA adapter on the packed in_proj (C -> 3C) and on out_proj of every image-tower attention block, as the
    north-star's "every attention and MLP linear" words it.  New state-dict keys:
    `image_encoder.transformer.resblocks.{i}.attn_in_lora.{lora_A,lora_S,lora_B}.weight`, `...attn_out_lora...`."""
    named = dict(model.named_modules())
    if adapt_attention and unfreeze_image_encoder:
        if lora_type != "FairLoRA":
            raise NotImplementedError("adapt_attention is built for FairLoRA adapters")
        for name, module in named.items():
            if name.startswith("image_encoder.") and hasattr(module, "attn") and hasattr(module, "attn_in_lora") \
                    and isinstance(module.attn, nn.MultiheadAttention):
                a = module.attn
                module.attn_in_lora = FairLoRALinear(FrozenLinearView(a.in_proj_weight, a.in_proj_bias), rank=rank,
                                                     alpha=alpha, global_s=global_s, num_attrs=num_attrs)
                module.attn_out_lora = FairLoRALinear(FrozenLinearView(a.out_proj.weight, a.out_proj.bias), rank=rank,
                                                      alpha=alpha, global_s=global_s, num_attrs=num_attrs)
    for name, module in named.items():
        if not (unfreeze_image_encoder and name.startswith("image_encoder.")):
            continue
        new = None
        if isinstance(module, nn.Linear) and ".mlp." in name:
            if lora_type == "LoRA":
                new = LoRALinear(module, rank=rank, alpha=alpha)
            elif lora_type == "SVLoRA":
                new = SVLoRALinear(module, rank=rank, alpha=alpha, global_s=global_s)
            elif lora_type == "FairLoRA":
                new = FairLoRALinear(module, rank=rank, alpha=alpha, global_s=global_s, num_attrs=num_attrs)
            else:
                raise NotImplementedError(lora_type)
        elif name.startswith("image_encoder.layer") or name.startswith("image_encoder.attnpool"):
            is_conv = isinstance(module, nn.Conv2d) and "conv" in name and tuple(module.weight.shape[-2:]) == (1, 1)
            is_pool = "attnpool" in name and isinstance(module, nn.Linear)
            if is_pool:
                new = LoRALinear(module, rank=rank, alpha=alpha)
            elif is_conv:
                if lora_type != "FairLoRA":
                    raise NotImplementedError(lora_type)
                new = FairLoRALinear(module, rank=rank, alpha=alpha, global_s=global_s, num_attrs=num_attrs)
        if new is None:
            continue
        parent = model
        parts = name.split(".")
        for part in parts[:-1]:
            parent = getattr(parent, part)
        setattr(parent, parts[-1], new)
    return model
