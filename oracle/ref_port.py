"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's FairLoRA hot path.

Plain torch-CPU fp32 / numpy restatement of the algorithms on the path, each function citing the
reference file:line it follows (paths relative to the reference repo, mounted at /root/reference in the
build container).  It is the checker for the CUDA path and the `cpu_baseline` / `--impl reference` arm of
bench.py; the product (fairfedmed_b200/) never imports it.

Pinned against the real reference: tests/golden/make_golden.py imports the reference through
oracle/shim.py, runs both on the same seeded inputs and commits the reference outputs as fixtures
(tests/golden/*.npz); tests/test_oracle_golden.py re-checks this file against those fixtures on every run.
Exception — "parity unpinned": DPD / EOD / AOD come from fairlearn / aif360, which are neither vendored nor
pinned by the reference nor installed here; they are restated from their public definitions.
"""
from __future__ import annotations

import math
from typing import Mapping, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

LAMBDA_GROUP = 0.7  # trainers/GLP_OT_SVLoRA.py:459


# =====================================================================================================
# FairLoRA / SVLoRA / LoRA linear                                trainers/GLP_OT_SVLoRA.py:203-500
# =====================================================================================================
def group_mixture(attr: Optional[torch.Tensor], num_groups: int, dtype=torch.float32, hard: bool = False):
    """pi [B, G] (or [1, G] when attr is None) — :453-462 (soft 0.7 mix) and :427-432 (hard one-hot, .weight())."""
    if attr is None:
        return torch.full((1, num_groups), 1.0 / num_groups, dtype=dtype)
    one_hot = F.one_hot(attr.long(), num_classes=num_groups).to(dtype)
    if hard:
        return one_hot
    return one_hot * LAMBDA_GROUP + (1 - one_hot) * (1 - LAMBDA_GROUP) / (num_groups - 1)


def effective_singular_values(attr, S: torch.Tensor, S_global: Optional[torch.Tensor] = None, hard=False):
    """s_eff [B or 1, r] = pi @ S (+ S_global) — :464-467."""
    pi = group_mixture(attr, S.shape[0], S.dtype, hard=hard).to(S.device)   # (device-agnostic: the GPU-eager baseline)
    s_eff = pi @ S
    if S_global is not None:
        s_eff = s_eff + S_global.reshape(1, -1)
    return s_eff


def fairlora_linear(x, W, bias, A, S, B, attr, scaling: float, S_global=None):
    """FairLoRALinear.forward for the nn.Linear case — :450-482.

    x [L, B', C_in] sequence-first; attr int64 [B] or None; B' = B * num_slices (:473-475).
    The reference materialises diag(s_eff[b]) and contracts with einsum('nbr,brr->nbr'), which selects the
    diagonal, i.e. an element-wise product with s_eff of the row's sample.
    """
    y = F.linear(x, W, bias)
    s_eff = effective_singular_values(attr, S, S_global)          # [B or 1, r]
    num_slices = x.shape[1] // s_eff.shape[0]
    s_rows = s_eff.repeat_interleave(num_slices, dim=0)            # [B', r]
    h = x @ A                                                       # [L, B', r]
    dy = ((h * s_rows.unsqueeze(0)) @ B) * scaling
    return y + dy


def fairlora_conv1x1(x, W, bias, A, S, B, attr, scaling: float, S_global=None):
    """FairLoRALinear.forward for a wrapped 1x1 nn.Conv2d (RN50 trunk) — :469-471, :479-480."""
    b, c_in, hh, ww = x.shape
    y = F.conv2d(x, W, bias)
    tokens = x.reshape(b, c_in, hh * ww).permute(2, 0, 1)          # [hw, b, c]
    s_eff = effective_singular_values(attr, S, S_global)
    num_slices = tokens.shape[1] // s_eff.shape[0]
    s_rows = s_eff.repeat_interleave(num_slices, dim=0)
    dy = (((tokens @ A) * s_rows.unsqueeze(0)) @ B) * scaling      # [hw, b, c_out]
    return y + dy.reshape(hh, ww, b, -1).permute(2, 3, 0, 1)


def svlora_linear(x, W, bias, A, s, B, scaling: float, s_global=None):
    """SVLoRALinear.forward (one global s, G = 1) — :308-312."""
    sv = s.reshape(-1) if s_global is None else (s + s_global).reshape(-1)
    return F.linear(x, W, bias) + (((x @ A) * sv) @ B) * scaling


def lora_linear(x, W, bias, A, B, scaling: float):
    """LoRALinear.forward — :241-242."""
    return F.linear(x, W, bias) + ((x @ A) @ B) * scaling


def lora_merged_weight(W, A, B, scaling: float):
    """LoRALinear.weight — :235-236."""
    return W + scaling * (A @ B).t()


def fairlora_merged_weight(W, A, S, B, attr, scaling: float, n_cols: int, S_global=None):
    """FairLoRALinear.weight: per-sample merged weight [B', out, in] with a HARD one-hot mixture — :425-445."""
    s_eff = effective_singular_values(attr, S, S_global, hard=True)
    num_slices = n_cols // s_eff.shape[0]
    s_rows = s_eff.repeat_interleave(num_slices, dim=0)            # [B', r]
    dw = (A.unsqueeze(0) * s_rows.unsqueeze(1)) @ B                # [B', in, out]
    return W.unsqueeze(0) + scaling * dw.transpose(1, 2)


def fairlora_init_S(num_groups: int, rank: int, dtype=torch.float32):
    """'same+cycle' initialisation of lora_S — :403-417."""
    half = rank // 2
    base = torch.linspace(0.5, 0.1, steps=half).to(dtype)
    shift = half // num_groups
    cyc = torch.stack([torch.roll(base, -i * shift) for i in range(num_groups)])
    return torch.cat([base.unsqueeze(0).repeat(num_groups, 1), 0.2 * cyc], dim=1)


# =====================================================================================================
# GLP_OT head                                                     trainers/GLP_OT_SVLoRA.py:615-757
# =====================================================================================================
def sinkhorn(K: torch.Tensor, u: torch.Tensor, v: torch.Tensor, thresh: float, max_iter: int):
    """CustomCLIP.Sinkhorn — :615-634. Returns (T, iterations). ONE stopping decision for the whole batch."""
    r = torch.ones_like(u)
    c = torch.ones_like(v)
    iters = 0
    Kt = K.transpose(1, 2)
    for _ in range(max_iter):
        r_prev = r
        r = u / torch.bmm(K, c.unsqueeze(-1)).squeeze(-1)
        c = v / torch.bmm(Kt, r.unsqueeze(-1)).squeeze(-1)
        iters += 1
        if (r - r_prev).abs().mean().item() < thresh:
            break
    return r.unsqueeze(-1) * c.unsqueeze(-2) * K, iters


def entropic_cot(a: torch.Tensor, b: torch.Tensor, K: torch.Tensor, thresh: float, max_iter: int):
    """CustomCLIP.entropic_COT_fast — :636-675 (K is already exp(-M/eps); `reg` is unused upstream)."""
    u = torch.ones_like(a)
    v = torch.ones_like(b)
    Kp = K / a.unsqueeze(-1)                       # diag(1/a) K
    Kq = K.transpose(1, 2) / b.unsqueeze(-1)       # diag(1/b) K^T
    iters = 0
    one = torch.ones_like(a)
    while iters < max_iter:
        v_prev = v
        u = torch.minimum(one / torch.bmm(Kp, v.unsqueeze(-1)).squeeze(-1), one)
        v = torch.ones_like(b) / torch.bmm(Kq, u.unsqueeze(-1)).squeeze(-1)
        iters += 1
        if (v - v_prev).abs().mean().item() < thresh:
            break
    return u.unsqueeze(-1) * K * v.unsqueeze(-2), iters


def ot_head(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale: torch.Tensor, *, n_cls: int,
            batch: int, ot: str = "Sinkhorn", eps: float = 0.1, thresh: float = 1e-3, max_iter: int = 100,
            top_percent: float = 0.8, return_aux: bool = False):
    """Head of CustomCLIP.forward — :696-757.

    image_features [M+1, B', D] (token 0 is the pooled token and is dropped, :696-697);
    text_features [N*n_cls, D] in prompt-major order (viewed [N, n_cls, D], :710); `batch` = b of :678.
    Returns logits [batch, n_cls], or None when the transport plan contains NaN (:738-743).
    """
    feats = image_features[1:]
    M, Bp, D = feats.shape
    N = text_features.shape[0] // n_cls
    txt = text_features.contiguous().view(N, n_cls, D)
    feats = F.normalize(feats, dim=2)
    txt = F.normalize(txt, dim=2)
    sim = torch.einsum("mbd,ncd->mnbc", feats, txt).contiguous()
    sim = sim.view(M, N, Bp * n_cls).permute(2, 0, 1)              # [B'*n_cls, M, N]
    iters = 0
    if ot == "None":
        T = None
        sim_op = sim.mean(dim=(1, 2))
    else:
        with torch.no_grad():
            KK = torch.exp(-(1.0 - sim) / eps)
            xx = torch.full((sim.shape[0], M), 1.0 / M, dtype=sim.dtype, device=sim.device)
            if ot == "Sinkhorn":
                yy = torch.full((sim.shape[0], N), 1.0 / N, dtype=sim.dtype, device=sim.device)
                T, iters = sinkhorn(KK, xx, yy, thresh, max_iter)
            elif ot == "COT":
                tp = min(float(xx.sum().item()), top_percent)     # :727
                yy = torch.full((sim.shape[0], N), 1.0 / N, dtype=sim.dtype, device=sim.device) * tp
                T, iters = entropic_cot(xx, yy, KK, thresh, max_iter)
            else:
                raise NotImplementedError(ot)
            if torch.isnan(T).any():
                return (None, None, sim, iters) if return_aux else None
        sim_op = (T * sim).sum(dim=(1, 2))
    sim_op = sim_op.contiguous().view(batch, -1, n_cls).mean(1)
    logits = logit_scale.exp() * sim_op
    if return_aux:
        return logits, T, sim, iters
    return logits


# =====================================================================================================
# CLIP ViT image encoder / text encoder / prompt learner (functional, keyed by the reference's state-dict names)
#                                    clip/model.py:304-449, trainers/GLP_OT_SVLoRA.py:46-200, :677-763
# =====================================================================================================
PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)
PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)


def _ln(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5).to(x.dtype)        # clip/model.py:304-310


def _quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)                                            # clip/model.py:313-315


def attention_core(qkv: torch.Tensor, n_head: int, mask=None):
    """The part of nn.MultiheadAttention between in_proj and out_proj (torch F.multi_head_attention_forward as called
    by clip/model.py:350-352, need_weights=False, no dropout): qkv [L, B, 3C] sequence-first, packed [q | k | v] with
    heads contiguous inside each third -> [L, B, C].  `mask`: additive [L, L] (the text tower's causal mask)."""
    L, Bn, c3 = qkv.shape
    C = c3 // 3
    q, k, v = qkv.chunk(3, dim=-1)
    hd = C // n_head

    def heads(t):
        return t.reshape(L, Bn * n_head, hd).transpose(0, 1)                       # [B*h, L, hd]

    q, k, v = heads(q), heads(k), heads(v)
    att = torch.bmm(q * (hd ** -0.5), k.transpose(1, 2))
    if mask is not None:
        att = att + mask
    att = torch.softmax(att, dim=-1)
    return torch.bmm(att, v).transpose(0, 1).reshape(L, Bn, C)


def _mha(x, p: Mapping[str, torch.Tensor], pre: str, n_head: int, mask=None):
    """nn.MultiheadAttention self-attention, sequence-first — clip/model.py:350-352."""
    qkv = F.linear(x, p[pre + "in_proj_weight"], p[pre + "in_proj_bias"])
    out = attention_core(qkv, n_head, mask)
    return F.linear(out, p[pre + "out_proj.weight"], p[pre + "out_proj.bias"])


def _adapted_linear(x, p, pre: str, attr, scaling: float, lora_type: str):
    """MLP.c_fc / c_proj after apply_lora_to_model (:503-540); falls back to a plain linear when un-adapted."""
    if pre + "original_linear.weight" not in p:
        return F.linear(x, p[pre + "weight"], p[pre + "bias"])
    W, b = p[pre + "original_linear.weight"], p[pre + "original_linear.bias"]
    A, Bm = p[pre + "lora_A.weight"], p[pre + "lora_B.weight"]
    if lora_type == "LoRA":
        return lora_linear(x, W, b, A, Bm, scaling)
    if lora_type == "SVLoRA":
        return svlora_linear(x, W, b, A, p[pre + "lora_S.weight"], Bm, scaling, p.get(pre + "lora_S_global.weight"))
    return fairlora_linear(x, W, b, A, p[pre + "lora_S.weight"], Bm, attr, scaling,
                           p.get(pre + "lora_S_global.weight"))


def transformer(x, p, pre: str, n_layers: int, n_head: int, attr=None, scaling=0.0, lora_type="FairLoRA", mask=None):
    """Transformer / ResidualAttentionBlock / MLP — clip/model.py:317-374."""
    for i in range(n_layers):
        bp = f"{pre}resblocks.{i}."
        x = x + _mha(_ln(x, p[bp + "ln_1.weight"], p[bp + "ln_1.bias"]), p, bp + "attn.", n_head, mask)
        hdn = _ln(x, p[bp + "ln_2.weight"], p[bp + "ln_2.bias"])
        hdn = _adapted_linear(hdn, p, bp + "mlp.c_fc.", attr, scaling, lora_type)
        hdn = _quick_gelu(hdn)
        hdn = _adapted_linear(hdn, p, bp + "mlp.c_proj.", attr, scaling, lora_type)
        x = x + hdn
    return x


def vit_image_encoder(image, p, attr, *, n_layers=12, n_head=12, patch=16, scaling=1.0 / 6.0, lora_type="FairLoRA",
                      pre="image_encoder."):
    """ModifiedVisionTransformer.forward: returns ALL tokens, sequence-first [197, B', 512] — clip/model.py:430-449."""
    x = F.conv2d(image, p[pre + "conv1.weight"], stride=patch)
    x = x.flatten(2).transpose(1, 2)                                               # [B', 196, width]
    cls = p[pre + "class_embedding"].to(x.dtype).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + p[pre + "positional_embedding"].to(x.dtype)
    x = _ln(x, p[pre + "ln_pre.weight"], p[pre + "ln_pre.bias"])
    x = x.transpose(0, 1)                                                          # -> [L, B', width]
    x = transformer(x, p, pre + "transformer.", n_layers, n_head, attr, scaling, lora_type)
    x = x.transpose(0, 1)
    x = _ln(x, p[pre + "ln_post.weight"], p[pre + "ln_post.bias"]) @ p[pre + "proj"]
    return x.transpose(0, 1)


def prompt_embeddings(p, n_prompts: int, n_cls: int, pre="prompt_learner."):
    """PromptLearner.forward with a generic context and class token at the end — :131-152."""
    ctx = p[pre + "ctx"]                                                           # [N, n_ctx, D]
    n_ctx, d = ctx.shape[1], ctx.shape[2]
    ctx = ctx.unsqueeze(0).expand(n_cls, -1, -1, -1).permute(1, 0, 2, 3).contiguous().view(n_prompts * n_cls, n_ctx, d)
    return torch.cat([p[pre + "token_prefix"], ctx, p[pre + "token_suffix"]], dim=1)


def text_encoder(prompts, eot_index, p, *, n_layers=12, n_head=8, pre="text_encoder."):
    """TextEncoder.forward — :55-66 (causal mask from clip/model.py:562-568)."""
    x = prompts + p[pre + "positional_embedding"].to(prompts.dtype)
    L = x.shape[1]
    mask = torch.full((L, L), float("-inf"), dtype=x.dtype, device=x.device).triu_(1)
    x = transformer(x.transpose(0, 1), p, pre + "transformer.", n_layers, n_head, mask=mask).transpose(0, 1)
    x = _ln(x, p[pre + "ln_final.weight"], p[pre + "ln_final.bias"])
    return x[torch.arange(x.shape[0], device=x.device), eot_index.to(x.device)] @ p[pre + "text_projection"]


def _bn_train(x, p, pre, eps=1e-5):
    """nn.BatchNorm2d in training mode (batch statistics, biased variance for the normalisation) — the reference never
    calls .eval() on the trunk while training (Dassl/dassl/engine/trainer.py:156-165 sets train mode)."""
    return F.batch_norm(x, None, None, p[pre + "weight"], p[pre + "bias"], True, 0.0, eps)


def _rn_conv1x1(x, p, pre, attr, scaling):
    """A 1x1 conv of a Bottleneck: FairLoRA-wrapped (keys under `original_linear`) or plain."""
    if pre + "lora_A.weight" in p:
        return fairlora_conv1x1(x, p[pre + "original_linear.weight"], None, p[pre + "lora_A.weight"],
                                p[pre + "lora_S.weight"], p[pre + "lora_B.weight"], attr, scaling)
    return F.conv2d(x, p[pre + "weight"])


def resnet_image_encoder(image, p, attr, *, layers=(3, 4, 6, 3), n_head=32, scaling=0.25, pre="image_encoder."):
    """ModifiedResNet_GLP_OT.forward (clip/model.py:270-301) with Bottleneck.forward (:41-60) and
    AttentionPool2d.forward (:75-118): returns all tokens [HW+1, B, output_dim]."""
    x = image
    for i, stride in ((1, 2), (2, 1), (3, 1)):                                           # 3-conv stem
        x = F.relu(_bn_train(F.conv2d(x, p[f"{pre}conv{i}.weight"], stride=stride, padding=1), p, f"{pre}bn{i}."))
    x = F.avg_pool2d(x, 2)
    for li, n_blocks in enumerate(layers, start=1):
        for bi in range(n_blocks):
            bp = f"{pre}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            identity = x
            out = F.relu(_bn_train(_rn_conv1x1(x, p, bp + "conv1.", attr, scaling), p, bp + "bn1."))
            out = F.relu(_bn_train(F.conv2d(out, p[bp + "conv2.weight"], padding=1), p, bp + "bn2."))
            if stride > 1:
                out = F.avg_pool2d(out, stride)
            out = _bn_train(_rn_conv1x1(out, p, bp + "conv3.", attr, scaling), p, bp + "bn3.")
            if bp + "downsample.0.weight" in p:
                identity = F.avg_pool2d(x, stride) if stride > 1 else x
                identity = _bn_train(F.conv2d(identity, p[bp + "downsample.0.weight"]), p, bp + "downsample.1.")
            x = F.relu(out + identity)
    # attention pool: every token (mean first) is a query; projections are plain-LoRA merged weights (:235-236)
    b, c, hh, ww = x.shape
    t = x.reshape(b, c, hh * ww).permute(2, 0, 1)
    t = torch.cat([t.mean(dim=0, keepdim=True), t], dim=0) + p[pre + "attnpool.positional_embedding"][:, None, :]
    ap = pre + "attnpool."

    def wb(name):
        if ap + name + ".lora_A.weight" in p:
            w = lora_merged_weight(p[ap + name + ".original_linear.weight"], p[ap + name + ".lora_A.weight"],
                                   p[ap + name + ".lora_B.weight"], scaling)
            return w, p[ap + name + ".original_linear.bias"]
        return p[ap + name + ".weight"], p[ap + name + ".bias"]
    L, hd = t.shape[0], c // n_head
    q, k, v = (F.linear(t, *wb(nm)).reshape(L, b, n_head, hd).permute(1, 2, 0, 3) for nm in ("q_proj", "k_proj", "v_proj"))
    att = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v                # [B, H, L, hd]
    out = att.permute(2, 0, 1, 3).reshape(L, b, c)
    return F.linear(out, *wb("c_proj"))


def custom_clip_forward(image, attr, p, eot_index, *, n_prompts=2, n_cls=2, ot="None", eps=0.1, thresh=1e-3,
                        max_iter=100, top_percent=0.8, vision_layers=12, vision_heads=12, text_layers=12,
                        text_heads=8, scaling=1.0 / 6.0, lora_type="FairLoRA", dim_per_3d_slice=None):
    """CustomCLIP.forward for FairFedMed / FedChexMimic inputs (raw 0..255 pixels) — :677-763."""
    b = image.shape[0]
    hh, ww = image.shape[-2:]
    image = image / 255.0
    if dim_per_3d_slice is not None:                                               # OCT volumes, :681-690
        image = image.reshape(-1, dim_per_3d_slice, hh, ww)
        image = F.conv2d(image, p["proj_per_3d_slice.weight"], p["proj_per_3d_slice.bias"], padding=2)
        lo = image.amin(dim=(1, 2, 3), keepdim=True)
        hi = image.amax(dim=(1, 2, 3), keepdim=True)
        image = (image - lo) / (hi - lo + 1e-5)
    mean = torch.tensor(PIXEL_MEAN, dtype=image.dtype, device=image.device).reshape(1, -1, 1, 1)
    std = torch.tensor(PIXEL_STD, dtype=image.dtype, device=image.device).reshape(1, -1, 1, 1)
    image = (image - mean) / std
    if isinstance(vision_layers, (tuple, list)):       # CLIP ResNet backbone: vision_heads = width * 32 // 64
        feats = resnet_image_encoder(image, p, attr, layers=tuple(vision_layers), n_head=vision_heads, scaling=scaling)
    else:
        feats = vit_image_encoder(image, p, attr, n_layers=vision_layers, n_head=vision_heads, scaling=scaling,
                                  lora_type=lora_type)
    prompts = prompt_embeddings(p, n_prompts, n_cls)
    txt = text_encoder(prompts, eot_index, p, n_layers=text_layers, n_head=text_heads)
    return ot_head(feats, txt, p["logit_scale"], n_cls=n_cls, batch=b, ot=ot, eps=eps, thresh=thresh,
                   max_iter=max_iter, top_percent=top_percent)


def sgd_double_step(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], bufs: list, lr: float,
                    momentum: float = 0.9, weight_decay: float = 5e-4, n_steps: int = 2):
    """torch.optim.SGD stepped `n_steps` times on the same gradient (the reference registers one optimizer under two
    model names: trainers/GLP_OT_SVLoRA.py:864-871, Dassl/dassl/engine/trainer.py:333-342)."""
    with torch.no_grad():
        for _ in range(n_steps):
            for i, (w, g) in enumerate(zip(params, grads)):
                d = g + weight_decay * w
                if bufs[i] is None:
                    bufs[i] = d.clone()
                else:
                    bufs[i].mul_(momentum).add_(d)
                w.add_(bufs[i], alpha=-lr)
    return bufs


# =====================================================================================================
# Federated aggregation                                            utils/fed_utils.py:6-100
# =====================================================================================================
def average_weights_ema(w_g, w, idxs_users, n_client, n_client_by_attr, epoch, max_epoch, beta=0.999,
                        shared_half_s=False):
    """average_weights_EMA (dict branch) — utils/fed_utils.py:42-100."""
    total = sum(n_client[k] for k in idxs_users)
    by_attr = None
    if n_client_by_attr is not None:
        by_attr = torch.tensor(n_client_by_attr)
        total_by_attr = by_attr[list(idxs_users)].sum(0)
    out = {}
    first = w[idxs_users[0]]
    for key in first:
        acc = None
        for k in idxs_users:
            t = w[k][key]
            if by_attr is not None and "lora_S" in key and t.shape[0] == by_attr.shape[1]:
                coef = (by_attr[k] / total_by_attr)[:, None].to(t.device)
            else:
                coef = n_client[k] / total
            term = t * coef
            acc = term if acc is None else acc + term
        out[key] = acc
    beta_decay = beta * (epoch / max(max_epoch, 1))
    for key in out:
        t = out[key]
        if shared_half_s and by_attr is not None and "lora_S" in key and t.shape[0] == by_attr.shape[1]:
            g, r = t.shape
            t = torch.cat([t[:, : r // 2].mean(dim=0, keepdim=True).repeat(g, 1), t[:, r // 2:]], dim=1)
        out[key] = (1 - beta_decay) * t + beta_decay * w_g[key]
    return out


def average_weights(w, idxs_users, n_client, n_client_by_attr=None):
    """average_weights (dict branch) — utils/fed_utils.py:6-40; the EMA variant with beta_decay = 0, no shared half."""
    zero_g = {k: torch.zeros_like(v) for k, v in w[idxs_users[0]].items()}
    return average_weights_ema(zero_g, w, idxs_users, n_client, n_client_by_attr, 0, 1, shared_half_s=False)


# =====================================================================================================
# Fairness metrics                                                 evaluation/metrics.py:197-356, 486-550
# =====================================================================================================
def mann_whitney_counts(score: np.ndarray, positive: np.ndarray):
    """(#{pos > neg}, #{pos == neg}, P, Nn) on the stored values — the integer core of roc_auc_score
    (sklearn AUC == (gt + eq/2) / (P*Nn) exactly)."""
    score = np.asarray(score)
    positive = np.asarray(positive).astype(bool)
    order = np.argsort(score, kind="stable")
    s, pos = score[order], positive[order]
    n = s.shape[0]
    # boundaries of tie groups
    new = np.ones(n, dtype=bool)
    new[1:] = s[1:] != s[:-1]
    grp = np.cumsum(new) - 1
    n_grp = int(grp[-1]) + 1 if n else 0
    neg_in = np.bincount(grp, weights=(~pos).astype(np.float64), minlength=n_grp).astype(np.int64)
    pos_in = np.bincount(grp, weights=pos.astype(np.float64), minlength=n_grp).astype(np.int64)
    neg_below = np.cumsum(neg_in) - neg_in
    gt = int((pos_in * neg_below).sum())
    eq = int((pos_in * neg_in).sum())
    return gt, eq, int(pos.sum()), int((~pos).sum())


def auc_from_counts(gt: int, eq: int, P: int, Nn: int) -> float:
    if P == 0 or Nn == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return (gt + 0.5 * eq) / (P * Nn)


def compute_auc(pred_prob: np.ndarray, y: np.ndarray, num_classes: int = 2) -> float:
    """compute_auc — evaluation/metrics.py:340-356. [N] scores: binary ROC AUC; [N, C] probs: macro one-vs-rest."""
    pred_prob = np.asarray(pred_prob)
    y = np.asarray(y)
    if num_classes == 2 and pred_prob.shape == y.shape:
        return auc_from_counts(*mann_whitney_counts(pred_prob, y == 1))
    vals = [auc_from_counts(*mann_whitney_counts(pred_prob[:, c], y.astype(int) == c)) for c in range(num_classes)]
    return float(np.mean(vals))


def accuracy_top1(prob: np.ndarray, y: np.ndarray) -> float:
    """accuracy(..., topk=(1,)) — evaluation/metrics.py:313-338 (argmax; ties -> lowest index like topk)."""
    if prob.ndim == 1:
        return float(np.sum((prob >= 0.5).astype(float) == y) / y.shape[0])
    return float(np.sum(np.argmax(prob, axis=1) == y) / y.shape[0])


def equity_scaled_accuracy(prob, y, attrs, alpha=1.0) -> float:
    """evaluation/metrics.py:486-511 — note: group -1 is NOT skipped here."""
    pred = np.argmax(prob, axis=1) if prob.ndim >= 2 else (prob >= 0.5).astype(float)
    overall = np.sum(pred == y) / y.shape[0]
    gap = 0.0
    for g in np.unique(attrs).astype(int):
        m = attrs == g
        gap += np.abs(np.sum(pred[m] == y[m]) / y[m].shape[0] - overall)
    return float(overall / (alpha * gap + 1))


def equity_scaled_auc(prob, y, attrs, alpha=1.0, num_classes=2) -> float:
    """evaluation/metrics.py:513-547 — group -1 IS skipped here."""
    overall = compute_auc(prob, y, num_classes)
    gap = 0.0
    for g in np.unique(attrs).astype(int):
        if g == -1:
            continue
        m = attrs == g
        gap += np.abs(compute_auc(prob[m], y[m], num_classes) - overall)
    return float(overall / (alpha * gap + 1))


def between_group_disparity(aucs, overall_auc):
    """evaluation/metrics.py:549-550."""
    return float(np.std(aucs) / overall_auc), float((np.max(aucs) - np.min(aucs)) / overall_auc)


def _rates_by_group(y_true, y_pred, sensitive):
    out = {}
    for g in np.unique(sensitive):
        m = sensitive == g
        yt, yp = y_true[m], y_pred[m]
        pos, neg = yt == 1, yt == 0
        out[g] = dict(
            sel=float(np.mean(yp == 1)) if m.sum() else float("nan"),
            tpr=float(np.mean(yp[pos] == 1)) if pos.sum() else float("nan"),
            fpr=float(np.mean(yp[neg] == 1)) if neg.sum() else float("nan"),
        )
    return out


def demographic_parity_difference(y_true, y_pred, *, sensitive_features):
    """fairlearn definition: max_g P(yhat=1|g) - min_g P(yhat=1|g). (parity unpinned)"""
    r = _rates_by_group(np.asarray(y_true), np.asarray(y_pred), np.asarray(sensitive_features))
    sel = [v["sel"] for v in r.values()]
    return float(np.max(sel) - np.min(sel))


def demographic_parity_ratio(y_true, y_pred, *, sensitive_features):
    r = _rates_by_group(np.asarray(y_true), np.asarray(y_pred), np.asarray(sensitive_features))
    sel = [v["sel"] for v in r.values()]
    return float(np.min(sel) / np.max(sel)) if np.max(sel) > 0 else float("nan")


def equalized_odds_difference(y_true, y_pred, *, sensitive_features):
    """fairlearn definition: max(TPR spread, FPR spread) over groups. (parity unpinned)"""
    r = _rates_by_group(np.asarray(y_true), np.asarray(y_pred), np.asarray(sensitive_features))
    tpr = [v["tpr"] for v in r.values()]
    fpr = [v["fpr"] for v in r.values()]
    return float(max(np.nanmax(tpr) - np.nanmin(tpr), np.nanmax(fpr) - np.nanmin(fpr)))


def equalized_odds_ratio(y_true, y_pred, *, sensitive_features):
    r = _rates_by_group(np.asarray(y_true), np.asarray(y_pred), np.asarray(sensitive_features))
    tpr = [v["tpr"] for v in r.values()]
    fpr = [v["fpr"] for v in r.values()]
    a = np.nanmin(tpr) / np.nanmax(tpr) if np.nanmax(tpr) > 0 else float("nan")
    b = np.nanmin(fpr) / np.nanmax(fpr) if np.nanmax(fpr) > 0 else float("nan")
    return float(np.nanmin([a, b]))


def average_odds_difference(y_true, y_pred, *, prot_attr, priv_group):
    """aif360.sklearn definition: ((FPR_unpriv - FPR_priv) + (TPR_unpriv - TPR_priv)) / 2, unpriv = all other
    samples pooled. (parity unpinned)"""
    y_true, y_pred, prot = np.asarray(y_true), np.asarray(y_pred), np.asarray(prot_attr)
    priv = prot == priv_group

    def rates(m):
        yt, yp = y_true[m], y_pred[m]
        pos, neg = yt == 1, yt == 0
        tpr = np.mean(yp[pos] == 1) if pos.sum() else 0.0
        fpr = np.mean(yp[neg] == 1) if neg.sum() else 0.0
        return tpr, fpr

    tpr_p, fpr_p = rates(priv)
    tpr_u, fpr_u = rates(~priv)
    return float(((fpr_u - fpr_p) + (tpr_u - tpr_p)) / 2)


def comprehensive_scores(prob: np.ndarray, y: np.ndarray, attrs: np.ndarray, num_classes: int = 2):
    """evalute_comprehensive_perf_scores (binary, [N,2] probabilities) — evaluation/metrics.py:197-311.

    Returns the same 9-tuple: overall_acc, esaccs, overall_auc, esaucs, aucs_by_attrs (list of arrays), dpds, eods,
    aods (list), between_group_disparity [n_attr, 2]."""
    overall_acc = accuracy_top1(prob, y)
    overall_auc = compute_auc(prob, y, num_classes)
    esaccs, esaucs, aucs_by_attrs, dpds, eods, aods, disp = [], [], [], [], [], [], []
    pred = prob.argmax(-1)
    for i in range(attrs.shape[0]):
        attr = attrs[i, :]
        esaccs.append(equity_scaled_accuracy(prob, y, attr))
        esaucs.append(equity_scaled_auc(prob, y, attr, num_classes=num_classes))
        g_aucs = [compute_auc(prob[attr == e], y[attr == e], num_classes)
                  for e in np.unique(attr).astype(int) if e != -1]
        aucs_by_attrs.append(np.array(g_aucs))
        disp.append(list(between_group_disparity(g_aucs, overall_auc)))
        dpds.append(demographic_parity_difference(y, pred, sensitive_features=attr))
        eods.append(equalized_odds_difference(y, pred, sensitive_features=attr))
        per_priv = [abs(average_odds_difference(y, pred, prot_attr=attr, priv_group=g)) for g in set(attr.tolist())]
        aods.append(sum(per_priv) / max(len(per_priv), 1))
    return (overall_acc, np.array(esaccs), overall_auc, np.array(esaucs), aucs_by_attrs, np.array(dpds),
            np.array(eods), aods, np.array(disp))


def confusion_counts(prob: np.ndarray, y: np.ndarray, mask: Optional[np.ndarray] = None):
    """(tp, fp, tn, fn) of pred = argmax(prob) restricted to mask."""
    pred = prob.argmax(-1)
    if mask is None:
        mask = np.ones_like(y, dtype=bool)
    p, t = pred[mask], y[mask]
    return (int(np.sum((p == 1) & (t == 1))), int(np.sum((p == 1) & (t == 0))),
            int(np.sum((p == 0) & (t == 0))), int(np.sum((p == 0) & (t == 1))))
