"""Seeded input recipes shared by the golden-vector generator and the parity tests.

Everything is produced with torch's CPU generator (bit-reproducible for a given torch build, and the GPU box runs
the same image), so the committed fixtures only hold the reference OUTPUTS.
"""
from __future__ import annotations

import numpy as np
import torch

CLASSNAMES = ("NOT Glaucoma", "Glaucoma")


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------- FairLoRA linear
FAIRLORA_CASES = {
    # sequence-first activations [L, B', C_in]
    "vit_small": dict(kind="FairLoRA", L=5, Bp=4, slices=1, c_in=64, c_out=96, rank=12, alpha=2.0, groups=3,
                      global_s=False, seed=11, merged=True),
    "vit_cfc": dict(kind="FairLoRA", L=7, Bp=8, slices=1, c_in=768, c_out=3072, rank=12, alpha=2.0, groups=3,
                    global_s=False, seed=12),
    "vit_cproj": dict(kind="FairLoRA", L=7, Bp=8, slices=1, c_in=3072, c_out=768, rank=12, alpha=2.0, groups=3,
                      global_s=False, seed=13),
    "oct_slices": dict(kind="FairLoRA", L=6, Bp=8, slices=4, c_in=128, c_out=64, rank=12, alpha=2.0, groups=2,
                       global_s=False, seed=14),
    "global_s": dict(kind="FairLoRA", L=4, Bp=6, slices=1, c_in=64, c_out=64, rank=8, alpha=2.0, groups=3,
                     global_s=True, seed=15),
    "no_attr": dict(kind="FairLoRA", L=4, Bp=6, slices=1, c_in=64, c_out=64, rank=12, alpha=2.0, groups=3,
                    global_s=False, seed=16, attr_none=True),
    "svlora": dict(kind="SVLoRA", L=5, Bp=4, slices=1, c_in=64, c_out=96, rank=12, alpha=2.0, groups=1,
                   global_s=False, seed=17),
    "lora": dict(kind="LoRA", L=5, Bp=4, slices=1, c_in=64, c_out=96, rank=12, alpha=2.0, groups=1,
                 global_s=False, seed=18),
}


def fairlora_inputs(rc):
    g = _gen(rc["seed"])
    B = rc["Bp"] // rc["slices"]
    r = rc["rank"]
    t = dict(
        x=torch.randn(rc["L"], rc["Bp"], rc["c_in"], generator=g),
        W=torch.randn(rc["c_out"], rc["c_in"], generator=g) * rc["c_in"] ** -0.5,
        bias=torch.randn(rc["c_out"], generator=g) * 0.1,
        A=torch.randn(rc["c_in"], r, generator=g) * 0.05,
        B=torch.randn(r, rc["c_out"], generator=g),
        dy=torch.randn(rc["L"], rc["Bp"], rc["c_out"], generator=g),
    )
    if rc["kind"] == "FairLoRA":
        t["S"] = torch.rand(rc["groups"], r, generator=g) + 0.1
        t["S_global"] = torch.rand(r, generator=g) + 0.1   # 1-D upstream after reset_parameters (:419-422)
        t["attr"] = None if rc.get("attr_none") else torch.randint(0, rc["groups"], (B,), generator=g)
    elif rc["kind"] == "SVLoRA":
        t["S"] = torch.rand(r, generator=g) + 0.1
        t["S_global"] = torch.rand(r, generator=g) + 0.1
        t["attr"] = None
    else:
        t["attr"] = None
    return t


# ------------------------------------------------------------------------------------------------- Sinkhorn / COT
SINKHORN_CASES = {
    "sk_small": dict(mode="Sinkhorn", P=6, M=49, N=2, eps=0.1, thresh=1e-3, max_iter=100, v_mass=1.0, seed=21),
    "sk_vit": dict(mode="Sinkhorn", P=16, M=196, N=2, eps=0.1, thresh=1e-3, max_iter=100, v_mass=1.0, seed=22),
    "sk_tight": dict(mode="Sinkhorn", P=4, M=196, N=2, eps=0.1, thresh=1e-7, max_iter=100, v_mass=1.0, seed=23),
    "sk_cap": dict(mode="Sinkhorn", P=4, M=64, N=4, eps=0.05, thresh=0.0, max_iter=7, v_mass=1.0, seed=24),
    "cot_vit": dict(mode="COT", P=16, M=196, N=2, eps=0.1, thresh=1e-3, max_iter=100, v_mass=0.8, seed=25),
    "cot_cap": dict(mode="COT", P=4, M=64, N=4, eps=0.05, thresh=0.0, max_iter=9, v_mass=0.8, seed=26),
}


def sinkhorn_inputs(rc):
    """K = exp(-(1 - sim)/eps) for cosine-like sims, u = 1/M, v = v_mass/N (trainers/GLP_OT_SVLoRA.py:721-735)."""
    g = _gen(rc["seed"])
    sim = torch.rand(rc["P"], rc["M"], rc["N"], generator=g) * 0.6 - 0.1
    K = torch.exp(-(1.0 - sim) / rc["eps"])
    u = torch.full((rc["P"], rc["M"]), 1.0 / rc["M"])
    v = torch.full((rc["P"], rc["N"]), 1.0 / rc["N"]) * rc["v_mass"]
    return K, u, v


# ------------------------------------------------------------------------------------------------- aggregation
FEDAVG_CASES = {
    "two_clients": dict(n_clients=2, idxs=[0, 1], groups=3, rank=4, epoch=1, max_epoch=2, shared_half_s=True,
                        seed=31, plain=True, n_k=[10, 30], n_kg=[[5, 5, 0], [5, 10, 15]]),
    "frac": dict(n_clients=4, idxs=[0, 2, 3], groups=3, rank=12, epoch=3, max_epoch=50, shared_half_s=True, seed=32,
                 n_k=[100, 50, 75, 20], n_kg=[[40, 30, 30], [10, 20, 20], [25, 25, 25], [5, 5, 10]]),
    "no_shared": dict(n_clients=3, idxs=[0, 1, 2], groups=2, rank=12, epoch=0, max_epoch=50, shared_half_s=False,
                      seed=33, n_k=[7, 9, 11], n_kg=[[3, 4], [4, 5], [6, 5]]),
    # ResNet trunk: the reference averages the WHOLE state dict, i.e. also BatchNorm running statistics and the int64
    # batch counter (which comes back as a float tensor), utils/fed_utils.py:63-98
    "rn50_bn": dict(n_clients=3, idxs=[0, 2], groups=3, rank=32, epoch=7, max_epoch=50, shared_half_s=True, seed=34,
                    n_k=[64, 32, 96], n_kg=[[20, 20, 24], [10, 12, 10], [40, 30, 26]], batchnorm=True),
}


def fedavg_inputs(rc):
    g = _gen(rc["seed"])
    G, r = rc["groups"], rc["rank"]

    def one():
        if rc.get("batchnorm"):
            p = "image_encoder.layer1.0."
            return {
                "prompt_learner.ctx": torch.randn(2, 4, 16, generator=g),
                p + "conv1.lora_A.weight": torch.randn(16, r, generator=g),
                p + "conv1.lora_S.weight": torch.rand(G, r, generator=g),
                p + "conv1.lora_B.weight": torch.randn(r, 8, generator=g),
                p + "bn1.weight": torch.randn(8, generator=g),
                p + "bn1.bias": torch.randn(8, generator=g),
                p + "bn1.running_mean": torch.randn(8, generator=g),
                p + "bn1.running_var": torch.rand(8, generator=g) + 0.5,
                p + "bn1.num_batches_tracked": torch.randint(1, 50, (), generator=g),
            }
        return {
            "prompt_learner.ctx": torch.randn(2, 4, 16, generator=g),
            "image_encoder.transformer.resblocks.0.mlp.c_fc.lora_A.weight": torch.randn(24, r, generator=g),
            "image_encoder.transformer.resblocks.0.mlp.c_fc.lora_S.weight": torch.rand(G, r, generator=g),
            "image_encoder.transformer.resblocks.0.mlp.c_fc.lora_B.weight": torch.randn(r, 40, generator=g),
            "image_encoder.transformer.resblocks.0.mlp.c_fc.lora_S_global.weight": torch.rand(r, generator=g),
            "image_encoder.transformer.resblocks.0.mlp.c_fc.original_linear.weight": torch.randn(40, 24, generator=g),
        }

    w_g = one()
    w_loc = [one() for _ in range(rc["n_clients"])]
    return w_g, w_loc, list(rc["n_k"]), [list(x) for x in rc["n_kg"]]


# ------------------------------------------------------------------------------------------------- metrics
METRIC_CASES = {
    "m_small": dict(N=200, n_attr=2, groups=[3, 2], unknown=0.0, ties=False, seed=41),
    "m_ties": dict(N=1000, n_attr=3, groups=[3, 3, 2], unknown=0.1, ties=True, seed=42),
    "m_large": dict(N=5000, n_attr=6, groups=[3, 3, 2, 2, 3, 2], unknown=0.05, ties=True, seed=43),
}


def metric_inputs(rc):
    """Softmax probabilities [N,2] float32 (with exact ties when asked), labels {0,1}, attrs [n_attr, N] (-1 unknown).
    Every (attribute, group) — including the -1 pseudo-group where the reference does not skip it — holds both
    classes, as the reference exit()s otherwise (evaluation/metrics.py:224-242)."""
    g = _gen(rc["seed"])
    N = rc["N"]
    y = (torch.arange(N) % 2).to(torch.int64)[torch.randperm(N, generator=g)]
    logits = torch.randn(N, 2, generator=g) + torch.nn.functional.one_hot(y, 2) * 0.8
    if rc["ties"]:
        logits = torch.round(logits * 4) / 4          # coarse grid => many exactly tied probabilities
    prob = torch.softmax(logits.float(), dim=-1)
    attrs = []
    for a, G in enumerate(rc["groups"]):
        col = torch.randint(0, G, (N,), generator=g)
        if rc["unknown"] > 0:
            col[torch.rand(N, generator=g) < rc["unknown"]] = -1
        attrs.append(col)
    return prob.numpy().astype(np.float32), y.numpy(), torch.stack(attrs).numpy()


# ------------------------------------------------------------------------------------------------- whole model
def _adapter_grads(name):
    return "lora_" in name or "prompt_learner.ctx" in name or "proj_per_3d_slice" in name


MODEL_CASES = {
    "tiny_none": dict(modality="slo_fundus", ot="None", res=32, embed=64, v_layers=2, v_width=128, t_width=64,
                      t_heads=2, t_layers=2, rank=12, alpha=2.0, lora_type="FairLoRA", groups=3, batch=4, seed=51,
                      grad_filter=_adapter_grads),
    "tiny_sinkhorn": dict(modality="slo_fundus", ot="Sinkhorn", res=32, embed=64, v_layers=2, v_width=128, t_width=64,
                          t_heads=2, t_layers=2, rank=12, alpha=2.0, lora_type="FairLoRA", groups=3, batch=4, seed=52,
                          grad_filter=_adapter_grads),
    "tiny_cot": dict(modality="slo_fundus", ot="COT", res=32, embed=64, v_layers=1, v_width=128, t_width=64,
                     t_heads=2, t_layers=1, rank=12, alpha=2.0, lora_type="FairLoRA", groups=3, batch=4, seed=53,
                     grad_filter=_adapter_grads),
    "tiny_oct": dict(modality="oct_bscans", ot="Sinkhorn", res=32, embed=64, v_layers=1, v_width=128, t_width=64,
                     t_heads=2, t_layers=1, rank=12, alpha=2.0, lora_type="FairLoRA", groups=2, batch=2, seed=54,
                     dim_per_3d_slice=8, grad_filter=_adapter_grads),
    # width 256 (4 heads): the vision tower takes the fused device path end to end — patchify+normalise kernel,
    # class/positional embedding + ln_pre + ln_1 kernel, add+LayerNorm kernels, fused MLP, text tower on a side stream
    "small_fused": dict(modality="slo_fundus", ot="Sinkhorn", res=64, embed=64, v_layers=2, v_width=256, t_width=64,
                        t_heads=2, t_layers=2, rank=12, alpha=2.0, lora_type="FairLoRA", groups=3, batch=4, seed=56,
                        grad_filter=_adapter_grads),
    # CLIP ResNet backbone (scope row a8): conv trunk with FairLoRA on the 1x1 convs (rank 32, alpha 8 — the RN50 script
    # values) + plain LoRA on the attention pool; BatchNorm in training mode with trainable affine parameters
    "tiny_rn50": dict(modality="slo_fundus", ot="Sinkhorn", res=128, embed=64, v_layers=(1, 1, 1, 1), v_width=16,
                      t_width=64, t_heads=2, t_layers=1, rank=32, alpha=8.0, lora_type="FairLoRA", groups=3, batch=8,
                      seed=55, train_bn=True,
                      grad_filter=lambda n: _adapter_grads(n) or ".bn" in n or "downsample.1" in n),
}


def model_params(rc, shapes: dict):
    """Random state dict for every key of the reference CustomCLIP (name -> shape given by the caller)."""
    g = _gen(rc["seed"])
    out = {}
    for k in shapes:
        shp = shapes[k]
        if k == "logit_scale":
            out[k] = torch.tensor(float(np.log(1 / 0.07)))
        elif k.endswith("running_var"):
            out[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shp, dtype=torch.int64)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(shp, generator=g)
        elif (".bn" in k or "downsample.1" in k) and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("ln_1.weight") or k.endswith("ln_2.weight") or k.endswith("ln_pre.weight") or \
                k.endswith("ln_post.weight") or k.endswith("ln_final.weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif "lora_S" in k:
            out[k] = torch.rand(shp, generator=g) + 0.1
        elif "lora_A" in k:
            out[k] = 0.05 * torch.randn(shp, generator=g)
        elif "lora_B" in k:
            out[k] = torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            out[k] = 0.02 * torch.randn(shp, generator=g)
        elif "token_prefix" in k or "token_suffix" in k or k.endswith("ctx") or "positional_embedding" in k or \
                "class_embedding" in k:
            out[k] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = shp[-1] if len(shp) == 2 else int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            out[k] = torch.randn(shp, generator=g) * (fan_in ** -0.5)
    return out


def model_batch(rc):
    g = _gen(rc["seed"] + 1000)
    b = rc["batch"]
    ch = 32 if rc["modality"] == "oct_bscans" else 3
    if ch == 3:
        image = torch.randint(0, 256, (b, 1, rc["res"], rc["res"]), generator=g).float().repeat(1, 3, 1, 1)
    else:
        image = torch.randint(0, 256, (b, ch, rc["res"], rc["res"]), generator=g).float()
    label = (torch.arange(b) % 2).to(torch.int64)
    attr = torch.randint(0, rc["groups"], (b,), generator=g)
    return image, label, attr
