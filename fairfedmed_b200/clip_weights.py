"""Host-side glue between a pretrained CLIP state dict and this package's CustomCLIP — what the reference does in
`load_clip_to_cpu` + `clip.build_model` + `PromptLearner.__init__` (trainers/GLP_OT_SVLoRA.py:23-43,68-132,
clip/model.py:628-700), restated as pure functions on state dicts so the trainer can be built from the reference's
own cfg (`cfg.MODEL.BACKBONE.NAME`) and weights.  Nothing here runs on the hot path.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

# clip/clip.py `_MODELS` names the reference's cfg uses (cfg.MODEL.BACKBONE.NAME) -> CLIP constructor arguments
# (clip/model.py:461-531): vision_layers, vision_width, patch, embed_dim; the text tower is the same for all four.
ARCH_BY_BACKBONE = {
    "ViT-B/16": dict(VISION_LAYERS=12, VISION_WIDTH=768, PATCH=16, EMBED=512),
    "ViT-B/32": dict(VISION_LAYERS=12, VISION_WIDTH=768, PATCH=32, EMBED=512),
    "RN50": dict(VISION_LAYERS=(3, 4, 6, 3), VISION_WIDTH=64, PATCH=None, EMBED=1024),
    "RN101": dict(VISION_LAYERS=(3, 4, 23, 3), VISION_WIDTH=64, PATCH=None, EMBED=512),
}
TEXT_ARCH = dict(TEXT_WIDTH=512, TEXT_LAYERS=12, TEXT_HEADS=8, CONTEXT=77)


def arch_for_backbone(name: str) -> dict:
    if name not in ARCH_BY_BACKBONE:
        raise KeyError(f"unknown CLIP backbone {name!r}; known: {sorted(ARCH_BY_BACKBONE)} "
                       "(or give cfg.MODEL_ARCH explicitly)")
    return {**ARCH_BY_BACKBONE[name], **TEXT_ARCH}


def arch_from_clip_state_dict(sd: Dict[str, torch.Tensor]) -> dict:
    """The shape inference of clip.build_model (clip/model.py:628-655)."""
    if "visual.proj" in sd:
        width = sd["visual.conv1.weight"].shape[0]
        layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
        patch = sd["visual.conv1.weight"].shape[-1]
    else:
        counts = [len({k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}")}) for b in (1, 2, 3, 4)]
        layers, width, patch = tuple(counts), sd["visual.layer1.0.conv1.weight"].shape[0], None
    t_width = sd["ln_final.weight"].shape[0]
    return dict(VISION_LAYERS=layers, VISION_WIDTH=width, PATCH=patch, EMBED=sd["text_projection"].shape[1],
                TEXT_WIDTH=t_width, TEXT_HEADS=t_width // 64, CONTEXT=sd["positional_embedding"].shape[0],
                TEXT_LAYERS=len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")}))


def map_clip_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """CLIP keys -> CustomCLIP keys, as `CustomCLIP.__init__` re-homes the sub-modules (:583-588): `visual.*` ->
    `image_encoder.*`; `transformer.*`, `positional_embedding`, `ln_final.*`, `text_projection` -> `text_encoder.*`;
    `logit_scale` stays.  `token_embedding.weight` is consumed by prompt_buffers_from_clip, not by the model."""
    out = {}
    for k, v in sd.items():
        if k.startswith("visual."):
            out["image_encoder." + k[len("visual."):]] = v
        elif k.startswith("transformer.") or k.startswith("ln_final.") or k in ("positional_embedding",
                                                                                  "text_projection"):
            out["text_encoder." + k] = v
        elif k == "logit_scale":
            out[k] = v
    return out


def tokenize_prompts(classnames: Sequence[str], n_ctx: int, tokenize: Callable[[str], torch.Tensor]) -> torch.Tensor:
    """"X X X X <classname>." per class (trainers/GLP_OT_SVLoRA.py:99-107) -> int tensor [n_cls, context_length]."""
    prefix = " ".join(["X"] * n_ctx)
    return torch.cat([tokenize(prefix + " " + name.replace("_", " ") + ".") for name in classnames])


def prompt_buffers_from_clip(token_embedding: torch.Tensor, tokenized: torch.Tensor, n_prompts: int,
                             n_ctx: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(token_prefix, token_suffix, eot_index) of PromptLearner (:106-119): the tokenized class prompts repeated N
    times, embedded with the frozen token embedding; SOS row, everything after the context slots, EOT position."""
    tok = tokenized.repeat(n_prompts, 1)
    emb = torch.nn.functional.embedding(tok, token_embedding.float())
    return emb[:, :1, :].clone(), emb[:, 1 + n_ctx:, :].clone(), tok.argmax(dim=-1)


def load_clip_into(model: torch.nn.Module, clip_sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
    """Copy a CLIP state dict into a CustomCLIP built for the same architecture (before adapters are applied).
    Every frozen tensor of both towers must be present; prompt-learner tensors are not part of a CLIP checkpoint."""
    mapped = {k: v.float() for k, v in map_clip_state_dict(clip_sd).items()}
    missing, unexpected = model.load_state_dict(mapped, strict=False)
    missing = [k for k in missing if not k.startswith("prompt_learner.") and "proj_per_3d_slice" not in k]
    if strict and (missing or unexpected):
        raise KeyError(f"CLIP checkpoint does not match the model: missing {missing[:8]}, unexpected {unexpected[:8]}")


def resolve_state_dict(obj) -> Optional[Dict[str, torch.Tensor]]:
    """Accept a state dict, an nn.Module / TorchScript archive with .state_dict(), or None."""
    if obj is None:
        return None
    if hasattr(obj, "state_dict") and callable(obj.state_dict):
        return dict(obj.state_dict())
    return dict(obj)
