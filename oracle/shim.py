"""TEST INFRASTRUCTURE ONLY — import the *real* reference (read-only at /root/reference) in-process.

Only usable in the build container (the GPU box has no /root/reference).  It is used by
tests/golden/make_golden.py to generate the committed golden vectors and by the `not gpu`
tests that re-validate the restated oracle (oracle/ref_port.py) against the reference when
the reference tree is present.  Nothing in the product imports this.

The reference cannot be imported as shipped: missing third-party modules, a
missing `datasets/WangGrant.py`, HuggingFace `datasets` shadowing the repo's `datasets/`, and an
import cycle.  `install()` injects minimal stand-ins into sys.modules, then imports the engine first.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("FFM_REFERENCE_ROOT", "/root/reference")

_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "trainers"))


def _module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _AttrDict(dict):
    """Smallest possible stand-in for yacs.config.CfgNode."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def clone(self):
        import copy
        return copy.deepcopy(self)


def install() -> None:
    """Make `import trainers.GLP_OT_SVLoRA`, `clip.model`, `utils.fed_utils`, `evaluation.metrics` work."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.nn as nn

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # a `utils` / `clip` / `evaluation` package of ours must not shadow the reference's
    for shadow in ("utils", "clip", "evaluation", "trainers", "datasets", "Dassl"):
        mod = sys.modules.get(shadow)
        if mod is not None and not str(getattr(mod, "__file__", "") or "").startswith(REF_ROOT):
            del sys.modules[shadow]

    _module("ftfy", fix_text=lambda s: s)
    _module("gdown")

    class PrettyTable:  # noqa: D401 - stand-in
        def __init__(self, *a, **k):
            self.rows = []

        def add_row(self, row):
            self.rows.append(row)

        def __str__(self):
            return "\n".join(map(str, self.rows))

    _module("prettytable", PrettyTable=PrettyTable)
    _module("yacs")
    _module("yacs.config", CfgNode=_AttrDict)
    sys.modules["yacs"].config = sys.modules["yacs.config"]
    timm = _module("timm")
    timm.models = _module("timm.models")
    timm.models.vision_transformer = _module(
        "timm.models.vision_transformer", VisionTransformer=type("VisionTransformer", (nn.Module,), {}))
    sk = _module("skimage")
    sk.transform = _module("skimage.transform", resize=None)

    # fairlearn / aif360 are un-vendored and un-pinned upstream; the stand-ins are the public definitions
    # restated in oracle/ref_port.py ("parity unpinned" for DPD / EOD / AOD).
    from oracle import ref_port as rp

    fl = _module("fairlearn")
    fl.metrics = _module(
        "fairlearn.metrics",
        demographic_parity_difference=rp.demographic_parity_difference,
        demographic_parity_ratio=rp.demographic_parity_ratio,
        equalized_odds_difference=rp.equalized_odds_difference,
        equalized_odds_ratio=rp.equalized_odds_ratio,
    )
    aif = _module("aif360")
    aif.sklearn = _module("aif360.sklearn")
    aif.sklearn.metrics = _module("aif360.sklearn.metrics", average_odds_difference=rp.average_odds_difference)

    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REF_ROOT, "datasets")]
    sys.modules["datasets"] = ds
    _module("datasets.WangGrant", WangGrant=type("WangGrant", (), {}))

    import Dassl.dassl.engine  # noqa: F401  (must come first: breaks the engine <-> trainer import cycle)

    _installed = True


def modules():
    """Return (trainer module, clip.model module, fed_utils module, metrics module) of the reference."""
    install()
    import trainers.GLP_OT_SVLoRA as T
    import clip.model as CM
    import utils.fed_utils as FU
    import evaluation.metrics as EM
    return T, CM, FU, EM


def make_cfg(modality="slo_fundus", ot="None", dataset="FairFedMed", n_prompts=2, n_ctx=4, eps=0.1, thresh=1e-3,
             max_iter=100, top_percent=0.8, dim_per_3d_slice=8):
    """Config object with exactly the fields CustomCLIP / PromptLearner read."""
    N = types.SimpleNamespace
    return N(
        INPUT=N(PIXEL_MEAN=[0.48145466, 0.4578275, 0.40821073], PIXEL_STD=[0.26862954, 0.26130258, 0.27577711],
                SIZE=(224, 224)),
        DATASET=N(NAME=dataset, MODALITY_TYPE=modality, DIM_PER_3D_SLICE=dim_per_3d_slice),
        TRAINER=N(GLP_OT=N(N_CTX=n_ctx, CTX_INIT="", CSC=False, N=n_prompts, CLASS_TOKEN_POSITION="end", EPS=eps,
                           THRESH=thresh, OT=ot, TOP_PERCENT=top_percent, MAX_ITER=max_iter, PREC="fp32")),
    )


def build_reference_model(cfg, classnames=("NOT Glaucoma", "Glaucoma"), rank=12, alpha=2, lora_type="FairLoRA",
                          num_attrs=3, global_s=False, vision_layers=12, vision_width=768, text_layers=12, seed=1):
    """Random-init CLIP ViT-B/16 (or a shrunken variant) + CustomCLIP + apply_lora_to_model, fp32, on CPU."""
    import torch
    T, CM, _, _ = modules()
    torch.manual_seed(seed)
    dd = {"trainer": "GLP_OT", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0}
    heads = max(1, 512 // 64)
    clip_model = CM.CLIP(512, 224, vision_layers, vision_width, 16, 77, 49408, 512, heads, text_layers, dd).float()
    model = T.CustomCLIP(cfg, list(classnames), clip_model)
    for name, p in model.named_parameters():
        p.requires_grad_("prompt_learner" in name or "proj_per_3d_slice" in name)
    T.apply_lora_to_model(model, True, rank=rank, alpha=alpha, lora_type=lora_type, global_s=global_s,
                          num_attrs=num_attrs)
    return model
