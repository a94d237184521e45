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
STAGED_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference.zip")

_installed = False


def _unpack_staged() -> bool:
    """No reference tree (GPU box): unpack oracle/_ref/reference.zip (oracle/stage_ref.py) into a scratch directory."""
    global REF_ROOT
    if not os.path.isfile(STAGED_ZIP):
        return False
    import hashlib
    import tempfile
    import zipfile
    tag = hashlib.sha1(f"{os.path.getsize(STAGED_ZIP)}-{os.path.getmtime(STAGED_ZIP)}".encode()).hexdigest()[:12]
    dst = os.path.join(tempfile.gettempdir(), f"ffm_reference_{tag}")
    if not os.path.isdir(os.path.join(dst, "trainers")):
        tmp = dst + f".{os.getpid()}"
        with zipfile.ZipFile(STAGED_ZIP) as z:
            z.extractall(tmp)
        try:
            os.replace(tmp, dst)
        except OSError:                      # another process won the race
            pass
    REF_ROOT = dst
    return os.path.isdir(os.path.join(REF_ROOT, "trainers"))


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "trainers")) or _unpack_staged()


def kind() -> str:
    """'reference' when the real reference modules can be imported (tree or staged archive), else 'port'."""
    return "reference" if available() else "port"


def _module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _AttrDict(dict):
    """Stand-in for yacs.config.CfgNode (yacs is not installed here): attribute access, clone, and the subset of the
    yacs API `federated_main.setup_cfg` uses — merge_from_file (yaml, string values literal_eval'ed like yacs does),
    merge_from_list, freeze / defrost."""

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        dict.__setattr__(self, "_frozen", False)
        for k, v in (init_dict or {}).items():
            self[k] = _AttrDict(v) if isinstance(v, dict) and not isinstance(v, _AttrDict) else v

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        if dict.__getattribute__(self, "__dict__").get("_frozen", False):
            raise AttributeError(f"Attempted to set {key} to {value}, but CfgNode is immutable")
        self[key] = value

    def clone(self):
        import copy
        return copy.deepcopy(self)

    @staticmethod
    def _decode(v):
        if isinstance(v, str):
            import ast
            try:
                return ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        return v

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                node = self.get(k)
                if not isinstance(node, _AttrDict):
                    node = _AttrDict()
                    dict.__setitem__(self, k, node)
                node._merge(v)
            else:
                dict.__setitem__(self, k, self._decode(v))

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        opts = list(opts or [])
        assert len(opts) % 2 == 0
        for full_key, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = full_key.split(".")
            for part in parts[:-1]:
                node = node[part]
            dict.__setitem__(node, parts[-1], self._decode(v))

    def freeze(self):
        dict.__getattribute__(self, "__dict__")["_frozen"] = True
        for v in self.values():
            if isinstance(v, _AttrDict):
                v.freeze()

    def defrost(self):
        dict.__getattribute__(self, "__dict__")["_frozen"] = False
        for v in self.values():
            if isinstance(v, _AttrDict):
                v.defrost()

    def is_frozen(self):
        return dict.__getattribute__(self, "__dict__").get("_frozen", False)


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
