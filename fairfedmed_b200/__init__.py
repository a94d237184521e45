"""fairfedmed_b200 — B200-native (sm_100a) implementation of the FairLoRA training hot path of
Harvard-AI-and-Robotics-Lab/FairFedMed.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of libffm_b200.so), `_cabi` (ctypes
binding), `ops` (torch custom ops + autograd), and the host-side mirrors of the reference interfaces:
`modules` (FairLoRALinear & co), `clip_model` (attr-aware CLIP ViT + CustomCLIP), `trainer`
(TRAINER_REGISTRY entry GLP_OT_SVLoRA), `fed_utils` / `federated` (aggregation, round loop), `metrics`.
"""
__version__ = "0.2.0"

from . import _cabi  # noqa: F401


def build():
    """Compile libffm_b200.so in-tree (nvcc, sm_100a)."""
    from .build import build as _build
    return _build()
