"""ctypes binding of libffm_b200.so (the C ABI declared in include/ffm_b200.h).

This is the stub a maintainer of the reference would add: raw device pointers, sizes and the current
CUDA stream go in, an int status comes out.  There is deliberately NO fallback — if the shared library
is missing or a call fails, a Python exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "lib" / "libffm_b200.so"
HEADER_PATH = Path(__file__).resolve().parent.parent / "include" / "ffm_b200.h"

_vp, _fp, _i, _f, _sz, _i64 = C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64

# name -> (restype, argtypes).  Pointers are passed as integers (tensor.data_ptr()).
SIGNATURES = {
    "ffm_last_error": (C.c_char_p, []),
    "ffm_version": (_i, []),
    "ffm_svlora_max_rank": (_i, []),
    "ffm_svlora_padded_rank": (_i, [_i]),
    "ffm_launch_count": (C.c_longlong, [_i]),
    "ffm_profile_enable": (_i, [_i]),
    "ffm_profile_read": (_i, [_vp, _vp, _i]),
    "ffm_svlora_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ffm_svlora_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ffm_svlora_prepare": (_i, [_fp, _fp, _fp, _vp, _sz, _i, _i, _i, _i, _f, _vp]),
    "ffm_svlora_fwd": (_i, [_vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _vp, _vp, _sz,
                            _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "ffm_svlora_bwd": (_i, [_vp, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp, _vp, _vp, _fp, _fp, _fp, _vp, _sz,
                            _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "ffm_svlora_bwd_phase": (_i, [_vp, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp, _vp, _vp, _fp, _fp, _fp, _vp, _sz,
                                  _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "ffm_frozen_linear": (_i, [_vp, _vp, _fp, _vp, _i, _i, _i, _vp]),
    "ffm_add_layernorm_fwd": (_i, [_vp, _vp, _fp, _fp, _vp, _vp, _fp, _fp, _i, _i, _f, _vp]),
    "ffm_add_layernorm_bwd": (_i, [_vp, _vp, _vp, _fp, _fp, _fp, _vp, _i, _i, _vp]),
    "ffm_attention_max_len": (_i, []),
    "ffm_attention_fwd": (_i, [_vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffm_attention_bwd": (_i, [_vp, _vp, _vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffm_patchify_normalize": (_i, [_fp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffm_vit_embed_ln": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _fp, _i, _i, _i, _f, _f, _vp]),
    "ffm_oct_minmax_patchify": (_i, [_fp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "ffm_oct_input_bwd_ws_bytes": (_sz, [_i]),
    "ffm_oct_input_bwd": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "ffm_lora_merged_weight_ws_bytes": (_sz, [_i, _i]),
    "ffm_lora_merged_weight": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _f, _vp]),
    "ffm_lora_merged_weight_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _sz, _i, _i, _i, _f, _vp]),
    "ffm_oct_slice_conv_wgrad_ws_bytes": (_sz, [_i]),
    "ffm_oct_slice_conv_fwd": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _f, _vp]),
    "ffm_oct_slice_conv_wgrad": (_i, [_fp, _fp, _fp, _fp, _vp, _sz, _i, _i, _i, _i, _i, _f, _vp]),
    "ffm_avgpool_nhwc_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffm_avgpool_nhwc_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ffm_widen_bf16": (_i, [_vp, _fp, _i64, _vp]),
    "ffm_bn_ws_bytes": (_sz, [_i]),
    "ffm_bn_relu_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _sz, _i64, _i, _f, _f, _i, _vp]),
    "ffm_bn_relu_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _sz, _i64, _i, _i, _vp]),
    "ffm_add_relu": (_i, [_fp, _fp, _fp, _i64, _vp]),
    "ffm_relu_mask": (_i, [_fp, _fp, _fp, _i64, _vp]),
    "ffm_seff": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _f, _vp]),
    "ffm_ds": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _f, _vp]),
    "ffm_ot_head_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ffm_ot_head_fwd": (_i, [_vp, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _vp, _sz,
                             _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _vp]),
    "ffm_ot_head_bwd": (_i, [_vp, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _vp, _sz,
                             _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ffm_sinkhorn_workspace_bytes": (_sz, [_i, _i, _i]),
    "ffm_sinkhorn": (_i, [_fp, _fp, _vp, _vp, _sz, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "ffm_fedavg_scale": (_i, [_fp, _fp, _vp, _vp, _vp, _i, _i64, _f, _fp, _i, _i, _vp]),
    "ffm_fedavg_epilogue": (_i, [_fp, _fp, _fp, _vp, _vp, _vp, _i, _i64, _f, _i, _i, _i, _vp]),
    "ffm_group_auc_workspace_bytes": (_sz, [_i, _i, _i]),
    "ffm_group_auc": (_i, [_fp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _vp]),
    "ffm_sgd_step": (_i, [_fp, _fp, _fp, _i64, _f, _f, _f, _i, _i, _vp]),
    "ffm_sgd_step_dev_lr": (_i, [_fp, _fp, _fp, _i64, _fp, _f, _f, _i, _vp]),
}


class FfmError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function name declared in include/ffm_b200.h."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffm_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load the shared library (once) and attach argument types. Raises if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise FfmError(
            f"{LIB_PATH} is missing: build it with `python -m fairfedmed_b200.build` "
            "(the CUDA library is the only compute path; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and os.environ.get("FFM_ALLOW_PARTIAL_LIB") != "1":  # bring-up escape hatch only
        raise FfmError(f"{LIB_PATH} is stale or incomplete, missing symbols: {missing}; rebuild it")
    _check_digest()
    _LIB = lib
    return lib


def _check_digest() -> None:
    """A library that still exports every symbol but was built from older sources must not be loaded silently:
    compare lib/.build_digest (written by build.py) with the digest of the sources next to it."""
    if os.environ.get("FFM_ALLOW_STALE_LIB") == "1":
        return
    import importlib
    _build = importlib.import_module(__package__ + ".build")   # the package attribute `build` is a function
    stamp = _build.LIB_DIR / ".build_digest"
    srcs = _build.sources()
    if not srcs:                      # binary-only deployment: nothing to compare with
        return
    deps = srcs + sorted(_build.CSRC.glob("*.cuh")) + [HEADER_PATH]
    want = _build._digest(deps)
    have = stamp.read_text() if stamp.exists() else None
    if have != want:
        raise FfmError(f"{LIB_PATH} was not built from the current sources (build digest mismatch): run "
                       "`python -m fairfedmed_b200.build` (FFM_ALLOW_STALE_LIB=1 overrides)")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ffm_last_error()
        raise FfmError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise FfmError on a non-zero status."""
    rc = getattr(load(), name)(*args)
    check(rc, name)
