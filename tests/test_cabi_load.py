"""The C-ABI library loads on a CPU-only box and exports exactly what include/ffm_b200.h declares."""
import ctypes

import pytest

from fairfedmed_b200 import _cabi
from fairfedmed_b200.build import build


@pytest.fixture(scope="module")
def lib():
    build()   # no-op when the in-tree .so is up to date; nvcc cross-compiles without a GPU
    return _cabi.load()


def test_header_symbols_are_exported(lib):
    raw = ctypes.CDLL(str(_cabi.LIB_PATH))
    declared = _cabi.declared_symbols()
    assert len(declared) >= 19
    missing = [s for s in declared if not hasattr(raw, s)]
    assert not missing, f"declared in include/ffm_b200.h but not exported: {missing}"


def test_binding_covers_header():
    declared = set(_cabi.declared_symbols())
    bound = set(_cabi.SIGNATURES)
    assert declared == bound, (declared - bound, bound - declared)


def test_binding_arity_and_kinds_match_header():
    """Every ctypes signature has the parameter count of the C declaration, pointers bind to pointers, ints to ints,
    floats to floats (an ABI drift here is silent memory corruption at run time)."""
    import re
    text = _cabi.HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = dict()
    for m in re.finditer(r"\b(?:int|size_t|long long|const char\*)\s+(ffm_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), " ".join(m.group(2).split())
        decls[name] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    assert set(decls) == set(_cabi.SIGNATURES)
    for name, params in decls.items():
        _, argtypes = _cabi.SIGNATURES[name]
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for p, t in zip(params, argtypes):
            if "*" in p or "ffm_stream_t" in p:
                assert t is ctypes.c_void_p, (name, p, t)
            elif p.startswith("float"):
                assert t is ctypes.c_float, (name, p, t)
            elif p.startswith("size_t"):
                assert t is ctypes.c_size_t, (name, p, t)
            elif p.startswith("int64_t"):
                assert t is ctypes.c_int64, (name, p, t)
            else:
                assert t is ctypes.c_int, (name, p, t)


def test_no_compute_metadata_calls(lib):
    assert lib.ffm_version() >= 100
    assert lib.ffm_svlora_max_rank() == 32
    assert [lib.ffm_svlora_padded_rank(r) for r in (0, 1, 12, 16, 17, 32, 33)] == [0, 16, 16, 16, 32, 32, 0]
    assert lib.ffm_svlora_fwd_workspace_bytes(1576, 768, 3072, 8) > 0
    assert lib.ffm_svlora_bwd_workspace_bytes(1576, 768, 3072, 8) > lib.ffm_svlora_fwd_workspace_bytes(1576, 768, 3072, 8)
    assert lib.ffm_ot_head_workspace_bytes(196, 64, 512, 2, 2) > 0
    assert lib.ffm_sinkhorn_workspace_bytes(128, 196, 2) > 0
    assert lib.ffm_group_auc_workspace_bytes(5000, 6, 3) > 0
    assert lib.ffm_attention_max_len() == 208


def test_argument_validation_without_gpu(lib):
    # null pointers are rejected before any CUDA call is made
    rc = lib.ffm_seff(0, 0, 0, 0, 1, 3, 12, 0.7, 0)
    assert rc == -22
    assert b"null pointer" in lib.ffm_last_error()


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from fairfedmed_b200.config import get_cfg_default
    from fairfedmed_b200.registry import build_trainer
    import fairfedmed_b200.trainer  # noqa: F401  (registers GLP_OT_SVLoRA)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        build_trainer(get_cfg_default())
    from fairfedmed_b200.modules import FairLoRALinear
    import torch.nn as nn
    layer = FairLoRALinear(nn.Linear(64, 64), rank=12, alpha=2.0, num_attrs=3)
    with pytest.raises(Exception):
        layer(torch.randn(4, 2, 64), torch.tensor([0, 1]))
