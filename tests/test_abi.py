"""The C-ABI library loads and exports every symbol include/resvg_b200.h declares (no compute, no GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "resvg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from resvg_b200 import _ffi

    syms = _declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(_ffi.lib, s)]
    assert not missing, missing
    unbound = [s for s in syms if s not in _ffi.SIGNATURES]
    assert not unbound, f"declared in the header but not bound in _ffi.SIGNATURES: {unbound}"


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    import pytest
    import resvg_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(resvg_b200.ResvgB200Error):
        resvg_b200.Context(0)


def test_product_never_references_oracle():
    """The product tree must not import, link or mention the oracle library."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "resvg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                if "oracle" in open(os.path.join(dirpath, f), errors="ignore").read().lower().replace("oracle restatement", ""):
                    bad.append(f)
    assert not bad, bad
