"""CPU tests: the C-ABI library loads, exports every symbol include/sba_b200.h declares and fails
loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from spherical_bundle_adjuster_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "sba_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sba_[a-z0-9_]+)\s*\(", hdr)) - {"sba_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sba_b200.h but not exported"
    assert sorted(_lib.EXPORTED) == declared


def test_version_and_error_text():
    lib = _lib.load()
    assert lib.sba_version() >= 100
    assert isinstance(lib.sba_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.sba_ctx_create(0, None, C.byref(h)) == -5           # SBA_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.sba_last_error()
    from spherical_bundle_adjuster_b200 import Context, SbaError
    with pytest.raises(SbaError):
        Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spherical_bundle_adjuster_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "sba_oracle" not in src and "libsba_ref" not in src, f
