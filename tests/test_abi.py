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


def test_cpp_facade_builds_and_fails_loudly_without_gpu(tmp_path):
    """The drop-in C++ classes link against the C ABI; with no GPU the demo must exit non-zero with the
    library's error text instead of computing anything on the CPU."""
    import subprocess
    import torch
    facade = os.path.join(ROOT, "spherical_bundle_adjuster_b200", "libsba_facade.so")
    demo = os.path.join(ROOT, "build", "facade_demo")
    assert os.path.exists(facade) and os.path.exists(demo), "run __graft_entry__.build()"
    out = subprocess.run(["nm", "-D", "--defined-only", facade], capture_output=True, text=True).stdout
    for sym in ["equi2cube7get_all", "feature_matcher15match_two_image", "equi2cube_surf6do_all", "equi2cube_surf15cube2equi_pixel",
                "spherical_bundle_adjuster20do_bundle_adjustment", "ba_spherical_costfunctor_rot_only5solve"]:
        assert sym in out, sym
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    open(tmp_path / "meta.txt", "w").write("8 4 2 0 0\n")
    (tmp_path / "im.bin").write_bytes(bytes(8 * 4 * 3))
    r = subprocess.run([demo, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_python_constants_mirror_the_header_enums():
    """_lib.py's algorithm / memory constants are the header's enum values (a silent mismatch would select the wrong kernel)."""
    import re
    from spherical_bundle_adjuster_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sba_b200.h")).read()
    def enum(name):
        m = re.search(name + r"\s*=\s*(-?\d+)", hdr)
        assert m, name
        return int(m.group(1))
    assert (_lib.MATCH_AUTO, _lib.MATCH_SIMT_EXACT, _lib.MATCH_TENSOR, _lib.MATCH_TENSOR_FP16) == \
        (enum("SBA_MATCH_AUTO"), enum("SBA_MATCH_SIMT_EXACT"), enum("SBA_MATCH_TENSOR"), enum("SBA_MATCH_TENSOR_FP16"))
    assert (_lib.SBA_MEM_HOST, _lib.SBA_MEM_DEVICE) == (enum("SBA_MEM_HOST"), enum("SBA_MEM_DEVICE"))
