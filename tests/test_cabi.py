"""The C-ABI library loads on a CPU-only box and exports every symbol include/modelcompose_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "modelcompose_b200.h")).read()
    return sorted(set(re.findall(r"MC_API\s+[A-Za-z_0-9\*\s]+?\b(mc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from modelcompose_b200 import build
    return build.build()


def test_header_symbols_exported(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 9
    handle = ctypes.CDLL(built_lib)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in the header but not exported"


def test_ctypes_table_matches_header(built_lib):
    from modelcompose_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_symbols()
    lib = _cabi.lib()
    assert lib.mc_abi_version() == 3


def test_argument_validation_without_gpu(built_lib):
    """Pure host-side argument checks (no compute, no device needed)."""
    from modelcompose_b200 import _cabi
    lib = _cabi.lib()
    plan = ctypes.c_void_p()
    rc = lib.mc_merge_plan_create(ctypes.byref(plan), 1, 9, None, None, None, _cabi.MC_BF16, _cabi.MC_BF16, 0)
    assert rc == -1 and b"n_src" in lib.mc_last_error()
    rc = lib.mc_merge_plan_create(ctypes.byref(plan), 1, 2, None, None, None, 7, _cabi.MC_BF16, 0)
    assert rc == -1 and b"dtype" in lib.mc_last_error()
    assert lib.mc_merge_plan_run(None, None, 0, None) == -1
    assert lib.mc_merge_plan_destroy(None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from modelcompose_b200 import _cabi
    monkeypatch.setattr(_cabi, "_LIB", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.McError, match="no CPU or PyTorch fallback"):
        _cabi.lib()
