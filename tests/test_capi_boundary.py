"""CPU (`not gpu`): the product library builds for sm_100a, loads, and exports every entry point that
include/mmego_b200.h declares; without a GPU it refuses to create a handle (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from mmego_b200 import ABI_VERSION, _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mmego_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmego_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_capi.SIGNATURES)


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (mmego_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing


def test_library_is_sm100a_native(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_loads_and_refuses_without_gpu(libpath):
    lib = _capi.Lib(libpath)
    assert lib.dll.mmego_abi_version() == ABI_VERSION
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hp = ctypes.c_void_p()
    rc = lib.dll.mmego_create(ctypes.byref(hp), 0)
    assert rc != 0 and not hp.value
    assert lib.dll.mmego_last_error(None)
    with pytest.raises(_capi.MMEgoError):
        _capi.Handle()


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(_capi.MMEgoError, match="no CPU fallback"):
        _capi.Lib(str(tmp_path / "libmmego_b200.so"))


def test_product_never_imports_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "mmego_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                s = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M) or "cuda_emul" in s and f != "cuda_compat.h":
                    bad.append(f)
    assert not bad, bad
