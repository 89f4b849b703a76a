"""The C-ABI library builds, loads and exports every symbol include/mdc.h declares.
No compute calls here (CPU-only box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from modulationdetectioncnn_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mdc.h")).read()
    return sorted(set(re.findall(r"MDC_API\s+[\w\s\*]+?\b(mdc_\w+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from modulationdetectioncnn_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 17
    assert sorted(_lib.EXPORTS) == syms
    for s in syms:
        assert hasattr(lib, s), f"libmdc.so does not export {s}"


def test_only_c_abi_is_exported():
    import subprocess
    from modulationdetectioncnn_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(exported) == declared_symbols()


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.mdc_version()
    assert isinstance(lib.mdc_last_error(), bytes)


def test_library_contains_sm100a_code():
    import subprocess
    from modulationdetectioncnn_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib):
    """Without a CUDA device, creating a handle must fail loudly (never silently compute on CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from modulationdetectioncnn_b200 import _lib
    with pytest.raises(_lib.MdcError):
        _lib.Handle(_lib.MODEL_TINY, 3, 3, _lib.MODE_FP32, 0)
    with pytest.raises(_lib.MdcError):
        from modulationdetectioncnn_b200.fwht import fwht
        import numpy as np
        fwht(np.zeros((1, 1024), dtype=np.int32))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "modulationdetectioncnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
