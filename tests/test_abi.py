"""The C-ABI library loads and exports every symbol include/gecco_crf_b200.h declares (no GPU needed)."""
import ctypes
import pathlib
import re

import numpy
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "gecco_crf_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcrf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_the_header():
    from gecco_b200 import _lib

    lib = _lib.load_library()
    declared = header_symbols()
    assert declared, "no declarations found in the header"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert lib.gcrf_version() == 1
    assert lib.gcrf_last_error() is not None


def test_no_cpu_fallback_without_a_device(weights):
    """On a box without a GPU the product path must fail loudly, not fall back."""
    from gecco_b200 import _lib

    lib = _lib.load_library()
    if lib.gcrf_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.GcrfError) as err:
        _lib.CRFEngine(weights)
    assert err.value.status == -2  # GCRF_ENODEVICE
    assert "no CPU path" in str(err.value)


def test_model_argument_validation_happens_before_device_use(weights):
    from gecco_b200 import _lib

    lib = _lib.load_library()
    handle = ctypes.c_void_p()
    w3 = numpy.zeros((4, 3))
    rc = lib.gcrf_model_create(w3.ctypes.data, 4, 3, numpy.zeros((3, 3)).ctypes.data, 1, 0, ctypes.byref(handle))
    assert rc == -5 and b"2-label" in lib.gcrf_last_error()
    rc = lib.gcrf_model_create(None, 4, 2, None, 1, 0, ctypes.byref(handle))
    assert rc == -1


def test_product_package_never_imports_the_oracle():
    """SPEC: only tests/, smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    for path in (ROOT / "gecco_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h", ".cpp"}:
            text = path.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
            assert "crf_oracle" not in text, path
