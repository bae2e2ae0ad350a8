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


def _build_example(tmp_path, name):
    import shutil
    import subprocess

    root = pathlib.Path(__file__).resolve().parent.parent
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / name
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{root / 'include'}", str(root / "examples" / f"{name}.c"),
                    f"-L{root / 'gecco_b200'}", "-lgecco_crf_b200", f"-Wl,-rpath,{root / 'gecco_b200'}", "-lm", "-o", str(exe)],
                   check=True, capture_output=True)
    return exe


def test_device_example_compiles_as_c99_and_refuses_without_a_gpu(tmp_path):
    """examples/marginals_device.c: the device half of the ABI from plain C.  On a box without a B200 the library says
    GCRF_ENODEVICE (exit code 3) — it has no CPU path to fall back to."""
    import subprocess

    import torch

    exe = _build_example(tmp_path, "marginals_device")
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert run.returncode == 3 and "no usable B200" in run.stdout


@pytest.mark.gpu
def test_device_example_runs_on_the_gpu(tmp_path):
    """The same program on a B200, checked against the oracle: FP32 arithmetic within 1e-5, f64 within 1e-12, the wire
    format equal to the CSR call."""
    import subprocess

    import numpy
    from oracle import crf_oracle

    exe = _build_example(tmp_path, "marginals_device")
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "wire_equals_csr=1" in run.stdout
    rows = [line.split() for line in run.stdout.splitlines() if line.startswith("gene ")]
    p32 = numpy.array([float(r[2][2:]) for r in rows])
    p64 = numpy.array([float(r[3][4:]) for r in rows])
    state_w = numpy.array([[0.5, -0.25], [-1.0, 2.0], [0.0, 0.75]])
    trans_w = numpy.array([[2.5, -2.5], [-2.5, 2.5]])
    want, _ = crf_oracle.marginals_windowed(state_w, trans_w, 1, numpy.array([0, 3, 9]), numpy.array([0, 1, 3, 3, 4, 4, 6, 7, 9, 10]),
                                            numpy.array([1, 0, 2, 1, 2, 1, -1, 0, 1, 2], dtype=numpy.int32), 5, 1, True)
    assert numpy.abs(p32 - want).max() <= 1e-5 and numpy.abs(p64 - want).max() <= 1e-12
