"""The double-double exponential of the f64 reference-order kernel (gecco_b200/csrc/gcrf_exp.cuh), built as HOST code
with g++ (the header compiles both ways) and checked against 60-digit decimals: it must return the correctly rounded
double.  No GPU needed; the device build of the same source is checked in tests/test_gpu_f64.py."""
import pathlib
import random
import shutil
import subprocess
from decimal import Decimal, getcontext

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = tmp_path_factory.mktemp("expdd") / "exp_dd_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), str(ROOT / "tools" / "exp_dd_check.cpp")],
                   check=True)
    return exe


def test_exp_dd_is_correctly_rounded(checker):
    getcontext().prec = 60
    rng = random.Random(7)
    xs = [rng.uniform(-45, 45) for _ in range(6000)] + [rng.uniform(-690, 690) for _ in range(1500)]
    xs += [0.0, 1.0, -1.0, 1e-10, -1e-300, 12.649, -6.296, 2.670, -2.602, 88.7, -87.3, 0.5 ** 40]
    out = subprocess.run([str(checker)], input="\n".join(x.hex() for x in xs) + "\n", capture_output=True, text=True,
                         check=True).stdout.split()
    ours = [float.fromhex(t) for t in out[0::2]]
    libm = [float.fromhex(t) for t in out[1::2]]
    exact = [float(Decimal(x).exp()) for x in xs]  # Decimal -> float rounds correctly
    assert ours == exact
    # context for the parity tests: the host libm is close to, but not always, correctly rounded
    assert sum(a != b for a, b in zip(libm, exact)) < len(xs) // 100


def test_tables_are_current():
    """gcrf_exp_tables.inc is what tools/gen_exp_tables.py prints."""
    import sys

    made = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_exp_tables.py")], capture_output=True, text=True, check=True).stdout
    assert made == (ROOT / "gecco_b200" / "csrc" / "gcrf_exp_tables.inc").read_text()
