import json
import pathlib
import sys

import numpy
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def weights():
    from gecco_b200 import model_io

    return model_io.load_tsv_model(model_io.bundled_model_dir())


@pytest.fixture(scope="session")
def bgc():
    """The reference's CLI fixture BGC0001866 with its python-crfsuite golden probabilities."""
    return json.loads((GOLDEN / "bgc0001866.json").read_text())


@pytest.fixture(scope="session")
def mibig():
    return numpy.load(GOLDEN / "mibig_proG2.npz")


@pytest.fixture(scope="session")
def ref_cases():
    return json.loads((GOLDEN / "ref_loop_cases.json").read_text())["cases"]


def pack_case(case, weights):
    """CSR arrays of one ref_loop_cases.json case in the reference's order (contig id, gene start)."""
    from gecco_b200.packer import pack_records

    records = []
    for contig in case["contigs"]:
        for gene in contig["genes"]:
            doms = sorted(gene["domains"], key=lambda d: d[1])
            records.append((contig["id"], gene["start"], gene["id"], [d[0] for d in doms]))
    return pack_records(records, weights.attr_index)


@pytest.fixture(scope="session")
def engine(weights):
    from gecco_b200._lib import CRFEngine

    eng = CRFEngine(weights, device=0)
    yield eng
    eng.close()
