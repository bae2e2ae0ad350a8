"""Next row N1: threshold + contiguous-segment extraction on the kernel's output."""
import json

import numpy
import pytest

from conftest import GOLDEN
from oracle import refine_oracle


def case_arrays(case):
    ids, prob, ann, ptr = [], [], [], [0]
    for contig in sorted(case["contigs"], key=lambda c: c["id"]):  # refine sorts by source id (refine.py:193)
        for g in contig["genes"]:
            ids.append(g["id"])
            prob.append(numpy.nan if g["p"] is None else g["p"])
            ann.append(g["annotated"])
        ptr.append(len(ids))
    return ids, numpy.array(ptr, dtype=numpy.int32), numpy.array(prob), numpy.array(ann, dtype=numpy.uint8)


@pytest.fixture(scope="module")
def refine_cases():
    return json.loads((GOLDEN / "refine_cases.json").read_text())["cases"]


def test_array_oracle_matches_the_reference_class(refine_cases):
    assert sum(len(c["clusters"]) for c in refine_cases) >= 50
    for case in refine_cases:
        ids, ptr, prob, ann = case_arrays(case)
        segs = refine_oracle.extract_segments(ptr, prob, ann, **case["settings"])
        got = [[ids[g] for g in range(b, e)] for _, b, e in segs]
        assert got == [cl["genes"] for cl in case["clusters"]], case["settings"]


@pytest.mark.gpu
def test_device_segments_match_the_reference_class(engine, refine_cases):
    for case in refine_cases:
        ids, ptr, prob, ann = case_arrays(case)
        segs = engine.extract_segments(ptr, prob, ann, **case["settings"])
        got = [[ids[g] for g in range(b, e)] for _, b, e in segs]
        assert got == [cl["genes"] for cl in case["clusters"]], case["settings"]


@pytest.mark.gpu
def test_device_segments_on_a_large_random_table(engine):
    rng = numpy.random.default_rng(5)
    lens = numpy.maximum(1, rng.poisson(60, size=3000))
    ptr = numpy.concatenate([[0], numpy.cumsum(lens)]).astype(numpy.int32)
    G = int(ptr[-1])
    walk = numpy.cumsum(rng.normal(0, 0.35, size=G))
    prob = 1 / (1 + numpy.exp(-(walk - numpy.convolve(walk, numpy.ones(200) / 200, mode="same")) * 2.5))
    prob[rng.random(G) < 0.02] = numpy.nan
    prob[ptr[7]:ptr[9]] = numpy.nan  # whole contigs without a probability
    ann = (rng.random(G) < 0.7).astype(numpy.uint8)
    for kw in (dict(threshold=0.8, n_cds=3, edge_distance=0, trim=True), dict(threshold=0.6, n_cds=5, edge_distance=3, trim=True),
               dict(threshold=0.7, n_cds=2, edge_distance=1, trim=False)):
        want = refine_oracle.extract_segments(ptr, prob, ann, **kw)
        got = engine.extract_segments(ptr, prob, ann, **kw)
        assert len(want) > 20
        assert got == want, kw
