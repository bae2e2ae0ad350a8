"""Next row N1: threshold + contiguous-segment extraction on the kernel's output."""
import json
import math

import numpy
import pytest

from conftest import GOLDEN
from oracle import refine_oracle


def case_arrays(case):
    ids, prob, ann, ptr, cids = [], [], [], [0], []
    for contig in sorted(case["contigs"], key=lambda c: c["id"]):  # refine sorts by source id (refine.py:193)
        for g in contig["genes"]:
            ids.append(g["id"])
            prob.append(numpy.nan if g["p"] is None else g["p"])
            ann.append(g["annotated"])
        ptr.append(len(ids))
        cids.append(contig["id"])
    return ids, numpy.array(ptr, dtype=numpy.int32), numpy.array(prob), numpy.array(ann, dtype=numpy.uint8), cids


@pytest.fixture(scope="module")
def refine_cases():
    return json.loads((GOLDEN / "refine_cases.json").read_text())["cases"]


def check_against_reference(case, clusters, key, ids, cids):
    """clusters: [(contig, begin, end, ordinal, average_p, max_p)] vs what the reference class produced."""
    want = case[key]
    assert [[ids[g] for g in range(b, e)] for _, b, e, *_ in clusters] == [cl["genes"] for cl in want], case["settings"]
    assert [f"{cids[c]}_cluster_{o}" for c, _, _, o, _, _ in clusters] == [cl["id"] for cl in want]
    for (_, _, _, _, avg, mx), cl in zip(clusters, want):
        if cl["average_p"] is None:
            assert math.isnan(avg) and math.isnan(mx)
        else:
            assert abs(avg - cl["average_p"]) <= 1e-12 and mx == cl["max_p"]


def test_array_oracle_matches_the_reference_class(refine_cases):
    assert sum(len(c["clusters"]) for c in refine_cases) >= 50
    assert sum(c["clusters"] != c["clusters_per_contig_call"] for c in refine_cases) >= 2  # the leak is exercised
    for case in refine_cases:
        ids, ptr, prob, ann, cids = case_arrays(case)
        for key, reset in (("clusters", False), ("clusters_per_contig_call", True)):
            got = refine_oracle.extract_clusters(ptr, prob, ann, reset_per_contig=reset, **case["settings"])
            check_against_reference(case, got, key, ids, cids)


def device_clusters(engine, ptr, prob, ann, **kw):
    seg = engine.segments(ptr, prob, ann, **kw)
    return list(zip(seg.contig.tolist(), seg.begin.tolist(), seg.end.tolist(), seg.ordinal.tolist(),
                    seg.average_p.tolist(), seg.max_p.tolist()))


@pytest.mark.gpu
def test_device_segments_match_the_reference_class(engine, refine_cases):
    for case in refine_cases:
        ids, ptr, prob, ann, cids = case_arrays(case)
        for key, reset in (("clusters", False), ("clusters_per_contig_call", True)):
            got = device_clusters(engine, ptr, prob, ann, reset_per_contig=reset, **case["settings"])
            check_against_reference(case, got, key, ids, cids)
        # float32 probabilities: same segments whenever no probability sits within rounding of the threshold
        got32 = engine.extract_segments(ptr, prob.astype(numpy.float32), ann, **case["settings"])
        assert got32 == refine_oracle.extract_segments(ptr, prob.astype(numpy.float32).astype(numpy.float64), ann,
                                                       **case["settings"])


def random_table(seed, contigs, mean_len, nan_rate=0.02):
    rng = numpy.random.default_rng(seed)
    lens = numpy.maximum(1, rng.poisson(mean_len, size=contigs))
    ptr = numpy.concatenate([[0], numpy.cumsum(lens)]).astype(numpy.int32)
    G = int(ptr[-1])
    walk = numpy.cumsum(rng.normal(0, 0.35, size=G))
    k = min(200, G)  # mode="same" returns max(len) samples: keep the kernel no longer than the walk
    prob = 1 / (1 + numpy.exp(-(walk - numpy.convolve(walk, numpy.ones(k) / k, mode="same")) * 2.5))
    prob[rng.random(G) < nan_rate] = numpy.nan
    if contigs > 9:
        prob[ptr[7]:ptr[9]] = numpy.nan  # whole contigs without a probability
    ann = (rng.random(G) < 0.7).astype(numpy.uint8)
    return ptr, prob, ann


SETTINGS = (dict(threshold=0.8, n_cds=3, edge_distance=0, trim=True), dict(threshold=0.6, n_cds=5, edge_distance=3, trim=True),
            dict(threshold=0.7, n_cds=2, edge_distance=1, trim=False), dict(threshold=0.5, n_cds=1, edge_distance=40, trim=True))


@pytest.mark.gpu
@pytest.mark.parametrize("contigs,mean_len", [(3000, 60), (40000, 3), (3, 70000), (1, 1), (1, 5000)])
def test_device_segments_on_random_tables(engine, contigs, mean_len):
    """Many chunks per scan (G up to 210k), tiny contigs (boundary handling), one-gene table."""
    ptr, prob, ann = random_table(5 + contigs, contigs, mean_len)
    for kw in SETTINGS:
        for reset in (False, True):
            want = refine_oracle.extract_clusters(ptr, prob, ann, reset_per_contig=reset, **kw)
            got = device_clusters(engine, ptr, prob, ann, reset_per_contig=reset, **kw)
            assert len(got) == len(want)
            assert [g[:4] for g in got] == [w[:4] for w in want], kw
            if want:
                assert numpy.allclose([g[4:] for g in got], [w[4:] for w in want], rtol=0, atol=1e-12, equal_nan=True)


@pytest.mark.gpu
def test_device_segments_capacity_and_errors(engine):
    from gecco_b200._lib import GcrfError

    ptr, prob, ann = random_table(11, 500, 80)
    kw = dict(threshold=0.5, n_cds=1, edge_distance=0, trim=False)
    want = refine_oracle.extract_segments(ptr, prob, ann, **kw)
    assert len(want) > 8
    assert engine.extract_segments(ptr, prob, ann, capacity=4, **kw) == want  # retried with room for all
    assert engine.extract_segments(ptr[:1], prob[:0], ann[:0], **kw) == []    # empty table
    with pytest.raises(GcrfError):
        engine.extract_segments(ptr, prob, ann, threshold=0.5, n_cds=1, edge_distance=-1)
    with pytest.raises(GcrfError):
        engine.extract_segments(numpy.array([0, 5, 5, len(prob)]), prob, ann, **kw)  # empty contig


# ---------------------------------------------------------------------------------------------- drop-in class
def genes_of_refine_case(case, seed=3):
    import random

    from fake_model import Domain, Gene, Protein, Source

    genes = []
    for contig in case["contigs"]:
        src = Source(contig["id"])
        for i, g in enumerate(contig["genes"]):
            doms = [Domain("PF00001", 1, 10)] if g["annotated"] else []
            genes.append(Gene(src, 100 + 1000 * i, 900 + 1000 * i, 1, Protein(g["id"], None, doms), g["p"]))
    random.Random(seed).shuffle(genes)
    return genes


class OracleSegmentsEngine:
    """Answers ``segments`` from the CPU oracle: lets the host side of the drop-in class be tested without a GPU."""

    def segments(self, contig_ptr, prob, annotated, **kw):
        from gecco_b200._lib import Segments

        rows = refine_oracle.extract_clusters(contig_ptr, prob, annotated, **kw)
        seg = Segments(len(rows)).truncated(len(rows))
        for i, (c, b, e, o, avg, mx) in enumerate(rows):
            seg.contig[i], seg.begin[i], seg.end[i], seg.ordinal[i], seg.average_p[i], seg.max_p[i] = c, b, e, o, avg, mx
        return seg


def check_dropin(refiner_of, refine_cases):
    from gecco_b200.refine import ClusterRefiner

    for case in refine_cases:
        refiner = ClusterRefiner(criterion="gecco", **case["settings"], **refiner_of)
        clusters = list(refiner.iter_clusters(genes_of_refine_case(case)))
        assert [cl.id for cl in clusters] == [cl["id"] for cl in case["clusters"]]
        assert [[g.id for g in cl.genes] for cl in clusters] == [cl["genes"] for cl in case["clusters"]]


def test_dropin_refiner_host_side(refine_cases):
    check_dropin({"engine": OracleSegmentsEngine()}, refine_cases)
    from gecco_b200.refine import ClusterRefiner

    assert list(ClusterRefiner(engine=OracleSegmentsEngine()).iter_clusters([])) == []
    with pytest.raises(ValueError):
        list(ClusterRefiner(criterion="nope", engine=OracleSegmentsEngine()).iter_clusters([]))


@pytest.mark.gpu
def test_dropin_refiner_on_device(refine_cases):
    check_dropin({}, refine_cases)


# ------------------------------------------------------------------------------ criterion "antismash" (:157-163)
@pytest.fixture(scope="module")
def antismash_cases():
    return json.loads((GOLDEN / "refine_antismash_cases.json").read_text())["cases"]


def genes_of_antismash_case(case, seed=5):
    import random

    from fake_model import Domain, Gene, Protein, Source

    genes = []
    for contig in case["contigs"]:
        src = Source(contig["id"])
        for i, g in enumerate(contig["genes"]):
            doms = [Domain(name, 1 + 10 * j, 9 + 10 * j) for j, name in enumerate(g["domains"])]
            genes.append(Gene(src, 100 + 1000 * i, 900 + 1000 * i, 1, Protein(g["id"], None, doms), g["p"]))
    random.Random(seed).shuffle(genes)
    return genes


def check_antismash(refiner_of, cases):
    """Clusters of the reference's own class with ``criterion="antismash"`` (tools/make_golden.py antismash)."""
    from gecco_b200.refine import BIO_PFAMS, ClusterRefiner

    assert len(BIO_PFAMS) == 130 and "PF00109" in BIO_PFAMS
    assert sum(len(c["clusters"]) for c in cases) >= 40
    for case in cases:
        refiner = ClusterRefiner(criterion="antismash", **case["settings"], **refiner_of)
        clusters = list(refiner.iter_clusters(genes_of_antismash_case(case)))
        assert [cl.id for cl in clusters] == [cl["id"] for cl in case["clusters"]], case["settings"]
        assert [[g.id for g in cl.genes] for cl in clusters] == [cl["genes"] for cl in case["clusters"]]


def test_antismash_criterion_host_side(antismash_cases):
    check_antismash({"engine": OracleSegmentsEngine()}, antismash_cases)


@pytest.mark.gpu
def test_antismash_criterion_on_device(antismash_cases):
    check_antismash({}, antismash_cases)
