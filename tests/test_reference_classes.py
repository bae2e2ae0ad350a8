"""The drop-in meets the reference's REAL classes: ``gecco.model.Gene / Protein / Domain`` objects (frozen dataclasses
with copy-on-write ``with_*`` helpers, ``gecco/model.py:110-387``) go through ``gecco_b200.crf.ClusterCRF`` and through
the reference's own caller ``gecco.cli.commands._common.predict_probabilities(..., crf_type=...)``
(``gecco/cli/commands/_common.py:565-592``), and the result is compared object by object with what the reference's
``ClusterCRF.predict_probabilities`` returns for the same genes (its third-party tagger answered by the CPU oracle).

Needs the reference checkout (mounted in the build container, absent on the GPU box — skipped there); Biopython is
replaced by the import stubs of tools/make_golden.py.  The device call is answered by the oracle here: this file pins
the host side of the boundary against the real types; the CUDA path behind the same call is pinned in the gpu tests.
"""
import io
import pathlib
import sys
import warnings

import numpy
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
REFERENCE = pathlib.Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REFERENCE / "gecco" / "model.py").exists(), reason="reference checkout not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, str(ROOT / "tools"))
    import make_golden

    crfmod, modelmod, SeqRecord = make_golden.import_reference_crf(REFERENCE)
    return make_golden, crfmod, modelmod, SeqRecord


def real_genes(modelmod, SeqRecord, contigs):
    genes = []
    for cid, cgenes in contigs:
        src = SeqRecord(id=cid)
        for gid, start, doms in cgenes:
            domains = [modelmod.Domain(n, s, s + 10, "Pfam", 1e-20, 1e-20) for n, s in doms]
            genes.append(modelmod.Gene(src, start, start + 99, modelmod.Strand.Coding, modelmod.Protein(gid, None, domains)))
    return genes


def contigs_of(weights, seed=3):
    rng = numpy.random.default_rng(seed)
    names = list(weights.attrs)
    contigs = []
    for c, n in enumerate([45, 7, 1, 20, 63]):
        cgenes = []
        for i in range(n):
            k = int(rng.choice([0, 0, 1, 2, 3, 6]))
            doms = [(str(rng.choice(names)) if rng.random() < 0.9 else "PF99999", int(rng.integers(1, 300))) for _ in range(k)]
            if k >= 2 and rng.random() < 0.3:
                doms.append(doms[0])  # the same domain twice in one gene: features are a set (features.py:32)
            cgenes.append((f"c{c}_{i + 1}", 1000 * (n - i), doms))  # given in reverse start order
        contigs.append((f"contig{4 - c}", cgenes))
    return contigs


def ours(weights, window=None, step=None):
    from gecco_b200.crf import ClusterCRF
    from test_crf_dropin import OracleEngine

    class OracleBackedClusterCRF(ClusterCRF):
        def _get_engine(self):
            if self._engine is None:
                self._engine = OracleEngine(self._weights)
            return self._engine

    return OracleBackedClusterCRF


@pytest.mark.parametrize("pad", [True, False])
def test_real_gene_objects_through_the_dropin(ref, weights, pad):
    make_golden, crfmod, modelmod, SeqRecord = ref
    contigs = contigs_of(weights)
    want, want_warnings = make_golden.reference_predict(crfmod, modelmod, SeqRecord, weights, contigs, pad=pad)
    genes = real_genes(modelmod, SeqRecord, contigs)
    crf = ours(weights).trained()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        got = crf.predict_probabilities(genes, pad=pad)
    assert sorted(str(w.message) for w in caught) == sorted(want_warnings)
    assert all(type(g) is modelmod.Gene and type(g.protein) is modelmod.Protein for g in got)
    assert [g.id for g in got] == [g.id for g in want]
    for a, b in zip(got, want):
        assert a.source.id == b.source.id and (a.start, a.end, a.strand) == (b.start, b.end, b.strand)
        if b.average_probability is None:
            assert a.average_probability is None
        else:
            assert abs(a.average_probability - b.average_probability) <= 1e-12
            assert a.maximum_probability == a.average_probability
        assert [type(d) for d in a.protein.domains] == [modelmod.Domain] * len(b.protein.domains)
        for da, db in zip(a.protein.domains, b.protein.domains):
            assert (da.name, da.start, da.end, da.hmm, da.i_evalue, da.pvalue) == (db.name, db.start, db.end, db.hmm, db.i_evalue, db.pvalue)
            assert da.cluster_weight == db.cluster_weight
            assert (da.probability is None) == (db.probability is None)
            if db.probability is not None:
                assert abs(da.probability - db.probability) <= 1e-12
    # the caller's objects are not replaced; only their domain lists got sorted in place (:200-201)
    assert all(g.average_probability is None for g in genes)


def test_through_the_reference_pipeline_step(ref, weights, bgc):
    """``_common.predict_probabilities`` as ``gecco run`` / ``gecco predict`` call it, with ``crf_type`` = the drop-in, on the
    reference's CLI fixture: python-crfsuite's golden probabilities come back on real Gene objects."""
    make_golden, crfmod, modelmod, SeqRecord = ref
    import gecco.cli.commands._common as common
    from gecco.cli._log import make_logger
    from rich.console import Console

    doms = {}
    for d in bgc["domains"]:
        if d["pvalue"] < 1e-9:
            doms.setdefault(d["protein_id"], []).append((d["domain"], d["domain_start"]))
    contigs = [("BGC0001866.1", [(g["protein_id"], g["start"], doms.get(g["protein_id"], [])) for g in bgc["genes"]])]
    genes = real_genes(modelmod, SeqRecord, contigs)
    logger = make_logger(Console(file=io.StringIO()), quiet=0, verbose=0)
    out = common.predict_probabilities(logger, genes, model=None, pad=True, crf_type=ours(weights))
    golden = {g["protein_id"]: g["average_p"] for g in bgc["genes"]}
    assert [g.id for g in out] == [g["protein_id"] for g in bgc["genes"]]
    assert all(type(g) is modelmod.Gene for g in out)
    assert [g.average_probability for g in out] == [golden[g.id] for g in out]  # the oracle is bit-exact on this fixture
    sf = ours(weights).trained().model.state_features_
    assert all(d.cluster_weight == sf.get((d.name, "1")) for g in out for d in g.protein.domains)
    # the next pipeline step consumes them: the reference's own refiner finds its one cluster (tests/test_cli/test_run.py:68-70)
    from gecco.refine import ClusterRefiner

    clusters = list(ClusterRefiner(threshold=0.8, n_cds=3).iter_clusters(out))
    assert len(clusters) == 1
