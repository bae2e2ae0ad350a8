"""The drop-in ClusterCRF: host logic on CPU (device call replaced by the oracle), the real thing on GPU."""
import warnings

import numpy
import pytest

from fake_model import Domain, Gene, Protein, Source, genes_of_case
from gecco_b200.crf import ClusterCRF, NotFittedError


class OracleEngine:
    """Test double for CRFEngine: answers ``marginals_windowed`` from the CPU oracle."""

    def __init__(self, weights):
        self.weights = weights
        self.calls = 0

    def max_window(self, f64_arith=False):
        return 2**31 - 1 if f64_arith else 128

    def marginals_windowed(self, contig_ptr, gene_ptr, attr_idx, *, window, step, pad, f64_arith=False):
        from oracle import crf_oracle

        self.calls += 1
        self.f64_arith = f64_arith
        p, _ = crf_oracle.marginals_windowed(self.weights.state_w, self.weights.trans_w, self.weights.label_id("1"),
                                             contig_ptr, gene_ptr, attr_idx, window, step, pad)
        return p


def make_crf(weights, case=None, oracle=True):
    crf = ClusterCRF.trained()
    if case is not None:
        crf.window_size, crf.window_step = case["window"], case["step"]
    if oracle:
        crf._engine = OracleEngine(weights)
    return crf


def check_case(crf, case, tol):
    genes = genes_of_case(case, shuffle_seed=5)
    calls = []
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        out = crf.predict_probabilities(genes, pad=case["pad"], progress=lambda i, n: calls.append((i, n)))
    expected = case["expected"]
    assert [g.id for g in out] == [e["id"] for e in expected]  # contig id, then start
    assert [g.source.id for g in out] == [e["contig"] for e in expected]
    for g, e in zip(out, expected):
        if e["p"] is None:
            assert g.average_probability is None
        else:
            assert abs(g.average_probability - e["p"]) <= tol
            assert all(d.probability == g.average_probability for d in g.protein.domains)
        assert [d.cluster_weight for d in g.protein.domains] == e["weights"]
        assert [d.start for d in g.protein.domains] == sorted(d.start for d in g.protein.domains)
    assert sorted(str(w.message) for w in caught) == sorted(case["warnings"])
    assert calls[0][0] == 0 and calls[-1] == (calls[0][1], calls[0][1])
    return out


def test_reference_cases_host_logic(ref_cases, weights):
    for case in ref_cases:
        check_case(make_crf(weights, case), case, tol=1e-12)


def test_trained_defaults_and_model_view(weights):
    crf = ClusterCRF.trained()
    assert (crf.feature_type, crf.window_size, crf.window_step) == ("protein", 20, 1)
    assert crf.model.state_features_[("PF00109", "1")] == 0.3281678016907602
    assert crf.model.transition_features_[("0", "1")] == -2.599571900486168
    assert len(crf.model.attributes_) == 2659 and crf.model.classes_ == ["0", "1"]


def test_constructor_and_unfitted_errors():
    with pytest.raises(ValueError, match="invalid feature type"):
        ClusterCRF("gene")
    with pytest.raises(ValueError, match="Window size must be strictly positive"):
        ClusterCRF(window_size=0)
    with pytest.raises(ValueError, match="Window step must be strictly positive"):
        ClusterCRF(window_size=5, window_step=6)
    with pytest.raises(NotFittedError):
        ClusterCRF().predict_probabilities([])
    assert issubclass(NotFittedError, ValueError)


def test_inputs_are_not_replaced_but_domains_get_sorted_in_place(weights, ref_cases):
    case = ref_cases[0]
    genes = genes_of_case(case)
    genes[0].protein.domains.reverse()
    before = [id(g) for g in genes]
    out = make_crf(weights, case).predict_probabilities(genes, pad=True)
    assert [id(g) for g in genes] == before and all(g._probability is None for g in genes)
    assert all(id(o) not in before for o in out)
    for g in genes:  # gecco/crf/__init__.py:200-201 mutates the caller's lists
        assert [d.start for d in g.protein.domains] == sorted(d.start for d in g.protein.domains)


def test_domain_feature_type_sizes_by_rows(weights):
    """One row per domain, one per domain-less gene (features.py:38-48); the reference breaks on this input."""
    from oracle import crf_oracle

    names = weights.attrs
    src = Source("ctg")
    genes = []
    layout = [2, 0, 1, 3, 0, 1, 1, 2, 0, 1, 1, 1]
    k = 0
    for i, nd in enumerate(layout):
        doms = [Domain(names[(7 * k + 3 * j) % len(names)], 10 * j + 1, 10 * j + 9) for j in range(nd)]
        k += nd
        genes.append(Gene(src, 100 * i, 100 * i + 90, 1, Protein(f"g{i}", None, doms)))
    crf = ClusterCRF.trained()
    crf.feature_type = "domain"
    crf.window_size = 5
    crf._engine = OracleEngine(weights)
    out = crf.predict_probabilities(genes)
    rows = []
    for g in genes:
        rows += [[weights.attr_index[d.name]] for d in g.protein.domains] or [[]]
    ptr = numpy.cumsum([0] + [len(r) for r in rows])
    idx = numpy.array([a for r in rows for a in r], dtype=numpy.int32)
    want, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, numpy.array([0, len(rows)]), ptr, idx, 5, 1, True)
    got = []
    for g in out:
        got += [d.probability for d in g.protein.domains] or [g.average_probability]
    assert numpy.abs(numpy.array(got) - want).max() < 1e-12


def test_training_is_delegated_or_refused(weights, tmp_path):
    crf = ClusterCRF.trained()
    crf.save(tmp_path)  # weight tables when not fitted by the reference class
    again = ClusterCRF.trained(tmp_path)
    assert numpy.array_equal(again._weights.state_w, weights.state_w) and again.window_size == 20
    try:
        import sklearn_crfsuite  # noqa: F401
    except ImportError:
        with pytest.raises(NotImplementedError, match="training is not part"):
            ClusterCRF().fit([])


def test_arithmetic_switch(weights, ref_cases, monkeypatch):
    """The drop-in asks for the reference's f64 arithmetic unless told otherwise (GECCO_B200_ARITHMETIC / .arithmetic)."""
    crf = make_crf(weights, ref_cases[0])
    assert crf.arithmetic == "f64"
    crf.predict_probabilities(genes_of_case(ref_cases[0]), pad=True)
    assert crf._engine.f64_arith is True
    monkeypatch.setenv("GECCO_B200_ARITHMETIC", "f32")
    crf = make_crf(weights, ref_cases[0])
    crf.predict_probabilities(genes_of_case(ref_cases[0]), pad=True)
    assert crf.arithmetic == "f32" and crf._engine.f64_arith is False
    monkeypatch.setenv("GECCO_B200_ARITHMETIC", "f16")
    with pytest.raises(ValueError, match="invalid arithmetic"):
        ClusterCRF()


@pytest.mark.gpu
@pytest.mark.parametrize("arithmetic,tol", [("f32", 1e-5), ("f64", 1e-12)])
def test_reference_cases_on_device(ref_cases, weights, arithmetic, tol):
    for case in ref_cases:
        crf = make_crf(weights, case, oracle=False)
        crf.arithmetic = arithmetic
        check_case(crf, case, tol=tol)


@pytest.mark.gpu
@pytest.mark.parametrize("arithmetic,tol", [("f32", 1e-5), ("f64", 1e-12)])
def test_domain_feature_type_on_device_vs_oracle(weights, arithmetic, tol):
    """feature_type="domain" (gecco/crf/features.py:38-48, 99-120) through the CUDA path against the CPU oracle: one
    row per domain, one empty row per domain-less gene, several contigs incl. one shorter than the window."""
    from oracle import crf_oracle

    rng = numpy.random.default_rng(11)
    names = weights.attrs
    genes, rows, contig_rows = [], [], []
    for c, n_genes in enumerate([40, 3, 75, 1, 12]):
        src = Source(f"ctg{c}")
        before = len(rows)
        for i in range(n_genes):
            nd = int(rng.choice([0, 1, 1, 2, 4]))
            doms = [Domain(str(rng.choice(names)) if rng.random() < 0.9 else "PF99999", 10 * j + 1, 10 * j + 9) for j in range(nd)]
            genes.append(Gene(src, 100 * i, 100 * i + 90, 1, Protein(f"c{c}g{i}", None, doms)))
            rows += [[weights.attr_index.get(d.name, -1)] for d in doms] or [[]]
        contig_rows.append(len(rows) - before)
    crf = ClusterCRF.trained()
    crf.feature_type, crf.window_size, crf.arithmetic = "domain", 5, arithmetic
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = crf.predict_probabilities(genes)
    ptr = numpy.cumsum([0] + [len(r) for r in rows])
    idx = numpy.array([a for r in rows for a in r], dtype=numpy.int32)
    cptr = numpy.cumsum([0] + contig_rows)
    want, _ = crf_oracle.marginals_windowed(weights.state_w, weights.trans_w, 1, cptr, ptr, idx, 5, 1, True)
    got = []
    for g in out:
        got += [d.probability for d in g.protein.domains] or [g.average_probability]
    assert len(got) == len(want) and numpy.abs(numpy.array(got) - want).max() <= tol
    # bulk cluster-weight annotation (gecco/crf/__init__.py:261-269)
    sf = crf.model.state_features_
    assert all(d.cluster_weight == sf.get((d.name, "1")) for g in out for d in g.protein.domains)


@pytest.mark.gpu
def test_golden_bgc0001866_through_the_dropin(bgc, weights):
    doms = {}
    for d in bgc["domains"]:
        if d["pvalue"] < 1e-9:
            doms.setdefault(d["protein_id"], []).append(Domain(d["domain"], d["domain_start"], d["domain_end"]))
    src = Source("BGC0001866.1")
    genes = [Gene(src, g["start"], g["end"], 1, Protein(g["protein_id"], None, doms.get(g["protein_id"], [])))
             for g in bgc["genes"]]
    golden = {g["protein_id"]: g["average_p"] for g in bgc["genes"]}
    crf = ClusterCRF.trained()
    assert crf.arithmetic == "f64"
    out = crf.predict_probabilities(genes)
    # the reference's own arithmetic on the device: python-crfsuite's numbers to the last bit
    assert [g.average_probability for g in out] == [golden[g.id] for g in out]
    crf.arithmetic = "f32"
    out = crf.predict_probabilities(genes)
    assert max(abs(g.average_probability - golden[g.id]) for g in out) <= 1e-5
